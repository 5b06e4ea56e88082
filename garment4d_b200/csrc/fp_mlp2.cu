// Two-layer shared MLP over point-major rows on tcgen05, weights STREAMED from L2 (sm_100a).
//
// Replaces, for the coarser feature-propagation levels in eval mode (PointnetFPModule.forward, pointnet2_modules.py:138-156,
// widths pointnet2encoder.py:91-96: FP1 352 -> 256 -> 128 on 1024 points per cloud, FP2 576 -> 512 -> 256 on 256 points):
//     2 x [cuDNN 1x1 Conv2d -> BatchNorm2d -> ReLU]                                   (pytorch_utils.py:5-32)
// which round 1 ran as two library GEMMs plus two bias/ReLU passes of ours (the hidden activation went through HBM).
// Input: the activation rows x (b*n, c_in) fp16 written by g4d_fp_interp_concat_rows_h (weights + 3-tap interpolation + skip
// concat).  Output: (b, c2, n) fp32 channel-major (the reference layout) and, optionally, (b, n, c2) fp16 point-major for the
// next level's gather.
//
// The weights do not fit shared memory (FP2: 590 + 262 KB), so both GEMMs run a K-loop over 32-channel STAGES of a ring:
//   stage = [A: 128 rows x 32 k fp16, canonical K-major, 8 KB | W: N rows x 32 k, canonical, N * 64 B]
//   A producers (4 warps)  one thread per tile row: 4 x cp.async of 16 B per stage (layer 1 only; in layer 2 the A operand is
//                          the hidden activation H, resident in shared memory)
//   W loader (1 lane)      one cp.async.bulk (TMA bulk copy) per stage from the host-packed blob, completing on the stage's
//                          full barrier (expect_tx)
//   MMA issuer (1 warp)    warp-uniform code, elected lane: per stage 2 k-steps x (N / 256 rounded up) tcgen05.mma, then
//                          tcgen05.commit -> the stage's empty barrier; D1 [128 x c1] and D2 [128 x c2] in TMEM (D2 reuses
//                          D1's columns: c1 can take all 512)
//   epilogue (8 warps)     two per TMEM lane quadrant, splitting the columns: D1 -> +b1, ReLU, fp16 -> H (canonical layout);
//                          D2 -> +b2, ReLU -> fp32 channel-major (lane = point: 128-byte coalesced stores per channel)
//                          and fp16 point-major (64 contiguous bytes per thread and 32 channels)
// One tile at a time per CTA (persistent, one CTA per SM): the tile's tensor work (FP2: 13 k cycles) dwarfs its hand-offs.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "umma.cuh"
#include "garment4d_b200.h"

namespace g4d {

constexpr int M2_TILE = 128;
constexpr int M2_KS = 32;                               // channels per stage
constexpr int M2_A_BYTES = M2_TILE * M2_KS * 2;        // 8 KB
constexpr int M2_MAX_STAGES = 4;
constexpr int M2_EPI_WARPS = 8;
constexpr int M2_THREADS = (M2_EPI_WARPS + 1 + 4 + 1) * 32;      // epilogue | issuer | A producers | W loader = 448

struct Mlp2Layout {
    int c_in, c1, c2, n1, n2, nst;                      // stages per tile in layer 1 / 2, ring depth
    uint32_t w1_stage, w2_stage, stage_bytes;           // bytes of one W1 / W2 stage; ring slot size (A + max W)
    uint32_t off_w1, off_w2, off_b1, off_b2, blob_bytes;            // inside the blob (global memory)
    uint32_t off_bias, off_h, off_ring, off_bar, total_smem, tmem_cols;
};

static bool mlp2_layout(const g4d_mlp2_desc* d, Mlp2Layout* L, const char** why) {
    if (d->c_in < 32 || d->c_in % 32 || d->c_in > 4096) { *why = "mlp2: c_in must be a multiple of 32 in [32, 4096]"; return false; }
    if (d->c1 < 64 || d->c1 % 32 || d->c1 > 512) { *why = "mlp2: c1 must be a multiple of 32 in [64, 512]"; return false; }
    if (d->c2 < 16 || d->c2 % 16 || d->c2 > 256) { *why = "mlp2: c2 must be a multiple of 16 in [16, 256]"; return false; }
    L->c_in = d->c_in; L->c1 = d->c1; L->c2 = d->c2;
    L->n1 = d->c_in / M2_KS; L->n2 = d->c1 / M2_KS;
    L->w1_stage = (uint32_t)d->c1 * M2_KS * 2; L->w2_stage = (uint32_t)d->c2 * M2_KS * 2;
    uint32_t o = 0;
    L->off_w1 = o; o += L->w1_stage * (uint32_t)L->n1;
    L->off_w2 = o; o += L->w2_stage * (uint32_t)L->n2;
    L->off_b1 = o; o += (uint32_t)d->c1 * 4;
    L->off_b2 = o; o += (uint32_t)d->c2 * 4;
    L->blob_bytes = o;
    L->stage_bytes = M2_A_BYTES + (L->w1_stage > L->w2_stage ? L->w1_stage : L->w2_stage);
    L->off_bias = 0;                                    // b1 | b2 (fp32) at the start of shared memory
    L->off_h = ((uint32_t)(d->c1 + d->c2) * 4 + 127) / 128 * 128;
    L->off_ring = L->off_h + (uint32_t)M2_TILE * d->c1 * 2;
    const uint32_t budget = 227u * 1024u - 1024u - 512u;
    const uint32_t bar_bytes = 8u * (4 + 2 * M2_MAX_STAGES) + 16u;
    if (L->off_ring + 2u * L->stage_bytes + bar_bytes > budget) { *why = "mlp2: shared memory footprint exceeds 227 KB"; return false; }   // two stages at least
    int nst = (int)((budget - bar_bytes - L->off_ring) / L->stage_bytes);
    if (nst > M2_MAX_STAGES) nst = M2_MAX_STAGES;
    L->nst = nst;
    L->off_bar = L->off_ring + (uint32_t)nst * L->stage_bytes;
    L->total_smem = L->off_bar + bar_bytes;
    uint32_t p2 = 32;
    while (p2 < (uint32_t)d->c1 || p2 < (uint32_t)d->c2) p2 <<= 1;
    L->tmem_cols = p2;
    return true;
}

struct Mlp2Args {
    Mlp2Layout L;
    long long rows;                  // b * n
    int n, ntiles;
    const __half* x;                 // (rows, c_in) fp16 row-major
    const unsigned char* blob;       // packed weights (device)
    float* out_cm;                   // (b, c2, n) fp32
    __half* out_pm;                  // (b, n, c2) fp16 or null
};

__device__ __forceinline__ void tmem_ld32_m2(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
}

__global__ void __launch_bounds__(M2_THREADS, 1)
fp_mlp2_kernel(const Mlp2Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Mlp2Layout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* b1 = reinterpret_cast<const float*>(smem + L.off_bias);
    const float* b2 = b1 + L.c1;
    // barriers: [0] biases, [1] tmem address, [2] d_full, [3] epi_done, then full[nst], empty[nst]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    const uint32_t bar_b = bar0, tmem_slot = bar0 + 8, bar_dfull = bar0 + 16, bar_epi = bar0 + 24;
    const uint32_t bar_full = bar0 + 32, bar_empty = bar_full + 8 * M2_MAX_STAGES;
    const uint32_t s_h = smem_u32(smem + L.off_h), s_ring = smem_u32(smem + L.off_ring);
    const int NST = L.nst, n1 = L.n1, n2 = L.n2, per_tile = n1 + n2;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_dfull, 1);
        mbar_init(bar_epi, M2_EPI_WARPS);
        for (int s = 0; s < NST; ++s) { mbar_init(bar_full + 8 * s, 5);   /* 4 A-producer warps + the W loader */ mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, L.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L.off_bar + 8);
    if (tid == 0) {
        mbar_expect_tx(bar_b, (uint32_t)(L.c1 + L.c2) * 4);
        bulk_g2s(smem_u32(smem + L.off_bias), a.blob + L.off_b1, (uint32_t)(L.c1 + L.c2) * 4, bar_b);
    }
    const int bx = (int)blockIdx.x, gx = (int)gridDim.x;
    const int nq = bx < a.ntiles ? (a.ntiles - bx + gx - 1) / gx : 0;          // tiles of this CTA: bx + q * gx
    // stage number g = q * per_tile + (stage within the tile); ring slot g % NST, phase (g / NST) & 1.  One filler side (the
    // producers and the loader, each in program order) and one drainer (the issuer) per slot: the single phase bit is safe.

    if (warp == M2_EPI_WARPS + 5) {
        // =========================== W LOADER ================================================================
        if (lane == 0) {
            uint32_t g = 0;
            for (int q = 0; q < nq; ++q)
                for (int st = 0; st < per_tile; ++st, ++g) {
                    const uint32_t slot = g % NST, ph = (g / NST) & 1;
                    mbar_wait(bar_empty + 8 * slot, ph ^ 1);       // (suspending wait: a nanosleep poll costs >= 256 ns per miss, and with 2-4 stages nearly every wait misses)
                    const bool l1 = st < n1;
                    const uint32_t bytes = l1 ? L.w1_stage : L.w2_stage;
                    const unsigned char* src = a.blob + (l1 ? L.off_w1 + (size_t)st * L.w1_stage : L.off_w2 + (size_t)(st - n1) * L.w2_stage);
                    mbar_expect_tx(bar_full + 8 * slot, bytes);                  // counts as this thread's arrival
                    bulk_g2s(s_ring + slot * L.stage_bytes + M2_A_BYTES, src, bytes, bar_full + 8 * slot);
                }
        }
    } else if (warp >= M2_EPI_WARPS + 1) {
        // =========================== A PRODUCERS: one thread per tile row ====================================
        // Software pipeline: the copies of up to LAG + 1 stages are in flight; a stage is handed over (arrive on its full barrier)
        // LAG stages after it was issued.  LAG <= NST - 1, so the empty slot a new stage waits for was handed over long before.
        const int r = (warp - (M2_EPI_WARPS + 1)) * 32 + lane;
        const int LAG = NST - 1 < 3 ? NST - 1 : 3;
        const uint32_t total = (uint32_t)nq * (uint32_t)per_tile;
        uint32_t g = 0;
        for (int q = 0; q < nq; ++q) {
            const long long R = (long long)(bx + q * gx) * M2_TILE + r;
            const bool live = R < a.rows;
            const char* srow = reinterpret_cast<const char*>(a.x + (size_t)(live ? R : 0) * L.c_in);
            for (int st = 0; st < per_tile; ++st, ++g) {
                const uint32_t slot = g % NST, ph = (g / NST) & 1;
                mbar_wait(bar_empty + 8 * slot, ph ^ 1);       // (suspending wait: a nanosleep poll costs >= 256 ns per miss, and with 2-4 stages nearly every wait misses)
                if (st < n1) {
                    const uint32_t sdst = s_ring + slot * L.stage_bytes + (uint32_t)r * 16;
#pragma unroll
                    for (int c = 0; c < 4; ++c)                                  // 4 chunks of 8 channels: [(k/8)][row][8]
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst + c * (M2_TILE * 16)),
                                     "l"(srow + (size_t)st * (M2_KS * 2) + c * 16), "r"(live ? 16 : 0) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");              // (an empty group for layer-2 stages: A = H)
                if (g >= (uint32_t)LAG) {
                    switch (LAG) {
                        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                        default: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * ((g - LAG) % NST));
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
        for (uint32_t t = total > (uint32_t)LAG ? total - LAG : 0; t < total; ++t)
            if (lane == 0) mbar_arrive(bar_full + 8 * (t % NST));
    } else if (warp == M2_EPI_WARPS) {
        // =========================== MMA ISSUER (warp-uniform, elected lane) =================================
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem, 0);
        const uint32_t idesc1 = umma_idesc(M2_TILE, L.c1 > 256 ? 256 : L.c1), idesc1b = umma_idesc(M2_TILE, L.c1 > 256 ? L.c1 - 256 : 16);
        const uint32_t idesc2 = umma_idesc(M2_TILE, L.c2);
        uint32_t g = 0, nepi = 0;
        for (int q = 0; q < nq; ++q) {
            // ---- layer 1: D1 = X . W1^T   (needs D drained by the previous tile's epilogue 2)
            mbar_wait_spin(bar_epi, (nepi + 1) & 1); ++nepi;
            tc_fence_after();
            for (int st = 0; st < n1; ++st, ++g) {
                const uint32_t slot = g % NST, ph = (g / NST) & 1;
                mbar_wait(bar_full + 8 * slot, ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t sa = s_ring + slot * L.stage_bytes, sw = sa + M2_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint64_t ad = desc64(desc_lo(sa + k * (2 * M2_TILE * 16), M2_TILE * 16));
                        umma_f16(tmem_u, ad, desc64(desc_lo(sw + k * (2 * L.c1 * 16), L.c1 * 16)), idesc1, (st | k) > 0);
                        if (L.c1 > 256)
                            umma_f16(tmem_u + 256, ad, desc64(desc_lo(sw + k * (2 * L.c1 * 16) + 256 * 16, L.c1 * 16)), idesc1b, (st | k) > 0);
                    }
                    umma_commit(bar_empty + 8 * slot);
                    if (st == n1 - 1) umma_commit(bar_dfull);
                }
                __syncwarp();
            }
            // ---- layer 2: D2 = H . W2^T   (needs H written by epilogue 1)
            mbar_wait_spin(bar_epi, (nepi + 1) & 1); ++nepi;
            tc_fence_after();
            for (int st = 0; st < n2; ++st, ++g) {
                const uint32_t slot = g % NST, ph = (g / NST) & 1;
                mbar_wait(bar_full + 8 * slot, ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t sw = s_ring + slot * L.stage_bytes + M2_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_f16(tmem_u, desc64(desc_lo(s_h + (uint32_t)(2 * st + k) * (2 * M2_TILE * 16), M2_TILE * 16)),
                                 desc64(desc_lo(sw + k * (2 * L.c2 * 16), L.c2 * 16)), idesc2, (st | k) > 0);
                    umma_commit(bar_empty + 8 * slot);
                    if (st == n2 - 1) umma_commit(bar_dfull);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== EPILOGUE: warps q and q+4 own TMEM lanes 32q..32q+31, half the columns each ====
        mbar_wait(bar_b, 0);
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
        uint4* hd = reinterpret_cast<uint4*>(smem + L.off_h);
        uint32_t nd = 0;
        for (int q = 0; q < nq; ++q) {
            const long long R = (long long)(bx + q * gx) * M2_TILE + row;
            const bool live = R < a.rows;
            // ---- epilogue 1: D1 -> +b1, ReLU, fp16 -> H
            mbar_wait(bar_dfull, nd & 1); ++nd;
            tc_fence_after();
            {
                const int c_lo = half * (L.c1 / 2), c_hi = c_lo + L.c1 / 2;      // c1 / 2 is a multiple of 16
#pragma unroll 1
                for (int c = c_lo; c < c_hi; c += 32) {
                    uint32_t v[32];
                    tmem_ld32_m2(taddr + c, v);                               // (the last chunk of a 16-multiple half reads 16 columns too many: ignored)
                    const int nu = (c_hi - c) >= 32 ? 4 : 2;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (u < nu) {
                            const float4 ba = *reinterpret_cast<const float4*>(b1 + c + 8 * u), bb = *reinterpret_cast<const float4*>(b1 + c + 8 * u + 4);
                            hd[(size_t)((c >> 3) + u) * M2_TILE + row] =
                                make_uint4(pack_relu_f16x2(__uint_as_float(v[8 * u]) + ba.x, __uint_as_float(v[8 * u + 1]) + ba.y),
                                           pack_relu_f16x2(__uint_as_float(v[8 * u + 2]) + ba.z, __uint_as_float(v[8 * u + 3]) + ba.w),
                                           pack_relu_f16x2(__uint_as_float(v[8 * u + 4]) + bb.x, __uint_as_float(v[8 * u + 5]) + bb.y),
                                           pack_relu_f16x2(__uint_as_float(v[8 * u + 6]) + bb.z, __uint_as_float(v[8 * u + 7]) + bb.w));
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_epi);
            // ---- epilogue 2: D2 -> +b2, ReLU -> fp32 channel-major + fp16 point-major
            mbar_wait(bar_dfull, nd & 1); ++nd;
            tc_fence_after();
            {
                const unsigned cloud = live ? (unsigned)((unsigned long long)R / (unsigned)a.n) : 0u;
                const int pt = live ? (int)(R - (long long)cloud * a.n) : 0;
                float* ocm = a.out_cm + ((size_t)cloud * L.c2) * a.n + pt;
                __half* opm = a.out_pm ? a.out_pm + (size_t)R * L.c2 : nullptr;
                const int c_lo = half * (L.c2 / 2), c_hi = c_lo + L.c2 / 2;      // c2 / 2 is a multiple of 8
#pragma unroll 1
                for (int c = c_lo; c < c_hi; c += 32) {
                    uint32_t v[32];
                    tmem_ld32_m2(taddr + c, v);
                    const int nv = (c_hi - c) >= 32 ? 32 : (c_hi - c);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (8 * u < nv) {
                            float o[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) o[i] = fmaxf(__uint_as_float(v[8 * u + i]) + b2[c + 8 * u + i], 0.f);
                            if (live) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) ocm[(size_t)(c + 8 * u + i) * a.n] = o[i];     // lane = point: coalesced per channel
                                if (opm)
                                    *reinterpret_cast<uint4*>(opm + c + 8 * u) =
                                        make_uint4(pack_f16x2(o[0], o[1]), pack_f16x2(o[2], o[3]), pack_f16x2(o[4], o[5]), pack_f16x2(o[6], o[7]));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_epi);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, L.tmem_cols);
}

}  // namespace g4d

using namespace g4d;

G4D_API size_t g4d_mlp2_param_bytes(const g4d_mlp2_desc* d) {
    Mlp2Layout L; const char* why = nullptr;
    if (!d || !mlp2_layout(d, &L, &why)) { set_error("%s", why ? why : "mlp2: null descriptor"); return 0; }
    return L.blob_bytes;
}

// w1 (c1, c_in), w2 (c2, c1) fp32 folded weights, b1 (c1), b2 (c2) -> blob (host memory): per 32-channel stage the UMMA
// canonical K-major image of that K range ([k/8][row][k%8] fp16), then the fp32 biases.  Fails ("fp16 range") when a weight does
// not fit fp16.
G4D_API int g4d_mlp2_pack_params(const g4d_mlp2_desc* d, const float* w1, const float* b1, const float* w2, const float* b2, void* blob) {
    Mlp2Layout L; const char* why = nullptr;
    if (!d || !mlp2_layout(d, &L, &why)) return bad_arg(why ? why : "mlp2: null descriptor");
    if (!w1 || !b1 || !w2 || !b2 || !blob) return bad_arg("mlp2_pack_params: null pointer");
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, L.blob_bytes);
    bool ok = true;
    auto put = [&](__half* base, int R, int r, int kl, float v) {       // kl: k within the stage (0..31)
        base[((size_t)(kl / 8) * R + r) * 8 + (kl % 8)] = __float2half_rn(v);
        ok &= fabsf(v) <= 65504.f;
    };
    for (int st = 0; st < L.n1; ++st) {
        __half* W = (__half*)(out + L.off_w1 + (size_t)st * L.w1_stage);
        for (int o = 0; o < L.c1; ++o)
            for (int kl = 0; kl < M2_KS; ++kl) put(W, L.c1, o, kl, w1[(size_t)o * L.c_in + st * M2_KS + kl]);
    }
    for (int st = 0; st < L.n2; ++st) {
        __half* W = (__half*)(out + L.off_w2 + (size_t)st * L.w2_stage);
        for (int o = 0; o < L.c2; ++o)
            for (int kl = 0; kl < M2_KS; ++kl) put(W, L.c2, o, kl, w2[(size_t)o * L.c1 + st * M2_KS + kl]);
    }
    memcpy(out + L.off_b1, b1, sizeof(float) * L.c1);
    memcpy(out + L.off_b2, b2, sizeof(float) * L.c2);
    if (!ok) return bad_arg("mlp2_pack_params: a folded weight is outside the fp16 range (|v| > 65504)");
    return 0;
}

// x (b*n, c_in) fp16 row-major -> out_cm (b, c2, n) fp32 = relu(W2 relu(W1 x + b1) + b2), out_pm (b, n, c2) fp16 (may be NULL).
G4D_API int g4d_mlp2_rows(const g4d_mlp2_desc* d, const void* params_dev, int b, int n, const void* x_h, float* out_cm, void* out_pm,
                          void* stream) {
    Mlp2Args a;
    const char* why = nullptr;
    if (!d || !mlp2_layout(d, &a.L, &why)) return bad_arg(why ? why : "mlp2: null descriptor");
    if (b < 0 || n < 0) return bad_arg("mlp2_rows: negative size");
    if (b == 0 || n == 0) return 0;
    if (!params_dev || !x_h || !out_cm) return bad_arg("mlp2_rows: null pointer");
    if (((uintptr_t)params_dev & 15) || ((uintptr_t)x_h & 15) || ((uintptr_t)out_pm & 15)) return bad_arg("mlp2_rows: params / x / out_pm must be 16-byte aligned");
    a.rows = (long long)b * n;
    if (a.rows > 0x7FFFFF00ll) return bad_arg("mlp2_rows: b*n must stay below 2^31");
    a.n = n;
    a.ntiles = (int)((a.rows + M2_TILE - 1) / M2_TILE);
    a.x = (const __half*)x_h; a.blob = (const unsigned char*)params_dev;
    a.out_cm = out_cm; a.out_pm = (__half*)out_pm;
    cudaError_t e = cudaFuncSetAttribute(fp_mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total_smem);
    if (e != cudaSuccess) { set_error("mlp2_rows: shared memory opt-in (%u B): %s", a.L.total_smem, cudaGetErrorString(e)); return (int)e; }
    int grid = sm_count();
    if (grid > a.ntiles) grid = a.ntiles;
    fp_mlp2_kernel<<<grid, M2_THREADS, a.L.total_smem, (cudaStream_t)stream>>>(a);
    return finish_launch("g4d mlp2_rows");
}
