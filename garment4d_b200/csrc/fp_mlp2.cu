// Two-layer shared MLP over point-major rows on tcgen05, operands STREAMED by TMA (sm_100a).
//
// Replaces, for the coarser feature-propagation levels in eval mode (PointnetFPModule.forward, pointnet2_modules.py:138-156,
// widths pointnet2encoder.py:91-96: FP1 352 -> 256 -> 128 on 1024 points per cloud, FP2 576 -> 512 -> 256 on 256 points):
//     2 x [cuDNN 1x1 Conv2d -> BatchNorm2d -> ReLU]                                   (pytorch_utils.py:5-32)
// which round 1 ran as two library GEMMs plus two bias/ReLU passes of ours (the hidden activation went through HBM).
// Input: the activation rows x (b*n, c_in) fp16 written by g4d_fp_interp_concat_rows_h (weights + 3-tap interpolation + skip
// concat).  Output: (b, c2, n) fp32 channel-major (the reference layout) and, optionally, (b, n, c2) fp16 point-major for the
// next level's gather.
//
// The weights do not fit shared memory (FP2: 590 + 262 KB), so both GEMMs run a K-loop over STAGES (32 or 16 channels) of a ring:
//   stage = [A: 128 rows x ks k fp16, canonical K-major | W: N rows x ks k, canonical]
//   loader (1 lane)        per stage ONE TMA tensor copy for A -- box (ks channels, 128 rows) of the (rows, c_in) activation matrix
//                          with the hardware swizzle whose span is the box row (ks = 64: SWIZZLE_128B), read by the MMA through a
//                          K-major descriptor of the same swizzle mode; rows and channels past the end are zero-filled by the
//                          hardware -- and one cp.async.bulk for the host-packed W stage (canonical no-swizzle image), both
//                          completing on the stage's full barrier (expect_tx).  (A first version fetched A as a 3-D box with
//                          16-byte inner rows straight into the no-swizzle layout: 1024 row requests per 16 KB made the TMA unit
//                          the bottleneck of the whole kernel.)
//                          In layer 2 the A operand is the hidden activation H, resident in shared memory: W only.
//   MMA issuer (1 warp)    warp-uniform code, elected lane: per stage ks/16 k-steps x (N / 256 rounded up) tcgen05.mma, then
//                          tcgen05.commit -> the stage's empty barrier; D1 [128 x c1] and D2 [128 x c2] in TMEM, in SEPARATE
//                          columns when c1 + c2 <= 512 (then layer 1 of the next tile runs under epilogue 2 of this one)
//   epilogue (8 warps)     two per TMEM lane quadrant, splitting the columns: D1 -> +b1, ReLU, fp16 -> H (canonical layout),
//                          handed to the issuer in 64-channel pieces so layer 2 starts under the rest of epilogue 1;
//                          D2 -> +b2, ReLU -> fp32 channel-major (lane = point: 128-byte coalesced stores per channel)
//                          and fp16 point-major (64 contiguous bytes per thread and 32 channels)
// Persistent, one CTA per SM.
#include <stdlib.h>
#include <string.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "umma.cuh"
#include "garment4d_b200.h"

namespace g4d {

constexpr int M2_TILE = 128;
constexpr int M2_MAX_STAGES = 8;
constexpr int M2_MAX_PIECES = 16;
constexpr int M2_EPI_WARPS = 8;
constexpr int M2_THREADS = (M2_EPI_WARPS + 1 + 1) * 32;          // epilogue | issuer | loader = 320

struct Mlp2Layout {
    int c_in, c1, c2, ks, n1, n2, nst;                  // channels per stage; stages per PASS in layer 1 / 2; ring depth
    int npass, c1p;                                     // hidden channels are produced c1p (<= 256) at a time
    int piece, npieces;                                 // H hand-over granularity (channels), pieces per pass
    uint32_t a_bytes, w1_stage, w2_stage, stage_bytes;  // bytes of one A / W1 / W2 stage; ring slot size (A + max W)
    uint32_t off_w1, off_w2, off_b1, off_b2, blob_bytes;            // inside the blob (global memory)
    uint32_t off_bias, off_h, off_ring, off_bar, total_smem, tmem_cols, d2_col;
};

constexpr uint32_t M2_BAR_BYTES = 8u * (6 + 2 * M2_MAX_STAGES + M2_MAX_PIECES) + 16u;

static bool mlp2_layout(const g4d_mlp2_desc* d, Mlp2Layout* L, const char** why) {
    if (d->c_in < 32 || d->c_in % 32 || d->c_in > 4096) { *why = "mlp2: c_in must be a multiple of 32 in [32, 4096]"; return false; }
    if (d->c1 < 64 || d->c1 % 64 || d->c1 > 512 || (d->c1 > 256 && d->c1 % 128)) { *why = "mlp2: c1 must be a multiple of 64 in [64, 256] or of 128 up to 512"; return false; }
    if (d->c2 < 16 || d->c2 % 16 || d->c2 > 256) { *why = "mlp2: c2 must be a multiple of 16 in [16, 256]"; return false; }
    L->c_in = d->c_in; L->c1 = d->c1; L->c2 = d->c2;
    // Hidden channels are produced in PASSES of at most 256: D1 [128 x c1p] and D2 [128 x c2] then always fit TMEM side by side,
    // H is one pass wide (<= 64 KB), and layer 2 accumulates D2 over the passes (K = c1 split by pass).
    L->npass = d->c1 > 256 ? 2 : 1;
    L->c1p = d->c1 / L->npass;
    L->off_bias = 0;                                    // b1 | b2 (fp32) at the start of shared memory
    L->off_h = ((uint32_t)(d->c1 + d->c2) * 4 + 127) / 128 * 128;
    L->off_ring = (L->off_h + (uint32_t)M2_TILE * L->c1p * 2 + 1023u) / 1024u * 1024u;      // swizzled A stages: 1024-byte aligned
    const uint32_t budget = 227u * 1024u - 1024u - 512u;
    // Channels per stage: as many as leave three ring slots.  Every stage costs the issuer ~500 cycles of waits, fences and commits
    // whatever it holds, and a 128 x 256 x 16 MMA runs 128 cycles (tools/microbench/mma_rate.cu: the full 8192 FLOP/cycle/SM from this
    // layout, ring and concurrent bulk copies included): 64 channels = 4 MMAs per stage keep the tensor pipe, not the issuer, busy.
    static const int ks_env = getenv("G4D_MLP2_KS") ? atoi(getenv("G4D_MLP2_KS")) : 0;
    int ks = 0, nst = 0;
    for (int cand = 64; cand >= 16; cand >>= 1) {
        if (ks_env && cand != ks_env) continue;
        const uint32_t wmax = (uint32_t)(L->c1p > d->c2 ? L->c1p : d->c2) * cand * 2;
        const uint32_t a_bytes = (uint32_t)M2_TILE * cand * 2, stage = (a_bytes + wmax + 1023u) / 1024u * 1024u;
        if (L->off_ring + 2u * stage + M2_BAR_BYTES > budget) continue;
        const int n = (int)((budget - M2_BAR_BYTES - L->off_ring) / stage);
        if (n >= 3 || cand == 16 || ks_env) { ks = cand; nst = n; L->a_bytes = a_bytes; L->stage_bytes = stage; break; }
    }
    if (!ks) { *why = "mlp2: shared memory footprint exceeds 227 KB"; return false; }
    if (nst > M2_MAX_STAGES) nst = M2_MAX_STAGES;
    L->ks = ks; L->nst = nst;
    L->n1 = (d->c_in + ks - 1) / ks; L->n2 = L->c1p / ks;   // stages per PASS (the last layer-1 stage may be partly past c_in: zeros)
    L->w1_stage = (uint32_t)L->c1p * ks * 2; L->w2_stage = (uint32_t)d->c2 * ks * 2;
    uint32_t o = 0;
    L->off_w1 = o; o += L->w1_stage * (uint32_t)L->n1 * (uint32_t)L->npass;       // [pass][stage]
    L->off_w2 = o; o += L->w2_stage * (uint32_t)L->n2 * (uint32_t)L->npass;       // [pass][stage]
    L->off_b1 = o; o += (uint32_t)d->c1 * 4;
    L->off_b2 = o; o += (uint32_t)d->c2 * 4;
    L->blob_bytes = o;
    L->off_bar = L->off_ring + (uint32_t)nst * L->stage_bytes;
    L->total_smem = L->off_bar + M2_BAR_BYTES;
    L->piece = (L->c1p % 128 == 0) ? 64 : L->c1p / 2;   // each half of the columns (one epilogue warp per quadrant) = whole pieces
    L->npieces = L->c1p / L->piece;
    L->d2_col = (uint32_t)L->c1p;
    uint32_t p2 = 32;
    while (p2 < (uint32_t)(L->c1p + d->c2)) p2 <<= 1;
    L->tmem_cols = p2;
    return true;
}

struct Mlp2Args {
    Mlp2Layout L;
    long long rows;                  // b * n
    int n, ntiles;
    const unsigned char* blob;       // packed weights (device)
    float* out_cm;                   // (b, c2, n) fp32
    __half* out_pm;                  // (b, n, c2) fp16 or null
    int prof;                        // G4D_MLP2_PROF=1: CTA 0 accumulates role-level cycle counters (g4d_debug_mlp2_counters)
};

__device__ long long g_mlp2_prof[16];
#define M2_T0() (pf ? clock64() : 0ll)
#define M2_ACC(slot, t0) do { if (pf) acc[slot] += clock64() - (t0); } while (0)

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

// One box of the activation tensor map -> shared memory, completing (bytes) on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// K-major operand descriptor for a tile whose rows are ONE swizzle span wide (span = ks * 2 bytes: 128 / 64 / 32): 8-row groups
// 8 * span bytes apart (SBO), leading-dimension field 1 (unused inside a span), version 1, layout type 2 / 4 / 6
// (cute/arch/mma_sm100_desc.hpp: SWIZZLE_128B / 64B / 32B).  A k-step of 16 halves advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t desc_swizzled(uint32_t saddr, uint32_t span_bytes) {
    const uint32_t layout = span_bytes == 128 ? 2u : (span_bytes == 64 ? 4u : 6u);
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = ((8u * span_bytes) >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(M2_THREADS, 1)
fp_mlp2_kernel(const Mlp2Args a, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const Mlp2Layout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* b1 = reinterpret_cast<const float*>(smem + L.off_bias);
    const float* b2 = b1 + L.c1;
    // barriers: [0] biases, [1] tmem address, [2] d1_full, [3] d2_full, [4] epi1 (unused slot kept for alignment), [5] epi2,
    // then full[MAX_STAGES], empty[MAX_STAGES], h_ready[MAX_PIECES]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    const uint32_t bar_b = bar0, tmem_slot = bar0 + 8, bar_d1 = bar0 + 16, bar_d2 = bar0 + 24, bar_epi2 = bar0 + 40;
    const uint32_t bar_full = bar0 + 48, bar_empty = bar_full + 8 * M2_MAX_STAGES, bar_h = bar_empty + 8 * M2_MAX_STAGES;
    const uint32_t s_h = smem_u32(smem + L.off_h), s_ring = smem_u32(smem + L.off_ring);
    const int NST = L.nst, n1 = L.n1, n2 = L.n2, NP = L.npass, ks = L.ks;
    const bool pf = a.prof && blockIdx.x == 0;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_d1, 1);
        mbar_init(bar_d2, 1);
        mbar_init(bar_epi2, M2_EPI_WARPS);
        for (int s = 0; s < NST; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int j = 0; j < L.npieces; ++j) mbar_init(bar_h + 8 * j, 4);     // the four quadrant warps of the half that owns the piece
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
    // Stage sequence of a tile: for each pass, n1 layer-1 stages then n2 layer-2 stages; running stage number g, ring slot g % NST,
    // phase (g / NST) & 1.  One filler (the loader) and one drainer (the issuer) per slot, both in program order: the single phase
    // bit is safe.  d1_full and the H pieces complete once per pass (phase number u = q * NP + pass), d2_full and epi2 once per
    // tile, and their waiters wait for every phase.

    if (warp == M2_EPI_WARPS + 1) {
        // =========================== LOADER ==================================================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
            uint32_t g = 0;
            long long acc[1] = {0};
            const long long tl = M2_T0();
            for (int q = 0; q < nq; ++q) {
                const int row0 = (bx + q * gx) * M2_TILE;
                for (int p = 0; p < NP; ++p)
                    for (int st = 0; st < n1 + n2; ++st, ++g) {
                        const uint32_t slot = g % NST, ph = (g / NST) & 1;
                        { const long long t0 = M2_T0(); mbar_wait(bar_empty + 8 * slot, ph ^ 1); M2_ACC(0, t0); }
                        const uint32_t dst = s_ring + slot * L.stage_bytes, full = bar_full + 8 * slot;
                        if (st < n1) {
                            mbar_expect_tx(full, L.a_bytes + L.w1_stage);            // counts as this thread's arrival
                            tma_load_2d(dst, &xmap, st * ks, row0, full);
                            bulk_g2s(dst + L.a_bytes, a.blob + L.off_w1 + (size_t)(p * n1 + st) * L.w1_stage, L.w1_stage, full);
                        } else {
                            mbar_expect_tx(full, L.w2_stage);
                            bulk_g2s(dst + L.a_bytes, a.blob + L.off_w2 + (size_t)(p * n2 + st - n1) * L.w2_stage, L.w2_stage, full);
                        }
                    }
            }
            if (pf) { g_mlp2_prof[8] = acc[0]; g_mlp2_prof[9] = clock64() - tl; }
        }
    } else if (warp == M2_EPI_WARPS) {
        // =========================== MMA ISSUER (warp-uniform, elected lane) =================================
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem, 0);
        const uint32_t idesc1 = umma_idesc(M2_TILE, L.c1p), idesc2 = umma_idesc(M2_TILE, L.c2);
        const int ksteps = ks >> 4;
        uint32_t g = 0, u = 0;
        long long acc[5] = {0, 0, 0, 0, 0};
        const long long ti = M2_T0();
        for (int q = 0; q < nq; ++q) {
            for (int p = 0; p < NP; ++p, ++u) {
                // ---- layer 1 of this pass: D1 = X . W1p^T.  (D1 was drained by the previous pass's epilogue 1: every H piece
                // was waited for below.)
                for (int st = 0; st < n1; ++st, ++g) {
                    const uint32_t slot = g % NST, ph = (g / NST) & 1;
                    { const long long t0 = M2_T0(); mbar_wait(bar_full + 8 * slot, ph); M2_ACC(0, t0); }
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t sa = s_ring + slot * L.stage_bytes, sw = sa + L.a_bytes;
                        for (int k = 0; k < ksteps; ++k)
                            umma_f16(tmem_u, desc_swizzled(sa + k * 32, (uint32_t)ks * 2),
                                     desc64(desc_lo(sw + k * (2 * L.c1p * 16), L.c1p * 16)), idesc1, (st | k) > 0);
                        umma_commit(bar_empty + 8 * slot);
                        if (st == n1 - 1) umma_commit(bar_d1);
                    }
                    __syncwarp();
                }
                // ---- layer 2 of this pass: D2 (+)= Hp . W2p^T, started piece by piece as epilogue 1 hands H over.  The first
                // pass overwrites D2: epilogue 2 of the previous tile must have drained it.
                if (p == 0 && q > 0) { const long long t0 = M2_T0(); mbar_wait_spin(bar_epi2, (q - 1) & 1); tc_fence_after(); M2_ACC(3, t0); }
                int have = 0;
                for (int st = 0; st < n2; ++st, ++g) {
                    const int pc = ((st + 1) * ks - 1) / L.piece;            // last H piece this stage reads
                    if (pc >= have) {
                        const long long t0 = M2_T0();
                        for (; have <= pc; ++have) mbar_wait_spin(bar_h + 8 * have, u & 1);
                        tc_fence_after();
                        M2_ACC(2, t0);
                    }
                    const uint32_t slot = g % NST, ph = (g / NST) & 1;
                    { const long long t0 = M2_T0(); mbar_wait(bar_full + 8 * slot, ph); M2_ACC(1, t0); }
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t sw = s_ring + slot * L.stage_bytes + L.a_bytes;
                        for (int k = 0; k < ksteps; ++k)
                            umma_f16(tmem_u + L.d2_col, desc64(desc_lo(s_h + (uint32_t)(ksteps * st + k) * (2 * M2_TILE * 16), M2_TILE * 16)),
                                     desc64(desc_lo(sw + k * (2 * L.c2 * 16), L.c2 * 16)), idesc2, (p | st | k) > 0);
                        umma_commit(bar_empty + 8 * slot);
                        if (st == n2 - 1 && p == NP - 1) umma_commit(bar_d2);
                    }
                    __syncwarp();
                }
            }
        }
        if (pf && lane == 0) { for (int i = 0; i < 4; ++i) g_mlp2_prof[i] = acc[i]; g_mlp2_prof[4] = clock64() - ti; g_mlp2_prof[5] = nq; }
    } else {
        // =========================== EPILOGUE: warps q and q+4 own TMEM lanes 32q..32q+31, half the columns each ====
        mbar_wait(bar_b, 0);
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
        uint4* hd = reinterpret_cast<uint4*>(smem + L.off_h);
        long long acc[4] = {0, 0, 0, 0};
        uint32_t u = 0;
        for (int q = 0; q < nq; ++q) {
            const long long R = (long long)(bx + q * gx) * M2_TILE + row;
            const bool live = R < a.rows;
            // ---- epilogue 1, once per pass: D1 -> +b1, ReLU, fp16 -> H.  (d1_full of a pass is committed after the previous pass's
            // layer-2 MMAs, so H is no longer being read.)
            for (int p = 0; p < NP; ++p, ++u) {
                { const long long t0 = M2_T0(); mbar_wait(bar_d1, u & 1); M2_ACC(0, t0); }
                tc_fence_after();
                const long long te1 = M2_T0();
                const float* b1p = b1 + p * L.c1p;
                const int c_lo = half * (L.c1p / 2), c_hi = c_lo + L.c1p / 2;    // c1p / 2 is a multiple of 32
#pragma unroll 1
                for (int c = c_lo; c < c_hi; c += 32) {
                    uint32_t v[32];
                    tmem_ld32_m2(taddr + c, v);
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const float4 ba = *reinterpret_cast<const float4*>(b1p + c + 8 * w), bb = *reinterpret_cast<const float4*>(b1p + c + 8 * w + 4);
                        hd[(size_t)((c >> 3) + w) * M2_TILE + row] =
                            make_uint4(pack_relu_f16x2(__uint_as_float(v[8 * w]) + ba.x, __uint_as_float(v[8 * w + 1]) + ba.y),
                                       pack_relu_f16x2(__uint_as_float(v[8 * w + 2]) + ba.z, __uint_as_float(v[8 * w + 3]) + ba.w),
                                       pack_relu_f16x2(__uint_as_float(v[8 * w + 4]) + bb.x, __uint_as_float(v[8 * w + 5]) + bb.y),
                                       pack_relu_f16x2(__uint_as_float(v[8 * w + 6]) + bb.z, __uint_as_float(v[8 * w + 7]) + bb.w));
                    }
                    if ((c + 32) % L.piece == 0) {                                // a piece of H is complete in this warp
                        tc_fence_before();
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_h + 8 * (c / L.piece));
                    }
                }
                M2_ACC(1, te1);
            }
            // ---- epilogue 2: D2 -> +b2, ReLU -> fp32 channel-major + fp16 point-major
            { const long long t0 = M2_T0(); mbar_wait(bar_d2, q & 1); M2_ACC(2, t0); }
            tc_fence_after();
            const long long te2 = M2_T0();
            {
                const unsigned cloud = live ? (unsigned)((unsigned long long)R / (unsigned)a.n) : 0u;
                const int pt = live ? (int)(R - (long long)cloud * a.n) : 0;
                float* ocm = a.out_cm + ((size_t)cloud * L.c2) * a.n + pt;
                __half* opm = a.out_pm ? a.out_pm + (size_t)R * L.c2 : nullptr;
                const int c_lo = half * (L.c2 / 2), c_hi = c_lo + L.c2 / 2;      // c2 / 2 is a multiple of 8
#pragma unroll 1
                for (int c = c_lo; c < c_hi; c += 32) {
                    uint32_t v[32];
                    tmem_ld32_m2(taddr + L.d2_col + c, v);                    // (the last chunk of a short half reads columns past it: ignored)
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
            if (lane == 0) mbar_arrive(bar_epi2);
            M2_ACC(3, te2);
        }
        if (pf && tid == 0) for (int i = 0; i < 4; ++i) g_mlp2_prof[10 + i] = acc[i];
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, L.tmem_cols);
}

// The activation rows as a TMA tensor: (c_in, rows) fp16, box (ks channels, 128 rows), swizzle span = the box row (ks * 2 bytes).
static int make_x_map(CUtensorMap* map, const void* x, long long rows, int c_in, int ks) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) { set_error("mlp2_rows: cuTensorMapEncodeTiled not available from the driver"); return 1; }
        encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)c_in, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)c_in * 2};
    const cuuint32_t box[2] = {(cuuint32_t)ks, (cuuint32_t)M2_TILE};
    const CUtensorMapSwizzle sw = ks == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (ks == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("mlp2_rows: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return 1; }
    return 0;
}

}  // namespace g4d

using namespace g4d;

// Measurement aid: role-level cycle counters of CTA 0 of the last g4d_mlp2_rows launch made with G4D_MLP2_PROF=1 (16 int64):
// issuer [0] wait full (layer 1) [1] wait full (layer 2) [2] wait H [3] wait epilogue 2 [4] loop [5] tiles; loader [8] wait empty [9] loop;
// epilogue warp 0 [10] wait D1 [11] epilogue 1 [12] wait D2 [13] epilogue 2.
G4D_API int g4d_debug_mlp2_counters(long long* out16) { return (int)cudaMemcpyFromSymbol(out16, g_mlp2_prof, sizeof(long long) * 16); }

G4D_API size_t g4d_mlp2_param_bytes(const g4d_mlp2_desc* d) {
    Mlp2Layout L; const char* why = nullptr;
    if (!d || !mlp2_layout(d, &L, &why)) { set_error("%s", why ? why : "mlp2: null descriptor"); return 0; }
    return L.blob_bytes;
}

// w1 (c1, c_in), w2 (c2, c1) fp32 folded weights, b1 (c1), b2 (c2) -> blob (host memory): per stage (32 or 16 channels) the UMMA
// canonical K-major image of that K range ([k/8][row][k%8] fp16), then the fp32 biases.  Fails ("fp16 range") when a weight does
// not fit fp16.
G4D_API int g4d_mlp2_pack_params(const g4d_mlp2_desc* d, const float* w1, const float* b1, const float* w2, const float* b2, void* blob) {
    Mlp2Layout L; const char* why = nullptr;
    if (!d || !mlp2_layout(d, &L, &why)) return bad_arg(why ? why : "mlp2: null descriptor");
    if (!w1 || !b1 || !w2 || !b2 || !blob) return bad_arg("mlp2_pack_params: null pointer");
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, L.blob_bytes);
    bool ok = true;
    const int KS = L.ks;
    auto put = [&](__half* base, int R, int r, int kl, float v) {       // kl: k within the stage
        base[((size_t)(kl / 8) * R + r) * 8 + (kl % 8)] = __float2half_rn(v);
        ok &= fabsf(v) <= 65504.f;
    };
    for (int p = 0; p < L.npass; ++p) {
        for (int st = 0; st < L.n1; ++st) {
            __half* W = (__half*)(out + L.off_w1 + (size_t)(p * L.n1 + st) * L.w1_stage);
            for (int o = 0; o < L.c1p; ++o)
                for (int kl = 0; kl < KS; ++kl) put(W, L.c1p, o, kl, st * KS + kl < L.c_in ? w1[(size_t)(p * L.c1p + o) * L.c_in + st * KS + kl] : 0.f);
        }
        for (int st = 0; st < L.n2; ++st) {
            __half* W = (__half*)(out + L.off_w2 + (size_t)(p * L.n2 + st) * L.w2_stage);
            for (int o = 0; o < L.c2; ++o)
                for (int kl = 0; kl < KS; ++kl) put(W, L.c2, o, kl, w2[(size_t)o * L.c1 + p * L.c1p + st * KS + kl]);
        }
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
    a.blob = (const unsigned char*)params_dev;
    static const int prof_env = getenv("G4D_MLP2_PROF") ? atoi(getenv("G4D_MLP2_PROF")) : 0;
    a.prof = prof_env;
    CUtensorMap xmap;
    if (make_x_map(&xmap, x_h, a.rows, a.L.c_in, a.L.ks)) return 1;
    a.out_cm = out_cm; a.out_pm = (__half*)out_pm;
    cudaError_t e = cudaFuncSetAttribute(fp_mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total_smem);
    if (e != cudaSuccess) { set_error("mlp2_rows: shared memory opt-in (%u B): %s", a.L.total_smem, cudaGetErrorString(e)); return (int)e; }
    int grid = sm_count();
    if (grid > a.ntiles) grid = a.ntiles;
    fp_mlp2_kernel<<<grid, M2_THREADS, a.L.total_smem, (cudaStream_t)stream>>>(a, xmap);
    return finish_launch("g4d mlp2_rows");
}
