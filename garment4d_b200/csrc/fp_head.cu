// Fused feature propagation + segmentation head on tcgen05 (sm_100a).
//
// Replaces, for a PointnetFPModule without skip features in eval mode followed (optionally) by the encoder's
// FC head (pointnet2_modules.py:131-156 as used by pointnet2encoder.py:138-143 for the finest level):
//     dist_recip = 1/(dist+1e-8); weight = dist_recip / sum  (3 torch kernels)            pointnet2_modules.py:141-143
//     three_interpolate                                                                     interpolate_gpu.cu:77-97
//     unsqueeze ; 2 x [cuDNN conv1x1 -> BatchNorm2d -> ReLU] ; squeeze                      pytorch_utils.py:5-32
//     Conv1d(64,32)+BN+ReLU ; Dropout (eval: identity) ; Conv1d(32,classes) ; transpose     pointnet2encoder.py:98-101,143
// (~16 kernels, each streaming a (B, C, N) activation of up to 1 GB through HBM at c3) with ONE persistent,
// warp-specialised kernel (one CTA per SM, 16 warps):
//
//   producer groups (2 x 4 warps)  take alternate 128-point tiles: inverse-distance weights from three_nn's squared
//                                  distances (meta loads prefetched one tile ahead), then the 3-tap interpolation is
//                                  evaluated while gathering the fp16 point-major rows (24 x 16-byte loads in flight per
//                                  lane; fp32 arithmetic in the reference's FMUL/FFMA order) and written as fp16 in the
//                                  UMMA canonical layout into a ring of A buffers (full/empty mbarriers).
//   consumer groups (2 x 4 warps)  take alternate tiles, each with its own TMEM accumulator (128 columns) and hidden
//                                  buffer: one lane issues the tcgen05.mma chain (up to four layers), the group's four
//                                  warps run the epilogues (tcgen05.ld -> bias, ReLU -> fp16 -> shared; layer 2 also
//                                  writes the FP output, the last layer the logits).  The hidden activations overwrite the
//                                  tile's A buffer in place; the tcgen05.commit of the last layer returns it to the ring, so
//                                  with 4 buffers the gather of the next two tiles overlaps the two chains in flight.
//   Only the FP output (channel-major fp32, returned by the encoder as l_features[0]) and the logits (point-major)
//   are written to HBM.  (The first version ran gather and layers back to back in one 256-thread CTA with block-wide
//   barriers: 31 % of its warp samples waited on the gather, 31 % at barriers -- profiles/r01_ncu_summary.md.)
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "umma.cuh"
#include "garment4d_b200.h"

namespace g4d {

constexpr int FP_TILE = 128;
constexpr int FP_CG = 2, FP_PG = 2;                   // consumer / producer groups of 4 warps each
constexpr int FP_THREADS = (FP_CG + FP_PG) * 128;     // warps [0, 4*FP_CG): consumers; the rest: producers
constexpr int FP_MAX_A = 4;                           // A-buffer ring depth (as many as fit, at least 2)

struct FpLayout {
    int c_in, c1, c2, h1, h2, h2p;      // h2p = 16 when a head is present (classes padded), else 0
    uint32_t off_w1, off_w2, off_w3, off_w4, off_b1, off_b2, off_b3, off_b4, blob_bytes;
    uint32_t off_a, a_bytes, na, off_h, h_bytes, off_meta, off_bar, total_smem, tcols, tmem_cols;
};

static bool fp_layout(const g4d_fp_desc* d, FpLayout* L, const char** why) {
    auto ok16 = [](int v) { return v >= 16 && v <= 256 && v % 16 == 0; };
    if (!ok16(d->c_in) || !ok16(d->c1) || !ok16(d->c2)) { *why = "fp: c_in, c1, c2 must be multiples of 16 in [16, 256]"; return false; }
    if (d->h1 != 0 && (!ok16(d->h1) || d->h2 < 1 || d->h2 > 16)) { *why = "fp: head needs h1 multiple of 16 in [16,256] and 1 <= h2 <= 16"; return false; }
    L->c_in = d->c_in; L->c1 = d->c1; L->c2 = d->c2; L->h1 = d->h1; L->h2 = d->h1 ? d->h2 : 0; L->h2p = d->h1 ? 16 : 0;
    uint32_t o = 0;
    L->off_w1 = o; o += (uint32_t)L->c_in * L->c1 * 2;
    L->off_w2 = o; o += (uint32_t)L->c1 * L->c2 * 2;
    L->off_w3 = o; o += (uint32_t)L->c2 * L->h1 * 2;
    L->off_w4 = o; o += (uint32_t)L->h1 * L->h2p * 2;
    L->off_b1 = o; o += (uint32_t)L->c1 * 4;
    L->off_b2 = o; o += (uint32_t)L->c2 * 4;
    L->off_b3 = o; o += (uint32_t)L->h1 * 4;
    L->off_b4 = o; o += (uint32_t)L->h2p * 4;
    L->blob_bytes = o;
    int hmax = L->c1;                                      // widest hidden activation (K of layers 2..4)
    if (L->c2 > hmax) hmax = L->c2;
    if (L->h1 > hmax) hmax = L->h1;
    // One ring of buffers serves as layer-1 operand (filled by the producers) AND, in place, as the hidden-activation buffer
    // of the consumer group that took it: it goes back to the producers when the tile's last MMA has read it.
    const int kmax = L->c_in > hmax ? L->c_in : hmax;
    L->a_bytes = (uint32_t)FP_TILE * kmax * 2;
    L->h_bytes = 0;
    L->off_h = 0;
    L->off_meta = (o + 127) / 128 * 128;                   // per producer group, double-buffered: 3 point ids + 3 weights per row
    L->off_bar = L->off_meta + FP_PG * 3 * FP_TILE * 6 * 4;  // + one raw staging area (cp.async prefetch of dist2 / idx) per group
    L->off_a = (L->off_bar + 8 * (3 + 2 * FP_MAX_A + FP_CG) + 127) / 128 * 128;
    // keep ~24 KB of the SM's shared memory free: Nsight Compute cannot replay a kernel that takes all 227 KB (the first
    // version of this layout did, and every capture of it hung), and the L1 that is left serves the gather
    const uint32_t budget = 203u * 1024u;
    if (L->off_a + 2 * L->a_bytes > 227u * 1024u) { *why = "fp: shared memory footprint exceeds 227 KB"; return false; }
    uint32_t na = budget > L->off_a ? (budget - L->off_a) / L->a_bytes : 0;
    // The ring depth must be EVEN: tile i lives in buffer i % na and is filled by producer group i % 2 and drained by consumer
    // group i % 2, so with an even depth every buffer has ONE filler and ONE drainer, each seeing its phases in order -- the
    // single phase bit of an mbarrier is only safe then (with 3 buffers a producer group could run a lap ahead of a lagging
    // consumer group and pass its "empty" wait on the wrong phase).  4 when they fit (one tile of run-ahead per group), else 2.
    (void)budget;
    na = (L->off_a + 4 * L->a_bytes <= 227u * 1024u) ? 4 : 2;     // (the encoder's shapes need 200 KB with 4: inside the profiling budget)
    L->na = na;
    L->total_smem = L->off_a + L->na * L->a_bytes;
    uint32_t p2 = 32;
    while (p2 < (uint32_t)hmax || p2 < (uint32_t)L->h2p) p2 <<= 1;
    L->tcols = p2;                                         // accumulator columns per consumer group
    L->tmem_cols = p2 * FP_CG;                             // <= 512
    return true;
}

struct FpArgs {
    FpLayout L;
    int n, m;
    long long total_rows;
    int ntiles;
    const float* dist2;          // (b, n, 3) squared distances from three_nn
    const int* idx;              // (b, n, 3)
    const __half* known_pm;      // (b, m, c_in) point-major fp16
    const unsigned char* params;
    float* out_feat;             // (b, c2, n) channel-major fp32
    float* out_head;             // (b, n, h2) or null
    unsigned char* out_label;    // (b, n) arg-max over the h2 head outputs, or null
    long long* dbg;              // optional cycle counters of CTA 0 (g4d_debug_fp_counters)
};

// one 256-bit read-only load (LDG.E.256, sm_100+): 32 bytes per lane, p 32-byte aligned
__device__ __forceinline__ void ldg256(const uint4* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p));
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// TMEM row -> +bias, ReLU, fp16 -> hidden buffer (canonical layout), optionally also fp32 channel-major to global.
// NB x 16 columns per step: all tcgen05.ld of a step are in flight before the single wait (the load -> wait -> use chain is
// the consumer's critical path; every access also queues behind the producers' gather in the LSU data pipe).
template <int NB>
__device__ __forceinline__ void fp_epilogue_step(uint32_t taddr, const float* bias, unsigned char* hbuf, int row, int c0,
                                                 float* gout, size_t gstride) {
    float v[16 * NB], bv[16 * NB];
#pragma unroll
    for (int i = 0; i < 4 * NB; ++i)                            // broadcast LDS.128: 4 per 16 columns instead of 16 LDS.32
        *reinterpret_cast<float4*>(bv + 4 * i) = *reinterpret_cast<const float4*>(bias + c0 + 4 * i);
    if (NB == 2) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v);
#pragma unroll
    for (int i = 0; i < 16 * NB; ++i) v[i] = fmaxf(v[i] + bv[i], 0.f);
    if (gout) {
#pragma unroll
        for (int i = 0; i < 16 * NB; ++i) gout[(size_t)(c0 + i) * gstride] = v[i];     // lanes = consecutive points: coalesced
    }
    uint4* dst = reinterpret_cast<uint4*>(hbuf);
#pragma unroll
    for (int q = 0; q < 2 * NB; ++q)
        dst[(size_t)(c0 / 8 + q) * FP_TILE + row] = make_uint4(pack_f16x2(v[8 * q], v[8 * q + 1]), pack_f16x2(v[8 * q + 2], v[8 * q + 3]),
                                                                 pack_f16x2(v[8 * q + 4], v[8 * q + 5]), pack_f16x2(v[8 * q + 6], v[8 * q + 7]));
}

__device__ __forceinline__ void fp_epilogue_relu(uint32_t lane_taddr, int ncols, const float* bias, unsigned char* hbuf, int row,
                                                 float* gout /* channel 0 of this row's point, or null */, size_t gstride) {
    int c0 = 0;
#pragma unroll 1
    for (; c0 + 32 <= ncols; c0 += 32) fp_epilogue_step<2>(lane_taddr, bias, hbuf, row, c0, gout, gstride);
    if (c0 < ncols) fp_epilogue_step<1>(lane_taddr, bias, hbuf, row, c0, gout, gstride);
}

// D[128 x N] (+)= A[128 x K] . W[N x K]^T, one tcgen05.mma per 16-wide K step (issued by ONE thread; descriptors advance
// by plain 32-bit adds on their low words: every instruction on this lane's path is latency of the whole chain)
__device__ __forceinline__ void fp_issue_layer(uint32_t tmem, uint32_t s_act, uint32_t s_w, int K, int N) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc(FP_TILE, N);
    uint32_t alo = desc_lo(s_act, FP_TILE * 16), blo = desc_lo(s_w, (uint32_t)N * 16);
    const uint32_t a_step = (uint32_t)(2 * FP_TILE * 16) >> 4, b_step = (uint32_t)(2 * N * 16) >> 4;
    const int ks = K >> 4;
#pragma unroll 1
    for (int k = 0; k < ks; ++k) {
        umma_f16(tmem, desc64(alo), desc64(blo), idesc, k > 0);
        alo += a_step; blo += b_step;
    }
}

// Barrier of the 128 threads of one consumer / producer group.  Immediate barrier ids on purpose: with a register operand
// ptxas reserves all 16 hardware barriers of the CTA (and Nsight Compute could not replay the kernel).
__device__ __forceinline__ void group_bar(int id) {
    switch (id) {
        case 1: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
        case 3: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
    }
}

__global__ void __launch_bounds__(FP_THREADS, 1)
fp_interp_mlp_kernel(const FpArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const FpLayout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* b1 = reinterpret_cast<const float*>(smem + L.off_b1);
    const float* b2 = reinterpret_cast<const float*>(smem + L.off_b2);
    const float* b3 = reinterpret_cast<const float*>(smem + L.off_b3);
    const float* b4 = reinterpret_cast<const float*>(smem + L.off_b4);
    // barriers: [0] weights, [1] tmem slot, [2 .. 2+MAX_A) full[a], [2+MAX_A .. 2+2*MAX_A) empty[a], then mma_done[group]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    const uint32_t bar_w = bar0, tmem_slot = bar0 + 8, bar_full = bar0 + 16, bar_empty = bar_full + 8 * FP_MAX_A,
                   bar_done = bar_empty + 8 * FP_MAX_A;
    const int NA = (int)L.na;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < NA; ++i) { mbar_init(bar_full + 8 * i, 4); /* the 4 warps of the filling producer group */ mbar_init(bar_empty + 8 * i, 1); }
        for (int g = 0; g < FP_CG; ++g) mbar_init(bar_done + 8 * g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, L.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L.off_bar + 8);
    if (tid == 0) {
        mbar_expect_tx(bar_w, L.blob_bytes);
        bulk_g2s(smem_u32(smem), a.params, L.blob_bytes, bar_w);
    }
    const int nseq = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles of this CTA: blockIdx.x + i * gridDim.x

    if (warp >= 4 * FP_CG) {
        // =========================== PRODUCERS ==================================================================
        const int pg = (warp - 4 * FP_CG) >> 2;                 // producer group
        const int pw = (warp - 4 * FP_CG) & 3;                  // warp within the group: tile rows 32*pw .. 32*pw+31
        const int prow = pw * 32 + lane;                         // the row whose meta data this thread prepares
        const int rl = lane & 7, cl = lane >> 3;
        const int nchunk = L.c_in >> 3;
        uint32_t* meta = reinterpret_cast<uint32_t*>(smem + L.off_meta) + pg * 3 * (FP_TILE * 6);
        // Raw dist2 / idx of the row this thread prepares are prefetched one tile ahead with cp.async into the thread's own
        // six staging words (no registers held across the gather -- they spilled, and the spill store waited for the load).
        float* raw = reinterpret_cast<float*>(meta + 2 * (FP_TILE * 6)) + prow * 6;
        const uint32_t s_raw = smem_u32(raw);
        long long Rn = ((long long)blockIdx.x + (long long)pg * gridDim.x) * FP_TILE + prow;
        if (pg < nseq && Rn < a.total_rows) {
            const float* dp = a.dist2 + (size_t)Rn * 3; const int* ip = a.idx + (size_t)Rn * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_raw + 4 * k), "l"(dp + k) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_raw + 12 + 4 * k), "l"(ip + k) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        int buf = 0;
        for (int i = pg; i < nseq; i += FP_PG, buf ^= 1) {
            const long long R = Rn;
            const bool live = R < a.total_rows;
            // ---- interpolation weights of row `prow`, exactly the reference's torch arithmetic (pointnet2_modules.py:141-143)
            uint32_t* rowpt = meta + buf * (FP_TILE * 6);
            float* roww = reinterpret_cast<float*>(rowpt + FP_TILE * 3);
            asm volatile("cp.async.wait_group 0;" ::: "memory");          // this thread's own staging words have landed
            {
                uint32_t p0 = 0xFFFFFFFFu, p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;
                float w0 = 0.f, w1 = 0.f, w2 = 0.f;
                if (live) {
                    const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(raw[0]), 1e-8f));
                    const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(raw[1]), 1e-8f));
                    const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(raw[2]), 1e-8f));
                    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
                    w0 = __fdiv_rn(r0, norm); w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm);
                    const uint32_t base = (uint32_t)((unsigned long long)R / (unsigned)a.n) * (uint32_t)a.m;
                    p0 = base + __float_as_uint(raw[3]); p1 = base + __float_as_uint(raw[4]); p2 = base + __float_as_uint(raw[5]);
                }
                rowpt[prow] = p0; rowpt[FP_TILE + prow] = p1; rowpt[2 * FP_TILE + prow] = p2;
                roww[prow] = w0; roww[FP_TILE + prow] = w1; roww[2 * FP_TILE + prow] = w2;
            }
            group_bar(1 + FP_CG + pg);          // meta[buf] complete (it is rewritten two tiles later, after the next barrier)
            // ---- prefetch the raw meta data of this group's next tile: the latency hides behind the gather
            Rn = R + (long long)FP_PG * gridDim.x * FP_TILE;
            if (i + FP_PG < nseq && Rn < a.total_rows) {
                const float* dp = a.dist2 + (size_t)Rn * 3; const int* ip = a.idx + (size_t)Rn * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_raw + 4 * k), "l"(dp + k) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_raw + 12 + 4 * k), "l"(ip + k) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            // ---- A buffer of this tile
            const int ab = i % NA;
            const uint32_t use = (uint32_t)(i / NA);
            mbar_wait(bar_empty + 8 * ab, (use & 1) ^ 1);       // freed by the commit of the last layer of tile i - NA (first lap passes)
            uint4* dst = reinterpret_cast<uint4*>(smem + L.off_a + (size_t)ab * L.a_bytes);
            // ---- gather + interpolate: warp pw owns rows 32pw..32pw+31 in two rounds of 16 rows = 2 groups of 8 rows.
            //      Lane (rl, cl) loads 32 bytes (two 8-channel chunks) per tap with ONE 256-bit load, so that a warp-wide load
            //      covers 8 rows x 128 B = 8 full cache lines: the L1TEX data pipe (one wavefront per line touched) is the
            //      bottleneck of this kernel -- with 16-byte loads over 8 rows x 64 B the same bytes cost twice the wavefronts.
            //      The 12 loads of a round (2 row groups x 3 taps x 2 column blocks = 12 KB per warp) are issued before their first use.
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t q[2][3];                                // chunk index of the tap's row start (point id * chunks per row)
                float wt[2][3];
                bool valid[2];
                const uint4* kp = reinterpret_cast<const uint4*>(a.known_pm);
#pragma unroll
                for (int rg = 0; rg < 2; ++rg) {
                    const int row = pw * 32 + h * 16 + rg * 8 + rl;
                    valid[rg] = rowpt[row] != 0xFFFFFFFFu;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        q[rg][k] = valid[rg] ? rowpt[k * FP_TILE + row] * (uint32_t)nchunk : 0u;      // b*m*c_in/8 < 2^32 (host check)
                        wt[rg][k] = roww[k * FP_TILE + row];
                    }
                }
#pragma unroll 1
                for (int cb = 2 * cl; cb < nchunk; cb += 16) {    // this lane's chunk pair of each 16-chunk column block pair
                    uint4 ld[2][3][2][2];                         // [row group][tap][column block][chunk of the pair]
#pragma unroll
                    for (int rg = 0; rg < 2; ++rg)
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int c = cb + 8 * u;
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                ld[rg][k][u][0] = ld[rg][k][u][1] = make_uint4(0, 0, 0, 0);
                                if (c < nchunk && valid[rg]) ldg256(kp + (q[rg][k] + (uint32_t)c), ld[rg][k][u][0], ld[rg][k][u][1]);
                            }
                        }
#pragma unroll
                    for (int rg = 0; rg < 2; ++rg) {
                        const int row = pw * 32 + h * 16 + rg * 8 + rl;
#pragma unroll
                        for (int u = 0; u < 2; ++u)
#pragma unroll
                            for (int e2 = 0; e2 < 2; ++e2) {
                                const int c = cb + 8 * u + e2;
                                if (c < nchunk) {
                                    float f0[8], f1[8], f2[8], r[8];
                                    unpack8(ld[rg][0][u][e2], f0); unpack8(ld[rg][1][u][e2], f1); unpack8(ld[rg][2][u][e2], f2);
#pragma unroll
                                    for (int e = 0; e < 8; ++e)     // interpolate_gpu.cu:96 in the reference build's order
                                        r[e] = __fmaf_rn(wt[rg][2], f2[e], __fmaf_rn(wt[rg][0], f0[e], __fmul_rn(wt[rg][1], f1[e])));
                                    uint4 o = make_uint4(0, 0, 0, 0);
                                    if (valid[rg])
                                        o = make_uint4(pack_f16x2(r[0], r[1]), pack_f16x2(r[2], r[3]), pack_f16x2(r[4], r[5]), pack_f16x2(r[6], r[7]));
                                    dst[(size_t)c * FP_TILE + row] = o;
                                }
                            }
                    }
                }
            }
            fence_proxy_async();                                 // generic-proxy stores -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * ab);
        }
    } else {
        // =========================== CONSUMERS ==================================================================
        const int cg = warp >> 2;                                // consumer group; its warps own TMEM lane quadrants warp & 3
        const int row = (warp & 3) * 32 + lane;                  // tile row == TMEM lane
        const bool iwarp = (warp & 3) == 0;                      // the group's MMA-issuing warp: converged code, one ELECTED lane issues
        const bool issuer = iwarp && lane == 0;                  // (an `if (lane == 0)` around tcgen05.mma compiles to an ELECT + R2UR.BROADCAST loop per instruction)
        const uint32_t tacc = tmem + (uint32_t)cg * L.tcols;     // this group's accumulator columns
        const uint32_t lane_taddr = tacc + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t s_w1 = smem_u32(smem + L.off_w1), s_w2 = smem_u32(smem + L.off_w2), s_w3 = smem_u32(smem + L.off_w3),
                       s_w4 = smem_u32(smem + L.off_w4);
        const uint32_t my_done = bar_done + 8 * cg;
        const int bar_id = 1 + cg;
        uint32_t phase = 0;
        mbar_wait(bar_w, 0);                                     // weights + biases resident
        const bool dbgc = a.dbg && blockIdx.x == 0 && cg == 0 && issuer;
        long long t_wait_full = 0, t_wait_done = 0, t_loop0 = dbgc ? clock64() : 0;
        long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tp = 0, e4a = 0, e4b = 0;   // per layer: issue | wait for the MMAs | epilogue + group barrier
#define FP_PH(i) do { if (dbgc) { const long long now_ = clock64(); ph[i] += now_ - tp; tp = now_; } } while (0)
        for (int i = cg; i < nseq; i += FP_CG) {
            const long long R = ((long long)blockIdx.x + (long long)i * gridDim.x) * FP_TILE + row;
            const bool live = R < a.total_rows;
            const unsigned cloud = live ? (unsigned)((unsigned long long)R / (unsigned)a.n) : 0u;
            const int pt = live ? (int)(R - (long long)cloud * a.n) : 0;
            const int ab = i % NA;
            const uint32_t use = (uint32_t)(i / NA);
            unsigned char* hbuf = smem + L.off_a + (size_t)ab * L.a_bytes;     // layer-1 operand, then (in place) the hidden activations
            const uint32_t s_h = smem_u32(hbuf);
            // ---- layer 1: A buffer -> D
            if (dbgc) tp = clock64();
            if (iwarp) {
                const long long tw0 = dbgc ? clock64() : 0;
                mbar_wait_spin(bar_full + 8 * ab, use & 1);
                if (dbgc) t_wait_full += clock64() - tw0;
                if (elect_one_sync()) {
                    fp_issue_layer(tacc, s_h, s_w1, L.c_in, L.c1);
                    umma_commit(my_done);
                }
            }
            __syncwarp();
            FP_PH(0);
            { const long long tw0 = dbgc ? clock64() : 0;
              mbar_wait(my_done, phase); phase ^= 1;
              if (dbgc) t_wait_done += clock64() - tw0; }
            tc_fence_after();
            FP_PH(1);
            fp_epilogue_relu(lane_taddr, L.c1, b1, hbuf, row, nullptr, 0);
            tc_fence_before(); fence_proxy_async(); group_bar(bar_id);
            FP_PH(2);
            // ---- layer 2 (FP output)
            if (iwarp && elect_one_sync()) {
                fp_issue_layer(tacc, s_h, s_w2, L.c1, L.c2);
                if (!L.h1) umma_commit(bar_empty + 8 * ab);      // last reader of the buffer: back to the producers when it retires
                umma_commit(my_done);
            }
            __syncwarp();
            FP_PH(3);
            mbar_wait(my_done, phase); phase ^= 1;
            tc_fence_after();
            FP_PH(4);
            fp_epilogue_relu(lane_taddr, L.c2, b2, hbuf, row, live ? a.out_feat + ((size_t)cloud * L.c2) * a.n + pt : nullptr, (size_t)a.n);
            tc_fence_before(); fence_proxy_async(); group_bar(bar_id);
            FP_PH(5);
            if (L.h1) {
                // ---- head layer 1
                if (iwarp && elect_one_sync()) { fp_issue_layer(tacc, s_h, s_w3, L.c2, L.h1); umma_commit(my_done); }
                __syncwarp();
                FP_PH(6);
                mbar_wait(my_done, phase); phase ^= 1;
                tc_fence_after();
                FP_PH(7);
                fp_epilogue_relu(lane_taddr, L.h1, b3, hbuf, row, nullptr, 0);
                tc_fence_before(); fence_proxy_async(); group_bar(bar_id);
                FP_PH(8);
                // ---- head layer 2: logits, no activation
                if (iwarp && elect_one_sync()) {
                    fp_issue_layer(tacc, s_h, s_w4, L.h1, L.h2p);
                    umma_commit(bar_empty + 8 * ab);             // last reader of the buffer: back to the producers when it retires
                    umma_commit(my_done);
                }
                __syncwarp();
                FP_PH(9);
                mbar_wait(my_done, phase); phase ^= 1;
                tc_fence_after();
                FP_PH(10);
                float v[16];
                tmem_ld16(lane_taddr, v);
                if (dbgc) e4a += clock64() - tp;
                if (live) {
                    // (staging these h2 floats per row through shared memory for fully coalesced stores was measured: 0.448 -> 0.477 ms;
                    // these 8 store instructions take ~2700 cycles to issue behind the producers' gather in the LSU queue, tools/fp_counters.py)
                    float* o = a.out_head + (size_t)R * L.h2;
                    float best = 0.f;
                    int lab = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < L.h2) {
                            const float z = v[j] + b4[j];
                            o[j] = z;
                            // torch.argmax: the first maximal class, NaN counts as the maximum (mesh_encoder.py:113)
                            if (j == 0 || z > best || (z != z && best == best)) { best = z; lab = j; }
                        }
                    if (a.out_label) a.out_label[R] = (unsigned char)lab;
                }
                if (dbgc) e4b += clock64() - tp;
                tc_fence_before(); group_bar(bar_id);            // every warp has drained D: the next tile's layer 1 may overwrite it
                FP_PH(11);
            }
        }
        if (dbgc) {
            a.dbg[2] = t_wait_full; a.dbg[3] = t_wait_done; a.dbg[4] = clock64() - t_loop0; a.dbg[5] = nseq;
            for (int i = 0; i < 12; ++i) a.dbg[8 + i] = ph[i];       // (buffer: >= 22 int64)
            a.dbg[20] = e4a; a.dbg[21] = e4b;
        }
#undef FP_PH
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, L.tmem_cols);
}

static void fp_put(__half* base, int R, int r, int k, float v) {
    base[((size_t)(k / 8) * R + r) * 8 + (k % 8)] = __float2half_rn(v);
}

}  // namespace g4d

using namespace g4d;

static long long* g_fp_dbg = nullptr;
// Debug aid: device buffer of >= 8 int64 that CTA 0 of the next g4d_fp_interp_mlp launches fills with cycle counters
// ([2] consumer group 0's issuer waiting for a full A buffer -- i.e. for the producers --, [3] waiting for layer 1's MMAs,
// [4] its whole loop, [5] tiles of the CTA; the producers carry no counters: they have no registers to spare); NULL = off.
G4D_API void g4d_debug_fp_counters(void* buf) { g_fp_dbg = (long long*)buf; }

G4D_API size_t g4d_fp_param_bytes(const g4d_fp_desc* d) {
    FpLayout L; const char* why = nullptr;
    if (!d || !fp_layout(d, &L, &why)) { set_error("%s", why ? why : "fp: null descriptor"); return 0; }
    return L.blob_bytes;
}

G4D_API int g4d_fp_pack_params(const g4d_fp_desc* d, const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* wh1, const float* bh1, const float* wh2, const float* bh2, void* blob) {
    FpLayout L; const char* why = nullptr;
    if (!d || !fp_layout(d, &L, &why)) return bad_arg(why ? why : "fp: null descriptor");
    if (!w1 || !b1 || !w2 || !b2 || !blob || (L.h1 && (!wh1 || !bh1 || !wh2 || !bh2))) return bad_arg("fp_pack_params: null pointer");
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, L.blob_bytes);
    __half* W1 = (__half*)(out + L.off_w1);
    for (int o = 0; o < L.c1; ++o) for (int k = 0; k < L.c_in; ++k) fp_put(W1, L.c1, o, k, w1[(size_t)o * L.c_in + k]);
    __half* W2 = (__half*)(out + L.off_w2);
    for (int o = 0; o < L.c2; ++o) for (int k = 0; k < L.c1; ++k) fp_put(W2, L.c2, o, k, w2[(size_t)o * L.c1 + k]);
    memcpy(out + L.off_b1, b1, 4 * (size_t)L.c1);
    memcpy(out + L.off_b2, b2, 4 * (size_t)L.c2);
    if (L.h1) {
        __half* W3 = (__half*)(out + L.off_w3);
        for (int o = 0; o < L.h1; ++o) for (int k = 0; k < L.c2; ++k) fp_put(W3, L.h1, o, k, wh1[(size_t)o * L.c2 + k]);
        __half* W4 = (__half*)(out + L.off_w4);
        for (int o = 0; o < L.h2; ++o) for (int k = 0; k < L.h1; ++k) fp_put(W4, L.h2p, o, k, wh2[(size_t)o * L.h1 + k]);
        memcpy(out + L.off_b3, bh1, 4 * (size_t)L.h1);
        memcpy(out + L.off_b4, bh2, 4 * (size_t)L.h2);
    }
    return 0;
}

static int fp_interp_mlp_impl(const g4d_fp_desc* d, const void* params_dev, int b, int n, int m, const float* dist2, const int* idx,
                              const void* known_pm, float* out_feat, float* out_head, unsigned char* out_label, void* stream) {
    FpArgs a;
    const char* why = nullptr;
    if (!d || !fp_layout(d, &a.L, &why)) return bad_arg(why ? why : "fp: null descriptor");
    if (b < 0 || n < 0 || m <= 0) return bad_arg("fp_interp_mlp: bad size");
    if (b == 0 || n == 0) return 0;
    if (!params_dev || !dist2 || !idx || !known_pm || !out_feat || (a.L.h1 && !out_head)) return bad_arg("fp_interp_mlp: null pointer");
    if ((long long)b * m * (d->c_in / 8) > 0xFFFFFFFEll) return bad_arg("fp_interp_mlp: b*m*c_in/8 exceeds 32-bit chunk ids");
    if (((uintptr_t)params_dev & 15) || ((uintptr_t)known_pm & 31)) return bad_arg("fp_interp_mlp: params must be 16-byte, known_pm 32-byte aligned");
    a.n = n; a.m = m;
    a.total_rows = (long long)b * n;
    const long long nt = (a.total_rows + FP_TILE - 1) / FP_TILE;
    if (nt > INT32_MAX) return bad_arg("fp_interp_mlp: too many tiles");
    a.ntiles = (int)nt;
    a.dist2 = dist2; a.idx = idx; a.known_pm = (const __half*)known_pm; a.params = (const unsigned char*)params_dev;
    a.out_feat = out_feat; a.out_head = out_head; a.out_label = a.L.h1 ? out_label : nullptr;
    a.dbg = g_fp_dbg;
    cudaError_t e = cudaFuncSetAttribute(fp_interp_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total_smem);
    if (e != cudaSuccess) { set_error("fp_interp_mlp: shared memory opt-in (%u B): %s", a.L.total_smem, cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(fp_interp_mlp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    long long grid = sm_count();                                   // one persistent CTA per SM (all of its shared memory, half its TMEM)
    if (grid > a.ntiles) grid = a.ntiles;
    fp_interp_mlp_kernel<<<(unsigned)grid, FP_THREADS, a.L.total_smem, (cudaStream_t)stream>>>(a);
    return finish_launch("g4d fp_interp_mlp");
}

G4D_API int g4d_fp_interp_mlp(const g4d_fp_desc* d, const void* params_dev, int b, int n, int m, const float* dist2, const int* idx,
                              const void* known_pm, float* out_feat, float* out_head, void* stream) {
    return fp_interp_mlp_impl(d, params_dev, b, n, m, dist2, idx, known_pm, out_feat, out_head, nullptr, stream);
}

// = g4d_fp_interp_mlp, and out_label (b, n) uint8 = argmax over the head's outputs of every point (torch.argmax semantics: first
// maximal class), written by the last epilogue: the segmentation the model consumes (modules/mesh_encoder.py:113).
G4D_API int g4d_fp_interp_mlp_labels(const g4d_fp_desc* d, const void* params_dev, int b, int n, int m, const float* dist2,
                                     const int* idx, const void* known_pm, float* out_feat, float* out_head,
                                     unsigned char* out_label, void* stream) {
    if (d && d->h1 && !out_label) return bad_arg("fp_interp_mlp_labels: null label pointer");
    return fp_interp_mlp_impl(d, params_dev, b, n, m, dist2, idx, known_pm, out_feat, out_head, out_label, stream);
}
