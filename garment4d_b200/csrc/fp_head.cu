// Fused feature propagation + segmentation head on tcgen05 (sm_100a).
//
// Replaces, for a PointnetFPModule without skip features in eval mode followed (optionally) by the encoder's
// FC head (pointnet2_modules.py:131-156 as used by pointnet2encoder.py:138-143 for the finest level):
//     dist_recip = 1/(dist+1e-8); weight = dist_recip / sum  (3 torch kernels)            pointnet2_modules.py:141-143
//     three_interpolate                                                                     interpolate_gpu.cu:77-97
//     unsqueeze ; 2 x [cuDNN conv1x1 -> BatchNorm2d -> ReLU] ; squeeze                      pytorch_utils.py:5-32
//     Conv1d(64,32)+BN+ReLU ; Dropout (eval: identity) ; Conv1d(32,classes) ; transpose     pointnet2encoder.py:98-101,143
// (~16 kernels, each streaming a (B, C, N) activation of up to 1 GB through HBM at c3) with ONE persistent kernel:
// per 128-point tile the 3-tap interpolation is evaluated while gathering (fp32, the reference's FMUL/FFMA order),
// written as fp16 in the UMMA canonical layout, and up to four chained tcgen05.mma layers run with fp32 TMEM
// accumulators; only the FP output (channel-major fp32, returned by the encoder as l_features[0]) and the logits
// (point-major) are written to HBM.
#include <string.h>
#include "common.cuh"
#include "umma.cuh"
#include "garment4d_b200.h"

namespace g4d {

constexpr int FP_TILE = 128;
constexpr int FP_THREADS = 256;     // warps 0-3: rows / TMEM lanes (weights, epilogues); all 8 warps gather

struct FpLayout {
    int c_in, c1, c2, h1, h2, h2p;      // h2p = 16 when a head is present (classes padded), else 0
    uint32_t off_w1, off_w2, off_w3, off_w4, off_b1, off_b2, off_b3, off_b4, blob_bytes;
    uint32_t off_act, off_meta, off_bar, total_smem, tmem_cols;
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
    int kmax = L->c_in;
    if (L->c1 > kmax) kmax = L->c1;
    if (L->c2 > kmax) kmax = L->c2;
    if (L->h1 > kmax) kmax = L->h1;
    L->off_act = (o + 127) / 128 * 128;
    L->off_meta = L->off_act + (uint32_t)FP_TILE * kmax * 2;
    L->off_bar = L->off_meta + FP_TILE * 6 * 4;           // 3 point ids + 3 weights per row
    L->total_smem = L->off_bar + 64;
    uint32_t cols = L->c1 > L->c2 ? L->c1 : L->c2;
    if ((uint32_t)L->h1 > cols) cols = L->h1;
    uint32_t p2 = 32;
    while (p2 < cols) p2 <<= 1;
    L->tmem_cols = p2;
    if (L->total_smem > 227 * 1024) { *why = "fp: shared memory footprint exceeds 227 KB"; return false; }
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
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// TMEM row -> +bias, ReLU, fp16 -> act buffer (canonical layout), optionally also fp32 channel-major to global
__device__ __forceinline__ void fp_epilogue_relu(uint32_t lane_taddr, int ncols, const float* bias, unsigned char* act, int tid,
                                                 float* gout /* channel 0 of this row's point, or null */, size_t gstride) {
    for (int c0 = 0; c0 < ncols; c0 += 16) {
        float v[16];
        tmem_ld16(lane_taddr + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + bias[c0 + i], 0.f);
        if (gout) {
#pragma unroll
            for (int i = 0; i < 16; ++i) gout[(size_t)(c0 + i) * gstride] = v[i];     // lanes = consecutive points: coalesced
        }
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
        uint4* dst = reinterpret_cast<uint4*>(act);
        dst[(size_t)(c0 / 8) * FP_TILE + tid] = make_uint4(h[0], h[1], h[2], h[3]);
        dst[(size_t)(c0 / 8 + 1) * FP_TILE + tid] = make_uint4(h[4], h[5], h[6], h[7]);
    }
}

__device__ __forceinline__ void fp_issue_layer(uint32_t tmem, uint32_t s_act, uint32_t s_w, int K, int N, uint32_t bar) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc(FP_TILE, N);
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t ad = umma_desc(s_act + (uint32_t)k * 2 * FP_TILE * 16, FP_TILE * 16, 128);
        const uint64_t bd = umma_desc(s_w + (uint32_t)k * 2 * N * 16, N * 16, 128);
        umma_f16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(bar);
}

__global__ void __launch_bounds__(FP_THREADS)
fp_interp_mlp_kernel(const FpArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const FpLayout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* act = smem + L.off_act;
    uint32_t* rowpt = reinterpret_cast<uint32_t*>(smem + L.off_meta);          // [3][128]
    float* roww = reinterpret_cast<float*>(smem + L.off_meta + FP_TILE * 3 * 4); // [3][128]
    const float* b1 = reinterpret_cast<const float*>(smem + L.off_b1);
    const float* b2 = reinterpret_cast<const float*>(smem + L.off_b2);
    const float* b3 = reinterpret_cast<const float*>(smem + L.off_b3);
    const float* b4 = reinterpret_cast<const float*>(smem + L.off_b4);
    const uint32_t bar_w = smem_u32(smem + L.off_bar), bar_mma = bar_w + 8, tmem_slot = bar_w + 16;
    const uint32_t s_act = smem_u32(act);
    const uint32_t s_w1 = smem_u32(smem + L.off_w1), s_w2 = smem_u32(smem + L.off_w2), s_w3 = smem_u32(smem + L.off_w3),
                   s_w4 = smem_u32(smem + L.off_w4);

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, L.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L.off_bar + 16);
    if (tid == 0) {
        mbar_expect_tx(bar_w, L.blob_bytes);
        bulk_g2s(smem_u32(smem), a.params, L.blob_bytes, bar_w);
    }
    mbar_wait(bar_w, 0);

    const int nchunk = L.c_in >> 3;
    const uint32_t lane_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const long long R = (long long)tile * FP_TILE + (tid & (FP_TILE - 1));
        const bool live = R < a.total_rows && tid < FP_TILE;
        const unsigned cloud = live ? (unsigned)((unsigned long long)R / (unsigned)a.n) : 0u;   // total_rows < 2^32 * n: one 64/32 division
        const int pt = live ? (int)(R - (long long)cloud * a.n) : 0;
        // ---- interpolation weights, exactly the reference's torch arithmetic (pointnet2_modules.py:141-143) ----
        {
            uint32_t p0 = 0xFFFFFFFFu, p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;
            float w0 = 0.f, w1 = 0.f, w2 = 0.f;
            if (live) {
                const float* d2 = a.dist2 + (size_t)R * 3;
                const int* id = a.idx + (size_t)R * 3;
                const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2)), 1e-8f));
                const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + 1)), 1e-8f));
                const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + 2)), 1e-8f));
                const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
                w0 = __fdiv_rn(r0, norm); w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm);
                const uint32_t base = cloud * (uint32_t)a.m;
                p0 = base + (uint32_t)__ldg(id); p1 = base + (uint32_t)__ldg(id + 1); p2 = base + (uint32_t)__ldg(id + 2);
            }
            if (tid < FP_TILE) {
                rowpt[tid] = p0; rowpt[FP_TILE + tid] = p1; rowpt[2 * FP_TILE + tid] = p2;
                roww[tid] = w0; roww[FP_TILE + tid] = w1; roww[2 * FP_TILE + tid] = w2;
            }
        }
        __syncthreads();
        // ---- gather + interpolate (all 8 warps) ----
        {
            const int rl = lane & 7, cl = lane >> 3;
            uint4* dst = reinterpret_cast<uint4*>(act);
            // warp w (of 8) owns rows 16w..16w+15 = 2 groups of 8 rows; 4 chunks per lane per group; the 24 loads of both
            // groups (3 taps x 4 chunks x 2) are issued before the first use: the gather is pure L2/HBM latency
            uint32_t q[2][3];
            float wt[2][3];
            const uint4* sp[2][3];
#pragma unroll
            for (int rg = 0; rg < 2; ++rg) {
                const int row = warp * 16 + rg * 8 + rl;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    q[rg][k] = rowpt[k * FP_TILE + row];
                    wt[rg][k] = roww[k * FP_TILE + row];
                    sp[rg][k] = reinterpret_cast<const uint4*>(a.known_pm + (size_t)q[rg][k] * L.c_in);
                }
            }
            for (int cb = cl; cb < nchunk; cb += 16) {
                uint4 ld[2][3][4];
#pragma unroll
                for (int rg = 0; rg < 2; ++rg)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = cb + 4 * u;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            ld[rg][k][u] = make_uint4(0, 0, 0, 0);
                            if (c < nchunk && q[rg][0] != 0xFFFFFFFFu) ld[rg][k][u] = __ldg(sp[rg][k] + c);
                        }
                    }
#pragma unroll
                for (int rg = 0; rg < 2; ++rg) {
                    const int row = warp * 16 + rg * 8 + rl;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = cb + 4 * u;
                        if (c < nchunk) {
                            float f0[8], f1[8], f2[8], r[8];
                            unpack8(ld[rg][0][u], f0); unpack8(ld[rg][1][u], f1); unpack8(ld[rg][2][u], f2);
#pragma unroll
                            for (int i = 0; i < 8; ++i)     // interpolate_gpu.cu:96 in the reference build's order
                                r[i] = __fmaf_rn(wt[rg][2], f2[i], __fmaf_rn(wt[rg][0], f0[i], __fmul_rn(wt[rg][1], f1[i])));
                            uint4 o = make_uint4(0, 0, 0, 0);
                            if (q[rg][0] != 0xFFFFFFFFu)
                                o = make_uint4(pack_f16x2(r[0], r[1]), pack_f16x2(r[2], r[3]), pack_f16x2(r[4], r[5]), pack_f16x2(r[6], r[7]));
                            dst[(size_t)c * FP_TILE + row] = o;
                        }
                    }
                }
            }
        }
        fence_proxy_async();
        __syncthreads();

        // ---- layer 1 ----
        if (tid == 0) fp_issue_layer(tmem, s_act, s_w1, L.c_in, L.c1, bar_mma);
        if (warp < 4) {
            mbar_wait(bar_mma, phase);
            tc_fence_after();
            fp_epilogue_relu(lane_taddr, L.c1, b1, act, tid, nullptr, 0);
        }
        phase ^= 1;
        tc_fence_before(); fence_proxy_async(); __syncthreads();
        // ---- layer 2 (FP output) ----
        if (tid == 0) fp_issue_layer(tmem, s_act, s_w2, L.c1, L.c2, bar_mma);
        if (warp < 4) {
            mbar_wait(bar_mma, phase);
            tc_fence_after();
            fp_epilogue_relu(lane_taddr, L.c2, b2, act, tid, live ? a.out_feat + ((size_t)cloud * L.c2) * a.n + pt : nullptr, (size_t)a.n);
        }
        phase ^= 1;
        tc_fence_before(); fence_proxy_async(); __syncthreads();
        if (L.h1) {
            // ---- head layer 1 ----
            if (tid == 0) fp_issue_layer(tmem, s_act, s_w3, L.c2, L.h1, bar_mma);
            if (warp < 4) {
                mbar_wait(bar_mma, phase);
                tc_fence_after();
                fp_epilogue_relu(lane_taddr, L.h1, b3, act, tid, nullptr, 0);
            }
            phase ^= 1;
            tc_fence_before(); fence_proxy_async(); __syncthreads();
            // ---- head layer 2: logits, no activation ----
            if (tid == 0) fp_issue_layer(tmem, s_act, s_w4, L.h1, L.h2p, bar_mma);
            phase ^= 1;
            if (warp < 4) {
                mbar_wait(bar_mma, phase ^ 1);
                tc_fence_after();
                float v[16];
                tmem_ld16(lane_taddr, v);
                if (live) {
                    float* o = a.out_head + (size_t)R * L.h2;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < L.h2) o[j] = v[j] + b4[j];
                }
            }
            tc_fence_before(); __syncthreads();
        }
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

G4D_API int g4d_fp_interp_mlp(const g4d_fp_desc* d, const void* params_dev, int b, int n, int m, const float* dist2, const int* idx,
                              const void* known_pm, float* out_feat, float* out_head, void* stream) {
    FpArgs a;
    const char* why = nullptr;
    if (!d || !fp_layout(d, &a.L, &why)) return bad_arg(why ? why : "fp: null descriptor");
    if (b < 0 || n < 0 || m <= 0) return bad_arg("fp_interp_mlp: bad size");
    if (b == 0 || n == 0) return 0;
    if (!params_dev || !dist2 || !idx || !known_pm || !out_feat || (a.L.h1 && !out_head)) return bad_arg("fp_interp_mlp: null pointer");
    if ((long long)b * m > 0xFFFFFFFEll) return bad_arg("fp_interp_mlp: b*m exceeds 32-bit point ids");
    if (((uintptr_t)params_dev & 15) || ((uintptr_t)known_pm & 15)) return bad_arg("fp_interp_mlp: params/known_pm must be 16-byte aligned");
    a.n = n; a.m = m;
    a.total_rows = (long long)b * n;
    const long long nt = (a.total_rows + FP_TILE - 1) / FP_TILE;
    if (nt > INT32_MAX) return bad_arg("fp_interp_mlp: too many tiles");
    a.ntiles = (int)nt;
    a.dist2 = dist2; a.idx = idx; a.known_pm = (const __half*)known_pm; a.params = (const unsigned char*)params_dev;
    a.out_feat = out_feat; a.out_head = out_head;
    cudaError_t e = cudaFuncSetAttribute(fp_interp_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total_smem);
    if (e != cudaSuccess) { set_error("fp_interp_mlp: shared memory opt-in (%u B): %s", a.L.total_smem, cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(fp_interp_mlp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    int occ = (int)((227u * 1024u) / (a.L.total_smem + 1024u));
    const int tmem_limit = 512 / (int)a.L.tmem_cols;
    if (occ > tmem_limit) occ = tmem_limit;
    if (occ > 2) occ = 2;          // 128 registers x 256 threads
    if (occ < 1) occ = 1;
    long long grid = (long long)sm_count() * occ;
    if (grid > a.ntiles) grid = a.ntiles;
    fp_interp_mlp_kernel<<<(unsigned)grid, FP_THREADS, a.L.total_smem, (cudaStream_t)stream>>>(a);
    return finish_launch("g4d fp_interp_mlp");
}
