// Grouped shared-MLP + max-pool on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces, for one scale of a set-abstraction module in eval mode (pointnet2_modules.py:37-51):
//     grouping_operation(xyz) - centroid ; grouping_operation(features) ; cat        (pointnet2_utils.py:250-258)
//     3 x [cuDNN 1x1 Conv2d -> BatchNorm2d -> ReLU]                                  (pytorch_utils.py:5-32)
//     F.max_pool2d(kernel=[1, nsample]) ; squeeze ; (torch.cat over scales)          (pointnet2_modules.py:42-55)
// i.e. 9+ kernels that each stream the (B, C, npoint, nsample) activation through HBM, with ONE persistent
// kernel that never materialises the grouped tensor:
//
//   per 128-row tile (= 128/nsample whole neighbourhoods):
//     gather  : rows [features(idx) | xyz(idx) - centroid] are written straight into the UMMA canonical
//               K-major shared-memory layout (fp16; the 3 relative coordinates are split hi+lo so that the
//               geometry enters the first layer at ~fp32 precision)
//     layer 1 : D1[128 x c1]  = A0[128 x k0] . W1^T      tcgen05.mma kind::f16, fp32 accumulate in TMEM
//     epi 1   : TMEM -> regs (+bias, ReLU, ->fp16) -> shared (canonical layout again)
//     layer 2 : D2[128 x c2]  = H1 . W2^T
//     epi 2   : as epi 1
//     layer 3 : D3[c3 x 128]  = W3 . H2^T   -- TRANSPOSED: channels on TMEM lanes, positions on columns, so the
//               max over a neighbourhood is a max over nsample consecutive columns inside ONE thread
//     epi 3   : max over each neighbourhood, + bias, ReLU (max and the monotone bias+ReLU commute), written
//               channel-major fp32 at the scale's channel offset (fuses torch.cat) and, optionally,
//               point-major fp16 for the next level's gather.
//   Folded weights of all three layers stay resident in shared memory (one cp.async.bulk / TMA bulk copy per CTA).
//
// Shared-memory operand layout ("canonical K-major, no swizzle", cute::UMMA::LayoutType::SWIZZLE_NONE):
//   element (row r, k) of an operand with R rows lives at byte ((k/8)*R + r)*16 + (k%8)*2
//   => core matrix = 8 rows x 16 B contiguous; SBO (next 8-row group) = 128 B; LBO (next K chunk) = R*16 B.
#include <string.h>
#include "common.cuh"
#include "umma.cuh"
#include "garment4d_b200.h"

namespace g4d {

constexpr int TILE_M = 128;
constexpr int SA_THREADS = 128;
constexpr int XYZ_SLOTS = 9;     // [hi(3) | lo(3) | hi(3)] against weights [wh | wh | wl]

struct SaMlpLayout {
    int k0, c1, c2, c3, c3p, nb3;
    uint32_t off_w1, off_w2, off_w3, off_b1, off_b2, off_b3, blob_bytes;   // inside the parameter blob == smem image
    uint32_t off_act, act_bytes, off_rowpt, off_bar, total_smem;
    uint32_t tmem_cols;
};

static inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

static bool make_layout(const g4d_sa_mlp_desc* d, SaMlpLayout* L, const char** why) {
    static const char* msgs[] = {
        "sa_mlp: c1 and c2 must be multiples of 16 in [16, 256]", "sa_mlp: c3 must be in [1, 256]",
        "sa_mlp: nsample must be 8, 16, 32, 64 or 128", "sa_mlp: c_in must be a non-negative multiple of 8",
        "sa_mlp: k0 does not match g4d_sa_mlp_k0(c_in)", "sa_mlp: shared memory footprint exceeds 227 KB"};
    if (d->c1 < 16 || d->c1 > 256 || d->c1 % 16 || d->c2 < 16 || d->c2 > 256 || d->c2 % 16) { *why = msgs[0]; return false; }
    if (d->c3 < 1 || d->c3 > 256) { *why = msgs[1]; return false; }
    if (!(d->nsample == 8 || d->nsample == 16 || d->nsample == 32 || d->nsample == 64 || d->nsample == 128)) { *why = msgs[2]; return false; }
    if (d->c_in < 0 || d->c_in % 8) { *why = msgs[3]; return false; }
    if (d->k0 != g4d_sa_mlp_k0(d->c_in)) { *why = msgs[4]; return false; }
    L->k0 = d->k0; L->c1 = d->c1; L->c2 = d->c2; L->c3 = d->c3;
    L->c3p = d->c3 <= 128 ? 128 : 256;
    L->nb3 = L->c3p / 128;
    uint32_t o = 0;
    L->off_w1 = o; o += (uint32_t)L->k0 * L->c1 * 2;
    L->off_w2 = o; o += (uint32_t)L->c1 * L->c2 * 2;
    L->off_w3 = o; o += (uint32_t)L->c2 * L->c3p * 2;
    L->off_b1 = o; o += (uint32_t)L->c1 * 4;
    L->off_b2 = o; o += (uint32_t)L->c2 * 4;
    L->off_b3 = o; o += (uint32_t)L->c3p * 4;
    L->blob_bytes = o;                                   // multiple of 16 by construction
    int kmax = L->k0 > L->c1 ? L->k0 : L->c1;
    if (L->c2 > kmax) kmax = L->c2;
    L->off_act = round_up(o, 128);
    L->act_bytes = (uint32_t)TILE_M * kmax * 2;
    L->off_rowpt = L->off_act + L->act_bytes;
    L->off_bar = L->off_rowpt + TILE_M * 4;
    L->total_smem = L->off_bar + 64;
    uint32_t cols = L->c1 > L->c2 ? L->c1 : L->c2;
    if ((uint32_t)(128 * L->nb3) > cols) cols = 128 * L->nb3;
    uint32_t p2 = 32;
    while (p2 < cols) p2 <<= 1;
    L->tmem_cols = p2;
    if (L->total_smem > 227 * 1024) { *why = msgs[5]; return false; }
    return true;
}

// ---------------------------------------------------------------------------------------------------------

struct SaMlpArgs {
    SaMlpLayout L;
    int c_in, nsample, n, m;
    long long total_rows;            // b * m * nsample
    int ntiles;
    const float* xyz;                // (b, n, 3)
    const float* new_xyz;            // (b, m, 3)
    const int* idx;                  // (b, m, nsample)
    const __half* feat_pm;           // (b, n, c_in) or null
    const unsigned char* params;     // packed blob (device)
    float* out_cm;                   // (b, ctot, m)
    __half* out_pm;                  // (b, m, ctot) or null
    int ctot, coff;
};

__global__ void __launch_bounds__(SA_THREADS)
sa_mlp_max_kernel(const SaMlpArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SaMlpLayout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* act = smem + L.off_act;
    uint32_t* rowpt = reinterpret_cast<uint32_t*>(smem + L.off_rowpt);
    const float* b1 = reinterpret_cast<const float*>(smem + L.off_b1);
    const float* b2 = reinterpret_cast<const float*>(smem + L.off_b2);
    const float* b3 = reinterpret_cast<const float*>(smem + L.off_b3);
    const uint32_t bar_w = smem_u32(smem + L.off_bar), bar_mma = bar_w + 8, tmem_slot = bar_w + 16;
    const uint32_t s_w1 = smem_u32(smem + L.off_w1), s_w2 = smem_u32(smem + L.off_w2), s_w3 = smem_u32(smem + L.off_w3);
    const uint32_t s_act = smem_u32(act);

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
        bulk_g2s(smem_u32(smem), a.params, L.blob_bytes, bar_w);      // folded weights + biases, once per CTA
    }
    mbar_wait(bar_w, 0);

    const int nchunk_feat = a.c_in >> 3;            // 16-byte chunks of feature channels per row
    const int nchunk_k0 = L.k0 >> 3;
    const int groups_per_tile = TILE_M / a.nsample;
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        // ---------------- gather: build A0[128 x k0] in canonical layout -------------------------------
        {
            const long long R = (long long)tile * TILE_M + tid;       // global row = (cloud, centroid, sample)
            uint32_t pt = 0xFFFFFFFFu;
            uint4 c0 = make_uint4(0, 0, 0, 0), c1v = make_uint4(0, 0, 0, 0);
            if (R < a.total_rows) {
                const long long gp = R / a.nsample;                   // global centroid
                const int cloud = (int)(gp / a.m);
                const int src = __ldg(a.idx + R);
                pt = (uint32_t)((long long)cloud * a.n + src);
                const float* p = a.xyz + (size_t)pt * 3;
                const float* q = a.new_xyz + (size_t)gp * 3;
                const float dx = __ldg(p) - __ldg(q), dy = __ldg(p + 1) - __ldg(q + 1), dz = __ldg(p + 2) - __ldg(q + 2);
                const __half hx = __float2half_rn(dx), hy = __float2half_rn(dy), hz = __float2half_rn(dz);
                const float lx = dx - __half2float(hx), ly = dy - __half2float(hy), lz = dz - __half2float(hz);
                const uint32_t uhx = __half_as_ushort(hx), uhy = __half_as_ushort(hy), uhz = __half_as_ushort(hz);
                const uint32_t ulx = __half_as_ushort(__float2half_rn(lx)), uly = __half_as_ushort(__float2half_rn(ly)),
                               ulz = __half_as_ushort(__float2half_rn(lz));
                // slots: hi.x hi.y hi.z lo.x lo.y lo.z hi.x hi.y | hi.z 0 0 0 0 0 0 0
                c0 = make_uint4(uhx | (uhy << 16), uhz | (ulx << 16), uly | (ulz << 16), uhx | (uhy << 16));
                c1v = make_uint4(uhz, 0, 0, 0);
            }
            rowpt[tid] = pt;
            uint4* dst = reinterpret_cast<uint4*>(act);
            dst[(size_t)nchunk_feat * TILE_M + tid] = c0;
            dst[(size_t)(nchunk_feat + 1) * TILE_M + tid] = c1v;
            for (int c = nchunk_feat + 2; c < nchunk_k0; ++c) dst[(size_t)c * TILE_M + tid] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (nchunk_feat) {
                // warp w owns rows 32w..32w+31: 8 rows x 4 chunks per step, 16 B per lane
                const int rl = lane & 7, cl = lane >> 3;
                for (int rg = 0; rg < 4; ++rg) {
                    const int row = warp * 32 + rg * 8 + rl;
                    const uint32_t rp = rowpt[row];
                    const uint4* srcrow = reinterpret_cast<const uint4*>(a.feat_pm + (size_t)rp * a.c_in);
                    for (int c = cl; c < nchunk_feat; c += 4) {
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (rp != 0xFFFFFFFFu) v = __ldg(srcrow + c);
                        dst[(size_t)c * TILE_M + row] = v;
                    }
                }
            }
        }
        fence_proxy_async();
        __syncthreads();

        // ---------------- layer 1 ----------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc(TILE_M, L.c1);
            for (int k = 0; k < L.k0 / 16; ++k) {
                const uint64_t ad = umma_desc(s_act + (uint32_t)k * 2 * TILE_M * 16, TILE_M * 16, 128);
                const uint64_t bd = umma_desc(s_w1 + (uint32_t)k * 2 * L.c1 * 16, L.c1 * 16, 128);
                umma_f16(tmem, ad, bd, idesc, k > 0);
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase); phase ^= 1;
        tc_fence_after();
        // epilogue 1: row tid, bias + ReLU -> fp16 -> act (A0 is dead: MMA 1 has completed)
        for (int c0 = 0; c0 < L.c1; c0 += 16) {
            float v[16];
            tmem_ld16(lane_taddr + c0, v);
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = pack_relu_f16x2(v[2 * i] + b1[c0 + 2 * i], v[2 * i + 1] + b1[c0 + 2 * i + 1]);
            uint4* dst = reinterpret_cast<uint4*>(act);
            dst[(size_t)(c0 / 8) * TILE_M + tid] = make_uint4(h[0], h[1], h[2], h[3]);
            dst[(size_t)(c0 / 8 + 1) * TILE_M + tid] = make_uint4(h[4], h[5], h[6], h[7]);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();

        // ---------------- layer 2 ----------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc(TILE_M, L.c2);
            for (int k = 0; k < L.c1 / 16; ++k) {
                const uint64_t ad = umma_desc(s_act + (uint32_t)k * 2 * TILE_M * 16, TILE_M * 16, 128);
                const uint64_t bd = umma_desc(s_w2 + (uint32_t)k * 2 * L.c2 * 16, L.c2 * 16, 128);
                umma_f16(tmem, ad, bd, idesc, k > 0);
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase); phase ^= 1;
        tc_fence_after();
        for (int c0 = 0; c0 < L.c2; c0 += 16) {
            float v[16];
            tmem_ld16(lane_taddr + c0, v);
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = pack_relu_f16x2(v[2 * i] + b2[c0 + 2 * i], v[2 * i + 1] + b2[c0 + 2 * i + 1]);
            uint4* dst = reinterpret_cast<uint4*>(act);
            dst[(size_t)(c0 / 8) * TILE_M + tid] = make_uint4(h[0], h[1], h[2], h[3]);
            dst[(size_t)(c0 / 8 + 1) * TILE_M + tid] = make_uint4(h[4], h[5], h[6], h[7]);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();

        // ---------------- layer 3, transposed: D3[c3p x 128] = W3 . H2^T --------------------------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc(128, TILE_M);
            for (int j = 0; j < L.nb3; ++j)
                for (int k = 0; k < L.c2 / 16; ++k) {
                    const uint64_t ad = umma_desc(s_w3 + (uint32_t)k * 2 * L.c3p * 16 + (uint32_t)j * 128 * 16, L.c3p * 16, 128);
                    const uint64_t bd = umma_desc(s_act + (uint32_t)k * 2 * TILE_M * 16, TILE_M * 16, 128);
                    umma_f16(tmem + j * 128, ad, bd, idesc, k > 0);
                }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase); phase ^= 1;
        tc_fence_after();
        // epilogue 3: lane = output channel; max over each neighbourhood's nsample consecutive columns
        for (int j = 0; j < L.nb3; ++j) {
            const int ch = j * 128 + tid;
            const float bias = b3[ch];
            float run = -INFINITY;
            for (int c0 = 0; c0 < TILE_M; c0 += 16) {
                float v[16];
                tmem_ld16(lane_taddr + j * 128 + c0, v);
                if (a.nsample == 8) {
#pragma unroll
                    for (int hgrp = 0; hgrp < 2; ++hgrp) {
                        float mx = v[hgrp * 8];
#pragma unroll
                        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[hgrp * 8 + i]);
                        const long long gp = (long long)tile * groups_per_tile + (c0 >> 3) + hgrp;
                        if (ch < L.c3 && gp * a.nsample < a.total_rows) {
                            const float o = fmaxf(mx + bias, 0.f);
                            const long long cloud = gp / a.m, p = gp - cloud * a.m;
                            a.out_cm[((size_t)cloud * a.ctot + a.coff + ch) * a.m + p] = o;
                            if (a.out_pm) a.out_pm[(size_t)gp * a.ctot + a.coff + ch] = __float2half_rn(fminf(o, 65504.f));
                        }
                    }
                } else {
                    float mx = v[0];
#pragma unroll
                    for (int i = 1; i < 16; ++i) mx = fmaxf(mx, v[i]);
                    run = fmaxf(run, mx);
                    if (((c0 + 16) % a.nsample) == 0) {
                        const long long gp = (long long)tile * groups_per_tile + c0 / a.nsample;
                        if (ch < L.c3 && gp * a.nsample < a.total_rows) {
                            const float o = fmaxf(run + bias, 0.f);
                            const long long cloud = gp / a.m, p = gp - cloud * a.m;
                            a.out_cm[((size_t)cloud * a.ctot + a.coff + ch) * a.m + p] = o;
                            if (a.out_pm) a.out_pm[(size_t)gp * a.ctot + a.coff + ch] = __float2half_rn(fminf(o, 65504.f));
                        }
                        run = -INFINITY;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();     // TMEM and act are free for the next tile
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, L.tmem_cols);
}

}  // namespace g4d

using namespace g4d;

G4D_API int g4d_sa_mlp_k0(int c_in) { return (c_in + XYZ_SLOTS + 15) / 16 * 16; }

G4D_API size_t g4d_sa_mlp_param_bytes(const g4d_sa_mlp_desc* d) {
    SaMlpLayout L; const char* why = nullptr;
    if (!d || !make_layout(d, &L, &why)) { set_error("%s", why ? why : "sa_mlp: null descriptor"); return 0; }
    return L.blob_bytes;
}

// canonical K-major image of a (rows x K) fp16 operand
static void put_canonical(__half* base, int R, int r, int k, float v) {
    base[((size_t)(k / 8) * R + r) * 8 + (k % 8)] = __float2half_rn(v);
}

G4D_API int g4d_sa_mlp_pack_params(const g4d_sa_mlp_desc* d, const float* w1, const float* b1, const float* w2, const float* b2,
                                   const float* w3, const float* b3, void* blob) {
    SaMlpLayout L; const char* why = nullptr;
    if (!d || !make_layout(d, &L, &why)) return bad_arg(why ? why : "sa_mlp: null descriptor");
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !blob) return bad_arg("sa_mlp_pack_params: null pointer");
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, L.blob_bytes);
    const int cin = d->c_in, ld1 = 3 + cin;     // reference column order: [xyz(3) | features(c_in)]  (pointnet2_utils.py:258)
    __half* W1 = (__half*)(out + L.off_w1);
    for (int o = 0; o < L.c1; ++o) {
        for (int k = 0; k < cin; ++k) put_canonical(W1, L.c1, o, k, w1[(size_t)o * ld1 + 3 + k]);
        for (int j = 0; j < 3; ++j) {
            const float w = w1[(size_t)o * ld1 + j];
            const float wh = __half2float(__float2half_rn(w));
            put_canonical(W1, L.c1, o, cin + j, wh);          // x_hi * w_hi
            put_canonical(W1, L.c1, o, cin + 3 + j, wh);      // x_lo * w_hi
            put_canonical(W1, L.c1, o, cin + 6 + j, w - wh);  // x_hi * w_lo
        }
    }
    __half* W2 = (__half*)(out + L.off_w2);
    for (int o = 0; o < L.c2; ++o)
        for (int k = 0; k < L.c1; ++k) put_canonical(W2, L.c2, o, k, w2[(size_t)o * L.c1 + k]);
    __half* W3 = (__half*)(out + L.off_w3);
    for (int o = 0; o < L.c3; ++o)
        for (int k = 0; k < L.c2; ++k) put_canonical(W3, L.c3p, o, k, w3[(size_t)o * L.c2 + k]);
    memcpy(out + L.off_b1, b1, sizeof(float) * L.c1);
    memcpy(out + L.off_b2, b2, sizeof(float) * L.c2);
    memcpy(out + L.off_b3, b3, sizeof(float) * L.c3);
    return 0;
}

G4D_API int g4d_sa_mlp_max(const g4d_sa_mlp_desc* d, const void* params_dev, int b, int n, int m, const float* xyz,
                           const float* new_xyz, const int* idx, const void* feat_pm, float* out_cm, void* out_pm,
                           int out_c_total, int out_c_off, void* stream) {
    SaMlpArgs a;
    const char* why = nullptr;
    if (!d || !make_layout(d, &a.L, &why)) return bad_arg(why ? why : "sa_mlp: null descriptor");
    if (b < 0 || n <= 0 || m < 0) return bad_arg("sa_mlp_max: bad size");
    if (b == 0 || m == 0) return 0;
    if (!params_dev || !xyz || !new_xyz || !idx || !out_cm || (d->c_in > 0 && !feat_pm)) return bad_arg("sa_mlp_max: null pointer");
    if (out_c_off < 0 || out_c_off + d->c3 > out_c_total) return bad_arg("sa_mlp_max: channel window outside the output");
    if ((long long)b * n > 0xFFFFFFFEll) return bad_arg("sa_mlp_max: b*n exceeds 32-bit point ids");
    if (((uintptr_t)params_dev & 15) || ((uintptr_t)feat_pm & 15)) return bad_arg("sa_mlp_max: params/feat_pm must be 16-byte aligned");
    a.c_in = d->c_in; a.nsample = d->nsample; a.n = n; a.m = m;
    a.total_rows = (long long)b * m * d->nsample;
    const long long nt = (a.total_rows + TILE_M - 1) / TILE_M;
    if (nt > INT32_MAX) return bad_arg("sa_mlp_max: too many tiles");
    a.ntiles = (int)nt;
    a.xyz = xyz; a.new_xyz = new_xyz; a.idx = idx; a.feat_pm = (const __half*)feat_pm;
    a.params = (const unsigned char*)params_dev;
    a.out_cm = out_cm; a.out_pm = (__half*)out_pm; a.ctot = out_c_total; a.coff = out_c_off;

    cudaError_t e = cudaFuncSetAttribute(sa_mlp_max_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total_smem);
    if (e != cudaSuccess) { set_error("sa_mlp_max: shared memory opt-in (%u B): %s", a.L.total_smem, cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(sa_mlp_max_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    // resident CTAs per SM: shared memory (227 KB usable, 1 KB reserved per CTA), TMEM columns (512 per SM), warps
    int occ = (int)((227u * 1024u) / (a.L.total_smem + 1024u));
    const int tmem_limit = 512 / (int)a.L.tmem_cols;
    if (occ > tmem_limit) occ = tmem_limit;
    if (occ > 8) occ = 8;
    if (occ < 1) occ = 1;
    long long grid = (long long)sm_count() * occ;
    if (grid > a.ntiles) grid = a.ntiles;
    sa_mlp_max_kernel<<<(unsigned)grid, SA_THREADS, a.L.total_smem, (cudaStream_t)stream>>>(a);
    return finish_launch("g4d sa_mlp_max");
}
