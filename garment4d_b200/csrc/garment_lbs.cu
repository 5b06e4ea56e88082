// Garment skinning by interpolated body weights, sm_100a.
//
// MeshEncoder.lbs_garment_interpolation (modules/mesh_encoder.py:312-410) -- the block the reference times itself (:434-441):
//   knn_points(garment, body, K)         K = cfg.NETWORK.LBSK (128 / 256), K64 = min(64, K), K = 1      (:321-324, chamferdist)
//   interp = 1 / dist2, inf -> 0, normalised over K, inf -> 0                                            (:341-345, :371-375)
//   nn_W = sum_k interp_k * W[idx_k]     the reference materialises (F, body_v, K, 24) for the gather   (:339-346, :377-379)
//   100 x  nn_W += 0.1 * Adj . nn_W      Adj = row-normalised mesh adjacency - I, torch.spmm            (:382-389)
//   vertices = (nn_W . A) [v ; 1]        = g4d_lbs_skin with per-frame weights                           (:391-408)
// Kernels here: the K-nearest-neighbour search (one warp per garment vertex: 4-pass radix select of the K-th smallest squared
// distance over the body vertices held in shared memory, ordered collection, bitonic sort by (distance, index)), the inverse-
// distance weights, the weighted gather of the skinning weights (no (F, V, K, J) intermediate), and one smoothing step.
// chamferdist (the package knn_points comes from) is neither vendored nor pinned by the reference: ordering among EQUAL
// distances is ours (smaller index first) and parity for this block is against the numpy restatement in oracle/mesh_ops.py only.
#include "common.cuh"

namespace g4d {

constexpr int KNN_WARPS = 16;
constexpr int KNN_MAXK = 256;

__device__ __forceinline__ unsigned knn_key(float qx, float qy, float qz, float x, float y, float z) {
    const float dx = qx - x, dy = qy - y, dz = qz - z;
    return __float_as_uint(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));   // dist += diff * diff, x then y then z
}

// query (b, nq, 3), ref (b, nr, 3) -> dist2 (b, nq, K) ascending, idx (b, nq, K); among equal distances the smaller index first.
__global__ void __launch_bounds__(KNN_WARPS * 32, 1)
knn_points_kernel(int nq, int nr, int K, const float* __restrict__ query_all, const float* __restrict__ ref_all,
                  float* __restrict__ dist2_all, int* __restrict__ idx_all) {
    extern __shared__ __align__(16) unsigned char knn_smem[];
    float* xs = reinterpret_cast<float*>(knn_smem);
    float* ys = xs + nr;
    float* zs = ys + nr;
    unsigned long long* cand_all = reinterpret_cast<unsigned long long*>(zs + nr + (nr & 1));      // 8-byte aligned
    unsigned* hist_all = reinterpret_cast<unsigned*>(cand_all + KNN_WARPS * KNN_MAXK);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t cloud = blockIdx.y;
    const float* ref = ref_all + cloud * (size_t)nr * 3;
    for (int i = tid; i < nr; i += blockDim.x) { xs[i] = __ldg(ref + 3 * i); ys[i] = __ldg(ref + 3 * i + 1); zs[i] = __ldg(ref + 3 * i + 2); }
    __syncthreads();
    unsigned long long* cand = cand_all + warp * KNN_MAXK;
    unsigned* hist = hist_all + warp * 256;
    int KP = 1;
    while (KP < K) KP <<= 1;
    for (int q = blockIdx.x * KNN_WARPS + warp; q < nq; q += gridDim.x * KNN_WARPS) {
        const float* qp = query_all + (cloud * nq + q) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        // ---- radix select: the key T of the K-th smallest distance and how many keys equal to T belong to the K
        unsigned prefix = 0;
        int kk = K;                                                   // rank still to find inside the current bucket (1-based)
#pragma unroll 1
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = lane; i < 256; i += 32) hist[i] = 0;
            __syncwarp();
            for (int i = lane; i < nr; i += 32) {
                const unsigned key = knn_key(qx, qy, qz, xs[i], ys[i], zs[i]);
                if (shift == 24 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            __syncwarp();
            // lane l owns bins 8l .. 8l+7
            unsigned c[8], s = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { c[e] = hist[8 * lane + e]; s += c[e]; }
            unsigned incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
            const unsigned excl = incl - s;
            const bool mine = (unsigned)kk > excl && (unsigned)kk <= incl;        // exactly one lane
            unsigned bin = 0, below = 0;
            if (mine) {
                unsigned run = excl;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if ((unsigned)kk > run && (unsigned)kk <= run + c[e]) { bin = 8 * lane + e; below = run; }
                    run += c[e];
                }
            }
            const int src = __ffs(__ballot_sync(0xFFFFFFFFu, mine)) - 1;
            bin = __shfl_sync(0xFFFFFFFFu, bin, src);
            below = __shfl_sync(0xFFFFFFFFu, below, src);
            prefix = (prefix << 8) | bin;
            kk -= (int)below;
            __syncwarp();
        }
        const unsigned T = prefix;
        // ---- ordered collection: every key < T, and the kk lowest-index keys == T
        int n_out = 0, n_eq = 0;
        for (int i0 = 0; i0 < nr; i0 += 32) {
            const int i = i0 + lane;
            unsigned key = 0xFFFFFFFFu;
            if (i < nr) key = knn_key(qx, qy, qz, xs[i], ys[i], zs[i]);
            const bool lt = i < nr && key < T, eq = i < nr && key == T;
            const unsigned meq = __ballot_sync(0xFFFFFFFFu, eq);
            const bool take = lt || (eq && n_eq + __popc(meq & ((1u << lane) - 1u)) < kk);
            const unsigned mt = __ballot_sync(0xFFFFFFFFu, take);
            if (take) cand[n_out + __popc(mt & ((1u << lane) - 1u))] = ((unsigned long long)key << 32) | (unsigned)i;
            n_out += __popc(mt);
            n_eq += __popc(meq);
        }
        for (int i = K + lane; i < KP; i += 32) cand[i] = ~0ull;
        __syncwarp();
        // ---- bitonic sort of KP keys (distance bits, index) in shared memory
        for (int size = 2; size <= KP; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = lane; t < (KP >> 1); t += 32) {
                    const int lo = ((t / stride) * stride << 1) + (t % stride), hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const unsigned long long a = cand[lo], b = cand[hi];
                    if ((a > b) == up) { cand[lo] = b; cand[hi] = a; }
                }
                __syncwarp();
            }
        float* dout = dist2_all + (cloud * nq + q) * (size_t)K;
        int* iout = idx_all + (cloud * nq + q) * (size_t)K;
        for (int i = lane; i < K; i += 32) {
            const unsigned long long v = cand[i];
            dout[i] = __uint_as_float((unsigned)(v >> 32));
            iout[i] = (int)(unsigned)v;
        }
        __syncwarp();
    }
}

// dist2 rows (rows, ld) -> w (rows, k): 1 / dist2, inf -> 0, divided by the row sum, inf -> 0  (mesh_encoder.py:341-345)
__global__ void knn_inverse_weights_kernel(long long rows, int ld, int k, const float* __restrict__ dist2, float* __restrict__ w) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* d = dist2 + r * ld;
    float* o = w + r * k;
    float sum = 0.f;
    for (int i = 0; i < k; ++i) {
        float v = __frcp_rn(__ldg(d + i));
        if (isinf(v)) v = 0.f;
        o[i] = v;
        sum += v;
    }
    for (int i = 0; i < k; ++i) {
        float v = __fdiv_rn(o[i], sum);
        if (isinf(v)) v = 0.f;
        o[i] = v;
    }
}

// out (B*T, nq, J) = sum_k w[b][q][k] * W[f][idx[b][q][k]][:], f = b * T + t; W (B*T, P, J).  One warp per (f, q), lane = joint.
__global__ void knn_blend_weights_kernel(int T, int nq, int P, int J, int K, int ld_idx, const int* __restrict__ idx,
                                         const float* __restrict__ w, const float* __restrict__ W, float* __restrict__ out,
                                         long long total) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= total) return;
    const long long f = gw / nq;
    const int q = (int)(gw - f * nq);
    const long long b = f / T;
    const int* ip = idx + (b * nq + q) * (long long)ld_idx;
    const float* wp = w + (b * nq + q) * (long long)K;
    const float* Wf = W + f * (long long)P * J;
    float acc = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int kl = k0 + lane;
        const int my_i = kl < K ? __ldg(ip + kl) : 0;
        const float my_w = kl < K ? __ldg(wp + kl) : 0.f;
        const int lim = min(32, K - k0);
        for (int k = 0; k < lim; ++k) {
            const int i = __shfl_sync(0xFFFFFFFFu, my_i, k);
            const float wk = __shfl_sync(0xFFFFFFFFu, my_w, k);
            if (lane < J) acc = __fmaf_rn(wk, __ldg(Wf + (long long)i * J + lane), acc);
        }
    }
    if (lane < J) out[gw * J + lane] = acc;
}

// y = x + coeff * Adj . x for every frame: x, y (F, G, J); Adj in CSR (rowptr G+1, col, val)   (mesh_encoder.py:389)
__global__ void smooth_weights_kernel(long long total, int G, int J, float coeff, const int* __restrict__ rowptr,
                                      const int* __restrict__ col, const float* __restrict__ val, const float* __restrict__ x,
                                      float* __restrict__ y) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int j = (int)(t % J);
    const long long fg = t / J;
    const int g = (int)(fg % G);
    const float* xf = x + (fg - g) * J;
    float acc = 0.f;
    const int e1 = __ldg(rowptr + g + 1);
    for (int e = __ldg(rowptr + g); e < e1; ++e) acc = __fmaf_rn(__ldg(val + e), __ldg(xf + (long long)__ldg(col + e) * J + j), acc);
    y[t] = __fmaf_rn(coeff, acc, __ldg(x + t));
}

}  // namespace g4d

using namespace g4d;

// K nearest reference points of every query point, ascending squared distance (ties: smaller index first).
// query (b, nq, 3), ref (b, nr, 3), dist2 (b, nq, K) fp32, idx (b, nq, K) int32.  1 <= K <= min(256, nr), nr <= 8192.
// Replaces chamferdist.knn_points as called at modules/mesh_encoder.py:321-324 (K = LBSK, min(64, LBSK) and 1 are prefixes of one call).
G4D_API int g4d_knn_points(int b, int nq, int nr, int K, const float* query, const float* ref, float* dist2, int* idx, void* stream) {
    if (b < 0 || nq < 0 || nr <= 0) return bad_arg("knn_points: need b >= 0, nq >= 0, nr > 0");
    if (K < 1 || K > KNN_MAXK || K > nr) return bad_arg("knn_points: need 1 <= K <= min(256, nr)");
    if (nr > 8192) return bad_arg("knn_points: nr > 8192");
    if (b == 0 || nq == 0) return 0;
    if (!query || !ref || !dist2 || !idx) return bad_arg("knn_points: null pointer");
    const size_t smem = (size_t)(nr + (nr & 1)) * 12 + 8 + (size_t)KNN_WARPS * KNN_MAXK * 8 + (size_t)KNN_WARPS * 256 * 4;
    cudaError_t e = cudaFuncSetAttribute(knn_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("knn_points: shared memory opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
    int gx = (nq + KNN_WARPS - 1) / KNN_WARPS;
    const int cap = (2 * sm_count() + b - 1) / b;                   // enough CTAs to fill the GPU; each loads the reference cloud once
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    knn_points_kernel<<<dim3(gx, b), KNN_WARPS * 32, smem, (cudaStream_t)stream>>>(nq, nr, K, query, ref, dist2, idx);
    return finish_launch("g4d knn_points");
}

// w (rows, k) from the first k columns of dist2 (rows, ld): 1 / dist2 with inf -> 0, normalised over k, inf -> 0.
G4D_API int g4d_knn_inverse_weights(long long rows, int ld, int k, const float* dist2, float* w, void* stream) {
    if (rows < 0 || k < 1 || ld < k) return bad_arg("knn_inverse_weights: need rows >= 0, 1 <= k <= ld");
    if (rows == 0) return 0;
    if (!dist2 || !w) return bad_arg("knn_inverse_weights: null pointer");
    knn_inverse_weights_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rows, ld, k, dist2, w);
    return finish_launch("g4d knn_inverse_weights");
}

// out (B*T, nq, J) = sum_k w (B, nq, K)[.., k] * W (B*T, P, J)[f, idx (B, nq, ld_idx)[.., k], :]   (J <= 32)
G4D_API int g4d_knn_blend_weights(int B, int T, int nq, int P, int J, int K, int ld_idx, const int* idx, const float* w,
                                  const float* W, float* out, void* stream) {
    if (B < 0 || T < 1 || nq < 0 || P < 1 || J < 1 || J > 32 || K < 1 || ld_idx < K) return bad_arg("knn_blend_weights: bad sizes (J <= 32, K <= ld_idx)");
    if (B == 0 || nq == 0) return 0;
    if (!idx || !w || !W || !out) return bad_arg("knn_blend_weights: null pointer");
    const long long total = (long long)B * T * nq;
    const long long blocks = (total * 32 + 255) / 256;
    if (blocks > 0x7FFFFFFFll) return bad_arg("knn_blend_weights: too many rows");
    knn_blend_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(T, nq, P, J, K, ld_idx, idx, w, W, out, total);
    return finish_launch("g4d knn_blend_weights");
}

// `iters` steps of x <- x + coeff * Adj . x on every frame (mesh_encoder.py:382-389).  x (F, G, J) in/out, tmp same size;
// Adj (G x G) in CSR.  The result is in x.
G4D_API int g4d_smooth_weights(int F, int G, int J, int iters, float coeff, const int* rowptr, const int* col, const float* val,
                               float* x, float* tmp, void* stream) {
    if (F < 0 || G < 1 || J < 1 || iters < 0) return bad_arg("smooth_weights: bad sizes");
    if (F == 0 || iters == 0) return 0;
    if (!rowptr || !col || !val || !x || !tmp) return bad_arg("smooth_weights: null pointer");
    const long long total = (long long)F * G * J;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7FFFFFFFll) return bad_arg("smooth_weights: too many elements");
    float *src = x, *dst = tmp;
    for (int it = 0; it < iters; ++it) {
        smooth_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total, G, J, coeff, rowptr, col, val, src, dst);
        float* t = src; src = dst; dst = t;
    }
    if (src != x) cudaMemcpyAsync(x, src, sizeof(float) * (size_t)total, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return finish_launch("g4d smooth_weights", iters);
}
