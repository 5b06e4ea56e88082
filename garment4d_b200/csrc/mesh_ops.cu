// Callers of the hot path inside the garment model, sm_100a.
//
// g4d_select_points -- garment point selection (PCAGarmentEncoderSeg.calc_segmentation_results, modules/mesh_encoder.py:109-125):
//     labels = argmax(sem_logits, 2); per frame: x[labels == g][:n], feature[labels == g][:n], zero-padded to n rows
// The reference runs it as a Python loop over the B*T frames with boolean-mask indexing (one device->host sync per frame);
// here one CTA per frame does an ORDER-PRESERVING stream compaction (ballot + prefix popcount per warp, warp totals through
// shared memory), fused with the arg-max over the class logits, and writes the selected coordinates and feature rows.
#include "common.cuh"

namespace g4d {

constexpr int SEL_THREADS = 256;

// sem_logits (c, n, ncls) fp32 or NULL when labels (c, n) uint8 are given.  features (c, cf, n) channel-major (the encoder's
// l_features[0]) or NULL.  out_xyz (c, n_out, 3), out_feat (c, n_out, cf) point-major, out_count (c) selected before clipping.
__global__ void __launch_bounds__(SEL_THREADS)
select_points_kernel(int n, int ncls, int cf, int target, int n_out, const float* __restrict__ logits_all,
                     const unsigned char* __restrict__ labels_all, const float* __restrict__ xyz_all, const float* __restrict__ feat_all,
                     float* __restrict__ out_xyz_all, float* __restrict__ out_feat_all, int* __restrict__ out_count) {
    __shared__ int warp_cnt[SEL_THREADS / 32];
    __shared__ int base_s;
    const size_t cloud = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* logits = logits_all ? logits_all + cloud * (size_t)n * ncls : nullptr;
    const unsigned char* labels = labels_all ? labels_all + cloud * (size_t)n : nullptr;
    const float* xyz = xyz_all + cloud * (size_t)n * 3;
    const float* feat = feat_all ? feat_all + cloud * (size_t)cf * n : nullptr;
    float* out_xyz = out_xyz_all + cloud * (size_t)n_out * 3;
    float* out_feat = out_feat_all ? out_feat_all + cloud * (size_t)n_out * cf : nullptr;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int p0 = 0; p0 < n; p0 += SEL_THREADS) {
        const int p = p0 + tid;
        bool sel = false;
        if (p < n) {
            int lab;
            if (labels) lab = labels[p];
            else {
                // torch.argmax: the first maximal class (NaN counts as the maximum, like torch)
                const float* l = logits + (size_t)p * ncls;
                float best = __ldg(l);
                lab = 0;
                for (int k = 1; k < ncls; ++k) {
                    const float v = __ldg(l + k);
                    if (v > best || (v != v && best == best)) { best = v; lab = k; }
                }
            }
            sel = lab == target;
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, sel);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = base_s, total = 0;
#pragma unroll
        for (int w = 0; w < SEL_THREADS / 32; ++w) { const int c = warp_cnt[w]; if (w < warp) before += c; total += c; }
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        if (sel && pos < n_out) {
            out_xyz[3 * pos] = __ldg(xyz + 3 * p); out_xyz[3 * pos + 1] = __ldg(xyz + 3 * p + 1); out_xyz[3 * pos + 2] = __ldg(xyz + 3 * p + 2);
            if (out_feat)
                for (int ch = 0; ch < cf; ++ch) out_feat[(size_t)pos * cf + ch] = __ldg(feat + (size_t)ch * n + p);
        }
        __syncthreads();
        if (tid == 0) base_s += total;
        __syncthreads();
        if (base_s >= n_out && !out_count) break;        // block-uniform: everything that fits has been written
    }
    const int count = base_s;
    if (out_count && tid == 0) out_count[cloud] = count;
    // zero padding of the rows that were not filled (mesh_encoder.py:121-122)
    const int filled = count < n_out ? count : n_out;
    for (int e = filled * 3 + tid; e < n_out * 3; e += SEL_THREADS) out_xyz[e] = 0.f;
    if (out_feat)
        for (size_t e = (size_t)filled * cf + tid; e < (size_t)n_out * cf; e += SEL_THREADS) out_feat[e] = 0.f;
}


// ---------------------------------------------------------------------------------------------------------
// g4d_pe_mlp_max -- one "positional encoding" unit of the GCN refinement (PCALBSGarmentUseSegEncoderSeg.forward,
// modules/mesh_encoder.py:450-466):
//     QueryAndGroup(xyz, new_xyz, features) (B, 3+C, P, K) -> permute -> Linear(3+C, H) -> ReLU -> Linear(H, H) -> max over K
// 18 of them run per step (6 x 3 iterations), with gradients.  The reference materialises the grouped tensor and the two
// activations ((B, P, K, H) each); here one kernel reads idx (from our ball query), gathers, runs the two small layers in fp32
// on the CUDA cores (the reference's Linear is true fp32: torch's matmul TF32 switch is off by default) and reduces over the
// neighbourhood, writing (B, P, H) and, for the backward pass, the arg-max sample of every channel.
//   lane = one SAMPLE (32 / K centroids per warp): the lane streams its neighbour's fp32 feature row (point-major copy),
//   keeps the H = 32 layer-1 accumulators in registers; weights are broadcast from shared memory (W^T, 128-bit loads:
//   4 FMA per shared-memory instruction).
constexpr int PE_H = 32;

template <int K>
__global__ void __launch_bounds__(256)
pe_mlp_max_kernel(int n, int p, int c, const float* __restrict__ xyz_all, const float* __restrict__ new_xyz_all,
                  const float* __restrict__ feat_pm_all, const int* __restrict__ idx_all, const float* __restrict__ w1t,
                  const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
                  float* __restrict__ out_all, signed char* __restrict__ arg_all) {
    extern __shared__ __align__(16) float sw[];         // W1^T [(3+c)][32] | W2^T [32][32] | b1 [32] | b2 [32]
    const int cin = 3 + c;
    float* s_w1 = sw;
    float* s_w2 = s_w1 + (size_t)cin * PE_H;
    float* s_b1 = s_w2 + PE_H * PE_H;
    float* s_b2 = s_b1 + PE_H;
    for (int e = threadIdx.x; e < cin * PE_H; e += blockDim.x) s_w1[e] = __ldg(w1t + e);
    for (int e = threadIdx.x; e < PE_H * PE_H; e += blockDim.x) s_w2[e] = __ldg(w2t + e);
    if (threadIdx.x < PE_H) { s_b1[threadIdx.x] = __ldg(b1 + threadIdx.x); s_b2[threadIdx.x] = __ldg(b2 + threadIdx.x); }
    __syncthreads();
    constexpr int CPW = 32 / K;                          // centroids per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cloud = blockIdx.y;
    const int s = lane % K;
    const unsigned gmask = K == 32 ? 0xFFFFFFFFu : (((1u << K) - 1u) << (lane / K * K));
    const float* xyz = xyz_all + cloud * (size_t)n * 3;
    const float* feat = feat_pm_all ? feat_pm_all + cloud * (size_t)n * c : nullptr;
    for (int q0 = (blockIdx.x * (blockDim.x >> 5) + warp) * CPW; q0 < p; q0 += gridDim.x * (blockDim.x >> 5) * CPW) {
        const int q = q0 + lane / K;
        const bool live = q < p;
        const int qs = live ? q : p - 1;
        const int src = __ldg(idx_all + (cloud * p + qs) * K + s);
        const float* nq = new_xyz_all + (cloud * p + qs) * 3;
        float acc[PE_H];
#pragma unroll
        for (int j = 0; j < PE_H; ++j) acc[j] = s_b1[j];
        {
            const float x0 = __ldg(xyz + 3 * src) - __ldg(nq), x1 = __ldg(xyz + 3 * src + 1) - __ldg(nq + 1), x2 = __ldg(xyz + 3 * src + 2) - __ldg(nq + 2);
#pragma unroll
            for (int j4 = 0; j4 < PE_H / 4; ++j4) {
                const float4 wa = *reinterpret_cast<const float4*>(s_w1 + 0 * PE_H + 4 * j4), wb = *reinterpret_cast<const float4*>(s_w1 + 1 * PE_H + 4 * j4),
                             wc = *reinterpret_cast<const float4*>(s_w1 + 2 * PE_H + 4 * j4);
                acc[4 * j4] = fmaf(x2, wc.x, fmaf(x1, wb.x, fmaf(x0, wa.x, acc[4 * j4])));
                acc[4 * j4 + 1] = fmaf(x2, wc.y, fmaf(x1, wb.y, fmaf(x0, wa.y, acc[4 * j4 + 1])));
                acc[4 * j4 + 2] = fmaf(x2, wc.z, fmaf(x1, wb.z, fmaf(x0, wa.z, acc[4 * j4 + 2])));
                acc[4 * j4 + 3] = fmaf(x2, wc.w, fmaf(x1, wb.w, fmaf(x0, wa.w, acc[4 * j4 + 3])));
            }
        }
        if (feat) {
            const float* frow = feat + (size_t)src * c;
            for (int i = 0; i < c; ++i) {                  // c is small for the body normals (3); rows are 16-byte aligned when c % 4 == 0
                const float f = __ldg(frow + i);
                const float* wr = s_w1 + (size_t)(3 + i) * PE_H;
#pragma unroll
                for (int j4 = 0; j4 < PE_H / 4; ++j4) {
                    const float4 w = *reinterpret_cast<const float4*>(wr + 4 * j4);
                    acc[4 * j4] = fmaf(f, w.x, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(f, w.y, acc[4 * j4 + 1]);
                    acc[4 * j4 + 2] = fmaf(f, w.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(f, w.w, acc[4 * j4 + 3]);
                }
            }
        }
        float o[PE_H];
#pragma unroll
        for (int k = 0; k < PE_H; ++k) o[k] = s_b2[k];
#pragma unroll
        for (int j = 0; j < PE_H; ++j) {
            const float h = fmaxf(acc[j], 0.f);
#pragma unroll
            for (int k4 = 0; k4 < PE_H / 4; ++k4) {
                const float4 w = *reinterpret_cast<const float4*>(s_w2 + j * PE_H + 4 * k4);
                o[4 * k4] = fmaf(h, w.x, o[4 * k4]); o[4 * k4 + 1] = fmaf(h, w.y, o[4 * k4 + 1]);
                o[4 * k4 + 2] = fmaf(h, w.z, o[4 * k4 + 2]); o[4 * k4 + 3] = fmaf(h, w.w, o[4 * k4 + 3]);
            }
        }
        // max over the K samples of the centroid (lanes of the group); the FIRST maximal sample is recorded for the backward pass
        float* orow = out_all + (cloud * p + qs) * PE_H;
        signed char* arow = arg_all ? arg_all + (cloud * p + qs) * PE_H : nullptr;
#pragma unroll
        for (int k = 0; k < PE_H; ++k) {
            float m = o[k];
#pragma unroll
            for (int off = K / 2; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
            const unsigned eq = __ballot_sync(0xFFFFFFFFu, o[k] == m) & gmask;
            if (live && s == 0) {
                orow[k] = m;
                if (arow) arow[k] = (signed char)(__ffs(eq) - 1 - (lane / K) * K);
            }
        }
    }
}


}  // namespace g4d

using namespace g4d;

G4D_API int g4d_select_points(int c, int n, int ncls, int cf, int target, int n_out, const float* sem_logits, const unsigned char* labels,
                              const float* xyz, const float* features, float* out_xyz, float* out_feat, int* out_count, void* stream) {
    if (c < 0 || n < 0 || n_out < 0 || cf < 0) return bad_arg("select_points: negative size");
    if (c == 0 || n_out == 0) return 0;
    if ((!sem_logits && !labels) || !xyz || !out_xyz || (cf > 0 && (!features || !out_feat))) return bad_arg("select_points: null pointer");
    if (sem_logits && ncls < 1) return bad_arg("select_points: ncls < 1");
    select_points_kernel<<<c, SEL_THREADS, 0, (cudaStream_t)stream>>>(n, ncls, cf, target, n_out, sem_logits, labels, xyz,
                                                                     cf > 0 ? features : nullptr, out_xyz, cf > 0 ? out_feat : nullptr, out_count);
    return finish_launch("g4d select_points");
}

// One positional-encoding unit (mesh_encoder.py:450-466).  xyz (b,n,3), new_xyz (b,p,3), feat_pm (b,n,c) fp32 POINT-major or NULL
// (c = 0), idx (b,p,nsample) from the ball query (rows without a hit are all zero, as the reference leaves them), w1t (3+c, 32) =
// Linear1.weight^T, b1 (32), w2t (32, 32) = Linear2.weight^T, b2 (32) -> out (b,p,32) fp32, argmax (b,p,32) int8 (may be NULL):
// the sample that holds each channel's maximum.  nsample in {4, 8, 16, 32}.
G4D_API int g4d_pe_mlp_max(int b, int n, int p, int c, int nsample, const float* xyz, const float* new_xyz, const float* feat_pm,
                           const int* idx, const float* w1t, const float* b1, const float* w2t, const float* b2, float* out,
                           signed char* argmax, void* stream) {
    if (b < 0 || n <= 0 || p < 0 || c < 0) return bad_arg("pe_mlp_max: bad size");
    if (b == 0 || p == 0) return 0;
    if (!xyz || !new_xyz || !idx || !w1t || !b1 || !w2t || !b2 || !out || (c > 0 && !feat_pm)) return bad_arg("pe_mlp_max: null pointer");
    if (b > 65535) return bad_arg("pe_mlp_max: b > 65535");
    const size_t smem = ((size_t)(3 + c) * PE_H + PE_H * PE_H + 2 * PE_H) * sizeof(float);
    if (smem > 200 * 1024) return bad_arg("pe_mlp_max: 3 + c too large for the shared-memory weight image");
    cudaStream_t s = (cudaStream_t)stream;
    int gx = (p + 7) / 8;                                 // 8 warps per CTA, >= 1 centroid per warp
    if (gx > 148 * 4) gx = 148 * 4;
    dim3 grid(gx, b);
#define G4D_PE(KK) { \
        cudaError_t e = cudaFuncSetAttribute(pe_mlp_max_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { set_error("pe_mlp_max: shared memory opt-in: %s", cudaGetErrorString(e)); return (int)e; } \
        pe_mlp_max_kernel<KK><<<grid, 256, smem, s>>>(n, p, c, xyz, new_xyz, c > 0 ? feat_pm : nullptr, idx, w1t, b1, w2t, b2, out, argmax); }
    switch (nsample) {
        case 4: G4D_PE(4) break;
        case 8: G4D_PE(8) break;
        case 16: G4D_PE(16) break;
        case 32: G4D_PE(32) break;
        default: return bad_arg("pe_mlp_max: nsample must be 4, 8, 16 or 32");
    }
#undef G4D_PE
    return finish_launch("g4d pe_mlp_max");
}
