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
