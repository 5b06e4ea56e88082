// Deterministic segmented reduction for the backward kernels (sm_100a).
//
// The reference's three backward kernels are atomicAdd scatters -- group_points_grad (group_points_gpu.cu:8-25),
// gather_points_grad (sampling_gpu.cu:46-63), three_interpolate_grad (interpolate_gpu.cu:120-142) -- so its gradients depend on
// the order in which the atomics land (fp32 addition is not associative) and differ from run to run.  All three are the same
// operation:   out[b, c, dst[b, e]] += w[b, e] * grad[b, c, e]      (w = 1 for group / gather, the 3 interpolation weights)
// Here it runs as a SEGMENTED reduction in a fixed order:
//   1. csr_count / csr_scan      per cloud: how many sources hit each destination, exclusive scan -> segment starts
//   2. csr_fill_stable           ONE warp per cloud walks the sources in order, 32 at a time; __match_any_sync groups equal
//                                destinations, the rank inside the chunk is a popcount and a per-destination counter carries the
//                                rank across chunks: a stable counting sort, no atomics -> every segment lists its sources in
//                                ascending order
//   3. seg_reduce                one thread per (destination, channel slab) adds its segment front to back
// The result is bit-identical from run to run and equal to the serial loop of the CPU oracle (ascending source order).
// The index structure (steps 1-2) depends on `dst` only: it is built once per index tensor and reused for every channel slab.
#include "common.cuh"

namespace g4d {

constexpr int SD_THREADS = 256;

__global__ void __launch_bounds__(SD_THREADS)
csr_count_kernel(int n_src, int n_dst, const int* __restrict__ dst_all, int* __restrict__ count_all) {
    const size_t b = blockIdx.y;
    const int e = blockIdx.x * SD_THREADS + threadIdx.x;
    if (e >= n_src) return;
    const int d = __ldg(dst_all + b * n_src + e);
    if ((unsigned)d < (unsigned)n_dst) atomicAdd(count_all + b * (size_t)(n_dst + 1) + d, 1);     // integer atomics: order-independent
}

// exclusive scan of count[0..n_dst) in place -> start[0..n_dst]; one CTA per cloud
__global__ void __launch_bounds__(SD_THREADS)
csr_scan_kernel(int n_dst, int* __restrict__ count_all) {
    __shared__ int warp_tot[SD_THREADS / 32];
    __shared__ int carry_s;
    int* cnt = count_all + blockIdx.x * (size_t)(n_dst + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_dst; base += SD_THREADS) {
        const int i = base + tid;
        const int v = i < n_dst ? cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = carry_s;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (i < n_dst) cnt[i] = before + incl - v;
        __syncthreads();
        if (tid == SD_THREADS - 1) carry_s = before + incl;
        __syncthreads();
    }
    if (tid == 0) cnt[n_dst] = carry_s;
}

// stable placement: entries[start[d] + rank] = e, rank = number of earlier sources with the same destination.
// One warp per cloud; cursor (n_dst ints per cloud, zeroed) carries the ranks across the 32-source chunks.
__global__ void __launch_bounds__(32)
csr_fill_stable_kernel(int n_src, int n_dst, const int* __restrict__ dst_all, const int* __restrict__ start_all,
                       int* __restrict__ cursor_all, int* __restrict__ entries_all) {
    const size_t b = blockIdx.x;
    const int lane = threadIdx.x;
    const int* dst = dst_all + b * n_src;
    const int* start = start_all + b * (size_t)(n_dst + 1);
    int* cursor = cursor_all + b * (size_t)n_dst;
    int* entries = entries_all + b * (size_t)n_src;
    for (int e0 = 0; e0 < n_src; e0 += 32) {
        const int e = e0 + lane;
        const bool live = e < n_src;
        int d = live ? __ldg(dst + e) : -1;
        if ((unsigned)d >= (unsigned)n_dst) d = -1;                  // out-of-range destinations are dropped
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (d >= 0) {
            const int base = cursor[d];                              // written by this warp only, in earlier iterations
            entries[__ldg(start + d) + base + rank] = e;
            __syncwarp(peers);
            if (rank == 0) cursor[d] = base + __popc(peers);         // the group's first lane advances the counter
        }
        __syncwarp();
    }
}

// out[b, ch, d] = sum over the segment of d, in ascending source order, of w[e] * grad[b, ch, e]
__global__ void __launch_bounds__(SD_THREADS)
seg_reduce_kernel(int c, int n_src, int n_dst, int grad_div, const int* __restrict__ start_all, const int* __restrict__ entries_all,
                  const float* __restrict__ weight_all, const float* __restrict__ grad_all, float* __restrict__ out_all) {
    const size_t b = blockIdx.z;
    const int d = blockIdx.x * SD_THREADS + threadIdx.x;
    if (d >= n_dst) return;
    const int* start = start_all + b * (size_t)(n_dst + 1);
    const int* entries = entries_all + b * (size_t)n_src;
    const float* w = weight_all ? weight_all + b * (size_t)n_src : nullptr;
    const int s0 = __ldg(start + d), s1 = __ldg(start + d + 1);
    for (int ch = blockIdx.y; ch < c; ch += gridDim.y) {
        const float* g = grad_all + (b * c + ch) * (size_t)(n_src / grad_div);
        float acc = 0.f;
        for (int j = s0; j < s1; ++j) {
            const int e = __ldg(entries + j);
            const float gv = __ldg(g + (grad_div == 1 ? e : e / grad_div));
            acc = __fadd_rn(acc, w ? __fmul_rn(gv, __ldg(w + e)) : gv);      // product rounded, then added: the reference's atomicAdd(.., g * w)
        }
        out_all[(b * c + ch) * (size_t)n_dst + d] = acc;
    }
}

}  // namespace g4d

using namespace g4d;

// Workspace: per cloud (n_dst + 1) segment starts + n_dst cursors + n_src entries, int32.
G4D_API size_t g4d_scatter_det_workspace_bytes(int b, int n_src, int n_dst) {
    if (b < 0 || n_src < 0 || n_dst < 0) return 0;
    return (size_t)b * ((size_t)(n_dst + 1) + (size_t)n_dst + (size_t)n_src) * 4;
}

// Builds the index structure for dst (b, n_src) int32 with values in [0, n_dst) into workspace (zeroed here).
G4D_API int g4d_scatter_det_build(int b, int n_src, int n_dst, const int* dst, void* workspace, void* stream) {
    if (b < 0 || n_src < 0 || n_dst < 0) return bad_arg("scatter_det_build: negative size");
    if (b == 0 || n_dst == 0) return 0;
    if (!dst || !workspace) return bad_arg("scatter_det_build: null pointer");
    if (b > 65535) return bad_arg("scatter_det_build: b > 65535");
    cudaStream_t s = (cudaStream_t)stream;
    int* start = (int*)workspace;
    int* cursor = start + (size_t)b * (n_dst + 1);
    int* entries = cursor + (size_t)b * n_dst;
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)b * ((size_t)(n_dst + 1) + (size_t)n_dst) * 4, s);
    if (e != cudaSuccess) { set_error("scatter_det_build: memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (n_src > 0) {
        dim3 grid((n_src + SD_THREADS - 1) / SD_THREADS, b);
        csr_count_kernel<<<grid, SD_THREADS, 0, s>>>(n_src, n_dst, dst, start);
    }
    csr_scan_kernel<<<b, SD_THREADS, 0, s>>>(n_dst, start);
    if (n_src > 0) csr_fill_stable_kernel<<<b, 32, 0, s>>>(n_src, n_dst, dst, start, cursor, entries);
    return finish_launch("g4d scatter_det_build", 3);
}

// out (b, c, n_dst) = deterministic sum of weight (b, n_src) [NULL = 1] * grad (b, c, n_src / grad_div)[e / grad_div] over the
// segments of `workspace` (built by g4d_scatter_det_build for the same b, n_src, n_dst).  grad_div = 3 serves three_interpolate
// (three sources per interpolated point share its gradient), 1 otherwise.  Every element of out is written (0 for empty segments).
G4D_API int g4d_scatter_det_apply(int b, int c, int n_src, int n_dst, int grad_div, const void* workspace, const float* weight,
                                  const float* grad, float* out, void* stream) {
    if (b < 0 || c < 0 || n_src < 0 || n_dst < 0) return bad_arg("scatter_det_apply: negative size");
    if (b == 0 || c == 0 || n_dst == 0) return 0;
    if (!workspace || !out || (n_src > 0 && !grad)) return bad_arg("scatter_det_apply: null pointer");
    if (b > 65535) return bad_arg("scatter_det_apply: b > 65535");
    if (grad_div < 1 || n_src % grad_div) return bad_arg("scatter_det_apply: grad_div must divide n_src");
    const int* start = (const int*)workspace;
    const int* entries = start + (size_t)b * (n_dst + 1) + (size_t)b * n_dst;
    int slabs = c < 16 ? c : 16;
    dim3 grid((n_dst + SD_THREADS - 1) / SD_THREADS, slabs, b);
    seg_reduce_kernel<<<grid, SD_THREADS, 0, (cudaStream_t)stream>>>(c, n_src, n_dst, grad_div, start, entries, weight, grad, out);
    return finish_launch("g4d scatter_det_apply");
}
