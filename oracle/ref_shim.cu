// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// extern "C" shim over the reference's own kernel launchers.  The four
// reference .cu files are compiled UNMODIFIED from /root/reference (see
// build_ref.sh); this file only re-declares the launcher prototypes that the
// reference headers give (sampling_gpu.h:12-27, ball_query_gpu.h:12-13,
// group_points_gpu.h:13-20, interpolate_gpu.h:13-28) so that the torch
// headers those files drag in are not needed here, and forwards to them.
// It replaces the reference's .cpp wrappers (ball_query.cpp, group_points.cpp,
// interpolate.cpp, sampling.cpp), which do not build against torch 2.11
// (THC/THC.h is gone).
#include <cuda_runtime.h>

// C++ linkage, exactly as in the reference headers.
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints,
    const float *points, const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
    const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void furthest_point_sampling_kernel_launcher(int b, int n, int m,
    const float *dataset, float *temp, int *idxs, cudaStream_t stream);
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
    const float *xyz, const float *new_xyz, int *idx, cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
    const float *points, const int *idx, float *out, cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
    const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown,
    const float *known, float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n,
    const float *points, const int *idx, const float *weight, float *out, cudaStream_t stream);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float *grad_out,
    const int *idx, const float *weight, float *grad_points, cudaStream_t stream);

#define REF_RET() return (int)cudaGetLastError()

extern "C" {

int ref_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream) {
    furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, (cudaStream_t)stream); REF_RET();
}
int ref_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, void *stream) {
    gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out, (cudaStream_t)stream); REF_RET();
}
int ref_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points, void *stream) {
    gather_points_grad_kernel_launcher_fast(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream); REF_RET();
}
// Argument order follows the reference call site ball_query.cpp:23 (new_xyz, then xyz).
int ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, void *stream) {
    ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)stream); REF_RET();
}
int ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out, void *stream) {
    group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out, (cudaStream_t)stream); REF_RET();
}
int ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points, void *stream) {
    group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, grad_out, idx, grad_points, (cudaStream_t)stream); REF_RET();
}
int ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *stream) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)stream); REF_RET();
}
int ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, void *stream) {
    three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out, (cudaStream_t)stream); REF_RET();
}
int ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, void *stream) {
    three_interpolate_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, weight, grad_points, (cudaStream_t)stream); REF_RET();
}

}  // extern "C"
