// Shared helpers for the sm_100a kernels of the Garment4D hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#define G4D_API extern "C" __attribute__((visibility("default")))

namespace g4d {

// last error text, readable through g4d_last_error()
void set_error(const char* fmt, ...);

// kernels launched through this library since load (g4d_launch_count); bench.py reports the per-step delta
void count_launches(int n);

inline int finish_launch(const char* what, int nkernels = 1) {
    count_launches(nkernels);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

inline int bad_arg(const char* what) {
    set_error("%s", what);
    return (int)cudaErrorInvalidValue;
}

// The reference's block-size rule (cuda_utils.h:10-14): largest power of two <= n,
// through the same double log ratio, clamped to [1, 1024].  It fixes the FPS tie-break.
inline int ref_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

int sm_count();

// Squared distance with the exact operation order nvcc 12.9 -O2 emits for the reference
// kernels' `dx*dx + dy*dy + dz*dz` (FMUL on y, FFMA on x, FFMA on z; checked in SASS).
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

}  // namespace g4d
