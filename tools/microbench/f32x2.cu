// Micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters) {
    float a[8], b = threadIdx.x * 1e-3f, c = 0.5f;
    for (int i = 0; i < 8; ++i) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __fmaf_rn(a[i], b, c);
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters) {
    unsigned long long a[4];
    float b = threadIdx.x * 1e-3f, c = 0.5f;
    unsigned long long bb, cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    for (int i = 0; i < 4; ++i) { float x = 2 * i + threadIdx.x, y = 2 * i + 1 + threadIdx.x; asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(x), "f"(y)); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb), "l"(cc));
    float s = 0;
    for (int i = 0; i < 4; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<148 * 4, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("scalar FFMA : %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 8 * iters * 148 * 4 * 1024 / ms / 1e9);
        cudaEventRecord(e0); k_packed<<<148 * 4, 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 8 * iters * 148 * 4 * 1024 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
