// Micro-benchmark: sustained tcgen05.mma rate of ONE CTA for the operand layouts the kernels use.
//   kind::f16, M = 128, N = 256 (or 128 / 64), K = 16 per instruction, operands in shared memory in the canonical NO-SWIZZLE K-major
//   layout ([k/8][row][8 halves], LBO = rows * 16 B, SBO = 128 B), walking a ring of `nstage` stages like fp_mlp2.cu does.
//   A second mode streams bulk copies into the ring concurrently (the TMA write traffic of a real K loop).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../../garment4d_b200/csrc/umma.cuh"
using namespace g4d;
namespace g4d { void set_error(const char*, ...) {} void count_launches(int) {} int sm_count() { return 148; } }

template <int N, int KSTEPS>
__global__ void __launch_bounds__(128) kmma(long long* out, const unsigned char* src, int nmma, int nstage, int copy) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t A_BYTES = 128 * 16 * 2 * KSTEPS, B_BYTES = N * 16 * 2 * KSTEPS, ST = A_BYTES + B_BYTES;
    const uint32_t base = smem_u32(smem);
    const uint32_t bar0 = base + nstage * ST, bar_d = bar0, slot = bar0 + 8, bar_c = bar0 + 16;
    for (uint32_t i = tid; i < nstage * ST / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(bar_d, 1); mbar_init(bar_c, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + nstage * ST + 8);
    if (warp == 1) {
        const uint32_t idesc = umma_idesc(128, N);
        long long t0 = 0;
        if (elect_one_sync()) {
            t0 = clock64();
            int st = 0;
            for (int i = 0; i < nmma; i += KSTEPS) {
                const uint32_t sa = base + st * ST, sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                    umma_f16(tmem, desc64(desc_lo(sa + k * (2 * 128 * 16), 128 * 16)), desc64(desc_lo(sb + k * (2 * N * 16), N * 16)), idesc, 1);
                if (++st == nstage) st = 0;
            }
            umma_commit(bar_d);
        }
        __syncwarp();
        mbar_wait(bar_d, 0);
        if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
    } else if (warp == 2 && copy) {
        // concurrent bulk copies into the ring (no synchronisation with the MMAs: garbage in, garbage out -- bandwidth only)
        if (elect_one_sync()) {
            int st = 0;
            const int ncopy = nmma / KSTEPS;
            for (int i = 0; i < ncopy; ++i) {
                mbar_expect_tx(bar_c, ST);
                bulk_g2s(base + st * ST, src + (size_t)((blockIdx.x * 7 + i) % 64) * ST, ST, bar_c);
                mbar_wait(bar_c, i & 1);
                if (++st == nstage) st = 0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int KSTEPS> void run(const char* name, long long* d, const unsigned char* src, int nstage, int copy, int grid) {
    const int nmma = 4096;
    const size_t smem = (size_t)nstage * (128 * 32 * KSTEPS + N * 32 * KSTEPS) + 64;
    cudaFuncSetAttribute(kmma<N, KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[148] = {0};
    for (int rep = 0; rep < 2; ++rep) { kmma<N, KSTEPS><<<grid, 128, smem>>>(d, src, nmma, nstage, copy); cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost); }
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = (double)mx / nmma, flop = 2.0 * 128 * N * 16;
    printf("%-58s %7.1f cycles / MMA  = %6.0f FLOP/cycle/SM  (%d CTAs, %s)\n", name, cyc, flop / cyc, grid, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* d; cudaMalloc(&d, 8 * 148);
    unsigned char* src; cudaMalloc(&src, 64 * 49152); cudaMemset(src, 0, 64 * 49152);
    run<256, 2>("N=256, ring of 1 stage (same operands every time)", d, src, 1, 0, 1);
    run<256, 2>("N=256, ring of 6 stages of 24 KB", d, src, 6, 0, 1);
    run<256, 2>("N=256, ring of 6 stages, all 148 SMs", d, src, 6, 0, 148);
    run<256, 2>("N=256, ring of 6 stages + concurrent bulk copies, 148 SMs", d, src, 6, 1, 148);
    run<128, 2>("N=128, ring of 6 stages", d, src, 6, 0, 1);
    run<128, 2>("N=128, ring of 6 stages + concurrent bulk copies, 148 SMs", d, src, 6, 1, 148);
    run<64, 2>("N=64, ring of 6 stages", d, src, 6, 0, 1);
    return 0;
}
