// Micro-benchmark: round-trip latency of the MMA-issuer <-> epilogue hand-off used by sa_mlp.cu / fp_head.cu.
#include <cstdio>
#include "../../garment4d_b200/csrc/umma.cuh"
using namespace g4d;
namespace g4d { void set_error(const char*, ...) {} void count_launches(int) {} int sm_count() { return 148; } }
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// issue loop cost: NM MMAs, optionally a commit after each (to a dummy barrier) and a try_wait on a completed barrier before each
template <int NM, int COMMIT_EACH, int WAIT_EACH, int FENCE_EACH>
__global__ void __launch_bounds__(160) kissue(long long* out, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_d = smem_u32(smem + 65536), bar_x = bar_d + 8, bar_done = bar_d + 24, slot = bar_d + 16;
    for (int i = tid; i < 16384; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(bar_d, 1); mbar_init(bar_x, 1); mbar_init(bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(slot, 128);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 16);
    if (tid == 128) {
        const uint32_t idesc = umma_idesc(128, 32);
        long long t0 = clock64();
        uint32_t nx = 0;
        for (int i = 0; i < iters; ++i) {
            for (int m = 0; m < NM; ++m) {
                if (WAIT_EACH) mbar_wait(bar_done, 1);      // parity 1 on a fresh barrier: already complete
                if (FENCE_EACH) tc_fence_after();
                const uint64_t ad = umma_desc(smem_u32(smem) + m * 64, 128 * 16, 128), bd = umma_desc(smem_u32(smem) + 32768 + m * 64, 32 * 16, 128);
                umma_f16(tmem, ad, bd, idesc, m > 0);
                if (COMMIT_EACH) { umma_commit(bar_x); }
            }
            umma_commit(bar_d);
            mbar_wait(bar_d, i & 1);
            if (COMMIT_EACH) { nx += NM; }
        }
        out[0] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}
template <int NM, int C, int W, int F> void runi(const char* name, long long* d) {
    cudaFuncSetAttribute(kissue<NM, C, W, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) { kissue<NM, C, W, F><<<1, 160, 66000>>>(d, 1000); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); }
    printf("%-52s %8.1f cycles / iteration (%d MMAs)  (%s)\n", name, h / 1000.0, NM, cudaGetErrorString(cudaGetLastError()));
}

template <int LD, int ST, int FENCE, int NMMA>
__global__ void __launch_bounds__(160) k(long long* out, int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_d = smem_u32(smem + 65536), bar_e = bar_d + 8, slot = bar_d + 16;
    for (int i = tid; i < 16384; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(bar_d, 1); mbar_init(bar_e, 4); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(slot, 128);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 16);
    if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, 32);
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                mbar_wait(bar_e, (i + 1) & 1);
                tc_fence_after();
                for (int m = 0; m < NMMA; ++m) {
                    const uint64_t ad = umma_desc(smem_u32(smem), 128 * 16, 128), bd = umma_desc(smem_u32(smem) + 32768, 32 * 16, 128);
                    umma_f16(tmem, ad, bd, idesc, m > 0);
                }
                umma_commit(bar_d);
            }
            mbar_wait(bar_e, (iters + 1) & 1);
            out[0] = clock64() - t0;
        }
        __syncwarp();
    } else {
        float acc = 0.f;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bar_d, i & 1);
            tc_fence_after();
            if (LD) { float v[16]; tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v); acc += v[0]; }
            if (ST) reinterpret_cast<uint4*>(smem)[tid] = make_uint4(i, 0, 0, 0);
            tc_fence_before();
            if (FENCE) fence_proxy_async();
            __syncwarp();
            if (lane == 0) arrive(bar_e);
        }
        if (acc == 123.f) out[1] = 1;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}
template <int LD, int ST, int FENCE, int NMMA> void run(const char* name, long long* d) {
    cudaFuncSetAttribute(k<LD, ST, FENCE, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000);
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) { k<LD, ST, FENCE, NMMA><<<1, 160, 66000>>>(d, 2000); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); }
    printf("%-44s %8.1f cycles / round trip   (%s)\n", name, h / 2000.0, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    run<0, 0, 0, 1>("commit/wait/arrive only, 1 MMA", d);
    run<1, 0, 0, 1>("+ tmem ld16", d);
    run<1, 1, 0, 1>("+ st.shared", d);
    run<1, 1, 1, 1>("+ fence.proxy.async", d);
    run<1, 1, 1, 8>("+ 8 MMAs", d);
    run<0, 0, 0, 0>("no MMA at all (commit only)", d);
    runi<1, 0, 0, 0>("issue: 1 MMA + commit + wait", d);
    runi<13, 0, 0, 0>("issue: 13 MMAs, one commit", d);
    runi<13, 1, 0, 0>("issue: 13 MMAs, commit after each", d);
    runi<13, 0, 1, 0>("issue: 13 MMAs, try_wait(complete) before each", d);
    runi<13, 0, 0, 1>("issue: 13 MMAs, tcgen05.fence before each", d);
    runi<13, 1, 1, 1>("issue: 13 MMAs, wait + fence + commit each", d);
    return 0;
}
