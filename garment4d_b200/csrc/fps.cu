// Furthest-point sampling for sm_100a -- replaces the reference's
// furthest_point_sampling_kernel (modules/pointnet2/pointnet2/src/sampling_gpu.cu:93-253)
// and, in its fused form, the gather_operation + two transpose copies that follow it in
// _PointnetSAModuleBase.forward (pointnet2_modules.py:30-35).
//
// Design (one CTA per cloud, state never leaves the SM):
//   * the cloud's coordinates are staged once into shared memory (SoA) and, for N <= 8192,
//     kept in registers together with the running min-distance `temp` (4 regs/point);
//   * each of the m-1 serial steps is: update temp, per-thread max, then ONE block-wide
//     arg-max built from redux.sync (REDUX) warp reductions + a single __syncthreads over a
//     double-buffered 32-slot exchange -- the reference needs a 10-level shared-memory tree
//     with 11 barriers per step;
//   * bit-exact tie-breaking: the reference's per-thread strict '>' and its shared-memory
//     tree make the winner, among points with equal maximal distance, the one with the
//     smallest (bitrev(k mod bs), k div bs), bs = opt_n_threads(N) (cuda_utils.h:10-14).
//     We reduce max over the float bits, then min over that 32-bit key.
//   * distance arithmetic is the exact FMUL/FFMA/FFMA sequence of the reference build.
#include <limits.h>
#include <math.h>
#include "common.cuh"

namespace g4d {

template <int T, int PPT, bool XYZ_REG, bool SIMPLE_KEY>
__global__ void __launch_bounds__(T, 1)
fps_kernel(int n, int m, int lg_bs, const float* __restrict__ xyz_all, float* __restrict__ temp_all,
           int* __restrict__ idx_all, float* __restrict__ new_xyz_all) {
    extern __shared__ float smem_f[];
    float* xs = smem_f;
    float* ys = xs + n;
    float* zs = ys + n;
    __shared__ int slot_v[2][32];
    __shared__ unsigned slot_k[2][32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = T / 32;
    const size_t cloud = blockIdx.x;
    const float* xyz = xyz_all + cloud * (size_t)n * 3;
    float* temp_io = temp_all ? temp_all + cloud * (size_t)n : nullptr;
    int* idx_out = idx_all + cloud * (size_t)m;
    float* new_xyz = new_xyz_all ? new_xyz_all + cloud * (size_t)m * 3 : nullptr;

    // stage the cloud: coalesced scalar loads of the contiguous (n,3) block, scattered to SoA
    for (int e = tid; e < 3 * n; e += T) {
        const float v = __ldg(xyz + e);
        const int k = e / 3, c = e - 3 * k;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    __syncthreads();

    const unsigned himask = lg_bs ? ~((1u << (32 - lg_bs)) - 1u) : 0u;
    const unsigned bs_mask = (1u << lg_bs) - 1u;

    float px[XYZ_REG ? PPT : 1], py[XYZ_REG ? PPT : 1], pz[XYZ_REG ? PPT : 1];
    float temp[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + T * i;
        const bool valid = k < n;
        if (XYZ_REG) {
            px[i] = valid ? xs[k] : 0.f;
            py[i] = valid ? ys[k] : 0.f;
            pz[i] = valid ? zs[k] : 0.f;
        }
        // invalid lanes carry -1: min(d,-1) stays -1 and never equals a block maximum (>= 0)
        temp[i] = valid ? (temp_io ? temp_io[k] : 1e10f) : -1.f;
    }

    float x1 = xs[0], y1 = ys[0], z1 = zs[0];
    if (tid == 0) {
        idx_out[0] = 0;
        if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
    }
    const unsigned brev_t = __brev((unsigned)tid);   // SIMPLE_KEY: bs == T == 1024 -> k mod bs == tid

    for (int j = 1; j < m; ++j) {
        float vmax = -1.f;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float x2, y2, z2;
            if (XYZ_REG) { x2 = px[i]; y2 = py[i]; z2 = pz[i]; }
            else {
                const int k = min(tid + T * i, n - 1);
                x2 = xs[k]; y2 = ys[k]; z2 = zs[k];
            }
            const float d = sqdist_ref(x2 - x1, y2 - y1, z2 - z1);
            const float t = fminf(d, temp[i]);
            temp[i] = t;
            vmax = fmaxf(vmax, t);
        }
        unsigned mk = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            unsigned key;
            if (SIMPLE_KEY) key = brev_t | (unsigned)i;
            else {
                const unsigned k = (unsigned)(tid + T * i);
                key = (__brev(k & bs_mask) & himask) | (k >> lg_bs);
            }
            mk = (temp[i] == vmax) ? min(mk, key) : mk;
        }
        const int vb = __float_as_int(vmax);
        const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
        const unsigned wk = __reduce_min_sync(0xFFFFFFFFu, vb == wv ? mk : 0xFFFFFFFFu);
        const int par = j & 1;
        if (lane == 0) { slot_v[par][warp] = wv; slot_k[par][warp] = wk; }
        __syncthreads();
        const int sv = lane < NW ? slot_v[par][lane] : INT_MIN;
        const unsigned sk = lane < NW ? slot_k[par][lane] : 0xFFFFFFFFu;
        const int bv = __reduce_max_sync(0xFFFFFFFFu, sv);
        const unsigned bk = __reduce_min_sync(0xFFFFFFFFu, sv == bv ? sk : 0xFFFFFFFFu);
        const int old = (int)(((bk & ~himask) << lg_bs) | __brev(bk & himask));
        x1 = xs[old]; y1 = ys[old]; z1 = zs[old];
        if (tid == 0) {
            idx_out[j] = old;
            if (new_xyz) { new_xyz[3 * j] = x1; new_xyz[3 * j + 1] = y1; new_xyz[3 * j + 2] = z1; }
        }
    }

    if (temp_io) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = tid + T * i;
            if (k < n) temp_io[k] = temp[i];
        }
    }
}

// Any N: temp stays in global memory (L2-resident), coordinates are re-read through the
// read-only path.  Same reduction and tie-break as above.
__global__ void __launch_bounds__(1024, 1)
fps_kernel_generic(int n, int m, int lg_bs, const float* __restrict__ xyz_all, float* __restrict__ temp_all,
                   int* __restrict__ idx_all, float* __restrict__ new_xyz_all) {
    __shared__ int slot_v[2][32];
    __shared__ unsigned slot_k[2][32];
    constexpr int T = 1024, NW = 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t cloud = blockIdx.x;
    const float* xyz = xyz_all + cloud * (size_t)n * 3;
    float* temp = temp_all + cloud * (size_t)n;
    int* idx_out = idx_all + cloud * (size_t)m;
    float* new_xyz = new_xyz_all ? new_xyz_all + cloud * (size_t)m * 3 : nullptr;
    const unsigned himask = lg_bs ? ~((1u << (32 - lg_bs)) - 1u) : 0u;
    const unsigned bs_mask = (1u << lg_bs) - 1u;

    float x1 = __ldg(xyz), y1 = __ldg(xyz + 1), z1 = __ldg(xyz + 2);
    if (tid == 0) {
        idx_out[0] = 0;
        if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
    }
    for (int j = 1; j < m; ++j) {
        float vmax = -1.f;
        unsigned mk = 0xFFFFFFFFu;
        for (int k = tid; k < n; k += T) {
            const float d = sqdist_ref(__ldg(xyz + 3 * k) - x1, __ldg(xyz + 3 * k + 1) - y1, __ldg(xyz + 3 * k + 2) - z1);
            const float t = fminf(d, temp[k]);
            temp[k] = t;
            const unsigned key = (__brev((unsigned)k & bs_mask) & himask) | ((unsigned)k >> lg_bs);
            if (t > vmax) { vmax = t; mk = key; }
            else if (t == vmax) mk = min(mk, key);
        }
        const int vb = __float_as_int(vmax);
        const int wv = __reduce_max_sync(0xFFFFFFFFu, vb);
        const unsigned wk = __reduce_min_sync(0xFFFFFFFFu, vb == wv ? mk : 0xFFFFFFFFu);
        const int par = j & 1;
        if (lane == 0) { slot_v[par][warp] = wv; slot_k[par][warp] = wk; }
        __syncthreads();
        const int sv = lane < NW ? slot_v[par][lane] : INT_MIN;
        const unsigned sk = lane < NW ? slot_k[par][lane] : 0xFFFFFFFFu;
        const int bv = __reduce_max_sync(0xFFFFFFFFu, sv);
        const unsigned bk = __reduce_min_sync(0xFFFFFFFFu, sv == bv ? sk : 0xFFFFFFFFu);
        const int old = (int)(((bk & ~himask) << lg_bs) | __brev(bk & himask));
        x1 = __ldg(xyz + 3 * old); y1 = __ldg(xyz + 3 * old + 1); z1 = __ldg(xyz + 3 * old + 2);
        if (tid == 0) {
            idx_out[j] = old;
            if (new_xyz) { new_xyz[3 * j] = x1; new_xyz[3 * j + 1] = y1; new_xyz[3 * j + 2] = z1; }
        }
    }
}

template <int T, int PPT, bool XYZ_REG, bool SIMPLE_KEY>
static int launch_fps(int b, int n, int m, int lg, const float* xyz, float* temp, int* idx, float* new_xyz, cudaStream_t s) {
    auto kern = fps_kernel<T, PPT, XYZ_REG, SIMPLE_KEY>;
    const size_t smem = (size_t)3 * n * sizeof(float);
    if (smem > 32 * 1024) {   // dynamic + static (the exchange slots) must stay under the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("fps: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    }
    kern<<<b, T, smem, s>>>(n, m, lg, xyz, temp, idx, new_xyz);
    return finish_launch("g4d fps kernel");
}

int fps_dispatch(int b, int n, int m, const float* xyz, float* temp, int* idx, float* new_xyz, cudaStream_t s) {
    if (b < 0 || n <= 0 || m < 0) return bad_arg("fps: need b >= 0, n > 0, m >= 0");
    if (b == 0 || m == 0) return 0;   // reference kernel returns at once for m <= 0 (sampling_gpu.cu:100)
    if (!xyz || !idx) return bad_arg("fps: null xyz/idx");
    const int bs = ref_opt_n_threads(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    if (n <= 256)   return launch_fps<256, 1, true, false>(b, n, m, lg, xyz, temp, idx, new_xyz, s);
    if (n <= 512)   return launch_fps<256, 2, true, false>(b, n, m, lg, xyz, temp, idx, new_xyz, s);
    if (n <= 1024)  return launch_fps<256, 4, true, false>(b, n, m, lg, xyz, temp, idx, new_xyz, s);
    if (n <= 2048)  return launch_fps<256, 8, true, false>(b, n, m, lg, xyz, temp, idx, new_xyz, s);
    if (n <= 4096)  return launch_fps<256, 16, true, false>(b, n, m, lg, xyz, temp, idx, new_xyz, s);
    if (n <= 8192)  return launch_fps<1024, 8, true, true>(b, n, m, lg, xyz, temp, idx, new_xyz, s);    // bs == 1024 here
    if (n <= 16384) return launch_fps<1024, 16, false, true>(b, n, m, lg, xyz, temp, idx, new_xyz, s);
    if (!temp) return bad_arg("fps: n > 16384 needs the (b,n) float scratch `temp` (pre-filled with 1e10)");
    fps_kernel_generic<<<b, 1024, 0, s>>>(n, m, lg, xyz, temp, idx, new_xyz);
    return finish_launch("g4d fps generic kernel");
}

}  // namespace g4d

// Drop-in for furthest_point_sampling_kernel_launcher (sampling_gpu.h:26-27 / sampling.cpp:36-46).
// temp (b,n): in = caller's pre-fill (1e10, pointnet2_utils.py:26), out = final min distances, as the reference.
G4D_API int g4d_furthest_point_sampling(int b, int n, int m, const float* xyz, float* temp, int* idx, void* stream) {
    if (!temp) return g4d::bad_arg("g4d_furthest_point_sampling: temp must not be null");
    return g4d::fps_dispatch(b, n, m, xyz, temp, idx, nullptr, (cudaStream_t)stream);
}

// Fused FPS + centroid gather: writes idx (b,m) and new_xyz (b,m,3) in one launch, temp implicit (1e10).
// Replaces furthest_point_sample -> gather_operation -> 2x transpose().contiguous() (pointnet2_modules.py:30-35).
// `scratch` (b,n) floats is only needed (pre-filled with 1e10) when n > 16384.
G4D_API int g4d_fps_gather(int b, int n, int m, const float* xyz, int* idx, float* new_xyz, float* scratch, void* stream) {
    return g4d::fps_dispatch(b, n, m, xyz, n > 16384 ? scratch : nullptr, idx, new_xyz, (cudaStream_t)stream);
}
