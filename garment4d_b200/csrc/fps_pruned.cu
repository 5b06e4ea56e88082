// Furthest-point sampling with exact spatial pruning (sm_100a) -- same indices as the reference kernel
// (sampling_gpu.cu:93-209), bit for bit, but most of the N min-distance updates of each serial step are skipped.
//
// Observation: a step changes temp[k] = min(temp[k], d(k, last)) only for points closer to the new sample than their
// current temp.  Threads own PPT CONSECUTIVE points of the cell-sorted order produced by g4d_grid_build, i.e. a compact
// clump with bounding sphere (c, rad).  If |last - c| >= rad + sqrt(max temp of the clump) (with a 2e-4 safety margin
// that dwarfs fp32 rounding), none of the clump's temps can change, the thread's cached (max temp, tie-break key,
// coordinates of its candidate) stay valid and the thread pays only the 7-instruction test.  Warps whose points are all
// far away skip the update entirely (measured on body scans: 3 % of threads, 20 % of warps update per step).
// The block-wide arg-max (4 redux.sync + 1 barrier; tie-break = smallest (bitrev(k mod bs), k div bs) among equal
// maxima, see fps.cu) runs on the cached per-thread candidates.
//
// The serial chain (update -> arg-max -> barrier -> next sample) is latency-bound, so the kernel is shaped for TWO
// resident CTAs per SM (512 threads x 16 points, <= 64 registers, 112 KB of SoA coordinates + indices in shared memory each):
// while one cloud waits on its barrier the other issues, and 240 clouds fit 148 SMs in one wave.
// Registers hold only the temps and the cached candidate; coordinates are read from shared memory by the (rare)
// updates, as is the original index (u16) that gives the tie-break key of the clump's maximal point.
#include <limits.h>
#include <stdlib.h>
#include "common.cuh"
#include "grid.cuh"

namespace g4d {

// PROF: clock() phase sums of every warp of cloud 0 (tools/fps_phases.py); a measurement build, never launched by the product path.
__device__ unsigned g_fps_prof[32 * 20];
__device__ __forceinline__ unsigned clk() { unsigned c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); return c; }

template <int T, int PPT, bool PROF = false>
__global__ void __launch_bounds__(T, (T <= 512 && !PROF) ? 2 : 1)
fps_pruned_kernel(int n, int m, int lg_bs, const float4* __restrict__ sorted_all, long long cloud_stride, int* __restrict__ idx_all,
                  float* __restrict__ new_xyz_all) {
    extern __shared__ __align__(16) float soa[];           // xs[PPT][T], ys[PPT][T], zs[PPT][T]
    float* xs = soa;
    float* ys = xs + PPT * T;
    float* zs = ys + PPT * T;
    unsigned short* ks = reinterpret_cast<unsigned short*>(zs + PPT * T);   // original index of every point (n <= 8192)
    constexpr int NW = T / 32;
    __shared__ int slot_v[2][NW];
    __shared__ int slot_t[2][NW];                           // where the warp's candidate lives in xs/ys/zs/ks
    __shared__ float first_xyz[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t cloud = blockIdx.x;
    const float4* sorted = sorted_all + cloud * cloud_stride;      // any spatially coherent order of the cloud: (x, y, z, bits(k))
    int* idx_out = idx_all + cloud * (size_t)m;
    float* new_xyz = new_xyz_all ? new_xyz_all + cloud * (size_t)m * 3 : nullptr;
    const unsigned himask = lg_bs ? ~((1u << (32 - lg_bs)) - 1u) : 0u;
    const unsigned bs_mask = (1u << lg_bs) - 1u;

    // ---- load this thread's clump, its bounding sphere; locate point 0 (the first sample) ----
    float temp[PPT];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int pos = tid * PPT + i;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        temp[i] = -1.f;                                    // empty slot: never a maximum, min(d, -1) stays -1
        if (pos < n) {
            p = __ldg(sorted + pos);
            if (__float_as_int(p.w) == 0) { first_xyz[0] = p.x; first_xyz[1] = p.y; first_xyz[2] = p.z; }
            temp[i] = 1e10f;                               // the reference's pre-fill (pointnet2_utils.py:26)
            lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
            lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
            lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        }
        xs[i * T + tid] = p.x; ys[i * T + tid] = p.y; zs[i * T + tid] = p.z;
        ks[i * T + tid] = (unsigned short)__float_as_int(p.w);
    }
    const bool has_pts = tid * PPT < n;
    const float ccx = 0.5f * (lo[0] + hi[0]), ccy = 0.5f * (lo[1] + hi[1]), ccz = 0.5f * (lo[2] + hi[2]);
    float rad = 0.f;
    if (has_pts) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const float ex = xs[i * T + tid] - ccx, ey = ys[i * T + tid] - ccy, ez = zs[i * T + tid] - ccz;
            if (tid * PPT + i < n) rad = fmaxf(rad, ex * ex + ey * ey + ez * ez);
        }
        rad = sqrtf(rad) * 1.0001f;
    }
    __syncthreads();
    float x1 = first_xyz[0], y1 = first_xyz[1], z1 = first_xyz[2];
    if (tid == 0) {
        idx_out[0] = 0;
        if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
    }

    // cached per-thread candidate and per-warp reduction
    float tmax = has_pts ? 1e10f : -1.f, thr = has_pts ? INFINITY : -1.f;
    int tpos = tid;                                        // position (slot * T + tid) of this thread's candidate
    int vb = __float_as_int(tmax), wv = 0;
    bool holder = false, stale_thr = false;
    int out_i = 0;
    float out_x = 0.f, out_y = 0.f, out_z = 0.f;
    // tie-break key of the point at a position: only needed when equal maxima meet (rare), so it is computed on demand
    auto key_at = [&](int pos) -> unsigned {
        const unsigned k = ks[pos];
        return (__brev(k & bs_mask) & himask) | (k >> lg_bs);
    };

    unsigned pf[2][8] = {}, pn[2] = {0, 0}, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
    for (int j = 1; j < m; ++j) {
        if (PROF) c0 = clk();
        const float dcx = ccx - x1, dcy = ccy - y1, dcz = ccz - z1;
        const float d2c = dcx * dcx + dcy * dcy + dcz * dcz;
        const bool need = d2c < thr;
        const bool wneed = __any_sync(0xFFFFFFFFu, need) || j == 1;
        if (PROF) c1 = c2 = c3 = clk();
        if (wneed) {
            if (need) {
#pragma unroll
                for (int i = 0; i < PPT; ++i) {
                    const float d = sqdist_ref(xs[i * T + tid] - x1, ys[i * T + tid] - y1, zs[i * T + tid] - z1);
                    temp[i] = fminf(d, temp[i]);
                }
                // tree-shaped max and equality mask (short dependency chains: this path is on the critical path)
                float mx[PPT];
#pragma unroll
                for (int i = 0; i < PPT; ++i) mx[i] = temp[i];
#pragma unroll
                for (int w = PPT / 2; w >= 1; w >>= 1)
#pragma unroll
                    for (int i = 0; i < w; ++i) mx[i] = fmaxf(mx[i], mx[i + w]);
                const float vmax = mx[0];
                unsigned em[PPT];
#pragma unroll
                for (int i = 0; i < PPT; ++i) em[i] = (temp[i] == vmax) ? (1u << i) : 0u;
#pragma unroll
                for (int w = PPT / 2; w >= 1; w >>= 1)
#pragma unroll
                    for (int i = 0; i < w; ++i) em[i] |= em[i + w];
                unsigned eq = em[0];
                // candidate = the clump's maximal point; several equal maxima (rare): the smallest tie-break key
                int bi = __ffs(eq) - 1;
                if (eq & (eq - 1)) {
                    unsigned mk = 0xFFFFFFFFu;
                    while (eq) {
                        const int i = __ffs(eq) - 1;
                        eq &= eq - 1;
                        const unsigned key = key_at(i * T + tid);
                        if (key < mk) { mk = key; bi = i; }
                    }
                }
                tpos = bi * T + tid;
                tmax = vmax;
                vb = __float_as_int(tmax);
                stale_thr = true;
            }
            if (PROF) { __syncwarp(); c2 = clk(); }
            // warp candidate: max value; the tie-break reduction only runs when several lanes share the maximum
            wv = __reduce_max_sync(0xFFFFFFFFu, vb);
            const unsigned m1 = __ballot_sync(0xFFFFFFFFu, vb == wv);
            if (__popc(m1) == 1) holder = (vb == wv);
            else {
                const unsigned tkey = vb == wv ? key_at(tpos) : 0xFFFFFFFFu;
                const unsigned wk = __reduce_min_sync(0xFFFFFFFFu, tkey);
                holder = (vb == wv) && (tkey == wk);
            }
            if (PROF) c3 = clk();
        }
        const int par = j & 1;
        if (holder) { slot_v[par][warp] = vb; slot_t[par][warp] = tpos; }      // cached between updates of this warp
        if (PROF) c4 = clk();
        __syncthreads();
        if (PROF) c5 = clk();
        const int sv = lane < NW ? slot_v[par][lane] : INT_MIN;
        const int bv = __reduce_max_sync(0xFFFFFFFFu, sv);
        unsigned m2 = __ballot_sync(0xFFFFFFFFu, sv == bv);
        if (__popc(m2) > 1) {                              // equal maxima in several warps: smallest key wins
            const unsigned sk = sv == bv ? key_at(slot_t[par][lane]) : 0xFFFFFFFFu;
            const unsigned bk2 = __reduce_min_sync(0xFFFFFFFFu, sk);
            m2 = __ballot_sync(0xFFFFFFFFu, sv == bv && sk == bk2);
        }
        const int wl = __ffs(m2) - 1;
        const int wpos = slot_t[par][wl];
        x1 = xs[wpos]; y1 = ys[wpos]; z1 = zs[wpos];
        if (PROF) { c6 = clk() + (__float_as_uint(x1) & 0u); }
        // Results are latched in the lanes of warp 0 and written out 32 steps at a time: a global store in every step
        // would make each barrier wait for its acknowledgement (BAR.SYNC drains the warp's outstanding stores).
        if (warp == 0) {
            if (lane == (j & 31)) { out_i = ks[wpos]; out_x = x1; out_y = y1; out_z = z1; }
            if ((j & 31) == 31 || j == m - 1) {
                const int jj = (j & ~31) + lane;
                if (jj >= 1 && jj <= j) {
                    idx_out[jj] = out_i;
                    if (new_xyz) { new_xyz[3 * jj] = out_x; new_xyz[3 * jj + 1] = out_y; new_xyz[3 * jj + 2] = out_z; }
                }
            }
        }
        if (stale_thr) {                                   // off the pre-barrier critical path
            const float s = rad + sqrtf(fmaxf(tmax, 0.f));
            thr = s * s * 1.0002f;
            stale_thr = false;
        }
        if (PROF && j > 1) {
            const unsigned dd[7] = {c1 - c0, c2 - c1, c3 - c2, c4 - c3, c5 - c4, c6 - c5, clk() - c6};
#pragma unroll
            for (int i = 0; i < 7; ++i) { pf[0][i] += wneed ? 0u : dd[i]; pf[1][i] += wneed ? dd[i] : 0u; }
            pn[0] += wneed ? 0u : 1u; pn[1] += wneed ? 1u : 0u;
        }
    }
    if (PROF && blockIdx.x == 0 && lane == 0) {
#pragma unroll
        for (int f = 0; f < 2; ++f) {
#pragma unroll
            for (int i = 0; i < 7; ++i) g_fps_prof[warp * 20 + f * 8 + i] = pf[f][i];
            g_fps_prof[warp * 20 + 16 + f] = pn[f];
        }
    }
}

static int g_fps_step_clouds = 0;      // g4d_fps_concurrency_hint

template <int T, int PPT, bool PROF = false>
static int launch_fps_pruned(int b, int n, int m, int lg, const float4* sorted, long long stride, int* idx, float* new_xyz, cudaStream_t s) {
    auto kern = fps_pruned_kernel<T, PPT, PROF>;
    size_t smem = (size_t)T * PPT * (3 * sizeof(float) + sizeof(unsigned short));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("fps_pruned: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    kern<<<b, T, smem, s>>>(n, m, lg, sorted, stride, idx, new_xyz);
    return finish_launch("g4d fps_pruned kernel");
}

// sorted: per cloud n float4 (x, y, z, bits(k)) in a spatially coherent order, clouds `stride` float4 apart.  n <= 8192.
int fps_pruned_sorted(int b, int n, int m, const float4* sorted, long long stride, int* idx, float* new_xyz, cudaStream_t s) {
    const int bs = ref_opt_n_threads(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    if (n <= 1024) return launch_fps_pruned<128, 8>(b, n, m, lg, sorted, stride, idx, new_xyz, s);
    if (n <= 2048) return launch_fps_pruned<128, 16>(b, n, m, lg, sorted, stride, idx, new_xyz, s);
    if (n <= 4096) return launch_fps_pruned<256, 16>(b, n, m, lg, sorted, stride, idx, new_xyz, s);
    // 1024 threads x 8 points per cloud (half the update path per step, but one CTA per SM): 0.65 ms against 0.71 ms for the
    // 512 x 16 form when every cloud of the step has an SM to itself, 1.27 ms against 0.92 ms when they do not (two waves instead
    // of two clouds per SM).  A launch cannot see the launches on other streams, so the caller says how many clouds the whole
    // step holds (g4d_fps_concurrency_hint); without a hint the launch's own count decides.  G4D_FPS_WIDE=0/1 forces the choice.
    static const int wide_env = getenv("G4D_FPS_WIDE") ? atoi(getenv("G4D_FPS_WIDE")) : -1;
    const int step_clouds = g_fps_step_clouds > 0 ? g_fps_step_clouds : b;
    const bool wide = wide_env >= 0 ? wide_env != 0 : step_clouds <= sm_count();
    static const bool prof = getenv("G4D_FPS_PROF") != nullptr;
    if (prof) return launch_fps_pruned<512, 16, true>(b, n, m, lg, sorted, stride, idx, new_xyz, s);
    if (wide) return launch_fps_pruned<1024, 8>(b, n, m, lg, sorted, stride, idx, new_xyz, s);
    return launch_fps_pruned<512, 16>(b, n, m, lg, sorted, stride, idx, new_xyz, s);
}

}  // namespace g4d

using namespace g4d;

// How many clouds the caller's whole step samples concurrently (all streams together); 0 = unknown (each launch decides from its own
// count).  Only steers the choice between the two pruned-FPS kernel shapes: results never depend on it.
G4D_API void g4d_fps_concurrency_hint(int step_clouds) { g_fps_step_clouds = step_clouds > 0 ? step_clouds : 0; }

// Measurement aid: phase sums written by the last G4D_FPS_PROF=1 launch (32 warps x 20 words), see tools/fps_phases.py.
G4D_API int g4d_debug_fps_phases(unsigned* out640) {
    return (int)cudaMemcpyFromSymbol(out640, g_fps_prof, sizeof(unsigned) * 640);
}

// = g4d_fps_gather (same idx and new_xyz) given a grid built over xyz by g4d_grid_build (any cell size): the
// cell-sorted order gives every thread a compact clump of points, which makes the exact pruning effective.
// 1 <= n <= 8192 (larger clouds: g4d_fps_gather).
G4D_API int g4d_fps_gather_grid(int b, int n, int m, const void* grid, int* idx, float* new_xyz, void* stream) {
    if (b < 0 || n <= 0 || m < 0) return bad_arg("fps_gather_grid: need b >= 0, n > 0, m >= 0");
    if (b == 0 || m == 0) return 0;
    if (!grid || !idx) return bad_arg("fps_gather_grid: null pointer");
    if (n > 8192) return bad_arg("fps_gather_grid: n > 8192 (use g4d_fps_gather)");
    return fps_pruned_sorted(b, n, m, grid_sorted((const float*)grid), (long long)(grid_cloud_words(n) / 4), idx, new_xyz, (cudaStream_t)stream);
}
