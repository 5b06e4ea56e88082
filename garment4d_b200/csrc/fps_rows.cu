// Furthest-point sampling with exact spatial pruning, warp-row form (sm_100a) -- same indices as the reference kernel
// (sampling_gpu.cu:93-209), bit for bit.
//
// A serial step of FPS changes temp[k] = min(temp[k], d(k, last)) only for the few points that are closer to the new sample
// than their current temp (44 of 8192 on average on a body scan), and the next sample is the arg-max of temp.  The points are
// first put in MORTON ORDER (fps_morton_kernel: 15-bit codes of a 32^3 grid over the cloud's bounding cube, counting sort in
// shared memory), so that 32 consecutive points form a compact CLUMP with a small bounding sphere (c, rad), and 16 consecutive
// clumps (one warp's share) a compact region.  If |last - c| >= rad + sqrt(max temp of the clump) -- with safety margins that
// dwarf fp32 rounding -- none of the clump's temps can change and its cached maximum stays valid.
//
//   layout    clump g = sorted positions 32g .. 32g+31 belongs to warp g % NW (spatial neighbours go to different warps);
//             clump c of warp w is g = c * NW + w; ONE POINT PER LANE:
//             lane l keeps temp[c] of its point of each clump in registers; lane i < 16 also keeps the record of clump i
//             (centre, radius, threshold, max temp, tie-break key and lane of its maximal point).
//   step      lanes 0..15 test their clump against the new sample (7 instructions), one ballot gives the clumps to update;
//             each flagged clump is updated by all 32 lanes at once (one point each: 3 LDS + 8 ALU, no divergence -- the thread-
//             per-clump form of round 1 ran a 280-instruction path with 3 of 32 lanes active), its new maximum is ONE
//             redux.sync (+ a key reduction only on ties), then the warp's best clump is one more redux over the 16 records.
//   block     arg-max over the warps' candidates exactly as in fps.cu: double-buffered slots, one __syncthreads per step,
//             every warp reduces the 16 slots redundantly; a warp rewrites its slot only for two steps after it changed.
//   ties      winner among equal maxima = smallest (bitrev(k mod bs), k div bs), bs = opt_n_threads(N) (cuda_utils.h:10-14):
//             reduced as max over float bits, then min over that key, at clump, warp and block level (keys are unique).
//
// Emulated on the CPU before it was written (tools/emul/fps_prune_emul.py): on body scans 15 flagged clumps per step in 4 of
// 16 warps (Morton order) against 26 thread-clumps in 6.3 warps for the cell-sorted thread-per-clump kernel.
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace g4d {

constexpr int MORTON_BITS = 5;                         // per axis
constexpr int MORTON_CELLS = 1 << (3 * MORTON_BITS);   // 32768
constexpr int MS_THREADS = 512;

__device__ __forceinline__ unsigned spread5(unsigned v) {      // bits 0..4 -> positions 0,3,6,9,12
    v &= 0x1Fu;
    v = (v | (v << 8)) & 0x100Fu;
    v = (v | (v << 4)) & 0x10C3u;
    v = (v | (v << 2)) & 0x1249u;
    return v;
}

__device__ __forceinline__ unsigned morton_code(float x, float y, float z, float ox, float oy, float oz, float inv_h) {
    const int cx = min(max(__float2int_rd((x - ox) * inv_h), 0), (1 << MORTON_BITS) - 1);      // NaN -> 0
    const int cy = min(max(__float2int_rd((y - oy) * inv_h), 0), (1 << MORTON_BITS) - 1);
    const int cz = min(max(__float2int_rd((z - oz) * inv_h), 0), (1 << MORTON_BITS) - 1);
    return spread5((unsigned)cx) | (spread5((unsigned)cy) << 1) | (spread5((unsigned)cz) << 2);
}

// One CTA per cloud: sorted[pos] = (x, y, z, bits(k)) in Morton order of the cells (order inside a cell: arbitrary).
// Only the ORDER matters to the caller (any permutation gives the same FPS result; a good one makes the pruning effective).
__global__ void __launch_bounds__(MS_THREADS)
fps_morton_kernel(int n, const float* __restrict__ xyz_all, float4* __restrict__ sorted_all) {
    extern __shared__ unsigned hist[];                 // MORTON_CELLS u16 counters, two per word
    __shared__ float red[6][MS_THREADS / 32];
    __shared__ float org[4];
    __shared__ unsigned warp_tot[MS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xyz = xyz_all + (size_t)blockIdx.x * n * 3;
    float4* sorted = sorted_all + (size_t)blockIdx.x * n;

    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = tid; k < n; k += MS_THREADS)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = __ldg(xyz + 3 * k + a); lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    for (int c = tid; c < MORTON_CELLS / 2; c += MS_THREADS) hist[c] = 0u;
    __syncthreads();
    if (tid == 0) {
        float L[3], H[3];
        for (int a = 0; a < 3; ++a) {
            L[a] = red[a][0]; H[a] = red[3 + a][0];
            for (int w = 1; w < MS_THREADS / 32; ++w) { L[a] = fminf(L[a], red[a][w]); H[a] = fmaxf(H[a], red[3 + a][w]); }
        }
        const float ext = fmaxf(fmaxf(H[0] - L[0], H[1] - L[1]), H[2] - L[2]);
        const bool sane = isfinite(ext) && ext > 0.f;
        org[0] = sane ? L[0] : 0.f; org[1] = sane ? L[1] : 0.f; org[2] = sane ? L[2] : 0.f;
        org[3] = sane ? (float)(1 << MORTON_BITS) / (ext * 1.0001f) : 0.f;        // degenerate cloud: everything in cell 0
    }
    __syncthreads();
    const float ox = org[0], oy = org[1], oz = org[2], inv_h = org[3];
    for (int k = tid; k < n; k += MS_THREADS) {
        const unsigned code = morton_code(__ldg(xyz + 3 * k), __ldg(xyz + 3 * k + 1), __ldg(xyz + 3 * k + 2), ox, oy, oz, inv_h);
        atomicAdd(&hist[code >> 1], 1u << (16 * (code & 1u)));          // n <= 65535: a half never carries into its neighbour
    }
    __syncthreads();
    // exclusive scan of the 32768 counters: 64 per thread (32 words)
    constexpr int WPT = MORTON_CELLS / 2 / MS_THREADS;
    unsigned sum = 0;
#pragma unroll 8
    for (int i = 0; i < WPT; ++i) { const unsigned w = hist[tid * WPT + i]; sum += (w & 0xFFFFu) + (w >> 16); }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned run = incl - sum;
    for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll 8
    for (int i = 0; i < WPT; ++i) {
        const unsigned w = hist[tid * WPT + i];
        const unsigned a = run, b = run + (w & 0xFFFFu);
        hist[tid * WPT + i] = a | (b << 16);
        run = b + (w >> 16);
    }
    __syncthreads();
    for (int k = tid; k < n; k += MS_THREADS) {
        const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
        const unsigned code = morton_code(x, y, z, ox, oy, oz, inv_h);
        const unsigned sh = 16 * (code & 1u);
        const unsigned old = atomicAdd(&hist[code >> 1], 1u << sh);
        sorted[(old >> sh) & 0xFFFFu] = make_float4(x, y, z, __int_as_float(k));
    }
}

// float -> int whose signed order is the float order (for redux.sync min/max of coordinates)
__device__ __forceinline__ int f2ord(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7FFFFFFF); }
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7FFFFFFF)); }

constexpr int FR_CPW = 16;                              // clumps per warp

// Update of one flagged clump C (a compile-time index: temp[] lives in registers).  All 32 lanes take part, one point each.
// No divergence: the owner lane's record is updated with selects, and the sqrt of the threshold overlaps the shuffles.
#define G4D_FR_UPDATE(C)                                                                                                   \
    {                                                                                                                      \
        const int pos = base + (C) * (NW * 32);                                                                            \
        const unsigned k = ks[pos];                                                                                        \
        const float d = sqdist_ref(xs[pos] - x1, ys[pos] - y1, zs[pos] - z1);                                              \
        const float t = fminf(d, temp[C]);                                                                                 \
        temp[C] = t;                                                                                                       \
        const int tb = __float_as_int(t);    /* t >= 0, or -1 for padding: int order == float order */                    \
        const int mb = __reduce_max_sync(FULL, tb);                                                                        \
        unsigned who = __ballot_sync(FULL, tb == mb);                                                                      \
        const unsigned key = (__brev(k & bs_mask) & himask) | (k >> lg_bs);                                                \
        const float sq = sqrtf(fmaxf(__int_as_float(mb), 0.f));                                                            \
        if (__popc(who) > 1) {               /* equal maxima inside the clump (duplicate points): smallest key wins */     \
            const unsigned kmin = __reduce_min_sync(FULL, tb == mb ? key : 0xFFFFFFFFu);                                   \
            who = __ballot_sync(FULL, tb == mb && key == kmin);                                                            \
        }                                                                                                                  \
        const int src = __ffs(who) - 1;                                                                                    \
        const unsigned kk = __shfl_sync(FULL, key, src);                                                                   \
        const bool own = lane == (C);                                                                                      \
        const float s_ = crad + sq;                                                                                        \
        cmax = own ? mb : cmax; ckey = own ? kk : ckey; csrc = own ? src : csrc;                                           \
        thr = own ? s_ * s_ * 1.0002f : thr;                                                                               \
    }

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW <= 16 ? 2 : 1)
fps_rows_kernel(int n, int m, int lg_bs, const float4* __restrict__ sorted_all, int* __restrict__ idx_all,
                float* __restrict__ new_xyz_all) {
    constexpr int CAP = NW * FR_CPW * 32;              // points this CTA can hold
    extern __shared__ __align__(16) float soa[];       // xs[CAP], ys[CAP], zs[CAP], ks[CAP] (u16)
    float* xs = soa;
    float* ys = xs + CAP;
    float* zs = ys + CAP;
    unsigned short* ks = reinterpret_cast<unsigned short*>(zs + CAP);
    __shared__ int slot_v[2][NW];
    __shared__ unsigned slot_k[2][NW];
    __shared__ float slot_p[2][NW][3];
    __shared__ float first_xyz[3];
    __shared__ unsigned latch_k[16];                   // results of the last <= 16 steps, flushed together by warp 0 (16, not 32:
    __shared__ float latch_p[16][3];                   //  two CTAs of 112 KB + static data must fit the SM's 228 KB)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    const size_t cloud = blockIdx.x;
    const float4* sorted = sorted_all + cloud * (size_t)n;
    int* idx_out = idx_all + cloud * (size_t)m;
    float* new_xyz = new_xyz_all ? new_xyz_all + cloud * (size_t)m * 3 : nullptr;
    const unsigned himask = lg_bs ? ~((1u << (32 - lg_bs)) - 1u) : 0u;
    const unsigned bs_mask = (1u << lg_bs) - 1u;
    // Clump g (32 consecutive points of the Morton order) belongs to warp g % NW: the clumps a new sample touches are
    // neighbours in that order, so they land in DIFFERENT warps and are updated in parallel (a warp handles its flagged clumps
    // one after the other, ~200 cycles of dependent latency each).  Clump c of this warp = g = c * NW + warp.
    const int base = warp * 32 + lane;                 // position of this lane's point of clump 0; clump c: + c * NW * 32

    // ---- load: one point per (clump, lane); clump records (bounding sphere) into lane c ----
    float temp[FR_CPW];
    float ccx = 0.f, ccy = 0.f, ccz = 0.f, crad = 0.f, thr = -1.f;
    int cmax = __float_as_int(-1.f);
    unsigned ckey = 0xFFFFFFFFu;
    int csrc = 0;
#pragma unroll
    for (int c = 0; c < FR_CPW; ++c) {
        const int pos = base + c * (NW * 32);
        const bool valid = pos < n;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            p = __ldg(sorted + pos);
            if (__float_as_int(p.w) == 0) { first_xyz[0] = p.x; first_xyz[1] = p.y; first_xyz[2] = p.z; }
        }
        xs[pos] = p.x; ys[pos] = p.y; zs[pos] = p.z;
        ks[pos] = (unsigned short)__float_as_int(p.w);
        temp[c] = valid ? 1e10f : -1.f;                 // the reference's pre-fill (pointnet2_utils.py:26); -1: never a maximum
        const int lox = __reduce_min_sync(FULL, valid ? f2ord(p.x) : INT_MAX), hix = __reduce_max_sync(FULL, valid ? f2ord(p.x) : INT_MIN);
        const int loy = __reduce_min_sync(FULL, valid ? f2ord(p.y) : INT_MAX), hiy = __reduce_max_sync(FULL, valid ? f2ord(p.y) : INT_MIN);
        const int loz = __reduce_min_sync(FULL, valid ? f2ord(p.z) : INT_MAX), hiz = __reduce_max_sync(FULL, valid ? f2ord(p.z) : INT_MIN);
        const bool any = __any_sync(FULL, valid);
        const float mx = 0.5f * (ord2f(lox) + ord2f(hix)), my = 0.5f * (ord2f(loy) + ord2f(hiy)), mz = 0.5f * (ord2f(loz) + ord2f(hiz));
        const float ex = p.x - mx, ey = p.y - my, ez = p.z - mz;
        const float e2 = valid ? ex * ex + ey * ey + ez * ez : 0.f;
        const float r2 = __int_as_float(__reduce_max_sync(FULL, __float_as_int(e2)));           // e2 >= 0: bits order like values
        if (lane == c && any) {
            ccx = mx; ccy = my; ccz = mz;
            crad = sqrtf(r2) * 1.0001f;
            thr = INFINITY;                              // first step: every non-empty clump is updated
            cmax = __float_as_int(1e10f);
        }
    }
    __syncthreads();
    float x1 = first_xyz[0], y1 = first_xyz[1], z1 = first_xyz[2];
    if (tid == 0) {
        idx_out[0] = 0;
        if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
    }

    // the warp's candidate (cached between its updates) and the output latch of warp 0
    int wv = __float_as_int(-1.f);
    unsigned wkey = 0xFFFFFFFFu;
    float wx = 0.f, wy = 0.f, wz = 0.f;
    int dirty = 2;                                       // slot writes still owed (one per parity buffer)

    for (int j = 1; j < m; ++j) {
        const float dcx = ccx - x1, dcy = ccy - y1, dcz = ccz - z1;
        const float d2c = dcx * dcx + dcy * dcy + dcz * dcz;
        unsigned mask = __ballot_sync(FULL, lane < FR_CPW && d2c < thr);
        if (mask) {
            do {                                         // warp-uniform walk over the flagged clumps
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                switch (c) {
                    case 0: G4D_FR_UPDATE(0) break;   case 1: G4D_FR_UPDATE(1) break;   case 2: G4D_FR_UPDATE(2) break;
                    case 3: G4D_FR_UPDATE(3) break;   case 4: G4D_FR_UPDATE(4) break;   case 5: G4D_FR_UPDATE(5) break;
                    case 6: G4D_FR_UPDATE(6) break;   case 7: G4D_FR_UPDATE(7) break;   case 8: G4D_FR_UPDATE(8) break;
                    case 9: G4D_FR_UPDATE(9) break;   case 10: G4D_FR_UPDATE(10) break; case 11: G4D_FR_UPDATE(11) break;
                    case 12: G4D_FR_UPDATE(12) break; case 13: G4D_FR_UPDATE(13) break; case 14: G4D_FR_UPDATE(14) break;
                    default: G4D_FR_UPDATE(15) break;
                }
            } while (mask);
            // the warp's best clump
            const int v = lane < FR_CPW ? cmax : INT_MIN;
            wv = __reduce_max_sync(FULL, v);
            unsigned m1 = __ballot_sync(FULL, v == wv);
            if (__popc(m1) > 1) {
                const unsigned wk = __reduce_min_sync(FULL, v == wv ? ckey : 0xFFFFFFFFu);
                m1 = __ballot_sync(FULL, v == wv && ckey == wk);
            }
            const int b = __ffs(m1) - 1;
            wkey = __shfl_sync(FULL, ckey, b);
            const int pos = (b * NW + warp) * 32 + __shfl_sync(FULL, csrc, b);
            wx = xs[pos]; wy = ys[pos]; wz = zs[pos];    // broadcast reads
            dirty = 2;
        }
        const int par = j & 1;
        if (dirty) {                                     // warp-uniform
            if (lane == 0) {
                slot_v[par][warp] = wv; slot_k[par][warp] = wkey;
                slot_p[par][warp][0] = wx; slot_p[par][warp][1] = wy; slot_p[par][warp][2] = wz;
            }
            --dirty;
        }
        __syncthreads();
        const int sv = lane < NW ? slot_v[par][lane] : INT_MIN;
        const int bv = __reduce_max_sync(FULL, sv);
        unsigned m2 = __ballot_sync(FULL, sv == bv);
        if (__popc(m2) > 1) {                            // equal maxima in several warps: smallest key wins
            const unsigned sk = lane < NW ? slot_k[par][lane] : 0xFFFFFFFFu;
            const unsigned bk2 = __reduce_min_sync(FULL, sv == bv ? sk : 0xFFFFFFFFu);
            m2 = __ballot_sync(FULL, sv == bv && sk == bk2);
        }
        const int wl = __ffs(m2) - 1;
        x1 = slot_p[par][wl][0]; y1 = slot_p[par][wl][1]; z1 = slot_p[par][wl][2];
        // Results are latched (shared memory) by warp 0 and written out 16 steps at a time: a global store in every step
        // would make each barrier wait for its acknowledgement (BAR.SYNC drains the warp's outstanding stores).
        if (warp == 0) {
            if (lane == 0) { latch_k[j & 15] = slot_k[par][wl]; latch_p[j & 15][0] = x1; latch_p[j & 15][1] = y1; latch_p[j & 15][2] = z1; }
            if ((j & 15) == 15 || j == m - 1) {
                __syncwarp();
                const int jj = (j & ~15) + lane;
                if (lane < 16 && jj >= 1 && jj <= j) {
                    const unsigned out_k = latch_k[lane];
                    idx_out[jj] = (int)(((out_k & ~himask) << lg_bs) | __brev(out_k & himask));
                    if (new_xyz) { new_xyz[3 * jj] = latch_p[lane][0]; new_xyz[3 * jj + 1] = latch_p[lane][1]; new_xyz[3 * jj + 2] = latch_p[lane][2]; }
                }
                __syncwarp();
            }
        }
    }
}
#undef G4D_FR_UPDATE

int fps_pruned_sorted(int b, int n, int m, const float4* sorted, long long stride, int* idx, float* new_xyz, cudaStream_t s);   // fps_pruned.cu

template <int NW>
static int launch_fps_rows(int b, int n, int m, int lg, const float4* sorted, int* idx, float* new_xyz, cudaStream_t s) {
    auto kern = fps_rows_kernel<NW>;
    const size_t smem = (size_t)NW * FR_CPW * 32 * (3 * sizeof(float) + sizeof(unsigned short));
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("fps_rows: cannot opt in to %zu B shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    kern<<<b, NW * 32, smem, s>>>(n, m, lg, sorted, idx, new_xyz);
    return finish_launch("g4d fps_rows kernel");
}

}  // namespace g4d

using namespace g4d;

// Workspace of g4d_fps_gather_ws: the Morton-ordered copy of the clouds, (b, n) float4.
G4D_API size_t g4d_fps_workspace_bytes(int b, int n) { return (size_t)(b < 0 ? 0 : b) * (size_t)(n < 0 ? 0 : n) * 16; }

// = g4d_fps_gather (same idx and new_xyz) through the Morton-ordered pruned kernel.  1 <= n <= 16384; workspace: device
// buffer of g4d_fps_workspace_bytes(b, n), 16-byte aligned (contents: scratch).
G4D_API int g4d_fps_gather_ws(int b, int n, int m, const float* xyz, int* idx, float* new_xyz, void* workspace, void* stream) {
    if (b < 0 || n <= 0 || m < 0) return bad_arg("fps_gather_ws: need b >= 0, n > 0, m >= 0");
    if (b == 0 || m == 0) return 0;
    if (!xyz || !idx || !workspace || ((uintptr_t)workspace & 15)) return bad_arg("fps_gather_ws: null or misaligned pointer");
    if (n > 16384) return bad_arg("fps_gather_ws: n > 16384 (use g4d_fps_gather)");
    const int bs = ref_opt_n_threads(n);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t hist_bytes = (size_t)MORTON_CELLS * 2;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(fps_morton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
        if (e != cudaSuccess) { set_error("fps_morton: shared memory opt-in: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    fps_morton_kernel<<<b, MS_THREADS, hist_bytes, s>>>(n, xyz, (float4*)workspace);
    int rc = finish_launch("g4d fps_morton kernel");
    if (rc) return rc;
    const float4* sorted = (const float4*)workspace;
    // n <= 8192: the thread-per-clump kernel (fps_pruned.cu) on the Morton order -- measured at 240 x 8192 (B200): 1.00 ms against
    // 1.26 ms for the warp-row kernel below and 1.29 ms for the same kernel on the cell-sorted order of g4d_grid_build.  The
    // warp-row kernel serves 8192 < n <= 16384 (one point per lane per clump keeps 16384 temps in registers).  G4D_FPS_WS=rows
    // forces it for every n.
    static const bool force_rows = getenv("G4D_FPS_WS") && !strcmp(getenv("G4D_FPS_WS"), "rows");
    if (!force_rows && n <= 8192) return fps_pruned_sorted(b, n, m, sorted, (long long)n, idx, new_xyz, s);
    if (n <= 2048) return launch_fps_rows<4>(b, n, m, lg, sorted, idx, new_xyz, s);
    if (n <= 4096) return launch_fps_rows<8>(b, n, m, lg, sorted, idx, new_xyz, s);
    if (n <= 8192) return launch_fps_rows<16>(b, n, m, lg, sorted, idx, new_xyz, s);
    return launch_fps_rows<32>(b, n, m, lg, sorted, idx, new_xyz, s);
}
