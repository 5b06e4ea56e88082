// Uniform-grid acceleration of the two neighbour searches of the hot path, with results IDENTICAL to the
// brute-force reference kernels (ball_query_gpu.cu:9-45, interpolate_gpu.cu:9-52):
//
//   g4d_grid_build          per cloud: bounding box -> cell size >= the search radius -> counting sort of the points
//                           by cell (one CTA per cloud, histogram and cursors in shared memory)
//   g4d_ball_query2_grid    one warp per query visits the 27 neighbouring cells (9 contiguous runs of the sorted
//                           array), tests candidates with the reference's exact distance arithmetic and sets bit k
//                           of a per-warp N-bit bitmap for every hit; the first nsample set bits ARE "the first
//                           nsample hits in ascending index order" -> extracted with popcount prefix sums, padded
//                           with the lowest set bit.  Both MSG radii share one candidate walk (two bitmaps).
//   g4d_three_nn_grid       one thread per unknown point walks shells of cells of a grid over the KNOWN points and
//                           keeps the 3 best under the total order (d, k) -- exactly what the reference's strict '<'
//                           insertion in ascending k produces -- until the shell's distance bound exceeds the third
//                           best (conservative margin, so rounding can never end the search early).
//
// Work drops from N*P (8192*1024 per cloud at SA1) distance tests to ~a few hundred per query.
#include <float.h>
#include <stdlib.h>
#include "common.cuh"
#include "grid.cuh"

namespace g4d {

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int dim) {
    const int c = __float2int_rd(__fmul_rn(__fsub_rn(v, o), inv_h));     // NaN -> 0
    return min(max(c, 0), dim - 1);
}

constexpr int GB_THREADS = 512;

__global__ void __launch_bounds__(GB_THREADS)
grid_build_kernel(int n, const float* __restrict__ xyz_all, float min_cell, float* __restrict__ grid_all) {
    __shared__ int hist[GRID_MAX_CELLS];
    __shared__ float red[6][GB_THREADS / 32];
    __shared__ GridHdr hdr;
    __shared__ int warp_tot[GB_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xyz = xyz_all + (size_t)blockIdx.x * n * 3;
    float* g = grid_all + (size_t)blockIdx.x * grid_cloud_words(n);

    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = tid; k < n; k += GB_THREADS)
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
    for (int c = tid; c < GRID_MAX_CELLS; c += GB_THREADS) hist[c] = 0;
    __syncthreads();
    if (tid == 0) {
        float L[3], H[3];
        for (int a = 0; a < 3; ++a) {
            L[a] = red[a][0]; H[a] = red[3 + a][0];
            for (int w = 1; w < GB_THREADS / 32; ++w) { L[a] = fminf(L[a], red[a][w]); H[a] = fmaxf(H[a], red[3 + a][w]); }
        }
        // cell edge strictly larger than the search radius, so |x - q| < r always lands within +-1 cell even after rounding
        const float ex = H[0] - L[0], ey = H[1] - L[1], ez = H[2] - L[2];
        // min_cell < 0: automatic, -min_cell cells along the longest axis (used when no search radius is implied)
        float h = min_cell > 0.f ? min_cell * 1.0001f : fmaxf(fmaxf(ex, ey), ez) / -min_cell;
        if (!(h > 1e-30f)) h = 1e-30f;
        int dx = 1, dy = 1, dz = 1;
        const bool sane = isfinite(ex) && isfinite(ey) && isfinite(ez) && h > 0.f && isfinite(h) && ex >= 0.f && ey >= 0.f && ez >= 0.f;
        if (sane) {
            for (int it = 0; it < 200; ++it) {
                const float fx = floorf(ex / h) + 1.f, fy = floorf(ey / h) + 1.f, fz = floorf(ez / h) + 1.f;
                if (fx * fy * fz <= (float)GRID_MAX_CELLS && fmaxf(fx, fmaxf(fy, fz)) <= 256.f) { dx = (int)fx; dy = (int)fy; dz = (int)fz; break; }
                h *= 1.25f;
                if (it == 199) { dx = dy = dz = 1; }
            }
        }
        // Rounding.  (1) Binning: c = floor((v - o) * inv_h) carries a relative error of ~3 * 2^-24, i.e. up to cells_per_axis *
        // 1.8e-7 cells; with at most 256 cells per axis (enforced above) that is < 5e-5 cells, inside the 1e-4 * h head-room of the
        // cell edge over the search radius, so a hit is never more than one cell away.  (2) The cell faces o + c*h used for culling
        // carry an absolute error of a few ulp of the largest coordinate, and a point may sit up to 5e-5 * h outside the faces of
        // the cell it was binned to: the searches subtract hdr.eps from every face gap they cull with.
        float amax = 0.f;
        for (int a = 0; a < 3; ++a) amax = fmaxf(amax, fmaxf(fabsf(L[a]), fabsf(H[a])));
        const float eps = sane ? amax * 9.5367431640625e-07f + 1e-4f * h : 0.f;      // 2^-20 * max|coord| (8 ulp) + binning slack
        hdr.ox = sane ? L[0] : 0.f; hdr.oy = sane ? L[1] : 0.f; hdr.oz = sane ? L[2] : 0.f;
        hdr.h = h; hdr.inv_h = (dx * dy * dz > 1) ? 1.0f / h : 0.f;
        hdr.dx = dx; hdr.dy = dy; hdr.dz = dz; hdr.ncells = dx * dy * dz;
        hdr.eps = eps;
        for (int i = 0; i < 6; ++i) hdr.pad[i] = 0;
        *reinterpret_cast<GridHdr*>(g) = hdr;
    }
    __syncthreads();
    const GridHdr H = hdr;
    for (int k = tid; k < n; k += GB_THREADS) {
        const int c = (cell_coord(__ldg(xyz + 3 * k + 2), H.oz, H.inv_h, H.dz) * H.dy + cell_coord(__ldg(xyz + 3 * k + 1), H.oy, H.inv_h, H.dy)) * H.dx +
                      cell_coord(__ldg(xyz + 3 * k), H.ox, H.inv_h, H.dx);
        atomicAdd(&hist[c], 1);
    }
    __syncthreads();
    // exclusive scan of hist[0..GRID_MAX_CELLS): 8 cells per thread
    constexpr int CPT = GRID_MAX_CELLS / GB_THREADS;
    int local[CPT], sum = 0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) { local[i] = hist[tid * CPT + i]; sum += local[i]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    int run = base + incl - sum;
    int* cell_start = reinterpret_cast<int*>(g) + GRID_HDR;
#pragma unroll
    for (int i = 0; i < CPT; ++i) { hist[tid * CPT + i] = run; cell_start[tid * CPT + i] = run; run += local[i]; }
    if (tid == GB_THREADS - 1) cell_start[GRID_MAX_CELLS] = run;
    __syncthreads();
    float4* sorted = reinterpret_cast<float4*>(g + (GRID_HDR + GRID_MAX_CELLS + 1 + 3) / 4 * 4);
    for (int k = tid; k < n; k += GB_THREADS) {
        const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
        const int c = (cell_coord(z, H.oz, H.inv_h, H.dz) * H.dy + cell_coord(y, H.oy, H.inv_h, H.dy)) * H.dx + cell_coord(x, H.ox, H.inv_h, H.dx);
        const int pos = atomicAdd(&hist[c], 1);
        sorted[pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

// ---------------------------------------------------------------------------------------------------------
constexpr int BQG_WARPS = 8;
constexpr int BQG_QPW = 4;

template <int NS>
__global__ void __launch_bounds__(BQG_WARPS * 32)
ball_query_grid_kernel(int n, int m, int nwords, const float* __restrict__ new_xyz_all, const float* __restrict__ grid_all,
                       const float* __restrict__ qgrid_all, float r2_0, int K0, int* __restrict__ idx0_all, float r2_1, int K1,
                       int* __restrict__ idx1_all) {
    extern __shared__ unsigned bitmap_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cloud = blockIdx.y;
    const float* g = grid_all + cloud * grid_cloud_words(n);
    const GridHdr H = *grid_hdr(g);
    const int* cell_start = grid_cell_start(g);
    const float4* sorted = grid_sorted(g);
    // Optional processing order: the cell-sorted record of a grid built over the QUERIES, so that the queries of a block are spatial
    // neighbours and their candidate runs hit L1 instead of L2 (FPS hands the centroids over in a far-apart order).
    const float4* qsorted = qgrid_all ? grid_sorted(qgrid_all + cloud * grid_cloud_words(m)) : nullptr;
    unsigned* bm[2];
    bm[0] = bitmap_all + (size_t)warp * NS * nwords;
    bm[1] = bm[0] + nwords;
    for (int s = 0; s < NS; ++s)
        for (int w = lane; w < nwords; w += 32) bm[s][w] = 0u;
    __syncwarp();
    const float r2[2] = {r2_0, r2_1};
    const float r2max = NS > 1 ? fmaxf(r2_0, r2_1) : r2_0;
    const int KK[2] = {K0, K1};
    int* outs[2] = {idx0_all + cloud * (size_t)m * K0, NS > 1 ? idx1_all + cloud * (size_t)m * K1 : nullptr};
    const int wpl = (nwords + 31) / 32;      // bitmap words per lane (consecutive)

    for (int qi = 0; qi < BQG_QPW; ++qi) {
        const int qslot = (blockIdx.x * BQG_WARPS + warp) * BQG_QPW + qi;
        if (qslot >= m) break;              // warp-uniform
        int q = qslot;
        float qx, qy, qz;
        if (qsorted) {
            const float4 qv = __ldg(qsorted + qslot);                  // (x, y, z, bits(query index)): the same floats as new_xyz[q]
            qx = qv.x; qy = qv.y; qz = qv.z; q = __float_as_int(qv.w);
        } else {
            const float* qp = new_xyz_all + (cloud * m + q) * 3;
            qx = __ldg(qp); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
        }
        const int cx = cell_coord(qx, H.ox, H.inv_h, H.dx), cy = cell_coord(qy, H.oy, H.inv_h, H.dy), cz = cell_coord(qz, H.oz, H.inv_h, H.dz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, H.dx - 1);
        // The 27 neighbour cells are 9 contiguous runs of the sorted array (x is the fastest cell index).  Lanes 0..8
        // fetch one run's bounds each (one memory latency instead of nine), the runs are concatenated with a prefix
        // sum, and all 32 lanes then walk the concatenation.
        int beg = 0, len = 0;
        if (lane < 9) {
            const int oz = lane / 3 - 1, oy = lane % 3 - 1;
            const int zz = cz + oz, yy = cy + oy;
            if (zz >= 0 && zz < H.dz && yy >= 0 && yy < H.dy) {
                // Conservative culling against the larger radius: gap from the query to the neighbouring row of cells
                // (0 for the query's own row; computed from the same cell faces the point binning used, 0.1 % slack).
                const float fy0 = H.oy + (float)cy * H.h, fz0 = H.oz + (float)cz * H.h, fx0 = H.ox + (float)cx * H.h;
                const float gy = oy < 0 ? qy - fy0 : (oy > 0 ? (fy0 + H.h) - qy : 0.f);
                const float gz = oz < 0 ? qz - fz0 : (oz > 0 ? (fz0 + H.h) - qz : 0.f);
                const float gyp = fmaxf(gy - H.eps, 0.f), gzp = fmaxf(gz - H.eps, 0.f);
                const float rem = r2max * 1.001f - gyp * gyp - gzp * gzp;
                if (rem >= 0.f && H.inv_h > 0.f) {
                    const float gxl = fmaxf(qx - fx0 - H.eps, 0.f), gxr = fmaxf((fx0 + H.h) - qx - H.eps, 0.f);
                    const int xa = (gxl * gxl > rem) ? cx : x0;          // left neighbour cell cannot hold a hit
                    const int xb = (gxr * gxr > rem) ? cx : x1;
                    const int row = (zz * H.dy + yy) * H.dx;
                    beg = __ldg(cell_start + row + xa);
                    len = __ldg(cell_start + row + xb + 1) - beg;
                } else if (H.inv_h == 0.f) {                             // degenerate single-cell grid: everything is a candidate
                    const int row = (zz * H.dy + yy) * H.dx;
                    beg = __ldg(cell_start + row + x0);
                    len = __ldg(cell_start + row + x1 + 1) - beg;
                }
            }
        }
        for (int r = 0; r < 9; ++r) {
            const int rb = __shfl_sync(0xFFFFFFFFu, beg, r), re = rb + __shfl_sync(0xFFFFFFFFu, len, r);
            for (int j = rb + lane; j < re; j += 32) {
                const float4 p = __ldg(sorted + j);
                const float d2 = sqdist_ref(qx - p.x, qy - p.y, qz - p.z);
                if (d2 < r2max) {                                   // warp-divergent, but most candidates fail here
                    const unsigned k = (unsigned)__float_as_int(p.w);
                    const unsigned bit = 1u << (k & 31);
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        if (d2 < r2[s]) atomicOr(&bm[s][k >> 5], bit);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int K = KK[s];
            int* row = outs[s] + (size_t)q * K;
            // each lane owns wpl consecutive words (held in registers when wpl == 8, i.e. n <= 8192, the hot case)
            unsigned wreg[8];
            int cnt = 0;
            if (wpl == 8 && nwords == 256) {
                const uint4 a4 = reinterpret_cast<const uint4*>(bm[s])[lane * 2], b4 = reinterpret_cast<const uint4*>(bm[s])[lane * 2 + 1];
                wreg[0] = a4.x; wreg[1] = a4.y; wreg[2] = a4.z; wreg[3] = a4.w; wreg[4] = b4.x; wreg[5] = b4.y; wreg[6] = b4.z; wreg[7] = b4.w;
                reinterpret_cast<uint4*>(bm[s])[lane * 2] = make_uint4(0, 0, 0, 0);          // ready for the next query
                reinterpret_cast<uint4*>(bm[s])[lane * 2 + 1] = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int w = 0; w < 8; ++w) cnt += __popc(wreg[w]);
            } else {
                for (int w = 0; w < wpl; ++w) { const int wi = lane * wpl + w; if (wi < nwords) cnt += __popc(bm[s][wi]); }
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total > 0) {       // rows without any hit are left untouched (ball_query_gpu.cu: idx stays as the caller zero-filled it)
                int rank = incl - cnt;
                int first_local = 0x7FFFFFFF;
                if (wpl == 8 && nwords == 256) {
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        unsigned bits = wreg[w];
                        const int wi = lane * 8 + w;
                        if (bits && first_local == 0x7FFFFFFF) first_local = wi * 32 + __ffs(bits) - 1;
                        while (bits && rank < K) {
                            row[rank++] = wi * 32 + __ffs(bits) - 1;
                            bits &= bits - 1;
                        }
                    }
                } else {
                    for (int w = 0; w < wpl; ++w) {
                        const int wi = lane * wpl + w;
                        if (wi >= nwords) break;
                        unsigned bits = bm[s][wi];
                        if (bits && first_local == 0x7FFFFFFF) first_local = wi * 32 + __ffs(bits) - 1;
                        while (bits && rank < K) {
                            row[rank++] = wi * 32 + __ffs(bits) - 1;
                            bits &= bits - 1;
                        }
                    }
                }
                const int first = __reduce_min_sync(0xFFFFFFFFu, first_local);
                for (int p = total + lane; p < K; p += 32) row[p] = first;
            }
            if (!(wpl == 8 && nwords == 256))
                for (int w = 0; w < wpl; ++w) { const int wi = lane * wpl + w; if (wi < nwords) bm[s][wi] = 0u; }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// three nearest neighbours through a grid over the KNOWN points.  order (optional): processing order of the
// unknown points (the 'k' column of a grid built over them) so that the threads of a warp are spatial neighbours.

// (d, k) packed as one 64-bit key: d >= 0, so its float bits order like the value and "smaller key" == the reference's
// strict '<' insertion while scanning k upwards (ties: lower index first).  NaN distances sort above +inf: never kept.
__device__ __forceinline__ unsigned long long nn_key(float d, int k) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)k;
}

__global__ void __launch_bounds__(256)
three_nn_grid_kernel(int n, int m, const float* __restrict__ unknown_all, const float* __restrict__ kgrid_all,
                     const float* __restrict__ ugrid_all, float* __restrict__ dist2_all, int* __restrict__ idx_all) {
    const size_t cloud = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int pt = t;
    if (ugrid_all) pt = __float_as_int(__ldg(&grid_sorted(ugrid_all + cloud * grid_cloud_words(n))[t].w));
    const float* g = kgrid_all + cloud * grid_cloud_words(m);
    const GridHdr H = *grid_hdr(g);
    const int* cell_start = grid_cell_start(g);
    const float4* sorted = grid_sorted(g);
    const float* u = unknown_all + (cloud * n + pt) * 3;
    const float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);
    const int cx = cell_coord(ux, H.ox, H.inv_h, H.dx), cy = cell_coord(uy, H.oy, H.inv_h, H.dy), cz = cell_coord(uz, H.oz, H.inv_h, H.dz);
    // fewer than 3 known points: the reference leaves distance 1e40 -> (float) inf and index 0 in the unused places
    const unsigned long long EMPTY = 0x7F80000000000000ull;
    unsigned long long k1 = EMPTY, k2 = EMPTY, k3 = EMPTY;
    float d3 = INFINITY;                                             // distance part of k3
    // own-cell faces, for the per-row lower bounds
    const float fx0 = H.ox + (float)cx * H.h, fy0 = H.oy + (float)cy * H.h, fz0 = H.oz + (float)cz * H.h;
    const int maxring = max(max(max(cx, H.dx - 1 - cx), max(cy, H.dy - 1 - cy)), max(cz, H.dz - 1 - cz));
    for (int R = 0; R <= maxring; ++R) {
        for (int zz = max(cz - R, 0); zz <= min(cz + R, H.dz - 1); ++zz) {
            const bool zshell = (zz == cz - R) || (zz == cz + R);
            // distance from the query to the slab of cells zz (0 inside the query's own slab; conservative when clamped)
            const float gz = zz < cz ? fmaxf(uz - (fz0 - (float)(cz - zz - 1) * H.h) - H.eps, 0.f) : (zz > cz ? fmaxf((fz0 + (float)(zz - cz) * H.h) - uz - H.eps, 0.f) : 0.f);
            for (int yy = max(cy - R, 0); yy <= min(cy + R, H.dy - 1); ++yy) {
                const float gy = yy < cy ? fmaxf(uy - (fy0 - (float)(cy - yy - 1) * H.h) - H.eps, 0.f) : (yy > cy ? fmaxf((fy0 + (float)(yy - cy) * H.h) - uy - H.eps, 0.f) : 0.f);
                // every point of this row of cells is at least sqrt(gy^2 + gz^2) away: skip it when that cannot beat the
                // current third best (0.998: slack for the rounding of cell faces / binning)
                const float b3 = __uint_as_float((unsigned)(k3 >> 32));
                if ((gy * gy + gz * gz) * 0.998f > b3) continue;
                const bool full = zshell || (yy == cy - R) || (yy == cy + R);
                const int row = (zz * H.dy + yy) * H.dx;
                // full row of the shell: one contiguous run; otherwise only the two end cells x = cx-R and x = cx+R
                const int nseg = full ? 1 : 2;
                for (int sgm = 0; sgm < nseg; ++sgm) {
                    int xa, xb;
                    if (full) { xa = max(cx - R, 0); xb = min(cx + R, H.dx - 1); }
                    else { xa = xb = (sgm == 0 ? cx - R : cx + R); if (xa < 0 || xa >= H.dx) continue; }
                    const int beg = __ldg(cell_start + row + xa), end = __ldg(cell_start + row + xb + 1);
                    for (int j = beg; j < end; ++j) {
                        const float4 p = __ldg(sorted + j);
                        const float d = sqdist_ref(ux - p.x, uy - p.y, uz - p.z);
                        if (d <= d3) {                               // one float compare rejects almost every candidate; NaN fails too
                            const unsigned long long key = nn_key(d, __float_as_int(p.w));
                            if (key < k3) {
                                k3 = key;
                                if (k3 < k2) { const unsigned long long tmp = k2; k2 = k3; k3 = tmp; }
                                if (k2 < k1) { const unsigned long long tmp = k1; k1 = k2; k2 = tmp; }
                                d3 = __uint_as_float((unsigned)(k3 >> 32));
                            }
                        }
                    }
                }
            }
        }
        // Every unvisited known point lies outside the visited block of cells along at least one axis, on a side where
        // the grid continues; its distance is at least the gap from the query to that face.  0.998 keeps the stop
        // test conservative under rounding of the cell assignment and of these face coordinates.
        float bound = INFINITY;
        if (cx - R > 0) bound = fminf(bound, ux - (fx0 - (float)R * H.h));
        if (cx + R < H.dx - 1) bound = fminf(bound, (fx0 + (float)(R + 1) * H.h) - ux);
        if (cy - R > 0) bound = fminf(bound, uy - (fy0 - (float)R * H.h));
        if (cy + R < H.dy - 1) bound = fminf(bound, (fy0 + (float)(R + 1) * H.h) - uy);
        if (cz - R > 0) bound = fminf(bound, uz - (fz0 - (float)R * H.h));
        if (cz + R < H.dz - 1) bound = fminf(bound, (fz0 + (float)(R + 1) * H.h) - uz);
        bound = fmaxf(bound - H.eps, 0.f);
        if (__uint_as_float((unsigned)(k3 >> 32)) < bound * bound * 0.998f) break;
    }
    float* od = dist2_all + (cloud * n + pt) * 3;
    int* oi = idx_all + (cloud * n + pt) * 3;
    od[0] = __uint_as_float((unsigned)(k1 >> 32)); od[1] = __uint_as_float((unsigned)(k2 >> 32)); od[2] = __uint_as_float((unsigned)(k3 >> 32));
    oi[0] = (int)(unsigned)k1; oi[1] = (int)(unsigned)k2; oi[2] = (int)(unsigned)k3;
}

}  // namespace g4d

using namespace g4d;

G4D_API size_t g4d_grid_bytes(int b, int n) { return (size_t)(b < 0 ? 0 : b) * grid_cloud_words(n < 0 ? 0 : n) * 4; }

// Builds one grid per cloud over xyz (b,n,3) with cell edge >= min_cell (grown until <= 4096 cells).  grid: device buffer of
// g4d_grid_bytes(b,n), 16-byte aligned.
G4D_API int g4d_grid_build(int b, int n, const float* xyz, float min_cell, void* grid, void* stream) {
    if (b < 0 || n < 0) return bad_arg("grid_build: negative size");
    if (b == 0) return 0;
    if (!xyz || !grid || ((uintptr_t)grid & 15)) return bad_arg("grid_build: null or misaligned pointer");
    if (!(min_cell > 0.f) && !(min_cell <= -1.f)) return bad_arg("grid_build: min_cell must be positive, or <= -1 for 'that many cells along the longest axis'");
    grid_build_kernel<<<b, GB_THREADS, 0, (cudaStream_t)stream>>>(n, xyz, min_cell, (float*)grid);
    return finish_launch("g4d grid_build");
}

// Same results as g4d_ball_query2 / g4d_ball_query (idx1 = NULL: one scale).  grid: g4d_grid_build over xyz with
// min_cell >= max(radius0, radius1).  n <= 65536.  query_grid (optional): g4d_grid_build over new_xyz (any cell size), used only
// as a spatially coherent processing order of the queries.
static int ball_query2_grid_impl(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1, int* idx1,
                                 const float* new_xyz, const void* grid, const void* query_grid, void* stream) {
    if (b < 0 || n < 0 || m < 0 || nsample0 <= 0 || (idx1 && nsample1 <= 0)) return bad_arg("ball_query2_grid: bad size");
    if (b == 0 || m == 0 || n == 0) return 0;
    if (!new_xyz || !grid || !idx0) return bad_arg("ball_query2_grid: null pointer");
    if (n > 65536) return bad_arg("ball_query2_grid: n > 65536 (use g4d_ball_query2)");
    const int nwords = (n + 31) / 32;
    const int ns = idx1 ? 2 : 1;
    const size_t smem = (size_t)BQG_WARPS * ns * nwords * 4;
    dim3 gridDim((m + BQG_WARPS * BQG_QPW - 1) / (BQG_WARPS * BQG_QPW), b);
    cudaStream_t s = (cudaStream_t)stream;
    if (ns == 2) {
        if (smem > 32 * 1024) cudaFuncSetAttribute(ball_query_grid_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ball_query_grid_kernel<2><<<gridDim, BQG_WARPS * 32, smem, s>>>(n, m, nwords, new_xyz, (const float*)grid, (const float*)query_grid, radius0 * radius0, nsample0,
                                                                        idx0, radius1 * radius1, nsample1, idx1);
    } else {
        if (smem > 32 * 1024) cudaFuncSetAttribute(ball_query_grid_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ball_query_grid_kernel<1><<<gridDim, BQG_WARPS * 32, smem, s>>>(n, m, nwords, new_xyz, (const float*)grid, (const float*)query_grid, radius0 * radius0, nsample0,
                                                                        idx0, 0.f, 1, nullptr);
    }
    return finish_launch("g4d ball_query2_grid");
}

G4D_API int g4d_ball_query2_grid(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1, int* idx1,
                                 const float* new_xyz, const void* grid, void* stream) {
    return ball_query2_grid_impl(b, n, m, radius0, nsample0, idx0, radius1, nsample1, idx1, new_xyz, grid, nullptr, stream);
}

G4D_API int g4d_ball_query2_grid_ordered(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1,
                                         int* idx1, const float* new_xyz, const void* grid, const void* query_grid, void* stream) {
    if (query_grid && ((uintptr_t)query_grid & 15)) return bad_arg("ball_query2_grid_ordered: misaligned query grid");
    return ball_query2_grid_impl(b, n, m, radius0, nsample0, idx0, radius1, nsample1, idx1, new_xyz, grid, query_grid, stream);
}

// Same results as g4d_three_nn.  known_grid: g4d_grid_build over the KNOWN points (b,m,3) (any positive min_cell; a good
// choice is the typical neighbour spacing).  unknown_grid (optional): a grid over the unknown points, used only as a
// spatially coherent processing order.
G4D_API int g4d_three_nn_grid(int b, int n, int m, const float* unknown, const void* known_grid, const void* unknown_grid,
                              float* dist2, int* idx, void* stream) {
    if (b < 0 || n < 0 || m < 0) return bad_arg("three_nn_grid: negative size");
    if (b == 0 || n == 0) return 0;
    if (!unknown || !known_grid || !dist2 || !idx) return bad_arg("three_nn_grid: null pointer");
    dim3 gridDim((n + 255) / 256, b);
    three_nn_grid_kernel<<<gridDim, 256, 0, (cudaStream_t)stream>>>(n, m, unknown, (const float*)known_grid, (const float*)unknown_grid, dist2, idx);
    return finish_launch("g4d three_nn_grid");
}
