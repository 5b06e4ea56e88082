// Ball query, grouping, gather and the fused QueryAndGroup for sm_100a.
//
// Replaces (reference, modules/pointnet2/pointnet2/src/):
//   ball_query_kernel_fast          ball_query_gpu.cu:9-45     (one THREAD per query, serial scan)
//   group_points_kernel_fast        group_points_gpu.cu:47-66
//   group_points_grad_kernel_fast   group_points_gpu.cu:8-25
//   gather_points_kernel_fast       sampling_gpu.cu:8-24
//   gather_points_grad_kernel_fast  sampling_gpu.cu:46-63
// and, fused, QueryAndGroup.forward (pointnet2_utils.py:243-265): ball_query -> transpose copy ->
// group(xyz) -> in-place subtract -> group(features) -> cat, i.e. 6 kernels and >= 4 passes over the
// grouped tensor in the reference, one kernel and one write here.
//
// Ball-query design: one WARP per QPW queries.  The source cloud streams through shared memory in
// coalesced tiles shared by all warps of the CTA; each lane tests one point per step against the
// warp's queries, hits are compacted in index order with ballot + prefix popcount, so the output is
// exactly "the first nsample hits in ascending k, padded with the first hit, untouched if none"
// (ball_query_gpu.cu:28-44).  Up to two radii (the MSG scales) are answered from the same scan.
#include "common.cuh"

namespace g4d {

constexpr int BQ_WARPS = 8;           // warps per CTA
constexpr int BQ_QPW = 4;             // queries per warp
constexpr int BQ_TILE = 2048;         // source points per shared-memory tile (24 KB)
constexpr int BQ_QPB = BQ_WARPS * BQ_QPW;

struct BallScale {
    float radius2;
    int nsample;
    int* idx;      // (b, m, nsample), rows with no hit are left untouched
};

// Scans the cloud for the CTA's queries.  NS = number of scales (1 or 2).
// cnt_out (optional, smem) receives min(cnt, nsample) per (scale, query) for the fused group stage.
template <int NS>
__device__ __forceinline__ void ball_scan(int n, int m, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                          const BallScale* sc, int q0, float* tile /*[BQ_TILE*3]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW];
    int cnt[NS][BQ_QPW], first[NS][BQ_QPW];
    bool live[BQ_QPW];
#pragma unroll
    for (int q = 0; q < BQ_QPW; ++q) {
        const int qi = q0 + warp * BQ_QPW + q;
        live[q] = qi < m;
        const int qs = live[q] ? qi : 0;
        qx[q] = __ldg(new_xyz + 3 * qs); qy[q] = __ldg(new_xyz + 3 * qs + 1); qz[q] = __ldg(new_xyz + 3 * qs + 2);
#pragma unroll
        for (int s = 0; s < NS; ++s) { cnt[s][q] = 0; first[s][q] = 0; }
    }
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int base = 0; base < n; base += BQ_TILE) {
        const int tn = min(BQ_TILE, n - base);
        __syncthreads();   // previous tile fully consumed
        for (int e = tid; e < tn * 3; e += BQ_WARPS * 32) tile[e] = __ldg(xyz + (size_t)base * 3 + e);
        __syncthreads();
        // warp-uniform: anything left to find?
        bool need = false;
#pragma unroll
        for (int q = 0; q < BQ_QPW; ++q)
#pragma unroll
            for (int s = 0; s < NS; ++s) need |= live[q] && cnt[s][q] < sc[s].nsample;
        const int all_done = __syncthreads_and(!need);
        if (all_done) break;
        if (!need) continue;
        for (int off = 0; off < tn; off += 32) {
            const int kl = off + lane;
            const bool inb = kl < tn;
            const int ks = inb ? kl : tn - 1;
            const float x = tile[3 * ks], y = tile[3 * ks + 1], z = tile[3 * ks + 2];   // stride-3 words: conflict-free
            const int k = base + kl;
            bool more = false;
#pragma unroll
            for (int q = 0; q < BQ_QPW; ++q) {
                const float d2 = sqdist_ref(qx[q] - x, qy[q] - y, qz[q] - z);
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (live[q] && cnt[s][q] < sc[s].nsample) {          // warp-uniform
                        const unsigned hits = __ballot_sync(0xFFFFFFFFu, inb && d2 < sc[s].radius2);
                        if (hits) {
                            const int K = sc[s].nsample;
                            if (cnt[s][q] == 0) first[s][q] = base + off + __ffs(hits) - 1;
                            const int pos = cnt[s][q] + __popc(hits & lt_mask);
                            if (((hits >> lane) & 1u) && pos < K)
                                sc[s].idx[(size_t)(q0 + warp * BQ_QPW + q) * K + pos] = k;
                            cnt[s][q] += __popc(hits);
                        }
                        more |= cnt[s][q] < sc[s].nsample;
                    }
                }
            }
            if (!more) break;
        }
    }
    // pad with the first hit (ball_query_gpu.cu:34-38); rows with no hit stay as the caller left them
#pragma unroll
    for (int q = 0; q < BQ_QPW; ++q)
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int K = sc[s].nsample;
            if (live[q] && cnt[s][q] > 0 && cnt[s][q] < K) {
                int* row = sc[s].idx + (size_t)(q0 + warp * BQ_QPW + q) * K;
                for (int p = cnt[s][q] + lane; p < K; p += 32) row[p] = first[s][q];
            }
        }
}

template <int NS>
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int n, int m, const float* __restrict__ xyz_all, const float* __restrict__ new_xyz_all,
                  BallScale s0, BallScale s1) {
    __shared__ float tile[BQ_TILE * 3];
    const size_t cloud = blockIdx.y;
    BallScale sc[2] = {s0, s1};
#pragma unroll
    for (int s = 0; s < NS; ++s) sc[s].idx += cloud * (size_t)m * sc[s].nsample;
    ball_scan<NS>(n, m, xyz_all + cloud * (size_t)n * 3, new_xyz_all + cloud * (size_t)m * 3, sc, blockIdx.x * BQ_QPB, tile);
}

// ---------------------------------------------------------------------------------------------
// gather / group (channel-major features, as the reference API)

__global__ void gather_points_kernel(int c, int n, int m, const float* __restrict__ points, const int* __restrict__ idx,
                                     float* __restrict__ out) {
    // out[b,c,j] = points[b,c,idx[b,j]]; one thread per j, looping over a slab of channels
    const size_t bi = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int src = __ldg(idx + bi * m + j);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        out[(bi * c + ci) * (size_t)m + j] = __ldg(points + (bi * c + ci) * (size_t)n + src);
}

__global__ void gather_points_grad_kernel(int c, int n, int m, const float* __restrict__ grad_out, const int* __restrict__ idx,
                                          float* __restrict__ grad_points) {
    const size_t bi = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int dst = __ldg(idx + bi * m + j);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        atomicAdd(grad_points + (bi * c + ci) * (size_t)n + dst, __ldg(grad_out + (bi * c + ci) * (size_t)m + j));
}

__global__ void group_points_kernel(int c, int n, int ps /* npoints*nsample */, const float* __restrict__ points,
                                    const int* __restrict__ idx, float* __restrict__ out) {
    const size_t bi = blockIdx.z;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ps) return;
    const int src = __ldg(idx + bi * ps + e);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        out[(bi * c + ci) * (size_t)ps + e] = __ldg(points + (bi * c + ci) * (size_t)n + src);
}

__global__ void group_points_grad_kernel(int c, int n, int ps, const float* __restrict__ grad_out, const int* __restrict__ idx,
                                         float* __restrict__ grad_points) {
    const size_t bi = blockIdx.z;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ps) return;
    const int dst = __ldg(idx + bi * ps + e);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        atomicAdd(grad_points + (bi * c + ci) * (size_t)n + dst, __ldg(grad_out + (bi * c + ci) * (size_t)ps + e));
}

// ---------------------------------------------------------------------------------------------
// Grouping stage of QueryAndGroup alone, given idx: (3+C) x P x K tensor written once with 16-byte stores; each thread owns
// 4 consecutive samples of one centroid (its 4 source indices stay in registers) and walks a slab of channels.
// Replaces grouping_operation(xyz) -> subtract -> grouping_operation(features) -> cat (pointnet2_utils.py:251-258).
constexpr int GF_CH_SLAB = 16;
__global__ void __launch_bounds__(256)
group_fused_kernel(int n, int m, int c, int K, int use_xyz, const float* __restrict__ xyz_all, const float* __restrict__ new_xyz_all,
                   const float* __restrict__ feat_all, const int* __restrict__ idx_all, float* __restrict__ out_all) {
    const size_t cloud = blockIdx.z;
    const size_t plane = (size_t)m * K;                       // elements per channel plane (multiple of 4: K % 4 == 0)
    const size_t e4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e4 * 4 >= plane) return;
    const int4 id = __ldg(reinterpret_cast<const int4*>(idx_all + cloud * plane) + e4);
    const int cxyz = use_xyz ? 3 : 0;
    float* out = out_all + cloud * (size_t)(cxyz + c) * plane;
    const int slab = blockIdx.y;
    if (use_xyz && slab == 0) {
        const float* xyz = xyz_all + cloud * (size_t)n * 3;
        const float* q = new_xyz_all + (cloud * m + (e4 * 4) / K) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float qd = __ldg(q + d);
            float4 v;
            v.x = __ldg(xyz + 3 * id.x + d) - qd; v.y = __ldg(xyz + 3 * id.y + d) - qd;
            v.z = __ldg(xyz + 3 * id.z + d) - qd; v.w = __ldg(xyz + 3 * id.w + d) - qd;
            reinterpret_cast<float4*>(out + d * plane)[e4] = v;
        }
    }
    if (c > 0) {
        const float* feat = feat_all + cloud * (size_t)c * n;
        float* of = out + cxyz * plane;
        const int c0 = slab * GF_CH_SLAB, c1 = min(c, c0 + GF_CH_SLAB);
#pragma unroll 4
        for (int ci = c0; ci < c1; ++ci) {
            const float* row = feat + (size_t)ci * n;
            float4 v;
            v.x = __ldg(row + id.x); v.y = __ldg(row + id.y); v.z = __ldg(row + id.z); v.w = __ldg(row + id.w);
            reinterpret_cast<float4*>(of + (size_t)ci * plane)[e4] = v;
        }
    }
}

// Grouping stage of QueryAndGroup from POINT-MAJOR fp32 features (b, n, c): a neighbour's channels are one contiguous row, so the
// gather moves 256-byte pieces (16 lanes x float4) instead of one 4-byte word per channel and lane, and the (b, 3+c, m, K) output
// -- channel-major, 98 % of the traffic -- is written through a shared-memory transpose as 256-byte runs per channel.
// (group_fused_kernel, reading the reference's channel-major layout directly, stalls on its scalar gathers at 41-47 % of DRAM
// bandwidth: profiles/r01_ncu_summary.md.)  One CTA = 64 consecutive (centroid, sample) positions x all channels, 64 at a time.
constexpr int GR_E = 64, GR_C = 64;
__global__ void __launch_bounds__(256)
group_rows_kernel(int n, int m, int c, int K, int use_xyz, const float* __restrict__ xyz_all, const float* __restrict__ new_xyz_all,
                  const float* __restrict__ feat_pm_all, const int* __restrict__ idx_all, float* __restrict__ out_all) {
    __shared__ int sidx[GR_E];
    __shared__ float tile[GR_C][GR_E + 1];
    const size_t cloud = blockIdx.y;
    const size_t plane = (size_t)m * K;
    const size_t e0 = (size_t)blockIdx.x * GR_E;
    const int ne = (int)((plane - e0) < (size_t)GR_E ? (plane - e0) : (size_t)GR_E);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cxyz = use_xyz ? 3 : 0;
    float* out = out_all + cloud * (size_t)(cxyz + c) * plane + e0;
    if (tid < GR_E) sidx[tid] = tid < ne ? __ldg(idx_all + cloud * plane + e0 + tid) : 0;
    __syncthreads();
    if (use_xyz && tid < 3 * GR_E) {
        const int d = tid / GR_E, e = tid - d * GR_E;           // lanes along e: coalesced plane writes
        if (e < ne) {
            const size_t q = (e0 + e) / K;
            out[(size_t)d * plane + e] = __ldg(xyz_all + (cloud * n + sidx[e]) * 3 + d) - __ldg(new_xyz_all + (cloud * m + q) * 3 + d);
        }
    }
    const float* feat = feat_pm_all + cloud * (size_t)n * c;
    const int hl = lane & 15, hs = lane >> 4;                    // 16 lanes x float4 = one 64-channel piece of a row; two rows per warp load
    for (int c0 = 0; c0 < c; c0 += GR_C) {
        const int nc = (c - c0) < GR_C ? (c - c0) : GR_C;
        // gather: warp w takes positions w, w+8, ... two at a time
#pragma unroll 2
        for (int e = 2 * warp + hs; e < GR_E; e += 16) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < ne && 4 * hl < nc) {
                const float* row = feat + (size_t)sidx[e] * c + c0 + 4 * hl;
                if (((c | c0) & 3) == 0 && 4 * hl + 3 < nc) v = __ldg(reinterpret_cast<const float4*>(row));
                else { v.x = __ldg(row); if (4 * hl + 1 < nc) v.y = __ldg(row + 1); if (4 * hl + 2 < nc) v.z = __ldg(row + 2); if (4 * hl + 3 < nc) v.w = __ldg(row + 3); }
            }
            tile[4 * hl][e] = v.x; tile[4 * hl + 1][e] = v.y; tile[4 * hl + 2][e] = v.z; tile[4 * hl + 3][e] = v.w;
        }
        __syncthreads();
        // write: one channel per warp iteration, 64 consecutive positions = 256 contiguous bytes
        for (int ch = warp; ch < nc; ch += 8) {
            float* o = out + (size_t)(cxyz + c0 + ch) * plane;
            if (lane < ne) o[lane] = tile[ch][lane];
            if (lane + 32 < ne) o[lane + 32] = tile[ch][lane + 32];
        }
        __syncthreads();
    }
}

static inline dim3 chan_grid(int work, int c, int b) {
    // y = channel slabs: enough CTAs to fill the machine without one CTA per channel re-reading idx
    int y = c < 8 ? c : 8;
    return dim3((work + 255) / 256, y < 1 ? 1 : y, b);
}

}  // namespace g4d

using namespace g4d;

// ball_query_kernel_launcher_fast (ball_query_gpu.h:12-13; argument order of the call site ball_query.cpp:23).
// idx rows of queries with no point in range are not written (caller zero-fills, pointnet2_utils.py:218).
G4D_API int g4d_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx, void* stream) {
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return bad_arg("ball_query: negative size");
    if (b == 0 || m == 0 || n == 0 || nsample == 0) return 0;
    if (!new_xyz || !xyz || !idx) return bad_arg("ball_query: null pointer");
    BallScale s0{radius * radius, nsample, idx};   // radius2 = rn(r*r) in float, ball_query_gpu.cu:23
    dim3 grid((m + BQ_QPB - 1) / BQ_QPB, b);
    ball_query_kernel<1><<<grid, BQ_WARPS * 32, 0, (cudaStream_t)stream>>>(n, m, xyz, new_xyz, s0, s0);
    return finish_launch("g4d ball_query");
}

// Two radii from one scan (the two MSG scales of a PointnetSAModuleMSG share xyz and new_xyz).
G4D_API int g4d_ball_query2(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1, int* idx1,
                            const float* new_xyz, const float* xyz, void* stream) {
    if (b < 0 || n < 0 || m < 0 || nsample0 <= 0 || nsample1 <= 0) return bad_arg("ball_query2: bad size");
    if (b == 0 || m == 0 || n == 0) return 0;
    if (!new_xyz || !xyz || !idx0 || !idx1) return bad_arg("ball_query2: null pointer");
    BallScale s0{radius0 * radius0, nsample0, idx0}, s1{radius1 * radius1, nsample1, idx1};
    dim3 grid((m + BQ_QPB - 1) / BQ_QPB, b);
    ball_query_kernel<2><<<grid, BQ_WARPS * 32, 0, (cudaStream_t)stream>>>(n, m, xyz, new_xyz, s0, s1);
    return finish_launch("g4d ball_query2");
}

G4D_API int g4d_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out, void* stream) {
    if (b <= 0 || c <= 0 || npoints <= 0) return (b < 0 || c < 0 || npoints < 0) ? bad_arg("gather_points: negative size") : 0;
    gather_points_kernel<<<chan_grid(npoints, c, b), 256, 0, (cudaStream_t)stream>>>(c, n, npoints, points, idx, out);
    return finish_launch("g4d gather_points");
}

G4D_API int g4d_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx, float* grad_points, void* stream) {
    if (b <= 0 || c <= 0 || npoints <= 0) return (b < 0 || c < 0 || npoints < 0) ? bad_arg("gather_points_grad: negative size") : 0;
    gather_points_grad_kernel<<<chan_grid(npoints, c, b), 256, 0, (cudaStream_t)stream>>>(c, n, npoints, grad_out, idx, grad_points);
    return finish_launch("g4d gather_points_grad");
}

G4D_API int g4d_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out, void* stream) {
    if (b <= 0 || c <= 0 || npoints <= 0 || nsample <= 0) return (b < 0 || c < 0 || npoints < 0 || nsample < 0) ? bad_arg("group_points: negative size") : 0;
    const long long ps = (long long)npoints * nsample;
    if (ps > INT32_MAX) return bad_arg("group_points: npoints*nsample exceeds int32");
    group_points_kernel<<<chan_grid((int)ps, c, b), 256, 0, (cudaStream_t)stream>>>(c, n, (int)ps, points, idx, out);
    return finish_launch("g4d group_points");
}

G4D_API int g4d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx, float* grad_points, void* stream) {
    if (b <= 0 || c <= 0 || npoints <= 0 || nsample <= 0) return (b < 0 || c < 0 || npoints < 0 || nsample < 0) ? bad_arg("group_points_grad: negative size") : 0;
    const long long ps = (long long)npoints * nsample;
    if (ps > INT32_MAX) return bad_arg("group_points_grad: npoints*nsample exceeds int32");
    group_points_grad_kernel<<<chan_grid((int)ps, c, b), 256, 0, (cudaStream_t)stream>>>(c, n, (int)ps, grad_out, idx, grad_points);
    return finish_launch("g4d group_points_grad");
}

// Grouping stage of QueryAndGroup.forward (pointnet2_utils.py:251-258) from a given idx (b,m,nsample), nsample % 4 == 0:
// out (b, 3+c, m, nsample) [use_xyz] or (b, c, m, nsample), one pass, 16-byte stores.
G4D_API int g4d_group_fused(int b, int n, int m, int c, int nsample, int use_xyz, const float* xyz, const float* new_xyz,
                            const float* features, const int* idx, float* out, void* stream) {
    if (b < 0 || n <= 0 || m < 0 || c < 0 || nsample <= 0) return bad_arg("group_fused: bad size");
    if (b == 0 || m == 0) return 0;
    if (nsample % 4) return bad_arg("group_fused: nsample must be a multiple of 4");
    if (!idx || !out || (use_xyz && (!xyz || !new_xyz)) || (c > 0 && !features)) return bad_arg("group_fused: null pointer");
    if (c == 0 && !use_xyz) return bad_arg("group_fused: no features and use_xyz = 0");
    if (((uintptr_t)idx & 15) || ((uintptr_t)out & 15)) return bad_arg("group_fused: idx/out must be 16-byte aligned");
    const long long plane4 = (long long)m * nsample / 4;
    int slabs = (c + GF_CH_SLAB - 1) / GF_CH_SLAB;
    if (slabs < 1) slabs = 1;
    dim3 grid((unsigned)((plane4 + 255) / 256), (unsigned)slabs, (unsigned)b);
    group_fused_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, c, nsample, use_xyz, xyz, new_xyz, features, idx, out);
    return finish_launch("g4d group_fused");
}

// = g4d_group_fused with the features given POINT-major: feat_pm (b, n, c) fp32 (a transposed copy of the reference's (b, c, n)
// tensor; the caller makes it once per feature tensor).  Same output, any nsample.
G4D_API int g4d_group_fused_pm(int b, int n, int m, int c, int nsample, int use_xyz, const float* xyz, const float* new_xyz,
                               const float* feat_pm, const int* idx, float* out, void* stream) {
    if (b < 0 || n <= 0 || m < 0 || c <= 0 || nsample <= 0) return bad_arg("group_fused_pm: bad size");
    if (b == 0 || m == 0) return 0;
    if (!idx || !out || !feat_pm || (use_xyz && (!xyz || !new_xyz))) return bad_arg("group_fused_pm: null pointer");
    if (b > 65535) return bad_arg("group_fused_pm: b > 65535");
    if ((uintptr_t)feat_pm & 15) return bad_arg("group_fused_pm: feat_pm must be 16-byte aligned");
    const long long plane = (long long)m * nsample;
    dim3 grid((unsigned)((plane + GR_E - 1) / GR_E), (unsigned)b);
    group_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, c, nsample, use_xyz, xyz, new_xyz, feat_pm, idx, out);
    return finish_launch("g4d group_fused_pm");
}
