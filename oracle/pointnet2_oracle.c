/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the pointnet2 hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (garment4d_b200/) never does, and fails loudly when its CUDA library is
 * missing.
 *
 * The reference (hongfz16/Garment4D) has NO CPU implementation of these ops:
 * they exist only as CUDA kernels under modules/pointnet2/pointnet2/src/.
 * Each function below restates one of those kernels in plain C, following the
 * kernel line by line (file:line cited per function), including the floating
 * point evaluation order that nvcc 12.9 -O2 emits for the reference source
 * (checked in SASS of oracle/_ref/libpointnet2_ref.so):
 *
 *     d = fma(dz, dz, fma(dx, dx, rn(dy * dy)))
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (see Makefile).  -ffp-contract=off
 * matters: every fused multiply-add here is an explicit fmaf().
 *
 * Parity pin: this restatement is checked bit-for-bit against the reference's
 * own kernels (compiled unmodified into oracle/_ref/ by build_ref.sh) on the
 * GPU box by tests/test_parity_gpu.py, and against the golden vectors those
 * kernels produced (tests/golden/pointnet2_ref_*.npz) on CPU.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* cuda_utils.h:10-14 -- largest power of two <= work_size, clamped to [1,1024],
 * computed through the same double log ratio the reference uses. */
int orc_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

/* bench.py's reference arm sets the thread count explicitly (torchrun exports OMP_NUM_THREADS=1 to every rank). */
void orc_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* squared distance exactly as the reference kernels evaluate it */
static inline float sqdist(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* sampling_gpu.cu:93-209 furthest_point_sampling_kernel<block_size>.
 * Simulates the block: thread t owns points k == t (mod bs); per-thread running
 * best uses strict '>' (:136-137); the shared-memory tree (:143-203, __update
 * :86-91) keeps the LOWER slot on ties.  temp must be pre-filled by the caller
 * (the Python side fills 1e10, pointnet2_utils.py:26). */
void orc_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp, int *idxs) {
    if (m <= 0) return;
    const int bs = orc_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        const float *ds = dataset + (size_t)bi * n * 3;
        float *tp = temp + (size_t)bi * n;
        int *out = idxs + (size_t)bi * m;
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
            for (int t = 0; t < bs; ++t) { dists[t] = -1.0f; dists_i[t] = 0; }
            for (int k = 0; k < n; ++k) {
                const int t = k & (bs - 1);   /* k mod bs (bs is a power of two); ascending k within a thread, as the strided loop :124 */
                const float x2 = ds[k * 3 + 0], y2 = ds[k * 3 + 1], z2 = ds[k * 3 + 2];
                const float d = sqdist(x2 - x1, y2 - y1, z2 - z1);
                const float d2 = fminf(d, tp[k]);
                tp[k] = d2;
                if (d2 > dists[t]) { dists[t] = d2; dists_i[t] = k; }
            }
            for (int s = bs >> 1; s >= 1; s >>= 1) {
                for (int t = 0; t < s; ++t) {
                    const float v1 = dists[t], v2 = dists[t + s];
                    const int i1 = dists_i[t], i2 = dists_i[t + s];
                    dists[t] = v1 > v2 ? v1 : v2;   /* max(v1, v2) */
                    dists_i[t] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists); free(dists_i);
    }
}

static inline unsigned bitrev_bits(unsigned v, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

/* Same result as orc_furthest_point_sampling, restated as three flat passes so
 * the compiler can vectorise it (used as the timed CPU baseline; tests check
 * it against the line-by-line simulation above).  Equivalence: the block's
 * winner is, among the points whose updated temp equals the global maximum,
 * the one with the smallest (bitrev(k mod bs), k div bs): strict '>' keeps the
 * lowest k inside a thread (sampling_gpu.cu:136-137) and the tree keeps the
 * lower slot on ties at strides bs/2 ... 1 (:143-203), i.e. orders thread ids
 * by their bit-reversal. */
void orc_furthest_point_sampling_fast(int b, int n, int m, const float *dataset, float *temp, int *idxs) {
    if (m <= 0) return;
    const int bs = orc_opt_n_threads(n);
    int lg = 0; while ((1 << lg) < bs) ++lg;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        const float *ds = dataset + (size_t)bi * n * 3;
        float *tp = temp + (size_t)bi * n;
        int *out = idxs + (size_t)bi * m;
        float *xs = (float *)malloc(sizeof(float) * n * 3);
        float *ys = xs + n, *zs = ys + n;
        for (int k = 0; k < n; ++k) { xs[k] = ds[k * 3]; ys[k] = ds[k * 3 + 1]; zs[k] = ds[k * 3 + 2]; }
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = xs[old], y1 = ys[old], z1 = zs[old];
            float vmax = -1.0f;
            for (int k = 0; k < n; ++k) {
                const float dx = xs[k] - x1, dy = ys[k] - y1, dz = zs[k] - z1;
                const float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                const float d2 = d < tp[k] ? d : tp[k];
                tp[k] = d2;
                vmax = d2 > vmax ? d2 : vmax;
            }
            unsigned long long bestkey = ~0ull; int besti = 0;
            for (int k = 0; k < n; ++k) {
                if (tp[k] == vmax) {
                    const unsigned long long key = ((unsigned long long)bitrev_bits((unsigned)(k & (bs - 1)), lg) << 32) | (unsigned)(k >> lg);
                    if (key < bestkey) { bestkey = key; besti = k; }
                }
            }
            old = besti;
            out[j] = old;
        }
        free(xs);
    }
}

/* sampling_gpu.cu:8-24 gather_points_kernel_fast: out[b,c,j] = points[b,c,idx[b,j]] */
void orc_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *p = points + ((size_t)bi * c + ci) * n;
            const int *id = idx + (size_t)bi * npoints;
            float *o = out + ((size_t)bi * c + ci) * npoints;
            for (int j = 0; j < npoints; ++j) o[j] = p[id[j]];
        }
}

/* sampling_gpu.cu:46-63 gather_points_grad_kernel_fast: atomicAdd scatter.  The
 * reference's float summation order is nondeterministic; this oracle adds in
 * ascending j.  grad_points must be pre-zeroed (pointnet2_utils.py:67). */
void orc_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * npoints;
            const int *id = idx + (size_t)bi * npoints;
            float *gp = grad_points + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < npoints; ++j) gp[id[j]] += g[j];
        }
}

/* ball_query_gpu.cu:9-45 ball_query_kernel_fast.  idx must be pre-zeroed
 * (pointnet2_utils.py:218): a query with no hit keeps its all-zero row. */
void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    const float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int pt = 0; pt < m; ++pt) {
            const float *q = new_xyz + ((size_t)bi * m + pt) * 3;
            const float *src = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + pt) * nsample;
            const float nx = q[0], ny = q[1], nz = q[2];
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                const float d2 = sqdist(nx - src[k * 3 + 0], ny - src[k * 3 + 1], nz - src[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0) for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* group_points_gpu.cu:47-66 group_points_kernel_fast: out[b,c,p,s] = points[b,c,idx[b,p,s]] */
void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *p = points + ((size_t)bi * c + ci) * n;
            const int *id = idx + (size_t)bi * npoints * nsample;
            float *o = out + ((size_t)bi * c + ci) * npoints * nsample;
            for (int j = 0; j < npoints * nsample; ++j) o[j] = p[id[j]];
        }
}

/* group_points_gpu.cu:8-25 group_points_grad_kernel_fast (atomicAdd scatter; ascending order here) */
void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *id = idx + (size_t)bi * npoints * nsample;
            float *gp = grad_points + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < npoints * nsample; ++j) gp[id[j]] += g[j];
        }
}

/* interpolate_gpu.cu:9-52 three_nn_kernel_fast.  Running bests are double
 * (init 1e40), the distance itself is float; strict '<' so the lowest index
 * wins ties.  Writes SQUARED distances; the Python side takes sqrt
 * (pointnet2_utils.py:98). */
void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int pt = 0; pt < n; ++pt) {
            const float *u = unknown + ((size_t)bi * n + pt) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            const float ux = u[0], uy = u[1], uz = u[2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = sqdist(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + pt) * 3;
            int *oi = idx + ((size_t)bi * n + pt) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
        }
}

/* interpolate_gpu.cu:77-97 three_interpolate_kernel_fast.
 * out = w0*p0 + w1*p1 + w2*p2, contracted by nvcc 12.9 -O2 as fma(w2,p2, fma(w0,p0, rn(w1*p1)))
 * (SASS of oracle/_ref: FMUL on the +4 operands, FFMA on +0, FFMA on +8). */
void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *p = points + ((size_t)bi * c + ci) * m;
            const int *id = idx + (size_t)bi * n * 3;
            const float *w = weight + (size_t)bi * n * 3;
            float *o = out + ((size_t)bi * c + ci) * n;
            for (int j = 0; j < n; ++j)
                o[j] = fmaf(w[j * 3 + 2], p[id[j * 3 + 2]], fmaf(w[j * 3 + 0], p[id[j * 3 + 0]], w[j * 3 + 1] * p[id[j * 3 + 1]]));
        }
}

/* interpolate_gpu.cu:120-142 three_interpolate_grad_kernel_fast (3 atomicAdds per output; ascending order here) */
void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            const int *id = idx + (size_t)bi * n * 3;
            const float *w = weight + (size_t)bi * n * 3;
            float *gp = grad_points + ((size_t)bi * c + ci) * m;
            for (int j = 0; j < n; ++j)
                for (int t = 0; t < 3; ++t) gp[id[j * 3 + t]] += g[j] * w[j * 3 + t];
        }
}
