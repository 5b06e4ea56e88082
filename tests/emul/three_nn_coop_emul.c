/* CPU emulation (test infrastructure) of the EXPERIMENTAL warp-cooperative three_nn kernel
 * (garment4d_b200/csrc/spatial_grid.cu: three_nn_coop_kernel) and of g4d_grid_build, statement by statement, one "warp" = 32
 * consecutive entries of the cell-sorted order of the unknown points.  It validates the LOGIC that cannot be checked on a
 * machine without a GPU: block-of-cells construction, scanning only the new cells after growing the block, and the stop test.
 * The result must equal the brute-force three nearest neighbours under the total order (distance bits, index).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see tests/test_emul_cpu.py). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_CELLS 4096

typedef struct { float ox, oy, oz, inv_h; int dx, dy, dz, ncells; float h; } Hdr;
typedef struct { Hdr H; int cell_start[MAX_CELLS + 1]; float* sorted; /* n x 4: x y z bits(k) */ } Grid;

static int cell_coord(float v, float o, float inv_h, int dim) {
    float f = (float)(v - o);                 /* __fsub_rn */
    f = (float)(f * inv_h);                   /* __fmul_rn */
    int c = (f != f) ? 0 : (int)floorf(f);    /* __float2int_rd, NaN -> 0 */
    if (c < 0) c = 0;
    if (c > dim - 1) c = dim - 1;
    return c;
}

static float sqdist_ref(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dx, dx, (float)(dy * dy))); }

/* g4d_grid_build for one cloud (spatial_grid.cu: grid_build_kernel) */
static void grid_build(int n, const float* xyz, float min_cell, Grid* G) {
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int k = 0; k < n; ++k)
        for (int a = 0; a < 3; ++a) { float v = xyz[3 * k + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    float h = min_cell > 0.f ? min_cell * 1.0001f : fmaxf(fmaxf(ex, ey), ez) / -min_cell;
    if (!(h > 1e-30f)) h = 1e-30f;
    int dx = 1, dy = 1, dz = 1;
    int sane = isfinite(ex) && isfinite(ey) && isfinite(ez) && h > 0.f && isfinite(h) && ex >= 0.f && ey >= 0.f && ez >= 0.f;
    if (sane)
        for (int it = 0; it < 200; ++it) {
            float fx = floorf(ex / h) + 1.f, fy = floorf(ey / h) + 1.f, fz = floorf(ez / h) + 1.f;
            if (fx * fy * fz <= (float)MAX_CELLS) { dx = (int)fx; dy = (int)fy; dz = (int)fz; break; }
            h *= 1.25f;
            if (it == 199) { dx = dy = dz = 1; }
        }
    Hdr* H = &G->H;
    H->ox = sane ? lo[0] : 0.f; H->oy = sane ? lo[1] : 0.f; H->oz = sane ? lo[2] : 0.f;
    H->h = h; H->inv_h = (dx * dy * dz > 1) ? 1.0f / h : 0.f;
    H->dx = dx; H->dy = dy; H->dz = dz; H->ncells = dx * dy * dz;
    int* hist = (int*)calloc(MAX_CELLS + 1, sizeof(int));
    int* cell = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) {
        int c = (cell_coord(xyz[3 * k + 2], H->oz, H->inv_h, dz) * dy + cell_coord(xyz[3 * k + 1], H->oy, H->inv_h, dy)) * dx +
                cell_coord(xyz[3 * k], H->ox, H->inv_h, dx);
        cell[k] = c; hist[c]++;
    }
    int run = 0;
    for (int c = 0; c < MAX_CELLS; ++c) { G->cell_start[c] = run; run += hist[c]; hist[c] = G->cell_start[c]; }
    G->cell_start[MAX_CELLS] = run;
    G->sorted = (float*)malloc(sizeof(float) * 4 * (n > 0 ? n : 1));
    /* the GPU fills cells with atomics (any order inside a cell): fill them back to front to make sure order does not matter */
    for (int k = n - 1; k >= 0; --k) {
        int pos = hist[cell[k]]++;
        float* s = G->sorted + 4 * (size_t)pos;
        s[0] = xyz[3 * k]; s[1] = xyz[3 * k + 1]; s[2] = xyz[3 * k + 2];
        int32_t kk = k; memcpy(&s[3], &kk, 4);
    }
    free(hist); free(cell);
}

static uint64_t nn_key(float d, int k) { uint32_t b; memcpy(&b, &d, 4); return ((uint64_t)b << 32) | (uint32_t)k; }
static float key_d(uint64_t k) { uint32_t b = (uint32_t)(k >> 32); float d; memcpy(&d, &b, 4); return d; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* one cloud: unknown (n,3), known (m,3); ucell / kcell: the min_cell arguments of the two grids.
 * dist2 (n,3), idx (n,3) out; returns the total number of candidate visits (work measure), passes_out = growth rounds summed */
static long long coop_core(int n, int m, const float* unknown, const float* known, const int* order, float kcell, float* dist2, int* idx,
                           long long* passes_out) {
    Grid K;
    grid_build(m, known, kcell, &K);
    const Hdr H = K.H;
    const uint64_t EMPTY = 0x7F80000000000000ull;
    long long visits = 0, passes = 0;
    for (int t0 = 0; t0 < n; t0 += 32) {
        int live[32], pt[32], cx[32], cy[32], cz[32];
        float ux[32], uy[32], uz[32];
        uint64_t k1[32], k2[32], k3[32];
        int X0 = 0x7FFFFFFF, X1 = -1, Y0 = 0x7FFFFFFF, Y1 = -1, Z0 = 0x7FFFFFFF, Z1 = -1;
        for (int l = 0; l < 32; ++l) {
            live[l] = t0 + l < n; pt[l] = 0; ux[l] = uy[l] = uz[l] = 0.f; cx[l] = cy[l] = cz[l] = 0;
            k1[l] = k2[l] = k3[l] = EMPTY;
            if (live[l]) {
                int32_t kk = order[t0 + l]; pt[l] = kk;
                ux[l] = unknown[3 * kk]; uy[l] = unknown[3 * kk + 1]; uz[l] = unknown[3 * kk + 2];
                cx[l] = cell_coord(ux[l], H.ox, H.inv_h, H.dx); cy[l] = cell_coord(uy[l], H.oy, H.inv_h, H.dy); cz[l] = cell_coord(uz[l], H.oz, H.inv_h, H.dz);
                X0 = imin(X0, cx[l]); X1 = imax(X1, cx[l]); Y0 = imin(Y0, cy[l]); Y1 = imax(Y1, cy[l]); Z0 = imin(Z0, cz[l]); Z1 = imax(Z1, cz[l]);
            }
        }
        int rx0 = imax(X0 - 1, 0), rx1 = imin(X1 + 1, H.dx - 1), ry0 = imax(Y0 - 1, 0), ry1 = imin(Y1 + 1, H.dy - 1);
        int rz0 = imax(Z0 - 1, 0), rz1 = imin(Z1 + 1, H.dz - 1);
        int px0 = 0, px1 = -1, py0 = 1, py1 = 0, pz0 = 1, pz1 = 0;
        for (;;) {
            ++passes;
            for (int zz = rz0; zz <= rz1; ++zz)
                for (int yy = ry0; yy <= ry1; ++yy) {
                    const int seen_row = zz >= pz0 && zz <= pz1 && yy >= py0 && yy <= py1;
                    const int row = (zz * H.dy + yy) * H.dx;
                    for (int sgm = 0; sgm < (seen_row ? 2 : 1); ++sgm) {
                        const int xa = !seen_row ? rx0 : (sgm == 0 ? rx0 : px1 + 1);
                        const int xb = !seen_row ? rx1 : (sgm == 0 ? px0 - 1 : rx1);
                        if (xa > xb) continue;
                        const int beg = K.cell_start[row + xa], end = K.cell_start[row + xb + 1];
                        for (int j = beg; j < end; ++j) {
                            const float* p = K.sorted + 4 * (size_t)j;
                            int32_t kk; memcpy(&kk, &p[3], 4);
                            for (int l = 0; l < 32; ++l) {          /* dead lanes compute too on the GPU; their results are dropped */
                                const uint64_t key = nn_key(sqdist_ref(ux[l] - p[0], uy[l] - p[1], uz[l] - p[2]), kk);
                                if (key < k3[l]) {
                                    k3[l] = key;
                                    if (k3[l] < k2[l]) { uint64_t tmp = k2[l]; k2[l] = k3[l]; k3[l] = tmp; }
                                    if (k2[l] < k1[l]) { uint64_t tmp = k1[l]; k1[l] = k2[l]; k2[l] = tmp; }
                                }
                            }
                            ++visits;
                        }
                    }
                }
            const int whole = rx0 == 0 && ry0 == 0 && rz0 == 0 && rx1 == H.dx - 1 && ry1 == H.dy - 1 && rz1 == H.dz - 1;
            if (whole) break;
            int all_done = 1;
            for (int l = 0; l < 32; ++l) {
                float bound = INFINITY;
                if (rx0 > 0) bound = fminf(bound, ux[l] - (H.ox + (float)rx0 * H.h));
                if (rx1 < H.dx - 1) bound = fminf(bound, (H.ox + (float)(rx1 + 1) * H.h) - ux[l]);
                if (ry0 > 0) bound = fminf(bound, uy[l] - (H.oy + (float)ry0 * H.h));
                if (ry1 < H.dy - 1) bound = fminf(bound, (H.oy + (float)(ry1 + 1) * H.h) - uy[l]);
                if (rz0 > 0) bound = fminf(bound, uz[l] - (H.oz + (float)rz0 * H.h));
                if (rz1 < H.dz - 1) bound = fminf(bound, (H.oz + (float)(rz1 + 1) * H.h) - uz[l]);
                bound = fmaxf(bound, 0.f);
                const int done = !live[l] || key_d(k3[l]) < bound * bound * 0.998f;
                all_done &= done;
            }
            if (all_done) break;
            px0 = rx0; px1 = rx1; py0 = ry0; py1 = ry1; pz0 = rz0; pz1 = rz1;
            rx0 = imax(rx0 - 1, 0); rx1 = imin(rx1 + 1, H.dx - 1);
            ry0 = imax(ry0 - 1, 0); ry1 = imin(ry1 + 1, H.dy - 1);
            rz0 = imax(rz0 - 1, 0); rz1 = imin(rz1 + 1, H.dz - 1);
        }
        for (int l = 0; l < 32; ++l)
            if (live[l]) {
                float* od = dist2 + 3 * (size_t)pt[l]; int* oi = idx + 3 * (size_t)pt[l];
                od[0] = key_d(k1[l]); od[1] = key_d(k2[l]); od[2] = key_d(k3[l]);
                oi[0] = (int)(uint32_t)k1[l]; oi[1] = (int)(uint32_t)k2[l]; oi[2] = (int)(uint32_t)k3[l];
            }
    }
    free(K.sorted);
    if (passes_out) *passes_out = passes;
    return visits;
}

long long three_nn_coop_emul(int n, int m, const float* unknown, const float* known, float ucell, float kcell, float* dist2, int* idx,
                             long long* passes_out) {
    Grid U;
    grid_build(n, unknown, ucell, &U);
    int* order = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int t = 0; t < n; ++t) { int32_t kk; memcpy(&kk, &U.sorted[4 * (size_t)t + 3], 4); order[t] = kk; }
    const long long v = coop_core(n, m, unknown, known, order, kcell, dist2, idx, passes_out);
    free(order); free(U.sorted);
    return v;
}

/* the same search with an explicit processing order of the unknown points (to study other orders, e.g. Morton) */
long long three_nn_coop_emul_order(int n, int m, const float* unknown, const float* known, const int* order, float kcell, float* dist2,
                                   int* idx, long long* passes_out) {
    return coop_core(n, m, unknown, known, order, kcell, dist2, idx, passes_out);
}
