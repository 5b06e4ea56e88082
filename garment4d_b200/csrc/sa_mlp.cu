// Grouped shared-MLP + max-pool on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces, for one scale of a set-abstraction module in eval mode (pointnet2_modules.py:37-51):
//     grouping_operation(xyz) - centroid ; grouping_operation(features) ; cat        (pointnet2_utils.py:250-258)
//     3 x [cuDNN 1x1 Conv2d -> BatchNorm2d -> ReLU]                                  (pytorch_utils.py:5-32)
//     F.max_pool2d(kernel=[1, nsample]) ; squeeze ; (torch.cat over scales)          (pointnet2_modules.py:42-55)
// i.e. 9+ kernels that each stream the (B, C, npoint, nsample) activation through HBM, with ONE persistent,
// warp-specialised kernel that never materialises the grouped tensor:
//
//   producer warps (8)  one thread per row of the 128-row tile (= 128/nsample whole neighbourhoods): idx -> point id,
//                       relative xyz (fp32 subtract, then hi/lo fp16 split so the geometry enters layer 1 at ~fp32
//                       precision), and the neighbour's fp16 feature row streamed 32 B at a time into a ring of
//                       16-channel K-slices in shared memory, already in the UMMA canonical K-major layout.
//                       full/empty mbarriers per ring slot; the producers run up to a whole tile ahead.
//   MMA warp (1 lane)   layer 1: D1[128 x c1] += slice . W1_slice^T, one tcgen05.mma (kind::f16, fp32 accumulate in
//                       TMEM) per slice as it lands; tcgen05.commit frees the slot.
//                       layer 2: D2[128 x c2] = H1 . W2^T ;  layer 3 TRANSPOSED: D3[c3 x 128] = W3 . H2^T, so that
//                       channels sit on TMEM lanes, positions on columns, and the max over a neighbourhood is a max
//                       over nsample consecutive columns inside ONE thread.
//   epilogue warps (8)  tcgen05.ld -> +bias, ReLU, fp16 (cvt.rn.relu.satfinite) -> shared (canonical layout) for
//                       layers 1-2; layer 3: neighbourhood max, +bias, ReLU (max and the monotone bias+ReLU commute),
//                       written channel-major fp32 at the scale's channel offset (fuses torch.cat) and, optionally,
//                       point-major fp16 for the next level's gather.
//   Folded weights of all three layers stay resident in shared memory (one cp.async.bulk / TMA bulk copy per CTA).
//
// Shared-memory operand layout ("canonical K-major, no swizzle", cute::UMMA::LayoutType::SWIZZLE_NONE):
//   element (row r, k) of an operand with R rows lives at byte ((k/8)*R + r)*16 + (k%8)*2
//   => core matrix = 8 rows x 16 B contiguous; SBO (next 8-row group) = 128 B; LBO (next K chunk) = R*16 B.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "umma.cuh"
#include "garment4d_b200.h"

namespace g4d {

constexpr int TILE_M = 128;
constexpr int XYZ_SLOTS = 9;                 // [hi(3) | lo(3) | hi(3)] against weights [wh | wh | wl]
// Warp groups per role, WG (template parameter of the kernel): WG x 4 epilogue warps + 1 MMA warp + WG x 4 producer warps.
//   WG = 2 (544 threads, 1 CTA per SM): two epilogue warps per TMEM lane quadrant, two producer groups on alternate tiles;
//   WG = 1 (288 threads, 2 CTAs per SM, half the TMEM each): two independent producer -> MMA -> epilogue chains per SM.  The
//          chain is latency-bound (every layer is MMA issue -> commit -> epilogue -> arrive), so the second CTA fills the bubbles.
__host__ __device__ constexpr int sa_threads(int wg) { return (8 * wg + 1) * 32; }
constexpr int SLICE_BYTES = TILE_M * 16 * 2;  // one K-slice: 128 rows x 16 channels fp16 = 4 KB
constexpr int MAX_RING = 16;

struct SaMlpLayout {
    int k0, c1, c2, c3, c3p, nb3, nslices, ring, tb, wg;        // tb = tiles per batch (hand-off latency amortised over tb tiles)
    uint32_t h_bytes, cstride;                              // per-tile H buffer bytes, per-tile TMEM column stride
    uint32_t off_w1, off_w2, off_w3, off_b1, off_b2, off_b3, blob_bytes;   // inside the parameter blob == smem image
    uint32_t off_h, off_ring, off_bar, total_smem;
    uint32_t tmem_cols;
};

static inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

static bool make_layout(const g4d_sa_mlp_desc* d, SaMlpLayout* L, const char** why) {
    static const char* msgs[] = {
        "sa_mlp: c1 and c2 must be multiples of 16 in [16, 256]", "sa_mlp: c3 must be in [1, 256]",
        "sa_mlp: nsample must be 8, 16, 32, 64 or 128", "sa_mlp: c_in must be a non-negative multiple of 8",
        "sa_mlp: k0 does not match g4d_sa_mlp_k0(c_in)", "sa_mlp: shared memory footprint exceeds 227 KB"};
    if (d->c1 < 16 || d->c1 > 256 || d->c1 % 16 || d->c2 < 16 || d->c2 > 256 || d->c2 % 16) { *why = msgs[0]; return false; }
    if (d->c3 < 1 || d->c3 > 256) { *why = msgs[1]; return false; }
    if (!(d->nsample == 8 || d->nsample == 16 || d->nsample == 32 || d->nsample == 64 || d->nsample == 128)) { *why = msgs[2]; return false; }
    if (d->c_in < 0 || d->c_in % 8) { *why = msgs[3]; return false; }
    if (d->k0 != g4d_sa_mlp_k0(d->c_in)) { *why = msgs[4]; return false; }
    L->k0 = d->k0; L->c1 = d->c1; L->c2 = d->c2; L->c3 = d->c3;
    L->c3p = d->c3 <= 128 ? 128 : 256;
    L->nb3 = L->c3p / 128;
    L->nslices = L->k0 / 16;
    uint32_t o = 0;
    L->off_w1 = o; o += (uint32_t)L->k0 * L->c1 * 2;
    L->off_w2 = o; o += (uint32_t)L->c1 * L->c2 * 2;
    L->off_w3 = o; o += (uint32_t)L->c2 * L->c3p * 2;
    L->off_b1 = o; o += (uint32_t)L->c1 * 4;
    L->off_b2 = o; o += (uint32_t)L->c2 * 4;
    L->off_b3 = o; o += (uint32_t)L->c3p * 4;
    L->blob_bytes = o;                                   // multiple of 16 by construction
    const int hk = L->c1 > L->c2 ? L->c1 : L->c2;
    L->h_bytes = (uint32_t)TILE_M * hk * 2;
    L->cstride = (uint32_t)(128 * L->nb3) > (uint32_t)hk ? 128 * L->nb3 : hk;
    L->off_h = round_up(o, 128);
    // Plan for `occ` resident CTAs per SM (occ = 2 <=> WG = 1): shared memory and the 512 TMEM columns are split evenly.
    auto plan = [&](int occ) -> bool {
        const uint32_t budget = (227u * 1024u) / occ - 1024u - 512u;             // 1 KB per CTA is reserved by the driver
        if (L->off_h + L->h_bytes + 2u * SLICE_BYTES > budget) return false;
        // tiles per batch: as many as TMEM and shared memory allow, up to 4 per SM (env G4D_SA_TB overrides)
        int tb = (512 / occ) / (int)L->cstride;
        if (tb < 1) return false;
        if (tb > 4 / occ) tb = 4 / occ;
        if (tb == 3) tb = 2;
        if (const char* e = getenv("G4D_SA_TB")) { const int v = atoi(e); if (v >= 1 && v <= tb) tb = v; }
        while (tb > 1 && L->off_h + (uint32_t)tb * L->h_bytes + 4u * SLICE_BYTES > budget) tb >>= 1;
        L->tb = tb;
        L->off_ring = L->off_h + (uint32_t)tb * L->h_bytes;
        // ring depth: two tiles' worth of slices when they fit (producers run ahead), at least 2, at most 16
        int ring = 2 * L->nslices;
        if (ring < 4) ring = 4;
        if (ring > MAX_RING) ring = MAX_RING;
        while (ring > 2 && L->off_ring + (uint32_t)ring * SLICE_BYTES > budget) --ring;
        if (occ > 1 && ring < 8 && ring < L->nslices) return false;              // too shallow to hide the gather: use the big CTA
        L->ring = ring;
        L->off_bar = L->off_ring + (uint32_t)ring * SLICE_BYTES;
        L->total_smem = L->off_bar + 8 * (2 * MAX_RING + 4) + 16;
        uint32_t p2 = 32;
        while (p2 < (uint32_t)tb * L->cstride) p2 <<= 1;
        L->tmem_cols = p2;
        L->wg = occ == 1 ? 2 : 1;
        return L->total_smem <= budget + 512u;
    };
    int want_wg = 0;
    if (const char* e = getenv("G4D_SA_WG")) want_wg = atoi(e);
    bool ok = false;
    if (want_wg != 2) ok = plan(2);
    if (!ok) ok = plan(1);
    if (!ok) { *why = msgs[5]; return false; }
    if (L->total_smem > 227 * 1024) { *why = msgs[5]; return false; }
    return true;
}

// NB x 16 consecutive TMEM columns of this thread's lane: all loads in flight, ONE wait
template <int NB>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
    uint32_t r[NB * 16];
#pragma unroll
    for (int b = 0; b < NB; ++b)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[16 * b + 0]), "=r"(r[16 * b + 1]), "=r"(r[16 * b + 2]), "=r"(r[16 * b + 3]), "=r"(r[16 * b + 4]),
              "=r"(r[16 * b + 5]), "=r"(r[16 * b + 6]), "=r"(r[16 * b + 7]), "=r"(r[16 * b + 8]), "=r"(r[16 * b + 9]),
              "=r"(r[16 * b + 10]), "=r"(r[16 * b + 11]), "=r"(r[16 * b + 12]), "=r"(r[16 * b + 13]), "=r"(r[16 * b + 14]),
              "=r"(r[16 * b + 15])
            : "r"(taddr + 16 * b));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NB * 16; ++i) {
        asm volatile("" : "+r"(r[i]));            // pin: no use of r[i] may be scheduled above the wait
        v[i] = __uint_as_float(r[i]);
    }
}


// ---------------------------------------------------------------------------------------------------------

struct SaMlpArgs {
    SaMlpLayout L;
    int c_in, nsample, n, m;
    int total_rows;                  // b * m * nsample (< 2^31, checked on the host)
    int lg_ns, lg_tb;                // log2(nsample), log2(tiles per batch): divisions become shifts
    int ntiles;
    const float* xyz;                // (b, n, 3)
    const float* new_xyz;            // (b, m, 3)
    const int* idx;                  // (b, m, nsample)
    const __half* feat_pm;           // (b, n, c_in) or null
    const unsigned char* params;     // packed blob (device)
    float* out_cm;                   // (b, ctot, m)
    __half* out_pm;                  // (b, m, ctot) or null
    int ctot, coff;
    long long* dbg;                  // optional timeline buffer (g4d_debug_timeline), CTA 0 only
};

// Unstaged channel-major store of one output value (few centroids per batch: nsample >= 64).  Out of line: code size.
__device__ __noinline__ void store_cm_direct(float* out_cm, unsigned cloud, unsigned p, unsigned m, int ctot, int ch, float o) {
    while (p >= m) { p -= m; ++cloud; }
    out_cm[((size_t)cloud * ctot + ch) * m + p] = o;
}

// One instantiation per (nsample, features present), and single call sites / rolled loops for everything that is not the
// inner arithmetic: with every variant a run-time branch and every helper inlined at each use the kernel was 107 KB of
// SASS, three times the 32 KB L1.5 instruction cache, shared by three roles that execute disjoint code.
template <int NS, bool FEAT, int WG>
__global__ void __launch_bounds__(sa_threads(WG), 3 - WG)
sa_mlp_max_kernel(const SaMlpArgs a) {
    constexpr int SA_EPI_WARPS = 4 * WG, SA_THREADS = sa_threads(WG);
    constexpr int LG_NS = NS == 8 ? 3 : NS == 16 ? 4 : NS == 32 ? 5 : NS == 64 ? 6 : 7;
    extern __shared__ __align__(128) unsigned char smem[];
    const SaMlpLayout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* hbuf = smem + L.off_h;
    const float* b1 = reinterpret_cast<const float*>(smem + L.off_b1);
    const float* b2 = reinterpret_cast<const float*>(smem + L.off_b2);
    const float* b3 = reinterpret_cast<const float*>(smem + L.off_b3);
    // barriers: [0] weights, [1] d_full (MMA -> epilogue), [2] epi_done (epilogue -> MMA), [3] tmem slot,
    //           [4 .. 4+MAX_RING) full[r], [4+MAX_RING .. 4+2*MAX_RING) empty[r]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    const uint32_t bar_w = bar0, bar_dfull = bar0 + 8, bar_epi = bar0 + 16, tmem_slot = bar0 + 24;
    const uint32_t bar_full = bar0 + 32, bar_empty = bar0 + 32 + 8 * MAX_RING;
    const uint32_t s_w1 = smem_u32(smem + L.off_w1), s_w2 = smem_u32(smem + L.off_w2), s_w3 = smem_u32(smem + L.off_w3);
    const uint32_t s_h = smem_u32(hbuf), s_ring = smem_u32(smem + L.off_ring);

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_dfull, 1);
        mbar_init(bar_epi, SA_EPI_WARPS);
        for (int r = 0; r < L.ring; ++r) { mbar_init(bar_full + 8 * r, 4);     /* the 4 warps of the producer group that owns the slot's tile */ mbar_init(bar_empty + 8 * r, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, L.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L.off_bar + 24);
    if (tid == 0) {
        mbar_expect_tx(bar_w, L.blob_bytes);
        bulk_g2s(smem_u32(smem), a.params, L.blob_bytes, bar_w);      // folded weights + biases, once per CTA
    }

    const int S = L.nslices, RING = L.ring;

    if (warp >= SA_EPI_WARPS + 1) {
        // =========================== PRODUCERS: one thread per tile row ===================================
        // Two groups of 4 warps take alternate tiles of this CTA's tile sequence (ring slots are addressed by the
        // global slice counter seq * S + s, so both groups fill the ring concurrently and in order).  Inside a group
        // the loads are software-pipelined: idx two tiles ahead, coordinates one tile ahead, feature chunks one batch
        // (8 x 16 B per thread) ahead of the stores.  All index arithmetic is 32-bit (sizes checked on the host) and
        // nsample / tiles-per-batch are powers of two; the code is written flat on purpose (no helper objects: the
        // per-tile instruction count of this role bounds the small layers).
        const int pw = warp - (SA_EPI_WARPS + 1);
        const int grp = pw >> 2;                                 // 0 .. WG-1
        const int r = (pw & 3) * 32 + lane;                      // tile row 0..127
        const int nchunk_feat = FEAT ? (a.c_in >> 3) : 0;
        const int lg_tb = a.lg_tb, tbm = L.tb - 1;
        constexpr int lg_ns = LG_NS;
        const int bx = (int)blockIdx.x, gx = (int)gridDim.x;
        const unsigned um = (unsigned)a.m;
#define G4D_TILE_OF(q) ((((bx + ((q) >> lg_tb) * gx)) << lg_tb) + ((q) & tbm))
        int seq = grp;
        int t_cur = G4D_TILE_OF(seq), t_nxt = G4D_TILE_OF(seq + WG);
        // pipeline registers: source index of the next two tiles, geometry + point id of the next tile
        int src_n = 0, src_nn = 0;
        float npx = 0.f, npy = 0.f, npz = 0.f, nqx = 0.f, nqy = 0.f, nqz = 0.f;
        unsigned npt = 0;
        if (t_cur < a.ntiles && t_cur * TILE_M + r < a.total_rows) src_n = __ldg(a.idx + t_cur * TILE_M + r);
        if (t_nxt < a.ntiles && t_nxt * TILE_M + r < a.total_rows) src_nn = __ldg(a.idx + t_nxt * TILE_M + r);
        if (t_cur < a.ntiles && t_cur * TILE_M + r < a.total_rows) {
            const unsigned gp = (unsigned)(t_cur * TILE_M + r) >> lg_ns;
            npt = (gp / um) * (unsigned)a.n + (unsigned)src_n;
            const float* p = a.xyz + (size_t)npt * 3;
            const float* q = a.new_xyz + (size_t)gp * 3;
            npx = __ldg(p); npy = __ldg(p + 1); npz = __ldg(p + 2);
            nqx = __ldg(q); nqy = __ldg(q + 1); nqz = __ldg(q + 2);
        }
        while (t_cur < a.ntiles) {
            const int tile = t_cur;
            const bool live = tile * TILE_M + r < a.total_rows;
            const float dx = npx - nqx, dy = npy - nqy, dz = npz - nqz;
            const unsigned pt = npt;
            const uint32_t it0 = (uint32_t)seq * (uint32_t)S;      // global slice counter of this tile's first slice
            // ---- advance the pipeline: issue the loads of the following tiles before touching this one
            seq += WG;
            t_cur = t_nxt;
            t_nxt = G4D_TILE_OF(seq + WG);
            src_n = src_nn;
            src_nn = 0;
            if (t_nxt < a.ntiles && t_nxt * TILE_M + r < a.total_rows) src_nn = __ldg(a.idx + t_nxt * TILE_M + r);
            if (t_cur < a.ntiles && t_cur * TILE_M + r < a.total_rows) {
                const unsigned gp = (unsigned)(t_cur * TILE_M + r) >> lg_ns;
                npt = (gp / um) * (unsigned)a.n + (unsigned)src_n;
                const float* p = a.xyz + (size_t)npt * 3;
                const float* q = a.new_xyz + (size_t)gp * 3;
                npx = __ldg(p); npy = __ldg(p + 1); npz = __ldg(p + 2);
                nqx = __ldg(q); nqy = __ldg(q + 1); nqz = __ldg(q + 2);
            }
            // ---- this tile: relative xyz, hi/lo split.  slots: hi.x hi.y hi.z lo.x lo.y lo.z hi.x hi.y | hi.z 0 ... 0
            uint4 xc0 = make_uint4(0, 0, 0, 0), xc1 = make_uint4(0, 0, 0, 0);
            if (live) {
                const __half hx = __float2half_rn(dx), hy = __float2half_rn(dy), hz = __float2half_rn(dz);
                const uint32_t uhx = __half_as_ushort(hx), uhy = __half_as_ushort(hy), uhz = __half_as_ushort(hz);
                const uint32_t ulx = __half_as_ushort(__float2half_rn(dx - __half2float(hx))),
                               uly = __half_as_ushort(__float2half_rn(dy - __half2float(hy))),
                               ulz = __half_as_ushort(__float2half_rn(dz - __half2float(hz)));
                xc0 = make_uint4(uhx | (uhy << 16), uhz | (ulx << 16), uly | (ulz << 16), uhx | (uhy << 16));
                xc1 = make_uint4(uhz, 0, 0, 0);
            }
            uint32_t it = it0;
            if (!FEAT) {
                // xyz-only level: a single slice per tile
                const uint32_t slot = it % RING, ph = (it / RING) & 1;
                mbar_wait_relaxed(bar_empty + 8 * slot, ph ^ 1);
                uint4* dst = reinterpret_cast<uint4*>(smem + L.off_ring + (size_t)slot * SLICE_BYTES);
                dst[r] = xc0;
                dst[TILE_M + r] = xc1;
                fence_proxy_async();                              // generic-proxy stores -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * slot);
            } else {
                // Feature levels: every 16-byte chunk of the neighbour's fp16 row goes global -> shared with cp.async (LDGSTS), one
                // commit group per K-slice, ALL slices of the tile in flight at once (no register staging: 128 threads x 2S copies
                // = the whole 128 x k0 tile outstanding, which is what hides the L2/HBM latency of the gather); then the slices are
                // handed to the MMA warp in order as their groups complete.
                const char* srcrow = reinterpret_cast<const char*>(a.feat_pm + (size_t)pt * a.c_in);
                // (waves of at most min(S, RING) slices: a wave must fit the ring, or waiting for its own slots would deadlock)
                const int wave = S < RING ? S : RING;
                uint32_t slot = it % RING, ph = (it / RING) & 1;               // one division per tile; then walked
#pragma unroll 1
                for (int w0 = 0; w0 < S; w0 += wave) {
                const int w1 = (w0 + wave < S) ? w0 + wave : S;
                const uint32_t slot_w = slot;
#pragma unroll 1
                for (int sl = w0; sl < w1; ++sl) {
                    mbar_wait_relaxed(bar_empty + 8 * slot, ph ^ 1);           // slot free (first lap passes at once)
                    const uint32_t sdst = s_ring + slot * SLICE_BYTES + (uint32_t)r * 16;
                    uint4* gdst = reinterpret_cast<uint4*>(smem + L.off_ring + (size_t)slot * SLICE_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = 2 * sl + h;                               // 16-byte chunk index along K
                        if (c < nchunk_feat)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst + h * TILE_M * 16), "l"(srcrow + (size_t)c * 16),
                                         "r"(live ? 16 : 0) : "memory");
                        else gdst[h * TILE_M + r] = (c == nchunk_feat) ? xc0 : ((c == nchunk_feat + 1) ? xc1 : make_uint4(0, 0, 0, 0));
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    if (++slot == (uint32_t)RING) { slot = 0; ph ^= 1; }
                }
                uint32_t slot_a = slot_w;
#pragma unroll 1
                for (int sl = w0; sl < w1; ++sl) {
                    switch (w1 - 1 - sl) {                                      // groups that may still be pending: the younger slices
                        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
                        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
                        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
                        case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
                        case 8: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
                        case 9: asm volatile("cp.async.wait_group 9;" ::: "memory"); break;
                        case 10: asm volatile("cp.async.wait_group 10;" ::: "memory"); break;
                        case 11: asm volatile("cp.async.wait_group 11;" ::: "memory"); break;
                        case 12: asm volatile("cp.async.wait_group 12;" ::: "memory"); break;
                        case 13: asm volatile("cp.async.wait_group 13;" ::: "memory"); break;
                        case 14: asm volatile("cp.async.wait_group 14;" ::: "memory"); break;
                        default: asm volatile("cp.async.wait_group 15;" ::: "memory"); break;
                    }
                    fence_proxy_async();                                        // copies + stores -> visible to tcgen05.mma
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * slot_a);
                    if (++slot_a == (uint32_t)RING) slot_a = 0;
                }
                }
            }
        }
#undef G4D_TILE_OF
    } else if (warp == SA_EPI_WARPS) {
        // =========================== MMA ISSUER (one lane) ==================================================
        if (lane == 0) {
            mbar_wait(bar_w, 0);
            uint32_t slot = 0, ph = 0, nepi = 0;                 // ring position / phase; epilogue hand-offs waited for
            const uint32_t ring_lo = desc_lo(s_ring, TILE_M * 16), w1_lo = desc_lo(s_w1, L.c1 * 16), w1_step = (uint32_t)(2 * L.c1 * 16) >> 4;
            const uint32_t w2_lo = desc_lo(s_w2, L.c2 * 16), w2_step = (uint32_t)(2 * L.c2 * 16) >> 4;
            const uint32_t w3_step = (uint32_t)(2 * L.c3p * 16) >> 4, h_step = (uint32_t)(2 * TILE_M * 16) >> 4;
            const uint32_t idesc1 = umma_idesc(TILE_M, L.c1), idesc2 = umma_idesc(TILE_M, L.c2), idesc3 = umma_idesc(128, TILE_M);
            for (long long tl = (long long)blockIdx.x * L.tb; tl < a.ntiles; tl += (long long)gridDim.x * L.tb) {
                const int nb = (int)((a.ntiles - tl) < L.tb ? (a.ntiles - tl) : L.tb);      // tiles in this batch
                // ---- layer 1: needs TMEM drained by the previous batch's epilogue 3
                mbar_wait_spin(bar_epi, (nepi + 1) & 1); ++nepi;
                tc_fence_after();
                long long* dbg = (a.dbg && blockIdx.x == 0 && nepi <= 3 * 24) ? a.dbg + (nepi / 3) * 16 : nullptr;
                if (dbg) dbg[0] = clock64();
                for (int bi = 0; bi < nb; ++bi) {
                    uint32_t blo = w1_lo;
                    for (int s = 0; s < S; ++s) {
                        if constexpr (FEAT) mbar_wait(bar_full + 8 * slot, ph);   // gathers in flight: leave the issue slots to the producers
                        else mbar_wait_spin(bar_full + 8 * slot, ph);
                        tc_fence_after();
                        umma_f16(tmem + bi * L.cstride, desc64(ring_lo + slot * (SLICE_BYTES >> 4)), desc64(blo), idesc1, s > 0);
                        umma_commit(bar_empty + 8 * slot);        // slot reusable once this (and earlier) MMAs retire
                        blo += w1_step;
                        if (++slot == (uint32_t)RING) { slot = 0; ph ^= 1; }
                    }
                }
                umma_commit(bar_dfull);
                if (dbg) dbg[1] = clock64();
                // ---- layer 2: needs H1 written by epilogue 1
                mbar_wait_spin(bar_epi, (nepi + 1) & 1); ++nepi;
                tc_fence_after();
                if (dbg) dbg[2] = clock64();
                for (int bi = 0; bi < nb; ++bi) {
                    uint32_t alo = desc_lo(s_h + bi * L.h_bytes, TILE_M * 16), blo = w2_lo;
                    for (int k = 0; k < L.c1 / 16; ++k) {
                        umma_f16(tmem + bi * L.cstride, desc64(alo), desc64(blo), idesc2, k > 0);
                        alo += h_step; blo += w2_step;
                    }
                }
                umma_commit(bar_dfull);
                if (dbg) dbg[3] = clock64();
                // ---- layer 3, transposed: D3[c3p x 128] = W3 . H2^T
                mbar_wait_spin(bar_epi, (nepi + 1) & 1); ++nepi;
                tc_fence_after();
                if (dbg) dbg[4] = clock64();
                for (int bi = 0; bi < nb; ++bi)
                    for (int j = 0; j < L.nb3; ++j) {
                        uint32_t alo = desc_lo(s_w3 + (uint32_t)j * 128 * 16, L.c3p * 16), blo = desc_lo(s_h + bi * L.h_bytes, TILE_M * 16);
                        for (int k = 0; k < L.c2 / 16; ++k) {
                            umma_f16(tmem + bi * L.cstride + j * 128, desc64(alo), desc64(blo), idesc3, k > 0);
                            alo += w3_step; blo += h_step;
                        }
                    }
                umma_commit(bar_dfull);
                if (dbg) dbg[5] = clock64();
            }
        }
        __syncwarp();                                            // reconverge before the block-wide barrier below
    } else {
        // =========================== EPILOGUE: warps q and q+4 own TMEM lanes 32q .. 32q+31 ==================
        mbar_wait(bar_w, 0);                                      // biases live in the weight blob
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;                         // tile row (layers 1-2) / channel within block (layer 3)
        const uint32_t lane_taddr = tmem + ((uint32_t)(quad * 32) << 16);
        constexpr int groups_per_tile = TILE_M / NS;
        uint32_t nd = 0;                                          // d_full hand-offs waited for
        // Work split between the two warps of a quadrant: with tb >= 2 tiles per batch each takes alternate tiles (all
        // columns); with a single tile per batch they split its columns.
        const bool split_cols = WG == 2 && L.tb == 1;
        auto relu_to_h = [&](int bi, int ncols, const float* bias) {
            const int units = ncols / 16;
            const int u0 = (!split_cols || half == 0) ? 0 : (units + 1) / 2;
            const int u1 = (!split_cols) ? units : (half == 0 ? (units + 1) / 2 : units);
            const uint32_t ta = lane_taddr + bi * L.cstride;
            uint4* hd = reinterpret_cast<uint4*>(hbuf + (size_t)bi * L.h_bytes);
            int u = u0;
            for (; u + 2 <= u1; u += 2) {
                float v[32];
                tmem_ld_cols<2>(ta + 16 * u, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t h[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        h[i] = pack_relu_f16x2(v[8 * q + 2 * i] + bias[16 * u + 8 * q + 2 * i], v[8 * q + 2 * i + 1] + bias[16 * u + 8 * q + 2 * i + 1]);
                    hd[(size_t)(2 * u + q) * TILE_M + row] = make_uint4(h[0], h[1], h[2], h[3]);
                }
            }
            for (; u < u1; ++u) {
                float v[16];
                tmem_ld_cols<1>(ta + 16 * u, v);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint32_t h[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        h[i] = pack_relu_f16x2(v[8 * q + 2 * i] + bias[16 * u + 8 * q + 2 * i], v[8 * q + 2 * i + 1] + bias[16 * u + 8 * q + 2 * i + 1]);
                    hd[(size_t)(2 * u + q) * TILE_M + row] = make_uint4(h[0], h[1], h[2], h[3]);
                }
            }
        };
        // centroid gp = gp0 + g of this tile; (cloud, p) of gp0 is divided out once per tile, then walked
        // Channel-major fp32 output: with lane = channel a direct store would touch 32 different sectors per instruction
        // (4 useful bytes each).  When it fits, the batch's (channel x centroid) block is staged in the H buffers (free
        // during epilogue 3) and written out with lanes along the centroid index: 64-128 contiguous bytes per channel.
        const int GS = L.tb * groups_per_tile;                   // centroids per batch
        const bool staged = (size_t)L.c3 * (GS + 1) * 4 <= (size_t)L.tb * L.h_bytes && GS >= 8;
        float* stage = reinterpret_cast<float*>(hbuf);
        unsigned e_gp0 = 0, e_cloud0 = 0, e_p0 = 0;
        int e_bi = 0;
        auto emit = [&](int ch, int g, float mx, float bias) {
            const unsigned gp = e_gp0 + (unsigned)g;
            if (ch < L.c3 && (gp << LG_NS) < (unsigned)a.total_rows) {
                const float o = fmaxf(mx + bias, 0.f);
                if (staged) stage[ch * (GS + 1) + e_bi * groups_per_tile + g] = o;
                else store_cm_direct(a.out_cm, e_cloud0, e_p0 + (unsigned)g, (unsigned)a.m, a.ctot, a.coff + ch, o);
                if (a.out_pm) a.out_pm[(size_t)gp * a.ctot + a.coff + ch] = __float2half_rn(fminf(o, 65504.f));
            }
        };
        auto set_tile = [&](int tile) {
            e_gp0 = (unsigned)tile * (unsigned)groups_per_tile;
            e_cloud0 = e_gp0 / (unsigned)a.m;
            e_p0 = e_gp0 - e_cloud0 * (unsigned)a.m;
        };
        // 64 consecutive positions (columns) of channel block j of tile bi: neighbourhood max + emit
        auto max_emit64 = [&](int bi, int tile, int j, int colhalf) {
            const int ch = j * 128 + row;
            const float bias = b3[ch];
            float v[64];
            tmem_ld_cols<4>(lane_taddr + bi * L.cstride + j * 128 + 64 * colhalf, v);
            set_tile(tile); e_bi = bi;
            const int gp0 = (64 * colhalf) >> LG_NS;
            if constexpr (NS == 8) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float mx = v[8 * g];
#pragma unroll
                    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[8 * g + i]);
                    emit(ch, gp0 + g, mx, bias);
                }
            } else if constexpr (NS == 16) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float mx = v[16 * g];
#pragma unroll
                    for (int i = 1; i < 16; ++i) mx = fmaxf(mx, v[16 * g + i]);
                    emit(ch, gp0 + g, mx, bias);
                }
            } else if constexpr (NS == 32) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    float mx = v[32 * g];
#pragma unroll
                    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[32 * g + i]);
                    emit(ch, gp0 + g, mx, bias);
                }
            } else {
                float mx = v[0];
#pragma unroll
                for (int i = 1; i < 64; ++i) mx = fmaxf(mx, v[i]);
                emit(ch, gp0, mx, bias);
            }
        };
        auto max_emit128 = [&](int bi, int tile, int j) {            // nsample == 128: one neighbourhood per tile
            const int ch = j * 128 + row;
            float mx = -INFINITY;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float v[64];
                tmem_ld_cols<4>(lane_taddr + bi * L.cstride + j * 128 + 64 * hh, v);
#pragma unroll
                for (int i = 0; i < 64; ++i) mx = fmaxf(mx, v[i]);
            }
            set_tile(tile); e_bi = bi;
            emit(ch, 0, mx, b3[ch]);
        };
        const int t_first = split_cols ? 0 : half, t_step = split_cols ? 1 : WG;
        for (long long tl = (long long)blockIdx.x * L.tb; tl < a.ntiles; tl += (long long)gridDim.x * L.tb) {
            const int nb = (int)((a.ntiles - tl) < L.tb ? (a.ntiles - tl) : L.tb);
            // ---- epilogues 1 and 2: D -> bias + ReLU -> fp16 -> H (H1 is dead when d_full fires for layer 2: MMA 2 has completed).
            // One body for both layers (code size).
            long long* dbg = (a.dbg && blockIdx.x == 0 && tid == 0 && nd < 3 * 24) ? a.dbg + (nd / 3) * 16 + 8 : nullptr;
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                mbar_wait(bar_dfull, nd & 1); ++nd;
                tc_fence_after();
                if (dbg) dbg[2 * layer] = clock64();
                const int ncols = layer ? L.c2 : L.c1;
                const float* bias = layer ? b2 : b1;
#pragma unroll 1
                for (int bi = t_first; bi < nb; bi += t_step) relu_to_h(bi, ncols, bias);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_epi);
                if (dbg) dbg[2 * layer + 1] = clock64();
            }
            // ---- epilogue 3: lane = output channel; max over each neighbourhood's nsample consecutive columns
            mbar_wait(bar_dfull, nd & 1); ++nd;
            tc_fence_after();
            if (dbg) dbg[4] = clock64();
            if constexpr (NS <= 64) {
                // work items (tile, channel block, column half); with one tile per batch the two warps of a quadrant split the halves
                const int h_first = split_cols ? half : 0, h_step = split_cols ? 2 : 1;
#pragma unroll 1
                for (int bi = t_first; bi < nb; bi += t_step)
#pragma unroll 1
                    for (int j = 0; j < L.nb3; ++j)
#pragma unroll 1
                        for (int ch2 = h_first; ch2 < 2; ch2 += h_step) max_emit64(bi, (int)tl + bi, j, ch2);
            } else {
                const int j_first = split_cols ? half : 0, j_step = split_cols ? 2 : 1;
#pragma unroll 1
                for (int bi = t_first; bi < nb; bi += t_step)
#pragma unroll 1
                    for (int j = j_first; j < L.nb3; j += j_step) max_emit128(bi, (int)tl + bi, j);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_epi);                  // TMEM drained: the next batch's layer 1 may start
            if (dbg) dbg[5] = clock64();
            if (staged) {
                asm volatile("bar.sync 1, %0;" ::"n"(SA_EPI_WARPS * 32) : "memory");     // the 8 epilogue warps only
                const unsigned gpb = (unsigned)tl * (unsigned)groups_per_tile;            // first centroid of the batch
                const unsigned cloudb = gpb / (unsigned)a.m, pb = gpb - cloudb * (unsigned)a.m;
                const int G = nb * groups_per_tile;
                const int et = warp * 32 + lane;                  // 0..255
                for (int i = et; i < L.c3 * G; i += SA_EPI_WARPS * 32) {
                    const int ch = i / G, g = i - ch * G;
                    if (((gpb + (unsigned)g) << LG_NS) < (unsigned)a.total_rows) {
                        unsigned cloud = cloudb, pp = pb + (unsigned)g;
                        while (pp >= (unsigned)a.m) { pp -= (unsigned)a.m; ++cloud; }
                        a.out_cm[((size_t)cloud * a.ctot + a.coff + ch) * a.m + pp] = stage[ch * (GS + 1) + g];
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(SA_EPI_WARPS * 32) : "memory");     // staging area is H again
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, L.tmem_cols);
}

}  // namespace g4d

using namespace g4d;

static long long* g_timeline = nullptr;
// Debug aid: device buffer of >= 25*16 int64 that CTA 0 of the next g4d_sa_mlp_max launches fills with clock64() stamps
// (per batch: MMA lane [0..5], epilogue warp 0 [8..13]); NULL switches it off.
G4D_API void g4d_debug_timeline(void* buf) { g_timeline = (long long*)buf; }

G4D_API int g4d_sa_mlp_k0(int c_in) { return (c_in + XYZ_SLOTS + 15) / 16 * 16; }

G4D_API size_t g4d_sa_mlp_param_bytes(const g4d_sa_mlp_desc* d) {
    SaMlpLayout L; const char* why = nullptr;
    if (!d || !make_layout(d, &L, &why)) { set_error("%s", why ? why : "sa_mlp: null descriptor"); return 0; }
    return L.blob_bytes;
}

// canonical K-major image of a (rows x K) fp16 operand
static void put_canonical(__half* base, int R, int r, int k, float v) {
    base[((size_t)(k / 8) * R + r) * 8 + (k % 8)] = __float2half_rn(v);
}

G4D_API int g4d_sa_mlp_pack_params(const g4d_sa_mlp_desc* d, const float* w1, const float* b1, const float* w2, const float* b2,
                                   const float* w3, const float* b3, void* blob) {
    SaMlpLayout L; const char* why = nullptr;
    if (!d || !make_layout(d, &L, &why)) return bad_arg(why ? why : "sa_mlp: null descriptor");
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !blob) return bad_arg("sa_mlp_pack_params: null pointer");
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, L.blob_bytes);
    const int cin = d->c_in, ld1 = 3 + cin;     // reference column order: [xyz(3) | features(c_in)]  (pointnet2_utils.py:258)
    __half* W1 = (__half*)(out + L.off_w1);
    for (int o = 0; o < L.c1; ++o) {
        for (int k = 0; k < cin; ++k) put_canonical(W1, L.c1, o, k, w1[(size_t)o * ld1 + 3 + k]);
        for (int j = 0; j < 3; ++j) {
            const float w = w1[(size_t)o * ld1 + j];
            const float wh = __half2float(__float2half_rn(w));
            put_canonical(W1, L.c1, o, cin + j, wh);          // x_hi * w_hi
            put_canonical(W1, L.c1, o, cin + 3 + j, wh);      // x_lo * w_hi
            put_canonical(W1, L.c1, o, cin + 6 + j, w - wh);  // x_hi * w_lo
        }
    }
    __half* W2 = (__half*)(out + L.off_w2);
    for (int o = 0; o < L.c2; ++o)
        for (int k = 0; k < L.c1; ++k) put_canonical(W2, L.c2, o, k, w2[(size_t)o * L.c1 + k]);
    __half* W3 = (__half*)(out + L.off_w3);
    for (int o = 0; o < L.c3; ++o)
        for (int k = 0; k < L.c2; ++k) put_canonical(W3, L.c3p, o, k, w3[(size_t)o * L.c2 + k]);
    memcpy(out + L.off_b1, b1, sizeof(float) * L.c1);
    memcpy(out + L.off_b2, b2, sizeof(float) * L.c2);
    memcpy(out + L.off_b3, b3, sizeof(float) * L.c3);
    return 0;
}

G4D_API int g4d_sa_mlp_max(const g4d_sa_mlp_desc* d, const void* params_dev, int b, int n, int m, const float* xyz,
                           const float* new_xyz, const int* idx, const void* feat_pm, float* out_cm, void* out_pm,
                           int out_c_total, int out_c_off, void* stream) {
    SaMlpArgs a;
    const char* why = nullptr;
    if (!d || !make_layout(d, &a.L, &why)) return bad_arg(why ? why : "sa_mlp: null descriptor");
    if (b < 0 || n <= 0 || m < 0) return bad_arg("sa_mlp_max: bad size");
    if (b == 0 || m == 0) return 0;
    if (!params_dev || !xyz || !new_xyz || !idx || !out_cm || (d->c_in > 0 && !feat_pm)) return bad_arg("sa_mlp_max: null pointer");
    if (out_c_off < 0 || out_c_off + d->c3 > out_c_total) return bad_arg("sa_mlp_max: channel window outside the output");
    if (((uintptr_t)params_dev & 15) || ((uintptr_t)feat_pm & 15)) return bad_arg("sa_mlp_max: params/feat_pm must be 16-byte aligned");
    a.c_in = d->c_in; a.nsample = d->nsample; a.n = n; a.m = m;
    const long long rows = (long long)b * m * d->nsample;
    if (rows > 0x7FFFFF00ll) return bad_arg("sa_mlp_max: b*m*nsample must stay below 2^31");
    a.total_rows = (int)rows;
    a.ntiles = (int)((rows + TILE_M - 1) / TILE_M);
    a.lg_ns = 0; while ((1 << a.lg_ns) < d->nsample) ++a.lg_ns;
    a.lg_tb = 0; while ((1 << a.lg_tb) < a.L.tb) ++a.lg_tb;
    a.xyz = xyz; a.new_xyz = new_xyz; a.idx = idx; a.feat_pm = (const __half*)feat_pm;
    a.params = (const unsigned char*)params_dev;
    a.out_cm = out_cm; a.out_pm = (__half*)out_pm; a.ctot = out_c_total; a.coff = out_c_off;
    a.dbg = g_timeline;

    typedef void (*kern_t)(const SaMlpArgs);
    kern_t kern = nullptr;
    const bool feat = d->c_in > 0;
#define G4D_SA_PICK(NS_) \
    kern = a.L.wg == 1 ? (feat ? sa_mlp_max_kernel<NS_, true, 1> : sa_mlp_max_kernel<NS_, false, 1>) \
                       : (feat ? sa_mlp_max_kernel<NS_, true, 2> : sa_mlp_max_kernel<NS_, false, 2>)
    switch (d->nsample) {
        case 8: G4D_SA_PICK(8); break;
        case 16: G4D_SA_PICK(16); break;
        case 32: G4D_SA_PICK(32); break;
        case 64: G4D_SA_PICK(64); break;
        default: G4D_SA_PICK(128); break;
    }
#undef G4D_SA_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total_smem);
    if (e != cudaSuccess) { set_error("sa_mlp_max: shared memory opt-in (%u B): %s", a.L.total_smem, cudaGetErrorString(e)); return (int)e; }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    const int occ = a.L.wg == 1 ? 2 : 1;                   // resident CTAs per SM the layout was planned for (make_layout)
    long long grid = (long long)sm_count() * occ;
    const long long nbatches = (a.ntiles + a.L.tb - 1) / a.L.tb;
    if (grid > nbatches) grid = nbatches;
    kern<<<(unsigned)grid, sa_threads(a.L.wg), a.L.total_smem, (cudaStream_t)stream>>>(a);
    return finish_launch("g4d sa_mlp_max");
}
