// Grouped shared-MLP + max-pool on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces, for one scale of a set-abstraction module in eval mode (pointnet2_modules.py:37-51):
//     grouping_operation(xyz) - centroid ; grouping_operation(features) ; cat        (pointnet2_utils.py:250-258)
//     3 x [cuDNN 1x1 Conv2d -> BatchNorm2d -> ReLU]                                  (pytorch_utils.py:5-32)
//     F.max_pool2d(kernel=[1, nsample]) ; squeeze ; (torch.cat over scales)          (pointnet2_modules.py:42-55)
// i.e. 9+ kernels that each stream the (B, C, npoint, nsample) activation through HBM, with ONE persistent,
// warp-specialised kernel (one CTA per SM) that never materialises the grouped tensor.
//
// A 128-row tile (= 128/nsample whole neighbourhoods) goes through three chained GEMMs with a hand-off to CUDA cores
// between them.  Each hand-off is a ~300-cycle round trip (tcgen05.commit -> mbarrier -> tcgen05.ld ... -> mbarrier), far
// longer than the MMAs of these small layers, so ONE tile in flight leaves the tensor pipe idle > 80 % of the time
// (profiles/r01: 4-21 % active).  Here up to FOUR tiles are in flight per SM, each in its own SLOT:
//
//   slot s (s < nslot)  = TMEM columns [s*cstride, (s+1)*cstride) + one H buffer in shared memory + 4 epilogue warps
//                         (warp 4s+q owns TMEM lanes 32q..32q+31) + barriers d_full[s] / epi_done[s].
//   MMA issuer warp(s)    one lane each; issuer i serves the slots s = i (mod ni) in a static round-robin over
//                         (layer, slot): L1(s0) L1(s1) .. | L2(s0) L2(s1) .. | L3(s0) ..  so that while the epilogue warps of
//                         slot s turn D_k into the next A operand, the tensor pipe runs the other slots' layers.
//   producer warps (8)    two groups of 4 warps on alternate tiles, one thread per tile row: idx -> point id, relative xyz
//                         (fp32 subtract, hi/lo fp16 split so the geometry enters layer 1 at ~fp32 precision), and the
//                         neighbour's fp16 feature row copied by cp.async (LDGSTS) straight into a ring of 16-channel
//                         K-slices already in the UMMA canonical K-major layout; full/empty mbarriers per ring slot.
//
//   layer 1   D1[128 x c1] = A . W1^T, one tcgen05.mma (kind::f16, fp32 accumulate in TMEM) per K-slice.  The bias rides
//             in the GEMM: two spare K positions of the xyz slice hold 1.0 and meet (b1_hi, b1_lo) in W1.
//   layer 2   D2[128 x c2] = H1 . W2^T + ones . [b2_hi b2_lo 0..]^T   (one extra K=16 MMA against a constant operand), so
//             both epilogues are just tcgen05.ld -> cvt.rn.relu.satfinite.f16x2 -> st.shared (no bias loads, no FADD).
//   layer 3   TRANSPOSED: D3[c3 x 128] = W3 . H2^T -- channels on TMEM lanes, positions on columns: the max over a
//             neighbourhood is a max over nsample consecutive columns inside ONE thread; bias + ReLU after the max
//             (they commute with it), written channel-major fp32 at the scale's channel offset (fuses torch.cat) through a
//             small staging area, and point-major fp16 for the next level's gather.
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
constexpr int XYZ_SLOTS = 11;                // [hi(3) | lo(3) | hi(3) | 1 | 1] against weights [wh | wh | wl | b_hi | b_lo]
constexpr int SLICE_BYTES = TILE_M * 16 * 2;  // one K-slice: 128 rows x 16 channels fp16 = 4 KB
constexpr int MAX_RING = 32;
constexpr int MAX_SLOTS = 4;
constexpr int PROD_WARPS = 8;                 // two groups of 4
// barrier block: [0] weights, [1] tmem address slot, [2 .. 2+MAX_SLOTS) d_full[s], [.. +MAX_SLOTS) epi_done[s],
//                then full[MAX_RING], empty[MAX_RING]
//                then h1_full[MAX_SLOTS][2], h_free[MAX_SLOTS][2] (xyz-only levels: layer 1 runs in the producers)
constexpr int BAR_WORDS = 2 + 2 * MAX_SLOTS + 2 * MAX_RING + 4 * MAX_SLOTS;

struct SaMlpLayout {
    int k0, c1, c2, c3, c3p, nb3, nslices, ring, nslot, ni, threads;
    uint32_t h_bytes, cstride;                              // per-slot H buffer bytes, per-slot TMEM column stride
    uint32_t off_w1, off_w2, off_w3, off_b3, off_ones, off_w1f, blob_bytes;   // inside the parameter blob == smem image
    uint32_t off_stage, stage_bytes;                        // per-slot staging area of the channel-major output
    int feat, nhbuf;                                        // features present; H buffers per slot (2 when layer 1 runs in the producers)
    int c3r, rep;                                           // layer 3 with c3 <= 64: W3 replicated rep = 128/c3r times down the 128 TMEM lanes
    uint32_t off_h, off_ring, off_bar, total_smem;
    uint32_t tmem_cols;
};

static inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

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
    L->off_w2 = o; o += (uint32_t)(L->c1 + 16) * L->c2 * 2;        // + one K-slice carrying the layer-2 bias
    L->off_w3 = o; o += (uint32_t)L->c2 * L->c3p * 2;
    L->off_b3 = o; o += (uint32_t)L->c3p * 4;
    L->off_ones = o; o += SLICE_BYTES;                            // constant A operand of the bias MMA
    L->off_w1f = o; o += d->c_in > 0 ? 0u : (uint32_t)L->c1 * 16;  // fp32 (wx, wy, wz, b) per channel: layer 1 of xyz-only levels on CUDA cores
    L->feat = d->c_in > 0;
    L->nhbuf = L->feat ? 1 : 2;
    // Layer 3 is transposed (channels on TMEM lanes).  With c3 <= 64 half or three quarters of the 128 lanes would hold zero
    // padding and their epilogue warps would idle while the others walk all 128 position columns: instead W3 is REPLICATED
    // down the lanes (same MMA, no extra cost) and copy cp serves the position columns [cp * 128/rep, (cp+1) * 128/rep) --
    // every warp reads 128/rep columns.  A neighbourhood (nsample columns) must not straddle two copies.
    L->c3r = 128; L->rep = 1;
    if (L->c3p == 128) {
        const int r = d->c3 <= 32 ? 32 : (d->c3 <= 64 ? 64 : 128);
        if (128 / r > 1 && d->nsample <= 128 / (128 / r)) { L->c3r = r; L->rep = 128 / r; }
    }
    L->blob_bytes = o;                                            // multiple of 16 by construction
    const int hk = L->c1 > L->c2 ? L->c1 : L->c2;
    L->h_bytes = (uint32_t)TILE_M * hk * 2;
    L->cstride = (uint32_t)(128 * L->nb3) > (uint32_t)hk ? 128 * L->nb3 : hk;
    L->off_h = round_up(o, 128);
    const uint32_t budget = 227u * 1024u - 1024u - 256u;          // 1 KB per CTA is reserved by the driver
    const uint32_t bar_bytes = 8u * BAR_WORDS + 16u;
    int nslot = 512 / (int)L->cstride;
    if (nslot > MAX_SLOTS) nslot = MAX_SLOTS;
    { const int v = env_int("G4D_SA_NSLOT", 0); if (v >= 1 && v < nslot) nslot = v; }      // tuning knob
    // as many slots as shared memory allows while the ring still holds two tiles' worth of slices (the producers must
    // run a tile ahead of the issuer); a single slot only needs a ring of two slices
    bool ok = false;
    // channel-major output staging: inside the H buffer when features are present (H is free during epilogue 3); a separate
    // area for xyz-only levels, whose H buffers are refilled by the producers meanwhile
    const int G = TILE_M / d->nsample;
    L->stage_bytes = L->feat ? 0u : round_up((uint32_t)L->c3 * (uint32_t)(G + 1) * 4u, 128u);
    for (; nslot >= 1; nslot >>= 1) {                             // 4, 2, 1: two issuers take the tiles of even / odd q
        const uint32_t fixed = L->off_h + (uint32_t)nslot * ((uint32_t)L->nhbuf * L->h_bytes + L->stage_bytes) + bar_bytes;
        if (fixed + 4u * SLICE_BYTES > budget) continue;
        int ring = (int)((budget - fixed) / SLICE_BYTES);
        if (ring > MAX_RING) ring = MAX_RING;
        if (nslot >= 2) ring &= ~1;                               // two issuers: two sub-rings of ring / 2 slices
        const int want = 2 * L->nslices < MAX_RING ? 2 * L->nslices : MAX_RING;
        if (nslot > 1 && ring < want) continue;
        L->nslot = nslot; L->ring = ring; ok = true;
        break;
    }
    if (!ok) { *why = msgs[5]; return false; }
    L->ni = L->nslot >= 2 ? 2 : 1;
    { const int v = env_int("G4D_SA_NI", 0); if (v >= 1 && v <= 2 && v <= L->nslot) L->ni = v; }
    L->threads = (4 * L->nslot + L->ni + PROD_WARPS) * 32;
    L->off_stage = L->off_h + (uint32_t)L->nslot * (uint32_t)L->nhbuf * L->h_bytes;
    L->off_ring = L->off_stage + (uint32_t)L->nslot * L->stage_bytes;
    L->off_bar = L->off_ring + (uint32_t)L->ring * SLICE_BYTES;
    L->total_smem = L->off_bar + bar_bytes;
    uint32_t p2 = 32;
    while (p2 < (uint32_t)L->nslot * L->cstride) p2 <<= 1;
    L->tmem_cols = p2;
    if (L->total_smem > 227 * 1024) { *why = msgs[5]; return false; }
    return true;
}

// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// wait for the loads above; the register pins keep every use of r[] below the wait
template <int NREG>
__device__ __forceinline__ void tmem_ld_wait(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NREG; ++i) asm volatile("" : "+r"(r[i]));
}
// two fp32 accumulator words (bit patterns) -> packed fp16x2 with ReLU and saturation (no inf can enter the next layer)
__device__ __forceinline__ uint32_t pack_relu_bits(uint32_t lo, uint32_t hi) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
    return d;
}

// ---------------------------------------------------------------------------------------------------------

struct SaMlpArgs {
    SaMlpLayout L;
    int c_in, nsample, n, m;
    int total_rows;                  // b * m * nsample (< 2^31, checked on the host)
    int ntiles;
    const float* xyz;                // (b, n, 3)
    const float* new_xyz;            // (b, m, 3)
    const int* idx;                  // (b, m, nsample)
    const __half* feat_pm;           // (b, n, c_in) or null
    const unsigned char* params;     // packed blob (device)
    float* out_cm;                   // (b, ctot, m)
    __half* out_pm;                  // (b, m, ctot) or null
    int ctot, coff;
    int flags;                       // bit 0: epilogue warps spin on d_full instead of the suspending wait (tuning knob G4D_SA_SPIN)
    long long* dbg;                  // optional clock64() timeline of CTA 0 (g4d_debug_timeline); null = off
};

// Unstaged channel-major store of one output value (one centroid per tile: nsample == 128).  Out of line: code size.
__device__ __noinline__ void store_cm_direct(float* out_cm, unsigned cloud, unsigned p, unsigned m, int ctot, int ch, float o) {
    while (p >= m) { p -= m; ++cloud; }
    out_cm[((size_t)cloud * ctot + ch) * m + p] = o;
}

// named barrier over the 128 threads (4 warps) of one slot; immediate ids (a register id makes ptxas reserve all 16)
__device__ __forceinline__ void slot_bar_sync(int slot) {
    switch (slot) {
        case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
    }
}

// One instantiation per (nsample, features present, big CTA).  BIG: up to 26 warps (4 slots); otherwise 13 warps (1 slot).
template <int NS, bool FEAT, bool BIG>
__global__ void __launch_bounds__(BIG ? 832 : 416, 1)
sa_mlp_max_kernel(const SaMlpArgs a) {
    constexpr int LG_NS = NS == 8 ? 3 : NS == 16 ? 4 : NS == 32 ? 5 : NS == 64 ? 6 : 7;
    extern __shared__ __align__(128) unsigned char smem[];
    const SaMlpLayout& L = a.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nslot = L.nslot, ni = L.ni;
    const int epi_warps = 4 * nslot;
    const float* b3 = reinterpret_cast<const float*>(smem + L.off_b3);
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    const uint32_t bar_w = bar0, tmem_slot = bar0 + 8;
    const uint32_t bar_dfull = bar0 + 16, bar_epi = bar_dfull + 8 * MAX_SLOTS;
    const uint32_t bar_full = bar_epi + 8 * MAX_SLOTS, bar_empty = bar_full + 8 * MAX_RING;
    const uint32_t bar_h1full = bar_empty + 8 * MAX_RING, bar_hfree = bar_h1full + 16 * MAX_SLOTS;      // [slot][buf]
    const uint32_t s_w1 = smem_u32(smem + L.off_w1), s_w2 = smem_u32(smem + L.off_w2), s_w3 = smem_u32(smem + L.off_w3);
    const uint32_t s_ones = smem_u32(smem + L.off_ones);
    const uint32_t s_h = smem_u32(smem + L.off_h), s_ring = smem_u32(smem + L.off_ring);

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int s = 0; s < nslot; ++s) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_epi + 8 * s, 4); }
        for (int s = 0; s < 2 * nslot; ++s) { mbar_init(bar_h1full + 8 * s, 4); mbar_init(bar_hfree + 8 * s, 1); }
        for (int r = 0; r < L.ring; ++r) { mbar_init(bar_full + 8 * r, 4);     /* the 4 warps of the producer group that owns the slot's tile */ mbar_init(bar_empty + 8 * r, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, L.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L.off_bar + 8);
    if (tid == 0) {
        mbar_expect_tx(bar_w, L.blob_bytes);
        bulk_g2s(smem_u32(smem), a.params, L.blob_bytes, bar_w);      // folded weights, biases, the constant ones slice: once per CTA
    }

    // With two issuers the ring is split into TWO SUB-RINGS, one per producer group (tiles of even / odd q).  A ring slot's barriers carry one
    // phase bit, which is only safe while each side of a slot sees its phases in order: ONE group fills a sub-ring in program
    // order and ONE issuer drains it in tile order (issuer q % 2 with two issuers -- nslot is even then --, the single issuer
    // alternating between the sub-rings otherwise).  With a shared ring a group could run a lap ahead of the other one's
    // unfilled slices, or an issuer see another issuer's lap: a wait then passes on the wrong phase (it hung with 3 slots).
    // A single issuer (one slot) is fed by ONE producer group over the whole ring.
    const int nring = L.ni == 2 ? 2 : 1;
    const int S = L.nslices, RING = L.ring / nring;           // slices per sub-ring
    const int bx = (int)blockIdx.x, gx = (int)gridDim.x;
    // this CTA's tile sequence: tile(q) = bx + q * gx, q = 0 .. nq-1; slot of tile q = q % nslot; its slices are the
    // slice numbers k*S .. k*S+S-1 in the sub-ring of its issuer, k = index of the tile among that issuer's tiles
    const int nq = bx < a.ntiles ? (a.ntiles - bx + gx - 1) / gx : 0;

    if (warp >= epi_warps + ni) {
        // =========================== PRODUCERS: one thread per tile row ===================================
        // Two groups of 4 warps take alternate tiles of the sequence and fill the ring concurrently, in order.  Inside a
        // group the loads are software-pipelined: idx two tiles ahead, coordinates one tile ahead, the feature row of the
        // current tile entirely in flight (cp.async) before the first slice is handed over.
        const int NG = FEAT ? nring : PROD_WARPS / 4;          // groups at work (xyz-only levels have no ring: both groups always)
        const int pw = warp - (epi_warps + ni);
        const int grp = pw >> 2;                                 // 0 .. PROD_WARPS/4-1
        const int r = (pw & 3) * 32 + lane;                      // tile row 0..127
        const int nchunk_feat = FEAT ? (a.c_in >> 3) : 0;
        const unsigned um = (unsigned)a.m;
        if (!FEAT) mbar_wait(bar_w, 0);                          // layer-1 weights live in the blob
        int q = grp < NG ? grp : nq;                           // a group without work falls straight through
        int t_cur = q < nq ? bx + q * gx : a.ntiles, t_nxt = q + NG < nq ? bx + (q + NG) * gx : a.ntiles;
        int src_n = 0, src_nn = 0;
        float npx = 0.f, npy = 0.f, npz = 0.f, nqx = 0.f, nqy = 0.f, nqz = 0.f;
        unsigned npt = 0;
        if (t_cur < a.ntiles && t_cur * TILE_M + r < a.total_rows) src_n = __ldg(a.idx + t_cur * TILE_M + r);
        if (t_nxt < a.ntiles && t_nxt * TILE_M + r < a.total_rows) src_nn = __ldg(a.idx + t_nxt * TILE_M + r);
        if (t_cur < a.ntiles && t_cur * TILE_M + r < a.total_rows) {
            const unsigned gp = (unsigned)(t_cur * TILE_M + r) >> LG_NS;
            npt = (gp / um) * (unsigned)a.n + (unsigned)src_n;
            const float* p = a.xyz + (size_t)npt * 3;
            const float* c = a.new_xyz + (size_t)gp * 3;
            npx = __ldg(p); npy = __ldg(p + 1); npz = __ldg(p + 2);
            nqx = __ldg(c); nqy = __ldg(c + 1); nqz = __ldg(c + 2);
        }
        while (t_cur < a.ntiles) {
            const int tile = t_cur;
            const bool live = tile * TILE_M + r < a.total_rows;
            const float dx = npx - nqx, dy = npy - nqy, dz = npz - nqz;
            const unsigned pt = npt;
            // sub-ring of this group (q % 2 == grp), and the tile's slice number in it
            const uint32_t rbase = (uint32_t)(grp * RING);
            const uint32_t it0 = (uint32_t)(nring == 2 ? q >> 1 : q) * (uint32_t)S;
            // ---- advance the pipeline: issue the loads of the following tiles before touching this one
            q += NG;
            t_cur = t_nxt;
            t_nxt = q + NG < nq ? bx + (q + NG) * gx : a.ntiles;
            src_n = src_nn;
            src_nn = 0;
            if (t_nxt < a.ntiles && t_nxt * TILE_M + r < a.total_rows) src_nn = __ldg(a.idx + t_nxt * TILE_M + r);
            if (t_cur < a.ntiles && t_cur * TILE_M + r < a.total_rows) {
                const unsigned gp = (unsigned)(t_cur * TILE_M + r) >> LG_NS;
                npt = (gp / um) * (unsigned)a.n + (unsigned)src_n;
                const float* p = a.xyz + (size_t)npt * 3;
                const float* c = a.new_xyz + (size_t)gp * 3;
                npx = __ldg(p); npy = __ldg(p + 1); npz = __ldg(p + 2);
                nqx = __ldg(c); nqy = __ldg(c + 1); nqz = __ldg(c + 2);
            }
            // ---- this tile: relative xyz, hi/lo split.  K positions: hi.x hi.y hi.z lo.x lo.y lo.z hi.x hi.y | hi.z 1 1 0 0 0 0 0
            // (the two 1.0 meet b1_hi, b1_lo in W1: the layer-1 bias is part of the GEMM).  Rows past the end stay all-zero.
            uint4 xc0 = make_uint4(0, 0, 0, 0), xc1 = make_uint4(0, 0, 0, 0);
            if (live) {
                const __half hx = __float2half_rn(dx), hy = __float2half_rn(dy), hz = __float2half_rn(dz);
                const uint32_t uhx = __half_as_ushort(hx), uhy = __half_as_ushort(hy), uhz = __half_as_ushort(hz);
                const uint32_t ulx = __half_as_ushort(__float2half_rn(dx - __half2float(hx))),
                               uly = __half_as_ushort(__float2half_rn(dy - __half2float(hy))),
                               ulz = __half_as_ushort(__float2half_rn(dz - __half2float(hz)));
                xc0 = make_uint4(uhx | (uhy << 16), uhz | (ulx << 16), uly | (ulz << 16), uhx | (uhy << 16));
                xc1 = make_uint4(uhz | (0x3C00u << 16), 0x3C00u, 0, 0);
            }
            if (!FEAT) {
                // xyz-only level: LAYER 1 RUNS HERE, on the CUDA cores (K = 3: 0.1 % of the FLOPs; on the tensor pipe it cost a
                // whole MMA -> TMEM -> epilogue hand-off, ~1500 cycles of the tile's serial chain).  Each thread computes its row of
                // H1 = relu(W1 . (dx, dy, dz) + b1) in fp32 and writes it, fp16, straight into the slot's H buffer (double-buffered:
                // the tile after next of this slot is prepared while the current one is in layers 2-3).
                const int q_this = q - NG;                         // q was advanced above
                const int sl = q_this % nslot, t = q_this / nslot; // slot and index of the tile within its slot
                const int buf = t & 1;
                const uint32_t hf = bar_hfree + 8 * (2 * sl + buf), h1f = bar_h1full + 8 * (2 * sl + buf);
                mbar_wait(hf, ((t >> 1) + 1) & 1);         // layer 3 of tile t-2 has read this buffer (passes at once for t < 2)
                uint4* hd = reinterpret_cast<uint4*>(smem + L.off_h + (size_t)(2 * sl + buf) * L.h_bytes);
                const float4* w1f = reinterpret_cast<const float4*>(smem + L.off_w1f);
#pragma unroll 1
                for (int j = 0; j < (L.c1 >> 3); ++j) {
                    uint32_t h[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 wa = w1f[8 * j + 2 * i], wb = w1f[8 * j + 2 * i + 1];          // broadcast reads
                        const float va = fmaf(wa.z, dz, fmaf(wa.y, dy, fmaf(wa.x, dx, wa.w)));
                        const float vb = fmaf(wb.z, dz, fmaf(wb.y, dy, fmaf(wb.x, dx, wb.w)));
                        h[i] = pack_relu_f16x2(va, vb);
                    }
                    hd[(size_t)j * TILE_M + r] = live ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(0, 0, 0, 0);
                }
                fence_proxy_async();                              // generic-proxy stores -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(h1f);
            } else {
                // Feature levels: every 16-byte chunk of the neighbour's fp16 row goes global -> shared with cp.async (LDGSTS), one
                // commit group per K-slice, a whole wave of slices in flight at once (no register staging: 128 threads x 2S copies
                // = the whole 128 x k0 tile outstanding, which is what hides the L2/HBM latency of the gather); then the slices are
                // handed over in order as their groups complete.  A wave is at most min(S, RING) slices: it must fit the ring.
                const char* srcrow = reinterpret_cast<const char*>(a.feat_pm + (size_t)pt * a.c_in);
                const int wave = S < RING ? S : RING;
                uint32_t slot = it0 % RING, ph = (it0 / RING) & 1;             // one division per tile; then walked
#pragma unroll 1
                for (int w0 = 0; w0 < S; w0 += wave) {
                const int w1 = (w0 + wave < S) ? w0 + wave : S;
                const uint32_t slot_w = slot;
#pragma unroll 1
                for (int sl = w0; sl < w1; ++sl) {
                    mbar_wait(bar_empty + 8 * (rbase + slot), ph ^ 1);           // slot free (first lap passes at once); suspending wait, not a nanosleep poll (>= 256 ns per miss)
                    const uint32_t sdst = s_ring + (rbase + slot) * SLICE_BYTES + (uint32_t)r * 16;
                    uint4* gdst = reinterpret_cast<uint4*>(smem + L.off_ring + (size_t)(rbase + slot) * SLICE_BYTES);
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
                    if (lane == 0) mbar_arrive(bar_full + 8 * (rbase + slot_a));
                    if (++slot_a == (uint32_t)RING) slot_a = 0;
                }
                }
            }
        }
    } else if (warp >= epi_warps) {
        // =========================== MMA ISSUERS ============================================================
        // Issuer i serves the slots s = i, i + ni, ...  Static schedule per round of nslot tiles:
        //   layer 1 of its slots, layer 2 of its slots, layer 3 of its slots; a layer of slot s waits for the epilogue of the
        //   previous stage of THAT slot only, which had the other slots' MMAs to finish under.
        // The WHOLE warp runs this code converged, on warp-uniform values only (loop counters, barrier addresses, descriptors),
        // and one elected lane issues: tcgen05.mma / tcgen05.commit take their operands from uniform registers, and with
        // per-thread operands inside an `if (lane == 0)` ptxas wraps EVERY one of them in an ELECT + 5 x R2UR.BROADCAST +
        // BRA.U.ANY loop (measured: 600-900 cycles to issue a layer of 1-3 MMAs; profiles/r02_sa_timeline.txt).
        const int issuer = __shfl_sync(0xFFFFFFFFu, warp - epi_warps, 0);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem, 0);
        mbar_wait(bar_w, 0);
        const uint32_t ring_lo = desc_lo(s_ring, TILE_M * 16), w1_lo = desc_lo(s_w1, L.c1 * 16), w1_step = (uint32_t)(2 * L.c1 * 16) >> 4;
        const uint32_t w2_lo = desc_lo(s_w2, L.c2 * 16), w2_step = (uint32_t)(2 * L.c2 * 16) >> 4;
        const uint32_t w3_step = (uint32_t)(2 * L.c3p * 16) >> 4, h_step = (uint32_t)(2 * TILE_M * 16) >> 4;
        const uint32_t ones_lo = desc_lo(s_ones, TILE_M * 16);
        const uint32_t idesc1 = umma_idesc(TILE_M, L.c1), idesc2 = umma_idesc(TILE_M, L.c2), idesc3 = umma_idesc(128, TILE_M);
        long long* dbg = (a.dbg && blockIdx.x == 0 && issuer == 0 && lane == 0) ? a.dbg : nullptr;    // timeline (debug): [0..5] = after-wait / after-commit stamps of layers 1,2,3 of slot 0
        const int wave = S < RING ? S : RING;
        const int nk1 = L.c1 / 16, nk2 = L.c2 / 16;
        // epilogue hand-offs of a slot complete in the order (tile 0: TMEM free*, e1, e2 | tile 1: e3 of tile 0, e1, e2 | ..):
        // the wait before layer l of round r is the (3r + l)-th of its slot, whatever the slot (* = passes at once)
        // (xyz-only levels have no layer 1 here: two hand-offs per tile)
        uint32_t nepi = 0;
        for (int q0 = 0, rnd = 0; q0 < nq; q0 += nslot, nepi += (FEAT ? 3 : 2), ++rnd) {
            // ---- layer 1: needs the slot's TMEM drained by the previous tile's epilogue 3, and the tile's K-slices
#pragma unroll 1
            for (int s = issuer; FEAT && s < nslot && q0 + s < nq; s += ni) {
                mbar_wait_spin(bar_epi + 8 * s, (nepi + 1) & 1);
                tc_fence_after();
                if (dbg && s == 0 && q0 < 16 * nslot) dbg[(q0 / nslot) * 16 + 0] = clock64();
                const uint32_t rbase = nring == 2 ? (uint32_t)(((q0 + s) & 1) * RING) : 0u;
                const uint32_t it0 = (uint32_t)(nring == 2 ? (q0 + s) >> 1 : q0 + s) * (uint32_t)S;      // the tile's slice number in its sub-ring
                uint32_t slot = it0 % RING, ph = (it0 / RING) & 1;
                uint32_t blo = w1_lo;
                const uint32_t dcol = tmem_u + s * L.cstride;
#pragma unroll 1
                for (int w0 = 0; w0 < S; w0 += wave) {
                    const int w1 = (w0 + wave < S) ? w0 + wave : S;
                    // every producer warp arrives on a tile's slices in order, so the LAST slice of a wave being complete
                    // means all of them are: one wait per wave instead of one per slice (each try_wait costs ~90 cycles)
                    {
                        uint32_t ls = slot + (uint32_t)(w1 - w0 - 1), lph = ph;
                        if (ls >= (uint32_t)RING) { ls -= RING; lph ^= 1; }
                        mbar_wait(bar_full + 8 * (rbase + ls), lph);   // gathers in flight: leave the issue slots to the producers
                        tc_fence_after();
                    }
                    if (elect_one_sync()) {
                        uint32_t e_slot = slot, e_blo = blo;      // private walk of the elected lane: the warp's copies stay uniform
#pragma unroll 1
                        for (int sl = w0; sl < w1; ++sl) {
                            umma_f16(dcol, desc64(ring_lo + (rbase + e_slot) * (SLICE_BYTES >> 4)), desc64(e_blo), idesc1, sl > 0);
                            umma_commit(bar_empty + 8 * (rbase + e_slot));      // slot reusable once this (and earlier) MMAs retire
                            e_blo += w1_step;
                            if (++e_slot == (uint32_t)RING) e_slot = 0;
                        }
                        if (w1 == S) umma_commit(bar_dfull + 8 * s);
                    }
                    __syncwarp();
                    slot += (uint32_t)(w1 - w0); if (slot >= (uint32_t)RING) { slot -= RING; ph ^= 1; }
                    blo += (uint32_t)(w1 - w0) * w1_step;
                }
                if (dbg && s == 0 && q0 < 16 * nslot) dbg[(q0 / nslot) * 16 + 1] = clock64();
            }
            // ---- layer 2: needs H1 written by epilogue 1; the bias enters through the constant ones operand
#pragma unroll 1
            for (int s = issuer; s < nslot && q0 + s < nq; s += ni) {
                mbar_wait_spin(bar_epi + 8 * s, (nepi + (FEAT ? 2 : 1)) & 1);     // xyz-only: the slot's TMEM drained by the previous tile
                const int hb = FEAT ? s : 2 * s + (rnd & 1);                     // H buffer of this tile
                if (!FEAT) mbar_wait_spin(bar_h1full + 8 * hb, (rnd >> 1) & 1);  // H1 written by the producers
                tc_fence_after();
                if (dbg && s == 0 && q0 < 16 * nslot) dbg[(q0 / nslot) * 16 + 2] = clock64();
                if (elect_one_sync()) {
                    uint32_t alo = desc_lo(s_h + hb * L.h_bytes, TILE_M * 16), blo = w2_lo;
                    const uint32_t dcol = tmem_u + s * L.cstride;
#pragma unroll 1
                    for (int k = 0; k < nk1; ++k) {
                        umma_f16(dcol, desc64(alo), desc64(blo), idesc2, k > 0);
                        alo += h_step; blo += w2_step;
                    }
                    umma_f16(dcol, desc64(ones_lo), desc64(blo), idesc2, 1);
                    umma_commit(bar_dfull + 8 * s);
                }
                __syncwarp();
                if (dbg && s == 0 && q0 < 16 * nslot) dbg[(q0 / nslot) * 16 + 3] = clock64();
            }
            // ---- layer 3, transposed: D3[c3p x 128] = W3 . H2^T
#pragma unroll 1
            for (int s = issuer; s < nslot && q0 + s < nq; s += ni) {
                mbar_wait_spin(bar_epi + 8 * s, (nepi + (FEAT ? 3 : 2)) & 1);
                tc_fence_after();
                if (dbg && s == 0 && q0 < 16 * nslot) dbg[(q0 / nslot) * 16 + 4] = clock64();
                const int hb = FEAT ? s : 2 * s + (rnd & 1);
                if (elect_one_sync()) {
                    const uint32_t dcol = tmem_u + s * L.cstride;
#pragma unroll 1
                    for (int j = 0; j < L.nb3; ++j) {
                        uint32_t alo = desc_lo(s_w3 + (uint32_t)j * 128 * 16, L.c3p * 16), blo = desc_lo(s_h + hb * L.h_bytes, TILE_M * 16);
#pragma unroll 1
                        for (int k = 0; k < nk2; ++k) {
                            umma_f16(dcol + j * 128, desc64(alo), desc64(blo), idesc3, k > 0);
                            alo += w3_step; blo += h_step;
                        }
                    }
                    umma_commit(bar_dfull + 8 * s);
                    if (!FEAT) umma_commit(bar_hfree + 8 * hb);   // the producers may refill this H buffer once layer 3 has read it
                }
                __syncwarp();
                if (dbg && s == 0 && q0 < 16 * nslot) dbg[(q0 / nslot) * 16 + 5] = clock64();
            }
        }
    } else {
        // =========================== EPILOGUE: warp 4s+q owns TMEM lanes 32q .. 32q+31 of slot s ===============
        mbar_wait(bar_w, 0);                                      // b3 lives in the weight blob
        const int slot = warp >> 2, quad = warp & 3;
        const int row = quad * 32 + lane;                         // tile row (layers 1-2) / channel within block (layer 3)
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + slot * L.cstride;
        unsigned char* const hbuf0 = smem + L.off_h + (size_t)slot * L.nhbuf * L.h_bytes;
        const uint32_t dfull = bar_dfull + 8 * slot, epi = bar_epi + 8 * slot;
        constexpr int G = TILE_M / NS;                            // centroids per tile
        uint32_t nd = 0;                                          // d_full hand-offs waited for
        // timeline (debug): slot 0 of CTA 0, first 16 tiles: [8..13] = wake/arrive stamps of epilogues 1,2,3 (warp 0, lane 0)
        long long* dbg = (a.dbg && blockIdx.x == 0 && warp == 0 && lane == 0) ? a.dbg : nullptr;
        // Channel-major fp32 output: with lane = channel a direct store would touch 32 different sectors per instruction
        // (4 useful bytes each).  The tile's (channel x centroid) block is staged in the slot's H buffer (free during
        // epilogue 3) and written out with lanes along the centroid index.
        const bool staged = G >= 2 && (size_t)L.c3 * (G + 1) * 4 <= (size_t)(FEAT ? L.h_bytes : L.stage_bytes);
        float* stage = reinterpret_cast<float*>(FEAT ? hbuf0 : smem + L.off_stage + (size_t)slot * L.stage_bytes);
#pragma unroll 1
        for (int q = slot; q < nq; q += nslot) {
            const int tile = bx + q * gx;
            unsigned char* hbuf = FEAT ? hbuf0 : hbuf0 + (size_t)((q / nslot) & 1) * L.h_bytes;      // xyz-only: double-buffered
            // ---- epilogues 1 and 2 (xyz-only levels: only 2 -- layer 1 ran in the producers): D -> ReLU -> fp16 -> H (bias already inside D; H1 is dead when d_full fires for layer 2)
#pragma unroll 1
            for (int layer = FEAT ? 0 : 1; layer < 2; ++layer) {
                if (a.flags & 1) mbar_wait_spin(dfull, nd & 1); else mbar_wait(dfull, nd & 1);
                ++nd;
                tc_fence_after();
                if (dbg && q < 16 * nslot) dbg[(q / nslot) * 16 + 8 + 2 * layer] = clock64();
                const int ncols = layer ? L.c2 : L.c1;
                uint4* hd = reinterpret_cast<uint4*>(hbuf);
                int c = 0;
                if constexpr (!BIG) {
                    // one-slot layout (13 warps, up to 157 registers): two TMEM loads in flight per wait -- the load -> wait ->
                    // convert -> store chain of this warp is the tile's critical path (nothing else runs meanwhile)
#pragma unroll 1
                    for (; c + 64 <= ncols; c += 64) {
                        uint32_t v[64];
                        tmem_ld32_raw(taddr + c, v);
                        tmem_ld32_raw(taddr + c + 32, v + 32);
                        tmem_ld_wait<64>(v);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            hd[(size_t)((c >> 3) + u) * TILE_M + row] =
                                make_uint4(pack_relu_bits(v[8 * u], v[8 * u + 1]), pack_relu_bits(v[8 * u + 2], v[8 * u + 3]),
                                           pack_relu_bits(v[8 * u + 4], v[8 * u + 5]), pack_relu_bits(v[8 * u + 6], v[8 * u + 7]));
                    }
                }
#pragma unroll 1
                for (; c + 32 <= ncols; c += 32) {
                    uint32_t v[32];
                    tmem_ld32_raw(taddr + c, v);
                    tmem_ld_wait<32>(v);
                    if (dbg && layer == 0 && c == 0 && q < 16 * nslot) dbg[(q / nslot) * 16 + 6] = clock64();
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        hd[(size_t)((c >> 3) + u) * TILE_M + row] =
                            make_uint4(pack_relu_bits(v[8 * u], v[8 * u + 1]), pack_relu_bits(v[8 * u + 2], v[8 * u + 3]),
                                       pack_relu_bits(v[8 * u + 4], v[8 * u + 5]), pack_relu_bits(v[8 * u + 6], v[8 * u + 7]));
                }
                if (c < ncols) {                                  // ncols % 32 == 16
                    uint32_t v[16];
                    tmem_ld16_raw(taddr + c, v);
                    tmem_ld_wait<16>(v);
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        hd[(size_t)((c >> 3) + u) * TILE_M + row] =
                            make_uint4(pack_relu_bits(v[8 * u], v[8 * u + 1]), pack_relu_bits(v[8 * u + 2], v[8 * u + 3]),
                                       pack_relu_bits(v[8 * u + 4], v[8 * u + 5]), pack_relu_bits(v[8 * u + 6], v[8 * u + 7]));
                }
                if (dbg && layer == 0 && q < 16 * nslot) dbg[(q / nslot) * 16 + 7] = clock64();
                tc_fence_before();
                fence_proxy_async();
                if (dbg && layer == 0 && q < 16 * nslot) dbg[(q / nslot) * 16 + 14] = clock64();
                __syncwarp();
                if (lane == 0) mbar_arrive(epi);
                if (dbg && q < 16 * nslot) dbg[(q / nslot) * 16 + 9 + 2 * layer] = clock64();
            }
            // ---- epilogue 3: lane = output channel; max over each neighbourhood's nsample consecutive columns
            if (a.flags & 1) mbar_wait_spin(dfull, nd & 1); else mbar_wait(dfull, nd & 1);
            ++nd;
            tc_fence_after();
            if (dbg && q < 16 * nslot) dbg[(q / nslot) * 16 + 12] = clock64();
            const unsigned gp0 = (unsigned)tile * (unsigned)G;    // first centroid of the tile
            const unsigned cloud0 = gp0 / (unsigned)a.m, p0 = gp0 - cloud0 * (unsigned)a.m;
            // rep > 1 (c3 <= 64): this warp's lanes hold copy cp of the channels and serve the columns [col0, col0 + ncol)
            const int cp = (quad * 32) / L.c3r, chq = (quad * 32) % L.c3r;
            const int ncol = TILE_M / L.rep, col0 = cp * ncol;
#pragma unroll 1
            for (int j = 0; j < L.nb3; ++j) {
                if (j * 128 + chq >= L.c3) continue;              // warp-uniform: these 32 lanes hold padding channels only
                const int ch = j * 128 + chq + lane;
                const float bias = ch < L.c3 ? b3[ch] : 0.f;
                float run = -INFINITY;
                // CH positions per step: 32, or 64 with two TMEM loads in flight in the one-slot layout (when the copy spans >= 64)
                constexpr int CHMAX = BIG ? 32 : 64;
                if (CHMAX == 64 && (ncol & 63) == 0) {
#pragma unroll 1
                    for (int cc = col0 >> 6; cc < ((col0 + ncol) >> 6); ++cc) {
                        uint32_t v[64];
                        tmem_ld32_raw(taddr + j * 128 + 64 * cc, v);
                        tmem_ld32_raw(taddr + j * 128 + 64 * cc + 32, v + 32);
                        tmem_ld_wait<64>(v);
                        constexpr int GPC = NS <= 64 ? 64 / NS : 1;
                        constexpr int W = NS <= 64 ? NS : 64;
#pragma unroll
                        for (int g = 0; g < GPC; ++g) {
                            float mx = __uint_as_float(v[W * g]);
#pragma unroll
                            for (int i = 1; i < W; ++i) mx = fmaxf(mx, __uint_as_float(v[W * g + i]));
                            int gi;
                            if constexpr (NS <= 64) { gi = cc * GPC + g; }
                            else {
                                run = fmaxf(run, mx);
                                if (((cc + 1) * 64) % NS != 0) continue;
                                mx = run; run = -INFINITY;
                                gi = (cc * 64) / NS;
                            }
                            const unsigned gp = gp0 + (unsigned)gi;
                            if (ch < L.c3 && (gp << LG_NS) < (unsigned)a.total_rows) {
                                const float o = fmaxf(mx + bias, 0.f);
                                if (staged) stage[ch * (G + 1) + gi] = o;
                                else store_cm_direct(a.out_cm, cloud0, p0 + (unsigned)gi, (unsigned)a.m, a.ctot, a.coff + ch, o);
                                if (a.out_pm) a.out_pm[(size_t)gp * a.ctot + a.coff + ch] = __float2half_rn(fminf(o, 65504.f));
                            }
                        }
                    }
                } else {
#pragma unroll 1
                for (int cc = col0 >> 5; cc < ((col0 + ncol) >> 5); ++cc) {      // 32 positions at a time
                    uint32_t v[32];
                    tmem_ld32_raw(taddr + j * 128 + 32 * cc, v);
                    tmem_ld_wait<32>(v);
                    constexpr int GPC = NS <= 32 ? 32 / NS : 1;   // whole neighbourhoods inside a 32-column chunk
                    constexpr int W = NS <= 32 ? NS : 32;
#pragma unroll
                    for (int g = 0; g < GPC; ++g) {
                        float mx = __uint_as_float(v[W * g]);
#pragma unroll
                        for (int i = 1; i < W; ++i) mx = fmaxf(mx, __uint_as_float(v[W * g + i]));
                        int gi;                                    // centroid index within the tile
                        if constexpr (NS <= 32) { gi = cc * GPC + g; }
                        else {
                            run = fmaxf(run, mx);
                            if (((cc + 1) * 32) % NS != 0) continue;
                            mx = run; run = -INFINITY;
                            gi = (cc * 32) / NS;
                        }
                        const unsigned gp = gp0 + (unsigned)gi;
                        if (ch < L.c3 && (gp << LG_NS) < (unsigned)a.total_rows) {
                            const float o = fmaxf(mx + bias, 0.f);
                            if (staged) stage[ch * (G + 1) + gi] = o;
                            else store_cm_direct(a.out_cm, cloud0, p0 + (unsigned)gi, (unsigned)a.m, a.ctot, a.coff + ch, o);
                            if (a.out_pm) a.out_pm[(size_t)gp * a.ctot + a.coff + ch] = __float2half_rn(fminf(o, 65504.f));
                        }
                    }
                }
                }
            }
            if (dbg && q < 16 * nslot) dbg[(q / nslot) * 16 + 15] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(epi);                      // TMEM drained: the slot's next tile may start layer 1
            if (dbg && q < 16 * nslot) dbg[(q / nslot) * 16 + 13] = clock64();
            if (staged) {
                slot_bar_sync(slot);                              // the 4 warps of this slot only
                const int et = quad * 32 + lane;                  // 0..127
                for (int i = et; i < L.c3 * G; i += 128) {
                    const int ch = i / G, g = i - ch * G;
                    if (((gp0 + (unsigned)g) << LG_NS) < (unsigned)a.total_rows) {
                        unsigned cloud = cloud0, pp = p0 + (unsigned)g;
                        while (pp >= (unsigned)a.m) { pp -= (unsigned)a.m; ++cloud; }
                        a.out_cm[((size_t)cloud * a.ctot + a.coff + ch) * a.m + pp] = stage[ch * (G + 1) + g];
                    }
                }
                slot_bar_sync(slot);                              // staging area is H again
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
// Debug aid: device buffer of >= 16*16 int64 that CTA 0 of the next g4d_sa_mlp_max launches fills with clock64() stamps of
// its slot 0 (per tile: issuer [0..5] = after-wait / after-commit of layers 1-3, epilogue warp 0 [8..13] = wake / arrive of
// epilogues 1-3); NULL switches it off.
G4D_API void g4d_debug_timeline(void* buf) { g_timeline = (long long*)buf; }

G4D_API int g4d_sa_mlp_k0(int c_in) { return (c_in + XYZ_SLOTS + 15) / 16 * 16; }

G4D_API size_t g4d_sa_mlp_param_bytes(const g4d_sa_mlp_desc* d) {
    SaMlpLayout L; const char* why = nullptr;
    if (!d || !make_layout(d, &L, &why)) { set_error("%s", why ? why : "sa_mlp: null descriptor"); return 0; }
    return L.blob_bytes;
}

// canonical K-major image of a (rows x K) fp16 operand
static bool put_canonical(__half* base, int R, int r, int k, float v) {
    base[((size_t)(k / 8) * R + r) * 8 + (k % 8)] = __float2half_rn(v);
    return fabsf(v) <= 65504.f;           // representable in fp16 (NaN fails too)
}

// Returns cudaErrorInvalidValue when a folded weight or bias does not fit fp16 (|v| > 65504): the caller must then take
// the operator route (the fused kernel would silently compute with inf).
G4D_API int g4d_sa_mlp_pack_params(const g4d_sa_mlp_desc* d, const float* w1, const float* b1, const float* w2, const float* b2,
                                   const float* w3, const float* b3, void* blob) {
    SaMlpLayout L; const char* why = nullptr;
    if (!d || !make_layout(d, &L, &why)) return bad_arg(why ? why : "sa_mlp: null descriptor");
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !blob) return bad_arg("sa_mlp_pack_params: null pointer");
    unsigned char* out = (unsigned char*)blob;
    memset(out, 0, L.blob_bytes);
    bool ok = true;
    const int cin = d->c_in, ld1 = 3 + cin;     // reference column order: [xyz(3) | features(c_in)]  (pointnet2_utils.py:258)
    __half* W1 = (__half*)(out + L.off_w1);
    for (int o = 0; o < L.c1; ++o) {
        for (int k = 0; k < cin; ++k) ok &= put_canonical(W1, L.c1, o, k, w1[(size_t)o * ld1 + 3 + k]);
        for (int j = 0; j < 3; ++j) {
            const float w = w1[(size_t)o * ld1 + j];
            const float wh = __half2float(__float2half_rn(w));
            ok &= put_canonical(W1, L.c1, o, cin + j, wh);          // x_hi * w_hi
            ok &= put_canonical(W1, L.c1, o, cin + 3 + j, wh);      // x_lo * w_hi
            ok &= put_canonical(W1, L.c1, o, cin + 6 + j, w - wh);  // x_hi * w_lo
        }
        const float bh = __half2float(__float2half_rn(b1[o]));
        ok &= put_canonical(W1, L.c1, o, cin + 9, bh);              // 1.0 * b_hi
        ok &= put_canonical(W1, L.c1, o, cin + 10, b1[o] - bh);     // 1.0 * b_lo
    }
    __half* W2 = (__half*)(out + L.off_w2);
    for (int o = 0; o < L.c2; ++o) {
        for (int k = 0; k < L.c1; ++k) ok &= put_canonical(W2, L.c2, o, k, w2[(size_t)o * L.c1 + k]);
        const float bh = __half2float(__float2half_rn(b2[o]));
        ok &= put_canonical(W2, L.c2, o, L.c1, bh);                 // the bias K-slice: meets the ones operand
        ok &= put_canonical(W2, L.c2, o, L.c1 + 1, b2[o] - bh);
    }
    __half* W3 = (__half*)(out + L.off_w3);
    for (int cp = 0; cp < L.rep; ++cp)                            // rep > 1: copies of the channels down the 128 lanes (see make_layout)
        for (int o = 0; o < L.c3; ++o)
            for (int k = 0; k < L.c2; ++k) ok &= put_canonical(W3, L.c3p, cp * L.c3r + o, k, w3[(size_t)o * L.c2 + k]);
    memcpy(out + L.off_b3, b3, sizeof(float) * L.c3);
    float* w1f = (float*)(out + L.off_w1f);                       // fp32 layer 1 for xyz-only levels (CUDA cores, in the producers)
    for (int o = 0; o < L.c1 && cin == 0; ++o) {
        for (int j = 0; j < 3; ++j) w1f[4 * o + j] = w1[(size_t)o * ld1 + j];
        w1f[4 * o + 3] = b1[o];
    }
    __half* ones = (__half*)(out + L.off_ones);
    for (int r = 0; r < TILE_M; ++r) { put_canonical(ones, TILE_M, r, 0, 1.f); put_canonical(ones, TILE_M, r, 1, 1.f); }
    if (!ok) return bad_arg("sa_mlp_pack_params: a folded weight or bias is outside the fp16 range (|v| > 65504)");
    return 0;
}

G4D_API int g4d_sa_mlp_max(const g4d_sa_mlp_desc* d, const void* params_dev, int b, int n, int m, const float* xyz,
                           const float* new_xyz, const int* idx, const void* feat_pm, float* out_cm, void* out_pm,
                           int out_c_total, int out_c_off, void* stream) {
    // the layout (a few getenv calls for tuning knobs) is planned once per descriptor and cached
    static g4d_sa_mlp_desc cached_d[16];
    static SaMlpLayout cached_L[16];
    static int ncached = 0;
    SaMlpArgs a;
    const char* why = nullptr;
    if (!d) return bad_arg("sa_mlp: null descriptor");
    int hit = -1;
    for (int i = 0; i < ncached; ++i) if (!memcmp(&cached_d[i], d, sizeof(*d))) { hit = i; break; }
    if (hit >= 0) a.L = cached_L[hit];
    else {
        if (!make_layout(d, &a.L, &why)) return bad_arg(why ? why : "sa_mlp: bad descriptor");
        if (ncached < 16) { cached_d[ncached] = *d; cached_L[ncached] = a.L; ++ncached; }
    }
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
    a.xyz = xyz; a.new_xyz = new_xyz; a.idx = idx; a.feat_pm = (const __half*)feat_pm;
    a.params = (const unsigned char*)params_dev;
    a.out_cm = out_cm; a.out_pm = (__half*)out_pm; a.ctot = out_c_total; a.coff = out_c_off;
    static const int spin = env_int("G4D_SA_SPIN", 0);
    a.flags = spin ? 1 : 0;
    a.dbg = g_timeline;

    typedef void (*kern_t)(const SaMlpArgs);
    kern_t kern = nullptr;
    const bool feat = d->c_in > 0;
    const bool big = a.L.threads > 416;
#define G4D_SA_PICK(NS_) \
    kern = big ? (feat ? sa_mlp_max_kernel<NS_, true, true> : sa_mlp_max_kernel<NS_, false, true>) \
               : (feat ? sa_mlp_max_kernel<NS_, true, false> : sa_mlp_max_kernel<NS_, false, false>)
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
    long long grid = (long long)sm_count();                // persistent: one CTA per SM
    if (grid > a.ntiles) grid = a.ntiles;
    kern<<<(unsigned)grid, a.L.threads, a.L.total_smem, (cudaStream_t)stream>>>(a);
    return finish_launch("g4d sa_mlp_max");
}
