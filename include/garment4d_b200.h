/*
 * garment4d_b200 -- C ABI of the B200 (sm_100a) implementation of Garment4D's data-parallel hot path.
 *
 * This header is the drop-in boundary.  Every entry point takes plain device pointers, sizes and a
 * CUDA stream (passed as void*, i.e. a cudaStream_t; NULL = default stream), launches asynchronously
 * on that stream, allocates nothing, keeps no state, and returns a cudaError_t as int (0 = success);
 * g4d_last_error() gives the text.  All floats are IEEE fp32, all indices int32, all tensors
 * contiguous, in the reference's layouts: coordinates (B,N,3), features channel-major (B,C,N),
 * neighbourhood indices (B,P,K).  The current CUDA device must be the one that owns the pointers
 * (as in the reference, which takes at::cuda::getCurrentCUDAStream() without a device guard).
 *
 * Section 1 mirrors, one to one, the kernel launchers that the reference's Python extension
 * `pointnet2_cuda` binds (modules/pointnet2/pointnet2/src/pointnet2_api.cpp:10-23); the caller-side
 * pre-conditions are the reference's: FPS `temp` pre-filled with 1e10 (pointnet2_utils.py:26),
 * ball-query `idx` zero-filled (:218), *_grad outputs zero-filled (:67,146,190).
 * Section 2 holds the fused forms that sit behind the same Python operators/modules.
 * Section 3 is SMPL linear-blend skinning (smplx/smplx/lbs.py).
 *
 * Unlike the reference, a failed launch never calls exit(-1) (e.g. sampling_gpu.cu:39-43): it
 * returns the error and the Python layer raises.
 */
#ifndef GARMENT4D_B200_H
#define GARMENT4D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 0. plumbing ------------------------------------------------------------------------------- */
const char* g4d_last_error(void);       /* text of the calling thread's last failure */
int g4d_abi_version(void);
int g4d_sm_count(void);
unsigned long long g4d_launch_count(void);   /* kernels launched through this library since it was loaded */

/* ---- 1. one-to-one replacements of the reference launchers ------------------------------------- */

/* furthest_point_sampling_kernel_launcher  (sampling_gpu.h:26-27, sampling_gpu.cu:211-253; bound at
 * sampling.cpp:36-46).  xyz (b,n,3) -> idx (b,m); temp (b,n) in/out scratch.  Bit-exact indices,
 * including the reference's tie-break order (bit-reversed thread slot of its shared-memory tree). */
int g4d_furthest_point_sampling(int b, int n, int m, const float* xyz, float* temp, int* idx, void* stream);

/* gather_points_kernel_launcher_fast  (sampling_gpu.h:14-15, sampling_gpu.cu:26-43; sampling.cpp:11-21).
 * out[b,c,j] = points[b,c,idx[b,j]] */
int g4d_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx, float* out, void* stream);

/* gather_points_grad_kernel_launcher_fast  (sampling_gpu.h:20-21, sampling_gpu.cu:65-83; sampling.cpp:24-34) */
int g4d_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx, float* grad_points, void* stream);

/* ball_query_kernel_launcher_fast  (ball_query_gpu.h:12-13, ball_query_gpu.cu:48-67; ball_query.cpp:14-25).
 * NOTE argument order: new_xyz (b,m,3) BEFORE xyz (b,n,3), as at the reference call site.
 * idx (b,m,nsample): first nsample points with d^2 < radius^2 in ascending index order, padded with the
 * first hit; rows without any hit are not written.  Bit-exact. */
int g4d_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz, const float* xyz, int* idx, void* stream);

/* group_points_kernel_launcher_fast  (group_points_gpu.h:13-14, group_points_gpu.cu:69-86; group_points.cpp:25-36).
 * out[b,c,p,s] = points[b,c,idx[b,p,s]] */
int g4d_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out, void* stream);

/* group_points_grad_kernel_launcher_fast  (group_points_gpu.h:19-20, group_points_gpu.cu:28-44; group_points.cpp:11-22) */
int g4d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out, const int* idx, float* grad_points, void* stream);

/* three_nn_kernel_launcher_fast  (interpolate_gpu.h:16-17, interpolate_gpu.cu:55-74; interpolate.cpp:14-23).
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) SQUARED distances, idx (b,n,3).  Bit-exact. */
int g4d_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx, void* stream);

/* three_interpolate_kernel_launcher_fast  (interpolate_gpu.h:21-22, interpolate_gpu.cu:99-117; interpolate.cpp:26-39).
 * points (b,c,m), idx/weight (b,n,3) -> out (b,c,n).  Bit-exact (same FMUL/FFMA order). */
int g4d_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out, void* stream);

/* three_interpolate_grad_kernel_launcher_fast  (interpolate_gpu.h:27-28, interpolate_gpu.cu:144-161; interpolate.cpp:42-54) */
int g4d_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points, void* stream);

/* ---- 2. fused forms (behind the same Python operators / nn.Modules) ----------------------------- */

/* FPS + centroid gather in one launch: idx (b,m) and new_xyz (b,m,3).  Replaces
 * furthest_point_sample -> gather_operation -> two transpose().contiguous() copies
 * (pointnet2_modules.py:30-35).  scratch (b,n) floats pre-filled with 1e10 is needed only for n > 16384. */
int g4d_fps_gather(int b, int n, int m, const float* xyz, int* idx, float* new_xyz, float* scratch, void* stream);

/* Two ball queries (the two scales of a PointnetSAModuleMSG, pointnet2_modules.py:37-38) from ONE scan. */
int g4d_ball_query2(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1, int* idx1,
                    const float* new_xyz, const float* xyz, void* stream);

/* Grouping stage of QueryAndGroup alone, from a given idx (b,m,nsample), nsample % 4 == 0: group(xyz) - centroid,
 * group(features) and the concatenation in one write pass (pointnet2_utils.py:251-258); with g4d_ball_query[2][_grid] in front of it
 * this is QueryAndGroup.forward (pointnet2_utils.py:243-265).  features (b,c,n) may be NULL with c = 0.
 * out: (b, 3+c, m, nsample) when use_xyz, else (b, c, m, nsample). */
int g4d_group_fused(int b, int n, int m, int c, int nsample, int use_xyz, const float* xyz, const float* new_xyz,
                    const float* features, const int* idx, float* out, void* stream);
/* = g4d_group_fused with the features POINT-major, feat_pm (b, n, c) fp32 (c > 0): contiguous 256-byte gathers and a
 * shared-memory transpose to the channel-major output instead of one 4-byte gather per channel. */
int g4d_group_fused_pm(int b, int n, int m, int c, int nsample, int use_xyz, const float* xyz, const float* new_xyz,
                       const float* feat_pm, const int* idx, float* out, void* stream);

/* Uniform-grid acceleration of the neighbour searches; results identical to the brute-force entry points above.
 * g4d_grid_build sorts each cloud's points by cell (cell edge >= min_cell, grown until <= 4096 cells; min_cell <= -1:
 * automatic, -min_cell cells along the longest axis).  grid: device buffer of g4d_grid_bytes(b,n), 16-byte aligned. */
size_t g4d_grid_bytes(int b, int n);
int g4d_grid_build(int b, int n, const float* xyz, float min_cell, void* grid, void* stream);
/* = g4d_fps_gather (identical idx / new_xyz) given any grid over xyz: threads own compact clumps of the cell-sorted
 * points and skip, exactly, every min-distance update that cannot change anything.  1 <= n <= 8192. */
int g4d_fps_gather_grid(int b, int n, int m, const void* grid, int* idx, float* new_xyz, void* stream);
/* = g4d_fps_gather (identical idx / new_xyz) through the Morton-ordered, warp-row pruned kernel (fps_rows.cu): the cloud is
 * first sorted by the 15-bit Morton code of a 32^3 grid over its bounding cube, 32 consecutive points form a clump that all
 * 32 lanes of a warp update together.  1 <= n <= 16384.  workspace: device buffer of g4d_fps_workspace_bytes(b, n),
 * 16-byte aligned (the sorted copy; scratch). */
/* optional: number of clouds the caller's whole step samples concurrently over all its streams (0 = unknown); steers the kernel
 * shape of g4d_fps_gather_ws / g4d_fps_gather_grid only, never the results */
void g4d_fps_concurrency_hint(int step_clouds);
size_t g4d_fps_workspace_bytes(int b, int n);
int g4d_fps_gather_ws(int b, int n, int m, const float* xyz, int* idx, float* new_xyz, void* workspace, void* stream);
/* = g4d_ball_query2 (idx1 = NULL: = g4d_ball_query) given a grid over xyz with min_cell >= max radius; n <= 65536 */
int g4d_ball_query2_grid(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1, int* idx1,
                         const float* new_xyz, const void* grid, void* stream);
/* = g4d_three_nn given a grid over the KNOWN points; unknown_grid (optional) only fixes a coherent processing order */
/* = g4d_ball_query2_grid; query_grid (optional): g4d_grid_build over new_xyz (b,m,3), any cell size -- only the processing order of
 * the queries (spatial neighbours share a thread block, their candidate runs stay in L1). */
int g4d_ball_query2_grid_ordered(int b, int n, int m, float radius0, int nsample0, int* idx0, float radius1, int nsample1, int* idx1,
                                 const float* new_xyz, const void* grid, const void* query_grid, void* stream);
int g4d_three_nn_grid(int b, int n, int m, const float* unknown, const void* known_grid, const void* unknown_grid,
                      float* dist2, int* idx, void* stream);

/* Grouped shared-MLP + max-pool on the tcgen05 tensor cores: for every centroid p, gathers its nsample
 * neighbours' [xyz - centroid, features] rows, runs the 3-layer 1x1-conv MLP (eval-mode BatchNorm folded
 * into weight/bias, ReLU) and max-pools over the neighbourhood -- SharedMLP + F.max_pool2d of
 * _PointnetSAModuleBase.forward (pointnet2_modules.py:37-51; pytorch_utils.py:5-32) without ever
 * materialising the (b, C, m, nsample) grouped tensor.  See g4d_sa_mlp_* below. */
typedef struct g4d_sa_mlp_desc {
    int c_in;          /* feature channels of the source points (0 = xyz only)                      */
    int c1, c2, c3;    /* MLP widths; c1,c2 multiples of 16 and <= 256; c3 <= 256                   */
    int nsample;       /* neighbourhood size: 16, 32, 64 or 128                                     */
    int k0;            /* padded input width of layer 1 in fp16 elements (from g4d_sa_mlp_k0)       */
} g4d_sa_mlp_desc;

/* Padded layer-1 K for a given c_in (c_in + 9 split-precision xyz slots + 2 bias slots, rounded up to 16). */
int g4d_sa_mlp_k0(int c_in);
/* Bytes of the packed parameter blob (fp16 UMMA-canonical weights with the layer-1/2 biases folded in as extra K
 * positions, fp32 layer-3 biases, one constant operand) for a descriptor. */
size_t g4d_sa_mlp_param_bytes(const g4d_sa_mlp_desc* d);
/* Packs host fp32 folded weights w1 (c1, 3+c_in), w2 (c2,c1), w3 (c3,c2) and biases into `blob` (host memory).
 * Fails (cudaErrorInvalidValue, "fp16 range") when a value does not fit fp16: take the operator route then. */
int g4d_sa_mlp_pack_params(const g4d_sa_mlp_desc* d, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* w3, const float* b3, void* blob);
/* xyz (b,n,3), new_xyz (b,m,3), idx (b,m,nsample), feat_pm: point-major fp16 features (b,n,c_in) or NULL.
 * out_cm: fp32 channel-major (b, out_c_total, m) written at channel offset out_c_off (fuses the torch.cat of
 * the MSG branches, pointnet2_modules.py:55); out_pm: optional fp16 point-major (b, m, out_c_total) copy for
 * the next level's gather (may be NULL). */
int g4d_sa_mlp_max(const g4d_sa_mlp_desc* d, const void* params_dev, int b, int n, int m, const float* xyz,
                   const float* new_xyz, const int* idx, const void* feat_pm, float* out_cm, void* out_pm,
                   int out_c_total, int out_c_off, void* stream);

/* debug aid: clock64() timeline of slot 0 of CTA 0 of the following g4d_sa_mlp_max launches (buf: >= 256 int64 on the device;
 * NULL = off) */
void g4d_debug_timeline(void* buf);
/* debug aid: clock() phase sums (32 warps x 20 words) of cloud 0 of the last pruned-FPS launch made with G4D_FPS_PROF=1 */
int g4d_debug_fps_phases(unsigned* out640);
/* debug aid: role-level cycle counters (16 int64) of CTA 0 of the last g4d_mlp2_rows launch made with G4D_MLP2_PROF=1 */
int g4d_debug_mlp2_counters(long long* out16);

/* Fused feature propagation (no skip features) + optional segmentation head on tcgen05: inverse-distance weights
 * from three_nn's squared distances, 3-tap interpolation, the FP module's 2-layer 1x1-conv MLP (eval BN folded, ReLU)
 * and, when h1 > 0, Conv1d(c2,h1)+BN+ReLU -> Conv1d(h1,h2) -- PointnetFPModule.forward (pointnet2_modules.py:131-156)
 * + FC_layer (modules/pointnet2encoder.py:98-101,143) for the finest level, in one kernel. */
typedef struct g4d_fp_desc {
    int c_in;          /* channels of the known (coarse) features = K of layer 1; multiple of 16, <= 256   */
    int c1, c2;        /* FP MLP widths, multiples of 16, <= 256                                          */
    int h1, h2;        /* head: hidden width (multiple of 16) and classes (<= 16); h1 = 0: no head        */
} g4d_fp_desc;
size_t g4d_fp_param_bytes(const g4d_fp_desc* d);
/* host fp32 folded weights: w1 (c1,c_in), w2 (c2,c1), wh1 (h1,c2), wh2 (h2,h1) and biases -> blob (host memory) */
int g4d_fp_pack_params(const g4d_fp_desc* d, const float* w1, const float* b1, const float* w2, const float* b2,
                       const float* wh1, const float* bh1, const float* wh2, const float* bh2, void* blob);
/* dist2/idx (b,n,3) as written by g4d_three_nn; known_pm: point-major fp16 (b,m,c_in);
 * out_feat (b,c2,n) fp32 channel-major; out_head (b,n,h2) fp32 (NULL when h1 = 0). */
int g4d_fp_interp_mlp(const g4d_fp_desc* d, const void* params_dev, int b, int n, int m, const float* dist2, const int* idx,
                      const void* known_pm, float* out_feat, float* out_head, void* stream);

/* = g4d_fp_interp_mlp plus out_label (b,n) uint8: arg-max over the h2 head outputs of every point (torch.argmax semantics), the
 * segmentation labels of modules/mesh_encoder.py:113, written by the kernel's last epilogue. */
int g4d_fp_interp_mlp_labels(const g4d_fp_desc* d, const void* params_dev, int b, int n, int m, const float* dist2, const int* idx,
                             const void* known_pm, float* out_feat, float* out_head, unsigned char* out_label, void* stream);
/* debug aid: cycle counters of CTA 0 of the following g4d_fp_interp_mlp launches (buf: >= 22 int64 on the device; NULL = off) */
void g4d_debug_fp_counters(void* buf);

/* y[b,c,:] = max(y[b,c,:] + bias[c], 0) in place (relu = 0: bias only); channel-major (b,c,n), b*c <= 65535.  One-pass
 * epilogue for the feature-propagation 1x1 convolutions that stay on the library GEMM (pointnet2_modules.py:154). */
int g4d_bias_relu_inplace(int b, int c, long long n, float* y, const float* bias, int relu, void* stream);

/* Front half of PointnetFPModule.forward (pointnet2_modules.py:138-152) in one pass: inverse-distance weights from
 * three_nn's SQUARED distances (1/(sqrt(d2)+1e-8), normalised; bit-identical to the torch expression), three_interpolate
 * of known_feats (b,c2,m) and the concatenation with the skip features skip (b,c1,n) (NULL with c1 = 0):
 * out (b, c2+c1, n) = cat([interpolated, skip], dim=1).  Replaces 5 torch kernels + three_interpolate + torch.cat. */
int g4d_fp_interp_concat(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const float* known_feats,
                         const float* skip, float* out, void* stream);

/* The same with the result as fp16 in (c, b, n) layout, out_h[(ci * b + bi) * n + pt]: the operand of ONE batched
 * (Cout x Cin) . (Cin x b*n) GEMM for the module's 1x1 convolutions (pointnet2_modules.py:153-154). */
int g4d_fp_interp_concat_cbn_h(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const float* known_feats,
                               const float* skip, void* out_h, void* stream);
/* g4d_fp_interp_concat_cbn_h with the known features given fp16 point-major, known_pm_h (b, m, c2), c2 % 8 == 0 (the copy
 * the fused levels emit for the next gather): three contiguous rows per point instead of 3 scattered words per channel. */
int g4d_fp_interp_concat_pm_cbn_h(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const void* known_pm_h,
                                  const float* skip, void* out_h, void* stream);
/* Point-major form of the same front half: known_pm_h (b, m, c2) and skip_pm_h (b, n, c1) fp16 point-major (NULL with c1 = 0),
 * c2 % 8 == c1 % 8 == 0 -> out_rows_h (b*n, c2+c1) fp16 row-major = the activation operand of x @ W^T. */
int g4d_fp_interp_concat_rows_h(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const void* known_pm_h,
                                const void* skip_pm_h, void* out_rows_h, void* stream);
/* y (rows, c) fp16 row-major, in place: y[r,ch] = max(y[r,ch] + bias[ch], 0) (relu = 0: bias only); c % 8 == 0 */
int g4d_bias_relu_rows_h(long long rows, int c, void* y_h, const float* bias, int relu, void* stream);
/* last layer of that route: yin_rows (b, n, c) fp32 pre-activations -> out_cm (b, c, n) fp32 = act(yin + bias) and, when
 * out_pm != NULL, the same values fp16 point-major (b, n, c) */
int g4d_bias_relu_rows_unpack(int b, int c, int n, const float* yin_rows, const float* bias, int relu, float* out_cm, void* out_pm,
                              void* stream);
/* y (c, len) fp16, in place: y[ch,:] = max(y[ch,:] + bias[ch], 0) (relu = 0: bias only); len % 8 == 0 */
int g4d_bias_relu_h(int c, long long len, void* y_h, const float* bias, int relu, void* stream);
/* epilogue of the last layer of that route: yin (c, b, n) pre-activations (fp32, or fp16 when in_half) ->
 * out_cm (b, c, n) fp32 = act(yin + bias) and, when out_pm != NULL, the same values fp16 point-major (b, n, c) */
int g4d_bias_relu_unpack(int b, int c, int n, const void* yin_cbn, int in_half, const float* bias, int relu, float* out_cm,
                         void* out_pm, void* stream);

/* g4d_bias_relu_inplace that ALSO writes the activated values as fp16 point-major out_pm (b,n,c): the gather layout of
 * g4d_fp_interp_mlp / g4d_sa_mlp_max (replaces transpose(1,2).to(half).contiguous() on the next level's input). */
int g4d_bias_relu_pm(int b, int c, int n, float* y, const float* bias, int relu, void* out_pm, void* stream);

/* ---- 3. SMPL linear-blend skinning (smplx/smplx/lbs.py) ----------------------------------------- */

/* Two-layer 1x1-conv MLP (eval BN folded, ReLU after both) over point-major fp16 rows on tcgen05 with the weights streamed
 * through a shared-memory ring -- the MLPs of the coarser PointnetFPModule levels (pointnet2_modules.py:138-156, widths
 * pointnet2encoder.py:91-96).  x (b*n, c_in) fp16 row-major (g4d_fp_interp_concat_rows_h) -> out_cm (b, c2, n) fp32 and,
 * when out_pm != NULL, (b, n, c2) fp16 point-major. */
typedef struct g4d_mlp2_desc {
    int c_in;          /* multiple of 32                                   */
    int c1;            /* hidden width, multiple of 32 in [64, 512]        */
    int c2;            /* output width, multiple of 16 in [16, 256]        */
} g4d_mlp2_desc;
size_t g4d_mlp2_param_bytes(const g4d_mlp2_desc* d);
/* w1 (c1, c_in), w2 (c2, c1): folded fp32 weights; fails ("fp16 range") when one does not fit fp16 */
int g4d_mlp2_pack_params(const g4d_mlp2_desc* d, const float* w1, const float* b1, const float* w2, const float* b2, void* blob);
int g4d_mlp2_rows(const g4d_mlp2_desc* d, const void* params_dev, int b, int n, const void* x_h, float* out_cm, void* out_pm,
                  void* stream);

/* ---- deterministic backward ----
 * The reference's three backward kernels (group_points_grad group_points_gpu.cu:8-25, gather_points_grad sampling_gpu.cu:46-63,
 * three_interpolate_grad interpolate_gpu.cu:120-142) are atomicAdd scatters: out[b,c,dst[b,e]] += w[b,e] * grad[b,c,e].  These two
 * calls compute the same sums as a segmented reduction in ascending source order: bit-identical from run to run.
 * build: dst (b, n_src) int32 in [0, n_dst) -> index structure in workspace (g4d_scatter_det_workspace_bytes), reusable for any
 * number of apply calls; apply: weight (b, n_src) or NULL (= 1), grad (b, c, n_src / grad_div) read at [e / grad_div] (grad_div = 3
 * for three_interpolate, else 1) -> out (b, c, n_dst), every element written. */
size_t g4d_scatter_det_workspace_bytes(int b, int n_src, int n_dst);
int g4d_scatter_det_build(int b, int n_src, int n_dst, const int* dst, void* workspace, void* stream);
int g4d_scatter_det_apply(int b, int c, int n_src, int n_dst, int grad_div, const void* workspace, const float* weight,
                          const float* grad, float* out, void* stream);

/* ---- callers of the hot path inside the garment model ----
 * Garment point selection (PCAGarmentEncoderSeg.calc_segmentation_results, modules/mesh_encoder.py:109-125): per frame the
 * points whose arg-max class equals `target`, in their original order, the first n_out of them, zero-padded.
 * sem_logits (c, n, ncls) fp32, or NULL with labels (c, n) uint8 given; xyz (c, n, 3); features (c, cf, n) channel-major or NULL
 * (cf = 0) -> out_xyz (c, n_out, 3), out_feat (c, n_out, cf) point-major, out_count (c) = selected points before clipping
 * (may be NULL). */
int g4d_select_points(int c, int n, int ncls, int cf, int target, int n_out, const float* sem_logits, const unsigned char* labels,
                      const float* xyz, const float* features, float* out_xyz, float* out_feat, int* out_count, void* stream);

/* One positional-encoding unit of the GCN refinement (modules/mesh_encoder.py:450-466): QueryAndGroup -> Linear(3+c,32) -> ReLU ->
 * Linear(32,32) -> max over the nsample neighbours, fp32, one kernel.  feat_pm (b,n,c) fp32 POINT-major (NULL with c = 0); idx
 * (b,p,nsample) from the ball query; w1t (3+c,32), w2t (32,32): the Linear weights transposed -> out (b,p,32), argmax (b,p,32) int8
 * (sample holding each maximum; may be NULL).  nsample in {4, 8, 16, 32}. */
int g4d_pe_mlp_max(int b, int n, int p, int c, int nsample, const float* xyz, const float* new_xyz, const float* feat_pm,
                   const int* idx, const float* w1t, const float* b1, const float* w2t, const float* b2, float* out,
                   signed char* argmax, void* stream);

/* ---- garment skinning by interpolated body weights: MeshEncoder.lbs_garment_interpolation (modules/mesh_encoder.py:312-410) ---- */
/* K nearest reference points of every query, ascending squared distance, equal distances by ascending index: replaces
 * chamferdist.knn_points as called at mesh_encoder.py:321-324 (its K = 64 and K = 1 results are prefixes of the K = LBSK one).
 * query (b,nq,3), ref (b,nr,3) -> dist2 (b,nq,K), idx (b,nq,K) int32.  1 <= K <= min(256, nr), nr <= 8192. */
int g4d_knn_points(int b, int nq, int nr, int K, const float* query, const float* ref, float* dist2, int* idx, void* stream);
/* mesh_encoder.py:341-345 / :371-375: w (rows,k) = 1 / dist2[:, :k] with inf -> 0, divided by the row sum, inf -> 0; dist2 (rows,ld) */
int g4d_knn_inverse_weights(long long rows, int ld, int k, const float* dist2, float* w, void* stream);
/* mesh_encoder.py:339-346 / :377-379 without the (F, body_v, K, J) intermediate: out (B*T,nq,J) = sum_k w (B,nq,K)[..,k] *
 * W (B*T,P,J)[f, idx (B,nq,ld_idx)[..,k], :] with f = b*T + t.  J <= 32. */
int g4d_knn_blend_weights(int B, int T, int nq, int P, int J, int K, int ld_idx, const int* idx, const float* w, const float* W,
                          float* out, void* stream);
/* mesh_encoder.py:382-389: `iters` steps of x <- x + coeff * Adj . x per frame; x (F,G,J) in/out, tmp (F,G,J) scratch,
 * Adj (G x G) in CSR (rowptr G+1, col, val). */
int g4d_smooth_weights(int F, int G, int J, int iters, float coeff, const int* rowptr, const int* col, const float* val, float* x,
                       float* tmp, void* stream);

/* batch_rodrigues (lbs.py:312-346): rot_vecs (n,3) -> rot_mats (n,3,3) */
int g4d_batch_rodrigues(int n, const float* rot_vecs, float* rot_mats, void* stream);
/* blend_shapes (smplx/smplx/lbs.py:288-309): betas (F,NB), shape_disps (V,3,NB) -> out (F,V,3) displacements */
int g4d_blend_shapes(int F, int V, int NB, const float* betas, const float* shape_disps, float* out, void* stream);
/* vertices2joints / vertices2jointsB (lbs.py:251-286): J_regressor (J,V) or, per_frame_regressor=1, (F,J,V) */
int g4d_vertices2joints(int F, int V, int J, int per_frame_regressor, const float* J_regressor, const float* vertices,
                        float* joints, void* stream);
/* batch_rigid_transform (lbs.py:362-419): rot_mats (F,J,3,3), joints (F,J,3), parents int32 (J) ->
 * posed_joints (F,J,3), rel_transforms (F,J,4,4) */
int g4d_batch_rigid_transform(int F, int J, const float* rot_mats, const float* joints, const int* parents,
                              float* posed_joints, float* rel_transforms, void* stream);
/* skinning tail (lbs.py:233-246): verts = (W.A)[v_posed;1]; W (V,J) or, per_frame_weights=1, (F,V,J) */
int g4d_lbs_skin(int F, int V, int J, int per_frame_weights, const float* v_posed, const float* A, const float* W,
                 float* verts, void* stream);
/* full lbs() (lbs.py:152-248); ws = device scratch of g4d_lbs_workspace_bytes(F,V,J) */
size_t g4d_lbs_workspace_bytes(int F, int V, int J);
int g4d_lbs(int F, int V, int J, int NB, int betas_rows, int pose2rot, const float* betas, const float* pose,
            const float* v_template, const float* shapedirs, const float* posedirs, const float* J_regressor,
            const int* parents, const float* lbs_weights, float* verts, float* joints, void* ws, size_t ws_bytes,
            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GARMENT4D_B200_H */
