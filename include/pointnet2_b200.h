/*
 * pointnet2_b200.h -- C ABI of lib3dvlp_b200.so, the B200 (sm_100a) implementation of 3DVLP's
 * PointNet++ op suite and nn_distance.
 *
 * Boundary.  In the reference the Python layer (lib/pointnet2/pointnet2_utils.py:25-33) imports the
 * pybind11 module `pointnet2._ext` (lib/pointnet2/_ext_src/src/bindings.cpp:11-24); inside it every
 * op crosses a C-style launcher seam (raw device pointers + ints).  The functions below replace
 * exactly those launchers, one for one, with three changes:
 *   - the CUDA stream is an explicit argument (`stream`, a cudaStream_t passed as void*); the
 *     reference reads at::cuda::getCurrentCUDAStream() inside every launcher;
 *   - every function returns a status (0 = PN2_OK) and leaves a message for pn2_last_error();
 *     the reference prints and calls exit(-1) on a failed launch (include/cuda_utils.h:35-44);
 *   - no torch type appears in any signature.
 * All pointers are DEVICE pointers on the current CUDA device; tensors are dense row-major fp32 /
 * int32 exactly as the reference's CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT demand
 * (include/utils.h:10-30).  Outputs are written by the callee and allocated by the caller; where the
 * reference relies on a zero-filled output (ball_query.cpp:24-26, the *_grad outputs) this is noted.
 * Calls are asynchronous on `stream`, stateless and re-entrant.
 */
#ifndef POINTNET2_B200_H_
#define POINTNET2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_OK 0
#define PN2_ERR_INVALID 1 /* bad argument (shape, null pointer, unsupported size) */
#define PN2_ERR_CUDA 2    /* a CUDA runtime call or kernel launch failed */

/* Library identity / diagnostics. */
int pn2_abi_version(void);
const char* pn2_last_error(void);     /* thread-local; valid until the next failing call */
int pn2_device_arch(int* major, int* minor); /* compute capability of the current device */
/* Synchronises `stream` and reads (then clears) the device-side protocol error flag: 0 = none,
 * 1 = FPS cluster exchange timed out, 2 = tensor-core pipeline timed out.  Debug aid. */
int pn2_poll_device_error(int* flag_out, void* stream);
/* Host-side launch statistics (process-wide, monotonic): how often a kernel variant was chosen --
 * "fps_cluster_launches", "fps_bucket_launches", "fps_pair_launches", "bq_grid_launches",
 * "bq_small_grid_launches", "bq_scan_launches", "nn_small_launches", "nn_tiled_launches",
 * "wgrad_tma_launches", "wgrad_cpasync_launches".
 * The parity tests read them to prove that the variant they mean to cover is the one that ran.
 * Returns -1 for an unknown name. */
long long pn2_get_stat(const char* name);
/* Run-time options (process-wide; initial value from the environment variable in brackets):
 *   "fps_bucket" [PN2_FPS_BUCKET, default 0]  1 = large scenes (16 384..131 072 points) use the single-SM
 *                bucketed FPS kernel with spatial pruning instead of the 16-SM cluster kernel: same
 *                output, ~2x the latency of one batch, 1/16 of the SM footprint (throughput mode);
 *   "bq_grid"    [PN2_BQ_GRID, default 1]     0 = ball query always uses the linear scan;
 *   "sa_tiles_per_group" [PN2_SA_TILES_PER_GROUP, default 0]  k > 0 = pn2_sa_group_mlp launches fewer
 *                persistent CTAs, each 128-thread group claiming >= k row tiles, so the per-CTA prologue
 *                (weight TMA, TMEM allocation) is paid less often: same output, longer single launch,
 *                less SM-time (throughput mode: several batches in flight);
 *   "wgrad_tma"  [PN2_WGRAD_TMA, default 1]   0 = pn2_wgrad_*_bf16 keeps the cp.async operand copies instead of
 *                the tiled-TMA (tensor-map) kernel: same result; also the automatic fallback when a tensor map
 *                cannot be encoded.
 * pn2_get_option returns the value, or -1 for an unknown name. */
int pn2_set_option(const char* name, int value);
int pn2_get_option(const char* name);
/* Per-thread override of an option for calls made by the CALLING host thread only (value < 0 removes
 * the override); pn2_get_option on that thread reports the override.  Entry points are re-entrant and
 * may be called from several host threads (the reference is driven from DataParallel worker threads):
 * a thread that captures a CUDA graph with other settings must not change what its neighbours launch. */
int pn2_set_option_thread(const char* name, int value);

/* ---------------------------------------------------------------------------------------------
 * The nine launchers of pointnet2._ext
 * ------------------------------------------------------------------------------------------- */

/* furthest_point_sampling_kernel_wrapper(int b,int n,int m,const float* dataset,float* temp,int* idxs)
 * -- src/sampling.cpp:16-18, kernel src/sampling_gpu.cu:75-178.
 * dataset (b,n,3); idxs (b,m) int32.  Greedy FPS from index 0 with the reference's skip of points
 * with |p|^2 <= 1e-3 and its thread/tree tie-break (bit-exact indices).  `temp` (b,n) is the
 * reference's running-minimum scratch: this implementation keeps the minima on chip and never
 * touches it (NULL is accepted).  m <= 0 returns PN2_OK without writing (sampling_gpu.cu:78). */
int pn2_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs,
                                void* stream);

/* Same, fused with the gather that always follows it (pointnet2_modules.py:233-240): also writes
 * new_xyz (b,m,3) = dataset[idxs].  new_xyz may be NULL. */
int pn2_furthest_point_sampling_gather(int b, int n, int m, const float* dataset, int* idxs,
                                       float* new_xyz, void* stream);

/* gather_points_kernel_wrapper(int b,int c,int n,int npoints,const float* points,const int* idx,
 * float* out) -- src/sampling.cpp:9-11, kernel sampling_gpu.cu:13-25.
 * points (b,c,n), idx (b,npoints) -> out (b,c,npoints). */
int pn2_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                      float* out, void* stream);

/* gather_points_grad_kernel_wrapper -- src/sampling.cpp:12-14, kernel sampling_gpu.cu:39-52.
 * grad_out (b,c,npoints), idx (b,npoints) -> grad_points (b,c,n) += ; caller zero-fills. */
int pn2_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out, const int* idx,
                           float* grad_points, void* stream);

/* query_ball_point_kernel_wrapper(int b,int n,int m,float radius,int nsample,const float* new_xyz,
 * const float* xyz,int* idx) -- src/ball_query.cpp:9-11, kernel ball_query_gpu.cu:14-49.
 * new_xyz (b,m,3), xyz (b,n,3) -> idx (b,m,nsample): the first nsample points (ascending index)
 * with d2 < radius*radius, padded with the first hit; an empty ball yields zeros (this
 * implementation writes them itself, no zero-fill needed). */
int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                   const float* xyz, int* idx, void* stream);

/* group_points_kernel_wrapper(int b,int c,int n,int npoints,int nsample,const float* points,
 * const int* idx,float* out) -- src/group_points.cpp:9-11, kernel group_points_gpu.cu:13-33.
 * points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample). */
int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                     const int* idx, float* out, void* stream);

/* group_points_grad_kernel_wrapper -- src/group_points.cpp:13-15, kernel group_points_gpu.cu:48-69.
 * grad_out (b,c,npoints,nsample) -> grad_points (b,c,n) += ; caller zero-fills. */
int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                          const int* idx, float* grad_points, void* stream);

/* three_nn_kernel_wrapper(int b,int n,int m,const float* unknown,const float* known,float* dist2,
 * int* idx) -- src/interpolate.cpp:9-11, kernel interpolate_gpu.cu:14-64.
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) SQUARED distances, idx (b,n,3) int32. */
int pn2_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                 int* idx, void* stream);

/* three_interpolate_kernel_wrapper(int b,int c,int m,int n,const float* points,const int* idx,
 * const float* weight,float* out) -- src/interpolate.cpp:12-13, kernel interpolate_gpu.cu:77-106.
 * points (b,c,m), idx/weight (b,n,3) -> out (b,c,n). */
int pn2_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                          const float* weight, float* out, void* stream);

/* three_interpolate_grad_kernel_wrapper(int b,int c,int n,int m,...) -- src/interpolate.cpp:14-17,
 * kernel interpolate_gpu.cu:121-148.  grad_out (b,c,n) -> grad_points (b,c,m) += ; caller
 * zero-fills.  NOTE: the reference BINDING (interpolate.cpp:95) launches the forward kernel by
 * mistake; this is the scatter-add its kernel file intends. */
int pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                               const float* weight, float* grad_points, void* stream);

/* ---------------------------------------------------------------------------------------------
 * nn_distance -- utils/nn_distance.py:32-59 (pure torch in the reference; one fused kernel here)
 * ------------------------------------------------------------------------------------------- */
#define PN2_NN_L2 0     /* sum (a-b)^2            (default)            */
#define PN2_NN_L1 1     /* sum |a-b|              (l1=True)            */
#define PN2_NN_HUBER 2  /* sum huber(a-b, delta)  (l1smooth=True)      */

/* pc1 (b,n,c), pc2 (b,m,c) -> dist1 (b,n) f32, idx1 (b,n) i64, dist2 (b,m) f32, idx2 (b,m) i64.
 * Ties keep the first index (torch.min). */
int pn2_nn_distance(int b, int n, int m, int c, const float* pc1, const float* pc2, int mode,
                    float delta, float* dist1, int64_t* idx1, float* dist2, int64_t* idx2,
                    void* stream);

/* Backward of the above w.r.t. pc1 and pc2 given grad_dist1 (b,n) / grad_dist2 (b,m) (either may
 * be NULL) and the forward indices.  grad_pc1 (b,n,c) / grad_pc2 (b,m,c) += ; caller zero-fills. */
int pn2_nn_distance_grad(int b, int n, int m, int c, const float* pc1, const float* pc2, int mode,
                         float delta, const int64_t* idx1, const int64_t* idx2,
                         const float* grad_dist1, const float* grad_dist2, float* grad_pc1,
                         float* grad_pc2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Axis-aligned box IoU / DIoU -- utils/box_util.py:488-528 `box3d_diou_batch_tensor` (element-wise
 * torch ops in the reference, driven by Python loops in lib/loss_helper/loss_grounding.py:205-231)
 * and the axis-aligned use of pytorch3d `box3d_overlap` in
 * models/constrast_module/constrast_module.py:105.  SURVEY.md 8f ranks 3-4.
 * ------------------------------------------------------------------------------------------- */

/* Boxes as centre + size, fp32 (.,3).  box1 = (b,k) proposals, box2 = (b,l) targets ->
 * iou / diou (b,l,k) (either output may be NULL).  Bit-identical to the reference's fp32 op tree.
 * The element-wise form of the reference (P independent pairs) is the call b=P, l=1, k=1. */
int pn2_box3d_diou_pairs(int b, int l, int k, const float* center1, const float* size1,
                         const float* center2, const float* size2, float* iou, float* diou,
                         void* stream);

/* Gradient w.r.t. the proposals, summed over the l targets: grad_iou / grad_diou (b,l,k) (either
 * may be NULL) -> grad_center1, grad_size1 (b,k,3), overwritten.  torch's tie rules for
 * max/min/clamp. */
int pn2_box3d_diou_pairs_grad(int b, int l, int k, const float* center1, const float* size1,
                              const float* center2, const float* size2, const float* grad_iou,
                              const float* grad_diou, float* grad_center1, float* grad_size1,
                              void* stream);

/* Element-wise gradient for both boxes (P pairs); any grad output may be NULL. */
int pn2_box3d_diou_grad(long long p, const float* center1, const float* size1, const float* center2,
                        const float* size2, const float* grad_iou, const float* grad_diou,
                        float* grad_center1, float* grad_size1, float* grad_center2,
                        float* grad_size2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SharedMLP on tcgen05 tensor cores -- pytorch_utils.py:11-36 (Conv2d 1x1 -> BatchNorm2d -> ReLU;
 * cuDNN in the reference).  BatchNorm (eval) is folded into W / bias by the host layer.
 * ------------------------------------------------------------------------------------------- */

/* W (n,k) fp32 row-major -> bf16, K zero-padded to kpad (multiple of 64), in the tile layout the
 * kernels below consume: kpad/64 panels of n rows x 128 bytes, SWIZZLE_128B.  out: n*kpad*2 bytes. */
int pn2_pack_weight_bf16(int n, int k, int kpad, const float* w, void* out, void* stream);

/* features (b,c,n) fp32 channel-major (the reference's feature layout) -> (b*n, kpad) bf16 rows. */
int pn2_pack_rows_bf16(int b, int c, int n, int kpad, const float* features, void* out, void* stream);

/* Rows of the input cloud as the dataset stores them -- (rows, ld) fp32, features in columns
 * [col0, col0+c) after x,y,z (models/base_module/backbone_module.py:69-74 `_break_up_pc`) -> the same
 * (rows, kpad) bf16 operand, without the transpose to (b,c,n) and back. */
int pn2_pack_point_rows_bf16(long long rows, int ld, int col0, int c, int kpad, const float* in, void* out,
                             void* stream);

/* Y = act(X W^T + bias).  X (m,k) bf16 row-major, k % 64 == 0; W packed as above (n % 16 == 0,
 * n <= 256); bias (n) fp32 or NULL; relu 0/1.  out_mode 0: Y fp32 (m,n) row-major; 1: Y bf16 (m,n)
 * row-major; 2: Y fp32 channel-major (m / rows_per_batch, n, rows_per_batch) -- the reference's
 * (B, C, npoint) feature layout. */
int pn2_linear_bf16(int m, int n, int k, const void* x_bf16, const void* w_packed, const float* bias,
                    int relu, int out_mode, int rows_per_batch, void* y, void* stream);

/* pn2_linear_bf16 on split-precision rows (training): x (m, terms * kb) bf16 stores each term once ([h | m | l],
 * blocks kb wide, kb % 64 == 0); w_packed holds 1, 3 or 6 K blocks of width kb in the product order given at
 * pn2_bn_relu_rows; logical K = {1, 3, 6} * kb. */
int pn2_linear_split_bf16(int m, int n, int kb, int terms, const void* x_bf16, const void* w_packed,
                          const float* bias, int relu, int out_mode, int rows_per_batch, void* y, void* stream);

/* Fused body of a set-abstraction layer (pointnet2_modules.py:243-257: QueryAndGroup -> SharedMLP
 * (3 layers) -> max_pool2d over nsample), eval mode, BatchNorm folded.  Grouping, the three 1x1
 * convolutions and the max-pool run in one kernel; only the pooled result is written.
 *   xyz (b,n,3), new_xyz (b,npoint,3), idx (b,npoint,nsample) from pn2_ball_query;
 *   radius_div: divisor of the recentred coordinates (radius when normalize_xyz, else 1);
 *   layer 1 = w1 (c1, 3 + c_direct) fp32 . [rel_xyz, feat] + b1 (+ g): `feat` (b,c_direct,n) are up
 *   to 4 raw channels evaluated in fp32; `g` (b*n, c1) fp32, if non-NULL, is the per-point
 *   contribution W1[:,3:] . features precomputed with pn2_linear_bf16 (wide inputs);
 *   w2_packed (c2,c1), w3_packed (c3,c2) from pn2_pack_weight_bf16; b1/b2/b3 fp32 folded biases;
 *   c1,c2 in {64,128}, c3 in {128,256}, nsample in {16,32,64,128}, b <= 64, npoint <= 4096.
 *   idx must follow pn2_ball_query's padding contract (slots past the last distinct hit repeat
 *   idx[..,0]; ball_query_gpu.cu:40-44): the repeated rows are recognised and not recomputed --
 *   their activations equal slot 0's, so the max-pool is unchanged.
 * Outputs: out_cm (b,c3,npoint) fp32 -- the reference's feature layout, zero-filled by the call --
 * and, if non-NULL, out_pm_bf16 (b*npoint, c3) bf16 point-major rows ready to be the next layer's
 * GEMM operand.  Launches four kernels (count, row plan, fused body, bf16 transpose) and takes its
 * workspace from the stream-ordered allocator: asynchronous, re-entrant, capturable. */
int pn2_sa_group_mlp(int b, int n, int npoint, int nsample, float radius_div, const float* xyz,
                     const float* new_xyz, const int* idx, int c_direct, const float* feat,
                     const float* g, int c1, int c2, int c3, const float* w1, const float* b1,
                     const void* w2_packed, const float* b2, const void* w3_packed, const float* b3,
                     float* out_cm, void* out_pm_bf16, void* stream);

/* Front end of a feature-propagation layer (pointnet2_modules.py:393-412: three_nn -> sqrt ->
 * 1/(d+1e-8) normalised weights -> three_interpolate -> cat([interpolated, unknow_feats])), fused:
 * unknown (b,n,3), known (b,m,3), unknow_feats (b,c1,n) or NULL (c1 = 0), known_feats (b,c2,m), all
 * fp32 in the reference's layouts -> out_bf16 (b*n, kpad) bf16 rows [interp c2 | skip c1 | zero pad],
 * kpad % 64 == 0: the activation operand of pn2_linear_bf16. */
int pn2_fp_prepare(int b, int n, int m, int c1, int c2, int kpad, const float* unknown,
                   const float* known, const float* unknow_feats, const float* known_feats,
                   void* out_bf16, void* stream);

/* pn2_furthest_point_sampling_gather for an input the caller believes to be in furthest-point order
 * already -- the samples of a previous FPS call, which is what SA2..SA4 of the backbone receive
 * (backbone_module.py:101-113).  The belief is verified on the device in parallel (n*m distance
 * tests without the serial chain); where it holds the result 0..m-1 is written directly, elsewhere the
 * regular kernel runs.  Output identical to pn2_furthest_point_sampling_gather in every case. */
int pn2_furthest_point_sampling_gather_ordered(int b, int n, int m, const float* dataset, int* idxs,
                                               float* new_xyz, void* stream);

/* three_nn (interpolate.cpp:19-45) fused with the inverse-distance weights of PointnetFPModule
 * (pointnet2_utils.py:138-140 sqrt, pointnet2_modules.py:395-397 1/(d+1e-8), normalise):
 * unknown (b,n,3), known (b,m,3) -> idx (b,n,3) i32, weight (b,n,3) f32.  Same neighbours, order and
 * ties as pn2_three_nn.  Depends on coordinates only: the backbone runs it on its sampling stream. */
int pn2_three_nn_weights(int b, int n, int m, const float* unknown, const float* known, int* idx,
                         float* weight, void* stream);

/* pn2_fp_prepare with the neighbours and weights given (idx, weight from pn2_three_nn_weights). */
int pn2_fp_prepare_indexed(int b, int n, int m, int c1, int c2, int kpad, const int* idx,
                           const float* weight, const float* unknow_feats, const float* known_feats,
                           void* out_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-mode SharedMLP (+ max-pool): batch-statistics BatchNorm and the backward pass
 * -- pytorch_utils.py:11-36,67-120 (Conv2d 1x1 without bias -> BatchNorm2d -> ReLU, cuDNN in the reference),
 * pointnet2_modules.py:251-257 (max_pool2d over nsample), lib/joint/solver_3dvlp.py:580-584 (train()).
 * Activations are point-major rows (P, C), P = B * npoint * nsample positions, row = (b*npoint + j)*nsample + s.
 * Forward of a layer: pn2_linear_bf16 (Y fp32) -> pn2_column_stats -> pn2_bn_finalize -> pn2_bn_relu_rows
 * (hidden) or pn2_bn_relu_maxpool (last).  Backward of a layer: pn2_bn_bwd_stats -> pn2_bn_bwd_rows ->
 * pn2_wgrad_bf16 and pn2_linear_bf16 with the transposed weight.  All asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------- */

/* sum[ch] += sum_r y[r][ch], sumsq[ch] += sum_r y[r][ch]^2 over rows of (rows, ld) fp32, ch < c <= 256.
 * Accumulators are double and are NOT cleared by the call. */
int pn2_column_stats(long long rows, int c, int ld, const float* y, double* sum, double* sumsq, void* stream);

/* BatchNorm statistics -> per-channel affine: scale = gamma * invstd, shift = beta - mean * scale.
 * count > 0: batch statistics from sum / sumsq over `count` positions (biased variance, torch semantics);
 * if running_mean / running_var are non-NULL they are updated in place with `momentum` (the running
 * variance takes the unbiased estimate).  count == 0: frozen statistics read from the running buffers.
 * gamma / beta may be NULL (1 / 0).  mean / invstd (c) are kept for the backward pass. */
int pn2_bn_finalize(int c, long long count, const double* sum, const double* sumsq, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* mean, float* invstd, void* stream);

/* Operand rows x = relu(y * scale + shift) of the next contraction: (rows, ld) fp32 -> bf16, c % 8 == 0.
 * A value is carried as `terms` bf16 terms (h = bf16(x), m = bf16(x - h), l = bf16(x - h - m)), each stored ONCE
 * per row: (rows, c) [h], (rows, 2c) [h | m] or (rows, 3c) [h | m | l].  pn2_linear_split_bf16 contracts them
 * against a weight whose K blocks spell out the leading products, each exact in the fp32 accumulator:
 *   terms 1: K blocks [h]                      x [w_h]                                   plain bf16
 *   terms 2: K blocks [h | m | h]              x [w_h | w_h | w_m]                       ~2^-16 relative
 *   terms 3: K blocks [h | m | h | l | h | m]  x [w_h | w_h | w_m | w_h | w_l | w_m]     ~2^-24 (fp32 level)
 * (the kernel maps a logical K block of the rows to the stored term block). */
int pn2_bn_relu_rows(long long rows, int c, int ld, const float* y, const float* scale, const float* shift,
                     int terms, void* x_bf16, void* stream);

/* pn2_pack_rows_bf16 in the split layouts above (terms 2 or 3, blocks of width kpad): features (b,c,n) fp32 ->
 * (b*n, 2*kpad) or (b*n, 3*kpad) bf16 rows. */
int pn2_pack_rows_split_bf16(int b, int c, int n, int kpad, int terms, const float* features, void* out,
                             void* stream);

/* Last layer: out_cm[b][ch][j] = max over the ns rows of group g = b*npoint + j of relu(y*scale + shift)
 * (fp32, the reference's (B, C, npoint) layout) and argmax[g][ch] = the first slot attaining it (int32,
 * point-major (groups, c)). */
int pn2_bn_relu_maxpool(int groups, int ns, int c, int ld, int npoint, const float* y, const float* scale,
                        const float* shift, float* out_cm, int* argmax, void* stream);

/* s1[ch] += sum dYhat, s2[ch] += sum dYhat * xhat over the rows, dYhat = upstream * [y*scale+shift > 0],
 * xhat = (y - mean) * invstd.  Upstream: dense `da` (rows, ld_da) fp32, or (da == NULL) the pooled gradient
 * dout_pm (groups, c) fp32 POINT-major (the transpose of the (B, c, npoint) gradient), routed to the arg-max
 * slot of each group of ns rows. */
int pn2_bn_bwd_stats(long long rows, int c, int ld_da, int ld_y, const float* da, const float* dout_pm,
                     const int* argmax, int ns, int npoint, const float* y, const float* scale,
                     const float* shift, const float* mean, const float* invstd, double* s1, double* s2,
                     void* stream);

/* dy = gamma * invstd * (dYhat - (s1 + xhat * s2) / count)  (count > 0: batch statistics) or
 * gamma * invstd * dYhat  (count == 0: frozen statistics) -> operand rows in the `terms` layout of
 * pn2_bn_relu_rows: (rows, c), (rows, 2c) or (rows, 3c) bf16. */
int pn2_bn_bwd_rows(long long rows, int c, int ld_da, int ld_y, const float* da, const float* dout_pm,
                    const int* argmax, int ns, int npoint, const float* y, const float* scale,
                    const float* shift, const float* mean, const float* invstd, const float* gamma,
                    const double* s1, const double* s2, long long count, int terms, void* dy_bf16,
                    void* stream);

/* dw[co][ci] += sum_r dy[r][co] * x[r][ci] on tcgen05 (bf16 operands, fp32 accumulation, fp32 atomics):
 * dy (rows, ld_dy) bf16 columns [0, c_out), x (rows, ld_x) bf16 columns [0, c_in), dw (c_out, c_in) fp32 --
 * caller zero-fills.  c_out % 8 == 0, ld_dy % 8 == 0, ld_x % 8 == 0, ld_x >= c_in rounded up to 8; pointers
 * 16-byte aligned. */
int pn2_wgrad_bf16(long long rows, int c_out, int c_in, int ld_dy, int ld_x, const void* dy_bf16,
                   const void* x_bf16, float* dw, void* stream);

/* The same contraction on split-precision operands in ONE launch: dy and x are rows in the `terms` layout of
 * pn2_bn_relu_rows / pn2_bn_bwd_rows ([h], [h|m], [h|m|l]; dy blocks c_out wide, x blocks ld_x / terms wide).
 * Every term of both operands is read once per 64-row chunk and all leading (dy term, x term) products -- 1, 3
 * or 6 -- accumulate into one TMEM tile: against one pn2_wgrad_bf16 call per product, half the operand traffic and
 * a sixth of the launches / fp32 atomics at terms = 3. */
int pn2_wgrad_split_bf16(long long rows, int c_out, int c_in, int ld_dy, int ld_x, const void* dy_bf16,
                         const void* x_bf16, int terms, float* dw, void* stream);

/* rows (b*n, ld) fp32, columns [0, c) -> out[b][ch][p] at out + b*batch_stride + ch*n + p: the gradient
 * w.r.t. a (B, C, n) tensor (or a channel slice of one) from its point-major form. */
int pn2_unpack_rows(int b, int c, int n, int ld, long long batch_stride, const float* rows, float* out,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POINTNET2_B200_H_ */
