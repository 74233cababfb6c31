/* neat_b200 -- C ABI of the B200-native NEAT attraction-field training step.
 *
 * The reference (cherubicXN/neat) has no native boundary: its hot path is Python calling ATen.
 * Each entry point below replaces one reference function (cited as file:line relative to the
 * reference tree) and is what a maintainer binds from the reference's Python side with ctypes
 * (see INTEGRATION.md).  Conventions:
 *   - all pointers are DEVICE pointers unless the name ends in _host; fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous, never synchronise and
 *     never allocate (scratch comes from the caller, sizes from neat_workspace_bytes);
 *   - return value 0 = success, otherwise a negative neat_status / positive cudaError_t.
 */
#ifndef NEAT_B200_H
#define NEAT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct neat_ctx neat_ctx;

enum neat_status { NEAT_OK = 0, NEAT_EINVAL = -1, NEAT_ENOMEM = -2, NEAT_ENODEV = -3, NEAT_EUNSUPPORTED = -4 };

/* Network shapes: code/confs/dtu.conf:28-70, code/model/networks/neat_wfr_rend_a.py:14-255 */
typedef struct {
  int sdf_layers;       /* number of Linear layers of ImplicitNetwork (len(dims)+1 = 9)            */
  int sdf_hidden;       /* hidden width (256); multiple of 32, <= 256                               */
  int sdf_skip;         /* skip_in[0] (4): input of this layer is [h, PE(x)] / sqrt(2); -1 = none  */
  int multires;         /* positional-encoding octaves of ImplicitNetwork (6)                       */
  int feat;             /* feature_vector_size (256)                                                */
  int head_layers;      /* Linear layers of the rendering / attraction nets (5)                     */
  int head_hidden;      /* their hidden width (256)                                                 */
  int multires_view;    /* view-direction octaves of RenderingNetwork (4); attraction net uses 0   */
  float sphere_radius;  /* scene_bounding_sphere (3.0)                                              */
  float sphere_scale;   /* ImplicitNetwork.sphere_scale (20.0)                                      */
} neat_net_config;

/* ---- context ------------------------------------------------------------------------------ */
int neat_create(const neat_net_config* cfg, neat_ctx** out);
void neat_destroy(neat_ctx* ctx);
const char* neat_last_error(void);
/* number of CUDA kernels this library has launched in this process (all contexts) */
long long neat_launch_count(void);

/* Flat fp32 parameter buffer shared with the host framework (effective weights, i.e. weight_norm
 * already applied: neat_wfr_rend_a.py:71-72):
 *   for net in (implicit, rendering, attraction): for layer l: W_l [out,in] row-major, then b_l [out].
 * The gradient buffer produced by the backward entry points has the same layout.              */
size_t neat_param_count(const neat_ctx* ctx);
/* net: 0 implicit, 1 rendering, 2 attraction; kind: 0 weight, 1 bias.  Returns float offset or -1. */
long neat_param_offset(const neat_ctx* ctx, int net, int layer, int kind);
/* in/out features of a layer; returns 0 or NEAT_EINVAL */
int neat_layer_dims(const neat_ctx* ctx, int net, int layer, int* in_features, int* out_features);

/* nn.utils.weight_norm for every layer in one launch (neat_wfr_rend_a.py:71-72): layers in the order of the flat
 * buffer (implicit, rendering, attraction; layer 0..).  g == NULL: plain Linear (v is the weight).  forward writes
 * the effective weights + biases into flat; backward reads flat_grad and writes gg / gv / gb.          */
typedef struct {
  const float *g, *v, *b; /* weight_g [rows], weight_v [rows, cols], bias [rows] */
  float *gg, *gv, *gb;    /* their gradients (backward only)                    */
  int rows, cols;
  long w_off, b_off;      /* float offsets in the flat buffer (filled by the library) */
} neat_wn_layer;
int neat_weight_norm_forward(neat_ctx* ctx, const neat_wn_layer* layers, int n_layers, float* flat, void* stream);
/* accumulate != 0: ADD into gg / gv / gb (a second backward before zero_grad, or gradient buffers that are
 * views of a zeroed data-parallel bucket); 0: overwrite.                                              */
int neat_weight_norm_backward(neat_ctx* ctx, const neat_wn_layer* layers, int n_layers, const float* flat_grad,
                              int accumulate, void* stream);

/* Arithmetic of the layer GEMMs: 0 (default) = bf16x3, operands split hi + lo, three MMAs per k-step, fp32
 * accumulation (meets the 1e-4 parity bound); 1 = plain bf16 operands, one MMA per k-step (~1e-2 on the SDF;
 * an explicitly opt-in fast mode, never used by the tests or the headline benchmark).             */
int neat_set_precision(neat_ctx* ctx, int fast);

/* Re-pack the flat parameters into tcgen05 operand slabs (bf16 hi/lo planes).  Call once per
 * optimizer step, before any of the entry points below.                                        */
int neat_pack_weights(neat_ctx* ctx, const float* flat_params, void* stream);

/* ---- ImplicitNetwork.get_sdf_vals (neat_wfr_rend_a.py:131-137) ----------------------------- */
/* points x[M,3] -> sdf[M] = min(net(x)[0], sphere_scale * (sphere_radius - |x|))               */
int neat_sdf_points(neat_ctx* ctx, const float* x, int M, float* sdf, void* stream);
/* the sampler's form (ray_sampler.py:146-151): x = o + z * d for z[R,n]; o is [R,3] (o_stride 3)
 * or one camera centre [3] (o_stride 0); sdf[R,n]                                              */
int neat_sdf_rays(neat_ctx* ctx, const float* rays_o, int o_stride, const float* rays_d, const float* z, int R,
                  int n, float* sdf, void* stream);
/* Dense-grid evaluation for mesh extraction (code/utils/plots.py:101-108, 318-329): the grid
 * points are generated in the kernel.  Axis k = np.linspace(lo[k], hi[k], n[k]) (float64, cast to float32); point order =
 * np.meshgrid(x, y, z) raveled (idx = (iy * nx + ix) * nz + iz), i.e. the reference's `grid_points`.  clamp = 0: the raw
 * network output implicit_network(x)[:, 0] that the mesh extraction uses; 1: get_sdf_vals' sphere clamp.  sdf [nx*ny*nz]. */
int neat_sdf_grid(neat_ctx* ctx, const double* lo, const double* hi, const int* n, int clamp, float* sdf, void* stream);

/* ---- ErrorBoundSampler.get_z_vals (code/model/ray_sampler.py:130-283) ------------------------ */
typedef struct {
  int n_eval;      /* N_samples_eval (128): samples added per iteration; <= 128                   */
  int n_final;     /* N_samples (64)                                                              */
  int n_extra;     /* N_samples_extra (32); n_final + 2 + n_extra <= 128                          */
  int beta_iters;  /* bisection steps (10)                                                        */
  int max_iters;   /* max_total_iters (5); n_eval * max_iters <= 640                              */
  float near_;     /* ray_sampler.near (0)                                                        */
  float far_;      /* 2 * scene_bounding_sphere                                                   */
  float eps;       /* error bound target (0.1)                                                    */
  float beta_min;  /* density.beta_min (1e-4): beta0 = |density.beta| + beta_min                  */
} neat_sampler_config;

size_t neat_sampler_workspace_bytes(int R);
/* Phase 1: stratified depths, the iterative SDF queries (tensor-core kernel), error-bound bisection and
 * inverse-CDF up-sampling until convergence; no host synchronisation (a device state word replaces
 * `beta.max() > beta0`, ray_sampler.py:200).  t_rand [R,n_eval] and u_final [R,n_final] are the two
 * torch.rand draws of a training call (ray_sampler.py:87,234) or NULL in eval mode (linspace).
 * n_iters_dev (device int, may be NULL) receives k.                                              */
int neat_sampler_run(neat_ctx* ctx, const neat_sampler_config* cfg, const float* rays_o, int o_stride,
                     const float* rays_d, int R, const float* beta_param, const float* t_rand,
                     const float* u_final, void* workspace, int* n_iters_dev, void* stream);
/* Phase 2: near / far / extra columns, final sort, eikonal depth (ray_sampler.py:259-276).
 * extra_idx: int64 table [max_iters][n_extra]; row k-1 holds the columns of z[R, n_eval*k] to add
 * (training: randperm(L)[:n_extra], eval: linspace(0, L-1, n_extra).long()).  eik_idx [R] int64 or NULL
 * (eval: column 0).  z_vals [R, n_final+2+n_extra], z_eik [R].                                   */
int neat_sampler_finish(neat_ctx* ctx, const neat_sampler_config* cfg, int R, const int64_t* extra_idx,
                        const int64_t* eik_idx, void* workspace, float* z_vals, float* z_eik, void* stream);

/* ---- render points ------------------------------------------------------------------------- */
/* A batch of M points, either explicit (x [M,3], optional per-point view dirs [M,3]) or generated on the
 * fly as x = o + z d, M = R*S, point index = ray*S + sample (neat_wfr_rend_a.py:392-398).        */
typedef struct {
  const float* x;       /* explicit points or NULL                                                 */
  const float* dirs;    /* explicit view directions or NULL (then rays_d[ray])                     */
  const float* rays_o;  /* [3] (o_stride 0) or [R,3] (o_stride 3)                                  */
  const float* rays_d;  /* [R,3]                                                                   */
  const float* z;       /* [R,S]                                                                   */
  int o_stride, R, S, M;
} neat_points;

/* ImplicitNetwork.get_outputs (clamp=1, neat_wfr_rend_a.py:111-129) / .gradient (clamp=0, :98-109):
 * sdf [M] (may be NULL), grad [M,3] = d sdf / d x by the analytic reverse pass, act [M] (NULL ok) = 1
 * where the network (not the bounding sphere) supplies the sdf, feat_tiles = the feature vectors as
 * bf16 hi/lo operand tiles for neat_head_forward (NULL ok).  `save` is scratch: with training=1 it is
 * the per-tile record the backward pass reads.                                                   */
size_t neat_feat_tiles_bytes(int M);
size_t neat_sdf_save_bytes(const neat_ctx* ctx, int M, int training);
int neat_sdf_outputs(neat_ctx* ctx, const neat_points* pts, int clamp, int training, float* sdf, float* grad,
                     float* act, void* feat_tiles, void* save, void* stream);

/* RenderingNetwork.forward (head 0, :235-255) -> rgb [M,3];  AttractionFieldNetwork.forward (head 1,
 * :175-197) -> lines3d [M,2,3].  normals [M,3] and feat_tiles come from neat_sdf_outputs.          */
size_t neat_head_save_bytes(const neat_ctx* ctx, int M);
int neat_head_forward(neat_ctx* ctx, int head, const neat_points* pts, const float* normals,
                      const void* feat_tiles, int training, void* save, float* out, void* stream);

/* rend_util.get_camera_params (code/utils/rend_util.py:55-81): uv [R,2], pose [4,4], K [4,4] ->
 * dirs [R,3] (unit), cam [3]                                                                     */
int neat_camera_rays(const float* uv, const float* pose, const float* K, int R, float* dirs, float* cam,
                     void* stream);

/* LaplaceDensity + VolSDFNetwork.volume_rendering + the weighted sums of VolSDFNetwork.forward
 * (density.py:21-30, neat_wfr_rend_a.py:404-429, 540-554, 530-536)                               */
typedef struct {
  int R, S;
  const float* z;          /* [R,S]   */
  const float* sdf;        /* [R,S]   */
  const float* rgb;        /* [R,S,3] or NULL: rgb_values is not computed  (the step composites the attraction  */
  const float* lines;      /* [R,S,6] or NULL: lines3d is not computed      head first, the colours later)     */
  const float* normals;    /* [R,S,3] or NULL (no normal map) */
  const float* rays_o;     /* [3]     */
  const float* rays_d;     /* [R,3]   */
  const float* beta_param; /* density.beta */
  float beta_min;
  float* weights;          /* [R,S]   or NULL */
  float* rgb_values;       /* [R,3]   */
  float* lines3d;          /* [R,6]   */
  float* depth;            /* [R]     or NULL */
  float* points3d;         /* [R,3]   or NULL */
  float* normal_map;       /* [R,3] or NULL */
  const float* bg_color;   /* [3] or NULL: white_bkgd, rgb_values += (1 - sum w) * bg_color (neat_wfr_rend_a.py:411-413) */
} neat_composite_args;
int neat_composite_forward(const neat_composite_args* a, void* stream);

/* project2D + the uv_proj ray / tangent-plane intersection (neat_wfr_rend_a.py:317-331, 433-456).
 * pose_inv [16] receives pose^-1.  lines2d, lines2d_calib [R,2,2]; l3d [R,3].                    */
int neat_line_geometry(int R, const float* pose, const float* K, const float* uv_proj, const float* points3d,
                       const float* grad3d, const float* lines3d, float* pose_inv, float* lines2d,
                       float* lines2d_calib, float* l3d, void* stream);

/* ---- junction clustering (VolSDFNetwork.cluster_dbscan, neat_wfr_rend_a.py:333-342) ------------- */
/* sklearn DBSCAN(eps, min_samples=2) + per-cluster mean == connected components (>= 2 points) of the
 * eps-graph + centroids.  points [N,3]; centroids [N/2,3] (first *n_clusters rows valid, clusters ordered by
 * their smallest point index, as sklearn labels them); n_clusters: device int.                     */
size_t neat_dbscan_workspace_bytes(int N);
int neat_dbscan(const float* points, int N, float eps, void* workspace, float* centroids, int* n_clusters,
                void* stream);

/* use_l3d junction candidates (neat_wfr_rend_a.py:454-455, 461-465; only read when dbscan_enabled is false):
 * score_i = |(l3d_i - a_i) x (l3d_i - b_i)| / |a_i - b_i| for the 3D line (a_i, b_i) = lines3d[i]; the rays with
 * score < max(median(score), 0.01) contribute both end points (ray order) followed by their l3d points.
 * lines3d [R,6], l3d [R,3] -> out [3R,3] (first *n_out rows valid; n_out: device int); score [R] or NULL.  R <= 4096. */
int neat_l3d_candidates(int R, const float* lines3d, const float* l3d, float* out, int* n_out, float* score,
                        void* stream);

/* ---- junction matching, HOST functions (no device work, no stream) -------------------------------------
 * The reference moves the clustered junctions to the CPU and solves two assignment problems with
 * scipy.optimize.linear_sum_assignment (neat_wfr_rend_a.py:466-484, loss_wfr.py:104-108).  These two entry
 * points are that host step in native code; all pointers are HOST pointers.
 *
 * neat_linear_sum_assignment: min-cost assignment of an n_rows x n_cols row-major cost matrix (shortest
 * augmenting paths, float64).  Writes min(n_rows, n_cols) pairs sorted by row (scipy's convention) and returns
 * their number, or a negative error code (NaN / -inf entries, infeasible problem).                          */
int neat_linear_sum_assignment(const double* cost, int n_rows, int n_cols, int* row_ind, int* col_ind);

/* neat_junction_match: centroids [C,3] (neat_dbscan output), gt_vertices [J,2] (WireframeGraph.vertices, pixels),
 * pose [16] (camera-to-world), intrinsics [16] (4x4), global_junctions [G,3] (ffn(latents)).
 *   1. project the centroids (pixel and calibrated coordinates), cost[j,c] = ||proj_c - gt_j||_2, assignment,
 *      keep pairs with cost < 10 px (use_median: < the median of the matched costs, 10 if there are none);
 *      local_out [min(J,C),7] rows = (x y z | u v | u_calib v_calib) in ground-truth order, *n_local rows valid
 *   2. the loss' assignment of the kept junctions to the global ones: cost = L1(3D) + 0.1 L1(calibrated 2D);
 *      global_rows / global_cols [n_local]; *n_close = pairs with cost < 10 (the loss' "jcount")            */
int neat_junction_match(const float* centroids, int n_centroids, const float* gt_vertices, int n_gt,
                        const float* pose, const float* intrinsics, const float* global_junctions, int n_global,
                        int use_median, float* local_out, int* n_local, int* global_rows, int* global_cols,
                        int* n_close, float* median_out);

/* ---- junction terms on the device ------------------------------------------------------------------------
 * neat_project_points: out_pix = project2D(K, R, T, X) and out_calib = project2D(I, R, T, X) of N points
 * (the global junctions, neat_wfr_rend_a.py:484-486); pose_inv [16] = world-to-camera (neat_line_geometry's output),
 * K: 3x3 with row stride k_ld.  Either output may be NULL.  _backward: g_X [N,3] = adjoint of both (either g NULL). */
int neat_project_points(int N, const float* pose_inv, const float* K, int k_ld, const float* X, float* out_pix,
                        float* out_calib, void* stream);
int neat_project_points_backward(int N, const float* pose_inv, const float* K, int k_ld, const float* X,
                                 const float* g_pix, const float* g_calib, float* g_X, void* stream);
/* neat_junction_terms (loss_wfr.py:110-121): for the n matched pairs (rows[i] -> local junction, cols[i] -> global
 * junction): out[0] = mean L1 3D distance (j3d_loss), out[1] = mean L1 calibrated 2D distance (j2d_loss), out[2] = the
 * same in pixels (statistics).  _backward: g_out [2] (device) -> gradients w.r.t. the GLOBAL junctions [n_global,3] /
 * their calibrated projections [n_global,2] (zero for unmatched rows).                                        */
int neat_junction_terms(int n, const float* j3d_local, const float* j3d_global, const float* j2d_local_calib,
                        const float* j2d_global_calib, const float* j2d_local, const float* j2d_global, const int* rows,
                        const int* cols, float* out, void* stream);
int neat_junction_terms_backward(int n, int n_global, const float* j3d_local, const float* j3d_global,
                                 const float* j2d_local_calib, const float* j2d_global_calib, const int* rows,
                                 const int* cols, const float* g_out, float* g_j3d_global, float* g_j2d_global_calib,
                                 void* stream);

/* ---- VolSDFLoss (code/model/networks/loss_wfr.py:34-79), forward fused with its own backward ---- */
/* loss_core = rgb L1 + eikonal_weight * eikonal + line_weight * calibrated line loss.  out[8] = {loss_core, rgb_loss,
 * eikonal_loss, line_loss, l2d_loss (uncalibrated, statistics only), count}.  g_* = d loss_core / d input.
 * lines_gt [R,5] = x1 y1 x2 y2 weight; labels [R] or NULL; K3: 3x3 with row stride k_ld; grad_theta may be NULL.
 * scratch: 8 + R floats.  The Hungarian-matched junction terms (loss_wfr.py:95-131) stay on the host.      */
typedef struct {
  int R, n_eik;
  const float *rgb_values, *rgb_gt, *lines2d, *lines2d_calib, *lines_gt, *labels, *K3;
  int k_ld;
  const float* grad_theta;
  float eikonal_weight, line_weight;
  float *scratch, *out, *g_rgb, *g_calib, *g_theta;
} neat_loss_args;
int neat_loss_forward_backward(const neat_loss_args* a, void* stream);
/* adjoint of lines2d_calib = project2D(I, R, T, lines3d) (neat_wfr_rend_a.py:442): g_calib [R,2,2] -> g_lines3d [R,2,3] */
int neat_project_calib_backward(int R, const float* pose_inv, const float* lines3d, const float* g_calib,
                                float* g_lines3d, void* stream);

/* ---- dataset-side attraction precompute (once per image) ------------------------ */
/* hawp.base._C.encodels (third-party/hawp/hawp/base/csrc/linesegment.cu:23-139): lines [num,4] (x1,y1,x2,y2) ->
 * map [6,H,W] f32, label [num,H,W] bool (bytes; zero-initialised here), tmap [1,H,W] f32.            */
int neat_encodels(const float* lines, int input_height, int input_width, int height, int width, int num_lines,
                  float* map, uint8_t* label, float* tmap, void* stream);
/* SceneDataset.compute_point_line_attraction (code/datasets/scene_hawp_dataset.py:92-146), fused:
 * mask [H*W] bool, labels [H*W] int64, proj_points [H*W,2] f32.                                   */
int neat_point_line_attraction(const float* lines, int num_lines, int height, int width, float distance,
                               uint8_t* mask, long long* labels, float* proj_points, void* stream);

/* ---- dataset pixel sampling: SceneDataset.__getitem__, code/datasets/scene_hawp_dataset.py:148-194 ----
 * The per-image tables stay on the device; one launch per step produces every per-ray input.
 * neat_mask_compact: `mask.nonzero()` (:173), once per image: out_idx [<= n] = ascending indices of the non-zero bytes of
 * mask [n], *n_out (device int) = how many.
 * neat_sample_pixels: R rays.  masked != NULL: pixel = masked[pos_j] with pos_j = perm[j] when perm != NULL (the prefix
 * of the reference's CPU `torch.randperm(n_masked)`, :176 -- same seed, same rays) or else the j-th value of a keyed
 * bijection of [0, n_masked) (seed, step): R distinct positions with no n-sized work (neat_pixel_permutation evaluates
 * the same bijection on the host).  masked == NULL: pixel = first + j (the full-image branch, sampling_idx None).
 * Outputs: uv [R,2] = (column, row) (:149-151), uv_proj [R,2] = att_points[pixel], rgb [R,3], labels_out [R],
 * lines2d [R,5] = lines[labels[pixel]] (NULL to skip), index_out [R] = pixel (NULL to skip).                        */
size_t neat_mask_compact_workspace_bytes(long long n);
int neat_mask_compact(const uint8_t* mask, long long n, void* workspace, int* out_idx, int* n_out, void* stream);
typedef struct {
  int R, W;
  long long first;
  const int* masked;
  int n_masked;
  const long long* perm;
  unsigned long long seed, step;
  const float* rgb_image;    /* [HW,3] */
  const long long* labels;   /* [HW]   */
  const float* att_points;   /* [HW,2] */
  const float* lines;        /* [n_lines,5] */
  int n_lines;
  float *uv, *uv_proj, *rgb, *lines2d;
  long long *labels_out, *index_out;
} neat_pixel_args;
int neat_sample_pixels(const neat_pixel_args* a, void* stream);
/* host function: out[i] = position drawn for ray first + i, i < count, of the (n, seed, step) bijection */
int neat_pixel_permutation(unsigned n, unsigned long long seed, unsigned long long step, unsigned first, unsigned count,
                           unsigned* out);

/* ---- backward (replaces loss.backward() through the model, code/training/volsdf_train.py:373) ---- */
/* Adjoint of neat_composite_forward for the outputs the reference losses consume (rgb_values, lines3d;
 * lines3d uses detached weights, neat_wfr_rend_a.py:410).  rgb_pre_bar [R,S,3] = dL/d(pre-sigmoid rgb),
 * lines_bar [R,S,6], sdf_bar [R,S] (masked by act), beta_bar[1] += dL/d(density.beta).            */
typedef struct {
  int R, S;
  const float *z, *sdf, *weights, *rgb, *act, *rgb_values_bar, *lines3d_bar, *beta_param;
  float beta_min;
  float *rgb_pre_bar, *lines_bar, *sdf_bar, *beta_bar;
  const float* bg_color;   /* [3] or NULL, as in neat_composite_args */
} neat_composite_bwd_args;
int neat_composite_backward(const neat_composite_bwd_args* a, void* stream);

/* Reverse sweep of a head.  out_bar [M,3|6] = dL/d(pre-activation output); fwd_save = the record written by
 * neat_head_forward(training=1); feat_bar (neat_feat_bar_bytes) and n_bar [M,3] are overwritten
 * (accumulate=0) or added to (accumulate=1); bwd_save feeds neat_weight_gradients.                */
size_t neat_head_bwd_save_bytes(const neat_ctx* ctx, int M);
size_t neat_feat_bar_bytes(int M);
int neat_head_backward(neat_ctx* ctx, int head, int M, const float* out_bar, const void* fwd_save, void* bwd_save,
                       float* feat_bar, float* n_bar, int accumulate, void* stream);

/* ImplicitNetwork double backward: n_bar [M,3] = dL/d(normal), s_bar [M] = dL/d(sdf)
 * (NULL = 0), feat_bar (NULL = 0), act [M] from neat_sdf_outputs (NULL = 1, eikonal points).        */
size_t neat_sdf_bwd_save_bytes(const neat_ctx* ctx, int M);
size_t neat_sdf_bwd_scratch_bytes(const neat_ctx* ctx, int M);
int neat_sdf_backward(neat_ctx* ctx, const neat_points* pts, const float* n_bar, const float* s_bar,
                      const float* feat_bar, const float* act, const void* fwd_save, void* bwd_save,
                      void* scratch, void* stream);

/* Weight / bias gradients of all three MLPs as tensor-core GEMMs over the saved operand tiles; ADDS into
 * flat_grad (layout of neat_param_count).  Any group may be absent (M = 0).                      */
typedef struct {
  int M;                     /* points of this group                                        */
  const void* sdf_fwd_save;  /* neat_sdf_outputs(training=1) record                          */
  const void* sdf_bwd_save;  /* neat_sdf_backward record                                     */
  const void* feat_tiles;    /* render points only (NULL for eikonal points)                */
  const void* head_fwd_save[2];
  const void* head_bwd_save[2];
} neat_grad_group;
int neat_weight_gradients(neat_ctx* ctx, const neat_grad_group* groups, int n_groups, float* flat_grad,
                          void* stream);

/* ---- wireframe finalisation: per-image line voting, code/neat-final-parsing.py:226-260 ----
 * lines2d [N,4], lines3d [N,2,3], points3d [N,3] (the `l3d` support points) are the eval-mode outputs of one image;
 * gt_lines [G,4] the image's 2D wireframe segments.  Every prediction votes twice (both end-point orders) for its nearest
 * ground-truth line; votes with squared distance < dis_threshold are kept.  Outputs per ground-truth line g:
 * counts[g] (votes), lines3d_mean[g] [2,3] (mean of the voting 3D lines, oriented like the vote), scores[g] (mean distance
 * of the votes' support points to that mean line); rows without votes are zero.                                  */
size_t neat_line_vote_workspace_bytes(int N, int G);
int neat_line_vote(const float* lines2d, const float* lines3d, const float* points3d, int N, const float* gt_lines, int G,
                   float dis_threshold, void* workspace, float* lines3d_mean, float* scores, float* counts, void* stream);

/* visibility_checking for ONE view (code/neat-final-parsing.py:305-337): lines3d [L,2,3] projected with project2D(K, R, T)
 * (pose_inv [16] = world-to-camera, K 3x3 with row stride k_ld); visible[l] is SET to 1 when the squared distance to the
 * nearest ground-truth 2D line gt_lines [G,4], in either end-point order, is < dis_threshold (never cleared: call once per
 * view on the same buffer to accumulate "seen in any view", or on a per-view column).  mindis [L] optional.         */
int neat_line_visibility(const float* lines3d, int L, const float* pose_inv, const float* K, int k_ld, const float* gt_lines,
                         int G, float dis_threshold, uint8_t* visible, float* mindis, void* stream);

/* get_wireframe_from_lines_and_junctions (code/neat-final-parsing.py:128-157): lines3d [N,2,3], junctions [J,3].
 * midx [N,2] = nearest junction of each end point, matched [N] = 1 when max(snap distances) < line length (and
 * rel_threshold <= 0: with a positive threshold the reference's `is_matched *= is_matched < thr` clears every match),
 * graph [J,J] f32 = symmetric adjacency (overwritten), upper [J,J] u8 = its upper triangle incl. the diagonal
 * (graph.triu() != 0; neat_mask_compact of it lists the wireframe's junction pairs in the reference's order).      */
int neat_line_junction_graph(const float* lines3d, int N, const float* junctions, int J, float rel_threshold, int* midx,
                             uint8_t* matched, float* graph, uint8_t* upper, void* stream);

/* ---- optimizer step: torch.optim.Adam(lr) of code/training/volsdf_train.py:178,374 ----
 * One launch for every parameter tensor: param -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * with m, v updated in place (torch's default Adam: no amsgrad, L2 weight decay added to the gradient).  grad is
 * multiplied by grad_scale first (1 / world size after the data-parallel all-reduce(SUM); 1 otherwise).
 * step = 1 for the first update.  At most 128 tensors per call.                                                */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
} neat_adam_tensor;
int neat_adam_step(const neat_adam_tensor* tensors, int n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, float grad_scale, void* stream);

/* ---- the training step without the host in it (round 2): everything below exists so that one step can be replayed as
 * CUDA graphs (neat_b200.trainer.FusedTrainStep): no eager PyTorch kernels, no host-side counters. -------------------- */
/* Every random draw of one training forward in ONE launch (Philox4x32-10 keyed by seed and a device-resident draw counter
 * that the kernel bumps itself: counter_dev = 2 x uint64, zero-initialised).  Same DISTRIBUTIONS as the reference's CPU
 * draws: t_rand [R,n_eval] ~ U[0,1) (code/model/ray_sampler.py:87), u_final [R,n_final] ~ U[0,1) (:234), extra_idx
 * [max_iters,n_extra] row k-1 = the first n_extra entries of a uniform random permutation of [0, n_eval*k) (:265),
 * eik_idx [R] ~ U{0..n_final+2+n_extra-1} (:275), eik_uniform [R,3] ~ U(-radius, radius)
 * (code/model/networks/neat_wfr_rend_a.py:518).                                                                    */
int neat_train_draws(const neat_sampler_config* cfg, int R, float radius, unsigned long long seed,
                     unsigned long long* counter_dev, float* t_rand, float* u_final, int64_t* extra_idx, int64_t* eik_idx,
                     float* eik_uniform, void* stream);
/* fp32 GEMM C[M,N] (+)= op(A) op(B) for the junction `ffn` (nn.Sequential of 3 Linear layers on the 1024 latents,
 * neat_wfr_rend_a.py:274-303, 488) and its backward: row-major with leading dimensions; ta: A stored [K,M]; tb: B stored
 * [N,K] (a Linear weight); optional bias [N], ReLU, ReLU-adjoint mask (C = mask > 0 ? C : 0), accumulate.          */
int neat_gemm_f32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int ta, int tb,
                  const float* bias, int relu, const float* mask, int ldm, int accumulate, void* stream);
/* out[n] (+)= sum_m X[m,n] (bias gradients) */
int neat_colsum_f32(const float* X, int M, int N, int ldx, float* out, int accumulate, void* stream);
/* neat_junction_terms + _backward with the pair count read from DEVICE memory (it changes every step): packed = [n | rows
 * [cap] | cols [cap] (int32) | local [cap,7] (float: xyz, uv, uv_calib)], the one host->device copy of the step.  out[3] =
 * j3d_loss, j2d_loss, j2d_stat; g_* = d (w3 out[0] + w2 out[1]) / d (global junctions [G,3], calibrated projections [G,2]). */
int neat_junction_step(const int* packed_dev, int cap, int n_global, const float* j3d_global, const float* j2d_global_calib,
                       const float* j2d_global, float w3, float w2, float* out, float* g_j3d_global,
                       float* g_j2d_global_calib, void* stream);
/* neat_adam_step with the step count and hyper-parameters in device memory: hyper_dev[6] = lr, beta1, beta2, eps,
 * weight_decay, grad_scale; state_dev[3] = step count (bumped here), 1 - beta1^step, 1/sqrt(1 - beta2^step).        */
int neat_adam_step_device(const neat_adam_tensor* tensors, int n, const float* hyper_dev, float* state_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEAT_B200_H */
