/*
 * nlb200.h -- C ABI of libnlb200.so, the B200 (sm_100a) implementation of the
 * NeRF-LiDAR zipnerf volume-rendering hot path.
 *
 * Every entry point takes plain DEVICE pointers and sizes (no torch types), is
 * asynchronous on the CUDA stream passed as `stream` (a cudaStream_t; NULL = the
 * legacy default stream, which is what the reference's kernels use), and returns
 * 0 on success or a negative NLB_E* code; nlb_last_error() gives the message of
 * the last failure on the calling thread.  Callers allocate all outputs, as in the
 * reference (gridencoder/grid.py:47-52,77-82).
 *
 * `Z/` = NeRF_LiDAR/zipnerf/ in the reference tree.  Each function names the
 * reference interface it replaces.
 */
#ifndef NLB200_H
#define NLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLB_OK 0
#define NLB_EINVAL (-1)      /* bad argument (the reference raises RuntimeError) */
#define NLB_EUNSUPPORTED (-2)/* valid in the reference, not built here (D>3, fp16/fp64 tables) */
#define NLB_ECUDA (-3)       /* CUDA launch error */

const char* nlb_last_error(void);
int nlb_version(void);
/* 1 when the library was compiled for sm_100a and a device of that arch is current. */
int nlb_device_ok(void);

/* ------------------------------------------------------------------ grid encoder
 * Replaces the pybind11 module `_gridencoder` (Z/gridencoder/src/bindings.cpp:5-9,
 * Z/gridencoder/src/gridencoder.h:12-15).  Argument order and meaning follow
 * grid_encode_forward / grid_encode_backward / grad_total_variation there.
 * fp32 tables; D in {2,3}; C in {1,2,4,8}.
 */
int nlb_grid_encode_forward(const float* inputs /*[B,D] in [0,1]*/, const float* embeddings /*[rows,C]*/,
                            const int32_t* offsets /*[L+1]*/, float* outputs /*[L,B,C]*/,
                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            float* dy_dx /*[B,L*D*C] or NULL*/, uint32_t gridtype, int align_corners,
                            uint32_t interp, void* stream);

int nlb_grid_encode_backward(const float* grad /*[L,B,C]*/, const float* inputs, const float* embeddings,
                             const int32_t* offsets, float* grad_embeddings /*[rows,C], pre-zeroed, accumulated*/,
                             uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             const float* dy_dx /*or NULL*/, float* grad_inputs /*[B,D] or NULL*/,
                             uint32_t gridtype, int align_corners, uint32_t interp, void* stream);

int nlb_grad_total_variation(const float* inputs, const float* embeddings, float* grad,
                             const int32_t* offsets, float weight, uint32_t B, uint32_t D, uint32_t C,
                             uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                             void* stream);

/* Parity probe: the 2^D level-local row indices kernel_grid would gather
 * (Z/gridencoder/src/gridencoder.cu:66-84,166-179).  indices[L,B,2^D];
 * out-of-range points get 0xFFFFFFFF. */
int nlb_grid_corner_indices(const float* inputs, const int32_t* offsets, uint32_t* indices,
                            uint32_t B, uint32_t D, uint32_t L, float S, uint32_t H,
                            uint32_t gridtype, int align_corners, void* stream);

/* ------------------------------------------------------------------ step-function resampling
 * One call = the per-level chain of Model.forward, Z/internal/models.py:320-372:
 * stepfun.max_dilate_weights (+[1:-1] trim) when `dilate`, anneal/log -> logits,
 * softmax, integrate_weights, math.sorted_interp at the sample centres,
 * midpoints + reflected end fenceposts (stepfun.sample_intervals), and the
 * power-transformation s_to_t (Z/internal/coord.py:103-162).
 *   sdist_in[N,n_in+1], weights_in[N,n_in]; u_base[S] = the reference's
 *   torch.linspace; jitter[N] (U[0,1), single_jitter) or NULL for rand=False.
 *   sample_idx[N,S] (optional) = lower CDF knot of every sample centre.
 */
int nlb_resample(const float* sdist_in, const float* weights_in, int n_in, int dilate, float dilation,
                 float anneal, float resample_padding, const float* u_base, const float* jitter,
                 float max_jitter, const float* near, const float* far, float lam, int S, int N,
                 float* sdist_out /*[N,S+1]*/, float* tdist_out /*[N,S+1]*/, int32_t* sample_idx,
                 void* stream);

/* math.sorted_interp (Z/internal/math.py:89-108) for sorted xp / non-decreasing fp.
 * x[N,nx], xp[N,np], fp[N,np] -> out[N,nx], idx[N,nx] (optional lower-knot index). */
int nlb_sorted_interp(const float* x, const float* xp, const float* fp, int N, int nx, int np,
                      float* out, int32_t* idx, void* stream);

/* ------------------------------------------------------------------ fused sample-point generation + encoding
 * render.cast_rays (Z/internal/render.py:129-168) + coord.contract_mean_std and /2
 * (Z/internal/coord.py:51-63, models.py:968-973) + GridEncoder forward + erf
 * re-weighting and mean over the n=7 multisamples (models.py:974-977), without
 * materialising means/stds or the [N,S,7,L,C] features.
 *   features[N*S, L*C] fp32.  deg_noise[N,S,7] or NULL (rand=False).
 */
typedef struct {
  const float* tdist;     /* [N,S+1] */
  const float* origins;   /* [N,3] */
  const float* directions;/* [N,3] */
  const float* radii;     /* [N]   */
  const float* base_x;    /* [N,3] */
  const float* base_y;    /* [N,3] */
  const float* deg_noise; /* [N,S,7] or NULL */
  int N, S;
  float std_scale;        /* Model.std_scale = 0.35 */
  float* points_cache;    /* [7, N*S, 4] grid-space sample points (x, y, z, std or -1), or NULL */
  int points_mode;        /* 0: compute; 1: compute and write points_cache (training forward);
                             2: read points_cache instead of computing (backward of the same level) */
} nlb_rays_t;

typedef struct {
  const float* embeddings;   /* [rows,C] */
  const int32_t* offsets;    /* [L+1] */
  const int32_t* grid_sizes; /* [L] (= python resolutions, scale+2) */
  int L, C;
  uint32_t H;                /* base resolution */
  float S;                   /* log2(per_level_scale) */
  const int32_t* offsets_host; /* [L+1] HOST copy of `offsets` (launch shapes, validation: hashed levels
                                  must be powers of two, as GridEncoder builds them) */
} nlb_table_t;

/* Parity probe: the sample points the fused kernels generate, points[N,S,7,4] =
 * (x,y,z in the [0,1] grid cube, contracted std / 2). */
int nlb_sample_points(const nlb_rays_t* rays, float* points, void* stream);

int nlb_encode_forward(const nlb_rays_t* rays, const nlb_table_t* table, float* features, void* stream);
/* grad_features[N*S, L*C] -> grad_embeddings[rows,C] (accumulated). */
int nlb_encode_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* grad_features,
                        float* grad_embeddings,
                        float* workspace /*nlb_encode_backward_workspace_bytes(), or NULL: no privatised coarse rows*/,
                        void* stream);
size_t nlb_encode_backward_workspace_bytes(const nlb_table_t* table);

/* Gradients w.r.t. the ray geometry, needed only while poses are refined (Z/train.py:200-221: origins,
 * directions, base_x, base_y become functions of LearnPose).  Replaces, for the fused path, the reference's
 * dy_dx output + kernel_input_backward (gridencoder.cu:201-244,343-369) and the autograd chain through the erf
 * weights (models.py:974-977), contract_mean_std (coord.py:51-63) and cast_rays (render.py:129-168):
 * grad_features[N*S, L*C] -> four [N,3] buffers, ACCUMULATED (pre-zero them; the three levels add into one set).
 * tdist carries no gradient (models.py:368-369).  The sample points are regenerated from the rays
 * (points_cache is not read). */
typedef struct {
  float* origins;    /* [N,3] */
  float* directions; /* [N,3] */
  float* base_x;     /* [N,3] */
  float* base_y;     /* [N,3] */
} nlb_ray_grads_t;
int nlb_encode_input_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* grad_features,
                              const nlb_ray_grads_t* grads, void* stream);

/* Proposal level: the above + PropMLP Linear(L,64)-ReLU-Linear(64,1), softplus(x-1)
 * (models.py:887-889,996-997,1116) in one kernel.  W0[64,L] b0[64] W1[64] b1[1]
 * in nn.Linear layout.  density[N,S]. */
int nlb_prop_forward(const nlb_rays_t* rays, const nlb_table_t* table, const float* W0, const float* b0,
                     const float* W1, const float* b1, float* density,
                     float* features /*[N*S,L] saved for the backward, or NULL*/, void* stream);
/* grad_density[N,S] + the features saved by the forward -> table / weight grads
 * (all accumulated into pre-zeroed buffers). */
int nlb_prop_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* W0, const float* b0,
                      const float* W1, const float* b1, const float* features, const float* grad_density,
                      float* grad_embeddings, float* gW0, float* gb0, float* gW1, float* gb1,
                      float* workspace /*nlb_prop_backward_workspace_bytes()*/, void* stream);
size_t nlb_prop_backward_workspace_bytes(int N, int S, const nlb_table_t* table);
/* Proposal level of the above: reads the feature gradients nlb_prop_backward left in `workspace` (call it right
 * after, same rays / table / stream) and accumulates the ray-geometry gradients. */
int nlb_prop_input_backward(const nlb_rays_t* rays, const nlb_table_t* table, const float* workspace,
                            const nlb_ray_grads_t* grads, void* stream);

/* ------------------------------------------------------------------ compositing
 * render.compute_alpha_weights (Z/internal/render.py:170-189) +
 * render.volumetric_rendering (:192-284) incl. stepfun.weighted_percentile
 * (Z/internal/stepfun.py:329-339).  NULL inputs/outputs are skipped.
 */
typedef struct {
  const float* density;   /* [N,S] */
  const float* tdist;     /* [N,S+1] */
  const float* directions;/* [N,3] */
  const float* rgb;       /* [N,S,3] or NULL (proposal levels: zeros) */
  const float* semantic;  /* [N,S,K] or NULL */
  const float* intensity; /* [N,S] or NULL */
  const float* far;       /* [N] */
  int N, S, K;
  float bg;               /* background colour (bg_intensity_range = (1,1)) */
  int opaque_background;
  int compute_extras;
} nlb_composite_in_t;

typedef struct {
  float* weights;   /* [N,S] */
  float* rgb;       /* [N,3] */
  float* depth;     /* [N] */
  float* acc;       /* [N] */
  float* semantic;  /* [N,K] or NULL */
  float* intensity; /* [N] or NULL */
  float* distance_mean;       /* [N] or NULL */
  float* distance_percentiles;/* [N,3] (5,50,95) or NULL */
} nlb_composite_out_t;

int nlb_composite_forward(const nlb_composite_in_t* in, const nlb_composite_out_t* out, void* stream);

typedef struct {
  const float* g_weights;  /* [N,S] direct gradient on the weights (losses on ray_history) or NULL */
  const float* g_rgb;      /* [N,3] or NULL */
  const float* g_depth;    /* [N] or NULL */
  const float* g_acc;      /* [N] or NULL */
  const float* g_semantic; /* [N,K] or NULL (weights detached: reaches only `semantic`) */
  const float* g_intensity;/* [N] or NULL (weights detached) */
} nlb_composite_grad_t;

int nlb_composite_backward(const nlb_composite_in_t* in, const float* weights /*[N,S] from forward*/,
                           const nlb_composite_grad_t* g, float* g_density /*[N,S]*/,
                           float* g_rgb /*[N,S,3] or NULL*/, float* g_semantic /*[N,S,K] or NULL*/,
                           float* g_intensity /*[N,S] or NULL*/, void* stream);

/* ------------------------------------------------------------------ NeRF MLP (tcgen05 / TMEM)
 * MLP.forward for NerfMLP, Z/internal/models.py:996-997,1116-1251:
 * 40->64->256 trunk, density softplus, semantic 256->64->19 softmax, intensity
 * 256->64->1, view branch cat[x,pos_enc(viewdirs)](283)->256->cat(539)->256->3
 * sigmoid with rgb padding.  bf16 operands, fp32 accumulation in TMEM.
 * Weights are passed as one packed bf16 blob built by nlb_nerf_mlp_pack().
 */
size_t nlb_nerf_mlp_packed_bytes(void);
typedef struct {
  const float *W_d0, *b_d0;   /* [64,40],[64]   density_layer.0 */
  const float *W_d2, *b_d2;   /* [256,64],[256] density_layer.2 */
  const float *W_s0, *b_s0;   /* [64,256],[64]  sem_layer.0 */
  const float *W_s2, *b_s2;   /* [19,64],[19]   sem_layer.2 */
  const float *W_i0, *b_i0;   /* [64,256],[64]  intensity_layer.0 */
  const float *W_i2, *b_i2;   /* [1,64],[1]     intensity_layer.2 */
  const float *W_v0, *b_v0;   /* [256,283],[256] lin_second_stage_0 */
  const float *W_v1, *b_v1;   /* [256,539],[256] lin_second_stage_1 */
  const float *W_rgb, *b_rgb; /* [3,256],[3]    rgb_layer */
} nlb_nerf_mlp_weights_t;
int nlb_nerf_mlp_pack(const nlb_nerf_mlp_weights_t* w, void* packed, void* stream);
/* Row-major bf16 activations the forward can save for the backward pass
 * (post-ReLU hidden states; x = bottleneck before the heads).  All-NULL = inference. */
typedef struct {
  void* h0; /* [M,64]  relu(density_layer.0) */
  void* x;  /* [M,256] density_layer.2 output */
  void* g;  /* [M,128] relu(sem_layer.0) | relu(intensity_layer.0) */
  void* h1; /* [M,256] relu(lin_second_stage_0) */
  void* h2; /* [M,256] relu(lin_second_stage_1) */
  void* f0; /* [M,64]  the input features, zero-padded from 40 columns (optional: nlb_nerf_mlp_wgrad needs it) */
} nlb_nerf_mlp_saved_t;
int nlb_nerf_mlp_forward(const float* features /*[M,40]*/, const float* viewdirs /*[N,3]*/, int M,
                         int rows_per_ray, const void* packed, float* density /*[M]*/, float* rgb /*[M,3]*/,
                         float* semantic /*[M,19]*/, float* intensity /*[M]*/,
                         const nlb_nerf_mlp_saved_t* saved /*or NULL*/, void* stream);

/* Data-gradient chain of the NerfMLP on tcgen05 (transposed packed weights from
 * nlb_nerf_mlp_pack_transposed): from the output gradients and the activations saved
 * by the forward it produces grad_features[M,40] (fp32) and the pre-activation
 * gradients of every layer as row-major bf16 (inputs of the weight-gradient GEMMs
 * dW = dZ^T A, which are plain GEMMs left to the caller). */
size_t nlb_nerf_mlp_packed_transposed_bytes(void);
int nlb_nerf_mlp_pack_transposed(const nlb_nerf_mlp_weights_t* w, void* packed_t, void* stream);
typedef struct {
  const float* g_density;   /* [M] or NULL */
  const float* g_rgb;       /* [M,3] or NULL */
  const float* g_semantic;  /* [M,19] or NULL */
  const float* g_intensity; /* [M] or NULL */
  const float* density;     /* [M]    forward outputs */
  const float* rgb;         /* [M,3]  */
  const float* semantic;    /* [M,19] */
} nlb_nerf_mlp_grad_in_t;
typedef struct {
  void* d_rgb; /* [M,16]  bf16: d(rgb pre-sigmoid) in cols 0..2 */
  void* d_v1;  /* [M,256] bf16: d(lin_second_stage_1 pre-relu) */
  void* d_v0;  /* [M,256] bf16: d(lin_second_stage_0 pre-relu) */
  void* d_hs1; /* [M,32]  bf16: d(sem logits) cols 0..18, d(intensity) col 19 */
  void* d_g;   /* [M,128] bf16: d(sem_layer.0 | intensity_layer.0 pre-relu) */
  void* d_x;   /* [M,256] bf16: d(bottleneck) */
  void* d_h0;  /* [M,64]  bf16: d(density_layer.0 pre-relu) */
  int ld_v1, ld_v0, ld_g; /* leading dimensions (elements) of d_v1 / d_v0 / d_g; 0 = dense (256 / 256 / 128).
                             The three share the A operand `x` in the weight-gradient GEMMs: written as
                             column slices of one [M,640] buffer they become ONE GEMM. */
} nlb_nerf_mlp_grad_out_t;
int nlb_nerf_mlp_backward(const nlb_nerf_mlp_grad_in_t* gin, const nlb_nerf_mlp_saved_t* saved, int M,
                          const void* packed_t, float* grad_features /*[M,40]*/,
                          const nlb_nerf_mlp_grad_out_t* gout, void* stream);

/* Weight gradients of the NerfMLP dense layers on tcgen05: dW += dZ^T A for every layer, with dZ = the bf16
 * pre-activation gradients written by nlb_nerf_mlp_backward and A = the bf16 activations saved by
 * nlb_nerf_mlp_forward (incl. f0), both read as MN-major operands straight from their row-major matrices.
 * The results are ADDED (fp32 atomics) into the gradient tensors, which have the shapes of the weights
 * (e.g. views of one flat gradient buffer that the optimizer pass clears).  nlb_nerf_mlp_wgrad covers the
 * columns fed by activations; nlb_nerf_mlp_wgrad_finish adds the bias gradients (cs_* = column sums of d_x[256],
 * d_g[128], d_h0[64], d_hs1[32], d_rgb[16] from nlb_colsum_bf16) and the view-direction columns
 * W_v0[:, 256:283], W_v1[:, 512:539] and b_v0 / b_v1 from the per-ray sums rs_v0 / rs_v1 [N,256]
 * (nlb_group_sum_bf16) and viewdirs [N,3]. */
typedef struct {
  float *W_d0, *b_d0, *W_d2, *b_d2, *W_s0, *b_s0, *W_s2, *b_s2, *W_i0, *b_i0, *W_i2, *b_i2;
  float *W_v0, *b_v0, *W_v1, *b_v1, *W_rgb, *b_rgb;   /* same order and shapes as nlb_nerf_mlp_weights_t */
} nlb_nerf_mlp_wgrads_t;
int nlb_nerf_mlp_wgrad(const nlb_nerf_mlp_saved_t* saved, const nlb_nerf_mlp_grad_out_t* dz, int M,
                       const nlb_nerf_mlp_wgrads_t* grads, void* stream);
int nlb_nerf_mlp_wgrad_finish(const float* rs_v0, const float* rs_v1, const float* viewdirs, int N,
                              const float* cs_x, const float* cs_g, const float* cs_h0, const float* cs_hs1,
                              const float* cs_rgb, const nlb_nerf_mlp_wgrads_t* grads, void* stream);

/* Reductions of the bf16 pre-activation gradients written by nlb_nerf_mlp_backward:
 * colsum: x[M, ld] (first `cols` columns; cols a power of two <= 512) -> out[cols]
 *   (overwritten) = bias gradients;
 * group_sum: x[groups*S, ld] (first `cols` columns) -> out[groups, cols] = per-ray sums over the S samples (the
 *   view-direction encoding is a per-ray constant, Z/internal/models.py:1192-1196). */
int nlb_colsum_bf16(const void* x, int64_t M, int cols, int ld, float* out, void* stream);
int nlb_group_sum_bf16(const void* x, int64_t groups, int S, int cols, int ld, float* out, void* stream);

/* The same reductions as jobs of ONE launch (what a NerfMLP backward issues: five bias gradients and two per-ray
 * sums, 661 MB at the bench size): x bf16 [rows, ld] read with 16-byte loads (x and out 16-byte aligned, ld a
 * multiple of 8, cols a power of two in [8, 256]).  group = 0: out[cols] = column sums (overwritten); group = S > 0:
 * out[rows / S, cols] = sums over S consecutive rows.  At most 8 jobs. */
typedef struct {
  const void* x;
  int64_t rows;
  int cols, ld, group;
  float* out;
} nlb_bf16_sum_job_t;
int nlb_bf16_sums(const nlb_bf16_sum_job_t* jobs, int njobs, void* stream);

/* Dev probe: clock64() stamps of block 0's MMA thread / epilogue thread for the first two
 * tiles of nlb_nerf_mlp_forward are written to buf[128] (int64); NULL switches it off. */
int nlb_debug_set_timeline(void* buf);

/* ------------------------------------------------------------------ per-ray regularisers
 * Value and gradient (w.r.t. the weights only; distances are detached in the
 * reference) in one pass.
 * distortion: stepfun.lossfun_distortion (Z/internal/stepfun.py:297-307);
 *   loss_ray[N], grad_w[N,S] = d loss_ray / d weights.
 * interlevel: train_utils.anti_interlevel_loss for ONE proposal level
 *   (Z/internal/train_utils.py:134-172, stepfun.blur_stepfun, math.sorted_interp_quad);
 *   (c[N,Sc+1], w[N,Sc]) = final level, (cp[N,Sp+1], wp[N,Sp]) = proposal level;
 *   loss_ray[N] = sum over the Sp intervals, grad_wp[N,Sp] = d loss_ray / d wp.
 */
int nlb_distortion_loss(const float* sdist, const float* weights, int N, int S, float* loss_ray, float* grad_w,
                        void* stream);
int nlb_interlevel_loss(const float* c, const float* w, int Sc, const float* cp, const float* wp, int Sp,
                        float pulse_width, int N, float* loss_ray, float* grad_wp, void* stream);

/* Loss assembly: the reference sums its loss dictionary and back-propagates the sum (Z/train.py:283-462); as torch
 * scalar arithmetic that is ~45 tiny launches per step.
 * weighted_sums (one block): for every term in order, out[out_index] += coef * sum_i x[i] * (w ? w[i] : 1), i < n;
 *   a term with x == NULL adds coef * (the value this call has formed so far for output index n).  Outputs that
 *   no term names keep their contents, named ones are overwritten.  At most 24 terms, 16 outputs.
 * scale_tensors (one launch): dst[i] = src[i] * (coef * g * s) + (src2 ? src2[i] * (coef2 * g * s2) : 0), with
 *   g / s / s2 device scalars (NULL = 1): every gradient a backward pass seeds.  At most 8 jobs. */
typedef struct {
  const float* x;
  const float* w;
  int64_t n;
  float coef;
  int out_index;
} nlb_sum_term_t;
int nlb_weighted_sums(const nlb_sum_term_t* terms, int nterms, float* out, int nout, void* stream);
typedef struct {
  const float* src;
  const float* src2;
  float* dst;
  int64_t n;
  const float* g;
  const float* s;
  const float* s2;
  float coef, coef2;
} nlb_scale_job_t;
int nlb_scale_tensors(const nlb_scale_job_t* jobs, int njobs, void* stream);

/* ------------------------------------------------------------------ supervision losses
 * Value and gradient of the losses of Z/train.py:283-455 on the final rendering:
 * losses[6] / scales[6] = {data, depth, sem, int, d_smo, s_smo} (device).  The gradient
 * buffers receive the UNNORMALISED per-ray terms; d loss_k / d output = scales[k] * g_k
 * (scales = multiplier / denominator, 0 where the reference's nan_to_num zeroes a term):
 *   g_rgb[N,3] (data), g_depth[N] (depth), g_sem[N,K] (sem), g_int[N] (int),
 *   g_depth_smo[N], g_sem_smo[N,K] (first num_patch*patch_size^2 rays: the patch rays
 *   lead the batch, Z/internal/datasets.py:356-366; the rest is left untouched).
 * Masks follow Z/train.py:286-327: `ray_valid` (NULL = every ray) is 1 where the loss applies, i.e. where the
 * dataset's batch['mask'] is non-zero (Z/train.py:287,307: `mask = mask == 0`, `rgb_mask = mask == 0`); with
 * Config.instance_obj = True the reference clears the mask (NULL here).  It gates the rgb / depth / semantic
 * terms and, as edge_aware_loss_v2's `mask` (Z/internal/train_utils.py:341-348), the smoothness edges: an edge
 * counts when both of its pixels are valid, and each direction is normalised by its number of counted edges.
 * A patch_mask that does not lead the batch (== 1 exactly on the first num_patch*patch_size^2 rays) turns
 * d_smo / s_smo into NaN instead of silently smoothing the wrong rays.
 */
typedef struct {
  const float* rgb;         /* [N,3] rendered */
  const float* depth;       /* [N] */
  const float* semantic;    /* [N,K] class probabilities or NULL */
  const float* intensity;   /* [N] or NULL */
  const float* t_rgb;       /* [N,3] targets */
  const float* t_depth;     /* [N] (> 0: supervised) */
  const float* t_semantic;  /* [N] float labels, 255 = unlabelled */
  const float* t_intensity; /* [N] */
  const float* patch_mask;  /* [N] == 1: patch ray (smoothness only) */
  const float* lidar_mask;  /* [N] == 1: LiDAR ray */
  const float* ray_valid;   /* [N] != 0: supervised (dataset mask, see above) or NULL = all */
  int N, K;
  int num_patch, patch_size;
  int lidar_supervision, only_lidar_supervision;
  int charb;                /* 1: Charbonnier (Config.data_loss_type = 'charb'), 0: MSE */
  float charb_padding;
  float depth_mult, sem_mult, int_mult, smooth_mult; /* 0.4|0.1|0, 0.04|0.01|0, 0.1, 0.01 (Z/train.py:330-371) */
  float smo_scale_x, smo_scale_y; /* unused (kept for layout compatibility): the edge counts are data-dependent */
} nlb_losses_in_t;

size_t nlb_render_losses_workspace_bytes(void);
int nlb_render_losses(const nlb_losses_in_t* in, float* losses /*[6]*/, float* scales /*[6]*/, float* g_rgb,
                      float* g_depth, float* g_sem, float* g_int, float* g_depth_smo, float* g_sem_smo,
                      float* workspace, void* stream);

/* ------------------------------------------------------------------ optimizer
 * One fused pass per table: hash-decay gradient (Model.hash_decay_loss,
 * Z/internal/models.py:203-223: d/dp of mult * mean_levels(mean_rows(p^2)))
 * + NaN scrub (train_utils.py:251-253) + Adam (train_utils.py:256-275:
 * betas .9/.99, eps 1e-15) + zeroing the gradient for the next step.
 */
int nlb_adam_table_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                        const int32_t* offsets_host /*[L+1], HOST memory*/, int L, int C, float decay_mult, float lr, float beta1,
                        float beta2, float eps, int step, float grad_scale,
                        float* level_sumsq /*[L] += per-level sum of squares of the UPDATED table, or NULL*/,
                        void* stream);
/* The same pass over the float range [first, first + count) of the table only (both multiples of 4): the slice of
 * the optimizer pass one rank owns in data-parallel runs (reduce-scatter of the gradients -> this -> all-gather of
 * the parameters).  param / grad point at the START of the table, exp_avg / exp_avg_sq at the slice-local moments;
 * level_sumsq receives the slice's contribution. */
int nlb_adam_table_step_range(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                              const int32_t* offsets_host, int L, int C, float decay_mult, float lr, float beta1,
                              float beta2, float eps, int step, float grad_scale, float* level_sumsq,
                              int64_t first, int64_t count, void* stream);
/* CUDA-graph support.  The scalars that change every training step -- the resampling
 * anneal (Z/internal/models.py:343-349) and Adam's learning rate / bias corrections --
 * are by-value arguments above, which a captured graph would freeze.  When a device buffer
 * dyn[3] = {anneal, lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)} is registered here,
 * nlb_resample, nlb_adam_table_step and nlb_adam_step read these three from it at kernel
 * execution time instead (NULL restores the by-value behaviour).  nlb_adam_bias_terms
 * computes dyn[1], dyn[2] exactly as the by-value path does (host, double precision). */
int nlb_set_dynamic_scalars(const float* dev);
int nlb_adam_bias_terms(float lr, float beta1, float beta2, int step, float* out2 /*host [2]*/);

int nlb_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

/* ------------------------------------------------------------------ on-GPU ray generation
 * Replaces the host-side numpy ray construction of the reference's data layer (the 8-worker DataLoader,
 * Z/train.py:111-118).  float64 arithmetic in the reference's operation order, ONE rounding to float32 at the
 * store (the reference casts the finished batch with `.float()`, Z/internal/datasets.py `_make_ray_batch`).
 * Outputs follow the batch schema Model.forward reads (Z/internal/camera_utils.py:603-617): every pointer is
 * [n,3] fp32 except radii [n] and imageplane [n,2] (optional, may be NULL).
 */
typedef struct {
  float* origins; float* directions; float* viewdirs; float* radii; float* imageplane; float* base_x; float* base_y;
} nlb_ray_out_t;

/* camera_utils.pixels_to_rays / cast_ray_batch (Z/internal/camera_utils.py:454-564,567-617) for a perspective
 * camera without distortion or NDC (the nuScenes loader).  pixtocams [n_pixtocams,3,3] and camtoworlds
 * [n_camtoworlds,3,4] are row-major float64; a count of 1 shares the matrix between all rays (`batch_index`,
 * camera_utils.py:591), otherwise cam_idx[n] selects it. */
int nlb_camera_rays(const int32_t* pix_x, const int32_t* pix_y, const int32_t* cam_idx /*or NULL*/,
                    const double* pixtocams, int n_pixtocams, const double* camtoworlds, int n_camtoworlds,
                    int64_t n, const nlb_ray_out_t* out, void* stream);

/* lidar_utils.get_directions (Z/internal/lidar_utils.py:559-568): out[n_beams*width,3] fp32, beam-major,
 * (cos t sin p, cos t cos p, sin t) with t = elevation in DEGREES, p = azimuth in radians. */
int nlb_lidar_directions(const double* elev_deg, int n_beams, const double* azim_rad, int width, float* out,
                         void* stream);

/* lidar_utils.cast_lidar_ray_batch (Z/internal/lidar_utils.py:8-33) incl. its quirks: viewdirs = directions /
 * the GLOBAL Frobenius norm of the whole [n,3] array, base_x = base_y = directions, radii = 5e-4.
 * workspace: 8 bytes of device memory (the sum of squares). */
int nlb_lidar_rays(const float* origins, const float* directions, int64_t n, double* workspace,
                   const nlb_ray_out_t* out, void* stream);

/* ------------------------------------------------------------------ dynamic-object branch
 * Replaces, inside Model.forward (Z/internal/models.py:306-315,401-477): obj_utils.get_pose
 * (Z/internal/obj_utils.py:431-475), obj_utils.box_pts / world2object / rotate_yaw_z / scale_frames
 * (:196-234,116-194,76-111,5-29), the boolean-index compaction with its host synchronisation
 * (`intersect_idx.sum() == 0`), the per-class ObjMLP evaluation (Z/internal/models.py:1000-1034,1036-1263 with
 * warp_fn=None, re_weights=False, fixed_semantic=True, split shape / texture latent) and the masked merge of
 * every result key (zeros_like + masked assignment + where).
 */
/* pose[N,n_obj,9] = per ray, per track: the two track entries closest to the ray's timestamp, blended
 * (get_pose).  tracks[n_obj,T,9] = centre(3), yaw about z, wlh(3), timestamp, track id; T >= 2. */
int nlb_obj_pose(const float* time /*[N]*/, const float* tracks, int N, int n_obj, int T, float* pose, void* stream);

typedef struct {
  const float *W_d0, *b_d0;   /* density_layer.0 [hidden, L*C + latent_shape] */
  const float *W_d2, *b_d2;   /* density_layer.2 [bottleneck, hidden] */
  const float *W_v0, *b_v0;   /* lin_second_stage_0 [view_width, bottleneck + dir + latent_tex] */
  const float *W_v1, *b_v1;   /* lin_second_stage_1 [view_width, view_width + bottleneck + dir + latent_tex] */
  const float *W_rgb, *b_rgb; /* rgb_layer [3, view_width] */
  const float* latent;        /* [latent_shape + latent_tex] (split_latent) or NULL */
  int hidden, bottleneck, view_width, deg_view, latent_shape, latent_tex;
  float density_bias, rgb_premultiplier, rgb_bias, rgb_padding;
  int class_type, class_num;  /* fixed_semantic one-hot; class 255 = all zeros */
} nlb_obj_mlp_t;

/* One track at one sampling level: every sample midpoint of every ray is tested against the track's box; hits
 * are evaluated by the ObjMLP and OVERWRITE density[N,S] (and rgb[N,S,3], semantic[N,S,class_num] when given:
 * NULL at the proposal levels) in place; obj_mask[N,S] (bytes) is set to 1 on hits.  Tracks are applied by
 * calling this once per track in track order, as the reference's loop does. */
int nlb_obj_forward(const float* tdist /*[N,S+1]*/, const float* origins, const float* directions,
                    const float* viewdirs, const float* pose /*[N,n_obj,9]*/, int n_obj, int track, int N, int S,
                    const nlb_table_t* table, const nlb_obj_mlp_t* mlp, float* density, float* rgb, float* semantic,
                    uint8_t* obj_mask, int32_t* owner /*[N,S] or NULL: the track that wrote each sample*/, void* stream);

/* Training backward of one track at the final level: gradients of the ObjMLP weights, the track's latent code and
 * the object table (all ACCUMULATED into the caller's buffers, same shapes as the parameters) from the gradients
 * w.r.t. the merged density[N,S] / rgb[N,S,3], for the samples the track owns (owner[pt] == track).  The forward
 * is recomputed; poses are constants. */
typedef struct {
  float *g_W_d0, *g_b_d0, *g_W_d2, *g_b_d2, *g_W_v0, *g_b_v0, *g_W_v1, *g_b_v1, *g_W_rgb, *g_b_rgb;
  float* g_latent;   /* or NULL */
  float* g_table;    /* [rows, C] or NULL */
  float* g_pose;     /* [N, n_obj, 9] or NULL: gradient w.r.t. the interpolated pose (centre, yaw, wlh) of THIS track's
                        column, accumulated -- track refinement, Z/train.py:244-257 */
} nlb_obj_grads_t;
int nlb_obj_backward(const float* tdist, const float* origins, const float* directions, const float* viewdirs,
                     const float* pose, int n_obj, int track, int N, int S, const nlb_table_t* table,
                     const nlb_obj_mlp_t* mlp, const int32_t* owner, const float* g_density /*or NULL*/,
                     const float* g_rgb /*or NULL*/, const nlb_obj_grads_t* grads, void* stream);

/* ------------------------------------------------------------------ stage-3 ray-drop: around the U-Net
 * `R/` = NeRF_LiDAR/NeRF_Lidar_code/.  The U-Net (R/src/unet/) is not part of the library: its [2,H,W] logits are
 * an input of nlb_raydrop_select.
 */
/* R/src/depth_filter.py:4-31: points[H,W,3] (beam-major sweep), semantic[H*W] or NULL -> mask[H*W] (bytes). */
int nlb_depth_filter(const float* points, const float* semantic, int H, int W, int width, float radius, int threshold,
                     uint8_t* mask, void* stream);

/* LaserScan.do_range_projection (R/src/lidar_utils.py:209-275).  Images are [H,W] (xyz / rgb [H,W,3]); any pointer
 * may be NULL.  Empty pixels: range / xyz / semantic = -1, rgb = 0, idx = -1; proj_mask = (proj_idx > 0). */
typedef struct {
  float* proj_range; float* proj_xyz; float* proj_semantic; float* proj_rgb; int32_t* proj_idx; float* proj_mask;
} nlb_range_image_t;
size_t nlb_range_projection_workspace_bytes(int H, int W);
int nlb_range_projection(const float* points /*[n,3]*/, const float* semantic /*[n] or NULL*/, const float* rgb /*[n,3] or NULL*/,
                         int n, int H, int W, float fov_up_deg, float fov_down_deg, int32_t* proj_x /*[n]*/,
                         int32_t* proj_y /*[n]*/, float* unproj_range /*[n]*/, const nlb_range_image_t* image,
                         void* workspace, void* stream);

/* The drop selection of R/src/drop_simulation_rays.py:104-140 (save_near=False): a point survives when the
 * softmax probability of class 1 at its pixel exceeds mask_thre, its pixel is occupied (proj_mask), the depth filter
 * kept it (filter_mask, or NULL), and it is neither sky (label 10) nor a road outlier (label 0, z < -3).  Survivors
 * are written in their original order to remain_points[*,3] / remain_labels[*] (capacity n); *remain_count (device
 * int) receives their number. */
size_t nlb_raydrop_select_workspace_bytes(int n);
int nlb_raydrop_select(const float* logits /*[2,H,W]*/, float mask_thre, const float* proj_mask, const int32_t* proj_x,
                       const int32_t* proj_y, const uint8_t* filter_mask, const float* points, const float* labels, int n,
                       int H, int W, float* remain_points, float* remain_labels, int* remain_count, void* workspace,
                       void* stream);

/* The ray-drop U-Net itself, inference (R/src/unet/unet_model.py:6-47, unet_parts.py:8-77; n_classes logits per
 * pixel, `regression` head not built).  BatchNorm is passed folded (eval mode): scale = gamma / sqrt(var + eps),
 * shift = beta - mean * scale.  Weights in torch layout: Conv2d [OC, C, 3, 3], ConvTranspose2d [C, C/2, 2, 2]. */
typedef struct {
  const float* weight; const float* scale; const float* shift;
  const float* packed;   /* nlb_unet_pack_conv() copy of `weight`, or NULL: with it the layer runs on the tensor cores
                            in TF32 -- what torch's default (torch.backends.cudnn.allow_tf32 = True) makes of the
                            reference's Conv2d on a GPU; without it in strict fp32 */
} nlb_unet_conv_t;
typedef struct {
  nlb_unet_conv_t inc[2];          /* DoubleConv(n_channels, 64) */
  nlb_unet_conv_t down[4][2];      /* Down(64,128) ... Down(512, 1024 / factor) */
  nlb_unet_conv_t up[4][2];        /* Up(1024, 512 / factor) ... Up(128, 64): DoubleConv after the concatenation */
  const float* up_weight[4];       /* ConvTranspose2d weights / biases (bilinear == 0), else NULL */
  const float* up_bias[4];
  const float* up_packed[4];       /* nlb_unet_pack_convtranspose() copies of up_weight (TF32 tensor-core path) or NULL */
  const float* outc_weight;        /* [n_classes, 64] */
  const float* outc_bias;          /* [n_classes] */
  const float* outr_weight;        /* [1, 64] range-regression head (UNet(regression=True)) or NULL */
  const float* outr_bias;          /* [1] */
  int bilinear;                    /* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) instead, factor = 2 */
  int n_classes;
} nlb_unet_weights_t;
size_t nlb_unet_workspace_bytes(int N, int H, int W);
/* weight[OC,C,3,3] -> packed[OC*C*9] (pre-swizzled [oc tile][32-channel chunk][tap] operand blocks); OC % 64 == 0 and
 * C % 32 == 0, else NLB_EUNSUPPORTED (such a layer keeps packed = NULL). */
int nlb_unet_pack_conv(const float* weight, int OC, int C, float* packed, void* stream);
/* ConvTranspose2d weight[C,OC,2,2] -> packed[C*OC*4], same conditions. */
int nlb_unet_pack_convtranspose(const float* weight, int C, int OC, float* packed, void* stream);
/* image[N,Cin,H,W] -> logits[N,n_classes,H,W] (+ regression[N,1,H,W] = sigmoid(outr(x)), or NULL); H, W multiples
 * of 16. */
int nlb_unet_forward(const float* image, const nlb_unet_weights_t* w, int N, int Cin, int H, int W, float* logits,
                     float* regression, float* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NLB200_H */
