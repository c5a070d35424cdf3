/*
 * rabi_b200 -- C ABI of the B200-native (sm_100a) backend for RABI's MAF hot path.
 *
 * The reference (mackelab/RABI, /root/reference) is pure Python and has NO FFI of its own: its seam is the
 * hydra "module_path + class_name" plugin boundary (src/rbibm/rbibm/scripts/rbibm_script.py:597-636,753-764).
 * This header is therefore the boundary a maintainer binds with ctypes; every entry point
 * names the reference code whose arithmetic it replaces.  Paths below are relative to /root/reference/src.
 *
 * Conventions
 *  - all tensors are fp32, row-major, contiguous; "_dev" pointers are device pointers owned by the caller
 *    (normally torch allocations), "_host" pointers are host memory;
 *  - the library owns only the packed weight image and a scratch workspace (grown on demand);
 *  - every compute call takes the CUDA stream to launch on (cudaStream_t passed as void*; 0 = default stream);
 *  - return value: 0 = ok, non-zero = error; rabi_last_error() gives the message.  No exceptions cross the ABI;
 *  - a handle belongs to one device and is not thread-safe;
 *  - sample-major layout: theta / z are [S, B, D] (sample, observation row, dimension); per-row results are [B];
 *    "A" is the layer-0 context projection of all K transforms, [K, H0, B] (row index fastest).
 */
#ifndef RABI_B200_H
#define RABI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rabi_maf rabi_maf_t;
typedef void* rabi_stream_t; /* cudaStream_t */

#define RABI_MAX_HIDDEN_LAYERS 4
#define RABI_MAX_TRANSFORMS 8

/* Library / build information ("sm_100a", version).  Never NULL. */
const char* rabi_build_info(void);

/* Message of the last failed call on this handle (h may be NULL: last create failure). */
const char* rabi_last_error(const rabi_maf_t* h);

/* ---- construction: replaces InverseAffineAutoregressiveModel.__init__ -> PyroFlowModel.__init__ building K
 * ConditionalAutoRegressiveNN(D, C, hidden) (rbi/rbi/models/flows.py:72-113, models/base.py:277-343,
 * models/transforms.py:92-98).  log_scale_{min,max}: AffineAutoregressive clamp (config/model/maf_pyro.yaml:6-7). */
int rabi_maf_create(rabi_maf_t** out, int device, int K, int D, int C, int n_hidden, const int* hidden,
                    float log_scale_min, float log_scale_max);
int rabi_maf_destroy(rabi_maf_t* h);

/* MADE structure of transform k: the `permutation` buffer (int64[D]) and the `mask` buffer of MaskedLinear
 * layer `layer` (0..n_hidden; fp32 [out, in], in = C + D for layer 0, out = 2*D for the last layer), exactly as
 * they sit in the reference state_dict (t{k}.nn.permutation, t{k}.nn.layers.{layer}.mask).  Host pointers. */
int rabi_maf_set_permutation(rabi_maf_t* h, int k, const int64_t* perm_host);
int rabi_maf_set_mask(rabi_maf_t* h, int k, int layer, const float* mask_host, int out_features, int in_features);
/* Optional Permute layer applied after transform k (shuffle=True: buffers permute{k}, models/base.py:326-333;
 * pyro Permute: y[i] = x[perm[i]]).  NULL removes it.  Folded into the kernels' gather/scatter indices. */
int rabi_maf_set_post_permutation(rabi_maf_t* h, int k, const int64_t* perm_host);
/* Validates that every mask is the MADE prefix structure Pyro's create_mask produces, builds the packed image
 * layout and uploads the index tables.  Must be called once after all permutations/masks are set. */
int rabi_maf_finalize_structure(rabi_maf_t* h);

/* Parameters of MaskedLinear layer `layer` of transform k (t{k}.nn.layers.{layer}.{weight,bias}): RAW (unmasked)
 * weights [out, in] and bias [out] as DEVICE pointers; the mask multiply of MaskedLinear.forward
 * (F.linear(x, W*mask, b)) and the re-layout are done once here instead of on every call. */
int rabi_maf_set_weights(rabi_maf_t* h, int k, int layer, const float* W_dev, const float* b_dev, rabi_stream_t s);

/* Layer-0 context projection for all K transforms: A[k, u, b] = b0_k[u] + sum_c W0_k[u, c] * ctx[b, c], with
 * ctx = (x - zs_mean) / zs_std when zs_* are non-NULL (ZScoreLayer, rbi/rbi/models/module.py:109-122), else ctx = x.
 * This is the part of ConditionalAutoRegressiveNN.forward that depends only on the observation row. */
int rabi_maf_ctx_project(rabi_maf_t* h, const float* x_dev, const float* zs_mean_dev, const float* zs_std_dev,
                         int B, float* A_dev, rabi_stream_t s);
/* Transposed projection (input gradient): g_x[b, c] (+)= (1/zs_std[c]) * sum_{k,u} W0_k[u, c] * gA[k, u, b]. */
int rabi_maf_ctx_project_bwd(rabi_maf_t* h, const float* gA_dev, const float* zs_std_dev, int B, float* gx_dev,
                             rabi_stream_t s);

/* q.log_prob(theta): K MADE passes (TransformedDistribution.log_prob over AffineAutoregressive._call,
 * models/base.py:417-446 + pyro AffineAutoregressive).  theta [S,B,D] -> logp [S,B]. */
int rabi_maf_log_prob_fwd(rabi_maf_t* h, const float* A_dev, const float* theta_dev, int S, int B, float* logp_dev,
                          rabi_stream_t s);
/* q.rsample from given base noise z [S,B,D] (AffineAutoregressive._inverse, one pass-equivalent instead of D
 * passes); logq_dev (optional) receives log q(theta) of the drawn samples (the "free" cache-hit log_prob, Q9). */
int rabi_maf_rsample_fwd(rabi_maf_t* h, const float* A_dev, const float* z_dev, int S, int B, float* theta_dev,
                         float* logq_dev, rabi_stream_t s);

/* MC reverse KL per row, no gradient: kl[b] = mean_s[log p(th_s) - log q(th_s)], th_s ~ p via z
 * (general_kl_divergence, rbi/rbi/loss/kl_div.py:38-56; ReverseKLLoss, loss/eval_loss.py:129-150).
 * A_p: projection of the distribution sampled from (first KL argument), A_q: of the second. */
int rabi_rkl_fwd(rabi_maf_t* h, const float* A_p_dev, const float* A_q_dev, const float* z_dev, int S, int B,
                 float* kl_dev, rabi_stream_t s);

/* One L2-PGD iteration on the reverse-KL attack loss, identity(+z-score) embedding (dx == C):
 *   loss = inv_B * sum_b KL(q(.|x+delta) || q(.|x));  g = d loss / d delta;
 *   g /= max(||g||_2, norm_floor);  delta += eps_iter*g;  delta = clamp(x+delta, clip_min, clip_max) - x;
 *   delta *= min(1, eps/||delta||_2)
 * (advertorch perturb_iterative ord=2 as driven by rbi/rbi/attacks/advertorch_attack.py:100-114 with
 * loss_fn = ReverseKLLoss(reduction="mean")).  A_q = rabi_maf_ctx_project of the clean x (iteration invariant).
 * kl_dev (optional, [B]) receives the per-row KL estimate of this iteration. */
int rabi_rkl_pgd_step(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* zs_mean_dev,
                      const float* zs_std_dev, const float* A_q_dev, const float* z_dev, int S, int B, float eps,
                      float eps_iter, float clip_min, float clip_max, float inv_B, float norm_floor, float* kl_dev,
                      rabi_stream_t s);
/* Gradient only: gx[b, :] = d/dx~ [ inv_B * KL(q(.|x~_b) || q_target_b) ] for x~ = x_dev (what ``loss.backward()`` leaves
 * in ``delta.grad`` inside perturb_iterative); kl_dev (optional) the per-row KL. */
int rabi_rkl_input_grad(rabi_maf_t* h, const float* x_dev, const float* zs_mean_dev, const float* zs_std_dev,
                        const float* A_q_dev, const float* z_dev, int S, int B, float inv_B, float* gx_dev,
                        float* kl_dev, rabi_stream_t s);

/* nb_iter PGD iterations (z_all_dev: [nb_iter, S, B, D]) followed by x_adv = clamp(x + delta, clip_min, clip_max).
 * A_q_dev: projection of the attack target; NULL = untargeted, i.e. the projection of the clean x (computed here).
 * A negative inv_B descends on the loss (targeted attack, perturb_iterative(minimize=True)). */
int rabi_rkl_pgd_run(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* zs_mean_dev,
                     const float* zs_std_dev, const float* A_q_dev, const float* z_all_dev, int nb_iter, int S, int B, float eps,
                     float eps_iter, float clip_min, float clip_max, float inv_B, float norm_floor,
                     float* x_adv_dev, rabi_stream_t s);

/* The update half of one advertorch perturb_iterative(ord=2) iteration on its own, for attacks whose gradient has to be
 * chained through a non-identity embedding net by the caller (x-space PGD, e.g. pyloric's CNN embedding):
 *   g = grad / max(||grad||_2, norm_floor);  delta += eps_iter*g;  delta = clamp(x+delta) - x;  delta *= min(1, eps/||delta||_2) */
int rabi_pgd_update(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* grad_dev, int B, int dx, float eps,
                    float eps_iter, float clip_min, float clip_max, float norm_floor, rabi_stream_t s);

/* Training side (replaces the backward pass torch autograd builds for NLLLoss, rbi/rbi/loss/train_loss.py:16-29, through
 * TransformedDistribution.log_prob, rbi/rbi/models/base.py:417-446): for theta_dev [S,B,D] (unconstrained space) and the rows'
 * context ctx_dev [B,C] (after the z-score / embedding) with projection A_dev [K,H0,B] (rabi_maf_ctx_project), writes
 * logq_dev [S,B] (optional) and, for the loss L with dL/dlogq = gout_dev [S,B]:
 *   g_theta_dev [S,B,D] = dL/dtheta (optional), g_A_dev [K,H0,B] = dL/dA summed over the S samples of a row (optional; chain it
 *   to x with rabi_maf_ctx_project_bwd), and the weight / bias gradients in the REFERENCE layout: g_weight_dev[k*(L+1)+l] is
 *   [out_l, in_l] like t{k}.nn.layers.{l}.weight (masked entries get exact zeros, as autograd of weight*mask gives),
 *   g_bias_dev[k*(L+1)+l] is [out_l].  g_weight_dev == NULL: no weight gradients.  All buffers are overwritten.
 * Two-hidden-layer conditioners (<= 120 padded units, D <= 12) without Permute layers; other shapes return an error and the
 * caller keeps its differentiable path. */
int rabi_maf_log_prob_bwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* theta_dev,
                          const float* gout_dev, int S, int B, float* logq_dev, float* g_theta_dev, float* g_A_dev,
                          float* const* g_weight_dev, float* const* g_bias_dev, rabi_stream_t s);

/* Reverse of the reparameterised sampling chain (TransformedDistribution.rsample, rbi/rbi/models/base.py:417-446, as
 * differentiated by the TRADES regulariser, rbi/rbi/defenses/trades_regularizers.py:58-64, and by the Fisher regulariser's path
 * through its samples, rbi/rbi/utils/fisher_info.py:213-221): for base noise z_dev [S,B,D], theta = T(z; x, W) and its own
 * log-density, and upstream gradients g_theta_dev [S,B,D] = dL/dtheta, g_logq_dev [S,B] = dL/dlog q(theta), writes theta_dev /
 * logq_dev (optional: the forward values), g_A_dev [K,H0,B] (optional) and the weight / bias gradients exactly as
 * rabi_maf_log_prob_bwd does.  Same shape restrictions. */
int rabi_maf_rsample_bwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* z_dev,
                         const float* g_theta_dev, const float* g_logq_dev, int S, int B, float* theta_dev, float* logq_dev,
                         float* g_A_dev, float* const* g_weight_dev, float* const* g_bias_dev, rabi_stream_t s);

/* Fisher-information trace of q(theta | x) w.r.t. the observation, MC score estimator (rbi/rbi/utils/fisher_info.py:185-280,
 * method "score_based", typ "trace"; the quantity FIMTraceRegularizer penalises, rbi/rbi/defenses/fisher_regularization.py:67-95):
 *   theta_s = T(z_s; x_b) (sampling chain), score_{s,b} = d log q(theta_s | x_b) / d x_b (density passes, their reverse and the
 *   back-projection with every (s, b) pair as a row), trace_b = mean_s ||score_{s,b}||^2 -- four launch groups, no autograd.
 * ctx_dev [B,C] context rows (after z-score / embedding), A_dev [K,H0,B] their projection, z_dev [S,B,D] base noise,
 * zs_std_dev [C] or NULL (the z-score's 1/std of the back-projection); outputs trace_dev [B] (optional), score_dev [S,B,C],
 * theta_dev [S,B,D] (both required: they double as workspace).  Same shape restrictions as rabi_maf_log_prob_bwd. */
int rabi_fisher_trace_fwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* z_dev, const float* zs_std_dev,
                          int S, int B, float* trace_dev, float* score_dev, float* theta_dev, rabi_stream_t s);

/* Weight gradient of the Fisher-trace regulariser R = sum_b w_b tr F(x_b) (w_row_dev [B] = dR / d tr F_b: beta / (n S) for the
 * rows the reduction keeps, 0 for the others) -- what FIMTraceRegularizer obtains from autograd.grad of the MC estimator with
 * create_graph=True (rbi/rbi/defenses/fisher_regularization.py:67-95, rbi/rbi/utils/fisher_info.py:185-280), as a closed form
 * without an autograd graph (derivation and fp64 check: oracle/fim_closed_form_check.py).  Must follow rabi_fisher_trace_fwd on
 * the same handle with the same ctx / A / z / S / B (it consumes the per-pair vectors that call left in the workspace) and the
 * scores it returned.  Launches: k_fim_cdot, k_fim_dual (tangent density pass + the two masked reverse passes per transform,
 * 64 pairs per CTA), k_wgrad (24 stacked weight products), then the sampling-chain reverse (k_tc<TC_RSB> + k_wgrad) seeded with
 * d phi / d theta.  wm_dev[k*3 + l]: weight * mask of transform k, layer l in the reference layout ([H0, C+D], [H1, H0],
 * [2D, H1]); bo_dev[k]: output-layer bias [2D]; g_weight_dev / g_bias_dev as in rabi_maf_log_prob_bwd (overwritten). */
int rabi_fisher_trace_bwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* z_dev, const float* zs_std_dev,
                          int S, int B, const float* score_dev, const float* w_row_dev, const float* const* wm_dev,
                          const float* const* bo_dev, float* const* g_weight_dev, float* const* g_bias_dev, rabi_stream_t s);

/* The per-pair vectors of the LAST training-side call on this handle (same S, B), copied unit-major ([.][P], P = S*B, the pair
 * index fastest) into dst_dev: [h0 | a0bar : K*H0*P each][h1 | a1bar : K*H1*P each][obar : K*2D*P][y : K*D*P] -- layer
 * activations, cotangents of the pre-activations / outputs (reference row order) and the theta-side inputs of the transforms.
 * Host code builds further products on them (the Fisher regulariser's closed-form weight gradient, rabi_b200/fisher.py). */
int rabi_train_emissions(rabi_maf_t* h, int S, int B, float* dst_dev, rabi_stream_t s);

/* Per-row attack radii for batched eps-sweeps (config/experiment/eval_l2_pyloric.yaml:13-16 sweeps eps over 6 values;
 * run_rob_eval.py:115-140 derives eps_abs and eps_iter per value): rows x eps values are independent, so one call can attack
 * the rows at all radii at once.  While set, every rabi_*_pgd_step / rabi_*_pgd_run / rabi_pgd_update on this handle uses
 * eps_rows_dev[b] / eps_iter_rows_dev[b] ([B] device arrays owned by the caller) instead of its scalar eps / eps_iter
 * arguments.  Pass NULL, NULL to return to the scalars. */
int rabi_pgd_set_row_scales(rabi_maf_t* h, const float* eps_rows_dev, const float* eps_iter_rows_dev);

/* Masked-feature attacks (RobustnessMetric(mask=...), rbibm/rbibm/metric/base.py:232-250: only the features selected by the mask
 * are perturbed -- the reference attacks the sub-vector x[..., mask] through a wrapper model).  While set, every PGD update on
 * this handle multiplies the input gradient by col_mask_dev [C] (1 = attacked, 0 = frozen) before the normalisation and keeps
 * delta == 0 on the frozen columns (they are not clamped either).  NULL clears it. */
int rabi_pgd_set_column_mask(rabi_maf_t* h, const float* col_mask_dev);

/* Forward-KL attack loss (ForwardKLLoss, rbi/rbi/loss/eval_loss.py:108-126 -- the default attack_loss_fn of
 * config/eval_rob/attack/l2pgd.yaml): loss = inv_B * sum_b KL(target_b || q(.|x_b + delta_b)); samples are drawn from the
 * fixed target (projection A_t_dev; NULL in the run = the clean x), the gradient flows through the density pass only.
 * Same argument meaning as the rabi_rkl_* counterparts. */
int rabi_fkl_input_grad(rabi_maf_t* h, const float* x_dev, const float* zs_mean_dev, const float* zs_std_dev,
                        const float* A_t_dev, const float* z_dev, int S, int B, float inv_B, float* gx_dev,
                        float* kl_dev, rabi_stream_t s);
int rabi_fkl_pgd_run(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* zs_mean_dev,
                     const float* zs_std_dev, const float* A_t_dev, const float* z_all_dev, int nb_iter, int S, int B,
                     float eps, float eps_iter, float clip_min, float clip_max, float inv_B, float norm_floor,
                     float* x_adv_dev, rabi_stream_t s);

/* Measurement support (bench.py): CUDA-event timing of the dominant kernel (the per-pair rKL forward+backward
 * kernel) on the stream it is launched on.  rabi_profile_read synchronises the device, returns the summed duration
 * and the number of timed launches since the last read, and resets the counters. */
int rabi_profile(rabi_maf_t* h, int enable);
int rabi_profile_read(rabi_maf_t* h, double* total_ms, unsigned long long* launches);
/* Measured fp32 FMA-pipe peak of `device` in TFLOP/s (best of `repeats` runs of a register-only FFMA kernel);
 * the roofline denominator of the FMA-bound kernels.  < 0 on error. */
double rabi_fma_peak_tflops(int device, int repeats);

/* fp32 GEMM primitive of the differentiable (training-side) path: C[M,N] = beta*C + op(A) * op(B); row-major with
 * leading dimensions, op = transpose when trans* != 0.  It carries every matrix product of the first- and second-
 * order gradients torch autograd builds for TrainLoss / FIMTraceRegularizer / L2PGDrKLTrades (rbi/rbi/loss/base.py:
 * 142-175, defenses/fisher_regularization.py:67-95, defenses/trades_regularizers.py:58-64); the masked products
 * F.linear(x, W*mask, b) of Pyro's MaskedLinear are instances of it. */
int rabi_sgemm(int device, int transA, int transB, int M, int N, int K, const float* A_dev, int lda,
               const float* B_dev, int ldb, float* C_dev, int ldc, float beta, rabi_stream_t s);
unsigned long long rabi_sgemm_launch_count(void);

/* Number of kernel launches issued through this handle so far (bench.py's gpu_launches). */
unsigned long long rabi_launch_count(const rabi_maf_t* h);

#ifdef __cplusplus
}
#endif
#endif /* RABI_B200_H */
