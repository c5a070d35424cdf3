// Tensor-core path (maf_tc.cu): interface towards the launch dispatch in rabi_maf.cu.
#pragma once
#include "rabi_internal.h"

namespace rabi {

// same numbering as FastMode (maf_fast.cuh)
enum TcMode { TC_FWD = 0, TC_GRAD = 1, TC_FKL = 2, TC_RSAMPLE = 3, TC_LOGPROB = 4, TC_NLL = 5, TC_RSB = 6 };

struct TcCall {
    const float* A_p;   // [K,H0,B] projection of the sampled-from posterior (LOGPROB: unused)
    const float* A_q;   // [K,H0,B] projection of the posterior whose density is evaluated
    const float* z;     // [S,B,D] base noise (LOGPROB: theta)
    float* diff;        // [S,B] log p - log q | log q(own sample) | log q(theta); optional in the gradient modes
    float* theta_out;   // RSAMPLE: [S,B,D]
    float* gA;          // GRAD / FKL: [K,H0,B], zeroed by the caller; the pairs add their parts with float reductions
    int S, B;
    float w;            // weight of one pair in the loss (inv_B / S)
    // TC_NLL (training side: log q(theta | x) and the reverse of its density chain): z = theta, A_q = the projection,
    // loss = - sum_p wrow[p] log q_p; besides gA (summed over the S samples of a row) the per-pair vectors the weight-gradient
    // products need are written unit-major ([.][npairs], the pair index fastest)
    const float* wrow;  // [S,B] per-pair weights (NULL: w)
    float* em_h0;       // [K][H0][P] layer-0 activations          float* em_a0: d loss / d (layer-0 pre-activation)
    float* em_a0;
    float* em_h1;       // [K][H1][P]
    float* em_a1;
    float* em_o;        // [K][2D][P] d loss / d (mean, log-scale) outputs, REAL row order (row d = mean of theta_d, D + d = log-scale)
    float* em_y;        // [K][D][P]  theta-side input of transform k, real index order
    float* theta_grad;  // [P][D]     TC_NLL: d loss / d theta (out);  TC_RSB: d loss / d theta_sample (in)
    // TC_RSB (reverse of the reparameterised sampling chain): z = base noise, A_p = the projection, loss = sum_p
    // <theta_grad_p, theta_p> + wrow_p log q(theta_p); the same per-pair vectors are emitted (em_y = the transforms' outputs)
};

// The whole L2-PGD attack (nb_iter iterations) as one persistent launch (two-hidden-layer conditioners, identity / z-score
// embedding): projection, flow passes, gradient back-projection and the perturb_iterative(ord=2) update inside the kernel.
struct TcRunCall {
    const float* A_p;     // reverse KL: [K,H0,B] slab the kernel rewrites every iteration; forward KL: the fixed target's projection
    const float* A_q;     // reverse KL: projection of the clean x (fixed); forward KL: the slab that follows delta
    const float* z_all;   // [nb_iter,S,B,D]
    float* gA;            // [K,H0,B], zeroed by the caller; left zeroed
    const float* x;       // [B,C]
    float* delta;         // [B,C] in / out
    const float* zmean;   // [C] or null (with zstd)
    const float* zstd;
    int nb_iter, S, B;
    float eps, eps_iter, cmin, cmax, inv_B, floor_;
    const float* eps_rows;       // optional [B] per-row radius / step size (batched eps-sweep); else the scalars
    const float* eps_iter_rows;
    const float* col_mask;       // optional [C] attack mask (1 = attacked, 0 = frozen)
};
int rabi_tc_run(rabi_maf* h, const TcRunCall& r, int mode, cudaStream_t st);

// returns 0 = launched, 1 = error (message on the handle), 2 = shape / launch not eligible for this path
int rabi_tc_launch(rabi_maf* h, const TcCall& c, int mode, cudaStream_t st);
void rabi_tc_destroy(rabi_maf* h);
// training side (maf_train.cu): log q with gradients w.r.t. theta, the context projection and every weight
int rabi_train_log_prob_bwd(rabi_maf* h, const float* ctx, const float* A, const float* theta, const float* gout, int S, int B,
                            float* logq, float* g_theta, float* g_A, float* const* g_weight, float* const* g_bias, cudaStream_t st);
void rabi_train_destroy(rabi_maf* h);
// three hidden layers up to 200 units wide, D <= 32 (maf_tcd.cu); reached through rabi_tc_launch
int rabi_tcd_launch(rabi_maf* h, const TcCall& c, int mode, cudaStream_t st);
void rabi_tcd_destroy(rabi_maf* h);

}  // namespace rabi
