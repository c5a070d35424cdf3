// Training side of the hot path: log q(theta | x) with gradients w.r.t. theta, the context projection and EVERY weight of the
// K masked-MADE conditioners, without an autograd graph (rabi_maf_log_prob_bwd).
//
// Reference arithmetic replaced: the backward pass torch autograd builds for NLLLoss (rbi/rbi/loss/train_loss.py:16-29)
// through TransformedDistribution.log_prob (rbi/rbi/models/base.py:417-446) over Pyro's ConditionalAutoRegressiveNN /
// AffineAutoregressive (rbi/rbi/models/transforms.py:37-46), i.e. for the loss  L = - sum_p w_p log q(theta_p | x_b(p)):
//   d L / d theta, d L / d A (A = layer-0 context projection; the caller chains it to x with rabi_maf_ctx_project_bwd) and,
//   per transform k,  dW_l = mask_l * (Abar_l^T H_{l-1}),  db_l = sum_p Abar_l[p]   for the layers l = 0, 1, output.
//
// Two launches:
//   1. k_tc<TC_NLL> (maf_tc.cu): one thread per (sample, row) pair runs the K density passes and their reverse on the tensor
//      cores and leaves, unit-major and coalesced ([unit][pair]), the vectors the weight products need: activations h0, h1, the
//      theta-side inputs y, and the cotangents of the pre-activations / outputs a0bar, a1bar, obar;
//   2. k_wgrad (below): ALL weight-gradient products of all transforms in one launch -- a list of small "NT" GEMMs
//      C[M x N] = A[M x P] B[N x P]^T whose reduction dimension is the pair index, split over CTAs (fp32 atomics into the zeroed
//      gradient buffers), with the MADE mask applied to the tile and the bias gradient (row sums of A) folded in.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "maf_tc.h"

namespace rabi {

constexpr int WG_MAXD = 32;    // products per launch
constexpr int WG_TM = 32, WG_TN = 32, WG_KC = 32;

struct WgDesc {
    const float* A;      // [M][lda]   cotangents, unit-major (pair index contiguous)
    const float* B;      // b_pmajor == 0: [N][ldb] unit-major;  1: [P][ldb] pair-major (context rows)
    float* out;          // [M][ldo] + column offset already applied
    const float* mask;   // [M][ldo] (+ the same offset) or null
    float* bias;         // [M] += sum_p A[m][p]  (null: none)
    int M, N, P, lda, ldb, ldo, b_pmajor;
    int b_mod;           // b_pmajor: row of B = p % b_mod (0: p)
    int tile0;           // first CTA of this product
    int tn, ksplit;      // tiles along N, K chunks
};

struct WgArgs {
    WgDesc d[WG_MAXD];
    int nd;
};

__global__ void __launch_bounds__(256) k_wgrad(const WgArgs a) {
    __shared__ float As[WG_KC][WG_TM + 1];
    __shared__ float Bs[WG_KC][WG_TN + 1];
    int di = 0;
    while (di + 1 < a.nd && (int)blockIdx.x >= a.d[di + 1].tile0) ++di;
    const WgDesc& d = a.d[di];
    int t = (int)blockIdx.x - d.tile0;
    const int ks = t % d.ksplit;
    t /= d.ksplit;
    const int m0 = (t / d.tn) * WG_TM, n0 = (t % d.tn) * WG_TN;
    const int per = ((d.P + d.ksplit - 1) / d.ksplit + WG_KC - 1) / WG_KC * WG_KC;
    const int p0 = ks * per, p1 = min(d.P, p0 + per);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 2 x 2 outputs each
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float bsum = 0.f;   // threads with tx == 0 own the bias rows m0 + 2 ty + {0, 1} ... summed over the tile's first column block
    for (int pc = p0; pc < p1; pc += WG_KC) {
        // A tile: rows m0..m0+31, pairs pc..pc+31 (pair index fastest in memory: lanes walk the pairs)
        for (int i = threadIdx.x; i < WG_TM * WG_KC; i += 256) {
            const int kk = i % WG_KC, mm = i / WG_KC;
            const int m = m0 + mm, p = pc + kk;
            As[kk][mm] = (m < d.M && p < p1) ? d.A[(size_t)m * d.lda + p] : 0.f;
        }
        if (d.b_pmajor) {   // context rows [P][ldb]: lanes walk the columns
            for (int i = threadIdx.x; i < WG_TN * WG_KC; i += 256) {
                const int nn = i % WG_TN, kk = i / WG_TN;
                const int n = n0 + nn, p = pc + kk;
                Bs[kk][nn] = (n < d.N && p < p1) ? d.B[(size_t)(d.b_mod ? p % d.b_mod : p) * d.ldb + n] : 0.f;
            }
        } else {
            for (int i = threadIdx.x; i < WG_TN * WG_KC; i += 256) {
                const int kk = i % WG_KC, nn = i / WG_KC;
                const int n = n0 + nn, p = pc + kk;
                Bs[kk][nn] = (n < d.N && p < p1) ? d.B[(size_t)n * d.ldb + p] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < WG_KC; ++kk) {
            const float a0 = As[kk][2 * ty], a1 = As[kk][2 * ty + 1];
            const float b0 = Bs[kk][2 * tx], b1 = Bs[kk][2 * tx + 1];
            acc[0][0] = fmaf(a0, b0, acc[0][0]);
            acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]);
            acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        if (d.bias && n0 == 0 && threadIdx.x < WG_TM) {
            float s = 0.f;
#pragma unroll 8
            for (int kk = 0; kk < WG_KC; ++kk) s += As[kk][threadIdx.x];
            bsum += s;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + 2 * ty + i;
        if (m >= d.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + 2 * tx + j;
            if (n >= d.N) continue;
            const size_t o = (size_t)m * d.ldo + n;
            const float v = d.mask ? acc[i][j] * d.mask[o] : acc[i][j];
            if (v != 0.f) atomicAdd(d.out + o, v);
        }
    }
    if (d.bias && n0 == 0 && threadIdx.x < WG_TM && m0 + (int)threadIdx.x < d.M) atomicAdd(d.bias + m0 + threadIdx.x, bsum);
}

__global__ void k_fill(float* __restrict__ out, float v, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

__global__ void k_negate(const float* __restrict__ in, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = -in[i];
}

struct TrainState {
    float* masks = nullptr;                 // every mask, reference layout, one after the other
    std::vector<size_t> mask_off;           // [K * (L + 1)]
    unsigned long long masks_of = ~0ull;    // structure version the masks were uploaded for (weights_version at upload)
    float* ws = nullptr;                    // emission buffers
    size_t ws_floats = 0;
    float* fim = nullptr;                   // buffers of the Fisher-gradient dual pass
    size_t fim_floats = 0;
};

static int layer_out(const rabi_maf* h, int layer) { return layer < h->L ? h->H[layer] : 2 * h->D; }
static int layer_in(const rabi_maf* h, int layer) { return layer == 0 ? h->C + h->D : h->H[layer - 1]; }

// Zeroes the gradient buffers: ONE memset when the caller laid them out as one block [W(k,l) | b(k,l)] in (k, l) order (the
// Python binding does), else one per tensor.
static int zero_grads(rabi_maf* h, float* const* g_weight, float* const* g_bias, cudaStream_t st) {
    const int n = h->K * (h->L + 1);
    bool contiguous = true;
    size_t total = 0;
    for (int i = 0; i < n && contiguous; ++i) {
        const int l = i % (h->L + 1);
        const size_t nw = (size_t)layer_out(h, l) * layer_in(h, l), nb = (size_t)layer_out(h, l);
        contiguous = g_bias[i] == g_weight[i] + nw && (i + 1 == n || g_weight[i + 1] == g_bias[i] + nb);
        total += nw + nb;
    }
    if (contiguous) {
        RABI_CU_OK(h, cudaMemsetAsync(g_weight[0], 0, total * sizeof(float), st));
        return 0;
    }
    for (int i = 0; i < n; ++i) {
        const int l = i % (h->L + 1);
        RABI_CU_OK(h, cudaMemsetAsync(g_weight[i], 0, (size_t)layer_out(h, l) * layer_in(h, l) * sizeof(float), st));
        RABI_CU_OK(h, cudaMemsetAsync(g_bias[i], 0, (size_t)layer_out(h, l) * sizeof(float), st));
    }
    return 0;
}

void rabi_train_destroy(rabi_maf* h) {
    TrainState* s = static_cast<TrainState*>(h->train);
    if (!s) return;
    if (s->masks) cudaFree(s->masks);
    if (s->ws) cudaFree(s->ws);
    if (s->fim) cudaFree(s->fim);
    delete s;
    h->train = nullptr;
}

// Common body of the two training-side entry points.
//   mode TC_NLL: in = theta, gout = dL/dlogq, logq / g_theta optional outputs;
//   mode TC_RSB: in = base noise z, gout = dL/dlogq(own sample), g_theta_in = dL/dtheta_sample, theta_out / logq optional outputs.
// 0 = done, 1 = error, 2 = this shape is not served by the fused training path (the caller keeps its autograd path)
static int train_bwd(rabi_maf* h, int mode, const float* ctx, const float* A, const float* in, const float* gout,
                     const float* g_theta_in, int S, int B, float* theta_out, float* logq, float* g_theta, float* g_A,
                     float* const* g_weight, float* const* g_bias, cudaStream_t st, float** em_a0_out = nullptr,
                     bool accumulate = false) {
    if (h->L != 2) return 2;
    for (int k = 0; k < h->K; ++k)
        if (!h->post[k].empty()) return 2;   // Permute layers between the transforms: index bookkeeping not done here
    TrainState* s = static_cast<TrainState*>(h->train);
    if (!s) h->train = s = new TrainState();
    const int K = h->K, D = h->D, C = h->C, H0 = h->H[0], H1 = h->H[1], L = h->L;
    const size_t P = (size_t)S * B;
    if (!s->masks) {
        std::vector<float> host;
        s->mask_off.clear();
        for (int k = 0; k < K; ++k)
            for (int l = 0; l <= L; ++l) {
                s->mask_off.push_back(host.size());
                host.insert(host.end(), h->masks[k][l].begin(), h->masks[k][l].end());
            }
        RABI_CU_OK(h, cudaMalloc(&s->masks, host.size() * sizeof(float)));
        RABI_CU_OK(h, cudaMemcpy(s->masks, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    // emission buffers: h0, a0bar [K][H0][P]; h1, a1bar [K][H1][P]; obar [K][2D][P]; y [K][D][P]; w [P]
    const size_t n0 = (size_t)K * H0 * P, n1 = (size_t)K * H1 * P, no = (size_t)K * 2 * D * P, ny = (size_t)K * D * P;
    const size_t need = 2 * n0 + 2 * n1 + no + ny + P + (size_t)K * H0 * B + P * D;
    if (need > s->ws_floats) {
        if (s->ws) {
            RABI_CU_OK(h, cudaDeviceSynchronize());
            RABI_CU_OK(h, cudaFree(s->ws));
            s->ws = nullptr;
            s->ws_floats = 0;
        }
        RABI_CU_OK(h, cudaMalloc(&s->ws, need * sizeof(float)));
        s->ws_floats = need;
    }
    float* em_h0 = s->ws;
    float* em_a0 = em_h0 + n0;
    float* em_h1 = em_a0 + n0;
    float* em_a1 = em_h1 + n1;
    float* em_o = em_a1 + n1;
    float* em_y = em_o + no;
    float* wrow = em_y + ny;
    float* gA = g_A ? g_A : wrow + P;
    float* gth = g_theta ? g_theta : wrow + P + (size_t)K * H0 * B;
    if (em_a0_out) *em_a0_out = em_a0;
    // the kernel differentiates  - sum_p w_p log q_p : w = - dL / d log q
    RABI_CU_OK(h, cudaMemsetAsync(gA, 0, (size_t)K * H0 * B * sizeof(float), st));
    TcCall c = {};
    if (mode == TC_NLL) {   // the kernel differentiates - sum_p w_p log q_p
        k_negate<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(gout, wrow, P);
        h->launches++;
        c.A_q = A;
        c.wrow = wrow;
    } else {                // ... and, for the sampling chain, + sum_p (<g_theta_p, theta_p> + w_p log q_p)
        c.A_p = A;
        c.A_q = A;
        c.wrow = gout;
        c.theta_out = theta_out;
    }
    c.z = in;
    c.diff = logq;
    c.gA = gA;
    c.S = S;
    c.B = B;
    c.w = 0.f;
    c.em_h0 = em_h0;
    c.em_a0 = em_a0;
    c.em_h1 = em_h1;
    c.em_a1 = em_a1;
    c.em_o = em_o;
    c.em_y = em_y;
    c.theta_grad = mode == TC_NLL ? gth : const_cast<float*>(g_theta_in);
    const int rc = rabi_tc_launch(h, c, mode, st);
    if (rc) return rc;
    if (!g_weight) return 0;
    // ---- all weight-gradient products in one launch
    WgArgs wa;
    memset(&wa, 0, sizeof(wa));
    int nd = 0, tiles = 0;
    auto add = [&](const float* Aop, int M, const float* Bop, int N, int Pn, int lda, int ldb, int b_pmajor, float* out, int ldo,
                   const float* mask, float* bias) {
        WgDesc& d = wa.d[nd++];
        d.A = Aop;
        d.B = Bop;
        d.out = out;
        d.mask = mask;
        d.bias = bias;
        d.M = M;
        d.N = N;
        d.P = Pn;
        d.lda = lda;
        d.ldb = ldb;
        d.ldo = ldo;
        d.b_pmajor = b_pmajor;
        d.tile0 = tiles;
        d.tn = (N + WG_TN - 1) / WG_TN;
        d.ksplit = std::max(1, std::min(32, (Pn + 255) / 256));
        tiles += ((M + WG_TM - 1) / WG_TM) * d.tn * d.ksplit;
    };
    if (K * 4 > WG_MAXD) return 2;
    for (int k = 0; k < K; ++k) {
        const float* m0 = s->masks + s->mask_off[(size_t)k * (L + 1) + 0];
        const float* m1 = s->masks + s->mask_off[(size_t)k * (L + 1) + 1];
        const float* mo = s->masks + s->mask_off[(size_t)k * (L + 1) + 2];
        float* W0 = g_weight[k * (L + 1) + 0];
        float* W1 = g_weight[k * (L + 1) + 1];
        float* Wo = g_weight[k * (L + 1) + 2];
        if (k == 0 && !accumulate && zero_grads(h, g_weight, g_bias, st)) return 1;
        // layer 0, context columns: the S samples of a row share the context -> the reduced gA [H0][B] times ctx [B][C]
        add(gA + (size_t)k * H0 * B, H0, ctx, C, B, B, C, 1, W0, C + D, m0, g_bias[k * (L + 1) + 0]);
        // layer 0, theta columns: per pair
        add(em_a0 + (size_t)k * H0 * P, H0, em_y + (size_t)k * D * P, D, (int)P, (int)P, (int)P, 0, W0 + C, C + D, m0 + C, nullptr);
        add(em_a1 + (size_t)k * H1 * P, H1, em_h0 + (size_t)k * H0 * P, H0, (int)P, (int)P, (int)P, 0, W1, H0, m1, g_bias[k * (L + 1) + 1]);
        add(em_o + (size_t)k * 2 * D * P, 2 * D, em_h1 + (size_t)k * H1 * P, H1, (int)P, (int)P, (int)P, 0, Wo, H1, mo, g_bias[k * (L + 1) + 2]);
    }
    wa.nd = nd;
    k_wgrad<<<tiles, 256, 0, st>>>(wa);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "k_wgrad launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int rabi_train_log_prob_bwd(rabi_maf* h, const float* ctx, const float* A, const float* theta, const float* gout, int S, int B,
                            float* logq, float* g_theta, float* g_A, float* const* g_weight, float* const* g_bias,
                            cudaStream_t st) {
    return train_bwd(h, TC_NLL, ctx, A, theta, gout, nullptr, S, B, nullptr, logq, g_theta, g_A, g_weight, g_bias, st);
}

}  // namespace rabi

namespace rabi {

// ------------------------------------------------------------------------------------------------------------------------
// Fisher-trace regulariser, weight gradient (rbi/rbi/defenses/fisher_regularization.py:67-95 takes autograd.grad of the MC
// estimator of rbi/rbi/utils/fisher_info.py:185-280 with create_graph=True).  Closed form, derivation and fp64 check in
// oracle/fim_closed_form_check.py: with the scores g_p held constant, dR = sum_p d phi_p, phi_p = the derivative of
// log q(theta_p | x) along xdot_p = 2 w_p g_p.  k_fim_dual runs, for a tile of FD_NP pairs per CTA,
//   * the TANGENT density pass (the primal pass with the stored ReLU masks instead of the sign tests, zero biases), and
//   * the reverse over the (primal, tangent) program seeded on the tangent of log q: two masked reverse passes per transform
//     (adjoints "t" of the tangent and "b" of the primal variables),
// as a chain of small dense products [units x FD_NP pairs] on masked reference-layout weights (packed FMAs, inputs staged in shared
// memory), and leaves every vector the stacked weight products need unit-major in global memory for k_wgrad.
constexpr int FD_NP = 16;            // pairs per CTA: 2560 pairs (batch 512 x 5 samples) spread over 160 CTAs
constexpr int FD_TX = FD_NP / 4;     // threads across the pairs of a tile (4 pairs each); the other 256 / FD_TX walk the rows

struct FimArgs {
    const float* W0[RABI_MAXK];   // [H0][C + D] weight * mask
    const float* W1[RABI_MAXK];   // [H1][H0]
    const float* Wo[RABI_MAXK];   // [2D][H1]
    const float* bo[RABI_MAXK];   // [2D]
    const float* h0;              // [K][H0][P] primal activations (emitted by k_tc<TC_NLL>)
    const float* h1;              // [K][H1][P]
    const float* y;               // [K][D][P]  theta-side inputs of the transforms
    const float* cdot;            // [C][P]     tangent of the (z-scored) context
    float *h0d, *h1d, *ydot;      // tangent activations / inputs, same shapes as h0, h1, y
    float *a0t, *a0b, *a1t, *a1b; // adjoints of the (tangent, primal) pre-activations
    float *ot, *ob;               // [K][2D][P] adjoints of the (tangent, primal) outputs (mean rows, then log-scale rows)
    float* aux;                   // [K][3][D][P]: exp(clamp(s)), sdot, u
    float* theta_bar;             // [P][D] d phi / d theta
    int K, D, C, H0, H1, P;
    float lo, hi;
};

// Out[m][0..FD_NP) (+)= sum_k W(m, k) In[k][0..FD_NP);  W(m, k) = TRANS ? W[k * ldw + m] : W[m * ldw + k]
template <bool TRANS>
__device__ __forceinline__ void fd_prod(const float* __restrict__ W, int ldw, int M, int Kd, const float* In, float* Out, bool acc) {
    const int tx = threadIdx.x % FD_TX, ty = threadIdx.x / FD_TX;
    for (int m = ty; m < M; m += 256 / FD_TX) {
        float4 a = acc ? *reinterpret_cast<const float4*>(Out + m * FD_NP + 4 * tx) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < Kd; ++k) {
            const float w = __ldg(TRANS ? W + (size_t)k * ldw + m : W + (size_t)m * ldw + k);
            const float4 v = *reinterpret_cast<const float4*>(In + k * FD_NP + 4 * tx);
            a.x = fmaf(w, v.x, a.x);
            a.y = fmaf(w, v.y, a.y);
            a.z = fmaf(w, v.z, a.z);
            a.w = fmaf(w, v.w, a.w);
        }
        *reinterpret_cast<float4*>(Out + m * FD_NP + 4 * tx) = a;
    }
}

__global__ void __launch_bounds__(256) k_fim_dual(const FimArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int K = a.K, D = a.D, C = a.C, H0 = a.H0, H1 = a.H1, P = a.P;
    const int HM = max(H0, H1);
    float* bC = sm;                       // [C][FD_NP]   cdot tile
    float* bX = bC + C * FD_NP;           // [HM][FD_NP]
    float* bY = bX + HM * FD_NP;          // [HM][FD_NP]
    float* bOd = bY + HM * FD_NP;         // [2D][FD_NP]  tangent outputs | tangent-adjoint outputs
    float* bOv = bOd + 2 * D * FD_NP;     // [2D][FD_NP]  primal outputs  | primal-adjoint outputs
    float* bYd = bOv + 2 * D * FD_NP;     // [D][FD_NP]   ydot chain | ut chain
    float* bUb = bYd + D * FD_NP;         // [D][FD_NP]   ub chain
    const int p0 = blockIdx.x * FD_NP, tid = threadIdx.x, T = blockDim.x;
    // element (row r, column c) of a [rows][FD_NP] tile <-> global [rows][P] at pair p0 + c: consecutive threads walk the pairs
    auto gl = [&](const float* base, int r, int c) { return p0 + c < P ? base[(size_t)r * P + p0 + c] : 0.f; };
    auto gs = [&](float* base, int r, int c, float v) { if (p0 + c < P) base[(size_t)r * P + p0 + c] = v; };
    for (int i = tid; i < C * FD_NP; i += T) bC[i] = gl(a.cdot, i / FD_NP, i % FD_NP);
    for (int i = tid; i < D * FD_NP; i += T) bYd[i] = 0.f;
    __syncthreads();
    // ================= tangent density pass: transform K-1 first (theta itself has no tangent)
    for (int k = K - 1; k >= 0; --k) {
        const float* W0 = a.W0[k];
        const int ld0 = C + D;
        fd_prod<false>(W0, ld0, H0, C, bC, bX, false);
        __syncthreads();
        fd_prod<false>(W0 + C, ld0, H0, D, bYd, bX, true);
        __syncthreads();
        for (int i = tid; i < H0 * FD_NP; i += T) {
            const int r = i / FD_NP, c = i % FD_NP;
            const float v = gl(a.h0 + (size_t)k * H0 * P, r, c) > 0.f ? bX[i] : 0.f;
            bX[i] = v;
            gs(a.h0d + (size_t)k * H0 * P, r, c, v);
        }
        for (int i = tid; i < D * FD_NP; i += T) gs(a.ydot + (size_t)k * D * P, i / FD_NP, i % FD_NP, bYd[i]);
        __syncthreads();
        fd_prod<false>(a.W1[k], H0, H1, H0, bX, bY, false);
        __syncthreads();
        for (int i = tid; i < H1 * FD_NP; i += T) {
            const int r = i / FD_NP, c = i % FD_NP;
            const float h = gl(a.h1 + (size_t)k * H1 * P, r, c);
            const float v = h > 0.f ? bY[i] : 0.f;
            bY[i] = v;
            gs(a.h1d + (size_t)k * H1 * P, r, c, v);
            bX[i] = h;                     // primal activations: the outputs (mean, log-scale) are recomputed from them
        }
        __syncthreads();
        fd_prod<false>(a.Wo[k], H1, 2 * D, H1, bY, bOd, false);
        fd_prod<false>(a.Wo[k], H1, 2 * D, H1, bX, bOv, false);
        __syncthreads();
        for (int i = tid; i < D * FD_NP; i += T) {
            const int d = i / FD_NP, c = i % FD_NP;
            const float mu = bOv[i] + __ldg(a.bo[k] + d);
            const float s = bOv[(D + d) * FD_NP + c] + __ldg(a.bo[k] + D + d);
            const float es = expf(fminf(fmaxf(s, a.lo), a.hi));
            const float yv = gl(a.y + (size_t)k * D * P, d, c);
            const float sd = bOd[(D + d) * FD_NP + c], mud = bOd[i], ydin = bYd[i];
            float* aux = a.aux + (size_t)k * 3 * D * P;
            gs(aux, d, c, es);
            gs(aux + (size_t)D * P, d, c, sd);
            gs(aux + (size_t)2 * D * P, d, c, fmaf(es, yv, mu));
            bYd[i] = fmaf(es, fmaf(sd, yv, ydin), mud);     // tangent of u = exp(s) y + mu: the next transform's ydot
        }
        __syncthreads();
    }
    // ================= reverse over the dual program; seed 1 on the tangent of log q = sum sdot - u0 . u0dot
    for (int i = tid; i < D * FD_NP; i += T) {
        const int d = i / FD_NP, c = i % FD_NP;
        bUb[i] = -bYd[i];                                   // adjoint of u0      = -u0dot
        bYd[i] = -gl(a.aux + (size_t)2 * D * P, d, c);      // adjoint of u0dot   = -u0      (aux of transform 0)
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const float* aux = a.aux + (size_t)k * 3 * D * P;
        for (int i = tid; i < D * FD_NP; i += T) {
            const int d = i / FD_NP, c = i % FD_NP;
            const float es = gl(aux, d, c), sd = gl(aux + (size_t)D * P, d, c);
            const float yv = gl(a.y + (size_t)k * D * P, d, c), yd = gl(a.ydot + (size_t)k * D * P, d, c);
            const float ut = bYd[i], ub = bUb[i];
            const float ot_m = ut, ot_s = fmaf(ut * es, yv, 1.f);
            const float ob_m = ub, ob_s = fmaf(ub * es, yv, ut * es * fmaf(sd, yv, yd));
            bOd[i] = ot_m;
            bOd[(D + d) * FD_NP + c] = ot_s;
            bOv[i] = ob_m;
            bOv[(D + d) * FD_NP + c] = ob_s;
            gs(a.ot + (size_t)k * 2 * D * P, d, c, ot_m);
            gs(a.ot + (size_t)k * 2 * D * P, D + d, c, ot_s);
            gs(a.ob + (size_t)k * 2 * D * P, d, c, ob_m);
            gs(a.ob + (size_t)k * 2 * D * P, D + d, c, ob_s);
            bYd[i] = ut * es;                               // direct parts of the adjoints of (ydot, y)
            bUb[i] = fmaf(ub, es, ut * es * sd);
        }
        __syncthreads();
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {           // 0: adjoints of the tangent variables, 1: of the primal ones
            const float* Oin = which == 0 ? bOd : bOv;
            float* a1out = (which == 0 ? a.a1t : a.a1b) + (size_t)k * H1 * P;
            float* a0out = (which == 0 ? a.a0t : a.a0b) + (size_t)k * H0 * P;
            float* chain = which == 0 ? bYd : bUb;
            fd_prod<true>(a.Wo[k], H1, H1, 2 * D, Oin, bX, false);
            __syncthreads();
            for (int i = tid; i < H1 * FD_NP; i += T) {
                const int r = i / FD_NP, c = i % FD_NP;
                const float v = gl(a.h1 + (size_t)k * H1 * P, r, c) > 0.f ? bX[i] : 0.f;
                bX[i] = v;
                gs(a1out, r, c, v);
            }
            __syncthreads();
            fd_prod<true>(a.W1[k], H0, H0, H1, bX, bY, false);
            __syncthreads();
            for (int i = tid; i < H0 * FD_NP; i += T) {
                const int r = i / FD_NP, c = i % FD_NP;
                const float v = gl(a.h0 + (size_t)k * H0 * P, r, c) > 0.f ? bY[i] : 0.f;
                bY[i] = v;
                gs(a0out, r, c, v);
            }
            __syncthreads();
            fd_prod<true>(a.W0[k] + C, C + D, D, H0, bY, chain, true);
            __syncthreads();
        }
    }
    for (int i = tid; i < D * FD_NP; i += T) {
        const int d = i / FD_NP, c = i % FD_NP;
        if (p0 + c < P) a.theta_bar[(size_t)(p0 + c) * D + d] = bUb[i];
    }
}

// cdot[c][p] = 2 w_row[p % B] score[p][c] / std[c]   (rows the reduction dropped carry weight 0 and contribute nothing)
__global__ void k_fim_cdot(const float* __restrict__ score, const float* __restrict__ w_row, const float* __restrict__ zstd,
                           int P, int B, int C, float* __restrict__ cdot) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)P * C) return;
    const int c = (int)(i / P), p = (int)(i % P);
    const float w = w_row[p % B];
    float v = 0.f;
    if (w != 0.f) {
        v = 2.f * w * score[(size_t)p * C + c];
        if (zstd) v /= zstd[c];
    }
    cdot[i] = v;
}

}  // namespace rabi

extern "C" int rabi_fisher_trace_bwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* z_dev,
                                     const float* zs_std_dev, int S, int B, const float* score_dev, const float* w_row_dev,
                                     const float* const* wm_dev, const float* const* bo_dev, float* const* g_weight_dev,
                                     float* const* g_bias_dev, rabi_stream_t s) {
    using namespace rabi;
    if (!h || !h->train) return rabi_fail(h, "rabi_fisher_trace_bwd: call rabi_fisher_trace_fwd on this handle first");
    if (cudaSetDevice(h->device) != cudaSuccess) return rabi_fail(h, "cudaSetDevice failed");
    if (S <= 0 || B <= 0) return 0;
    if (h->L != 2 || h->K > RABI_MAXK) return rabi_fail(h, "rabi_fisher_trace_bwd serves two-hidden-layer conditioners");
    cudaStream_t st = (cudaStream_t)s;
    TrainState* ts = static_cast<TrainState*>(h->train);
    const int K = h->K, D = h->D, C = h->C, H0 = h->H[0], H1 = h->H[1], L = h->L;
    const size_t P = (size_t)S * B;
    const size_t n0 = (size_t)K * H0 * P, n1 = (size_t)K * H1 * P, no = (size_t)K * 2 * D * P, ny = (size_t)K * D * P;
    if (!ts->ws || 2 * n0 + 2 * n1 + no + ny > ts->ws_floats) return rabi_fail(h, "rabi_fisher_trace_bwd: S, B do not match the forward call");
    const float* em_h0 = ts->ws;
    const float* em_h1 = ts->ws + 2 * n0;
    const float* em_y = ts->ws + 2 * n0 + 2 * n1 + no;
    // own buffers: cdot [C][P] | h0d, a0t, a0b [n0] | h1d, a1t, a1b [n1] | ot, ob [no] | ydot [ny] | aux [3 ny] | theta_bar [P D]
    const size_t need = (size_t)C * P + 3 * n0 + 3 * n1 + 2 * no + 4 * ny + P * D + P;
    if (need > ts->fim_floats) {
        if (ts->fim) {
            RABI_CU_OK(h, cudaDeviceSynchronize());
            RABI_CU_OK(h, cudaFree(ts->fim));
            ts->fim = nullptr;
            ts->fim_floats = 0;
        }
        RABI_CU_OK(h, cudaMalloc(&ts->fim, need * sizeof(float)));
        ts->fim_floats = need;
    }
    float* cdot = ts->fim;
    float* h0d = cdot + (size_t)C * P;
    float* a0t = h0d + n0;
    float* a0b = a0t + n0;
    float* h1d = a0b + n0;
    float* a1t = h1d + n1;
    float* a1b = a1t + n1;
    float* ot = a1b + n1;
    float* ob = ot + no;
    float* ydot = ob + no;
    float* aux = ydot + ny;
    float* theta_bar = aux + 3 * ny;
    float* zero_w = theta_bar + P * D;    // [P] dL/dlog q of the sample-path reverse: none
    k_fim_cdot<<<(unsigned)((P * C + 255) / 256), 256, 0, st>>>(score_dev, w_row_dev, zs_std_dev, (int)P, B, C, cdot);
    h->launches++;
    FimArgs fa;
    memset(&fa, 0, sizeof(fa));
    for (int k = 0; k < K; ++k) {
        fa.W0[k] = wm_dev[k * 3 + 0];
        fa.W1[k] = wm_dev[k * 3 + 1];
        fa.Wo[k] = wm_dev[k * 3 + 2];
        fa.bo[k] = bo_dev[k];
    }
    fa.h0 = em_h0;
    fa.h1 = em_h1;
    fa.y = em_y;
    fa.cdot = cdot;
    fa.h0d = h0d;
    fa.h1d = h1d;
    fa.ydot = ydot;
    fa.a0t = a0t;
    fa.a0b = a0b;
    fa.a1t = a1t;
    fa.a1b = a1b;
    fa.ot = ot;
    fa.ob = ob;
    fa.aux = aux;
    fa.theta_bar = theta_bar;
    fa.K = K;
    fa.D = D;
    fa.C = C;
    fa.H0 = H0;
    fa.H1 = H1;
    fa.P = (int)P;
    fa.lo = h->lo;
    fa.hi = h->hi;
    const size_t smem = ((size_t)C + 2 * (size_t)std::max(H0, H1) + 4 * (size_t)D + 2 * (size_t)D) * FD_NP * sizeof(float);
    if (smem > (size_t)h->max_smem_optin) return rabi_fail(h, "rabi_fisher_trace_bwd: conditioner too wide for the dual kernel");
    RABI_CU_OK(h, cudaFuncSetAttribute(k_fim_dual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fim_dual<<<(unsigned)((P + FD_NP - 1) / FD_NP), 256, smem, st>>>(fa);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "k_fim_dual launch failed: %s", cudaGetErrorString(e));
    // ---- stacked weight products (zeroing the gradient buffers), then the path through the samples accumulates on top
    if (!ts->masks) return rabi_fail(h, "rabi_fisher_trace_bwd: masks not uploaded (forward call missing)");
    WgArgs wa;
    memset(&wa, 0, sizeof(wa));
    int nd = 0, tiles = 0;
    auto add = [&](const float* Aop, int M, const float* Bop, int N, int Pn, int ldb, int b_pmajor, int b_mod, float* out, int ldo,
                   const float* mask, float* bias) {
        WgDesc& d = wa.d[nd++];
        d.A = Aop;
        d.B = Bop;
        d.out = out;
        d.mask = mask;
        d.bias = bias;
        d.M = M;
        d.N = N;
        d.P = Pn;
        d.lda = Pn;
        d.ldb = ldb;
        d.ldo = ldo;
        d.b_pmajor = b_pmajor;
        d.b_mod = b_mod;
        d.tile0 = tiles;
        d.tn = (N + WG_TN - 1) / WG_TN;
        d.ksplit = std::max(1, std::min(32, (Pn + 255) / 256));
        tiles += ((M + WG_TM - 1) / WG_TM) * d.tn * d.ksplit;
    };
    if (K * 8 > WG_MAXD) return rabi_fail(h, "rabi_fisher_trace_bwd: too many transforms for one weight-product launch");
    const int Pn = (int)P;
    for (int k = 0; k < K; ++k) {
        const float* m0 = ts->masks + ts->mask_off[(size_t)k * (L + 1) + 0];
        const float* m1 = ts->masks + ts->mask_off[(size_t)k * (L + 1) + 1];
        const float* mo = ts->masks + ts->mask_off[(size_t)k * (L + 1) + 2];
        float* W0 = g_weight_dev[k * (L + 1) + 0];
        float* W1 = g_weight_dev[k * (L + 1) + 1];
        float* Wo = g_weight_dev[k * (L + 1) + 2];
        if (k == 0 && zero_grads(h, g_weight_dev, g_bias_dev, st)) return 1;
        const size_t o0 = (size_t)k * H0 * P, o1 = (size_t)k * H1 * P, oo = (size_t)k * 2 * D * P, oy = (size_t)k * D * P;
        add(a0t + o0, H0, cdot, C, Pn, Pn, 0, 0, W0, C + D, m0, nullptr);
        add(a0b + o0, H0, ctx_dev, C, Pn, C, 1, B, W0, C + D, m0, g_bias_dev[k * (L + 1) + 0]);
        add(a0t + o0, H0, ydot + oy, D, Pn, Pn, 0, 0, W0 + C, C + D, m0 + C, nullptr);
        add(a0b + o0, H0, em_y + oy, D, Pn, Pn, 0, 0, W0 + C, C + D, m0 + C, nullptr);
        add(a1t + o1, H1, h0d + o0, H0, Pn, Pn, 0, 0, W1, H0, m1, nullptr);
        add(a1b + o1, H1, em_h0 + o0, H0, Pn, Pn, 0, 0, W1, H0, m1, g_bias_dev[k * (L + 1) + 1]);
        add(ot + oo, 2 * D, h1d + o1, H1, Pn, Pn, 0, 0, Wo, H1, mo, nullptr);
        add(ob + oo, 2 * D, em_h1 + o1, H1, Pn, Pn, 0, 0, Wo, H1, mo, g_bias_dev[k * (L + 1) + 2]);
    }
    wa.nd = nd;
    k_wgrad<<<tiles, 256, 0, st>>>(wa);
    h->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "k_wgrad launch failed: %s", cudaGetErrorString(e));
    // the path through the reparameterised samples: reverse of the sampling chain seeded with d phi / d theta (overwrites the
    // primal emissions, which the launch above has consumed), accumulating into the same gradient buffers
    RABI_CU_OK(h, cudaMemsetAsync(zero_w, 0, P * sizeof(float), st));
    const int rc = train_bwd(h, TC_RSB, ctx_dev, A_dev, z_dev, zero_w, theta_bar, S, B, nullptr, nullptr, nullptr, nullptr,
                             g_weight_dev, g_bias_dev, st, nullptr, true);
    if (rc == 2) return rabi_fail(h, "rabi_fisher_trace_bwd: shape not served by the sampling-chain reverse");
    return rc;
}

// trace[b] = (1/S) sum_s ||score[s, b, :]||^2 ; one warp per row
__global__ void k_score_trace(const float* __restrict__ score, int S, int B, int C, float* __restrict__ trace) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
        const float* r = score + ((size_t)s * B + row) * C;
        for (int c = lane; c < C; c += 32) acc = fmaf(r[c], r[c], acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) trace[row] = acc / (float)S;
}

// Copies the per-pair vectors the last rabi_maf_log_prob_bwd / rabi_maf_rsample_bwd / rabi_fisher_trace_fwd call left in the
// library's workspace into a caller buffer: [h0 | a0bar : K*H0*P each][h1 | a1bar : K*H1*P each][obar : K*2D*P][y : K*D*P], P = S*B.
extern "C" int rabi_train_emissions(rabi_maf_t* h, int S, int B, float* dst_dev, rabi_stream_t s) {
    if (!h || !h->train) return rabi_fail(h, "rabi_train_emissions: no training-side call was made on this handle");
    rabi::TrainState* ts = static_cast<rabi::TrainState*>(h->train);
    const size_t P = (size_t)S * B;
    const size_t n = (size_t)h->K * P * (2 * (size_t)h->H[0] + 2 * (size_t)h->H[1] + 3 * (size_t)h->D);
    if (!ts->ws || n > ts->ws_floats) return rabi_fail(h, "rabi_train_emissions: S, B do not match the last training-side call");
    if (cudaMemcpyAsync(dst_dev, ts->ws, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)s) != cudaSuccess)
        return rabi_fail(h, "rabi_train_emissions: copy failed");
    return 0;
}

extern "C" int rabi_fisher_trace_fwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* z_dev,
                                     const float* zs_std_dev, int S, int B, float* trace_dev, float* score_dev,
                                     float* theta_dev, rabi_stream_t s) {
    if (!h) return 1;
    if (!h->finalized) return rabi_fail(h, "rabi_maf_finalize_structure was not called");
    if (cudaSetDevice(h->device) != cudaSuccess) return rabi_fail(h, "cudaSetDevice failed");
    if (S <= 0 || B <= 0) return 0;
    if (!score_dev || !theta_dev) return rabi_fail(h, "rabi_fisher_trace_fwd: score_dev [S,B,C] and theta_dev [S,B,D] are required");
    cudaStream_t st = (cudaStream_t)s;
    // 1. theta_s = T(z_s; x): the sampling chain
    int rc = rabi_maf_rsample_fwd(h, A_dev, z_dev, S, B, theta_dev, nullptr, s);
    if (rc) return rc;
    // 2. per pair: density passes + their reverse for L = sum_p log q(theta_p | x_b): d L / d (layer-0 pre-activations)
    const size_t P = (size_t)S * B;
    float* ones = nullptr;   // dL/dlogq = 1: reuse the score buffer's head (overwritten by step 3)
    if ((size_t)h->C < 1) return 1;
    ones = score_dev;
    rabi::k_fill<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(ones, 1.f, P);
    h->launches++;
    float* em_a0 = nullptr;
    rc = rabi::train_bwd(h, rabi::TC_NLL, ctx_dev, A_dev, theta_dev, ones, nullptr, S, B, nullptr, nullptr, nullptr, nullptr, nullptr,
                         nullptr, st, &em_a0);
    if (rc == 2) return rabi_fail(h, "rabi_fisher_trace_fwd serves two-hidden-layer conditioners (up to 120 padded units, D <= 12) "
                                     "without Permute layers");
    if (rc) return rc;
    // 3. score_p = d log q / d x = (1/std) W0c^T a0bar_p, every pair a row of the back-projection
    rc = rabi_maf_ctx_project_bwd(h, em_a0, zs_std_dev, (int)P, score_dev, s);
    if (rc) return rc;
    // 4. tr F(x_b) ~= mean_s ||score_{s,b}||^2
    if (trace_dev) {
        k_score_trace<<<(B + 3) / 4, 128, 0, st>>>(score_dev, S, B, h->C, trace_dev);
        h->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "rabi_fisher_trace_fwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rabi_maf_rsample_bwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* z_dev,
                                    const float* g_theta_dev, const float* g_logq_dev, int S, int B, float* theta_dev,
                                    float* logq_dev, float* g_A_dev, float* const* g_weight_dev, float* const* g_bias_dev,
                                    rabi_stream_t s) {
    if (!h) return 1;
    if (!h->finalized) return rabi_fail(h, "rabi_maf_finalize_structure was not called");
    if (cudaSetDevice(h->device) != cudaSuccess) return rabi_fail(h, "cudaSetDevice failed");
    if (S <= 0 || B <= 0) return 0;
    const int rc = rabi::train_bwd(h, rabi::TC_RSB, ctx_dev, A_dev, z_dev, g_logq_dev, g_theta_dev, S, B, theta_dev, logq_dev,
                                   nullptr, g_A_dev, g_weight_dev, g_bias_dev, (cudaStream_t)s);
    if (rc == 2) return rabi_fail(h, "rabi_maf_rsample_bwd serves two-hidden-layer conditioners (up to 120 padded units, D <= 12) "
                                     "without Permute layers; use the differentiable path for this model");
    return rc;
}

extern "C" int rabi_maf_log_prob_bwd(rabi_maf_t* h, const float* ctx_dev, const float* A_dev, const float* theta_dev,
                                     const float* gout_dev, int S, int B, float* logq_dev, float* g_theta_dev, float* g_A_dev,
                                     float* const* g_weight_dev, float* const* g_bias_dev, rabi_stream_t s) {
    if (!h) return 1;
    if (!h->finalized) return rabi_fail(h, "rabi_maf_finalize_structure was not called");
    if (cudaSetDevice(h->device) != cudaSuccess) return rabi_fail(h, "cudaSetDevice failed");
    if (S <= 0 || B <= 0) return 0;
    const int rc = rabi::rabi_train_log_prob_bwd(h, ctx_dev, A_dev, theta_dev, gout_dev, S, B, logq_dev, g_theta_dev, g_A_dev,
                                                 g_weight_dev, g_bias_dev, (cudaStream_t)s);
    if (rc == 2) return rabi_fail(h, "rabi_maf_log_prob_bwd serves two-hidden-layer conditioners (up to 120 padded units, D <= 12) "
                                     "without Permute layers; use the differentiable path for this model");
    return rc;
}
