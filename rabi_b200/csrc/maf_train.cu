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
                Bs[kk][nn] = (n < d.N && p < p1) ? d.B[(size_t)p * d.ldb + n] : 0.f;
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
};

static int layer_out(const rabi_maf* h, int layer) { return layer < h->L ? h->H[layer] : 2 * h->D; }
static int layer_in(const rabi_maf* h, int layer) { return layer == 0 ? h->C + h->D : h->H[layer - 1]; }

void rabi_train_destroy(rabi_maf* h) {
    TrainState* s = static_cast<TrainState*>(h->train);
    if (!s) return;
    if (s->masks) cudaFree(s->masks);
    if (s->ws) cudaFree(s->ws);
    delete s;
    h->train = nullptr;
}

// Common body of the two training-side entry points.
//   mode TC_NLL: in = theta, gout = dL/dlogq, logq / g_theta optional outputs;
//   mode TC_RSB: in = base noise z, gout = dL/dlogq(own sample), g_theta_in = dL/dtheta_sample, theta_out / logq optional outputs.
// 0 = done, 1 = error, 2 = this shape is not served by the fused training path (the caller keeps its autograd path)
static int train_bwd(rabi_maf* h, int mode, const float* ctx, const float* A, const float* in, const float* gout,
                     const float* g_theta_in, int S, int B, float* theta_out, float* logq, float* g_theta, float* g_A,
                     float* const* g_weight, float* const* g_bias, cudaStream_t st, float** em_a0_out = nullptr) {
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
        for (int l = 0; l <= L; ++l) {
            RABI_CU_OK(h, cudaMemsetAsync(g_weight[k * (L + 1) + l], 0, (size_t)layer_out(h, l) * layer_in(h, l) * sizeof(float), st));
            RABI_CU_OK(h, cudaMemsetAsync(g_bias[k * (L + 1) + l], 0, (size_t)layer_out(h, l) * sizeof(float), st));
        }
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
