// Generic (any K, D, C, hidden) per-pair evaluation of the MAF posterior: one thread owns one
// (observation row, MC sample) pair and keeps its activations in a private shared-memory column.
//
// Reference arithmetic replaced (paths relative to /root/reference/src; Pyro = pyro_ppl 1.8.4):
//   made_full_*  : ConditionalAutoRegressiveNN._forward on [ctx, theta] (one MADE pass) and its input-VJP
//   made_seq_*   : AffineAutoregressive._inverse -- D sequential steps in `permutation` order -- done as ONE
//                  pass-equivalent by evaluating, at step j, only the hidden units of MADE degree j
//   chains       : TransformedDistribution.rsample / log_prob over the K conditioned transforms
//                  (rbi/rbi/models/base.py:417-446, rbi/rbi/utils/distributions.py:464-492)
//   RKL          : general_kl_divergence (rbi/rbi/loss/kl_div.py:38-56) incl. the "free" log p
//                  and the reverse-mode input gradient torch autograd would produce for it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "maf_image.h"

namespace rabi {

enum PairMode { MODE_LOGPROB = 0, MODE_RSAMPLE = 1, MODE_RKL_FWD = 2, MODE_RKL_GRAD = 3 };

struct PairArgs {
    const float* A_p;   // [K,H0,B] projection for the sampled-from distribution (or the only one)
    const float* A_q;   // [K,H0,B] projection for the second distribution (RKL modes)
    const float* in;    // theta (LOGPROB) or z (others), [S,B,D]
    float* out_theta;   // RSAMPLE: [S,B,D]
    float* out_lp;      // LOGPROB: log q(theta); RSAMPLE: log q(own sample) (optional); RKL: log p - log q  [S,B]
    float* gA;          // RKL_GRAD: [K,H0,B] += d loss / d A_p   (atomicAdd; caller zeroes)
    int S, B;
    float w;            // RKL_GRAD: weight of each pair in the loss (= inv_B / S)
};

// Thread-private views into dynamic shared memory (column `tid` of [n][T] arrays).
struct TP {
    float* buf[RABI_MAXL];  // activations / gradient accumulators per hidden layer, [Hpad][T]
    float* vec;             // [NV*D][T]
    uint32_t* bits;         // [NSLOT*WL][T] ReLU masks
    int T, tid, WL;
};

#define RB_BUF(l, i) tp.buf[l][(size_t)(i) * tp.T + tp.tid]
#define RB_VEC(slot, j) tp.vec[((size_t)(slot) * g.D + (j)) * tp.T + tp.tid]

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int pair_nvec(int K) { return 4 * K + 4; }
__host__ __device__ inline int pair_nbuf(int L) { return L < 2 ? 2 : L; }
__host__ __device__ inline size_t pair_smem_floats_per_thread(int K, int D, int L, int Hmax) {
    int WL = (Hmax + 31) / 32;
    return (size_t)pair_nbuf(L) * (round4(Hmax) + 8) + (size_t)pair_nvec(K) * D + (size_t)2 * K * L * WL;
}

__device__ __forceinline__ void bit_put(const TP& tp, int slot, int unit, bool on) {
    uint32_t* p = &tp.bits[((size_t)slot * tp.WL + (unit >> 5)) * tp.T + tp.tid];
    uint32_t m = 1u << (unit & 31);
    *p = on ? (*p | m) : (*p & ~m);
}
__device__ __forceinline__ bool bit_get(const TP& tp, int slot, int unit) {
    return (tp.bits[((size_t)slot * tp.WL + (unit >> 5)) * tp.T + tp.tid] >> (unit & 31)) & 1u;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// acc[r] += sum_{i<e} W[r*ld + i] * h[i]   for 4 consecutive rows (rows beyond the layer exist as zero padding)
__device__ __forceinline__ void dot_rows4(const float* __restrict__ W, int ld, int e, const float* h, int T, int tid,
                                          float acc[4]) {
    int e4 = e & ~3;
    for (int i = 0; i < e4; i += 4) {
        float h0 = h[(size_t)(i + 0) * T + tid], h1 = h[(size_t)(i + 1) * T + tid];
        float h2 = h[(size_t)(i + 2) * T + tid], h3 = h[(size_t)(i + 3) * T + tid];
        float4 a = ldg4(W + i), b = ldg4(W + ld + i), c = ldg4(W + 2 * ld + i), d = ldg4(W + 3 * ld + i);
        acc[0] = fmaf(a.w, h3, fmaf(a.z, h2, fmaf(a.y, h1, fmaf(a.x, h0, acc[0]))));
        acc[1] = fmaf(b.w, h3, fmaf(b.z, h2, fmaf(b.y, h1, fmaf(b.x, h0, acc[1]))));
        acc[2] = fmaf(c.w, h3, fmaf(c.z, h2, fmaf(c.y, h1, fmaf(c.x, h0, acc[2]))));
        acc[3] = fmaf(d.w, h3, fmaf(d.z, h2, fmaf(d.y, h1, fmaf(d.x, h0, acc[3]))));
    }
    for (int i = e4; i < e; ++i) {
        float hv = h[(size_t)i * T + tid];
        acc[0] = fmaf(__ldg(W + i), hv, acc[0]);
        acc[1] = fmaf(__ldg(W + ld + i), hv, acc[1]);
        acc[2] = fmaf(__ldg(W + 2 * ld + i), hv, acc[2]);
        acc[3] = fmaf(__ldg(W + 3 * ld + i), hv, acc[3]);
    }
}

// acc[c] += sum_{v in [va,vb)} W[v*ld + c] * gout[v]   for 4 consecutive columns (W already offset to the column)
__device__ __forceinline__ void tdot_cols4(const float* __restrict__ W, int ld, int va, int vb, const float* gout,
                                           int T, int tid, float acc[4]) {
    int v = va;
    for (; v + 1 < vb; v += 2) {
        float g0 = gout[(size_t)v * T + tid], g1 = gout[(size_t)(v + 1) * T + tid];
        float4 w0 = ldg4(W + (size_t)v * ld), w1 = ldg4(W + (size_t)(v + 1) * ld);
        acc[0] = fmaf(w1.x, g1, fmaf(w0.x, g0, acc[0]));
        acc[1] = fmaf(w1.y, g1, fmaf(w0.y, g0, acc[1]));
        acc[2] = fmaf(w1.z, g1, fmaf(w0.z, g0, acc[2]));
        acc[3] = fmaf(w1.w, g1, fmaf(w0.w, g0, acc[3]));
    }
    if (v < vb) {
        float g0 = gout[(size_t)v * T + tid];
        float4 w0 = ldg4(W + (size_t)v * ld);
        acc[0] = fmaf(w0.x, g0, acc[0]);
        acc[1] = fmaf(w0.y, g0, acc[1]);
        acc[2] = fmaf(w0.z, g0, acc[2]);
        acc[3] = fmaf(w0.w, g0, acc[3]);
    }
}

// (m_j, s_j) of output position j from the last hidden layer, prefix length e.
__device__ __forceinline__ void out_pos(const MafImg& g, const TImg& t, const TP& tp, int j, int e, float& m,
                                        float& s) {
    const float* F = g.f;
    const float* wm = F + t.Wo + (size_t)j * t.ldo;
    const float* ws = F + t.Wo + (size_t)(g.D + j) * t.ldo;
    const int l = g.L - 1;
    float am = __ldg(F + t.bo + j), as = __ldg(F + t.bo + g.D + j);
    int e4 = e & ~3;
    for (int i = 0; i < e4; i += 4) {
        float h0 = RB_BUF(l, i), h1 = RB_BUF(l, i + 1), h2 = RB_BUF(l, i + 2), h3 = RB_BUF(l, i + 3);
        float4 a = ldg4(wm + i), b = ldg4(ws + i);
        am = fmaf(a.w, h3, fmaf(a.z, h2, fmaf(a.y, h1, fmaf(a.x, h0, am))));
        as = fmaf(b.w, h3, fmaf(b.z, h2, fmaf(b.y, h1, fmaf(b.x, h0, as))));
    }
    for (int i = e4; i < e; ++i) {
        float hv = RB_BUF(l, i);
        am = fmaf(__ldg(wm + i), hv, am);
        as = fmaf(__ldg(ws + i), hv, as);
    }
    m = am;
    s = as;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---------------------------------------------------------------------------------------------------------
// One full MADE pass of transform k on the order-space input in vec slot `yin`; writes m/s (raw) to slots.
// bits_slot0 < 0: do not record ReLU masks.
__device__ void made_full_fwd(const MafImg& g, int k, const TP& tp, int yin, const float* __restrict__ A,
                              size_t Astride, int mout, int sout, int bits_slot0) {
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int D = g.D, L = g.L;
    const int H0 = g.H[0];
    for (int u = 0; u < H0; ++u) {
        float pre = A[(size_t)u * Astride];
        int dg = M[t.deg0 + u];
        const float* w = F + t.W0t + (size_t)u * t.ldt;
        for (int j = 0; j < dg; ++j) pre = fmaf(__ldg(w + j), RB_VEC(yin, j), pre);
        bool on = pre > 0.f;
        if (bits_slot0 >= 0) bit_put(tp, bits_slot0, u, on);
        RB_BUF(0, u) = on ? pre : 0.f;
    }
    for (int l = 1; l < L; ++l) {
        const int Hl = g.H[l];
        const int ld = t.ld[l];
        for (int v = 0; v < Hl; v += 4) {
            int nv = min(4, Hl - v);
            int e = M[t.pre[l] + v + nv - 1];
            float acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = __ldg(F + t.b[l] + v + r);
            dot_rows4(F + t.W[l] + (size_t)v * ld, ld, e, tp.buf[l - 1], tp.T, tp.tid, acc);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r < nv) {
                    bool on = acc[r] > 0.f;
                    if (bits_slot0 >= 0) bit_put(tp, bits_slot0 + l, v + r, on);
                    RB_BUF(l, v + r) = on ? acc[r] : 0.f;
                }
            }
        }
    }
    for (int j = 0; j < D; ++j) {
        float m, s;
        out_pos(g, t, tp, j, M[t.preo + j], m, s);
        RB_VEC(mout, j) = m;
        RB_VEC(sout, j) = s;
    }
}

// Input-VJP of one full MADE pass: given mbar/sbar (order space) adds d/d(theta position) into slot `ybar`.
__device__ void made_full_bwd(const MafImg& g, int k, const TP& tp, int mbar, int sbar, int bits_slot0, int ybar) {
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int D = g.D, L = g.L;
    {
        const int l = L - 1, HL = g.H[l];
        for (int v = 0; v < HL; v += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = M[t.firsto + v]; j < D; ++j) {
                float mb = RB_VEC(mbar, j), sb = RB_VEC(sbar, j);
                float4 wm = ldg4(F + t.Wo + (size_t)j * t.ldo + v);
                float4 ws = ldg4(F + t.Wo + (size_t)(D + j) * t.ldo + v);
                acc[0] = fmaf(ws.x, sb, fmaf(wm.x, mb, acc[0]));
                acc[1] = fmaf(ws.y, sb, fmaf(wm.y, mb, acc[1]));
                acc[2] = fmaf(ws.z, sb, fmaf(wm.z, mb, acc[2]));
                acc[3] = fmaf(ws.w, sb, fmaf(wm.w, mb, acc[3]));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (v + c < HL) RB_BUF(l, v + c) = bit_get(tp, bits_slot0 + l, v + c) ? acc[c] : 0.f;
        }
    }
    for (int l = L - 1; l >= 1; --l) {
        const int Hin = g.H[l - 1], Hl = g.H[l], ld = t.ld[l];
        for (int u = 0; u < Hin; u += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            tdot_cols4(F + t.W[l] + u, ld, M[t.first[l] + u], Hl, tp.buf[l], tp.T, tp.tid, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (u + c < Hin) RB_BUF(l - 1, u + c) = bit_get(tp, bits_slot0 + l - 1, u + c) ? acc[c] : 0.f;
        }
    }
    const int H0 = g.H[0];
    for (int j = 0; j < D; ++j) {
        float acc = 0.f;
        for (int u = M[t.first0 + j]; u < H0; ++u) acc = fmaf(__ldg(F + t.W0t + (size_t)u * t.ldt + j), RB_BUF(0, u), acc);
        RB_VEC(ybar, j) += acc;
    }
}

// Sequential inverse of transform k (sampling direction): zin (order space) -> yout, clamped log-scales -> shat.
__device__ void made_seq_fwd(const MafImg& g, int k, const TP& tp, int zin, int yout, int shat,
                             const float* __restrict__ A, size_t Astride, int bits_slot0) {
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int D = g.D, L = g.L;
    for (int j = 0; j < D; ++j) {
        {
            int ua = M[t.grp[0] + j], ub = M[t.grp[0] + j + 1];
            for (int u = ua; u < ub; ++u) {
                float pre = A[(size_t)u * Astride];
                const float* w = F + t.W0t + (size_t)u * t.ldt;
                for (int jj = 0; jj < j; ++jj) pre = fmaf(__ldg(w + jj), RB_VEC(yout, jj), pre);
                bool on = pre > 0.f;
                if (bits_slot0 >= 0) bit_put(tp, bits_slot0, u, on);
                RB_BUF(0, u) = on ? pre : 0.f;
            }
        }
        for (int l = 1; l < L; ++l) {
            int va = M[t.grp[l] + j], vb = M[t.grp[l] + j + 1];
            int e = M[t.grp[l - 1] + j + 1];
            const int ld = t.ld[l];
            for (int v = va; v < vb; v += 4) {
                int nv = min(4, vb - v);
                float acc[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = __ldg(F + t.b[l] + v + r);
                dot_rows4(F + t.W[l] + (size_t)v * ld, ld, e, tp.buf[l - 1], tp.T, tp.tid, acc);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r < nv) {
                        bool on = acc[r] > 0.f;
                        if (bits_slot0 >= 0) bit_put(tp, bits_slot0 + l, v + r, on);
                        RB_BUF(l, v + r) = on ? acc[r] : 0.f;
                    }
                }
            }
        }
        float m, s;
        out_pos(g, t, tp, j, M[t.grp[L - 1] + j + 1], m, s);
        float sh = clampf(s, g.lo, g.hi);
        RB_VEC(shat, j) = sh;
        RB_VEC(yout, j) = (RB_VEC(zin, j) - m) * expf(-sh);
    }
}

// Reverse-mode through made_seq_fwd.  In: ybar (order space, grad wrt outputs y; consumed), yout, shat, wld =
// coefficient of sum_j shat_j in the loss.  Out: zbar (grad wrt z inputs), gA += grad wrt layer-0 pre-activations.
__device__ void made_seq_bwd(const MafImg& g, int k, const TP& tp, int ybar, int yout, int shat, float wld,
                             int bits_slot0, int zbar, float* __restrict__ gA, size_t Astride) {
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int D = g.D, L = g.L;
    for (int l = 0; l < L; ++l) {
        int n = round4(g.H[l]) + 4;
        for (int i = 0; i < n; ++i) RB_BUF(l, i) = 0.f;
    }
    for (int j = D - 1; j >= 0; --j) {
        float yb = RB_VEC(ybar, j);
        float zb = yb * expf(-RB_VEC(shat, j));
        RB_VEC(zbar, j) = zb;
        float mb = -zb;
        float sb = fmaf(-yb, RB_VEC(yout, j), wld);
        {
            const int l = L - 1;
            int e = M[t.grp[l] + j + 1];
            const float* wm = F + t.Wo + (size_t)j * t.ldo;
            const float* ws = F + t.Wo + (size_t)(D + j) * t.ldo;
            for (int v = 0; v < e; v += 4) {  // columns >= e are zero-masked; buffers are padded
                float4 a = ldg4(wm + v), b = ldg4(ws + v);
                RB_BUF(l, v + 0) += fmaf(b.x, sb, a.x * mb);
                RB_BUF(l, v + 1) += fmaf(b.y, sb, a.y * mb);
                RB_BUF(l, v + 2) += fmaf(b.z, sb, a.z * mb);
                RB_BUF(l, v + 3) += fmaf(b.w, sb, a.w * mb);
            }
        }
        for (int l = L - 1; l >= 1; --l) {
            int va = M[t.grp[l] + j], vb = M[t.grp[l] + j + 1];
            for (int v = va; v < vb; ++v)
                if (!bit_get(tp, bits_slot0 + l, v)) RB_BUF(l, v) = 0.f;
            int e = M[t.grp[l - 1] + j + 1];
            const int ld = t.ld[l];
            for (int u = 0; u < e; u += 4) {
                float acc[4] = {RB_BUF(l - 1, u), RB_BUF(l - 1, u + 1), RB_BUF(l - 1, u + 2), RB_BUF(l - 1, u + 3)};
                tdot_cols4(F + t.W[l] + u, ld, va, vb, tp.buf[l], tp.T, tp.tid, acc);
                RB_BUF(l - 1, u) = acc[0];
                RB_BUF(l - 1, u + 1) = acc[1];
                RB_BUF(l - 1, u + 2) = acc[2];
                RB_BUF(l - 1, u + 3) = acc[3];
            }
        }
        int ua = M[t.grp[0] + j], ub = M[t.grp[0] + j + 1];
        for (int u = ua; u < ub; ++u) {
            float g0 = bit_get(tp, bits_slot0, u) ? RB_BUF(0, u) : 0.f;
            if (gA) atomicAdd(gA + (size_t)u * Astride, g0);
            const float* w = F + t.W0t + (size_t)u * t.ldt;
            for (int jj = 0; jj < j; ++jj) RB_VEC(ybar, jj) = fmaf(__ldg(w + jj), g0, RB_VEC(ybar, jj));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void k_pairs(const MafImg g, const PairArgs a) {
    extern __shared__ __align__(16) float smem[];
    TP tp;
    tp.T = blockDim.x;
    tp.tid = threadIdx.x;
    tp.WL = (g.Hmax + 31) / 32;
    const int K = g.K, D = g.D, L = g.L;
    {
        const int NB = pair_nbuf(L);
        const size_t bufsz = (size_t)(round4(g.Hmax) + 8) * tp.T;
        float* p = smem;
        for (int l = 0; l < RABI_MAXL; ++l) tp.buf[l] = p + (size_t)(l < NB ? l : 0) * bufsz;
        p += (size_t)NB * bufsz;
        tp.vec = p;
        p += (size_t)pair_nvec(K) * D * tp.T;
        tp.bits = reinterpret_cast<uint32_t*>(p);
    }
    const long long npairs = (long long)a.S * a.B;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int b = (int)(p % a.B);
    const size_t Astride = (size_t)a.B;
    const int H0 = g.H[0];
    const int* M = g.m;
    const float HALF_LOG_2PI = 0.91893853320467274178f;

    // vec slots
    const int SV = 0;            // v_k, k = 0..K (variable space; v_0 = z, v_K = theta)
    const int SU = K + 1;        // u_k, k = 0..K-1 (q chain, variable space; u_K == v_K)
    const int SHP = 2 * K + 1;   // clamped log-scales of the sampling pass, order space of transform k
    const int SHQ = 3 * K + 1;   // clamped log-scales of the density pass
    const int TA = 4 * K + 1, TB = 4 * K + 2, TC = 4 * K + 3;
    const bool record = (MODE == MODE_RKL_GRAD);
    auto bslot = [&](int pass, int k) { return record ? (pass * K + k) * L : -1; };

    // zero activation buffers once (prefix tails may be read against zero weights)
    for (int l = 0; l < pair_nbuf(L); ++l) {
        int n = round4(g.Hmax) + 8;
        for (int i = 0; i < n; ++i) RB_BUF(l, i) = 0.f;
    }
    if (record) {
        int nw = 2 * K * L * tp.WL;
        for (int i = 0; i < nw; ++i) tp.bits[(size_t)i * tp.T + tp.tid] = 0u;
    }

    float logp = 0.f;  // log-density under the sampled-from distribution
    if (MODE != MODE_LOGPROB) {
        // ---- sampling chain: v_0 = z, v_{k+1} = T_k(v_k)
        float ssq = 0.f;
        for (int d = 0; d < D; ++d) {
            float z = a.in[(size_t)p * D + d];
            RB_VEC(SV, d) = z;
            ssq = fmaf(z, z, ssq);
        }
        logp = -0.5f * ssq - (float)D * HALF_LOG_2PI;
        for (int k = 0; k < K; ++k) {
            const TImg& t = g.t[k];
            for (int j = 0; j < D; ++j) RB_VEC(TA, j) = RB_VEC(SV + k, M[t.perm + j]);
            made_seq_fwd(g, k, tp, TA, TB, SHP + k, a.A_p + (size_t)k * H0 * a.B + b, Astride, bslot(0, k));
            float sl = 0.f;
            for (int j = 0; j < D; ++j) {
                RB_VEC(SV + k + 1, M[t.permy + j]) = RB_VEC(TB, j);
                sl += RB_VEC(SHP + k, j);
            }
            logp += sl;
        }
        if (MODE == MODE_RSAMPLE) {
            for (int d = 0; d < D; ++d) a.out_theta[(size_t)p * D + d] = RB_VEC(SV + K, d);
            if (a.out_lp) a.out_lp[p] = logp;
            return;
        }
    } else {
        for (int d = 0; d < D; ++d) RB_VEC(SV + K, d) = a.in[(size_t)p * D + d];
    }

    // ---- density chain under A_q (A_p for LOGPROB): u_K = theta, u_k = exp(s) * u_{k+1} + m
    const float* Aq = (MODE == MODE_LOGPROB) ? a.A_p : a.A_q;
    float logq = 0.f;
    for (int k = K - 1; k >= 0; --k) {
        const TImg& t = g.t[k];
        const int src = (k == K - 1) ? SV + K : SU + k + 1;
        for (int j = 0; j < D; ++j) RB_VEC(TA, j) = RB_VEC(src, M[t.permy + j]);
        made_full_fwd(g, k, tp, TA, Aq + (size_t)k * H0 * a.B + b, Astride, TB, TC, bslot(1, k));
        float sl = 0.f;
        for (int j = 0; j < D; ++j) {
            float sh = clampf(RB_VEC(TC, j), g.lo, g.hi);
            RB_VEC(SHQ + k, j) = sh;
            sl += sh;
            RB_VEC(SU + k, M[t.perm + j]) = fmaf(expf(sh), RB_VEC(TA, j), RB_VEC(TB, j));
        }
        logq += sl;
    }
    {
        float ssq = 0.f;
        for (int d = 0; d < D; ++d) {
            float u = RB_VEC(SU, d);
            ssq = fmaf(u, u, ssq);
        }
        logq += -0.5f * ssq - (float)D * HALF_LOG_2PI;
    }
    if (MODE == MODE_LOGPROB) {
        a.out_lp[p] = logq;
        return;
    }
    if (a.out_lp) a.out_lp[p] = logp - logq;
    if (MODE != MODE_RKL_GRAD) return;

    // ---- reverse mode.  loss contribution of this pair: w * (log p - log q)
    const float w = a.w;
    // seed: d(-w log q)/d u_0 = w * u_0   (kept in slot TB in variable space)
    // We reuse SU+0 as the running gradient in variable space (u_0's value is consumed first).
    for (int d = 0; d < D; ++d) RB_VEC(SU, d) = w * RB_VEC(SU, d);
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        const int src = (k == K - 1) ? SV + K : SU + k + 1;  // u_{k+1}: input of transform k in the density chain
        // order-space: TA = ubar (-> becomes ybar accumulator), TB = mbar, TC = sbar
        for (int j = 0; j < D; ++j) {
            float ub = RB_VEC(SU, M[t.perm + j]);
            float es = expf(RB_VEC(SHQ + k, j));
            float yin = RB_VEC(src, M[t.permy + j]);
            RB_VEC(TB, j) = ub;
            RB_VEC(TC, j) = fmaf(ub * es, yin, -w);
            RB_VEC(TA, j) = ub * es;
        }
        made_full_bwd(g, k, tp, TB, TC, bslot(1, k), TA);
        for (int j = 0; j < D; ++j) RB_VEC(SU, M[t.permy + j]) = RB_VEC(TA, j);
    }
    // SU now holds d loss / d theta (through -log q).  Back through the sampling chain (+w * sum shat^p).
    for (int k = K - 1; k >= 0; --k) {
        const TImg& t = g.t[k];
        for (int j = 0; j < D; ++j) {
            RB_VEC(TA, j) = RB_VEC(SU, M[t.permy + j]);        // ybar (order space)
            RB_VEC(TB, j) = RB_VEC(SV + k + 1, M[t.permy + j]);  // y values
        }
        made_seq_bwd(g, k, tp, TA, TB, SHP + k, w, bslot(0, k), TC, a.gA + (size_t)k * H0 * a.B + b, Astride);
        for (int j = 0; j < D; ++j) RB_VEC(SU, M[t.perm + j]) = RB_VEC(TC, j);
    }
}

}  // namespace rabi
