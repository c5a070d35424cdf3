// Small-batch variant: ONE WARP per (row, sample) pair.
//
// The thread-per-pair kernels (maf_fast.cuh, maf_deep.cuh, maf_pairs.cuh) run a launch in the time ONE thread needs for
// its pair (~0.4 ms for the attack gradient at lotka_volterra shapes), however few pairs there are.  Training-time
// attacks (TRADES, adversarial training: batch 512, one MC sample, 20 iterations per step; rbi/rbi/defenses/
// trades_regularizers.py:58-64, adversarial_training.py:34-50) have 512 pairs per launch: 99 % of the lanes idle and the
// inner attack dominates the step.  Here the 32 lanes of a warp split the units of every layer of one pair:
//   * forward layers: lane <-> output unit (row of W, 128-bit loads along the row; with ld = 100 the eight lanes of a
//     quarter-warp hit disjoint banks), previous activations broadcast from the warp's shared-memory buffer;
//   * reverse layers: lane <-> input unit (column of W, consecutive lanes read consecutive addresses), cotangents
//     broadcast;
//   * ReLU masks by warp ballot (full passes) / shared-memory atomics (sequential passes), chain vectors in shared memory.
// Same arithmetic and the same one-pass-equivalent sequential sampling as maf_pairs.cuh; any L, D <= 32.
// Weights come straight from the L2/L1-resident packed image (a handful of warps per SM: no staging needed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "maf_image.h"
#include "maf_pairs.cuh"

namespace rabi {

enum WarpMode { WARP_RKL_GRAD = 0, WARP_FKL_GRAD = 1 };

struct WP {
    float* buf[RABI_MAXL];  // [Hpad] activations / cotangents per hidden layer
    float* vec;             // [NV][D]
    uint32_t* bits;         // [2*K*L][WL]
    int lane, WL;
};

__host__ __device__ inline int warp_nvec(int K) { return 4 * K + 6; }
__host__ __device__ inline size_t warp_smem_floats(int K, int D, int L, int Hmax) {
    const int WL = (Hmax + 31) / 32;
    return (size_t)L * (round4(Hmax) + 8) + (size_t)warp_nvec(K) * D + (size_t)2 * K * L * WL + 8;
}

#define WB(l, i) wp.buf[l][i]
#define WV(slot, j) wp.vec[(slot) * D + (j)]
#define WBITS(slot, word) wp.bits[(slot) * wp.WL + (word)]
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// dot of a weight row with a shared-memory vector over [0, e) (e rounded up to 4: masked weights and padding are zero)
__device__ __forceinline__ float row_dot(const float* __restrict__ w, const float* h, int e, float acc) {
    const int e4 = (e + 3) & ~3;
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four independent chains: a lone warp has no other warp to hide FMA latency
#pragma unroll 4
    for (int i = 0; i < e4; i += 4) {
        const float4 ww = *reinterpret_cast<const float4*>(w + i);
        const float4 hh = *reinterpret_cast<const float4*>(h + i);
        acc = fmaf(ww.x, hh.x, acc);
        a1 = fmaf(ww.y, hh.y, a1);
        a2 = fmaf(ww.z, hh.z, a2);
        a3 = fmaf(ww.w, hh.w, a3);
    }
    return (acc + a1) + (a2 + a3);
}

// sum_{v in [va, vb)} w[v*ld] * g[v]  (one column of a weight matrix against a shared-memory vector), four chains
__device__ __forceinline__ float col_dot(const float* __restrict__ w, int ld, const float* g, int va, int vb) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int v = va;
#pragma unroll 2
    for (; v + 3 < vb; v += 4) {
        a0 = fmaf(w[(size_t)v * ld], g[v], a0);
        a1 = fmaf(w[(size_t)(v + 1) * ld], g[v + 1], a1);
        a2 = fmaf(w[(size_t)(v + 2) * ld], g[v + 2], a2);
        a3 = fmaf(w[(size_t)(v + 3) * ld], g[v + 3], a3);
    }
    for (; v < vb; ++v) a0 = fmaf(w[(size_t)v * ld], g[v], a0);
    return (a0 + a1) + (a2 + a3);
}

// ---- one full MADE pass (density direction).  In: order-space input in slot yin.  Out: raw m / s in slots mout / sout.
__device__ void w_full_fwd(const MafImg& g, const float* __restrict__ F, int k, const WP& wp, int yin, const float* __restrict__ A, size_t Astride,
                           int mout, int sout, int bslot0) {
    const TImg& t = g.t[k];
    const int* M = g.m;
    const int D = g.D, L = g.L, H0 = g.H[0];
    for (int r = 0; r * 32 < ((H0 + 3) & ~3); ++r) {
        const int u = r * 32 + wp.lane;
        float pre = 0.f;
        if (u < H0) {
            pre = A[(size_t)u * Astride];
            const int dg = M[t.deg0 + u];
            const float* w = F + t.W0t + (size_t)u * t.ldt;
            for (int j = 0; j < dg; ++j) pre = fmaf(w[j], WV(yin, j), pre);
        }
        const bool on = u < H0 && pre > 0.f;
        if (u < ((H0 + 3) & ~3)) WB(0, u) = on ? pre : 0.f;
        const unsigned m = __ballot_sync(FULL, on);
        if (bslot0 >= 0 && wp.lane == 0) WBITS(bslot0, r) = m;
    }
    __syncwarp();
    for (int l = 1; l < L; ++l) {
        const int Hl = g.H[l], Hlp = (Hl + 3) & ~3, ld = t.ld[l];
        for (int r = 0; r * 32 < Hlp; ++r) {
            const int v = r * 32 + wp.lane;
            float acc = 0.f;
            if (v < Hl) acc = row_dot(F + t.W[l] + (size_t)v * ld, wp.buf[l - 1], M[t.pre[l] + v], *(F + t.b[l] + v));
            const bool on = v < Hl && acc > 0.f;
            if (v < Hlp) WB(l, v) = on ? acc : 0.f;
            const unsigned m = __ballot_sync(FULL, on);
            if (bslot0 >= 0 && wp.lane == 0) WBITS(bslot0 + l, r) = m;
        }
        __syncwarp();
    }
    for (int r = wp.lane; r < 2 * D; r += 32) {
        const int j = r < D ? r : r - D;
        const float acc = row_dot(F + t.Wo + (size_t)r * t.ldo, wp.buf[L - 1], M[t.preo + j], *(F + t.bo + r));
        if (r < D) WV(mout, j) = acc;
        else WV(sout, j) = acc;
    }
    __syncwarp();
}

// ---- input-VJP of one full MADE pass.  In: mbar / sbar slots.  ybar slot += d/d(theta positions); gA (optional) +=
// d/d(layer-0 pre-activations) (the forward-KL attack differentiates this pass).
__device__ void w_full_bwd(const MafImg& g, const float* __restrict__ F, int k, const WP& wp, int mbar, int sbar, int bslot0, int ybar,
                           float* __restrict__ gA, size_t Astride) {
    const TImg& t = g.t[k];
    const int* M = g.m;
    const int D = g.D, L = g.L, H0 = g.H[0];
    {
        const int l = L - 1, HL = g.H[l], HLp = (HL + 3) & ~3;
        for (int r = 0; r * 32 < HLp; ++r) {
            const int v = r * 32 + wp.lane;
            float acc = 0.f;
            if (v < HL) {
                for (int j = M[t.firsto + v]; j < D; ++j) {
                    acc = fmaf(*(F + t.Wo + (size_t)j * t.ldo + v), WV(mbar, j), acc);
                    acc = fmaf(*(F + t.Wo + (size_t)(D + j) * t.ldo + v), WV(sbar, j), acc);
                }
            }
            const bool on = v < HL && ((WBITS(bslot0 + l, r) >> wp.lane) & 1u);
            if (v < HLp) WB(l, v) = on ? acc : 0.f;
        }
        __syncwarp();
    }
    for (int l = L - 1; l >= 1; --l) {
        const int Hin = g.H[l - 1], Hinp = (Hin + 3) & ~3, Hl = g.H[l], ld = t.ld[l];
        for (int r = 0; r * 32 < Hinp; ++r) {
            const int u = r * 32 + wp.lane;
            float acc = 0.f;
            if (u < Hin) acc = col_dot(F + t.W[l] + u, ld, wp.buf[l], M[t.first[l] + u], Hl);
            const bool on = u < Hin && ((WBITS(bslot0 + l - 1, r) >> wp.lane) & 1u);
            if (u < Hinp) WB(l - 1, u) = on ? acc : 0.f;
        }
        __syncwarp();
    }
    if (gA)
        for (int u = wp.lane; u < H0; u += 32) atomicAdd(gA + (size_t)u * Astride, WB(0, u));
    for (int j = wp.lane; j < D; j += 32) {
        WV(ybar, j) += col_dot(F + t.W0t + j, t.ldt, wp.buf[0], M[t.first0 + j], H0);
    }
    __syncwarp();
}

// ---- sequential inverse of transform k (sampling direction): zin (order space) -> yout, clamped log-scales -> shat
__device__ void w_seq_fwd(const MafImg& g, const float* __restrict__ F, int k, const WP& wp, int zin, int yout, int shat, const float* __restrict__ A,
                          size_t Astride, int bslot0) {
    const TImg& t = g.t[k];
    const int* M = g.m;
    const int D = g.D, L = g.L;
    for (int j = 0; j < D; ++j) {
        {
            const int ua = M[t.grp[0] + j], ub = M[t.grp[0] + j + 1];
            for (int u = ua + wp.lane; u < ub; u += 32) {
                float pre = A[(size_t)u * Astride];
                const float* w = F + t.W0t + (size_t)u * t.ldt;
                for (int jj = 0; jj < j; ++jj) pre = fmaf(w[jj], WV(yout, jj), pre);
                const bool on = pre > 0.f;
                if (bslot0 >= 0 && on) atomicOr(&WBITS(bslot0, u >> 5), 1u << (u & 31));
                WB(0, u) = on ? pre : 0.f;
            }
        }
        __syncwarp();
        for (int l = 1; l < L; ++l) {
            const int va = M[t.grp[l] + j], vb = M[t.grp[l] + j + 1], e = M[t.grp[l - 1] + j + 1], ld = t.ld[l];
            for (int v = va + wp.lane; v < vb; v += 32) {
                const float acc = row_dot(F + t.W[l] + (size_t)v * ld, wp.buf[l - 1], e, *(F + t.b[l] + v));
                const bool on = acc > 0.f;
                if (bslot0 >= 0 && on) atomicOr(&WBITS(bslot0 + l, v >> 5), 1u << (v & 31));
                WB(l, v) = on ? acc : 0.f;
            }
            __syncwarp();
        }
        const int e = M[t.grp[L - 1] + j + 1];
        float am = 0.f, as = 0.f;
        const float* wm = F + t.Wo + (size_t)j * t.ldo;
        const float* ws = F + t.Wo + (size_t)(D + j) * t.ldo;
        for (int i = wp.lane; i < e; i += 32) {
            const float hv = WB(L - 1, i);
            am = fmaf(wm[i], hv, am);
            as = fmaf(ws[i], hv, as);
        }
        am = warp_sum(am) + *(F + t.bo + j);
        as = warp_sum(as) + *(F + t.bo + D + j);
        if (wp.lane == 0) {
            const float sh = clampf(as, g.lo, g.hi);
            WV(shat, j) = sh;
            WV(yout, j) = (WV(zin, j) - am) * expf(-sh);
        }
        __syncwarp();
    }
}

// ---- reverse of w_seq_fwd, pull style.  In: ybar (consumed), yout, shat, wld = coefficient of sum_j shat_j in the loss.
// Out: zbar; gA += d/d(layer-0 pre-activations).  mb / sb: scratch slots.
__device__ void w_seq_bwd(const MafImg& g, const float* __restrict__ F, int k, const WP& wp, int ybar, int yout, int shat, float wld, int bslot0, int zbar,
                          int mb, int sb, float* __restrict__ gA, size_t Astride) {
    const TImg& t = g.t[k];
    const int* M = g.m;
    const int D = g.D, L = g.L;
    for (int j = wp.lane; j < D; j += 32) WV(mb, j) = WV(sb, j) = 0.f;
    __syncwarp();
    for (int j = D - 1; j >= 0; --j) {
        if (wp.lane == 0) {
            const float yb = WV(ybar, j);
            const float zb = yb * expf(-WV(shat, j));
            WV(zbar, j) = zb;
            WV(mb, j) = -zb;
            WV(sb, j) = fmaf(-yb, WV(yout, j), wld);
        }
        __syncwarp();
        for (int l = L - 1; l >= 0; --l) {
            const int ua = M[t.grp[l] + j], ub = M[t.grp[l] + j + 1];
            for (int u = ua + wp.lane; u < ub; u += 32) {
                float acc = 0.f;
                if (l == L - 1) {
                    for (int jj = j; jj < D; ++jj) {
                        acc = fmaf(*(F + t.Wo + (size_t)jj * t.ldo + u), WV(mb, jj), acc);
                        acc = fmaf(*(F + t.Wo + (size_t)(D + jj) * t.ldo + u), WV(sb, jj), acc);
                    }
                } else {
                    acc = col_dot(F + t.W[l + 1] + u, t.ld[l + 1], wp.buf[l + 1], M[t.grp[l + 1] + j], g.H[l + 1]);
                }
                const bool on = (WBITS(bslot0 + l, u >> 5) >> (u & 31)) & 1u;
                WB(l, u) = on ? acc : 0.f;
            }
            __syncwarp();
        }
        {
            const int ua = M[t.grp[0] + j], ub = M[t.grp[0] + j + 1];
            if (gA)
                for (int u = ua + wp.lane; u < ub; u += 32) atomicAdd(gA + (size_t)u * Astride, WB(0, u));
            for (int jj = wp.lane; jj < j; jj += 32) {
                WV(ybar, jj) += col_dot(F + t.W0t + jj, t.ldt, wp.buf[0], ua, ub);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// MODE WARP_RKL_GRAD: loss w * (log p - log q), samples from A_p, gradient w.r.t. A_p (both passes).
// MODE WARP_FKL_GRAD: samples from A_p (the fixed target), loss -w * log q under A_q, gradient w.r.t. A_q.
// ``wblock`` > 0: the weights of ONE transform ([t.W0t, t.bo + ...), wblock floats, the same size for every transform) are
// staged in shared memory in front of the per-warp state and shared by the warps of the CTA (L2 latency on every weight
// load is what bounds a lone warp otherwise); 0: read the image in global memory (models too large to stage).
// ``wall``: all K transforms are resident (staged once); ``nmeta`` > 0: the index tables (prefix lengths, group bounds:
// loop bounds on the critical path) are staged as well.
template <int MODE>
__global__ void __launch_bounds__(256) k_warp(const MafImg gin, const PairArgs a, const int wblock, const int wall,
                                              const int nmeta) {
    extern __shared__ __align__(16) float smem_all[];
    float* const Ws = smem_all;
    int* const Ms = reinterpret_cast<int*>(smem_all + (size_t)wblock * (wall ? gin.K : 1));
    float* const smem = reinterpret_cast<float*>(Ms + nmeta);
    MafImg g = gin;
    if (nmeta > 0) {
        for (int i = threadIdx.x; i < nmeta; i += blockDim.x) Ms[i] = __ldg(gin.m + i);
        g.m = Ms;
    }
    if (wall) {
        for (int k = 0; k < gin.K; ++k) {
            const float4* src = reinterpret_cast<const float4*>(gin.f + gin.t[k].W0t);
            float4* dst = reinterpret_cast<float4*>(Ws + (size_t)k * wblock);
            for (int i = threadIdx.x; i < (wblock >> 2); i += blockDim.x) dst[i] = __ldg(src + i);
        }
    }
    if (wall || nmeta > 0) __syncthreads();
    const int K = g.K, D = g.D, L = g.L, H0 = g.H[0];
    const int wpc = blockDim.x >> 5, wid = threadIdx.x >> 5;
    WP wp;
    wp.lane = threadIdx.x & 31;
    wp.WL = (g.Hmax + 31) / 32;
    const size_t per = (warp_smem_floats(K, D, L, g.Hmax) + 3) & ~(size_t)3;
    float* base = smem + (size_t)wid * per;
    const int bufsz = round4(g.Hmax) + 8;
    for (int l = 0; l < RABI_MAXL; ++l) wp.buf[l] = base + (size_t)(l < L ? l : 0) * bufsz;
    wp.vec = base + (size_t)L * bufsz;
    wp.bits = reinterpret_cast<uint32_t*>(wp.vec + (size_t)warp_nvec(K) * D);
    const long long npairs = (long long)a.S * a.B;
    const long long p_raw = (long long)blockIdx.x * wpc + wid;
    const bool live = p_raw < npairs;   // idle warps of the last CTA still take part in the CTA-wide weight staging
    const long long p = live ? p_raw : npairs - 1;
    const int b = (int)(p % a.B);
    int resident = -1;
    auto weights = [&](int k) -> const float* {  // base pointer such that base + (image offset) addresses transform k
        if (wblock == 0) return g.f;
        if (wall) return Ws + (size_t)k * wblock - g.t[k].W0t;
        if (resident != k) {
            __syncthreads();
            const float4* src = reinterpret_cast<const float4*>(g.f + g.t[k].W0t);
            float4* dst = reinterpret_cast<float4*>(Ws);
            for (int i = threadIdx.x; i < (wblock >> 2); i += blockDim.x) dst[i] = __ldg(src + i);
            __syncthreads();
            resident = k;
        }
        return Ws - g.t[k].W0t;
    };
    const size_t Astride = (size_t)a.B;
    const int* M = g.m;
    const float HALF_LOG_2PI = 0.91893853320467274178f;
    const int SV = 0, SU = K + 1, SHP = 2 * K + 1, SHQ = 3 * K + 1;
    const int TA = 4 * K + 1, TB = 4 * K + 2, TC = 4 * K + 3, TD = 4 * K + 4, TE = 4 * K + 5;
    auto bslot = [&](int pass, int k) { return (pass * K + k) * L; };
    for (int i = wp.lane; i < (int)per; i += 32) base[i] = 0.f;  // activations, vectors and ReLU words
    __syncwarp();
    const int lane = wp.lane;

    // ---- sampling chain under A_p
    float ssq = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float z = a.in[(size_t)p * D + d];
        WV(SV, d) = z;
        ssq = fmaf(z, z, ssq);
    }
    float logp = -0.5f * warp_sum(ssq) - (float)D * HALF_LOG_2PI;
    __syncwarp();
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        for (int j = lane; j < D; j += 32) WV(TA, j) = WV(SV + k, M[t.perm + j]);
        __syncwarp();
        w_seq_fwd(g, weights(k), k, wp, TA, TB, SHP + k, a.A_p + (size_t)k * H0 * a.B + b, Astride,
                  MODE == WARP_RKL_GRAD ? bslot(0, k) : -1);
        float sl = 0.f;
        for (int j = lane; j < D; j += 32) {
            WV(SV + k + 1, M[t.permy + j]) = WV(TB, j);
            sl += WV(SHP + k, j);
        }
        logp += warp_sum(sl);
        __syncwarp();
    }
    // ---- density chain under A_q
    float logq = 0.f;
    for (int k = K - 1; k >= 0; --k) {
        const TImg& t = g.t[k];
        const int src = (k == K - 1) ? SV + K : SU + k + 1;
        for (int j = lane; j < D; j += 32) WV(TA, j) = WV(src, M[t.permy + j]);
        __syncwarp();
        w_full_fwd(g, weights(k), k, wp, TA, a.A_q + (size_t)k * H0 * a.B + b, Astride, TB, TC, bslot(1, k));
        float sl = 0.f;
        for (int j = lane; j < D; j += 32) {
            const float sh = clampf(WV(TC, j), g.lo, g.hi);
            WV(SHQ + k, j) = sh;
            sl += sh;
            WV(SU + k, M[t.perm + j]) = fmaf(expf(sh), WV(TA, j), WV(TB, j));
        }
        logq += warp_sum(sl);
        __syncwarp();
    }
    {
        float s2 = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float u = WV(SU, d);
            s2 = fmaf(u, u, s2);
        }
        logq += -0.5f * warp_sum(s2) - (float)D * HALF_LOG_2PI;
    }
    if (a.out_lp && lane == 0 && live) a.out_lp[p] = logp - logq;
    // ---- reverse mode
    const float w = a.w;
    for (int d = lane; d < D; d += 32) WV(SU, d) = w * WV(SU, d);  // seed: d(-w log q)/d u_0
    __syncwarp();
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        const int src = (k == K - 1) ? SV + K : SU + k + 1;
        for (int j = lane; j < D; j += 32) {
            const float ub = WV(SU, M[t.perm + j]);
            const float es = expf(WV(SHQ + k, j));
            const float yin = WV(src, M[t.permy + j]);
            WV(TB, j) = ub;
            WV(TC, j) = fmaf(ub * es, yin, -w);
            WV(TA, j) = ub * es;
        }
        __syncwarp();
        w_full_bwd(g, weights(k), k, wp, TB, TC, bslot(1, k), TA,
                   (MODE == WARP_FKL_GRAD && live) ? a.gA + (size_t)k * H0 * a.B + b : nullptr, Astride);
        for (int j = lane; j < D; j += 32) WV(TD, j) = WV(TA, j);
        __syncwarp();
        for (int j = lane; j < D; j += 32) WV(SU, M[t.permy + j]) = WV(TD, j);
        __syncwarp();
    }
    if (MODE == WARP_FKL_GRAD) return;
    for (int k = K - 1; k >= 0; --k) {
        const TImg& t = g.t[k];
        for (int j = lane; j < D; j += 32) {
            WV(TA, j) = WV(SU, M[t.permy + j]);
            WV(TB, j) = WV(SV + k + 1, M[t.permy + j]);
        }
        __syncwarp();
        w_seq_bwd(g, weights(k), k, wp, TA, TB, SHP + k, w, bslot(0, k), TC, TD, TE,
                  live ? a.gA + (size_t)k * H0 * a.B + b : nullptr, Astride);
        for (int j = lane; j < D; j += 32) WV(SU, M[t.perm + j]) = WV(TC, j);
        __syncwarp();
    }
}

#undef WB
#undef WV
#undef WBITS

}  // namespace rabi
