// Deep / wide conditioners (3+ hidden layers or D > 12: config 5, pyloric: D = 31, hidden [200, 200, 200]) whose weights
// and per-pair state do not fit on chip together.  Same arithmetic as maf_pairs.cuh (one thread owns one (row, sample)
// pair; sampling costs one pass-equivalent; reverse mode for the input gradient only), different placement:
//
//  * per-pair state (activations / cotangents of every layer, chain vectors, ReLU bits) lives in a GLOBAL-memory
//    scratch laid out [quad][pair][4] / [entry][pair], so that a warp touches consecutive addresses: the state is
//    L2-resident (126 MB) instead of limiting occupancy through shared memory (the generic kernel keeps 5 kB per pair
//    in shared memory => one warp per SM);
//  * the weights a CTA needs next are staged in shared memory once per CTA and read as warp-broadcast 128-bit loads:
//    a whole layer for the full (density) passes, the rows / column strips of one MADE degree group per step of the
//    sequential (sampling) passes;
//  * products are register-blocked over 16 (8) units x 4 inputs with packed fma.rn.f32x2 on scratch vectors, with one
//    batch of inputs of look-ahead (gdot_block / gtdot_block, and the stride-templated gdotz / gtdotz of the sequential
//    passes);
//  * every (transform, pass) is its own launch: 4K launches per attack iteration, state carried by the scratch.
//
// Reference arithmetic replaced: as maf_pairs.cuh (ConditionalAutoRegressiveNN._forward, AffineAutoregressive._call /
// _inverse, TransformedDistribution.rsample / log_prob, general_kl_divergence and the autograd input gradient).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "maf_fast.cuh"
#include "maf_image.h"
#include "maf_pairs.cuh"

namespace rabi {

template <>
struct ActBuf<2> {
    float* p;   // &buf[0][pair][0]
    size_t q;   // floats between consecutive quads of one pair (= 4 * padded pair count)
    __device__ __forceinline__ float4 ld4(int i) const { return *reinterpret_cast<const float4*>(p + (size_t)(i >> 2) * q); }
    __device__ __forceinline__ void st4(int i, float4 v) const { *reinterpret_cast<float4*>(p + (size_t)(i >> 2) * q) = v; }
    __device__ __forceinline__ void st1(int i, float v) const { p[(size_t)(i >> 2) * q + (i & 3)] = v; }
    __device__ __forceinline__ float ld1(int i) const { return p[(size_t)(i >> 2) * q + (i & 3)]; }
    __device__ __forceinline__ static void wait_ld(float4&) {}
    __device__ __forceinline__ static void wait_st() {}
};


// Register-blocked products on a scratch-resident vector.  Same arithmetic as dot_block / tdot_block (maf_fast.cuh), but
// the activation quads are fetched in batches of PD with the next batch in flight while the current one is consumed:
// the scratch is L2-resident, and one quad of look-ahead leaves the pass bound by L2 latency.
//   out[c] = sum_{i<e4} W[min(c,cmax)*ld + i] * x[i]
template <int NR>
__device__ __forceinline__ void gdot_block(const float* __restrict__ W, int ld, int cmax, int e4, const ActBuf<2>& x,
                                           float (&out)[NR]) {
    constexpr int PD = 4;
    float2 acc[NR];
    const float* wr[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) {
        acc[c] = make_float2(0.f, 0.f);
        wr[c] = W + (size_t)min(c, cmax) * ld;
    }
    float4 cur[PD], nxt[PD];
    const float* xp = x.p;  // next quad to fetch (pointer stepping instead of 64-bit index arithmetic per load)
#pragma unroll
    for (int q = 0; q < PD; ++q) {
        cur[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        nxt[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * q < e4) cur[q] = *reinterpret_cast<const float4*>(xp);
        xp += x.q;
    }
    for (int i = 0; i < e4; i += 4 * PD) {
#pragma unroll
        for (int q = 0; q < PD; ++q) {
            if (i + 4 * (PD + q) < e4) nxt[q] = *reinterpret_cast<const float4*>(xp);
            xp += x.q;
        }
#pragma unroll
        for (int q = 0; q < PD; ++q)
            if (i + 4 * q < e4) {
                const float4 h = cur[q];
#pragma unroll
                for (int c = 0; c < NR; ++c) {
                    const float4 w = *reinterpret_cast<const float4*>(wr[c] + i + 4 * q);
                    acc[c] = ffma2(make_float2(w.x, w.y), make_float2(h.x, h.y), acc[c]);
                    acc[c] = ffma2(make_float2(w.z, w.w), make_float2(h.z, h.w), acc[c]);
                }
            }
#pragma unroll
        for (int q = 0; q < PD; ++q) cur[q] = nxt[q];
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) out[c] = acc[c].x + acc[c].y;
}
//   out[c] = sum_{v in [va4, vb4)} W[v*ld + c] * x[v]      (NC columns, multiple of 4; W offset to the first column)
template <int NC>
__device__ __forceinline__ void gtdot_block(const float* __restrict__ W, int ld, int va4, int vb4, const ActBuf<2>& x,
                                            float (&out)[NC]) {
    constexpr int PD = 4;
    float2 acc[NC / 2];
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) acc[c] = make_float2(0.f, 0.f);
    float4 cur[PD], nxt[PD];
    const float* xp = x.p + (size_t)(va4 >> 2) * x.q;
#pragma unroll
    for (int q = 0; q < PD; ++q) {
        cur[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        nxt[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (va4 + 4 * q < vb4) cur[q] = *reinterpret_cast<const float4*>(xp);
        xp += x.q;
    }
    for (int v = va4; v < vb4; v += 4 * PD) {
#pragma unroll
        for (int q = 0; q < PD; ++q) {
            if (v + 4 * (PD + q) < vb4) nxt[q] = *reinterpret_cast<const float4*>(xp);
            xp += x.q;
        }
#pragma unroll
        for (int q = 0; q < PD; ++q)
            if (v + 4 * q < vb4) {
                const float4 g4 = cur[q];
                const float* wbase = W + (size_t)(v + 4 * q) * ld;
#pragma unroll
                for (int dv = 0; dv < 4; ++dv) {
                    const float gv = dv == 0 ? g4.x : dv == 1 ? g4.y : dv == 2 ? g4.z : g4.w;
                    const float2 gg = make_float2(gv, gv);
#pragma unroll
                    for (int cc = 0; cc < NC / 4; ++cc) {
                        const float4 w = *reinterpret_cast<const float4*>(wbase + dv * ld + 4 * cc);
                        acc[2 * cc] = ffma2(make_float2(w.x, w.y), gg, acc[2 * cc]);
                        acc[2 * cc + 1] = ffma2(make_float2(w.z, w.w), gg, acc[2 * cc + 1]);
                    }
                }
            }
#pragma unroll
        for (int q = 0; q < PD; ++q) cur[q] = nxt[q];
    }
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) {
        out[2 * c] = acc[c].x;
        out[2 * c + 1] = acc[c].y;
    }
}

// DEEP_FKL: forward-KL attack loss KL(target || q(.|x+delta)): samples from the fixed target (A_p), gradient through the
// density pass only (d/dA_q, emitted by kd_full_bwd).  DEEP_LOGPROB / DEEP_RSAMPLE: one chain only.
// ---- sequential passes: staged weights at COMPILE-TIME strides, zero-padded to whole batches -----------------------
// The staged rows are zero beyond the prefix / group and the scratch vectors carry slack (DEEP_SLACK floats), so the
// loops below run over whole batches of 16 inputs without clamps or predicates, every shared-memory address is
// base + immediate, and the next batch of the scratch vector is always in flight.
constexpr int DEEP_SLD = 256;   // row stride (floats) of staged weight rows   => hidden widths <= 256
constexpr int DEEP_SW = 16;     // row stride of staged column strips (one strip = up to 16 producer units)
constexpr int DEEP_SLACK = 36;  // readable floats past the padded width of every scratch vector

//   out[c] = sum_{i<e16} W[c*STRIDE + i] * x[i]            (e16 multiple of 16; W zero beyond the prefix)
// One batch (16 inputs) of look-ahead, fetched unconditionally: the last batch over-reads 16 floats of slack.  (A deeper
// ring with guarded loads was measured slower: the guards serialise the schedule.)
template <int NR, int STRIDE>
__device__ __forceinline__ void gdotz(const float* __restrict__ W, int e16, const ActBuf<2>& x, float (&out)[NR]) {
    float2 acc[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) acc[c] = make_float2(0.f, 0.f);
    float4 cur[4], nxt[4];
    const float* xp = x.p;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        cur[q] = *reinterpret_cast<const float4*>(xp);
        xp += x.q;
    }
    for (int i = 0; i < e16; i += 16, W += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            nxt[q] = *reinterpret_cast<const float4*>(xp);
            xp += x.q;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 h = cur[q];
#pragma unroll
            for (int c = 0; c < NR; ++c) {
                const float4 w = *reinterpret_cast<const float4*>(W + c * STRIDE + 4 * q);
                acc[c] = ffma2(make_float2(w.x, w.y), make_float2(h.x, h.y), acc[c]);
                acc[c] = ffma2(make_float2(w.z, w.w), make_float2(h.z, h.w), acc[c]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) out[c] = acc[c].x + acc[c].y;
}
//   out[c] = sum_{r<n16} S[r*STRIDE + c] * x[v0 + r]        (n16 multiple of 16, v0 multiple of 4; S zero where masked)
template <int NC, int STRIDE>
__device__ __forceinline__ void gtdotz(const float* __restrict__ S, int v0, int n16, const ActBuf<2>& x, float (&out)[NC]) {
    float2 acc[NC / 2];
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) acc[c] = make_float2(0.f, 0.f);
    float4 cur[4], nxt[4];
    const float* xp = x.p + (size_t)(v0 >> 2) * x.q;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        cur[q] = *reinterpret_cast<const float4*>(xp);
        xp += x.q;
    }
    for (int r = 0; r < n16; r += 16, S += 16 * STRIDE) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            nxt[q] = *reinterpret_cast<const float4*>(xp);
            xp += x.q;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 g4 = cur[q];
#pragma unroll
            for (int dv = 0; dv < 4; ++dv) {
                const float gv = dv == 0 ? g4.x : dv == 1 ? g4.y : dv == 2 ? g4.z : g4.w;
                const float2 gg = make_float2(gv, gv);
#pragma unroll
                for (int cc = 0; cc < NC / 4; ++cc) {
                    const float4 w = *reinterpret_cast<const float4*>(S + (4 * q + dv) * STRIDE + 4 * cc);
                    acc[2 * cc] = ffma2(make_float2(w.x, w.y), gg, acc[2 * cc]);
                    acc[2 * cc + 1] = ffma2(make_float2(w.z, w.w), gg, acc[2 * cc + 1]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    }
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) {
        out[2 * c] = acc[c].x;
        out[2 * c + 1] = acc[c].y;
    }
}
// dst[r][0..e16) (row stride DEEP_SLD) = src[(r0 + r)*ld + 0..e16) for r < n, zero for r in [n, n8) and columns >= ld
__device__ __forceinline__ void deep_stage_rows(float* dst, const float* __restrict__ src, int ld, int r0, int n, int n8,
                                                int e16) {
    const int q = e16 >> 2;
    for (int i = threadIdx.x; i < n8 * q; i += blockDim.x) {
        const int r = i / q, c4 = i - r * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n && 4 * c4 < ld) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * ld) + c4);
        *reinterpret_cast<float4*>(dst + (size_t)r * DEEP_SLD + 4 * c4) = v;
    }
}
// as deep_stage_rows, but the copied rows are zero at columns >= ccut (partner rows of a paired step: old prefix only)
__device__ __forceinline__ void deep_stage_rows_cut(float* dst, const float* __restrict__ src, int ld, int r0, int n, int n8,
                                                    int e16, int ccut) {
    const int q = e16 >> 2;
    for (int i = threadIdx.x; i < n8 * q; i += blockDim.x) {
        const int r = i / q, c4 = i - r * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n && 4 * c4 < ld && 4 * c4 < ccut) {
            v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(r0 + r) * ld) + c4);
            if (4 * c4 + 1 >= ccut) v.y = 0.f;
            if (4 * c4 + 2 >= ccut) v.z = 0.f;
            if (4 * c4 + 3 >= ccut) v.w = 0.f;
        }
        *reinterpret_cast<float4*>(dst + (size_t)r * DEEP_SLD + 4 * c4) = v;
    }
}
// dst[r][c] (8 x 8) = src[(r0 + r)*ld + c0 + c] for r < n, c < nc, else 0   (diagonal block of a paired step)
__device__ __forceinline__ void deep_stage_diag(float* dst, const float* __restrict__ src, int ld, int r0, int n, int c0,
                                                int nc) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        dst[i] = (r < n && c < nc) ? __ldg(src + (size_t)(r0 + r) * ld + c0 + c) : 0.f;
    }
}
// strip[r][0..16) (row stride DEEP_SW) = src[(v0 + r)*ld + c0 .. c0+16) for v0 + r < nrows_src and column < ld, else 0
__device__ __forceinline__ void deep_stage_strip16(float* dst, const float* __restrict__ src, int ld, int v0, int n16,
                                                   int nrows_src, int c0) {
    for (int i = threadIdx.x; i < n16 * 4; i += blockDim.x) {
        const int r = i >> 2, c4 = i & 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v0 + r < nrows_src && c0 + 4 * c4 < ld)
            v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(v0 + r) * ld + c0) + c4);
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}

enum DeepMode { DEEP_FWD = 0, DEEP_GRAD = 1, DEEP_FKL = 2, DEEP_LOGPROB = 3, DEEP_RSAMPLE = 4 };

struct DeepArgs {
    const float* A_p;   // [K,H0,B]
    const float* A_q;
    const float* z;     // [S*B][D]
    float* diff;        // [S*B] log p - log q (optional)
    float* gA;          // [K,H0,B] += d loss / d A_p (GRAD) or d A_q (FKL)  (atomics; caller zeroes)
    float* theta_out;   // RSAMPLE: [S*B][D]
    int S, B;
    long long p0;       // first pair of this batch
    int n;              // pairs in this batch
    int Np;             // padded pair count of the scratch (>= gridDim.x * blockDim.x)
    float w;
    float* act;         // [L][Hq][Np][4]
    float* obuf;        // [Oq][Np][4]   (mbar_0.., sbar_0..) of the reverse passes
    float* vec;         // [NV*D][Np]
    uint32_t* bits;     // [2*K*L*WL][Np]
    float* lp;          // [2][Np]  log p, log q
    int Hq, Oq;
    int record;         // store ReLU masks (GRAD)
};

__host__ __device__ inline int deep_nvec(int K) { return 4 * K + 4; }

struct DeepThread {
    int col;
    bool live;
    long long p;
    int b;
};

__device__ __forceinline__ DeepThread deep_thread(const DeepArgs& a) {
    DeepThread t;
    t.col = blockIdx.x * blockDim.x + threadIdx.x;
    t.live = t.col < a.n;
    t.p = a.p0 + (t.live ? t.col : 0);
    t.b = (int)(t.p % a.B);
    return t;
}

#define DV(slot, j) a.vec[((size_t)(slot) * D + (j)) * a.Np + th.col]

__device__ __forceinline__ ActBuf<2> deep_act(const DeepArgs& a, int l, int col) {
    ActBuf<2> h;
    h.p = a.act + ((size_t)l * a.Hq * a.Np + col) * 4;
    h.q = (size_t)4 * a.Np;
    return h;
}
__device__ __forceinline__ ActBuf<2> deep_obuf(const DeepArgs& a, int col) {
    ActBuf<2> h;
    h.p = a.obuf + (size_t)col * 4;
    h.q = (size_t)4 * a.Np;
    return h;
}

// contiguous global -> shared copy of n4 float4 (all threads of the CTA; callers bracket it with __syncthreads)
__device__ __forceinline__ void deep_stage(float* dst, const float* __restrict__ src, int n4) {
    const float4* s = reinterpret_cast<const float4*>(src);
    float4* d = reinterpret_cast<float4*>(dst);
    const int T = blockDim.x;
    for (int base = threadIdx.x; base < n4; base += 4 * T) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (base + q * T < n4) v[q] = __ldg(s + base + q * T);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (base + q * T < n4) d[base + q * T] = v[q];
    }
}
struct BitAcc {  // ReLU masks of one layer, units arrive in increasing order
    uint32_t w;
    uint32_t* base;  // &bits[slot*WL][col]
    size_t stride;   // Np
    int H;
    bool on;
    __device__ __forceinline__ void put(int unit, bool relu_on) {
        if (!on) return;
        if (relu_on) w |= 1u << (unit & 31);
        if ((unit & 31) == 31 || unit == H - 1) {
            base[(size_t)(unit >> 5) * stride] = w;
            w = 0u;
        }
    }
};
__device__ __forceinline__ uint32_t* deep_bits(const DeepArgs& a, const MafImg& g, int pass, int k, int l, int col) {
    const int WL = (g.Hmax + 31) >> 5;
    return a.bits + ((size_t)((pass * g.K + k) * g.L + l) * WL) * a.Np + col;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sampling pass of transform k under A_p: D steps; step j evaluates the units of MADE degree j of every layer.
template <int DP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) kd_seq_fwd(const MafImg g, const DeepArgs a, const int k) {
    extern __shared__ __align__(16) float sm[];
    const DeepThread th = deep_thread(a);
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int K = g.K, D = g.D, L = g.L, H0 = g.H[0];
    const int SV = 0, SHP = 2 * K + 1;
    const float HALF_LOG_2PI = 0.91893853320467274178f;
    float logp;
    if (k == 0) {
        float ssq = 0.f;
        for (int d = 0; d < D; ++d) {
            const float zv = th.live ? a.z[(size_t)th.p * D + d] : 0.f;
            DV(SV, d) = zv;
            ssq = fmaf(zv, zv, ssq);
        }
        logp = -0.5f * ssq - (float)D * HALF_LOG_2PI;
    } else {
        logp = a.lp[th.col];
    }
    float y[DP];
#pragma unroll
    for (int jj = 0; jj < DP; ++jj) y[jj] = 0.f;
    const float* Ap = a.A_p + (size_t)k * H0 * a.B + th.b;
    ActBuf<2> h[RABI_MAXL];
    BitAcc bacc[RABI_MAXL];
#pragma unroll
    for (int l = 0; l < RABI_MAXL; ++l) {
        if (l >= L) continue;
        h[l] = deep_act(a, l, th.col);
        bacc[l].w = 0u;
        bacc[l].base = deep_bits(a, g, 0, k, l, th.col);
        bacc[l].stride = (size_t)a.Np;
        bacc[l].H = g.H[l];
        bacc[l].on = a.record != 0;
    }
    for (int j = 0; j < D; ++j) {
        // ---- paired steps (j, j+1) when every group of both degrees has <= 8 units (pyloric: 6-7): ONE sweep over the
        // prefix of each layer serves the rows of both degrees (16 rows per block instead of 8, half the prefix
        // traffic, half the barriers); the rows of degree j+1 only lack the contribution of the degree-(j+1) units of
        // the layer below, which step j+1 adds from registers through an 8 x 8 diagonal block.
        bool pair = j + 1 < D;
#pragma unroll
        for (int l = 0; l < RABI_MAXL; ++l)
            if (l < L) pair = pair && M[t.grp[l] + j + 1] - M[t.grp[l] + j] <= 8 && M[t.grp[l] + j + 2] - M[t.grp[l] + j + 1] <= 8;
        if (pair) {
            // shared-memory layout: [W0t rows of both groups][per layer l >= 1: 16 sweep rows | 8 x 8 diagonal]
            //                       [4 output sweep rows (m_j, s_j, m_j+1, s_j+1) | 2 x 8 output diagonal]
            int offs[RABI_MAXL + 1], eA[RABI_MAXL + 1];
            const int n0a = M[t.grp[0] + j + 1] - M[t.grp[0] + j], n0b = M[t.grp[0] + j + 2] - M[t.grp[0] + j + 1];
            {
                int o = (n0a + n0b) * t.ldt;
#pragma unroll
                for (int l = 1; l <= RABI_MAXL; ++l) {
                    offs[l == RABI_MAXL ? 0 : l] = o;  // slot 0 holds the output block's offset
                    eA[l == RABI_MAXL ? 0 : l] = 0;
                    if (l < L) {
                        eA[l] = (M[t.grp[l - 1] + j + 1] + 15) & ~15;
                        o += 16 * DEEP_SLD + 64;
                    }
                }
                eA[0] = (M[t.grp[L - 1] + j + 1] + 15) & ~15;
            }
            __syncthreads();
            deep_stage(sm, F + t.W0t + (size_t)M[t.grp[0] + j] * t.ldt, ((n0a + n0b) * t.ldt) >> 2);
#pragma unroll
            for (int l = 1; l < RABI_MAXL; ++l) {
                if (l >= L) continue;
                const int ra = M[t.grp[l] + j], rb = M[t.grp[l] + j + 1], rc = M[t.grp[l] + j + 2];
                const int ca = M[t.grp[l - 1] + j + 1], cb = M[t.grp[l - 1] + j + 2];
                deep_stage_rows(sm + offs[l], F + t.W[l], t.ld[l], ra, rb - ra, 8, eA[l]);
                deep_stage_rows_cut(sm + offs[l] + 8 * DEEP_SLD, F + t.W[l], t.ld[l], rb, rc - rb, 8, eA[l], ca);
                deep_stage_diag(sm + offs[l] + 16 * DEEP_SLD, F + t.W[l], t.ld[l], rb, rc - rb, ca, cb - ca);
            }
            {
                const int ca = M[t.grp[L - 1] + j + 1], cb = M[t.grp[L - 1] + j + 2];
                float* so = sm + offs[0];
                deep_stage_rows(so, F + t.Wo, t.ldo, j, 1, 1, eA[0]);
                deep_stage_rows(so + DEEP_SLD, F + t.Wo, t.ldo, D + j, 1, 1, eA[0]);
                deep_stage_rows_cut(so + 2 * DEEP_SLD, F + t.Wo, t.ldo, j + 1, 1, 1, eA[0], ca);
                deep_stage_rows_cut(so + 3 * DEEP_SLD, F + t.Wo, t.ldo, D + j + 1, 1, 1, eA[0], ca);
                for (int i = threadIdx.x; i < 16; i += blockDim.x) {
                    const int r = i >> 3, c = i & 7;
                    so[4 * DEEP_SLD + i] = c < cb - ca ? __ldg(F + t.Wo + (size_t)(r == 0 ? j + 1 : D + j + 1) * t.ldo + ca + c) : 0.f;
                }
            }
            __syncthreads();
            float part[RABI_MAXL][8], pout[2] = {0.f, 0.f};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int jj0 = j + half;
                // ---- layer 0 of degree jj0 (the new activations stay in registers for the diagonal blocks)
                float prev[8];
                {
                    const int ua = M[t.grp[0] + jj0], n = M[t.grp[0] + jj0 + 1] - ua;
                    const float* w0 = sm + (size_t)(half ? n0a : 0) * t.ldt;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        prev[c] = 0.f;
                        if (c < n) {
                            float pre = th.live ? __ldg(Ap + (size_t)(ua + c) * a.B) : 0.f;
                            const float* w = w0 + (size_t)c * t.ldt;
#pragma unroll
                            for (int q4 = 0; q4 < DP / 4; ++q4) {
                                const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * q4);
                                pre = fmaf(w4.x, y[4 * q4], pre);
                                pre = fmaf(w4.y, y[4 * q4 + 1], pre);
                                pre = fmaf(w4.z, y[4 * q4 + 2], pre);
                                pre = fmaf(w4.w, y[4 * q4 + 3], pre);
                            }
                            const bool on = pre > 0.f;
                            bacc[0].put(ua + c, on);
                            prev[c] = on ? pre : 0.f;
                            h[0].st1(ua + c, prev[c]);
                        }
                    }
                }
                // ---- hidden layers
#pragma unroll
                for (int l = 1; l < RABI_MAXL; ++l) {
                    if (l >= L) continue;
                    const int va = M[t.grp[l] + jj0], n = M[t.grp[l] + jj0 + 1] - va;
                    float val[8];
                    if (half == 0) {
                        float o[16];
                        gdotz<16, DEEP_SLD>(sm + offs[l], eA[l], h[l - 1], o);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            val[c] = o[c];
                            part[l][c] = o[8 + c];
                        }
                    } else {
                        const float* dg = sm + offs[l] + 16 * DEEP_SLD;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 d0 = *reinterpret_cast<const float4*>(dg + 8 * c);
                            const float4 d1 = *reinterpret_cast<const float4*>(dg + 8 * c + 4);
                            float acc = part[l][c];
                            acc = fmaf(d0.x, prev[0], acc); acc = fmaf(d0.y, prev[1], acc);
                            acc = fmaf(d0.z, prev[2], acc); acc = fmaf(d0.w, prev[3], acc);
                            acc = fmaf(d1.x, prev[4], acc); acc = fmaf(d1.y, prev[5], acc);
                            acc = fmaf(d1.z, prev[6], acc); acc = fmaf(d1.w, prev[7], acc);
                            val[c] = acc;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        prev[c] = 0.f;
                        if (c < n) {
                            const float x = val[c] + __ldg(F + t.b[l] + va + c);
                            const bool on = x > 0.f;
                            bacc[l].put(va + c, on);
                            prev[c] = on ? x : 0.f;
                            h[l].st1(va + c, prev[c]);
                        }
                    }
                }
                // ---- output position jj0
                float m, sraw;
                if (half == 0) {
                    float o[4];
                    gdotz<4, DEEP_SLD>(sm + offs[0], eA[0], deep_act(a, L - 1, th.col), o);
                    m = o[0];
                    sraw = o[1];
                    pout[0] = o[2];
                    pout[1] = o[3];
                } else {
                    const float* dg = sm + offs[0] + 4 * DEEP_SLD;
                    m = pout[0];
                    sraw = pout[1];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        m = fmaf(dg[c], prev[c], m);
                        sraw = fmaf(dg[8 + c], prev[c], sraw);
                    }
                }
                m += __ldg(F + t.bo + jj0);
                sraw += __ldg(F + t.bo + D + jj0);
                const float zj = DV(SV + k, M[t.perm + jj0]);
                const float sh = fminf(fmaxf(sraw, g.lo), g.hi);
                const float yj = (zj - m) * expf(-sh);
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) y[jj] = jj == jj0 ? yj : y[jj];
                logp += sh;
                DV(SHP + k, jj0) = sh;
                DV(SV + k + 1, M[t.permy + jj0]) = yj;
            }
            ++j;  // the partner step is done
            continue;
        }
        // ---- stage the rows of degree j: [W0t rows | W[1] | ... | W[L-1] | Wo rows j, D+j]; hidden / output rows at the
        // fixed stride DEEP_SLD, groups padded with zero rows to a multiple of 8, prefixes to a multiple of 16
        int off[RABI_MAXL + 1], e16[RABI_MAXL + 1];
        {
            int o = 0;
#pragma unroll
            for (int l = 0; l < RABI_MAXL; ++l) {
                off[l] = o;
                e16[l] = 0;
                if (l < L) {
                    const int n = M[t.grp[l] + j + 1] - M[t.grp[l] + j];
                    if (l == 0) {
                        o += n * t.ldt;
                    } else {
                        e16[l] = (M[t.grp[l - 1] + j + 1] + 15) & ~15;
                        o += ((n + 7) & ~7) * DEEP_SLD;
                    }
                }
            }
            off[RABI_MAXL] = o;  // output rows follow the last layer's rows
            e16[RABI_MAXL] = (M[t.grp[L - 1] + j + 1] + 15) & ~15;
        }
        __syncthreads();
        {
            const int r0 = M[t.grp[0] + j], n = M[t.grp[0] + j + 1] - r0;
            deep_stage(sm + off[0], F + t.W0t + (size_t)r0 * t.ldt, (n * t.ldt) >> 2);
        }
#pragma unroll
        for (int l = 1; l < RABI_MAXL; ++l) {
            if (l >= L) continue;
            const int r0 = M[t.grp[l] + j], n = M[t.grp[l] + j + 1] - r0;
            deep_stage_rows(sm + off[l], F + t.W[l], t.ld[l], r0, n, (n + 7) & ~7, e16[l]);
        }
        deep_stage_rows(sm + off[RABI_MAXL], F + t.Wo, t.ldo, j, 1, 1, e16[RABI_MAXL]);
        deep_stage_rows(sm + off[RABI_MAXL] + DEEP_SLD, F + t.Wo, t.ldo, D + j, 1, 1, e16[RABI_MAXL]);
        __syncthreads();
        // ---- layer 0
        {
            const int ua = M[t.grp[0] + j], ub = M[t.grp[0] + j + 1];
            for (int u = ua; u < ub; ++u) {
                float pre = th.live ? __ldg(Ap + (size_t)u * a.B) : 0.f;
                const float* w = sm + off[0] + (size_t)(u - ua) * t.ldt;
#pragma unroll
                for (int q4 = 0; q4 < DP / 4; ++q4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * q4);
                    pre = fmaf(w4.x, y[4 * q4], pre);
                    pre = fmaf(w4.y, y[4 * q4 + 1], pre);
                    pre = fmaf(w4.z, y[4 * q4 + 2], pre);
                    pre = fmaf(w4.w, y[4 * q4 + 3], pre);
                }
                const bool on = pre > 0.f;
                bacc[0].put(u, on);
                h[0].st1(u, on ? pre : 0.f);
            }
        }
        // ---- hidden layers 1..L-1
#pragma unroll
        for (int l = 1; l < RABI_MAXL; ++l) {
            if (l >= L) continue;
            const int va = M[t.grp[l] + j], vb = M[t.grp[l] + j + 1];
            for (int v0 = va; v0 < vb; v0 += 8) {
                const int nv = min(8, vb - v0);
                float o[8];
                gdotz<8, DEEP_SLD>(sm + off[l] + (v0 - va) * DEEP_SLD, e16[l], h[l - 1], o);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < nv) {
                        const float val = o[c] + __ldg(F + t.b[l] + v0 + c);
                        const bool on = val > 0.f;
                        bacc[l].put(v0 + c, on);
                        h[l].st1(v0 + c, on ? val : 0.f);
                    }
            }
        }
        // ---- output position j, then y_j = (z_j - m_j) exp(-clamp(s_j))
        {
            float o[2];
            gdotz<2, DEEP_SLD>(sm + off[RABI_MAXL], e16[RABI_MAXL], deep_act(a, L - 1, th.col), o);
            const float m = o[0] + __ldg(F + t.bo + j), s = o[1] + __ldg(F + t.bo + D + j);
            const float zj = DV(SV + k, M[t.perm + j]);
            const float sh = fminf(fmaxf(s, g.lo), g.hi);
            const float yj = (zj - m) * expf(-sh);
#pragma unroll
            for (int jj = 0; jj < DP; ++jj) y[jj] = jj == j ? yj : y[jj];
            logp += sh;
            DV(SHP + k, j) = sh;
            DV(SV + k + 1, M[t.permy + j]) = yj;
        }
    }
    a.lp[th.col] = logp;
}

// ---------------------------------------------------------------------------------------------------------------------
// Density pass of transform k under A_q (one full MADE pass) + the chain update u_k = exp(clamp(s)) u_{k+1} + m.
template <int DP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) kd_full_fwd(const MafImg g, const DeepArgs a, const int k, const int mode) {
    extern __shared__ __align__(16) float sm[];
    const DeepThread th = deep_thread(a);
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int K = g.K, D = g.D, L = g.L, H0 = g.H[0];
    const int SV = 0, SU = K + 1, SHQ = 3 * K + 1, TB = 4 * K + 2, TC = 4 * K + 3;
    const float HALF_LOG_2PI = 0.91893853320467274178f;
    const int src = (k == K - 1) ? SV + K : SU + k + 1;
    float y[DP];
#pragma unroll
    for (int jj = 0; jj < DP; ++jj) y[jj] = jj < D ? DV(src, M[t.permy + jj]) : 0.f;
    const float* Aq = a.A_q + (size_t)k * H0 * a.B + th.b;
    // ---- layer 0
    {
        const int H0p = (H0 + 3) & ~3;
        __syncthreads();
        deep_stage(sm, F + t.W0t, (H0p * t.ldt) >> 2);
        __syncthreads();
        ActBuf<2> h0 = deep_act(a, 0, th.col);
        BitAcc bacc = {0u, deep_bits(a, g, 1, k, 0, th.col), (size_t)a.Np, H0, a.record != 0};
        for (int u0 = 0; u0 < H0p; u0 += 4) {
            float val[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int u = u0 + c;
                float pre = (th.live && u < H0) ? __ldg(Aq + (size_t)u * a.B) : 0.f;
                const float* w = sm + (size_t)u * t.ldt;
#pragma unroll
                for (int q4 = 0; q4 < DP / 4; ++q4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * q4);
                    pre = fmaf(w4.x, y[4 * q4], pre);
                    pre = fmaf(w4.y, y[4 * q4 + 1], pre);
                    pre = fmaf(w4.z, y[4 * q4 + 2], pre);
                    pre = fmaf(w4.w, y[4 * q4 + 3], pre);
                }
                const bool on = u < H0 && pre > 0.f;
                if (u < H0) bacc.put(u, on);
                val[c] = on ? pre : 0.f;
            }
            h0.st4(u0, make_float4(val[0], val[1], val[2], val[3]));
        }
    }
    // ---- hidden layers
    for (int l = 1; l < L; ++l) {
        const int Hl = g.H[l], Hlp = (Hl + 3) & ~3, ld = t.ld[l];
        __syncthreads();
        deep_stage(sm, F + t.W[l], (Hlp * ld) >> 2);
        __syncthreads();
        const ActBuf<2> in = deep_act(a, l - 1, th.col);
        ActBuf<2> out = deep_act(a, l, th.col);
        BitAcc bacc = {0u, deep_bits(a, g, 1, k, l, th.col), (size_t)a.Np, Hl, a.record != 0};
        for (int v0 = 0; v0 < Hlp; v0 += 16) {
            const int nv = min(16, Hlp - v0);  // padded rows are zero rows of the image
            const int e4 = (M[t.pre[l] + min(v0 + nv, Hl) - 1] + 3) & ~3;
            float o[16];
            gdot_block<16>(sm + (size_t)v0 * ld, ld, nv - 1, e4, in, o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * q < nv) {
                    float val[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int v = v0 + 4 * q + c;
                        const float x = o[4 * q + c] + __ldg(F + t.b[l] + min(v, Hl - 1));
                        const bool on = v < Hl && x > 0.f;
                        if (v < Hl) bacc.put(v, on);
                        val[c] = on ? x : 0.f;
                    }
                    out.st4(v0 + 4 * q, make_float4(val[0], val[1], val[2], val[3]));
                }
        }
    }
    // ---- output layer: rows 0..D-1 means, D..2D-1 log-scales
    {
        const int R = 2 * D, HL = g.H[L - 1], e4 = (HL + 3) & ~3;
        const int Rp = (R + 3) & ~3;
        __syncthreads();
        deep_stage(sm, F + t.Wo, (Rp * t.ldo) >> 2);
        __syncthreads();
        const ActBuf<2> in = deep_act(a, L - 1, th.col);
        for (int r0 = 0; r0 < R; r0 += 16) {
            const int nr = min(16, R - r0);
            float o[16];
            gdot_block<16>(sm + (size_t)r0 * t.ldo, t.ldo, nr - 1, e4, in, o);
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < nr) {
                    const int r = r0 + c;
                    const float val = o[c] + __ldg(F + t.bo + r);
                    if (r < D) DV(TB, r) = val;
                    else DV(TC, r - D) = val;
                }
        }
    }
    float logq = (k == K - 1) ? 0.f : a.lp[(size_t)a.Np + th.col];
#pragma unroll
    for (int jj = 0; jj < DP; ++jj)
        if (jj < D) {
            const float sh = fminf(fmaxf(DV(TC, jj), g.lo), g.hi);
            DV(SHQ + k, jj) = sh;
            logq += sh;
            DV(SU + k, M[t.perm + jj]) = fmaf(expf(sh), y[jj], DV(TB, jj));
        }
    if (k == 0) {
        float ssq = 0.f;
        for (int d = 0; d < D; ++d) {
            const float u = DV(SU, d);
            ssq = fmaf(u, u, ssq);
            if (mode == DEEP_GRAD || mode == DEEP_FKL) DV(SU, d) = a.w * u;  // reverse-pass seed: d(-w log q)/d u_0
        }
        logq += -0.5f * ssq - (float)D * HALF_LOG_2PI;
        if (a.diff && th.live) a.diff[th.p] = mode == DEEP_LOGPROB ? logq : a.lp[th.col] - logq;
    }
    a.lp[(size_t)a.Np + th.col] = logq;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reverse of the density pass of transform k (input gradient only).  The running cotangent lives in vec slot SU.
template <int DP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) kd_full_bwd(const MafImg g, const DeepArgs a, const int k, const int emit_gA) {
    extern __shared__ __align__(16) float sm[];
    const DeepThread th = deep_thread(a);
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int K = g.K, D = g.D, L = g.L, H0 = g.H[0];
    const int WL = (g.Hmax + 31) >> 5;
    const int SV = 0, SU = K + 1, SHQ = 3 * K + 1;
    const int src = (k == K - 1) ? SV + K : SU + k + 1;
    const float w = a.w;
    float yb[DP];
    ActBuf<2> ob = deep_obuf(a, th.col);
    const int R = 2 * D, Rp = (R + 3) & ~3;
    for (int i = R; i < Rp; ++i) ob.st1(i, 0.f);
#pragma unroll
    for (int jj = 0; jj < DP; ++jj) {
        if (jj < D) {
            const float ub = DV(SU, M[t.perm + jj]);
            const float es = expf(DV(SHQ + k, jj));
            const float yin = DV(src, M[t.permy + jj]);
            ob.st1(jj, ub);
            ob.st1(D + jj, fmaf(ub * es, yin, -w));
            yb[jj] = ub * es;
        } else {
            yb[jj] = 0.f;
        }
    }
    // ---- last hidden layer from the output layer
    {
        const int HL = g.H[L - 1], HLp = (HL + 3) & ~3;
        __syncthreads();
        deep_stage(sm, F + t.Wo, (Rp * t.ldo) >> 2);
        __syncthreads();
        ActBuf<2> out = deep_act(a, L - 1, th.col);
        const uint32_t* bw = deep_bits(a, g, 1, k, L - 1, th.col);
        for (int v0 = 0; v0 < HLp; v0 += 16) {
            float o[16];
            gtdot_block<16>(sm + v0, t.ldo, 0, Rp, ob, o);  // columns >= ldo are never stored
            const uint32_t word = bw[(size_t)(v0 >> 5) * a.Np] >> (v0 & 31);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (v0 + 4 * q < HLp) {
                    float val[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) val[c] = ((word >> (4 * q + c)) & 1u) ? o[4 * q + c] : 0.f;
                    out.st4(v0 + 4 * q, make_float4(val[0], val[1], val[2], val[3]));
                }
        }
    }
    // ---- down the hidden layers
    for (int l = L - 1; l >= 1; --l) {
        const int Hin = g.H[l - 1], Hinp = (Hin + 3) & ~3, Hl = g.H[l], Hlp = (Hl + 3) & ~3, ld = t.ld[l];
        __syncthreads();
        deep_stage(sm, F + t.W[l], (Hlp * ld) >> 2);
        __syncthreads();
        const ActBuf<2> in = deep_act(a, l, th.col);
        ActBuf<2> out = deep_act(a, l - 1, th.col);
        const uint32_t* bw = deep_bits(a, g, 1, k, l - 1, th.col);
        for (int u0 = 0; u0 < Hinp; u0 += 16) {
            const int va4 = min(M[t.first[l] + u0], Hlp - 4) & ~3;
            float o[16];
            gtdot_block<16>(sm + u0, ld, va4, Hlp, in, o);
            const uint32_t word = bw[(size_t)(u0 >> 5) * a.Np] >> (u0 & 31);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (u0 + 4 * q < Hinp) {
                    float val[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) val[c] = ((word >> (4 * q + c)) & 1u) ? o[4 * q + c] : 0.f;
                    out.st4(u0 + 4 * q, make_float4(val[0], val[1], val[2], val[3]));
                }
        }
    }
    // ---- theta positions from layer 0
    {
        const int H0p = (H0 + 3) & ~3;
        __syncthreads();
        deep_stage(sm, F + t.W0t, (H0p * t.ldt) >> 2);
        __syncthreads();
        if (emit_gA && th.live) {  // forward-KL attack: the differentiated pass is this one
            const ActBuf<2> g0 = deep_act(a, 0, th.col);
            float* gAk = a.gA + (size_t)k * H0 * a.B + th.b;
            for (int u0 = 0; u0 < H0; u0 += 4) {
                const float4 v = g0.ld4(u0);
                atomicAdd(gAk + (size_t)u0 * a.B, v.x);
                if (u0 + 1 < H0) atomicAdd(gAk + (size_t)(u0 + 1) * a.B, v.y);
                if (u0 + 2 < H0) atomicAdd(gAk + (size_t)(u0 + 2) * a.B, v.z);
                if (u0 + 3 < H0) atomicAdd(gAk + (size_t)(u0 + 3) * a.B, v.w);
            }
        }
        float o[DP];
        gtdot_block<DP>(sm, t.ldt, 0, H0p, deep_act(a, 0, th.col), o);
#pragma unroll
        for (int jj = 0; jj < DP; ++jj) yb[jj] += o[jj];
    }
#pragma unroll
    for (int jj = 0; jj < DP; ++jj)
        if (jj < D) DV(SU, M[t.permy + jj]) = yb[jj];
    (void)WL;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reverse of the sampling pass of transform k: D steps in descending order, pull style (every cotangent is produced
// once, when all its consumers are final); emits d loss / d A_p.
template <int DP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) kd_seq_bwd(const MafImg g, const DeepArgs a, const int k) {
    extern __shared__ __align__(16) float sm[];
    const DeepThread th = deep_thread(a);
    const TImg& t = g.t[k];
    const float* F = g.f;
    const int* M = g.m;
    const int K = g.K, D = g.D, L = g.L, H0 = g.H[0];
    const int SV = 0, SU = K + 1, SHP = 2 * K + 1, TC = 4 * K + 3;
    const float w = a.w;
    float yb[DP];
#pragma unroll
    for (int jj = 0; jj < DP; ++jj) yb[jj] = jj < D ? DV(SU, M[t.permy + jj]) : 0.f;
    ActBuf<2> ob = deep_obuf(a, th.col);
    const int R = 2 * D, Rp = (R + 3) & ~3;
    for (int i = 0; i < Rp; ++i) ob.st1(i, 0.f);
    ActBuf<2> h[RABI_MAXL];
#pragma unroll
    for (int l = 0; l < RABI_MAXL; ++l)
        if (l < L) h[l] = deep_act(a, l, th.col);
    float* gAk = a.gA + (size_t)k * H0 * a.B + th.b;
    for (int j = D - 1; j >= 0; --j) {
        {
            float ybj = 0.f;
#pragma unroll
            for (int jj = 0; jj < DP; ++jj) ybj = jj == j ? yb[jj] : ybj;
            const float shj = DV(SHP + k, j);
            const float yj = DV(SV + k + 1, M[t.permy + j]);
            const float zbj = ybj * expf(-shj);
            ob.st1(j, -zbj);
            ob.st1(D + j, fmaf(-ybj, yj, w));
            DV(TC, j) = zbj;
        }
        // ---- stage: for the degree-j units of every layer (producers) the column strips of their consumer matrix
        // (Wo for the last hidden layer, W[l+1] otherwise), 16 columns per strip at stride DEEP_SW, consumer rows from
        // r0 (first row that can see the group, rounded down to 4) in whole batches of 16, zero beyond the matrix;
        // plus the layer-0 rows of the group for the push into the theta cotangents
        int off[RABI_MAXL + 1], c0[RABI_MAXL], nblk[RABI_MAXL], r0[RABI_MAXL], n16[RABI_MAXL];
        {
            int o = 0;
#pragma unroll
            for (int l = RABI_MAXL - 1; l >= 0; --l) {
                off[l + 1] = o;
                c0[l] = nblk[l] = r0[l] = n16[l] = 0;
                if (l >= L) continue;
                const int ua = M[t.grp[l] + j], ub = M[t.grp[l] + j + 1];
                c0[l] = ua & ~3;
                nblk[l] = ub > ua ? (ub - c0[l] + 15) >> 4 : 0;
                r0[l] = l == L - 1 ? 0 : (M[t.grp[l + 1] + j] & ~3);
                const int rows = l == L - 1 ? Rp : ((g.H[l + 1] + 3) & ~3);
                n16[l] = (max(rows - r0[l], 0) + 15) & ~15;
                o += nblk[l] * n16[l] * DEEP_SW;
            }
            off[0] = o;  // layer-0 rows of the group: [n0][ldt]
        }
        __syncthreads();
#pragma unroll
        for (int l = RABI_MAXL - 1; l >= 0; --l) {
            if (l >= L) continue;
            const float* src = l == L - 1 ? F + t.Wo : F + t.W[l + 1];
            const int ld = l == L - 1 ? t.ldo : t.ld[l + 1];
            const int rows = l == L - 1 ? Rp : ((g.H[l + 1] + 3) & ~3);
            for (int bk = 0; bk < nblk[l]; ++bk)
                deep_stage_strip16(sm + off[l + 1] + bk * n16[l] * DEEP_SW, src, ld, r0[l], n16[l], rows, c0[l] + 16 * bk);
        }
        {
            const int ua = M[t.grp[0] + j], n = M[t.grp[0] + j + 1] - ua;
            deep_stage(sm + off[0], F + t.W0t + (size_t)ua * t.ldt, (n * t.ldt) >> 2);
        }
        __syncthreads();
#pragma unroll
        for (int l = RABI_MAXL - 1; l >= 0; --l) {
            if (l >= L) continue;
            const int ua = M[t.grp[l] + j], ub = M[t.grp[l] + j + 1];
            if (ub <= ua) continue;
            const uint32_t* bw = deep_bits(a, g, 0, k, l, th.col);
            const ActBuf<2> in = l == L - 1 ? ob : h[l + 1 < RABI_MAXL ? l + 1 : l];
            auto emit = [&](int cbase, int c, float val) {
                const int u = cbase + c;
                if (u >= ua && u < ub) {
                    const bool on = (bw[(size_t)(u >> 5) * a.Np] >> (u & 31)) & 1u;
                    const float gv = on ? val : 0.f;
                    if (l > 0) {
                        h[l].st1(u, gv);
                    } else {
                        if (th.live) atomicAdd(gAk + (size_t)u * a.B, gv);
                        const float* wr = sm + off[0] + (size_t)(u - ua) * t.ldt;
#pragma unroll
                        for (int q4 = 0; q4 < DP / 4; ++q4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * q4);
                            yb[4 * q4] = fmaf(w4.x, gv, yb[4 * q4]);
                            yb[4 * q4 + 1] = fmaf(w4.y, gv, yb[4 * q4 + 1]);
                            yb[4 * q4 + 2] = fmaf(w4.z, gv, yb[4 * q4 + 2]);
                            yb[4 * q4 + 3] = fmaf(w4.w, gv, yb[4 * q4 + 3]);
                        }
                    }
                }
            };
            for (int bk = 0; bk < nblk[l]; ++bk) {
                const float* S = sm + off[l + 1] + bk * n16[l] * DEEP_SW;
                const int cbase = c0[l] + 16 * bk;
                const int width = ((ub + 3) & ~3) - cbase;  // columns of this strip that hold units of the group
                if (width > 12) {
                    float o[16];
                    gtdotz<16, DEEP_SW>(S, r0[l], n16[l], in, o);
#pragma unroll
                    for (int c = 0; c < 16; ++c) emit(cbase, c, o[c]);
                } else if (width > 8) {
                    float o[12];
                    gtdotz<12, DEEP_SW>(S, r0[l], n16[l], in, o);
#pragma unroll
                    for (int c = 0; c < 12; ++c) emit(cbase, c, o[c]);
                } else if (width > 4) {
                    float o[8];
                    gtdotz<8, DEEP_SW>(S, r0[l], n16[l], in, o);
#pragma unroll
                    for (int c = 0; c < 8; ++c) emit(cbase, c, o[c]);
                } else {
                    float o[4];
                    gtdotz<4, DEEP_SW>(S, r0[l], n16[l], in, o);
#pragma unroll
                    for (int c = 0; c < 4; ++c) emit(cbase, c, o[c]);
                }
            }
        }
    }
#pragma unroll
    for (int jj = 0; jj < DP; ++jj)
        if (jj < D) DV(SU, M[t.perm + jj]) = DV(TC, jj);
}

// LOGPROB: theta -> chain slot v_K;   RSAMPLE: v_K -> theta_out, log q(own sample) -> diff
__global__ void kd_load_theta(const MafImg g, const DeepArgs a) {
    const DeepThread th = deep_thread(a);
    const int D = g.D, K = g.K;
    for (int d = 0; d < D; ++d) DV(K, d) = th.live ? a.z[(size_t)th.p * D + d] : 0.f;
}
__global__ void kd_store_theta(const MafImg g, const DeepArgs a) {
    const DeepThread th = deep_thread(a);
    const int D = g.D, K = g.K;
    if (!th.live) return;
    for (int d = 0; d < D; ++d) a.theta_out[(size_t)th.p * D + d] = DV(K, d);
    if (a.diff) a.diff[th.p] = a.lp[th.col];
}

#undef DV

}  // namespace rabi
