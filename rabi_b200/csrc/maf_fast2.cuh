// k_fast2: the reverse-KL attack gradient of k_fast (maf_fast.cuh, FAST_GRAD) with TWO LANES PER PAIR.
//
// Why: one PGD iteration of the lotka_volterra workload has 10 000 rows x 5 samples = 338 pairs per SM.  With a thread
// per pair that is 11 warps per SM (2.75 per scheduler) and every warp runs the whole 164 k-instruction program of a
// pair: the kernel is latency-bound (issue slots 47 % busy, ncu).  Here the two lanes 2i, 2i+1 of a warp share pair i:
//  * the INPUT side of every product is split ("k-split"): lane L owns the aligned quads q = 2m + L of the layer-0
//    activations (forward) and of the layer-1 cotangents (reverse) in its private TMEM row (52 instead of 100 columns),
//    computes the partial dot products over its quads and the two partial sums meet in one __shfl_xor per output unit;
//  * the OUTPUT side is split by theta position: lane L accumulates (mean, log-scale) of positions 2L, 2L+1 in the
//    forward and the cotangents of positions 2L, 2L+1 in the reverse direction, so no further reduction is needed;
//  * everything of size D (chain vectors, log-densities) is replicated, the per-pair shared-memory slots are shared.
// 22 warps per SM with half the instructions each.  Same arithmetic per pair as k_fast up to the association of the
// two partial sums.  Shapes: D <= 4, two hidden layers of at most 104 units (config/model/maf_pyro.yaml at the
// lotka_volterra task); everything else stays on k_fast.
#pragma once
#include "maf_fast.cuh"

namespace rabi {
#ifndef RABI_F2_DOT_UNROLL
#define RABI_F2_DOT_UNROLL 2
#endif
#ifndef RABI_F2_TDOT_UNROLL
#define RABI_F2_TDOT_UNROLL 2
#endif
constexpr int kF2DotUnroll = RABI_F2_DOT_UNROLL, kF2TdotUnroll = RABI_F2_TDOT_UNROLL;

// partial[c] = sum over the own quads m < M of W[c*ld + 8m .. 8m+3] . act[4m .. 4m+3]     (W already offset by 4L)
template <int NR, int LDC>
__device__ __forceinline__ void dot2_block(const float* __restrict__ W, int ld_rt, int M, const ActBuf<1>& ab,
                                           float (&out)[NR]) {
    float2 acc[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) acc[c] = make_float2(0.f, 0.f);
    const int ld = LDC ? LDC : ld_rt;
    float4 hn = ab.ld4(0);
#pragma unroll(kF2DotUnroll)
    for (int m = 0; m < M; ++m) {
        ActBuf<1>::wait_ld(hn);
        const float4 h = hn;
        hn = ab.ld4(4 * (m + 1));  // unconditional look-ahead: the quad after the last one is inside the CTA's TMEM allocation, never used
#pragma unroll
        for (int c = 0; c < NR; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(W + (size_t)c * ld + 8 * m);
            acc[c] = ffma2(make_float2(w.x, w.y), make_float2(h.x, h.y), acc[c]);
            acc[c] = ffma2(make_float2(w.z, w.w), make_float2(h.z, h.w), acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) out[c] = acc[c].x + acc[c].y;
}

// partial[c] = sum over the own quads m0 <= m < m1, dv < 4 of W[(8m + dv)*ld + c] * g[4m + dv]   (W offset by 4L rows)
template <int NC, int LDC>
__device__ __forceinline__ void tdot2_block(const float* __restrict__ W, int ld_rt, int m0, int m1, const ActBuf<1>& ab,
                                            float (&out)[NC]) {
    float2 acc[NC / 2];
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) acc[c] = make_float2(0.f, 0.f);
    const int ld = LDC ? LDC : ld_rt;
    const float* wbase = W + (size_t)(8 * m0) * ld;
    float4 gn = ab.ld4(4 * m0);
#pragma unroll(kF2TdotUnroll)
    for (int m = m0; m < m1; ++m, wbase += 8 * ld) {
        ActBuf<1>::wait_ld(gn);
        const float4 g4 = gn;
        gn = ab.ld4(4 * (m + 1));
#pragma unroll
        for (int dv = 0; dv < 4; ++dv) {
            const float* wrow = wbase + dv * ld;
            const float gv = dv == 0 ? g4.x : dv == 1 ? g4.y : dv == 2 ? g4.z : g4.w;
            const float2 gg = make_float2(gv, gv);
#pragma unroll
            for (int cc = 0; cc < NC / 4; ++cc) {
                const float4 w = *reinterpret_cast<const float4*>(wrow + 4 * cc);
                acc[2 * cc] = ffma2(make_float2(w.x, w.y), gg, acc[2 * cc]);
                acc[2 * cc + 1] = ffma2(make_float2(w.z, w.w), gg, acc[2 * cc + 1]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) {
        out[2 * c] = acc[c].x;
        out[2 * c + 1] = acc[c].y;
    }
}

template <int LDC>
__global__ void __launch_bounds__(704, 1) k_fast2(const MafImg g, const FastArgs a, const FastRel rel) {
    extern __shared__ __align__(16) float sm[];
    constexpr int DP = 4;
    constexpr unsigned FULL = 0xffffffffu;
    const int T = blockDim.x, tid = threadIdx.x;
    const int NPS = T >> 1;              // pair slots of the tile
    const int slot = tid >> 1, L = tid & 1;
    const int lane = tid & 31;
    const int K = g.K, D = g.D, H0 = g.H[0], H1 = g.H[1];
    const int WL = (g.Hmax + 31) >> 5;
    float* Ws = sm;
    int* Ms = reinterpret_cast<int*>(Ws + a.wfloats);
    float* vec = reinterpret_cast<float*>(Ms + a.wints);                        // [NPS][VS]
    uint32_t* bits = reinterpret_cast<uint32_t*>(vec + (size_t)NPS * a.VS);     // [NPS][BWS]

    const float* W0t = Ws + rel.W0t;
    const float* b1 = Ws + rel.b1;
    const float* W1 = Ws + rel.W1;
    const float* WoT = Ws + rel.WoT;   // [H1][8]: (m_0, s_0, m_1, s_1, m_2, s_2, m_3, s_3)
    const float* bo = Ws + rel.bo;
    const int ld1 = LDC ? LDC : rel.ld1;
    const int* perm = Ms + rel.perm;
    const int* permy = Ms + rel.permy;
    const int* grp0 = Ms + rel.grp0;
    const int* grp1 = Ms + rel.grp1;

    __shared__ uint32_t tmem_base_s;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                     "r"((uint32_t)a.tm_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int CS = a.AS;        // own TMEM columns of a lane (4 per own quad)
    const int CQ = CS >> 2;     // own quads
    ActBuf<1> ab;
    {
        const int warp = tid >> 5;
        ab.t = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * CS);
    }
    float* vecp = vec + (size_t)slot * a.VS;
    uint32_t* bitp = bits + (size_t)slot * a.BWS;
    for (int i = 0; i < CS; i += 4) ab.st4(i, make_float4(0.f, 0.f, 0.f, 0.f));
    ActBuf<1>::wait_st();

    const int oV = 0;
    const int oU = (K + 1) * D;
    const int oSHP = (2 * K + 1) * D;
    const int oSHQ = (3 * K + 1) * D;
    const int oG = (4 * K + 1) * D;
    const float HALF_LOG_2PI = 0.91893853320467274178f;
    const float w = a.w;

    int resident = -1;
    auto ensure_weights = [&](int k) {
        if (resident == k) return;
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(g.f + g.t[k].W0t);
            float4* dst = reinterpret_cast<float4*>(Ws);
            const int n4 = a.wfloats >> 2;
            for (int base = tid; base < n4; base += 4 * T) {
                float4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (base + q * T < n4) v[q] = __ldg(src + base + q * T);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (base + q * T < n4) dst[base + q * T] = v[q];
            }
            const int* msrc = g.m + g.t[k].perm;
            for (int i = tid; i < a.wints; i += T) Ms[i] = __ldg(msrc + i);
        }
        __syncthreads();
        resident = k;
    };

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int b0 = tile * a.TB;
        const int rows_here = min(a.TB, a.B - b0);
        const bool live = slot < rows_here * a.S;
        const int s = live ? slot / rows_here : 0;
        const int rrow = live ? slot % rows_here : 0;
        const long long pidx = (long long)s * a.B + b0 + rrow;
        const float* Ap = a.A_p + (b0 + rrow);
        const float* Aq = a.A_q + (b0 + rrow);
        float logp, logq = 0.f;
        {
            float ssq = 0.f;
            for (int d = 0; d < D; ++d) {
                const float zv = live ? a.z[(size_t)pidx * D + d] : 0.f;
                vecp[oV + d] = zv;  // both lanes of the pair store the same value
                ssq = fmaf(zv, zv, ssq);
            }
            logp = -0.5f * ssq - (float)D * HALF_LOG_2PI;
        }

        const int nphase = 4 * K;
        for (int ph = 0; ph < nphase; ++ph) {
            const int kind = ph / K;
            const int idx = ph - kind * K;
            const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
            ensure_weights(k);
            __syncwarp();  // chain vectors written by the partner lane in the previous phase
            const bool seq = (kind == 0 || kind == 3);
            const int bs = (seq ? k : K + k) * 2;  // relu-bit slot of (pass, k): WL words of layer 0, WL words of layer 1
            uint32_t* bw0 = bitp + bs * WL;
            uint32_t* bw1 = bitp + (bs + 1) * WL;
            const int src = (k == K - 1) ? oV + K * D : oU + (k + 1) * D;

            if (kind < 2) {
                // ===================== forward pass of transform k =====================
                const float* Aptr = (seq ? Ap : Aq) + (size_t)k * H0 * a.B;
                float zos[DP], y[DP];
                float2 oms[2];  // (mean, log-scale) accumulators of the own positions 2L, 2L + 1
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) {
                    zos[jj] = (seq && jj < D) ? vecp[oV + k * D + perm[jj]] : 0.f;
                    y[jj] = (!seq && jj < D) ? vecp[src + permy[jj]] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int jj = 2 * L + i;
                    oms[i] = jj < D ? make_float2(bo[2 * jj], bo[2 * jj + 1]) : make_float2(0.f, 0.f);
                }
                for (int i = 0; i < 2 * WL; ++i) bw0[i] = 0u;  // both layers' words of this (pass, k)
                __syncwarp();
                // ---- context projection of the own quads -> TMEM (16 loads in flight; idle slots read row 0 of the tile)
                {
                    const float* Al = Aptr + (size_t)(4 * L) * a.B;
                    const int nreal = (((H0 + 3) >> 2) + 1) >> 1;  // own quads of lane 0 (warp-uniform loop: lane 1 pads with zeros)
                    for (int m0 = 0; m0 < nreal; m0 += 4) {
                        float av[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int u = 8 * (m0 + (c >> 2)) + (c & 3);  // + 4L through Al
                            av[c] = (u + 4 * L < H0) ? __ldg(Al + (size_t)u * a.B) : 0.f;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (m0 + q < nreal) ab.st4(4 * (m0 + q), make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]));
                    }
                }
                ActBuf<1>::wait_st();
                for (int j = 0; j < D; ++j) {
                    // ---- layer-0 units of degree j in the own quads: act = relu(A[u] + W0t[u] . y)
                    {
                        const int ua = grp0[j], ub = grp0[j + 1];
                        uint32_t wacc = 0u;  // relu bits of the own units in the current 32-unit word
                        const int mend = (ub + 7) >> 3;
                        for (int m = ua >> 3; m < mend; ++m) {
                            const int u0 = 8 * m + 4 * L;
                            float4 av = ab.ld4(4 * m);
                            ActBuf<1>::wait_ld(av);
                            float pre[4] = {av.x, av.y, av.z, av.w};
                            uint32_t nib = 0u;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4 wt = *reinterpret_cast<const float4*>(W0t + (size_t)(u0 + c) * DP);
                                float p = pre[c];
                                p = fmaf(wt.x, y[0], p);
                                p = fmaf(wt.y, y[1], p);
                                p = fmaf(wt.z, y[2], p);
                                p = fmaf(wt.w, y[3], p);
                                const int u = u0 + c;
                                if (u >= ua && u < ub) {
                                    const bool on = p > 0.f;
                                    pre[c] = on ? p : 0.f;
                                    nib |= on ? (1u << c) : 0u;
                                }
                            }
                            ab.st4(4 * m, make_float4(pre[0], pre[1], pre[2], pre[3]));
                            wacc |= nib << (8 * (m & 3) + 4 * L);
                            if ((m & 3) == 3 || m == mend - 1) {  // warp-uniform: the two lanes meet, both store the word
                                wacc |= __shfl_xor_sync(FULL, wacc, 1);
                                bw0[m >> 2] |= wacc;
                                wacc = 0u;
                            }
                        }
                        ActBuf<1>::wait_st();
                    }
                    // ---- layer-1 units of degree j: partial products over the own quads, one shuffle per unit,
                    //      relu bits, (mean, log-scale) of the own positions
                    {
                        const int va = grp1[j], vb = grp1[j + 1];
                        const int M = (((grp0[j + 1] + 3) >> 2) + 1) >> 1;  // own quads covering the prefix
                        auto block = [&](auto NRtag, int v0) {
                            constexpr int NR = decltype(NRtag)::value;
                            float h[NR];
                            dot2_block<NR, LDC>(W1 + (size_t)v0 * ld1 + 4 * L, ld1, M, ab, h);
                            uint32_t bm = 0u;
#pragma unroll
                            for (int c = 0; c < NR; ++c) {
                                const int v = v0 + c;
                                const float val = h[c] + __shfl_xor_sync(FULL, h[c], 1) + b1[v];
                                const bool on = val > 0.f;
                                const float hv = on ? val : 0.f;
                                const float2 hh = make_float2(hv, hv);
                                const float4 wq = *reinterpret_cast<const float4*>(WoT + (size_t)v * 2 * DP + 4 * L);
                                oms[0] = ffma2(make_float2(wq.x, wq.y), hh, oms[0]);
                                oms[1] = ffma2(make_float2(wq.z, wq.w), hh, oms[1]);
                                bm |= on ? (1u << c) : 0u;
                            }
                            const int sh = v0 & 31;
                            bw1[v0 >> 5] |= bm << sh;
                            if (sh + NR > 32) bw1[(v0 >> 5) + 1] |= bm >> (32 - sh);
                        };
                        int v = va;
                        while (v < vb) {
                            const int n = vb - v;
                            switch (n >= 13 ? 8 : n) {
                                case 12: block(std::integral_constant<int, 12>{}, v); v += 12; break;
                                case 11: block(std::integral_constant<int, 11>{}, v); v += 11; break;
                                case 10: block(std::integral_constant<int, 10>{}, v); v += 10; break;
                                case 9: block(std::integral_constant<int, 9>{}, v); v += 9; break;
                                case 8: block(std::integral_constant<int, 8>{}, v); v += 8; break;
                                case 7: case 6: case 5: case 4: block(std::integral_constant<int, 4>{}, v); v += 4; break;
                                case 3: case 2: block(std::integral_constant<int, 2>{}, v); v += 2; break;
                                default: block(std::integral_constant<int, 1>{}, v); v += 1; break;
                            }
                        }
                    }
                    if (seq) {  // position j is final: y_j = (z_j - m_j) exp(-clamp(s_j)); the owner lane is j >> 1
                        const float mo = (j & 1) ? oms[1].x : oms[0].x;
                        const float so = (j & 1) ? oms[1].y : oms[0].y;
                        const int owner = (lane & ~1) | (j >> 1);
                        const float mj = __shfl_sync(FULL, mo, owner);
                        const float sj = __shfl_sync(FULL, so, owner);
                        float zj = 0.f;
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) zj = jj == j ? zos[jj] : zj;
                        const float sh = fminf(fmaxf(sj, g.lo), g.hi);
                        const float yj = (zj - mj) * expf(-sh);
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) y[jj] = jj == j ? yj : y[jj];
                        logp += sh;
                        vecp[oSHP + k * D + j] = sh;
                        vecp[oV + (k + 1) * D + permy[j]] = yj;
                    }
                }
                if (!seq) {  // u_k = exp(clamp(s)) * u_{k+1} + m, the own positions
                    float lsum = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int jj = 2 * L + i;
                        const float yj = L ? y[2 + i] : y[i];
                        if (jj < D) {
                            const float sh = fminf(fmaxf(oms[i].y, g.lo), g.hi);
                            vecp[oSHQ + k * D + jj] = sh;
                            lsum += sh;
                            vecp[oU + k * D + perm[jj]] = fmaf(expf(sh), yj, oms[i].x);
                        }
                    }
                    logq += lsum + __shfl_xor_sync(FULL, lsum, 1);
                    if (k == 0) {  // end of the density chain: base log-density, KL term, reverse-mode seed
                        __syncwarp();
                        float ssq = 0.f;
                        for (int d = 0; d < D; ++d) {
                            const float u = vecp[oU + d];
                            ssq = fmaf(u, u, ssq);
                            vecp[oG + d] = w * u;  // d(-w log q)/d u_0
                        }
                        logq += -0.5f * ssq - (float)D * HALF_LOG_2PI;
                        if (a.diff && live && L == 0) a.diff[pidx] = logp - logq;
                    }
                }
            } else {
                // ===================== reverse pass of transform k =====================
                float yv[DP], sh[DP], mb[DP], sb[DP], zb[DP];
                float ybo[2];  // cotangents of the own positions 2L, 2L + 1
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) {
                    zb[jj] = 0.f;
                    if (seq) {
                        yv[jj] = jj < D ? vecp[oV + (k + 1) * D + permy[jj]] : 0.f;
                        sh[jj] = jj < D ? vecp[oSHP + k * D + jj] : 0.f;
                        mb[jj] = sb[jj] = 0.f;
                    } else if (jj < D) {
                        const float ub = vecp[oG + perm[jj]];
                        const float es = expf(vecp[oSHQ + k * D + jj]);
                        const float yin = vecp[src + permy[jj]];
                        mb[jj] = ub;
                        sb[jj] = fmaf(ub * es, yin, -w);
                        yv[jj] = ub * es;  // full d/dy seed; the own half goes to ybo below
                        sh[jj] = 0.f;
                    } else {
                        mb[jj] = sb[jj] = yv[jj] = sh[jj] = 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int jj = 2 * L + i;
                    if (seq) ybo[i] = jj < D ? vecp[oG + permy[jj]] : 0.f;
                    else ybo[i] = L ? yv[2 + i] : yv[i];
                }
                __syncwarp();  // all reads of oG are done before either lane overwrites it at the end of the phase
                for (int j = D - 1; j >= 0; --j) {
                    if (seq) {
                        const float yo = (j & 1) ? ybo[1] : ybo[0];
                        const float ybj = __shfl_sync(FULL, yo, (lane & ~1) | (j >> 1));
                        float shj = 0.f, yj = 0.f;
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) {
                            shj = jj == j ? sh[jj] : shj;
                            yj = jj == j ? yv[jj] : yj;
                        }
                        const float zbj = ybj * expf(-shj);
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) {
                            zb[jj] = jj == j ? zbj : zb[jj];
                            mb[jj] = jj == j ? -zbj : mb[jj];
                            sb[jj] = jj == j ? fmaf(-ybj, yj, w) : sb[jj];
                        }
                    }
                    // ---- layer-1 cotangents of the degree-j units in the own quads: g = relu'(.) * (WoT[v] . (mbar, sbar))
                    {
                        const int va = grp1[j], vb = grp1[j + 1];
                        for (int m = va >> 3; m < ((vb + 7) >> 3); ++m) {
                            const int v0 = 8 * m + 4 * L;
                            float4 old = ab.ld4(4 * m);
                            const uint32_t word = __funnelshift_r(bw1[v0 >> 5], bw1[(v0 >> 5) + 1], v0 & 31);
                            float gs[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)(v0 + c) * 2 * DP);
                                const float4 wa = wo[0], wc = wo[1];
                                float2 g2 = make_float2(0.f, 0.f);
                                g2 = ffma2(make_float2(wa.x, wa.y), make_float2(mb[0], sb[0]), g2);
                                g2 = ffma2(make_float2(wa.z, wa.w), make_float2(mb[1], sb[1]), g2);
                                g2 = ffma2(make_float2(wc.x, wc.y), make_float2(mb[2], sb[2]), g2);
                                g2 = ffma2(make_float2(wc.z, wc.w), make_float2(mb[3], sb[3]), g2);
                                gs[c] = ((word >> c) & 1u) ? g2.x + g2.y : 0.f;
                            }
                            ActBuf<1>::wait_ld(old);
                            float nv[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int v = v0 + c;
                                nv[c] = (v >= va && v < vb) ? gs[c] : nv[c];
                            }
                            ab.st4(4 * m, make_float4(nv[0], nv[1], nv[2], nv[3]));
                        }
                        ActBuf<1>::wait_st();
                    }
                    // ---- layer-0 cotangents of the degree-j units (consumed by layer-1 rows >= grp1[j])
                    {
                        const int ua = grp0[j], ub = grp0[j + 1];
                        const int m0 = grp1[j] >> 3, m1 = (((H1 + 3) >> 2) + 1) >> 1;
                        auto cblock = [&](auto NCtag, int u0) {
                            constexpr int NC = decltype(NCtag)::value;
                            float g0[NC];
                            tdot2_block<NC, LDC>(W1 + u0 + (size_t)(4 * L) * ld1, ld1, m0, m1, ab, g0);
                            const uint32_t wbits = __funnelshift_r(bw0[u0 >> 5], bw0[(u0 >> 5) + 1], u0 & 31);
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                const int u = u0 + c;
                                const float full = g0[c] + __shfl_xor_sync(FULL, g0[c], 1);
                                if (u >= ua && u < ub) {
                                    const float gv = ((wbits >> c) & 1u) ? full : 0.f;
                                    const float2 wt = *reinterpret_cast<const float2*>(W0t + (size_t)u * DP + 2 * L);
                                    ybo[0] = fmaf(wt.x, gv, ybo[0]);
                                    ybo[1] = fmaf(wt.y, gv, ybo[1]);
                                    if (seq && live && (c & 1) == L) atomicAdd(a.gA + ((size_t)k * H0 + u) * a.B + b0 + rrow, gv);
                                }
                            }
                        };
                        int u = ua & ~3;
                        const int uend = (ub + 3) & ~3;
                        while (u < uend) {
                            const int left = uend - u;
                            if (left >= 16) { cblock(std::integral_constant<int, 16>{}, u); u += 16; }
                            else if (left == 12) { cblock(std::integral_constant<int, 12>{}, u); u += 12; }
                            else if (left >= 8) { cblock(std::integral_constant<int, 8>{}, u); u += 8; }
                            else { cblock(std::integral_constant<int, 4>{}, u); u += 4; }
                        }
                    }
                }
                if (seq) {
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj)
                        if (jj < D) vecp[oG + perm[jj]] = zb[jj];
                } else {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (2 * L + i < D) vecp[oG + permy[2 * L + i]] = ybo[i];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"((uint32_t)a.tm_cols)
                     : "memory");
}

}  // namespace rabi
