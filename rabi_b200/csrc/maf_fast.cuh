// Fast path for two-hidden-layer MADE conditioners (config/model/maf_pyro.yaml: hidden_dims [100, 100]) whose
// per-transform weights fit in shared memory: fused MC reverse-KL forward (+ reverse-mode input gradient).
//
// Layout / schedule (DESIGN.md "k_fast"):
//  * a CTA owns a tile of pairs; every thread owns RP pairs (slots tid, tid+T, ...) and keeps ONE private
//    activation buffer per pair in shared memory ([slot][AS], AS/4 odd => conflict-free 128-bit access);
//  * the weights of ONE transform (theta part of layer 0, layer 1, unit-major output layer) are staged in shared
//    memory and read as warp-broadcast 128-bit loads; phases visit the transforms in the order
//    p0 p1 p2 | q2 q1 q0 | q0' q1' q2' | p2' p1' p0'  so that only 9 of 12 phase changes reload weights;
//  * layer 1 is never stored: each block of NR units is produced in registers (packed fma.rn.f32x2 along the
//    dot-product direction), ReLU'd into a bit mask and folded straight into the 2D output accumulators;
//  * the sequential sampling pass evaluates, at step j, only the units of MADE degree j (one pass-equivalent);
//  * reverse mode is pull-style: the buffer holds the layer-1 cotangents, layer-0 cotangents are produced in
//    registers per block, pushed into d/dy and reduced over the S samples of a row with global float reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "maf_image.h"

// Tuning knobs of the product loops (compile-time; the defaults are the measured best on lotka_volterra shapes).
#define RABI_TM_CLOBBER "memory"
#ifndef RABI_X_DOT_UNROLL
#define RABI_X_DOT_UNROLL 2
#endif
#ifndef RABI_X_TDOT_UNROLL
#define RABI_X_TDOT_UNROLL 2
#endif

namespace rabi {
constexpr int kDotUnroll = RABI_X_DOT_UNROLL, kTdotUnroll = RABI_X_TDOT_UNROLL;

struct FastArgs {
    const float* A_p;
    const float* A_q;
    const float* z;     // [S,B,D]
    float* diff;        // [S,B] log p - log q (optional in GRAD mode); RSAMPLE: log q(own sample); LOGPROB: log q(theta)
    float* theta_out;   // RSAMPLE: [S,B,D]
    float* gA;          // GRAD: [K,H0,B] d loss / d A_p, zeroed by the caller: the S samples of a row add their parts with
                        // global reductions (shared-memory float atomics are CAS loops, and a tile needs flush barriers)
    int S, B;
    float w;
    int TB;             // GRAD: rows per tile; FWD: unused
    int NP;             // FWD: pairs per tile
    int ntiles;
    int AS, VS, BWS;    // per-slot strides: activations (floats), vectors (floats), relu bit words
    int wfloats, wints; // per-transform weight / index block sizes (identical for all transforms)
    int tm_cols;        // TMEM columns to allocate (power of two >= 32) when the activations live in tensor memory
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// ---- activation buffer of one pair: shared memory, or a private row of TMEM columns -----------------------
// TMEM (tensor memory, 512 columns x 128 lanes x 32 bit per SM) is idle in this FMA-bound kernel; .32x32b accesses
// give every thread of a warp its own lane, so a column range is a thread-private array with 128-bit vector access.
// Keeping the activations there frees ~45% of the per-pair shared memory => twice the resident warps.
// kind: 0 = shared memory, 1 = tensor memory, 2 = global-memory scratch (maf_deep.cuh)
template <int AK>
struct ActBuf;
template <>
struct ActBuf<0> {
    float* p;
    __device__ __forceinline__ float4 ld4(int i) const { return *reinterpret_cast<const float4*>(p + i); }
    __device__ __forceinline__ void st4(int i, float4 v) const { *reinterpret_cast<float4*>(p + i) = v; }
    __device__ __forceinline__ void st1(int i, float v) const { p[i] = v; }
    __device__ __forceinline__ static void wait_ld(float4&) {}
    __device__ __forceinline__ static void wait_st() {}
};
template <>
struct ActBuf<1> {
    uint32_t t;  // TMEM address of column 0 of this thread's row: (lane quadrant base << 16) | first column
    __device__ __forceinline__ float4 ld4(int i) const {
        float4 v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(t + (uint32_t)i)
                     : RABI_TM_CLOBBER);
        return v;
    }
    __device__ __forceinline__ void st4(int i, float4 v) const {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t + (uint32_t)i), "f"(v.x),
                     "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
    __device__ __forceinline__ void st1(int i, float v) const {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(t + (uint32_t)i), "f"(v) : "memory");
    }
    // the loaded registers are only valid after wait::ld; tie them to the wait so nothing is scheduled above it
    __device__ __forceinline__ static void wait_ld(float4& v) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)::RABI_TM_CLOBBER);
    }
    __device__ __forceinline__ static void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
};

// out[r][c] = sum_{i<e4} W[(min(c,cmax))*ld + i] * act_r[i]     (NR rows x RP pairs, e4 multiple of 4, e4 >= 4)
template <int RP, int NR, int TM, int LDC>
__device__ __forceinline__ void dot_block(const float* __restrict__ W, int ld_rt, int cmax, int e4,
                                          const ActBuf<TM> (&ab)[RP], float (&out)[RP][NR]) {
    float2 acc[RP][NR];
#pragma unroll
    for (int r = 0; r < RP; ++r)
#pragma unroll
        for (int c = 0; c < NR; ++c) acc[r][c] = make_float2(0.f, 0.f);
    const int ld = LDC ? LDC : ld_rt;  // compile-time leading dimension: row addresses become immediates
    const float* wr[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) wr[c] = W + (size_t)min(c, cmax) * ld;
    float4 hn[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) hn[r] = ab[r].ld4(0);
#pragma unroll(kDotUnroll)
    for (int i = 0; i < e4; i += 4) {
        float4 h[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            ActBuf<TM>::wait_ld(hn[r]);
            h[r] = hn[r];
        }
        {  // unconditional look-ahead: the quad after the last one is inside the CTA's shared / tensor memory, never used
#pragma unroll
            for (int r = 0; r < RP; ++r) hn[r] = ab[r].ld4(i + 4);
        }
#pragma unroll
        for (int c = 0; c < NR; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(wr[c] + i);
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                acc[r][c] = ffma2(make_float2(w.x, w.y), make_float2(h[r].x, h[r].y), acc[r][c]);
                acc[r][c] = ffma2(make_float2(w.z, w.w), make_float2(h[r].z, h[r].w), acc[r][c]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RP; ++r)
#pragma unroll
        for (int c = 0; c < NR; ++c) out[r][c] = acc[r][c].x + acc[r][c].y;
}

// out[r][c] = sum_{v in [va4, vb4)} W[v*ld + c] * g_r[v]     (NC columns x RP pairs; W offset to the first column;
// va4 < vb4 multiples of 4; rows >= the layer's width are zero padding)
template <int RP, int NC, int TM, int LDC>
__device__ __forceinline__ void tdot_block(const float* __restrict__ W, int ld_rt, int va4, int vb4,
                                           const ActBuf<TM> (&ab)[RP], float (&out)[RP][NC]) {
    float2 acc[RP][NC / 2];
#pragma unroll
    for (int r = 0; r < RP; ++r)
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) acc[r][c] = make_float2(0.f, 0.f);
    const int ld = LDC ? LDC : ld_rt;
    const float* wbase = W + (size_t)va4 * ld;
    float4 gn[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) gn[r] = ab[r].ld4(va4);
#pragma unroll(kTdotUnroll)
    for (int v = va4; v < vb4; v += 4, wbase += 4 * ld) {
        float4 g4[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            ActBuf<TM>::wait_ld(gn[r]);
            g4[r] = gn[r];
        }
        {
#pragma unroll
            for (int r = 0; r < RP; ++r) gn[r] = ab[r].ld4(v + 4);
        }
#pragma unroll
        for (int dv = 0; dv < 4; ++dv) {
            const float* wrow = wbase + dv * ld;
#pragma unroll
            for (int cc = 0; cc < NC / 4; ++cc) {
                const float4 w = *reinterpret_cast<const float4*>(wrow + 4 * cc);
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    const float gv = dv == 0 ? g4[r].x : dv == 1 ? g4[r].y : dv == 2 ? g4[r].z : g4[r].w;
                    const float2 gg = make_float2(gv, gv);
                    acc[r][2 * cc] = ffma2(make_float2(w.x, w.y), gg, acc[r][2 * cc]);
                    acc[r][2 * cc + 1] = ffma2(make_float2(w.z, w.w), gg, acc[r][2 * cc + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RP; ++r)
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) {
            out[r][2 * c] = acc[r][c].x;
            out[r][2 * c + 1] = acc[r][c].y;
        }
}

// FAST_GRAD: reverse KL, gradient w.r.t. the projection of the sampled-from posterior (A_p).
// FAST_FKL : forward-KL attack loss KL(target || q(.|x+delta)): samples come from the fixed target (A_p), the gradient
//            flows only through the density pass (A_q): d/dA_q of -w*log q, emitted by the reverse density phases.
// FAST_RSAMPLE: sampling passes only (theta and its log-density); FAST_LOGPROB: density passes only on given theta.
enum FastMode { FAST_FWD = 0, FAST_GRAD = 1, FAST_FKL = 2, FAST_RSAMPLE = 3, FAST_LOGPROB = 4 };
__host__ __device__ constexpr bool fast_is_grad(int mode) { return mode == FAST_GRAD || mode == FAST_FKL; }

// Relative offsets (floats / ints) of the sections inside one transform's staged block.
struct FastRel {
    int W0t, b1, W1, ld1, WoT, bo;                  // floats, relative to t.W0t
    int perm, permy, deg0, pre1, grp0, grp1, first1; // ints, relative to t.perm
};

template <int MODE, int RP, int DP, bool TM, int LDC>
__global__ void __launch_bounds__(RP != 1 ? 256 : (fast_is_grad(MODE) ? 384 : 512), 1) k_fast(const MafImg g, const FastArgs a, const FastRel rel) {
    extern __shared__ __align__(16) float sm[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int K = g.K, D = g.D, H0 = g.H[0], H1 = g.H[1];
    const int WL = (g.Hmax + 31) >> 5;
    float* Ws = sm;
    int* Ms = reinterpret_cast<int*>(Ws + a.wfloats);
    float* act = reinterpret_cast<float*>(Ms + a.wints);                    // [RP*T][AS], 16-B aligned
    float* vec = act + (TM ? 0 : (size_t)RP * T * a.AS);                    // [RP*T][VS]  (TM: no smem activations)
    uint32_t* bits = reinterpret_cast<uint32_t*>(vec + (size_t)RP * T * a.VS);  // [RP*T][BWS] (GRAD only)

    const float* W0t = Ws + rel.W0t;
    const float* b1 = Ws + rel.b1;
    const float* W1 = Ws + rel.W1;
    const float* WoT = Ws + rel.WoT;   // [H1][2*DP]: m-weights of positions 0..DP-1, then s-weights
    const float* bo = Ws + rel.bo;     // [2*DP]
    const int ld1 = LDC ? LDC : rel.ld1;
    const int* perm = Ms + rel.perm;
    const int* permy = Ms + rel.permy;
    const int* grp0 = Ms + rel.grp0;
    const int* grp1 = Ms + rel.grp1;

    __shared__ uint32_t tmem_base_s;
    if (TM) {  // one warp allocates the tensor-memory columns for the whole CTA
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
    }
    ActBuf<TM> ab[RP];
    float* vecp[RP];
    uint32_t* bitp[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) {
        const int slot = tid + r * T;
        if constexpr (TM) {
            const int warp = tid >> 5;  // lanes 32*(warp%4).. ; columns ((warp/4)*RP + r)*AS ..
            ab[r].t = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(((warp >> 2) * RP + r) * a.AS);
        } else {
            ab[r].p = act + (size_t)slot * a.AS;
        }
        vecp[r] = vec + (size_t)slot * a.VS;
        bitp[r] = bits + (size_t)slot * a.BWS;
        for (int i = 0; i < a.AS; i += 4) ab[r].st4(i, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    ActBuf<TM>::wait_st();
    // vec slots (floats)
    const int oV = 0;                 // v_k, k = 0..K     (variable space)
    const int oU = (K + 1) * D;       // u_k, k = 0..K-1
    const int oSHP = (2 * K + 1) * D; // clamped log-scales, sampling pass (order space)
    const int oSHQ = (3 * K + 1) * D; // clamped log-scales, density pass
    const int oG = (4 * K + 1) * D;   // running cotangent (variable space)
    const float HALF_LOG_2PI = 0.91893853320467274178f;

    int resident = -1;
    auto ensure_weights = [&](int k) {
        if (resident == k) return;
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(g.f + g.t[k].W0t);
            float4* dst = reinterpret_cast<float4*>(Ws);
            const int n4 = a.wfloats >> 2;
            for (int base = tid; base < n4; base += 8 * T) {  // 8 loads in flight per thread
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (base + q * T < n4) v[q] = __ldg(src + base + q * T);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (base + q * T < n4) dst[base + q * T] = v[q];
            }
            const int* msrc = g.m + g.t[k].perm;
            for (int i = tid; i < a.wints; i += T) Ms[i] = __ldg(msrc + i);
        }
        __syncthreads();
        resident = k;
    };

    // -------------------------------------------------------------------------------------------------------
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        bool live[RP];
        int rrow[RP];                 // row inside the tile (GRAD)
        long long pidx[RP];           // flat pair index s*B + b
        const float* Ap[RP];
        const float* Aq[RP];
        int rows_here = 0, b0 = 0;
        if (fast_is_grad(MODE)) {
            b0 = tile * a.TB;
            rows_here = min(a.TB, a.B - b0);
        }
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            const int q = tid + r * T;
            if (fast_is_grad(MODE)) {
                live[r] = q < rows_here * a.S;
                const int s = live[r] ? q / rows_here : 0;
                rrow[r] = live[r] ? q % rows_here : 0;
                pidx[r] = (long long)s * a.B + b0 + rrow[r];
            } else {
                const long long p = (long long)tile * a.NP + q;
                live[r] = q < a.NP && p < (long long)a.S * a.B;
                pidx[r] = live[r] ? p : 0;
                rrow[r] = 0;
            }
            const int b = (int)(pidx[r] % a.B);
            Ap[r] = a.A_p + b;
            Aq[r] = a.A_q + b;
        }
        float logp[RP], logq[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            float ssq = 0.f;
            for (int d = 0; d < D; ++d) {
                const float zv = live[r] ? a.z[(size_t)pidx[r] * D + d] : 0.f;
                vecp[r][oV + (MODE == FAST_LOGPROB ? K * D : 0) + d] = zv;  // LOGPROB: the input is theta = v_K
                ssq = fmaf(zv, zv, ssq);
            }
            logp[r] = MODE == FAST_LOGPROB ? 0.f : -0.5f * ssq - (float)D * HALF_LOG_2PI;
            logq[r] = 0.f;
        }

        // One loop over the 2K (forward) / 4K (forward + reverse) phases; every building block below is
        // instantiated exactly once.  kind: 0 = sampling pass under A_p (k ascending), 1 = density pass under A_q
        // (k descending), 2 = reverse of the density pass (k ascending), 3 = reverse of the sampling pass.
        const int nphase = (MODE == FAST_GRAD ? 4 : MODE == FAST_FKL ? 3 : MODE == FAST_RSAMPLE ? 1 : 2) * K;
        for (int ph = (MODE == FAST_LOGPROB ? K : 0); ph < nphase; ++ph) {
            const int kind = ph / K;
            const int idx = ph - kind * K;
            const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
            ensure_weights(k);
            const bool seq = (kind == 0 || kind == 3);
            const int bs = ((kind == 0 || kind == 3) ? k : K + k) * 2;  // relu-bit slot of (pass, k)

            if (kind < 2) {
                // ===================== forward pass of transform k =====================
                const float* Aptr[RP];
                float zos[RP][DP], y[RP][DP];
                float2 oms[RP][DP];  // (mean, log-scale) accumulators of every position
                uint32_t w0[RP], w1[RP];
                const int src = (k == K - 1) ? oV + K * D : oU + (k + 1) * D;
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    Aptr[r] = (seq ? Ap[r] : Aq[r]) + (size_t)k * H0 * a.B;
                    w0[r] = w1[r] = 0u;
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj) {
                        zos[r][jj] = (seq && jj < D) ? vecp[r][oV + k * D + perm[jj]] : 0.f;
                        y[r][jj] = (!seq && jj < D) ? vecp[r][src + permy[jj]] : 0.f;
                        oms[r][jj] = jj < D ? make_float2(bo[2 * jj], bo[2 * jj + 1]) : make_float2(0.f, 0.f);
                    }
                }
                // ---- the context projection of this transform goes into the activation buffer first (16 loads in
                // flight per pair); layer 0 then updates it in place, so the hot loop has no global-memory dependence
                for (int u0 = 0; u0 < H0; u0 += 16) {
                    float av[RP][16];
                    const float* ap[RP];  // walks the rows u0, u0 + 1, ... of A (idle slots read row 0 of the batch)
#pragma unroll
                    for (int r = 0; r < RP; ++r) ap[r] = Aptr[r] + (size_t)u0 * a.B;
                    if (u0 + 16 <= H0) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
#pragma unroll
                            for (int r = 0; r < RP; ++r) {
                                av[r][c] = __ldg(ap[r]);
                                ap[r] += a.B;
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
#pragma unroll
                            for (int r = 0; r < RP; ++r) {
                                av[r][c] = __ldg(ap[r]);
                                if (u0 + c + 1 < H0) ap[r] += a.B;  // units past the layer repeat its last row
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (u0 + 4 * q < H0) {
#pragma unroll
                            for (int r = 0; r < RP; ++r)
                                ab[r].st4(u0 + 4 * q, make_float4(av[r][4 * q], av[r][4 * q + 1], av[r][4 * q + 2], av[r][4 * q + 3]));
                        }
                }
                ActBuf<TM>::wait_st();
                for (int j = 0; j < D; ++j) {
                    // ---- layer-0 units of degree j: act = relu(A[u] + W0t[u] . y), aligned quads updated in place
                    {
                        const int ua = grp0[j], ub = grp0[j + 1];
                        for (int u0 = ua & ~3; u0 < ub; u0 += 4) {
                            float4 av[RP];
#pragma unroll
                            for (int r = 0; r < RP; ++r) av[r] = ab[r].ld4(u0);
                            float pre[RP][4];
#pragma unroll
                            for (int r = 0; r < RP; ++r) {
                                ActBuf<TM>::wait_ld(av[r]);
                                pre[r][0] = av[r].x; pre[r][1] = av[r].y; pre[r][2] = av[r].z; pre[r][3] = av[r].w;
                            }
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int u = u0 + c;  // W0t has round4(H0) + 4 rows: no clamp
#pragma unroll
                                for (int q4 = 0; q4 < DP / 4; ++q4) {
                                    const float4 w = *reinterpret_cast<const float4*>(W0t + (size_t)u * DP + 4 * q4);
#pragma unroll
                                    for (int r = 0; r < RP; ++r) {
                                        pre[r][c] = fmaf(w.x, y[r][4 * q4], pre[r][c]);
                                        pre[r][c] = fmaf(w.y, y[r][4 * q4 + 1], pre[r][c]);
                                        pre[r][c] = fmaf(w.z, y[r][4 * q4 + 2], pre[r][c]);
                                        pre[r][c] = fmaf(w.w, y[r][4 * q4 + 3], pre[r][c]);
                                    }
                                }
                            }
#pragma unroll
                            for (int r = 0; r < RP; ++r) {
                                float nv[4] = {av[r].x, av[r].y, av[r].z, av[r].w};
                                uint32_t nib = 0u;  // relu bits of the quad (units outside [ua, ub) contribute nothing)
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const int u = u0 + c;
                                    if (u >= ua && u < ub) {
                                        const bool on = pre[r][c] > 0.f;
                                        nv[c] = on ? pre[r][c] : 0.f;
                                        nib |= on ? (1u << c) : 0u;
                                    }
                                }
                                ab[r].st4(u0, make_float4(nv[0], nv[1], nv[2], nv[3]));
                                if (fast_is_grad(MODE)) {  // the word is complete after its last quad / the layer's last quad
                                    w0[r] |= nib << (u0 & 31);
                                    if (ub >= min(u0 + 4, H0) && ((u0 & 31) == 28 || u0 + 4 >= H0)) {
                                        bitp[r][bs * WL + (u0 >> 5)] = w0[r];
                                        w0[r] = 0u;
                                    }
                                }
                            }
                        }
                        ActBuf<TM>::wait_st();
                    }
                    // ---- layer-1 units of degree j (prefix grp0[j+1]) -> relu bits + output accumulators
                    {
                        const int va = grp1[j], vb = grp1[j + 1];
                        const int e4 = (grp0[j + 1] + 3) & ~3;
                        auto block = [&](auto NRtag, int v0, int nv) {
                            constexpr int NR = decltype(NRtag)::value;
                            float h[RP][NR];
                            dot_block<RP, NR, TM, LDC>(W1 + (size_t)v0 * ld1, ld1, NR - 1, e4, ab, h);  // blocks are exact: nv == NR
                            uint32_t bm[RP];
#pragma unroll
                            for (int r = 0; r < RP; ++r) bm[r] = 0u;
#pragma unroll
                            for (int c = 0; c < NR; ++c) {
                                {
                                    const int v = v0 + c;
                                    const float bias = b1[v];
                                    const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)v * 2 * DP);
                                    float4 wq[DP / 2];  // (m_0, s_0, m_1, s_1), ...
#pragma unroll
                                    for (int q2 = 0; q2 < DP / 2; ++q2) wq[q2] = wo[q2];
#pragma unroll
                                    for (int r = 0; r < RP; ++r) {
                                        const float val = h[r][c] + bias;
                                        const bool on = val > 0.f;
                                        const float hv = on ? val : 0.f;
                                        const float2 hh = make_float2(hv, hv);
#pragma unroll
                                        for (int q2 = 0; q2 < DP / 2; ++q2) {
                                            oms[r][2 * q2] = ffma2(make_float2(wq[q2].x, wq[q2].y), hh, oms[r][2 * q2]);
                                            oms[r][2 * q2 + 1] = ffma2(make_float2(wq[q2].z, wq[q2].w), hh, oms[r][2 * q2 + 1]);
                                        }
                                        if (fast_is_grad(MODE)) bm[r] |= on ? (1u << c) : 0u;
                                    }
                                }
                            }
                            if (fast_is_grad(MODE)) {  // relu bits of the block -> the running word; store it when it is complete
                                const int sh = v0 & 31;
#pragma unroll
                                for (int r = 0; r < RP; ++r) {
                                    w1[r] |= bm[r] << sh;
                                    if (sh + NR >= 32) {
                                        bitp[r][(bs + 1) * WL + (v0 >> 5)] = w1[r];
                                        w1[r] = sh + NR > 32 ? bm[r] >> (32 - sh) : 0u;
                                    }
                                    if (v0 + NR == H1 && (H1 & 31) != 0) {
                                        bitp[r][(bs + 1) * WL + ((H1 - 1) >> 5)] = w1[r];
                                        w1[r] = 0u;
                                    }
                                }
                            }
                        };
                        int v = va;
                        for (; v + 16 <= vb; v += 16) block(std::integral_constant<int, 16>{}, v, 16);
                        // a remainder of 9..15 units as ONE exact block (25 units per degree at lotka_volterra shapes = 16 + 9,
                        // 10 at gaussian_linear shapes): every extra block is another pass over the activations
#define RABI_REM(N) case N: block(std::integral_constant<int, N>{}, v, N); v += N; break;
                        switch (vb - v) {
                            RABI_REM(9) RABI_REM(10) RABI_REM(11) RABI_REM(12) RABI_REM(13) RABI_REM(14) RABI_REM(15)
                            default: break;
                        }
#undef RABI_REM
                        if (v + 8 <= vb) { block(std::integral_constant<int, 8>{}, v, 8); v += 8; }
                        if (v + 4 <= vb) { block(std::integral_constant<int, 4>{}, v, 4); v += 4; }
                        if (v + 2 <= vb) { block(std::integral_constant<int, 2>{}, v, 2); v += 2; }
                        if (v < vb) block(std::integral_constant<int, 1>{}, v, 1);
                    }
                    if (seq) {  // position j is final: y_j = (z_j - m_j) exp(-clamp(s_j))
#pragma unroll
                        for (int r = 0; r < RP; ++r) {
                            float m = 0.f, s = 0.f, zj = 0.f;
#pragma unroll
                            for (int jj = 0; jj < DP; ++jj) {
                                m = jj == j ? oms[r][jj].x : m;
                                s = jj == j ? oms[r][jj].y : s;
                                zj = jj == j ? zos[r][jj] : zj;
                            }
                            const float sh = fminf(fmaxf(s, g.lo), g.hi);
                            const float yj = (zj - m) * expf(-sh);
#pragma unroll
                            for (int jj = 0; jj < DP; ++jj) y[r][jj] = jj == j ? yj : y[r][jj];
                            logp[r] += sh;
                            vecp[r][oSHP + k * D + j] = sh;
                            vecp[r][oV + (k + 1) * D + permy[j]] = yj;
                        }
                    }
                }
                if (!seq) {  // u_k = exp(clamp(s)) * u_{k+1} + m
#pragma unroll
                    for (int r = 0; r < RP; ++r) {
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) {
                            if (jj < D) {
                                const float sh = fminf(fmaxf(oms[r][jj].y, g.lo), g.hi);
                                vecp[r][oSHQ + k * D + jj] = sh;
                                logq[r] += sh;
                                vecp[r][oU + k * D + perm[jj]] = fmaf(expf(sh), y[r][jj], oms[r][jj].x);
                            }
                        }
                    }
                    if (k == 0) {  // end of the density chain: base log-density, KL term, reverse-mode seed
#pragma unroll
                        for (int r = 0; r < RP; ++r) {
                            float ssq = 0.f;
                            for (int d = 0; d < D; ++d) {
                                const float u = vecp[r][oU + d];
                                ssq = fmaf(u, u, ssq);
                                if (fast_is_grad(MODE)) vecp[r][oG + d] = a.w * u;  // d(-w log q)/d u_0
                            }
                            logq[r] += -0.5f * ssq - (float)D * HALF_LOG_2PI;
                            if (a.diff && live[r]) a.diff[pidx[r]] = MODE == FAST_LOGPROB ? logq[r] : logp[r] - logq[r];
                        }
                    }
                }
            } else {
                // ===================== reverse pass of transform k =====================
                const float w = a.w;
                float yb[RP][DP], yv[RP][DP], sh[RP][DP], mb[RP][DP], sb[RP][DP], zb[RP][DP];
                const int src = (k == K - 1) ? oV + K * D : oU + (k + 1) * D;
#pragma unroll
                for (int r = 0; r < RP; ++r)
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj) {
                        zb[r][jj] = 0.f;
                        if (seq) {
                            yb[r][jj] = jj < D ? vecp[r][oG + permy[jj]] : 0.f;
                            yv[r][jj] = jj < D ? vecp[r][oV + (k + 1) * D + permy[jj]] : 0.f;
                            sh[r][jj] = jj < D ? vecp[r][oSHP + k * D + jj] : 0.f;
                            mb[r][jj] = sb[r][jj] = 0.f;
                        } else if (jj < D) {
                            const float ub = vecp[r][oG + perm[jj]];
                            const float es = expf(vecp[r][oSHQ + k * D + jj]);
                            const float yin = vecp[r][src + permy[jj]];
                            mb[r][jj] = ub;
                            sb[r][jj] = fmaf(ub * es, yin, -w);
                            yb[r][jj] = ub * es;
                            yv[r][jj] = sh[r][jj] = 0.f;
                        } else {
                            mb[r][jj] = sb[r][jj] = yb[r][jj] = yv[r][jj] = sh[r][jj] = 0.f;
                        }
                    }
                for (int j = D - 1; j >= 0; --j) {
                    if (seq) {
#pragma unroll
                        for (int r = 0; r < RP; ++r) {
                            float ybj = 0.f, shj = 0.f, yj = 0.f;
#pragma unroll
                            for (int jj = 0; jj < DP; ++jj) {
                                ybj = jj == j ? yb[r][jj] : ybj;
                                shj = jj == j ? sh[r][jj] : shj;
                                yj = jj == j ? yv[r][jj] : yj;
                            }
                            const float zbj = ybj * expf(-shj);
#pragma unroll
                            for (int jj = 0; jj < DP; ++jj) {
                                zb[r][jj] = jj == j ? zbj : zb[r][jj];
                                mb[r][jj] = jj == j ? -zbj : mb[r][jj];
                                sb[r][jj] = jj == j ? fmaf(-ybj, yj, w) : sb[r][jj];
                            }
                        }
                    }
                    // ---- layer-1 cotangents of the degree-j units: g = relu'(.) * (WoT[v] . (mbar, sbar)) -> buffer
                    {
                        const int va = grp1[j], vb = grp1[j + 1];
                        for (int v0 = va & ~3; v0 < vb; v0 += 4) {  // aligned quads: one relu nibble and one 128-bit store each
                            float gs[RP][4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)(v0 + c) * 2 * DP);  // rows < round4(H1) + 4
                                float2 g2[RP];
#pragma unroll
                                for (int r = 0; r < RP; ++r) g2[r] = make_float2(0.f, 0.f);
#pragma unroll
                                for (int q2 = 0; q2 < DP / 2; ++q2) {
                                    const float4 wq = wo[q2];  // (m_j, s_j, m_j+1, s_j+1)
#pragma unroll
                                    for (int r = 0; r < RP; ++r) {
                                        g2[r] = ffma2(make_float2(wq.x, wq.y), make_float2(mb[r][2 * q2], sb[r][2 * q2]), g2[r]);
                                        g2[r] = ffma2(make_float2(wq.z, wq.w), make_float2(mb[r][2 * q2 + 1], sb[r][2 * q2 + 1]), g2[r]);
                                    }
                                }
#pragma unroll
                                for (int r = 0; r < RP; ++r) gs[r][c] = g2[r].x + g2[r].y;
                            }
#pragma unroll
                            for (int r = 0; r < RP; ++r) {
                                const uint32_t nib = bitp[r][(bs + 1) * WL + (v0 >> 5)] >> (v0 & 31);
#pragma unroll
                                for (int c = 0; c < 4; ++c) gs[r][c] = ((nib >> c) & 1u) ? gs[r][c] : 0.f;
                            }
                            if (v0 >= va && v0 + 4 <= vb) {
#pragma unroll
                                for (int r = 0; r < RP; ++r) ab[r].st4(v0, make_float4(gs[r][0], gs[r][1], gs[r][2], gs[r][3]));
                            } else {  // a quad shared with a neighbouring degree group: only the own units
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const int v = v0 + c;
                                    if (v >= va && v < vb) {
#pragma unroll
                                        for (int r = 0; r < RP; ++r) ab[r].st1(v, gs[r][c]);
                                    }
                                }
                            }
                        }
                        ActBuf<TM>::wait_st();
                    }
                    // ---- layer-0 cotangents of the degree-j units (consumed by layer-1 rows >= grp1[j])
                    {
                        const int ua = grp0[j], ub = grp0[j + 1];
                        const int va4 = grp1[j] & ~3, vb4 = (H1 + 3) & ~3;
                        auto cblock = [&](auto NCtag, int u0) {
                            constexpr int NC = decltype(NCtag)::value;
                            float g0[RP][NC];
                            tdot_block<RP, NC, TM, LDC>(W1 + u0, ld1, va4, vb4, ab, g0);
                            uint32_t wb[RP];
#pragma unroll
                            for (int r = 0; r < RP; ++r)  // 32 mask bits starting at unit u0 (a block may straddle two words; the
                                                          // slot array ends with a padding word, so word + 1 is readable)
                                wb[r] = __funnelshift_r(bitp[r][bs * WL + (u0 >> 5)], bitp[r][bs * WL + (u0 >> 5) + 1], u0 & 31);
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                const int u = u0 + c;
                                if (u >= ua && u < ub) {
                                    float wrow[DP];
#pragma unroll
                                    for (int q4 = 0; q4 < DP / 4; ++q4) {
                                        const float4 w4 = *reinterpret_cast<const float4*>(W0t + (size_t)u * DP + 4 * q4);
                                        wrow[4 * q4] = w4.x; wrow[4 * q4 + 1] = w4.y; wrow[4 * q4 + 2] = w4.z; wrow[4 * q4 + 3] = w4.w;
                                    }
#pragma unroll
                                    for (int r = 0; r < RP; ++r) {
                                        const float gv = ((wb[r] >> c) & 1u) ? g0[r][c] : 0.f;
#pragma unroll
                                        for (int jj = 0; jj < DP; ++jj) yb[r][jj] = fmaf(wrow[jj], gv, yb[r][jj]);
                                        if ((seq || MODE == FAST_FKL) && live[r])  // sum over the S samples of the row: RED.ADD.F32 into the zeroed gA
                                            atomicAdd(a.gA + ((size_t)k * H0 + u) * a.B + b0 + rrow[r], gv);
                                    }
                                }
                            }
                        };
                        int u = ua & ~3;
                        const int uend = (ub + 3) & ~3;
                        while (u < uend) {  // as few passes over the layer-1 cotangents as possible
                            const int left = uend - u;
                            if (left >= 16) { cblock(std::integral_constant<int, 16>{}, u); u += 16; }
                            else if (left == 12) { cblock(std::integral_constant<int, 12>{}, u); u += 12; }
                            else if (left >= 8) { cblock(std::integral_constant<int, 8>{}, u); u += 8; }
                            else { cblock(std::integral_constant<int, 4>{}, u); u += 4; }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < RP; ++r)
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj)
                        if (jj < D) {
                            if (seq) vecp[r][oG + perm[jj]] = zb[r][jj];
                            else vecp[r][oG + permy[jj]] = yb[r][jj];
                        }
            }
        }
        if (MODE == FAST_RSAMPLE) {
#pragma unroll
            for (int r = 0; r < RP; ++r)
                if (live[r]) {
                    for (int d = 0; d < D; ++d) a.theta_out[(size_t)pidx[r] * D + d] = vecp[r][oV + K * D + d];
                    if (a.diff) a.diff[pidx[r]] = logp[r];
                }
        }
    }
    if (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"((uint32_t)a.tm_cols)
                         : "memory");
    }
}

}  // namespace rabi
