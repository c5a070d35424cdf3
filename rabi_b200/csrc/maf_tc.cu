// Tensor-core path for two-hidden-layer MADE conditioners (config/model/maf_pyro.yaml: hidden_dims [100, 100]).
//
// Same arithmetic as k_fast (maf_fast.cuh; reference: Pyro ConditionalAutoRegressiveNN / AffineAutoregressive as driven
// by rbi/rbi/models/transforms.py:37-46, general_kl_divergence rbi/rbi/loss/kl_div.py:38-56), but the layer-0 ->
// layer-1 contraction (91 % of the masked MACs at lotka_volterra shapes) runs on the 5th-generation tensor cores:
//
//  * a CTA holds up to 3 TILES of 128 (observation row, MC sample) pairs, one warpgroup each; thread = pair = TMEM lane;
//  * per tile 168 tensor-memory columns: [0, 104) fp32 accumulators of the layer being produced (column = unit),
//    [104, 136) / [136, 168) the TF32 (hi, lo) halves of up to 32 staged inputs;
//  * "group push": the threads compute a group of layer-0 activations (ReLU(A[u] + W0t[u].y), FMA code, D is tiny),
//    split them into (hi, lo), store them to the staging columns (tcgen05.st); ONE elected thread issues the 3xTF32
//    chain  acc[:, n0:] += h_hi*W_hi + h_lo*W_hi + h_hi*W_lo  (tcgen05.mma.cta_group::1.kind::tf32, M = 128 pairs,
//    A from tensor memory, B = weight strip in shared memory, K-major no-swizzle [k/4][n][4] layout) and commits to the
//    tile's mbarrier; the threads wait and read the accumulator columns that are final now (tcgen05.ld);
//  * the sequential sampling pass pushes, at autoregressive step j, only the layer-0 units of MADE degree j into the
//    accumulator columns of the layer-1 units of degree >= j (one pass-equivalent, exactly the masked MACs, N shrinks
//    with j); the density pass pushes all units in chunks of 32; the reverse passes run the same pattern on the
//    transposed weight copy (cotangents of layer 1 pushed into the columns of layer 0);
//  * the weights of one (transform, orientation) -- hi and lo halves plus the small FMA-side arrays and the chunk
//    tables -- are ONE contiguous block of the packed tensor-core image and arrive in shared memory by cp.async.bulk
//    (TMA bulk copy, completion on an mbarrier); phases visit the transforms as p0 p1 p2 | q2 q1 q0 | q0' q1' q2' |
//    p2' p1' p0', so 10 blocks are loaded per attack iteration.
//
// 3xTF32: a = a_hi + a_lo with a_hi = a & 0xffffe000 (exact TF32), products a_hi*b_hi + a_lo*b_hi + a_hi*b_lo accumulate
// in fp32; measured 1.4e-6 relative to the exact product (profiles/microbench/tc_push_probe.cu), i.e. fp32-level.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "maf_tc.h"

namespace rabi {

constexpr int TC_KC = 32;          // staged inputs per push
constexpr int TC_ACC = 104;        // accumulator columns (max units per hidden layer on this path)
constexpr int TC_SH = TC_ACC;      // staging columns, hi half
constexpr int TC_SL = TC_ACC + TC_KC;
constexpr int TC_TILE_COLS = TC_ACC + 2 * TC_KC;  // 168
constexpr int TC_MAXTILES = 3;
constexpr int TC_WL = 4;           // relu-mask words per (pass, transform, layer)
constexpr int TC_MAXCHUNK = 64;    // capacity of one chunk list

struct TcChunk {
    int u0, nu;   // staged (aligned) input-unit range [u0, u0 + nu), nu a multiple of 8, <= 32
    int ua, ub;   // units of the range that really belong to this push (the others are staged as zeros)
    int n0, N;    // accumulator columns [n0, n0 + N) receive the product (n0, N multiples of 8)
    int pad0, pad1;
};

// word offsets inside one (transform, orientation) block; identical for all transforms
struct TcBlk {
    int whi, wlo;        // [NPK/4][NPN][4] TF32 hi / lo halves of layer 1 (orientation 0: n = layer-1 unit, k = layer-0 unit)
    int NPN, NPK;
    int w0t, b1, wot, bo;  // [R0][DP], [R1], [R1][2*DP], [2*DP]
    int perm, permy, grp0, grp1;
    int nfull, cfull, seqstart, cseq;
    int words;
};

struct TcArgs {
    const float* A_p;
    const float* A_q;
    const float* z;
    float* diff;
    float* theta_out;
    float* gA;
    int S, B;
    float w;
    int ntiles;
    long long npairs;
    const float* img;    // [K][2][blk_words]
    TcBlk blk[2];
    int blk_words;
    uint32_t* bits;      // [2K*2][TC_WL][bits_stride]
    int bits_stride;
    int VS;
};

// ------------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // K direction: next 4-float k chunk
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // N direction: next group of 8 rows
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M x N
__device__ __forceinline__ uint32_t tc_make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    long long t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(tc_smem_u32(mbar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins == 1024u) t0 = clock64();
        if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 4000000000LL) {  // ~2 s: never hang the GPU
            printf("rabi_b200 k_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ float4 tc_ld4(uint32_t addr) {
    float4 v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_wait_ld(float4& v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) :: "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float tc_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float2 tc_ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
// relu-mask words of one layer (<= 128 units) held in registers
__device__ __forceinline__ void tc_bits_merge(uint32_t (&bw)[TC_WL], int u0, uint32_t m) {
    const int w = u0 >> 5, sh = u0 & 31;
    const uint32_t lo = m << sh, hi = sh ? (m >> (32 - sh)) : 0u;
#pragma unroll
    for (int i = 0; i < TC_WL; ++i) bw[i] |= (i == w ? lo : 0u) | (i == w + 1 ? hi : 0u);
}
__device__ __forceinline__ uint32_t tc_bits_extract(const uint32_t (&bw)[TC_WL], int u0) {  // 32 mask bits from unit u0 on
    const int w = u0 >> 5, sh = u0 & 31;
    uint32_t a = 0u, b = 0u;
#pragma unroll
    for (int i = 0; i < TC_WL; ++i) {
        a = i == w ? bw[i] : a;
        b = i == w + 1 ? bw[i] : b;
    }
    return __funnelshift_r(a, b, sh);
}

// ------------------------------------------------------------------------------------------------- the kernel
template <int MODE, int DP>
__global__ void __launch_bounds__(TC_MAXTILES * 128, 1) k_tc(const MafImg g, const TcArgs a) {
    extern __shared__ __align__(128) float sm[];
    constexpr bool GRADM = MODE == TC_GRAD || MODE == TC_FKL;
    float* Wb = sm;                                   // the resident (transform, orientation) block
    const int* Wi = reinterpret_cast<const int*>(sm);
    float* vec = sm + a.blk_words;                    // [threads][VS]
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar[TC_MAXTILES + 1];   // one per tile (MMA completion) + the block loader
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2, lane128 = tid & 127;
    const int NT = blockDim.x >> 7;
    const int K = g.K, D = g.D, H0 = g.H[0], H1 = g.H[1];
    const float HALF_LOG_2PI = 0.91893853320467274178f;

    if (tid == 0) {
        for (int i = 0; i <= TC_MAXTILES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(&mbar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(tc_smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tcol = tmem_base_s + (uint32_t)(wg * TC_TILE_COLS);          // lane 0, first column of the tile
    const uint32_t trow = tcol + ((uint32_t)(32 * (warp & 3)) << 16);           // this thread's lane quadrant
    float* vecp = vec + (size_t)tid * a.VS;
    const int gslot = blockIdx.x * blockDim.x + tid;
    uint64_t* my_mbar = &mbar[wg];
    uint32_t mphase = 0, wphase = 0;

    const int oV = 0, oU = (K + 1) * D, oSHP = (2 * K + 1) * D, oSHQ = (3 * K + 1) * D, oG = (4 * K + 1) * D;

    int resident = -1;
    auto ensure_block = [&](int k, int ori) {
        const int id = 2 * k + ori;
        if (resident == id) return;
        __syncthreads();   // every tile is done with the old block (their MMAs completed: each push is waited for)
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)a.blk[ori].words * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tc_smem_u32(&mbar[TC_MAXTILES])), "r"(bytes) : "memory");
            const char* src = reinterpret_cast<const char*>(a.img + (size_t)id * a.blk_words);
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t n = min(bytes - off, 32768u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(tc_smem_u32(sm) + off), "l"(src + off), "r"(n), "r"(tc_smem_u32(&mbar[TC_MAXTILES])) : "memory");
            }
        }
        tc_mbar_wait(&mbar[TC_MAXTILES], wphase);
        wphase ^= 1u;
        resident = id;
    };

    // one push: the staged inputs of chunk c (this tile) times the weight strip, into accumulator columns [n0, n0 + N)
    auto push = [&](const TcBlk& bk, const TcChunk& c, bool first) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
        if (lane128 == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = tc_make_idesc(128, c.N);
            const uint32_t lbo = (uint32_t)bk.NPN * 16u;
            const uint32_t strip = (uint32_t)(((c.u0 >> 2) * bk.NPN + c.n0) * 16);
            uint32_t acc = first ? 0u : 1u;
#pragma unroll
            for (int t = 0; t < 3; ++t) {   // (h_hi, W_hi), (h_lo, W_hi), (h_hi, W_lo)
                const uint32_t wbase = tc_smem_u32(Wb) + 4u * (uint32_t)(t == 2 ? bk.wlo : bk.whi) + strip;
                const uint32_t acol = tcol + (uint32_t)(t == 1 ? TC_SL : TC_SH);
                for (int kk = 0; kk < c.nu; kk += 8) {
                    tc_mma_ts(tcol + (uint32_t)c.n0, acol + (uint32_t)kk, tc_make_desc(wbase + (uint32_t)(kk >> 2) * lbo, lbo, 128u), idesc, acc);
                    acc = 1u;
                }
            }
            tc_commit(my_mbar);
        }
        tc_mbar_wait(my_mbar, mphase);
        mphase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    const int rounds = (a.ntiles + (int)gridDim.x * NT - 1) / ((int)gridDim.x * NT);
    for (int round = 0; round < rounds; ++round) {
        const int tile = (round * NT + wg) * (int)gridDim.x + (int)blockIdx.x;
        const bool tile_live = tile < a.ntiles;
        const long long p = (long long)tile * 128 + lane128;
        const bool live = tile_live && p < a.npairs;
        const long long pidx = live ? p : 0;
        const int b = (int)(pidx % a.B);
        float logp = 0.f, logq = 0.f;
        if (tile_live) {
            float ssq = 0.f;
            for (int d = 0; d < D; ++d) {
                const float zv = live ? a.z[(size_t)pidx * D + d] : 0.f;
                vecp[oV + (MODE == TC_LOGPROB ? K * D : 0) + d] = zv;
                ssq = fmaf(zv, zv, ssq);
            }
            logp = MODE == TC_LOGPROB ? 0.f : -0.5f * ssq - (float)D * HALF_LOG_2PI;
        }
        const int nphase = (MODE == TC_GRAD ? 4 : MODE == TC_FKL ? 3 : MODE == TC_RSAMPLE ? 1 : 2) * K;
        for (int ph = (MODE == TC_LOGPROB ? K : 0); ph < nphase; ++ph) {
            const int kind = ph / K;
            const int idx = ph - kind * K;
            const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
            const bool seq = (kind == 0 || kind == 3);
            const int ori = kind < 2 ? 0 : 1;
            ensure_block(k, ori);
            if (!tile_live) continue;
            const TcBlk& bk = a.blk[ori];
            const float* W0t = Wb + bk.w0t;
            const float* b1 = Wb + bk.b1;
            const float* WoT = Wb + bk.wot;
            const float* bo = Wb + bk.bo;
            const int* perm = Wi + bk.perm;
            const int* permy = Wi + bk.permy;
            const int* grp0 = Wi + bk.grp0;
            const int* grp1 = Wi + bk.grp1;
            const int* seqstart = Wi + bk.seqstart;
            const TcChunk* cseq = reinterpret_cast<const TcChunk*>(Wi + bk.cseq);
            const TcChunk* cfull = reinterpret_cast<const TcChunk*>(Wi + bk.cfull);
            const int nfull = Wi[bk.nfull];
            const int bs = (seq ? k : K + k) * 2;   // relu-mask slots of (pass, k): layer 0, layer 1
            const int src = (k == K - 1) ? oV + K * D : oU + (k + 1) * D;
            const int nsteps = seq ? D : 1;

            if (kind < 2) {
                // ===================== forward pass of transform k =====================
                const float* Aptr = (seq ? a.A_p : a.A_q) + (size_t)k * H0 * a.B + b;
                float zos[DP], y[DP];
                float2 oms[DP];
                uint32_t bw0[TC_WL], bw1[TC_WL];
#pragma unroll
                for (int i = 0; i < TC_WL; ++i) bw0[i] = bw1[i] = 0u;
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) {
                    zos[jj] = (seq && jj < D) ? vecp[oV + k * D + perm[jj]] : 0.f;
                    y[jj] = (!seq && jj < D) ? vecp[src + permy[jj]] : 0.f;
                    oms[jj] = jj < D ? make_float2(bo[2 * jj], bo[2 * jj + 1]) : make_float2(0.f, 0.f);
                }
                bool first = true;
                for (int step = 0; step < nsteps; ++step) {
                    const int c_lo = seq ? seqstart[step] : 0, c_hi = seq ? seqstart[step + 1] : nfull;
                    for (int ci = c_lo; ci < c_hi; ++ci) {
                        const TcChunk c = seq ? cseq[ci] : cfull[ci];
                        // ---- layer-0 units of the chunk: h = relu(A[u] + W0t[u] . y), staged as (hi, lo)
                        float av[TC_KC];
#pragma unroll
                        for (int i = 0; i < TC_KC; ++i)
                            av[i] = i < c.nu ? __ldg(Aptr + (size_t)min(c.u0 + i, H0 - 1) * a.B) : 0.f;
                        uint32_t cm = 0u;
#pragma unroll
                        for (int q = 0; q < TC_KC; q += 4) {
                            if (q < c.nu) {
                                float hv[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int u = c.u0 + q + e;
                                    float pre = av[q + e];
#pragma unroll
                                    for (int q4 = 0; q4 < DP / 4; ++q4) {
                                        const float4 w = *reinterpret_cast<const float4*>(W0t + (size_t)u * DP + 4 * q4);
                                        pre = fmaf(w.x, y[4 * q4], pre);
                                        pre = fmaf(w.y, y[4 * q4 + 1], pre);
                                        pre = fmaf(w.z, y[4 * q4 + 2], pre);
                                        pre = fmaf(w.w, y[4 * q4 + 3], pre);
                                    }
                                    const bool on = u >= c.ua && u < c.ub && pre > 0.f;
                                    hv[e] = on ? pre : 0.f;
                                    cm |= on ? (1u << (q + e)) : 0u;
                                }
                                const float h0 = tc_hi(hv[0]), h1 = tc_hi(hv[1]), h2 = tc_hi(hv[2]), h3 = tc_hi(hv[3]);
                                tc_st4(trow + TC_SH + q, h0, h1, h2, h3);
                                tc_st4(trow + TC_SL + q, hv[0] - h0, hv[1] - h1, hv[2] - h2, hv[3] - h3);
                            }
                        }
                        if (GRADM) tc_bits_merge(bw0, c.u0, cm);
                        push(bk, c, first);
                        first = false;
                    }
                    // ---- layer-1 units that are final now: relu(acc + b1) folded into the (mean, log-scale) accumulators
                    {
                        const int va = seq ? grp1[step] : 0, vb = seq ? grp1[step + 1] : H1;
                        float4 nxt = tc_ld4(trow + (uint32_t)(va & ~3));
                        for (int v0 = va & ~3; v0 < vb; v0 += 4) {
                            tc_wait_ld(nxt);
                            const float4 acc = nxt;
                            nxt = tc_ld4(trow + (uint32_t)(v0 + 4));   // look-ahead (stays inside the tile's columns)
                            const float pre[4] = {acc.x, acc.y, acc.z, acc.w};
                            uint32_t nib = 0u;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int v = v0 + e;
                                const float val = pre[e] + b1[v];
                                const bool on = v >= va && v < vb && val > 0.f;
                                const float hv = on ? val : 0.f;
                                nib |= on ? (1u << e) : 0u;
                                const float2 hh = make_float2(hv, hv);
                                const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)v * 2 * DP);
#pragma unroll
                                for (int q2 = 0; q2 < DP / 2; ++q2) {
                                    const float4 wq = wo[q2];   // (m_j, s_j, m_j+1, s_j+1)
                                    oms[2 * q2] = tc_ffma2(make_float2(wq.x, wq.y), hh, oms[2 * q2]);
                                    oms[2 * q2 + 1] = tc_ffma2(make_float2(wq.z, wq.w), hh, oms[2 * q2 + 1]);
                                }
                            }
                            if (GRADM) tc_bits_merge(bw1, v0, nib);
                        }
                        tc_wait_ld(nxt);
                    }
                    if (seq) {  // position `step` is final: y_j = (z_j - m_j) exp(-clamp(s_j))
                        const int j = step;
                        float m = 0.f, s = 0.f, zj = 0.f;
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) {
                            m = jj == j ? oms[jj].x : m;
                            s = jj == j ? oms[jj].y : s;
                            zj = jj == j ? zos[jj] : zj;
                        }
                        const float sh = fminf(fmaxf(s, g.lo), g.hi);
                        const float yj = (zj - m) * expf(-sh);
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) y[jj] = jj == j ? yj : y[jj];
                        logp += sh;
                        vecp[oSHP + k * D + j] = sh;
                        vecp[oV + (k + 1) * D + permy[j]] = yj;
                    }
                }
                if (!seq) {  // u_k = exp(clamp(s)) * u_{k+1} + m
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj) {
                        if (jj < D) {
                            const float sh = fminf(fmaxf(oms[jj].y, g.lo), g.hi);
                            vecp[oSHQ + k * D + jj] = sh;
                            logq += sh;
                            vecp[oU + k * D + perm[jj]] = fmaf(expf(sh), y[jj], oms[jj].x);
                        }
                    }
                    if (k == 0) {  // end of the density chain: base log-density, KL term, reverse-mode seed
                        float ssq = 0.f;
                        for (int d = 0; d < D; ++d) {
                            const float u = vecp[oU + d];
                            ssq = fmaf(u, u, ssq);
                            if (GRADM) vecp[oG + d] = a.w * u;  // d(-w log q)/d u_0
                        }
                        logq += -0.5f * ssq - (float)D * HALF_LOG_2PI;
                        if (a.diff && live) a.diff[pidx] = MODE == TC_LOGPROB ? logq : logp - logq;
                    }
                }
                if (GRADM) {
#pragma unroll
                    for (int i = 0; i < TC_WL; ++i) {
                        a.bits[((size_t)(bs * TC_WL + i)) * a.bits_stride + gslot] = bw0[i];
                        a.bits[((size_t)((bs + 1) * TC_WL + i)) * a.bits_stride + gslot] = bw1[i];
                    }
                }
            } else {
                // ===================== reverse pass of transform k =====================
                const float w = a.w;
                float yb[DP], yv[DP], sh[DP], mb[DP], sb[DP], zb[DP];
                uint32_t bw0[TC_WL], bw1[TC_WL];
#pragma unroll
                for (int i = 0; i < TC_WL; ++i) {
                    bw0[i] = a.bits[((size_t)(bs * TC_WL + i)) * a.bits_stride + gslot];
                    bw1[i] = a.bits[((size_t)((bs + 1) * TC_WL + i)) * a.bits_stride + gslot];
                }
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) {
                    zb[jj] = 0.f;
                    if (seq) {
                        yb[jj] = jj < D ? vecp[oG + permy[jj]] : 0.f;
                        yv[jj] = jj < D ? vecp[oV + (k + 1) * D + permy[jj]] : 0.f;
                        sh[jj] = jj < D ? vecp[oSHP + k * D + jj] : 0.f;
                        mb[jj] = sb[jj] = 0.f;
                    } else if (jj < D) {
                        const float ub = vecp[oG + perm[jj]];
                        const float es = expf(vecp[oSHQ + k * D + jj]);
                        const float yin = vecp[src + permy[jj]];
                        mb[jj] = ub;
                        sb[jj] = fmaf(ub * es, yin, -w);
                        yb[jj] = ub * es;
                        yv[jj] = sh[jj] = 0.f;
                    } else {
                        mb[jj] = sb[jj] = yb[jj] = yv[jj] = sh[jj] = 0.f;
                    }
                }
                bool first = true;
                for (int st = 0; st < nsteps; ++st) {
                    const int step = seq ? D - 1 - st : 0;
                    if (seq) {
                        const int j = step;
                        float ybj = 0.f, shj = 0.f, yj = 0.f;
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) {
                            ybj = jj == j ? yb[jj] : ybj;
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
                    // ---- layer-1 cotangents of the chunk's units: g = relu'(.) * (WoT[v] . (mbar, sbar)), staged as (hi, lo)
                    const int c_lo = seq ? seqstart[step] : 0, c_hi = seq ? seqstart[step + 1] : nfull;
                    for (int ci = c_lo; ci < c_hi; ++ci) {
                        const TcChunk c = seq ? cseq[ci] : cfull[ci];
                        const uint32_t cm = tc_bits_extract(bw1, c.u0);
#pragma unroll
                        for (int q = 0; q < TC_KC; q += 4) {
                            if (q < c.nu) {
                                float gs[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int v = c.u0 + q + e;
                                    const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)v * 2 * DP);
                                    float2 g2 = make_float2(0.f, 0.f);
#pragma unroll
                                    for (int q2 = 0; q2 < DP / 2; ++q2) {
                                        const float4 wq = wo[q2];
                                        g2 = tc_ffma2(make_float2(wq.x, wq.y), make_float2(mb[2 * q2], sb[2 * q2]), g2);
                                        g2 = tc_ffma2(make_float2(wq.z, wq.w), make_float2(mb[2 * q2 + 1], sb[2 * q2 + 1]), g2);
                                    }
                                    const bool on = v >= c.ua && v < c.ub && ((cm >> (q + e)) & 1u);
                                    gs[e] = on ? g2.x + g2.y : 0.f;
                                }
                                const float h0 = tc_hi(gs[0]), h1 = tc_hi(gs[1]), h2 = tc_hi(gs[2]), h3 = tc_hi(gs[3]);
                                tc_st4(trow + TC_SH + q, h0, h1, h2, h3);
                                tc_st4(trow + TC_SL + q, gs[0] - h0, gs[1] - h1, gs[2] - h2, gs[3] - h3);
                            }
                        }
                        push(bk, c, first);
                        first = false;
                    }
                    // ---- layer-0 cotangents that are final now: mask, push into d/dy, emit d/dA
                    {
                        const int ua = seq ? grp0[step] : 0, ub = seq ? grp0[step + 1] : H0;
                        const bool emit = (seq || MODE == TC_FKL) && live;
                        float* gA = a.gA + (size_t)k * H0 * a.B + b;
                        float4 nxt = tc_ld4(trow + (uint32_t)(ua & ~3));
                        for (int u0 = ua & ~3; u0 < ub; u0 += 4) {
                            tc_wait_ld(nxt);
                            const float4 acc = nxt;
                            nxt = tc_ld4(trow + (uint32_t)(u0 + 4));
                            const float g0[4] = {acc.x, acc.y, acc.z, acc.w};
                            const uint32_t nib = tc_bits_extract(bw0, u0);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int u = u0 + e;
                                const bool on = u >= ua && u < ub && ((nib >> e) & 1u);
                                const float gv = on ? g0[e] : 0.f;
#pragma unroll
                                for (int q4 = 0; q4 < DP / 4; ++q4) {
                                    const float4 w4 = *reinterpret_cast<const float4*>(W0t + (size_t)u * DP + 4 * q4);
                                    yb[4 * q4] = fmaf(w4.x, gv, yb[4 * q4]);
                                    yb[4 * q4 + 1] = fmaf(w4.y, gv, yb[4 * q4 + 1]);
                                    yb[4 * q4 + 2] = fmaf(w4.z, gv, yb[4 * q4 + 2]);
                                    yb[4 * q4 + 3] = fmaf(w4.w, gv, yb[4 * q4 + 3]);
                                }
                                if (emit && on) atomicAdd(gA + (size_t)u * a.B, gv);   // RED.ADD.F32: sum over the S samples of the row
                            }
                        }
                        tc_wait_ld(nxt);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < DP; ++jj)
                    if (jj < D) {
                        if (seq) vecp[oG + perm[jj]] = zb[jj];
                        else vecp[oG + permy[jj]] = yb[jj];
                    }
            }
        }
        if (MODE == TC_RSAMPLE && live) {
            for (int d = 0; d < D; ++d) a.theta_out[(size_t)pidx * D + d] = vecp[oV + K * D + d];
            if (a.diff) a.diff[pidx] = logp;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}

// ------------------------------------------------------------------------------------------------- image packing
// float part of every block, from the masked order-space image (maf_image.h): grid (blocks, 2 orientations, K)
__global__ void k_tc_pack(const MafImg g, TcBlk b0, TcBlk b1, int blk_words, int DP, int R0, int R1, float* __restrict__ img) {
    const int k = blockIdx.z, ori = blockIdx.y;
    const TcBlk bk = ori == 0 ? b0 : b1;
    const TImg t = g.t[k];
    const int H0 = g.H[0], H1 = g.H[1], ld1 = t.ld[1];
    float* out = img + (size_t)(2 * k + ori) * blk_words;
    const float* f = g.f;
    const int nW = (bk.NPK / 4) * bk.NPN * 4;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int o = t0; o < nW; o += stride) {
        const int kc = o / (bk.NPN * 4), n = (o >> 2) % bk.NPN, r = o & 3, kidx = 4 * kc + r;
        // orientation 0: n = layer-1 unit v, k = layer-0 unit u;  orientation 1: n = u, k = v;  W1 is [v][u]
        const int v = ori == 0 ? n : kidx, u = ori == 0 ? kidx : n;
        const float w = (v < H1 && u < H0) ? f[t.W[1] + (size_t)v * ld1 + u] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
        out[bk.whi + o] = hi;
        out[bk.wlo + o] = w - hi;
    }
    for (int o = t0; o < R0 * DP; o += stride) {
        const int u = o / DP, jj = o % DP;
        out[bk.w0t + o] = (u < H0 && jj < t.ldt) ? f[t.W0t + (size_t)u * t.ldt + jj] : 0.f;
    }
    for (int o = t0; o < R1; o += stride) out[bk.b1 + o] = o < H1 ? f[t.b[1] + o] : 0.f;
    for (int o = t0; o < R1 * 2 * DP; o += stride) {
        const int v = o / (2 * DP), c = o % (2 * DP);
        out[bk.wot + o] = (v < H1 && c < 2 * t.ldt) ? f[t.WoT + (size_t)v * 2 * t.ldt + c] : 0.f;
    }
    for (int o = t0; o < 2 * DP; o += stride) out[bk.bo + o] = o < 2 * t.ldt ? f[t.boF + o] : 0.f;
}

struct TcState {
    bool built = false, eligible = false;
    int DP = 4, R0 = 0, R1 = 0;
    TcBlk blk[2];
    int blk_words = 0;
    float* img = nullptr;
    unsigned long long packed_version = ~0ull;
    uint32_t* bits = nullptr;
    size_t bits_words = 0;
};

static inline int r8(int x) { return (x + 7) & ~7; }

// chunk lists of one orientation.  in_grp / out_grp: [D+1] group boundaries of the staged (input) layer and of the
// accumulator (output) layer.  forward: inputs of degree j feed outputs of degree >= j; reverse: cotangents of degree j
// feed the units of degree <= j of the layer below.
static void tc_chunks(int D, const int* in_grp, int Hin, const int* out_grp, int NPN, bool reverse,
                      std::vector<TcChunk>& seq, std::vector<int>& seqstart, std::vector<TcChunk>& full) {
    seq.clear();
    full.clear();
    seqstart.assign(D + 1, 0);
    for (int j = 0; j < D; ++j) {
        seqstart[j] = (int)seq.size();
        const int ua = in_grp[j], ub = in_grp[j + 1];
        if (ub > ua) {
            const int a8 = ua & ~7, b8 = r8(ub);
            for (int u0 = a8; u0 < b8; u0 += TC_KC) {
                TcChunk c;
                c.u0 = u0;
                c.nu = std::min(TC_KC, b8 - u0);
                c.ua = std::max(ua, u0);
                c.ub = std::min(ub, u0 + c.nu);
                if (!reverse) {
                    c.n0 = out_grp[j] & ~7;
                    c.N = NPN - c.n0;
                } else {
                    c.n0 = 0;
                    c.N = std::min(NPN, r8(out_grp[j + 1]));
                }
                c.pad0 = c.pad1 = 0;
                seq.push_back(c);
            }
        }
    }
    seqstart[D] = (int)seq.size();
    for (int u0 = 0; u0 < r8(Hin); u0 += TC_KC) {
        TcChunk c;
        c.u0 = u0;
        c.nu = std::min(TC_KC, r8(Hin) - u0);
        c.ua = u0;
        c.ub = std::min(Hin, u0 + c.nu);
        c.n0 = 0;
        c.N = NPN;
        c.pad0 = c.pad1 = 0;
        full.push_back(c);
    }
}

static int tc_build(rabi_maf* h, TcState* s) {
    s->built = true;
    s->eligible = false;
    const int K = h->K, D = h->D, H0 = h->H[0], H1 = h->H[1];
    if (h->L != 2 || D > 12 || H0 > TC_ACC || H1 > TC_ACC) return 0;
    const std::vector<int>& M = h->meta_host;
    const MafImg& g = h->img;
    for (int k = 0; k < K; ++k)
        for (int l = 0; l < 2; ++l)
            for (int j = 0; j < D; ++j)
                if (M[g.t[k].grp[l] + j + 1] <= M[g.t[k].grp[l] + j]) return 0;   // an empty degree group: not a Pyro mask shape we tile
    const int DP = (D + 3) & ~3;
    s->DP = DP;
    s->R0 = r8(H0) + 8;
    s->R1 = r8(H1) + 8;
    // block layout (words); the two orientations share everything but the matrix shape
    for (int ori = 0; ori < 2; ++ori) {
        TcBlk& b = s->blk[ori];
        b.NPN = ori == 0 ? r8(H1) : r8(H0);
        b.NPK = ori == 0 ? r8(H0) : r8(H1);
        int o = 0;
        auto take = [&](int n) {
            int at = o;
            o += (n + 3) & ~3;
            return at;
        };
        const int nW = (b.NPK / 4) * b.NPN * 4;
        b.whi = take(nW);
        b.wlo = take(nW);
        b.w0t = take(s->R0 * DP);
        b.b1 = take(s->R1);
        b.wot = take(s->R1 * 2 * DP);
        b.bo = take(2 * DP);
        b.perm = take(DP);
        b.permy = take(DP);
        b.grp0 = take(D + 1);
        b.grp1 = take(D + 1);
        b.nfull = take(4);
        b.seqstart = take(D + 1);
        b.cfull = take(8 * 8);
        b.cseq = take(TC_MAXCHUNK * 8);
        b.words = o;
    }
    s->blk_words = std::max(s->blk[0].words, s->blk[1].words);
    // integer tables per (k, orientation)
    std::vector<float> host((size_t)K * 2 * s->blk_words, 0.f);
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        for (int ori = 0; ori < 2; ++ori) {
            const TcBlk& b = s->blk[ori];
            int* w = reinterpret_cast<int*>(host.data() + (size_t)(2 * k + ori) * s->blk_words);
            for (int j = 0; j < D; ++j) {
                w[b.perm + j] = M[t.perm + j];
                w[b.permy + j] = M[t.permy + j];
            }
            for (int j = 0; j <= D; ++j) {
                w[b.grp0 + j] = M[t.grp[0] + j];
                w[b.grp1 + j] = M[t.grp[1] + j];
            }
            std::vector<TcChunk> seq, full;
            std::vector<int> ss;
            if (ori == 0) tc_chunks(D, &M[t.grp[0]], H0, &M[t.grp[1]], b.NPN, false, seq, ss, full);
            else tc_chunks(D, &M[t.grp[1]], H1, &M[t.grp[0]], b.NPN, true, seq, ss, full);
            if ((int)seq.size() > TC_MAXCHUNK || (int)full.size() > 8) return 0;
            w[b.nfull] = (int)full.size();
            for (int j = 0; j <= D; ++j) w[b.seqstart + j] = ss[j];
            memcpy(w + b.cfull, full.data(), full.size() * sizeof(TcChunk));
            memcpy(w + b.cseq, seq.data(), seq.size() * sizeof(TcChunk));
        }
    }
    if (s->img) cudaFree(s->img);
    s->img = nullptr;
    RABI_CU_OK(h, cudaMalloc(&s->img, host.size() * sizeof(float)));
    RABI_CU_OK(h, cudaMemcpy(s->img, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    s->packed_version = ~0ull;
    s->eligible = true;
    return 0;
}

template <int MODE, int DP>
static int tc_launch_inst(rabi_maf* h, const TcArgs& a, int grid, int T, size_t smem, cudaStream_t st) {
    RABI_CU_OK(h, cudaFuncSetAttribute(k_tc<MODE, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = MODE == TC_GRAD || MODE == TC_FKL;
    if (prof && rabi_prof_begin(h, st)) return 1;
    k_tc<MODE, DP><<<grid, T, smem, st>>>(h->img, a);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "k_tc launch failed: %s", cudaGetErrorString(e));
    if (prof && rabi_prof_end(h, st)) return 1;
    return 0;
}

template <int DP>
static int tc_launch_mode(rabi_maf* h, const TcArgs& a, int mode, int grid, int T, size_t smem, cudaStream_t st) {
    switch (mode) {
        case TC_FWD: return tc_launch_inst<TC_FWD, DP>(h, a, grid, T, smem, st);
        case TC_GRAD: return tc_launch_inst<TC_GRAD, DP>(h, a, grid, T, smem, st);
        case TC_FKL: return tc_launch_inst<TC_FKL, DP>(h, a, grid, T, smem, st);
        case TC_RSAMPLE: return tc_launch_inst<TC_RSAMPLE, DP>(h, a, grid, T, smem, st);
        case TC_LOGPROB: return tc_launch_inst<TC_LOGPROB, DP>(h, a, grid, T, smem, st);
    }
    return 2;
}

int rabi_tc_launch(rabi_maf* h, const TcCall& c, int mode, cudaStream_t st) {
    const long long npairs = (long long)c.S * c.B;
    if (npairs <= 0) return 0;
    if (!rabi_env_int("RABI_TC", 1)) return 2;
    TcState* s = static_cast<TcState*>(h->tc);
    if (!s) h->tc = s = new TcState();
    if (!s->built && tc_build(h, s)) return 1;
    if (!s->eligible) return 2;
    if (npairs < (long long)rabi_env_int("RABI_TC_MIN_PAIRS", 1)) return 2;
    if (s->packed_version != h->weights_version) {
        dim3 grid(16, 2, h->K);
        k_tc_pack<<<grid, 256, 0, st>>>(h->img, s->blk[0], s->blk[1], s->blk_words, s->DP, s->R0, s->R1, s->img);
        h->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return rabi_fail(h, "k_tc_pack launch failed: %s", cudaGetErrorString(e));
        s->packed_version = h->weights_version;
    }
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.A_p = c.A_p;
    a.A_q = c.A_q;
    a.z = c.z;
    a.diff = c.diff;
    a.theta_out = c.theta_out;
    a.gA = c.gA;
    a.S = c.S;
    a.B = c.B;
    a.w = c.w;
    a.npairs = npairs;
    const long long nt = (npairs + 127) / 128;
    if (nt > 0x7fffffff) return 2;
    a.ntiles = (int)nt;
    a.img = s->img;
    a.blk[0] = s->blk[0];
    a.blk[1] = s->blk[1];
    a.blk_words = s->blk_words;
    a.VS = ((4 * h->K + 2) * h->D) | 1;
    // tiles per CTA: as many as the per-pair state next to the weight block allows, no more than the launch needs
    const size_t budget = (size_t)h->max_smem_optin - 1024;
    int NT = std::min(TC_MAXTILES, std::max(1, rabi_env_int("RABI_TC_TILES", TC_MAXTILES)));
    while (NT > 1 && ((size_t)s->blk_words + (size_t)NT * 128 * a.VS) * 4 > budget) --NT;
    if (((size_t)s->blk_words + (size_t)NT * 128 * a.VS) * 4 > budget) return 2;
    const int grid = (int)std::min<long long>(nt, h->sm_count);
    while (NT > 1 && (long long)grid * (NT - 1) >= nt) --NT;   // spread the tiles over the SMs first
    const int T = NT * 128;
    const size_t smem = ((size_t)s->blk_words + (size_t)T * a.VS) * 4;
    if (mode == TC_GRAD || mode == TC_FKL) {
        const size_t need = (size_t)2 * h->K * 2 * TC_WL * (size_t)h->sm_count * TC_MAXTILES * 128;
        if (need > s->bits_words) {
            if (s->bits) {
                RABI_CU_OK(h, cudaDeviceSynchronize());
                RABI_CU_OK(h, cudaFree(s->bits));
                s->bits = nullptr;
                s->bits_words = 0;
            }
            RABI_CU_OK(h, cudaMalloc(&s->bits, need * sizeof(uint32_t)));
            s->bits_words = need;
        }
        a.bits = s->bits;
        a.bits_stride = h->sm_count * TC_MAXTILES * 128;
    }
    switch (s->DP) {
        case 4: return tc_launch_mode<4>(h, a, mode, grid, T, smem, st);
#ifndef RABI_LEAN
        case 8: return tc_launch_mode<8>(h, a, mode, grid, T, smem, st);
        case 12: return tc_launch_mode<12>(h, a, mode, grid, T, smem, st);
#endif
    }
    return 2;
}

void rabi_tc_destroy(rabi_maf* h) {
    TcState* s = static_cast<TcState*>(h->tc);
    if (!s) return;
    if (s->img) cudaFree(s->img);
    if (s->bits) cudaFree(s->bits);
    delete s;
    h->tc = nullptr;
}

}  // namespace rabi
