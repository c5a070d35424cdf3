// Tensor-core path for two-hidden-layer MADE conditioners (config/model/maf_pyro.yaml: hidden_dims [100, 100]).
//
// Same arithmetic as k_fast (maf_fast.cuh; reference: Pyro ConditionalAutoRegressiveNN / AffineAutoregressive as driven
// by rbi/rbi/models/transforms.py:37-46, general_kl_divergence rbi/rbi/loss/kl_div.py:38-56), but the layer-0 ->
// layer-1 contraction (91 % of the masked MACs at lotka_volterra shapes) runs on the 5th-generation tensor cores:
//
//  * a CTA holds up to 3 TILES of 128 (observation row, MC sample) pairs, one warpgroup each; thread = pair = TMEM lane;
//  * per tile 168 tensor-memory columns: [0, 120) fp32 accumulators of the layer being produced (column = unit, the MADE
//    degree groups each padded to a multiple of 4), [120, 144) / [144, 168) the TF32 (hi, lo) halves of up to 24 staged
//    inputs;
//  * "group push": the threads compute a group of layer-0 activations (ReLU(A[u] + W0t[u].y), FMA code, D is tiny),
//    split them into (hi, lo), store them to the staging columns (tcgen05.st); ONE elected thread issues the 3xTF32
//    chain  acc[:, n0:] += h_hi*W_hi + h_lo*W_hi + h_hi*W_lo  (tcgen05.mma.cta_group::1.kind::tf32, M = 128 pairs,
//    A from tensor memory, B = weight strip in shared memory, K-major no-swizzle [k/4][n][4] layout) and commits to the
//    tile's mbarrier; the threads wait and read the accumulator columns that are final now (tcgen05.ld);
//  * the sequential sampling pass pushes, at autoregressive step j, only the layer-0 units of MADE degree j into the
//    accumulator columns of the layer-1 units of degree >= j (one pass-equivalent, exactly the masked MACs, N shrinks
//    with j); the density pass runs the same pushes without the intermediate read-backs; the reverse passes run the
//    same pattern on the transposed weight copy (cotangents of layer 1 pushed into the columns of layer 0);
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

#ifndef RABI_TC_FASTPOLL
#define RABI_TC_FASTPOLL 64
#endif
constexpr int TC_KC = 24;          // staged inputs per push (columns of the hi half and of the lo half)
constexpr int TC_ACC = 120;        // accumulator columns = max padded units per hidden layer on this path
constexpr int TC_SH = TC_ACC;      // staging columns, hi half
constexpr int TC_SL = TC_ACC + TC_KC;
constexpr int TC_TILE_COLS = TC_ACC + 2 * TC_KC;  // 168
constexpr int TC_MAXTILES = 3;
constexpr int TC_MAXP = 24;        // pieces per layer

// Unit layout inside the tile ("padded"): the MADE degree groups of a layer are laid out one after the other, each
// padded to a multiple of 4 units, so that 128-bit tensor-memory accesses never straddle two groups.  Padding units have
// zero weights everywhere (rows and columns), which makes whatever is staged for them harmless.
// A PIECE is a run of <= TC_KC padded units of one degree group: the unit of staging + push (as the K side of a
// product) and of read-back (as the N side of the product that feeds the layer).
struct TcPiece {
    int u0;      // first padded unit (multiple of 4)
    int nu;      // staged columns (multiple of 8, <= TC_KC); columns [nz, nu) are staged as zeros
    int nz;      // padded units of the piece (multiple of 4)
    int r0, nr;  // real units [r0, r0 + nr) <-> padded [u0, u0 + nr)
    int n0, N;   // accumulator columns [n0, n0 + N) that receive the piece's product (multiples of 8)
    int flags;   // degree (bits 0..7) | last piece of its degree group (bit 8) | first piece of its group (bit 9)
};

// word offsets inside one (transform, orientation) block; identical for all transforms
struct TcBlk {
    int whi, wlo;        // [NPK/4][NPN][4] TF32 hi / lo halves of layer 1 (orientation 0: n = layer-1 unit, k = layer-0 unit)
    int NPN, NPK;
    int w0t, b1, wot, bo;  // [R0][DP], [R1], [R1][2*DP], [2*DP]   (padded unit order)
    int perm, permy;
    int np, p0start, p1start, P0, P1;   // np: (np0, np1); p?start[D+1]; piece tables
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
    int region_words;    // shared-memory words in front of the per-pair region (>= the largest block; RUN: also the context weights)
    uint32_t* bits;      // [2K][2][TC_MAXP][bits_stride] relu masks, one word per piece
    int bits_stride;
    int VS;
    // ---- persistent attack (template RUN): a CTA owns blocks of R observation rows with all their S samples for all
    // nb_iter L2-PGD iterations; projection, gradient back-projection and the PGD update run inside the kernel
    const float* x;       // [B][C]
    float* delta;         // [B][C] in: delta_0, out: delta_nb_iter
    const float* zmean;   // [C] or null
    const float* zstd;
    int nb_iter, R, RP, nblocks;   // R rows per block (multiple of 4), RP = shared-memory row stride (round8(R) + 4)
    float eps, eps_iter, cmin, cmax, floor_;
    const float* eps_rows;      // optional per-row radius / step size
    const float* eps_iter_rows;
    const float* col_mask;      // optional [C] attack mask
    // tensor-core tail (tct): the projection and the back-projection of a block as 3xTF32 MMAs with A AND B in shared memory
    int tct;              // 0: the packed-FMA products
    const float* pimg;    // [2 orientations][K][strip] context weights as (hi | lo) strips in the K-major core-matrix layout
    int strip_words;      // floats per strip (both orientations padded to the same size)
    int nch0, NB0;        // projection:      K dimension = C  in nch0 chunks of 4, N = NB0 (H0 rounded to 8)
    int nch1, NB1;        // back-projection: K dimension = H0 in nch1 chunks of 4, N = NB1 (C rounded to 8)
    long long zstride;    // floats between the base noise of consecutive iterations (S * B * D)
    const float* wrow;    // TC_NLL: per-pair loss weights and the per-pair vectors for the weight-gradient products (maf_tc.h)
    float* em_h0;
    float* em_a0;
    float* em_h1;
    float* em_a1;
    float* em_o;
    float* em_y;
    float* theta_grad;
    int vec_words;        // words of the per-pair region (the tail's tiles live there between the flow phases)
    int xs_off;           // word offset of Xs inside it; the partial gradient tiles of the thread groups lie in front
    long long* dbg;       // optional [16] cycle counters of CTA 0 (RABI_TC_RUN_DBG=1): where an iteration's time goes
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
// A and B from shared memory (both K-major, no swizzle)
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// generic-proxy writes to shared memory (st.shared) become visible to the async proxy (the MMA's operand reads)
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
#ifndef RABI_TC_NO_FASTPOLL
    // the waits of this kernel are short (one MMA chain): poll without suspending first -- try_wait parks the thread for a
    // hardware-chosen interval, which is long compared to a push round trip
#pragma unroll 1
    for (int i = 0; i < RABI_TC_FASTPOLL; ++i) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(tc_smem_u32(mbar)), "r"(parity) : "memory");
        if (done) return;
    }
#endif
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
// stage one quad of values as TF32 (hi, lo) halves
__device__ __forceinline__ void tc_stage4(uint32_t trow, int q, const float (&v)[4]) {
    const float h0 = tc_hi(v[0]), h1 = tc_hi(v[1]), h2 = tc_hi(v[2]), h3 = tc_hi(v[3]);
    tc_st4(trow + TC_SH + q, h0, h1, h2, h3);
    tc_st4(trow + TC_SL + q, v[0] - h0, v[1] - h1, v[2] - h2, v[3] - h3);
}

// ------------------------------------------------------------------------------------------------- the kernel
template <int MODE, int DP, bool RUN>
__global__ void __launch_bounds__(TC_MAXTILES * 128, 1) k_tc(const MafImg g, const TcArgs a) {
    extern __shared__ __align__(128) float sm[];
    constexpr bool GRADM = MODE == TC_GRAD || MODE == TC_FKL || MODE == TC_NLL || MODE == TC_RSB;
    constexpr bool NLLM = MODE == TC_NLL;
    constexpr bool RSBM = MODE == TC_RSB;
    constexpr bool EMITM = NLLM || RSBM;   // training side: per-pair vectors for the weight-gradient products are written out
    static_assert(!RUN || (MODE == TC_GRAD || MODE == TC_FKL), "the persistent attack exists for the two attack losses only");
    float* Wb = sm;                                   // the resident (transform, orientation) block
    const int* Wi = reinterpret_cast<const int*>(sm);
    float* vec = sm + a.region_words;                 // [threads][VS]
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar[TC_MAXTILES + 2];   // one per tile (MMA completion) + the block loader + the tail's MMAs
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2, lane128 = tid & 127;
    const int NT = blockDim.x >> 7;
    const int K = g.K, D = g.D, H0 = g.H[0];
    const float HALF_LOG_2PI = 0.91893853320467274178f;

    if (tid == 0) {
        for (int i = 0; i <= TC_MAXTILES + 1; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(&mbar[i])) : "memory");
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
    uint32_t* bitp = a.bits + (blockIdx.x * blockDim.x + tid);
    uint64_t* my_mbar = &mbar[wg];
    uint32_t mphase = 0, wphase = 0, tphase = 0;

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

    // one push: the staged columns of piece c (this tile) times the weight strip, into accumulator columns [n0, n0 + N)
    auto push = [&](const TcBlk& bk, const TcPiece& c, bool first) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
        if (lane128 == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = tc_make_idesc(128, c.N);
            const uint32_t lbo = (uint32_t)bk.NPN * 16u;
            const uint32_t strip = (uint32_t)(((c.u0 >> 2) * bk.NPN + c.n0) * 16);
            const uint64_t dhi = tc_make_desc(tc_smem_u32(Wb) + 4u * (uint32_t)bk.whi + strip, lbo, 128u);
            const uint64_t dlo = tc_make_desc(tc_smem_u32(Wb) + 4u * (uint32_t)bk.wlo + strip, lbo, 128u);
            const uint64_t dstep = (uint64_t)((2u * lbo) >> 4);   // one K = 8 step = two k chunks
            const uint32_t dcol = tcol + (uint32_t)c.n0;
            uint32_t acc = first ? 0u : 1u;
            const int nk = c.nu >> 3;
            for (int kk = 0; kk < nk; ++kk) {   // (h_hi, W_hi)
                tc_mma_ts(dcol, tcol + TC_SH + 8 * kk, dhi + kk * dstep, idesc, acc);
                acc = 1u;
            }
            for (int kk = 0; kk < nk; ++kk) tc_mma_ts(dcol, tcol + TC_SL + 8 * kk, dhi + kk * dstep, idesc, 1u);  // (h_lo, W_hi)
            for (int kk = 0; kk < nk; ++kk) tc_mma_ts(dcol, tcol + TC_SH + 8 * kk, dlo + kk * dstep, idesc, 1u);  // (h_hi, W_lo)
            tc_commit(my_mbar);
        }
        tc_mbar_wait(my_mbar, mphase);
        mphase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    // generic smem <- global bulk copies into the block region (same mbarrier protocol as ensure_block)
    auto bulk_expect = [&](uint32_t bytes) {
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tc_smem_u32(&mbar[TC_MAXTILES])), "r"(bytes) : "memory");
    };
    auto bulk_piece = [&](uint32_t dst_off, const void* srcp, uint32_t bytes) {
        if (tid == 0) {
            const char* src = reinterpret_cast<const char*>(srcp);
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t n = min(bytes - off, 32768u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(tc_smem_u32(sm) + dst_off + off), "l"(src + off), "r"(n), "r"(tc_smem_u32(&mbar[TC_MAXTILES])) : "memory");
            }
        }
    };
    auto bulk_wait = [&]() {
        tc_mbar_wait(&mbar[TC_MAXTILES], wphase);
        wphase ^= 1u;
    };
    // RUN: geometry of the in-kernel projection products (8 rows x 8 outputs per thread, packed FMAs, operands in shared memory)
    const int C = g.C;
    const int RP = RUN ? a.RP : 0;                   // row stride of the staged tiles: RT + 4, so that column walks spread over 8 banks
    const int RT = RP - 4;                           // rows covered by the 8-row tiles (rows >= nr are zero)
    float* Xs = vec + a.xs_off;                      // [C][RP]    z-scored x + delta of the block's rows, transposed (behind the gradient tiles)
    float* Xn = Xs;
    float* Gs = vec;                                 // [K H0][RP] cotangents d/dA of the block's rows; then up to G partial
                                                     //            gradient tiles [G][C][RP] of the thread groups
    float* pertA = const_cast<float*>(MODE == TC_GRAD ? a.A_p : a.A_q);   // the projection that follows delta
    const int nwarps = (int)(blockDim.x >> 5), lane = tid & 31;
    auto load_ctx_weights = [&]() {                  // W0cT [C][ldKH] of all transforms -> weight region
        const uint32_t wbytes = (uint32_t)C * (uint32_t)g.ldKH * 4u;
        bulk_expect(wbytes);
        bulk_piece(0u, g.f + g.W0cT, wbytes);
    };
    // ---- tensor-core tail (a.tct): both products as 3xTF32 MMAs with A and B in shared memory, D in the (idle) tensor memory
    const int NCA = RUN ? max(a.nch0, a.nch1) : 0;   // chunks of 4 K-indices in the A-operand buffers
    const int ACS = RT * 4 + 4;                      // floats per chunk of the A operand: RT rows x 4 + one quad of padding, so that
                                                     // walks along the K index hit 32 different banks
    float* Ahi = vec;                                // [NCA][ACS] hi halves of the A operand: z-scored x + delta rows | gA chunk
    float* Alo = vec + (size_t)NCA * ACS;            //            lo halves
    float* Gt2 = vec + (size_t)2 * NCA * ACS;        // [RT][CG]   input gradient of the block's rows
    const uint32_t trow0 = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16);   // this warp's lane quadrant, column 0
    auto put_a = [&](int kidx, int r, float v) {     // element (row r, K-index kidx) of the A operand as TF32 (hi, lo)
        const int off = (kidx >> 2) * ACS + r * 4 + (kidx & 3);
        const float hi = tc_hi(v);
        Ahi[off] = hi;
        Alo[off] = v - hi;
    };
    auto load_strip = [&](int ori, int kk) {         // (hi | lo) weight strip of transform kk -> weight region
        const uint32_t bytes = (uint32_t)a.strip_words * 4u;
        bulk_expect(bytes);
        bulk_piece(0u, a.pimg + ((size_t)ori * K + kk) * a.strip_words, bytes);
    };
    // D[128 x NB] (+)= A[128 x 4 nch] B^T as a 3xTF32 chain; callers made the A buffers visible (fence + barrier) and hold the strip
    auto tail_mma = [&](uint32_t dcol, int nch, int NB, bool first) {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = tc_make_idesc(128, NB);
            const uint32_t lboA = (uint32_t)ACS * 4u, lboB = (uint32_t)NB * 16u;
            const uint64_t ahi = tc_make_desc(tc_smem_u32(Ahi), lboA, 128u), alo = tc_make_desc(tc_smem_u32(Alo), lboA, 128u);
            const uint64_t bhi = tc_make_desc(tc_smem_u32(Wb), lboB, 128u);
            const uint64_t blo = tc_make_desc(tc_smem_u32(Wb) + (uint32_t)nch * lboB, lboB, 128u);
            const uint64_t sa = (uint64_t)((2u * lboA) >> 4), sb = (uint64_t)((2u * lboB) >> 4);   // one K = 8 step = two chunks
            const int nks = nch >> 1;
            uint32_t acc = first ? 0u : 1u;
            for (int ks = 0; ks < nks; ++ks) {
                tc_mma_ss(tmem_base_s + dcol, ahi + ks * sa, bhi + ks * sb, idesc, acc);
                acc = 1u;
            }
            for (int ks = 0; ks < nks; ++ks) tc_mma_ss(tmem_base_s + dcol, alo + ks * sa, bhi + ks * sb, idesc, 1u);
            for (int ks = 0; ks < nks; ++ks) tc_mma_ss(tmem_base_s + dcol, ahi + ks * sa, blo + ks * sb, idesc, 1u);
            tc_commit(&mbar[TC_MAXTILES + 1]);
        }
        tc_mbar_wait(&mbar[TC_MAXTILES + 1], tphase);
        tphase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    long long tmark = 0;
    auto mark = [&](int slot) {
        if (RUN && a.dbg && blockIdx.x == 0 && tid == 0) {
            const long long now = clock64();
            if (slot >= 0) a.dbg[slot] += now - tmark;
            tmark = now;
        }
    };

    const int rounds = RUN ? a.nblocks : (a.ntiles + (int)gridDim.x * NT - 1) / ((int)gridDim.x * NT);
    for (int round = RUN ? (int)blockIdx.x : 0; round < rounds; round += RUN ? (int)gridDim.x : 1) {
        // ---- which pair this thread owns
        bool tile_live, live;
        long long pidx;
        int b, b0 = 0, nr = 0;
        if constexpr (RUN) {
            b0 = round * a.R;
            nr = min(a.R, a.B - b0);
            const int lp = wg * 128 + lane128;       // local pair: sample-major over the block's rows
            const int sidx = lp / nr, rr = lp - sidx * nr;
            tile_live = wg * 128 < a.S * nr;
            live = sidx < a.S;
            b = b0 + (live ? rr : 0);
            pidx = (long long)(live ? sidx : 0) * a.B + b;
        } else {
            const int tile = (round * NT + wg) * (int)gridDim.x + (int)blockIdx.x;
            tile_live = tile < a.ntiles;
            const long long p = (long long)tile * 128 + lane128;
            live = tile_live && p < a.npairs;
            pidx = live ? p : 0;
            b = (int)(pidx % a.B);
        }
        if constexpr (RUN) {   // Xs <- ((x + delta_0) - mean) / std, rows of the block (lanes walk the context dimension)
            __syncthreads();
            resident = -1;
            if (a.tct) load_strip(0, 0);
            else load_ctx_weights();
            for (int r = warp; r < RT; r += nwarps) {
                const bool in = r < nr;
                for (int c = lane; c < 4 * NCA; c += 32) {   // tct: the K-index padding [C, 4 NCA) is zero-filled as well
                    float v = 0.f;
                    if (in && c < C) {
                        v = a.x[(size_t)(b0 + r) * C + c] + a.delta[(size_t)(b0 + r) * C + c];
                        if (a.zmean) v = (v - a.zmean[c]) / a.zstd[c];
                    }
                    if (a.tct) put_a(c, r, v);
                    else if (c < C) Xs[c * RP + r] = v;
                }
                if (!a.tct)
                    for (int c = 4 * NCA + lane; c < C; c += 32) {   // (NCA may be 0 without the tensor-core tail)
                        float v = 0.f;
                        if (in) {
                            v = a.x[(size_t)(b0 + r) * C + c] + a.delta[(size_t)(b0 + r) * C + c];
                            if (a.zmean) v = (v - a.zmean[c]) / a.zstd[c];
                        }
                        Xs[c * RP + r] = v;
                    }
            }
            if (a.tct) tc_fence_async_smem();
        }
      for (int it = 0; it < (RUN ? a.nb_iter : 1); ++it) {
        if constexpr (RUN) {
            // ================= projection of the block's rows: A[k,u,b] = b0_k[u] + sum_c W0_k[u,c] ctx[b,c] =================
            mark(-1);
            if (a.tct) {
                // ---- tensor cores: per transform D[:, kk NB0 ..] = X~ W0c_kk^T (A = the (hi, lo) rows in shared memory, B = the strip)
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();                      // the A operand is complete and fenced for the async proxy
                for (int kk = 0; kk < K; ++kk) {
                    if (kk) load_strip(0, kk);
                    bulk_wait();
                    tail_mma((uint32_t)(kk * a.NB0), a.nch0, a.NB0, true);
                }
                mark(0);
                // epilogue: D (lane = row) -> + bias -> transposed through the (now dead) weight region [K H0][RT] -> 128-bit stores of
                // four consecutive rows into the CTA's slab A[k,u,b0 + .]
                const int row = 32 * (warp & 3) + lane, part = warp >> 2, nparts = nwarps >> 2, H0q = (H0 + 3) >> 2;
                // (item i <-> (transform kk, unit quad uq) is advanced incrementally: no integer divisions in the loop)
                int kk_i = part / H0q, uq_i = part - kk_i * H0q;
                for (int i0 = part; i0 < K * H0q; i0 += 5 * nparts) {
                    float4 v[5];
                    int kks[5], uqs[5];
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        kks[j] = kk_i;
                        uqs[j] = uq_i;
                        v[j] = kk_i < K ? tc_ld4(trow0 + (uint32_t)(kk_i * a.NB0 + 4 * uq_i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        uq_i += nparts;
                        while (uq_i >= H0q) {
                            uq_i -= H0q;
                            ++kk_i;
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[0].z), "+f"(v[0].w), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[1].z), "+f"(v[1].w),
                                   "+f"(v[2].x), "+f"(v[2].y), "+f"(v[2].z), "+f"(v[2].w), "+f"(v[3].x), "+f"(v[3].y), "+f"(v[3].z), "+f"(v[3].w),
                                   "+f"(v[4].x), "+f"(v[4].y), "+f"(v[4].z), "+f"(v[4].w) :: "memory");
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        if (kks[j] < K && row < RT) {
                            const int u = 4 * uqs[j];
                            float* dst = Wb + (kks[j] * H0 + u) * RT + row;
                            const float vals[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (u + e < H0) dst[e * RT] = vals[e];   // (the bias joins in the store loop)
                        }
                    }
                }
                __syncthreads();
                {
                    const int RQ = RT >> 2, n = K * H0 * RQ;
                    const bool vec4 = (a.B & 3) == 0;   // b0 is a multiple of 4
                    for (int i = tid; i < n; i += (int)blockDim.x) {
                        const int j = i / RQ, r0 = 4 * (i - j * RQ), kk = j / H0;
                        const float bias = __ldg(g.f + g.t[kk].b0 + (j - kk * H0));
                        float4 v = *reinterpret_cast<const float4*>(Wb + j * RT + r0);
                        v = make_float4(v.x + bias, v.y + bias, v.z + bias, v.w + bias);
                        float* out = pertA + (size_t)j * a.B + b0 + r0;
                        if (vec4 && r0 + 3 < nr) {
                            __stcg(reinterpret_cast<float4*>(out), v);
                        } else {
                            const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (r0 + e < nr) __stcg(out + e, vals[e]);
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();                      // the projection is visible to the block; tensor memory free for the tiles
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mark(1);
            } else {
            __syncthreads();                          // Xs complete
            bulk_wait();
            mark(0);                              // context weights (issued before the update / the first Xs)
            const int ldw = g.ldKH, ncol = K * g.H0p, NG = (ncol + 7) >> 3, RG = RT >> 3;
            for (int t = tid; t < NG * RG; t += (int)blockDim.x) {
                const int rg = t % RG, ug = t / RG;
                float2 acc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 2
                for (int c = 0; c < C; ++c) {
                    const float4 x0 = *reinterpret_cast<const float4*>(Xs + c * RP + 8 * rg);
                    const float4 x1 = *reinterpret_cast<const float4*>(Xs + c * RP + 8 * rg + 4);
                    const float4 w0 = *reinterpret_cast<const float4*>(Wb + c * ldw + 8 * ug);
                    const float4 w1 = *reinterpret_cast<const float4*>(Wb + c * ldw + 8 * ug + 4);
                    const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 xx = make_float2(xr[i], xr[i]);
                        acc[i][0] = tc_ffma2(make_float2(w0.x, w0.y), xx, acc[i][0]);
                        acc[i][1] = tc_ffma2(make_float2(w0.z, w0.w), xx, acc[i][1]);
                        acc[i][2] = tc_ffma2(make_float2(w1.x, w1.y), xx, acc[i][2]);
                        acc[i][3] = tc_ffma2(make_float2(w1.z, w1.w), xx, acc[i][3]);
                    }
                }
                const bool vec4 = (a.B & 3) == 0;     // b0 is a multiple of 4
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = 8 * ug + j;
                    if (col >= ncol) continue;
                    const int kk = col / g.H0p, u = col - kk * g.H0p;
                    if (u >= H0) continue;
                    const float bias = __ldg(g.f + g.t[kk].b0 + u);
                    float* out = pertA + ((size_t)kk * H0 + u) * a.B + b0 + 8 * rg;
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = bias + ((j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x);
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        const int r0 = 8 * rg + 4 * h4;
                        if (vec4 && r0 + 3 < nr) {
                            __stcg(reinterpret_cast<float4*>(out + 4 * h4), make_float4(o[4 * h4], o[4 * h4 + 1], o[4 * h4 + 2], o[4 * h4 + 3]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (r0 + i < nr) __stcg(out + 4 * h4 + i, o[4 * h4 + i]);
                        }
                    }
                }
            }
            __syncthreads();                          // the projection is visible to the block; Xs and the weights are dead
            mark(1);
            }
        }
        const float* zit = a.z + (RUN ? (size_t)it * (size_t)a.zstride : (size_t)0);
        float logp = 0.f, logq = 0.f;
        if (tile_live) {
            float ssq = 0.f;
            for (int d = 0; d < D; ++d) {
                const float zv = live ? zit[(size_t)pidx * D + d] : 0.f;
                vecp[oV + ((MODE == TC_LOGPROB || NLLM) ? K * D : 0) + d] = zv;
                ssq = fmaf(zv, zv, ssq);
            }
            logp = (MODE == TC_LOGPROB || NLLM) ? 0.f : -0.5f * ssq - (float)D * HALF_LOG_2PI;
            if (RSBM)
                for (int d = 0; d < D; ++d) vecp[oG + d] = live ? a.theta_grad[(size_t)pidx * D + d] : 0.f;
        }
        const int nphase = ((MODE == TC_GRAD || RSBM) ? 4 : (MODE == TC_FKL || NLLM) ? 3 : MODE == TC_RSAMPLE ? 1 : 2) * K;
        const float wp = a.wrow ? (live ? a.wrow[pidx] : 0.f) : a.w;   // weight of this pair in the loss
        const size_t NP = (size_t)a.npairs;
        for (int ph = ((MODE == TC_LOGPROB || NLLM) ? K : 0); ph < nphase; ++ph) {
            const int kind = ph / K;
            if (RSBM && (kind == 1 || kind == 2)) continue;   // sampling chain and its reverse only
            const int idx = ph - kind * K;
            const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
            const bool seq = (kind == 0 || kind == 3);
            const int ori = kind < 2 ? 0 : 1;
            mark(-1);
            ensure_block(k, ori);
            mark(8);
            if (!tile_live) {
                mark(9 + kind);
                continue;
            }
            const TcBlk& bk = a.blk[ori];
            const float* W0t = Wb + bk.w0t;
            const float* b1 = Wb + bk.b1;
            const float* WoT = Wb + bk.wot;
            const float* bo = Wb + bk.bo;
            const int* perm = Wi + bk.perm;
            const int* permy = Wi + bk.permy;
            const int np0 = Wi[bk.np], np1 = Wi[bk.np + 1];
            const int* p0start = Wi + bk.p0start;
            const int* p1start = Wi + bk.p1start;
            const TcPiece* P0 = reinterpret_cast<const TcPiece*>(Wi + bk.P0);
            const TcPiece* P1 = reinterpret_cast<const TcPiece*>(Wi + bk.P1);
            // relu masks of (pass, k): one word per piece, layer 0 then layer 1
            uint32_t* bits0 = bitp + (size_t)((seq ? k : K + k) * 2 * TC_MAXP) * a.bits_stride;
            uint32_t* bits1 = bits0 + (size_t)TC_MAXP * a.bits_stride;
            const int src = (k == K - 1) ? oV + K * D : oU + (k + 1) * D;

            if (kind < 2) {
                // ===================== forward pass of transform k =====================
                const float* Abase = (seq ? a.A_p : a.A_q) + (size_t)k * H0 * a.B + b;
                const bool rewritten = RUN && (MODE == TC_GRAD ? seq : !seq);   // this projection follows delta: no read-only path
                float zos[DP], y[DP];
                float2 oms[DP];
#pragma unroll
                for (int jj = 0; jj < DP; ++jj) {
                    zos[jj] = (seq && jj < D) ? vecp[oV + k * D + perm[jj]] : 0.f;
                    y[jj] = (!seq && jj < D) ? vecp[src + permy[jj]] : 0.f;
                    oms[jj] = jj < D ? make_float2(bo[2 * jj], bo[2 * jj + 1]) : make_float2(0.f, 0.f);
                }
                if (NLLM && live)
                    for (int d = 0; d < D; ++d) a.em_y[((size_t)k * D + d) * NP + pidx] = vecp[src + d];
                float an[TC_KC];   // context projection of the NEXT piece's units (loads in flight during the push)
                {
                    const TcPiece c0 = P0[0];
                    const float* ap = Abase + (size_t)c0.r0 * a.B;
#pragma unroll
                    for (int i = 0; i < TC_KC; ++i) {
                        an[i] = i < c0.nr ? (rewritten ? __ldcg(ap) : __ldg(ap)) : 0.f;
                        ap += a.B;
                    }
                }
                for (int pi = 0; pi < np0; ++pi) {
                    const TcPiece c = P0[pi];
                    float av[TC_KC];
#pragma unroll
                    for (int i = 0; i < TC_KC; ++i) av[i] = an[i];
                    if (pi + 1 < np0) {
                        const TcPiece cn = P0[pi + 1];
                        const float* ap = Abase + (size_t)cn.r0 * a.B;
#pragma unroll
                        for (int i = 0; i < TC_KC; ++i) {
                            an[i] = i < cn.nr ? (rewritten ? __ldcg(ap) : __ldg(ap)) : 0.f;
                            ap += a.B;
                        }
                    }
                    // ---- layer-0 units of the piece: h = relu(A[u] + W0t[u] . y), staged as (hi, lo)
                    uint32_t cm = 0u;
#pragma unroll
                    for (int q = 0; q < TC_KC; q += 4) {
                        if (q < c.nz) {
                            float hv[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float pre = av[q + e];
#pragma unroll
                                for (int q4 = 0; q4 < DP / 4; ++q4) {
                                    const float4 w = *reinterpret_cast<const float4*>(W0t + (size_t)(c.u0 + q + e) * DP + 4 * q4);
                                    pre = fmaf(w.x, y[4 * q4], pre);
                                    pre = fmaf(w.y, y[4 * q4 + 1], pre);
                                    pre = fmaf(w.z, y[4 * q4 + 2], pre);
                                    pre = fmaf(w.w, y[4 * q4 + 3], pre);
                                }
                                hv[e] = fmaxf(pre, 0.f);
                                cm |= pre > 0.f ? (1u << (q + e)) : 0u;
                                if (EMITM && live && q + e < c.nr) a.em_h0[((size_t)k * H0 + c.r0 + q + e) * NP + pidx] = hv[e];
                            }
                            tc_stage4(trow, q, hv);
                        } else if (q < c.nu) {   // columns of the next group inside the last K = 8 step: not known yet
                            tc_st4(trow + TC_SH + q, 0.f, 0.f, 0.f, 0.f);
                            tc_st4(trow + TC_SL + q, 0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    if (GRADM) bits0[(size_t)pi * a.bits_stride] = cm;
                    push(bk, c, pi == 0);
                    const bool group_done = (c.flags >> 8) & 1;
                    if (!(seq ? group_done : pi == np0 - 1)) continue;
                    // ---- layer-1 units that are final now: relu(acc + b1) folded into the (mean, log-scale) accumulators
                    const int j = c.flags & 255;
                    const int pj_lo = seq ? p1start[j] : 0, pj_hi = seq ? p1start[j + 1] : np1;
                    for (int pj = pj_lo; pj < pj_hi; ++pj) {
                        const TcPiece r = P1[pj];
                        uint32_t rm = 0u;
                        float4 nxt = tc_ld4(trow + (uint32_t)r.u0);
#pragma unroll
                        for (int q = 0; q < TC_KC; q += 4) {
                            if (q < r.nz) {
                                tc_wait_ld(nxt);
                                const float4 acc = nxt;
                                nxt = tc_ld4(trow + (uint32_t)(r.u0 + q + 4));   // look-ahead (stays inside the tile's columns)
                                const float4 bias = *reinterpret_cast<const float4*>(b1 + r.u0 + q);
                                const float val[4] = {acc.x + bias.x, acc.y + bias.y, acc.z + bias.z, acc.w + bias.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float hv = fmaxf(val[e], 0.f);
                                    rm |= val[e] > 0.f ? (1u << (q + e)) : 0u;
                                    if (EMITM && live && q + e < r.nr) a.em_h1[((size_t)k * g.H[1] + r.r0 + q + e) * NP + pidx] = hv;
                                    const float2 hh = make_float2(hv, hv);
                                    const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)(r.u0 + q + e) * 2 * DP);
#pragma unroll
                                    for (int q2 = 0; q2 < DP / 2; ++q2) {
                                        const float4 wq = wo[q2];   // (m_j, s_j, m_j+1, s_j+1)
                                        oms[2 * q2] = tc_ffma2(make_float2(wq.x, wq.y), hh, oms[2 * q2]);
                                        oms[2 * q2 + 1] = tc_ffma2(make_float2(wq.z, wq.w), hh, oms[2 * q2 + 1]);
                                    }
                                }
                            }
                        }
                        tc_wait_ld(nxt);
                        if (GRADM) bits1[(size_t)pj * a.bits_stride] = rm;
                    }
                    if (seq) {  // position j is final: y_j = (z_j - m_j) exp(-clamp(s_j))
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
                if (seq && RSBM && live)
                    for (int d = 0; d < D; ++d) a.em_y[((size_t)k * D + d) * NP + pidx] = vecp[oV + (k + 1) * D + d];
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
                            if (GRADM) vecp[oG + d] = wp * u;  // d(-w log q)/d u_0
                        }
                        logq += -0.5f * ssq - (float)D * HALF_LOG_2PI;
                        if (a.diff && live) a.diff[pidx] = (MODE == TC_LOGPROB || NLLM) ? logq : logp - logq;
                    }
                }
            } else {
                // ===================== reverse pass of transform k =====================
                const float w = wp;
                float yb[DP], yv[DP], sh[DP], mb[DP], sb[DP], zb[DP];
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
                        if (NLLM && live) {   // d loss / d (mean, log-scale) of the variable this position visits, real row order
                            a.em_o[((size_t)k * 2 * D + perm[jj]) * NP + pidx] = mb[jj];
                            a.em_o[((size_t)k * 2 * D + D + perm[jj]) * NP + pidx] = sb[jj];
                        }
                    } else {
                        mb[jj] = sb[jj] = yb[jj] = yv[jj] = sh[jj] = 0.f;
                    }
                }
                float* gAbase = a.gA + (size_t)k * H0 * a.B + b;
                const bool emit = (seq || MODE == TC_FKL || NLLM) && live;
                // layer-1 pieces in DESCENDING order (the first push covers every accumulator column)
                uint32_t cmn = bits1[(size_t)(np1 - 1) * a.bits_stride];
                for (int pi = np1 - 1; pi >= 0; --pi) {
                    const TcPiece c = P1[pi];
                    const int j = c.flags & 255;
                    const uint32_t cm = cmn;
                    if (pi > 0) cmn = bits1[(size_t)(pi - 1) * a.bits_stride];
                    if (seq && ((c.flags >> 8) & 1)) {   // entering degree group j (its last piece comes first): seed of position j
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
                        if (RSBM && live) {
                            a.em_o[((size_t)k * 2 * D + perm[j]) * NP + pidx] = -zbj;
                            a.em_o[((size_t)k * 2 * D + D + perm[j]) * NP + pidx] = fmaf(-ybj, yj, w);
                        }
                    }
                    // ---- layer-1 cotangents of the piece: g = relu'(.) * (WoT[v] . (mbar, sbar)), staged as (hi, lo)
#pragma unroll
                    for (int q = 0; q < TC_KC; q += 4) {
                        if (q < c.nz) {
                            float gs[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4* wo = reinterpret_cast<const float4*>(WoT + (size_t)(c.u0 + q + e) * 2 * DP);
                                float2 g2 = make_float2(0.f, 0.f);
#pragma unroll
                                for (int q2 = 0; q2 < DP / 2; ++q2) {
                                    const float4 wq = wo[q2];
                                    g2 = tc_ffma2(make_float2(wq.x, wq.y), make_float2(mb[2 * q2], sb[2 * q2]), g2);
                                    g2 = tc_ffma2(make_float2(wq.z, wq.w), make_float2(mb[2 * q2 + 1], sb[2 * q2 + 1]), g2);
                                }
                                gs[e] = ((cm >> (q + e)) & 1u) ? g2.x + g2.y : 0.f;
                                if (EMITM && live && q + e < c.nr) a.em_a1[((size_t)k * g.H[1] + c.r0 + q + e) * NP + pidx] = gs[e];
                            }
                            tc_stage4(trow, q, gs);
                        } else if (q < c.nu) {
                            tc_st4(trow + TC_SH + q, 0.f, 0.f, 0.f, 0.f);
                            tc_st4(trow + TC_SL + q, 0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    const bool group_done = (c.flags >> 9) & 1;   // descending: the group's first piece is staged last
                    const bool read_now = seq ? group_done : pi == 0;
                    const int pj_lo = seq ? p0start[j] : 0, pj_hi = seq ? p0start[j + 1] : np0;
                    uint32_t rmn = 0u;
                    if (read_now) rmn = bits0[(size_t)pj_lo * a.bits_stride];   // in flight during the push
                    push(bk, c, pi == np1 - 1);
                    if (!read_now) continue;
                    // ---- layer-0 cotangents that are final now: mask, push into d/dy, emit d/dA
                    for (int pj = pj_lo; pj < pj_hi; ++pj) {
                        const TcPiece r = P0[pj];
                        const uint32_t rm = rmn & (r.nr >= 32 ? 0xFFFFFFFFu : ((1u << r.nr) - 1u));   // padding units: no gradient
                        if (pj + 1 < pj_hi) rmn = bits0[(size_t)(pj + 1) * a.bits_stride];
                        float* gAp = gAbase + (size_t)r.r0 * a.B;
                        float4 nxt = tc_ld4(trow + (uint32_t)r.u0);
#pragma unroll
                        for (int q = 0; q < TC_KC; q += 4) {
                            if (q < r.nz) {
                                tc_wait_ld(nxt);
                                const float4 acc = nxt;
                                nxt = tc_ld4(trow + (uint32_t)(r.u0 + q + 4));
                                const float g0[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const bool on = (rm >> (q + e)) & 1u;
                                    const float gv = on ? g0[e] : 0.f;
#pragma unroll
                                    for (int q4 = 0; q4 < DP / 4; ++q4) {
                                        const float4 w4 = *reinterpret_cast<const float4*>(W0t + (size_t)(r.u0 + q + e) * DP + 4 * q4);
                                        yb[4 * q4] = fmaf(w4.x, gv, yb[4 * q4]);
                                        yb[4 * q4 + 1] = fmaf(w4.y, gv, yb[4 * q4 + 1]);
                                        yb[4 * q4 + 2] = fmaf(w4.z, gv, yb[4 * q4 + 2]);
                                        yb[4 * q4 + 3] = fmaf(w4.w, gv, yb[4 * q4 + 3]);
                                    }
                                    if (emit && on) atomicAdd(gAp + (size_t)(q + e) * a.B, gv);   // RED.ADD.F32 over the S samples of the row
                                    if (EMITM && live && q + e < r.nr) a.em_a0[((size_t)k * H0 + r.r0 + q + e) * NP + pidx] = gv;
                                }
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
            mark(9 + kind);
        }
        if (NLLM && live)
            for (int d = 0; d < D; ++d) a.theta_grad[(size_t)pidx * D + d] = vecp[oG + d];
        if ((MODE == TC_RSAMPLE || RSBM) && live) {
            if (a.theta_out)
                for (int d = 0; d < D; ++d) a.theta_out[(size_t)pidx * D + d] = vecp[oV + K * D + d];
            if (a.diff) a.diff[pidx] = logp;
        }
        if constexpr (RUN) {
            // ================= input gradient of the block's rows: gx[b,c] = (1/std[c]) sum_{k,u} W0_k[u,c] gA[k,u,b] =================
            __threadfence();                          // this thread's reductions into gA are performed before the block reads them
            __syncthreads();
            mark(2);
            resident = -1;
            const int CG = (C + 7) & ~7;
            const float* Gsrc = a.tct ? Gt2 : vec;    // the gradient tile [row][CG] the update reads
            if (a.tct) {
                // ---- tensor cores: D[:, 0 .. NB1) = sum_kk gA_kk^T W0c_kk (A = the cotangent chunk as (hi, lo) rows, B = the strip)
                for (int kk = 0; kk < K; ++kk) {
                    load_strip(1, kk);
                    const int n = NCA * RT, T = (int)blockDim.x;   // items = (chunk of 4 units, row), row fastest: coalesced slab reads
                    for (int base = tid; base < n; base += 4 * T) {
                        float v[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int idx = base + i * T;
                            const int q = idx / RT, r = idx - q * RT;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int u = 4 * q + e;
                                v[i][e] = (idx < n && u < H0 && r < nr) ? __ldcg(a.gA + ((size_t)kk * H0 + u) * a.B + b0 + r) : 0.f;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int idx = base + i * T;
                            if (idx < n) {
                                const int q = idx / RT, r = idx - q * RT;
                                const float h0 = tc_hi(v[i][0]), h1 = tc_hi(v[i][1]), h2 = tc_hi(v[i][2]), h3 = tc_hi(v[i][3]);
                                *reinterpret_cast<float4*>(Ahi + q * ACS + r * 4) = make_float4(h0, h1, h2, h3);
                                *reinterpret_cast<float4*>(Alo + q * ACS + r * 4) =
                                    make_float4(v[i][0] - h0, v[i][1] - h1, v[i][2] - h2, v[i][3] - h3);
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (4 * q + e < H0 && r < nr) __stcg(a.gA + ((size_t)kk * H0 + 4 * q + e) * a.B + b0 + r, 0.f);
                            }
                        }
                    }
                    tc_fence_async_smem();
                    bulk_wait();
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncthreads();
                    tail_mma(0u, a.nch1, a.NB1, kk == 0);
                }
                mark(3);
                const int row = 32 * (warp & 3) + lane, part = warp >> 2, nparts = nwarps >> 2;
                for (int i = part; i < (CG >> 2); i += nparts) {
                    float4 v = tc_ld4(trow0 + (uint32_t)(4 * i));
                    tc_wait_ld(v);
                    if (row < RT) *reinterpret_cast<float4*>(Gt2 + (size_t)row * CG + 4 * i) = v;
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();                      // gradient tile complete; A buffers, strip and tensor memory free
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mark(4);
                if (it + 1 < a.nb_iter) load_strip(0, 0);   // the next projection's first strip arrives while the update runs
            } else {
            const int ldc = g.t[0].ldc, KH = K * H0;  // the K context blocks [H0][ldc] land back to back: row kk * H0 + u
            {
                const uint32_t nbytes = (uint32_t)H0 * (uint32_t)ldc * 4u;
                bulk_expect(nbytes * (uint32_t)K);
                for (int kk = 0; kk < K; ++kk) bulk_piece(nbytes * (uint32_t)kk, g.f + g.t[kk].W0c, nbytes);
            }
            // cotangents of the block's rows -> Gs (24 loads in flight per thread); the slab is zeroed for the next iteration
            {
                const int n = KH * RP, T = (int)blockDim.x;
                for (int base = tid; base < n; base += 24 * T) {
                    float v[24];
#pragma unroll
                    for (int i = 0; i < 24; ++i) {
                        const int idx = base + i * T;
                        const int j = idx / RP, r = idx - j * RP;
                        v[i] = (idx < n && r < nr) ? __ldcg(a.gA + (size_t)j * a.B + b0 + r) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 24; ++i) {
                        const int idx = base + i * T;
                        if (idx < n) {
                            const int j = idx / RP, r = idx - j * RP;
                            Gs[idx] = v[i];
                            if (r < nr) __stcg(a.gA + (size_t)j * a.B + b0 + r, 0.f);
                        }
                    }
                }
            }
            bulk_wait();
            __syncthreads();
            mark(3);
            // (8 rows x 8 columns) tiles; the thread groups split the K H0 reduction rows among them
            const int RG = RT >> 3, NGC = (C + 7) >> 3, ntile = RG * NGC;
            const int G = max(1, min((int)blockDim.x / ntile, KH));
            const int grp = tid / ntile, tile = tid - grp * ntile;
            const bool worker = grp < G;
            const int cg = tile % NGC, rg = tile / NGC;   // lanes walk the columns: broadcast cotangent loads, 2-way tile stores
            float2 acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
            if (worker) {
                const int j0 = (int)((long long)KH * grp / G), j1 = (int)((long long)KH * (grp + 1) / G);
#pragma unroll 2
                for (int j = j0; j < j1; ++j) {
                    const float4 g0 = *reinterpret_cast<const float4*>(Gs + j * RP + 8 * rg);
                    const float4 g1 = *reinterpret_cast<const float4*>(Gs + j * RP + 8 * rg + 4);
                    const float4 w0 = *reinterpret_cast<const float4*>(Wb + j * ldc + 8 * cg);
                    const float4 w1 = *reinterpret_cast<const float4*>(Wb + j * ldc + 8 * cg + 4);
                    const float gr[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 gg = make_float2(gr[i], gr[i]);
                        acc[i][0] = tc_ffma2(make_float2(w0.x, w0.y), gg, acc[i][0]);
                        acc[i][1] = tc_ffma2(make_float2(w0.z, w0.w), gg, acc[i][1]);
                        acc[i][2] = tc_ffma2(make_float2(w1.x, w1.y), gg, acc[i][2]);
                        acc[i][3] = tc_ffma2(make_float2(w1.z, w1.w), gg, acc[i][3]);
                    }
                }
            }
            __syncthreads();                          // Gs and the context blocks are dead
            mark(4);
            if (it + 1 < a.nb_iter) load_ctx_weights();   // arrives while the update runs
            // the groups add their partial sums into the row-major gradient tile Gt[row][CG] in turn (128-bit accesses)
            for (int gi = 0; gi < G; ++gi) {
                if (worker && grp == gi) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4* o = reinterpret_cast<float4*>(vec + (size_t)(8 * rg + i) * CG + 8 * cg);
                        float4 v0 = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
                        float4 v1 = make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y);
                        if (gi) {
                            const float4 p0 = o[0], p1 = o[1];
                            v0 = make_float4(v0.x + p0.x, v0.y + p0.y, v0.z + p0.z, v0.w + p0.w);
                            v1 = make_float4(v1.x + p1.x, v1.y + p1.y, v1.z + p1.z, v1.w + p1.w);
                        }
                        o[0] = v0;
                        o[1] = v1;
                    }
                }
                __syncthreads();
            }
            }
            mark(5);
            // ================= advertorch perturb_iterative(ord=2) update, one warp per row; Xs for the next projection =================
            {
                constexpr int NE = 4;                  // context entries per lane (C <= 128: checked by the launcher)
                constexpr int NB = 6;                  // rows per warp handled together: their dependent chains (two warp
                                                       // reductions, sqrt, divisions) interleave instead of running one after the other
                float isd[NE], mu[NE], mk[NE];
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int c = lane + 32 * e;
                    isd[e] = (a.zstd && c < C) ? a.zstd[c] : 1.f;
                    mu[e] = (a.zmean && c < C) ? a.zmean[c] : 0.f;
                    mk[e] = (a.col_mask && c < C) ? a.col_mask[c] : 1.f;   // masked-feature attack: 0 = frozen column
                }
                for (int rb = warp; rb < RT; rb += NB * nwarps) {
                    float xv[NB][NE], dv[NB][NE], gv[NB][NE], n2[NB], d2[NB];
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int r = rb + q * nwarps;
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            const int c = lane + 32 * e;
                            const bool in = r < nr && c < C;
                            xv[q][e] = in ? a.x[(size_t)(b0 + r) * C + c] : 0.f;
                            dv[q][e] = in ? a.delta[(size_t)(b0 + r) * C + c] : 0.f;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int r = rb + q * nwarps;
                        n2[q] = 0.f;
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            const int c = lane + 32 * e;
                            float v = 0.f;
                            if (r < nr && c < C) {
                                v = Gsrc[(size_t)r * CG + c];
                                if (a.zstd) v /= isd[e];
                                v *= mk[e];
                            }
                            gv[q][e] = v;
                            n2[q] = fmaf(v, v, n2[q]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1)
#pragma unroll
                        for (int q = 0; q < NB; ++q) n2[q] += __shfl_xor_sync(0xffffffffu, n2[q], o);
                    if (n2[0] == 123.456f) Xn[0] = 1.f;
                    mark(13);
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const float inv = 1.f / fmaxf(sqrtf(n2[q]), a.floor_);
                        const int rq = rb + q * nwarps;
                        const float step = (a.eps_iter_rows && rq < nr) ? a.eps_iter_rows[b0 + rq] : a.eps_iter;
                        d2[q] = 0.f;
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            float d = dv[q][e] + (gv[q][e] * inv) * step;
                            d = fminf(fmaxf(xv[q][e] + d, a.cmin), a.cmax) - xv[q][e];
                            dv[q][e] = (lane + 32 * e < C && mk[e] != 0.f) ? d : 0.f;
                            d2[q] = fmaf(dv[q][e], dv[q][e], d2[q]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1)
#pragma unroll
                        for (int q = 0; q < NB; ++q) d2[q] += __shfl_xor_sync(0xffffffffu, d2[q], o);
                    if (d2[0] == 123.456f) Xn[0] = 1.f;
                    mark(14);
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int r = rb + q * nwarps;
                        if (r >= RT) continue;
                        const float radius = (a.eps_rows && r < nr) ? a.eps_rows[b0 + r] : a.eps;
                        const float factor = fminf(radius / sqrtf(d2[q]), 1.f);
#pragma unroll
                        for (int e = 0; e < NE; ++e) {
                            const int c = lane + 32 * e;
                            if (c < C) {
                                float v = 0.f;
                                if (r < nr) {
                                    const float d = dv[q][e] * factor;
                                    a.delta[(size_t)(b0 + r) * C + c] = d;
                                    v = xv[q][e] + d;
                                    if (a.zmean) v = (v - mu[e]) / isd[e];
                                }
                                if (a.tct) put_a(c, r, v);
                                else Xn[c * RP + r] = v;
                            }
                        }
                        if (a.tct)   // K-index padding of the A operand (the cotangent chunks used the same buffers)
                            for (int c = C + lane; c < 4 * NCA; c += 32) put_a(c, r, 0.f);
                    }
                }
                if (a.tct) tc_fence_async_smem();
            }
        }
        if constexpr (RUN) {
            mark(15);
            __syncthreads();
            mark(6);
        }
      }   // iterations
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}

// ------------------------------------------------------------------------------------------------- launchers
// The kernel instantiations of one padded dimension DP (7 modes + 2 persistent ones) are one translation unit's worth of
// compile time: this file is compiled once per DP (rabi_b200/_lib.py: -DRABI_TC_SIDE_DP=8 / =12 build only
// tc_launch_mode<DP>; the plain compile holds DP = 4 and all host code), and the objects build in parallel.
template <int MODE, int DP, bool RUN>
static int tc_launch_inst(rabi_maf* h, const TcArgs& a, int grid, int T, size_t smem, cudaStream_t st) {
    RABI_CU_OK(h, cudaFuncSetAttribute(k_tc<MODE, DP, RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = MODE == TC_GRAD || MODE == TC_FKL;   // the attack kernels are the profiled ones
    if (prof && rabi_prof_begin(h, st)) return 1;
    k_tc<MODE, DP, RUN><<<grid, T, smem, st>>>(h->img, a);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "k_tc launch failed: %s", cudaGetErrorString(e));
    if (prof && rabi_prof_end(h, st)) return 1;
    return 0;
}

template <int DP>
int tc_launch_mode(rabi_maf* h, const TcArgs& a, int mode, bool run, int grid, int T, size_t smem, cudaStream_t st) {
    if (run) {
        if (mode == TC_GRAD) return tc_launch_inst<TC_GRAD, DP, true>(h, a, grid, T, smem, st);
        if (mode == TC_FKL) return tc_launch_inst<TC_FKL, DP, true>(h, a, grid, T, smem, st);
        return 2;
    }
    switch (mode) {
        case TC_FWD: return tc_launch_inst<TC_FWD, DP, false>(h, a, grid, T, smem, st);
        case TC_GRAD: return tc_launch_inst<TC_GRAD, DP, false>(h, a, grid, T, smem, st);
        case TC_FKL: return tc_launch_inst<TC_FKL, DP, false>(h, a, grid, T, smem, st);
        case TC_RSAMPLE: return tc_launch_inst<TC_RSAMPLE, DP, false>(h, a, grid, T, smem, st);
        case TC_LOGPROB: return tc_launch_inst<TC_LOGPROB, DP, false>(h, a, grid, T, smem, st);
        case TC_NLL: return tc_launch_inst<TC_NLL, DP, false>(h, a, grid, T, smem, st);
        case TC_RSB: return tc_launch_inst<TC_RSB, DP, false>(h, a, grid, T, smem, st);
    }
    return 2;
}

#ifdef RABI_TC_SIDE_DP
#ifndef RABI_LEAN
template int tc_launch_mode<RABI_TC_SIDE_DP>(rabi_maf* h, const TcArgs& a, int mode, bool run, int grid, int T, size_t smem, cudaStream_t st);
#endif
#else
#ifndef RABI_LEAN
extern template int tc_launch_mode<8>(rabi_maf* h, const TcArgs& a, int mode, bool run, int grid, int T, size_t smem, cudaStream_t st);
extern template int tc_launch_mode<12>(rabi_maf* h, const TcArgs& a, int mode, bool run, int grid, int T, size_t smem, cudaStream_t st);
#endif

// ------------------------------------------------------------------------------------------------- image packing
struct TcPackMaps {
    const int* pu0;   // [K][H0] padded index of layer-0 unit u
    const int* pu1;   // [K][H1]
};
// float part of every block, from the masked order-space image (maf_image.h): grid (blocks, 2 orientations, K).
// The block was zero-filled by the host; only real units are written.
__global__ void k_tc_pack(const MafImg g, TcBlk b0, TcBlk b1, int blk_words, int DP, TcPackMaps mp, float* __restrict__ img) {
    const int k = blockIdx.z, ori = blockIdx.y;
    const TcBlk bk = ori == 0 ? b0 : b1;
    const TImg t = g.t[k];
    const int H0 = g.H[0], H1 = g.H[1], ld1 = t.ld[1];
    float* out = img + (size_t)(2 * k + ori) * blk_words;
    const float* f = g.f;
    const int* pu0 = mp.pu0 + (size_t)k * H0;
    const int* pu1 = mp.pu1 + (size_t)k * H1;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int o = t0; o < H1 * H0; o += stride) {
        const int v = o / H0, u = o % H0;
        // orientation 0: n = layer-1 unit v, k = layer-0 unit u;  orientation 1: n = u, k = v;  W1 is [v][u]
        const int n = ori == 0 ? pu1[v] : pu0[u], kidx = ori == 0 ? pu0[u] : pu1[v];
        const float w = f[t.W[1] + (size_t)v * ld1 + u];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
        const int at = ((kidx >> 2) * bk.NPN + n) * 4 + (kidx & 3);
        out[bk.whi + at] = hi;
        out[bk.wlo + at] = w - hi;
    }
    for (int o = t0; o < H0 * DP; o += stride) {
        const int u = o / DP, jj = o % DP;
        out[bk.w0t + pu0[u] * DP + jj] = jj < t.ldt ? f[t.W0t + (size_t)u * t.ldt + jj] : 0.f;
    }
    for (int o = t0; o < H1; o += stride) out[bk.b1 + pu1[o]] = f[t.b[1] + o];
    for (int o = t0; o < H1 * 2 * DP; o += stride) {
        const int v = o / (2 * DP), c = o % (2 * DP);
        out[bk.wot + pu1[v] * 2 * DP + c] = c < 2 * t.ldt ? f[t.WoT + (size_t)v * 2 * t.ldt + c] : 0.f;
    }
    for (int o = t0; o < 2 * DP; o += stride) out[bk.bo + o] = o < 2 * t.ldt ? f[t.boF + o] : 0.f;
}

// Context weights of the persistent kernel's tensor-core tail: per (orientation, transform) a (hi | lo) strip in the K-major
// core-matrix layout [k/4][n][4] (zero-filled by the host):
//   orientation 0 (projection):      n = layer-0 unit u, k = context column c      strip[hl][c/4][u][c%4] = W0c_k[u][c]
//   orientation 1 (back-projection): n = context column c, k = layer-0 unit u      strip[hl][u/4][c][u%4] = W0c_k[u][c]
__global__ void k_tc_pack_ctx(const MafImg g, int strip_words, int nch0, int NB0, int nch1, int NB1, float* __restrict__ img) {
    const int k = blockIdx.y, C = g.C, H0 = g.H[0];
    const TImg t = g.t[k];
    float* s0 = img + (size_t)k * strip_words;
    float* s1 = img + ((size_t)g.K + k) * strip_words;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < H0 * C; o += gridDim.x * blockDim.x) {
        const int u = o / C, c = o % C;
        const float w = g.f[t.W0c + (size_t)u * t.ldc + c];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u), lo = w - hi;
        const int a0 = ((c >> 2) * NB0 + u) * 4 + (c & 3), a1 = ((u >> 2) * NB1 + c) * 4 + (u & 3);
        s0[a0] = hi;
        s0[(size_t)nch0 * NB0 * 4 + a0] = lo;
        s1[a1] = hi;
        s1[(size_t)nch1 * NB1 * 4 + a1] = lo;
    }
}

struct TcState {
    bool built = false, eligible = false;
    int DP = 4;
    TcBlk blk[2];
    int blk_words = 0;
    float* img = nullptr;  // [template: integer tables, zeros elsewhere][packed image]
    size_t img_words = 0;
    int* maps = nullptr;   // pu0 [K][H0] then pu1 [K][H1]
    unsigned long long packed_version = ~0ull;
    uint32_t* bits = nullptr;
    size_t bits_words = 0;
    float* pimg = nullptr;   // strips of the persistent kernel's tensor-core tail
    size_t pimg_words = 0;
    unsigned long long pimg_version = ~0ull;
};

static inline int r4(int x) { return (x + 3) & ~3; }
static inline int r8(int x) { return (x + 7) & ~7; }

// pieces of one layer.  grp: [D+1] real group boundaries; gp: [D+1] padded group starts.
// forward (K side = this layer, N side = next layer with padded starts np_next / width NPN): inputs of degree j feed
// outputs of degree >= j;  reverse: cotangents of degree j feed the units of degree <= j of the layer below.
static void tc_pieces(int D, const int* grp, const int* gp, const int* gp_other, int NPN, bool reverse,
                      std::vector<TcPiece>& out, std::vector<int>& start) {
    out.clear();
    start.assign(D + 1, 0);
    for (int j = 0; j < D; ++j) {
        start[j] = (int)out.size();
        const int len = grp[j + 1] - grp[j], plen = r4(len);
        for (int o = 0; o < plen; o += TC_KC) {
            TcPiece c;
            c.u0 = gp[j] + o;
            c.nz = std::min(TC_KC, plen - o);
            c.nu = r8(c.nz);
            c.r0 = grp[j] + o;
            c.nr = std::min(TC_KC, len - o);
            if (!reverse) {
                c.n0 = gp_other[j] & ~7;
                c.N = NPN - c.n0;
            } else {
                c.n0 = 0;
                c.N = std::min(NPN, r8(gp_other[j + 1]));
            }
            c.flags = j | ((o + TC_KC >= plen) ? 256 : 0) | (o == 0 ? 512 : 0);
            out.push_back(c);
        }
    }
    start[D] = (int)out.size();
}

static int tc_build(rabi_maf* h, TcState* s) {
    s->built = true;
    s->eligible = false;
    const int K = h->K, D = h->D, H0 = h->H[0], H1 = h->H[1];
    if (h->L != 2 || D > 12) return 0;
    const std::vector<int>& M = h->meta_host;
    const MafImg& g = h->img;
    // padded layouts (per transform; the block layout must be the same for all of them)
    std::vector<int> gp0((size_t)K * (D + 1)), gp1((size_t)K * (D + 1)), maps((size_t)K * (H0 + H1));
    int HP0 = 0, HP1 = 0;
    for (int k = 0; k < K; ++k) {
        for (int l = 0; l < 2; ++l) {
            std::vector<int>& gp = l == 0 ? gp0 : gp1;
            const int* grp = &M[g.t[k].grp[l]];
            int at = 0;
            for (int j = 0; j < D; ++j) {
                if (grp[j + 1] <= grp[j]) return 0;   // an empty degree group: not a shape this path tiles
                gp[(size_t)k * (D + 1) + j] = at;
                for (int u = grp[j]; u < grp[j + 1]; ++u) maps[(size_t)(l == 0 ? 0 : K * H0) + (size_t)k * (l == 0 ? H0 : H1) + u] = at + (u - grp[j]);
                at += r4(grp[j + 1] - grp[j]);
            }
            gp[(size_t)k * (D + 1) + D] = at;
            (l == 0 ? HP0 : HP1) = std::max(l == 0 ? HP0 : HP1, at);
        }
    }
    if (HP0 > TC_ACC || HP1 > TC_ACC) return 0;
    const int DP = (D + 3) & ~3;
    s->DP = DP;
    const int R0 = r8(HP0) + 8, R1 = r8(HP1) + 8;
    for (int ori = 0; ori < 2; ++ori) {
        TcBlk& b = s->blk[ori];
        b.NPN = ori == 0 ? r8(HP1) : r8(HP0);
        b.NPK = ori == 0 ? R0 : R1;   // the last piece's K = 8 steps may run past the last group
        int o = 0;
        auto take = [&](int n) {
            int at = o;
            o += (n + 3) & ~3;
            return at;
        };
        const int nW = (b.NPK / 4) * b.NPN * 4;
        b.whi = take(nW);
        b.wlo = take(nW);
        b.w0t = take(R0 * DP);
        b.b1 = take(R1);
        b.wot = take(R1 * 2 * DP);
        b.bo = take(2 * DP);
        b.perm = take(DP);
        b.permy = take(DP);
        b.np = take(4);
        b.p0start = take(D + 1);
        b.p1start = take(D + 1);
        b.P0 = take(TC_MAXP * 8);
        b.P1 = take(TC_MAXP * 8);
        b.words = o;
    }
    s->blk_words = std::max(s->blk[0].words, s->blk[1].words);
    std::vector<float> host((size_t)K * 2 * s->blk_words, 0.f);
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        const int* grp0 = &M[t.grp[0]];
        const int* grp1 = &M[t.grp[1]];
        const int* p0 = &gp0[(size_t)k * (D + 1)];
        const int* p1 = &gp1[(size_t)k * (D + 1)];
        for (int ori = 0; ori < 2; ++ori) {
            const TcBlk& b = s->blk[ori];
            int* w = reinterpret_cast<int*>(host.data() + (size_t)(2 * k + ori) * s->blk_words);
            for (int j = 0; j < D; ++j) {
                w[b.perm + j] = M[t.perm + j];
                w[b.permy + j] = M[t.permy + j];
            }
            std::vector<TcPiece> P0, P1;
            std::vector<int> s0, s1;
            // layer-0 pieces carry the forward push (into layer-1 columns), layer-1 pieces the reverse push (into layer-0 columns)
            tc_pieces(D, grp0, p0, p1, s->blk[0].NPN, false, P0, s0);
            tc_pieces(D, grp1, p1, p0, s->blk[1].NPN, true, P1, s1);
            if ((int)P0.size() > TC_MAXP || (int)P1.size() > TC_MAXP) return 0;
            w[b.np] = (int)P0.size();
            w[b.np + 1] = (int)P1.size();
            for (int j = 0; j <= D; ++j) {
                w[b.p0start + j] = s0[j];
                w[b.p1start + j] = s1[j];
            }
            memcpy(w + b.P0, P0.data(), P0.size() * sizeof(TcPiece));
            memcpy(w + b.P1, P1.data(), P1.size() * sizeof(TcPiece));
        }
    }
    if (s->img) cudaFree(s->img);
    if (s->maps) cudaFree(s->maps);
    s->img = nullptr;
    s->maps = nullptr;
    s->img_words = host.size();
    RABI_CU_OK(h, cudaMalloc(&s->img, 2 * host.size() * sizeof(float)));
    RABI_CU_OK(h, cudaMemcpy(s->img, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    RABI_CU_OK(h, cudaMalloc(&s->maps, maps.size() * sizeof(int)));
    RABI_CU_OK(h, cudaMemcpy(s->maps, maps.data(), maps.size() * sizeof(int), cudaMemcpyHostToDevice));
    s->packed_version = ~0ull;
    s->eligible = true;
    return 0;
}

static int tc_dispatch_dp(rabi_maf* h, TcState* s, const TcArgs& a, int mode, bool run, int grid, int T, size_t smem, cudaStream_t st) {
    switch (s->DP) {
        case 4: return tc_launch_mode<4>(h, a, mode, run, grid, T, smem, st);
#ifndef RABI_LEAN
        case 8: return tc_launch_mode<8>(h, a, mode, run, grid, T, smem, st);
        case 12: return tc_launch_mode<12>(h, a, mode, run, grid, T, smem, st);
#endif
    }
    return 2;
}

// state, packed image and the mode-independent arguments; 0 = ready, 1 = error, 2 = not eligible
static int tc_prepare(rabi_maf* h, const TcCall& c, int mode, cudaStream_t st, TcState** out, TcArgs* ap) {
    TcState* s = static_cast<TcState*>(h->tc);
    if (!s) h->tc = s = new TcState();
    if (!s->built && tc_build(h, s)) return 1;
    if (!s->eligible) return 2;
    if (s->packed_version != h->weights_version) {
        // padding entries stay zero: start from the template, then write the real units
        RABI_CU_OK(h, cudaMemcpyAsync(s->img + s->img_words, s->img, s->img_words * sizeof(float), cudaMemcpyDeviceToDevice, st));
        dim3 grid(16, 2, h->K);
        TcPackMaps mp = {s->maps, s->maps + (size_t)h->K * h->H[0]};
        k_tc_pack<<<grid, 256, 0, st>>>(h->img, s->blk[0], s->blk[1], s->blk_words, s->DP, mp, s->img + s->img_words);
        h->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return rabi_fail(h, "k_tc_pack launch failed: %s", cudaGetErrorString(e));
        s->packed_version = h->weights_version;
    }
    TcArgs& a = *ap;
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
    a.wrow = c.wrow;
    a.em_h0 = c.em_h0;
    a.em_a0 = c.em_a0;
    a.em_h1 = c.em_h1;
    a.em_a1 = c.em_a1;
    a.em_o = c.em_o;
    a.em_y = c.em_y;
    a.theta_grad = c.theta_grad;
    a.npairs = (long long)c.S * c.B;
    a.img = s->img + s->img_words;
    a.blk[0] = s->blk[0];
    a.blk[1] = s->blk[1];
    a.blk_words = s->blk_words;
    a.region_words = s->blk_words;
    a.VS = ((4 * h->K + 2) * h->D) | 1;
    if (mode == TC_GRAD || mode == TC_FKL || mode == TC_NLL || mode == TC_RSB) {
        const size_t need = (size_t)2 * h->K * 2 * TC_MAXP * (size_t)h->sm_count * TC_MAXTILES * 128;
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
    *out = s;
    return 0;
}

int rabi_tc_launch(rabi_maf* h, const TcCall& c, int mode, cudaStream_t st) {
    const long long npairs = (long long)c.S * c.B;
    if (npairs <= 0) return 0;
    if (!rabi_env_int("RABI_TC", 1)) return 2;
    if (h->L == 3) return rabi_tcd_launch(h, c, mode, st);
    if (npairs < (long long)rabi_env_int("RABI_TC_MIN_PAIRS", 1)) return 2;
    TcState* s = nullptr;
    TcArgs a;
    const int rc = tc_prepare(h, c, mode, st, &s, &a);
    if (rc) return rc;
    const long long nt = (npairs + 127) / 128;
    if (nt > 0x7fffffff) return 2;
    a.ntiles = (int)nt;
    // tiles per CTA: as many as the per-pair state next to the weight block allows, no more than the launch needs
    const size_t budget = (size_t)h->max_smem_optin - 1024;
    int NT = std::min(TC_MAXTILES, std::max(1, rabi_env_int("RABI_TC_TILES", TC_MAXTILES)));
    while (NT > 1 && ((size_t)s->blk_words + (size_t)NT * 128 * a.VS) * 4 > budget) --NT;
    if (((size_t)s->blk_words + (size_t)NT * 128 * a.VS) * 4 > budget) return 2;
    const int grid = (int)std::min<long long>(nt, h->sm_count);
    while (NT > 1 && (long long)grid * (NT - 1) >= nt) --NT;   // spread the tiles over the SMs first
    const int T = NT * 128;
    const size_t smem = ((size_t)s->blk_words + (size_t)T * a.VS) * 4;
    return tc_dispatch_dp(h, s, a, mode, false, grid, T, smem, st);
}

// The whole L2-PGD attack in ONE launch (north_star kernel 3: "PGD step, L2 projection and eps-ball clamp in one launch"):
// rows are independent, so a CTA keeps its row blocks for all nb_iter iterations and runs projection -> 12 flow phases ->
// gradient back-projection -> advertorch perturb_iterative(ord=2) update on chip / in its own L2-resident slabs.
// Reference loop body: advertorch perturb_iterative as called at rbi/rbi/attacks/advertorch_attack.py:100-114.
int rabi_tc_run(rabi_maf* h, const TcRunCall& r, int mode, cudaStream_t st) {
    if (r.S <= 0 || r.B <= 0 || r.nb_iter <= 0) return 2;
    if (!rabi_env_int("RABI_TC", 1) || !rabi_env_int("RABI_TC_RUN", 1)) return 2;
    if (h->L != 2) return 2;
    if ((long long)r.S * r.B < (long long)rabi_env_int("RABI_TC_RUN_MIN_PAIRS", 1025)) return 2;
    TcCall c = {r.A_p, r.A_q, r.z_all, nullptr, nullptr, r.gA, r.S, r.B, r.inv_B / (float)r.S};
    TcState* s = nullptr;
    TcArgs a;
    const int rc = tc_prepare(h, c, mode, st, &s, &a);
    if (rc) return rc;
    const MafImg& g = h->img;
    const int C = h->C, H0 = h->H[0], K = h->K;
    // the transposed context weights of all transforms (projection) and the K context blocks (back-projection) take turns in
    // the weight-block region; bulk copies need 16-byte granularity
    if ((g.ldKH & 3) || (g.W0cT & 3)) return 2;
    size_t wsum = 0;
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        if ((t.ldc & 3) || (t.W0c & 3)) return 2;
        wsum += (size_t)H0 * t.ldc;
    }
    // (the 8-wide unit / column tiles of the two FMA products may read a few floats past a row end: still inside shared memory,
    //  and those accumulators are never stored)
    // tensor-core tail: strips of both orientations padded to one size; D needs K * NB0 tensor-memory columns
    const int nch0 = 2 * ((C + 7) / 8), NB0 = (H0 + 7) & ~7, nch1 = 2 * ((H0 + 7) / 8), NB1 = (C + 7) & ~7;
    const size_t strip_words = (size_t)2 * 4 * std::max((size_t)nch0 * NB0, (size_t)nch1 * NB1);
    const bool tct = rabi_env_int("RABI_TC_TAIL", 1) && K * NB0 <= 512 && NB0 <= 256 && NB1 <= 256;
    const size_t region = ((tct ? std::max((size_t)s->blk_words, strip_words)
                                : std::max(std::max((size_t)s->blk_words, (size_t)C * g.ldKH), wsum)) + 31) & ~(size_t)31;
    a.region_words = (int)region;
    const size_t budget = (size_t)h->max_smem_optin - 1024;
    int NT = TC_MAXTILES;
    while (NT > 1 && (region + (size_t)NT * 128 * a.VS) * 4 > budget) --NT;
    if ((region + (size_t)NT * 128 * a.VS) * 4 > budget) return 2;
    // rows per block: all S samples of a row in one CTA; as few rounds of blocks as possible, every CTA the same number of rows.
    // The per-pair region also holds, in turn, the cotangent tile [K H0][RP] and the gradient tile + Xs (2 C RP).
    if (C > 128) return 2;                               // the update keeps a row in 4 registers per lane
    const size_t spare = budget / 4 - region;            // words available behind the weight region
    auto tail_words = [&](int R_) {
        const size_t RP_ = (size_t)((R_ + 7) & ~7) + 4;
        const size_t CG_ = (size_t)((C + 7) & ~7);
        if (tct)   // A-operand buffers (hi | lo) [NCA][RT][4] + gradient tile [RT][CG] (+ the MMA's reads past row RT: finite garbage)
            return (size_t)2 * std::max(nch0, nch1) * ((RP_ - 4) * 4 + 4) + (RP_ - 4) * CG_ + 1024;
        return std::max((size_t)K * H0 * RP_, RP_ * CG_ + (size_t)C * RP_);   // cotangent tile | gradient tile [RP][CG] + Xs [C][RP]
    };
    int Rmax = ((NT * 128) / r.S) & ~3;
    // (tensor-core tail: the projection epilogue transposes [K H0][round8(R)] through the weight region)
    while (Rmax >= 4 && (tail_words(Rmax) > spare || (tct && (size_t)K * H0 * ((Rmax + 7) & ~7) > region))) Rmax -= 4;
    if (Rmax < 4) return 2;
    const int rounds = (r.B + h->sm_count * Rmax - 1) / (h->sm_count * Rmax);
    int R = ((r.B + h->sm_count * rounds - 1) / (h->sm_count * rounds) + 3) & ~3;
    R = std::min(R, Rmax);
    const int RP = ((R + 7) & ~7) + 4;
    const int nblocks = (r.B + R - 1) / R;
    const int gtiles = ((C + 7) / 8) * ((RP - 4) / 8);         // one (8 rows x 8 columns) gradient tile per thread
    if (gtiles > NT * 128) return 2;
    while (NT > 1 && (NT - 1) * 128 >= r.S * R && (NT - 1) * 128 >= gtiles) --NT;
    a.x = r.x;
    a.delta = r.delta;
    a.zmean = r.zmean;
    a.zstd = r.zstd;
    a.nb_iter = r.nb_iter;
    a.R = R;
    a.RP = RP;
    a.nblocks = nblocks;
    a.eps = r.eps;
    a.eps_iter = r.eps_iter;
    a.cmin = r.cmin;
    a.cmax = r.cmax;
    a.floor_ = r.floor_;
    a.eps_rows = r.eps_rows;
    a.eps_iter_rows = r.eps_iter_rows;
    a.col_mask = r.col_mask;
    a.zstride = (long long)r.S * r.B * h->D;
    const int grid = std::min(nblocks, h->sm_count);
    const int T = NT * 128;
    a.tct = tct ? 1 : 0;
    if (tct) {
        const size_t need = (size_t)2 * K * strip_words;
        if (need > s->pimg_words) {
            if (s->pimg) {
                RABI_CU_OK(h, cudaDeviceSynchronize());
                RABI_CU_OK(h, cudaFree(s->pimg));
                s->pimg = nullptr;
            }
            RABI_CU_OK(h, cudaMalloc(&s->pimg, need * sizeof(float)));
            s->pimg_words = need;
            s->pimg_version = ~0ull;
        }
        if (s->pimg_version != h->weights_version) {
            RABI_CU_OK(h, cudaMemsetAsync(s->pimg, 0, need * sizeof(float), st));
            k_tc_pack_ctx<<<dim3(16, K), 256, 0, st>>>(h->img, (int)strip_words, nch0, NB0, nch1, NB1, s->pimg);
            h->launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return rabi_fail(h, "k_tc_pack_ctx launch failed: %s", cudaGetErrorString(e));
            s->pimg_version = h->weights_version;
        }
        a.pimg = s->pimg;
        a.strip_words = (int)strip_words;
        a.nch0 = nch0;
        a.NB0 = NB0;
        a.nch1 = nch1;
        a.NB1 = NB1;
    }
    a.vec_words = (int)std::max((size_t)T * a.VS, tail_words(R));
    a.xs_off = a.vec_words - C * RP;
    const size_t smem = (region + (size_t)a.vec_words) * 4;
    if (smem > budget) return 2;
    // The slabs the kernel rewrites every iteration ([A_pert | gA], contiguous in the caller's workspace) are CTA-private and
    // re-read within microseconds, while 0.8 MB of fresh base noise per iteration streams through L2: ask for the slabs to stay
    // resident (persisting hits, everything else streaming) for the duration of this launch.
    struct L2Window {
        cudaStream_t st;
        bool on = false;
        L2Window(rabi_maf* h, cudaStream_t st_, const void* base, size_t bytes) : st(st_) {
            if (!rabi_env_int("RABI_TC_L2_PERSIST", 1)) return;
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
            if (max_persist <= 0 || max_window <= 0) return;
            static size_t limit_set = 0;
            const size_t want = std::min<size_t>(bytes, (size_t)max_persist);
            // the carve-out is set-associative: some slack over the window keeps colliding lines resident too
            const size_t carve = std::min<size_t>((size_t)max_persist, bytes * (size_t)rabi_env_int("RABI_TC_L2_CARVE_PCT", 100) / 100);
            if (carve > limit_set) {
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) {
                    cudaGetLastError();
                    return;
                }
                limit_set = carve;
            }
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            av.accessPolicyWindow.base_ptr = const_cast<void*>(base);
            av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            on = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
            if (!on) cudaGetLastError();
        }
        ~L2Window() {
            if (!on) return;
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
        }
    } l2win(h, st, mode == TC_GRAD ? (const void*)r.A_p : (const void*)r.A_q, (size_t)2 * K * H0 * r.B * sizeof(float));
    if (rabi_env_int("RABI_TC_RUN_DBG", 0)) {   // development aid: cycle counters of CTA 0 per section of an iteration
        long long* d = nullptr;
        RABI_CU_OK(h, cudaMalloc(&d, 16 * sizeof(long long)));
        RABI_CU_OK(h, cudaMemsetAsync(d, 0, 16 * sizeof(long long), st));
        a.dbg = d;
        int rc2 = tc_dispatch_dp(h, s, a, mode, true, grid, T, smem, st);
        long long hc[16];
        RABI_CU_OK(h, cudaStreamSynchronize(st));
        RABI_CU_OK(h, cudaMemcpy(hc, d, sizeof(hc), cudaMemcpyDeviceToHost));
        cudaFree(d);
        static const char* names[7] = {"wait ctx weights", "projection", "flow phases", "stage gA", "gx product", "reduce groups", "update"};
        fprintf(stderr, "rabi_tc_run: R=%d RP=%d blocks=%d T=%d smem=%zu; cycles per iteration of CTA 0:", R, RP, nblocks, T, smem);
        for (int i = 0; i < 7; ++i) fprintf(stderr, " %s %.0f;", names[i], (double)hc[i] / r.nb_iter);
        fprintf(stderr, " | inside the flow phases: block loads %.0f; sampling fwd %.0f; density fwd %.0f; density bwd %.0f; sampling bwd %.0f;",
                (double)hc[8] / r.nb_iter, (double)hc[9] / r.nb_iter, (double)hc[10] / r.nb_iter, (double)hc[11] / r.nb_iter, (double)hc[12] / r.nb_iter);
        fprintf(stderr, " | update of warp 0: operands + first reduction %.0f; second reduction %.0f; stores %.0f; wait for the other warps %.0f",
                (double)hc[13] / r.nb_iter, (double)hc[14] / r.nb_iter, (double)hc[15] / r.nb_iter, (double)hc[6] / r.nb_iter);
        fprintf(stderr, "\n");
        return rc2;
    }
    return tc_dispatch_dp(h, s, a, mode, true, grid, T, smem, st);
}

void rabi_tc_destroy(rabi_maf* h) {
    rabi_tcd_destroy(h);
    rabi_train_destroy(h);
    TcState* s = static_cast<TcState*>(h->tc);
    if (!s) return;
    if (s->img) cudaFree(s->img);
    if (s->maps) cudaFree(s->maps);
    if (s->bits) cudaFree(s->bits);
    if (s->pimg) cudaFree(s->pimg);
    delete s;
    h->tc = nullptr;
}

#endif  // !RABI_TC_SIDE_DP

}  // namespace rabi
