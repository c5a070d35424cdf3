// Tensor-core path for THREE-hidden-layer MADE conditioners up to 200 units wide, D <= 32 (the pyloric posterior:
// config/experiment/train_pyloric.yaml:11-16, hidden [200, 200, 200], D = 31), replacing the global-memory-scratch deep
// path (maf_deep.cuh) whose per-pair state made 25x the algorithmic DRAM traffic.
//
// Reference arithmetic: Pyro ConditionalAutoRegressiveNN / AffineAutoregressive._call / ._inverse as driven by
// rbi/rbi/models/transforms.py:37-46, chained by TransformedDistribution (rbi/rbi/models/base.py:417-446), and the MC
// KL estimators (rbi/rbi/loss/kl_div.py:38-56); same one-pass-equivalent sampling and phase order as maf_fast.cuh.
//
// Layout of one CTA (160 threads): warps 0-3 own ONE tile of 128 (observation row, MC sample) pairs (thread = pair =
// tensor-memory lane); warp 4 is the weight PRODUCER.
//  * tensor memory (512 columns): R0 = [0,200) and R1 = [200,400) hold the fp32 accumulators of two hidden layers
//    (column = unit), R2 = [400,464) the (mean, log-scale) accumulators of the D positions, [464,496) two staging
//    buffers of 8 inputs as TF32 (hi, lo) halves.  ALL per-pair activations stay on chip.
//  * every hidden -> hidden and hidden -> output product is a "group push": the threads ReLU the <= 8 units of one
//    MADE degree group, stage them (tcgen05.st), one elected thread issues acc[:, n0:n0+N] += h_hi W_hi + h_lo W_hi +
//    h_hi W_lo (3 x tcgen05.mma.cta_group::1.kind::tf32, M = 128, K = 8, A from tensor memory) and commits to the
//    staging buffer's mbarrier; the next layer's units of the same degree are final after the push and are read back
//    with tcgen05.ld.  In the sequential sampling pass that is 2 exposed round trips per autoregressive step (the own
//    group's part of the output layer is 16 FMAs, its push for the later positions runs behind the next step).
//  * the weight strip of each push ([2 k-chunks][N][4] floats, hi then lo, <= 12.8 KB, packed per push in consumption
//    order) STREAMS from the L2-resident packed image through a 6-slot shared-memory ring: the producer thread issues
//    cp.async.bulk (TMA bulk copy) with mbarrier complete_tx, the issuing thread waits on the slot's "full" barrier and
//    releases it with a second tcgen05.commit onto its "empty" barrier.
//  * layer 0 (theta part, K = D) and the reverse of the output layer (2 values per position) are FMA code on weights
//    resident in shared memory; the per-pair chain vectors live in a coalesced global scratch ([slot][pair]), the
//    ReLU masks as one word per piece.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "maf_tc.h"

namespace rabi {

constexpr int TD_R0 = 0, TD_R1 = 200, TD_R2 = 400, TD_ST = 464;   // staging buffer b (of TD_NB): hi at TD_ST + 16 b, lo 8 columns further
constexpr int TD_MAXH = 200;
constexpr int TD_NSLOT = 6;
constexpr int TD_SLOT = 13312;     // bytes per ring slot (>= 2 halves x 2 k-chunks x 200 rows x 16 B), multiple of 128
constexpr int TD_MAXP = 64;        // pieces per layer
constexpr int TD_THREADS = 192;    // warps 0-3: the tile, warp 4: weight producer, warp 5: MMA issuer
constexpr int TD_NB = 3;           // staging buffers of 8 inputs (hi, lo)

// piece flags
constexpr int TD_F_FWD = 1;   // the piece pushes into the next layer / the outputs (forward stream)
constexpr int TD_F_REV = 2;   // the piece pushes its cotangents into the previous layer (reverse stream)
constexpr int TD_F_DGN = 4;   // its product with the SAME degree group of the NEXT layer is an FMA block (zeroed in the strips)
constexpr int TD_F_DGP = 8;   // the same for the group of the PREVIOUS layer
struct TdPiece { int r0, nr, j, flags; };   // real units [r0, r0 + nr), nr <= 8, of degree j
struct TdPush { int off16, bytes, dcol, N; };   // strip at image + 16 * off16; accumulator columns [dcol & 0xffff, .. + N); bit 16: overwrite

struct TdBlk {   // word offsets inside the resident per-transform block
    int w0t, b1, b2, bo, wot;       // [H0][DP], [H1], [H2], [2 DP] (m, s interleaved), [H2][2 DP]
    int dg01, dg12;                 // [D][8][8] own-degree blocks of W1 / W2 (rows: next layer's group, columns: this layer's)
    int perm, permy;
    int np, pstart, pieces;         // np[3]; pstart[3][D + 1]; pieces[3][TD_MAXP]
    int dgi;                        // [2][D][4] (row0, rows, col0, cols) of the blocks above; rows = 0: no block
    int words;
};

struct TdArgs {
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
    const float* blocks;   // [K][blk.words]
    TdBlk blk;
    const char* strips;    // packed strips
    const TdPush* push;    // push tables
    int ps_off[RABI_MAXK][2], ps_cnt[RABI_MAXK][2];   // per transform: forward and reverse push streams
    float* vec;            // [(4K+2) D][stride]
    uint8_t* bits;         // [2K][3][TD_MAXP][stride] ReLU masks, one BYTE per piece (a piece has <= 8 units)
    int stride;            // grid * 128
};

// ------------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t td_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t td_make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t td_make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void td_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void td_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(td_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void td_mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(td_smem_u32(mbar)), "r"(count) : "memory");
}
// wait of the producer / issuer threads: they share a scheduler with a compute warp, so they back off between polls
__device__ __forceinline__ void td_mbar_wait_bg(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(td_smem_u32(mbar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(32);
        if (++spins > 40000000u) {   // ~2 s: never hang the GPU
            printf("rabi_b200 k_tcd: background mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void td_mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    long long t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(td_smem_u32(mbar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins == 1024u) t0 = clock64();
        if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 4000000000LL) {  // ~2 s: never hang the GPU
            printf("rabi_b200 k_tcd: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void td_ld8(uint32_t addr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]) :: "memory");
}
__device__ __forceinline__ void td_ld2(uint32_t addr, float& a, float& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b) :: "memory");
}
__device__ __forceinline__ void td_st8(uint32_t addr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ float td_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float2 td_ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// ------------------------------------------------------------------------------------------------- the kernel
template <int MODE, int DP>
__global__ void __launch_bounds__(TD_THREADS, 1) k_tcd(const MafImg g, const TdArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr bool GRADM = MODE == TC_GRAD || MODE == TC_FKL;
    unsigned char* ring = smraw;                                                       // [TD_NSLOT][TD_SLOT]
    float* Wb = reinterpret_cast<float*>(smraw + TD_NSLOT * TD_SLOT);                   // resident block of transform k
    const int* Wi = reinterpret_cast<const int*>(Wb);
    float2* ms = reinterpret_cast<float2*>(Wb + a.blk.words);                           // [DP][128] (mbar, sbar) per position
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar_blk, bar_staged[TD_NB], bar_done[TD_NB], bar_full[TD_NSLOT], bar_empty[TD_NSLOT];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = g.K, D = g.D, H0 = g.H[0];
    const float HALF_LOG_2PI = 0.91893853320467274178f;

    if (tid == 0) {
        td_mbar_init(&bar_blk, 1);
        for (int i = 0; i < TD_NB; ++i) {
            td_mbar_init(&bar_staged[i], 128);
            td_mbar_init(&bar_done[i], 1);
        }
        for (int i = 0; i < TD_NSLOT; ++i) {
            td_mbar_init(&bar_full[i], 1);
            td_mbar_init(&bar_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(td_smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int rounds = (a.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nphase = (MODE == TC_GRAD ? 4 : MODE == TC_FKL ? 3 : MODE == TC_RSAMPLE ? 1 : 2) * K;
    const int ph0 = MODE == TC_LOGPROB ? K : 0;

    if (warp == 4) {
        // ===================== weight producer: streams the strips of every push, in consumption order =====================
        if (tid == 128) {
            unsigned e = 0;
            for (int round = 0; round < rounds; ++round) {
                if (round * (int)gridDim.x + (int)blockIdx.x >= a.ntiles) break;
                for (int ph = ph0; ph < nphase; ++ph) {
                    const int kind = ph / K, idx = ph - kind * K;
                    const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
                    const TdPush* tbl = a.push + a.ps_off[k][kind >> 1];
                    const int n = a.ps_cnt[k][kind >> 1];
                    for (int i = 0; i < n; ++i, ++e) {
                        const unsigned slot = e % TD_NSLOT, par = (e / TD_NSLOT) & 1u;
                        td_mbar_wait_bg(&bar_empty[slot], par ^ 1u);
                        const int off16 = __ldg(&tbl[i].off16), bytes = __ldg(&tbl[i].bytes);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(td_smem_u32(&bar_full[slot])), "r"((uint32_t)bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"(td_smem_u32(ring + slot * TD_SLOT)), "l"(a.strips + (size_t)off16 * 16), "r"((uint32_t)bytes),
                                        "r"(td_smem_u32(&bar_full[slot])) : "memory");
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer: one thread turns every staged group into a 3xTF32 push =====================
        if (tid == 160) {
            const uint32_t tcol = tmem_base_s;
            unsigned e = 0;
            for (int round = 0; round < rounds; ++round) {
                if (round * (int)gridDim.x + (int)blockIdx.x >= a.ntiles) break;
                for (int ph = ph0; ph < nphase; ++ph) {
                    const int kind = ph / K, idx = ph - kind * K;
                    const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
                    const TdPush* tbl = a.push + a.ps_off[k][kind >> 1];
                    const int n = a.ps_cnt[k][kind >> 1];
                    for (int i = 0; i < n; ++i, ++e) {
                        const unsigned slot = e % TD_NSLOT, sb = e % TD_NB;
                        const int N = __ldg(&tbl[i].N), dcol = __ldg(&tbl[i].dcol);
                        const uint32_t acc = (dcol >> 16) ? 0u : 1u;   // bit 16: first push into its region (overwrite)
                        td_mbar_wait_bg(&bar_staged[sb], (e / TD_NB) & 1u);
                        td_mbar_wait_bg(&bar_full[slot], (e / TD_NSLOT) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t idesc = td_make_idesc(128, N);
                        const uint32_t lbo = (uint32_t)N * 16u;
                        const uint32_t sbase = td_smem_u32(ring + slot * TD_SLOT);
                        const uint64_t dhi = td_make_desc(sbase, lbo, 128u), dlo = td_make_desc(sbase + 2u * lbo, lbo, 128u);
                        const uint32_t ah = tcol + TD_ST + 16 * sb, al = ah + 8;
                        const uint32_t dst = tcol + (uint32_t)(dcol & 0xFFFF);
                        td_mma_ts(dst, ah, dhi, idesc, acc);
                        td_mma_ts(dst, al, dhi, idesc, 1u);
                        td_mma_ts(dst, ah, dlo, idesc, 1u);
                        td_commit(&bar_done[sb]);
                        td_commit(&bar_empty[slot]);
                    }
                }
            }
        }
    } else {
        // ===================== the tile =====================
        const uint32_t tcol = tmem_base_s;
        const uint32_t trow = tcol + ((uint32_t)(32 * warp) << 16);
        float* vecp = a.vec + (blockIdx.x * 128 + tid);
        uint8_t* bitp = a.bits + (blockIdx.x * 128 + tid);
        const size_t VS = (size_t)a.stride;
        float2* msp = ms + tid;
        const int oV = 0, oU = (K + 1) * D, oSHP = (2 * K + 1) * D, oSHQ = (3 * K + 1) * D, oG = (4 * K + 1) * D;
        unsigned e = 0;        // pushes staged so far (identical in every thread, the producer and the issuer)
        unsigned waited = 0;   // pushes [0, waited) are known to be complete
        uint32_t bph = 0;      // parity of the block barrier
        int resident = -1;

        // pushes complete in issue order; every thread observes every phase of the bar_done barriers in order
        auto wait_upto = [&](unsigned n) {
            if (waited < n) {
                do {
                    td_mbar_wait(&bar_done[waited % TD_NB], (waited / TD_NB) & 1u);
                    ++waited;
                } while (waited < n);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
        };
        // stage 8 values as TF32 (hi, lo) halves into the next staging buffer and hand the push to the issuer
        auto stage_push = [&](const float (&v)[8]) {
            const unsigned sb = e % TD_NB;
            if (e >= TD_NB) wait_upto(e - TD_NB + 1);   // the push that used this buffer before
            float hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                hi[q] = td_hi(v[q]);
                lo[q] = v[q] - hi[q];
            }
            td_st8(trow + TD_ST + 16 * sb, hi);
            td_st8(trow + TD_ST + 16 * sb + 8, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(td_smem_u32(&bar_staged[sb])) : "memory");
            ++e;
        };

        for (int round = 0; round < rounds; ++round) {
            const int tile = round * (int)gridDim.x + (int)blockIdx.x;
            if (tile >= a.ntiles) break;
            const long long p = (long long)tile * 128 + tid;
            const bool live = p < a.npairs;
            // pairs are walked ROW-major (the S samples of an observation row sit in neighbouring lanes): the context projections
            // A[k,u,b] of a row are then read once, by one warp, instead of once per sample by tiles of different rounds
            // (4.6x fewer DRAM reads of A_p / A_q at pyloric size); pidx is the canonical index s * B + b of z / diff / theta
            const int b = live ? (int)(p / a.S) : 0;
            const long long pidx = live ? (long long)(p - (long long)b * a.S) * a.B + b : 0;
            float logp = 0.f, logq = 0.f;
            {
                float ssq = 0.f;
                for (int d = 0; d < D; ++d) {
                    const float zv = live ? a.z[(size_t)pidx * D + d] : 0.f;
                    vecp[(size_t)(oV + (MODE == TC_LOGPROB ? K * D : 0) + d) * VS] = zv;
                    ssq = fmaf(zv, zv, ssq);
                }
                logp = MODE == TC_LOGPROB ? 0.f : -0.5f * ssq - (float)D * HALF_LOG_2PI;
            }
            for (int ph = ph0; ph < nphase; ++ph) {
                const int kind = ph / K, idx = ph - kind * K;
                const int k = (kind == 0 || kind == 2) ? idx : K - 1 - idx;
                const bool seq = (kind == 0 || kind == 3);
                if (resident != k) {   // resident block of transform k
                    asm volatile("bar.sync 1, 128;" ::: "memory");   // every thread is done with the previous block
                    if (tid == 0) {
                        const uint32_t bytes = (uint32_t)a.blk.words * 4u;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(td_smem_u32(&bar_blk)), "r"(bytes) : "memory");
                        const char* srcp = reinterpret_cast<const char*>(a.blocks + (size_t)k * a.blk.words);
                        for (uint32_t off = 0; off < bytes; off += 32768u) {
                            const uint32_t n = min(bytes - off, 32768u);
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"(td_smem_u32(Wb) + off), "l"(srcp + off), "r"(n), "r"(td_smem_u32(&bar_blk)) : "memory");
                        }
                    }
                    td_mbar_wait(&bar_blk, bph);
                    bph ^= 1u;
                    resident = k;
                }
                const TdBlk& bk = a.blk;
                const float* W0t = Wb + bk.w0t;
                const float* b1 = Wb + bk.b1;
                const float* b2 = Wb + bk.b2;
                const float* bo = Wb + bk.bo;
                const float* WoT = Wb + bk.wot;
                const float* dg01 = Wb + bk.dg01;                                     // [D][8 out][8 in]
                const float* dg12 = Wb + bk.dg12;
                const int* perm = Wi + bk.perm;
                const int* permy = Wi + bk.permy;
                const int* pstart = Wi + bk.pstart;                                   // [3][D + 1]
                const TdPiece* pieces = reinterpret_cast<const TdPiece*>(Wi + bk.pieces);   // [3][TD_MAXP]
                uint8_t* bits = bitp + (size_t)((seq ? k : K + k) * 3 * TD_MAXP) * VS;  // [layer][piece]
                const int src = (k == K - 1) ? oV + K * D : oU + (k + 1) * D;
                const int* ps0 = pstart;
                const int* ps1 = pstart + (D + 1);
                const int* ps2 = pstart + 2 * (D + 1);

                if (kind < 2) {
                    // ===================== forward pass of transform k (step-structured; seq: y_j produced by step j) =====================
                    const float* Abase = (seq ? a.A_p : a.A_q) + (size_t)k * H0 * a.B + b;
                    float y[DP];
#pragma unroll
                    for (int jj = 0; jj < DP; ++jj) y[jj] = (!seq && jj < D) ? vecp[(size_t)(src + permy[jj]) * VS] : 0.f;
                    float an[8];
                    auto load_A = [&](const TdPiece& c) {
                        const float* ap = Abase + (size_t)c.r0 * a.B;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            an[i] = i < c.nr ? __ldg(ap) : 0.f;
                            ap += a.B;
                        }
                    };
                    unsigned markp0 = e, markp1 = e, markp2 = e;
                    float hv0[8], hv1[8], hv2[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) hv0[i] = hv1[i] = hv2[i] = 0.f;
                    load_A(pieces[0]);
                    float zn = seq ? vecp[(size_t)(oV + k * D + perm[0]) * VS] : 0.f;
                    for (int j = 0; j < D; ++j) {
                        const float zj = zn;
                        if (seq && j + 1 < D) zn = vecp[(size_t)(oV + k * D + perm[j + 1]) * VS];
                        // ---- layer 0, group j: h = relu(A[u] + W0t[u] . y) (only positions < j have non-zero weights)
                        for (int pi = ps0[j]; pi < ps0[j + 1]; ++pi) {
                            const TdPiece c = pieces[pi];
                            float pre[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) pre[i] = an[i];
                            if (pi + 1 < ps0[D]) load_A(pieces[pi + 1]);
#pragma unroll
                            for (int q4 = 0; q4 < DP / 4; ++q4) {
                                if (4 * q4 < j) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float4 w = *reinterpret_cast<const float4*>(W0t + (size_t)(c.r0 + i) * DP + 4 * q4);
                                        pre[i] = fmaf(w.x, y[4 * q4], pre[i]);
                                        pre[i] = fmaf(w.y, y[4 * q4 + 1], pre[i]);
                                        pre[i] = fmaf(w.z, y[4 * q4 + 2], pre[i]);
                                        pre[i] = fmaf(w.w, y[4 * q4 + 3], pre[i]);
                                    }
                                }
                            }
                            uint32_t cm = 0u;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const bool on = i < c.nr && pre[i] > 0.f;
                                hv0[i] = on ? pre[i] : 0.f;
                                cm |= on ? (1u << i) : 0u;
                            }
                            if (GRADM) bits[(size_t)pi * VS] = (uint8_t)cm;
                            if (c.flags & TD_F_FWD) stage_push(hv0);
                        }
                        const unsigned mark0 = e;
                        // ---- layer 1, group j: accumulated pushes of the groups < j (+ own group: FMA block or waited push)
                        for (int pi = ps1[j]; pi < ps1[j + 1]; ++pi) {
                            const TdPiece c = pieces[TD_MAXP + pi];
                            float acc[8];
                            if (c.flags & TD_F_DGP) {
                                wait_upto(markp0);
                                if (j > 0) {
                                    td_ld8(trow + TD_R0 + (uint32_t)c.r0, acc);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                                }
                                const float4* dq = reinterpret_cast<const float4*>(dg01 + (size_t)j * 64);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 w0 = dq[2 * i], w1 = dq[2 * i + 1];
                                    acc[i] = fmaf(w0.x, hv0[0], acc[i]);
                                    acc[i] = fmaf(w0.y, hv0[1], acc[i]);
                                    acc[i] = fmaf(w0.z, hv0[2], acc[i]);
                                    acc[i] = fmaf(w0.w, hv0[3], acc[i]);
                                    acc[i] = fmaf(w1.x, hv0[4], acc[i]);
                                    acc[i] = fmaf(w1.y, hv0[5], acc[i]);
                                    acc[i] = fmaf(w1.z, hv0[6], acc[i]);
                                    acc[i] = fmaf(w1.w, hv0[7], acc[i]);
                                }
                            } else {
                                wait_upto(mark0);
                                td_ld8(trow + TD_R0 + (uint32_t)c.r0, acc);
                            }
                            uint32_t cm = 0u;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float val = acc[i] + b1[c.r0 + i];
                                const bool on = i < c.nr && val > 0.f;
                                hv1[i] = on ? val : 0.f;
                                cm |= on ? (1u << i) : 0u;
                            }
                            if (GRADM) bits[(size_t)(TD_MAXP + pi) * VS] = (uint8_t)cm;
                            if (c.flags & TD_F_FWD) stage_push(hv1);
                        }
                        const unsigned mark1 = e;
                        // ---- layer 2, group j; its part of (m_j, s_j) by FMA, the later positions by a push
                        float2 msj = make_float2(bo[2 * j], bo[2 * j + 1]);
                        for (int pi = ps2[j]; pi < ps2[j + 1]; ++pi) {
                            const TdPiece c = pieces[2 * TD_MAXP + pi];
                            float acc[8];
                            if (c.flags & TD_F_DGP) {
                                wait_upto(markp1);
                                if (j > 0) {
                                    td_ld8(trow + TD_R1 + (uint32_t)c.r0, acc);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                                }
                                const float4* dq = reinterpret_cast<const float4*>(dg12 + (size_t)j * 64);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 w0 = dq[2 * i], w1 = dq[2 * i + 1];
                                    acc[i] = fmaf(w0.x, hv1[0], acc[i]);
                                    acc[i] = fmaf(w0.y, hv1[1], acc[i]);
                                    acc[i] = fmaf(w0.z, hv1[2], acc[i]);
                                    acc[i] = fmaf(w0.w, hv1[3], acc[i]);
                                    acc[i] = fmaf(w1.x, hv1[4], acc[i]);
                                    acc[i] = fmaf(w1.y, hv1[5], acc[i]);
                                    acc[i] = fmaf(w1.z, hv1[6], acc[i]);
                                    acc[i] = fmaf(w1.w, hv1[7], acc[i]);
                                }
                            } else {
                                wait_upto(mark1);
                                td_ld8(trow + TD_R1 + (uint32_t)c.r0, acc);
                            }
                            uint32_t cm = 0u;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float val = acc[i] + b2[c.r0 + i];
                                const bool on = i < c.nr && val > 0.f;
                                hv2[i] = on ? val : 0.f;
                                cm |= on ? (1u << i) : 0u;
                            }
                            if (GRADM) bits[(size_t)(2 * TD_MAXP + pi) * VS] = (uint8_t)cm;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float2 wj = *reinterpret_cast<const float2*>(WoT + (size_t)(c.r0 + i) * 2 * DP + 2 * j);
                                msj = td_ffma2(wj, make_float2(hv2[i], hv2[i]), msj);
                            }
                            if (c.flags & TD_F_FWD) stage_push(hv2);
                        }
                        const unsigned mark2 = e;
                        // ---- position j: contributions of the groups < j come from the accumulator
                        if (j > 0) {
                            wait_upto(markp2);
                            float m0, s0;
                            td_ld2(trow + TD_R2 + 2 * j, m0, s0);
                            msj.x += m0;
                            msj.y += s0;
                        }
                        const float sh = fminf(fmaxf(msj.y, g.lo), g.hi);
                        if (seq) {
                            const float yj = (zj - msj.x) * expf(-sh);
#pragma unroll
                            for (int jj = 0; jj < DP; ++jj) y[jj] = jj == j ? yj : y[jj];
                            logp += sh;
                            vecp[(size_t)(oSHP + k * D + j) * VS] = sh;
                            vecp[(size_t)(oV + (k + 1) * D + permy[j]) * VS] = yj;
                        } else {   // u_k = exp(clamp(s)) * u_{k+1} + m
                            float yj = 0.f;
#pragma unroll
                            for (int q = 0; q < DP; ++q) yj = q == j ? y[q] : yj;
                            vecp[(size_t)(oSHQ + k * D + j) * VS] = sh;
                            logq += sh;
                            vecp[(size_t)(oU + k * D + perm[j]) * VS] = fmaf(expf(sh), yj, msj.x);
                        }
                        markp0 = mark0;
                        markp1 = mark1;
                        markp2 = mark2;
                    }
                    wait_upto(e);   // the next pass overwrites the accumulators
                    if (!seq && k == 0) {  // end of the density chain: base log-density, KL term, reverse-mode seed
                        float ssq = 0.f;
                        for (int d = 0; d < D; ++d) {
                            const float u = vecp[(size_t)(oU + d) * VS];
                            ssq = fmaf(u, u, ssq);
                            if (GRADM) vecp[(size_t)(oG + d) * VS] = a.w * u;  // d(-w log q)/d u_0
                        }
                        logq += -0.5f * ssq - (float)D * HALF_LOG_2PI;
                        if (a.diff && live) a.diff[pidx] = MODE == TC_LOGPROB ? logq : logp - logq;
                    }
                } else {
                    // ===================== reverse pass of transform k (steps in descending order) =====================
                    const float w = a.w;
                    float yb[DP];
                    if (seq) {
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) yb[jj] = jj < D ? vecp[(size_t)(oG + permy[jj]) * VS] : 0.f;
                    } else {
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj) {
                            if (jj < D) {
                                const float ub = vecp[(size_t)(oG + perm[jj]) * VS];
                                const float es = expf(vecp[(size_t)(oSHQ + k * D + jj) * VS]);
                                const float yin = vecp[(size_t)(src + permy[jj]) * VS];
                                msp[jj * 128] = make_float2(ub, fmaf(ub * es, yin, -w));
                                yb[jj] = ub * es;
                            } else {
                                yb[jj] = 0.f;
                            }
                        }
                    }
                    float* gAbase = a.gA + (size_t)k * H0 * a.B + b;
                    const bool emit = (seq || MODE == TC_FKL) && live;
                    unsigned markp2 = e, markp1 = e;
                    float gs2[8], gs1[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) gs2[i] = gs1[i] = 0.f;
                    float yvn = 0.f, shn = 0.f;
                    if (seq) {
                        yvn = vecp[(size_t)(oV + (k + 1) * D + permy[D - 1]) * VS];
                        shn = vecp[(size_t)(oSHP + k * D + D - 1) * VS];
                    }
                    for (int j = D - 1; j >= 0; --j) {
                        if (seq) {
                            const float yvj = yvn, shj = shn;
                            if (j > 0) {
                                yvn = vecp[(size_t)(oV + (k + 1) * D + permy[j - 1]) * VS];
                                shn = vecp[(size_t)(oSHP + k * D + j - 1) * VS];
                            }
                            float ybj = 0.f;
#pragma unroll
                            for (int jj = 0; jj < DP; ++jj) ybj = jj == j ? yb[jj] : ybj;
                            const float zbj = ybj * expf(-shj);
                            msp[j * 128] = make_float2(-zbj, fmaf(-ybj, yvj, w));
                            vecp[(size_t)(oG + perm[j]) * VS] = zbj;   // the old oG values are all in yb already
                        }
                        // ---- layer 2, group j: g = relu'(.) * sum_{jj >= j} WoT[v][jj] . (mbar, sbar)_jj
                        for (int pi = ps2[j + 1] - 1; pi >= ps2[j]; --pi) {
                            const TdPiece c = pieces[2 * TD_MAXP + pi];
                            const uint32_t cm = bits[(size_t)(2 * TD_MAXP + pi) * VS];
                            float2 acc[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
                            for (int jj = j; jj < D; ++jj) {
                                const float2 mv = msp[jj * 128];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 wj = *reinterpret_cast<const float2*>(WoT + (size_t)(c.r0 + i) * 2 * DP + 2 * jj);
                                    acc[i] = td_ffma2(wj, mv, acc[i]);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) gs2[i] = ((cm >> i) & 1u) ? acc[i].x + acc[i].y : 0.f;
                            if (c.flags & TD_F_REV) stage_push(gs2);
                        }
                        const unsigned mark2 = e;
                        // ---- layer 1, group j: cotangents pushed by the layer-2 groups > j (+ own group: FMA block or waited push)
                        for (int pi = ps1[j + 1] - 1; pi >= ps1[j]; --pi) {
                            const TdPiece c = pieces[TD_MAXP + pi];
                            const uint32_t cm = bits[(size_t)(TD_MAXP + pi) * VS];
                            float acc[8];
                            if (c.flags & TD_F_DGN) {
                                wait_upto(markp2);
                                if (j < D - 1) {
                                    td_ld8(trow + TD_R0 + (uint32_t)c.r0, acc);
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                                }
                                const float4* dq = reinterpret_cast<const float4*>(dg12 + (size_t)j * 64);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 w0 = dq[2 * i], w1 = dq[2 * i + 1];
                                    acc[0] = fmaf(w0.x, gs2[i], acc[0]);
                                    acc[1] = fmaf(w0.y, gs2[i], acc[1]);
                                    acc[2] = fmaf(w0.z, gs2[i], acc[2]);
                                    acc[3] = fmaf(w0.w, gs2[i], acc[3]);
                                    acc[4] = fmaf(w1.x, gs2[i], acc[4]);
                                    acc[5] = fmaf(w1.y, gs2[i], acc[5]);
                                    acc[6] = fmaf(w1.z, gs2[i], acc[6]);
                                    acc[7] = fmaf(w1.w, gs2[i], acc[7]);
                                }
                            } else {
                                wait_upto(mark2);
                                td_ld8(trow + TD_R0 + (uint32_t)c.r0, acc);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) gs1[i] = ((cm >> i) & 1u) ? acc[i] : 0.f;
                            if (c.flags & TD_F_REV) stage_push(gs1);
                        }
                        const unsigned mark1 = e;
                        // ---- layer 0, group j: d/dA, and d/dy of the positions < j
                        for (int pi = ps0[j]; pi < ps0[j + 1]; ++pi) {
                            const TdPiece c = pieces[pi];
                            const uint32_t cm = bits[(size_t)pi * VS];
                            float acc[8], gv[8];
                            if (c.flags & TD_F_DGN) {
                                wait_upto(markp1);
                                if (j < D - 1) {
                                    td_ld8(trow + TD_R1 + (uint32_t)c.r0, acc);
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                                }
                                const float4* dq = reinterpret_cast<const float4*>(dg01 + (size_t)j * 64);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 w0 = dq[2 * i], w1 = dq[2 * i + 1];
                                    acc[0] = fmaf(w0.x, gs1[i], acc[0]);
                                    acc[1] = fmaf(w0.y, gs1[i], acc[1]);
                                    acc[2] = fmaf(w0.z, gs1[i], acc[2]);
                                    acc[3] = fmaf(w0.w, gs1[i], acc[3]);
                                    acc[4] = fmaf(w1.x, gs1[i], acc[4]);
                                    acc[5] = fmaf(w1.y, gs1[i], acc[5]);
                                    acc[6] = fmaf(w1.z, gs1[i], acc[6]);
                                    acc[7] = fmaf(w1.w, gs1[i], acc[7]);
                                }
                            } else {
                                wait_upto(mark1);
                                td_ld8(trow + TD_R1 + (uint32_t)c.r0, acc);
                            }
                            float* gAp = gAbase + (size_t)c.r0 * a.B;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const bool on = (cm >> i) & 1u;
                                gv[i] = on ? acc[i] : 0.f;
                                if (emit && on) atomicAdd(gAp + (size_t)i * a.B, gv[i]);
                            }
#pragma unroll
                            for (int q4 = 0; q4 < DP / 4; ++q4) {
                                if (4 * q4 < j) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float4 w4 = *reinterpret_cast<const float4*>(W0t + (size_t)(c.r0 + i) * DP + 4 * q4);
                                        yb[4 * q4] = fmaf(w4.x, gv[i], yb[4 * q4]);
                                        yb[4 * q4 + 1] = fmaf(w4.y, gv[i], yb[4 * q4 + 1]);
                                        yb[4 * q4 + 2] = fmaf(w4.z, gv[i], yb[4 * q4 + 2]);
                                        yb[4 * q4 + 3] = fmaf(w4.w, gv[i], yb[4 * q4 + 3]);
                                    }
                                }
                            }
                        }
                        markp2 = mark2;
                        markp1 = mark1;
                    }
                    wait_upto(e);
                    if (!seq) {
#pragma unroll
                        for (int jj = 0; jj < DP; ++jj)
                            if (jj < D) vecp[(size_t)(oG + permy[jj]) * VS] = yb[jj];
                    }
                }
            }
            if (MODE == TC_RSAMPLE && live) {
                for (int d = 0; d < D; ++d) a.theta_out[(size_t)pidx * D + d] = vecp[(size_t)(oV + K * D + d) * VS];
                if (a.diff) a.diff[pidx] = logp;
            }
        }
        wait_upto(e);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s) : "memory");
}

// ------------------------------------------------------------------------------------------------- image packing
struct TdStrip { int dir, l, r0, nr, n0, N, off16, zcut; };
// dir 0: forward (K side = layer l units, N side = layer l+1 units / outputs); columns c < zcut are zeroed (own-degree FMA block)
// dir 1: reverse (K side = layer l units, N side = layer l-1 units); columns c >= zcut are zeroed
// one block per strip: [2 k-chunks][N][4] hi, then lo; element (n, kk) = weight between K-side unit r0 + kk and N-side column n0 + n
__global__ void k_tcd_pack_strips(const MafImg g, int k, const TdStrip* __restrict__ strips, char* __restrict__ img) {
    const TdStrip s = strips[blockIdx.x];
    const TImg t = g.t[k];
    const float* f = g.f;
    const int D = g.D;
    float* hi = reinterpret_cast<float*>(img + (size_t)s.off16 * 16);
    float* lo = hi + 2 * s.N * 4;
    for (int o = threadIdx.x; o < 2 * s.N * 4; o += blockDim.x) {
        const int kc = o / (s.N * 4), n = (o >> 2) % s.N, r = o & 3, kk = 4 * kc + r;
        float w = 0.f;
        if (kk < s.nr) {
            const int u = s.r0 + kk, c = s.n0 + n;
            if (s.dir == 0 && s.l < 2) {           // W[l+1] is [v][u]: row = unit of layer l+1 (column c), col = unit of layer l
                if (c < g.H[s.l + 1] && c >= s.zcut) w = f[t.W[s.l + 1] + (size_t)c * t.ld[s.l + 1] + u];
            } else if (s.dir == 0) {               // output layer: column c = 2 * position + half; Wo rows: position j = mean, D + j = log-scale
                const int jj = c >> 1, half = c & 1;
                if (jj < D && c >= s.zcut) w = f[t.Wo + (size_t)(half * D + jj) * t.ldo + u];
            } else {                               // reverse: K side = unit of layer l (row of W[l]), N side = unit of layer l-1
                if (c < g.H[s.l - 1] && c < s.zcut) w = f[t.W[s.l] + (size_t)u * t.ld[s.l] + c];
            }
        }
        const float h = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
        hi[o] = h;
        lo[o] = w - h;
    }
}
// float part of the resident block of transform k (integer tables come from the host template, already copied into `out`)
__global__ void k_tcd_pack_block(const MafImg g, int k, TdBlk bk, int DP, float* __restrict__ out) {
    const TImg t = g.t[k];
    const float* f = g.f;
    const int D = g.D;
    const int H0 = g.H[0], H1 = g.H[1], H2 = g.H[2];
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int o = t0; o < H0 * DP; o += stride) {
        const int u = o / DP, jj = o % DP;
        out[bk.w0t + o] = jj < t.ldt ? f[t.W0t + (size_t)u * t.ldt + jj] : 0.f;
    }
    for (int o = t0; o < H1; o += stride) out[bk.b1 + o] = f[t.b[1] + o];
    for (int o = t0; o < H2; o += stride) out[bk.b2 + o] = f[t.b[2] + o];
    for (int o = t0; o < 2 * DP; o += stride) out[bk.bo + o] = o < 2 * t.ldt ? f[t.boF + o] : 0.f;
    for (int o = t0; o < H2 * 2 * DP; o += stride) {
        const int v = o / (2 * DP), c = o % (2 * DP);
        out[bk.wot + o] = c < 2 * t.ldt ? f[t.WoT + (size_t)v * 2 * t.ldt + c] : 0.f;
    }
    const int* dgi = reinterpret_cast<const int*>(out) + bk.dgi;
    for (int o = t0; o < 2 * D * 64; o += stride) {
        const int lp = o / (D * 64), j = (o >> 6) % D, i = (o >> 3) & 7, q = o & 7;
        const int* tab = dgi + (lp * D + j) * 4;
        float w = 0.f;
        if (i < tab[1] && q < tab[3]) w = f[t.W[lp + 1] + (size_t)(tab[0] + i) * t.ld[lp + 1] + tab[2] + q];
        out[(lp ? bk.dg12 : bk.dg01) + (o - lp * D * 64)] = w;
    }
}

struct TdState {
    bool built = false, eligible = false;
    int DP = 32;
    TdBlk blk;
    float* blocks = nullptr;        // [template][packed]: K * blk.words each
    size_t blocks_words = 0;
    char* strips = nullptr;
    size_t strips_bytes = 0;
    TdStrip* strip_tbl = nullptr;   // device copy, per transform contiguous
    std::vector<int> strip_off, strip_cnt;
    TdPush* push = nullptr;
    int ps_off[RABI_MAXK][2], ps_cnt[RABI_MAXK][2];
    unsigned long long packed_version = ~0ull;
    float* vec = nullptr;
    uint8_t* bits = nullptr;   // inside the allocation of vec
    size_t scratch_bytes = 0, l2_window = 0;
    size_t scratch_pairs = 0;
};

static inline int td_r4(int x) { return (x + 3) & ~3; }
static inline int td_r8(int x) { return (x + 7) & ~7; }

static void td_free(TdState* s) {
    if (s->blocks) cudaFree(s->blocks);
    if (s->strips) cudaFree(s->strips);
    if (s->strip_tbl) cudaFree(s->strip_tbl);
    if (s->push) cudaFree(s->push);
    if (s->vec) cudaFree(s->vec);
    s->blocks = nullptr;
    s->strips = nullptr;
    s->strip_tbl = nullptr;
    s->push = nullptr;
    s->vec = nullptr;
    s->bits = nullptr;
    s->scratch_pairs = 0;
}

static int td_build(rabi_maf* h, TdState* s) {
    s->built = true;
    s->eligible = false;
    const int K = h->K, D = h->D;
    if (h->L != 3 || D > 32 || D < 2) return 0;
    for (int l = 0; l < 3; ++l)
        if (h->H[l] > TD_MAXH) return 0;
    const std::vector<int>& M = h->meta_host;
    const MafImg& g = h->img;
    const int DP = td_r4(D);
    s->DP = DP;
    const int H0 = h->H[0], H1 = h->H[1], H2 = h->H[2];
    const bool use_diag = rabi_env_int("RABI_TCD_DIAG", 1) != 0;
    // resident block layout
    {
        TdBlk& b = s->blk;
        int o = 0;
        auto take = [&](int n) {
            int at = o;
            o += (n + 3) & ~3;
            return at;
        };
        // 8 zero rows of slack behind every per-unit array: a piece always touches 8 unit slots
        b.w0t = take((H0 + 8) * DP);
        b.b1 = take(TD_MAXH + 8);
        b.b2 = take(TD_MAXH + 8);
        b.bo = take(2 * DP);
        b.wot = take((H2 + 8) * 2 * DP);
        b.dg01 = take(D * 64);
        b.dg12 = take(D * 64);
        b.perm = take(DP);
        b.permy = take(DP);
        b.np = take(4);
        b.pstart = take(3 * (D + 1));
        b.pieces = take(3 * TD_MAXP * 4);
        b.dgi = take(2 * D * 4);
        b.words = (o + 31) & ~31;
    }
    std::vector<float> blocks((size_t)K * s->blk.words, 0.f);
    std::vector<TdStrip> strips;
    std::vector<TdPush> pushes;
    s->strip_off.assign(K, 0);
    s->strip_cnt.assign(K, 0);
    size_t img_bytes = 0;
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        const int* grp[3] = {&M[t.grp[0]], &M[t.grp[1]], &M[t.grp[2]]};
        const int Hl[3] = {H0, H1, H2};
        int* w = reinterpret_cast<int*>(blocks.data() + (size_t)k * s->blk.words);
        for (int j = 0; j < D; ++j) {
            w[s->blk.perm + j] = M[t.perm + j];
            w[s->blk.permy + j] = M[t.permy + j];
        }
        std::vector<TdPiece> pc[3];
        std::vector<int> ps[3];
        for (int l = 0; l < 3; ++l) {
            ps[l].assign(D + 1, 0);
            for (int j = 0; j < D; ++j) {
                ps[l][j] = (int)pc[l].size();
                const int len = grp[l][j + 1] - grp[l][j];
                if (len <= 0) return 0;   // an empty degree group: not a shape this path tiles
                for (int o = 0; o < len; o += 8) {
                    TdPiece c;
                    c.r0 = grp[l][j] + o;
                    c.nr = std::min(8, len - o);
                    c.j = j;
                    c.flags = 0;
                    pc[l].push_back(c);
                }
            }
            ps[l][D] = (int)pc[l].size();
            if ((int)pc[l].size() > TD_MAXP) return 0;
        }
        // own-degree blocks: when group j is a single piece in layer l and in layer l + 1, their product is an FMA block on
        // the values the thread still holds, and the strips leave it out -- no push has to complete inside a step
        bool dg[2][32];
        for (int lp = 0; lp < 2; ++lp)
            for (int j = 0; j < D; ++j) {
                dg[lp][j] = use_diag && ps[lp][j + 1] - ps[lp][j] == 1 && ps[lp + 1][j + 1] - ps[lp + 1][j] == 1;
                int* tab = w + s->blk.dgi + (lp * D + j) * 4;
                tab[0] = grp[lp + 1][j];
                tab[1] = dg[lp][j] ? grp[lp + 1][j + 1] - grp[lp + 1][j] : 0;
                tab[2] = grp[lp][j];
                tab[3] = dg[lp][j] ? grp[lp][j + 1] - grp[lp][j] : 0;
                if (dg[lp][j]) {
                    pc[lp][ps[lp][j]].flags |= TD_F_DGN;
                    pc[lp + 1][ps[lp + 1][j]].flags |= TD_F_DGP;
                }
            }
        // strips + push streams.  Forward stream: per step j the pieces of layer 0, 1, 2; reverse: j descending, layer 2, 1.
        s->strip_off[k] = (int)strips.size();
        auto add_strip = [&](int dir, int l, const TdPiece& c, int n0, int N, int zcut) {
            TdStrip st;
            st.dir = dir;
            st.l = l;
            st.r0 = c.r0;
            st.nr = c.nr;
            st.n0 = n0;
            st.N = N;
            st.off16 = (int)(img_bytes / 16);
            st.zcut = zcut;
            img_bytes += (size_t)2 * 2 * N * 16;
            img_bytes = (img_bytes + 127) & ~(size_t)127;
            strips.push_back(st);
            return (int)strips.size() - 1;
        };
        const int region[3] = {TD_R0, TD_R1, TD_R2};
        int bad = 0;
        auto emit = [&](int si, int reg, bool first) {
            const TdStrip& st = strips[si];
            TdPush p;
            p.off16 = st.off16;
            p.bytes = 2 * 2 * st.N * 16;
            p.dcol = (reg + st.n0) | (first ? (1 << 16) : 0);
            p.N = st.N;
            if (p.bytes > TD_SLOT || st.N < 8 || st.N > 256) bad = 1;
            pushes.push_back(p);
        };
        s->ps_off[k][0] = (int)pushes.size();
        bool firstF[3] = {true, true, true};
        for (int j = 0; j < D; ++j)
            for (int l = 0; l < 3; ++l)
                for (int pi = ps[l][j]; pi < ps[l][j + 1]; ++pi) {
                    TdPiece& c = pc[l][pi];
                    int n0, N, zcut, real_end;
                    if (l < 2) {
                        zcut = dg[l][j] ? grp[l + 1][j + 1] : 0;
                        n0 = (dg[l][j] ? grp[l + 1][j + 1] : grp[l + 1][j]) & ~7;
                        N = td_r8(Hl[l + 1]) - n0;
                        real_end = Hl[l + 1];
                    } else {   // own position by FMA
                        zcut = 2 * (j + 1);
                        n0 = zcut & ~7;
                        N = td_r8(2 * D) - n0;
                        real_end = 2 * D;
                    }
                    if (N <= 0 || std::max(zcut, n0) >= real_end) continue;   // nothing left to push into
                    c.flags |= TD_F_FWD;
                    emit(add_strip(0, l, c, n0, N, zcut), region[l], firstF[l]);
                    firstF[l] = false;
                }
        s->ps_cnt[k][0] = (int)pushes.size() - s->ps_off[k][0];
        s->ps_off[k][1] = (int)pushes.size();
        bool firstR[3] = {true, true, true};
        for (int j = D - 1; j >= 0; --j)
            for (int l = 2; l >= 1; --l)
                for (int pi = ps[l][j + 1] - 1; pi >= ps[l][j]; --pi) {
                    TdPiece& c = pc[l][pi];
                    const bool d = dg[l - 1][j];
                    const int zcut = d ? grp[l - 1][j] : (1 << 30);
                    const int N = td_r8(d ? grp[l - 1][j] : grp[l - 1][j + 1]);
                    if (N <= 0) continue;
                    c.flags |= TD_F_REV;
                    emit(add_strip(1, l, c, 0, N, zcut), l == 2 ? TD_R0 : TD_R1, firstR[l]);
                    firstR[l] = false;
                }
        s->ps_cnt[k][1] = (int)pushes.size() - s->ps_off[k][1];
        s->strip_cnt[k] = (int)strips.size() - s->strip_off[k];
        if (bad) return 0;
        for (int l = 0; l < 3; ++l) {
            w[s->blk.np + l] = (int)pc[l].size();
            for (int j = 0; j <= D; ++j) w[s->blk.pstart + l * (D + 1) + j] = ps[l][j];
            memcpy(w + s->blk.pieces + l * TD_MAXP * 4, pc[l].data(), pc[l].size() * sizeof(TdPiece));
        }
    }
    td_free(s);
    s->blocks_words = blocks.size();
    RABI_CU_OK(h, cudaMalloc(&s->blocks, 2 * blocks.size() * sizeof(float)));
    RABI_CU_OK(h, cudaMemcpy(s->blocks, blocks.data(), blocks.size() * sizeof(float), cudaMemcpyHostToDevice));
    s->strips_bytes = img_bytes;
    RABI_CU_OK(h, cudaMalloc(&s->strips, img_bytes));
    RABI_CU_OK(h, cudaMalloc(&s->strip_tbl, strips.size() * sizeof(TdStrip)));
    RABI_CU_OK(h, cudaMemcpy(s->strip_tbl, strips.data(), strips.size() * sizeof(TdStrip), cudaMemcpyHostToDevice));
    RABI_CU_OK(h, cudaMalloc(&s->push, pushes.size() * sizeof(TdPush)));
    RABI_CU_OK(h, cudaMemcpy(s->push, pushes.data(), pushes.size() * sizeof(TdPush), cudaMemcpyHostToDevice));
    s->packed_version = ~0ull;
    s->eligible = true;
    return 0;
}

template <int MODE, int DP>
static int td_launch_inst(rabi_maf* h, const TdArgs& a, int grid, size_t smem, cudaStream_t st) {
    RABI_CU_OK(h, cudaFuncSetAttribute(k_tcd<MODE, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = MODE == TC_GRAD || MODE == TC_FKL;
    if (prof && rabi_prof_begin(h, st)) return 1;
    k_tcd<MODE, DP><<<grid, TD_THREADS, smem, st>>>(h->img, a);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rabi_fail(h, "k_tcd launch failed: %s", cudaGetErrorString(e));
    if (prof && rabi_prof_end(h, st)) return 1;
    return 0;
}

template <int DP>
static int td_launch_mode(rabi_maf* h, const TdArgs& a, int mode, int grid, size_t smem, cudaStream_t st) {
    switch (mode) {
        case TC_FWD: return td_launch_inst<TC_FWD, DP>(h, a, grid, smem, st);
        case TC_GRAD: return td_launch_inst<TC_GRAD, DP>(h, a, grid, smem, st);
        case TC_FKL: return td_launch_inst<TC_FKL, DP>(h, a, grid, smem, st);
        case TC_RSAMPLE: return td_launch_inst<TC_RSAMPLE, DP>(h, a, grid, smem, st);
        case TC_LOGPROB: return td_launch_inst<TC_LOGPROB, DP>(h, a, grid, smem, st);
    }
    return 2;
}

int rabi_tcd_launch(rabi_maf* h, const TcCall& c, int mode, cudaStream_t st) {
    const long long npairs = (long long)c.S * c.B;
    if (npairs <= 0) return 0;
    if (!rabi_env_int("RABI_TC", 1) || !rabi_env_int("RABI_TCD", 1)) return 2;
    TdState* s = static_cast<TdState*>(h->tcd);
    if (!s) h->tcd = s = new TdState();
    if (!s->built && td_build(h, s)) return 1;
    if (!s->eligible) return 2;
    const int K = h->K, D = h->D;
    if (s->packed_version != h->weights_version) {
        RABI_CU_OK(h, cudaMemcpyAsync(s->blocks + s->blocks_words, s->blocks, s->blocks_words * sizeof(float), cudaMemcpyDeviceToDevice, st));
        for (int k = 0; k < K; ++k) {
            k_tcd_pack_block<<<16, 256, 0, st>>>(h->img, k, s->blk, s->DP, s->blocks + s->blocks_words + (size_t)k * s->blk.words);
            k_tcd_pack_strips<<<s->strip_cnt[k], 256, 0, st>>>(h->img, k, s->strip_tbl + s->strip_off[k], s->strips);
            h->launches += 2;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return rabi_fail(h, "k_tcd_pack launch failed: %s", cudaGetErrorString(e));
        s->packed_version = h->weights_version;
    }
    TdArgs a;
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
    a.blocks = s->blocks + s->blocks_words;
    a.blk = s->blk;
    a.strips = s->strips;
    a.push = s->push;
    memcpy(a.ps_off, s->ps_off, sizeof(a.ps_off));
    memcpy(a.ps_cnt, s->ps_cnt, sizeof(a.ps_cnt));
    const int grid = (int)std::min<long long>(nt, h->sm_count);
    const size_t pairs = (size_t)h->sm_count * 128;
    if (s->scratch_pairs < pairs) {
        if (s->vec) {
            RABI_CU_OK(h, cudaDeviceSynchronize());
            cudaFree(s->vec);
            s->vec = nullptr;
            s->bits = nullptr;
        }
        // ONE allocation [chain vectors | ReLU bytes] so that a single L2 access-policy window covers the whole per-pair scratch
        const size_t vec_bytes = (size_t)(4 * K + 2) * D * pairs * sizeof(float);
        const size_t bits_bytes = (size_t)2 * K * 3 * TD_MAXP * pairs * sizeof(uint8_t);
        RABI_CU_OK(h, cudaMalloc(&s->vec, vec_bytes + bits_bytes));
        s->bits = reinterpret_cast<uint8_t*>(s->vec) + vec_bytes;
        s->scratch_bytes = vec_bytes + bits_bytes;
        s->scratch_pairs = pairs;
        // the scratch is rewritten and re-read by every tile round while A_p / A_q / gA stream through: ask for it to stay in L2
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
        s->l2_window = 0;
        if (max_persist > 0 && max_window > 0 && rabi_env_int("RABI_TCD_L2_PERSIST", 1)) {
            const size_t want = std::min<size_t>(s->scratch_bytes, (size_t)max_persist);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
                s->l2_window = std::min<size_t>(s->scratch_bytes, (size_t)max_window);
            cudaGetLastError();
        }
    }
    a.vec = s->vec;
    a.bits = s->bits;
    a.stride = (int)pairs;
    const size_t smem = (size_t)TD_NSLOT * TD_SLOT + (size_t)s->blk.words * 4 + (size_t)s->DP * 128 * sizeof(float2);
    if (smem > (size_t)h->max_smem_optin) return 2;
    // L2 access-policy window over the per-pair scratch for this launch (persisting hits; everything else streams)
    if (s->l2_window) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = s->vec;
        av.accessPolicyWindow.num_bytes = s->l2_window;
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
    int rc = 2;
    switch (s->DP) {
#ifndef RABI_LEAN
        case 4: rc = td_launch_mode<4>(h, a, mode, grid, smem, st); break;
        case 8: rc = td_launch_mode<8>(h, a, mode, grid, smem, st); break;
        case 12: rc = td_launch_mode<12>(h, a, mode, grid, smem, st); break;
        case 16: rc = td_launch_mode<16>(h, a, mode, grid, smem, st); break;
#endif
        case 32: rc = td_launch_mode<32>(h, a, mode, grid, smem, st); break;
    }
    if (s->l2_window) {   // later kernels on this stream see no window
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.num_bytes = 0;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
    return rc;
}

void rabi_tcd_destroy(rabi_maf* h) {
    TdState* s = static_cast<TdState*>(h->tcd);
    if (!s) return;
    td_free(s);
    delete s;
    h->tcd = nullptr;
}

}  // namespace rabi
