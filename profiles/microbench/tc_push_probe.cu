// Round-2 probe for the tensor-core MADE kernels: the "group push" pattern and its round-trip cost.
//
// One CTA, NT tiles of 128 pairs (one warpgroup each, thread = pair = TMEM lane).  A hidden->hidden product with a
// block-lower-triangular weight matrix is evaluated group by group: the threads ReLU a group of G inputs, split them
// into TF32 (hi, lo) and store them to a TMEM staging area; one elected thread issues the 3xTF32 tcgen05.mma chain that
// pushes the group into the accumulator columns of all later groups (A from TMEM, B from shared memory, K-major
// no-swizzle [k/4][n][4] layout), commits to an mbarrier; the threads wait, read the now-final accumulator columns of
// the next group and go on.  Measures clocks per step and checks the result against an fp64 CPU product.
//
// Also: B delivered by cp.async.bulk + mbarrier complete_tx (bmode 1), A from shared memory (SS form, test "ss"),
// and N % 8 shapes at M = 128 (test "n8").
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* mbar, uint32_t parity, int* err) {
    uint32_t done = 0;
    long long spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
        if (++spins > 4000000LL) { atomicExch(err, 1); return false; }
    }
    return true;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct P {
    int NT, HK, NP, G, bmode, reps;
    const float* h;      // [NT*128][HK]
    const float* Bimg;   // pre-formatted [2][HK/4][NP][4]  (hi, lo)
    float* out;          // [NT*128][NP]
    long long* clk;      // [NT]
    int* err;
};
__host__ __device__ inline int n0_of(int g, int G, int HK, int NP) {
    int NG = HK / G;
    if (g >= NG) return NP;
    return ((g * NP / NG) / 16) * 16;
}

__global__ void __launch_bounds__(384, 1) k_push(P p) {
    extern __shared__ __align__(128) float sm[];
    const int nB = (p.HK / 4) * p.NP * 4;
    float* Bhi = sm;
    float* Blo = sm + nB;
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t mbar[4];   // [0..2] per tile, [3] bulk copy
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.bmode == 0) {
        for (int i = tid; i < 2 * nB; i += blockDim.x) sm[i] = p.Bimg[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    } else {
        if (tid == 0) {
            const uint32_t bytes = 2u * nB * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar[3])), "r"(bytes) : "memory");
            // bulk copies of <= 64 KB each
            uint32_t off = 0;
            while (off < bytes) {
                uint32_t n = min(bytes - off, 32768u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(sm) + off), "l"((const char*)p.Bimg + off), "r"(n), "r"(smem_u32(&mbar[3])) : "memory");
                off += n;
            }
        }
        mbar_wait(&mbar[3], 0, p.err);
    }
    const int TCOLS = 512 / 3 / 8 * 8;                 // 168 columns per tile
    const uint32_t tcol = tbase_s + (uint32_t)(wg * TCOLS);
    const uint32_t DCOL = 0, SH = p.NP, SL = p.NP + p.G;   // NP + 2G <= 168
    const uint32_t trow = tcol + ((uint32_t)(32 * (warp & 3)) << 16);
    const int row = wg * 128 + (tid & 127);
    const int NG = p.HK / p.G;
    uint32_t phase = 0;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < p.reps; ++rep) {
        if (rep == 1 || p.reps == 1) t0 = clock64();
        for (int g = 0; g < NG; ++g) {
            // ---- the group's inputs: relu, split, stage
            for (int c = 0; c < p.G; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(p.h + (size_t)row * p.HK + g * p.G + c);
                float x[4] = {fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)};
                float hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { hi[q] = tf32_hi(x[q]); lo[q] = x[q] - hi[q]; }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(trow + SH + c), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(trow + SL + c), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
            if ((tid & 127) == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int n0 = n0_of(g, p.G, p.HK, p.NP);
                const uint32_t idesc = make_idesc(128, p.NP - n0);
                uint32_t acc = g > 0 ? 1u : 0u;
                for (int t = 0; t < 3; ++t) {     // (A_hi,B_hi) (A_lo,B_hi) (A_hi,B_lo)
                    const float* Bs = t == 2 ? Blo : Bhi;
                    const uint32_t acol = t == 1 ? SL : SH;
                    for (int kk = 0; kk < p.G; kk += 8) {
                        const uint32_t addr = smem_u32(Bs) + (uint32_t)((((g * p.G + kk) / 4) * p.NP + n0) * 16);
                        const uint64_t desc = make_desc(addr, (uint32_t)p.NP * 16u, 128u);
                        mma_ts(tcol + DCOL + n0, tcol + acol + kk, desc, idesc, acc);
                        acc = 1u;
                    }
                }
                commit(&mbar[wg]);
            }
            if (!mbar_wait(&mbar[wg], phase, p.err)) return;
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- columns that are final now
            const int c0 = n0_of(g, p.G, p.HK, p.NP), c1 = n0_of(g + 1, p.G, p.HK, p.NP);
            for (int c = c0; c < c1; c += 4) {
                float4 v;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(trow + DCOL + c) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) :: "memory");
                *reinterpret_cast<float4*>(p.out + (size_t)row * p.NP + c) = v;
            }
        }
    }
    t1 = clock64();
    if ((tid & 127) == 0) p.clk[wg] = (t1 - t0) / (p.reps > 1 ? p.reps - 1 : 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase_s) : "memory");
}

// ---- single product D[128][N] = A[128][K] * W[N][K]^T, A from shared memory (SS form) or TMEM, any N multiple of 8
__global__ void __launch_bounds__(128, 1) k_single(const float* A, const float* Bimg, float* out, int K, int N, int NP, int ss, int* err) {
    extern __shared__ __align__(128) float sm[];
    const int nB = (K / 4) * NP * 4;
    float* Bhi = sm;
    float* Blo = sm + nB;
    float* Ahi = sm + 2 * nB;             // [K/4][128][4]
    float* Alo = Ahi + (K / 4) * 128 * 4;
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int i = tid; i < 2 * nB; i += blockDim.x) sm[i] = Bimg[i];
    const uint32_t trow = tbase_s + ((uint32_t)(32 * warp) << 16);
    const uint32_t DCOL = 0, SH = 256, SL = 384;
    for (int c = 0; c < K; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(A + (size_t)tid * K + c);
        float x[4] = {v.x, v.y, v.z, v.w}, hi[4], lo[4];
        for (int q = 0; q < 4; ++q) { hi[q] = tf32_hi(x[q]); lo[q] = x[q] - hi[q]; }
        if (ss) {
            *reinterpret_cast<float4*>(Ahi + ((c / 4) * 128 + tid) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(Alo + ((c / 4) * 128 + tid) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        } else {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(trow + SH + c), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(trow + SL + c), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]) : "memory");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc(128, N);
        uint32_t acc = 0;
        for (int t = 0; t < 3; ++t) {
            const float* Bs = t == 2 ? Blo : Bhi;
            for (int kk = 0; kk < K; kk += 8) {
                const uint64_t bdesc = make_desc(smem_u32(Bs) + (uint32_t)((kk / 4) * NP * 16), (uint32_t)NP * 16u, 128u);
                if (ss) {
                    const float* As = t == 1 ? Alo : Ahi;
                    const uint64_t adesc = make_desc(smem_u32(As) + (uint32_t)((kk / 4) * 128 * 16), 128u * 16u, 128u);
                    mma_ss(tbase_s + DCOL, adesc, bdesc, idesc, acc);
                } else {
                    mma_ts(tbase_s + DCOL, tbase_s + (t == 1 ? SL : SH) + kk, bdesc, idesc, acc);
                }
                acc = 1;
            }
        }
        commit(&mbar);
    }
    if (!mbar_wait(&mbar, 0, err)) return;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < NP; c += 4) {
        float4 v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(trow + DCOL + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) :: "memory");
        *reinterpret_cast<float4*>(out + (size_t)tid * NP + c) = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase_s) : "memory");
}

static float hi_of(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }
static void make_Bimg(const std::vector<float>& W, int HK, int NP, int Nreal, std::vector<float>& img) {
    const int nB = (HK / 4) * NP * 4;
    img.assign(2 * nB, 0.f);
    for (int k = 0; k < HK; ++k)
        for (int n = 0; n < Nreal; ++n) {
            float w = W[(size_t)n * HK + k], hi = hi_of(w);
            img[((k / 4) * NP + n) * 4 + (k % 4)] = hi;
            img[nB + ((k / 4) * NP + n) * 4 + (k % 4)] = w - hi;
        }
}

static int run_push(int NT, int HK, int NP, int G, int bmode, int reps, int grid) {
    std::vector<float> W((size_t)NP * HK), h((size_t)NT * 128 * HK), img;
    srand(3);
    const int NG = HK / G;
    for (int n = 0; n < NP; ++n)
        for (int k = 0; k < HK; ++k) W[(size_t)n * HK + k] = (n >= n0_of(k / G, G, HK, NP)) ? (rand() / (float)RAND_MAX - 0.5f) : 0.f;
    for (auto& v : h) v = rand() / (float)RAND_MAX * 2.f - 0.7f;
    make_Bimg(W, HK, NP, NP, img);
    P p;
    p.NT = NT; p.HK = HK; p.NP = NP; p.G = G; p.bmode = bmode; p.reps = reps;
    float *hd, *bd, *od; long long* cd; int* ed;
    cudaMalloc(&hd, h.size() * 4); cudaMalloc(&bd, img.size() * 4); cudaMalloc(&od, (size_t)NT * 128 * NP * 4);
    cudaMalloc(&cd, 8 * 4); cudaMalloc(&ed, 4);
    cudaMemcpy(hd, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(bd, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(od, 0, (size_t)NT * 128 * NP * 4); cudaMemset(ed, 0, 4); cudaMemset(cd, 0, 32);
    p.h = hd; p.Bimg = bd; p.out = od; p.clk = cd; p.err = ed;
    size_t smem = img.size() * 4;
    cudaFuncSetAttribute(k_push, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_push<<<grid, NT * 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out((size_t)NT * 128 * NP); long long clk[4] = {0, 0, 0, 0}; int err = 0;
    cudaMemcpy(out.data(), od, out.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(clk, cd, 32, cudaMemcpyDeviceToHost); cudaMemcpy(&err, ed, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, scale = 0;
    for (int t = 0; t < NT * 128; ++t)
        for (int n = 0; n < NP; ++n) {
            double ex = 0;
            for (int k = 0; k < HK; ++k) ex += (double)W[(size_t)n * HK + k] * fmax(h[(size_t)t * HK + k], 0.f);
            maxerr = fmax(maxerr, fabs(out[(size_t)t * NP + n] - ex)); scale = fmax(scale, fabs(ex));
        }
    printf("push NT %d HK %3d NP %3d G %2d bmode %d grid %3d: %s err %d  max|D-exact| %.3e (scale %.2f)  clk/pass %lld %lld %lld = %.0f clk/step\n",
           NT, HK, NP, G, bmode, grid, cudaGetErrorString(e), err, maxerr, scale, clk[0], clk[1], clk[2], (double)clk[0] / NG);
    cudaFree(hd); cudaFree(bd); cudaFree(od); cudaFree(cd); cudaFree(ed);
    return e != cudaSuccess;
}

static int run_single(int K, int N, int NP, int ss) {
    std::vector<float> W((size_t)N * K), A((size_t)128 * K), img;
    srand(5);
    for (auto& v : W) v = rand() / (float)RAND_MAX - 0.5f;
    for (auto& v : A) v = rand() / (float)RAND_MAX * 2.f - 1.f;
    make_Bimg(W, K, NP, N, img);
    float *ad, *bd, *od; int* ed;
    cudaMalloc(&ad, A.size() * 4); cudaMalloc(&bd, img.size() * 4); cudaMalloc(&od, (size_t)128 * NP * 4); cudaMalloc(&ed, 4);
    cudaMemcpy(ad, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(bd, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(od, 0, (size_t)128 * NP * 4); cudaMemset(ed, 0, 4);
    size_t smem = img.size() * 4 + 2 * (size_t)(K / 4) * 128 * 16;
    cudaFuncSetAttribute(k_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_single<<<1, 128, smem>>>(ad, bd, od, K, N, NP, ss, ed);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out((size_t)128 * NP); int err = 0;
    cudaMemcpy(out.data(), od, out.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&err, ed, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, scale = 0, spill = 0;
    for (int t = 0; t < 128; ++t)
        for (int n = 0; n < NP; ++n) {
            double ex = 0;
            if (n < N) for (int k = 0; k < K; ++k) ex += (double)W[(size_t)n * K + k] * A[(size_t)t * K + k];
            if (n < N) { maxerr = fmax(maxerr, fabs(out[(size_t)t * NP + n] - ex)); scale = fmax(scale, fabs(ex)); }
            else spill = fmax(spill, fabs(out[(size_t)t * NP + n]));
        }
    printf("single K %3d N %3d NP %3d %s: %s err %d  max|D-exact| %.3e (scale %.2f)  max|cols >= N| %.3e\n", K, N, NP, ss ? "SS" : "TS",
           cudaGetErrorString(e), err, maxerr, scale, spill);
    cudaFree(ad); cudaFree(bd); cudaFree(od); cudaFree(ed);
    return e != cudaSuccess;
}

int run_ldalign();
int run_rt();
int main(int argc, char** argv) {
    const char* t = argc > 1 ? argv[1] : "push";
    if (!strcmp(t, "rt")) return run_rt();
    if (!strcmp(t, "ldalign")) return run_ldalign();
    if (!strcmp(t, "push")) {
        int bad = 0;
        bad |= run_push(1, 96, 112, 24, 0, 1, 1);      // lotka_volterra-like chunks
        bad |= run_push(1, 96, 112, 24, 1, 1, 1);      // B by cp.async.bulk
        bad |= run_push(1, 96, 112, 24, 1, 20, 1);
        bad |= run_push(2, 96, 112, 24, 1, 20, 1);
        bad |= run_push(3, 96, 112, 24, 1, 20, 1);
        bad |= run_push(3, 96, 112, 24, 1, 20, 148);
        bad |= run_push(1, 64, 144, 8, 1, 20, 1);      // pyloric-like: one K=8 instruction chain per step
        bad |= run_push(3, 64, 144, 8, 1, 20, 1);
        bad |= run_push(1, 64, 16, 8, 1, 20, 1);       // smallest N
        bad |= run_push(3, 96, 112, 16, 1, 20, 1);
        bad |= run_push(3, 96, 96, 32, 1, 20, 1);
        return bad;
    }
    if (!strcmp(t, "ss")) return run_single(32, 48, 48, 1) | run_single(32, 48, 48, 0);
    if (!strcmp(t, "n8")) {
        int N = argc > 2 ? atoi(argv[2]) : 104;
        return run_single(32, N, ((N + 15) / 16) * 16, 0);
    }
    return 2;
}

// ---- "ldalign": tcgen05.ld / st .32x32b.x8 / .x2 at column addresses that are not multiples of the vector width
__global__ void __launch_bounds__(128, 1) k_ldalign(float* out, int* bad) {
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tbase_s + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < 128; ++c) {
        const float v = (float)(tid * 1000 + c);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(trow + c), "f"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    int nbad = 0;
    for (int c0 = 0; c0 < 24; ++c0) {
        float v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(trow + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) nbad += v[e] != (float)(tid * 1000 + c0 + e);
        float w0, w1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(w0), "=f"(w1) : "r"(trow + 40 + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        nbad += (w0 != (float)(tid * 1000 + 40 + c0)) + (w1 != (float)(tid * 1000 + 41 + c0));
    }
    // unaligned x8 store, read back with x1
    for (int c0 = 65; c0 < 70; ++c0) {
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = (float)(-(tid * 1000 + c0 * 10 + e));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(trow + c0), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) {
            float r;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(r) : "r"(trow + c0 + e) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            nbad += r != v[e];
        }
    }
    if (nbad) atomicAdd(bad, nbad);
    out[tid] = (float)nbad;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tbase_s) : "memory");
}
int run_ldalign() {
    float* od; int* bd; int bad = -1;
    cudaMalloc(&od, 128 * 4); cudaMalloc(&bd, 4); cudaMemset(bd, 0, 4);
    k_ldalign<<<1, 128>>>(od, bd);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&bad, bd, 4, cudaMemcpyDeviceToHost);
    printf("ldalign: %s  mismatches %d (x8 / x2 loads and x8 stores at arbitrary column offsets)\n", cudaGetErrorString(e), bad);
    return e != cudaSuccess || bad != 0;
}

// ---- "rt": the bare staging -> MMA -> read-back round trip (no global memory in the loop), one tile of 128 threads:
// each step stages 8 columns (hi, lo) derived from the previous read-back, pushes them (3 x tcgen05.mma, K = 8, N as
// given) and reads 8 accumulator columns back.
__global__ void __launch_bounds__(128, 1) k_rt(int N, int steps, long long* clk, float* sink) {
    extern __shared__ __align__(128) float sm[];   // B: [2 k chunks][N][4], hi == lo image (timing only)
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * N * 4; i += 128) sm[i] = 0.001f * (float)(i % 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tcol = tbase_s, trow = tcol + ((uint32_t)(32 * warp) << 16);
    const uint32_t SH = 256, SL = 264;
    const uint64_t desc = make_desc(smem_u32(sm), (uint32_t)N * 16u, 128u);
    const uint32_t idesc = make_idesc(128, N);
    float v[8];
    for (int e = 0; e < 8; ++e) v[e] = 0.01f * (float)(tid + e);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        float hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float x = fmaxf(v[e], 0.f); hi[e] = tf32_hi(x); lo[e] = x - hi[e]; }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(trow + SH), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]), "f"(hi[4]), "f"(hi[5]), "f"(hi[6]), "f"(hi[7]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     :: "r"(trow + SL), "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]), "f"(lo[4]), "f"(lo[5]), "f"(lo[6]), "f"(lo[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            mma_ts(tcol, tcol + SH, desc, idesc, 0u);
            mma_ts(tcol, tcol + SL, desc, idesc, 1u);
            mma_ts(tcol, tcol + SH, desc, idesc, 1u);
            commit(&mbar);
        }
        {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(trow + (uint32_t)((s * 8) % (N - 7))) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]) :: "memory");
    }
    const long long t1 = clock64();
    if (tid == 0) clk[0] = (t1 - t0) / steps;
    sink[tid] = v[0] + v[7];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase_s) : "memory");
}
int run_rt() {
    long long* cd; float* sd;
    cudaMalloc(&cd, 8); cudaMalloc(&sd, 128 * 4);
    int bad = 0;
    for (int N : {8, 64, 104, 200}) {
        size_t smem = 2 * (size_t)N * 16;
        k_rt<<<1, 128, smem>>>(N, 2000, cd, sd);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, cd, 8, cudaMemcpyDeviceToHost);
        printf("rt N %3d: %s  %lld clk per stage->mma->readback round trip\n", N, cudaGetErrorString(e), c);
        bad |= e != cudaSuccess;
    }
    return bad;
}
