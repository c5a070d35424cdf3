// Standalone validation of tcgen05.mma kind::tf32 with A from TMEM (thread-private rows) and B from shared memory in
// the [k-or-n chunk of 4][row][4] layout, both as K-major (transposed product) and MN-major (forward product); 3xTF32.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int H = 100, UP = 104, NCH = (H + 3) / 4 + 8;   // W[v][u]; layout Bs[(v/4)][u][v%4], UP rows per chunk
constexpr int A_HI = 0, A_LO = 104, DCOL = 208, NMMA = 48;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// mode 0: forward  D[t][n] = sum_{k<K} a[t][k] * W[v0+n][k]          (B MN-major over the layout)
// mode 1: backward D[t][n] = sum_{v in [va, va+K)} a[t][v] * W[v][u0+n]   (B K-major)
__global__ void k(const float* __restrict__ Wg, const float* __restrict__ Ag, float* __restrict__ Dg, int mode, int K, int off_n, int off_k, int terms, int variant) {
    extern __shared__ __align__(128) float sm[];
    float* Bhi = sm;                         // [NCH][UP][4]
    float* Blo = sm + NCH * UP * 4;
    __shared__ uint32_t tbase;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < NCH * UP * 4; i += blockDim.x) {
        int c = i / (UP * 4), u = (i / 4) % UP, q = i % 4;
        int v = 4 * c + q;
        float w = (v < H && u < H) ? Wg[v * H + u] : 0.f;
        float hi = tf32_trunc(w);
        Bhi[i] = hi;
        Blo[i] = w - hi;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    // A rows (thread = row): hi / lo
    for (int c = 0; c < 104; c += 4) {
        float v[4], h[4], l[4];
        for (int q = 0; q < 4; ++q) { v[q] = (c + q < H) ? Ag[tid * H + c + q] : 0.f; h[q] = tf32_trunc(v[q]); l[q] = v[q] - h[q]; }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(tb + A_HI + c), "f"(h[0]), "f"(h[1]), "f"(h[2]), "f"(h[3]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(tb + A_LO + c), "f"(l[0]), "f"(l[1]), "f"(l[2]), "f"(l[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem writes (generic proxy) -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc(128, NMMA, mode == 0 ? 1 : 0);
        uint32_t acc = 0;
        for (int t = 0; t < terms; ++t) {  // (A_hi,B_hi), (A_hi,B_lo), (A_lo,B_hi)
            const float* Bs = (t == 1) ? Blo : Bhi;
            const uint32_t acol = (t == 2) ? A_LO : A_HI;
            for (int kk = 0; kk < K; kk += 8) {
                uint64_t desc;
                uint32_t a_t;
                if (mode == 0) {   // N = v (chunked dim, MN-major): start at chunk off_n/4; K = u rows kk..kk+7
                    uint32_t addr = smem_u32(Bs) + ((off_n / 4) * UP + (off_k + kk)) * 16;
                    if (variant == 0) desc = make_desc(addr, 8 * 16, UP * 16);
                    else if (variant == 1) desc = make_desc(addr, UP * 16, 8 * 16);
                    else if (variant == 2) desc = make_desc(addr, UP * 16, UP * 16);
                    else desc = make_desc(addr, 16, UP * 16);
                    a_t = tbase + acol + off_k + kk;
                } else {           // N = u rows, K = v (chunked dim, K-major): chunks (off_k+kk)/4, (off_k+kk)/4+1
                    uint32_t addr = smem_u32(Bs) + (((off_k + kk) / 4) * UP + off_n) * 16;
                    desc = make_desc(addr, /*LBO: between the two k chunks*/ UP * 16, /*SBO: between groups of 8 n rows*/ 8 * 16);
                    a_t = tbase + acol + off_k + kk;
                }
                mma_tf32_ts(tbase + DCOL, a_t, desc, idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    // wait (bounded spin)
    {
        uint32_t done = 0; long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            if (++spins > 20000000LL) { if (tid == 0) printf("TIMEOUT waiting for MMA\n"); break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < NMMA; c += 4) {
        float4 v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(tb + DCOL + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) :: "memory");
        Dg[tid * NMMA + c] = v.x; Dg[tid * NMMA + c + 1] = v.y; Dg[tid * NMMA + c + 2] = v.z; Dg[tid * NMMA + c + 3] = v.w;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }
int main() {
    std::vector<float> W(H * H), A(128 * H), D(128 * NMMA);
    srand(1);
    for (auto& w : W) w = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& a : A) a = (rand() / (float)RAND_MAX) * 2.f;
    float *Wd, *Ad, *Dd;
    cudaMalloc(&Wd, W.size() * 4); cudaMalloc(&Ad, A.size() * 4); cudaMalloc(&Dd, D.size() * 4);
    cudaMemcpy(Wd, W.data(), W.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(Ad, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    size_t smem = 2 * NCH * UP * 4 * sizeof(float);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Case { int mode, K, off_n, off_k, terms, variant; };
    Case cases[] = {{0, 8, 16, 0, 1, 0}, {0, 8, 16, 0, 1, 1}, {0, 8, 16, 0, 1, 2}, {0, 8, 16, 0, 1, 3}, {0, 56, 16, 0, 1, 1}, {1, 56, 48, 48, 1, 0}};
    for (auto c : cases) {
        cudaMemset(Dd, 0, D.size() * 4);
        k<<<1, 128, smem>>>(Wd, Ad, Dd, c.mode, c.K, c.off_n, c.off_k, c.terms, c.variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), Dd, D.size() * 4, cudaMemcpyDeviceToHost);
        double max_exact = 0, max_trunc = 0, scale = 0;
        for (int t = 0; t < 128; ++t)
            for (int n = 0; n < NMMA; ++n) {
                double ex = 0, tr = 0;
                for (int kk = 0; kk < c.K; ++kk) {
                    float a, w;
                    if (c.mode == 0) { int u = c.off_k + kk, v = c.off_n + n; a = u < H ? A[t * H + u] : 0; w = (v < H && u < H) ? W[v * H + u] : 0; }
                    else { int v = c.off_k + kk, u = c.off_n + n; a = v < H ? A[t * H + v] : 0; w = (v < H && u < H) ? W[v * H + u] : 0; }
                    ex += (double)a * w; tr += (double)trunc_tf32(a) * trunc_tf32(w);
                }
                double got = D[t * NMMA + n];
                max_exact = fmax(max_exact, fabs(got - ex)); max_trunc = fmax(max_trunc, fabs(got - tr)); scale = fmax(scale, fabs(ex));
            }
        if (c.mode == 0) {
            printf("  variant %d  row0 got:", c.variant);
            for (int n = 0; n < 8; ++n) printf(" %8.4f", D[n]);
            printf("\n             row0 exp:");
            for (int n = 0; n < 8; ++n) {
                double tr = 0;
                for (int kk = 0; kk < c.K; ++kk) tr += (double)trunc_tf32(A[c.off_k + kk]) * trunc_tf32(W[(c.off_n + n) * H + c.off_k + kk]);
                printf(" %8.4f", tr);
            }
            printf("\n");
        }
        printf("mode %d K %3d off_n %2d off_k %2d terms %d: %s  max|D-exact| %.3e  max|D-tf32ref| %.3e  (scale %.2f)\n", c.mode, c.K, c.off_n, c.off_k, c.terms, cudaGetErrorString(e), max_exact, max_trunc, scale);
    }
    return 0;
}
