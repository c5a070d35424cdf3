// validate: TMEM as thread-private scratch via tcgen05.st / tcgen05.ld .32x32b (no MMA involved)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tm_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tm_st1(uint32_t taddr, float a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "f"(a) : "memory");
}
__device__ __forceinline__ float4 tm_ld4(uint32_t taddr) {
    float4 v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void k(float* out, int ncols_per_group, long long* clk) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tbase;
    // this warp's private region: lanes 32*(warp%4).., columns (warp/4)*ncols ..
    const uint32_t my = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * ncols_per_group);
    for (int c = 0; c < ncols_per_group; c += 4)
        tm_st4(my + c, threadIdx.x * 1000.f + c, threadIdx.x * 1000.f + c + 1, threadIdx.x * 1000.f + c + 2, threadIdx.x * 1000.f + c + 3);
    tm_wait_st();
    // in-place single-column update
    tm_st1(my + 5, -1.f * threadIdx.x);
    tm_wait_st();
    float bad = 0.f;
    long long t0 = clock64();
    float acc = 0.f;
    for (int rep = 0; rep < 64; ++rep)
        for (int c = 0; c < ncols_per_group; c += 4) {
            float4 v = tm_ld4(my + c);
            tm_wait_ld();
            acc += v.x + v.y + v.z + v.w;
            if (rep == 0) {
                float e0 = threadIdx.x * 1000.f + c;
                if (v.x != e0 || (c + 1 == 5 ? v.y != -1.f * threadIdx.x : v.y != e0 + 1) || v.z != e0 + 2 || v.w != e0 + 3) bad += 1.f;
            }
        }
    long long t1 = clock64();
    out[threadIdx.x] = bad;
    out[blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 2 * 512 * sizeof(float)); cudaMalloc(&clk, 8);
    for (int T : {128, 352}) {
        k<<<148, T>>>(out, 100, clk);
        cudaError_t e = cudaDeviceSynchronize();
        float h[1024]; long long c;
        cudaMemcpy(h, out, 2 * T * sizeof(float), cudaMemcpyDeviceToHost); cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        float bad = 0; for (int i = 0; i < T; ++i) bad += h[i];
        printf("T=%d status=%s mismatches=%g  clocks for 64x25 dependent ld4+wait per thread: %lld (%.1f clk per ld4)\n", T, cudaGetErrorString(e), bad, c, c / (64.0 * 25));
    }
    return 0;
}
