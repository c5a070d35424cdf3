// rabi_b200: C-ABI library (include/rabi_b200.h) -- host side + launchers.  sm_100a only.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "rabi_internal.h"
#include "maf_pairs.cuh"
#include "maf_fast.cuh"
#include "maf_fast2.cuh"
#include "maf_deep.cuh"
#include "maf_warp.cuh"
#include "maf_tc.h"

using namespace rabi;

static std::string g_create_error;

int rabi_fail(rabi_maf* h, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h)
        h->err = buf;
    else
        g_create_error = buf;
    return 1;
}
#define fail rabi_fail

int rabi_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
#define env_int rabi_env_int

#define CU_OK(h, expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) return fail(h, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

extern "C" const char* rabi_build_info(void) { return "rabi_b200 0.1 (sm_100a, CUDA " __DATE__ ")"; }
extern "C" const char* rabi_last_error(const rabi_maf_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" unsigned long long rabi_launch_count(const rabi_maf_t* h) { return h ? h->launches : 0ull; }

extern "C" int rabi_maf_create(rabi_maf_t** out, int device, int K, int D, int C, int n_hidden, const int* hidden,
                               float log_scale_min, float log_scale_max) {
    if (!out) return fail(nullptr, "out is NULL");
    *out = nullptr;
    if (K < 1 || K > RABI_MAXK) return fail(nullptr, "num_transforms must be in [1,%d], got %d", RABI_MAXK, K);
    if (n_hidden < 1 || n_hidden > RABI_MAXL)
        return fail(nullptr, "number of hidden layers must be in [1,%d], got %d", RABI_MAXL, n_hidden);
    if (D < 1 || C < 1) return fail(nullptr, "D and C must be positive (C == 0: unconditional MADE is not on the path)");
    for (int l = 0; l < n_hidden; ++l)
        if (hidden[l] < D)  // pyro/nn/auto_reg_nn.py: "Hidden dimension must not be less than input dimension."
            return fail(nullptr, "Hidden dimension must not be less than input dimension.");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, "no CUDA device available (rabi_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, "invalid device %d", device);
    rabi_maf* h = new rabi_maf();
    h->device = device;
    h->K = K;
    h->D = D;
    h->C = C;
    h->L = n_hidden;
    for (int l = 0; l < n_hidden; ++l) h->H[l] = hidden[l];
    h->lo = log_scale_min;
    h->hi = log_scale_max;
    h->perm.assign(K, std::vector<int64_t>());
    h->post.assign(K, std::vector<int64_t>());
    h->masks.assign(K, std::vector<std::vector<float>>(n_hidden + 1));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete h;
        return fail(nullptr, "cudaGetDeviceProperties failed");
    }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    *out = h;
    return 0;
}

extern "C" int rabi_maf_destroy(rabi_maf_t* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    rabi_tc_destroy(h);
    if (h->f_dev) cudaFree(h->f_dev);
    if (h->m_dev) cudaFree(h->m_dev);
    if (h->ws) cudaFree(h->ws);
    if (h->deep_ws) cudaFree(h->deep_ws);
    for (auto& e : h->prof_events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    delete h;
    return 0;
}

static int layer_out(const rabi_maf* h, int layer) { return layer < h->L ? h->H[layer] : 2 * h->D; }
static int layer_in(const rabi_maf* h, int layer) { return layer == 0 ? h->C + h->D : h->H[layer - 1]; }

extern "C" int rabi_maf_set_permutation(rabi_maf_t* h, int k, const int64_t* perm_host) {
    if (!h) return 1;
    if (k < 0 || k >= h->K) return fail(h, "transform index %d out of range", k);
    std::vector<char> seen(h->D, 0);
    for (int j = 0; j < h->D; ++j) {
        int64_t v = perm_host[j];
        if (v < 0 || v >= h->D || seen[v]) return fail(h, "permutation of transform %d is not a permutation of 0..D-1", k);
        seen[v] = 1;
    }
    h->perm[k].assign(perm_host, perm_host + h->D);
    h->finalized = false;
    return 0;
}

extern "C" int rabi_maf_set_post_permutation(rabi_maf_t* h, int k, const int64_t* perm_host) {
    if (!h) return 1;
    if (k < 0 || k >= h->K) return fail(h, "transform index %d out of range", k);
    if (!perm_host) {
        h->post[k].clear();
        h->finalized = false;
        return 0;
    }
    std::vector<char> seen(h->D, 0);
    for (int j = 0; j < h->D; ++j) {
        int64_t v = perm_host[j];
        if (v < 0 || v >= h->D || seen[v]) return fail(h, "permute%d is not a permutation of 0..D-1", k);
        seen[v] = 1;
    }
    h->post[k].assign(perm_host, perm_host + h->D);
    h->finalized = false;
    return 0;
}

extern "C" int rabi_maf_set_mask(rabi_maf_t* h, int k, int layer, const float* mask_host, int out_features,
                                 int in_features) {
    if (!h) return 1;
    if (k < 0 || k >= h->K) return fail(h, "transform index %d out of range", k);
    if (layer < 0 || layer > h->L) return fail(h, "layer index %d out of range", layer);
    if (out_features != layer_out(h, layer) || in_features != layer_in(h, layer))
        return fail(h, "mask of t%d.layers.%d has shape [%d,%d], expected [%d,%d]", k, layer, out_features, in_features,
                    layer_out(h, layer), layer_in(h, layer));
    h->masks[k][layer].assign(mask_host, mask_host + (size_t)out_features * in_features);
    h->finalized = false;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int rabi_maf_finalize_structure(rabi_maf_t* h) {
    if (!h) return 1;
    const int K = h->K, D = h->D, C = h->C, L = h->L;
    CU_OK(h, cudaSetDevice(h->device));
    MafImg& g = h->img;
    memset(&g, 0, sizeof(g));
    g.K = K;
    g.D = D;
    g.C = C;
    g.L = L;
    g.lo = h->lo;
    g.hi = h->hi;
    g.Hmax = 0;
    for (int l = 0; l < L; ++l) {
        g.H[l] = h->H[l];
        g.Hmax = std::max(g.Hmax, h->H[l]);
    }
    size_t nf = 0;
    std::vector<int>& meta = h->meta_host;
    meta.clear();
    auto falloc = [&](size_t n) {
        size_t o = nf;
        nf += (n + 3) & ~(size_t)3;
        return (int)o;
    };
    auto ialloc = [&](size_t n) {
        size_t o = meta.size();
        meta.resize(o + n, 0);
        return (int)o;
    };
    const int H0 = h->H[0], H0p = round4(H0);
    g.H0p = H0p;
    g.ldKH = K * H0p;
    g.W0cT = falloc((size_t)C * g.ldKH);
    for (int k = 0; k < K; ++k) {
        if ((int)h->perm[k].size() != D) return fail(h, "permutation of transform %d was not set", k);
        for (int l = 0; l <= L; ++l)
            if (h->masks[k][l].empty()) return fail(h, "mask of t%d.layers.%d was not set", k, l);
        TImg& t = g.t[k];
        const std::vector<int64_t>& perm = h->perm[k];
        t.ldc = round4(C);
        t.W0c = falloc((size_t)(H0p + 4) * t.ldc);
        // [fast block: W0t, b[1], W[1], WoT, boF] is contiguous so that k_fast stages it with one copy
        t.ldt = round4(D);
        t.W0t = falloc((size_t)(H0p + 4) * t.ldt);
        for (int l = 1; l < L; ++l) {
            t.ld[l] = round4(h->H[l - 1]);
            t.b[l] = falloc(round4(h->H[l]) + 4);
            t.W[l] = falloc((size_t)(round4(h->H[l]) + 4) * t.ld[l]);
        }
        t.WoT = falloc((size_t)(round4(h->H[L - 1]) + 4) * 2 * t.ldt);
        t.boF = falloc(2 * t.ldt);
        t.fast_end = (int)nf;
        t.b0 = falloc(H0p + 4);
        t.ldo = round4(h->H[L - 1]);
        t.Wo = falloc((size_t)(round4(2 * D) + 4) * t.ldo);
        t.bo = falloc(round4(2 * D) + 4);

        t.perm = ialloc(D);
        for (int j = 0; j < D; ++j) meta[t.perm + j] = (int)perm[j];
        t.permy = ialloc(D);
        if (h->post[k].empty()) {
            for (int j = 0; j < D; ++j) meta[t.permy + j] = (int)perm[j];
        } else {  // Permute: y[i] = x[post[i]]  ->  variable perm[j] lands at i with post[i] == perm[j]
            std::vector<int> inv(D);
            for (int i = 0; i < D; ++i) inv[h->post[k][i]] = i;
            for (int j = 0; j < D; ++j) meta[t.permy + j] = inv[perm[j]];
        }
        // ---- layer 0: context columns must be fully connected; theta columns a prefix in order space
        t.deg0 = ialloc(H0);
        {
            const std::vector<float>& m = h->masks[k][0];
            const int in = C + D;
            for (int u = 0; u < H0; ++u) {
                for (int c = 0; c < C; ++c)
                    if (m[(size_t)u * in + c] != 1.f)
                        return fail(h, "t%d.layers.0.mask: context column %d of unit %d is masked; not a Pyro MADE mask", k, c, u);
                int dg = 0;
                for (int j = 0; j < D; ++j) dg += m[(size_t)u * in + C + perm[j]] != 0.f;
                for (int j = 0; j < D; ++j)
                    if ((m[(size_t)u * in + C + perm[j]] != 0.f) != (j < dg))
                        return fail(h, "t%d.layers.0.mask: row %d is not a prefix in permutation order", k, u);
                meta[t.deg0 + u] = dg;
                if (u > 0 && dg < meta[t.deg0 + u - 1])
                    return fail(h, "t%d.layers.0.mask: unit degrees are not non-decreasing", k);
            }
        }
        // groups of layer 0
        std::vector<std::vector<int>> deg(L);
        deg[0].assign(meta.begin() + t.deg0, meta.begin() + t.deg0 + H0);
        auto build_groups = [&](int l) -> int {
            int off = ialloc(D + 1);
            const std::vector<int>& d = deg[l];
            int n = (int)d.size();
            int u = 0;
            for (int j = 0; j <= D; ++j) {
                while (u < n && d[u] < j) ++u;
                meta[off + j] = u;
            }
            meta[off + D] = n;
            return off;
        };
        for (int u = 0; u < H0; ++u)
            if (deg[0][u] < 0 || deg[0][u] > D - 1)
                return fail(h, "t%d.layers.0.mask: unit %d sees all %d variables (degree out of range)", k, u, D);
        t.grp[0] = build_groups(0);
        // ---- hidden layers
        for (int l = 1; l < L; ++l) {
            const int Hl = h->H[l], Hin = h->H[l - 1];
            const std::vector<float>& m = h->masks[k][l];
            t.pre[l] = ialloc(Hl);
            deg[l].assign(Hl, 0);
            for (int v = 0; v < Hl; ++v) {
                int e = 0;
                for (int u = 0; u < Hin; ++u) e += m[(size_t)v * Hin + u] != 0.f;
                for (int u = 0; u < Hin; ++u)
                    if ((m[(size_t)v * Hin + u] != 0.f) != (u < e))
                        return fail(h, "t%d.layers.%d.mask: row %d is not a prefix", k, l, v);
                meta[t.pre[l] + v] = e;
                // degree: the j with grp_{l-1}[j+1] == e
                int dj = -1;
                for (int j = 0; j < D; ++j)
                    if (meta[t.grp[l - 1] + j + 1] == e) {
                        dj = j;
                        break;
                    }
                if (dj < 0) return fail(h, "t%d.layers.%d.mask: row %d does not match a MADE degree", k, l, v);
                deg[l][v] = dj;
                if (v > 0 && dj < deg[l][v - 1]) return fail(h, "t%d.layers.%d.mask: degrees not non-decreasing", k, l);
            }
            t.grp[l] = build_groups(l);
            t.first[l] = ialloc(round4(Hin) + 4);
            for (int u = 0; u < round4(Hin) + 4; ++u) {
                int v = 0;
                while (v < Hl && meta[t.pre[l] + v] <= u) ++v;
                meta[t.first[l] + u] = v;
            }
        }
        // ---- output layer (rows perm[j] and D + perm[j])
        {
            const int HL = h->H[L - 1];
            const std::vector<float>& m = h->masks[k][L];
            t.meta_fast_end = (int)meta.size();
            t.preo = ialloc(D);
            for (int j = 0; j < D; ++j) {
                int e = meta[t.grp[L - 1] + j + 1];
                for (int half = 0; half < 2; ++half) {
                    size_t row = (size_t)(half * D + perm[j]);
                    for (int v = 0; v < HL; ++v)
                        if ((m[row * HL + v] != 0.f) != (v < e))
                            return fail(h, "t%d.layers.%d.mask: output row %d is not the MADE prefix", k, L, (int)row);
                }
                meta[t.preo + j] = e;
            }
            t.firsto = ialloc(round4(HL) + 4);
            for (int v = 0; v < round4(HL) + 4; ++v) {
                int j = 0;
                while (j < D && meta[t.preo + j] <= v) ++j;
                meta[t.firsto + v] = j;
            }
            t.first0 = ialloc(D);
            for (int j = 0; j < D; ++j) {
                int u = 0;
                while (u < H0 && meta[t.deg0 + u] <= j) ++u;
                meta[t.first0 + j] = u;
            }
        }
    }
    if (h->f_dev) cudaFree(h->f_dev);
    if (h->m_dev) cudaFree(h->m_dev);
    h->n_float = nf;
    h->n_int = meta.size();
    CU_OK(h, cudaMalloc(&h->f_dev, nf * sizeof(float)));
    CU_OK(h, cudaMemset(h->f_dev, 0, nf * sizeof(float)));
    CU_OK(h, cudaMalloc(&h->m_dev, (meta.size() + 4) * sizeof(int)));  // +4: k_warp stages it in whole quads
    CU_OK(h, cudaMemset(h->m_dev, 0, (meta.size() + 4) * sizeof(int)));
    CU_OK(h, cudaMemcpy(h->m_dev, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
    g.f = h->f_dev;
    g.m = h->m_dev;
    h->finalized = true;
    rabi_tc_destroy(h);   // derived images are rebuilt for the new structure
    h->weights_version++;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// pack kernels: raw torch parameters -> masked, order-space image
__global__ void k_pack_layer0(MafImg g, int k, const float* __restrict__ W, const float* __restrict__ b, float* f) {
    const TImg t = g.t[k];
    const int H0 = g.H[0], C = g.C, D = g.D, in = C + D;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H0 * in) return;
    int u = idx / in, c = idx % in;
    if (c < C) {
        float w = W[(size_t)u * in + c];
        f[t.W0c + (size_t)u * t.ldc + c] = w;
        f[g.W0cT + (size_t)c * g.ldKH + k * g.H0p + u] = w;
        if (c == 0) f[t.b0 + u] = b[u];
    } else {
        int j = c - C;
        int dg = g.m[t.deg0 + u];
        f[t.W0t + (size_t)u * t.ldt + j] = (j < dg) ? W[(size_t)u * in + C + g.m[t.perm + j]] : 0.f;
    }
}
__global__ void k_pack_hidden(MafImg g, int k, int l, const float* __restrict__ W, const float* __restrict__ b, float* f) {
    const TImg t = g.t[k];
    const int Hl = g.H[l], Hin = g.H[l - 1];
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Hl * Hin) return;
    int v = idx / Hin, u = idx % Hin;
    f[t.W[l] + (size_t)v * t.ld[l] + u] = (u < g.m[t.pre[l] + v]) ? W[(size_t)v * Hin + u] : 0.f;
    if (u == 0) f[t.b[l] + v] = b[v];
}
__global__ void k_pack_out(MafImg g, int k, const float* __restrict__ W, const float* __restrict__ b, float* f) {
    const TImg t = g.t[k];
    const int D = g.D, HL = g.H[g.L - 1];
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * D * HL) return;
    int r = idx / HL, v = idx % HL;
    int half = r / D, j = r % D;
    int src = half * D + g.m[t.perm + j];
    const float wv = (v < g.m[t.preo + j]) ? W[(size_t)src * HL + v] : 0.f;
    f[t.Wo + (size_t)r * t.ldo + v] = wv;
    f[t.WoT + (size_t)v * 2 * t.ldt + 2 * j + half] = wv;   // (mean, log-scale) pairs per position
    if (v == 0) {
        f[t.bo + r] = b[src];
        f[t.boF + 2 * j + half] = b[src];
    }
}

#define REQUIRE_READY(h)                                                               \
    do {                                                                               \
        if (!(h)) return 1;                                                            \
        if (!(h)->finalized) return fail(h, "rabi_maf_finalize_structure was not called"); \
        CU_OK(h, cudaSetDevice((h)->device));                                          \
    } while (0)

#define LAUNCH_CHECK(h)                                                        \
    do {                                                                       \
        (h)->launches++;                                                       \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) return fail(h, "kernel launch failed: %s", cudaGetErrorString(_e)); \
    } while (0)

extern "C" int rabi_maf_set_weights(rabi_maf_t* h, int k, int layer, const float* W_dev, const float* b_dev,
                                    rabi_stream_t s) {
    REQUIRE_READY(h);
    if (k < 0 || k >= h->K) return fail(h, "transform index %d out of range", k);
    if (layer < 0 || layer > h->L) return fail(h, "layer index %d out of range", layer);
    cudaStream_t st = (cudaStream_t)s;
    int n = layer_out(h, layer) * layer_in(h, layer);
    int nb = (n + 255) / 256;
    if (layer == 0)
        k_pack_layer0<<<nb, 256, 0, st>>>(h->img, k, W_dev, b_dev, h->f_dev);
    else if (layer < h->L)
        k_pack_hidden<<<nb, 256, 0, st>>>(h->img, k, layer, W_dev, b_dev, h->f_dev);
    else
        k_pack_out<<<nb, 256, 0, st>>>(h->img, k, W_dev, b_dev, h->f_dev);
    LAUNCH_CHECK(h);
    h->weights_version++;
    return 0;
}

static int ensure_ws(rabi_maf* h, size_t bytes) {
    if (bytes <= h->ws_bytes) return 0;
    // grow geometrically; synchronises (rare: only when a larger batch than ever before is seen)
    size_t nb = std::max(bytes, h->ws_bytes * 2);
    if (h->ws) {
        CU_OK(h, cudaDeviceSynchronize());
        CU_OK(h, cudaFree(h->ws));
        h->ws = nullptr;
        h->ws_bytes = 0;
    }
    CU_OK(h, cudaMalloc(&h->ws, nb));
    h->ws_bytes = nb;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// context projection:  A[k,u,b] = b0_k[u] + sum_c W0_k[u,c] * ((x[b,c] (+ delta[b,c]) - mean[c]) / std[c])
// Register-tiled, shared-memory-staged GEMMs (4 rows x 8 outputs per thread, packed fma.rn.f32x2):
//   forward : A[k,u,b] = b0_k[u] + sum_c W0_k[u,c] * ctx[b,c],  ctx = ((x (+ delta)) - mean) / std     (grid.y = k)
//   backward: gx[b,c]  = (1/std[c]) * sum_{k,u} W0_k[u,c] * gA[k,u,b]
constexpr int CP_MAXROWS = 80;  // max rows per block (forward); the launcher picks a multiple of 4 that fills one wave
constexpr int CP_LD = 84;       // padded row stride of the transposed x tile
constexpr int CPB_ROWS = 32;    // rows per block (backward)
__global__ void __launch_bounds__(256) k_ctx_project(MafImg g, const float* __restrict__ x, const float* __restrict__ delta,
                                                     const float* __restrict__ zmean, const float* __restrict__ zstd, int B,
                                                     float* __restrict__ A, int rows) {
    extern __shared__ __align__(16) float sm[];
    const int C = g.C, H0 = g.H[0], k = blockIdx.y;
    const int ldw = g.ldKH;                  // W0cT: [C][ldKH], column k*H0p + u
    const int Hp8 = (H0 + 7) & ~7;
    float* Xs = sm;                          // [C][CP_LD]   ctx tile, transposed
    float* Wsm = sm + (size_t)C * CP_LD;     // [C][Hp8]
    const int b0 = blockIdx.x * rows;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
    // x tile: lanes walk the context dimension (coalesced), warps walk the rows; loads are issued in batches of 10
    for (int c = tx; c < C; c += 32) {
        const float mu = zmean ? zmean[c] : 0.f, sd = zmean ? zstd[c] : 1.f;
        float v[CP_MAXROWS / 8];
#pragma unroll
        for (int it = 0; it < CP_MAXROWS / 8; ++it) {
            const int r = ty + 8 * it, b = b0 + r;
            v[it] = (r < rows && b < B) ? x[(size_t)b * C + c] : 0.f;
        }
        if (delta) {
#pragma unroll
            for (int it = 0; it < CP_MAXROWS / 8; ++it) {
                const int r = ty + 8 * it, b = b0 + r;
                if (r < rows && b < B) v[it] += delta[(size_t)b * C + c];
            }
        }
#pragma unroll
        for (int it = 0; it < CP_MAXROWS / 8; ++it) {
            const int r = ty + 8 * it;
            if (r < rows) Xs[c * CP_LD + r] = (b0 + r < B) ? (zmean ? (v[it] - mu) / sd : v[it]) : 0.f;
        }
    }
    // weight slice of transform k, 128-bit rows, 13 loads in flight per thread
    const int n4 = Hp8 >> 2, h4 = g.H0p >> 2;
    for (int q = tx; q < n4; q += 32) {
        for (int c0 = ty; c0 < C; c0 += 8 * 13) {
            float4 w[13];
#pragma unroll
            for (int it = 0; it < 13; ++it) {
                const int c = c0 + 8 * it;
                w[it] = (c < C && q < h4) ? __ldg(reinterpret_cast<const float4*>(g.f + g.W0cT + (size_t)c * ldw + k * g.H0p) + q)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < 13; ++it) {
                const int c = c0 + 8 * it;
                if (c < C) reinterpret_cast<float4*>(Wsm + (size_t)c * Hp8)[q] = w[it];
            }
        }
    }
    __syncthreads();
    const int ng = Hp8 / 8, RG = rows >> 2;  // unit groups of 8, row groups of 4
    for (int t = threadIdx.x; t < ng * RG; t += blockDim.x) {
        const int rg = t % RG, ug = t / RG;
        float2 acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            const float4 xv = *reinterpret_cast<const float4*>(Xs + c * CP_LD + 4 * rg);
            const float4 w0 = *reinterpret_cast<const float4*>(Wsm + c * Hp8 + 8 * ug);
            const float4 w1 = *reinterpret_cast<const float4*>(Wsm + c * Hp8 + 8 * ug + 4);
            const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 xx = make_float2(xr[i], xr[i]);
                acc[i][0] = ffma2(make_float2(w0.x, w0.y), xx, acc[i][0]);
                acc[i][1] = ffma2(make_float2(w0.z, w0.w), xx, acc[i][1]);
                acc[i][2] = ffma2(make_float2(w1.x, w1.y), xx, acc[i][2]);
                acc[i][3] = ffma2(make_float2(w1.z, w1.w), xx, acc[i][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int u = 8 * ug + j;
            if (u >= H0) continue;
            const float bias = g.f[g.t[k].b0 + u];
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = bias + ((j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x);
            float* out = A + ((size_t)k * H0 + u) * B + b0 + 4 * rg;
            if (b0 + 4 * rg + 3 < B && (B & 3) == 0) {
                *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (b0 + 4 * rg + i < B) out[i] = o[i];
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_ctx_project_bwd(MafImg g, const float* __restrict__ gA,
                                                         const float* __restrict__ zstd, int B, float* __restrict__ gx,
                                                         int accumulate) {
    extern __shared__ __align__(16) float sm[];
    const int C = g.C, K = g.K, H0 = g.H[0];
    const int Cp8 = (C + 7) & ~7;
    float* Gs = sm;                            // [H0][CPB_ROWS]
    float* Wsm = sm + (size_t)H0 * CPB_ROWS;   // [H0][Cp8]
    const int b0 = blockIdx.x * CPB_ROWS;
    const int ng = Cp8 / 8, RG = CPB_ROWS / 4;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 4 warps
    // each thread owns up to two (row group, column group) tiles (covers C <= 256 with 128 threads)
    float2 acc[2][4][4];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][i][j] = make_float2(0.f, 0.f);
    for (int k = 0; k < K; ++k) {
        __syncthreads();
        // cotangent tile: lanes walk the rows (coalesced); all loads of a thread are issued before the stores
        for (int u0 = ty; u0 < H0; u0 += 4 * 13) {
            float gv[13];
#pragma unroll
            for (int it = 0; it < 13; ++it) {
                const int u = u0 + 4 * it;
                gv[it] = (u < H0 && b0 + tx < B) ? gA[((size_t)k * H0 + u) * B + b0 + tx] : 0.f;
            }
#pragma unroll
            for (int it = 0; it < 13; ++it) {
                const int u = u0 + 4 * it;
                if (u < H0) Gs[u * CPB_ROWS + tx] = gv[it];
            }
        }
        const float* W = g.f + g.t[k].W0c;
        const int ldc = g.t[k].ldc, n4 = Cp8 >> 2, c4 = ldc >> 2;
        for (int q = tx; q < n4; q += 32) {
            for (int u0 = ty; u0 < H0; u0 += 4 * 13) {
                float4 w[13];
#pragma unroll
                for (int it = 0; it < 13; ++it) {
                    const int u = u0 + 4 * it;
                    w[it] = (u < H0 && q < c4) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)u * ldc) + q)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int it = 0; it < 13; ++it) {
                    const int u = u0 + 4 * it;
                    if (u < H0) reinterpret_cast<float4*>(Wsm + (size_t)u * Cp8)[q] = w[it];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int t = threadIdx.x + s * blockDim.x;
            if (t >= ng * RG) continue;
            const int rg = t % RG, cg = t / RG;
#pragma unroll 4
            for (int u = 0; u < H0; ++u) {
                const float4 gv = *reinterpret_cast<const float4*>(Gs + u * CPB_ROWS + 4 * rg);
                const float4 w0 = *reinterpret_cast<const float4*>(Wsm + u * Cp8 + 8 * cg);
                const float4 w1 = *reinterpret_cast<const float4*>(Wsm + u * Cp8 + 8 * cg + 4);
                const float gr[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 gg = make_float2(gr[i], gr[i]);
                    acc[s][i][0] = ffma2(make_float2(w0.x, w0.y), gg, acc[s][i][0]);
                    acc[s][i][1] = ffma2(make_float2(w0.z, w0.w), gg, acc[s][i][1]);
                    acc[s][i][2] = ffma2(make_float2(w1.x, w1.y), gg, acc[s][i][2]);
                    acc[s][i][3] = ffma2(make_float2(w1.z, w1.w), gg, acc[s][i][3]);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int t = threadIdx.x + s * blockDim.x;
        if (t >= ng * RG) continue;
        const int rg = t % RG, cg = t / RG;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = b0 + 4 * rg + i;
            if (b >= B) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * cg + j;
                if (c >= C) continue;
                float v = (j & 1) ? acc[s][i][j >> 1].y : acc[s][i][j >> 1].x;
                if (zstd) v /= zstd[c];
                float* o = gx + (size_t)b * C + c;
                *o = accumulate ? *o + v : v;
            }
        }
    }
}

// mean over the sample axis: out[b] = (1/S) sum_s in[s,b]
__global__ void k_mean_over_s(const float* __restrict__ in, int S, int B, float* __restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += in[(size_t)s * B + b];
    out[b] = acc / (float)S;
}

// advertorch perturb_iterative(ord=2) update, one warp per row.
__global__ void k_pgd_update(const float* __restrict__ x, float* __restrict__ delta, const float* __restrict__ grad,
                             int B, int dx, float eps, float eps_iter, float cmin, float cmax, float floor_,
                             const float* __restrict__ eps_rows, const float* __restrict__ eps_iter_rows,
                             const float* __restrict__ col_mask) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    if (eps_rows) {   // batched eps-sweep: every row carries its own radius and step size
        eps = eps_rows[warp];
        eps_iter = eps_iter_rows[warp];
    }
    const float* gr = grad + (size_t)warp * dx;
    const float* xr = x + (size_t)warp * dx;
    float* dr = delta + (size_t)warp * dx;
    float n2 = 0.f;
    for (int c = lane; c < dx; c += 32) {
        const float gm = col_mask ? gr[c] * col_mask[c] : gr[c];   // masked-feature attack: frozen columns carry no gradient
        n2 = fmaf(gm, gm, n2);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    float inv = 1.f / fmaxf(sqrtf(n2), floor_);
    float d2 = 0.f;
    for (int c = lane; c < dx; c += 32) {
        const float mk = col_mask ? col_mask[c] : 1.f;
        float d = dr[c] + (gr[c] * mk * inv) * eps_iter;
        d = mk != 0.f ? fminf(fmaxf(xr[c] + d, cmin), cmax) - xr[c] : 0.f;
        dr[c] = d;
        d2 = fmaf(d, d, d2);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    float factor = fminf(eps / sqrtf(d2), 1.f);
    for (int c = lane; c < dx; c += 32) dr[c] *= factor;
}

__global__ void k_clamp_add(const float* __restrict__ x, const float* __restrict__ delta, size_t n, float cmin,
                            float cmax, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fminf(fmaxf(x[i] + delta[i], cmin), cmax);
}

// ---------------------------------------------------------------------------------------------------------
static int launch_ctx_project(rabi_maf* h, const float* x, const float* delta, const float* zm, const float* zs, int B,
                              float* A, cudaStream_t st) {
    if ((zm == nullptr) != (zs == nullptr)) return fail(h, "zs_mean and zs_std must both be given or both be NULL");
    const int Hp8 = (h->H[0] + 7) & ~7;
    size_t smem = ((size_t)h->C * CP_LD + (size_t)h->C * Hp8) * sizeof(float);
    if (smem > (size_t)h->max_smem_optin) return fail(h, "context dimension %d too large for the projection kernel", h->C);
    CU_OK(h, cudaFuncSetAttribute(k_ctx_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // rows per block: a multiple of 4 such that grid.x * K blocks fill whole waves (blocks resident per SM from smem)
    const int per_sm = std::max(1, (int)((size_t)(h->max_smem_optin) / (smem + 1024)));
    const int slots = std::max(1, h->sm_count * per_sm / h->K);
    int rows = 64;
    {
        int waves = (B + slots * CP_MAXROWS - 1) / (slots * CP_MAXROWS);  // as few waves as the largest tile allows
        int tiles = std::max(1, waves * slots);
        rows = std::min(CP_MAXROWS, std::max(4, ((B + tiles - 1) / tiles + 3) & ~3));
    }
    dim3 grid((B + rows - 1) / rows, h->K);
    k_ctx_project<<<grid, 256, smem, st>>>(h->img, x, delta, zm, zs, B, A, rows);
    LAUNCH_CHECK(h);
    return 0;
}

extern "C" int rabi_maf_ctx_project(rabi_maf_t* h, const float* x_dev, const float* zs_mean_dev,
                                    const float* zs_std_dev, int B, float* A_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (B <= 0) return 0;
    return launch_ctx_project(h, x_dev, nullptr, zs_mean_dev, zs_std_dev, B, A_dev, (cudaStream_t)s);
}

static int launch_ctx_project_bwd(rabi_maf* h, const float* gA, const float* zs, int B, float* gx, int accumulate,
                                  cudaStream_t st) {
    const int Cp8 = (h->C + 7) & ~7;
    if ((Cp8 / 8) * (CPB_ROWS / 4) > 256) return fail(h, "context dimension %d too large for the projection kernel", h->C);
    size_t smem = ((size_t)h->H[0] * CPB_ROWS + (size_t)h->H[0] * Cp8) * sizeof(float);
    if (smem > (size_t)h->max_smem_optin) return fail(h, "context dimension %d too large for the projection kernel", h->C);
    CU_OK(h, cudaFuncSetAttribute(k_ctx_project_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ctx_project_bwd<<<(B + CPB_ROWS - 1) / CPB_ROWS, 128, smem, st>>>(h->img, gA, zs, B, gx, accumulate);
    LAUNCH_CHECK(h);
    return 0;
}

extern "C" int rabi_maf_ctx_project_bwd(rabi_maf_t* h, const float* gA_dev, const float* zs_std_dev, int B,
                                        float* gx_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (B <= 0) return 0;
    return launch_ctx_project_bwd(h, gA_dev, zs_std_dev, B, gx_dev, 0, (cudaStream_t)s);
}

template <int MODE>
static int launch_pairs(rabi_maf* h, const PairArgs& a, cudaStream_t st) {
    const long long npairs = (long long)a.S * a.B;
    if (npairs <= 0) return 0;
    size_t per_thread = pair_smem_floats_per_thread(h->K, h->D, h->L, h->img.Hmax) * sizeof(float);
    int T = 256;
    const size_t budget = std::min<size_t>((size_t)h->max_smem_optin, 200 * 1024);
    while (T > 32 && (size_t)T * per_thread > budget) T -= 32;
    if ((size_t)T * per_thread > (size_t)h->max_smem_optin)
        return fail(h, "model too large for the per-pair kernel (%zu B of shared memory per pair)", per_thread);
    size_t smem = (size_t)T * per_thread;
    CU_OK(h, cudaFuncSetAttribute(k_pairs<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long nblk = (npairs + T - 1) / T;
    const bool prof = h->profiling && MODE == MODE_RKL_GRAD;
    if (prof) {
        if (h->prof_used == h->prof_events.size()) {
            cudaEvent_t e0, e1;
            CU_OK(h, cudaEventCreate(&e0));
            CU_OK(h, cudaEventCreate(&e1));
            h->prof_events.push_back({e0, e1});
        }
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].first, st));
    }
    k_pairs<MODE><<<(unsigned)nblk, T, smem, st>>>(h->img, a);
    LAUNCH_CHECK(h);
    if (prof) {
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].second, st));
        h->prof_used++;
    }
    return 0;
}

extern "C" int rabi_profile(rabi_maf_t* h, int enable) {
    if (!h) return 1;
    h->profiling = enable != 0;
    h->prof_used = 0;
    return 0;
}

extern "C" int rabi_profile_read(rabi_maf_t* h, double* total_ms, unsigned long long* launches) {
    if (!h) return 1;
    CU_OK(h, cudaSetDevice(h->device));
    CU_OK(h, cudaDeviceSynchronize());
    double tot = 0.0;
    for (size_t i = 0; i < h->prof_used; ++i) {
        float ms = 0.f;
        CU_OK(h, cudaEventElapsedTime(&ms, h->prof_events[i].first, h->prof_events[i].second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = h->prof_used;
    h->prof_used = 0;
    return 0;
}

int rabi_prof_begin(rabi_maf* h, cudaStream_t st) {
    if (!h->profiling) return 0;
    if (h->prof_used == h->prof_events.size()) {
        cudaEvent_t e0, e1;
        CU_OK(h, cudaEventCreate(&e0));
        CU_OK(h, cudaEventCreate(&e1));
        h->prof_events.push_back({e0, e1});
    }
    CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].first, st));
    return 0;
}
int rabi_prof_end(rabi_maf* h, cudaStream_t st) {
    if (!h->profiling) return 0;
    CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].second, st));
    h->prof_used++;
    return 0;
}

// fp32 FMA-pipe peak (roofline denominator for the FMA-bound kernels; MEASURED_PEAKS.json has no fp32 entry)
__global__ void k_fma_peak(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = (float)(threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(acc[j], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    if (s == 12345.678f) out[0] = s;
}

extern "C" double rabi_fma_peak_tflops(int device, int repeats) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    float* out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fma_peak<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);  // warm-up
    double best = 0.0;
    for (int r = 0; r < (repeats > 0 ? repeats : 5); ++r) {
        cudaEventRecord(e0);
        k_fma_peak<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

template <int MODE>
static int launch_fast(rabi_maf* h, FastArgs fa, cudaStream_t st);  // defined below
static int launch_deep(rabi_maf* h, const DeepArgs& a, int mode, cudaStream_t st);

extern "C" int rabi_maf_log_prob_fwd(rabi_maf_t* h, const float* A_dev, const float* theta_dev, int S, int B,
                                     float* logp_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    PairArgs a = {};
    a.A_p = A_dev;
    a.in = theta_dev;
    a.out_lp = logp_dev;
    a.S = S;
    a.B = B;
    {
        FastArgs fa = {};
        fa.A_p = A_dev;
        fa.A_q = A_dev;
        fa.z = theta_dev;
        fa.diff = logp_dev;
        fa.S = S;
        fa.B = B;
        int rc = 2;
        if (!env_int("RABI_FORCE_DEEP", 0)) {
            TcCall tc = {nullptr, A_dev, theta_dev, logp_dev, nullptr, nullptr, S, B, 0.f};
            rc = rabi_tc_launch(h, tc, TC_LOGPROB, (cudaStream_t)s);
            if (rc == 2) rc = launch_fast<FAST_LOGPROB>(h, fa, (cudaStream_t)s);
        }
        if (rc != 2) return rc;
        DeepArgs da = {};
        da.A_p = A_dev;
        da.A_q = A_dev;
        da.z = theta_dev;
        da.diff = logp_dev;
        da.S = S;
        da.B = B;
        rc = launch_deep(h, da, DEEP_LOGPROB, (cudaStream_t)s);
        if (rc != 2) return rc;
    }
    return launch_pairs<MODE_LOGPROB>(h, a, (cudaStream_t)s);
}

extern "C" int rabi_maf_rsample_fwd(rabi_maf_t* h, const float* A_dev, const float* z_dev, int S, int B,
                                    float* theta_dev, float* logq_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    PairArgs a = {};
    a.A_p = A_dev;
    a.in = z_dev;
    a.out_theta = theta_dev;
    a.out_lp = logq_dev;
    a.S = S;
    a.B = B;
    {
        FastArgs fa = {};
        fa.A_p = A_dev;
        fa.A_q = A_dev;
        fa.z = z_dev;
        fa.theta_out = theta_dev;
        fa.diff = logq_dev;
        fa.S = S;
        fa.B = B;
        int rc = 2;
        if (!env_int("RABI_FORCE_DEEP", 0)) {
            TcCall tc = {A_dev, A_dev, z_dev, logq_dev, theta_dev, nullptr, S, B, 0.f};
            rc = rabi_tc_launch(h, tc, TC_RSAMPLE, (cudaStream_t)s);
            if (rc == 2) rc = launch_fast<FAST_RSAMPLE>(h, fa, (cudaStream_t)s);
        }
        if (rc != 2) return rc;
        DeepArgs da = {};
        da.A_p = A_dev;
        da.A_q = A_dev;
        da.z = z_dev;
        da.theta_out = theta_dev;
        da.diff = logq_dev;
        da.S = S;
        da.B = B;
        rc = launch_deep(h, da, DEEP_RSAMPLE, (cudaStream_t)s);
        if (rc != 2) return rc;
    }
    return launch_pairs<MODE_RSAMPLE>(h, a, (cudaStream_t)s);
}

// ---------------------------------------------------------------------------------------------------------
// fast path launcher (maf_fast.cuh).  Returns 0 = launched, 1 = error, 2 = configuration not eligible (caller
// falls back to the generic per-pair kernel).

template <int MODE, int RP, int DP, bool TM, int LDC>
static int launch_fast_inst(rabi_maf* h, const FastArgs& fa, const FastRel& rel, int grid, int T, size_t smem,
                            cudaStream_t st) {
    CU_OK(h, cudaFuncSetAttribute(k_fast<MODE, RP, DP, TM, LDC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = h->profiling && fast_is_grad(MODE);
    if (prof) {
        if (h->prof_used == h->prof_events.size()) {
            cudaEvent_t e0, e1;
            CU_OK(h, cudaEventCreate(&e0));
            CU_OK(h, cudaEventCreate(&e1));
            h->prof_events.push_back({e0, e1});
        }
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].first, st));
    }
    k_fast<MODE, RP, DP, TM, LDC><<<grid, T, smem, st>>>(h->img, fa, rel);
    LAUNCH_CHECK(h);
    if (prof) {
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].second, st));
        h->prof_used++;
    }
    return 0;
}

template <int MODE>
static int launch_fast(rabi_maf* h, FastArgs fa, cudaStream_t st) {
    if ((long long)fa.S * fa.B <= 0) return 0;  // empty batch: nothing to launch
    if (env_int("RABI_DISABLE_FAST", 0)) return 2;
    if (h->L != 2 || h->D > 12) return 2;
    const MafImg& g = h->img;
    const TImg& t0 = g.t[0];
    const int K = h->K, D = h->D, H0 = h->H[0], H1 = h->H[1];
    const int DP = round4(D);
    FastRel rel;
    rel.W0t = 0;
    rel.b1 = t0.b[1] - t0.W0t;
    rel.W1 = t0.W[1] - t0.W0t;
    rel.ld1 = t0.ld[1];
    rel.WoT = t0.WoT - t0.W0t;
    rel.bo = t0.boF - t0.W0t;
    rel.perm = 0;
    rel.permy = t0.permy - t0.perm;
    rel.deg0 = t0.deg0 - t0.perm;
    rel.pre1 = t0.pre[1] - t0.perm;
    rel.grp0 = t0.grp[0] - t0.perm;
    rel.grp1 = t0.grp[1] - t0.perm;
    rel.first1 = t0.first[1] - t0.perm;
    fa.wfloats = t0.fast_end - t0.W0t;
    fa.wints = (t0.meta_fast_end - t0.perm + 3) & ~3;
    const int RP = 1;  // two pairs per thread (shared weight loads) was measured 1.8x slower: half the warps, more spills
    // activations in tensor memory (default) or in shared memory
    const bool TM = env_int("RABI_FAST_TMEM", 1) != 0 && RP * std::max(round4(H0), round4(H1)) <= 512 && (RP == 1 || DP == 4);
    int AS = std::max(round4(H0), round4(H1));
    if (!TM && ((AS / 4) & 1) == 0) AS += 4;  // AS/4 odd: 128-bit accesses of consecutive slots hit distinct bank groups
    fa.AS = AS;
    fa.VS = ((4 * K + 2) * D) | 1;
    const int WL = (g.Hmax + 31) / 32;
    fa.BWS = fast_is_grad(MODE) ? ((2 * K * 2 * WL) | 1) : 0;
    const size_t per_slot = (size_t)((TM ? 0 : fa.AS) + fa.VS + fa.BWS) * sizeof(float);
    const size_t fixed = (size_t)(fa.wfloats + fa.wints) * sizeof(float);
    const size_t budget = (size_t)h->max_smem_optin - 1024;
    // matches __launch_bounds__ of k_fast; the forward-only modes keep less state per thread and run 512-thread tiles
    // (a tile takes the time one thread needs for its pair, so wider tiles mean fewer waves)
    const int Tmax = RP != 1 ? 256 : (fast_is_grad(MODE) ? 384 : 512);
    int T = env_int("RABI_FAST_T", Tmax);
    T = std::max(32, std::min(Tmax, (T / 32) * 32));
    if (TM) {  // every group of 4 warps needs RP*AS of the 512 tensor-memory columns
        while (T > 32 && ((T / 32 + 3) / 4) * RP * AS + 4 > 512) T -= 32;  // + 4: the look-ahead quad of the product loops
    }
    int grid = 0;
    size_t smem = 0;
    const long long npairs = (long long)fa.S * fa.B;
    if (fast_is_grad(MODE)) {
        // largest T whose slots (+ the gA tile for the rows they hold) fit
        for (;; T -= 32) {
            if (T < 32) return 2;
            int slots = RP * T;
            int tb = slots / fa.S;
            if (tb < 1) return 2;  // one row's samples do not fit in a tile
            smem = fixed + (size_t)slots * per_slot;
            if (smem <= budget) break;
        }
        int tb_cap = (RP * T) / fa.S;
        int ntiles = (fa.B + tb_cap - 1) / tb_cap;
        int waves = (ntiles + h->sm_count - 1) / h->sm_count;
        int want = std::min(fa.B, waves * h->sm_count);           // balanced: every SM gets `waves` tiles
        fa.TB = std::min(tb_cap, (fa.B + want - 1) / want);
        fa.ntiles = (fa.B + fa.TB - 1) / fa.TB;
        int Tneed = ((fa.TB * fa.S + RP - 1) / RP + 31) & ~31;     // no idle warps
        T = std::min(T, std::max(32, Tneed));
        smem = fixed + (size_t)RP * T * per_slot;
        grid = std::min(fa.ntiles, h->sm_count);
    } else {
        for (;; T -= 32) {
            if (T < 32) return 2;
            smem = fixed + (size_t)RP * T * per_slot;
            if (smem <= budget) break;
        }
        fa.NP = RP * T;
        long long nt = (npairs + fa.NP - 1) / fa.NP;
        if (nt > 0x7fffffff) return 2;
        fa.ntiles = (int)nt;
        grid = (int)std::min<long long>(nt, h->sm_count);
    }
    fa.tm_cols = 32;
    if (TM) {
        int need = ((T / 32 + 3) / 4) * RP * AS + 4;
        while (fa.tm_cols < need) fa.tm_cols *= 2;
    }
#define RABI_FAST_CASE(RPv, DPv, TMv, LDv) \
    if (RP == RPv && DP == DPv && TM == TMv && (LDv == 0 || rel.ld1 == LDv))  \
        return launch_fast_inst<MODE, RPv, DPv, TMv, LDv>(h, fa, rel, grid, T, smem, st);
    RABI_FAST_CASE(1, 4, true, 100)   // lotka_volterra: hidden [100, 100]
#ifndef RABI_LEAN
    RABI_FAST_CASE(1, 12, true, 100)  // gaussian_linear
    RABI_FAST_CASE(1, 4, true, 0)
    RABI_FAST_CASE(1, 8, true, 0)
    RABI_FAST_CASE(1, 12, true, 0)
    RABI_FAST_CASE(1, 4, false, 0)
#endif
#undef RABI_FAST_CASE
    return 2;
}

// ---- k_fast2 (maf_fast2.cuh): the attack gradient with two lanes per pair, for launches that leave the SMs short of warps
template <int LDC>
static int launch_fast2_inst(rabi_maf* h, const FastArgs& fa, const FastRel& rel, int grid, int T, size_t smem, cudaStream_t st) {
    CU_OK(h, cudaFuncSetAttribute(k_fast2<LDC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = h->profiling;
    if (prof) {
        if (h->prof_used == h->prof_events.size()) {
            cudaEvent_t e0, e1;
            CU_OK(h, cudaEventCreate(&e0));
            CU_OK(h, cudaEventCreate(&e1));
            h->prof_events.push_back({e0, e1});
        }
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].first, st));
    }
    k_fast2<LDC><<<grid, T, smem, st>>>(h->img, fa, rel);
    LAUNCH_CHECK(h);
    if (prof) {
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].second, st));
        h->prof_used++;
    }
    return 0;
}

// returns 2 when the shape or the launch size is not one for k_fast2 (the caller then takes k_fast)
static int launch_fast2(rabi_maf* h, FastArgs fa, cudaStream_t st) {
    if ((long long)fa.S * fa.B <= 0) return 0;
    if (!env_int("RABI_FAST2", 1) || env_int("RABI_DISABLE_FAST", 0)) return 2;
    if (h->L != 2 || h->D > 4 || h->H[0] > 104 || h->H[1] > 104) return 2;
    const MafImg& g = h->img;
    const TImg& t0 = g.t[0];
    const int K = h->K, D = h->D, H0 = h->H[0], H1 = h->H[1];
    // pays only while a thread per pair leaves the schedulers nearly empty (measured on lotka_volterra shapes after k_fast's
    // own trimming: 2-3 % below ~40 pairs per SM, even at ~150, 14 % slower with the SMs full); RABI_FAST2=2 forces it
    if (env_int("RABI_FAST2", 1) != 2 && (long long)fa.S * fa.B > (long long)h->sm_count * env_int("RABI_FAST2_MAX_PER_SM", 40)) return 2;
    if ((long long)fa.S * fa.B > (long long)h->sm_count * 352) return 2;  // one wave of 704-thread tiles
    FastRel rel;
    rel.W0t = 0;
    rel.b1 = t0.b[1] - t0.W0t;
    rel.W1 = t0.W[1] - t0.W0t;
    rel.ld1 = t0.ld[1];
    rel.WoT = t0.WoT - t0.W0t;
    rel.bo = t0.boF - t0.W0t;
    rel.perm = 0;
    rel.permy = t0.permy - t0.perm;
    rel.deg0 = t0.deg0 - t0.perm;
    rel.pre1 = t0.pre[1] - t0.perm;
    rel.grp0 = t0.grp[0] - t0.perm;
    rel.grp1 = t0.grp[1] - t0.perm;
    rel.first1 = t0.first[1] - t0.perm;
    if (t0.ldt != 4) return 2;
    fa.wfloats = t0.fast_end - t0.W0t;
    fa.wints = (t0.meta_fast_end - t0.perm + 3) & ~3;
    const int nq = (std::max(H0, H1) + 3) / 4;
    fa.AS = 4 * ((nq + 1) / 2);  // own TMEM columns of a lane
    fa.VS = ((4 * K + 2) * D) | 1;
    const int WL = (g.Hmax + 31) / 32;
    fa.BWS = (2 * K * 2 * WL) | 1;
    const size_t per_slot = (size_t)(fa.VS + fa.BWS) * sizeof(float);
    const size_t fixed = (size_t)(fa.wfloats + fa.wints) * sizeof(float);
    const size_t budget = (size_t)h->max_smem_optin - 1024;
    int T = 704;
    while (T > 64 && ((T / 32 + 3) / 4) * fa.AS + 4 > 512) T -= 64;  // + 4: the look-ahead quad of the product loops
    size_t smem = 0;
    for (;; T -= 64) {
        if (T < 64) return 2;
        const int tb = (T / 2) / fa.S;
        if (tb < 1) return 2;
        smem = fixed + (size_t)(T / 2) * per_slot;
        if (smem <= budget) break;
    }
    const int tb_cap = (T / 2) / fa.S;
    const int ntiles0 = (fa.B + tb_cap - 1) / tb_cap;
    const int waves = (ntiles0 + h->sm_count - 1) / h->sm_count;
    const int want = std::min(fa.B, waves * h->sm_count);
    fa.TB = std::min(tb_cap, (fa.B + want - 1) / want);
    fa.ntiles = (fa.B + fa.TB - 1) / fa.TB;
    const int Tneed = (2 * fa.TB * fa.S + 63) & ~63;
    T = std::min(T, std::max(64, Tneed));
    smem = fixed + (size_t)(T / 2) * per_slot;
    const int grid = std::min(fa.ntiles, h->sm_count);
    fa.tm_cols = 32;
    const int need = ((T / 32 + 3) / 4) * fa.AS + 4;
    while (fa.tm_cols < need) fa.tm_cols *= 2;
    if (rel.ld1 == 100) return launch_fast2_inst<100>(h, fa, rel, grid, T, smem, st);
    return launch_fast2_inst<0>(h, fa, rel, grid, T, smem, st);
}

// ---- deep / wide conditioners: per-(transform, pass) launches over a global-memory scratch (maf_deep.cuh) ----------
// MAXT: __launch_bounds__ of the kernel variant (384: 168 registers per thread; 512: 128).  A batch runs in the time one
// thread needs for its pair, so the launcher minimises the NUMBER of batches first and only then picks the variant.
template <int DP, int MAXT>
static int launch_deep_dp(rabi_maf* h, DeepArgs a, int mode, long long N, cudaStream_t st) {
    const int K = h->K, D = h->D, L = h->L;
    const MafImg& g = h->img;
    const std::vector<int>& M = h->meta_host;
    auto r4 = [](int x) { return (x + 3) & ~3; };
    // shared-memory needs (floats)
    size_t full = (size_t)r4(h->H[0]) * g.t[0].ldt;
    for (int l = 1; l < L; ++l) full = std::max(full, (size_t)r4(h->H[l]) * g.t[0].ld[l]);
    full = std::max(full, (size_t)r4(2 * D) * g.t[0].ldo) + 64;
    size_t seqf = 0, seqb = 0;
    for (int k = 0; k < K; ++k) {
        const TImg& t = g.t[k];
        for (int j = 0; j < D; ++j) {
            size_t f = 2 * (size_t)DEEP_SLD, b = 0;
            for (int l = 0; l < L; ++l) {
                const int ua = M[t.grp[l] + j], ub = M[t.grp[l] + j + 1];
                f += l == 0 ? (size_t)(ub - ua) * t.ldt : (size_t)((ub - ua + 7) & ~7) * DEEP_SLD;
                const int nblk = ub > ua ? (ub - (ua & ~3) + 15) >> 4 : 0;
                const int rows = l == L - 1 ? r4(2 * D) : r4(h->H[l + 1]);
                const int r0 = l == L - 1 ? 0 : (M[t.grp[l + 1] + j] & ~3);
                const int n16 = (std::max(rows - r0, 0) + 15) & ~15;
                b += (size_t)nblk * n16 * DEEP_SW;
                if (l == 0) b += (size_t)(ub - ua) * t.ldt;
            }
            seqf = std::max(seqf, f);
            seqb = std::max(seqb, b);
            if (j + 1 < D) {  // paired step of kd_seq_fwd (groups of <= 8 units): W0t rows of both groups, per hidden
                              // layer 16 sweep rows + an 8 x 8 diagonal, 4 output rows + a 2 x 8 diagonal
                size_t fp = (size_t)(M[t.grp[0] + j + 2] - M[t.grp[0] + j]) * t.ldt + (size_t)(L - 1) * (16 * DEEP_SLD + 64) +
                            4 * DEEP_SLD + 16;
                seqf = std::max(seqf, fp);
            }
        }
    }
    seqf += 64;
    seqb += 64;
    const size_t budget = (size_t)h->max_smem_optin - 1024;
    if (full * 4 > budget || seqf * 4 > budget || seqb * 4 > budget) return 2;
    CU_OK(h, cudaFuncSetAttribute(kd_seq_fwd<DP, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(seqf * 4)));
    CU_OK(h, cudaFuncSetAttribute(kd_full_fwd<DP, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(full * 4)));
    CU_OK(h, cudaFuncSetAttribute(kd_full_bwd<DP, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(full * 4)));
    CU_OK(h, cudaFuncSetAttribute(kd_seq_bwd<DP, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(seqb * 4)));
    // batches: at most one CTA of <= Tmax threads per SM, every SM gets the same number of pairs
    const int Tmax = std::max(64, std::min(MAXT, env_int("RABI_DEEP_T", MAXT) & ~31));
    const long long cap = (long long)h->sm_count * Tmax;
    const long long nbatch = (N + cap - 1) / cap;
    const int nmax = (int)((N + nbatch - 1) / nbatch);
    int T = (int)((((nmax + h->sm_count - 1) / h->sm_count) + 31) & ~31);
    T = std::max(64, std::min(Tmax, T));
    const int grid_max = (nmax + T - 1) / T;
    const int Np = grid_max * T;
    const int WL = (g.Hmax + 31) / 32;
    const int Hq = (r4(g.Hmax) + DEEP_SLACK) / 4, Oq = (r4(2 * D) + DEEP_SLACK) / 4;
    const size_t n_act = (size_t)L * Hq * Np * 4, n_ob = (size_t)Oq * Np * 4, n_vec = (size_t)deep_nvec(K) * D * Np;
    const size_t n_bits = (size_t)2 * K * L * WL * Np, n_lp = (size_t)2 * Np;
    const size_t bytes = (n_act + n_ob + n_vec + n_bits + n_lp) * sizeof(float);
    if (bytes > h->deep_ws_bytes) {
        if (h->deep_ws) {
            CU_OK(h, cudaDeviceSynchronize());
            CU_OK(h, cudaFree(h->deep_ws));
            h->deep_ws = nullptr;
            h->deep_ws_bytes = 0;
        }
        CU_OK(h, cudaMalloc(&h->deep_ws, bytes));
        h->deep_ws_bytes = bytes;
        CU_OK(h, cudaMemsetAsync(h->deep_ws, 0, bytes, st));
    }
    // stale activations are only ever multiplied by exact-zero (masked / padded) weights, so they must be finite:
    // clear them per call (a diverged earlier call may have left NaN behind)
    CU_OK(h, cudaMemsetAsync(h->deep_ws, 0, (n_act + n_ob) * sizeof(float), st));
    a.act = (float*)h->deep_ws;
    a.obuf = a.act + n_act;
    a.vec = a.obuf + n_ob;
    a.bits = (uint32_t*)(a.vec + n_vec);
    a.lp = (float*)(a.bits + n_bits);
    a.Np = Np;
    a.Hq = Hq;
    a.Oq = Oq;
    a.record = mode == DEEP_GRAD || mode == DEEP_FKL;
    const bool prof = h->profiling && a.record;
    if (prof) {
        if (h->prof_used == h->prof_events.size()) {
            cudaEvent_t e0, e1;
            CU_OK(h, cudaEventCreate(&e0));
            CU_OK(h, cudaEventCreate(&e1));
            h->prof_events.push_back({e0, e1});
        }
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].first, st));
    }
    for (long long p0 = 0; p0 < N; p0 += nmax) {
        a.p0 = p0;
        a.n = (int)std::min<long long>(nmax, N - p0);
        const int grid = (a.n + T - 1) / T;
        if (mode == DEEP_LOGPROB) {
            kd_load_theta<<<grid, T, 0, st>>>(g, a);
            LAUNCH_CHECK(h);
        } else {
            DeepArgs as = a;
            if (mode == DEEP_FKL) as.record = 0;  // the sampling pass is not differentiated
            for (int k = 0; k < K; ++k) {
                kd_seq_fwd<DP, MAXT><<<grid, T, seqf * 4, st>>>(g, as, k);
                LAUNCH_CHECK(h);
            }
        }
        if (mode == DEEP_RSAMPLE) {
            kd_store_theta<<<grid, T, 0, st>>>(g, a);
            LAUNCH_CHECK(h);
            continue;
        }
        for (int k = K - 1; k >= 0; --k) {
            kd_full_fwd<DP, MAXT><<<grid, T, full * 4, st>>>(g, a, k, mode);
            LAUNCH_CHECK(h);
        }
        if (mode == DEEP_GRAD || mode == DEEP_FKL) {
            for (int k = 0; k < K; ++k) {
                kd_full_bwd<DP, MAXT><<<grid, T, full * 4, st>>>(g, a, k, mode == DEEP_FKL);
                LAUNCH_CHECK(h);
            }
        }
        if (mode == DEEP_GRAD) {
            for (int k = K - 1; k >= 0; --k) {
                kd_seq_bwd<DP, MAXT><<<grid, T, seqb * 4, st>>>(g, a, k);
                LAUNCH_CHECK(h);
            }
        }
    }
    if (prof) {
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].second, st));
        h->prof_used++;
    }
    return 0;
}

// returns 0 = launched, 1 = error, 2 = not applicable (caller falls back to the generic per-pair kernel)
static int launch_deep(rabi_maf* h, const DeepArgs& a, int mode, cudaStream_t st) {
    const long long N = (long long)a.S * a.B;
    if (N <= 0) return 0;
    if (env_int("RABI_DISABLE_DEEP", 0)) return 2;
#ifdef RABI_LEAN  // development builds (kernel experiments): only the lotka_volterra k_fast instantiations, 1 min instead of 6
    return 2;
#else
    if (h->L < 2 || h->D > 32 || ((h->img.Hmax + 15) & ~15) > DEEP_SLD) return 2;
    // fewest batches: one CTA per SM of up to 512 threads; the 384-thread variant (more registers) when it needs no more
    const long long nb512 = (N + (long long)h->sm_count * 512 - 1) / ((long long)h->sm_count * 512);
    const long long nb384 = (N + (long long)h->sm_count * 384 - 1) / ((long long)h->sm_count * 384);
    const bool big = nb512 < nb384 && env_int("RABI_DEEP_T", 512) > 384;
    if (h->D <= 12) return big ? launch_deep_dp<12, 512>(h, a, mode, N, st) : launch_deep_dp<12, 384>(h, a, mode, N, st);
    return big ? launch_deep_dp<32, 512>(h, a, mode, N, st) : launch_deep_dp<32, 384>(h, a, mode, N, st);
#endif
}

extern "C" int rabi_rkl_fwd(rabi_maf_t* h, const float* A_p_dev, const float* A_q_dev, const float* z_dev, int S, int B,
                            float* kl_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (S <= 0 || B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    if (ensure_ws(h, (size_t)S * B * sizeof(float))) return 1;
    PairArgs a = {};
    a.A_p = A_p_dev;
    a.A_q = A_q_dev;
    a.in = z_dev;
    a.out_lp = (float*)h->ws;
    a.S = S;
    a.B = B;
    {
        FastArgs fa = {};
        fa.A_p = A_p_dev;
        fa.A_q = A_q_dev;
        fa.z = z_dev;
        fa.diff = (float*)h->ws;
        fa.S = S;
        fa.B = B;
        int rc = 2;
        if (!env_int("RABI_FORCE_DEEP", 0)) {
            TcCall tc = {A_p_dev, A_q_dev, z_dev, (float*)h->ws, nullptr, nullptr, S, B, 0.f};
            rc = rabi_tc_launch(h, tc, TC_FWD, st);
            if (rc == 2) rc = launch_fast<FAST_FWD>(h, fa, st);
        }
        if (rc == 1) return 1;
        if (rc == 2) {
            DeepArgs da = {};
            da.A_p = A_p_dev;
            da.A_q = A_q_dev;
            da.z = z_dev;
            da.diff = (float*)h->ws;
            da.S = S;
            da.B = B;
            rc = launch_deep(h, da, DEEP_FWD, st);
            if (rc == 1) return 1;
        }
        if (rc == 2 && launch_pairs<MODE_RKL_FWD>(h, a, st)) return 1;
    }
    k_mean_over_s<<<(B + 127) / 128, 128, 0, st>>>((const float*)h->ws, S, B, kl_dev);
    LAUNCH_CHECK(h);
    return 0;
}

// ---- small batches: one warp per pair (maf_warp.cuh).  Returns 0 = launched, 1 = error, 2 = not applicable.
template <int MODE>
static int launch_warp(rabi_maf* h, const PairArgs& a, cudaStream_t st) {
    const long long N = (long long)a.S * a.B;
    if (N <= 0) return 0;
    if (N > env_int("RABI_WARP_MAX_PAIRS", 1024)) return 2;
    // warps per CTA: about one CTA per SM (every CTA stages the weights once)
    const int wpc = (int)std::max<long long>(2, std::min<long long>(8, (N + h->sm_count - 1) / h->sm_count));
    const size_t per = (warp_smem_floats(h->K, h->D, h->L, h->img.Hmax) + 3) & ~(size_t)3;
    // weights [W0t .. bo] of all K transforms resident if they fit, else one transform at a time, else global memory
    const TImg& t0 = h->img.t[0];
    int wblock = (t0.bo + round4(2 * h->D) + 4 - t0.W0t + 3) & ~3;
    const int nmeta = ((int)h->n_int + 3) & ~3;
    const size_t budget = (size_t)h->max_smem_optin - 2048;
    int wall = 1;
    if (env_int("RABI_WARP_STAGE", 1) == 0) {
        wblock = 0;
        wall = 0;
    } else if (((size_t)wblock * h->K + nmeta + wpc * per) * sizeof(float) > budget) {
        wall = 0;
        if (((size_t)wblock + nmeta + wpc * per) * sizeof(float) > 100 * 1024) wblock = 0;
    }
    const size_t smem = ((size_t)wblock * (wall ? h->K : 1) + nmeta + wpc * per) * sizeof(float);
    if (smem > (size_t)h->max_smem_optin) return 2;
    CU_OK(h, cudaFuncSetAttribute(k_warp<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = h->profiling;
    if (prof) {
        if (h->prof_used == h->prof_events.size()) {
            cudaEvent_t e0, e1;
            CU_OK(h, cudaEventCreate(&e0));
            CU_OK(h, cudaEventCreate(&e1));
            h->prof_events.push_back({e0, e1});
        }
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].first, st));
    }
    k_warp<MODE><<<(unsigned)((N + wpc - 1) / wpc), wpc * 32, smem, st>>>(h->img, a, wblock, wall, nmeta);
    LAUNCH_CHECK(h);
    if (prof) {
        CU_OK(h, cudaEventRecord(h->prof_events[h->prof_used].second, st));
        h->prof_used++;
    }
    return 0;
}

// workspace layout for the PGD step: [A_p: K*H0*B][gA: K*H0*B][gx: B*C][diff: S*B]
static int pgd_step_impl(rabi_maf* h, const float* x, float* delta, const float* zm, const float* zs, const float* A_q,
                         const float* z, int S, int B, float eps, float eps_iter, float cmin, float cmax, float inv_B,
                         float floor_, float* kl, cudaStream_t st, int forward_kl = 0) {
    const size_t nA = (size_t)h->K * h->H[0] * B;
    float* A_p = (float*)h->ws;
    float* gA = A_p + nA;
    float* gx = gA + nA;
    float* diff = gx + (size_t)B * h->C;
    if (launch_ctx_project(h, x, delta, zm, zs, B, A_p, st)) return 1;
    FastArgs fa = {};
    fa.A_p = A_p;
    fa.A_q = A_q;
    fa.z = z;
    fa.diff = kl ? diff : nullptr;
    fa.gA = gA;
    fa.S = S;
    fa.B = B;
    fa.w = inv_B / (float)S;
    int rc;
    // every kernel family adds the S samples' parts of d/dA_p into gA (float reductions in global memory)
    CU_OK(h, cudaMemsetAsync(gA, 0, nA * sizeof(float), st));
    if (!env_int("RABI_FORCE_DEEP", 0) && (long long)S * B >= env_int("RABI_TC_GRAD_MIN_PAIRS", 1025)) {
        // tensor-core path (maf_tc.cu): two-hidden-layer conditioners up to 120 (padded) units wide; launches of up to 1024
        // pairs stay on the warp-per-pair kernel (measured: 512 pairs 183 us there, 212 us here; 1500 pairs 320 vs 216 us)
        TcCall tc = {forward_kl ? A_q : A_p, forward_kl ? A_p : A_q, z, kl ? diff : nullptr, nullptr, gA, S, B, inv_B / (float)S};
        rc = rabi_tc_launch(h, tc, forward_kl ? TC_FKL : TC_GRAD, st);
        if (rc == 1) return 1;
        if (rc == 0) goto grad_done;
    }
    {   // small batches (training-time attacks): one warp per pair
        PairArgs wa = {};
        wa.A_p = forward_kl ? A_q : A_p;   // forward KL: samples come from the fixed target
        wa.A_q = forward_kl ? A_p : A_q;
        wa.in = z;
        wa.out_lp = kl ? diff : nullptr;
        wa.gA = gA;
        wa.S = S;
        wa.B = B;
        wa.w = inv_B / (float)S;
        if ((long long)S * B <= env_int("RABI_WARP_MAX_PAIRS", 1024) && !env_int("RABI_FORCE_DEEP", 0)) {
            rc = forward_kl ? launch_warp<WARP_FKL_GRAD>(h, wa, st) : launch_warp<WARP_RKL_GRAD>(h, wa, st);
            if (rc == 1) return 1;
            if (rc == 0) goto grad_done;
        }
    }
    if (forward_kl) {  // KL(target || q(.|x+delta)): sample from the target's projection, differentiate the density pass
        fa.A_p = A_q;
        fa.A_q = A_p;
        rc = env_int("RABI_FORCE_DEEP", 0) ? 2 : launch_fast<FAST_FKL>(h, fa, st);
        if (rc == 2) {
            DeepArgs da = {};
            da.A_p = A_q;   // samples come from the fixed target
            da.A_q = A_p;   // the density pass under x + delta is the differentiated one
            da.z = z;
            da.diff = kl ? diff : nullptr;
            da.gA = gA;
            da.S = S;
            da.B = B;
            da.w = inv_B / (float)S;
            rc = launch_deep(h, da, DEEP_FKL, st);
            if (rc == 2) return fail(h, "the forward-KL attack loss needs the fast or the deep path (2..4 hidden layers, D <= 32)");
            if (rc == 0) rc = 3;
        }
    } else {
        rc = env_int("RABI_FORCE_DEEP", 0) ? 2 : launch_fast2(h, fa, st);
        if (rc == 2 && !env_int("RABI_FORCE_DEEP", 0)) rc = launch_fast<FAST_GRAD>(h, fa, st);
    }
    if (rc == 1) return 1;
    if (rc == 2) {  // deep path (global-memory scratch), gA accumulated with atomics
        DeepArgs da = {};
        da.A_p = A_p;
        da.A_q = A_q;
        da.z = z;
        da.diff = kl ? diff : nullptr;
        da.gA = gA;
        da.S = S;
        da.B = B;
        da.w = inv_B / (float)S;
        rc = launch_deep(h, da, DEEP_GRAD, st);
        if (rc == 1) return 1;
        if (rc == 0) rc = 3;  // done
    }
    if (rc == 2) {  // generic per-pair kernel (accumulates gA with atomics)
        PairArgs a = {};
        a.A_p = A_p;
        a.A_q = A_q;
        a.in = z;
        a.out_lp = kl ? diff : nullptr;
        a.gA = gA;
        a.S = S;
        a.B = B;
        a.w = inv_B / (float)S;
        if (launch_pairs<MODE_RKL_GRAD>(h, a, st)) return 1;
    }
grad_done:
    if (launch_ctx_project_bwd(h, gA, zs, B, gx, 0, st)) return 1;
    if (!delta) {  // gradient-only mode (rabi_rkl_input_grad)
        if (kl) {
            k_mean_over_s<<<(B + 127) / 128, 128, 0, st>>>(diff, S, B, kl);
            LAUNCH_CHECK(h);
        }
        return 0;
    }
    int threads = 128, rows_per_block = threads / 32;
    k_pgd_update<<<(B + rows_per_block - 1) / rows_per_block, threads, 0, st>>>(x, delta, gx, B, h->C, eps, eps_iter,
                                                                                   cmin, cmax, floor_, h->row_eps, h->row_eps_iter, h->col_mask);
    LAUNCH_CHECK(h);
    if (kl) {
        k_mean_over_s<<<(B + 127) / 128, 128, 0, st>>>(diff, S, B, kl);
        LAUNCH_CHECK(h);
    }
    return 0;
}

static size_t pgd_ws_bytes(const rabi_maf* h, int S, int B) {
    return (2 * (size_t)h->K * h->H[0] * B + (size_t)B * h->C + (size_t)S * B) * sizeof(float);
}

extern "C" int rabi_rkl_pgd_step(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* zs_mean_dev,
                                 const float* zs_std_dev, const float* A_q_dev, const float* z_dev, int S, int B,
                                 float eps, float eps_iter, float clip_min, float clip_max, float inv_B,
                                 float norm_floor, float* kl_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (S <= 0 || B <= 0) return 0;
    if (ensure_ws(h, pgd_ws_bytes(h, S, B))) return 1;
    return pgd_step_impl(h, x_dev, delta_dev, zs_mean_dev, zs_std_dev, A_q_dev, z_dev, S, B, eps, eps_iter, clip_min,
                         clip_max, inv_B, norm_floor, kl_dev, (cudaStream_t)s);
}

extern "C" int rabi_pgd_set_row_scales(rabi_maf_t* h, const float* eps_rows_dev, const float* eps_iter_rows_dev) {
    if (!h) return 1;
    if ((eps_rows_dev == nullptr) != (eps_iter_rows_dev == nullptr))
        return fail(h, "rabi_pgd_set_row_scales: pass both arrays or NULL, NULL");
    h->row_eps = eps_rows_dev;
    h->row_eps_iter = eps_iter_rows_dev;
    return 0;
}

extern "C" int rabi_pgd_set_column_mask(rabi_maf_t* h, const float* col_mask_dev) {
    if (!h) return 1;
    h->col_mask = col_mask_dev;
    return 0;
}

extern "C" int rabi_pgd_update(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* grad_dev, int B, int dx,
                               float eps, float eps_iter, float clip_min, float clip_max, float norm_floor,
                               rabi_stream_t s) {
    REQUIRE_READY(h);
    if (B <= 0 || dx <= 0) return 0;
    int threads = 128, rows_per_block = threads / 32;
    k_pgd_update<<<(B + rows_per_block - 1) / rows_per_block, threads, 0, (cudaStream_t)s>>>(
        x_dev, delta_dev, grad_dev, B, dx, eps, eps_iter, clip_min, clip_max, norm_floor, h->row_eps, h->row_eps_iter, h->col_mask);
    LAUNCH_CHECK(h);
    return 0;
}

extern "C" int rabi_rkl_input_grad(rabi_maf_t* h, const float* x_dev, const float* zs_mean_dev, const float* zs_std_dev,
                                   const float* A_q_dev, const float* z_dev, int S, int B, float inv_B, float* gx_dev,
                                   float* kl_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (S <= 0 || B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    if (ensure_ws(h, pgd_ws_bytes(h, S, B))) return 1;
    if (pgd_step_impl(h, x_dev, nullptr, zs_mean_dev, zs_std_dev, A_q_dev, z_dev, S, B, -1.f, 0.f, 0.f, 0.f, inv_B, 0.f,
                      kl_dev, st))
        return 1;
    const float* gx = (const float*)h->ws + 2 * (size_t)h->K * h->H[0] * B;
    CU_OK(h, cudaMemcpyAsync(gx_dev, gx, (size_t)B * h->C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int rabi_fkl_input_grad(rabi_maf_t* h, const float* x_dev, const float* zs_mean_dev, const float* zs_std_dev,
                                   const float* A_t_dev, const float* z_dev, int S, int B, float inv_B, float* gx_dev,
                                   float* kl_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (S <= 0 || B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    if (ensure_ws(h, pgd_ws_bytes(h, S, B))) return 1;
    if (pgd_step_impl(h, x_dev, nullptr, zs_mean_dev, zs_std_dev, A_t_dev, z_dev, S, B, -1.f, 0.f, 0.f, 0.f, inv_B, 0.f,
                      kl_dev, st, 1))
        return 1;
    const float* gx = (const float*)h->ws + 2 * (size_t)h->K * h->H[0] * B;
    CU_OK(h, cudaMemcpyAsync(gx_dev, gx, (size_t)B * h->C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// The whole attack as ONE persistent launch where the tensor-core path supports it (maf_tc.cu: rabi_tc_run); 2 = not eligible,
// the caller then runs the per-iteration launch groups.  Workspace as in pgd_step_impl: [A_p][gA][gx][diff].
static int pgd_run_fused(rabi_maf* h, const float* x, float* delta, const float* zm, const float* zs, const float* A_fixed,
                         const float* z_all, int nb_iter, int S, int B, float eps, float eps_iter, float cmin, float cmax,
                         float inv_B, float floor_, cudaStream_t st, int forward_kl) {
    if ((zm == nullptr) != (zs == nullptr)) return fail(h, "zs_mean and zs_std must both be given or both be NULL");
    if (env_int("RABI_FORCE_DEEP", 0)) return 2;
    const size_t nA = (size_t)h->K * h->H[0] * B;
    float* A_pert = (float*)h->ws;
    float* gA = A_pert + nA;
    CU_OK(h, cudaMemsetAsync(gA, 0, nA * sizeof(float), st));
    rabi::TcRunCall r = {};
    r.A_p = forward_kl ? A_fixed : A_pert;
    r.A_q = forward_kl ? A_pert : A_fixed;
    r.z_all = z_all;
    r.gA = gA;
    r.x = x;
    r.delta = delta;
    r.zmean = zm;
    r.zstd = zs;
    r.nb_iter = nb_iter;
    r.S = S;
    r.B = B;
    r.eps = eps;
    r.eps_iter = eps_iter;
    r.cmin = cmin;
    r.cmax = cmax;
    r.inv_B = inv_B;
    r.floor_ = floor_;
    r.eps_rows = h->row_eps;
    r.eps_iter_rows = h->row_eps_iter;
    r.col_mask = h->col_mask;
    return rabi::rabi_tc_run(h, r, forward_kl ? rabi::TC_FKL : rabi::TC_GRAD, st);
}

extern "C" int rabi_fkl_pgd_run(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* zs_mean_dev,
                                const float* zs_std_dev, const float* A_t_dev, const float* z_all_dev, int nb_iter, int S,
                                int B, float eps, float eps_iter, float clip_min, float clip_max, float inv_B,
                                float norm_floor, float* x_adv_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (S <= 0 || B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const size_t nA = (size_t)h->K * h->H[0] * B;
    if (ensure_ws(h, pgd_ws_bytes(h, S, B) + nA * sizeof(float))) return 1;
    const float* A_t = A_t_dev;
    if (!A_t) {
        float* own = (float*)((char*)h->ws + pgd_ws_bytes(h, S, B));
        if (launch_ctx_project(h, x_dev, nullptr, zs_mean_dev, zs_std_dev, B, own, st)) return 1;
        A_t = own;
    }
    const size_t zstride = (size_t)S * B * h->D;
    int fused = pgd_run_fused(h, x_dev, delta_dev, zs_mean_dev, zs_std_dev, A_t, z_all_dev, nb_iter, S, B, eps, eps_iter, clip_min,
                              clip_max, inv_B, norm_floor, st, 1);
    if (fused == 1) return 1;
    for (int it = 0; fused == 2 && it < nb_iter; ++it)
        if (pgd_step_impl(h, x_dev, delta_dev, zs_mean_dev, zs_std_dev, A_t, z_all_dev + (size_t)it * zstride, S, B, eps,
                          eps_iter, clip_min, clip_max, inv_B, norm_floor, nullptr, st, 1))
            return 1;
    if (x_adv_dev) {
        size_t n = (size_t)B * h->C;
        k_clamp_add<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_dev, delta_dev, n, clip_min, clip_max, x_adv_dev);
        LAUNCH_CHECK(h);
    }
    return 0;
}

extern "C" int rabi_rkl_pgd_run(rabi_maf_t* h, const float* x_dev, float* delta_dev, const float* zs_mean_dev,
                                const float* zs_std_dev, const float* A_q_dev, const float* z_all_dev, int nb_iter, int S,
                                int B, float eps,
                                float eps_iter, float clip_min, float clip_max, float inv_B, float norm_floor,
                                float* x_adv_dev, rabi_stream_t s) {
    REQUIRE_READY(h);
    if (S <= 0 || B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const size_t nA = (size_t)h->K * h->H[0] * B;
    if (ensure_ws(h, pgd_ws_bytes(h, S, B) + nA * sizeof(float))) return 1;
    const float* A_q = A_q_dev;
    if (!A_q) {
        float* own = (float*)((char*)h->ws + pgd_ws_bytes(h, S, B));
        if (launch_ctx_project(h, x_dev, nullptr, zs_mean_dev, zs_std_dev, B, own, st)) return 1;
        A_q = own;
    }
    const size_t zstride = (size_t)S * B * h->D;
    int fused = pgd_run_fused(h, x_dev, delta_dev, zs_mean_dev, zs_std_dev, A_q, z_all_dev, nb_iter, S, B, eps, eps_iter, clip_min,
                              clip_max, inv_B, norm_floor, st, 0);
    if (fused == 1) return 1;
    for (int it = 0; fused == 2 && it < nb_iter; ++it)
        if (pgd_step_impl(h, x_dev, delta_dev, zs_mean_dev, zs_std_dev, A_q, z_all_dev + (size_t)it * zstride, S, B, eps,
                          eps_iter, clip_min, clip_max, inv_B, norm_floor, nullptr, st))
            return 1;
    if (x_adv_dev) {
        size_t n = (size_t)B * h->C;
        k_clamp_add<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_dev, delta_dev, n, clip_min, clip_max, x_adv_dev);
        LAUNCH_CHECK(h);
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------------------
// General fp32 GEMM primitive for the differentiable (training-side) path: C[M,N] = op(A) * op(B), row-major with
// leading dimensions (so sliced views of the masked weights / activation prefixes need no copies).  64x64 tile,
// 16-deep K chunks in shared memory, 4x4 outputs per thread.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, const float* __restrict__ A, int lda,
                                               const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                               float beta) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // split-K: blockIdx.z owns the k-range [kz0, kz1); partial tiles are combined with atomicAdd into a zeroed C
    const int kper = ((K + gridDim.z - 1) / gridDim.z + BK - 1) / BK * BK;
    const int kz0 = blockIdx.z * kper, kz1 = min(K, kz0 + kper);
    for (int k0 = kz0; k0 < kz1; k0 += BK) {
        for (int i = threadIdx.x; i < BM * BK; i += 256) {
            int mm, kk;
            if (TA) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < kz1) ? (TA ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]) : 0.f;
        }
        for (int i = threadIdx.x; i < BN * BK; i += 256) {
            int nn, kk;
            if (TB) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < kz1) ? (TB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * ty + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tx + j;
            if (n < N) {
                float* c = C + (size_t)m * ldc + n;
                if (gridDim.z > 1) atomicAdd(c, acc[i][j]);
                else *c = beta != 0.f ? beta * *c + acc[i][j] : acc[i][j];
            }
        }
    }
}

static unsigned long long g_sgemm_launches = 0;
extern "C" unsigned long long rabi_sgemm_launch_count(void) { return g_sgemm_launches; }

extern "C" int rabi_sgemm(int device, int transA, int transB, int M, int N, int K, const float* A_dev, int lda,
                          const float* B_dev, int ldb, float* C_dev, int ldc, float beta, rabi_stream_t s) {
    if (M <= 0 || N <= 0) return 0;
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, "cudaSetDevice(%d) failed", device);
    cudaStream_t st = (cudaStream_t)s;
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    // few output tiles and a long reduction (weight gradients: the batch is the K dimension): split K over blocks
    if (beta == 0.f && grid.x * grid.y <= 16 && K >= 512) {
        grid.z = std::min(64, (K + 127) / 128);
        for (int m = 0; m < M; ++m)
            if (ldc == N) {
                if (cudaMemsetAsync(C_dev, 0, (size_t)M * N * sizeof(float), st) != cudaSuccess) return fail(nullptr, "memset failed");
                break;
            } else if (cudaMemsetAsync(C_dev + (size_t)m * ldc, 0, (size_t)N * sizeof(float), st) != cudaSuccess)
                return fail(nullptr, "memset failed");
    }
    if (transA && transB) k_sgemm<true, true><<<grid, 256, 0, st>>>(M, N, K, A_dev, lda, B_dev, ldb, C_dev, ldc, beta);
    else if (transA) k_sgemm<true, false><<<grid, 256, 0, st>>>(M, N, K, A_dev, lda, B_dev, ldb, C_dev, ldc, beta);
    else if (transB) k_sgemm<false, true><<<grid, 256, 0, st>>>(M, N, K, A_dev, lda, B_dev, ldb, C_dev, ldc, beta);
    else k_sgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, A_dev, lda, B_dev, ldb, C_dev, ldc, beta);
    g_sgemm_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, "sgemm launch failed: %s", cudaGetErrorString(e));
    return 0;
}
