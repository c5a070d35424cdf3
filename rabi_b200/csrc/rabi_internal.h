// Host-side state behind a rabi_maf_t handle, shared by the translation units of librabi_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/rabi_b200.h"
#include "maf_image.h"

struct rabi_maf {
    int device = 0;
    int K = 0, D = 0, C = 0, L = 0;
    int H[RABI_MAXL] = {0, 0, 0, 0};
    float lo = -5.f, hi = 3.f;
    bool finalized = false;
    std::vector<std::vector<int64_t>> perm;               // [K][D]
    std::vector<std::vector<int64_t>> post;               // [K][D] optional Permute layer after transform k
    std::vector<std::vector<std::vector<float>>> masks;   // [K][L+1][out*in]
    MafImg img;                                           // device pointers inside
    std::vector<int> meta_host;
    float* f_dev = nullptr;
    int* m_dev = nullptr;
    size_t n_float = 0, n_int = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    void* deep_ws = nullptr;   // per-pair scratch of the deep path (maf_deep.cuh), zero-initialised on allocation
    size_t deep_ws_bytes = 0;
    unsigned long long launches = 0;
    // optional CUDA-event timing of the dominant kernel on its launch stream
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    std::string err;
    int sm_count = 148;
    int max_smem_optin = 0;
    void* tc = nullptr;        // tensor-core path state (maf_tc.cu), owned by rabi_tc_*
    void* train = nullptr;     // training-side state (maf_train.cu): device masks, emission buffers
    void* tcd = nullptr;       // three-hidden-layer tensor-core path state (maf_tcd.cu), owned by rabi_tcd_*
    const float* row_eps = nullptr;        // optional per-row attack radii / step sizes (rabi_pgd_set_row_scales)
    const float* row_eps_iter = nullptr;
    const float* col_mask = nullptr;       // optional per-column attack mask (rabi_pgd_set_column_mask)
    unsigned long long weights_version = 0;   // bumped by rabi_maf_set_weights / finalize: derived images re-pack lazily
};

int rabi_fail(rabi_maf* h, const char* fmt, ...);
int rabi_env_int(const char* name, int dflt);

#define RABI_CU_OK(h, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) return rabi_fail(h, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
    } while (0)

// CUDA-event bracket around the dominant kernel(s) of a call when profiling is on (rabi_profile)
int rabi_prof_begin(rabi_maf* h, cudaStream_t st);
int rabi_prof_end(rabi_maf* h, cudaStream_t st);
