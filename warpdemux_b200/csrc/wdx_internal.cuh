// wdx_internal.cuh — pieces shared by the translation units of libwdx_b200.so
// (not part of the C ABI): error reporting, grow-only buffers, the model handle.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "../../include/wdx_b200.h"
#include "wdx_types.cuh"

namespace wdx {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return ::wdx::fail(WDX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                                 \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
};

struct HostBuf {  // pinned
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
};

// 0 pageable host (or unknown), 1 pinned host, 2 device / managed
int mem_kind(const void* p);

}  // namespace wdx

struct wdx_model {
    using ModelDev = wdx::ModelDev;
    using DevBuf = wdx::DevBuf;
    using HostBuf = wdx::HostBuf;
    ModelDev dev{};
    int device = 0;
    int sm_count = 148;
    int k = 0, L = 0, n_sv = 0, n_pairs = 0;
    bool specialised = false;  // L == 25 && window == 15
    double guard = 5e-5;
    std::mutex mu;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D prefetch of the next chunk
    std::vector<void*> owned;            // model arrays on the device
    // workspaces (grow-only)
    DevBuf part, part2, near_idx, counters;
    DevBuf xdev[2], lab_dev[2], conf_dev[2], prob_dev[2], flag_dev[2], dist_dev[2];
    HostBuf xpin[2];
    DevBuf small_dev;   // live-sized batches: input and all results in one device block ...
    HostBuf small_pin;  // ... mirrored by one pinned block (one copy each way, one synchronisation)
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr},
                ev_d2h[2] = {nullptr, nullptr};
    // timing of the fused kernel
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev;
    std::vector<char> tev_exact;
    size_t tev_used = 0;
    int64_t chunk_reads = (int64_t)1 << 22;
    int forced_splits = 0;
};

namespace wdx {
// One chunk, everything on the device: Xd [n][L] -> labels/conf/prob/flags (device).  wdx_b200.cu
int predict_chunk_device(wdx_model* m, const void* Xd, int x_is_f32, int64_t n, int mode, int64_t* labels_d,
                         double* conf_d, double* prob_d, uint8_t* flags_d, float* dist_d, cudaStream_t st);
}  // namespace wdx
