// block_select.cuh — CTA-wide exact selection helpers shared by the fingerprint and the boundary-CNN
// kernels: order-preserving float keys, block scan, warp-aggregated histogram, 8-bit radix selection
// and numpy's float32 median.  All loops assume blockDim.x == FP_THREADS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wdx {

#ifndef WDX_FP_THREADS
#define WDX_FP_THREADS 512
#endif
constexpr int FP_THREADS = WDX_FP_THREADS;  // 512 x 2 CTAs/SM = 32 warps/SM at <= 64 registers
constexpr int FP_WARPS = FP_THREADS / 32;
constexpr int FP_MAX_EVENTS = 254;   // num_events bound (cpts has num_events + 2 entries)
#ifndef WDX_FP_MAX_LEN
#define WDX_FP_MAX_LEN 16000
#endif
constexpr int FP_MAX_LEN = WDX_FP_MAX_LEN;    // longest adapter slice a CTA can hold in shared memory (14 B per sample)


// ---- order-preserving keys ---------------------------------------------------
__device__ __forceinline__ uint32_t f32_key(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- block helpers -------------------------------------------------------------
// Phase clock of the fingerprint kernel (experiments: -DWDX_FP_PROF): thread 0 adds the cycles since its previous mark
// to counter i; read back with wdx_fp_prof_dump.
#ifdef WDX_FP_PROF
__device__ unsigned long long g_fp_prof[32];
#define FP_T(s_, i_)                                                                  \
    do {                                                                              \
        if (threadIdx.x == 0) {                                                       \
            const long long t__ = clock64();                                          \
            (s_).prof[i_] += (unsigned long long)(t__ - (s_).t_prev);                 \
            (s_).t_prev = t__;                                                        \
        }                                                                             \
    } while (0)
#define FP_T_BEGIN(s_)                                                                \
    do {                                                                              \
        if (threadIdx.x == 0) {                                                       \
            for (int i__ = 0; i__ < 32; i__++) (s_).prof[i__] = 0;                    \
            (s_).t_prev = clock64();                                                  \
        }                                                                             \
    } while (0)
#define FP_T_END(s_)                                                                  \
    do {                                                                              \
        if (threadIdx.x == 0)                                                         \
            for (int i__ = 0; i__ < 32; i__++)                                        \
                if ((s_).prof[i__]) atomicAdd(&g_fp_prof[i__], (s_).prof[i__]);       \
    } while (0)
#else
#define FP_T(s_, i_) do { } while (0)
#define FP_T_BEGIN(s_) do { } while (0)
#define FP_T_END(s_) do { } while (0)
#endif

struct FpScratch {
    uint32_t hist[256];
    uint32_t warp_tmp[FP_WARPS];
    uint32_t sel_prefix;   // radix select: key prefix found so far
    uint32_t sel_k;        // radix select: rank still to resolve inside the prefix
    unsigned long long sel_prefix64;
    int flag;
    int first_nan;
    int n_kept;
    uint32_t vmin_key, vmax_key;      // order keys of the slice minimum / maximum
    int sel_bin;                       // linear-bin median: bin holding the wanted rank
    uint32_t sel_below, sel_count;     //   elements in lower bins / in that bin
    uint32_t ncand, ncand2;            //   gathered candidates (first / second median of a pair)
    uint32_t amin, amax;               // smallest non-zero / largest |x| (float bit patterns) of the winsorised slice
#ifdef WDX_FP_PROF
    long long t_prev;                  // clock of the previous phase mark (thread 0)
    unsigned long long prof[32];       // cycles per phase of this read
#endif
    uint32_t key_lo, key_hi;
};

constexpr int FP_MED_BINS = 2048;      // histogram bins of the linear-bin median (aliases the score array)
constexpr int FP_MED_CAND = 512;       // candidate keys kept from the bin that holds the median

// Exclusive prefix sum of one value per thread over the CTA; returns the exclusive
// prefix, *total receives the CTA sum.  Contains __syncthreads.
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, FpScratch& s, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // warp_tmp free
    if (lane == 31) s.warp_tmp[warp] = inc;
    __syncthreads();
    // every warp scans the FP_WARPS warp totals itself (one value per lane) instead of every thread adding them up
    const uint32_t t = (lane < FP_WARPS) ? s.warp_tmp[lane] : 0u;
    uint32_t ti = t;
#pragma unroll
    for (int o = 1; o < FP_WARPS; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, ti, FP_WARPS - 1);
    const uint32_t base = __shfl_sync(0xffffffffu, ti - t, warp);
    *total = tot;
    return base + inc - v;
}

// Histogram add with warp aggregation (the top bytes of signal samples are
// nearly identical, so naive shared atomics would serialise).
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool valid) {
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}

// k-th smallest (0-based) of key(i), i in [0,n), by 4 passes of 8-bit radix
// selection.  KEY is a functor int -> uint32_t.  All threads get the result.
template <typename KEY>
__device__ uint32_t block_select_u32(int n, uint32_t k, KEY key, FpScratch& s) {
    uint32_t prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        __syncthreads();
        if (threadIdx.x < 256) s.hist[threadIdx.x] = 0;
        __syncthreads();
        const int n_round = (n + 31) & ~31;  // keep whole warps in the loop for the ballot
        for (int i = threadIdx.x; i < n_round; i += FP_THREADS) {
            uint32_t kv = 0;
            bool ok = false;
            if (i < n) {
                kv = key(i);
                ok = (kv & mask) == prefix;
            }
            hist_add(s.hist, (kv >> shift) & 255u, ok);
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // warp 0 scans the 256 bins (8 per lane)
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                c[q] = s.hist[threadIdx.x * 8 + q];
                sum += c[q];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += t;
            }
            uint32_t run = inc - sum;  // elements in lower bins
            if (k >= run && k < inc) {  // exactly one lane
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (k >= run && k < run + c[q]) {
                        s.sel_prefix = prefix | ((uint32_t)(threadIdx.x * 8 + q) << shift);
                        s.sel_k = k - run;
                    }
                    run += c[q];
                }
            }
        }
        __syncthreads();
        prefix = s.sel_prefix;
        k = s.sel_k;
        mask |= 255u << shift;
    }
    return prefix;
}

// numpy median of n float32 values given through KEY (np.median / np.nanmedian on
// a NaN-free 1-D float32 array): middle element, or fl32(fl32(a+b)/2) for even n.
template <typename KEY>
__device__ float block_median_f32(int n, KEY key, FpScratch& s) {
    const uint32_t k_lo = (uint32_t)((n - 1) / 2);
    const uint32_t key_lo = block_select_u32(n, k_lo, key, s);
    const float v_lo = f32_unkey(key_lo);
    if (n & 1) return v_lo;
    // the next order statistic: v_lo again if enough copies, else the smallest key above it
    __syncthreads();
    if (threadIdx.x == 0) {
        s.hist[0] = 0;            // count(key <= key_lo)
        s.hist[1] = 0xffffffffu;  // min key > key_lo
    }
    __syncthreads();
    uint32_t cnt = 0, mn = 0xffffffffu;
    for (int i = threadIdx.x; i < n; i += FP_THREADS) {
        const uint32_t kv = key(i);
        if (kv <= key_lo) cnt++;
        else mn = min(mn, kv);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s.hist[0], cnt);
        atomicMin(&s.hist[1], mn);
    }
    __syncthreads();
    const float v_hi = (s.hist[0] >= k_lo + 2) ? v_lo : f32_unkey(s.hist[1]);
    return __fdiv_rn(__fadd_rn(v_lo, v_hi), 2.0f);  // np.mean of two float32: float32 add, then /2
}

}  // namespace wdx
