// cnn_kernels.cuh — adapter / poly(A) boundary CNN of ADAPTed on the device (sm_100a).
//
// Restates the reference's `cnn_detect` chain, the step that produces the adapter boundaries
// the fingerprint stage consumes:
//   prepare_data    warpdemux/adapted/adapted/detect/cnn.py:71-85   (+ downscale.py:4-41)
//   BoundariesCNN   cnn.py:16-52    Conv1d(1,64,7,s3,p3) ReLU Conv1d(64,64,7,p3) ReLU Conv1d(64,64,7,p3) ReLU
//                                   ConvTranspose1d(64,2,7,s3,p3)
//   cnn_predict     cnn.py:104-162  argmax / masking / find_peaks(distance=5) on the FLATTENED batch / top-k
//   cnn_detect      cnn.py:165-183  * downscale_factor + min_obs_adapter, == min_obs_adapter -> 0
//
// This header holds the arithmetic that is shared by both modes: the float32 input preparation
// (bit-identical to numpy: block means in numpy's pairwise order, exact medians), the float32
// CUDA-core convolutions of the EXACT mode, and the integer boundary logic.  The tensor-core
// (tcgen05) convolution of the FAST mode lives in cnn_tc_kernel.cuh.
//
// Activation layout: h[read][t][64] float32 (time-major, channels contiguous).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "block_select.cuh"  // FpScratch, block_select_u32, f32_key (FP_THREADS CTAs)

namespace wdx {

constexpr int CNN_C = 64;   // channels
constexpr int CNN_K = 7;    // kernel size
constexpr int CNN_S = 3;    // stride of the first / last layer (kernel_size // 2)
constexpr int CNN_P = 3;    // padding (kernel_size // 2)
constexpr int CNN_MAX_T = 4096;   // downscaled samples per read the kernels are sized for
constexpr int CNN_MAX_TOPK = 16;  // polya_cand_k bound
constexpr int CNN_HALO = 64;      // flat samples either side of a read seen by the peak kernel
constexpr float CNN_SCORE_EXCL = -5.0f;  // cnn.py:13

enum { CNN_FLAG_NONFINITE = 1, CNN_FLAG_RECOMPUTED = 2, CNN_FLAG_CHAIN = 4, CNN_FLAG_RANGE = 8 };

struct CnnDims {
    int min_obs;   // core.min_obs_adapter
    int factor;    // core.downscale_factor
    int span;      // (max_obs_adapter - min_obs_adapter) / factor : adapter-end search range
    int topk;      // cnn_boundaries.polya_cand_k
    int T;         // downscaled input length  = ceil((stride - min_obs) / factor)
    int T1;        // hidden length            = (T - 1) / 3 + 1
    int To;        // score length             = 3 * T1 - 2
};

// torch.relu keeps NaN (fmaxf would drop it)
__device__ __forceinline__ float cnn_relu(float x) { return x < 0.0f ? 0.0f : x; }

// numpy's float32 add.reduce over n <= 128 contiguous values (pairwise_sum_FLOAT): sequential for
// n < 8, else 8 strided accumulators, a fixed combination tree and a sequential tail.
template <typename F>
__device__ __forceinline__ float np_pairwise_sum_f32(int n, F at) {
    if (n < 8) {
        float r = at(0);  // reduce starts from the first element
        for (int i = 1; i < n; i++) r = __fadd_rn(r, at(i));
        return r;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = at(j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], at(i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __fadd_rn(res, at(i));
    return res;
}

// np.nanmedian of T float32 values given as order keys (NaN -> 0xffffffff), n_valid of them not NaN.
template <typename KEY>
__device__ float block_nanmedian_f32(int T, int n_valid, KEY key, FpScratch& s) {
    if (n_valid <= 0) return __uint_as_float(0x7fc00000u);
    const uint32_t k_lo = (uint32_t)((n_valid - 1) / 2);
    const uint32_t key_lo = block_select_u32(T, k_lo, key, s);
    const float v_lo = f32_unkey(key_lo);
    if (n_valid & 1) return v_lo;
    // the next order statistic without a second radix selection: v_lo again if enough copies, else the smallest key
    // above it (rank k_lo + 1 < n_valid, so that key belongs to a valid sample, never to a NaN)
    __syncthreads();
    if (threadIdx.x == 0) {
        s.hist[0] = 0;            // count(key <= key_lo)
        s.hist[1] = 0xffffffffu;  // min key > key_lo
    }
    __syncthreads();
    uint32_t cnt = 0, mn = 0xffffffffu;
    for (int i = threadIdx.x; i < T; i += FP_THREADS) {
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
    __syncthreads();
    return __fdiv_rn(__fadd_rn(v_lo, v_hi), 2.0f);  // np.mean of the two middle float32 values
}

// ---- prepare_data: one CTA per read ------------------------------------------------------------
// signals [n][stride] float32 (NaN padded)  ->  x [n][T] float32
__global__ void __launch_bounds__(FP_THREADS) cnn_prepare_kernel(const float* __restrict__ signals, int64_t stride, int64_t n,
                                                                 CnnDims d, float* __restrict__ x) {
    extern __shared__ __align__(16) float cnn_ds[];  // [T]
    __shared__ FpScratch s;
    __shared__ int n_nan;
    const int tid = threadIdx.x;
    const int64_t read = blockIdx.x;
    if (read >= n) return;
    const float* row = signals + read * stride + d.min_obs;
    const int64_t width = stride - d.min_obs;
    if (tid == 0) n_nan = 0;
    __syncthreads();
    int cnt = 0;
    for (int b = tid; b < d.T; b += FP_THREADS) {
        const int64_t base = (int64_t)b * d.factor;
        // zero padding up to a multiple of the factor (downscale.py:22-27)
        const float sum = np_pairwise_sum_f32(d.factor, [&](int i) { return (base + i < width) ? __ldg(row + base + i) : 0.0f; });
        const float v = __fdiv_rn(sum, (float)d.factor);
        cnn_ds[b] = v;
        cnt += (v != v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((tid & 31) == 0 && cnt) atomicAdd(&n_nan, cnt);
    __syncthreads();
    const int n_valid = d.T - n_nan;
    const float med = block_nanmedian_f32(d.T, n_valid, [&](int i) {
        const float v = cnn_ds[i];
        return (v != v) ? 0xffffffffu : f32_key(v);
    }, s);
    const float mad = block_nanmedian_f32(d.T, n_valid, [&](int i) {
        const float v = cnn_ds[i];
        return (v != v) ? 0xffffffffu : f32_key(fabsf(__fsub_rn(v, med)));
    }, s);
    float* out = x + read * d.T;
    for (int b = tid; b < d.T; b += FP_THREADS) {
        float v = __fdiv_rn(__fsub_rn(cnn_ds[b], med), mad);
        if (v != v) v = CNN_SCORE_EXCL;                       // torch nan_to_num(nan=-5)
        else if (isinf(v)) v = copysignf(3.402823466e38f, v);  //   +-inf -> +-float32 max
        out[b] = v;
    }
}

// ---- EXACT mode: float32 CUDA-core convolutions ------------------------------------------------
// conv1: x [n][T] -> h [n][T1][64], stride 3, padding 3, ReLU.   w0 [64][7], b0 [64]
// n_dev (all EXACT kernels): optional read count on the device; reads at or beyond it are skipped, so a launch sized for
// the worst case needs no host round trip to learn how many reads there really are (GUARDED re-run).
__global__ void __launch_bounds__(256) cnn_conv1_f32_kernel(const float* __restrict__ x, const float* __restrict__ w0,
                                                            const float* __restrict__ b0, CnnDims d, float* __restrict__ h,
                                                            const int32_t* __restrict__ n_dev = nullptr) {
    __shared__ float ws[CNN_C * CNN_K], bs[CNN_C];
    if (n_dev && (int64_t)blockIdx.y >= (int64_t)*n_dev) return;
    const int tid = threadIdx.x;
    for (int i = tid; i < CNN_C * CNN_K; i += 256) ws[i] = w0[i];
    if (tid < CNN_C) bs[tid] = b0[tid];
    __syncthreads();
    const int64_t read = blockIdx.y;
    const float* xr = x + read * d.T;
    const int co = tid & 63;
    for (int t = blockIdx.x * 4 + (tid >> 6); t < d.T1; t += gridDim.x * 4) {
        float acc = bs[co];
#pragma unroll
        for (int k = 0; k < CNN_K; k++) {
            const int i = t * CNN_S + k - CNN_P;
            const float v = (i >= 0 && i < d.T) ? __ldg(xr + i) : 0.0f;
            acc = fmaf(v, ws[co * CNN_K + k], acc);
        }
        h[(read * d.T1 + t) * CNN_C + co] = cnn_relu(acc);
    }
}

// conv 64 -> 64, kernel 7, padding 3, ReLU.  wt [7][64 ci][64 co] (tap-major repack), bias [64].
// Persistent CTAs: the 112 KB of weights are staged once per CTA, then the CTA walks (read, 64-row tile)
// pairs; every thread owns a 4 (time) x 4 (channel) block of outputs.
constexpr int CV_TT = 64;
constexpr int CV_LD = 68;  // padded row of the input tile (floats): conflict-free broadcast reads
inline size_t cnn_conv64_smem_bytes() { return (size_t)(CNN_K * CNN_C * CNN_C + (CV_TT + 2 * CNN_P) * CV_LD) * 4; }

__global__ void __launch_bounds__(256) cnn_conv64_f32_kernel(const float* __restrict__ hin, float* __restrict__ hout,
                                                             const float* __restrict__ wt, const float* __restrict__ bias,
                                                             int T1, int tiles_per_read, int64_t n_tiles,
                                                             const int32_t* __restrict__ n_dev = nullptr) {
    extern __shared__ __align__(16) float cv_sm[];
    if (n_dev) n_tiles = min(n_tiles, (int64_t)*n_dev * tiles_per_read);
    if ((int64_t)blockIdx.x >= n_tiles) return;
    float* w_s = cv_sm;                          // [7][64][64]
    float* in_s = cv_sm + CNN_K * CNN_C * CNN_C;  // [70][CV_LD]
    const int tid = threadIdx.x;
    for (int i = tid; i < CNN_K * CNN_C * CNN_C / 4; i += 256)
        reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
    const int tx = tid & 15, ty = tid >> 4;  // channels 4*tx.., rows 4*ty..
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + tx);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t read = tile / tiles_per_read;
        const int t0 = (int)(tile % tiles_per_read) * CV_TT;
        __syncthreads();
        for (int i = tid; i < (CV_TT + 2 * CNN_P) * 16; i += 256) {
            const int r = i >> 4, c4 = i & 15;
            const int t = t0 - CNN_P + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < T1) v = __ldg(reinterpret_cast<const float4*>(hin + (read * T1 + t) * CNN_C) + c4);
            *reinterpret_cast<float4*>(in_s + r * CV_LD + c4 * 4) = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            acc[a][0] = b4.x; acc[a][1] = b4.y; acc[a][2] = b4.z; acc[a][3] = b4.w;
        }
#pragma unroll 2
        for (int ci = 0; ci < CNN_C; ci++) {
            float a[4 + CNN_K - 1];
#pragma unroll
            for (int j = 0; j < 4 + CNN_K - 1; j++) a[j] = in_s[(ty * 4 + j) * CV_LD + ci];
#pragma unroll
            for (int k = 0; k < CNN_K; k++) {
                const float4 w = *reinterpret_cast<const float4*>(w_s + (k * CNN_C + ci) * CNN_C + tx * 4);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    acc[q][0] = fmaf(a[q + k], w.x, acc[q][0]);
                    acc[q][1] = fmaf(a[q + k], w.y, acc[q][1]);
                    acc[q][2] = fmaf(a[q + k], w.z, acc[q][2]);
                    acc[q][3] = fmaf(a[q + k], w.w, acc[q][3]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int t = t0 + ty * 4 + q;
            if (t < T1)
                *reinterpret_cast<float4*>(hout + (read * T1 + t) * CNN_C + tx * 4) =
                    make_float4(cnn_relu(acc[q][0]), cnn_relu(acc[q][1]), cnn_relu(acc[q][2]), cnn_relu(acc[q][3]));
        }
    }
}

// ConvTranspose1d(64, 2, 7, stride 3, padding 3): h [n][T1][64] -> scores [n][2][To].
// out[c][u] = b[c] + sum_ci sum_{k = u + 3 - 3t} h[t][ci] * w[ci][c][k];   wT repacked [7][64][2].
template <typename HAT>  // hat(t, ci) -> float: the last hidden layer
__device__ __forceinline__ void cnn_convT_point(HAT hat, int T1, int u, const float* __restrict__ wT_s, float b0, float b1, float* o0,
                                                float* o1) {
    float a0 = b0, a1 = b1;
    for (int k = (u + CNN_P) % CNN_S; k < CNN_K; k += CNN_S) {
        const int t = (u + CNN_P - k) / CNN_S;
        if (t < 0 || t >= T1) continue;
        const float* w = wT_s + k * CNN_C * 2;
#pragma unroll 8
        for (int ci = 0; ci < CNN_C; ci++) {
            const float h = hat(t, ci);
            a0 = fmaf(h, w[ci * 2 + 0], a0);
            a1 = fmaf(h, w[ci * 2 + 1], a1);
        }
    }
    *o0 = a0;
    *o1 = a1;
}

__global__ void __launch_bounds__(128) cnn_convT_f32_kernel(const float* __restrict__ h3, const float* __restrict__ wT,
                                                            const float* __restrict__ b3, CnnDims d, float* __restrict__ scores,
                                                            const int32_t* __restrict__ n_dev = nullptr) {
    __shared__ float w_s[CNN_K * CNN_C * 2];
    if (n_dev && (int64_t)blockIdx.y >= (int64_t)*n_dev) return;
    for (int i = threadIdx.x; i < CNN_K * CNN_C * 2; i += 128) w_s[i] = wT[i];
    __syncthreads();
    const int64_t read = blockIdx.y;
    const int u = blockIdx.x * 128 + threadIdx.x;
    if (u >= d.To) return;
    float o0, o1;
    const float* hr = h3 + read * d.T1 * CNN_C;
    cnn_convT_point([&](int t, int ci) { return __ldg(hr + (size_t)t * CNN_C + ci); }, d.T1, u, w_s, __ldg(b3), __ldg(b3 + 1), &o0, &o1);
    scores[(read * 2 + 0) * d.To + u] = o0;
    scores[(read * 2 + 1) * d.To + u] = o1;
}

// ---- cnn_predict, part 1: per-read argmax and masking (cnn.py:115-137) ---------------------------
// One warp per read.  masked row m[To] receives channel 1 with SCORE_EXCL before the adapter end and
// after the poly(A) end.  *margin = min(top1 - top2) over the two argmax decisions (guard input).
__device__ __forceinline__ void cnn_argmax_read(const float* __restrict__ c0, const float* __restrict__ c1, const CnnDims& d,
                                                float* __restrict__ m, int* ae_out, int* pe_out, float* margin, bool* bad_out) {
    const int lane = threadIdx.x & 31;
    const float ninf = __uint_as_float(0xff800000u);
    // np.argmax: first maximum; NaN counts as the maximum (first NaN wins)
    auto better = [](float v, int i, float bv, int bi) {
        const bool vn = v != v, bn = bv != bv;
        if (vn != bn) return vn;
        if (vn) return i < bi;
        return v > bv || (v == bv && i < bi);
    };
    auto warp_argmax = [&](auto value, int len, float* best2) -> int {
        float bv = ninf, sv = ninf;  // best, second best
        int bi = 0x7fffffff;
        for (int i = lane; i < len; i += 32) {
            const float v = value(i);
            if (better(v, i, bv, bi)) { sv = bv; bv = v; bi = i; }
            else if (v > sv) sv = v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o), os = __shfl_xor_sync(0xffffffffu, sv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { sv = fmaxf(bv, os); bv = ov; bi = oi; }
            else sv = fmaxf(sv, ov);
        }
        *best2 = bv - sv;
        return bi == 0x7fffffff ? 0 : bi;
    };
    bool bad = false;
    for (int i = lane; i < d.To; i += 32) bad |= !isfinite(c0[i]) || !isfinite(c1[i]);
    *bad_out = __any_sync(0xffffffffu, bad);
    float g0, g1;
    const int span = min(d.span, d.To);
    const int ae = warp_argmax([&](int i) { return c0[i]; }, span, &g0);
    const int pe = warp_argmax([&](int i) { return i < ae ? CNN_SCORE_EXCL : c1[i]; }, d.To, &g1);
    for (int i = lane; i < d.To; i += 32) m[i] = (i < ae || i > pe) ? CNN_SCORE_EXCL : c1[i];
    *ae_out = ae;
    *pe_out = pe;
    *margin = fminf(g0, g1);
}

__global__ void __launch_bounds__(128) cnn_argmax_kernel(const float* __restrict__ scores, int64_t n, CnnDims d,
                                                         float* __restrict__ masked, int32_t* __restrict__ a_end,
                                                         int32_t* __restrict__ p_end, float* __restrict__ margin,
                                                         uint8_t* __restrict__ flags) {
    const int64_t read = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (read >= n) return;
    int ae, pe;
    float mg;
    bool bad;
    cnn_argmax_read(scores + (read * 2 + 0) * d.To, scores + (read * 2 + 1) * d.To, d, masked + read * d.To, &ae, &pe, &mg, &bad);
    if ((threadIdx.x & 31) == 0) {
        a_end[read] = ae;
        p_end[read] = pe;
        margin[read] = mg;
        if (bad) flags[read] |= CNN_FLAG_NONFINITE;
    }
}

// GUARDED re-run: scores of the q-th listed read are in row q; results go to read idx[q].
__global__ void __launch_bounds__(128) cnn_argmax_idx_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, int64_t m,
                                                             CnnDims d, float* __restrict__ masked, int32_t* __restrict__ a_end,
                                                             int32_t* __restrict__ p_end, uint8_t* __restrict__ flags,
                                                             float* __restrict__ scores_out, const int32_t* __restrict__ n_dev = nullptr) {
    const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (n_dev) m = min(m, (int64_t)*n_dev);
    if (q >= m) return;
    const int64_t read = idx[q];
    const float* c0 = scores + (q * 2 + 0) * d.To;
    int ae, pe;
    float mg;
    bool bad;
    cnn_argmax_read(c0, c0 + d.To, d, masked + read * d.To, &ae, &pe, &mg, &bad);
    if (scores_out)
        for (int i = threadIdx.x & 31; i < 2 * d.To; i += 32) scores_out[read * 2 * d.To + i] = c0[i];
    if ((threadIdx.x & 31) == 0) {
        a_end[read] = ae;
        p_end[read] = pe;
        flags[read] = (uint8_t)((flags[read] & ~(CNN_FLAG_RANGE | CNN_FLAG_NONFINITE)) | CNN_FLAG_RECOMPUTED | (bad ? CNN_FLAG_NONFINITE : 0));
    }
}

// GUARDED: list the reads to redo in EXACT_F32 (ordered list not needed).
__global__ void cnn_guard_list_kernel(const float* __restrict__ margin, const uint8_t* __restrict__ flags, int64_t n, float guard,
                                      int32_t* __restrict__ idx, int32_t* __restrict__ count) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const float mg = margin[r];
        if (!(mg >= guard) || (flags[r] & (CNN_FLAG_RANGE | CNN_FLAG_NONFINITE))) idx[atomicAdd(count, 1)] = (int32_t)r;
    }
}

__global__ void cnn_gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int T, float* __restrict__ out,
                                       const int32_t* __restrict__ n_dev = nullptr) {
    if (n_dev && (int64_t)blockIdx.x >= (int64_t)*n_dev) return;
    const float* src = x + (int64_t)idx[blockIdx.x] * T;
    float* dst = out + (int64_t)blockIdx.x * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) dst[i] = src[i];
}

// ---- cnn_predict, part 2: scipy find_peaks(distance) on the flattened batch ------------------------
// One CTA per read over the flat window [r*To - HALO, (r+1)*To + HALO).  Strict local maxima with
// plateau midpoints (scipy _local_maxima_1d), distance suppression as the unique fixed point of
// "a peak stays iff no STAYING peak of higher priority lies closer than `distance`" (what scipy's
// highest-first sweep computes; equal heights: the higher index wins), then the k highest surviving
// peaks of the read in (descending height, ascending position) order (np.lexsort, cnn.py:142-146).
// Peaks closer than `distance` to a window edge that is not an edge of the flat array have unknown
// neighbours; "unknown" propagates, and a read whose own peaks depend on one gets CNN_FLAG_CHAIN
// instead of a silently different answer.
constexpr int PK_THREADS = 256;
constexpr int PK_MAXPK = (CNN_MAX_T + 2 * CNN_HALO) / 2 + 2;

__global__ void __launch_bounds__(PK_THREADS) cnn_peaks_kernel(const float* __restrict__ masked, int64_t n, CnnDims d, int distance,
                                                               int32_t* __restrict__ cand /*[n][topk]*/, int32_t* __restrict__ n_cand,
                                                               uint8_t* __restrict__ flags) {
    __shared__ float v[CNN_MAX_T + 2 * CNN_HALO + 2];
    __shared__ int32_t pk[PK_MAXPK];
    __shared__ uint8_t st[PK_MAXPK];
    __shared__ int extra_lo, kept_total;
    __shared__ int def_n, def_i[8], def_p[8];
    __shared__ long long scan_hit;
    __shared__ uint32_t wtot[PK_THREADS / 32];
    const int tid = threadIdx.x;
    const int64_t read = blockIdx.x;
    const int64_t N = n * d.To;
    const int64_t own_lo = read * d.To, own_hi = own_lo + d.To;
    const int64_t lo = max((int64_t)0, own_lo - CNN_HALO), hi = min(N, own_hi + CNN_HALO);
    const int len = (int)(hi - lo);
    // v[1 + i] = flat[lo + i]; v[0] / v[len + 1] = the neighbours outside (or a copy of the edge value)
    for (int i = tid; i < len + 2; i += PK_THREADS) {
        int64_t g = lo - 1 + i;
        g = min(max(g, (int64_t)0), N - 1);
        v[i] = masked[g];
    }
    if (tid == 0) {
        extra_lo = -1;
        def_n = 0;
        scan_hit = 0x7fffffffffffffffLL;
    }
    __syncthreads();
    // The masked scores are flat (SCORE_EXCL) from one read's poly(A) end to the next read's adapter end: hundreds of equal
    // samples, and a rising edge into such a plateau (a score below SCORE_EXCL before it) or a window that starts inside one
    // sends scipy's plateau walk across them.  One thread doing that walk through global memory took 3/4 of this kernel's
    // time; the CTA does it together: first position != x at or after `from` (forward) / at or before `from` (backward).
    auto scan_forward = [&](int64_t from, float x) -> int64_t {      // first a >= from with flat[a] != x, or N - 1 (all threads)
        int64_t base = from;
        for (;;) {
            const int64_t a = base + tid;
            const bool hit = a >= N - 1 || masked[a] != x;
            if (hit) atomicMin((unsigned long long*)&scan_hit, (unsigned long long)min(a, N - 1));
            __syncthreads();
            const int64_t h = scan_hit;
            __syncthreads();
            if (h != 0x7fffffffffffffffLL) {
                if (tid == 0) scan_hit = 0x7fffffffffffffffLL;
                __syncthreads();
                return h;
            }
            base += PK_THREADS;
        }
    };
    // A plateau that starts left of the window but whose midpoint may fall inside.
    if (lo > 0 && v[0] == v[1]) {      // uniform
        const float x = v[1];
        // backward: last position b < lo with flat[b] != x (or -1): the plateau starts at b + 1
        int64_t base = lo - 1, b = -2;
        for (;;) {
            const int64_t q = base - tid;
            const bool hit = q < 0 || masked[q] != x;
            if (hit) atomicMin((unsigned long long*)&scan_hit, (unsigned long long)(lo - 1 - max(q, (int64_t)-1)));   // distance back
            __syncthreads();
            const int64_t h = scan_hit;
            __syncthreads();
            if (h != 0x7fffffffffffffffLL) {
                if (tid == 0) scan_hit = 0x7fffffffffffffffLL;
                __syncthreads();
                b = lo - 1 - h;
                break;
            }
            base -= PK_THREADS;
        }
        const int64_t j = b + 1;
        if (j > 0 && masked[j - 1] < x) {  // a rising edge at j (uniform: every thread reads the same element): find the end
            const int64_t e = scan_forward(lo, x);
            if (tid == 0 && masked[e] < x) {
                const int64_t mid = (j + e - 1) / 2;
                if (mid >= lo && mid < hi) extra_lo = (int)(mid - lo);
            }
        }
    }
    __syncthreads();
    // local maxima, position-ordered: thread t scans a contiguous chunk, block scan for the offsets
    const int chunk = (len + PK_THREADS - 1) / PK_THREADS;
    const int i0 = min(len, tid * chunk), i1 = min(len, i0 + chunk);
    auto at = [&](int64_t g) -> float {  // flat[g]: shared memory inside [lo-1, hi], global memory beyond (rare)
        const int64_t w = g - lo + 1;
        return (w >= 0 && w <= len + 1) ? v[w] : masked[g];
    };
    // rising edges whose plateau is longer than a few samples are set aside and resolved by the whole CTA
    constexpr int PK_SHORT = 6, PK_DEFER = 8;
    for (int i = i0; i < i1; i++) {
        const int64_t g = lo + i;
        if (g < 1 || g > N - 2) continue;
        const float x = v[1 + i];
        if (!(v[i] < x)) continue;
        int64_t a = g + 1;
        while (a < N - 1 && a - g <= PK_SHORT && at(a) == x) a++;
        if (a < N - 1 && a - g > PK_SHORT && at(a) == x) {
            const int slot = atomicAdd(&def_n, 1);
            if (slot < PK_DEFER) def_i[slot] = i;
        }
    }
    __syncthreads();
    {
        const int nd = min(def_n, PK_DEFER);
        for (int k = 0; k < nd; k++) {      // uniform
            const int i = def_i[k];
            const int64_t g = lo + i;
            const float x = v[1 + i];
            const int64_t a = scan_forward(g + 1, x);
            if (tid == 0) {
                int p = -1;
                if (masked[a] < x) {
                    const int64_t mid = (g + a - 1) / 2;
                    if (mid < hi) p = (int)(mid - lo);
                }
                def_p[k] = p;
            }
        }
        __syncthreads();
    }
    // scipy _local_maxima_1d: a rising edge at g, the plateau [g, a), a falling edge at a -> midpoint
    auto peak_from = [&](int i) -> int {
        const int64_t g = lo + i;
        if (g < 1 || g > N - 2) return -1;
        const float x = v[1 + i];
        if (!(v[i] < x)) return -1;
        int64_t a = g + 1;
        while (a < N - 1 && a - g <= PK_SHORT && at(a) == x) a++;
        if (a < N - 1 && a - g > PK_SHORT && at(a) == x) {      // a long plateau: resolved above (the serial walk only past the list's end)
            const int nd = min(def_n, PK_DEFER);
            for (int k = 0; k < nd; k++)
                if (def_i[k] == i) return def_p[k];
            while (a < N - 1 && at(a) == x) a++;
        }
        if (!(at(a) < x)) return -1;
        const int64_t mid = (g + a - 1) / 2;
        return (mid < hi) ? (int)(mid - lo) : -1;
    };
    uint32_t my = 0;
    for (int i = i0; i < i1; i++) my += (peak_from(i) >= 0);
    if (tid == 0 && extra_lo >= 0) my++;
    // block exclusive scan of `my`
    uint32_t inc = my;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) wtot[tid >> 5] = inc;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PK_THREADS / 32; w++) {
        if (w < (tid >> 5)) base += wtot[w];
        total += wtot[w];
    }
    uint32_t off = base + inc - my;
    if (tid == 0 && extra_lo >= 0) { pk[off] = extra_lo; st[off] = 1; off++; }
    for (int i = i0; i < i1; i++) {
        const int p = peak_from(i);
        if (p >= 0) { pk[off] = p; st[off] = 1; off++; }
    }
    __syncthreads();
    const int P = (int)total;
    // states: 1 undecided, 2 kept, 3 removed, 4 unknown (depends on samples outside the window)
    const bool open_lo = lo > 0, open_hi = hi < N;
    if (distance > 1) {
        for (int j = tid; j < P; j += PK_THREADS) {
            const int p = pk[j];
            if ((open_lo && p < distance - 1) || (open_hi && len - 1 - p < distance - 1)) st[j] = 4;
        }
        __syncthreads();
        // A verdict is written only when it is final (removed: a higher KEPT peak is near; kept / unknown: every higher
        // peak near has its final state), and final states never change, so it does not matter how fresh the neighbours'
        // states are when they are read: no barrier between the rounds, every warp iterates over its own undecided peaks
        // until none is left (the highest undecided peak can always be decided, so the loops end).
        volatile uint8_t* vst = st;
        bool mine_left = true;
        for (int spin = 0; __any_sync(0xffffffffu, mine_left); spin++) {
            if (spin > (1 << 22)) __trap();   // a protocol bug must end as a launch failure, never as a hung GPU
            mine_left = false;
            for (int j = tid; j < P; j += PK_THREADS) {
                if (vst[j] != 1) continue;
                const int pj = pk[j];
                const float x = v[1 + pj];
                bool killed = false, blocked = false, unknown = false;
                for (int q = j - 1; q >= 0 && pj - pk[q] < distance; q--) {
                    const int sq = vst[q];
                    if (sq == 3) continue;
                    if (v[1 + pk[q]] > x) {
                        if (sq == 2) killed = true;
                        else if (sq == 4) unknown = true;
                        else blocked = true;
                    }
                }
                for (int q = j + 1; q < P && pk[q] - pj < distance; q++) {
                    const int sq = vst[q];
                    if (sq == 3) continue;
                    if (v[1 + pk[q]] >= x) {
                        if (sq == 2) killed = true;
                        else if (sq == 4) unknown = true;
                        else blocked = true;
                    }
                }
                if (killed) vst[j] = 3;
                else if (blocked) mine_left = true;
                else vst[j] = unknown ? 4 : 2;
            }
        }
        __syncthreads();
    } else {
        for (int j = tid; j < P; j += PK_THREADS) st[j] = 2;
        __syncthreads();
    }
    // own surviving peaks, ranked by (height desc, position asc); only the topk highest are reported.  Every warp
    // takes its own topk best (topk rounds of a warp maximum over order keys: height, then lower position), the few
    // warp candidates are ranked against each other — a peak of global rank r < topk is outranked by r peaks, each of
    // which is among ITS warp's r + 1 best, so the rank among the candidates is the global rank.
    const int w_lo = (int)(own_lo - lo), w_hi = (int)(own_hi - lo);
    int own_kept = 0, own_unknown = 0;
    constexpr int PK_TOP = 8;
    __shared__ unsigned long long topc[PK_THREADS / 32][PK_TOP];
    auto key_of = [&](int p) -> unsigned long long {
        return ((unsigned long long)f32_key(v[1 + p] + 0.0f) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p);
    };
    for (int j = tid; j < P; j += PK_THREADS) {
        const int p = pk[j];
        if (p < w_lo || p >= w_hi) continue;
        if (st[j] == 4) own_unknown = 1;
        if (st[j] == 2) own_kept++;
    }
    if (d.topk <= PK_TOP) {
        unsigned long long prev = ~0ull;
        for (int r = 0; r < d.topk; r++) {
            unsigned long long best = 0ull;
            for (int j = tid; j < P; j += PK_THREADS) {
                const int p = pk[j];
                if (p < w_lo || p >= w_hi || st[j] != 2) continue;
                const unsigned long long key = key_of(p);
                if (key < prev && key > best) best = key;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if ((tid & 31) == 0) topc[tid >> 5][r] = best;   // 0: the warp has no further peak
            prev = best;
        }
        __syncthreads();
        if (tid < (PK_THREADS / 32) * d.topk) {
            const unsigned long long key = topc[tid / d.topk][tid % d.topk];
            if (key) {
                int rank = 0;
                for (int w = 0; w < PK_THREADS / 32; w++)
                    for (int r = 0; r < d.topk; r++) rank += topc[w][r] > key;
                if (rank < d.topk) cand[read * d.topk + rank] = (int)(0xffffffffu - (uint32_t)key) - w_lo;
            }
        }
    } else {
        for (int j = tid; j < P; j += PK_THREADS) {
            const int p = pk[j];
            if (p < w_lo || p >= w_hi || st[j] != 2) continue;
            const float x = v[1 + p];
            int rank = 0;
            for (int q = 0; q < P && rank < d.topk; q++) {   // only ranks below topk are used: stop counting there
                const int pq = pk[q];
                if (st[q] != 2 || pq < w_lo || pq >= w_hi) continue;
                const float y = v[1 + pq];
                rank += (y > x) || (y == x && pq < p);
            }
            if (rank < d.topk) cand[read * d.topk + rank] = p - w_lo;
        }
    }
    own_unknown = __syncthreads_or(own_unknown);
    if (tid == 0) kept_total = 0;
    __syncthreads();
    if (own_kept) atomicAdd(&kept_total, own_kept);
    __syncthreads();
    if (tid == 0) {
        n_cand[read] = kept_total;
        if (own_unknown && flags) flags[read] |= CNN_FLAG_CHAIN;
    }
    for (int j = tid; j < d.topk; j += PK_THREADS)
        if (j >= kept_total) cand[read * d.topk + j] = 0;  // zero padding (cnn.py:153-158)
}

// ---- cnn_predict part 3 + cnn_detect: rows (with the reference's group shift), scaling ---------------
// The reference writes the i-th GROUP of candidates (reads that have at least one peak, in order)
// to row i (cnn.py:147-158).  Single CTA: scan of has-peak flags, then scatter.
__global__ void __launch_bounds__(1024) cnn_rows_kernel(const int32_t* __restrict__ a_end, const int32_t* __restrict__ cand,
                                                        const int32_t* __restrict__ n_cand, int64_t n, CnnDims d,
                                                        int64_t* __restrict__ preds /*[n][1+topk]*/) {
    __shared__ uint32_t wsum[32];
    __shared__ int64_t carry;
    const int tid = threadIdx.x;
    const int ld = 1 + d.topk;
    if (tid == 0) carry = 0;
    __syncthreads();
    auto scale = [&](int pos) -> int64_t {
        const int64_t s = (int64_t)pos * d.factor + d.min_obs;
        return s == d.min_obs ? 0 : s;  // cnn.py:180
    };
    for (int64_t r0 = 0; r0 < n; r0 += 1024) {
        const int64_t r = r0 + tid;
        const uint32_t has = (r < n && n_cand[r] > 0) ? 1u : 0u;
        uint32_t inc = has;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((tid & 31) >= o) inc += t;
        }
        if ((tid & 31) == 31) wsum[tid >> 5] = inc;
        __syncthreads();
        uint32_t base = 0, tot = 0;
        for (int w = 0; w < 32; w++) {
            if (w < (tid >> 5)) base += wsum[w];
            tot += wsum[w];
        }
        if (r < n) {
            preds[r * ld] = scale(a_end[r]);
            if (has) {
                const int64_t row = carry + base + inc - 1;
                for (int j = 0; j < d.topk; j++) preds[row * ld + 1 + j] = scale(cand[r * d.topk + j]);
            }
        }
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    // rows beyond the number of groups keep zero candidates -> scale(0) = 0
    const int64_t groups = carry;
    for (int64_t r = groups + tid; r < n; r += 1024)
        for (int j = 0; j < d.topk; j++) preds[r * ld + 1 + j] = 0;
}

}  // namespace wdx
