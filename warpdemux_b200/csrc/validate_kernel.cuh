// validate_kernel.cuh — validation of CNN boundary predictions, one CTA per read (sm_100a).
//
// The step between the boundary CNN and the fingerprint stage.  Restates, for
// mvs_detect_overwrite = false (every shipped configuration), the reference's per-read Python chain
//   validate_boundaries                 warpdemux/adapted/adapted/detect/combined.py:409-683
//     adapter median / MAD (float32)    combined.py:452-467
//     find_open_pores                   adapted/detect/anomalies.py:16-35
//     real_range_check                  adapted/detect/real_range.py:34-63   (np.mean float32, np.percentile)
//     mean_var_shift_polyA_check        adapted/detect/mvs.py:42-159          (moving mean / variance medians,
//                                                                              median, percentiles, median shift)
//     median-shift check                combined.py:612-629
// as ONE persistent kernel: a CTA stages the row of one read in shared memory (read from HBM once,
// coalesced) and evaluates every check on-chip; per read 1 byte of verdict, three boundaries and a few
// statistics go back.  HBM-bound by construction: algorithmic bytes per read = 4 * min(len, stride).
//
// Exactness (this TU is compiled with -fmad=false): medians and percentiles are exact order statistics
// found by histogram / radix selection on order-preserving keys; float32 / float64 sums follow numpy's pairwise
// summation order; np.percentile's linear interpolation is evaluated in float64 as numpy does.  The
// sequential open-pore scan is replaced by its closed form: position i is kept iff an earlier open-pore
// sample exists and none lies within min_obs_diff - 1 samples before i.
// bottleneck.move_mean / move_var (third-party, float32 out) are computed per window in float64
// (exact window mean; two-pass variance), as documented — not bottleneck's running update.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

// CTA shape of this translation unit (block_select.cuh helpers follow WDX_FP_THREADS)
#ifndef WDX_VAL_THREADS
#define WDX_VAL_THREADS 512
#endif
#ifndef WDX_VAL_MIN_CTAS
#define WDX_VAL_MIN_CTAS 2
#endif
#define WDX_FP_THREADS WDX_VAL_THREADS
#include "block_select.cuh"

namespace wdx {

enum {
    VAL_OK = 0, VAL_NO_ADAPTER = 1, VAL_ADAPTER_MAD = 2, VAL_OPEN_PORE = 3, VAL_REAL_RANGE = 4, VAL_NO_POLYA = 5,
    VAL_MVS_NO_SIGNAL = 6, VAL_MVS_CHECKS = 7, VAL_MED_SHIFT = 8, VAL_HAS_NAN = 9
};
constexpr int VAL_NVALS = 12;
constexpr int VAL_NPART = 18;  // adapter, polya, rna_preloaded x (start, len, mean, std, med, mad)
constexpr int VAL_PORES_LD = 64;  // DetectResults.open_pores row: count (-1 = None), then up to 63 positions

struct ValCfg {
    int min_obs_adapter;
    int detect_open_pores, real_signal_check, mean_window, max_obs_local_range;
    double mean_start_lo, mean_start_hi, mean_end_lo, mean_end_hi, local_range_lo, local_range_hi, mad_lo, mad_hi;
    float open_pore_min;
    int open_pore_min_obs_diff;
    int mvs_detect_check, pa_mean_window, pa_var_window, median_shift_window;
    double var_lo, var_hi, shift_lo, shift_hi, pmed_lo, pmed_hi, plr_lo, plr_hi, mean_lo, mean_hi, scale_lo, scale_hi;
    int mean_from_scale;  // pA_mean_range empty and pA_mean_adapter_med_scale_range set (combined.py:505-519)
    int detect_med_shift, med_shift_window;
    double ms_lo, ms_hi;
};

struct ValArgs {
    const float* signals;      // [n][stride] calibrated pA, NaN padded
    int64_t stride;
    const int32_t* full_len;   // [n] full_signal_lens (may exceed stride)
    const int64_t* preds;      // [n][ld] adapter end, poly(A) end candidates (cnn_detect's output)
    int ld;
    int64_t n;
    uint8_t* success;          // [n]
    int32_t* info;             // [n][4] fail code, check bits (bit i = check i passed), open pores kept, 0
    int64_t* bounds;           // [n][3] adapter_start, adapter_end, polya_end
    double* vals;              // [n][VAL_NVALS] or nullptr
    double* parts;             // [n][VAL_NPART] partition statistics (signal_partitions.py:65-96) or nullptr
    int32_t* pores;            // [n][VAL_PORES_LD] DetectResults.open_pores (count, positions ascending) or nullptr
    float* scratch;            // [gridDim.x][stride] moving-window statistics
    int verdict_only;          // stop at the first failing poly(A) candidate: same success / boundaries (success is never
                               // set back once a candidate failed), fail code and statistics of THAT candidate instead of the last
    unsigned long long* next;  // work counter (zeroed before the launch): reads are handed out dynamically, because a
                               // read whose first poly(A) candidate fails costs several times a read that validates
    // re-validation of the LLR fallback's proposals (llr_kernel.cuh; combined.py:222-290): work lists on the device
    const int* list;           // or nullptr: only the reads list[0 .. *list_count) are validated, the others keep their results
    const int* list_count;
    int commit_on_success;     // != 0: results are written only when the validation succeeds (combined.py:288-289)
    int src_tag;               // list mode: written to the low two bits of info[.][3] with the results (LLR_SRC_*)
    int* fail_list;            // or nullptr: reads whose validation fails (other than by the NaN error) are appended here
    int* fail_count;
};

// numpy's pairwise summation (np.add.reduce on a contiguous 1-D array): a block of n <= 128 elements is
// summed with 8 strided accumulators, a fixed combination tree and a sequential tail ...
template <typename T, typename F>
__device__ __forceinline__ T np_pairwise_leaf(int lo, int n, F at) {
    if (n < 8) {
        T r = (T)0;
        for (int i = 0; i < n; i++) r = r + at(lo + i);
        return r;
    }
    T r0 = at(lo), r1 = at(lo + 1), r2 = at(lo + 2), r3 = at(lo + 3), r4 = at(lo + 4), r5 = at(lo + 5), r6 = at(lo + 6),
      r7 = at(lo + 7);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 = r0 + at(lo + i);
        r1 = r1 + at(lo + i + 1);
        r2 = r2 + at(lo + i + 2);
        r3 = r3 + at(lo + i + 3);
        r4 = r4 + at(lo + i + 4);
        r5 = r5 + at(lo + i + 5);
        r6 = r6 + at(lo + i + 6);
        r7 = r7 + at(lo + i + 7);
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; i++) res = res + at(lo + i);
    return res;
}

// ... and longer arrays are halved recursively (left half rounded down to a multiple of 8); the recursion
// is unrolled onto an explicit stack here.
template <typename T, typename F>
__device__ T np_pairwise(int lo, int n, F at) {
    if (n <= 128) return np_pairwise_leaf<T>(lo, n, at);
    struct Frame {
        int lo, n, state;
        T left;
    };
    Frame st[24];
    int sp = 0;
    st[sp++] = Frame{lo, n, 0, (T)0};
    T result = (T)0;
    while (sp > 0) {
        Frame& f = st[sp - 1];
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.n <= 128) {
            result = np_pairwise_leaf<T>(f.lo, f.n, at);
            sp--;
        } else if (f.state == 0) {
            f.state = 1;
            st[sp++] = Frame{f.lo, n2, 0, (T)0};
        } else if (f.state == 1) {
            f.left = result;
            f.state = 2;
            st[sp++] = Frame{f.lo + n2, f.n - n2, 0, (T)0};
        } else {
            result = f.left + result;
            sp--;
        }
    }
    return result;
}

// ---- order statistics ---------------------------------------------------------------------------------
// Several exact order statistics of the same n float32 values in three light passes (instead of five
// radix passes per statistic): (1) min / max, (2) a 2048-bin histogram over [min, max] — a monotone map,
// so every element of a lower bin is <= every element of a higher bin, (3) the few members of the bins
// that hold the wanted ranks are gathered and ranked exactly on their order keys.  A crowded bin
// (degenerate data) falls back to radix selection for that rank.
constexpr int VAL_BINS = 2048;
constexpr int VAL_MAXQ = 6;      // ranks per call
constexpr int VAL_CAND = 320;    // members kept per selected bin

struct ValSel {
    uint32_t hist[VAL_BINS];
    uint32_t cand[VAL_MAXQ][VAL_CAND];
    uint32_t slot_n[VAL_MAXQ];
    int q_bin[VAL_MAXQ];
    uint32_t q_r[VAL_MAXQ], q_cnt[VAL_MAXQ], q_key[VAL_MAXQ];
    uint32_t mnk, mxk;
};

// out[q] = the ranks[q]-th smallest (0-based) value, q < nq <= VAL_MAXQ; ranks < n, n >= 1.  All threads get all results.
template <typename VAL>
__device__ __forceinline__ void block_ranks_t(int n, VAL val, int nq, const uint32_t* ranks, float* out, ValSel& vs, FpScratch& s) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) {
        vs.mnk = 0xffffffffu;
        vs.mxk = 0u;
    }
    if (tid < VAL_MAXQ) vs.slot_n[tid] = 0;
    for (int b = tid; b < VAL_BINS; b += FP_THREADS) vs.hist[b] = 0;
    __syncthreads();
    {
        uint32_t mn = 0xffffffffu, mx = 0u;
        for (int i = tid; i < n; i += FP_THREADS) {
            const uint32_t kv = f32_key(val(i));
            mn = min(mn, kv);
            mx = max(mx, kv);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(&vs.mnk, mn);
            atomicMax(&vs.mxk, mx);
        }
    }
    __syncthreads();
    const float vmin = f32_unkey(vs.mnk), vmax = f32_unkey(vs.mxk);
    if (!(vmax > vmin)) {   // all values equal
        for (int q = 0; q < nq; q++) out[q] = vmin;
        return;
    }
    const float scale = __fdiv_rn((float)VAL_BINS, __fsub_rn(vmax, vmin));
    auto key_of = [&](int i) { return f32_key(val(i)); };
    if (!(scale < 1e30f)) {
        for (int q = 0; q < nq; q++) out[q] = f32_unkey(block_select_u32(n, ranks[q], key_of, s));
        return;
    }
    auto bin_of = [&](float x) { return min(VAL_BINS - 1, (int)__fmul_rn(__fsub_rn(x, vmin), scale)); };
    for (int i = tid; i < n; i += FP_THREADS) atomicAdd(&vs.hist[bin_of(val(i))], 1u);
    __syncthreads();
    {   // thread t owns bins [t*B, (t+1)*B)
        constexpr int B = VAL_BINS / FP_THREADS;
        uint32_t c[B], sum = 0;
#pragma unroll
        for (int j = 0; j < B; j++) {
            c[j] = vs.hist[tid * B + j];
            sum += c[j];
        }
        uint32_t total;
        const uint32_t run0 = block_exscan(sum, s, &total);
        for (int q = 0; q < nq; q++) {
            const uint32_t k = ranks[q];
            if (k >= run0 && k < run0 + sum) {  // exactly one thread per rank
                uint32_t run = run0;
#pragma unroll
                for (int j = 0; j < B; j++) {
                    if (k >= run && k < run + c[j]) {
                        vs.q_bin[q] = tid * B + j;
                        vs.q_r[q] = k - run;
                        vs.q_cnt[q] = c[j];
                    }
                    run += c[j];
                }
            }
        }
    }
    __syncthreads();
    // distinct selected bins -> slots (every thread derives the same mapping)
    int slot_of[VAL_MAXQ], slot_bin[VAL_MAXQ], n_slots = 0;
    bool crowded = false;
    for (int q = 0; q < nq; q++) {
        const int b = vs.q_bin[q];
        int sl = -1;
        for (int j = 0; j < n_slots; j++)
            if (slot_bin[j] == b) sl = j;
        if (vs.q_cnt[q] > (uint32_t)VAL_CAND) {
            crowded = true;
            slot_of[q] = -1;
            continue;
        }
        if (sl < 0) {
            sl = n_slots++;
            slot_bin[sl] = b;
        }
        slot_of[q] = sl;
    }
    for (int i = tid; i < n; i += FP_THREADS) {
        const float x = val(i);
        const int b = bin_of(x);
        for (int j = 0; j < n_slots; j++)
            if (b == slot_bin[j]) vs.cand[j][atomicAdd(&vs.slot_n[j], 1u)] = f32_key(x);
    }
    __syncthreads();
    for (int q = 0; q < nq; q++) {
        const int sl = slot_of[q];
        if (sl < 0) continue;
        const uint32_t m = vs.q_cnt[q], r = vs.q_r[q];
        for (uint32_t t = tid; t < m; t += FP_THREADS) {
            const uint32_t x = vs.cand[sl][t];
            uint32_t rank = 0;
            for (uint32_t u = 0; u < m; u++) {
                const uint32_t y = vs.cand[sl][u];
                rank += (y < x) || (y == x && u < t);
            }
            if (rank == r) vs.q_key[q] = x;
        }
    }
    __syncthreads();
    for (int q = 0; q < nq; q++)
        if (slot_of[q] >= 0) out[q] = f32_unkey(vs.q_key[q]);
    if (crowded) {   // uniform decision
        for (int q = 0; q < nq; q++)
            if (slot_of[q] < 0) out[q] = f32_unkey(block_select_u32(n, ranks[q], key_of, s));
    }
}

// np.median / np.nanmedian of NaN-free float32 values from their two middle order statistics.
__device__ __forceinline__ float median_of(int n, float lo, float hi) { return (n & 1) ? lo : __fdiv_rn(__fadd_rn(lo, hi), 2.0f); }

// np.percentile(x, q), method "linear": the two ranks it interpolates between, and the float64 result.
__device__ __forceinline__ void pct_ranks(int n, double q, uint32_t* r0, uint32_t* r1, double* g) {
    const double vi = __dmul_rn((double)(n - 1), __ddiv_rn(q, 100.0));
    int prev = (int)floor(vi);
    prev = max(0, min(prev, n - 1));
    *g = __dsub_rn(vi, (double)prev);
    *r0 = (uint32_t)prev;
    *r1 = (uint32_t)min(prev + 1, n - 1);
}
__device__ __forceinline__ double pct_lerp(float a, float b, double g) {
    const double d = (double)__fsub_rn(b, a);   // subtract(b, a) on float32
    double r = (g >= 0.5) ? __dsub_rn((double)b, __dmul_rn(d, __dsub_rn(1.0, g))) : __dadd_rn((double)a, __dmul_rn(d, g));
    if (b == a) r = (double)a;                  // _lerp: where(b == a, a, lerp)
    return r;
}

// Every selection of this kernel reads either p[i] or |p[i] - med| (float32): ONE out-of-line copy of the selection code
// serves all call sites.  (Inlined per call site the kernel was 58 800 SASS instructions and stalled on instruction
// fetch - "no instruction" 3.3 per issue in profiles/r02_ncu_full_validate_real_reads_before.json.)
struct ValSrc {
    const float* p;
    float med;
    int absdev;
    __device__ __forceinline__ float operator()(int i) const { return absdev ? fabsf(__fsub_rn(p[i], med)) : p[i]; }
};
__device__ __forceinline__ ValSrc val_src(const float* p) { return ValSrc{p, 0.f, 0}; }
__device__ __forceinline__ ValSrc val_src_absdev(const float* p, float med) { return ValSrc{p, med, 1}; }

__device__ __noinline__ void block_ranks(int n, ValSrc src, int nq, const uint32_t* ranks, float* out, ValSel& vs, FpScratch& s) {
    block_ranks_t(n, src, nq, ranks, out, vs, s);
}

// np.median of n NaN-free float32 values (NaN for n == 0, as numpy returns for an empty slice).
__device__ __forceinline__ float val_median(int n, ValSrc src, ValSel& vs, FpScratch& s) {
    if (n <= 0) return __int_as_float(0x7fc00000);
    const uint32_t rk[2] = {(uint32_t)((n - 1) / 2), (uint32_t)(n / 2)};
    float o[2];
    block_ranks(n, src, 2, rk, o, vs, s);
    return median_of(n, o[0], o[1]);
}

// p85 - p15 of n >= 1 values (np.subtract(*np.percentile(x, (85, 15)))), optionally with the median.
__device__ __forceinline__ double val_local_range(int n, ValSrc src, ValSel& vs, FpScratch& s, float* median) {
    uint32_t rk[6];
    double g85, g15;
    pct_ranks(n, 85.0, &rk[0], &rk[1], &g85);
    pct_ranks(n, 15.0, &rk[2], &rk[3], &g15);
    rk[4] = (uint32_t)((n - 1) / 2);
    rk[5] = (uint32_t)(n / 2);
    float o[6];
    block_ranks(n, src, median ? 6 : 4, rk, o, vs, s);
    if (median) *median = median_of(n, o[4], o[5]);
    return __dsub_rn(pct_lerp(o[0], o[1], g85), pct_lerp(o[2], o[3], g15));
}

// ---- numpy's pairwise sum of a long float32 array, in parallel ---------------------------------------------
// np.add.reduce halves the array recursively (left half rounded down to a multiple of 8) until a block has <= 128
// elements.  Thread 0 lists the leaf blocks in order (VAL_MAX_LEAVES covers any row that fits in shared memory), the
// CTA sums the leaves with the 8-accumulator block routine, thread 0 adds the leaf sums up the same tree.
constexpr int VAL_MAX_LEAVES = 1024;
struct ValTree {
    int off[VAL_MAX_LEAVES];
    float sum[VAL_MAX_LEAVES];
    short len[VAL_MAX_LEAVES];
    int n_leaves;
    float result;
};

__device__ void val_tree_build(int n, ValTree& t) {   // one thread
    int lo_s[24], n_s[24], sp = 0, nl = 0;
    lo_s[sp] = 0;
    n_s[sp++] = n;
    while (sp > 0) {
        const int lo = lo_s[--sp], m = n_s[sp];
        if (m <= 128) {
            if (nl < VAL_MAX_LEAVES) {
                t.off[nl] = lo;
                t.len[nl] = (short)m;
            }
            nl++;
        } else {
            int n2 = m / 2;
            n2 -= n2 % 8;
            lo_s[sp] = lo + n2;      // right half (popped second)
            n_s[sp++] = m - n2;
            lo_s[sp] = lo;           // left half (popped first)
            n_s[sp++] = n2;
        }
    }
    t.n_leaves = nl;
}

__device__ float val_tree_combine(int n, const ValTree& t) {   // one thread; same walk, leaf sums consumed in order
    struct Frame {
        int n, state;
        float left;
    };
    Frame st[24];
    int sp = 0, next = 0;
    st[sp++] = Frame{n, 0, 0.f};
    float result = 0.f;
    while (sp > 0) {
        Frame& f = st[sp - 1];
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.n <= 128) {
            result = t.sum[next++];
            sp--;
        } else if (f.state == 0) {
            f.state = 1;
            st[sp++] = Frame{n2, 0, 0.f};
        } else if (f.state == 1) {
            f.left = result;
            f.state = 2;
            st[sp++] = Frame{f.n - n2, 0, 0.f};
        } else {
            result = __fadd_rn(f.left, result);
            sp--;
        }
    }
    return result;
}

// Sum of f(lo + i), i < n, in numpy's order; the tree for this n must have been built (val_tree_build + barrier).
template <typename F>
__device__ float block_np_sum_f32(int lo, int n, F f, ValTree& t) {
    const int nl = t.n_leaves;
    for (int k = threadIdx.x; k < nl; k += FP_THREADS) t.sum[k] = np_pairwise_leaf<float>(lo + t.off[k], (int)t.len[k], f);
    __syncthreads();
    if (threadIdx.x == 0) t.result = val_tree_combine(n, t);
    __syncthreads();
    const float r = t.result;
    __syncthreads();
    return r;
}

// calc_partition_stats (signal_partitions.py:80-96) of sig[start:end] for a row of L samples -> out[6] (thread 0 writes).
__device__ __noinline__ void val_partition(const float* vsig, int L, int64_t start, int64_t end, double* out, ValTree& t, ValSel& vs, FpScratch& s) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int tid = threadIdx.x;
    if (end <= start) {
        if (tid == 0) {
            out[0] = (double)start;
            for (int j = 1; j < 6; j++) out[j] = qnan;
        }
        return;
    }
    const int b = (int)max((int64_t)0, min(start, (int64_t)L)), e = (int)max((int64_t)0, min(end, (int64_t)L));
    const int n = max(0, e - b);
    float mean, sd, med, mad;
    if (n == 0 || n > VAL_MAX_LEAVES * 64) {   // empty slice: numpy returns NaN for all four
        mean = sd = med = mad = __int_as_float(0x7fc00000);
    } else {
        __syncthreads();
        if (tid == 0) val_tree_build(n, t);
        __syncthreads();
        mean = __fdiv_rn(block_np_sum_f32(b, n, [&](int i) { return vsig[i]; }, t), (float)n);
        const float ss = block_np_sum_f32(b, n, [&](int i) {
            const float d = __fsub_rn(vsig[i], mean);
            return __fmul_rn(d, d);
        }, t);
        sd = __fsqrt_rn(__fdiv_rn(ss, (float)n));
        med = val_median(n, val_src(vsig + b), vs, s);
        mad = val_median(n, val_src_absdev(vsig + b, med), vs, s);
    }
    if (tid == 0) {
        out[0] = (double)start;
        out[1] = (double)(end - start);
        out[2] = (double)mean;
        out[3] = (double)sd;
        out[4] = (double)med;
        out[5] = (double)mad;
    }
}

// ---- median of a moving-window statistic, exactly, without evaluating the statistic exactly everywhere ---------------
// mean_var_shift_polyA_check takes np.nanmedian over bottleneck.move_var(seg, 100) / move_mean(seg, 20) (mvs.py:93-107): one
// value per window position, each an O(window) float64 reduction in numpy's pairwise order (two passes for the variance).
// On real reads the poly(A) stretch is ~4 800 samples, i.e. ~1.9 M float64 operations + 0.9 M float32->float64
// conversions per read for a single median — half of this kernel's time.  Only the MIDDLE order statistics are needed:
//  (1) every window gets a cheap approximation A[p] (sliding sums of the centred samples in float64, restarted every
//      VAL_WIN_CHUNK windows; relative error < 1e-8, stored as float32),
//  (2) the approximate middle values a are selected from A,
//  (3) windows with A outside a band of relative half-width 1e-6 around a are certainly below / above the exact middle
//      values (an order statistic moves by at most the largest perturbation of the data, and float32 rounding of the exact
//      value is monotone); only the few windows inside the band are evaluated exactly,
//  (4) the exact middle values are the (k - #below)-th smallest exact values inside the band.
// The result is bit-identical to selecting from the exactly evaluated array (tests: tests/test_validate.py against the
// reference-generated fixture and the CPU restatement).  Degenerate data (band overflows VAL_BAND_CAP, non-finite
// approximations) falls back to evaluating every window exactly.
constexpr int VAL_WIN_CHUNK = 8;
constexpr int VAL_BAND_CAP = 384;
struct ValBand {
    int idx[VAL_BAND_CAP];
    float val[VAL_BAND_CAP];
    int n, below;
    float out[2];
};

template <bool VAR>
__device__ __forceinline__ float val_window_exact(const float* vsig, int lo, int w) {
    const double mu = __ddiv_rn(np_pairwise<double>(lo, w, [&](int i) { return (double)vsig[i]; }), (double)w);
    if (!VAR) return (float)mu;
    const double ss = np_pairwise<double>(lo, w, [&](int i) {
        const double d = __dsub_rn((double)vsig[i], mu);
        return __dmul_rn(d, d);
    });
    return (float)__ddiv_rn(ss, (double)w);
}

// np.nanmedian(move_var / move_mean (vsig[e : e + m], w)) as float32 (NaN when no window fits).  scratch: >= m floats.
template <bool VAR>
__device__ __noinline__ float val_window_median(const float* vsig, int e, int m, int w, float* scratch, ValBand& bd, ValSel& vs, FpScratch& s) {
    const int tid = threadIdx.x;
    const int cnt = m - w + 1;
    if (cnt <= 0) return __int_as_float(0x7fc00000);
    __syncthreads();
    bool exact_all = cnt <= 4 * VAL_BAND_CAP / 3 || w < 4;
    if (!exact_all) {
        const double cref = (double)vsig[e], wd = (double)w;
        for (int p0 = tid * VAL_WIN_CHUNK; p0 < cnt; p0 += FP_THREADS * VAL_WIN_CHUNK) {
            double s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < w; i++) {
                const double d = (double)vsig[e + p0 + i] - cref;
                s1 += d;
                if (VAR) s2 += d * d;
            }
            const int pe = min(p0 + VAL_WIN_CHUNK, cnt);
            for (int p = p0;; p++) {
                const double mu = s1 / wd;
                scratch[p] = VAR ? (float)(s2 / wd - mu * mu) : (float)(cref + mu);
                if (p + 1 >= pe) break;
                const double dn = (double)vsig[e + p + w] - cref, dl = (double)vsig[e + p] - cref;
                s1 += dn - dl;
                if (VAR) s2 += dn * dn - dl * dl;
            }
        }
        __syncthreads();
        const uint32_t rk[2] = {(uint32_t)((cnt - 1) / 2), (uint32_t)(cnt / 2)};
        float am[2];
        block_ranks(cnt, val_src(scratch), 2, rk, am, vs, s);
        // half-width of the band: 1e-6 relative — the float32 storage of A (6e-8) and the sliding float64 sums of the centred
        // samples (< 1e-10 of the statistic for pA-scale data) stay far inside it; a statistic that is ~0 against data
        // that is not (constant signal) puts every window into the band and takes the exact path below
        const double a_lo = (double)fminf(am[0], am[1]), a_hi = (double)fmaxf(am[0], am[1]);
        const double tol = 1e-6 * fmax(fabs(a_lo), fabs(a_hi)) + 1e-12;
        const double b_lo = a_lo - tol, b_hi = a_hi + tol;
        if (tid == 0) {
            bd.n = 0;
            bd.below = 0;
        }
        __syncthreads();
        int below = 0;
        for (int p = tid; p < cnt; p += FP_THREADS) {
            const double A = (double)scratch[p];
            if (A < b_lo) below++;
            else if (A <= b_hi) {
                const int pos = atomicAdd(&bd.n, 1);
                if (pos < VAL_BAND_CAP) bd.idx[pos] = p;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if ((tid & 31) == 0 && below) atomicAdd(&bd.below, below);
        __syncthreads();
        const int nb = bd.n, r0 = (int)rk[0] - bd.below, r1 = (int)rk[1] - bd.below;
        if (!(a_lo == a_lo) || !(a_hi - a_lo <= 1e300) || nb > VAL_BAND_CAP || r0 < 0 || r1 >= nb) {
            exact_all = true;    // uniform decision
        } else {
            for (int j = tid; j < nb; j += FP_THREADS) bd.val[j] = val_window_exact<VAR>(vsig, e + bd.idx[j], w);
            __syncthreads();
            for (int j = tid; j < nb; j += FP_THREADS) {
                const float x = bd.val[j];
                int rank = 0;
                for (int u = 0; u < nb; u++) {
                    const float y = bd.val[u];
                    rank += (y < x) || (y == x && u < j);
                }
                if (rank == r0) bd.out[0] = x;
                if (rank == r1) bd.out[1] = x;
            }
            __syncthreads();
            const float res = median_of(cnt, bd.out[0], bd.out[1]);
            __syncthreads();
            return res;
        }
        __syncthreads();
    }
    for (int p = tid; p < cnt; p += FP_THREADS) scratch[p] = val_window_exact<VAR>(vsig, e + p, w);
    __syncthreads();
    return val_median(cnt, val_src(scratch), vs, s);
}

__device__ __forceinline__ bool val_in_range(double v, double lo, double hi) { return lo <= v && v <= hi; }

__global__ void __launch_bounds__(FP_THREADS, WDX_VAL_MIN_CTAS) validate_kernel(const ValArgs a, const ValCfg c) {
#ifndef WDX_VAL_GLOBAL_ROW
    extern __shared__ float vsig[];
#endif
    __shared__ FpScratch s;
    __shared__ ValSel vs;
    __shared__ ValTree tree;
    __shared__ ValBand band;
    __shared__ int sh_i[6];
    __shared__ int sh_pores[VAL_PORES_LD];
    __shared__ double sh_d[4];
    __shared__ double sh_v[VAL_NVALS];
    const int tid = threadIdx.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    float* scratch = a.scratch + (size_t)blockIdx.x * a.stride;

    __shared__ unsigned long long sh_next;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            unsigned long long nx = atomicAdd(a.next, 1ULL);
            if (a.list) nx = (nx < (unsigned long long)*a.list_count) ? (unsigned long long)a.list[nx] : (unsigned long long)a.n;
            sh_next = nx;
        }
        __syncthreads();
        const int64_t r = (int64_t)sh_next;
        if (r >= a.n) break;
        const float* row = a.signals + (size_t)r * a.stride;
        const int64_t fl = a.full_len[r];
        const int L = (int)max((int64_t)0, min(fl, a.stride));
        int has_nan = 0;
        if (tid < VAL_NVALS) sh_v[tid] = qnan;
#ifdef WDX_VAL_GLOBAL_ROW
        // The row stays in global memory: this NaN scan is its one trip from HBM (coalesced), every later pass finds it in
        // L2 (46 KB per read, ~1200 reads in flight).  No shared-memory staging = small CTAs, eight per SM, and the many
        // short barrier-separated phases of one read overlap with those of seven others.
        const float* vsig = row;
        for (int i0 = tid; i0 < L; i0 += 8 * FP_THREADS) {
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * FP_THREADS;
                xv[u] = (i < L) ? __ldg(row + i) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) has_nan |= (xv[u] != xv[u]);
        }
#else
        for (int i0 = tid; i0 < L; i0 += 4 * FP_THREADS) {   // four loads in flight per thread before the first use
            float xv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * FP_THREADS;
                xv[u] = (i < L) ? __ldg(row + i) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * FP_THREADS;
                if (i < L) {
                    vsig[i] = xv[u];
                    has_nan |= (xv[u] != xv[u]);
                }
            }
        }
#endif
        has_nan = __syncthreads_or(has_nan);

        const int64_t* pr = a.preds + (size_t)r * a.ld;
        const int64_t a1 = pr[0];
        int64_t a0 = 0;
        int64_t pe_best = a.ld > 1 ? pr[1] : 0;
        int code = VAL_OK, checks = 0, n_pores = 0;
        int pores_cnt = -1, pores_single = 0;   // open_pores as the reference reports it: None / the kept positions
        // reported statistics live in shared memory (thread 0 writes them as they are found); the fill above is
        // ordered before those writes by the barrier of the NaN vote

        if (has_nan) code = VAL_HAS_NAN;
        const int hi = (int)max((int64_t)0, min(a1, (int64_t)L));   // sig[a0:a1] ends here
        float med = 0.f, mad = 0.f;
        if (code == VAL_OK) {
            if (a1 == 0) code = VAL_NO_ADAPTER;
            else {
                med = val_median(hi, val_src(vsig), vs, s);
                mad = val_median(hi, val_src_absdev(vsig, med), vs, s);
                if (tid == 0) {
                    sh_v[0] = (double)med;
                    sh_v[1] = (double)mad;
                }
            }
        }
        if (code == VAL_OK && mad != 0.f && !val_in_range((double)mad, c.mad_lo, c.mad_hi)) code = VAL_ADAPTER_MAD;

        if (code == VAL_OK && c.detect_open_pores) {
            __syncthreads();
            if (tid == 0) {
                sh_i[0] = 0;           // open-pore samples
                sh_i[1] = 0x7fffffff;  // first
                sh_i[2] = -1;          // last
                sh_i[3] = -1;          // last kept
                sh_i[4] = 0;           // kept
            }
            __syncthreads();
            const float lo = c.open_pore_min;
            for (int i = tid; i < hi; i += FP_THREADS) {
                if (vsig[i] >= lo) {
                    atomicAdd(&sh_i[0], 1);
                    atomicMin(&sh_i[1], i);
                    atomicMax(&sh_i[2], i);
                }
            }
            __syncthreads();
            const int first = sh_i[1];
            const int D = c.open_pore_min_obs_diff;
            for (int i = tid; i < hi; i += FP_THREADS) {
                if (vsig[i] >= lo && first < i) {
                    bool near = false;
                    for (int q = max(0, i - D + 1); q < i; q++) near |= (vsig[q] >= lo);
                    if (!near) {
                        atomicMax(&sh_i[3], i);
                        const int slot = atomicAdd(&sh_i[4], 1);
                        if (slot < VAL_PORES_LD - 1) sh_pores[1 + slot] = i;
                    }
                }
            }
            __syncthreads();
            const int cnt = sh_i[0];
            if (cnt > 1) {
                n_pores = sh_i[4] > 0 ? sh_i[4] : 1;
                a0 = sh_i[4] > 0 ? sh_i[3] : sh_i[2];
            } else if (cnt == 1) {
                n_pores = 1;
                a0 = first;
            }
            if (cnt > 0 && a1 - a0 < c.min_obs_adapter) code = VAL_OPEN_PORE;
            // find_open_pores(...).ravel() (anomalies.py:16-35): the kept positions, else the last open-pore sample, else
            // the single one, else an empty array
            pores_cnt = cnt > 1 ? (sh_i[4] > 0 ? sh_i[4] : 1) : cnt;
            pores_single = (cnt > 1 && sh_i[4] > 0) ? -1 : (int)a0;
        }

        if (code == VAL_OK && c.real_signal_check) {
            const int b0 = (int)min(a0, (int64_t)hi);
            const int nseg = hi - b0;
            bool ok = false;
            if (nseg >= 2 * c.mean_window) {
                __syncthreads();
                if (tid == 0 || tid == 32) {
                    const int base = tid == 0 ? b0 : hi - c.mean_window;
                    const float sm = np_pairwise<float>(base, c.mean_window, [&](int i) { return vsig[i]; });
                    sh_d[tid == 0 ? 0 : 1] = (double)__fdiv_rn(sm, (float)c.mean_window);
                }
                __syncthreads();
                const double mean_start = sh_d[0], mean_end = sh_d[1];
                if (tid == 0) {
                    sh_v[2] = mean_start;
                    sh_v[3] = mean_end;
                }
                if (val_in_range(mean_start, c.mean_start_lo, c.mean_start_hi) && val_in_range(mean_end, c.mean_end_lo, c.mean_end_hi)) {
                    const int nn = min(c.max_obs_local_range, nseg);
                    const int base = hi - nn;
                    const double lr = val_local_range(nn, val_src(vsig + base), vs, s, nullptr);
                    if (tid == 0) sh_v[4] = lr;
                    ok = val_in_range(lr, c.local_range_lo, c.local_range_hi);
                }
            }
            if (!ok) code = VAL_REAL_RANGE;
        }

        if (code == VAL_OK && c.mvs_detect_check) {
            if (pe_best == 0) code = VAL_NO_POLYA;
            else {
                double mlo = c.mean_lo, mhi = c.mean_hi;
                if (c.mean_from_scale) {
                    mlo = __dmul_rn(c.scale_lo, (double)med);
                    mhi = __dmul_rn(c.scale_hi, (double)med);
                }
                const int e = (int)a1;   // a1 <= L here or the size test below fails first
                bool have_shift = false;
                double shift_cached = 0.0;
                for (int j = 1; j < a.ld; j++) {
                    const int64_t pe = pr[j];
                    if (pe == 0) break;
                    double r_mean = 0.0, r_var = 0.0, r_med = 0.0, r_lr = 0.0, r_shift = 0.0;
                    bool okc = false;
                    int bits = 0;
                    const bool early = pe < a1 || pe - a1 <= 2 || (int64_t)L < a1 + c.median_shift_window;
                    if (!early) {
                        const int pend = (int)min(pe, (int64_t)L);
                        const int m = pend - e;
                        const int64_t nominal = pe - a1;
                        // variance
                        if (nominal <= c.pa_var_window + 2) {
                            __syncthreads();
                            if (tid == 0) {
                                const float mu = __fdiv_rn(np_pairwise<float>(e, m, [&](int i) { return vsig[i]; }), (float)m);
                                const float ss = np_pairwise<float>(e, m, [&](int i) {
                                    const float d = __fsub_rn(vsig[i], mu);
                                    return __fmul_rn(d, d);
                                });
                                sh_d[0] = (double)__fdiv_rn(ss, (float)m);
                            }
                            __syncthreads();
                            r_var = sh_d[0];
                        } else {
                            r_var = (double)val_window_median<true>(vsig, e, m, c.pa_var_window, scratch, band, vs, s);
                        }
                        // mean
                        if (nominal <= c.pa_mean_window + 2) {
                            __syncthreads();
                            if (tid == 0)
                                sh_d[0] = (double)__fdiv_rn(np_pairwise<float>(e, m, [&](int i) { return vsig[i]; }), (float)m);
                            __syncthreads();
                            r_mean = sh_d[0];
                        } else {
                            r_mean = (double)val_window_median<false>(vsig, e, m, c.pa_mean_window, scratch, band, vs, s);
                        }
                        float pmed;
                        r_lr = val_local_range(m, val_src(vsig + e), vs, s, &pmed);
                        r_med = (double)pmed;
                        if (!have_shift) {   // depends on the adapter end only: once per read, not per candidate
                            const int up = min(e + c.median_shift_window, L), dn = max(e - c.median_shift_window, 0);
                            const float m_after = val_median(up - e, val_src(vsig + e), vs, s);
                            const float m_before = val_median(e - dn, val_src(vsig + dn), vs, s);
                            shift_cached = (double)__fsub_rn(m_after, m_before);
                            have_shift = true;
                        }
                        r_shift = shift_cached;
                        bits = (val_in_range(r_mean, mlo, mhi) ? 1 : 0) | (val_in_range(r_var, c.var_lo, c.var_hi) ? 2 : 0) |
                               (val_in_range(r_med, c.pmed_lo, c.pmed_hi) ? 4 : 0) | (val_in_range(r_lr, c.plr_lo, c.plr_hi) ? 8 : 0) |
                               (val_in_range(r_shift, c.shift_lo, c.shift_hi) ? 16 : 0);
                        okc = bits == 31;
                    }
                    if (tid == 0) {
                        sh_v[5] = r_mean;
                        sh_v[6] = r_var;
                        sh_v[7] = r_med;
                        sh_v[8] = r_lr;
                        sh_v[9] = r_shift;
                    }
                    if (!okc) {   // `success` is never set back to True: later candidates only overwrite the report
                        if (r_mean == 0.0) {
                            code = VAL_MVS_NO_SIGNAL;
                            checks = 0;
                        } else {
                            code = VAL_MVS_CHECKS;
                            checks = bits;
                        }
                    }
                    if (code == VAL_OK) {
                        pe_best = pe;
                        break;
                    }
                    if (a.verdict_only) break;
                }
            }
        }

        if (code == VAL_OK && c.detect_med_shift) {
            const int e = (int)min(a1, (int64_t)L);
            const int up = (int)min(min(a1 + c.med_shift_window, fl), (int64_t)L), dn = (int)max(a1 - c.med_shift_window, (int64_t)0);
            const float m_after = val_median(max(0, up - e), val_src(vsig + e), vs, s);
            const float m_before = val_median(max(0, e - min(dn, e)), val_src(vsig + min(dn, e)), vs, s);
            const double ms = (double)__fsub_rn(m_after, m_before);
            if (tid == 0) sh_v[10] = ms;
            if (!val_in_range(ms, c.ms_lo, c.ms_hi)) code = VAL_MED_SHIFT;
        }

        if (a.fail_list && tid == 0 && code != VAL_OK && code != VAL_HAS_NAN) a.fail_list[atomicAdd(a.fail_count, 1)] = (int)r;
        if (a.commit_on_success && code != VAL_OK) continue;
        if (a.parts) {   // calc_partitions_from_vals(signal, adapter_start, adapter_end, polya_end_best), combined.py:631-636
            double* po = a.parts + r * VAL_NPART;
            if (code == VAL_HAS_NAN) {
                if (tid < VAL_NPART) po[tid] = qnan;
            } else {
                val_partition(vsig, L, a0, a1, po, tree, vs, s);
                val_partition(vsig, L, a1, pe_best, po + 6, tree, vs, s);
                val_partition(vsig, L, pe_best, (int64_t)L, po + 12, tree, vs, s);
            }
        }
        if (a.pores) {
            __syncthreads();
            if (tid == 0) {
                int32_t* po = a.pores + r * VAL_PORES_LD;
                po[0] = pores_cnt;
                if (pores_cnt > 0 && pores_single >= 0) po[1] = pores_single;
                else if (pores_cnt > 0) {   // found in parallel: a handful of entries, insertion sort
                    const int m = min(pores_cnt, VAL_PORES_LD - 1);
                    for (int i = 1; i < m; i++) {
                        const int v = sh_pores[1 + i];
                        int j = i - 1;
                        for (; j >= 0 && sh_pores[1 + j] > v; j--) sh_pores[2 + j] = sh_pores[1 + j];
                        sh_pores[2 + j] = v;
                    }
                    for (int i = 0; i < m; i++) po[1 + i] = sh_pores[1 + i];
                }
            }
        }
        if (tid == 0) {
            a.success[r] = code == VAL_OK;
            a.info[r * 4 + 0] = code;
            a.info[r * 4 + 1] = checks;
            a.info[r * 4 + 2] = n_pores;
            a.info[r * 4 + 3] = a.list ? ((a.info[r * 4 + 3] & ~3) | a.src_tag) : 0;
            a.bounds[r * 3 + 0] = a0;
            a.bounds[r * 3 + 1] = a1;
            a.bounds[r * 3 + 2] = pe_best;
            if (a.vals)
                for (int j = 0; j < VAL_NVALS; j++) a.vals[r * VAL_NVALS + j] = sh_v[j];
        }
    }
}

}  // namespace wdx
