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

// ---- several selections at once -------------------------------------------------------------------------------------
// The checks of one read need ~10 exact order statistics over different ranges of the same row.  Run one after the
// other (block_ranks) each costs ~10 CTA-wide barriers for a few microseconds of work: the kernel was barrier- and
// latency-bound (profiles/r02_ncu_full_validate_real_reads_before.json).  Here the independent selections of a read
// are JOBS of one round: every pass (histogram, gather) walks all jobs back to back with no barrier in between, the
// per-job histogram scans and the per-rank exact rankings run on different warps at the same time.  A job is a range
// of samples, |sample - median|, 16-bit codes, or the exact moving mean of the range; bins are a monotone map of the
// value (linear over bounds of the range known beforehand, so no min / max pass), the members of the bins that hold the
// wanted ranks are gathered and ranked exactly on their order keys.  Crowded bins fall back to radix selection.
constexpr int MS_BINS = 1024;          // 16-bit counters, two per word
constexpr int MS_MAXJ = 9;             // jobs per round
constexpr int MS_MAXQ = 24;            // ranks per round
constexpr int MS_MAXS = 16;            // distinct (job, bin) candidate lists per round
constexpr int MS_CAND = 224;           // members kept per list
constexpr int MS_CHUNKS = (MS_MAXS * MS_CAND * 4) / 16;   // chunk sums (two doubles each) of the moving variance alias the candidate lists

struct MsJob {
    const float* p;               // first sample of the range
    int n;                        // values
    int mode;                     // 0 sample, 1 |sample - med|
    float med, vmin, scale;       // bin = (value - vmin) * scale
    int q0, nq;                   // ranks q0 .. q0 + nq - 1
};

struct MsState {
    uint32_t hist[MS_MAXJ][MS_BINS / 2];
    uint32_t cand[MS_MAXS][MS_CAND];
    MsJob job[MS_MAXJ];
    uint32_t slot_n[MS_MAXS];
    uint32_t q_rank[MS_MAXQ], q_r[MS_MAXQ], q_cnt[MS_MAXQ], q_key[MS_MAXQ];
    int q_job[MS_MAXQ], q_bin[MS_MAXQ], q_slot[MS_MAXQ];    // q_slot: candidate list, -1 = crowded (radix fallback)
    int nj, nq, n_slots, crowded;
};

template <int MODE>
__device__ __forceinline__ float ms_value(const MsJob& j, int i) {
    const float x = j.p[i];
    return (MODE == 1) ? fabsf(__fsub_rn(x, j.med)) : x;
}
__device__ __forceinline__ int ms_bin(const MsJob& j, float x) {
    return max(0, min(MS_BINS - 1, (int)__fmul_rn(__fsub_rn(x, j.vmin), j.scale)));
}

template <int MODE>
__device__ __forceinline__ void ms_hist_pass(const MsJob& j, uint32_t* hist) {
    const int n = j.n;
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * FP_THREADS) {     // four independent elements per trip
        int bin[4];
#pragma unroll
        for (int u = 0; u < 4; u++) bin[u] = ms_bin(j, ms_value<MODE>(j, min(i0 + u * FP_THREADS, n - 1)));
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + u * FP_THREADS < n) atomicAdd(&hist[bin[u] >> 1], (bin[u] & 1) ? 0x10000u : 1u);
    }
}

// After the scan a job's histogram holds, per bin, 0 or 1 + the candidate list its members go to.
template <int MODE>
__device__ __forceinline__ void ms_gather_pass(const MsJob& j, MsState& ms, const uint32_t* hist) {
    const unsigned short* to_list = reinterpret_cast<const unsigned short*>(hist);
    const int n = j.n;
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * FP_THREADS) {
        float x[4];
        int sl[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x[u] = ms_value<MODE>(j, min(i0 + u * FP_THREADS, n - 1));
            sl[u] = (i0 + u * FP_THREADS < n) ? to_list[ms_bin(j, x[u])] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (sl[u]) {
                const uint32_t pos = atomicAdd(&ms.slot_n[sl[u] - 1], 1u);
                if (pos < (uint32_t)MS_CAND) ms.cand[sl[u] - 1][pos] = f32_key(x[u]);
            }
    }
}

// One warp: exclusive scan of a job's histogram, the bin / rank inside the bin / population of every wanted rank, and
// the candidate lists (ranks of one job that fall into the same bin share a list).  A job has at most six ranks.
__device__ __forceinline__ void ms_scan_job(MsState& ms, int jj) {
    const int lane = threadIdx.x & 31;
    const MsJob& j = ms.job[jj];
    constexpr int WPL = MS_BINS / 2 / 32;     // words per lane
    uint32_t* h = ms.hist[jj] + lane * WPL;
    uint32_t w[WPL], sum = 0;
#pragma unroll
    for (int t = 0; t < WPL; t++) {
        w[t] = h[t];
        sum += (w[t] & 0xffffu) + (w[t] >> 16);
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t run0 = inc - sum;
    // The lane's 32 counts become inclusive prefixes IN PLACE (16-bit pairs again: a lane holds fewer than 2^16 values), so
    // that the bin of a rank is found by the whole warp at once — lane t looks at bin t of the lane whose range holds the
    // rank — instead of one lane walking its 32 bins through a chain of dependent additions per rank.
    {
        uint32_t run = 0;
#pragma unroll
        for (int t = 0; t < WPL; t++) {
            run += w[t] & 0xffffu;
            const uint32_t p0 = run;
            run += w[t] >> 16;
            h[t] = p0 | (run << 16);
        }
    }
    __syncwarp();
    const unsigned short* pre = reinterpret_cast<const unsigned short*>(ms.hist[jj]);
    const int q0 = j.q0, nq = j.nq;
    // lane q < nq owns rank q
    const uint32_t myk = (lane < nq) ? ms.q_rank[q0 + lane] : 0xffffffffu;
    int mybin = -1;
    uint32_t myr = 0, mycnt = 0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        const uint32_t k = __shfl_sync(0xffffffffu, myk, q);
        if (q >= nq) continue;              // uniform
        const bool hit = k >= run0 && k < run0 + sum;
        const unsigned who = __ballot_sync(0xffffffffu, hit);
        int bin = -1;
        uint32_t r = 0, cnt = 0;
        if (who) {                          // uniform
            const int src = __ffs(who) - 1;
            const uint32_t kl = k - __shfl_sync(0xffffffffu, run0, src);    // rank inside that lane's 32 bins
            const uint32_t incl = pre[src * 2 * WPL + lane];
            const uint32_t excl = lane ? pre[src * 2 * WPL + lane - 1] : 0u;
            const unsigned who2 = __ballot_sync(0xffffffffu, kl >= excl && kl < incl);   // exactly one lane (kl < that lane's sum)
            const int l2 = __ffs(who2) - 1;
            bin = src * 2 * WPL + l2;
            r = kl - __shfl_sync(0xffffffffu, excl, l2);
            cnt = __shfl_sync(0xffffffffu, incl - excl, l2);
        }
        if (lane == q) {
            mybin = bin;
            myr = r;
            mycnt = cnt;
        }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < WPL; t++) h[t] = 0;  // from here on: bin -> 1 + candidate list (ms_gather_pass)
    // candidate lists: the first rank of a bin allocates, the others share
    int first = -1;
#pragma unroll
    for (int u = 0; u < 6; u++) {
        const int b = __shfl_sync(0xffffffffu, mybin, u);
        if (u < lane && lane < nq && b == mybin && first < 0) first = u;
    }
    int sl = -1;
    if (lane < nq && first < 0 && mybin >= 0 && mycnt <= (uint32_t)MS_CAND) {
        sl = atomicAdd(&ms.n_slots, 1);
        if (sl >= MS_MAXS) sl = -1;
    }
    const int shared_sl = __shfl_sync(0xffffffffu, sl, first < 0 ? 0 : first);
    if (first >= 0) sl = shared_sl;
    __syncwarp();                            // the histogram words are cleared
    if (lane < nq) {
        ms.q_bin[q0 + lane] = mybin;
        ms.q_r[q0 + lane] = myr;
        ms.q_cnt[q0 + lane] = mycnt;
        ms.q_slot[q0 + lane] = sl;
        if (sl < 0) ms.crowded = 1;
        else if (first < 0) reinterpret_cast<unsigned short*>(ms.hist[jj])[mybin] = (unsigned short)(sl + 1);
    }
}

// The q_r-th smallest key of every rank's candidate list: the (list, candidate) pairs are dealt out over the whole CTA,
// a candidate's rank inside its list is matched against every rank that reads the list.
__device__ __forceinline__ void ms_rank_all(MsState& ms) {
    const int nq = ms.nq, ns = min(ms.n_slots, MS_MAXS);
    int total = 0;
    for (int sl = 0; sl < ns; sl++) total += (int)min(ms.slot_n[sl], (uint32_t)MS_CAND);
    for (int idx = threadIdx.x; idx < total; idx += FP_THREADS) {
        int sl = 0, t = idx;
        for (;; sl++) {
            const int m = (int)min(ms.slot_n[sl], (uint32_t)MS_CAND);
            if (t < m) break;
            t -= m;
        }
        const uint32_t m = min(ms.slot_n[sl], (uint32_t)MS_CAND);
        const uint32_t* cd = ms.cand[sl];
        const uint32_t x = cd[t];
        uint32_t rank = 0;
#pragma unroll 4
        for (uint32_t u = 0; u < m; u++) {
            const uint32_t y = cd[u];
            rank += (y < x) || (y == x && u < (uint32_t)t);
        }
        for (int q = 0; q < nq; q++)
            if (ms.q_slot[q] == sl && ms.q_r[q] == rank) ms.q_key[q] = x;
    }
}

// A round: jobs and ranks are in `ms` (written before the caller's last barrier), histograms / list counters are zero.
// On return (after a barrier) ms.q_key[q] holds the order key (or the code) of every rank.  `extra_*` let the caller run
// its own work inside the phases of the round (no barrier of its own needed).
template <typename F1, typename F2, typename F3>
__device__ __forceinline__ void ms_round(MsState& ms, FpScratch& s, F1 extra_hist, F2 extra_scan, F3 extra_gather, int pb = 5) {
    const int warp = threadIdx.x >> 5;
    const int nj = ms.nj, nq = ms.nq;
    (void)pb;   // phase-clock slots of this round (WDX_FP_PROF builds)
    for (int jj = 0; jj < nj; jj++) {
        const MsJob& j = ms.job[jj];
        if (j.mode == 0) ms_hist_pass<0>(j, ms.hist[jj]);
        else ms_hist_pass<1>(j, ms.hist[jj]);
    }
    extra_hist();
    __syncthreads();
    FP_T(s, pb + 0);   // histogram pass
    if (warp < nj) ms_scan_job(ms, warp);
    extra_scan();
    __syncthreads();
    FP_T(s, pb + 1);   // scans
    for (int jj = 0; jj < nj; jj++) {
        const MsJob& j = ms.job[jj];
        if (j.mode == 0) ms_gather_pass<0>(j, ms, ms.hist[jj]);
        else ms_gather_pass<1>(j, ms, ms.hist[jj]);
    }
    extra_gather();
    __syncthreads();
    FP_T(s, pb + 2);   // gather pass
    ms_rank_all(ms);
    __syncthreads();
    FP_T(s, pb + 3);   // ranking
    if (ms.crowded) {     // uniform; degenerate data only
        for (int q = 0; q < nq; q++) {
            if (ms.q_slot[q] >= 0) continue;
            const MsJob j = ms.job[ms.q_job[q]];
            const uint32_t k = ms.q_rank[q];
            auto key_of = [&](int i) { return f32_key(j.mode == 0 ? ms_value<0>(j, i) : ms_value<1>(j, i)); };
            const uint32_t key = block_select_u32(j.n, k, key_of, s);
            __syncthreads();
            if (threadIdx.x == 0) ms.q_key[q] = key;
        }
        __syncthreads();
    }
}

// job / rank bookkeeping: job jj with its ranks at q0 .. (one thread per job; the layout is known to every thread)
__device__ __forceinline__ void ms_set_job(MsState& ms, int jj, const float* p, int n, int mode, float med, float vmin, float vmax,
                                           int q0, int nq, const uint32_t* ranks) {
    MsJob& j = ms.job[jj];
    j.p = p;
    j.n = n;
    j.mode = mode;
    j.med = med;
    j.vmin = vmin;
    const float d = __fsub_rn(vmax, vmin);
    float sc = (d > 0.f) ? __fdiv_rn((float)MS_BINS, d) : 0.f;
    if (!(sc < 1e30f)) sc = 0.f;       // a point range or an overflowing scale: everything lands in bin 0 (radix fallback)
    j.scale = sc;
    j.q0 = q0;
    j.nq = nq;
    for (int t = 0; t < nq; t++) {
        ms.q_rank[q0 + t] = ranks[t];
        ms.q_job[q0 + t] = jj;
        ms.q_slot[q0 + t] = -1;
        ms.q_key[q0 + t] = 0;
    }
}

// np.add.reduce of n <= 512 float32 values in numpy's pairwise order by ONE warp: the recursion splits down to at most
// four blocks of <= 128 elements, each summed with numpy's eight strided accumulators (eight lanes per block), combined
// in numpy's tree, then the sequential tail.  All lanes get the result.  (One thread: 300 dependent additions.)
template <typename F>
__device__ __forceinline__ float warp_np_sum_f32(int lo, int n, F at) {
    const int lane = threadIdx.x & 31, blk = lane >> 3, k = lane & 7;
    // blocks 0, 1 = the left half (block 1 empty unless the half is longer than 128), blocks 2, 3 = the right half
    int off[4] = {0, 0, 0, 0}, len[4] = {0, 0, 0, 0};
    if (n <= 128) {
        len[0] = n;
    } else {
        int h = n / 2;
        h -= h % 8;
        if (h <= 128) {
            len[0] = h;
        } else {
            int q = h / 2;
            q -= q % 8;
            len[0] = q;
            off[1] = q;
            len[1] = h - q;
        }
        const int r = n - h;
        off[2] = h;
        if (r <= 128) {
            len[2] = r;
        } else {
            int q2 = r / 2;
            q2 -= q2 % 8;
            len[2] = q2;
            off[3] = h + q2;
            len[3] = r - q2;
        }
    }
    const int b0 = lo + (blk == 0 ? off[0] : blk == 1 ? off[1] : blk == 2 ? off[2] : off[3]);
    const int m = blk == 0 ? len[0] : blk == 1 ? len[1] : blk == 2 ? len[2] : len[3];
    const int m8 = (m < 8) ? 0 : m - (m % 8);
    float r = 0.f;
    if (m8 > 0) {
        r = at(b0 + k);
        for (int i = 8; i < m8; i += 8) r = __fadd_rn(r, at(b0 + i + k));
    }
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));      // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7))
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    float res = r;                                             // 0 for a block of fewer than eight elements
    for (int i = m8; i < m; i++) res = __fadd_rn(res, at(b0 + i));
    const float s0 = __shfl_sync(0xffffffffu, res, 0), s1 = __shfl_sync(0xffffffffu, res, 8), s2 = __shfl_sync(0xffffffffu, res, 16),
                s3 = __shfl_sync(0xffffffffu, res, 24);
    const float left = len[1] ? __fadd_rn(s0, s1) : s0;
    if (n <= 128) return left;
    const float right = len[3] ? __fadd_rn(s2, s3) : s2;
    return __fadd_rn(left, right);
}

// ---- the common read, all checks in two rounds --------------------------------------------------------------------
// A read whose CNN boundaries allow every check to run (adapter end inside the row, first poly(A) candidate long
// enough for the windowed statistics) is validated speculatively: all statistics of the checks are computed in two
// rounds of concurrent selections, then the verdict is derived in the reference's order, so that code / check bits /
// reported values are those of the sequential evaluation.  Anything else (and further poly(A) candidates after a failed
// first one) takes the sequential path of validate_kernel.
//
// The medians of the moving variance / moving mean (mvs.py:93-107): every window gets a cheap approximation (sliding
// float64 sums; the variance's seeded from chunk sums), reduced to an 8-BIT BIN that is kept on chip: bin 0 / 255 =
// below / above the FOCUS, the range between the 25 % and 75 % quantiles of 64 sampled windows, bins 1..254 linear inside
// it.  A 256-bin histogram (filled while the bins are made) gives the bins of the middle ranks; the windows of those
// bins +- 1 are the band that is evaluated exactly and ranked, everything else is certainly below / above (a bin is
// orders of magnitude wider than the approximation error; same argument as val_window_median).  Result = the exact
// float32 order statistic, bit for bit.
struct ValFastOut {      // written by thread 0, read by every thread after the closing barrier
    int code, checks, n_pores, pores_cnt, pores_single, resume;
    long long a0, pe_best;
    float med;
    double shift;
};

struct ValSeg {          // order keys of the minimum / maximum of [0, a1), [a1, pend), [pend, L)
    uint32_t mn[3], mx[3];
};

struct ValCode8 {        // per code job (0 moving variance, 1 moving mean)
    uint32_t hist[2][128];          // 256 bins, 16-bit counters
    int b_lo[2], b_hi[2], below[2]; // bins of the two middle ranks; windows in bins below b_lo - 1
};

__device__ __forceinline__ bool val_fast_eligible(const ValCfg& c, int L, int64_t stride, const int64_t* pr, int ld) {
    if (!c.mvs_detect_check || ld < 2) return false;
    const int64_t a1 = pr[0], pe = pr[1];
    if (a1 <= 0 || pe == 0 || a1 > (int64_t)L) return false;
    if (pe < a1 || pe - a1 <= 2 || (int64_t)L < a1 + c.median_shift_window) return false;        // "early" candidate
    const int64_t nominal = pe - a1;
    if (nominal <= c.pa_var_window + 2 || nominal <= c.pa_mean_window + 2) return false;
    const int m = (int)(min(pe, (int64_t)L) - a1);
    const int cntv = m - c.pa_var_window + 1, cntm = m - c.pa_mean_window + 1;
    if (cntv < 1 || cntm < 1 || (int64_t)cntv + cntm > 2 * stride) return false;
    const int C = (cntv + FP_THREADS - 1) / FP_THREADS;
    if ((m + C - 1) / C > MS_CHUNKS) return false;
    return true;
}

__device__ __forceinline__ int val_var_code(double var) {     // float32 bit pattern >> 12: 2048 codes per octave, monotone
    float v32 = (float)var;
    if (!(v32 > 0.f)) v32 = 0.f;
    return (int)(__float_as_uint(v32) >> 12);
}
__device__ __forceinline__ int val_mean_code(double mean, uint32_t kb) {   // distance of the float32 order key from a lower bound's
    const uint32_t k = f32_key((float)mean);
    return k > kb ? (int)min(k - kb, 0x7ffffff0u) : 0;
}
__device__ __forceinline__ int val_bin8(int code, int flo, int fsh) { return code < flo ? 0 : min(255, 1 + ((code - flo) >> fsh)); }

__device__ __noinline__ void val_fast(const ValArgs& a, const ValCfg& c, const float* vsig, unsigned char* bins8, int L, int64_t fl,
                                      const int64_t* pr, ValSeg& seg, MsState& ms, ValCode8& c8, ValBand& bd, ValBand& bdm, FpScratch& s,
                                      int* sh_i, int* sh_pores, double* sh_d, double* sh_v, float* scratch, ValSel& vs_old, ValFastOut& fo) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a1 = (int)pr[0], e = a1, hi = a1;
    const int64_t pe = pr[1];
    const int pend = (int)min(pe, (int64_t)L);
    const int m = pend - e;
    const int wv = c.pa_var_window, wm = c.pa_mean_window;
    const int cntv = m - wv + 1, cntm = m - wm + 1;
    unsigned char* bv = bins8;             // [cntv] bins of the moving variance
    unsigned char* bm = bins8 + cntv;      // [cntm] bins of the moving mean
    const bool all_v = cntv <= VAL_BAND_CAP, all_m = cntm <= VAL_BAND_CAP;     // few windows: all of them are "the band"

    // ---- phase 1a: chunk sums of the centred poly(A) samples (seeds of the sliding sums of the moving variance), open
    // pores, minimum / maximum of the three ranges of the row (bounds of the bin maps)
    const int C = (cntv + FP_THREADS - 1) / FP_THREADS;     // windows per thread
    const int nch = (m + C - 1) / C;
    const int Cm = (cntm + FP_THREADS - 1) / FP_THREADS;
    double* S = reinterpret_cast<double*>(&ms.cand[0][0]);
    double* Q = S + MS_CHUNKS;
    const double cref = (double)vsig[e];
    for (int ch = tid; ch < nch; ch += FP_THREADS) {
        const int i1 = min((ch + 1) * C, m);
        double s1 = 0.0, s2 = 0.0;
        for (int i = ch * C; i < i1; i++) {
            const double d = (double)vsig[e + i] - cref;
            s1 += d;
            s2 += d * d;
        }
        S[ch] = s1;
        Q[ch] = s2;
    }
    {
        const int lo3[3] = {0, a1, pend}, hi3[3] = {a1, pend, L};
#pragma unroll
        for (int r = 0; r < 3; r++) {
            float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
            for (int i = lo3[r] + tid; i < hi3[r]; i += FP_THREADS) {
                const float x = vsig[i];
                mn = fminf(mn, x);
                mx = fmaxf(mx, x);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if (lane == 0) {
                atomicMin(&seg.mn[r], f32_key(mn));
                atomicMax(&seg.mx[r], f32_key(mx));
            }
        }
    }
    if (c.detect_open_pores) {
        // one pass: position i is kept iff it is an open-pore sample with none within min_obs_diff - 1 samples before it and
        // it is not the first one (find_open_pores starts at the second position); the first one is removed afterwards
        const float lo = c.open_pore_min;
        const int D = c.open_pore_min_obs_diff;
        for (int i = tid; i < hi; i += FP_THREADS) {
            if (vsig[i] >= lo) {
                atomicAdd(&sh_i[0], 1);
                atomicMin(&sh_i[1], i);
                atomicMax(&sh_i[2], i);
                bool near = false;
                for (int q = max(0, i - D + 1); q < i; q++) near |= (vsig[q] >= lo);
                if (!near) {
                    atomicMax(&sh_i[3], i);
                    const int slot = atomicAdd(&sh_i[4], 1);
                    if (slot < VAL_PORES_LD - 1) sh_pores[1 + slot] = i;
                }
            }
        }
    }
    __syncthreads();

    const float mnA = f32_unkey(seg.mn[0]), mxA = f32_unkey(seg.mx[0]);
    const float mnP = f32_unkey(seg.mn[1]), mxP = f32_unkey(seg.mx[1]);
    const float mnR = (pend < L) ? fminf(mnP, f32_unkey(seg.mn[2])) : mnP, mxR = (pend < L) ? fmaxf(mxP, f32_unkey(seg.mx[2])) : mxP;
    const uint32_t kbm = f32_key(mnP);                      // every mean of poly(A) samples is >= their minimum
    // open pores -> adapter start
    int64_t a0 = 0;
    int n_pores = 0, pores_cnt = -1, pores_single = 0;
    bool pore_fail = false;
    if (c.detect_open_pores) {
        const int cnt = sh_i[0], first = sh_i[1], last = sh_i[2], kept = max(0, sh_i[4] - 1);   // sh_i[4] counts the first one too
        if (cnt > 1) {
            n_pores = kept > 0 ? kept : 1;
            a0 = kept > 0 ? sh_i[3] : last;
        } else if (cnt == 1) {
            n_pores = 1;
            a0 = first;
        }
        pore_fail = cnt > 0 && a1 - a0 < c.min_obs_adapter;
        pores_cnt = cnt > 1 ? n_pores : cnt;
        pores_single = (cnt > 1 && kept > 0) ? -2 : (int)a0;    // -2: the list in sh_pores, minus its smallest entry
    }
    const int b0 = (int)min(a0, (int64_t)hi);
    const int nseg = hi - b0;
    const bool have_means = c.real_signal_check && nseg >= 2 * c.mean_window;
    int* smp = bdm.idx;                    // 2 x 64 sample codes (the mean's band list is free until round B)

    // the sums of a thread's first variance window from the chunk sums
    auto var_seed = [&](int p0, double& s1, double& s2) {
        const int k = wv / C;
        s1 = 0.0;
        s2 = 0.0;
        for (int q = 0; q < k; q++) {
            s1 += S[p0 / C + q];
            s2 += Q[p0 / C + q];
        }
        for (int i = p0 + k * C; i < p0 + wv; i++) {
            const double d = (double)vsig[e + i] - cref;
            s1 += d;
            s2 += d * d;
        }
    };
    const double inv_wv = 1.0 / (double)wv, inv_wm = 1.0 / (double)wm;

    FP_T(s, 1);   // phase 1a
    // ---- phase 1b: samples for the focus (every eighth thread's first window), real_range_check's two means
    if ((tid & 7) == 0) {
        int cv = -1, cm = -1;
        if (tid * C < cntv) {
            double s1, s2;
            var_seed(tid * C, s1, s2);
            const double mu = s1 * inv_wv;
            cv = val_var_code(s2 * inv_wv - mu * mu);
        }
        if (tid * Cm < cntm) {
            double s1 = 0.0;
            for (int i = 0; i < wm; i++) s1 += (double)vsig[e + tid * Cm + i] - cref;
            cm = val_mean_code(cref + s1 * inv_wm, kbm);
        }
        smp[tid >> 3] = cv;
        smp[64 + (tid >> 3)] = cm;
    }
    if (have_means && (warp == FP_WARPS - 1 || warp == FP_WARPS - 2)) {     // float32, numpy's pairwise order: one warp each
        const bool first_one = warp == FP_WARPS - 1;
        const int base_i = first_one ? b0 : hi - c.mean_window;
        const float sm = (c.mean_window <= 512) ? warp_np_sum_f32(base_i, c.mean_window, [&](int i) { return vsig[i]; })
                                                : np_pairwise<float>(base_i, c.mean_window, [&](int i) { return vsig[i]; });
        if (lane == 0) sh_d[first_one ? 0 : 1] = (double)__fdiv_rn(sm, (float)c.mean_window);
    }
    __syncthreads();

    FP_T(s, 2);   // phase 1b
    // ---- focus of the two code histograms: the 25 % / 75 % quantiles of the samples (the middle ranks lie 4 sigma inside;
    // a miss only costs the exact evaluation of every window).  Warps 0-1: variance, warps 2-3: mean.  Histograms cleared.
    uint32_t* hz = &ms.hist[0][0];
    for (int i = tid; i < MS_MAXJ * MS_BINS / 2; i += FP_THREADS) hz[i] = 0;
    if (tid < MS_MAXS) ms.slot_n[tid] = 0;
    if (tid < 256) (&c8.hist[0][0])[tid] = 0;
    if (tid < 128) {
        const int* sp = smp + (tid & 64);
        const int t = tid & 63, x = sp[t], which = tid >> 6;
        int ns = 0, rank = 0;
#pragma unroll 8
        for (int u = 0; u < 64; u++) {
            const int y = sp[u];
            ns += (y >= 0);
            rank += (y >= 0) && ((y < x) || (y == x && u < t));
        }
        if (x >= 0) {
            if (rank == ns / 4) sh_i[8 + which * 2] = x;
            if (rank == (ns * 3) / 4) sh_i[9 + which * 2] = x;
        }
    }
    __syncthreads();

    FP_T(s, 3);   // focus
    // ---- phase 1c: every window's approximation -> bin -> histogram; the jobs of round A are set up next to it
    const int flo_v = sh_i[8], flo_m = sh_i[10];
    int fsh_v = 0, fsh_m = 2;               // the mean's codes are 1 ulp apart: a bin of >= 4 ulp keeps the +-1-bin band safe
    while ((((long long)sh_i[9] - flo_v) >> fsh_v) > 253) fsh_v++;
    while ((((long long)sh_i[11] - flo_m) >> fsh_m) > 253) fsh_m++;
    // rank layout of round A (every thread derives it; the lane 0 of warp jj writes job jj)
    const int nn = min(c.max_obs_local_range, nseg);
    const bool lr_on = c.real_signal_check && nn >= 1;
    const bool lr_merged = lr_on && nn == hi;          // real_range_check's percentiles over the whole adapter: ranks of the same job
    const int up = min(e + c.median_shift_window, L), dn = max(e - c.median_shift_window, 0);
    const int ms_e = min(a1, L);
    const int ms_up = (int)min(min((int64_t)a1 + c.med_shift_window, fl), (int64_t)L), ms_dn = (int)max((int64_t)a1 - c.med_shift_window, (int64_t)0);
    const int ms_n1 = max(0, ms_up - ms_e), ms_n2 = max(0, ms_e - min(ms_dn, ms_e));
    int jn = 0, qn = 0;
    const int ja = jn++, qa = qn;
    qn += 2;
    int jl = -1, ql = -1;
    if (lr_on) {
        ql = qn;
        qn += 4;
        if (!lr_merged) jl = jn++;
    }
    const int jp = jn++, qp = qn;
    qn += 6;
    const int js1 = jn++, qs1 = qn;
    qn += 2;
    const int js2 = jn++, qs2 = qn;
    qn += 2;
    int jm1 = -1, qm1 = -1, jm2 = -1, qm2 = -1;
    if (c.detect_med_shift && ms_n1 >= 1) {
        jm1 = jn++;
        qm1 = qn;
        qn += 2;
    }
    if (c.detect_med_shift && ms_n2 >= 1) {
        jm2 = jn++;
        qm2 = qn;
        qn += 2;
    }
    double g85l = 0, g15l = 0, g85p = 0, g15p = 0;
    uint32_t rl[4] = {0, 0, 0, 0}, rp[6];
    if (lr_on) {
        pct_ranks(nn, 85.0, &rl[0], &rl[1], &g85l);
        pct_ranks(nn, 15.0, &rl[2], &rl[3], &g15l);
    }
    pct_ranks(m, 85.0, &rp[0], &rp[1], &g85p);
    pct_ranks(m, 15.0, &rp[2], &rp[3], &g15p);
    rp[4] = (uint32_t)((m - 1) / 2);
    rp[5] = (uint32_t)(m / 2);
    if (lane == 0 && warp < jn) {
        auto median_job = [&](int jj, int q0, const float* p, int n, float lo, float hi_) {
            const uint32_t rk[2] = {(uint32_t)((n - 1) / 2), (uint32_t)(n / 2)};
            ms_set_job(ms, jj, p, n, 0, 0.f, lo, hi_, q0, 2, rk);
        };
        if (warp == ja) {
            if (lr_merged) {
                const uint32_t rk[6] = {(uint32_t)((hi - 1) / 2), (uint32_t)(hi / 2), rl[0], rl[1], rl[2], rl[3]};
                ms_set_job(ms, ja, vsig, hi, 0, 0.f, mnA, mxA, qa, 6, rk);
            } else {
                median_job(ja, qa, vsig, hi, mnA, mxA);
            }
        }
        if (warp == jl) ms_set_job(ms, jl, vsig + (hi - nn), nn, 0, 0.f, mnA, mxA, ql, 4, rl);
        if (warp == jp) ms_set_job(ms, jp, vsig + e, m, 0, 0.f, mnP, mxP, qp, 6, rp);
        if (warp == js1) median_job(js1, qs1, vsig + e, up - e, mnR, mxR);
        if (warp == js2) median_job(js2, qs2, vsig + dn, e - dn, mnA, mxA);
        if (warp == jm1) median_job(jm1, qm1, vsig + ms_e, ms_n1, mnR, mxR);
        if (warp == jm2) median_job(jm2, qm2, vsig + min(ms_dn, ms_e), ms_n2, mnA, mxA);
        if (warp == 0) {
            ms.nj = jn;
            ms.nq = qn;
            ms.n_slots = 0;
            ms.crowded = 0;
            bd.n = 0;
            bd.below = 0;
            bdm.n = 0;
            bdm.below = 0;
        }
    }
    {
        if (!all_v && tid * C < cntv) {
            const int p0 = tid * C, pe_ = min(p0 + C, cntv);
            double s1, s2;
            var_seed(p0, s1, s2);
            for (int p = p0;; p++) {
                const double mu = s1 * inv_wv;
                const int b = val_bin8(val_var_code(s2 * inv_wv - mu * mu), flo_v, fsh_v);
                bv[p] = (unsigned char)b;
                atomicAdd(&c8.hist[0][b >> 1], (b & 1) ? 0x10000u : 1u);
                if (p + 1 >= pe_) break;
                const double dn_ = (double)vsig[e + p + wv] - cref, dl = (double)vsig[e + p] - cref;
                s1 += dn_ - dl;
                s2 += dn_ * dn_ - dl * dl;
            }
        }
        if (!all_m && tid * Cm < cntm) {
            const int p0 = tid * Cm, pe_ = min(p0 + Cm, cntm);
            double s1 = 0.0;
            for (int i = 0; i < wm; i++) s1 += (double)vsig[e + p0 + i] - cref;
            for (int p = p0;; p++) {
                const int b = val_bin8(val_mean_code(cref + s1 * inv_wm, kbm), flo_m, fsh_m);
                bm[p] = (unsigned char)b;
                atomicAdd(&c8.hist[1][b >> 1], (b & 1) ? 0x10000u : 1u);
                if (p + 1 >= pe_) break;
                s1 += (double)vsig[e + p + wm] - (double)vsig[e + p];
            }
        }
    }
    __syncthreads();

    FP_T(s, 4);   // phase 1c
    // ---- round A: every selection on samples that does not need another one's result; two more warps find the bins of
    // the middle ranks of the two code histograms
    const uint32_t rkv0 = (uint32_t)((cntv - 1) / 2), rkv1 = (uint32_t)(cntv / 2);
    const uint32_t rkm0 = (uint32_t)((cntm - 1) / 2), rkm1 = (uint32_t)(cntm / 2);
    ms_round(ms, s, [] {},
        [&] {
            const int which = warp - (FP_WARPS - 2);      // warps 14, 15
            if (which < 0) return;
            const uint32_t* h = c8.hist[which] + lane * 4;     // 8 bins per lane
            uint32_t cnt8[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                cnt8[2 * t] = h[t] & 0xffffu;
                cnt8[2 * t + 1] = h[t] >> 16;
                sum += cnt8[2 * t] + cnt8[2 * t + 1];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            uint32_t run = inc - sum;
            const uint32_t k0 = which ? rkm0 : rkv0, k1 = which ? rkm1 : rkv1;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                if (k0 >= run && k0 < run + cnt8[t]) c8.b_lo[which] = lane * 8 + t;
                if (k1 >= run && k1 < run + cnt8[t]) c8.b_hi[which] = lane * 8 + t;
                run += cnt8[t];
            }
        },
        [] {});
    auto qf = [&](int q) { return f32_unkey(ms.q_key[q]); };
    const float med = median_of(hi, qf(qa), qf(qa + 1));
    const double lr = lr_on ? __dsub_rn(pct_lerp(qf(ql), qf(ql + 1), g85l), pct_lerp(qf(ql + 2), qf(ql + 3), g15l)) : 0.0;
    const double r_lr = __dsub_rn(pct_lerp(qf(qp), qf(qp + 1), g85p), pct_lerp(qf(qp + 2), qf(qp + 3), g15p));
    const double r_med = (double)median_of(m, qf(qp + 4), qf(qp + 5));
    const double r_shift = (double)__fsub_rn(median_of(up - e, qf(qs1), qf(qs1 + 1)), median_of(e - dn, qf(qs2), qf(qs2 + 1)));
    const float qnan32 = __int_as_float(0x7fc00000);
    const double msv = c.detect_med_shift ? (double)__fsub_rn(qm1 >= 0 ? median_of(ms_n1, qf(qm1), qf(qm1 + 1)) : qnan32,
                                                              qm2 >= 0 ? median_of(ms_n2, qf(qm2), qf(qm2 + 1)) : qnan32) : 0.0;
    const double mean_start = sh_d[0], mean_end = sh_d[1];
    // the bands of the two code jobs: bins of the middle ranks +- 1 (bins 0 / 255 = outside the focus: taken whole, which
    // overflows the band and sends the read to the exact evaluation of every window)
    const int bv_lo = all_v ? 0 : max(0, c8.b_lo[0] - 1), bv_hi = all_v ? 255 : min(255, c8.b_hi[0] + 1);
    const int bm_lo = all_m ? 0 : max(0, c8.b_lo[1] - 1), bm_hi = all_m ? 255 : min(255, c8.b_hi[1] + 1);
    __syncthreads();      // everyone has read the results of round A

    FP_T(s, 9);   // results of round A
    // ---- round B: the adapter MAD (needs the median) and the exact evaluation of the windows around the middle
    for (int i = tid; i < MS_BINS / 2; i += FP_THREADS) hz[i] = 0;
    if (tid < MS_MAXS) ms.slot_n[tid] = 0;
    if (tid == 0) {
        ms.nj = 1;
        ms.nq = 2;
        ms.n_slots = 0;
        ms.crowded = 0;
        const float dmax = fmaxf(__fsub_rn(mxA, med), __fsub_rn(med, mnA));
        const uint32_t rk[2] = {(uint32_t)((hi - 1) / 2), (uint32_t)(hi / 2)};
        ms_set_job(ms, 0, vsig, hi, 1, med, 0.f, dmax, 0, 2, rk);
    }
    __syncthreads();
    auto band_pass = [&](const unsigned char* b8, int cnt, bool all, int lo8, int hi8, ValBand& B) {
        if (all) {
            for (int p = tid; p < cnt; p += FP_THREADS) B.idx[p] = p;
            if (tid == 0) B.n = cnt;
            return;
        }
        int below = 0;
        for (int p = tid; p < cnt; p += FP_THREADS) {
            const int b = (int)b8[p];
            if (b < lo8) below++;
            else if (b <= hi8) {
                const int pos = atomicAdd(&B.n, 1);
                if (pos < VAL_BAND_CAP) B.idx[pos] = p;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below) atomicAdd(&B.below, below);
    };
    auto band_rank = [&](ValBand& B, uint32_t k0, uint32_t k1, int first, int step) {
        const int nb = B.n, r0 = (int)k0 - B.below, r1 = (int)k1 - B.below;
        if (nb > VAL_BAND_CAP || r0 < 0 || r1 >= nb) return;
        for (int j = first; j < nb; j += step) {
            const float x = B.val[j];
            int rank = 0;
            for (int u = 0; u < nb; u++) {
                const float y = B.val[u];
                rank += (y < x) || (y == x && u < j);
            }
            if (rank == r0) B.out[0] = x;
            if (rank == r1) B.out[1] = x;
        }
    };
    ms_round(ms, s,
        [&] {   // next to the histogram pass: windows certainly below the middle, windows inside the band
            band_pass(bv, cntv, all_v, bv_lo, bv_hi, bd);
            band_pass(bm, cntm, all_m, bm_lo, bm_hi, bdm);
        },
        [&] {   // next to the scan (warp 0): exact statistic of the bands' windows
            if (tid < 32) return;
            const int nb = bd.n, nbm = bdm.n;
            if (nb <= VAL_BAND_CAP)
                for (int j = tid - 32; j < nb; j += FP_THREADS - 32) bd.val[j] = val_window_exact<true>(vsig, e + bd.idx[j], wv);
            if (nbm <= VAL_BAND_CAP)
                for (int j = FP_THREADS - 1 - tid; j < nbm; j += FP_THREADS - 32) bdm.val[j] = val_window_exact<false>(vsig, e + bdm.idx[j], wm);
        },
        [&] {   // next to the gather: exact ranking inside the bands
            band_rank(bd, rkv0, rkv1, tid, FP_THREADS);
            band_rank(bdm, rkm0, rkm1, FP_THREADS - 1 - tid, FP_THREADS);
        }, 10);
    bool band_ok, mband_ok;
    {
        const int nb = bd.n, r0 = (int)rkv0 - bd.below, r1 = (int)rkv1 - bd.below;
        band_ok = nb <= VAL_BAND_CAP && r0 >= 0 && r1 < nb;
        const int nbm = bdm.n, m0 = (int)rkm0 - bdm.below, m1 = (int)rkm1 - bdm.below;
        mband_ok = nbm <= VAL_BAND_CAP && m0 >= 0 && m1 < nbm;
    }
    const float mad = median_of(hi, qf(0), qf(1));
    double r_var = band_ok ? (double)median_of(cntv, bd.out[0], bd.out[1]) : 0.0;
    double r_mean = mband_ok ? (double)median_of(cntm, bdm.out[0], bdm.out[1]) : 0.0;
    // uniform; a crowded band (a constant stretch, a middle rank outside the focus): every window exactly, as the sequential path does
    if (!band_ok || !mband_ok) {
        __syncthreads();
        if (!band_ok) r_var = (double)val_window_median<true>(vsig, e, m, wv, scratch, bd, vs_old, s);
        if (!mband_ok) r_mean = (double)val_window_median<false>(vsig, e, m, wm, scratch, bd, vs_old, s);
    }

    FP_T(s, 13);  // round B incl. its set-up
    // ---- the verdict, in the reference's order (combined.py:452-629)
    if (tid == 0) {
        int code = VAL_OK, checks = 0, resume = 0;
        long long pe_best = pe;
        sh_v[0] = (double)med;
        sh_v[1] = (double)mad;
        if (mad != 0.f && !val_in_range((double)mad, c.mad_lo, c.mad_hi)) code = VAL_ADAPTER_MAD;
        const bool pores_ran = code == VAL_OK && c.detect_open_pores;
        if (pores_ran && pore_fail) code = VAL_OPEN_PORE;
        if (code == VAL_OK && c.real_signal_check) {
            bool ok = false;
            if (have_means) {
                sh_v[2] = mean_start;
                sh_v[3] = mean_end;
                if (val_in_range(mean_start, c.mean_start_lo, c.mean_start_hi) && val_in_range(mean_end, c.mean_end_lo, c.mean_end_hi)) {
                    sh_v[4] = lr;
                    ok = val_in_range(lr, c.local_range_lo, c.local_range_hi);
                }
            }
            if (!ok) code = VAL_REAL_RANGE;
        }
        if (code == VAL_OK) {
            double mlo = c.mean_lo, mhi = c.mean_hi;
            if (c.mean_from_scale) {
                mlo = __dmul_rn(c.scale_lo, (double)med);
                mhi = __dmul_rn(c.scale_hi, (double)med);
            }
            const int bits = (val_in_range(r_mean, mlo, mhi) ? 1 : 0) | (val_in_range(r_var, c.var_lo, c.var_hi) ? 2 : 0) |
                             (val_in_range(r_med, c.pmed_lo, c.pmed_hi) ? 4 : 0) | (val_in_range(r_lr, c.plr_lo, c.plr_hi) ? 8 : 0) |
                             (val_in_range(r_shift, c.shift_lo, c.shift_hi) ? 16 : 0);
            sh_v[5] = r_mean;
            sh_v[6] = r_var;
            sh_v[7] = r_med;
            sh_v[8] = r_lr;
            sh_v[9] = r_shift;
            if (bits != 31) {
                if (r_mean == 0.0) {
                    code = VAL_MVS_NO_SIGNAL;
                    checks = 0;
                } else {
                    code = VAL_MVS_CHECKS;
                    checks = bits;
                }
                resume = !a.verdict_only;     // the reference goes on to the next candidates (their report overwrites this one)
            }
        }
        if (code == VAL_OK && c.detect_med_shift) {
            sh_v[10] = msv;
            if (!val_in_range(msv, c.ms_lo, c.ms_hi)) code = VAL_MED_SHIFT;
        }
        fo.code = code;
        fo.checks = checks;
        fo.resume = resume;
        fo.pe_best = pe_best;
        fo.a0 = pores_ran ? a0 : 0;
        fo.n_pores = pores_ran ? n_pores : 0;
        fo.pores_cnt = pores_ran ? pores_cnt : -1;
        fo.pores_single = pores_single;
        fo.med = med;
        fo.shift = r_shift;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FP_THREADS, WDX_VAL_MIN_CTAS) validate_kernel(const ValArgs a, const ValCfg c) {
    extern __shared__ float vsig[];     // the row, then [2 * stride] 8-bit bins of the moving variance / mean windows
    __shared__ FpScratch s;
    // the concurrent selections of the common read and the selection / pairwise-sum state of the sequential path are
    // never live at the same time
    __shared__ union ValShared {
        MsState ms;
        struct {
            ValSel vs;
            ValTree tree;
        } seq;
        __device__ ValShared() {}
    } shu;
    ValSel& vs = shu.seq.vs;
    ValTree& tree = shu.seq.tree;
    __shared__ ValBand band, band2;
    __shared__ ValSeg seg;
    __shared__ ValCode8 c8;
    __shared__ ValFastOut fo;
    __shared__ int sh_i[12];
    __shared__ int sh_pores[VAL_PORES_LD];
    __shared__ double sh_d[4];
    __shared__ double sh_v[VAL_NVALS];
    const int tid = threadIdx.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    float* scratch = a.scratch + (size_t)blockIdx.x * a.stride;
    unsigned char* bins8 = reinterpret_cast<unsigned char*>(vsig + ((a.stride + 3) & ~(int64_t)3));   // [2 * stride]

    __shared__ unsigned long long sh_next;
    FP_T_BEGIN(s);
    for (;;) {
        FP_T(s, 15);   // rest of the previous read (sequential path, reports)
        FP_T_END(s);
        __syncthreads();
        if (tid == 0) {
            unsigned long long nx = atomicAdd(a.next, 1ULL);
            if (a.list) nx = (nx < (unsigned long long)*a.list_count) ? (unsigned long long)a.list[nx] : (unsigned long long)a.n;
            sh_next = nx;
            sh_i[0] = 0;           // open-pore samples
            sh_i[1] = 0x7fffffff;  // first
            sh_i[2] = -1;          // last
            sh_i[3] = -1;          // last kept
            sh_i[4] = 0;           // kept
            sh_i[5] = 65535;       // smallest / largest code of the moving variance
            sh_i[6] = 0;
            sh_i[8] = sh_i[10] = 0;                 // focus of the code histograms (30 % / 70 % sample quantiles)
            sh_i[9] = sh_i[11] = 0x7ffffff0;
        }
        if (tid < 3) {
            seg.mn[tid] = 0xffffffffu;
            seg.mx[tid] = 0u;
        }
        __syncthreads();
        const int64_t r = (int64_t)sh_next;
        if (r >= a.n) break;
        FP_T_BEGIN(s);
        const float* row = a.signals + (size_t)r * a.stride;
        const int64_t fl = a.full_len[r];
        const int L = (int)max((int64_t)0, min(fl, a.stride));
        int has_nan = 0;
        if (tid < VAL_NVALS) sh_v[tid] = qnan;
        const int64_t* pr = a.preds + (size_t)r * a.ld;
        const bool fast = val_fast_eligible(c, L, a.stride, pr, a.ld);
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
        has_nan = __syncthreads_or(has_nan);

        const int64_t a1 = pr[0];
        int64_t a0 = 0;
        int64_t pe_best = a.ld > 1 ? pr[1] : 0;
        int code = VAL_OK, checks = 0, n_pores = 0;
        // reported statistics live in shared memory (thread 0 writes them as they are found); the fill above is
        // ordered before those writes by the barrier of the NaN vote

        if (has_nan) code = VAL_HAS_NAN;
        const int hi = (int)max((int64_t)0, min(a1, (int64_t)L));   // sig[a0:a1] ends here
        float med = 0.f, mad = 0.f;
        int pores_cnt = -1, pores_single = 0;   // open_pores as the reference reports it: None / the kept positions
        FP_T(s, 0);   // row load + range scan
        const bool use_fast = fast && code == VAL_OK;
        int resume = 0;                        // poly(A) candidates from here on go through the sequential loop
        if (use_fast) {
            val_fast(a, c, vsig, bins8, L, fl, pr, seg, shu.ms, c8, band, band2, s, sh_i, sh_pores, sh_d, sh_v, scratch, vs, fo);
            FP_T(s, 14);  // verdict
            code = fo.code;
            checks = fo.checks;
            n_pores = fo.n_pores;
            pores_cnt = fo.pores_cnt;
            pores_single = fo.pores_single;
            a0 = fo.a0;
            pe_best = fo.pe_best;
            med = fo.med;
            resume = fo.resume;
        }
        if (!use_fast && code == VAL_OK) {
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
        if (!use_fast && code == VAL_OK && mad != 0.f && !val_in_range((double)mad, c.mad_lo, c.mad_hi)) code = VAL_ADAPTER_MAD;

        if (!use_fast && code == VAL_OK && c.detect_open_pores) {
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

        if (!use_fast && code == VAL_OK && c.real_signal_check) {
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

        if ((use_fast ? resume != 0 : code == VAL_OK) && c.mvs_detect_check) {
            if (pe_best == 0) code = VAL_NO_POLYA;
            else {
                double mlo = c.mean_lo, mhi = c.mean_hi;
                if (c.mean_from_scale) {
                    mlo = __dmul_rn(c.scale_lo, (double)med);
                    mhi = __dmul_rn(c.scale_hi, (double)med);
                }
                const int e = (int)a1;   // a1 <= L here or the size test below fails first
                bool have_shift = use_fast;
                double shift_cached = use_fast ? fo.shift : 0.0;
                for (int j = use_fast ? 2 : 1; j < a.ld; j++) {
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

        if (!use_fast && code == VAL_OK && c.detect_med_shift) {
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
                    const int extra = pores_single == -2 ? 1 : 0;    // single-pass scan: the list still holds the first open-pore sample
                    const int m = min(pores_cnt + extra, VAL_PORES_LD - 1);
                    for (int i = 1; i < m; i++) {
                        const int v = sh_pores[1 + i];
                        int j = i - 1;
                        for (; j >= 0 && sh_pores[1 + j] > v; j--) sh_pores[2 + j] = sh_pores[1 + j];
                        sh_pores[2 + j] = v;
                    }
                    for (int i = extra; i < m; i++) po[1 + i - extra] = sh_pores[1 + i];
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
