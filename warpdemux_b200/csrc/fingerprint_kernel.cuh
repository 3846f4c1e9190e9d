// fingerprint_kernel.cuh — barcode fingerprint extraction, one CTA per read (sm_100a).
//
// Restates, for the configuration every shipped DTW-SVM model uses
// (rna004_130bps@v1.0.toml), the reference's per-read Python/Cython chain
//   detect_results_to_fpt                       warpdemux/sig_proc.py:394-605
//     extract_adapter                           sig_proc.py:382-391
//     nanmedian / MAD winsorisation (float32)   sig_proc.py:421-431
//     segment_signal                            sig_proc.py:201-254
//       c_windowed_t_test (float64)             segmentation/_c_segmentation.pyx:124-161
//       scipy find_peaks(distance=...)          sig_proc.py:183  (local maxima + greedy distance suppression)
//       top num_events peaks by score, sorted   sig_proc.py:188-198
//       c_new_means (float64)                   _c_segmentation.pyx:41-53
//     normalize(..., "mean") (numpy pairwise)   sig_proc.py:99-111, 546-552
//     six adapter statistics                    sig_proc.py:562-567
//     keep the last barcode_num_events          sig_proc.py:569-594
// as ONE kernel: the adapter slice is read from HBM exactly once (coalesced)
// into shared memory and everything else happens on-chip; 200 B of fingerprint
// (+ optional dwell times / statistics) go back.  HBM-bound by construction:
// algorithmic bytes per read = 4 * n_adapter in + 8 * barcode_num_events out.
//
// Exactness: every float operation is issued in the reference's order and
// precision (float32 for the winsorisation, float64 for scores and means; this
// TU is compiled with -fmad=false), so change points — integer work — are
// identical and the float64 outputs are bit-identical to the CPU chain.
// The sequential pieces are replaced by order-independent equivalents:
//   * medians: exact radix selection on order-preserving integer keys;
//   * find_peaks distance suppression (highest first, sequential): fixed point of
//     "a peak stays iff no STAYING higher peak lies within distance" — unique, and
//     equal to the greedy result, for any total priority order;
//   * top-k by score: radix selection of the k-th largest score.
// Ties between EQUAL scores are broken towards the higher index (what a stable
// argsort would do); numpy's default argsort leaves them unspecified.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "block_select.cuh"

namespace wdx {

enum { FP_OK = 0, FP_FAIL_SEGMENTATION = 1, FP_FAIL_DETECT = 2, FP_FAIL_NORMALIZE = 3, FP_FAIL_TOO_LONG = 4, FP_FAIL_CONSENSUS = 5,
       FP_PENDING = 100 /* internal: between the kernels of the three-launch consensus form */ };
constexpr int FP_MAX_QUERY = 128;  // longest consensus query (4 rows per lane of one warp)

struct FpConfig {
    int padding;             // sig_extract.padding
    float outlier_thresh;    // core.sig_norm_outlier_thresh (numpy 2: weak Python float * float32 -> float32)
    double outlier_thresh_d; // the same value as given
    int numpy1_promotion;    // winsorisation bounds as numpy < 2 forms them (the reference pins numpy 1.26.4): med -+ thresh * mad
                             // in float64 (np.float32 scalar * Python float -> float64), cast to float32 ONCE by np.clip
    int min_obs_per_base;    // segmentation.min_obs_per_base
    int running_stat_width;  // segmentation.running_stat_width
    int num_events;          // segmentation.num_events
    int barcode_num_events;  // segmentation.barcode_num_events (consensus mode: barcode_num_events[1], events kept)
    // consensus-guided barcode refinement (segmentation.consensus_refinement, sig_proc.py:257-378, 451-521)
    int cons_len;            // length of the consensus query; 0 = refinement off
    int cons_seg_events;     // barcode_num_events[0]: change points of the second segmentation
    double cons_pen2;        // consensus_subseq_match_penalty squared
    int cons_psi_q, cons_psi_s;              // consensus_subseq_match_psi[0], [2] (start relaxation: query, series)
    int cons_ub_start, cons_lb_end, cons_ub_end;
};

struct FpArgs {
    const float* signals;       // [n][stride]
    float* signals_mut;         // == signals when the winsorised slice is to be written back, else nullptr
    int64_t stride;
    const int32_t* sig_len;     // [n] valid samples per row, or nullptr (trailing NaN padding is detected)
    const int64_t* adapter_start;
    const int64_t* adapter_end;
    const uint8_t* detect_ok;   // [n] DetectResults.success, or nullptr (all true)
    int64_t n;
    int cap;                    // samples of shared memory per CTA
    int retry_status;           // != 0: second pass with a larger `cap` — only reads whose status equals this are processed
    double* fpt;                // [n][barcode_num_events]
    int64_t* dwell;             // [n][barcode_num_events] or nullptr
    double* stats;              // [n][6] or nullptr
    int32_t* status;            // [n]
    const double* cons_query;   // [cons_len] consensus query (device), consensus mode only
    int32_t* cons;              // [n][3] seg_cons_query_start, seg_cons_query_end, sig_barcode_start, or nullptr
    unsigned char* park;        // three-launch consensus form: fp_park_bytes(cap) of state per read between the kernels
};

// State of a read between the kernels of the three-launch consensus form (global memory): the winsorised slice, the
// t-test scores, the first segmentation's change points, the normalised event series, a few scalars, the match.
struct FpParkHead {
    int n, nc, w, n_seg, exact_sums, match[2], pad;
    double ev_mean, ev_std;
};
__host__ __device__ inline size_t fp_park_bytes(int cap) {
    return sizeof(FpParkHead) + (size_t)cap * 12 + (size_t)(FP_MAX_EVENTS + 2) * 12 + 16;
}
__device__ __forceinline__ FpParkHead* park_head(unsigned char* p) { return reinterpret_cast<FpParkHead*>(p); }
__device__ __forceinline__ double* park_score(unsigned char* p) { return reinterpret_cast<double*>(p + sizeof(FpParkHead)); }
__device__ __forceinline__ double* park_series(unsigned char* p, int cap) { return park_score(p) + cap; }
__device__ __forceinline__ float* park_sig(unsigned char* p, int cap) { return reinterpret_cast<float*>(park_series(p, cap) + FP_MAX_EVENTS + 2); }
__device__ __forceinline__ int* park_cpts(unsigned char* p, int cap) { return reinterpret_cast<int*>(park_sig(p, cap) + cap); }

// Same result as block_median_f32, found with two light passes instead of five
// heavy ones: values are binned linearly over [vmin, vmax] (a monotone map, so
// every element of a lower bin is <= every element of a higher bin), the bin that
// holds rank (n-1)/2 is located by a scan of the histogram, its few members are
// gathered and ranked exactly on their order keys.  Falls back to the radix
// selection when the bin is crowded (degenerate signals).
//   cand: FP_MED_CAND uint32 of scratch in shared memory
//   hist: FP_MED_BINS uint32 ZEROED by the caller, *ncand zeroed too (both behind a barrier; two medians in a row
//   use two histograms, so neither starts with a clearing pass and its barriers)
template <typename VAL>
__device__ float block_median_f32_linear(int n, VAL val, float vmin, float vmax, uint32_t* hist, uint32_t* cand,
                                         uint32_t* ncand, FpScratch& s) {
    if (!(vmax > vmin)) return vmin;  // all values equal
    const float scale = __fdiv_rn((float)FP_MED_BINS, __fsub_rn(vmax, vmin));
    auto key_of = [&](int i) { return f32_key(val(i)); };
    if (!(scale < 1e30f)) return block_median_f32(n, key_of, s);
    auto bin_of = [&](float x) { return min(FP_MED_BINS - 1, (int)__fmul_rn(__fsub_rn(x, vmin), scale)); };
    const int tid = threadIdx.x;
    const uint32_t k_lo = (uint32_t)((n - 1) / 2);
    for (int i = tid; i < n; i += FP_THREADS) atomicAdd(&hist[bin_of(val(i))], 1u);
    __syncthreads();
    {   // thread t owns bins [t*B, (t+1)*B)
        constexpr int B = FP_MED_BINS / FP_THREADS;
        uint32_t c[B], sum = 0;
#pragma unroll
        for (int q = 0; q < B; q++) {
            c[q] = hist[tid * B + q];
            sum += c[q];
        }
        uint32_t total;
        uint32_t run = block_exscan(sum, s, &total);
        if (k_lo >= run && k_lo < run + sum) {  // exactly one thread
#pragma unroll
            for (int q = 0; q < B; q++) {
                if (k_lo >= run && k_lo < run + c[q]) {
                    s.sel_bin = tid * B + q;
                    s.sel_below = run;
                    s.sel_count = c[q];
                }
                run += c[q];
            }
        }
    }
    __syncthreads();
    const int sel_bin = s.sel_bin;
    const uint32_t below = s.sel_below, m = s.sel_count;
    if (m > (uint32_t)FP_MED_CAND) return block_median_f32(n, key_of, s);  // uniform decision
    for (int i = tid; i < n; i += FP_THREADS) {
        const float x = val(i);
        if (bin_of(x) == sel_bin) cand[atomicAdd(ncand, 1u)] = f32_key(x);
    }
    __syncthreads();
    const uint32_t r = k_lo - below;  // wanted rank inside the bin
    const bool even = (n & 1) == 0;
    for (uint32_t t = tid; t < m; t += FP_THREADS) {
        const uint32_t x = cand[t];
        uint32_t rank = 0;
        for (uint32_t u = 0; u < m; u++) {
            const uint32_t y = cand[u];
            rank += (y < x) || (y == x && u < t);
        }
        if (rank == r) s.key_lo = x;
        if (rank == r + 1) s.key_hi = x;
    }
    __syncthreads();
    const float v_lo = f32_unkey(s.key_lo);
    if (!even) return v_lo;
    float v_hi;
    if (r + 1 < m) {
        v_hi = f32_unkey(s.key_hi);
    } else {  // the upper middle element is the smallest member of the following bins
        if (tid == 0) s.hist[1] = 0xffffffffu;
        __syncthreads();
        uint32_t mn = 0xffffffffu;
        for (int i = tid; i < n; i += FP_THREADS) {
            const float x = val(i);
            if (bin_of(x) > sel_bin) mn = min(mn, f32_key(x));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((tid & 31) == 0) atomicMin(&s.hist[1], mn);
        __syncthreads();
        v_hi = f32_unkey(s.hist[1]);
    }
    return __fdiv_rn(__fadd_rn(v_lo, v_hi), 2.0f);
}

// x / w for a small positive integer w, correctly rounded, in three FP64 operations
// (Markstein: with r = RN(1/w) and q0 = RN(x*r), q0 + RN(x - w*q0)*r rounds to RN(x/w);
// checked against IEEE division on 4.8e8 operands per w in 1..12).  Bit-identical to
// the reference's `m1 /= running_stat_width`.
__device__ __forceinline__ double div_small_int(double x, double w, double r) {
    const double q0 = __dmul_rn(x, r);
    const double rem = __fma_rn(-w, q0, x);
    return __fma_rn(rem, r, q0);
}

// Mean and sum of squared deviations of the W samples at p, in the reference's order
// (_c_segmentation.pyx:133-149).  The statistics of window [pos+W, pos+2W) at position pos are
// bit for bit those of window [pos', pos'+W) at pos' = pos + W (same operands, same sequential
// order), so a thread that walks pos, pos+W, pos+2W, ... computes every window once.
// TREE: the sum is known to be exact (see `exact_sums` in the kernel), so it may be formed as a tree — a dependent chain
// of 4 additions instead of 12.
template <int W, bool TREE = false>
__device__ __forceinline__ void window_stat_fixed(const float* p, double wd, double wr, double& m, double& v) {
    double x[W];
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = (double)p[i];
    if (TREE && W == 12) {
        const double s01 = __dadd_rn(x[0], x[1]), s23 = __dadd_rn(x[2], x[3]), s45 = __dadd_rn(x[4], x[5]);
        const double s67 = __dadd_rn(x[6], x[7]), s89 = __dadd_rn(x[8], x[9]), sab = __dadd_rn(x[10], x[11]);
        m = __dadd_rn(__dadd_rn(__dadd_rn(s01, s23), __dadd_rn(s45, s67)), __dadd_rn(s89, sab));
        m = __dadd_rn(m, 0.0);   // an exact zero sum is +0.0 in the reference (it starts from +0.0); -0.0 + 0.0 = +0.0
    } else {
        m = 0.0;
#pragma unroll
        for (int i = 0; i < W; i++) m = __dadd_rn(m, x[i]);
    }
    m = (W <= 12) ? div_small_int(m, wd, wr) : __ddiv_rn(m, wd);   // the three-operation division is verified for w <= 12 only
    v = 0.0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const double pd = __dsub_rn(x[i], m);
        v = __dadd_rn(v, __dmul_rn(pd, pd));
    }
}

// One t-test score (_c_segmentation.pyx:151-156) from the statistics of its two windows.
__device__ __forceinline__ double ttest_combine(double m1, double var1, double m2, double var2) {
    const double vs = __dadd_rn(var1, var2);
    if (vs == 0.0) return 0.0;
    const double num = (m1 > m2) ? __dsub_rn(m1, m2) : __dsub_rn(m2, m1);
    return __ddiv_rn(num, __dsqrt_rn(vs));
}

// The t-test over residue-class chains (see the kernel): thread = (residue r = pos mod W, segment of the chain r, r + W, ...);
// the second window of one position is the first window of the next.  nblk = ceil(nc / W), seg_len = windows per thread.
template <int W, bool TREE>
__device__ __forceinline__ void ttest_residue_chains(const float* sig, double* score, int nc, int nblk, int seg_len) {
    const int tid = threadIdx.x;
    const double wd = (double)W, wr = 1.0 / (double)W;
    const int n_seg = (nblk + seg_len - 1) / seg_len;  // <= FP_THREADS / W: one item per thread
    if (tid < W * n_seg) {
        int pos = tid % W + W * seg_len * (tid / W);
        if (pos < nc) {
            double m1, v1;
            window_stat_fixed<W, TREE>(sig + pos, wd, wr, m1, v1);
            for (int j = 0; j < seg_len && pos < nc; j++, pos += W) {
                double m2, v2;
                window_stat_fixed<W, TREE>(sig + pos + W, wd, wr, m2, v2);
                score[pos] = ttest_combine(m1, v1, m2, v2);
                m1 = m2;
                v1 = v2;
            }
        }
    }
}

// numpy's pairwise summation for n <= 128 contiguous float64 (np.add.reduce):
// 8 strided accumulators, fixed combination tree, sequential tail.
template <typename F>
__device__ double np_pairwise_sum(int n, F at) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r = __dadd_rn(r, at(i));
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = at(j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], at(i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, at(i));
    return res;
}

// The same sum by one warp (all 32 lanes call it, all get the result): lane j < 8 owns accumulator j, so the
// dependent chain is n / 8 additions instead of n; the combination tree and the tail are those of the serial code.
template <typename F>
__device__ double np_pairwise_sum_warp(int n, F at) {
    const int lane = threadIdx.x & 31;
    if (n < 8) return np_pairwise_sum(n, at);
    double r = 0.0;
    const int n8 = n - (n % 8);
    if (lane < 8) {
        r = at(lane);
        int i = 8 + lane;
        for (; i + 24 < n8; i += 32) {   // four operands in flight ahead of the dependent additions
            const double v0 = at(i), v1 = at(i + 8), v2 = at(i + 16), v3 = at(i + 24);
            r = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r, v0), v1), v2), v3);
        }
        for (; i < n8; i += 8) r = __dadd_rn(r, at(i));
    }
    const double r0 = __shfl_sync(0xffffffffu, r, 0), r1 = __shfl_sync(0xffffffffu, r, 1), r2 = __shfl_sync(0xffffffffu, r, 2),
                 r3 = __shfl_sync(0xffffffffu, r, 3), r4 = __shfl_sync(0xffffffffu, r, 4), r5 = __shfl_sync(0xffffffffu, r, 5),
                 r6 = __shfl_sync(0xffffffffu, r, 6), r7 = __shfl_sync(0xffffffffu, r, 7);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (int i = n8; i < n; i++) res = __dadd_rn(res, at(i));
    return res;
}

// Python round() of a non-negative double: round half to even.
__device__ __forceinline__ int py_round(double x) { return (int)rint(x); }

// Middle element (rank (n-1)/2, n odd) or mean of the two middle elements of
// v[0..n) (n <= 256), by rank counting; the CTA cooperates, result in *out.
__device__ void small_median(const double* v, int n, double* out, double* tmp2) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < n) {
        const double x = v[t];
        int rank = 0;
        for (int r = 0; r < n; r++) {
            const double y = v[r];
            rank += (y < x) || (y == x && r < t);
        }
        if (rank == (n - 1) / 2) tmp2[0] = x;
        if (rank == n / 2) tmp2[1] = x;
    }
    __syncthreads();
    if (t == 0) *out = (n & 1) ? tmp2[0] : __ddiv_rn(__dadd_rn(tmp2[0], tmp2[1]), 2.0);
    __syncthreads();
}


// ---- scipy find_peaks(scores[lo:nc], distance=m_obs): local maxima + distance suppression ----------
// Works on the positions (lo, nc - 1) of score[] (find_peaks of the sub-array scores[lo:]: its first
// and last samples are never peaks and a plateau is judged by the same neighbours, so the maxima of
// the sub-array are the maxima of the full array whose plateau starts after lo).  Leaves the kept
// peaks, in order, in kp[0..P) and returns P; also prepares select_top_k's scratch (histogram zeroed,
// float range of the kept peaks' scores in s.vmin_key / s.vmax_key).
//   bm:  three bitmaps (peak / kept / removed), FP_BM_WORDS(cap) words each, position i = bit i + 32
//        (a 32-bit window around any position needs no bounds checks);
//   kp:  (nc - lo)/2 + 8 entries; until the list is written it holds, per peak, the set of its
//        higher-priority neighbours (two peaks are never adjacent, so position >> 1 is a unique slot).
static_assert((FP_MAX_LEN + FP_THREADS - 1) / FP_THREADS <= 64, "a thread's positions must fit a 64-bit mask");
static_assert(FP_THREADS >= 256, "the 256-bin histograms are zeroed by one thread per bin");
constexpr int FP_NEAR = 7;   // neighbour distances that fit the 15-bit windows (m_obs <= FP_NEAR + 1)
__host__ __device__ constexpr int FP_BM_WORDS(int cap) { return cap / 32 + 4; }

// bits pos .. pos + 31 of a bitmap (pos >= -32)
template <typename W>
__device__ __forceinline__ uint32_t bm_window32(const W* bm, int pos) {
    const int bi = pos + 32;
    return __funnelshift_r(bm[bi >> 5], bm[(bi >> 5) + 1], bi & 31);
}
__device__ __forceinline__ void bm_set(uint32_t* bm, int pos) { atomicOr(&bm[(pos + 32) >> 5], 1u << (pos & 31)); }

// Clears the bitmaps and select_top_k's scratch; a barrier must lie between this and find_kept_peaks.
__device__ __forceinline__ void peaks_scratch_clear(uint32_t* bm, int cap, FpScratch& s) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * FP_BM_WORDS(cap); i += FP_THREADS) bm[i] = 0u;
    if (tid < 256) s.hist[tid] = 0;
    if (tid == 0) {
        s.vmin_key = 0xffffffffu;
        s.vmax_key = 0u;
        s.ncand = 0;
        s.flag = 0;
    }
}

constexpr int FP_TOPK_CAND = FP_MAX_EVENTS + 2;   // keys of the threshold's histogram bin that can be ranked exactly
__device__ void select_top_k(const double* score, const uint16_t* kp, uint8_t* state, int P, int k_events, int add, int* out,
                             FpScratch& s, unsigned long long* cand);

// 256 logarithmic bins (1/8 octave, 2^-12 .. 2^20) of a non-negative score: monotone, needs no minimum / maximum
__device__ __forceinline__ int topk_bin(double x) {
    const int b = (int)(__float_as_uint((float)x) >> 20) - ((127 - 12) << 3);
    return min(255, max(0, b));
}

// find_peaks + the k_events highest-scoring kept peaks (sig_proc.py:176-198) in one go.  Returns the number P of kept
// peaks; if P >= k_events, out[0..k_events) = position + add of the selected peaks, in order.
__device__ int kept_peaks_select(const double* score, int lo, int nc, int m_obs, uint16_t* kp, uint32_t* bm, int cap,
                                 int k_events, int add, int* out, uint8_t* state, unsigned long long* cand, FpScratch& s) {
    const int tid = threadIdx.x;
    const int bw = FP_BM_WORDS(cap);
    uint32_t* pk = bm;
    uint32_t* kept_bm = bm + bw;
    uint32_t* rem_bm = bm + 2 * bw;
    // scipy _local_maxima_1d (strict maxima, plateaus -> midpoint): thread t walks the contiguous positions
    // [lo + t*chunk, lo + (t+1)*chunk) once (one load per position) and marks the midpoint of every maximum
    // that STARTS there.
    const int chunk = (nc - lo + FP_THREADS - 1) / FP_THREADS;
    const int o_begin = min(nc, lo + tid * chunk), o_end = min(nc, o_begin + chunk);
    {
        const int p_begin = max(lo + 1, o_begin), p_end = min(nc - 1, o_end);
        if (p_begin < p_end) {
            double prev = score[p_begin - 1];
            for (int i = p_begin; i < p_end; i++) {
                const double x = score[i];
                if (prev < x) {
                    int ahead = i + 1;
                    while (ahead < nc - 1 && score[ahead] == x) ahead++;
                    if (score[ahead] < x) bm_set(pk, (i + ahead - 1) >> 1);
                }
                prev = x;
            }
        }
    }
    __syncthreads();
    FP_T(s, 4);   // local maxima
    // the peaks at this thread's positions
    unsigned long long mine = 0ull;
    if (o_begin < o_end) {
        mine = (unsigned long long)bm_window32(pk, o_begin) | ((unsigned long long)bm_window32(pk, o_begin + 32) << 32);
        const int len = o_end - o_begin;
        if (len < 64) mine &= (1ull << len) - 1ull;
    }

    // ---- scipy _select_by_peak_distance ------------------------------------------------------------
    // A peak stays iff no STAYING peak of higher priority (score, then index) lies closer than m_obs
    // samples; the greedy highest-first sweep of scipy computes exactly this set, and the set is unique.
    // A verdict is written only when it is final (removed: a higher KEPT peak is near; kept: every higher
    // peak near is REMOVED) and the kept / removed sets only grow, so it does not matter how fresh they are
    // when they are read: no barrier between the rounds, every warp iterates over its own undecided peaks
    // until none is left (the highest undecided peak of the read can always be decided, so the loops end).
    unsigned long long kept = mine;
    if (m_obs > 1) {
        unsigned long long und = mine;
        kept = 0ull;
        if (m_obs <= FP_NEAR + 1) {
            // the higher-priority neighbours of every peak, once (scores are not looked at again), as a window:
            // bit b = the peak at position p - FP_NEAR + b; none -> kept at once
            const uint32_t near_mask = ((1u << (2 * m_obs - 1)) - 1u) << (FP_NEAR - (m_obs - 1)) & ~(1u << FP_NEAR);
            unsigned long long m = mine;
            while (m) {
                const int k = __ffsll((long long)m) - 1;
                m &= m - 1;
                const int p = o_begin + k;
                const double x = score[p];
                uint32_t nb = bm_window32(pk, p - FP_NEAR) & near_mask, h = 0;
                while (nb) {
                    const int b = __ffs((int)nb) - 1;
                    nb &= nb - 1;
                    const double y = score[p - FP_NEAR + b];
                    if (b < FP_NEAR ? (y > x) : (y >= x)) h |= 1u << b;   // equal scores: the higher index wins
                }
                if (h == 0) {
                    bm_set(kept_bm, p);
                    und &= ~(1ull << k);
                    kept |= 1ull << k;
                } else {
                    kp[p >> 1] = (uint16_t)h;
                }
            }
            FP_T(s, 5);   // neighbour sets
            const volatile uint32_t* vkept = kept_bm;
            const volatile uint32_t* vrem = rem_bm;
            for (int spin = 0; __any_sync(0xffffffffu, und != 0ull); spin++) {
                if (spin > (1 << 22)) __trap();   // a protocol bug must end as a launch failure, never as a hung GPU
                m = und;
                while (m) {
                    const int k = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const int p = o_begin + k;
                    const uint32_t h = kp[p >> 1];
                    if (bm_window32(vkept, p - FP_NEAR) & h) {
                        bm_set(rem_bm, p);
                        und &= ~(1ull << k);
                    } else if ((h & ~bm_window32(vrem, p - FP_NEAR)) == 0u) {
                        bm_set(kept_bm, p);
                        und &= ~(1ull << k);
                        kept |= 1ull << k;
                    }
                }
            }
        } else if (m_obs <= 16) {
            // wider neighbourhoods (tRNA: 9): the same rounds on 31-bit windows (bit b = position p - 15 + b); the sets of
            // higher neighbours do not fit the 16-bit slots, so a visit compares the scores of the neighbours that are
            // not removed yet again
            const volatile uint32_t* vkept = kept_bm;
            const volatile uint32_t* vrem = rem_bm;
            const uint32_t near_mask = (uint32_t)(((1ull << (2 * m_obs - 1)) - 1ull) << (15 - (m_obs - 1))) & ~(1u << 15);
            for (int spin = 0; __any_sync(0xffffffffu, und != 0ull); spin++) {
                if (spin > (1 << 22)) __trap();   // a protocol bug must end as a launch failure, never as a hung GPU
                unsigned long long m = und;
                while (m) {
                    const int k = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const int p = o_begin + k;
                    const double x = score[p];
                    uint32_t nb = bm_window32(pk, p - 15) & near_mask & ~bm_window32(vrem, p - 15);
                    const uint32_t kw = bm_window32(vkept, p - 15);
                    bool killed = false, blocked = false;
                    while (nb) {
                        const int b = __ffs((int)nb) - 1;
                        nb &= nb - 1;
                        const double y = score[p - 15 + b];
                        if (b < 15 ? (y > x) : (y >= x)) {  // equal scores: the higher index wins
                            if ((kw >> b) & 1u) killed = true;
                            else blocked = true;
                        }
                    }
                    if (killed) {
                        bm_set(rem_bm, p);
                        und &= ~(1ull << k);
                    } else if (!blocked) {
                        bm_set(kept_bm, p);
                        und &= ~(1ull << k);
                        kept |= 1ull << k;
                    }
                }
            }
        } else {
            const volatile uint32_t* vkept = kept_bm;
            const volatile uint32_t* vrem = rem_bm;
            for (int spin = 0; __any_sync(0xffffffffu, und != 0ull); spin++) {
                if (spin > (1 << 22)) __trap();   // a protocol bug must end as a launch failure, never as a hung GPU
                unsigned long long m = und;
                while (m) {
                    const int k = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const int p = o_begin + k;
                    const double x = score[p];
                    bool killed = false, blocked = false;
                    const int qlo = max(lo, p - m_obs + 1), qhi = min(nc - 1, p + m_obs - 1);
                    for (int q = qlo; q <= qhi; q++) {
                        if (q == p || !((pk[(q + 32) >> 5] >> (q & 31)) & 1u)) continue;
                        if ((vrem[(q + 32) >> 5] >> (q & 31)) & 1u) continue;
                        const double y = score[q];
                        if (q < p ? (y > x) : (y >= x)) {  // equal scores: the higher index wins
                            if ((vkept[(q + 32) >> 5] >> (q & 31)) & 1u) killed = true;
                            else blocked = true;
                        }
                    }
                    if (killed) {
                        bm_set(rem_bm, p);
                        und &= ~(1ull << k);
                    } else if (!blocked) {
                        bm_set(kept_bm, p);
                        und &= ~(1ull << k);
                        kept |= 1ull << k;
                    }
                }
            }
        }
    }
    FP_T(s, 6);   // suppression rounds (thread 0's warp)

    // ---- the k_events highest-scoring kept peaks -------------------------------------------------------
    // Usual case, without ever listing the kept peaks: every thread adds ITS kept peaks to a histogram over the
    // logarithmic bins, warp 0 finds the bin that holds rank k from the top (and the number of kept peaks), the few
    // members of that bin are ranked exactly on their 64-bit keys, and the peaks at or above the threshold key are
    // written in position order (one scan).  A crowded threshold bin, or a threshold score of which only some copies
    // are selected (ties -> the higher indices), takes the list-based selection below.
    {
        unsigned long long kk = kept;
        while (kk) {
            const int k = __ffsll((long long)kk) - 1;
            kk &= kk - 1;
            atomicAdd(&s.hist[topk_bin(score[o_begin + k])], 1u);
        }
    }
    __syncthreads();
    if (tid < 32) {  // warp 0 scans the 256 bins from the top, 8 per lane (lane 0 = bins 255..248)
        const uint32_t k = (uint32_t)k_events - 1;  // 0-based rank from the top
        uint32_t cnt[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            cnt[q] = s.hist[255 - (tid * 8 + q)];
            sum += cnt[q];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += t;
        }
        if (tid == 31) s.n_kept = (int)inc;
        uint32_t run = inc - sum;  // peaks in higher bins
        if (k >= run && k < inc) {  // at most one lane
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (k >= run && k < run + cnt[q]) {
                    s.sel_bin = 255 - (tid * 8 + q);
                    s.sel_below = k - run;     // rank from the top inside the bin
                    s.sel_count = cnt[q];
                }
                run += cnt[q];
            }
        }
    }
    __syncthreads();
    const int P = s.n_kept;
    if (P < k_events) return P;   // uniform
    bool fast = s.sel_count <= (uint32_t)FP_TOPK_CAND;
    if (fast) {
        const int sel_bin = s.sel_bin;
        const uint32_t r_in = s.sel_below, m = s.sel_count;
        unsigned long long kk = kept;
        while (kk) {
            const int k = __ffsll((long long)kk) - 1;
            kk &= kk - 1;
            const double x = score[o_begin + k];
            if (topk_bin(x) == sel_bin) cand[atomicAdd(&s.ncand, 1u)] = (unsigned long long)__double_as_longlong(x);
        }
        __syncthreads();
        for (uint32_t t = tid; t < m; t += FP_THREADS) {
            const unsigned long long x = cand[t];
            uint32_t g = 0, e = 0;
            for (uint32_t u = 0; u < m; u++) {
                const unsigned long long y = cand[u];
                g += (y > x);
                e += (y == x);
            }
            if (r_in >= g && r_in < g + e) {   // every copy of the threshold key writes the same values
                s.sel_prefix64 = x;
                s.flag = (r_in - g + 1 == e) ? 2 : 1;   // 2: every copy of the threshold is selected
            }
        }
        __syncthreads();
        fast = s.flag == 2;
    }
    if (fast) {
        const unsigned long long thr_key = s.sel_prefix64;
        unsigned long long sel = 0ull, kk = kept;
        while (kk) {
            const int k = __ffsll((long long)kk) - 1;
            kk &= kk - 1;
            if ((unsigned long long)__double_as_longlong(score[o_begin + k]) >= thr_key) sel |= 1ull << k;
        }
        uint32_t tot_sel = 0;
        uint32_t so = block_exscan((uint32_t)__popcll(sel), s, &tot_sel);
        while (sel) {
            const int k = __ffsll((long long)sel) - 1;
            sel &= sel - 1;
            out[so++] = o_begin + k + add;
        }
        __syncthreads();
        return P;
    }

    // ---- list-based selection: kept peaks in order (the neighbour sets in kp are dead: barriers lie in between),
    // float range of their scores and a cleared histogram for select_top_k
    if (tid < 256) s.hist[tid] = 0;
    if (tid == 0) {
        s.vmin_key = 0xffffffffu;
        s.vmax_key = 0u;
        s.ncand = 0;
        s.flag = 0;
    }
    uint32_t total = 0;
    uint32_t koff = block_exscan((uint32_t)__popcll(kept), s, &total);
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    while (kept) {
        const int k = __ffsll((long long)kept) - 1;
        kept &= kept - 1;
        kp[koff++] = (uint16_t)(o_begin + k);
        const uint32_t kf = f32_key((float)score[o_begin + k]);
        kmin = min(kmin, kf);
        kmax = max(kmax, kf);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((tid & 31) == 0 && kmax >= kmin) {
        atomicMin(&s.vmin_key, kmin);
        atomicMax(&s.vmax_key, kmax);
    }
    __syncthreads();
    select_top_k(score, kp, state, P, k_events, add, out, s, cand);
    return P;
}

constexpr int FP_RANK_MAX = 192;

// ---- the k highest-scoring of the P kept peaks (sig_proc.py:188), in position order ---------------
// out[0..k) = kp[i] + add for the selected peaks.  Radix-selects the k-th largest score (scores are
// >= 0, so their bit patterns order like the values); ties -> the higher indices.  Requires P >= k.
__device__ void select_top_k(const double* score, const uint16_t* kp, uint8_t* state, int P, int k_events, int add,
                             int* out, FpScratch& s, unsigned long long* cand /* FP_TOPK_CAND keys of scratch */) {
    const int tid = threadIdx.x;
    if (P <= FP_RANK_MAX) {
        // A few hundred peaks: rank every peak against all others (descending score, ties -> the higher index first)
        // and keep ranks < k - one pass and two barriers instead of the eight radix passes below.
        __syncthreads();
        for (int i = tid; i < P; i += FP_THREADS) {
            const unsigned long long ki = (unsigned long long)__double_as_longlong(score[kp[i]]);
            int rank = 0;
            for (int j = 0; j < P; j++) {
                const unsigned long long kj = (unsigned long long)__double_as_longlong(score[kp[j]]);
                rank += (kj > ki) || (kj == ki && j > i);
            }
            state[i] = rank < k_events ? 4 : 0;
        }
        __syncthreads();
        const int pchunk = (P + FP_THREADS - 1) / FP_THREADS;
        const int i0 = min(P, tid * pchunk), i1 = min(P, i0 + pchunk);
        uint32_t mysel = 0;
        for (int i = i0; i < i1; i++) mysel += (state[i] == 4);
        uint32_t tot_sel = 0;
        uint32_t so = block_exscan(mysel, s, &tot_sel);
        for (int i = i0; i < i1; i++)
            if (state[i] == 4) out[so++] = (int)kp[i] + add;  // already sorted
        __syncthreads();
        return;
    }
    unsigned long long thr_key;
    bool have_thr = false, all_ties = false;
    {
        // The k-th largest score in two light passes: a 256-bin histogram over the float range of the peaks' scores
        // (minimum / maximum and the zeroed histogram come from find_kept_peaks; float(x) and the binning are monotone,
        // so a higher bin holds only larger scores), and the exact ranking of the few members of the bin that holds
        // rank k on their 64-bit keys.  Falls through to the radix selection below for degenerate score sets.
        const float vmin = f32_unkey(s.vmin_key), vmax = f32_unkey(s.vmax_key);
        const float scale = __fdiv_rn(256.0f, __fsub_rn(vmax, vmin));
        if (vmax > vmin && scale < 1e30f) {   // uniform
            auto bin_of = [&](double x) { return min(255, (int)__fmul_rn(__fsub_rn((float)x, vmin), scale)); };
            for (int i = tid; i < P; i += FP_THREADS) atomicAdd(&s.hist[bin_of(score[kp[i]])], 1u);
            __syncthreads();
            if (tid < 32) {  // warp 0 scans the 256 bins from the top, 8 per lane (lane 0 = bins 255..248)
                const uint32_t k = (uint32_t)k_events - 1;  // 0-based rank from the top
                uint32_t cnt[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    cnt[q] = s.hist[255 - (tid * 8 + q)];
                    sum += cnt[q];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (tid >= o) inc += t;
                }
                uint32_t run = inc - sum;  // peaks in higher bins
                if (k >= run && k < inc) {  // exactly one lane
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (k >= run && k < run + cnt[q]) {
                            s.sel_bin = 255 - (tid * 8 + q);
                            s.sel_below = k - run;     // rank from the top inside the bin
                            s.sel_count = cnt[q];
                        }
                        run += cnt[q];
                    }
                }
            }
            __syncthreads();
            const int sel_bin = s.sel_bin;
            const uint32_t r_in = s.sel_below, m = s.sel_count;
            if (m <= (uint32_t)FP_TOPK_CAND) {   // uniform
                for (int i = tid; i < P; i += FP_THREADS) {
                    const double x = score[kp[i]];
                    if (bin_of(x) == sel_bin) cand[atomicAdd(&s.ncand, 1u)] = (unsigned long long)__double_as_longlong(x);
                }
                __syncthreads();
                for (uint32_t t = tid; t < m; t += FP_THREADS) {
                    const unsigned long long x = cand[t];
                    uint32_t g = 0, e = 0;
                    for (uint32_t u = 0; u < m; u++) {
                        const unsigned long long y = cand[u];
                        g += (y > x);
                        e += (y == x);
                    }
                    if (r_in >= g && r_in < g + e) {   // every copy of the threshold key writes the same values
                        s.sel_prefix64 = x;
                        s.sel_k = r_in - g;            // ties ranked above the selected one
                        s.flag = (r_in - g + 1 == e) ? 2 : 1;   // 2: every copy of the threshold is selected (no tie ranking)
                    }
                }
                __syncthreads();
                have_thr = s.flag != 0;
                all_ties = s.flag == 2;
            }
        }
    }
    if (have_thr) {
        thr_key = s.sel_prefix64;
    } else {
        unsigned long long prefix = 0, mask = 0;
        uint32_t k = (uint32_t)k_events - 1;  // 0-based rank from the top
        for (int shift = 56; shift >= 0; shift -= 8) {
            __syncthreads();
            if (tid < 256) s.hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < P; i += FP_THREADS) {
                const unsigned long long kv = (unsigned long long)__double_as_longlong(score[kp[i]]);
                if ((kv & mask) == prefix) atomicAdd(&s.hist[(uint32_t)(kv >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 32) {  // warp 0 scans the 256 bins from the top, 8 per lane (lane 0 = bins 255..248)
                uint32_t cnt[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    cnt[q] = s.hist[255 - (tid * 8 + q)];
                    sum += cnt[q];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (tid >= o) inc += t;
                }
                uint32_t run = inc - sum;  // elements in higher bins
                if (k >= run && k < inc) {  // exactly one lane
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (k >= run && k < run + cnt[q]) {
                            s.sel_prefix64 = prefix | ((unsigned long long)(255 - (tid * 8 + q)) << shift);
                            s.sel_k = k - run;
                        }
                        run += cnt[q];
                    }
                }
            }
            __syncthreads();
            prefix = s.sel_prefix64;
            k = s.sel_k;
            mask |= 255ull << shift;
        }
        thr_key = prefix;
        // k = how many elements EQUAL to the threshold rank above the selected one, i.e.
        // (k + 1) of the ties are taken; ties -> the higher indices
    }
    const uint32_t ties_needed = s.sel_k + 1;
    __syncthreads();
    // mark the selection, then compact in order
    const int pchunk = (P + FP_THREADS - 1) / FP_THREADS;
    const int i0 = min(P, tid * pchunk), i1 = min(P, i0 + pchunk);
    uint32_t tot_ties = 0, tie_off = 0;
    if (all_ties) {   // uniform: the usual case (one copy of the threshold score) needs no ranking of the ties
        tot_ties = ties_needed;
    } else {
        uint32_t my_ties = 0;
        for (int i = i0; i < i1; i++) my_ties += ((unsigned long long)__double_as_longlong(score[kp[i]]) == thr_key);
        tie_off = block_exscan(my_ties, s, &tot_ties);  // ties before this thread's chunk
    }
    uint32_t mysel = 0;
    for (int i = i0; i < i1; i++) {
        const unsigned long long kv = (unsigned long long)__double_as_longlong(score[kp[i]]);
        bool sel = kv > thr_key;
        if (kv == thr_key) {
            const uint32_t ties_after = tot_ties - tie_off - 1;  // ties at higher index
            sel = ties_after < ties_needed;
            tie_off++;
        }
        state[i] = sel ? 4 : 0;
        mysel += sel;
    }
    uint32_t tot_sel = 0;
    uint32_t so = block_exscan(mysel, s, &tot_sel);
    for (int i = i0; i < i1; i++)
        if (state[i] == 4) out[so++] = (int)kp[i] + add;  // already sorted
    __syncthreads();
}

// ---- consensus sub-sequence match (sig_proc.py:288-312) ----------------------------------------------
// dtaidistance warping_paths(query, series, penalty, psi = (psi_q, 0, psi_s, 0)) without window:
//     P[0][0..psi_s] = 0, P[0..psi_q][0] = 0, +inf elsewhere on the border,
//     P[i+1][j+1] = (q[i] - x[j])^2 + min(P[i][j], P[i][j+1] + pen2, P[i+1][j] + pen2),
// then SubsequenceAlignment: matching[j] = sqrt(P[Q][j+1]) / Q, end = first argmin, start = column of the
// first cell of best_path(paths, col = end + 1) (walk back to the FIRST minimum of sqrt(diagonal),
// sqrt(up), sqrt(left), no penalty).  The matrix is never stored: every lane owns one query row and
// sweeps the columns as an anti-diagonal wavefront (values handed down the lanes by __shfl_up_sync);
// the start column of the walk-back travels FORWARD with every cell (origin of a cell = origin of the
// predecessor the walk-back would choose, or the cell's own column when that predecessor lies on the
// border), so the answer for every end column is known when the last row is reached.
// sqrt(a) < sqrt(b) as the walk-back compares them (correctly rounded square roots can collide):
__device__ __forceinline__ bool lt_sqrt(double a, double b) {
    if (!(a < b)) return false;
    if (__dsub_rn(b, a) > __dmul_rn(b, 1.7763568394002505e-15 /* 2^-49 */)) return true;  // more than 8 ulp apart: the roots differ
    // a finite, b = +inf (border cells: every anti-diagonal of the first Q steps has one): inf - a > inf * 2^-49 is false, and
    // without this line the whole warp would take the two software square roots below for it; sqrt(a) < inf holds exactly
    if (b == __longlong_as_double(0x7ff0000000000000LL)) return true;
    return __dsqrt_rn(a) < __dsqrt_rn(b);
}

// ONE ROW PER LANE over NW warps (Q <= 32 * NW): a single warp would walk ceil(Q / 32) dependent rows per lane and
// step, and with one warp per CTA at work that dependency chain was 60 % of the consensus kernel's time
// (profiles/r01z_ncu_trna_lines.txt).  Lane g owns row g; inside a warp the hand-down is the same __shfl_up_sync, between
// warps lane 31 leaves its value in a double-buffered shared slot and the NW warps meet at a named barrier once per
// anti-diagonal step.  Cell arithmetic, walk-back origin and the final argmin are unchanged (bit-identical results).
template <int NW>
__device__ void consensus_match_rows(const double* __restrict__ query, int Q, const double* series, int Cn, double pen2,
                                     int psi_q, int psi_s, double* lastrow, int* lastorg, int* result /*[2] start, end*/,
                                     double* hand_v /*[2][NW]*/, int* hand_o /*[2][NW]*/) {
    const int g = threadIdx.x;   // caller: threadIdx.x < 32 * NW
    const int lane = g & 31, warp = g >> 5;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const bool own = g < Q;
    const double qv = own ? query[g] : 0.0;
    double left = (g + 1 <= psi_q) ? 0.0 : inf;   // P[g+1][0]
    int lorg = -1;
    double diag_in = (g <= psi_q) ? 0.0 : inf;    // P[g][0]
    int diag_org = -1;
    double pass_v = inf;
    int pass_o = -1;
    if (lane == 31) {
        hand_v[warp] = inf;
        hand_o[warp] = -1;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
    const int steps = Cn + Q - 1;
    for (int st = 0; st < steps; st++) {
        double up_in = __shfl_up_sync(0xffffffffu, pass_v, 1);
        int up_org = __shfl_up_sync(0xffffffffu, pass_o, 1);
        if (lane == 0 && warp > 0) {
            up_in = hand_v[(st & 1) * NW + warp - 1];
            up_org = hand_o[(st & 1) * NW + warp - 1];
        }
        const int j = st - g;
        if (g == 0) {
            up_in = (j + 1 <= psi_s) ? 0.0 : inf;  // P[0][j+1]
            up_org = -1;
        }
        if (j >= 0 && j < Cn && own) {
            const double x = series[j];
            const double up = up_in, dg = diag_in, lf = left;
            const int uo = up_org, dgo = diag_org, lfo = lorg;
            const double df = __dsub_rn(qv, x);
            const double d = __dmul_rn(df, df);
            double m = dg;
            double t = __dadd_rn(up, pen2);
            if (t < m) m = t;
            t = __dadd_rn(lf, pen2);
            if (t < m) m = t;
            const double val = __dadd_rn(d, m);
            // walk-back choice at this cell
            double best = dg;
            int po = dgo;
            if (lt_sqrt(up, best)) { best = up; po = uo; }
            if (lt_sqrt(lf, best)) { best = lf; po = lfo; }
            const int org = (po < 0) ? j : po;
            left = val;
            lorg = org;
            if (g == Q - 1) { lastrow[j] = val; lastorg[j] = org; }
            pass_v = val;
            pass_o = org;
            diag_in = up_in;
            diag_org = up_org;
        }
        if (lane == 31) {
            hand_v[((st + 1) & 1) * NW + warp] = pass_v;
            hand_o[((st + 1) & 1) * NW + warp] = pass_o;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NW * 32) : "memory");
    }
    if (warp == 0) {
        // matching = sqrt(last row) / Q; first minimum (np.argmin)
        const double Qd = (double)Q;
        double bv = inf;
        int bi = 0x7fffffff;
        for (int j = lane; j < Cn; j += 32) {
            const double v = __ddiv_rn(__dsqrt_rn(lastrow[j]), Qd);
            if (v < bv) { bv = v; bi = j; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (bi == 0x7fffffff) bi = 0;  // every entry +inf/NaN: argmin returns 0
        if (lane == 0) {
            result[0] = lastorg[bi];
            result[1] = bi;
        }
    }
}

// A read fails: NaN fingerprint, zero dwell times, NaN statistics, its status code (the whole CTA calls this).
__device__ __forceinline__ void fp_fail(const FpArgs& a, int64_t read, int nb, int code, bool cons) {
    const int tid = threadIdx.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int q = tid; q < nb; q += FP_THREADS) {
        a.fpt[read * nb + q] = qnan;
        if (a.dwell) a.dwell[read * nb + q] = 0;
    }
    if (a.stats && tid < 6) a.stats[read * 6 + tid] = qnan;
    if (cons && a.cons && tid < 3) a.cons[read * 3 + tid] = 0;
    if (tid == 0) a.status[read] = code;
}

// c_new_means (_c_segmentation.pyx:41-53): sequential float64 sums; eight lanes per segment when the sums are exact
__device__ __forceinline__ void fp_segment_means(const float* base, int nseg, const int* cpts, double* ev, bool exact_sums) {
    const int tid = threadIdx.x;
    if (exact_sums) {
        // four segments per warp and trip, eight lanes each
        const int lane = tid & 31, warp = tid >> 5, sub = lane >> 3, sl = lane & 7;
        for (int q0 = warp * 4; q0 < nseg; q0 += FP_WARPS * 4) {
            const int q = q0 + sub;
            const int b = (q < nseg) ? cpts[q] : 0, e = (q < nseg) ? cpts[q + 1] : 0;
            double sum = 0.0;
            for (int i = b + sl; i < e; i += 8) sum = __dadd_rn(sum, (double)base[i]);
#pragma unroll
            for (int o = 4; o; o >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
            if (sl == 0 && q < nseg) ev[q] = __ddiv_rn(sum, (double)(e - b));
        }
    } else {
        for (int q = tid; q < nseg; q += FP_THREADS) {
            double sum = 0.0;
            const int b = cpts[q], e = cpts[q + 1];
            for (int i = b; i < e; i++) sum = __dadd_rn(sum, (double)base[i]);
            ev[q] = __ddiv_rn(sum, (double)(e - b));
        }
    }
}

// The last barcode_num_events of the n_out_seg segments (bounds in cpts, means in ev), normalised by the mean / std of
// the adapter events (sig_proc.py:569-594; normalize "mean" :546-552, or normalize_wrt for the refined events :482-484).
__device__ __forceinline__ void fp_write_fingerprint(const FpArgs& a, int64_t read, int nb, int n_out_seg, const int* cpts, const double* ev,
                                                     double ev_mean, double ev_std) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const int keep = min(nb, n_out_seg);
    for (int q = threadIdx.x; q < nb; q += FP_THREADS) {
        const int srcq = n_out_seg - keep + (q - (nb - keep));
        double v = qnan;  // front NaN padding if fewer events than asked (unreachable with accept_less_cpts=false)
        int64_t dw = 0;
        if (q >= nb - keep) {
            v = __ddiv_rn(__dsub_rn(ev[srcq], ev_mean), ev_std);
            dw = (int64_t)(cpts[srcq + 1] - cpts[srcq]);
        }
        a.fpt[read * nb + q] = v;
        if (a.dwell) a.dwell[read * nb + q] = dw;
    }
}

// Consensus refinement behind the alignment (sig_proc.py:334-378, 500-521): second segmentation of the scores from the
// barcode start on, event means, outlier filter, fingerprint.  cpts holds the FIRST segmentation's change points on entry.
// Returns false when the read failed (status written).  The whole CTA calls this.
__device__ bool cons_tail(const FpConfig& c, const FpArgs& a, FpScratch& s, int64_t read, int cap, const double* score, const float* sig,
                          uint16_t* kp, uint8_t* state, uint32_t* bm, int* cpts, double* ev, double* dv, int n, int nc, int w,
                          bool exact_sums, double ev_mean, double ev_std, int q_start, int q_end) {
    const int tid = threadIdx.x;
    const int nb = c.barcode_num_events;
    const int sbs = cpts[q_end];  // sig_barcode_start = sum(adapter_dwell_times[:q_end]) (sig_proc.py:334)
    peaks_scratch_clear(bm, cap, s);
    __syncthreads();              // cpts / ev are rewritten below
    // second segmentation on barcode_scores = adapter_scores[sbs:] with the UNCAPPED min_obs_per_base
    // and running_stat_width (sig_proc.py:336-365)
    const int ke = c.cons_seg_events;
    const int P2 = (nc - sbs >= 3) ? kept_peaks_select(score, sbs, nc, c.min_obs_per_base, kp, bm, cap, ke, w - sbs /* relative to raw_signal[sbs:] */,
                                                       cpts + 1, state, reinterpret_cast<unsigned long long*>(dv), s)
                                   : 0;
    if (P2 < ke) {
        fp_fail(a, read, nb, FP_FAIL_SEGMENTATION, true);
        return false;
    }
    if (tid == 0) {
        cpts[0] = 0;
        cpts[ke + 1] = n - sbs;   // scores.size + 2 * running_stat_width = (nc - sbs) + 2 w
    }
    __syncthreads();
    fp_segment_means(sig + sbs, ke + 1, cpts, ev, exact_sums);  // compute_base_means(raw_signal[sbs:], cpts) (:368)
    __syncthreads();
    if (a.cons && tid == 0) {
        a.cons[read * 3 + 0] = q_start;
        a.cons[read * 3 + 1] = q_end;
        a.cons[read * 3 + 2] = sbs;
    }
    if (q_start > c.cons_ub_start || q_end < c.cons_lb_end || q_end > c.cons_ub_end) {  // sig_proc.py:500-521
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        for (int q = tid; q < nb; q += FP_THREADS) {
            a.fpt[read * nb + q] = qnan;
            if (a.dwell) a.dwell[read * nb + q] = 0;
        }
        if (tid == 0) a.status[read] = FP_FAIL_CONSENSUS;
        return false;
    }
    fp_write_fingerprint(a, read, nb, ke + 1, cpts, ev, ev_mean, ev_std);
    if (tid == 0) a.status[read] = FP_OK;
    return true;
}

// CONS = consensus-guided barcode refinement (tRNA configurations, sig_proc.py:257-378, 451-521).
// PHASE 0: the whole chain in one launch.  PHASE 1 (consensus only, big batches): up to the normalised event series, the
// read's state is parked in global memory and the alignment / refinement follow as fingerprint_cons_align_kernel (many small
// CTAs per SM: the alignment keeps three warps busy for ~40 % of a read's time here) and fingerprint_cons_tail_kernel.
template <bool CONS, int PHASE = 0>
#ifndef WDX_FP_MIN_CTAS
#define WDX_FP_MIN_CTAS (1024 / FP_THREADS)
#endif
__global__ void __launch_bounds__(FP_THREADS, WDX_FP_MIN_CTAS) fingerprint_kernel(const __grid_constant__ FpConfig c,
                                                                   const __grid_constant__ FpArgs a) {
    extern __shared__ __align__(16) unsigned char fp_smem[];
    const int cap = a.cap;
    double* score = reinterpret_cast<double*>(fp_smem);                       // [cap]
    float* sig_al = reinterpret_cast<float*>(fp_smem + (size_t)cap * 8);      // [cap + 8]: the slice starts at sig_al + (0..3)
    uint16_t* kp = reinterpret_cast<uint16_t*>(fp_smem + (size_t)cap * 12 + 32);   // [cap/2 + 8] kept peak positions
    uint8_t* state = fp_smem + (size_t)cap * 12 + 32 + ((size_t)(cap / 2 + 8) * 2);  // [cap/2 + 8] one byte per kept peak (top-k)
    uint32_t* bm = reinterpret_cast<uint32_t*>(fp_smem + (size_t)cap * 12 + 32 + ((size_t)(cap / 2 + 8) * 3));  // peak / kept / removed bitmaps
    __shared__ FpScratch s;
    __shared__ int cpts[FP_MAX_EVENTS + 2];
    __shared__ double ev[FP_MAX_EVENTS + 2];   // event means, later normalised
    __shared__ double dv[FP_MAX_EVENTS + 2];   // scratch for the statistics
    __shared__ double red[8];

    const int tid = threadIdx.x;
    const int64_t read = blockIdx.x;
    if (read >= a.n) return;
    if (a.retry_status && a.status[read] != a.retry_status) return;
    FP_T_BEGIN(s);
    const int nb = c.barcode_num_events;
    double* fpt_out = a.fpt + read * nb;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    auto fail = [&](int code) { fp_fail(a, read, nb, code, CONS); };  // whole CTA calls this (uniform)

    // the read's four scalars are requested together (one round trip to memory instead of three dependent ones)
    const uint8_t ok_in = a.detect_ok ? a.detect_ok[read] : (uint8_t)1;
    const int64_t a_start = a.adapter_start[read], a_end = a.adapter_end[read];
    const int64_t len_in = a.sig_len ? (int64_t)a.sig_len[read] : a.stride;
    if (!ok_in) {  // sig_proc.py:400-407
        fail(FP_FAIL_DETECT);
        return;
    }

    // ---- extract_adapter (sig_proc.py:382-391) --------------------------------
    const int64_t len_row = min(len_in, a.stride);
    int64_t start = a_start - c.padding;
    if (start < 0) start = 0;
    int64_t stop = a_end + c.padding;
    if (stop > len_row) stop = len_row;
    int64_t n64 = stop - start;
    if (n64 < 0) n64 = 0;
    if (n64 > cap) {
        fail(FP_FAIL_TOO_LONG);
        return;
    }
    int n = (int)n64;
    const float* src = a.signals + read * a.stride + start;
    if (tid == 0) {
        s.first_nan = n;
        s.vmin_key = 0xffffffffu;
        s.vmax_key = 0u;
        s.amin = 0xffffffffu;
        s.amax = 0u;
        s.ncand = 0;
        s.ncand2 = 0;
    }
    // the two median histograms (the score array is idle until the t-test) are cleared while the slice is on its way
    uint32_t* med_hist = reinterpret_cast<uint32_t*>(score);
    const bool lin_fits = cap >= (2 * FP_MED_BINS + FP_MED_CAND) / 2;
    if (lin_fits)
        for (int b = tid; b < 2 * FP_MED_BINS; b += FP_THREADS) med_hist[b] = 0u;
    __syncthreads();
    // The one HBM read of the slice: 16-byte loads from the aligned address below the slice's first sample (the
    // staging buffer keeps the same misalignment, so the shared-memory stores are whole vectors too; the up to three
    // samples in front of and behind the slice that come along are never looked at), minimum / maximum on the way.
    const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    float* sig = sig_al + mis;
    {
        const float4* src4 = reinterpret_cast<const float4*>(src - mis);
        float4* dst4 = reinterpret_cast<float4*>(sig_al);
        const int nv = (n + mis + 3) >> 2;
        uint32_t kmin = 0xffffffffu, kmax = 0u;
        for (int v0 = tid; v0 < nv; v0 += 2 * FP_THREADS) {   // two vectors in flight per thread before the first use
            float4 xv[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int v = v0 + u * FP_THREADS;
                xv[u] = (v < nv) ? __ldg(src4 + v) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int v = v0 + u * FP_THREADS;
                if (v >= nv) break;
                dst4[v] = xv[u];
                const float xe[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int i = 4 * v + e - mis;
                    if (i < 0 || i >= n) continue;
                    const float x = xe[e];
                    if (x != x) {
                        atomicMin(&s.first_nan, i);
                    } else {
                        const uint32_t kx = f32_key(x);
                        kmin = min(kmin, kx);
                        kmax = max(kmax, kx);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(&s.vmin_key, kmin);
            atomicMax(&s.vmax_key, kmax);
        }
    }
    __syncthreads();
    FP_T(s, 0);   // load
    // NaN padding inside the slice (a read that ends less than `padding` samples after its adapter):
    // the reference hands the padded minibatch row to detect_results_to_fpt (file_proc.py:418-428), so
    // the slice keeps its full length for the segmentation parameters, medians ignore the NaNs
    // (np.nanmedian), scores next to the NaNs are NaN and never peaks, and the last event mean is NaN:
    // the read fails with "segment normalization failed" unless it already failed for too few peaks.
    const int n_total = n;
    const bool trimmed = s.first_nan < n;
    n = min(n, s.first_nan);

    // ---- segmentation parameters (sig_proc.py:526-533; Python round = half to even)
    const int m_obs = min(c.min_obs_per_base, py_round((double)n_total / (double)c.num_events / 2.0));
    const int w = min(c.running_stat_width, py_round((double)n_total / (double)c.num_events));
    const int nc = n - 2 * w;  // number of finite t-test positions
    if (m_obs < 1 || w < 1 || nc < 3) {  // find_peaks(distance < 1) raises -> the reference reports a failed read
        fail(FP_FAIL_SEGMENTATION);
        return;
    }

    // ---- winsorise at med +- thresh * MAD, float32 (sig_proc.py:421-431) --------
    uint32_t* med_cand = med_hist + 2 * FP_MED_BINS;
    const bool lin = !trimmed && lin_fits;  // min/max cover exactly the slice; scratch fits
    float med, mad;
    if (lin) {
        const float vmin = f32_unkey(s.vmin_key), vmax = f32_unkey(s.vmax_key);
        med = block_median_f32_linear(n, [&](int i) { return sig[i]; }, vmin, vmax, med_hist, med_cand, &s.ncand, s);
        const float ymax = fmaxf(__fsub_rn(vmax, med), __fsub_rn(med, vmin));  // >= every |x - med| (rounding is monotone)
        mad = block_median_f32_linear(n, [&](int i) { return fabsf(__fsub_rn(sig[i], med)); }, 0.0f, ymax,
                                      med_hist + FP_MED_BINS, med_cand, &s.ncand2, s);
    } else {
        med = block_median_f32(n, [&](int i) { return f32_key(sig[i]); }, s);
        mad = block_median_f32(n, [&](int i) { return f32_key(fabsf(__fsub_rn(sig[i], med))); }, s);
    }
    FP_T(s, 1);   // medians
    float lo, hi;
    if (c.numpy1_promotion) {
        const double tmd = __dmul_rn(c.outlier_thresh_d, (double)mad);
        lo = (float)__dsub_rn((double)med, tmd);
        hi = (float)__dadd_rn((double)med, tmd);
    } else {   // numpy >= 2 (NEP 50): every step in float32
        const float tm = __fmul_rn(c.outlier_thresh, mad);
        lo = __fsub_rn(med, tm);
        hi = __fadd_rn(med, tm);
    }
    __syncthreads();
    {
        uint32_t amin = 0xffffffffu, amax = 0u;   // |x| range of the winsorised slice (non-zero values), as bit patterns
        for (int i = tid; i < n; i += FP_THREADS) {
            float x = sig[i];
            x = (x < lo) ? lo : x;  // np.clip = min(max(x, lo), hi)
            x = (x > hi) ? hi : x;
            sig[i] = x;
            if (a.signals_mut) a.signals_mut[read * a.stride + start + i] = x;  // the reference clips its view in place
            const uint32_t ab = __float_as_uint(x) & 0x7fffffffu;
            if (ab) {
                amin = min(amin, ab);
                amax = max(amax, ab);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            amin = min(amin, __shfl_xor_sync(0xffffffffu, amin, o));
            amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(&s.amin, amin);
            atomicMax(&s.amax, amax);
        }
    }
    __syncthreads();
    // Sums of these samples in float64 are EXACT — and therefore independent of the order of the additions — when the
    // binary exponents of the non-zero samples span at most 14: every sample is a multiple of u = 2^(e_lo - 23), every
    // partial sum of fewer than 2^14 samples is a multiple of u below 2^(e_hi + 15), i.e. an integer of fewer than
    // 53 bits times u.  (pA signals: 40 .. 200.)  The reference's sequential sums (window means of the t-test, event
    // means) may then be formed in any order, bit for bit; otherwise they are formed in the reference's order.
    bool exact_sums;
    {
        const int e_lo = (int)(s.amin >> 23), e_hi = (int)(s.amax >> 23);
        exact_sums = s.amax == 0u || (e_lo >= 1 && e_hi <= 254 && e_hi - e_lo <= 14);
    }
    static_assert(FP_MAX_LEN <= (1 << 14), "exact-sum bound");

    FP_T(s, 2);   // clip
    // ---- c_windowed_t_test (_c_segmentation.pyx:124-161), float64, reference order
    peaks_scratch_clear(bm, cap, s);   // for find_kept_peaks behind the t-test's barrier (the median scratch is done with)
    const double wd = (double)w;
    if (w == 12) {  // the capped width (every adapter of >= 1265 samples): unrolled, window in registers
        // A lane owns the twelve consecutive window starts 12*blk .. 12*blk + 11: the window slides through its registers
        // (one new sample per window), and the second window of position pos — the window that starts at pos + 12 — is
        // the one its right-hand neighbour lane holds at the same step (same operands, same order: bit for bit the
        // statistics the reference computes twice).  A warp covers 31 blocks of positions per pass; lane 31 only supplies
        // second windows.  With exact sums the window sum slides too (minus the sample that leaves, plus the one
        // that enters).
        const double wr = 1.0 / 12.0;
        const int nblk = (nc + 11) / 12;
        const int lane = tid & 31, warp = tid >> 5;
        // Shorter slices (most real adapters): a lane of the scheme above still walks twelve windows while half the
        // warps idle.  There every thread takes one residue class r = pos mod 12 and a segment of its chain r, r + 12,
        // r + 24, ... instead (the second window of one position is the first window of the next), which spreads the
        // windows over all threads: ceil(chain / 42) + 1 windows each, at ~1.3 x the instructions per window.
        const int seg_len = (nblk + FP_THREADS / 12 - 1) / (FP_THREADS / 12);
        if (seg_len <= 8) {
            if (exact_sums) ttest_residue_chains<12, true>(sig, score, nc, nblk, seg_len);   // uniform
            else ttest_residue_chains<12, false>(sig, score, nc, nblk, seg_len);
        } else
        for (int g = warp * 31; g < nblk; g += FP_WARPS * 31) {
            const int s0 = 12 * (g + lane);
            double x[12];
#pragma unroll
            for (int i = 0; i < 12; i++) x[i] = (double)sig[min(s0 + i, n - 1)];
            double msum = 0.0;
#pragma unroll
            for (int i = 0; i < 12; i++) msum = __dadd_rn(msum, x[i]);
#pragma unroll
            for (int k = 0; k < 12; k++) {
                if (k > 0) {   // window start s0 + k: x[(k + i) % 12], i = 0..11, oldest first
                    const double xn = (double)sig[min(s0 + k + 11, n - 1)];
                    if (exact_sums) {
                        msum = __dadd_rn(__dsub_rn(msum, x[k - 1]), xn);
                        x[k - 1] = xn;
                    } else {
                        x[k - 1] = xn;
                        msum = 0.0;
#pragma unroll
                        for (int i = 0; i < 12; i++) msum = __dadd_rn(msum, x[(k + i) % 12]);
                    }
                }
                const double m = div_small_int(msum, 12.0, wr);
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const double pd = __dsub_rn(x[(k + i) % 12], m);
                    v = __dadd_rn(v, __dmul_rn(pd, pd));
                }
                const double m2 = __shfl_down_sync(0xffffffffu, m, 1), v2 = __shfl_down_sync(0xffffffffu, v, 1);
                const int pos = s0 + k;
                if (lane < 31 && pos < nc) score[pos] = ttest_combine(m, v, m2, v2);
            }
        }
    } else if (w == 18) {   // the tRNA configuration's width: residue-class chains (every window's statistics once)
        const int nblk = (nc + 17) / 18;
        ttest_residue_chains<18, false>(sig, score, nc, nblk, (nblk + FP_THREADS / 18 - 1) / (FP_THREADS / 18));
    } else {
        for (int pos = tid; pos < nc; pos += FP_THREADS) {
            double m1 = 0.0, m2 = 0.0, var1 = 0.0, var2 = 0.0;
            for (int i = 0; i < w; i++) m1 = __dadd_rn(m1, (double)sig[pos + i]);
            m1 = __ddiv_rn(m1, wd);
            for (int i = 0; i < w; i++) m2 = __dadd_rn(m2, (double)sig[pos + w + i]);
            m2 = __ddiv_rn(m2, wd);
            for (int i = 0; i < w; i++) {
                const double pd = __dsub_rn((double)sig[pos + i], m1);
                var1 = __dadd_rn(var1, __dmul_rn(pd, pd));
            }
            for (int i = 0; i < w; i++) {
                const double pd = __dsub_rn((double)sig[pos + w + i], m2);
                var2 = __dadd_rn(var2, __dmul_rn(pd, pd));
            }
            const double vs = __dadd_rn(var1, var2);
            double sc;
            if (vs == 0.0) sc = 0.0;
            else if (m1 > m2) sc = __ddiv_rn(__dsub_rn(m1, m2), __dsqrt_rn(vs));
            else sc = __ddiv_rn(__dsub_rn(m2, m1), __dsqrt_rn(vs));
            score[pos] = sc;
        }
    }
    __syncthreads();


    FP_T(s, 3);   // t-test
    // ---- change points: find_peaks + the num_events highest scores (sig_proc.py:176-198) ------------
    const int P = kept_peaks_select(score, 0, nc, m_obs, kp, bm, cap, c.num_events, w /* + running_stat_width */, cpts + 1, state,
                                    reinterpret_cast<unsigned long long*>(dv), s);
    FP_T(s, 8);   // kept peaks + top-k
    if (P < c.num_events) {  // sig_proc.py:185-186 -> "event segmentation failed"
        fail(FP_FAIL_SEGMENTATION);
        return;
    }
    if (trimmed) {  // the last event would reach into the NaN padding (sig_proc.py:553-560)
        fail(FP_FAIL_NORMALIZE);
        return;
    }
    if (tid == 0) {
        cpts[0] = 0;                      // peaks lie in [1, nc-2] and w >= 1: 0 and n are never present
        cpts[c.num_events + 1] = n;
    }
    __syncthreads();
    FP_T(s, 9);   // top-k
    const int n_seg = c.num_events + 1;

    // ---- c_new_means (_c_segmentation.pyx:41-53): sequential float64 sums; one warp per segment when the sums are exact
    fp_segment_means(sig, n_seg, cpts, ev, exact_sums);
    __syncthreads();

    FP_T(s, 10);  // means
    // ---- mean_normalize (sig_proc.py:99-111) with numpy's summation order -----------
    if (tid < 32) {   // warp 0: the eight accumulators of numpy's block sum live in lanes 0..7
#ifdef WDX_FP_EXPERIMENT_NONORM   // timing experiment only (wrong results): what the two sums cost
        const double mean = ev[0];
        const double ss = ev[1];
#else
        const double mean = __ddiv_rn(np_pairwise_sum_warp(n_seg, [&](int i) { return ev[i]; }), (double)n_seg);
        FP_T(s, 13);
        const double ss = np_pairwise_sum_warp(n_seg, [&](int i) {
            const double d = __dsub_rn(ev[i], mean);
            return __dmul_rn(d, d);
        });
#endif
        FP_T(s, 14);
        if (tid == 0) {
            red[0] = mean;
            red[1] = __dsqrt_rn(__ddiv_rn(ss, (double)n_seg));
        }
        FP_T(s, 15);
    }
    __syncthreads();
    const double ev_mean = red[0], ev_std = red[1];

    FP_T(s, 11);  // normalize
    // ---- statistics (sig_proc.py:562-567 / 490-498: always over the ADAPTER events) ---------------
    if (a.stats) {
        if (tid < n_seg) dv[tid] = (double)(cpts[tid + 1] - cpts[tid]);
        small_median(dv, n_seg, &red[2], &red[6]);  // adapter_dt_med
        const double dt_med = red[2];
        if (tid < n_seg) dv[tid] = fabs(__dsub_rn((double)(cpts[tid + 1] - cpts[tid]), dt_med));
        small_median(dv, n_seg, &red[3], &red[6]);  // adapter_dt_mad
        small_median(ev, n_seg, &red[4], &red[6]);  // adapter_event_med
        const double e_med = red[4];
        if (tid < n_seg) dv[tid] = fabs(__dsub_rn(ev[tid], e_med));
        small_median(dv, n_seg, &red[5], &red[6]);  // adapter_event_mad
    }
    auto write_stats = [&]() {
        if (a.stats && tid == 0) {
            double* st = a.stats + read * 6;
            st[0] = red[2];
            st[1] = red[3];
            st[2] = ev_mean;
            st[3] = ev_std;
            st[4] = red[4];
            st[5] = red[5];
        }
    };

    if constexpr (CONS) {
        // ---- consensus-guided barcode refinement (sig_proc.py:257-378) ------------------------------
        // The second segmentation hands compute_base_means change points up to scores.size + 2 *
        // running_stat_width; with a narrower adapter window that lies beyond the signal (the
        // reference's bounds-checked Cython raises): such reads fail.
        if (w != c.running_stat_width) {
            fail(FP_FAIL_SEGMENTATION);
            return;
        }
        __syncthreads();
        if (tid < n_seg) dv[tid] = __ddiv_rn(__dsub_rn(ev[tid], ev_mean), ev_std);  // normalize(series, "mean")
        __syncthreads();
        write_stats();   // (a read that fails further on overwrites them with NaN, like every failed read)
        if constexpr (PHASE == 1) {
            unsigned char* pk_ = a.park + (size_t)read * fp_park_bytes(cap);
            FpParkHead* hd = park_head(pk_);
            if (tid == 0) {
                hd->n = n;
                hd->nc = nc;
                hd->w = w;
                hd->n_seg = n_seg;
                hd->exact_sums = exact_sums ? 1 : 0;
                hd->ev_mean = ev_mean;
                hd->ev_std = ev_std;
                a.status[read] = FP_PENDING;
            }
            double* ps = park_score(pk_);
            for (int i = tid; i < nc; i += FP_THREADS) ps[i] = score[i];
            float* pg = park_sig(pk_, cap);
            for (int i = tid; i < n; i += FP_THREADS) pg[i] = sig[i];
            double* pr = park_series(pk_, cap);
            int* pc = park_cpts(pk_, cap);
            for (int i = tid; i < n_seg; i += FP_THREADS) pr[i] = dv[i];
            for (int i = tid; i <= n_seg; i += FP_THREADS) pc[i] = cpts[i];
            return;
        } else {
            __shared__ double lastrow[FP_MAX_EVENTS + 2];
            __shared__ int lastorg[FP_MAX_EVENTS + 2];
            __shared__ int match[2];
            __shared__ double hand_v[2 * (FP_MAX_QUERY / 32)];
            __shared__ int hand_o[2 * (FP_MAX_QUERY / 32)];
            if (tid < FP_MAX_QUERY)   // FP_MAX_QUERY / 32 warps, one query row per lane
                consensus_match_rows<FP_MAX_QUERY / 32>(a.cons_query, c.cons_len, dv, n_seg, c.cons_pen2, c.cons_psi_q, c.cons_psi_s,
                                                        lastrow, lastorg, match, hand_v, hand_o);
            __syncthreads();
            cons_tail(c, a, s, read, cap, score, sig, kp, state, bm, cpts, ev, dv, n, nc, w, exact_sums, ev_mean, ev_std, match[0], match[1]);
            return;
        }
    }
    write_stats();
    fp_write_fingerprint(a, read, nb, n_seg, cpts, ev, ev_mean, ev_std);
    FP_T(s, 12);  // stats + output
    FP_T_END(s);
    if (tid == 0) a.status[read] = FP_OK;
}

// ---- three-launch consensus form, second kernel: the sub-sequence alignment alone.  One small CTA per read (one query row
// per lane over FP_MAX_QUERY / 32 warps), so that an SM holds many alignments at once: the 200 dependent anti-diagonal
// steps of one read are latency, not work.
template <int NW>   // warps = ceil(query length / 32)
__global__ void __launch_bounds__(NW * 32) fingerprint_cons_align_kernel(const __grid_constant__ FpConfig c, const __grid_constant__ FpArgs a) {
    __shared__ double series[FP_MAX_EVENTS + 2];
    __shared__ double lastrow[FP_MAX_EVENTS + 2];
    __shared__ int lastorg[FP_MAX_EVENTS + 2];
    __shared__ int match[2];
    __shared__ double hand_v[2 * NW];
    __shared__ int hand_o[2 * NW];
    const int64_t read = blockIdx.x;
    if (read >= a.n || a.status[read] != FP_PENDING) return;
    unsigned char* pk_ = a.park + (size_t)read * fp_park_bytes(a.cap);
    FpParkHead* hd = park_head(pk_);
    const int n_seg = hd->n_seg;
    const double* pr = park_series(pk_, a.cap);
    for (int i = threadIdx.x; i < n_seg; i += NW * 32) series[i] = pr[i];
    __syncthreads();
    consensus_match_rows<NW>(a.cons_query, c.cons_len, series, n_seg, c.cons_pen2, c.cons_psi_q, c.cons_psi_s, lastrow, lastorg,
                                            match, hand_v, hand_o);
    __syncthreads();
    if (threadIdx.x < 2) hd->match[threadIdx.x] = match[threadIdx.x];
}

// ---- third kernel: the refinement behind the alignment on the parked state (same CTA shape and shared-memory layout as
// fingerprint_kernel).
__global__ void __launch_bounds__(FP_THREADS, WDX_FP_MIN_CTAS) fingerprint_cons_tail_kernel(const __grid_constant__ FpConfig c,
                                                                                             const __grid_constant__ FpArgs a) {
    extern __shared__ __align__(16) unsigned char fp_smem[];
    const int cap = a.cap;
    double* score = reinterpret_cast<double*>(fp_smem);
    float* sig = reinterpret_cast<float*>(fp_smem + (size_t)cap * 8);
    uint16_t* kp = reinterpret_cast<uint16_t*>(fp_smem + (size_t)cap * 12 + 32);
    uint8_t* state = fp_smem + (size_t)cap * 12 + 32 + ((size_t)(cap / 2 + 8) * 2);
    uint32_t* bm = reinterpret_cast<uint32_t*>(fp_smem + (size_t)cap * 12 + 32 + ((size_t)(cap / 2 + 8) * 3));
    __shared__ FpScratch s;
    __shared__ int cpts[FP_MAX_EVENTS + 2];
    __shared__ double ev[FP_MAX_EVENTS + 2];
    __shared__ double dv[FP_MAX_EVENTS + 2];
    const int tid = threadIdx.x;
    const int64_t read = blockIdx.x;
    if (read >= a.n || a.status[read] != FP_PENDING) return;
    unsigned char* pk_ = a.park + (size_t)read * fp_park_bytes(cap);
    const FpParkHead hd = *park_head(pk_);
    const int* pc = park_cpts(pk_, cap);
    for (int i = tid; i <= hd.n_seg; i += FP_THREADS) cpts[i] = pc[i];
    // only what lies behind the barcode start is looked at again (scores from sbs - 1: the left neighbour of the first candidate)
    const int q_end_ = min(max(hd.match[1], 0), hd.n_seg);
    const int from = max(0, min(pc[q_end_], hd.nc) - 1);
    const double* ps = park_score(pk_);
    for (int i = from + tid; i < hd.nc; i += FP_THREADS) score[i] = ps[i];
    const float* pg = park_sig(pk_, cap);
    for (int i = from + tid; i < hd.n; i += FP_THREADS) sig[i] = pg[i];
    __syncthreads();
    cons_tail(c, a, s, read, cap, score, sig, kp, state, bm, cpts, ev, dv, hd.n, hd.nc, hd.w, hd.exact_sums != 0, hd.ev_mean, hd.ev_std,
              hd.match[0], hd.match[1]);
}

// Longest adapter slice of a batch whose bounds live in device memory (sizes the shared memory).
__global__ void max_slice_kernel(const int64_t* __restrict__ a0, const int64_t* __restrict__ a1, const int32_t* __restrict__ sig_len,
                                 int64_t n, int64_t stride, int padding, int* __restrict__ out) {
    int best = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t len_row = sig_len ? min((int64_t)sig_len[r], stride) : stride;
        int64_t b = a0[r] - padding, e = a1[r] + padding;
        if (b < 0) b = 0;
        if (e > len_row) e = len_row;
        if (e - b > best) best = (int)min((int64_t)0x7fffffff, e - b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(out, best);
}

inline size_t fingerprint_smem_bytes(int cap) {
    return (size_t)cap * 12 + 32 + (size_t)(cap / 2 + 8) * 3 + (size_t)FP_BM_WORDS(cap) * 12;
}

}  // namespace wdx
