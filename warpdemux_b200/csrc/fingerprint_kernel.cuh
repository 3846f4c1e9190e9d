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

enum { FP_OK = 0, FP_FAIL_SEGMENTATION = 1, FP_FAIL_DETECT = 2, FP_FAIL_NORMALIZE = 3, FP_FAIL_TOO_LONG = 4, FP_FAIL_CONSENSUS = 5 };
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
};

// Same result as block_median_f32, found with two light passes instead of five
// heavy ones: values are binned linearly over [vmin, vmax] (a monotone map, so
// every element of a lower bin is <= every element of a higher bin), the bin that
// holds rank (n-1)/2 is located by a scan of the histogram, its few members are
// gathered and ranked exactly on their order keys.  Falls back to the radix
// selection when the bin is crowded (degenerate signals).
//   hist: FP_MED_BINS uint32,  cand: FP_MED_CAND uint32  (scratch in shared memory)
template <typename VAL>
__device__ float block_median_f32_linear(int n, VAL val, float vmin, float vmax, uint32_t* hist, uint32_t* cand,
                                         FpScratch& s) {
    if (!(vmax > vmin)) return vmin;  // all values equal
    const float scale = __fdiv_rn((float)FP_MED_BINS, __fsub_rn(vmax, vmin));
    auto key_of = [&](int i) { return f32_key(val(i)); };
    if (!(scale < 1e30f)) return block_median_f32(n, key_of, s);
    auto bin_of = [&](float x) { return min(FP_MED_BINS - 1, (int)__fmul_rn(__fsub_rn(x, vmin), scale)); };
    const int tid = threadIdx.x;
    const uint32_t k_lo = (uint32_t)((n - 1) / 2);
    __syncthreads();
    for (int b = tid; b < FP_MED_BINS; b += FP_THREADS) hist[b] = 0;
    if (tid == 0) s.ncand = 0;
    __syncthreads();
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
        if (bin_of(x) == sel_bin) cand[atomicAdd(&s.ncand, 1u)] = f32_key(x);
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
template <int W>
__device__ __forceinline__ void window_stat_fixed(const float* p, double wd, double wr, double& m, double& v) {
    double x[W];
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = (double)p[i];
    m = 0.0;
#pragma unroll
    for (int i = 0; i < W; i++) m = __dadd_rn(m, x[i]);
    m = div_small_int(m, wd, wr);
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
        for (int i = 8 + lane; i < n8; i += 8) r = __dadd_rn(r, at(i));
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
// peaks, in order, in kp[0..P) and returns P.  kp / state: scratch of (nc - lo)/2 + 8 entries.
__device__ int find_kept_peaks(const double* score, int lo, int nc, int m_obs, uint16_t* kp, uint8_t* state,
                               FpScratch& s) {
    const int tid = threadIdx.x;
    // scipy _local_maxima_1d (strict maxima, plateaus -> midpoint), compacted in order:
    // thread t scans the contiguous positions [lo + t*chunk, lo + (t+1)*chunk); a plateau belongs to the
    // thread that owns its first sample, which keeps the list sorted by position.
    const int chunk = (nc - lo + FP_THREADS - 1) / FP_THREADS;
    const int p_begin = min(nc - 1, max(lo + 1, lo + tid * chunk)), p_end = min(nc - 1, lo + (tid + 1) * chunk);
    auto peak_at = [&](int i) -> int {  // midpoint of the maximum that starts at i, or -1
        const double x = score[i];
        if (!(score[i - 1] < x)) return -1;
        int ahead = i + 1;
        while (ahead < nc - 1 && score[ahead] == x) ahead++;
        return (score[ahead] < x) ? ((i + ahead - 1) >> 1) : -1;
    };
    uint32_t my = 0;
    for (int i = p_begin; i < p_end; i++) my += (peak_at(i) >= 0);
    uint32_t total = 0;
    uint32_t off = block_exscan(my, s, &total);
    for (int i = p_begin; i < p_end; i++) {
        const int pk = peak_at(i);
        if (pk >= 0) {
            kp[off] = (uint16_t)pk;
            state[off] = 1;  // per-peak state: 1 undecided, 2 kept, 3 removed
            off++;
        }
    }
    __syncthreads();
    const int P0 = (int)total;

    // ---- scipy _select_by_peak_distance as a fixed point over the peak list ---------
    // A peak stays iff no STAYING peak of higher priority (score, then index) lies closer than
    // m_obs samples; the greedy highest-first sweep of scipy computes exactly this set.
    if (m_obs > 1) {
        for (;;) {
            int undecided = 0;
            for (int j = tid; j < P0; j += FP_THREADS) {
                if ((state[j] & 15) != 1) continue;
                const int pj = kp[j];
                const double x = score[pj];
                bool killed = false, blocked = false;
                for (int q = j - 1; q >= 0 && pj - (int)kp[q] < m_obs; q--) {
                    const int st = state[q] & 15;
                    if (st == 3) continue;
                    if (score[kp[q]] > x) {  // equal scores: the higher index wins, q < j loses
                        if (st == 2) killed = true;
                        else blocked = true;
                    }
                }
                for (int q = j + 1; q < P0 && (int)kp[q] - pj < m_obs; q++) {
                    const int st = state[q] & 15;
                    if (st == 3) continue;
                    if (score[kp[q]] >= x) {
                        if (st == 2) killed = true;
                        else blocked = true;
                    }
                }
                const int ns = killed ? 3 : (blocked ? 1 : 2);
                if (ns == 1) undecided = 1;
                state[j] = (uint8_t)(1 | (ns << 4));  // verdict parked in the high nibble (nobody else reads it)
            }
            const int any = __syncthreads_or(undecided);
            for (int j = tid; j < P0; j += FP_THREADS)
                if (state[j] >> 4) state[j] = state[j] >> 4;
            __syncthreads();
            if (!any) break;
        }
    } else {
        for (int j = tid; j < P0; j += FP_THREADS) state[j] = 2;
        __syncthreads();
    }

    // ---- kept peaks, in order (in place: the write index never passes the read index)
    const int pchunk0 = (P0 + FP_THREADS - 1) / FP_THREADS;
    const int j0 = min(P0, tid * pchunk0), j1 = min(P0, j0 + pchunk0);
    uint32_t mk = 0;
    for (int j = j0; j < j1; j++) mk += (state[j] == 2);
    uint32_t koff = block_exscan(mk, s, &total);
    uint16_t keep_local[FP_MAX_LEN / 2 / FP_THREADS + 2];
    int nk = 0;
    for (int j = j0; j < j1; j++)
        if (state[j] == 2) keep_local[nk++] = kp[j];
    __syncthreads();  // everybody has read its part of the list
    for (int q = 0; q < nk; q++) kp[koff + q] = keep_local[q];
    __syncthreads();
    return (int)total;
}

constexpr int FP_RANK_MAX = 192;
constexpr int FP_TOPK_CAND = FP_MAX_EVENTS + 2;   // keys of the threshold's histogram bin that can be ranked exactly (scratch = dv)   // up to this many kept peaks the top-k is found by all-pairs ranking

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
    bool have_thr = false;
    {
        // The k-th largest score in three light passes: float minimum / maximum of the peaks' scores, a 256-bin histogram
        // over that range (float(x) and the binning are monotone, so a higher bin holds only larger scores), and the
        // exact ranking of the few members of the bin that holds rank k on their 64-bit keys.  Falls through to the
        // radix selection below for degenerate score sets.
        __syncthreads();
        if (tid < 256) s.hist[tid] = 0;
        if (tid == 0) {
            s.vmin_key = 0xffffffffu;
            s.vmax_key = 0u;
            s.ncand = 0;
            s.flag = 0;
        }
        __syncthreads();
        uint32_t kmin = 0xffffffffu, kmax = 0u;
        for (int i = tid; i < P; i += FP_THREADS) {
            const uint32_t kf = f32_key((float)score[kp[i]]);
            kmin = min(kmin, kf);
            kmax = max(kmax, kf);
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
        __syncthreads();
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
                        s.flag = 1;
                    }
                }
                __syncthreads();
                have_thr = s.flag != 0;
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
    uint32_t my_ties = 0;
    for (int i = i0; i < i1; i++) my_ties += ((unsigned long long)__double_as_longlong(score[kp[i]]) == thr_key);
    uint32_t tot_ties = 0;
    uint32_t tie_off = block_exscan(my_ties, s, &tot_ties);  // ties before this thread's chunk
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

// CONS = consensus-guided barcode refinement (tRNA configurations, sig_proc.py:257-378, 451-521).
template <bool CONS>
__global__ void __launch_bounds__(FP_THREADS, 1024 / FP_THREADS) fingerprint_kernel(const __grid_constant__ FpConfig c,
                                                                   const __grid_constant__ FpArgs a) {
    extern __shared__ __align__(16) unsigned char fp_smem[];
    const int cap = a.cap;
    double* score = reinterpret_cast<double*>(fp_smem);                       // [cap]
    float* sig = reinterpret_cast<float*>(fp_smem + (size_t)cap * 8);         // [cap]
    uint16_t* kp = reinterpret_cast<uint16_t*>(fp_smem + (size_t)cap * 12);   // [cap/2 + 8] kept peak positions
    uint8_t* state = fp_smem + (size_t)cap * 12 + ((size_t)(cap / 2 + 8) * 2);  // [cap]
    __shared__ FpScratch s;
    __shared__ int cpts[FP_MAX_EVENTS + 2];
    __shared__ double ev[FP_MAX_EVENTS + 2];   // event means, later normalised
    __shared__ double dv[FP_MAX_EVENTS + 2];   // scratch for the statistics
    __shared__ double red[8];

    const int tid = threadIdx.x;
    const int64_t read = blockIdx.x;
    if (read >= a.n) return;
    if (a.retry_status && a.status[read] != a.retry_status) return;
    const int nb = c.barcode_num_events;
    double* fpt_out = a.fpt + read * nb;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    auto fail = [&](int code) {  // whole CTA calls this (uniform)
        for (int q = tid; q < nb; q += FP_THREADS) {
            fpt_out[q] = qnan;
            if (a.dwell) a.dwell[read * nb + q] = 0;
        }
        if (a.stats && tid < 6) a.stats[read * 6 + tid] = qnan;
        if (CONS && a.cons && tid < 3) a.cons[read * 3 + tid] = 0;
        if (tid == 0) a.status[read] = code;
    };

    if (a.detect_ok && !a.detect_ok[read]) {  // sig_proc.py:400-407
        fail(FP_FAIL_DETECT);
        return;
    }

    // ---- extract_adapter (sig_proc.py:382-391) --------------------------------
    const int64_t len_row = a.sig_len ? min((int64_t)a.sig_len[read], a.stride) : a.stride;
    int64_t start = a.adapter_start[read] - c.padding;
    if (start < 0) start = 0;
    int64_t stop = a.adapter_end[read] + c.padding;
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
    }
    __syncthreads();
    {   // the one HBM read of the slice, coalesced; minimum / maximum on the way
        uint32_t kmin = 0xffffffffu, kmax = 0u;
        for (int i0 = tid; i0 < n; i0 += 4 * FP_THREADS) {   // four loads in flight per thread before the first use
            float xv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * FP_THREADS;
                xv[u] = (i < n) ? __ldg(src + i) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * FP_THREADS;
                if (i >= n) break;
                const float x = xv[u];
                sig[i] = x;
                if (x != x) {
                    atomicMin(&s.first_nan, i);
                } else {
                    const uint32_t kx = f32_key(x);
                    kmin = min(kmin, kx);
                    kmax = max(kmax, kx);
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
    uint32_t* med_hist = reinterpret_cast<uint32_t*>(score);  // the score array is idle until the t-test
    uint32_t* med_cand = med_hist + FP_MED_BINS;
    const bool lin = !trimmed && cap >= (FP_MED_BINS + FP_MED_CAND) / 2;  // min/max cover exactly the slice; scratch fits
    float med, mad;
    if (lin) {
        const float vmin = f32_unkey(s.vmin_key), vmax = f32_unkey(s.vmax_key);
        med = block_median_f32_linear(n, [&](int i) { return sig[i]; }, vmin, vmax, med_hist, med_cand, s);
        const float ymax = fmaxf(__fsub_rn(vmax, med), __fsub_rn(med, vmin));  // >= every |x - med| (rounding is monotone)
        mad = block_median_f32_linear(n, [&](int i) { return fabsf(__fsub_rn(sig[i], med)); }, 0.0f, ymax, med_hist,
                                      med_cand, s);
    } else {
        med = block_median_f32(n, [&](int i) { return f32_key(sig[i]); }, s);
        mad = block_median_f32(n, [&](int i) { return f32_key(fabsf(__fsub_rn(sig[i], med))); }, s);
    }
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
    for (int i = tid; i < n; i += FP_THREADS) {
        float x = sig[i];
        x = (x < lo) ? lo : x;  // np.clip = min(max(x, lo), hi)
        x = (x > hi) ? hi : x;
        sig[i] = x;
        if (a.signals_mut) a.signals_mut[read * a.stride + start + i] = x;  // the reference clips its view in place
    }
    __syncthreads();

    // ---- c_windowed_t_test (_c_segmentation.pyx:124-161), float64, reference order
    const double wd = (double)w;
    if (w == 12) {  // the capped width (every adapter of >= 1265 samples): unrolled, window in registers
        // Positions r, r+12, r+24, ... share windows: a thread owns `seg_len` consecutive positions of one
        // residue class r and carries the second window's statistics over as the next position's first.
        const double wr = 1.0 / 12.0;
        const int chain_len = (nc + 11) / 12;
        const int seg_len = (chain_len + FP_THREADS / 12 - 1) / (FP_THREADS / 12);
        const int n_seg = (chain_len + seg_len - 1) / seg_len;  // <= FP_THREADS / 12: one item per thread
        if (tid < 12 * n_seg) {
            int pos = tid % 12 + 12 * seg_len * (tid / 12);
            if (pos < nc) {
                double m1, v1;
                window_stat_fixed<12>(sig + pos, 12.0, wr, m1, v1);
                for (int j = 0; j < seg_len && pos < nc; j++, pos += 12) {
                    double m2, v2;
                    window_stat_fixed<12>(sig + pos + 12, 12.0, wr, m2, v2);
                    score[pos] = ttest_combine(m1, v1, m2, v2);
                    m1 = m2;
                    v1 = v2;
                }
            }
        }
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


    // ---- change points: find_peaks + the num_events highest scores (sig_proc.py:176-198) ------------
    const int P = find_kept_peaks(score, 0, nc, m_obs, kp, state, s);
    if (P < c.num_events) {  // sig_proc.py:185-186 -> "event segmentation failed"
        fail(FP_FAIL_SEGMENTATION);
        return;
    }
    if (trimmed) {  // the last event would reach into the NaN padding (sig_proc.py:553-560)
        fail(FP_FAIL_NORMALIZE);
        return;
    }
    select_top_k(score, kp, state, P, c.num_events, w, cpts + 1, s, reinterpret_cast<unsigned long long*>(dv));  // + running_stat_width
    if (tid == 0) {
        cpts[0] = 0;                      // peaks lie in [1, nc-2] and w >= 1: 0 and n are never present
        cpts[c.num_events + 1] = n;
    }
    __syncthreads();
    const int n_seg = c.num_events + 1;

    // ---- c_new_means (_c_segmentation.pyx:41-53): sequential float64 sums ----------
    for (int q = tid; q < n_seg; q += FP_THREADS) {
        double sum = 0.0;
        const int b = cpts[q], e = cpts[q + 1];
        for (int i = b; i < e; i++) sum = __dadd_rn(sum, (double)sig[i]);
        ev[q] = __ddiv_rn(sum, (double)(e - b));
    }
    __syncthreads();

    // ---- mean_normalize (sig_proc.py:99-111) with numpy's summation order -----------
    if (tid < 32) {   // warp 0: the eight accumulators of numpy's block sum live in lanes 0..7
        const double mean = __ddiv_rn(np_pairwise_sum_warp(n_seg, [&](int i) { return ev[i]; }), (double)n_seg);
        const double ss = np_pairwise_sum_warp(n_seg, [&](int i) {
            const double d = __dsub_rn(ev[i], mean);
            return __dmul_rn(d, d);
        });
        if (tid == 0) {
            red[0] = mean;
            red[1] = __dsqrt_rn(__ddiv_rn(ss, (double)n_seg));
        }
    }
    __syncthreads();
    const double ev_mean = red[0], ev_std = red[1];

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

    int n_out_seg = n_seg;   // segments the fingerprint is cut from (their bounds in cpts, means in ev)
    if constexpr (CONS) {
        // ---- consensus-guided barcode refinement (sig_proc.py:257-378) ------------------------------
        // The second segmentation hands compute_base_means change points up to scores.size + 2 *
        // running_stat_width; with a narrower adapter window that lies beyond the signal (the
        // reference's bounds-checked Cython raises): such reads fail.
        if (w != c.running_stat_width) {
            fail(FP_FAIL_SEGMENTATION);
            return;
        }
        __shared__ double lastrow[FP_MAX_EVENTS + 2];
        __shared__ int lastorg[FP_MAX_EVENTS + 2];
        __shared__ int match[2];
        __syncthreads();
        if (tid < n_seg) dv[tid] = __ddiv_rn(__dsub_rn(ev[tid], ev_mean), ev_std);  // normalize(series, "mean")
        __syncthreads();
        __shared__ double hand_v[2 * (FP_MAX_QUERY / 32)];
        __shared__ int hand_o[2 * (FP_MAX_QUERY / 32)];
        if (tid < FP_MAX_QUERY)   // FP_MAX_QUERY / 32 warps, one query row per lane
            consensus_match_rows<FP_MAX_QUERY / 32>(a.cons_query, c.cons_len, dv, n_seg, c.cons_pen2, c.cons_psi_q, c.cons_psi_s,
                                                    lastrow, lastorg, match, hand_v, hand_o);
        __syncthreads();
        const int q_start = match[0], q_end = match[1];
        const int sbs = cpts[q_end];  // sig_barcode_start = sum(adapter_dwell_times[:q_end]) (sig_proc.py:334)
        __syncthreads();              // cpts / ev are rewritten below
        // second segmentation on barcode_scores = adapter_scores[sbs:] with the UNCAPPED min_obs_per_base
        // and running_stat_width (sig_proc.py:336-365)
        const int ke = c.cons_seg_events;
        const int P2 = (nc - sbs >= 3) ? find_kept_peaks(score, sbs, nc, c.min_obs_per_base, kp, state, s) : 0;
        if (P2 < ke) {
            fail(FP_FAIL_SEGMENTATION);
            return;
        }
        select_top_k(score, kp, state, P2, ke, w - sbs, cpts + 1, s, reinterpret_cast<unsigned long long*>(dv));  // relative to raw_signal[sbs:]
        if (tid == 0) {
            cpts[0] = 0;
            cpts[ke + 1] = n - sbs;   // scores.size + 2 * running_stat_width = (nc - sbs) + 2 w
        }
        __syncthreads();
        for (int q = tid; q < ke + 1; q += FP_THREADS) {  // compute_base_means(raw_signal[sbs:], cpts) (:368)
            double sum = 0.0;
            const int b = cpts[q], e = cpts[q + 1];
            for (int i = b; i < e; i++) sum = __dadd_rn(sum, (double)sig[sbs + i]);
            ev[q] = __ddiv_rn(sum, (double)(e - b));
        }
        __syncthreads();
        n_out_seg = ke + 1;
        if (a.cons && tid == 0) {
            a.cons[read * 3 + 0] = q_start;
            a.cons[read * 3 + 1] = q_end;
            a.cons[read * 3 + 2] = sbs;
        }
        if (q_start > c.cons_ub_start || q_end < c.cons_lb_end || q_end > c.cons_ub_end) {  // sig_proc.py:500-521
            for (int q = tid; q < nb; q += FP_THREADS) {
                fpt_out[q] = qnan;
                if (a.dwell) a.dwell[read * nb + q] = 0;
            }
            write_stats();
            if (tid == 0) a.status[read] = FP_FAIL_CONSENSUS;
            return;
        }
    }
    write_stats();

    // ---- keep the last barcode_num_events (sig_proc.py:569-594); normalised by the mean / std of the
    // adapter events (normalize "mean" :546-552, or normalize_wrt for the refined barcode events :482-484)
    const int keep = min(nb, n_out_seg);
    for (int q = tid; q < nb; q += FP_THREADS) {
        const int srcq = n_out_seg - keep + (q - (nb - keep));
        double v = qnan;  // front NaN padding if fewer events than asked (unreachable with accept_less_cpts=false)
        int64_t dw = 0;
        if (q >= nb - keep) {
            v = __ddiv_rn(__dsub_rn(ev[srcq], ev_mean), ev_std);
            dw = (int64_t)(cpts[srcq + 1] - cpts[srcq]);
        }
        fpt_out[q] = v;
        if (a.dwell) a.dwell[read * nb + q] = dw;
    }
    if (tid == 0) a.status[read] = FP_OK;
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
    return (size_t)cap * 12 + (size_t)(cap / 2 + 8) * 2 + (size_t)cap;
}

}  // namespace wdx
