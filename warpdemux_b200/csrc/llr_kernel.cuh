// llr_kernel.cuh — the LLR fallback of the boundary detection, one CTA per read that failed validation (sm_100a).
//
// Restates, for the reads whose CNN boundaries fail validate_boundaries, the re-detection branch of the reference's
// combined_detect_cnn (warpdemux/adapted/adapted/detect/combined.py:222-296):
//   normalize_signal(row[:min(max_obs_trace, full_len)], with_nan=True)   detect/normalize.py:15-63 (float32)
//   stage 0 "hail mary" (combined.py:232-274): poly(A) end re-detected on the CNN's [adapter_end, polya_end) stretch
//   stage 1 full LLR    (combined.py:275-290 -> detect_llr_on_downscaled_signal, combined.py:39-129)
// with
//   downscale_single_read_excl_nan         combined.py:132-142, downscale.py:4-41 (zero-padded block means, numpy order)
//   c_llr_trace / _gains / var_c           detect/_c_llr.pyx:24-38, 66-88, 184-230 (sequential float64 cumulative sums)
//   LLRTrace._trace_start_end              detect/llr.py:113-121
//   find_peaks_in_trace, adapter_end_from_trace, correct_for_plateau, correct_for_split_peak   llr.py:124-240
//   detect_full_polya_trace_peak_with_spike                                                    llr.py:385-455
// and the pieces of scipy those call: signal.find_peaks with distance / prominence / width / rel_height
// (_local_maxima_1d, _select_by_peak_distance, _peak_prominences, _peak_widths), np.nanstd, np.nan_to_num,
// stats.linregress' r value.
//
// The kernel only PROPOSES boundaries (preds_out, todo_list); they are validated by validate_kernel in list mode, and
// wdx_validate_run_ex commits the results the way the reference assigns `validated`.  Which reads each launch looks at
// comes from device-side lists filled by the launch before it (failed reads of the validation -> stage 0 -> reads with
// a proposal / without one -> ...), so a minibatch with 5 % fallback reads costs 5 % of the work and no host round trip.
//
// Arithmetic: float64 without contraction in the reference's operation order, so everything except `log` is
// bit-identical to the CPU; `log` is CUDA's (<= 1 ulp; the reference's is libm's, itself machine-dependent).  The
// outputs are integer positions.  Ties between equal peak heights in the distance rule are resolved towards the later
// peak (numpy's argsort leaves them unspecified).
//
// Included after validate_kernel.cuh (same translation unit, same CTA shape FP_THREADS).
#pragma once
#include "validate_kernel.cuh"

namespace wdx {

enum { VAL_MAD_ZERO = 10, VAL_LLR_ERROR = 11 };
enum { LLR_SRC_CNN = 0, LLR_SRC_HAIL_MARY = 1, LLR_SRC_LLR = 2, LLR_BIT_HM_RAN = 4, LLR_BIT_LLR_RAN = 8 };

struct LlrCfg {
    int max_obs_trace, min_obs_adapter, max_obs_adapter, factor;
    double outlier_thresh, peak_prominence, peak_rel_height;
    int peak_width;   // adapter_peak_width // downscale_factor
    int fallback_to_llr, fallback_short_reads;
};

struct LlrArgs {
    const float* signals;      // [n][stride]
    int64_t stride;
    const int32_t* full_len;   // [n]
    const int64_t* cnn_preds;  // [n][ld] the CNN's boundaries
    int ld;
    int64_t n;
    int32_t* info;             // [n][4]; [0] fail code (MAD_ZERO is written here), [3] source / progress bits
    int64_t* preds_out;        // [n][ld] boundaries proposed for the next validation
    float* medmad;             // [n][2] median / MAD of the trace part of the row (stage 0 writes, stage 1 reads)
    int stage;                 // 0 hail mary, 1 full LLR
    int nmax;                  // capacity (downscaled samples) of the shared float64 arrays
    int lt_max;                // capacity (samples) of the shared row
    // device-side work lists (no host round trip, no scan over the minibatch): the reads to look at, the reads with a
    // proposal (validated next), and - stage 0 only - the reads without one, which go straight on to the full LLR stage
    const int* in_list;
    const int* in_count;
    int* todo_list;
    int* todo_count;
    int* pass_list;            // or nullptr
    int* pass_count;
    unsigned long long* next;  // cursor into in_list
};

struct LlrPeaks {       // work arrays of find_peaks, capacity nmax / 2 + 2 each
    int* pk;
    int* lbase;
    int* rbase;
    int* order;
    double* prom;
    uint8_t* keep;
};

struct LlrShared {
    int count, first[2];
    int start, end, found;
    double dres[2];
};

// ---- numpy's pairwise sum of a float64 array of a few hundred elements, in parallel (np.nanstd of the LLR trace) --------
// np.add.reduce halves the array recursively (left half rounded down to a multiple of 8) until a block has <= 128 elements:
// thread 0 lists the leaf blocks in order, one WARP sums each leaf (the eight block accumulators in eight lanes), thread 0
// adds the leaf sums up the same tree.  Same additions in the same order as the serial np_pairwise<double>.
constexpr int LLR_MAX_LEAVES = 64;
struct LlrTree {
    int n_leaves;
    int off[LLR_MAX_LEAVES];
    short len[LLR_MAX_LEAVES];
    double sum[LLR_MAX_LEAVES];
    double result;
};
__device__ void llr_tree_build(int n, LlrTree& t) {   // one thread
    int lo_s[24], n_s[24], sp = 0, nl = 0;
    lo_s[sp] = 0;
    n_s[sp++] = n;
    while (sp > 0) {
        const int lo = lo_s[--sp], m = n_s[sp];
        if (m <= 128) {
            if (nl < LLR_MAX_LEAVES) {
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
__device__ double llr_tree_combine(int n, const LlrTree& t) {   // one thread; same walk, leaf sums consumed in order
    struct Frame {
        int n, state;
        double left;
    };
    Frame st[24];
    int sp = 0, next = 0;
    st[sp++] = Frame{n, 0, 0.0};
    double result = 0.0;
    while (sp > 0) {
        Frame& f = st[sp - 1];
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.n <= 128) {
            result = t.sum[next++];
            sp--;
        } else if (f.state == 0) {
            f.state = 1;
            st[sp++] = Frame{n2, 0, 0.0};
        } else if (f.state == 1) {
            f.left = result;
            f.state = 2;
            st[sp++] = Frame{f.n - n2, 0, 0.0};
        } else {
            result = __dadd_rn(f.left, result);
            sp--;
        }
    }
    return result;
}
// np_pairwise_leaf<double> (n <= 128) by one warp: lane j < 8 owns accumulator j; every lane returns the sum
template <typename F>
__device__ __forceinline__ double warp_np_leaf_f64(int lo, int n, F at) {
    const int lane = threadIdx.x & 31;
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r = __dadd_rn(r, at(lo + i));
        return r;
    }
    const int n8 = n - (n % 8);
    double r = 0.0;
    if (lane < 8) {
        r = at(lo + lane);
        for (int i = 8 + lane; i < n8; i += 8) r = __dadd_rn(r, at(lo + i));
    }
    const unsigned full = 0xffffffffu;
    const double r0 = __shfl_sync(full, r, 0), r1 = __shfl_sync(full, r, 1), r2 = __shfl_sync(full, r, 2), r3 = __shfl_sync(full, r, 3),
                 r4 = __shfl_sync(full, r, 4), r5 = __shfl_sync(full, r, 5), r6 = __shfl_sync(full, r, 6), r7 = __shfl_sync(full, r, 7);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (int i = n8; i < n; i++) res = __dadd_rn(res, at(lo + i));
    return res;
}
// Sum of f(lo + i), i < n, in numpy's order; the tree for this n must have been built (llr_tree_build + barrier).
template <typename F>
__device__ double block_np_sum_f64(int lo, int n, F f, LlrTree& t) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < t.n_leaves; k += FP_WARPS) {
        const double sum = warp_np_leaf_f64(lo + t.off[k], (int)t.len[k], f);
        if (lane == 0) t.sum[k] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) t.result = llr_tree_combine(n, t);
    __syncthreads();
    const double r = t.result;
    __syncthreads();
    return r;
}

__device__ __forceinline__ double llr_var_c(int start, int end, const double* c, const double* c2) {   // _c_llr.pyx:24-38
    if (start == end) return 0.0;
    if (start == 0) {
        const double m = __ddiv_rn(c[end - 1], (double)end);
        return __dsub_rn(__ddiv_rn(c2[end - 1], (double)end), __dmul_rn(m, m));
    }
    const double len = (double)(end - start);
    const double m = __ddiv_rn(__dsub_rn(c[end - 1], c[start - 1]), len);
    return __dsub_rn(__ddiv_rn(__dsub_rn(c2[end - 1], c2[start - 1]), len), __dmul_rn(m, m));
}

// _gains(start, end, c, c2, offset_head, offset_tail, stride = 1) over an array of n entries (zeros elsewhere).
__device__ __noinline__ void llr_gains(const double* c, const double* c2, int n, int start, int end, int head, int tail, double* g) {
    const double vs = __dmul_rn((double)(end - start), log(llr_var_c(start, end, c, c2)));
    for (int i = threadIdx.x; i < n; i += FP_THREADS) {
        double v = 0.0;
        if (i >= start + head && i < end - tail) {
            const double h = __dmul_rn((double)(i - start), log(llr_var_c(start, i, c, c2)));
            const double t = __dmul_rn((double)(end - i), log(llr_var_c(i, end, c, c2)));
            v = __dsub_rn(vs, __dadd_rn(h, t));
        }
        g[i] = v;
    }
    __syncthreads();
}

// scipy.signal.find_peaks(x[0:n], distance (0 = None), prominence = pmin, width = wmin, rel_height): number of peaks,
// the first two in sh.first.  All threads call; results are uniform.
__device__ __noinline__ int llr_find_peaks(const double* x, int n, int distance, double pmin, double wmin, double rel_height, LlrPeaks& w,
                              LlrShared& sh, FpScratch& s) {
    const int tid = threadIdx.x;
    __syncthreads();
    // _local_maxima_1d: a rising edge followed by a plateau that ends in a falling edge; the plateau's midpoint
    const int span = max(0, n - 2);                        // candidate left edges p in [1, n - 1)
    const int per = (span + FP_THREADS - 1) / FP_THREADS;
    const int p_lo = 1 + tid * per, p_hi = min(n - 1, p_lo + per);
    auto peak_at = [&](int p) -> int {
        if (!(x[p - 1] < x[p])) return -1;
        int q = p + 1;
        while (q < n - 1 && x[q] == x[p]) q++;
        return (x[q] < x[p]) ? (p + q - 1) / 2 : -1;
    };
    uint32_t cnt = 0;
    for (int p = p_lo; p < p_hi; p++) cnt += peak_at(p) >= 0;
    uint32_t total;
    uint32_t at = block_exscan(cnt, s, &total);
    for (int p = p_lo; p < p_hi; p++) {
        const int m = peak_at(p);
        if (m >= 0) w.pk[at++] = m;
    }
    __syncthreads();
    int P = (int)total;
    if (P == 0) {
        __syncthreads();
        return 0;
    }
    if (distance >= 1) {   // _select_by_peak_distance, priority = (height, then the higher index)
        // scipy walks the peaks from the highest down and removes the neighbours of every peak still kept; the kept set is
        // the unique one with "kept iff no kept peak of higher priority lies closer than `distance`".  A verdict is written
        // only when it is final (removed: a higher KEPT peak is near; kept: every higher peak near is REMOVED) and final
        // states never change, so stale reads are harmless: every warp iterates over its own undecided peaks without a
        // barrier between the rounds (before: an all-pairs ranking of the peaks and ONE thread walking them in order).
        for (int j = tid; j < P; j += FP_THREADS) w.keep[j] = 1;     // 1 undecided, 2 kept, 3 removed
        __syncthreads();
        {
            volatile uint8_t* vk = w.keep;
            bool mine_left = true;
            for (int spin = 0; __any_sync(0xffffffffu, mine_left); spin++) {
                if (spin > (1 << 22)) __trap();   // a protocol bug must end as a launch failure, never as a hung GPU
                mine_left = false;
                for (int j = tid; j < P; j += FP_THREADS) {
                    if (vk[j] != 1) continue;
                    const int pj = w.pk[j];
                    const double hj = x[pj];
                    bool killed = false, blocked = false;
                    for (int k = j - 1; k >= 0 && pj - w.pk[k] < distance; k--) {
                        const int st = vk[k];
                        if (st != 3 && x[w.pk[k]] > hj) {       // equal heights: the higher index wins, k < j loses
                            if (st == 2) killed = true;
                            else blocked = true;
                        }
                    }
                    for (int k = j + 1; k < P && w.pk[k] - pj < distance; k++) {
                        const int st = vk[k];
                        if (st != 3 && x[w.pk[k]] >= hj) {
                            if (st == 2) killed = true;
                            else blocked = true;
                        }
                    }
                    if (killed) vk[j] = 3;
                    else if (blocked) mine_left = true;
                    else vk[j] = 2;
                }
            }
        }
        __syncthreads();
        if (P <= FP_THREADS) {      // kept peaks, in order (in place: every thread has read its peak before the scan's barriers)
            const int pj = tid < P ? w.pk[tid] : 0;
            const uint32_t kf = (tid < P && w.keep[tid] == 2) ? 1u : 0u;
            uint32_t m = 0;
            const uint32_t pos = block_exscan(kf, s, &m);
            if (kf) w.pk[pos] = pj;
            if (tid == 0) sh.count = (int)m;
        } else if (tid == 0) {
            int m = 0;
            for (int j = 0; j < P; j++)
                if (w.keep[j] == 2) w.pk[m++] = w.pk[j];
            sh.count = m;
        }
        __syncthreads();
        P = sh.count;
    }
    // _peak_prominences (wlen = None), prominence filter, _peak_widths, width filter: one WARP per peak, the walks
    // away from the peak advance 32 samples per step (a thread per peak would leave the CTA waiting for the one thread
    // whose peak is the highest and walks the whole trace)
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    for (int j = warp; j < P; j += FP_WARPS) {
        const int peak = w.pk[j];
        const double xp = x[peak];
        // walk(dir): samples peak, peak+dir, ... while inside [0, n) and x <= xp; minimum of them and the position where it
        // is first reached (the one closest to the peak), as the sequential loop finds it
        double mins[2];
        int bases[2];
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const int dir = side ? 1 : -1;
            // every lane keeps the minimum of the samples IT visits (strictly smaller only: its visits move away from the
            // peak) and how far from the peak it lies; ONE reduction over (value, then distance) behind the walk gives the
            // minimum and the position closest to the peak where it is reached (lane 0's first visit is the peak itself:
            // value xp, distance 0).  Before: a five-round (value, index) reduction in every 32-sample step.
            double lv = INFINITY;
            int ld = 0x7fffffff;
            for (int base = peak;; base += 32 * dir) {
                const int idx = base + lane * dir;
                const bool inside = idx >= 0 && idx < n;
                const double v = inside ? x[idx] : 0.0;
                const bool go = inside && (v <= xp);
                const unsigned stop = __ballot_sync(full, !go);
                const int first = stop ? (__ffs(stop) - 1) : 32;       // lanes below `first` belong to the walk
                if (lane < first && v < lv) {
                    lv = v;
                    ld = (idx - peak) * dir;
                }
                if (stop) break;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(full, lv, o);
                const int od = __shfl_xor_sync(full, ld, o);
                if (ov < lv || (ov == lv && od < ld)) {
                    lv = ov;
                    ld = od;
                }
            }
            mins[side] = lv;
            bases[side] = peak + ld * dir;
        }
        const double lmin = mins[0], rmin = mins[1];
        const int lb = bases[0], rb = bases[1];
        const double prom = __dsub_rn(xp, (lmin > rmin) ? lmin : rmin);
        bool keep = pmin <= prom;
        if (keep) {
            const double height = __dsub_rn(xp, __dmul_rn(prom, rel_height));
            int ends[2];
#pragma unroll
            for (int side = 0; side < 2; side++) {   // i = peak; while (lb < i && height < x[i]) i--;   (and mirrored)
                const int dir = side ? 1 : -1;
                int found = peak;
                for (int base = peak;; base += 32 * dir) {
                    const int idx = base + lane * dir;
                    const bool inside = side ? (idx < rb) : (lb < idx);
                    const bool go = inside && (height < x[min(max(idx, 0), n - 1)]);
                    const unsigned stop = __ballot_sync(full, !go);
                    if (stop) {
                        found = base + (__ffs(stop) - 1) * dir;
                        break;
                    }
                }
                ends[side] = found;
            }
            int i = ends[0];
            double left_ip = (double)i;
            if (x[i] < height) left_ip = __dadd_rn(left_ip, __ddiv_rn(__dsub_rn(height, x[i]), __dsub_rn(x[i + 1], x[i])));
            i = ends[1];
            double right_ip = (double)i;
            if (x[i] < height) right_ip = __dsub_rn(right_ip, __ddiv_rn(__dsub_rn(height, x[i]), __dsub_rn(x[i - 1], x[i])));
            keep = wmin <= __dsub_rn(right_ip, left_ip);
        }
        if (lane == 0) w.keep[j] = keep;
    }
    __syncthreads();
    if (P <= FP_THREADS) {      // number of peaks that passed the filters and the first two of them, by a scan
        const uint32_t kf = (tid < P && w.keep[tid]) ? 1u : 0u;
        uint32_t tot = 0;
        const uint32_t pos = block_exscan(kf, s, &tot);
        if (kf && pos < 2) sh.first[pos] = w.pk[tid];
        if (tid == 0) sh.count = (int)tot;
    } else if (tid == 0) {
        int m = 0;
        for (int j = 0; j < P; j++) {
            if (w.keep[j]) {
                if (m < 2) sh.first[m] = w.pk[j];
                m++;
            }
        }
        sh.count = m;
    }
    __syncthreads();
    const int m = sh.count;
    __syncthreads();
    return m;
}

// detect_full_polya_trace_peak_with_spike(trace[0:n]) (llr.py:385-455): xn = scratch for nan_to_num(trace).
__device__ __noinline__ int llr_polya_peak(const double* tr, int n, double* xn, LlrPeaks& w, LlrShared& sh, FpScratch& s) {
    const int tid = threadIdx.x;
    const double dmax = 1.7976931348623157e308;
    for (int i = tid; i < n; i += FP_THREADS) {
        const double v = tr[i];
        xn[i] = (v != v) ? 0.0 : (v > dmax ? dmax : (v < -dmax ? -dmax : v));
    }
    __syncthreads();
    const int cnt = llr_find_peaks(xn, n, 10, 1.0, 10.0, 0.5, w, sh, s);
    if (cnt == 0) return 0;
    const int p0 = sh.first[0], p1 = sh.first[1];
    __syncthreads();
    if (cnt == 1) return p0;
    const double h0 = tr[p0], h1 = tr[p1];
    if (h1 > h0) return p1;
    if (h1 < __dmul_rn(h0, 0.5)) return p0;
    // linear regression of the trace between the minimum after the first peak and the second peak
    if (tid == 0) {
        int idx = p0;            // np.argmin: the first NaN if there is one, else the first minimum
        bool nan_seen = false;
        double mn = tr[p0];
        for (int i = p0; i < p1 && !nan_seen; i++) {
            const double v = tr[i];
            if (v != v) {
                idx = i;
                nan_seen = true;
            } else if (v < mn) {
                mn = v;
                idx = i;
            }
        }
        const int m = p1 - idx;
        double r2 = -1.0;        // "not >= threshold"
        if (m >= 2) {
            const double nn = (double)m;
            const double xm = __ddiv_rn(np_pairwise<double>(idx, m, [&](int i) { return (double)i; }), nn);
            const double ym = __ddiv_rn(np_pairwise<double>(idx, m, [&](int i) { return tr[i]; }), nn);
            double sxx = 0.0, syy = 0.0, sxy = 0.0;
            for (int i = idx; i < p1; i++) {
                const double dx = __dsub_rn((double)i, xm), dy = __dsub_rn(tr[i], ym);
                sxx = __dadd_rn(sxx, __dmul_rn(dx, dx));
                syy = __dadd_rn(syy, __dmul_rn(dy, dy));
                sxy = __dadd_rn(sxy, __dmul_rn(dx, dy));
            }
            sxx = __ddiv_rn(sxx, nn);
            syy = __ddiv_rn(syy, nn);
            sxy = __ddiv_rn(sxy, nn);
            if (!(sxx == 0.0 || syy == 0.0)) {
                double r = __ddiv_rn(sxy, sqrt(__dmul_rn(sxx, syy)));
                r = fmin(fmax(r, -1.0), 1.0);
                r2 = __dmul_rn(r, r);
            }
        }
        sh.found = (r2 >= 0.99) ? p1 : 0;
    }
    __syncthreads();
    const int res = sh.found;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(FP_THREADS, 2) llr_kernel(const LlrArgs a, const LlrCfg c) {
    extern __shared__ __align__(16) unsigned char llr_smem[];
    __shared__ FpScratch s;
    __shared__ ValSel vs;
    __shared__ LlrShared sh;
    __shared__ unsigned long long sh_next;
    const int tid = threadIdx.x;
    // dynamic shared memory: four float64 arrays of nmax, the find_peaks work arrays, the float32 row
    double* xs = reinterpret_cast<double*>(llr_smem);     // downscaled signal; later nan_to_num scratch
    double* cs = xs + a.nmax;
    double* c2 = cs + a.nmax;
    double* g = c2 + a.nmax;
    const int pcap = a.nmax / 2 + 2;
    LlrPeaks w;
    w.prom = g + a.nmax;
    w.pk = reinterpret_cast<int*>(w.prom + pcap);
    w.lbase = w.pk + pcap;
    w.rbase = w.lbase + pcap;
    w.order = w.rbase + pcap;
    w.keep = reinterpret_cast<uint8_t*>(w.order + pcap);
    float* vsig = reinterpret_cast<float*>(llr_smem + ((size_t)a.nmax * 32 + (size_t)pcap * 25 + 15) / 16 * 16);

    FP_T_BEGIN(s);
    for (;;) {
        FP_T(s, 25);   // rest of the previous read
        FP_T_END(s);
        __syncthreads();
        if (tid == 0) {
            const unsigned long long i = atomicAdd(a.next, 1ULL);
            sh_next = (i < (unsigned long long)*a.in_count) ? (unsigned long long)a.in_list[i] : (unsigned long long)a.n;
        }
        __syncthreads();
        const int64_t r = (int64_t)sh_next;
        if (r >= a.n) break;
        FP_T_BEGIN(s);
        const float* row = a.signals + (size_t)r * a.stride;
        const int64_t fl = a.full_len[r];
        const int Lt = (int)max((int64_t)0, min(min((int64_t)c.max_obs_trace, fl), min(a.stride, (int64_t)a.lt_max)));
        const int64_t* cp = a.cnn_preds + (size_t)r * a.ld;
        const int64_t ae = cp[0], pe = a.ld > 1 ? cp[1] : 0;
        int64_t* po = a.preds_out + (size_t)r * a.ld;
        if (tid < a.ld) po[tid] = 0;
        int todo = 0;

        const bool hm = a.stage == 0 && ae > 0 && pe > 0 && pe - ae > 1000 && fl < 2 * (int64_t)c.max_obs_adapter && c.fallback_short_reads;
        const bool full = a.stage == 1 && c.fallback_to_llr;
        float med, mad;
        if (a.stage == 0 || !a.medmad) {
            for (int i = tid; i < Lt; i += FP_THREADS) vsig[i] = __ldg(row + i);
            __syncthreads();
            if (Lt > 0) {   // normalize_signal: nanmedian / MAD of the NaN-free float32 slice
                med = val_median(Lt, val_src(vsig), vs, s);
                mad = val_median(Lt, val_src_absdev(vsig, med), vs, s);
            } else {
                med = mad = 1.0f;   // an empty slice normalises to an empty array (normalize.py:49-50)
            }
            if (tid == 0 && a.medmad) {
                a.medmad[r * 2] = med;
                a.medmad[r * 2 + 1] = mad;
            }
            if (mad == 0.f) {       // ValueError("MAD normalization failed: scale is 0") -> the read's result
                if (tid == 0) {
                    a.info[r * 4] = VAL_MAD_ZERO;
                    a.info[r * 4 + 1] = 0;
                }
                continue;
            }
        } else {
            med = a.medmad[r * 2];
            mad = a.medmad[r * 2 + 1];
            if (full)
                for (int i = tid; i < Lt; i += FP_THREADS) vsig[i] = __ldg(row + i);
            __syncthreads();
        }
        if (!(hm || full)) {
            if (tid == 0 && a.pass_list) a.pass_list[atomicAdd(a.pass_count, 1)] = (int)r;
            continue;
        }
        FP_T(s, 16);   // row + median / MAD
        // clip bounds: python floats (float64), cast once to float32 by np.clip; (clip - med) / mad in float32
        const double tm = __dmul_rn((double)mad, c.outlier_thresh);
        const float lo = (float)__dsub_rn((double)med, tm), hi = (float)__dadd_rn((double)med, tm);
        const int b0 = hm ? (int)min(ae, (int64_t)Lt) : 0, b1 = hm ? (int)min(pe, (int64_t)Lt) : Lt;
        const int len = max(0, b1 - b0);
        const int m = min((len + c.factor - 1) / c.factor, a.nmax);
        for (int j = tid; j < m; j += FP_THREADS) {   // zero-padded block means in numpy's pairwise order, float32
            const int base = b0 + j * c.factor;
            const float sum = np_pairwise_leaf<float>(0, c.factor, [&](int u) {
                const int i = base + u;
                if (i >= b1) return 0.0f;
                const float v = fminf(fmaxf(vsig[i], lo), hi);
                return __fdiv_rn(__fsub_rn(v, med), mad);
            });
            xs[j] = (double)__fdiv_rn(sum, (float)c.factor);
        }
        __syncthreads();
        FP_T(s, 17);   // downscaling
        // c = cumsum(x), c2 = cumsum(x * x): sequential float64 adds (two warps, one each); the loads of eight terms are
        // issued before their dependent chain of adds
        if (tid == 0 || tid == 32) {
            const bool sq = tid == 32;
            double* dst = sq ? c2 : cs;
            double acc = 0.0;
            int i = 0;
            for (; i + 8 <= m; i += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = xs[i + u];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    acc = __dadd_rn(acc, sq ? __dmul_rn(v[u], v[u]) : v[u]);
                    v[u] = acc;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) dst[i + u] = v[u];
            }
            for (; i < m; i++) {
                acc = __dadd_rn(acc, sq ? __dmul_rn(xs[i], xs[i]) : xs[i]);
                dst[i] = acc;
            }
        }
        __syncthreads();

        FP_T(s, 18);   // cumulative sums
        int bits = 0;
        if (hm) {
            bits = LLR_BIT_HM_RAN;
            int pd = 0;
            if (m >= 12) {   // fewer samples leave the trace all zero (range(5, m - 6) is empty): no peak
                llr_gains(cs, c2, m, 0, m - 1, 5, 5, g);
                pd = llr_polya_peak(g, m, xs, w, sh, s);
            }
            if (pd > 0) {
                if (tid == 0) {
                    po[0] = ae;
                    po[1] = (int64_t)pd * c.factor + ae;
                }
                todo = 1;
            }
        } else {
            bits = LLR_BIT_LLR_RAN;
            int a_ds = 0, p_ds = 0;
            if (m >= 2) {
                llr_gains(cs, c2, m, 0, m - 1, 1 + c.min_obs_adapter / c.factor, 1, g);
                // LLRTrace._trace_start_end: first / last entry that is not <= 0
                if (tid == 0) {
                    sh.start = 0x7fffffff;
                    sh.end = -1;
                }
                __syncthreads();
                int fst = 0x7fffffff, lst = -1;
                for (int i = tid; i < m; i += FP_THREADS) {
                    if (!(g[i] <= 0.0)) {
                        fst = min(fst, i);
                        lst = max(lst, i);
                    }
                }
                if (lst >= 0) {
                    atomicMin(&sh.start, fst);
                    atomicMax(&sh.end, lst);
                }
                __syncthreads();
                const int t_start = sh.end < 0 ? 0 : sh.start, t_end = sh.end < 0 ? m - 1 : sh.end;
                const int nc = t_end - t_start;          // trace.signal[start:end]
                __syncthreads();
                FP_T(s, 19);   // gains + trace start / end
                int cnt = 0;
                if (nc > 0) {
                    // np.nanstd(clip): np.var's two pairwise sums; NaN entries count as absent.  By the CTA (leaf blocks of
                    // numpy's recursion summed by warps) — one thread took ~2 x nc dependent steps here with 511 waiting.
                    __shared__ LlrTree tree;
                    int n_nan = 0;
                    for (int i0 = 0; i0 < nc; i0 += FP_THREADS) {
                        const int i = i0 + tid;
                        n_nan += __syncthreads_count(i < nc && g[t_start + i] != g[t_start + i]);
                    }
                    const double cntd = (double)(nc - n_nan);
                    if (tid == 0) llr_tree_build(nc, tree);
                    __syncthreads();
                    if (tree.n_leaves <= LLR_MAX_LEAVES) {    // uniform
                        const double mean = __ddiv_rn(block_np_sum_f64(t_start, nc, [&](int i) { return g[i] != g[i] ? 0.0 : g[i]; }, tree), cntd);
                        const double ss = block_np_sum_f64(t_start, nc, [&](int i) {
                            if (g[i] != g[i]) return 0.0;
                            const double d = __dsub_rn(g[i], mean);
                            return __dmul_rn(d, d);
                        }, tree);
                        if (tid == 0) sh.dres[0] = sqrt(__ddiv_rn(ss, cntd));
                    } else if (tid == 0) {
                        const double mean = __ddiv_rn(np_pairwise<double>(t_start, nc, [&](int i) { return g[i] != g[i] ? 0.0 : g[i]; }), cntd);
                        const double ss = np_pairwise<double>(t_start, nc, [&](int i) {
                            if (g[i] != g[i]) return 0.0;
                            const double d = __dsub_rn(g[i], mean);
                            return __dmul_rn(d, d);
                        });
                        sh.dres[0] = sqrt(__ddiv_rn(ss, cntd));
                    }
                    __syncthreads();
                    FP_T(s, 20);   // nanstd
                    const double pmin = __dmul_rn(c.peak_prominence, sh.dres[0]);
                    cnt = llr_find_peaks(g + t_start, nc, 0, pmin, (double)c.peak_width, c.peak_rel_height, w, sh, s);
                }
                FP_T(s, 21);   // find_peaks of the trace
                if (cnt > 0) {
                    int peak = sh.first[0] + t_start;
                    __syncthreads();
                    {   // correct_for_plateau(trace, peak, s = 10, t = 0.9, window = 500)
                        const int wl = min(peak + 500, m) - peak, nch = wl - 1;
                        if (tid == 0) sh.found = -1;
                        __syncthreads();
                        const double thr = __dmul_rn(0.9, g[peak]);
                        int best = -1;
                        for (int i = tid; i <= nch - 10; i += FP_THREADS) {
                            bool ok = g[peak + i + 9] > thr;
                            for (int u = 0; u < 9 && ok; u++) ok = __dsub_rn(g[peak + i + u + 1], g[peak + i + u]) >= 0.0;
                            if (ok) best = i;
                        }
                        if (best >= 0) atomicMax(&sh.found, best);
                        __syncthreads();
                        if (sh.found >= 0) peak += sh.found + 9;
                        __syncthreads();
                    }
                    FP_T(s, 22);   // plateau correction
                    {   // correct_for_split_peak(trace, peak, s = 10, t = 0.9, window = 500, prominence = 1.0)
                        const int wl = min(peak + 500, m) - peak;
                        const int c2n = llr_find_peaks(g + peak, wl, 0, 1.0, 10.0, 0.5, w, sh, s);
                        if (c2n > 0) {
                            const int q = sh.first[0] + peak;
                            if (g[q] >= __dmul_rn(0.9, g[peak])) peak = q;
                        }
                        __syncthreads();
                    }
                    FP_T(s, 23);   // split-peak correction
                    if (peak > 0) {
                        a_ds = peak;
                        llr_gains(cs, c2, m, a_ds, m - 1, 1, 1, g);
                        p_ds = llr_polya_peak(g, m, xs, w, sh, s);
                    }
                }
            }
            FP_T(s, 24);   // second gains + poly(A) peak
            if (a_ds > 0) {
                if (tid == 0) {
                    po[0] = (int64_t)a_ds * c.factor;
                    if (p_ds > 0) po[1] = (int64_t)p_ds * c.factor;
                }
                todo = 1;
            }
        }
        if (tid == 0) {
            a.info[r * 4 + 3] |= bits;
            if (todo) a.todo_list[atomicAdd(a.todo_count, 1)] = (int)r;
            else if (a.pass_list) a.pass_list[atomicAdd(a.pass_count, 1)] = (int)r;
        }
    }
}

inline size_t llr_smem_bytes(int nmax, int lt_max) {
    const size_t pcap = (size_t)nmax / 2 + 2;
    return ((size_t)nmax * 32 + pcap * 25 + 15) / 16 * 16 + (size_t)lt_max * 4;
}

}  // namespace wdx
