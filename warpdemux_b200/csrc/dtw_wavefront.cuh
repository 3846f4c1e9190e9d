// dtw_wavefront.cuh — windowed DTW for series longer than a register row (L > MAXL): one WARP per
// (x, y) pair, anti-diagonal wavefront over column strips, neighbours exchanged with __shfl_sync.
//
// Same recurrence, band and minimum order as dtw_generic (dtw_band.cuh) i.e. the published algorithm
// of dtaidistance 2.3.13 `dtw_distance` (call sites warpdemux/parallel_distances.py:34,59):
//   D[i+1][j+1] = (x_i - y_j)^2 + min(D[i][j], D[i][j+1] + p^2, D[i+1][j] + p^2)   for |i - j| < window
// EXACT: float64, no contraction, bit-identical to the thread-per-pair kernels.  FAST: float32, FMA.
//
// Layout of one pair.  The columns are cut into chunks of 32*C; inside a chunk lane l owns the C
// columns [jb + l*C, jb + (l+1)*C): their y values and the running DP row live in registers.  Lane l
// works on row i at step i + l, so that at every step the 32 lanes sit on one anti-diagonal of strips:
//   * the row value x_i enters at lane 0 and moves one lane down per step (one __shfl_up_sync);
//   * the strip's left neighbour D[i+1][j0] is the last column lane l-1 produced one step earlier
//     (one __shfl_up_sync); the diagonal neighbour D[i][j0] is what the lane received the step before.
// A chunk only walks the rows whose band touches it (jb - window < i < jb + 32C + window), so a narrow
// window costs about L*(32C + 2*window) cells instead of L^2.  Between chunks the last column of the
// chunk goes through a per-warp scratch row in global memory (L values; lane 31 writes, lane 0 reads
// 32 rows at a time and feeds them out by shuffle).
#pragma once
#include <math.h>
#include <stdint.h>

#include <type_traits>

namespace wdx {

constexpr int WF_THREADS = 256;
constexpr int WF_WARPS = WF_THREADS / 32;
constexpr int WF_MAX_L = 16384;

template <typename T>
__device__ __forceinline__ T wf_cell(T a, T s, T m) {
    if constexpr (sizeof(T) == 8) {
        const T diff = __dsub_rn(a, s);
        return __dadd_rn(__dmul_rn(diff, diff), m);
    } else {
        const T diff = a - s;
        return __fmaf_rn(diff, diff, m);
    }
}

template <typename T>
__device__ __forceinline__ T wf_add(T a, T b) {
    if constexpr (sizeof(T) == 8) return __dadd_rn(a, b);
    else return __fadd_rn(a, b);
}

// One row of one strip: C cells, left to right.  EXACT keeps the reference's compare-and-replace minimum
// (diagonal, then vertical, then horizontal; a NaN never replaces) so non-finite inputs propagate
// identically; FAST uses one 3-input minimum.  CHECK = some strip of the warp straddles a band edge:
// bit c of `valid` says whether column j0 + c is inside the band (one LOP3 + select per cell).
template <typename T, int C, bool CHECK>
__device__ __forceinline__ void wf_strip(T (&row)[C], const T (&s)[C], T a, T diag, T left, T p2, uint32_t valid) {
    const T inf = (T)INFINITY;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const T up = row[c];
        T m;
        if constexpr (sizeof(T) == 8) {
            m = diag;
            T tt = __dadd_rn(up, p2);
            if (tt < m) m = tt;
            tt = __dadd_rn(left, p2);
            if (tt < m) m = tt;
        } else {
            m = fminf(fminf(diag, __fadd_rn(up, p2)), __fadd_rn(left, p2));
        }
        T v = wf_cell(a, s[c], m);
        if constexpr (CHECK) {
            if (!(valid & (1u << c))) v = inf;
        }
        diag = up;
        left = v;
        row[c] = v;
    }
}

template <bool EXACT, int C, typename OutT>
__global__ void __launch_bounds__(WF_THREADS) dtw_wavefront_kernel(const double* __restrict__ X, int64_t nX,
                                                                    const double* __restrict__ Y, int64_t nY, int L,
                                                                    int window, double p2d, OutT* __restrict__ out,
                                                                    void* __restrict__ edge_scratch) {
    using T = typename std::conditional<EXACT, double, float>::type;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = 32 * C;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * WF_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * WF_WARPS;
    const T inf = (T)INFINITY;
    const T p2 = (T)p2d;
    const int n_chunks = (L + W - 1) / W;
    T* edge[2] = {nullptr, nullptr};
    if (n_chunks > 1) {
        edge[0] = reinterpret_cast<T*>(edge_scratch) + warp * 2 * (int64_t)L;
        edge[1] = edge[0] + L;
    }
    const int64_t n_pairs = nX * nY;
    for (int64_t pair = warp; pair < n_pairs; pair += n_warps) {
        const double* __restrict__ x = X + (pair / nY) * L;
        const double* __restrict__ y = Y + (pair % nY) * L;
        T result = inf;
        for (int k = 0; k < n_chunks; k++) {
            const int jb = k * W;
            const int j0 = jb + lane * C;
            const int i_begin = max(0, jb - window + 1);
            const int i_end = min(L, jb + W + window - 1);
            const int prev_end = min(L, jb + window - 1);  // rows for which the previous chunk left an edge value
            const T* ein = edge[(k + 1) & 1];
            T* eout = edge[k & 1];
            const bool feeds_next = (k + 1 < n_chunks);
            T s[C], row[C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                s[c] = (j0 + c < L) ? (T)y[j0 + c] : (T)0;
                row[c] = inf;  // D[i_begin][j+1]: row 0 of the matrix, or right of the band of row i_begin - 1
            }
            // diagonal neighbour of the strip's first column at the lane's first row
            T prev_recv = inf;
            if (lane == 0) {
                if (k == 0) prev_recv = (T)0;                         // D[0][0]
                else if (i_begin > 0) prev_recv = ein[i_begin - 1];   // cell (i_begin-1, jb-1): upper band edge
            }
            auto fetch_a = [&](int idx) -> T { return (idx < i_end) ? (T)x[idx] : (T)0; };
            auto fetch_e = [&](int idx) -> T { return (k > 0 && idx < prev_end) ? ein[idx] : inf; };
            T a_next = fetch_a(i_begin + lane), e_next = fetch_e(i_begin + lane);
            T a_buf = (T)0, e_buf = inf, a_cur = (T)0, last = inf;
            const int n_steps = (i_end - i_begin) + 31;
            for (int t = 0; t < n_steps; t++) {
                if ((t & 31) == 0) {  // rows i_begin+t .. +31 for lane 0; the block after is already in flight
                    a_buf = a_next;
                    e_buf = e_next;
                    a_next = fetch_a(i_begin + t + 32 + lane);
                    e_next = fetch_e(i_begin + t + 32 + lane);
                }
                const T a_in = __shfl_sync(FULL, a_buf, t & 31);
                const T e_in = __shfl_sync(FULL, e_buf, t & 31);
                a_cur = __shfl_up_sync(FULL, a_cur, 1);
                T recv = __shfl_up_sync(FULL, last, 1);
                if (lane == 0) {
                    a_cur = a_in;
                    recv = e_in;
                }
                const int i = i_begin + t - lane;
                const bool active = (i >= i_begin && i < i_end);
                // columns of this strip inside the band of row i: [c_lo, c_hi)
                const int c_lo = max(0, (i - window + 1) - j0), c_hi = min(C, min(L, i + window) - j0);
                const bool compute = active && c_lo < c_hi;
                const bool partial = compute && (c_lo > 0 || c_hi < C);
                if (__any_sync(FULL, partial)) {  // warp-uniform: no lane pays for both variants
                    if (compute) {
                        const uint32_t valid = (c_hi >= 32 ? 0xffffffffu : ((1u << c_hi) - 1u)) & ~((1u << c_lo) - 1u);
                        wf_strip<T, C, true>(row, s, a_cur, prev_recv, recv, p2, valid);
                    }
                } else if (compute) {
                    wf_strip<T, C, false>(row, s, a_cur, prev_recv, recv, p2, 0u);
                }
                if (active) {
                    if (c_hi <= 0) {
                        // strip right of the band: untouched (+inf since the chunk began)
                    } else if (c_lo >= C) {
                        row[C - 1] = inf;  // strip left of the band: only its last column is ever read again
                    }
                    last = row[C - 1];
                    prev_recv = recv;
                    if (feeds_next && lane == 31) eout[i] = last;
                    if (i == L - 1 && !feeds_next) {
                        const int cf = (L - 1 - jb) - lane * C;
#pragma unroll
                        for (int c = 0; c < C; c++)
                            if (c == cf) result = row[c];
                    }
                }
            }
            __syncwarp();  // lane 31's edge stores are visible to the loads of the next chunk
        }
        const int lf = ((L - 1) % W) / C;
        result = __shfl_sync(FULL, result, lf);
        if (lane == 0) {
            if constexpr (EXACT) out[pair] = (OutT)__dsqrt_rn(result);
            else out[pair] = (OutT)sqrtf(result);
        }
    }
}

}  // namespace wdx
