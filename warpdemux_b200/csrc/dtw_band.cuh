// dtw_band.cuh — the Sakoe-Chiba-banded DTW recurrence, one (read, support
// vector) pair per thread, the whole DP row in registers.
//
// Arithmetic restated from dtaidistance 2.3.13 dtw_distance (the library the
// reference calls at warpdemux/parallel_distances.py:59-66):
//   D[i][j] = (a[i]-s[j])^2 + min(D[i-1][j-1], D[i-1][j]+p2, D[i][j-1]+p2),  p2 = penalty^2
//   band: max(0,i-w+1) <= j < min(L,i+w);  result sqrt(D[L-1][L-1]).
//
// L and the window are template parameters, so after full unrolling every
// band-edge test folds away and the straight-line code touches only registers:
//   FAST  (float):  FADD, FFMA, FADD, FMNMX3            = 4 issue slots / cell
//   EXACT (double): DADD, DMUL, DADD, DADD + 2 x (DSETP + 2 FSEL); no FMA, so
//                   every cell rounds exactly like the x86-64 CPU build.
#pragma once
#include <cuda_runtime.h>

namespace wdx {

template <int L, int W>
struct Band {
    static constexpr int w = (W <= 0 || W > L) ? L : W;
    __host__ __device__ static constexpr int jlo(int i) { return (i - w + 1 > 0) ? (i - w + 1) : 0; }
    __host__ __device__ static constexpr int jhi(int i) { return (i + w < L) ? (i + w) : L; }
    __host__ __device__ static constexpr int cells() {
        int c = 0;
        for (int i = 0; i < L; i++) c += jhi(i) - jlo(i);
        return c;
    }
};

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // FMNMX3 (sm_100+)
    return d;
}

// 3-input minimum, selectable implementation (chosen by measurement, see
// DESIGN.md "min3"): 0 = FMNMX3 (float), 1 = VIMNMX3 on the bit patterns (all DP
// values are non-negative floats, for which integer order == float order; a NaN
// pattern compares above +inf and is ignored exactly like fminf ignores it),
// 2 = two FMNMX, 3 = two VIMNMX.
template <int MI>
__device__ __forceinline__ float min3sel(float a, float b, float c) {
    if constexpr (MI == 0) {
        return fmin3(a, b, c);
    } else if constexpr (MI == 1) {
        return __int_as_float(min(min(__float_as_int(a), __float_as_int(b)), __float_as_int(c)));
    } else if constexpr (MI == 2) {
        float t, d;  // volatile: keep ptxas from re-fusing the pair into FMNMX3
        asm volatile("min.f32 %0, %1, %2;" : "=f"(t) : "f"(b), "f"(c));
        asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(t));
        return d;
    } else {
        int t, d;
        asm("min.s32 %0, %1, %2;" : "=r"(t) : "r"(__float_as_int(b)), "r"(__float_as_int(c)));
        asm("min.s32 %0, %1, %2;" : "=r"(d) : "r"(__float_as_int(a)), "r"(t));
        return __int_as_float(d);
    }
}
template <int MI>
__device__ __forceinline__ float min2sel(float a, float b) {
    if constexpr (MI == 1 || MI == 3) return __int_as_float(min(__float_as_int(a), __float_as_int(b)));
    else return fminf(a, b);
}

// ---- FAST: float32, FMA, v and v+p2 both kept so each cell is 4 instructions.
// Returns D[L-1][L-1] (squared-cost sum; caller takes the sqrt).
template <int L, int W, int MI = 0>
__device__ __forceinline__ float dtw_band_f32(const float (&a)[L], const float (&s)[L], const float p2) {
    using B = Band<L, W>;
    float v[L];   // D[i][j]       of the row being overwritten in place
    float vp[L];  // D[i][j] + p2
#pragma unroll
    for (int i = 0; i < L; i++) {
        const int jlo = B::jlo(i), jhi = B::jhi(i);
        const int prev_hi = (i > 0) ? B::jhi(i - 1) : 0;
        float diag = 0.f;  // D[i-1][j-1] carried across the in-place update
        if (i > 0 && jlo > 0) diag = v[jlo - 1];
#pragma unroll
        for (int j = 0; j < L; j++) {
            if (j >= jlo && j < jhi) {
                const float diff = a[i] - s[j];
                const bool has_diag = (i == 0) ? (j == 0) : (j > 0);
                const bool has_up = (i > 0) && (j < prev_hi);
                const bool has_left = (j > jlo);
                float old = 0.f;
                if (has_up) old = v[j];
                float m;
                if (has_diag && has_up && has_left) m = min3sel<MI>(diag, vp[j], vp[j - 1]);
                else if (has_diag && has_up) m = min2sel<MI>(diag, vp[j]);
                else if (has_diag && has_left) m = min2sel<MI>(diag, vp[j - 1]);
                else if (has_up && has_left) m = min2sel<MI>(vp[j], vp[j - 1]);
                else if (has_diag) m = diag;
                else if (has_up) m = vp[j];
                else m = vp[j - 1];
                v[j] = __fmaf_rn(diff, diff, m);
                vp[j] = v[j] + p2;
                diag = old;
            }
        }
    }
    return v[L - 1];
}

// ---- EXACT: float64, explicit round-to-nearest mul/add (never contracted).
// min(up+p2, left+p2) == min(up,left)+p2 bit-for-bit (rounding is monotone),
// which saves keeping a second row.  Strict-'<' update order of the reference
// only matters for ties, and ties are value-equal.
//
// The minimum is selectable (chosen by measurement, scripts/ubench_dtw64.cu):
//   0  (b < a) ? b : a on doubles      -> DSETP (FP64 pipe) + 2 SEL
//   1  the same comparison on the 64-bit patterns as unsigned integers -> 2 ISETP + 2 SEL on the
//      ALU pipe, leaving the FP64 pipe to the four arithmetic ops of the cell.  All DP values are
//      non-negative (sums of squares, +0, +inf), for which unsigned order == numeric order; a NaN
//      pattern compares above +inf, and a row that contains one NaN is NaN throughout (every
//      cell's cost term is NaN), exactly as with the floating-point comparison.
//   2  fmin()
// Measured (profiles/r01_ubench_instruction_mix.txt): 0 is fastest; DSETP + 2 SEL cost 5.3 cycles per
// minimum, the four arithmetic ops 2 cycles each: 18.6 cycles per cell is the FP64 ceiling of this cell.
template <int MI>
__device__ __forceinline__ double dmin_(double a, double b) {
    if constexpr (MI == 1) {
        const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
        return __longlong_as_double((long long)((ub < ua) ? ub : ua));
    } else if constexpr (MI == 2) {
        return fmin(a, b);
    } else {
        return (b < a) ? b : a;
    }
}

template <int L, int W, int MI = 0>
__device__ __forceinline__ double dtw_band_f64(const double (&a)[L], const double (&s)[L], const double p2) {
    using B = Band<L, W>;
    double v[L];
#pragma unroll
    for (int i = 0; i < L; i++) {
        const int jlo = B::jlo(i), jhi = B::jhi(i);
        const int prev_hi = (i > 0) ? B::jhi(i - 1) : 0;
        double diag = 0.0;
        if (i > 0 && jlo > 0) diag = v[jlo - 1];
#pragma unroll
        for (int j = 0; j < L; j++) {
            if (j >= jlo && j < jhi) {
                const double diff = __dsub_rn(a[i], s[j]);
                const double d = __dmul_rn(diff, diff);
                const bool has_diag = (i == 0) ? (j == 0) : (j > 0);
                const bool has_up = (i > 0) && (j < prev_hi);
                const bool has_left = (j > jlo);
                double old = 0.0;
                if (has_up) old = v[j];
                double m;
                if (has_up && has_left) m = __dadd_rn(dmin_<MI>(v[j], v[j - 1]), p2);
                else if (has_up) m = __dadd_rn(v[j], p2);
                else if (has_left) m = __dadd_rn(v[j - 1], p2);
                if (has_diag && (has_up || has_left)) m = dmin_<MI>(diag, m);
                else if (has_diag) m = diag;
                v[j] = __dadd_rn(d, m);
                diag = old;
            }
        }
    }
    return v[L - 1];
}

// ---- Generic (runtime L <= MAXL, any window): rolling rows in local memory.
// Correctness fallback for model shapes other than the shipped L=25/w=15.
template <typename T, int MAXL>
__device__ __noinline__ T dtw_generic(const T* a, const T* s, int L, int window, T p2) {
    T prev[MAXL + 1], cur[MAXL + 1];
    const T inf = (T)INFINITY;
    if (window <= 0 || window > L) window = L;
    for (int j = 0; j <= L; j++) prev[j] = inf;
    prev[0] = (T)0;
    for (int i = 0; i < L; i++) {
        int jlo = i - window + 1; if (jlo < 0) jlo = 0;
        int jhi = i + window; if (jhi > L) jhi = L;
        for (int j = 0; j <= L; j++) cur[j] = inf;
        for (int j = jlo; j < jhi; j++) {
            T diff, d, m, t;
            if constexpr (sizeof(T) == 8) {
                diff = __dsub_rn(a[i], s[j]); d = __dmul_rn(diff, diff);
                m = prev[j];
                t = __dadd_rn(prev[j + 1], p2); if (t < m) m = t;
                t = __dadd_rn(cur[j], p2); if (t < m) m = t;
                cur[j + 1] = __dadd_rn(d, m);
            } else {
                diff = a[i] - s[j];
                m = prev[j];
                t = prev[j + 1] + p2; if (t < m) m = t;
                t = cur[j] + p2; if (t < m) m = t;
                cur[j + 1] = __fmaf_rn(diff, diff, m);
            }
        }
        for (int j = 0; j <= L; j++) prev[j] = cur[j];
    }
    return prev[L];
}

}  // namespace wdx
