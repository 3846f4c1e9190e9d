// dtw_band_x2.cuh — FAST-mode DTW recurrence with packed f32x2 arithmetic.
//
// Blackwell (sm_100) has two-lane FP32 instructions (FADD2 / FFMA2: one issue
// slot, two results).  The scalar recurrence in dtw_band.cuh is issue-bound at
// 4 slots per cell (FADD, FFMA, FADD, FMNMX3); here two INDEPENDENT cells of the
// same DP are packed into every FADD2/FFMA2: rows are processed in pairs
// (r0, r0+1) with the lower row lagging one column, i.e. cells (r0, t) and
// (r0+1, t-1) — neighbours on an anti-diagonal — are computed together:
//     FADD2   diff  = (a[r0], a[r0+1]) - (s[t], s[t-1])
//     FMNMX3  m.x   = min(diag, up+p, left+p)   of cell (r0,   t)
//     FMNMX3  m.y   = ...                        of cell (r0+1, t-1)
//     FFMA2   v     = diff*diff + m
//     FADD2   v+p   = v + (p2, p2)
// = 5 issue slots per 2 cells.  The FMA pipe still performs 3 lane-operations
// per cell, which becomes the new bound.  Same arithmetic per cell as the
// scalar FAST path (identical results bit for bit).
//
// Operand layout: ap[q] = (a[2q], a[2q+1]);  sp[t] = (s[t], s[t-1]) for t >= 1.
#pragma once
#include <cuda_runtime.h>

#include "dtw_band.cuh"

namespace wdx {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float x, float y) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ float lo2(u64 v) { float x, y; unpack2(v, x, y); return x; }
__device__ __forceinline__ float hi2(u64 v) { float x, y; unpack2(v, x, y); return y; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// min over whichever of (diag, up, left) exist — resolved at compile time.
template <int MI = 0>
__device__ __forceinline__ float min_avail(bool hd, bool hu, bool hl, float d, float u, float l) {
    if (hd && hu && hl) return min3sel<MI>(d, u, l);
    if (hd && hu) return min2sel<MI>(d, u);
    if (hd && hl) return min2sel<MI>(d, l);
    if (hu && hl) return min2sel<MI>(u, l);
    if (hd) return d;
    if (hu) return u;
    return l;
}

template <int L, int W, int MI = 0>
__device__ __forceinline__ float dtw_band_f32_x2(const u64 (&ap)[(L + 1) / 2], const u64 (&sp)[L], const float p2) {
    static_assert(L >= 2, "packed recurrence needs at least two rows");
    using B = Band<L, W>;
    constexpr int NP = L / 2;  // row pairs; a trailing single row if L is odd
    const u64 p22 = pack2(p2, p2);
    float Pv[L], Pvp[L];  // previous row: D and D+p2
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const int r0 = 2 * q, r1 = r0 + 1;
        const int lo0 = B::jlo(r0), hi0 = B::jhi(r0), lo1 = B::jlo(r1), hi1 = B::jhi(r1);
        const int phi = (r0 > 0) ? B::jhi(r0 - 1) : 0;
        const int t_last = (hi0 - 1 > hi1) ? (hi0 - 1) : hi1;
        float Xv[L], Xvp[L], Yv[L], Yvp[L];
#pragma unroll
        for (int t = 0; t <= L; t++) {
            if (t >= lo0 && t <= t_last) {
                const int jy = t - 1;
                const bool xv = (t < hi0);
                const bool yv = (jy >= lo1 && jy < hi1);
                // cell (r0, t)
                const bool hdx = (r0 == 0) ? (t == 0) : (t > 0);
                const bool hux = (r0 > 0) && (t < phi);
                const bool hlx = (t > lo0);
                // cell (r1, jy)
                const bool hdy = (jy > 0);
                const bool huy = (jy < hi0);
                const bool hly = (jy > lo1);
                float mx = 0.f, my = 0.f;
                if (xv) mx = min_avail<MI>(hdx, hux, hlx, (r0 == 0) ? 0.f : Pv[(t > 0) ? t - 1 : 0], Pvp[(t < L) ? t : 0], Xvp[(t > 0) ? t - 1 : 0]);
                if (yv) my = min_avail<MI>(hdy, huy, hly, Xv[(jy > 0) ? jy - 1 : 0], Xvp[(jy >= 0) ? jy : 0], Yvp[(jy > 0) ? jy - 1 : 0]);
                if (xv && yv) {
                    const u64 dd = sub2(ap[q], sp[t]);
                    const u64 v2 = fma2(dd, dd, pack2(mx, my));
                    const u64 vp2 = add2(v2, p22);
                    unpack2(v2, Xv[t], Yv[jy]);
                    unpack2(vp2, Xvp[t], Yvp[jy]);
                } else if (xv) {
                    const float s_t = (t >= 1) ? lo2(sp[(t >= 1) ? t : 1]) : hi2(sp[1]);
                    const float diff = lo2(ap[q]) - s_t;
                    Xv[t] = __fmaf_rn(diff, diff, mx);
                    Xvp[t] = Xv[t] + p2;
                } else if (yv) {
                    const float s_j = (jy + 1 < L) ? hi2(sp[(jy + 1 < L) ? jy + 1 : 1]) : lo2(sp[(jy >= 1) ? jy : 1]);
                    const float diff = hi2(ap[q]) - s_j;
                    Yv[jy] = __fmaf_rn(diff, diff, my);
                    Yvp[jy] = Yv[jy] + p2;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < L; j++) {
            if (j >= lo1 && j < hi1) {
                Pv[j] = Yv[j];
                Pvp[j] = Yvp[j];
            }
        }
    }
    if constexpr ((L & 1) == 0) {
        return Pv[L - 1];
    } else {
        // trailing single row (scalar)
        const int i = L - 1;
        const int jlo = B::jlo(i), jhi = B::jhi(i);
        const int phi = (i > 0) ? B::jhi(i - 1) : 0;
        const float a_i = lo2(ap[(L - 1) / 2]);
        float Zv[L], Zvp[L];
#pragma unroll
        for (int j = 0; j < L; j++) {
            if (j >= jlo && j < jhi) {
                const bool hd = (i == 0) ? (j == 0) : (j > 0);
                const bool hu = (i > 0) && (j < phi);
                const bool hl = (j > jlo);
                const float m = min_avail<MI>(hd, hu, hl, (i == 0) ? 0.f : Pv[(j > 0) ? j - 1 : 0], Pvp[j], Zvp[(j > 0) ? j - 1 : 0]);
                const float s_j = (j >= 1) ? lo2(sp[(j >= 1) ? j : 1]) : hi2(sp[1]);
                const float diff = a_i - s_j;
                Zv[j] = __fmaf_rn(diff, diff, m);
                Zvp[j] = Zv[j] + p2;
            }
        }
        return Zv[L - 1];
    }
}


// ---------------------------------------------------------------------------
// Offset form of the packed recurrence ("E form").
//
// With E[i][j] = D[i][j] + p2*(i+j) the two penalised moves and the diagonal move
// all see the same offset:
//     D_diag      = E_diag - p2*(i+j-2)
//     D_up   + p2 = E_up   - p2*(i+j-2)
//     D_left + p2 = E_left - p2*(i+j-2)
// hence  E[i][j] = (a_i - s_j)^2 + 2*p2 + min3(E_diag, E_up, E_left)   and
//        D[L-1][L-1] = E[L-1][L-1] - p2*(2L-2).
// Same instruction count per 2 cells as dtw_band_f32_x2 (FADD2, FFMA2, FADD2,
// 2 x FMNMX3) but
//   * only ONE value per cell is kept (no D+p2 twin): half the row registers,
//   * FADD2 diff and FFMA2 diff^2 + 2*p2 do not depend on the neighbours, so the
//     loop-carried chain is FMNMX3 -> FADD2 instead of FMNMX3 -> FFMA2 -> FADD2.
// Rounding differs from the plain form by O(2^-24 * (D + p2*(2L-2))) absolute in
// D: inside the FAST tolerance (1e-5 relative on the distance) whenever
// D >= DTW_E_FORM_MIN_D2; the caller recomputes smaller results with the plain
// recurrence (identical-fingerprint pairs, D ~ 0, would otherwise lose all
// relative accuracy to cancellation).
// ---------------------------------------------------------------------------
constexpr float DTW_E_FORM_MIN_D2 = 0.0625f;

template <int L, int W, int MI = 0>
__device__ __forceinline__ float dtw_band_f32_x2e(const u64 (&ap)[(L + 1) / 2], const u64 (&sp)[L], const float p2) {
    static_assert(L >= 2, "packed recurrence needs at least two rows");
    using B = Band<L, W>;
    constexpr int NP = L / 2;
    const float c2 = p2 + p2;
    const u64 c22 = pack2(c2, c2);
    float Pv[L];  // previous row (E values)
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const int r0 = 2 * q, r1 = r0 + 1;
        const int lo0 = B::jlo(r0), hi0 = B::jhi(r0), lo1 = B::jlo(r1), hi1 = B::jhi(r1);
        const int phi = (r0 > 0) ? B::jhi(r0 - 1) : 0;
        const int t_last = (hi0 - 1 > hi1) ? (hi0 - 1) : hi1;
        float Xv[L], Yv[L];
#pragma unroll
        for (int t = 0; t <= L; t++) {
            if (t >= lo0 && t <= t_last) {
                const int jy = t - 1;
                const bool xv = (t < hi0);
                const bool yv = (jy >= lo1 && jy < hi1);
                const bool hdx = (r0 == 0) ? (t == 0) : (t > 0);
                const bool hux = (r0 > 0) && (t < phi);
                const bool hlx = (t > lo0);
                const bool hdy = (jy > 0);
                const bool huy = (jy < hi0);
                const bool hly = (jy > lo1);
                float mx = 0.f, my = 0.f;
                // virtual corner E[-1][-1] = -2*p2 makes E[0][0] = diff^2
                if (xv) mx = min_avail<MI>(hdx, hux, hlx, (r0 == 0) ? -c2 : Pv[(t > 0) ? t - 1 : 0], Pv[(t < L) ? t : 0], Xv[(t > 0) ? t - 1 : 0]);
                if (yv) my = min_avail<MI>(hdy, huy, hly, Xv[(jy > 0) ? jy - 1 : 0], Xv[(jy >= 0) ? jy : 0], Yv[(jy > 0) ? jy - 1 : 0]);
                if (xv && yv) {
                    const u64 dd = sub2(ap[q], sp[t]);
                    const u64 dc = fma2(dd, dd, c22);
                    const u64 e2 = add2(dc, pack2(mx, my));
                    unpack2(e2, Xv[t], Yv[jy]);
                } else if (xv) {
                    const float s_t = (t >= 1) ? lo2(sp[(t >= 1) ? t : 1]) : hi2(sp[1]);
                    const float diff = lo2(ap[q]) - s_t;
                    Xv[t] = __fmaf_rn(diff, diff, c2) + mx;
                } else if (yv) {
                    const float s_j = (jy + 1 < L) ? hi2(sp[(jy + 1 < L) ? jy + 1 : 1]) : lo2(sp[(jy >= 1) ? jy : 1]);
                    const float diff = hi2(ap[q]) - s_j;
                    Yv[jy] = __fmaf_rn(diff, diff, c2) + my;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < L; j++)
            if (j >= lo1 && j < hi1) Pv[j] = Yv[j];
    }
    float e_last;
    if constexpr ((L & 1) == 0) {
        e_last = Pv[L - 1];
    } else {
        const int i = L - 1;
        const int jlo = B::jlo(i), jhi = B::jhi(i);
        const int phi = (i > 0) ? B::jhi(i - 1) : 0;
        const float a_i = lo2(ap[(L - 1) / 2]);
        float Zv[L];
#pragma unroll
        for (int j = 0; j < L; j++) {
            if (j >= jlo && j < jhi) {
                const bool hd = (i == 0) ? (j == 0) : (j > 0);
                const bool hu = (i > 0) && (j < phi);
                const bool hl = (j > jlo);
                const float m = min_avail<MI>(hd, hu, hl, (i == 0) ? -c2 : Pv[(j > 0) ? j - 1 : 0], Pv[j], Zv[(j > 0) ? j - 1 : 0]);
                const float s_j = (j >= 1) ? lo2(sp[(j >= 1) ? j : 1]) : hi2(sp[1]);
                const float diff = a_i - s_j;
                Zv[j] = __fmaf_rn(diff, diff, c2) + m;
            }
        }
        e_last = Zv[L - 1];
    }
    const float d2 = e_last - p2 * (float)(2 * L - 2);
    return (d2 < 0.f) ? 0.f : d2;  // not fmaxf: a NaN must stay a NaN
}

}  // namespace wdx
