// fused_kernels.cuh — the fused DTW + SVC-decision kernel and the probability
// finishing kernel (sm_100a).
//
// Mapping: one thread owns one read for the whole launch; its
// fingerprint lives in registers.  The CTA walks the model's support vectors in
// class-sorted order, TILE_SV at a time, each tile brought into shared memory
// by one TMA bulk copy (cp.async.bulk + mbarrier, double-buffered).  A support
// vector is warp-uniform, so its 25 values reach the registers of all 32 lanes
// by broadcast LDS.128.  After each (read, SV) DTW the distance goes straight
// into the libsvm one-vs-one decision sums — exp kernel, then k-1 multiply-adds
// with that SV's dual coefficients — so the distance matrix never exists in HBM.
// The k-1 running sums of the CURRENT class stay in registers; at a class
// boundary they are parked in / fetched from a small per-read scratch array in
// global memory (k-1 doubles per read per class: noise next to 1.3 M DTW cells).
//
// Summation order is libsvm's (svm.cpp:2868-2896): for pair (i<j) first class-i
// SVs ascending, then class-j SVs ascending — exactly the order a walk over the
// class-sorted SV list produces, so with one SV split the decision values are
// formed in the same order as the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dtw_band.cuh"
#include "dtw_band_x2.cuh"
#include "wdx_types.cuh"

namespace wdx {

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (global -> shared)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ inline int pair_index(int i, int j, int k) {  // i < j
    return i * k - i * (i + 1) / 2 + (j - i - 1);
}

// Does SV range [b,e) contain any SV of class c?
__device__ __forceinline__ bool range_touches(const ModelDev& m, int b, int e, int c) {
    const int lo = max(b, m.class_start[c]), hi = min(e, m.class_start[c + 1]);
    return lo < hi;
}

template <typename T>
__device__ __forceinline__ T kernel_value(T d2, const ModelDev& m);

// K = float32(exp(-gamma * float32(d)^pwr)) as the reference forms it
// (parallel_distances.py:67 casts d to float32; dtw_svm.py:21-22 evaluates the
// kernel in float32); returned widened to double like sklearn's upcast.
__device__ __forceinline__ double kernel_from_dist_f32(float d, float gamma, int pwr) {
    float pw = d;
    if (pwr != 1) {
        pw = 1.0f;
        for (int q = 0; q < pwr; q++) pw *= d;
    }
    return (double)expf(-gamma * pw);
}
// EXACT mode: exp evaluated in double and rounded once to float32 — the
// correctly rounded float32 exp except in double-rounding corner cases.
__device__ __forceinline__ double kernel_from_dist_f32_exact(float d, float gamma, int pwr) {
    float pw = d;
    if (pwr != 1) {
        pw = 1.0f;
        for (int q = 0; q < pwr; q++) pw = __fmul_rn(pw, d);
    }
    const float x = __fmul_rn(-gamma, pw);
    return (double)(float)exp((double)x);
}

// ---------------------------------------------------------------------------
// Fused DTW + SVC decision-sum kernel.
//   EXACT : DTW in float64 (bit-exact), decision sums with separate mul/add
//   !EXACT: DTW in float32, decision sums with DFMA
//   L_, W_: compile-time fingerprint length / window (25 / 15 for every shipped
//           model); L_ == 0 selects the generic runtime-shape fallback.
//   KM1   : compile-time bound on k-1 (number of running decision sums)
//   X2    : FAST only — 0 scalar recurrence, 1 packed f32x2 (dtw_band_x2.cuh), 2 packed f32x2 in
//           offset ("E") form with a plain-recurrence redo of near-zero distances
//   MINB  : CTAs per SM the register allocation is bounded for
//   ACCS  : running decision sums live in shared memory instead of registers
// grid = (ceil(n / CTA_THREADS), n_splits)
// ---------------------------------------------------------------------------
template <bool EXACT, int L_, int W_, int KM1, int X2, int MINB, bool ACCS>
__global__ void __launch_bounds__(CTA_THREADS, MINB)
dtw_svc_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ PredictArgs a) {
    using T = typename std::conditional<EXACT, double, float>::type;
    static_assert(!(X2 && (EXACT || L_ == 0)), "packed recurrence is FAST + specialised shape only");
    constexpr bool GENERIC = (L_ == 0);
    constexpr int LR = GENERIC ? MAXL : L_;             // register/local array length
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int L = GENERIC ? m.L : L_;
    const int ldsv = EXACT ? m.ldd : (X2 ? m.ldp : m.ldf);  // elements per SV row in the tile
    const int sv_row_bytes = ldsv * (int)sizeof(T);
    const int coef_row_bytes = m.ldc * 8;
    const int tile_sv_bytes = TILE_SV * sv_row_bytes;
    const int tile_coef_bytes = TILE_SV * coef_row_bytes;
    const int stage_bytes = tile_sv_bytes + tile_coef_bytes;   // multiple of 16
    unsigned char* stage0 = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * stage_bytes);
    // [ACCS] per-thread column of running sums, after the stage/staging area (offset passed by the host)
    double* acc_s = reinterpret_cast<double*>(smem_raw + a.acc_smem_offset) + threadIdx.x;

    const int64_t n_eff = a.n_idx ? min((int64_t)(*a.n_idx), a.n) : a.n;
    const int64_t cta_first = (int64_t)blockIdx.x * CTA_THREADS;
    if (cta_first >= n_eff) return;                     // whole CTA idle (uniform)
    const int64_t slot = cta_first + tid;               // position in this launch
    const bool active = slot < n_eff;
    const int64_t row = active ? (a.read_idx ? (int64_t)a.read_idx[slot] : slot) : 0;

    const int split = blockIdx.y;
    const int sv_begin = split * a.sv_per_split;
    const int sv_end = min(m.n_sv, sv_begin + a.sv_per_split);
    if (sv_begin >= sv_end) return;

    // ---- fingerprints: coalesced 16-byte loads into shared memory, then each
    // thread lifts its own row into registers.
    T x[X2 ? 1 : LR];
    u64 ap[X2 ? (LR + 1) / 2 : 1];
    {
        const int esz = a.x_is_f32 ? 4 : 8;
        const int row_bytes = L * esz;
        unsigned char* xs = smem_raw;  // aliases the SV stages; released before the pipeline starts
        if (a.read_idx == nullptr) {
            const int64_t rows_here = min((int64_t)CTA_THREADS, n_eff - cta_first);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.X) + cta_first * row_bytes;
            const int64_t total = rows_here * row_bytes;
            const bool al16 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
            if (al16) {
                const int64_t n16 = total >> 4;
                for (int64_t q = tid; q < n16; q += CTA_THREADS)
                    reinterpret_cast<uint4*>(xs)[q] = __ldg(reinterpret_cast<const uint4*>(src) + q);
                for (int64_t q = (n16 << 4) + tid; q < total; q += CTA_THREADS) xs[q] = src[q];
            } else {
                const int64_t n4 = total >> 2;  // rows are at least 4-byte aligned
                for (int64_t q = tid; q < n4; q += CTA_THREADS)
                    reinterpret_cast<uint32_t*>(xs)[q] = __ldg(reinterpret_cast<const uint32_t*>(src) + q);
            }
        } else if (active) {  // gathered rows (recompute list): per-thread copy
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.X) + row * row_bytes;
            for (int q = 0; q < row_bytes; q += 4)
                *reinterpret_cast<uint32_t*>(xs + (int64_t)tid * row_bytes + q) = *reinterpret_cast<const uint32_t*>(src + q);
        }
        __syncthreads();
        auto xin = [&](int j) -> T {
            if (j >= L || !active) return (T)0;
            if (a.x_is_f32) return (T) reinterpret_cast<const float*>(xs)[tid * L + j];
            return (T) reinterpret_cast<const double*>(xs)[tid * L + j];
        };
        if constexpr (X2) {
#pragma unroll
            for (int q = 0; q < (LR + 1) / 2; q++) ap[q] = pack2((float)xin(2 * q), (float)xin(2 * q + 1));
        } else {
#pragma unroll
            for (int j = 0; j < LR; j++) x[j] = xin(j);
        }
        __syncthreads();
    }

    // ---- TMA pipeline over SV tiles
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();  // generic writes to the staging area above precede async-proxy writes
    }
    __syncthreads();
    const int n_tiles = (sv_end - sv_begin + TILE_SV - 1) / TILE_SV;
    auto issue_tile = [&](int t) {
        const int base = sv_begin + t * TILE_SV;
        const int cnt = min(TILE_SV, sv_end - base);
        unsigned char* st = stage0 + (t & 1) * stage_bytes;
        const uint32_t b_sv = cnt * sv_row_bytes, b_cf = cnt * coef_row_bytes;
        mbar_expect_tx(&bars[t & 1], b_sv + b_cf);
        const unsigned char* gsv = EXACT ? reinterpret_cast<const unsigned char*>(m.sv_f64)
                                         : (X2 ? reinterpret_cast<const unsigned char*>(m.sv_x2) : reinterpret_cast<const unsigned char*>(m.sv_f32));
        tma_load_1d(st, gsv + (size_t)base * sv_row_bytes, b_sv, &bars[t & 1]);
        tma_load_1d(st + tile_sv_bytes, reinterpret_cast<const unsigned char*>(m.coef) + (size_t)base * coef_row_bytes, b_cf, &bars[t & 1]);
    };
    if (tid == 0) {
        issue_tile(0);
        if (n_tiles > 1) issue_tile(1);
    }

    const T p2 = (T)m.p2;
    const float gamma_f = (float)m.gamma;
    const int k = m.k, km1 = k - 1;
    double acc[ACCS ? 1 : KM1];
    auto acc_get = [&](int r) -> double { if constexpr (ACCS) return acc_s[r * CTA_THREADS]; else return acc[r]; };
    auto acc_set = [&](int r, double v) { if constexpr (ACCS) acc_s[r * CTA_THREADS] = v; else acc[r] = v; };
#pragma unroll
    for (int r = 0; r < KM1; r++) acc_set(r, 0.0);

    // class of the first SV of this split
    int cls = 0;
    while (m.class_start[cls + 1] <= sv_begin) cls++;
    int next_boundary = m.class_start[cls + 1];

    double* part = a.part + (size_t)split * m.n_pairs * a.part_stride + slot;
    auto pidx = [&](int c, int r) {
        const int o = (r < c) ? r : r + 1;
        return (c < o) ? pair_index(c, o, k) : pair_index(o, c, k);
    };
    auto park = [&](int c) {  // store the running sums of class c
        if (!active) return;
#pragma unroll
        for (int r = 0; r < KM1; r++)
            if (r < km1) part[(size_t)pidx(c, r) * a.part_stride] = acc_get(r);
    };
    auto fetch = [&](int c) {  // resume (or start) the running sums of class c
#pragma unroll
        for (int r = 0; r < KM1; r++) {
            double v0 = 0.0;
            if (r < km1 && active) {
                const int o = (r < c) ? r : r + 1;
                // pair (o,c), o<c, already holds class o's contribution iff class o intersects this split
                if (o < c && range_touches(m, sv_begin, sv_end, o)) v0 = part[(size_t)pidx(c, r) * a.part_stride];
            }
            acc_set(r, v0);
        }
    };

    // A NaN fingerprint gives NaN distances, NaN kernel values and therefore NaN
    // decision sums (0 * NaN = NaN too); the finishing kernel flags the read.
    for (int t = 0; t < n_tiles; t++) {
        const unsigned char* st = stage0 + (t & 1) * stage_bytes;
        mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
        const int base = sv_begin + t * TILE_SV;
        const int cnt = min(TILE_SV, sv_end - base);
        for (int q = 0; q < cnt; q++) {
            const int s_glob = base + q;
            if (s_glob == next_boundary) {  // warp-uniform
                park(cls);
                do { cls++; } while (m.class_start[cls + 1] <= s_glob);
                next_boundary = m.class_start[cls + 1];
                fetch(cls);
            }
            // support vector -> registers (broadcast 16-byte shared loads)
            const T* srow = reinterpret_cast<const T*>(st + q * sv_row_bytes);
            T d2;
            if constexpr (X2) {
                u64 sp[LR];
                sp[0] = 0;
#pragma unroll
                for (int t = 1; t < L_; t += 2) {  // (s[t],s[t-1],s[t+1],s[t]) per 16-byte load
                    const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(srow) + (t - 1) * 2);
                    sp[t] = pack2(w.x, w.y);
                    if (t + 1 < L_) sp[t + 1] = pack2(w.z, w.w);
                }
                if constexpr (X2 == 2) {
                    d2 = dtw_band_f32_x2e<L_, W_>(ap, sp, p2);
                    if (d2 < DTW_E_FORM_MIN_D2) {  // rare: cancellation would cost relative accuracy
                        float xa[MAXL], sa[MAXL];
#pragma unroll
                        for (int j = 0; j < L_; j++) xa[j] = (j & 1) ? hi2(ap[j >> 1]) : lo2(ap[j >> 1]);
                        sa[0] = reinterpret_cast<const float*>(srow)[1];
                        for (int j = 1; j < L_; j++) sa[j] = reinterpret_cast<const float*>(srow)[2 * (j - 1)];
                        d2 = dtw_generic<float, MAXL>(xa, sa, L_, m.window, p2);
                    }
                } else {
                    d2 = dtw_band_f32_x2<L_, W_>(ap, sp, p2);
                }
            } else if constexpr (!GENERIC) {
                T s[LR];
                constexpr int VEC = 16 / sizeof(T);
#pragma unroll
                for (int j = 0; j < L_; j += VEC) {
                    if constexpr (EXACT) {
                        const double2 w = *reinterpret_cast<const double2*>(srow + j);
                        s[j] = w.x;
                        if (j + 1 < L_) s[j + 1] = w.y;
                    } else {
                        const float4 w = *reinterpret_cast<const float4*>(srow + j);
                        s[j] = w.x;
                        if (j + 1 < L_) s[j + 1] = w.y;
                        if (j + 2 < L_) s[j + 2] = w.z;
                        if (j + 3 < L_) s[j + 3] = w.w;
                    }
                }
                if constexpr (EXACT) d2 = dtw_band_f64<L_, W_>(x, s, p2);
                else d2 = dtw_band_f32<L_, W_>(x, s, p2);
            } else {
                T s[LR];
                for (int j = 0; j < L; j++) s[j] = srow[j];
                d2 = dtw_generic<T, MAXL>(x, s, L, m.window, p2);
            }
            // distance -> float32 (the reference's cast) -> kernel value
            float d32;
            if constexpr (EXACT) d32 = (float)__dsqrt_rn(d2);
            else d32 = sqrtf(d2);
            if (a.dist && active) a.dist[(size_t)row * m.n_sv + s_glob] = d32;
            double Kd;
            if constexpr (EXACT) Kd = kernel_from_dist_f32_exact(d32, gamma_f, m.pwr_dist);
            else Kd = kernel_from_dist_f32(d32, gamma_f, m.pwr_dist);
            // decision sums of the current class (svm.cpp:2880-2884)
            const double* crow = reinterpret_cast<const double*>(st + tile_sv_bytes + q * coef_row_bytes);
#pragma unroll
            for (int r = 0; r < KM1; r += 2) {
                if (r < km1) {  // ldc is even, so the 16-byte load stays inside the row
                    const double2 c2 = *reinterpret_cast<const double2*>(crow + r);
                    if constexpr (EXACT) acc_set(r, __dadd_rn(acc_get(r), __dmul_rn(c2.x, Kd)));
                    else acc_set(r, __fma_rn(c2.x, Kd, acc_get(r)));
                    if (r + 1 < km1) {
                        if constexpr (EXACT) acc_set(r + 1, __dadd_rn(acc_get(r + 1), __dmul_rn(c2.y, Kd)));
                        else acc_set(r + 1, __fma_rn(c2.y, Kd, acc_get(r + 1)));
                    }
                }
            }
        }
        __syncthreads();  // every warp is done with this stage
        if (tid == 0 && t + 2 < n_tiles) issue_tile(t + 2);
    }
    park(cls);
}

// ---------------------------------------------------------------------------
// Finishing kernel: decision sums -> Platt sigmoid -> pairwise coupling
// (Wu-Lin-Weng method 2) -> argmax / margin / threshold.  One thread per read,
// float64, no FMA contraction (this TU is compiled with -fmad=false).
// Restates svm.cpp:2035-2104, 2921-2964 and models/utils.py:45-61.
// ---------------------------------------------------------------------------
struct FinishArgs {
    const double* part;
    int64_t part_stride;
    int n_splits, sv_per_split;
    const int* read_idx;   // optional: slot -> output row
    const int* n_idx;      // optional device count
    int64_t n;
    int64_t* labels;
    double* conf;
    double* prob;          // [n][k] or nullptr
    uint8_t* flags;        // or nullptr
    // GUARDED: collect reads close to a decision boundary
    int* near_idx;         // or nullptr
    int* near_count;       // [0] = list length (clamped by the consumer), [1] = reads that did not fit
    int near_cap;
    double guard;
    uint8_t flag_or;       // bits OR-ed into flags (WDX_FLAG_RECOMPUTED on the exact re-run)
};

// Small batches cut the support-vector list into ranges (grid.y) so that the GPU is filled; this
// kernel folds the per-range partial decision sums into range 0's plane so that the finishing kernel
// does not walk n_splits x n_pairs dependent loads on its own.  One (pair, read) per group of
// FOLD_SUB threads: thread `sub` adds the ranges sub, sub + FOLD_SUB, ... in ascending order, then the
// FOLD_SUB partial sums are added in ascending `sub` order - a fixed order for a given n_splits, and a
// dependency chain FOLD_SUB times shorter than one thread walking all ranges (the fold was 20-40 us of a
// live-sized call).  A warp covers 32 consecutive reads of one `sub`, so the loads stay coalesced.
constexpr int FOLD_SUB = 8;
__global__ void __launch_bounds__(256) svc_fold_splits_kernel(const __grid_constant__ ModelDev m, double* __restrict__ part,
                                                              int64_t part_stride, int n_splits, int sv_per_split,
                                                              const int* __restrict__ n_idx, int64_t n) {
    __shared__ double acc[FOLD_SUB][32];
    const int64_t n_eff = n_idx ? min((int64_t)(*n_idx), n) : n;
    const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * 32 + lane;
    const int pi = blockIdx.y;
    if ((int64_t)blockIdx.x * 32 >= n_eff) return;   // whole CTA
    int i = 0, rem = pi;  // pair index -> (i, j), i < j
    while (rem >= m.k - 1 - i) {
        rem -= m.k - 1 - i;
        i++;
    }
    const int j = i + 1 + rem;
    double sum = 0.0;
    if (slot < n_eff) {
#pragma unroll 4
        for (int s = sub; s < n_splits; s += FOLD_SUB) {
            const int b = s * sv_per_split, e = min(m.n_sv, b + sv_per_split);
            const bool live = b < e && (range_touches(m, b, e, i) || range_touches(m, b, e, j));
            if (live) sum += part[((size_t)s * m.n_pairs + pi) * part_stride + slot];
        }
    }
    acc[sub][lane] = sum;
    __syncthreads();
    if (sub == 0 && slot < n_eff) {
        double t = acc[0][lane];
#pragma unroll
        for (int q = 1; q < FOLD_SUB; q++) t += acc[q][lane];
        part[(size_t)pi * part_stride + slot] = t;
    }
}

// Same arithmetic as svc_finish_kernel below, one WARP per read: for live-sized batches the single-thread
// coupling iteration (up to 100 sweeps over a k x k system with a division per element) is the latency of the
// whole call (150 us at k = 11).  Lane t owns row t of Q, p[t] and Qp[t]; every sum runs in the serial order
// (operands fetched with shuffles), the Gauss-Seidel sweep stays sequential in t but its k-wide updates and
// divisions run across the lanes - bit-identical results, ~k times shorter dependency chain.
constexpr int FINISH_WARPS = 4;
__global__ void __launch_bounds__(FINISH_WARPS * 32) svc_finish_warp_kernel(const __grid_constant__ ModelDev m,
                                                                            const __grid_constant__ FinishArgs a) {
    __shared__ double Rs[FINISH_WARPS][MAXK][MAXK];
    __shared__ double Qs[FINISH_WARPS][MAXK][MAXK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t n_eff = a.n_idx ? min((int64_t)(*a.n_idx), a.n) : a.n;
    const int64_t slot = (int64_t)blockIdx.x * FINISH_WARPS + w;
    if (slot >= n_eff) return;   // whole warp
    const int64_t row = a.read_idx ? (int64_t)a.read_idx[slot] : slot;
    const int k = m.k;
    const unsigned full = 0xffffffffu;
    double (*R)[MAXK] = Rs[w];
    double (*Q)[MAXK] = Qs[w];

    bool bad = false;
    {
        int pi = 0;
        for (int i = 0; i < k; i++)
            for (int j = i + 1; j < k; j++, pi++) {
                if ((pi & 31) != lane) continue;
                double sum = 0.0;
                for (int sp = 0; sp < a.n_splits; sp++) {
                    const int b = sp * a.sv_per_split, e = min(m.n_sv, b + a.sv_per_split);
                    if (b < e && (range_touches(m, b, e, i) || range_touches(m, b, e, j)))
                        sum += a.part[((size_t)sp * m.n_pairs + pi) * a.part_stride + slot];
                }
                const double dec = sum - m.rho[pi];
                bad |= !(fabs(dec) <= 1.7e308);
                const double fApB = dec * m.probA[pi] + m.probB[pi];
                double r = (fApB >= 0) ? exp(-fApB) / (1.0 + exp(-fApB)) : 1.0 / (1 + exp(fApB));
                const double min_prob = 1e-7;
                r = (r < min_prob) ? min_prob : r;
                r = ((1 - min_prob) < r) ? (1 - min_prob) : r;
                R[i][j] = r;
                R[j][i] = 1 - r;
            }
    }
    bad = __any_sync(full, bad);
    __syncwarp();

    const bool act = lane < k;
    const int t_own = act ? lane : 0;
    double p = 0.0, Qp = 0.0;
    bool fragile = false;   // the sweep count of the coupling iteration could change under a FAST_F32-sized perturbation
    if (!bad) {
        const int max_iter = (k > 100) ? k : 100;
        const double eps = 0.005 / k;
        if (act) {
            const int t = lane;
            double qtt = 0;
            for (int j = 0; j < t; j++) { qtt += R[j][t] * R[j][t]; Q[t][j] = -R[j][t] * R[t][j]; }
            for (int j = t + 1; j < k; j++) { qtt += R[j][t] * R[j][t]; Q[t][j] = -R[j][t] * R[t][j]; }
            Q[t][t] = qtt;
        }
        p = 1.0 / k;
        __syncwarp();
        for (int iter = 0; iter < max_iter; iter++) {
            Qp = 0;
            for (int j = 0; j < k; j++) {
                const double pj = __shfl_sync(full, p, j);
                Qp += Q[t_own][j] * pj;
            }
            double pQp = 0;
            for (int t = 0; t < k; t++) pQp += __shfl_sync(full, p, t) * __shfl_sync(full, Qp, t);
            double err = act ? fabs(Qp - pQp) : 0.0;
#pragma unroll
            for (int o = 16; o; o >>= 1) err = fmax(err, __shfl_xor_sync(full, err, o));
            fragile |= fabs(err - eps) < a.guard;
            if (err < eps) break;
            for (int t = 0; t < k; t++) {
                const double Qp_t = __shfl_sync(full, Qp, t);
                const double qtt = Q[t][t];
                const double diff = (-Qp_t + pQp) / qtt;
                if (lane == t) p += diff;
                pQp = (pQp + diff * (diff * qtt + 2 * Qp_t)) / (1 + diff) / (1 + diff);
                if (act) {
                    Qp = (Qp + diff * Q[t][lane]) / (1 + diff);
                    p /= (1 + diff);
                }
            }
        }
    }
    // process_probs (models/utils.py:45-61): first maximum, top1 - top2 — every lane walks the classes in order
    int best = k - 1;
    double top1 = 0.0, top2 = 0.0;
    if (!bad) {
        best = 0;
        double pb = __shfl_sync(full, p, 0);
        for (int c = 1; c < k; c++) {
            const double pc = __shfl_sync(full, p, c);
            if (pc > pb) { pb = pc; best = c; }
        }
        top1 = pb;
        top2 = -INFINITY;
        for (int c = 0; c < k; c++) {
            const double pc = __shfl_sync(full, p, c);
            if (c != best && pc > top2) top2 = pc;
        }
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double conf = bad ? nan : (top1 - top2);
    const double thr = m.thresholds[best];
    if (a.prob && act) a.prob[(size_t)row * k + lane] = bad ? nan : p;
    if (lane == 0) {
        int64_t label = m.label_map[best];
        if (bad || conf < thr) label = -1;
        a.labels[row] = label;
        if (a.conf) a.conf[row] = conf;
        if (a.flags) a.flags[row] = (uint8_t)((bad ? 1 : 0) | a.flag_or);
        if (a.near_idx && !bad) {
            const double band = fragile ? 100.0 * a.guard : a.guard;   // see svc_finish_kernel
            const bool near = (fabs(conf - thr) < band) || (conf < band);
            if (near) {
                const int pos = atomicAdd(a.near_count, 1);
                if (pos < a.near_cap) {
                    a.near_idx[pos] = (int)row;
                } else {
                    atomicAdd(a.near_count + 1, 1);
                    if (a.flags) a.flags[row] |= 4;  // WDX_FLAG_GUARD_OVERFLOW
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) svc_finish_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ FinishArgs a) {
    const int64_t n_eff = a.n_idx ? min((int64_t)(*a.n_idx), a.n) : a.n;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_eff) return;
    const int64_t row = a.read_idx ? (int64_t)a.read_idx[slot] : slot;
    const int k = m.k;

    double R[MAXK][MAXK];
    double Q[MAXK][MAXK];
    double Qp[MAXK], p[MAXK];
    bool bad = false;
    int pi = 0;
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++, pi++) {
            double sum = 0.0;
            for (int s = 0; s < a.n_splits; s++) {
                const int b = s * a.sv_per_split, e = min(m.n_sv, b + a.sv_per_split);
                if (b < e && (range_touches(m, b, e, i) || range_touches(m, b, e, j)))
                    sum += a.part[((size_t)s * m.n_pairs + pi) * a.part_stride + slot];
            }
            const double dec = sum - m.rho[pi];
            bad |= !(fabs(dec) <= 1.7e308);
            const double fApB = dec * m.probA[pi] + m.probB[pi];
            double r = (fApB >= 0) ? exp(-fApB) / (1.0 + exp(-fApB)) : 1.0 / (1 + exp(fApB));
            const double min_prob = 1e-7;
            r = (r < min_prob) ? min_prob : r;            // max(r, min_prob)
            r = ((1 - min_prob) < r) ? (1 - min_prob) : r;  // min(r, 1-min_prob)
            R[i][j] = r;
            R[j][i] = 1 - r;
        }

    int best = k - 1;
    double top1 = 0.0, top2 = 0.0;
    bool fragile = false;
    if (!bad) {
        // multiclass_probability (svm.cpp:2046-2104)
        const int max_iter = (k > 100) ? k : 100;
        const double eps = 0.005 / k;
        for (int t = 0; t < k; t++) {
            p[t] = 1.0 / k;
            Q[t][t] = 0;
            for (int j = 0; j < t; j++) { Q[t][t] += R[j][t] * R[j][t]; Q[t][j] = Q[j][t]; }
            for (int j = t + 1; j < k; j++) { Q[t][t] += R[j][t] * R[j][t]; Q[t][j] = -R[j][t] * R[t][j]; }
        }
        for (int iter = 0; iter < max_iter; iter++) {
            double pQp = 0;
            for (int t = 0; t < k; t++) {
                Qp[t] = 0;
                for (int j = 0; j < k; j++) Qp[t] += Q[t][j] * p[j];
                pQp += p[t] * Qp[t];
            }
            double max_error = 0;
            for (int t = 0; t < k; t++) {
                const double err = fabs(Qp[t] - pQp);
                if (err > max_error) max_error = err;
            }
            fragile |= fabs(max_error - eps) < a.guard;
            if (max_error < eps) break;
            for (int t = 0; t < k; t++) {
                const double diff = (-Qp[t] + pQp) / Q[t][t];
                p[t] += diff;
                pQp = (pQp + diff * (diff * Q[t][t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
                for (int j = 0; j < k; j++) {
                    Qp[j] = (Qp[j] + diff * Q[t][j]) / (1 + diff);
                    p[j] /= (1 + diff);
                }
            }
        }
        // process_probs (models/utils.py:45-61): first maximum, top1 - top2
        best = 0;
        for (int c = 1; c < k; c++)
            if (p[c] > p[best]) best = c;
        top1 = p[best];
        top2 = -INFINITY;
        for (int c = 0; c < k; c++)
            if (c != best && p[c] > top2) top2 = p[c];
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double conf = bad ? nan : (top1 - top2);
    const double thr = m.thresholds[best];
    int64_t label = m.label_map[best];
    if (bad || conf < thr) label = -1;
    a.labels[row] = label;
    if (a.conf) a.conf[row] = conf;
    if (a.prob)
        for (int c = 0; c < k; c++) a.prob[(size_t)row * k + c] = bad ? nan : p[c];
    if (a.flags) a.flags[row] = (uint8_t)((bad ? 1 : 0) | a.flag_or);
    if (a.near_idx && !bad) {
        // GUARDED: which reads are redone in EXACT_F64.  FAST_F32 moves a confidence by < 1e-5 (measured over 12.5 M WDX10
        // reads) as long as libsvm's coupling iteration stops after the same number of sweeps; a read whose stopping test
        // (max_error < eps) was within `guard` of flipping at some sweep can differ by one sweep, i.e. by up to ~1e-4 (one read
        // in 12.5 M: 5.8e-5).  So: the plain band `guard` around the threshold / the top-2 tie for stable reads, a band 100 x
        // wider for the fragile ones.
        const double band = fragile ? 100.0 * a.guard : a.guard;
        const bool near = (fabs(conf - thr) < band) || (conf < band);
        if (near) {
            const int pos = atomicAdd(a.near_count, 1);
            if (pos < a.near_cap) {
                a.near_idx[pos] = (int)row;
            } else {  // list full: keep the FAST result, tell the caller
                atomicAdd(a.near_count + 1, 1);
                if (a.flags) a.flags[row] |= 4;  // WDX_FLAG_GUARD_OVERFLOW
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Plain distance-matrix kernel (secondary seam, parallel_distances.py:48-84):
// thread = row of X, loop over the rows of Y staged through shared memory.
// ---------------------------------------------------------------------------
template <bool EXACT, int L_, int W_, typename OutT>
__global__ void __launch_bounds__(CTA_THREADS) dtw_matrix_kernel(const double* __restrict__ X, int64_t nX,
                                                                  const double* __restrict__ Y, int64_t nY, int L,
                                                                  int window, double p2d, OutT* __restrict__ out,
                                                                  int y_per_split) {
    using T = typename std::conditional<EXACT, double, float>::type;
    constexpr bool GENERIC = (L_ == 0);
    constexpr int LR = GENERIC ? MAXL : L_;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* ys = reinterpret_cast<T*>(smem_raw);  // [TILE_SV][L]
    const int tid = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * CTA_THREADS + tid;
    const bool active = row < nX;
    const int64_t y_begin = (int64_t)blockIdx.y * y_per_split;
    const int64_t y_end = min(nY, y_begin + (int64_t)y_per_split);
    T x[LR];
#pragma unroll
    for (int j = 0; j < LR; j++) x[j] = (j < L && active) ? (T)X[row * L + j] : (T)0;
    const T p2 = (T)p2d;
    for (int64_t base = y_begin; base < y_end; base += TILE_SV) {
        const int cnt = (int)min((int64_t)TILE_SV, y_end - base);
        __syncthreads();
        for (int q = tid; q < cnt * L; q += CTA_THREADS) ys[q] = (T)Y[base * L + q];
        __syncthreads();
        for (int q = 0; q < cnt; q++) {
            T s[LR];
            T d2;
            if constexpr (!GENERIC) {
#pragma unroll
                for (int j = 0; j < L_; j++) s[j] = ys[q * L_ + j];
                if constexpr (EXACT) d2 = dtw_band_f64<L_, W_>(x, s, p2);
                else d2 = dtw_band_f32<L_, W_>(x, s, p2);
            } else {
                for (int j = 0; j < L; j++) s[j] = ys[q * L + j];
                d2 = dtw_generic<T, MAXL>(x, s, L, window, p2);
            }
            if (active) {
                if constexpr (EXACT) out[(size_t)row * nY + base + q] = (OutT)__dsqrt_rn(d2);
                else out[(size_t)row * nY + base + q] = (OutT)sqrtf(d2);
            }
        }
    }
}

}  // namespace wdx
