// wdx_types.cuh — plain structs and constants shared by the kernels and the host code.
#pragma once
#include <stdint.h>

namespace wdx {

constexpr int CTA_THREADS = 128;
constexpr int TILE_SV = 32;
constexpr int MAXK = 16;
constexpr int MAXL = 64;

struct ModelDev {
    // support vectors, tile-friendly rows (16-byte multiples for TMA bulk copies)
    const float* sv_f32;    // [n_sv][ldf]   ldf = roundup(L,4)
    const float* sv_x2;     // [n_sv][ldp]   pre-paired rows (s[t], s[t-1]), t = 1..L-1; ldp = roundup(2(L-1),4)
    const double* sv_f64;   // [n_sv][ldd]   ldd = roundup(L,2)
    const double* coef;     // [n_sv][ldc]   coef[s][r] = dual_coef[r][s], ldc = roundup(k-1,2)
    const double* rho;      // [n_pairs]
    const double* probA;    // [n_pairs]
    const double* probB;    // [n_pairs]
    const double* thresholds;  // [k]
    const int64_t* label_map;  // [k]
    int class_start[MAXK + 1];
    int n_sv, L, k, n_pairs, ldf, ldd, ldc, ldp;
    int window;
    double p2;      // penalty^2
    double gamma;
    int pwr_dist;
};

struct PredictArgs {
    const void* X;        // [n][L] row-major, f64 or f32
    int x_is_f32;
    const int* read_idx;  // optional indirection (GUARDED recompute list) or nullptr
    const int* n_idx;     // device count for read_idx (nullptr => n)
    int64_t n;            // reads in this launch (upper bound when n_idx != nullptr)
    int64_t part_stride;  // elements between consecutive (split,pair) planes (>= n)
    double* part;         // [n_splits][n_pairs][part_stride] partial decision sums
    float* dist;          // optional [n][n_sv] float32 distances (debug / secondary seam), or nullptr
    int n_splits, sv_per_split;
    int acc_smem_offset;  // byte offset of the shared-memory accumulator block (ACCS kernels)
};

}  // namespace wdx
