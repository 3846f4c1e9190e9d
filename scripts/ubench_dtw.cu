// Micro-benchmark of the DTW inner recurrences (scalar vs packed f32x2).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I warpdemux_b200/csrc scripts/ubench_dtw.cu -o /tmp/ubench_dtw
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cstring>
#include "dtw_band.cuh"
#include "dtw_band_x2.cuh"
using namespace wdx;
constexpr int L = 25, W = 15, NSV = 512;

template <int MINB, int MI, int PI = 0>
__global__ void __launch_bounds__(128, MINB) k_scalar(const float* __restrict__ X, const float* __restrict__ SV, float* out, int n, float p2rt) {
    const float p2v = PI ? 0.01f : p2rt;
    extern __shared__ float sm[];
    for (int q = threadIdx.x; q < NSV * 28; q += blockDim.x) sm[q] = SV[q];
    __syncthreads();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    float a[L];
#pragma unroll
    for (int j = 0; j < L; j++) a[j] = X[(size_t)r * L + j];
    float acc = 0.f;
    for (int s = 0; s < NSV; s++) {
        float sv[L];
        const float* row = sm + s * 28;
#pragma unroll
        for (int j = 0; j < L; j += 4) {
            float4 w = *reinterpret_cast<const float4*>(row + j);
            sv[j] = w.x; if (j + 1 < L) sv[j + 1] = w.y; if (j + 2 < L) sv[j + 2] = w.z; if (j + 3 < L) sv[j + 3] = w.w;
        }
        acc += dtw_band_f32<L, W, MI>(a, sv, p2v);
    }
    out[r] = acc;
}

template <int MINB, int MI, int PI = 0>
__global__ void __launch_bounds__(128, MINB) k_x2(const float* __restrict__ X, const float* __restrict__ SVP, float* out, int n, float p2rt) {
    const float p2v = PI ? 0.01f : p2rt;
    extern __shared__ float sm[];
    for (int q = threadIdx.x; q < NSV * 48; q += blockDim.x) sm[q] = SVP[q];
    __syncthreads();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    u64 ap[(L + 1) / 2];
#pragma unroll
    for (int q = 0; q < (L + 1) / 2; q++) ap[q] = pack2(X[(size_t)r * L + 2 * q], (2 * q + 1 < L) ? X[(size_t)r * L + 2 * q + 1] : 0.f);
    float acc = 0.f;
    for (int s = 0; s < NSV; s++) {
        u64 sp[L];
        const float* row = sm + s * 48;
#pragma unroll
        for (int t = 1; t < L; t += 2) {  // pairs t, t+1 -> one 16-byte load
            float4 w = *reinterpret_cast<const float4*>(row + (t - 1) * 2);
            sp[t] = pack2(w.x, w.y);
            if (t + 1 < L) sp[t + 1] = pack2(w.z, w.w);
        }
        sp[0] = 0;
        acc += dtw_band_f32_x2<L, W, MI>(ap, sp, p2v);
    }
    out[r] = acc;
}

// offset ("E") form, the production default; MI = 1: 3-input minimum as VIMNMX3 on the bit patterns
template <int MINB, int MI>
__global__ void __launch_bounds__(128, MINB) k_x2e(const float* __restrict__ X, const float* __restrict__ SVP, float* out, int n, float p2rt) {
    extern __shared__ float sm[];
    for (int q = threadIdx.x; q < NSV * 48; q += blockDim.x) sm[q] = SVP[q];
    __syncthreads();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    u64 ap[(L + 1) / 2];
#pragma unroll
    for (int q = 0; q < (L + 1) / 2; q++) ap[q] = pack2(X[(size_t)r * L + 2 * q], (2 * q + 1 < L) ? X[(size_t)r * L + 2 * q + 1] : 0.f);
    float acc = 0.f;
    for (int s = 0; s < NSV; s++) {
        u64 sp[L];
        const float* row = sm + s * 48;
#pragma unroll
        for (int t = 1; t < L; t += 2) {
            float4 w = *reinterpret_cast<const float4*>(row + (t - 1) * 2);
            sp[t] = pack2(w.x, w.y);
            if (t + 1 < L) sp[t + 1] = pack2(w.z, w.w);
        }
        sp[0] = 0;
        acc += dtw_band_f32_x2e<L, W, MI>(ap, sp, p2rt);
    }
    out[r] = acc;
}

int main(int argc, char** argv) {
    int n = 148 * 4 * 128 * 4;
    std::vector<float> X((size_t)n * L), SV(NSV * 28, 0.f), SVP(NSV * 48, 0.f);
    srand(1);
    for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (int s = 0; s < NSV; s++) {
        float t[L];
        for (int j = 0; j < L; j++) { t[j] = (rand() / (float)RAND_MAX - 0.5f) * 4.f; SV[s * 28 + j] = t[j]; }
        for (int j = 1; j < L; j++) { SVP[s * 48 + (j - 1) * 2] = t[j]; SVP[s * 48 + (j - 1) * 2 + 1] = t[j - 1]; }
    }
    float *dX, *dSV, *dSVP, *o1, *o2;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dSV, SV.size() * 4); cudaMalloc(&dSVP, SVP.size() * 4);
    cudaMalloc(&o1, n * 4); cudaMalloc(&o2, n * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dSV, SV.data(), SV.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dSVP, SVP.data(), SVP.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double cells = (double)n * NSV * Band<L, W>::cells();
    auto timeit = [&](const char* name, auto kern, size_t smem, float* o) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int w = 0; w < 2; w++) kern<<<n / 128, 128, smem>>>(dX, name[0] == 's' ? dSV : dSVP, o, n, 0.01f);
        cudaEventRecord(e0);
        for (int w = 0; w < 3; w++) kern<<<n / 128, 128, smem>>>(dX, name[0] == 's' ? dSV : dSVP, o, n, 0.01f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        cudaError_t e = cudaGetLastError();
        printf("%-14s %8.3f ms  %8.1f GCUPS  cycles/cell/SMSP@1965MHz=%.3f  %s\n", name, ms, cells / ms / 1e6,
               1.965e9 * 148 * 4 * 32 / (cells / (ms * 1e-3)), cudaGetErrorString(e));
    };
    timeit("scalar_occ4", k_scalar<4, 0, 0>, NSV * 28 * 4, o1);
    timeit("scalar_occ4_imm", k_scalar<4, 0, 1>, NSV * 28 * 4, o1);
    timeit("scalar_occ5_imm", k_scalar<5, 0, 1>, NSV * 28 * 4, o1);
    timeit("scalar_occ6_imm", k_scalar<6, 0, 1>, NSV * 28 * 4, o1);
    timeit("scalar_occ4_2min", k_scalar<4, 2, 0>, NSV * 28 * 4, o1);
    timeit("x2_occ3", k_x2<3, 0, 0>, NSV * 48 * 4, o2);
    timeit("x2_occ3_imm", k_x2<3, 0, 1>, NSV * 48 * 4, o2);
    timeit("x2_occ2_imm", k_x2<2, 0, 1>, NSV * 48 * 4, o2);
    timeit("x2_occ3_2min", k_x2<3, 2, 0>, NSV * 48 * 4, o2);
    timeit("x2_occ2_2min", k_x2<2, 2, 0>, NSV * 48 * 4, o2);
    timeit("x2e_occ4_fmnmx3", k_x2e<4, 0>, NSV * 48 * 4, o2);
    timeit("x2e_occ4_vimnmx3", k_x2e<4, 1>, NSV * 48 * 4, o2);
    timeit("x2e_occ3_vimnmx3", k_x2e<3, 1>, NSV * 48 * 4, o2);
    timeit("x2_occ3_vimnmx3", k_x2<3, 1, 0>, NSV * 48 * 4, o2);
    {
        k_x2e<4, 0><<<n / 128, 128, NSV * 48 * 4>>>(dX, dSVP, o1, n, 0.01f);
        k_x2e<4, 1><<<n / 128, 128, NSV * 48 * 4>>>(dX, dSVP, o2, n, 0.01f);
        std::vector<float> g1(n), g2(n);
        cudaMemcpy(g1.data(), o1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(g2.data(), o2, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < n; i++) if (memcmp(&g1[i], &g2[i], 4)) bad++;
        printf("E form, VIMNMX3 vs FMNMX3 bitwise mismatches: %d of %d\n", bad, n);
    }
    k_scalar<4, 0><<<n / 128, 128, NSV * 28 * 4>>>(dX, dSV, o1, n, 0.01f);
    k_x2<3, 2><<<n / 128, 128, NSV * 48 * 4>>>(dX, dSVP, o2, n, 0.01f);
    std::vector<float> h1(n), h2(n);
    cudaMemcpy(h1.data(), o1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), o2, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < n; i++) if (h1[i] != h2[i]) bad++;
    printf("x2 vs scalar mismatches: %d of %d (sample %.6f %.6f)\n", bad, n, h1[5], h2[5]);
    return 0;
}
