// Micro-benchmark of the EXACT (float64) DTW recurrence with different minimum implementations.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I warpdemux_b200/csrc scripts/ubench_dtw64.cu -o scripts/ubench_dtw64.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dtw_band.cuh"
using namespace wdx;
constexpr int L = 25, W = 15, NSV = 256;

template <int MINB, int MI>
__global__ void __launch_bounds__(128, MINB) k_f64(const double* __restrict__ X, const double* __restrict__ SV, double* out, int n, double p2) {
    extern __shared__ double smd[];
    for (int q = threadIdx.x; q < NSV * 26; q += blockDim.x) smd[q] = SV[q];
    __syncthreads();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    double a[L];
#pragma unroll
    for (int j = 0; j < L; j++) a[j] = X[(size_t)r * L + j];
    double acc = 0.0;
    for (int s = 0; s < NSV; s++) {
        double sv[L];
        const double* row = smd + s * 26;
#pragma unroll
        for (int j = 0; j < L; j += 2) {
            double2 w = *reinterpret_cast<const double2*>(row + j);
            sv[j] = w.x; if (j + 1 < L) sv[j + 1] = w.y;
        }
        acc += dtw_band_f64<L, W, MI>(a, sv, p2);
    }
    out[r] = acc;
}

int main() {
    int n = 148 * 4 * 128 * 2;
    std::vector<double> X((size_t)n * L), SV(NSV * 26, 0.0);
    srand(1);
    for (auto& v : X) v = (rand() / (double)RAND_MAX - 0.5) * 4.0;
    for (auto& v : SV) v = (rand() / (double)RAND_MAX - 0.5) * 4.0;
    double *dX, *dSV, *o[3];
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dSV, SV.size() * 8);
    for (auto& p : o) cudaMalloc(&p, n * 8);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dSV, SV.data(), SV.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double cells = (double)n * NSV * Band<L, W>::cells();
    const size_t smem = NSV * 26 * 8;
    auto timeit = [&](const char* name, auto kern, double* out) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<n / 128, 128, smem>>>(dX, dSV, out, n, 0.01);
        cudaEventRecord(e0);
        for (int w = 0; w < 3; w++) kern<<<n / 128, 128, smem>>>(dX, dSV, out, n, 0.01);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        printf("%-16s %8.3f ms  %8.1f GCUPS  cycles/cell/SMSP@1965MHz=%.3f  %s\n", name, ms, cells / ms / 1e6,
               1.965e9 * 148 * 4 * 32 / (cells / (ms * 1e-3)), cudaGetErrorString(cudaGetLastError()));
    };
    timeit("f64_dsetp_occ3", k_f64<3, 0>, o[0]);
    timeit("f64_dsetp_occ4", k_f64<4, 0>, o[0]);
    timeit("f64_imin_occ3", k_f64<3, 1>, o[1]);
    timeit("f64_imin_occ4", k_f64<4, 1>, o[1]);
    timeit("f64_fmin_occ3", k_f64<3, 2>, o[2]);
    std::vector<double> h0(n), h1(n), h2(n);
    cudaMemcpy(h0.data(), o[0], n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h1.data(), o[1], n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), o[2], n * 8, cudaMemcpyDeviceToHost);
    int b1 = 0, b2 = 0;
    for (int i = 0; i < n; i++) { b1 += h0[i] != h1[i]; b2 += h0[i] != h2[i]; }
    printf("bitwise mismatches vs dsetp: imin %d, fmin %d of %d\n", b1, b2, n);
    return 0;
}
