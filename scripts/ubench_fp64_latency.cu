// Dependent-issue latency of float64 / float32 / integer instructions on one warp (no contention) and with the other warps of
// the SM busy on the FP64 pipe.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench_fp64_latency.bin scripts/ubench_fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, int iters, int busy_warps) {
    const int warp = threadIdx.x >> 5;
    double a = out[0], b = out[1];
    float fa = (float)a, fb = (float)b;
    if (warp == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (OP == 0) a = __dadd_rn(a, b);
                if (OP == 1) a = __dmul_rn(a, b);
                if (OP == 2) a = __fma_rn(a, b, b);
                if (OP == 3) fa = __fadd_rn(fa, fb);
                if (OP == 4) a = (a < b) ? __dadd_rn(a, 1.0) : b;   // compare + select + add
            }
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) cyc[0] = t1 - t0;
    } else if (warp <= busy_warps) {   // independent FP64 streams on the other warps (pipe contention)
        double x0 = a, x1 = b, x2 = a + 1, x3 = b + 1;
        for (int i = 0; i < iters * 4; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b);
            }
        }
        a += x0 + x1 + x2 + x3;
    }
    out[2 + threadIdx.x] = a + fa;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 8);
    double h[2] = {1.0, 1e-9};
    cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
    const char* names[5] = {"DADD", "DMUL", "DFMA", "FADD", "DSETP+SEL+DADD"};
    const int iters = 4096;
    for (int busy = 0; busy <= 31; busy = busy ? busy * 2 + 1 : 3) {
        for (int op = 0; op < 5; op++) {
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                    case 0: chain<0><<<1, 1024>>>(out, cyc, iters, busy); break;
                    case 1: chain<1><<<1, 1024>>>(out, cyc, iters, busy); break;
                    case 2: chain<2><<<1, 1024>>>(out, cyc, iters, busy); break;
                    case 3: chain<3><<<1, 1024>>>(out, cyc, iters, busy); break;
                    case 4: chain<4><<<1, 1024>>>(out, cyc, iters, busy); break;
                }
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-16s busy FP64 warps %2d : %.2f cycles per dependent op\n", names[op], busy, (double)c / (iters * 16.0));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
