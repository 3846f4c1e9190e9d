// cnn_tc_kernel.cuh — FAST mode of the boundary CNN: the whole forward pass of one read in one
// persistent CTA, the two 64->64 convolutions (97 % of the flops) on the 5th-generation tensor
// cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Restates BoundariesCNN.forward (warpdemux/adapted/adapted/detect/cnn.py:16-52) on the prepared
// input of cnn.py:71-85.  Per read (T1 = 584 hidden positions for the 18 500-sample rows WarpDemuX
// preloads):
//   conv1 (1->64, stride 3)          CUDA cores, float32       -> activations as fp16 hi + lo in smem
//   conv 64->64, k = 7, twice        7 taps x 5 M-tiles x 3 split products x 4 K-steps of
//                                    tcgen05.mma.kind::f16 128x64x16, float32 accumulate in TMEM
//   epilogue                         tcgen05.ld -> bias, ReLU -> fp16 hi + lo back into smem (in place)
//   ConvTranspose1d (64->2, stride 3) CUDA cores, float32      -> scores [2][To] to HBM
//
// The convolution as GEMMs: out[t][co] = sum_tap sum_ci act[t + tap - 3][ci] * W[co][ci][tap].  Activations
// live in shared memory time-major in the canonical NO-SWIZZLE K-major UMMA layout with a uniform 16-byte
// row pitch: element (row, ci) of split s at  s*A_SPLIT + (ci/8)*LBO + row*16 + (ci%8)*2  (row = t + 3, three
// zero rows of padding either side).  Because consecutive rows are 16 bytes apart for the whole column, the
// operand of tap k is the SAME buffer with the descriptor's start address advanced by k rows: no im2col copy.
// Weights stream through a 3-stage ring (one tap = fp16 hi + lo = 16 KB) filled by cp.async.bulk + mbarrier.
//
// Precision: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) keeps 22 mantissa bits; the three products
// hi*hi + lo*hi + hi*lo drop only lo*lo (2^-22 relative) and accumulate in float32.  Weights are pre-scaled
// by a power of two so their low parts stay normal fp16.  Activations beyond the fp16 range raise
// CNN_FLAG_RANGE (GUARDED mode recomputes those reads with the float32 kernels).
#pragma once
#include <cuda_fp16.h>

#include "cnn_kernels.cuh"

namespace wdx {

// 8 warps.  16 warps (-DWDX_TC_THREADS=512: conv1 interleaved over two warps per channel group, epilogue columns in four groups)
// make the kernel itself 3 % faster (1.60 -> 1.55 ms per 8 000 reads) but take the whole register file of every SM, so the
// prepare / argmax kernels of the other chunks no longer run next to it: CNN stage 1.93 -> 2.20 ms.  Measured, kept off.
#ifndef WDX_TC_THREADS
#define WDX_TC_THREADS 256
#endif
constexpr int TC_THREADS = WDX_TC_THREADS;
constexpr int TC_WARPS = TC_THREADS / 32;
static_assert(TC_THREADS == 256 || TC_THREADS == 512, "epilogue column split: 2 or 4 column groups of 32 / 16");
constexpr int TC_MAX_TILES = 5;                              // 128-row M tiles per read (640 hidden positions: the 18 500-sample preload)
constexpr int TC_MAX_T1 = TC_MAX_TILES * 128;
constexpr int TC_W_HALF = CNN_C * CNN_C * 2;                 // 8 192 B: one tap, one split
constexpr int TC_W_TAP = 2 * TC_W_HALF;                      // 16 384 B
constexpr int TC_STAGES = 3;
constexpr int TC_CT_N = 16;                                  // ConvTranspose as an MMA: 3 residues x 2 channels = 6 columns, N >= 16 at M = 128
constexpr int TC_CT_BLOCK = CNN_C * TC_CT_N * 2;             // 2 048 B: one row shift, one split of its operand B (K = 64, N = 16)
constexpr int TC_CT_BYTES = 3 * 2 * TC_CT_BLOCK;             // 12 288 B: shifts j = 0..2 x (hi, lo); travels as one weight-ring item
// small block: w0 [64*7], b0 [64], b1 [64], b2 [64], b3 [2 (+2 pad)], barriers 8 x u64, tmem ptr, flag
constexpr int TC_SMALL_FLOATS = CNN_C * CNN_K + 3 * CNN_C + 4;
static_assert(TC_CT_BYTES <= TC_W_TAP, "the ConvTranspose operand must fit in one ring stage");

// Shared-memory / TMEM layout for reads of up to TILES x 128 hidden positions.  The CLI preloads 11 500 samples (350
// hidden positions): three tiles = 100 KB of activations instead of 166 KB, which leaves room on the SM for the CTAs
// of the neighbouring kernels (prepare / argmax of the other chunks run next to the tensor-core kernel).
template <int TILES>
struct TcLayout {
    static constexpr int MAX_T1 = TILES * 128;
    static constexpr int ROWS = MAX_T1 + 2 * CNN_P + 2;          // rows of 16 B per k-chunk column
    static constexpr int LBO = ROWS * 16;                        // bytes between k-chunk columns of A
    static constexpr int A_SPLIT = (CNN_C / 8) * LBO;            // one split (hi or lo) of the activations
    static constexpr int A_BYTES = 2 * A_SPLIT;
    static constexpr int XS = 3 * MAX_T1 + 16;                   // padded input row (floats)
    static constexpr int TMEM_COLS = TILES * (CNN_C + TC_CT_N) <= 256 ? 256 : 512;   // power of two
    static constexpr int CT_COL0 = TILES * CNN_C;                // first TMEM column of the ConvTranspose accumulators (16 per tile)
    static constexpr int OFF_W = A_BYTES;
    static constexpr int OFF_XS = OFF_W + TC_STAGES * TC_W_TAP;
    static constexpr int OFF_SMALL = OFF_XS + XS * 4;
    static constexpr int OFF_BARS = OFF_SMALL + TC_SMALL_FLOATS * 4;
    static constexpr size_t SMEM_BYTES = OFF_BARS + 8 * 8 + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory per CTA");
    static_assert(CT_COL0 + TILES * TC_CT_N <= TMEM_COLS, "TMEM columns");
    static_assert(OFF_BARS % 8 == 0 && OFF_W % 128 == 0, "alignment");
};

struct TcArgs {
    const float* x;      // [n][T] prepared input
    int64_t n;
    CnnDims d;
    const float *w0, *b0, *b1, *b2, *b3;
    const __half* wtc;   // [2 layers][7 taps][hi, lo][tc_b_offset(co, ci)] fp16, scaled by 1 / inv_wscale
    float inv_wscale;
    const __half* wct;   // ConvTranspose operand B: [3 shifts][hi, lo][tc_ct_offset(n, ci)] fp16, scaled by 1 / inv_ctscale
    float inv_ctscale;
    float* scores;       // [n][2][To]
    uint8_t* flags;      // [n]
};

// element offset of W[co][ci] inside one 64x64 operand-B block (K-major, no swizzle: 8x8 core matrices,
// 16-byte rows, 128 B between 8-row groups, 1024 B between k-chunks)
__host__ __device__ inline size_t tc_b_offset(int co, int ci) { return (size_t)(ci / 8) * (CNN_C * 8) + (size_t)co * 8 + (ci % 8); }

// element offset of column n (= 2 * residue + channel) and input channel ci inside one K = 64, N = 16 ConvTranspose block
__host__ __device__ inline size_t tc_ct_offset(int n, int ci) { return (size_t)(ci / 8) * (TC_CT_N * 8) + (size_t)n * 8 + (ci % 8); }

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end as a launch failure, never as a hung GPU.
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    if (tc_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!tc_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tc_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, float32 accumulate, M = 128, N = 64, K = 16
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start address [0,14), leading byte offset [16,30) (between the two 16-byte k-chunks of a K-step),
// stride byte offset [32,46) (between 8-row groups), version 1 at [46,48), layout type 0 at [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (InstrDescriptor): c_format F32 (1) at [4,6), a/b format F16 (0), K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(CNN_C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t TC_IDESC_CT = (1u << 4) | ((uint32_t)(TC_CT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// one lane of a converged warp (the compiler then keeps the descriptors in uniform registers)
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}\n"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld(uint32_t taddr, uint32_t (&r)[32]) { tc_ld32(taddr, r); }
__device__ __forceinline__ void tc_ld(uint32_t taddr, uint32_t (&r)[16]) { tc_ld16(taddr, r); }

// fp16 hi/lo split of 8 consecutive channels -> two 16-byte vectors
__device__ __forceinline__ void tc_split8(const float (&v)[8], uint4* hi, uint4* lo, bool* range) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float a = v[2 * q], b = v[2 * q + 1];
        *range |= (a > 65000.0f) || (b > 65000.0f) || (a != a) || (b != b);
        const __half2 hh = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
        h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *hi = make_uint4(h[0], h[1], h[2], h[3]);
    *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int TILES>
__global__ void __launch_bounds__(TC_THREADS, 1) cnn_tc_kernel(const __grid_constant__ TcArgs a) {
    using LY = TcLayout<TILES>;
    constexpr int TC_ROWS = LY::ROWS, TC_LBO = LY::LBO, TC_A_SPLIT = LY::A_SPLIT, TC_A_BYTES = LY::A_BYTES, TC_TMEM_COLS = LY::TMEM_COLS,
                  TC_CT_COL0 = LY::CT_COL0, TC_OFF_W = LY::OFF_W, TC_OFF_XS = LY::OFF_XS, TC_OFF_SMALL = LY::OFF_SMALL,
                  TC_OFF_BARS = LY::OFF_BARS;
    extern __shared__ __align__(128) unsigned char tc_sm[];
    unsigned char* A = tc_sm;                                    // activations: [2 splits][8 k-chunks][TC_ROWS][16 B]
    unsigned char* W = tc_sm + TC_OFF_W;                          // weight ring
    float* xs = reinterpret_cast<float*>(tc_sm + TC_OFF_XS);      // xs[j] = x[j - 3], zero padded
    float* w0_s = reinterpret_cast<float*>(tc_sm + TC_OFF_SMALL);  // [64][7]
    float* b0_s = w0_s + CNN_C * CNN_K;
    float* b1_s = b0_s + CNN_C;
    float* b2_s = b1_s + CNN_C;
    float* b3_s = b2_s + CNN_C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_sm + TC_OFF_BARS);
    uint64_t* full = bars;             // [3] weights of a tap have landed
    uint64_t* empty = bars + 3;        // [3] the MMAs that read the stage have completed
    uint64_t* layer_done = bars + 6;   // all MMAs of a layer have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    uint32_t* range_flag = tmem_slot + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const CnnDims d = a.d;
    const int T1 = d.T1;
    const int m_tiles = (T1 + 127) >> 7;   // 128-row M tiles that hold hidden positions (3 of 5 at the CLI's preload size)
    const int q_tiles = (T1 + 128) >> 7;   // tiles of the ConvTranspose rows q = 0 .. T1 (<= TILES: the host requires T1 < TILES * 128)

    // ---- one-time setup ------------------------------------------------------------------------------
    for (int i = tid; i < CNN_C * CNN_K; i += TC_THREADS) w0_s[i] = a.w0[i];
    for (int i = tid; i < CNN_C; i += TC_THREADS) {
        b0_s[i] = a.b0[i];
        b1_s[i] = a.b1[i];
        b2_s[i] = a.b2[i];
    }
    if (tid < 2) b3_s[tid] = a.b3[tid];
    for (int i = tid; i < TC_A_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(A)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) {
            tc_mbar_init(&full[s], 1);
            tc_mbar_init(&empty[s], 1);
        }
        tc_mbar_init(layer_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t a_base = tc_smem_u32(A), w_base = tc_smem_u32(W);

    uint32_t p_item = 0, c_item = 0, done_phase = 0;  // ring positions of the producer / MMA thread, phase of layer_done

    // input row of a read into xs (zero padded), by `nthr` threads of which this one is number `t0`
    auto load_xs = [&](int64_t rd, int t0, int nthr) {
        const float* xr = a.x + rd * d.T;
        for (int j0 = t0; j0 < 3 * T1 + 8; j0 += 4 * nthr) {   // four loads in flight per thread before the first use
            float xv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = j0 + u * nthr - CNN_P;
                xv[u] = (i >= 0 && i < d.T) ? __ldg(xr + i) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u * nthr;
                if (j < 3 * T1 + 8) xs[j] = xv[u];
            }
        }
    };
    // Weight ring, producer side (lane 0 of warp 1): item `it` of a read = tap it % 7 of layer it / 7, item 14 = the
    // ConvTranspose operand.  The producer runs three items AHEAD of the layer it is in — the first taps of a layer are
    // requested while the previous layer's last MMAs and its epilogue still run (for layer 0: under the input load and
    // conv1) — so no layer starts by waiting for its weights.
    auto issue_item = [&](int it) {
        const uint32_t s = p_item % TC_STAGES, ph = (p_item / TC_STAGES) & 1;
        tc_mbar_wait(&empty[s], ph ^ 1);
        if (it < 2 * CNN_K) {
            tc_mbar_expect_tx(&full[s], TC_W_TAP);
            tc_bulk_load(W + s * TC_W_TAP, a.wtc + (size_t)it * (TC_W_TAP / 2), TC_W_TAP, &full[s]);
        } else {
            tc_mbar_expect_tx(&full[s], TC_CT_BYTES);
            tc_bulk_load(W + s * TC_W_TAP, a.wct, TC_CT_BYTES, &full[s]);
        }
        p_item++;
    };

    if ((int64_t)blockIdx.x < a.n) load_xs(blockIdx.x, tid, TC_THREADS);
    for (int64_t read = blockIdx.x; read < a.n; read += gridDim.x) {
        // ---- conv1 -> activations (fp16 hi + lo); the input row is in xs already (prefetched under the previous read's MMAs).
        // The padding rows [0,3) and [T1+3, ROWS) of the activation buffer stay zero from the one-time clearing: conv1
        // writes rows 3 .. T1+2, the epilogues write zeros to the rows beyond T1 of the tiles they cover.
        if (tid == 0) *range_flag = 0;
        if (warp == 1 && lane == 0) {
            issue_item(0);
            issue_item(1);
            issue_item(2);
        }
        __syncthreads();
        bool range = false;
        {   // warp w owns channels 8 (w & 7) .. + 7 (weights in registers); lanes walk the time positions, the warps of a
            // channel group interleaved in blocks of 32
            const int cg = warp & 7;
            float wr[8][CNN_K], br[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                br[q] = b0_s[cg * 8 + q];
#pragma unroll
                for (int k = 0; k < CNN_K; k++) wr[q][k] = w0_s[(cg * 8 + q) * CNN_K + k];
            }
            unsigned char* col = A + cg * TC_LBO + CNN_P * 16;
            for (int t = lane + 32 * (warp >> 3); t < T1; t += 32 * (TC_WARPS / 8)) {
                float xv[CNN_K];
#pragma unroll
                for (int k = 0; k < CNN_K; k++) xv[k] = xs[t * CNN_S + k];
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float acc = br[q];
#pragma unroll
                    for (int k = 0; k < CNN_K; k++) acc = fmaf(xv[k], wr[q][k], acc);
                    v[q] = cnn_relu(acc);
                }
                uint4 hi, lo;
                tc_split8(v, &hi, &lo, &range);
                *reinterpret_cast<uint4*>(col + t * 16) = hi;
                *reinterpret_cast<uint4*>(col + t * 16 + TC_A_SPLIT) = lo;
            }
        }
        tc_fence_proxy_async();  // generic-proxy stores above are read by the tensor core through the async proxy
        __syncthreads();

        for (int layer = 0; layer < 2; layer++) {
            if (warp == 1) {  // ---- weight producer: the rest of this layer's taps and the first three items of what follows ----
                if (lane == 0) {
                    const int last = min(layer * CNN_K + CNN_K + 2, 2 * CNN_K);
                    for (int it = layer * CNN_K + 3; it <= last; it++) issue_item(it);
                }
                __syncwarp();
            } else if (layer == 1 && warp >= 2) {  // ---- the next read's input row, under this layer's MMAs (xs is idle since conv1) ----
                if (read + gridDim.x < a.n) load_xs(read + gridDim.x, tid - 64, TC_THREADS - 64);
            }
            if (warp == 0) {  // ---- MMA issuer: the warp walks the loops, one elected lane issues ----
                tc_fence_after();
                // descriptors advance in 16-byte units inside the 14-bit start-address field (no carry: smem < 256 KB)
                const uint64_t a_desc0 = tc_desc(a_base, TC_LBO, 128);
                const uint64_t b_desc0 = tc_desc(w_base, CNN_C * 16, 128);
                for (int tap = 0; tap < CNN_K; tap++, c_item++) {
                    const uint32_t s = c_item % TC_STAGES, ph = (c_item / TC_STAGES) & 1;
                    tc_mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (tc_elect_one()) {
                        const uint64_t bd0 = b_desc0 + (uint64_t)(s * (TC_W_TAP / 16));
#pragma unroll 1
                        for (int m = 0; m < m_tiles; m++) {
                            const uint64_t ad0 = a_desc0 + (uint64_t)(m * 128 + tap);
                            const uint32_t dcol = tmem + m * CNN_C;
#pragma unroll
                            for (int prod = 0; prod < 3; prod++) {  // hi*hi, lo*hi, hi*lo
#pragma unroll
                                for (int ks = 0; ks < CNN_C / 16; ks++) {
                                    const uint64_t ad = ad0 + (uint64_t)((prod == 1 ? TC_A_SPLIT / 16 : 0) + 2 * ks * (TC_LBO / 16));
                                    const uint64_t bd = bd0 + (uint64_t)((prod == 2 ? TC_W_HALF / 16 : 0) + 2 * ks * (CNN_C * 16 / 16));
                                    tc_mma_f16(dcol, ad, bd, TC_IDESC, (tap | prod | ks) != 0);
                                }
                            }
                        }
                        tc_commit(&empty[s]);  // frees the ring stage once these MMAs have read it
                        if (tap == CNN_K - 1) tc_commit(layer_done);
                    }
                    __syncwarp();
                }
            }
            tc_mbar_wait(layer_done, done_phase);
            done_phase ^= 1;
            tc_fence_after();

            // ---- epilogue: TMEM -> registers -> bias, ReLU -> next operand ------------------------------
            constexpr int EC = CNN_C / (TC_WARPS / 4);   // columns per warp: 32 (8 warps) or 16 (16 warps)
            const int g = warp & 3, hc = warp >> 2;     // TMEM lane group of this warp, column group
            const float* bias = (layer == 0 ? b1_s : b2_s) + hc * EC;
            const float sc = a.inv_wscale;
#pragma unroll 1
            for (int m = 0; m < m_tiles; m++) {
                uint32_t r[EC];
                tc_ld(tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)(m * CNN_C + hc * EC), r);
                const int t = m * 128 + g * 32 + lane;
#pragma unroll
                for (int c4 = 0; c4 < EC / 8; c4++) {   // both layers: the result is the next MMA's operand A (fp16 hi + lo, in place)
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float acc = fmaf(__uint_as_float(r[c4 * 8 + q]), sc, bias[c4 * 8 + q]);
                        v[q] = (t < T1) ? cnn_relu(acc) : 0.0f;
                    }
                    uint4 hi, lo;
                    tc_split8(v, &hi, &lo, &range);
                    unsigned char* p = A + (hc * (EC / 8) + c4) * TC_LBO + (t + CNN_P) * 16;
                    *reinterpret_cast<uint4*>(p) = hi;
                    *reinterpret_cast<uint4*>(p + TC_A_SPLIT) = lo;
                }
            }
            tc_fence_before();
            tc_fence_proxy_async();
            __syncthreads();
        }

        // ---- ConvTranspose1d (64 -> 2, stride 3) on the tensor cores ---------------------------------------------
        //   out[c][3q - 3 + r] = b[c] + sum_{j = 0..2} sum_ci h[q - j][ci] * w[ci][c][r + 3j]      (tap r + 3j <= 6)
        // = three row-shifted GEMMs  D[q][2r + c] += A[rows q - j] * B_j  with K = 64, N = 16 (6 columns used): the operand of
        // shift j is the activation buffer with the descriptor's start moved back j rows, exactly like a convolution tap.
        if (warp == 0) {   // (the operand B blocks were requested by the producer during layer 1: ring item 14)
            tc_fence_after();
            const uint32_t s = c_item % TC_STAGES, ph = (c_item / TC_STAGES) & 1;
            tc_mbar_wait(&full[s], ph);
            tc_fence_after();
            if (tc_elect_one()) {
                const uint64_t a_desc0 = tc_desc(a_base, TC_LBO, 128);
                const uint64_t b_desc0 = tc_desc(w_base + s * TC_W_TAP, TC_CT_N * 16, 128);
#pragma unroll 1
                for (int m = 0; m < q_tiles; m++) {
                    const uint32_t dcol = tmem + TC_CT_COL0 + m * TC_CT_N;
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const uint64_t ad0 = a_desc0 + (uint64_t)(m * 128 + CNN_P - j);
#pragma unroll
                        for (int prod = 0; prod < 3; prod++) {  // hi*hi, lo*hi, hi*lo
#pragma unroll
                            for (int ks = 0; ks < CNN_C / 16; ks++) {
                                const uint64_t ad = ad0 + (uint64_t)((prod == 1 ? TC_A_SPLIT / 16 : 0) + 2 * ks * (TC_LBO / 16));
                                const uint64_t bd = b_desc0 + (uint64_t)(((j * 2 + (prod == 2 ? 1 : 0)) * TC_CT_BLOCK) / 16 +
                                                                         2 * ks * (TC_CT_N * 16 / 16));
                                tc_mma_f16(dcol, ad, bd, TC_IDESC_CT, (j | prod | ks) != 0);
                            }
                        }
                    }
                }
                tc_commit(&empty[s]);
                tc_commit(layer_done);
            }
            c_item++;
            __syncwarp();
        }
        tc_mbar_wait(layer_done, done_phase);
        done_phase ^= 1;
        tc_fence_after();
        if (range) atomicOr(range_flag, 1u);
        {
            float* s0 = a.scores + (read * 2 + 0) * d.To;
            float* s1 = a.scores + (read * 2 + 1) * d.To;
            const float bb0 = b3_s[0], bb1 = b3_s[1], sct = a.inv_ctscale;
            const int g = warp & 3;
            for (int m = warp >> 2; m < q_tiles; m += TC_WARPS / 4) {
                uint32_t r[8];
                tc_ld8(tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)(TC_CT_COL0 + m * TC_CT_N), r);
                const int q = m * 128 + g * 32 + lane;
                if (q <= T1) {
#pragma unroll
                    for (int rr = 0; rr < 3; rr++) {
                        const int u = 3 * q - 3 + rr;
                        if (u >= 0 && u < d.To) {
                            s0[u] = fmaf(__uint_as_float(r[2 * rr]), sct, bb0);
                            s1[u] = fmaf(__uint_as_float(r[2 * rr + 1]), sct, bb1);
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0 && *range_flag && a.flags) a.flags[read] |= CNN_FLAG_RANGE;
    }

    // ---- teardown ----------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}

}  // namespace wdx
