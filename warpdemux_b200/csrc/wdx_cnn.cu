// wdx_cnn.cu — C-ABI entry points of the boundary-CNN stage (include/wdx_b200.h: wdx_cnn_*):
// handle with the repacked weights, staging of host minibatches, launch of the preparation,
// convolution and boundary kernels.  No CPU compute path.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda_fp16.h>

#include "cnn_kernels.cuh"
#include "cnn_tc_kernel.cuh"
#include "wdx_internal.cuh"

using namespace wdx;

struct wdx_cnn {
    int device = 0, sm_count = 148;
    int min_obs = 0, max_obs = 0, factor = 0, topk = 0;
    double guard = 1e-3;
    std::mutex mu;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t pre_stream = nullptr, post_stream = nullptr;   // prepare / argmax of the other chunks next to the tensor-core kernel
    cudaStream_t conv_stream = nullptr;                         // ... which runs at the highest stream priority: its CTAs are placed first
    cudaEvent_t ev_start = nullptr;
    std::vector<cudaEvent_t> ev_chunk;                           // three per chunk: prepared, convolved, reduced
    DevBuf scores2;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    // weights on the device
    DevBuf w0, b0, wt1, b1, wt2, b2, wT, b3;  // float32 (EXACT mode and the first / last layer of both modes)
    DevBuf wtc;                                // fp16 hi/lo split weights of the two 64->64 layers (FAST mode)
    float tc_wscale = 1.0f;
    DevBuf wct;                                // fp16 hi/lo split ConvTranspose operand (FAST mode), 3 row shifts
    float tc_ctscale = 1.0f;
    // workspaces (grow-only)
    DevBuf sig[2], x, hA, hB, scores, masked, a_end, p_end, margin, cand, n_cand, flags, preds, redo_idx, redo_cnt, xg, sg;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev;
    size_t tev_used = 0;
    int conv_smem_ok = 0, tc_smem_ok = 0;
};

namespace {

int upload_f32(DevBuf& b, const std::vector<float>& h) {
    int rc = b.reserve(std::max<size_t>(16, h.size() * 4));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(b.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    return WDX_OK;
}

struct Timer {
    wdx_cnn* c;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    int begin() {
        if (!c->timing) return WDX_OK;
        if (c->tev_used == c->tev.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            c->tev.emplace_back(a, b);
        }
        CUDA_TRY(cudaEventRecord(c->tev[c->tev_used].first, st));
        e1 = c->tev[c->tev_used].second;
        c->tev_used++;
        return WDX_OK;
    }
    int end() {
        if (e1) CUDA_TRY(cudaEventRecord(e1, st));
        e1 = nullptr;
        return WDX_OK;
    }
};

int make_dims(const wdx_cnn* c, int64_t stride, CnnDims* d) {
    if (stride <= c->min_obs) return fail(WDX_ERR_INVALID, "row stride %lld <= min_obs_adapter %d", (long long)stride, c->min_obs);
    const int64_t T = (stride - c->min_obs + c->factor - 1) / c->factor;
    if (T < CNN_K || T > CNN_MAX_T) return fail(WDX_ERR_UNSUPPORTED, "downscaled length %lld outside [%d,%d]", (long long)T, CNN_K, CNN_MAX_T);
    d->min_obs = c->min_obs;
    d->factor = c->factor;
    d->span = (c->max_obs - c->min_obs) / c->factor;
    d->topk = c->topk;
    d->T = (int)T;
    d->T1 = (d->T - 1) / CNN_S + 1;
    d->To = CNN_S * d->T1 - 2;
    return WDX_OK;
}

// x [cn][T] -> scores [cn][2][To], float32 CUDA-core path
constexpr int64_t ASYNC_REDO_MAX_READS = 4096;
bool async_redo_disabled() {
    static const bool off = [] { const char* e = getenv("WDX_CNN_SYNC_REDO"); return e && e[0] == '1'; }();
    return off;
}

int forward_exact(wdx_cnn* c, const float* x, int64_t cn, const CnnDims& d, float* scores, cudaStream_t st,
                  const int32_t* n_dev = nullptr) {
    int rc;
    if ((rc = c->hA.reserve((size_t)cn * d.T1 * CNN_C * 4))) return rc;
    if ((rc = c->hB.reserve((size_t)cn * d.T1 * CNN_C * 4))) return rc;
    float* hA = (float*)c->hA.p;
    float* hB = (float*)c->hB.p;
    Timer tm{c, st};
    if ((rc = tm.begin())) return rc;
    {
        dim3 grid((unsigned)std::min(n_dev ? 8 : 64, (d.T1 + 3) / 4), (unsigned)cn);   // few, mostly idle CTAs in the device-counted re-run
        cnn_conv1_f32_kernel<<<grid, 256, 0, st>>>(x, (const float*)c->w0.p, (const float*)c->b0.p, d, hA, n_dev);
        CUDA_TRY(cudaGetLastError());
    }
    const int tiles_per_read = (d.T1 + CV_TT - 1) / CV_TT;
    const int64_t n_tiles = cn * tiles_per_read;
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, c->sm_count);
    cnn_conv64_f32_kernel<<<grid, 256, cnn_conv64_smem_bytes(), st>>>(hA, hB, (const float*)c->wt1.p, (const float*)c->b1.p, d.T1,
                                                                     tiles_per_read, n_tiles, n_dev);
    CUDA_TRY(cudaGetLastError());
    cnn_conv64_f32_kernel<<<grid, 256, cnn_conv64_smem_bytes(), st>>>(hB, hA, (const float*)c->wt2.p, (const float*)c->b2.p, d.T1,
                                                                     tiles_per_read, n_tiles, n_dev);
    CUDA_TRY(cudaGetLastError());
    {
        dim3 g2((unsigned)((d.To + 127) / 128), (unsigned)cn);
        cnn_convT_f32_kernel<<<g2, 128, 0, st>>>(hA, (const float*)c->wT.p, (const float*)c->b3.p, d, scores, n_dev);
        CUDA_TRY(cudaGetLastError());
    }
    g_launches += 4;
    return tm.end();
}

int forward_fast(wdx_cnn* c, const float* x, int64_t cn, const CnnDims& d, float* scores, uint8_t* flags, cudaStream_t st) {
    if (!c->tc_smem_ok) return fail(WDX_ERR_UNSUPPORTED, "tensor-core CNN kernel unavailable on this device");
    if (d.T1 >= TC_MAX_T1) return fail(WDX_ERR_UNSUPPORTED, "hidden length %d >= %d: use WDX_CNN_EXACT_F32", d.T1, TC_MAX_T1);
    Timer tm{c, st};
    int rc;
    if ((rc = tm.begin())) return rc;
    const unsigned grid = (unsigned)std::min<int64_t>(cn, c->sm_count);
    TcArgs a{};
    a.x = x;
    a.n = cn;
    a.d = d;
    a.w0 = (const float*)c->w0.p;
    a.b0 = (const float*)c->b0.p;
    a.b1 = (const float*)c->b1.p;
    a.b2 = (const float*)c->b2.p;
    a.b3 = (const float*)c->b3.p;
    a.wtc = (const __half*)c->wtc.p;
    a.inv_wscale = 1.0f / c->tc_wscale;
    a.wct = (const __half*)c->wct.p;
    a.inv_ctscale = 1.0f / c->tc_ctscale;
    a.scores = scores;
    a.flags = flags;
    if (d.T1 < TcLayout<3>::MAX_T1) cnn_tc_kernel<3><<<grid, TC_THREADS, TcLayout<3>::SMEM_BYTES, st>>>(a);
    else cnn_tc_kernel<TC_MAX_TILES><<<grid, TC_THREADS, TcLayout<TC_MAX_TILES>::SMEM_BYTES, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    g_launches += 1;
    return tm.end();
}

// find_peaks on the flattened batch + per-read top-k + row assignment (cnn.py:137-160, 176-181)
int finish_predict(wdx_cnn* c, int64_t n, const CnnDims& d, uint8_t* flags_d, int64_t* preds_d, cudaStream_t st) {
    cnn_peaks_kernel<<<(unsigned)n, PK_THREADS, 0, st>>>((const float*)c->masked.p, n, d, 5, (int32_t*)c->cand.p, (int32_t*)c->n_cand.p,
                                                        flags_d);
    CUDA_TRY(cudaGetLastError());
    cnn_rows_kernel<<<1, 1024, 0, st>>>((const int32_t*)c->a_end.p, (const int32_t*)c->cand.p, (const int32_t*)c->n_cand.p, n, d, preds_d);
    CUDA_TRY(cudaGetLastError());
    g_launches += 2;
    return WDX_OK;
}

int reserve_predict(wdx_cnn* c, int64_t n, const CnnDims& d) {
    int rc;
    if ((rc = c->masked.reserve((size_t)n * d.To * 4)) || (rc = c->a_end.reserve((size_t)n * 4)) || (rc = c->p_end.reserve((size_t)n * 4)) ||
        (rc = c->margin.reserve((size_t)n * 4)) || (rc = c->cand.reserve((size_t)n * d.topk * 4)) || (rc = c->n_cand.reserve((size_t)n * 4)))
        return rc;
    return WDX_OK;
}

}  // namespace

extern "C" {

int wdx_cnn_create(const wdx_cnn_config* cfg, const float* w0, const float* b0, const float* w1, const float* b1,
                   const float* w2, const float* b2, const float* w3, const float* b3, int device, wdx_cnn** out) {
    if (!out) return fail(WDX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!cfg || !w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3) return fail(WDX_ERR_INVALID, "NULL argument");
    if (cfg->channels != CNN_C || cfg->kernel_size != CNN_K)
        return fail(WDX_ERR_UNSUPPORTED, "only BoundariesCNN(channels=64, kernel_size=7) is implemented (got %d, %d)", cfg->channels,
                    cfg->kernel_size);
    if (cfg->min_obs_adapter < 0 || cfg->max_obs_adapter <= cfg->min_obs_adapter || cfg->downscale_factor < 1 ||
        cfg->downscale_factor > 128)
        return fail(WDX_ERR_INVALID, "bad core configuration");
    if (cfg->polya_cand_k < 2 || cfg->polya_cand_k > CNN_MAX_TOPK)
        return fail(WDX_ERR_UNSUPPORTED, "polya_cand_k=%d outside [2,%d]", cfg->polya_cand_k, CNN_MAX_TOPK);
    int ndev = wdx_device_count();
    if (ndev <= 0) return fail(WDX_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(WDX_ERR_INVALID, "device %d of %d", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    wdx_cnn* c = new (std::nothrow) wdx_cnn();
    if (!c) return fail(WDX_ERR_NOMEM, "host allocation failed");
    c->device = device;
    c->min_obs = cfg->min_obs_adapter;
    c->max_obs = cfg->max_obs_adapter;
    c->factor = cfg->downscale_factor;
    c->topk = cfg->polya_cand_k;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    auto bail = [&](int rc) {
        wdx_cnn_destroy(c);
        return rc;
    };
    int rc;
    // repack: conv weights torch [co][ci][k] -> [k][ci][co]; transposed conv torch [ci][c][k] -> [k][ci][c]
    auto repack = [](const float* w) {
        std::vector<float> o((size_t)CNN_K * CNN_C * CNN_C);
        for (int co = 0; co < CNN_C; co++)
            for (int ci = 0; ci < CNN_C; ci++)
                for (int k = 0; k < CNN_K; k++) o[((size_t)k * CNN_C + ci) * CNN_C + co] = w[((size_t)co * CNN_C + ci) * CNN_K + k];
        return o;
    };
    std::vector<float> wT((size_t)CNN_K * CNN_C * 2);
    for (int ci = 0; ci < CNN_C; ci++)
        for (int ch = 0; ch < 2; ch++)
            for (int k = 0; k < CNN_K; k++) wT[((size_t)k * CNN_C + ci) * 2 + ch] = w3[((size_t)ci * 2 + ch) * CNN_K + k];
    if ((rc = upload_f32(c->w0, std::vector<float>(w0, w0 + CNN_C * CNN_K))) || (rc = upload_f32(c->b0, std::vector<float>(b0, b0 + CNN_C))) ||
        (rc = upload_f32(c->wt1, repack(w1))) || (rc = upload_f32(c->b1, std::vector<float>(b1, b1 + CNN_C))) ||
        (rc = upload_f32(c->wt2, repack(w2))) || (rc = upload_f32(c->b2, std::vector<float>(b2, b2 + CNN_C))) ||
        (rc = upload_f32(c->wT, wT)) || (rc = upload_f32(c->b3, std::vector<float>(b3, b3 + 2))))
        return bail(rc);
    {   // FAST mode operand B: per layer, per tap, fp16 high and low parts of W * 2^s in the canonical
        // no-swizzle K-major core-matrix layout (cnn_tc_kernel.cuh)
        float wmax = 0.f;
        for (size_t i = 0; i < (size_t)CNN_C * CNN_C * CNN_K; i++) wmax = std::max(wmax, std::max(std::fabs(w1[i]), std::fabs(w2[i])));
        int e = 0;
        if (wmax > 0.f) e = (int)std::floor(std::log2(1024.0f / wmax));  // |W| * 2^e <= 1024: low parts stay normal fp16
        e = std::max(-8, std::min(14, e));
        c->tc_wscale = std::ldexp(1.0f, e);
        std::vector<__half> hw((size_t)2 * CNN_K * 2 * CNN_C * CNN_C);
        const float* ws[2] = {w1, w2};
        for (int layer = 0; layer < 2; layer++)
            for (int k = 0; k < CNN_K; k++)
                for (int co = 0; co < CNN_C; co++)
                    for (int ci = 0; ci < CNN_C; ci++) {
                        const float v = ws[layer][((size_t)co * CNN_C + ci) * CNN_K + k] * c->tc_wscale;
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        const size_t base = ((size_t)(layer * CNN_K + k) * 2) * CNN_C * CNN_C;
                        const size_t off = tc_b_offset(co, ci);
                        hw[base + off] = hi;
                        hw[base + (size_t)CNN_C * CNN_C + off] = lo;
                    }
        if ((rc = c->wtc.reserve(hw.size() * sizeof(__half)))) return bail(rc);
        if (cudaMemcpy(c->wtc.p, hw.data(), hw.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(WDX_ERR_CUDA, "weight upload failed"));
    }
    {   // FAST mode ConvTranspose as three row-shifted GEMMs (cnn_tc_kernel.cuh): operand B of shift j holds, in column
        // n = 2 * r + ch, the tap r + 3 j of torch's w3 [ci][ch][k] (zero where r + 3 j > 6 and in the padding columns)
        float wmax = 0.f;
        for (size_t i = 0; i < (size_t)CNN_C * 2 * CNN_K; i++) wmax = std::max(wmax, std::fabs(w3[i]));
        int e = 0;
        if (wmax > 0.f) e = (int)std::floor(std::log2(1024.0f / wmax));
        e = std::max(-8, std::min(14, e));
        c->tc_ctscale = std::ldexp(1.0f, e);
        std::vector<__half> hc((size_t)TC_CT_BYTES / 2, __float2half_rn(0.0f));
        for (int j = 0; j < 3; j++)
            for (int r = 0; r < 3; r++) {
                const int k = r + 3 * j;
                if (k >= CNN_K) continue;
                for (int ch = 0; ch < 2; ch++)
                    for (int ci = 0; ci < CNN_C; ci++) {
                        const float v = w3[((size_t)ci * 2 + ch) * CNN_K + k] * c->tc_ctscale;
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        const size_t off = tc_ct_offset(2 * r + ch, ci);
                        hc[(size_t)(j * 2 + 0) * (TC_CT_BLOCK / 2) + off] = hi;
                        hc[(size_t)(j * 2 + 1) * (TC_CT_BLOCK / 2) + off] = lo;
                    }
            }
        if ((rc = c->wct.reserve(hc.size() * sizeof(__half)))) return bail(rc);
        if (cudaMemcpy(c->wct.p, hc.data(), hc.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(WDX_ERR_CUDA, "weight upload failed"));
    }
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (cudaFuncSetAttribute((const void*)cnn_conv64_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)cnn_conv64_smem_bytes()) != cudaSuccess)
        return bail(fail(WDX_ERR_CUDA, "cudaFuncSetAttribute(conv64) failed: %s", cudaGetErrorString(cudaGetLastError())));
    c->conv_smem_ok = 1;
    if ((int)TcLayout<TC_MAX_TILES>::SMEM_BYTES <= optin &&
        cudaFuncSetAttribute((const void*)cnn_tc_kernel<TC_MAX_TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)TcLayout<TC_MAX_TILES>::SMEM_BYTES) == cudaSuccess &&
        cudaFuncSetAttribute((const void*)cnn_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)TcLayout<3>::SMEM_BYTES) == cudaSuccess)
        c->tc_smem_ok = 1;
    else
        cudaGetLastError();
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->pre_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->post_stream, cudaStreamNonBlocking) == cudaSuccess &&
              [&] {
                  int least = 0, greatest = 0;
                  cudaDeviceGetStreamPriorityRange(&least, &greatest);
                  return cudaStreamCreateWithPriority(&c->conv_stream, cudaStreamNonBlocking, greatest) == cudaSuccess;
              }() &&
              cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return bail(fail(WDX_ERR_CUDA, "stream/event creation failed"));
    *out = c;
    return WDX_OK;
}

void wdx_cnn_destroy(wdx_cnn* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (DevBuf* b : {&c->w0, &c->b0, &c->wt1, &c->b1, &c->wt2, &c->b2, &c->wT, &c->b3, &c->wtc, &c->wct, &c->sig[0], &c->sig[1], &c->x, &c->hA,
                      &c->hB, &c->scores, &c->masked, &c->a_end, &c->p_end, &c->margin, &c->cand, &c->n_cand, &c->flags, &c->preds,
                      &c->redo_idx, &c->redo_cnt, &c->xg, &c->sg})
        b->release();
    for (int i = 0; i < 2; i++) {
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    for (auto& e : c->tev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    for (cudaStream_t q : {c->pre_stream, c->post_stream, c->conv_stream})
        if (q) {
            cudaStreamSynchronize(q);
            cudaStreamDestroy(q);
        }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    for (cudaEvent_t e : c->ev_chunk) cudaEventDestroy(e);
    c->scores2.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int wdx_cnn_score_len(wdx_cnn* c, int64_t stride, int32_t* t_in, int32_t* t_out) {
    if (!c) return fail(WDX_ERR_INVALID, "NULL CNN handle");
    CnnDims d;
    int rc = make_dims(c, stride, &d);
    if (rc) return rc;
    if (t_in) *t_in = d.T;
    if (t_out) *t_out = d.To;
    return WDX_OK;
}

int wdx_cnn_set_guard(wdx_cnn* c, double guard) {
    if (!c || !(guard >= 0)) return fail(WDX_ERR_INVALID, "bad guard");
    c->guard = guard;
    return WDX_OK;
}

int wdx_cnn_detect(wdx_cnn* c, const float* signals, int64_t n, int64_t stride, int mode, int64_t* preds, float* scores,
                   uint8_t* flags, void* stream) {
    if (!c) return fail(WDX_ERR_INVALID, "NULL CNN handle");
    if (n < 0) return fail(WDX_ERR_INVALID, "n=%lld", (long long)n);
    if (n == 0) return WDX_OK;
    if (!signals || !preds) return fail(WDX_ERR_INVALID, "signals and preds are required");
    if (mode < WDX_CNN_EXACT_F32 || mode > WDX_CNN_GUARDED) return fail(WDX_ERR_INVALID, "mode=%d", mode);
    CnnDims d;
    int rc = make_dims(c, stride, &d);
    if (rc) return rc;
    if ((int64_t)n * d.To > (int64_t)0x7ffffff0 * 64) return fail(WDX_ERR_UNSUPPORTED, "batch too large");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->device));
    c->tev_used = 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const bool sig_dev = mem_kind(signals) == 2;
    const bool preds_dev = mem_kind(preds) == 2;
    const bool scores_dev = scores && mem_kind(scores) == 2;
    const bool flags_dev = flags && mem_kind(flags) == 2;
    const int ld = 1 + d.topk;

    const int64_t chunk = std::min<int64_t>(n, 1024);
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    // Device-resident batches of several chunks: the prepare kernels run ahead on their own stream and the argmax kernels
    // behind on another, next to the convolution kernel of the neighbouring chunks (they fit beside its CTA on every SM)
    // instead of between them on one stream.
    const bool overlap = sig_dev && n_chunks > 1 && (!scores || scores_dev);
    const bool keep_x = mode == WDX_CNN_GUARDED || overlap;
    if ((rc = c->x.reserve((size_t)(keep_x ? n : chunk) * d.T * 4))) return rc;
    if (!scores_dev && (rc = c->scores.reserve((size_t)chunk * 2 * d.To * 4))) return rc;
    if (overlap && !scores_dev && (rc = c->scores2.reserve((size_t)chunk * 2 * d.To * 4))) return rc;
    if (overlap) {
        while ((int64_t)c->ev_chunk.size() < 3 * n_chunks) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->ev_chunk.push_back(e);
        }
    }
    if ((rc = reserve_predict(c, n, d))) return rc;
    if (!flags_dev && (rc = c->flags.reserve((size_t)n))) return rc;
    if (!preds_dev && (rc = c->preds.reserve((size_t)n * ld * 8))) return rc;
    uint8_t* flags_d = flags_dev ? flags : (uint8_t*)c->flags.p;
    int64_t* preds_d = preds_dev ? preds : (int64_t*)c->preds.p;
    CUDA_TRY(cudaMemsetAsync(flags_d, 0, (size_t)n, st));
    const int nbuf = n_chunks > 1 ? 2 : 1;
    if (!sig_dev)
        for (int b = 0; b < nbuf; b++)
            if ((rc = c->sig[b].reserve((size_t)chunk * stride * 4))) return rc;

    auto stage_in = [&](int64_t ci) -> int {
        const int b = (int)(ci & 1);
        const int64_t r0 = ci * chunk, cn = std::min(chunk, n - r0);
        CUDA_TRY(cudaEventSynchronize(c->ev_free[b]));
        CUDA_TRY(cudaMemcpyAsync(c->sig[b].p, signals + (size_t)r0 * stride, (size_t)cn * stride * 4, cudaMemcpyHostToDevice, c->copy_stream));
        CUDA_TRY(cudaEventRecord(c->ev_h2d[b], c->copy_stream));
        return WDX_OK;
    };
    if (!sig_dev) {
        CUDA_TRY(cudaEventRecord(c->ev_free[0], st));
        CUDA_TRY(cudaEventRecord(c->ev_free[1], st));
        if ((rc = stage_in(0))) return rc;
    }
    if (overlap) {
        cudaStream_t cs = c->conv_stream;
        CUDA_TRY(cudaEventRecord(c->ev_start, st));
        CUDA_TRY(cudaStreamWaitEvent(c->pre_stream, c->ev_start, 0));
        CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_start, 0));
        for (int64_t ci = 0; ci < n_chunks; ci++) {        // every chunk's prepare, in order, ahead of the convolutions
            const int64_t r0 = ci * chunk, cn = std::min(chunk, n - r0);
            cnn_prepare_kernel<<<(unsigned)cn, FP_THREADS, (size_t)d.T * 4, c->pre_stream>>>(signals + (size_t)r0 * stride, stride, cn, d,
                                                                                           (float*)c->x.p + (size_t)r0 * d.T);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
            CUDA_TRY(cudaEventRecord(c->ev_chunk[3 * ci], c->pre_stream));
        }
        for (int64_t ci = 0; ci < n_chunks; ci++) {
            const int64_t r0 = ci * chunk, cn = std::min(chunk, n - r0);
            float* x_d = (float*)c->x.p + (size_t)r0 * d.T;
            float* sc_d = scores_dev ? scores + (size_t)r0 * 2 * d.To : (float*)((ci & 1) ? c->scores2.p : c->scores.p);
            CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_chunk[3 * ci], 0));
            if (!scores_dev && ci >= 2) CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_chunk[3 * (ci - 2) + 2], 0));   // the score buffer is free again
            if (mode == WDX_CNN_EXACT_F32) rc = forward_exact(c, x_d, cn, d, sc_d, cs);
            else rc = forward_fast(c, x_d, cn, d, sc_d, flags_d + r0, cs);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(c->ev_chunk[3 * ci + 1], cs));
            CUDA_TRY(cudaStreamWaitEvent(c->post_stream, c->ev_chunk[3 * ci + 1], 0));
            cnn_argmax_kernel<<<(unsigned)((cn + 3) / 4), 128, 0, c->post_stream>>>(sc_d, cn, d, (float*)c->masked.p + (size_t)r0 * d.To,
                                                                                   (int32_t*)c->a_end.p + r0, (int32_t*)c->p_end.p + r0,
                                                                                   (float*)c->margin.p + r0, flags_d + r0);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
            CUDA_TRY(cudaEventRecord(c->ev_chunk[3 * ci + 2], c->post_stream));
        }
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[3 * (n_chunks - 1) + 2], 0));
    }
    for (int64_t ci = 0; ci < (overlap ? 0 : n_chunks); ci++) {
        const int b = (int)(ci & 1);
        const int64_t r0 = ci * chunk, cn = std::min(chunk, n - r0);
        const float* sig_d = sig_dev ? signals + (size_t)r0 * stride : (const float*)c->sig[b].p;
        if (!sig_dev) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_h2d[b], 0));
        float* x_d = (float*)c->x.p + (keep_x ? (size_t)r0 * d.T : 0);
        cnn_prepare_kernel<<<(unsigned)cn, FP_THREADS, (size_t)d.T * 4, st>>>(sig_d, stride, cn, d, x_d);
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        if (!sig_dev) {
            CUDA_TRY(cudaEventRecord(c->ev_free[b], st));
            if (ci + 1 < n_chunks && (rc = stage_in(ci + 1))) return rc;
        }
        float* sc_d = scores_dev ? scores + (size_t)r0 * 2 * d.To : (float*)c->scores.p;
        if (mode == WDX_CNN_EXACT_F32) rc = forward_exact(c, x_d, cn, d, sc_d, st);
        else rc = forward_fast(c, x_d, cn, d, sc_d, flags_d + r0, st);
        if (rc) return rc;
        cnn_argmax_kernel<<<(unsigned)((cn + 3) / 4), 128, 0, st>>>(sc_d, cn, d, (float*)c->masked.p + (size_t)r0 * d.To,
                                                                   (int32_t*)c->a_end.p + r0, (int32_t*)c->p_end.p + r0,
                                                                   (float*)c->margin.p + r0, flags_d + r0);
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        if (scores && !scores_dev) {
            CUDA_TRY(cudaMemcpyAsync(scores + (size_t)r0 * 2 * d.To, sc_d, (size_t)cn * 2 * d.To * 4, cudaMemcpyDeviceToHost, st));
            if (n_chunks > 1) CUDA_TRY(cudaStreamSynchronize(st));  // the staging buffer is reused by the next chunk
        }
    }
    if (mode == WDX_CNN_GUARDED) {
        // reads whose argmax margins are inside the guard band, or whose activations left the fp16 range: EXACT_F32
        if ((rc = c->redo_idx.reserve((size_t)n * 4)) || (rc = c->redo_cnt.reserve(16))) return rc;
        CUDA_TRY(cudaMemsetAsync(c->redo_cnt.p, 0, 4, st));
        cnn_guard_list_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>(
            (const float*)c->margin.p, flags_d, n, (float)c->guard, (int32_t*)c->redo_idx.p, (int32_t*)c->redo_cnt.p);
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        int m = 0;
        const bool counted_on_device = n <= ASYNC_REDO_MAX_READS && (!scores || scores_dev) && !async_redo_disabled();
        if (counted_on_device) {
            // minibatch-sized call: the re-run is launched for the worst case (every read listed) and the kernels skip
            // what lies beyond the device-side count - no host synchronisation, the call stays asynchronous
            const int32_t* cnt = (const int32_t*)c->redo_cnt.p;
            if ((rc = c->xg.reserve((size_t)n * d.T * 4)) || (rc = c->sg.reserve((size_t)n * 2 * d.To * 4))) return rc;
            cnn_gather_rows_kernel<<<(unsigned)n, 256, 0, st>>>((const float*)c->x.p, (const int32_t*)c->redo_idx.p, d.T, (float*)c->xg.p, cnt);
            CUDA_TRY(cudaGetLastError());
            if ((rc = forward_exact(c, (const float*)c->xg.p, n, d, (float*)c->sg.p, st, cnt))) return rc;
            cnn_argmax_idx_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>((const float*)c->sg.p, (const int32_t*)c->redo_idx.p, n, d,
                                                                          (float*)c->masked.p, (int32_t*)c->a_end.p, (int32_t*)c->p_end.p,
                                                                          flags_d, scores_dev ? scores : nullptr, cnt);
            CUDA_TRY(cudaGetLastError());
            g_launches += 2;
        } else {
            CUDA_TRY(cudaMemcpyAsync(&m, c->redo_cnt.p, 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        for (int64_t g0 = 0; g0 < m; g0 += chunk) {
            const int64_t gn = std::min<int64_t>(chunk, m - g0);
            if ((rc = c->xg.reserve((size_t)gn * d.T * 4)) || (rc = c->sg.reserve((size_t)gn * 2 * d.To * 4))) return rc;
            const int32_t* idx = (const int32_t*)c->redo_idx.p + g0;
            cnn_gather_rows_kernel<<<(unsigned)gn, 256, 0, st>>>((const float*)c->x.p, idx, d.T, (float*)c->xg.p);
            CUDA_TRY(cudaGetLastError());
            if ((rc = forward_exact(c, (const float*)c->xg.p, gn, d, (float*)c->sg.p, st))) return rc;
            cnn_argmax_idx_kernel<<<(unsigned)((gn + 3) / 4), 128, 0, st>>>((const float*)c->sg.p, idx, gn, d, (float*)c->masked.p,
                                                                           (int32_t*)c->a_end.p, (int32_t*)c->p_end.p, flags_d,
                                                                           scores_dev ? scores : nullptr);
            CUDA_TRY(cudaGetLastError());
            g_launches += 2;
            if (scores && !scores_dev) {  // rare: patch the host rows one by one
                std::vector<int32_t> hidx((size_t)gn);
                CUDA_TRY(cudaMemcpyAsync(hidx.data(), idx, (size_t)gn * 4, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                for (int64_t q = 0; q < gn; q++)
                    CUDA_TRY(cudaMemcpyAsync(scores + (size_t)hidx[q] * 2 * d.To, (const float*)c->sg.p + (size_t)q * 2 * d.To,
                                             (size_t)2 * d.To * 4, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
            }
        }
    }
    if ((rc = finish_predict(c, n, d, flags_d, preds_d, st))) return rc;
    if (!preds_dev) CUDA_TRY(cudaMemcpyAsync(preds, preds_d, (size_t)n * ld * 8, cudaMemcpyDeviceToHost, st));
    if (flags && !flags_dev) CUDA_TRY(cudaMemcpyAsync(flags, flags_d, (size_t)n, cudaMemcpyDeviceToHost, st));
    const bool any_host = !sig_dev || !preds_dev || (scores && !scores_dev) || (flags && !flags_dev);
    if (any_host || !stream) CUDA_TRY(cudaStreamSynchronize(st));
    return WDX_OK;
}

int wdx_cnn_prepare(wdx_cnn* c, const float* signals, int64_t n, int64_t stride, float* x, void* stream) {
    if (!c) return fail(WDX_ERR_INVALID, "NULL CNN handle");
    if (n < 0) return fail(WDX_ERR_INVALID, "n=%lld", (long long)n);
    if (n == 0) return WDX_OK;
    if (!signals || !x) return fail(WDX_ERR_INVALID, "signals and x are required");
    CnnDims d;
    int rc = make_dims(c, stride, &d);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const bool sig_dev = mem_kind(signals) == 2, x_dev = mem_kind(x) == 2;
    const int64_t chunk = std::min<int64_t>(n, 1024);
    if (!sig_dev && (rc = c->sig[0].reserve((size_t)chunk * stride * 4))) return rc;
    if (!x_dev && (rc = c->x.reserve((size_t)chunk * d.T * 4))) return rc;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t cn = std::min(chunk, n - r0);
        const float* sig_d = signals + (size_t)r0 * stride;
        if (!sig_dev) {
            CUDA_TRY(cudaMemcpyAsync(c->sig[0].p, sig_d, (size_t)cn * stride * 4, cudaMemcpyHostToDevice, st));
            sig_d = (const float*)c->sig[0].p;
        }
        float* x_d = x_dev ? x + (size_t)r0 * d.T : (float*)c->x.p;
        cnn_prepare_kernel<<<(unsigned)cn, FP_THREADS, (size_t)d.T * 4, st>>>(sig_d, stride, cn, d, x_d);
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        if (!x_dev) {
            CUDA_TRY(cudaMemcpyAsync(x + (size_t)r0 * d.T, x_d, (size_t)cn * d.T * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    if (!sig_dev || !x_dev || !stream) CUDA_TRY(cudaStreamSynchronize(st));
    return WDX_OK;
}

int wdx_cnn_predict(wdx_cnn* c, const float* scores, int64_t n, int32_t t_out, int scaled, int64_t* preds, uint8_t* flags,
                    void* stream) {
    if (!c) return fail(WDX_ERR_INVALID, "NULL CNN handle");
    if (n < 0) return fail(WDX_ERR_INVALID, "n=%lld", (long long)n);
    if (n == 0) return WDX_OK;
    if (!scores || !preds) return fail(WDX_ERR_INVALID, "scores and preds are required");
    if (t_out < 3 || t_out > CNN_MAX_T) return fail(WDX_ERR_UNSUPPORTED, "score length %d outside [3,%d]", t_out, CNN_MAX_T);
    CnnDims d{};
    d.min_obs = scaled ? c->min_obs : 0;
    d.factor = scaled ? c->factor : 1;
    d.span = (c->max_obs - c->min_obs) / c->factor;
    d.topk = c->topk;
    d.T = d.To = t_out;
    d.T1 = 0;
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const bool sc_dev = mem_kind(scores) == 2, preds_dev = mem_kind(preds) == 2, flags_dev = flags && mem_kind(flags) == 2;
    const int ld = 1 + d.topk;
    int rc;
    if ((rc = reserve_predict(c, n, d))) return rc;
    const int64_t chunk = std::min<int64_t>(n, 4096);
    if (!sc_dev && (rc = c->scores.reserve((size_t)chunk * 2 * d.To * 4))) return rc;
    if (!flags_dev && (rc = c->flags.reserve((size_t)n))) return rc;
    if (!preds_dev && (rc = c->preds.reserve((size_t)n * ld * 8))) return rc;
    uint8_t* flags_d = flags_dev ? flags : (uint8_t*)c->flags.p;
    int64_t* preds_d = preds_dev ? preds : (int64_t*)c->preds.p;
    CUDA_TRY(cudaMemsetAsync(flags_d, 0, (size_t)n, st));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t cn = std::min(chunk, n - r0);
        const float* sc_d = scores + (size_t)r0 * 2 * d.To;
        if (!sc_dev) {
            CUDA_TRY(cudaMemcpyAsync(c->scores.p, sc_d, (size_t)cn * 2 * d.To * 4, cudaMemcpyHostToDevice, st));
            sc_d = (const float*)c->scores.p;
        }
        cnn_argmax_kernel<<<(unsigned)((cn + 3) / 4), 128, 0, st>>>(sc_d, cn, d, (float*)c->masked.p + (size_t)r0 * d.To,
                                                                   (int32_t*)c->a_end.p + r0, (int32_t*)c->p_end.p + r0,
                                                                   (float*)c->margin.p + r0, flags_d + r0);
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        if (!sc_dev && r0 + chunk < n) CUDA_TRY(cudaStreamSynchronize(st));
    }
    if ((rc = finish_predict(c, n, d, flags_d, preds_d, st))) return rc;
    if (!preds_dev) CUDA_TRY(cudaMemcpyAsync(preds, preds_d, (size_t)n * ld * 8, cudaMemcpyDeviceToHost, st));
    if (flags && !flags_dev) CUDA_TRY(cudaMemcpyAsync(flags, flags_d, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (!sc_dev || !preds_dev || (flags && !flags_dev) || !stream) CUDA_TRY(cudaStreamSynchronize(st));
    return WDX_OK;
}

int wdx_cnn_enable_timing(wdx_cnn* c, int on) {
    if (!c) return fail(WDX_ERR_INVALID, "NULL CNN handle");
    c->timing = on != 0;
    return WDX_OK;
}

int wdx_cnn_last_kernel_ms(wdx_cnn* c, double* ms, int* launches) {
    if (!c || !ms) return fail(WDX_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CUDA_TRY(cudaSetDevice(c->device));
    double tot = 0;
    for (size_t i = 0; i < c->tev_used; i++) {
        CUDA_TRY(cudaEventSynchronize(c->tev[i].second));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, c->tev[i].first, c->tev[i].second));
        tot += t;
    }
    *ms = tot;
    if (launches) *launches = (int)c->tev_used;
    return WDX_OK;
}

}  // extern "C"
