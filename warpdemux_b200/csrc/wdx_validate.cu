// wdx_validate.cu — C-ABI entry points of the boundary validation step (include/wdx_b200.h:
// wdx_validate_*): handle, staging of host buffers, launch of validate_kernel.  No CPU compute path.
#include <algorithm>
#include <cmath>
#include <new>

#include "validate_kernel.cuh"
#include "llr_kernel.cuh"
#include "wdx_internal.cuh"

using namespace wdx;

struct wdx_validate {
    ValCfg cfg{};
    int device = 0;
    int sm_count = 148;
    int smem_max = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    DevBuf sig, len, preds, success, info, bounds, vals, parts, pores, scratch, counter;
    // LLR fallback (wdx_validate_set_llr): proposed boundaries, mask, per-read median / MAD
    bool llr_on = false;
    LlrCfg llr{};
    int llr_nmax = 0, llr_lt_max = 0, llr_ctas_per_sm = 1;
    size_t llr_smem = 0;
    DevBuf preds2, todo, medmad;
    bool timing = false;
    bool verdict_only = false;
    uint8_t* early_ok = nullptr;          // one-shot (wdx_validate_set_early): copy of `success` behind the first validation
    cudaEvent_t early_ev = nullptr;       //           and the event recorded there
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

namespace {
// pA = (adc + offset) * scale in float32, as the pod5 reader calibrates on the host (the reference receives
// `read_record.signal_pa`, file_proc.py:227-262); samples at or beyond n_valid become the NaN padding of a minibatch row.
__global__ void __launch_bounds__(256) calibrate_kernel(const int16_t* __restrict__ adc, int64_t stride_in, const int32_t* __restrict__ n_valid,
                                                        const float* __restrict__ offset, const float* __restrict__ scale,
                                                        float* __restrict__ out, int64_t stride_out) {
    const int64_t r = blockIdx.x;
    const int64_t nv = min((int64_t)n_valid[r], min(stride_in, stride_out));
    const float o = offset[r], sc = scale[r];
    const int16_t* src = adc + r * stride_in;
    float* dst = out + r * stride_out;
    const float qnan = __int_as_float(0x7fc00000);
    for (int64_t i = threadIdx.x; i < stride_out; i += 256)
        dst[i] = (i < nv) ? __fmul_rn(__fadd_rn((float)src[i], o), sc) : qnan;
}

bool range_empty(const double* r) { return r[0] == -INFINITY && r[1] == INFINITY; }
}  // namespace

extern "C" {

int wdx_calibrate_rows(const int16_t* adc, int64_t n, int64_t stride_in, const int32_t* n_valid, const float* offset,
                       const float* scale, float* out, int64_t stride_out, int device, void* stream) {
    if (n < 0 || stride_in < 1 || stride_out < 1) return fail(WDX_ERR_INVALID, "n=%lld strides %lld, %lld", (long long)n, (long long)stride_in, (long long)stride_out);
    if (n == 0) return WDX_OK;
    if (!adc || !n_valid || !offset || !scale || !out) return fail(WDX_ERR_INVALID, "NULL argument");
    for (const void* p : {(const void*)adc, (const void*)n_valid, (const void*)offset, (const void*)scale, (const void*)out})
        if (mem_kind(p) != 2) return fail(WDX_ERR_INVALID, "wdx_calibrate_rows takes device pointers (it is the first kernel after the upload)");
    CUDA_TRY(cudaSetDevice(device));
    calibrate_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(adc, stride_in, n_valid, offset, scale, out, stride_out);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return WDX_OK;
}

int wdx_validate_create(const wdx_validate_config* cfg, int device, wdx_validate** out) {
    if (!cfg || !out) return fail(WDX_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->mvs_detect_overwrite) return fail(WDX_ERR_UNSUPPORTED, "mvs_detect_overwrite = true is not implemented");
    if (cfg->mean_window < 1 || cfg->max_obs_local_range < 1 || cfg->pA_mean_window < 1 || cfg->pA_var_window < 1 ||
        cfg->median_shift_window < 1 || cfg->med_shift_window < 1 || cfg->min_obs_adapter < 0 || cfg->open_pore_min_obs_diff < 1)
        return fail(WDX_ERR_INVALID, "window sizes must be positive");
    const bool from_scale = range_empty(cfg->pA_mean_range) && !range_empty(cfg->pA_mean_adapter_med_scale_range);
    if (cfg->mvs_detect_check && !from_scale && range_empty(cfg->pA_mean_range))
        return fail(WDX_ERR_INVALID, "pA_mean_range is not specified");   // combined.py:521-522 raises the same
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(WDX_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(WDX_ERR_INVALID, "device %d of %d", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    wdx_validate* h = new (std::nothrow) wdx_validate();
    if (!h) return fail(WDX_ERR_NOMEM, "out of host memory");
    h->device = device;
    ValCfg& c = h->cfg;
    c.min_obs_adapter = cfg->min_obs_adapter;
    c.detect_open_pores = cfg->detect_open_pores;
    c.real_signal_check = cfg->real_signal_check;
    c.mean_window = cfg->mean_window;
    c.max_obs_local_range = cfg->max_obs_local_range;
    c.mean_start_lo = cfg->mean_start_range[0];
    c.mean_start_hi = cfg->mean_start_range[1];
    c.mean_end_lo = cfg->mean_end_range[0];
    c.mean_end_hi = cfg->mean_end_range[1];
    c.local_range_lo = cfg->local_range[0];
    c.local_range_hi = cfg->local_range[1];
    c.mad_lo = cfg->adapter_mad_range[0];
    c.mad_hi = cfg->adapter_mad_range[1];
    c.open_pore_min = (float)cfg->open_pore_min;
    c.open_pore_min_obs_diff = cfg->open_pore_min_obs_diff;
    c.mvs_detect_check = cfg->mvs_detect_check;
    c.pa_mean_window = cfg->pA_mean_window;
    c.pa_var_window = cfg->pA_var_window;
    c.median_shift_window = cfg->median_shift_window;
    c.var_lo = cfg->pA_var_range[0];
    c.var_hi = cfg->pA_var_range[1];
    c.shift_lo = cfg->median_shift_range[0];
    c.shift_hi = cfg->median_shift_range[1];
    c.pmed_lo = cfg->polyA_med_range[0];
    c.pmed_hi = cfg->polyA_med_range[1];
    c.plr_lo = cfg->polyA_local_range[0];
    c.plr_hi = cfg->polyA_local_range[1];
    c.mean_lo = cfg->pA_mean_range[0];
    c.mean_hi = cfg->pA_mean_range[1];
    c.scale_lo = cfg->pA_mean_adapter_med_scale_range[0];
    c.scale_hi = cfg->pA_mean_adapter_med_scale_range[1];
    c.mean_from_scale = from_scale;
    c.detect_med_shift = cfg->detect_med_shift;
    c.med_shift_window = cfg->med_shift_window;
    c.ms_lo = cfg->med_shift_range[0];
    c.ms_hi = cfg->med_shift_range[1];
    if ((double)c.open_pore_min != cfg->open_pore_min) {
        delete h;
        return fail(WDX_ERR_UNSUPPORTED, "open_pore_min must be representable in float32");
    }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, validate_kernel));
    h->smem_max = (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes - 1024;
    CUDA_TRY(cudaFuncSetAttribute(validate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_max));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    *out = h;
    return WDX_OK;
}

void wdx_validate_destroy(wdx_validate* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (DevBuf* b : {&h->sig, &h->len, &h->preds, &h->success, &h->info, &h->bounds, &h->vals, &h->parts, &h->pores, &h->scratch, &h->counter,
                      &h->preds2, &h->todo, &h->medmad})
        b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wdx_validate_set_early(wdx_validate* h, uint8_t* success_snapshot, void* event) {
    if (!h) return fail(WDX_ERR_INVALID, "NULL validation handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->early_ok = success_snapshot;
    h->early_ev = (cudaEvent_t)event;
    return WDX_OK;
}

int wdx_validate_set_verdict_only(wdx_validate* h, int on) {
    if (!h) return fail(WDX_ERR_INVALID, "NULL validation handle");
    h->verdict_only = on != 0;
    return WDX_OK;
}

int wdx_validate_set_llr(wdx_validate* h, const wdx_llr_config* cfg) {
    if (!h) return fail(WDX_ERR_INVALID, "NULL validation handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!cfg) {
        h->llr_on = false;
        return WDX_OK;
    }
    if (cfg->downscale_factor < 1 || cfg->downscale_factor > 128 || cfg->max_obs_trace < 1 || cfg->min_obs_adapter < 0 ||
        cfg->max_obs_adapter < 1 || cfg->adapter_peak_width < 0)
        return fail(WDX_ERR_INVALID, "wdx_llr_config: sizes out of range (1 <= downscale_factor <= 128)");
    LlrCfg& c = h->llr;
    c.max_obs_trace = cfg->max_obs_trace;
    c.min_obs_adapter = cfg->min_obs_adapter;
    c.max_obs_adapter = cfg->max_obs_adapter;
    c.factor = cfg->downscale_factor;
    c.outlier_thresh = cfg->sig_norm_outlier_thresh;
    c.peak_prominence = cfg->adapter_peak_prominence;
    c.peak_rel_height = cfg->adapter_peak_rel_height;
    c.peak_width = cfg->adapter_peak_width / cfg->downscale_factor;
    c.fallback_to_llr = cfg->fallback_to_llr != 0;
    c.fallback_short_reads = cfg->fallback_to_llr_short_reads != 0;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, llr_kernel));
    const int room = (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes - 1024;
    h->llr_lt_max = cfg->max_obs_trace;
    h->llr_nmax = (cfg->max_obs_trace + cfg->downscale_factor - 1) / cfg->downscale_factor + 8;
    h->llr_smem = llr_smem_bytes(h->llr_nmax, h->llr_lt_max);
    if ((int64_t)h->llr_smem > room)
        return fail(WDX_ERR_UNSUPPORTED, "max_obs_trace = %d needs %zu B of shared memory, device allows %d", cfg->max_obs_trace, h->llr_smem, room);
    CUDA_TRY(cudaFuncSetAttribute(llr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->llr_smem));
    int per_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, llr_kernel, FP_THREADS, h->llr_smem));
    h->llr_ctas_per_sm = std::max(1, per_sm);
    h->llr_on = c.fallback_to_llr || c.fallback_short_reads;
    return WDX_OK;
}

int wdx_validate_enable_timing(wdx_validate* h, int on) {
    if (!h) return fail(WDX_ERR_INVALID, "NULL validation handle");
    h->timing = on != 0;
    return WDX_OK;
}

int wdx_validate_last_kernel_ms(wdx_validate* h, double* ms, int* launches) {
    if (!h || !ms) return fail(WDX_ERR_INVALID, "NULL argument");
    *ms = 0.0;
    if (launches) *launches = 0;
    if (!h->timed) return WDX_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    float t = 0.f;
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev0, h->ev1));
    *ms = t;
    if (launches) *launches = 1;
    return WDX_OK;
}

int wdx_validate_run_report(wdx_validate* h, const float* signals, int64_t n, int64_t stride, const int32_t* full_len,
                            const int64_t* preds, int32_t ld, uint8_t* success, int32_t* info, int64_t* bounds, double* vals,
                            double* parts, int32_t* open_pores, void* stream);

int wdx_validate_run_ex(wdx_validate* h, const float* signals, int64_t n, int64_t stride, const int32_t* full_len,
                        const int64_t* preds, int32_t ld, uint8_t* success, int32_t* info, int64_t* bounds, double* vals,
                        double* parts, void* stream) {
    return wdx_validate_run_report(h, signals, n, stride, full_len, preds, ld, success, info, bounds, vals, parts, nullptr, stream);
}

int wdx_validate_run(wdx_validate* h, const float* signals, int64_t n, int64_t stride, const int32_t* full_len,
                     const int64_t* preds, int32_t ld, uint8_t* success, int32_t* info, int64_t* bounds, double* vals,
                     void* stream) {
    return wdx_validate_run_ex(h, signals, n, stride, full_len, preds, ld, success, info, bounds, vals, nullptr, stream);
}

int wdx_validate_run_report(wdx_validate* h, const float* signals, int64_t n, int64_t stride, const int32_t* full_len,
                            const int64_t* preds, int32_t ld, uint8_t* success, int32_t* info, int64_t* bounds, double* vals,
                            double* parts, int32_t* open_pores, void* stream) {
    if (!h) return fail(WDX_ERR_INVALID, "NULL validation handle");
    if (n < 0 || stride < 1 || ld < 1) return fail(WDX_ERR_INVALID, "n=%lld stride=%lld ld=%d", (long long)n, (long long)stride, ld);
    if (n == 0) return WDX_OK;
    if (!signals || !full_len || !preds || !success || !info || !bounds)
        return fail(WDX_ERR_INVALID, "signals, full_len, preds, success, info and bounds are required");
    // the row (float32) + one 16-bit code per sample (moving variance of the poly(A) windows, validate_kernel.cuh)
    const size_t smem = (size_t)((stride + 3) & ~(int64_t)3) * 4 + (((size_t)stride * 2 + 15) & ~(size_t)15);
    if ((int64_t)smem > h->smem_max)
        return fail(WDX_ERR_UNSUPPORTED, "rows of %lld samples need %zu B of shared memory, device allows %d", (long long)stride, smem, h->smem_max);
    std::lock_guard<std::mutex> lk(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    int rc;
    h->timed = false;

    const float* sig_d = signals;
    int sig_kind = mem_kind(signals);
    if (sig_kind == 1) {   // pinned host rows: read once by the kernel, straight over PCIe
        void* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, const_cast<float*>(signals), 0) == cudaSuccess && dptr) {
            sig_d = (const float*)dptr;
            sig_kind = 2;
        } else {
            cudaGetLastError();
        }
    }
    if (sig_kind != 2) {
        if ((rc = h->sig.reserve((size_t)n * stride * 4))) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->sig.p, signals, (size_t)n * stride * 4, cudaMemcpyHostToDevice, st));
        sig_d = (const float*)h->sig.p;
    }
    const int32_t* len_d = full_len;
    if (mem_kind(full_len) != 2) {
        if ((rc = h->len.reserve((size_t)n * 4))) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->len.p, full_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        len_d = (const int32_t*)h->len.p;
    }
    const int64_t* preds_d = preds;
    if (mem_kind(preds) != 2) {
        if ((rc = h->preds.reserve((size_t)n * ld * 8))) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->preds.p, preds, (size_t)n * ld * 8, cudaMemcpyHostToDevice, st));
        preds_d = (const int64_t*)h->preds.p;
    }
    const bool suc_dev = mem_kind(success) == 2, info_dev = mem_kind(info) == 2, bnd_dev = mem_kind(bounds) == 2,
               val_dev = vals && mem_kind(vals) == 2, part_dev = parts && mem_kind(parts) == 2,
               pore_dev = open_pores && mem_kind(open_pores) == 2;
    if (!suc_dev && (rc = h->success.reserve((size_t)n))) return rc;
    if (!info_dev && (rc = h->info.reserve((size_t)n * 16))) return rc;
    if (!bnd_dev && (rc = h->bounds.reserve((size_t)n * 24))) return rc;
    if (vals && !val_dev && (rc = h->vals.reserve((size_t)n * VAL_NVALS * 8))) return rc;
    if (parts && !part_dev && (rc = h->parts.reserve((size_t)n * VAL_NPART * 8))) return rc;
    if (open_pores && !pore_dev && (rc = h->pores.reserve((size_t)n * VAL_PORES_LD * 4))) return rc;

    // persistent grid: as many CTAs as fit (shared memory bound), each owning one scratch row
    int per_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, validate_kernel, FP_THREADS, smem));
    per_sm = std::max(1, per_sm);
    const int grid = (int)std::min<int64_t>(n, (int64_t)per_sm * h->sm_count);
    if ((rc = h->scratch.reserve((size_t)grid * stride * 4))) return rc;

    ValArgs a{};
    a.signals = sig_d;
    a.stride = stride;
    a.full_len = len_d;
    a.preds = preds_d;
    a.ld = ld;
    a.n = n;
    a.success = suc_dev ? success : (uint8_t*)h->success.p;
    a.info = info_dev ? info : (int32_t*)h->info.p;
    a.bounds = bnd_dev ? bounds : (int64_t*)h->bounds.p;
    a.vals = vals ? (val_dev ? vals : (double*)h->vals.p) : nullptr;
    a.parts = parts ? (part_dev ? parts : (double*)h->parts.p) : nullptr;
    a.pores = open_pores ? (pore_dev ? open_pores : (int32_t*)h->pores.p) : nullptr;
    a.scratch = (float*)h->scratch.p;
    a.verdict_only = h->verdict_only ? 1 : 0;
    if ((rc = h->counter.reserve(64))) return rc;
    CUDA_TRY(cudaMemsetAsync(h->counter.p, 0, 64, st));
    unsigned long long* counters = (unsigned long long*)h->counter.p;
    a.next = counters;
    int* L[4] = {nullptr, nullptr, nullptr, nullptr};
    int* lc = nullptr;
    if (h->llr_on) {
        if ((rc = h->preds2.reserve((size_t)n * ld * 8))) return rc;
        if ((rc = h->todo.reserve((size_t)n * 4 * 4 + 64))) return rc;    // four work lists + their counters
        if ((rc = h->medmad.reserve((size_t)n * 8))) return rc;
        lc = (int*)h->todo.p;
        for (int q = 0; q < 4; q++) L[q] = lc + 16 + (size_t)q * n;
        CUDA_TRY(cudaMemsetAsync(lc, 0, 64, st));
        a.fail_list = L[0];
        a.fail_count = lc + 0;
    }
    if (h->timing) CUDA_TRY(cudaEventRecord(h->ev0, st));
    validate_kernel<<<grid, FP_THREADS, smem, st>>>(a, h->cfg);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    {   // wdx_validate_set_early: the verdicts of the reads that pass are final here — what follows only revisits the failed ones
        uint8_t* early_ok = h->early_ok;
        cudaEvent_t early_ev = h->early_ev;
        h->early_ok = nullptr;
        h->early_ev = nullptr;
        if (early_ok) CUDA_TRY(cudaMemcpyAsync(early_ok, a.success, (size_t)n, cudaMemcpyDeviceToDevice, st));
        if (early_ev) CUDA_TRY(cudaEventRecord(early_ev, st));
    }
    if (h->llr_on) {
        // combined.py:222-290: the reads that failed get a poly(A) re-detection on the CNN's adapter end ("hail mary",
        // result assigned unconditionally), then a full LLR detection (result assigned only when it validates).
        // Device-side work lists chain the launches: L0 = failed reads (written by the validation above) -> llr stage 0 ->
        // L1 (hail-mary proposal: validated next) / L2 (no proposal) -> validation of L1 appends its failures to L2 ->
        // llr stage 1 on L2 -> L3 (LLR proposal) -> validation of L3, committed on success.
        LlrArgs la{};
        la.signals = sig_d;
        la.stride = stride;
        la.full_len = len_d;
        la.cnn_preds = preds_d;
        la.ld = ld;
        la.n = n;
        la.info = a.info;
        la.preds_out = (int64_t*)h->preds2.p;
        la.medmad = (float*)h->medmad.p;
        la.nmax = h->llr_nmax;
        la.lt_max = h->llr_lt_max;
        const int llr_grid = (int)std::min<int64_t>(n, (int64_t)h->llr_ctas_per_sm * h->sm_count);
        ValArgs b = a;
        b.preds = (const int64_t*)h->preds2.p;
        b.verdict_only = 0;   // single-candidate lists: nothing to cut short
        const bool hm_on = h->llr.fallback_short_reads, full_on = h->llr.fallback_to_llr;
        if (hm_on) {
            la.stage = 0;
            la.in_list = L[0]; la.in_count = lc + 0;
            la.todo_list = L[1]; la.todo_count = lc + 1;
            la.pass_list = L[2]; la.pass_count = lc + 2;
            la.next = counters + 1;
            llr_kernel<<<llr_grid, FP_THREADS, h->llr_smem, st>>>(la, h->llr);
            CUDA_TRY(cudaGetLastError());
            b.list = L[1]; b.list_count = lc + 1;
            b.fail_list = L[2]; b.fail_count = lc + 2;
            b.next = counters + 2;
            b.commit_on_success = 0;
            b.src_tag = LLR_SRC_HAIL_MARY;
            validate_kernel<<<std::min(grid, llr_grid), FP_THREADS, smem, st>>>(b, h->cfg);
            CUDA_TRY(cudaGetLastError());
            g_launches += 2;
        }
        if (full_on) {
            la.stage = 1;
            la.in_list = hm_on ? L[2] : L[0]; la.in_count = hm_on ? lc + 2 : lc + 0;
            la.todo_list = L[3]; la.todo_count = lc + 3;
            la.pass_list = nullptr; la.pass_count = nullptr;
            if (!hm_on) la.medmad = nullptr;   // stage 0 did not run: the median / MAD are computed here
            la.next = counters + 3;
            llr_kernel<<<llr_grid, FP_THREADS, h->llr_smem, st>>>(la, h->llr);
            CUDA_TRY(cudaGetLastError());
            b.list = L[3]; b.list_count = lc + 3;
            b.fail_list = nullptr; b.fail_count = nullptr;
            b.next = counters + 4;
            b.commit_on_success = 1;
            b.src_tag = LLR_SRC_LLR;
            validate_kernel<<<std::min(grid, llr_grid), FP_THREADS, smem, st>>>(b, h->cfg);
            CUDA_TRY(cudaGetLastError());
            g_launches += 2;
        }
    }
    if (h->timing) {
        CUDA_TRY(cudaEventRecord(h->ev1, st));
        h->timed = true;
    }
    bool any_host = false;
    if (!suc_dev) { CUDA_TRY(cudaMemcpyAsync(success, a.success, (size_t)n, cudaMemcpyDeviceToHost, st)); any_host = true; }
    if (!info_dev) { CUDA_TRY(cudaMemcpyAsync(info, a.info, (size_t)n * 16, cudaMemcpyDeviceToHost, st)); any_host = true; }
    if (!bnd_dev) { CUDA_TRY(cudaMemcpyAsync(bounds, a.bounds, (size_t)n * 24, cudaMemcpyDeviceToHost, st)); any_host = true; }
    if (vals && !val_dev) { CUDA_TRY(cudaMemcpyAsync(vals, a.vals, (size_t)n * VAL_NVALS * 8, cudaMemcpyDeviceToHost, st)); any_host = true; }
    if (parts && !part_dev) { CUDA_TRY(cudaMemcpyAsync(parts, a.parts, (size_t)n * VAL_NPART * 8, cudaMemcpyDeviceToHost, st)); any_host = true; }
    if (open_pores && !pore_dev) { CUDA_TRY(cudaMemcpyAsync(open_pores, a.pores, (size_t)n * VAL_PORES_LD * 4, cudaMemcpyDeviceToHost, st)); any_host = true; }
    // host buffers (inputs staged from pageable memory, results) are only safe to touch after the stream drained
    if (any_host || sig_kind != 2 || len_d != full_len || preds_d != preds) CUDA_TRY(cudaStreamSynchronize(st));
    return WDX_OK;
}

#ifdef WDX_FP_PROF
// experiments only: cycles per phase of validate_kernel summed over all reads since the last reset
int wdx_validate_prof_dump(unsigned long long* out32, int reset) {
    if (out32) CUDA_TRY(cudaMemcpyFromSymbol(out32, wdx::g_fp_prof, 32 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long z[32] = {};
        CUDA_TRY(cudaMemcpyToSymbol(wdx::g_fp_prof, z, sizeof z));
    }
    return WDX_OK;
}
#endif
}  // extern "C"
