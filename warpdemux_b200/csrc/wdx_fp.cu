// wdx_fp.cu — C-ABI entry points of the fingerprint stage (include/wdx_b200.h:
// wdx_fp_*): handle, staging of host buffers, launch of fingerprint_kernel, and
// the fused minibatch step fingerprint -> DTW+SVC with the fingerprints kept on
// the device.  No CPU compute path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "fingerprint_kernel.cuh"
#include "wdx_internal.cuh"

using namespace wdx;

struct wdx_fp {
    FpConfig cfg{};
    int max_slice_len = 0;
    int long_slice_len = 0;   // > max_slice_len: reads too long for the first pass are redone with this capacity
    int resume_status = 0;    // one-shot (wdx_fp_set_resume_status): the next call's fingerprint pass only revisits reads with this status
    int device = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;       // kernels + result copies
    cudaStream_t copy_stream = nullptr;  // H2D of the next chunk
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    DevBuf sig[2], len[2], a0[2], a1[2], ok[2], maxlen;
    DevBuf fpt[2], dwell[2], stats[2], status[2], lab[2], conf[2], prob[2], flags[2], cons[2];
    DevBuf cons_query;                   // consensus query (consensus-guided mode)
    DevBuf park;                         // three-launch consensus form: the reads' state between the kernels
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev;
    size_t tev_used = 0;
    int smem_max = 0;  // opt-in dynamic shared memory granted to fingerprint_kernel on this device
};

namespace {

struct FpCall {
    const float* signals;
    int64_t n, stride;
    const int32_t* sig_len;
    const int64_t *a0, *a1;
    const uint8_t* ok;
    int clip_in_place;
    double* fpt;
    int64_t* dwell;
    double* stats;
    int32_t* status;
    int32_t* cons;
    // fused predict (m == nullptr: extraction only)
    wdx_model* m;
    int mode;
    int64_t* labels;
    double *conf, *prob;
    uint8_t* flags;
    cudaStream_t user_stream;
};

bool zero_copy_disabled() {
    static const bool off = [] { const char* e = getenv("WDX_FP_NO_ZERO_COPY"); return e && e[0] == '1'; }();
    return off;
}

constexpr int64_t FP_SPLIT_MIN_READS = 2048;   // below: one launch (lower latency); above: the three-launch consensus form
bool split_disabled() {
    const char* e = getenv("WDX_FP_NO_SPLIT");
    return e && *e && *e != '0';
}

int launch_fp(wdx_fp* f, const FpArgs& fa, cudaStream_t st) {
    const size_t smem = fingerprint_smem_bytes(fa.cap);
    if ((int)smem > f->smem_max) return fail(WDX_ERR_UNSUPPORTED, "adapter slice of %d samples needs %zu B of shared memory, device allows %d", fa.cap, smem, f->smem_max);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (f->timing) {
        if (f->tev_used == f->tev.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            f->tev.emplace_back(a, b);
        }
        e0 = f->tev[f->tev_used].first;
        e1 = f->tev[f->tev_used].second;
        f->tev_used++;
        CUDA_TRY(cudaEventRecord(e0, st));
    }
    if (f->cfg.cons_len > 0 && fa.park && fa.retry_status == 0) {
        // big consensus batches: front part, alignment (small CTAs, many per SM), refinement — see fingerprint_kernel
        fingerprint_kernel<true, 1><<<(unsigned)fa.n, FP_THREADS, smem, st>>>(f->cfg, fa);
        CUDA_TRY(cudaGetLastError());
        switch ((f->cfg.cons_len + 31) / 32) {
            case 1: fingerprint_cons_align_kernel<1><<<(unsigned)fa.n, 32, 0, st>>>(f->cfg, fa); break;
            case 2: fingerprint_cons_align_kernel<2><<<(unsigned)fa.n, 64, 0, st>>>(f->cfg, fa); break;
            case 3: fingerprint_cons_align_kernel<3><<<(unsigned)fa.n, 96, 0, st>>>(f->cfg, fa); break;
            default: fingerprint_cons_align_kernel<4><<<(unsigned)fa.n, 128, 0, st>>>(f->cfg, fa); break;
        }
        CUDA_TRY(cudaGetLastError());
        fingerprint_cons_tail_kernel<<<(unsigned)fa.n, FP_THREADS, smem, st>>>(f->cfg, fa);
        g_launches += 2;
    } else if (f->cfg.cons_len > 0) {
        FpArgs fw = fa;
        fw.park = nullptr;
        fingerprint_kernel<true><<<(unsigned)fa.n, FP_THREADS, smem, st>>>(f->cfg, fw);
    } else {
        fingerprint_kernel<false><<<(unsigned)fa.n, FP_THREADS, smem, st>>>(f->cfg, fa);
    }
    CUDA_TRY(cudaGetLastError());
    if (f->timing) CUDA_TRY(cudaEventRecord(e1, st));
    g_launches++;
    return WDX_OK;
}

int run(wdx_fp* f, const FpCall& c) {
    if (!f) return fail(WDX_ERR_INVALID, "NULL fingerprint handle");
    if (c.n < 0 || c.stride < 1) return fail(WDX_ERR_INVALID, "n=%lld stride=%lld", (long long)c.n, (long long)c.stride);
    if (c.n == 0) return WDX_OK;
    if (!c.signals || !c.a0 || !c.a1 || !c.status) return fail(WDX_ERR_INVALID, "signals, adapter_start, adapter_end and status are required");
    if (!c.m && !c.fpt) return fail(WDX_ERR_INVALID, "fpt is required");
    if (c.m && !c.labels) return fail(WDX_ERR_INVALID, "labels is required");
    const int nb = f->cfg.barcode_num_events;
    if (c.m) {
        if (c.m->device != f->device) return fail(WDX_ERR_INVALID, "model lives on device %d, fingerprint handle on %d", c.m->device, f->device);
        if (c.m->L != nb) return fail(WDX_ERR_INVALID, "model fingerprint length %d != barcode_num_events %d", c.m->L, nb);
        if (c.mode < WDX_MODE_EXACT_F64 || c.mode > WDX_MODE_FAST_F32_GUARDED) return fail(WDX_ERR_INVALID, "mode=%d", c.mode);
    }
    std::lock_guard<std::mutex> lk(f->mu);
    std::unique_lock<std::mutex> lkm;
    if (c.m) lkm = std::unique_lock<std::mutex>(c.m->mu);
    CUDA_TRY(cudaSetDevice(f->device));
    f->tev_used = 0;
    if (c.m) c.m->tev_used = 0;
    cudaStream_t st = c.user_stream ? c.user_stream : f->stream;

    const int k = c.m ? c.m->k : 0;
    // Signals in PINNED host memory are read by the kernel straight over PCIe (zero copy): each CTA
    // touches only its adapter slice, so about half the bytes of the NaN-padded rows never move.
    const float* signals = c.signals;
    int sig_kind = mem_kind(c.signals);
    if (sig_kind == 1 && !zero_copy_disabled()) {
        void* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, const_cast<float*>(c.signals), 0) == cudaSuccess && dptr) {
            signals = (const float*)dptr;
            sig_kind = 2;
        } else {
            cudaGetLastError();
        }
    }
    const bool sig_dev = sig_kind == 2;
    const bool len_dev = c.sig_len && mem_kind(c.sig_len) == 2;
    const bool a0_dev = mem_kind(c.a0) == 2, a1_dev = mem_kind(c.a1) == 2;
    const bool ok_dev = c.ok && mem_kind(c.ok) == 2;
    const bool fpt_dev = c.fpt && mem_kind(c.fpt) == 2;
    const bool dwell_dev = c.dwell && mem_kind(c.dwell) == 2;
    const bool stats_dev = c.stats && mem_kind(c.stats) == 2;
    const bool status_dev = mem_kind(c.status) == 2;
    const bool lab_dev = c.labels && mem_kind(c.labels) == 2;
    const bool conf_dev = c.conf && mem_kind(c.conf) == 2;
    const bool prob_dev = c.prob && mem_kind(c.prob) == 2;
    const bool flags_dev = c.flags && mem_kind(c.flags) == 2;
    const bool cons_dev = c.cons && mem_kind(c.cons) == 2;
    const bool any_host = !sig_dev || (c.sig_len && !len_dev) || !a0_dev || !a1_dev || (c.ok && !ok_dev) ||
                          (c.fpt && !fpt_dev) || (c.dwell && !dwell_dev) || (c.stats && !stats_dev) || !status_dev ||
                          (c.labels && !lab_dev) || (c.conf && !conf_dev) || (c.prob && !prob_dev) ||
                          (c.flags && !flags_dev) || (c.cons && !cons_dev);

    const int resume_status = f->resume_status;
    f->resume_status = 0;
    if (resume_status && !status_dev) return fail(WDX_ERR_INVALID, "wdx_fp_set_resume_status needs the status array of the earlier pass on the device");
    // shared-memory capacity per read
    int64_t cap64 = f->max_slice_len;
    if (cap64 <= 0) {
        if (!a0_dev && !a1_dev) {
            cap64 = 0;
            for (int64_t r = 0; r < c.n; r++) {
                int64_t b = c.a0[r] - f->cfg.padding, e = c.a1[r] + f->cfg.padding;
                if (b < 0) b = 0;
                if (e > c.stride) e = c.stride;
                cap64 = std::max(cap64, e - b);
            }
        } else if (a0_dev && a1_dev && (!c.sig_len || len_dev)) {
            // bounds live on the device: one small reduction + a 4-byte read back (a host sync; give
            // wdx_fp_config.max_slice_len to stay asynchronous)
            int rc0 = f->maxlen.reserve(sizeof(int));
            if (rc0) return rc0;
            CUDA_TRY(cudaMemsetAsync(f->maxlen.p, 0, sizeof(int), st));
            const unsigned blocks = (unsigned)std::min<int64_t>(1184, (c.n + 255) / 256);
            max_slice_kernel<<<blocks, 256, 0, st>>>(c.a0, c.a1, c.sig_len, c.n, c.stride, f->cfg.padding, (int*)f->maxlen.p);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
            int h = 0;
            CUDA_TRY(cudaMemcpyAsync(&h, f->maxlen.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            cap64 = h;
        } else {
            cap64 = c.stride;
        }
    }
    cap64 = std::min<int64_t>(FP_MAX_LEN, std::max<int64_t>(64, (cap64 + 63) & ~(int64_t)63));
    const int cap = (int)cap64;

    // host rows are staged in 256 MB pieces (the next piece copying while this one computes); rows already on the device
    // need no staging, so the batch is cut only where the result buffers of a host caller would grow without bound
    int64_t chunk = std::max<int64_t>(1, ((int64_t)256 << 20) / (c.stride * 4));
    if (sig_dev) chunk = (int64_t)1 << 20;
    if (c.m) chunk = std::min<int64_t>(chunk, c.m->chunk_reads);
    if (f->cfg.cons_len > 0) chunk = std::min<int64_t>(chunk, (int64_t)1 << 15);   // bounds the parked state (~70 KB per read)
    chunk = std::min(chunk, c.n);
    const int64_t n_chunks = (c.n + chunk - 1) / chunk;
    const int nbuf = n_chunks > 1 ? 2 : 1;
    int rc;
    for (int b = 0; b < nbuf; b++) {
        if (!sig_dev && (rc = f->sig[b].reserve((size_t)chunk * c.stride * 4))) return rc;
        if (c.sig_len && !len_dev && (rc = f->len[b].reserve((size_t)chunk * 4))) return rc;
        if (!a0_dev && (rc = f->a0[b].reserve((size_t)chunk * 8))) return rc;
        if (!a1_dev && (rc = f->a1[b].reserve((size_t)chunk * 8))) return rc;
        if (c.ok && !ok_dev && (rc = f->ok[b].reserve((size_t)chunk))) return rc;
        if (!fpt_dev && (rc = f->fpt[b].reserve((size_t)chunk * nb * 8))) return rc;  // also the hand-over buffer of the fused step
        if (c.dwell && !dwell_dev && (rc = f->dwell[b].reserve((size_t)chunk * nb * 8))) return rc;
        if (c.stats && !stats_dev && (rc = f->stats[b].reserve((size_t)chunk * 6 * 8))) return rc;
        if (!status_dev && (rc = f->status[b].reserve((size_t)chunk * 4))) return rc;
        if (c.labels && !lab_dev && (rc = f->lab[b].reserve((size_t)chunk * 8))) return rc;
        if (c.conf && !conf_dev && (rc = f->conf[b].reserve((size_t)chunk * 8))) return rc;
        if (c.prob && !prob_dev && (rc = f->prob[b].reserve((size_t)chunk * k * 8))) return rc;
        if (c.flags && !flags_dev && (rc = f->flags[b].reserve((size_t)chunk))) return rc;
        if (c.cons && !cons_dev && (rc = f->cons[b].reserve((size_t)chunk * 3 * 4))) return rc;
    }

    const bool stage_any = !sig_dev || (c.sig_len && !len_dev) || !a0_dev || !a1_dev || (c.ok && !ok_dev);
    auto stage_in = [&](int64_t ci) -> int {
        const int b = (int)(ci & 1);
        const int64_t r0 = ci * chunk, cn = std::min(chunk, c.n - r0);
        CUDA_TRY(cudaEventSynchronize(f->ev_free[b]));  // the chunk that used these buffers is finished
        if (!sig_dev)
            CUDA_TRY(cudaMemcpyAsync(f->sig[b].p, signals + (size_t)r0 * c.stride, (size_t)cn * c.stride * 4,
                                     cudaMemcpyHostToDevice, f->copy_stream));
        if (c.sig_len && !len_dev)
            CUDA_TRY(cudaMemcpyAsync(f->len[b].p, c.sig_len + r0, (size_t)cn * 4, cudaMemcpyHostToDevice, f->copy_stream));
        if (!a0_dev) CUDA_TRY(cudaMemcpyAsync(f->a0[b].p, c.a0 + r0, (size_t)cn * 8, cudaMemcpyHostToDevice, f->copy_stream));
        if (!a1_dev) CUDA_TRY(cudaMemcpyAsync(f->a1[b].p, c.a1 + r0, (size_t)cn * 8, cudaMemcpyHostToDevice, f->copy_stream));
        if (c.ok && !ok_dev) CUDA_TRY(cudaMemcpyAsync(f->ok[b].p, c.ok + r0, (size_t)cn, cudaMemcpyHostToDevice, f->copy_stream));
        CUDA_TRY(cudaEventRecord(f->ev_h2d[b], f->copy_stream));
        return WDX_OK;
    };
    if (stage_any) {
        CUDA_TRY(cudaEventRecord(f->ev_free[0], st));
        CUDA_TRY(cudaEventRecord(f->ev_free[1], st));
        if ((rc = stage_in(0))) return rc;
    }

    for (int64_t ci = 0; ci < n_chunks; ci++) {
        const int b = (int)(ci & 1);
        const int64_t r0 = ci * chunk, cn = std::min(chunk, c.n - r0);
        if (stage_any) CUDA_TRY(cudaStreamWaitEvent(st, f->ev_h2d[b], 0));
        FpArgs fa{};
        float* sig_d = sig_dev ? const_cast<float*>(signals) + (size_t)r0 * c.stride : (float*)f->sig[b].p;
        fa.signals = sig_d;
        fa.signals_mut = c.clip_in_place ? sig_d : nullptr;
        fa.stride = c.stride;
        fa.sig_len = c.sig_len ? (len_dev ? c.sig_len + r0 : (const int32_t*)f->len[b].p) : nullptr;
        fa.adapter_start = a0_dev ? c.a0 + r0 : (const int64_t*)f->a0[b].p;
        fa.adapter_end = a1_dev ? c.a1 + r0 : (const int64_t*)f->a1[b].p;
        fa.detect_ok = c.ok ? (ok_dev ? c.ok + r0 : (const uint8_t*)f->ok[b].p) : nullptr;
        fa.n = cn;
        fa.cap = cap;
        double* fpt_d = fpt_dev ? c.fpt + (size_t)r0 * nb : (double*)f->fpt[b].p;
        fa.fpt = fpt_d;
        fa.dwell = c.dwell ? (dwell_dev ? c.dwell + (size_t)r0 * nb : (int64_t*)f->dwell[b].p) : nullptr;
        fa.stats = c.stats ? (stats_dev ? c.stats + (size_t)r0 * 6 : (double*)f->stats[b].p) : nullptr;
        fa.status = status_dev ? c.status + r0 : (int32_t*)f->status[b].p;
        fa.cons_query = (const double*)f->cons_query.p;
        fa.cons = c.cons ? (cons_dev ? c.cons + (size_t)r0 * 3 : (int32_t*)f->cons[b].p) : nullptr;
        fa.park = nullptr;
        if (f->cfg.cons_len > 0 && cn >= FP_SPLIT_MIN_READS && !split_disabled()) {
            if ((rc = f->park.reserve((size_t)cn * fp_park_bytes(cap)))) return rc;
            fa.park = (unsigned char*)f->park.p;
        }
        fa.retry_status = resume_status;
        if ((rc = launch_fp(f, fa, st))) return rc;
        fa.retry_status = 0;
        if (f->max_slice_len > 0 && f->long_slice_len > cap) {   // the few reads longer than the first pass's capacity
            FpArgs fl = fa;
            fl.cap = std::min<int>(FP_MAX_LEN, (f->long_slice_len + 63) & ~63);
            fl.retry_status = FP_FAIL_TOO_LONG;
            if ((rc = launch_fp(f, fl, st))) return rc;
        }

        int64_t* lab_d = nullptr;
        double *conf_d = nullptr, *prob_d = nullptr;
        uint8_t* flags_d = nullptr;
        if (c.m) {
            lab_d = lab_dev ? c.labels + r0 : (int64_t*)f->lab[b].p;
            conf_d = c.conf ? (conf_dev ? c.conf + r0 : (double*)f->conf[b].p) : nullptr;
            prob_d = c.prob ? (prob_dev ? c.prob + (size_t)r0 * k : (double*)f->prob[b].p) : nullptr;
            flags_d = c.flags ? (flags_dev ? c.flags + r0 : (uint8_t*)f->flags[b].p) : nullptr;
            if ((rc = predict_chunk_device(c.m, fpt_d, 0, cn, c.mode, lab_d, conf_d, prob_d, flags_d, nullptr, st))) return rc;
        }
        // results (small) go back on the compute stream
        if (c.clip_in_place && !sig_dev)
            CUDA_TRY(cudaMemcpyAsync(const_cast<float*>(c.signals) + (size_t)r0 * c.stride, sig_d, (size_t)cn * c.stride * 4,
                                     cudaMemcpyDeviceToHost, st));
        if (c.fpt && !fpt_dev) CUDA_TRY(cudaMemcpyAsync(c.fpt + (size_t)r0 * nb, fpt_d, (size_t)cn * nb * 8, cudaMemcpyDeviceToHost, st));
        if (c.dwell && !dwell_dev)
            CUDA_TRY(cudaMemcpyAsync(c.dwell + (size_t)r0 * nb, fa.dwell, (size_t)cn * nb * 8, cudaMemcpyDeviceToHost, st));
        if (c.stats && !stats_dev)
            CUDA_TRY(cudaMemcpyAsync(c.stats + (size_t)r0 * 6, fa.stats, (size_t)cn * 6 * 8, cudaMemcpyDeviceToHost, st));
        if (!status_dev) CUDA_TRY(cudaMemcpyAsync(c.status + r0, fa.status, (size_t)cn * 4, cudaMemcpyDeviceToHost, st));
        if (c.cons && !cons_dev)
            CUDA_TRY(cudaMemcpyAsync(c.cons + (size_t)r0 * 3, fa.cons, (size_t)cn * 3 * 4, cudaMemcpyDeviceToHost, st));
        if (c.m) {
            if (!lab_dev) CUDA_TRY(cudaMemcpyAsync(c.labels + r0, lab_d, (size_t)cn * 8, cudaMemcpyDeviceToHost, st));
            if (c.conf && !conf_dev) CUDA_TRY(cudaMemcpyAsync(c.conf + r0, conf_d, (size_t)cn * 8, cudaMemcpyDeviceToHost, st));
            if (c.prob && !prob_dev)
                CUDA_TRY(cudaMemcpyAsync(c.prob + (size_t)r0 * k, prob_d, (size_t)cn * k * 8, cudaMemcpyDeviceToHost, st));
            if (c.flags && !flags_dev) CUDA_TRY(cudaMemcpyAsync(c.flags + r0, flags_d, (size_t)cn, cudaMemcpyDeviceToHost, st));
        }
        if (stage_any) {
            CUDA_TRY(cudaEventRecord(f->ev_free[b], st));
            if (ci + 1 < n_chunks && (rc = stage_in(ci + 1))) return rc;  // overlaps the kernels just enqueued
        } else if (n_chunks > 1 && any_host) {
            CUDA_TRY(cudaStreamSynchronize(st));  // result buffers are reused two chunks later
        }
    }
    if (any_host || !c.user_stream) CUDA_TRY(cudaStreamSynchronize(st));
    return WDX_OK;
}

}  // namespace

extern "C" {

int wdx_fp_create(const wdx_fp_config* cfg, int device, wdx_fp** out) {
    if (!out) return fail(WDX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!cfg) return fail(WDX_ERR_INVALID, "cfg is NULL");
    if (cfg->num_events < 1 || cfg->num_events > FP_MAX_EVENTS) return fail(WDX_ERR_INVALID, "num_events=%d outside [1,%d]", cfg->num_events, FP_MAX_EVENTS);
    if (cfg->barcode_num_events < 1 || cfg->barcode_num_events > FP_THREADS)
        return fail(WDX_ERR_INVALID, "barcode_num_events=%d", cfg->barcode_num_events);
    if (cfg->padding < 0 || cfg->min_obs_per_base < 1 || cfg->running_stat_width < 1 || !(cfg->outlier_thresh >= 0))
        return fail(WDX_ERR_INVALID, "bad fingerprint configuration");
    if (cfg->max_slice_len < 0 || cfg->max_slice_len > FP_MAX_LEN)
        return fail(WDX_ERR_INVALID, "max_slice_len=%d outside [0,%d]", cfg->max_slice_len, FP_MAX_LEN);
    int ndev = wdx_device_count();
    if (ndev <= 0) return fail(WDX_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(WDX_ERR_INVALID, "device %d of %d", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    wdx_fp* f = new (std::nothrow) wdx_fp();
    if (!f) return fail(WDX_ERR_NOMEM, "host allocation failed");
    f->device = device;
    f->cfg.padding = cfg->padding;
    f->cfg.outlier_thresh = (float)cfg->outlier_thresh;
    f->cfg.outlier_thresh_d = cfg->outlier_thresh;
    f->cfg.numpy1_promotion = 0;
    f->cfg.min_obs_per_base = cfg->min_obs_per_base;
    f->cfg.running_stat_width = cfg->running_stat_width;
    f->cfg.num_events = cfg->num_events;
    f->cfg.barcode_num_events = cfg->barcode_num_events;
    f->max_slice_len = cfg->max_slice_len;
    {   // The attribute belongs to the function (per device), not to a handle: always ask for the
        // device maximum so that handles cannot shrink each other's limit.
        cudaFuncAttributes fa;
        int optin = 0;
        cudaFuncAttributes fb;
        if (cudaFuncGetAttributes(&fa, (const void*)fingerprint_kernel<false>) != cudaSuccess ||
            cudaFuncGetAttributes(&fb, (const void*)fingerprint_kernel<true>) != cudaSuccess ||
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) {
            delete f;
            return fail(WDX_ERR_CUDA, "cannot query shared-memory limits: %s", cudaGetErrorString(cudaGetLastError()));
        }
        f->smem_max = optin - (int)std::max(fa.sharedSizeBytes, fb.sharedSizeBytes);
        if (cudaFuncSetAttribute((const void*)fingerprint_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 f->smem_max) != cudaSuccess ||
            cudaFuncSetAttribute((const void*)fingerprint_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 f->smem_max) != cudaSuccess ||
            cudaFuncSetAttribute((const void*)fingerprint_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 f->smem_max) != cudaSuccess ||
            cudaFuncSetAttribute((const void*)fingerprint_cons_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 f->smem_max) != cudaSuccess) {
            delete f;
            return fail(WDX_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    bool ok = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
        ok = cudaEventCreateWithFlags(&f->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&f->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        wdx_fp_destroy(f);
        return fail(WDX_ERR_CUDA, "stream/event creation failed");
    }
    *out = f;
    return WDX_OK;
}

void wdx_fp_destroy(wdx_fp* f) {
    if (!f) return;
    cudaSetDevice(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    if (f->copy_stream) cudaStreamSynchronize(f->copy_stream);
    for (int i = 0; i < 2; i++) {
        for (DevBuf* b : {&f->sig[i], &f->len[i], &f->a0[i], &f->a1[i], &f->ok[i], &f->fpt[i], &f->dwell[i], &f->stats[i],
                          &f->status[i], &f->lab[i], &f->conf[i], &f->prob[i], &f->flags[i], &f->cons[i]})
            b->release();
        if (f->ev_h2d[i]) cudaEventDestroy(f->ev_h2d[i]);
        if (f->ev_free[i]) cudaEventDestroy(f->ev_free[i]);
    }
    f->maxlen.release();
    f->cons_query.release();
    for (auto& e : f->tev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (f->stream) cudaStreamDestroy(f->stream);
    if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
    delete f;
}

int wdx_fp_set_numpy1_promotion(wdx_fp* f, int on) {
    if (!f) return fail(WDX_ERR_INVALID, "NULL fingerprint handle");
    std::lock_guard<std::mutex> lk(f->mu);
    f->cfg.numpy1_promotion = on != 0;
    return WDX_OK;
}

int wdx_fp_set_resume_status(wdx_fp* f, int32_t status) {
    if (!f) return fail(WDX_ERR_INVALID, "NULL fingerprint handle");
    if (status < 0 || status > FP_FAIL_CONSENSUS) return fail(WDX_ERR_INVALID, "resume status %d", status);
    std::lock_guard<std::mutex> lk(f->mu);
    f->resume_status = status;
    return WDX_OK;
}

int wdx_fp_set_long_slice_len(wdx_fp* f, int32_t len) {
    if (!f) return fail(WDX_ERR_INVALID, "NULL fingerprint handle");
    if (len < 0 || len > FP_MAX_LEN) return fail(WDX_ERR_INVALID, "long_slice_len=%d outside [0,%d]", len, FP_MAX_LEN);
    std::lock_guard<std::mutex> lk(f->mu);
    f->long_slice_len = len;
    return WDX_OK;
}

int wdx_fp_set_consensus(wdx_fp* f, const wdx_fp_consensus* cc) {
    if (!f) return fail(WDX_ERR_INVALID, "NULL fingerprint handle");
    std::lock_guard<std::mutex> lk(f->mu);
    if (!cc || cc->query_len == 0) {  // back to the plain segmentation
        f->cfg.cons_len = 0;
        return WDX_OK;
    }
    if (!cc->query || cc->query_len < 1 || cc->query_len > FP_MAX_QUERY)
        return fail(WDX_ERR_INVALID, "consensus query of %d values outside [1,%d]", cc->query_len, FP_MAX_QUERY);
    if (cc->barcode_segm_events < 1 || cc->barcode_segm_events > FP_MAX_EVENTS)
        return fail(WDX_ERR_INVALID, "barcode_segm_events=%d outside [1,%d]", cc->barcode_segm_events, FP_MAX_EVENTS);
    if (!(cc->penalty >= 0) || cc->psi_query_begin < 0 || cc->psi_series_begin < 0)
        return fail(WDX_ERR_INVALID, "bad consensus penalty / psi");
    CUDA_TRY(cudaSetDevice(f->device));
    int rc = f->cons_query.reserve((size_t)cc->query_len * sizeof(double));
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    CUDA_TRY(cudaMemcpy(f->cons_query.p, cc->query, (size_t)cc->query_len * sizeof(double), cudaMemcpyHostToDevice));
    f->cfg.cons_len = cc->query_len;
    f->cfg.cons_seg_events = cc->barcode_segm_events;
    f->cfg.cons_pen2 = cc->penalty * cc->penalty;
    f->cfg.cons_psi_q = cc->psi_query_begin;
    f->cfg.cons_psi_s = cc->psi_series_begin;
    f->cfg.cons_ub_start = cc->ub_start;
    f->cfg.cons_lb_end = cc->lb_end;
    f->cfg.cons_ub_end = cc->ub_end;
    return WDX_OK;
}

int wdx_fp_extract_ex(wdx_fp* f, const float* signals, int64_t n, int64_t stride, const int32_t* sig_len,
                      const int64_t* adapter_start, const int64_t* adapter_end, const uint8_t* detect_ok,
                      int clip_in_place, double* fpt, int64_t* dwell, double* stats, int32_t* status, int32_t* cons,
                      void* stream) {
    if (cons && f && f->cfg.cons_len == 0) return fail(WDX_ERR_INVALID, "`cons` needs wdx_fp_set_consensus first");
    FpCall c{};
    c.signals = signals; c.n = n; c.stride = stride; c.sig_len = sig_len; c.a0 = adapter_start; c.a1 = adapter_end;
    c.ok = detect_ok; c.clip_in_place = clip_in_place; c.fpt = fpt; c.dwell = dwell; c.stats = stats; c.status = status;
    c.cons = cons; c.m = nullptr; c.user_stream = (cudaStream_t)stream;
    return run(f, c);
}

int wdx_fp_extract(wdx_fp* f, const float* signals, int64_t n, int64_t stride, const int32_t* sig_len,
                   const int64_t* adapter_start, const int64_t* adapter_end, const uint8_t* detect_ok,
                   int clip_in_place, double* fpt, int64_t* dwell, double* stats, int32_t* status, void* stream) {
    return wdx_fp_extract_ex(f, signals, n, stride, sig_len, adapter_start, adapter_end, detect_ok, clip_in_place, fpt,
                             dwell, stats, status, nullptr, stream);
}

int wdx_fp_predict(wdx_fp* f, wdx_model* m, const float* signals, int64_t n, int64_t stride, const int32_t* sig_len,
                   const int64_t* adapter_start, const int64_t* adapter_end, const uint8_t* detect_ok, int mode,
                   int64_t* labels, double* conf, double* prob, uint8_t* flags, double* fpt, int32_t* status,
                   void* stream) {
    if (!m) return fail(WDX_ERR_INVALID, "NULL model");
    FpCall c{};
    c.signals = signals; c.n = n; c.stride = stride; c.sig_len = sig_len; c.a0 = adapter_start; c.a1 = adapter_end;
    c.ok = detect_ok; c.clip_in_place = 0; c.fpt = fpt; c.dwell = nullptr; c.stats = nullptr; c.status = status;
    c.cons = nullptr; c.m = m; c.mode = mode; c.labels = labels; c.conf = conf; c.prob = prob; c.flags = flags;
    c.user_stream = (cudaStream_t)stream;
    return run(f, c);
}

int wdx_fp_enable_timing(wdx_fp* f, int on) {
    if (!f) return fail(WDX_ERR_INVALID, "NULL fingerprint handle");
    f->timing = on != 0;
    return WDX_OK;
}

int wdx_fp_last_kernel_ms(wdx_fp* f, double* ms, int* launches) {
    if (!f || !ms) return fail(WDX_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(f->mu);
    CUDA_TRY(cudaSetDevice(f->device));
    double tot = 0;
    for (size_t i = 0; i < f->tev_used; i++) {
        CUDA_TRY(cudaEventSynchronize(f->tev[i].second));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, f->tev[i].first, f->tev[i].second));
        tot += t;
    }
    *ms = tot;
    if (launches) *launches = (int)f->tev_used;
    return WDX_OK;
}

#ifdef WDX_FP_PROF
// experiments only: cycles per phase summed over all CTAs since the last reset
int wdx_fp_prof_dump(unsigned long long* out32, int reset) {
    if (out32) CUDA_TRY(cudaMemcpyFromSymbol(out32, wdx::g_fp_prof, 32 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long z[32] = {};
        CUDA_TRY(cudaMemcpyToSymbol(wdx::g_fp_prof, z, sizeof z));
    }
    return WDX_OK;
}
#endif
}  // extern "C"
