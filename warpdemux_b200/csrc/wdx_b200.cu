// wdx_b200.cu — C-ABI implementation (include/wdx_b200.h): model handle,
// staging, launch plumbing.  All arithmetic is in the kernels
// (fused_kernels.cuh, dtw_band.cuh).  There is no CPU compute path here: if
// CUDA is unavailable every entry point returns WDX_ERR_CUDA.
#include "../../include/wdx_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "dtw_wavefront.cuh"
#include "fused_kernels.cuh"
#include "wdx_internal.cuh"

#ifndef WDX_DEFAULT_EXACT_VARIANT
#define WDX_DEFAULT_EXACT_VARIANT 2
#endif
#ifndef WDX_DEFAULT_FAST_VARIANT
#define WDX_DEFAULT_FAST_VARIANT 9
#endif

using namespace wdx;

namespace wdx {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return WDX_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        return fail(WDX_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return WDX_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

int HostBuf::reserve(size_t bytes) {
    if (bytes <= cap) return WDX_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) {
        p = nullptr;
        return fail(WDX_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = bytes;
    return WDX_OK;
}
void HostBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

int mem_kind(const void* p) {
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return 2;
    if (at.type == cudaMemoryTypeHost) return 1;
    return 0;
}

}  // namespace wdx

namespace {
bool is_device_ptr(const void* p, int) { return wdx::mem_kind(p) == 2; }
}  // namespace

namespace {

template <typename T>
int upload(wdx_model* m, const std::vector<T>& h, const T** out) {
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::max<size_t>(h.size() * sizeof(T), 16)));
    m->owned.push_back(d);
    CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T*>(d);
    return WDX_OK;
}

using FusedFn = void (*)(const ModelDev, const PredictArgs);

struct FusedChoice {
    FusedFn fn;
    bool x2, accs;
    int km1_cap;
};

// FAST-path variants (selected by measurement, profiles/r01_fast_variants.jsonl and
// profiles/r01_ubench_instruction_mix.txt; WDX_FAST_VARIANT overrides for experiments):
//   0 scalar recurrence,               4 CTAs/SM, sums in registers
//   2 packed f32x2,                    3 CTAs/SM, sums in shared memory
//   6 packed f32x2 in offset (E) form, 4 CTAs/SM, sums in shared memory
//   7 packed f32x2 in offset (E) form, 5 CTAs/SM, sums in shared memory (spills once k-1 > 4)
//   9 (default) 7 for models with at most 5 classes, else 6
int fast_variant() {
    static int v = [] {
        const char* e = getenv("WDX_FAST_VARIANT");
        return e ? atoi(e) : WDX_DEFAULT_FAST_VARIANT;
    }();
    return v;
}

template <bool EXACT, int L_, int W_, int X2, int MINB, bool ACCS>
FusedChoice pick_km1(int km1) {
    FusedFn f;
    int cap;
    if (km1 <= 4) { f = dtw_svc_kernel<EXACT, L_, W_, 4, X2, MINB, ACCS>; cap = 4; }
    else if (km1 <= 6) { f = dtw_svc_kernel<EXACT, L_, W_, 6, X2, MINB, ACCS>; cap = 6; }
    else if (km1 <= 10) { f = dtw_svc_kernel<EXACT, L_, W_, 10, X2, MINB, ACCS>; cap = 10; }
    else { f = dtw_svc_kernel<EXACT, L_, W_, 16, X2, MINB, ACCS>; cap = 16; }
    return FusedChoice{f, X2 != 0, ACCS, cap};
}

FusedChoice pick_fused(const wdx_model* m, bool exact) {
    const int km1 = m->k - 1;
    if (!m->specialised) {
        if (exact) return FusedChoice{dtw_svc_kernel<true, 0, 0, 16, 0, 3, false>, false, false, 16};
        return FusedChoice{dtw_svc_kernel<false, 0, 0, 16, 0, 4, false>, false, false, 16};
    }
    if (exact) {  // WDX_EXACT_VARIANT: 0 = 3 CTAs/SM, sums in registers; 1 = 4 CTAs/SM, sums in shared memory;
                  // 2 (default) = 0 for models with at most 5 classes, else 1 (measured: WDX4 1930 vs 1876, WDX10 1806 vs 1892 GCUPS)
        static const int ev = [] { const char* e = getenv("WDX_EXACT_VARIANT"); return e ? atoi(e) : WDX_DEFAULT_EXACT_VARIANT; }();
        if (ev == 0 || (ev == 2 && km1 <= 4)) return pick_km1<true, 25, 15, 0, 3, false>(km1);
        return pick_km1<true, 25, 15, 0, 4, true>(km1);
    }
    switch (fast_variant()) {
        case 0: return pick_km1<false, 25, 15, 0, 4, false>(km1);
        case 2: return pick_km1<false, 25, 15, 1, 3, true>(km1);
        case 6: return pick_km1<false, 25, 15, 2, 4, true>(km1);
        case 7: return pick_km1<false, 25, 15, 2, 5, true>(km1);
        default:
            if (km1 <= 4) return FusedChoice{dtw_svc_kernel<false, 25, 15, 4, 2, 5, true>, true, true, 4};
            return pick_km1<false, 25, 15, 2, 4, true>(km1);
    }
}

// dynamic shared memory: [ SV/coef stages x2 + mbarriers | aliased fingerprint staging ][ running sums ]
size_t fused_smem_bytes(const wdx_model* m, bool exact, bool x_f32, const FusedChoice& ch, int* acc_off) {
    const size_t sv_row = exact ? (size_t)m->dev.ldd * 8 : (ch.x2 ? (size_t)m->dev.ldp * 4 : (size_t)m->dev.ldf * 4);
    const size_t stage = TILE_SV * (sv_row + (size_t)m->dev.ldc * 8);
    const size_t pipe = 2 * stage + 16;
    const size_t xs = (size_t)CTA_THREADS * m->L * (x_f32 ? 4 : 8);
    size_t base = (std::max(pipe, xs) + 15) & ~(size_t)15;
    *acc_off = (int)base;
    if (ch.accs) base += (size_t)ch.km1_cap * CTA_THREADS * 8;
    return base;
}

int launch_fused(wdx_model* m, bool exact, const PredictArgs& pa_in, int64_t grid_rows, cudaStream_t st) {
    const FusedChoice ch = pick_fused(m, exact);
    FusedFn fn = ch.fn;
    PredictArgs pa = pa_in;
    int acc_off = 0;
    const size_t smem = fused_smem_bytes(m, exact, pa.x_is_f32 != 0, ch, &acc_off);
    pa.acc_smem_offset = acc_off;
    CUDA_TRY(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((grid_rows + CTA_THREADS - 1) / CTA_THREADS), (unsigned)pa.n_splits);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (m->timing) {
        if (m->tev_used == m->tev.size()) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            m->tev.emplace_back(a, b);
            m->tev_exact.push_back(0);
        }
        e0 = m->tev[m->tev_used].first;
        e1 = m->tev[m->tev_used].second;
        m->tev_exact[m->tev_used] = exact ? 1 : 0;
        m->tev_used++;
        CUDA_TRY(cudaEventRecord(e0, st));
    }
    fn<<<grid, CTA_THREADS, smem, st>>>(m->dev, pa);
    CUDA_TRY(cudaGetLastError());
    if (m->timing) CUDA_TRY(cudaEventRecord(e1, st));
    g_launches++;
    return WDX_OK;
}

constexpr int64_t FINISH_WARP_MAX_ROWS = 16384;
bool finish_warp_disabled() {
    static const bool off = [] { const char* e = getenv("WDX_NO_FINISH_WARP"); return e && e[0] == '1'; }();
    return off;
}

int launch_finish(wdx_model* m, const FinishArgs& fa_in, int64_t grid_rows, cudaStream_t st) {
    FinishArgs fa = fa_in;
    if (fa.n_splits > 1) {  // fold the SV ranges first (parallel over pairs x reads), then finish as one range
        dim3 grid((unsigned)((grid_rows + 31) / 32), (unsigned)m->n_pairs);   // 32 reads x FOLD_SUB sub-sums per CTA
        svc_fold_splits_kernel<<<grid, 256, 0, st>>>(m->dev, const_cast<double*>(fa.part), fa.part_stride, fa.n_splits,
                                                    fa.sv_per_split, fa.n_idx, fa.n);
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        fa.n_splits = 1;
        fa.sv_per_split = m->n_sv;
    }
    if (grid_rows <= FINISH_WARP_MAX_ROWS && !finish_warp_disabled()) {   // live-sized batch: one warp per read (latency)
        const unsigned blocks = (unsigned)((grid_rows + FINISH_WARPS - 1) / FINISH_WARPS);
        svc_finish_warp_kernel<<<blocks, FINISH_WARPS * 32, 0, st>>>(m->dev, fa);
    } else {
        const unsigned blocks = (unsigned)((grid_rows + 127) / 128);
        svc_finish_kernel<<<blocks, 128, 0, st>>>(m->dev, fa);
    }
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return WDX_OK;
}

// A thread walks its SV range serially, so for a live-sized batch the range length IS the latency of the call:
// 41 SVs per range (64 ranges of WDX10) cost 74 us in FAST and 330 us in EXACT arithmetic, measured; shorter
// ranges are paid for by the fold over the partial sums (profiles/r01o_latency_sweep.jsonl: best at 128 / 512).
void choose_splits(const wdx_model* m, int64_t n, int max_splits, int* n_splits, int* sv_per_split) {
    const int64_t ctas_x = (n + CTA_THREADS - 1) / CTA_THREADS;
    const int64_t target = (int64_t)m->sm_count * 4 * 2;  // two full waves of 4 CTAs/SM
    int splits = 1;
    if (ctas_x < target) splits = (int)std::min<int64_t>(max_splits, (target + ctas_x - 1) / ctas_x);
    if (m->forced_splits > 0) splits = m->forced_splits;
    splits = std::max(1, std::min(splits, m->n_sv));
    int per = (m->n_sv + splits - 1) / splits;
    splits = (m->n_sv + per - 1) / per;
    *n_splits = splits;
    *sv_per_split = per;
}

}  // namespace

// One chunk, everything on the device: Xd [n][L] -> labels/conf/prob/flags (device).
int wdx::predict_chunk_device(wdx_model* m, const void* Xd, int x_is_f32, int64_t n, int mode, int64_t* labels_d,
                         double* conf_d, double* prob_d, uint8_t* flags_d, float* dist_d, cudaStream_t st) {
    int n_splits, per;
    choose_splits(m, n, mode == WDX_MODE_EXACT_F64 ? 512 : 128, &n_splits, &per);
    const int64_t stride = (n + 31) & ~(int64_t)31;
    int rc = m->part.reserve((size_t)n_splits * m->n_pairs * stride * sizeof(double));
    if (rc) return rc;
    const bool exact_first = (mode == WDX_MODE_EXACT_F64);
    const bool guarded = (mode == WDX_MODE_FAST_F32_GUARDED);

    PredictArgs pa{};
    pa.X = Xd;
    pa.x_is_f32 = x_is_f32;
    pa.read_idx = nullptr;
    pa.n_idx = nullptr;
    pa.n = n;
    pa.part_stride = stride;
    pa.part = (double*)m->part.p;
    pa.dist = dist_d;
    pa.n_splits = n_splits;
    pa.sv_per_split = per;
    rc = launch_fused(m, exact_first, pa, n, st);
    if (rc) return rc;

    FinishArgs fa{};
    fa.part = (const double*)m->part.p;
    fa.part_stride = stride;
    fa.n_splits = n_splits;
    fa.sv_per_split = per;
    fa.n = n;
    fa.labels = labels_d;
    fa.conf = conf_d;
    fa.prob = prob_d;
    fa.flags = flags_d;
    fa.guard = m->guard;
    // GUARDED: reads close to a decision boundary are listed by the finishing
    // kernel (count stays on the device) and re-run in EXACT_F64.  The list is
    // capped at max(4096, n/64) entries; the re-run is sized for that worst case
    // (idle CTAs exit at once) and cut into SV ranges so that a handful of reads
    // still fills the GPU.  Reads that do not fit keep their FAST result and get
    // WDX_FLAG_GUARD_OVERFLOW.
    const int64_t cap = std::min<int64_t>(n, std::max<int64_t>(4096, n / 64));
    if (guarded) {
        rc = m->near_idx.reserve((size_t)cap * sizeof(int));
        if (rc) return rc;
        rc = m->counters.reserve(64);
        if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(m->counters.p, 0, 64, st));
        fa.near_idx = (int*)m->near_idx.p;
        fa.near_count = (int*)m->counters.p;
        fa.near_cap = (int)cap;
    }
    rc = launch_finish(m, fa, n, st);
    if (rc) return rc;

    if (guarded) {
        int s2, per2;
        // expect the list to be mostly empty; short ranges (latency) only for live-sized batches - the partial-sum planes
        // of the re-run are sized splits x pairs x cap
        choose_splits(m, std::max<int64_t>(1, cap / 16), cap <= 1024 ? 512 : (cap <= 4096 ? 256 : 64), &s2, &per2);
        const int64_t stride2 = (cap + 31) & ~(int64_t)31;
        rc = m->part2.reserve((size_t)s2 * m->n_pairs * stride2 * sizeof(double));
        if (rc) return rc;
        PredictArgs pb = pa;
        pb.read_idx = (const int*)m->near_idx.p;
        pb.n_idx = (const int*)m->counters.p;
        pb.n = cap;
        pb.dist = nullptr;
        pb.part = (double*)m->part2.p;
        pb.part_stride = stride2;
        pb.n_splits = s2;
        pb.sv_per_split = per2;
        rc = launch_fused(m, true, pb, cap, st);
        if (rc) return rc;
        FinishArgs fb = fa;
        fb.part = (const double*)m->part2.p;
        fb.part_stride = stride2;
        fb.n_splits = s2;
        fb.sv_per_split = per2;
        fb.read_idx = (const int*)m->near_idx.p;
        fb.n_idx = (const int*)m->counters.p;
        fb.n = cap;
        fb.near_idx = nullptr;
        fb.near_count = nullptr;
        fb.flag_or = WDX_FLAG_RECOMPUTED;
        rc = launch_finish(m, fb, cap, st);
        if (rc) return rc;
    }
    return WDX_OK;
}

extern "C" {

const char* wdx_last_error(void) { return g_err; }
const char* wdx_version(void) { return "wdx_b200 0.1 (sm_100a)"; }
int64_t wdx_kernel_launch_count(void) { return g_launches.load(); }

int wdx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int wdx_model_create(const double* sv, int n_sv, int L, const int32_t* n_sv_class, int k, const double* dual_coef,
                     const double* rho, const double* probA, const double* probB, const double* thresholds,
                     const int64_t* label_map, int window, double penalty, double gamma, int pwr_dist, int device,
                     wdx_model** out) {
    if (!out) return fail(WDX_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!sv || !n_sv_class || !dual_coef || !rho || !probA || !probB || !thresholds || !label_map)
        return fail(WDX_ERR_INVALID, "NULL model array");
    if (k < 2 || k > MAXK) return fail(WDX_ERR_INVALID, "k=%d outside [2,%d]", k, MAXK);
    if (L < 1 || L > MAXL) return fail(WDX_ERR_INVALID, "L=%d outside [1,%d]", L, MAXL);
    if (n_sv < 1) return fail(WDX_ERR_INVALID, "n_sv=%d", n_sv);
    if (pwr_dist < 0) return fail(WDX_ERR_INVALID, "pwr_dist=%d", pwr_dist);
    if (window < 0) return fail(WDX_ERR_INVALID, "window=%d", window);
    long tot = 0;
    for (int c = 0; c < k; c++) {
        if (n_sv_class[c] < 0) return fail(WDX_ERR_INVALID, "negative n_sv_class");
        tot += n_sv_class[c];
    }
    if (tot != n_sv) return fail(WDX_ERR_INVALID, "n_sv_class sums to %ld, n_sv=%d", tot, n_sv);
    int ndev = wdx_device_count();
    if (ndev <= 0) return fail(WDX_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(WDX_ERR_INVALID, "device %d of %d", device, ndev);
    CUDA_TRY(cudaSetDevice(device));

    wdx_model* m = new (std::nothrow) wdx_model();
    if (!m) return fail(WDX_ERR_NOMEM, "host allocation failed");
    m->device = device;
    m->k = k;
    m->L = L;
    m->n_sv = n_sv;
    m->n_pairs = k * (k - 1) / 2;
    const int w_eff = (window <= 0 || window > L) ? L : window;
    m->specialised = (L == 25 && w_eff == 15);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->sm_count = prop.multiProcessorCount;

    ModelDev& d = m->dev;
    d.n_sv = n_sv;
    d.L = L;
    d.k = k;
    d.n_pairs = m->n_pairs;
    d.ldf = (L + 3) & ~3;
    d.ldd = (L + 1) & ~1;
    d.ldc = (k - 1 + 1) & ~1;
    d.ldp = (std::max(2 * (L - 1), 4) + 3) & ~3;
    d.window = w_eff;
    d.p2 = penalty * penalty;
    d.gamma = gamma;
    d.pwr_dist = pwr_dist;
    d.class_start[0] = 0;
    for (int c = 0; c < k; c++) d.class_start[c + 1] = d.class_start[c] + n_sv_class[c];
    for (int c = k + 1; c <= MAXK; c++) d.class_start[c] = n_sv;

    std::vector<float> svf((size_t)n_sv * d.ldf, 0.f);
    std::vector<double> svd((size_t)n_sv * d.ldd, 0.0);
    std::vector<double> coef((size_t)n_sv * d.ldc, 0.0);
    std::vector<float> svp((size_t)n_sv * d.ldp, 0.f);
    for (int s = 0; s < n_sv; s++) {
        for (int t = 1; t < L; t++) {  // pre-paired layout of the packed recurrence: (s[t], s[t-1])
            svp[(size_t)s * d.ldp + 2 * (t - 1)] = (float)sv[(size_t)s * L + t];
            svp[(size_t)s * d.ldp + 2 * (t - 1) + 1] = (float)sv[(size_t)s * L + t - 1];
        }
        for (int j = 0; j < L; j++) {
            svf[(size_t)s * d.ldf + j] = (float)sv[(size_t)s * L + j];
            svd[(size_t)s * d.ldd + j] = sv[(size_t)s * L + j];
        }
        for (int r = 0; r < k - 1; r++) coef[(size_t)s * d.ldc + r] = dual_coef[(size_t)r * n_sv + s];
    }
    int rc = WDX_OK;
    auto bail = [&](int code) {
        wdx_model_destroy(m);
        return code;
    };
    if ((rc = upload(m, svf, &d.sv_f32))) return bail(rc);
    if ((rc = upload(m, svd, &d.sv_f64))) return bail(rc);
    if ((rc = upload(m, svp, &d.sv_x2))) return bail(rc);
    if ((rc = upload(m, coef, &d.coef))) return bail(rc);
    if ((rc = upload(m, std::vector<double>(rho, rho + m->n_pairs), &d.rho))) return bail(rc);
    if ((rc = upload(m, std::vector<double>(probA, probA + m->n_pairs), &d.probA))) return bail(rc);
    if ((rc = upload(m, std::vector<double>(probB, probB + m->n_pairs), &d.probB))) return bail(rc);
    if ((rc = upload(m, std::vector<double>(thresholds, thresholds + k), &d.thresholds))) return bail(rc);
    if ((rc = upload(m, std::vector<int64_t>(label_map, label_map + k), &d.label_map))) return bail(rc);

    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(WDX_ERR_CUDA, "cudaStreamCreate failed"));
    for (int i = 0; i < 2; i++) {
        if (cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_free[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_d2h[i], cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(WDX_ERR_CUDA, "cudaEventCreate failed"));
    }
    *out = m;
    return WDX_OK;
}

void wdx_model_destroy(wdx_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
    for (void* p : m->owned) cudaFree(p);
    m->part.release();
    m->part2.release();
    m->near_idx.release();
    m->counters.release();
    m->small_dev.release();
    m->small_pin.release();
    for (int i = 0; i < 2; i++) {
        m->xdev[i].release();
        m->xpin[i].release();
        if (m->ev_h2d[i]) cudaEventDestroy(m->ev_h2d[i]);
        if (m->ev_free[i]) cudaEventDestroy(m->ev_free[i]);
        if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
        if (m->ev_d2h[i]) cudaEventDestroy(m->ev_d2h[i]);
        m->lab_dev[i].release();
        m->conf_dev[i].release();
        m->prob_dev[i].release();
        m->flag_dev[i].release();
        m->dist_dev[i].release();
    }
    for (auto& e : m->tev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    delete m;
}

int wdx_model_set_guard(wdx_model* m, double guard) {
    if (!m || !(guard >= 0)) return fail(WDX_ERR_INVALID, "bad guard");
    m->guard = guard;
    return WDX_OK;
}

int wdx_model_set_chunk_reads(wdx_model* m, int64_t chunk) {
    if (!m || chunk < 1) return fail(WDX_ERR_INVALID, "bad chunk");
    m->chunk_reads = chunk;
    return WDX_OK;
}

int wdx_model_set_sv_splits(wdx_model* m, int splits) {
    if (!m || splits < 0) return fail(WDX_ERR_INVALID, "bad splits");
    m->forced_splits = splits;
    return WDX_OK;
}

int wdx_model_enable_timing(wdx_model* m, int on) {
    if (!m) return fail(WDX_ERR_INVALID, "NULL model");
    m->timing = on != 0;
    return WDX_OK;
}

static int kernel_ms_impl(wdx_model* m, int which, double* ms, int* launches) {
    if (!m || !ms) return fail(WDX_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(m->mu);
    CUDA_TRY(cudaSetDevice(m->device));
    double tot = 0;
    int cnt = 0;
    for (size_t i = 0; i < m->tev_used; i++) {
        if (which >= 0 && (int)m->tev_exact[i] != which) continue;
        CUDA_TRY(cudaEventSynchronize(m->tev[i].second));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, m->tev[i].first, m->tev[i].second));
        tot += t;
        cnt++;
    }
    *ms = tot;
    if (launches) *launches = cnt;
    return WDX_OK;
}

int wdx_model_last_kernel_ms(wdx_model* m, double* ms, int* launches) { return kernel_ms_impl(m, -1, ms, launches); }

int wdx_model_last_kernel_ms_mode(wdx_model* m, int exact, double* ms, int* launches) {
    return kernel_ms_impl(m, exact ? 1 : 0, ms, launches);
}

namespace {
constexpr int64_t SMALL_BATCH_MAX = 8192;

bool small_path_disabled() {
    static const bool off = [] { const char* e = getenv("WDX_NO_SMALL_PATH"); return e && e[0] == '1'; }();
    return off;
}

// Live-sized batch with every buffer in host memory (the read-until caller, live_balancing/worker.py:117-120):
// input and results share ONE device block mirrored by ONE pinned block, so the call costs one H2D, the
// kernels, one D2H and one stream synchronisation instead of a copy and a wait per output array.
int predict_small_host(wdx_model* m, const void* X, int64_t n, int esz, int x_kind, int mode, int64_t* labels, double* conf,
                       double* prob, uint8_t* flags, cudaStream_t st) {
    const int L = m->L, k = m->k;
    const size_t x_bytes = ((size_t)n * L * esz + 15) & ~(size_t)15;
    const size_t o_lab = x_bytes, o_conf = o_lab + (size_t)n * 8, o_prob = o_conf + (size_t)n * 8,
                 o_flag = o_prob + (size_t)n * k * 8, total = (o_flag + (size_t)n + 15) & ~(size_t)15;
    int rc;
    if ((rc = m->small_dev.reserve(total)) || (rc = m->small_pin.reserve(total))) return rc;
    char* d = (char*)m->small_dev.p;
    char* h = (char*)m->small_pin.p;
    const void* src = X;
    if (x_kind == 0) {   // pageable: stage through the pinned block
        std::memcpy(h, X, (size_t)n * L * esz);
        src = h;
    }
    CUDA_TRY(cudaMemcpyAsync(d, src, (size_t)n * L * esz, cudaMemcpyHostToDevice, st));
    rc = wdx::predict_chunk_device(m, d, esz == 4, n, mode, (int64_t*)(d + o_lab), (double*)(d + o_conf), (double*)(d + o_prob),
                                   (uint8_t*)(d + o_flag), nullptr, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h + o_lab, d + o_lab, total - o_lab, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::memcpy(labels, h + o_lab, (size_t)n * 8);
    if (conf) std::memcpy(conf, h + o_conf, (size_t)n * 8);
    if (prob) std::memcpy(prob, h + o_prob, (size_t)n * k * 8);
    if (flags) std::memcpy(flags, h + o_flag, (size_t)n);
    return WDX_OK;
}
}  // namespace

int wdx_predict(wdx_model* m, const void* X, int64_t n, int x_dtype, int mode, int64_t* labels, double* conf,
                double* prob, uint8_t* flags, float* dist, void* stream) {
    if (!m) return fail(WDX_ERR_INVALID, "NULL model");
    if (n < 0) return fail(WDX_ERR_INVALID, "n=%lld", (long long)n);
    if (n == 0) return WDX_OK;
    if (!X || !labels) return fail(WDX_ERR_INVALID, "X and labels are required");
    if (x_dtype != WDX_F64 && x_dtype != WDX_F32) return fail(WDX_ERR_INVALID, "x_dtype=%d", x_dtype);
    if (mode < WDX_MODE_EXACT_F64 || mode > WDX_MODE_FAST_F32_GUARDED) return fail(WDX_ERR_INVALID, "mode=%d", mode);
    std::lock_guard<std::mutex> lk(m->mu);
    CUDA_TRY(cudaSetDevice(m->device));
    m->tev_used = 0;

    const int esz = (x_dtype == WDX_F32) ? 4 : 8;
    const int L = m->L, k = m->k;
    const int x_kind = mem_kind(X);
    const bool x_dev = x_kind == 2;
    const bool lab_dev = mem_kind(labels) == 2;
    const bool conf_devp = conf && mem_kind(conf) == 2;
    const bool prob_devp = prob && mem_kind(prob) == 2;
    const bool flag_devp = flags && mem_kind(flags) == 2;
    const bool dist_devp = dist && mem_kind(dist) == 2;
    const bool out_host = !lab_dev || (conf && !conf_devp) || (prob && !prob_devp) || (flags && !flag_devp) ||
                          (dist && !dist_devp);
    const bool any_host = !x_dev || out_host;
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    if (n <= SMALL_BATCH_MAX && !x_dev && !lab_dev && !conf_devp && !prob_devp && !flag_devp && !dist && !small_path_disabled())
        return predict_small_host(m, X, n, esz, x_kind, mode, labels, conf, prob, flags, st);

    int64_t chunk = m->chunk_reads;
    if (any_host) chunk = std::min<int64_t>(chunk, (int64_t)1 << 21);  // finer pipeline when copies are involved
    if (dist) chunk = std::min<int64_t>(chunk, (int64_t)1 << 16);
    chunk = std::min<int64_t>(chunk, n);
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    const int nb = n_chunks > 1 ? 2 : 1;
    int rc;
    // device-side result buffers (double-buffered) for whatever lives on the host
    for (int b = 0; b < nb; b++) {
        if (!lab_dev && (rc = m->lab_dev[b].reserve((size_t)chunk * 8))) return rc;
        if (conf && !conf_devp && (rc = m->conf_dev[b].reserve((size_t)chunk * 8))) return rc;
        if (prob && !prob_devp && (rc = m->prob_dev[b].reserve((size_t)chunk * k * 8))) return rc;
        if (flags && !flag_devp && (rc = m->flag_dev[b].reserve((size_t)chunk))) return rc;
        if (dist && !dist_devp && (rc = m->dist_dev[b].reserve((size_t)chunk * m->n_sv * 4))) return rc;
        if (!x_dev) {
            if ((rc = m->xdev[b].reserve((size_t)chunk * L * esz))) return rc;
            if (x_kind == 0 && (rc = m->xpin[b].reserve((size_t)chunk * L * esz))) return rc;
        }
    }

    // Host-resident input: chunk c+1 goes user memory -> (pinned staging ->) device
    // on the copy stream while chunk c computes.
    auto stage_in = [&](int64_t c) -> int {
        const int b = (int)(c & 1);
        const int64_t r0 = c * chunk, cn = std::min(chunk, n - r0);
        const size_t bytes = (size_t)cn * L * esz;
        const char* src = (const char*)X + (size_t)r0 * L * esz;
        CUDA_TRY(cudaEventSynchronize(m->ev_free[b]));  // the kernels that read this buffer (2 chunks ago) are done
        if (x_kind == 0) {
            std::memcpy(m->xpin[b].p, src, bytes);
            src = (const char*)m->xpin[b].p;
        }
        CUDA_TRY(cudaMemcpyAsync(m->xdev[b].p, src, bytes, cudaMemcpyHostToDevice, m->copy_stream));
        CUDA_TRY(cudaEventRecord(m->ev_h2d[b], m->copy_stream));
        return WDX_OK;
    };
    // Results of chunk c back to the caller's host buffers on the copy stream,
    // issued after chunk c+1's kernels are already queued.
    auto drain = [&](int64_t c) -> int {
        const int b = (int)(c & 1);
        const int64_t r0 = c * chunk, cn = std::min(chunk, n - r0);
        CUDA_TRY(cudaStreamWaitEvent(m->copy_stream, m->ev_done[b], 0));
        if (!lab_dev)
            CUDA_TRY(cudaMemcpyAsync(labels + r0, m->lab_dev[b].p, (size_t)cn * 8, cudaMemcpyDeviceToHost, m->copy_stream));
        if (conf && !conf_devp)
            CUDA_TRY(cudaMemcpyAsync(conf + r0, m->conf_dev[b].p, (size_t)cn * 8, cudaMemcpyDeviceToHost, m->copy_stream));
        if (prob && !prob_devp)
            CUDA_TRY(cudaMemcpyAsync(prob + (size_t)r0 * k, m->prob_dev[b].p, (size_t)cn * k * 8, cudaMemcpyDeviceToHost,
                                     m->copy_stream));
        if (flags && !flag_devp)
            CUDA_TRY(cudaMemcpyAsync(flags + r0, m->flag_dev[b].p, (size_t)cn, cudaMemcpyDeviceToHost, m->copy_stream));
        if (dist && !dist_devp)
            CUDA_TRY(cudaMemcpyAsync(dist + (size_t)r0 * m->n_sv, m->dist_dev[b].p, (size_t)cn * m->n_sv * 4,
                                     cudaMemcpyDeviceToHost, m->copy_stream));
        CUDA_TRY(cudaEventRecord(m->ev_d2h[b], m->copy_stream));
        return WDX_OK;
    };
    if (!x_dev) {
        CUDA_TRY(cudaEventRecord(m->ev_free[0], st));
        CUDA_TRY(cudaEventRecord(m->ev_free[1], st));
        if ((rc = stage_in(0))) return rc;
    }
    if (out_host) {
        CUDA_TRY(cudaEventRecord(m->ev_d2h[0], m->copy_stream));
        CUDA_TRY(cudaEventRecord(m->ev_d2h[1], m->copy_stream));
    }

    for (int64_t c = 0; c < n_chunks; c++) {
        const int b = (int)(c & 1);
        const int64_t r0 = c * chunk, cn = std::min(chunk, n - r0);
        const void* Xd;
        if (x_dev) {
            Xd = (const char*)X + (size_t)r0 * L * esz;
        } else {
            Xd = m->xdev[b].p;
            CUDA_TRY(cudaStreamWaitEvent(st, m->ev_h2d[b], 0));
        }
        if (out_host) CUDA_TRY(cudaStreamWaitEvent(st, m->ev_d2h[b], 0));  // result buffers b were drained
        int64_t* lab_d = lab_dev ? labels + r0 : (int64_t*)m->lab_dev[b].p;
        double* conf_d = conf ? (conf_devp ? conf + r0 : (double*)m->conf_dev[b].p) : nullptr;
        double* prob_d = prob ? (prob_devp ? prob + (size_t)r0 * k : (double*)m->prob_dev[b].p) : nullptr;
        uint8_t* flag_d = flags ? (flag_devp ? flags + r0 : (uint8_t*)m->flag_dev[b].p) : nullptr;
        float* dist_d = dist ? (dist_devp ? dist + (size_t)r0 * m->n_sv : (float*)m->dist_dev[b].p) : nullptr;
        rc = predict_chunk_device(m, Xd, esz == 4, cn, mode, lab_d, conf_d, prob_d, flag_d, dist_d, st);
        if (rc) return rc;
        if (out_host) CUDA_TRY(cudaEventRecord(m->ev_done[b], st));
        if (!x_dev) {
            CUDA_TRY(cudaEventRecord(m->ev_free[b], st));
            if (c + 1 < n_chunks && (rc = stage_in(c + 1))) return rc;  // overlaps the kernels just enqueued
        }
        if (out_host && c >= 1 && (rc = drain(c - 1))) return rc;
    }
    if (out_host) {
        if ((rc = drain(n_chunks - 1))) return rc;
        CUDA_TRY(cudaStreamSynchronize(m->copy_stream));
    }
    if (any_host || !stream) CUDA_TRY(cudaStreamSynchronize(st));
    return WDX_OK;
}

int wdx_distance_matrix_to(const double* X, int64_t nX, const double* Y, int64_t nY, int L, int window,
                           double penalty, int mode, void* out, int out_dtype, int device, void* stream) {
    if (!X || !Y || !out) return fail(WDX_ERR_INVALID, "NULL argument");
    if (nX < 0 || nY < 0) return fail(WDX_ERR_INVALID, "negative size");
    if (L < 1 || L > WF_MAX_L) return fail(WDX_ERR_INVALID, "L=%d outside [1,%d]", L, WF_MAX_L);
    if (out_dtype != WDX_F32 && out_dtype != WDX_F64) return fail(WDX_ERR_INVALID, "out_dtype=%d", out_dtype);
    if (mode != WDX_MODE_EXACT_F64 && mode != WDX_MODE_FAST_F32) return fail(WDX_ERR_INVALID, "mode=%d", mode);
    if (nX == 0 || nY == 0) return WDX_OK;
    int ndev = wdx_device_count();
    if (ndev <= 0) return fail(WDX_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(WDX_ERR_INVALID, "device %d of %d", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool xd = is_device_ptr(X, device), yd = is_device_ptr(Y, device), od = is_device_ptr(out, device);
    const size_t osz = out_dtype == WDX_F32 ? 4 : 8;
    DevBuf bx, by, bo;
    int rc = WDX_OK;
    const double *Xd = X, *Yd = Y;
    void* Od = out;
    auto cleanup = [&](int code) {
        bx.release();
        by.release();
        bo.release();
        return code;
    };
    if (!xd) {
        if ((rc = bx.reserve((size_t)nX * L * 8))) return cleanup(rc);
        if (cudaMemcpyAsync(bx.p, X, (size_t)nX * L * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return cleanup(fail(WDX_ERR_CUDA, "H2D copy of X failed"));
        Xd = (const double*)bx.p;
    }
    if (!yd) {
        if ((rc = by.reserve((size_t)nY * L * 8))) return cleanup(rc);
        if (cudaMemcpyAsync(by.p, Y, (size_t)nY * L * 8, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return cleanup(fail(WDX_ERR_CUDA, "H2D copy of Y failed"));
        Yd = (const double*)by.p;
    }
    if (!od) {
        if ((rc = bo.reserve((size_t)nX * nY * osz))) return cleanup(rc);
        Od = bo.p;
    }
    const int w_eff = (window <= 0 || window > L) ? L : window;
    const bool spec = (L == 25 && w_eff == 15);
    const bool exact = mode == WDX_MODE_EXACT_F64;
    const double p2 = penalty * penalty;
    // Series longer than a register row (L > MAXL) take the warp-wide anti-diagonal wavefront (dtw_wavefront.cuh);
    // so do non-specialised shapes from L = 57: there the thread-per-pair kernel keeps its rows in local memory
    // and is 2.3x slower (profiles/r01h_wavefront_probe.jsonl).  WDX_WAVEFRONT_MIN_L overrides (probing only).
    int wf_min_l = 57;
    if (const char* ev = getenv("WDX_WAVEFRONT_MIN_L")) wf_min_l = std::max(1, std::min(MAXL + 1, atoi(ev)));
    if (L >= wf_min_l && !spec) {
        // strip width C: the one with the fewest issue slots, ~7 (14 in float64) per cell plus the per-step
        // shuffles and bookkeeping; 2 columns per lane only when one chunk of 64 columns covers the series
        int best_c = 4;
        double best_cost = 0;
        for (int c : {2, 4, 8, 16}) {
            if (c == 2 && L > 64) continue;
            const int W = 32 * c;
            double cost = 0;
            for (int jb = 0; jb < L; jb += W) {
                const int ib = std::max(0, jb - w_eff + 1), ie = std::min(L, jb + W + w_eff - 1);
                cost += (double)(ie - ib + 31) * (c * (exact ? 14.0 : 7.0) + (exact ? 45.0 : 30.0));
            }
            if (best_cost == 0 || cost < best_cost) best_cost = cost, best_c = c;
        }
        const int n_chunks = (L + 32 * best_c - 1) / (32 * best_c);
        const int64_t pairs = nX * nY;
        const int64_t max_ctas = n_chunks > 1 ? 148 * 2 : 148 * 8;
        const unsigned ctas = (unsigned)std::max<int64_t>(1, std::min<int64_t>((pairs + WF_WARPS - 1) / WF_WARPS, max_ctas));
        DevBuf be;
        auto cleanup_wf = [&](int code) {
            be.release();
            return cleanup(code);
        };
        if (n_chunks > 1 && (rc = be.reserve((size_t)ctas * WF_WARPS * 2 * L * (exact ? 8 : 4)))) return cleanup_wf(rc);
#define WDX_LAUNCH_WF(EX, CC, OT) \
    dtw_wavefront_kernel<EX, CC, OT><<<ctas, WF_THREADS, 0, st>>>(Xd, nX, Yd, nY, L, w_eff, p2, (OT*)Od, be.p)
#define WDX_LAUNCH_WF_C(EX, OT)                      \
    do {                                             \
        if (best_c == 2) WDX_LAUNCH_WF(EX, 2, OT);   \
        else if (best_c == 4) WDX_LAUNCH_WF(EX, 4, OT); \
        else if (best_c == 8) WDX_LAUNCH_WF(EX, 8, OT); \
        else WDX_LAUNCH_WF(EX, 16, OT);              \
    } while (0)
        if (exact && out_dtype == WDX_F64) WDX_LAUNCH_WF_C(true, double);
        else if (exact) WDX_LAUNCH_WF_C(true, float);
        else if (out_dtype == WDX_F64) WDX_LAUNCH_WF_C(false, double);
        else WDX_LAUNCH_WF_C(false, float);
#undef WDX_LAUNCH_WF_C
#undef WDX_LAUNCH_WF
        if (cudaGetLastError() != cudaSuccess) return cleanup_wf(fail(WDX_ERR_CUDA, "dtw_wavefront_kernel launch failed"));
        g_launches++;
        if (!od && cudaMemcpyAsync(out, Od, (size_t)nX * nY * osz, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            return cleanup_wf(fail(WDX_ERR_CUDA, "D2H copy failed"));
        cudaError_t e = cudaStreamSynchronize(st);  // the scratch is released on return
        if (e != cudaSuccess) return cleanup_wf(fail(WDX_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e)));
        return cleanup_wf(WDX_OK);
    }
    const int64_t ctas_x = (nX + CTA_THREADS - 1) / CTA_THREADS;
    int splits = (int)std::max<int64_t>(1, std::min<int64_t>((nY + TILE_SV - 1) / TILE_SV, (148 * 8 + ctas_x - 1) / ctas_x));
    int per = (int)((nY + splits - 1) / splits);
    per = ((per + TILE_SV - 1) / TILE_SV) * TILE_SV;
    splits = (int)((nY + per - 1) / per);
    dim3 grid((unsigned)ctas_x, (unsigned)splits);
    const size_t smem = (size_t)TILE_SV * L * (exact ? 8 : 4);
#define WDX_LAUNCH_DM(EX, LL, WW, OT)                                                                              \
    dtw_matrix_kernel<EX, LL, WW, OT><<<grid, CTA_THREADS, smem, st>>>(Xd, nX, Yd, nY, L, w_eff, p2, (OT*)Od, per)
    if (spec) {
        if (exact && out_dtype == WDX_F64) WDX_LAUNCH_DM(true, 25, 15, double);
        else if (exact) WDX_LAUNCH_DM(true, 25, 15, float);
        else if (out_dtype == WDX_F64) WDX_LAUNCH_DM(false, 25, 15, double);
        else WDX_LAUNCH_DM(false, 25, 15, float);
    } else {
        if (exact && out_dtype == WDX_F64) WDX_LAUNCH_DM(true, 0, 0, double);
        else if (exact) WDX_LAUNCH_DM(true, 0, 0, float);
        else if (out_dtype == WDX_F64) WDX_LAUNCH_DM(false, 0, 0, double);
        else WDX_LAUNCH_DM(false, 0, 0, float);
    }
#undef WDX_LAUNCH_DM
    if (cudaGetLastError() != cudaSuccess) return cleanup(fail(WDX_ERR_CUDA, "dtw_matrix_kernel launch failed"));
    g_launches++;
    if (!od && cudaMemcpyAsync(out, Od, (size_t)nX * nY * osz, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return cleanup(fail(WDX_ERR_CUDA, "D2H copy failed"));
    if (!xd || !yd || !od || !stream) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return cleanup(fail(WDX_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e)));
    }
    return cleanup(WDX_OK);
}

}  // extern "C"
