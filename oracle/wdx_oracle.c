/*
 * wdx_oracle.c — CPU restatement of the WarpDemuX classification hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in warpdemux_b200/ (the product) may
 * import, link or execute this file.  It is used by tests/, by
 * __graft_entry__.smoke() as the checker, and by bench.py's cpu_baseline /
 * `--impl reference` leg as the timed CPU arm.
 *
 * PARITY PINNING.  The reference ships no tests and no golden vectors
 * (SURVEY.md §4).  The DTW arithmetic lives in a third-party dependency that
 * is absent from /root/reference and from this image: dtaidistance 2.3.13
 * (environment.yml:13), C function `dtw_distance` in
 * dtaidistance/lib/DTAIDistanceC/DTAIDistanceC/dd_dtw.c.  This file restates
 * its published algorithm (squared-Euclidean inner distance, Sakoe-Chiba
 * window, squared penalty on the two non-diagonal moves, final sqrt) and is
 * pinned by
 *   (1) the KKT margin known-answer test on the shipped SVC models
 *       (tests/test_oracle_kkt.py; free support vectors satisfy
 *       |y f(x) - 1| <= 1e-3 only with this exact window/penalty semantics),
 *   (2) the reference's own Python (`warpdemux/models/dtw_svm.py`,
 *       `warpdemux/parallel_distances.py`, imported unmodified from
 *       /root/reference in the build container through a dtaidistance shim
 *       backed by this file) + the real sklearn/libsvm binary, whose outputs
 *       are committed as tests/golden/ fixtures (oracle/make_golden.py),
 *   (3) a live comparison of the libsvm restatement below against
 *       sklearn.svm.SVC.predict_proba on a freshly fitted model
 *       (tests/test_oracle_svc.py).
 * Bit-level agreement with upstream dtaidistance C is NOT verifiable here
 * (no source, no wheel): "bit-exact FP64" in this project means bit-exact
 * against this frozen restatement.
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off (no FMA contraction: the
 * reference wheels are x86-64 baseline builds).  Single-threaded per call,
 * like the reference (parallel=False, parallel_distances.py:37,61); callers
 * that want all cores run minibatches on a thread pool (ctypes drops the GIL),
 * which is how production parallelises (file_proc.py:1197-1245).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------
 * DTW distance.  Follows dtaidistance 2.3.13 dtw_distance() as called from
 * warpdemux/parallel_distances.py:59-66 (window, penalty; psi=0, no
 * max_dist/max_step/max_length_diff, no pruning).  SURVEY.md App. A.1.
 *
 *   band: j in [max(0, i - max(0,l1-l2) - window + 1),
 *               min(l2, i + max(0,l2-l1) + window))
 *   D[i][j] = (s1[i]-s2[j])^2 + min(D[i-1][j-1],
 *                                   D[i-1][j] + penalty^2,
 *                                   D[i][j-1] + penalty^2)
 *   result  = sqrt(D[l1-1][l2-1])
 * window == 0 means "no window"; penalty may be 0.
 * Two rolling rows of l2+1 cells, +inf outside the band.
 * ---------------------------------------------------------------------- */
double wdx_oracle_dtw_distance(const double *s1, int l1, const double *s2,
                               int l2, int window, double penalty)
{
    if (l1 <= 0 || l2 <= 0) return 0.0;
    if (window <= 0) window = (l1 > l2) ? l1 : l2;
    const double p2 = penalty * penalty;
    const int dl1 = (l1 > l2) ? (l1 - l2) : 0; /* max(0, r-c) */
    const int dl2 = (l2 > l1) ? (l2 - l1) : 0; /* max(0, c-r) */
    const int len = l2 + 1;
    double stackbuf[2 * 130];
    double *rows = (2 * len <= 260) ? stackbuf
                                    : (double *)malloc(sizeof(double) * 2 * len);
    double *prev = rows, *cur = rows + len;
    for (int j = 0; j < len; j++) prev[j] = INFINITY;
    prev[0] = 0.0; /* D[-1][-1] */
    for (int i = 0; i < l1; i++) {
        int jlo = i - dl1 - window + 1;
        if (jlo < 0) jlo = 0;
        int jhi = i + dl2 + window;
        if (jhi > l2) jhi = l2;
        for (int j = 0; j < len; j++) cur[j] = INFINITY;
        const double a = s1[i];
        for (int j = jlo; j < jhi; j++) {
            const double diff = a - s2[j];
            const double d = diff * diff;
            double m = prev[j];          /* D[i-1][j-1] (diagonal) */
            double t = prev[j + 1] + p2; /* D[i-1][j]   (vertical) */
            if (t < m) m = t;
            t = cur[j] + p2;             /* D[i][j-1]   (horizontal) */
            if (t < m) m = t;
            cur[j + 1] = d + m;
        }
        double *tmp = prev; prev = cur; cur = tmp;
    }
    const double r = sqrt(prev[l2]);
    if (rows != stackbuf) free(rows);
    return r;
}

/* Distance matrix between the rows of X [nX,L] and Y [nY,L] — the block
 * ((0,nX),(nX,nX+nY)) of dtw.distance_matrix(vstack([X,Y])) that
 * parallel_distances.py:58-67 slices out.  float64 out [nX,nY]. */
void wdx_oracle_dtw_matrix(const double *X, int64_t nX, const double *Y,
                           int64_t nY, int L, int window, double penalty,
                           double *out)
{
    for (int64_t r = 0; r < nX; r++)
        for (int64_t c = 0; c < nY; c++)
            out[r * nY + c] = wdx_oracle_dtw_distance(X + r * L, L, Y + c * L, L,
                                                      window, penalty);
}

/* Same, cast to float32 on store (parallel_distances.py:67 `.astype(np.float32)`). */
void wdx_oracle_dtw_matrix_f32(const double *X, int64_t nX, const double *Y,
                               int64_t nY, int L, int window, double penalty,
                               float *out)
{
    for (int64_t r = 0; r < nX; r++)
        for (int64_t c = 0; c < nY; c++)
            out[r * nY + c] = (float)wdx_oracle_dtw_distance(
                X + r * L, L, Y + c * L, L, window, penalty);
}

/* ------------------------------------------------------------------------
 * libsvm probability prediction with a precomputed kernel row.
 * Follows sklearn/svm/src/libsvm/svm.cpp: predict_values (:2868-2896),
 * sigmoid_predict (:2035-2043), multiclass_probability (:2046-2104),
 * predict_probability (:2921-2964).  SURVEY.md App. A.3.
 *
 *   Krow      [n_sv]            kernel values of one read vs every SV (float64)
 *   n_sv_cls  [k]               SVs per class (model._n_support)
 *   dual_coef [(k-1), n_sv]     model._dual_coef_ (row-major)
 *   rho       [k(k-1)/2]        = -model._intercept_
 *   probA/B   [k(k-1)/2]
 *   dec_out   [k(k-1)/2] or NULL
 *   prob_out  [k]
 * ---------------------------------------------------------------------- */
#define WDX_MAXK 32

static double sigmoid_predict(double dec, double A, double B)
{
    double fApB = dec * A + B;
    if (fApB >= 0) return exp(-fApB) / (1.0 + exp(-fApB));
    return 1.0 / (1 + exp(fApB));
}

static void multiclass_probability(int k, double r[WDX_MAXK][WDX_MAXK], double *p)
{
    int t, j, iter = 0, max_iter = (k > 100) ? k : 100;
    static __thread double Q[WDX_MAXK][WDX_MAXK];
    double Qp[WDX_MAXK];
    double pQp, eps = 0.005 / k;
    for (t = 0; t < k; t++) {
        p[t] = 1.0 / k;
        Q[t][t] = 0;
        for (j = 0; j < t; j++) {
            Q[t][t] += r[j][t] * r[j][t];
            Q[t][j] = Q[j][t];
        }
        for (j = t + 1; j < k; j++) {
            Q[t][t] += r[j][t] * r[j][t];
            Q[t][j] = -r[j][t] * r[t][j];
        }
    }
    for (iter = 0; iter < max_iter; iter++) {
        pQp = 0;
        for (t = 0; t < k; t++) {
            Qp[t] = 0;
            for (j = 0; j < k; j++) Qp[t] += Q[t][j] * p[j];
            pQp += p[t] * Qp[t];
        }
        double max_error = 0;
        for (t = 0; t < k; t++) {
            double error = fabs(Qp[t] - pQp);
            if (error > max_error) max_error = error;
        }
        if (max_error < eps) break;
        for (t = 0; t < k; t++) {
            double diff = (-Qp[t] + pQp) / Q[t][t];
            p[t] += diff;
            pQp = (pQp + diff * (diff * Q[t][t] + 2 * Qp[t])) / (1 + diff) / (1 + diff);
            for (j = 0; j < k; j++) {
                Qp[j] = (Qp[j] + diff * Q[t][j]) / (1 + diff);
                p[j] /= (1 + diff);
            }
        }
    }
}

int wdx_oracle_svc_predict_proba_row(const double *Krow, int n_sv, int k,
                                     const int *n_sv_cls, const double *dual_coef,
                                     const double *rho, const double *probA,
                                     const double *probB, double *dec_out,
                                     double *prob_out)
{
    if (k < 2 || k > WDX_MAXK) return -1;
    int start[WDX_MAXK];
    start[0] = 0;
    for (int i = 1; i < k; i++) start[i] = start[i - 1] + n_sv_cls[i - 1];
    double dec[WDX_MAXK * (WDX_MAXK - 1) / 2];
    int p = 0;
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++) {
            double sum = 0;
            const int si = start[i], sj = start[j];
            const int ci = n_sv_cls[i], cj = n_sv_cls[j];
            const double *coef1 = dual_coef + (size_t)(j - 1) * n_sv;
            const double *coef2 = dual_coef + (size_t)i * n_sv;
            for (int t = 0; t < ci; t++) sum += coef1[si + t] * Krow[si + t];
            for (int t = 0; t < cj; t++) sum += coef2[sj + t] * Krow[sj + t];
            sum -= rho[p];
            dec[p] = sum;
            p++;
        }
    if (dec_out) memcpy(dec_out, dec, sizeof(double) * p);
    const double min_prob = 1e-7;
    double R[WDX_MAXK][WDX_MAXK];
    p = 0;
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++) {
            double v = sigmoid_predict(dec[p], probA[p], probB[p]);
            v = (v > min_prob) ? v : min_prob;             /* max(v, min_prob) */
            v = (v < 1 - min_prob) ? v : (1 - min_prob);   /* min(., 1-min_prob) */
            R[i][j] = v;
            R[j][i] = 1 - v;
            p++;
        }
    multiclass_probability(k, R, prob_out);
    return 0;
}

/* Batch version over a float32 kernel matrix K [n, n_sv] (the dtype the
 * reference hands to sklearn, dtw_svm.py:90-92; sklearn upcasts exactly). */
int wdx_oracle_svc_predict_proba(const float *K, int64_t n, int n_sv, int k,
                                 const int *n_sv_cls, const double *dual_coef,
                                 const double *rho, const double *probA,
                                 const double *probB, double *dec_out,
                                 double *prob_out)
{
    if (k < 2 || k > WDX_MAXK) return -1;
    const int npair = k * (k - 1) / 2;
    {
        double *row = (double *)malloc(sizeof(double) * n_sv);
        for (int64_t r = 0; r < n; r++) {
            for (int s = 0; s < n_sv; s++) row[s] = (double)K[r * n_sv + s];
            wdx_oracle_svc_predict_proba_row(row, n_sv, k, n_sv_cls, dual_coef, rho,
                                             probA, probB,
                                             dec_out ? dec_out + r * npair : NULL,
                                             prob_out + r * k);
        }
        free(row);
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * Decision.  Follows warpdemux/models/utils.py:19-22,45-61 (process_probs,
 * confidence_margin): argmax (first maximum), conf = top1 - top2 of the
 * sorted row, label = label_map[argmax], -1 if conf < thresholds[argmax].
 * ---------------------------------------------------------------------- */
void wdx_oracle_process_probs(const double *prob, int64_t n, int k,
                              const int64_t *label_map, const double *thresholds,
                              int64_t *pred, double *conf)
{
    for (int64_t r = 0; r < n; r++) {
        const double *p = prob + r * k;
        int best = 0;
        for (int c = 1; c < k; c++)
            if (p[c] > p[best]) best = c;
        double top1 = p[best], top2 = -INFINITY;
        for (int c = 0; c < k; c++)
            if (c != best && p[c] > top2) top2 = p[c];
        conf[r] = top1 - top2;
        pred[r] = label_map[best];
        if (thresholds && conf[r] < thresholds[best]) pred[r] = -1;
    }
}

/* ------------------------------------------------------------------------
 * Whole predict path in C for the timed CPU arm (bench.py cpu_baseline /
 * --impl reference): distance matrix (float64 -> float32 cast) -> kernel
 * exp(-gamma * d^pwr) in float32 (dtw_svm.py:21-22) -> predict_proba ->
 * process_probs.  One minibatch; threads over reads.  expf() here stands in
 * for numpy's SIMD float32 exp (low bits ISA-dependent).
 * ---------------------------------------------------------------------- */
int wdx_oracle_predict(const double *X, int64_t n, const double *SV, int n_sv,
                       int L, int window, double penalty, double gamma,
                       int pwr_dist, int k, const int *n_sv_cls,
                       const double *dual_coef, const double *rho,
                       const double *probA, const double *probB,
                       const int64_t *label_map, const double *thresholds,
                       int64_t *pred, double *conf, double *prob)
{
    if (k < 2 || k > WDX_MAXK) return -1;
    int rc = 0;
    {
        double *row = (double *)malloc(sizeof(double) * n_sv);
        for (int64_t r = 0; r < n; r++) {
            for (int s = 0; s < n_sv; s++) {
                float d = (float)wdx_oracle_dtw_distance(X + r * L, L, SV + (size_t)s * L,
                                                         L, window, penalty);
                float pw = d;
                for (int q = 1; q < pwr_dist; q++) pw *= d;
                if (pwr_dist == 0) pw = 1.0f;
                row[s] = (double)expf(-(float)gamma * pw);
            }
            wdx_oracle_svc_predict_proba_row(row, n_sv, k, n_sv_cls, dual_coef, rho,
                                             probA, probB, NULL, prob + r * k);
        }
        free(row);
    }
    wdx_oracle_process_probs(prob, n, k, label_map, thresholds, pred, conf);
    return rc;
}

/* ------------------------------------------------------------------------
 * Segmentation kernels.  Follow warpdemux/segmentation/_c_segmentation.pyx:
 * c_windowed_t_test (:124-161) and c_new_means (:41-53), float64, naive
 * per-position recomputation in the reference's summation order.
 * ---------------------------------------------------------------------- */
void wdx_oracle_windowed_t_test(const double *x, int64_t n, int64_t w, double *scores)
{
    const int64_t nc = n - 2 * w;
    for (int64_t pos = 0; pos < nc; pos++) {
        double m1 = 0, m2 = 0, var1 = 0, var2 = 0, pd;
        for (int64_t i = 0; i < w; i++) m1 += x[pos + i];
        m1 /= w;
        for (int64_t i = 0; i < w; i++) m2 += x[pos + w + i];
        m2 /= w;
        for (int64_t i = 0; i < w; i++) { pd = x[pos + i] - m1; var1 += pd * pd; }
        for (int64_t i = 0; i < w; i++) { pd = x[pos + w + i] - m2; var2 += pd * pd; }
        if (var1 + var2 == 0) scores[pos] = 0.0;
        else if (m1 > m2) scores[pos] = (m1 - m2) / sqrt(var1 + var2);
        else scores[pos] = (m2 - m1) / sqrt(var1 + var2);
    }
}

void wdx_oracle_new_means(const double *x, const int64_t *segs, int64_t n_segs, double *means)
{
    for (int64_t q = 0; q < n_segs; q++) {
        double s = 0;
        for (int64_t i = segs[q]; i < segs[q + 1]; i++) s += x[i];
        means[q] = s / (double)(segs[q + 1] - segs[q]);
    }
}

/* ------------------------------------------------------------------------
 * Warping-paths matrix for the consensus-guided (tRNA) fingerprint.
 * Follows dtaidistance 2.3.13 `dtw.warping_paths` / `dtw_cc.warping_paths`
 * (dd_dtw.c: dtw_warping_paths, expanded by dtw_expand_wps) as called from
 * warpdemux/sig_proc.py:298-305:
 *     warping_paths_fast(query, norm_series, penalty=..., psi=(b1, 0, b2, 0),
 *                        compact=False, psi_neg=False)
 * i.e. no window (window = max(r, c) covers the whole matrix), no max_step /
 * max_dist, squared-Euclidean inner distance, squared penalty on the two
 * non-diagonal moves, start relaxation only:
 *     P[0][0..psi_2b] = 0,  P[0..psi_1b][0] = 0,  +inf elsewhere on the border
 *     P[i+1][j+1] = (s1[i]-s2[j])^2 + min(P[i][j], P[i][j+1] + pen^2, P[i+1][j] + pen^2)
 * and the WHOLE matrix goes through sqrt at the end.  out: (r+1) x (c+1), row-major.
 * PARITY UNPINNED: dtaidistance is absent from /root/reference and this image;
 * the reference ships no vectors for this call.
 * ---------------------------------------------------------------------- */
void wdx_oracle_warping_paths(const double *s1, int r, const double *s2, int c,
                              double penalty, int psi_1b, int psi_2b, double *out)
{
    const int W = c + 1;
    const double pen = penalty * penalty;
    for (int64_t t = 0; t < (int64_t)(r + 1) * W; t++) out[t] = INFINITY;
    for (int j = 0; j <= psi_2b && j <= c; j++) out[j] = 0.0;
    for (int i = 0; i <= psi_1b && i <= r; i++) out[(int64_t)i * W] = 0.0;
    for (int i = 0; i < r; i++) {
        const double *p0 = out + (int64_t)i * W;
        double *p1 = out + (int64_t)(i + 1) * W;
        for (int j = 0; j < c; j++) {
            const double df = s1[i] - s2[j];
            const double d = df * df;
            double m = p0[j];
            double t = p0[j + 1] + pen;
            if (t < m) m = t;
            t = p1[j] + pen;
            if (t < m) m = t;
            p1[j + 1] = d + m;
        }
    }
    for (int64_t t = 0; t < (int64_t)(r + 1) * W; t++) out[t] = sqrt(out[t]);
}

/* ------------------------------------------------------------------------
 * LLR change-point trace of the ADAPTed fallback detector.  Follows
 * warpdemux/adapted/adapted/detect/_c_llr.pyx: var_c (:24-38), _gains (:66-88,
 * stride 1, no early stopping — the only variant combined.py:39-129,232-274
 * reaches) on the cumulative sums c = cumsum(x), c2 = cumsum(x*x) that
 * c_llr_trace (:214-230) builds with np.cumsum (sequential float64 adds).
 *   var(s, e)  = c2[e-1]/e - (c[e-1]/e)^2                          (s == 0)
 *              = (c2[e-1]-c2[s-1])/(e-s) - ((c[e-1]-c[s-1])/(e-s))^2
 *   gains[i]   = (e-s) log var(s,e) - ((i-s) log var(s,i) + (e-i) log var(i,e)),
 *                i in [s + offset_head, e - offset_tail); 0 elsewhere.
 * log(0) = -inf and log(<0) = NaN propagate exactly as in the reference
 * (its callers silence the RuntimeWarnings).  Pinned against the reference's
 * own compiled Cython (oracle/_ref/ref_c_llr, tests/test_oracle_llr.py).
 * ---------------------------------------------------------------------- */
void wdx_oracle_cumsums(const double *x, int64_t n, double *c, double *c2)
{
    double a = 0.0, b = 0.0;
    for (int64_t i = 0; i < n; i++) {
        a += x[i];
        b += x[i] * x[i];
        c[i] = a;
        c2[i] = b;
    }
}

static double llr_var_c(int64_t start, int64_t end, const double *c, const double *c2)
{
    if (start == end) return 0.0;
    if (start == 0) {
        const double m = c[end - 1] / (double)end;
        return c2[end - 1] / (double)end - m * m;
    }
    const double m = (c[end - 1] - c[start - 1]) / (double)(end - start);
    return (c2[end - 1] - c2[start - 1]) / (double)(end - start) - m * m;
}

void wdx_oracle_llr_gains(const double *c, const double *c2, int64_t n, int64_t start, int64_t end,
                          int64_t offset_head, int64_t offset_tail, double *gains)
{
    for (int64_t i = 0; i < n; i++) gains[i] = 0.0;
    const double var_summed = (double)(end - start) * log(llr_var_c(start, end, c, c2));
    for (int64_t i = start + offset_head; i < end - offset_tail; i++) {
        const double head = (double)(i - start) * log(llr_var_c(start, i, c, c2));
        const double tail = (double)(end - i) * log(llr_var_c(i, end, c, c2));
        gains[i] = var_summed - (head + tail);
    }
}
