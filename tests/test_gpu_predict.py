"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the
golden outputs of the reference's own DTW_SVM.predict.

Bars: EXACT_F64 distances bit-identical to the oracle; float32 casts identical;
labels / threshold decisions identical; probabilities within 2e-6 (the float32
exp of the kernel is not bit-reproducible across implementations);
FAST_F32 distances within 1e-5 relative (BASELINE.json north_star)."""
import pickle

import numpy as np
import pytest

from wdx_testutil import synth_fingerprints

pytestmark = pytest.mark.gpu

PROB_ATOL = 2e-6
FAST_RTOL = 1e-5


@pytest.fixture(scope="module")
def dev_models(models):
    from warpdemux_b200.device_model import DeviceModel

    out = {n: DeviceModel(m, 0) for n, m in models.items()}
    yield out
    for d in out.values():
        d.close()


def test_distance_matrix_exact_is_bitwise(models):
    from oracle import wdx_oracle as o
    from warpdemux_b200.device_model import distance_matrix

    m = models["WDX4_rna004_v1_0"]
    X = synth_fingerprints(m.sv, 300, seed=3)
    want = o.dtw_matrix(X, m.sv, m.window, m.penalty)
    got = distance_matrix(X, m.sv, m.window, m.penalty, mode="exact", out_dtype=np.float64)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    got32 = distance_matrix(X, m.sv, m.window, m.penalty, mode="exact", out_dtype=np.float32)
    assert np.array_equal(got32, want.astype(np.float32))
    fast = distance_matrix(X, m.sv, m.window, m.penalty, mode="fast", out_dtype=np.float64)
    rel = np.abs(fast - want) / np.maximum(want, 1e-30)
    assert rel.max() <= FAST_RTOL, rel.max()


def test_distance_matrix_to_seam(models):
    """drop-in `distance_matrix_to` (parallel_distances.py:48-84): float32 out."""
    from oracle import wdx_oracle as o
    from warpdemux_b200.parallel_distances import distance_matrix_to

    m = models["WDX6_rna004_v1_0"]
    X = synth_fingerprints(m.sv, 130, seed=4)
    got = distance_matrix_to(X, m.sv, window=m.window, penalty=m.penalty, block_size=500, n_jobs=1)
    assert got.dtype == np.float32
    assert np.array_equal(got, o.distance_matrix_to(X, m.sv, m.window, m.penalty))
    # symmetric use (SV vs SV): zero diagonal, bit-wise symmetric
    D = distance_matrix_to(m.sv[:200], m.sv[:200], window=m.window, penalty=m.penalty)
    assert np.all(np.diag(D) == 0) and np.array_equal(D, D.T)
    # the block-parallel variants of the reference (parallel_distances.py:24-45, 87-198) give the same numbers
    from warpdemux_b200.parallel_distances import compute_block_distance, parallel_distance_matrix, parallel_distance_matrix_to

    assert np.array_equal(parallel_distance_matrix_to(X, m.sv, block_size=50, n_jobs=3, window=m.window, penalty=m.penalty), got)
    Z = np.vstack([X, m.sv])
    sub = parallel_distance_matrix(Z, block_size=64, subset=((0, len(X)), (len(X), len(Z))), window=m.window, penalty=m.penalty)
    assert np.array_equal(sub, got)
    i, j, blk = compute_block_distance((np.arange(5, 25), np.arange(len(X), len(X) + 40)), Z, window=m.window, penalty=m.penalty)
    assert np.array_equal(blk, got[5:25, :40]) and blk.dtype == np.float32


@pytest.mark.parametrize("shape", [(12, 5, 0.1), (30, 0, 0.0), (25, 25, 0.3), (25, 14, 0.1), (1, 1, 0.1), (64, 20, 0.05)])
def test_distance_matrix_generic_shapes(shape):
    from oracle import wdx_oracle as o
    from warpdemux_b200.device_model import distance_matrix

    L, w, pen = shape
    rng = np.random.default_rng(L * 100 + w)
    X, Y = rng.standard_normal((70, L)), rng.standard_normal((150, L))
    want = o.dtw_matrix(X, Y, w, pen)
    got = distance_matrix(X, Y, w, pen, mode="exact", out_dtype=np.float64)
    assert np.array_equal(got, want)
    fast = distance_matrix(X, Y, w, pen, mode="fast", out_dtype=np.float64)
    # float32 rounding of the inputs bounds the absolute error (cancellation when L is tiny)
    assert np.allclose(fast, want, rtol=FAST_RTOL, atol=2e-6)


# (L, window, penalty): strips of 4 / 8 / 16 columns per lane, one and several column chunks, windows
# narrower than a strip, wider than a chunk, absent (0 = no window), and L at the chunk boundaries
WAVEFRONT_SHAPES = [(57, 0, 0.1), (60, 9, 0.2), (64, 64, 0.0), (65, 15, 0.1), (65, 0, 0.1), (100, 30, 0.0), (128, 128, 0.3), (129, 0, 0.1), (130, 7, 0.1),
                    (257, 1, 0.1), (300, 40, 0.05), (300, 0, 0.1), (530, 50, 0.1), (530, 0, 0.2), (700, 3, 0.1),
                    (1024, 0, 0.1), (1500, 200, 0.1)]


@pytest.mark.parametrize("shape", WAVEFRONT_SHAPES)
def test_distance_matrix_wavefront_long_series(shape):
    """Series longer than a register row (L > 64) take the warp-wide anti-diagonal wavefront
    (dtw_wavefront.cuh): EXACT bit-identical to the oracle, FAST within 1e-5 relative."""
    from oracle import wdx_oracle as o
    from warpdemux_b200.device_model import distance_matrix

    L, w, pen = shape
    rng = np.random.default_rng(L * 1000 + w)
    nX, nY = (23, 37) if L <= 600 else (5, 9)
    X, Y = rng.standard_normal((nX, L)), rng.standard_normal((nY, L))
    Y[0] = X[0]  # distance exactly 0
    X[1, ::7] += 25.0  # spikes: large cell costs next to small ones
    want = o.dtw_matrix(X, Y, w, pen)
    got = distance_matrix(X, Y, w, pen, mode="exact", out_dtype=np.float64)
    assert np.array_equal(got, want)
    assert got[0, 0] == 0.0
    got32 = distance_matrix(X, Y, w, pen, mode="exact", out_dtype=np.float32)
    assert np.array_equal(got32, want.astype(np.float32))
    fast = distance_matrix(X, Y, w, pen, mode="fast", out_dtype=np.float64)
    assert np.allclose(fast, want, rtol=FAST_RTOL, atol=2e-6)


def test_distance_matrix_wavefront_nonfinite_and_limits():
    from oracle import wdx_oracle as o
    from warpdemux_b200.device_model import distance_matrix

    rng = np.random.default_rng(5)
    L = 200
    X, Y = rng.standard_normal((6, L)), rng.standard_normal((4, L))
    X[2, 50] = np.nan
    X[3, 10] = np.inf
    Y[1, 199] = np.nan
    want = o.dtw_matrix(X, Y, 25, 0.1)
    got = distance_matrix(X, Y, 25, 0.1, mode="exact", out_dtype=np.float64)
    assert np.array_equal(got, want, equal_nan=True)
    # the same pair through both code paths (thread-per-pair for L <= 64, wavefront above): pad with a
    # constant tail that adds zero cost along the diagonal
    Xs, Ys = rng.standard_normal((9, 60)), rng.standard_normal((11, 60))
    Xl, Yl = np.hstack([Xs, np.zeros((9, 40))]), np.hstack([Ys, np.zeros((11, 40))])
    assert np.array_equal(distance_matrix(Xl, Yl, 0, 0.1, mode="exact", out_dtype=np.float64), o.dtw_matrix(Xl, Yl, 0, 0.1))
    with pytest.raises(ValueError):
        distance_matrix(np.zeros((1, 16385)), np.zeros((1, 16385)), 10, 0.1)


def test_golden_predict_exact(models, dev_models, golden_predict):
    """Same inputs the reference's DTW_SVM.predict was run on (oracle/make_golden.py)."""
    for name, g in golden_predict.items():
        labels, prob, conf, flags, dist = dev_models[name].predict(g["X"], mode="exact", want_dist=True)
        assert np.array_equal(labels, g["y_pred"]), name
        np.testing.assert_allclose(prob, g["y_prob"], rtol=0, atol=PROB_ATOL)
        nd = g["D"].shape[0]
        assert np.array_equal(dist[:nd], g["D"]), name  # float32 distances bit-identical
        assert not flags.any()
        np.testing.assert_allclose(prob.sum(1), 1.0, atol=1e-9)


@pytest.mark.parametrize("n,splits", [(1, 0), (37, 0), (1000, 0), (1000, 1), (513, 7), (2500, 3)])
def test_predict_matches_oracle_across_split_geometries(models, dev_models, n, splits):
    from oracle import wdx_oracle as o

    for name in ("WDX4_rna004_v1_0", "WDX10_rna004_v1_0"):
        if name.startswith("WDX10") and n > 1000:
            continue
        m, d = models[name], dev_models[name]
        X = synth_fingerprints(m.sv, n, seed=n + splits)
        want_pred, want_prob, want_conf, want_D = o.predict(m, X)
        d.set_sv_splits(splits)
        try:
            labels, prob, conf, flags, dist = d.predict(X, mode="exact", want_dist=True)
        finally:
            d.set_sv_splits(0)
        assert np.array_equal(dist, want_D)
        np.testing.assert_allclose(prob, want_prob, rtol=0, atol=PROB_ATOL)
        np.testing.assert_allclose(conf, want_conf, rtol=0, atol=2 * PROB_ATOL)
        # labels identical except where the oracle itself is within float32-exp noise of a boundary
        diff = labels != want_pred
        if diff.any():
            thr = m.thresholds[np.argmax(want_prob, 1)]
            assert np.all(np.minimum(np.abs(want_conf - thr), want_conf)[diff] < 4 * PROB_ATOL)


def test_fast_and_guarded_modes(models, dev_models):
    name = "WDX4_rna004_v1_0"
    m, d = models[name], dev_models[name]
    X = synth_fingerprints(m.sv, 30000, seed=11)
    le, pe, ce, _ = d.predict(X, mode="exact")
    lf, pf, cf, _, = d.predict(X, mode="fast")
    lg, pg, cg, fg = d.predict(X, mode="guarded")
    # fast: probabilities close, a few boundary labels may differ
    assert np.abs(pf - pe).max() < 5e-4
    assert (lf != le).mean() < 2e-3
    # guarded: decisions identical to exact; recomputed reads carry exact values
    assert np.array_equal(lg, le)
    rec = (fg & 2) != 0
    assert rec.sum() < 0.05 * len(X)
    np.testing.assert_allclose(pg[rec], pe[rec], rtol=0, atol=1e-12)
    np.testing.assert_allclose(cg[rec], ce[rec], rtol=0, atol=1e-12)
    assert np.array_equal(pg[~rec], pf[~rec])


def test_guard_band_is_wide_enough_and_overflow_is_repaired(models, dev_models):
    """The default guard (5e-5) must dominate the FAST-vs-EXACT confidence error with
    a 10x margin; and when the device-side boundary list overflows, the host wrapper
    re-runs the leftover reads so that GUARDED == EXACT still holds."""
    for name, n in (("WDX4_rna004_v1_0", 200000), ("WDX10_rna004_v1_0", 60000)):
        m, d = models[name], dev_models[name]
        X = synth_fingerprints(m.sv, n, seed=21)
        le, pe, ce, _ = d.predict(X, mode="exact")
        lf, pf, cf, _ = d.predict(X, mode="fast")
        assert np.abs(cf - ce).max() < 5e-6, np.abs(cf - ce).max()
        lg, pg, cg, fg = d.predict(X, mode="guarded")
        assert np.array_equal(lg, le)
        assert ((fg & 2) != 0).mean() < 2e-3
    m, d = models["WDX4_rna004_v1_0"], dev_models["WDX4_rna004_v1_0"]
    X = synth_fingerprints(m.sv, 10000, seed=22)
    le, pe, ce, _ = d.predict(X, mode="exact")
    d.set_guard(10.0)  # every read is "near a boundary": the 4096-entry list overflows
    try:
        lg, pg, cg, fg = d.predict(X, mode="guarded")
    finally:
        d.set_guard(5e-5)
    assert np.all((fg & 2) != 0) and not np.any(fg & 4)
    # same arithmetic, possibly another SV-split geometry (summation association): equal to ~1 ulp
    assert np.array_equal(lg, le)
    np.testing.assert_allclose(pg, pe, rtol=0, atol=1e-12)
    np.testing.assert_allclose(cg, ce, rtol=0, atol=1e-12)


def test_dropin_class_api(models, golden_predict):
    from warpdemux_b200.models.dtw_svm import DTW_SVM

    name = "WDX4_rna004_v1_0"
    g = golden_predict[name]
    mdl = DTW_SVM(models[name], mode="exact")
    y_pred, y_prob = mdl.predict(g["X"], nproc=1, return_df=False)
    assert y_pred.dtype == np.int64 and y_prob.dtype == np.float64 and y_prob.shape == (len(g["X"]), 5)
    assert np.array_equal(y_pred, g["y_pred"])
    df = mdl.predict(g["X"], pbar=False, nproc=1, return_df=True)  # the call file_proc.py:445-450 makes
    assert list(df.columns) == list(g["df_columns"])
    assert np.array_equal(df["predicted_barcode"].to_numpy(), g["y_pred"])
    np.testing.assert_allclose(df.to_numpy(dtype=np.float64)[:, 1:], g["df_values"][:, 1:], atol=1.01e-3)
    # 1-D input is one read (dtw_svm.py:70-71), the live caller's shape (worker.py:117-120)
    y1, p1 = mdl.predict(g["X"][5])
    assert y1.shape == (1,) and p1.shape == (1, 5) and y1[0] == g["y_pred"][5]
    # error behaviour (dtw_svm.py:65-77)
    with pytest.raises(ValueError, match="same number of columns"):
        mdl.predict(np.zeros((3, 24)))
    with pytest.raises(ValueError, match="Model not trained yet."):
        DTW_SVM(None).predict(np.zeros((1, 25)))
    # picklable like the reference model that is shipped to every worker (file_proc.py:1232-1243)
    clone = pickle.loads(pickle.dumps(mdl))
    y2, p2 = clone.predict(g["X"][:16], nproc=1)
    assert np.array_equal(y2, g["y_pred"][:16])
    # empty batch
    y0, p0 = mdl.predict(np.zeros((0, 25)))
    assert y0.shape == (0,) and p0.shape == (0, 5)


def test_nonfinite_fingerprints(models):
    from warpdemux_b200.models.dtw_svm import DTW_SVM

    m = models["WDX4_rna004_v1_0"]
    X = synth_fingerprints(m.sv, 40, seed=1)
    X[7, 3] = np.nan
    with pytest.raises(ValueError, match="NaN"):  # sklearn's check in predict_proba raises in the reference
        DTW_SVM(m, mode="exact").predict(X)
    y, p = DTW_SVM(m, mode="exact", on_nonfinite="noise").predict(X)
    assert y[7] == -1 and np.isnan(p[7]).all()
    ok = np.ones(40, bool)
    ok[7] = False
    y_ref, p_ref = DTW_SVM(m, mode="exact").predict(X[ok])
    assert np.array_equal(y[ok], y_ref) and np.array_equal(p[ok], p_ref)
    # +inf gives distance inf, kernel 0: finite probabilities, like the reference
    X2 = synth_fingerprints(m.sv, 8, seed=2)
    X2[3, 0] = np.inf
    y2, p2 = DTW_SVM(m, mode="exact").predict(X2)
    assert np.isfinite(p2).all()


def test_float32_input_and_device_pointers(models, dev_models):
    import torch

    from warpdemux_b200 import _lib

    name = "WDX6_rna004_v1_0"
    m, d = models[name], dev_models[name]
    X = synth_fingerprints(m.sv, 700, seed=5)
    l64, p64, c64, _ = d.predict(X, mode="exact")
    X32 = X.astype(np.float32)
    l32, p32, _, _ = d.predict(X32, mode="exact")
    l32b, p32b, _, _ = d.predict(X32.astype(np.float64), mode="exact")
    assert np.array_equal(l32, l32b) and np.array_equal(p32, p32b)  # f32 input widened exactly
    # device-resident buffers (torch only as the allocator)
    Xd = torch.from_numpy(X).cuda()
    lab = torch.empty(len(X), dtype=torch.int64, device="cuda")
    prob = torch.empty((len(X), m.k), dtype=torch.float64, device="cuda")
    conf = torch.empty(len(X), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    d.predict_raw(Xd, len(X), _lib.WDX_F64, _lib.MODE_EXACT_F64, lab, conf, prob, None, None, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(lab.cpu().numpy(), l64) and np.array_equal(prob.cpu().numpy(), p64)


def test_synthetic_model_shapes(models):
    """Shapes beyond the shipped ones: 13 classes (deprecated WDX12 shape), tiny classes, L != 25."""
    from oracle import wdx_oracle as o
    from warpdemux_b200 import model_io
    from warpdemux_b200.device_model import DeviceModel

    for n_sv_class, L, w in [([40] * 13, 25, 15), ([3, 1, 70, 2], 25, 15), ([30, 50, 20], 16, 6), ([5, 5], 25, 15),
                             ([20] * 16, 25, 15)]:
        m = model_io.synthetic_model(n_sv_class, L=L, window=w, seed=len(n_sv_class))
        X = synth_fingerprints(m.sv, 300, seed=9, sigma=0.5)
        want_pred, want_prob, want_conf, want_D = o.predict(m, X)
        d = DeviceModel(m, 0)
        for splits in (0, 1, 5):
            d.set_sv_splits(splits)
            labels, prob, conf, flags, dist = d.predict(X, mode="exact", want_dist=True)
            assert np.array_equal(dist, want_D)
            np.testing.assert_allclose(prob, want_prob, rtol=0, atol=PROB_ATOL)
            diff = labels != want_pred
            assert diff.mean() < 0.01
        d.close()


def test_fast_distances_tolerance_through_fused_kernel(models, dev_models):
    """FAST_F32 distances as the fused kernel forms them (packed recurrence in offset form) stay
    within 1e-5 relative of the float64 oracle for ordinary reads AND for reads (almost) identical
    to a support vector, where the kernel redoes the pair with the plain recurrence."""
    from oracle import wdx_oracle as o

    for name in ("WDX4_rna004_v1_0", "WDX10_rna004_v1_0"):
        m, d = models[name], dev_models[name]
        rng = np.random.default_rng(5)
        parts = [synth_fingerprints(m.sv, 200, seed=21)]
        for sigma in (0.0, 1e-4, 1e-3, 1e-2, 5e-2, 0.15):
            idx = rng.integers(0, m.n_sv, 60)
            parts.append(m.sv[idx] + sigma * rng.standard_normal((60, m.L)))
        X = np.vstack(parts)
        # inputs representable in float32, so that the comparison isolates the recurrence
        X = X.astype(np.float32).astype(np.float64)
        want = o.dtw_matrix(X, m.sv.astype(np.float32).astype(np.float64), m.window, m.penalty)
        _, _, _, _, dist = d.predict(X, mode="fast", want_dist=True)
        got = dist.astype(np.float64)
        ref32 = want.astype(np.float32).astype(np.float64)
        rel = np.abs(got - ref32) / np.maximum(ref32, 1e-30)
        rel[ref32 == 0] = np.where(got[ref32 == 0] == 0, 0.0, np.inf)
        assert rel.max() <= FAST_RTOL + 1.2e-7, (name, rel.max())   # + one float32 rounding of the stored distance
        assert (want < 0.25).sum() > 50                               # the near-zero branch was exercised


def test_small_batch_host_path_equals_chunked_path(models, dev_models):
    """Live-sized batches with host buffers take the one-copy-each-way path (<= 8192 reads); larger ones the
    pipelined chunk path.  Same kernels, so labels are identical and probabilities agree to summation geometry."""
    import torch

    from warpdemux_b200 import _lib

    name = "WDX4_rna004_v1_0"
    m, d = models[name], dev_models[name]
    X = synth_fingerprints(m.sv, 8300, seed=9)
    for mode in ("exact", "guarded"):
        la, pa, ca, _ = d.predict(X, mode=mode)                 # chunk path
        lb, pb, cb, _ = d.predict(X[:8192], mode=mode)          # small path at its upper edge
        lc, pc, cc, _ = d.predict(X[:3], mode=mode)
        assert np.array_equal(la[:8192], lb) and np.array_equal(la[:3], lc)
        assert np.abs(pa[:8192] - pb).max() < 2e-6 and np.abs(pa[:3] - pc).max() < 2e-6
        assert np.abs(ca[:8192] - cb).max() < 4e-6
    # labels only (conf / prob / flags NULL), pinned and float32 inputs
    lab = np.empty(100, dtype=np.int64)
    Xp = torch.from_numpy(X[:100].copy()).pin_memory()
    d.predict_raw(Xp, 100, _lib.WDX_F64, _lib.MODE_FAST_F32_GUARDED, lab)
    assert np.array_equal(lab, la[:100])
    lab32 = np.empty(100, dtype=np.int64)
    d.predict_raw(X[:100].astype(np.float32), 100, _lib.WDX_F32, _lib.MODE_EXACT_F64, lab32)
    want, _, _, _ = d.predict(X[:100].astype(np.float32).astype(np.float64), mode="exact")
    assert np.array_equal(lab32, want)


def test_warp_finish_is_bit_identical_to_thread_finish(models, dev_models):
    """Batches <= 16384 reads finish with one warp per read (latency), larger ones with one thread per read
    (throughput).  With the same SV-range geometry the two must agree to the last bit."""
    for name in ("WDX4_rna004_v1_0", "WDX10_rna004_v1_0"):
        m, d = models[name], dev_models[name]
        X = synth_fingerprints(m.sv, 16500, seed=11)
        X[5] = np.nan                      # non-finite fingerprint -> label -1, NaN probabilities on both paths
        X[7] = m.sv[3]                     # zero distance to a support vector
        d.set_sv_splits(8)
        try:
            for mode in ("exact", "fast"):
                lb, pb, cb, fb = d.predict(X, mode=mode)            # thread-per-read finish
                ls, ps, cs, fs = d.predict(X[:4000], mode=mode)     # warp-per-read finish
                assert np.array_equal(lb[:4000], ls) and np.array_equal(fb[:4000], fs)
                assert np.array_equal(pb[:4000], ps, equal_nan=True) and np.array_equal(cb[:4000], cs, equal_nan=True)
        finally:
            d.set_sv_splits(0)


def test_concurrent_predict_from_four_threads(models):
    """The live path classifies from 4 threads (live_balancing/session.py:165-167): calls on ONE model object are
    serialised by the handle's mutex, calls on per-thread objects run concurrently; both give the single-thread result."""
    import threading

    from warpdemux_b200.models.dtw_svm import DTW_SVM

    m = models["WDX6_rna004_v1_0"]
    X = synth_fingerprints(m.sv, 64 * 24, seed=13)
    shared = DTW_SVM(m, device=0, mode="guarded")
    want, want_p = shared.predict(X, nproc=1)
    for per_thread in (False, True):
        out, errs = {}, []

        def work(t):
            try:
                mdl = DTW_SVM(m, device=0, mode="guarded") if per_thread else shared
                for r in range(6):
                    lo = (t * 6 + r) * 64
                    y, p = mdl.predict(X[lo:lo + 64], nproc=1)
                    out[(t, r)] = (lo, y, p)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ths = [threading.Thread(target=work, args=(t,)) for t in range(4)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        assert not errs, errs
        assert len(out) == 24
        for lo, y, p in out.values():
            assert np.array_equal(y, want[lo:lo + 64])
            assert np.abs(p - want_p[lo:lo + 64]).max() < 2e-6
