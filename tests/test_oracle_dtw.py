"""Pins the CPU oracle's DTW: identities, an independently written Python form,
the KKT known-answer test on the shipped models, and the golden distances the
reference's own `distance_matrix_to` produced (oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import wdx_oracle as o


def test_identities(models):
    m = models["WDX4_rna004_v1_0"]
    sv = m.sv[:120]
    D = o.dtw_matrix(sv, sv, m.window, m.penalty)
    assert np.all(np.diag(D) == 0.0)
    assert np.array_equal(D, D.T)  # bit-wise symmetric
    assert np.all(D >= 0)


@pytest.mark.parametrize("window,penalty", [(15, 0.1), (3, 0.0), (0, 0.5), (25, 0.1), (1, 0.1)])
def test_c_matches_python_form(window, penalty):
    rng = np.random.default_rng(5)
    for L1, L2 in [(25, 25), (25, 25), (7, 7), (12, 9), (9, 12), (1, 1)]:
        a, b = rng.standard_normal(L1), rng.standard_normal(L2)
        assert o.dtw_distance(a, b, window, penalty) == o.dtw_distance_py(a, b, window, penalty)


def test_band_cell_count(models):
    assert models["WDX10_rna004_v1_0"].band_cells() == 515


def _kkt_worst(m, window=None, penalty=None):
    window = m.window if window is None else window
    penalty = m.penalty if penalty is None else penalty
    from concurrent.futures import ThreadPoolExecutor

    cuts = np.linspace(0, m.n_sv, 9).astype(int)       # ctypes drops the GIL: row blocks on a few threads
    with ThreadPoolExecutor(8) as ex:
        D = np.vstack(list(ex.map(lambda ab: o.distance_matrix_to(m.sv[ab[0]:ab[1]], m.sv, window, penalty), zip(cuts[:-1], cuts[1:]))))
    K = o.pdist_kernel(D, m.gamma, m.pwr_dist)
    _, dec = o.svc_predict_proba(K, m)
    start = np.concatenate([[0], np.cumsum(m.n_sv_class)])
    worst, p = 0.0, 0
    for i in range(m.k):
        for j in range(i + 1, m.k):
            for c, row, y in ((i, j - 1, 1.0), (j, i, -1.0)):
                idx = np.arange(start[c], start[c + 1])
                a = np.abs(m.dual_coef[row, idx])
                free = (a > 0) & (a < a.max() * (1 - 1e-9))
                if free.any():
                    worst = max(worst, float(np.abs(y * dec[idx, p] - 1)[free].max()))
            p += 1
    return worst


def test_kkt_known_answer(models):
    """Free support vectors of the shipped SVC sit on the margin (|y f(x) - 1|
    <= libsvm tol 1e-3) only if DTW window/penalty semantics, the float32
    kernel and the dual_coef/intercept layout are all restated correctly
    (SURVEY.md §4 item 1)."""
    for name, mm in models.items():      # every shipped DTW_SVM model (WDX4, WDX4b, WDX4c, WDX6, WDX10)
        assert _kkt_worst(mm) < 1e-3, name
    m = models["WDX4_rna004_v1_0"]
    # perturbations must break it (the test has teeth)
    assert _kkt_worst(m, window=14) > 2e-2
    assert _kkt_worst(m, window=16) > 2e-2
    assert _kkt_worst(m, penalty=np.sqrt(0.1)) > 1e-1  # i.e. penalty not squared


def test_golden_distances(models, golden_predict):
    for name, g in golden_predict.items():
        m = models[name]
        D = o.distance_matrix_to(g["X"][: g["D"].shape[0]], m.sv, m.window, m.penalty)
        assert D.dtype == np.float32
        assert np.array_equal(D, g["D"])
        assert g["D"][0, 0] == 0.0  # row 0 is SV 0 itself
