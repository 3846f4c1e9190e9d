"""GPU parity of the consensus-guided (tRNA) fingerprint (wdx_fp_set_consensus + wdx_fp_extract_ex /
wdx_fp_predict through the C ABI), SURVEY.md §8f rank 3 / BASELINE.json configs[3]:
  * against the fixture produced by the reference's own `detect_results_to_fpt` with the
    WDX4_tRNA configuration on 500 REAL adapter signals and on synthetic edge cases
    (tests/golden/fingerprint_trna.npz, oracle/make_golden_trna.py), and
  * against the CPU oracle on seeded synthetic signals.
Bars: status, match positions (query start / end, barcode start) and dwell times identical;
float64 fingerprints and adapter statistics bit-identical.  The dtaidistance sub-sequence
alignment is a restatement on BOTH sides (PARITY UNPINNED for that piece, oracle/wdx_oracle.c)."""
import json
import os
import pickle

import numpy as np
import pytest

from wdx_testutil import oracle_fingerprints_consensus, real_fixture_rows, synth_trna_signals

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_trna():
    with np.load(os.path.join(ROOT, "tests", "golden", "fingerprint_trna.npz")) as z:
        return {k: z[k] for k in z.files}


def make_fp(g, **kw):
    from warpdemux_b200.sig_proc import Fingerprinter, FingerprintConfig

    cfg = json.loads(str(g["cfg"]))
    fc = FingerprintConfig(padding=cfg["padding"], outlier_thresh=cfg["outlier_thresh"],
                           min_obs_per_base=cfg["min_obs_per_base"], running_stat_width=cfg["running_stat_width"],
                           num_events=cfg["num_events"], barcode_num_events=cfg["barcode_num_events"][1],
                           consensus=tuple(g["consensus"].tolist()), barcode_segm_events=cfg["barcode_num_events"][0],
                           consensus_penalty=cfg["penalty"], consensus_psi=tuple(cfg["psi"]),
                           consensus_ub_start=cfg["ub_start"], consensus_lb_end=cfg["lb_end"],
                           consensus_ub_end=cfg["ub_end"], **kw)
    return Fingerprinter(fc, device=0), cfg


def same(b, status, fpt, dwell, stats, cons):
    assert np.array_equal(b.status, status), np.flatnonzero(b.status != status)
    ok = status == 0
    rep = ok | (status == 5)
    assert np.array_equal(b.cons[rep], cons[rep]), np.flatnonzero((b.cons != cons).any(axis=1) & rep)
    assert np.array_equal(b.dwell[ok], dwell[ok])
    assert np.array_equal(b.fpt[ok], fpt[ok]), np.abs(b.fpt[ok] - fpt[ok]).max()
    assert np.array_equal(b.stats[rep], stats[rep])
    assert np.isnan(b.fpt[~ok]).all()


def test_reference_golden_real_reads(golden_trna, golden_real):
    g = golden_trna
    fp, _ = make_fp(g)
    rows = real_fixture_rows(golden_real)
    b = fp.extract(rows, golden_real["adapter_start"], golden_real["adapter_end"], detect_ok=golden_real["detect_ok"])
    same(b, g["real_status"], g["real_fpt"], g["real_dwell"], g["real_stats"], g["real_cons"])
    assert (b.status == 0).sum() >= 400
    fp.close()


def test_three_launch_form_on_big_batches(golden_trna, golden_real):
    """Batches of >= 2048 reads take the three-launch consensus form (front part -> alignment kernel of small CTAs ->
    refinement on the parked state): the same fixture, tiled, must come out bit for bit; the single-launch form
    (WDX_FP_NO_SPLIT) must agree with it."""
    g = golden_trna
    rows = real_fixture_rows(golden_real)
    reps = -(-2304 // rows.shape[0])
    tile = lambda x: np.concatenate([x] * reps, axis=0)
    big = tile(rows)
    a0, a1, ok = tile(golden_real["adapter_start"]), tile(golden_real["adapter_end"]), tile(golden_real["detect_ok"])
    assert big.shape[0] >= 2048
    fp, _ = make_fp(g)
    b = fp.extract(big, a0, a1, detect_ok=ok)
    same(b, tile(g["real_status"]), tile(g["real_fpt"]), tile(g["real_dwell"]), tile(g["real_stats"]), tile(g["real_cons"]))
    os.environ["WDX_FP_NO_SPLIT"] = "1"
    try:
        b1 = fp.extract(big, a0, a1, detect_ok=ok)
    finally:
        del os.environ["WDX_FP_NO_SPLIT"]
    assert np.array_equal(b.status, b1.status) and np.array_equal(b.fpt, b1.fpt, equal_nan=True)
    assert np.array_equal(b.dwell, b1.dwell) and np.array_equal(b.cons, b1.cons) and np.array_equal(b.stats, b1.stats, equal_nan=True)
    fp.close()


def test_reference_golden_synthetic(golden_trna):
    g = golden_trna
    fp, _ = make_fp(g)
    sig, a0, a1 = synth_trna_signals(g["consensus"], 48, seed=5)
    b = fp.extract(sig, a0, a1)
    same(b, g["syn_status"], g["syn_fpt"], g["syn_dwell"], g["syn_stats"], g["syn_cons"])
    fp.close()


@pytest.mark.parametrize("seed", [31, 32])
def test_synthetic_signals_match_oracle(golden_trna, seed):
    g = golden_trna
    fp, cfg = make_fp(g)
    sig, a0, a1 = synth_trna_signals(g["consensus"], 200, seed=seed)
    status, fpt, dwell, stats, cons = oracle_fingerprints_consensus(sig, a0, a1, g["consensus"], **cfg)
    assert (status == 0).sum() > 80 and (status == 5).sum() > 5
    b = fp.extract(sig, a0, a1)
    same(b, status, fpt, dwell, stats, cons)
    # device-resident buffers, no statistics / dwell times requested
    import torch

    sd = torch.from_numpy(sig).cuda()
    fd = torch.empty((sig.shape[0], 25), dtype=torch.float64, device="cuda")
    st = torch.empty(sig.shape[0], dtype=torch.int32, device="cuda")
    cd = torch.empty((sig.shape[0], 3), dtype=torch.int32, device="cuda")
    fp.extract_raw(sd, sig.shape[0], sig.shape[1], torch.from_numpy(a0).cuda(), torch.from_numpy(a1).cuda(), fd, st,
                   cons=cd, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(st.cpu().numpy(), status)
    ok = status == 0
    assert np.array_equal(fd.cpu().numpy()[ok], fpt[ok])
    assert np.array_equal(cd.cpu().numpy()[ok], cons[ok])
    fp.close()


def test_other_query_lengths_and_relaxations(golden_trna):
    """Rows-per-lane 1..4 of the wavefront, different psi / penalty / event counts."""
    g = golden_trna
    rng = np.random.default_rng(9)
    sig, a0, a1 = synth_trna_signals(g["consensus"], 40, seed=33)
    for qlen, psi, pen, segm, keep in [(20, (0, 0, 0, 0), 0.0, 25, 25), (33, (3, 0, 90, 0), 0.7, 30, 20),
                                       (84, (5, 0, 40, 0), 1.5, 25, 25), (128, (10, 0, 10, 0), 2.0, 26, 25)]:
        q = np.resize(g["consensus"], qlen) + 0.01 * rng.standard_normal(qlen)
        cfg = dict(padding=100, outlier_thresh=5.0, min_obs_per_base=9, running_stat_width=18, num_events=120,
                   barcode_num_events=[segm, keep], penalty=pen, psi=list(psi), ub_start=60, lb_end=10, ub_end=119)
        status, fpt, dwell, stats, cons = oracle_fingerprints_consensus(sig, a0, a1, q, **cfg)
        from warpdemux_b200.sig_proc import Fingerprinter, FingerprintConfig

        fp = Fingerprinter(FingerprintConfig(min_obs_per_base=9, running_stat_width=18, num_events=120,
                                             barcode_num_events=keep, consensus=tuple(q.tolist()),
                                             barcode_segm_events=segm, consensus_penalty=pen, consensus_psi=psi,
                                             consensus_ub_start=60, consensus_lb_end=10, consensus_ub_end=119), device=0)
        b = fp.extract(sig, a0, a1)
        same(b, status, fpt, dwell, stats, cons)
        if qlen <= 84:   # a 128-point query ends at the end of the 121 events: (almost) no barcode events left
            assert (status == 0).sum() > 10, (qlen, np.unique(status, return_counts=True))
        fp.close()


def test_reference_shaped_calls_and_fused_predict(golden_trna, golden_real, models):
    """`detect_results_to_fpt(signal, spc, detect_results, consensus_query)` with the reference's
    signature, and the fused signals -> calls step with a DTW-SVM model as the shape proxy of the
    (CatBoost, out of scope) tRNA classifier — SURVEY.md F4."""
    import types

    from warpdemux_b200.models.dtw_svm import DTW_SVM
    from warpdemux_b200.sig_proc import DetectResults, FingerprintConfig, detect_results_to_fpt

    g = golden_trna
    cfg = json.loads(str(g["cfg"]))
    seg = types.SimpleNamespace(min_obs_per_base=cfg["min_obs_per_base"], running_stat_width=cfg["running_stat_width"],
                                num_events=cfg["num_events"], normalization="mean", accept_less_cpts=False,
                                barcode_num_events=cfg["barcode_num_events"], consensus_refinement=True,
                                consensus_subseq_match_normalization="mean",
                                consensus_subseq_match_penalty=cfg["penalty"], consensus_subseq_match_psi=cfg["psi"],
                                consensus_subseq_match_ub_start=cfg["ub_start"],
                                consensus_subseq_match_lb_end=cfg["lb_end"], consensus_subseq_match_ub_end=cfg["ub_end"],
                                refinement_optimal_cpts=False)
    spc = types.SimpleNamespace(segmentation=seg, sig_extract=types.SimpleNamespace(padding=100, normalization="none"),
                                core=types.SimpleNamespace(sig_norm_outlier_thresh=5.0))
    rows = real_fixture_rows(golden_real)
    seen = set()
    for r in range(60):
        if not golden_real["detect_ok"][r]:
            continue
        valid = rows[r][: int(golden_real["row_samples"][r])].copy()
        if np.isnan(valid[max(0, int(golden_real["adapter_start"][r]) - 100): int(golden_real["adapter_end"][r]) + 100]).any():
            continue
        dr = DetectResults(success=True, adapter_start=int(golden_real["adapter_start"][r]),
                           adapter_end=int(golden_real["adapter_end"][r]))
        res = detect_results_to_fpt(valid, spc, dr, g["consensus"])
        st = int(g["real_status"][r])
        seen.add(st)
        assert res.success == (st == 0)
        if st == 0:
            assert np.array_equal(res.barcode_fpt, g["real_fpt"][r]) and np.array_equal(res.dwell_times, g["real_dwell"][r])
        if st in (0, 5):
            assert (res.seg_cons_query_start, res.seg_cons_query_end, res.sig_barcode_start) == tuple(g["real_cons"][r])
            assert res.adapter_event_std == g["real_stats"][r, 3]
        if st == 5:
            assert res.fail_reason == "consensus query outlier"
    assert 0 in seen
    with pytest.raises(ValueError):
        FingerprintConfig.from_spc(spc)          # consensus refinement without a query
    # picklable configuration / fused step
    fp, _ = make_fp(g)
    fp2 = pickle.loads(pickle.dumps(fp))
    mdl = DTW_SVM(models["WDX4_rna004_v1_0"], device=0, mode="exact")
    labels, prob, conf, status, fpt = fp2.extract_and_predict(mdl, rows, golden_real["adapter_start"],
                                                              golden_real["adapter_end"],
                                                              detect_ok=golden_real["detect_ok"], want_fpt=True)
    assert np.array_equal(status, g["real_status"])
    ok = status == 0
    assert np.array_equal(fpt[ok], g["real_fpt"][ok])
    y_pred, y_prob = mdl.predict(g["real_fpt"][ok], nproc=1)
    assert np.array_equal(labels[ok], y_pred) and (labels[~ok] == -1).all()
    fp.close()
    fp2.close()
