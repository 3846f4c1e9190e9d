"""Boundary validation (reference combined.py:409-683).
CPU: the numpy restatement (oracle/wdx_oracle_validate.py) against the fixture produced by the reference's
own validate_boundaries (tests/golden/validate_rna004.npz: 64 real reads + 540 synthetic rows that reach
every fail_reason).  GPU: validate_kernel through the C ABI against the same fixture and the oracle.
Bar: success flags, fail reasons, boundaries and open-pore counts identical; the float statistics
bit-identical (they are float32 order statistics / numpy-ordered sums), float64 tolerance 0."""
import dataclasses
import json
import os

import numpy as np
import pytest

from conftest import GOLD
from wdx_testutil import pack_rows, validate_golden_inputs


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(GOLD, "validate_rna004.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def gold_cnn():
    with np.load(os.path.join(GOLD, "cnn_detect_rna004.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def inputs(gold, gold_cnn):
    return validate_golden_inputs(gold, gold_cnn)


def _ocfg(gold):
    from oracle import wdx_oracle_validate as ov

    d = json.loads(str(gold["cfg"]))
    return ov.ValidateConfig(**{k: tuple(v) if isinstance(v, list) else v for k, v in d.items()})


def test_oracle_matches_reference_validate_boundaries(gold, inputs):
    from oracle import wdx_oracle_validate as ov

    rows, lens, preds = inputs
    cfg = _ocfg(gold)
    assert len(set(gold["reasons"].tolist())) >= 12          # every fail_reason of the config is in the fixture
    assert gold["success"][: int(gold["n_real"])].sum() >= 40  # most real reads validate
    for i in range(len(rows)):
        r = ov.validate_one(rows[i], int(lens[i]), int(preds[i, 0]), preds[i, 1:], cfg)
        assert bool(r["success"]) == bool(gold["success"][i]), i
        assert (ov.fail_reason(r["code"], r["checks"]) or "") == str(gold["reasons"][i]), i
        if gold["success"][i]:
            assert (r["adapter_start"], r["adapter_end"], r["polya_end"]) == tuple(gold["bounds"][i]), i
        assert np.array_equal(r["vals"][2:11], gold["vals"][i, 2:11], equal_nan=True), i
        assert r["n_open_pores"] == gold["n_open_pores"][i], i


def test_config_mirror_matches_oracle_config(gold):
    """The product's ValidateConfig carries the same fields and defaults as the oracle's."""
    from warpdemux_b200.detect import combined

    want = json.loads(str(gold["cfg"]))
    got = dataclasses.asdict(combined.ValidateConfig())
    for k, v in want.items():
        g = got[k]
        assert (list(g) if isinstance(g, tuple) else g) == v, k
    assert combined.fail_reason(7, 0b10100) == "MVS polya check failed: mean var range"
    assert combined.fail_reason(2) == "adapter MAD check failed"


@pytest.mark.gpu
def test_gpu_validate_matches_reference_and_oracle(gold, inputs):
    from oracle import wdx_oracle_validate as ov
    from warpdemux_b200.detect import combined

    rows, lens, preds = inputs
    sig = pack_rows(rows)
    cfg = _ocfg(gold)
    v = combined.Validator(combined.ValidateConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg)}), device=0)
    vb = v.validate(sig, lens, preds, partitions=True)
    n = len(rows)
    # partition statistics of DetectResults against the reference's own values (start, len, mean, std, med, mad x 3)
    has = np.array([not str(r).startswith("Validate boundaries failed") for r in gold["reasons"]])
    pbad = np.flatnonzero(~np.all((vb.parts == gold["partitions"]) | (np.isnan(vb.parts) & np.isnan(gold["partitions"])), axis=1) & has)
    assert pbad.size == 0, (pbad[:5], vb.parts[pbad[:2]], gold["partitions"][pbad[:2]])
    assert np.isnan(vb.parts[~has]).all()
    bad = []
    for i in range(n):
        reason = vb.fail_reason(i) or ""
        ok = (bool(vb.success[i]) == bool(gold["success"][i]) and reason == str(gold["reasons"][i])
              and (not gold["success"][i] or tuple(vb.bounds[i]) == tuple(gold["bounds"][i]))
              and vb.n_open_pores[i] == gold["n_open_pores"][i]
              and np.array_equal(vb.vals[i, 2:11], gold["vals"][i, 2:11], equal_nan=True))
        if not ok:
            bad.append((i, reason, str(gold["reasons"][i]), vb.bounds[i].tolist(), gold["bounds"][i].tolist(),
                        vb.vals[i, 2:11].tolist(), gold["vals"][i, 2:11].tolist()))
    assert not bad, bad[:5]
    # against the oracle: also the thresholded adapter median / MAD and the boundaries of failed reads
    osuc, ocode, ochk, obounds, ovals, opores = ov.validate_batch(sig, lens, preds, cfg)
    assert np.array_equal(vb.success, osuc) and np.array_equal(vb.code, ocode) and np.array_equal(vb.checks, ochk)
    assert np.array_equal(vb.bounds, obounds) and np.array_equal(vb.n_open_pores, opores)
    assert np.array_equal(vb.vals, ovals, equal_nan=True)
    _check_open_pores(vb, ov.open_pores_batch(sig, lens, preds, cfg), "fixture")
    v.close()
    # verdict-only mode (what the chained pipeline uses): same success and boundaries, report of the first failing candidate
    v1 = combined.Validator(v.cfg, device=0, verdict_only=True)
    vb1 = v1.validate(sig, lens, preds)
    assert np.array_equal(vb1.success, vb.success) and np.array_equal(vb1.bounds, vb.bounds)
    o1 = ov.validate_batch(sig, lens, preds, cfg, verdict_only=True)
    assert np.array_equal(vb1.code, o1[1]) and np.array_equal(vb1.checks, o1[2]) and np.array_equal(vb1.vals, o1[4], equal_nan=True)
    assert (vb1.code != vb.code).sum() + (vb1.checks != vb.checks).sum() > 0      # the fixture does exercise the difference
    v1.close()


@pytest.mark.gpu
def test_gpu_validate_device_buffers_and_variants(gold, inputs):
    """Device-resident buffers (the chained CNN -> validation -> fingerprint use), a config with the median-shift
    check on and explicit pA_mean_range, empty batch, oversize rows."""
    import torch

    from oracle import wdx_oracle_validate as ov
    from warpdemux_b200 import _lib
    from warpdemux_b200.detect import combined

    rows, lens, preds = inputs
    sel = list(range(0, len(rows), 3))
    sig = pack_rows([rows[i] for i in sel])
    lens, preds = lens[sel], preds[sel]
    ocfg = dataclasses.replace(_ocfg(gold), detect_med_shift=True, med_shift_window=700, med_shift_range=(8.0, ov.INF),
                               pA_mean_range=(95.0, 160.0), polyA_local_range=(2.0, 14.0), polyA_med_range=(90.0, 150.0),
                               mean_start_range=(60.0, 100.0), max_obs_local_range=1500, mean_window=250)
    v = combined.Validator(combined.ValidateConfig(**{f.name: getattr(ocfg, f.name) for f in dataclasses.fields(ocfg)}), device=0)
    n = sig.shape[0]
    d_sig = torch.from_numpy(sig).cuda()
    d_len = torch.from_numpy(lens.astype(np.int32)).cuda()
    d_preds = torch.from_numpy(preds).cuda()
    d_suc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_info = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_bounds = torch.zeros((n, 3), dtype=torch.int64, device="cuda")
    d_vals = torch.zeros((n, combined.N_VALS), dtype=torch.float64, device="cuda")
    v.run_raw(d_sig, n, sig.shape[1], d_len, d_preds, preds.shape[1], d_suc, d_info, d_bounds, d_vals,
              stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    osuc, ocode, ochk, obounds, ovals, opores = ov.validate_batch(sig, lens, preds, ocfg)
    assert len(set(ocode.tolist())) >= 7
    assert np.array_equal(d_suc.cpu().numpy(), osuc)
    info = d_info.cpu().numpy()
    assert np.array_equal(info[:, 0], ocode) and np.array_equal(info[:, 1], ochk) and np.array_equal(info[:, 2], opores)
    assert np.array_equal(d_bounds.cpu().numpy(), obounds)
    gv = d_vals.cpu().numpy()
    badv = np.argwhere(~((gv == ovals) | (np.isnan(gv) & np.isnan(ovals))))
    assert badv.size == 0, [(int(i), int(j), float(gv[i, j]), float(ovals[i, j]), int(ocode[i]), preds[i].tolist(), int(lens[i])) for i, j in badv[:8]]
    # empty batch is a no-op; rows too long for shared memory are refused, not truncated
    v.run_raw(d_sig, 0, sig.shape[1], d_len, d_preds, preds.shape[1], d_suc, d_info, d_bounds)
    with pytest.raises(_lib.WdxError):
        v.run_raw(d_sig, 1, 100000, d_len, d_preds, preds.shape[1], d_suc, d_info, d_bounds)
    v.close()
    with pytest.raises(_lib.WdxError):
        combined.Validator(combined.ValidateConfig(mvs_detect_overwrite=True), device=0)._handle()


@pytest.mark.gpu
def test_gpu_validate_fuzz_against_oracle():
    """Randomised rows, boundaries and configurations (incl. degenerate ones: constant signals, duplicates, rows
    shorter than the windows, adapter ends beyond the signal, candidates at / before the adapter end, a single
    candidate column): kernel == oracle on every output, both report modes."""
    from oracle import wdx_oracle_validate as ov
    from warpdemux_b200.detect import combined

    rng = np.random.default_rng(123)
    stride = 9000
    n = 160
    sig = np.full((n, stride), np.nan, dtype=np.float32)
    lens = np.zeros(n, dtype=np.int64)
    k = 4
    preds = np.zeros((n, 1 + k), dtype=np.int64)
    for i in range(n):
        L = int(rng.choice([0, 1, 5, 150, 700, 1300, 2500, 6000, stride]))
        kind = i % 8
        if kind == 0:
            x = np.full(L, 80.0)                                          # constant
        elif kind == 1:
            x = rng.integers(70, 90, L).astype(np.float64)                # heavy duplicates (crowded bins, ties)
        elif kind == 2:
            x = rng.normal(80, 6, L)
            x[rng.random(L) < 0.02] = rng.choice([1e6, -1e6, 250.0])      # outliers stretch the histogram range
        else:
            a = int(L * rng.uniform(0.3, 0.7))
            x = np.concatenate([rng.normal(80, rng.uniform(3, 12), a), rng.normal(115, rng.uniform(1, 7), L - a)])
        sig[i, :L] = x.astype(np.float32)
        lens[i] = L if rng.random() < 0.8 else L + int(rng.integers(0, 3000))
        if lens[i] > L:
            sig[i, L:min(stride, lens[i])] = 81.5                         # the row really holds full_len samples
        a_end = int(rng.choice([0, 3, 600, 1000, 1200, 2000, 4000, 8000, 12000]))
        preds[i, 0] = a_end
        c = [a_end + int(rng.choice([-50, 0, 1, 2, 3, 30, 101, 102, 103, 400, 2000, 6000])) for _ in range(k)]
        c = [max(0, v) for v in c]
        if rng.random() < 0.3:
            c[int(rng.integers(0, k))] = 0
        preds[i, 1:] = c
    cfgs = [ov.ValidateConfig(),
            ov.ValidateConfig(min_obs_adapter=100, mean_window=40, max_obs_local_range=333, local_range=(0.0, 1e9),
                              adapter_mad_range=(0.0, 1e9), pA_mean_window=5, pA_var_window=17, median_shift_window=64,
                              pA_var_range=(-ov.INF, 1e9), median_shift_range=(-1e9, ov.INF), open_pore_min=90.0,
                              open_pore_min_obs_diff=3, detect_med_shift=True, med_shift_window=77, med_shift_range=(-5.0, 60.0)),
            ov.ValidateConfig(detect_open_pores=False, real_signal_check=False, min_obs_adapter=0, adapter_mad_range=(0.0, 1e9),
                              pA_mean_adapter_med_scale_range=(0.5, ov.INF), median_shift_window=200)]
    n_lists = 0
    for ci, cfg in enumerate(cfgs):
        pcfg = combined.ValidateConfig(**{f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg)})
        for verdict_only in (False, True):
            v = combined.Validator(pcfg, device=0, verdict_only=verdict_only)
            for cols in ((1 + k), 2):
                pr = np.ascontiguousarray(preds[:, :cols])
                vb = v.validate(sig, lens, pr, partitions=True)
                o = ov.validate_batch(sig, lens, pr, cfg, verdict_only=verdict_only)
                oparts = np.array([ov.validate_one(sig[i], int(lens[i]), int(pr[i, 0]), pr[i, 1:], cfg, verdict_only=verdict_only)["parts"]
                                   for i in range(n)])
                pb = np.flatnonzero(~np.all((vb.parts == oparts) | (np.isnan(vb.parts) & np.isnan(oparts)), axis=1))
                assert pb.size == 0, ("partitions", (ci, verdict_only, cols), pb[:5], vb.parts[pb[:2]], oparts[pb[:2]])
                tag = (ci, verdict_only, cols)
                assert np.array_equal(vb.success, o[0]), tag
                assert np.array_equal(vb.code, o[1]) and np.array_equal(vb.checks, o[2]), tag
                assert np.array_equal(vb.bounds, o[3]) and np.array_equal(vb.n_open_pores, o[5]), tag
                bad = np.flatnonzero(~np.all((vb.vals == o[4]) | (np.isnan(vb.vals) & np.isnan(o[4])), axis=1))
                assert bad.size == 0, (tag, bad[:5], vb.vals[bad[:2]], o[4][bad[:2]])
                if cols == 2:   # DetectResults.open_pores: every kept position, None where the step did not run
                    n_lists += _check_open_pores(vb, ov.open_pores_batch(sig, lens, pr, cfg), tag)
            v.close()
    assert n_lists > 20
    assert len(set(o[1].tolist())) >= 3


def _check_open_pores(vb, expected, tag):
    """vb.open_pores rows against the oracle's lists; returns how many reads had more than one position."""
    from warpdemux_b200.detect import combined

    multi = 0
    for i, exp in enumerate(expected):
        if int(vb.code[i]) == 9:      # NaN error: the reference raises, no DetectResults fields at all
            continue
        if exp is not None and len(exp) > combined.PORES_LD - 1:      # more than a row lists: the count is still exact
            assert int(vb.open_pores[i, 0]) == len(exp), (tag, i)
            with pytest.raises(ValueError):
                combined.open_pores_array(vb.open_pores[i])
            continue
        got = combined.open_pores_array(vb.open_pores[i])
        if exp is None:
            assert got is None, (tag, i, got)
        else:
            assert got is not None and got.tolist() == [int(x) for x in exp], (tag, i, got, exp)
            multi += len(exp) > 1
    return multi


def test_detect_results_mirror_has_the_reference_fields():
    """DetectResults of the mirror lists the reference's fields in the reference's order (container_types.py:14-77), so
    to_dict() yields the columns of its detected_boundaries table."""
    from warpdemux_b200.detect import combined

    ref = ["success", "signal_len", "preloaded", "adapter_start", "adapter_end", "adapter_len", "adapter_mean", "adapter_std",
           "adapter_med", "adapter_mad", "polya_start", "polya_end", "polya_len", "polya_mean", "polya_std", "polya_med", "polya_mad",
           "polya_candidates", "rna_preloaded_start", "rna_preloaded_len", "rna_preloaded_mean", "rna_preloaded_std",
           "rna_preloaded_med", "rna_preloaded_mad", "start_peak_idx", "start_peak_pa", "start_peak_next_max_idx",
           "start_peak_next_max_pa", "start_peak_open_pore_idx", "adapter_rna_median_shift", "llr_adapter_end", "llr_polya_end",
           "cnn_adapter_end", "cnn_polya_end", "start_peak_adapter_end", "start_peak_polya_end", "llr_trace", "mvs_adapter_end",
           "mvs_detect_mean_at_loc", "mvs_detect_var_at_loc", "mvs_detect_polya_med", "mvs_detect_polya_local_range",
           "mvs_detect_med_shift", "real_adapter_mean_start", "real_adapter_mean_end", "real_adapter_local_range", "open_pores",
           "fail_reason"]
    assert list(combined.DetectResults(success=True).to_dict().keys()) == ref
    # to_detect_results fills them from a batch (host-only logic)
    n = 2
    vb = combined.ValidationBatch(success=np.array([1, 0], np.uint8), code=np.array([0, 7], np.int32), checks=np.array([0, 0b11100], np.int32),
                                  n_open_pores=np.array([1, 0], np.int32), bounds=np.array([[644, 4247, 5672], [0, 3000, 3500]]),
                                  vals=np.full((n, combined.N_VALS), np.nan), parts=np.arange(36, dtype=np.float64).reshape(2, 18))
    vb.parts[1, 1:6] = np.nan
    preds = np.array([[4247, 5672, 0], [3000, 3500, 3600]])
    ds = combined.to_detect_results(vb, preds, [20000, 9000], 11500)
    assert ds[0].success and ds[0].adapter_start == 644 and ds[0].cnn_adapter_end == 4247 and ds[0].preloaded == 11500
    assert ds[0].adapter_len == 1 and ds[0].polya_start == 6 and ds[0].rna_preloaded_mad == 17.0 and ds[0].open_pores.tolist() == [644]
    assert not ds[1].success and ds[1].fail_reason == "MVS polya check failed: mean var" and ds[1].needs_llr_fallback
    assert ds[1].adapter_len is None and ds[1].adapter_mean is None and ds[1].preloaded == 9000 and ds[1].open_pores is None
