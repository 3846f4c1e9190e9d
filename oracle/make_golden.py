#!/usr/bin/env python
"""Generate tests/golden/ fixtures by running the REFERENCE'S OWN Python.

Runs only in the build container (needs /root/reference).  What is executed:
  * `warpdemux.models.dtw_svm.DTW_SVM.predict`, `warpdemux.parallel_distances
    .distance_matrix_to`, `warpdemux.sig_proc.detect_results_to_fpt` imported
    UNMODIFIED from /root/reference,
  * the real sklearn/libsvm binary on the real shipped SVC pickles,
  * the reference's Cython `_c_segmentation.pyx` compiled by pyximport,
  * scipy.signal.find_peaks,
with one substitution: `dtaidistance` (absent) is the shim in
oracle/shim backed by oracle/wdx_oracle.c.

Outputs (small, committed):
  tests/golden/models/<name>.npz       model parameters as plain arrays
  tests/golden/predict_<name>.npz      X, y_pred, y_prob, D[:n_d]  from DTW_SVM.predict
  tests/golden/fingerprint_rna004.npz  signals, boundaries, fpt, dwell, status from detect_results_to_fpt
  tests/golden/MANIFEST.json           versions + sha256 of every fixture
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np
import pandas  # noqa: F401  (before oracle/shim is on sys.path: pandas probes for the real `bottleneck`)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "warpdemux", "adapted"))

MODELS = ["WDX4_rna004_v1_0", "WDX4b_rna004_v1_0", "WDX4c_rna004_v1_0", "WDX6_rna004_v1_0", "WDX10_rna004_v1_0",   # every shipped DTW_SVM
          "WDX12_rna002_v0_4_4"]   # + the largest deprecated rna002 model (DEPRECATED/model_files: gamma 1.2, 13 classes, 3617 SVs)
# WDX_GOLDEN_MODELS="a,b": (re)generate only those model / predict fixtures and leave everything else as it is
ONLY = [m for m in os.environ.get("WDX_GOLDEN_MODELS", "").split(",") if m]


def synth_fingerprints(sv: np.ndarray, n: int, seed: int = 0, sigma: float = 0.35) -> np.ndarray:
    """S1 of SURVEY.md §8(d): support vectors + gaussian noise."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, sv.shape[0], size=n)
    return sv[idx] + sigma * rng.standard_normal((n, sv.shape[1]))


def synth_adapter_signals(n: int, seed: int = 2, width: int = 9000):
    """S4 of SURVEY.md §8(d): piecewise-constant adapter-like signals (pA,
    float32), NaN-padded rows like the reference's minibatches
    (file_proc.py:333-354), plus adapter boundaries."""
    rng = np.random.default_rng(seed)
    sig = np.full((n, width), np.nan, dtype=np.float32)
    a0 = np.zeros(n, dtype=np.int64)
    a1 = np.zeros(n, dtype=np.int64)
    for r in range(n):
        n_levels = 121 + int(rng.integers(0, 60))
        levels = rng.standard_normal(n_levels) * 12.0 + 80.0
        dwell = 9 + rng.geometric(1.0 / 25.0, size=n_levels)
        x = np.repeat(levels, dwell)
        x = x + rng.normal(0.0, 2.0, size=x.size)
        spikes = rng.random(x.size) < 0.01
        x[spikes] += rng.choice([-60.0, 60.0], size=int(spikes.sum()))
        lead = int(rng.integers(0, 300))
        tail = int(rng.integers(200, 1500))
        full = np.concatenate([rng.normal(110.0, 3.0, lead), x, rng.normal(95.0, 6.0, tail)])[:width]
        sig[r, : full.size] = full.astype(np.float32)
        a0[r] = lead
        a1[r] = min(lead + x.size, full.size)
    # a few edge rows: tiny adapter (segmentation fails), adapter at the very start
    a1[0] = a0[0] + 40
    a0[1] = 0
    return sig, a0, a1


def main():
    import joblib
    import scipy
    import sklearn

    from warpdemux_b200 import model_io
    from warpdemux.models.dtw_svm import DTW_SVM  # noqa: F401  (reference class, needed by joblib.load)
    from warpdemux.parallel_distances import distance_matrix_to as ref_distance_matrix_to

    os.makedirs(os.path.join(GOLD, "models"), exist_ok=True)
    manifest = {
        "generator": "oracle/make_golden.py",
        "reference": "KleistLab/WarpDemuX v1.0.0 @ /root/reference (python modules imported unmodified)",
        "dtaidistance": "ABSENT - shim backed by oracle/wdx_oracle.c (restatement of 2.3.13 dtw_distance)",
        "numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__,
        "files": {},
    }

    if ONLY:
        with open(os.path.join(GOLD, "MANIFEST.json")) as fh:
            manifest = json.load(fh)
    for name in (ONLY or MODELS):
        path = os.path.join(REF, "warpdemux", "models", "model_files", name + ".joblib")
        if not os.path.exists(path):
            path = os.path.join(REF, "DEPRECATED", "model_files", name + ".joblib")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_model = joblib.load(path)  # the real reference class + real sklearn SVC
            params = model_io.from_reference_model(ref_model, name=name)
            model_io.save_npz(params, os.path.join(GOLD, "models", name + ".npz"))
            n = 192 if params.n_sv > 2000 else 256
            X = synth_fingerprints(params.sv, n, seed=0)
            # edge rows: exactly an SV (d=0), all zeros, large magnitude
            X[0] = params.sv[0]
            X[1] = 0.0
            X[2] = params.sv[-1] * 50.0
            y_pred, y_prob = ref_model.predict(X, nproc=1, return_df=False)
            df = ref_model.predict(X, nproc=1, return_df=True)
            n_d = 6
            D = ref_distance_matrix_to(X[:n_d], ref_model._X, window=ref_model.window,
                                       penalty=ref_model.penalty, n_jobs=1)
        np.savez_compressed(
            os.path.join(GOLD, f"predict_{name}.npz"),
            X=X, y_pred=y_pred.astype(np.int64), y_prob=y_prob, D=D,
            df_columns=np.array(list(df.columns)), df_values=df.to_numpy(dtype=np.float64),
        )
        print(name, "labels", dict(zip(*np.unique(y_pred, return_counts=True))))
        if ONLY:
            for rel in (os.path.join("models", name + ".npz"), f"predict_{name}.npz"):
                p = os.path.join(GOLD, rel)
                manifest["files"][rel] = {"sha256": hashlib.sha256(open(p, "rb").read()).hexdigest(), "bytes": os.path.getsize(p)}
    if ONLY:
        with open(os.path.join(GOLD, "MANIFEST.json"), "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)
        return

    # ---- fingerprints through the reference's sig_proc ------------------
    # The reference's config dataclasses use mutable defaults, which Python
    # 3.12 rejects at import (the reference pins 3.10).  sig_proc.py only uses
    # them as type names, so give it name-only stand-ins and build the config
    # object it READS from the two TOML files the reference layers
    # (config/utils.py:42-55): ADAPTed chemistry TOML overlaid by WarpDemuX's.
    import types

    import toml

    stub = types.ModuleType("warpdemux.config.sig_proc")
    stub.SegmentationConfig = type("SegmentationConfig", (), {})
    stub.SigProcConfig = type("SigProcConfig", (), {})
    sys.modules["warpdemux.config.sig_proc"] = stub
    from adapted.container_types import DetectResults
    from warpdemux.sig_proc import detect_results_to_fpt

    adapted_cfg = toml.load(os.path.join(
        REF, "warpdemux/adapted/adapted/config/config_files/rna004_130bps@v0.2.4.toml"))
    wdx_cfg = toml.load(os.path.join(REF, "warpdemux/config/config_files/rna004_130bps@v1.0.toml"))
    for sec, vals in wdx_cfg.items():
        adapted_cfg.setdefault(sec, {}).update(vals)
    adapted_cfg["segmentation"].setdefault("consensus_refinement", False)  # config/sig_proc.py default
    spc = types.SimpleNamespace(**{k: types.SimpleNamespace(**v) for k, v in adapted_cfg.items()})
    sig, a0, a1 = synth_adapter_signals(32, seed=2)
    n = sig.shape[0]
    fpt = np.full((n, 25), np.nan)
    dwell = np.zeros((n, 25), dtype=np.int64)
    status = np.zeros(n, dtype=np.int32)
    stats = np.full((n, 6), np.nan)
    work = sig.copy()  # the reference clips the minibatch row in place (sig_proc.py:426-431)
    for r in range(n):
        dr = DetectResults(success=True, adapter_start=int(a0[r]), adapter_end=int(a1[r]))
        valid = work[r][~np.isnan(work[r])]
        res = detect_results_to_fpt(valid, spc, dr)
        if res.success:
            fpt[r] = res.barcode_fpt
            dwell[r] = res.dwell_times
            stats[r] = [res.adapter_dt_med, res.adapter_dt_mad, res.adapter_event_mean,
                        res.adapter_event_std, res.adapter_event_med, res.adapter_event_mad]
        else:
            status[r] = 1
    cfg = dict(
        padding=spc.sig_extract.padding, outlier_thresh=spc.core.sig_norm_outlier_thresh,
        min_obs_per_base=spc.segmentation.min_obs_per_base,
        running_stat_width=spc.segmentation.running_stat_width,
        num_events=spc.segmentation.num_events, barcode_num_events=spc.segmentation.barcode_num_events,
    )
    np.savez_compressed(
        os.path.join(GOLD, "fingerprint_rna004.npz"),
        signals=sig, adapter_start=a0, adapter_end=a1, fpt=fpt, dwell=dwell, status=status, stats=stats,
        cfg=np.array(json.dumps(cfg)),
    )
    print("fingerprints: ok", int((status == 0).sum()), "failed", int((status != 0).sum()), cfg)

    for dirpath, _, files in os.walk(GOLD):
        for f in sorted(files):
            if f.endswith(".npz"):
                p = os.path.join(dirpath, f)
                manifest["files"][os.path.relpath(p, GOLD)] = {
                    "sha256": hashlib.sha256(open(p, "rb").read()).hexdigest(),
                    "bytes": os.path.getsize(p),
                }
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps({k: v["bytes"] for k, v in manifest["files"].items()}, indent=1))


if __name__ == "__main__":
    main()
