#!/usr/bin/env python
"""Golden fixture from REAL reads: BASELINE.json configs[0] ("WDX4_rna004_v1_0 demux of
test_data/demux reads on CPU, reference path, label parity check").

Runs in the build container only (needs /root/reference).  For the first N reads of
/root/reference/test_data/demux/4000_rna004.pod5 it executes the reference's own code,
imported unmodified:

    adapted.detect.combined.combined_detect_cnn      adapter boundaries (CNN + validation + LLR fallback)
    warpdemux.sig_proc.detect_results_to_fpt         fingerprints            <- accelerated path starts here
    warpdemux.models.dtw_svm.DTW_SVM.predict         barcode calls           <- ... and ends here

exactly as `file_proc.worker_detect_and_predict_on_preloaded_signals` chains them
(file_proc.py:380-455): NaN-padded float32 minibatch rows of sig_preload_size samples.
Third-party pieces absent from the image are stood in for by test-only shims:
  pod5       -> warpdemux_b200/io/pod5_min.py (signal decode; pA = (adc + offset) * scale in float32)
  bottleneck -> oracle/shim/bottleneck (numpy moving mean/var; detection only)
  dtaidistance -> oracle/shim/dtaidistance (restated DTW, oracle/wdx_oracle.c)
  _c_llr.pyx -> compiled from the reference sources into oracle/_ref (oracle/build_ref.py)
Python 3.12 rejects the reference's dataclass defaults (it pins 3.10); the dataclass
decorator is wrapped to add `unsafe_hash=True`, which lifts that check and nothing else.

Detection is UPSTREAM of the accelerated path: its boundaries are stored as inputs.  The
fixture keeps, per read, the int16 ADC samples of the adapter slice (+- padding) so that the
float32 pA row can be rebuilt bit-for-bit, the boundaries, and the reference's outputs.
"""
import dataclasses
import glob
import hashlib
import importlib.util
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
N_READS = int(os.environ.get("WDX_GOLDEN_REAL_READS", "600"))

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


def main():
    import joblib
    import pandas  # noqa: F401  (third-party modules first: only the reference's dataclasses get the wrapper)
    import scipy.signal  # noqa: F401
    import sklearn.svm  # noqa: F401
    import toml  # noqa: F401
    import torch
    import attrs  # noqa: F401

    dataclasses.dataclass = _dataclass
    for p in (ROOT, os.path.join(ROOT, "oracle", "shim"), REF, os.path.join(REF, "warpdemux", "adapted")):
        sys.path.insert(0, p)
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_c_llr*.so"))
    if not hits:
        raise SystemExit("run oracle/build_ref.py first")
    spec = importlib.util.spec_from_file_location("ref_c_llr", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["adapted.detect._c_llr"] = mod

    from adapted.detect.cnn import load_cnn_model
    from adapted.detect.combined import combined_detect_cnn
    from warpdemux.config.utils import get_model_spc_config
    from warpdemux.sig_proc import detect_results_to_fpt

    dataclasses.dataclass = _orig_dataclass
    from warpdemux_b200.io.pod5_min import Pod5File

    torch.set_num_threads(8)
    model_name = "WDX4_rna004_v1_0"
    spc = get_model_spc_config(model_name)
    spc.update_sig_preload_size() if hasattr(spc, "update_sig_preload_size") else None
    m = int(spc.sig_preload_size)
    print("sig_preload_size", m, "primary_method", spc.primary_method)
    ref_model = joblib.load(os.path.join(REF, "warpdemux", "models", "model_files", model_name + ".joblib"))
    cnn = load_cnn_model(spc.cnn_boundaries.model_name)

    pf = Pod5File(os.path.join(REF, "test_data", "demux", "4000_rna004.pod5"))
    reads = []
    for r in pf.reads():
        reads.append(r)
        if len(reads) >= N_READS:
            break
    n = len(reads)
    signals = np.full((n, m), np.nan, dtype=np.float32)                  # file_proc.py:241-262
    full_lengths = np.empty(n, dtype=np.int32)
    adc_rows = []
    for i, r in enumerate(reads):
        adc = r.signal
        _m = min(m, r.num_samples)
        full_lengths[i] = r.num_samples
        signals[i, :_m] = r.signal_pa[:_m]
        adc_rows.append(adc[:_m])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        detect_results = combined_detect_cnn(batch_of_signals=signals, full_signal_lens=full_lengths, model=cnn, spc=spc)
    ok = np.array([bool(d.success) for d in detect_results])
    print("detection: success", int(ok.sum()), "of", n)

    pad = int(spc.sig_extract.padding)
    nb = int(spc.segmentation.barcode_num_events)
    fpt = np.full((n, nb), np.nan)
    dwell = np.zeros((n, nb), dtype=np.int64)
    stats = np.full((n, 6), np.nan)
    status = np.zeros(n, dtype=np.int32)
    a0 = np.zeros(n, dtype=np.int64)
    a1 = np.zeros(n, dtype=np.int64)
    reasons = {}
    work = signals.copy()          # the reference winsorises the minibatch row in place
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, d in enumerate(detect_results):
            if d.success:
                a0[i], a1[i] = int(d.adapter_start), int(d.adapter_end)
            res = detect_results_to_fpt(work[i], spc, d)         # file_proc.py:418-428 passes the padded row
            if res.success:
                fpt[i], dwell[i] = res.barcode_fpt, res.dwell_times
                stats[i] = [res.adapter_dt_med, res.adapter_dt_mad, res.adapter_event_mean, res.adapter_event_std,
                            res.adapter_event_med, res.adapter_event_mad]
            else:
                status[i] = 2 if not d.success else (3 if "normalization" in str(res.fail_reason) else 1)
                reasons[str(res.fail_reason)] = reasons.get(str(res.fail_reason), 0) + 1
    print("fingerprints: ok", int((status == 0).sum()), "fail reasons", reasons)
    good = status == 0
    y_pred, y_prob = ref_model.predict(fpt[good], nproc=1, return_df=False)     # file_proc.py:443-450
    print("labels", dict(zip(*np.unique(y_pred, return_counts=True))))

    # compact storage: ADC samples of [slice_start, slice_stop) only
    starts = np.where(ok, np.maximum(0, a0 - pad), 0).astype(np.int64)
    stops = np.where(ok, np.minimum(m, a1 + pad), 0).astype(np.int64)
    in_len = np.array([len(a) for a in adc_rows], dtype=np.int64)        # samples present in the row (rest NaN)
    chunks, offs = [], [0]
    for i in range(n):
        hi = min(int(stops[i]), int(in_len[i]))
        c = adc_rows[i][int(starts[i]):hi] if hi > starts[i] else np.zeros(0, np.int16)
        chunks.append(c.astype(np.int16))
        offs.append(offs[-1] + c.size)
    out = os.path.join(GOLD, "real_rna004_WDX4.npz")
    np.savez_compressed(
        out,
        read_ids=np.array([r.read_id for r in reads]),
        adc=np.concatenate(chunks), adc_offsets=np.array(offs, dtype=np.int64),
        slice_start=starts, row_samples=in_len, preload_size=np.int64(m),
        calibration_offset=np.array([r.calibration_offset for r in reads], dtype=np.float32),
        calibration_scale=np.array([r.calibration_scale for r in reads], dtype=np.float32),
        detect_ok=ok.astype(np.uint8), adapter_start=a0, adapter_end=a1,
        status=status, fpt=fpt, dwell=dwell, stats=stats,
        y_pred=y_pred.astype(np.int64), y_prob=y_prob,
        cfg=np.array(json.dumps(dict(padding=pad, outlier_thresh=float(spc.core.sig_norm_outlier_thresh),
                                     min_obs_per_base=int(spc.segmentation.min_obs_per_base),
                                     running_stat_width=int(spc.segmentation.running_stat_width),
                                     num_events=int(spc.segmentation.num_events), barcode_num_events=nb,
                                     model=model_name))),
    )
    man_path = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(man_path))
    man["files"]["real_rna004_WDX4.npz"] = {"sha256": hashlib.sha256(open(out, "rb").read()).hexdigest(),
                                            "bytes": os.path.getsize(out),
                                            "generator": "oracle/make_golden_real.py",
                                            "source": "test_data/demux/4000_rna004.pod5, first %d reads" % n}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
