#!/usr/bin/env python
"""Golden fixture over ALL 4000 reads of the reference's test file — BASELINE.json configs[0]
("WDX4_rna004_v1_0 demux of test_data/demux reads on CPU, reference path, label parity check").

Runs in the build container only (needs /root/reference).  Executes the reference's own code,
imported unmodified and chained as `file_proc.worker_detect_and_predict_on_preloaded_signals`
(file_proc.py:380-455) chains it, in production minibatches of 1000 rows of sig_preload_size samples:

    adapted.detect.combined.combined_detect_cnn      CNN + validation + hail-mary / LLR fallback
    warpdemux.sig_proc.detect_results_to_fpt         fingerprints
    warpdemux.models.dtw_svm.DTW_SVM.predict         barcode calls

The detection PATH of every read (CNN validated / poly(A) re-detection on the CNN adapter end
("hail mary", combined.py:232-274) / full LLR detection (combined.py:275-290) / failed) is
recorded by wrapping the module attribute `combined.validate_boundaries` with a pass-through
logger — the reference's code is not modified.

Third-party pieces absent from the image are stood in for by test-only shims (see
oracle/make_golden_real.py): pod5 -> warpdemux_b200/io/pod5_min.py, bottleneck -> numpy shim,
dtaidistance -> restated DTW, `_c_llr.pyx` compiled from the reference sources (oracle/_ref).

Outputs
  tests/golden/real4000_rna004_WDX4.npz        (committed) per-read reference results of all 4000 reads + the int16
                                               ADC rows of a SUBSET (every read off the plain CNN path + every 16th read)
  tests/golden/_local/real4000_adc_rows.npz    (git-ignored, travels to the GPU box with the snapshot) the int16 ADC
                                               rows (first sig_preload_size samples) of all 4000 reads
"""
import dataclasses
import glob
import hashlib
import importlib.util
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
MINIBATCH = 1000          # parser.py:170-176 (production minibatch size)

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


FLOAT_FIELDS = ("mvs_detect_mean_at_loc", "mvs_detect_var_at_loc", "mvs_detect_polya_med", "mvs_detect_polya_local_range",
                "mvs_detect_med_shift", "adapter_rna_median_shift", "real_adapter_mean_start", "real_adapter_mean_end",
                "real_adapter_local_range")
PART_FIELDS = tuple(f"{p}_{f}" for p in ("adapter", "polya", "rna_preloaded") for f in ("start", "len", "mean", "std", "med", "mad"))


def main():
    import joblib
    import pandas  # noqa: F401
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

    import adapted.detect.combined as comb
    from adapted.detect.cnn import cnn_detect_boundaries, load_cnn_model
    from warpdemux.config.utils import get_model_spc_config
    from warpdemux.sig_proc import detect_results_to_fpt

    dataclasses.dataclass = _orig_dataclass
    from warpdemux_b200.io.pod5_min import Pod5File

    torch.set_num_threads(8)
    model_name = "WDX4_rna004_v1_0"
    spc = get_model_spc_config(model_name)
    spc.update_sig_preload_size() if hasattr(spc, "update_sig_preload_size") else None
    m = int(spc.sig_preload_size)
    print("sig_preload_size", m, "primary_method", spc.primary_method, "max_obs_trace", spc.core.max_obs_trace)
    ref_model = joblib.load(os.path.join(REF, "warpdemux", "models", "model_files", model_name + ".joblib"))
    cnn = load_cnn_model(spc.cnn_boundaries.model_name)

    # pass-through logger around validate_boundaries: (primary_method, adapter_end, polya_end, success) per call
    calls = []
    _vb = comb.validate_boundaries

    def logged_validate(signal, boundaries, spc_, full_signal_len):
        res = _vb(signal, boundaries, spc_, full_signal_len)
        calls.append((str(spc_.primary_method), int(boundaries.adapter_end or 0), int(boundaries.polya_end or 0), bool(res.success)))
        return res

    comb.validate_boundaries = logged_validate
    _dl = comb.detect_llr_on_downscaled_signal

    def logged_detect_llr(ds_signal, spc_):
        calls.append(("detect_llr",))          # marker: the next validate call checks the full-LLR boundaries
        return _dl(ds_signal, spc_)

    comb.detect_llr_on_downscaled_signal = logged_detect_llr

    pf = Pod5File(os.path.join(REF, "test_data", "demux", "4000_rna004.pod5"))
    reads = list(pf.reads())
    n = len(reads)
    k_cand = int(spc.cnn_boundaries.polya_cand_k)
    pad = int(spc.sig_extract.padding)
    nb = int(spc.segmentation.barcode_num_events)

    full_lengths = np.array([r.num_samples for r in reads], dtype=np.int64)
    adc_rows = [np.ascontiguousarray(r.signal[:m]) for r in reads]
    cal_off = np.array([r.calibration_offset for r in reads], dtype=np.float32)
    cal_sc = np.array([r.calibration_scale for r in reads], dtype=np.float32)

    success = np.zeros(n, np.uint8)
    path = np.zeros(n, np.uint8)            # 0 cnn, 1 hail mary (validated result replaced), 2 full llr validated, see n_validate
    n_validate = np.zeros(n, np.uint8)
    hm_tried = np.zeros(n, np.uint8)
    llr_tried = np.zeros(n, np.uint8)
    bounds = np.zeros((n, 3), np.int64)
    cnn_preds = np.zeros((n, 1 + k_cand), np.int64)
    llr_bounds = np.zeros((n, 2), np.int64)      # DetectResults.llr_adapter_end / llr_polya_end (0 = None)
    hm_polya = np.zeros(n, np.int64)
    fvals = np.full((n, len(FLOAT_FIELDS)), np.nan)
    parts = np.full((n, len(PART_FIELDS)), np.nan)
    fail_reason = np.array([""] * n, dtype=object)
    fpt = np.full((n, nb), np.nan)
    dwell = np.zeros((n, nb), dtype=np.int64)
    stats = np.full((n, 6), np.nan)
    status = np.zeros(n, dtype=np.int32)
    fp_reason = {}

    t0 = time.time()
    for lo in range(0, n, MINIBATCH):
        hi = min(n, lo + MINIBATCH)
        signals = np.full((hi - lo, m), np.nan, dtype=np.float32)                  # file_proc.py:241-262
        for j, i in enumerate(range(lo, hi)):
            a = adc_rows[i]
            signals[j, : a.size] = (a.astype(np.float32) + cal_off[i]) * cal_sc[i]
        fl = full_lengths[lo:hi].astype(np.int32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            bl = cnn_detect_boundaries(signals, cnn, spc.cnn_boundaries, spc.core)
            for j, b in enumerate(bl):
                cnn_preds[lo + j, 0] = int(b.adapter_end or 0)
                tk = np.asarray(b.polya_end_topk if b.polya_end_topk is not None else [], dtype=np.int64)
                cnn_preds[lo + j, 1:1 + min(k_cand, tk.size)] = tk[:k_cand]
            # one read at a time so that the validate calls can be attributed (combined_detect_cnn is per-read after the CNN;
            # cnn_predict's flattened-batch peak search is reproduced by feeding the SAME boundaries: checked below)
            calls.clear()
            dets = comb.combined_detect_cnn(batch_of_signals=signals, full_signal_lens=fl, model=cnn, spc=spc)
        # attribute the logged calls: every read starts with one 'cnn' call
        starts = [c for c, x in enumerate(calls) if x[0] == "cnn"]
        assert len(starts) == hi - lo, (len(starts), hi - lo)
        starts.append(len(calls))
        work = signals            # the reference winsorises the minibatch row in place (detect_results_to_fpt)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for j, d in enumerate(dets):
                i = lo + j
                seq = calls[starts[j]:starts[j + 1]]
                assert seq[0][1] == cnn_preds[i, 0] and seq[0][2] == cnn_preds[i, 1], (i, seq, cnn_preds[i])
                n_validate[i] = sum(1 for x in seq if x[0] != "detect_llr")
                rest = seq[1:]
                # a validate call that is not preceded by the detect_llr marker is the hail mary's (combined.py:267-272)
                if rest and rest[0][0] != "detect_llr":
                    assert rest[0][1] == seq[0][1] and rest[0][2] > 0
                    hm_tried[i] = 1
                    hm_polya[i] = rest[0][2]
                    if rest[0][3]:
                        path[i] = 1
                    rest = rest[1:]
                if rest:
                    assert len(rest) == 2 and rest[0][0] == "detect_llr" and rest[1][0] == "llr"
                    llr_tried[i] = 1
                    if rest[1][3]:
                        path[i] = 2
                success[i] = bool(d.success)
                bounds[i] = (int(d.adapter_start or 0), int(d.adapter_end or 0), int(d.polya_end or 0))
                llr_bounds[i] = (int(d.llr_adapter_end or 0), int(d.llr_polya_end or 0))
                fail_reason[i] = "" if d.fail_reason is None else str(d.fail_reason)
                for q, f in enumerate(FLOAT_FIELDS):
                    v = getattr(d, f)
                    fvals[i, q] = np.nan if v is None else float(v)
                for q, f in enumerate(PART_FIELDS):
                    v = getattr(d, f)
                    parts[i, q] = np.nan if v is None else float(v)
                res = detect_results_to_fpt(work[j], spc, d)         # file_proc.py:418-428 passes the padded row
                if res.success:
                    fpt[i], dwell[i] = res.barcode_fpt, res.dwell_times
                    stats[i] = [res.adapter_dt_med, res.adapter_dt_mad, res.adapter_event_mean, res.adapter_event_std,
                                res.adapter_event_med, res.adapter_event_mad]
                else:
                    status[i] = 2 if not d.success else (3 if "normalization" in str(res.fail_reason) else 1)
                    fp_reason[str(res.fail_reason)] = fp_reason.get(str(res.fail_reason), 0) + 1
        print(f"minibatch {lo}:{hi} done, {time.time() - t0:.0f}s", flush=True)
    comb.validate_boundaries = _vb
    comb.detect_llr_on_downscaled_signal = _dl

    print("detection success", int(success.sum()), "of", n)
    print("paths: cnn", int(((path == 0) & (success == 1)).sum()), "hail-mary", int((path == 1).sum()), "llr", int((path == 2).sum()),
          "failed", int((success == 0).sum()), "| hm tried", int(hm_tried.sum()), "llr tried", int(llr_tried.sum()))
    fr, fc = np.unique(fail_reason[success == 0], return_counts=True)
    print("fail reasons", dict(zip(fr.tolist(), fc.tolist())))
    print("fingerprints ok", int((status == 0).sum()), "fail reasons", fp_reason)
    good = status == 0
    y_pred, y_prob = ref_model.predict(fpt[good], nproc=1, return_df=False)     # file_proc.py:443-450
    print("labels", dict(zip(*[x.tolist() for x in np.unique(y_pred, return_counts=True)])))

    def pack(idx):
        chunks, offs = [], [0]
        for i in idx:
            chunks.append(adc_rows[i].astype(np.int16))
            offs.append(offs[-1] + adc_rows[i].size)
        return np.concatenate(chunks) if chunks else np.zeros(0, np.int16), np.array(offs, dtype=np.int64)

    # Reads whose change points depend on how numpy's (unstable) argsort orders EQUAL t-test scores: the reference's
    # result is machine-dependent there.  The fixture keeps which reads have such ties and, where a stable sort (the CUDA
    # kernels' rule: the later of two equal peaks wins) gives another fingerprint than this machine's numpy did, that
    # fingerprint too (oracle/wdx_oracle.py::fingerprint(stable_ties=True)).
    from oracle import wdx_oracle as _o

    fp_keys = dict(padding=pad, outlier_thresh=float(spc.core.sig_norm_outlier_thresh), min_obs_per_base=int(spc.segmentation.min_obs_per_base),
                   running_stat_width=int(spc.segmentation.running_stat_width), num_events=int(spc.segmentation.num_events),
                   barcode_num_events=nb)
    tie_reads, stable_idx, stable_fpt, stable_label = [], [], [], []
    for i in np.flatnonzero(status == 0):
        row = np.full(m, np.nan, dtype=np.float32)
        a = adc_rows[i]
        row[: a.size] = (a.astype(np.float32) + cal_off[i]) * cal_sc[i]
        if _o.fingerprint_has_ties(row, int(bounds[i, 0]), int(bounds[i, 1]), **fp_keys):
            tie_reads.append(i)
            st_, f_, _, _ = _o.fingerprint(row, int(bounds[i, 0]), int(bounds[i, 1]), stable_ties=True, **fp_keys)
            plain = _o.fingerprint(row, int(bounds[i, 0]), int(bounds[i, 1]), **fp_keys)
            assert np.array_equal(plain[1], fpt[i]), i          # the oracle reproduces the reference on this machine
            if st_ != 0 or not np.array_equal(f_, fpt[i]):
                stable_idx.append(i)
                stable_fpt.append(f_)
                stable_label.append(int(ref_model.predict(f_[None, :], nproc=1, return_df=False)[0][0]) if st_ == 0 else -1)
    print("reads with score ties:", len(tie_reads), "of which the stable tie rule changes the fingerprint:", stable_idx)
    subset = np.flatnonzero((n_validate > 1) | (success == 0) | (np.arange(n) % 16 == 0) | np.isin(np.arange(n), stable_idx))
    adc_sub, offs_sub = pack(subset)

    def planes(adc):
        # zig-zag deltas split into byte planes compress ~2.3x under deflate (wdx_testutil.unpack_adc_planes inverts it)
        d = np.diff(adc.astype(np.int32), prepend=np.int32(0)).astype(np.int16).astype(np.int32)
        z = ((d << 1) ^ (d >> 31)).astype(np.uint32) & 0xFFFF
        return (z & 0xFF).astype(np.uint8), (z >> 8).astype(np.uint8)

    sub_lo, sub_hi = planes(adc_sub)
    cfg = dict(padding=pad, outlier_thresh=float(spc.core.sig_norm_outlier_thresh),
               min_obs_per_base=int(spc.segmentation.min_obs_per_base), running_stat_width=int(spc.segmentation.running_stat_width),
               num_events=int(spc.segmentation.num_events), barcode_num_events=nb, model=model_name,
               max_obs_trace=int(spc.core.max_obs_trace), min_obs_adapter=int(spc.core.min_obs_adapter),
               max_obs_adapter=int(spc.core.max_obs_adapter), downscale_factor=int(spc.core.downscale_factor),
               polya_cand_k=k_cand, adapter_peak_prominence=float(spc.llr_boundaries.adapter_peak_prominence),
               adapter_peak_rel_height=float(spc.llr_boundaries.adapter_peak_rel_height),
               adapter_peak_width=int(spc.llr_boundaries.adapter_peak_width),
               fallback_to_llr=bool(spc.cnn_boundaries.fallback_to_llr),
               fallback_to_llr_short_reads=bool(spc.cnn_boundaries.fallback_to_llr_short_reads))
    out = os.path.join(GOLD, "real4000_rna004_WDX4.npz")
    np.savez_compressed(
        out,
        read_ids=np.array([r.read_id for r in reads]), full_lengths=full_lengths, preload_size=np.int64(m),
        calibration_offset=cal_off, calibration_scale=cal_sc,
        cnn_preds=cnn_preds, success=success, path=path, n_validate=n_validate, hm_tried=hm_tried, hm_polya=hm_polya,
        llr_tried=llr_tried, bounds=bounds, llr_bounds=llr_bounds, fvals=fvals, parts=parts,
        float_fields=np.array(FLOAT_FIELDS), part_fields=np.array(PART_FIELDS),
        fail_reason=np.array(fail_reason.tolist()),
        status=status, fpt=fpt, dwell=dwell, stats=stats, y_pred=y_pred.astype(np.int64), y_prob=y_prob,
        tie_reads=np.array(tie_reads, dtype=np.int64), stable_idx=np.array(stable_idx, dtype=np.int64),
        stable_fpt=np.array(stable_fpt, dtype=np.float64).reshape(-1, nb), stable_label=np.array(stable_label, dtype=np.int64),
        subset=subset.astype(np.int64), adc_lo=sub_lo, adc_hi=sub_hi, adc_offsets=offs_sub,
        cfg=np.array(json.dumps(cfg)),
    )
    os.makedirs(os.path.join(GOLD, "_local"), exist_ok=True)
    adc_all, offs_all = pack(range(n))
    big = os.path.join(GOLD, "_local", "real4000_adc_rows.npz")
    all_lo, all_hi = planes(adc_all)
    np.savez_compressed(big, adc_lo=all_lo, adc_hi=all_hi, adc_offsets=offs_all,
                        sha256=np.array(hashlib.sha256(adc_all.tobytes()).hexdigest()))
    man_path = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(man_path))
    man["files"]["real4000_rna004_WDX4.npz"] = {
        "sha256": hashlib.sha256(open(out, "rb").read()).hexdigest(), "bytes": os.path.getsize(out),
        "generator": "oracle/make_golden_real4000.py",
        "source": "test_data/demux/4000_rna004.pod5, all %d reads (ADC rows of %d of them; the rest in the git-ignored "
                  "tests/golden/_local/real4000_adc_rows.npz, sha256 of the int16 samples %s)" % (
                      n, subset.size, hashlib.sha256(adc_all.tobytes()).hexdigest())}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print(out, os.path.getsize(out), "bytes;", big, os.path.getsize(big), "bytes; subset", subset.size)


if __name__ == "__main__":
    main()
