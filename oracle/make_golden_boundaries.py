#!/usr/bin/env python
"""Golden fixture of the reference's boundary / failed-read tables (second half):

    detected_boundaries_{i}.csv.gz   file_proc.py:682-724 save_detect_results("pass") -> adapted/output.py:26-51
    failed_reads_{i}.csv.gz          file_proc.py:650-665 save_batch_outputs_fail

Runs in the build container only (needs /root/reference).  The reference's own code, imported unmodified, processes ONE
minibatch — the reads of tests/golden/real4000_rna004_WDX4.npz that carry their ADC rows (every read off the plain CNN
path, every failed read, every 16th read) — exactly as `worker_detect_and_predict_on_preloaded_signals` does
(file_proc.py:380-431: combined_detect_cnn, then barcode_fpt_wrapper per read, results split into pass / fail), and its
own writers produce the two tables.  Stored: the decompressed bytes of both files, the per-read `to_summary_dict()`
records they were written from (JSON; the CPU-side writer test rebuilds them), and DetectResults.open_pores per read.

Third-party stand-ins as in oracle/make_golden_real4000.py (bottleneck -> numpy shim, dtaidistance -> restated DTW,
`_c_llr.pyx` compiled from the reference sources into oracle/_ref).

Output: tests/golden/boundaries_rna004.npz
"""
import dataclasses
import glob
import gzip
import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


def _jsonable(v):
    if v is None or isinstance(v, (bool, str)):
        return v
    if isinstance(v, np.ndarray):
        return {"__nd__": v.tolist(), "dtype": str(v.dtype)}
    if isinstance(v, (np.integer,)):
        return {"__np_int__": int(v)}
    if isinstance(v, (np.floating,)):
        return {"__np_float__": float(v), "dtype": str(v.dtype)}
    if isinstance(v, (int, float)):
        return v
    raise TypeError(type(v))


def main():
    import pandas  # noqa: F401
    import scipy.signal  # noqa: F401
    import toml  # noqa: F401
    import torch
    import attrs  # noqa: F401

    dataclasses.dataclass = _dataclass
    for p in (ROOT, os.path.join(ROOT, "oracle", "shim"), os.path.join(ROOT, "tests"), REF, os.path.join(REF, "warpdemux", "adapted")):
        sys.path.insert(0, p)
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_c_llr*.so"))
    if not hits:
        raise SystemExit("run oracle/build_ref.py first")
    spec = importlib.util.spec_from_file_location("ref_c_llr", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["adapted.detect._c_llr"] = mod

    for name in ("pod5", "pod5.reader", "catboost"):      # absent third-party modules none of the functions used here touches
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pod5.reader"].Reader = object
    sys.modules["catboost"].CatBoostClassifier = object
    import adapted.detect.combined as comb
    from adapted.detect.cnn import load_cnn_model
    from warpdemux.config.utils import get_model_spc_config
    from warpdemux.file_proc import barcode_fpt_wrapper, save_detect_results

    dataclasses.dataclass = _orig_dataclass
    from wdx_testutil import real4000_rows

    torch.set_num_threads(8)
    spc = get_model_spc_config("WDX4_rna004_v1_0")
    spc.update_sig_preload_size() if hasattr(spc, "update_sig_preload_size") else None     # as the CLI does (11 500 samples)
    m = int(spc.sig_preload_size)
    cnn = load_cnn_model(spc.cnn_boundaries.model_name)
    with np.load(os.path.join(GOLD, "real4000_rna004_WDX4.npz")) as z:
        g = {k: z[k] for k in z.files}
    idx, rows, _, _ = real4000_rows(g)           # the committed subset, float32 pA rows NaN padded to the preload size
    assert rows.shape[1] == m, (rows.shape, m)
    read_ids = [str(x) for x in g["read_ids"][idx]]
    fl = g["full_lengths"][idx].astype(np.int32)

    signals = rows.copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dets = comb.combined_detect_cnn(batch_of_signals=signals, full_signal_lens=fl, model=cnn, spc=spc)
        success, fail = [], []
        for sig, rid, d in zip(signals, read_ids, dets):      # file_proc.py:418-431
            res = barcode_fpt_wrapper(signal=sig, read_id=rid, detect_results=d, spc=spc)
            (success if res.success else fail).append(res)
    print("reads", len(read_ids), "pass", len(success), "fail", len(fail))

    with tempfile.TemporaryDirectory() as tmp:
        save_detect_results("pass", results=success, batch_idx=0, save_fpts=False, save_dwell_time=False, save_boundaries=True,
                            output_dir_boundaries=tmp, output_dir_fpts=tmp)
        save_detect_results("fail", results=fail, batch_idx=0, output_dir_fail=tmp, save_fpts=False, save_dwell_time=False,
                            save_boundaries=True)
        csv_pass = gzip.open(os.path.join(tmp, "detected_boundaries_0.csv.gz"), "rb").read()
        csv_fail = gzip.open(os.path.join(tmp, "failed_reads_0.csv.gz"), "rb").read()
    print("detected_boundaries", len(csv_pass), "bytes; failed_reads", len(csv_fail), "bytes")

    def records(results):
        out = []
        for r in results:
            d = r.to_summary_dict()
            d.pop("llr_trace", None)       # dropped by the writer (output.py:38), large
            out.append({k: _jsonable(v) for k, v in d.items()})
        return out

    by_id = {r.read_id: r for r in success + fail}
    pores_n = np.full(len(read_ids), -1, np.int32)
    pores_flat, pores_off = [], [0]
    for j, rid in enumerate(read_ids):
        dr = by_id[rid].detect_results
        op = None if dr is None else dr.open_pores
        if op is not None:
            op = np.asarray(op).ravel()
            pores_n[j] = op.size
            pores_flat.extend(int(x) for x in op)
        pores_off.append(len(pores_flat))
    print("reads with open pores:", int((pores_n > 0).sum()), "max per read", int(pores_n.max()))

    out = os.path.join(GOLD, "boundaries_rna004.npz")
    np.savez_compressed(
        out, subset=idx.astype(np.int64), csv_pass=np.frombuffer(csv_pass, dtype=np.uint8), csv_fail=np.frombuffer(csv_fail, dtype=np.uint8),
        pass_ids=np.array([r.read_id for r in success]), fail_ids=np.array([r.read_id for r in fail]),
        records_pass=np.array(json.dumps(records(success))), records_fail=np.array(json.dumps(records(fail))),
        open_pores_n=pores_n, open_pores_flat=np.array(pores_flat, dtype=np.int64), open_pores_off=np.array(pores_off, dtype=np.int64))
    man_path = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(man_path))
    man["files"]["boundaries_rna004.npz"] = {
        "sha256": hashlib.sha256(open(out, "rb").read()).hexdigest(), "bytes": os.path.getsize(out),
        "generator": "oracle/make_golden_boundaries.py",
        "source": "the %d reads of real4000_rna004_WDX4.npz that carry their ADC rows, as one minibatch through the reference's "
                  "combined_detect_cnn + barcode_fpt_wrapper + save_detect_results" % len(read_ids)}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
