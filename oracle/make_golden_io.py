#!/usr/bin/env python
"""Golden fixture for the result formats either side of the hot path.

Runs in the build container only (needs /root/reference).  Imports the reference's own
`warpdemux.file_proc` unmodified (absent third-party modules pod5 / catboost stubbed as empty
modules: none of the functions used here touches them) and, on the real-read fixture
tests/golden/real_rna004_WDX4.npz, executes

    save_fpts_signals                 file_proc.py:725-754   fingerprints/barcode_fpts_{i}.npz
    DTW_SVM.predict(return_df=True)   models/dtw_svm.py:54-98 (dtaidistance shim -> oracle/wdx_oracle.c)
    add_read_id_col_to_predictions    file_proc.py:769-780
    save_predictions                  file_proc.py:757-766   predictions/barcode_predictions_{i}.csv.gz
    scan_processed_reads              file_proc.py:129-169   (resume)
    yield_fpts_from_npz               file_proc.py:282-330   (input of `warpdemux predict`)

and stores what they wrote / returned: tests/golden/io_formats.npz.
"""
import dataclasses
import glob
import gzip
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
N_PER_FILE = (70, 50)  # two fingerprint files, so that yield_fpts_from_npz has to carry a remainder over

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


def main():
    import attrs  # noqa: F401
    import joblib
    import pandas  # noqa: F401
    import scipy.signal  # noqa: F401
    import sklearn.svm  # noqa: F401
    import toml  # noqa: F401
    import torch  # noqa: F401

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
    for name in ("pod5", "pod5.reader", "catboost"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pod5.reader"].Reader = object
    sys.modules["catboost"].CatBoostClassifier = object
    import warpdemux.file_proc as fp
    from warpdemux.sig_proc import ReadResult

    dataclasses.dataclass = _orig_dataclass

    with np.load(os.path.join(GOLD, "real_rna004_WDX4.npz")) as z:
        g = {k: z[k] for k in z.files}
    good = np.flatnonzero(g["status"] == 0)[: sum(N_PER_FILE)]
    read_ids, fpt, dwell = g["read_ids"][good], g["fpt"][good], g["dwell"][good]
    ref_model = joblib.load(os.path.join(REF, "warpdemux", "models", "model_files", "WDX4_rna004_v1_0.joblib"))

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "fingerprints"))
        os.makedirs(os.path.join(tmp, "predictions"))
        lo = 0
        for bidx, cnt in enumerate(N_PER_FILE):
            sl = slice(lo, lo + cnt)
            lo += cnt
            results = [ReadResult(read_id=str(r), success=True, barcode_fpt=f, dwell_times=d)
                       for r, f, d in zip(read_ids[sl], fpt[sl], dwell[sl])]
            fn = os.path.join(tmp, "fingerprints", f"barcode_fpts_{bidx}.npz")
            fp.save_fpts_signals(results, fn, save_dwell_time=(bidx == 0))
            with np.load(fn) as w:
                for k in w.files:
                    out[f"fpts{bidx}_{k}"] = w[k]
                out[f"fpts{bidx}_keys"] = np.array(json.dumps(list(w.files)))
            df = ref_model.predict(fpt[sl], pbar=False, nproc=1, return_df=True)        # file_proc.py:488-493
            df = fp.add_read_id_col_to_predictions(df, read_ids[sl])
            fn = os.path.join(tmp, "predictions", f"barcode_predictions_{bidx}.csv.gz")
            fp.save_predictions(df, fn)
            with gzip.open(fn, "rt") as fh:
                out[f"pred{bidx}_csv"] = np.array(fh.read())
        processed, max_pass, max_fail = fp.scan_processed_reads(tmp, scan_failed=False, result_type="predictions")
        out["scan_pred_ids"] = np.array(sorted(processed))
        out["scan_pred_bidx"] = np.array([max_pass, max_fail])
        processed, max_pass, max_fail = fp.scan_processed_reads(tmp, scan_failed=False, result_type="fingerprints")
        out["scan_fpts_ids"] = np.array(sorted(str(r) for r in processed))
        out["scan_fpts_bidx"] = np.array([max_pass, max_fail])
        files = [os.path.join(tmp, "fingerprints", f"barcode_fpts_{b}.npz") for b in range(len(N_PER_FILE))]
        # every file must hold at least one listed read: with none, the reference indexes with an empty
        # float64 array (np.array([])) and raises IndexError (file_proc.py:303-312)
        excl = set(str(r) for r in read_ids[5:15]) | set(str(r) for r in read_ids[75:80])
        incl = set(str(r) for r in read_ids[60:90])
        for tag, kw in (("all", dict(read_ids_incl=set(), read_ids_excl=set())),
                        ("excl", dict(read_ids_incl=set(), read_ids_excl=excl)),
                        ("incl", dict(read_ids_incl=incl, read_ids_excl=set()))):
            batches = list(fp.yield_fpts_from_npz(files, batch_size=32, **kw))
            out[f"yield_{tag}_sizes"] = np.array([len(b[1]) for b in batches])
            out[f"yield_{tag}_ids"] = np.concatenate([np.asarray(b[1]).astype(str) for b in batches])
            out[f"yield_{tag}_fpts"] = np.concatenate([b[0] for b in batches], axis=0)
    out["read_ids"], out["fpt"], out["dwell"] = read_ids, fpt, dwell
    out["excl"], out["incl"] = np.array(sorted(excl)), np.array(sorted(incl))
    out["n_per_file"] = np.array(N_PER_FILE)
    path = os.path.join(GOLD, "io_formats.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    print(str(out["pred0_csv"])[:400])


if __name__ == "__main__":
    main()
