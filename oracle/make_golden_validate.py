#!/usr/bin/env python
"""Golden fixture for the boundary VALIDATION step (combined.py:409-683).

Runs in the build container only (needs /root/reference).  Imports the reference's own
`adapted.detect.combined.validate_boundaries` unmodified (bottleneck, absent from the image, comes from
oracle/shim) and calls it exactly as `combined_detect_cnn` does (combined.py:211-221):
    validate_boundaries(signal[:full_signal_len], Boundaries(0, pred[0], pred[1], pred[1:]), spc, full_signal_len)
on
  * the real reads of tests/golden/cnn_detect_rna004.npz (first reads of test_data/demux/4000_rna004.pod5,
    boundaries = the reference CNN's own predictions stored in that fixture), and
  * synthetic rows from oracle/wdx_oracle_validate.synthetic_case(seed) that reach every branch.
Only the OUTPUTS are stored (the inputs are rebuilt from the CNN fixture / from the seeds):
  tests/golden/validate_rna004.npz
"""
import dataclasses
import hashlib
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WDX_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
N_SYN = int(os.environ.get("WDX_GOLDEN_VALIDATE_SYN", "540"))

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)  # Python 3.12 vs the reference's 3.10 dataclass defaults
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


# DetectResults.adapter_med / adapter_mad are the PARTITION statistics (recomputed after the open-pore shift of
# adapter_start, combined.py:631-636,677); the values validate_boundaries thresholds on are not exposed, so the
# first two columns stay NaN in the fixture and the partition statistics are kept separately.
FIELDS = (None, None, "real_adapter_mean_start", "real_adapter_mean_end", "real_adapter_local_range",
          "mvs_detect_mean_at_loc", "mvs_detect_var_at_loc", "mvs_detect_polya_med", "mvs_detect_polya_local_range",
          "mvs_detect_med_shift", "adapter_rna_median_shift")


PART_FIELDS = tuple("%s_%s" % (p, f) for p in ("adapter", "polya", "rna_preloaded") for f in ("start", "len", "mean", "std", "med", "mad"))


def main():
    import pandas  # noqa: F401
    import scipy.signal  # noqa: F401
    import toml  # noqa: F401
    import torch  # noqa: F401
    import attrs  # noqa: F401

    dataclasses.dataclass = _dataclass
    for p in (ROOT, os.path.join(ROOT, "oracle", "shim"), REF, os.path.join(REF, "warpdemux", "adapted")):
        sys.path.insert(0, p)
    import glob
    import importlib.util
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_c_llr*.so"))
    if hits:
        spec = importlib.util.spec_from_file_location("ref_c_llr", hits[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["adapted.detect._c_llr"] = mod
    from adapted.container_types import Boundaries
    from adapted.detect.combined import validate_boundaries
    from warpdemux.config.utils import get_model_spc_config

    dataclasses.dataclass = _orig_dataclass
    from oracle import wdx_oracle_validate as ov

    spc = get_model_spc_config("WDX4_rna004_v1_0")
    cfg = ov.ValidateConfig.from_spc(spc)
    print(cfg)

    g = np.load(os.path.join(GOLD, "cnn_detect_rna004.npz"))
    m = int(g["preload_size"])
    offs = g["adc_offsets"]
    n_real = len(offs) - 1
    rows, lens, preds = [], [], []
    for i in range(n_real):
        adc = g["adc"][offs[i]:offs[i + 1]]
        row = np.full(m, np.nan, dtype=np.float32)
        row[:adc.size] = (adc.astype(np.float32) + g["calibration_offset"][i]) * g["calibration_scale"][i]
        rows.append(row)
        lens.append(adc.size if adc.size < m else m + 1000)   # full length unknown here when >= preload: any value > m
        preds.append(g["preds"][i])
    k = g["preds"].shape[1] - 1
    for s in range(N_SYN):
        row, fl, pr = ov.synthetic_case(s, stride=m, k=k)
        rows.append(row)
        lens.append(int(fl))
        preds.append(pr)
    n = len(rows)
    success = np.zeros(n, np.uint8)
    bounds = np.zeros((n, 3), np.int64)
    vals = np.full((n, ov.N_VALS), np.nan)
    reasons = []
    n_pores = np.zeros(n, np.int32)
    parts = np.full((n, len(PART_FIELDS)), np.nan)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            pr = preds[i]
            b = Boundaries(adapter_start=0, adapter_end=int(pr[0]), polya_end=int(pr[1]), polya_end_topk=np.array(pr[1:]))
            try:
                d = validate_boundaries(rows[i][:lens[i]], b, spc, lens[i])
            except Exception as e:  # combined.py:293-294
                reasons.append(str(e))
                bounds[i] = (0, int(pr[0]), int(pr[1]))
                continue
            success[i] = bool(d.success)
            reasons.append("" if d.fail_reason is None else str(d.fail_reason))
            bounds[i] = (0 if d.adapter_start is None else int(d.adapter_start), 0 if d.adapter_end is None else int(d.adapter_end),
                         0 if d.polya_end is None else int(d.polya_end))
            for j, f in enumerate(FIELDS):
                v = getattr(d, f) if f else None
                vals[i, j] = np.nan if v is None else float(v)
            for j, f in enumerate(PART_FIELDS):
                v = getattr(d, f)
                parts[i, j] = np.nan if v is None else float(v)
            n_pores[i] = 0 if d.open_pores is None else int(np.size(d.open_pores))
    uniq, cnt = np.unique(np.array(reasons), return_counts=True)
    print("real reads", n_real, "synthetic", N_SYN, "success", int(success.sum()))
    for u, c in zip(uniq, cnt):
        print("  %4d  %r" % (c, u))

    # the restatement must agree before the fixture is worth anything
    bad = 0
    for i in range(n):
        r = ov.validate_one(rows[i], lens[i], int(preds[i][0]), preds[i][1:], cfg)
        want_reason = reasons[i]
        got_reason = ov.fail_reason(r["code"], r["checks"]) or ""
        same = (bool(r["success"]) == bool(success[i]) and got_reason == want_reason
                and (not success[i] or tuple(bounds[i]) == (r["adapter_start"], r["adapter_end"], r["polya_end"]))
                and np.array_equal(r["vals"][2:len(FIELDS)], vals[i, 2:len(FIELDS)], equal_nan=True)
                and r["n_open_pores"] == n_pores[i])
        if not same:
            bad += 1
            if bad < 10:
                print("MISMATCH", i, "ref:", success[i], want_reason, bounds[i], vals[i], "\n   oracle:", r)
    print("oracle vs reference mismatches:", bad, "of", n)

    out = os.path.join(GOLD, "validate_rna004.npz")
    np.savez_compressed(out, n_real=np.int64(n_real), n_syn=np.int64(N_SYN), k=np.int64(k), success=success, bounds=bounds,
                        vals=vals, partitions=parts, n_open_pores=n_pores, reasons=np.array(reasons),
                        full_lens=np.array(lens, dtype=np.int64), cfg=np.array(json.dumps(dataclasses.asdict(cfg))))
    man_path = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(man_path))
    man["files"]["validate_rna004.npz"] = {"sha256": hashlib.sha256(open(out, "rb").read()).hexdigest(),
                                           "bytes": os.path.getsize(out), "generator": "oracle/make_golden_validate.py",
                                           "source": "reference validate_boundaries on cnn_detect_rna004.npz reads + %d synthetic rows" % N_SYN}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print(out, os.path.getsize(out), "bytes")
    if bad:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
