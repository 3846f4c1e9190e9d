#!/usr/bin/env python
"""Golden fixture for the consensus-guided (tRNA) fingerprint, SURVEY.md §8f rank 3 /
BASELINE.json configs[3].

Runs in the build container only (needs /root/reference).  Executes the reference's own
`warpdemux.sig_proc.detect_results_to_fpt` (imported unmodified) with the configuration of
WDX4_tRNA_rna004_v1_0 (`rna004_130bps@v1.0_tRNA.toml`: consensus_refinement = true) and the
reference's own consensus query (`warpdemux/_consensus.py`), as `barcode_fpt_wrapper` does
(file_proc.py:187-223), on
  (1) the REAL adapter signals of tests/golden/real_rna004_WDX4.npz (reads of
      test_data/demux/4000_rna004.pod5 carry the same RNA004 adapter the consensus models), and
  (2) a few synthetic consensus-shaped signals incl. edge cases (tests/wdx_testutil.py).
dtaidistance (absent) is the test-only shim in oracle/shim backed by oracle/wdx_oracle.c, so the
sub-sequence alignment itself is PARITY UNPINNED; everything around it (both segmentations,
normalisations, the choice of change points, the outlier filter) is the reference's own code.
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

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


def run(detect_results_to_fpt, DetectResults, spc, consensus, rows, a0, a1, ok):
    n = rows.shape[0]
    nb = int(spc.segmentation.barcode_num_events[1])
    fpt = np.full((n, nb), np.nan)
    dwell = np.zeros((n, nb), dtype=np.int64)
    stats = np.full((n, 6), np.nan)
    cons = np.zeros((n, 3), dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    reasons = {}
    work = rows.copy()
    for i in range(n):
        d = DetectResults(success=bool(ok[i]), adapter_start=int(a0[i]), adapter_end=int(a1[i]))
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = detect_results_to_fpt(work[i], spc, d, consensus)
        except (ValueError, IndexError) as e:
            # ValueError: NaN padding inside the slice, normalize() raises (sig_proc.py:288-290) -> status 3;
            # IndexError: adapter window narrower than running_stat_width, compute_base_means is handed change
            # points beyond the signal end (sig_proc.py:368, bounds-checked Cython) -> status 1
            status[i] = 3 if isinstance(e, ValueError) else 1
            reasons["raised: " + str(e)] = reasons.get("raised: " + str(e), 0) + 1
            continue
        if res.success:
            fpt[i], dwell[i] = res.barcode_fpt, res.dwell_times
        else:
            fr = str(res.fail_reason)
            status[i] = 2 if not ok[i] else (5 if "consensus" in fr else (3 if "normalization" in fr else 1))
            reasons[fr] = reasons.get(fr, 0) + 1
        if res.adapter_dt_med is not None:
            stats[i] = [res.adapter_dt_med, res.adapter_dt_mad, res.adapter_event_mean, res.adapter_event_std,
                        res.adapter_event_med, res.adapter_event_mad]
            cons[i] = [res.seg_cons_query_start, res.seg_cons_query_end, res.sig_barcode_start]
    return dict(status=status, fpt=fpt, dwell=dwell, stats=stats, cons=cons), reasons


def main():
    import pandas  # noqa: F401
    import scipy.signal  # noqa: F401
    import toml  # noqa: F401
    import attrs  # noqa: F401

    dataclasses.dataclass = _dataclass
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "shim"), REF,
              os.path.join(REF, "warpdemux", "adapted")):
        sys.path.insert(0, p)
    from adapted.container_types import DetectResults
    from warpdemux._consensus import ALL as CONSENSUS_ALL
    from warpdemux.config.utils import get_model_spc_config
    from warpdemux.sig_proc import detect_results_to_fpt

    dataclasses.dataclass = _orig_dataclass
    from wdx_testutil import real_fixture_rows, synth_trna_signals

    spc = get_model_spc_config("WDX4_tRNA_rna004_v1_0")
    seg = spc.segmentation
    assert seg.consensus_refinement and not seg.refinement_optimal_cpts
    consensus = CONSENSUS_ALL[seg.consensus_model]
    cfg = dict(padding=int(spc.sig_extract.padding), outlier_thresh=float(spc.core.sig_norm_outlier_thresh),
               min_obs_per_base=int(seg.min_obs_per_base), running_stat_width=int(seg.running_stat_width),
               num_events=int(seg.num_events), barcode_num_events=[int(v) for v in seg.barcode_num_events],
               penalty=float(seg.consensus_subseq_match_penalty), psi=[int(v) for v in seg.consensus_subseq_match_psi],
               ub_start=int(seg.consensus_subseq_match_ub_start), lb_end=int(seg.consensus_subseq_match_lb_end),
               ub_end=int(seg.consensus_subseq_match_ub_end), consensus_model=str(seg.consensus_model))
    print(cfg)

    g = np.load(os.path.join(GOLD, "real_rna004_WDX4.npz"))
    rows = real_fixture_rows(g)
    real, reasons = run(detect_results_to_fpt, DetectResults, spc, consensus, rows, g["adapter_start"],
                        g["adapter_end"], g["detect_ok"].astype(bool))
    print("real reads: ok", int((real["status"] == 0).sum()), "of", rows.shape[0], reasons)
    ok = real["status"] == 0
    print("  query end: min/median/max", real["cons"][ok, 1].min(), np.median(real["cons"][ok, 1]), real["cons"][ok, 1].max())

    sig, a0, a1 = synth_trna_signals(consensus, 48, seed=5)
    syn, reasons = run(detect_results_to_fpt, DetectResults, spc, consensus, sig, a0, a1, np.ones(len(a0), bool))
    print("synthetic: ok", int((syn["status"] == 0).sum()), "of", sig.shape[0], reasons)

    out = os.path.join(GOLD, "fingerprint_trna.npz")
    np.savez_compressed(
        out, consensus=np.asarray(consensus, dtype=np.float64), cfg=np.array(json.dumps(cfg)),
        **{"real_" + k: v for k, v in real.items()}, **{"syn_" + k: v for k, v in syn.items()},
        syn_signals_sha256=np.array(hashlib.sha256(np.ascontiguousarray(sig).tobytes()).hexdigest()),
        syn_adapter_start=a0, syn_adapter_end=a1,
    )
    man_path = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(man_path))
    man["files"]["fingerprint_trna.npz"] = {
        "sha256": hashlib.sha256(open(out, "rb").read()).hexdigest(), "bytes": os.path.getsize(out),
        "generator": "oracle/make_golden_trna.py",
        "source": "real_rna004_WDX4.npz adapter signals + wdx_testutil.synth_trna_signals(consensus, 48, seed=5)",
        "parity": "dtaidistance sub-sequence alignment is the oracle's restatement (UNPINNED); the rest is reference code"}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
