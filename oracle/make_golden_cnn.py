#!/usr/bin/env python
"""Golden fixture for the boundary CNN from REAL reads.

Runs in the build container only (needs /root/reference).  Imports the reference's own
`adapted.detect.cnn` unmodified and runs `prepare_data`, `cnn_score` and `cnn_detect`
(cnn.py:71-183) on the first N reads of test_data/demux/4000_rna004.pod5, batched exactly as
`file_proc.yield_signals_from_pod5` builds a minibatch (float32 rows of sig_preload_size
samples, NaN padded; file_proc.py:241-262).  Also exports the CNN weights (a torch state dict
in the reference tree) as a plain npz for the product and the oracle.

Outputs:
  tests/golden/models/cnn_rna004_130bps_v0.2.4.npz   w0,b0 .. w3,b3 (float32, torch layouts)
  tests/golden/cnn_detect_rna004.npz                 int16 ADC rows + calibration, med/mad-normalised
                                                     inputs (hash + a few rows), scores of the first rows,
                                                     boundaries of every row
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
N_READS = int(os.environ.get("WDX_GOLDEN_CNN_READS", "64"))
N_SCORE_ROWS = 12

_orig_dataclass = dataclasses.dataclass


def _dataclass(cls=None, **kw):
    kw.setdefault("unsafe_hash", True)  # Python 3.12 vs the reference's 3.10 dataclass defaults
    if cls is None:
        return lambda c: _orig_dataclass(c, **kw)
    return _orig_dataclass(cls, **kw)


def main():
    import pandas  # noqa: F401
    import scipy.signal  # noqa: F401
    import toml  # noqa: F401
    import torch
    import attrs  # noqa: F401

    dataclasses.dataclass = _dataclass
    for p in (ROOT, os.path.join(ROOT, "oracle", "shim"), REF, os.path.join(REF, "warpdemux", "adapted")):
        sys.path.insert(0, p)
    from adapted.detect.cnn import cnn_detect, cnn_score, load_cnn_model, prepare_data
    from warpdemux.config.utils import get_model_spc_config

    dataclasses.dataclass = _orig_dataclass
    from warpdemux_b200.io.pod5_min import Pod5File

    torch.set_num_threads(8)
    spc = get_model_spc_config("WDX4_rna004_v1_0")
    m = int(spc.sig_preload_size)
    cnn = load_cnn_model(spc.cnn_boundaries.model_name)
    sd = cnn.state_dict()
    wpath = os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz")
    np.savez_compressed(wpath, w0=sd["0.weight"].numpy(), b0=sd["0.bias"].numpy(), w1=sd["2.weight"].numpy(),
                        b1=sd["2.bias"].numpy(), w2=sd["4.weight"].numpy(), b2=sd["4.bias"].numpy(),
                        w3=sd["6.weight"].numpy(), b3=sd["6.bias"].numpy())

    pf = Pod5File(os.path.join(REF, "test_data", "demux", "4000_rna004.pod5"))
    reads = []
    for r in pf.reads():
        reads.append(r)
        if len(reads) >= N_READS:
            break
    n = len(reads)
    signals = np.full((n, m), np.nan, dtype=np.float32)
    adc_rows = []
    for i, r in enumerate(reads):
        _m = min(m, r.num_samples)
        signals[i, :_m] = r.signal_pa[:_m]
        adc_rows.append(np.asarray(r.signal[:_m], dtype=np.int16))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x = prepare_data(signals, spc.core)
        with torch.no_grad():
            scores = cnn_score(x, cnn).numpy()
        preds = cnn_detect(signals, cnn, spc.cnn_boundaries, spc.core)
    x = x.numpy()[:, 0, :].astype(np.float32)
    offs = np.concatenate([[0], np.cumsum([len(a) for a in adc_rows])]).astype(np.int64)
    out = os.path.join(GOLD, "cnn_detect_rna004.npz")
    np.savez_compressed(
        out,
        adc=np.concatenate(adc_rows), adc_offsets=offs, preload_size=np.int64(m),
        calibration_offset=np.array([r.calibration_offset for r in reads], dtype=np.float32),
        calibration_scale=np.array([r.calibration_scale for r in reads], dtype=np.float32),
        x_sha256=np.array(hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()),
        x_rows=x[:N_SCORE_ROWS], scores=scores[:N_SCORE_ROWS].astype(np.float32), preds=preds.astype(np.int64),
        cfg=np.array(json.dumps(dict(min_obs_adapter=int(spc.core.min_obs_adapter),
                                     max_obs_adapter=int(spc.core.max_obs_adapter),
                                     downscale_factor=int(spc.core.downscale_factor),
                                     polya_cand_k=int(spc.cnn_boundaries.polya_cand_k)))),
    )
    man_path = os.path.join(GOLD, "MANIFEST.json")
    man = json.load(open(man_path))
    for path, src in ((out, "test_data/demux/4000_rna004.pod5, first %d reads" % n),
                      (wpath, "warpdemux/adapted/adapted/models/rna004_130bps@v0.2.4.pth")):
        man["files"][os.path.relpath(path, GOLD)] = {"sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(),
                                                     "bytes": os.path.getsize(path),
                                                     "generator": "oracle/make_golden_cnn.py", "source": src}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print(out, os.path.getsize(out), "bytes;", wpath, os.path.getsize(wpath), "bytes")
    print("preds[:4]", preds[:4].tolist())


if __name__ == "__main__":
    main()
