"""Device-side throughput probe of the boundary-validation kernel and of the chained raw-signal pipeline
(CNN -> validation -> fingerprint + DTW/SVC), all buffers resident on the device (not the bench)."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from warpdemux_b200 import model_io  # noqa: E402
from warpdemux_b200.detect import cnn, combined  # noqa: E402
from warpdemux_b200.models.dtw_svm import DTW_SVM  # noqa: E402
from warpdemux_b200.sig_proc import Fingerprinter, FingerprintConfig  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def synth_reads(n, stride, seed=7):
    """Adapter (piecewise-constant, ~80 pA) + poly(A) plateau (~1.4x) + RNA, float32 NaN-padded rows."""
    rng = np.random.default_rng(seed)
    sig = np.full((n, stride), np.nan, dtype=np.float32)
    lens = np.zeros(n, dtype=np.int32)
    a_end = np.zeros(n, dtype=np.int64)
    p_end = np.zeros(n, dtype=np.int64)
    for r in range(n):
        a_len, p_len, r_len = int(rng.integers(2500, 5500)), int(rng.integers(150, 1200)), int(rng.integers(3000, 9000))
        lv = np.repeat(rng.normal(80, 11, a_len // 22 + 1), 22)[:a_len]
        x = np.concatenate([lv + rng.normal(0, 1.5, a_len), rng.normal(115, 2.5, p_len),
                            np.repeat(rng.normal(88, 12, r_len // 12 + 1), 12)[:r_len] + rng.normal(0, 2, r_len)])[:stride]
        sig[r, : x.size] = x
        lens[r] = x.size
        a_end[r], p_end[r] = a_len, a_len + p_len
    return sig, lens, a_end, p_end


def main(n_base=256, reps=int(os.environ.get("VAL_REPS", "32")), stride=int(os.environ.get("VAL_STRIDE", "16000"))):
    sig, lens, a_end, p_end = synth_reads(n_base, stride)
    n = n_base * reps
    k = 5
    rng = np.random.default_rng(1)
    preds = np.zeros((n_base, 1 + k), dtype=np.int64)
    preds[:, 0] = (a_end // 10) * 10
    preds[:, 1] = (p_end // 10) * 10
    preds[:, 2] = preds[:, 1] + 10 * rng.integers(5, 60, n_base)
    d_sig = torch.from_numpy(sig).cuda().repeat(reps, 1).contiguous()
    d_len = torch.from_numpy(lens).cuda().repeat(reps).contiguous()
    d_preds = torch.from_numpy(preds).cuda().repeat(reps, 1).contiguous()
    d_suc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_info = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    d_bounds = torch.zeros((n, 3), dtype=torch.int64, device="cuda")
    d_vals = torch.zeros((n, combined.N_VALS), dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()      # a real (non-default) stream: handle 0 would mean "the handle's own stream"
    torch.cuda.synchronize()
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    v = combined.Validator(device=0)
    v.enable_timing(True)
    best = 1e30
    for it in range(5):
        v.run_raw(d_sig, n, stride, d_len, d_preds, 1 + k, d_suc, d_info, d_bounds, d_vals, stream=stream)
        torch.cuda.synchronize()
        if it:
            best = min(best, v.last_kernel_ms())
    alg = int(lens.astype(np.int64).sum()) * 4 * reps
    codes, cnt = np.unique(d_info[:, 0].cpu().numpy(), return_counts=True)
    print(json.dumps(dict(stage="validate", reads=n, stride=stride, kernel_ms=round(best, 3), reads_per_s=round(n / best * 1e3),
                          alg_GBps=round(alg / best / 1e6, 1), codes=dict(zip(codes.tolist(), cnt.tolist())))), flush=True)

    if os.environ.get("VAL_ONLY") == "1":
        return
    v.close()
    v = combined.Validator(device=0, verdict_only=True)     # what MinibatchDemuxer uses: stop at the first failing candidate
    # chained pipeline on the same rows: CNN (guarded) -> validation -> fingerprint + DTW/SVC (guarded)
    model = cnn.load_cnn_model(os.path.join(GOLD, "models", "cnn_rna004_130bps_v0.2.4.npz"), device=0)
    core, cb = cnn.CoreConfig(), cnn.CNNBoundariesConfig(polya_cand_k=k)
    fp = Fingerprinter(FingerprintConfig(), device=0)
    mdl = DTW_SVM(model_io.load_npz(os.path.join(GOLD, "models", "WDX4_rna004_v1_0.npz")), device=0, mode="guarded")
    dm = mdl._device_model()
    d_cpreds = torch.zeros((n, 1 + k), dtype=torch.int64, device="cuda")
    d_lab = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    res = None
    for it in range(4):
        ev[0].record()
        cnn.detect_raw(model, core, k, d_sig, n, stride, d_cpreds, stream=stream)
        ev[1].record()
        v.run_raw(d_sig, n, stride, d_len, d_cpreds, 1 + k, d_suc, d_info, d_bounds, None, stream=stream)
        ev[2].record()
        a0 = d_bounds[:, 0].contiguous()
        a1 = d_bounds[:, 1].contiguous()
        fp.predict_raw(dm, d_sig, n, stride, a0, a1, 2, d_lab, d_st, detect_ok=d_suc, stream=stream)
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        if it and (res is None or sum(t) < sum(res)):
            res = t
    lab, cnt = np.unique(d_lab.cpu().numpy(), return_counts=True)
    print(json.dumps(dict(stage="chain cnn->validate->fingerprint+predict (WDX4, guarded)", reads=n, cnn_ms=round(res[0], 2),
                          validate_ms=round(res[1], 2), fp_predict_ms=round(res[2], 2), reads_per_s=round(n / sum(res) * 1e3),
                          validated=int(d_suc.sum().item()), fp_ok=int((d_st == 0).sum().item()),
                          labels=dict(zip(lab.tolist(), cnt.tolist())))), flush=True)

    # the same chain through the public minibatch API from PINNED HOST memory (upload + 3 stages + download)
    import time

    from warpdemux_b200.file_proc import MinibatchDemuxer

    torch.cuda.set_stream(torch.cuda.default_stream())
    dmx = MinibatchDemuxer(mdl, model, core=core, cnn_boundaries=cb, device=0, cnn_mode=os.environ.get("CNN_MODE") or None)
    h_sig = torch.from_numpy(np.tile(sig, (reps, 1))).pin_memory()
    h_len = np.tile(lens, reps)
    best = 1e30
    for it in range(4):
        t0 = time.perf_counter()
        r = dmx.run(h_sig, h_len, return_df=False)
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    print(json.dumps(dict(stage="MinibatchDemuxer.run from pinned host rows (e2e)", reads=n, ms=round(best * 1e3, 2),
                          reads_per_s=round(n / best), h2d_GBps=round(h_sig.numel() * 4 / best / 1e9, 1),
                          validated=int(r.detect_success.sum()), fp_ok=int((r.fp_status == 0).sum()))), flush=True)
    # raw pinned-host -> device copy rate of this box (the bound of the end-to-end numbers above and below)
    d_tmp = torch.empty_like(h_sig, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rates = {}
    for label, rows in (("whole", n), ("1000_rows", min(1000, n))):
        bestc = 1e30
        for it in range(4):
            e0.record()
            d_tmp[:rows].copy_(h_sig[:rows], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            bestc = min(bestc, e0.elapsed_time(e1))
        rates[label] = round(rows * stride * 4 / bestc / 1e6, 1)
    print(json.dumps(dict(stage="pinned H2D copy rate (GB/s)", **rates)), flush=True)
    del d_tmp
    if os.environ.get("VAL_PHASES") == "1":     # host time spent in each phase of the pipelined stream
        acc = {"_upload": 0.0, "_launch": 0.0, "_finalize": 0.0}
        for name in acc:
            orig = getattr(dmx, name)

            def wrap(*a, _o=orig, _n=name, **k):
                t0 = time.perf_counter()
                r_ = _o(*a, **k)
                acc[_n] += time.perf_counter() - t0
                return r_
            setattr(dmx, name, wrap)
    mb = 1000
    mbs = [(h_sig[a:a + mb], h_len[a:a + mb]) for a in range(0, n - mb + 1, mb)] * 4
    for it in range(2):
        t0 = time.perf_counter()
        got = sum(int(x.labels.size) for x in dmx.stream(mbs, return_df=False))
        dt = time.perf_counter() - t0
    if os.environ.get("VAL_PHASES") == "1":
        print(json.dumps({"host_ms_per_minibatch_in_phase (2 passes of the stream)": {k: round(v / (2 * len(mbs)) * 1e3, 3) for k, v in acc.items()}}), flush=True)
    t0 = time.perf_counter()
    got2 = sum(int(dmx.run(a, b, return_df=False).labels.size) for a, b in mbs)
    dt2 = time.perf_counter() - t0
    print(json.dumps(dict(stage="1000-read minibatches from pinned host rows", minibatches=len(mbs), stream_reads_per_s=round(got / dt),
                          stream_ms_per_minibatch=round(dt / len(mbs) * 1e3, 3), run_reads_per_s=round(got2 / dt2),
                          run_ms_per_minibatch=round(dt2 / len(mbs) * 1e3, 3))), flush=True)


if __name__ == "__main__":
    main()
