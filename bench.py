#!/usr/bin/env python
"""bench.py — reads/s of the WarpDemuX classification hot path on B200.

    python bench.py --gpus N --steps K --warmup W          # this framework
    python bench.py --impl reference --steps K --warmup W  # reference CPU path (oracle port)

Workload (BASELINE.json metric / configs[2]): WDX10_rna004_v1_0 (2601 support
vectors, 11 classes, L=25, window 15) on synthetic barcode fingerprints S1
(support vector + 0.35*N(0,1)).  100 M reads across 8 GPUs =
12.5 M reads per GPU per step; reads are sharded by contiguous index range, one
process per GPU, no data-path collective (weak scaling).

A "step" = one pass of the fused path (DTW distance to every support vector ->
DTW-kernel SVC probabilities -> thresholded barcode call) over this rank's
batch.  `value` is timed with CUDA events with the batch already in HBM;
`e2e` goes through the reference-facing API `DTW_SVM.predict(X_host)` with
host buffers (H2D and D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "WDX10_rna004_v1_0"
METRIC = "reads/s demuxed (WDX10 DTW+SVC)"
READS_PER_GPU = 12_500_000  # 100 M / 8
SIGMA = 0.35


def load_params():
    from warpdemux_b200 import model_io

    return model_io.load_npz(os.path.join(ROOT, "tests", "golden", "models", MODEL + ".npz"))


def synth_host(params, n, seed):
    """S1 fingerprints, generated in 1 M-row blocks (bounded host memory spikes)."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, params.L), dtype=np.float64)
    for r0 in range(0, n, 1 << 20):
        r1 = min(n, r0 + (1 << 20))
        idx = rng.integers(0, params.n_sv, size=r1 - r0)
        X[r0:r1] = params.sv[idx] + SIGMA * rng.standard_normal((r1 - r0, params.L))
    return X


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); pw.append(float(r[2]))
            except Exception:  # noqa: BLE001
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "power_w_max": float(np.max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(params, threads, seconds_target, steps=1, warmup=0):
    """The reference's CPU path as restated in oracle/ (kind "port"): production
    style minibatches of 1000 reads over `threads` single-threaded workers."""
    from oracle import wdx_oracle as o

    o.lib()
    probe = synth_host(params, 64, seed=123)
    t0 = time.perf_counter()
    o.predict_c(params, probe)
    per_read = (time.perf_counter() - t0) / 64
    n = int(max(threads * 50, min(200_000, seconds_target / per_read * threads)))
    n = max(threads, (n // threads) * threads)
    mb = min(1000, max(1, n // threads))
    X = synth_host(params, n, seed=7)
    for _ in range(warmup):
        o.predict_threaded(params, X[: max(threads, n // 8)], threads, mb)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.predict_threaded(params, X, threads, mb)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = load_params()
    threads = os.cpu_count() or 1
    total_budget = 120.0
    per_step = max(2.0, min(20.0, total_budget / max(1, args.steps + args.warmup)))
    rps, n, dt = cpu_arm(params, threads, per_step, steps=args.steps, warmup=min(args.warmup, 1))
    cells = params.n_sv * params.band_cells()
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{MODEL} on synthetic S1 fingerprints; bounded sample of {n} reads per step",
                   "model": MODEL, "n_sv": params.n_sv, "classes": params.k, "L": params.L, "window": params.window},
        "gcups": rps * cells / 1e9,
        "cpu_baseline": {"value": rps, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": f"{n} S1 reads/step, minibatches of <=1000 over {threads} single-threaded workers "
                                   "(oracle/wdx_oracle.c: restated dtaidistance DTW + libsvm predict_proba)"},
        "e2e": {"value": rps, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from warpdemux_b200 import _lib
    from warpdemux_b200.device_model import DeviceModel
    from warpdemux_b200.models.dtw_svm import DTW_SVM
    from warpdemux_b200.sharding import shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    params = load_params()
    n_total = args.reads_per_gpu * world
    lo, hi = shard_bounds(n_total, world)[rank]  # contiguous index range of this rank
    n = hi - lo
    mode = args.mode
    k = params.k
    cells_per_read = params.n_sv * params.band_cells()

    # this rank's shard of the synthetic set (seeded per shard so any N gives a reproducible set)
    X_host_t = torch.empty((n, params.L), dtype=torch.float64).pin_memory()
    X_host = X_host_t.numpy()
    X_host[:] = synth_host(params, n, seed=1000 + rank)
    X_dev = X_host_t.cuda(non_blocking=False)
    lab_d = torch.empty(n, dtype=torch.int64, device="cuda")
    conf_d = torch.empty(n, dtype=torch.float64, device="cuda")
    prob_d = torch.empty((n, k), dtype=torch.float64, device="cuda")

    dm = DeviceModel(params, local)
    dm.enable_timing(True)
    stream = torch.cuda.current_stream().cuda_stream
    MODE = _lib.MODES[mode]

    def step_device():
        dm.predict_raw(X_dev, n, _lib.WDX_F64, MODE, lab_d, conf_d, prob_d, None, None, stream=stream)

    # ---- value: inputs resident in HBM, CUDA events on the launch stream ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, kernel_launches = 0.0, 0
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = sum_over_ranks(_lib.kernel_launch_count() - launches0)
    # fused DTW+SVC launches of the LAST step, CUDA events on the launch stream; in GUARDED mode the
    # dominant kernel is the FAST_F32 pass (the EXACT re-run of boundary reads is reported separately)
    kms, kl = dm.last_kernel_ms_mode(mode == "exact")
    kms_all, kl_all = dm.last_kernel_ms()
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # labels gathered host-side in shard order (the only cross-rank exchange of the path)
    labels_host = lab_d.cpu().numpy()
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, labels_host[:1000])
        label_sample = np.concatenate(gathered)
    else:
        label_sample = labels_host[:1000]

    # ---- e2e: host buffers through the reference-facing API ------------------
    mdl = DTW_SVM(params, device=local, mode=mode)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    mdl.predict(X_host[: min(n, 1 << 18)], nproc=1)  # warm-up: creates the device replica, staging buffers
    mdl.predict(X_host, nproc=1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y_pred, y_prob = mdl.predict(X_host, nproc=1, return_df=False)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    barrier()
    e2e_value = n_total / e2e_s
    e2e_match = bool(np.array_equal(y_pred, labels_host))

    # ---- other arithmetic modes on a smaller batch (context for `value`) ------
    modes = {}
    if rank == 0 and args.extra_modes:
        n_small = min(n, 1 << 20)
        for mname in ("fast", "exact", "guarded"):
            if mname == mode:
                continue
            M2 = _lib.MODES[mname]
            for _ in range(2):
                dm.predict_raw(X_dev, n_small, _lib.WDX_F64, M2, lab_d, conf_d, prob_d, None, None, stream=stream)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dm.predict_raw(X_dev, n_small, _lib.WDX_F64, M2, lab_d, conf_d, prob_d, None, None, stream=stream)
            b.record()
            torch.cuda.synchronize()
            modes[mname] = {"reads_per_s": n_small / (a.elapsed_time(b) * 1e-3), "batch": n_small}
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused DTW+SVC, FP32 CUDA-core issue) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    exact = mode == "exact"
    slots, lanes = (6, 64) if exact else (5, 128)
    # fused launches of one step process n reads (GUARDED adds the small exact re-run; count the first launch's cells)
    cells_per_launch_set = n * cells_per_read
    kernel_s = kms * 1e-3
    achieved = cells_per_launch_set * slots / kernel_s / 1e12           # T lane-ops/s
    peak = n_sm * lanes * sm_max * 1e6 / 1e12
    sm_now = (clocks or {}).get("sm_mhz") or sm_max
    roofline = {
        "bound": "fp64_alu" if exact else "fp32_alu",
        "kernel": "dtw_svc_kernel<%s,25,15,%d>" % ("EXACT" if exact else "FAST", 10),
        "achieved": achieved, "peak": peak, "unit": "T lane-op/s", "frac": achieved / peak,
        "frac_at_measured_clock": achieved / (n_sm * lanes * sm_now * 1e6 / 1e12),
        "definition": f"DTW cells/s x {slots} issue slots per cell / ({n_sm} SM x {lanes} lanes x f_SM); "
                      f"peak uses clocks.max.sm={sm_max:.0f} MHz (MEASURED_PEAKS.json), SURVEY.md 8(d)",
        "gcups": cells_per_launch_set / kernel_s / 1e9,
        "kernel_ms_per_step": kms, "kernel_launches_per_step": kl,
        "kernel_share_of_step": kms / ms_per_step,
        "all_fused_launches_ms_per_step": kms_all, "all_fused_launches_per_step": kl_all,
        "traffic": None,
        "hbm_algorithmic_bytes_per_launch": n * (params.L * 8 + 2 * (k - 1) * k * 8),
    }
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        roofline["traffic"] = prof.get("dram_bytes_per_launch")
        roofline["traffic_note"] = prof.get("note")
    except Exception:  # noqa: BLE001
        pass

    # ---- CPU baseline on this box's host cores (bounded sample) ------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rps, ns, dt = cpu_arm(params, threads, 15.0)
        cpu = {"value": rps, "unit": "reads/s", "cores": threads, "kind": "port",
               "sample": f"{ns} S1 reads of the same workload, minibatches of <=1000 over {threads} single-threaded "
                         f"workers, {dt:.1f} s (oracle/wdx_oracle.c)"}

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if exact else "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL}: {args.reads_per_gpu} synthetic S1 fingerprints per GPU per step "
                               f"({n_total} total), mode {mode}",
                   "model": MODEL, "n_sv": params.n_sv, "classes": k, "L": params.L, "window": params.window,
                   "reads_per_gpu": args.reads_per_gpu, "mode": mode, "parallelism": f"reads sharded x{world}",
                   "cache": f"input {n * params.L * 8 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)"},
        "gcups": value * cells_per_read / 1e9,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(n * params.L * 8),
                "d2h_bytes_per_step": int(n * (8 + 8 + 8 * k + 1)), "steps": e2e_steps,
                "api": "DTW_SVM.predict(X_host, nproc=1) -> (y_pred, y_prob); pinned host input, "
                       "host perf_counter around the blocking calls, max over ranks",
                "labels_equal_device_run": e2e_match},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "modes": modes,
        "label_histogram": {str(int(a)): int(b) for a, b in zip(*np.unique(label_sample, return_counts=True))},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("WDX_BENCH_MODE", "guarded"), choices=["exact", "fast", "guarded"])
    ap.add_argument("--reads-per-gpu", type=int, default=READS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-modes", dest="extra_modes", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
