"""Curated summary of one kernel of an .ncu-rep (ncu --set full) as JSON, for profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [launch_index] > profiles/x_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
launch__waves_per_multiprocessor sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed.avg.per_cycle_active smsp__inst_executed.sum
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active
sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu.sum
sm__cycles_elapsed.avg.per_second""".split()


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    col = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = {"source": f"ncu --set full --clock-control none ({rep})", "kernel": col["Kernel Name"][0]}
    for k in KEYS:
        for name, (v, u) in col.items():
            if name == k or name.endswith("." + k):
                out[k] = f"{v} {u}".strip()
                break
    for name, (v, u) in col.items():
        if "issue_stalled" in name and name.endswith("per_issue_active.ratio") and "average_warps" in name:
            try:
                if float(v) >= 0.15:
                    out[name] = f"{v} {u}".strip()
            except ValueError:
                pass
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
