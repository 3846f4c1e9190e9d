"""bench.raw_signal_chain alone (stage times on the real reads, run / stream throughput): one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from warpdemux_b200 import model_io  # noqa: E402

small = model_io.load_npz(os.path.join(ROOT, "tests", "golden", "models", "WDX4_rna004_v1_0.npz"))
out = bench.raw_signal_chain(small, 0)
out.pop("validate_cpu_baseline", None)
print(json.dumps(out))
