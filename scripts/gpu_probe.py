"""Exploration sweep run under gpurun: calibration + per-kernel timings -> gpurun_out/probe.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context  # noqa: E402

out = {}
with Context(0) as ctx:
    out["fr_mul_per_s"] = [ctx.calibrate_fr_mul(4000) for _ in range(3)]
    rows = []
    for which, name in [(0, "bind_l2h"), (1, "bind_h2l"), (2, "eval_mul"), (3, "eval_dot2"), (4, "eval_add")]:
        for log_n in (14, 18, 20, 22, 24, 26):
            for npoly in ((1, 2, 4) if which <= 1 and log_n == 24 else (1,)):
                ms = ctx.bench_kernel(which, log_n, npoly, 10 if log_n >= 24 else 30)
                n = 1 << log_n
                if which <= 1:
                    bytes_ = 48 * (n // 2) * 2 * npoly / 2   # 32n read + 16n write per poly
                    bytes_ = (32 * n + 16 * n) * npoly
                else:
                    bytes_ = 32 * n * 2
                rows.append({"kernel": name, "log_n": log_n, "n_polys": npoly, "ms": ms, "GBps": bytes_ / ms / 1e6})
                print(rows[-1], flush=True)
    out["kernels"] = rows
    out["launches"] = ctx.launch_count()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out["fr_mul_per_s"]))
