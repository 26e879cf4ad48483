"""TMA-staged vs register-staged fused round kernels: ms per launch and algorithmic GB/s at sizes beyond L2."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
out = []
with Context(0) as ctx:
    for name, which, npoly in (("add regs", 0, 2), ("add tma", 7, 2), ("ident regs", 2, 1), ("ident tma", 8, 1)):
        for log_n in (22, 24, 26):
            ms = ctx.bench_fused(which, log_n, 20)
            gb = 48 * (1 << log_n) * npoly / ms / 1e6
            out.append({"kernel": name, "log_n": log_n, "ms": round(ms, 4), "GBps": round(gb, 1), "frac_hbm": round(gb / peak, 3)})
            print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tma_probe.json", "w"), indent=1)
