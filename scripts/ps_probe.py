"""Wall-clock cost of the ps_shout T-sized passes (ja_psshout_*) per phase at nanoGPT / GPT-2 node sizes -> gpurun_out/ps_probe.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import api as A, Context, workload as W
from jolt_atlas_b200 import parallel as PAR
out = {}
with Context(0) as ctx:
    for log_t in (12, 14, 16, 20):
        rng = np.random.default_rng(log_t)
        T = 1 << log_t
        small = rng.integers(-(1 << 20), 1 << 20, size=T); big = rng.integers(-(1 << 40), 1 << 40, size=T)
        idx = np.where(rng.integers(0, 8, size=T) == 0, big, small).astype(np.int64).view(np.uint64)
        r = W._challenges(rng, log_t)
        vs = [W._challenges(rng, 256) for _ in range(8)]
        rows = []
        for rep in range(4):
            t0 = time.perf_counter(); ps = A.PrefixSuffixShout(ctx, idx, r); t_new = time.perf_counter() - t0
            ph = []
            for phase in range(8):
                t0 = time.perf_counter(); ps.init_phase(phase, vs[phase - 1] if phase else None, W.PS_SUFFIXES, 31); ph.append((time.perf_counter() - t0) * 1e6)
            t0 = time.perf_counter(); ra = ps.materialize_ra(np.concatenate(vs)); ctx.sync(); t_ra = time.perf_counter() - t0
            ra.free(); ps.free()
            rows = {"new_us": round(t_new * 1e6, 1), "phase_us": [round(x, 1) for x in ph], "ra_us": round(t_ra * 1e6, 1)}
        out["log_t=%d" % log_t] = rows
        print(log_t, rows, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ps_probe.json", "w"), indent=1)
