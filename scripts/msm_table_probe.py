"""Fixed-base window table sweep: width of the second (wide) table (0 = none) x job size (and the HyperKZG opening at the SRS size).
usage: python scripts/msm_table_probe.py LOG_SRS c1,c2,...  ->  gpurun_out/msm_table_probe_<LOG_SRS>.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from jolt_atlas_b200 import api as A
from jolt_atlas_b200 import Context, MultilinearPolynomial, SRS, msm_fr
from jolt_atlas_b200 import workload as W

log_srs = int(sys.argv[1]) if len(sys.argv) > 1 else 22
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 20]
rows = []
with Context(0) as ctx:
    for c in cs:
        os.environ["JA_MSM_TABLE2_C"] = str(c)
        t0 = time.time()
        srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << log_srs).precompute()
        ctx.sync()
        setup = time.time() - t0
        for log_n in range(max(16, log_srs - 6), log_srs + 1, 2):
            p = MultilinearPolynomial.random(ctx, 1 << log_n, 7)
            msm_fr(ctx, srs, p)
            best = 1e9
            for _ in range(3):
                ctx.timer_begin(); msm_fr(ctx, srs, p); best = min(best, ctx.timer_end())
            rows.append({"log_srs": log_srs, "table2_c": c, "log_n": log_n, "ms": round(best, 3), "Mscalar_per_s": round((1 << log_n) / best / 1e3, 1)})
            print(rows[-1], flush=True)
            p.free()
        rng = np.random.default_rng(5)
        point = W._challenges(rng, log_srs)
        poly = MultilinearPolynomial.random(ctx, 1 << log_srs, 9)
        best = 1e9
        for _ in range(2):
            q = poly.clone()
            t = A.Blake2bTranscriptState(b"probe")
            ctx.timer_begin(); A.hyperkzg_open(ctx, srs, q, point, t); best = min(best, ctx.timer_end())
            q.free()
        rows.append({"log_srs": log_srs, "table2_c": c, "hyperkzg_open_ms": round(best, 2), "setup_s": round(setup, 2)})
        print(rows[-1], flush=True)
        poly.free(); srs.free()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/msm_table_probe_%d.json" % log_srs, "w"), indent=1)
