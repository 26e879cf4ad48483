"""Exploration sweep run under gpurun: MSM timings over n, window width c and run length T -> gpurun_out/msm_probe.json."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context, MultilinearPolynomial, SRS, msm_fr  # noqa: E402

G1 = np.zeros(8, dtype=np.uint64)
# generator (1, 2) in Montgomery form (R mod q, 2R mod q)
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = (1 << 256) % Q
for k in range(4):
    G1[k] = (R >> (64 * k)) & ((1 << 64) - 1)
    G1[4 + k] = ((2 * R % Q) >> (64 * k)) & ((1 << 64) - 1)
BETA = np.array([0x1234567890abcdef, 0x0fedcba987654321, 0x1111111111111111, 0x0222222222222222], dtype=np.uint64)

max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rows = []
with Context(0) as ctx:
    t0 = time.time()
    srs = SRS.generate(ctx, G1, BETA, 1 << max_log)
    print("srs generate 2^%d: %.3f s" % (max_log, time.time() - t0), flush=True)
    min_log = int(os.environ.get('MIN_LOG', '16'))
    for log_n in range(min_log, max_log + 1, 2):
        p = MultilinearPolynomial.random(ctx, 1 << log_n, 7)
        cs = [None] if len(sys.argv) <= 2 else [int(x) for x in sys.argv[2].split(",")]
        ts = [None] if len(sys.argv) <= 3 else [int(x) for x in sys.argv[3].split(",")]
        occs = [None] if len(sys.argv) <= 4 else [int(x) for x in sys.argv[4].split(",")]
        for c, T, occ in [(c, T, o) for c in cs for T in ts for o in occs]:
            if True:
                if occ is not None:
                    os.environ["JA_MSM_OCC"] = str(occ)
                if c is not None:
                    os.environ["JA_MSM_C"] = str(c)
                if T is not None:
                    os.environ["JA_MSM_T"] = str(T)
                msm_fr(ctx, srs, p)
                best = 1e9
                for _ in range(3):
                    ctx.timer_begin()
                    msm_fr(ctx, srs, p)
                    best = min(best, ctx.timer_end())
                rows.append({"log_n": log_n, "c": c, "T": T, "occ": occ, "ms": best, "Mscalar_per_s": (1 << log_n) / best / 1e3})
                print(rows[-1], flush=True)
        p.free()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/msm_probe.json", "w"), indent=1)
