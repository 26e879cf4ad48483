"""Live per-launch time of the fused round kernels at SMALL slabs (back-to-back launches, warm caches): the latency floor
of a round kernel, as opposed to the cold-cache ncu figure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context
names = {0: "add", 1: "mul", 2: "ident", 3: "prod4", 4: "prod16", 5: "bool16", 6: "open_h2l", 9: "prod16w", 10: "pair16", 11: "pair16w", 12: "pair16s"}
with Context(0) as ctx:
    for which in (2, 0, 1, 3, 4, 9, 5, 10, 11, 12):
        row = []
        for log_n in (6, 8, 10, 11, 12, 13, 14, 15, 16, 18):
            row.append("%5.1f" % (ctx.bench_fused(which, log_n, 200) * 1e3))
        print("%-8s us/launch at log_n 6,8,10,11,12,13,14,15,16,18: %s" % (names[which], " ".join(row)), flush=True)
