import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context
with Context(0) as ctx:
    for which, log_n in ((0, 24), (2, 24), (0, 26), (2, 26), (1, 24), (6, 24)):
        print(which, log_n, round(ctx.bench_fused(which, log_n, 10), 4))
