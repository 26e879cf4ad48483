"""Single launches of the TMA-staged round kernels (ADD, IDENT) at 2^24 Fr per polynomial for `ncu --set full`
(never a bench number).  Each ja_bench_fused call = 3 warm-up launches + 1 timed launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context  # noqa: E402

with Context(0) as ctx:
    for which, log_n in ((7, 24), (8, 24), (0, 24)):
        print(which, log_n, ctx.bench_fused(which, log_n, 1))
