"""Single launches of the fused bind+evaluate round kernels at sizes larger than L2, for `ncu --set full` (never a bench
number).  Each ja_bench_fused call = 3 warm-up launches + 1 timed launch of one kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context  # noqa: E402

with Context(0) as ctx:
    for which, log_n in ((0, 24), (1, 24), (2, 24), (6, 24), (3, 22), (4, 20), (5, 20)):
        print(which, log_n, ctx.bench_fused(which, log_n, 1))
