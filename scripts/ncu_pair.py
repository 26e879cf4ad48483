"""Single launches of the paired RA-check kernel (product of 16 + booleanity over 16) in its three forms, for
`ncu --set full` (never a bench number): 64 threads per pair at 2^8 pairs, 128-thread blocks at 2^10 pairs, 256-thread blocks
at 2^14 pairs.  Each ja_bench_fused call = 3 warm-up launches + 1 timed launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import Context  # noqa: E402

FORMS = {"wide": (11, 10), "small": (12, 12), "large": (10, 16)}
with Context(0) as ctx:
    for name in (sys.argv[1:] or list(FORMS)):
        which, log_n = FORMS[name]
        print(name, which, log_n, ctx.bench_fused(which, log_n, 1))
