"""One prove-shaped pass on device-resident inputs, for ncu (launch list / --set full captures).  Never a bench number.
usage: python scripts/ncu_pass.py [config] [passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jolt_atlas_b200 import SRS, Context, workload as W  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "nanoGPT"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with Context(0) as ctx:
    inputs = W.build_inputs(config)
    srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << inputs["ell"]).precompute()
    res = W.make_resident(ctx, inputs)
    ctx.sync()
    l0 = ctx.launch_count()
    for _ in range(passes):
        W.run_device(ctx, srs, inputs, resident=res)
    print("launches per pass:", (ctx.launch_count() - l0) // passes, "setup launches:", l0)
    W.free_resident(res)
    srs.free()
