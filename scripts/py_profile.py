"""cProfile of the Python driver of one nanoGPT-shaped pass (where does the host-language glue spend its time)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jolt_atlas_b200 import SRS, Context, workload as W
with Context(0) as ctx:
    inputs = W.build_inputs(sys.argv[1] if len(sys.argv) > 1 else "nanoGPT")
    srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << inputs["ell"]).precompute()
    res = W.make_resident(ctx, inputs)
    for _ in range(3):
        W.run_device(ctx, srs, inputs, resident=res)
    ctx.sync()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        W.run_device(ctx, srs, inputs, resident=res)
    ctx.sync()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
