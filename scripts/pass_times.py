"""Per-pass wall times with a per-stage breakdown (API-level timers) to locate sporadic slow passes."""
import os, sys, time, collections, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jolt_atlas_b200 import api as A
from jolt_atlas_b200 import SRS, Context, workload as W
acc = collections.defaultdict(float)
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[name] += (time.perf_counter() - t0) * 1e3
    setattr(mod, name, g)
for n in ("batched_sumcheck_prove", "sumcheck_prove", "hyperkzg_open", "commit_one_hot_batches", "tensor_fold_i32"):
    wrap(A, n)
for n in ("ra_evals", "gather"):
    wrap(A.OneHotAddresses, n)
if len(sys.argv) > 2: gc.disable()
with Context(0) as ctx:
    inputs = W.build_inputs("nanoGPT")
    srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << inputs["ell"]).precompute()
    res = W.make_resident(ctx, inputs)
    for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
        acc.clear()
        t0 = time.perf_counter(); W.run_device(ctx, srs, inputs, resident=res); ctx.sync()
        dt = (time.perf_counter() - t0) * 1e3
        print("pass %2d %6.0f ms | " % (i, dt) + " ".join("%s=%.0f" % (k[:14], v) for k, v in sorted(acc.items())), flush=True)
