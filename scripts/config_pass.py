"""One prove-shaped pass of a named config (microgpt / nanoGPT / gpt2) on one GPU with an API-level stage breakdown.
usage: python scripts/config_pass.py gpt2 [passes]   ->  gpurun_out/pass_<config>.json"""
import collections, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jolt_atlas_b200 import api as A
from jolt_atlas_b200 import SRS, Context, workload as W

config = sys.argv[1] if len(sys.argv) > 1 else "gpt2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
acc = collections.defaultdict(float)


def wrap(obj, name):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[name] += (time.perf_counter() - t0) * 1e3
    setattr(obj, name, g)


for n in ("batched_sumcheck_prove", "sumcheck_prove", "hyperkzg_open", "commit_one_hot_batches", "tensor_fold_i32"):
    wrap(A, n)
for n in ("ra_evals", "gather"):
    wrap(A.OneHotAddresses, n)
wrap(A.TensorI32, "fold")
wrap(A.MultilinearPolynomial, "rlc_add_onehot")
out = {"config": config, "passes": []}
with Context(0) as ctx:
    t0 = time.perf_counter()
    inputs = W.build_inputs(config)
    out["build_inputs_s"] = round(time.perf_counter() - t0, 2)
    t0 = time.perf_counter()
    srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << inputs["ell"]).precompute()
    ctx.sync()
    out["srs_generate_and_window_table_s"] = round(time.perf_counter() - t0, 2)
    t0 = time.perf_counter()
    res = W.make_resident(ctx, inputs)
    ctx.sync()
    out["upload_inputs_s"] = round(time.perf_counter() - t0, 2)
    out["units"] = W.count_units(inputs)
    for i in range(passes):
        acc.clear()
        ctx.timer_begin()
        t0 = time.perf_counter()
        r = W.run_device(ctx, srs, inputs, resident=res)
        dev_ms = ctx.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        row = {"pass": i, "wall_ms": round(wall, 1), "device_ms": round(dev_ms, 1), "stages_ms": {k: round(v, 1) for k, v in sorted(acc.items())}}
        out["passes"].append(row)
        print(row, flush=True)
    out["final_state"] = r["states"][-1].hex()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/pass_%s.json" % config, "w"), indent=1)
