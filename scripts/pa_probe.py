import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from jolt_atlas_b200 import api as A, Context, workload as W
with Context(0) as ctx:
    inputs = W.build_inputs("nanoGPT")
    ni = inputs["nodes"][0]
    for rep in range(4):
        ps = A.PrefixSuffixShout(ctx, ni.acc, ni.eq_w)
        t = A.Blake2bTranscriptState(b"x")
        ctx.sync()
        t0 = time.perf_counter(); ps.prove_address(t, ni.gammas[0], 31); print("prove_address us", (time.perf_counter() - t0) * 1e6, file=sys.stderr)
        ps.free()
