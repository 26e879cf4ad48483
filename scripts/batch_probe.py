"""Where a [RaVirtual, Hamming, Booleanity] batch of one node spends its host-visible time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jolt_atlas_b200 import api as A
from jolt_atlas_b200 import Context, workload as W
with Context(0) as ctx:
    inputs = W.build_inputs("nanoGPT")
    ni = inputs["nodes"][0]
    addr = A.OneHotAddresses(ctx, ni.hot_k[:16], 16)
    t = A.Blake2bTranscriptState(b"x")
    claim = inputs["claim"]
    T = {}
    def lap(name, t0):
        T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e6
    reps = 30
    for rep in range(reps + 3):
        if rep == 3: T.clear()
        t0 = time.perf_counter(); G = addr.ra_evals(ni.eq_w); lap("ra_evals", t0)
        t0 = time.perf_counter(); ra = addr.gather(ni.tables[:16]); lap("gather", t0)
        t0 = time.perf_counter(); first = ra[0].clone(); lap("clone", t0)
        t0 = time.perf_counter()
        r = A.batched_sumcheck_prove(ctx, [
            {"kind": A.EvalKernel.PROD, "polys": ra, "eq_w": ni.eq_w, "claim": claim},
            {"kind": A.InstanceKind.HAMMING_TABLES, "tables": G, "aux_fr": ni.gammas[:16], "claim": claim},
            {"kind": A.InstanceKind.BOOLEANITY, "tables": G, "addr": addr, "eq_w": ni.eq_w, "gammas": ni.gammas[:16], "r_address": ni.r_addr}], t)
        lap("batched_sumcheck_prove", t0)
        t0 = time.perf_counter()
        for q in ra: q.free()
        first.free(); lap("free", t0)
    print({k: round(v / reps, 1) for k, v in T.items()})
