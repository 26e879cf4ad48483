"""Wall-clock (host-visible) cost of the individual stages of a nanoGPT-shaped node -> gpurun_out/stage_probe.json."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jolt_atlas_b200 import api as A  # noqa: E402
from jolt_atlas_b200 import SRS, Context, workload as W  # noqa: E402

out = {}
with Context(0) as ctx:
    inputs = W.build_inputs("nanoGPT")
    ell = inputs["ell"]
    srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << ell).precompute()
    ni = inputs["nodes"][0]
    T = 1 << ni.spec.log_t
    hot = A.OneHotBatch(ctx, W.onehot_index_lists(ni))
    ra = [A.MultilinearPolynomial.from_lookup(ctx, ni.tables[j], ni.hot_k[j]) for j in range(ni.d_hot)]
    hot16 = A.OneHotAddresses(ctx, ni.hot_k[:16], 16)
    hot4 = A.OneHotAddresses(ctx, ni.hot_k[16:], 16)
    t = A.Blake2bTranscriptState(b"x")

    def timeit(name, fn, reps=20):
        fn(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps * 1e6
        out[name] = dt
        print("%-40s %10.1f us" % (name, dt), flush=True)

    timeit("onehot_commit 20x2^14", lambda: hot.commit(srs))
    timeit("addr16.commit 16x2^14", lambda: hot16.commit(srs))
    timeit("addr16.ra_evals", lambda: hot16.ra_evals(ni.eq_w))
    timeit("addr16.gather", lambda: [q.free() for q in hot16.gather(ni.tables[:16])])
    out_d = {"finals": [], "msg_bytes": 0}
    timeit("RA checks batch d=16 (18 rounds)", lambda: W._ra_checks(A, ctx, hot16, ni, 0, 16, inputs["claim"], t, out_d, None).free())
    timeit("RA checks batch d=4 (18 rounds)", lambda: W._ra_checks(A, ctx, hot4, ni, 16, 20, inputs["claim"], t, out_d, None).free())
    timeit("clone 2^14", lambda: ra[0].clone().free())
    timeit("clone x16 2^14", lambda: [q.free() for q in [p.clone() for p in ra[:16]]])

    def sc(kind, polys, **kw):
        ps = [p.clone() for p in polys]
        A.sumcheck_prove(ctx, kind, ps, inputs["claim"], t, **kw)
        for p in ps:
            p.free()
    timeit("sumcheck IDENT 2^14 (14 rounds)", lambda: sc(A.EvalKernel.IDENT, ra[:1], eq_w=ni.eq_w))
    timeit("sumcheck SUM1 x16 2^14", lambda: sc(A.EvalKernel.SUM1, ra[:16], gammas=ni.gammas[:16]))
    timeit("sumcheck PROD16 2^14", lambda: sc(A.EvalKernel.PROD, ra[:16], eq_w=ni.eq_w))
    timeit("sumcheck PROD4 2^14", lambda: sc(A.EvalKernel.PROD, ra[16:20], eq_w=ni.eq_w))
    timeit("sumcheck MUL 2^14", lambda: sc(A.EvalKernel.MUL, ra[:2], eq_w=ni.eq_w))
    for lg in (4, 8, 10, 12):
        small = [A.MultilinearPolynomial.random(ctx, 1 << lg, 3 + i) for i in range(2)]
        timeit("sumcheck MUL 2^%d" % lg, lambda: sc(A.EvalKernel.MUL, small, eq_w=ni.eq_w[:lg]))
        timeit("sumcheck DOT2 2^%d" % lg, lambda: sc(A.EvalKernel.DOT2, small))
    rlc = A.MultilinearPolynomial.random(ctx, 1 << ell, 9)
    timeit("hyperkzg_open ell=18", lambda: A.hyperkzg_open(ctx, srs, rlc, inputs["open_point"], t), reps=5)
    timeit("msm_fr 2^18", lambda: A.msm_fr(ctx, srs, rlc), reps=5)
    timeit("eq_evals m=6", lambda: A.EqPolynomial.evals(ctx, ni.eq_rows).free())
    e = A.EqPolynomial.evals(ctx, ni.eq_rows)
    timeit("tensor_fold 64x64", lambda: A.tensor_fold_i32(ctx, ni.A, e, transpose=False).free())
    timeit("full pass resident", lambda: None, reps=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/stage_probe.json", "w"), indent=1)
