"""A handful of single launches of the hot kernels at sizes larger than L2, for `ncu --set full` (never a bench number).
Order: bind LowToHigh 2^24, bind HighToLow 2^24, round-eval MUL 2^24, round-eval ADD 2^24, product-of-16 2^20,
product-of-4 2^22, one-hot point sums 20 x 2^16, MSM 2^20 (accumulate)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jolt_atlas_b200 import api as A  # noqa: E402
from jolt_atlas_b200 import SRS, Context, workload as W  # noqa: E402

rng = np.random.default_rng(5)
with Context(0) as ctx:
    ch = W._challenges(rng, 40)
    a = A.MultilinearPolynomial.random(ctx, 1 << 24, 1)
    b = A.MultilinearPolynomial.random(ctx, 1 << 24, 2)
    eq = A.GruenSplitEqPolynomial(ctx, ch[:24], 0)
    A.round_eval(ctx, A.EvalKernel.MUL, [a, b], eq)
    A.round_eval(ctx, A.EvalKernel.ADD, [a, b], eq)
    a.bind_parallel(ch[30], 0)
    b.bind_parallel(ch[30], 1)
    a.free(); b.free(); eq.free()
    ps = [A.MultilinearPolynomial.random(ctx, 1 << 20, 10 + i) for i in range(16)]
    eq = A.GruenSplitEqPolynomial(ctx, ch[:20], 0)
    A.round_eval(ctx, A.EvalKernel.PROD, ps, eq, n_out=16)
    A.bind_many(ctx, ps, ch[31], 0)
    for p in ps:
        p.free()
    eq.free()
    srs = SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << 20)
    T = 1 << 16
    lists = [rng.integers(0, 16, size=T, dtype=np.uint64) * np.uint64(T) + np.arange(T, dtype=np.uint64) for _ in range(20)]
    A.g1_sum_indexed_batch(ctx, srs, lists)
    s = A.MultilinearPolynomial.random(ctx, 1 << 20, 77)
    A.msm_fr(ctx, srs, s)
    s.free(); srs.free()
print("done")
