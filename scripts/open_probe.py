import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from jolt_atlas_b200 import api as A, workload as W
ELL = int(sys.argv[1]) if len(sys.argv) > 1 else 18      # JA_MSM_PROFILE=1 prints the per-stage MSM timings
with A.Context(0) as ctx:
    srs = A.SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << ELL).precompute()
    rlc = A.MultilinearPolynomial.random(ctx, 1 << ELL, 9)
    pt = W._challenges(np.random.default_rng(1), ELL)
    for i in range(3):
        t = A.Blake2bTranscriptState(b"x"); t0 = time.perf_counter(); A.hyperkzg_open(ctx, srs, rlc, pt, t); print("open ms", (time.perf_counter()-t0)*1e3)
