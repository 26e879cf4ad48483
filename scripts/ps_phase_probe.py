"""Per-launch time of the ps_shout phase pass (k_ps_phase, device time by CUDA events through the library's profile counters is not
exposed per launch: wall time of init_phase incl. the flag wait) at nanoGPT / GPT-2 node sizes, clamp (m = 256, 6 suffixes) and
remainder (m = 128, 2 suffixes) shapes.  usage: JA_PS_TILE_LOG=<k> python scripts/ps_phase_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jolt_atlas_b200 import api as A, Context, workload as W

with Context(0) as ctx:
    for log_t in (12, 14, 16):
        rng = np.random.default_rng(log_t)
        T = 1 << log_t
        q = rng.integers(-64, 64, size=T).astype(np.int64).view(np.uint64)
        rem = rng.integers(0, 1 << 14, size=T, dtype=np.uint64)
        r = W._challenges(rng, log_t)
        for name, idx, log_k, phases, kinds, bound in (("clamp", q, 64, 8, W.PS_SUFFIXES, 31), ("rem", rem, 14, 2, (5, 4), 0)):
            m = 1 << (log_k // phases)
            vs = [W._challenges(rng, m) for _ in range(phases)]
            best = [1e9] * phases
            for rep in range(6):
                ps = A.PrefixSuffixShout(ctx, idx, r, log_k, phases)
                ctx.sync()
                for ph in range(phases):
                    t0 = time.perf_counter(); ps.init_phase(ph, vs[ph - 1] if ph else None, kinds, bound); dt = (time.perf_counter() - t0) * 1e6
                    best[ph] = min(best[ph], dt)
                ps.free()
            print("tile_log", os.environ.get("JA_PS_TILE_LOG", "10"), name, "log_t", log_t, "us per phase (best of 6):", [round(x, 1) for x in best], flush=True)
