"""GPT-2-scale single stages on one GPU (SURVEY 8d config 4/5 shapes): RA one-hot checks at T = 2^20, one-hot commitments of
20 x 2^20, opening-reduction group, HyperKZG open at ell = 22/24, MSM 2^24.  Wall-clock per call -> gpurun_out/scale_probe.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jolt_atlas_b200 import api as A
from jolt_atlas_b200 import Context, workload as W
out = {}
def timeit(ctx, name, fn, reps=3):
    fn(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.sync()
    out[name] = round((time.perf_counter() - t0) / reps * 1e3, 3)
    print("%-46s %10.3f ms" % (name, out[name]), flush=True)
log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ell = log_t + 4
rng = np.random.default_rng(1)
with Context(0) as ctx:
    T = 1 << log_t
    t0 = time.perf_counter()
    srs = A.SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), 1 << ell)
    print("SRS generate 2^%d: %.2f s" % (ell, time.perf_counter() - t0), flush=True)
    t0 = time.perf_counter(); srs.precompute(); print("window table: %.2f s" % (time.perf_counter() - t0), flush=True)
    k = rng.integers(0, 16, size=(16, T), dtype=np.uint32)
    eq_w, r_addr, gam = W._challenges(rng, log_t), W._challenges(rng, 4), W._challenges(rng, 16)
    tables = np.stack([W._challenges(rng, 16) for _ in range(16)])
    claim = W._challenges(rng, 1)[0]
    t0 = time.perf_counter(); addr = A.OneHotAddresses(ctx, k, 16); out["upload 16 x 2^%d addresses" % log_t] = round((time.perf_counter() - t0) * 1e3, 3)
    tr = A.Blake2bTranscriptState(b"x")
    timeit(ctx, "one-hot commit 16 x 2^%d" % log_t, lambda: addr.commit(srs))
    timeit(ctx, "compute_ra_evals 16 x 2^%d" % log_t, lambda: addr.ra_evals(eq_w))
    def ra_batch():
        G = addr.ra_evals(eq_w)
        ra = addr.gather(tables)
        A.batched_sumcheck_prove(ctx, [
            {"kind": A.EvalKernel.PROD, "polys": ra, "eq_w": eq_w, "claim": claim},
            {"kind": A.InstanceKind.HAMMING_TABLES, "tables": G, "aux_fr": gam, "claim": claim},
            {"kind": A.InstanceKind.BOOLEANITY, "tables": G, "addr": addr, "eq_w": eq_w, "gammas": gam, "r_address": r_addr}], tr)
        for q in ra: q.free()
    timeit(ctx, "RA one-hot checks batch d=16, T=2^%d (%d rounds)" % (log_t, log_t + 4), ra_batch)
    def opening():
        A.batched_sumcheck_prove(ctx, [{"kind": A.InstanceKind.OPENING_ONEHOT, "addr": addr, "eq_w": eq_w, "r_address": r_addr,
                                        "claims": np.broadcast_to(claim, (16, 4))}], tr)
    timeit(ctx, "opening reduction group d=16, T=2^%d" % log_t, opening)
    a = A.MultilinearPolynomial.random(ctx, T, 5); b = A.MultilinearPolynomial.random(ctx, T, 6)
    def mul_sc():
        A.sumcheck_prove(ctx, A.EvalKernel.MUL, [a.clone(), b.clone()], claim, tr, eq_w=eq_w)
    timeit(ctx, "Mul sumcheck T=2^%d" % log_t, mul_sc)
    rlc = A.MultilinearPolynomial.random(ctx, 1 << ell, 9)
    pt = W._challenges(rng, ell)
    timeit(ctx, "HyperKZG open ell=%d" % ell, lambda: A.hyperkzg_open(ctx, srs, rlc, pt, tr), reps=2)
    timeit(ctx, "MSM 2^%d" % ell, lambda: A.msm_fr(ctx, srs, rlc), reps=2)
    out["MSM Mscalar/s"] = round((1 << ell) / out["MSM 2^%d" % ell] / 1e3, 1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/scale_probe_%d.json" % log_t, "w"), indent=1)
print(out)
