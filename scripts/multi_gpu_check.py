"""torchrun job (one rank per GPU, NCCL): the sharded MSM, the sharded HyperKZG opening and sharded sumcheck round
evaluations equal their single-GPU results bit for bit on every rank.  Also times the sharded opening.
usage: python -m torch.distributed.run --nproc-per-node N scripts/multi_gpu_check.py [ell]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from jolt_atlas_b200 import api as A  # noqa: E402
from jolt_atlas_b200 import parallel as PAR  # noqa: E402
from jolt_atlas_b200 import workload as W  # noqa: E402

ell = int(sys.argv[1]) if len(sys.argv) > 1 else 16
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = PAR.Comm(device=torch.device("cuda", local))
rng = np.random.default_rng(9)           # same seed on every rank: replicated inputs
with A.Context(local) as ctx:
    n = 1 << ell
    srs = A.SRS.generate(ctx, bench.g1_generator_mont(), bench.tau_mont(), n).precompute()
    poly = A.MultilinearPolynomial.random(ctx, n, 21)
    # 1. MSM split by index range
    want, winf = A.msm_fr(ctx, srs, poly)
    got, ginf = PAR.sharded_msm_fr(ctx, srs, poly, comm)
    assert ginf == winf and np.array_equal(got, want), "sharded MSM differs"
    # 2. HyperKZG::open, MSMs sharded, two all-gathers
    point = W._challenges(rng, ell)
    t1, t2 = A.Blake2bTranscriptState(b"open"), A.Blake2bTranscriptState(b"open")
    ref = A.hyperkzg_open(ctx, srs, poly, point, t1)
    out = PAR.sharded_hyperkzg_open(ctx, srs, poly, point, t2, comm)
    for k in ("com", "w", "v"):
        assert np.array_equal(ref[k], out[k]), "sharded opening differs in " + k
    assert t1.state == t2.state
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        PAR.sharded_hyperkzg_open(ctx, srs, poly, point, A.Blake2bTranscriptState(b"open"), comm)
    dist.barrier(); t_sh = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(3):
        A.hyperkzg_open(ctx, srs, poly, point, A.Blake2bTranscriptState(b"open"))
    t_one = (time.perf_counter() - t0) / 3
    # 3. sumcheck rounds over contiguous hypercube slices (Mul body + product of 4)
    m = min(ell, 14)
    N = 1 << m
    host = rng.integers(0, 1 << 63, size=(4, N, 4), dtype=np.uint64)
    host[..., 3] &= np.uint64((1 << 60) - 1)
    lo, hi = comm.plan.slice_range(N)
    full = [A.MultilinearPolynomial.from_fr(ctx, host[i]) for i in range(4)]
    mine = [A.MultilinearPolynomial.from_fr(ctx, host[i, lo:hi]) for i in range(4)]
    eq = A.GruenSplitEqPolynomial(ctx, W._challenges(rng, m), 0)
    for rnd in range(comm.plan.local_rounds(N)):
        assert np.array_equal(PAR.sharded_round_eval(ctx, 2, mine[:2], eq, comm, 2), A.round_eval(ctx, 2, full[:2], eq)), rnd
        assert np.array_equal(PAR.sharded_round_eval(ctx, 4, mine, eq, comm, 4), A.round_eval(ctx, 4, full, eq, n_out=4)), rnd
        ch = W._challenges(rng, 1)[0]
        eq.bind(ch)
        A.bind_many(ctx, full, ch, 0)
        A.bind_many(ctx, mine, ch, 0)
    # 4. the full sharded Sumcheck::prove loop over NCCL (Mul at 2^m, product of 4)
    host2 = rng.integers(0, 1 << 63, size=(4, N, 4), dtype=np.uint64)
    host2[..., 3] &= np.uint64((1 << 60) - 1)
    w2, claim2 = W._challenges(rng, m), W._challenges(rng, 1)[0]
    for kind, npoly in ((2, 2), (4, 4)):
        tr_one, tr_sh = A.Blake2bTranscriptState(b"sc"), A.Blake2bTranscriptState(b"sc")
        want = A.sumcheck_prove(ctx, kind, [A.MultilinearPolynomial.from_fr(ctx, host2[i]) for i in range(npoly)], claim2, tr_one, eq_w=w2)
        got = PAR.sharded_sumcheck_prove(ctx, kind, [A.MultilinearPolynomial.from_fr(ctx, host2[i, lo:hi]) for i in range(npoly)],
                                         claim2, tr_sh, comm, w2)
        assert all(np.array_equal(a, b) for a, b in zip(got["coeffs"], want["coeffs"])), "sharded sumcheck differs"
        assert np.array_equal(got["final_claims"], want["final_claims"]) and tr_one.state == tr_sh.state
    # 5. the library's own NCCL communicator (csrc/comm.cu): the ORDINARY single-GPU calls run sharded, the exchange is inside them
    lc = PAR.LibComm(ctx)
    allg = lc.all_gather(np.arange(5, dtype=np.uint64) + 100 * lc.rank)
    assert allg.shape == (lc.world, 5) and all(int(allg[r, 0]) == 100 * r for r in range(lc.world)), "ja_comm_allgather"
    want_msm, want_inf = A.msm_fr(ctx, srs, poly)
    lc.shard_on()
    got, ginf = A.msm_fr(ctx, srs, poly)
    t3 = A.Blake2bTranscriptState(b"open")
    out3 = A.hyperkzg_open(ctx, srs, poly, point, t3)
    lc.shard_off()
    assert ginf == want_inf and np.array_equal(got, want_msm), "in-library sharded MSM differs"
    for k in ("com", "w", "v"):
        assert np.array_equal(ref[k], out3[k]), "in-library sharded opening differs in " + k
    assert t3.state == t1.state
    # one-hot commitments dealt to the ranks inside ja_addr_commit_many
    hk = rng.integers(0, 16, size=(5, n // 16), dtype=np.uint32)
    batches = [A.OneHotAddresses(ctx, hk[:3], 16), A.OneHotAddresses(ctx, hk[3:], 16)]
    ref_c = A.commit_one_hot_batches(ctx, srs, batches)
    lc.shard_on()
    got_c = A.commit_one_hot_batches(ctx, srs, batches)
    lc.shard_off()
    for (a, ai), (b, bi) in zip(ref_c, got_c):
        assert np.array_equal(a, b) and np.array_equal(np.asarray(ai, dtype=bool), np.asarray(bi, dtype=bool)), "in-library dealt commitments differ"
    # sharded Sumcheck::prove with the partial sums all-gathered by the library
    for kind, npoly in ((2, 2), (4, 4)):
        tr_one, tr_sh = A.Blake2bTranscriptState(b"sc"), A.Blake2bTranscriptState(b"sc")
        want2 = A.sumcheck_prove(ctx, kind, [A.MultilinearPolynomial.from_fr(ctx, host2[i]) for i in range(npoly)], claim2, tr_one, eq_w=w2)
        got2 = PAR.sharded_sumcheck_prove(ctx, kind, [A.MultilinearPolynomial.from_fr(ctx, host2[i, lo:hi]) for i in range(npoly)],
                                          claim2, tr_sh, lc, w2)
        assert all(np.array_equal(a, b) for a, b in zip(got2["coeffs"], want2["coeffs"])), "in-library sharded sumcheck differs"
        assert np.array_equal(got2["final_claims"], want2["final_claims"]) and tr_one.state == tr_sh.state
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    lc.shard_on()
    for _ in range(3):
        A.hyperkzg_open(ctx, srs, poly, point, A.Blake2bTranscriptState(b"open"))
    lc.shard_off()
    dist.barrier(); t_lib = (time.perf_counter() - t0) / 3
    lc.close()
    if comm.rank == 0:
        print("in-library exchange: open sharded %.2f ms" % (t_lib * 1e3), flush=True)
        print("multi-gpu ok: world=%d ell=%d  open sharded %.2f ms vs single-GPU %.2f ms" % (comm.world, ell, t_sh * 1e3, t_one * 1e3), flush=True)
dist.barrier()
dist.destroy_process_group()
