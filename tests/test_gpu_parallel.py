"""Sharded entry points on the device: MSM index ranges, hypercube-slice round evaluation, the sharded HyperKZG opening.
On a 1-GPU box the ranks are "virtual" (every slice runs on the same GPU, the combine is the real host code); with
>= 2 GPUs the same checks run as a 2-process NCCL job (scripts/multi_gpu_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cpu as ORC
from tests.util import to_mont_array

pytestmark = pytest.mark.gpu
TAU = 0x1234567890abcdef1122334455667788
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chal(rng, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def test_msm_index_ranges_add_up(ctx):
    import ctypes as C
    from jolt_atlas_b200 import SRS, MultilinearPolynomial, _lib, msm_fr
    from jolt_atlas_b200 import parallel as PAR
    n = 1 << 11
    srs = SRS(ctx, ORC.srs_powers(to_mont_array([TAU])[0], n))
    p = MultilinearPolynomial.random(ctx, n, 5)
    want, winf = msm_fr(ctx, srs, p)
    for world in (2, 3, 8):
        xy = np.zeros((world, 8), dtype=np.uint64)
        inf = np.zeros(world, dtype=np.int32)
        for r in range(world):
            lo, hi = PAR.ShardPlan(r, world).index_range(n)
            one = np.zeros(8, dtype=np.uint64)
            f = C.c_int32()
            _lib.check(ctx._lib.ja_msm_fr_range(ctx._h, srs._h, p._h, lo, hi, one.ctypes.data_as(_lib.u64p), C.byref(f)))
            xy[r], inf[r] = one, f.value
        got, ginf = PAR.g1_sum_affine(xy, inf)
        assert ginf == winf and np.array_equal(got, want)
    p.free(); srs.free()


@pytest.mark.parametrize("kind,npoly,n_out", [(2, 2, 2), (0, 2, 1), (6, 1, 1), (4, 4, 4), (4, 16, 16)])
def test_round_eval_slices_match_unsharded(ctx, kind, npoly, n_out):
    import ctypes as C
    from jolt_atlas_b200 import GruenSplitEqPolynomial, MultilinearPolynomial, _lib, bind_many, round_eval
    from jolt_atlas_b200 import api as A
    from jolt_atlas_b200 import parallel as PAR
    rng = np.random.default_rng(kind * 10 + npoly)
    m, world = 10, 4
    N = 1 << m
    host = rng.integers(0, 1 << 63, size=(npoly, N, 4), dtype=np.uint64)
    host[..., 3] &= np.uint64((1 << 60) - 1)
    full = [MultilinearPolynomial.from_fr(ctx, host[i]) for i in range(npoly)]
    per = N // world
    slices = [[MultilinearPolynomial.from_fr(ctx, host[i, r * per:(r + 1) * per]) for i in range(npoly)] for r in range(world)]
    w = _chal(rng, m)
    eq = GruenSplitEqPolynomial(ctx, w, 0)
    for rnd in range(PAR.ShardPlan(0, world).local_rounds(N)):
        want = round_eval(ctx, kind, full, eq, n_out=n_out)
        parts = np.zeros((world, n_out, 4), dtype=np.uint64)
        for r in range(world):
            n_local = len(slices[r][0])
            arr = (C.c_void_p * npoly)(*[p._h for p in slices[r]])
            _lib.check(ctx._lib.ja_round_eval_slice(ctx._h, kind, arr, npoly, eq._h, 0, r * (n_local // 2), A._u64p(parts[r]), n_out))
        assert np.array_equal(PAR.fr_sum(parts), want), rnd
        ch = _chal(rng, 1)[0]
        eq.bind(ch)
        bind_many(ctx, full, ch, 0)
        for r in range(world):
            bind_many(ctx, slices[r], ch, 0)              # binds are local to a slice
    # after the local rounds the slices are the (world) remaining coefficients of every MLE
    for i in range(npoly):
        rest = np.concatenate([slices[r][i].to_host() for r in range(world)])
        assert np.array_equal(rest, full[i].to_host())
    for p in full + [q for s in slices for q in s]:
        p.free()
    eq.free()


def test_sharded_open_single_rank_equals_all_in_one(ctx):
    from jolt_atlas_b200 import SRS, Blake2bTranscriptState, MultilinearPolynomial, hyperkzg_open
    from jolt_atlas_b200 import parallel as PAR
    ell = 9
    n = 1 << ell
    srs = SRS(ctx, ORC.srs_powers(to_mont_array([TAU])[0], n))
    p = MultilinearPolynomial.random(ctx, n, 3)
    point = _chal(np.random.default_rng(1), ell)
    t1, t2 = Blake2bTranscriptState(b"open"), Blake2bTranscriptState(b"open")
    a = hyperkzg_open(ctx, srs, p, point, t1)
    b = PAR.sharded_hyperkzg_open(ctx, srs, p, point, t2, PAR.LocalComm())
    for k in ("com", "w", "v"):
        assert np.array_equal(a[k], b[k]), k
    assert t1.state == t2.state and t1.n_rounds == t2.n_rounds
    p.free(); srs.free()


def test_two_gpu_nccl_job():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "multi-gpu ok" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("kind,npoly,world,m", [(2, 2, 2, 10), (6, 1, 4, 9), (4, 4, 2, 8), (0, 2, 8, 5), (4, 16, 2, 7)])
def test_sharded_sumcheck_prove_equals_single(kind, npoly, world, m):
    """The full sharded Sumcheck::prove loop (local fused rounds on hypercube slices, partial sums all-gathered through the
    callback, hand-over to replicated rounds when a slice reaches one coefficient) equals the unsharded proof on every
    rank.  Ranks are threads with their own Context on this GPU (tests the engine path; NCCL run: scripts/multi_gpu_check.py)."""
    import threading
    from jolt_atlas_b200 import Blake2bTranscriptState, Context, MultilinearPolynomial, sumcheck_prove
    from jolt_atlas_b200 import parallel as PAR
    rng = np.random.default_rng(kind * 100 + world)
    N = 1 << m
    host = rng.integers(0, 1 << 63, size=(npoly, N, 4), dtype=np.uint64)
    host[..., 3] &= np.uint64((1 << 60) - 1)
    w, claim = _chal(rng, m), _chal(rng, 1)[0]
    with Context(0) as c0:
        t0 = Blake2bTranscriptState(b"shard")
        want = sumcheck_prove(c0, kind, [MultilinearPolynomial.from_fr(c0, host[i]) for i in range(npoly)], claim, t0, eq_w=w)
    group = PAR.ThreadComm.Group(world)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            comm = PAR.ThreadComm(group, rank)
            lo, hi = comm.plan.slice_range(N)
            with Context(0) as c:
                t = Blake2bTranscriptState(b"shard")
                polys = [MultilinearPolynomial.from_fr(c, host[i, lo:hi]) for i in range(npoly)]
                results[rank] = (PAR.sharded_sumcheck_prove(c, kind, polys, claim, t, comm, w), t.state)
        except Exception as e:   # noqa: BLE001
            errors.append((rank, repr(e)))
            try:
                group.barrier.abort()
            except Exception:    # noqa: BLE001
                pass
    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    for rank in range(world):
        got, state = results[rank]
        assert len(got["coeffs"]) == m
        for a, b in zip(got["coeffs"], want["coeffs"]):
            assert np.array_equal(a, b), rank
        assert np.array_equal(got["challenges"], want["challenges"])
        assert np.array_equal(got["final_claims"], want["final_claims"])
        assert state == t0.state


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_prove_pass_equals_single_gpu(world):
    """One microgpt-shaped proof on `world` ranks (threads, one Context each): commitments sharded by polynomial, opening
    MSMs by index range, sumchecks replicated -> every rank produces the single-GPU commitments, opening and transcript."""
    import threading
    import bench
    from jolt_atlas_b200 import SRS, Context
    from jolt_atlas_b200 import parallel as PAR
    from jolt_atlas_b200 import workload as W
    inputs = W.build_inputs("microgpt")
    with Context(0) as c0:
        srs0 = SRS.generate(c0, bench.g1_generator_mont(), bench.tau_mont(), 1 << inputs["ell"]).precompute()
        want = W.run_device(c0, srs0, inputs)
        srs0.free()
    group = PAR.ThreadComm.Group(world)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            comm = PAR.ThreadComm(group, rank)
            with Context(0) as c:
                srs = SRS.generate(c, bench.g1_generator_mont(), bench.tau_mont(), 1 << inputs["ell"]).precompute()
                results[rank] = W.run_device(c, srs, inputs, comm=comm)
                srs.free()
        except Exception as e:   # noqa: BLE001
            errors.append((rank, repr(e)))
            try:
                group.barrier.abort()
            except Exception:    # noqa: BLE001
                pass
    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errors, errors
    for rank in range(world):
        got = results[rank]
        assert got["states"] == want["states"], rank
        assert len(got["commitments"]) == len(want["commitments"])
        for (gxy, ginf), (wxy, winf) in zip(got["commitments"], want["commitments"]):
            assert np.array_equal(gxy, wxy) and np.array_equal(np.asarray(ginf, dtype=bool), np.asarray(winf, dtype=bool)), rank
        for k in ("com", "w", "v"):
            assert np.array_equal(got["open"][k], want["open"][k]), (rank, k)
