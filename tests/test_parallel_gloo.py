"""CPU-only: the N>1 host logic (shard plan, exchange, combine, caller-owned transcript) on world_size 2 with the gloo
backend.  Local partial results come from the oracle; the combine goes through the product library's GPU-free functions."""
import os
import random
import socket

import numpy as np
import pytest

from oracle.pyref import field as F
from tests.util import to_mont_array

TAU = 0x1234567890abcdef1122334455667788


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jolt_atlas_b200 import parallel as PAR
        from oracle import cpu as ORC
        from oracle.pyref import poly as PL
        comm = PAR.Comm()
        assert (comm.rank, comm.world) == (rank, world)
        # -- MSM split by index range: partial points from the oracle, all-gather, host-side addition
        n = 300                                            # not a multiple of the world size on purpose
        srs = ORC.srs_powers(to_mont_array([TAU])[0], n)
        rng = random.Random(11)
        scal = to_mont_array([rng.randrange(F.P) for _ in range(n)])
        lo, hi = comm.plan.index_range(n)
        part, pinf = ORC.msm_fr(srs[lo:hi], scal[lo:hi])
        got, ginf = PAR.combine_points(comm, part[None], np.array([int(pinf)]))
        want, winf = ORC.msm_fr(srs, scal)
        assert not winf and not ginf[0] and np.array_equal(got[0], want)
        # a partial that is the identity on one rank (empty contribution) still combines
        z = np.zeros((1, 8), dtype=np.uint64)
        got, ginf = PAR.combine_points(comm, part[None] if rank == 0 else z, np.array([int(pinf) if rank == 0 else 1]))
        w0, _ = ORC.msm_fr(srs[:n * 1 // world], scal[:n * 1 // world])
        assert np.array_equal(got[0], w0)
        # -- one split-eq round evaluation over contiguous hypercube slices (Mul body), partial sums added as field elements
        m = 6
        N = 1 << m
        w = [F.challenge_to_fr(rng.getrandbits(128) & F.CHALLENGE_MASK) for _ in range(m)]
        a = [rng.randrange(F.P) for _ in range(N)]
        b = [rng.randrange(F.P) for _ in range(N)]
        eq = PL.GruenSplitEq(w, 0)
        body = lambda g: [a[2 * g] * b[2 * g] % F.P, (a[2 * g + 1] - a[2 * g]) * (b[2 * g + 1] - b[2 * g]) % F.P]
        full = eq.fold(body, 2)
        s_lo, s_hi = comm.plan.slice_range(N)
        e_out, e_in = eq.E_out(), eq.E_in()
        bits_in = len(e_in).bit_length() - 1
        part = [0, 0]
        for g in range(s_lo // 2, s_hi // 2):
            wgt = e_out[g >> bits_in] * e_in[g & ((1 << bits_in) - 1)] % F.P
            v = body(g)
            part = [(part[k] + wgt * v[k]) % F.P for k in range(2)]
        tot = PAR.fr_sum(comm.all_gather(to_mont_array(part)))
        from tests.util import from_mont_array
        assert from_mont_array(tot) == [x % F.P for x in full]
        assert comm.plan.local_rounds(N) == m - 1
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_msm_and_round_combine():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_plan_partitions():
    from jolt_atlas_b200.parallel import ShardPlan
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 300, 1 << 18):
            rs = [ShardPlan(r, world).index_range(n) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    for world in (1, 2, 4, 8):
        n = 1 << 10
        rs = [ShardPlan(r, world).slice_range(n) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(hi - lo == n // world for lo, hi in rs)


def test_caller_owned_transcript_matches_oracle():
    """ja_transcript_* (the library's Blake2b transcript for callers that own the state) against the Python twin."""
    from jolt_atlas_b200.parallel import Transcript
    from oracle.pyref import curve as CV
    from oracle.pyref import transcript as TR
    rng = random.Random(3)
    t = Transcript(b"HyperKZG")
    o = TR.Blake2bTranscript(b"HyperKZG")
    assert t.state == o.state
    pts = [CV.scalar_mul(CV.G1, rng.randrange(1, F.P)) for _ in range(3)] + [None]
    xy = np.zeros((4, 8), dtype=np.uint64)
    inf = np.zeros(4, dtype=np.int32)
    for i, p in enumerate(pts):
        if p is None:
            inf[i] = 1
        else:
            xy[i, :4] = F.fq_to_mont(p[0]); xy[i, 4:] = F.fq_to_mont(p[1])
    t.append_points(xy, inf); o.append_points(pts)
    assert t.state == o.state
    r = t.challenge_scalar()
    assert F.fr_from_mont([int(x) for x in r]) == o.challenge_scalar()
    sc = [rng.randrange(F.P) for _ in range(5)]
    t.append_scalars(to_mont_array(sc)); o.append_scalars(sc)
    qp = t.challenge_scalar_powers(4)
    assert [F.fr_from_mont([int(x) for x in row]) for row in qp] == o.challenge_scalar_powers(4)
    assert t.state == o.state and t.n_rounds == o.n_rounds
