"""Multi-GPU layer: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo on CPU) as plumbing.

The reference is single-process rayon (no distributed code to mirror); what shards is dictated by the algorithms:
  * MSM (joltworks/src/msm/mod.rs:27-181, hyperkzg commit/open): split the (scalar, base) pairs by INDEX RANGE, every GPU
    keeps the SRS resident; one all-gather of a partial POINT per GPU and MSM, then world-1 additions.
  * sumcheck fold / round evaluation with LowToHigh binding (dense_mlpoly.rs:219-239): contiguous hypercube slices; binds
    are local, a round exchanges <= 17 partial field sums per GPU.
The exchange is tiny (<= 544 B per GPU and step): it is latency, not bandwidth, so nothing is fused into kernels.
Independent proofs need no exchange at all — that is what `bench.py --gpus N` measures (weak scaling).

Host-side combine goes through the C ABI's GPU-free functions (ja_g1_sum_affine, ja_fr_sum, ja_transcript_*), so the
world_size-2 gloo tests run this logic on CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check


@dataclass(frozen=True)
class ShardPlan:
    """Which part of an index space / hypercube this rank owns."""
    rank: int
    world: int

    def index_range(self, n: int) -> tuple[int, int]:
        """Balanced contiguous range of [0, n) — the MSM split (same rule as the library's ja_set_msm_shard)."""
        return n * self.rank // self.world, n * (self.rank + 1) // self.world

    def slice_range(self, n: int) -> tuple[int, int]:
        """Contiguous hypercube slice of a length-n MLE for LowToHigh binding: n / world coefficients (pairs 2i, 2i+1 stay
        local for log2(n / world) rounds).  n and world must be powers of two with n >= 2 * world."""
        assert n & (n - 1) == 0 and self.world & (self.world - 1) == 0 and n >= 2 * self.world
        per = n // self.world
        return self.rank * per, (self.rank + 1) * per

    def local_rounds(self, n: int) -> int:
        """Rounds a length-n instance can run before its slices shrink to one element per rank."""
        return (n // self.world).bit_length() - 1


class Comm:
    """all_gather of small numpy arrays over torch.distributed (backend nccl -> staged through a CUDA tensor)."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        self.plan = ShardPlan(self.rank, self.world)

    def all_gather(self, a: np.ndarray) -> np.ndarray:
        """(world,) + a.shape, identical on every rank, rank-major."""
        import torch
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy())
        if self.device is not None:
            t = t.to(self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in out])


class LibComm:
    """The library's OWN exchange (csrc/comm.cu): an NCCL communicator owned by the context, collectives on the context's stream,
    no torch tensor and no numpy staging in the data path.  torch.distributed is used once, for the rendezvous (broadcast of the
    128-byte NCCL unique id).  With `shard_on()` every MSM / one-hot commitment of the context is split over the ranks and its
    partial points are combined INSIDE the library call: the ordinary single-GPU API then runs one proof on all GPUs."""
    in_library = True

    def __init__(self, ctx):
        import torch.distributed as dist
        self.ctx = ctx
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.plan = ShardPlan(self.rank, self.world)
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            check(lib.ja_comm_unique_id(buf))
        obj = [buf.raw if self.rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        check(lib.ja_comm_init(ctx._h, self.rank, self.world, obj[0]))

    def all_gather(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a)
        out = np.empty((self.world,) + a.shape, dtype=a.dtype)
        check(self.ctx._lib.ja_comm_allgather(self.ctx._h, a.ctypes.data, a.nbytes, out.ctypes.data))
        return out

    def shard_on(self):
        check(self.ctx._lib.ja_set_msm_shard(self.ctx._h, self.rank, self.world))

    def shard_off(self):
        check(self.ctx._lib.ja_set_msm_shard(self.ctx._h, 0, 1))

    def close(self):
        self.ctx._lib.ja_comm_free(self.ctx._h)


class LocalComm:
    """world == 1 stand-in (and the shape tests use it to run the sharded code path in one process)."""
    rank, world = 0, 1
    plan = ShardPlan(0, 1)

    def all_gather(self, a: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(a)[None]


# ---- host-side combine (GPU-free C ABI) -----------------------------------------------------------------------------
def g1_sum_affine(xy: np.ndarray, is_inf: np.ndarray):
    """Sum of affine points ((n, 8) Montgomery limbs + (n,) infinity flags) -> (xy (8,), is_infinity)."""
    lib = _lib.load()
    xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 8)
    inf = np.ascontiguousarray(is_inf, dtype=np.int32).reshape(-1)
    out = np.zeros(8, dtype=np.uint64)
    oinf = C.c_int32()
    check(lib.ja_g1_sum_affine(xy.ctypes.data_as(_lib.u64p), inf.ctypes.data_as(_lib.i32p), xy.shape[0],
                               out.ctypes.data_as(_lib.u64p), C.byref(oinf)))
    return out, bool(oinf.value)


def fr_sum(parts: np.ndarray) -> np.ndarray:
    """(n_parts, n_vals, 4) partial field sums -> (n_vals, 4)."""
    lib = _lib.load()
    parts = np.ascontiguousarray(parts, dtype=np.uint64)
    n_parts, n_vals = parts.shape[0], parts.shape[1]
    out = np.zeros((n_vals, 4), dtype=np.uint64)
    check(lib.ja_fr_sum(parts.ctypes.data_as(_lib.u64p), n_parts, n_vals, out.ctypes.data_as(_lib.u64p)))
    return out


def combine_points(comm, xy: np.ndarray, inf: np.ndarray):
    """All-gather this rank's partial points ((k, 8), (k,)) and add them point-wise -> ((k, 8), (k,) bool)."""
    xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 8)
    inf = np.ascontiguousarray(inf, dtype=np.int32).reshape(-1)
    packed = np.concatenate([xy, inf.astype(np.uint64)[:, None]], axis=1)          # one exchange for points + flags
    allp = comm.all_gather(packed)                                                 # (world, k, 9)
    out = np.zeros_like(xy)
    oinf = np.zeros(xy.shape[0], dtype=bool)
    for j in range(xy.shape[0]):
        out[j], oinf[j] = g1_sum_affine(allp[:, j, :8], allp[:, j, 8].astype(np.int32))
    return out, oinf


class Transcript:
    """Blake2bTranscript state driven through the library (blake2b.rs) for callers that own the transcript."""

    def __init__(self, label: bytes = None, state: bytes = None, n_rounds: int = 0):
        self._lib = _lib.load()
        self._st = C.create_string_buffer(32)
        self._nr = C.c_uint32(n_rounds)
        if state is not None:
            self._st.raw = state
        else:
            self._lib.ja_transcript_new(label, self._st, C.byref(self._nr))

    @property
    def state(self) -> bytes:
        return self._st.raw

    @property
    def n_rounds(self) -> int:
        return self._nr.value

    def append_points(self, xy, inf):
        xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 8)
        inf = np.ascontiguousarray(inf, dtype=np.int32).reshape(-1)
        self._lib.ja_transcript_append_points(self._st, C.byref(self._nr), xy.ctypes.data_as(_lib.u64p),
                                              inf.ctypes.data_as(_lib.i32p), xy.shape[0])

    def append_scalars(self, fr):
        fr = np.ascontiguousarray(fr, dtype=np.uint64).reshape(-1, 4)
        self._lib.ja_transcript_append_scalars(self._st, C.byref(self._nr), fr.ctypes.data_as(_lib.u64p), fr.shape[0])

    def challenge_scalar(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        self._lib.ja_transcript_challenge_scalar(self._st, C.byref(self._nr), out.ctypes.data_as(_lib.u64p))
        return out

    def challenge_optimized(self, n: int = 1) -> np.ndarray:
        """n x challenge_scalar_optimized (blake2b.rs:233-238): (n, 4) limbs {0, 0, lo, hi}."""
        out = np.zeros((n, 4), dtype=np.uint64)
        self._lib.ja_transcript_challenge_optimized(self._st, C.byref(self._nr), n, out.ctypes.data)
        return out

    def challenge_scalar_powers(self, n: int) -> np.ndarray:
        out = np.zeros((n, 4), dtype=np.uint64)
        self._lib.ja_transcript_challenge_scalar_powers(self._st, C.byref(self._nr), n, out.ctypes.data_as(_lib.u64p))
        return out


def expanding_table(challenges: np.ndarray, order: int = 1) -> np.ndarray:
    """ExpandingTable after len(challenges) updates from [1] (utils/expanding_table.rs:62-89); order 1 = HighToLow."""
    ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((1 << ch.shape[0], 4), dtype=np.uint64)
    check(_lib.load().ja_expanding_table(ch.ctypes.data_as(_lib.u64p), ch.shape[0], order, out.ctypes.data_as(_lib.u64p)))
    return out


# ---- sharded device operations ---------------------------------------------------------------------------------------------
def sharded_msm_fr(ctx, srs, scalars, comm):
    """UnivariateKZG::commit_as_univariate with the pairs split by index range over the ranks.  `scalars` is the FULL
    device polynomial on every rank (replicated streaming data, sharded group arithmetic)."""
    from . import api as A
    lo, hi = comm.plan.index_range(len(scalars))
    out = np.zeros((1, 8), dtype=np.uint64)
    inf = np.zeros(1, dtype=np.int32)
    check(ctx._lib.ja_msm_fr_range(ctx._h, srs._h, scalars._h, lo, hi, A._u64p(out), inf.ctypes.data_as(_lib.i32p)))
    xy, oinf = combine_points(comm, out, inf)
    return xy[0], bool(oinf[0])


def sharded_commit_one_hot_batches(ctx, srs, batches, comm):
    """commit_to_polynomials (prover.rs:236-249) with the address batches dealt round-robin to the ranks: the commitments
    of a proof are independent of each other and of the transcript, so a rank commits batches rank, rank + world, ...
    (all its lists in one pair of launches) and ONE all-gather of the affine points makes every rank hold all of them.
    Same return value as api.commit_one_hot_batches, identical on every rank."""
    from . import api as A
    mine = [b for i, b in enumerate(batches) if i % comm.world == comm.rank]
    res = A.commit_one_hot_batches(ctx, srs, mine) if mine else []
    per_rank = (len(batches) + comm.world - 1) // comm.world
    dmax = max(b.d for b in batches)
    packed = np.zeros((per_rank, dmax, 9), dtype=np.uint64)                # x||y limbs + infinity flag, padded
    for j, (xy, inf) in enumerate(res):
        packed[j, : xy.shape[0], :8] = xy
        packed[j, : xy.shape[0], 8] = np.asarray(inf, dtype=np.uint64)
    allp = comm.all_gather(packed)                                         # (world, per_rank, dmax, 9)
    out = []
    for i, b in enumerate(batches):
        blk = allp[i % comm.world, i // comm.world]
        out.append((np.ascontiguousarray(blk[: b.d, :8]), blk[: b.d, 8].astype(bool)))
    return out


def sharded_hyperkzg_open(ctx, srs, poly, point, transcript_state, comm):
    """HyperKZG::open (hyperkzg/mod.rs:400-447) with every MSM split by index range over the ranks: the folds, the
    univariate evaluations and the quotient recurrence are replicated (HBM streaming, no exchange), the commitments are
    partial points combined by two all-gathers (l-1 points, then 3).  `transcript_state` is api.Blake2bTranscriptState.
    Every rank returns the same proof and ends with the same transcript."""
    from . import api as A
    check(ctx._lib.ja_set_msm_shard(ctx._h, comm.rank, comm.world))
    try:
        op = A.HyperKZGOpening(ctx, srs, poly, point)                                   # phase 1: folds + partial commitments
        com, com_inf = combine_points(comm, op.com, op.com_inf)
        t = Transcript(state=transcript_state.state, n_rounds=transcript_state.n_rounds)
        t.append_points(com, com_inf)                                                   # mod.rs:439
        r = t.challenge_scalar()                                                        # mod.rs:440
        v = op.evals(r)
        t.append_scalars(v.reshape(-1, 4))                                              # mod.rs:258
        q = t.challenge_scalar_powers(op.ell)                                           # mod.rs:260
        w_part, w_inf_part = op.witness(r, q)
        w, w_inf = combine_points(comm, w_part, w_inf_part)
        t.append_points(w, w_inf)                                                       # mod.rs:276
        t.challenge_scalar()                                                            # mod.rs:277
        op.free()
    finally:
        check(ctx._lib.ja_set_msm_shard(ctx._h, 0, 1))
    transcript_state.state, transcript_state.n_rounds = t.state, t.n_rounds
    return {"com": com, "com_inf": com_inf.astype(np.int32), "w": w, "w_inf": w_inf.astype(np.int32), "v": v}


def sharded_round_eval(ctx, kernel_id, slice_polys, eq, comm, n_out, aux_u32=0):
    """One round's reduced sums when every MLE lives as contiguous hypercube slices (this rank holds `slice_polys`,
    `eq` is the replicated split-eq of the whole instance): local partial sums, one all-gather, field addition."""
    from . import api as A
    n_local = len(slice_polys[0])
    g_offset = comm.rank * (n_local // 2)
    arr = (C.c_void_p * len(slice_polys))(*[p._h for p in slice_polys])
    part = np.zeros((n_out, 4), dtype=np.uint64)
    check(ctx._lib.ja_round_eval_slice(ctx._h, kernel_id, arr, len(slice_polys), eq._h, aux_u32, g_offset, A._u64p(part), n_out))
    return fr_sum(comm.all_gather(part))


ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def _allgather_callback(comm):
    """ja_allgather_fn over a Comm: send `nbytes`, receive world x nbytes (rank-major)."""
    def cb(user, send, nbytes, recv):
        try:
            buf = np.frombuffer((C.c_uint8 * nbytes).from_address(send), dtype=np.uint8).copy()
            allp = np.ascontiguousarray(comm.all_gather(buf))
            C.memmove(recv, allp.ctypes.data, comm.world * nbytes)
            return 0
        except Exception:       # noqa: BLE001 - nothing may unwind into the C caller
            return -1
    return ALLGATHER_FN(cb)


def sharded_sumcheck_prove(ctx, kind, slice_polys, claim, transcript, comm, eq_w, aux_u32=0, max_coeffs=40):
    """Sumcheck::prove (sumcheck.rs:565-599) with every MLE sharded into contiguous hypercube slices: `slice_polys` hold
    this rank's len / world coefficients, eq point / claim / transcript are replicated.  Split-eq (LowToHigh) bodies and
    PROD / POW.  Returns the same dict as api.sumcheck_prove, identical on every rank."""
    from . import api as A
    n_local = len(slice_polys[0])
    rounds = (n_local * comm.world).bit_length() - 1
    if getattr(comm, "in_library", False):       # the context's own NCCL communicator: no callback
        cb = None
        check(ctx._lib.ja_set_sumcheck_shard(ctx._h, comm.rank, comm.world, None, None))
    else:
        cb = _allgather_callback(comm)
        check(ctx._lib.ja_set_sumcheck_shard(ctx._h, comm.rank, comm.world, C.cast(cb, C.c_void_p), None))
    try:
        arr = (C.c_void_p * len(slice_polys))(*[p._h for p in slice_polys])
        w = A._fr_arg(eq_w).reshape(-1, 4)
        assert w.shape[0] == rounds
        coeffs = np.zeros((rounds, max_coeffs, 4), dtype=np.uint64)
        ncoeffs = np.zeros(rounds, dtype=np.uint32)
        chal = np.zeros((rounds, 4), dtype=np.uint64)
        fin = np.zeros((len(slice_polys), 4), dtype=np.uint64)
        st = C.create_string_buffer(transcript.state, 32)
        nr = C.c_uint32(transcript.n_rounds)
        check(ctx._lib.ja_sumcheck_prove(ctx._h, kind, arr, len(slice_polys), A._u64p(w), w.shape[0], None, 0, aux_u32,
                                         A._u64p(A._fr_arg(claim)), st, C.byref(nr), max_coeffs, A._u64p(coeffs),
                                         ncoeffs.ctypes.data_as(_lib.u32p), A._u64p(chal), A._u64p(fin)))
    finally:
        check(ctx._lib.ja_set_sumcheck_shard(ctx._h, 0, 1, None, None))
    transcript.state, transcript.n_rounds = st.raw, nr.value
    return {"coeffs": [coeffs[i, : ncoeffs[i]].copy() for i in range(rounds)], "challenges": chal, "final_claims": fin}


class ThreadComm:
    """In-process all-gather between `world` threads (one Context per thread on the same GPU): lets the sharded code
    paths run on a 1-GPU box.  Create one shared ThreadComm.Group and one ThreadComm(rank) per thread."""

    class Group:
        def __init__(self, world):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, group, rank):
        self.group, self.rank, self.world = group, rank, group.world
        self.plan = ShardPlan(rank, group.world)

    def all_gather(self, a: np.ndarray) -> np.ndarray:
        g = self.group
        g.slots[self.rank] = np.ascontiguousarray(a).copy()
        g.barrier.wait(timeout=60)
        out = np.stack(g.slots)
        g.barrier.wait(timeout=60)
        return out
