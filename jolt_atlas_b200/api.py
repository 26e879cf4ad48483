"""Host-side mirror of the reference's polynomial interface on top of the C ABI.

Names and argument meaning follow the reference (joltworks/src/poly/*.rs) so parity tests read like
the reference's own tests:
  MultilinearPolynomial.{from_fr, from_i32, bind_parallel, final_claim, evaluate, len}
        multilinear_polynomial.rs:22-35, :657-667, :728-762, :766-862
  EqPolynomial.evals                      eq_poly.rs:77-101
  GruenSplitEqPolynomial.{new, bind, merge, get_current_scalar, get_current_w}   split_eq_poly.rs:86-504
  BindingOrder                            multilinear_polynomial.rs (LowToHigh / HighToLow)

Field elements cross this layer as numpy uint64 arrays of shape (..., 4): little-endian limbs in
Montgomery form, exactly ark_bn254::Fr's memory layout; challenges are {0, 0, lo, hi}.
Errors from the C ABI raise JoltAtlasError (the reference prover panics at the same points).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import JoltAtlasError, check  # noqa: F401


class BindingOrder:
    LowToHigh = 0
    HighToLow = 1


class EvalKernel:
    ADD, SUB, MUL, SQUARE, PROD, POW, IDENT = 0, 1, 2, 3, 4, 5, 6
    IFF, DIV, RSQRT, LIN3 = 8, 9, 10, 11           # ops/iff.rs:189, ops/div.rs:329, ops/rsqrt.rs:390, neural_teleport/division.rs:231
    WIDENT = 12                                    # softmax_last_axis/recip_mult.rs:196 (phase 1); ja_round_eval only
    DOT2, DOT3, SUM1, SUMHI, OPEN = 16, 17, 18, 19, 20
    # two-phase / table-weighted bodies (ja_round_eval only; the table or eq polynomial is the last polynomial, aux_u32 = shift)
    WSUM, WDOT2, DOT2_L2H, SQ_EQHI, DOT2_EQHI, DOT2_EQLOW = 21, 22, 23, 24, 25, 26
    N_OUT = {0: 1, 1: 1, 2: 2, 3: 2, 6: 1, 8: 2, 9: 2, 10: 2, 11: 1, 12: 1, 16: 2, 17: 3, 18: 1, 19: 1,
             21: 1, 22: 3, 23: 2, 24: 3, 25: 3, 26: 3}
    FAMILY_S = (0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12)


class _Addr(C.c_void_p):
    """Address of an ndarray's buffer that keeps the array alive for the duration of the call it is passed to."""
    __slots__ = ("_keep",)


def _u64p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    p = _Addr(a.ctypes.data)
    p._keep = a
    return p


def _fr_arg(x) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


class Context:
    """Owns one CUDA device context/stream of the library (ja_init / ja_shutdown)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.ja_init(device, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            self._lib.ja_shutdown(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        check(self._lib.ja_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.ja_launch_count(self._h))

    def timer_begin(self):
        check(self._lib.ja_timer_begin(self._h))

    def timer_end(self) -> float:
        out = C.c_float()
        check(self._lib.ja_timer_end(self._h, C.byref(out)))
        return out.value

    def profile_begin(self):
        check(self._lib.ja_profile_begin(self._h))

    def profile_end(self) -> dict:
        """{class name: {"launches": n, "ms": summed device ms}} for the launches since profile_begin."""
        k = int(self._lib.ja_profile_class_count())
        cnt = np.zeros(k, dtype=np.uint64)
        ms = np.zeros(k, dtype=np.float64)
        check(self._lib.ja_profile_end(self._h, _u64p(cnt), ms.ctypes.data_as(C.POINTER(C.c_double)), k))
        return {self._lib.ja_profile_class_name(i).decode(): {"launches": int(cnt[i]), "ms": float(ms[i])}
                for i in range(k) if cnt[i]}

    def bench_kernel(self, which: int, log_n: int, n_polys: int = 1, iters: int = 20) -> float:
        """Average device milliseconds per launch of one kernel on resident synthetic operands."""
        out = C.c_float()
        check(self._lib.ja_bench_kernel(self._h, which, log_n, n_polys, iters, C.byref(out)))
        return out.value

    def bench_fused(self, which: int, log_n: int, iters: int = 10) -> float:
        """Average device milliseconds per launch of one fused bind+eval round kernel (ja_bench_fused)."""
        out = C.c_float()
        check(self._lib.ja_bench_fused(self._h, which, log_n, iters, C.byref(out)))
        return out.value

    def calibrate_fr_mul(self, iters: int = 2000) -> float:
        out = C.c_double()
        check(self._lib.ja_calibrate_fr_mul(self._h, iters, C.byref(out)))
        return out.value

    def test_field_ops(self, op: int, a, b) -> np.ndarray:
        """ja_test_field_ops: the device field arithmetic element-wise on host arrays (n, 4) of Montgomery limbs."""
        a, b = _fr_arg(a), _fr_arg(b)
        assert a.shape == b.shape
        out = np.empty_like(a)
        check(self._lib.ja_test_field_ops(self._h, op, _u64p(a), _u64p(b), a.shape[0], _u64p(out)))
        return out


class MultilinearPolynomial:
    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    # -- constructors
    @staticmethod
    def from_fr(ctx: Context, z) -> "MultilinearPolynomial":
        z = _fr_arg(z)
        h = C.c_void_p()
        check(ctx._lib.ja_poly_from_fr(ctx._h, _u64p(z), z.shape[0], C.byref(h)))
        return MultilinearPolynomial(ctx, h)

    @staticmethod
    def from_i32(ctx: Context, z) -> "MultilinearPolynomial":
        z = np.ascontiguousarray(z, dtype=np.int32)
        h = C.c_void_p()
        check(ctx._lib.ja_poly_from_i32(ctx._h, z.ctypes.data_as(_lib.i32p), z.shape[0], C.byref(h)))
        return MultilinearPolynomial(ctx, h)

    @staticmethod
    def from_i32_many(ctx: Context, mat) -> list:
        """One polynomial per row of an (npoly, n) int32 matrix: one copy and one synchronisation for all of them."""
        mat = np.ascontiguousarray(mat, dtype=np.int32)
        hs = (C.c_void_p * mat.shape[0])()
        check(ctx._lib.ja_poly_from_i32_many(ctx._h, mat.ctypes.data_as(_lib.i32p), mat.shape[0], mat.shape[1], hs))
        return [MultilinearPolynomial(ctx, C.c_void_p(h)) for h in hs]

    @staticmethod
    def from_lookup(ctx: Context, table, idx) -> "MultilinearPolynomial":
        """RaPolynomial materialisation: out[t] = table[idx[t]] (idx 0xFFFFFFFF = None -> 0)."""
        table = _fr_arg(table).reshape(-1, 4)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        h = C.c_void_p()
        check(ctx._lib.ja_poly_from_lookup(ctx._h, _u64p(table), table.shape[0], idx.ctypes.data_as(_lib.u32p), idx.shape[0],
                                           C.byref(h)))
        return MultilinearPolynomial(ctx, h)

    @staticmethod
    def random(ctx: Context, n: int, seed: int = 1) -> "MultilinearPolynomial":
        """Device-generated pseudo-random canonical Fr coefficients (synthetic bench operands)."""
        h = C.c_void_p()
        check(ctx._lib.ja_poly_random(ctx._h, n, seed, C.byref(h)))
        return MultilinearPolynomial(ctx, h)

    @staticmethod
    def zeros(ctx: Context, n: int) -> "MultilinearPolynomial":
        h = C.c_void_p()
        check(ctx._lib.ja_poly_zeros(ctx._h, n, C.byref(h)))
        return MultilinearPolynomial(ctx, h)

    def rlc_add_onehot(self, addr, coeffs):
        """build_materialized_rlc, sparse half: self[k_i[t] * T + t] += coeffs[i] for the batch's d one-hot polynomials."""
        co = np.ascontiguousarray(_fr_arg(coeffs).reshape(-1, 4))
        assert co.shape[0] == addr.d
        check(self.ctx._lib.ja_rlc_add_onehot(self.ctx._h, self._h, addr._h, _u64p(co)))

    def rlc_add_dense(self, poly, coeff):
        """build_materialized_rlc, dense half: self[i] += coeff * poly[i]."""
        check(self.ctx._lib.ja_rlc_add_dense(self.ctx._h, self._h, poly._h, _u64p(np.ascontiguousarray(_fr_arg(coeff)))))

    def clone(self) -> "MultilinearPolynomial":
        h = C.c_void_p()
        check(self.ctx._lib.ja_poly_clone(self.ctx._h, self._h, C.byref(h)))
        return MultilinearPolynomial(self.ctx, h)

    def free(self):
        if self._h:
            self.ctx._lib.ja_poly_free(self.ctx._h, self._h)
            self._h = None

    @staticmethod
    def free_many(polys):
        """Release several polynomials of one context in one call (ja_poly_free_many)."""
        live = [p for p in polys if p._h]
        if not live:
            return
        arr = (C.c_void_p * len(live))(*[p._h for p in live])
        live[0].ctx._lib.ja_poly_free_many(live[0].ctx._h, arr, len(live))
        for p in live:
            p._h = None

    def __len__(self):
        return int(self.ctx._lib.ja_poly_len(self._h))

    def to_host(self) -> np.ndarray:
        n = len(self)
        out = np.empty((n, 4), dtype=np.uint64)
        check(self.ctx._lib.ja_poly_to_host(self.ctx._h, self._h, _u64p(out), n))
        return out

    # -- PolynomialBinding
    def bind_parallel(self, r, order: int):
        r = _fr_arg(r)
        check(self.ctx._lib.ja_bind(self.ctx._h, self._h, _u64p(r), order))

    bind = bind_parallel

    def final_claim(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        check(self.ctx._lib.ja_final_claim(self.ctx._h, self._h, _u64p(out)))
        return out

    # -- PolynomialEvaluation
    def evaluate(self, point) -> np.ndarray:
        point = _fr_arg(point).reshape(-1, 4)
        out = np.empty(4, dtype=np.uint64)
        check(self.ctx._lib.ja_poly_evaluate(self.ctx._h, self._h, _u64p(point), point.shape[0], _u64p(out)))
        return out


def bind_many(ctx: Context, polys, r, order: int):
    """One `ingest_challenge`: bind several polynomials with the same challenge in one launch."""
    r = _fr_arg(r)
    arr = (C.c_void_p * len(polys))(*[p._h for p in polys])
    check(ctx._lib.ja_bind_many(ctx._h, arr, len(polys), _u64p(r), order))


class EqPolynomial:
    @staticmethod
    def evals(ctx: Context, r, scaling=None) -> MultilinearPolynomial:
        r = _fr_arg(r).reshape(-1, 4)
        h = C.c_void_p()
        sc = _u64p(_fr_arg(scaling)) if scaling is not None else None
        check(ctx._lib.ja_eq_evals(ctx._h, _u64p(r) if r.shape[0] else None, r.shape[0], sc, C.byref(h)))
        return MultilinearPolynomial(ctx, h)


class GruenSplitEqPolynomial:
    def __init__(self, ctx: Context, w, order: int, scaling=None):
        w = _fr_arg(w).reshape(-1, 4)
        self.ctx = ctx
        h = C.c_void_p()
        sc = _u64p(_fr_arg(scaling)) if scaling is not None else None
        check(ctx._lib.ja_spliteq_new(ctx._h, _u64p(w) if w.shape[0] else None, w.shape[0], order, sc, C.byref(h)))
        self._h = h

    new = classmethod(lambda cls, ctx, w, order: cls(ctx, w, order))

    def bind(self, r):
        check(self.ctx._lib.ja_spliteq_bind(self.ctx._h, self._h, _u64p(_fr_arg(r))))

    def get_current_scalar(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        check(self.ctx._lib.ja_spliteq_current_scalar(self._h, _u64p(out)))
        return out

    def get_current_w(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        check(self.ctx._lib.ja_spliteq_current_w(self._h, _u64p(out)))
        return out

    def merge(self) -> MultilinearPolynomial:
        h = C.c_void_p()
        check(self.ctx._lib.ja_spliteq_merge(self.ctx._h, self._h, C.byref(h)))
        return MultilinearPolynomial(self.ctx, h)

    def free(self):
        if self._h:
            self.ctx._lib.ja_spliteq_free(self.ctx._h, self._h)
            self._h = None


def round_eval(ctx: Context, kernel_id: int, polys, eq: GruenSplitEqPolynomial | None = None,
               aux_fr=None, aux_u32: int = 0, n_out: int | None = None) -> np.ndarray:
    """Reduced sums of one `compute_message` body (see EvalKernel / include/jolt_atlas_b200.h)."""
    if n_out is None:
        n_out = EvalKernel.N_OUT[kernel_id]
    arr = (C.c_void_p * len(polys))(*[p._h for p in polys])
    out = np.empty((n_out, 4), dtype=np.uint64)
    aux = _fr_arg(aux_fr).reshape(-1, 4) if aux_fr is not None else None
    check(ctx._lib.ja_round_eval(ctx._h, kernel_id, arr, len(polys), eq._h if eq is not None else None,
                                 _u64p(aux) if aux is not None else None, aux.shape[0] if aux is not None else 0,
                                 aux_u32, _u64p(out), n_out))
    return out


def eval_reduction_h(ctx: Context, mle: MultilinearPolynomial, points) -> np.ndarray:
    """compute_h (evaluation_reduction.rs:223-249): coefficients of h = mle o l for the curve l through `points` (n, m, 4)."""
    pts = np.ascontiguousarray(_fr_arg(points))
    n, m = pts.shape[0], pts.shape[1]
    out = np.zeros((m * (n - 1) + 1, 4), dtype=np.uint64)
    cnt = C.c_size_t()
    check(ctx._lib.ja_eval_reduction_h(ctx._h, mle._h, _u64p(pts), n, m, _u64p(out), C.byref(cnt)))
    return out[: cnt.value]


def tensor_fold_i32(ctx: Context, A, eq: MultilinearPolynomial, transpose: bool) -> MultilinearPolynomial:
    A = np.ascontiguousarray(A, dtype=np.int32)
    rows, cols = A.shape
    h = C.c_void_p()
    check(ctx._lib.ja_tensor_fold_i32(ctx._h, A.ctypes.data_as(_lib.i32p), rows, cols, eq._h, int(transpose), C.byref(h)))
    return MultilinearPolynomial(ctx, h)


class TensorI32:
    """An i32 tensor (rows x cols) resident on the device: fold it with eq tables any number of times."""

    def __init__(self, ctx: Context, A):
        A = np.ascontiguousarray(A, dtype=np.int32)
        assert A.ndim == 2
        self.ctx, self.shape = ctx, A.shape
        h = C.c_void_p()
        check(ctx._lib.ja_tensor_i32_upload(ctx._h, A.ctypes.data_as(_lib.i32p), A.shape[0], A.shape[1], C.byref(h)))
        self._h = h

    def fold(self, eq: MultilinearPolynomial, transpose: bool) -> MultilinearPolynomial:
        h = C.c_void_p()
        check(self.ctx._lib.ja_tensor_fold_resident(self.ctx._h, self._h, eq._h, int(transpose), C.byref(h)))
        return MultilinearPolynomial(self.ctx, h)

    def free(self):
        if self._h:
            self.ctx._lib.ja_tensor_i32_free(self.ctx._h, self._h)
            self._h = None


# ---- commitment half: SRS residency, MSM, one-hot point sums (joltworks/src/msm/mod.rs, hyperkzg/) ----
class MsmWidth:
    FR, U8, U16, U32, U64, I32, I64 = range(7)
    DTYPE = {1: np.uint8, 2: np.uint16, 3: np.uint32, 4: np.uint64, 5: np.int32, 6: np.int64}


class SRS:
    """Device-resident g1_powers of a KZGProverKey (kzg.rs:108-143): (n, 8) uint64 affine x||y Montgomery limbs."""

    def __init__(self, ctx: Context, g1_affine_xy):
        pts = np.ascontiguousarray(g1_affine_xy, dtype=np.uint64).reshape(-1, 8)
        self.ctx = ctx
        h = C.c_void_p()
        check(ctx._lib.ja_srs_upload(ctx._h, _u64p(pts), pts.shape[0], C.byref(h)))
        self._h = h

    @classmethod
    def generate(cls, ctx: Context, g1_xy, beta, n: int) -> "SRS":
        """SRS::setup's fixed-base loop on device: g1_powers[i] = beta^i * g1."""
        self = cls.__new__(cls)
        self.ctx = ctx
        h = C.c_void_p()
        g = np.ascontiguousarray(g1_xy, dtype=np.uint64).reshape(8)
        check(ctx._lib.ja_srs_generate(ctx._h, _u64p(g), _u64p(_fr_arg(beta)), n, C.byref(h)))
        self._h = h
        return self

    def precompute(self) -> "SRS":
        """Build the fixed-base window table (16x the SRS in HBM): later Fr MSMs use one bucket set and no doubling tail."""
        check(self.ctx._lib.ja_srs_precompute(self.ctx._h, self._h))
        return self

    def to_host(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = len(self) - first if count is None else count
        out = np.empty((count, 8), dtype=np.uint64)
        check(self.ctx._lib.ja_srs_to_host(self.ctx._h, self._h, first, count, _u64p(out)))
        return out

    def __len__(self):
        return int(self.ctx._lib.ja_srs_len(self._h))

    def free(self):
        if self._h:
            self.ctx._lib.ja_srs_free(self.ctx._h, self._h)
            self._h = None


def _pt_out(count: int):
    return np.zeros((count, 8), dtype=np.uint64), np.zeros(count, dtype=np.int32)


def msm_fr(ctx: Context, srs: SRS, scalars: MultilinearPolynomial):
    """UnivariateKZG::commit_as_univariate on a device polynomial -> (xy limbs, is_infinity)."""
    out, inf = _pt_out(1)
    check(ctx._lib.ja_msm_fr(ctx._h, srs._h, scalars._h, _u64p(out), inf.ctypes.data_as(_lib.i32p)))
    return out[0], bool(inf[0])


def msm_fr_batch(ctx: Context, srs: SRS, polys):
    """UnivariateKZG::commit_variable_batch: one bucket pipeline for all polynomials."""
    out, inf = _pt_out(max(len(polys), 1))
    arr = (C.c_void_p * max(len(polys), 1))(*[p._h for p in polys])
    check(ctx._lib.ja_msm_fr_batch(ctx._h, srs._h, arr, len(polys), _u64p(out), inf.ctypes.data_as(_lib.i32p)))
    return out[: len(polys)], inf[: len(polys)].astype(bool)


def msm_host(ctx: Context, srs: SRS, scalars, width: int, base_offset: int = 0):
    """VariableBaseMSM::msm over host scalars of the given width tag (MsmWidth)."""
    if width == MsmWidth.FR:
        s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    else:
        s = np.ascontiguousarray(scalars, dtype=MsmWidth.DTYPE[width]).reshape(-1)
    out, inf = _pt_out(1)
    check(ctx._lib.ja_msm_host(ctx._h, srs._h, base_offset, s.ctypes.data_as(C.c_void_p), width, s.shape[0],
                               _u64p(out), inf.ctypes.data_as(_lib.i32p)))
    return out[0], bool(inf[0])


def g1_sum_indexed(ctx: Context, srs: SRS, indices):
    """HyperKZG::commit_one_hot: sum of g1_powers[indices]."""
    idx = np.ascontiguousarray(indices, dtype=np.uint64).reshape(-1)
    out, inf = _pt_out(1)
    check(ctx._lib.ja_g1_sum_indexed(ctx._h, srs._h, _u64p(idx) if idx.shape[0] else None, idx.shape[0], _u64p(out),
                                     inf.ctypes.data_as(_lib.i32p)))
    return out[0], bool(inf[0])


def g1_sum_indexed_batch(ctx: Context, srs: SRS, index_lists):
    """HyperKZG::batch_commit_one_hot."""
    offs = np.zeros(len(index_lists) + 1, dtype=np.uint64)
    for i, l in enumerate(index_lists):
        offs[i + 1] = offs[i] + len(l)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.uint64) for l in index_lists])
                                if index_lists else np.zeros(0), dtype=np.uint64)
    out, inf = _pt_out(max(len(index_lists), 1))
    check(ctx._lib.ja_g1_sum_indexed_batch(ctx._h, srs._h, _u64p(flat) if flat.shape[0] else None, _u64p(offs),
                                           len(index_lists), _u64p(out), inf.ctypes.data_as(_lib.i32p)))
    return out[: len(index_lists)], inf[: len(index_lists)].astype(bool)


class OneHotBatch:
    """Device-resident batch of one-hot index lists (k*T + t per non-None entry, hyperkzg/mod.rs:536-542)."""

    def __init__(self, ctx: Context, index_lists):
        self.ctx, self.count = ctx, len(index_lists)
        offs = np.zeros(self.count + 1, dtype=np.uint64)
        for i, l in enumerate(index_lists):
            offs[i + 1] = offs[i] + len(l)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.uint64) for l in index_lists]), dtype=np.uint64)
        h = C.c_void_p()
        check(ctx._lib.ja_onehot_upload(ctx._h, _u64p(flat) if flat.shape[0] else None, _u64p(offs), self.count, C.byref(h)))
        self._h = h

    def commit(self, srs: SRS):
        """HyperKZG::batch_commit_one_hot over the resident lists."""
        out, inf = _pt_out(self.count)
        check(self.ctx._lib.ja_onehot_commit(self.ctx._h, srs._h, self._h, _u64p(out), inf.ctypes.data_as(_lib.i32p)))
        return out, inf.astype(bool)

    def free(self):
        if self._h:
            self.ctx._lib.ja_onehot_free(self.ctx._h, self._h)
            self._h = None


# ---- HyperKZG::open (hyperkzg/mod.rs:400-447) ----
class Blake2bTranscriptState:
    """Running state + round counter of a Blake2bTranscript (transcripts/blake2b.rs:11-26), as the library's
    all-in-one entry points exchange it.  new(label) follows blake2b.rs:81-100."""

    def __init__(self, label: bytes):
        import hashlib
        assert len(label) <= 32
        self.state = hashlib.blake2b(label + b"\0" * (32 - len(label)), digest_size=32).digest()
        self.n_rounds = 0


def hyperkzg_open(ctx: Context, srs: SRS, poly: MultilinearPolynomial, point, transcript: Blake2bTranscriptState):
    """HyperKZG::open -> dict(com, com_inf, w, w_inf, v); advances `transcript` exactly as the reference does."""
    point = _fr_arg(point).reshape(-1, 4)
    ell = point.shape[0]
    com, com_inf = _pt_out(max(ell - 1, 1))
    w, w_inf = _pt_out(3)
    v = np.zeros((3, ell, 4), dtype=np.uint64)
    st = C.create_string_buffer(transcript.state, 32)
    nr = C.c_uint32(transcript.n_rounds)
    check(ctx._lib.ja_hyperkzg_open(ctx._h, srs._h, poly._h, _u64p(point), ell, st, C.byref(nr), _u64p(com),
                                    com_inf.ctypes.data_as(_lib.i32p), _u64p(w), w_inf.ctypes.data_as(_lib.i32p), _u64p(v)))
    transcript.state = st.raw
    transcript.n_rounds = nr.value
    return {"com": com[: ell - 1], "com_inf": com_inf[: ell - 1], "w": w, "w_inf": w_inf, "v": v}


class HyperKZGOpening:
    """The split form of HyperKZG::open for callers that own the transcript (the Rust shim's shape)."""

    def __init__(self, ctx: Context, srs: SRS, poly: MultilinearPolynomial, point):
        point = _fr_arg(point).reshape(-1, 4)
        self.ctx, self.ell = ctx, point.shape[0]
        self.com, self.com_inf = _pt_out(max(self.ell - 1, 1))
        h = C.c_void_p()
        check(ctx._lib.ja_hyperkzg_open_begin(ctx._h, srs._h, poly._h, _u64p(point), self.ell, C.byref(h),
                                              _u64p(self.com), self.com_inf.ctypes.data_as(_lib.i32p)))
        self._h = h
        self.com, self.com_inf = self.com[: self.ell - 1], self.com_inf[: self.ell - 1]

    def evals(self, r) -> np.ndarray:
        v = np.zeros((3, self.ell, 4), dtype=np.uint64)
        check(self.ctx._lib.ja_hyperkzg_open_evals(self.ctx._h, self._h, _u64p(_fr_arg(r)), _u64p(v)))
        return v

    def witness(self, r, q_powers):
        w, w_inf = _pt_out(3)
        q = _fr_arg(q_powers).reshape(-1, 4)
        check(self.ctx._lib.ja_hyperkzg_open_witness(self.ctx._h, self._h, _u64p(_fr_arg(r)), _u64p(q), _u64p(w),
                                                     w_inf.ctypes.data_as(_lib.i32p)))
        return w, w_inf

    def free(self):
        if self._h:
            self.ctx._lib.ja_hyperkzg_open_free(self.ctx._h, self._h)
            self._h = None


# ---- Sumcheck::prove (subprotocols/sumcheck.rs:565-599) ----
class _ScResult(dict):
    """Result of a sumcheck call; the per-round coefficient views (`coeffs`) are built on first access (a proof makes
    hundreds of calls and only tests / serialisation look at them)."""

    def __missing__(self, key):
        if key == "coeffs":
            buf, nc = self["_coeffs"], self["_ncoeffs"]
            v = [buf[i, :n] for i, n in enumerate(nc)]
            self["coeffs"] = v
            return v
        raise KeyError(key)


def sumcheck_prove(ctx: Context, kind: int, polys, claim, transcript: Blake2bTranscriptState, eq_w=None, gammas=None,
                   pow_d: int = 0, max_coeffs: int = 40):
    """One instance through the device kernels + the library transcript.  Consumes `polys`.
    Returns dict(coeffs=[per-round compressed coefficient arrays], challenges, final_claims)."""
    n = len(polys[0])
    rounds = n.bit_length() - 1
    arr = (C.c_void_p * len(polys))(*[p._h for p in polys])
    w = _fr_arg(eq_w).reshape(-1, 4) if eq_w is not None else None
    g = _fr_arg(gammas).reshape(-1, 4) if gammas is not None else None
    coeffs = np.empty((rounds, max_coeffs, 4), dtype=np.uint64)       # the library writes ncoeffs[i] entries of row i
    ncoeffs = np.zeros(rounds, dtype=np.uint32)
    chal = np.zeros((rounds, 4), dtype=np.uint64)
    fin = np.zeros((len(polys), 4), dtype=np.uint64)
    st = C.create_string_buffer(transcript.state, 32)
    nr = C.c_uint32(transcript.n_rounds)
    check(ctx._lib.ja_sumcheck_prove(ctx._h, kind, arr, len(polys), _u64p(w) if w is not None else None,
                                     w.shape[0] if w is not None else 0, _u64p(g) if g is not None else None,
                                     g.shape[0] if g is not None else 0, pow_d, _u64p(_fr_arg(claim)), st, C.byref(nr),
                                     max_coeffs, _u64p(coeffs), ncoeffs.ctypes.data_as(_lib.u32p), _u64p(chal), _u64p(fin)))
    transcript.state = st.raw
    transcript.n_rounds = nr.value
    nc = ncoeffs.tolist()
    return _ScResult(_coeffs=coeffs, _ncoeffs=nc, challenges=chal, final_claims=fin, msg_bytes=32 * sum(nc))


# ---- one-hot address batches (witness.rs:84-99 -> OneHotPolynomial; hyperkzg/mod.rs:558-596; shout.rs:549-598) ----
class OneHotAddresses:
    """The d chunk-address lists of a node, resident on the device: (d, T) uint32 in [0, K), 0xFFFFFFFF = None."""

    def __init__(self, ctx: Context, k, K: int):
        k = np.ascontiguousarray(k, dtype=np.uint32)
        assert k.ndim == 2
        self.ctx, self.d, self.T, self.K = ctx, k.shape[0], k.shape[1], K
        h = C.c_void_p()
        check(ctx._lib.ja_addr_upload(ctx._h, k.ctypes.data_as(_lib.u32p), self.d, self.T, K, C.byref(h)))
        self._h = h

    @classmethod
    def _view(cls, ctx: Context, handle, d: int, T: int, K: int):
        """A batch owned by another object (FusedWitness): free() is a no-op."""
        o = cls.__new__(cls)
        o.ctx, o.d, o.T, o.K, o._h, o._borrowed = ctx, d, T, K, C.c_void_p(handle), True
        return o

    @classmethod
    def upload_many(cls, ctx: Context, ks, K: int):
        """All address batches of a proof in one call (ja_addr_upload_many): copies and device-side validation enqueued back
        to back, one synchronisation.  ks: list of (d_i, T_i) uint32 arrays -> list of OneHotAddresses."""
        ks = [np.ascontiguousarray(k, dtype=np.uint32) for k in ks]
        n = len(ks)
        assert n and all(k.ndim == 2 for k in ks)
        ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in ks])
        ds = (C.c_size_t * n)(*[k.shape[0] for k in ks])
        ts = (C.c_size_t * n)(*[k.shape[1] for k in ks])
        kk = (C.c_size_t * n)(*([K] * n))
        outs = (C.c_void_p * n)()
        check(ctx._lib.ja_addr_upload_many(ctx._h, ptrs, ds, ts, kk, n, outs))
        res = []
        for i, k in enumerate(ks):
            o = cls.__new__(cls)
            o.ctx, o.d, o.T, o.K, o._h = ctx, k.shape[0], k.shape[1], K, C.c_void_p(outs[i])
            res.append(o)
        return res

    def commit(self, srs: SRS):
        """HyperKZG::batch_commit_one_hot over the resident lists -> ((d, 8) xy limbs, (d,) is_infinity)."""
        out, inf = _pt_out(self.d)
        check(self.ctx._lib.ja_addr_commit(self.ctx._h, srs._h, self._h, _u64p(out), inf.ctypes.data_as(_lib.i32p)))
        return out, inf.astype(bool)

    def gather(self, tables):
        """RaPolynomial::new(indices, table_i) for every list in one launch: tables (d, K, 4) -> d polynomials."""
        tables = np.ascontiguousarray(tables, dtype=np.uint64).reshape(self.d, self.K, 4)
        arr = (C.c_void_p * self.d)()
        check(self.ctx._lib.ja_addr_gather(self.ctx._h, self._h, _u64p(tables), arr))
        return [MultilinearPolynomial(self.ctx, C.c_void_p(arr[i])) for i in range(self.d)]

    def ra_evals(self, r_cycle) -> np.ndarray:
        """compute_ra_evals: G[i][k] = sum_{t: k_i[t] == k} eq(r_cycle, t) -> (d, K, 4)."""
        r = _fr_arg(r_cycle).reshape(-1, 4)
        out = np.zeros((self.d, self.K, 4), dtype=np.uint64)
        check(self.ctx._lib.ja_addr_ra_evals(self.ctx._h, self._h, _u64p(r), r.shape[0], _u64p(out)))
        return out

    def free(self):
        if self._h and not getattr(self, "_borrowed", False):
            self.ctx._lib.ja_addr_free(self.ctx._h, self._h)
        self._h = None


def commit_one_hot_batches(ctx: Context, srs: SRS, batches):
    """commit_to_polynomials over every one-hot polynomial of a proof: all address batches in one pair of launches.
    Returns [(xy (d_i, 8), is_infinity (d_i,)) per batch]."""
    total = sum(b.d for b in batches)
    out, inf = _pt_out(total)
    arr = (C.c_void_p * len(batches))(*[b._h for b in batches])
    check(ctx._lib.ja_addr_commit_many(ctx._h, srs._h, arr, len(batches), _u64p(out), inf.ctypes.data_as(_lib.i32p)))
    res, o = [], 0
    for b in batches:
        res.append((out[o:o + b.d], inf[o:o + b.d].astype(bool)))
        o += b.d
    return res


def fr_add(a, b) -> np.ndarray:
    """(a + b) mod r on Montgomery limbs (addition does not depend on the representation): host glue between two library calls."""
    R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    to_i = lambda x: sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(x, dtype=np.uint64).reshape(4)))
    s_ = (to_i(a) + to_i(b)) % R_MOD
    return np.array([(s_ >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fr_div(a, b) -> np.ndarray:
    """a / b on Montgomery limbs (a_m * b_m^-1 * R mod r): host glue between two library calls."""
    R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    to_i = lambda x: sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(x, dtype=np.uint64).reshape(4)))
    q = to_i(a) * pow(to_i(b), -1, R_MOD) % R_MOD * pow(2, 256, R_MOD) % R_MOD
    return np.array([(q >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def transcript_append_scalar_each(ctx: Context, transcript: "Blake2bTranscriptState", fr):
    """cache_openings of one instance: one Transcript::append_scalar per claim (poly/opening_proof.rs:281, :338, :398)."""
    fr = _fr_arg(fr).reshape(-1, 4)
    st = C.create_string_buffer(transcript.state, 32)
    nr = C.c_uint32(transcript.n_rounds)
    ctx._lib.ja_transcript_append_scalar_each(st, C.byref(nr), _u64p(fr), fr.shape[0])
    transcript.state, transcript.n_rounds = st.raw, nr.value


class SuffixKind:
    """Suffix MLEs of the clamp-table family (joltworks/src/lookup_tables/suffixes/) and the identity suffix of the raf decomposition."""
    ONE, HIGHER_ALL_ZERO, HZERO_MUL_LWORD, HONE_MUL_LWORD, IDENTITY, SHIFT = 0, 1, 2, 3, 4, 5


class PrefixSuffixShout:
    """The T-sized passes of ReadRafSumcheckProver (joltworks/src/subprotocols/ps_shout/mod.rs): lookup indices and u_evals resident on
    the device; init_phase(phase, v_prev) -> the m-entry suffix polynomials Q of the phase; materialize_ra(v) -> the cycle-round polynomial."""

    def __init__(self, ctx: Context, lookup_indices, r_cycle, log_k: int = 64, phases: int = 8):
        idx = np.ascontiguousarray(lookup_indices, dtype=np.uint64)
        r = _fr_arg(r_cycle).reshape(-1, 4)
        self.ctx, self.T, self.log_k, self.phases, self.m = ctx, idx.shape[0], log_k, phases, 1 << (log_k // phases)
        h = C.c_void_p()
        check(ctx._lib.ja_psshout_new(ctx._h, _u64p(idx), idx.shape[0], _u64p(r), r.shape[0], log_k, phases, C.byref(h)))
        self._h = h

    def init_phase(self, phase: int, v_prev, suffix_kinds, bound: int) -> np.ndarray:
        kinds = np.ascontiguousarray(suffix_kinds, dtype=np.uint32)
        out = np.empty((kinds.shape[0], self.m, 4), dtype=np.uint64)
        v = _fr_arg(v_prev).reshape(-1, 4) if v_prev is not None else None
        check(self.ctx._lib.ja_psshout_init_phase(self.ctx._h, self._h, phase, _u64p(v) if v is not None else None,
                                                  kinds.ctypes.data_as(_lib.u32p), kinds.shape[0], bound, _u64p(out)))
        return out

    def materialize_ra(self, v=None, scale=None) -> "MultilinearPolynomial":
        """v = None: the expanding tables of the last prove_address.  scale: the constant val + raf_val of the cycle rounds, folded into ra."""
        v = _fr_arg(v).reshape(self.phases * self.m, 4) if v is not None else None
        sc = _fr_arg(scale).reshape(4) if scale is not None else None
        h = C.c_void_p()
        check(self.ctx._lib.ja_psshout_materialize_ra(self.ctx._h, self._h, _u64p(v) if v is not None else None,
                                                      _u64p(sc) if sc is not None else None, C.byref(h)))
        return MultilinearPolynomial(self.ctx, h)

    def prove_address(self, transcript: "Blake2bTranscriptState", gamma, bound: int, claim=None) -> dict:
        """The LOG_K address rounds of the read-raf sumcheck (ps_shout/mod.rs:337-418, :491-560 under sumcheck.rs:565-599): the phase
        passes on the device, the 256-entry rounds on the host next to the library transcript.  claim = None: the prover's own sum."""
        n = self.log_k
        out = dict(coeffs=np.zeros((n, 2, 4), dtype=np.uint64), ncoeffs=np.zeros(n, dtype=np.uint32),
                   challenges=np.zeros((n, 4), dtype=np.uint64), input_claim=np.zeros(4, dtype=np.uint64),
                   val=np.zeros(4, dtype=np.uint64), raf_val=np.zeros(4, dtype=np.uint64), claim=np.zeros(4, dtype=np.uint64))
        st = C.create_string_buffer(transcript.state, 32)
        nr = C.c_uint32(transcript.n_rounds)
        g = _fr_arg(gamma).reshape(4)
        cl = _fr_arg(claim).reshape(4) if claim is not None else None
        check(self.ctx._lib.ja_psshout_prove_address(self.ctx._h, self._h, bound, _u64p(g), _u64p(cl) if cl is not None else None, st,
                                                     C.byref(nr), _u64p(out["coeffs"]), out["ncoeffs"].ctypes.data_as(_lib.u32p),
                                                     _u64p(out["challenges"]), _u64p(out["input_claim"]), _u64p(out["val"]),
                                                     _u64p(out["raf_val"]), _u64p(out["claim"])))
        transcript.state, transcript.n_rounds = st.raw, nr.value
        out["msg_bytes"] = 32 * int(out["ncoeffs"].sum())
        return out

    def prove_identity_rc(self, transcript: "Blake2bTranscriptState", claim=None) -> dict:
        """IdentityRCProver's LOG_K address rounds (identity_range_check.rs:140-325): the remainder range check."""
        n = self.log_k
        out = dict(coeffs=np.zeros((n, 2, 4), dtype=np.uint64), ncoeffs=np.zeros(n, dtype=np.uint32),
                   challenges=np.zeros((n, 4), dtype=np.uint64), input_claim=np.zeros(4, dtype=np.uint64),
                   raf_val=np.zeros(4, dtype=np.uint64), claim=np.zeros(4, dtype=np.uint64))
        st = C.create_string_buffer(transcript.state, 32)
        nr = C.c_uint32(transcript.n_rounds)
        cl = _fr_arg(claim).reshape(4) if claim is not None else None
        check(self.ctx._lib.ja_psshout_prove_identity_rc(self.ctx._h, self._h, _u64p(cl) if cl is not None else None, st, C.byref(nr),
                                                         _u64p(out["coeffs"]), out["ncoeffs"].ctypes.data_as(_lib.u32p),
                                                         _u64p(out["challenges"]), _u64p(out["input_claim"]), _u64p(out["raf_val"]),
                                                         _u64p(out["claim"])))
        transcript.state, transcript.n_rounds = st.raw, nr.value
        out["msg_bytes"] = 32 * int(out["ncoeffs"].sum())
        return out

    def tables(self) -> np.ndarray:
        v = np.zeros((self.phases, self.m, 4), dtype=np.uint64)
        check(self.ctx._lib.ja_psshout_tables(self.ctx._h, self._h, _u64p(v)))
        return v

    def free(self):
        if self._h:
            self.ctx._lib.ja_psshout_free(self.ctx._h, self._h)
            self._h = None


class FusedWitness:
    """generate_node_witnesses (jolt-atlas-core/src/onnx_proof/witness.rs:142-214) for a fused node, on the device: from the resident
    i32 operands to the ClampRaD / RescaleRemainderRaD address batches, the clamp lookup indices and the clamped output.
    op: 0 einsum mk,kn->mn, 1 Mul, 2 Add, 3 Sub.  `.clamp` / `.rem` are OneHotAddresses views owned by the witness."""
    EINSUM_MK_KN, MUL, ADD, SUB = 0, 1, 2, 3

    def __init__(self, ctx: Context, op: int, A: "TensorI32", B: "TensorI32", scale_bits: int, T: int):
        self.ctx, self.T, self.scale_bits = ctx, T, scale_bits
        h = C.c_void_p()
        check(ctx._lib.ja_witness_fused(ctx._h, op, A._h, B._h, scale_bits, T, C.byref(h)))
        self._h = h
        self.clamp = OneHotAddresses._view(ctx, ctx._lib.ja_witness_clamp_addr(h), 16, T, 16)
        d_rem = (scale_bits + 3) // 4
        self.rem = OneHotAddresses._view(ctx, ctx._lib.ja_witness_rem_addr(h), d_rem, T, 16) if d_rem else None

    def ps_shout(self, r_cycle, log_k: int = 64, phases: int = 8) -> "PrefixSuffixShout":
        r = _fr_arg(r_cycle).reshape(-1, 4)
        ps = PrefixSuffixShout.__new__(PrefixSuffixShout)
        ps.ctx, ps.T, ps.log_k, ps.phases, ps.m = self.ctx, self.T, log_k, phases, 1 << (log_k // phases)
        h = C.c_void_p()
        check(self.ctx._lib.ja_psshout_from_witness(self.ctx._h, self._h, _u64p(r), r.shape[0], log_k, phases, C.byref(h)))
        ps._h = h
        return ps

    def rem_shout(self, r_cycle, phases: int) -> "PrefixSuffixShout":
        """ps_shout state of the remainder range check (LOG_K = the rescale bits)."""
        r = _fr_arg(r_cycle).reshape(-1, 4)
        ps = PrefixSuffixShout.__new__(PrefixSuffixShout)
        ps.ctx, ps.T, ps.log_k, ps.phases, ps.m = self.ctx, self.T, self.scale_bits, phases, 1 << (self.scale_bits // phases)
        h = C.c_void_p()
        check(self.ctx._lib.ja_psshout_from_witness_rem(self.ctx._h, self._h, _u64p(r), r.shape[0], phases, C.byref(h)))
        ps._h = h
        return ps

    def to_host(self):
        """(lookup indices (T,) u64, clamped output (T,) i32, clamp chunks (16, T) u32, remainder chunks (d_rem, T) u32 or None)."""
        idx = np.empty(self.T, dtype=np.uint64)
        o = np.empty(self.T, dtype=np.int32)
        ck = np.empty((16, self.T), dtype=np.uint32)
        rk = np.empty((self.rem.d, self.T), dtype=np.uint32) if self.rem is not None else None
        check(self.ctx._lib.ja_witness_to_host(self.ctx._h, self._h, idx.ctypes.data, o.ctypes.data, ck.ctypes.data,
                                               rk.ctypes.data if rk is not None else None))
        return idx, o, ck, rk

    def free(self):
        if self._h:
            self.ctx._lib.ja_witness_free(self.ctx._h, self._h)
            self._h = None


class InstanceKind:
    BOOLEANITY, HAMMING_TABLES, OPENING_ONEHOT = 32, 33, 34


def batched_sumcheck_prove(ctx: Context, instances, transcript: Blake2bTranscriptState, max_coeffs: int = 40):
    """BatchedSumcheck::prove (sumcheck.rs:30-184).  `instances`: list of dicts
         device kinds   {"kind": EvalKernel.*, "polys": [MultilinearPolynomial...], "eq_w", "aux_fr", "aux_u32", "claim"}
         BOOLEANITY     {"kind": 32, "tables": G (d, K, 4), "addr": OneHotAddresses, "eq_w": r_cycle, "gammas", "r_address"}
         HAMMING_TABLES {"kind": 33, "tables": G (d, K, 4), "aux_fr": gamma powers, "claim"}
         OPENING_ONEHOT {"kind": 34, "addr": OneHotAddresses (d lists), "eq_w": r_cycle, "r_address", "claims": (d, 4)}
                        expands to d instances; its final_claims entry is (d, 4)
    Device polynomials are consumed.  Returns dict(coeffs, challenges, final_claims=[per instance (n, 4)])."""
    n = len(instances)
    arr = (_lib.ScInstance * n)()
    keep, finals, max_rounds = [], [], 0
    zero_claim = np.zeros(4, dtype=np.uint64)
    for i, d in enumerate(instances):
        kind = int(d["kind"])
        ai = arr[i]                                   # one struct view per instance (every arr[i] builds a new one)
        ai.kind = kind
        ai.aux_u32 = int(d.get("aux_u32", 0))
        if kind == InstanceKind.OPENING_ONEHOT:
            d = dict(d)
            d["tables"] = _fr_arg(d["claims"]).reshape(-1, 1, 4)
            d["aux_fr"] = d["r_address"]
        if kind in (InstanceKind.BOOLEANITY, InstanceKind.HAMMING_TABLES, InstanceKind.OPENING_ONEHOT):
            tabs = np.ascontiguousarray(d["tables"], dtype=np.uint64)
            keep.append(tabs)
            ai.n_polys, ai.table_len = tabs.shape[0], tabs.shape[1]
            ai.host_tables = tabs.ctypes.data
            npoly = tabs.shape[0]
            rounds = tabs.shape[1].bit_length() - 1
        else:
            polys = d["polys"]
            hs = (C.c_void_p * len(polys))(*[p._h for p in polys])
            keep.append(hs)
            ai.n_polys = len(polys)
            ai.polys = C.cast(hs, C.c_void_p)
            npoly = len(polys)
            rounds = len(polys[0]).bit_length() - 1
        aux = d.get("aux_fr")
        if kind == InstanceKind.OPENING_ONEHOT:
            ai.addr = d["addr"]._h
            rounds = _fr_arg(aux).reshape(-1, 4).shape[0] + _fr_arg(d["eq_w"]).reshape(-1, 4).shape[0]
        if kind == InstanceKind.BOOLEANITY:
            ai.addr = d["addr"]._h
            ra = _fr_arg(d["r_address"]).reshape(-1, 4)
            aux = np.concatenate([_fr_arg(d["gammas"]).reshape(-1, 4), ra])
            ai.aux_u32 = ra.shape[0]
            rounds = ra.shape[0] + _fr_arg(d["eq_w"]).reshape(-1, 4).shape[0]
        for key, val in (("eq_w", d.get("eq_w")), ("aux_fr", aux)):
            if val is not None:
                v = np.ascontiguousarray(_fr_arg(val).reshape(-1, 4))
                keep.append(v)
                setattr(ai, key, v.ctypes.data)
                setattr(ai, "eq_m" if key == "eq_w" else "n_aux", v.shape[0])
        claim = _fr_arg(d.get("claim", zero_claim))
        C.memmove(ai.claim, claim.ctypes.data, 32)
        fc = np.zeros((npoly, 4), dtype=np.uint64)
        finals.append(fc)
        ai.out_final_claims = fc.ctypes.data
        max_rounds = max(max_rounds, rounds)
    coeffs = np.empty((max_rounds, max_coeffs, 4), dtype=np.uint64)   # the library writes ncoeffs[i] entries of row i
    ncoeffs = np.zeros(max_rounds, dtype=np.uint32)
    chal = np.zeros((max_rounds, 4), dtype=np.uint64)
    st = C.create_string_buffer(transcript.state, 32)
    nr = C.c_uint32(transcript.n_rounds)
    check(ctx._lib.ja_batched_sumcheck_prove(ctx._h, arr, n, st, C.byref(nr), max_coeffs, _u64p(coeffs),
                                             ncoeffs.ctypes.data_as(_lib.u32p), _u64p(chal)))
    transcript.state = st.raw
    transcript.n_rounds = nr.value
    nc = ncoeffs.tolist()
    return _ScResult(_coeffs=coeffs, _ncoeffs=nc, challenges=chal, final_claims=finals, msg_bytes=32 * sum(nc))
