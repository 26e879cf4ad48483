"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the C++ CPU oracle (oracle/cpp -> oracle/lib/liboracle_cpu.so).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

u64p = C.POINTER(C.c_uint64)
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = build_oracle.LIB
        if not os.path.exists(path):
            path = build_oracle.build()
        _lib = C.CDLL(path)
        _lib.orc_bench_kernel.restype = C.c_double
        _lib.orc_bench_kernel.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_sumcheck_prove.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs ask for every host core explicitly."""
    lib().orc_set_threads(C.c_int(n))


def blake2b256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_blake2b256(data, C.c_size_t(len(data)), out)
    return out.raw


def fr_binop(op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.empty_like(a)
    lib().orc_fr_binop(op, _p(a), _p(b), C.c_size_t(a.shape[0]), _p(out))
    return out


def fr_from_i64(v) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.int64)
    out = np.empty((v.shape[0], 4), dtype=np.uint64)
    lib().orc_fr_from_i64(_p(v), C.c_size_t(v.shape[0]), _p(out))
    return out


def bind(z: np.ndarray, r: np.ndarray, order: int) -> np.ndarray:
    z = np.ascontiguousarray(z, dtype=np.uint64).copy()
    r = np.ascontiguousarray(r, dtype=np.uint64)
    lib().orc_bind(_p(z), C.c_size_t(z.shape[0]), _p(r), order)
    return z[: z.shape[0] // 2].copy()


def eq_evals(r: np.ndarray, scale=None) -> np.ndarray:
    r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((1 << r.shape[0], 4), dtype=np.uint64)
    sc = _p(np.ascontiguousarray(scale, dtype=np.uint64)) if scale is not None else None
    lib().orc_eq_evals(_p(r), C.c_size_t(r.shape[0]), sc, _p(out))
    return out


def evaluate(z: np.ndarray, point: np.ndarray) -> np.ndarray:
    z = np.ascontiguousarray(z, dtype=np.uint64)
    point = np.ascontiguousarray(point, dtype=np.uint64).reshape(-1, 4)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_evaluate(_p(z), C.c_size_t(z.shape[0]), _p(point), C.c_size_t(point.shape[0]), _p(out))
    return out


def sumcheck_prove(family: int, kind: int, polys: np.ndarray, w, claim: np.ndarray, label: bytes, pow_d: int = 0):
    """polys: (npoly, n, 4) Montgomery limbs.  Returns dict(coeffs=[per-round arrays], challenges, final_claims, state)."""
    polys = np.ascontiguousarray(polys, dtype=np.uint64)
    npoly, n, _ = polys.shape
    rounds = n.bit_length() - 1
    w = np.ascontiguousarray(w if w is not None else np.zeros((0, 4)), dtype=np.uint64).reshape(-1, 4)
    claim = np.ascontiguousarray(claim, dtype=np.uint64)
    maxc = 40
    coeffs = np.zeros((rounds, maxc, 4), dtype=np.uint64)
    ncoeffs = np.zeros(rounds, dtype=np.uint32)
    chal = np.zeros((rounds, 4), dtype=np.uint64)
    fin = np.zeros((npoly, 4), dtype=np.uint64)
    state = C.create_string_buffer(32)
    rc = lib().orc_sumcheck_prove(family, kind, C.c_uint(pow_d), _p(polys), C.c_size_t(npoly), C.c_size_t(n), _p(w),
                                  C.c_size_t(w.shape[0]), _p(claim), label, C.c_size_t(maxc), _p(coeffs), _p(ncoeffs),
                                  _p(chal), _p(fin), state)
    assert rc == rounds, rc
    return {"coeffs": [coeffs[i, : ncoeffs[i]].copy() for i in range(rounds)], "challenges": chal,
            "final_claims": fin, "state": state.raw}


def tensor_fold_i32(A: np.ndarray, eq: np.ndarray, transpose: bool) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.int32)
    rows, cols = A.shape
    eq = np.ascontiguousarray(eq, dtype=np.uint64)
    out = np.empty((rows if transpose else cols, 4), dtype=np.uint64)
    lib().orc_tensor_fold_i32(_p(A), C.c_size_t(rows), C.c_size_t(cols), _p(eq), int(transpose), _p(out))
    return out


class TranscriptState:
    """Running Blake2bTranscript state + round counter (blake2b.rs:11-26); new(label) per blake2b.rs:81-100."""

    def __init__(self, label: bytes):
        import hashlib
        self.state = hashlib.blake2b(label + b"\0" * (32 - len(label)), digest_size=32).digest()
        self.n_rounds = 0


def sumcheck_prove_st(family: int, kind: int, polys: np.ndarray, w, claim: np.ndarray, t: TranscriptState, pow_d: int = 0,
                      gammas=None):
    """Like sumcheck_prove but resumes / advances the running transcript `t`.  family 2 = Hamming weight (gammas)."""
    polys = np.ascontiguousarray(polys, dtype=np.uint64)
    npoly, n, _ = polys.shape
    rounds = n.bit_length() - 1
    w = np.ascontiguousarray(w if w is not None else np.zeros((0, 4)), dtype=np.uint64).reshape(-1, 4)
    claim = np.ascontiguousarray(claim, dtype=np.uint64)
    g = _p(np.ascontiguousarray(gammas, dtype=np.uint64)) if gammas is not None else None
    maxc = 40
    coeffs = np.zeros((rounds, maxc, 4), dtype=np.uint64)
    ncoeffs = np.zeros(rounds, dtype=np.uint32)
    chal = np.zeros((rounds, 4), dtype=np.uint64)
    fin = np.zeros((npoly, 4), dtype=np.uint64)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    rc = lib().orc_sumcheck_prove_st(family, kind, C.c_uint(pow_d), _p(polys), C.c_size_t(npoly), C.c_size_t(n), _p(w),
                                     C.c_size_t(w.shape[0]), _p(claim), g, st, C.byref(nr), C.c_size_t(maxc), _p(coeffs),
                                     _p(ncoeffs), _p(chal), _p(fin))
    assert rc == rounds, rc
    t.state, t.n_rounds = st.raw, nr.value
    return {"coeffs": [coeffs[i, : ncoeffs[i]].copy() for i in range(rounds)], "challenges": chal, "final_claims": fin}


def hyperkzg_open_st(srs: np.ndarray, poly: np.ndarray, point: np.ndarray, t: TranscriptState):
    n = poly.shape[0]
    ell = point.shape[0]
    com = np.zeros((max(ell - 1, 1), 8), dtype=np.uint64)
    com_inf = np.zeros(max(ell - 1, 1), dtype=np.int32)
    w = np.zeros((3, 8), dtype=np.uint64)
    w_inf = np.zeros(3, dtype=np.int32)
    v = np.zeros((3, ell, 4), dtype=np.uint64)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    lib().orc_hyperkzg_open_st(_p(srs), C.c_size_t(n), _p(np.ascontiguousarray(poly, dtype=np.uint64)),
                               _p(np.ascontiguousarray(point, dtype=np.uint64)), C.c_size_t(ell), st, C.byref(nr),
                               _p(com), _p(com_inf), _p(w), _p(w_inf), _p(v))
    t.state, t.n_rounds = st.raw, nr.value
    return {"com": com[: ell - 1], "com_inf": com_inf[: ell - 1], "w": w, "w_inf": w_inf, "v": v}


def srs_powers(tau_mont: np.ndarray, n: int) -> np.ndarray:
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_srs_powers(_p(np.ascontiguousarray(tau_mont, dtype=np.uint64)), C.c_size_t(n), _p(out))
    return out


def msm_fr(bases: np.ndarray, scalars: np.ndarray):
    out = np.empty(8, dtype=np.uint64)
    inf = C.c_int32()
    lib().orc_msm_fr(_p(bases), _p(np.ascontiguousarray(scalars, dtype=np.uint64)), C.c_size_t(scalars.shape[0]), _p(out), C.byref(inf))
    return out, bool(inf.value)


def msm_i64(bases: np.ndarray, scalars) -> tuple:
    s = np.ascontiguousarray(scalars, dtype=np.int64)
    out = np.empty(8, dtype=np.uint64)
    inf = C.c_int32()
    lib().orc_msm_i64(_p(bases), _p(s), C.c_size_t(s.shape[0]), _p(out), C.byref(inf))
    return out, bool(inf.value)


def sum_indexed(bases: np.ndarray, idx) -> tuple:
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    inf = C.c_int32()
    lib().orc_sum_indexed(_p(bases), C.c_size_t(bases.shape[0]), _p(idx), C.c_size_t(idx.shape[0]), _p(out), C.byref(inf))
    return out, bool(inf.value)


def hyperkzg_open(srs: np.ndarray, poly: np.ndarray, point: np.ndarray, label: bytes):
    n = poly.shape[0]
    ell = point.shape[0]
    com = np.zeros((max(ell - 1, 1), 8), dtype=np.uint64)
    com_inf = np.zeros(max(ell - 1, 1), dtype=np.int32)
    w = np.zeros((3, 8), dtype=np.uint64)
    w_inf = np.zeros(3, dtype=np.int32)
    v = np.zeros((3, ell, 4), dtype=np.uint64)
    state = C.create_string_buffer(32)
    lib().orc_hyperkzg_open(_p(srs), C.c_size_t(n), _p(np.ascontiguousarray(poly, dtype=np.uint64)),
                            _p(np.ascontiguousarray(point, dtype=np.uint64)), C.c_size_t(ell), label,
                            _p(com), _p(com_inf), _p(w), _p(w_inf), _p(v), state)
    return {"com": com[: ell - 1], "com_inf": com_inf[: ell - 1], "w": w, "w_inf": w_inf, "v": v, "state": state.raw}


def bench_kernel(which: int, log_n: int, iters: int = 3) -> float:
    return lib().orc_bench_kernel(which, log_n, iters)


# ---- BatchedSumcheck::prove over instance descriptors (oracle/cpp/capi.cpp orc_batched_sumcheck_prove) ----
class _OrcInst(C.Structure):
    _fields_ = [("kind", C.c_int32), ("aux_u32", C.c_uint32), ("n_polys", C.c_uint64), ("poly_len", C.c_uint64),
                ("polys", C.c_void_p), ("idx", C.c_void_p), ("eq_w", C.c_void_p), ("eq_m", C.c_uint64),
                ("aux_fr", C.c_void_p), ("n_aux", C.c_uint64), ("claim", C.c_uint64 * 4), ("final_claims", C.c_void_p)]


def batched_sumcheck_prove(instances, t: TranscriptState):
    """instances: list of dicts with keys kind, polys ((n_polys, len, 4) uint64; Booleanity: G tables (d, K, 4)),
    optional eq_w, aux_fr, aux_u32, idx ((d, T) uint32), claim.  Returns dict(coeffs, challenges, final_claims=[per instance])."""
    n = len(instances)
    arr = (_OrcInst * n)()
    keep, finals, max_rounds = [], [], 0
    for i, d in enumerate(instances):
        kind = int(d["kind"])
        polys = np.ascontiguousarray(d["polys"] if kind != 34 else np.zeros((1, 1, 4)), dtype=np.uint64)
        arr[i].kind = kind
        arr[i].aux_u32 = int(d.get("aux_u32", 0))
        arr[i].n_polys, arr[i].poly_len = polys.shape[0], polys.shape[1]
        arr[i].polys = polys.ctypes.data
        keep.append(polys)
        for key, fld in (("eq_w", "eq_w"), ("aux_fr", "aux_fr")):
            v = d.get(key)
            if v is not None:
                v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4)
                keep.append(v)
                setattr(arr[i], fld, v.ctypes.data)
                setattr(arr[i], "eq_m" if key == "eq_w" else "n_aux", v.shape[0])
        if d.get("idx") is not None:
            ix = np.ascontiguousarray(d["idx"], dtype=np.uint32)
            keep.append(ix)
            arr[i].idx = ix.ctypes.data
        claim = np.ascontiguousarray(d.get("claim", np.zeros(4, dtype=np.uint64)), dtype=np.uint64)
        for k in range(4):
            arr[i].claim[k] = int(claim[k])
        fc = np.zeros((polys.shape[0], 4), dtype=np.uint64)
        finals.append(fc)
        arr[i].final_claims = fc.ctypes.data
        if kind == 34:
            arr[i].polys = None
            arr[i].n_polys = 1
        if kind in (32, 34):
            rounds = int(d["aux_u32"]) + arr[i].eq_m
        elif kind == 20:
            rounds = arr[i].eq_m
        elif kind <= 6:
            rounds = arr[i].eq_m
        else:
            rounds = int(polys.shape[1]).bit_length() - 1
        max_rounds = max(max_rounds, rounds)
    maxc = 40
    coeffs = np.zeros((max_rounds, maxc, 4), dtype=np.uint64)
    ncoeffs = np.zeros(max_rounds, dtype=np.uint32)
    chal = np.zeros((max_rounds, 4), dtype=np.uint64)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    fn = lib().orc_batched_sumcheck_prove
    fn.restype = C.c_int
    rc = fn(arr, C.c_size_t(n), st, C.byref(nr), C.c_size_t(maxc), _p(coeffs), _p(ncoeffs), _p(chal))
    assert rc == max_rounds, rc
    t.state, t.n_rounds = st.raw, nr.value
    return {"coeffs": [coeffs[i, : ncoeffs[i]].copy() for i in range(max_rounds)], "challenges": chal, "final_claims": finals}


def compute_ra_evals(idx: np.ndarray, K: int, r_cycle: np.ndarray) -> np.ndarray:
    """shout.rs:549-598.  idx: (d, T) uint32; returns (d, K, 4)."""
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    d, T = idx.shape
    rc = np.ascontiguousarray(r_cycle, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((d, K, 4), dtype=np.uint64)
    lib().orc_compute_ra_evals(_p(idx), C.c_size_t(d), C.c_size_t(T), C.c_size_t(K), _p(rc), C.c_size_t(rc.shape[0]), _p(out))
    return out


def rlc_add_onehot(joint: np.ndarray, idx: np.ndarray, coeffs: np.ndarray):
    """rlc_polynomial.rs:59-74, in place: joint[idx[i][t] * T + t] += coeffs[i]."""
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    co = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    assert joint.flags["C_CONTIGUOUS"] and joint.dtype == np.uint64
    lib().orc_rlc_add_onehot(_p(joint), _p(idx), C.c_size_t(idx.shape[0]), C.c_size_t(idx.shape[1]), _p(co))


def rlc_add_dense(joint: np.ndarray, poly: np.ndarray, coeff: np.ndarray):
    poly = np.ascontiguousarray(poly, dtype=np.uint64)
    lib().orc_rlc_add_dense(_p(joint), _p(poly), C.c_size_t(poly.shape[0]), _p(np.ascontiguousarray(coeff, dtype=np.uint64)))


def transcript_append_scalars(t: TranscriptState, fr: np.ndarray):
    fr = np.ascontiguousarray(fr, dtype=np.uint64).reshape(-1, 4)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    lib().orc_transcript_append_scalars(st, C.byref(nr), _p(fr), C.c_size_t(fr.shape[0]))
    t.state, t.n_rounds = st.raw, nr.value


def transcript_append_scalar_each(t: TranscriptState, fr: np.ndarray):
    """cache_openings: one Transcript::append_scalar per claim (poly/opening_proof.rs:281, :338, :398)."""
    fr = np.ascontiguousarray(fr, dtype=np.uint64).reshape(-1, 4)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    lib().orc_transcript_append_scalar_each(st, C.byref(nr), _p(fr), C.c_size_t(fr.shape[0]))
    t.state, t.n_rounds = st.raw, nr.value


def transcript_challenge_scalar_powers(t: TranscriptState, n: int) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    lib().orc_transcript_challenge_scalar_powers(st, C.byref(nr), C.c_size_t(n), _p(out))
    t.state, t.n_rounds = st.raw, nr.value
    return out


def eval_reduction_h(mle: np.ndarray, points: np.ndarray) -> np.ndarray:
    """compute_h (evaluation_reduction.rs:223-249) by the reference's polynomial-valued fold.  points: (n, m, 4)."""
    mle = np.ascontiguousarray(mle, dtype=np.uint64)
    pts = np.ascontiguousarray(points, dtype=np.uint64)
    n, m = pts.shape[0], pts.shape[1]
    out = np.zeros((m * max(n - 1, 1) + 2, 4), dtype=np.uint64)
    fn = lib().orc_eval_reduction_h
    fn.restype = C.c_int
    k = fn(_p(mle), C.c_size_t(mle.shape[0]), _p(pts), C.c_size_t(n), C.c_size_t(m), _p(out), C.c_size_t(out.shape[0]))
    assert k > 0
    return out[:k]


class PsShout:
    """T-sized passes of the prefix-suffix Shout prover (oracle/cpp/psshout.hpp; ps_shout/mod.rs:269-335,420-446)."""

    def __init__(self, lookup_indices, r_cycle, log_k: int = 64, phases: int = 8):
        idx = np.ascontiguousarray(lookup_indices, dtype=np.uint64)
        r = np.ascontiguousarray(r_cycle, dtype=np.uint64).reshape(-1, 4)
        self.T, self.phases, self.m = idx.shape[0], phases, 1 << (log_k // phases)
        fn = lib().orc_psshout_new
        fn.restype = C.c_void_p
        self._h = C.c_void_p(fn(_p(idx), C.c_size_t(idx.shape[0]), _p(r), C.c_size_t(r.shape[0]), C.c_uint(log_k), C.c_uint(phases)))

    def init_phase(self, phase: int, v_prev, suffix_kinds, bound: int) -> np.ndarray:
        kinds = np.ascontiguousarray(suffix_kinds, dtype=np.uint32)
        out = np.empty((kinds.shape[0], self.m, 4), dtype=np.uint64)
        v = _p(np.ascontiguousarray(v_prev, dtype=np.uint64)) if v_prev is not None else None
        lib().orc_psshout_init_phase(self._h, C.c_uint(phase), v, _p(kinds), C.c_size_t(kinds.shape[0]), C.c_uint(bound), _p(out))
        return out

    def materialize_ra(self, v) -> np.ndarray:
        out = np.empty((self.T, 4), dtype=np.uint64)
        lib().orc_psshout_materialize_ra(self._h, _p(np.ascontiguousarray(v, dtype=np.uint64)), _p(out))
        return out

    def prove_address(self, t: "TranscriptState", gamma, claim, bound: int) -> dict:
        """The LOG_K address rounds of the read-raf sumcheck (ps_shout/mod.rs:337-418, :491-560 under sumcheck.rs:565-599), on a
        FRESH state (phase 0 not yet initialised).  -> coeffs (rounds, 2, 4) compressed [c0, c2], ncoeffs, challenges, v, val,
        raf_val, claim (the running claim handed to the cycle rounds)."""
        log_k = self.phases * (self.m.bit_length() - 1)
        out = dict(coeffs=np.zeros((log_k, 2, 4), dtype=np.uint64), ncoeffs=np.zeros(log_k, dtype=np.uint32),
                   challenges=np.zeros((log_k, 4), dtype=np.uint64), v=np.zeros((self.phases, self.m, 4), dtype=np.uint64),
                   val=np.zeros(4, dtype=np.uint64), raf_val=np.zeros(4, dtype=np.uint64), claim=np.zeros(4, dtype=np.uint64))
        st = C.create_string_buffer(t.state, 32)
        nr = C.c_uint32(t.n_rounds)
        lib().orc_psshout_prove_address(self._h, C.c_uint(bound), _p(np.ascontiguousarray(gamma, dtype=np.uint64)),
                                        _p(np.ascontiguousarray(claim, dtype=np.uint64)) if claim is not None else None, st, C.byref(nr), _p(out["coeffs"]),
                                        _p(out["ncoeffs"]), _p(out["challenges"]), _p(out["v"]), _p(out["val"]), _p(out["raf_val"]),
                                        _p(out["claim"]))
        t.state, t.n_rounds = st.raw, nr.value
        return out

    def prove_identity_rc(self, t: "TranscriptState", claim) -> dict:
        """IdentityRCProver's LOG_K address rounds (identity_range_check.rs:140-325) on a fresh state; claim None = derived."""
        log_k = self.phases * (self.m.bit_length() - 1)
        out = dict(coeffs=np.zeros((log_k, 2, 4), dtype=np.uint64), ncoeffs=np.zeros(log_k, dtype=np.uint32),
                   challenges=np.zeros((log_k, 4), dtype=np.uint64), v=np.zeros((self.phases, self.m, 4), dtype=np.uint64),
                   raf_val=np.zeros(4, dtype=np.uint64), claim=np.zeros(4, dtype=np.uint64))
        st = C.create_string_buffer(t.state, 32)
        nr = C.c_uint32(t.n_rounds)
        lib().orc_psshout_prove_identity_rc(self._h, _p(np.ascontiguousarray(claim, dtype=np.uint64)) if claim is not None else None, st,
                                            C.byref(nr), _p(out["coeffs"]), _p(out["ncoeffs"]), _p(out["challenges"]), _p(out["v"]),
                                            _p(out["raf_val"]), _p(out["claim"]))
        t.state, t.n_rounds = st.raw, nr.value
        return out

    def free(self):
        if self._h:
            lib().orc_psshout_free(self._h)
            self._h = None


def clamp_evaluate_mle(r, xlen: int, bound: int) -> np.ndarray:
    """ClampBoundedTable<XLEN, BOUND, true>::evaluate_mle (lookup_tables/clamp.rs:140-192); r = xlen Montgomery coordinates, MSB first."""
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_clamp_evaluate_mle(_p(np.ascontiguousarray(r, dtype=np.uint64)), C.c_uint(xlen), C.c_uint(bound), _p(out))
    return out


def signed_identity_evaluate(r, n: int) -> np.ndarray:
    """SignedIdentityPoly::evaluate (poly/signed_identity_poly.rs:43-59)."""
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_signed_identity_evaluate(_p(np.ascontiguousarray(r, dtype=np.uint64)), C.c_uint(n), _p(out))
    return out


def suffix_mle(kind: int, bits: int, length: int, xlen: int, bound: int) -> int:
    fn = lib().orc_suffix_mle
    fn.restype = C.c_uint64
    return int(fn(C.c_int(kind), C.c_uint64(bits), C.c_uint(length), C.c_uint(xlen), C.c_uint(bound)))


def transcript_challenge_optimized(t: TranscriptState, n: int) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    st = C.create_string_buffer(t.state, 32)
    nr = C.c_uint32(t.n_rounds)
    lib().orc_transcript_challenge_optimized(st, C.byref(nr), C.c_size_t(n), _p(out))
    t.state, t.n_rounds = st.raw, nr.value
    return out


def expanding_table_h2l(challenges: np.ndarray) -> np.ndarray:
    ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)
    out = np.empty((1 << ch.shape[0], 4), dtype=np.uint64)
    lib().orc_expanding_table_h2l(_p(ch), C.c_size_t(ch.shape[0]), _p(out))
    return out
