"""GPU parity tests (through the C ABI) for the polynomial layer: bind, eq tables, split-eq, MLE
evaluation, tensor folds, round evaluation.  Oracle = oracle/pyref (Python big-int restatement of
joltworks/src/poly/*).  Bar: bit-exact canonical Montgomery limbs.
Mirrors the reference's own invariants: eq serial==parallel==cached (eq_poly.rs:265-313), split-eq
merge == dense eq after every bind in both orders (split_eq_poly.rs:623-669), F*Challenge ==
F*Fr(challenge) (transcripts/blake2b.rs:286-316)."""
import random

import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from oracle.pyref import poly as PL
from tests.util import challenge_array, from_mont_array, rand_challenge, rand_fr, to_mont_array

pytestmark = pytest.mark.gpu

P = F.P
EDGE = [0, 1, P - 1, F.R % P, (P - 1) // 2, 2, P - 2, (1 << 253) % P, F.R2 % P]


def _poly(ctx, vals):
    from jolt_atlas_b200 import MultilinearPolynomial
    return MultilinearPolynomial.from_fr(ctx, to_mont_array(vals))


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("logn", [1, 2, 5, 10, 13])
def test_bind_matches_oracle(ctx, order, logn):
    rng = random.Random(100 * order + logn)
    n = 1 << logn
    z = rand_fr(rng, n)
    for i, e in enumerate(EDGE[: min(n, len(EDGE))]):
        z[i] = e
    p = _poly(ctx, z)
    cur = z
    for rnd in range(logn):
        c = rand_challenge(rng) if rnd else F.CHALLENGE_MASK      # max challenge first
        p.bind_parallel(challenge_array(c), order)
        cur = PL.bind(cur, F.challenge_to_fr(c), order)
        assert len(p) == len(cur)
        if rnd < 3 or len(cur) <= 4:
            assert from_mont_array(p.to_host()) == cur
    assert from_mont_array(p.final_claim()) == cur
    p.free()


def test_bind_edge_challenges(ctx):
    rng = random.Random(7)
    z = rand_fr(rng, 64)
    for c in [0, 1, F.CHALLENGE_MASK, 1 << 124, (1 << 64) - 1, 1 << 64]:
        for order in (0, 1):
            p = _poly(ctx, z)
            p.bind_parallel(challenge_array(c), order)
            assert from_mont_array(p.to_host()) == PL.bind(z, F.challenge_to_fr(c), order)
            p.free()


def test_bind_errors(ctx):
    from jolt_atlas_b200 import JoltAtlasError, MultilinearPolynomial
    p = _poly(ctx, [5])
    with pytest.raises(JoltAtlasError):
        p.bind_parallel(challenge_array(3), 0)        # already fully bound
    with pytest.raises(JoltAtlasError):
        MultilinearPolynomial.from_fr(ctx, to_mont_array([1, 2, 3]))   # not a power of two
    q = _poly(ctx, [1, 2])
    with pytest.raises(JoltAtlasError):
        q.final_claim()                                # len != 1
    with pytest.raises(JoltAtlasError):
        q.bind_parallel(np.array([1, 0, 0, 0], dtype=np.uint64), 0)   # not a MontU128Challenge
    p.free(); q.free()


def test_from_i32_matches_field_embedding(ctx):
    from jolt_atlas_b200 import MultilinearPolynomial
    vals = [0, 1, -1, 127, -128, 2**31 - 1, -2**31, 65535, 65536, -65535, -65536, 12345, -54321, 2, -2, 7]
    p = MultilinearPolynomial.from_i32(ctx, np.array(vals, dtype=np.int32))
    assert from_mont_array(p.to_host()) == [v % P for v in vals]
    p.free()


@pytest.mark.parametrize("m", [0, 1, 2, 3, 7, 12, 15])
def test_eq_evals(ctx, m):
    from jolt_atlas_b200 import EqPolynomial
    rng = random.Random(m)
    cs = [rand_challenge(rng) for _ in range(m)]
    r_fr = [F.challenge_to_fr(c) for c in cs]
    r_arr = np.array([F.challenge_limbs(c) for c in cs], dtype=np.uint64).reshape(-1, 4)
    t = EqPolynomial.evals(ctx, r_arr)
    assert from_mont_array(t.to_host()) == PL.eq_evals(r_fr)
    t.free()
    # full-width Fr point with scaling
    pt = rand_fr(rng, m)
    sc = rng.randrange(P)
    t = EqPolynomial.evals(ctx, to_mont_array(pt).reshape(-1, 4), to_mont_array([sc])[0])
    assert from_mont_array(t.to_host()) == PL.eq_evals(pt, sc)
    t.free()


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("m", [1, 2, 5, 8, 11])
def test_spliteq_merge_after_every_bind(ctx, order, m):
    from jolt_atlas_b200 import GruenSplitEqPolynomial
    rng = random.Random(31 * m + order)
    w = rand_fr(rng, m)
    ref = PL.GruenSplitEq(w, order)
    g = GruenSplitEqPolynomial(ctx, to_mont_array(w), order)
    for rnd in range(m):
        mg = g.merge()
        assert from_mont_array(mg.to_host()) == ref.merge()
        mg.free()
        assert from_mont_array(g.get_current_w()) == [ref.current_w()]
        c = rand_challenge(rng)
        g.bind(challenge_array(c))
        ref.bind(F.challenge_to_fr(c))
        assert from_mont_array(g.get_current_scalar()) == [ref.current_scalar]
    g.free()


@pytest.mark.parametrize("m", [0, 1, 4, 9, 12])
def test_evaluate(ctx, m):
    rng = random.Random(m + 5)
    z = rand_fr(rng, 1 << m)
    cs = [rand_challenge(rng) for _ in range(m)]
    p = _poly(ctx, z)
    pt = np.array([F.challenge_limbs(c) for c in cs], dtype=np.uint64).reshape(-1, 4)
    got = from_mont_array(p.evaluate(pt))
    assert got == [PL.evaluate(z, [F.challenge_to_fr(c) for c in cs])]
    p.free()


@pytest.mark.parametrize("kind,kid,npoly", [("add", 0, 2), ("sub", 1, 2), ("mul", 2, 2), ("square", 3, 1), ("ident", 6, 1)])
@pytest.mark.parametrize("m", [1, 2, 3, 6, 11])
def test_round_eval_split_eq_every_round(ctx, kind, kid, npoly, m):
    """GPU sums == oracle par_fold_out_in for every round of a LowToHigh sumcheck."""
    from jolt_atlas_b200 import GruenSplitEqPolynomial, bind_many, round_eval
    rng = random.Random(m * 17 + kid)
    w = [F.challenge_to_fr(rand_challenge(rng)) for _ in range(m)]
    zs = [[rng.randrange(-2**31, 2**31) % P for _ in range(1 << m)] for _ in range(npoly)]
    if m >= 3:
        zs[0][:4] = [0, P - 1, 1, 0]
    ref_eq = PL.GruenSplitEq(w, 0)
    g = GruenSplitEqPolynomial(ctx, to_mont_array(w), 0)
    polys = [_poly(ctx, z) for z in zs]
    cur = [list(z) for z in zs]
    for rnd in range(m):
        def body(gi):
            a0 = cur[0][2 * gi]
            if kind == "add":
                return [a0 + cur[1][2 * gi]]
            if kind == "sub":
                return [a0 - cur[1][2 * gi]]
            if kind == "mul":
                b0 = cur[1][2 * gi]
                return [a0 * b0 % P, (cur[0][2 * gi + 1] - a0) * (cur[1][2 * gi + 1] - b0) % P]
            if kind == "square":
                d = cur[0][2 * gi + 1] - a0
                return [a0 * a0 % P, d * d % P]
            return [a0]
        n_out = 2 if kind in ("mul", "square") else 1
        want = [x % P for x in ref_eq.fold(body, n_out)]
        got = from_mont_array(round_eval(ctx, kid, polys, g))
        assert got == want, (kind, m, rnd)
        c = rand_challenge(rng)
        g.bind(challenge_array(c))
        ref_eq.bind(F.challenge_to_fr(c))
        bind_many(ctx, polys, challenge_array(c), 0)
        cur = [PL.bind(z, F.challenge_to_fr(c), 0) for z in cur]
    for p, z in zip(polys, cur):
        assert from_mont_array(p.final_claim()) == z
        p.free()
    g.free()


@pytest.mark.parametrize("npoly", [2, 3])
@pytest.mark.parametrize("m", [1, 2, 5, 10])
def test_round_eval_dot_every_round(ctx, npoly, m):
    from jolt_atlas_b200 import bind_many, round_eval
    rng = random.Random(m * 3 + npoly)
    zs = [rand_fr(rng, 1 << m) for _ in range(npoly)]
    polys = [_poly(ctx, z) for z in zs]
    cur = [list(z) for z in zs]
    for rnd in range(m):
        half = len(cur[0]) // 2
        want = [0] * npoly
        for i in range(half):
            evs = [PL.sumcheck_evals(z, i, npoly, 1) for z in cur]
            for k in range(npoly):
                t = 1
                for e in evs:
                    t = t * e[k] % P
                want[k] = (want[k] + t) % P
        got = from_mont_array(round_eval(ctx, 16 if npoly == 2 else 17, polys))
        assert got == want
        c = rand_challenge(rng)
        bind_many(ctx, polys, challenge_array(c), 1)
        cur = [PL.bind(z, F.challenge_to_fr(c), 1) for z in cur]
    for p in polys:
        p.free()


@pytest.mark.parametrize("rows,cols", [(4, 8), (64, 64), (16, 256), (256, 16), (128, 2)])
def test_tensor_fold_i32(ctx, rows, cols):
    """einsum fold mk,kn->mn (ops/einsum/mk_kn_mn.rs:47-79): left[j] = sum_i A[i,j] eq_m[i]; right[j] = sum_h B[j,h] eq_n[h]."""
    from jolt_atlas_b200 import EqPolynomial, tensor_fold_i32
    rng = random.Random(rows * 1000 + cols)
    A = np.array([[rng.randrange(-128, 128) for _ in range(cols)] for _ in range(rows)], dtype=np.int32)
    A[0, 0] = -2**31
    A[rows - 1, cols - 1] = 2**31 - 1
    for transpose in (False, True):
        k = cols if transpose else rows
        m = k.bit_length() - 1
        cs = [rand_challenge(rng) for _ in range(m)]
        eq = EqPolynomial.evals(ctx, np.array([F.challenge_limbs(c) for c in cs], dtype=np.uint64).reshape(-1, 4))
        eqv = PL.eq_evals([F.challenge_to_fr(c) for c in cs])
        out = tensor_fold_i32(ctx, A, eq, transpose)
        if transpose:
            want = [sum(int(A[i, j]) * eqv[j] for j in range(cols)) % P for i in range(rows)]
        else:
            want = [sum(int(A[i, j]) * eqv[i] for i in range(rows)) % P for j in range(cols)]
        assert from_mont_array(out.to_host()) == want
        out.free(); eq.free()


def test_large_bind_linearity_property(ctx):
    """Size-independent property at 2^22: bind(a)+bind(b) == bind(a+b) spot-checked, and the fully bound
    value equals MLE evaluation at the reversed challenge vector (LowToHigh binds the LSB first)."""
    from jolt_atlas_b200 import MultilinearPolynomial
    logn = 22
    rng = np.random.default_rng(5)
    small = rng.integers(-2**31, 2**31, size=1 << logn, dtype=np.int64).astype(np.int32)
    p = MultilinearPolynomial.from_i32(ctx, small)
    q = p.clone()
    prng = random.Random(9)
    cs = [rand_challenge(prng) for _ in range(logn)]
    for c in cs:
        p.bind_parallel(challenge_array(c), 0)
    pt = np.array([F.challenge_limbs(c) for c in reversed(cs)], dtype=np.uint64).reshape(-1, 4)
    assert from_mont_array(p.final_claim()) == from_mont_array(q.evaluate(pt))
    # HighToLow binds the MSB first -> same point, natural order
    q2 = q.clone()
    for c in reversed(cs):
        q2.bind_parallel(challenge_array(c), 1)
    assert from_mont_array(q2.final_claim()) == from_mont_array(p.final_claim())
    p.free(); q.free(); q2.free()


@pytest.mark.parametrize("n,m", [(2, 10), (3, 6), (1, 5), (2, 1), (5, 3)])
def test_eval_reduction_h(ctx, n, m):
    """NodeEvalReduction / compute_h (evaluation_reduction.rs:223-249): device evaluate-and-interpolate against the oracle's
    polynomial-valued fold, on an i32 node output (EvalReductionWitness::from_tensor)."""
    from jolt_atlas_b200 import MultilinearPolynomial, eval_reduction_h
    rng = np.random.default_rng(n * 100 + m)
    z = rng.integers(-128, 128, size=1 << m, dtype=np.int32)
    pts = rng.integers(0, 1 << 63, size=(n, m, 4), dtype=np.uint64)
    pts[..., 3] &= np.uint64((1 << 60) - 1)
    p = MultilinearPolynomial.from_i32(ctx, z)
    got = eval_reduction_h(ctx, p, pts)
    want = ORC.eval_reduction_h(ORC.fr_from_i64(z), pts)
    assert np.array_equal(got, want)
    p.free()
