"""CPU-only: the oracle's restatement of the three-to-five-operand split-eq bodies (Iff, Div, Rsqrt, teleport division /
ScalarConstDiv) is pinned the way the reference pins its operators (ops/test.rs:10-20: prove -> verify): the C++ oracle
proves on TRUE claims computed here with Python big integers straight from the operator relations, and the reference's
verifier logic (sumcheck.rs:653-686) replayed over the round polynomials must end at eq(w, r) * body(final openings) —
which no wrong round body can satisfy.  Relations: ops/iff.rs:162-171, ops/div.rs:302-311, ops/rsqrt.rs:343-367,
neural_teleport/division.rs:205-213, ops/scalar_const_div.rs:227-239."""
import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from oracle.pyref import transcript as TR
from oracle.pyref.unipoly import CompressedUniPoly
from tests.util import from_mont_array, to_mont_array

P = F.P
IFF, DIV, RSQRT, LIN3, SUB = 8, 9, 10, 11, 1


def _chal(rng, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def eq_table(w):
    """EqPolynomial::evals, big-endian index (eq_poly.rs:77-101)."""
    t = [1]
    for wj in w:
        t = [x for e in t for x in (e * (1 - wj) % P, e * wj % P)]
    return t


def body(kind, vals, aux):
    if kind == IFF:
        m, a, b = vals
        return (m * a + (1 - m) * b) % P
    if kind == DIV:
        l, r, q, R = vals
        return (r * q + R - l) % P
    if kind == RSQRT:
        x, quot, out, dr, sr = vals
        return (x * quot + dr - aux[1] + aux[0] * (out * out + sr - quot)) % P
    if kind == LIN3:
        inp, q, r = vals
        return (aux[0] * q + r - inp) % P
    l, R = vals
    return (l - R) % P


CASES = [(IFF, 3, 0), (DIV, 4, 0), (RSQRT, 5, 2), (LIN3, 3, 1), (SUB, 2, 0)]


def make_case(kind, npoly, naux, m, seed):
    rng = np.random.default_rng(seed)
    n = 1 << m
    cols = [[int(v) % P for v in rng.integers(-(1 << 20), 1 << 20, size=n)] for _ in range(npoly)]
    if kind == IFF:
        cols[0] = [int(v) for v in rng.integers(0, 2, size=n)]                      # a 0/1 mask, as the tracer produces
    aux = [int(v) for v in rng.integers(1, 1 << 40, size=naux)]
    w = _chal(rng, m)
    w_fr = from_mont_array(w)
    eq = eq_table(w_fr)
    claim = sum(eq[i] * body(kind, [c[i] for c in cols], aux) for i in range(n)) % P
    return cols, aux, w, w_fr, claim


def verify(coeffs, claim, label):
    t = TR.Blake2bTranscript(label)
    t.append_scalar(claim)
    e, rs = claim % P, []
    for c in coeffs:
        cp = CompressedUniPoly(from_mont_array(c))
        cp.append_to_transcript(t)
        r = F.challenge_to_fr(t.challenge_scalar_optimized())
        e = cp.eval_from_hint(e, r)
        rs.append(r)
    return e, rs


@pytest.mark.parametrize("kind,npoly,naux", CASES)
@pytest.mark.parametrize("m", [1, 2, 5, 8])
def test_oracle_proof_verifies_against_the_operator_relation(kind, npoly, naux, m):
    cols, aux, w, w_fr, claim = make_case(kind, npoly, naux, m, 1000 * kind + m)
    polys = np.stack([to_mont_array(c) for c in cols])
    t = ORC.TranscriptState(b"body")
    got = ORC.sumcheck_prove_st(0, kind, polys, w, to_mont_array([claim])[0], t, gammas=to_mont_array(aux) if naux else None)
    deg = 2 if kind in (LIN3, SUB) else 3
    assert all(c.shape[0] == deg for c in got["coeffs"])                          # compressed: degree + 1 - 1 coefficients
    e, rs = verify(got["coeffs"], claim, b"body")
    finals = from_mont_array(got["final_claims"])
    # LowToHigh binding: challenge j binds variable m-1-j, so the opening point is the reversed challenge vector
    point = rs[::-1]
    eq_at_r = 1
    for wj, rj in zip(w_fr, point):
        eq_at_r = eq_at_r * ((wj * rj + (1 - wj) * (1 - rj)) % P) % P
    assert e == eq_at_r * body(kind, finals, aux) % P
    # ... and the final openings are the MLEs of the operands at that point
    eqp = eq_table(point)
    for c, f in zip(cols, finals):
        assert f == sum(x * y for x, y in zip(eqp, c)) % P
