"""GPU parity tests (through the C ABI) for HyperKZG::open (hyperkzg/mod.rs:400-447): commitments of the folded
polynomials, the 3 x l evaluations, the witness commitments and the transcript state after the opening must equal
the oracle's (oracle/cpp) and the committed golden vectors (tests/golden/hyperkzg.json, from the Python twin).
Also the split (caller-owned transcript) form == the all-in-one form, and a pairing-free soundness check with the
known test tau: e(C - v G, H) = e(W, (tau - u) H)  <=>  C - v*G == (tau - u) * W  on G1."""
import json
import os
import random

import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import curve as CV
from oracle.pyref import field as F
from oracle.pyref.transcript import Blake2bTranscript
from tests.util import from_mont_array, rand_challenge, rand_fr, to_mont_array

pytestmark = pytest.mark.gpu

P = F.P
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAU = 0x1234567890abcdef1122334455667788


def _pts(arr, inf):
    return [None if i else (F.fq_from_mont(p[:4]), F.fq_from_mont(p[4:])) for p, i in zip(arr, inf)]


@pytest.fixture(scope="module")
def srs_host():
    return ORC.srs_powers(to_mont_array([TAU])[0], 1 << 12)


@pytest.fixture(scope="module")
def srs(ctx, srs_host):
    from jolt_atlas_b200 import SRS
    s = SRS(ctx, srs_host)
    yield s
    s.free()


@pytest.mark.parametrize("ell", [1, 2, 3, 5, 8, 10, 12])
def test_open_matches_oracle(ctx, srs, srs_host, ell):
    from jolt_atlas_b200 import Blake2bTranscriptState, MultilinearPolynomial, hyperkzg_open
    rng = random.Random(ell)
    n = 1 << ell
    z = to_mont_array(rand_fr(rng, n))
    pt = np.array([F.challenge_limbs(rand_challenge(rng)) for _ in range(ell)], dtype=np.uint64)
    want = ORC.hyperkzg_open(srs_host[:n].copy(), z, pt, b"TestEval")
    poly = MultilinearPolynomial.from_fr(ctx, z)
    t = Blake2bTranscriptState(b"TestEval")
    got = hyperkzg_open(ctx, srs, poly, pt, t)
    assert _pts(got["com"], got["com_inf"]) == _pts(want["com"], want["com_inf"])
    assert np.array_equal(got["v"], want["v"])
    assert _pts(got["w"], got["w_inf"]) == _pts(want["w"], want["w_inf"])
    assert t.state == want["state"]
    assert len(poly) == n            # the opened polynomial is not consumed
    poly.free()


def test_open_golden(ctx):
    from jolt_atlas_b200 import SRS, Blake2bTranscriptState, MultilinearPolynomial, hyperkzg_open
    g = json.load(open(os.path.join(G, "hyperkzg.json")))
    srs_h = ORC.srs_powers(to_mont_array([int(g["tau"], 16)])[0], 32)
    s = SRS(ctx, srs_h)
    for case in g["cases"]:
        poly = MultilinearPolynomial.from_fr(ctx, to_mont_array([int(x, 16) for x in case["poly"]]))
        pt = np.array([F.challenge_limbs(int(c, 16)) for c in case["point"]], dtype=np.uint64)
        t = Blake2bTranscriptState(b"TestEval")
        got = hyperkzg_open(ctx, s, poly, pt, t)
        assert [[hex(v) for v in p] for p in _pts(got["com"], got["com_inf"])] == case["com"]
        assert [[hex(v) for v in p] for p in _pts(got["w"], got["w_inf"])] == case["w"]
        assert [[hex(v) for v in from_mont_array(row)] for row in got["v"]] == case["v"]
        assert t.state.hex() == case["transcript_state"]
        poly.free()
    s.free()


def test_split_form_with_caller_transcript(ctx, srs, srs_host):
    """The Rust shim's shape: the caller owns the transcript (here the Python twin) between the three calls."""
    from jolt_atlas_b200 import Blake2bTranscriptState, HyperKZGOpening, MultilinearPolynomial, hyperkzg_open
    rng = random.Random(77)
    ell = 9
    z = to_mont_array(rand_fr(rng, 1 << ell))
    pt = np.array([F.challenge_limbs(rand_challenge(rng)) for _ in range(ell)], dtype=np.uint64)
    poly = MultilinearPolynomial.from_fr(ctx, z)
    t = Blake2bTranscript(b"split")
    op = HyperKZGOpening(ctx, srs, poly, pt)
    t.append_points(_pts(op.com, op.com_inf))
    r = t.challenge_scalar()
    v = op.evals(to_mont_array([r])[0])
    t.append_scalars([x for row in v for x in from_mont_array(row)])
    q = t.challenge_scalar_powers(ell)
    w, w_inf = op.witness(to_mont_array([r])[0], to_mont_array(q))
    t.append_points(_pts(w, w_inf))
    t.challenge_scalar()
    op.free()
    ts = Blake2bTranscriptState(b"split")
    one = hyperkzg_open(ctx, srs, poly, pt, ts)
    assert np.array_equal(one["v"], v) and np.array_equal(one["w"], w) and np.array_equal(one["com"], op.com)
    assert ts.state == t.state and ts.n_rounds == t.n_rounds
    # pairing-free KZG check with the known tau: B(tau) - B(u_i) == (tau - u_i) * h_i(tau)  <=>  C_B - B(u_i) G == (tau - u_i) W_i
    u = [r, (-r) % P, r * r % P]
    com_pts = _pts(one["com"], one["com_inf"])
    c0 = ORC.msm_fr(srs_host[: 1 << ell].copy(), z)
    cs = [(F.fq_from_mont(c0[0][:4]), F.fq_from_mont(c0[0][4:]))] + com_pts
    CB = None
    for k in range(ell):
        CB = CV.add_affine(CB, CV.scalar_mul(cs[k], q[k]))
    gen = (1, 2)
    for i in range(3):
        Bu = sum(q[k] * from_mont_array(one["v"][i])[k] for k in range(ell)) % P
        lhs = CV.add_affine(CB, CV.scalar_mul(gen, (-Bu) % P))
        wi = _pts(one["w"][i: i + 1], one["w_inf"][i: i + 1])[0]
        assert lhs == CV.scalar_mul(wi, (TAU - u[i]) % P)
    poly.free()


def test_open_errors(ctx, srs):
    from jolt_atlas_b200 import Blake2bTranscriptState, JoltAtlasError, MultilinearPolynomial, hyperkzg_open
    z = to_mont_array([1] * 16)
    poly = MultilinearPolynomial.from_fr(ctx, z)
    pt3 = np.array([F.challenge_limbs(5)] * 3, dtype=np.uint64)
    with pytest.raises(JoltAtlasError):            # n != 2^ell (reference: assert_eq!(n, 1 << ell))
        hyperkzg_open(ctx, srs, poly, pt3, Blake2bTranscriptState(b"x"))
    big = MultilinearPolynomial.random(ctx, 1 << 13, 3)
    pt13 = np.array([F.challenge_limbs(5)] * 13, dtype=np.uint64)
    with pytest.raises(JoltAtlasError) as e:       # SRS has 2^12 powers
        hyperkzg_open(ctx, srs, big, pt13, Blake2bTranscriptState(b"x"))
    assert e.value.code == -3
    poly.free(); big.free()
