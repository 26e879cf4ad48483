"""GPU parity tests (through the C ABI) for Sumcheck::prove (subprotocols/sumcheck.rs:565-599): every round's compressed
polynomial, every challenge, the final MLE claims and the transcript state after the proof must equal the committed
golden vectors (tests/golden/sumcheck.json, from the Python twin) and the C++ oracle on fresh seeded inputs.
The per-round invariant H(0)+H(1)==claim (sumcheck.rs:131-142) is implied by equality with the oracle, whose
verifier-side check runs in tests/test_oracle_py.py.  Also the raw round-evaluation kernels PROD / POW / SUM1 / SUMHI
against the pyref formulas (mles_product_sum.rs:61-129, hamming_weight.rs:118-139, ops/sum/axis.rs:220-233)."""
import json
import os
import random

import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from oracle.pyref import poly as PL
from tests.util import from_mont_array, rand_challenge, rand_fr, to_mont_array

pytestmark = pytest.mark.gpu

P = F.P
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KIND = {"add": 0, "sub": 1, "mul": 2, "square": 3, "prod": 4, "cube": 5, "dot2": 16, "dot3": 17}
ORC_KIND = {"add": (0, 0), "sub": (0, 1), "mul": (0, 2), "square": (0, 3), "prod": (0, 4), "cube": (0, 5), "dot2": (1, 0), "dot3": (1, 0)}


def _run_gpu(ctx, kind, polys_fr, w, claim, label, pow_d=0):
    from jolt_atlas_b200 import Blake2bTranscriptState, MultilinearPolynomial, sumcheck_prove
    ps = [MultilinearPolynomial.from_fr(ctx, z) for z in polys_fr]
    t = Blake2bTranscriptState(label)
    res = sumcheck_prove(ctx, KIND[kind], ps, claim, t, eq_w=w if KIND[kind] < 16 else None, pow_d=pow_d)
    for p in ps:
        p.free()
    return res, t


def test_sumcheck_golden(ctx):
    for case in json.load(open(os.path.join(G, "sumcheck.json"))):
        kind = case["kind"]
        polys = [to_mont_array([v % P for v in z]) for z in case["polys_i32"]]
        w = np.array([F.challenge_limbs(int(c, 16)) for c in case["w_challenges"]], dtype=np.uint64).reshape(-1, 4)
        res, t = _run_gpu(ctx, kind, polys, w, to_mont_array([int(case["claim"], 16)])[0], case["label"].encode(),
                          pow_d=3 if kind == "cube" else 0)
        assert [[hex(v) for v in from_mont_array(cp)] for cp in res["coeffs"]] == case["round_polys"], kind
        assert [hex((int(r[3]) << 64) | int(r[2])) for r in res["challenges"]] == case["challenges"], kind
        assert [hex(v) for v in from_mont_array(res["final_claims"])] == case["final_poly_claims"], kind
        assert t.state.hex() == case["transcript_state"], kind


@pytest.mark.parametrize("kind,npoly,m", [("add", 2, 11), ("sub", 2, 6), ("mul", 2, 12), ("square", 1, 9), ("prod", 4, 8),
                                          ("prod", 16, 7), ("prod", 5, 6), ("cube", 1, 8), ("dot2", 2, 10), ("dot3", 3, 9),
                                          ("mul", 2, 1), ("dot2", 2, 1), ("prod", 2, 3)])
def test_sumcheck_matches_oracle(ctx, kind, npoly, m):
    rng = random.Random(hash((kind, npoly, m)) & 0xffff)
    n = 1 << m
    polys = np.stack([to_mont_array(rand_fr(rng, n)) for _ in range(npoly)])
    w = np.array([F.challenge_limbs(rand_challenge(rng)) for _ in range(m)], dtype=np.uint64)
    claim = to_mont_array([rng.randrange(P)])[0]     # the prover does not check the claim; any value exercises the same path
    fam, ok = ORC_KIND[kind]
    pow_d = 3 if kind == "cube" else 0
    want = ORC.sumcheck_prove(fam, ok, polys, w, claim, b"parity", pow_d=pow_d)
    res, t = _run_gpu(ctx, kind, list(polys), w, claim, b"parity", pow_d=pow_d)
    assert len(res["coeffs"]) == m
    for r in range(m):
        assert np.array_equal(res["coeffs"][r], want["coeffs"][r]), (kind, r)
    assert np.array_equal(res["challenges"], want["challenges"])
    assert np.array_equal(res["final_claims"], want["final_claims"])
    assert t.state == want["state"]


def test_round_eval_prod_and_sums(ctx):
    from jolt_atlas_b200 import EvalKernel, GruenSplitEqPolynomial, MultilinearPolynomial, round_eval
    rng = random.Random(31)
    m = 7
    n = 1 << m
    for d in (2, 3, 4, 7, 8, 16, 17, 32):
        zs = [rand_fr(rng, n) for _ in range(d)]
        w = [F.challenge_to_fr(rand_challenge(rng)) for _ in range(m)]
        ps = [MultilinearPolynomial.from_fr(ctx, to_mont_array(z)) for z in zs]
        eq = GruenSplitEqPolynomial(ctx, to_mont_array(w), 0)
        ref = PL.GruenSplitEq(w, 0)

        def per_g(g):
            out = []
            for k in range(d):
                acc = 1
                for z in zs:
                    p0, dp = z[2 * g], (z[2 * g + 1] - z[2 * g]) % P
                    acc = acc * (dp if k == d - 1 else (p0 + (k + 1) * dp)) % P
                out.append(acc)
            return out
        want = ref.fold(per_g, d)
        got = from_mont_array(round_eval(ctx, EvalKernel.PROD, ps, eq, n_out=d))
        assert got == want, d
        for p in ps:
            p.free()
        eq.free()
    # SUM1 (Hamming weight) with gammas, SUMHI
    zs = [rand_fr(rng, n) for _ in range(5)]
    gam = rand_fr(rng, 5)
    ps = [MultilinearPolynomial.from_fr(ctx, to_mont_array(z)) for z in zs]
    got = from_mont_array(round_eval(ctx, EvalKernel.SUM1, ps, None, aux_fr=to_mont_array(gam), n_out=1))
    assert got == [sum(g * sum(z[0::2]) for g, z in zip(gam, zs)) % P]
    got = from_mont_array(round_eval(ctx, EvalKernel.SUMHI, ps[:1], None, n_out=1))
    assert got == [sum(zs[0][: n // 2]) % P]
    for p in ps:
        p.free()


@pytest.mark.parametrize("kind_id,fam_ok,npoly,m", [(0, (0, 0), 2, 19), (1, (0, 1), 2, 18), (6, (0, 6), 1, 19)])
def test_tma_staged_round_kernels_match_oracle(ctx, kind_id, fam_ok, npoly, m):
    """Slabs of >= 2^16 pairs go through the TMA-staged kernel (csrc/tma_round.cuh): same proof as the C++ oracle and as
    the register-staged kernel (JA_NO_TMA=1) on the same inputs."""
    from jolt_atlas_b200 import Blake2bTranscriptState, MultilinearPolynomial, sumcheck_prove
    rng = np.random.default_rng(kind_id * 100 + m)
    polys = rng.integers(0, 1 << 63, size=(npoly, 1 << m, 4), dtype=np.uint64)
    polys[..., 3] &= np.uint64((1 << 60) - 1)                                       # canonical: top limb < 2^60 < p's top limb
    w = np.zeros((m, 4), dtype=np.uint64)
    w[:, 2] = rng.integers(0, 1 << 63, size=m, dtype=np.uint64)
    w[:, 3] = rng.integers(0, 1 << 61, size=m, dtype=np.uint64)
    claim = polys[0, 0].copy()
    want = ORC.sumcheck_prove(fam_ok[0], fam_ok[1], polys, w, claim, b"tma")
    outs = []
    for no_tma in (False, True):
        if no_tma:
            os.environ["JA_NO_TMA"] = "1"
        try:
            ps = [MultilinearPolynomial.from_fr(ctx, z) for z in polys]
            t = Blake2bTranscriptState(b"tma")
            res = sumcheck_prove(ctx, kind_id, ps, claim, t, eq_w=w)
            for p in ps:
                p.free()
        finally:
            os.environ.pop("JA_NO_TMA", None)
        outs.append((res, t.state))
    for res, state in outs:
        assert len(res["coeffs"]) == m
        for r in range(m):
            assert np.array_equal(res["coeffs"][r], want["coeffs"][r]), r
        assert np.array_equal(res["challenges"], want["challenges"])
        assert np.array_equal(res["final_claims"], want["final_claims"])
        assert state == want["state"]


def test_prelaunched_rounds_survive_mailbox_reuse(ctx):
    """The engine enqueues round j+1 ahead of its challenge and hands it over through a ring of 64 tagged mailbox
    entries (csrc/fused_kernels.cuh: MailRef).  Thousands of rounds on one context wrap every counter involved; every
    proof must equal the first one (a stale-tag match would bind a stale challenge or clobber the result slot)."""
    from jolt_atlas_b200 import Blake2bTranscriptState, MultilinearPolynomial, sumcheck_prove
    rng = np.random.default_rng(77)
    m = 5
    z = rng.integers(0, 1 << 63, size=(2, 1 << m, 4), dtype=np.uint64)
    z[..., 3] &= np.uint64((1 << 60) - 1)
    w = np.zeros((m, 4), dtype=np.uint64)
    w[:, 2] = rng.integers(0, 1 << 63, size=m, dtype=np.uint64)
    w[:, 3] = rng.integers(0, 1 << 61, size=m, dtype=np.uint64)
    claim = z[0, 0].copy()
    want = ORC.sumcheck_prove(0, 2, z, w, claim, b"ring")
    first = None
    for it in range(5000):                       # 4 pre-launched rounds each: 20000 mailbox uses, > 256 per entry
        ps = [MultilinearPolynomial.from_fr(ctx, z[0]), MultilinearPolynomial.from_fr(ctx, z[1])]
        t = Blake2bTranscriptState(b"ring")
        res = sumcheck_prove(ctx, 2, ps, claim, t, eq_w=w)
        for p in ps:
            p.free()
        key = (b"".join(c.tobytes() for c in res["coeffs"]), res["final_claims"].tobytes(), t.state)
        if first is None:
            first = key
            for r in range(m):
                assert np.array_equal(res["coeffs"][r], want["coeffs"][r]), r
            assert t.state == want["state"]
        else:
            assert key == first, it


@pytest.mark.parametrize("kind,npoly,m", [("mul", 2, 9), ("dot2", 2, 7), ("prod", 4, 6)])
def test_cache_openings_appends_the_final_claims(ctx, kind, npoly, m):
    """ja_set_cache_openings: the driver appends every final claim with Transcript::append_scalar at the end of the proof
    (SumcheckInstanceProver::cache_openings -> opening_proof.rs:281, :338, :398).  Same state as the oracle's transcript after the
    oracle's proof + one append_scalar per claim; OFF (the default) leaves the transcript where the last round left it."""
    from jolt_atlas_b200 import api as A
    rng = random.Random(77 + m)
    n = 1 << m
    polys = [to_mont_array(rand_fr(rng, n)) for _ in range(npoly)]
    w = np.array([F.challenge_limbs(rand_challenge(rng)) for _ in range(m)], dtype=np.uint64)
    claim = to_mont_array([rng.randrange(P)])[0]
    fam, kid = ORC_KIND[kind]
    tc = ORC.TranscriptState(b"cache")
    want = ORC.sumcheck_prove_st(fam, kid, np.stack(polys), w if fam == 0 else None, claim, tc)
    state_plain = tc.state
    ORC.transcript_append_scalar_each(tc, want["final_claims"])
    states = {}
    for on in (0, 1):
        A.check(ctx._lib.ja_set_cache_openings(ctx._h, on))
        try:
            ps = [A.MultilinearPolynomial.from_fr(ctx, z) for z in polys]
            t = A.Blake2bTranscriptState(b"cache")
            got = A.sumcheck_prove(ctx, KIND[kind], ps, claim, t, eq_w=w if KIND[kind] < 16 else None)
            states[on] = (t.state, t.n_rounds)
            assert np.array_equal(got["final_claims"], want["final_claims"])
        finally:
            ctx._lib.ja_set_cache_openings(ctx._h, 0)
    assert states[0][0] == state_plain
    assert states[1] == (tc.state, tc.n_rounds)
    # the caller-side form of the same appends
    t2 = A.Blake2bTranscriptState(b"cache")
    t2.state, t2.n_rounds = states[0]
    A.transcript_append_scalar_each(ctx, t2, want["final_claims"])
    assert (t2.state, t2.n_rounds) == states[1]
