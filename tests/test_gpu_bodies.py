"""GPU parity of the three-to-five-operand split-eq bodies (JA_EVAL_IFF / DIV / RSQRT / LIN3; ScalarConstDiv = SUB) against the
C++ oracle, bit-exact, through every path the library has for them: the round-resident kernel (default), the per-round
fused kernels with pre-launched rounds (JA_NO_PERSIST=1), plain per-round launches (JA_NO_AHEAD=1), and the un-fused
ja_round_eval + ja_bind_many pair a Rust caller with its own transcript would drive."""
import numpy as np
import pytest

from oracle import cpu as ORC
from tests.test_oracle_bodies import CASES, make_case
from tests.util import to_mont_array

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["persist", "per_round_ahead", "per_round_plain"])
@pytest.mark.parametrize("kind,npoly,naux", CASES)
@pytest.mark.parametrize("m", [1, 4, 11, 14])
def test_sumcheck_prove_matches_oracle(ctx, monkeypatch, mode, kind, npoly, naux, m):
    from jolt_atlas_b200 import Blake2bTranscriptState, MultilinearPolynomial, sumcheck_prove
    if mode != "persist":
        monkeypatch.setenv("JA_NO_PERSIST", "1")
    if mode == "per_round_plain":
        monkeypatch.setenv("JA_NO_AHEAD", "1")
    cols, aux, w, w_fr, claim = make_case(kind, npoly, naux, m, 77 * kind + m)
    host = np.stack([to_mont_array(c) for c in cols])
    claim_m = to_mont_array([claim])[0]
    aux_m = to_mont_array(aux) if naux else None
    t_dev, t_cpu = Blake2bTranscriptState(b"bodies"), ORC.TranscriptState(b"bodies")
    polys = [MultilinearPolynomial.from_fr(ctx, host[i]) for i in range(npoly)]
    got = sumcheck_prove(ctx, kind, polys, claim_m, t_dev, eq_w=w, gammas=aux_m)
    want = ORC.sumcheck_prove_st(0, kind, host, w, claim_m, t_cpu, gammas=aux_m)
    assert len(got["coeffs"]) == len(want["coeffs"]) == m
    for i, (a, b) in enumerate(zip(got["coeffs"], want["coeffs"])):
        assert np.array_equal(a, b), f"round {i}"
    assert np.array_equal(got["challenges"], want["challenges"])
    assert np.array_equal(got["final_claims"], want["final_claims"])
    assert t_dev.state == t_cpu.state and t_dev.n_rounds == t_cpu.n_rounds
    for p in polys:
        p.free()


@pytest.mark.parametrize("kind,npoly,naux", CASES)
def test_unfused_round_eval_matches_fused_rounds(ctx, kind, npoly, naux):
    """ja_round_eval sums of every round == what the fused path's round polynomials were assembled from: drive the rounds by
    hand with the oracle's challenges and compare the reduced sums with the oracle's own fold of the same arrays."""
    from jolt_atlas_b200 import GruenSplitEqPolynomial, MultilinearPolynomial, bind_many, round_eval
    from oracle.pyref import field as F
    from oracle.pyref import poly as PL
    from tests.test_oracle_bodies import body
    from tests.util import challenge_array, from_mont_array, rand_challenge
    import random
    m = 7
    cols, aux, w, w_fr, claim = make_case(kind, npoly, naux, m, 5 * kind + 1)
    polys = [MultilinearPolynomial.from_fr(ctx, to_mont_array(c)) for c in cols]
    eq = GruenSplitEqPolynomial(ctx, w, 0)
    ref = PL.GruenSplitEq(w_fr, 0)
    aux_m = to_mont_array(aux) if naux else None
    rng = random.Random(kind)
    cur = [list(c) for c in cols]
    P = F.P
    for _ in range(m):
        got = from_mont_array(round_eval(ctx, kind, polys, eq, aux_fr=aux_m))
        def f(g):
            lo = [c[2 * g] for c in cur]
            hi = [c[2 * g + 1] for c in cur]
            c0 = body(kind, lo, aux)
            if kind in (11, 1):
                return [c0]
            # quadratic coefficient = body(lo + X d) at X^2: body(hi) - 2 body(mid) ... use the closed forms
            d = [(h - l) % P for l, h in zip(lo, hi)]
            if kind == 8:
                return [c0, d[0] * (d[1] - d[2]) % P]
            if kind == 9:
                return [c0, d[1] * d[2] % P]
            return [c0, (d[0] * d[1] + aux[0] * d[2] * d[2]) % P]
        want = ref.fold(f, len(got))
        assert got == want
        c = rand_challenge(rng)
        eq.bind(challenge_array(c)); ref.bind(F.challenge_to_fr(c))
        bind_many(ctx, polys, challenge_array(c), 0)
        cur = [PL.bind(x, F.challenge_to_fr(c), 0) for x in cur]
    for p in polys:
        p.free()
