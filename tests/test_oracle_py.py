"""CPU-only: pin the Python oracle with everything the reference's own tests pin for this path
(SURVEY.md §8c — no golden bytes exist upstream, so these are the reference's equivalence invariants
plus public constants): eq serial==cached (eq_poly.rs:265-313), split-eq merge==dense after every bind in
both orders (split_eq_poly.rs:623-669), F*Challenge == F*Fr(challenge) (blake2b.rs:286-316), 125-bit
challenge bound (:265-283), per-round H(0)+H(1)==claim (sumcheck.rs:131-142), prove->verify acceptance,
product-sum optimised==naive (mles_product_sum.rs:1397-1422), HyperKZG prove/verify + tamper rejection
(hyperkzg/tests.rs:19-169), proof length 368 B for l=2 (:108-110), sparse one-hot commit == dense MSM
(:544-680), all-None one-hot commits to identity (:722-745)."""
import hashlib
import random

import pytest

from oracle.pyref import curve as C
from oracle.pyref import field as F
from oracle.pyref import hyperkzg as HK
from oracle.pyref import poly as PL
from oracle.pyref import sumcheck as SC
from oracle.pyref import transcript as TR
from oracle.pyref.unipoly import UniPoly

P = F.P


def test_public_constants():
    # BN254: r and q are prime-looking 254-bit values with q = 36u^4+36u^3+24u^2+6u+1, r = 36u^4+36u^3+18u^2+6u+1
    u = 4965661367192848881
    assert F.Q == 36 * u**4 + 36 * u**3 + 24 * u**2 + 6 * u + 1
    assert F.P == 36 * u**4 + 36 * u**3 + 18 * u**2 + 6 * u + 1
    assert C.is_on_curve(C.G1)
    assert C.scalar_mul(C.G1, P) is None           # generator has order r
    assert (F.INV64 * F.P) % (1 << 64) == (1 << 64) - 1
    assert F.R == pow(2, 256, P) and F.R2 == pow(2, 512, P)


def test_challenge_is_montgomery_reinterpretation():
    rng = random.Random(0)
    for _ in range(100):
        c = rng.getrandbits(128)
        masked = c & F.CHALLENGE_MASK
        assert masked < (1 << 125)
        limbs = F.challenge_limbs(c)
        assert limbs[0] == 0 and limbs[1] == 0
        a = rng.randrange(P)
        # F * challenge == F * Fr::from(challenge) where the conversion reinterprets limbs as Montgomery form
        assert a * F.challenge_to_fr(c) % P == a * F.fr_from_mont(limbs) % P
        assert F.from_limbs(limbs) < P


def test_eq_tables_agree():
    rng = random.Random(1)
    for n in (1, 2, 5, 9):
        r = [rng.randrange(P) for _ in range(n)]
        ev = PL.eq_evals(r)
        cached = PL.eq_evals_cached(r)
        for j in range(n + 1):
            assert cached[j] == PL.eq_evals(r[:j])
        rev = PL.eq_evals_cached_rev(r)
        for j in range(n + 1):
            assert rev[j] == PL.eq_evals(r[n - j:])
        for i in (0, len(ev) - 1, len(ev) // 3):
            bits = [(i >> (n - 1 - k)) & 1 for k in range(n)]
            assert ev[i] == PL.eq_mle(r, bits)
        assert sum(ev) % P == 1


@pytest.mark.parametrize("order", [0, 1])
def test_split_eq_merge_equals_dense_after_every_bind(order):
    rng = random.Random(2 + order)
    n = 7
    w = [rng.randrange(P) for _ in range(n)]
    se = PL.GruenSplitEq(w, order)
    dense = PL.eq_evals(w)
    for _ in range(n):
        assert se.merge() == dense
        # factorisation used by par_fold_out_in: eq = E_out x E_in x (current linear factor)
        eo, ei = se.E_out(), se.E_in()
        assert len(eo) * len(ei) * 2 == len(dense)
        r = rng.randrange(P)
        se.bind(r)
        dense = PL.bind(dense, r, order)
    assert se.merge() == dense


def _verify_split(kind, polys, w, cl, label=b"t"):
    inst = SC.SplitEqInstance(kind, w, polys, cl)
    t = TR.Blake2bTranscript(label)
    cps, rs, fin = SC.sumcheck_prove(inst, t)
    t2 = TR.Blake2bTranscript(label)
    t2.append_scalar(cl)
    e, rs2 = SC.sumcheck_verify(cps, cl, len(w), inst.degree, t2)
    assert rs == rs2 and e == fin and t.state == t2.state
    # H(0)+H(1) == claim each round
    claim = cl
    for cp, c in zip(cps, rs):
        uni = cp.decompress(claim)
        assert (uni.evaluate(0) + uni.evaluate(1)) % P == claim
        claim = uni.evaluate(F.challenge_to_fr(c))
    return inst, rs, e


def test_sumcheck_instances_prove_verify():
    rng = random.Random(3)
    n = 5
    w = [F.challenge_to_fr(rng.getrandbits(128)) for _ in range(n)]
    a = [rng.randrange(-128, 128) % P for _ in range(1 << n)]
    b = [rng.randrange(-128, 128) % P for _ in range(1 << n)]
    cases = {
        "mul": ([a, b], [x * y % P for x, y in zip(a, b)], lambda f: f[0] * f[1]),
        "add": ([a, b], [(x + y) % P for x, y in zip(a, b)], lambda f: f[0] + f[1]),
        "sub": ([a, b], [(x - y) % P for x, y in zip(a, b)], lambda f: f[0] - f[1]),
        "square": ([a], [x * x % P for x in a], lambda f: f[0] ** 2),
        "ident": ([a], a, lambda f: f[0]),
        "cube": ([a], [x ** 3 % P for x in a], lambda f: f[0] ** 3),
        "prod": ([a, b, a], [x * y * x % P for x, y in zip(a, b)], lambda f: f[0] * f[1] * f[2]),
    }
    for kind, (polys, out, fin_f) in cases.items():
        cl = PL.evaluate(out, w)
        inst, rs, e = _verify_split(kind, polys, w, cl)
        rf = [F.challenge_to_fr(c) for c in rs]
        eqv = PL.eq_mle(w, list(reversed(rf)))
        fc = inst.final_claims()
        assert e == eqv * fin_f(fc) % P
        assert fc[0] == PL.evaluate(polys[0], list(reversed(rf)))


def test_product_sum_grid_equals_naive_interpolation():
    """mles_product_sum optimised == naive: the degree-(d+1) round poly equals the direct evaluation."""
    rng = random.Random(4)
    n, d = 4, 5
    w = [rng.randrange(P) for _ in range(n)]
    polys = [[rng.randrange(P) for _ in range(1 << n)] for _ in range(d)]
    out = [1] * (1 << n)
    for z in polys:
        out = [x * y % P for x, y in zip(out, z)]
    cl = PL.evaluate(out, w)
    inst = SC.SplitEqInstance("prod", w, polys, cl)
    uni = inst.compute_message(0, cl)
    # naive: g(X) = sum_j eq(w, (j, X)) prod_i p_i(j, X)   (LowToHigh: X is the LSB)
    for X in range(0, d + 3):
        tot = 0
        eqt = PL.eq_evals(w[:-1])
        lin = ((1 - w[-1]) * (1 - X) + w[-1] * X) % P
        for j in range(1 << (n - 1)):
            pr = 1
            for z in polys:
                pr = pr * (z[2 * j] + X * (z[2 * j + 1] - z[2 * j])) % P
            tot = (tot + eqt[j] * pr) % P
        assert uni.evaluate(X) == tot * lin % P


def test_batched_sumcheck_front_loaded():
    rng = random.Random(5)
    w5 = [rng.randrange(P) for _ in range(5)]
    w3 = [rng.randrange(P) for _ in range(3)]
    a5 = [rng.randrange(P) for _ in range(32)]
    b5 = [rng.randrange(P) for _ in range(32)]
    a3 = [rng.randrange(P) for _ in range(8)]
    i1 = SC.SplitEqInstance("mul", w5, [a5, b5], PL.evaluate([x * y % P for x, y in zip(a5, b5)], w5))
    i2 = SC.SplitEqInstance("square", w3, [a3], PL.evaluate([x * x % P for x in a3], w3))
    i3 = SC.HammingInstance([a3, a3[::-1]], [3, 5], (3 * sum(a3) + 5 * sum(a3)) % P)
    t = TR.Blake2bTranscript(b"batched")
    cps, rs, coeffs, claims = SC.batched_sumcheck_prove([i1, i2, i3], t)
    # verifier side (sumcheck.rs:190-259)
    t2 = TR.Blake2bTranscript(b"batched")
    for i in (i1, i2, i3):
        t2.append_scalar(i.claim)
    co2 = t2.challenge_vector(3)
    assert co2 == coeffs
    claim = (i1.claim * co2[0] + F.mul_pow_2(i2.claim, 2) * co2[1] + F.mul_pow_2(i3.claim, 2) * co2[2]) % P
    e, rs2 = SC.sumcheck_verify(cps, claim, 5, 3, t2)
    assert rs2 == rs
    rf = [F.challenge_to_fr(c) for c in rs]
    f1, f2, f3 = i1.final_claims(), i2.final_claims(), i3.final_claims()
    exp = (PL.eq_mle(w5, rf[::-1]) * f1[0] * f1[1] * co2[0]
           + PL.eq_mle(w3, rf[2:][::-1]) * f2[0] ** 2 * co2[1]
           + (3 * f3[0] + 5 * f3[1]) * co2[2]) % P
    assert e == exp


def test_transcript_against_hashlib_layout():
    t = TR.Blake2bTranscript(b"ONNXProof")
    s0 = hashlib.blake2b(b"ONNXProof" + b"\0" * 23, digest_size=32).digest()
    assert t.state == s0
    t.append_scalar(5)
    s1 = hashlib.blake2b(s0 + b"\0" * 28 + (0).to_bytes(4, "big") + (5).to_bytes(32, "big"), digest_size=32).digest()
    assert t.state == s1
    c = t.challenge_u128()
    s2 = hashlib.blake2b(s1 + b"\0" * 28 + (1).to_bytes(4, "big"), digest_size=32).digest()
    assert c == int.from_bytes(s2[:16], "little") and t.state == s2
    x = t.challenge_scalar()
    s3 = hashlib.blake2b(s2 + b"\0" * 28 + (2).to_bytes(4, "big"), digest_size=32).digest()
    assert x == int.from_bytes(s3[:16], "big")


def test_unipoly_interpolation_paths_agree():
    rng = random.Random(6)
    for deg in (1, 2, 3, 4, 7):
        co = [rng.randrange(P) for _ in range(deg + 1)]
        u = UniPoly(co)
        ev = [u.evaluate(i) for i in range(deg + 1)]
        assert UniPoly.from_evals(ev).coeffs == co
        hint = (ev[0] + ev[1]) % P
        assert UniPoly.from_evals_and_hint(hint, [ev[0]] + ev[2:]).coeffs == co
        toom = [u.evaluate(i) for i in range(deg)] + [co[-1]]
        assert UniPoly.from_evals_toom(toom).coeffs == co
        cp = u.compress()
        assert cp.decompress(hint).coeffs == co
        r = rng.randrange(P)
        assert cp.eval_from_hint(hint, r) == u.evaluate(r)
    # fixed-length deg-2/3 interpolation keeps zero leading coefficients; general path trims
    assert len(UniPoly.from_evals([1, 2, 3]).coeffs) == 3
    assert len(UniPoly.from_evals([1, 2, 3, 4]).coeffs) == 4
    assert len(UniPoly.from_evals([1, 2, 3, 4, 5]).coeffs) == 2


def test_hyperkzg_open_verify_and_pins():
    rng = random.Random(7)
    for ell in (2, 3, 5):
        n = 1 << ell
        srs = HK.srs_powers(n)
        poly = [rng.randrange(P) for _ in range(n)]
        pt = [rng.getrandbits(128) & F.CHALLENGE_MASK for _ in range(ell)]
        cm = HK.commit(srs, poly)
        assert cm == C.msm_naive(srs, poly)
        y = PL.evaluate(poly, [F.challenge_to_fr(c) for c in pt])
        t = TR.Blake2bTranscript(b"TestEval")
        pr = HK.open(srs, poly, pt, t)
        t2 = TR.Blake2bTranscript(b"TestEval")
        assert HK.verify(srs[0], HK.TEST_TAU, cm, pt, y, pr, t2)
        assert t.state == t2.state
        # tamper -> reject (hyperkzg/tests.rs:112-125)
        bad = dict(pr); bad["v"] = [list(r) for r in pr["v"]]; bad["v"][0][0] = (bad["v"][0][0] + 1) % P
        assert not HK.verify(srs[0], HK.TEST_TAU, cm, pt, y, bad, TR.Blake2bTranscript(b"TestEval"))
        assert not HK.verify(srs[0], HK.TEST_TAU, cm, pt, (y + 1) % P, pr, TR.Blake2bTranscript(b"TestEval"))
        if ell == 2:
            assert len(HK.serialize_proof(pr)) == 368   # the reference's only byte-level pin


def test_one_hot_commit_equals_dense_msm():
    rng = random.Random(8)
    K, T = 4, 8
    srs = HK.srs_powers(K * T)
    idx = [rng.randrange(K) if rng.random() < 0.8 else None for _ in range(T)]
    dense = [0] * (K * T)
    for t, k in enumerate(idx):
        if k is not None:
            dense[k * T + t] = 1
    assert HK.commit_one_hot(srs, idx, K) == HK.commit(srs, dense)
    assert HK.commit_one_hot(srs, [None] * T, K) is None
    # signed small scalars == pos/neg split (msm/mod.rs:93-176)
    ints = [rng.randrange(-2**31, 2**31) for _ in range(K * T)]
    pos = C.msm_pippenger(srs, [v if v > 0 else 0 for v in ints])
    neg = C.msm_pippenger(srs, [-v if v < 0 else 0 for v in ints])
    assert C.msm_i(srs, ints) == C.add_affine(pos, C.neg_affine(neg))
