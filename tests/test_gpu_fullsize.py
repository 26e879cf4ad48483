"""Size-independent properties at the nanoGPT sizes of BASELINE.json (T = 2^14, K T = 2^18), where the oracle would be too
slow to recompute everything:
  * the device sumcheck proofs VERIFY: the reference's verifier logic (sumcheck.rs:653-686) replayed over the round
    polynomials, and the final claim equals eq(w, r) * body(final openings) — for Mul, product-of-16 and the batched RA
    one-hot checks (Booleanity's claim is 0 on one-hot data);
  * commitment homomorphism ("checksum of checksums"): the commitment of the materialised RLC polynomial equals
    sum_i gamma_i * C_i over all one-hot commitments of the batch (HyperKZG::combine_commitments,
    commitment_scheme.rs:91-101) — ties ja_addr_commit, ja_rlc_add_onehot and the MSM together;
  * errors surface as JoltAtlasError with the reference's meaning."""
import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from oracle.pyref import poly as PL
from oracle.pyref import transcript as TR
from oracle.pyref.unipoly import CompressedUniPoly
from tests.util import from_mont_array, to_mont_array

pytestmark = pytest.mark.gpu
P = F.P


def _chal(rng, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def _verify(coeffs, claim, label, pre=None):
    """SumcheckInstanceProof::verify: returns (final claim, challenges as Fr ints) with the transcript replayed."""
    t = TR.Blake2bTranscript(label)
    if pre:
        pre(t)
    e, rs = claim % P, []
    for c in coeffs:
        cp = CompressedUniPoly(from_mont_array(c))
        cp.append_to_transcript(t)
        r = F.challenge_to_fr(t.challenge_scalar_optimized())
        e = cp.eval_from_hint(e, r)
        rs.append(r)
    return e, rs, t


def test_mul_and_product_sumchecks_verify_at_full_size(ctx):
    from jolt_atlas_b200 import Blake2bTranscriptState, EvalKernel, MultilinearPolynomial, sumcheck_prove
    rng = np.random.default_rng(2024)
    m = 14
    w = _chal(rng, m)
    w_fr = from_mont_array(w)
    # Mul: claim = sum_x eq(w, x) a[x] b[x] on i8-range tensors
    a = rng.integers(-128, 128, size=1 << m, dtype=np.int32)
    b = rng.integers(-128, 128, size=1 << m, dtype=np.int32)
    ab = ORC.fr_binop(2, ORC.fr_from_i64(a), ORC.fr_from_i64(b))
    claim = ORC.evaluate(ab, w)
    t = Blake2bTranscriptState(b"mul")
    res = sumcheck_prove(ctx, EvalKernel.MUL, [MultilinearPolynomial.from_i32(ctx, a), MultilinearPolynomial.from_i32(ctx, b)], claim, t, eq_w=w)
    claim_i = from_mont_array(claim[None])[0]
    e, rs, tt = _verify(res["coeffs"], claim_i, b"mul", pre=lambda tr: tr.append_scalar(claim_i))
    fa, fb = from_mont_array(res["final_claims"])
    assert e == PL.eq_mle(w_fr, rs[::-1]) * fa % P * fb % P        # LowToHigh: round j binds the j-th lowest variable
    assert tt.state == t.state
    # product of 16 RA polynomials gathered from 16-entry tables
    d, K = 16, 16
    k = rng.integers(0, K, size=(d, 1 << m), dtype=np.uint32)
    tables = _chal(rng, d * K).reshape(d, K, 4)
    polys_host = [tables[i][k[i]] for i in range(d)]
    prod = polys_host[0]
    for i in range(1, d):
        prod = ORC.fr_binop(2, prod, np.ascontiguousarray(polys_host[i]))
    claim = ORC.evaluate(np.ascontiguousarray(prod), w)
    t = Blake2bTranscriptState(b"prod")
    res = sumcheck_prove(ctx, EvalKernel.PROD, [MultilinearPolynomial.from_fr(ctx, np.ascontiguousarray(p)) for p in polys_host], claim, t, eq_w=w)
    claim_i = from_mont_array(claim[None])[0]
    e, rs, tt = _verify(res["coeffs"], claim_i, b"prod", pre=lambda tr: tr.append_scalar(claim_i))
    fin = 1
    for x in from_mont_array(res["final_claims"]):
        fin = fin * x % P
    assert e == PL.eq_mle(w_fr, rs[::-1]) * fin % P
    assert tt.state == t.state


def test_rlc_commitment_is_combination_of_commitments(ctx):
    from jolt_atlas_b200 import SRS, MsmWidth, MultilinearPolynomial, OneHotAddresses, commit_one_hot_batches, msm_fr, msm_host
    rng = np.random.default_rng(7)
    K, tau = 16, 0x1234567890abcdef1122334455667788
    g = np.zeros(8, dtype=np.uint64)
    rq = (1 << 256) % F.Q
    g[:4] = F.to_limbs(rq); g[4:] = F.to_limbs(2 * rq % F.Q)
    ell = 18
    srs = SRS.generate(ctx, g, to_mont_array([tau])[0], 1 << ell).precompute()
    batches = []
    for d, log_t in ((16, 14), (4, 14), (16, 12), (4, 12)):
        k = rng.integers(0, K, size=(d, 1 << log_t), dtype=np.uint32)
        k[0, :7] = 0xFFFFFFFF
        batches.append(OneHotAddresses(ctx, k, K))
    coms = commit_one_hot_batches(ctx, srs, batches)
    # the batched upload gives the same resident lists as one upload per batch
    ks2 = []
    rng2 = np.random.default_rng(7)
    for d, log_t in ((16, 14), (4, 14), (16, 12), (4, 12)):
        k = rng2.integers(0, K, size=(d, 1 << log_t), dtype=np.uint32)
        k[0, :7] = 0xFFFFFFFF
        ks2.append(k)
    many = OneHotAddresses.upload_many(ctx, ks2, K)
    coms2 = commit_one_hot_batches(ctx, srs, many)
    for (a, ai), (b, bi) in zip(coms, coms2):
        assert np.array_equal(a, b) and np.array_equal(ai, bi)
    for h in many:
        h.free()
    n_poly = sum(b.d for b in batches)
    gammas = np.ascontiguousarray(rng.integers(0, 1 << 63, size=(n_poly, 4), dtype=np.uint64))
    gammas[:, 3] &= np.uint64((1 << 60) - 1)
    joint = MultilinearPolynomial.zeros(ctx, 1 << ell)
    o = 0
    for b in batches:
        joint.rlc_add_onehot(b, gammas[o:o + b.d])
        o += b.d
    lhs, linf = msm_fr(ctx, srs, joint)                              # commit(sum_i gamma_i P_i)
    pts = np.concatenate([c[0] for c in coms])
    assert not np.concatenate([c[1] for c in coms]).any()
    com_srs = SRS(ctx, pts)                                          # sum_i gamma_i * C_i as an MSM over the commitments
    rhs, rinf = msm_host(ctx, com_srs, gammas, MsmWidth.FR)
    assert not linf and not rinf and np.array_equal(lhs, rhs)
    for b in batches:
        b.free()
    joint.free(); com_srs.free(); srs.free()


def test_engine_error_paths(ctx):
    from jolt_atlas_b200 import (Blake2bTranscriptState, EvalKernel, InstanceKind, JoltAtlasError, MultilinearPolynomial,
                                 OneHotAddresses, batched_sumcheck_prove, sumcheck_prove)
    rng = np.random.default_rng(3)
    t = Blake2bTranscriptState(b"err")
    claim = _chal(rng, 1)[0]
    p = MultilinearPolynomial.random(ctx, 1 << 6, 1)
    with pytest.raises(JoltAtlasError):                               # eq point shorter than the number of rounds
        sumcheck_prove(ctx, EvalKernel.IDENT, [p], claim, t, eq_w=_chal(rng, 5))
    with pytest.raises(JoltAtlasError):                               # MUL needs two polynomials
        sumcheck_prove(ctx, EvalKernel.MUL, [p], claim, t, eq_w=_chal(rng, 6))
    q = MultilinearPolynomial.random(ctx, 1 << 5, 2)
    with pytest.raises(JoltAtlasError):                               # length mismatch inside one instance
        sumcheck_prove(ctx, EvalKernel.ADD, [p, q], claim, t, eq_w=_chal(rng, 6))
    with pytest.raises(JoltAtlasError):                               # address outside [0, K)
        OneHotAddresses(ctx, np.full((2, 8), 16, dtype=np.uint32), 16)
    good = rng.integers(0, 16, size=(3, 64), dtype=np.uint32)
    bad = good.copy(); bad[2, 63] = 16
    with pytest.raises(JoltAtlasError):                               # one bad batch fails the whole batched upload (device-side validation)
        OneHotAddresses.upload_many(ctx, [good, bad, good], 16)
    none_ok = good.copy(); none_ok[1, 5] = 0xFFFFFFFF                 # None entries are valid
    for h in OneHotAddresses.upload_many(ctx, [good, none_ok], 16):
        h.free()
    addr = OneHotAddresses(ctx, rng.integers(0, 16, size=(2, 8), dtype=np.uint32), 16)
    with pytest.raises(JoltAtlasError):                               # booleanity: r_cycle does not match T
        batched_sumcheck_prove(ctx, [{"kind": InstanceKind.BOOLEANITY, "tables": _chal(rng, 32).reshape(2, 16, 4), "addr": addr,
                                      "eq_w": _chal(rng, 4), "gammas": _chal(rng, 2), "r_address": _chal(rng, 4)}], t)
    assert t.state == Blake2bTranscriptState(b"err").state           # a failed call leaves the caller's transcript untouched
    addr.free(); p.free(); q.free()
