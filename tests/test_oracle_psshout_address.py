"""CPU-only: the oracle's restatement of the 64 ADDRESS rounds of the prefix-suffix read-raf sumcheck (oracle/cpp/psshout.hpp
PsReadRaf; joltworks/src/subprotocols/ps_shout/mod.rs:337-418, :491-560) is pinned the way the reference's own test pins its prover
(ps_shout/unary.rs `test_read_raf_sumcheck`: prove, then Sumcheck::verify with ReadRafSumcheckVerifier::expected_output_claim,
mod.rs:601-620): every round polynomial must satisfy s(0) + s(1) = claim, and after the LOG_K address rounds the running claim must
equal  sum_j eq(r_cycle, j) * ra(r_address, j) * (val + gamma * raf)  with val = ClampBoundedTable::evaluate_mle(r_address)
(clamp.rs:140-192) and raf = SignedIdentityPoly::evaluate(r_address) (signed_identity_poly.rs:43-59) - two closed forms that share
no code with the prefix/suffix machinery.  The input claim is the brute-force  rv(r_cycle) + gamma * operand(r_cycle)."""
import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from tests.util import from_mont_array, to_mont_array

P = F.P
XLEN = 64


def clamp_entry(index: int, bound: int) -> int:
    val = index - (1 << XLEN) if index >> (XLEN - 1) else index
    return max(-(1 << bound), min(val, (1 << bound) - 1))


def lookup_indices(rng, T: int, bound: int) -> np.ndarray:
    small = rng.integers(-(1 << 12), 1 << 12, size=T)
    edge = rng.integers(-(1 << (bound + 1)), 1 << (bound + 1), size=T)
    wild = rng.integers(-(1 << 62), 1 << 62, size=T)
    pick = rng.integers(0, 4, size=T)
    v = np.where(pick == 0, wild, np.where(pick == 1, edge, small)).astype(np.int64)
    v[:4] = [0, -1, (1 << bound) - 1, -(1 << bound)]
    return v.view(np.uint64)


def run_case(seed: int, log_t: int, bound: int):
    rng = np.random.default_rng(seed)
    T = 1 << log_t
    idx = lookup_indices(rng, T, bound)
    r_cycle = to_mont_array([int(x) for x in rng.integers(1, 1 << 62, size=log_t)])
    gamma_i = int(rng.integers(1, 1 << 62)) * int(rng.integers(1, 1 << 62)) % P
    eq = from_mont_array(ORC.eq_evals(r_cycle))
    signed = [int(x) - (1 << 64) if int(x) >> 63 else int(x) for x in idx]
    rv = sum(e * clamp_entry(int(k), bound) for e, k in zip(eq, idx)) % P
    operand = sum(e * s for e, s in zip(eq, signed)) % P
    claim = (rv + gamma_i * operand) % P
    t = ORC.TranscriptState(b"ps_shout_test")
    ps = ORC.PsShout(idx, r_cycle)
    out = ps.prove_address(t, to_mont_array([gamma_i])[0], to_mont_array([claim])[0], bound)
    ps.free()
    return dict(idx=idx, eq=eq, gamma=gamma_i, claim=claim, out=out, t=t)


@pytest.mark.parametrize("seed,log_t,bound", [(1, 5, 31), (2, 7, 31), (3, 6, 9)])
def test_address_rounds_verify(seed, log_t, bound):
    c = run_case(seed, log_t, bound)
    out, claim = c["out"], c["claim"]
    ch = from_mont_array(out["challenges"])
    co = [from_mont_array(out["coeffs"][i]) for i in range(XLEN)]
    for j in range(XLEN):
        assert int(out["ncoeffs"][j]) == 2
        c0, c2 = co[j]
        c1 = (claim - 2 * c0 - c2) % P                      # the verifier's decompression: s(0) + s(1) = claim
        claim = (c0 + c1 * ch[j] + c2 * ch[j] * ch[j]) % P
    assert claim == from_mont_array(out["claim"].reshape(1, 4))[0]
    # expected_output_claim before the cycle rounds
    r_addr = out["challenges"]
    val = from_mont_array(ORC.clamp_evaluate_mle(r_addr, XLEN, bound).reshape(1, 4))[0]
    raf = from_mont_array(ORC.signed_identity_evaluate(r_addr, XLEN).reshape(1, 4))[0]
    assert val == from_mont_array(out["val"].reshape(1, 4))[0]
    assert c["gamma"] * raf % P == from_mont_array(out["raf_val"].reshape(1, 4))[0]
    v = [from_mont_array(out["v"][ph]) for ph in range(8)]
    total = 0
    for e, k in zip(c["eq"], c["idx"]):
        ra = 1
        for ph in range(8):
            ra = ra * v[ph][(int(k) >> (8 * (7 - ph))) & 255] % P
        total += e * ra
    assert claim == total % P * ((val + c["gamma"] * raf) % P) % P


def test_derived_input_claim_is_the_true_claim():
    """claim = None: the oracle derives s(0) + s(1) from its phase-0 tables; it must be rv(r) + gamma * operand(r) (same proof)."""
    c = run_case(9, 6, 31)
    rng = np.random.default_rng(9)
    idx = lookup_indices(rng, 64, 31)
    r_cycle = to_mont_array([int(x) for x in rng.integers(1, 1 << 62, size=6)])
    ps = ORC.PsShout(idx, r_cycle)
    t = ORC.TranscriptState(b"ps_shout_test")
    out = ps.prove_address(t, to_mont_array([c["gamma"]])[0], None, 31)
    ps.free()
    assert np.array_equal(out["coeffs"], c["out"]["coeffs"]) and t.state == c["t"].state


def test_expanding_tables_match_the_challenges():
    c = run_case(5, 4, 31)
    out = c["out"]
    for ph in range(8):
        assert np.array_equal(out["v"][ph], ORC.expanding_table_h2l(out["challenges"][8 * ph: 8 * ph + 8]))


@pytest.mark.parametrize("seed,log_t,log_k,phases", [(1, 5, 14, 7), (2, 7, 16, 4), (3, 4, 8, 2)])
def test_identity_range_check_rounds_verify(seed, log_t, log_k, phases):
    """IdentityRCProver (identity_range_check.rs:140-325) pinned by IdentityRCVerifier::expected_output_claim (:369-389): after the
    LOG_K address rounds the claim is  sum_j eq(r, j) * ra(r_address, j) * IdentityPolynomial(r_address); the input claim is the
    brute-force MLE of the remainders at r.  Phases per IdentityRCProvider::phases (:416-431)."""
    rng = np.random.default_rng(seed)
    T = 1 << log_t
    idx = rng.integers(0, 1 << log_k, size=T, dtype=np.uint64)
    idx[:2] = [0, (1 << log_k) - 1]
    r_cycle = to_mont_array([int(x) for x in rng.integers(1, 1 << 62, size=log_t)])
    eq = from_mont_array(ORC.eq_evals(r_cycle))
    claim = sum(e * int(k) for e, k in zip(eq, idx)) % P
    t = ORC.TranscriptState(b"identity_rc")
    ps = ORC.PsShout(idx, r_cycle, log_k, phases)
    out = ps.prove_identity_rc(t, to_mont_array([claim])[0])
    ps.free()
    ps2, t2 = ORC.PsShout(idx, r_cycle, log_k, phases), ORC.TranscriptState(b"identity_rc")
    out2 = ps2.prove_identity_rc(t2, None)                          # derived claim == true claim: same proof
    ps2.free()
    assert np.array_equal(out["coeffs"], out2["coeffs"]) and t.state == t2.state
    ch = from_mont_array(out["challenges"])
    for j in range(log_k):
        c0, c2 = from_mont_array(out["coeffs"][j])
        c1 = (claim - 2 * c0 - c2) % P
        claim = (c0 + c1 * ch[j] + c2 * ch[j] * ch[j]) % P
    assert claim == from_mont_array(out["claim"].reshape(1, 4))[0]
    ident = sum(ch[i] << (log_k - 1 - i) for i in range(log_k)) % P          # IdentityPolynomial::evaluate, big-endian point
    assert ident == from_mont_array(out["raf_val"].reshape(1, 4))[0]
    log_m = log_k // phases
    v = [from_mont_array(out["v"][ph]) for ph in range(phases)]
    total = 0
    for e, k in zip(eq, idx):
        ra = 1
        for ph in range(phases):
            ra = ra * v[ph][(int(k) >> (log_m * (phases - 1 - ph))) & ((1 << log_m) - 1)] % P
        total += e * ra
    assert claim == total % P * ident % P
