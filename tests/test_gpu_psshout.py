"""GPU parity of the T-sized passes of the prefix-suffix Shout prover (ja_psshout_*: init_phase / init_suffix_polys / init_Q /
init_log_t_rounds, joltworks/src/subprotocols/ps_shout/mod.rs:269-335,420-446) against the C++ oracle, bit-exact, over all 8
phases of a 64-bit clamp lookup (SaturationTable suffixes + the unary raf suffixes) and a 32-bit table with 4 bits per phase."""
import numpy as np
import pytest

from oracle import cpu as ORC
from tests.test_oracle_psshout import _chal, clamp_like_indices

pytestmark = pytest.mark.gpu
ONE, HAZ, HZML, HOML, IDENT = range(5)


@pytest.mark.parametrize("log_t,log_k,bound", [(4, 64, 31), (10, 64, 31), (14, 64, 31), (11, 64, 9), (9, 32, 9)])
def test_all_phases_match_oracle(ctx, log_t, log_k, bound):
    from jolt_atlas_b200.api import PrefixSuffixShout
    rng = np.random.default_rng(100 * log_t + log_k)
    T, phases = 1 << log_t, 8
    m = 1 << (log_k // phases)
    idx = clamp_like_indices(rng, T)
    if log_k < 64:
        idx = idx & np.uint64((1 << log_k) - 1)
    r = _chal(rng, log_t)
    kinds = [HAZ, HZML, HOML, ONE, ONE, IDENT]                 # SaturationTable read-checking suffixes, then the raf's [One, Identity]
    dev, cpu = PrefixSuffixShout(ctx, idx, r, log_k, phases), ORC.PsShout(idx, r, log_k, phases)
    vs = []
    for phase in range(phases):
        v_prev = vs[-1] if phase else None
        got = dev.init_phase(phase, v_prev, kinds, bound)
        want = cpu.init_phase(phase, v_prev, kinds, bound)
        assert np.array_equal(got, want), phase
        vs.append(_chal(rng, m))                               # stands for ExpandingTable after the phase's rounds: any m field elements
    v_all = np.concatenate(vs)
    ra = dev.materialize_ra(v_all)
    assert np.array_equal(ra.to_host(), cpu.materialize_ra(v_all))
    ra.free(); dev.free(); cpu.free()


def test_phase_order_is_enforced(ctx):
    from jolt_atlas_b200 import JoltAtlasError
    from jolt_atlas_b200.api import PrefixSuffixShout
    rng = np.random.default_rng(1)
    dev = PrefixSuffixShout(ctx, clamp_like_indices(rng, 16), _chal(rng, 4))
    with pytest.raises(JoltAtlasError):
        dev.init_phase(1, _chal(rng, 256), [ONE], 31)          # phase 0 first
    dev.free()


@pytest.mark.parametrize("log_t,bound,given_claim", [(5, 31, True), (10, 31, False), (14, 31, True), (9, 9, False)])
def test_address_rounds_match_oracle(ctx, log_t, bound, given_claim):
    """ja_psshout_prove_address (phase passes on the device, regrouped 256-entry rounds on the host) against the oracle's per-b
    restatement of ps_shout/mod.rs:337-418, :491-560: round polynomials, challenges, transcript, expanding tables, val, raf_val,
    running claim - then the cycle rounds over ra * (val + raf_val) against the oracle's scaled IDENT sumcheck."""
    from jolt_atlas_b200 import api as A
    from tests.test_oracle_psshout_address import P, clamp_entry, lookup_indices
    from tests.util import from_mont_array, to_mont_array
    rng = np.random.default_rng(7 * log_t + bound)
    T = 1 << log_t
    idx = lookup_indices(rng, T, bound)
    r = _chal(rng, log_t)
    gamma = _chal(rng, 1)[0]
    # the true input claim rv(r) + gamma * operand(r), brute force
    eq = from_mont_array(ORC.eq_evals(r))
    g_i = from_mont_array(gamma.reshape(1, 4))[0]
    signed = [int(x) - (1 << 64) if int(x) >> 63 else int(x) for x in idx]
    true_claim = (sum(e * clamp_entry(int(k), bound) for e, k in zip(eq, idx)) + g_i * sum(e * s for e, s in zip(eq, signed))) % P
    claim = to_mont_array([true_claim])[0]
    dev, cpu = A.PrefixSuffixShout(ctx, idx, r), ORC.PsShout(idx, r)
    td, tc = A.Blake2bTranscriptState(b"ps_addr"), ORC.TranscriptState(b"ps_addr")
    got = dev.prove_address(td, gamma, bound, claim if given_claim else None)
    want = cpu.prove_address(tc, gamma, claim, bound)
    assert np.array_equal(got["input_claim"], claim)               # the sum the prover derives from its phase-0 tables IS the claim
    assert np.array_equal(got["ncoeffs"], want["ncoeffs"])
    assert np.array_equal(got["coeffs"], want["coeffs"])
    assert np.array_equal(got["challenges"], want["challenges"])
    assert td.state == tc.state and td.n_rounds == tc.n_rounds
    assert np.array_equal(dev.tables(), want["v"])
    for k in ("val", "raf_val", "claim"):
        assert np.array_equal(got[k], want[k]), k
    # cycle rounds: gruen_poly_deg_2(eval_at_0 * (val + raf_val), claim) (mod.rs:464-488) == IDENT over ra * (val + raf_val)
    scale = ORC.fr_binop(0, got["val"].reshape(1, 4), got["raf_val"].reshape(1, 4))[0]
    ra = dev.materialize_ra(scale=scale)
    ra_cpu = cpu.materialize_ra(want["v"].reshape(-1, 4))
    ra_scaled = ORC.fr_binop(2, ra_cpu, np.broadcast_to(scale, ra_cpu.shape).copy())
    assert np.array_equal(ra.to_host(), ra_scaled)
    rd = A.sumcheck_prove(ctx, A.EvalKernel.IDENT, [ra], got["claim"], td, eq_w=r)
    rc = ORC.sumcheck_prove_st(0, 6, ra_scaled[None], r, want["claim"], tc)
    assert all(np.array_equal(a, b) for a, b in zip(rd["coeffs"], rc["coeffs"]))
    assert td.state == tc.state
    dev.free(); cpu.free()


@pytest.mark.parametrize("log_t,log_k,phases,given_claim", [(5, 14, 7, True), (12, 14, 7, False), (14, 16, 4, False), (9, 8, 2, True)])
def test_identity_range_check_rounds_match_oracle(ctx, log_t, log_k, phases, given_claim):
    """ja_psshout_prove_identity_rc against the oracle's per-b IdentityRCProver (identity_range_check.rs:140-325), then the cycle
    rounds over ra * raf_val against the oracle's IDENT sumcheck (gruen_poly_deg_2(eval_at_0 * raf_val, claim), :253-283)."""
    from jolt_atlas_b200 import api as A
    from tests.test_oracle_psshout_address import P
    from tests.util import from_mont_array, to_mont_array
    rng = np.random.default_rng(13 * log_t + log_k)
    T = 1 << log_t
    idx = rng.integers(0, 1 << log_k, size=T, dtype=np.uint64)
    r = _chal(rng, log_t)
    eq = from_mont_array(ORC.eq_evals(r))
    claim = to_mont_array([sum(e * int(k) for e, k in zip(eq, idx)) % P])[0]
    dev, cpu = A.PrefixSuffixShout(ctx, idx, r, log_k, phases), ORC.PsShout(idx, r, log_k, phases)
    td, tc = A.Blake2bTranscriptState(b"identity_rc"), ORC.TranscriptState(b"identity_rc")
    got = dev.prove_identity_rc(td, claim if given_claim else None)
    want = cpu.prove_identity_rc(tc, claim)
    assert np.array_equal(got["input_claim"], claim)
    for k in ("ncoeffs", "coeffs", "challenges", "raf_val", "claim"):
        assert np.array_equal(got[k], want[k]), k
    assert td.state == tc.state and td.n_rounds == tc.n_rounds
    assert np.array_equal(dev.tables(), want["v"])
    ra = dev.materialize_ra(scale=got["raf_val"])
    ra_cpu = cpu.materialize_ra(want["v"].reshape(-1, 4))
    ra_scaled = ORC.fr_binop(2, ra_cpu, np.broadcast_to(want["raf_val"], ra_cpu.shape).copy())
    assert np.array_equal(ra.to_host(), ra_scaled)
    rd = A.sumcheck_prove(ctx, A.EvalKernel.IDENT, [ra], got["claim"], td, eq_w=r)
    rc = ORC.sumcheck_prove_st(0, 6, ra_scaled[None], r, want["claim"], tc)
    assert all(np.array_equal(a, b) for a, b in zip(rd["coeffs"], rc["coeffs"]))
    assert td.state == tc.state
    dev.free(); cpu.free()


def test_identity_rc_phase_count_is_free(ctx):
    """The device prover with 2 phases of 7 bits emits the transcript of the reference's 7 phases of 2 bits (LOG_K = 14): round
    polynomials, challenges, raf_val, running claim and the materialised ra are identical - the chunking is a prover-side choice."""
    from jolt_atlas_b200 import api as A
    from jolt_atlas_b200.workload import device_rc_phases, identity_rc_phases
    rng = np.random.default_rng(5)
    log_t, log_k = 11, 14
    assert (device_rc_phases(log_k), identity_rc_phases(log_k)) == (2, 7)
    idx = rng.integers(0, 1 << log_k, size=1 << log_t, dtype=np.uint64)
    r = _chal(rng, log_t)
    dev, cpu = A.PrefixSuffixShout(ctx, idx, r, log_k, 2), ORC.PsShout(idx, r, log_k, 7)
    td, tc = A.Blake2bTranscriptState(b"rc"), ORC.TranscriptState(b"rc")
    got, want = dev.prove_identity_rc(td), cpu.prove_identity_rc(tc, None)
    for k in ("ncoeffs", "coeffs", "challenges", "raf_val", "claim"):
        assert np.array_equal(got[k], want[k]), k
    assert td.state == tc.state
    ra = dev.materialize_ra()
    assert np.array_equal(ra.to_host(), cpu.materialize_ra(want["v"].reshape(-1, 4)))
    ra.free(); dev.free(); cpu.free()


@pytest.mark.parametrize("log_t,bound,lo,hi", [(10, 31, -64, 64), (12, 31, -(1 << 20), 1 << 20), (9, 31, -(1 << 31), 1 << 31), (8, 31, -1, 1),
                                               (11, 31, 0, 1000), (10, 31, -5000, 0), (10, 9, -300, 300), (9, 31, -(1 << 40), 1 << 40),
                                               (2, 31, -5, 5), (1, 31, -(1 << 40), 1 << 40)])
def test_sign_extension_phases_match_oracle(ctx, log_t, bound, lo, hi):
    """Small signed lookup values: ja_psshout_prove_address builds the suffix polynomials of the sign-extension phases on the host
    from the four class sums of the phase-0 pass (no T-sized pass for them) - the proof must stay the oracle's, which runs all eight
    passes.  Ranges: a few bits, 20 bits, the full i32 range (sig = BOUND), {-1, 0}, one-sided, BOUND = 9 with values beyond it, and
    values wider than BOUND (fast path off).  JA_PS_NO_SKIP is the library's switch for the all-passes form."""
    from jolt_atlas_b200 import api as A
    rng = np.random.default_rng(31 * log_t + (hi - lo) % 1000)
    T = 1 << log_t
    idx = rng.integers(lo, hi, size=T).astype(np.int64)
    idx[0], idx[1] = lo, hi - 1
    idx = idx.view(np.uint64)
    r = _chal(rng, log_t)
    gamma = _chal(rng, 1)[0]
    dev, cpu = A.PrefixSuffixShout(ctx, idx, r), ORC.PsShout(idx, r)
    td, tc = A.Blake2bTranscriptState(b"ps_sign"), ORC.TranscriptState(b"ps_sign")
    got = dev.prove_address(td, gamma, bound)
    want = cpu.prove_address(tc, gamma, None, bound)
    for k in ("ncoeffs", "coeffs", "challenges", "val", "raf_val", "claim"):
        assert np.array_equal(got[k], want[k]), k
    assert td.state == tc.state and td.n_rounds == tc.n_rounds
    assert np.array_equal(dev.tables(), want["v"])
    scale = ORC.fr_binop(0, got["val"].reshape(1, 4), got["raf_val"].reshape(1, 4))[0]
    ra = dev.materialize_ra(scale=scale)
    ra_cpu = cpu.materialize_ra(want["v"].reshape(-1, 4))
    assert np.array_equal(ra.to_host(), ORC.fr_binop(2, ra_cpu, np.broadcast_to(scale, ra_cpu.shape).copy()))
    ra.free(); dev.free(); cpu.free()
