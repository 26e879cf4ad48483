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
