"""CPU-only: the oracle's suffix MLEs (restated as the reference's bit loops, oracle/cpp/psshout.hpp) are pinned by the property
that DEFINES a prefix-suffix decomposition (joltworks/src/lookup_tables/clamp.rs:88-130 `combine`, checked there by the reference's
own property tests): with the prefix part of an index fixed to BOOLEAN bits, combine(prefix indicators, suffix MLEs) must equal
ClampBoundedTable::materialize_entry (clamp.rs:145-157) of the whole index — for every split point the 8 phases use.  The prefix
values on boolean inputs are the plain indicators their docs state (clamp.rs:47-53): higher-all-zero, higher-all-one, lower word,
msb.  And the passes themselves (init_phase / init_Q / init_log_t_rounds) are compared with a direct Python big-int loop."""
import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from tests.util import from_mont_array, to_mont_array

P = F.P
ONE, HAZ, HZML, HOML, IDENT = range(5)


def clamp_entry(index: int, xlen: int, bound: int, symmetric: bool) -> int:
    """ClampBoundedTable::materialize_entry (clamp.rs:145-157), as a signed value."""
    val = index - (1 << xlen) if index >> (xlen - 1) else index
    lower = -(1 << bound) if symmetric else 0
    return max(lower, min(val, (1 << bound) - 1))


def combine(prefix_bits: int, prefix_len: int, suffix_bits: int, suffix_len: int, xlen: int, bound: int, symmetric: bool) -> int:
    """clamp.rs:103-129 with the prefix MLEs evaluated on boolean prefix bits."""
    const_upper = (1 << bound) - 1
    lower_coeff = 2 * const_upper + 1 if symmetric else const_upper
    # bits of the prefix by significance: prefix bit i (from the top) has global index i
    hi_bits = [(prefix_bits >> (prefix_len - 1 - i)) & 1 for i in range(prefix_len)]
    bound_index = xlen - bound - 1
    pre_haz = int(all(b == 0 for i, b in enumerate(hi_bits) if i <= bound_index))
    pre_hao = int(symmetric and all(b == 1 for i, b in enumerate(hi_bits) if i <= bound_index))
    pre_lw = sum(b << (xlen - 1 - i) for i, b in enumerate(hi_bits) if i > bound_index)
    pre_msb = hi_bits[0] if prefix_len else 0
    suf = [ORC.suffix_mle(k, suffix_bits, suffix_len, xlen, bound) for k in (HAZ, HZML, HOML if symmetric else ONE, ONE)]
    suf_haz, suf_hzml, suf_homl, suf_one = suf
    return (suf_one * const_upper - pre_msb * suf_one * lower_coeff
            + pre_haz * (suf_hzml + pre_lw * suf_one - suf_haz * const_upper)
            + pre_hao * (suf_homl + pre_lw * suf_one))


@pytest.mark.parametrize("xlen,bound,symmetric", [(64, 31, True), (64, 9, True), (64, 20, False), (32, 9, True)])
def test_suffix_mles_recompose_the_clamp_table(xlen, bound, symmetric):
    rng = np.random.default_rng(xlen * 100 + bound)
    log_m = xlen // 8
    edge = [0, 1, (1 << bound) - 1, 1 << bound, (1 << bound) + 5, (1 << (xlen - 1)) - 1, 1 << (xlen - 1), (1 << xlen) - 1,
            (1 << xlen) - (1 << bound), (1 << xlen) - (1 << bound) - 1, (1 << xlen) - 2]
    idxs = edge + [int(x) & ((1 << xlen) - 1) for x in rng.integers(0, 1 << 63, size=300, dtype=np.uint64) * 2 + 1] + \
        [int(x) % (1 << (bound + 2)) for x in rng.integers(0, 1 << 62, size=100)] + \
        [((1 << xlen) - 1 - int(x) % (1 << (bound + 2))) for x in rng.integers(0, 1 << 62, size=100)]
    for index in idxs:
        want = clamp_entry(index, xlen, bound, symmetric)
        for phase in range(8):
            suffix_len = (8 - 1 - phase) * log_m
            prefix_len = xlen - suffix_len
            got = combine(index >> suffix_len, prefix_len, index & ((1 << suffix_len) - 1), suffix_len, xlen, bound, symmetric)
            assert got == want, (hex(index), phase)


def _chal(rng, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def clamp_like_indices(rng, T):
    """64-bit two's-complement accumulations as the clamp lookups see them: mostly small magnitudes, some saturating."""
    small = rng.integers(-(1 << 20), 1 << 20, size=T)
    big = rng.integers(-(1 << 40), 1 << 40, size=T)
    pick = rng.integers(0, 8, size=T) == 0
    return np.where(pick, big, small).astype(np.int64).view(np.uint64)


def test_passes_match_direct_loops():
    rng = np.random.default_rng(5)
    log_t, log_k, phases, bound = 6, 64, 8, 31
    T, m = 1 << log_t, 256
    idx = clamp_like_indices(rng, T)
    r = _chal(rng, log_t)
    kinds = [HAZ, HZML, HOML, ONE, ONE, IDENT]
    ps = ORC.PsShout(idx, r, log_k, phases)
    u = from_mont_array(ORC.eq_evals(r))
    vs = []
    for phase in range(phases):
        v_prev = vs[-1] if phase else None
        Q = ps.init_phase(phase, to_mont_array(v_prev) if v_prev is not None else None, kinds, bound)
        if phase:
            for j in range(T):
                u[j] = u[j] * v_prev[(int(idx[j]) >> ((phases - phase) * 8)) & 255] % P
        suffix_len = (phases - 1 - phase) * 8
        want = [[0] * m for _ in kinds]
        for j in range(T):
            k = int(idx[j])
            y, sb = (k >> suffix_len) & 255, k & ((1 << suffix_len) - 1)
            for s, kind in enumerate(kinds):
                want[s][y] = (want[s][y] + u[j] * ORC.suffix_mle(kind, sb, suffix_len, 64, bound)) % P
        for s in range(len(kinds)):
            assert from_mont_array(Q[s]) == want[s], (phase, s)
        vs.append([int(x) for x in rng.integers(1, 1 << 62, size=m)])
    ra = from_mont_array(ps.materialize_ra(to_mont_array([x for v in vs for x in v])))
    for j in range(T):
        acc = 1
        for phase in range(phases):
            acc = acc * vs[phase][(int(idx[j]) >> ((phases - 1 - phase) * 8)) & 255] % P
        assert ra[j] == acc
    ps.free()
