"""GPU parity of witness generation (ja_witness_fused: i64 accumulation, floor rebase, remainder, one-hot chunk lists, clamped
output) against the numpy restatement, bit-exact, and of everything built on the device-born addresses: the one-hot commitments
equal those of uploaded addresses, and a ps_shout state created from the witness equals one created from host indices."""
import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import witness as WT
from tests.util import to_mont_array

pytestmark = pytest.mark.gpu
TAU = 0x1234567890abcdef1122334455667788


@pytest.mark.parametrize("op,shape_a,shape_b,S", [(0, (64, 64), (64, 192), 14), (0, (16, 768), (768, 96), 14), (0, (5, 33), (33, 7), 7),
                                                   (1, (64, 256), (64, 256), 14), (2, (32, 128), (32, 128), 0), (3, (1, 100), (1, 100), 0)])
def test_fused_witness_matches_oracle(ctx, op, shape_a, shape_b, S):
    from jolt_atlas_b200 import FusedWitness, TensorI32
    rng = np.random.default_rng(op * 100 + S + shape_a[0])
    lim = 1 << (30 if op else 24)
    A = rng.integers(-lim, lim, size=shape_a, dtype=np.int64).astype(np.int32)
    B = rng.integers(-lim, lim, size=shape_b, dtype=np.int64).astype(np.int32)
    n = WT.accumulate(op, A, B).shape[0]
    T = 1 << (n - 1).bit_length()
    ta, tb = TensorI32(ctx, A), TensorI32(ctx, B)
    w = FusedWitness(ctx, op, ta, tb, S, T)
    got = w.to_host()
    want = WT.fused_witness(op, A, B, S, T)
    for g, x in zip(got, want):
        assert (g is None and x is None) or np.array_equal(g, x)
    w.free(); ta.free(); tb.free()


def test_device_born_addresses_commit_and_ps_shout(ctx):
    from jolt_atlas_b200 import SRS, FusedWitness, OneHotAddresses, TensorI32, commit_one_hot_batches
    from jolt_atlas_b200.api import PrefixSuffixShout
    from tests.test_oracle_psshout import _chal
    rng = np.random.default_rng(9)
    A = rng.integers(-128, 128, size=(16, 32), dtype=np.int32)
    B = rng.integers(-128, 128, size=(32, 16), dtype=np.int32)
    T, S = 256, 14
    ta, tb = TensorI32(ctx, A), TensorI32(ctx, B)
    w = FusedWitness(ctx, 0, ta, tb, S, T)
    idx, out, ck, rk = w.to_host()
    srs = SRS(ctx, ORC.srs_powers(to_mont_array([TAU])[0], 16 * T))
    up = [OneHotAddresses(ctx, ck, 16), OneHotAddresses(ctx, rk, 16)]
    a = commit_one_hot_batches(ctx, srs, [w.clamp, w.rem])
    b = commit_one_hot_batches(ctx, srs, up)
    for (x, xi), (y, yi) in zip(a, b):
        assert np.array_equal(x, y) and np.array_equal(xi, yi)
    r = _chal(rng, 8)
    p1, p2 = w.ps_shout(r), PrefixSuffixShout(ctx, idx, r)
    kinds = [1, 2, 3, 0, 0, 4]
    assert np.array_equal(p1.init_phase(0, None, kinds, 31), p2.init_phase(0, None, kinds, 31))
    v = _chal(rng, 256)
    assert np.array_equal(p1.init_phase(1, v, kinds, 31), p2.init_phase(1, v, kinds, 31))
    for o in (p1, p2, *up, srs, ta, tb):
        o.free()
    w.free()


def test_remainder_range_check_from_device_witness(ctx):
    """The remainder lookup indices born on the device (ja_psshout_from_witness_rem) drive IdentityRCProver's address rounds to the
    same proof as the oracle over the numpy witness' remainders."""
    from jolt_atlas_b200 import api as A
    from jolt_atlas_b200.workload import identity_rc_phases
    from oracle.pyref import witness as WT
    rng = np.random.default_rng(77)
    a = rng.integers(-128, 128, size=(16, 32), dtype=np.int32)
    b = rng.integers(-128, 128, size=(32, 16), dtype=np.int32)
    S, T = 14, 256
    ta, tb = A.TensorI32(ctx, a), A.TensorI32(ctx, b)
    w = A.FusedWitness(ctx, 0, ta, tb, S, T)
    acc = WT.accumulate(0, a, b)
    rem = (acc & ((1 << S) - 1)).astype(np.uint64)
    r = np.ascontiguousarray(rng.integers(1, 1 << 60, size=(8, 4), dtype=np.uint64))
    r[:, 3] &= np.uint64((1 << 58) - 1)
    dev, cpu = w.rem_shout(r, identity_rc_phases(S)), ORC.PsShout(rem, r, S, identity_rc_phases(S))
    td, tc = A.Blake2bTranscriptState(b"rem"), ORC.TranscriptState(b"rem")
    got, want = dev.prove_identity_rc(td), cpu.prove_identity_rc(tc, None)
    assert np.array_equal(got["coeffs"], want["coeffs"]) and td.state == tc.state
    dev.free(); cpu.free(); w.free(); ta.free(); tb.free()
