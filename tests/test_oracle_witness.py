"""CPU-only: the oracle's witness generation satisfies the identities the reference's fused-rescale argument rests on
(jolt-atlas-core/src/onnx_proof/fused_rebase.rs: acc = q * 2^S + R, 0 <= R < 2^S; clamp_lookups: output = SatClamp(q)) and the
one-hot chunks recompose the 64-bit lookup index / the remainder (joltworks/src/config.rs:75-77)."""
import numpy as np
import pytest

from oracle.pyref import witness as WT


@pytest.mark.parametrize("op,shape_a,shape_b,S", [(0, (8, 16), (16, 4), 14), (0, (5, 33), (33, 7), 7), (1, (3, 11), (3, 11), 14), (2, (4, 4), (4, 4), 0), (3, (2, 9), (2, 9), 0)])
def test_identities(op, shape_a, shape_b, S):
    rng = np.random.default_rng(op * 10 + S)
    big = 1 << 30
    A = rng.integers(-big, big, size=shape_a, dtype=np.int64).astype(np.int32)
    B = rng.integers(-big, big, size=shape_b, dtype=np.int64).astype(np.int32)
    acc = [int(x) for x in WT.accumulate(op, A, B)]
    T = 1 << (len(acc) - 1).bit_length()
    idx, out, ck, rk = WT.fused_witness(op, A, B, S, T)
    d_rem = (S + 3) // 4
    for t in range(T):
        a = acc[t] if t < len(acc) else 0
        q = int(idx[t]) - (1 << 64) if int(idx[t]) >> 63 else int(idx[t])
        r = sum(int(rk[d, t]) << (4 * (d_rem - 1 - d)) for d in range(d_rem)) if d_rem else 0
        assert a == q * (1 << S) + r and 0 <= r < (1 << S) or (S == 0 and a == q)
        assert sum(int(ck[d, t]) << (4 * (15 - d)) for d in range(16)) == int(idx[t])
        assert int(out[t]) == max(-(1 << 31), min(q, (1 << 31) - 1))


def test_workload_host_witness_is_the_oracle_witness():
    """workload.host_witness (the numpy synthesis of a node's committed polynomials, exact f64 BLAS product for the einsum) equals the
    oracle's integer restatement for the three fused operators; phase counts of the remainder range check."""
    from jolt_atlas_b200 import workload as W
    rng = np.random.default_rng(3)
    for spec, A, B in ((W.NodeSpec("einsum", 8, 12, 64, 16), rng.integers(-128, 128, size=(12, 64), dtype=np.int32), rng.integers(-128, 128, size=(64, 16), dtype=np.int32)),
                       (W.NodeSpec("mul", 7), rng.integers(-128, 128, size=100, dtype=np.int32), rng.integers(-128, 128, size=100, dtype=np.int32)),
                       (W.NodeSpec("add", 7), rng.integers(-128, 128, size=128, dtype=np.int32), rng.integers(-128, 128, size=128, dtype=np.int32))):
        T = 1 << spec.log_t
        op, S = W.witness_op(spec)
        idx, rem, rows = W.host_witness(spec, A, B, T)
        a2, b2 = (A, B) if A.ndim == 2 else (A.reshape(1, -1), B.reshape(1, -1))
        widx, _, ck, rk = WT.fused_witness(op, a2, b2, S, T)
        assert np.array_equal(idx, widx) and np.array_equal(rows[:16], ck)
        assert (rk is None and rows.shape[0] == 16) or np.array_equal(rows[16:], rk)
        acc = WT.accumulate(op, a2, b2)
        assert np.array_equal(rem[: acc.shape[0]], (acc & ((1 << S) - 1)).astype(np.uint64))
    assert (W.identity_rc_phases(14), W.device_rc_phases(14)) == (7, 2)
    assert (W.identity_rc_phases(16), W.device_rc_phases(16)) == (4, 2)
    assert (W.identity_rc_phases(2), W.device_rc_phases(8)) == (1, 1)
