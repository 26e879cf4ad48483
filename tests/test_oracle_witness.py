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
