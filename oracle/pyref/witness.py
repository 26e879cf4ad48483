"""TEST INFRASTRUCTURE ONLY — numpy restatement (exact integer arithmetic) of the witness generation of a fused node
(jolt-atlas-core/src/onnx_proof/witness.rs:142-214).  Parity unpinned against the Rust prover (no toolchain, no KATs); the
tests pin it by the identities the reference itself relies on: acc == quotient * 2^S + remainder with 0 <= remainder < 2^S
(fused_rebase.rs), the chunks recompose the index (joltworks/src/config.rs:75-77), output == clamp_to_i32(quotient)."""
from __future__ import annotations

import numpy as np


def accumulate(op: int, A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """einsum_acc_i64 for mk,kn->mn (atlas-onnx-tracer/src/ops/einsum.rs:248-258), mul_acc_i64 (ops/mul.rs:51-60), left +- right
    (sat_binop_intermediate): i64, flattened row-major."""
    a, b = A.astype(np.int64), B.astype(np.int64)
    if op == 0:
        # exact in f64 (BLAS) whenever every partial sum stays below 2^53; the integer product otherwise
        bound = int(np.abs(a).max(initial=0)) * int(np.abs(b).max(initial=0)) * a.shape[1]
        if bound < (1 << 52):
            return (A.astype(np.float64) @ B.astype(np.float64)).astype(np.int64).reshape(-1)
        return (a @ b).reshape(-1)
    if op == 1:
        return (a * b).reshape(-1)
    return (a + b).reshape(-1) if op == 2 else (a - b).reshape(-1)


def fused_witness(op: int, A: np.ndarray, B: np.ndarray, scale_bits: int, T: int):
    """-> (lookup indices (T,) u64, clamped output (T,) i32, ClampRaD chunks (16, T) u32, RescaleRemainderRaD chunks (ceil(S/4), T) u32)."""
    acc = accumulate(op, A, B)
    pad = np.zeros(T, dtype=np.int64)                       # padded_next_power_of_two (zeros)
    pad[: acc.shape[0]] = acc
    q = pad >> scale_bits                                    # floor_rebase_i64: div_euclid(2^S) (ops/mod.rs:224-232)
    r = pad & ((1 << scale_bits) - 1)                        # rebase_remainder_i32: rem_euclid(2^S) (ops/mod.rs:237-249)
    idx = q.view(np.uint64)                                  # LookupBits::new(v as u64, 64) (clamp_lookups/mod.rs:245-252)
    clamp_k = np.stack([((idx >> np.uint64(4 * (15 - d))) & np.uint64(15)).astype(np.uint32) for d in range(16)])   # config.rs:75-77
    d_rem = (scale_bits + 3) // 4
    rem_k = np.stack([((r >> (4 * (d_rem - 1 - d))) & 15).astype(np.uint32) for d in range(d_rem)]) if d_rem else None
    out = np.clip(q, -(1 << 31), (1 << 31) - 1).astype(np.int32)
    return idx, out, clamp_k, rem_k
