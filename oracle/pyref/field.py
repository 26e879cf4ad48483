"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python big-int twin). Never imported by the product path.

BN254 scalar field Fr / base field Fq arithmetic restated with Python ints, plus the
limb conventions the C ABI uses.

Parity status: **parity unpinned at the byte level** — the reference holds no golden vectors
for this path and cannot be compiled here (no Rust toolchain; arithmetic lives
in un-vendored `a16z/arkworks-algebra@76bb3a4`).  What pins this file: the public BN254
constants, the Montgomery constants of `ark_bn254::FrConfig` (R, R2, INV are functions of the
modulus), and the reference's own equivalence invariants re-run in tests/.

Follows:
  joltworks/src/field/ark.rs:16-298            (JoltField for ark_bn254::Fr: from_i32.., from_bytes)
  joltworks/src/field/challenge/mont_ark_u128.rs:51-92 (125-bit challenge stored as Montgomery limbs [0,0,lo,hi])
  joltworks/src/field/challenge/macros.rs:274-286      (F x Challenge = mul_hi_bigint_u128 == F * Fr(challenge))
"""
from __future__ import annotations

# BN254 (alt_bn128) public constants
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr modulus r
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq modulus q
R_BITS = 256
R = (1 << R_BITS) % P           # Montgomery R for Fr  (ark MontConfig::R)
R2 = (R * R) % P                # ark MontConfig::R2
R_INV = pow(1 << R_BITS, -1, P)
RQ = (1 << R_BITS) % Q
RQ_INV = pow(1 << R_BITS, -1, Q)
INV64 = (-pow(P, -1, 1 << 64)) % (1 << 64)     # ark MontConfig::INV  (-p^{-1} mod 2^64)
INV32 = (-pow(P, -1, 1 << 32)) % (1 << 32)
INVQ64 = (-pow(Q, -1, 1 << 64)) % (1 << 64)
INVQ32 = (-pow(Q, -1, 1 << 32)) % (1 << 32)
MASK64 = (1 << 64) - 1
CHALLENGE_MASK = ((1 << 128) - 1) >> 3          # mont_ark_u128.rs:55  value & (u128::MAX >> 3)


def to_limbs(x: int, n: int = 4) -> list[int]:
    return [(x >> (64 * i)) & MASK64 for i in range(n)]


def from_limbs(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


# ---- Fr <-> Montgomery limbs (the in-memory form of ark_bn254::Fr: BigInt<4>, value*R mod p) ----
def fr_to_mont(x: int) -> list[int]:
    return to_limbs((x % P) * R % P)


def fr_from_mont(limbs) -> int:
    return from_limbs(limbs) * R_INV % P


def fq_to_mont(x: int) -> list[int]:
    return to_limbs((x % Q) * RQ % Q)


def fq_from_mont(limbs) -> int:
    return from_limbs(limbs) * RQ_INV % Q


# ---- challenge type (MontU128Challenge) ----
def challenge_limbs(u128: int) -> list[int]:
    """mont_ark_u128.rs:51-63: mask to 125 bits, store as limbs [0, 0, lo, hi]."""
    v = u128 & CHALLENGE_MASK
    return [0, 0, v & MASK64, v >> 64]


def challenge_to_fr(u128: int) -> int:
    """mont_ark_u128.rs:79-84: limbs are *reinterpreted as the Montgomery representation*
    (from_bigint_unchecked), hence the field value is  (masked << 128) * R^-1 mod p."""
    return from_limbs(challenge_limbs(u128)) * R_INV % P


def fr_from_i(v: int) -> int:
    """field/ark.rs:125-162 from_i32/from_i64/from_i128: sign handled by field negation."""
    return v % P


def fr_from_le_bytes_mod_order(b: bytes) -> int:
    """field/ark.rs:240-243 from_bytes -> Fr::from_le_bytes_mod_order."""
    return int.from_bytes(b, "little") % P


def fr_inv(x: int) -> int:
    return pow(x, -1, P)


def fr_to_le_bytes(x: int) -> bytes:
    """ark-serialize of Fr: 32-byte little-endian canonical (non-Montgomery) integer."""
    return (x % P).to_bytes(32, "little")


def mul_pow_2(x: int, k: int) -> int:
    """field/mod.rs mul_pow_2: x * 2^k."""
    return x * pow(2, k, P) % P
