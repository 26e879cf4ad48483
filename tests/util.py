"""Test helpers: int <-> Montgomery-limb numpy arrays (uses the oracle's constants; tests only)."""
import random

import numpy as np

from oracle.pyref import field as F


def to_mont_array(xs) -> np.ndarray:
    out = np.empty((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        v = (x % F.P) * F.R % F.P
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & F.MASK64
    return out


def from_mont_array(a: np.ndarray):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in a:
        v = 0
        for k in range(4):
            v |= int(row[k]) << (64 * k)
        assert v < F.P, "non-canonical Montgomery limbs (>= p)"
        out.append(v * F.R_INV % F.P)
    return out


def challenge_array(c_u128: int) -> np.ndarray:
    return np.array(F.challenge_limbs(c_u128), dtype=np.uint64)


def rand_challenge(rng: random.Random) -> int:
    return rng.getrandbits(128) & F.CHALLENGE_MASK


def rand_fr(rng: random.Random, n: int):
    return [rng.randrange(F.P) for _ in range(n)]
