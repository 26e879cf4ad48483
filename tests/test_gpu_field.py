"""Device field arithmetic (csrc/fp.cuh) against the golden edge vectors of tests/golden/field.json and against the
Python big-int oracle on random operands, DIRECTLY through the ja_test_field_ops hook (not through bind / round kernels).
Reference semantics: joltworks/src/field/ark.rs:76-297, field/challenge/mont_ark_u128.rs:51-84,
field/challenge/macros.rs:274-286 (F * challenge), field/mod.rs:286-310 (delayed reduction).  Bit-exact."""
import json
import os
import random

import numpy as np
import pytest

from oracle.pyref import field as F
from tests.util import from_mont_array, to_mont_array

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "field.json")))
MUL, ADD, SUB, MUL_CH, FROM_I64, MUL_WIDE, NEG, SQR = range(8)


def _h(x):
    return int(x, 16)


def test_golden_mul_add_sub(ctx):
    rows = GOLD["mul"]
    a = to_mont_array([_h(r["a"]) for r in rows])
    b = to_mont_array([_h(r["b"]) for r in rows])
    assert from_mont_array(ctx.test_field_ops(MUL, a, b)) == [_h(r["ab"]) for r in rows]
    assert from_mont_array(ctx.test_field_ops(ADD, a, b)) == [_h(r["a_plus_b"]) for r in rows]
    assert from_mont_array(ctx.test_field_ops(SUB, a, b)) == [_h(r["a_minus_b"]) for r in rows]
    assert from_mont_array(ctx.test_field_ops(MUL_WIDE, a, b)) == [16 * _h(r["ab"]) % F.P for r in rows]
    assert from_mont_array(ctx.test_field_ops(SQR, a, a)) == [_h(r["a"]) ** 2 % F.P for r in rows]
    assert from_mont_array(ctx.test_field_ops(NEG, a, a)) == [(-_h(r["a"])) % F.P for r in rows]


def test_golden_montgomery_limbs_roundtrip(ctx):
    # x + 0 through the device returns the golden Montgomery limbs (canonical, < p)
    rows = GOLD["mont"]
    a = to_mont_array([_h(r["x"]) for r in rows])
    want = np.array([[_h(l) for l in r["limbs"]] for r in rows], dtype=np.uint64)
    assert np.array_equal(a, want)
    z = to_mont_array([0] * len(rows))
    assert np.array_equal(ctx.test_field_ops(ADD, a, z), want)


def test_golden_challenge_mul(ctx):
    rows = GOLD["challenge"]
    a = to_mont_array([_h(r["a"]) for r in rows])
    c = np.array([[_h(l) for l in r["limbs"]] for r in rows], dtype=np.uint64)
    assert from_mont_array(ctx.test_field_ops(MUL_CH, a, c)) == [_h(r["a_times_c"]) for r in rows]
    # F * challenge == F * Fr(challenge)  (transcripts/blake2b.rs:286-316): the challenge limbs ARE a valid Montgomery Fr
    assert np.array_equal(ctx.test_field_ops(MUL_CH, a, c), ctx.test_field_ops(MUL, a, c))


def test_golden_from_i64(ctx):
    rows = GOLD["from_i64"]
    a = np.zeros((len(rows), 4), dtype=np.uint64)
    a[:, 0] = np.array([r["v"] for r in rows], dtype=np.int64).view(np.uint64)
    assert from_mont_array(ctx.test_field_ops(FROM_I64, a, a)) == [_h(r["fr"]) for r in rows]


def test_random_and_edge_pairs_vs_bigint(ctx):
    rng = random.Random(0xF1E1D)
    edge = [0, 1, 2, F.P - 1, F.P - 2, F.R % F.P, (F.R * F.R) % F.P, (1 << 253), (1 << 128) - 1, (1 << 64), F.P >> 1]
    xs = [x for x in edge for _ in edge] + [rng.randrange(F.P) for _ in range(4096)]
    ys = [y for _ in edge for y in edge] + [rng.randrange(F.P) for _ in range(4096)]
    a, b = to_mont_array(xs), to_mont_array(ys)
    assert from_mont_array(ctx.test_field_ops(MUL, a, b)) == [x * y % F.P for x, y in zip(xs, ys)]
    assert from_mont_array(ctx.test_field_ops(ADD, a, b)) == [(x + y) % F.P for x, y in zip(xs, ys)]
    assert from_mont_array(ctx.test_field_ops(SUB, a, b)) == [(x - y) % F.P for x, y in zip(xs, ys)]
    assert from_mont_array(ctx.test_field_ops(MUL_WIDE, a, b)) == [16 * x * y % F.P for x, y in zip(xs, ys)]
    cs = [rng.getrandbits(128) & F.CHALLENGE_MASK for _ in xs]
    cs[:4] = [0, 1, F.CHALLENGE_MASK, 1 << 124]
    c = np.array([F.challenge_limbs(v) for v in cs], dtype=np.uint64)
    assert from_mont_array(ctx.test_field_ops(MUL_CH, a, c)) == [x * F.challenge_to_fr(v) % F.P for x, v in zip(xs, cs)]
