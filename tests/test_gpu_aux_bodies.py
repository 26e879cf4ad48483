"""GPU parity of the two-phase / table-weighted round bodies (JA_EVAL_WSUM, WDOT2, WIDENT, DOT2_L2H, SQ_EQHI, DOT2_EQHI,
DOT2_EQLOW) against the Python restatement of the reference loops (oracle/pyref/bodies.py), bit-exact, including the phase where
the eq polynomial of the mean-of-squares / einsum schedules has collapsed to its final claim."""
import random

import numpy as np
import pytest

from oracle.pyref import bodies as B
from oracle.pyref import field as F
from oracle.pyref import poly as PL
from tests.util import from_mont_array, rand_challenge, to_mont_array

pytestmark = pytest.mark.gpu
P = F.P


def _polys(ctx, cols):
    from jolt_atlas_b200 import MultilinearPolynomial
    return [MultilinearPolynomial.from_fr(ctx, to_mont_array(c)) for c in cols]


@pytest.mark.parametrize("m,s", [(3, 0), (6, 2), (11, 4), (13, 13)])
def test_table_weighted_low_to_high(ctx, m, s):
    from jolt_atlas_b200 import EvalKernel, round_eval
    rng = random.Random(100 * m + s)
    n = 1 << m
    s = min(s, m - 1)
    x, e = [rng.randrange(P) for _ in range(n)], [rng.randrange(P) for _ in range(n)]
    tab = [rng.randrange(P) for _ in range(max(1, (n // 2) >> s))]
    px, pe, pt = _polys(ctx, [x, e, tab])
    assert from_mont_array(round_eval(ctx, EvalKernel.WSUM, [px, pt], aux_u32=s)) == B.wsum(x, tab, s)
    assert from_mont_array(round_eval(ctx, EvalKernel.WDOT2, [px, pe, pt], aux_u32=s)) == B.wdot2(x, e, tab, s)
    assert from_mont_array(round_eval(ctx, EvalKernel.DOT2_L2H, [px, pe])) == B.dot2_l2h(x, e)
    for q in (px, pe, pt):
        q.free()


@pytest.mark.parametrize("m,s", [(4, 1), (9, 3), (12, 5)])
def test_wident_split_eq(ctx, m, s):
    from jolt_atlas_b200 import EvalKernel, GruenSplitEqPolynomial, round_eval
    rng = random.Random(7 * m + s)
    n = 1 << m
    p = [rng.randrange(P) for _ in range(n)]
    tab = [rng.randrange(P) for _ in range((n // 2) >> s)]
    w = [F.challenge_to_fr(rand_challenge(rng)) for _ in range(m)]
    pp, pt = _polys(ctx, [p, tab])
    eq = GruenSplitEqPolynomial(ctx, to_mont_array(w), 0)
    ref = PL.GruenSplitEq(w, 0)
    assert from_mont_array(round_eval(ctx, EvalKernel.WIDENT, [pp, pt], eq, aux_u32=s)) == B.wident(p, tab, s, ref.fold)
    pp.free(); pt.free()


@pytest.mark.parametrize("m,log_eq,s", [(5, 2, 3), (10, 4, 6), (12, 1, 11), (8, 0, 0)])
def test_eq_scheduled_high_to_low(ctx, m, log_eq, s):
    from jolt_atlas_b200 import EvalKernel, round_eval
    rng = random.Random(31 * m + log_eq)
    n = 1 << m
    l, r = [rng.randrange(P) for _ in range(n)], [rng.randrange(P) for _ in range(n)]
    eq = [rng.randrange(P) for _ in range(1 << log_eq)]              # length 1 = the cached eq_bound_claim
    pl, pr, pe = _polys(ctx, [l, r, eq])
    assert from_mont_array(round_eval(ctx, EvalKernel.SQ_EQHI, [pl, pe], aux_u32=s)) == B.eq_hi([l], eq, s, True)
    assert from_mont_array(round_eval(ctx, EvalKernel.DOT2_EQHI, [pl, pr, pe], aux_u32=s)) == B.eq_hi([l, r], eq, s, False)
    log_b = min(log_eq, m - 1)
    assert from_mont_array(round_eval(ctx, EvalKernel.DOT2_EQLOW, [pl, pr, pe], aux_u32=log_b)) == B.dot2_eq_low(l, r, eq, log_b)
    for q in (pl, pr, pe):
        q.free()


def test_shape_errors(ctx):
    from jolt_atlas_b200 import EvalKernel, JoltAtlasError, round_eval
    pa, pt = _polys(ctx, [[1] * 16, [1] * 2])
    with pytest.raises(JoltAtlasError):            # table shorter than (len / 2) >> shift
        round_eval(ctx, EvalKernel.WSUM, [pa, pt], aux_u32=1)
    pa.free(); pt.free()
