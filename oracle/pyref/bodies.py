"""TEST INFRASTRUCTURE ONLY — Python big-int restatement of the two-phase / table-weighted round bodies (SURVEY 8a addendum).
Each function returns the reduced sums one `compute_message` forms BEFORE interpolation, following the reference loop literally
(paths under jolt-atlas-core/src/onnx_proof/).  Parity is unpinned against the Rust prover (no toolchain, no KATs in the
reference); these are direct transcriptions of the cited loops and are cross-checked by the operator relations in tests."""
from __future__ import annotations

from .field import P


def sumcheck_evals(z, i, degree, low_to_high):
    """MultilinearPolynomial::sumcheck_evals (joltworks/src/poly/multilinear_polynomial.rs:873-905): [e(0), e(2), ..., e(degree)]."""
    half = len(z) // 2
    a, b = (z[2 * i], z[2 * i + 1]) if low_to_high else (z[i], z[i + half])
    m = (b - a) % P
    out, e = [a % P], b % P
    for _ in range(2, degree + 1):
        e = (e + m) % P
        out.append(e)
    return out


def wsum(p, tab, shift):
    """ExpSumProver::compute_phase_1_message (ops/softmax_last_axis/exp_sum.rs:146-158): k = kj >> (log_N - m)."""
    return [sum(p[2 * kj] * tab[kj >> shift] for kj in range(len(p) // 2)) % P]


def wdot2(x, e, tab, shift):
    """MaxIndicatorProver::compute_phase_1_message (ops/softmax_last_axis/max.rs:185-206): DEGREE_BOUND = 3 evaluations."""
    acc = [0, 0, 0]
    for kj in range(len(x) // 2):
        ev, xv = sumcheck_evals(e, kj, 3, True), sumcheck_evals(x, kj, 3, True)
        for k in range(3):
            acc[k] = (acc[k] + tab[kj >> shift] * xv[k] * ev[k]) % P
    return acc


def dot2_l2h(a, b):
    """SliceSumcheckProver::compute_message (ops/slice.rs:254-272); Gather with b = dictionary + gamma * identity (ops/gather/mod.rs:232-258)."""
    acc = [0, 0]
    for i in range(len(a) // 2):
        av, bv = sumcheck_evals(a, i, 2, True), sumcheck_evals(b, i, 2, True)
        for k in range(2):
            acc[k] = (acc[k] + av[k] * bv[k]) % P
    return acc


def eq_hi(ops, eq, shift, square):
    """MeanOfSquaresReductionProver::compute_message (ops/mean_of_squares.rs:363-386; square, one operand) and
    EinsumDotProver::compute_message, EqSchedule::High (ops/einsum/dot.rs:306-326; two operands).  len(eq) == 1: the cached
    eq_bound_claim."""
    half = len(ops[0]) // 2
    acc = [0, 0, 0]
    for i in range(half):
        ev = [eq[0]] * 3 if len(eq) == 1 else sumcheck_evals(eq, i >> shift, 3, False)
        vals = [sumcheck_evals(z, i, 3, False) for z in ops]
        for k in range(3):
            t = vals[0][k] * vals[0][k] if square else vals[0][k] * vals[1][k]
            acc[k] = (acc[k] + t * ev[k]) % P
    return acc


def dot2_eq_low(l, r, tab, log_b):
    """EinsumDotProver::compute_message, EqSchedule::Low while round < log_k (ops/einsum/dot.rs:328-347)."""
    half = len(l) // 2
    acc = [0, 0, 0]
    for jh in range(half):
        e = tab[jh & ((1 << log_b) - 1)]
        lv, rv = sumcheck_evals(l, jh, 3, False), sumcheck_evals(r, jh, 3, False)
        for k in range(3):
            acc[k] = (acc[k] + lv[k] * rv[k] * e) % P
    return acc


def wident(p, tab, shift, fold):
    """RecipMultProver::compute_phase_1_message (ops/softmax_last_axis/recip_mult.rs:196-216): split-eq fold of exp_q(k, 0) * inv_sum(k).
    `fold` is GruenSplitEq.fold of pyref/poly.py."""
    return fold(lambda kj: [p[2 * kj] * tab[kj >> shift] % P], 1)
