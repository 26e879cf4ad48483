"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python twin) of the sumcheck drivers and exemplar instances.

Follows:
  joltworks/src/subprotocols/sumcheck.rs:565-599 (Sumcheck::prove), :30-184 (BatchedSumcheck::prove),
        :653-686 (SumcheckInstanceProof::verify)
  joltworks/src/subprotocols/sumcheck_prover.rs:10-68 (SumcheckInstanceProver trait)
  jolt-atlas-core/src/onnx_proof/ops/mul.rs:160-185, add.rs:283-304, sub.rs:267, square.rs:163,
        einsum/dot.rs:290-375 (EqSchedule::None), cube.rs:159-166
  joltworks/src/subprotocols/mles_product_sum.rs:15-129,330-376 (product of d MLEs, split-eq weighted)
  joltworks/src/subprotocols/hamming_weight.rs:118-139 (sum_i gamma_i * ra_i, one eval)
Challenges are carried as masked u128 ints; field values are ints mod P.
Parity unpinned at the byte level (no reference KATs).
"""
from __future__ import annotations

from . import field as F
from .field import P
from .poly import GruenSplitEq, bind, sumcheck_evals, LOW_TO_HIGH, HIGH_TO_LOW
from .unipoly import UniPoly, CompressedUniPoly


# ----------------------------------------------------------------------------- instances
class Instance:
    """Mirror of SumcheckInstanceProver: num_rounds/input_claim/compute_message/ingest_challenge."""
    degree = 0

    def num_rounds(self): raise NotImplementedError
    def input_claim(self): raise NotImplementedError
    def compute_message(self, rnd, previous_claim): raise NotImplementedError
    def ingest_challenge(self, c_u128, rnd): raise NotImplementedError
    def final_claims(self): return []


def _log2(n):
    assert n & (n - 1) == 0 and n > 0
    return n.bit_length() - 1


class SplitEqInstance(Instance):
    """Family S (SURVEY §8a addendum): split-eq weighted, LowToHigh binding.
    kind in {add, sub, mul, square, cube(same-MLE power 3), prod (product of d MLEs)}"""

    def __init__(self, kind, w_fr, polys, claim):
        self.kind = kind
        self.eq = GruenSplitEq(w_fr, LOW_TO_HIGH)
        self.polys = [list(p) for p in polys]
        self.claim = claim % P
        self.degree = {"add": 2, "sub": 2, "ident": 2, "mul": 3, "square": 3, "cube": 4}.get(kind, len(polys) + 1)

    def num_rounds(self): return len(self.eq.w)
    def input_claim(self): return self.claim

    def compute_message(self, rnd, prev):
        ps = self.polys
        k = self.kind
        if k in ("add", "sub"):
            sgn = 1 if k == "add" else -1
            [q0] = self.eq.fold(lambda g: [ps[0][2 * g] + sgn * ps[1][2 * g]], 1)
            return self.eq.gruen_poly_deg_2(q0, prev)
        if k == "ident":
            (q0,) = self.eq.fold(lambda g: [ps[0][2 * g]], 1)
            return self.eq.gruen_poly_deg_2(q0, prev)
        if k == "mul":
            def f(g):
                l0, r0 = ps[0][2 * g], ps[1][2 * g]
                return [l0 * r0 % P, (ps[0][2 * g + 1] - l0) * (ps[1][2 * g + 1] - r0) % P]
            q0, q2 = self.eq.fold(f, 2)
            return self.eq.gruen_poly_deg_3(q0, q2, prev)
        if k == "square":
            def f(g):
                o0 = ps[0][2 * g]
                d = ps[0][2 * g + 1] - o0
                return [o0 * o0 % P, d * d % P]
            q0, q2 = self.eq.fold(f, 2)
            return self.eq.gruen_poly_deg_3(q0, q2, prev)
        # product of d linear factors on the grid {1..d-1, inf}  (mles_product_sum.rs)
        if k == "cube":
            factors = [ps[0]] * 3
        else:
            factors = ps
        d = len(factors)

        def f(g):
            out = []
            for x in list(range(1, d)) + [None]:
                acc = 1
                for z in factors:
                    p0, p1 = z[2 * g], z[2 * g + 1]
                    acc = acc * ((p1 - p0) if x is None else (p0 + x * (p1 - p0))) % P
                out.append(acc)
            return out
        sums = [s * self.eq.current_scalar % P for s in self.eq.fold(f, d)]
        return finish_mles_product_sum_from_evals(sums, prev, self.eq)

    def ingest_challenge(self, c, rnd):
        r = F.challenge_to_fr(c)
        self.eq.bind(r)
        self.polys = [bind(p, r, LOW_TO_HIGH) for p in self.polys]

    def final_claims(self):
        return [p[0] for p in self.polys]


def finish_mles_product_sum_from_evals(sum_evals, claim, eq):
    """mles_product_sum.rs:330-376."""
    r = eq.current_w()
    eq0, eq1 = (1 - r) % P, r
    if len(sum_evals) == 1:
        at0 = (claim - eq1 * sum_evals[0]) % P
    else:
        at0 = (claim - eq1 * sum_evals[0]) * F.fr_inv(eq0) % P
    tmp = UniPoly.from_evals_toom([at0] + list(sum_evals)).coeffs
    cc, xc = (1 - r) % P, (2 * r - 1) % P
    coeffs = [0] * (len(tmp) + 1)
    for i, c in enumerate(tmp):
        coeffs[i] = (coeffs[i] + c * cc) % P
        coeffs[i + 1] = (coeffs[i + 1] + c * xc) % P
    return UniPoly.from_coeff(coeffs)


class DotInstance(Instance):
    """Family D: plain products via sumcheck_evals, HighToLow (einsum/dot.rs EqSchedule::None, degree 2;
    with a third bound MLE (eq) of the same length, degree 3)."""

    def __init__(self, polys, claim):
        self.polys = [list(p) for p in polys]
        self.claim = claim % P
        self.degree = len(polys)

    def num_rounds(self): return _log2(len(self.polys[0]))
    def input_claim(self): return self.claim

    def compute_message(self, rnd, prev):
        half = len(self.polys[0]) // 2
        deg = self.degree
        acc = [0] * deg
        for i in range(half):
            evs = [sumcheck_evals(p, i, deg, HIGH_TO_LOW) for p in self.polys]
            for k in range(deg):
                t = 1
                for e in evs:
                    t = t * e[k] % P
                acc[k] = (acc[k] + t) % P
        return UniPoly.from_evals_and_hint(prev, acc)

    def ingest_challenge(self, c, rnd):
        r = F.challenge_to_fr(c)
        self.polys = [bind(p, r, HIGH_TO_LOW) for p in self.polys]

    def final_claims(self):
        return [p[0] for p in self.polys]


class HammingInstance(Instance):
    """hamming_weight.rs:118-139: sum_j sum_i gamma_i * ra_i[j]; degree 1, LowToHigh."""
    degree = 1

    def __init__(self, polys, gammas, claim):
        self.polys = [list(p) for p in polys]
        self.gammas = list(gammas)
        self.claim = claim % P

    def num_rounds(self): return _log2(len(self.polys[0]))
    def input_claim(self): return self.claim

    def compute_message(self, rnd, prev):
        half = len(self.polys[0]) // 2
        acc = 0
        for p, g in zip(self.polys, self.gammas):
            acc = (acc + g * sum(p[2 * j] for j in range(half))) % P
        return UniPoly.from_evals_and_hint(prev, [acc])

    def ingest_challenge(self, c, rnd):
        r = F.challenge_to_fr(c)
        self.polys = [bind(p, r, LOW_TO_HIGH) for p in self.polys]

    def final_claims(self):
        return [p[0] for p in self.polys]


# ----------------------------------------------------------------------------- drivers
def sumcheck_prove(inst, transcript):
    """Sumcheck::prove (sumcheck.rs:565-599). Returns (compressed polys, challenges u128, final claim)."""
    n = inst.num_rounds()
    claim = inst.input_claim()
    transcript.append_scalar(claim)
    prev = claim
    rs, cps = [], []
    for rnd in range(n):
        uni = inst.compute_message(rnd, prev)
        cp = uni.compress()
        cp.append_to_transcript(transcript)
        c = transcript.challenge_scalar_optimized()
        rs.append(c)
        prev = uni.evaluate(F.challenge_to_fr(c))
        inst.ingest_challenge(c, rnd)
        cps.append(cp)
    return cps, rs, prev


def batched_sumcheck_prove(insts, transcript):
    """BatchedSumcheck::prove (sumcheck.rs:30-184), front-loaded batching."""
    max_rounds = max(i.num_rounds() for i in insts)
    for i in insts:
        transcript.append_scalar(i.input_claim())
    coeffs = transcript.challenge_vector(len(insts))
    claims = [F.mul_pow_2(i.input_claim(), max_rounds - i.num_rounds()) for i in insts]
    rs, cps = [], []
    for rnd in range(max_rounds):
        remaining = max_rounds - rnd
        unis = []
        for inst, prev in zip(insts, claims):
            nr = inst.num_rounds()
            if remaining > nr:
                unis.append(UniPoly.from_coeff([F.mul_pow_2(inst.input_claim(), remaining - nr - 1)]))
            else:
                unis.append(inst.compute_message(rnd - (max_rounds - nr), prev))
        batched = UniPoly.from_coeff([])
        for u, cf in zip(unis, coeffs):
            batched.add_assign(u.scaled(cf))
        cp = batched.compress()
        cp.append_to_transcript(transcript)
        c = transcript.challenge_scalar_optimized()
        rs.append(c)
        rf = F.challenge_to_fr(c)
        claims = [u.evaluate(rf) for u in unis]
        for inst in insts:
            nr = inst.num_rounds()
            if remaining <= nr:
                inst.ingest_challenge(c, rnd - (max_rounds - nr))
        cps.append(cp)
    return cps, rs, coeffs, claims


def sumcheck_verify(cps, claim, num_rounds, degree_bound, transcript):
    """SumcheckInstanceProof::verify (sumcheck.rs:653-686). Returns (final claim, challenges)."""
    assert len(cps) == num_rounds
    e = claim % P
    rs = []
    for cp in cps:
        if cp.degree() > degree_bound:
            raise ValueError("InvalidInputLength")
        cp.append_to_transcript(transcript)
        c = transcript.challenge_scalar_optimized()
        rs.append(c)
        e = cp.eval_from_hint(e, F.challenge_to_fr(c))
    return e, rs


# ----------------------------------------------------------------------------- RA one-hot checks
def compute_ra_evals(idx_lists, K, r_cycle_fr):
    """compute_ra_evals (joltworks/src/subprotocols/shout.rs:549-598): G[i][k] = sum_{j: idx_i[j]==k} eq(r_cycle, j)."""
    from .poly import eq_evals
    eq = eq_evals(r_cycle_fr)
    G = [[0] * K for _ in idx_lists]
    for i, idx in enumerate(idx_lists):
        for j, k in enumerate(idx):
            if k is not None:
                G[i][k] = (G[i][k] + eq[j]) % P
    return G


class BooleanityInstance(Instance):
    """BooleanitySumcheckProver (joltworks/src/subprotocols/booleanity.rs:153-372): log_k address rounds over the G
    tables with the expanding table F, then log_t cycle rounds over H_i[j] = F[idx_i[j]]; degree 3, input claim 0."""
    degree = 3

    def __init__(self, G, idx_lists, gammas_fr, r_address_fr, r_cycle_fr):
        self.G = [list(g) for g in G]
        self.idx = [list(ix) for ix in idx_lists]
        self.gammas = list(gammas_fr)
        self.log_k, self.log_t = len(r_address_fr), len(r_cycle_fr)
        self.B = GruenSplitEq(r_address_fr, LOW_TO_HIGH)
        self.D = GruenSplitEq(r_cycle_fr, LOW_TO_HIGH)
        self.F = [1]
        self.H = []
        self.eq_r_r = 0

    def num_rounds(self): return self.log_k + self.log_t
    def input_claim(self): return 0

    def compute_message(self, rnd, prev):
        if rnd < self.log_k:
            m = rnd + 1

            def f(kp):
                c0 = c1 = 0
                for Gi, gam in zip(self.G, self.gammas):
                    s0 = s1 = 0
                    for k in range(1 << m):
                        Gk = Gi[(kp << m) + k]
                        Fk = self.F[k % (1 << (m - 1))]
                        GF = Gk * Fk % P
                        e_inf = GF * Fk % P
                        if (k >> (m - 1)) == 0:
                            s0 = (s0 + e_inf - GF) % P
                        s1 = (s1 + e_inf) % P
                    c0 = (c0 + gam * s0) % P
                    c1 = (c1 + gam * s1) % P
                return [c0, c1]
            q0, q2 = self.B.fold(f, 2)
            return self.B.gruen_poly_deg_3(q0, q2, prev)

        def f2(j):
            c0 = c1 = 0
            for h, gam in zip(self.H, self.gammas):
                h0 = h[2 * j]
                b = (h[2 * j + 1] - h0) % P
                c0 = (c0 + gam * h0 % P * (h0 - 1)) % P
                c1 = (c1 + gam * b % P * b) % P
            return [c0, c1]
        q0, q2 = self.D.fold(f2, 2)
        adjusted = prev * F.fr_inv(self.eq_r_r) % P
        return self.D.gruen_poly_deg_3(q0, q2, adjusted).scaled(self.eq_r_r)

    def ingest_challenge(self, c, rnd):
        r = F.challenge_to_fr(c)
        if rnd < self.log_k:
            self.B.bind(r)
            hi = [x * r % P for x in self.F]
            self.F = [(x - y) % P for x, y in zip(self.F, hi)] + hi
            if rnd == self.log_k - 1:
                self.eq_r_r = self.B.current_scalar
                self.H = [[0 if k is None else self.F[k] for k in ix] for ix in self.idx]
        else:
            self.D.bind(r)
            self.H = [bind(h, r, LOW_TO_HIGH) for h in self.H]

    def final_claims(self):
        return [h[0] for h in self.H]


# ----------------------------------------------------------------------------- batched opening reduction
def _open_q0(D, z):
    """opening_reduction.rs:355-403 / :630-673: first half only, j = (x_in << out_bits) | x_out."""
    eo, ei = D.E_out(), D.E_in()
    out_bits = len(eo).bit_length() - 1
    tot = 0
    for xi, e_in in enumerate(ei):
        inner = 0
        for xo, e_out in enumerate(eo):
            inner = (inner + e_out * z[(xi << out_bits) | xo]) % P
        tot = (tot + e_in * inner) % P
    return tot


class DenseOpeningInstance(Instance):
    """DensePolynomialProverOpening (opening_reduction.rs:337-424): sum_j eq(r, j) P[j], HighToLow, degree 2."""
    degree = 2

    def __init__(self, r_fr, poly, claim):
        self.D = GruenSplitEq(r_fr, HIGH_TO_LOW)
        self.poly = list(poly)
        self.claim = claim % P

    def num_rounds(self): return len(self.D.w)
    def input_claim(self): return self.claim
    def compute_message(self, rnd, prev): return self.D.gruen_poly_deg_2(_open_q0(self.D, self.poly), prev)

    def ingest_challenge(self, c, rnd):
        r = F.challenge_to_fr(c)
        self.D.bind(r)
        self.poly = bind(self.poly, r, HIGH_TO_LOW)

    def final_claims(self): return [self.poly[0]]


class OneHotOpeningInstance(Instance):
    """OneHotPolynomialProverOpening (opening_reduction.rs:503-723)."""
    degree = 2

    def __init__(self, idx, r_address_fr, r_cycle_fr, claim):
        from .poly import eq_evals
        self.idx = list(idx)
        self.log_k, self.log_t = len(r_address_fr), len(r_cycle_fr)
        self.D = GruenSplitEq(r_cycle_fr, HIGH_TO_LOW)
        self.B = eq_evals(r_address_fr)
        self.F = [1]
        dm = self.D.merge()
        self.G = [0] * (1 << self.log_k)
        for j, k in enumerate(self.idx):
            if k is not None:
                self.G[k] = (self.G[k] + dm[j]) % P
        self.H = []
        self.claim = claim % P

    def num_rounds(self): return self.log_k + self.log_t
    def input_claim(self): return self.claim

    def compute_message(self, rnd, prev):
        if rnd < self.log_k:
            nu = self.log_k - rnd
            half = len(self.B) // 2
            e0 = e2 = 0
            for kp in range(half):
                b0, b1 = self.B[kp], self.B[kp + half]
                b2 = (2 * b1 - b0) % P
                i0 = i2 = 0
                for k in range(kp, len(self.G), half):
                    GF = self.G[k] * self.F[k >> nu] % P
                    if ((k >> (nu - 1)) & 1) == 0:
                        i0 = (i0 + GF) % P
                        i2 = (i2 - GF) % P
                    else:
                        i2 = (i2 + 2 * GF) % P
                e0 = (e0 + b0 * i0) % P
                e2 = (e2 + b2 * i2) % P
            return UniPoly.from_evals_and_hint(prev, [e0, e2])
        ea = self.B[0]
        return self.D.gruen_poly_deg_2(_open_q0(self.D, self.H), prev * F.fr_inv(ea) % P).scaled(ea)

    def ingest_challenge(self, c, rnd):
        r = F.challenge_to_fr(c)
        if rnd < self.log_k:
            self.B = bind(self.B, r, HIGH_TO_LOW)
            nf = []
            for v in self.F:
                e1 = r * v % P
                nf += [(v - e1) % P, e1]
            self.F = nf
            if rnd == self.log_k - 1:
                self.H = [0 if k is None else self.F[k] for k in self.idx]
        else:
            self.D.bind(r)
            self.H = bind(self.H, r, HIGH_TO_LOW)

    def final_claims(self): return [self.H[0]]
