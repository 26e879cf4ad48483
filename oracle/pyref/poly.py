"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python twin) of the multilinear-polynomial layer.

Follows (all under joltworks/src/poly/):
  eq_poly.rs:23-36 (mle), :149-167 (evals_serial, big-endian: r[0] = MSB), :174-217 (cached / cached_rev)
  split_eq_poly.rs:86-145 (new_with_scaling), :331-372 (bind), :379-471 (gruen_poly_deg_3/2),
                   :473-493 (merge), :526-597 (par_fold_out_in[_unreduced])
  dense_mlpoly.rs:126-141 (bind HighToLow), :219-239 (bind LowToHigh), :265-305 (split_eq_evaluate)
  multilinear_polynomial.rs:873-905 (sumcheck_evals: values at X = 0, 2, 3, ...)
  compact_polynomial.rs:272-353 (first bind of small-scalar polys == the same affine map)
Field values are ints mod P; `r` arguments are FIELD values (convert challenges with
field.challenge_to_fr first).  Parity unpinned at the byte level (no reference KATs).
"""
from __future__ import annotations

from .field import P, fr_inv
from .unipoly import UniPoly

LOW_TO_HIGH = 0
HIGH_TO_LOW = 1


def eq_mle(x, y):
    acc = 1
    for a, b in zip(x, y):
        acc = acc * ((a * b + (1 - a) * (1 - b)) % P) % P
    return acc


def eq_evals(r, scaling=1):
    """EqPolynomial::evals (big-endian index: r[0] is the MSB of the table index)."""
    ev = [scaling % P]
    for rj in r:
        nxt = [0] * (2 * len(ev))
        for i, s in enumerate(ev):
            hi = s * rj % P
            nxt[2 * i + 1] = hi
            nxt[2 * i] = (s - hi) % P
        ev = nxt
    return ev


def eq_evals_cached(r, scaling=1):
    """result[j] = eq(r[..j], .) — eq_poly.rs:174-194."""
    out = [[scaling % P]]
    for j in range(len(r)):
        out.append(eq_evals(r[: j + 1], scaling))
    return out


def eq_evals_cached_rev(r, scaling=1):
    """result[j][x] = eq(r[n-j..], x) with index bit0 <-> r[n-1] — eq_poly.rs:198-217."""
    rev = list(reversed(r))
    out = [[scaling % P]]
    size = 1
    for j in range(len(r)):
        prev = out[j]
        nxt = [0] * (2 * size)
        for i in range(size):
            hi = prev[i] * rev[j] % P
            nxt[i + (1 << j)] = hi
            nxt[i] = (prev[i] - hi) % P
        out.append(nxt)
        size *= 2
    return out


def bind(z, r, order):
    """DensePolynomial bind: HighToLow a[i] + r(a[i+n/2]-a[i]); LowToHigh a[2i] + r(a[2i+1]-a[2i])."""
    n = len(z) // 2
    if order == HIGH_TO_LOW:
        return [(z[i] + r * (z[i + n] - z[i])) % P for i in range(n)]
    return [(z[2 * i] + r * (z[2 * i + 1] - z[2 * i])) % P for i in range(n)]


def sumcheck_evals(z, index, degree, order):
    """multilinear_polynomial.rs:873-905: [p(0), p(2), p(3), ...] (degree entries)."""
    n = len(z)
    if order == HIGH_TO_LOW:
        a, b = z[index], z[index + n // 2]
    else:
        a, b = z[2 * index], z[2 * index + 1]
    ev = [a % P]
    m = (b - a) % P
    cur = b
    for _ in range(1, degree):
        cur = (cur + m) % P
        ev.append(cur)
    return ev


def evaluate(z, r):
    """MultilinearPolynomial::evaluate (multilinear_polynomial.rs:766-862): r[0] binds the MSB."""
    m = len(r) // 2
    eq1, eq2 = eq_evals(r[:m]), eq_evals(r[m:])
    acc = 0
    for x1, e1 in enumerate(eq1):
        part = 0
        base = x1 * len(eq2)
        for x2, e2 in enumerate(eq2):
            part += e2 * z[base + x2]
        acc = (acc + e1 * (part % P)) % P
    return acc


class GruenSplitEq:
    """GruenSplitEqPolynomial (split_eq_poly.rs:67-598). w: list of field values."""

    def __init__(self, w, order, scaling=1):
        self.w = [x % P for x in w]
        self.order = order
        self.current_scalar = scaling % P
        n = len(w)
        if n == 0:
            self.current_index = 0
            self.E_in_vec, self.E_out_vec = [[1]], [[1]]
            return
        m = n // 2
        if order == LOW_TO_HIGH:
            wprime = self.w[:-1]
            w_out, w_in = wprime[:m], wprime[m:]
            self.E_out_vec, self.E_in_vec = eq_evals_cached(w_out), eq_evals_cached(w_in)
            self.current_index = n
        else:
            wprime = self.w[1:]
            w_in, w_out = wprime[:m], wprime[m:]
            self.E_in_vec, self.E_out_vec = eq_evals_cached_rev(w_in), eq_evals_cached_rev(w_out)
            self.current_index = 0

    def E_in(self):
        return self.E_in_vec[-1]

    def E_out(self):
        return self.E_out_vec[-1]

    def current_w(self):
        return self.w[self.current_index - 1] if self.order == LOW_TO_HIGH else self.w[self.current_index]

    def bind(self, r):
        w = self.current_w()
        self.current_scalar = self.current_scalar * ((1 - w - r + 2 * w * r) % P) % P
        n = len(self.w)
        if self.order == LOW_TO_HIGH:
            self.current_index -= 1
            if n // 2 < self.current_index and len(self.E_in_vec) > 1:
                self.E_in_vec.pop()
            elif 0 < self.current_index and len(self.E_out_vec) > 1:
                self.E_out_vec.pop()
        else:
            self.current_index += 1
            if self.current_index <= n // 2 and len(self.E_in_vec) > 1:
                self.E_in_vec.pop()
            elif self.current_index <= n and len(self.E_out_vec) > 1:
                self.E_out_vec.pop()

    def merge(self):
        if self.order == LOW_TO_HIGH:
            return eq_evals(self.w[: self.current_index], self.current_scalar)
        return eq_evals(self.w[self.current_index:], self.current_scalar)

    def fold(self, per_g, num_out):
        """par_fold_out_in_unreduced: sum_{x_out} E_out[x_out] * sum_{x_in} E_in[x_in] * per_g(g)."""
        e_out, e_in = self.E_out(), self.E_in()
        bits_in = (len(e_in) - 1).bit_length()
        acc = [0] * num_out
        for xo, eo in enumerate(e_out):
            inner = [0] * num_out
            for xi, ei in enumerate(e_in):
                vals = per_g((xo << bits_in) | xi)
                for k in range(num_out):
                    inner[k] += ei * vals[k]
            for k in range(num_out):
                acc[k] = (acc[k] + eo * (inner[k] % P)) % P
        return acc

    def gruen_poly_deg_3(self, q_constant, q_quadratic, s01):
        eq1 = self.current_scalar * self.current_w() % P
        eq0 = (self.current_scalar - eq1) % P
        eqm = (eq1 - eq0) % P
        eq2 = (eq1 + eqm) % P
        eq3 = (eq2 + eqm) % P
        c0 = eq0 * q_constant % P
        c1 = (s01 - c0) % P
        q1 = c1 * fr_inv(eq1) % P
        e2 = 2 * q_quadratic % P
        q2 = (q1 + q1 - q_constant + e2) % P
        q3 = (q2 + q1 - q_constant + e2 + e2) % P
        return UniPoly.from_evals([c0, c1, eq2 * q2 % P, eq3 * q3 % P])

    def gruen_poly_deg_2(self, q0, prev_claim):
        eq1 = self.current_scalar * self.current_w() % P
        eq0 = (self.current_scalar - eq1) % P
        eqm = (eq1 - eq0) % P
        eq2 = (eq1 + eqm) % P
        c0 = eq0 * q0 % P
        c1 = (prev_claim - c0) % P
        l1 = c1 * fr_inv(eq1) % P
        l2 = (l1 + l1 - q0) % P
        return UniPoly.from_evals([c0, c1, eq2 * l2 % P])
