"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python twin) of the round polynomials.

Follows joltworks/src/poly/unipoly.rs (from_coeff :39-52, from_evals :55-64,
from_evals_degree2/3 :66-92, from_evals_and_hint :96-101, from_evals_toom :104-134,
vandermonde_interpolation :136-153, evaluate :219-245, compress :307-318,
CompressedUniPoly::{decompress,eval_from_hint} :503-533, append_to_transcript :550-558)
and joltworks/src/utils/gaussian_elimination.rs:9-70.
Field values are plain ints mod P.  Parity unpinned at the byte level (no reference KATs).
"""
from __future__ import annotations

from .field import P, fr_inv


def _gauss(m):
    """gaussian_elimination.rs:9-70, same elimination order (result is unique when non-singular)."""
    size = len(m)
    assert size == len(m[0]) - 1
    for i in range(size - 1):
        for j in range(i, size - 1):
            if m[i][i] != 0:
                f = m[j + 1][i] * fr_inv(m[i][i]) % P
                for k in range(i, size + 1):
                    m[j + 1][k] = (m[j + 1][k] - f * m[i][k]) % P
    for i in range(size - 1, 0, -1):
        if m[i][i] != 0:
            for j in range(i, 0, -1):
                f = m[j - 1][i] * fr_inv(m[i][i]) % P
                for k in range(size, -1, -1):
                    m[j - 1][k] = (m[j - 1][k] - f * m[i][k]) % P
    return [m[i][size] * fr_inv(m[i][i]) % P for i in range(size)]


class UniPoly:
    def __init__(self, coeffs):
        self.coeffs = [c % P for c in coeffs]

    @staticmethod
    def from_coeff(coeffs):
        c = [x % P for x in coeffs]
        while c and c[-1] == 0:
            c.pop()
        if not c:
            c = [0]
        return UniPoly(c)

    @staticmethod
    def from_evals(evals):
        n = len(evals)
        if n == 3:
            e0, e1, e2 = evals
            two_inv = fr_inv(2)
            c2 = (e0 - e1 - e1 + e2) * two_inv % P
            c1 = (e1 - e0 - c2) % P
            return UniPoly([e0, c1, c2])
        if n == 4:
            e0, e1, e2, e3 = evals
            two_inv, six_inv = fr_inv(2), fr_inv(6)
            c3 = (e3 - e0 + (e1 - e2) * 3) * six_inv % P
            c2 = ((e0 - e1 - e1 + e2) * two_inv - 3 * c3) % P
            c1 = (e1 - e0 - c2 - c3) % P
            return UniPoly([e0, c1, c2, c3])
        rows = []
        for i in range(n):
            row = [pow(i, j, P) for j in range(n)]
            row.append(evals[i] % P)
            rows.append(row)
        return UniPoly.from_coeff(_gauss(rows))

    @staticmethod
    def from_evals_and_hint(hint, evals):
        ev = list(evals)
        ev.insert(1, (hint - ev[0]) % P)
        return UniPoly.from_evals(ev)

    @staticmethod
    def from_evals_toom(evals):
        """evals on [0, 1, ..., n-2, inf]; no trimming (unipoly.rs:131-133)."""
        n = len(evals)
        rows = []
        for i in range(n - 1):
            row = [pow(i, j, P) for j in range(n)]
            row.append(evals[i] % P)
            rows.append(row)
        rows.append([0] * (n - 1) + [1, evals[n - 1] % P])
        return UniPoly(_gauss(rows))

    def degree(self):
        return len(self.coeffs) - 1

    def evaluate(self, r):
        acc, pw = 0, 1
        for c in self.coeffs:
            acc = (acc + c * pw) % P
            pw = pw * r % P
        return acc

    def compress(self):
        if len(self.coeffs) < 2:
            return CompressedUniPoly(list(self.coeffs))
        return CompressedUniPoly(self.coeffs[:1] + self.coeffs[2:])

    def scaled(self, s):
        """Mul<F> for &UniPoly (unipoly.rs:455-461): goes through from_coeff, i.e. TRIMS."""
        return UniPoly.from_coeff([c * s % P for c in self.coeffs])

    def add_assign(self, other):
        """AddAssign<&UniPoly> (unipoly.rs:430-447): pads to the longer length, no trimming."""
        n = max(len(self.coeffs), len(other.coeffs))
        a = self.coeffs + [0] * (n - len(self.coeffs))
        for i, c in enumerate(other.coeffs):
            a[i] = (a[i] + c) % P
        self.coeffs = a


class CompressedUniPoly:
    def __init__(self, coeffs_except_linear_term):
        self.coeffs_except_linear_term = list(coeffs_except_linear_term)

    def degree(self):
        return len(self.coeffs_except_linear_term)

    def decompress(self, hint):
        c = self.coeffs_except_linear_term
        lin = (hint - 2 * c[0] - sum(c[1:])) % P
        return UniPoly([c[0], lin] + c[1:])

    def eval_from_hint(self, hint, x):
        return self.decompress(hint).evaluate(x)

    def append_to_transcript(self, t):
        t.append_message(b"UniPoly_begin")
        for c in self.coeffs_except_linear_term:
            t.append_scalar(c)
        t.append_message(b"UniPoly_end")
