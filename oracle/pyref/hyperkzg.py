"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python twin) of the HyperKZG commit/open path.

Follows joltworks/src/poly/commitment/hyperkzg/:
  mod.rs:400-447 (HyperKZG::open), :231-280 (kzg_open_batch), :192-229 (kzg_batch_open_no_rem,
  compute_witness_polynomial), :451-509 (verify_inner), :282-372 (kzg_verify_batch),
  :520-554 (commit_one_hot);  kzg.rs:227-243,285-298 (commit_variable_batch / commit_as_univariate);
  commitment_scheme.rs:54-73 (commit dispatch);  ../../unipoly.rs:247-305 (eval_as_univariate).

SRS: the reference derives (beta, g1, g2) from ChaCha20 + arkworks UniformRand (external sampling
rules), so this oracle uses its own SRS  g1_powers[i] = tau^i * G  with a fixed test tau.  Because
tau is known, the pairing check  e(L, g2) == e(R, tau*g2)  is replaced by the equivalent G1 check
L == tau * R.  Parity unpinned at the byte level; only the 368-byte length pin exists
(hyperkzg/tests.rs:108-110).
"""
from __future__ import annotations

from . import curve as C
from . import field as F
from .field import P

TEST_TAU = 0x1d3f5b7a9c2e4f6081a3c5e7092b4d6f8123456789abcdef0fedcba987654321 % P


def srs_powers(n: int, tau: int = TEST_TAU):
    out, t = [], 1
    for _ in range(n):
        out.append(C.scalar_mul(C.G1, t))
        t = t * tau % P
    return out


def commit(srs, coeffs):
    """commit_as_univariate: MSM of the coefficient vector over the SRS prefix."""
    return C.msm_pippenger(srs[: len(coeffs)], coeffs)


def commit_one_hot(srs, indices, K):
    """indices[t] in [0,K) or None; coefficient k*T + t is 1."""
    T = len(indices)
    assert len(srs) >= K * T
    return C.sum_indexed(srs, [k * T + t for t, k in enumerate(indices) if k is not None])


def eval_as_univariate(coeffs, r):
    acc, pw = 0, 1
    for c in coeffs:
        acc = (acc + c * pw) % P
        pw = pw * r % P
    return acc


def witness_polynomial(f, u):
    d = len(f)
    h = [0] * d
    for i in range(d - 1, 0, -1):
        h[i - 1] = (f[i] + h[i] * u) % P
    return h


def open(srs, poly, point_u128, transcript):
    """HyperKZG::open. point: challenges (masked u128). Returns dict(com, w, v)."""
    ell = len(point_u128)
    assert len(poly) == 1 << ell
    point = [F.challenge_to_fr(c) for c in point_u128]
    polys = [list(poly)]
    for i in range(ell - 1):
        prev = polys[i]
        x = point[ell - i - 1]
        polys.append([(x * (prev[2 * j + 1] - prev[2 * j]) + prev[2 * j]) % P for j in range(len(prev) // 2)])
    com = [commit(srs, p) for p in polys[1:]]
    transcript.append_points(com)
    r = transcript.challenge_scalar()
    u = [r, (-r) % P, r * r % P]
    # kzg_open_batch
    v = [[eval_as_univariate(f, ui) for f in polys] for ui in u]
    transcript.append_scalars([x for row in v for x in row])
    q_powers = transcript.challenge_scalar_powers(len(polys))
    Bp = [0] * len(poly)
    for f, q in zip(polys, q_powers):
        for j, c in enumerate(f):
            Bp[j] = (Bp[j] + q * c) % P
    w = [commit(srs, witness_polynomial(Bp, ui)) for ui in u]
    transcript.append_points(w)
    _d0 = transcript.challenge_scalar()
    return {"com": com, "w": w, "v": v}


def verify(srs_g1, tau, commitment, point_u128, y, proof, transcript) -> bool:
    """verify_inner + kzg_verify_batch with the pairing replaced by the known-tau check."""
    ell = len(point_u128)
    point = [F.challenge_to_fr(c) for c in point_u128]
    com = list(proof["com"])
    transcript.append_points(com)
    r = transcript.challenge_scalar()
    if r == 0 or commitment is None:
        return False
    com.insert(0, commitment)
    u = [r, (-r) % P, r * r % P]
    v = proof["v"]
    if len(v) != 3 or any(len(row) != ell for row in v):
        return False
    ypos, yneg, Y = v[0], v[1], list(v[2]) + [y % P]
    for i in range(ell):
        x = point[ell - i - 1]
        lhs = 2 * r * Y[i + 1] % P
        rhs = (r * (1 - x) * (ypos[i] + yneg[i]) + x * (ypos[i] - yneg[i])) % P
        if lhs != rhs:
            return False
    # kzg_verify_batch
    k = len(com)
    transcript.append_scalars([x for row in v for x in row])
    q_powers = transcript.challenge_scalar_powers(k)
    W = proof["w"]
    transcript.append_points(W)
    d0 = transcript.challenge_scalar()
    d1 = d0 * d0 % P
    mult = (1 + d0 + d1) % P
    B_u = [sum(a * b for a, b in zip(row, q_powers)) % P for row in v]
    bases = com + [W[0], W[1], W[2], srs_g1]
    scalars = [q * mult % P for q in q_powers] + [
        u[0], u[1] * d0 % P, u[2] * d1 % P, (-(B_u[0] + d0 * B_u[1] + d1 * B_u[2])) % P]
    L = C.msm_naive(bases, scalars)
    Rp = C.msm_naive([W[0], W[1], W[2]], [1, d0, d1])
    return L == C.scalar_mul(Rp, tau)


def serialize_proof(proof) -> bytes:
    """ark-serialize (compressed) of HyperKZGProof{com: Vec<G1Affine>, w: Vec<G1Affine>, v: Vec<Vec<Fr>>}."""
    out = bytearray()
    for pts in (proof["com"], proof["w"]):
        out += len(pts).to_bytes(8, "little")
        for p in pts:
            out += C.serialize_compressed(p)
    out += len(proof["v"]).to_bytes(8, "little")
    for row in proof["v"]:
        out += len(row).to_bytes(8, "little")
        for x in row:
            out += F.fr_to_le_bytes(x)
    return bytes(out)
