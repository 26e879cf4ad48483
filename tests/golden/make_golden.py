"""Generates tests/golden/*.json from the Python big-int oracle (oracle/pyref), seeded.  Committed together with its
outputs.  The reference has no golden vectors for this path and cannot run here, so these vectors pin
the *restatement* (two independent implementations + the GPU must all reproduce them), not the Rust binary.
Run:  PYTHONPATH=. python tests/golden/make_golden.py"""
import json
import os
import random

from oracle.pyref import curve as C
from oracle.pyref import field as F
from oracle.pyref import hyperkzg as HK
from oracle.pyref import poly as PL
from oracle.pyref import sumcheck as SC
from oracle.pyref import transcript as TR

HERE = os.path.dirname(os.path.abspath(__file__))
P = F.P
hx = lambda x: hex(x)  # noqa: E731


def field_vectors():
    rng = random.Random(0xF1E1D)
    edge = [0, 1, 2, P - 1, P - 2, F.R % P, F.R2 % P, (P - 1) // 2, (1 << 253) % P, (1 << 128) - 1]
    vals = edge + [rng.randrange(P) for _ in range(22)]
    pairs = [(vals[i], vals[(i * 7 + 3) % len(vals)]) for i in range(len(vals))]
    chal = [0, 1, F.CHALLENGE_MASK, 1 << 124, (1 << 64) - 1, 1 << 64] + [rng.getrandbits(128) for _ in range(10)]
    return {
        "modulus": hx(P),
        "mont": [{"x": hx(a), "limbs": [hx(l) for l in F.fr_to_mont(a)]} for a in vals],
        "mul": [{"a": hx(a), "b": hx(b), "ab": hx(a * b % P), "a_plus_b": hx((a + b) % P), "a_minus_b": hx((a - b) % P)} for a, b in pairs],
        "challenge": [{"u128": hx(c), "limbs": [hx(l) for l in F.challenge_limbs(c)], "fr": hx(F.challenge_to_fr(c)),
                       "a": hx(vals[i % len(vals)]), "a_times_c": hx(vals[i % len(vals)] * F.challenge_to_fr(c) % P)} for i, c in enumerate(chal)],
        "from_i64": [{"v": v, "fr": hx(v % P)} for v in [0, 1, -1, 127, -128, 2**31 - 1, -2**31, 2**32, -2**32, 2**63 - 1, -2**63]],
    }


def transcript_vectors():
    t = TR.Blake2bTranscript(b"ONNXProof")
    states = [t.state.hex()]
    t.append_message(b"UniPoly_begin"); states.append(t.state.hex())
    t.append_scalar(12345678901234567890123456789 % P); states.append(t.state.hex())
    t.append_u64(0xDEADBEEF); states.append(t.state.hex())
    t.append_scalars([1, 2, P - 1]); states.append(t.state.hex())
    t.append_point(C.G1); states.append(t.state.hex())
    t.append_point(None); states.append(t.state.hex())
    t.append_points([C.G1, C.scalar_mul(C.G1, 5)]); states.append(t.state.hex())
    c = t.challenge_scalar_optimized(); states.append(t.state.hex())
    s = t.challenge_scalar(); states.append(t.state.hex())
    pw = t.challenge_scalar_powers(4); states.append(t.state.hex())
    return {"states": states, "challenge_u128_masked": hx(c), "challenge_scalar": hx(s), "powers": [hx(x) for x in pw],
            "g1x5": [hx(v) for v in C.scalar_mul(C.G1, 5)]}


def sumcheck_vectors():
    out = []
    for kind, npoly, m, seed in [("add", 2, 4, 1), ("sub", 2, 5, 2), ("mul", 2, 6, 3), ("square", 1, 5, 4), ("cube", 1, 4, 5),
                                 ("prod", 4, 4, 6), ("mul", 2, 1, 7), ("dot2", 2, 5, 8), ("dot3", 3, 4, 9)]:
        rng = random.Random(0x5C + seed)
        polys = [[rng.randrange(-128, 128) for _ in range(1 << m)] for _ in range(npoly)]
        label = ("golden_" + kind).encode()
        t = TR.Blake2bTranscript(label)
        if kind.startswith("dot"):
            pf = [[v % P for v in z] for z in polys]
            claim = 0
            for i in range(1 << m):
                term = 1
                for z in pf:
                    term = term * z[i] % P
                claim = (claim + term) % P
            inst = SC.DotInstance(pf, claim)
            w_c = []
        else:
            w_c = [rng.getrandbits(128) & F.CHALLENGE_MASK for _ in range(m)]
            w = [F.challenge_to_fr(c) for c in w_c]
            pf = [[v % P for v in z] for z in polys]
            f = {"add": lambda a: a[0] + a[1], "sub": lambda a: a[0] - a[1], "mul": lambda a: a[0] * a[1], "square": lambda a: a[0] ** 2,
                 "cube": lambda a: a[0] ** 3, "prod": lambda a: a[0] * a[1] * a[2] * a[3]}[kind]
            outp = [f([z[i] for z in pf]) % P for i in range(1 << m)]
            claim = PL.evaluate(outp, w)
            inst = SC.SplitEqInstance(kind, w, pf, claim)
        cps, rs, fin = SC.sumcheck_prove(inst, t)
        out.append({"kind": kind, "m": m, "label": label.decode(), "polys_i32": polys, "w_challenges": [hx(c) for c in w_c],
                    "claim": hx(claim), "round_polys": [[hx(c) for c in cp.coeffs_except_linear_term] for cp in cps],
                    "challenges": [hx(c) for c in rs], "final_claim": hx(fin), "final_poly_claims": [hx(x) for x in inst.final_claims()],
                    "transcript_state": t.state.hex()})
    return out


def hyperkzg_vectors():
    rng = random.Random(0xA11CE)
    out = {"tau": hx(HK.TEST_TAU), "cases": []}
    srs = HK.srs_powers(32)
    out["srs_first4"] = [[hx(p[0]), hx(p[1])] for p in srs[:4]]
    for ell in (2, 3, 5):
        n = 1 << ell
        poly = [rng.randrange(P) for _ in range(n)]
        pt = [rng.getrandbits(128) & F.CHALLENGE_MASK for _ in range(ell)]
        cm = HK.commit(srs, poly)
        y = PL.evaluate(poly, [F.challenge_to_fr(c) for c in pt])
        t = TR.Blake2bTranscript(b"TestEval")
        pr = HK.open(srs, poly, pt, t)
        assert HK.verify(srs[0], HK.TEST_TAU, cm, pt, y, pr, TR.Blake2bTranscript(b"TestEval"))
        enc = lambda p: None if p is None else [hx(p[0]), hx(p[1])]  # noqa: E731
        out["cases"].append({"ell": ell, "poly": [hx(x) for x in poly], "point": [hx(c) for c in pt], "commitment": enc(cm), "eval": hx(y),
                             "com": [enc(p) for p in pr["com"]], "w": [enc(p) for p in pr["w"]], "v": [[hx(x) for x in row] for row in pr["v"]],
                             "transcript_state": t.state.hex(), "proof_bytes": HK.serialize_proof(pr).hex()})
    # one-hot + small-scalar commits
    K, T = 4, 8
    idx = [rng.randrange(K) if rng.random() < 0.8 else None for _ in range(T)]
    ints = [rng.randrange(-2**31, 2**31) for _ in range(32)]
    out["one_hot"] = {"K": K, "T": T, "indices": idx, "commitment": [hx(v) for v in HK.commit_one_hot(srs, idx, K)]}
    out["msm_i32"] = {"scalars": ints, "result": [hx(v) for v in C.msm_i(srs, ints)]}
    return out


if __name__ == "__main__":
    for name, fn in [("field", field_vectors), ("transcript", transcript_vectors), ("sumcheck", sumcheck_vectors), ("hyperkzg", hyperkzg_vectors)]:
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=0)
        print("wrote", name)
