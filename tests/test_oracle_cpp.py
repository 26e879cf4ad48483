"""CPU-only: the C++ oracle (oracle/cpp) against the committed golden vectors (generated from the independent Python
big-int twin by tests/golden/make_golden.py) and against hashlib for Blake2b."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import curve as C
from oracle.pyref import field as F
from oracle.pyref import hyperkzg as HK
from oracle.pyref import poly as PL
from tests.util import from_mont_array, to_mont_array

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P = F.P


def load(name):
    return json.load(open(os.path.join(G, name + ".json")))


def fq_arr(pts):
    out = np.zeros((len(pts), 8), dtype=np.uint64)
    for i, p in enumerate(pts):
        if p is None:
            continue
        out[i, :4] = F.fq_to_mont(p[0]); out[i, 4:] = F.fq_to_mont(p[1])
    return out


def pt_from(xy, inf=False):
    if inf:
        return None
    return (F.fq_from_mont(xy[:4]), F.fq_from_mont(xy[4:]))


def test_blake2b_matches_hashlib():
    rng = random.Random(1)
    for n in (0, 1, 31, 32, 64, 96, 127, 128, 129, 255, 256, 1000):
        data = bytes(rng.getrandbits(8) for _ in range(n))
        assert ORC.blake2b256(data) == hashlib.blake2b(data, digest_size=32).digest()


def test_field_golden():
    g = load("field")
    for e in g["mont"]:
        assert [int(x, 16) for x in e["limbs"]] == F.fr_to_mont(int(e["x"], 16))
    a = to_mont_array([int(e["a"], 16) for e in g["mul"]])
    b = to_mont_array([int(e["b"], 16) for e in g["mul"]])
    assert from_mont_array(ORC.fr_binop(2, a, b)) == [int(e["ab"], 16) for e in g["mul"]]
    assert from_mont_array(ORC.fr_binop(0, a, b)) == [int(e["a_plus_b"], 16) for e in g["mul"]]
    assert from_mont_array(ORC.fr_binop(1, a, b)) == [int(e["a_minus_b"], 16) for e in g["mul"]]
    a = to_mont_array([int(e["a"], 16) for e in g["challenge"]])
    c = np.array([[int(x, 16) for x in e["limbs"]] for e in g["challenge"]], dtype=np.uint64)
    assert from_mont_array(ORC.fr_binop(2, a, c)) == [int(e["a_times_c"], 16) for e in g["challenge"]]
    assert from_mont_array(ORC.fr_from_i64([e["v"] for e in g["from_i64"]])) == [int(e["fr"], 16) for e in g["from_i64"]]


def test_sumcheck_golden():
    kinds = {"add": (0, 0), "sub": (0, 1), "mul": (0, 2), "square": (0, 3), "prod": (0, 4), "cube": (0, 5), "dot2": (1, 0), "dot3": (1, 0)}
    for case in load("sumcheck"):
        fam, kind = kinds[case["kind"]]
        polys = np.stack([to_mont_array([v % P for v in z]) for z in case["polys_i32"]])
        w = np.array([F.challenge_limbs(int(c, 16)) for c in case["w_challenges"]], dtype=np.uint64).reshape(-1, 4)
        res = ORC.sumcheck_prove(fam, kind, polys, w, to_mont_array([int(case["claim"], 16)])[0], case["label"].encode(),
                                 pow_d=3 if case["kind"] == "cube" else 0)
        assert [[hex(v) for v in from_mont_array(cp)] for cp in res["coeffs"]] == case["round_polys"], case["kind"]
        assert [hex((int(r[3]) << 64) | int(r[2])) for r in res["challenges"]] == case["challenges"]
        assert [hex(v) for v in from_mont_array(res["final_claims"])] == case["final_poly_claims"]
        assert res["state"].hex() == case["transcript_state"]


def test_poly_layer_matches_python_twin():
    rng = random.Random(3)
    for m in (1, 4, 9):
        z = [rng.randrange(P) for _ in range(1 << m)]
        cs = [rng.getrandbits(128) & F.CHALLENGE_MASK for _ in range(m)]
        r_arr = np.array([F.challenge_limbs(c) for c in cs], dtype=np.uint64)
        rf = [F.challenge_to_fr(c) for c in cs]
        assert from_mont_array(ORC.eq_evals(r_arr)) == PL.eq_evals(rf)
        assert from_mont_array(ORC.evaluate(to_mont_array(z), r_arr)) == [PL.evaluate(z, rf)]
        for order in (0, 1):
            assert from_mont_array(ORC.bind(to_mont_array(z), r_arr[0], order)) == PL.bind(z, rf[0], order)


def test_curve_and_hyperkzg_golden():
    g = load("hyperkzg")
    tau = int(g["tau"], 16)
    srs = ORC.srs_powers(to_mont_array([tau])[0], 32)
    for i, p in enumerate(g["srs_first4"]):
        assert pt_from(srs[i]) == (int(p[0], 16), int(p[1], 16))
    ref_srs = [pt_from(s) for s in srs]
    assert all(C.is_on_curve(p) for p in ref_srs)
    for case in g["cases"]:
        ell = case["ell"]
        poly = to_mont_array([int(x, 16) for x in case["poly"]])
        pt = np.array([F.challenge_limbs(int(c, 16)) for c in case["point"]], dtype=np.uint64)
        cm, inf = ORC.msm_fr(srs[: 1 << ell].copy(), poly)
        assert pt_from(cm, inf) == tuple(int(v, 16) for v in case["commitment"])
        res = ORC.hyperkzg_open(srs[: 1 << ell].copy(), poly, pt, b"TestEval")
        assert [[hex(v) for v in pt_from(c)] for c in res["com"]] == case["com"]
        assert [[hex(v) for v in pt_from(c)] for c in res["w"]] == case["w"]
        assert [[hex(v) for v in from_mont_array(row)] for row in res["v"]] == case["v"]
        assert res["state"].hex() == case["transcript_state"]
    oh = g["one_hot"]
    idx = [k * oh["T"] + t for t, k in enumerate(oh["indices"]) if k is not None]
    s, inf = ORC.sum_indexed(srs, idx)
    assert [hex(v) for v in pt_from(s, inf)] == oh["commitment"]
    s, inf = ORC.msm_i64(srs, g["msm_i32"]["scalars"])
    assert [hex(v) for v in pt_from(s, inf)] == g["msm_i32"]["result"]


def test_msm_random_vs_python():
    rng = random.Random(5)
    n = 200
    srs_py = HK.srs_powers(8)
    # bases: random multiples of G built from the first SRS points
    bases = [srs_py[i % 8] for i in range(n)]
    sc = [rng.randrange(P) for _ in range(n)]
    sc[0] = 0; sc[1] = 1; sc[2] = P - 1
    got, inf = ORC.msm_fr(fq_arr(bases), to_mont_array(sc))
    assert pt_from(got, inf) == C.msm_pippenger(bases, sc, c=7)


def test_ident_instance_matches_python_twin():
    """S_IDENT (ps_shout / identity-RC cycle rounds, dense opening reduction) is not in the golden file: compare the two
    oracles directly on a seeded case."""
    from oracle.pyref import sumcheck as SC
    from oracle.pyref import transcript as TR
    rng = random.Random(21)
    m = 6
    z = [rng.randrange(P) for _ in range(1 << m)]
    w_c = [rng.getrandbits(128) & F.CHALLENGE_MASK for _ in range(m)]
    w = [F.challenge_to_fr(c) for c in w_c]
    claim = PL.evaluate(z, w)
    t = TR.Blake2bTranscript(b"ident")
    inst = SC.SplitEqInstance("ident", w, [z], claim)
    cps, rs, fin = SC.sumcheck_prove(inst, t)
    res = ORC.sumcheck_prove(0, 6, np.stack([to_mont_array(z)]), np.array([F.challenge_limbs(c) for c in w_c], dtype=np.uint64),
                             to_mont_array([claim])[0], b"ident")
    assert [from_mont_array(cp) for cp in res["coeffs"]] == [list(cp.coeffs_except_linear_term) for cp in cps]
    assert res["state"] == t.state
    assert from_mont_array(res["final_claims"]) == inst.final_claims()
