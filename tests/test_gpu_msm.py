"""GPU parity tests (through the C ABI) for the commitment half: batched Pippenger MSM over every scalar width the
reference dispatches on (joltworks/src/msm/mod.rs:27-181), one-hot point sums (hyperkzg/mod.rs:520-596), edge cases.
Oracle = oracle/cpp (OpenMP restatement) and the committed golden vectors (tests/golden/hyperkzg.json, made by the
independent Python twin).  Bar: identical affine Montgomery limbs + infinity flag.
Mirrors the reference's own pins: sparse one-hot commit == dense MSM commit (hyperkzg/tests.rs:544-680),
batch == individual (:682-720), all-None one-hot commits to the identity (:722-745)."""
import json
import os
import random

import numpy as np
import pytest

from oracle import cpu as ORC
from oracle.pyref import field as F
from tests.util import rand_fr, to_mont_array

pytestmark = pytest.mark.gpu

P = F.P
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAU = 0x1234567890abcdef1122334455667788


@pytest.fixture(scope="module")
def srs_host():
    return ORC.srs_powers(to_mont_array([TAU])[0], 1 << 13)


@pytest.fixture(scope="module")
def srs(ctx, srs_host):
    from jolt_atlas_b200 import SRS
    s = SRS(ctx, srs_host)
    yield s
    s.free()


def _same(got, want):
    (gxy, ginf), (wxy, winf) = got, want
    assert bool(ginf) == bool(winf)
    if not winf:
        assert [int(v) for v in gxy] == [int(v) for v in wxy]


@pytest.mark.parametrize("n", [1, 2, 3, 17, 256, 1000, 4096, 8192])
def test_msm_fr_matches_oracle(ctx, srs, srs_host, n):
    from jolt_atlas_b200 import MsmWidth, msm_host
    rng = random.Random(n)
    sc = rand_fr(rng, n)
    for i, e in enumerate([0, 1, P - 1, 2, (1 << 253) % P, P - 2][: n]):
        sc[i] = e
    s = to_mont_array(sc)
    _same(msm_host(ctx, srs, s, MsmWidth.FR), ORC.msm_fr(srs_host[:n].copy(), s))


def test_msm_fr_device_poly_and_batch(ctx, srs, srs_host):
    from jolt_atlas_b200 import MultilinearPolynomial, msm_fr, msm_fr_batch
    rng = random.Random(99)
    polys, hosts = [], []
    for logn in (12, 11, 10, 5, 2, 1, 7):
        z = to_mont_array(rand_fr(rng, 1 << logn))
        hosts.append(z)
        polys.append(MultilinearPolynomial.from_fr(ctx, z))
    out, inf = msm_fr_batch(ctx, srs, polys)
    for i, z in enumerate(hosts):
        want = ORC.msm_fr(srs_host[: z.shape[0]].copy(), z)
        _same((out[i], inf[i]), want)
        _same(msm_fr(ctx, srs, polys[i]), want)       # batch == individual (hyperkzg/tests.rs:682-720)
    for p in polys:
        p.free()


def test_msm_fr_with_offset_and_window_overrides(ctx, srs, srs_host, monkeypatch):
    from jolt_atlas_b200 import MsmWidth, msm_host
    rng = random.Random(5)
    n, off = 777, 1234
    s = to_mont_array(rand_fr(rng, n))
    want = ORC.msm_fr(srs_host[off: off + n].copy(), s)
    for c, t in [("2", "4"), ("5", "7"), ("11", "64"), ("16", "32"), ("20", "500")]:
        monkeypatch.setenv("JA_MSM_C", c)
        monkeypatch.setenv("JA_MSM_T", t)
        _same(msm_host(ctx, srs, s, MsmWidth.FR, base_offset=off), want)


def test_msm_degenerate_scalars(ctx, srs, srs_host):
    """all-zero -> identity; all-equal scalars (one bucket per window: the skewed path); duplicated bases via sums."""
    from jolt_atlas_b200 import MsmWidth, msm_host
    n = 3000
    z = to_mont_array([0] * n)
    _, inf = msm_host(ctx, srs, z, MsmWidth.FR)
    assert inf
    for v in (1, 5, P - 1, 0xdeadbeefcafebabe1234):
        s = to_mont_array([v] * n)
        _same(msm_host(ctx, srs, s, MsmWidth.FR), ORC.msm_fr(srs_host[:n].copy(), s))
    # p-1 and 1 on the same base range cancel pairwise: sum_i (p-1) G_i + sum_i G_i = identity
    half = to_mont_array([P - 1] * 10 + [1] * 10)
    bases = np.concatenate([srs_host[:10], srs_host[:10]])
    from jolt_atlas_b200 import SRS
    s2 = SRS(ctx, bases)
    _, inf = msm_host(ctx, s2, half, MsmWidth.FR)
    assert inf
    # doubling inside a bucket: the same base twice with the same scalar
    both = to_mont_array([7, 7])
    s3 = SRS(ctx, np.concatenate([srs_host[3:4], srs_host[3:4]]))
    _same(msm_host(ctx, s3, both, MsmWidth.FR), ORC.msm_fr(srs_host[3:4].copy(), to_mont_array([14])))
    s2.free(); s3.free()


@pytest.mark.parametrize("width,lo,hi", [(1, 0, 1 << 8), (2, 0, 1 << 16), (3, 0, 1 << 32), (4, 0, 1 << 64),
                                         (5, -(1 << 31), 1 << 31), (6, -(1 << 63), 1 << 63)])
def test_msm_small_scalars(ctx, srs, srs_host, width, lo, hi):
    from jolt_atlas_b200 import MsmWidth, msm_host
    rng = random.Random(width)
    n = 2500
    vals = [rng.randrange(lo, hi) for _ in range(n)]
    vals[:4] = [lo, hi - 1, 0, 1]
    # activations are concentrated near zero: skew half of them
    for i in range(4, n, 2):
        vals[i] = rng.randrange(max(lo, -3), min(hi, 4))
    arr = np.array(vals, dtype=MsmWidth.DTYPE[width])
    got = msm_host(ctx, srs, arr, width)
    want = ORC.msm_fr(srs_host[:n].copy(), to_mont_array([v % P for v in vals]))
    _same(got, want)
    if width in (5, 6):
        _same(got, ORC.msm_i64(srs_host[:n].copy(), [int(v) for v in vals]))


def test_one_hot_commit_matches_dense_msm(ctx, srs, srs_host):
    """hyperkzg/tests.rs:544-680: sparse commit == dense MSM over the materialised 0/1 polynomial."""
    from jolt_atlas_b200 import MsmWidth, g1_sum_indexed, g1_sum_indexed_batch, msm_host
    rng = random.Random(11)
    K, T = 16, 512
    lists, dense_want = [], []
    for _ in range(5):
        ks = [rng.randrange(K) if rng.random() > 0.1 else None for _ in range(T)]
        idx = [k * T + t for t, k in enumerate(ks) if k is not None]
        lists.append(idx)
        dense = np.zeros(K * T, dtype=np.uint8)
        dense[idx] = 1
        dense_want.append(msm_host(ctx, srs, dense, MsmWidth.U8))
    lists.append([])                                  # all None -> identity (hyperkzg/tests.rs:722-745)
    out, inf = g1_sum_indexed_batch(ctx, srs, lists)
    for i in range(5):
        want = ORC.sum_indexed(srs_host, lists[i])
        _same((out[i], inf[i]), want)
        _same((out[i], inf[i]), dense_want[i])
        _same(g1_sum_indexed(ctx, srs, lists[i]), want)
    assert inf[5]
    assert g1_sum_indexed(ctx, srs, [])[1]
    # repeated index (doubling) and a long single list (the wide-bucket kernel)
    _same(g1_sum_indexed(ctx, srs, [5, 5]), ORC.msm_fr(srs_host[5:6].copy(), to_mont_array([2])))
    long = [rng.randrange(1 << 13) for _ in range(20000)]
    _same(g1_sum_indexed(ctx, srs, long), ORC.sum_indexed(srs_host, long))


def test_golden_commitments(ctx):
    from jolt_atlas_b200 import SRS, MsmWidth, g1_sum_indexed, msm_host
    g = json.load(open(os.path.join(G, "hyperkzg.json")))
    srs_h = ORC.srs_powers(to_mont_array([int(g["tau"], 16)])[0], 32)
    s = SRS(ctx, srs_h)
    for case in g["cases"]:
        poly = to_mont_array([int(x, 16) for x in case["poly"]])
        xy, inf = msm_host(ctx, s, poly, MsmWidth.FR)
        assert not inf
        assert (F.fq_from_mont(xy[:4]), F.fq_from_mont(xy[4:])) == tuple(int(v, 16) for v in case["commitment"])
    oh = g["one_hot"]
    idx = [k * oh["T"] + t for t, k in enumerate(oh["indices"]) if k is not None]
    xy, inf = g1_sum_indexed(ctx, s, idx)
    assert [hex(F.fq_from_mont(xy[:4])), hex(F.fq_from_mont(xy[4:]))] == oh["commitment"]
    xy, inf = msm_host(ctx, s, np.array(g["msm_i32"]["scalars"], dtype=np.int32), MsmWidth.I32)
    assert [hex(F.fq_from_mont(xy[:4])), hex(F.fq_from_mont(xy[4:]))] == g["msm_i32"]["result"]
    s.free()


def test_msm_errors(ctx, srs):
    from jolt_atlas_b200 import JoltAtlasError, MsmWidth, g1_sum_indexed, msm_host
    with pytest.raises(JoltAtlasError) as e:
        msm_host(ctx, srs, to_mont_array([1] * 10), MsmWidth.FR, base_offset=(1 << 13) - 5)
    assert e.value.code == -3 and "KeyLengthError" in str(e.value)
    with pytest.raises(JoltAtlasError) as e:
        g1_sum_indexed(ctx, srs, [1 << 13])
    assert e.value.code == -3


def test_msm_linearity_large(ctx):
    """Size-independent property at a bench-scale size: MSM(a) + MSM(b) == MSM(a + b) and MSM(k * a) == k * MSM(a)
    checked through the identity MSM(a) + MSM(p - a) == identity, on 2^18 pairs (oracle too slow to be the checker)."""
    from jolt_atlas_b200 import SRS, MsmWidth, msm_host
    n = 1 << 18
    rng = np.random.default_rng(7)
    # bases: the 2^13 oracle SRS tiled (repeated bases are legal MSM inputs)
    base = ORC.srs_powers(to_mont_array([TAU])[0], 1 << 10)
    bases = np.tile(base, (n // (1 << 10), 1))
    s = SRS(ctx, bases)
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)          # < 2^252 < p: valid canonical-range Montgomery limbs
    # negation in Montgomery form is limb-wise p - a (a != 0)
    pl = [(P >> (64 * k)) & F.MASK64 for k in range(4)]
    neg = np.zeros_like(a)
    borrow = np.zeros(n, dtype=np.uint64)
    for k in range(4):
        ak = a[:, k]
        t = np.uint64(pl[k]) - ak - borrow
        borrow = ((np.uint64(pl[k]) < ak) | ((np.uint64(pl[k]) == ak) & (borrow == 1)) | ((np.uint64(pl[k]) - ak) < borrow)).astype(np.uint64)
        neg[:, k] = t
    both = np.concatenate([a, neg])
    s2 = SRS(ctx, np.concatenate([bases, bases]))
    _, inf = msm_host(ctx, s2, both, MsmWidth.FR)
    assert inf, "MSM(a) + MSM(-a) must be the identity"
    xy1, inf1 = msm_host(ctx, s, a, MsmWidth.FR)
    assert not inf1
    # result must be on the curve: y^2 = x^3 + 3
    x, y = F.fq_from_mont(xy1[:4]), F.fq_from_mont(xy1[4:])
    assert (y * y - x * x * x - 3) % F.Q == 0
    s.free(); s2.free()


def test_srs_generate_matches_oracle(ctx):
    """SRS::setup's fixed-base loop (kzg.rs:45-66): g1_powers[i] = beta^i * g1, against the oracle's scalar_mul."""
    from jolt_atlas_b200 import SRS
    n = 300
    want = ORC.srs_powers(to_mont_array([TAU])[0], n)
    g1 = np.concatenate([np.array(F.fq_to_mont(1), dtype=np.uint64), np.array(F.fq_to_mont(2), dtype=np.uint64)])
    s = SRS.generate(ctx, g1, to_mont_array([TAU])[0], n)
    assert len(s) == n
    assert np.array_equal(s.to_host(), want)
    s.free()


def test_fixed_base_window_table_gives_same_points(ctx):
    """ja_srs_precompute: full-width MSMs through the 2^(16 w) * G_i table (one bucket set, no doubling tail) return the
    same group elements as the classic per-window pipeline and as the oracle; batches and index ranges included."""
    import ctypes as C
    from jolt_atlas_b200 import SRS, MultilinearPolynomial, _lib, msm_fr, msm_fr_batch
    from oracle import cpu as ORC
    n = 1 << 10
    srs_host = ORC.srs_powers(to_mont_array([0x1234567890abcdef1122334455667788])[0], n)
    plain, tab = SRS(ctx, srs_host), SRS(ctx, srs_host).precompute()
    polys = [MultilinearPolynomial.random(ctx, 1 << k, 40 + k) for k in (10, 9, 3, 1, 0)]
    # edge scalars: 0, 1, p - 1 (largest canonical value: exercises the top-window carry)
    from oracle.pyref import field as F
    edge = to_mont_array([0, 1, F.P - 1, 2 ** 253, 2 ** 16 - 1, 2 ** 15, 2 ** 15 + 1, 12345] * 2)
    polys.append(MultilinearPolynomial.from_fr(ctx, edge))
    a, ainf = msm_fr_batch(ctx, plain, polys)
    b, binf = msm_fr_batch(ctx, tab, polys)
    assert np.array_equal(ainf, binf) and np.array_equal(a, b)
    for p, xy in zip(polys, b):
        want, winf = ORC.msm_fr(srs_host[: len(p)], p.to_host())
        assert not winf and np.array_equal(xy, want)
    one = np.zeros(8, dtype=np.uint64)
    f = C.c_int32()
    _lib.check(ctx._lib.ja_msm_fr_range(ctx._h, tab._h, polys[0]._h, 100, 900, one.ctypes.data_as(_lib.u64p), C.byref(f)))
    want, _ = ORC.msm_fr(srs_host[100:900], polys[0].to_host()[100:900])
    assert np.array_equal(one, want)
    for p in polys:
        p.free()
    plain.free(); tab.free()
