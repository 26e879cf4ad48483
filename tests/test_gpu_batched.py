"""BatchedSumcheck::prove on the device (fused bind+eval round kernels, host-mapped result slots) against the C++ oracle:
the RA one-hot checks batch [RaVirtual product-of-d, HammingWeight over G, Booleanity], mixed batches with instances of
different lengths (late start + 2^k claim scaling), and the ja_addr consumers (commit, gather, G scatter).  Bit-exact."""
import numpy as np
import pytest

from oracle import cpu as ORC
from tests.util import to_mont_array

pytestmark = pytest.mark.gpu
TAU = 0x1234567890abcdef1122334455667788


def _chal(rng, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def _rand_fr(rng, shape):
    a = rng.integers(0, 1 << 63, size=tuple(shape) + (4,), dtype=np.uint64)
    a[..., 3] &= np.uint64((1 << 60) - 1)
    return a


def _same(got, want):
    assert len(got["coeffs"]) == len(want["coeffs"])
    for a, b in zip(got["coeffs"], want["coeffs"]):
        assert np.array_equal(a, b)
    assert np.array_equal(got["challenges"], want["challenges"])
    for a, b in zip(got["final_claims"], want["final_claims"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("d,log_k,log_t,none_frac", [(16, 4, 10, 0.0), (4, 4, 12, 0.01), (3, 2, 5, 0.2), (16, 4, 1, 0.0), (1, 1, 3, 0.0)])
def test_ra_onehot_checks_batch(ctx, d, log_k, log_t, none_frac):
    from jolt_atlas_b200 import Blake2bTranscriptState, EvalKernel, InstanceKind, OneHotAddresses, batched_sumcheck_prove
    rng = np.random.default_rng(100 + d + log_t)
    K, T = 1 << log_k, 1 << log_t
    k = rng.integers(0, K, size=(d, T), dtype=np.uint32)
    k[rng.random((d, T)) < none_frac] = 0xFFFFFFFF
    r_cycle, r_addr, gam = _chal(rng, log_t), _chal(rng, log_k), _chal(rng, d)
    hw_g = _rand_fr(rng, (d,))
    tables = _rand_fr(rng, (d, K))
    claims = _rand_fr(rng, (2,))
    addr = OneHotAddresses(ctx, k, K)
    G = addr.ra_evals(r_cycle)
    Gw = ORC.compute_ra_evals(k, K, r_cycle)
    assert np.array_equal(G, Gw)
    ra = addr.gather(tables)
    ra_host = np.stack([np.where((k[i] == 0xFFFFFFFF)[:, None], np.uint64(0), tables[i][np.minimum(k[i], K - 1)]) for i in range(d)])
    for i in range(d):
        assert np.array_equal(ra[i].to_host(), ra_host[i])
    t_dev, t_cpu = Blake2bTranscriptState(b"ra_onehot"), ORC.TranscriptState(b"ra_onehot")
    got = batched_sumcheck_prove(ctx, [
        {"kind": EvalKernel.PROD, "polys": ra, "eq_w": r_cycle, "claim": claims[0]},
        {"kind": InstanceKind.HAMMING_TABLES, "tables": G, "aux_fr": hw_g, "claim": claims[1]},
        {"kind": InstanceKind.BOOLEANITY, "tables": G, "addr": addr, "eq_w": r_cycle, "gammas": gam, "r_address": r_addr},
    ], t_dev)
    want = ORC.batched_sumcheck_prove([
        {"kind": 4, "polys": ra_host, "eq_w": r_cycle, "claim": claims[0]},
        {"kind": 18, "polys": Gw, "aux_fr": hw_g, "claim": claims[1]},
        {"kind": 32, "polys": Gw, "idx": k, "eq_w": r_cycle, "aux_u32": log_k, "aux_fr": np.concatenate([gam, r_addr])},
    ], t_cpu)
    _same(got, want)
    assert t_dev.state == t_cpu.state and t_dev.n_rounds == t_cpu.n_rounds
    for p in ra:
        p.free()
    addr.free()


@pytest.mark.parametrize("d,log_t", [(13, 9), (9, 7), (16, 11)])
@pytest.mark.parametrize("no_wide", [False, True])
def test_ra_checks_small_slab_variants(ctx, monkeypatch, d, log_t, no_wide):
    """The RA-check launch has three forms (256-thread blocks, 128-thread blocks, product of 9..16 factors on 64 threads per
    pair with padded lanes for d < 16): each must give the oracle's transcript.  JA_NO_WIDE=1 forces the 256-thread form."""
    if no_wide:
        monkeypatch.setenv("JA_NO_WIDE", "1")
    else:
        monkeypatch.delenv("JA_NO_WIDE", raising=False)
    test_ra_onehot_checks_batch(ctx, d, 4, log_t, 0.02)


def test_mixed_batch_different_lengths(ctx):
    """MUL (2^9), DOT2 (2^6, HighToLow), ADD (2^11), IDENT (2^3), POW d=3 (2^7) in one batch: late starts, one exchange per round."""
    from jolt_atlas_b200 import Blake2bTranscriptState, EvalKernel, MultilinearPolynomial, batched_sumcheck_prove
    rng = np.random.default_rng(77)
    spec = [(EvalKernel.MUL, 2, 9), (EvalKernel.DOT2, 2, 6), (EvalKernel.ADD, 2, 11), (EvalKernel.IDENT, 1, 3), (EvalKernel.POW, 1, 7)]
    dev, cpu = [], []
    for kind, npoly, lg in spec:
        host = _rand_fr(rng, (npoly, 1 << lg))
        claim = _rand_fr(rng, (1,))[0]
        d = {"kind": kind, "polys": [MultilinearPolynomial.from_fr(ctx, host[i]) for i in range(npoly)], "claim": claim}
        c = {"kind": kind, "polys": host, "claim": claim}
        if kind != EvalKernel.DOT2:
            w = _chal(rng, lg)
            d["eq_w"] = w; c["eq_w"] = w
        if kind == EvalKernel.POW:
            d["aux_u32"] = 3; c["aux_u32"] = 3
        dev.append(d); cpu.append(c)
    t_dev, t_cpu = Blake2bTranscriptState(b"mixed"), ORC.TranscriptState(b"mixed")
    got = batched_sumcheck_prove(ctx, dev, t_dev)
    want = ORC.batched_sumcheck_prove(cpu, t_cpu)
    _same(got, want)
    assert t_dev.state == t_cpu.state
    for d in dev:
        for p in d["polys"]:
            p.free()


def test_addr_commit_matches_indexed_sums(ctx):
    from jolt_atlas_b200 import SRS, OneHotAddresses, g1_sum_indexed_batch
    rng = np.random.default_rng(5)
    d, K, T = 5, 16, 1 << 9
    srs_host = ORC.srs_powers(to_mont_array([TAU])[0], K * T)
    srs = SRS(ctx, srs_host)
    k = rng.integers(0, K, size=(d, T), dtype=np.uint32)
    k[1, :] = 0xFFFFFFFF                      # all-None list commits to the identity (hyperkzg/tests.rs:722-745)
    k[2, 7] = 0xFFFFFFFF
    addr = OneHotAddresses(ctx, k, K)
    xy, inf = addr.commit(srs)
    lists = [np.array([int(k[i, t]) * T + t for t in range(T) if k[i, t] != 0xFFFFFFFF], dtype=np.uint64) for i in range(d)]
    xy2, inf2 = g1_sum_indexed_batch(ctx, srs, lists)
    assert np.array_equal(inf, inf2) and inf[1] and np.array_equal(xy, xy2)
    for i in (0, 2):
        w, winf = ORC.sum_indexed(srs_host, lists[i])
        assert not winf and np.array_equal(xy[i], w)
    addr.free()
    srs.free()


def test_opening_reduction_batch(ctx):
    """Batched opening reduction (opening_proof.rs:500-532): two one-hot groups (16 and 4 polynomials, different T) and
    three dense polynomials of different sizes in ONE BatchedSumcheck, against the C++ oracle."""
    from jolt_atlas_b200 import (Blake2bTranscriptState, EvalKernel, InstanceKind, MultilinearPolynomial, OneHotAddresses,
                                 batched_sumcheck_prove)
    rng = np.random.default_rng(404)
    log_k, K = 4, 16
    dev, cpu, keep = [], [], []
    for d, log_t in ((16, 9), (4, 11)):
        T = 1 << log_t
        k = rng.integers(0, K, size=(d, T), dtype=np.uint32)
        k[0, 5] = 0xFFFFFFFF
        r_cycle, r_addr = _chal(rng, log_t), _chal(rng, log_k)
        claims = _rand_fr(rng, (d,))
        addr = OneHotAddresses(ctx, k, K)
        keep.append(addr)
        dev.append({"kind": InstanceKind.OPENING_ONEHOT, "addr": addr, "eq_w": r_cycle, "r_address": r_addr, "claims": claims})
        for i in range(d):
            cpu.append({"kind": 34, "polys": None, "idx": k[i:i + 1], "eq_w": r_cycle, "aux_fr": r_addr, "aux_u32": log_k, "claim": claims[i]})
    for lg in (13, 6, 1):
        host = _rand_fr(rng, (1, 1 << lg))
        w = _chal(rng, lg)
        claim = _rand_fr(rng, (1,))[0]
        dev.append({"kind": EvalKernel.OPEN, "polys": [MultilinearPolynomial.from_fr(ctx, host[0])], "eq_w": w, "claim": claim})
        cpu.append({"kind": 20, "polys": host, "eq_w": w, "claim": claim})
    t_dev, t_cpu = Blake2bTranscriptState(b"opening"), ORC.TranscriptState(b"opening")
    got = batched_sumcheck_prove(ctx, dev, t_dev)
    want = ORC.batched_sumcheck_prove(cpu, t_cpu)
    flat = []
    for f in got["final_claims"]:
        flat.extend([f[i:i + 1] for i in range(f.shape[0])])
    assert len(flat) == len(want["final_claims"])
    for a, b in zip(got["coeffs"], want["coeffs"]):
        assert np.array_equal(a, b)
    assert np.array_equal(got["challenges"], want["challenges"])
    for a, b in zip(flat, want["final_claims"]):
        assert np.array_equal(a, b)
    assert t_dev.state == t_cpu.state
    for a in keep:
        a.free()
    for d in dev:
        for p in d.get("polys", []):
            p.free()
