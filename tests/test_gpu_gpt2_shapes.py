"""Parity at the GPT-2 shapes of BASELINE.json config 4 (the shapes `bench.py --config gpt2` measures), GPU vs the C++ oracle,
bit-exact, each under a time budget the oracle meets on the box's host cores:
  * the lm_head node's RA one-hot checks at T = 2^20 (d = 16: product of 16 + Hamming weight + Booleanity in one
    BatchedSumcheck) — the large-slab kernels (256-thread product blocks, TMA-eligible sizes) that nanoGPT never reaches;
  * an opening-reduction batch with >= 128 one-hot instances of mixed lengths (2^12 .. 2^16): the two-launch row split
    (k_round_open_rows long / short rows) and the OpenMP host glue that only switches on from 128 instances;
  * HyperKZG::open at ell = 21 on a 2^21-point SRS with the 20-bit fixed-base window table (built only from 2^21 points).
SRS note: the 2^21-point SRS is generated on the device (ja_srs_generate, itself checked against the oracle on a prefix here) and
copied to the host for the oracle: the oracle's own fixed-base loop would take a minute for it."""
import os

import numpy as np
import pytest

from oracle import cpu as ORC
from tests.util import to_mont_array

pytestmark = pytest.mark.gpu
TAU = 0x1234567890abcdef1122334455667788
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def _chal(rng, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def _same_proof(got, want):
    assert len(got["coeffs"]) == len(want["coeffs"])
    for i, (a, b) in enumerate(zip(got["coeffs"], want["coeffs"])):
        assert np.array_equal(a, b), f"round {i}"
    assert np.array_equal(got["challenges"], want["challenges"])


def test_lm_head_ra_checks_T20(ctx):
    from jolt_atlas_b200 import Blake2bTranscriptState, EvalKernel, InstanceKind, OneHotAddresses, batched_sumcheck_prove
    ORC.set_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(2020)
    d, log_k, log_t = 16, 4, 20
    k = rng.integers(0, 1 << log_k, size=(d, 1 << log_t), dtype=np.uint32)
    r_cycle, r_addr, gam, claims = _chal(rng, log_t), _chal(rng, log_k), _chal(rng, d), _chal(rng, 2)
    tables = np.stack([_chal(rng, 16) for _ in range(d)])
    addr = OneHotAddresses(ctx, k, 1 << log_k)
    G = addr.ra_evals(r_cycle)
    assert np.array_equal(G, ORC.compute_ra_evals(k, 1 << log_k, r_cycle))
    ra = addr.gather(tables)
    t_dev, t_cpu = Blake2bTranscriptState(b"lm_head"), ORC.TranscriptState(b"lm_head")
    got = batched_sumcheck_prove(ctx, [
        {"kind": EvalKernel.PROD, "polys": ra, "eq_w": r_cycle, "claim": claims[0]},
        {"kind": InstanceKind.HAMMING_TABLES, "tables": G, "aux_fr": gam, "claim": claims[1]},
        {"kind": InstanceKind.BOOLEANITY, "tables": G, "addr": addr, "eq_w": r_cycle, "gammas": gam, "r_address": r_addr}], t_dev)
    polys = np.stack([np.ascontiguousarray(tables[i][k[i]]) for i in range(d)])
    want = ORC.batched_sumcheck_prove([
        {"kind": 4, "polys": polys, "eq_w": r_cycle, "claim": claims[0]},
        {"kind": 18, "polys": G, "aux_fr": gam, "claim": claims[1]},
        {"kind": 32, "polys": G, "idx": k, "eq_w": r_cycle, "aux_u32": log_k, "aux_fr": np.concatenate([gam, r_addr])}], t_cpu)
    _same_proof(got, want)
    for a, b in zip(got["final_claims"], want["final_claims"]):
        assert np.array_equal(a, b)
    assert t_dev.state == t_cpu.state
    for q in ra:
        q.free()
    addr.free()


def test_opening_reduction_128_instances_mixed_lengths(ctx):
    from jolt_atlas_b200 import Blake2bTranscriptState, InstanceKind, OneHotAddresses, batched_sumcheck_prove
    ORC.set_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(128)
    log_k, K = 4, 16
    dev, cpu, keep = [], [], []
    for d, log_t in ((16, 16), (16, 12), (16, 14), (4, 14), (16, 12), (16, 16), (4, 12), (16, 13), (16, 14), (16, 12)):
        T = 1 << log_t
        k = rng.integers(0, K, size=(d, T), dtype=np.uint32)
        r_cycle, r_addr = _chal(rng, log_t), _chal(rng, log_k)
        claims = _chal(rng, d)
        addr = OneHotAddresses(ctx, k, K)
        keep.append(addr)
        dev.append({"kind": InstanceKind.OPENING_ONEHOT, "addr": addr, "eq_w": r_cycle, "r_address": r_addr, "claims": claims})
        for i in range(d):
            cpu.append({"kind": 34, "polys": None, "idx": k[i:i + 1], "eq_w": r_cycle, "aux_fr": r_addr, "aux_u32": log_k, "claim": claims[i]})
    assert len(cpu) >= 128
    t_dev, t_cpu = Blake2bTranscriptState(b"opening128"), ORC.TranscriptState(b"opening128")
    got = batched_sumcheck_prove(ctx, dev, t_dev)
    want = ORC.batched_sumcheck_prove(cpu, t_cpu)
    _same_proof(got, want)
    flat = np.concatenate(got["final_claims"])
    assert np.array_equal(flat, np.concatenate(want["final_claims"]))
    assert t_dev.state == t_cpu.state
    for a in keep:
        a.free()


def test_hyperkzg_open_ell21_wide_window_table(ctx):
    from jolt_atlas_b200 import SRS, Blake2bTranscriptState, MultilinearPolynomial, hyperkzg_open
    ORC.set_threads(os.cpu_count() or 1)
    rq = (1 << 256) % Q
    g1 = np.array([(rq >> (64 * k)) & ((1 << 64) - 1) for k in range(4)] + [((2 * rq % Q) >> (64 * k)) & ((1 << 64) - 1) for k in range(4)],
                  dtype=np.uint64)
    ell = 21
    tau = to_mont_array([TAU])[0]
    srs = SRS.generate(ctx, g1, tau, 1 << ell)
    srs_host = srs.to_host()
    assert np.array_equal(srs_host[:4096], ORC.srs_powers(tau, 4096))          # the device SRS is the oracle's on a prefix
    srs.precompute()                                                          # 16-bit table + the 20-bit table (n >= 2^21)
    rng = np.random.default_rng(21)
    z = rng.integers(0, 1 << 63, size=(1 << ell, 4), dtype=np.uint64)
    z[:, 3] &= np.uint64((1 << 60) - 1)                                         # canonical (< p)
    pt = _chal(rng, ell)
    poly = MultilinearPolynomial.from_fr(ctx, z)
    t_dev, t_cpu = Blake2bTranscriptState(b"open21"), ORC.TranscriptState(b"open21")
    got = hyperkzg_open(ctx, srs, poly, pt, t_dev)
    want = ORC.hyperkzg_open_st(srs_host, z, pt, t_cpu)
    for key in ("com", "v", "w"):
        assert np.array_equal(got[key], want[key]), key
    assert t_dev.state == t_cpu.state
    poly.free()
    srs.free()
