"""TEST INFRASTRUCTURE ONLY — CPU twin of jolt_atlas_b200/workload.py::run_device on the C++ oracle (OpenMP), stage for
stage and with the same chained transcript.  Used by tests (parity of every commitment / round polynomial / final
claim / transcript state), by __graft_entry__.smoke() and by bench.py's cpu_baseline and --impl reference legs."""
from __future__ import annotations

import numpy as np

from . import cpu as ORC
from .pyref import witness as WT

D_CLAMP = 16


def _index_lists(ni):
    T = ni.hot_k.shape[1]
    t = np.arange(T, dtype=np.uint64)
    return [ni.hot_k[i].astype(np.uint64) * np.uint64(T) + t for i in range(ni.d_hot)]


R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _fr_div(a, b):
    """a / b on Montgomery limbs."""
    to_i = lambda x: sum(int(v) << (64 * i) for i, v in enumerate(np.asarray(x, dtype=np.uint64).reshape(4)))
    q = to_i(a) * pow(to_i(b), -1, R_MOD) % R_MOD * pow(2, 256, R_MOD) % R_MOD
    return np.array([(q >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def _cache(t, final_claims):
    """cache_openings (opening_proof.rs:281, :338, :398): every cached claim is appended to the transcript."""
    for fc in (final_claims if isinstance(final_claims, (list, tuple)) else [final_claims]):
        ORC.transcript_append_scalar_each(t, fc)


def _ra_checks(ni, lo, hi, claim, t, out):
    k = np.ascontiguousarray(ni.hot_k[lo:hi])
    G = ORC.compute_ra_evals(k, 16, ni.eq_w)
    ra = np.stack([np.ascontiguousarray(ni.tables[j][ni.hot_k[j]]) for j in range(lo, hi)])
    r = ORC.batched_sumcheck_prove([
        {"kind": 4, "polys": ra, "eq_w": ni.eq_w, "claim": claim},
        {"kind": 18, "polys": G, "aux_fr": ni.gammas[lo:hi], "claim": claim},
        {"kind": 32, "polys": G, "idx": k, "eq_w": ni.eq_w, "aux_u32": 4, "aux_fr": np.concatenate([ni.gammas[lo:hi], ni.r_addr])},
    ], t)
    _cache(t, r["final_claims"])
    out["finals"].extend(r["final_claims"])
    return ra[0]


def run_cpu(srs_host: np.ndarray, inputs, rlc_host=None, node_limit: int | None = None, do_open: bool = True, iop: bool = True):
    """srs_host: (n, 8) affine Montgomery limbs.  node_limit bounds the number of nodes processed, iop=False skips the
    per-node stage (bench samples time the stages separately); rlc_host is unused (the joint polynomial is built here)."""
    t = ORC.TranscriptState(b"ONNXProof")
    out = {"commitments": [], "states": [], "finals": []}
    claim = inputs["claim"]
    nodes = inputs["nodes"] if node_limit is None else inputs["nodes"][:node_limit]
    for node_index, ni in enumerate(nodes if iop else []):
        spec = ni.spec
        # witness generation is part of prove (ONNXProof::commit_witness_polynomials, prover.rs:72-87 -> witness.rs:142-214): the twin
        # re-derives the chunk lists from the operands and checks them against the workload's
        from jolt_atlas_b200.workload import witness_op
        op, S = witness_op(spec)
        idx, _, ck, rk = WT.fused_witness(op, ni.A if ni.A.ndim == 2 else ni.A.reshape(1, -1), ni.B if ni.B.ndim == 2 else ni.B.reshape(1, -1),
                                          S, 1 << spec.log_t)
        assert np.array_equal(idx, ni.acc) and np.array_equal(ck, ni.hot_k[:D_CLAMP]) and (rk is None or np.array_equal(rk, ni.hot_k[D_CLAMP:]))
        lists = _index_lists(ni)
        for lo, hi in ((0, D_CLAMP), (D_CLAMP, ni.d_hot)):
            if hi > lo:
                coms = [ORC.sum_indexed(srs_host, idx) for idx in lists[lo:hi]]
                out["commitments"].append((np.stack([c[0] for c in coms]), np.array([c[1] for c in coms])))
        # clamp lookup read-raf: the 64 address rounds (8 phase passes + per-b prefix/suffix rounds), then the cycle rounds on
        # ra * (val + raf_val) from the running claim (see workload.run_device)
        ps = ORC.PsShout(ni.acc, ni.eq_w, 64, 8)
        pa = ps.prove_address(t, ni.gammas[0], None, 31)
        out["finals"].append(np.stack([pa["val"], pa["raf_val"], pa["claim"]]))
        scale = ORC.fr_binop(0, pa["val"].reshape(1, 4), pa["raf_val"].reshape(1, 4))[0]
        ra_ps = ps.materialize_ra(pa["v"].reshape(-1, 4))
        ra_ps = ORC.fr_binop(2, ra_ps, np.broadcast_to(scale, ra_ps.shape).copy())
        ps.free()
        r = ORC.sumcheck_prove_st(0, 6, np.stack([ra_ps]), ni.eq_w, pa["claim"], t)
        out["finals"].append(r["final_claims"])
        ra_claim = _fr_div(r["final_claims"][0], scale)          # the instance caches ra(r), not the scaled polynomial's claim
        _cache(t, ra_claim)
        out["finals"].append(ra_claim.reshape(1, 4))
        _ra_checks(ni, 0, D_CLAMP, claim, t, out)
        if spec.kind == "einsum":
            left = ORC.tensor_fold_i32(ni.A, ORC.eq_evals(ni.eq_rows), False)
            right = ORC.tensor_fold_i32(ni.B, ORC.eq_evals(ni.eq_cols), True)
            r = ORC.sumcheck_prove_st(1, 0, np.stack([left, right]), None, claim, t)
        else:
            a, b = ORC.fr_from_i64(ni.A), ORC.fr_from_i64(ni.B)
            r = ORC.sumcheck_prove_st(0, 2 if spec.kind == "mul" else 0, np.stack([a, b]), ni.eq_w, claim, t)
        _cache(t, r["final_claims"])
        out["finals"].append(r["final_claims"])
        if ni.d_hot > D_CLAMP:
            _ra_checks(ni, D_CLAMP, ni.d_hot, claim, t, out)
            # remainder range check: IdentityRCProver's 14 address rounds (7 phases), then the cycle rounds on ra * raf_val
            from jolt_atlas_b200.workload import identity_rc_phases
            rp = ORC.PsShout(ni.rem, ni.eq_w, S, identity_rc_phases(S))
            pr = rp.prove_identity_rc(t, None)
            out["finals"].append(np.stack([pr["raf_val"], pr["claim"]]))
            ra_rem = rp.materialize_ra(pr["v"].reshape(-1, 4))
            ra_rem = ORC.fr_binop(2, ra_rem, np.broadcast_to(pr["raf_val"], ra_rem.shape).copy())
            rp.free()
            r = ORC.sumcheck_prove_st(0, 6, np.stack([ra_rem]), ni.eq_w, pr["claim"], t)
            out["finals"].append(r["final_claims"])
            ra_claim = _fr_div(r["final_claims"][0], pr["raf_val"])
            _cache(t, ra_claim)
            out["finals"].append(ra_claim.reshape(1, 4))
        # the layer's operators without lookups (mask, softmax normalisation, layer norms)
        for ax in inputs.get("aux", {}).get(node_index, ()):
            polys = np.stack([ORC.fr_from_i64(col) for col in ax.polys])
            r = ORC.sumcheck_prove_st(0, ax.kind, polys, ax.eq_w, claim, t, gammas=ax.aux)
            _cache(t, r["final_claims"])
            out["finals"].append(r["final_claims"])
        out["states"].append(t.state)
    if do_open:
        # prove_reduced_openings: batched opening reduction over every one-hot polynomial, gamma powers, RLC, HyperKZG open
        n = 1 << inputs["ell"]
        descs = []
        for ni in nodes:
            for j in range(ni.d_hot):
                descs.append({"kind": 34, "polys": None, "idx": ni.hot_k[j:j + 1], "eq_w": ni.eq_w, "aux_fr": ni.r_addr, "aux_u32": 4,
                              "claim": claim})
        r = ORC.batched_sumcheck_prove(descs, t)
        claims = np.concatenate(r["final_claims"])
        out["finals"].append(claims)
        ORC.transcript_append_scalars(t, claims)
        gammas = ORC.transcript_challenge_scalar_powers(t, claims.shape[0])
        joint = np.zeros((n, 4), dtype=np.uint64)
        o = 0
        for ni in nodes:
            ORC.rlc_add_onehot(joint, ni.hot_k, gammas[o:o + ni.d_hot])
            o += ni.d_hot
        out["rlc"] = joint
        out["open"] = ORC.hyperkzg_open_st(srs_host[:n], joint, r["challenges"], t)
        out["states"].append(t.state)
    return out
