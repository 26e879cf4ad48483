"""Prove-shaped synthetic workloads for the configs BASELINE.json names.

No ONNX model can be traced in this environment (no Rust, no tract/onnx), so a workload is the SHAPE of what
`ONNXProof::prove` (jolt-atlas-core/src/onnx_proof/mod.rs:152-200) does for a model, on seeded synthetic tensors:
per fused node the witness commitment of its one-hot polynomials, the cycle-round sumchecks of
the lookup arguments, the RA one-hot checks (product of d factors + Hamming weight), the arithmetic sumcheck (einsum
operand folds + dot rounds, or Mul / Add rounds), the remainder range-check rounds; then one HyperKZG opening.
Every stage runs through the public API of this package (one Fiat–Shamir transcript chained through all of them) and
has a CPU twin in oracle/ used by tests and by bench.py's cpu_baseline / --impl reference legs.
What is NOT reproduced: the claim wiring between operators, the ps_shout address rounds, booleanity, evaluation
reduction and the batched opening reduction (SURVEY.md §8f "next") — the stage list says so instead of pretending.

`build_inputs` is host-only (numpy); `run_device` drives the GPU; the oracle twin lives in oracle/workload_cpu.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

K_CHUNK = 16          # common/src/consts/general.rs:2-3 (LOG_K_CHUNK = 4)
D_CLAMP = 16          # 64-bit clamp lookups: 64 / LOG_K_CHUNK one-hot chunks (clamp_lookups/mod.rs:57)
D_REM = 4             # remainder range check: ceil(14 / 4) chunks (MODEL_SCALE = 14)
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % P
MASK64 = (1 << 64) - 1
CH_MASK = (1 << 125) - 1


@dataclass
class NodeSpec:
    kind: str          # "einsum" | "mul" | "add"
    log_t: int         # log2 of the pow2-padded number of outputs
    m: int = 0         # einsum: rows of A (m x k)
    k: int = 0         # einsum: contraction length
    n: int = 0         # einsum: columns of B (k x n)


def nanogpt_nodes() -> list[NodeSpec]:
    """nanoGPT (n_embd 64, 4 heads, 4 layers, block 64, vocab 65 -> 128;
    atlas-onnx-tracer/models/nanoGPT/gen.py:208-209).  Batched per-head einsums are listed with the batch folded into m."""
    layer = [
        NodeSpec("einsum", 14, 64, 64, 192),    # qkv projection (12288 -> 2^14 outputs)
        NodeSpec("einsum", 14, 256, 16, 64),    # 4 heads x (64x16 . 16x64) scores
        NodeSpec("mul", 14),                    # softmax-side elementwise product
        NodeSpec("einsum", 12, 256, 64, 16),    # 4 heads x (64x64 . 64x16)
        NodeSpec("einsum", 12, 64, 64, 64),     # output projection
        NodeSpec("add", 12),                    # residual
        NodeSpec("einsum", 14, 64, 64, 256),    # MLP up
        NodeSpec("mul", 14),                    # activation-side product
        NodeSpec("einsum", 12, 64, 256, 64),    # MLP down
        NodeSpec("add", 12),                    # residual
    ]
    return layer * 4 + [NodeSpec("einsum", 13, 64, 64, 128)]   # lm_head


def microgpt_nodes() -> list[NodeSpec]:
    """microgpt (n_embd 16, 4 heads, 1 layer, block 16, vocab 32; jolt-atlas-core/examples/microgpt.rs:22-31)."""
    return [
        NodeSpec("einsum", 10, 16, 16, 48),
        NodeSpec("einsum", 10, 64, 4, 16),
        NodeSpec("mul", 10),
        NodeSpec("einsum", 8, 64, 16, 4),
        NodeSpec("einsum", 8, 16, 16, 16),
        NodeSpec("add", 8),
        NodeSpec("einsum", 10, 16, 16, 64),
        NodeSpec("mul", 10),
        NodeSpec("einsum", 8, 16, 64, 16),
        NodeSpec("add", 8),
        NodeSpec("einsum", 9, 16, 16, 32),
    ]


CONFIGS = {
    "microgpt": {"nodes": microgpt_nodes, "ell": 14, "seed": 0x42},
    "nanoGPT": {"nodes": nanogpt_nodes, "ell": 18, "seed": 0x1096},
}


def _challenges(rng: np.random.Generator, n: int) -> np.ndarray:
    """n random 125-bit challenges as Montgomery limbs {0, 0, lo, hi} (mont_ark_u128.rs:51-63)."""
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 2] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    out[:, 3] = rng.integers(0, 1 << 61, size=n, dtype=np.uint64)
    return out


def _mont_small(vals) -> np.ndarray:
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        x = (int(v) % P) * R % P
        for k in range(4):
            out[i, k] = (x >> (64 * k)) & MASK64
    return out


@dataclass
class NodeInputs:
    spec: NodeSpec
    d_hot: int                 # one-hot polynomials committed for this node (16 clamp [+ 4 remainder])
    hot_k: np.ndarray          # (d_hot, T) uint32 addresses k in [0, 16)
    tables: np.ndarray         # (d_hot, 16, 4) Fr eq tables the RA polynomials are materialised from
    eq_w: np.ndarray           # (log_t, 4) eq point of the node's split-eq sumchecks
    gammas: np.ndarray         # (d_hot, 4) Hamming-weight batching coefficients
    A: np.ndarray | None = None    # einsum left operand (m x k) i32 / mul, add: left operand (T,) i32
    B: np.ndarray | None = None
    eq_rows: np.ndarray | None = None   # einsum: eq point over the m rows / the n columns
    eq_cols: np.ndarray | None = None


def build_inputs(config: str, seed: int | None = None):
    """Seeded synthetic inputs of the traced shapes: i32 tensors in the i8 range (Tensor::random_small,
    atlas-onnx-tracer/src/tensor/mod.rs:178-183), u32 one-hot addresses, 125-bit challenges."""
    cfg = CONFIGS[config]
    rng = np.random.default_rng(cfg["seed"] if seed is None else seed)
    nodes = []
    for spec in cfg["nodes"]():
        T = 1 << spec.log_t
        d_hot = D_CLAMP + (0 if spec.kind == "add" else D_REM)
        ni = NodeInputs(spec=spec, d_hot=d_hot,
                        hot_k=rng.integers(0, K_CHUNK, size=(d_hot, T), dtype=np.uint32),
                        tables=np.stack([_challenges(rng, K_CHUNK) for _ in range(d_hot)]),
                        eq_w=_challenges(rng, spec.log_t), gammas=_challenges(rng, d_hot))
        if spec.kind == "einsum":
            ni.A = rng.integers(-128, 128, size=(spec.m, spec.k), dtype=np.int32)
            ni.B = rng.integers(-128, 128, size=(spec.k, spec.n), dtype=np.int32)
            ni.eq_rows = _challenges(rng, (spec.m - 1).bit_length())
            ni.eq_cols = _challenges(rng, (spec.n - 1).bit_length())
        else:
            ni.A = rng.integers(-128, 128, size=T, dtype=np.int32)
            ni.B = rng.integers(-128, 128, size=T, dtype=np.int32)
        nodes.append(ni)
    ell = cfg["ell"]
    open_point = _challenges(rng, ell)
    claim = _challenges(rng, 1)[0]       # the prover never checks its input claim; one fixed value feeds every instance
    return {"config": config, "nodes": nodes, "ell": ell, "open_point": open_point, "claim": claim,
            "rlc_seed": int(rng.integers(1, 1 << 31))}


def onehot_index_lists(ni: NodeInputs):
    """hyperkzg/mod.rs:536-542: coefficient index k*T + t for every timestep t."""
    T = ni.hot_k.shape[1]
    t = np.arange(T, dtype=np.uint64)
    return [ni.hot_k[i].astype(np.uint64) * np.uint64(T) + t for i in range(ni.d_hot)]


def h2d_bytes(inputs) -> int:
    """Bytes of per-proof inputs that cross host->device in the end-to-end path."""
    total = 0
    for ni in inputs["nodes"]:
        total += ni.d_hot * ni.hot_k.shape[1] * (8 + 4)            # commit index lists (u64) + RA addresses (u32)
        total += ni.tables.nbytes + ni.A.nbytes + ni.B.nbytes
    return total


def run_device(ctx, srs, inputs, resident=None):
    """One prove-shaped pass on the GPU.  Returns dict(commitments, states, open) for parity checks.
    `resident` (from make_resident) supplies device-resident copies of the per-proof inputs; without it every input is
    uploaded from the host arrays inside this call (the end-to-end path)."""
    from . import api as A
    t = A.Blake2bTranscriptState(b"ONNXProof")
    out = {"commitments": [], "states": [], "finals": [], "msg_bytes": 0}
    claim = inputs["claim"]

    def _sc(*a, **kw):
        r = A.sumcheck_prove(*a, **kw)
        out["msg_bytes"] += sum(c.nbytes for c in r["coeffs"])     # round polynomials read back from the device
        return r
    for i, ni in enumerate(inputs["nodes"]):
        spec = ni.spec
        res = resident["nodes"][i] if resident else None
        # A. witness commitment: HyperKZG::batch_commit_one_hot (prover.rs:236-249 -> hyperkzg/mod.rs:558-596)
        if res:
            com, inf = res["hot"].commit(srs)
        else:
            hot = A.OneHotBatch(ctx, onehot_index_lists(ni))
            com, inf = hot.commit(srs)
            hot.free()
        out["commitments"].append((com, inf))
        # RA polynomials materialised from the addresses (ra_poly.rs; shout.rs:549-598)
        if res:
            ra = [p.clone() for p in res["ra"]]
        else:
            ra = [A.MultilinearPolynomial.from_lookup(ctx, ni.tables[j], ni.hot_k[j]) for j in range(ni.d_hot)]
        # B. lookup read-raf cycle rounds (ps_shout/mod.rs:464-488): [ra0], degree 2
        p = ra[0].clone()
        r = _sc(ctx, A.EvalKernel.IDENT, [p], claim, t, eq_w=ni.eq_w)
        out["finals"].append(r["final_claims"]); p.free()
        # C. RA one-hot checks (shout.rs:399-466): Hamming weight over all chunks, then RA virtualisation = product of d
        hw = [q.clone() for q in ra[:D_CLAMP]]
        r = _sc(ctx, A.EvalKernel.SUM1, hw, claim, t, gammas=ni.gammas[:D_CLAMP])
        out["finals"].append(r["final_claims"])
        for q in hw:
            q.free()
        r = _sc(ctx, A.EvalKernel.PROD, ra[:D_CLAMP], claim, t, eq_w=ni.eq_w)
        out["finals"].append(r["final_claims"])
        # D. the operator's own sumcheck
        if spec.kind == "einsum":
            # EinsumDotProver::initialize (einsum/dot.rs:259-283): fold both operands with the eq tables, then k dot rounds
            eq_r = A.EqPolynomial.evals(ctx, ni.eq_rows)
            eq_c = A.EqPolynomial.evals(ctx, ni.eq_cols)
            left = A.tensor_fold_i32(ctx, ni.A, eq_r, transpose=False)     # (m x k) folded over rows -> k
            right = A.tensor_fold_i32(ctx, ni.B, eq_c, transpose=True)     # (k x n) folded over columns -> k
            r = _sc(ctx, A.EvalKernel.DOT2, [left, right], claim, t)
            out["finals"].append(r["final_claims"])
            for q in (eq_r, eq_c, left, right):
                q.free()
        else:
            if res:
                a, b = res["A"].clone(), res["B"].clone()
            else:
                a, b = A.MultilinearPolynomial.from_i32(ctx, ni.A), A.MultilinearPolynomial.from_i32(ctx, ni.B)
            kind = A.EvalKernel.MUL if spec.kind == "mul" else A.EvalKernel.ADD
            r = _sc(ctx, kind, [a, b], claim, t, eq_w=ni.eq_w)
            out["finals"].append(r["final_claims"])
            a.free(); b.free()
        if ni.d_hot > D_CLAMP:
            # E. remainder range check cycle rounds (identity_range_check.rs:332-358)
            p = ra[D_CLAMP].clone()
            r = _sc(ctx, A.EvalKernel.IDENT, [p], claim, t, eq_w=ni.eq_w)
            out["finals"].append(r["final_claims"]); p.free()
            # F. remainder RA checks: product of d = 4
            r = _sc(ctx, A.EvalKernel.PROD, ra[D_CLAMP:], claim, t, eq_w=ni.eq_w)
            out["finals"].append(r["final_claims"])
        for q in ra:
            q.free()
        out["states"].append(t.state)
    # G. joint opening: HyperKZG::open of a 2^ell polynomial (prover.rs:164-170); the RLC polynomial is synthetic
    rlc = resident["rlc"] if resident else A.MultilinearPolynomial.random(ctx, 1 << inputs["ell"], inputs["rlc_seed"])
    out["open"] = A.hyperkzg_open(ctx, srs, rlc, inputs["open_point"], t)
    if not resident:
        rlc.free()
    out["states"].append(t.state)
    return out


def make_resident(ctx, inputs):
    """Upload every per-proof input once (the device-resident leg of the bench clones from these)."""
    from . import api as A
    nodes = []
    for ni in inputs["nodes"]:
        d = {"hot": A.OneHotBatch(ctx, onehot_index_lists(ni)),
             "ra": [A.MultilinearPolynomial.from_lookup(ctx, ni.tables[j], ni.hot_k[j]) for j in range(ni.d_hot)]}
        if ni.spec.kind != "einsum":
            d["A"] = A.MultilinearPolynomial.from_i32(ctx, ni.A)
            d["B"] = A.MultilinearPolynomial.from_i32(ctx, ni.B)
        nodes.append(d)
    return {"nodes": nodes, "rlc": A.MultilinearPolynomial.random(ctx, 1 << inputs["ell"], inputs["rlc_seed"])}


def free_resident(res):
    for d in res["nodes"]:
        d["hot"].free()
        for p in d["ra"]:
            p.free()
        for k in ("A", "B"):
            if k in d:
                d[k].free()
    res["rlc"].free()


def count_units(inputs) -> dict:
    """Work units of one pass (for throughput figures): sumcheck rounds, one-hot point additions, MSM pairs."""
    rounds = adds = 0
    for ni in inputs["nodes"]:
        lt = ni.spec.log_t
        adds += ni.d_hot * (1 << lt)
        rounds += 3 * lt + (2 * lt if ni.d_hot > D_CLAMP else 0)
        rounds += (ni.spec.k - 1).bit_length() if ni.spec.kind == "einsum" else lt
    return {"sumcheck_rounds": rounds, "onehot_point_additions": adds, "open_msm_pairs": 4 << inputs["ell"]}


# ---- measurement helpers (bench.py) -------------------------------------------------------------------------------
def sumcheck_list(inputs):
    """(kernel class of the round evaluation, number of polynomials, initial length) of every sumcheck of one pass, in order."""
    out = []
    for ni in inputs["nodes"]:
        T = 1 << ni.spec.log_t
        out.append(("round_eval_split_eq", 1, T))
        out.append(("round_sum", D_CLAMP, T))
        out.append(("round_eval_product", D_CLAMP, T))
        if ni.spec.kind == "einsum":
            out.append(("round_eval_dot", 2, 1 << (ni.spec.k - 1).bit_length()))
        else:
            out.append(("round_eval_split_eq", 2, T))
        if ni.d_hot > D_CLAMP:
            out.append(("round_eval_split_eq", 1, T))
            out.append(("round_eval_product", D_REM, T))
    return out


def algorithmic_bytes(inputs) -> dict:
    """Algorithmic HBM bytes of one pass per kernel class: bind of a length-n polynomial reads 32n and writes
    16n; a round evaluation reads 32n per participating polynomial (the Hamming-weight sum reads the even half: 16n)."""
    b = {"bind": 0, "round_eval_split_eq": 0, "round_eval_product": 0, "round_eval_dot": 0, "round_sum": 0}
    for cls, npoly, n0 in sumcheck_list(inputs):
        rounds = n0.bit_length() - 1
        tot = sum(n0 >> j for j in range(rounds))          # sum of the current lengths over the rounds
        b["bind"] += 48 * tot * npoly
        b[cls] += (16 if cls == "round_sum" else 32) * tot * npoly
    n = 1 << inputs["ell"]
    b["bind"] += 48 * sum(n >> j for j in range(inputs["ell"] - 1))        # HyperKZG folds (hyperkzg/mod.rs:415-428)
    return b


def algorithmic_fieldmuls(inputs) -> dict:
    """Nominal Fq / Fr multiplications of the compute-bound classes: 10 Fq-mul per mixed addition (XYZZ),
    d^2 Fr-mul per pair for a product-of-d round."""
    adds = sum(ni.d_hot * (1 << ni.spec.log_t) for ni in inputs["nodes"])
    n = 1 << inputs["ell"]
    msm_pairs = 4 * n
    nwin = 16                                         # 255 bits / c = 16 signed windows
    prod = 0
    for cls, npoly, n0 in sumcheck_list(inputs):
        if cls == "round_eval_product":
            prod += npoly * npoly * (n0 - 1)          # sum over rounds of (n/2) pairs = n0 - 1
    return {"msm_accumulate": 10 * (adds + msm_pairs * nwin), "round_eval_product": prod}


def config_dict(config: str, inputs, world: int = 1) -> dict:
    u = count_units(inputs)
    return {"workload": "%s-shaped prove pass: %d nodes (one-hot commits K=16, lookup/RA/operator/range-check sumchecks, "
                        "chained Blake2b transcript) + HyperKZG open ell=%d; synthetic i8-range tensors" %
                        (config, len(inputs["nodes"]), inputs["ell"]),
            "sumcheck_rounds": u["sumcheck_rounds"], "onehot_point_additions": u["onehot_point_additions"],
            "open_msm_pairs": u["open_msm_pairs"],
            "l2": "no explicit flush: one pass streams %.0f MB of polynomial data through a 126 MB L2" %
                  (sum(algorithmic_bytes(inputs).values()) / 1e6),
            "parallelism": "1 GPU" if world == 1 else "%d replicas, one independent proof per GPU (no data-path collective)" % world}


def roofline_from_profile(prof: dict, inputs, peaks: dict, peaks_kind: str, ctx) -> dict:
    """Per-class achieved rates of one profiled pass + a large-n sweep of the streaming kernels."""
    ab = algorithmic_bytes(inputs)
    am = algorithmic_fieldmuls(inputs)
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    hbm = float(peaks["hbm_gbs"])
    mul_peak = ctx.calibrate_fr_mul(2000) / 1e9       # Gmul/s, register-resident Montgomery loop on this GPU
    classes = {}
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4)}
        if name in ab and v["ms"] > 0:
            row["algorithmic_bytes"] = ab[name]
            row["GBps"] = round(ab[name] / v["ms"] / 1e6, 2)
        if name in am and v["ms"] > 0:
            row["field_muls"] = am[name]
            row["Gmul_per_s"] = round(am[name] / v["ms"] / 1e6, 2)
        classes[name] = row
    dom_name = next(iter(classes))
    dom = classes[dom_name]
    if dom_name in am:
        dominant = {"kernel": dom_name, "bound": "int32 field-mul (no tensor cores on this path)", "achieved": dom["Gmul_per_s"],
                    "peak": round(mul_peak, 2), "unit": "Gmul/s", "frac": round(dom["Gmul_per_s"] / mul_peak, 4),
                    "traffic": None, "peak_source": "calibrated live: register-resident Montgomery product loop",
                    "per_launch_ms": round(dom["ms"] / dom["launches"], 5), "share_of_step": dom["share"]}
    else:
        gb = dom.get("GBps", 0.0)
        dominant = {"kernel": dom_name, "bound": "hbm", "achieved": gb, "peak": hbm, "unit": "GB/s", "frac": round(gb / hbm, 4),
                    "traffic": None, "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs)" % peaks_kind,
                    "per_launch_ms": round(dom["ms"] / dom["launches"], 5), "share_of_step": dom["share"]}
    sweep = []
    for which, name, bytes_per_n in ((0, "bind_low_to_high", 48), (1, "bind_high_to_low", 48), (2, "round_eval_mul", 64)):
        for log_n in (24, 26):
            ms = ctx.bench_kernel(which, log_n, 1, 10)
            gb = bytes_per_n * (1 << log_n) / ms / 1e6
            sweep.append({"kernel": name, "log_n": log_n, "ms": round(ms, 4), "GBps": round(gb, 1), "frac_hbm": round(gb / hbm, 4)})
    return {"dominant": dominant, "classes": classes, "sweep": sweep, "fr_mul_peak_Gmul_s": mul_peak}
