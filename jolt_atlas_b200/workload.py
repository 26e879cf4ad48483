"""Prove-shaped synthetic workloads for the configs BASELINE.json names.

No ONNX model can be traced in this environment (no Rust, no tract/onnx), so a workload is the SHAPE of what
`ONNXProof::prove` (jolt-atlas-core/src/onnx_proof/mod.rs:152-200) does for a model, on seeded synthetic tensors:
per fused node the witness commitment of its one-hot polynomials, the cycle-round sumchecks of
the lookup arguments, the RA one-hot checks (product of d factors + Hamming weight), the arithmetic sumcheck (einsum
operand folds + dot rounds, or Mul / Add rounds), the remainder range-check rounds; then one HyperKZG opening.
Every stage runs through the public API of this package (one Fiat–Shamir transcript chained through all of them) and
has a CPU twin in oracle/ used by tests and by bench.py's cpu_baseline / --impl reference legs.
Reproduced per node: the ps_shout T-sized phase passes + cycle rounds, the batched RA one-hot checks (RaVirtual, Hamming weight,
Booleanity), the operator sumcheck, the remainder checks; then the batched opening reduction, the RLC and the HyperKZG opening.
NOT reproduced: the claim wiring between operators (claims are one synthetic value; the prover never checks them), the O(256)
host math of the 64 ps_shout / identity-RC ADDRESS rounds (prefix MLEs + checkpoints: joltworks/src/lookup_tables/, the Rust
prover's unchanged code - each phase appears as its transcript traffic only), evaluation reduction, and every operator body
other than einsum-dot / Mul / Add.  The stage list says so instead of pretending.

`build_inputs` is host-only (numpy); `run_device` drives the GPU; the oracle twin lives in oracle/workload_cpu.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

K_CHUNK = 16          # common/src/consts/general.rs:2-3 (LOG_K_CHUNK = 4)
LOG_K = 4
D_CLAMP = 16          # 64-bit clamp lookups: 64 / LOG_K_CHUNK one-hot chunks (clamp_lookups/mod.rs:57)
D_REM = 4             # remainder range check: ceil(14 / 4) chunks (MODEL_SCALE = 14)
CLAMP_LOG_K = 64      # clamp_lookups/mod.rs:57
PS_PHASES = 8         # ps_shout/mod.rs:56 NUM_PHASES
SAT_BOUND = 31        # SaturationTable: clamp to [-2^31, 2^31 - 1] (lookup_tables/clamp.rs, SIGN_BIT_I32)
# suffixes of the pass: SaturationTable read-checking suffixes (clamp.rs:78-86) then the unary raf's [One, Identity] (signed_identity_poly.rs:160-167)
PS_SUFFIXES = (1, 2, 3, 0, 0, 4)
_PS_KINDS = np.array(PS_SUFFIXES, dtype=np.uint32)
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


def gpt2_nodes() -> list[NodeSpec]:
    """GPT-2 125M at seq_len 16 (n_embd 768, 12 heads, 12 layers, vocab 50257 -> 65536; jolt-atlas-core/examples/gpt2.rs:86-112,
    README.md:139).  Output counts are padded to powers of two as the tracer does (atlas-onnx-tracer/src/model/mod.rs:273-300):
    16 x 2304 -> 2^16, 16 x 768 -> 2^14, 16 x 3072 -> 2^16, logits 16 x 65536 = 2^20 (=> committed polynomials up to 2^24).
    Contraction lengths 768 / 3072 are zero-padded to 1024 / 4096."""
    layer = [
        NodeSpec("einsum", 16, 16, 768, 2304),    # qkv projection
        NodeSpec("einsum", 12, 192, 64, 16),      # 12 heads x (16x64 . 64x16) scores
        NodeSpec("mul", 12),                      # softmax-side elementwise product
        NodeSpec("einsum", 14, 192, 16, 64),      # 12 heads x (16x16 . 16x64)
        NodeSpec("einsum", 14, 16, 768, 768),     # output projection
        NodeSpec("add", 14),                      # residual
        NodeSpec("einsum", 16, 16, 768, 3072),    # MLP up
        NodeSpec("mul", 16),                      # activation-side product
        NodeSpec("einsum", 14, 16, 3072, 768),    # MLP down
        NodeSpec("add", 14),                      # residual
    ]
    return layer * 12 + [NodeSpec("einsum", 20, 16, 768, 65536)]   # lm_head


CONFIGS = {
    "microgpt": {"nodes": microgpt_nodes, "ell": 14, "seed": 0x42},
    "nanoGPT": {"nodes": nanogpt_nodes, "ell": 18, "seed": 0x1096},
    "gpt2": {"nodes": gpt2_nodes, "ell": 24, "seed": 42},
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
    gammas: np.ndarray         # (d_hot, 4) batching coefficients (booleanity gammas; Hamming-weight gamma powers stand-ins)
    r_addr: np.ndarray = None  # (log K, 4) booleanity address point
    rem: np.ndarray = None     # (T,) uint64: the rescale remainders = lookup indices of the remainder range check
    acc: np.ndarray = None     # (T,) uint64: the pre-clamp i64 accumulations (two's complement) = lookup indices of the clamp read-raf
    A: np.ndarray | None = None    # einsum left operand (m x k) i32 / mul, add: left operand (T,) i32
    B: np.ndarray | None = None
    eq_rows: np.ndarray | None = None   # einsum: eq point over the m rows / the n columns
    eq_cols: np.ndarray | None = None


# Transformer operators WITHOUT lookups that a real graph carries per layer next to the fused nodes: the causal mask (Iff, ops/iff.rs:189),
# the softmax normalisation (Div, ops/div.rs:329), the two layer norms (Rsqrt with its gamma pair, ops/rsqrt.rs:390) and their means
# (ScalarConstDiv / neural-teleport division = the three-operand linear body, ops/scalar_const_div.rs:227, neural_teleport/division.rs:231).
# (kernel id, polynomials, aux scalars, "scores" or "embed" sized); sizes follow the layer's score / projection nodes.
AUX_PER_LAYER = ((8, 3, 0, "scores"), (9, 4, 0, "scores"), (10, 5, 2, "embed"), (11, 3, 1, "embed"), (10, 5, 2, "embed"), (11, 3, 1, "embed"))
NODES_PER_LAYER = 10


@dataclass
class AuxInputs:
    kind: int
    log_t: int
    polys: np.ndarray          # (npoly, T) int32 operands (small integers; the Iff condition is a 0/1 mask)
    aux: np.ndarray | None     # (naux, 4) Fr scalars of the body (Rsqrt gammas, the constant of the linear body)
    eq_w: np.ndarray           # (log_t, 4)


def build_aux(rng: np.random.Generator, layer_specs) -> list:
    sizes = {"scores": layer_specs[1].log_t, "embed": layer_specs[4].log_t}
    out = []
    for kind, npoly, naux, size in AUX_PER_LAYER:
        log_t = sizes[size]
        polys = rng.integers(-(1 << 12), 1 << 12, size=(npoly, 1 << log_t), dtype=np.int32)
        if kind == 8:
            polys[0] = rng.integers(0, 2, size=1 << log_t, dtype=np.int32)
        out.append(AuxInputs(kind, log_t, np.ascontiguousarray(polys), _challenges(rng, naux) if naux else None, _challenges(rng, log_t)))
    return out


MODEL_SCALE = 14      # common/src/consts/general.rs (rescale bits S of Einsum / Mul; Add carries no rescale)


def witness_op(spec: NodeSpec):
    """(FusedWitness op code, rescale bits) of a node."""
    return {"einsum": (0, MODEL_SCALE), "mul": (1, MODEL_SCALE), "add": (2, 0)}[spec.kind]


def host_witness(spec: NodeSpec, A: np.ndarray, B: np.ndarray, T: int):
    """numpy twin of ja_witness_fused for workload synthesis: (lookup indices (T,) u64, remainders (T,) u64, chunk lists (16 [+ 4], T) u32)."""
    op, S = witness_op(spec)
    a, b = A.astype(np.int64), B.astype(np.int64)
    if op == 0:      # i8-range operands, contraction <= 2^10: every partial sum < 2^53, the f64 (BLAS) product is exact
        acc = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.int64).reshape(-1)
    else:
        acc = (a * b).reshape(-1) if op == 1 else (a + b).reshape(-1)
    pad = np.zeros(T, dtype=np.int64)
    pad[: acc.shape[0]] = acc
    q, r = pad >> S, pad & ((1 << S) - 1)
    idx = q.view(np.uint64)
    rows = [((idx >> np.uint64(4 * (15 - d))) & np.uint64(15)).astype(np.uint32) for d in range(16)]
    d_rem = (S + 3) // 4
    rows += [((r >> (4 * (d_rem - 1 - d))) & 15).astype(np.uint32) for d in range(d_rem)]
    return np.ascontiguousarray(idx), np.ascontiguousarray(r.astype(np.uint64)), np.ascontiguousarray(np.stack(rows))


def device_rc_phases(log_k: int) -> int:
    """Phase count of the device prover for a LOG_K-bit identity range check.  The phase structure is the prover's own choice: a
    round polynomial is a sum over the remaining address bits and the cycles whatever chunking computes it, and
    ra = prod_phase v[phase][chunk] = eq(r_address, k) for every chunking - so the device uses the FEWEST phases its kernel takes
    (chunks of at most 8 bits: 2 T-sized passes instead of the reference's 7 for LOG_K = 14) and emits the reference's transcript bit
    for bit (tests/test_gpu_psshout.py::test_identity_rc_phase_count_is_free)."""
    n = 1
    while log_k % n or log_k // n > 8:
        n += 1
    return n


def identity_rc_phases(log_k: int) -> int:
    """IdentityRCProvider::phases (joltworks/src/subprotocols/identity_range_check.rs:416-431)."""
    if log_k <= 2:
        return 1
    if log_k % 4 == 0:
        return log_k // 4
    assert log_k % 2 == 0, "odd LOG_K is not supported by the prefix-suffix decomposition"
    return log_k // 2


def build_inputs(config: str, seed: int | None = None):
    """Seeded synthetic inputs of the traced shapes: i32 tensors in the i8 range (Tensor::random_small,
    atlas-onnx-tracer/src/tensor/mod.rs:178-183), u32 one-hot addresses, 125-bit challenges."""
    cfg = CONFIGS[config]
    rng = np.random.default_rng(cfg["seed"] if seed is None else seed)
    nodes = []
    for spec in cfg["nodes"]():
        T = 1 << spec.log_t
        d_hot = D_CLAMP + (0 if spec.kind == "add" else D_REM)
        ni = NodeInputs(spec=spec, d_hot=d_hot, hot_k=None,
                        tables=np.stack([_challenges(rng, K_CHUNK) for _ in range(d_hot)]),
                        eq_w=_challenges(rng, spec.log_t), gammas=_challenges(rng, d_hot), r_addr=_challenges(rng, LOG_K))
        if spec.kind == "einsum":
            kp = 1 << (spec.k - 1).bit_length()            # contraction axis zero-padded to a power of two (MLE length)
            ni.A = np.zeros((spec.m, kp), dtype=np.int32)
            ni.B = np.zeros((kp, spec.n), dtype=np.int32)
            ni.A[:, :spec.k] = rng.integers(-128, 128, size=(spec.m, spec.k), dtype=np.int32)
            ni.B[:spec.k] = rng.integers(-128, 128, size=(spec.k, spec.n), dtype=np.int32)
            ni.eq_rows = _challenges(rng, (spec.m - 1).bit_length())
            ni.eq_cols = _challenges(rng, (spec.n - 1).bit_length())
        else:
            ni.A = rng.integers(-128, 128, size=T, dtype=np.int32)
            ni.B = rng.integers(-128, 128, size=T, dtype=np.int32)
        # the node's committed one-hot polynomials are the WITNESS of its operands (witness.rs:142-214): 4-bit chunks of the floor-rebased
        # i64 accumulation (the clamp lookup index) and of the rescale remainder.  Host copy here (numpy, exact integers) for the
        # resident leg's uploads and the CPU twin; the end-to-end leg derives them on the device from the operands (FusedWitness).
        ni.acc, ni.rem, ni.hot_k = host_witness(spec, ni.A, ni.B, T)
        nodes.append(ni)
    ell = cfg["ell"]
    open_point = _challenges(rng, ell)
    claim = _challenges(rng, 1)[0]       # the prover never checks its input claim; one fixed value feeds every instance
    rlc_seed = int(rng.integers(1, 1 << 31))
    # per-layer operators without lookups (after the fused nodes of the layer); drawn last so that the fused nodes keep their inputs
    specs = [ni.spec for ni in nodes]
    aux = {}
    for layer in range(len(nodes) // NODES_PER_LAYER):
        aux[(layer + 1) * NODES_PER_LAYER - 1] = build_aux(rng, specs[layer * NODES_PER_LAYER:(layer + 1) * NODES_PER_LAYER])
    return {"config": config, "nodes": nodes, "aux": aux, "ell": ell, "open_point": open_point, "claim": claim, "rlc_seed": rlc_seed}


def onehot_index_lists(ni: NodeInputs):
    """hyperkzg/mod.rs:536-542: coefficient index k*T + t for every timestep t."""
    T = ni.hot_k.shape[1]
    t = np.arange(T, dtype=np.uint64)
    return [ni.hot_k[i].astype(np.uint64) * np.uint64(T) + t for i in range(ni.d_hot)]


def h2d_bytes(inputs) -> int:
    """Bytes of per-proof inputs that cross host->device in the end-to-end path."""
    total = 0
    for ni in inputs["nodes"]:
        # the one-hot addresses and the clamp lookup indices are generated on the device from the operands (FusedWitness): only
        # the operands (once for the witness, once more for the operator's own polynomials / folds) and the small tables cross PCIe
        total += 2 * ni.tables.nbytes + 2 * (ni.A.nbytes + ni.B.nbytes)
    for lst in inputs.get("aux", {}).values():
        total += sum(ax.polys.nbytes for ax in lst)
    return total


def _ra_checks(A, ctx, addr, ni, lo, hi, claim, t, out, sc, keep_first=True):
    """RaOneHotChecks / RescaleRemainderRaChecks (shout.rs:399-466): BatchedSumcheck[RaVirtual (product of d),
    HammingWeight over the G tables, Booleanity] sharing one G = compute_ra_evals and one address batch."""
    G = addr.ra_evals(ni.eq_w)                                                  # shout.rs:549-598
    ra = addr.gather(ni.tables[lo:hi])                                          # ra_virtual.rs:113-134
    first = ra[0].clone() if keep_first else None
    r = A.batched_sumcheck_prove(ctx, [
        {"kind": A.EvalKernel.PROD, "polys": ra, "eq_w": ni.eq_w, "claim": claim},
        {"kind": A.InstanceKind.HAMMING_TABLES, "tables": G, "aux_fr": ni.gammas[lo:hi], "claim": claim},
        {"kind": A.InstanceKind.BOOLEANITY, "tables": G, "addr": addr, "eq_w": ni.eq_w, "gammas": ni.gammas[lo:hi],
         "r_address": ni.r_addr},
    ], t)
    out["finals"].extend(r["final_claims"])
    out["msg_bytes"] += r["msg_bytes"]
    A.MultilinearPolynomial.free_many(ra)
    return first


def run_device(ctx, srs, inputs, resident=None, comm=None, ps_shout=True):
    """One prove-shaped pass on the GPU.  Returns dict(commitments, states, finals, open) for parity checks.
    `resident` (from make_resident) supplies device-resident copies of the per-proof inputs; without it every input is
    uploaded from the host arrays inside this call (the end-to-end path).
    `comm` (parallel.Comm, world > 1): ONE proof on several GPUs — every rank holds the same inputs and runs the
    (latency-bound, strictly sequential) sumchecks replicated, while the group arithmetic is sharded: the witness
    commitments by polynomial, every MSM of the HyperKZG opening by index range (parallel.py).  All ranks end with the
    same transcript and the same proof."""
    from . import api as A
    from . import parallel as PAR
    t = A.Blake2bTranscriptState(b"ONNXProof")
    out = {"commitments": [], "states": [], "finals": [], "msg_bytes": 0}
    claim = inputs["claim"]
    # cache_openings (opening_proof.rs:281, :338, :398): the drivers append every instance's final claims to the transcript
    A.check(ctx._lib.ja_set_cache_openings(ctx._h, 1))
    try:
        return _run_device(A, PAR, ctx, srs, inputs, resident, comm, ps_shout, t, out, claim)
    finally:
        ctx._lib.ja_set_cache_openings(ctx._h, 0)


def _run_device(A, PAR, ctx, srs, inputs, resident, comm, ps_shout, t, out, claim):

    def _sc(*a, **kw):
        r = A.sumcheck_prove(*a, **kw)
        out["msg_bytes"] += r["msg_bytes"]                         # round polynomials read back from the device
        out["finals"].append(r["final_claims"])
        return r
    def _sc_scaled(ra_scaled, run_claim, scale, ni):
        # cycle rounds of a read-raf / identity range check over ra * scale; the instance caches ra(r) itself (mod.rs:562-578,
        # identity_range_check.rs:326-341), so this call runs with the drivers' appends off and the unscaled claim is appended here
        ctx._lib.ja_set_cache_openings(ctx._h, 0)
        r = _sc(ctx, A.EvalKernel.IDENT, [ra_scaled], run_claim, t, eq_w=ni.eq_w)
        ctx._lib.ja_set_cache_openings(ctx._h, 1)
        ra_claim = A.fr_div(r["final_claims"][0], scale)
        A.transcript_append_scalar_each(ctx, t, ra_claim)
        out["finals"].append(ra_claim.reshape(1, 4))
        ra_scaled.free()
    # A. witness commitment of EVERY one-hot polynomial before the IOP, as ONNXProof::prove does
    #    (mod.rs:152-200 step 4: commit_witness_polynomials -> prover.rs:236-249 -> hyperkzg/mod.rs:558-596)
    hots, wits = [], []
    if resident:
        for i, ni in enumerate(inputs["nodes"]):
            hots.append((resident["nodes"][i]["hot16"], resident["nodes"][i]["hot4"]))
    else:
        # generate_node_witnesses on the device (witness.rs:142-214): the operands go up, the address batches and the clamp lookup
        # indices are born in HBM - no index array crosses PCIe
        for ni in inputs["nodes"]:
            op, S = witness_op(ni.spec)
            ta = A.TensorI32(ctx, ni.A if ni.A.ndim == 2 else ni.A.reshape(1, -1))
            tb = A.TensorI32(ctx, ni.B if ni.B.ndim == 2 else ni.B.reshape(1, -1))
            w = A.FusedWitness(ctx, op, ta, tb, S, 1 << ni.spec.log_t)
            wits.append((w, ta, tb))
            hots.append((w.clamp, w.rem))
    sharded = comm is not None and comm.world > 1
    in_lib = sharded and getattr(comm, "in_library", False)
    if in_lib:
        comm.shard_on()        # from here every MSM / commitment of the context is split over the ranks and combined inside the library
    all_hots = [h for pair in hots for h in pair if h is not None]
    if sharded and not in_lib:
        from . import parallel as PAR0
        out["commitments"] = PAR0.sharded_commit_one_hot_batches(ctx, srs, all_hots, comm)
    else:
        out["commitments"] = A.commit_one_hot_batches(ctx, srs, all_hots)
    for i, ni in enumerate(inputs["nodes"]):
        spec = ni.spec
        res = resident["nodes"][i] if resident else None
        hot16, hot4 = hots[i]
        # B. clamp lookup read-raf (ps_shout/mod.rs; ReadRafSumcheckProver over the saturating clamp table + UnaryRafPS): the 64
        #    address rounds in 8 phases - the T-sized passes at the phase boundaries on the device (init_phase, init_suffix_polys,
        #    raf init_Q), the 256-entry rounds on the host next to the transcript (ja_psshout_prove_address) - then the log T cycle
        #    rounds on the materialised ra * (val + raf_val) (mod.rs:420-446, :464-488).  A REAL sumcheck: the input claim is the
        #    prover's own rv(r_cycle) + gamma * operand(r_cycle); the cycle rounds continue from the running claim.
        #    ps_shout=False (bench.py's continuity leg): the round-1 stage list - cycle rounds on the first RA polynomial, no phase passes.
        if ps_shout:
            ps = res["ps"].restart() if res else wits[i][0].ps_shout(ni.eq_w, CLAMP_LOG_K, PS_PHASES)
            pa = ps.prove_address(t, ni.gammas[0], SAT_BOUND)
            out["msg_bytes"] += pa["msg_bytes"] + PS_PHASES * len(PS_SUFFIXES) * (1 << (CLAMP_LOG_K // PS_PHASES)) * 32   # round polynomials + the Q rows of every phase (mapped memory)
            out["finals"].append(np.stack([pa["val"], pa["raf_val"], pa["claim"]]))
            scale = A.fr_add(pa["val"], pa["raf_val"])
            ra_ps = ps.materialize_ra(scale=scale)
            if not res:
                ps.free()
            _sc_scaled(ra_ps, pa["claim"], scale, ni)
        # C. RA one-hot checks of the clamp lookup (batched: product of 16, Hamming weight, booleanity)
        ra0 = _ra_checks(A, ctx, hot16, ni, 0, D_CLAMP, claim, t, out, _sc, keep_first=not ps_shout)
        if not ps_shout:
            _sc(ctx, A.EvalKernel.IDENT, [ra0], claim, t, eq_w=ni.eq_w)
            ra0.free()
        # D. the operator's own sumcheck
        if spec.kind == "einsum":
            # EinsumDotProver::initialize (einsum/dot.rs:259-283): fold both operands with the eq tables, then log k dot rounds
            eq_r = A.EqPolynomial.evals(ctx, ni.eq_rows)
            eq_c = A.EqPolynomial.evals(ctx, ni.eq_cols)
            if res:                                                        # weights / node inputs resident on the device
                left, right = res["A"].fold(eq_r, transpose=False), res["B"].fold(eq_c, transpose=True)
            else:                                                          # the tensors uploaded for the witness serve the folds too
                left, right = wits[i][1].fold(eq_r, transpose=False), wits[i][2].fold(eq_c, transpose=True)
            _sc(ctx, A.EvalKernel.DOT2, [left, right], claim, t)
            A.MultilinearPolynomial.free_many([eq_r, eq_c, left, right])
        else:
            if res:
                a, b = res["A"].clone(), res["B"].clone()
            else:
                a, b = A.MultilinearPolynomial.from_i32_many(ctx, np.stack([ni.A, ni.B]))
            _sc(ctx, A.EvalKernel.MUL if spec.kind == "mul" else A.EvalKernel.ADD, [a, b], claim, t, eq_w=ni.eq_w)
            A.MultilinearPolynomial.free_many([a, b])
        if hot4 is not None:
            # F. remainder RA checks (batched: product of d = 4, Hamming weight, booleanity)
            rem0 = _ra_checks(A, ctx, hot4, ni, D_CLAMP, ni.d_hot, claim, t, out, _sc, keep_first=not ps_shout)
            # E. remainder range check (IdentityRCProver, identity_range_check.rs:140-325): LOG_K = 14 address rounds in 7 phases
            #    (phase passes on the device, 4-entry rounds on the host), then the cycle rounds on ra * raf_val (:253-283)
            if ps_shout:
                rp = res["ps_rem"].restart() if res else wits[i][0].rem_shout(ni.eq_w, device_rc_phases(MODEL_SCALE))
                pr = rp.prove_identity_rc(t)
                out["msg_bytes"] += pr["msg_bytes"] + rp.phases * 2 * rp.m * 32
                out["finals"].append(np.stack([pr["raf_val"], pr["claim"]]))
                ra_rem = rp.materialize_ra(scale=pr["raf_val"])
                if not res:
                    rp.free()
                _sc_scaled(ra_rem, pr["claim"], pr["raf_val"], ni)
            else:
                _sc(ctx, A.EvalKernel.IDENT, [rem0], claim, t, eq_w=ni.eq_w)
                rem0.free()
        # H. the layer's operators without lookups (mask, softmax normalisation, layer norms): one split-eq sumcheck each
        if ps_shout:
            for j, ax in enumerate(inputs.get("aux", {}).get(i, ())):
                if res:
                    polys = [p_.clone() for p_ in resident["aux"][i][j]]
                else:
                    polys = A.MultilinearPolynomial.from_i32_many(ctx, ax.polys)
                _sc(ctx, ax.kind, polys, claim, t, eq_w=ax.eq_w, gammas=ax.aux)
                A.MultilinearPolynomial.free_many(polys)
        out["states"].append(t.state)
    # G. prove_reduced_openings (prover.rs:141-176): ONE BatchedSumcheck over every committed polynomial
    #    (opening_proof.rs:500-532), gamma powers (:611-643), the materialised RLC (rlc_polynomial.rs:13-78) and the
    #    single HyperKZG opening at r_sumcheck.  Every polynomial is opened at its node's (r_address, r_cycle).
    #    The reduction's instances cache nothing: its claims go to the transcript as ONE vector (:611-617), below.
    ctx._lib.ja_set_cache_openings(ctx._h, 0)
    groups, batches = [], []
    for (hot16, hot4), ni in zip(hots, inputs["nodes"]):
        for h, lo, hi in ((hot16, 0, D_CLAMP), (hot4, D_CLAMP, ni.d_hot)):
            if h is not None:
                groups.append({"kind": A.InstanceKind.OPENING_ONEHOT, "addr": h, "eq_w": ni.eq_w, "r_address": ni.r_addr,
                               "claims": np.broadcast_to(claim, (hi - lo, 4))})
                batches.append(h)
    r = A.batched_sumcheck_prove(ctx, groups, t)
    out["msg_bytes"] += r["msg_bytes"]
    claims = np.concatenate(r["final_claims"])
    out["finals"].append(claims)
    tr = PAR.Transcript(state=t.state, n_rounds=t.n_rounds)
    tr.append_scalars(claims)                                            # opening_proof.rs:628
    gammas = tr.challenge_scalar_powers(claims.shape[0])                 # :630
    t.state, t.n_rounds = tr.state, tr.n_rounds
    rlc = A.MultilinearPolynomial.zeros(ctx, 1 << inputs["ell"])
    o = 0
    for h in batches:
        rlc.rlc_add_onehot(h, gammas[o:o + h.d])
        o += h.d
    if sharded and not in_lib:
        out["open"] = PAR.sharded_hyperkzg_open(ctx, srs, rlc, r["challenges"], t, comm)
    else:
        out["open"] = A.hyperkzg_open(ctx, srs, rlc, r["challenges"], t)    # PCS::prove(rlc, r_sumcheck), prover.rs:164-170
    rlc.free()
    for w, ta, tb in wits:
        w.free(); ta.free(); tb.free()
    out["states"].append(t.state)
    if in_lib:
        comm.shard_off()
    return out


def A_free_many(polys):
    from . import api as A
    A.MultilinearPolynomial.free_many(polys)


def make_resident(ctx, inputs):
    """Upload every per-proof input once (the device-resident leg of the bench reuses these)."""
    from . import api as A
    nodes = []
    for ni in inputs["nodes"]:
        # the node's witness stays on the device (lookup indices of both ps_shout instances): a pass starts its states device to device
        op, S = witness_op(ni.spec)
        ta = A.TensorI32(ctx, ni.A if ni.A.ndim == 2 else ni.A.reshape(1, -1))
        tb = A.TensorI32(ctx, ni.B if ni.B.ndim == 2 else ni.B.reshape(1, -1))
        wit = A.FusedWitness(ctx, op, ta, tb, S, 1 << ni.spec.log_t)
        ta.free(); tb.free()
        d = {"wit": wit, "ps": _ResidentPs(ctx, A, ni, wit),
             "ps_rem": _ResidentPs(ctx, A, ni, wit, rem=True) if ni.d_hot > D_CLAMP else None,
             "hot16": A.OneHotAddresses(ctx, ni.hot_k[:D_CLAMP], K_CHUNK),
             "hot4": A.OneHotAddresses(ctx, ni.hot_k[D_CLAMP:], K_CHUNK) if ni.d_hot > D_CLAMP else None}
        if ni.spec.kind != "einsum":
            d["A"] = A.MultilinearPolynomial.from_i32(ctx, ni.A)
            d["B"] = A.MultilinearPolynomial.from_i32(ctx, ni.B)
        else:
            d["A"] = A.TensorI32(ctx, ni.A)
            d["B"] = A.TensorI32(ctx, ni.B)
        nodes.append(d)
    aux = {i: [[A.MultilinearPolynomial.from_i32(ctx, col) for col in ax.polys] for ax in lst] for i, lst in inputs.get("aux", {}).items()}
    return {"nodes": nodes, "aux": aux}


class _ResidentPs:
    """A fresh ps_shout state per pass over the lookup indices of the node's device-resident witness (device-to-device copy)."""

    def __init__(self, ctx, A, ni, wit, rem=False):
        self.ctx, self.A, self.ni, self.wit, self.cur, self.rem = ctx, A, ni, wit, None, rem

    def restart(self):
        self.free()
        if self.rem:
            self.cur = self.wit.rem_shout(self.ni.eq_w, device_rc_phases(MODEL_SCALE))
        else:
            self.cur = self.wit.ps_shout(self.ni.eq_w, CLAMP_LOG_K, PS_PHASES)
        return self.cur

    def free(self):
        if self.cur is not None:
            self.cur.free()
            self.cur = None


def free_resident(res):
    for lst in res.get("aux", {}).values():
        for polys in lst:
            A_free_many(polys)
    for d in res["nodes"]:
        d["ps"].free()
        if d["ps_rem"] is not None:
            d["ps_rem"].free()
        d["wit"].free()
        d["hot16"].free()
        if d["hot4"] is not None:
            d["hot4"].free()
        for k in ("A", "B"):
            if k in d:
                d[k].free()


def count_units(inputs) -> dict:
    """Work units of one pass (for throughput figures): sumcheck rounds, one-hot point additions, MSM pairs."""
    rounds = adds = 0
    for ni in inputs["nodes"]:
        lt = ni.spec.log_t
        adds += ni.d_hot * (1 << lt)
        rounds += (LOG_K + lt) + lt + ((LOG_K + lt) + lt if ni.d_hot > D_CLAMP else 0)     # batched RA checks + cycle rounds
        rounds += (ni.spec.k - 1).bit_length() if ni.spec.kind == "einsum" else lt
    rounds += sum(ax.log_t for lst in inputs.get("aux", {}).values() for ax in lst)      # mask / div / rsqrt / linear bodies
    rounds += inputs["ell"]                                             # the batched opening reduction
    n_rem = sum(1 for ni in inputs["nodes"] if ni.d_hot > D_CLAMP)
    # T-sized passes the device actually runs for the clamp read-raf: the sign-extension phases of small signed lookup values are
    # built on the host from the phase-0 pass (ja_psshout_prove_address), see DESIGN 4b
    dev_passes = 0
    for ni in inputs["nodes"]:
        v = ni.acc.view(np.int64)
        mag = np.where(v < 0, ~v, v).astype(np.uint64)
        sig = int(mag.max()).bit_length() if mag.size else 0
        log_m = CLAMP_LOG_K // PS_PHASES
        h = sum(1 for ph in range(PS_PHASES) if CLAMP_LOG_K - (ph + 1) * log_m >= sig) if sig <= SAT_BOUND else 0
        dev_passes += PS_PHASES - (h - 1 if h >= 2 else 0)
    addr = CLAMP_LOG_K * len(inputs["nodes"]) + MODEL_SCALE * n_rem     # read-raf + remainder range-check address rounds (host, small tables)
    return {"sumcheck_rounds": rounds + addr, "sumcheck_rounds_device": rounds, "ps_shout_address_rounds": addr,
            "onehot_point_additions": adds, "open_msm_pairs": 4 << inputs["ell"],
            "ps_shout_phase_passes": PS_PHASES * len(inputs["nodes"]) + device_rc_phases(MODEL_SCALE) * n_rem,
            "ps_shout_device_passes": dev_passes + device_rc_phases(MODEL_SCALE) * n_rem}


# ---- measurement helpers (bench.py) -------------------------------------------------------------------------------
def sumcheck_list(inputs):
    """(round body, number of polynomials, initial length) of every device sumcheck instance of one pass, in order."""
    out = []
    for ni in inputs["nodes"]:
        T = 1 << ni.spec.log_t
        out.append(("product", D_CLAMP, T))       # RaVirtual d = 16
        out.append(("booleanity", D_CLAMP, T))    # Booleanity phase 2 over 16 H polynomials
        out.append(("split_eq", 1, T))            # lookup cycle rounds
        if ni.spec.kind == "einsum":
            out.append(("dot", 2, 1 << (ni.spec.k - 1).bit_length()))
        else:
            out.append(("split_eq", 2, T))
        if ni.d_hot > D_CLAMP:
            out.append(("product", D_REM, T))
            out.append(("booleanity", D_REM, T))
            out.append(("split_eq", 1, T))
    for ni in inputs["nodes"]:
        out.append(("opening", ni.d_hot, 1 << ni.spec.log_t))      # batched opening reduction, cycle rounds
    return out


def algorithmic_bytes(inputs) -> dict:
    """Algorithmic HBM bytes of one pass per kernel class.  A fused round kernel over arrays of current length
    L reads 32 L and writes 16 L per polynomial (bind) and evaluates the bound arrays in the same pass; the first round of
    an instance only reads (32 n0).  `bind` = the plain bind launches: last bind of every instance + the HyperKZG folds."""
    b = {"bind": 0, "sumcheck_fused": 0, "onehot_point_sum": 0, "convert_gather": 0, "scatter_add": 0}
    for body, npoly, n0 in sumcheck_list(inputs):
        rounds = n0.bit_length() - 1
        b["sumcheck_fused"] += npoly * (32 * n0 + 48 * sum(n0 >> (j - 1) for j in range(1, rounds)))
        b["bind"] += npoly * 48 * 2
    n = 1 << inputs["ell"]
    b["bind"] += 48 * sum(n >> j for j in range(inputs["ell"] - 1))        # HyperKZG folds (hyperkzg/mod.rs:415-428)
    for ni in inputs["nodes"]:
        T = 1 << ni.spec.log_t
        b["onehot_point_sum"] += ni.d_hot * T * (4 + 64)                   # address + gathered affine base
        b["convert_gather"] += 2 * ni.d_hot * T * (4 + 32)                 # RA + H materialisations
        b["scatter_add"] += ni.d_hot * T * (4 + 32)                        # G tables: address + eq value per entry
    return b


def algorithmic_fieldmuls(inputs) -> dict:
    """Nominal field multiplications of the compute-bound classes: 10 Fq-mul per mixed addition (XYZZ);
    per pair of a round: d^2 for a product-of-d body (+d for the fused bind), 5 per polynomial for booleanity,
    4 + 2 for MUL, 2 + npoly for the linear bodies."""
    adds = sum(ni.d_hot * (1 << ni.spec.log_t) for ni in inputs["nodes"])
    n = 1 << inputs["ell"]
    fused = 0
    for body, npoly, n0 in sumcheck_list(inputs):
        pairs = n0 - 1                                # sum over the rounds of the pairs evaluated
        per_pair = {"product": npoly * npoly + npoly, "booleanity": 5 * npoly, "split_eq": 2 * npoly + 2, "dot": 4,
                    "opening": 3 * npoly}[body]
        fused += per_pair * pairs
    # HyperKZG open: MSMs of 2^(ell-1), ..., 2 pairs (phase 1) and 3 x 2^ell (witness); one mixed addition per scalar and
    # window: 13 windows of 20 bits through the wide fixed-base table (SRS >= 2^21 points, job >= 2^21), else 16 of 16 bits
    def windows(size):
        return 13 if (n >= (1 << 21) and size >= (1 << 21)) else 16
    msm_adds = 3 * n * windows(n) + sum((n >> j) * windows(n >> j) for j in range(1, inputs["ell"]))
    return {"onehot_point_sum": 10 * adds, "msm_accumulate": 10 * msm_adds, "sumcheck_fused": fused}


def config_dict(config: str, inputs, world: int = 1, shard: bool = False) -> dict:
    u = count_units(inputs)
    n_aux = sum(len(v) for v in inputs.get("aux", {}).values())
    return {"workload": "%s-shaped prove pass: %d fused nodes (witness generation, one-hot commits K=16, read-raf / RA / operator / range-check "
                        "sumchecks) + %d mask / div / rsqrt / linear operator sumchecks, chained Blake2b transcript, HyperKZG open ell=%d; "
                        "synthetic i8-range tensors" % (config, len(inputs["nodes"]), n_aux, inputs["ell"]),
            "sumcheck_rounds": u["sumcheck_rounds"], "onehot_point_additions": u["onehot_point_additions"],
            "open_msm_pairs": u["open_msm_pairs"],
            "l2": "no explicit flush: one pass streams %.0f MB of polynomial data through a 126 MB L2" %
                  (sum(algorithmic_bytes(inputs).values()) / 1e6),
            "parallelism": "1 GPU" if world == 1 else
                           ("one proof on %d GPUs: commitments dealt by polynomial, opening MSMs split by index range, partial points exchanged inside the "
                            "library (ncclAllGather on the context stream), Fiat-Shamir-sequential sumchecks replicated" % world if shard else
                            "%d replicas, one independent proof per GPU (no data-path collective)" % world)}


def roofline_from_profile(prof: dict, inputs, peaks: dict, peaks_kind: str, ctx, sweep: bool = True) -> dict:
    """Per-class achieved rates of one profiled pass + a large-n sweep of the streaming kernels.  A class that has both an
    algorithmic byte count and a field-mul count is bound by whichever roofline time is larger."""
    ab = algorithmic_bytes(inputs)
    am = algorithmic_fieldmuls(inputs)
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    hbm = float(peaks["hbm_gbs"])
    mul_peak = ctx.calibrate_fr_mul(2000) / 1e9       # Gmul/s, register-resident Montgomery loop on this GPU
    classes = {}
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4)}
        if ab.get(name) and v["ms"] > 0:
            row["algorithmic_bytes"] = ab[name]
            row["GBps"] = round(ab[name] / v["ms"] / 1e6, 2)
        if am.get(name) and v["ms"] > 0:
            row["field_muls"] = am[name]
            row["Gmul_per_s"] = round(am[name] / v["ms"] / 1e6, 2)
        classes[name] = row
    dom_name = next(iter(classes))
    dom = classes[dom_name]
    t_hbm = dom.get("algorithmic_bytes", 0) / (hbm * 1e9)
    t_mul = dom.get("field_muls", 0) / (mul_peak * 1e9)
    common = {"kernel": dom_name, "traffic": None, "per_launch_ms": round(dom["ms"] / dom["launches"], 5),
              "share_of_step": dom["share"], "roofline_ms": round(max(t_hbm, t_mul) * 1e3, 4), "measured_ms": dom["ms"]}
    if t_mul > t_hbm:
        dominant = dict(common, bound="int32 field-mul (no tensor cores on this path)", achieved=dom["Gmul_per_s"],
                        peak=round(mul_peak, 2), unit="Gmul/s", frac=round(dom["Gmul_per_s"] / mul_peak, 4),
                        peak_source="calibrated live: register-resident Montgomery product loop")
    else:
        gb = dom.get("GBps", 0.0)
        dominant = dict(common, bound="hbm", achieved=gb, peak=hbm, unit="GB/s", frac=round(gb / hbm, 4),
                        peak_source="%s (MEASURED_PEAKS.json hbm_gbs)" % peaks_kind)
    do_sweep, sweep = sweep, []
    if not do_sweep:
        return {"dominant": dominant, "classes": classes, "sweep": sweep, "fr_mul_peak_Gmul_s": mul_peak}
    for which, name, bytes_per_n in ((0, "bind_low_to_high", 48), (1, "bind_high_to_low", 48), (4, "round_eval_add", 64), (2, "round_eval_mul", 64)):
        for log_n in (24, 26):
            ms = ctx.bench_kernel(which, log_n, 1, 10)
            gb = bytes_per_n * (1 << log_n) / ms / 1e6
            sweep.append({"kernel": name, "log_n": log_n, "ms": round(ms, 4), "GBps": round(gb, 1), "frac_hbm": round(gb / hbm, 4)})
    # fused bind + evaluate round kernels: 48 n bytes per polynomial and launch; the mul-bound bodies also as Gmul/s.
    # Products per pair = (full Montgomery products, 4-row challenge products of the bind) THE KERNEL EXECUTES (ADD / IDENT weight one
    # eq-free value per pair with e_in; e_out is applied once per x_out group): a challenge product is half the IMAD.WIDE rows of a
    # full one and is counted as 0.5, and so is the UNREDUCED full product of the TMA-staged kernels (delayed Montgomery reduction: 64 of
    # 136 rows), so that frac_mul is a fraction of the calibrated full-product peak and stays below 1.  product16: 10 products per lane
    # since the quadratic first stage (lane_product<16>, was 15) + the e_in weighting; its ~30 field additions per lane are not counted.
    fused = ((7, "fused_round_add_tma", 2, 24, (0.5, 4)), (7, "fused_round_add_tma", 2, 26, (0.5, 4)), (8, "fused_round_ident_tma", 1, 26, (0.5, 2)),
             (0, "fused_round_add", 2, 24, (1, 4)), (1, "fused_round_mul", 2, 24, (4, 4)), (2, "fused_round_ident", 1, 24, (1, 2)),
             (6, "fused_round_open_h2l", 1, 24, (2, 2)), (3, "fused_round_product4", 4, 22, (16 + 4, 8)), (4, "fused_round_product16", 16, 20, (160 + 16, 32)),
             (5, "fused_round_booleanity16", 16, 20, (16 * 4 + 2, 32)))
    for which, name, npoly, log_n, (full_muls, chal_muls) in fused:
        ms = ctx.bench_fused(which, log_n, 10)
        n = 1 << log_n
        gb = 48 * n * npoly / ms / 1e6
        gm = (full_muls + 0.5 * chal_muls) * (n // 4) / ms / 1e6
        sweep.append({"kernel": name, "log_n": log_n, "n_polys": npoly, "ms": round(ms, 4), "GBps": round(gb, 1),
                      "frac_hbm": round(gb / hbm, 4), "full_products_per_pair": full_muls, "challenge_products_per_pair": chal_muls,
                      "Gmul_per_s": round(gm, 2), "frac_mul": round(gm / mul_peak, 4)})
    return {"dominant": dominant, "classes": classes, "sweep": sweep, "fr_mul_peak_Gmul_s": mul_peak}
