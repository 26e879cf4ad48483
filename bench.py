#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the proving hot path.

  python bench.py --gpus N --steps K --warmup W                  this repo's CUDA path (C ABI through jolt_atlas_b200)
  python bench.py --impl reference --gpus N --steps K --warmup W  the CPU restatement (oracle/, OpenMP) on the host cores

A "step" is ONE prove-shaped pass (jolt_atlas_b200/workload.py): for every node of the nanoGPT-shaped graph the one-hot
witness commitments, the lookup / RA / operator / range-check sumchecks with a chained Blake2b transcript, then one
HyperKZG opening (ell = 18).  metric = seconds per pass (BASELINE.json: "ONNXProof::prove sec (nanoGPT ...)").

  value     device-resident inputs (uploaded once before the timed region), CUDA events on the library's stream
  e2e       same pass through the public API with HOST (pinned) buffers: H2D of every input and D2H of every
            commitment / round polynomial / claim inside the timed region
  roofline  the dominant kernel class of the pass, timed live with CUDA events (ja_profile_*), vs MEASURED_PEAKS.json
  cpu_baseline  the C++ oracle on the host cores, bounded sample (rank 0, N = 1 only)

No number here is taken under a profiler.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
TAU = 0x1234567890abcdef1122334455667788          # fixed test "toxic waste": SRS = [tau^i] G (SURVEY §8c SRS note)
def metric_name(config: str) -> str:
    return "ONNXProof::prove sec (%s-shaped prove pass)" % config


def _limbs(x: int, n: int = 4) -> list[int]:
    return [(x >> (64 * k)) & ((1 << 64) - 1) for k in range(n)]


def g1_generator_mont() -> np.ndarray:
    rq = (1 << 256) % Q
    return np.array(_limbs(rq) + _limbs(2 * rq % Q), dtype=np.uint64)       # (1, 2) in Montgomery form


def tau_mont() -> np.ndarray:
    from jolt_atlas_b200.workload import P, R
    return np.array(_limbs(TAU * R % P), dtype=np.uint64)


def peaks() -> tuple[dict, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md, the clocks line)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def pin_inputs(inputs):
    """Move the per-proof host inputs into pinned memory (torch is plumbing here: cudaHostAlloc)."""
    import torch

    def pin(a: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    keep = []
    for ni in inputs["nodes"]:
        for name in ("hot_k", "tables", "A", "B", "eq_w", "gammas", "eq_rows", "eq_cols"):
            v = getattr(ni, name)
            if v is not None:
                setattr(ni, name, pin(v))
    inputs["open_point"] = pin(inputs["open_point"])
    return keep


def d2h_bytes(result) -> int:
    total = 0
    for com, inf in result["commitments"]:
        total += com.nbytes + np.asarray(inf).nbytes
    for f in result["finals"]:
        total += f.nbytes
    total += result.get("msg_bytes", 0)
    for k in ("com", "v", "w"):
        total += result["open"][k].nbytes
    return total


# ---------------------------------------------------------------------------------------------------------------
LAST_CPU_PASS = {"out": None}      # result of the last FULL oracle pass (parity check of the device arm)


def parity_with_cpu_pass(dev_out, cpu_out) -> bool:
    """The device pass and the oracle pass of the same inputs agree bit for bit: every commitment, every final claim, the
    transcript state after every node (which pins every round polynomial and challenge) and the HyperKZG opening."""
    if cpu_out is None or dev_out is None:
        return False
    ok = dev_out["states"] == cpu_out["states"] and len(dev_out["finals"]) == len(cpu_out["finals"])
    ok = ok and all(np.array_equal(a, b) for a, b in zip(dev_out["finals"], cpu_out["finals"]))
    ok = ok and len(dev_out["commitments"]) == len(cpu_out["commitments"])
    ok = ok and all(np.array_equal(a[0], b[0]) and np.array_equal(np.asarray(a[1], dtype=bool), np.asarray(b[1], dtype=bool))
                    for a, b in zip(dev_out["commitments"], cpu_out["commitments"]))
    ok = ok and all(np.array_equal(dev_out["open"][k], cpu_out["open"][k]) for k in ("com", "v", "w"))
    return bool(ok)


def cpu_pass_seconds(srs_host, inputs, rlc_host, budget_s: float):
    """Time the C++ oracle (OpenMP, all host threads) on the workload.  Full pass when it fits the budget, else
    layer 0 x (number of identical layers) + lm_head + the opening stage (reduction sumcheck, RLC, HyperKZG open),
    each timed once."""
    from oracle import cpu as ORC
    from oracle import workload_cpu as WC
    ORC.set_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1
    nodes = inputs["nodes"]
    per_layer = 10
    t0 = time.perf_counter()
    WC.run_cpu(srs_host, inputs, node_limit=per_layer, do_open=False)
    t_layer = time.perf_counter() - t0
    n_layers = (len(nodes) - 1) // per_layer
    if t_layer * n_layers * 1.6 <= budget_s:
        t0 = time.perf_counter()
        full_out = WC.run_cpu(srs_host, inputs)
        LAST_CPU_PASS["out"] = full_out
        return time.perf_counter() - t0, "full pass: %d nodes + opening reduction + HyperKZG open ell=%d" % (len(nodes), inputs["ell"])
    head = dict(inputs)
    head["nodes"] = nodes[n_layers * per_layer:]
    t0 = time.perf_counter()
    WC.run_cpu(srs_host, head, do_open=False)
    t_head = time.perf_counter() - t0
    t0 = time.perf_counter()
    WC.run_cpu(srs_host, inputs, iop=False)
    t_open = time.perf_counter() - t0
    return (t_layer * n_layers + t_head + t_open,
            "layer 0 (%d of %d nodes) timed once and counted x%d, + lm_head, + opening reduction / RLC / HyperKZG open ell=%d over "
            "all nodes, each timed once" % (per_layer, len(nodes), n_layers, inputs["ell"]))


def cpu_sample_large(inputs):
    """Bounded CPU sample for the GPT-2-sized pass (one measurement, ~1-2 minutes): the oracle cannot even generate the 2^24-point
    SRS in the time budget.  Layer 0 (10 nodes, T <= 2^16, SRS 2^20) is timed and counted once per layer; the lm_head node
    (T = 2^20) is counted as 16 x the same-shaped T = 2^16 einsum node (every stage of a node is linear in T); the opening stage
    (reduction sumcheck, RLC, HyperKZG open) is timed over layer 0's polynomials at ell = 20 and scaled by 2^(ell - 20), which is
    also the ratio of all committed elements to layer 0's.  The figure is an EXTRAPOLATION and labelled as one."""
    from oracle import cpu as ORC
    from oracle import workload_cpu as WC
    ORC.set_threads(os.cpu_count() or 1)
    nodes = inputs["nodes"]
    per_layer = 10
    n_layers = (len(nodes) - 1) // per_layer
    srs = ORC.srs_powers(tau_mont(), 1 << 20)
    t0 = time.perf_counter()
    WC.run_cpu(srs, inputs, node_limit=per_layer, do_open=False)
    t_layer = time.perf_counter() - t0
    one = dict(inputs); one["nodes"] = nodes[:1]
    t0 = time.perf_counter()
    WC.run_cpu(srs, one, do_open=False)
    t_node0 = time.perf_counter() - t0
    sub = dict(inputs); sub["nodes"] = nodes[:per_layer]; sub["ell"] = 20
    t0 = time.perf_counter()
    WC.run_cpu(srs, sub, iop=False)
    t_open = time.perf_counter() - t0
    scale = 1 << (inputs["ell"] - 20)
    return (t_layer * n_layers + 16 * t_node0 + t_open * scale,
            "EXTRAPOLATED from one bounded sample: layer 0 (%d of %d nodes) timed once and counted x%d; lm_head (T=2^20) = 16 x the T=2^16 einsum node; "
            "opening stage timed over layer 0 at ell=20 and scaled x%d (measured %.1f + %.1f + %.1f s)" %
            (per_layer, len(nodes), n_layers, scale, t_layer, t_node0, t_open))


def synthetic_rlc_host(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)           # canonical (< p)
    return a


def run_reference(args):
    """--impl reference: the CPU restatement of the same pass on the host cores (oracle/cpp, OpenMP).  The reference
    itself is Rust with un-vendored git dependencies and cannot be built in this image."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu as ORC
    from oracle import workload_cpu as WC
    from jolt_atlas_b200 import workload as W
    ORC.set_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1
    inputs = W.build_inputs(args.config)
    n = 1 << inputs["ell"]
    total_steps = args.steps + args.warmup
    if inputs["ell"] > 20:
        # GPT-2 size: one bounded, extrapolated sample stands for every step (a single CPU pass would take tens of minutes)
        val, sample = cpu_sample_large(inputs)
        full, times = False, [val]
    else:
        srs_host = ORC.srs_powers(tau_mont(), n)
        rlc_host = None
        budget = 150.0 / max(total_steps, 1)
        # calibrate once, then decide between the full pass and the bounded sample
        secs, sample = cpu_pass_seconds(srs_host, inputs, rlc_host, budget)
        full = sample.startswith("full")
        times = []
        # a FULL pass is timed every step while the whole run fits ~2.5 minutes; otherwise the steps already measured stand
        t_run0 = time.perf_counter()
        for i in range(total_steps):
            if time.perf_counter() - t_run0 > 150.0 and times:
                break
            if full:
                t0 = time.perf_counter()
                WC.run_cpu(srs_host, inputs)
                dt = time.perf_counter() - t0
            else:
                dt, _ = cpu_pass_seconds(srs_host, inputs, rlc_host, 0.0)
            if i >= args.warmup or not times:
                times.append(dt)
        val = float(np.mean(times)) if times else secs
        if not full:
            sample = "EXTRAPOLATED: " + sample
        sample += " (%d timed passes)" % len(times)
    cores = ORC.num_threads()
    line = {"impl": "reference", "metric": metric_name(args.config) + ("" if full else " [CPU value extrapolated from a bounded sample]"), "value": val, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "u64x4 Montgomery (BN254 Fr/Fq)", "data": "synthetic",
            "config": W.config_dict(args.config, inputs, args.gpus, args.gpus > 1),
            "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def run_device_arm(args):
    import torch
    import torch.distributed as dist

    from jolt_atlas_b200 import SRS, Context, MultilinearPolynomial
    from jolt_atlas_b200 import workload as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # host worker threads of the library (large-batch glue): split the box's cores between the ranks
    os.environ.setdefault("JA_HOST_THREADS", str(max(1, min(8, (os.cpu_count() or 1) // world))))
    ctx = Context(local)
    # independent proofs per rank (weak scaling: the path has no cross-proof exchange); see DESIGN.md §Multi-GPU
    # --shard: ONE proof on all ranks (same inputs everywhere; commitments and opening MSMs sharded, strong scaling)
    shard = world > 1 and not args.replicas
    comm = None
    if shard:
        from jolt_atlas_b200 import parallel as PAR
        comm = PAR.LibComm(ctx)            # the library's own NCCL communicator: the exchange happens inside the C-ABI calls
    inputs = W.build_inputs(args.config, seed=None if (rank == 0 or shard) else W.CONFIGS[args.config]["seed"] + rank)
    n = 1 << inputs["ell"]
    srs = SRS.generate(ctx, g1_generator_mont(), tau_mont(), n).precompute()
    resident = W.make_resident(ctx, inputs)
    pin_inputs(inputs)
    ctx.sync()

    single_same = None
    if shard:
        # the same proof on ONE GPU (every rank runs it on its own, no exchange): the denominator of the strong-scaling figure
        W.run_device(ctx, srs, inputs, resident=resident)
        barrier()
        ctx.timer_begin()
        for _ in range(2):
            W.run_device(ctx, srs, inputs, resident=resident)
        t1 = torch.tensor([ctx.timer_end() / 2], device="cuda", dtype=torch.float64)
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        single_same = float(t1.item()) / 1e3
    # ---- device-resident leg ----
    for _ in range(args.warmup):
        W.run_device(ctx, srs, inputs, resident=resident, comm=comm)
    sampler = ClockSampler(local)
    barrier()
    if not os.environ.get("JA_BENCH_NO_CLOCKS"):
        sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_begin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        W.run_device(ctx, srs, inputs, resident=resident, comm=comm)
    dev_ms = ctx.timer_end()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    ms = max(dev_ms, 0.0)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps

    # ---- continuity leg (one GPU): the ROUND-1 stage list (no ps_shout phase passes), resident inputs, same timing rules ----
    r1_ms = None
    if world == 1:
        W.run_device(ctx, srs, inputs, resident=resident, ps_shout=False)
        barrier()
        ctx.timer_begin()
        for _ in range(args.steps):
            W.run_device(ctx, srs, inputs, resident=resident, ps_shout=False)
        r1_ms = ctx.timer_end() / args.steps
    # ---- end-to-end leg: host buffers in, proof data out, every step ----
    W.run_device(ctx, srs, inputs, comm=comm)          # warm the upload path
    barrier()
    ctx.timer_begin()
    last = None
    for _ in range(args.steps):
        last = W.run_device(ctx, srs, inputs, comm=comm)
    e2e_ms = ctx.timer_end()
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_per_step = e2e_ms / args.steps
    h2d = W.h2d_bytes(inputs)
    d2h = d2h_bytes(last)

    line = None
    if rank == 0:
        # ---- live per-class kernel profile (one extra pass, not part of any headline number) ----
        # per-kernel event times must not include a kernel's wait for the host: the same round bodies as one launch per round
        # (a round-resident kernel's duration includes the host's transcript time, a pre-launched one its wait for the challenge)
        os.environ["JA_NO_AHEAD"] = "1"
        os.environ["JA_NO_PERSIST"] = "1"
        ctx.profile_begin()
        W.run_device(ctx, srs, inputs, resident=resident)
        prof = ctx.profile_end()
        os.environ.pop("JA_NO_AHEAD", None)
        os.environ.pop("JA_NO_PERSIST", None)
        pk, pk_kind = peaks()
        roof = W.roofline_from_profile(prof, inputs, pk, pk_kind, ctx, sweep=not (args.no_sweep or world > 1))
        # `traffic`: DRAM bytes per launch of the class's dominant kernel form from the committed `ncu --set full` capture of ONE launch
        # (profiles/r2_ncu_pair.json, scripts/r2_measure.sh; cold-cache replay) next to that launch's algorithmic bytes; the class itself
        # mixes thousands of launches of different sizes, whose live average is `per_launch_ms`.
        probe_file = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_ncu_pair.json")
        if os.path.exists(probe_file):
            try:
                forms = json.load(open(probe_file))["forms"]
                wide = forms["wide"]                    # k_round_prod_bool<16, 1, 128, 1>: the form with the largest share of the launch list
                roof["dominant"]["traffic"] = int(wide["dram_read"] + wide["dram_write"])
                roof["dominant"]["traffic_probe"] = {
                    "source": "profiles/r2_ncu_pair.json (ncu --set full, one launch per form, cold-cache replay); `traffic` is the 'wide' form",
                    "launches": [{"kernel": f["kernel"], "pairs": f["pairs"], "dram_bytes": int(f["dram_read"] + f["dram_write"]),
                                  "algorithmic_bytes": f["algorithmic_read"] + f["algorithmic_write"]} for f in forms.values()]}
            except (OSError, KeyError, ValueError):
                pass
        units = W.count_units(inputs)
        div = 1 if shard else world        # replicas: world proofs per step; --shard: one proof per step
        line = {"metric": metric_name(args.config), "value": ms_per_step / 1e3 / div, "unit": "s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong" if shard else "weak",
                "vs_baseline": None, "dtype": "u64x4 Montgomery (BN254 Fr/Fq)", "data": "synthetic",
                "config": W.config_dict(args.config, inputs, world, shard),
                "clocks": clocks,
                "e2e": {"value": e2e_per_step / 1e3 / div, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches),
                "wall_ms_per_step": wall_ms / args.steps,
                "units_per_step": units,
                "roofline": roof["dominant"], "kernel_classes": roof["classes"], "kernel_sweep": roof["sweep"]}
        if r1_ms is not None:
            line["value_round1_stage_list_s"] = r1_ms / 1e3      # the same pass WITHOUT the ps_shout phase passes added in round 2 (BENCH_r01's workload)
        if shard:
            line["one_gpu_same_config_s"] = single_same          # this config on one GPU of the same box, same run
            line["speedup_vs_one_gpu"] = round(single_same / (ms_per_step / 1e3), 3)
            line["limiter"] = ("the Fiat-Shamir chain: %d strictly sequential sumcheck rounds run replicated on every rank; only the group "
                               "arithmetic (commitments, opening MSMs) is divided by the number of GPUs" % units["sumcheck_rounds"])
        line["msm_sweep"] = []
        if world == 1 and not args.no_cpu:
            from oracle import cpu as ORC
            secs, sample = cpu_pass_seconds(srs.to_host(), inputs, None, 30.0)
            line["cpu_baseline"] = {"value": secs, "unit": "s", "cores": ORC.num_threads(), "kind": "port", "sample": sample}
            if LAST_CPU_PASS["out"] is not None:
                # the timed device pass (last end-to-end step) against the oracle's full pass of the same inputs, bit for bit
                line["parity_checked"] = parity_with_cpu_pass(last, LAST_CPU_PASS["out"])
                if not line["parity_checked"]:
                    raise SystemExit("bench.py: the device pass differs from the CPU oracle pass (transcript states / claims / commitments / opening)")
            else:
                line["parity_checked"] = False   # the oracle only ran a bounded sample of this workload
    # ---- MSM sweep (BASELINE.json config 5): Mscalar/s on pseudo-random 254-bit scalars.  One GPU: 2^18 .. 2^26 against an SRS
    # with the fixed-base window tables (up to 2^24: 29x the SRS in HBM) and without them (plain Pippenger, every size), table build
    # time reported, and a CPU Pippenger column (the oracle, all host cores) at the sizes it finishes in seconds.  N GPUs: the pairs
    # split by index range over the ranks (ja_set_msm_shard + the library's communicator), on the pass's resident 2^24 SRS.
    msm_rows = []
    if not args.no_sweep:
        from jolt_atlas_b200 import msm_fr

        def time_msm(srs_h, log_n, reps=3):
            p = MultilinearPolynomial.random(ctx, 1 << log_n, 7)
            msm_fr(ctx, srs_h, p)
            best = 1e9
            for _ in range(reps):
                barrier()
                ctx.timer_begin(); msm_fr(ctx, srs_h, p); best = min(best, ctx.timer_end())
            if world > 1:
                t = torch.tensor([best], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = float(t.item())
            p.free()
            return best
        if world > 1 and shard:
            comm.shard_on()
            for log_n in (20, 22, 24):
                if (1 << log_n) <= len(srs):
                    ms = time_msm(srs, log_n)
                    msm_rows.append({"log_n": log_n, "n_gpus": world, "table": True, "ms": round(ms, 3), "Mscalar_per_s": round((1 << log_n) / ms / 1e3, 1)})
            comm.shard_off()
        elif world == 1:
            W.free_resident(resident); resident = None
            srs.free(); srs = None
            for top, sizes, table in ((24, (18, 20, 22, 24), True), (26, (18, 20, 22, 24, 26), False)):
                t0 = time.perf_counter()
                big = SRS.generate(ctx, g1_generator_mont(), tau_mont(), 1 << top)
                ctx.sync()
                t_gen = time.perf_counter() - t0
                t_tab = None
                if table:
                    t0 = time.perf_counter(); big.precompute(); ctx.sync(); t_tab = time.perf_counter() - t0
                for log_n in sizes:
                    ms = time_msm(big, log_n)
                    msm_rows.append({"log_n": log_n, "n_gpus": 1, "table": table, "srs_log_n": top, "ms": round(ms, 3),
                                     "Mscalar_per_s": round((1 << log_n) / ms / 1e3, 1), "srs_generate_s": round(t_gen, 2),
                                     "table_build_s": round(t_tab, 2) if t_tab is not None else None})
                if table and not args.no_cpu:
                    from oracle import cpu as ORC
                    ORC.set_threads(os.cpu_count() or 1)
                    host_srs = big.to_host(0, 1 << 20)
                    for log_n in (18, 20):
                        sc = synthetic_rlc_host(1 << log_n, 5)
                        t0 = time.perf_counter(); ORC.msm_fr(host_srs[: 1 << log_n], sc); dt = time.perf_counter() - t0
                        msm_rows.append({"log_n": log_n, "impl": "cpu oracle Pippenger (ark-ec style windows, OpenMP)", "cores": ORC.num_threads(),
                                         "ms": round(dt * 1e3, 1), "Mscalar_per_s": round((1 << log_n) / dt / 1e6, 3)})
                big.free()
    if line is not None:
        line["msm_sweep"] = msm_rows
    if resident is not None:
        W.free_resident(resident)
    if srs is not None:
        srs.free()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=["nanoGPT", "microgpt", "gpt2"],
                    help="default: nanoGPT on one GPU (BASELINE.json configs[1]); gpt2 for --gpus N > 1 (configs[3]: one GPT-2-shaped proof on N GPUs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--shard", action="store_true", help="(default for N > 1) one proof on all GPUs: commitments / opening MSMs sharded, exchange inside the library")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent proofs, one per GPU (no exchange) instead of one sharded proof")
    ap.add_argument("--no-sweep", action="store_true", help="skip the large-n kernel sweep and the MSM sweep")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    if args.config is None:
        args.config = "gpt2" if (world > 1 and not args.replicas) else "nanoGPT"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_device_arm(args)


if __name__ == "__main__":
    main()
