#!/usr/bin/env python
"""bench.py — env-steps/s of the fused merging-gym step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" advances every env of the rank's 2^20-env shard by one dT with external uint8 actions (pvp, auto-reset):
BASELINE.json configs[2].  The shard is a `MergeVecEnv(lanes=2)` — a step is two `mg_step` launches of 2^19 envs on the
lanes' two CUDA streams, each ordered only behind its own lane's previous step (`--headline serialized`: one launch per
step on one stream; whichever is not the headline is reported beside it as `serialized` / `laned`).
Actions are pre-generated on the device (Philox, global env ids), so inputs are HBM-resident.
L2 rule: one shard's inputs (56.6 MB) would fit the 126 MB L2, so the timed region steps `--shards` (4) independent
2^20-env shards round-robin — inputs larger than L2, no flush — and every launch reads its state from HBM.  Timed with
CUDA events on the launching (current torch) stream behind a spin-kernel gate (so that the host has queued the region
before the device starts it), barrier + synchronize on both sides, max over ranks.

The JSON line (rank 0) carries, besides the contract keys:
  roofline       HBM, algorithmic bytes = 156 B/env-step
  e2e            the same step through the host-buffer API (`MergeVecEnv.step_host_async/_wait`): actions in pinned
                 host memory, obs/rew/done/info copied back every step; + the synchronous call, the reduced-field
                 variant and the box's measured device->host copy ceiling
  sustained      the same measurement at 2000 steps whatever --steps says (clocks are sampled over it as well)
  serialized, l2_warm, l2_flushed, overlapped_streams, lean_no_returns, rollout_fused      extra device-side figures
  extra_errors   an extra that failed (null when none did): extras never cost the line its contract keys
  policy_in_loop BASELINE configs[4] (2^18 envs, DQN forward in the loop)
  strong_8m      BASELINE configs[3] (2^23 envs over all ranks, NCCL statistics reduction every 64 steps) with a
                 digest of the reduced statistics that must be equal for every world size
  cpu_baseline   (N=1, rank 0) the plain-C oracle port on all host cores, and the reference's OWN unmodified Python
                 env timed on this box (one process, and one process per core)

`--impl reference` times the reference's CPU implementation of the path on this box's host cores: `value` is the
plain-C float64 port of the oracle with all host threads (`kind: "port"`, the stronger baseline); the reference's own
Python `MergeEnv.step` (from the byte-identical copy in baseline/_ref, run against oracle/ref_shims) is timed in the
same run and printed beside it (`cpu_baseline.reference_python_*`).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
BYTES_PER_ENV_STEP = 156          # read 54 B + write 102 B (pvp, u8 actions, auto-reset)
BYTES_PER_ENV_STEP_LEAN = 124     # track_returns=False: the two float64 accumulators are neither read nor written
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
STRONG_TOTAL_ENVS = 1 << 23       # BASELINE configs[3]
STRONG_STEPS, STRONG_REDUCE_EVERY = 256, 64


def workload_config(n):
    """The workload, worded identically by both arms (the driver compares the dicts)."""
    return {"workload": "pvp, 2^20 envs per GPU, auto-reset, uniform-random uint8 actions (BASELINE.json configs[2]); "
                        "the reference arm steps one 2^20-env shard on the host cores regardless of --gpus",
            "envs_per_gpu": n, "mode": "pvp", "auto_reset": True, "actions": "uniform-random uint8, pre-generated",
            "l2": "GPU arm: inputs larger than L2, no flush (4 independent 2^20-env shards per GPU stepped round-robin, "
                  "218 MB of float64 state: every launch re-reads its shard's state from HBM); CPU arm: not applicable"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """Per-launch DRAM bytes of the step kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa
            self.err = repr(e)

    def sample_now(self):
        """One sample from the calling thread (used while the GPU is still busy with the timed work)."""
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            self.sample_now()
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------ CPU legs
def cpu_port_throughput(n_envs, min_seconds, threads, warmup_steps=2, max_steps=None, fixed_steps=None):
    """Times the plain-C oracle port (oracle/merge_oracle.c) on host cores.  Checker code used
    as the reported CPU baseline only — never on the product path."""
    from oracle import c_oracle
    env = c_oracle.CVecEnv(n_envs, pvp=True, auto_reset=True, nthreads=threads)
    a1, a2 = c_oracle.philox_actions(n_envs, 0x5EED, 0, 0)
    for _ in range(warmup_steps):
        env.step(a1, a2)
    steps, t0 = 0, time.perf_counter()
    while True:
        env.step(a1, a2)
        steps += 1
        dt = time.perf_counter() - t0
        if fixed_steps is not None:
            if steps >= fixed_steps:
                break
        elif dt >= min_seconds or (max_steps and steps >= max_steps):
            break
    return n_envs * steps / dt, steps, dt


def python_scalar_port_throughput(seconds=2.0):
    """The scalar Python restatement (one env object, like the reference's own per-env loop)."""
    import numpy as np
    from oracle import merge_oracle as mo
    env = mo.RefEnv()
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, (4096, 2)).tolist()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a = acts[n % 4096]
        _, _, d, _ = env.step(a[0], a[1])
        if d:
            env.reset()
        n += 1
    return n / (time.perf_counter() - t0)


def _reference_python_worker(args):
    """One process stepping the reference's OWN `MergeEnv` (unmodified merging_env.py + helper.py, third-party
    packages replaced by oracle/ref_shims): BASELINE configs[0] workload (random actions, manual reset on done)."""
    pvp, steps, seed = args
    import warnings
    warnings.simplefilter("ignore")
    import numpy as np
    from oracle.ref_loader import load_reference_env, quiet
    env = load_reference_env()
    acts = np.random.default_rng(seed).integers(0, 5, (steps, 2)).tolist()
    env.reset()
    with quiet():                                            # the env prints on every collision (merging_env.py:204)
        t0 = time.perf_counter()
        for a1, a2 in acts:
            _, _, done, _ = env.step(a1, a2 if pvp else None)
            if done:
                env.reset()
        dt = time.perf_counter() - t0
    return steps / dt


def reference_python_throughput(steps=3000, cores=None):
    """env-steps/s of the unmodified reference env on this box: one process, and one process per core (each its own
    env, aggregate over the wall clock of the whole pool incl. start-up of the slowest).  None if the reference copy
    is not present (baseline/_ref is made by baseline/install_ref.py / __graft_entry__.build())."""
    try:
        from oracle import ref_loader
        if not ref_loader.reference_available():
            return {"unavailable": f"no reference tree at {ref_loader.REFERENCE_ROOT}"}
    except Exception as e:  # noqa
        return {"unavailable": repr(e)}
    import multiprocessing as mp
    cores = cores or host_threads()
    out = {"source": ref_loader.REFERENCE_ROOT, "steps_per_process": steps, "cores": cores,
           "note": "the reference's own merging_env.py + helper.py, byte-identical copy; gym/pygame/shapely/qpsolvers are "
                   "not installed, so it runs against oracle/ref_shims (lighter than the real packages: this over-estimates "
                   "the reference's speed)"}
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(1) as pool:
            out["pve_1_process"] = pool.map(_reference_python_worker, [(False, steps, 0)])[0]
            out["pvp_1_process"] = pool.map(_reference_python_worker, [(True, steps, 0)])[0]
        with ctx.Pool(cores) as pool:
            pool.map(_reference_python_worker, [(True, 50, s) for s in range(cores)])       # imports done, workers warm
            for name, pvp in (("pve", False), ("pvp", True)):
                t0 = time.perf_counter()
                pool.map(_reference_python_worker, [(pvp, steps, s) for s in range(cores)])
                out[f"{name}_all_cores"] = cores * steps / (time.perf_counter() - t0)
    except Exception as e:  # noqa
        out["error"] = repr(e)
    return out


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_object(seconds, fixed_steps=None, warmup=2, with_python=True, ref_steps=3000):
    threads = min(host_threads(), 64)
    v, steps, dt = cpu_port_throughput(ENVS_PER_GPU, seconds, threads, warmup_steps=warmup, fixed_steps=fixed_steps)
    cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{ENVS_PER_GPU} envs x {steps} steps in {dt:.1f} s (pvp, auto-reset, pre-generated uint8 actions, "
                     f"float64, plain-C oracle port, {threads} OpenMP threads)",
           "ms_per_step": 1e3 * dt / steps}
    if with_python:
        cpu["single_thread_value"] = cpu_port_throughput(1 << 16, min(3.0, seconds), 1)[0]
        cpu["python_scalar_port_value"] = python_scalar_port_throughput(2.0)
        ref = reference_python_throughput(ref_steps) if ref_steps > 0 else {"unavailable": "--ref-python-steps 0"}
        cpu["reference_python"] = ref
        cpu["reference_python_value"] = ref.get("pvp_1_process")
        cpu["reference_python_all_cores_value"] = ref.get("pvp_all_cores")
        cpu["note"] = ("value = the C port (the strongest CPU baseline available); reference_python_* = the reference's "
                       "own Python env on this box (pvp, like the GPU workload; pve = BASELINE configs[0] is inside "
                       "reference_python); python_scalar_port_value = the oracle's scalar Python restatement")
    return cpu, dt, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    n = ENVS_PER_GPU
    cpu, dt, steps = cpu_baseline_object(0, fixed_steps=args.steps, warmup=args.warmup, ref_steps=args.ref_python_steps)
    val = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(n), "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import merging_gym_b200 as mg
    from merging_gym_b200.affinity import bind_to_gpu
    from merging_gym_b200.sharding import init_distributed, shard_range

    rank, local_rank, world = init_distributed()
    if world != args.gpus and rank == 0:
        print(f"# note: WORLD_SIZE={world} differs from --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    binding = bind_to_gpu(local_rank) if world > 1 else {"bound": False, "note": "single rank: not bound"}
    n = args.envs
    S, R = args.slots, args.shards
    # one distinct pre-generated action set per step of the captured graph
    unit = S * R // math.gcd(S, R)
    A = max(unit, (min(args.steps, args.graph_steps) // unit) * unit) if args.action_sets <= 0 else args.action_sets
    # R independent n-env shards (default 2^20) stepped round-robin: 4 x 54.5 MB of float64 state cannot stay in
    # the 126 MB L2 between two launches on the same shard, so every launch reads its inputs from HBM.
    envs = [mg.MergeVecEnv(n, mode="pvp", device=dev, auto_reset=True, seed=0x5EED,
                           env_id_base=(rank * R + r) * n, out_slots=S, episode_info=False, track_stats=True)
            for r in range(R)]
    env = envs[0]
    laned_headline = args.headline == "laned" and args.lanes > 1
    lenvs = None
    if args.lanes > 1:
        # the product API for overlapping launches: each 2^20-env shard is L lanes of 2^20/L envs on L streams
        lenvs = [mg.MergeVecEnv(n, mode="pvp", device=dev, auto_reset=True, seed=0x5EED, env_id_base=(rank * R + r) * n,
                                out_slots=S, episode_info=False, track_stats=True, lanes=args.lanes) for r in range(R)]
    henvs = lenvs if laned_headline else envs
    # pre-generated, HBM-resident action sets (Philox over global env ids), cycled over steps
    acts1 = torch.empty(A, n, dtype=torch.uint8, device=dev)
    acts2 = torch.empty(A, n, dtype=torch.uint8, device=dev)
    for i in range(A):
        a1, a2 = envs[i % R].sample_actions(i)
        acts1[i].copy_(a1); acts2[i].copy_(a2)
    # de-synchronise episodes so the timed region sees the steady-state reset rate (~1/210 per step)
    for e in envs + (lenvs or []):
        e.rollout(args.mix_steps, step0=1000)
    torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    peak, peak_src = load_peaks()

    # the path's only collective: NCCL all-reduce of the int64 statistics, on a side stream.  Banked: the launching
    # stream carries no statistics kernel at all (the side stream drains the retired bank).
    reducer = mg.AsyncStatsReducer(henvs[0], banked=True) if world > 1 else None
    if reducer is not None:                               # NCCL communicator set-up happens here, untimed
        reducer.submit()
        reducer.latest()
        torch.cuda.synchronize()
        reducer.submissions = 0

    gate_hz = 1e3 * float(getattr(torch.cuda.get_device_properties(dev), "clock_rate", 1.9e6))   # kHz -> Hz
    last_host_ms = [0.0]

    def timed_region(shards, K, W, sample_clocks, n_streams=1, graph_steps=None, lanes=False, with_reducer=True):
        """W warm-up + K timed step launches round-robin over `shards`; returns (ms, G, eager, clocks).
        n_streams > 1: shard r is stepped on stream r % n_streams (forked from / joined to the current
        stream around every batch), so launches of independent shards may overlap.
        lanes: every shard is a laned env; a step issues one launch per lane, each on the lane's own stream."""
        nsh = len(shards)
        idx = [0]
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else None
        gsteps = graph_steps or args.graph_steps

        def do_steps(k):
            main = torch.cuda.current_stream()
            if streams and k > 0:
                for ln in streams:
                    ln.wait_stream(main)
            for _ in range(k):
                i = idx[0]
                sh = shards[i % nsh]
                if lanes:
                    for l, sl in enumerate(sh.lane_slices):
                        sh.step_lane_async(l, acts1[i % A][sl], acts2[i % A][sl])
                elif streams:
                    with torch.cuda.stream(streams[(i % nsh) % n_streams]):
                        sh.step_async(acts1[i % A], acts2[i % A])
                else:
                    sh.step_async(acts1[i % A], acts2[i % A])
                idx[0] = i + 1
            if streams and k > 0:
                for ln in streams:
                    main.wait_stream(ln)
            if lanes and k > 0:
                for sh in shards:
                    sh.join_lanes()

        def rewind():
            idx[0] = 0
            for e in shards:
                e._slot = 0
                e._lane_slot = [0] * e.lanes

        do_steps(W)                                       # warm-up, eager
        torch.cuda.synchronize()
        G, graphs = 0, None
        cyc = S * nsh // math.gcd(S, nsh)
        cyc = cyc * A // math.gcd(cyc, A)                 # slot ring, shard ring and action ring line up
        if not args.no_graph and K >= cyc:
            G = (min(K, max(gsteps, cyc)) // cyc) * cyc
        use_reducer = reducer is not None and sample_clocks and with_reducer and shards[0] is reducer.env
        if G > 0:
            rewind()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                do_steps(cyc)                             # warm-up on a side stream before capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

            def issue():
                rewind()
                do_steps(G)
            if use_reducer:
                graphs = reducer.capture_per_bank(issue)  # one graph per statistics bank
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    issue()
                graphs = [g]
            graphs[0 if not use_reducer else env._stats_active].replay()   # one untimed replay
            torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        # Gate: a spin kernel keeps the stream busy for ~args.gate_ms while the host queues the whole region behind
        # it, so the events time the DEVICE executing K back-to-back launches and not how fast this host thread
        # (one of N ranks sharing the box's cores) issues a graph launch / an NCCL call.  Device time only either way.
        if args.gate_ms > 0:
            torch.cuda._sleep(int(args.gate_ms * 1e-3 * gate_hz))
        h0 = time.perf_counter()
        e0.record()
        done_steps = 0
        if graphs is not None:
            for _ in range(K // G):
                if use_reducer:
                    reducer.submit()                      # headline region only: the async NCCL statistics reduce of
                    reducer.replay(graphs)                # everything so far runs on its side stream UNDER these launches
                else:
                    graphs[0].replay()
                done_steps += G
        do_steps(K - done_steps)
        e1.record()
        host_ms = 1e3 * (time.perf_counter() - h0)
        if sampler:
            sampler.sample_now()                          # the queue is still draining: a sample under load
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        last_host_ms[0] = host_ms
        return float(t.item()), G, K - done_steps, clocks

    total_envs = n * world                               # envs advanced per step (one launch per GPU)

    def rate(ms, k, bytes_per=BYTES_PER_ENV_STEP):
        gbs = n * k * bytes_per / (ms * 1e-3) / 1e9
        return {"value": total_envs * k / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / k, "steps": k,
                "algorithmic_gbs_per_gpu": gbs, "frac_of_peak": gbs / peak}

    extra_errors = {}

    class guard:
        """`with guard("name"):` around an extra measurement: a failure must never cost the line its contract keys; it
        is reported under `extra_errors` and the extra stays null."""

        def __init__(self, name):
            self.name = name

        def __enter__(self):
            return self

        def __exit__(self, et, ev, tb):
            if et is None or not issubclass(et, Exception):
                return False
            extra_errors[self.name] = repr(ev)
            print(f"# bench.py: extra measurement {self.name!r} failed: {ev!r}", file=sys.stderr)
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa
                pass
            return True

    ms, G, eager, clocks = timed_region(henvs, K, W, True, lanes=laned_headline)
    host_issue_ms = last_host_ms[0]
    value = total_envs * K / (ms * 1e-3)
    per_gpu_gbs = n * K * BYTES_PER_ENV_STEP / (ms * 1e-3) / 1e9
    n_reductions = reducer.submissions if reducer else 0
    # the other launch scheme over the same K steps, reported beside the headline
    other = None
    with guard("serialized_at_headline_steps"):
        if lenvs is not None:
            ms_o, G_o, _, _ = timed_region(envs if laned_headline else lenvs, K, W, False, lanes=not laned_headline)
            other = dict(rate(ms_o, K), launch=f"CUDA graph of {G_o} steps" if G_o else "eager")

    # ---- sustained: the same region at 2000 steps (graph of 200) whatever --steps says ------------------------
    sustained = None
    with guard("sustained"):
        if args.sustained_steps > 0:
            if K >= args.sustained_steps:
                sustained = dict(rate(ms, K), note="the headline region itself")
            else:
                # the driver's --steps region lasts half a millisecond, too short for NVML to see: the clocks are sampled
                # over this longer region as well (`clocks.sustained`)
                ms_s, G_s, _, clocks_s = timed_region(henvs, args.sustained_steps, W, True, lanes=laned_headline, with_reducer=False)
                if clocks is not None and clocks_s is not None:
                    clocks["sustained"] = clocks_s
                    clocks["reasons"] = sorted(set(clocks["reasons"]) | set(clocks_s["reasons"]))
                sustained = dict(rate(ms_s, args.sustained_steps),
                                 note=f"same shards and launches as `value`, {args.sustained_steps} steps as a CUDA graph of {G_s} "
                                      "replayed; shows what the short driver-run region (--steps) cannot amortise")
    # same kernel, ONE shard stepped in place: its 54.5 MB state is partly L2-resident between steps
    ms_warm, Kw = None, max(K, 200)
    with guard("l2_warm"):
        ms_warm, _, _, _ = timed_region(envs[:1], Kw, W, False)

    # ---- extra: the same shards on two CUDA streams, so the ramp-up of one shard's launch overlaps the
    #      drain of another's (a forked CUDA graph); whole-GPU throughput, not a per-launch figure -----------
    overlapped = None
    Ko = max(K, 400)
    with guard("overlapped_streams"):
        if args.overlap_streams > 1 and R % args.overlap_streams == 0:
            ms_ov, _, _, _ = timed_region(envs, Ko, W, False, n_streams=args.overlap_streams)
            overlapped = dict(rate(ms_ov, Ko), streams=args.overlap_streams,
                              note=f"the same {R} shards, shard r on stream r % {args.overlap_streams}: launches of "
                                   "independent shards overlap, which hides the per-launch ramp-up/drain that separates "
                                   "the serialised 2^20-env launch from the copy peak; launches are concurrent, so this "
                                   "is aggregate throughput and is not used for value/roofline")

    # ---- extra: the other launch scheme at the longer step count -------------------------------------------------
    laned = serialized = None
    with guard("other_scheme"):
        if lenvs is not None:
            ms_l, _, _, _ = timed_region(envs if laned_headline else lenvs, Ko, W, False, lanes=not laned_headline)
            o = dict(rate(ms_l, Ko), at_headline_steps=other)
            if laned_headline:
                serialized = dict(o, note="one mg_step launch per 2^20-env step, all on ONE stream (round 1's headline): each launch "
                                          "pays its own ramp-up and drain")
            else:
                laned = dict(o, lanes=args.lanes, launches_per_step=args.lanes,
                             note=f"MergeVecEnv(lanes={args.lanes}): each step issued as {args.lanes} launches on the lanes' streams")
    # ---- extra: track_returns=False (no float64 return accumulators): 124 B/env-step ------------------------------
    lean = None
    with guard("lean_no_returns"):
        if args.lean:
            nenvs = [mg.MergeVecEnv(n, mode="pvp", device=dev, auto_reset=True, seed=0x5EED, env_id_base=(rank * R + r) * n,
                                    out_slots=S, episode_info=False, track_stats=True, track_returns=False) for r in range(R)]
            for e in nenvs:
                e.rollout(args.mix_steps, step0=1000)
            ms_n, _, _, _ = timed_region(nenvs, Ko, W, False)
            lean = dict(rate(ms_n, Ko, BYTES_PER_ENV_STEP_LEAN), bytes_per_env_step=BYTES_PER_ENV_STEP_LEAN,
                        note="MergeVecEnv(track_returns=False): r1_accumulate / r2_accumulate (merging_env.py:191-192) are not "
                             "kept, 124 instead of 156 B of HBM traffic per env-step; frac_of_peak uses 124 B")
            del nenvs

    # ---- cross-check of the L2 methodology: ONE shard stepped in place with L2 flushed (a 512 MB
    #      buffer overwritten) before every timed launch, each launch bracketed by its own events ------
    flushed = None
    with guard("l2_flushed"):
        if args.flush_steps > 0:
            fl = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
            fr = torch.zeros(64 << 20, dtype=torch.int64, device=dev)          # 512 MB, only ever read
            sink = torch.zeros((), dtype=torch.int64, device=dev)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(args.flush_steps)]
            for i, (a, b) in enumerate(evs):
                fl.fill_(i & 0xFF)                            # evicts the shard's state and outputs from L2 ...
                sink += fr.sum()                              # ... and a read pass leaves L2 full of CLEAN lines, so
                a.record()                                    # the timed launch does not pay for the flush's write-back
                env.step_async(acts1[i % A], acts2[i % A])
                b.record()
            torch.cuda.synchronize()
            fms = sorted(a.elapsed_time(b) for a, b in evs)
            fmed = fms[len(fms) // 2]
            flushed = {"ms_per_step_median": fmed, "ms_per_step_min": fms[0], "steps": len(fms),
                       "value": n / (fmed * 1e-3), "unit": UNIT + " per GPU",
                       "frac_of_peak": n * BYTES_PER_ENV_STEP / (fmed * 1e-3) / 1e9 / peak,
                       "note": "single shard in place; before every launch 512 MB are written and then 512 MB read to "
                               "flush the 126 MB L2 (leaving clean lines); one event pair per eager launch, so the "
                               "figure includes launch/event overhead and has no PDL or graph overlap"}
            del fl, fr

    # ---- extra: K fused steps per launch with in-kernel Philox actions (mg_rollout), all outputs on ----
    RK = args.rollout_k
    rollout = None
    with guard("rollout_fused"):
        if RK > 0:
            ro = torch.empty(RK, n, 10, device=dev); rr = torch.empty(RK, n, 2, device=dev)
            rd = torch.empty(RK, n, dtype=torch.uint8, device=dev); ri = torch.empty(RK, n, dtype=torch.uint8, device=dev)
            for _ in range(2):
                env.rollout(RK, obs=ro, rew=rr, done=rd, info=ri, refresh_obs=False)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            r0.record()
            for _ in range(8):
                env.rollout(RK, obs=ro, rew=rr, done=rd, info=ri, refresh_obs=False)
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / (8 * RK)
            t = torch.tensor([rms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rms = float(t.item())
            rb = 50 + 104.0 / RK
            rollout = {"value": total_envs / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms, "k_steps_per_launch": RK,
                       "bytes_per_env_step": rb, "achieved_gbs_per_gpu": n * rb / (rms * 1e-3) / 1e9,
                       "note": "mg_rollout: state stays in registers for K steps, actions from in-kernel Philox, "
                               "obs/rew/done/info written time-major every step (the last row is the current observation); "
                               "instruction-issue bound, not HBM bound"}
            del ro, rr, rd, ri

    # ---- end-to-end through the host-buffer API ------------------------------------------------------------------
    E = max(3, min(K, args.e2e_steps))
    out_bytes = n * (40 + 8 + 1 + 1)
    slots = [env.host_action_buffers(0), env.host_action_buffers(1)]   # pinned host memory: the steps' inputs live here
    for k_, (h1, h2) in enumerate(slots):
        h1[:] = acts1[k_ % A].cpu().numpy(); h2[:] = acts2[k_ % A].cpu().numpy()

    def e2e_loop(fn, steps):
        fn(2)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(steps)
        s = time.perf_counter() - t0
        t = torch.tensor([s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync_steps(k):
        for _ in range(k):
            env.step_host(slots[0][0], slots[0][1])          # synchronises inside

    def pipelined(fields):
        def run(k):
            for i in range(k):
                if i >= 2:
                    env.step_host_wait()                     # the host now owns step i-2's outputs
                h1, h2 = env.host_action_buffers()           # the pinned slot this step will use (pre-filled above)
                env.step_host_async(h1, h2, fields=fields)
            for _ in range(min(k, 2)):
                env.step_host_wait()
        return run

    s_sync = e2e_loop(sync_steps, E)
    s_pipe = e2e_loop(pipelined(None), E)
    s_small = e2e_loop(pipelined(("rew", "done", "info")), E)

    # what the box moves device->host when every rank copies at once: one plain cudaMemcpyAsync per copy, back to back
    hbuf = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(out_bytes, dtype=torch.uint8, device=dev)

    def plain_copies(k):
        for _ in range(k):
            hbuf.copy_(dbuf, non_blocking=True)
        torch.cuda.synchronize()
    s_ceil = e2e_loop(plain_copies, E)
    ceiling_gbs = world * out_bytes * E / s_ceil / 1e9
    del hbuf, dbuf
    e2e = {"value": total_envs * E / s_pipe, "unit": UNIT, "steps": E,
           "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": out_bytes,
           "api": "MergeVecEnv.step_host_async / step_host_wait -> mg_step_host_async (two pinned slots: the step kernel "
                  "reads this step's actions straight from pinned host memory across PCIe, obs|rew|done|info are copied "
                  "into pinned host memory by one cudaMemcpyAsync on a copy stream; the host waits for step t-2 before it "
                  "queues step t, all steps have landed when the clock stops)",
           "pcie_gbs_all_gpus": world * (2 * n + out_bytes) * E / s_pipe / 1e9,
           "ceiling_gbs": ceiling_gbs,
           "frac_of_ceiling": world * out_bytes * E / s_pipe / 1e9 / ceiling_gbs,
           "ceiling_note": f"measured in this run: every rank copies {out_bytes} B device->pinned host {E}x back to back with "
                           "plain cudaMemcpyAsync, all ranks at once; aggregate GB/s over the slowest rank "
                           "(profiles/d2h_ceiling.py measures the same stand-alone for N = 1/2/4/8)",
           "sync_value": total_envs * E / s_sync,
           "sync_api": "MergeVecEnv.step_host -> mg_step_host (same copies, host synchronises every step; round-1 figure)",
           "rew_done_info_value": total_envs * E / s_small,
           "rew_done_info_d2h_bytes_per_step": n * 10,
           "rew_done_info_note": "fields=('rew','done','info'): for callers whose policy reads the device-resident "
                                 "observation; NOT the headline (the reference's step returns obs too)",
           "cpu_binding": binding}

    # ---- episode statistics: the one collective on this path (tiny int64 all-reduce over NCCL) --
    stats = env.stats(reduce=world > 1)

    # ---- BASELINE configs[3]: 2^23 envs over all ranks, statistics all-reduced every 64 steps -------------------
    strong = None
    with guard("strong_8m"):
        if args.strong_envs > 0:
            base, cnt = shard_range(args.strong_envs, rank, world)
            senv = mg.MergeVecEnv(cnt, mode="pvp", device=dev, auto_reset=True, seed=0x5EED, env_id_base=base, out_slots=2,
                                  episode_info=False, track_stats=True)
            SA = 4
            sa1 = torch.empty(SA, cnt, dtype=torch.uint8, device=dev); sa2 = torch.empty(SA, cnt, dtype=torch.uint8, device=dev)
            for i in range(SA):
                a1, a2 = senv.sample_actions(i)
                sa1[i].copy_(a1); sa2[i].copy_(a2)
            sred = mg.AsyncStatsReducer(senv, banked=True)
            ctr = [0]

            def s_issue():
                senv._slot = 0
                for i in range(STRONG_REDUCE_EVERY):
                    senv.step_async(sa1[i % SA], sa2[i % SA])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(4):
                    senv.step_async(sa1[i % SA], sa2[i % SA])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            sgraphs = sred.capture_per_bank(s_issue)
            senv.reset()                                          # the digest run starts from reset, statistics zeroed
            senv.stats(reset=True)
            senv._slot = 0
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(STRONG_STEPS // STRONG_REDUCE_EVERY):
                sred.replay(sgraphs)
                sred.submit()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sms = float(t.item())
            totals = [int(v) for v in sred.latest().cpu().tolist()]
            strong = {"value": args.strong_envs * STRONG_STEPS / (sms * 1e-3), "unit": UNIT, "scaling": "strong",
                      "total_envs": args.strong_envs, "envs_per_gpu": cnt, "steps": STRONG_STEPS, "ms_per_step": sms / STRONG_STEPS,
                      "reductions": sred.submissions, "reduce_every": STRONG_REDUCE_EVERY,
                      "algorithmic_gbs_per_gpu": cnt * STRONG_STEPS * BYTES_PER_ENV_STEP / (sms * 1e-3) / 1e9,
                      "stats_totals": totals,
                      "stats_digest": hashlib.sha256(json.dumps(totals).encode()).hexdigest()[:16],
                      "note": f"BASELINE configs[3]: {args.strong_envs} envs in total sharded contiguously over the ranks, "
                              f"{STRONG_STEPS} steps from reset as CUDA graphs of {STRONG_REDUCE_EVERY} mg_step launches, the int64 "
                              "statistics all-reduced (NCCL, side stream, banked: no statistics kernel on the launching stream) "
                              "after every graph; stats_digest = sha256 of the reduced totals after the last step — it must be "
                              "the same at every world size (actions are Philox over GLOBAL env ids)"}
            del senv, sa1, sa2

    # ---- BASELINE configs[4]: DQN policy forward in the loop, 2^18 envs per GPU --------------------------------------
    policy = None
    with guard("policy_in_loop"):
        if args.policy_envs > 0 and rank == 0:
            import bench_policy
            policy = {"note": "BASELINE configs[4]: pve, policy forward + arg-max -> mg_step on the device, random starts; "
                              "graph-timed (bench_policy.measure); rank 0 only"}
            for be in ("fused", "tf32x3", "f16x3"):
                try:
                    policy[be] = bench_policy.measure(args.policy_envs, "dqn", be, device=dev)
                except Exception as e:  # noqa
                    policy[be] = {"error": repr(e)}
            for be in ("tf32x3", "f16x3"):   # the launch-bound end of the same loop: 4096 envs, where one launch per step (mg_policy_step) pays
                try:
                    small = bench_policy.measure(4096, "dqn", be, device=dev, k=32, replays=8)
                    policy[be + "_4096_envs"] = {k_: small[k_] for k_ in ("value", "ms_per_step", "fused_step", "envs")}
                except Exception as e:  # noqa
                    policy[be + "_4096_envs"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    LPS = henvs[0].lanes
    tr = load_traffic() or {}
    traffic_per_step = None
    if tr.get("dram_bytes_per_launch") and tr.get("envs_per_launch"):     # ncu bytes of one launch, scaled to one 2^20-env step
        traffic_per_step = tr["dram_bytes_per_launch"] * (n / tr["envs_per_launch"])
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_baseline_object(args.cpu_seconds, ref_steps=args.ref_python_steps)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n),
            "measurement": {"total_envs": total_envs, "envs_per_step": n, "launches_per_step": LPS,
                            "envs_per_launch": n // LPS, "shards_per_gpu": R,
                            "scheme": (f"MergeVecEnv(lanes={LPS}): a step of a 2^20-env shard = {LPS} mg_step launches of "
                                       f"{n // LPS} envs, lane l on its own CUDA stream and ordered only behind lane l's previous "
                                       "step, so one lane's launch ramps up while the other's drains (bit-identical to one "
                                       "launch: tests/test_gpu_round2.py); `serialized` = one launch per step on one stream"
                                       if laned_headline else "one mg_step launch per step on one stream"),
                            "launch": (f"CUDA graph of {G} steps ({G * LPS} mg_step launches) replayed {K // G}x + {eager} eager steps"
                                       if G else "eager ctypes launches"),
                            "gate": (f"a {args.gate_ms} ms spin kernel precedes the start event, so the host has queued the region "
                                     "before the device starts it" if args.gate_ms > 0 else "none"),
                            "host_issue_ms": host_issue_ms,
                            "state_mb": R * n * 52 / 1e6, "action_sets": A, "ring_slots": S,
                            "parallelism": f"env-sharded x{world}, no data-path collective; NCCL all-reduce of 16 int64 "
                                           f"stats on a side stream before every graph replay inside the timed region, running "
                                           f"under the step launches ({n_reductions} reductions; banked: no statistics kernel "
                                           "on the timed streams)"},
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s",
                         "frac": per_gpu_gbs / peak, "traffic": traffic_per_step,
                         "kernel": "mg::merge_step_kernel<2, uint8_t, true, false, true>",
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "peak_source": peak_src,
                         "per": ("GPU; achieved = 156 B x envs stepped in the timed region / its duration (CUDA events); with "
                                 f"{LPS} lanes the launches overlap, so this is the rate of the kernel over the region, not of "
                                 "one isolated launch (that one is `serialized`); traffic = ncu DRAM bytes of one 2^20-env launch"
                                 if laned_headline else
                                 "GPU; achieved = 156 B x envs_per_gpu / (timed ms / steps)")},
            "serialized": serialized,
            "sustained": sustained,
            "l2_warm": None if ms_warm is None else dict(
                rate(ms_warm, Kw), note="one 2^20-env shard stepped in place (the literal 1M-envs/GPU deployment): its "
                                        "54.5 MB state is partly L2-resident between steps, so it runs faster than the "
                                        "HBM roofline allows; not used for value/roofline"),
            "l2_flushed": flushed,
            "overlapped_streams": overlapped,
            "laned": laned,
            "lean_no_returns": lean,
            "rollout_fused": rollout,
            "policy_in_loop": policy,
            "strong_8m": strong,
            "extra_errors": extra_errors or None,
            "e2e": e2e, "gpu_launches": K * LPS, "clocks": clocks,
            "episode_stats": {k: stats[k] for k in ("episodes", "collision_rate", "merge_success_rate",
                                                    "mean_length", "mean_return1", "mean_return2")}}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def guard_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line there.  Everything but `emit()`
    goes to stderr from here on."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--slots", type=int, default=2, help="output ring slots per shard")
    ap.add_argument("--shards", type=int, default=4, help="independent 2^20-env shards per GPU, stepped round-robin")
    ap.add_argument("--action-sets", type=int, default=0, help="0 = one per graph step")
    ap.add_argument("--graph-steps", type=int, default=200)
    ap.add_argument("--mix-steps", type=int, default=400)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--sustained-steps", type=int, default=2000, help="extra region at this many steps (0 = skip)")
    ap.add_argument("--flush-steps", type=int, default=100, help="launches of the L2-flushed cross-check (0 = skip)")
    ap.add_argument("--rollout-k", type=int, default=32, help="steps per mg_rollout launch for the extra rollout_fused figure (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--overlap-streams", type=int, default=2,
                    help="extra measurement: the shards on this many CUDA streams (0/1 = skip)")
    ap.add_argument("--lanes", type=int, default=2, help="extra measurement: MergeVecEnv(lanes=L) (0/1 = skip)")
    ap.add_argument("--lean", type=int, default=1, help="extra measurement: track_returns=False (0 = skip)")
    ap.add_argument("--strong-envs", type=int, default=STRONG_TOTAL_ENVS, help="configs[3] total envs (0 = skip)")
    ap.add_argument("--policy-envs", type=int, default=1 << 18, help="configs[4] envs per GPU (0 = skip)")
    ap.add_argument("--ref-python-steps", type=int, default=3000,
                    help="steps per process when timing the reference's own Python env (0 = skip)")
    ap.add_argument("--gate-ms", type=float, default=4.0,
                    help="spin-kernel gate in front of every timed region so that the host has queued the region before "
                         "the device starts it (0 = none)")
    ap.add_argument("--headline", default="laned", choices=["laned", "serialized"],
                    help="what `value` measures: MergeVecEnv(lanes=L) (each step = L launches on L streams) or one "
                         "launch per step on one stream; the other one is reported beside it")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
