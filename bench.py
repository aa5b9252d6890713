#!/usr/bin/env python
"""bench.py — env-steps/s of the fused merging-gym step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is ONE `mg_step` launch advancing every env of the rank's shard by one dT with external
uint8 actions (pvp, auto-reset): BASELINE.json configs[2], 2^20 envs per launch per GPU.
Actions are pre-generated on the device (Philox, global env ids), so inputs are HBM-resident.
L2 rule: one shard's inputs (56.6 MB) would fit the 126 MB L2, so the timed region steps
`--shards` (4) independent 2^20-env shards round-robin — inputs larger than L2, no flush — and
every launch reads its state from HBM.  The in-place single-shard figure is reported separately as
`l2_warm`.  Timed with CUDA events on the launching (current torch) stream, barrier +
synchronize on both sides, max over ranks.

Extra objects in the JSON line: `roofline` (HBM, algorithmic bytes = 156 B/env-step, see
DESIGN.md), `cpu_baseline` (the C port of the oracle on this box's host cores, rank 0, N=1),
`e2e` (the same step through `mg_step_host`: pinned host actions -> H2D -> step -> D2H of
obs/rewards/done/info, synchronised), `clocks`, `gpu_launches`.

`--impl reference` times the reference's CPU implementation of the path.  The reference is
pure Python and cannot travel to the GPU box (and needs gym/pygame/shapely/qpsolvers, absent
from the image), so this arm runs the oracle port: the plain-C restatement with all host
threads (`kind: "port"`).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
BYTES_PER_ENV_STEP = 156          # read 54 B + write 102 B (pvp, u8 actions, auto-reset)
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


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
    import numpy as np
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


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    threads = min(host_threads(), 64)
    n = ENVS_PER_GPU
    val, steps, dt = cpu_port_throughput(n, 0, threads, warmup_steps=args.warmup, fixed_steps=args.steps)
    sample = f"{n} envs x {steps} steps (pvp, auto-reset, pre-generated uint8 actions), float64, {threads} OpenMP threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "pvp, 2^20 envs, auto-reset, uniform-random uint8 actions (BASELINE.json configs[2]); "
                                   "CPU sample is one 2^20-env shard regardless of --gpus",
                       "envs": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "reference is pure Python needing gym/pygame/shapely/qpsolvers (absent) and "
                                     "cannot travel to the GPU box; this is the plain-C oracle port, which is "
                                     "far faster than the reference's ~3e3 steps/s Python loop"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import merging_gym_b200 as mg
    from merging_gym_b200.sharding import init_distributed

    rank, local_rank, world = init_distributed()
    if world != args.gpus and rank == 0:
        print(f"# note: WORLD_SIZE={world} differs from --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n = args.envs
    S, R = args.slots, args.shards
    import math
    # one distinct pre-generated action set per step of the captured graph
    unit = S * R // math.gcd(S, R)
    A = max(unit, (min(args.steps, args.graph_steps) // unit) * unit) if args.action_sets <= 0 else args.action_sets
    # R independent n-env shards (default 2^20) stepped round-robin: 4 x 54.5 MB of float64 state cannot stay in
    # the 126 MB L2 between two launches on the same shard, so every launch reads its inputs from HBM.
    envs = [mg.MergeVecEnv(n, mode="pvp", device=dev, auto_reset=True, seed=0x5EED,
                           env_id_base=(rank * R + r) * n, out_slots=S, episode_info=False, track_stats=True)
            for r in range(R)]
    env = envs[0]
    # pre-generated, HBM-resident action sets (Philox over global env ids), cycled over steps
    acts1 = torch.empty(A, n, dtype=torch.uint8, device=dev)
    acts2 = torch.empty(A, n, dtype=torch.uint8, device=dev)
    for i in range(A):
        a1, a2 = envs[i % R].sample_actions(i)
        acts1[i].copy_(a1); acts2[i].copy_(a2)
    # de-synchronise episodes so the timed region sees the steady-state reset rate (~1/210 per step)
    for e in envs:
        e.rollout(args.mix_steps, step0=1000)
    torch.cuda.synchronize()

    K, W = args.steps, args.warmup

    reducer = mg.AsyncStatsReducer(env) if world > 1 else None
    if reducer is not None:                               # NCCL communicator set-up happens here, untimed
        reducer.submit()
        reducer.latest()
        torch.cuda.synchronize()
        reducer.submissions = 0

    def timed_region(shards, K, W, sample_clocks, n_streams=1):
        """W warm-up + K timed mg_step launches round-robin over `shards`; returns (ms, G, eager, clocks).
        n_streams > 1: shard r is stepped on stream r % n_streams (forked from / joined to the current
        stream around every batch), so launches of independent shards may overlap."""
        nsh = len(shards)
        idx = [0]
        lanes = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else None

        def do_steps(k):
            main = torch.cuda.current_stream()
            if lanes and k > 0:
                for ln in lanes:
                    ln.wait_stream(main)
            for _ in range(k):
                i = idx[0]
                if lanes:
                    with torch.cuda.stream(lanes[(i % nsh) % n_streams]):
                        shards[i % nsh].step_async(acts1[i % A], acts2[i % A])
                else:
                    shards[i % nsh].step_async(acts1[i % A], acts2[i % A])
                idx[0] = i + 1
            if lanes and k > 0:
                for ln in lanes:
                    main.wait_stream(ln)

        do_steps(W)                                       # warm-up, eager
        torch.cuda.synchronize()
        G, graph = 0, None
        cyc = S * nsh // math.gcd(S, nsh)
        cyc = cyc * A // math.gcd(cyc, A)                 # slot ring, shard ring and action ring line up
        if not args.no_graph and K >= cyc:
            G = (min(K, max(args.graph_steps, cyc)) // cyc) * cyc
        if G > 0:
            def rewind():
                idx[0] = 0
                for e in shards:
                    e._slot = 0
            rewind()
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                do_steps(cyc)                             # warm-up on a side stream before capture
            torch.cuda.current_stream().wait_stream(side)
            rewind()
            with torch.cuda.graph(graph):
                do_steps(G)
            graph.replay()                                # one untimed replay
            torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0.record()
        done_steps = 0
        if graph is not None:
            for _ in range(K // G):
                graph.replay()
                done_steps += G
                if reducer is not None and sample_clocks:   # headline region only: async NCCL stats reduce
                    reducer.submit()
        do_steps(K - done_steps)
        e1.record()
        if sampler:
            sampler.sample_now()                          # the queue is still draining: a sample under load
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), G, K - done_steps, clocks

    ms, G, eager, clocks = timed_region(envs, K, W, True)
    total_envs = n * world                               # envs advanced per step (one launch per GPU)
    value = total_envs * K / (ms * 1e-3)
    per_gpu_gbs = n * K * BYTES_PER_ENV_STEP / (ms * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    # same kernel, ONE shard stepped in place: its 54.5 MB state is partly L2-resident between steps
    ms_warm, _, _, _ = timed_region(envs[:1], K, W, False)

    # ---- extra: the same shards on two CUDA streams, so the ramp-up of one shard's launch overlaps the
    #      drain of another's (a forked CUDA graph); whole-GPU throughput, not a per-launch figure -----------
    overlapped = None
    if args.overlap_streams > 1 and R % args.overlap_streams == 0:
        ms_ov, _, _, _ = timed_region(envs, K, W, False, n_streams=args.overlap_streams)
        ov_gbs = n * K * BYTES_PER_ENV_STEP / (ms_ov * 1e-3) / 1e9
        overlapped = {"value": total_envs * K / (ms_ov * 1e-3), "unit": UNIT, "streams": args.overlap_streams,
                      "ms_per_step_effective": ms_ov / K, "algorithmic_gbs_per_gpu": ov_gbs,
                      "frac_of_peak": ov_gbs / peak,
                      "note": f"the same {R} shards, shard r on stream r % {args.overlap_streams}: launches of "
                              "independent shards overlap, which hides the per-launch ramp-up/drain that separates "
                              "the serialised 2^20-env launch from the copy peak; launches are concurrent, so this "
                              "is aggregate throughput and is not used for value/roofline"}

    # ---- cross-check of the L2 methodology: ONE shard stepped in place with L2 flushed (a 512 MB
    #      buffer overwritten) before every timed launch, each launch bracketed by its own events ------
    flushed = None
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
    if RK > 0:
        ro = torch.empty(RK, n, 10, device=dev); rr = torch.empty(RK, n, 2, device=dev)
        rd = torch.empty(RK, n, dtype=torch.uint8, device=dev); ri = torch.empty(RK, n, dtype=torch.uint8, device=dev)
        for _ in range(2):
            env.rollout(RK, obs=ro, rew=rr, done=rd, info=ri)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record()
        for _ in range(8):
            env.rollout(RK, obs=ro, rew=rr, done=rd, info=ri)
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
                           "obs/rew/done/info written time-major every step; instruction-issue bound, not HBM bound"}
        del ro, rr, rd, ri

    # ---- end-to-end through the host-buffer C-ABI entry (mg_step_host) -------------------------
    import numpy as np
    E = max(3, min(K, args.e2e_steps))
    h1, h2 = env.host_action_buffers()                    # pinned host memory: this step's inputs live here
    h1[:] = acts1[0].cpu().numpy(); h2[:] = acts2[0].cpu().numpy()
    for _ in range(2):
        env.step_host(h1, h2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(E):
        env.step_host(h1, h2)                             # synchronises inside
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": total_envs * E / e2e_s, "unit": UNIT, "steps": E,
           "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": n * (40 + 8 + 1 + 1),
           "api": "MergeVecEnv.step_host -> mg_step_host (actions in pinned host memory, read across PCIe by the step "
                  "kernel itself; outputs copied into pinned host memory; stream synchronised every step)",
           "pcie_gbs": total_envs / world * 52 * E / e2e_s / 1e9,
           "bound": "PCIe: 52 B/env-step cross the bus (2 up, 50 down); a plain 52 MB device->host copy reaches "
                    "~56 GB/s on this pool (profiles/README.md)"}

    # ---- episode statistics: the one collective on this path (tiny int64 all-reduce over NCCL) --
    stats = env.stats(reduce=world > 1)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = min(host_threads(), 64)
        v, steps, dt = cpu_port_throughput(ENVS_PER_GPU, args.cpu_seconds, threads)
        v1, steps1, dt1 = cpu_port_throughput(1 << 16, min(3.0, args.cpu_seconds), 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{ENVS_PER_GPU} envs x {steps} steps in {dt:.1f} s (pvp, auto-reset, float64, plain-C oracle port, OpenMP)",
               "single_thread_value": v1,
               "python_scalar_port_value": python_scalar_port_throughput(2.0),
               "note": "python_scalar_port_value is the per-env Python loop the reference itself runs "
                       "(its QP/pygame/shapely calls replaced by closed forms, so it over-estimates the reference)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "pvp, 2^20 envs per GPU, auto-reset, uniform-random uint8 actions "
                                   "pre-generated on device (BASELINE.json configs[2])",
                       "envs_per_gpu": n, "total_envs": total_envs, "mode": "pvp", "auto_reset": True,
                       "envs_per_launch": n, "shards_per_gpu": R,
                       "launch": (f"CUDA graph of {G} mg_step launches replayed {K // G}x + {eager} eager"
                                  if G else "eager ctypes launches"),
                       "l2": f"inputs larger than L2, no flush: {R} independent {n}-env shards per GPU stepped "
                             f"round-robin ({R * n * 52 / 1e6:.0f} MB of float64 state + {A} action sets "
                             f"{A * n * 2 / 1e6:.0f} MB, outputs to {R}x{S} ring slots of {n * 50 / 1e6:.0f} MB); "
                             "each launch re-reads its shard's state from HBM",
                       "parallelism": f"env-sharded x{world}, no data-path collective; NCCL all-reduce of 16 int64 "
                                      f"stats on a side stream every {G or K} steps inside the timed region "
                                      f"({reducer.submissions if reducer else 0} reductions)"},
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s",
                         "frac": per_gpu_gbs / peak, "traffic": (load_traffic() or {}).get("dram_bytes_per_launch"),
                         "kernel": "mg::merge_step_kernel<2, uint8_t, true>",
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "peak_source": peak_src,
                         "per": "GPU; achieved = 156 B x envs_per_gpu / (timed ms / steps)"},
            "l2_warm": {"value": total_envs * K / (ms_warm * 1e-3), "unit": UNIT, "ms_per_step": ms_warm / K,
                        "note": "one 2^20-env shard stepped in place (the literal 1M-envs/GPU deployment): its "
                                "54.5 MB state is partly L2-resident between steps, so it runs faster than the "
                                "HBM roofline allows; not used for value/roofline"},
            "l2_flushed": flushed,
            "overlapped_streams": overlapped,
            "rollout_fused": rollout,
            "e2e": e2e, "gpu_launches": K, "clocks": clocks,
            "episode_stats": {k: stats[k] for k in ("episodes", "collision_rate", "merge_success_rate",
                                                    "mean_length", "mean_return1", "mean_return2")}}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--flush-steps", type=int, default=100, help="launches of the L2-flushed cross-check (0 = skip)")
    ap.add_argument("--rollout-k", type=int, default=32, help="steps per mg_rollout launch for the extra rollout_fused figure (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--overlap-streams", type=int, default=2,
                    help="extra measurement: the shards on this many CUDA streams (0/1 = skip)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
