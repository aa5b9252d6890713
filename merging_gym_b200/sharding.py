"""Multi-GPU plumbing: one process per GPU, one independent shard of envs per rank.

The path has no exchange step — envs never interact — so ranks share nothing on the data
path.  The only collective is the optional SUM all-reduce of the 16-element int64 episode
statistics vector (NCCL over NVLink on GPUs, gloo in the CPU tests).  Because the statistics are
integers (returns in 2^24 fixed point) the reduced result is identical for every world size.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

from ._native import STAT_NAMES


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of global env ids [0, total) -> (first id, count) of `rank`.

    The first `total % world` ranks get one extra env; ids are global so that Philox action
    streams (key = seed, counter = (global env id, step)) do not depend on the sharding.
    """
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    q, r = divmod(int(total_envs), int(world_size))
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_distributed(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from env:// if WORLD_SIZE > 1.  NCCL when CUDA is present."""
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def all_reduce_stats(stats: torch.Tensor, async_op: bool = False):
    """SUM the int64 statistics vector over all ranks (no-op without an initialised group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    stats = stats.contiguous()
    work = dist.all_reduce(stats, op=dist.ReduceOp.SUM, async_op=async_op)
    return (stats, work) if async_op else stats


class AsyncStatsReducer:
    """Periodic, off-critical-path reduction of the episode statistics over all ranks — the path's only collective.

    `submit()` takes a snapshot of this rank's int64[16] totals that is consistent with everything launched on the
    current stream so far, and all-reduces it on a side stream (NCCL on GPUs); the env's launches are never blocked.
    `latest()` waits for the most recent submission and returns the reduced totals.

    banked=False: the snapshot is one small row-sum kernel on the launching stream.
    banked=True:  the launching stream carries NO statistics work at all: `submit()` retires the env's active
        statistics bank (later launches add into the other one), records an event, and the side stream drains the
        retired bank — row sum into the env's running totals, zero — before the all-reduce.  Launches pick the bank
        up when they are issued, so a caller that replays CUDA graphs must capture one graph per bank and replay the
        one that matches `env._stats_active` (bench.py does); eager launches need nothing.
    """

    def __init__(self, env, banked: bool = False):
        self.env = env
        self.banked = bool(banked)
        self.side = torch.cuda.Stream(device=env.device)
        self._work = None
        self._buf = None
        self._drained = {}                                  # bank id -> event after which it is empty again
        self.submissions = 0

    def submit(self) -> None:
        import torch.distributed as dist
        env = self.env
        cur = torch.cuda.current_stream(env.device)
        if self.banked:
            retired = env._retire_stats_bank()
            ev = self._drained.get(id(env.stats_buf))       # the bank that becomes active was drained by an earlier
            if ev is not None:                              # submit: order its reuse behind that (long finished)
                cur.wait_event(ev)
            snap = None
        else:
            retired = None
            snap = env.stats_tensor()                       # tiny sum kernel on the current stream
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            if self.banked:
                env._stats_total += retired.sum(dim=0)
                retired.zero_()
                done = torch.cuda.Event()
                done.record(self.side)
                self._drained[id(retired)] = done
                env._stats_side_event = done
                snap = env._stats_total.clone()
            else:
                snap.record_stream(self.side)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self._work = dist.all_reduce(snap, op=dist.ReduceOp.SUM, async_op=True)
            self._buf = snap
        self.submissions += 1

    def capture_per_bank(self, issue):
        """For CUDA-graph callers of a banked reducer: captures `issue()` (a callable that launches this env's steps)
        once per statistics bank and returns [graph_bank0, graph_bank1]; `replay(graphs)` replays the one whose
        launches add into the currently active bank.  `issue` must put the env's host-side cursors (output slot,
        action index) back to the same start itself."""
        env = self.env
        if len(env._stats_banks) == 1:
            env._retire_stats_bank()
            env._retire_stats_bank()
        keep, graphs = env._stats_active, []
        for b in (0, 1):
            env._stats_active = b
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                issue()
            graphs.append(g)
        env._stats_active = keep
        return graphs

    def replay(self, graphs) -> None:
        graphs[self.env._stats_active if self.banked else 0].replay()

    def latest(self) -> torch.Tensor:
        if self._buf is None:
            raise RuntimeError("AsyncStatsReducer.latest() before submit()")
        if self._work is not None:
            self._work.wait()
        torch.cuda.current_stream(self.env.device).wait_stream(self.side)
        return self._buf


def stats_to_dict(totals: np.ndarray, ret_scale: float) -> dict:
    """int64 totals -> named dict with derived rates (what scripts/main.py:203-227 tracks by hand)."""
    t = [int(v) for v in np.asarray(totals).reshape(-1)[:len(STAT_NAMES)]]
    d = dict(zip(STAT_NAMES, t))
    d["sum_return1"] = d.pop("sum_return1_fx") / ret_scale
    d["sum_return2"] = d.pop("sum_return2_fx") / ret_scale
    ep = max(d["episodes"], 1)
    d["collision_rate"] = d["collisions"] / ep
    d["merge_success_rate"] = d["merges_ok"] / ep
    d["win_rate_p1"] = d["wins_p1"] / ep
    d["win_rate_p2"] = d["wins_p2"] / ep
    d["mean_length"] = d["sum_length"] / ep
    d["mean_return1"] = d["sum_return1"] / ep
    d["mean_return2"] = d["sum_return2"] / ep
    return d
