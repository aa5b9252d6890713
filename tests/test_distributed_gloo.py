"""The N>1 path on CPU: world_size-2 gloo run of the sharding + statistics all-reduce logic,
driven by the oracle so the numbers are real (each rank steps its own shard of global env ids
with the Philox action stream; the reduced statistics must equal a single-process run)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from merging_gym_b200 import _native as nat
from merging_gym_b200.sharding import all_reduce_stats, shard_range, stats_to_dict
from oracle import merge_oracle as mo

TOTAL, STEPS, SEED = 600, 260, 0x5EED


def _run_shard(base, count):
    env = mo.RefVecEnv(count, pvp=True, auto_reset=True)
    for t in range(STEPS):
        a1, a2 = mo.philox_actions(count, SEED, base, t)
        env.step(a1, a2)
    s = env.stats
    v = np.zeros(nat.STATS_COLS, np.int64)
    v[:8] = [s[k] for k in nat.STAT_NAMES[:8]]
    v[8] = int(np.rint(s["sum_return1"] * 2 ** 24)); v[9] = int(np.rint(s["sum_return2"] * 2 ** 24))
    return v


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, count = shard_range(TOTAL, rank, world)
    t = all_reduce_stats(torch.from_numpy(_run_shard(base, count)))
    q.put((rank, t.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_stats_reduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    single = _run_shard(0, TOTAL)
    assert np.array_equal(res[0], res[1])
    assert np.array_equal(res[0][:8], single[:8])                 # counts: world-size invariant
    assert np.abs(res[0][8:10] - single[8:10]).max() <= 4         # fixed-point of float sums: +-ulp noise
    d = stats_to_dict(res[0], 2.0 ** 24)
    assert d["episodes"] > 0 and 0.2 < d["collision_rate"] < 0.6


def test_all_reduce_is_noop_without_group():
    t = torch.arange(16)
    assert all_reduce_stats(t) is t
