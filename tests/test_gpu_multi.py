"""Multi-GPU behaviour (skipped on single-GPU boxes): envs on two devices driven from one process,
and the torchrun / NCCL statistics reduction with world-size invariance of the reduced totals."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")


@needs2
def test_two_devices_one_process():
    import merging_gym_b200 as mg
    n = 4096
    whole = mg.MergeVecEnv(2 * n, device="cuda:0", seed=3)
    halves = [mg.MergeVecEnv(n, device=f"cuda:{d}", seed=3, env_id_base=d * n) for d in (0, 1)]
    whole.rollout(300)
    for h in halves:
        for t in range(300):
            h.step(*h.sample_actions())
    assert halves[1].pos1.device.index == 1
    got = torch.cat([halves[0].pos1.cpu(), halves[1].pos1.cpu()])
    assert torch.equal(got, whole.pos1.cpu())
    tot = halves[0].stats_tensor().cpu() + halves[1].stats_tensor().cpu()
    assert torch.equal(tot, whole.stats_tensor().cpu())


WORKER = r"""
import json, os, sys, torch
sys.path.insert(0, os.environ["MG_ROOT"])
import merging_gym_b200 as mg
rank, local, world = mg.init_distributed()
total, K = 1 << 16, 260
base, count = mg.shard_range(total, rank, world)
env = mg.MergeVecEnv(count, device=f"cuda:{local}", seed=11, env_id_base=base)
red = mg.AsyncStatsReducer(env, banked=bool(int(os.environ["MG_BANKED"])))
for t in range(K):
    env.step(*env.sample_actions())
    if t % 64 == 63:
        red.submit()                      # async NCCL all-reduce on a side stream
env.rollout(40)
red.submit()
tot = red.latest().cpu().tolist()
if rank == 0:
    print("STATS " + json.dumps(tot))
torch.distributed.barrier(); torch.distributed.destroy_process_group()
"""


@needs2
@pytest.mark.parametrize("banked", [0, 1])
def test_torchrun_nccl_stats_are_world_size_invariant(tmp_path, banked):
    """banked=1: the launching stream carries no statistics kernel (the side stream drains the retired bank)."""
    import merging_gym_b200 as mg
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MG_ROOT=ROOT, MG_BANKED=str(banked))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("STATS ")][-1]
    reduced = json.loads(line[6:])
    single = mg.MergeVecEnv(1 << 16, seed=11)
    for t in range(260):
        single.step(*single.sample_actions())
    single.rollout(40)
    assert reduced == single.stats_tensor().cpu().tolist()
