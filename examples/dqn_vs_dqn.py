#!/usr/bin/env python
"""Two of the reference's shipped DQN checkpoints playing each other in 4096 envs: the agent sees the observation,
the opponent its mirror image (`state[5:] + state[:5]`, scripts/main.py:196-199), both explore with the scripts'
`randn() <= 0.7` rule, everything stays on the GPU.  Weights come from tests/golden/dqn_policies.npz.

    python examples/dqn_vs_dqn.py [--backend fused|tf32x3|f16x3]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import merging_gym_b200 as mg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="fused", choices=["fused", "tf32x3", "f16x3"])
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=1000)
args = ap.parse_args()

z = np.load(os.path.join(ROOT, "tests", "golden", "dqn_policies.npz"))
weights = lambda tag: {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/") and "traj" not in k and "result" not in k}
agent = mg.MLPPolicy(10, 5, state_dict=weights("L2_2133"), backend=args.backend)
opponent = mg.MLPPolicy(10, 5, state_dict=weights("L1_2136"), backend=args.backend)

env = mg.MergeVecEnv(args.envs, mode="pvp", auto_reset=True)
gen = torch.Generator(device="cuda").manual_seed(0)
obs = env.reset()
for _ in range(args.steps):
    a1 = mg.explore(agent.act(obs), 5, generator=gen)
    a2 = mg.explore(opponent.act(obs, mirror=True), 5, generator=gen)      # = opponent.act(env.opponent_view(obs)), fused
    obs, rew, done, info = env.step(a1, a2)
s = env.stats()
print(f"{s['episodes']} episodes: agent wins {s['win_rate_p1']:.3f}, opponent wins {s['win_rate_p2']:.3f}, "
      f"collisions {s['collision_rate']:.3f}, mean return {s['mean_return1']:.3f} / {s['mean_return2']:.3f}")
