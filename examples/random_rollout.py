#!/usr/bin/env python
"""Random-action rollout of N two-player merge envs on one GPU, the way the reference scripts drive a single env
(`env.step(env.action_space.sample(), ...)`, scripts/main.py:26), and the episode statistics they track by hand.

    python examples/random_rollout.py [--envs 1048576] [--steps 500]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=500)
args = ap.parse_args()

env = mg.MergeVecEnv(args.envs, mode="pvp", auto_reset=True)
obs = env.reset()                                        # f32[N,10] on the GPU
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(args.steps):
    a1, a2 = env.sample_actions()                        # uint8[N] each, Philox on the device
    obs, rew, done, info = env.step(a1, a2)              # device tensors, no host round trip
torch.cuda.synchronize(); dt = time.perf_counter() - t0
s = env.stats()
print(f"{args.envs} envs x {args.steps} steps in {dt:.3f} s = {args.envs * args.steps / dt:.3e} env-steps/s (eager launches)")
print(f"episodes {s['episodes']}  collision rate {s['collision_rate']:.3f}  merge success {s['merge_success_rate']:.3f}  "
      f"mean length {s['mean_length']:.1f}  P1 / P2 win rate {s['win_rate_p1']:.3f} / {s['win_rate_p2']:.3f}")
