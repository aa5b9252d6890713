#!/usr/bin/env python
"""Times the small kernels that are not on the per-step path: mg_reset (all envs / masked) and
mg_sample_actions at 2^20 envs."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

n = 1 << 20
env = mg.MergeVecEnv(n, auto_reset=False, episode_info=False)
envr = mg.MergeVecEnv(n, auto_reset=False, episode_info=False, reset_mode="random", reset_seed=7)
mask = (torch.arange(n, device="cuda") % 7 == 0)


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return round(1e3 * e0.elapsed_time(e1) / iters, 2)


res = {"reset_all_us": timed(lambda: env.reset()), "reset_masked_us": timed(lambda: env.reset(mask)),
       "reset_random_us": timed(lambda: envr.reset()), "sample_actions_us": timed(lambda: env.sample_actions())}
print(json.dumps({"envs": n, **res}))
