#!/usr/bin/env python
"""Times mg_rollout (K fused steps per launch, in-kernel Philox actions) at 2^20 envs:
with all per-step outputs (obs/rew/done/info time-major), without outputs, and with actions too."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

n, K = 1 << 20, int(os.environ.get("K", 64))
env = mg.MergeVecEnv(n, episode_info=False)
env.rollout(300)
obs = torch.empty(K, n, 10, device="cuda"); rew = torch.empty(K, n, 2, device="cuda")
done = torch.empty(K, n, dtype=torch.uint8, device="cuda"); info = torch.empty(K, n, dtype=torch.uint8, device="cuda")
acts = torch.empty(K, n, 2, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {}
for name, kw in [("all_outputs", dict(obs=obs, rew=rew, done=done, info=info)),
                 ("all_outputs+actions", dict(obs=obs, rew=rew, done=done, info=info, actions=acts)),
                 ("no_outputs", {}), ("obs_only", dict(obs=obs))]:
    ms = timed(lambda: env.rollout(K, refresh_obs=False, **kw))
    res[name] = {"us_per_step": 1e3 * ms / K, "env_steps_per_s": n * K / (ms * 1e-3)}
print(json.dumps({"K": K, "envs": n, "lib": os.environ.get("MERGING_B200_LIB", "default"), **res}))
