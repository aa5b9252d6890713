#!/usr/bin/env python
"""Small end-to-end exercise of every kernel, for `compute-sanitizer --tool memcheck`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

for n in (1, 77, 256, 1000):
    for mode in ("pvp", "pve"):
        for rm in ("fixed", "random"):
            env = mg.MergeVecEnv(n, mode=mode, out_slots=2, reset_mode=rm)
            rec = mg.TransitionRecorder(env, 500, track_env_ids=True)
            obs = env.reset()
            for t in range(40):
                a1, a2 = env.sample_actions()
                out = env.step(a1, a2)
                rec.record(obs, a1, a2 if a2 is not None else None, out)
                obs = out[0]
            env.step(a1.to(torch.int64), None if a2 is None else a2.to(torch.int64))
            K = 8
            o = torch.empty(K, n, 10, device="cuda"); r = torch.empty(K, n, 2, device="cuda")
            d = torch.empty(K, n, dtype=torch.uint8, device="cuda"); i = torch.empty(K, n, dtype=torch.uint8, device="cuda")
            ac = torch.empty(K, n, 2, dtype=torch.uint8, device="cuda")
            env.rollout(K, obs=o, rew=r, done=d, info=i, actions=ac)
            env.rollout(K)
            m = torch.zeros(n, dtype=torch.bool); m[::2] = True
            env.reset(m)
            h = np.zeros(n, np.uint8)
            env.step_host(h, h if mode == "pvp" else None)
            env.stats()
    pol = mg.MLPPolicy(10, 5, seed=1); hd = mg.HDQNPolicy(seed=2)
    q = torch.empty(n, 5, device="cuda")
    pol.act(obs.contiguous(), q_out=q); hd.act(obs.contiguous())
torch.cuda.synchronize()
print("sanitize target done")
