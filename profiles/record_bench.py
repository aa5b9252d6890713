#!/usr/bin/env python
"""Throughput of mg_record_transitions (replay rows) at 2^20 envs."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg

n = 1 << 20
env = mg.MergeVecEnv(n, out_slots=2)
rec = mg.TransitionRecorder(env, 8 * n)
env.rollout(300)
obs = env.obs_buf[env._slot]
a1, a2 = env.sample_actions()
out = env.step(a1, a2)
for _ in range(3):
    rec.record(obs, a1, a2, out)
torch.cuda.synchronize()
c0 = int(rec.counter.item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    rec.record(obs, a1, a2, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
rows = (int(rec.counter.item()) - c0) / 50
byt = n * (1 + 1) + rows * (40 + 40 + 2 + 4 + 88)      # info + done per env; s, s', a, r read and the 88-byte row written per stored row
print(json.dumps({"envs": n, "rows_per_call": rows, "us_per_call": ms * 1e3, "rows_per_s": rows / (ms * 1e-3),
                  "algorithmic_GBps": byt / (ms * 1e-3) / 1e9}))
