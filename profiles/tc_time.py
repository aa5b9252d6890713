#!/usr/bin/env python
"""Times mg_mlp_act_tc (tf32x3 backend) at 2^18 envs and checks it against the fused fp32 kernel.
MERGING_B200_LIB selects a prebuilt library variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

n = 1 << 18
env = mg.MergeVecEnv(n, seed=3); env.rollout(120); obs = env.step(*env.sample_actions())[0].clone()
f = mg.MLPPolicy(10, 5, seed=7)
tc = mg.MLPPolicy(10, 5, state_dict=f.state_dict(), backend=os.environ.get("TC_BACKEND", "tf32x3"))
qf = torch.empty(n, 5, device="cuda"); qt = torch.empty(n, 5, device="cuda")
af = f.act(obs, q_out=qf); at = tc.act(obs, q_out=qt)
torch.cuda.synchronize()
err = ((qf - qt).abs().max() / qf.abs().max()).item()
act = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(5):
    tc.act(obs, out=act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    tc.act(obs, out=act)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"backend": os.environ.get("TC_BACKEND", "tf32x3"), "lib": os.path.basename(os.environ.get("MERGING_B200_LIB", "default")), "us": round(1e3 * e0.elapsed_time(e1) / 50, 2),
                  "rel_err_vs_fused": err, "action_agreement": (af == at).float().mean().item()}))
