#!/usr/bin/env python
"""f16x3 vs tf32x3 vs the fp32 kernel: Q error against an fp64 evaluation, action agreement, time at 2^18 envs."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402


def q64(p, x):
    x = x.double().cpu()
    h = torch.relu(x @ p.w1.double().cpu().t() + p.b1.double().cpu())
    h = torch.relu(h @ p.w2.double().cpu().t() + p.b2.double().cpu())
    return h @ p.w3.double().cpu().t() + p.b3.double().cpu()


res = []
for in_dim, out_dim, n in [(10, 5, 5000), (11, 3, 1031), (10, 5, 1), (10, 5, 1 << 18)]:
    env = mg.MergeVecEnv(n, seed=3); env.rollout(120); obs = env.step(*env.sample_actions())[0].clone()
    goal = torch.randint(0, 3, (n,), dtype=torch.uint8, device="cuda") if in_dim == 11 else None
    f = mg.MLPPolicy(in_dim, out_dim, seed=7)
    x = obs if goal is None else torch.cat([goal.float().unsqueeze(1), obs], 1)
    ref = q64(f, x); scale = ref.abs().max().item()
    row = {"in": in_dim, "out": out_dim, "n": n}
    for be in ("fused", "tf32x3", "f16x3"):
        p = mg.MLPPolicy(in_dim, out_dim, state_dict=f.state_dict(), backend=be)
        q = torch.empty(n, out_dim, device="cuda")
        a = p.act(obs, goal=goal, q_out=q)
        torch.cuda.synchronize()
        row[be] = {"max_rel_err": (q.double().cpu() - ref).abs().max().item() / scale,
                   "agree_fp64_argmax": (a.cpu().long() == ref.argmax(1)).float().mean().item()}
        if n == 1 << 18:
            act = torch.empty(n, dtype=torch.uint8, device="cuda")
            for _ in range(5):
                p.act(obs, out=act)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                p.act(obs, out=act)
            e1.record(); torch.cuda.synchronize()
            row[be]["us"] = round(1e3 * e0.elapsed_time(e1) / 50, 2)
    res.append(row)
    print(json.dumps(row), flush=True)
