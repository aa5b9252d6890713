#!/usr/bin/env python
"""One-off stress of the f16x3 policy kernel and its fused env variant: many launches at random sizes (ragged tiles, 1 ...
many tiles per SM), every launch checked against the fp32 kernel (Q within 2e-5 of the largest Q, same action wherever the
top-2 margin is clear) and for bitwise repeatability; then 300 fused steps ≡ act + step.  A protocol race would show up
as a trap (sticky CUDA error), a stale tile or a non-repeatable launch."""
import json, os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

random.seed(1)
f = mg.MLPPolicy(10, 5, seed=5)
h = mg.MLPPolicy(10, 5, state_dict=f.state_dict(), backend="f16x3")
launches, worst = 0, 0.0
for trial in range(60):
    n = random.choice([1, 127, 128, 129, 4096, 18944, 18945, 70001, 148 * 128 * 2 + 5, 262144, random.randint(1, 300000)])
    env = mg.MergeVecEnv(n, seed=trial, reset_mode="random")
    env.rollout(random.randint(0, 200))
    qf = torch.empty(n, 5, device="cuda"); qh = torch.empty(n, 5, device="cuda"); q2 = torch.empty(n, 5, device="cuda")
    for t in range(25):
        obs = env.step(*env.sample_actions())[0]
        af = f.act(obs, q_out=qf); ah = h.act(obs, q_out=qh).clone(); h.act(obs, q_out=q2)
        launches += 2
        scale = qf.abs().max()
        err = ((qf - qh).abs().max() / scale).item()
        worst = max(worst, err)
        assert err < 2e-5, (n, t, err)
        assert torch.equal(qh, q2), (n, t)
        top2 = qf.topk(2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * scale
        assert torch.equal(af[clear], ah[clear]), (n, t)
# fused: mg_policy_step(f16x3) == act + step, bit for bit, over 300 steps with auto-reset and random starts
n = 148 * 128 * 3 + 77
ea = mg.MergeVecEnv(n, mode="pve", seed=9, reset_mode="random", out_slots=1)
eb = mg.MergeVecEnv(n, mode="pve", seed=9, reset_mode="random", out_slots=1)
for t in range(300):
    oa = ea.step(h.act(ea.obs_buf[0]), None)
    ob = eb.policy_step(h)
    launches += 2
    assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1]) and torch.equal(oa[2], ob[2]), t
assert torch.equal(ea.pos1, eb.pos1) and torch.equal(ea.meta, eb.meta)
torch.cuda.synchronize()
print(json.dumps({"policy_launches_checked": launches, "worst_rel_q_err_vs_fp32_kernel": worst, "fused_steps": 300, "ok": True}))
