#!/usr/bin/env python
"""Policy-in-the-loop step rate: eager Python loop vs `GraphedPolicyRollout` (K = 32 steps per CUDA-graph replay) with
separate policy / env launches vs `GraphedPolicyRollout(fused=True)` (one `mg_policy_step` launch per step), greedy DQN
checkpoint vs the constant-speed opponent, with and without the replay recorder in the loop."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import merging_gym_b200 as mg

z = np.load(os.path.join(ROOT, "tests", "golden", "dqn_policies.npz"))
sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("L1_1445/") and "traj" not in k and "result" not in k}
res = {}
K = 32


def timed(roll, reps=10):
    roll.run(2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); roll.run(reps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * K)


for n in (4096, 1 << 16, 1 << 18):
    for backend in ("fused", "tf32x3", "f16x3"):
        for with_rec in (False, True):
            pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
            env = mg.MergeVecEnv(n, mode="pve", out_slots=1, reset_mode="random")
            rec = mg.TransitionRecorder(env, 1 << 22) if with_rec else None
            dt_g = timed(mg.GraphedPolicyRollout(env, pol, k_steps=K, after_step=rec.record if rec else None))
            obs = env.obs_buf[0]
            def eager():
                prev = obs.clone() if rec else None
                a = pol.act(obs); out = env.step(a, None)
                if rec: rec.record(prev, a, None, out)
            for _ in range(10): eager()
            torch.cuda.synchronize(); t = time.perf_counter()
            for _ in range(100): eager()
            torch.cuda.synchronize(); dt_e = (time.perf_counter() - t) / 100
            env2 = mg.MergeVecEnv(n, mode="pve", out_slots=2 if with_rec else 1, reset_mode="random")
            rec2 = mg.TransitionRecorder(env2, 1 << 22) if with_rec else None
            dt_f = timed(mg.GraphedPolicyRollout(env2, pol, k_steps=K, after_step=rec2.record if rec2 else None, fused=True))
            res[f"n={n} {backend}{' +recorder' if with_rec else ''}"] = {
                "graphed_us_per_step": round(dt_g * 1e6, 2), "graphed_fused_us_per_step": round(dt_f * 1e6, 2),
                "eager_us_per_step": round(dt_e * 1e6, 1), "graphed_fused_env_steps_per_s": round(n / dt_f)}
print(json.dumps(res, indent=1))
