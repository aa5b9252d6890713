#!/usr/bin/env python
"""Policy-in-the-loop step rate, eager Python loop vs `GraphedPolicyRollout` (K = 32 steps per CUDA-graph replay),
greedy DQN checkpoint vs the constant-speed opponent, with and without the replay recorder in the loop."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import merging_gym_b200 as mg

z = np.load(os.path.join(ROOT, "tests", "golden", "dqn_policies.npz"))
sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("L1_1445/") and "traj" not in k and "result" not in k}
res = {}
for n in (4096, 1 << 16, 1 << 18):
    for backend in ("fused", "tf32x3"):
        for with_rec in (False, True):
            env = mg.MergeVecEnv(n, mode="pve", out_slots=1)
            pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
            rec = mg.TransitionRecorder(env, 1 << 22) if with_rec else None
            K = 32
            roll = mg.GraphedPolicyRollout(env, pol, k_steps=K, after_step=rec.record if rec else None)
            roll.run(2); torch.cuda.synchronize()
            t = time.perf_counter(); roll.run(10); torch.cuda.synchronize(); dt_g = (time.perf_counter() - t) / (10 * K)
            obs = env.obs_buf[0]
            def eager():
                prev = obs.clone() if rec else None
                a = pol.act(obs); out = env.step(a, None)
                if rec: rec.record(prev, a, None, out)
            for _ in range(10): eager()
            torch.cuda.synchronize(); t = time.perf_counter()
            for _ in range(100): eager()
            torch.cuda.synchronize(); dt_e = (time.perf_counter() - t) / 100
            res[f"n={n} {backend}{' +recorder' if with_rec else ''}"] = {"graphed_us_per_step": round(dt_g * 1e6, 1), "eager_us_per_step": round(dt_e * 1e6, 1),
                                                                     "graphed_env_steps_per_s": round(n / dt_g)}
print(json.dumps(res, indent=1))
