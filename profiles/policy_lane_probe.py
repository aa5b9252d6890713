#!/usr/bin/env python
"""Can the env step of one lane run UNDER the policy kernel of another?  Times, at 2^17 envs per lane (tf32x3 policy):
the policy kernel alone, the env step alone (regular and MG_FLAG_NO_SMEM / coresident variant), and both at once on two
streams (policy first / env first), as CUDA graphs of 20 pairs."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import merging_gym_b200 as mg

n = 1 << 17
z = np.load(os.path.join(ROOT, "tests", "golden", "dqn_policies.npz"))
sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("L1_1445/") and "traj" not in k and "result" not in k}
pol = mg.MLPPolicy(10, 5, state_dict=sd, backend="tf32x3")
envA = mg.MergeVecEnv(n, mode="pve", reset_mode="random", episode_info=False)
res = {}


def graph_ms(fn, k=20, reps=5):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (k * reps) * 1e3


act = torch.zeros(n, dtype=torch.uint8, device="cuda")
obsA = envA.obs_buf[0]
res["policy_alone_us"] = graph_ms(lambda: pol.act(obsA, out=act))
for co in (False, True):
    envB = mg.MergeVecEnv(n, mode="pve", reset_mode="random", episode_info=False, coresident=co)
    envB.rollout(100)
    a = envB.sample_actions()[0].clone()
    tag = "coresident" if co else "regular"
    res[f"env_alone_{tag}_us"] = graph_ms(lambda: envB.step_async(a, None))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both(policy_first):
        def fn():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur); s2.wait_stream(cur)
            order = [(s1, lambda: pol.act(obsA, out=act)), (s2, lambda: envB.step_async(a, None))]
            if not policy_first:
                order.reverse()
            for st, f in order:
                with torch.cuda.stream(st):
                    f()
            cur.wait_stream(s1); cur.wait_stream(s2)
        return fn
    res[f"both_{tag}_policy_first_us"] = graph_ms(both(True))
    res[f"both_{tag}_env_first_us"] = graph_ms(both(False))
print(json.dumps(res, indent=1))
