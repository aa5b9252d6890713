#!/usr/bin/env python
"""bench_policy.py — BASELINE.json configs[4]: pve rollout with the DQN policy forward in the loop,
2^18 envs per GPU, observations and actions handed over on the device (no host round-trip).

    python bench_policy.py [--envs 262144] [--steps 400] [--policy dqn|hdqn] [--backend fused|tf32x3|torch]

One step = fused Q-network forward + arg-max (`mg_mlp_act`) -> `mg_step` (pve, auto-reset), captured
in a CUDA graph.  Prints one JSON line with env-steps/s and the share of the step spent in the env
kernel vs the policy kernel (each timed separately with CUDA events).  Not the headline bench
(`bench.py`); weights are the reference's shipped DQN checkpoint when the fixture is present,
random-init (the reference's `uniform_(0,1)` init) otherwise — h-DQN weights were never shipped.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import merging_gym_b200 as mg  # noqa: E402


def timed(fn, iters, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--policy", default="dqn", choices=["dqn", "hdqn"])
    ap.add_argument("--backend", default="fused", choices=["fused", "tf32x3", "torch"])
    args = ap.parse_args()
    n = args.envs
    env = mg.MergeVecEnv(n, mode="pve", auto_reset=True, episode_info=False)
    weights = "random-init (reference init)"
    if args.policy == "dqn":
        fx = os.path.join(ROOT, "tests", "golden", "dqn_policies.npz")
        sd = None
        if os.path.exists(fx):
            z = np.load(fx)
            sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("L1_1445/") and "traj" not in k and "result" not in k}
            weights = "test_params/dqn/2022--03--31 14:45:59.../eval.pth"
        pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=args.backend)
        flops = 2 * (10 * 200 + 200 * 100 + 100 * 5)
    else:
        pol = mg.HDQNPolicy(backend=args.backend)
        flops = 2 * (10 * 200 + 200 * 100 + 100 * 3) + 2 * (11 * 200 + 200 * 100 + 100 * 5)
    act = torch.empty(n, dtype=torch.uint8, device="cuda")
    obs0 = env.reset()

    def one_step(obs):
        pol.act(obs, out=act)
        return env.step(act, None)[0]

    obs = obs0
    for _ in range(20):
        obs = one_step(obs)
    torch.cuda.synchronize()
    # with out_slots=1 the obs buffer is fixed, so a captured step can be replayed indefinitely
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        one_step(env.obs_buf[0])
    ms_step = timed(g.replay, args.steps)
    ms_pol = timed(lambda: pol.act(env.obs_buf[0], out=act), 50)
    ms_env = timed(lambda: env.step(act, None), 50)
    st = env.stats()
    line = {"metric": "env_steps_per_sec", "value": n / (ms_step * 1e-3), "unit": "env-steps/s", "n_gpus": 1,
            "steps": args.steps, "ms_per_step": ms_step, "dtype": "f64 env / f32 policy", "data": "synthetic",
            "config": {"workload": f"pve, {n} envs, {args.policy} greedy policy in the loop, auto-reset "
                                   "(BASELINE.json configs[4])", "backend": args.backend, "weights": weights,
                       "launch": "CUDA graph of one policy+env step"},
            "policy_kernel_ms": ms_pol, "env_kernel_ms": ms_env,
            "env_share": ms_env / (ms_env + ms_pol),
            "policy_tflops": n * flops / (ms_pol * 1e-3) / 1e12,
            "episode_stats": {k: st[k] for k in ("episodes", "collision_rate", "win_rate_p1", "mean_length", "mean_return1")}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
