#!/usr/bin/env python
"""bench_policy.py — BASELINE.json configs[4]: pve rollout with the DQN / h-DQN policy forward in the loop,
2^18 envs per GPU, observations and actions handed over on the device (no host round-trip).

    python bench_policy.py [--envs 262144] [--steps 400] [--policy dqn|hdqn] [--backend fused|tf32x3|f16x3|torch]

One step = fused Q-network forward + arg-max (`mg_mlp_act` / `mg_mlp_act_tc`) -> `mg_step` (pve, auto-reset, RANDOM
starts so that the envs de-synchronise: with the fixed start and a greedy policy all envs would run in lockstep).
Everything is timed as CUDA-graph replays: graph A holds K x (policy + env step), graph B the same K policy launches
alone; the env kernel's share of a step is (A - B) / A, free of Python launch latency.  `measure()` is what
`bench.py` puts into its `policy_in_loop` object.  Weights: the reference's shipped DQN checkpoint from the committed
fixture `tests/golden/dqn_policies.npz` (h-DQN weights were never shipped: the reference's `uniform_(0,1)` init).
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

DQN_FLOPS = 2 * (10 * 200 + 200 * 100 + 100 * 5)
HDQN_FLOPS = 2 * (10 * 200 + 200 * 100 + 100 * 3) + 2 * (11 * 200 + 200 * 100 + 100 * 5)


def graph_ms(fn, k, replays=5):
    """ms per call of `fn`, measured as `replays` replays of a CUDA graph of k calls."""
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (k * replays)


def load_policy(policy, backend, device="cuda"):
    if policy == "dqn":
        fx = os.path.join(ROOT, "tests", "golden", "dqn_policies.npz")
        sd, weights = None, "random-init (reference init)"
        if os.path.exists(fx):
            z = np.load(fx)
            sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith("L1_1445/") and "traj" not in k and "result" not in k}
            weights = "test_params/dqn/2022--03--31 14:45:59.../eval.pth (tests/golden/dqn_policies.npz)"
        return mg.MLPPolicy(10, 5, device=device, state_dict=sd, backend=backend), DQN_FLOPS, weights
    return mg.HDQNPolicy(device=device, backend=backend), HDQN_FLOPS, "random-init (reference init; no h-DQN weights ship)"


def measure(n=1 << 18, policy="dqn", backend="fused", k=50, replays=4, device="cuda", reset_mode="random", mix_steps=300,
            pdl=True):
    env = mg.MergeVecEnv(n, mode="pve", device=device, auto_reset=True, episode_info=False, reset_mode=reset_mode)
    pol, flops, weights = load_policy(policy, backend, device)
    for q in (pol, getattr(pol, "meta", None), getattr(pol, "ctrl", None)):    # rollout loop: the kernel before a policy
        if hasattr(q, "pdl"):                                                  # kernel is the env step -> PDL is valid
            q.pdl = pdl
    act = torch.empty(n, dtype=torch.uint8, device=device)
    obs = env.obs_buf[0]                           # out_slots=1: one fixed observation buffer, replayable

    def policy_only():
        pol.act(obs, out=act)

    def one_step():
        pol.act(obs, out=act)
        env.step_async(act, None)

    def one_step_fused():                          # mg_policy_step: forward + arg-max + env step in ONE launch
        if policy == "dqn":
            env.policy_step(pol)
        else:
            pol.step(env)

    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                  # warm-up (module loading) and episode mixing outside any capture
        for _ in range(mix_steps):
            one_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    env.stats(reset=True)
    ms_step = graph_ms(one_step, k, replays)
    st = env.stats()
    ms_pol = graph_ms(policy_only, k, replays)
    fused = None
    if backend in ("fused", "tf32x3", "f16x3"):
        for _ in range(3):
            one_step_fused()
        ms_f = graph_ms(one_step_fused, k, replays)
        fused = {"value": n / (ms_f * 1e-3), "ms_per_step": ms_f, "launches_per_step": 1 if policy == "dqn" else 2,
                 "note": "mg_policy_step: the env step is the epilogue of the policy kernel (no action array, no second "
                         "launch); bit-identical to the separate launches (tests/test_gpu_fused_policy.py)"}
    return {"value": n / (ms_step * 1e-3), "fused_step": fused, "unit": "env-steps/s per GPU", "envs": n, "policy": policy, "backend": backend,
            "ms_per_step": ms_step, "policy_ms": ms_pol, "env_ms": ms_step - ms_pol,
            "env_share": (ms_step - ms_pol) / ms_step,
            "policy_tflops": n * flops / (ms_pol * 1e-3) / 1e12, "weights": weights,
            "pdl": pdl, "launch": f"CUDA graph of {k} x (policy forward+argmax, mg_step) vs a graph of {k} policy launches alone; "
                      "env_ms is the difference", "reset_mode": reset_mode,
            "episode_stats": {q: st[q] for q in ("episodes", "collision_rate", "win_rate_p1", "mean_length", "mean_return1")}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--policy", default="dqn", choices=["dqn", "hdqn"])
    ap.add_argument("--backend", default="fused", choices=["fused", "tf32x3", "f16x3", "torch"])
    ap.add_argument("--reset-mode", default="random", choices=["fixed", "random"])
    ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch of the policy kernels (0/1)")
    args = ap.parse_args()
    k = 50
    r = measure(args.envs, args.policy, args.backend, k=k, replays=max(1, args.steps // k), reset_mode=args.reset_mode, pdl=bool(args.pdl))
    line = {"metric": "env_steps_per_sec", "unit": "env-steps/s", "n_gpus": 1, "steps": args.steps,
            "dtype": "f64 env / f32 policy", "data": "synthetic",
            "config": {"workload": f"pve, {args.envs} envs, {args.policy} greedy policy in the loop, auto-reset, "
                                   f"{args.reset_mode} starts (BASELINE.json configs[4])"}}
    line.update(r)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
