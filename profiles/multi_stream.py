#!/usr/bin/env python
"""Experiment: R independent 2^20-env shards stepped round-robin on S CUDA streams (captured as a forked
CUDA graph), so that the ramp-up of one shard's launch overlaps the drain of another's.  Prints the
effective µs per launch and the algorithmic GB/s for S = 1, 2, 4.  (S = 1 is bench.py's headline set-up.)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

n, R = 1 << 20, int(os.environ.get("R", 4))
G, REPS = 200, 10
dev = torch.device("cuda", 0)
envs = [mg.MergeVecEnv(n, env_id_base=r * n, episode_info=False, out_slots=2) for r in range(R)]
acts = [envs[i % R].sample_actions(i) for i in range(8)]
acts = [(a.clone(), b.clone()) for a, b in acts]
for e in envs:
    e.rollout(300, step0=1000)
torch.cuda.synchronize()

res = {}
for S in (1, 2, 4):
    if R % S:
        continue
    streams = [torch.cuda.Stream() for _ in range(S)]

    def do_steps(k):
        main = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(main)
        for i in range(k):
            with torch.cuda.stream(streams[i % S]):
                envs[i % R].step_async(*acts[i % 8])
        for s in streams:
            main.wait_stream(s)

    do_steps(8)
    torch.cuda.synchronize()
    for e in envs:
        e._slot = 0
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        do_steps(G)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (G * REPS)
    res[f"streams_{S}"] = {"us_per_launch": round(us, 3), "algorithmic_GBps": round(156 * n / us / 1e3, 1),
                           "env_steps_per_s": n / (us * 1e-6)}
print(json.dumps({"envs_per_shard": n, "shards": R, "graph_launches": G, **res}))
