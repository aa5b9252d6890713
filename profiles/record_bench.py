#!/usr/bin/env python
"""Throughput of mg_record_transitions (replay rows) at 2^20 envs (bandwidth) and 4096 envs (launch count), eager and
as a CUDA graph of 32 calls."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(n, iters, graph):
    import merging_gym_b200 as mg
    env = mg.MergeVecEnv(n, out_slots=2)
    rec = mg.TransitionRecorder(env, 8 * n)
    env.rollout(300)
    obs = env.obs_buf[env._slot]
    a1, a2 = env.sample_actions()
    out = env.step(a1, a2)
    for _ in range(3):
        rec.record(obs, a1, a2, out)
    torch.cuda.synchronize()
    per = 1
    if graph:
        per = 32
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(per):
                    rec.record(obs, a1, a2, out)
        torch.cuda.current_stream().wait_stream(side)
        fn = g.replay
    else:
        fn = lambda: rec.record(obs, a1, a2, out)
    fn(); torch.cuda.synchronize()
    c0 = int(rec.counter.item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (iters * per)
    rows = (int(rec.counter.item()) - c0) / (iters * per)
    # info + done per env; s, s', a, r read and the 88-byte row written per stored row
    byt = n * (1 + 1) + rows * (40 + 40 + 2 + 4 + 88)
    return {"envs": n, "graph": graph, "rows_per_call": rows, "us_per_call": ms * 1e3, "rows_per_s": rows / (ms * 1e-3),
            "algorithmic_GBps": byt / (ms * 1e-3) / 1e9}


if __name__ == "__main__":
    res = [one(1 << 20, 50, False), one(1 << 20, 5, True), one(4096, 200, False), one(4096, 20, True)]
    print(json.dumps({"results": res}))
