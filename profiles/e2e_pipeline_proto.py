#!/usr/bin/env python
"""Prototype behind mg_step_host's chunked pipeline: what does the host-buffer step cost when (a) the actions are
already in pinned memory (no per-step copy into the pinned staging buffers), and (b) the n envs are stepped in C
chunks so that chunk c's D2H overlaps chunk c+1's H2D + kernel?"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg

n = 1 << 20
env = mg.MergeVecEnv(n, episode_info=False); env.rollout(300)
a1, a2 = env.sample_actions(); h1 = a1.cpu().numpy().copy(); h2 = a2.cpu().numpy().copy()


def timed(fn, iters=30):
    for _ in range(3): fn()
    t = time.perf_counter()
    for _ in range(iters): fn()
    dt = (time.perf_counter() - t) / iters
    return f"{dt * 1e3:.3f} ms/step  {n / dt:.3e} env-steps/s  {n * 52 / dt / 1e9:.1f} GB/s"


print("step_host (pageable numpy actions copied into pinned staging every step):", timed(lambda: env.step_host(h1, h2)))
t = time.perf_counter()
for _ in range(30):
    env._host["a1"].numpy()[:] = h1; env._host["a2"].numpy()[:] = h2
print("  of which the numpy -> pinned staging copies: %.3f ms/step" % ((time.perf_counter() - t) / 30 * 1e3))

# chunked prototype with torch streams: same kernels through MergeVecEnv views is not possible, so emulate the
# traffic: per chunk H2D(actions) -> a kernel of the same duration (mg_step on a separate small env) -> D2H(outputs)
pa1 = torch.from_numpy(h1).pin_memory(); pa2 = torch.from_numpy(h2).pin_memory()
hob = torch.empty(n, 10).pin_memory(); hrw = torch.empty(n, 2).pin_memory()
hdn = torch.empty(n, dtype=torch.uint8).pin_memory(); hin = torch.empty(n, dtype=torch.uint8).pin_memory()
for C in (1, 2, 4, 8):
    m = n // C
    subs = [mg.MergeVecEnv(m, episode_info=False, env_id_base=c * m) for c in range(C)]
    for s in subs: s.rollout(300)
    side = torch.cuda.Stream()
    evs = [torch.cuda.Event() for _ in range(C)]

    def step():
        main = torch.cuda.current_stream()
        for c, s in enumerate(subs):
            sl = slice(c * m, (c + 1) * m)
            s.act1.copy_(pa1[sl], non_blocking=True); s.act2.copy_(pa2[sl], non_blocking=True)
            o, r, d, i = s.step(s.act1, s.act2)
            evs[c].record(main)
            side.wait_event(evs[c])
            with torch.cuda.stream(side):
                hob[sl].copy_(o, non_blocking=True); hrw[sl].copy_(r, non_blocking=True)
                hdn[sl].copy_(d, non_blocking=True); hin[sl].copy_(i["flags"], non_blocking=True)
        side.synchronize()
    print(f"chunks={C}:", timed(step))
