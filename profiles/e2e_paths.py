#!/usr/bin/env python
"""The host-buffer step (`MergeVecEnv.step_host` -> `mg_step_host`) at 2^20 envs: pageable vs pinned action
arrays, 1 / 2 / 4 / 8 pipeline pieces, and the zero-copy variant; plus a plain 52 MB D2H copy for reference."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg

n = 1 << 20
env = mg.MergeVecEnv(n, episode_info=False); env.rollout(300)
a1, a2 = env.sample_actions(); h1 = a1.cpu().numpy().copy(); h2 = a2.cpu().numpy().copy()
p1, p2 = env.host_action_buffers(); p1[:] = h1; p2[:] = h2


def timed(fn, iters=30):
    for _ in range(3): fn()
    t = time.perf_counter()
    for _ in range(iters): fn()
    dt = (time.perf_counter() - t) / iters
    return f"{dt * 1e3:.3f} ms/step  {n / dt:.3e} env-steps/s  {n * 52 / dt / 1e9:.1f} GB/s"


print("pageable actions (staged), 1 piece :", timed(lambda: env.step_host(h1, h2, chunks=1)))
for c in (1, 2, 4, 8, 16):
    print(f"pinned actions, {c:2d} piece(s)        :", timed(lambda: env.step_host(p1, p2, chunks=c)))
print("pinned actions read by the kernel   :", timed(lambda: env.step_host(p1, p2, direct_actions=True)))
print("zero-copy (kernel stores to host)  :", timed(lambda: env.step_host(p1, p2, zero_copy=True)))
dev = torch.empty(52 << 20, dtype=torch.uint8, device="cuda"); host = torch.empty(52 << 20, dtype=torch.uint8).pin_memory()
for _ in range(3): host.copy_(dev, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
print("plain 52 MiB D2H copy              : %.1f GB/s" % (20 * (52 << 20) / (time.perf_counter() - t) / 1e9))
