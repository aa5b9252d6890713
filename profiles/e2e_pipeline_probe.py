#!/usr/bin/env python
"""Where the time of one pipelined host-buffer step goes: device-side timestamps of the kernel and of the copy of
each step (CUDA events with timing on the two streams), for the product path (kernel reads pinned host actions) and
for variants: actions already on the device; actions uploaded by an explicit cudaMemcpyAsync on a third stream."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402
from merging_gym_b200 import _native as nat  # noqa: E402

n, T = 1 << 20, 24
dev = torch.device("cuda")
env = mg.MergeVecEnv(n, episode_info=False)
env.rollout(300)
a1d, a2d = (x.clone() for x in env.sample_actions(5))
out_bytes = n * 50
res = {}


def product(T):
    for i in range(T):
        if i >= 2:
            env.step_host_wait()
        h1, h2 = env.host_action_buffers()
        env.step_host_async(h1, h2)
    env.step_host_wait(); env.step_host_wait()


for s in range(2):
    h1, h2 = env.host_action_buffers(s)
    h1[:] = a1d.cpu().numpy(); h2[:] = a2d.cpu().numpy()
product(4)
torch.cuda.synchronize()
t0 = time.perf_counter(); product(T); res["product_ms_per_step"] = (time.perf_counter() - t0) / T * 1e3

# ---- hand-built pipeline with timing events ------------------------------------------------------------------------
main, cs, us = torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()
dblock = [torch.empty(out_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
hblock = [torch.empty(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
hact = [torch.empty(2 * n, dtype=torch.uint8).pin_memory() for _ in range(2)]
dact = [torch.empty(2 * n, dtype=torch.uint8, device=dev) for _ in range(2)]
for h in hact:
    h[:n] = a1d.cpu(); h[n:] = a2d.cpu()


def views(block):
    f = block.view(torch.float32)
    return (f[:n * 10].view(n, 10), f[n * 10:n * 12].view(n, 2), block[n * 48:n * 49], block[n * 49:n * 50])


outs = []
for b in dblock:
    o, r, d, i = views(b)
    outs.append(nat.MgOut(o.data_ptr(), r.data_ptr(), d.data_ptr(), i.data_ptr(), None, None, None))


def run(variant, T):
    def ev():
        return torch.cuda.Event(enable_timing=True)
    marks = []
    done = [None, None]
    for t in range(T):
        k = t % 2
        if done[k] is not None:
            done[k].synchronize()                       # host owns slot k again
        k0, k1, c0, c1 = ev(), ev(), ev(), ev()
        if variant == "upload":
            with torch.cuda.stream(us):
                dact[k].copy_(hact[k], non_blocking=True)
                up = torch.cuda.Event(); up.record(us)
            main.wait_event(up)
        if variant == "host":
            p1, p2 = hact[k].data_ptr(), hact[k].data_ptr() + n
        else:
            p1, p2 = dact[k].data_ptr(), dact[k].data_ptr() + n
        k0.record(main)
        nat.check(env._lib.mg_step(C.byref(env._state), n, C.c_void_p(p1), C.c_void_p(p2), nat.ACT_U8, C.byref(env._rw),
                                   C.byref(outs[k]), None, env._flags(), C.byref(env._rs), C.c_void_p(main.cuda_stream)), "step")
        k1.record(main)
        cs.wait_event(k1)
        with torch.cuda.stream(cs):
            c0.record(cs)
            hblock[k].copy_(dblock[k], non_blocking=True)
            c1.record(cs)
        done[k] = c1
        marks.append((k0, k1, c0, c1))
    torch.cuda.synchronize()
    base = marks[0][0]
    return [[round(base.elapsed_time(e), 3) for e in m] for m in marks]


for variant in ("host", "device", "upload"):
    run(variant, 4)
    t0 = time.perf_counter(); rows = run(variant, T); wall = (time.perf_counter() - t0) / T * 1e3
    kern = [r[1] - r[0] for r in rows[4:]]
    copy = [r[3] - r[2] for r in rows[4:]]
    gap = [rows[i + 1][2] - rows[i][3] for i in range(4, T - 1)]
    period = (rows[-1][3] - rows[4][3]) / (len(rows) - 5)
    res[variant] = {"device_period_ms": period, "wall_ms_per_step": wall, "kernel_ms_median": float(np.median(kern)), "copy_ms_median": float(np.median(copy)),
                    "gap_between_copies_ms_median": float(np.median(gap)), "timeline_first_rows": rows[4:10]}
if "--full" not in sys.argv:
    for v in ("host", "device", "upload"):
        res[v].pop("timeline_first_rows")
# device-side period of the copy stream = what the pipeline sustains, free of this script's host-side bookkeeping
print(json.dumps(res))
