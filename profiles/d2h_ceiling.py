#!/usr/bin/env python
"""What the box can move device -> host when N GPUs copy at once (the ceiling of `bench.py`'s `e2e` at N GPUs).

    python profiles/d2h_ceiling.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/d2h_ceiling.py [--out gpurun_out/d2h_ceiling_nN.json]

Every rank owns one GPU and copies `--mb` (default 50 MiB = one 2^20-env step's obs|rew|done|info) from device memory
into its own pinned host buffer with ONE plain `cudaMemcpyAsync` per copy (libcudart through ctypes; nothing of the
product is involved), all ranks starting together behind a barrier:
  sync_each     copy, cudaStreamSynchronize, repeat   — what the synchronous `step_host` does every step
  back_to_back  all copies queued, one synchronize    — what the pipelined `step_host_async` can reach
each for two kinds of pinned memory: `cudaHostAlloc` (torch pin_memory, what the product uses) and `cudaHostRegister`
of a first-touched NumPy buffer, allocated after the rank has been bound to its GPU's local cores
(`merging_gym_b200.affinity.bind_to_gpu`; `--no-bind` to skip).  Also times the 2 MiB host -> device action upload.
Rank 0 prints one JSON line: per-rank GB/s (min / max) and the aggregate = N x bytes x iters / max-over-ranks time.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from merging_gym_b200.affinity import bind_to_gpu  # noqa: E402


def cudart():
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    raise RuntimeError("libcudart not found (import torch first: it ships one)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=50.0)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--no-bind", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bind = {"bound": False, "skipped": True} if args.no_bind else bind_to_gpu(local)
    rt = cudart()
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
    rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    D2H, H2D = 2, 1
    nbytes = int(args.mb * (1 << 20))
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda").fill_(7)
    stream = torch.cuda.Stream()
    sp = C.c_void_p(stream.cuda_stream)

    pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    raw = np.empty(nbytes + 4096, dtype=np.uint8)
    raw[:] = 1                                               # first touch on this (bound) thread
    off = (-raw.ctypes.data) % 4096
    reg_ptr = raw.ctypes.data + off
    rc = rt.cudaHostRegister(C.c_void_p(reg_ptr), nbytes, 0)
    hosts = {"cudaHostAlloc": pinned.data_ptr()}
    if rc == 0:
        hosts["cudaHostRegister_first_touch"] = reg_ptr

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        fn(3)                                                # warm-up
        barrier()
        t0 = time.perf_counter()
        fn(args.iters)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return float(lo.item()), float(hi.item())

    def copier(dst, src, n, kind, sync_each):
        def run(k):
            for _ in range(k):
                assert rt.cudaMemcpyAsync(C.c_void_p(dst), C.c_void_p(src), n, kind, sp) == 0
                if sync_each:
                    assert rt.cudaStreamSynchronize(sp) == 0
            assert rt.cudaStreamSynchronize(sp) == 0
        return run

    res = {}
    for hname, hp in hosts.items():
        for mode, se in (("sync_each", True), ("back_to_back", False)):
            lo, hi = timed(copier(hp, dev.data_ptr(), nbytes, D2H, se))
            res[f"d2h_{hname}_{mode}"] = {"per_rank_gbs_min": nbytes * args.iters / hi / 1e9,
                                          "per_rank_gbs_max": nbytes * args.iters / lo / 1e9,
                                          "aggregate_gbs": world * nbytes * args.iters / hi / 1e9}
    up = 2 << 20
    lo, hi = timed(copier(dev.data_ptr(), pinned.data_ptr(), up, H2D, True))
    res["h2d_2MiB_sync_each"] = {"per_rank_gbs_min": up * args.iters / hi / 1e9, "us_per_copy_max": hi / args.iters * 1e6}
    # both directions at once: the D2H stream keeps copying while a second stream uploads 2 MiB action blocks
    s2 = torch.cuda.Stream()
    sp2 = C.c_void_p(s2.cuda_stream)

    def duplex(k):
        for _ in range(k):
            assert rt.cudaMemcpyAsync(C.c_void_p(pinned.data_ptr()), C.c_void_p(dev.data_ptr()), nbytes, D2H, sp) == 0
            assert rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + (nbytes // 2)), C.c_void_p(reg_ptr if rc == 0 else pinned.data_ptr()),
                                      up, H2D, sp2) == 0
        assert rt.cudaStreamSynchronize(sp) == 0 and rt.cudaStreamSynchronize(sp2) == 0
    lo, hi = timed(duplex)
    res["d2h_with_2MiB_h2d_alongside"] = {"aggregate_gbs_d2h": world * nbytes * args.iters / hi / 1e9}

    if rank == 0:
        line = {"what": "concurrent device->host copy ceiling", "n_gpus": world, "mb_per_copy": args.mb, "iters": args.iters,
                "host_cores_visible": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(), "cpu_binding_rank0": bind,
                "gpu": torch.cuda.get_device_name(0), "results": res}
        txt = json.dumps(line)
        print(txt, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(txt + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
