"""CPU placement of a rank next to its GPU (host-buffer path only).

`mg_step_host*` moves 52 bytes per env-step across PCIe into pinned host memory.  Pinned pages are placed by the
kernel's NUMA policy of the *allocating thread*, so a rank that will talk to GPU g should run on — and allocate its
pinned buffers from — the cores NVML reports as local to g (`nvmlDeviceGetCpuAffinity`).  `bind_to_gpu()` does that
with `os.sched_setaffinity`; call it before the first `step_host` / `host_action_buffers` call (those allocate).
On a single-NUMA-node box (NVML returns every core for every GPU) it changes nothing.
"""
from __future__ import annotations

import os
from typing import List, Optional


def gpu_local_cpus(device_index: int) -> Optional[List[int]]:
    """Cores NVML lists as local to the GPU, intersected with this process's allowed set; None if unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
    except Exception:
        return None
    cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
    try:
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
    except (AttributeError, OSError):
        pass
    return cpus or None


def bind_to_gpu(device_index: int) -> dict:
    """Pin the calling process to the GPU-local cores.  Returns what was done (for logs / bench lines)."""
    cpus = gpu_local_cpus(device_index)
    info = {"device": int(device_index), "gpu_local_cpus": None if cpus is None else len(cpus), "bound": False}
    if cpus is None:
        return info
    try:
        before = os.sched_getaffinity(0)
        if set(cpus) != set(before):
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        info["cpus_before"], info["cpus_after"] = len(before), len(os.sched_getaffinity(0))
    except (AttributeError, OSError) as e:      # not Linux / not permitted: placement stays as it is
        info["error"] = repr(e)
    return info
