"""NVTX ranges around the env's launches (tracing hook).

Off by default — an NVTX push/pop costs a few hundred nanoseconds of host time per call, which matters for a launch
that takes 25 us.  `enable()` (or `MG_NVTX=1` in the environment) turns them on; an Nsight Systems / `ncu --nvtx`
timeline then shows `mg.reset`, `mg.step`, `mg.rollout[k]`, `mg.policy_step`, `mg.step_host`, `mg.record` ranges
on the calling thread, each covering the launches of one call.
"""
from __future__ import annotations

import contextlib
import os

_enabled = os.environ.get("MG_NVTX", "0") not in ("", "0")
_null = contextlib.nullcontext()


def enable(on: bool = True) -> None:
    global _enabled
    _enabled = bool(on)


def enabled() -> bool:
    return _enabled


class _Range:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        import torch
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        import torch
        torch.cuda.nvtx.range_pop()
        return False


def nvtx_range(name: str):
    """Context manager: an NVTX range when tracing is enabled, nothing otherwise."""
    return _Range(name) if _enabled else _null
