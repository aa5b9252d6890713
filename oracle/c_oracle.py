"""TEST INFRASTRUCTURE — ctypes binding of `oracle/merge_oracle.c` (checker / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmerge_oracle.so")
ST_NAMES = ["episodes", "collisions", "wins_p1", "wins_p2", "timeouts", "merges_ok",
            "sum_length", "bad_actions"]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "merge_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


class _State(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "time_stamp", "steps", "winner", "done")]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mgo_step.argtypes = [C.POINTER(_State), C.c_int64] + [C.c_void_p] * 2 + [C.c_int] + \
            [C.c_void_p] * 9 + [C.c_int]
        _lib.mgo_step.restype = None
        _lib.mgo_reset.argtypes = [C.POINTER(_State), C.c_int64, C.c_void_p, C.c_void_p]
        _lib.mgo_reset.restype = None
        _lib.mgo_philox_actions.argtypes = [C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64,
                                            C.c_void_p, C.c_void_p]
        _lib.mgo_philox_actions.restype = None
        _lib.mgo_check_div.argtypes = [C.c_void_p, C.c_int64, C.c_double]
        _lib.mgo_check_div.restype = C.c_int64
        _lib.mgo_atan2_h_r.restype = C.c_double
        _lib.mgo_stats_len.restype = C.c_int
    return _lib


class CVecEnv:
    """Same surface as `merge_oracle.RefVecEnv`, backed by the C restatement."""

    def __init__(self, num_envs, pvp=True, auto_reset=True, nthreads=1):
        n = self.n = int(num_envs)
        self.pvp, self.auto_reset, self.nthreads = bool(pvp), bool(auto_reset), int(nthreads)
        self.pos1 = np.empty(n); self.vel1 = np.empty(n); self.pos2 = np.empty(n); self.vel2 = np.empty(n)
        self.ret1 = np.zeros(n); self.ret2 = np.zeros(n); self.time_stamp = np.zeros(n)
        self.steps = np.zeros(n, np.int32); self.winner = np.zeros(n, np.uint8); self.done = np.zeros(n, np.uint8)
        self._st = _State(*[_p(getattr(self, k)) for k, _ in _State._fields_])
        self.obs = np.zeros((n, 10)); self.rew = np.zeros((n, 2))
        self.done_out = np.zeros(n, np.uint8); self.info = np.zeros(n, np.uint8)
        self.terminal_obs = np.zeros((n, 10)); self.ep_ret = np.zeros((n, 2)); self.ep_len = np.zeros(n, np.int32)
        self._stats = np.zeros(lib().mgo_stats_len(), np.int64); self._sumret = np.zeros(2)
        self.reset()

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().mgo_reset(C.byref(self._st), self.n, _p(m), _p(self.obs))
        return self.obs

    def step(self, a1, a2=None):
        a1 = np.ascontiguousarray(a1, dtype=np.uint8)
        a2 = np.ascontiguousarray(a2, dtype=np.uint8) if self.pvp else None
        lib().mgo_step(C.byref(self._st), self.n, _p(a1), _p(a2), int(self.auto_reset),
                       _p(self.obs), _p(self.rew), _p(self.done_out), _p(self.info),
                       _p(self.terminal_obs), _p(self.ep_ret), _p(self.ep_len),
                       _p(self._stats), _p(self._sumret), self.nthreads)
        return self.obs, self.rew, self.done_out.astype(bool), self.info

    @property
    def stats(self):
        d = {k: int(v) for k, v in zip(ST_NAMES, self._stats)}
        d["sum_return1"], d["sum_return2"] = float(self._sumret[0]), float(self._sumret[1])
        return d


def philox_actions(n, seed, env_id_base, step):
    a1 = np.empty(n, np.uint8); a2 = np.empty(n, np.uint8)
    lib().mgo_philox_actions(n, seed, env_id_base, step, _p(a1), _p(a2))
    return a1, a2


def check_div(x, d):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return int(lib().mgo_check_div(_p(x), x.size, float(d)))
