"""ctypes binding of the C ABI in include/merging_b200.h (libmerging_b200.so).

There is deliberately no fallback: if the shared library is missing or a call fails, the
product raises.  `MERGING_B200_LIB` may point at an alternative build of the library.
"""
from __future__ import annotations

import ctypes as C
import os

from ._paths import LIB_PATH

MG_ABI_VERSION = 7
OBS_DIM = 10
NUM_ACTIONS = 5
STATS_ROWS = 1024
STATS_COLS = 16

ACT_U8, ACT_I32, ACT_I64 = 0, 1, 2
FLAG_AUTO_RESET = 0x1
FLAG_NO_RETURNS = 0x2
POLICY_FLAG_EXPLORE = 0x100
POLICY_FLAG_PDL = 0x200
POLICY_FLAG_GOAL_IN_SLOT = 0x400
MLP_FLAG_MIRROR, MLP_FLAG_PDL = 0x1, 0x2
MLP_FLAG_OBS_SOA, MLP_FLAG_OBS_GOAL_SLOT, MLP_FLAG_WRITE_GOAL = 0x4, 0x8, 0x10
MLP_FLAG_F16X3 = 0x20
FLAG_OBS_SOA, FLAG_OBS_GOAL_SLOT = 0x10, 0x20
OBS_LAYOUTS = ("aos", "soa", "goal_slot")
OBS_LAYOUT_FLAG = {"aos": 0, "soa": FLAG_OBS_SOA, "goal_slot": FLAG_OBS_GOAL_SLOT}
MLP_LAYOUT_FLAG = {"aos": 0, "soa": MLP_FLAG_OBS_SOA, "goal_slot": MLP_FLAG_OBS_GOAL_SLOT}


def soa_stride(n: int) -> int:
    """MG_OBS_SOA_STRIDE: elements per column of the [10][stride] observation layout."""
    return (int(n) + 15) & ~15
POLICY_BACKEND_FP32, POLICY_BACKEND_TF32X3, POLICY_BACKEND_F16X3 = 0, 1, 2
FIELD_OBS, FIELD_REW, FIELD_DONE, FIELD_INFO, FIELD_ALL = 0x1, 0x2, 0x4, 0x8, 0xF
FIELD_BITS = {"obs": FIELD_OBS, "rew": FIELD_REW, "done": FIELD_DONE, "info": FIELD_INFO}

INFO_COLLISION = 0x01
INFO_WINNER_SHIFT = 1
INFO_WINNER_MASK = 0x06
INFO_TIMEOUT = 0x08
INFO_DONE = 0x10
INFO_BAD_ACTION = 0x80

META_STEPS_MASK = 0x0FFF
META_WINNER_SHIFT = 12
META_DONE = 0x4000
META_RESETS_SHIFT = 15

RESET_FIXED, RESET_RANDOM = 0, 1

STAT_NAMES = ["episodes", "collisions", "wins_p1", "wins_p2", "timeouts", "merges_ok",
              "sum_length", "bad_actions", "sum_return1_fx", "sum_return2_fx"]


class MgState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta")]


class MgOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("obs", "rew", "done", "info", "term_obs", "ep_ret", "ep_len")]


class MgHostSlot(C.Structure):
    _fields_ = [("h_a1", C.c_void_p), ("h_a2", C.c_void_p), ("d_a1", C.c_void_p), ("d_a2", C.c_void_p),
                ("d_out", MgOut), ("h_out", MgOut),
                ("ev_uploaded", C.c_void_p), ("ev_stepped", C.c_void_p), ("ev_done", C.c_void_p)]


class MgRewards(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("r_first", "r_second", "r_collision", "vel_penalty", "time_penalty")]


class MgResetSpec(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("reserved", C.c_uint32), ("seed", C.c_uint64), ("env_id_base", C.c_uint64)]


class MgExplore(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("step", C.c_uint64), ("keep_u32", C.c_uint32), ("reserved", C.c_uint32)]


class MgConstants(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("R", "H", "W", "dT", "start_point", "end_point",
                                          "prediction_t", "init_vel", "action_dv")] + \
               [(k, C.c_int32) for k in ("vehicle_w", "vehicle_h", "max_steps", "num_actions",
                                         "obs_dim", "stats_rows", "stats_cols")] + \
               [("return_fixed_point_scale", C.c_double)]


class NativeError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return os.environ.get("MERGING_B200_LIB", LIB_PATH)


def load():
    """Load libmerging_b200.so (once).  Raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(
            f"{path} not found: build it with `python -m merging_gym_b200.build` "
            "(nvcc, sm_100a).  merging_gym_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    vp, i64, u64, i32, u32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32
    lib.mg_version.restype = C.c_int
    lib.mg_last_error.restype = C.c_char_p
    lib.mg_get_constants.argtypes = [C.POINTER(MgConstants)]
    lib.mg_default_rewards.argtypes = [C.POINTER(MgRewards)]
    rsp = C.POINTER(MgResetSpec)
    lib.mg_reset.argtypes = [C.POINTER(MgState), i64, vp, vp, u32, rsp, vp]
    lib.mg_step.argtypes = [C.POINTER(MgState), i64, vp, vp, C.c_int, C.POINTER(MgRewards),
                            C.POINTER(MgOut), vp, u32, rsp, vp]
    lib.mg_sample_actions.argtypes = [vp, vp, i64, u64, u64, u64, vp]
    lib.mg_rollout.argtypes = [C.POINTER(MgState), i64, C.c_int, u64, u64, u64, i32,
                               C.POINTER(MgRewards), C.POINTER(MgOut), vp, vp, u32, rsp, vp]
    lib.mg_step_host.argtypes = [C.POINTER(MgState), i64, vp, vp, vp, vp, C.POINTER(MgRewards),
                                 C.POINTER(MgOut), C.POINTER(MgOut), vp, u32, rsp, vp, vp, i32]
    lib.mg_step_host_async.argtypes = [C.POINTER(MgState), i64, C.POINTER(MgHostSlot), u32, C.POINTER(MgRewards), vp,
                                       u32, rsp, vp, vp, vp]
    lib.mg_step_host_wait.argtypes = [vp]
    lib.mg_mlp_act.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]
    lib.mg_mlp_act.restype = C.c_int
    lib.mg_mlp_act_tc.argtypes = lib.mg_mlp_act.argtypes
    lib.mg_mlp_act_tc.restype = C.c_int
    lib.mg_policy_step.argtypes = [C.POINTER(MgState), i64, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(MgRewards),
                                   C.POINTER(MgOut), vp, u32, rsp, C.POINTER(MgExplore), vp, vp, vp]
    lib.mg_policy_step.restype = C.c_int
    lib.mg_explore.argtypes = [vp, i64, i32, C.POINTER(MgExplore), vp, u64, u32, vp]
    lib.mg_explore.restype = C.c_int
    lib.mg_option_update.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.mg_option_update.restype = C.c_int
    lib.mg_record_transitions.argtypes = [vp] * 10 + [i64, i32, i32, i32, vp, i64, vp, vp, vp, vp]
    lib.mg_record_transitions.restype = C.c_int
    lib.mg_record_scratch_words.argtypes = [i64]
    lib.mg_record_scratch_words.restype = i64
    for f in (lib.mg_get_constants, lib.mg_default_rewards, lib.mg_reset, lib.mg_step,
              lib.mg_sample_actions, lib.mg_rollout, lib.mg_step_host, lib.mg_step_host_async,
              lib.mg_step_host_wait):
        f.restype = C.c_int
    if lib.mg_version() != MG_ABI_VERSION:
        raise NativeError(f"ABI mismatch: library {lib.mg_version()} != binding {MG_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mg_last_error().decode(errors="replace")
        raise NativeError(f"{what} failed (code {rc}): {msg}")


def constants() -> MgConstants:
    c = MgConstants()
    check(load().mg_get_constants(C.byref(c)), "mg_get_constants")
    return c


def default_rewards() -> MgRewards:
    r = MgRewards()
    check(load().mg_default_rewards(C.byref(r)), "mg_default_rewards")
    return r
