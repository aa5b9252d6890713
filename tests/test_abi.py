"""The C-ABI shared library: it loads, exports every symbol include/merging_b200.h declares,
agrees with the Python binding on struct layouts and constants, and rejects bad arguments
without touching a GPU.  (No compute calls here — those are the `-m gpu` tests.)"""
import ctypes as C
import os
import re
import subprocess

import pytest

from merging_gym_b200 import _native as nat
from merging_gym_b200 import build as mgbuild
from merging_gym_b200._paths import INCLUDE_DIR, LIB_PATH
from oracle import merge_oracle as mo

HEADER = os.path.join(INCLUDE_DIR, "merging_b200.h")


@pytest.fixture(scope="module")
def lib():
    mgbuild.build()
    return nat.load()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MG_API\s+[\w\s\*]+?\b(mg_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    assert declared_symbols() == sorted([
        "mg_version", "mg_last_error", "mg_get_constants", "mg_default_rewards", "mg_reset",
        "mg_step", "mg_sample_actions", "mg_rollout", "mg_step_host", "mg_step_host_async", "mg_step_host_wait",
        "mg_mlp_act", "mg_mlp_act_tc", "mg_record_transitions", "mg_record_scratch_words", "mg_policy_step", "mg_explore", "mg_option_update"])


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (mg_\w+)", out))
    assert exported == set(declared_symbols())       # nothing else leaks (visibility=hidden)


def test_record_scratch_words(lib):
    # [0, n/256) block counts, then the 64-bit pre-increment counter behind entry (n+31)/32 (8-byte aligned)
    for n in (0, 1, 31, 32, 33, 255, 256, 257, 4096, (1 << 20) + 1):
        m = (n + 31) // 32
        words = lib.mg_record_scratch_words(n)
        assert words == m + 4 and ((m + 1) & ~1) + 2 <= words and (n + 255) // 256 <= m
    assert lib.mg_record_scratch_words(-5) == 0


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_version_and_constants(lib):
    assert lib.mg_version() == nat.MG_ABI_VERSION
    c = nat.constants()
    assert (c.R, c.H, c.W, c.dT) == (mo.R, mo.H, mo.W, mo.dT)
    assert (c.start_point, c.end_point, c.prediction_t) == (mo.START_POINT, mo.END_POINT, mo.prediction_t)
    assert (c.vehicle_w, c.vehicle_h, c.max_steps) == (mo.VEHICLE_W, mo.VEHICLE_H, 2501)
    assert (c.num_actions, c.obs_dim) == (mo.NUM_ACTIONS, mo.OBS_DIM)
    assert (c.stats_rows, c.stats_cols) == (nat.STATS_ROWS, nat.STATS_COLS)
    assert c.return_fixed_point_scale == 2.0 ** 24
    r = nat.default_rewards()
    assert (r.r_first, r.r_second, r.r_collision, r.vel_penalty, r.time_penalty) == \
        (mo.RFirst, mo.RSecond, mo.RCollision, mo.vel_penalty, mo.time_penalty)


def test_header_macros_match_binding():
    src = open(HEADER).read()
    def macro(name):
        return int(re.search(rf"#define {name} (0x[0-9A-Fa-f]+|\d+)", src).group(1), 0)
    assert macro("MG_ABI_VERSION") == nat.MG_ABI_VERSION
    assert macro("MG_OBS_DIM") == nat.OBS_DIM and macro("MG_NUM_ACTIONS") == nat.NUM_ACTIONS
    assert macro("MG_STATS_ROWS") == nat.STATS_ROWS and macro("MG_STATS_COLS") == nat.STATS_COLS
    for n, v in [("MG_INFO_COLLISION", nat.INFO_COLLISION), ("MG_INFO_TIMEOUT", nat.INFO_TIMEOUT),
                 ("MG_INFO_DONE", nat.INFO_DONE), ("MG_INFO_BAD_ACTION", nat.INFO_BAD_ACTION),
                 ("MG_INFO_WINNER_MASK", nat.INFO_WINNER_MASK), ("MG_META_DONE", nat.META_DONE),
                 ("MG_META_STEPS_MASK", nat.META_STEPS_MASK), ("MG_FLAG_AUTO_RESET", nat.FLAG_AUTO_RESET),
                 ("MG_META_RESETS_SHIFT", nat.META_RESETS_SHIFT), ("MG_FLAG_NO_RETURNS", nat.FLAG_NO_RETURNS),
                 ("MG_FIELD_OBS", nat.FIELD_OBS), ("MG_FIELD_REW", nat.FIELD_REW), ("MG_FIELD_DONE", nat.FIELD_DONE),
                 ("MG_FIELD_INFO", nat.FIELD_INFO), ("MG_FIELD_ALL", nat.FIELD_ALL)]:
        assert macro(n) == v, n
    # and with the oracle's copy of the info bits
    assert (nat.INFO_COLLISION, nat.INFO_TIMEOUT, nat.INFO_DONE, nat.INFO_BAD_ACTION) == \
        (mo.INFO_COLLISION, mo.INFO_TIMEOUT, mo.INFO_DONE, mo.INFO_BAD_ACTION)


def test_struct_sizes():
    assert C.sizeof(nat.MgState) == 7 * 8 and C.sizeof(nat.MgOut) == 7 * 8
    assert C.sizeof(nat.MgRewards) == 5 * 8 and C.sizeof(nat.MgResetSpec) == 24
    assert C.sizeof(nat.MgHostSlot) == 4 * 8 + 2 * 7 * 8 + 3 * 8
    assert C.sizeof(nat.MgConstants) == 9 * 8 + 7 * 4 + 4 + 8      # 4 bytes padding before the double


def test_argument_errors_without_gpu(lib):
    st, out, rw = nat.MgState(), nat.MgOut(), nat.default_rewards()
    assert lib.mg_step(C.byref(st), -1, None, None, 0, C.byref(rw), C.byref(out), None, 0, None, None) == -2
    assert lib.mg_step(C.byref(st), 8, None, None, 0, C.byref(rw), C.byref(out), None, 0x80, None, None) == -4
    assert lib.mg_step(C.byref(st), 8, None, None, 7, C.byref(rw), C.byref(out), None, 0, None, None) == -5
    assert lib.mg_step(C.byref(st), 8, None, None, 0, C.byref(rw), C.byref(out), None, 0, None, None) == -1
    assert b"NULL" in lib.mg_last_error()
    bad_rs = nat.MgResetSpec(7, 0, 0, 0)
    assert lib.mg_step(C.byref(st), 8, None, None, 0, C.byref(rw), C.byref(out), None, 0, C.byref(bad_rs), None) == -4
    st = nat.MgState(*([0x1008] * 7))                                # not 16-byte aligned
    assert lib.mg_step(C.byref(st), 8, None, None, 0, C.byref(rw), C.byref(out), None, 0, None, None) == -3
    assert lib.mg_reset(None, 8, None, None, 0, None, None) == -1
    assert lib.mg_get_constants(None) == -1
    # n == 0 is a no-op that needs no device
    assert lib.mg_step(C.byref(nat.MgState()), 0, None, None, 0, None, C.byref(out), None, 1, None, None) == 0
    assert lib.mg_reset(C.byref(nat.MgState()), 0, None, None, 0, None, None) == 0
    assert lib.mg_rollout(C.byref(nat.MgState()), 0, 1, 0, 0, 0, 4, None, C.byref(out), None, None, 1, None, None) == 0
    assert lib.mg_sample_actions(None, None, 0, 0, 0, 0, None) == 0
    assert lib.mg_mlp_act(None, None, 0, 10, 5, None, None, None, None, None, None, None, None, 0, None) == 0
    assert lib.mg_mlp_act(None, None, 8, 9, 5, None, None, None, None, None, None, None, None, 0, None) == -2
    assert lib.mg_mlp_act(None, None, 8, 10, 5, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.mg_mlp_act(None, None, 8, 10, 5, None, None, None, None, None, None, None, None, 0x10, None) == -4
    assert lib.mg_mlp_act_tc(None, None, 8, 10, 5, None, None, None, None, None, None, None, None, 0x10, None) == -4
    assert lib.mg_record_transitions(*([None] * 10), 8, 0, 3, 1, None, 16, None, None, None, None) == -4
    assert lib.mg_step_host(None, 8, *([None] * 8), 0, None, None, None, 1) == -1
    assert lib.mg_step_host_async(None, 8, None, 0xF, None, None, 0, None, None, None, None) == -1
    slot = nat.MgHostSlot()
    slot.h_a1 = 0x1000
    assert lib.mg_step_host_async(C.byref(nat.MgState()), 8, C.byref(slot), 0xF, None, None, 0, None, None, None, None) == -1  # no copy stream
    assert lib.mg_step_host_async(C.byref(nat.MgState()), 8, C.byref(slot), 0x1F, None, None, 0, None, None, 0x10, None) == -1  # no events
    assert lib.mg_step_host_wait(None) == -1
    # the lean state: MG_FLAG_NO_RETURNS accepts NULL ret1/ret2 but refuses an episode-return output
    lean = nat.MgState(0x1000, 0x1000, 0x1000, 0x1000, None, None, 0x1000)
    o2 = nat.MgOut(0x1000, 0x1000, None, 0x1000, None, 0x1000, None)
    assert lib.mg_step(C.byref(lean), 8, None, None, 0, C.byref(rw), C.byref(o2), None, nat.FLAG_NO_RETURNS, None, None) == -4
    assert lib.mg_step(C.byref(lean), 8, None, None, 0, C.byref(rw), C.byref(o2), None, 0, None, None) == -1      # ret1 NULL without the flag


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("MERGING_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(nat, "_lib", None)
    with pytest.raises(nat.NativeError, match="no CPU or PyTorch fallback"):
        nat.load()


def test_python_constants_match_the_header():
    """Every flag / enum value the ctypes binding hard-codes equals the header's #define (the header is the contract)."""
    src = open(HEADER).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(MG_\w+)\s+(0x[0-9a-fA-F]+|\d+)u?\b", src)}
    pairs = {"MG_ABI_VERSION": nat.MG_ABI_VERSION, "MG_OBS_DIM": nat.OBS_DIM, "MG_NUM_ACTIONS": nat.NUM_ACTIONS,
             "MG_FLAG_AUTO_RESET": nat.FLAG_AUTO_RESET, "MG_FLAG_NO_RETURNS": nat.FLAG_NO_RETURNS,
             "MG_FLAG_OBS_SOA": nat.FLAG_OBS_SOA, "MG_FLAG_OBS_GOAL_SLOT": nat.FLAG_OBS_GOAL_SLOT,
             "MG_POLICY_FLAG_EXPLORE": nat.POLICY_FLAG_EXPLORE, "MG_POLICY_FLAG_PDL": nat.POLICY_FLAG_PDL,
             "MG_POLICY_FLAG_GOAL_IN_SLOT": nat.POLICY_FLAG_GOAL_IN_SLOT,
             "MG_POLICY_BACKEND_FP32": nat.POLICY_BACKEND_FP32, "MG_POLICY_BACKEND_TF32X3": nat.POLICY_BACKEND_TF32X3,
             "MG_MLP_FLAG_MIRROR": nat.MLP_FLAG_MIRROR, "MG_MLP_FLAG_PDL": nat.MLP_FLAG_PDL,
             "MG_MLP_FLAG_OBS_SOA": nat.MLP_FLAG_OBS_SOA, "MG_MLP_FLAG_OBS_GOAL_SLOT": nat.MLP_FLAG_OBS_GOAL_SLOT,
             "MG_MLP_FLAG_WRITE_GOAL": nat.MLP_FLAG_WRITE_GOAL, "MG_MLP_FLAG_F16X3": nat.MLP_FLAG_F16X3,
             "MG_FIELD_OBS": nat.FIELD_OBS, "MG_FIELD_REW": nat.FIELD_REW, "MG_FIELD_DONE": nat.FIELD_DONE,
             "MG_FIELD_INFO": nat.FIELD_INFO, "MG_FIELD_ALL": nat.FIELD_ALL,
             "MG_INFO_COLLISION": nat.INFO_COLLISION, "MG_INFO_WINNER_SHIFT": nat.INFO_WINNER_SHIFT,
             "MG_INFO_WINNER_MASK": nat.INFO_WINNER_MASK, "MG_INFO_TIMEOUT": nat.INFO_TIMEOUT, "MG_INFO_DONE": nat.INFO_DONE,
             "MG_INFO_BAD_ACTION": nat.INFO_BAD_ACTION, "MG_META_STEPS_MASK": nat.META_STEPS_MASK,
             "MG_META_WINNER_SHIFT": nat.META_WINNER_SHIFT, "MG_META_DONE": nat.META_DONE,
             "MG_META_RESETS_SHIFT": nat.META_RESETS_SHIFT, "MG_RESET_FIXED": nat.RESET_FIXED, "MG_RESET_RANDOM": nat.RESET_RANDOM,
             "MG_STATS_ROWS": nat.STATS_ROWS, "MG_STATS_COLS": nat.STATS_COLS}
    for name, value in pairs.items():
        assert name in defs, f"{name} is not #defined in the header"
        assert defs[name] == value, f"{name}: header {defs[name]} != binding {value}"
    assert nat.soa_stride(1) == 16 and nat.soa_stride(16) == 16 and nat.soa_stride(4099) == 4112
    assert C.sizeof(nat.MgExplore) == 24 and C.sizeof(nat.MgHostSlot) == 4 * 8 + 2 * 56 + 3 * 8


def test_policy_step_and_layout_argument_errors_without_a_gpu(lib):
    """Argument validation happens before any CUDA call."""
    st, out, rw = nat.MgState(), nat.MgOut(), nat.default_rewards()
    vp = C.c_void_p
    z = [None] * 6
    assert lib.mg_policy_step(C.byref(st), -1, None, None, 0, *z, None, C.byref(rw), C.byref(out), None, 0, None, None, None, None, None) == -2
    assert lib.mg_policy_step(C.byref(st), 8, None, None, 7, *z, None, C.byref(rw), C.byref(out), None, 0, None, None, None, None, None) == -4
    assert lib.mg_policy_step(C.byref(st), 8, None, None, 0, *z, None, C.byref(rw), C.byref(out), None, nat.POLICY_FLAG_EXPLORE,
                              None, None, None, None, None) == -1          # MG_POLICY_FLAG_EXPLORE without an MgExplore
    assert lib.mg_policy_step(C.byref(st), 8, None, None, 0, *z, None, C.byref(rw), C.byref(out), None,
                              nat.FLAG_OBS_SOA | nat.FLAG_OBS_GOAL_SLOT, None, None, None, None, None) == -4
    assert lib.mg_policy_step(C.byref(st), 0, None, None, 0, *z, None, C.byref(rw), C.byref(out), None, 0, None, None, None, None, None) == 0
    assert lib.mg_reset(C.byref(st), 8, None, None, nat.FLAG_AUTO_RESET, None, None) == -4       # not a layout flag
    assert lib.mg_mlp_act(vp(16), None, 8, 10, 5, *[vp(16)] * 6, vp(16), None, nat.MLP_FLAG_WRITE_GOAL, None) == -4   # needs the goal-slot layout
    assert lib.mg_mlp_act(vp(16), None, 8, 11, 5, *[vp(16)] * 6, vp(16), None, 0, None) == -2                         # 11-float rows need the layout flag
    # MG_MLP_FLAG_F16X3: a flag of mg_mlp_act_tc only (w1t / b1 may then be NULL: they travel inside the operand blob)
    assert lib.mg_mlp_act(vp(16), None, 8, 10, 5, *[vp(16)] * 6, vp(16), None, nat.MLP_FLAG_F16X3, None) == -4
    assert lib.mg_mlp_act_tc(vp(16), None, 8, 10, 5, None, None, None, vp(16), vp(16), vp(16), vp(16), None, nat.MLP_FLAG_F16X3, None) == -1
    assert lib.mg_mlp_act_tc(vp(16), None, 8, 10, 5, None, None, vp(24), vp(16), vp(16), vp(16), vp(16), None, nat.MLP_FLAG_F16X3, None) == -3
    assert lib.mg_mlp_act_tc(vp(16), None, 8, 10, 5, None, None, vp(16), vp(16), vp(16), vp(16), vp(16), None, 0, None) == -1
    assert lib.mg_explore(None, 8, 5, None, None, 0, 0, None) == -1 and lib.mg_explore(None, 0, 5, None, None, 0, 0, None) == 0
    assert lib.mg_option_update(None, None, None, None, None, 8, None, None, None, None, None) == -1
    assert b"NULL" in lib.mg_last_error()
