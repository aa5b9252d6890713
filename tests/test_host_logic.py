import os
"""Host-side logic that needs no GPU: spaces, sharding arithmetic, statistics decoding, lazy info,
and that the product refuses to run without CUDA (no CPU fallback)."""
import numpy as np
import pytest
import torch

import merging_gym_b200 as mg
from merging_gym_b200 import _native as nat
from merging_gym_b200.sharding import shard_range, stats_to_dict
from merging_gym_b200.spaces import merge_action_space, merge_observation_space


def test_spaces_match_reference():
    """merging_env.py:75-78, 101-102."""
    a = merge_action_space()
    assert a.n == 5 and all(0 <= a.sample() < 5 for _ in range(50)) and 4 in a and 5 not in a
    o = merge_observation_space()
    assert o.shape == (10,) and o.dtype == np.float16
    assert list(o.low) == [-1000, -300, -100, 0, 0] * 2 and list(o.high) == [1000, 300, 100, 1000, 100] * 2
    assert o.contains([0, -30.06, 0, 900, 20, 0, 30.06, 0, 900, 20])


@pytest.mark.parametrize("total,world", [(2 ** 23, 8), (2 ** 23, 4), (10, 4), (7, 8), (0, 2), (1000003, 3)])
def test_shard_range_partitions(total, world):
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == total
    for (b0, c0), (b1, _) in zip(spans, spans[1:]):
        assert b0 + c0 == b1
    assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(total, world, world)


def test_stats_to_dict():
    t = np.zeros(16, np.int64)
    t[:10] = [10, 4, 3, 2, 1, 5, 2100, 0, int(-12.5 * 2 ** 24), int(3.25 * 2 ** 24)]
    d = stats_to_dict(t, 2.0 ** 24)
    assert d["episodes"] == 10 and d["collision_rate"] == 0.4 and d["merge_success_rate"] == 0.5
    assert d["mean_length"] == 210 and d["sum_return1"] == -12.5 and d["mean_return2"] == 0.325
    assert stats_to_dict(np.zeros(16, np.int64), 2.0 ** 24)["collision_rate"] == 0.0


def test_step_info_lazy_decoding():
    flags = torch.tensor([0, nat.INFO_COLLISION | nat.INFO_DONE, (2 << 1) | nat.INFO_DONE,
                          nat.INFO_TIMEOUT | nat.INFO_DONE | (1 << 1), nat.INFO_BAD_ACTION], dtype=torch.uint8)
    info = mg.StepInfo(flags, {"episode_length": torch.arange(5)})
    assert info["collision"].tolist() == [False, True, False, False, False]
    assert info["winner"].tolist() == [0, 0, 2, 1, 0]
    assert info["timeout"].tolist() == [False, False, False, True, False]
    assert info["bad_action"].tolist() == [False] * 4 + [True]
    assert set(info) == {"collision", "winner", "timeout", "bad_action", "flags", "episode_length"}
    assert len(info) == 6 and "collision" in info


def test_opponent_view():
    obs = torch.arange(20.0).reshape(2, 10)
    v = mg.MergeVecEnv.opponent_view(obs)
    assert v[0].tolist() == [5, 6, 7, 8, 9, 0, 1, 2, 3, 4]     # state[5:] + state[:5], main.py:199


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(mg.NativeError, match="no CPU fallback"):
        mg.MergeVecEnv(8)
    with pytest.raises(mg.NativeError):
        mg.make("merging_env-v0")
    with pytest.raises(KeyError):
        mg.make("CartPole-v0")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; the package must not reference it."""
    import os
    pkg = os.path.dirname(mg.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "merge_oracle" not in src or f.endswith((".cuh", ".cu")), f


def test_exploration_rule_constants():
    """`np.random.randn() <= 0.7` (main.py:103) holds with probability Phi(0.7); the device compares a u32 with floor(p * 2^32)."""
    import math
    from merging_gym_b200.policy import EPISILO, Exploration
    ex = Exploration()
    assert ex.threshold == EPISILO == 0.7
    assert abs(ex.keep_prob - 0.5 * (1 + math.erf(0.7 / math.sqrt(2)))) < 1e-15 and abs(ex.keep_prob - 0.7580363) < 1e-7
    assert ex.keep_u32 == int(ex.keep_prob * 2 ** 32)
    assert Exploration(threshold=10.0).keep_u32 == 0xFFFFFFFF and Exploration(threshold=-10.0).keep_u32 == 0
    sp = Exploration(seed=5, step=9).spec()
    assert (sp.seed, sp.step, sp.keep_u32) == (5, 9, ex.keep_u32)


def test_nvtx_ranges_are_off_by_default():
    import contextlib
    from merging_gym_b200 import tracing
    assert not tracing.enabled() or os.environ.get("MG_NVTX", "0") not in ("", "0")
    if not tracing.enabled():
        assert isinstance(tracing.nvtx_range("x"), contextlib.nullcontext)
    tracing.enable(True)
    assert tracing.enabled() and tracing.nvtx_range("x").name == "x"
    tracing.enable(False)


def test_f16x3_operand_blob_round_trips_the_weights():
    """`MLPPolicy(backend="f16x3")` packs fc1 (+ its bias in K slot 15) and fc2 as fp16 hi / lo operands, each scaled by a
    power of two, behind the two constants the kernel multiplies back with (include/merging_b200.h, MG_MLP_FLAG_F16X3).
    Unpacked on the CPU, hi + lo must give the weights back to 2^-22 of the largest entry, the scaled operands must sit in
    fp16's comfortable range, and the padding must be zero.  (Built on the CPU: packing needs no GPU.)"""
    import torch
    from merging_gym_b200.policy import MLPPolicy
    for in_dim, out_dim, seed in ((10, 5, 3), (11, 3, 4)):
        p = MLPPolicy(in_dim, out_dim, device="cpu", seed=seed, backend="f16x3")
        blob = p.w2_f16
        assert blob.dtype == torch.uint8 and blob.numel() == 64 + 13312 + 93184
        c1, c2 = (float(v) for v in blob[:64].view(torch.float32)[:2])
        s1, s2 = -3 - np.log2(c1), 3 - np.log2(c2)
        assert s1 == round(s1) and s2 == round(s2)                          # exact powers of two
        op1 = blob[64:64 + 13312].view(torch.float16).view(52, 2, 8, 8).permute(0, 2, 1, 3).reshape(416, 16).float()
        hi1, lo1 = op1[:208], op1[208:]
        assert 256 <= float(hi1.abs().max()) <= 512 and float(lo1.abs().max()) <= 0.26
        w1op = (hi1 + lo1) / 2.0 ** s1
        tol1 = 2.0 ** -22 * max(float(p.w1.abs().max()), float(p.b1.abs().max()))
        assert float((w1op[:200, :in_dim] - p.w1).abs().max()) <= tol1 and float((w1op[:200, 15] - p.b1).abs().max()) <= tol1
        assert float(w1op[200:].abs().max()) == 0 and float(w1op[:, in_dim:15].abs().max()) == 0
        op2 = blob[64 + 13312:].view(torch.float16).view(13, 28, 2, 8, 8).permute(1, 3, 0, 2, 4).reshape(224, 208).float()
        hi2, lo2 = op2[:112], op2[112:]
        assert 256 <= float(hi2.abs().max()) <= 512
        w2 = (hi2 + lo2) / 2.0 ** s2
        assert float((w2[:100, :200] - p.w2).abs().max()) <= 2.0 ** -22 * float(p.w2.abs().max())
        assert float(w2[100:].abs().max()) == 0 and float(w2[:, 200:].abs().max()) == 0
