"""Observation layouts (SURVEY.md §7.4 / §8b `obs_layout`): "soa" [10,S] and "goal_slot" [N,11] (`[goal] + state`,
hdqn.py:291) carry exactly the values of the default [N,10] rows — through reset, step (full warps and ragged tails,
fixed and random starts, auto-reset), the policy kernels' read side on both backends, the fused `policy_step`, and the
h-DQN loop in which the goal network writes slot 0 and the controller reads the row as it is."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import merging_gym_b200
    return merging_gym_b200


def shipped(tag="L1_1445"):
    z = np.load(os.path.join(GOLDEN, "dqn_policies.npz"))
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/") and "traj" not in k and "result" not in k}


@pytest.mark.parametrize("layout", ["soa", "goal_slot"])
@pytest.mark.parametrize("mode,n,reset_mode", [("pvp", 4099, "random"), ("pve", 64, "fixed"), ("pvp", 1, "fixed"), ("pve", 1000, "random")])
def test_step_and_reset_write_the_same_values_in_every_layout(mg, layout, mode, n, reset_mode):
    kw = dict(mode=mode, seed=3, reset_mode=reset_mode, out_slots=2)
    ea, eb = mg.MergeVecEnv(n, **kw), mg.MergeVecEnv(n, obs_layout=layout, **kw)
    assert eb.obs_buf.shape[1:] == ((10, (n + 15) & ~15) if layout == "soa" else (n, 11))
    assert torch.equal(ea.observation(), eb.observation())
    if layout == "goal_slot":
        eb.obs_buf[:, :, 0] = 7.0                                   # the env must never touch the goal slot
    for t in range(260):
        a1, a2 = ea.sample_actions(t)
        oa = ea.step(a1, a2)
        ob = eb.step(a1, a2)
        assert torch.equal(oa[0], eb.observation()), f"obs differs at step {t}"
        assert torch.equal(oa[1], ob[1]) and torch.equal(oa[2], ob[2]) and torch.equal(oa[3]["flags"], ob[3]["flags"])
    assert torch.equal(ea.pos1, eb.pos1) and torch.equal(ea.stats_tensor(), eb.stats_tensor()) and ea.stats()["episodes"] > 0
    mask = (torch.arange(n, device="cuda") % 3 == 0)
    ra = ea.reset(mask); eb.reset(mask)
    assert torch.equal(ra, eb.observation())
    ea.rollout(5); eb.rollout(5)                                    # rollout refreshes the observation in the env's layout
    assert torch.equal(ea.observation(), eb.observation())
    if layout == "goal_slot":
        assert bool((eb.obs_buf[:, :, 0] == 7.0).all())


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
@pytest.mark.parametrize("layout", ["soa", "goal_slot"])
def test_policy_kernels_read_every_layout(mg, backend, layout):
    n = 3001
    sd = shipped()
    ea, eb = (mg.MergeVecEnv(n, mode="pve", seed=5, reset_mode="random", obs_layout=l) for l in ("aos", layout))
    for e in (ea, eb):
        e.rollout(120)
    pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
    qa, qb = torch.empty(n, 5, device="cuda"), torch.empty(n, 5, device="cuda")
    a = pol.act(ea.obs_buf[0], q_out=qa)
    b = pol.act(eb.obs_buf[0], q_out=qb, obs_layout=layout, n=n)
    assert torch.equal(a, b) and torch.equal(qa, qb)
    am = pol.act(ea.obs_buf[0], mirror=True)                        # the opponent's view, swapped while reading
    bm = pol.act(eb.obs_buf[0], mirror=True, obs_layout=layout, n=n)
    assert torch.equal(am, bm)
    ctrl = mg.MLPPolicy(11, 5, seed=2, backend=backend)
    goal = torch.randint(0, 3, (n,), dtype=torch.uint8, device="cuda")
    assert torch.equal(ctrl.act(ea.obs_buf[0], goal=goal), ctrl.act(eb.obs_buf[0], goal=goal, obs_layout=layout, n=n))
    if layout == "goal_slot":                                       # the row's own slot 0 as the goal column
        eb.obs_buf[0, :, 0] = goal.float()
        assert torch.equal(ctrl.act(ea.obs_buf[0], goal=goal), ctrl.act(eb.obs_buf[0], obs_layout=layout))


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
@pytest.mark.parametrize("layout", ["soa", "goal_slot"])
def test_policy_step_in_every_layout(mg, backend, layout):
    n = 2500
    pol = mg.MLPPolicy(10, 5, state_dict=shipped(), backend=backend)
    ea = mg.MergeVecEnv(n, mode="pve", seed=7, reset_mode="random", out_slots=2)
    eb = mg.MergeVecEnv(n, mode="pve", seed=7, reset_mode="random", out_slots=2, obs_layout=layout)
    for e in (ea, eb):
        e.rollout(150)
    for t in range(90):
        oa = ea.policy_step(pol)
        ob = eb.policy_step(pol)
        assert torch.equal(oa[0], eb.observation()), f"obs differs at step {t}"
        assert torch.equal(oa[1], ob[1]) and torch.equal(oa[3]["flags"], ob[3]["flags"])
    assert torch.equal(ea.pos1, eb.pos1) and torch.equal(ea.meta, eb.meta)


@pytest.mark.parametrize("backend,fused_step", [("fused", False), ("fused", True), ("tf32x3", False), ("tf32x3", True), ("f16x3", False), ("f16x3", True)])
def test_hdqn_on_goal_slot_rows(mg, backend, fused_step):
    """`[goal] + state` rows: the goal network stores its choice into slot 0 (MG_MLP_FLAG_WRITE_GOAL) and the controller
    reads the 11-float row as it is == the loop with a separate goal array on the default rows."""
    n = 1500
    ha = mg.HDQNPolicy(seed=9, backend=backend)
    hb = mg.HDQNPolicy(meta_state=ha.meta.state_dict(), ctrl_state=ha.ctrl.state_dict(), backend=backend)
    ea = mg.MergeVecEnv(n, mode="pve", seed=1, reset_mode="random", out_slots=1)
    eb = mg.MergeVecEnv(n, mode="pve", seed=1, reset_mode="random", out_slots=1, obs_layout="goal_slot")
    for e in (ea, eb):
        e.rollout(80)
    for t in range(70):
        oa = ea.step(ha.act(ea.obs_buf[0]), None)
        if fused_step:
            ob = hb.step(eb)
        else:
            ob = eb.step(hb.act(eb.obs_buf[0], obs_layout="goal_slot"), None)
        assert torch.equal(ha.goal, hb.goal), f"goals differ at step {t}"
        assert torch.equal(oa[0], eb.observation()) and torch.equal(oa[1], ob[1])
    assert torch.equal(ea.pos2, eb.pos2)


def test_layout_argument_errors(mg):
    with pytest.raises(ValueError):
        mg.MergeVecEnv(64, obs_layout="rows")
    with pytest.raises(ValueError):
        mg.MergeVecEnv(64, obs_layout="soa", lanes=2)
    env = mg.MergeVecEnv(64, obs_layout="soa")
    with pytest.raises(ValueError):
        env.step_host(np.zeros(64, np.uint8), np.zeros(64, np.uint8))
    with pytest.raises(ValueError):
        env.rollout(2, obs=torch.empty(2, 64, 10, device="cuda"))
    with pytest.raises(ValueError):
        mg.TransitionRecorder(env, 128)
    with pytest.raises(ValueError):
        mg.MLPPolicy(10, 5).act(env.obs_buf[0], obs_layout="soa")            # n is required
    with pytest.raises(mg.NativeError):                                        # int64 actions: default rows only
        env.step(torch.zeros(64, dtype=torch.int64, device="cuda"), torch.zeros(64, dtype=torch.int64, device="cuda"))
