"""`mg_policy_step`: policy forward + arg-max + exploration + MergeEnv.step in ONE launch.

The fused launch must equal the two-launch sequence it replaces — `MLPPolicy.act(obs)` then `MergeVecEnv.step` —
bit for bit (same MLP arithmetic, same env arithmetic), for both backends, pve and pvp, ragged sizes, random starts,
the h-DQN controller input, and must reproduce the episodes recorded with the reference's shipped checkpoints in the
unmodified reference env.  The device exploration rule is checked against its distribution, for determinism, for
shard invariance and for fresh draws under CUDA-graph replay."""
import math
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
    sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/") and "traj" not in k and "result" not in k}
    return sd, z


def state_of(env):
    return [getattr(env, k).clone() for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta")]


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
@pytest.mark.parametrize("mode,n,reset_mode", [("pve", 5000, "random"), ("pvp", 4099, "fixed"), ("pve", 130, "fixed"),
                                               ("pvp", 1, "random"), ("pve", 66000, "random")])
def test_policy_step_equals_act_then_step(mg, backend, mode, n, reset_mode):
    sd, _ = shipped()
    pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
    kw = dict(mode=mode, seed=5, reset_mode=reset_mode, out_slots=2)
    ea, eb = mg.MergeVecEnv(n, **kw), mg.MergeVecEnv(n, **kw)
    for e in (ea, eb):
        e.rollout(150)                                  # mid-episode, de-synchronised; finishes and resets follow
    taken = torch.empty(n, dtype=torch.uint8, device="cuda")
    qa, qb = torch.empty(n, 5, device="cuda"), torch.empty(n, 5, device="cuda")
    for t in range(120):
        a2 = ea.sample_actions(1000 + t)[1].clone() if mode == "pvp" else None
        obs_a = ea.obs_buf[ea._slot]
        act = pol.act(obs_a, q_out=qa)
        oa = ea.step(act, a2)
        ob = eb.policy_step(pol, a2=a2, actions_out=taken, q_out=qb)
        assert torch.equal(taken, act), f"actions differ at step {t}"
        assert torch.equal(qa, qb), f"Q-values differ at step {t}"
        for x, y, name in zip(oa[:3], ob[:3], ("obs", "rew", "done")):
            assert torch.equal(x, y), f"{name} differs at step {t}"
        assert torch.equal(oa[3]["flags"], ob[3]["flags"]), f"info differs at step {t}"
    for x, y in zip(state_of(ea), state_of(eb)):
        assert torch.equal(x, y)
    assert torch.equal(ea.stats_tensor(), eb.stats_tensor())
    assert ea.stats()["episodes"] > 0
    for k in ("terminal_observation", "episode_return", "episode_length"):
        assert torch.equal(ea._extras[k], eb._extras[k])


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
def test_policy_step_in_place_without_returns_and_sticky_done(mg, backend):
    """out_slots=1 (the next observation overwrites the rows the policy read), track_returns=False, auto_reset=False."""
    sd, _ = shipped()
    pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
    n = 3000
    kw = dict(mode="pve", seed=9, out_slots=1, track_returns=False, auto_reset=False, episode_info=False)
    ea, eb = mg.MergeVecEnv(n, **kw), mg.MergeVecEnv(n, **kw)
    for t in range(260):                                # the greedy episode ends after 225 steps: sticky done follows
        oa = ea.step(pol.act(ea.obs_buf[0]), None)
        ob = eb.policy_step(pol)
        assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1]) and torch.equal(oa[3]["flags"], ob[3]["flags"])
    assert bool(ob[2].all())
    assert torch.equal(ea.meta, eb.meta) and torch.equal(ea.pos1, eb.pos1)
    assert torch.equal(ea.stats_tensor(), eb.stats_tensor())


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
def test_policy_step_hdqn_controller_input(mg, backend):
    """`[goal] + state` (hdqn.py:291): HDQNPolicy.step = goal launch + fused controller/env launch."""
    n = 2100
    h1 = mg.HDQNPolicy(seed=4, backend=backend)
    h2 = mg.HDQNPolicy(meta_state=h1.meta.state_dict(), ctrl_state=h1.ctrl.state_dict(), backend=backend)
    ea, eb = (mg.MergeVecEnv(n, mode="pve", seed=2, reset_mode="random", out_slots=2) for _ in range(2))
    for e in (ea, eb):
        e.rollout(60)
    for t in range(80):
        oa = ea.step(h1.act(ea.obs_buf[ea._slot]), None)
        ob = h2.step(eb)
        assert torch.equal(h1.goal, h2.goal)
        assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1]) and torch.equal(oa[3]["flags"], ob[3]["flags"])
    assert torch.equal(ea.pos2, eb.pos2)


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
def test_policy_step_reproduces_reference_episode(mg, backend):
    """The shipped DQN checkpoint played greedily against the L0 opponent through the fused launch reproduces the
    episode recorded in the unmodified reference env (tests/golden/dqn_policies.npz) action for action."""
    from conftest import rel_err
    sd, z = shipped()
    ref_actions, ref_obs, result = z["L1_1445/traj_actions"], z["L1_1445/traj_obs"], z["L1_1445/result"]
    pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
    env = mg.MergeVecEnv(64, mode="pve", auto_reset=False, out_slots=2)
    taken = torch.empty(64, dtype=torch.uint8, device="cuda")
    obs = env.obs_buf[env._slot]
    T = len(ref_actions)
    for t in range(T):
        assert rel_err(obs[0].cpu().numpy(), ref_obs[t]).max() <= 1e-5, t
        obs, rew, done, info = env.policy_step(pol, actions_out=taken)
        assert taken.cpu().tolist() == [int(ref_actions[t])] * 64, f"action differs at step {t}"
    steps, winner, col, R1, R2 = result
    assert T == steps == 225 and bool(done.all()) and not bool(info["collision"].any())
    assert env.winner.cpu().tolist() == [1] * 64
    assert abs(float(env.ret1[0]) - R1) <= 1e-9 and abs(float(env.ret2[0]) - R2) <= 1e-9


def test_device_exploration_rule(mg):
    """`randn() <= 0.7 ? greedy : randint(0, 5)` (main.py:103-110) as Philox draws: keep rate Phi(0.7), uniform random
    actions, deterministic, independent of sharding, and fresh under identical launch parameters as the env advances."""
    n = 1 << 18
    ex = mg.Exploration(seed=11)
    assert abs(ex.keep_prob - 0.7580363) < 1e-6
    env = mg.MergeVecEnv(n, mode="pve", seed=1)
    greedy = torch.full((n,), 7, dtype=torch.uint8, device="cuda")          # 7 = "kept" marker outside 0..4
    a = ex.apply(greedy.clone(), 5, env)
    kept = (a == 7).float().mean().item()
    assert abs(kept - ex.keep_prob) < 4 * math.sqrt(0.76 * 0.24 / n)
    rnd = a[a != 7].long()
    counts = torch.bincount(rnd, minlength=5).float() / rnd.numel()
    assert counts.numel() == 5 and (counts - 0.2).abs().max().item() < 0.01
    assert torch.equal(a, ex.apply(greedy.clone(), 5, env))                 # same (seed, env ids, clocks) -> same draws
    # shard invariance: the second half as its own shard
    half = mg.MergeVecEnv(n // 2, mode="pve", seed=1, env_id_base=n // 2)
    b = ex.apply(greedy[: n // 2].clone(), 5, half)
    assert torch.equal(b, a[n // 2:])
    # the env clock moves -> new draws with the very same parameter block
    env.rollout(1)
    c = ex.apply(greedy.clone(), 5, env)
    assert (c != a).float().mean().item() > 0.2
    # another salt (the h-DQN goal stream) is a different stream
    d = ex.apply(greedy.clone(), 5, env, salt=1)
    assert (d != c).float().mean().item() > 0.2


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
def test_fused_exploration_equals_unfused(mg, backend):
    """policy_step(explore=...) == act -> mg_explore -> step, and the actions it reports are the ones it took."""
    sd, _ = shipped()
    pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
    n = 6000
    ex = mg.Exploration(seed=3)
    ea, eb = (mg.MergeVecEnv(n, mode="pve", seed=8, reset_mode="random", out_slots=2) for _ in range(2))
    taken = torch.empty(n, dtype=torch.uint8, device="cuda")
    changed = 0
    for t in range(100):
        g = pol.act(ea.obs_buf[ea._slot])
        a = ex.apply(g.clone(), 5, ea)
        changed += int((a != g).sum())
        oa = ea.step(a, None)
        ob = eb.policy_step(pol, explore=ex, actions_out=taken)
        assert torch.equal(taken, a), f"explored actions differ at step {t}"
        assert torch.equal(oa[0], ob[0]) and torch.equal(oa[3]["flags"], ob[3]["flags"])
    assert 0.15 < changed / (100 * n) < 0.25                             # (1 - Phi(0.7)) * 4/5 = 0.194
    assert torch.equal(ea.pos1, eb.pos1)


@pytest.mark.parametrize("backend,hdqn", [("fused", False), ("tf32x3", False), ("tf32x3", True), ("f16x3", False), ("f16x3", True)])
def test_graphed_fused_rollout_equals_eager(mg, backend, hdqn):
    """GraphedPolicyRollout(fused=True): K fused launches in one CUDA graph == the eager unfused loop, including the
    replay rows the recorder stores and exploration (fresh draws on every replay)."""
    sd, _ = shipped()
    n, K = 4096, 8
    ex1, ex2 = mg.Exploration(seed=21), mg.Exploration(seed=21)
    if hdqn:
        p1 = mg.HDQNPolicy(seed=6, backend=backend)
        p2 = mg.HDQNPolicy(meta_state=p1.meta.state_dict(), ctrl_state=p1.ctrl.state_dict(), backend=backend)
    else:
        p1 = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
        p2 = p1
    ea = mg.MergeVecEnv(n, mode="pve", seed=4, reset_mode="random", out_slots=2)
    eb = mg.MergeVecEnv(n, mode="pve", seed=4, reset_mode="random", out_slots=2)
    ra, rb = mg.TransitionRecorder(ea, 1 << 16), mg.TransitionRecorder(eb, 1 << 16)
    roll = mg.GraphedPolicyRollout(eb, p2, k_steps=K, after_step=rb.record, warmup_steps=2, fused=True, explore=ex2)

    def eager_step():
        obs = ea.obs_buf[ea._slot]
        if hdqn:
            p1.meta.act(obs, out=p1_goal)
            ex1.apply(p1_goal, 3, ea, salt=1)
            a = p1.ctrl.act(obs, goal=p1_goal)
        else:
            a = p1.act(obs)
        a = ex1.apply(a, 5, ea)
        out = ea.step(a, None)
        ra.record(obs, a, None, out)
    p1_goal = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eager_step()
    for rep in range(3):
        for _ in range(K):
            eager_step()
        roll.run()
        assert torch.equal(ea.pos1, eb.pos1), f"state differs after replay {rep}"
        assert torch.equal(ea.obs_buf[ea._slot], eb.obs_buf[eb._slot])
    torch.cuda.synchronize()
    assert int(ra.counter) == int(rb.counter) > 0
    assert torch.equal(ra.ring, rb.ring)


def test_policy_step_argument_errors(mg):
    env = mg.MergeVecEnv(64, mode="pve")
    with pytest.raises(ValueError):
        env.policy_step(mg.MLPPolicy(10, 3))                    # 3 outputs are goals, not actions
    with pytest.raises(ValueError):
        env.policy_step(mg.MLPPolicy(11, 5))                    # controller without a goal column
    with pytest.raises(ValueError):
        env.policy_step(mg.MLPPolicy(10, 5, backend="torch"))
    with pytest.raises(ValueError):
        env.policy_step(mg.MLPPolicy(10, 5), a2=torch.zeros(3, dtype=torch.uint8, device="cuda"))
