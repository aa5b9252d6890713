"""Parity of the CUDA path (through the C ABI, via MergeVecEnv/ctypes) with the CPU oracle.

Bar (BASELINE.json north_star): done / collision / winner / timeout flags and step counts
bit-exact; observations, rewards and returns within 1e-5 relative, in the form
|d| <= 1e-5 * max(|ref|, 1e-3)  (SURVEY.md §8d config 2).  The float64 state (pos, vel, returns)
is additionally required to be BIT-IDENTICAL to the oracle's, which is what makes the flags exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import merge_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def mg():
    import merging_gym_b200
    return merging_gym_b200


def assert_step_equal(out, ref, t=None, state=None):
    obs, rew, done, info = out
    robs, rrew, rdone, rinfo = ref
    flags = info["flags"].cpu().numpy() if not isinstance(info, np.ndarray) else info
    assert np.array_equal(done.cpu().numpy() if hasattr(done, "cpu") else done, rdone), f"done differs at step {t}"
    assert np.array_equal(flags, rinfo), f"info flags differ at step {t}"
    e = rel_err(obs.cpu().numpy() if hasattr(obs, "cpu") else obs, robs).max()
    assert e <= TOL, f"obs rel err {e} at step {t}"
    e = rel_err(rew.cpu().numpy() if hasattr(rew, "cpu") else rew, rrew).max()
    assert e <= TOL, f"reward rel err {e} at step {t}"


def assert_state_bit_exact(env, ref):
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2"):
        a = getattr(env, k).cpu().numpy()
        assert np.array_equal(a, getattr(ref, k)), f"state {k} is not bit-identical"
    assert np.array_equal(env.steps.cpu().numpy(), np.minimum(ref.steps, 4095))
    assert np.array_equal(env.winner.cpu().numpy(), ref.winner)


def test_reset_observation(mg):
    env = mg.MergeVecEnv(130)
    obs = env.reset().cpu().numpy()
    ref = mo.RefVecEnv(130).reset()
    assert rel_err(obs, ref).max() <= 1e-7
    assert obs[0].tolist() == np.float32(ref[0]).tolist()       # fp32 cast of the fp64 oracle value
    assert obs[0, 1] == np.float32(-30.057386826127185)


@pytest.mark.parametrize("pvp", [True, False])
def test_config2_trajectory_parity_4096(mg, pvp):
    """BASELINE.json configs[1]: 4096 envs, random discrete actions, 640 steps, auto-reset."""
    N, T = 4096, 640
    g = torch.Generator(device="cuda").manual_seed(1234)
    acts = torch.randint(0, 5, (T, N, 2), dtype=torch.uint8, device="cuda", generator=g)
    acts_h = acts.cpu().numpy()
    env = mg.MergeVecEnv(N, mode="pvp" if pvp else "pve")
    ref = mo.RefVecEnv(N, pvp=pvp)
    n_done = 0
    for t in range(T):
        out = env.step(acts[t, :, 0], acts[t, :, 1] if pvp else None)
        r = ref.step(acts_h[t, :, 0], acts_h[t, :, 1] if pvp else None)
        assert_step_equal(out, r, t)
        m = r[2]
        if m.any():
            n_done += int(m.sum())
            info = out[3]
            assert rel_err(info["terminal_observation"].cpu().numpy()[m], ref.terminal_obs[m]).max() <= TOL
            assert np.array_equal(info["episode_length"].cpu().numpy()[m], ref.ep_len[m])
            assert rel_err(info["episode_return"].cpu().numpy()[m], ref.ep_ret[m]).max() <= TOL
    assert n_done > 2 * N
    assert_state_bit_exact(env, ref)
    s, rs = env.stats(), ref.stats
    for k in ("episodes", "collisions", "wins_p1", "wins_p2", "timeouts", "merges_ok", "sum_length", "bad_actions"):
        assert s[k] == rs[k], k
    assert s["episodes"] == n_done
    assert abs(s["sum_return1"] - rs["sum_return1"]) <= 1e-6 * n_done
    assert abs(s["sum_return2"] - rs["sum_return2"]) <= 1e-6 * n_done


def test_golden_vec16_from_reference_file(mg):
    """Kernel vs the trajectory recorded from the unmodified reference file (16 envs x 640)."""
    tr = dict(np.load(os.path.join(GOLDEN, "pvp_vec16_trace.npz")))
    T, N = tr["done"].shape
    env = mg.MergeVecEnv(N)
    acts = torch.from_numpy(tr["actions"]).cuda()
    for t in range(T):
        obs, rew, done, info = env.step(acts[t, :, 0], acts[t, :, 1])
        assert np.array_equal(done.cpu().numpy(), tr["done"][t]), t
        assert np.array_equal(info["collision"].cpu().numpy(), tr["collision"][t]), t
        assert np.array_equal(info["winner"].cpu().numpy(), tr["winner"][t]), t
        assert rel_err(obs.cpu().numpy(), tr["obs"][t]).max() <= TOL, t
        assert rel_err(rew.cpu().numpy(), tr["rewards"][t]).max() <= TOL, t
        m = tr["done"][t]
        if m.any():
            assert np.array_equal(info["episode_length"].cpu().numpy()[m], tr["ep_len"][t][m])
            assert rel_err(info["episode_return"].cpu().numpy()[m], tr["ep_ret"][t][m]).max() <= TOL
            assert rel_err(info["terminal_observation"].cpu().numpy()[m], tr["step_obs"][t][m]).max() <= TOL


@pytest.mark.parametrize("name,pvp", [("config1_pve_trace.npz", False), ("pvp_trace.npz", True)])
def test_cuda_path_replays_the_reference_traces(mg, golden, name, pvp):
    """BASELINE.json configs[0] straight on the CUDA path: the 10 000-step pve trace (and the 6 000-step pvp trace) that
    the UNMODIFIED reference env produced — random actions, `env.reset()` after every `done` — replayed through `mg_step` /
    `mg_reset` with the recorded actions: done / collision / winner bit-exact at every step, observations, rewards and
    the float64 return accumulators within 1e-5 relative.  70 envs replay the same trace (two full warps of the
    vector path and a ragged tail of the scalar path); envs 0, 63 and 69 are checked."""
    tr = golden(name)
    assert bool(tr["pvp"]) == pvp
    n, probe = 70, [0, 63, 69]
    env = mg.MergeVecEnv(n, mode="pvp" if pvp else "pve", auto_reset=False, episode_info=False)
    obs0 = env.reset().cpu().numpy()
    assert rel_err(obs0[probe], np.broadcast_to(tr["reset_obs"], (3, 10))).max() <= TOL
    T = len(tr["done"])
    acts = torch.as_tensor(tr["actions"]).cuda()                       # [T, 2] uint8
    ones = torch.ones(n, dtype=torch.uint8, device="cuda")
    episodes = 0
    for t in range(T):
        a1 = ones * acts[t, 0]
        a2 = ones * acts[t, 1] if pvp else None
        obs, rew, done, info = env.step(a1, a2)
        d = done[probe].cpu().numpy()
        assert d.tolist() == [bool(tr["done"][t])] * 3, f"done differs at step {t}"
        assert info["collision"][probe].cpu().tolist() == [bool(tr["collision"][t])] * 3, f"collision differs at step {t}"
        assert info["winner"][probe].cpu().tolist() == [int(tr["winner"][t])] * 3, f"winner differs at step {t}"
        if t % 7 == 0 or d[0]:                                         # continuous outputs: every 7th step and every terminal step
            assert rel_err(obs[probe].cpu().numpy(), np.broadcast_to(tr["obs"][t], (3, 10))).max() <= TOL, t
            assert rel_err(rew[probe].cpu().numpy(), np.broadcast_to(tr["rewards"][t], (3, 2))).max() <= TOL, t
            ret = torch.stack([env.r1_accumulate[probe], env.r2_accumulate[probe]], 1).cpu().numpy()
            assert rel_err(ret, np.broadcast_to(tr["returns"][t], (3, 2))).max() <= TOL, t
        if d[0]:
            env.reset()
            episodes += 1
    assert episodes == int(tr["done"].sum()) > 10


def test_known_answer_episodes_sticky_done(mg):
    """All scripted KAT episodes side by side in one launch, auto_reset off, run to 2600 steps
    (covers the 2501-step time limit, '>=' vs '>', truncation vs rounding, winner-keeps-driving)."""
    kats = json.load(open(os.path.join(GOLDEN, "kat.json")))[:-1]
    kats = [k for k in kats if k["a2"] is not None]              # pvp scripts in this launch
    N, T = len(kats), 2600

    def act(s, t):
        return s if isinstance(s, int) else s[t % len(s)]
    env = mg.MergeVecEnv(N, auto_reset=False)
    ref = mo.RefVecEnv(N, pvp=True, auto_reset=False)
    first_done = np.full(N, -1)
    snap = {}
    for t in range(T):
        a1 = np.array([act(k["a1"], t) for k in kats], np.uint8)
        a2 = np.array([act(k["a2"], t) for k in kats], np.uint8)
        out = env.step(a1, a2)
        assert_step_equal(out, ref.step(a1, a2), t)
        d = out[2].cpu().numpy()
        for e in np.nonzero(d & (first_done < 0))[0]:
            first_done[e] = t + 1
            snap[e] = (int(env.winner[e]), bool(out[3]["collision"][e]), float(env.ret1[e]), float(env.ret2[e]),
                       float(env.pos1[e]), float(env.pos2[e]))
    assert_state_bit_exact(env, ref)
    for e, k in enumerate(kats):
        w, col, R1, R2, p1, p2 = snap[e]
        assert first_done[e] == k["steps"], k
        assert (w or None) == k["winner"] and col == k["collision"], k
        assert rel_err([R1, R2, p1, p2], [k["R1"], k["R2"], k["pos1"], k["pos2"]]).max() <= 1e-9, k


def test_known_answer_episodes_pve(mg):
    kats = json.load(open(os.path.join(GOLDEN, "kat.json")))[:-1]
    kats = [k for k in kats if k["a2"] is None]
    N = len(kats)
    env = mg.MergeVecEnv(N, mode="pve", auto_reset=False)
    ref = mo.RefVecEnv(N, pvp=False, auto_reset=False)
    first_done = np.full(N, -1)
    for t in range(2600):
        a1 = np.array([k["a1"] for k in kats], np.uint8)
        out = env.step(a1, None)
        assert_step_equal(out, ref.step(a1, None), t)
        d = out[2].cpu().numpy()
        first_done[d & (first_done < 0)] = t + 1
    assert first_done.tolist() == [k["steps"] for k in kats]
    assert_state_bit_exact(env, ref)


@pytest.mark.parametrize("n", [1, 2, 3, 63, 64, 65, 127, 511, 513, 1000])
def test_ragged_sizes(mg, n):
    """Sizes that are not a multiple of the 64-env warp tile take the scalar tail path."""
    rng = np.random.default_rng(n)
    env = mg.MergeVecEnv(n)
    ref = mo.RefVecEnv(n)
    for t in range(260):
        a = rng.integers(0, 5, (n, 2)).astype(np.uint8)
        assert_step_equal(env.step(a[:, 0], a[:, 1]), ref.step(a[:, 0], a[:, 1]), t)
    assert_state_bit_exact(env, ref)
    assert env.stats()["episodes"] == ref.stats["episodes"]


def test_empty_env(mg):
    env = mg.MergeVecEnv(0)
    obs, rew, done, info = env.step(torch.zeros(0, dtype=torch.uint8), torch.zeros(0, dtype=torch.uint8))
    assert obs.shape == (0, 10) and rew.shape == (0, 2) and done.shape == (0,)
    assert env.stats()["episodes"] == 0


def test_action_dtypes_and_bad_actions(mg):
    n = 300
    rng = np.random.default_rng(0)
    envs = [mg.MergeVecEnv(n) for _ in range(3)]
    ref = mo.RefVecEnv(n)
    for t in range(50):
        a = rng.integers(0, 5, (n, 2))
        if t == 10:
            a[5, 0] = 7; a[6, 1] = 200; a[7, 0] = 5           # KeyError in the reference
        outs = [envs[0].step(torch.tensor(a[:, 0], dtype=torch.int64).cuda(), torch.tensor(a[:, 1], dtype=torch.int64).cuda()),
                envs[1].step(torch.tensor(a[:, 0], dtype=torch.int32).cuda(), torch.tensor(a[:, 1], dtype=torch.int32).cuda()),
                envs[2].step(np.clip(a[:, 0], 0, 255).astype(np.uint8), (a[:, 1] % 256).astype(np.uint8))]
        r = ref.step(a[:, 0], a[:, 1])
        for o in outs:
            assert_step_equal(o, r, t)
        if t == 10:
            bad = outs[0][3]["bad_action"].cpu().numpy()
            assert bad[5] and bad[6] and bad[7] and bad.sum() == 3
    assert envs[0].stats()["bad_actions"] == 3
    # negative actions exist only for the signed dtypes: clamped to 0 and flagged
    e, r = mg.MergeVecEnv(8), mo.RefVecEnv(8)
    a = np.array([-1, 0, 1, 2, 3, 4, -7, 2])
    out = e.step(torch.tensor(a, dtype=torch.int64).cuda(), torch.tensor(a[::-1].copy(), dtype=torch.int32).cuda())
    assert_step_equal(out, r.step(a, a[::-1]))
    assert out[3]["bad_action"].cpu().numpy().tolist() == [True, True, False, False, False, False, True, True]
    v = mg.MergeVecEnv(4, validate_actions=True)
    with pytest.raises(KeyError):
        v.step(np.array([0, 1, 9, 2]), np.array([0, 0, 0, 0]))
    with pytest.raises(ValueError):
        v.step(np.zeros(3, np.uint8), np.zeros(3, np.uint8))
    with pytest.raises(TypeError):
        v.step(np.zeros(4, np.float32), np.zeros(4, np.float32))


def test_sample_actions_bit_exact(mg):
    for seed, base in [(0x5EED, 0), (2 ** 40 + 17, 2 ** 33 + 5)]:
        env = mg.MergeVecEnv(5000, seed=seed, env_id_base=base)
        for step in (0, 1, 2 ** 35 + 1):
            a1, a2 = env.sample_actions(step)
            r1, r2 = mo.philox_actions(5000, seed, base, step)
            assert np.array_equal(a1.cpu().numpy(), r1) and np.array_equal(a2.cpu().numpy(), r2)


@pytest.mark.parametrize("pvp,n", [(True, 4096), (False, 1000), (True, 77)])
def test_rollout_equals_stepwise_and_oracle(mg, pvp, n):
    K, mode = 300, "pvp" if pvp else "pve"
    a = mg.MergeVecEnv(n, mode=mode, seed=99, env_id_base=12345)
    b = mg.MergeVecEnv(n, mode=mode, seed=99, env_id_base=12345)
    ref = mo.RefVecEnv(n, pvp=pvp)
    obs = torch.empty(K, n, 10, device="cuda"); rew = torch.empty(K, n, 2, device="cuda")
    done = torch.empty(K, n, dtype=torch.uint8, device="cuda"); info = torch.empty(K, n, dtype=torch.uint8, device="cuda")
    acts = torch.empty(K, n, 2, dtype=torch.uint8, device="cuda")
    a.rollout(K, obs=obs, rew=rew, done=done, info=info, actions=acts)
    for t in range(K):
        a1, a2 = b.sample_actions()
        assert torch.equal(acts[t, :, 0], a1)
        o, r, d, i = b.step(a1, a2)
        assert torch.equal(o, obs[t]) and torch.equal(r, rew[t])
        assert torch.equal(d.view(torch.uint8), done[t]) and torch.equal(i["flags"], info[t])
        r1, r2 = mo.philox_actions(n, 99, 12345, t)
        assert_step_equal((o, r, d, i), ref.step(r1, r2 if pvp else None), t)
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert_state_bit_exact(a, ref)
    assert a.stats() == b.stats()
    assert a.stats()["episodes"] == ref.stats["episodes"]
    # outputs are optional
    c = mg.MergeVecEnv(n, mode=mode, seed=99, env_id_base=12345)
    c.rollout(K)
    assert torch.equal(c.pos1, a.pos1) and c.stats() == a.stats()


def test_sticky_done_parity(mg):
    n = 512
    rng = np.random.default_rng(3)
    env = mg.MergeVecEnv(n, auto_reset=False)
    ref = mo.RefVecEnv(n, auto_reset=False)
    for t in range(400):
        a = rng.integers(0, 5, (n, 2)).astype(np.uint8)
        assert_step_equal(env.step(a[:, 0], a[:, 1]), ref.step(a[:, 0], a[:, 1]), t)
    assert bool(env.done.all())
    assert env.stats()["episodes"] == n == ref.stats["episodes"]
    # masked manual reset (the reference's caller-side `env.reset()` after done, main.py:190)
    mask = torch.zeros(n, dtype=torch.bool); mask[::3] = True
    obs = env.reset(mask).cpu().numpy()
    robs = ref.reset(mask.numpy())
    assert rel_err(obs, robs).max() <= TOL
    assert_state_bit_exact(env, ref)
    assert env.done.cpu().numpy().tolist() == (~mask.numpy()).tolist()


def test_custom_rewards(mg):
    rw = dict(r_first=5.0, r_second=0.5, r_collision=-3.0, vel_penalty=0.01, time_penalty=0.002)
    n = 256
    env = mg.MergeVecEnv(n, rewards=rw)
    assert env.show_reward() == (5.0, 0.5, -3.0, 0.01)
    old = (mo.RFirst, mo.RSecond, mo.RCollision, mo.vel_penalty, mo.time_penalty)
    try:
        mo.RFirst, mo.RSecond, mo.RCollision, mo.vel_penalty, mo.time_penalty = 5.0, 0.5, -3.0, 0.01, 0.002
        ref = mo.RefVecEnv(n)
        rng = np.random.default_rng(8)
        for t in range(300):
            a = rng.integers(0, 5, (n, 2)).astype(np.uint8)
            assert_step_equal(env.step(a[:, 0], a[:, 1]), ref.step(a[:, 0], a[:, 1]), t)
    finally:
        mo.RFirst, mo.RSecond, mo.RCollision, mo.vel_penalty, mo.time_penalty = old


def test_out_slots_ring_and_state_dict(mg):
    n = 128
    env = mg.MergeVecEnv(n, out_slots=3)
    a1, a2 = env.sample_actions()
    o1 = env.step(a1, a2)[0]
    ck = env.state_dict()
    a1, a2 = env.sample_actions()
    o2 = env.step(a1, a2)[0]
    assert o1.data_ptr() != o2.data_ptr() and not torch.equal(o1, o2)
    keep = o2.clone(); st = env.stats()
    env.load_state_dict(ck)
    a1, a2 = env.sample_actions()
    assert torch.equal(env.step(a1, a2)[0], keep) and env.stats() == st


def test_cuda_graph_replay_matches_eager(mg):
    n, K = 2048, 40
    a = mg.MergeVecEnv(n, seed=5); b = mg.MergeVecEnv(n, seed=5)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a1, a2 = a.sample_actions(0); a.step(a1, a2)          # warm-up outside capture
        a.load_state_dict(mg.MergeVecEnv(n, seed=5).state_dict())
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for t in range(K):
            a1, a2 = a.sample_actions(t)
            out = a.step(a1, a2)
    g.replay()
    for t in range(K):
        a1, a2 = b.sample_actions(t)
        ref = b.step(a1, a2)
    torch.cuda.synchronize()
    assert torch.equal(out[0], ref[0]) and torch.equal(a.pos1, b.pos1) and torch.equal(a.meta, b.meta)


def test_step_host_matches_device_path(mg):
    n = 1000
    rng = np.random.default_rng(4)
    a = mg.MergeVecEnv(n); b = mg.MergeVecEnv(n)
    for t in range(120):
        act = rng.integers(0, 5, (n, 2)).astype(np.uint8)
        ho, hr, hd, hi = a.step_host(act[:, 0], act[:, 1], direct_actions=bool(t % 2))   # with / without the explicit upload
        o, r, d, i = b.step(act[:, 0], act[:, 1])
        assert np.array_equal(ho, o.cpu().numpy()) and np.array_equal(hr, r.cpu().numpy())
        assert np.array_equal(hd, d.cpu().numpy()) and np.array_equal(hi, i["flags"].cpu().numpy())
    for t in range(40):                                         # zero-copy variant: kernel writes pinned host memory
        act = rng.integers(0, 5, (n, 2)).astype(np.uint8)
        ho, hr, hd, hi = a.step_host(act[:, 0], act[:, 1], zero_copy=True)
        o, r, d, i = b.step(act[:, 0], act[:, 1])
        assert np.array_equal(ho, o.cpu().numpy()) and np.array_equal(hr, r.cpu().numpy())
        assert np.array_equal(hd, d.cpu().numpy()) and np.array_equal(hi, i["flags"].cpu().numpy())
    ho, *_ = a.step_host(act[:, 0], None)                       # pve through the host path
    o, *_ = b.step(act[:, 0], None)
    assert np.array_equal(ho, o.cpu().numpy())


@pytest.mark.parametrize("chunks", [1, 3, 16])
def test_step_host_pipeline_pieces(mg, chunks):
    """`mg_step_host` stepping the envs in pieces on two streams (D2H of one piece under the upload + kernel of the
    next) gives exactly the single-launch result — ragged last piece, random-start resets (their Philox counter
    is the GLOBAL env id, so each piece must carry its offset), statistics, actions passed in the pinned buffers."""
    n = 100_037
    kw = dict(reset_mode="random", reset_seed=5, env_id_base=1 << 33)
    a = mg.MergeVecEnv(n, **kw); b = mg.MergeVecEnv(n, **kw)
    p1, p2 = a.host_action_buffers()
    rng = np.random.default_rng(chunks)
    for t in range(260):
        p1[:] = rng.integers(0, 5, n); p2[:] = rng.integers(0, 5, n)
        ho, hr, hd, hi = a.step_host(p1, p2, chunks=chunks, direct_actions=bool((t // 7) % 2))
        o, r, d, i = b.step(p1.copy(), p2.copy())
        if t % 20 == 0 or t > 250:
            assert np.array_equal(ho, o.cpu().numpy()) and np.array_equal(hr, r.cpu().numpy())
            assert np.array_equal(hd, d.cpu().numpy()) and np.array_equal(hi, i["flags"].cpu().numpy())
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert a.stats() == b.stats() and a.stats()["episodes"] > n // 2
    assert torch.equal(a.terminal_obs, b.terminal_obs)


def test_scalar_reference_interface(mg):
    """`gym.make("merging_env-v0")`-style use, as scripts/main.py:190-218 drives it."""
    env = mg.make("merging_env-v0")
    assert env.action_space.n == 5 and env.observation_space.shape[0] == 10
    assert env.show_reward() == (2.0, 1.0, -10, 0.001)
    ref = mo.RefEnv()
    obs = env.reset()
    assert isinstance(obs, list) and len(obs) == 10 and rel_err(obs, ref.reset()).max() <= TOL
    n = 0
    while True:
        n += 1
        o, r, d, info = env.step(2, 2)
        o2, r2, d2, info2 = ref.step(2, 2)
        assert d == d2 and info == info2 and rel_err(o, o2).max() <= TOL and rel_err(r, r2).max() <= TOL
        if d:
            break
    assert n == 151 and env.winner is None and info == {"collision": True}
    assert env.r1_accumulate == -10.0 and env.state1["pos"] == 654.0
    env.reset()
    for t in range(225):
        o, r, d, info = env.step(3)                              # pve: action2=None
    assert d and env.winner == 1 and env.state2["pos"] == 950.0 and env.r2_accumulate == 1.0
    with pytest.raises(KeyError):
        env.step(5)
    assert state_swap_ok(o)


def state_swap_ok(o):
    import merging_gym_b200 as m
    v = m.MergeVecEnv.opponent_view(torch.tensor([o]))[0].tolist()
    return v == o[5:] + o[:5]


def test_random_start_reset(mg):
    """reset_mode="random": the reference's commented-out random start (merging_env.py:219-221)."""
    n, seed, base = 4096, 77, 1 << 33
    env = mg.MergeVecEnv(n, reset_mode="random", reset_seed=seed, env_id_base=base)
    ref = mo.RefVecEnv(n, reset_mode="random", reset_seed=seed, env_id_base=base)
    obs = env.obs_buf[0].cpu().numpy()
    for k in ("pos1", "vel1"):                                     # Box-Muller: log/sincospi differ by ulps
        assert rel_err(getattr(env, k).cpu().numpy(), getattr(ref, k)).max() <= 1e-12, k
    for k in ("pos2", "vel2"):                                     # uniforms use exact arithmetic only
        assert np.array_equal(getattr(env, k).cpu().numpy(), getattr(ref, k)), k
    assert rel_err(obs, ref.observe()).max() <= 1e-5
    assert env.resets.cpu().tolist() == [1] * n
    g = torch.Generator(device="cuda").manual_seed(5)
    n_resets = 0
    for t in range(500):
        # continue from the GPU's float64 start so the dynamics can be compared bit for bit
        for k in ("pos1", "vel1"):
            getattr(ref, k)[:] = getattr(env, k).cpu().numpy()
        a = torch.randint(0, 5, (n, 2), dtype=torch.uint8, device="cuda", generator=g)
        out = env.step(a[:, 0], a[:, 1])
        r = ref.step(a[:, 0].cpu().numpy(), a[:, 1].cpu().numpy())
        assert_step_equal(out, r, t)
        n_resets += int(r[2].sum())
    assert n_resets > n
    assert np.array_equal(env.resets.cpu().numpy(), ref.resets)
    assert np.array_equal(env.pos2.cpu().numpy(), ref.pos2) and np.array_equal(env.vel2.cpu().numpy(), ref.vel2)
    p1 = env.pos1.cpu().numpy()
    assert 30 < np.percentile(p1, 1)                               # sanity: every env keeps moving forward


def test_random_start_is_shard_and_launch_invariant(mg):
    n, K = 2048, 300
    kw = dict(reset_mode="random", reset_seed=9, seed=3)
    full = mg.MergeVecEnv(n, **kw); full.rollout(K)
    halves = [mg.MergeVecEnv(n // 2, env_id_base=b, **kw) for b in (0, n // 2)]
    for h in halves:
        for t in range(K):
            h.step(*h.sample_actions())
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "meta"):
        assert torch.equal(torch.cat([getattr(h, k) for h in halves]), getattr(full, k)), k
    assert int(full.resets.max()) >= 2
    m = torch.zeros(n, dtype=torch.bool); m[:100] = True
    before = full.resets.clone()
    full.reset(m)
    assert torch.equal(full.resets[:100], before[:100] + 1) and torch.equal(full.resets[100:], before[100:])


def test_far_beyond_the_merge_zone(mg):
    """Sticky-done envs that keep being stepped: longitudes past 24 000 leave the polynomial's range
    (library sincos fallback) and the step counter saturates at 4095 instead of wrapping."""
    n, T = 6, 4300
    env = mg.MergeVecEnv(n, auto_reset=False)
    ref = mo.RefVecEnv(n, auto_reset=False)
    a1 = np.array([4, 4, 3, 0, 2, 4], np.uint8); a2 = np.array([4, 0, 4, 4, 2, 1], np.uint8)
    for t in range(T):
        out = env.step(a1, a2)
        r = ref.step(a1, a2)
        if t % 50 == 0 or t > T - 20:
            assert_step_equal(out, r, t)
    assert float(env.pos1.max()) > 24000.0 and int(ref.steps.max()) == T
    assert_state_bit_exact(env, ref)                    # pos/vel stay bit-identical; steps compare saturated
    assert env.steps.cpu().tolist() == [4095] * n


def test_gym_registration_drop_in(mg, monkeypatch):
    """The reference's plugin boundary is gym's registry (merging_gym/__init__.py:3-6): after
    `register_gym()`, `gym.make("merging_env-v0")` hands the scripts the GPU-backed env.  gym itself
    is absent from the image, so the registry used here is the stand-in from oracle/ref_shims."""
    import sys
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_shims")
    monkeypatch.syspath_prepend(shims)
    for m in [k for k in sys.modules if k == "gym" or k.startswith("gym.")]:
        monkeypatch.delitem(sys.modules, m)
    import gym
    assert mg.register_gym() == "merging_env-v0"
    env = gym.make("merging_env-v0")
    env = env.unwrapped                                           # scripts/main.py:20-23
    assert isinstance(env, mg.MergeEnv)
    assert env.action_space.n == 5 and env.observation_space.shape[0] == 10      # main.py:24-25
    state = env.reset()
    for t in range(151):                                          # main.py:192-218 loop body
        action_op = None
        next_state, rewards, done, info = env.step(2, action_op)
        reward, reward_op = rewards
        state = next_state
    assert done and info["collision"] and env.winner is None and t == 150


@pytest.mark.parametrize("pvp", [True, False])
def test_injected_states_parity(mg, pvp):
    """States no trajectory from the fixed start reaches: both cars anywhere on [0, 1100) with any speed in
    [0, 45), many of them inside the 8 x 4 collision box around the merge point, on either side of END_POINT,
    or crossing it together.  The float64 state is written into both implementations, then 12 steps with
    random actions: flags bit-exact, float64 state bit-identical, obs / rewards within 1e-5."""
    n, T = 1 << 15, 12
    rng = np.random.default_rng(77 + pvp)
    p1 = rng.uniform(0.0, 1100.0, n); p2 = rng.uniform(0.0, 1100.0, n)
    k = n // 2                                                   # half of them: close pairs near the merge zone
    p1[:k] = rng.uniform(900.0, 1010.0, k); p2[:k] = p1[:k] + rng.normal(0.0, 6.0, k)
    p1[k:k + 4096] = 950.0 - rng.uniform(0.0, 9.0, 4096); p2[k:k + 4096] = 950.0 - rng.uniform(0.0, 9.0, 4096)   # photo finishes
    v1 = rng.uniform(0.0, 45.0, n); v2 = rng.uniform(0.0, 45.0, n)
    v1[::7] = 0.0
    env = mg.MergeVecEnv(n, mode="pvp" if pvp else "pve", auto_reset=True)
    ref = mo.RefVecEnv(n, pvp=pvp, auto_reset=True)
    for name, val in (("pos1", p1), ("vel1", v1), ("pos2", p2), ("vel2", v2)):
        getattr(env, name).copy_(torch.from_numpy(val))
        getattr(ref, name)[:] = val
    acts = rng.integers(0, 5, (T, 2, n)).astype(np.uint8)
    seen_collision = seen_w1 = seen_w2 = 0
    for t in range(T):
        a1, a2 = acts[t]
        out = env.step(torch.from_numpy(a1).cuda(), torch.from_numpy(a2).cuda() if pvp else None)
        r = ref.step(a1, a2 if pvp else None)
        assert_step_equal(out, r, t)
        seen_collision += int((r[3] & 1).sum()); w = (r[3] & mo.INFO_WINNER_MASK) >> 1
        seen_w1 += int((w == 1).sum()); seen_w2 += int((w == 2).sum())
    assert_state_bit_exact(env, ref)
    assert seen_collision > 1000 and seen_w1 > 1000 and seen_w2 > 1000     # the interesting branches were all taken


def test_injected_states_against_the_reference_fixture(mg, golden):
    """tests/golden/injected_states.npz: one step of the UNMODIFIED reference env from 6 000 injected states
    (oracle/make_golden.py --injected).  The CUDA kernel, fed the same float64 state and winner through the
    state arrays, must return the reference's collision / done / winner bit for bit, its float64 state after
    the step, and its observations and rewards within 1e-5."""
    tr = golden("injected_states.npz")
    for pvp in (True, False):
        idx = np.nonzero(tr["pvp"] == pvp)[0]
        n = len(idx)
        env = mg.MergeVecEnv(n, mode="pvp" if pvp else "pve", auto_reset=False)
        for j, k in enumerate(("pos1", "vel1", "pos2", "vel2")):
            getattr(env, k).copy_(torch.from_numpy(tr["pos"][idx, j]))
        env.meta.copy_(torch.from_numpy(tr["winner_before"][idx].astype(np.int32) << 12).to(env.meta.dtype))
        a1 = torch.from_numpy(tr["actions"][idx, 0]).cuda(); a2 = torch.from_numpy(tr["actions"][idx, 1]).cuda()
        obs, rew, done, info = env.step(a1, a2 if pvp else None)
        assert np.array_equal(done.cpu().numpy().astype(bool), tr["done"][idx])
        assert np.array_equal(info["collision"].cpu().numpy(), tr["collision"][idx])
        assert np.array_equal(env.winner.cpu().numpy(), tr["winner"][idx])
        after = torch.stack([env.pos1, env.vel1, env.pos2, env.vel2], 1).cpu().numpy()
        assert rel_err(after, tr["state_after"][idx]).max() < 1e-12
        assert rel_err(obs.cpu().numpy(), tr["obs"][idx]).max() <= TOL
        assert rel_err(rew.cpu().numpy(), tr["rewards"][idx]).max() <= TOL
