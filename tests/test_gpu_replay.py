"""Device-resident transition ring and CSV episode logs against a
sequential NumPy emulation of the reference's `store_transition` / CSV writer driven by the oracle."""
import csv

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import merge_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import merging_gym_b200
    return merging_gym_b200


@pytest.mark.parametrize("n,cap,fmt,player", [(777, 5000, "replay", 1), (64, 100000, "replay", 2),
                                               (1000, 300, "log", 1), (4096, 50000, "replay", 1)])
def test_transition_ring_matches_sequential_store(mg, n, cap, fmt, player):
    T = 260
    env = mg.MergeVecEnv(n, out_slots=2, seed=21)
    rec = mg.TransitionRecorder(env, cap, format=fmt, player=player, track_env_ids=True)
    ref = mo.RefVecEnv(n)
    width = 22 if fmt == "replay" else 14
    mem = np.zeros((cap, width)); ids = np.full(cap, -1); counter = 0
    obs_prev = env.reset(); robs_prev = ref.reset()
    for t in range(T):
        a1, a2 = env.sample_actions()
        out = env.step(a1, a2)
        rec.record(obs_prev, a1, a2, out)
        h1, h2 = a1.cpu().numpy(), a2.cpu().numpy()
        robs, rrew, rdone, rinfo = ref.step(h1, h2)
        winner = (rinfo & mo.INFO_WINNER_MASK) >> 1
        for e in np.nonzero(winner != 1)[0]:                    # main.py:209 `if env.winner is not 1`
            if fmt == "replay":                                 # main.py:116 hstack((state, [action, reward], next_state))
                nxt = ref.terminal_obs[e] if rdone[e] else robs[e]
                row = np.hstack((robs_prev[e], [(h1, h2)[player - 1][e], rrew[e, player - 1]], nxt))
            else:                                               # human_player.py:181 state + [a, a_op] + rewards
                row = np.hstack((robs_prev[e], [h1[e], h2[e]], rrew[e]))
            mem[counter % cap] = row; ids[counter % cap] = e; counter += 1
        obs_prev, robs_prev = out[0], robs
    assert int(rec.counter.item()) == counter and len(rec) == min(counter, cap)
    assert np.array_equal(rec.env_ids.cpu().numpy(), ids)
    assert rel_err(rec.ring.cpu().numpy(), mem).max() <= 1e-5
    rows = rec.rows().cpu().numpy()
    assert rows.shape == (min(counter, cap), width)
    if counter > cap:
        assert rel_err(rows[0], mem[counter % cap]).max() <= 1e-5      # oldest row first
    assert rec.sample(128).shape == (128, width)


def test_transition_ring_large_n_multi_pass_scan(mg):
    """n = 5 * 2^20 + 300 envs: 20 482 blocks, so the single-CTA scan of the block counts takes several passes and
    carries between them.  Checked on the device: the ring must hold the selected envs' rows in env-id order
    (the order `torch.nonzero` gives), appended call after call."""
    n = 5 * (1 << 20) + 300
    env = mg.MergeVecEnv(n, out_slots=2, seed=5, episode_info=True)
    env.rollout(215)                                           # mid-run: many envs finish every step
    rec = mg.TransitionRecorder(env, 3 * n, track_env_ids=True)
    obs_prev = env.obs_buf[env._slot].clone()
    expect_ids, expect_rows = [], []
    for t in range(2):
        a1, a2 = env.sample_actions()
        out = env.step(a1, a2)
        rec.record(obs_prev, a1, a2, out)
        obs, rew, done, info = out
        keep = torch.nonzero(info["winner"] != 1).squeeze(1)
        nxt = torch.where(done.bool().unsqueeze(1), info["terminal_observation"], obs)
        expect_ids.append(keep)
        expect_rows.append(torch.cat([obs_prev[keep], a1[keep].float().unsqueeze(1), rew[keep, 0:1], nxt[keep]], 1))
        obs_prev = obs.clone()
    ids, rows = torch.cat(expect_ids), torch.cat(expect_rows)
    assert int(rec.counter.item()) == ids.numel() and 0 < ids.numel() < 2 * n
    assert torch.equal(rec.env_ids[:ids.numel()].long(), ids)
    assert torch.equal(rec.ring[:ids.numel()], rows)


@pytest.mark.parametrize("n,extra,density", [(100_003, 37, 0.5), (1 << 18, 0, 0.03), (300, 1, 0.9), (70_000, 5000, 1.0),
                                             (4099, -1000, 0.7)])
def test_writer_wraps_and_replays_in_a_graph(mg, n, extra, density):
    """The ring wraps at a different row of every call (capacity = n + extra), the explicit mask changes from call to
    call (the count pass reads it 16 bytes per thread), and the last calls are replays of one captured CUDA graph (count,
    scan and the programmatically dependent write pass): every call must leave exactly what a sequential
    `store_transition` loop over env ids leaves."""
    cap = n + extra
    env = mg.MergeVecEnv(n, out_slots=2, seed=3, episode_info=True)
    env.rollout(100)
    rec = mg.TransitionRecorder(env, cap, mask="explicit", track_env_ids=True)
    obs_prev = env.obs_buf[env._slot].clone()
    a1, a2 = env.sample_actions()
    out = env.step(a1, a2)
    obs, rew, done, info = out
    nxt = torch.where(done.bool().unsqueeze(1), info["terminal_observation"], obs)
    all_rows = torch.cat([obs_prev, a1.float().unsqueeze(1), rew[:, 0:1], nxt], 1)
    gen = torch.Generator(device="cuda").manual_seed(n)
    # odd sizes: the mask starts one byte into its allocation — the count pass then takes its byte-by-byte path instead of
    # the 16-byte loads
    select = torch.zeros(n + 1, dtype=torch.uint8, device="cuda")[1:] if n % 2 else torch.zeros(n, dtype=torch.uint8, device="cuda")
    want_ring = torch.zeros(cap, 22, device="cuda"); want_ids = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
    counter = 0

    def expect():
        nonlocal counter
        keep = torch.nonzero(select).squeeze(1)
        slots = (counter + torch.arange(keep.numel(), device="cuda")) % cap
        want_ring[slots] = all_rows[keep]; want_ids[slots] = keep.int()
        counter += keep.numel()

    def new_mask():
        select.copy_((torch.rand(n, device="cuda", generator=gen) < density).to(torch.uint8))

    for call in range(7):
        new_mask()
        rec.record(obs_prev, a1, a2, out, select=select)
        expect()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        rec.record(obs_prev, a1, a2, out, select=select); expect()          # warm-up on the capture stream
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            rec.record(obs_prev, a1, a2, out, select=select)
            rec.record(obs_prev, a1, a2, out, select=select)
    torch.cuda.current_stream().wait_stream(side)
    for replay in range(4):
        new_mask()
        g.replay(); expect(); expect()
    torch.cuda.synchronize()
    assert int(rec.counter.item()) == counter and (counter > cap or density < 0.1)
    assert torch.equal(rec.env_ids, want_ids)
    assert torch.equal(rec.ring, want_ring)


def test_recorder_all_mask_and_sticky_done(mg):
    n = 200
    env = mg.MergeVecEnv(n, out_slots=2, auto_reset=False, episode_info=False)
    rec = mg.TransitionRecorder(env, 100000, mask="all")
    obs_prev = env.reset()
    for t in range(50):
        a1, a2 = env.sample_actions()
        out = env.step(a1, a2)
        rec.record(obs_prev, a1, a2, out)
        obs_prev = out[0]
    assert len(rec) == 50 * n
    rows = rec.rows()
    assert torch.equal(rows[-n:, 12:], out[0])          # s' of the last step is the returned obs (no auto-reset)
    assert torch.equal(rows[-n:, 10], a1.float())


def test_csv_episode_logger(mg, tmp_path):
    n, ids = 32, [3, 17]
    env = mg.MergeVecEnv(n, out_slots=2, seed=5)
    log = mg.CsvEpisodeLogger(env, ids, str(tmp_path))
    refs = {e: mo.RefEnv() for e in ids}
    want = {e: [[]] for e in ids}
    state = {e: refs[e].reset() for e in ids}
    obs_prev = env.reset()
    for t in range(450):
        a1, a2 = env.sample_actions()
        out = env.step(a1, a2)
        log.log(obs_prev, a1, a2, out)
        obs_prev = out[0]
        for e in ids:
            x1, x2 = int(a1[e]), int(a2[e])
            nxt, rewards, done, info = refs[e].step(x1, x2)
            if refs[e].winner != 1:                                # human_player.py:180
                want[e][-1].append([float(v) for v in state[e]] + [x1, x2] + [float(v) for v in rewards])
            state[e] = nxt
            if done:
                state[e] = refs[e].reset()
                want[e].append([])
    log.close()
    for e in ids:
        episodes = [w for w in want[e] if w]
        files = sorted([f for f in log.files if f.endswith(f"env{e}.csv")], key=lambda p: int(p.split("episode")[-1].split(" ")[0]))
        assert len(files) == len(episodes) >= 2
        for path, rows in zip(files, episodes):
            got = list(csv.reader(open(path)))
            assert got[0] == mg.replay.CSV_HEADER and len(got) - 1 == len(rows)
            g = np.array([[float(v) for v in r] for r in got[1:]])
            assert rel_err(g, np.array(rows)).max() <= 1e-5


def test_replay_rows_against_the_reference_learner(mg):
    """tests/golden/replay_memory.npz: three episodes driven by the reference's OWN `DQN` class (scripts/main.py:
    `choose_action` with its epsilon rule under a fixed NumPy seed, `store_transition`, the `env.winner is not 1`
    guard of the episode loop) in the unmodified env.  Replaying the recorded actions through MergeVecEnv +
    TransitionRecorder must leave the same rows in the ring: 517 of 680 steps stored."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "replay_memory.npz"))
    acts, dones, mem, counter = z["actions"], z["done"], z["memory"], int(z["counter"])
    env = mg.MergeVecEnv(1, mode="pvp", auto_reset=False, out_slots=2)
    rec = mg.TransitionRecorder(env, 2000)                      # MEMORY_CAPACITY, main.py:16
    obs = env.reset().clone()
    for t in range(len(acts)):
        a1 = torch.tensor([acts[t, 0]], dtype=torch.uint8, device="cuda")
        a2 = torch.tensor([acts[t, 1]], dtype=torch.uint8, device="cuda")
        out = env.step(a1, a2)
        rec.record(obs, a1, a2, out)
        assert bool(out[2][0]) == bool(dones[t]), t
        obs = (env.reset() if dones[t] else out[0]).clone()
    assert int(rec.counter.item()) == counter == 517 and mem.shape == (517, 22)
    ring = rec.ring[:counter].cpu().numpy()
    assert np.array_equal(ring[:, 10], mem[:, 10])                               # the action column, exactly
    assert rel_err(ring, mem).max() <= 1e-5


@pytest.mark.parametrize("seed", [7, 36])
def test_hdqn_controller_rows_against_the_reference(mg, seed):
    """tests/golden/hdqn_policies.npz `controller_memory`: the rows the reference's OWN `HDQN.store_transition`
    (scripts/hdqn.py:180-184) left in `lower.memory` over one greedy episode driven by its nested option loop
    (:276-324) — `[g, s, a, r_int, g', s']`, r_int = 1 where the re-chosen goal equals goal_status(s).
    HDQNPolicy + TransitionRecorder(format="hdqn") must leave the same rows in the ring."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "hdqn_policies.npz"))
    tag = f"seed{seed}"
    sd = lambda name: {k.split("/")[-1]: z[k] for k in z.files if k.startswith(f"{tag}/{name}/")}
    mem = z[f"{tag}/L0/controller_memory"]
    pol = mg.HDQNPolicy(meta_state=sd("meta"), ctrl_state=sd("ctrl"))
    env = mg.MergeVecEnv(1, mode="pve", auto_reset=False, out_slots=2)
    rec = mg.TransitionRecorder(env, 2000, format="hdqn", mask="all")         # MEMORY_CAPACITY, hdqn.py:21
    obs = env.reset().clone()
    done = torch.zeros(1, dtype=torch.bool)
    for t in range(len(mem)):
        a = pol.act(obs).clone(); g = pol.goal.clone()
        out = env.step(a, None)
        g_next = pol.meta.act(out[0]).clone()                                   # hdqn.py:303
        rec.record(obs, a, None, out, goal_prev=g, goal_next=g_next)
        obs = out[0].clone()
    assert bool(out[2][0]) and int(rec.counter.item()) == len(mem)
    ring = rec.ring[:len(mem)].cpu().numpy()
    for col in (0, 11, 12, 13):                                                 # g, a, r_int, g': exact
        assert np.array_equal(ring[:, col], mem[:, col]), col
    assert rel_err(ring, mem).max() <= 1e-5
    assert mem[:, 12].sum() == (71 if seed == 36 else 0)


@pytest.mark.parametrize("seed", [7, 36])
def test_hdqn_meta_rows_against_the_reference(mg, seed):
    """tests/golden/hdqn_policies.npz `meta_memory`: the rows the reference's OWN `Goal_DQN.store_transition`
    (scripts/hdqn.py:97-101, called at :318 when an option ends: `done or goal == goal_status(state)`, :316) left in
    `upper.memory` — `[s_end, goal, sum of ego rewards over the option, s_end]`.  `OptionRecorder` driven by
    HDQNPolicy must leave the same rows (seed 36: 70 options in 225 steps; seed 7: one option = the whole episode)."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "hdqn_policies.npz"))
    tag = f"seed{seed}"
    sd = lambda name: {k.split("/")[-1]: z[k] for k in z.files if k.startswith(f"{tag}/{name}/")}
    mem, steps = z[f"{tag}/L0/meta_memory"], len(z[f"{tag}/L0/controller_memory"])
    pol = mg.HDQNPolicy(meta_state=sd("meta"), ctrl_state=sd("ctrl"))
    env = mg.MergeVecEnv(3, mode="pve", auto_reset=False, out_slots=2)
    opt = mg.OptionRecorder(env, 3 * 200, track_env_ids=True)                    # 3 envs x GOAL_MEMORY_CAPACITY (hdqn.py:22)
    obs = env.reset().clone()
    ended_total = 0
    for t in range(steps):
        a = pol.act(obs).clone()
        out = env.step(a, None)
        g_next = pol.meta.act(out[0]).clone()
        ended_total += int(opt.record(out, g_next).sum())
        obs = out[0].clone()
    assert bool(out[2].all()) and int(opt.counter.item()) == 3 * len(mem) == ended_total
    ring = opt.ring[:3 * len(mem)].cpu().numpy()
    for e in range(3):                                                           # three identical envs, rows interleaved
        rows = ring[e::3]
        assert np.array_equal(rows[:, 10], mem[:, 10])                           # the goal column, exactly
        assert rel_err(rows, mem).max() <= 1e-5
    assert len(mem) == (70 if seed == 36 else 1)
