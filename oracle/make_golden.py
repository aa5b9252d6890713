#!/usr/bin/env python
"""TEST INFRASTRUCTURE — records golden vectors from the UNMODIFIED reference env.

Run in the build container only (needs `/root/reference`):

    python oracle/make_golden.py            # rewrites tests/golden/*

It executes `/root/reference/merging_gym/envs/merging_env.py` and
`/root/reference/scripts/helper.py` as they are, with the third-party imports that
are missing from the image replaced by `oracle/ref_shims/` (see its README), and
stores what `reset()` / `step()` return:

* `config1_pve_trace.npz`  BASELINE.json configs[0]: pve, one env, 10 000 steps,
  actions `np.random.default_rng(0).integers(5)`, manual reset on done.
* `pvp_trace.npz`          same protocol, two players, rng seed 1, 6 000 steps.
* `pvp_vec16_trace.npz`    16 envs x 640 steps, pvp, per-env manual reset replayed with
  the gym-0.20 vector convention (a finished env returns its reset observation),
  the fixture the CUDA kernel is compared with directly.
* `kat.json`               known-answer episodes for scripted action pairs
  (SURVEY.md §8c): steps, winner, collision, returns, final positions, last obs.
* `injected_states.npz`    (`--injected`) 6 000 single steps from states written into the reference env
  (`state1/state2['pos'|'vel']`) — close pairs around the merge point, photo finishes at END_POINT, stopped
  cars, far-apart cars — with the winner the reference would hold there; pins the closed-form collision /
  winner / reward logic of the oracle away from the trajectories the fixed start can reach.
* `replay_memory.npz`      (`--replay`) three episodes driven by the reference's own `DQN` learner class
  (scripts/main.py: epsilon rule under a fixed NumPy seed, `store_transition`): actions and the memory rows.
* `hdqn_policies.npz`      (`--hdqn`) the reference's own h-DQN classes (scripts/hdqn.py) under fixed torch seeds,
  greedy, against the constant-speed opponent and in self-play: weights, goals and actions per step.
* `dqn_policies.npz`       (`--policies`) weights of two shipped DQN checkpoints and their greedy
  episodes against the L0 opponent in the reference env (policy-in-the-loop).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_loader import load_reference_env, quiet  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _f(v):
    return np.asarray(v, dtype=np.float64)


def trace(env, pvp, seed, steps):
    rng = np.random.default_rng(seed)
    acts = np.zeros((steps, 2), np.uint8)
    obs = np.zeros((steps, 10)); rew = np.zeros((steps, 2))
    done = np.zeros(steps, bool); col = np.zeros(steps, bool)
    win = np.zeros(steps, np.uint8)
    ret = np.zeros((steps, 2))
    reset_obs = _f(env.reset())
    with quiet():
        for t in range(steps):
            a1 = int(rng.integers(5))
            a2 = int(rng.integers(5)) if pvp else None
            acts[t] = (a1, a2 if pvp else 0)
            o, r, d, info = env.step(a1, a2)
            obs[t], rew[t], done[t], col[t] = _f(o), _f(r), d, info["collision"]
            win[t] = env.winner or 0
            ret[t] = (env.r1_accumulate, env.r2_accumulate)
            if d:
                env.reset()
    return dict(actions=acts, obs=obs, rewards=rew, done=done, collision=col, winner=win,
                returns=ret, reset_obs=reset_obs, pvp=np.array(pvp), seed=np.array(seed))


def vec_trace(env, n_envs, steps, seed):
    rng = np.random.default_rng(seed)
    acts = rng.integers(0, 5, size=(steps, n_envs, 2)).astype(np.uint8)
    obs = np.zeros((steps, n_envs, 10)); term = np.zeros((steps, n_envs, 10))
    rew = np.zeros((steps, n_envs, 2))
    done = np.zeros((steps, n_envs), bool); col = np.zeros((steps, n_envs), bool)
    win = np.zeros((steps, n_envs), np.uint8)
    eplen = np.zeros((steps, n_envs), np.int32); epret = np.zeros((steps, n_envs, 2))
    with quiet():
        for e in range(n_envs):
            env.reset()
            n = 0
            for t in range(steps):
                o, r, d, info = env.step(int(acts[t, e, 0]), int(acts[t, e, 1]))
                n += 1
                rew[t, e], done[t, e], col[t, e] = _f(r), d, info["collision"]
                win[t, e] = env.winner or 0
                term[t, e] = _f(o)
                if d:
                    eplen[t, e] = n
                    epret[t, e] = (env.r1_accumulate, env.r2_accumulate)
                    o = env.reset()
                    n = 0
                obs[t, e] = _f(o)
    return dict(actions=acts, obs=obs, step_obs=term, rewards=rew, done=done, collision=col,
                winner=win, ep_len=eplen, ep_ret=epret)


def kat(env, script1, script2, max_steps=2700):
    """script_i: constant int, None, or a list cycled over steps."""
    def act(s, t):
        if s is None or isinstance(s, int):
            return s
        return s[t % len(s)]
    env.reset()
    with quiet():
        for t in range(max_steps):
            o, r, d, info = env.step(act(script1, t), act(script2, t))
            if d:
                break
    return dict(a1=script1, a2=script2, steps=t + 1, winner=env.winner, collision=info["collision"],
                R1=float(env.r1_accumulate), R2=float(env.r2_accumulate),
                pos1=float(env.state1['pos']), pos2=float(env.state2['pos']),
                vel1=float(env.state1['vel']), vel2=float(env.state2['vel']),
                time_stamp=float(env.time_stamp),
                last_obs=[float(v) for v in o], last_rewards=[float(v) for v in r])


def main():
    os.makedirs(OUT, exist_ok=True)
    env = load_reference_env()
    np.savez_compressed(os.path.join(OUT, "config1_pve_trace.npz"), **trace(env, False, 0, 10000))
    np.savez_compressed(os.path.join(OUT, "pvp_trace.npz"), **trace(env, True, 1, 6000))
    np.savez_compressed(os.path.join(OUT, "pvp_vec16_trace.npz"), **vec_trace(env, 16, 640, 2))
    scripts = [(2, None), (2, 2), (3, None), (4, None), (1, None), (0, None), (0, 0), (4, 0),
               (4, 4), (3, 1), (1, 3), (4, 3), (3, 4), ([0, 4], 2), (0, 4), (2, 3), (3, 2),
               ([4, 4, 0], [0, 4, 4]), (1, 1), (3, 3)]
    kats = [kat(env, a, b) for a, b in scripts]
    kats.append(dict(reset_obs=[float(v) for v in env.reset()]))
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kats, f, indent=1)
    for k in kats[:-1]:
        print(k["a1"], k["a2"], k["steps"], k["winner"], k["collision"], k["R1"], k["R2"], k["pos1"], k["pos2"])


def injected(env, m, seed):
    """One reference step from each of m injected states.  `winner` before the step is what the reference's
    own logic implies for the injected positions: a car already past END_POINT can only be there as the winner
    (both past it would have ended the episode), so states with both cars past END_POINT are not generated."""
    rng = np.random.default_rng(seed)
    p1 = rng.uniform(0.0, 1100.0, m); p2 = rng.uniform(0.0, 1100.0, m)
    k = m // 2
    p1[:k] = rng.uniform(900.0, 1010.0, k); p2[:k] = p1[:k] + rng.normal(0.0, 6.0, k)
    q = m // 8
    p1[k:k + q] = 950.0 - rng.uniform(0.0, 9.0, q); p2[k:k + q] = 950.0 - rng.uniform(0.0, 9.0, q)
    v1 = rng.uniform(0.0, 45.0, m); v2 = rng.uniform(0.0, 45.0, m)
    v1[::7] = 0.0; v2[3::11] = 0.0
    both = (p1 > 950.0) & (p2 >= 950.0)
    p2[both] = rng.uniform(800.0, 949.0, int(both.sum()))
    pvp = rng.random(m) < 0.75
    acts = rng.integers(0, 5, (m, 2)).astype(np.uint8)
    win0 = np.where(p1 > 950.0, 1, np.where(p2 >= 950.0, 2, 0)).astype(np.uint8)
    out = dict(pos=np.stack([p1, v1, p2, v2], 1), pvp=pvp, actions=acts, winner_before=win0,
               obs=np.zeros((m, 10)), rewards=np.zeros((m, 2)), done=np.zeros(m, bool), collision=np.zeros(m, bool),
               winner=np.zeros(m, np.uint8), state_after=np.zeros((m, 4)))
    with quiet():
        for i in range(m):
            env.reset()
            env.state1['pos'], env.state1['vel'] = float(p1[i]), float(v1[i])
            env.state2['pos'], env.state2['vel'] = float(p2[i]), float(v2[i])
            env.winner = int(win0[i]) or None
            o, r, d, info = env.step(int(acts[i, 0]), int(acts[i, 1]) if pvp[i] else None)
            out["obs"][i], out["rewards"][i], out["done"][i], out["collision"][i] = _f(o), _f(r), d, info["collision"]
            out["winner"][i] = env.winner or 0
            out["state_after"][i] = (env.state1['pos'], env.state1['vel'], env.state2['pos'], env.state2['vel'])
    return out


if __name__ == "__main__" and "--injected" in sys.argv:
    _inj = injected(load_reference_env(), 6000, 11)
    import warnings
    with warnings.catch_warnings(), quiet():
        warnings.simplefilter("ignore")
        import hdqn as _ref_hdqn                    # the reference's goal_status (scripts/hdqn.py:223-236) on those observations
    _inj["goal_status"] = np.array([_ref_hdqn.goal_status(list(o)) for o in _inj["obs"]], np.uint8)
    np.savez_compressed(os.path.join(OUT, "injected_states.npz"), **_inj)
elif __name__ == "__main__" and not {"--policies", "--hdqn", "--replay"} & set(sys.argv):
    main()


# ------------------------------------------------------------------------------------------------
# Policy-in-the-loop fixtures: the reference's shipped DQN checkpoints
# (test_params/dqn/*/eval.pth; Net = 10 -> 200 -> 100 -> 5, scripts/main.py:30-47) played greedily
# against the constant-speed "L0" opponent (action2=None, main.py:196-197) in the unmodified env.
def policy_fixtures():
    import glob
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class Net(nn.Module):                      # same layer names as main.py:30-47 so state_dicts load
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Linear(10, 200); self.fc2 = nn.Linear(200, 100); self.out = nn.Linear(100, 5)

        def forward(self, x):
            return self.out(F.relu(self.fc2(F.relu(self.fc1(x)))))

    env = load_reference_env()
    out = {}
    root = os.path.join(os.environ.get("MERGING_GYM_REFERENCE", "/root/reference"), "test_params", "dqn")
    for tag, prefix in (("L1_1445", "2022--03--31 14:45:59"), ("L0_2037", "2022--03--31 20:37:39")):
        d = [p for p in sorted(glob.glob(os.path.join(root, "*"))) if os.path.basename(p).startswith(prefix)][0]
        sd = torch.load(os.path.join(d, "eval.pth"), map_location="cpu", weights_only=True)
        net = Net(); net.load_state_dict(sd); net.eval()
        for k, v in sd.items():
            out[f"{tag}/{k}"] = v.numpy().astype(np.float32)
        state = env.reset()
        acts, obs, qs = [], [], []
        with quiet(), torch.no_grad():
            while True:
                q = net(torch.FloatTensor(state).unsqueeze(0))
                a = int(torch.max(q, 1)[1][0])                      # main.py:105
                obs.append(_f(state)); qs.append(q[0].numpy().astype(np.float64)); acts.append(a)
                state, r, done, info = env.step(a, None)
                if done:
                    break
        out[f"{tag}/traj_obs"] = np.array(obs); out[f"{tag}/traj_q"] = np.array(qs)
        out[f"{tag}/traj_actions"] = np.array(acts, np.uint8)
        out[f"{tag}/result"] = np.array([len(acts), env.winner or 0, int(info["collision"]),
                                         env.r1_accumulate, env.r2_accumulate])
        print(tag, len(acts), env.winner, info, env.r1_accumulate, env.r2_accumulate)
    # pvp: the "OP:L2" agent (21:33:10) against the second "OP:L1" net (21:36:59) as opponent, which sees the
    # mirrored observation `state[NUM_STATES//2:] + state[:NUM_STATES//2]` (main.py:196-199); both greedy.
    # 432 steps: the agent holds action 4 and wins, the opponent switches between actions 0, 1 and 2.
    nets = {}
    for tag, prefix in (("L2_2133", "2022--03--31 21:33:10"), ("L1_2136", "2022--03--31 21:36:59")):
        d = [p for p in sorted(glob.glob(os.path.join(root, "*"))) if os.path.basename(p).startswith(prefix)][0]
        sd = torch.load(os.path.join(d, "eval.pth"), map_location="cpu", weights_only=True)
        net = Net(); net.load_state_dict(sd); nets[tag] = net.eval()
        for k, v in sd.items():
            out[f"{tag}/{k}"] = v.numpy().astype(np.float32)
    state = env.reset()
    acts, obs = [], []
    with quiet(), torch.no_grad():
        while True:
            a = int(torch.max(nets["L2_2133"](torch.FloatTensor(state).unsqueeze(0)), 1)[1][0])
            mirrored = state[5:] + state[:5]
            b = int(torch.max(nets["L1_2136"](torch.FloatTensor(mirrored).unsqueeze(0)), 1)[1][0])
            obs.append(_f(state)); acts.append((a, b))
            state, r, done, info = env.step(a, b)
            if done:
                break
    out["pvp_L2_vs_L1/traj_obs"] = np.array(obs)
    out["pvp_L2_vs_L1/traj_actions"] = np.array(acts, np.uint8)
    out["pvp_L2_vs_L1/result"] = np.array([len(acts), env.winner or 0, int(info["collision"]),
                                           env.r1_accumulate, env.r2_accumulate])
    print("pvp_L2_vs_L1", len(acts), env.winner, info, env.r1_accumulate, env.r2_accumulate)
    np.savez_compressed(os.path.join(OUT, "dqn_policies.npz"), **out)


# h-DQN fixtures: the reference's OWN classes (scripts/hdqn.py: Net, Goal_DQN.choose_goal, HDQN.choose_action)
# with their own initialisation under a fixed torch seed, made greedy by pinning the `np.random.randn() <= EPISILO`
# draw, driven by the episode loop of hdqn.py:276-312 (goal from the state, action from [goal] + state, goal
# re-chosen from next_state after every step; the opponent on the mirrored observation in self-play).
def hdqn_fixtures():
    import torch
    env = load_reference_env()                      # also puts the shims and scripts/ on sys.path
    import warnings
    with warnings.catch_warnings(), quiet():
        warnings.simplefilter("ignore")
        import hdqn as ref                          # the unmodified scripts/hdqn.py (creates its own env at import)
    ref.USE_CUDA = False
    ref.np.random.randn = lambda *a: -1e9           # `np.random.randn() <= EPISILO` is always true: greedy
    out = {}
    for seed in (7, 36):       # 7: the controller switches action mid-episode; 36: the meta-controller switches goal
        torch.manual_seed(seed)
        upper, lower = ref.Goal_DQN(None), ref.HDQN(None)
        tag = f"seed{seed}"
        for name, net in (("meta", upper.meta_eval_net), ("ctrl", lower.eval_net)):
            for k, v in net.state_dict().items():
                out[f"{tag}/{name}/{k}"] = v.numpy().astype(np.float32)
        for mode in ("L0", "selfplay"):
            state = env.reset()
            rows, obs = [], []
            done = False
            with quiet(), torch.no_grad():
                goal = upper.choose_goal(state)
                goal_op = upper.choose_goal(state[5:] + state[:5]) if mode == "selfplay" else 0
                while not done:
                    gs = torch.unsqueeze(torch.FloatTensor([goal] + state), dim=0)
                    action = int(lower.choose_action(gs))
                    action_op = None
                    if mode == "selfplay":
                        gso = torch.unsqueeze(torch.FloatTensor([goal_op] + state[5:] + state[:5]), dim=0)
                        action_op = int(lower.choose_action(gso))
                    obs.append(_f(state)); rows.append((goal, action, goal_op, action_op or 0))
                    state, rewards, done, info = env.step(action, action_op)
                    goal = upper.choose_goal(state)
                    if mode == "selfplay":
                        goal_op = upper.choose_goal(state[5:] + state[:5])
            if mode == "L0":
                # the same episode once more through the reference's nested option loop (hdqn.py:276-324) with its own
                # store_transition calls: the controller's memory rows [g, s, a, r_int, g', s'] (24 floats)
                lower.memory_counter = 0
                upper.memory_counter = 0
                state = env.reset(); done = False
                with quiet(), torch.no_grad():
                    while not done:
                        goal = upper.choose_goal(state)
                        extrinsic_reward = 0
                        while not done:
                            goal_state = torch.unsqueeze(torch.FloatTensor([goal] + state), dim=0)
                            action = lower.choose_action(goal_state)
                            next_state, rewards, done, info = env.step(action, None)
                            goal = upper.choose_goal(next_state)
                            next_goal_state = torch.unsqueeze(torch.FloatTensor([goal] + next_state), dim=0)
                            extrinsic_reward += rewards[0]
                            intrinsic_reward = 1.0 if goal == ref.goal_status(state) else 0.0
                            lower.store_transition(goal_state, action, intrinsic_reward, next_goal_state)
                            state = next_state
                            if done or goal == ref.goal_status(state):
                                break
                        upper.store_transition(state, goal, extrinsic_reward, next_state)      # hdqn.py:318
                assert lower.memory_counter == len(rows) and upper.memory_counter < ref.GOAL_MEMORY_CAPACITY
                out[f"{tag}/L0/controller_memory"] = lower.memory[:lower.memory_counter].copy()
                out[f"{tag}/L0/meta_memory"] = upper.memory[:upper.memory_counter].copy()
            out[f"{tag}/{mode}/traj_obs"] = np.array(obs)
            out[f"{tag}/{mode}/traj"] = np.array(rows, np.uint8)          # goal, action, goal_op, action_op
            out[f"{tag}/{mode}/result"] = np.array([len(rows), env.winner or 0, int(info["collision"]),
                                                    env.r1_accumulate, env.r2_accumulate])
            r = np.array(rows)
            print(tag, mode, len(rows), env.winner, info, [np.bincount(r[:, c], minlength=5).tolist() for c in range(4)])
    np.savez_compressed(os.path.join(OUT, "hdqn_policies.npz"), **out)


# Replay-memory fixture: the reference's OWN learner class (scripts/main.py `DQN`: choose_action with its
# epsilon rule under a fixed NumPy seed, store_transition) in the episode loop of main.py:190-221 — the agent
# loaded from the "OP:L1" checkpoint, the opponent from the "OP:L0" one acting on the mirrored observation —
# for three episodes (no learn(): the memory stays below MEMORY_CAPACITY).  Records both players' actions per
# step and the resulting `dqn.memory` rows.
def replay_fixture():
    import glob
    import torch
    env = load_reference_env()
    import warnings
    with warnings.catch_warnings(), quiet():
        warnings.simplefilter("ignore")
        import main as ref                          # the unmodified scripts/main.py
    ref.USE_CUDA = False
    root = os.path.join(os.environ.get("MERGING_GYM_REFERENCE", "/root/reference"), "test_params", "dqn")
    find = lambda prefix: [p for p in sorted(glob.glob(os.path.join(root, "*"))) if os.path.basename(p).startswith(prefix)][0]
    _load = torch.load
    torch.load = lambda f, *a, **k: _load(f, map_location="cpu", weights_only=True)     # checkpoints were saved from CUDA
    dqn, opponent = ref.DQN(find("2022--03--31 14:45:59")), ref.DQN(find("2022--03--31 03:37:35"))
    torch.load = _load
    np.random.seed(3)
    acts, dones = [], []
    with quiet(), torch.no_grad():
        for ep in range(3):
            state = env.reset()
            while True:
                action = int(dqn.choose_action(state))
                action_op = int(opponent.choose_action(state[5:] + state[:5]))
                next_state, rewards, done, info = env.step(action, action_op)
                if env.winner is not 1:
                    dqn.store_transition(state, action, rewards[0], next_state)
                acts.append((action, action_op)); dones.append(done)
                if done:
                    break
                state = next_state
    c = dqn.memory_counter
    assert c < ref.MEMORY_CAPACITY
    a = np.array(acts, np.uint8)
    print("replay fixture:", len(acts), "steps,", c, "rows stored, action histograms",
          np.bincount(a[:, 0], minlength=5).tolist(), np.bincount(a[:, 1], minlength=5).tolist())
    np.savez_compressed(os.path.join(OUT, "replay_memory.npz"), actions=a, done=np.array(dones),
                        memory=dqn.memory[:c].copy(), counter=np.array(c))


if __name__ == "__main__" and "--replay" in sys.argv:
    replay_fixture()
if __name__ == "__main__" and "--hdqn" in sys.argv:
    hdqn_fixtures()
if __name__ == "__main__" and "--policies" in sys.argv:
    policy_fixtures()
