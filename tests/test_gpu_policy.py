"""Policy-in-the-loop: the fused Q-network kernel `mg_mlp_act` against a plain
PyTorch reference of the same op (fp32 on the GPU, fp64 on the CPU), and the reference's shipped
DQN checkpoints played greedily in the env against the recorded reference-env episodes."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import merging_gym_b200
    return merging_gym_b200


@pytest.fixture(scope="module")
def ckpt():
    z = np.load(os.path.join(GOLDEN, "dqn_policies.npz"))
    def get(tag):
        sd = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/") and "traj" not in k and "result" not in k}
        return sd, {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/traj") or k.endswith("result") and k.startswith(tag)}
    return get


def mid_episode_obs(mg, n, steps=120, seed=3):
    env = mg.MergeVecEnv(n, seed=seed)
    env.rollout(steps)
    a1, a2 = env.sample_actions()
    return env.step(a1, a2)[0].clone()


def q_fp64(policy, x):
    sd = {k: v.double().cpu() for k, v in policy.state_dict().items()}
    h = torch.relu(x @ sd["fc1.weight"].t() + sd["fc1.bias"])
    h = torch.relu(h @ sd["fc2.weight"].t() + sd["fc2.bias"])
    return h @ sd["out.weight"].t() + sd["out.bias"]


@pytest.mark.parametrize("in_dim,out_dim,n", [(10, 5, 5000), (10, 3, 4099), (11, 5, 1031), (10, 5, 3), (11, 3, 256)])
def test_fused_mlp_matches_torch_reference(mg, in_dim, out_dim, n):
    obs = mid_episode_obs(mg, n)
    goal = torch.randint(0, 3, (n,), dtype=torch.uint8, device="cuda") if in_dim == 11 else None
    pol = mg.MLPPolicy(in_dim, out_dim, seed=in_dim * 10 + out_dim)
    ref = mg.MLPPolicy(in_dim, out_dim, state_dict=pol.state_dict(), backend="torch")
    q = torch.empty(n, out_dim, device="cuda")
    a = pol.act(obs, goal=goal, q_out=q)
    x = obs if goal is None else torch.cat([goal.float().unsqueeze(1), obs], 1)
    q64 = q_fp64(pol, x.double().cpu())
    q32 = ref.q_values_torch(obs, goal)
    scale = q64.abs().max().item()
    assert (q.double().cpu() - q64).abs().max().item() <= 2e-5 * scale      # fp32 accumulation noise
    assert (q32.double().cpu() - q64).abs().max().item() <= 2e-5 * scale
    # arg-max: identical wherever the top-2 margin is above the fp32 noise floor; first maximum on ties
    top2 = q64.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * scale
    assert clear.float().mean() > 0.9
    assert torch.equal(a.cpu()[clear].long(), q64.argmax(1)[clear])
    assert torch.equal(a.long(), q.argmax(1))                               # consistent with its own Q
    assert torch.equal(ref.act(obs, goal=goal).cpu()[clear], a.cpu()[clear])


@pytest.mark.parametrize("tc_backend", ["tf32x3", "f16x3"])
@pytest.mark.parametrize("in_dim,out_dim,n", [(10, 5, 5000), (10, 3, 129), (11, 5, 1031), (11, 3, 4096), (10, 5, 1), (10, 5, 148 * 128 * 3 + 77)])
def test_tensor_core_backend_matches_fp32(mg, in_dim, out_dim, n, tc_backend):
    """`backend="tf32x3"` (tcgen05 + TMEM, error-compensated 3xTF32 for the 200x100 layer) and `backend="f16x3"` (both
    hidden layers on the tensor cores as three-product sums of fp16 hi / lo operands): Q-values within fp32-level error
    of the fp64 reference and the same actions wherever the margin is clear."""
    obs = mid_episode_obs(mg, n, seed=in_dim + n)
    goal = torch.randint(0, 3, (n,), dtype=torch.uint8, device="cuda") if in_dim == 11 else None
    f = mg.MLPPolicy(in_dim, out_dim, seed=3 * in_dim + out_dim)
    tc = mg.MLPPolicy(in_dim, out_dim, state_dict=f.state_dict(), backend=tc_backend)
    qf = torch.empty(n, out_dim, device="cuda"); qt = torch.empty(n, out_dim, device="cuda")
    af, at = f.act(obs, goal=goal, q_out=qf), tc.act(obs, goal=goal, q_out=qt)
    x = obs if goal is None else torch.cat([goal.float().unsqueeze(1), obs], 1)
    q64 = q_fp64(f, x.double().cpu())
    scale = q64.abs().max().item()
    assert (qt.double().cpu() - q64).abs().max().item() <= 2e-5 * scale      # measured 3e-6 (both); fp32 FFMA: 2e-7
    top2 = q64.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4 * scale
    assert torch.equal(at.cpu()[clear].long(), q64.argmax(1)[clear])
    assert torch.equal(at.long(), qt.argmax(1))
    assert (af == at).float().mean().item() > 0.999


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
@pytest.mark.parametrize("in_dim,out_dim", [(10, 5), (11, 5), (10, 3)])
def test_mirrored_read_equals_opponent_view(mg, backend, in_dim, out_dim):
    """`mirror=True` (MG_MLP_FLAG_MIRROR): the kernels read every row as the opponent sees it, `state[5:] + state[:5]`
    (main.py:199; with a goal column: `[goal_op] + state[5:] + state[:5]`, hdqn.py:299) — bit-identical to running
    them on a materialised `env.opponent_view(obs)`."""
    n = 3001
    env = mg.MergeVecEnv(n, seed=2); env.rollout(130)
    obs = env.step(*env.sample_actions())[0].clone()
    goal = torch.randint(0, 3, (n,), dtype=torch.uint8, device="cuda") if in_dim == 11 else None
    pol = mg.MLPPolicy(in_dim, out_dim, seed=9, backend=backend)
    q1 = torch.empty(n, out_dim, device="cuda"); q2 = torch.empty(n, out_dim, device="cuda")
    a1 = pol.act(obs, goal=goal, q_out=q1, mirror=True).clone()
    a2 = pol.act(env.opponent_view(obs).contiguous(), goal=goal, q_out=q2).clone()
    assert torch.equal(a1, a2) and torch.equal(q1, q2)
    pol.act(obs, goal=goal, q_out=q2)
    assert not torch.equal(q1, q2)                                              # and it differs from the unmirrored read


@pytest.mark.parametrize("tc_backend", ["tf32x3", "f16x3"])
def test_tensor_core_backend_is_stable_under_repetition(mg, tc_backend):
    """The tensor-core kernel is a multi-role pipeline (8 producer warps, an MMA-issuing warp, 4 epilogue warps
    that hand the accumulator buffers back zeroed, 20 mbarriers).  A protocol race would show up as an occasional lost
    update or a stale tile: 150 launches over many tiles per SM (2^17 envs = 1024 tiles on 148 SMs) on changing
    inputs must all agree with the fp32 kernel — and a repeated launch must reproduce its Q-values BIT FOR BIT (the
    accumulation order in tensor memory is the K-step order; with two issuing warps, MG_TC_MMA_WARPS=2, it is not)."""
    n = 1 << 17
    env = mg.MergeVecEnv(n, seed=11)
    f = mg.MLPPolicy(10, 5, seed=5)
    tc = mg.MLPPolicy(10, 5, state_dict=f.state_dict(), backend=tc_backend)
    qf = torch.empty(n, 5, device="cuda"); qt = torch.empty(n, 5, device="cuda")
    worst = torch.zeros((), device="cuda")
    for t in range(150):
        obs = env.step(*env.sample_actions())[0]
        f.act(obs, q_out=qf); tc.act(obs, q_out=qt)
        worst = torch.maximum(worst, (qf - qt).abs().max() / qf.abs().max())
        if t % 10 == 0:
            q2 = torch.empty_like(qt); tc.act(obs, q_out=q2)
            assert torch.equal(q2, qt), t
    assert worst.item() < 5e-5                                               # measured 3e-6


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
@pytest.mark.parametrize("tag", ["L1_1445", "L0_2037"])
def test_shipped_dqn_checkpoint_greedy_vs_L0(mg, ckpt, tag, backend):
    """Greedy DQN (test_params/dqn/*/eval.pth) vs the constant-speed opponent: 225 steps, P1 wins,
    no collision, R1 = 0.589997, R2 = 1.0 — the episode recorded in the unmodified reference env."""
    sd, traj = ckpt(tag)
    pol = mg.MLPPolicy(10, 5, state_dict=sd, backend=backend)
    env = mg.MergeVecEnv(64, mode="pve", auto_reset=False)
    obs = env.reset()
    q = torch.empty(64, 5, device="cuda")
    T = len(traj["traj_actions"])
    for t in range(T):
        a = pol.act(obs, q_out=q)
        assert a.cpu().tolist() == [int(traj["traj_actions"][t])] * 64, t
        qr = traj["traj_q"][t]                                       # torch CPU fp32 forward in the reference loop
        # summation order differs and the hidden activations are O(100) while these Q-values are O(1)
        # (cancellation): loose on Q — looser still for the 3xTF32 backend — but exact on the action
        q_tol = 1e-4 if backend == "fused" else 1e-3
        assert np.abs(q[0].cpu().numpy() - qr).max() <= q_tol * max(1.0, np.abs(qr).max()), t
        assert rel_err(obs[0].cpu().numpy(), traj["traj_obs"][t]).max() <= 1e-5, t
        obs, rew, done, info = env.step(a, None)
    steps, winner, col, R1, R2 = traj["result"]
    assert T == steps == 225 and bool(done.all()) and not bool(info["collision"].any())
    assert env.winner.cpu().tolist() == [int(winner)] * 64 == [1] * 64
    assert abs(float(env.ret1[0]) - R1) <= 1e-9 and abs(float(env.ret2[0]) - R2) <= 1e-9
    assert abs(R1 - 0.5899968243808502) < 1e-12 and R2 == 1.0


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
def test_two_shipped_checkpoints_play_each_other(mg, ckpt, backend):
    """pvp with a policy on both sides, the opponent acting on the mirrored observation
    `state[5:] + state[:5]` (main.py:196-199 = `env.opponent_view`): the "OP:L2" checkpoint against the
    second "OP:L1" checkpoint, both greedy, in the unmodified reference env — 432 steps, the agent
    holds action 4 and wins, the opponent switches between actions 0, 1 and 2."""
    sd1, _ = ckpt("L2_2133"); sd2, _ = ckpt("L1_2136"); _, traj = ckpt("pvp_L2_vs_L1")
    p1 = mg.MLPPolicy(10, 5, state_dict=sd1, backend=backend)
    p2 = mg.MLPPolicy(10, 5, state_dict=sd2, backend=backend)
    env = mg.MergeVecEnv(64, mode="pvp", auto_reset=False)
    obs = env.reset()
    T = len(traj["traj_actions"])
    assert len(set(traj["traj_actions"][:, 1].tolist())) == 3
    for t in range(T):
        a1, a2 = p1.act(obs), (p2.act(obs, mirror=True) if t % 2 else p2.act(env.opponent_view(obs)))
        assert a1.cpu().tolist() == [int(traj["traj_actions"][t, 0])] * 64, t
        assert a2.cpu().tolist() == [int(traj["traj_actions"][t, 1])] * 64, t
        assert rel_err(obs[0].cpu().numpy(), traj["traj_obs"][t]).max() <= 1e-5, t
        obs, rew, done, info = env.step(a1, a2)
    steps, winner, col, R1, R2 = traj["result"]
    assert T == steps == 432 and bool(done.all()) and not bool(info["collision"].any()) and not col
    assert env.winner.cpu().tolist() == [int(winner)] * 64 == [1] * 64
    assert abs(float(env.ret1[0]) - R1) <= 1e-9 and abs(float(env.ret2[0]) - R2) <= 1e-9


def test_policy_in_the_loop_cuda_graph(mg, ckpt):
    """obs -> fused MLP -> uint8 actions -> mg_step, captured in one CUDA graph: no host sync, and
    the replay equals the eager loop."""
    sd, _ = ckpt("L1_1445")
    n, K = 4096, 32
    pol = mg.MLPPolicy(10, 5, state_dict=sd)
    opp = mg.MLPPolicy(10, 5, state_dict=ckpt("L0_2037")[0])

    def run(env, graph):
        act1 = torch.empty(n, dtype=torch.uint8, device="cuda"); act2 = torch.empty_like(act1)
        obs_box = [env.obs_buf[env._slot]]
        def loop():
            obs = obs_box[0]
            for _ in range(K):
                pol.act(obs, out=act1)
                opp.act(mg.MergeVecEnv.opponent_view(obs).contiguous(), out=act2)   # main.py:199
                obs = env.step(act1, act2)[0]
            return obs
        if not graph:
            return loop().clone()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            pol.act(obs_box[0], out=act1)                           # warm-up (sets the smem attribute)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = loop()
        g.replay()
        torch.cuda.synchronize()
        return out.clone()

    a = mg.MergeVecEnv(n, seed=11); a.rollout(100)
    b = mg.MergeVecEnv(n, seed=11); b.rollout(100)
    a.step(*a.sample_actions()); b.step(*b.sample_actions())
    oa, ob = run(a, False), run(b, True)
    assert torch.equal(oa, ob) and torch.equal(a.pos1, b.pos1) and torch.equal(a.meta, b.meta)


def test_explore_rule_and_goal_status(mg):
    g = torch.Generator(device="cuda").manual_seed(0)
    greedy = torch.full((400000,), 3, dtype=torch.uint8, device="cuda")
    a = mg.explore(greedy, 5, generator=g)
    kept = (a == 3).float().mean().item()
    # P(randn <= 0.7) = 0.758, plus 1/5 of the random draws land on 3 again
    assert abs(kept - (0.7580 + 0.2420 / 5)) < 0.005
    obs = torch.zeros(4, 10, device="cuda")
    obs[:, 0] = torch.tensor([-20.0, -5.0, 5.0, 20.0]); obs[:, 9] = 20.0
    assert mg.goal_status(obs).cpu().tolist() == [0, 1, 1, 2]          # hdqn.py:223-236
    # ... and on the 6 000 observations of the injected-state fixture, labelled by the reference's own function
    z = np.load(os.path.join(GOLDEN, "injected_states.npz"))
    got = mg.goal_status(torch.from_numpy(z["obs"]).float().cuda()).cpu().numpy()
    exact = z["goal_status"]
    knife = np.abs(np.abs(z["obs"][:, 0]) - 0.5 * z["obs"][:, 9]) < 1e-4 * np.maximum(1.0, np.abs(z["obs"][:, 0]))   # fp32 vs fp64 at the threshold
    assert np.array_equal(got[~knife], exact[~knife]) and knife.mean() < 0.01 and len(set(exact.tolist())) == 3


@pytest.mark.parametrize("backend", ["fused", "tf32x3", "f16x3"])
@pytest.mark.parametrize("seed,mode", [(7, "L0"), (7, "selfplay"), (36, "L0"), (36, "selfplay")])
def test_hdqn_against_the_reference_classes(mg, seed, mode, backend):
    """tests/golden/hdqn_policies.npz: episodes played by the reference's OWN h-DQN classes (scripts/hdqn.py
    `Goal_DQN.choose_goal`, `HDQN.choose_action`, their initialisation under a fixed torch seed, greedy) in
    the unmodified env — against the constant-speed opponent and in self-play on the mirrored observation.
    `HDQNPolicy` must pick the same goal and the same action at every step (seed 7: the controller switches
    action mid-episode; seed 36: the meta-controller switches goal)."""
    z = np.load(os.path.join(GOLDEN, "hdqn_policies.npz"))
    tag = f"seed{seed}"
    sd = lambda name: {k.split("/")[-1]: z[k] for k in z.files if k.startswith(f"{tag}/{name}/")}
    traj, tobs, result = z[f"{tag}/{mode}/traj"], z[f"{tag}/{mode}/traj_obs"], z[f"{tag}/{mode}/result"]
    pol = mg.HDQNPolicy(meta_state=sd("meta"), ctrl_state=sd("ctrl"), backend=backend)
    pvp = mode == "selfplay"
    env = mg.MergeVecEnv(32, mode="pvp" if pvp else "pve", auto_reset=False)
    obs = env.reset()
    for t in range(len(traj)):
        g, a, g_op, a_op = (int(v) for v in traj[t])
        a1 = pol.act(obs).clone(); g1 = pol.goal.clone()
        assert g1.cpu().tolist() == [g] * 32 and a1.cpu().tolist() == [a] * 32, t
        a2 = None
        if pvp:
            a2 = pol.act(env.opponent_view(obs)).clone()
            assert pol.goal.cpu().tolist() == [g_op] * 32 and a2.cpu().tolist() == [a_op] * 32, t
        assert rel_err(obs[0].cpu().numpy(), tobs[t]).max() <= 1e-5, t
        obs, rew, done, info = env.step(a1, a2)
    steps, winner, col, R1, R2 = result
    assert len(traj) == steps and bool(done.all()) and bool(info["collision"].all()) == bool(col)
    assert env.winner.cpu().tolist() == [int(winner)] * 32
    assert abs(float(env.ret1[0]) - R1) <= 1e-9 * max(1.0, abs(R1)) and abs(float(env.ret2[0]) - R2) <= 1e-9 * max(1.0, abs(R2))


def test_hdqn_policy_fused_equals_torch(mg):
    n = 3000
    obs = mid_episode_obs(mg, n, seed=9)
    f = mg.HDQNPolicy(seed=5)
    t = mg.HDQNPolicy(seed=5, backend="torch")
    af, at = f.act(obs), t.act(obs)
    assert (f.goal == t.goal).float().mean() > 0.98
    same_goal = f.goal == t.goal
    assert ((af == at) | ~same_goal).float().mean() > 0.98
    env = mg.MergeVecEnv(n, mode="pve")
    env.step(af, None)


@pytest.mark.parametrize("script,args", [("random_rollout.py", ["--envs", "65536", "--steps", "300"]),
                                         ("dqn_vs_dqn.py", ["--steps", "400"]), ("dqn_vs_dqn.py", ["--steps", "400", "--backend", "tf32x3"]),
                                         ("dqn_vs_dqn.py", ["--steps", "400", "--backend", "f16x3"])])
def test_examples_run(script, args):
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", script), *args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "episodes" in out.stdout


@pytest.mark.parametrize("pvp", [False, True])
def test_graphed_policy_rollout_equals_the_eager_loop(mg, ckpt, pvp):
    """`GraphedPolicyRollout`: K steps of policy -> env.step -> recorder.record captured in one CUDA graph and
    replayed; state, statistics and the recorded replay rows equal the same loop run eagerly."""
    n, K, R = 2048, 8, 40
    sd1, _ = ckpt("L2_2133"); sd2, _ = ckpt("L1_2136")
    def build():
        env = mg.MergeVecEnv(n, mode="pvp" if pvp else "pve", out_slots=1, seed=1)
        env.rollout(150)                                        # de-synchronise the envs first
        env.step(env.sample_actions()[0], env.sample_actions()[1] if pvp else None)
        p1 = mg.MLPPolicy(10, 5, state_dict=sd1); p2 = mg.MLPPolicy(10, 5, state_dict=sd2) if pvp else None
        rec = mg.TransitionRecorder(env, 4 * n * K, track_env_ids=True)
        return env, p1, p2, rec
    env, p1, p2, rec = build()
    roll = mg.GraphedPolicyRollout(env, p1, (lambda o: p2.act(o, mirror=True)) if pvp else None, k_steps=K,
                                   after_step=rec.record, warmup_steps=0)
    out = roll.run(R)
    ref, q1, q2, rrec = build()
    for t in range(K * R):
        obs = ref.obs_buf[0].clone()
        a1 = q1.act(obs); a2 = q2.act(obs, mirror=True) if pvp else None
        rout = ref.step(a1, a2)
        rrec.record(obs, a1, a2, rout)
    torch.cuda.synchronize()
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        assert torch.equal(getattr(env, k), getattr(ref, k)), k
    assert env.stats() == ref.stats() and env.stats()["episodes"] > n
    assert torch.equal(out[0], rout[0]) and torch.equal(rec.counter, rrec.counter)
    c = min(int(rec.counter.item()), rec.capacity)
    assert torch.equal(rec.ring[:c], rrec.ring[:c]) and torch.equal(rec.env_ids[:c], rrec.env_ids[:c])
