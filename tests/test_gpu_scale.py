"""Size-independent properties at BASELINE.json's full sizes (1M envs), where the oracle would
take too long to compare every step: determinism, shard invariance (global env ids), statistics
consistency, and agreement of a sub-sample with the oracle."""
import numpy as np
import pytest
import torch

from oracle import merge_oracle as mo
from conftest import rel_err

pytestmark = pytest.mark.gpu
N = 1 << 20
K = 256


@pytest.fixture(scope="module")
def mg():
    import merging_gym_b200
    return merging_gym_b200


@pytest.fixture(scope="module")
def full_run(mg):
    env = mg.MergeVecEnv(N, seed=0x5EED)
    done_count = torch.zeros((), dtype=torch.int64, device="cuda")
    for t in range(K):
        a1, a2 = env.sample_actions()
        _, _, done, _ = env.step(a1, a2)
        done_count += done.sum()
    torch.cuda.synchronize()
    return env, int(done_count)


def test_statistics_are_consistent(full_run):
    env, done_count = full_run
    s = env.stats()
    assert s["episodes"] == done_count > 0
    assert s["collisions"] + s["merges_ok"] + s["timeouts"] >= s["episodes"]
    assert s["merges_ok"] == s["episodes"] - s["collisions"]          # no timeouts within 256 steps
    assert s["timeouts"] == 0 and s["bad_actions"] == 0
    assert s["wins_p1"] + s["wins_p2"] <= s["episodes"]
    # SURVEY.md §6 [probe]: random pvp play -> mean length ~210, collision rate ~0.367
    assert 195 < s["mean_length"] < 225 and 0.33 < s["collision_rate"] < 0.40
    assert abs(s["win_rate_p1"] - s["win_rate_p2"]) < 0.01


def test_deterministic_and_rollout_equivalent(mg, full_run):
    env, _ = full_run
    other = mg.MergeVecEnv(N, seed=0x5EED)
    other.rollout(K)                                   # fused K-step launch, same Philox stream
    torch.cuda.synchronize()
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        assert torch.equal(getattr(env, k), getattr(other, k)), k
    assert env.stats() == other.stats()


def test_shard_invariance(mg, full_run):
    """Two half-size shards with global env-id bases reproduce the single-shard run bit for bit
    (what makes the multi-GPU statistics world-size invariant)."""
    env, _ = full_run
    halves = [mg.MergeVecEnv(N // 2, seed=0x5EED, env_id_base=b) for b in (0, N // 2)]
    for h in halves:
        h.rollout(K)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([h.pos1 for h in halves]), env.pos1)
    assert torch.equal(torch.cat([h.ret2 for h in halves]), env.ret2)
    tot = halves[0].stats_tensor() + halves[1].stats_tensor()
    assert torch.equal(tot, env.stats_tensor())


def test_shards_on_separate_streams_match_the_single_stream_run(mg, full_run):
    """bench.py's `overlapped_streams` set-up: independent shards stepped concurrently, each on its own
    CUDA stream (every entry point launches on the caller's current stream; the library keeps no global
    state), must reproduce the single-stream run bit for bit — state, statistics and outputs."""
    env, _ = full_run
    R = 4
    lanes = [torch.cuda.Stream() for _ in range(2)]
    shards = [mg.MergeVecEnv(N // R, seed=0x5EED, env_id_base=r * (N // R)) for r in range(R)]
    for ln in lanes:
        ln.wait_stream(torch.cuda.current_stream())
    last = [None] * R
    for t in range(K):
        for r, sh in enumerate(shards):
            with torch.cuda.stream(lanes[r % 2]):
                last[r] = sh.step(*sh.sample_actions())
    for ln in lanes:
        torch.cuda.current_stream().wait_stream(ln)
    torch.cuda.synchronize()
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        assert torch.equal(torch.cat([getattr(sh, k) for sh in shards]), getattr(env, k)), k
    assert torch.equal(sum(sh.stats_tensor() for sh in shards), env.stats_tensor())
    assert torch.equal(torch.cat([o[0] for o in last]), env.obs_buf[env._slot])


def test_subsample_matches_oracle(mg):
    """Envs [base, base+512) of the 1M-env Philox stream against the oracle, every step."""
    base, n = 777_216, 512
    env = mg.MergeVecEnv(n, seed=0x5EED, env_id_base=base)
    ref = mo.RefVecEnv(n)
    for t in range(K):
        a1, a2 = env.sample_actions()
        obs, rew, done, info = env.step(a1, a2)
        r1, r2 = mo.philox_actions(n, 0x5EED, base, t)
        robs, rrew, rdone, rinfo = ref.step(r1, r2)
        assert np.array_equal(done.cpu().numpy(), rdone) and np.array_equal(info["flags"].cpu().numpy(), rinfo)
        assert rel_err(obs.cpu().numpy(), robs).max() <= 1e-5 and rel_err(rew.cpu().numpy(), rrew).max() <= 1e-5
    assert np.array_equal(env.pos1.cpu().numpy(), ref.pos1)


def test_player_mirror_symmetry(mg):
    """Swapping the two players' action streams mirrors the float64 state bit for bit (the two
    lanes are mirror images, merging_env.py:48-58) and leaves `done` / collisions unchanged.  Rewards
    mirror too, except on steps where both cars cross END_POINT together: player 1 is evaluated
    first and takes RFirst in both runs (merging_env.py:163-181)."""
    n, K = N, 230
    a = mg.MergeVecEnv(n, seed=42, episode_info=False)
    b = mg.MergeVecEnv(n, seed=42, episode_info=False)
    mismatch = torch.zeros((), dtype=torch.int64, device="cuda")
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    for t in range(K):
        a1, a2 = a.sample_actions(t)
        oa, ra, da, ia = a.step(a1, a2)
        ob, rb, db, ib = b.step(a2, a1)
        assert torch.equal(da, db)
        assert torch.equal(ia["collision"], ib["collision"])
        diff = (ra[:, 0] != rb[:, 1]) | (ra[:, 1] != rb[:, 0])
        mismatch += diff.sum()
        # every asymmetric step is a simultaneous crossing: the bonuses 2 and 1 are exchanged
        tie = ((ra[:, 0] - rb[:, 1]).abs() - 1.0).abs() < 1e-6
        bad += (diff & ~tie).sum()
    torch.cuda.synchronize()
    assert torch.equal(a.pos1, b.pos2) and torch.equal(a.vel1, b.vel2)
    assert torch.equal(a.pos2, b.pos1) and torch.equal(a.vel2, b.vel1)
    assert int(bad) == 0
    assert int(mismatch) < 0.01 * n * K
    sa, sb = a.stats(), b.stats()
    assert sa["episodes"] == sb["episodes"] and sa["collisions"] == sb["collisions"]


def test_soak_parity_8e8_env_steps(mg):
    """262 144 envs x 3 000 steps (7.9e8 env-steps, ~3.7e6 episodes) against the plain-C oracle, every
    step: done / info flags bit-exact, rewards and observations within 1e-5, and at the end the float64
    state bit-identical.  A knife-edge trunc()/threshold disagreement anywhere would show up here."""
    from oracle import c_oracle
    n, T = int(__import__("os").environ.get("MG_SOAK_ENVS", 1 << 18)), 3000   # MG_SOAK_ENVS=4194304: the 1.3e10-step run in profiles/README.md
    env = mg.MergeVecEnv(n, seed=2024, episode_info=False)
    ref = c_oracle.CVecEnv(n, nthreads=min(16, max(1, __import__("os").cpu_count() or 1)))
    worst_obs = worst_rew = 0.0
    for t in range(T):
        a1, a2 = env.sample_actions()
        obs, rew, done, info = env.step(a1, a2)
        robs, rrew, rdone, rinfo = ref.step(a1.cpu().numpy(), a2.cpu().numpy())
        assert np.array_equal(info["flags"].cpu().numpy(), rinfo), f"flags differ at step {t}"
        assert np.array_equal(done.cpu().numpy(), rdone)
        if t % 25 == 0:                                  # continuous outputs: every 25th step (D2H volume)
            worst_obs = max(worst_obs, rel_err(obs.cpu().numpy(), robs).max())
            worst_rew = max(worst_rew, rel_err(rew.cpu().numpy(), rrew).max())
    assert worst_obs <= 1e-5 and worst_rew <= 1e-5
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2"):
        assert np.array_equal(getattr(env, k).cpu().numpy(), getattr(ref, k)), k
    s, rs = env.stats(), ref.stats
    for k in ("episodes", "collisions", "wins_p1", "wins_p2", "timeouts", "merges_ok", "sum_length"):
        assert s[k] == rs[k], k
    assert s["episodes"] > 3000000


def test_quarter_billion_envs_index_arithmetic(mg):
    """Largest size tested: 2^28 + 37 envs (27 GB of state + outputs; obs element offsets pass 2^31, a
    ragged last warp at the far end).  All envs get the same actions, so every env must equal env 0,
    which must equal a 64-env run — any 32-bit overflow in the index arithmetic would break that."""
    n = (1 << 28) + 37
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~35 GB of free device memory")
    env = mg.MergeVecEnv(n, episode_info=False)
    small = mg.MergeVecEnv(64, episode_info=False)
    a1 = torch.full((n,), 4, dtype=torch.uint8, device="cuda"); a2 = torch.full((n,), 1, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        obs, rew, done, info = env.step(a1, a2)
        so, sr, sd, si = small.step(a1[:64], a2[:64])
    torch.cuda.synchronize()
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        v = getattr(env, k)
        assert bool((v == getattr(small, k)[0]).all()), k
    assert bool((obs == so[0]).all()) and bool((rew == sr[0]).all())
    assert not bool(done.any()) and not bool(info["flags"].any())
    assert int(env.steps.min()) == int(env.steps.max()) == 5 and env.stats()["episodes"] == 0
    del env, obs, rew, done, info, a1, a2
    torch.cuda.empty_cache()
