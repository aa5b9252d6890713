"""The oracle against golden vectors recorded from the UNMODIFIED reference file
(oracle/make_golden.py; /root/reference/merging_gym/envs/merging_env.py + scripts/helper.py run
against oracle/ref_shims).  Discrete outputs bit-exact, continuous to 1e-9 relative (the shimmed
QP solve differs from the closed form by O(1e-15))."""
import numpy as np
import pytest

from oracle import merge_oracle as mo
from conftest import rel_err

CONT_TOL = 1e-9


def _replay_scalar(tr, pvp):
    env = mo.RefEnv()
    assert np.array_equal(np.asarray(env.reset(), dtype=float), tr["reset_obs"])
    T = len(tr["done"])
    for t in range(T):
        a1 = int(tr["actions"][t, 0]); a2 = int(tr["actions"][t, 1]) if pvp else None
        o, r, d, info = env.step(a1, a2)
        assert d == bool(tr["done"][t]), t
        assert info["collision"] == bool(tr["collision"][t]), t
        assert (env.winner or 0) == int(tr["winner"][t]), t
        assert rel_err(o, tr["obs"][t]).max() < CONT_TOL, t
        assert rel_err(r, tr["rewards"][t]).max() < CONT_TOL, t
        assert rel_err([env.r1_accumulate, env.r2_accumulate], tr["returns"][t]).max() < CONT_TOL, t
        if d:
            env.reset()


def test_config1_pve_trace(golden):
    """BASELINE.json configs[0]: pve, single env, random actions, 10k steps."""
    tr = golden("config1_pve_trace.npz")
    assert len(tr["done"]) == 10000 and not bool(tr["pvp"])
    _replay_scalar(tr, pvp=False)


def test_pvp_trace(golden):
    _replay_scalar(golden("pvp_trace.npz"), pvp=True)


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_vec16_trace(golden, impl):
    """Vector oracle (gym-0.20 auto-reset convention) vs 16 reference envs x 640 steps."""
    tr = golden("pvp_vec16_trace.npz")
    T, N = tr["done"].shape
    if impl == "numpy":
        env = mo.RefVecEnv(N, pvp=True, auto_reset=True)
    else:
        from oracle import c_oracle
        env = c_oracle.CVecEnv(N, pvp=True, auto_reset=True)
    for t in range(T):
        obs, rew, done, info = env.step(tr["actions"][t, :, 0], tr["actions"][t, :, 1])
        assert np.array_equal(done, tr["done"][t]), t
        assert np.array_equal((info & mo.INFO_COLLISION) != 0, tr["collision"][t]), t
        assert np.array_equal((info & mo.INFO_WINNER_MASK) >> 1, tr["winner"][t]), t
        assert rel_err(obs, tr["obs"][t]).max() < CONT_TOL, t
        assert rel_err(rew, tr["rewards"][t]).max() < CONT_TOL, t
        m = tr["done"][t]
        if m.any():
            assert rel_err(env.terminal_obs[m], tr["step_obs"][t][m]).max() < CONT_TOL
            assert np.array_equal(env.ep_len[m], tr["ep_len"][t][m])
            assert rel_err(env.ep_ret[m], tr["ep_ret"][t][m]).max() < CONT_TOL
    st = env.stats
    assert st["episodes"] == int(tr["done"].sum())
    assert st["collisions"] == int((tr["done"] & tr["collision"]).sum())
    assert st["sum_length"] == int(tr["ep_len"].sum())


def _run_script(env, s1, s2, max_steps=2700):
    def act(s, t):
        return s if (s is None or isinstance(s, int)) else s[t % len(s)]
    env.reset()
    for t in range(max_steps):
        o, r, d, info = env.step(act(s1, t), act(s2, t))
        if d:
            break
    return t + 1, o, r, info


def test_known_answer_episodes(golden):
    """SURVEY.md §8c KATs: truncation vs rounding (2,2), '>=' vs '>' (3,None), timeout (0,*) ..."""
    kats = golden("kat.json")
    assert np.array_equal(np.asarray(mo.RefEnv().reset(), dtype=float), kats[-1]["reset_obs"])
    for k in kats[:-1]:
        env = mo.RefEnv()
        steps, o, r, info = _run_script(env, k["a1"], k["a2"])
        assert steps == k["steps"], k
        assert env.winner == k["winner"], k
        assert info["collision"] == k["collision"], k
        assert rel_err([env.r1_accumulate, env.r2_accumulate], [k["R1"], k["R2"]]).max() < CONT_TOL, k
        assert rel_err([env.state1["pos"], env.state2["pos"]], [k["pos1"], k["pos2"]]).max() < CONT_TOL, k
        assert rel_err(o, k["last_obs"]).max() < CONT_TOL, k


def test_kat_values_from_survey(golden):
    """A few literal values, so the fixture file itself is pinned."""
    k = {(str(x["a1"]), str(x["a2"])): x for x in golden("kat.json")[:-1]}
    a = k[("2", "2")]
    assert (a["steps"], a["winner"], a["collision"], a["pos1"]) == (151, None, True, 654.0)
    assert a["last_obs"][1] == -3.981956335635914
    b = k[("3", "None")]
    assert (b["steps"], b["winner"], b["collision"], b["pos2"], b["R2"]) == (225, 1, False, 950.0, 1.0)
    c = k[("0", "None")]
    assert (c["steps"], c["winner"], c["time_stamp"]) == (2501, 2, 500.19999999998015)
    d = k[("4", "4")]
    assert (d["steps"], d["collision"]) == (83, True)


def test_time_limit_is_step_2501():
    """merging_env.py:141-142: float64 `time_stamp += 0.2; > 500` first holds at step 2501."""
    ts, n = 0.0, 0
    while not ts > mo.TIME_LIMIT:
        ts += mo.dT
        n += 1
    assert n == 2501 and ts == 500.19999999998015


def test_qp_closed_form():
    """scripts/helper.py:152-191: the QP's first control equals (vt - v0)/3."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        v0 = rng.uniform(0, 40); vt = float(rng.integers(5) * 10)
        u = mo.solve_qp_kkt(v0, vt)
        assert np.allclose(u, u[0], rtol=0, atol=1e-12)
        assert abs(u[0] - mo.mpc_1d_acc(v0, vt)) <= 1e-13 * max(1.0, abs(u[0]))
    assert np.all(mo.solve_qp_kkt(20.0, 20.0) == 0.0)


def test_collision_rule_truncation_and_touching():
    """trunc toward zero + closed rectangles (touching counts)."""
    f = mo.is_collided_xy
    assert f(100.9, 150.9, 108.0, 154.99)          # |100-108| = 8, |150-154| = 4 -> touching
    assert not f(100.9, 150.9, 109.0, 154.99)
    assert not f(100.9, 150.9, 108.0, 155.0)
    assert f(-0.9, 150.0, 0.9, 150.0) and f(-8.9, 150, 0.5, 150)   # trunc(-8.9) = -8
    assert not f(-9.0, 150, 0.5, 150)


def test_vec_oracle_equals_scalar_oracle():
    N, T = 32, 400
    rng = np.random.default_rng(11)
    for pvp in (True, False):
        vec = mo.RefVecEnv(N, pvp=pvp)
        sc = [mo.RefEnv() for _ in range(N)]
        for t in range(T):
            a = rng.integers(0, 5, (N, 2))
            obs, rew, done, info = vec.step(a[:, 0], a[:, 1] if pvp else None)
            for e in range(N):
                o, r, d, i = sc[e].step(int(a[e, 0]), int(a[e, 1]) if pvp else None)
                assert d == done[e] and i["collision"] == bool(info[e] & 1)
                assert np.array_equal(np.asarray(r, dtype=float), rew[e])
                if d:
                    o = sc[e].reset()
                assert np.array_equal(np.asarray(o, dtype=float), obs[e])


def test_sticky_done_without_auto_reset():
    env = mo.RefVecEnv(4, pvp=True, auto_reset=False)
    a = np.full(4, 2)
    seen = np.zeros(4, bool)
    for t in range(200):
        _, _, done, _ = env.step(a, a)
        assert np.all(done | ~seen)      # once done, always done
        seen |= done
    assert seen.all() and (env.steps == 200).all()


def test_random_start_distribution():
    """merging_env.py:219-221 (commented-out random start): N(50, 5), N(20, 3), U(46, 54), U(15, 30)."""
    n = 200000
    p1, v1, p2, v2 = mo.random_start_draw(123, np.arange(n), np.zeros(n))
    assert abs(p1.mean() - 50) < 0.05 and abs(p1.std() - 5) < 0.05
    assert abs(v1.mean() - 20) < 0.03 and abs(v1.std() - 3) < 0.03
    assert 46 <= p2.min() < 46.01 and 53.99 < p2.max() <= 54 and abs(p2.mean() - 50) < 0.02
    assert 15 <= v2.min() < 15.01 and 29.99 < v2.max() <= 30 and abs(v2.mean() - 22.5) < 0.04
    assert abs(np.corrcoef(p1, v1)[0, 1]) < 0.01
    q1, *_ = mo.random_start_draw(123, np.arange(n), np.ones(n))       # next episode: fresh draws
    assert abs(np.corrcoef(p1, q1)[0, 1]) < 0.01
    env = mo.RefVecEnv(64, reset_mode="random", reset_seed=123)
    assert np.array_equal(env.pos1, p1[:64]) and env.resets.tolist() == [1] * 64


def test_collision_rule_equals_reference_polygons():
    """`is_collided_xy` against the reference's own `corners()` + Polygon.intersects code path
    (merging_env.py:198-206, 232-239) driven with random float centres, through the stand-in
    pygame Rect / shapely Polygon (needs /root/reference, i.e. the build container)."""
    from oracle.ref_loader import load_reference_env, quiet, reference_available
    if not reference_available():
        pytest.skip("reference tree not present")
    env = load_reference_env()
    from shapely.geometry import Polygon          # the shim put on sys.path by the loader
    rng = np.random.default_rng(0)
    n = 4000
    x1 = rng.uniform(-30, 1000, n); y1 = rng.uniform(120, 180, n)
    # second car close to the first so that both outcomes are frequent, incl. touching cases
    x2 = x1 + rng.choice([-9.5, -8.6, -8.0, -7.4, 0.0, 3.3, 7.9, 8.0, 8.2, 9.9], n) + rng.uniform(-0.6, 0.6, n)
    y2 = y1 + rng.choice([-5.5, -4.4, -4.0, -3.6, 0.0, 2.2, 3.9, 4.0, 4.1, 5.2], n) + rng.uniform(-0.6, 0.6, n)
    want = mo.is_collided_xy(x1, y1, x2, y2)
    got = np.zeros(n, bool)
    with quiet():
        for i in range(n):
            p1 = Polygon([(p.x, p.y) for p in env.corners(env.ego, x1[i], y1[i], 0)])
            p2 = Polygon([(p.x, p.y) for p in env.corners(env.opponent, x2[i], y2[i], 0)])
            got[i] = p1.intersects(p2)
    assert np.array_equal(got, want)
    assert 0.2 < want.mean() < 0.8


def _injected_batches(tr):
    for pvp in (True, False):
        idx = np.nonzero(tr["pvp"] == pvp)[0]
        yield pvp, idx


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_injected_states(golden, impl):
    """One step of the unmodified reference from 6 000 injected states (close pairs around the merge point,
    photo finishes, stopped cars, a winner already past END_POINT): collision / done / winner bit-exact, the
    float64 state after the step bit-identical, observations and rewards to 1e-9."""
    tr = golden("injected_states.npz")
    assert tr["collision"].sum() > 1000 and (tr["winner"] == 2).sum() > 300
    for pvp, idx in _injected_batches(tr):
        n = len(idx)
        if impl == "numpy":
            env = mo.RefVecEnv(n, pvp=pvp, auto_reset=False)
        else:
            from oracle import c_oracle
            env = c_oracle.CVecEnv(n, pvp=pvp, auto_reset=False)
        for j, k in enumerate(("pos1", "vel1", "pos2", "vel2")):
            getattr(env, k)[:] = tr["pos"][idx, j]
        env.winner[:] = tr["winner_before"][idx]
        obs, rew, done, info = env.step(tr["actions"][idx, 0], tr["actions"][idx, 1])
        assert np.array_equal(done, tr["done"][idx])
        assert np.array_equal((info & 1).astype(bool), tr["collision"][idx])
        assert np.array_equal(np.asarray(env.winner), tr["winner"][idx])
        after = np.stack([env.pos1, env.vel1, env.pos2, env.vel2], 1)
        assert rel_err(after, tr["state_after"][idx]).max() < CONT_TOL
        assert rel_err(obs, tr["obs"][idx]).max() < CONT_TOL
        assert rel_err(rew, tr["rewards"][idx]).max() < CONT_TOL
