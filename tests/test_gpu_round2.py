"""Round-2 features of the env path, each held to the same bar as the step itself (through the C ABI):
lanes, the pipelined host-buffer step and its field selection, the lean (no-returns) state, optional `done`, banked
statistics reduction, observation refresh after rollout / load_state_dict, and the CUDA path replayed directly against
the BASELINE configs[0] fixture recorded from the unmodified reference."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import merge_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def mg():
    import merging_gym_b200
    return merging_gym_b200


# ---------------------------------------------------------------------------------------------- configs[0] fixture
@pytest.mark.parametrize("name,pvp", [("config1_pve_trace.npz", False), ("pvp_trace.npz", True)])
def test_cuda_path_replays_the_reference_traces(mg, golden, name, pvp):
    """BASELINE configs[0] (pve, one env, 10 000 random steps, manual reset on done) recorded from the unmodified
    reference, replayed through the scalar `make("merging_env-v0")` front end of the CUDA kernel — no oracle in between."""
    tr = golden(name)
    env = mg.make("merging_env-v0")
    assert rel_err(np.asarray(env.reset(), dtype=float), tr["reset_obs"]).max() <= 1e-7
    T = len(tr["done"])
    for t in range(T):
        a1 = int(tr["actions"][t, 0]); a2 = int(tr["actions"][t, 1]) if pvp else None
        o, r, d, info = env.step(a1, a2)
        assert d == bool(tr["done"][t]), t
        assert info["collision"] == bool(tr["collision"][t]), t
        assert (env.winner or 0) == int(tr["winner"][t]), t
        assert rel_err(o, tr["obs"][t]).max() <= TOL, t
        assert rel_err(r, tr["rewards"][t]).max() <= TOL, t
        assert rel_err([env.r1_accumulate, env.r2_accumulate], tr["returns"][t]).max() <= 1e-9, t   # float64 state
        if d:
            env.reset()


# ---------------------------------------------------------------------------------------------- lanes
@pytest.mark.parametrize("n,lanes,mode,reset_mode", [(4096, 2, "pvp", "fixed"), (5000, 3, "pvp", "random"),
                                                     (1031, 2, "pve", "fixed"), (700, 4, "pvp", "fixed")])
def test_laned_env_is_bit_identical(mg, n, lanes, mode, reset_mode):
    kw = dict(mode=mode, seed=21, out_slots=2, reset_mode=reset_mode)
    a = mg.MergeVecEnv(n, **kw)
    b = mg.MergeVecEnv(n, lanes=lanes, **kw)
    assert b.lanes == min(lanes, (n + 255) // 256) and b.lane_bounds[-1][1] == n
    for e in (a, b):
        e.rollout(140)
    for t in range(120):
        a1, a2 = a.sample_actions(1000 + t)
        oa = a.step(a1, a2)
        ob = b.step(a1, a2)
        for x, y in zip(oa[:3], ob[:3]):
            assert torch.equal(x, y), t
        assert torch.equal(oa[3]["flags"], ob[3]["flags"]), t
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert torch.equal(a.stats_tensor(), b.stats_tensor())
    assert torch.equal(a.terminal_obs, b.terminal_obs) and torch.equal(a.episode_length, b.episode_length)


def test_lanes_stepped_independently(mg):
    """The use lanes exist for: each lane advances on its own (lane 0 runs ahead of lane 1), a lane's step depending
    only on that lane's previous one.  Per-lane results equal the single-lane env's rows of the same step."""
    n, T = 2048, 60
    ref = mg.MergeVecEnv(n, seed=5, out_slots=1)
    env = mg.MergeVecEnv(n, seed=5, out_slots=2, lanes=2)
    acts = [tuple(x.clone() for x in ref.sample_actions(t)) for t in range(T)]
    want = []
    for t in range(T):
        o, r, d, i = ref.step(*acts[t])
        want.append((o.clone(), r.clone(), d.clone(), i["flags"].clone()))
    order = [(0, t) for t in range(10)]                      # lane 0 runs 10 steps ahead, then they alternate
    order += [x for t in range(10, T) for x in ((1, t - 10), (0, t))] + [(1, t) for t in range(T - 10, T)]
    for lane, t in order:
        sl = env.lane_slices[lane]
        env.step_lane_async(lane, acts[t][0][sl], acts[t][1][sl])
        o, r, d, i = env.step_lane_wait(lane)
        for got, w in zip((o, r, d, i["flags"]), want[t]):
            assert torch.equal(got, w[sl]), (lane, t)
    env.join_lanes()
    assert torch.equal(env.pos1, ref.pos1) and torch.equal(env.meta, ref.meta)
    assert torch.equal(env.stats_tensor(), ref.stats_tensor())


def test_lanes_inside_a_cuda_graph(mg):
    n = 4096
    ref = mg.MergeVecEnv(n, seed=9)
    env = mg.MergeVecEnv(n, seed=9, lanes=2)
    a1, a2 = (x.clone() for x in ref.sample_actions(3))
    env.step(a1, a2); ref.step(a1, a2)                       # warm-up outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for l, sl in enumerate(env.lane_slices):
            env.step_lane_async(l, a1[sl], a2[sl])
        env.join_lanes()
    ref.load_state_dict(env.state_dict())                    # the capture itself does not execute
    for _ in range(50):
        g.replay()
        ref.step(a1, a2)
    torch.cuda.synchronize()
    assert torch.equal(env.pos1, ref.pos1) and torch.equal(env.ret2, ref.ret2) and torch.equal(env.meta, ref.meta)


# ---------------------------------------------------------------------------------------------- host-buffer pipeline
@pytest.mark.parametrize("upload", [True, False])
def test_step_host_async_pipeline_matches_device_path(mg, upload):
    """upload=True: the actions travel by cudaMemcpyAsync on an upload stream; False: the kernel reads the pinned buffers."""
    n, T = 3000, 40
    dev = mg.MergeVecEnv(n, seed=4)
    env = mg.MergeVecEnv(n, seed=4)
    acts = [tuple(x.cpu().numpy().copy() for x in dev.sample_actions(t)) for t in range(T)]
    want = []
    for t in range(T):
        o, r, d, i = dev.step(torch.from_numpy(acts[t][0]).cuda(), torch.from_numpy(acts[t][1]).cuda())
        want.append([x.cpu().numpy().copy() for x in (o, r, d, i["flags"])])
    got = []
    for t in range(T):
        if t >= 2:
            got.append([x.copy() for x in env.step_host_wait()])
        h1, h2 = env.host_action_buffers()                   # written in place: no staging copy
        h1[:] = acts[t][0]; h2[:] = acts[t][1]
        env.step_host_async(h1, h2, upload=upload)
    with pytest.raises(RuntimeError, match="in flight"):
        env.step_host_async(acts[0][0], acts[0][1])
    with pytest.raises(RuntimeError, match="in flight"):
        env.step_host(acts[0][0], acts[0][1])
    got += [[x.copy() for x in env.step_host_wait()] for _ in range(2)]
    with pytest.raises(RuntimeError, match="without"):
        env.step_host_wait()
    for t in range(T):
        for g, w in zip(got[t], want[t]):
            assert np.array_equal(g, w), t
    assert torch.equal(env.pos1, dev.pos1) and torch.equal(env.stats_tensor(), dev.stats_tensor())


def test_step_host_field_selection(mg):
    n = 1000
    dev = mg.MergeVecEnv(n, seed=8, mode="pve")
    env = mg.MergeVecEnv(n, seed=8, mode="pve")
    dev.rollout(150); env.rollout(150)
    for t, fields in enumerate([("rew", "done", "info"), ("obs",), ("info", "obs"), "done", None, ("rew", "info")]):
        a1 = dev.sample_actions(500 + t)[0]
        o, r, d, i = dev.step(a1, None)
        want = dict(obs=o, rew=r, done=d, info=i["flags"])
        out = env.step_host(a1.cpu().numpy(), None, fields=fields)
        sel = ("obs", "rew", "done", "info") if fields is None else ((fields,) if isinstance(fields, str) else fields)
        for k, v in zip(("obs", "rew", "done", "info"), out):
            if k in sel:
                assert np.array_equal(v, want[k].cpu().numpy()), (fields, k)
            else:
                assert fields is not None and v is None
    with pytest.raises(ValueError):
        env.step_host(a1.cpu().numpy(), None, fields=("obs", "nope"))
    assert torch.equal(env.pos2, dev.pos2)


# ---------------------------------------------------------------------------------------------- lean state, optional done
@pytest.mark.parametrize("pvp", [True, False])
def test_track_returns_false_changes_nothing_else(mg, pvp):
    n = 4096 + 77
    mode = "pvp" if pvp else "pve"
    a = mg.MergeVecEnv(n, mode=mode, seed=13)
    b = mg.MergeVecEnv(n, mode=mode, seed=13, track_returns=False)
    assert b.ret1 is None and "episode_return" not in b.step(*b.sample_actions(0))[3]
    a.step(*a.sample_actions(0))
    with pytest.raises(AttributeError):
        b.r1_accumulate
    a.rollout(200, step0=1); b.rollout(200, step0=1)
    for t in range(60):
        acts = a.sample_actions(300 + t)
        oa, ob = a.step(*acts), b.step(*acts)
        for x, y in zip(oa[:3], ob[:3]):
            assert torch.equal(x, y), t
        assert torch.equal(oa[3]["flags"], ob[3]["flags"])
    for k in ("pos1", "vel1", "pos2", "vel2", "meta"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    sa, sb = a.stats(), b.stats()
    for k in ("episodes", "collisions", "wins_p1", "wins_p2", "timeouts", "merges_ok", "sum_length"):
        assert sa[k] == sb[k] and sa[k] > 0 or k == "timeouts", k
    assert sb["sum_return1"] == 0 and sa["sum_return1"] != 0
    assert torch.equal(a.episode_length, b.episode_length)


def test_done_output_is_optional(mg):
    """MgOut.done = NULL: the same bit is MG_INFO_DONE of the info byte."""
    from merging_gym_b200 import _native as nat
    n = 2048 + 5
    a = mg.MergeVecEnv(n, seed=2)
    b = mg.MergeVecEnv(n, seed=2)
    a.rollout(200); b.rollout(200)
    acts = a.sample_actions(777)
    o, r, d, i = a.step(*acts)
    canary = torch.full_like(b.done_buf[0], 0x5A)
    b.done_buf[0].copy_(canary)
    out = nat.MgOut(b.obs_buf[0].data_ptr(), b.rew_buf[0].data_ptr(), None, b.info_buf[0].data_ptr(), None, None, None)
    nat.check(b._lib.mg_step(C.byref(b._state), n, C.c_void_p(acts[0].data_ptr()), C.c_void_p(acts[1].data_ptr()), nat.ACT_U8,
                             C.byref(b._rw), C.byref(out), None, b._flags(), C.byref(b._rs), b._stream()), "mg_step")
    assert torch.equal(b.done_buf[0], canary)                               # untouched
    assert torch.equal(b.info_buf[0], i["flags"]) and torch.equal(b.obs_buf[0], o) and torch.equal(b.rew_buf[0], r)
    assert torch.equal((b.info_buf[0] & nat.INFO_DONE).bool(), d)


def test_mixed_action_dtypes_widen(mg):
    """a1 uint8 + a2 int64: both go to the wider type, so an out-of-range 256 stays out of range (not wrapped to 0)."""
    n = 300
    env = mg.MergeVecEnv(n)
    a1 = torch.full((n,), 2, dtype=torch.uint8, device="cuda")
    a2 = torch.full((n,), 2, dtype=torch.int64, device="cuda")
    a2[7] = 256
    _, _, _, info = env.step(a1, a2)
    bad = info["bad_action"].cpu().numpy()
    assert bad[7] and bad.sum() == 1


# ---------------------------------------------------------------------------------------------- statistics banks
def test_banked_reducer_equals_plain_totals(mg):
    n = 8192
    a = mg.MergeVecEnv(n, seed=31)
    b = mg.MergeVecEnv(n, seed=31)
    red = mg.AsyncStatsReducer(b, banked=True)
    a.rollout(150); b.rollout(150)
    for t in range(96):
        acts = a.sample_actions(400 + t)
        a.step(*acts); b.step(*acts)
        if t % 16 == 15:
            red.submit()
            if t % 32 == 31:
                assert torch.equal(red.latest(), a.stats_tensor()), t
    for t in range(5):                                       # launches after the last submit sit in the active bank
        acts = a.sample_actions(900 + t)
        a.step(*acts); b.step(*acts)
    assert torch.equal(a.stats_tensor(), b.stats_tensor())
    b2 = mg.MergeVecEnv(n, seed=31)
    b2.load_state_dict(b.state_dict())
    assert torch.equal(b2.stats_tensor(), a.stats_tensor())
    assert a.stats(reset=True)["episodes"] == b.stats(reset=True)["episodes"] > 0
    assert int(b.stats_tensor().abs().sum()) == 0


def test_banked_reducer_with_one_graph_per_bank(mg):
    n, G = 4096, 8
    a = mg.MergeVecEnv(n, seed=6)
    b = mg.MergeVecEnv(n, seed=6)
    red = mg.AsyncStatsReducer(b, banked=True)
    a.rollout(180); b.rollout(180)
    acts = [tuple(x.clone() for x in a.sample_actions(50 + t)) for t in range(G)]
    b.step(*acts[0]); a.step(*acts[0])
    torch.cuda.synchronize()

    def issue():
        for t in range(G):
            b.step_async(*acts[t])
    graphs = red.capture_per_bank(issue)
    for rep in range(6):
        red.replay(graphs)
        red.submit()
        for t in range(G):
            a.step_async(*acts[t])
        assert torch.equal(red.latest(), a.stats_tensor()), rep
    assert torch.equal(a.pos1, b.pos1)


# ---------------------------------------------------------------------------------------------- observation refresh
def test_rollout_and_load_state_dict_refresh_the_observation(mg):
    n = 1500
    env = mg.MergeVecEnv(n, seed=17)
    ref = mo.RefVecEnv(n)
    env.rollout(215)
    for t in range(215):
        robs = ref.step(*mo.philox_actions(n, 17, 0, t))[0]
    assert rel_err(env.obs_buf[env._slot].cpu().numpy(), robs).max() <= TOL       # the policy's next input
    stale = env.obs_buf[env._slot].clone()
    env.rollout(3, refresh_obs=False)
    assert torch.equal(env.obs_buf[env._slot], stale)
    sd = env.state_dict()
    assert sd["philox_seed"] == 17 and sd["env_id_base"] == 0 and sd["reset_mode"] == "fixed" and "reset_seed" in sd
    other = mg.MergeVecEnv(n, seed=99)
    other.load_state_dict(sd)
    for t in range(3):
        robs = ref.step(*mo.philox_actions(n, 17, 0, 215 + t))[0]
    assert rel_err(other.obs_buf[other._slot].cpu().numpy(), robs).max() <= TOL
    assert other.philox_seed == 17 and other.step_count == env.step_count
    a, b = env.step(*env.sample_actions()), other.step(*other.sample_actions())
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_recorder_rejects_aliased_observations(mg):
    env = mg.MergeVecEnv(512, out_slots=1)
    rec = mg.TransitionRecorder(env, 4096)
    obs = env.reset()
    a1, a2 = env.sample_actions()
    out = env.step(a1, a2)
    with pytest.raises(ValueError, match="aliases"):
        rec.record(obs, a1, a2, out)                        # out_slots=1: `obs` IS out[0] now
    with pytest.raises(ValueError, match="contiguous float32"):
        rec.record(obs.double(), a1, a2, out)
    rec.record(obs.clone(), a1, a2, out)
    assert len(rec) > 0


def test_batched_spaces_are_views(mg):
    env = mg.MergeVecEnv(1 << 20, episode_info=False, track_stats=False)
    o, a = env.observation_space, env.action_space
    assert o.shape == (1 << 20, 10) and a.shape == (1 << 20,)
    assert o.low.strides[0] == 0 and a.nvec.strides[0] == 0            # broadcast views: O(1) host memory
    assert o.low[12345].tolist() == env.single_observation_space.low.tolist()
