"""TEST INFRASTRUCTURE — CPU oracle for the merging-gym env hot path (float64, NumPy).

This module is the *checker*, never the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.  The
product package (`merging_gym_b200`) has no CPU path and must never route here.

It restates, operation for operation and in the reference's evaluation order,

* `merging_gym/envs/merging_env.py:22-58`   module constants, `lon2coord`
* `merging_gym/envs/merging_env.py:118-132` `MergeEnv.observe`
* `merging_gym/envs/merging_env.py:138-195` `MergeEnv.step`
* `merging_gym/envs/merging_env.py:198-206,232-239` `is_collided` / `corners`
* `merging_gym/envs/merging_env.py:208-230` `MergeEnv.reset`
* `scripts/helper.py:152-191` `mpc_1d` (as the closed form of its QP, see `mpc_1d_acc`)

without gym / pygame / shapely / qpsolvers (all absent from the image).

Parity pin: `tests/golden/*.npz` hold trajectories recorded by running the
UNMODIFIED reference file against the stand-ins in `oracle/ref_shims/`
(`oracle/make_golden.py`); `tests/test_oracle_golden.py` checks this module
against them (discrete outputs bit-exact, continuous to 1e-9).  The reference has
no tests or golden vectors of its own, and the arithmetic that lives in its
un-vendored dependencies — pygame 2.1.2 `Rect` float->int truncation, Shapely
1.8.1 `intersects` on touching rectangles, quadprog 0.1.11's QP solve — is
therefore frozen by definition here: **parity at those three
third-party boundaries is unpinned**.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------- constants
# merging_env.py:22-46
R = 30000
H, W = 1000, 300
dT = 0.2
RFirst = 2.0
RSecond = 1.0
RCollision = -10
vel_penalty = 0.001
time_penalty = 0
START_POINT = 50
END_POINT = H - 50
VEHICLE_W, VEHICLE_H = 4, 8
prediction_t = 3.0
TIME_LIMIT = 500            # merging_env.py:142  `if self.time_stamp > 500`
ACTION_DICT = {0: 0, 1: 10, 2: 20, 3: 30, 4: 40}   # merging_env.py:101
NUM_ACTIONS = 5
OBS_DIM = 10

# info bit-field written by the CUDA kernel and by RefVecEnv (include/merging_b200.h)
INFO_COLLISION = 0x01
INFO_WINNER_SHIFT = 1       # bits 1-2: winner after this step (0 = None, 1, 2)
INFO_WINNER_MASK = 0x06
INFO_TIMEOUT = 0x08
INFO_DONE = 0x10
INFO_BAD_ACTION = 0x80


def lon2coord(lon, ego: bool):
    """merging_env.py:48-58.  `lon` scalar or ndarray (float64)."""
    angle = np.arctan2(H, R) - lon / R
    x = R * np.sin(angle)
    if ego:
        y = W / 2 + (R - R * np.cos(angle))
    else:
        y = W / 2 - (R - R * np.cos(angle))
    return x, y


def mpc_1d_acc(v0, vt):
    """Closed form of `mpc_1d(x0, v0, xt, vt, t=3).action()` (scripts/helper.py:152-191).

    The QP is  min 1/2 u'(D'D + 0.01 I)u  s.t.  sum(dt * u) = vt - v0  with D the 9x10
    first-difference matrix (helper.py:175-179) and dt = t/10 (helper.py:160).  Only
    row 1 of the terminal condition is kept (helper.py:173, 182), so x0 and xt drop
    out.  D.1 = 0  =>  P.1 = 0.01.1  =>  the KKT point is the constant vector
    u = (vt - v0)/(10 dt) = (vt - v0)/t.  `solve_qp_kkt` below is the numeric check.
    """
    return (vt - v0) / prediction_t


def solve_qp_kkt(v0, vt, t=prediction_t):
    """Numeric restatement of helper.py:159-182 (matrix build + equality-constrained QP)."""
    T_len = 10
    dt = t / T_len
    a = np.array([[1, dt], [0, 1]])
    b = np.array([[0], [dt]])
    A = np.zeros([2, T_len])
    tmp = np.eye(2)
    for i in range(T_len)[::-1]:
        A[:, i] = np.matmul(tmp, b).T
        tmp = np.matmul(a, tmp)
    B = (np.array([[0.0], [vt]]) - np.matmul(tmp, np.array([[0.0], [v0]])))[1].reshape(1,)
    p = np.zeros([T_len - 1, T_len])
    for i in range(T_len - 1):
        p[i][i] = 1
        p[i][i + 1] = -1
    P = np.matmul(p.T, p) + np.eye(T_len) * 0.01
    K = np.zeros((T_len + 1, T_len + 1))
    K[:T_len, :T_len] = P
    K[:T_len, T_len] = A[1, :]
    K[T_len, :T_len] = A[1, :]
    rhs = np.zeros(T_len + 1)
    rhs[T_len] = B[0]
    return np.linalg.solve(K, rhs)[:T_len]


def _trunc_int(v):
    """pygame 2.1.2 Rect coordinate conversion: C `(int)double`, truncation toward zero."""
    return np.trunc(v).astype(np.int64)


def is_collided_xy(x1, y1, x2, y2):
    """merging_env.py:198-206 + 232-239.

    `corners(agent, x_i, y_i, 0)` binds y:=x_i, x:=y_i (arg swap at :201/:232), so the
    integer Rect has centre (trunc(y_i), trunc(x_i)), width VEHICLE_W=4 along the
    lateral axis and height VEHICLE_H=8 along the longitudinal axis; rotate(0) is the
    identity.  Two closed rectangles intersect iff both axis projections overlap
    (touching counts, Shapely/GEOS "not disjoint").
    """
    ty1, tx1, ty2, tx2 = _trunc_int(y1), _trunc_int(x1), _trunc_int(y2), _trunc_int(x2)
    return (np.abs(ty1 - ty2) <= VEHICLE_W) & (np.abs(tx1 - tx2) <= VEHICLE_H)


# ----------------------------------------------------------------------------- scalar env
class RefEnv:
    """Scalar restatement of `MergeEnv` (merging_env.py:72-230), same public members."""

    class _Discrete:
        n = NUM_ACTIONS

    class _Box:
        shape = (OBS_DIM,)

    action_space = _Discrete()
    observation_space = _Box()

    def __init__(self):
        self.r1_accumulate = 0
        self.r2_accumulate = 0
        self.time_stamp = 0
        self.n_steps = 0
        self.reset()

    def show_reward(self):                                   # merging_env.py:115-116
        return RFirst, RSecond, RCollision, vel_penalty

    def observe(self):                                       # merging_env.py:118-132
        x1, y1 = lon2coord(self.state1['pos'], True)
        x2, y2 = lon2coord(self.state2['pos'], False)
        return [x2 - x1, y2 - y1, self.state2['vel'] - self.state1['vel'],
                END_POINT - self.state1['pos'], self.state1['vel'],
                x1 - x2, y1 - y2, self.state1['vel'] - self.state2['vel'],
                END_POINT - self.state2['pos'], self.state2['vel']]

    def step(self, action1, action2=None):                   # merging_env.py:138-195
        self.time_stamp += dT
        self.n_steps += 1
        if self.time_stamp > TIME_LIMIT:
            self.done = True
        info = {"collision": False}

        s1, s2 = self.state1, self.state2
        s1['acc'] = mpc_1d_acc(s1['vel'], ACTION_DICT[action1])
        s1['vel'] = max(0, s1['vel'] + s1['acc'] * dT)
        s1['pos'] += s1['vel'] * dT
        s2['acc'] = 0 if action2 is None else mpc_1d_acc(s2['vel'], ACTION_DICT[action2])
        s2['vel'] = max(0, s2['vel'] + s2['acc'] * dT)
        s2['pos'] += s2['vel'] * dT

        obs = self.observe()
        reward1 = - time_penalty - vel_penalty * np.abs(s1['vel'] - 20.0)
        reward2 = - time_penalty - vel_penalty * np.abs(s2['vel'] - 20.0)

        if s1['pos'] > END_POINT:                            # :163  strict
            if self.winner is None:
                self.winner = 1
                reward1 += RFirst
            elif self.winner == 1:
                reward1 = 0
            else:
                reward1 += RSecond
                self.done = True
        if s2['pos'] >= END_POINT:                           # :173  non-strict
            if self.winner is None:
                self.winner = 2
                reward2 += RFirst
            elif self.winner == 2:
                reward2 = 0
            else:
                reward2 += RSecond
                self.done = True

        x1, y1 = lon2coord(s1['pos'], True)
        x2, y2 = lon2coord(s2['pos'], False)
        if bool(is_collided_xy(x1, y1, x2, y2)):             # :183-187
            self.done = True
            reward1 += RCollision
            reward2 += RCollision
            info["collision"] = True

        self.r1_accumulate += reward1
        self.r2_accumulate += reward2
        return obs, [reward1, reward2], self.done, info

    def reset(self):                                         # merging_env.py:208-230
        self.done = False
        self.winner = None
        self.time_stamp = 0
        self.n_steps = 0
        self.state1 = {'pos': START_POINT, 'vel': 20.0, 'acc': 0.0}
        self.state2 = {'pos': START_POINT, 'vel': 20.0, 'acc': 0.0}
        self.r1_accumulate = 0
        self.r2_accumulate = 0
        return self.observe()


# ----------------------------------------------------------------------------- vector env
class RefVecEnv:
    """NumPy float64 restatement over N independent envs, same evaluation order as `RefEnv`.

    `auto_reset=True` follows gym 0.20 `SyncVectorEnv.step_wait`: a finished env is
    reset inside the same call; the returned observation row is the *reset*
    observation while reward / done / info are the terminal step's.  The terminal
    observation, the finished episode's returns and its length are kept in
    `terminal_obs`, `ep_ret`, `ep_len` (rows of envs that have not finished keep
    their previous content).  `auto_reset=False` reproduces the reference's sticky
    `done` (merging_env.py:143,171,181,184,211).

    The time limit is carried as the float64 accumulator the reference uses
    (`time_stamp += 0.2; > 500`, :141-142) *and* as an integer step count, so tests can
    assert that "time_stamp > 500" first holds at step 2501.
    """

    def __init__(self, num_envs: int, pvp: bool = True, auto_reset: bool = True,
                 reset_mode: str = "fixed", reset_seed: int = 0x5EED, env_id_base: int = 0):
        self.n = int(num_envs)
        self.pvp = bool(pvp)
        self.auto_reset = bool(auto_reset)
        self.reset_mode, self.reset_seed, self.env_id_base = reset_mode, int(reset_seed), int(env_id_base)
        n = self.n
        self.resets = np.zeros(n, dtype=np.int64)     # resets each env has had: the random-start counter
        self.pos1 = np.empty(n); self.vel1 = np.empty(n)
        self.pos2 = np.empty(n); self.vel2 = np.empty(n)
        self.ret1 = np.zeros(n); self.ret2 = np.zeros(n)
        self.time_stamp = np.zeros(n)
        self.steps = np.zeros(n, dtype=np.int32)
        self.winner = np.zeros(n, dtype=np.uint8)
        self.done = np.zeros(n, dtype=bool)
        self.terminal_obs = np.zeros((n, OBS_DIM))
        self.ep_ret = np.zeros((n, 2))
        self.ep_len = np.zeros(n, dtype=np.int32)
        self.stats = dict(episodes=0, collisions=0, wins_p1=0, wins_p2=0, timeouts=0,
                          merges_ok=0, sum_length=0, sum_return1=0.0, sum_return2=0.0,
                          bad_actions=0)
        self.reset()

    # -- helpers
    def _reset_rows(self, m):
        if self.reset_mode == "random":
            ids = np.nonzero(m)[0]
            p1, v1, p2, v2 = random_start_draw(self.reset_seed, self.env_id_base + ids, self.resets[ids])
            self.pos1[ids] = p1; self.vel1[ids] = v1; self.pos2[ids] = p2; self.vel2[ids] = v2
        else:
            self.pos1[m] = START_POINT; self.vel1[m] = 20.0
            self.pos2[m] = START_POINT; self.vel2[m] = 20.0
        self.resets[m] = (self.resets[m] + 1) & 0x1FFFF
        self.ret1[m] = 0.0; self.ret2[m] = 0.0
        self.time_stamp[m] = 0.0
        self.steps[m] = 0
        self.winner[m] = 0
        self.done[m] = False

    def observe(self):
        x1, y1 = lon2coord(self.pos1, True)
        x2, y2 = lon2coord(self.pos2, False)
        return np.stack([x2 - x1, y2 - y1, self.vel2 - self.vel1, END_POINT - self.pos1, self.vel1,
                         x1 - x2, y1 - y2, self.vel1 - self.vel2, END_POINT - self.pos2, self.vel2],
                        axis=1)

    def reset(self, mask=None):
        m = np.ones(self.n, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
        self._reset_rows(m)
        return self.observe()

    def step(self, a1, a2=None):
        """Returns (obs[N,10] f64, rewards[N,2] f64, done[N] bool, info[N] u8 bit-field)."""
        n = self.n
        a1 = np.asarray(a1).astype(np.int64)
        bad = (a1 < 0) | (a1 >= NUM_ACTIONS)
        if self.pvp:
            if a2 is None:
                raise ValueError("pvp oracle needs action2")
            a2 = np.asarray(a2).astype(np.int64)
            bad |= (a2 < 0) | (a2 >= NUM_ACTIONS)
            a2 = np.clip(a2, 0, NUM_ACTIONS - 1)
        a1 = np.clip(a1, 0, NUM_ACTIONS - 1)

        self.time_stamp += dT
        self.steps += 1
        timeout = self.time_stamp > TIME_LIMIT
        prev_done = self.done.copy()
        done = self.done | timeout

        vt1 = (10 * a1).astype(np.float64)
        acc1 = mpc_1d_acc(self.vel1, vt1)
        self.vel1 = np.maximum(0.0, self.vel1 + acc1 * dT)
        self.pos1 = self.pos1 + self.vel1 * dT
        if self.pvp:
            vt2 = (10 * a2).astype(np.float64)
            acc2 = mpc_1d_acc(self.vel2, vt2)
        else:
            acc2 = 0.0
        self.vel2 = np.maximum(0.0, self.vel2 + acc2 * dT)
        self.pos2 = self.pos2 + self.vel2 * dT

        x1, y1 = lon2coord(self.pos1, True)
        x2, y2 = lon2coord(self.pos2, False)
        obs = np.stack([x2 - x1, y2 - y1, self.vel2 - self.vel1, END_POINT - self.pos1, self.vel1,
                        x1 - x2, y1 - y2, self.vel1 - self.vel2, END_POINT - self.pos2, self.vel2],
                       axis=1)

        r1 = - time_penalty - vel_penalty * np.abs(self.vel1 - 20.0)
        r2 = - time_penalty - vel_penalty * np.abs(self.vel2 - 20.0)
        w = self.winner.copy()

        f1 = self.pos1 > END_POINT
        c_none = f1 & (w == 0); c_self = f1 & (w == 1); c_other = f1 & (w == 2)
        r1 = np.where(c_none, r1 + RFirst, np.where(c_self, 0.0, np.where(c_other, r1 + RSecond, r1)))
        w = np.where(c_none, 1, w).astype(np.uint8)
        done = done | c_other

        f2 = self.pos2 >= END_POINT
        c_none = f2 & (w == 0); c_self = f2 & (w == 2); c_other = f2 & (w == 1)
        r2 = np.where(c_none, r2 + RFirst, np.where(c_self, 0.0, np.where(c_other, r2 + RSecond, r2)))
        w = np.where(c_none, 2, w).astype(np.uint8)
        done = done | c_other

        col = is_collided_xy(x1, y1, x2, y2)
        done = done | col
        r1 = np.where(col, r1 + RCollision, r1)
        r2 = np.where(col, r2 + RCollision, r2)

        self.winner = w
        self.done = done
        self.ret1 = self.ret1 + r1
        self.ret2 = self.ret2 + r2

        info = (col.astype(np.uint8) * INFO_COLLISION) | (w << INFO_WINNER_SHIFT) \
            | (timeout.astype(np.uint8) * INFO_TIMEOUT) | (done.astype(np.uint8) * INFO_DONE) \
            | (bad.astype(np.uint8) * INFO_BAD_ACTION)
        info = info.astype(np.uint8)
        self.stats["bad_actions"] += int(bad.sum())
        rewards = np.stack([r1, r2], axis=1)
        done_out = done.copy()

        fin = done_out & ~prev_done           # done became true in this step (always so under auto-reset)
        if fin.any():
            m = fin
            self.terminal_obs[m] = obs[m]
            self.ep_ret[m, 0] = self.ret1[m]; self.ep_ret[m, 1] = self.ret2[m]
            self.ep_len[m] = self.steps[m]
            st = self.stats
            st["episodes"] += int(m.sum())
            st["collisions"] += int((m & col).sum())
            st["wins_p1"] += int((m & (w == 1)).sum())
            st["wins_p2"] += int((m & (w == 2)).sum())
            st["timeouts"] += int((m & timeout).sum())
            st["merges_ok"] += int((m & ~col & ~timeout).sum())
            st["sum_length"] += int(self.steps[m].sum())
            st["sum_return1"] += float(self.ret1[m].sum())
            st["sum_return2"] += float(self.ret2[m].sum())
        if self.auto_reset and done_out.any():
            m = done_out                      # a copy: _reset_rows clears self.done in place
            self._reset_rows(m)
            obs = obs.copy()
            obs[m] = self.observe()[m]
        return obs, rewards, done_out, info


# ----------------------------------------------------------------------------- Philox4x32-10
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
ACTION_SEED_DEFAULT = 0x5EED


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123).  `ctr`: uint32[...,4], `key`: uint32[...,2]."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = PHILOX_M0 * c[0]
        p1 = PHILOX_M1 * c[2]
        hi0, lo0 = p0 >> s32, p0 & mask
        hi1, lo1 = p1 >> s32, p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def philox_actions(n, seed, env_id_base, step):
    """Uniform actions for envs [env_id_base, env_id_base+n) at rollout step `step`.

    key = (seed lo32, seed hi32); ctr = (env_id lo32, env_id hi32, step lo32, step hi32);
    action_p = (out[p-1] * 5) >> 32 for p in {1, 2}  (multiply-high, no modulo bias
    beyond 2^-32).  Identical in `csrc/merge_kernels.cu::philox_actions`.
    """
    ids = np.arange(n, dtype=np.uint64) + np.uint64(env_id_base)
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = (ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (ids >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32((step >> 32) & 0xFFFFFFFF)
    key = np.zeros((n, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    out = philox4x32_10(ctr, key).astype(np.uint64)
    a1 = ((out[:, 0] * np.uint64(5)) >> np.uint64(32)).astype(np.uint8)
    a2 = ((out[:, 1] * np.uint64(5)) >> np.uint64(32)).astype(np.uint8)
    return a1, a2


RESET_KEY_XOR = 0x52535445      # "RSTE": separates the reset stream from the action stream


def random_start_draw(seed, env_ids, counts):
    """The reference's commented-out random start (merging_env.py:219-221)

        state1 = {'pos': START_POINT + np.random.randn() * 5, 'vel': 20.0 + np.random.randn() * 3}
        state2 = {'pos': START_POINT + np.random.uniform(-VEHICLE_H/2, VEHICLE_H/2),
                  'vel': 20.0 + np.random.uniform(-5, 10)}

    with counter-based draws: Philox4x32-10, key = (seed lo32, seed hi32 ^ RESET_KEY_XOR), counter =
    (env id lo32, env id hi32, reset count, 0); outputs o0..o3:
    Box-Muller in float64 on u1 = (o0+1)/2^32, u2 = o1/2^32 gives the two normals; o2, o3 the uniforms.
    Same formulas as `merge_device.cuh::random_start` (transcendentals differ by an ulp or two).
    """
    env_ids = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    counts = np.asarray(counts, dtype=np.uint64).reshape(-1)
    n = env_ids.size
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = (env_ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (env_ids >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = (counts & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    key = np.zeros((n, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32(((seed >> 32) & 0xFFFFFFFF) ^ RESET_KEY_XOR)
    o = philox4x32_10(ctr, key).astype(np.float64)
    k32 = 1.0 / 4294967296.0
    u1 = (o[:, 0] + 1.0) * k32
    u2 = o[:, 1] * k32
    r = np.sqrt(-2.0 * np.log(u1))
    z1, z2 = r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)
    p1 = START_POINT + z1 * 5.0
    v1 = 20.0 + z2 * 3.0
    p2 = START_POINT + (-4.0 + 8.0 * ((o[:, 2] + 0.5) * k32))
    v2 = 20.0 + (-5.0 + 15.0 * ((o[:, 3] + 0.5) * k32))
    return p1, v1, p2, v2
