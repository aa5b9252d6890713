"""`MergeVecEnv` — gym-0.20-style vectorised front end of the fused CUDA env step.

Host-side mirror of the reference's `MergeEnv` interface (merging_gym/envs/merging_env.py:72-230)
for N independent env instances that live entirely in GPU memory:

    reference (one env, Python floats)                 this class (N envs, device tensors)
    ------------------------------------------------   ---------------------------------------------
    env.reset() -> list[10]                  (:208)    reset() -> obs f32[N,10]
    env.step(a1, a2|None) -> obs, [r1,r2],   (:138)    step(a1[N], a2[N]|None) -> obs f32[N,10],
                             done, {"collision"}                 rewards f32[N,2], done bool[N], info
    env.winner / r1_accumulate / r2_accumulate         .winner u8[N] / .r1_accumulate f64[N] / ...
    env.action_space.n == 5, observation_space (10,)   single_action_space / single_observation_space
    env.show_reward()                        (:115)    show_reward()

All tensors are owned by PyTorch (caching allocator, streams, CUDA graphs just work); the
kernels come from libmerging_b200.so through the C ABI (include/merging_b200.h) and are launched
on the current torch stream.  Nothing here synchronises with the host unless stated.  There is
no CPU path: constructing the env without CUDA raises.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Mapping
from typing import Optional

import numpy as np
import torch

from . import _native as nat
from .spaces import Box, Discrete, MultiDiscrete, merge_action_space, merge_observation_space
from .tracing import nvtx_range

ACTION_SEED_DEFAULT = 0x5EED


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _slot_views(block: torch.Tensor, K: int, n: int, n_pad: int):
    """obs f32[K,n,10], rew f32[K,n,2], done u8[K,n], info u8[K,n] views of a uint8 block of K slots of
    n_pad * 50 bytes laid out [obs | rew | done | info] (n_pad a multiple of 16: every piece 16-byte aligned)."""
    f = block.view(torch.float32)
    sf = n_pad * 50 // 4                                   # slot stride in floats
    obs = torch.as_strided(f, (K, n, 10), (sf, 10, 1), 0)
    rew = torch.as_strided(f, (K, n, 2), (sf, 2, 1), n_pad * 10)
    done = torch.as_strided(block, (K, n), (n_pad * 50, 1), n_pad * 48)
    info = torch.as_strided(block, (K, n), (n_pad * 50, 1), n_pad * 49)
    return obs, rew, done, info


class StepInfo(Mapping):
    """Lazy `info` of one vector step: device tensors decoded from the info byte on access.

    Keys: "collision" bool[N] (merging_env.py:144,187), "winner" u8[N] (0 = None; :164-181),
    "timeout" bool[N] (:142), "bad_action" bool[N], "flags" u8[N] (raw MG_INFO_* bits), and — when
    the env keeps them — "terminal_observation" f32[N,10], "episode_return" f32[N,2],
    "episode_length" i32[N]; these three hold, per env, the values of the most recently finished
    episode and are meaningful where `done` is (or has been) set.
    """

    _KEYS = ("collision", "winner", "timeout", "bad_action", "flags")

    def __init__(self, flags: torch.Tensor, extras: dict):
        self._flags = flags
        self._extras = extras

    def __getitem__(self, k):
        f = self._flags
        if k == "flags":
            return f
        if k == "collision":
            return (f & nat.INFO_COLLISION).bool()
        if k == "winner":
            return (f & nat.INFO_WINNER_MASK) >> nat.INFO_WINNER_SHIFT
        if k == "timeout":
            return (f & nat.INFO_TIMEOUT).bool()
        if k == "bad_action":
            return (f & nat.INFO_BAD_ACTION).bool()
        return self._extras[k]

    def __iter__(self):
        yield from self._KEYS
        yield from self._extras

    def __len__(self):
        return len(self._KEYS) + len(self._extras)


class MergeVecEnv:
    """N device-resident merging envs stepped by one fused kernel launch.

    Parameters
    ----------
    num_envs      number of env instances on this device (this rank's shard).
    mode          "pvp" (two action vectors) or "pve" (`action2=None`: player 2 keeps its speed,
                  merging_env.py:152).  Only selects the default for `sample_actions`/`rollout`;
                  `step(a1, None)` is always pve and `step(a1, a2)` always pvp, like the reference.
    auto_reset    True: gym-0.20 `SyncVectorEnv` convention (finished env is reset inside `step`
                  and returns its reset observation).  False: the reference's sticky `done`.
    seed, env_id_base   Philox key and first *global* env id of this shard; trajectories of
                  synthetic rollouts depend only on (seed, global env id, step), never on sharding.
    out_slots     number of output buffer sets rotated over steps (a rollout ring, [T,N,...]).
    episode_info  keep terminal_observation / episode_return / episode_length buffers.
    track_stats   accumulate episode statistics on the device (see `stats()`).
    rewards       dict overriding RFirst/RSecond/RCollision/vel_penalty/time_penalty (:28-32).
    reset_mode    "fixed": pos=50, vel=20 (merging_env.py:216-217, the live code); "random": the
                  reference's commented-out random start (:219-221) drawn per (reset_seed, global
                  env id, reset count) with Philox + Box-Muller, for explicit and automatic resets.
    track_returns False drops the float64 `r1_accumulate` / `r2_accumulate` state (merging_env.py:191-192;
                  only human_player.py:189-193 and `render` read them): 124 instead of 156 bytes of HBM
                  traffic per env-step.  `episode_return`, the return statistics and the two attributes
                  are then unavailable.
    host_slots    pinned host buffer sets of the host-buffer path (`step_host_async` keeps up to this
                  many steps in flight).
    obs_layout    "aos" (default): `obs f32[N,10]`, the gym-shaped rows.  "soa": `obs f32[10,S]`, S = N rounded up to 16 —
                  one column per feature, what a fused policy consumer reads best.  "goal_slot": `obs f32[N,11]`, rows
                  `[goal] + state` (hdqn.py:291) whose slot 0 belongs to the goal policy (`HDQNPolicy` writes its goal
                  there and its controller reads the row as it is; the env never touches slot 0).  `observation()` gives
                  an [N,10] view of any of them.  Supported by reset / step / policy_step and the policy kernels;
                  `rollout(obs=...)`, `step_host*` and the recorders need the default rows.
    lanes         L > 1 splits the envs into L contiguous sub-shards ("lanes", whole 256-env blocks), each
                  stepped by its own launch on its own CUDA stream: `step_lane_async(l, ...)` /
                  `step_lane_wait(l)`.  A lane depends only on its own previous step, so the launch of one
                  lane starts while another lane's launch drains — that overlap is what separates the
                  serialised 2^20-env launch (0.91 of the HBM copy peak) from the peak.  Results are
                  bit-identical to the single-lane env; `step()` steps all lanes and joins them.
    """

    def __init__(self, num_envs: int, mode: str = "pvp", device="cuda", auto_reset: bool = True,
                 seed: int = ACTION_SEED_DEFAULT, env_id_base: int = 0, out_slots: int = 1,
                 episode_info: bool = True, track_stats: bool = True, rewards: Optional[dict] = None,
                 validate_actions: bool = False, reset_mode: str = "fixed", reset_seed: Optional[int] = None,
                 track_returns: bool = True, host_slots: int = 2, lanes: int = 1, obs_layout: str = "aos"):
        if mode not in ("pvp", "pve"):
            raise ValueError("mode must be 'pvp' or 'pve'")
        if num_envs < 0 or out_slots < 1:
            raise ValueError("num_envs must be >= 0 and out_slots >= 1")
        if reset_mode not in ("fixed", "random"):
            raise ValueError("reset_mode must be 'fixed' or 'random'")
        if host_slots < 1:
            raise ValueError("host_slots must be >= 1")
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        if obs_layout not in nat.OBS_LAYOUTS:
            raise ValueError(f"obs_layout must be one of {nat.OBS_LAYOUTS}")
        if obs_layout != "aos" and (lanes != 1 or not track_returns):
            raise ValueError("obs_layout 'soa' / 'goal_slot' needs lanes=1 and track_returns=True")
        self.obs_layout = obs_layout
        self._lib = nat.load()                       # raises if the CUDA library is not built
        if not torch.cuda.is_available():
            raise nat.NativeError("merging_gym_b200 needs a CUDA device (no CPU fallback exists)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise nat.NativeError(f"device must be a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = n = int(num_envs)
        self.mode, self.auto_reset = mode, bool(auto_reset)
        self.philox_seed, self.env_id_base = int(seed), int(env_id_base)
        self.validate_actions = bool(validate_actions)
        self.reset_mode = reset_mode
        self._rs = nat.MgResetSpec(nat.RESET_RANDOM if reset_mode == "random" else nat.RESET_FIXED, 0,
                                   int(seed if reset_seed is None else reset_seed), int(env_id_base))
        self.out_slots = int(out_slots)
        self.track_returns = bool(track_returns)
        self.host_slots = int(host_slots)
        self._slot = 0
        self.step_count = 0                          # Philox step counter for sample_actions/rollout
        self.closed = False

        self._rw = nat.default_rewards()
        for k, v in (rewards or {}).items():
            if not hasattr(self._rw, k):
                raise KeyError(f"unknown reward parameter {k!r}")
            setattr(self._rw, k, float(v))
        self.constants = nat.constants()

        dev = self.device
        # one contiguous state allocation: 6 float64 arrays + the uint32 meta array, each padded to a
        # multiple of 32 envs (keeps every array 256-byte aligned)
        n32 = (n + 31) // 32 * 32
        self._state_block = torch.zeros(n32 * 52, dtype=torch.uint8, device=dev)
        f = self._state_block[:n32 * 48].view(torch.float64).view(6, n32)
        self.pos1, self.vel1, self.pos2, self.vel2, self.ret1, self.ret2 = (f[i, :n] for i in range(6))
        self.meta = self._state_block[n32 * 48:].view(torch.int32)[:n]
        if not self.track_returns:
            self.ret1 = self.ret2 = None
        self._state = nat.MgState(*[None if t is None else t.data_ptr() for t in
                                    (self.pos1, self.vel1, self.pos2, self.vel2, self.ret1, self.ret2, self.meta)])
        K = self.out_slots
        n_pad = (n + 15) // 16 * 16                  # keeps every slot's base pointer 16-byte aligned
        # one allocation, slot-major: [obs | rew | done | info] of a slot are back to back (50 bytes per env), so the
        # host-buffer path can fetch a whole slot with ONE device-to-host copy
        self._out_block = torch.zeros(K * n_pad * 50, dtype=torch.uint8, device=dev)
        self.obs_buf, self.rew_buf, self.done_buf, self.info_buf = _slot_views(self._out_block, K, n, n_pad)
        if obs_layout == "soa":                      # [10, S] per slot; S * 4 bytes keeps every column 64-byte aligned
            self.obs_buf = torch.zeros(K, nat.OBS_DIM, nat.soa_stride(n), dtype=torch.float32, device=dev)
        elif obs_layout == "goal_slot":              # [N, 11] rows `[goal] + state`
            self.obs_buf = torch.zeros(K, n_pad, nat.OBS_DIM + 1, dtype=torch.float32, device=dev)[:, :n]   # slots 16-byte aligned
        self._extras = {}
        if episode_info:
            self.terminal_obs = torch.zeros(n, nat.OBS_DIM, dtype=torch.float32, device=dev)
            self.episode_return = torch.zeros(n, 2, dtype=torch.float32, device=dev) if self.track_returns else None
            self.episode_length = torch.zeros(n, dtype=torch.int32, device=dev)
            self._extras = {"terminal_observation": self.terminal_obs, "episode_length": self.episode_length}
            if self.track_returns:
                self._extras["episode_return"] = self.episode_return
        else:
            self.terminal_obs = self.episode_return = self.episode_length = None
        self._outs = [nat.MgOut(self.obs_buf[k].data_ptr(), self.rew_buf[k].data_ptr(),
                                self.done_buf[k].data_ptr(), self.info_buf[k].data_ptr(),
                                *[None if t is None else t.data_ptr() for t in
                                  (self.terminal_obs, self.episode_return, self.episode_length)])
                      for k in range(K)]
        # Episode statistics: the kernels add into the ACTIVE bank of per-block partial rows.  A banked
        # `AsyncStatsReducer` retires the active bank and drains it (row sum into `_stats_total`, zero) on its side
        # stream, so the launching stream never runs a statistics kernel; without one, bank 0 is all there is.
        self._stats_banks = ([torch.zeros(nat.STATS_ROWS, nat.STATS_COLS, dtype=torch.int64, device=dev)]
                             if track_stats else None)
        self._stats_active = 0
        self._stats_total = torch.zeros(nat.STATS_COLS, dtype=torch.int64, device=dev) if track_stats else None
        self._stats_side_event = None                # last work a reducer queued on its side stream
        self._act_block = torch.zeros(2 * n_pad, dtype=torch.uint8, device=dev)   # sample_actions / step_host scratch
        self.act1, self.act2 = self._act_block[:n], self._act_block[n_pad:n_pad + n]
        self._host = None
        self._hslots = None
        self._hfly = []                              # host slots with a step in flight, oldest first
        self._hnext = 0
        self._copy_stream = None
        self._pending = None
        self._zero_mask = None
        self._init_lanes(int(lanes))

        self.single_observation_space: Box = merge_observation_space()
        self.single_action_space: Discrete = merge_action_space()
        self._batched_spaces = None
        self.reset()

    @property
    def observation_space(self) -> Box:
        """Batched Box[N,10] (gym 0.20 `VectorEnv.observation_space`); the bounds are broadcast views, O(1) memory."""
        return self._spaces()[0]

    @property
    def action_space(self) -> MultiDiscrete:
        return self._spaces()[1]

    def _spaces(self):
        if self._batched_spaces is None:
            n, so = self.num_envs, self.single_observation_space
            self._batched_spaces = (Box(np.broadcast_to(so.low, (n,) + so.low.shape), np.broadcast_to(so.high, (n,) + so.high.shape),
                                        dtype=np.float16),
                                    MultiDiscrete(np.broadcast_to(np.int64(nat.NUM_ACTIONS), (n,))))
        return self._batched_spaces

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _flags(self):
        return (nat.FLAG_AUTO_RESET if self.auto_reset else 0) | (0 if self.track_returns else nat.FLAG_NO_RETURNS)

    def _step_flags(self):
        """`_flags()` + the observation layout (mg_step / mg_policy_step write their observations in it)."""
        return self._flags() | nat.OBS_LAYOUT_FLAG[self.obs_layout]

    def _need_rows(self, what: str) -> None:
        if self.obs_layout != "aos":
            raise ValueError(f"{what} works on the default [N,10] observation rows (obs_layout='aos'), "
                             f"this env has obs_layout={self.obs_layout!r}")

    def observation(self, slot: Optional[int] = None) -> torch.Tensor:
        """The observations of output slot `slot` (default: the current one) as an [N,10] tensor whatever the layout —
        the buffer itself for "aos", a strided view for "soa" / "goal_slot" (`.contiguous()` copies)."""
        o = self.obs_buf[self._slot if slot is None else slot]
        if self.obs_layout == "soa":
            return o[:, :self.num_envs].t()
        if self.obs_layout == "goal_slot":
            return o[:, 1:]
        return o

    def lane_stream(self, lane: int):
        """The CUDA stream lane `lane`'s launches run on (the current stream for a single-lane env)."""
        return self._lane_streams[lane] if self.lanes > 1 else torch.cuda.current_stream(self.device)

    def _as_action(self, a, name):
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a))
        if a.dtype not in (torch.uint8, torch.int32, torch.int64):
            if a.dtype in (torch.int8, torch.int16, torch.bool):
                a = a.to(torch.int32)
            else:
                raise TypeError(f"{name}: actions must be integer, got {a.dtype}")
        if a.device != self.device:
            a = a.to(self.device, non_blocking=True)
        a = a.reshape(-1)
        if a.numel() != self.num_envs:
            raise ValueError(f"{name}: expected {self.num_envs} actions, got {a.numel()}")
        if not a.is_contiguous():
            a = a.contiguous()
        if self.validate_actions and a.numel():
            # the reference raises KeyError from action_dict[action] (merging_env.py:101,147)
            lo, hi = int(a.min()), int(a.max())            # host sync: debug option only
            if lo < 0 or hi >= nat.NUM_ACTIONS:
                raise KeyError(lo if lo < 0 else hi)
        return a

    _ACT_DTYPE = {torch.uint8: nat.ACT_U8, torch.int32: nat.ACT_I32, torch.int64: nat.ACT_I64}

    # ------------------------------------------------------------------ gym API
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`MergeEnv.reset()` (merging_env.py:208-230) for all envs, or where `mask` is set.

        Returns obs f32[N,10] (current slot).  Start state per `reset_mode`: fixed pos=50, vel=20 for
        both cars (the reference's live code), or the random start of its commented-out lines.
        """
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).reshape(-1).contiguous()
            m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)     # bool is one byte: no conversion kernel
            if m.numel() != self.num_envs:
                raise ValueError("mask must have num_envs elements")
        obs = self.obs_buf[self._slot]
        self._join_lanes()
        with torch.cuda.device(self.device), nvtx_range("mg.reset"):
            nat.check(self._lib.mg_reset(C.byref(self._state), self.num_envs, _ptr(m), _ptr(obs),
                                         nat.OBS_LAYOUT_FLAG[self.obs_layout], C.byref(self._rs), self._stream()), "mg_reset")
        return obs

    # ------------------------------------------------------------------ lanes
    def _init_lanes(self, L: int) -> None:
        n = self.num_envs
        per = ((n + L - 1) // L + 255) // 256 * 256 if L > 1 else n       # whole blocks: every sub-array stays 16-byte aligned
        bounds = [(min(l * per, n), min((l + 1) * per, n)) for l in range(L)]
        self.lane_bounds = [b for b in bounds if b[1] > b[0]] or [(0, n)]
        self.lanes = len(self.lane_bounds)
        self.lane_slices = [slice(lo, hi) for lo, hi in self.lane_bounds]
        self._lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)] if self.lanes > 1 else []
        self._lane_ev = [None] * self.lanes          # event after the lane's last launch (None: nothing outstanding)
        self._lane_slot = [0] * self.lanes
        self._lane_pending = [None] * self.lanes
        if self.lanes == 1:
            return
        st = (self.pos1, self.vel1, self.pos2, self.vel2, self.ret1, self.ret2, self.meta)
        self._lane_state, self._lane_outs, self._lane_rs = [], [], []
        for lo, hi in self.lane_bounds:
            self._lane_state.append(nat.MgState(*[None if t is None else t[lo:hi].data_ptr() for t in st]))
            self._lane_outs.append([nat.MgOut(self.obs_buf[k, lo:hi].data_ptr(), self.rew_buf[k, lo:hi].data_ptr(),
                                              self.done_buf[k, lo:hi].data_ptr(), self.info_buf[k, lo:hi].data_ptr(),
                                              *[None if t is None else t[lo:hi].data_ptr() for t in
                                                (self.terminal_obs, self.episode_return, self.episode_length)])
                                    for k in range(self.out_slots)])
            self._lane_rs.append(nat.MgResetSpec(self._rs.mode, 0, self._rs.seed, self.env_id_base + lo))

    def join_lanes(self) -> None:
        """Order the current stream behind every outstanding lane launch (call before reading `env.pos1` etc. after
        `step_lane_async`; the env's own whole-env calls do it themselves)."""
        self._join_lanes()

    def _join_lanes(self) -> None:
        cur = None
        for l, ev in enumerate(self._lane_ev):
            if ev is not None:
                cur = cur or torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                self._lane_ev[l] = None

    def step_lane_async(self, lane: int, a1, a2=None) -> None:
        """`step` for the envs of one lane (`lane_slices[lane]`), launched on the lane's own stream behind
        everything queued on the current stream so far (so actions computed there are visible)."""
        if self.lanes == 1:
            return self.step_async(a1, a2)
        lo, hi = self.lane_bounds[lane]
        a1 = self._as_lane_action(a1, hi - lo, "action1")
        if a2 is not None:
            a2 = self._as_lane_action(a2, hi - lo, "action2")
            if a2.dtype != a1.dtype:
                wide = a1.dtype if a1.element_size() >= a2.element_size() else a2.dtype
                a1, a2 = a1.to(wide), a2.to(wide)
        k = self._lane_slot[lane] = (self._lane_slot[lane] + 1) % self.out_slots
        cur, ls = torch.cuda.current_stream(self.device), self._lane_streams[lane]
        fork = torch.cuda.Event()
        fork.record(cur)
        ls.wait_event(fork)
        with torch.cuda.device(self.device), nvtx_range("mg.step[lane]"):
            nat.check(self._lib.mg_step(C.byref(self._lane_state[lane]), hi - lo, _ptr(a1), _ptr(a2),
                                        self._ACT_DTYPE[a1.dtype], C.byref(self._rw), C.byref(self._lane_outs[lane][k]),
                                        _ptr(self.stats_buf), self._flags(), C.byref(self._lane_rs[lane]),
                                        C.c_void_p(ls.cuda_stream)), "mg_step (lane)")
        ev = torch.cuda.Event()
        ev.record(ls)
        self._lane_ev[lane] = ev
        self._lane_pending[lane] = k

    def step_lane_wait(self, lane: int):
        """Orders the current stream behind the lane's last launch and returns that step's
        (obs, rewards, done, info) for the lane's envs (views of the env-wide output slot)."""
        if self.lanes == 1:
            return self.step_wait()
        k = self._lane_pending[lane]
        if k is None:
            raise RuntimeError("step_lane_wait() called without step_lane_async()")
        self._lane_pending[lane] = None
        if self._lane_ev[lane] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._lane_ev[lane])
            self._lane_ev[lane] = None
        sl = self.lane_slices[lane]
        extras = {q: t[sl] for q, t in self._extras.items()}
        return (self.obs_buf[k, sl], self.rew_buf[k, sl], self.done_buf[k, sl].view(torch.bool),
                StepInfo(self.info_buf[k, sl], extras))

    def _as_lane_action(self, a, m, name):
        if not isinstance(a, torch.Tensor) or a.device != self.device or a.dtype not in self._ACT_DTYPE:
            raise TypeError(f"{name}: lane actions must be uint8 / int32 / int64 tensors on {self.device}")
        a = a.reshape(-1)
        if a.numel() != m:
            raise ValueError(f"{name}: expected {m} actions for this lane, got {a.numel()}")
        if not a.is_contiguous() or a.data_ptr() % 16:
            a = a.contiguous().clone() if a.data_ptr() % 16 else a.contiguous()
        return a

    def step_async(self, a1, a2=None) -> None:
        a1 = self._as_action(a1, "action1")
        if a2 is not None:
            a2 = self._as_action(a2, "action2")
            if a2.dtype != a1.dtype:                 # one element type per launch: widen, never narrow (256 must stay bad)
                wide = a1.dtype if a1.element_size() >= a2.element_size() else a2.dtype
                a1, a2 = a1.to(wide), a2.to(wide)
        if self.lanes > 1:
            # lane slots advance together here, so the env-wide views of slot k are whole
            for l, sl in enumerate(self.lane_slices):
                self._lane_slot[l] = self._slot
                self.step_lane_async(l, a1[sl], None if a2 is None else a2[sl])
            self._slot = self._lane_slot[0]
            self._pending = self._slot
            return
        self._slot = (self._slot + 1) % self.out_slots
        k = self._slot
        with torch.cuda.device(self.device), nvtx_range("mg.step"):
            nat.check(self._lib.mg_step(C.byref(self._state), self.num_envs, _ptr(a1), _ptr(a2),
                                        self._ACT_DTYPE[a1.dtype], C.byref(self._rw),
                                        C.byref(self._outs[k]), _ptr(self.stats_buf), self._step_flags(),
                                        C.byref(self._rs), self._stream()), "mg_step")
        self._pending = k

    def step_wait(self):
        k = self._pending
        if k is None:
            raise RuntimeError("step_wait() called without step_async()")
        self._pending = None
        if self.lanes > 1:
            self._join_lanes()
            self._lane_pending = [None] * self.lanes
        return (self.obs_buf[k], self.rew_buf[k], self.done_buf[k].view(torch.bool),
                StepInfo(self.info_buf[k], self._extras))

    def step(self, a1, a2=None):
        """`MergeEnv.step(action1, action2=None)` (merging_env.py:138-195) for all envs.

        a1, a2: integer tensors/arrays of N actions in 0..4 (uint8, int32 or int64; device tensors
        are used in place).  Returns (obs f32[N,10], rewards f32[N,2], done bool[N], info) — device
        tensors that alias this env's output slot; they are overwritten `out_slots` steps later.
        """
        self.step_async(a1, a2)
        return self.step_wait()

    # ------------------------------------------------------------------ policy in the loop, one launch per step
    def policy_step(self, policy, goal: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None,
                    explore=None, actions_out: Optional[torch.Tensor] = None, q_out: Optional[torch.Tensor] = None):
        """One iteration of the reference scripts' inner loop (scripts/main.py:194-211, hdqn.py:288-316)
            action = policy(obs)  [exploration rule]  ->  obs, rewards, done, info = env.step(action, a2)
        in ONE kernel (`mg_policy_step`): the Q-network forward + arg-max with `MergeEnv.step` as its epilogue.  The
        policy reads the env's current observation buffer; the next observation goes into the next output slot
        (with `out_slots=1`: in place).  Bit-identical to `policy.act(obs)` followed by `step(...)`.

        policy       an `MLPPolicy` with 5 outputs and backend "fused" (fp32), "tf32x3" or "f16x3" (tensor cores)
        goal         uint8[N]: the h-DQN controller's `[goal] + state` column (`MLPPolicy(11, 5)`, hdqn.py:291)
        a2           uint8[N] actions of player 2, None = pve (the constant-speed opponent, merging_env.py:152)
        explore      None = greedy; a `policy.Exploration` = the scripts' `randn() <= EPISILO` rule (main.py:103-110)
                     drawn on the device from Philox over (seed, global env id, step)
        actions_out  uint8[N]: receives the action taken (what `store_transition` records)
        Returns (obs, rewards, done, info) like `step`.
        """
        from .policy import POLICY_BACKENDS
        if policy.out_dim != nat.NUM_ACTIONS or policy.backend not in POLICY_BACKENDS:
            raise ValueError("policy_step needs an MLPPolicy with 5 outputs and backend 'fused', 'tf32x3' or 'f16x3'")
        goal_in_slot = goal is None and self.obs_layout == "goal_slot" and policy.in_dim == nat.OBS_DIM + 1
        if policy.in_dim != nat.OBS_DIM + (0 if goal is None else 1) and not goal_in_slot:
            raise ValueError("policy input width does not match obs (+ goal)")
        n = self.num_envs

        def u8(t, name):
            if t is None:
                return None
            if not (isinstance(t, torch.Tensor) and t.dtype == torch.uint8 and t.device == self.device
                    and t.is_contiguous() and t.numel() == n):
                raise ValueError(f"{name} must be a contiguous uint8 tensor of {n} elements on {self.device}")
            return t
        goal, a2, actions_out = u8(goal, "goal"), u8(a2, "a2"), u8(actions_out, "actions_out")
        if q_out is not None and (tuple(q_out.shape) != (n, nat.NUM_ACTIONS) or q_out.dtype != torch.float32
                                  or not q_out.is_contiguous() or q_out.device != self.device):
            raise ValueError("q_out must be a contiguous float32 [N,5] tensor on the env's device")
        self._join_lanes()
        obs_in = self.obs_buf[self._slot]
        self._slot = (self._slot + 1) % self.out_slots
        k = self._slot
        flags = self._step_flags()
        if goal is None and policy.in_dim == nat.OBS_DIM + 1:
            flags |= nat.POLICY_FLAG_GOAL_IN_SLOT        # the controller reads `[goal] + state` rows as they are
        ex = None
        if explore is not None:
            ex = explore.spec()
            flags |= nat.POLICY_FLAG_EXPLORE
        if policy.pdl:
            flags |= nat.POLICY_FLAG_PDL
        w2 = policy.w2_native
        with torch.cuda.device(self.device), nvtx_range("mg.policy_step"):
            nat.check(self._lib.mg_policy_step(C.byref(self._state), n, _ptr(obs_in), _ptr(goal), POLICY_BACKENDS[policy.backend],
                                               _ptr(policy.w1_t), _ptr(policy.b1), _ptr(w2), _ptr(policy.b2), _ptr(policy.w3),
                                               _ptr(policy.b3), _ptr(a2), C.byref(self._rw), C.byref(self._outs[k]),
                                               _ptr(self.stats_buf), flags, C.byref(self._rs),
                                               None if ex is None else C.byref(ex), _ptr(actions_out), _ptr(q_out),
                                               self._stream()), "mg_policy_step")
        return (self.obs_buf[k], self.rew_buf[k], self.done_buf[k].view(torch.bool), StepInfo(self.info_buf[k], self._extras))

    def close(self):
        self.closed = True

    def seed(self, seed=None):
        """Sets the Philox key of the synthetic action stream (the env itself has no randomness)."""
        if seed is not None:
            self.philox_seed = int(seed)
        return [seed]

    def show_reward(self):
        """merging_env.py:115-116."""
        r = self._rw
        return r.r_first, r.r_second, r.r_collision, r.vel_penalty

    # ------------------------------------------------------------------ reference attributes
    @property
    def winner(self) -> torch.Tensor:
        """u8[N]: 0 = None, 1, 2 (merging_env.py:212)."""
        return ((self.meta >> nat.META_WINNER_SHIFT) & 3).to(torch.uint8)

    @property
    def steps(self) -> torch.Tensor:
        """i32[N] steps since reset (time_stamp / dT, merging_env.py:141)."""
        return self.meta & nat.META_STEPS_MASK

    @property
    def done(self) -> torch.Tensor:
        return (self.meta & nat.META_DONE).bool()

    @property
    def resets(self) -> torch.Tensor:
        """i32[N] number of resets each env has had, mod 2^17 (the random-start Philox counter)."""
        return (self.meta >> nat.META_RESETS_SHIFT) & 0x1FFFF

    @property
    def r1_accumulate(self) -> torch.Tensor:
        if self.ret1 is None:
            raise AttributeError("r1_accumulate: the env was created with track_returns=False")
        return self.ret1

    @property
    def r2_accumulate(self) -> torch.Tensor:
        if self.ret2 is None:
            raise AttributeError("r2_accumulate: the env was created with track_returns=False")
        return self.ret2

    @staticmethod
    def opponent_view(obs: torch.Tensor) -> torch.Tensor:
        """`state[5:] + state[:5]` — the observation from player 2's seat (scripts/main.py:199)."""
        return torch.cat([obs[..., 5:], obs[..., :5]], dim=-1)

    # ------------------------------------------------------------------ synthetic actions / rollouts
    def sample_actions(self, step: Optional[int] = None):
        """Uniform-random actions for this shard at rollout step `step` (default: internal counter).

        Counter-based Philox4x32-10: key = seed, counter = (global env id, step).  Returns the
        internal uint8 device buffers (a1, a2); a2 is None in pve mode.
        """
        if step is None:
            step = self.step_count
            self.step_count += 1
        a2 = self.act2 if self.mode == "pvp" else None
        with torch.cuda.device(self.device):
            nat.check(self._lib.mg_sample_actions(_ptr(self.act1), _ptr(a2), self.num_envs, self.philox_seed,
                                                  self.env_id_base, int(step), self._stream()),
                      "mg_sample_actions")
        return self.act1, a2

    def rollout(self, k_steps: int, obs: Optional[torch.Tensor] = None, rew: Optional[torch.Tensor] = None,
                done: Optional[torch.Tensor] = None, info: Optional[torch.Tensor] = None,
                actions: Optional[torch.Tensor] = None, step0: Optional[int] = None, refresh_obs: bool = True):
        """`k_steps` random-action steps in ONE launch (state stays in registers between steps).

        Time-major output tensors are optional: obs f32[k,N,10], rew f32[k,N,2], done u8[k,N],
        info u8[k,N], actions u8[k,N,2].  The action stream equals `sample_actions` at steps
        step0..step0+k-1, so `rollout(k)` and k x (`sample_actions` + `step`) give identical results.
        `refresh_obs`: afterwards write the observation of the final state into the env's current
        observation buffer (`observe()`, one extra small launch), so that `env.obs_buf[slot]` is what a
        policy should act on next; pass False when only the time-major outputs are used.
        """
        n, k = self.num_envs, int(k_steps)
        if obs is not None:
            self._need_rows("rollout(obs=...)")
        if step0 is None:
            step0 = self.step_count
            self.step_count += k

        def chk(t, shape, dtype, name):
            if t is None:
                return None
            if tuple(t.shape) != shape or t.dtype != dtype or t.device != self.device or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dtype} tensor of shape {shape} on {self.device}")
            return t
        chk(obs, (k, n, nat.OBS_DIM), torch.float32, "obs"); chk(rew, (k, n, 2), torch.float32, "rew")
        chk(done, (k, n), torch.uint8, "done"); chk(info, (k, n), torch.uint8, "info")
        chk(actions, (k, n, 2), torch.uint8, "actions")
        out = nat.MgOut(*[None if t is None else t.data_ptr() for t in
                          (obs, rew, done, info, self.terminal_obs, self.episode_return, self.episode_length)])
        self._join_lanes()
        with torch.cuda.device(self.device), nvtx_range(f"mg.rollout[{k}]"):
            nat.check(self._lib.mg_rollout(C.byref(self._state), n, int(self.mode == "pvp"), self.philox_seed,
                                           self.env_id_base, int(step0), k, C.byref(self._rw),
                                           C.byref(out), _ptr(actions), _ptr(self.stats_buf),
                                           self._flags(), C.byref(self._rs), self._stream()), "mg_rollout")
        if refresh_obs and k > 0:
            self.refresh_observation()

    def refresh_observation(self) -> torch.Tensor:
        """`observe()` (merging_env.py:118-132) of the current state into the current observation buffer: an
        `mg_reset` with an all-zero mask resets nothing and writes every env's observation."""
        if self._zero_mask is None:
            self._zero_mask = torch.zeros(max(self.num_envs, 1), dtype=torch.uint8, device=self.device)
        obs = self.obs_buf[self._slot]
        self._join_lanes()
        with torch.cuda.device(self.device):
            nat.check(self._lib.mg_reset(C.byref(self._state), self.num_envs, _ptr(self._zero_mask), _ptr(obs),
                                         nat.OBS_LAYOUT_FLAG[self.obs_layout], C.byref(self._rs), self._stream()),
                      "mg_reset (observe)")
        return obs

    # ------------------------------------------------------------------ host-buffer path
    def _host_buffers(self) -> dict:
        """Pinned mirrors of host slot 0 (the synchronous `step_host` path uses only this one)."""
        return self._host_slots()[0]

    def _host_slots(self) -> list:
        if self._hslots is None:
            # Per host slot: pinned mirrors with the device layout (actions: [a1 | a2]; outputs: [obs | rew | done |
            # info]) so that one copy per direction suffices, a device output block of its own (the asynchronous
            # path's kernel may run while the env's regular output slots are still being read), and two events.
            n, H, dev = self.num_envs, self.host_slots, self.device
            n_pad = (n + 15) // 16 * 16
            self._hd_block = torch.zeros(H * n_pad * 50, dtype=torch.uint8, device=dev)
            d_obs, d_rew, d_done, d_info = _slot_views(self._hd_block, H, n, n_pad)
            self._copy_stream = self._copy_stream or torch.cuda.Stream(device=dev)
            self._upload_stream = torch.cuda.Stream(device=dev)
            slots = []
            for k in range(H):
                acts = torch.zeros(2 * n_pad, dtype=torch.uint8, pin_memory=True)
                outs = torch.zeros(n_pad * 50, dtype=torch.uint8, pin_memory=True)
                obs, rew, done, info = _slot_views(outs, 1, n, n_pad)
                h = dict(a1=acts[:n], a2=acts[n_pad:n_pad + n], obs=obs[0], rew=rew[0], done=done[0], info=info[0],
                         _blocks=(acts, outs))
                h["np_acts"] = (h["a1"].numpy(), h["a2"].numpy())
                h["np_out"] = (h["obs"].numpy(), h["rew"].numpy(), h["done"].numpy().view(np.bool_), h["info"].numpy())
                h["h_out"] = nat.MgOut(h["obs"].data_ptr(), h["rew"].data_ptr(), h["done"].data_ptr(),
                                       h["info"].data_ptr(), None, None, None)
                h["d_out"] = nat.MgOut(d_obs[k].data_ptr(), d_rew[k].data_ptr(), d_done[k].data_ptr(), d_info[k].data_ptr(),
                                       *[None if t is None else t.data_ptr() for t in
                                         (self.terminal_obs, self.episode_return, self.episode_length)])
                h["d_acts"] = torch.zeros(2 * n_pad, dtype=torch.uint8, device=dev)    # [a1 | a2] like the pinned side
                ev = (torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event())
                with torch.cuda.device(dev):
                    for e in ev:                     # torch creates the CUDA event lazily, on its first record
                        e.record(self._copy_stream)
                h["ev"] = ev
                h["slot"] = nat.MgHostSlot(h["a1"].data_ptr(), h["a2"].data_ptr(), h["d_acts"].data_ptr(),
                                           h["d_acts"][n_pad:].data_ptr(), h["d_out"], h["h_out"],
                                           ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event)
                h["slot_pve"] = nat.MgHostSlot(h["a1"].data_ptr(), None, h["d_acts"].data_ptr(), None, h["d_out"], h["h_out"],
                                               ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event)
                h["fields"] = nat.FIELD_ALL
                slots.append(h)
            self._copy_stream.synchronize()
            self._hslots = slots
            self._host = slots[0]
            self._host_np = slots[0]["np_acts"]
            self._host_out = slots[0]["h_out"]
        return self._hslots

    def _copy_stream_ptr(self, chunks: int):
        if chunks <= 1:
            return None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return C.c_void_p(self._copy_stream.cuda_stream)

    def host_action_buffers(self, slot: Optional[int] = None):
        """(a1, a2): uint8[N] NumPy views of the PINNED staging buffers the step kernel reads its actions from.  A
        caller that writes its actions into them and passes them back to `step_host` / `step_host_async` saves the
        extra host-side copy.  `slot=None`: the slot the next `step_host_async` (or `step_host`) call will use."""
        slots = self._host_slots()
        return slots[self._hnext if slot is None else slot]["np_acts"]

    @staticmethod
    def _field_mask(fields) -> int:
        if fields is None:
            return nat.FIELD_ALL
        m = 0
        for f in ([fields] if isinstance(fields, str) else fields):
            if f not in nat.FIELD_BITS:
                raise ValueError(f"unknown field {f!r}: choose from {sorted(nat.FIELD_BITS)}")
            m |= nat.FIELD_BITS[f]
        if not m:
            raise ValueError("fields must name at least one of obs / rew / done / info")
        return m

    def step_host_async(self, a1: np.ndarray, a2: Optional[np.ndarray] = None, fields=None, upload: bool = True) -> None:
        """Pipelined host-buffer step (`mg_step_host_async`): queues the upload of the uint8 actions from this slot's
        pinned buffers (one cudaMemcpyAsync on a private upload stream; `upload=False`: the kernel reads them straight
        from pinned host memory instead, which costs the concurrent device-to-host copy 1-3 % on the boxes measured,
        `profiles/e2e_pipeline_probe.py`), the fused step, and the device-to-host copies of the selected `fields`
        (any of "obs", "rew", "done", "info"; default all four) on a private copy stream, and returns at once.
        Up to `host_slots` steps may be in flight; `step_host_wait()` hands back the oldest.  With two slots the
        kernel and action fetch of step t+1 run under the copies of step t, so the PCIe link never idles.
        The reference's contract per call is `return obs, rewards, done, info` (merging_env.py:195): a caller that
        only logs rewards / dones (its policy reading the device-resident observation) moves 10 B per env instead of 50.
        """
        self._need_rows("step_host_async")
        slots = self._host_slots()
        self._join_lanes()
        if len(self._hfly) >= self.host_slots:
            raise RuntimeError(f"{self.host_slots} host steps already in flight: call step_host_wait() first")
        k = self._hnext
        h = slots[k]
        mask = self._field_mask(fields)
        npa = h["np_acts"]
        # actions written in place into `host_action_buffers()` are used as they are; anything else is staged
        # into those pinned buffers first (an extra host copy of n bytes per player)
        if a1 is not npa[0]:
            npa[0][:] = np.asarray(a1, dtype=np.uint8).reshape(-1)
        if a2 is not None and a2 is not npa[1]:
            npa[1][:] = np.asarray(a2, dtype=np.uint8).reshape(-1)
        with torch.cuda.device(self.device), nvtx_range("mg.step_host_async"):
            nat.check(self._lib.mg_step_host_async(C.byref(self._state), self.num_envs,
                                                   C.byref(h["slot"] if a2 is not None else h["slot_pve"]), mask,
                                                   C.byref(self._rw), _ptr(self.stats_buf), self._flags(), C.byref(self._rs),
                                                   self._stream(), C.c_void_p(self._copy_stream.cuda_stream),
                                                   C.c_void_p(self._upload_stream.cuda_stream) if upload else None),
                      "mg_step_host_async")
        h["fields"] = mask
        self._hfly.append(k)
        self._hnext = (k + 1) % self.host_slots

    def step_host_wait(self):
        """Blocks until the oldest in-flight `step_host_async` has landed in host memory; returns
        (obs, rewards, done, info_flags) as NumPy views of that slot's pinned buffers — None for fields that were
        not requested.  The views are overwritten when the slot is reused, `host_slots` calls later."""
        if not self._hfly:
            raise RuntimeError("step_host_wait() called without step_host_async()")
        h = self._hslots[self._hfly.pop(0)]
        nat.check(self._lib.mg_step_host_wait(C.c_void_p(h["ev"][2].cuda_event)), "mg_step_host_wait")
        m = h["fields"]
        return tuple(v if m & (1 << i) else None for i, v in enumerate(h["np_out"]))

    def step_host(self, a1: np.ndarray, a2: Optional[np.ndarray] = None, zero_copy: bool = False,
                  chunks: int = 1, direct_actions: bool = True, fields=None):
        """Drop-in for host-resident callers: uint8 NumPy actions in, NumPy outputs out.

        Default: one `mg_step_host` call = the fused step reading the actions straight from the pinned
        staging buffers (`direct_actions=False`: after an explicit H2D copy of them instead), D2H of
        obs/rew/done/info into pinned host buffers, stream synchronise.  `fields` (any of "obs", "rew",
        "done", "info") restricts the copy-back; it is served by `step_host_async` + `step_host_wait`.
        `chunks > 1` steps the envs in that many pieces so that the device-to-host copy of one piece (on a
        private copy stream) overlaps the upload and the kernel of the next; on a PCIe 5 x16 B200 this does
        not pay (the upload + kernel it can hide are 0.07 ms of 1.04 ms, the extra small copies cost more:
        `profiles/e2e_paths.py`), so the default is one piece.
        `zero_copy=True`: `mg_step` is handed the pinned host buffers themselves (pinned memory is
        device-addressable under UVA), so the kernel reads the actions and streams its outputs
        across PCIe while it computes — no staging copy in HBM, no separate memcpy.
        Returns (obs, rewards, done, info_flags) as NumPy views of the pinned buffers (overwritten
        by the next call).
        """
        self._need_rows("step_host")
        if self._hfly:
            raise RuntimeError("step_host() while step_host_async() calls are in flight: call step_host_wait() first")
        if fields is not None:
            self._hnext = 0
            self.step_host_async(a1, a2, fields, upload=not direct_actions)
            self._hnext = 0
            return self.step_host_wait()
        n = self.num_envs
        h = self._host_buffers()
        self._join_lanes()
        self._hnext = 0
        if a1 is not self._host_np[0]:
            self._host_np[0][:] = np.asarray(a1, dtype=np.uint8).reshape(-1)
        if a2 is not None and a2 is not self._host_np[1]:
            self._host_np[1][:] = np.asarray(a2, dtype=np.uint8).reshape(-1)
        with torch.cuda.device(self.device), nvtx_range("mg.step_host"):
            if zero_copy:
                nat.check(self._lib.mg_step(C.byref(self._state), n, _ptr(h["a1"]),
                                            _ptr(h["a2"]) if a2 is not None else None, nat.ACT_U8,
                                            C.byref(self._rw), C.byref(self._host_out), _ptr(self.stats_buf),
                                            self._flags(), C.byref(self._rs), self._stream()), "mg_step (zero-copy)")
                torch.cuda.current_stream(self.device).synchronize()
            else:
                self._slot = (self._slot + 1) % self.out_slots
                nat.check(self._lib.mg_step_host(C.byref(self._state), n, _ptr(h["a1"]),
                                                 _ptr(h["a2"]) if a2 is not None else None,
                                                 None if direct_actions else _ptr(self.act1),
                                                 None if direct_actions else _ptr(self.act2), C.byref(self._rw),
                                                 C.byref(self._outs[self._slot]), C.byref(self._host_out),
                                                 _ptr(self.stats_buf), self._flags(), C.byref(self._rs),
                                                 self._stream(), self._copy_stream_ptr(chunks), int(chunks)),
                          "mg_step_host")
        return h["np_out"]

    # ------------------------------------------------------------------ statistics
    @property
    def stats_buf(self) -> Optional[torch.Tensor]:
        """int64[1024,16] partial rows the kernels launched from now on add into (the active bank)."""
        return None if self._stats_banks is None else self._stats_banks[self._stats_active]

    def _retire_stats_bank(self) -> torch.Tensor:
        """Make the other bank active and return the retired one (banked `AsyncStatsReducer` only).  Kernels already
        queued — and CUDA graphs captured earlier — keep adding into the bank that was active when they were launched
        or captured."""
        if len(self._stats_banks) == 1:
            self._stats_banks.append(torch.zeros_like(self._stats_banks[0]))
        retired = self._stats_banks[self._stats_active]
        self._stats_active ^= 1
        return retired

    def stats_tensor(self) -> torch.Tensor:
        """int64[16] device totals (retired banks' totals + the sum of every bank's partial rows); no host sync."""
        if self._stats_banks is None:
            raise RuntimeError("track_stats=False")
        self._join_lanes()
        if self._stats_side_event is not None:       # a reducer may still be draining a bank on its side stream
            torch.cuda.current_stream(self.device).wait_event(self._stats_side_event)
        t = self._stats_total.clone()
        for b in self._stats_banks:
            t += b.sum(dim=0)
        return t

    def stats(self, reduce: bool = False, reset: bool = False) -> dict:
        """Episode statistics as a dict (host sync).  `reduce=True` sums over all ranks first."""
        from .sharding import all_reduce_stats, stats_to_dict
        t = self.stats_tensor()
        if reduce:
            t = all_reduce_stats(t)
        d = stats_to_dict(t.cpu().numpy(), self.constants.return_fixed_point_scale)
        if reset:
            self._stats_total.zero_()
            for b in self._stats_banks:
                b.zero_()
        return d

    # ------------------------------------------------------------------ checkpoint
    _STATE_KEYS = ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "meta")

    def state_dict(self) -> dict:
        """Everything a bit-exact continuation needs: the env state arrays, the Philox step counter and keys (action
        stream, random starts), the first global env id, and the statistics partials."""
        self._join_lanes()
        d = {k: getattr(self, k).clone() for k in self._STATE_KEYS if getattr(self, k) is not None}
        d["step_count"] = self.step_count
        d["philox_seed"], d["reset_seed"], d["env_id_base"] = self.philox_seed, int(self._rs.seed), self.env_id_base
        d["reset_mode"] = self.reset_mode
        if self._stats_banks is not None:
            st = self._stats_banks[0].clone()
            for b in self._stats_banks[1:]:
                st += b
            st[0] += self._stats_total
            d["stats"] = st
        return d

    def load_state_dict(self, d: dict) -> None:
        """Restores `state_dict()` and re-derives the current observation (`observe()` of the loaded state) into
        the current observation buffer."""
        self._join_lanes()
        for k in self._STATE_KEYS:
            if getattr(self, k) is not None:
                getattr(self, k).copy_(d[k])
        self.step_count = int(d.get("step_count", 0))
        self.philox_seed = int(d.get("philox_seed", self.philox_seed))
        self.env_id_base = int(d.get("env_id_base", self.env_id_base))
        if d.get("reset_mode", self.reset_mode) != self.reset_mode:
            raise ValueError(f"state_dict was saved with reset_mode={d['reset_mode']!r}, this env has {self.reset_mode!r}")
        self._rs.seed = int(d.get("reset_seed", self._rs.seed))
        self._rs.env_id_base = self.env_id_base
        if self._stats_banks is not None and "stats" in d:
            self._stats_total.zero_()
            for b in self._stats_banks:
                b.zero_()
            self.stats_buf.copy_(d["stats"])
        self.refresh_observation()

    def __repr__(self):
        return (f"MergeVecEnv(num_envs={self.num_envs}, mode={self.mode!r}, device={self.device}, "
                f"auto_reset={self.auto_reset})")
