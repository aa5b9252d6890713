"""Transition recording for learners and logs.

`TransitionRecorder` — device-resident ring of the reference's replay rows
`[s(10), a, r, s'(10)]` (`DQN.store_transition`, scripts/main.py:115-119), appended for every env
and step while `env.winner is not 1` (main.py:209-211) by the deterministic stream-compaction
kernels behind `mg_record_transitions`; nothing leaves the GPU.

`CsvEpisodeLogger` — host-side per-episode CSV files for a few selected envs in the column order
of scripts/human_player.py:111 (`[obs(10), action1, action2, reward1, reward2]`, rows written
while `env.winner is not 1`, :180-181), so the reference's analysis notebook can read GPU episodes.
"""
from __future__ import annotations

import csv
import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _native as nat
from .tracing import nvtx_range

REPLAY_WIDTH = 2 * nat.OBS_DIM + 2      # main.py:92  NUM_STATES * 2 + 2
LOG_WIDTH = nat.OBS_DIM + 4
HDQN_WIDTH = 2 * (nat.OBS_DIM + 1) + 2   # hdqn.py:158  (NUM_STATES + 1) * 2 + 2
_FORMATS = {"replay": (0, REPLAY_WIDTH), "log": (1, LOG_WIDTH), "hdqn": (2, HDQN_WIDTH)}
_MASKS = {"all": 0, "winner_not_1": 1, "explicit": 2}

# header of scripts/human_player.py:111, verbatim column names
CSV_HEADER = ["x2 - x1", "y2 - y1", "self.state2['vel'] - self.state1['vel']", "END_POINT - self.state1['pos']",
              "self.state1['vel']", "x1 - x2", "y1 - y2", "self.state1['vel'] - self.state2['vel']",
              "END_POINT - self.state2['pos']", "self.state2['vel']", "action1", "action2", "reward1", "reward2"]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class TransitionRecorder:
    """Ring buffer of transition rows on the device.

    format "replay": rows `[s, a_p, r_p, s']` (22 floats) for `player` p in {1, 2}.
    format "log":    rows `[s, a1, a2, r1, r2]` (14 floats).
    format "hdqn":   rows `[g, s, a_p, r_int, g', s']` (24 floats), the h-DQN controller's transitions
                     (hdqn.py:180-184,291-316): pass the goals chosen from s and from s' to `record`; the
                     intrinsic reward `1 if g' == goal_status(s) else 0` (hdqn.py:314) is computed on the device.
                     The reference stores these every step: use mask="all".
    mask "winner_not_1" is the reference's store condition; "all" stores every env every step; "explicit"
    stores the envs selected by the uint8 `select` tensor passed to `record`.
    """

    def __init__(self, env, capacity: int, format: str = "replay", player: int = 1,
                 mask: str = "winner_not_1", track_env_ids: bool = False):
        if format not in _FORMATS or mask not in _MASKS or player not in (1, 2):
            raise ValueError("format in {'replay','log','hdqn'}, mask in {'winner_not_1','all','explicit'}, player in {1,2}")
        if getattr(env, "obs_layout", "aos") != "aos":
            raise ValueError("the recorders read the default [N,10] observation rows: create the env with obs_layout='aos'")
        if format != "log" and env.auto_reset and env.terminal_obs is None:
            raise ValueError("replay rows need the terminal observation: create the env with episode_info=True")
        self.env, self.capacity = env, int(capacity)
        self.format, self.player, self.mask = format, player, mask
        self.width = _FORMATS[format][1]
        dev = env.device
        self.ring = torch.zeros(self.capacity, self.width, dtype=torch.float32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)        # memory_counter, main.py:91
        self.env_ids = torch.full((self.capacity,), -1, dtype=torch.int32, device=dev) if track_env_ids else None
        self._lib = nat.load()
        self._scratch = torch.zeros(int(self._lib.mg_record_scratch_words(env.num_envs)), dtype=torch.int32, device=dev)
        self._same_obs_ok = False       # OptionRecorder stores [s_end, goal, sum_r, s_end]: s == s' on purpose

    def record(self, obs_prev: torch.Tensor, a1: torch.Tensor, a2: Optional[torch.Tensor], step_out,
               goal_prev: Optional[torch.Tensor] = None, goal_next: Optional[torch.Tensor] = None,
               select: Optional[torch.Tensor] = None) -> None:
        """Append the transitions of one `env.step`: `obs_prev` is the observation the actions were
        chosen from, `step_out` the tuple `env.step` returned.  uint8 actions (and, for format "hdqn",
        uint8 goals chosen from `obs_prev` and from the new observation); no host sync."""
        obs, rew, done, info = step_out
        env = self.env
        term = env.terminal_obs if env.auto_reset else None
        for t in (a1, a2, goal_prev, goal_next, select):
            if t is not None and t.dtype != torch.uint8:
                raise TypeError("TransitionRecorder.record expects uint8 action / goal / select tensors")
        if (self.mask == "explicit") != (select is not None):
            raise ValueError("`select` is required for, and only for, mask='explicit'")
        if (self.format == "hdqn") != (goal_prev is not None and goal_next is not None):
            raise ValueError("goal_prev and goal_next are required for, and only for, format 'hdqn'")
        if not (isinstance(obs_prev, torch.Tensor) and obs_prev.dtype == torch.float32 and obs_prev.is_contiguous()
                and obs_prev.device == env.device and tuple(obs_prev.shape) == (env.num_envs, nat.OBS_DIM)):
            raise ValueError(f"obs_prev must be a contiguous float32[{env.num_envs}, {nat.OBS_DIM}] tensor on {env.device} "
                             "(the kernel reads it with 128-bit loads)")
        if self.format != "log" and not self._same_obs_ok and obs_prev.data_ptr() == obs.data_ptr():
            # with out_slots=1 `env.step` returns the buffer it was given: s and s' would be the same row
            raise ValueError("obs_prev aliases the new observation (the env's single output slot was overwritten by this "
                             "step): create the env with out_slots >= 2 or pass a copy of the previous observation")
        with torch.cuda.device(env.device), nvtx_range("mg.record"):
            nat.check(self._lib.mg_record_transitions(
                _ptr(obs_prev), _ptr(obs), _ptr(term), _ptr(a1), _ptr(a2), _ptr(rew),
                _ptr(done.view(torch.uint8)), _ptr(select if select is not None else info["flags"]),
                _ptr(goal_prev), _ptr(goal_next), env.num_envs,
                _MASKS[self.mask], _FORMATS[self.format][0], self.player,
                _ptr(self.ring), self.capacity, _ptr(self.env_ids), _ptr(self.counter), _ptr(self._scratch),
                C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)), "mg_record_transitions")

    def __len__(self) -> int:
        """Rows currently valid (host sync)."""
        return min(int(self.counter.item()), self.capacity)

    def rows(self) -> torch.Tensor:
        """Valid rows, oldest first (host sync)."""
        c = int(self.counter.item())
        if c <= self.capacity:
            return self.ring[:c]
        k = c % self.capacity
        return torch.cat([self.ring[k:], self.ring[:k]], dim=0)

    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """`np.random.choice(MEMORY_CAPACITY, BATCH_SIZE)` (main.py:130) over a full ring, on device."""
        idx = torch.randint(0, self.capacity, (batch_size,), device=self.ring.device, generator=generator)
        return self.ring[idx]


class OptionRecorder:
    """The h-DQN meta-controller's transitions (scripts/hdqn.py:283-320) for N envs on the device.

    An option runs from one goal choice until `done or goal == goal_status(state)` (hdqn.py:316); the reference
    then stores `upper.store_transition(state, goal, extrinsic_reward, next_state)` (:318) where — because
    `state = next_state` has already been executed (:315) — BOTH observations are the one the option ended in, `goal`
    is the goal re-chosen from it, and `extrinsic_reward` is the sum of the ego rewards over the option.  Rows:
    `[s_end(10), goal, sum_r, s_end(10)]`, 22 floats, appended in env-id order like every other ring here.
    """

    def __init__(self, env, capacity: int, track_env_ids: bool = False):
        self.env = env
        self.rec = TransitionRecorder(env, capacity, format="replay", player=1, mask="explicit", track_env_ids=track_env_ids)
        self.rec._same_obs_ok = True
        n, dev = env.num_envs, env.device
        self.extrinsic = torch.zeros(n, dtype=torch.float32, device=dev)
        self._rew = torch.zeros(n, 2, dtype=torch.float32, device=dev)
        self._s_end = torch.zeros(n, nat.OBS_DIM, dtype=torch.float32, device=dev)
        self._ended = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._not_done = torch.zeros(n, dtype=torch.uint8, device=dev)

    @property
    def ring(self):
        return self.rec.ring

    @property
    def counter(self):
        return self.rec.counter

    def record(self, step_out, goal_next: torch.Tensor) -> torch.Tensor:
        """Call after every env step with the goal chosen from the NEW state; returns the uint8[N] mask of the
        envs whose option ended in this step (a buffer that the next call overwrites).  Two launches —
        `mg_option_update` (end-of-option test, reward sum, row observation) and the ring append — no host sync."""
        obs, rew, done, info = step_out
        env = self.env
        if goal_next.dtype != torch.uint8 or not goal_next.is_contiguous():
            raise TypeError("goal_next must be a contiguous uint8 tensor")
        term = env.terminal_obs if env.auto_reset else None      # the stepped state's observation, not the reset one
        with torch.cuda.device(env.device):
            nat.check(self.rec._lib.mg_option_update(_ptr(obs), _ptr(term), _ptr(rew), _ptr(done.view(torch.uint8)),
                                                     _ptr(goal_next), env.num_envs, _ptr(self.extrinsic), _ptr(self._s_end),
                                                     _ptr(self._rew), _ptr(self._ended),
                                                     C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)),
                      "mg_option_update")
        # row = [state, goal, extrinsic_reward, next_state] with state == next_state == s_end (hdqn.py:315-318)
        self.rec.record(self._s_end, goal_next, None, (self._s_end, self._rew, self._not_done.view(torch.bool), info),
                        select=self._ended)
        return self._ended


class CsvEpisodeLogger:
    """Per-episode CSV logs of selected envs in the reference's `human_player.py` format."""

    def __init__(self, env, env_ids: Sequence[int], directory: str, prefix: str = "episode"):
        self.env = env
        self.ids = torch.as_tensor(list(env_ids), dtype=torch.int64, device=env.device)
        self.directory, self.prefix = directory, prefix
        os.makedirs(directory, exist_ok=True)
        self._rows = {int(i): [] for i in env_ids}
        self._episode = {int(i): 0 for i in env_ids}
        self.files = []

    def log(self, obs_prev: torch.Tensor, a1, a2, step_out) -> None:
        """Call once per `env.step` with the observation the actions were chosen from (host sync)."""
        obs, rew, done, info = step_out
        ids = self.ids
        s = obs_prev.index_select(0, ids).cpu().tolist()
        act1 = torch.as_tensor(a1, device=obs.device).index_select(0, ids).cpu().tolist()
        act2 = [None] * len(s) if a2 is None else torch.as_tensor(a2, device=obs.device).index_select(0, ids).cpu().tolist()
        r = rew.index_select(0, ids).cpu().tolist()
        w = info["winner"].index_select(0, ids).cpu().tolist()
        d = done.index_select(0, ids).cpu().tolist()
        for k, e in enumerate(ids.cpu().tolist()):
            if w[k] != 1:                                   # human_player.py:180 `if env.winner is not 1`
                self._rows[e].append(s[k] + [act1[k], act2[k]] + r[k])
            if d[k]:
                self._flush(e)

    def _flush(self, e: int) -> None:
        path = os.path.join(self.directory, f"{self.prefix}{self._episode[e]} env{e}.csv")
        with open(path, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(CSV_HEADER)
            wr.writerows(self._rows[e])
        self.files.append(path)
        self._rows[e] = []
        self._episode[e] += 1

    def close(self) -> None:
        for e, rows in self._rows.items():
            if rows:
                self._flush(e)
