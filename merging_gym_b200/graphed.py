"""K policy-in-the-loop steps per host call ("multi-step rollout with fused greedy policy").

`GraphedPolicyRollout` captures K iterations of
    a1 = policy1(obs)            [a2 = policy2(opponent's view of obs)]        -> env.step -> [recorder.record]
into one CUDA graph: a replay advances every env K steps with no host work in between, which is what makes the
policy loop launch-latency free at small env counts.  The loop body is exactly the reference scripts' inner loop
(scripts/main.py:194-211, hdqn.py:288-316) with greedy (or `explore`d) actions.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .vec_env import MergeVecEnv


class GraphedPolicyRollout:
    """`run()` = K env steps.  The env must have been constructed with `out_slots=1` (one fixed observation buffer,
    so that every captured step reads the buffer the previous one wrote) — or, with `fused=True`, any `out_slots`
    that divides K.

    fused=True: player 1's step is ONE launch, `MergeVecEnv.policy_step` (`mg_policy_step`: forward + arg-max +
    exploration + env step), instead of policy launch + action copy + `mg_step`; `policy1` must then be an `MLPPolicy`
    (backend "fused", "tf32x3" or "f16x3") or an `HDQNPolicy` (two launches: goal, then controller + env).  `explore` (an
    `Exploration`) applies the scripts' `randn() <= EPISILO` rule on the device.  The single launch pays off where launches
    dominate — up to roughly 65 536 envs (f16x3: 9.1 vs 12.3 us per step at 4096 envs); from there up the two-launch loop is
    as fast or faster (profiles/r02_policy_small_batches.jsonl).  With `after_step` the env needs
    `out_slots >= 2`, so that the observation the actions were chosen from is still intact when the recorder reads it.

    policy1(obs) -> uint8[N] actions of player 1; policy2 (pvp only) receives the same observation buffer and must
    mirror it itself (`MLPPolicy.act(obs, mirror=True)`).  `after_step(obs_prev, a1, a2, step_out)`, if given, runs
    inside the graph after every step (e.g. `TransitionRecorder.record`); `obs_prev` is a graph-private copy of the
    observation the actions were chosen from.  Construction runs `warmup_steps` real steps (they advance the env and
    feed `after_step`) before the capture; pass `warmup_steps=0` once the kernels have been used before.
    """

    def __init__(self, env: MergeVecEnv, policy1: Callable, policy2: Optional[Callable] = None, k_steps: int = 32,
                 after_step: Optional[Callable] = None, warmup_steps: int = 3, fused: bool = False, explore=None):
        self.fused, self.explore = bool(fused), explore
        if fused:
            if int(k_steps) % env.out_slots:
                raise ValueError("fused GraphedPolicyRollout: k_steps must be a multiple of env.out_slots")
            if after_step is not None and env.out_slots < 2:
                raise ValueError("fused GraphedPolicyRollout with after_step needs an env with out_slots >= 2")
            if explore is not None and not hasattr(explore, "spec"):
                raise ValueError("explore must be a merging_gym_b200.Exploration")
        elif env.out_slots != 1:
            raise ValueError("GraphedPolicyRollout needs an env with out_slots=1")
        elif explore is not None:
            raise ValueError("explore is applied by the fused step: pass fused=True")
        if (policy2 is not None) != (env.mode == "pvp"):
            raise ValueError("policy2 is required for, and only for, a pvp env")
        self.env, self.k_steps = env, int(k_steps)
        # Inside this loop the kernel in front of a policy kernel is the env step (or the recorder), never one that writes
        # the weights, so the policy kernels are captured with programmatic dependent launch; the policies' own `pdl`
        # setting is put back after the capture (eager calls behind e.g. load_state_dict() must not use it).
        nets = [q for p in (policy1, policy2) for q in (p, getattr(p, "meta", None), getattr(p, "ctrl", None)) if hasattr(q, "pdl")]
        saved_pdl = [q.pdl for q in nets]
        for q in nets:
            q.pdl = True
        self._p1, self._p2, self._after = policy1, policy2, after_step
        n, dev = env.num_envs, env.device
        self._a1 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._a2 = torch.zeros(n, dtype=torch.uint8, device=dev) if policy2 is not None else None
        self._prev = torch.empty_like(env.obs_buf[0]) if after_step is not None and not fused else None
        if fused and warmup_steps % env.out_slots:
            warmup_steps += env.out_slots - warmup_steps % env.out_slots     # the slot ring is back at its start for the capture
        self.last = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up outside capture (lazy initialisation, autotuning)
            for _ in range(warmup_steps):
                self._one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            for _ in range(self.k_steps):
                self.last = self._one_step()
        for q, v in zip(nets, saved_pdl):
            q.pdl = v

    def _one_step(self):
        env = self.env
        if self.fused:
            obs = env.obs_buf[env._slot]
            if self._a2 is not None:
                self._a2.copy_(self._p2(obs))
            want = self._a1 if self._after is not None else None
            if hasattr(self._p1, "ctrl"):                       # HDQNPolicy: goal launch + fused controller/env launch
                out = self._p1.step(env, a2=self._a2, explore=self.explore, actions_out=want)
            else:
                out = env.policy_step(self._p1, a2=self._a2, explore=self.explore, actions_out=want)
            if self._after is not None:
                self._after(obs, self._a1, self._a2, out)
            return out
        obs = env.obs_buf[0]
        if self._prev is not None:
            self._prev.copy_(obs)
        self._a1.copy_(self._p1(obs))
        if self._a2 is not None:
            self._a2.copy_(self._p2(obs))
        out = env.step(self._a1, self._a2)
        if self._after is not None:
            self._after(self._prev, self._a1, self._a2, out)
        return out

    def run(self, times: int = 1):
        """Advance `times * k_steps` steps; returns the (obs, rew, done, info) of the last step (device tensors that
        alias the env's output buffers).  No host synchronisation."""
        for _ in range(times):
            self.graph.replay()
        return self.last
