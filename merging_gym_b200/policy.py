"""Policy-in-the-loop: the reference's Q-networks evaluated for N envs on the
device, emitting actions that `MergeVecEnv.step` consumes with no host round-trip.

Reference pieces mirrored here
  * `Net` 10 -> 200 -> 100 -> 5 (scripts/main.py:30-47) and `Net(num_inputs, num_outputs)`
    (scripts/hdqn.py:38-55; Goal_DQN 10 -> 3, HDQN controller 11 -> 5 on `[goal] + state`).
  * greedy action `torch.max(action_value, 1)[1]` (main.py:104-105, hdqn.py:86-87, 170-171).
  * the scripts' exploration rule `np.random.randn() <= EPISILO` with EPISILO = 0.7, i.e. greedy
    with probability Phi(0.7) ~ 0.758 (main.py:103, hdqn.py:84, 168) — note it is NOT epsilon-greedy.
  * `goal_status` (hdqn.py:223-236) and the opponent view `state[5:] + state[:5]` (main.py:199).

`backend="fused"` (default) runs the hand-written CUDA kernel `mg_mlp_act` (one launch: three layers,
bias, ReLU, arg-max, exactly the reference's fp32 arithmetic on FFMA2); `backend="tf32x3"` runs
`mg_mlp_act_tc`, the same operator with the 200x100 layer on the tcgen05 tensor cores as an
error-compensated 3xTF32 product (fp32-level accuracy, not bit-identical); `backend="f16x3"`
(`MG_MLP_FLAG_F16X3`, csrc/mlp_tc16_kernels.cu) puts BOTH hidden layers on the tensor cores as three-product sums of
fp16 hi / lo operands (the same ~22 significant bits per product; activations limited to fp16's range after scaling —
the fastest policy kernel); `backend="torch"` is the
plain PyTorch fp32 reference of the same op (cuBLAS), kept for the numerics tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native as nat

EPISILO = 0.7          # scripts/main.py:16, hdqn.py:20
HIDDEN1, HIDDEN2 = 200, 100
POLICY_BACKENDS = {"fused": nat.POLICY_BACKEND_FP32, "tf32x3": nat.POLICY_BACKEND_TF32X3, "f16x3": nat.POLICY_BACKEND_F16X3}   # mg_policy_step
TC_BACKENDS = ("tf32x3", "f16x3")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class MLPPolicy:
    """`Net`: Linear(in,200) - ReLU - Linear(200,100) - ReLU - Linear(100,out), batched arg-max."""

    def __init__(self, in_dim: int = 10, out_dim: int = 5, device="cuda", state_dict: Optional[dict] = None,
                 seed: Optional[int] = None, backend: str = "fused"):
        if backend not in ("fused", "tf32x3", "f16x3", "torch"):
            raise ValueError("backend must be 'fused', 'tf32x3', 'f16x3' or 'torch'")
        self.in_dim, self.out_dim, self.backend = int(in_dim), int(out_dim), backend
        self.device = torch.device(device)
        self.pdl = False                     # see act(): set True when no kernel that writes the weights precedes act()
        if state_dict is None:
            # the reference's init: weights uniform_(0,1), biases nn.Linear default (main.py:34-39)
            g = torch.Generator().manual_seed(0 if seed is None else seed)
            def lin(i, o):
                w = torch.rand(o, i, generator=g)
                b = (torch.rand(o, generator=g) * 2 - 1) / np.sqrt(i)
                return w, b
            w1, b1 = lin(in_dim, HIDDEN1); w2, b2 = lin(HIDDEN1, HIDDEN2); w3, b3 = lin(HIDDEN2, out_dim)
            state_dict = {"fc1.weight": w1, "fc1.bias": b1, "fc2.weight": w2, "fc2.bias": b2,
                          "out.weight": w3, "out.bias": b3}
        self.load_state_dict(state_dict)

    # -- weights ---------------------------------------------------------------------------------
    def load_state_dict(self, sd: dict) -> None:
        def get(k):
            return torch.as_tensor(np.asarray(sd[k]) if not isinstance(sd[k], torch.Tensor) else sd[k]) \
                .to(device=self.device, dtype=torch.float32).contiguous()
        self.w1, self.b1 = get("fc1.weight"), get("fc1.bias")
        self.w2, self.b2 = get("fc2.weight"), get("fc2.bias")
        self.w3, self.b3 = get("out.weight"), get("out.bias")
        if tuple(self.w1.shape) != (HIDDEN1, self.in_dim) or tuple(self.w2.shape) != (HIDDEN2, HIDDEN1) \
                or tuple(self.w3.shape) != (self.out_dim, HIDDEN2):
            raise ValueError("state_dict does not match Net(%d, %d)" % (self.in_dim, self.out_dim))
        # layouts of the fused kernel, built once: W1 K-major; W2 K-major with its 100 output columns split
        # into 4 groups of 25, each zero-padded to 28 floats (16-byte aligned rows in shared memory)
        self.w1_t = self.w1.t().contiguous()
        w2p = torch.zeros(HIDDEN1, 4, 28, dtype=torch.float32, device=self.device)
        w2p[:, :, :25] = self.w2.t().reshape(HIDDEN1, 4, 25)
        self.w2_p = w2p.contiguous()
        # tensor-core backend: W2 as the UMMA B operand (N = 112 rows = 100 neurons + zero pad, K-major),
        # split into tf32 hi (top 19 bits) and lo = w - hi, each in the canonical core-matrix layout
        # [K-step 25][row group 28 = 14 hi + 14 lo][k half 2][row 8][k 4]
        bp = torch.zeros(112, HIDDEN1, dtype=torch.float32, device=self.device)
        bp[:HIDDEN2] = self.w2
        hi = (bp.view(torch.int32) & -8192).view(torch.float32)
        lo = bp - hi

        cat = torch.cat([hi, lo], dim=0)                      # 224 rows: hi 0-111, lo 112-223 (one stacked B operand)
        self.w2_tc = cat.view(28, 8, 25, 2, 4).permute(2, 0, 3, 1, 4).contiguous().view(-1)
        # f16x3 (csrc/mlp_tc16_kernels.cu): both hidden layers as fp16 hi / lo operands in ONE blob (include/merging_b200.h):
        # [c1, c2 | layer 1: N = 208 x K = 16 with fc1.bias in K slot 15 | layer 2: 13 K-steps of N = 112 x K = 16], each
        # operand times a power of two that puts its largest entry in [256, 512), hi = fp16(v), lo = fp16(v - hi)
        def scale_exp(t):
            m = float(t.abs().max())
            return 0 if m == 0.0 else 8 - int(np.floor(np.log2(m)))
        w1op = torch.zeros(208, 16, dtype=torch.float32, device=self.device)
        w1op[:HIDDEN1, :self.in_dim] = self.w1
        w1op[:HIDDEN1, 15] = self.b1
        s1, s2 = scale_exp(w1op), scale_exp(self.w2)
        w1op = torch.ldexp(w1op, torch.tensor(s1, device=self.device))
        hi1 = w1op.to(torch.float16)
        lo1 = (w1op - hi1.float()).to(torch.float16)
        op1 = torch.cat([hi1, lo1], dim=0).view(52, 8, 2, 8).permute(0, 2, 1, 3).contiguous()      # [row group][K half][row][k]
        w2op = torch.zeros(112, 208, dtype=torch.float32, device=self.device)
        w2op[:HIDDEN2, :HIDDEN1] = torch.ldexp(self.w2, torch.tensor(s2, device=self.device))
        hi2 = w2op.to(torch.float16)
        lo2 = (w2op - hi2.float()).to(torch.float16)
        op2 = torch.cat([hi2, lo2], dim=0).view(28, 8, 13, 2, 8).permute(2, 0, 3, 1, 4).contiguous()  # [K-step][row group][K half][row][k]
        hdr = torch.zeros(16, dtype=torch.float32, device=self.device)
        hdr[0], hdr[1] = 2.0 ** (-s1 - 3), 2.0 ** (3 - s2)
        self.w2_f16 = torch.cat([hdr.view(torch.uint8), op1.view(torch.uint8).view(-1), op2.view(torch.uint8).view(-1)])
        assert self.w2_f16.numel() == 64 + 13312 + 93184

    @property
    def w2_native(self) -> torch.Tensor:
        """fc2.weight in the layout the backend's kernel reads."""
        return {"tf32x3": self.w2_tc, "f16x3": self.w2_f16}.get(self.backend, self.w2_p)

    def state_dict(self) -> dict:
        return {"fc1.weight": self.w1, "fc1.bias": self.b1, "fc2.weight": self.w2, "fc2.bias": self.b2,
                "out.weight": self.w3, "out.bias": self.b3}

    @classmethod
    def load(cls, path: str, in_dim: int = 10, out_dim: int = 5, **kw):
        """Load a reference checkpoint (`eval.pth`, a plain state_dict; main.py:85-87)."""
        sd = torch.load(path, map_location="cpu", weights_only=True)
        return cls(in_dim, out_dim, state_dict=sd, **kw)

    # -- forward ---------------------------------------------------------------------------------
    def _inputs(self, obs: torch.Tensor, goal: Optional[torch.Tensor]):
        if goal is None:
            if obs.shape[-1] != self.in_dim:
                raise ValueError(f"expected obs[..., {self.in_dim}]")
            return obs
        # hdqn.py:291 `[goal] + state`
        return torch.cat([goal.to(torch.float32).unsqueeze(-1), obs], dim=-1)

    def q_values_torch(self, obs: torch.Tensor, goal: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._inputs(obs, goal)
        h = torch.relu(torch.addmm(self.b1, x, self.w1.t()))
        h = torch.relu(torch.addmm(self.b2, h, self.w2.t()))
        return torch.addmm(self.b3, h, self.w3.t())

    def act(self, obs: torch.Tensor, goal: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            q_out: Optional[torch.Tensor] = None, mirror: bool = False, pdl: Optional[bool] = None,
            obs_layout: str = "aos", n: Optional[int] = None, write_goal: bool = False) -> torch.Tensor:
        """Greedy actions uint8[N] = argmax_a Q(obs)[a] (first maximum wins, as torch.max does).
        `mirror=True` evaluates the network on the opponent's view `state[5:] + state[:5]` (main.py:199) of every
        row; the fused kernels swap the halves while they read the row.
        `pdl` (default `self.pdl`): programmatic dependent launch — the kernel stages its weights while the previous
        kernel of the stream is still draining (`MG_MLP_FLAG_PDL`).  Only valid when that kernel does not write this
        policy's weights; in a rollout loop it is the env step, so `GraphedPolicyRollout` and `bench_policy` turn it on.
        `obs_layout` ("aos" [N,10] / "soa" [10,S] with `n` given / "goal_slot" [N,11]): the layout of `obs`
        (`MergeVecEnv(obs_layout=...)`).  On "goal_slot" rows a 10-input network reads slots 1..10 and, with
        `write_goal=True`, also stores its arg-max into slot 0 (`goal = choose_goal(state)` feeding `[goal] + state`,
        hdqn.py:283,291); an 11-input network given no `goal` reads the whole row."""
        if obs_layout != "aos":
            return self._act_layout(obs, goal, out, q_out, mirror, pdl, obs_layout, n, write_goal)
        n = obs.shape[0]
        if mirror and self.backend == "torch":
            obs = torch.cat([obs[:, 5:], obs[:, :5]], dim=1)
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=self.device)
        if self.backend == "torch":
            q = self.q_values_torch(obs, goal)
            if q_out is not None:
                q_out.copy_(q)
            out.copy_(q.argmax(dim=1))
            return out
        lib = nat.load()
        if not (obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()):
            raise ValueError("obs must be a contiguous float32 CUDA tensor")
        obs_dim = obs.shape[1]
        if obs_dim + (0 if goal is None else 1) != self.in_dim:
            raise ValueError("obs/goal widths do not match the network input")
        if goal is not None and not (goal.dtype == torch.uint8 and goal.is_contiguous()):
            raise ValueError("goal must be a contiguous uint8 tensor")
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        flags = (nat.MLP_FLAG_MIRROR if mirror else 0) | (nat.MLP_FLAG_PDL if (self.pdl if pdl is None else pdl) else 0)
        with torch.cuda.device(self.device):
            if self.backend in TC_BACKENDS:
                flags |= nat.MLP_FLAG_F16X3 if self.backend == "f16x3" else 0
                nat.check(lib.mg_mlp_act_tc(_ptr(obs), _ptr(goal), n, obs_dim, self.out_dim,
                                            _ptr(self.w1_t), _ptr(self.b1), _ptr(self.w2_native), _ptr(self.b2),
                                            _ptr(self.w3), _ptr(self.b3), _ptr(out), _ptr(q_out), flags, stream),
                          "mg_mlp_act_tc")
            else:
                nat.check(lib.mg_mlp_act(_ptr(obs), _ptr(goal), n, obs_dim, self.out_dim,
                                         _ptr(self.w1_t), _ptr(self.b1), _ptr(self.w2_p), _ptr(self.b2),
                                         _ptr(self.w3), _ptr(self.b3), _ptr(out), _ptr(q_out), flags, stream),
                          "mg_mlp_act")
        return out

    __call__ = act

    def _act_layout(self, obs, goal, out, q_out, mirror, pdl, obs_layout, n, write_goal):
        if self.backend == "torch":
            raise ValueError("backend 'torch' reads the default [N,10] rows only")
        if obs_layout not in nat.OBS_LAYOUTS:
            raise ValueError(f"obs_layout must be one of {nat.OBS_LAYOUTS}")
        if not (obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()):
            raise ValueError("obs must be a contiguous float32 CUDA tensor")
        if obs_layout == "soa":
            if n is None or tuple(obs.shape) != (nat.OBS_DIM, nat.soa_stride(n)):
                raise ValueError("obs_layout='soa': pass n and obs of shape [10, (n + 15) & ~15]")
        else:
            n = obs.shape[0]
            if obs.shape[1] != nat.OBS_DIM + 1:
                raise ValueError("obs_layout='goal_slot': obs must be [N, 11]")
        row11 = obs_layout == "goal_slot" and goal is None and self.in_dim == nat.OBS_DIM + 1
        if self.in_dim != nat.OBS_DIM + (0 if goal is None else 1) and not row11:
            raise ValueError("obs/goal widths do not match the network input")
        if goal is not None and not (goal.dtype == torch.uint8 and goal.is_contiguous()):
            raise ValueError("goal must be a contiguous uint8 tensor")
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=self.device)
        flags = (nat.MLP_FLAG_MIRROR if mirror else 0) | (nat.MLP_FLAG_PDL if (self.pdl if pdl is None else pdl) else 0) | \
            nat.MLP_LAYOUT_FLAG[obs_layout] | (nat.MLP_FLAG_WRITE_GOAL if write_goal else 0)
        fn = nat.load().mg_mlp_act_tc if self.backend in TC_BACKENDS else nat.load().mg_mlp_act
        flags |= nat.MLP_FLAG_F16X3 if self.backend == "f16x3" else 0
        w2 = self.w2_native
        with torch.cuda.device(self.device):
            nat.check(fn(_ptr(obs), _ptr(goal), n, nat.OBS_DIM + (1 if row11 else 0), self.out_dim, _ptr(self.w1_t), _ptr(self.b1),
                         _ptr(w2), _ptr(self.b2), _ptr(self.w3), _ptr(self.b3), _ptr(out), _ptr(q_out), flags,
                         C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "mg_mlp_act (layout)")
        return out


class Exploration:
    """The scripts' exploration rule as device code (main.py:103-110, hdqn.py:84-92,168-176):
        `if np.random.randn() <= EPISILO:` greedy choice `else:` `np.random.randint(0, num_choices)`.
    `randn() <= t` holds with probability Phi(t) (0.758 for t = 0.7 — NOT epsilon-greedy with 0.7), so per env and step
    one Philox uniform is compared with Phi(t) and a second one picks the random choice; the counter is (global env
    id, step), so the draws do not depend on sharding or launch shape.  `step` advances by one per `spec()` /
    `apply()` call unless given."""

    def __init__(self, threshold: float = EPISILO, seed: int = 0, step: int = 0):
        import math
        self.threshold, self.seed, self.step = float(threshold), int(seed), int(step)
        self.keep_prob = 0.5 * (1.0 + math.erf(self.threshold / math.sqrt(2.0)))
        self.keep_u32 = min(int(self.keep_prob * 4294967296.0), 0xFFFFFFFF)

    def spec(self, step: Optional[int] = None) -> "nat.MgExplore":
        """The launch parameter block.  The device XORs `step` with each env's own clock (its meta word: reset count,
        winner, steps since reset), so `step` only has to change where envs' clocks can coincide — it does not have to
        advance inside a CUDA graph.  `self.step` is used (and NOT advanced) unless `step` is given."""
        return nat.MgExplore(self.seed, int(self.step if step is None else step), self.keep_u32, 0)

    def apply(self, choices: torch.Tensor, num_choices: int, env=None, env_id_base: int = 0, salt: int = 0,
              step: Optional[int] = None) -> torch.Tensor:
        """In place on a uint8 device tensor of greedy choices (`mg_explore`, one small kernel).  `env` (a MergeVecEnv)
        supplies the env clocks and the global env id base.  salt 0 is the stream `MergeVecEnv.policy_step(explore=...)`
        uses for actions; the h-DQN meta-controller's goals use salt 1."""
        if not (choices.is_cuda and choices.dtype == torch.uint8 and choices.is_contiguous()):
            raise ValueError("choices must be a contiguous uint8 CUDA tensor")
        ex = self.spec(step)
        meta = None
        if env is not None:
            meta, env_id_base = env.meta, env.env_id_base
        with torch.cuda.device(choices.device):
            nat.check(nat.load().mg_explore(_ptr(choices), choices.numel(), int(num_choices), C.byref(ex), _ptr(meta),
                                            int(env_id_base), int(salt),
                                            C.c_void_p(torch.cuda.current_stream(choices.device).cuda_stream)), "mg_explore")
        return choices


def explore(greedy: torch.Tensor, num_actions: int, generator: Optional[torch.Generator] = None,
            threshold: float = EPISILO, exploration: Optional[Exploration] = None, env=None) -> torch.Tensor:
    """The scripts' exploration rule, batched: keep the greedy action where randn() <= 0.7, else a uniform random
    action (main.py:103-110).  With `exploration` (an `Exploration`) the draw is the hand-written Philox kernel
    (`mg_explore`, in place on a uint8 tensor); without it, the plain PyTorch restatement of the same rule on torch's
    generator, kept as the reference the kernel's distribution is tested against."""
    if exploration is not None:
        return exploration.apply(greedy, num_actions, env)
    n, dev = greedy.shape[0], greedy.device
    keep = torch.randn(n, device=dev, generator=generator) <= threshold
    rnd = torch.randint(0, num_actions, (n,), device=dev, generator=generator, dtype=torch.int64).to(greedy.dtype)
    return torch.where(keep, greedy, rnd)


def goal_status(obs: torch.Tensor) -> torch.Tensor:
    """hdqn.py:223-236 for a batch: 0 if dx1 < -0.5*v2, 1 if dx1 < 0.5*v2, else 2  (uint8[N])."""
    dx1, v2 = obs[:, 0], obs[:, 9]
    return torch.where(dx1 < -0.5 * v2, 0, torch.where(dx1 < 0.5 * v2, 1, 2)).to(torch.uint8)


class HDQNPolicy:
    """Two-level h-DQN acting greedily: the meta-controller `Net(10,3)` picks a goal and the
    controller `Net(11,5)` acts on `[goal] + state` (hdqn.py:58-139, 142-221).  In the reference's
    loop the goal is re-chosen from the new state after every env step (hdqn.py:303), so acting
    greedily is `ctrl([meta(state)] + state)` at every step."""

    def __init__(self, device="cuda", meta_state: Optional[dict] = None, ctrl_state: Optional[dict] = None,
                 seed: int = 0, backend: str = "fused"):
        self.meta = MLPPolicy(10, 3, device, meta_state, seed, backend)
        self.ctrl = MLPPolicy(11, 5, device, ctrl_state, seed + 1, backend)
        self.goal: Optional[torch.Tensor] = None

    def act(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None, obs_layout: str = "aos",
            n: Optional[int] = None) -> torch.Tensor:
        n = obs.shape[0] if obs_layout != "soa" else n
        if self.goal is None or self.goal.shape[0] != n:
            self.goal = torch.empty(n, dtype=torch.uint8, device=obs.device)
        if obs_layout == "goal_slot":
            # `[goal] + state` rows: the goal network drops its choice into slot 0, the controller reads the row as it is
            self.meta.act(obs, out=self.goal, obs_layout=obs_layout, write_goal=True)
            return self.ctrl.act(obs, out=out, obs_layout=obs_layout)
        self.meta.act(obs, out=self.goal, obs_layout=obs_layout, n=n)          # hdqn.py:283,303  choose_goal
        return self.ctrl.act(obs, goal=self.goal, out=out, obs_layout=obs_layout, n=n)   # hdqn.py:291-292  [goal] + state

    __call__ = act

    def step(self, env, a2: Optional[torch.Tensor] = None, explore: Optional[Exploration] = None,
             actions_out: Optional[torch.Tensor] = None):
        """One iteration of the h-DQN loop (hdqn.py:288-303) in two launches: `choose_goal(state)` (meta forward +
        arg-max [+ exploration, salt 1]) and `mg_policy_step` with the controller on `[goal] + state` — forward,
        arg-max, exploration and `env.step` fused.  Returns the step tuple; `self.goal` holds the goals acted on."""
        obs, n = env.obs_buf[env._slot], env.num_envs
        if self.goal is None or self.goal.shape[0] != n:
            self.goal = torch.empty(n, dtype=torch.uint8, device=obs.device)
        if env.obs_layout == "goal_slot" and explore is None:
            self.meta.act(obs, out=self.goal, obs_layout="goal_slot", write_goal=True)
            return env.policy_step(self.ctrl, a2=a2, actions_out=actions_out)          # the controller reads slot 0
        self.meta.act(obs, out=self.goal, obs_layout=env.obs_layout, n=n)
        if explore is not None:
            explore.apply(self.goal, self.meta.out_dim, env, salt=1)
        return env.policy_step(self.ctrl, goal=self.goal, a2=a2, explore=explore, actions_out=actions_out)
