#!/usr/bin/env python
"""One kernel at bench size, a few launches — the command ncu captures (`-k regex:<name>`).

    python profiles/ncu_target.py rollout|step|step_lean|mlp_tc|mlp|policy_step|policy_step_tc|record [--envs N]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "rollout"
n = int(sys.argv[sys.argv.index("--envs") + 1]) if "--envs" in sys.argv else (1 << 20 if what in ("rollout", "step", "step_lean", "record") else 1 << 18)
dev = "cuda"
if what == "rollout":
    env = mg.MergeVecEnv(n, episode_info=False)
    env.rollout(400, refresh_obs=False)
    K = 32
    o = torch.empty(K, n, 10, device=dev); r = torch.empty(K, n, 2, device=dev)
    d = torch.empty(K, n, dtype=torch.uint8, device=dev); i = torch.empty(K, n, dtype=torch.uint8, device=dev)
    for _ in range(4):
        env.rollout(K, obs=o, rew=r, done=d, info=i, refresh_obs=False)
elif what in ("step", "step_lean"):
    envs = [mg.MergeVecEnv(n, episode_info=False, env_id_base=k * n, track_returns=what == "step") for k in range(4)]
    for e in envs:
        e.rollout(400, refresh_obs=False)
    acts = [tuple(x.clone() for x in envs[0].sample_actions(t)) for t in range(4)]
    for t in range(24):
        envs[t % 4].step_async(*acts[t % 4])
elif what in ("mlp", "mlp_tc"):
    env = mg.MergeVecEnv(n, mode="pve", episode_info=False, reset_mode="random")
    env.rollout(200)
    pol = mg.MLPPolicy(10, 5, seed=1, backend="tf32x3" if what == "mlp_tc" else "fused")
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(4):
        pol.act(env.obs_buf[0], out=out)
elif what in ("policy_step", "policy_step_tc"):
    env = mg.MergeVecEnv(n, mode="pve", episode_info=False, reset_mode="random")
    env.rollout(200)
    pol = mg.MLPPolicy(10, 5, seed=1, backend="tf32x3" if what == "policy_step_tc" else "fused")
    for _ in range(4):
        env.policy_step(pol)
elif what == "record":
    env = mg.MergeVecEnv(n, out_slots=2)
    env.rollout(215)
    rec = mg.TransitionRecorder(env, 4 * n)
    obs = env.obs_buf[env._slot].clone()
    for _ in range(3):
        a1, a2 = env.sample_actions()
        out = env.step(a1, a2)
        rec.record(obs, a1, a2, out)
        obs = out[0].clone()
torch.cuda.synchronize()
print("ncu target done:", what, n)
