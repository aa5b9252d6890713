"""merging_gym_b200 — B200-native batched implementation of merging-gym's env hot path.

The package holds only what the hot path needs:
  csrc/        fused CUDA kernels for sm_100a + the C ABI (include/merging_b200.h)
  _native.py   ctypes binding of that ABI (no fallback: missing library => error)
  vec_env.py   `MergeVecEnv`, the gym-0.20-style vector env over device tensors
  scalar_env.py `MergeEnv` / `make("merging_env-v0")`, the reference's scalar interface
  sharding.py  one-shard-per-rank helpers and the NCCL statistics all-reduce
  spaces.py    Discrete(5) / Box(10) stand-ins (gym is not a dependency)
  replay.py    device-resident transition ring (replay rows) and per-episode CSV logs
  policy.py    policy-in-the-loop: the reference's Q-networks as one fused forward+argmax kernel
  graphed.py   K policy+env(+recorder) steps captured in one CUDA graph
  tracing.py   optional NVTX ranges around the launches
"""
from ._native import NativeError  # noqa: F401
from .graphed import GraphedPolicyRollout  # noqa: F401
from .policy import Exploration, HDQNPolicy, MLPPolicy, explore, goal_status  # noqa: F401
from .replay import CsvEpisodeLogger, OptionRecorder, TransitionRecorder  # noqa: F401
from .scalar_env import ENV_ID, MergeEnv, make, make_vec, register_gym  # noqa: F401
from .sharding import AsyncStatsReducer, all_reduce_stats, init_distributed, shard_range  # noqa: F401
from .spaces import Box, Discrete, MultiDiscrete  # noqa: F401
from .tracing import enable as enable_nvtx, nvtx_range  # noqa: F401
from .vec_env import MergeVecEnv, StepInfo  # noqa: F401

__version__ = "0.1.0"
