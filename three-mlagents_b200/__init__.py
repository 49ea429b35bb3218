"""three-mlagents_b200 — B200-native (sm_100a) data-parallel hot path of three-mlagents.

Same surface as the reference's `mlagents` package (backend/mlagents/__init__.py:8-10):
`TaskSpec, get_task, list_tasks, make_env`; plus `training` (TrainConfig / train_task / ...),
`cli`, `vec_env.CudaVecEnv` and the `native` ctypes binding of libtmla.so.
"""
from .registry import TaskSpec, get_task, list_tasks, make_env

__all__ = ["TaskSpec", "get_task", "list_tasks", "make_env"]
__version__ = "0.1.0"
