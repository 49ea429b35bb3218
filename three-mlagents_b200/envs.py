"""Single-environment Gymnasium-style view of the CUDA backend.

The reference's `make_env(task)` returns one `gym.Env` (BasicMoveToGoalEnv or
LegacySingleAgentGymAdapter, backend/mlagents/envs.py:30-159) used by `cli inspect`, the eval env
(training.py:92-95) and the contract tests (tests/test_mlagents.py:32-72).  `CudaTaskEnv` gives the
same 5-tuple API on top of a one-env `CudaVecEnv`; it is a convenience surface, not the hot path —
throughput comes from `CudaVecEnv` with thousands of envs per launch.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .spaces import spaces_for


class CudaTaskEnv:
    metadata = {"render_modes": []}

    def __init__(self, task_id: str, *, device: int = 0):
        self.task_id = task_id
        self.device = device
        self.observation_space, self.action_space = spaces_for(task_id)
        self._vec = None
        self._seed = 1
        self._episodes = 0

    def _make(self, seed: int):
        from .vec_env import CudaVecEnv

        if self._vec is not None:
            self._vec.close()
        self._vec = CudaVecEnv(self.task_id, 1, seed=seed, device=self.device)

    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None):
        if seed is not None or self._vec is None:
            self._seed = self._seed if seed is None else int(seed)
            self._make(self._seed)
        obs = self._vec.reset()[0]
        if self.task_id == "basic" and options and "position" in options:   # envs.py:55-56
            st = self._vec.get_state()
            st["pos"] = int(np.clip(int(options["position"]), 0, 20))
            self._vec.set_state(st)
            obs = np.zeros(21, np.float32)
            obs[st["pos"][0]] = 1.0
        return obs, self._info()

    def step(self, action):
        a = int(action)
        if not 0 <= a < self.action_space.n:
            raise IndexError(f"action {a} outside Discrete({self.action_space.n})")
        pre = self._vec.get_state() if self.task_id == "basic" else None
        obs, rew, done, infos = self._vec.step(np.array([a], np.int32))
        info_i = infos[0]
        truncated = bool(info_i["TimeLimit.truncated"])
        terminated = bool(done[0]) and not truncated
        if done[0]:
            # single-env Gymnasium semantics: return the terminal observation, caller resets
            out_obs = info_i["terminal_observation"]
            info = {"steps": info_i["steps"]}
            if self.task_id == "basic":
                info["position"] = int(np.argmax(out_obs))
        else:
            out_obs = obs[0]
            info = self._info()
        return out_obs, float(rew[0]), terminated, truncated, info

    def _info(self) -> dict[str, Any]:
        st = self._vec.get_state()
        info: dict[str, Any] = {"steps": int(st["steps"][0])}
        if self.task_id == "basic":                       # envs.py:83-84
            info = {"position": int(st["pos"][0]), "steps": int(st["steps"][0])}
        return info

    def close(self) -> None:
        if self._vec is not None:
            self._vec.close()
            self._vec = None


def make_basic_env() -> CudaTaskEnv:        # envs.py:162-163
    return CudaTaskEnv("basic")


def make_ball3d_env() -> CudaTaskEnv:       # envs.py:166-175
    return CudaTaskEnv("ball3d")


def make_gridworld_env() -> CudaTaskEnv:    # envs.py:178-187
    return CudaTaskEnv("gridworld")


def make_push_env() -> CudaTaskEnv:         # envs.py:190-199
    return CudaTaskEnv("push")


def make_walljump_env() -> CudaTaskEnv:     # envs.py:202-213
    return CudaTaskEnv("walljump")


def make_brick_break_env() -> CudaTaskEnv:  # envs.py:216-227
    return CudaTaskEnv("brickbreak")


def make_bicycle_env() -> CudaTaskEnv:      # envs.py:230-241
    return CudaTaskEnv("bicycle")


def make_glider_env() -> CudaTaskEnv:       # envs.py:244-255
    return CudaTaskEnv("glider")
