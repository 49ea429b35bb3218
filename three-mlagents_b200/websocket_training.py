"""WebSocket helpers for training and serving policies from the CUDA backend — same surface as
backend/mlagents/websocket_training.py:19-193 (`WebSocketProgressCallback`, `train_task_for_websocket`,
`predict_discrete_action`, `predict_policy_action`, `run_policy_for_websocket`, `send_error`) and the same JSON
payloads, so the reference's FastAPI routes (main.py:149-168, 311-356) and browser client keep working.

Differences, all forced by the device-resident rollout:
  * the progress callback fires per rollout (`CudaPPO.learn` calls `callback.on_rollout(model)`), not per env
    step: `progress_freq` is honoured at that granularity (SURVEY.md §7 "per-step host hooks");
  * the socket is duck-typed (`send_json`, `application_state`): fastapi/starlette are imported lazily and only
    to read `WebSocketState.CONNECTED`, so this module also works with a test double.
`train_task` owns its CUDA context work on whatever thread runs it; `asyncio.to_thread` is used exactly as in the
reference (websocket_training.py:98) and every call creates and closes its own `CudaVecEnv`.
"""
from __future__ import annotations

import asyncio
import contextlib
from dataclasses import asdict
from typing import Any, Callable

import numpy as np

from .registry import get_task


class WebSocketProgressCallback:
    """Send coarse learning progress to an already accepted socket (websocket_training.py:19-51)."""

    def __init__(self, websocket, loop: asyncio.AbstractEventLoop, *, total_timesteps: int, progress_freq: int = 2_000):
        self.websocket = websocket
        self.loop = loop
        self.total_timesteps = max(1, total_timesteps)
        self.progress_freq = max(1, progress_freq)
        self._last_emit = 0
        self.num_timesteps = 0
        self.model = None

    def payload(self) -> dict[str, Any]:
        return {
            "type": "progress",
            "episode": int(self.num_timesteps),
            "reward": None,
            "loss": None,
            "timesteps": int(self.num_timesteps),
            "progress": min(1.0, self.num_timesteps / self.total_timesteps),
            "algorithm": self.model.__class__.__name__,
        }

    def on_rollout(self, model) -> bool:
        """Called by CudaPPO.learn after every rollout + update (the reference's `_on_step`, per rollout)."""
        self.model = model
        self.num_timesteps = int(model.num_timesteps)
        if self.num_timesteps - self._last_emit < self.progress_freq:
            return True
        self._last_emit = self.num_timesteps
        asyncio.run_coroutine_threadsafe(self.websocket.send_json(self.payload()), self.loop)
        return True


async def train_task_for_websocket(websocket, task_id: str, *, total_timesteps: int | None = None, algorithm: str | None = None,
                                   seed: int = 1, n_envs: int | None = None, eval_episodes: int | None = None,
                                   eval_freq: int = 10_000, progress_freq: int = 2_000) -> dict[str, Any]:
    """websocket_training.py:54-113: initial progress frame, training in a worker thread, final `trained` frame."""
    from .training import TrainConfig, train_task

    task = get_task(task_id)
    config = TrainConfig(task_id=task_id, total_timesteps=total_timesteps, algorithm=algorithm, seed=seed, n_envs=n_envs,
                         eval_episodes=eval_episodes, eval_freq=eval_freq, verbose=0)
    effective_timesteps = total_timesteps or task.total_timesteps
    loop = asyncio.get_running_loop()
    callback = WebSocketProgressCallback(websocket, loop, total_timesteps=effective_timesteps, progress_freq=progress_freq)
    await websocket.send_json({
        "type": "progress", "episode": 0, "reward": None, "loss": None, "timesteps": 0, "progress": 0.0,
        "algorithm": algorithm or "default", "task_id": task.id,
    })
    result = await asyncio.to_thread(train_task, config, callback=callback)
    await websocket.send_json({
        "type": "trained",
        "file_url": f"/policies/{result.model_filename}",
        "model_filename": result.model_filename,
        "timestamp": result.run_id,
        "session_uuid": result.run_id.rsplit("_", 1)[-1],
        "algorithm": result.algorithm,
        "mean_reward": result.mean_reward,
        "std_reward": result.std_reward,
        "eval_episodes": result.eval_episodes,
        "run_dir": result.run_dir,
        "metadata_path": result.metadata_path,
    })
    return asdict(result)


def predict_discrete_action(task_id: str, obs, model_filename: str | None = None) -> int:
    """websocket_training.py:116-128."""
    from .training import predict_action

    action = predict_action(task_id, np.asarray(obs, dtype=np.float32), model_filename)
    if isinstance(action, list):
        if len(action) != 1:
            raise ValueError(f"Expected one discrete action for {task_id}, got {action}")
        return int(action[0])
    return int(action)


def predict_policy_action(task_id: str, obs, model_filename: str | None = None):
    """websocket_training.py:131-138."""
    from .training import predict_action

    return predict_action(task_id, np.asarray(obs, dtype=np.float32), model_filename)


def _connected(websocket) -> bool:
    """`websocket.application_state == WebSocketState.CONNECTED` (websocket_training.py:164); sockets that are not
    starlette objects (test doubles) may expose the state as a bool or a string."""
    state = getattr(websocket, "application_state", None)
    try:
        from starlette.websockets import WebSocketState

        if isinstance(state, WebSocketState):
            return state == WebSocketState.CONNECTED
    except ImportError:
        pass
    if isinstance(state, bool):
        return state
    return str(state).upper().endswith("CONNECTED") and not str(state).upper().endswith("DISCONNECTED")


async def run_policy_for_websocket(websocket, task_id: str, env_factory: Callable[[], Any], *, model_filename: str | None = None,
                                   action_transform: Callable[[Any], Any] | None = None, sleep_seconds: float = 0.03,
                                   max_steps: int | None = None) -> None:
    """Run a saved policy in a visualisation environment (websocket_training.py:141-188): legacy `reset/step`
    3-tuples and Gymnasium 5-tuples are both accepted, `get_state_for_viz` is streamed when present.  The policy
    is loaded ONCE (the reference reloads the zip on every step); `max_steps` bounds the loop for tests."""
    from .training import load_model

    task = get_task(task_id)
    model = load_model(task, model_filename)
    env = env_factory()
    episode = 0
    reset_result = env.reset()
    obs = reset_result[0] if isinstance(reset_result, tuple) else reset_result
    transform = action_transform or (lambda action: action)
    steps = 0
    try:
        while _connected(websocket) and (max_steps is None or steps < max_steps):
            action, _ = model.predict(np.asarray(obs, dtype=np.float32), deterministic=True)
            action = int(action.item()) if isinstance(action, np.ndarray) and action.ndim == 0 else int(action)
            result = env.step(transform(action))
            if len(result) == 5:
                next_obs, _, terminated, truncated, _ = result
                done = bool(terminated or truncated)
            else:
                next_obs, _, done = result
            state_for_viz = getattr(env, "get_state_for_viz", None)
            payload: dict[str, Any] = {"type": "run_step", "episode": episode + 1}
            if callable(state_for_viz):
                payload["state"] = state_for_viz()
            await websocket.send_json(payload)
            await asyncio.sleep(sleep_seconds)
            if done:
                episode += 1
                reset_result = env.reset()
                obs = reset_result[0] if isinstance(reset_result, tuple) else reset_result
            else:
                obs = next_obs
            steps += 1
            await asyncio.sleep(0)
    finally:
        model.env.close()


async def send_error(websocket, exc: Exception) -> None:
    """websocket_training.py:191-193."""
    with contextlib.suppress(Exception):
        await websocket.send_json({"type": "error", "message": str(exc)})
