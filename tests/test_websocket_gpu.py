"""train_task_for_websocket / run_policy_for_websocket on the CUDA backend: the frames a browser client of the
reference receives (backend/mlagents/websocket_training.py:82-112, 141-188)."""
import asyncio

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class FakeSocket:
    application_state = "CONNECTED"

    def __init__(self):
        self.frames = []

    async def send_json(self, payload):
        self.frames.append(payload)


def test_train_and_run_policy_over_websocket(tmp_path, monkeypatch):
    from three_mlagents_b200 import websocket_training as wst
    from three_mlagents_b200.registry import make_env

    monkeypatch.chdir(tmp_path)
    sock = FakeSocket()
    result = asyncio.run(wst.train_task_for_websocket(sock, "gridworld", total_timesteps=16 * 1024 * 3, algorithm="ppo",
                                                      n_envs=16, eval_episodes=8, progress_freq=16 * 1024))
    kinds = [f["type"] for f in sock.frames]
    assert kinds[0] == "progress" and sock.frames[0]["timesteps"] == 0 and sock.frames[0]["task_id"] == "gridworld"
    assert kinds[-1] == "trained" and kinds.count("progress") == 1 + 3       # initial frame + one per 16K timesteps
    trained = sock.frames[-1]
    assert trained["file_url"] == f"/policies/{result['model_filename']}" and trained["algorithm"] == "ppo"
    assert trained["eval_episodes"] == 8 and np.isfinite(trained["mean_reward"])
    assert trained["session_uuid"] == result["run_id"].rsplit("_", 1)[-1]

    obs = np.array([0.25, -0.5, 1.0, 0.0], np.float32)
    assert wst.predict_discrete_action("gridworld", obs, result["model_filename"]) in range(5)
    assert wst.predict_policy_action("gridworld", obs, result["model_filename"]) in range(5)

    sock2 = FakeSocket()
    asyncio.run(wst.run_policy_for_websocket(sock2, "gridworld", lambda: make_env("gridworld"),
                                             model_filename=result["model_filename"], sleep_seconds=0.0, max_steps=12))
    assert len(sock2.frames) == 12 and all(f["type"] == "run_step" and f["episode"] >= 1 for f in sock2.frames)
