"""WebSocket helper layer (three-mlagents_b200/websocket_training.py) against the payload contract of
backend/mlagents/websocket_training.py:36-51, 82-112, 191-193 — CPU part: callback throttling and frames with
a socket double and a stand-in model (no CUDA call is made)."""
import asyncio
import threading

from three_mlagents_b200 import websocket_training as wst


class FakeSocket:
    application_state = "CONNECTED"

    def __init__(self):
        self.frames = []

    async def send_json(self, payload):
        self.frames.append(payload)


class CudaPPO:                                   # the frame carries model.__class__.__name__
    num_timesteps = 0


def test_progress_callback_throttles_and_reports_fraction():
    async def main():
        sock, model = FakeSocket(), CudaPPO()
        cb = wst.WebSocketProgressCallback(sock, asyncio.get_running_loop(), total_timesteps=10_000, progress_freq=2_000)

        def worker():                            # learn() runs in a worker thread, as under asyncio.to_thread
            for ts in (1024, 2048, 3072, 4096, 8192, 12288):
                model.num_timesteps = ts
                assert cb.on_rollout(model) is True

        t = threading.Thread(target=worker)
        t.start()
        while t.is_alive():
            await asyncio.sleep(0.01)
        t.join()
        await asyncio.sleep(0.05)
        return sock.frames

    frames = asyncio.run(main())
    # emitted when >= 2000 new timesteps accumulated: 2048, 4096, 8192, 12288
    assert [f["timesteps"] for f in frames] == [2048, 4096, 8192, 12288]
    assert all(f["type"] == "progress" and f["reward"] is None and f["loss"] is None for f in frames)
    assert frames[0]["progress"] == 2048 / 10_000 and frames[-1]["progress"] == 1.0
    assert frames[0]["algorithm"] == "CudaPPO" and frames[0]["episode"] == 2048
    assert set(frames[0]) == {"type", "episode", "reward", "loss", "timesteps", "progress", "algorithm"}


def test_send_error_swallows_socket_failures():
    class Broken:
        async def send_json(self, payload):
            raise RuntimeError("closed")

    sock = FakeSocket()
    asyncio.run(wst.send_error(sock, ValueError("bad task")))
    assert sock.frames == [{"type": "error", "message": "bad task"}]
    asyncio.run(wst.send_error(Broken(), ValueError("x")))          # must not raise


def test_connected_accepts_starlette_state_and_doubles():
    class S:
        pass

    s = S()
    s.application_state = "CONNECTED"
    assert wst._connected(s)
    s.application_state = "DISCONNECTED"
    assert not wst._connected(s)
    s.application_state = True
    assert wst._connected(s)
    try:
        from starlette.websockets import WebSocketState
    except ImportError:
        return
    s.application_state = WebSocketState.CONNECTED
    assert wst._connected(s)
    s.application_state = WebSocketState.DISCONNECTED
    assert not wst._connected(s)
