"""CudaVecEnv — the SB3 `VecEnv` surface over libtmla.so.

Drop-in for what `make_vector_env` returns in the reference (backend/mlagents/training.py:71-89:
`DummyVecEnv([Monitor(adapter(env)) ...])`): same `reset()/step()/seed()/close()` contract,
NumPy in / NumPy out, `dones = terminated | truncated`, auto-reset with
`infos[i]["terminal_observation"]`, `infos[i]["TimeLimit.truncated"]` and Monitor's
`infos[i]["episode"] = {"r","l","t"}`.  All n environments advance in ONE kernel launch
(`tmla_step`); `step_tensor` is the zero-copy device path the PPO trainer uses.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import Any, Sequence

import numpy as np

from . import native
from .native import lib, check, ptr
from .spaces import spaces_for

STATE_DTYPES = {   # wire structs of include/tmla.h
    "basic": np.dtype([("pos", "<i4"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "ball3d": np.dtype([("rot", "<f8", (2,)), ("pos", "<f4", (2,)), ("vel", "<f4", (2,)),
                        ("steps", "<i4"), ("ep_return", "<f4"), ("episode", "<i4"), ("pad_", "<i4")]),
    "gridworld": np.dtype([("agent", "<i4", (2,)), ("green", "<i4", (2,)), ("red", "<i4", (2,)),
                           ("goal_type", "<i4"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "push": np.dtype([("agent", "<i4", (2,)), ("box", "<i4", (2,)), ("goal_x", "<i4"),
                      ("steps", "<i4"), ("ep_return", "<f4")]),
    "walljump": np.dtype([("agent_x", "<i4"), ("in_air", "<i4"), ("wall", "<i4"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "brickbreak": np.dtype([("pos", "<f8", (2,)), ("vel", "<f8", (2,)), ("paddle", "<f8"), ("bricks", "u1", (40,)),
                            ("steps", "<i4"), ("ep_return", "<f4")]),
}


class LazyInfos(Sequence):
    """`infos` of one vec-step, materialised per index on demand (64K dicts per step would
    dominate the step time; SB3 only reads the entries of finished episodes).  Holds a snapshot of the
    finished envs' payload only (`finished` sorted env indices -> compact rows)."""

    def __init__(self, n, done, truncated, finished, terminal_obs, ep_return, ep_length, t_elapsed):
        self._n, self._done, self._trunc = n, done, truncated
        self._idx, self._tobs, self._ret, self._len, self._t = finished, terminal_obs, ep_return, ep_length, t_elapsed

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        info: dict[str, Any] = {"TimeLimit.truncated": bool(self._trunc[i])}
        if self._done[i]:
            k = int(np.searchsorted(self._idx, i))
            info["terminal_observation"] = self._tobs[k]
            info["episode"] = {"r": round(float(self._ret[k]), 6), "l": int(self._len[k]), "t": round(self._t, 6)}
            info["steps"] = int(self._len[k])
        return info

    def finished(self):
        """Indices of envs whose episode ended on this step."""
        return self._idx if self._idx is not None else np.nonzero(self._done)[0]


class CudaVecEnv:
    def __init__(self, task_id: str, n_envs: int, seed: int = 1, *, device: int = 0, env_id_base: int = 0,
                 monitor_dir: str | os.PathLike | None = None):
        if task_id not in native.TASK_IDS:
            raise ValueError(f"Task '{task_id}' has no CUDA backend (have: {sorted(native.TASK_IDS)}).")
        self.task_id = task_id
        self.num_envs = int(n_envs)
        self.device_index = int(device)
        self.observation_space, self.action_space = spaces_for(task_id)
        self.obs_dim = self.observation_space.shape[0]
        self.n_actions = self.action_space.n
        self.max_episode_steps = lib.tmla_task_max_steps(native.TASK_IDS[task_id])
        self._h = native.vp()
        check(lib.tmla_create(native.TASK_IDS[task_id], self.num_envs, int(seed) & (2**64 - 1), int(env_id_base),
                              self.device_index, C.byref(self._h)))
        n, d = self.num_envs, self.obs_dim
        # NumPy views straight onto the handle's pinned host block (tmla_host_views): the D2H copy of a
        # step lands in these arrays, no intermediate memcpy.
        ptrs = [native.vp() for _ in range(8)]
        check(lib.tmla_host_views(self._h, *[C.byref(q) for q in ptrs]))

        def view(q, ctype, shape):
            count = int(np.prod(shape))
            return np.ctypeslib.as_array((ctype * count).from_address(q.value)).reshape(shape)

        self._pin_act = view(ptrs[0], C.c_int32, (n,))
        self._obs = view(ptrs[1], C.c_float, (n, d))
        self._rew = view(ptrs[2], C.c_float, (n,))
        self._done = view(ptrs[3], C.c_uint8, (n,))
        self._trunc = view(ptrs[4], C.c_uint8, (n,))
        self._tobs = view(ptrs[5], C.c_float, (n, d))
        self._ret = view(ptrs[6], C.c_float, (n,))
        self._len = view(ptrs[7], C.c_int32, (n,))
        recp, recw = native.vp(), native.i32(0)
        check(lib.tmla_host_records(self._h, C.byref(recp), C.byref(recw)))
        rec = view(recp, C.c_float, (n, recw.value))          # compact episode-end records {idx, ret, len, tobs[d]}
        self._rec_f, self._rec_i = rec, rec.view(np.int32)
        self._actions = None
        self._t0 = time.time()
        self._dev = None      # device-side buffers for step_tensor, allocated lazily
        self._monitor = None
        if monitor_dir is not None:
            os.makedirs(monitor_dir, exist_ok=True)
            self._monitor = open(os.path.join(monitor_dir, "0.monitor.csv"), "w", encoding="utf-8")
            self._monitor.write("#" + json.dumps({"t_start": self._t0, "env_id": task_id, "n_envs": n}) + "\nr,l,t\n")

    # ------------------------------------------------------------------ SB3 VecEnv surface (NumPy)
    def seed(self, seed: int | None = None):
        if seed is not None:
            check(lib.tmla_seed(self._h, int(seed) & (2**64 - 1)))
        return [None if seed is None else seed + i for i in range(min(self.num_envs, 1))]

    def reset(self) -> np.ndarray:
        check(lib.tmla_reset_host(self._h, ptr(self._obs)))
        return self._obs.copy()

    def step_async(self, actions) -> None:
        a = np.asarray(actions).reshape(-1)
        if a.shape[0] != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.shape[0]}")
        np.copyto(self._pin_act, a, casting="unsafe")      # int64 -> int32 straight into pinned memory
        self._actions = self._pin_act

    def step_wait(self):
        nd = native.i64(0)
        check(lib.tmla_step_pinned(self._h, C.byref(nd)))
        done = self._done.astype(bool)                      # fresh arrays: the pinned block is reused next step
        trunc = self._trunc.astype(bool)
        obs, rew = self._obs.copy(), self._rew.copy()
        if nd.value:
            k = int(nd.value)                               # compact records of the finished envs, sorted by env index
            order = np.argsort(self._rec_i[:k, 0], kind="stable")
            rec_f, rec_i = self._rec_f[:k][order], self._rec_i[:k][order]
            fin, ret, length = rec_i[:, 0].astype(np.int64), rec_f[:, 1], rec_i[:, 2]
            infos = LazyInfos(self.num_envs, done, trunc, fin, rec_f[:, 3:], ret, length, time.time() - self._t0)
            if self._monitor is not None:
                t = round(time.time() - self._t0, 6)
                for r, l in zip(ret, length):
                    self._monitor.write(f"{round(float(r), 6)},{int(l)},{t}\n")
        else:
            infos = LazyInfos(self.num_envs, done, trunc, None, None, None, None, 0.0)
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib.tmla_destroy(self._h)
            self._h = native.vp()
        if getattr(self, "_monitor", None) is not None:
            self._monitor.close()
            self._monitor = None

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def get_attr(self, name, indices=None):
        return [getattr(self, name)] * len(self._indices(indices))

    def set_attr(self, name, value, indices=None):
        raise AttributeError(f"CudaVecEnv has no per-env Python attribute '{name}' to set")

    def env_method(self, method_name, *args, indices=None, **kwargs):
        raise AttributeError(f"CudaVecEnv envs are CUDA threads; no Python method '{method_name}'")

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * len(self._indices(indices))

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    # --------------------------------------------------------------- device (torch tensor) fast path
    @property
    def handle(self):
        return self._h

    @property
    def step_count(self) -> int:
        return int(lib.tmla_step_count(self._h))

    def _device_buffers(self):
        if self._dev is None:
            import torch

            dev = torch.device("cuda", self.device_index)
            n, d = self.num_envs, self.obs_dim
            self._dev = {
                "obs": torch.empty((n, d), dtype=torch.float32, device=dev),
                "rew": torch.empty(n, dtype=torch.float32, device=dev),
                "done": torch.empty(n, dtype=torch.uint8, device=dev),
                "trunc": torch.empty(n, dtype=torch.uint8, device=dev),
                "tobs": torch.zeros((n, d), dtype=torch.float32, device=dev),
                "ret": torch.zeros(n, dtype=torch.float32, device=dev),
                "len": torch.zeros(n, dtype=torch.int32, device=dev),
            }
        return self._dev

    def reset_tensor(self):
        b = self._device_buffers()
        check(lib.tmla_reset(self._h, ptr(b["obs"]), native.current_stream()))
        return b["obs"]

    def step_tensor(self, actions, out: dict | None = None):
        """actions: int32 CUDA tensor [n].  Returns the dict of device tensors
        obs / rew / done / trunc / tobs / ret / len (overwritten by the next call)."""
        b = out or self._device_buffers()
        check(lib.tmla_step(self._h, ptr(actions), ptr(b["obs"]), ptr(b["rew"]), ptr(b["done"]), ptr(b["trunc"]),
                            ptr(b.get("tobs")), ptr(b.get("ret")), ptr(b.get("len")), native.current_stream()))
        return b

    def check_actions(self) -> None:
        check(lib.tmla_check_actions(self._h, native.current_stream()))

    def rollout_random(self, T: int, obs=None, act=None, rew=None, done=None) -> None:
        """Fused T-step random-policy rollout into [T,n,...] CUDA tensors (any may be None)."""
        check(lib.tmla_rollout_random(self._h, int(T), ptr(obs), ptr(act), ptr(rew), ptr(done), native.current_stream()))

    def get_state(self) -> np.ndarray:
        import torch

        dt = STATE_DTYPES[self.task_id]
        buf = torch.empty(self.num_envs * dt.itemsize, dtype=torch.uint8, device=torch.device("cuda", self.device_index))
        check(lib.tmla_get_state(self._h, ptr(buf), native.current_stream()))
        return buf.cpu().numpy().view(dt).copy()

    def set_state(self, state: np.ndarray) -> None:
        import torch

        dt = STATE_DTYPES[self.task_id]
        st = np.ascontiguousarray(state, dtype=dt)
        if st.shape != (self.num_envs,):
            raise ValueError(f"state must have shape ({self.num_envs},)")
        buf = torch.from_numpy(st.view(np.uint8).copy()).to(torch.device("cuda", self.device_index))
        check(lib.tmla_set_state(self._h, ptr(buf), native.current_stream()))
        torch.cuda.current_stream().synchronize()
