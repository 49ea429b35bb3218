"""CudaVecEnv — the SB3 `VecEnv` surface over libtmla.so.

Drop-in for what `make_vector_env` returns in the reference (backend/mlagents/training.py:71-89:
`DummyVecEnv([Monitor(adapter(env)) ...])`): same `reset()/step()/seed()/close()` contract,
NumPy in / NumPy out, `dones = terminated | truncated`, auto-reset with
`infos[i]["terminal_observation"]`, `infos[i]["TimeLimit.truncated"]` and Monitor's
`infos[i]["episode"] = {"r","l","t"}`.  All n environments advance in ONE kernel launch
(`tmla_step`); `step_tensor` is the zero-copy device path the PPO trainer uses.

Host step (`step` / `step_wait`): the GPU writes obs / rewards / dones / episode-end records of a step straight into a pooled
pinned *result block* (`tmla_step_block`) and the arrays returned to the caller are slices of that block — no host memcpy.
A block is handed out again only when the caller has dropped every array over it (`_ResultBlocks`), so results stay valid for
as long as they are referenced, like DummyVecEnv's fresh copies; `infos` is a lazy sequence over the compact records.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
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
    "bicycle": np.dtype([("x", "<f8"), ("z", "<f8"), ("theta", "<f8"), ("phi", "<f8"), ("phi_dot", "<f8"), ("delta", "<f8"),
                         ("goal", "<f8", (2,)), ("dist", "<f8"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "glider": np.dtype([("pos", "<f8", (3,)), ("vel", "<f8", (3,)), ("rot", "<f8", (3,)), ("ang_vel", "<f8", (3,)),
                        ("waypoint", "<i4"), ("steps", "<i4"), ("ep_return", "<f4"), ("pad_", "<i4")]),
}


class LazyInfos(Sequence):
    """`infos` of one vec-step, materialised per index on demand (64K dicts per step would
    dominate the step time; SB3 only reads the entries of finished episodes).  Holds the finished envs' payload as the
    compact records the step kernel wrote ({env index, ep_return, ep_length, terminal_obs[D]}, unordered); they are
    sorted by env index on first access."""

    def __init__(self, n, done, truncated, records, t_elapsed):
        self._n, self._done, self._trunc, self._rec, self._t = n, done, truncated, records, t_elapsed
        self._idx = None

    def _sort(self):
        if self._idx is None:
            if self._rec is None or len(self._rec) == 0:
                self._idx = np.zeros(0, np.int64)
                self._tobs, self._ret, self._len = np.zeros((0, 0), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32)
            else:
                rec_i = self._rec.view(np.int32)
                order = np.argsort(rec_i[:, 0], kind="stable")
                rec = self._rec[order]                       # a copy: the pinned block can go back to the pool
                rec_i = rec.view(np.int32)
                self._idx, self._ret, self._len, self._tobs = rec_i[:, 0].astype(np.int64), rec[:, 1], rec_i[:, 2], rec[:, 3:]
            self._rec = None

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
            self._sort()
            k = int(np.searchsorted(self._idx, i))
            info["terminal_observation"] = self._tobs[k]
            info["episode"] = {"r": round(float(self._ret[k]), 6), "l": int(self._len[k]), "t": round(self._t, 6)}
            info["steps"] = int(self._len[k])
        return info

    def finished(self):
        """Indices of envs whose episode ended on this step (ascending)."""
        self._sort()
        return self._idx

    def episode_stats(self):
        """(returns, lengths) of the episodes that ended on this step, ordered like `finished()`."""
        self._sort()
        return self._ret, self._len


class _ResultBlocks:
    """Pool of pinned result blocks (include/tmla.h `tmla_result_block_*`).  A step's D2H copy lands directly in the block
    whose slices are returned to the caller as obs / rewards / dones / infos.  DummyVecEnv returns fresh copies every step
    (SB3 dummy_vec_env.py `step_wait`), so a block is handed out again only when no array over it is alive — every view of
    a block holds a reference to its `raw` array, and `sys.getrefcount(raw)` says when they are all gone.  A caller that
    keeps more than `cap` steps' results alive gets ordinary NumPy copies from then on."""

    def __init__(self, handle, n, d, cap=8):
        self._h, self._cap = handle, cap
        off = (native.i64 * 6)()
        nbytes = native.i64(0)
        check(lib.tmla_result_block_layout(handle, off, C.byref(nbytes)))
        self.off, self.nbytes = [int(x) for x in off], int(nbytes.value)
        recp, recw = native.vp(), native.i32(0)
        check(lib.tmla_host_records(handle, C.byref(recp), C.byref(recw)))
        self.rec_words = int(recw.value)              # record stride: {idx, ret, len, tobs[d]} padded to a multiple of 4 words
        self._raw, self._fixed, self._ptr = [], [], []
        self.n, self.d = n, d
        self.scratch = self._alloc()      # never handed out: the fallback copies out of it
        self._alloc(); self._alloc()      # the usual case — the caller holds one step's results while asking for the next

    def _alloc(self):
        q = native.vp()
        check(lib.tmla_result_block_alloc(self._h, C.byref(q)))
        raw = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(q.value))
        o_obs, o_rew, o_done, o_trunc = self.off[:4]
        n, d = self.n, self.d
        # the four fixed-shape slices are made once per block and handed out again with it (5 us per step otherwise)
        fixed = (raw[o_obs:o_obs + 4 * n * d].view(np.float32).reshape(n, d), raw[o_rew:o_rew + 4 * n].view(np.float32),
                 raw[o_done:o_done + n].view(np.bool_), raw[o_trunc:o_trunc + n].view(np.bool_))
        self._raw.append(raw)
        self._fixed.append(fixed)
        self._ptr.append(q)
        k = len(self._raw) - 1
        del raw, fixed
        # reference counts of an unused block, measured through the same expressions acquire() evaluates
        self._idle = (sys.getrefcount(self._raw[k]), sys.getrefcount(self._fixed[k][0]))
        return k

    def acquire(self):
        """Index of a block nobody references, or None when `cap` blocks are all still in use.  A caller can hold a block
        through one of its four cached slices (their own reference count rises) or through anything derived from them or
        from the record slice (NumPy makes `raw` the base of every derived view: its count rises)."""
        raws, fixed = self._raw, self._fixed
        idle_raw, idle_view = self._idle
        for k in range(1, len(raws)):
            if sys.getrefcount(raws[k]) == idle_raw:
                f = fixed[k]
                if (sys.getrefcount(f[0]) == idle_view and sys.getrefcount(f[1]) == idle_view
                        and sys.getrefcount(f[2]) == idle_view and sys.getrefcount(f[3]) == idle_view):
                    return k
        return self._alloc() if len(raws) <= self._cap else None

    def views(self, k, n_done):
        obs, rew, done, trunc = self._fixed[k]
        rec = None
        if n_done:
            o_rec, rw = self.off[5], self.rec_words
            rec = self._raw[k][o_rec:o_rec + 4 * rw * n_done].view(np.float32).reshape(n_done, rw)[:, :3 + self.d]
        return obs, rew, done, trunc, rec

    def close(self):
        busy = [k for k in range(len(self._raw)) if not self._is_idle(k)]
        raws, fixed, ptrs = self._raw, self._fixed, self._ptr
        self._raw, self._fixed, self._ptr = [], [], []
        for k in range(len(raws)):
            if k not in busy:
                fixed[k] = None
                lib.tmla_result_block_free(ptrs[k])
            # else: a caller still holds arrays over this block; leave the pinned memory to process teardown

    def _is_idle(self, k):
        idle_raw, idle_view = self._idle
        return sys.getrefcount(self._raw[k]) == idle_raw and all(sys.getrefcount(self._fixed[k][j]) == idle_view for j in range(4))


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
        # NumPy view of the handle's pinned action buffer (tmla_host_views); results land in pooled pinned result blocks
        q = native.vp()
        check(lib.tmla_host_views(self._h, C.byref(q), None, None, None, None, None, None, None))
        self._pin_act = np.ctypeslib.as_array((C.c_int32 * n).from_address(q.value))
        self._blocks = _ResultBlocks(self._h, n, d)
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
        obs = np.empty((self.num_envs, self.obs_dim), np.float32)
        check(lib.tmla_reset_host(self._h, ptr(obs)))
        return obs

    def step_async(self, actions) -> None:
        a = np.asarray(actions).reshape(-1)
        if a.shape[0] != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.shape[0]}")
        np.copyto(self._pin_act, a, casting="unsafe")      # int64 -> int32 straight into pinned memory
        self._actions = self._pin_act

    def step_wait(self):
        blocks = self._blocks
        k = blocks.acquire()
        nd = native.i64(0)
        check(lib.tmla_step_block(self._h, blocks._ptr[blocks.scratch if k is None else k], C.byref(nd)))
        if k is None:     # the caller holds every pooled block: ordinary copies out of the scratch block
            obs, rew, done, trunc, rec = (None if v is None else v.copy() for v in blocks.views(blocks.scratch, int(nd.value)))
        else:
            obs, rew, done, trunc, rec = blocks.views(k, int(nd.value))
        if rec is not None:
            infos = LazyInfos(self.num_envs, done, trunc, rec, time.time() - self._t0)
            if self._monitor is not None:
                t = round(time.time() - self._t0, 6)
                for r, l in zip(*infos.episode_stats()):
                    self._monitor.write(f"{round(float(r), 6)},{int(l)},{t}\n")
        else:
            infos = LazyInfos(self.num_envs, done, trunc, None, 0.0)
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._blocks.close()
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
