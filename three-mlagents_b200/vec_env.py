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
import time
import weakref
from typing import Any, Sequence

import numpy as np

from . import native
from .native import lib, check, ptr
from .spaces import spaces_for

_F32, _BOOL = np.dtype(np.float32), np.dtype(np.bool_)

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

    def __init__(self, n, done, truncated, records, t_elapsed, step_no=0, last_reset=None):
        self._n, self._done, self._trunc, self._rec, self._t = n, done, truncated, records, t_elapsed
        self._idx = None
        # `steps` of a running env (the reference's `_info`, envs.py:154-159, reports it every step) = vec-steps since its last
        # reset; `last_reset` is the env's live bookkeeping array, so the value is exact until the NEXT step() call
        self._step_no, self._last_reset = step_no, last_reset

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
        elif self._last_reset is not None:
            info["steps"] = max(0, int(self._step_no - self._last_reset[i]))
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
    """Pool of pinned result blocks (include/tmla.h `tmla_result_block_*`).  A step's results land directly in a block whose
    slices are returned to the caller as obs / rewards / dones / infos.  DummyVecEnv returns fresh copies every step (SB3
    dummy_vec_env.py `step_wait`), so a block is handed out again only when no array over it is alive.  Every hand-out is a
    LEASE: a fresh base ndarray over the block's memory that all returned arrays (and anything sliced from them) keep alive
    through `.base`; a `weakref` callback on that base returns the block to the free list when the last of them is dropped.
    No reference counts are inspected, so a debugger or tracer holding extra references only delays the reuse.  A caller
    that keeps more than `cap` steps' results alive gets ordinary NumPy copies from then on."""

    def __init__(self, handle, n, d, cap=8):
        self._h, self._cap = handle, cap
        off = (native.i64 * 6)()
        nbytes = native.i64(0)
        check(lib.tmla_result_block_layout(handle, off, C.byref(nbytes)))
        self.off, self.nbytes = [int(x) for x in off], int(nbytes.value)
        recp, recw = native.vp(), native.i32(0)
        check(lib.tmla_host_records(handle, C.byref(recp), C.byref(recw)))
        self.rec_words = int(recw.value)              # record stride: {idx, ret, len, tobs[d]} padded to a multiple of 4 words
        self._mem, self._ptr, self._lease, self._release = [], [], [], []
        self._free: list[int] = []
        self.n, self.d = n, d
        self.scratch = self._alloc()      # never handed out: the fallback copies out of it
        self._free.clear()
        self._alloc(); self._alloc()      # the usual case — the caller holds one step's results while asking for the next

    def _alloc(self):
        q = native.vp()
        check(lib.tmla_result_block_alloc(self._h, C.byref(q)))
        k = len(self._mem)
        self._mem.append((C.c_uint8 * self.nbytes).from_address(q.value))
        self._ptr.append(q)
        self._lease.append(None)
        free = self._free
        self._release.append(lambda _ref, k=k: free.append(k))
        free.append(k)
        return k

    def acquire(self):
        """Index of a block no live array refers to, or None when `cap` blocks are all still in use."""
        if not self._free:
            if len(self._mem) > self._cap:
                return None
            self._alloc()
        return self._free.pop()

    def release(self, k):
        """Give back a block acquired but never handed out (the step failed)."""
        self._free.append(k)

    def views(self, k, n_done, lease=True):
        """(obs, rew, done, trunc, records) over block k.  With `lease`, the arrays share one fresh base whose death
        (the caller dropped them all) puts the block back on the free list."""
        raw = np.frombuffer(self._mem[k], dtype=np.uint8)
        if lease:
            self._lease[k] = weakref.ref(raw, self._release[k])
        n, d = self.n, self.d
        o_obs, o_rew, o_done, o_trunc = self.off[:4]
        obs = np.ndarray((n, d), _F32, raw, o_obs)
        rew = np.ndarray((n,), _F32, raw, o_rew)
        done = np.ndarray((n,), _BOOL, raw, o_done)
        trunc = np.ndarray((n,), _BOOL, raw, o_trunc)
        rec = None
        if n_done:
            rec = np.ndarray((n_done, self.rec_words), _F32, raw, self.off[5])[:, :3 + self.d]
        return obs, rew, done, trunc, rec

    def close(self):
        free = set(self._free) | {self.scratch}
        mem, ptrs = self._mem, self._ptr
        self._mem, self._ptr, self._lease, self._free = [], [], [], []
        for k in range(len(mem)):
            if k in free:
                lib.tmla_result_block_free(ptrs[k])
            # else: a caller still holds arrays over this block; leave the pinned memory to process teardown


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
        self._blocks = _ResultBlocks(self._h, n, d)      # results land in pooled pinned result blocks
        self._actions = None
        self._pending = None                             # result block of the step in flight (None: the scratch block)
        self._in_flight = False
        self._t0 = time.time()
        self._vec_step = 0                               # host-path vec-steps taken (for infos[i]["steps"])
        self._last_reset = np.zeros(n, np.int64)         # vec-step index at which env i last (auto-)reset
        self._ep_log = None                              # device episode log of the policy-driven path (Monitor rows)
        self._dev = None      # device-side buffers for step_tensor, allocated lazily
        self._monitor = None
        if monitor_dir is not None:
            os.makedirs(monitor_dir, exist_ok=True)
            self._monitor = open(os.path.join(monitor_dir, "0.monitor.csv"), "w", encoding="utf-8")
            self._monitor.write("#" + json.dumps({"t_start": self._t0, "env_id": task_id, "n_envs": n}) + "\nr,l,t\n")

    # ------------------------------------------------------------------ SB3 VecEnv surface (NumPy)
    def seed(self, seed: int | None = None):
        """SB3 `VecEnv.seed`: env i gets seed + i; returns the num_envs seeds (a random base seed when None).  Here the base
        seed keys the Philox streams and the env's global id is the sub-sequence, which is the same separation."""
        if seed is None:
            seed = int(np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32))
        check(lib.tmla_seed(self._h, int(seed) & (2**64 - 1)))
        return [int(seed) + i for i in range(self.num_envs)]

    def reset(self) -> np.ndarray:
        obs = np.empty((self.num_envs, self.obs_dim), np.float32)
        check(lib.tmla_reset_host(self._h, ptr(obs)))
        self._last_reset[:] = self._vec_step
        return obs

    def step_async(self, actions) -> None:
        a = np.asarray(actions).reshape(-1)
        if a.shape[0] != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.shape[0]}")
        if a.dtype not in (np.int32, np.int64) or not a.flags.c_contiguous:
            a = np.ascontiguousarray(a, dtype=np.int64)
        # chunk by chunk: range check + narrowing to one byte per action in the pinned stage, then that chunk's step kernel —
        # the staging of chunk c runs while chunk c-1's results already travel over PCIe.  An out-of-range action raises
        # IndexError with every env in its pre-step state (like the reference's ACTION_DELTAS[action]).
        blocks = self._blocks
        k = blocks.acquire()
        try:
            check(lib.tmla_step_block_begin(self._h, a.ctypes.data, a.dtype.itemsize, blocks._ptr[blocks.scratch if k is None else k]))
        except BaseException:
            if k is not None:
                blocks.release(k)
            raise
        self._actions, self._pending, self._in_flight = a, k, True

    def step_wait(self):
        if not self._in_flight:
            raise RuntimeError("step_wait() without a step_async() in flight")
        self._in_flight = False
        blocks = self._blocks
        k = self._pending
        nd = native.i64(0)
        try:
            check(lib.tmla_step_block_end(self._h, blocks._ptr[blocks.scratch if k is None else k], C.byref(nd)))
        except BaseException:
            if k is not None:
                blocks.release(k)
            raise
        n_done = int(nd.value)
        if k is None:     # the caller holds every pooled block: ordinary copies out of the scratch block
            obs, rew, done, trunc, rec = (None if v is None else v.copy() for v in blocks.views(blocks.scratch, n_done, lease=False))
        else:
            obs, rew, done, trunc, rec = blocks.views(k, n_done)
        self._vec_step += 1
        if rec is not None:
            infos = LazyInfos(self.num_envs, done, trunc, rec, time.time() - self._t0, self._vec_step, self._last_reset)
            self._last_reset[rec[:, 0].view(np.int32)] = self._vec_step
            if self._monitor is not None:
                t = round(time.time() - self._t0, 6)
                for r, l in zip(*infos.episode_stats()):
                    self._monitor.write(f"{round(float(r), 6)},{int(l)},{t}\n")
        else:
            infos = LazyInfos(self.num_envs, done, trunc, None, 0.0, self._vec_step, self._last_reset)
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._ep_log = None
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

    def _stream(self) -> int:
        """torch's current stream ON THIS ENV'S DEVICE (not the thread's current device)."""
        import torch

        return torch.cuda.current_stream(self.device_index).cuda_stream

    def reset_tensor(self):
        b = self._device_buffers()
        check(lib.tmla_reset(self._h, ptr(b["obs"]), self._stream()))
        return b["obs"]

    def attach_episode_log(self, capacity: int) -> None:
        """Monitor for the policy-driven device path (`tmla_step_policy` / `tmla_rollout`): finished episodes append
        (return, length) records to a device buffer that `flush_episode_log` turns into monitor.csv rows."""
        import torch

        dev = torch.device("cuda", self.device_index)
        cap = int(max(1, min(capacity, 1 << 22)))
        self._ep_log = {"rec": torch.zeros((cap, 2), dtype=torch.float32, device=dev),
                        "count": torch.zeros(1, dtype=torch.int32, device=dev), "cap": cap, "dropped": 0}
        check(lib.tmla_set_episode_log(self._h, ptr(self._ep_log["rec"]), cap, ptr(self._ep_log["count"])))

    def flush_episode_log(self):
        """(returns, lengths) of the episodes logged since the last flush (order within a rollout is arbitrary); writes
        them to monitor.csv when this env has a monitor file.  Synchronises the env's stream."""
        log = self._ep_log
        if log is None:
            return np.zeros(0, np.float32), np.zeros(0, np.int64)
        count = int(log["count"].item())
        k = min(count, log["cap"])
        log["dropped"] += count - k
        rec = log["rec"][:k].cpu().numpy()
        log["count"].zero_()
        rets, lens = rec[:, 0].copy(), rec[:, 1].astype(np.int64)
        if self._monitor is not None and k:
            t = round(time.time() - self._t0, 6)
            self._monitor.write("".join(f"{round(float(r), 6)},{int(l)},{t}\n" for r, l in zip(rets, lens)))
            self._monitor.flush()
        return rets, lens

    def step_tensor(self, actions, out: dict | None = None):
        """actions: int32 CUDA tensor [n].  Returns the dict of device tensors
        obs / rew / done / trunc / tobs / ret / len (overwritten by the next call)."""
        b = out or self._device_buffers()
        check(lib.tmla_step(self._h, ptr(actions), ptr(b["obs"]), ptr(b["rew"]), ptr(b["done"]), ptr(b["trunc"]),
                            ptr(b.get("tobs")), ptr(b.get("ret")), ptr(b.get("len")), self._stream()))
        return b

    def check_actions(self) -> None:
        check(lib.tmla_check_actions(self._h, self._stream()))

    def rollout_random(self, T: int, obs=None, act=None, rew=None, done=None) -> None:
        """Fused T-step random-policy rollout into [T,n,...] CUDA tensors (any may be None)."""
        check(lib.tmla_rollout_random(self._h, int(T), ptr(obs), ptr(act), ptr(rew), ptr(done), self._stream()))

    def get_state(self) -> np.ndarray:
        import torch

        dt = STATE_DTYPES[self.task_id]
        buf = torch.empty(self.num_envs * dt.itemsize, dtype=torch.uint8, device=torch.device("cuda", self.device_index))
        check(lib.tmla_get_state(self._h, ptr(buf), self._stream()))
        return buf.cpu().numpy().view(dt).copy()

    def set_state(self, state: np.ndarray) -> None:
        import torch

        dt = STATE_DTYPES[self.task_id]
        st = np.ascontiguousarray(state, dtype=dt)
        if st.shape != (self.num_envs,):
            raise ValueError(f"state must have shape ({self.num_envs},)")
        buf = torch.from_numpy(st.view(np.uint8).copy()).to(torch.device("cuda", self.device_index))
        check(lib.tmla_set_state(self._h, ptr(buf), self._stream()))
        torch.cuda.current_stream(self.device_index).synchronize()
        self._last_reset[:] = self._vec_step - np.asarray(st["steps"], np.int64)
