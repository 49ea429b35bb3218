"""Multi-GPU protocol of the hot path: one process per GPU, environments sharded by global
env id, ONE data-path collective — a sum all-reduce of the flat fp32 gradient per minibatch — plus one tiny
all-reduce per epoch of the advantage statistics of all its minibatches ([n_minibatches, 3] doubles) so that
normalisation and the loss mean keep global-minibatch semantics.  The reference is single-process (no torch.distributed anywhere); this module is
new surface, kept tiny so the same functions run under gloo on CPU (tests) and NCCL on GPUs.
"""
from __future__ import annotations

import os


def dist_state():
    """(dist module or None, rank, world_size) of the initialised default process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def init_from_env(backend: str = "nccl", device_index: int | None = None):
    """Initialise the default process group from torchrun's environment (RANK/WORLD_SIZE/MASTER_*)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return dist_state()
    kwargs = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else device_index
        torch.cuda.set_device(local)
        kwargs["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, **kwargs)
    return dist_state()


def bind_to_gpu_numa_node(device_index: int) -> list[int] | None:
    """Pin the calling thread to the CPUs NVML reports as local to GPU `device_index` (its NUMA node).  Call it BEFORE the
    first `CudaVecEnv` of the process: the pinned result blocks of the host step are then allocated on the GPU's own node, so
    the kernel's PCIe stores do not cross the socket interconnect (matters with one process per GPU on a two-socket box).
    Returns the CPU list, or None when NVML is unavailable (the binding is an optimisation, never a requirement)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 - NVML missing, containers without the sysfs topology, ...
        return None


def env_shard(rank: int, world: int, envs_per_rank: int) -> tuple[int, int]:
    """[first, last) global env ids owned by `rank`; the Philox sub-sequence of an env is its global id,
    so trajectories do not depend on `world`."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    return rank * envs_per_rank, (rank + 1) * envs_per_rank


def allreduce_sum_(tensor):
    """In-place sum over ranks (no-op for a single process). Gradients are produced already scaled by
    1/global_rows, so the SUM is the exact global-minibatch gradient."""
    dist, _, world = dist_state()
    if world > 1:
        dist.all_reduce(tensor)
    return tensor


def global_rows(local_rows: int) -> int:
    """Rows of the global minibatch when every rank contributes `local_rows` (equal shards)."""
    return local_rows * dist_state()[2]


def adv_mean_std(sums):
    """(mean, unbiased std) from the all-reduced (sum, sum of squares, count) triple — what
    csrc/ppo_kernels.cu:ppo_loss_kernel computes on the device."""
    s, ss, n = (float(x) for x in sums)
    mean = s / n
    var = max((ss - s * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5


class PeerComm:
    """Gradient exchange over NVLink peer memory (include/tmla.h `tmla_comm_*`, csrc/comm.cu): every rank's exchange slot is
    IPC-mapped into every other rank, and the all-reduce runs inside the clip + Adam launches (`tmla_adam_clip_allreduce`).
    `create` is collective: it returns None on EVERY rank when any rank could not set the mapping up (ranks on different
    hosts, no peer access, a single process) — callers then stay on NCCL."""

    def __init__(self, handle, rank: int, world: int):
        self.handle, self.rank, self.world = handle, rank, world

    @classmethod
    def create(cls, device_index: int, num_floats: int):
        import ctypes as C

        import torch

        from . import native

        dist, rank, world = dist_state()
        if world == 1 or os.environ.get("TMLA_ALLREDUCE", "peer").lower() == "nccl":
            return None
        h = native.vp()
        mine = (C.c_uint8 * 64)()
        ok, err = 1, ""
        try:
            native.check(native.lib.tmla_comm_create(rank, world, int(device_index), int(num_floats), C.byref(h), mine))
        except Exception as e:  # noqa: BLE001
            ok, err = 0, str(e)
        dev = torch.device("cuda", int(device_index))
        handles = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(handles, torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8).to(dev))
        if ok:
            blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in handles)
            try:
                native.check(native.lib.tmla_comm_connect(h, blob))
            except Exception as e:  # noqa: BLE001
                ok, err = 0, str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if h:
                native.lib.tmla_comm_destroy(h)
            if err and rank == 0:
                import warnings

                warnings.warn(f"peer-memory gradient exchange unavailable ({err}); using NCCL all-reduce", stacklevel=2)
            return None
        dist.barrier()
        return cls(h, rank, world)

    def check(self) -> None:
        from . import native

        native.check(native.lib.tmla_comm_check(self.handle, native.current_stream()))

    def close(self) -> None:
        from . import native

        if self.handle:
            native.lib.tmla_comm_destroy(self.handle)
            self.handle = None
