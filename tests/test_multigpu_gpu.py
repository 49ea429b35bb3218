"""Two-rank NCCL tests (skipped on a single-GPU box; run with `gpurun --gpus 2`): env sharding by global
env id is invisible in the trajectories, and data-parallel PPO keeps the replicas bit-identical."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, allreduce="peer"):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      TMLA_ALLREDUCE=allreduce)
    from three_mlagents_b200 import distributed as D
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    D.init_from_env("nccl", device_index=rank)
    dev = torch.device("cuda", rank)
    N, T = 512, 64
    first, _ = D.env_shard(rank, world, N)
    env = CudaVecEnv("ball3d", N, seed=3, device=rank, env_id_base=first)
    obs = torch.empty((T, N, 6), device=dev)
    act = torch.empty((T, N), dtype=torch.int32, device=dev)
    rew = torch.empty((T, N), device=dev)
    done = torch.empty((T, N), dtype=torch.uint8, device=dev)
    env.rollout_random(T, obs, act, rew, done)
    gathered = [torch.empty_like(obs) for _ in range(world)]
    dist.all_gather(gathered, obs)
    if rank == 0:
        out["obs"] = torch.cat(gathered, dim=1).cpu().numpy()
    env.close()

    env = CudaVecEnv("ball3d", N, seed=3, device=rank, env_id_base=first)
    model = CudaPPO("MlpPolicy", env, seed=3, n_steps=32, batch_size=4096, n_epochs=2, ent_coef=0.01)
    model.learn(world * N * 32)
    params = [torch.empty_like(model.params) for _ in range(world)]
    dist.all_gather(params, model.params)
    if rank == 0:
        out["params_equal"] = all(torch.equal(params[0], p) for p in params[1:])
        out["finite"] = bool(torch.isfinite(params[0]).all())
        out["timesteps"] = model.num_timesteps
        out["allreduce_impl"] = model.allreduce_impl
        out["params"] = params[0].cpu().numpy()
    dist.barrier()
    model.close()
    env.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharding_and_data_parallel_ppo():
    import torch.multiprocessing as mp
    from three_mlagents_b200.vec_env import CudaVecEnv

    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        obs2, eq, finite, ts = out["obs"], out["params_equal"], out["finite"], out["timesteps"]
    N, T = 512, 64
    env = CudaVecEnv("ball3d", world * N, seed=3)
    obs = torch.empty((T, world * N, 6), device="cuda")
    act = torch.empty((T, world * N), dtype=torch.int32, device="cuda")
    rew = torch.empty((T, world * N), device="cuda")
    done = torch.empty((T, world * N), dtype=torch.uint8, device="cuda")
    env.rollout_random(T, obs, act, rew, done)
    assert np.array_equal(obs.cpu().numpy(), obs2)          # sharding is invisible in the trajectories
    assert eq and finite and ts == world * N * 32
    env.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_allreduce_matches_nccl_and_keeps_replicas_identical():
    """The gradient all-reduce fused into clip + Adam over NVLink peer memory (csrc/comm.cu) against the NCCL path: both keep the
    two replicas bit-identical; the two paths agree to float rounding (rank-order sum vs NCCL's order, then 16 Adam steps:
    stated tolerance 2 * lr per step on a handful of sign-sensitive coordinates, cosine of the update > 0.999)."""
    import torch.multiprocessing as mp
    from oracle import ppo_oracle as po

    res = {}
    for impl in ("peer", "nccl"):
        world, port = 2, _free_port()
        with mp.Manager() as mgr:
            out = mgr.dict()
            mp.spawn(_worker, args=(world, port, out, impl), nprocs=world, join=True)
            res[impl] = dict(out)
    assert "peer-memory" in res["peer"]["allreduce_impl"] and res["nccl"]["allreduce_impl"] == "nccl"
    for impl in res:
        assert res[impl]["params_equal"] and res[impl]["finite"], impl
    p0 = po.init_params(6, 5, 3)
    a, b = res["peer"]["params"] - p0, res["nccl"]["params"] - p0
    cos = float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))
    print(f"peer vs nccl parameter update: cosine {cos:.6f}, max |dp| {np.abs(a - b).max():.2e}")
    assert cos > 0.999 and np.abs(a - b).max() <= 2 * 3e-4 * 16
