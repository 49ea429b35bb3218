"""world_size-2 gloo tests (CPU) of the multi-GPU protocol: sharding by global env id, gradient SUM
all-reduce with 1/global_rows scaling and all-reduced advantage statistics reproduce the single-process
global-minibatch update exactly (up to float summation order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import envs_oracle as eo, ppo_oracle as po


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_batch(seed=0, B=512, d=6, A=5):
    rng = np.random.default_rng(seed)
    obs = rng.normal(size=(B, d)).astype(np.float32)
    act = rng.integers(0, A, B)
    adv = (rng.normal(size=B) * 2 + 0.5).astype(np.float32)
    ret = rng.normal(size=B).astype(np.float32)
    params = po.init_params(d, A, 1) + rng.normal(scale=0.02, size=po.init_params(d, A, 1).shape).astype(np.float32)
    flat = torch.from_numpy(params)
    with torch.no_grad():
        logits, _ = po.forward(flat, torch.from_numpy(obs), d, A)
        old_logp, _ = po.categorical(logits, torch.from_numpy(act))
    old_logp = old_logp.numpy() + rng.normal(scale=0.2, size=B).astype(np.float32)
    return params, obs, act, adv, old_logp, ret


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from three_mlagents_b200 import distributed as D

    D.init_from_env(backend="gloo")
    d, A = 6, 5
    params, obs, act, adv, old_logp, ret = _make_batch()
    B = len(adv)
    lo, hi = rank * B // world, (rank + 1) * B // world
    sl = slice(lo, hi)
    # (sum, sumsq, count) all-reduce -> global normalisation statistics
    sums = torch.tensor([adv[sl].astype(np.float64).sum(), (adv[sl].astype(np.float64) ** 2).sum(), hi - lo], dtype=torch.float64)
    D.allreduce_sum_(sums)
    mean, std = D.adv_mean_std(sums)
    flat = torch.from_numpy(params.copy()).requires_grad_(True)
    logits, values = po.forward(flat, torch.from_numpy(obs[sl]), d, A)
    a_n = (torch.from_numpy(adv[sl]) - np.float32(mean)) / (np.float32(std) + 1e-8)
    # local loss terms summed (not averaged), scaled by 1/global_rows: the SUM over ranks is the global mean
    logp, ent = po.categorical(logits, torch.from_numpy(act[sl]))
    ratio = torch.exp(logp - torch.from_numpy(old_logp[sl]))
    pg = -torch.min(a_n * ratio, a_n * torch.clamp(ratio, 0.8, 1.2)).sum()
    vl = ((torch.from_numpy(ret[sl]) - values) ** 2).sum()
    loss = (pg + 0.01 * (-ent.sum()) + 0.5 * vl) / D.global_rows(hi - lo)
    loss.backward()
    g = flat.grad.detach().clone()
    D.allreduce_sum_(g)
    assert D.env_shard(rank, world, 100) == (rank * 100, rank * 100 + 100)
    if rank == 0:
        out["grad"] = g.numpy()
        out["mean_std"] = (mean, std)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_global_minibatch():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        grad2, (mean, std) = out["grad"], out["mean_std"]
    params, obs, act, adv, old_logp, ret = _make_batch()
    assert abs(mean - adv.astype(np.float64).mean()) < 1e-9 and abs(std - adv.astype(np.float64).std(ddof=1)) < 1e-9
    flat = torch.from_numpy(params.copy()).requires_grad_(True)
    logits, values = po.forward(flat, torch.from_numpy(obs), 6, 5)
    loss, _ = po.ppo_loss(logits, values, torch.from_numpy(act), torch.from_numpy(adv), torch.from_numpy(old_logp),
                          torch.from_numpy(ret))
    loss.backward()
    want = flat.grad.numpy()
    assert np.abs(grad2 - want).max() <= 1e-6 * max(1.0, np.abs(want).max())


def test_sharded_envs_equal_the_unsharded_batch():
    """Oracle twin of the sharding rule: ranks owning [r*N,(r+1)*N) reproduce rows of one 2N-env batch."""
    from three_mlagents_b200.distributed import env_shard

    N, seed = 64, 9
    full = eo.OracleVecEnv("gridworld", 2 * N, seed=seed)
    parts = [eo.OracleVecEnv("gridworld", N, seed=seed, env_id_base=env_shard(r, 2, N)[0]) for r in range(2)]
    for t in range(150):
        a = eo.random_actions("gridworld", seed, full.env_ids, t)
        o, r, d, _, _ = full.step(a)
        for k, p in enumerate(parts):
            op, rp, dp, _, _ = p.step(a[k * N:(k + 1) * N])
            assert np.array_equal(op, o[k * N:(k + 1) * N]) and np.array_equal(rp, r[k * N:(k + 1) * N])
            assert np.array_equal(dp, d[k * N:(k + 1) * N])
    with pytest.raises(ValueError):
        env_shard(2, 2, 10)


def test_numa_binding_is_optional_and_never_widens_the_affinity():
    """bench.py pins each rank to the CPUs local to its GPU before allocating pinned host buffers; without NVML (this
    container) the helper must report None and leave the process alone."""
    import os

    from three_mlagents_b200.distributed import bind_to_gpu_numa_node

    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    assert cpus is None or (set(cpus) <= before and set(cpus) == after)
    assert after <= before
    os.sched_setaffinity(0, before)
