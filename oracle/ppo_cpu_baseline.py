"""TEST / BENCH INFRASTRUCTURE ONLY — the reference's CPU PPO loop restated end to end, for the timed CPU
baseline that stands beside the GPU PPO number (SURVEY.md §8(d) "(2) PPO end-to-end"): SB3's
`OnPolicyAlgorithm.learn` = collect_rollouts -> compute_returns_and_advantage -> PPO.train, reached from
backend/mlagents/training.py:150,166 with the hyper-parameters of training.py:379-389 and the registry's
n_envs (registry.py:61-64 basic = 1, :77-80 ball3d = 8).

Built from oracle/ref_port.py (scalar envs with the reference's cost structure, serial DummyVecEnv loop) and
oracle/ppo_oracle.py (torch-CPU autograd, clip_grad_norm_, Adam).  The product path never imports this file;
bench.py uses it only for `cpu_baseline.ppo` and `--impl reference`.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ppo_oracle as po
from . import ref_port

OBS_DIMS = {"basic": 21, "ball3d": 6, "gridworld": 4, "push": 4}


class CpuPPO:
    def __init__(self, task: str, n_envs: int, *, seed: int = 1, n_steps: int = 1024, batch_size: int = 256, n_epochs: int = 10,
                 gamma: float = 0.99, gae_lambda: float = 0.95, ent_coef: float = 0.01):
        self.task, self.n_envs, self.n_steps, self.batch_size, self.n_epochs = task, n_envs, n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda = gamma, gae_lambda
        torch.manual_seed(seed)
        self.env = ref_port.SerialVecEnv(task, n_envs, seed)
        self.d, self.a = OBS_DIMS[task], ref_port.TASKS[task].n_actions
        self.learner = po.OraclePPO(self.d, self.a, seed=seed, ent_coef=ent_coef)
        self.rng = np.random.default_rng(seed)
        self.last_obs = self.env.reset()
        self.num_timesteps = 0

    def collect_rollouts(self):
        T, n, d = self.n_steps, self.n_envs, self.d
        obs = np.zeros((T, n, d), np.float32)
        act = np.zeros((T, n), np.int64)
        rew, val, logp = (np.zeros((T, n), np.float32) for _ in range(3))
        done = np.zeros((T, n), bool)
        for t in range(T):
            logits, values = self.learner.evaluate(self.last_obs)
            dist = torch.distributions.Categorical(logits=logits)
            a = dist.sample()
            obs[t], act[t], val[t], logp[t] = self.last_obs, a.numpy(), values.numpy(), dist.log_prob(a).numpy()
            self.last_obs, r, dn, infos = self.env.step(act[t])
            for i, info in enumerate(infos):                       # SB3 collect_rollouts: timeout bootstrap
                if dn[i] and info.get("TimeLimit.truncated", False):
                    _, tv = self.learner.evaluate(info["terminal_observation"][None])
                    r[i] += self.gamma * float(tv[0])
            rew[t], done[t] = r, dn
        _, last_values = self.learner.evaluate(self.last_obs)
        adv, ret = po.gae(rew, val, done, last_values.numpy(), self.gamma, self.gae_lambda)
        self.num_timesteps += T * n
        return obs, act, logp, adv, ret

    def train(self, obs, act, logp, adv, ret):
        total = self.n_steps * self.n_envs
        sw = lambda x: x.swapaxes(0, 1).reshape(total, *x.shape[2:])   # RolloutBuffer.get: swap_and_flatten
        fo, fa, fl, fadv, fr = sw(obs), sw(act), sw(logp), sw(adv), sw(ret)
        for _ in range(self.n_epochs):
            perm = self.rng.permutation(total)
            for s in range(0, total, self.batch_size):
                i = perm[s:s + self.batch_size]
                self.learner.minibatch_step(fo[i], fa[i], fadv[i], fl[i], fr[i])

    def learn(self, total_timesteps: int):
        while self.num_timesteps < total_timesteps:
            self.train(*self.collect_rollouts())
        return self

    def evaluate(self, episodes: int, seed: int = 10_001):
        """evaluate_policy(deterministic=True) on a fresh single env."""
        env = ref_port.SerialVecEnv(self.task, 1, seed)
        obs = env.reset()
        returns, cur = [], 0.0
        while len(returns) < episodes:
            logits, _ = self.learner.evaluate(obs)
            obs, r, dn, _ = env.step(np.array([int(torch.argmax(logits, dim=1)[0])]))
            cur += float(r[0])
            if dn[0]:
                returns.append(cur)
                cur = 0.0
        return np.asarray(returns)


def time_ppo(task: str = "ball3d", n_envs: int = 8, n_steps: int = 1024, batch_size: int = 256, n_epochs: int = 10,
             iterations: int = 3, seed: int = 1) -> dict:
    """Samples/s of `iterations` full PPO iterations at the reference defaults (one warm-up rollout excluded)."""
    m = CpuPPO(task, n_envs, seed=seed, n_steps=n_steps, batch_size=batch_size, n_epochs=n_epochs)
    t_roll = t_train = 0.0
    for _ in range(iterations):
        t0 = time.perf_counter()
        batch = m.collect_rollouts()
        t1 = time.perf_counter()
        m.train(*batch)
        t2 = time.perf_counter()
        t_roll += t1 - t0
        t_train += t2 - t1
    samples = iterations * n_envs * n_steps
    return {"value": samples / (t_roll + t_train), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "rollout_s": t_roll / iterations, "update_s": t_train / iterations,
            "sample": f"{task}: {iterations} iterations of n_envs={n_envs}, n_steps={n_steps}, batch {batch_size}, {n_epochs} epochs "
                      f"(SB3 PPO restated on torch-CPU autograd + scalar Python envs; torch threads = {torch.get_num_threads()})"}


def train_config1(total_timesteps: int = 25_000, seed: int = 1, eval_episodes: int = 50) -> dict:
    """BASELINE configs[0]: `three-mlagents train basic -a ppo -t 25000 --seed 1` on the CPU restatement."""
    t0 = time.perf_counter()
    m = CpuPPO("basic", 1, seed=seed).learn(total_timesteps)
    wall = time.perf_counter() - t0
    r = m.evaluate(eval_episodes, seed=seed + 10_000)
    return {"wall_s": wall, "timesteps": m.num_timesteps, "eval_mean_reward": float(r.mean()), "eval_std_reward": float(r.std()),
            "eval_episodes": eval_episodes, "cores": torch.get_num_threads()}
