"""TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy + torch-CPU autograd) of the Stable-Baselines3
2.9.0 PPO arithmetic the reference runs (backend/mlagents/training.py:150,166 with hyper-parameters
training.py:361-391).  The product path never imports this module.

PARITY UNPINNED: SB3 is a third-party dependency pinned in backend/uv.lock:1686-1688
(stable-baselines3==2.9.0 on torch==2.12.1) and is neither vendored in /root/reference nor
installed in this image (no network), and the reference's own tests assert nothing numerical about
PPO (tests/test_mlagents.py:74-101 only checks `action is not None`).  This file restates SB3's
published algorithm and is anchored on the reference's call sites; its
self-checks (tests/test_oracle_cpu.py) are GAE vs the O(T^2) definition, lambda=1 -> Monte-Carlo
returns, ratio==1 -> clip inactive, analytic loss gradients vs torch autograd.

Sections: A.1 policy, A.2 rollout bootstrap, A.3 GAE, A.4 minibatch order, A.5 update.
"""
from __future__ import annotations

import math

import numpy as np
import torch

H = 256  # net_arch dict(pi=[256,256], vf=[256,256]), training.py:363-365


# ---- A.1 policy ---------------------------------------------------------------------------------
def param_shapes(obs_dim: int, n_actions: int):
    """policy.parameters() order of SB3's ActorCriticPolicy (also the flat layout of include/tmla.h)."""
    return [
        ("mlp_extractor.policy_net.0.weight", (H, obs_dim)), ("mlp_extractor.policy_net.0.bias", (H,)),
        ("mlp_extractor.policy_net.2.weight", (H, H)), ("mlp_extractor.policy_net.2.bias", (H,)),
        ("mlp_extractor.value_net.0.weight", (H, obs_dim)), ("mlp_extractor.value_net.0.bias", (H,)),
        ("mlp_extractor.value_net.2.weight", (H, H)), ("mlp_extractor.value_net.2.bias", (H,)),
        ("action_net.weight", (n_actions, H)), ("action_net.bias", (n_actions,)),
        ("value_net.weight", (1, H)), ("value_net.bias", (1,)),
    ]


def init_params(obs_dim: int, n_actions: int, seed: int) -> np.ndarray:
    """Orthogonal init with SB3's gains: sqrt(2) towers, 0.01 action head, 1.0 value head; zero biases."""
    g = torch.Generator().manual_seed(int(seed))
    gains = {"mlp_extractor": math.sqrt(2.0), "action_net": 0.01, "value_net": 1.0}
    chunks = []
    for name, shape in param_shapes(obs_dim, n_actions):
        t = torch.zeros(shape, dtype=torch.float32)
        if name.endswith("weight"):
            torch.nn.init.orthogonal_(t, gain=gains[name.split(".")[0]], generator=g)
        chunks.append(t.reshape(-1))
    return torch.cat(chunks).numpy()


def unflatten(flat: torch.Tensor, obs_dim: int, n_actions: int) -> dict:
    out, p = {}, 0
    for name, shape in param_shapes(obs_dim, n_actions):
        n = int(np.prod(shape))
        out[name] = flat[p:p + n].view(shape)
        p += n
    assert p == flat.numel()
    return out


def forward(flat: torch.Tensor, obs: torch.Tensor, obs_dim: int, n_actions: int):
    """ActorCriticPolicy.forward/evaluate_actions: separate tanh towers -> logits [B,A], values [B]."""
    P = unflatten(flat, obs_dim, n_actions)
    lin = torch.nn.functional.linear
    hp = torch.tanh(lin(obs, P["mlp_extractor.policy_net.0.weight"], P["mlp_extractor.policy_net.0.bias"]))
    hp = torch.tanh(lin(hp, P["mlp_extractor.policy_net.2.weight"], P["mlp_extractor.policy_net.2.bias"]))
    hv = torch.tanh(lin(obs, P["mlp_extractor.value_net.0.weight"], P["mlp_extractor.value_net.0.bias"]))
    hv = torch.tanh(lin(hv, P["mlp_extractor.value_net.2.weight"], P["mlp_extractor.value_net.2.bias"]))
    logits = lin(hp, P["action_net.weight"], P["action_net.bias"])
    values = lin(hv, P["value_net.weight"], P["value_net.bias"]).flatten()
    return logits, values


def categorical(logits: torch.Tensor, actions: torch.Tensor):
    """torch.distributions.Categorical(logits=...): log_prob(a), entropy."""
    logp_all = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    logp = logp_all.gather(1, actions.long().view(-1, 1)).flatten()
    entropy = -(logp_all.exp() * logp_all).sum(-1)
    return logp, entropy


def sample_actions(logits: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Twin of csrc/env_kernels.cu:categorical (inverse CDF on exp(l - max), float32)."""
    l = logits.astype(np.float32)
    e = np.exp(l - l.max(1, keepdims=True)).astype(np.float32)
    c = np.zeros(len(l), np.float32)
    s = np.zeros(len(l), np.float32)
    for j in range(l.shape[1]):
        s = (s + e[:, j]).astype(np.float32)
    target = (u.astype(np.float32) * s).astype(np.float32)
    a = np.full(len(l), l.shape[1] - 1, np.int32)
    found = np.zeros(len(l), bool)
    for j in range(l.shape[1]):
        c = (c + e[:, j]).astype(np.float32)
        hit = ~found & (target < c)
        a[hit] = j
        found |= hit
    return a


# ---- A.3 GAE --------------------------------------------------------------------------------------
def gae(rewards, values, dones, last_values, gamma=0.99, gae_lambda=0.95):
    """RolloutBuffer.compute_returns_and_advantage, float32 NumPy, SB3's expression verbatim.
    `dones[t]` (done after step t) is SB3's episode_starts[t+1]; the final `dones` argument of SB3 is dones[T-1]."""
    rewards = np.asarray(rewards, np.float32)
    values = np.asarray(values, np.float32)
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last_gae_lam = 0
    for step in reversed(range(T)):
        next_non_terminal = np.float32(1.0) - dones[step].astype(np.float32)
        next_values = last_values.astype(np.float32) if step == T - 1 else values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        adv[step] = last_gae_lam
    return adv, adv + values


def gae_bruteforce(rewards, values, dones, last_values, gamma, lam):
    """O(T^2) definition in float64: A_t = sum_l (gamma*lam)^l delta_{t+l}, stopped at episode ends."""
    r = np.asarray(rewards, np.float64)
    v = np.asarray(values, np.float64)
    T, n = r.shape
    nv = np.concatenate([v[1:], np.asarray(last_values, np.float64)[None]], 0)
    nnt = 1.0 - dones.astype(np.float64)
    delta = r + gamma * nv * nnt - v
    adv = np.zeros((T, n))
    for t in range(T):
        w = np.ones(n)
        for l in range(t, T):
            adv[t] += w * delta[l]
            w = w * gamma * lam * nnt[l]
    return adv


# ---- A.4 minibatch order ----------------------------------------------------------------------------
def _mix(v):
    v = (v * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(15)
    v = (v * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(13)
    return v


def permutation(seed: int, epoch: int, T: int, n: int) -> np.ndarray:
    """Twin of csrc/ppo_kernels.cu:permutation_kernel — keyed 6-round Feistel + cycle walking over
    [0, T*n), mapped from SB3's flat sample index env*T+t (swapaxes(0,1).reshape) to buffer offset t*n+env."""
    from . import philox as px

    total = T * n
    bits = 2
    while (1 << bits) < total:
        bits += 1
    bits += bits & 1
    half = bits // 2
    mask = np.uint64((1 << half) - 1)
    key = px.philox4x32_10(
        (np.array([epoch & 0xFFFFFFFF]), np.array([epoch >> 32]), np.array([0]), np.array([px.TAG_PERM])),
        (seed & 0xFFFFFFFF, seed >> 32))
    key = [np.uint64(int(k[0])) for k in key]
    x = np.arange(total, dtype=np.uint64)
    todo = np.ones(total, bool)
    while todo.any():
        xs = x[todo]
        L, R = xs >> np.uint64(half), xs & mask
        for r in range(6):                                   # rounds 4, 5 reuse keys 0, 1 with a round constant
            F = _mix(R ^ key[r & 3] ^ np.uint64(0x9E3779B9 if r >= 4 else 0)) & mask
            L, R = R, L ^ F
        xs = (L << np.uint64(half)) | R
        x[todo] = xs
        todo[todo] = xs >= total
    s = x.astype(np.int64)
    env, t = s // T, s % T
    return (t * n + env).astype(np.int32)


# ---- A.5 update ---------------------------------------------------------------------------------------
def ppo_loss(logits, values, actions, advantages, old_logp, returns, *, clip=0.2, ent_coef=0.01, vf_coef=0.5,
             normalize=True):
    """PPO.train's loss for one minibatch (torch, differentiable). Returns (loss, stats dict)."""
    adv = advantages
    if normalize and adv.numel() > 1:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    logp, entropy = categorical(logits, actions)
    ratio = torch.exp(logp - old_logp)
    pl1 = adv * ratio
    pl2 = adv * torch.clamp(ratio, 1 - clip, 1 + clip)
    pg_loss = -torch.min(pl1, pl2).mean()
    value_loss = torch.nn.functional.mse_loss(returns, values)
    ent_loss = -entropy.mean()
    loss = pg_loss + ent_coef * ent_loss + vf_coef * value_loss
    with torch.no_grad():
        lr_ = logp - old_logp
        stats = {
            "pg_loss": float(pg_loss), "value_loss": float(value_loss), "entropy_loss": float(ent_loss),
            "approx_kl": float(((torch.exp(lr_) - 1) - lr_).mean()),
            "clip_fraction": float((torch.abs(ratio - 1) > clip).float().mean()), "loss": float(loss),
        }
    return loss, stats


class OraclePPO:
    """Flat-parameter PPO learner: evaluate -> loss -> backward -> clip_grad_norm_(0.5) -> Adam(3e-4, eps 1e-5)."""

    def __init__(self, obs_dim, n_actions, seed=1, lr=3e-4, max_grad_norm=0.5, clip=0.2, ent_coef=0.01, vf_coef=0.5,
                 params=None):
        self.obs_dim, self.n_actions = obs_dim, n_actions
        p = init_params(obs_dim, n_actions, seed) if params is None else np.asarray(params, np.float32)
        self.flat = torch.nn.Parameter(torch.from_numpy(p.copy()))
        self.opt = torch.optim.Adam([self.flat], lr=lr, eps=1e-5)
        self.max_grad_norm, self.clip, self.ent_coef, self.vf_coef = max_grad_norm, clip, ent_coef, vf_coef

    def evaluate(self, obs):
        with torch.no_grad():
            return forward(self.flat, torch.as_tensor(obs, dtype=torch.float32), self.obs_dim, self.n_actions)

    def minibatch_step(self, obs, actions, advantages, old_logp, returns):
        t = lambda x, dt=torch.float32: torch.as_tensor(np.asarray(x), dtype=dt)
        logits, values = forward(self.flat, t(obs), self.obs_dim, self.n_actions)
        loss, stats = ppo_loss(logits, values, t(actions, torch.int64), t(advantages), t(old_logp), t(returns),
                               clip=self.clip, ent_coef=self.ent_coef, vf_coef=self.vf_coef)
        self.opt.zero_grad()
        loss.backward()
        grad = self.flat.grad.detach().clone().numpy()
        norm = float(torch.nn.utils.clip_grad_norm_([self.flat], self.max_grad_norm))
        self.opt.step()
        stats["grad_norm"] = norm
        return stats, grad
