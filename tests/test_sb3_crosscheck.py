"""Opportunistic cross-check of the PPO arithmetic against the REAL Stable-Baselines3.

SB3 (uv.lock: stable-baselines3==2.9.0) and gymnasium are third-party dependencies of the reference, absent from
/root/reference and from this image, so these tests are skipped wherever the packages cannot be imported; on a machine
that has them they lift the "parity unpinned" caveat of oracle/ppo_oracle.py:

  * CPU half  (not gpu): one SB3 `collect_rollouts` + `PPO.train()` (reference call sites backend/mlagents/training.py:150,166,
    hyper-parameters :379-389, one full-buffer minibatch so the order of `RolloutBuffer.get` does not matter) against
    `ppo_oracle.gae` (bit-exact) and `OraclePPO.minibatch_step` (parameters within 1e-6: same torch primitives).
  * GPU half  (gpu): the same SB3 buffer through the kernels — `tmla_gae` bit-exact, the fp32 minibatch path within 2e-5 and
    the default fused bf16 path within the tolerances of tests/test_default_path_oracle_gpu.py.
"""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as po

sb3 = pytest.importorskip("stable_baselines3")
gym = pytest.importorskip("gymnasium")

OBS_DIM, N_ACTIONS, N_ENVS, N_STEPS = 6, 5, 8, 64


class _ToyEnv(gym.Env):
    """6 floats in, 5 discrete actions, random-walk dynamics with terminations and a 50-step limit handled by the caller:
    enough to produce episode boundaries inside one rollout."""

    observation_space = gym.spaces.Box(-np.inf, np.inf, (OBS_DIM,), np.float32)
    action_space = gym.spaces.Discrete(N_ACTIONS)

    def __init__(self, seed):
        self._rng = np.random.default_rng(seed)
        self._x = np.zeros(OBS_DIM, np.float32)
        self._t = 0

    def reset(self, *, seed=None, options=None):
        self._x = self._rng.standard_normal(OBS_DIM).astype(np.float32)
        self._t = 0
        return self._x.copy(), {}

    def step(self, action):
        self._x = (0.9 * self._x + 0.1 * self._rng.standard_normal(OBS_DIM) + 0.05 * (int(action) - 2)).astype(np.float32)
        self._t += 1
        terminated = bool(abs(self._x[0]) > 1.5)
        truncated = self._t >= 50 and not terminated
        return self._x.copy(), float(1.0 - abs(self._x[1])), terminated, truncated, {}


def _sb3_rollout_and_train():
    from stable_baselines3 import PPO
    from stable_baselines3.common.vec_env import DummyVecEnv

    venv = DummyVecEnv([(lambda i=i: _ToyEnv(100 + i)) for i in range(N_ENVS)])
    model = PPO("MlpPolicy", venv, seed=1, device="cpu", learning_rate=3e-4, n_steps=N_STEPS, batch_size=N_STEPS * N_ENVS,
                n_epochs=1, gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5,
                policy_kwargs=dict(net_arch=dict(pi=[256, 256], vf=[256, 256])))
    total, callback = model._setup_learn(N_STEPS * N_ENVS, None)
    p0 = torch.cat([p.detach().reshape(-1) for p in model.policy.parameters()]).numpy().copy()
    assert [tuple(p.shape) for p in model.policy.parameters()] == [s for _, s in po.param_shapes(OBS_DIM, N_ACTIONS)]
    model.collect_rollouts(model.env, callback, model.rollout_buffer, n_rollout_steps=N_STEPS)
    buf = model.rollout_buffer
    with torch.no_grad():
        last_values = model.policy.predict_values(torch.as_tensor(model._last_obs)).flatten().numpy().copy()
    roll = {
        "obs": buf.observations.copy(), "act": buf.actions.reshape(N_STEPS, N_ENVS).astype(np.int32),
        "rew": buf.rewards.copy(), "val": buf.values.copy(), "logp": buf.log_probs.copy(),
        "adv": buf.advantages.copy(), "ret": buf.returns.copy(),
        # ppo_oracle.gae's dones[t] (done after step t) = SB3's episode_starts[t+1]; the last row is `dones` of the final step
        "done": np.concatenate([buf.episode_starts[1:], model._last_episode_starts[None].astype(np.float32)], 0) > 0.5,
        "last_values": last_values,
    }
    model.train()
    p1 = torch.cat([p.detach().reshape(-1) for p in model.policy.parameters()]).numpy().copy()
    return p0, roll, p1


@pytest.fixture(scope="module")
def sb3_case():
    return _sb3_rollout_and_train()


def _flat(roll):
    total = N_STEPS * N_ENVS
    return (roll["obs"].reshape(total, OBS_DIM), roll["act"].reshape(total), roll["adv"].reshape(total),
            roll["logp"].reshape(total), roll["ret"].reshape(total))


def test_oracle_matches_sb3_gae_and_train_step(sb3_case):
    p0, roll, p1 = sb3_case
    adv, ret = po.gae(roll["rew"], roll["val"], roll["done"], roll["last_values"], 0.99, 0.95)
    assert np.array_equal(adv.view(np.uint32), roll["adv"].view(np.uint32))
    assert np.array_equal(ret.view(np.uint32), roll["ret"].view(np.uint32))
    learner = po.OraclePPO(OBS_DIM, N_ACTIONS, params=p0)
    logits, values = learner.evaluate(roll["obs"].reshape(-1, OBS_DIM))
    np.testing.assert_allclose(values.numpy(), roll["val"].reshape(-1), rtol=0, atol=1e-6)
    lp, _ = po.categorical(logits, torch.from_numpy(roll["act"].reshape(-1)))
    np.testing.assert_allclose(lp.numpy(), roll["logp"].reshape(-1), rtol=0, atol=1e-6)
    learner.minibatch_step(*_flat(roll))
    np.testing.assert_allclose(learner.flat.detach().numpy(), p1, rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_kernels_match_sb3_gae_and_train_step(sb3_case):
    from three_mlagents_b200 import ops

    p0, roll, p1 = sb3_case
    dev = torch.device("cuda")
    t = lambda x, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(dt).contiguous()
    adv, ret = ops.gae(t(roll["rew"]), t(roll["val"]), t(roll["done"].astype(np.uint8), torch.uint8), t(roll["last_values"]), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy().view(np.uint32), roll["adv"].view(np.uint32))
    assert np.array_equal(ret.cpu().numpy().view(np.uint32), roll["ret"].view(np.uint32))
    total = N_STEPS * N_ENVS
    obs_flat = t(roll["obs"].reshape(total, OBS_DIM))
    act, logp = t(roll["act"], torch.int32), t(roll["logp"])
    step = np.abs(p1 - p0)
    for impl in ("fp32", "bf16"):
        params = t(p0)
        m, v, grads = torch.zeros_like(params), torch.zeros_like(params), torch.zeros_like(params)
        if impl == "fp32":
            logits, values, cache = ops.mlp_forward(params, obs_flat, OBS_DIM, N_ACTIONS)
            dl, dv, _ = ops.ppo_loss(logits, values, act, adv, logp, ret, clip_range=0.2, ent_coef=0.01, vf_coef=0.5)
            ops.mlp_backward(params, obs_flat, OBS_DIM, N_ACTIONS, cache, dl, dv, grads=grads)
        else:
            wpack = ops.mlp_pack(params, OBS_DIM, N_ACTIONS)
            ops.ppo_minibatch(params, wpack, obs_flat, OBS_DIM, N_ACTIONS, act, adv, logp, ret, rows=total, clip_range=0.2,
                              ent_coef=0.01, vf_coef=0.5, grads=grads)
        ops.adam_clip(params, grads, m, v, 1, max_grad_norm=0.5, lr=3e-4, eps=1e-5)
        torch.cuda.synchronize()
        got = params.cpu().numpy()
        if impl == "fp32":
            np.testing.assert_allclose(got, p1, rtol=0, atol=2e-5)
        else:      # first Adam step = lr * sign(g) for |g| >> eps: compare the direction, bound the distance
            cos = float(np.dot(got - p0, p1 - p0) / (np.linalg.norm(got - p0) * np.linalg.norm(p1 - p0)))
            assert cos > 0.98 and np.abs(got - p1).max() <= 2 * 3e-4 + 1e-6 and step.max() > 1e-5
