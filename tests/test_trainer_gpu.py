"""End-to-end parity of one PPO iteration: CudaPPO (all kernels, device-resident) against the CPU
restatement of SB3 (oracle/ppo_oracle.py) fed with the SAME rollout and the SAME minibatch order.

Tolerances: values/log-probs stored by the rollout 2e-5 abs; timeout bootstrap + GAE 1e-4 abs
(they inherit the value error, gamma-discounted over the horizon); parameters after
n_epochs x n_minibatches Adam steps 2e-5 abs (lr 3e-4 per step bounds the drift)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import envs_oracle as eo, ppo_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("task,n,T,B", [("ball3d", 64, 32, 256), ("gridworld", 96, 120, 1000), ("basic", 50, 60, 512)])
def test_one_iteration_matches_sb3_restatement(task, n, T, B):
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    env = CudaVecEnv(task, n, seed=3)
    model = CudaPPO("MlpPolicy", env, seed=3, n_steps=T, batch_size=B, n_epochs=2, ent_coef=0.01, mlp_impl="fp32")
    d, a = env.obs_dim, env.n_actions
    p0 = model.params.cpu().numpy().copy()
    np.testing.assert_array_equal(p0, po.init_params(d, a, 3))          # same orthogonal init
    ora_env = eo.OracleVecEnv(task, n, seed=3)
    model.collect_rollouts()
    torch.cuda.synchronize()
    obs, act = model.obs.cpu().numpy(), model.act.cpu().numpy()
    rew, done = model.rew.cpu().numpy(), model.done.cpu().numpy().astype(bool)
    val, logp = model.val.cpu().numpy(), model.logp.cpu().numpy()
    learner = po.OraclePPO(d, a, params=p0)
    # --- rollout replay on the oracle env with the device's sampled actions ------------------------
    tol = 1e-5 if task == "ball3d" else 0.0
    cur = eo.observe(task, ora_env.state)
    want_rew = np.zeros((T, n), np.float32)
    n_boot = 0
    for t in range(T):
        assert np.abs(obs[t] - cur).max() <= tol, t
        logits, values = learner.evaluate(obs[t])
        np.testing.assert_allclose(val[t], values.numpy(), rtol=0, atol=2e-5)
        lp, _ = po.categorical(logits, torch.from_numpy(act[t]))
        np.testing.assert_allclose(logp[t], lp.numpy(), rtol=0, atol=2e-5)
        cur, r, dn, tl, info = ora_env.step(act[t])
        assert np.array_equal(done[t], dn)
        if tl.any():                                                      # collect_rollouts timeout bootstrap
            _, tv = learner.evaluate(info["terminal_obs"][tl])
            r = r.copy()
            r[tl] = r[tl] + np.float32(0.99) * tv.numpy()
            n_boot += int(tl.sum())
        want_rew[t] = r
    np.testing.assert_allclose(rew, want_rew, rtol=0, atol=1e-4)
    assert int(model.trunc_count.item()) == n_boot
    if task != "ball3d":
        assert n_boot > 0
    _, last_v = learner.evaluate(obs[T])
    np.testing.assert_allclose(model.last_values.cpu().numpy(), last_v.numpy(), rtol=0, atol=2e-5)
    # --- GAE on the device's own buffers is bit-exact; against oracle-valued buffers within 1e-4 -----
    adv_dev, ret_dev = model.adv.cpu().numpy(), model.ret.cpu().numpy()
    w_adv, w_ret = po.gae(rew, val, done, model.last_values.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(adv_dev.view(np.uint32), w_adv.view(np.uint32))
    # --- update: same minibatch order, compare parameters ------------------------------------------------
    n_mb = model.train()
    torch.cuda.synchronize()
    total = T * n
    fo = obs[:T].reshape(total, d)
    fl = lambda x: x.reshape(total)
    steps = 0
    for epoch in range(2):
        perm = po.permutation(3, epoch, T, n)
        for s in range(0, total, B):
            idx = perm[s:s + B]
            stats, _ = learner.minibatch_step(fo[idx], fl(act)[idx], fl(adv_dev)[idx], fl(logp)[idx], fl(ret_dev)[idx])
            steps += 1
    assert steps == n_mb
    got, want = model.params.cpu().numpy(), learner.flat.detach().numpy()
    assert np.abs(want - p0).max() > 1e-4                      # the update moved the parameters
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    last = model.stats.cpu().numpy()
    assert abs(last[5] - stats["loss"]) <= 1e-4 * max(1.0, abs(stats["loss"]))
    env.close()


def test_learn_improves_basic_and_roundtrips(tmp_path):
    """BASELINE config 1 in miniature: PPO on `basic` reaches the large goal; save/load/predict/evaluate work."""
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    env = CudaVecEnv("basic", 256, seed=1)
    model = CudaPPO("MlpPolicy", env, seed=1, n_steps=64, batch_size=2048, n_epochs=10, ent_coef=0.01)
    r0, _ = model.evaluate(50, seed=10_001)
    model.learn(256 * 64 * 12)
    r1, l1 = model.evaluate(50, seed=10_001)
    assert r1.mean() > 0.85 > r0.mean()                        # registry.py:64 reward_threshold for basic
    assert (l1 >= 7).all()
    path = tmp_path / "basic_policy_test.zip"
    model.save(path)
    loaded = CudaPPO.load(path)
    assert torch.equal(loaded.params, model.params)
    obs = np.zeros(21, np.float32); obs[10] = 1.0
    a, _ = loaded.predict(obs, deterministic=True)
    assert int(a) == 2                                          # move right, towards the large goal at 17
    acts, _ = loaded.predict(np.stack([obs, obs]), deterministic=True)
    assert acts.shape == (2,)
    rows = model.logger_rows
    assert rows[-1]["rollout/ep_rew_mean"] > rows[0]["rollout/ep_rew_mean"]
    env.close(); loaded.env.close()


@pytest.mark.parametrize("task,n,T,B", [("ball3d", 512, 32, 4096), ("gridworld", 300, 40, 1000)])
def test_bf16_tensor_core_path_tracks_fp32(task, n, T, B):
    """Same seeds, same rollout protocol: the tcgen05 bf16 path must stay close to the fp32 path.
    Stated tolerance: logits/values 3e-2 abs (bf16 activations, 8-bit mantissa, over two 256-wide layers),
    gradient cosine similarity > 0.999 on one minibatch."""
    from three_mlagents_b200 import ops
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    env = CudaVecEnv(task, n, seed=5)
    m = CudaPPO("MlpPolicy", env, seed=5, n_steps=T, batch_size=B, n_epochs=1, ent_coef=0.01, mlp_impl="bf16")
    d, a = env.obs_dim, env.n_actions
    with torch.no_grad():
        m.params += 0.05 * torch.randn_like(m.params)          # non-trivial biases / heads
    m._repack()
    m.collect_rollouts()
    torch.cuda.synchronize()
    obs_flat = m.obs[:T].reshape(T * n, d)
    rows = min(B, T * n)
    idx = ops.permutation(1, 0, T, n)[:rows].contiguous()
    l16, v16, c16 = ops.mlp_forward(m.params, obs_flat, d, a, index=idx, wpack=m.wpack)
    l32, v32, c32 = ops.mlp_forward(m.params, obs_flat, d, a, index=idx)
    assert float((l16 - l32).abs().max()) < 3e-2 and float((v16 - v32).abs().max()) < 3e-2
    dl, dv, _ = ops.ppo_loss(l32, v32, m.act, m.adv, m.logp, m.ret, index=idx)
    g16 = ops.mlp_backward(m.params, obs_flat, d, a, c16, dl, dv, index=idx, wpack=m.wpack)
    g32 = ops.mlp_backward(m.params, obs_flat, d, a, c32, dl, dv, index=idx)
    cos = float(torch.nn.functional.cosine_similarity(g16, g32, dim=0))
    rel = float((g16 - g32).norm() / g32.norm())
    print(f"{task}: bf16 vs fp32 gradient cosine {cos:.6f}, relative error {rel:.4f}")
    assert cos > 0.999 and rel < 0.05
    m.train()
    torch.cuda.synchronize()
    assert torch.isfinite(m.params).all()
    env.close()
