"""The path that bench.py times — CudaPPO with its DEFAULTS (mlp_impl="auto" -> bf16 tcgen05 towers, fused
forward+loss+backward kernel, fused clip+Adam) — compared DIRECTLY with the CPU restatement of SB3's PPO
(oracle/ppo_oracle.py, torch-CPU fp32/fp64 autograd), not with another CUDA path of this repo.

Reference call sites: backend/mlagents/training.py:150 (`PPO("MlpPolicy", ...)`), :166 (`model.learn`), hyper-parameters
:379-389.  The oracle is a restatement of SB3 2.9.0 (PARITY UNPINNED, see its header); what this file pins is that the
benchmarked kernels compute what the restatement computes.

Stated tolerances (bf16 operands with 8-bit mantissas through two 256-wide layers, fp32 accumulation):
  * values / log-probs stored by the rollout, logits / values of a minibatch ............ 3e-2 abs
  * bootstrapped rewards (gamma * V(terminal_obs)) ..................................... 3e-2 abs
  * GAE on the device's own buffers ..................................................... bit-exact
  * minibatch gradient, whole vector ..................... cosine > 0.999, relative L2 error < 5e-2
  * minibatch gradient, per parameter tensor ........ ||err_t|| <= 5e-2 ||g_t|| + 2e-3 ||g|| (a slice whose
    gradient nearly cancels carries only rounding noise, hence the absolute term)
  * loss statistics (pg / value / entropy / total) ....................... 2e-2 * max(1, |ref|)
  * parameters after n_epochs x n_minibatches Adam steps: cosine of the update (p - p0) > 0.98 and
    |p_dev - p_oracle| <= 2 * lr * steps (Adam moves a coordinate by at most lr per step; elements whose gradient is
    below the bf16 noise floor may take the opposite sign)
"""
import numpy as np
import pytest
import torch

from oracle import envs_oracle as eo, ppo_oracle as po

pytestmark = pytest.mark.gpu


def _tensor_slices(d, a):
    out, off = [], 0
    for name, shape in po.param_shapes(d, a):
        n = int(np.prod(shape))
        out.append((name, off, off + n))
        off += n
    return out


def _oracle_grad(params, obs, act, adv, old_logp, ret, d, a, *, clip=0.2, ent_coef=0.01, vf_coef=0.5, dtype=torch.float32,
                 chunk=32768):
    """Gradient and statistics of PPO.train's loss for ONE minibatch, accumulated over row chunks (autograd on the CPU).
    Advantage normalisation and the means use whole-minibatch statistics, like SB3."""
    flat = torch.tensor(np.asarray(params), dtype=dtype, requires_grad=True)
    rows = len(obs)
    advt = torch.as_tensor(np.asarray(adv), dtype=dtype)
    advn = (advt - advt.mean()) / (advt.std() + 1e-8)
    tot = {"pg_loss": 0.0, "value_loss": 0.0, "entropy_loss": 0.0}
    logits_all, values_all = [], []
    for s in range(0, rows, chunk):
        e = min(rows, s + chunk)
        o = torch.as_tensor(np.asarray(obs[s:e]), dtype=dtype)
        logits, values = po.forward(flat, o, d, a)
        logp, entropy = po.categorical(logits, torch.as_tensor(np.asarray(act[s:e]), dtype=torch.int64))
        ratio = torch.exp(logp - torch.as_tensor(np.asarray(old_logp[s:e]), dtype=dtype))
        pg = -torch.min(advn[s:e] * ratio, advn[s:e] * torch.clamp(ratio, 1 - clip, 1 + clip)).sum() / rows
        vl = ((torch.as_tensor(np.asarray(ret[s:e]), dtype=dtype) - values) ** 2).sum() / rows
        el = -entropy.sum() / rows
        (pg + ent_coef * el + vf_coef * vl).backward()
        tot["pg_loss"] += float(pg); tot["value_loss"] += float(vl); tot["entropy_loss"] += float(el)
        logits_all.append(logits.detach()); values_all.append(values.detach())
    tot["loss"] = tot["pg_loss"] + ent_coef * tot["entropy_loss"] + vf_coef * tot["value_loss"]
    return flat.grad.detach().to(torch.float32).numpy(), tot, torch.cat(logits_all).float().numpy(), torch.cat(values_all).float().numpy()


def _check_grad(got, want, d, a, label):
    g, w = torch.from_numpy(np.asarray(got, np.float32)).double(), torch.from_numpy(np.asarray(want, np.float32)).double()
    cos = float(torch.nn.functional.cosine_similarity(g, w, dim=0))
    rel = float((g - w).norm() / w.norm())
    print(f"{label}: fused-bf16 vs oracle gradient: cosine {cos:.7f}, relative L2 error {rel:.5f}")
    assert cos > 0.999 and rel < 5e-2, (label, cos, rel)
    wn = float(w.norm())
    for name, lo, hi in _tensor_slices(d, a):
        err, gn = float((g[lo:hi] - w[lo:hi]).norm()), float(w[lo:hi].norm())
        c = float(torch.nn.functional.cosine_similarity(g[lo:hi], w[lo:hi], dim=0)) if gn > 0 else 1.0
        print(f"    {name:40s} |g| {gn:.3e}  rel err {err / max(gn, 1e-30):.2e}  cos {c:.6f}")
        assert err <= 5e-2 * gn + 2e-3 * wn, (label, name, err, gn, wn)


def _check_stats(stats_dev, want, label):
    got = {"pg_loss": float(stats_dev[0]), "value_loss": float(stats_dev[1]), "entropy_loss": float(stats_dev[2]), "loss": float(stats_dev[5])}
    for k, v in got.items():
        assert abs(v - want[k]) <= 2e-2 * max(1.0, abs(want[k])), (label, k, v, want[k])


@pytest.mark.parametrize("task,n,T,B", [("ball3d", 256, 32, 2048), ("gridworld", 300, 40, 3000)])
def test_default_iteration_matches_oracle(task, n, T, B):
    from three_mlagents_b200 import ops
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    seed, n_epochs = 4, 2
    env = CudaVecEnv(task, n, seed=seed)
    model = CudaPPO("MlpPolicy", env, seed=seed, n_steps=T, batch_size=B, n_epochs=n_epochs, ent_coef=0.01)   # all defaults
    assert model.mlp_impl == "bf16" and model.fused_update, "the default path must be the fused tensor-core one"
    d, a = env.obs_dim, env.n_actions
    p0 = model.params.cpu().numpy().copy()
    np.testing.assert_array_equal(p0, po.init_params(d, a, seed))
    ora_env = eo.OracleVecEnv(task, n, seed=seed)
    model.collect_rollouts()
    torch.cuda.synchronize()
    obs, act = model.obs.cpu().numpy(), model.act.cpu().numpy()
    rew, done = model.rew.cpu().numpy(), model.done.cpu().numpy().astype(bool)
    val, logp = model.val.cpu().numpy(), model.logp.cpu().numpy()
    learner = po.OraclePPO(d, a, params=p0)
    # ---- rollout: env transitions exact (same actions), stored values / log-probs within the bf16 tolerance
    tol_env = 1e-5 if task == "ball3d" else 0.0
    cur = eo.observe(task, ora_env.state)
    want_rew = np.zeros((T, n), np.float32)
    for t in range(T):
        assert np.abs(obs[t] - cur).max() <= tol_env, t
        logits, values = learner.evaluate(obs[t])
        np.testing.assert_allclose(val[t], values.numpy(), rtol=0, atol=3e-2)
        lp, _ = po.categorical(logits, torch.from_numpy(act[t]))
        np.testing.assert_allclose(logp[t], lp.numpy(), rtol=0, atol=3e-2)
        cur, r, dn, tl, info = ora_env.step(act[t])
        assert np.array_equal(done[t], dn)
        if tl.any():
            _, tv = learner.evaluate(info["terminal_obs"][tl])
            r = r.copy()
            r[tl] = r[tl] + np.float32(0.99) * tv.numpy()
        want_rew[t] = r
    np.testing.assert_allclose(rew, want_rew, rtol=0, atol=3e-2)
    adv_dev, ret_dev = model.adv.cpu().numpy(), model.ret.cpu().numpy()
    w_adv, _ = po.gae(rew, val, done, model.last_values.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(adv_dev.view(np.uint32), w_adv.view(np.uint32))
    # ---- one minibatch: fused kernel gradient and statistics vs CPU autograd on the same rows
    total = T * n
    fo = obs[:T].reshape(total, d)
    fl = lambda x: x.reshape(total)
    perm0 = po.permutation(seed, 0, T, n)
    np.testing.assert_array_equal(ops.permutation(seed, 0, T, n).cpu().numpy(), perm0)
    idx = perm0[:B]
    g_want, st_want, l_want, v_want = _oracle_grad(p0, fo[idx], fl(act)[idx], fl(adv_dev)[idx], fl(logp)[idx], fl(ret_dev)[idx], d, a)
    idx_dev = torch.from_numpy(idx).cuda()
    logits_dev = torch.empty((B, a), device="cuda")
    values_dev = torch.empty(B, device="cuda")
    g_dev, st_dev = ops.ppo_minibatch(model.params, model.wpack, model.obs[:T].reshape(total, d), d, a, model.act, model.adv,
                                      model.logp, model.ret, index=idx_dev, rows=B, logits=logits_dev, values=values_dev)
    torch.cuda.synchronize()
    assert float(np.abs(logits_dev.cpu().numpy() - l_want).max()) < 3e-2
    assert float(np.abs(values_dev.cpu().numpy() - v_want).max()) < 3e-2
    _check_grad(g_dev.cpu().numpy(), g_want, d, a, f"{task} B={B}")
    _check_stats(st_dev.cpu().numpy(), st_want, task)
    # ---- whole update: same minibatch order, Adam-step by Adam-step on the oracle
    n_mb = model.train()
    torch.cuda.synchronize()
    steps = 0
    for epoch in range(n_epochs):
        perm = po.permutation(seed, epoch, T, n)
        for s in range(0, total, B):
            i = perm[s:s + B]
            learner.minibatch_step(fo[i], fl(act)[i], fl(adv_dev)[i], fl(logp)[i], fl(ret_dev)[i])
            steps += 1
    assert steps == n_mb
    got, want = model.params.cpu().numpy(), learner.flat.detach().numpy()
    du_dev, du_ora = torch.from_numpy(got - p0).double(), torch.from_numpy(want - p0).double()
    cos = float(torch.nn.functional.cosine_similarity(du_dev, du_ora, dim=0))
    print(f"{task}: parameter update after {steps} Adam steps: cosine {cos:.5f}, max |dp| {np.abs(got - want).max():.2e}")
    assert float(du_ora.abs().max()) > 1e-4
    assert cos > 0.98
    assert np.abs(got - want).max() <= 2 * 3e-4 * steps + 1e-6
    env.close()


def test_baseline_minibatch_262144_rows_vs_float64_oracle():
    """One fused minibatch at the BASELINE config-3 size (ball3d, 262 144 rows per GPU and optimizer step): weight
    gradients against CPU autograd with float64 accumulation over ALL rows, head outputs on a 4096-row subset."""
    from three_mlagents_b200 import ops

    d, a, rows = 6, 5, 262144
    rng = np.random.default_rng(7)
    params = po.init_params(d, a, 7) + 0.05 * rng.standard_normal(ops.num_params(d, a)).astype(np.float32)
    obs = (rng.standard_normal((rows, d)) * np.array([0.2, 0.2, 1.2, 1.2, 0.8, 0.8])).astype(np.float32)   # ball3d ranges
    act = rng.integers(0, a, rows).astype(np.int32)
    adv = (rng.standard_normal(rows) * 2.0 + 0.3).astype(np.float32)
    ret = (rng.standard_normal(rows) * 3.0).astype(np.float32)
    with torch.no_grad():
        logits0, _ = po.forward(torch.from_numpy(params), torch.from_numpy(obs), d, a)
        lp0, _ = po.categorical(logits0, torch.from_numpy(act))
    old_logp = (lp0.numpy() + 0.1 * rng.standard_normal(rows)).astype(np.float32)      # ratios spread around the clip range
    g_want, st_want, l_want, v_want = _oracle_grad(params, obs, act, adv, old_logp, ret, d, a, dtype=torch.float64)

    dev = torch.device("cuda")
    p_dev = torch.from_numpy(params).to(dev)
    wpack = ops.mlp_pack(p_dev, d, a)
    T, N = 8, rows // 8                                  # the [T,N] buffers the kernels index into
    perm = torch.from_numpy(rng.permutation(rows).astype(np.int32)).to(dev)
    inv = np.empty(rows, np.int64); inv[perm.cpu().numpy()] = np.arange(rows)      # row k of the minibatch = buffer row perm[k]
    to_buf = lambda x, dt: torch.from_numpy(x[inv]).to(dev).to(dt).contiguous()
    obs_b = to_buf(obs, torch.float32)
    act_b, adv_b = to_buf(act, torch.int32).view(T, N), to_buf(adv, torch.float32).view(T, N)
    lp_b, ret_b = to_buf(old_logp, torch.float32).view(T, N), to_buf(ret, torch.float32).view(T, N)
    logits = torch.empty((rows, a), device=dev)
    values = torch.empty(rows, device=dev)
    g_dev, st_dev = ops.ppo_minibatch(p_dev, wpack, obs_b, d, a, act_b, adv_b, lp_b, ret_b, index=perm, rows=rows,
                                      logits=logits, values=values)
    torch.cuda.synchronize()
    sub = rng.choice(rows, 4096, replace=False)
    dl = float(np.abs(logits.cpu().numpy()[sub] - l_want[sub]).max())
    dv = float(np.abs(values.cpu().numpy()[sub] - v_want[sub]).max())
    print(f"262144-row minibatch: head outputs vs fp64 oracle on 4096 rows: logits {dl:.2e}, values {dv:.2e}")
    assert dl < 3e-2 and dv < 3e-2
    _check_grad(g_dev.cpu().numpy(), g_want, d, a, "ball3d B=262144")
    _check_stats(st_dev.cpu().numpy(), st_want, "ball3d B=262144")
