"""GPU parity tests of the PPO kernels against the CPU restatement of SB3 2.9.0 (oracle/ppo_oracle.py).

Stated tolerances (north_star: "GAE/loss values within a stated float tolerance"):
  GAE            bit-exact (same float32 operation order as NumPy, no FMA)
  permutation    bit-exact vs the NumPy twin; bijection
  MLP fp32       |d| <= 2e-5 abs on logits/values (tanhf/FMA order), gradients rel 2e-3 of max |g|
  loss head      stats rel 1e-5; dlogits/dvalues abs 1e-7
  Adam+clip      params abs 1e-6 after 3 steps
"""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as po

pytestmark = pytest.mark.gpu

SHAPES = [(6, 5), (4, 5), (21, 3), (4, 4), (45, 3), (7, 3), (16, 5)]   # ball3d, gridworld/push, basic, walljump, brickbreak, bicycle, glider


def _ops():
    from three_mlagents_b200 import ops

    return ops


@pytest.mark.parametrize("T,n", [(128, 1000), (5, 33), (1, 7), (64, 4096)])
def test_gae_bit_exact(T, n):
    rng = np.random.default_rng(T * 1000 + n)
    r = rng.normal(size=(T, n)).astype(np.float32)
    v = rng.normal(size=(T, n)).astype(np.float32)
    d = (rng.random((T, n)) < 0.03)
    lv = rng.normal(size=n).astype(np.float32)
    adv, ret = _ops().gae(torch.from_numpy(r).cuda(), torch.from_numpy(v).cuda(),
                          torch.from_numpy(d.astype(np.uint8)).cuda(), torch.from_numpy(lv).cuda(), 0.99, 0.95)
    want_adv, want_ret = po.gae(r, v, d, lv, 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy().view(np.uint32), want_adv.view(np.uint32))
    assert np.array_equal(ret.cpu().numpy().view(np.uint32), want_ret.view(np.uint32))


def test_gae_full_size_properties():
    T, n = 128, 65536
    g = torch.Generator(device="cuda").manual_seed(0)
    r = torch.randn((T, n), device="cuda", generator=g)
    v = torch.randn((T, n), device="cuda", generator=g)
    d = (torch.rand((T, n), device="cuda", generator=g) < 0.01).to(torch.uint8)
    lv = torch.randn(n, device="cuda", generator=g)
    adv, ret = _ops().gae(r, v, d, lv, 0.99, 0.95)
    assert torch.equal(ret, adv + v)                        # returns = advantages + values
    # linearity in (r, V, last_values) at fixed dones
    adv2, _ = _ops().gae(2 * r, 2 * v, d, 2 * lv, 0.99, 0.95)
    assert torch.equal(adv2, 2 * adv)                       # scaling by 2 is exact in binary fp
    # where done[t]: A_t = r_t - V_t exactly (episode boundary cuts the recursion)
    m = d.bool()
    assert torch.equal(adv[m], (r - v)[m])
    cols = torch.arange(0, n, 4099, device="cuda")
    want, _ = po.gae(r[:, cols].cpu().numpy(), v[:, cols].cpu().numpy(), d[:, cols].cpu().numpy().astype(bool),
                     lv[cols].cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(adv[:, cols].cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("T,n", [(128, 1000), (7, 13), (1, 5), (128, 65536)])
def test_permutation(T, n):
    p = _ops().permutation(7, 3, T, n).cpu().numpy()
    if T * n <= 200_000:
        assert np.array_equal(p, po.permutation(7, 3, T, n))
    assert np.array_equal(np.sort(p), np.arange(T * n, dtype=np.int32))


@pytest.mark.parametrize("d,a", SHAPES)
@pytest.mark.parametrize("rows", [1, 100, 1000])
def test_mlp_forward_backward_vs_autograd(d, a, rows):
    ops = _ops()
    rng = np.random.default_rng(d * 100 + rows)
    params = po.init_params(d, a, 3)
    params += rng.normal(scale=0.02, size=params.shape).astype(np.float32)      # non-zero biases
    buf_rows = rows * 3 + 5
    xbuf = rng.normal(size=(buf_rows, d)).astype(np.float32)
    index = rng.permutation(buf_rows)[:rows].astype(np.int32)
    P, X, I = torch.from_numpy(params).cuda(), torch.from_numpy(xbuf).cuda(), torch.from_numpy(index).cuda()
    logits, values, cache = ops.mlp_forward(P, X, d, a, index=I)
    flat = torch.from_numpy(params).requires_grad_(True)
    wl, wv = po.forward(flat, torch.from_numpy(xbuf[index]), d, a)
    np.testing.assert_allclose(logits.cpu().numpy(), wl.detach().numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(values.cpu().numpy(), wv.detach().numpy(), rtol=0, atol=2e-5)
    dl = rng.normal(size=(rows, a)).astype(np.float32) / rows
    dv = rng.normal(size=rows).astype(np.float32) / rows
    grads = ops.mlp_backward(P, X, d, a, cache, torch.from_numpy(dl).cuda(), torch.from_numpy(dv).cuda(), index=I)
    ((wl * torch.from_numpy(dl)).sum() + (wv * torch.from_numpy(dv)).sum()).backward()
    want = flat.grad.numpy()
    got = grads.cpu().numpy()
    p = 0
    for name, shape in po.param_shapes(d, a):
        k = int(np.prod(shape))
        scale = np.abs(want[p:p + k]).max() + 1e-12
        assert np.abs(got[p:p + k] - want[p:p + k]).max() <= 2e-3 * scale + 1e-7, name
        p += k
    # value-only pass with a device-side row count (the timeout-bootstrap path)
    cnt = torch.tensor([max(1, rows // 2)], dtype=torch.int32, device="cuda")
    _, v2, _ = ops.mlp_forward(P, X, d, a, index=I, rows_dev=cnt, want_logits=False,
                               values=torch.full((rows,), 7.0, device="cuda"))
    k = int(cnt.item())
    np.testing.assert_allclose(v2[:k].cpu().numpy(), wv.detach().numpy()[:k], rtol=0, atol=2e-5)
    assert (v2[k:] == 7.0).all()


@pytest.mark.parametrize("A", [3, 5])
@pytest.mark.parametrize("normalize", [True, False])
def test_ppo_loss_vs_autograd(A, normalize):
    ops = _ops()
    rng = np.random.default_rng(A)
    B, total = 4000, 9000
    logits = rng.normal(size=(B, A)).astype(np.float32)
    values = rng.normal(size=B).astype(np.float32)
    index = rng.permutation(total)[:B].astype(np.int32)
    actions = rng.integers(0, A, total).astype(np.int32)
    adv = rng.normal(size=total).astype(np.float32) * 3 + 1
    ret = rng.normal(size=total).astype(np.float32)
    tl = torch.from_numpy(logits).requires_grad_(True)
    tv = torch.from_numpy(values).requires_grad_(True)
    with torch.no_grad():
        lp, _ = po.categorical(tl, torch.from_numpy(actions[index]))
    old_logp_mb = lp.numpy() + rng.normal(scale=0.25, size=B).astype(np.float32)     # exercise both clip sides
    old_logp = np.zeros(total, np.float32)
    old_logp[index] = old_logp_mb
    loss, want = po.ppo_loss(tl, tv, torch.from_numpy(actions[index]), torch.from_numpy(adv[index]),
                             torch.from_numpy(old_logp_mb), torch.from_numpy(ret[index]), normalize=normalize)
    loss.backward()
    assert 0.05 < want["clip_fraction"] < 0.95
    c = lambda x: torch.from_numpy(x).cuda()
    dl, dv, stats = ops.ppo_loss(c(logits), c(values), c(actions), c(adv), c(old_logp), c(ret), index=c(index),
                                 normalize=normalize)
    s = stats.cpu().numpy()
    for i, key in enumerate(("pg_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "loss")):
        assert abs(s[i] - want[key]) <= 1e-5 * max(1.0, abs(want[key])), (key, s[i], want[key])
    np.testing.assert_allclose(dl.cpu().numpy(), tl.grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(dv.cpu().numpy(), tv.grad.numpy(), rtol=1e-5, atol=1e-8)
    if normalize:
        assert abs(s[6] - adv[index].mean()) < 1e-5 and abs(s[7] - adv[index].std(ddof=1)) < 1e-4


def test_adam_clip_vs_torch():
    ops = _ops()
    rng = np.random.default_rng(0)
    n = 136710
    p0 = rng.normal(size=n).astype(np.float32)
    flat = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([flat], lr=3e-4, eps=1e-5)
    P = torch.from_numpy(p0.copy()).cuda()
    m, v = torch.zeros_like(P), torch.zeros_like(P)
    for step in range(1, 4):
        g = rng.normal(size=n).astype(np.float32) * (0.01 if step == 2 else 0.001)   # step 2 clips, others not
        flat.grad = torch.from_numpy(g.copy())
        norm = float(torch.nn.utils.clip_grad_norm_([flat], 0.5))
        opt.step()
        out = ops.adam_clip(P, torch.from_numpy(g).cuda(), m, v, step)
        assert abs(float(out[0]) - norm) <= 1e-5 * norm
    np.testing.assert_allclose(P.cpu().numpy(), flat.detach().numpy(), rtol=0, atol=1e-6)


def test_adam_zero_grads_and_operand_image_refresh():
    """tmla_adam_clip_fused: same update as the plain call, the consumed gradient is cleared, and the bf16 operand images of
    both hidden-layer matrices inside `wpack` equal a full tmla_mlp_pack_bf16 of the new parameters (bit for bit)."""
    ops = _ops()
    d, a = 6, 5
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 136710
    P = torch.randn(n, device="cuda", generator=g) * 0.1
    G = torch.randn(n, device="cuda", generator=g) * 0.01
    P2, G2 = P.clone(), G.clone()
    m, v, m2, v2 = (torch.zeros_like(P) for _ in range(4))
    wpack = ops.mlp_pack(P, d, a)
    stale = wpack.clone()
    ops.adam_clip(P, G, m, v, 1)
    ops.adam_clip(P2, G2, m2, v2, 1, zero_grads=True, wpack=wpack, obs_dim=d, n_actions=a)
    torch.cuda.synchronize()
    assert torch.equal(P, P2) and torch.equal(m, m2) and torch.equal(v, v2)
    assert float(G2.abs().max()) == 0.0 and float(G.abs().max()) > 0.0
    fresh = ops.mlp_pack(P2, d, a)
    w = wpack.view(6, 256, 256)
    assert torch.equal(w[4:].view(torch.int16), fresh.view(6, 256, 256)[4:].view(torch.int16))       # the two operand images
    assert torch.equal(w[:4].view(torch.int16), stale.view(6, 256, 256)[:4].view(torch.int16))       # row-major copies untouched
    assert not torch.equal(w[4:].view(torch.int16), stale.view(6, 256, 256)[4:].view(torch.int16))


def test_step_policy_sampling_and_bootstrap_records():
    """Policy-driven step: sampled actions follow the inverse-CDF twin, log-probs match log_softmax,
    truncation records carry the terminal observation of exactly the time-limit envs."""
    from three_mlagents_b200.vec_env import CudaVecEnv
    from oracle import envs_oracle as eo, philox as px

    ops = _ops()
    n, seed = 3000, 4
    env = CudaVecEnv("gridworld", n, seed=seed)
    ora = eo.OracleVecEnv("gridworld", n, seed=seed)
    rng = np.random.default_rng(0)
    T = 100
    obs_next = torch.empty((n, 4), device="cuda")
    act = torch.empty(n, dtype=torch.int32, device="cuda")
    logp = torch.empty(n, device="cuda")
    rew = torch.empty(n, device="cuda")
    done = torch.empty(n, dtype=torch.uint8, device="cuda")
    cap = 2 * n
    tcount = torch.zeros(1, dtype=torch.int32, device="cuda")
    tindex = torch.zeros(cap, dtype=torch.int32, device="cuda")
    tobs = torch.zeros((cap, 4), device="cuda")
    ep = torch.zeros(4, device="cuda")
    n_trunc = 0
    for t in range(T):
        # logits that make the agent idle (so that many envs reach the 100-step limit) but not always
        logits = rng.normal(size=(n, 5)).astype(np.float32)
        logits[:, 0] += 4.0
        ops.step_policy(env, torch.from_numpy(logits).cuda(), t, obs_next, act, logp, rew, done, trunc_count=tcount,
                        trunc_index=tindex, trunc_obs=tobs, ep_stats=ep)
        u = px.u24(px.stream_block(seed, np.arange(n, dtype=np.uint64), t, px.TAG_SAMPLE)[0])
        a = po.sample_actions(logits, u)
        got_a = act.cpu().numpy()
        assert (got_a != a).mean() < 1e-3            # exp() ulp differences may flip a boundary sample
        tl = torch.from_numpy(logits)
        want_lp, _ = po.categorical(tl, torch.from_numpy(got_a))
        np.testing.assert_allclose(logp.cpu().numpy(), want_lp.numpy(), rtol=0, atol=2e-6)
        o, r, d, tr, info = ora.step(got_a)
        assert np.array_equal(done.cpu().numpy().astype(bool), d)
        assert np.array_equal(obs_next.cpu().numpy(), o) and np.array_equal(rew.cpu().numpy(), r)
        if tr.any():
            k0, k1 = n_trunc, n_trunc + int(tr.sum())
            rec_idx = tindex[k0:k1].cpu().numpy()
            order = np.argsort(rec_idx)
            assert np.array_equal(rec_idx[order], t * n + np.nonzero(tr)[0])
            assert np.array_equal(tobs[k0:k1].cpu().numpy()[order], info["terminal_obs"][tr])
            n_trunc = k1
    assert n_trunc > 0 and int(tcount.item()) == n_trunc
    vals = torch.arange(cap, dtype=torch.float32, device="cuda")
    buf = torch.zeros(T * n, device="cuda")
    ops.bootstrap_add(buf, tcount, tindex, vals, 0.99)
    want = np.zeros(T * n, np.float32)
    want[tindex[:n_trunc].cpu().numpy()] = np.float32(0.99) * np.arange(n_trunc, dtype=np.float32)
    assert np.array_equal(buf.cpu().numpy(), want)
    env.close()


def test_adv_stats_batched_matches_per_minibatch():
    """One launch for all minibatches of an epoch == tmla_adv_stats per minibatch (double accumulation: 1e-9 relative)."""
    from three_mlagents_b200 import ops

    T, n, B = 37, 300, 1000
    g = torch.Generator(device="cuda").manual_seed(3)
    adv = torch.randn((T, n), device="cuda", generator=g) * 3 + 0.7
    perm = ops.permutation(5, 2, T, n)
    total = T * n
    sums = ops.adv_stats_batched(adv, perm, total, B)
    torch.cuda.synchronize()
    assert sums.shape == ((total + B - 1) // B, 3)
    for mb, start in enumerate(range(0, total, B)):
        rows = min(B, total - start)
        one = ops.adv_stats(adv, perm[start:start + rows], rows)
        assert torch.allclose(sums[mb], one, rtol=1e-9, atol=1e-9), (mb, sums[mb], one)
        ref = adv.reshape(-1)[perm[start:start + rows].long()].double()
        assert abs(float(sums[mb, 0]) - float(ref.sum())) < 1e-6 and int(sums[mb, 2]) == rows
