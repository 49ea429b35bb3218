"""CPU self-checks of the SB3-PPO restatement (oracle/ppo_oracle.py).  No reference test pins PPO
numerics ("parity unpinned"), so the oracle carries its own invariants."""
import numpy as np
import torch

from oracle import ppo_oracle as po


def _rollout(T=64, n=37, seed=0, p_done=0.05):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(T, n)).astype(np.float32), rng.normal(size=(T, n)).astype(np.float32),
            (rng.random((T, n)) < p_done), rng.normal(size=n).astype(np.float32))


def test_gae_matches_bruteforce_definition():
    r, v, d, lv = _rollout()
    adv, ret = po.gae(r, v, d, lv, 0.99, 0.95)
    assert adv.dtype == np.float32
    np.testing.assert_allclose(adv, po.gae_bruteforce(r, v, d, lv, 0.99, 0.95), rtol=2e-5, atol=2e-5)
    np.testing.assert_array_equal(ret, adv + v)


def test_gae_lambda_one_is_monte_carlo_return():
    r, v, d, lv = _rollout(p_done=0.0)
    adv, ret = po.gae(r, v, d, lv, 0.9, 1.0)
    T = r.shape[0]
    G = lv.astype(np.float64)
    for t in reversed(range(T)):
        G = r[t] + 0.9 * G
        np.testing.assert_allclose(ret[t], G, rtol=1e-4, atol=1e-4)


def test_gae_scalar_rounding_recipe():
    """The scalar recipe the CUDA kernel implements == NumPy's evaluation."""
    r, v, d, lv = _rollout(T=64, n=257, p_done=0.1)
    adv, _ = po.gae(r, v, d, lv, 0.99, 0.95)
    f = np.float32
    g, gl = f(0.99), f(0.99 * 0.95)
    last = np.zeros(257, f)
    nv = lv.copy()
    for t in reversed(range(64)):
        nnt = (f(1.0) - d[t].astype(f)).astype(f)
        delta = ((r[t] + ((g * nv).astype(f) * nnt).astype(f)).astype(f) - v[t]).astype(f)
        last = (delta + ((gl * nnt).astype(f) * last).astype(f)).astype(f)
        assert np.array_equal(last.view(np.uint32), adv[t].view(np.uint32)), t
        nv = v[t]


def test_policy_shapes_and_init():
    for d, a, n in ((6, 5, 136710), (4, 5, 135686), (21, 3, 143876)):
        p = po.init_params(d, a, 1)
        assert p.shape == (n,) and p.dtype == np.float32
        P = po.unflatten(torch.from_numpy(p), d, a)
        W = P["mlp_extractor.policy_net.2.weight"]
        np.testing.assert_allclose((W @ W.T).numpy(), 2.0 * np.eye(256), atol=1e-4)       # gain sqrt(2)
        Wa = P["action_net.weight"]
        np.testing.assert_allclose((Wa @ Wa.T).numpy(), 1e-4 * np.eye(a), atol=1e-7)      # gain 0.01
        assert float(P["value_net.bias"].abs().sum()) == 0.0


def test_loss_clip_inactive_at_ratio_one_and_gradients():
    torch.manual_seed(0)
    B, A = 64, 5
    logits = torch.randn(B, A, requires_grad=True)
    values = torch.randn(B, requires_grad=True)
    actions = torch.randint(0, A, (B,))
    adv, ret = torch.randn(B), torch.randn(B)
    with torch.no_grad():
        old_logp, _ = po.categorical(logits, actions)
    loss, st = po.ppo_loss(logits, values, actions, adv, old_logp, ret)
    assert st["clip_fraction"] == 0.0 and abs(st["approx_kl"]) < 1e-7
    loss.backward()
    # analytic gradient used by csrc/ppo_kernels.cu:ppo_loss_kernel
    with torch.no_grad():
        a_n = (adv - adv.mean()) / (adv.std() + 1e-8)
        lp = logits - torch.logsumexp(logits, -1, keepdim=True)
        p = lp.exp()
        ent = -(p * lp).sum(-1, keepdim=True)
        onehot = torch.nn.functional.one_hot(actions, A).float()
        g = (-a_n[:, None] * (onehot - p) + 0.01 * p * (lp + ent)) / B
        np.testing.assert_allclose(logits.grad.numpy(), g.numpy(), rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(values.grad.numpy(), (0.5 * 2 * (values - ret) / B).numpy(), rtol=1e-5, atol=1e-8)


def test_permutation_is_a_bijection_with_sb3_flatten_order():
    for T, n in ((128, 100), (7, 13), (1, 5), (64, 64)):
        p = po.permutation(1, 0, T, n)
        assert np.array_equal(np.sort(p), np.arange(T * n))
        q = po.permutation(1, 1, T, n)
        assert not np.array_equal(p, q) or T * n < 3
    # looks shuffled: mean absolute displacement of a uniform permutation is ~ total/3
    p = po.permutation(3, 5, 128, 1000).astype(np.int64)
    assert abs(np.abs(p - np.arange(p.size)).mean() / p.size - 1 / 3) < 0.02


def test_sample_actions_inverse_cdf():
    rng = np.random.default_rng(0)
    logits = rng.normal(size=(200000, 5)).astype(np.float32) * 0 + np.log(np.array([0.1, 0.2, 0.3, 0.25, 0.15], np.float32))
    a = po.sample_actions(logits, rng.random(200000).astype(np.float32))
    np.testing.assert_allclose(np.bincount(a, minlength=5) / a.size, [0.1, 0.2, 0.3, 0.25, 0.15], atol=5e-3)
