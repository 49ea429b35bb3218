"""Fused PPO minibatch (csrc/mlp_train.cu: forward + loss + backward in one kernel per tower, MN-major UMMA
descriptors) against the unfused bf16 path (same arithmetic, different kernels) and the fp32 path.

Stated tolerances:
  * descriptor probe / wgrad: fp32 accumulation of exact bf16 products -> |d| <= 2e-3 * max(1, |ref|)
  * head outputs (logits, values) fused vs unfused bf16: 5e-3 abs (the fused head GEMM reads H2 rounded to bf16,
    2^-9 relative per element over 256 terms; the unfused head reads the fp32 tanh outputs)
  * gradients fused vs unfused bf16: cosine > 0.9999, relative L2 error < 1e-2 (one extra bf16 rounding of dH1); per
    parameter tensor the fused error against fp32 is at most 1.5x the unfused-bf16 error + 0.2 % of the gradient norm
  * gradients fused vs fp32: cosine > 0.999, relative L2 error < 5e-2 (the bf16-path tolerance of test_trainer_gpu)
  * loss statistics: 1e-4 abs/rel
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nat():
    from three_mlagents_b200 import native

    return native


def _bf16(x):
    return x.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_mn_major_descriptor_probe(mode):
    nat = _nat()
    g = torch.Generator(device="cuda").manual_seed(11 + mode)
    A = _bf16(torch.randn((128, 256), device="cuda", generator=g) * 0.5)
    B = _bf16(torch.randn((256, 256), device="cuda", generator=g) * 0.25)
    rows_out = 256 if mode == 2 else 128
    out = torch.full((rows_out, 256), float("nan"), device="cuda")
    nat.check(nat.lib.tmla_tc_probe(nat.ptr(A), nat.ptr(B), nat.ptr(out), mode, nat.current_stream()))
    torch.cuda.synchronize()
    Af, Bf = A.float(), B.float()
    ref = {0: Af @ Bf.t(), 1: Af @ Bf, 2: Af.t() @ Bf[:128]}[mode]
    err = (out - ref).abs()
    tol = 2e-3 * torch.clamp(ref.abs(), min=1.0)
    assert bool((err <= tol).all()), f"mode {mode}: max err {float(err.max())}, nan {int(torch.isnan(out).sum())}"


def _tile_image(x):
    """[rows,256] bf16 (rows % 128 == 0) -> tile images: per 128 rows, 16-byte chunk (r, cb) at (r/8)*4096 + cb*128 + (r%8)*16."""
    tiles = x.shape[0] // 128
    return x.view(tiles, 16, 8, 32, 8).permute(0, 1, 3, 2, 4).contiguous()


@pytest.mark.parametrize("rows", [128, 1024, 128 * 148 + 128, 128 * 148 * 3 + 384])
def test_wgrad_tiled_matches_torch(rows):
    nat = _nat()
    g = torch.Generator(device="cuda").manual_seed(rows)
    X = _bf16(torch.randn((rows, 256), device="cuda", generator=g) * 0.1)
    Y = _bf16(torch.randn((rows, 256), device="cuda", generator=g))
    Xt, Yt = _tile_image(X), _tile_image(Y)
    G = torch.zeros((256, 256), device="cuda")
    nat.check(nat.lib.tmla_tc_wgrad_tiled(nat.ptr(Xt), nat.ptr(Yt), nat.ptr(G), rows, nat.current_stream()))
    torch.cuda.synchronize()
    ref = X.float().t() @ Y.float()
    err = (G - ref).abs()
    scale = float(ref.abs().max())
    assert float(err.max()) <= 2e-3 * max(1.0, scale), f"rows {rows}: max err {float(err.max())} (scale {scale})"
    # accumulates (+=) like tmla_tc_wgrad
    nat.check(nat.lib.tmla_tc_wgrad_tiled(nat.ptr(Xt), nat.ptr(Yt), nat.ptr(G), rows, nat.current_stream()))
    torch.cuda.synchronize()
    assert float((G - 2 * ref).abs().max()) <= 4e-3 * max(1.0, scale)


def _setup(task, n, T, seed=5):
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    env = CudaVecEnv(task, n, seed=seed)
    m = CudaPPO("MlpPolicy", env, seed=seed, n_steps=T, batch_size=n * T, n_epochs=1, ent_coef=0.01, mlp_impl="bf16")
    with torch.no_grad():                                       # non-trivial biases / heads (own generator: reproducible)
        g = torch.Generator(device=m.params.device).manual_seed(1234 + seed)
        m.params += 0.05 * torch.randn(m.params.shape, device=m.params.device, generator=g)
    m._repack()
    m.collect_rollouts()
    torch.cuda.synchronize()
    return env, m


@pytest.mark.parametrize("task,n,T,rows", [
    ("ball3d", 64, 8, 100),                   # one partial tile
    ("ball3d", 512, 32, 4096),
    ("gridworld", 300, 40, 1000),
    ("push", 512, 90, 128 * 148 * 2 + 77),    # several tiles per CTA + ragged tail
    ("ball3d", 1024, 64, 128 * 148 * 3),      # exactly three full rounds
    ("walljump", 400, 60, 5000),              # 4 actions (NOUT = 4)
    ("bicycle", 400, 60, 128 * 148 + 333),    # 7 inputs (28-byte rows, 4-byte cp.async pieces), 3 actions
])
def test_fused_minibatch_matches_unfused(task, n, T, rows):
    from three_mlagents_b200 import ops

    env, m = _setup(task, n, T)
    d, a = env.obs_dim, env.n_actions
    assert ops.ppo_minibatch_supported(d, a)
    obs_flat = m.obs[:T].reshape(T * n, d)
    idx = ops.permutation(1, 0, T, n)[:rows].contiguous()
    sums = ops.adv_stats(m.adv, idx, rows)
    # unfused bf16 reference path
    l16, v16, c16 = ops.mlp_forward(m.params, obs_flat, d, a, index=idx, wpack=m.wpack)
    dl, dv, st_ref = ops.ppo_loss(l16, v16, m.act, m.adv, m.logp, m.ret, index=idx, adv_sums=sums)
    g_ref = ops.mlp_backward(m.params, obs_flat, d, a, c16, dl, dv, index=idx, wpack=m.wpack).clone()
    # fp32 path
    l32, v32, c32 = ops.mlp_forward(m.params, obs_flat, d, a, index=idx)
    dl32, dv32, _ = ops.ppo_loss(l32, v32, m.act, m.adv, m.logp, m.ret, index=idx, adv_sums=sums)
    g32 = ops.mlp_backward(m.params, obs_flat, d, a, c32, dl32, dv32, index=idx)
    # fused
    logits = torch.full((rows, a), float("nan"), device="cuda")
    values = torch.full((rows,), float("nan"), device="cuda")
    g_f, st_f = ops.ppo_minibatch(m.params, m.wpack, obs_flat, d, a, m.act, m.adv, m.logp, m.ret, index=idx, rows=rows,
                                  adv_sums=sums, logits=logits, values=values)
    torch.cuda.synchronize()
    dl_max, dv_max = float((logits - l16).abs().max()), float((values - v16).abs().max())
    print(f"{task} rows={rows}: head outputs fused vs unfused-bf16: logits {dl_max:.2e} values {dv_max:.2e}")
    assert dl_max < 5e-3 and dv_max < 5e-3
    assert torch.allclose(st_f[:6], st_ref[:6], rtol=1e-4, atol=1e-4), (st_f, st_ref)
    assert torch.allclose(st_f[6:], st_ref[6:], rtol=1e-6, atol=1e-7)
    off = 0
    names = []
    for t in ("pi", "vf"):
        for nm, sz in (("W1", 256 * d), ("b1", 256), ("W2", 65536), ("b2", 256)):
            names.append((f"{t}.{nm}", off, off + sz)); off += sz
    for nm, sz in (("Wa", a * 256), ("ba", a), ("Wv", 256), ("bv", 1)):
        names.append((nm, off, off + sz)); off += sz
    assert off == g_f.numel()
    for nm, lo, hi in names:
        # per tensor, both bf16 paths against the fp32 gradient: the fused kernel must be as accurate as the three-call path
        # (1.5x its error plus 0.2 % of the whole gradient's norm).  Comparing fused and unfused directly is meaningless for a
        # slice whose gradient nearly cancels (walljump's pi.W2 with its 40 distinct observations carries only bf16 noise).
        err_f, err_u = float((g_f[lo:hi] - g32[lo:hi]).norm()), float((g_ref[lo:hi] - g32[lo:hi]).norm())
        assert err_f <= 1.5 * err_u + 2e-3 * float(g32.norm()), f"{task} {nm}: fused error {err_f} vs unfused-bf16 error {err_u} (|g| = {float(g32.norm())})"
    cos = float(torch.nn.functional.cosine_similarity(g_f, g_ref, dim=0))
    rel = float((g_f - g_ref).norm() / g_ref.norm())
    cos32 = float(torch.nn.functional.cosine_similarity(g_f, g32, dim=0))
    rel32 = float((g_f - g32).norm() / g32.norm())
    print(f"{task} rows={rows}: fused vs bf16 cos {cos:.7f} rel {rel:.5f}; vs fp32 cos {cos32:.6f} rel {rel32:.4f}")
    assert cos > 0.9999 and rel < 1e-2
    assert cos32 > 0.999 and rel32 < 5e-2
    env.close()


def test_fused_training_tracks_unfused_and_learns():
    """Same seeds: one train() with the fused update lands next to the unfused update; ball3d returns improve."""
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    deltas = []
    for fused in (True, False):
        env = CudaVecEnv("ball3d", 1024, seed=2)
        m = CudaPPO("MlpPolicy", env, seed=2, n_steps=32, batch_size=8192, n_epochs=1, ent_coef=0.01, fused_update=fused)
        assert m.fused_update == fused
        p0 = m.params.clone()
        m.collect_rollouts()
        m.train()
        torch.cuda.synchronize()
        assert torch.isfinite(m.params).all()
        deltas.append((m.params - p0).clone())
        env.close()
    cos = float(torch.nn.functional.cosine_similarity(deltas[0], deltas[1], dim=0))
    print(f"parameter-update cosine fused vs unfused after 4 Adam steps: {cos:.5f}")
    assert cos > 0.98
    env = CudaVecEnv("ball3d", 4096, seed=1)
    m = CudaPPO("MlpPolicy", env, seed=1, n_steps=64, batch_size=32768, n_epochs=4, ent_coef=0.01)
    m.learn(4096 * 64 * 10)
    rows = m.logger_rows
    print("ball3d ep_rew_mean:", [round(r["rollout/ep_rew_mean"], 2) for r in rows])
    assert rows[-1]["rollout/ep_rew_mean"] > rows[0]["rollout/ep_rew_mean"] + 5.0
    env.close()


_VARIANT_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r})
from three_mlagents_b200 import ops
from three_mlagents_b200.ppo import orthogonal_init
D, A, rows = {D}, {A}, {rows}
g = torch.Generator(device="cuda").manual_seed(5)
params = orthogonal_init(D, A, 1).cuda()
wpack = ops.mlp_pack(params, D, A)
T, N = 8, rows // 4
obs = torch.randn((T * N, D), device="cuda", generator=g)
act = torch.randint(0, A, (T, N), device="cuda", dtype=torch.int32, generator=g)
adv = torch.randn((T, N), device="cuda", generator=g); ret = torch.randn((T, N), device="cuda", generator=g)
logp = -torch.rand((T, N), device="cuda", generator=g)
idx = torch.randperm(T * N, device="cuda", generator=g)[:rows].to(torch.int32).contiguous()
grads = torch.zeros_like(params); stats = torch.zeros(8, device="cuda")
scratch = torch.empty(4 * ((rows + 127) // 128 * 128) * 256, dtype=torch.bfloat16, device="cuda")
sums = ops.adv_stats(adv, idx, rows)
m = torch.zeros_like(params); v = torch.zeros_like(params)
for step in (1, 2, 3):                         # three minibatches with the optimizer step between them (exercises every boundary)
    ops.ppo_minibatch(params, wpack, obs, D, A, act, adv, logp, ret, index=idx, rows=rows, adv_sums=sums, grads=grads, scratch=scratch,
                      stats=stats, grads_zeroed=True, accumulate_stats=True)
    if step == 1:
        g1 = grads.clone()
    ops.adam_clip(params, grads, m, v, step, max_grad_norm=0.5, lr=3e-4, eps=1e-5, zero_grads=True, wpack=wpack, obs_dim=D, n_actions=A)
torch.cuda.synchronize()
torch.save({{"g1": g1.cpu(), "params": params.cpu(), "stats": stats.cpu()}}, {out!r})
"""


@pytest.mark.parametrize("env", [{"TMLA_PDL": "0", "TMLA_OPT_PDL": "0"}, {"TMLA_PDL": "1"}, {"TMLA_WGRAD_H1": "recompute"}, {"TMLA_ADAM": "split"}],
                         ids=["ordinary-launches", "dependent-towers-only", "wgrad-recomputes-h1", "two-launch-optimizer"])
def test_launch_and_wgrad_variants_agree_with_the_default(env, tmp_path):
    """The A/B switches kept in the library (read once per process, hence subprocesses): ordinary launches instead of programmatic
    dependent ones, the weight-gradient kernel that recomputes H1, the two-launch optimizer step.  Same seeded minibatch, three
    update steps: the first gradient and the parameters after three steps must agree with the default build of the path
    (float atomics reorder sums, hence 2e-5 of the gradient norm, and all but a handful of parameters within 1e-6, not bit equality)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, extra in (("default", {}), ("variant", env)):
        out = str(tmp_path / f"{name}.pt")
        script = _VARIANT_SCRIPT.format(root=root, D=6, A=5, rows=128 * 148 + 77, out=out)
        e = {k: v for k, v in os.environ.items() if not k.startswith("TMLA_")}
        e.update(extra)
        r = subprocess.run([sys.executable, "-c", script], env=e, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = torch.load(out)
    a, b = outs["default"], outs["variant"]
    gn = float(a["g1"].norm())
    assert gn > 0 and float((a["g1"] - b["g1"]).norm()) <= 2e-5 * gn
    dp = (a["params"] - b["params"]).abs()      # Adam's first steps move a weight by ~lr whatever its gradient: a near-zero gradient
    assert float((dp > 1e-6).float().mean()) < 1e-3 and float(dp.norm()) <= 1e-4 * float(a["params"].norm())   # may flip sign
    assert torch.allclose(a["stats"], b["stats"], rtol=1e-4, atol=1e-5)
