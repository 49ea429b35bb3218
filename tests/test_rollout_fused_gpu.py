"""`tmla_rollout` (csrc/rollout.cu): the whole PPO rollout — SB3's collect_rollouts + compute_returns_and_advantage, entered
from backend/mlagents/training.py:166 — as ONE call replaying a CUDA graph, against the same launches issued one by one from
Python (`rollout_impl="steps"`: tmla_mlp_forward[_bf16] + tmla_step_policy per step, then tmla_bootstrap_add + tmla_gae).
Integer / byte / index buffers and every float buffer must be BIT-IDENTICAL (same kernels, same Philox counters), over several
consecutive rollouts (graph replays advance the device-side step counter) and interleaved with updates."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BUFFERS = ("obs", "act", "logp", "rew", "val", "done", "adv", "ret", "last_values")


def _pair(task, n, T, impl, seed=9):
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.vec_env import CudaVecEnv

    models = []
    for rollout_impl in ("fused", "steps"):
        env = CudaVecEnv(task, n, seed=seed)
        models.append(CudaPPO("MlpPolicy", env, seed=seed, n_steps=T, batch_size=n * T // 2, n_epochs=1, ent_coef=0.01, mlp_impl=impl,
                              rollout_impl=rollout_impl))
    return models


@pytest.mark.parametrize("task,n,T,impl", [("ball3d", 512, 32, "bf16"), ("gridworld", 300, 120, "bf16"), ("basic", 200, 64, "bf16"),
                                           ("push", 256, 130, "fp32"), ("bicycle", 256, 40, "bf16")])
def test_fused_rollout_is_bit_identical_to_per_step_calls(task, n, T, impl):
    fused, steps = _pair(task, n, T, impl)
    assert fused.rollout_impl == "fused" and steps.rollout_impl == "steps"
    for it in range(3):
        for m in (fused, steps):
            m.collect_rollouts()
        torch.cuda.synchronize()
        for name in BUFFERS:
            a, b = getattr(fused, name), getattr(steps, name)
            assert torch.equal(a.view(torch.uint8) if a.dtype == torch.uint8 else a.view(torch.int32),
                               b.view(torch.uint8) if b.dtype == torch.uint8 else b.view(torch.int32)), (task, it, name)
        assert int(fused.trunc_count.item()) == int(steps.trunc_count.item())
        # episode statistics are float atomicAdd sums: same addends, order not fixed
        assert torch.allclose(fused.ep_stats, steps.ep_stats, rtol=1e-5, atol=0) and fused.ep_stats[2] == steps.ep_stats[2]
        assert fused.env.step_count == steps.env.step_count == (it + 1) * T
        if it == 1:                                            # an update between rollouts: the graph reads the parameters in place
            for m in (fused, steps):
                m.train()
            torch.cuda.synchronize()
            # (the update accumulates gradients with float atomics: two runs agree to rounding, not bit for bit)
            assert torch.allclose(fused.params, steps.params, rtol=0, atol=2e-5)
            steps.params.copy_(fused.params)                   # keep the two policies in lockstep for the next rollout
            steps._repack()
    if task not in ("ball3d", "bicycle"):                      # (time limits 200 / 2000 steps: no timeout inside these rollouts)
        assert int(fused.trunc_count.item()) > 0               # the timeout-bootstrap branch ran
    for m in (fused, steps):
        m.env.close()


def test_fused_rollout_follows_reseeding_and_host_steps():
    """Calls that move the handle's step index outside the graph (a host-path VecEnv.step) are picked up by the next replay."""
    fused, steps = _pair("gridworld", 128, 16, "bf16")
    for m in (fused, steps):
        m.collect_rollouts()
        m.env.step(np.zeros(128, np.int64))                    # host path: advances the step index and the env state
        m._last_obs_valid = False                              # restart from a fresh reset observation
        m.collect_rollouts()
    torch.cuda.synchronize()
    for name in BUFFERS:
        assert torch.equal(getattr(fused, name), getattr(steps, name)), name
    for m in (fused, steps):
        m.env.close()
