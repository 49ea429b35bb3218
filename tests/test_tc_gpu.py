"""Tensor-core (tcgen05/TMEM) building blocks of the bf16 MLP against a float32 torch reference computed
from the SAME bf16-rounded operands.  Tolerance: bf16 output rounding (2^-9 relative) + fp32 accumulation
order + tanh.approx (2^-11): |d| <= 1e-2 * max(1, |ref|) elementwise, and a much tighter mean error."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from three_mlagents_b200 import native

    return native


def _bf16(x):
    return x.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("M", [128, 1000, 128 * 148 + 77])
def test_tc_linear_forward_and_dgrad(M):
    nat = _lib()
    g = torch.Generator(device="cuda").manual_seed(M)
    A = _bf16(torch.randn((M, 256), device="cuda", generator=g) * 0.5)
    W = _bf16(torch.randn((256, 256), device="cuda", generator=g) / 16)
    bias = (torch.randn(256, device="cuda", generator=g) * 0.1).contiguous()
    aux = _bf16(torch.tanh(torch.randn((M, 256), device="cuda", generator=g)))
    ref_f = torch.tanh(A.float() @ W.float().t() + bias)
    ref_d = (A.float() @ W.float().t()) * (1 - aux.float() ** 2)
    for epi, ref in ((0, ref_f), (1, ref_d)):
        ok = False
        for swap in (0, 1):
            nat.check(nat.lib.tmla_tc_debug(swap))
            out = torch.full((M, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
            nat.check(nat.lib.tmla_tc_linear(epi, nat.ptr(A), nat.ptr(W), nat.ptr(bias), nat.ptr(aux), nat.ptr(out), M,
                                             None, nat.current_stream()))
            torch.cuda.synchronize()
            err = (out.float() - ref).abs()
            tol = 1e-2 * torch.clamp(ref.abs(), min=1.0)
            if bool((err <= tol).all()):
                ok = True
                print(f"tc_linear epi={epi} M={M}: descriptor swap={swap} max err {float(err.max()):.4g} mean {float(err.mean()):.4g}")
                assert swap == 0, "LBO/SBO are exchanged relative to the documented K-major layout"
                assert float(err.mean()) < 2e-3
                break
        nat.check(nat.lib.tmla_tc_debug(0))
        assert ok, f"tc_linear epi={epi} M={M}: max err {float(err.max())}, nan {int(torch.isnan(out.float()).sum())}"


@pytest.mark.parametrize("rows", [128, 1000, 128 * 300 + 5])
def test_tc_wgrad(rows):
    nat = _lib()
    g = torch.Generator(device="cuda").manual_seed(rows)
    X = _bf16(torch.randn((rows, 256), device="cuda", generator=g) * 0.1)
    Y = _bf16(torch.randn((rows, 256), device="cuda", generator=g))
    G = torch.zeros((256, 256), device="cuda")
    nat.check(nat.lib.tmla_tc_debug(0))
    nat.check(nat.lib.tmla_tc_wgrad(nat.ptr(X), nat.ptr(Y), nat.ptr(G), rows, nat.current_stream()))
    torch.cuda.synchronize()
    ref = X.float().t() @ Y.float()
    err = (G - ref).abs()
    scale = float(ref.abs().max())
    print(f"tc_wgrad rows={rows}: max err {float(err.max()):.4g} (ref scale {scale:.3g})")
    assert float(err.max()) <= 2e-4 * scale + 1e-5          # fp32 accumulation of exact bf16 products
    # accumulation semantics: a second call adds
    nat.check(nat.lib.tmla_tc_wgrad(nat.ptr(X), nat.ptr(Y), nat.ptr(G), rows, nat.current_stream()))
    torch.cuda.synchronize()
    assert float((G - 2 * ref).abs().max()) <= 4e-4 * scale + 2e-5


def test_tc_linear_device_row_count():
    nat = _lib()
    M = 1000
    A = _bf16(torch.randn((M, 256), device="cuda"))
    W = _bf16(torch.randn((256, 256), device="cuda") / 16)
    bias = torch.zeros(256, device="cuda")
    out = torch.full((M, 256), 7.0, device="cuda", dtype=torch.bfloat16)
    cnt = torch.tensor([300], dtype=torch.int32, device="cuda")
    nat.check(nat.lib.tmla_tc_linear(0, nat.ptr(A), nat.ptr(W), nat.ptr(bias), None, nat.ptr(out), M, nat.ptr(cnt),
                                     nat.current_stream()))
    torch.cuda.synchronize()
    ref = torch.tanh(A[:300].float() @ W.float().t())
    assert float((out[:300].float() - ref).abs().max()) < 2e-2
    assert bool((out[300:] == 7.0).all())


@pytest.mark.parametrize("task_shape,rows", [((6, 5), 65536), ((6, 5), 128 * 74 * 3 + 77), ((4, 5), 32768), ((4, 4), 1000), ((6, 5), 100)])
def test_pipelined_rollout_forward_is_bit_identical_to_the_classic_kernel(task_shape, rows):
    """csrc/mlp_fwd_pipe.cu (inference: double-buffered TMEM accumulator, driver warp, MMAs of tile k+1 under the epilogue of
    tile k) against csrc/mlp_tc.cu's tower kernel (one tile at a time; selected here by keeping the activations): same FMA order,
    same tanh, same K-step order, same head summation tree -> logits and values bit for bit, full and ragged tile counts."""
    from three_mlagents_b200 import ops
    from three_mlagents_b200.ppo import orthogonal_init

    d, a = task_shape
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    params = orthogonal_init(d, a, 3).cuda()
    params += 0.05 * torch.randn(params.shape, device="cuda", generator=g)
    wpack = ops.mlp_pack(params, d, a)
    x = torch.randn((rows, d), device="cuda", generator=g)
    l_pipe, v_pipe, _ = ops.mlp_forward(params, x, d, a, wpack=wpack, keep_act=False)
    l_ref, v_ref, _ = ops.mlp_forward(params, x, d, a, wpack=wpack, keep_act=True)
    torch.cuda.synchronize()
    assert torch.equal(l_pipe.view(torch.int32), l_ref.view(torch.int32))
    assert torch.equal(v_pipe.view(torch.int32), v_ref.view(torch.int32))
    idx = torch.randperm(rows, device="cuda", generator=g)[: max(rows // 3, 1)].to(torch.int32).contiguous()   # gathered rows
    l_pipe, v_pipe, _ = ops.mlp_forward(params, x, d, a, index=idx, wpack=wpack, keep_act=False)
    l_ref, v_ref, _ = ops.mlp_forward(params, x, d, a, index=idx, wpack=wpack, keep_act=True)
    torch.cuda.synchronize()
    assert torch.equal(l_pipe, l_ref) and torch.equal(v_pipe, v_ref)
