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
