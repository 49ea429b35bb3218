"""Micro-driver for ncu: the tcgen05 kernels at one BASELINE minibatch (262144 rows), timed with CUDA events too."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from three_mlagents_b200 import native as nat

M = int(os.environ.get("M", "262144"))
A = (torch.randn((M, 256), device="cuda") * 0.5).to(torch.bfloat16)
aux = torch.tanh(torch.randn((M, 256), device="cuda")).to(torch.bfloat16)
W = (torch.randn((256, 256), device="cuda") / 16).to(torch.bfloat16)
bias = torch.zeros(256, device="cuda")
out = torch.empty_like(A)
G = torch.zeros((256, 256), device="cuda")
s = nat.current_stream()


def run(which):
    if which == 0:
        nat.check(nat.lib.tmla_tc_linear(0, nat.ptr(A), nat.ptr(W), nat.ptr(bias), None, nat.ptr(out), M, None, s))
    elif which == 1:
        nat.check(nat.lib.tmla_tc_linear(1, nat.ptr(A), nat.ptr(W), None, nat.ptr(aux), nat.ptr(out), M, None, s))
    else:
        nat.check(nat.lib.tmla_tc_wgrad(nat.ptr(A), nat.ptr(aux), nat.ptr(G), M, s))


for which, name in ((0, "tc_linear fwd"), (1, "tc_linear dgrad"), (2, "tc_wgrad")):
    for _ in range(3):
        run(which)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run(which)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    flop = 2.0 * M * 256 * 256
    print(f"{name}: {us:.1f} us/launch  {flop / us / 1e6:.1f} TFLOP/s")
