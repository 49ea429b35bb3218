import os, sys
sys.path.insert(0, os.getcwd())
import torch
from three_mlagents_b200 import ops
T, N = 128, 65536
g = torch.Generator(device="cuda").manual_seed(0)
rew = torch.randn((T, N), device="cuda", generator=g); val = torch.randn((T, N), device="cuda", generator=g)
done = (torch.rand((T, N), device="cuda", generator=g) < 0.01).to(torch.uint8); lastv = torch.randn(N, device="cuda", generator=g)
adv, ret = torch.empty_like(rew), torch.empty_like(rew)
for _ in range(5): ops.gae(rew, val, done, lastv, 0.99, 0.95, adv, ret)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(50): ops.gae(rew, val, done, lastv, 0.99, 0.95, adv, ret)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 50
print(os.environ.get("TMLA_GAE_UNROLL", "default"), "us", round(us, 2), "GB/s", round(17.0 * T * N / us / 1e3, 1))
