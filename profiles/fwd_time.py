"""Rollout forward (both towers, 65 536 ball3d rows, inference): us per launch.  TMLA_FWD=classic selects the one-tile-at-a-time kernel."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from three_mlagents_b200 import ops
from three_mlagents_b200.ppo import orthogonal_init
for d, a, rows in ((6, 5, 65536), (4, 5, 32768)):
    params = orthogonal_init(d, a, 1).cuda(); wpack = ops.mlp_pack(params, d, a)
    x = torch.randn((rows, d), device="cuda"); lg = torch.empty((rows, a), device="cuda"); vl = torch.empty(rows, device="cuda")
    f = lambda: ops.mlp_forward(params, x, d, a, rows=rows, logits=lg, values=vl, wpack=wpack, keep_act=False)
    for _ in range(5): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): f()
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("TMLA_FWD", "pipe"), (d, a, rows), f"{e0.elapsed_time(e1) * 1e3 / 200:.2f} us per dual-tower forward", flush=True)
