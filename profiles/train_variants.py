"""One fused PPO minibatch (tmla_ppo_minibatch_bf16) on synthetic ball3d-shaped data: a tiny call first (hang detector), then
the BASELINE size timed with CUDA events.  Run per library variant:  TMLA_LIB=three-mlagents_b200/lib/libtmla_<v>.so python profiles/train_variants.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from three_mlagents_b200 import ops
from three_mlagents_b200.ppo import orthogonal_init
dev = torch.device("cuda")
D, A = 6, 5
g = torch.Generator(device=dev).manual_seed(0)
params = orthogonal_init(D, A, 1).to(dev)
wpack = ops.mlp_pack(params, D, A)
def run(rows, reps):
    T, N = 8, max(rows // 8, 16)
    obs = torch.randn((T * N, D), device=dev, generator=g)
    act = torch.randint(0, A, (T, N), device=dev, dtype=torch.int32)
    adv = torch.randn((T, N), device=dev, generator=g); ret = torch.randn((T, N), device=dev, generator=g)
    logp = -torch.rand((T, N), device=dev, generator=g)
    idx = torch.randperm(T * N, device=dev)[:rows].to(torch.int32).contiguous()
    grads = torch.empty_like(params); stats = torch.zeros(8, device=dev)
    scratch = torch.empty(4 * ((rows + 127) // 128 * 128) * 256, dtype=torch.bfloat16, device=dev)
    sums = ops.adv_stats(adv, idx, rows)
    f = lambda: ops.ppo_minibatch(params, wpack, obs, D, A, act, adv, logp, ret, index=idx, rows=rows, adv_sums=sums, grads=grads, scratch=scratch, stats=stats)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps, float(grads.norm())
print(os.environ.get("TMLA_LIB", "default"), "rows=100 ok", run(100, 1), flush=True)
print("rows=4096 ok", run(4096, 1), flush=True)
for rows in [int(a) for a in sys.argv[1:]] or [262144]:
    us, gn = run(rows, 20)
    print(f"rows={rows}: {us:.1f} us per minibatch (2 towers + wgrad), |g| {gn:.5f}", flush=True)
