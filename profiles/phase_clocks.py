"""Cycle budget of one tile of tc_tower_train_kernel, measured with clock64 marks (not sampling):
    TMLA_EXTRA_NVCC_FLAGS=-DTMLA_PHASE_CLOCKS python three-mlagents_b200/build.py --force && python profiles/phase_clocks.py
Thread 0 of CTA 0 accumulates the cycles between consecutive marks of the tile loop over one BASELINE config-3
minibatch per tower (262 144 rows, 14 tiles for CTA 0).  Rebuild without the flag afterwards."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from three_mlagents_b200 import native, ops
from three_mlagents_b200.ppo import HIDDEN, orthogonal_init

NAMES = ["wait M4(prev)", "P1 store + XB -> barrier", "M1 window (prefetch, H1 stash) -> MMA1 done", "H1 image read + barrier",
         "P3 (bias, tanh, H2) -> barrier", "MH round trip", "P4 loss -> barrier", "M2 round trip", "P5 dZ2 -> barrier",
         "M3 window (layer 1 of the next tile) -> dgrad done", "dZ2 image read + barrier", "P7 dZ1 -> barrier"]
lib = native.lib
lib.tmla_debug_phase_cycles.restype = C.c_int
lib.tmla_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda", 0)
T, N, D, A = 128, 65536, 6, 5
B = T * N // 32
g = torch.Generator(device=dev).manual_seed(0)
obs = torch.randn((T * N, D), device=dev, generator=g)
act = torch.randint(0, A, (T, N), device=dev, dtype=torch.int32)
adv = torch.randn((T, N), device=dev, generator=g)
logp = -torch.rand((T, N), device=dev, generator=g)
ret = torch.randn((T, N), device=dev, generator=g)
idx = ops.permutation(1, 0, T, N)[:B].contiguous()
params = orthogonal_init(D, A, 1).to(dev)
wpack = ops.mlp_pack(params, D, A)
sums = ops.adv_stats(adv, idx, B)
for _ in range(3):
    ops.ppo_minibatch(params, wpack, obs, D, A, act, adv, logp, ret, index=idx, rows=B, adv_sums=sums)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
lib.tmla_debug_phase_cycles(None, 1)
reps = 5
for _ in range(reps):
    ops.ppo_minibatch(params, wpack, obs, D, A, act, adv, logp, ret, index=idx, rows=B, adv_sums=sums)
torch.cuda.synchronize()
lib.tmla_debug_phase_cycles(buf, 0)
tiles = 14 * 2 * reps            # CTA 0: 14 tiles per tower kernel, 2 towers
tot = sum(buf[i] for i in range(12))
print(f"cycles per tile (CTA 0, both towers averaged): {tot / tiles:.0f}")
for i, nm in enumerate(NAMES):
    print(f"  {buf[i] / tiles:8.0f}  {100.0 * buf[i] / tot:5.1f} %  {nm}")
launches = 2 * reps
print(f"CTA 0 whole kernels: {buf[25]} cycles in {buf[24]} ns of %globaltimer -> SM clock {1e3 * buf[25] / max(buf[24], 1):.0f} MHz during the tower kernels")
print(f"whole kernel, CTA 0, cycles per launch: prologue {buf[16] / launches:.0f}  tile loop {buf[17] / launches:.0f}  flush {buf[18] / launches:.0f}")
print(f"maximum over CTAs and launches: prologue {buf[20]}  tile loop {buf[21]}  flush {buf[22]}  whole kernel {buf[23]}  (1965 cycles = 1 us)")
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(20):
    ops.ppo_minibatch(params, wpack, obs, D, A, act, adv, logp, ret, index=idx, rows=B, adv_sums=sums)
t1.record(); torch.cuda.synchronize()
print(f"ppo_minibatch (2 towers + 2 wgrads, instrumented build): {t0.elapsed_time(t1) / 20 * 1e3:.1f} us")
cta = (C.c_uint * 512)()
lib.tmla_debug_phase_cycles(cta, 2)
import numpy as np
arr = np.array(list(cta), dtype=np.int64).reshape(256, 2)[:148]
loop, sm = arr[:, 0], arr[:, 1]
ntile = np.where(np.arange(148) < 2048 - 148 * 13, 14, 13)
per = loop / ntile
print("per-CTA tile-loop cycles of the last launch (value tower): min/median/max", loop.min(), int(np.median(loop)), loop.max())
print("cycles per tile by CTA: min/median/max", int(per.min()), int(np.median(per)), int(per.max()))
order = np.argsort(per)
print("slowest 12 CTAs (cta, sm, tiles, cycles/tile):", [(int(c), int(sm[c]), int(ntile[c]), int(per[c])) for c in order[-12:]])
print("fastest 6:", [(int(c), int(sm[c]), int(ntile[c]), int(per[c])) for c in order[:6]])
print("mean cycles/tile SM<74:", int(per[sm < 74].mean()), " SM>=74:", int(per[sm >= 74].mean()))
