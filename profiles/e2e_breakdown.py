"""Where one host step (CudaVecEnv.step, NumPy in / NumPy out) spends its time at 65 536 ball3d envs."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from three_mlagents_b200.vec_env import CudaVecEnv, LazyInfos
from three_mlagents_b200 import native
from three_mlagents_b200.native import lib, check
env = CudaVecEnv("ball3d", 65536, seed=1)
env.reset()
acts = np.random.default_rng(0).integers(0, 5, size=(64, 65536)).astype(np.int32)
acts64 = acts.astype(np.int64)
for i in range(20): env.step(acts[i % 64])
def t(fn, n=300):
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    return (time.perf_counter() - t0) / n * 1e6
print("full step (int32 actions)  us", round(t(lambda i: env.step(acts[i % 64])), 1))
print("full step (int64 actions)  us", round(t(lambda i: env.step(acts64[i % 64])), 1))
nd = native.i64(0)
blk = env._blocks
print("stage_actions int32        us", round(t(lambda i: check(lib.tmla_stage_actions(env._h, acts[i % 64].ctypes.data, 4))), 1))
print("stage_actions int64        us", round(t(lambda i: check(lib.tmla_stage_actions(env._h, acts64[i % 64].ctypes.data, 8))), 1))
print("step_block only (u8 act)   us", round(t(lambda i: (lib.tmla_stage_actions(env._h, acts[i % 64].ctypes.data, 4), check(lib.tmla_step_block(env._h, blk._ptr[blk.scratch], C.byref(nd))))), 1))
def lease(i):
    k = blk.acquire(); v = blk.views(k, 500); del v
print("acquire + leased views     us", round(t(lease), 1))
k = blk.acquire(); obs, rew, done, trunc, rec = blk.views(k, 500)
rec.view(np.int32)[:, 0] = np.arange(500)
def infos(i):
    inf = LazyInfos(65536, done, trunc, rec, 0.0, 7, env._last_reset)
    env._last_reset[rec[:, 0].view(np.int32)] = 7
print("LazyInfos + reset bookkeeping us", round(t(infos), 1))
