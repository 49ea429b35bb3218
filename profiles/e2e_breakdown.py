import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from three_mlagents_b200.vec_env import CudaVecEnv
from three_mlagents_b200 import native
from three_mlagents_b200.native import lib, check
env = CudaVecEnv("ball3d", 65536, seed=1)
env.reset()
acts = np.random.default_rng(0).integers(0, 5, size=(64, 65536)).astype(np.int32)
for i in range(20): env.step(acts[i % 64])
def t(fn, n=300):
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    return (time.perf_counter() - t0) / n * 1e6
print("full step            us", round(t(lambda i: env.step(acts[i % 64])), 1))
nd = native.i64(0)
blk = env._blocks
print("step_block only      us", round(t(lambda i: check(lib.tmla_step_block(env._h, blk._ptr[blk.scratch], C.byref(nd)))), 1))
print("copyto actions       us", round(t(lambda i: np.copyto(env._pin_act, acts[i % 64], casting="unsafe")), 1))
print("acquire + views      us", round(t(lambda i: blk.views(blk.acquire(), 500)), 1))
