"""Host-step latency of tmla_step_block against the number of envs (run once per TMLA_HOST_STEP mode)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from three_mlagents_b200.vec_env import CudaVecEnv
from three_mlagents_b200 import native
from three_mlagents_b200.native import lib, check
print("mode", os.environ.get("TMLA_HOST_STEP", "mapped"))
for n in (1024, 8192, 65536, 262144):
    env = CudaVecEnv("ball3d", n, seed=1)
    env.reset()
    acts = np.random.default_rng(0).integers(0, 5, size=(16, n)).astype(np.int32)
    for i in range(20): env.step(acts[i % 16])
    nd = native.i64(0)
    blk = env._blocks
    t0 = time.perf_counter()
    for i in range(300): check(lib.tmla_step_block(env._h, blk._ptr[blk.scratch], C.byref(nd)))
    t1 = time.perf_counter()
    for i in range(300): env.step(acts[i % 16])
    t2 = time.perf_counter()
    for i in range(300): o, r, d, inf = env.step(acts[i % 16])
    t3 = time.perf_counter()
    print(f"n={n:7d}  step_block {1e6*(t1-t0)/300:7.1f} us   env.step (dropped) {1e6*(t2-t1)/300:7.1f} us   env.step (held) {1e6*(t3-t2)/300:7.1f} us")
    env.close()
