"""tmla_step_block_begin / _end timed separately (65 536 ball3d envs, int32 and int64 actions) — run per TMLA_HOST_CHUNKS."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from three_mlagents_b200.vec_env import CudaVecEnv
from three_mlagents_b200 import native
from three_mlagents_b200.native import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = CudaVecEnv("ball3d", n, seed=1)
env.reset()
acts = np.random.default_rng(0).integers(0, 5, size=(64, n)).astype(np.int32)
acts64 = acts.astype(np.int64)
for i in range(20): env.step(acts[i % 64])
blk = env._blocks
ptr = blk._ptr[blk.scratch]
nd = native.i64(0)
for name, a in (("int32", acts), ("int64", acts64), ("int32 same array (cache-hot)", acts[:1])):
    tb = te = 0.0
    reps = 300
    for i in range(reps):
        x = a[i % len(a)]
        t0 = time.perf_counter()
        check(lib.tmla_step_block_begin(env._h, x.ctypes.data, x.dtype.itemsize, ptr))
        t1 = time.perf_counter()
        check(lib.tmla_step_block_end(env._h, ptr, C.byref(nd)))
        t2 = time.perf_counter()
        tb += t1 - t0; te += t2 - t1
    print(f"chunks={os.environ.get('TMLA_HOST_CHUNKS','default')} n={n} {name}: begin {tb/reps*1e6:.1f} us  end {te/reps*1e6:.1f} us  total {(tb+te)/reps*1e6:.1f} us", flush=True)
