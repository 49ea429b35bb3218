"""Per-step latency distribution of CudaVecEnv.step with the results held (two result blocks alternate)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from three_mlagents_b200.vec_env import CudaVecEnv
print("mode", os.environ.get("TMLA_HOST_STEP", "mapped"))
n = 65536
env = CudaVecEnv("ball3d", n, seed=1)
env.reset()
acts = np.random.default_rng(0).integers(0, 5, size=(16, n)).astype(np.int32)
for i in range(20): env.step(acts[i % 16])
for label in ("dropped", "held", "held again"):
    ts = []
    for i in range(400):
        t0 = time.perf_counter()
        if label == "dropped":
            env.step(acts[i % 16])
        else:
            o, r, d, inf = env.step(acts[i % 16])
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"{label:11s} mean {ts.mean():7.1f}  p10 {np.percentile(ts,10):6.1f}  p50 {np.percentile(ts,50):6.1f}  p90 {np.percentile(ts,90):6.1f}  max {ts.max():8.1f}  first5 {np.round(ts[:5],1)}  blocks {len(env._blocks._raw)}")
env.close()
