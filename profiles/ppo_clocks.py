"""SM clock / power during the PPO update loop (nvidia-smi sampled every 50 ms in the background)."""
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
import torch
from three_mlagents_b200.ppo import bench_ppo
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
t0 = time.time()
out = bench_ppo(0, 0, 1, iters=8)
t1 = time.time()
time.sleep(0.2)
p.terminate()
lines = [l.strip().split(", ") for l in p.stdout.read().strip().splitlines()]
clk = [float(l[0]) for l in lines if len(l) >= 2]
pw = [float(l[1]) for l in lines if len(l) >= 2]
print("ppo", {k: out[k] for k in ("value", "ms_per_iter", "rollout_ms", "update_ms")})
print("samples", len(clk), "clock min/median/max", min(clk), sorted(clk)[len(clk) // 2], max(clk), "power median/max", sorted(pw)[len(pw) // 2], max(pw))
print("reasons", sorted(set(l[2] for l in lines if len(l) >= 3)))
print("trace clk", clk[::4])
print("trace pw ", pw[::4])
