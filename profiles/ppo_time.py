import sys, json, os
sys.path.insert(0, os.getcwd())
from three_mlagents_b200.ppo import bench_ppo
for fused in (True, False):
    r = bench_ppo(0, 0, 1, iters=2, fused_update=fused)
    print(json.dumps({"fused": fused, "sps": r["value"], "rollout_ms": r["rollout_ms"], "update_ms": r["update_ms"], "ep_rew_mean": r["ep_rew_mean"], "kl": r["approx_kl"]}), flush=True)
