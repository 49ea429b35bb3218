"""End-to-end learning check on the CUDA backend: PPO with the reference's hyper-parameters (training.py:379-389) except the
rollout/minibatch geometry, which is scaled to thousands of device-resident envs.  Prints the evaluation return against the
registry's reward_threshold (registry.py: ball3d 150.0, gridworld 0.75, push 0.65, walljump 0.7)."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from three_mlagents_b200.registry import get_task
from three_mlagents_b200.training import TrainConfig, train_task

os.chdir(tempfile.mkdtemp())
for task, steps in (("ball3d", 40_000_000), ("gridworld", 40_000_000), ("push", 60_000_000), ("walljump", 40_000_000)):
    t0 = time.time()
    res = train_task(TrainConfig(task, total_timesteps=steps, algorithm="ppo", n_envs=4096, eval_episodes=256, eval_freq=10**12,
                                 verbose=0, run_name=f"{task}_check"),
                     model_kwargs={"n_steps": 128, "batch_size": 32768})
    print(json.dumps({"task": task, "timesteps": steps, "mean_reward": round(res.mean_reward, 3), "std_reward": round(res.std_reward, 3),
                      "reward_threshold": get_task(task).reward_threshold, "wall_s": round(time.time() - t0, 1)}), flush=True)
