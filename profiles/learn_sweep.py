import json, os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from three_mlagents_b200.registry import get_task
from three_mlagents_b200.training import TrainConfig, train_task
os.chdir(tempfile.mkdtemp())
for seed in (1, 2, 3):
    for task, steps in (("gridworld", 80_000_000), ("push", 120_000_000)):
        t0 = time.time()
        res = train_task(TrainConfig(task, total_timesteps=steps, algorithm="ppo", n_envs=4096, eval_episodes=2048, eval_freq=10**12,
                                     verbose=0, seed=seed, run_name=f"{task}_{seed}"), model_kwargs={"n_steps": 128, "batch_size": 32768})
        print(json.dumps({"task": task, "seed": seed, "mean_reward": round(res.mean_reward, 3), "std": round(res.std_reward, 3),
                          "thr": get_task(task).reward_threshold, "wall_s": round(time.time() - t0, 1)}), flush=True)
