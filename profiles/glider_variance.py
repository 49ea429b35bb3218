"""Run-to-run spread of tests/test_training_gpu.py::test_train_task_glider_runs_on_the_unfused_path: the same config, several
runs and seeds (gradient accumulation uses float atomics, so equal seeds do not give equal runs)."""
import os, sys, tempfile
sys.path.insert(0, os.getcwd())
from three_mlagents_b200.training import TrainConfig, train_task
os.chdir(tempfile.mkdtemp())
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
for seed in (1, 1, 1, 2, 3, 4):
    res = train_task(TrainConfig("glider", total_timesteps=steps, n_envs=2048, eval_episodes=64, eval_freq=10**12, verbose=0, seed=seed,
                                 run_name=f"gl{seed}"), model_kwargs={"n_steps": 128, "batch_size": 32768})
    print(f"seed {seed} steps {steps}: eval mean {res.mean_reward:.2f} std {res.std_reward:.2f}", flush=True)
