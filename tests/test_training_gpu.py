"""The reference's CLI/training boundary on the CUDA backend (BASELINE config 1:
`three-mlagents train basic --algorithm ppo --timesteps 25000 --seed 1`)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _train_until(ok, make_config, model_kwargs, seeds=(1, 2, 3)):
    """Learning-outcome tests: gradient accumulation uses float atomics, so equal seeds do not give equal runs, and PPO's
    outcome after a few million steps is heavy-tailed (profiles/glider_variance.py, glider after 5 M steps: 135 / 99 / 205 /
    115 / 327 / 64 over six runs, and one run in seven ended at -142).  A criterion is therefore given up to three seeds; the
    first run that meets it ends the test, and the results of all runs are reported when none does."""
    from three_mlagents_b200.training import train_task

    seen = []
    for seed in seeds:
        res = train_task(make_config(seed), model_kwargs=model_kwargs)
        seen.append(res.mean_reward)
        if ok(res):
            return res
    raise AssertionError(f"no run met the criterion; eval mean rewards per seed: {seen}")


def test_cli_train_basic_ppo_config1(tmp_path, monkeypatch, capsys):
    from three_mlagents_b200 import cli, training

    monkeypatch.chdir(tmp_path)
    cli.main(["train", "basic", "--algorithm", "ppo", "--timesteps", "25000", "--seed", "1", "--run-name", "t1", "--quiet"])
    result = json.loads(capsys.readouterr().out)
    # TrainResult fields of training.py:56-68
    assert set(result) == {"task_id", "algorithm", "run_id", "model_filename", "model_path", "run_dir", "mean_reward",
                           "std_reward", "eval_episodes", "total_timesteps", "metadata_path"}
    assert result["task_id"] == "basic" and result["algorithm"] == "ppo" and result["eval_episodes"] == 50
    assert result["model_filename"] == "basic_policy_t1.zip" and os.path.exists(result["model_path"])
    meta = json.loads(open(result["metadata_path"]).read())
    assert meta["task"]["id"] == "basic" and len(meta["episode_rewards"]) == 50 and meta["config"]["seed"] == 1
    run_dir = result["run_dir"]
    assert os.path.exists(os.path.join(run_dir, "eval", "evaluations.npz"))
    assert os.path.exists(os.path.join(run_dir, "monitor", "0.monitor.csv"))
    assert os.path.exists(os.path.join(run_dir, "tb", "progress.jsonl"))
    # n_envs=1 (registry.py:63), n_steps=1024 -> 25 iterations of 1024 steps
    rows = [json.loads(l) for l in open(os.path.join(run_dir, "tb", "progress.jsonl"))]
    assert len(rows) == 25 and rows[-1]["time/total_timesteps"] == 25600
    capsys.readouterr()
    assert result["mean_reward"] > 0.0

    cli.main(["evaluate", "basic", result["model_filename"], "--episodes", "10"])
    ev = json.loads(capsys.readouterr().out)
    assert ev["episodes"] == 10 and len(ev["episode_rewards"]) == 10 and ev["task_id"] == "basic"
    obs = np.zeros(21, np.float32); obs[10] = 1.0
    assert training.predict_action("basic", obs, result["model_filename"]) in (0, 1, 2)
    assert training.latest_model_filename("basic") == "basic_policy_t1.zip"

    cli.main(["inspect", "ball3d"])
    insp = json.loads(capsys.readouterr().out)
    assert insp["task"] == "ball3d" and "Discrete(5)" in insp["action_space"]


def test_train_task_ball3d_defaults_small(tmp_path, monkeypatch):
    from three_mlagents_b200.training import TrainConfig, train_task

    monkeypatch.chdir(tmp_path)
    res = train_task(TrainConfig("ball3d", total_timesteps=8 * 1024 * 2, eval_episodes=8, verbose=0, run_name="b1"))
    assert res.algorithm == "ppo" and res.total_timesteps == 16384 and np.isfinite(res.mean_reward)


def test_train_task_brickbreak_runs_on_the_unfused_path(tmp_path, monkeypatch):
    """brickbreak (45 inputs, 3 actions) trains through the three-call bf16 path (the fused kernel covers obs_dim <= 6)."""
    from three_mlagents_b200.training import TrainConfig, train_task

    monkeypatch.chdir(tmp_path)
    res = train_task(TrainConfig("brickbreak", total_timesteps=2 * 256 * 128, n_envs=256, eval_episodes=16, eval_freq=10**12, verbose=0,
                                 run_name="bb"), model_kwargs={"n_steps": 128, "batch_size": 8192})
    assert res.algorithm == "ppo" and np.isfinite(res.mean_reward) and res.eval_episodes == 16


def test_train_task_glider_runs_on_the_unfused_path(tmp_path, monkeypatch):
    """glider (16 inputs, 5 actions) trains through the three-call bf16 path (the fused kernel covers obs_dim <= 7) and beats
    the hands-off return: 5 M steps of PPO must lift the evaluation return well above a stall at -50."""
    from three_mlagents_b200.training import TrainConfig, train_task

    monkeypatch.chdir(tmp_path)
    res = _train_until(lambda r: r.mean_reward > 0.0,
                       lambda seed: TrainConfig("glider", total_timesteps=5_000_000, n_envs=2048, eval_episodes=64, eval_freq=10**12,
                                                verbose=0, seed=seed, run_name=f"gl{seed}"), {"n_steps": 128, "batch_size": 32768})
    assert res.algorithm == "ppo" and np.isfinite(res.mean_reward) and res.eval_episodes == 64


def test_train_task_bicycle_learns_to_stay_up(tmp_path, monkeypatch):
    """bicycle (7 inputs, 3 actions; no registry threshold) through the fused update: a random policy falls after ~37 steps
    (tests/golden/bicycle.npz: 863 episodes in 32 000 steps, about -10 + 36 * 0.3 per episode); 20 M steps of PPO must keep the
    bike up for much longer, i.e. collect a clearly positive return."""
    from three_mlagents_b200.ppo import CudaPPO
    from three_mlagents_b200.training import TrainConfig, train_task

    monkeypatch.chdir(tmp_path)
    _train_until(lambda r: r.mean_reward > 20.0,
                 lambda seed: TrainConfig("bicycle", total_timesteps=20_000_000, algorithm="ppo", n_envs=4096, eval_episodes=256,
                                          eval_freq=10**12, verbose=0, seed=seed, run_name=f"bk{seed}"), {"n_steps": 128, "batch_size": 32768})


@pytest.mark.parametrize("task,steps,episodes", [("ball3d", 40_000_000, 256), ("gridworld", 80_000_000, 8192),
                                                 ("push", 120_000_000, 2048), ("walljump", 40_000_000, 256)])
def test_ppo_reaches_registry_reward_threshold(task, steps, episodes, tmp_path, monkeypatch):
    """The whole device-resident pipeline learns: evaluation return above the reference registry's reward_threshold
    (registry.py:74-132: ball3d 150.0, gridworld 0.75, push 0.65, walljump 0.7) with the reference's PPO hyper-parameters
    and a rollout geometry scaled to 4096 envs (profiles/r1_learn_check.txt).  gridworld saturates near 0.78 (three seeds:
    0.777 / 0.777 / 0.781), so it is evaluated on 8192 episodes (standard error 0.007) to keep the 0.03 margin meaningful."""
    from three_mlagents_b200.registry import get_task
    from three_mlagents_b200.training import TrainConfig, train_task

    monkeypatch.chdir(tmp_path)
    thr = get_task(task).reward_threshold
    _train_until(lambda r: r.mean_reward >= thr,
                 lambda seed: TrainConfig(task, total_timesteps=steps, algorithm="ppo", n_envs=4096, eval_episodes=episodes,
                                          eval_freq=10**12, verbose=0, seed=seed, run_name=f"thr{seed}"), {"n_steps": 128, "batch_size": 32768})


@pytest.mark.parametrize("task", ["basic", "ball3d", "gridworld", "push", "walljump", "brickbreak", "bicycle", "glider"])
def test_train_task_with_the_default_algorithm(task, tmp_path, monkeypatch):
    """`algorithm=None` is what the CLI, the REST route and the websocket pass by default.  The registry keeps the reference's
    per-task default (dqn for basic / gridworld / push / walljump, registry.py:61-112); on this backend it resolves to PPO
    (with a warning for the dqn tasks) instead of raising.  Monitor rows are written from the device path."""
    import warnings

    from three_mlagents_b200.registry import get_task
    from three_mlagents_b200.training import TrainConfig, train_task

    monkeypatch.chdir(tmp_path)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        res = train_task(TrainConfig(task, total_timesteps=2 * 64 * 32, n_envs=64, eval_episodes=4, eval_freq=10**12, verbose=0,
                                     run_name="dflt"), model_kwargs={"n_steps": 32, "batch_size": 1024})
    assert res.algorithm == "ppo" and np.isfinite(res.mean_reward)
    assert any("trains it with 'ppo'" in str(w.message) for w in caught) == (get_task(task).default_algorithm != "ppo")
    rows = open(os.path.join(res.run_dir, "monitor", "0.monitor.csv")).read().splitlines()
    assert rows[0].startswith("#") and rows[1] == "r,l,t"
    if task in ("basic", "gridworld", "walljump", "bicycle"):      # short episodes: some finish within 64 steps
        assert len(rows) > 2 and all(len(r.split(",")) == 3 for r in rows[2:])
