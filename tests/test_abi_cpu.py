"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/tmla.h declares; metadata calls (no GPU needed) answer; compute calls fail loudly
without a GPU instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tmla.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tmla_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from three_mlagents_b200 import native

    names = _declared()
    assert len(names) >= 30
    for name in names:
        assert hasattr(native.lib, name), f"{name} declared in tmla.h but not exported by libtmla.so"
        assert name in native.SIGNATURES, f"{name} has no ctypes signature in native.py"
    assert set(native.SIGNATURES) == set(names)


def test_metadata_calls_match_reference_spaces():
    from three_mlagents_b200 import native

    lib = native.lib
    assert lib.tmla_version() == 100
    want = {"basic": (21, 3, 50, 12), "ball3d": (6, 5, 200, 48), "gridworld": (4, 5, 100, 36), "push": (4, 5, 120, 28),
            "walljump": (4, 4, 150, 20), "brickbreak": (45, 3, 2000, 88), "bicycle": (7, 3, 2000, 80), "glider": (16, 5, 4000, 112)}
    for name, (d, a, m, sz) in want.items():
        t = lib.tmla_task_from_name(name.encode())
        assert t == native.TASK_IDS[name]
        assert (lib.tmla_task_obs_dim(t), lib.tmla_task_num_actions(t), lib.tmla_task_max_steps(t),
                lib.tmla_task_state_size(t)) == (d, a, m, sz)
    assert lib.tmla_task_from_name(b"labyrinth") == native.TMLA_EINVAL
    assert b"labyrinth" in lib.tmla_last_error()
    # parameter counts of SURVEY.md A8
    assert lib.tmla_mlp_num_params(6, 256, 5) == 136710
    assert lib.tmla_mlp_num_params(4, 256, 5) == 135686
    assert lib.tmla_mlp_num_params(21, 256, 3) == 143876
    assert lib.tmla_mlp_num_params(4, 256, 4) == 135429          # walljump: 135686 - (256 + 1)
    assert lib.tmla_mlp_num_params(45, 256, 3) == 143876 + 2 * 256 * 24   # brickbreak: 45 instead of 21 inputs per tower
    assert lib.tmla_ppo_minibatch_supported(4, 256, 4) == 1 and lib.tmla_ppo_minibatch_supported(21, 256, 3) == 0
    assert lib.tmla_ppo_minibatch_supported(7, 256, 3) == 1 and lib.tmla_mlp_num_params(7, 256, 3) == 143876 - 2 * 256 * 14   # bicycle


def test_wire_structs_match_numpy_dtypes():
    from three_mlagents_b200 import native
    from three_mlagents_b200.vec_env import STATE_DTYPES
    from oracle.envs_oracle import STATE_DTYPES as ORACLE_DTYPES

    for name, tid in native.TASK_IDS.items():
        assert STATE_DTYPES[name].itemsize == native.lib.tmla_task_state_size(tid)
        assert STATE_DTYPES[name] == ORACLE_DTYPES[name]


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from three_mlagents_b200 import native
    from three_mlagents_b200.vec_env import CudaVecEnv

    with pytest.raises(native.TmlaError):
        CudaVecEnv("ball3d", 8)


def test_registry_surface():
    import three_mlagents_b200 as m
    from three_mlagents_b200 import registry

    assert len(registry.TASKS) == 19
    assert registry.CUDA_TASKS == ("basic", "ball3d", "gridworld", "push", "walljump", "brickbreak", "bicycle", "glider")
    assert m.get_task("brick-break").id == "brickbreak"            # tests/test_mlagents.py:47-49
    assert m.get_task("self_driving_car").id == "self-driving-car"
    with pytest.raises(KeyError):
        m.get_task("nope")
    with pytest.raises(ValueError):
        m.make_env("fish")
    card = m.get_task("basic").card()
    assert card["trainable"] is True and "env_factory" not in card
    assert [t.id for t in m.list_tasks(include_roadmap=False)] == ["glider", "brickbreak", "ball3d", "bicycle", "basic", "gridworld", "push", "walljump"]
    fams = [(t.family, t.id) for t in m.list_tasks()]
    assert fams == sorted(fams)


def test_model_path_errors_match_reference(tmp_path, monkeypatch):
    # backend/tests/test_mlagents.py:104-122
    import numpy as np
    from three_mlagents_b200 import training
    from three_mlagents_b200.registry import get_task

    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        training.predict_action("basic", np.zeros(21, np.float32), "missing.zip")
    task = get_task("basic")
    model_path = training.POLICIES_DIR / "resolver_regression_test.zip"
    model_path.parent.mkdir(parents=True, exist_ok=True)
    model_path.write_bytes(b"placeholder")
    assert training._resolve_model_path(task, str(model_path)) == model_path
    assert training._resolve_model_path(task, model_path.name) == model_path
    with pytest.raises(ValueError):
        training.train_task(training.TrainConfig("basic", algorithm="dqn"))   # asked for explicitly: no CUDA backend
    # algorithm=None (what the CLI / REST / websocket pass): the reference's per-task default dqn resolves to PPO with a
    # warning instead of raising; without a GPU the run then stops at tmla_create (no CPU fallback), not at the algorithm
    import warnings
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        try:
            training.train_task(training.TrainConfig("basic", total_timesteps=64, verbose=0))
        except ValueError:
            raise
        except Exception:   # noqa: BLE001 - TmlaError without a CUDA device
            pass
    assert any("trains it with 'ppo'" in str(w.message) for w in caught)
    with pytest.raises(ValueError):
        training.train_task(training.TrainConfig("fish"))
    with pytest.raises(KeyError):
        training.train_task(training.TrainConfig("nope"))
