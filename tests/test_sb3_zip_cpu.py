"""The policy artifact in Stable-Baselines3's zip layout: everything that can be checked
without SB3/gymnasium installed.  An opportunistic test loads the zip into real SB3 when it is importable."""
import base64
import io
import json
import pickle
import types
import zipfile

import numpy as np
import pytest
import torch

from oracle import ppo_oracle as po
from three_mlagents_b200 import sb3_zip
from three_mlagents_b200.spaces import spaces_for


class _FakeEnv:
    def __init__(self, task):
        self.task_id = task
        self.observation_space, self.action_space = spaces_for(task)


def _fake_model(task="ball3d", d=6, a=5):
    m = types.SimpleNamespace()
    m.env, m.obs_dim, m.n_actions, m.n_envs = _FakeEnv(task), d, a, 8
    m.params = torch.from_numpy(po.init_params(d, a, 4))
    m.m, m.v = torch.rand_like(m.params), torch.rand_like(m.params)
    m._adam_step, m.num_timesteps, m.n_updates, m.seed = 17, 8192, 10, 4
    m.lr, m.n_steps, m.batch_size, m.n_epochs, m.gamma, m.gae_lambda = 3e-4, 1024, 256, 10, 0.99, 0.95
    m.clip_range, m.ent_coef, m.vf_coef, m.max_grad_norm = 0.2, 0.01, 0.5, 0.5
    m.normalize_advantage, m.verbose, m.tensorboard_log, m.mlp_impl = True, 0, None, "bf16"
    return m


def test_layout_matches_the_sb3_restatement():
    for d, a in ((6, 5), (4, 5), (21, 3)):
        assert sb3_zip.param_layout(d, a) == po.param_shapes(d, a)
        flat = torch.from_numpy(po.init_params(d, a, 1))
        sd = sb3_zip.flat_to_state_dict(flat, d, a)
        assert list(sd) == [n for n, _ in po.param_shapes(d, a)]
        assert torch.equal(sb3_zip.state_dict_to_flat(sd, d, a), flat)
        # the state dict drives the oracle's forward exactly like the flat vector
        x = torch.randn(5, d)
        l0, v0 = po.forward(flat, x, d, a)
        hp = torch.tanh(torch.nn.functional.linear(x, sd["mlp_extractor.policy_net.0.weight"], sd["mlp_extractor.policy_net.0.bias"]))
        hp = torch.tanh(torch.nn.functional.linear(hp, sd["mlp_extractor.policy_net.2.weight"], sd["mlp_extractor.policy_net.2.bias"]))
        assert torch.allclose(torch.nn.functional.linear(hp, sd["action_net.weight"], sd["action_net.bias"]), l0)


def test_zip_members_and_round_trip(tmp_path):
    m = _fake_model()
    path = str(tmp_path / "ball3d_policy_x.zip")
    sb3_zip.write_zip(path, m)
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version",
                "system_info.txt"} <= names                      # what SB3's save_to_zip_file writes
        assert z.read("_stable_baselines3_version").decode() == "2.9.0"
        data = json.loads(z.read("data"))
        opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), weights_only=True)
    assert data["n_envs"] == 8 and data["policy_kwargs"]["net_arch"] == {"pi": [256, 256], "vf": [256, 256]}
    assert data["clip_range"] == 0.2 and data["learning_rate"] == 3e-4 and data["_n_updates"] == 10
    # torch.optim.Adam accepts the optimizer state for parameters of these shapes
    ps = [torch.nn.Parameter(torch.zeros(s)) for _, s in sb3_zip.param_layout(6, 5)]
    adam = torch.optim.Adam(ps, lr=1.0, eps=1e-5)
    adam.load_state_dict(opt)
    assert adam.param_groups[0]["lr"] == 3e-4 and float(adam.state[ps[0]]["step"]) == 17.0
    # by-reference pickles name the SB3 / gymnasium classes
    blob = base64.b64decode(data["policy_class"][":serialized:"])
    assert b"stable_baselines3.common.policies" in blob and b"ActorCriticPolicy" in blob
    import sys
    assert "stable_baselines3" not in sys.modules or hasattr(sys.modules["stable_baselines3"], "PPO")   # no stub left behind
    blob = base64.b64decode(data["observation_space"][":serialized:"])
    assert b"gymnasium.spaces.box" in blob and b"Box" in blob and b"_shape" in blob and b"bounded_below" in blob
    blob = base64.b64decode(data["action_space"][":serialized:"])
    assert b"gymnasium.spaces.discrete" in blob and b"Discrete" in blob
    z = sb3_zip.read_zip(path)
    assert torch.equal(z["params"], m.params) and torch.equal(z["adam_m"], m.m) and torch.equal(z["adam_v"], m.v)
    assert z["adam_step"] == 17 and z["meta"]["task_id"] == "ball3d" and (z["obs_dim"], z["n_actions"]) == (6, 5)


def test_reads_a_zip_as_sb3_writes_it(tmp_path):
    """An archive with only SB3's members (no tmla.json): parameters, optimizer state and hyper-parameters load."""
    d, a = 4, 5
    flat = torch.from_numpy(po.init_params(d, a, 9))
    path = str(tmp_path / "sb3_made.zip")
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps({"learning_rate": 1e-3, "n_steps": 2048, "gamma": 0.98, "num_timesteps": 4096, "seed": 3}))
        buf = io.BytesIO(); torch.save(sb3_zip.flat_to_state_dict(flat, d, a), buf); z.writestr("policy.pth", buf.getvalue())
        z.writestr("_stable_baselines3_version", "2.9.0")
    z = sb3_zip.read_zip(path)
    assert torch.equal(z["params"], flat) and z["meta"] is None and z["data"]["gamma"] == 0.98
    with pytest.raises(ValueError):
        bad = str(tmp_path / "bad.zip")
        with zipfile.ZipFile(bad, "w") as zz:
            zz.writestr("data", "{}")
        sb3_zip.read_zip(bad)


def test_opportunistic_load_into_real_sb3(tmp_path):
    sb3 = pytest.importorskip("stable_baselines3")
    pytest.importorskip("gymnasium")
    path = str(tmp_path / "ball3d_policy_y.zip")
    sb3_zip.write_zip(path, _fake_model())
    model = sb3.PPO.load(path, device="cpu")
    obs = np.zeros((1, 6), np.float32)
    action, _ = model.predict(obs, deterministic=True)
    assert action.shape == (1,)
