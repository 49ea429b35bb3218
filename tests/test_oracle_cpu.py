"""CPU tests: the oracle restatement against the reference-made golden vectors,
Philox known-answer vectors, reset distributions against reference samples."""
import numpy as np
import pytest

from oracle import envs_oracle as eo
from oracle import philox as px
import replay_util as _replay

TASKS = ("basic", "ball3d", "gridworld", "push", "walljump", "brickbreak", "bicycle", "glider")


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = px.philox4x32_10([np.array([c]) for c in ctr], key)
        assert tuple(int(x[0]) for x in got) == want


@pytest.mark.parametrize("task", TASKS)
def test_oracle_matches_reference_golden(task):
    g = _replay.load(task)
    acts = g["actions"]
    T, E = acts.shape
    st = _replay.initial_state(task, g, eo.STATE_DTYPES[task])
    np.testing.assert_array_equal(eo.observe(task, st), g["init_obs"])
    inexact = 0
    for t in range(T):
        obs, rew, term, trunc = eo.transition(task, st, acts[t])
        assert np.array_equal(term, g["terminated"][t]), (task, t)
        assert np.array_equal(trunc, g["truncated"][t]), (task, t)
        # bit-exact: integer tasks by construction, ball3d because the restatement
        # reproduces NumPy's promotion/rounding order
        if task in _replay.LIBM_TASKS:   # bit-exact on the host that made the fixture; elsewhere libm/SVML/BLAS may differ by an ulp
            tol = _replay.LIBM_TASKS[task]
            inexact += int((obs.view(np.uint32) != g["obs"][t].view(np.uint32)).sum()) + int((rew.view(np.uint32) != g["reward"][t].view(np.uint32)).sum())
            np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=tol["obs_atol"])
            np.testing.assert_allclose(rew, g["reward"][t], rtol=0, atol=tol["reward_atol"])
        else:
            assert np.array_equal(obs.view(np.uint32), g["obs"][t].view(np.uint32)), (task, t)
            assert np.array_equal(rew.view(np.uint32), g["reward"][t].view(np.uint32)), (task, t)
        done = term | trunc
        st = _replay.inject_resets(task, g, t, st, done)
        if done.any():
            idx = np.nonzero(done)[0]
            np.testing.assert_array_equal(eo.observe(task, st)[idx], g["reset_obs"][t][idx])
    print(f"{task}: {inexact} values differ in the last bits from the fixture")
    assert inexact <= 1e-3 * T * E * (obs.shape[1] + 1)


def test_reward_luts_are_f32_of_double():
    g = _replay.load("push")
    assert np.array_equal(g["reward"], g["reward64"].astype(np.float32))
    assert set(np.unique(g["reward"])).issubset(set(eo.PUSH_LUT.tolist()) | {np.float32(1.0)})
    assert eo.BASIC_LUT.view(np.uint32).tolist() == [0xBC23D70A, 0x3DB851EC, 0x3F7D70A4]
    w = _replay.load("walljump")
    assert np.array_equal(w["reward"], w["reward64"].astype(np.float32))
    assert set(np.unique(w["reward"]).tolist()) == set(eo.WALLJUMP_LUT.tolist())          # all four values occur in the trace
    assert eo.WALLJUMP_LUT.view(np.uint32).tolist() == [0xBC23D70A, 0xBCF5C28F, 0xBD23D70A, 0x3F800000]   # csrc/envs.cuh
    # the device divides in f32: f32(k/19.0) == f32(k)/f32(19) for every reachable numerator
    assert all(np.float32(k / 19.0) == np.float32(k) / np.float32(19) for k in range(-9, 20))


def test_reference_known_answer_basic():
    # backend/tests/test_mlagents.py:32-45 — reset -> position 10; step(2) -> 11, not done
    st = eo.draw_reset("basic", 1, np.arange(1), 0)
    assert int(st["pos"][0]) == 10
    obs, rew, term, trunc = eo.transition("basic", st, np.array([2]))
    assert int(st["pos"][0]) == 11 and not term[0] and not trunc[0]
    assert obs.shape == (1, 21) and obs[0, 11] == 1.0 and obs.sum() == 1.0


def test_walljump_reset_distribution_matches_reference():
    """walljump.py:39-45: agent at 0, grounded, wall present with probability 0.7."""
    ref = np.load(_replay.GOLDEN + "/walljump_resets.npz")
    n = 200_000
    st = eo.draw_reset("walljump", 7, np.arange(n), 3)
    assert (st["agent_x"] == 0).all() and (st["in_air"] == 0).all() and (ref["agent_x"] == 0).all() and (ref["in_air"] == 0).all()
    assert set(np.unique(st["wall"]).tolist()) == {0, 1} and set(np.unique(ref["wall"]).tolist()) == {0, 1}
    assert abs(st["wall"].mean() - 0.7) < 4 * np.sqrt(0.21 / n) and abs(ref["wall"].mean() - 0.7) < 4 * np.sqrt(0.21 / len(ref["wall"]))


def test_brickbreak_reset_distribution_matches_reference():
    """brick_break.py:39-46: ball at (20, 10), paddle at 20, all 40 bricks, serve angle ~ U(pi/4, 3pi/4) at speed 1.5."""
    ref = np.load(_replay.GOLDEN + "/brickbreak_resets.npz")
    n = 200_000
    st = eo.draw_reset("brickbreak", 7, np.arange(n), 3)
    for s in (st, ref):
        assert (s["pos"] == np.array([20.0, 10.0])).all() and (s["paddle"] == 20.0).all() and (s["bricks"] == 1).all()
        speed = np.linalg.norm(s["vel"], axis=1)
        assert np.abs(speed - 1.5).max() < 1e-14
        ang = np.arctan2(s["vel"][:, 1], s["vel"][:, 0])
        assert ang.min() >= np.pi / 4 - 1e-12 and ang.max() <= 3 * np.pi / 4 + 1e-12
    ang = np.arctan2(st["vel"][:, 1], st["vel"][:, 0])
    assert abs(ang.mean() - np.pi / 2) < 4 * (np.pi / 2) / np.sqrt(12 * n) and abs(ang.std() - (np.pi / 2) / np.sqrt(12)) < 0.005
    # the deterministic polynomial agrees with libm to the last bits
    y = np.linspace(-np.pi / 4, np.pi / 4, 100001)
    s_, c_ = eo.sin_cos_quarter(y)
    assert np.abs(s_ - np.sin(y)).max() < 3e-16 and np.abs(c_ - np.cos(y)).max() < 3e-16


def test_bicycle_reset_distribution_matches_reference():
    """bicycle.py:40-58: origin, zero heading/steer, lean and lean rate ~ U(-0.1, 0.1), goal at radius U(15, 25) and bearing
    U(-pi/4, pi/4), dist_to_goal = |goal|.  20 000 Philox draws (dot2 goes through np.dot per env) vs 4096 reference resets."""
    ref = np.load(_replay.GOLDEN + "/bicycle_resets.npz")
    n = 20_000
    st = eo.draw_reset("bicycle", 7, np.arange(n), 3)
    for s in (st, ref):
        for k in ("x", "z", "theta", "delta"):
            assert (s[k] == 0.0).all()
        assert np.abs(s["phi"]).max() <= 0.1 and np.abs(s["phi_dot"]).max() <= 0.1
        rad, ang = np.linalg.norm(s["goal"], axis=1), np.arctan2(s["goal"][:, 1], s["goal"][:, 0])
        assert rad.min() >= 15.0 - 1e-12 and rad.max() <= 25.0 + 1e-12 and np.abs(ang).max() <= np.pi / 4 + 1e-12
        assert np.abs(s["dist"] - rad).max() < 1e-13
        m = len(rad)
        assert abs(rad.mean() - 20.0) < 4 * 10.0 / np.sqrt(12 * m) and abs(ang.mean()) < 4 * (np.pi / 2) / np.sqrt(12 * m)
        assert abs(s["phi"].mean()) < 4 * 0.2 / np.sqrt(12 * m) and abs(s["phi"].std() - 0.2 / np.sqrt(12)) < 0.002
        assert abs(np.corrcoef(s["phi"], s["phi_dot"])[0, 1]) < 5 / np.sqrt(m)


@pytest.mark.parametrize("task", ("ball3d", "gridworld", "push"))
def test_reset_distribution_matches_reference(task):
    ref = np.load(_replay.GOLDEN + f"/{task}_resets.npz")
    n = 200_000
    st = eo.draw_reset(task, 7, np.arange(n), 3, episode=np.full(n, 3))
    if task == "ball3d":
        for key, half in (("rot", eo.MAX_TILT / 2), ("pos", 1.5), ("vel", 1.0)):
            x = st[key].astype(np.float64)
            assert np.abs(x).max() <= half * (1 + 1e-6) and np.abs(ref[key]).max() <= half * (1 + 1e-6)
            assert abs(x.mean()) < 4 * half / np.sqrt(3 * n) * 2
            assert abs(x.std() - half / np.sqrt(3)) < 0.01 * half
            assert abs(ref[key].std() - half / np.sqrt(3)) < 0.05 * half
            # both are f32-rounded doubles
            assert np.array_equal(x.astype(np.float32).astype(np.float64), x)
    elif task == "gridworld":
        cells = lambda a: a[:, 0] * 5 + a[:, 1]
        for s in (st, ref):
            a, g_, r = cells(s["agent"]), cells(s["green"]), cells(s["red"])
            assert ((a != g_) & (a != r) & (g_ != r)).all()
            assert a.min() >= 0 and a.max() <= 24 and r.max() <= 24 and g_.max() <= 24
        for key in ("agent", "green", "red"):
            h = np.bincount(cells(st[key]), minlength=25) / n
            assert np.abs(h - 1 / 25).max() < 0.003
        assert abs(st["goal_type"].mean() - 0.5) < 0.01 and abs(ref["goal_type"].mean() - 0.5) < 0.05
        # joint: red given (agent, green) is uniform over the 23 free cells
        pair = cells(st["agent"]) * 25 + cells(st["green"])
        sel = pair == pair[0]
        assert len(np.unique(cells(st["red"])[sel])) == 23
    else:
        cells = lambda a: a[:, 0] * 6 + a[:, 1]
        for s in (st, ref):
            assert (cells(s["agent"]) != cells(s["box"])).all()
            assert s["goal_x"].min() >= 0 and s["goal_x"].max() <= 5
        for key in ("agent", "box"):
            h = np.bincount(cells(st[key]), minlength=36) / n
            assert np.abs(h - 1 / 36).max() < 0.003
        h = np.bincount(st["goal_x"], minlength=6) / n
        assert np.abs(h - 1 / 6).max() < 0.005


@pytest.mark.parametrize("task", TASKS)
def test_oracle_vecenv_autoreset_contract(task):
    env = eo.OracleVecEnv(task, 64, seed=3)
    rng = np.random.default_rng(0)
    ep_ret = np.zeros(64, np.float64)
    ep_len = np.zeros(64, np.int64)
    seen_done = 0
    for t in range(300):
        a = rng.integers(0, env.n_actions, 64)
        obs, rew, done, tl_trunc, info = env.step(a)
        ep_ret += rew
        ep_len += 1
        assert obs.shape == (64, env.obs_dim) and obs.dtype == np.float32
        if done.any():
            i = np.nonzero(done)[0]
            seen_done += i.size
            assert np.allclose(info["episode_return"][i], ep_ret[i], rtol=1e-5, atol=1e-5)
            assert np.array_equal(info["episode_length"][i], ep_len[i])
            assert (env.state["steps"][i] == 0).all()
            assert (ep_len[i] <= env.max_steps).all()
            assert (tl_trunc[i] == (ep_len[i] == env.max_steps) & ~info["terminated"][i]).all()
            ep_ret[i] = 0
            ep_len[i] = 0
    assert seen_done > 0


@pytest.mark.parametrize("task", ("ball3d", "gridworld", "push", "basic"))
def test_scalar_port_matches_reference_golden(task):
    """oracle/ref_port.py (the timed CPU baseline) reproduces the reference transitions bit for bit."""
    from oracle import ref_port

    g = _replay.load(task)
    T, E = 300, 6
    envs = []
    for i in range(E):
        e = ref_port.TASKS[task]()
        if task == "basic":
            e.p = int(g["init_pos"][i])
        elif task == "ball3d":
            e.rot, e.pos, e.vel = g["init_rot"][i].astype(np.float32), g["init_pos"][i].copy(), g["init_vel"][i].copy()
        elif task == "gridworld":
            e.agent, e.green, e.red, e.kind = tuple(g["init_agent"][i]), tuple(g["init_green"][i]), tuple(g["init_red"][i]), int(g["init_goal_type"][i])
        else:
            e.agent, e.box, e.goal = tuple(g["init_agent"][i]), tuple(g["init_box"][i]), (int(g["init_goal_x"][i]), 5)
        e.t = 0
        envs.append(e)
    steps = [0] * E
    for t in range(T):
        for i, e in enumerate(envs):
            obs, r, done = e.step(int(g["actions"][t, i]))
            steps[i] += 1
            hit = steps[i] >= e.limit
            term, trunc = (done, hit and not done) if task == "basic" else (bool(done and not hit), bool(hit))
            assert term == g["terminated"][t, i] and trunc == g["truncated"][t, i], (task, t, i)
            assert np.array_equal(np.asarray(obs, np.float32).view(np.uint32), g["obs"][t, i].view(np.uint32)), (task, t, i)
            assert np.float32(r).view(np.uint32) == g["reward"][t, i].view(np.uint32), (task, t, i)
            if term or trunc:
                steps[i] = 0
                e.t = 0
                if task == "basic":
                    e.p = int(g["reset_pos"][t, i])
                elif task == "ball3d":
                    e.rot, e.pos, e.vel = g["reset_rot"][t, i].astype(np.float32), g["reset_pos"][t, i].copy(), g["reset_vel"][t, i].copy()
                elif task == "gridworld":
                    e.agent, e.green, e.red, e.kind = tuple(g["reset_agent"][t, i]), tuple(g["reset_green"][t, i]), tuple(g["reset_red"][t, i]), int(g["reset_goal_type"][t, i])
                else:
                    e.agent, e.box, e.goal = tuple(g["reset_agent"][t, i]), tuple(g["reset_box"][t, i]), (int(g["reset_goal_x"][t, i]), 5)
