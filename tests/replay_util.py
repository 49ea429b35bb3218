"""Golden-trace replay shared by the CPU (oracle) and GPU (CUDA) parity tests.

A trace (tests/golden/<task>.npz, made by oracle/make_golden.py from the unmodified
reference) holds, for E envs x T steps: the injected initial state, the actions, and
the reference's obs / reward / terminated / truncated plus the state the reference
reset to after every finished episode.  `replay` drives any backend exposing
get_state()/set_state()/step_noreset() through the same actions, re-injecting the
reference's reset states, and returns what the backend produced.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATE_KEYS = {
    "basic": ("pos",),
    "ball3d": ("rot", "pos", "vel"),
    "gridworld": ("agent", "green", "red", "goal_type"),
    "push": ("agent", "box", "goal_x"),
    "walljump": ("agent_x", "in_air", "wall"),
    "brickbreak": ("pos", "vel", "paddle", "bricks"),
    "bicycle": ("x", "z", "theta", "phi", "phi_dot", "delta", "goal", "dist"),
    "glider": ("pos", "vel", "rot", "ang_vel", "waypoint"),
}
# Tasks whose reference arithmetic goes through host-dependent libm / SVML / BLAS routines (np.tan, pow, ddot): compared
# within a stated tolerance instead of bit for bit.  obs: 2e-6 absolute (|obs| <= ~8, one f32 ulp is <= 4.8e-7);
# reward: 1e-5 absolute (|reward| <= 50).
LIBM_TASKS = {"bicycle": {"obs_atol": 2e-6, "reward_atol": 1e-5}, "glider": {"obs_atol": 2e-6, "reward_atol": 1e-5}}


def load(task):
    return np.load(os.path.join(GOLDEN, f"{task}.npz"))


def initial_state(task, g, dtype):
    E = g["actions"].shape[1]
    st = np.zeros(E, dtype)
    for k in STATE_KEYS[task]:
        st[k] = g[f"init_{k}"]
    return st


def inject_step_state(task, g, t, st, done):
    """Traces with per-step reference states (glider): overwrite the state of the envs that are still in their episode with
    the reference's state after step t, so that every step is compared from identical inputs (one-step parity)."""
    if f"step_{STATE_KEYS[task][0]}" not in g:
        return st
    live = np.nonzero(~done)[0]
    for k in STATE_KEYS[task]:
        st[k][live] = g[f"step_{k}"][t][live]
    return st


def inject_resets(task, g, t, st, done):
    idx = np.nonzero(done)[0]
    if idx.size == 0:
        return st
    for k in STATE_KEYS[task]:
        st[k][idx] = g[f"reset_{k}"][t][idx]
    st["steps"][idx] = 0
    st["ep_return"][idx] = 0.0
    return st
