"""TEST INFRASTRUCTURE ONLY — scalar, object-per-environment port of the reference's CPU path, used
as the timed CPU baseline (bench.py `cpu_baseline` / `--impl reference`) because the reference itself
(Python under /root/reference) cannot travel to the GPU box.

It keeps the reference's COST STRUCTURE on purpose: one Python object per environment, small NumPy
arrays and per-step `np.clip` / `np.linalg.norm` / `np.array` calls for ball3d
(examples/ball3d.py:74-113), tuple arithmetic for the grid tasks (gridworld.py:67-95,
push.py:62-125), the adapter's truncation logic (mlagents/envs.py:125-152) and a serial
DummyVecEnv-style loop with auto-reset (SB3 `DummyVecEnv.step_wait`, reached from
training.py:89).  Results are checked against tests/golden in tests/test_oracle_cpu.py; resets use
NumPy's global RNG exactly like the reference (np.random.*), so this port is for timing and
transition parity, not for reproducing this repo's Philox streams.
"""
from __future__ import annotations

import numpy as np

_TILT_MAX = np.deg2rad(25.0)
_TILT_STEP = np.deg2rad(3.0)
_BALL_MOVES = [np.array(v) for v in ([_TILT_STEP, 0.0], [-_TILT_STEP, 0.0], [0.0, _TILT_STEP], [0.0, -_TILT_STEP], [0.0, 0.0])]
_GRID_MOVES = [(0, 0), (0, 1), (0, -1), (-1, 0), (1, 0)]


class Ball3D:
    limit, n_actions = 200, 5

    def __init__(self):
        self.reset()

    def reset(self):
        u = np.random.uniform
        self.rot = u(-_TILT_MAX * 0.5, _TILT_MAX * 0.5, size=2).astype(np.float32)
        self.pos = u(-1.5, 1.5, size=2).astype(np.float32)
        self.vel = u(-1.0, 1.0, size=2).astype(np.float32)
        self.t = 0
        return self.obs()

    def obs(self):
        return np.array([self.rot[0], self.rot[1], self.pos[0], self.pos[1], self.vel[0], self.vel[1]], dtype=np.float32)

    def step(self, a):
        self.rot += _BALL_MOVES[a]
        self.rot = np.clip(self.rot, -_TILT_MAX, _TILT_MAX)
        self.vel[0] += 9.81 * np.sin(self.rot[0]) * 0.02
        self.vel[1] += 9.81 * np.sin(self.rot[1]) * 0.02
        self.vel *= 0.98
        self.pos += self.vel * 0.02
        self.t += 1
        fell = abs(self.pos[0]) > 3.0 or abs(self.pos[1]) > 3.0
        late = self.t >= 200
        dist = np.linalg.norm(self.pos)
        r = 1.0 - dist / 3.0
        if fell or late:
            r = 1.0 if (late and not fell) else -1.0
        r += -0.02 * np.linalg.norm(self.pos)
        return self.obs(), r, fell or late


class GridWorld:
    limit, n_actions = 100, 5

    def __init__(self):
        self.reset()

    def reset(self):
        cells = [(x, y) for x in range(5) for y in range(5)]
        np.random.shuffle(cells)
        self.agent, self.green, self.red = cells[0], cells[1], cells[2]
        self.kind = np.random.choice([0, 1])
        self.t = 0
        return self.obs()

    def obs(self):
        g = self.green if self.kind == 0 else self.red
        return np.array([(g[0] - self.agent[0]) / 4, (g[1] - self.agent[1]) / 4,
                         1.0 if self.kind == 0 else 0.0, 0.0 if self.kind == 0 else 1.0], dtype=np.float32)

    def step(self, a):
        dx, dy = _GRID_MOVES[a]
        self.agent = (int(np.clip(self.agent[0] + dx, 0, 4)), int(np.clip(self.agent[1] + dy, 0, 4)))
        self.t += 1
        r, done = -0.01, False
        if self.agent == self.green:
            r, done = (1.0 if self.kind == 0 else -1.0), True
        elif self.agent == self.red:
            r, done = (1.0 if self.kind == 1 else -1.0), True
        return self.obs(), r, done or self.t >= 100


class Push:
    limit, n_actions = 120, 5

    def __init__(self):
        self.reset()

    def reset(self):
        cells = [(x, y) for x in range(6) for y in range(6)]
        np.random.shuffle(cells)
        self.agent, self.box = cells[0], cells[1]
        self.goal = (np.random.randint(0, 6), 5)
        self.t = 0
        return self.obs()

    def obs(self):
        return np.array([(self.box[0] - self.agent[0]) / 5, (self.box[1] - self.agent[1]) / 5,
                         (self.goal[0] - self.box[0]) / 5, (self.goal[1] - self.box[1]) / 5], dtype=np.float32)

    @staticmethod
    def _l1(p, q):
        return abs(p[0] - q[0]) + abs(p[1] - q[1])

    def step(self, a):
        dx, dy = _GRID_MOVES[a]
        nxt = (int(np.clip(self.agent[0] + dx, 0, 5)), int(np.clip(self.agent[1] + dy, 0, 5)))
        d_bg, d_ab = self._l1(self.goal, self.box), self._l1(self.box, self.agent)
        box, bad = self.box, False
        if nxt == self.box:
            cand = (self.box[0] + dx, self.box[1] + dy)
            if 0 <= cand[0] < 6 and 0 <= cand[1] < 6:
                box = cand
            else:
                nxt, bad = self.agent, True
        self.agent, self.box = nxt, box
        self.t += 1
        r = -0.01
        r += 0.05 * (d_ab - self._l1(self.box, self.agent))
        r += 0.3 * (d_bg - self._l1(self.goal, self.box))
        if bad:
            r -= 0.05
        done = False
        if self.box[1] == 5:
            r, done = 1.0, True
        return self.obs(), r, done or self.t >= 120


class Basic:
    limit, n_actions = 50, 3

    def __init__(self):
        self.reset()

    def reset(self):
        self.p, self.t = 10, 0
        return self.obs()

    def obs(self):
        o = np.zeros(21, dtype=np.float32)
        o[int(np.clip(self.p, 0, 20))] = 1.0
        return o

    def step(self, a):
        self.p = int(np.clip(self.p + (-1, 0, 1)[a], 0, 20))
        self.t += 1
        r, done = -0.01, False
        if self.p == 7:
            r, done = r + 0.1, True
        elif self.p == 17:
            r, done = r + 1.0, True
        return self.obs(), r, done


TASKS = {"basic": Basic, "ball3d": Ball3D, "gridworld": GridWorld, "push": Push}


class SerialVecEnv:
    """DummyVecEnv[adapter(env)] equivalent: serial loop, fresh object + reset on done (envs.py:120-121)."""

    def __init__(self, task: str, n_envs: int, seed: int = 1):
        self.cls = TASKS[task]
        self.n = n_envs
        np.random.seed(seed)
        self.envs = [self.cls() for _ in range(n_envs)]
        self.steps = [0] * n_envs
        d = len(self.envs[0].obs())
        self.buf_obs = np.zeros((n_envs, d), np.float32)
        self.buf_rew = np.zeros(n_envs, np.float32)
        self.buf_done = np.zeros(n_envs, bool)

    def reset(self):
        for i in range(self.n):
            self.envs[i] = self.cls()
            self.buf_obs[i] = self.envs[i].reset()
            self.steps[i] = 0
        return self.buf_obs.copy()

    def step(self, actions):
        infos = []
        for i in range(self.n):
            env = self.envs[i]
            obs, r, done = env.step(int(actions[i]))
            self.steps[i] += 1
            hit = self.steps[i] >= env.limit
            if self.cls is Basic:
                terminated, truncated = done, hit and not done
            else:
                terminated, truncated = bool(done and not hit), bool(hit)
            info = {"steps": self.steps[i]}
            if terminated or truncated:
                info["terminal_observation"] = obs
                info["TimeLimit.truncated"] = truncated and not terminated
                self.envs[i] = self.cls()          # adapter.reset(): ctor (draw 1) + reset() (draw 2)
                obs = self.envs[i].reset()
                self.steps[i] = 0
            self.buf_obs[i] = obs
            self.buf_rew[i] = float(r)
            self.buf_done[i] = terminated or truncated
            infos.append(info)
        return self.buf_obs.copy(), self.buf_rew.copy(), self.buf_done.copy(), infos


def time_random_policy(task: str, n_envs: int, vec_steps: int, seed: int = 1) -> float:
    """Seconds to run `vec_steps` serial vec-steps with uniform random actions."""
    import time

    env = SerialVecEnv(task, n_envs, seed)
    env.reset()
    rng = np.random.default_rng(seed)
    acts = rng.integers(0, env.cls.n_actions, size=(vec_steps, n_envs))
    t0 = time.perf_counter()
    for t in range(vec_steps):
        env.step(acts[t])
    return time.perf_counter() - t0
