"""TEST INFRASTRUCTURE ONLY — CPU restatement (vectorised NumPy) of the four
reference tasks on the hot path (plus walljump), plus the adapter / DummyVecEnv / Monitor
semantics around them.  The product path never imports this module; only
tests/, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` leg do.

Pinned (tests/test_oracle_cpu.py) against tests/golden/*.npz, which were made by
running the UNMODIFIED reference classes (oracle/make_golden.py): 32 envs x 1000
steps per task with state injection at every reset — bit-exact for all four
tasks on the build container.

Reference anchors (relative to /root/reference/backend):
  basic      mlagents/envs.py:17-84
  ball3d     examples/ball3d.py:10-113
  gridworld  examples/gridworld.py:14-95
  push       examples/push.py:10-125
  walljump   examples/walljump.py:14-98
  brickbreak examples/brick_break.py:11-133
  bicycle    examples/bicycle.py:11-146
  glider     examples/glider.py:11-265
  adapter    mlagents/envs.py:87-159  (time-limit truncation, terminated/truncated split)
  vec/auto-reset + Monitor: SB3 DummyVecEnv/Monitor semantics, SURVEY.md §8(a) A7
Reset draws use this repo's Philox streams (oracle/philox.py), not MT19937.
"""
from __future__ import annotations

import numpy as np

from . import philox as px

# ---- constants -----------------------------------------------------------------------
# ball3d.py:10-37
G = 9.81
DT = 0.02
MAX_TILT = float(np.deg2rad(25.0))      # 0x3fdbecde5da115a9
TILT_DELTA = float(np.deg2rad(3.0))     # 0x3faacee9f37bebd6
PLATFORM_HALF = 3.0
BALL3D_DELTAS = np.array(
    [[TILT_DELTA, 0.0], [-TILT_DELTA, 0.0], [0.0, TILT_DELTA], [0.0, -TILT_DELTA], [0.0, 0.0]],
    dtype=np.float64,
)
# gridworld.py:19-25 and push.py:14-20 share the move table
GRID_DELTAS = np.array([[0, 0], [0, 1], [0, -1], [-1, 0], [1, 0]], dtype=np.int32)

TASKS = {
    #            obs_dim n_actions max_steps
    "basic":     (21, 3, 50),     # envs.py:17-44
    "ball3d":    (6, 5, 200),     # ball3d.py:12,24,38 ; envs.py:169-175
    "gridworld": (4, 5, 100),     # gridworld.py:14-30 ; envs.py:181-187
    "push":      (4, 5, 120),     # push.py:10-24 ; envs.py:193-199
    "walljump":  (4, 4, 150),     # walljump.py:14-21 ; envs.py:202-213
    "brickbreak": (45, 3, 2000),  # brick_break.py:14-38 (2 + 2 + 1 + 5*8 obs) ; envs.py:216-227
    "bicycle":   (7, 3, 2000),    # bicycle.py:130-145 ; envs.py:230-241
    "glider":    (16, 5, 4000),   # glider.py:241-265 (9 + 3 + 3 + 1) ; envs.py:244-255
}

STATE_DTYPES = {   # identical to the C structs in include/tmla.h (wire format of get/set_state)
    "basic": np.dtype([("pos", "<i4"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "ball3d": np.dtype([("rot", "<f8", (2,)), ("pos", "<f4", (2,)), ("vel", "<f4", (2,)),
                        ("steps", "<i4"), ("ep_return", "<f4"), ("episode", "<i4"), ("pad_", "<i4")]),
    "gridworld": np.dtype([("agent", "<i4", (2,)), ("green", "<i4", (2,)), ("red", "<i4", (2,)),
                           ("goal_type", "<i4"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "push": np.dtype([("agent", "<i4", (2,)), ("box", "<i4", (2,)), ("goal_x", "<i4"),
                      ("steps", "<i4"), ("ep_return", "<f4")]),
    "walljump": np.dtype([("agent_x", "<i4"), ("in_air", "<i4"), ("wall", "<i4"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "brickbreak": np.dtype([("pos", "<f8", (2,)), ("vel", "<f8", (2,)), ("paddle", "<f8"), ("bricks", "u1", (40,)),
                            ("steps", "<i4"), ("ep_return", "<f4")]),
    "bicycle": np.dtype([("x", "<f8"), ("z", "<f8"), ("theta", "<f8"), ("phi", "<f8"), ("phi_dot", "<f8"), ("delta", "<f8"),
                         ("goal", "<f8", (2,)), ("dist", "<f8"), ("steps", "<i4"), ("ep_return", "<f4")]),
    "glider": np.dtype([("pos", "<f8", (3,)), ("vel", "<f8", (3,)), ("rot", "<f8", (3,)), ("ang_vel", "<f8", (3,)),
                        ("waypoint", "<i4"), ("steps", "<i4"), ("ep_return", "<f4"), ("pad_", "<i4")]),
}

f32 = np.float32

# bicycle.py:15-31 — the constants as the reference's Python expressions evaluate them
BIKE_DT = 0.02
BIKE_GH = 9.8 / 0.8                  # self.g / self.h
BIKE_VLH = 5.0 ** 2 / (1.0 * 0.8)    # self.v**2 / (self.L * self.h)
BIKE_VL = 5.0 / 1.0                  # self.v / self.L
BIKE_MAX_PHI = np.pi / 4
BIKE_MAX_DELTA = np.pi / 6


# glider.py:16-50 — constants as the reference's Python expressions evaluate them
GL_DT, GL_MASS, GL_G, GL_RHO, GL_S = 0.02, 1.5, 9.81, 1.225, 0.5
GL_CL_ALPHA, GL_CD0, GL_CDK = 2 * np.pi, 0.02, 0.05
GL_WAYPOINTS = np.array([[-160.0, 0.0, 70.0], [160.0, 0.0, 70.0]])
GL_F1, GL_F2, GL_C1, GL_MAG1, GL_MAG2, GL_C3 = 1.0 / 250.0, 1.0 / 400.0, 8.0, 1.0, 0.7, 50.0
GL_MAX_ROLL, GL_MAX_PITCH, GL_MAX_AOA = np.pi / 2, np.pi / 4, np.deg2rad(15)
GL_TORQUES = np.array([[0.0, 0.0, 0.0], [-15.0, 0.0, 4.0], [15.0, 0.0, -4.0], [0.0, 10.0, 0.0], [0.0, -10.0, 0.0]])  # glider.py:92-103


def _glider_step_one(pos, vel, rot, ang_vel, wp, action):
    """One env, one step of glider.py:87-238 on float64 3-vectors, through the same NumPy entry points as the reference
    (np.linalg.norm, np.dot, `@` go to this host's BLAS; np.arctan2 / np.sin / np.cos to its libm or SVML) — the restatement is
    pinned bit for bit on the host that made the fixture and held to a tolerance elsewhere, like the CUDA kernel."""
    t = GL_TORQUES[action]
    ang_vel = ang_vel.copy()
    ang_vel[0] += t[0] * GL_DT                                      # :107-109
    ang_vel[1] += t[1] * GL_DT
    ang_vel[2] += t[2] * GL_DT
    ang_vel = ang_vel * 0.95                                        # :110
    rot = rot + ang_vel * GL_DT                                     # :111
    rot[0] = np.clip(rot[0], -GL_MAX_ROLL, GL_MAX_ROLL)             # :114-115
    rot[1] = np.clip(rot[1], -GL_MAX_PITCH, GL_MAX_PITCH)
    x, y = pos[0], pos[1]                                           # :55-77 wind at the position before integration
    up1 = np.sin(x * GL_F1 * 2 * np.pi) * np.cos(y * GL_F1 * 2 * np.pi) * GL_C1 * GL_MAG1
    up2 = np.sin(x * GL_F2 * 2 * np.pi / 1.5) * np.cos(y * GL_F1 * 2 * np.pi / 1.5) * GL_C1 * GL_MAG2
    v_air = vel - np.array([1.0, 0.5, up1 + up2])                   # :118-119
    v_air_mag = np.linalg.norm(v_air)                               # :120
    aoa = np.arctan2(-v_air[2], v_air[0]) if v_air[0] != 0 else 0   # :123
    if v_air_mag > 0.1:                                             # :125-160
        CL = GL_CL_ALPHA * aoa
        CD = GL_CD0 + GL_CDK * CL**2
        lift = 0.5 * GL_RHO * v_air_mag**2 * GL_S * CL
        drag = 0.5 * GL_RHO * v_air_mag**2 * GL_S * CD
        c0, s0, c1, s1, c2, s2 = np.cos(rot[0]), np.sin(rot[0]), np.cos(rot[1]), np.sin(rot[1]), np.cos(rot[2]), np.sin(rot[2])
        R_roll = np.array([[1, 0, 0], [0, c0, -s0], [0, s0, c0]])
        R_pitch = np.array([[c1, 0, s1], [0, 1, 0], [-s1, 0, c1]])
        R_yaw = np.array([[c2, -s2, 0], [s2, c2, 0], [0, 0, 1]])
        aero = (R_yaw @ R_pitch @ R_roll) @ (np.array([0, 0, lift]) + np.array([-drag, 0, 0]))
    else:
        aero = np.zeros(3)
        aoa = 0
    total = aero + np.array([0, 0, -GL_MASS * GL_G])                # :162-163
    vel = vel + (total / GL_MASS) * GL_DT                           # :166
    pos = pos + vel * GL_DT                                         # :167
    vec = GL_WAYPOINTS[wp] - pos                                    # :173-175
    dist = np.linalg.norm(vec)
    if dist < 15.0:                                                 # :177-180
        wp = (wp + 1) % 2
    vel_dir = vel / (np.linalg.norm(vel) + 1e-8)                    # :183-185
    heading = np.dot(vel_dir, vec / (dist + 1e-8))
    H = (heading + 1) / 2                                           # :188
    E = np.clip(np.linalg.norm(vel) / 30.0, 0, 2.0)                 # :190-192
    reward = E * (H - E + 1)                                        # :196
    lateral = abs(pos[1])                                           # :201-206
    if lateral > 250.0:
        reward -= 2.0 * (((lateral - 250.0) / 100.0) ** 2)
    if pos[2] > 250.0:                                              # :209-215
        reward -= 2.0 * ((pos[2] - 250.0) / 50.0) ** 2
    elif pos[2] < 25.0:
        reward -= 0.5
    done = False
    if pos[2] < 5.0:                                                # :218-220
        reward, done = -50.0, True
    if abs(aoa) > GL_MAX_AOA:                                       # :223-225
        reward, done = -50.0, True
    if dist > 500:                                                  # :228-230
        reward, done = -50.0, True
    return pos, vel, rot, ang_vel, wp, float(reward), done


def dot2(a, b):
    """Row-wise np.dot of [n,2] arrays THROUGH np.dot (the reference's np.linalg.norm / np.dot on 2-vectors go to BLAS ddot,
    which on FMA hosts rounds as fma(a1, b1, a0*b0) — measured; csrc/envs.cuh:BicycleTask uses exactly that)."""
    return np.array([np.dot(a[i], b[i]) for i in range(a.shape[0])], dtype=np.float64).reshape(a.shape[0])


def push_reward_lut():
    """18-entry LUT indexed [(d_ab+1)*6 + (d_bg+1)*2 + invalid], built with the
    reference's own expression order in Python doubles, then rounded once to f32
    (push.py:77,111-115)."""
    lut = np.zeros(18, np.float32)
    for dab in (-1, 0, 1):
        for dbg in (-1, 0, 1):
            for inv in (0, 1):
                r = -0.01
                r += 0.05 * dab
                r += 0.3 * dbg
                if inv:
                    r -= 0.05
                lut[(dab + 1) * 6 + (dbg + 1) * 2 + inv] = np.float32(r)
    return lut


def basic_reward_lut():
    """envs.py:65-72: -0.01, (-0.01)+0.1, (-0.01)+1.0 in doubles -> f32."""
    a = -0.01
    b = -0.01
    b += 0.1
    c = -0.01
    c += 1.0
    return np.array([a, b, c], dtype=np.float32)


def walljump_reward_lut():
    """walljump.py:58,77,81,91: -0.01, (-0.01)-0.02, (-0.01)-0.03 in doubles -> f32; index 3 = goal."""
    a = -0.01
    b = -0.01
    b -= 0.02
    c = -0.01
    c -= 0.03
    return np.array([a, b, c, 1.0], dtype=np.float32)


PUSH_LUT = push_reward_lut()
WALLJUMP_LUT = walljump_reward_lut()
WJ_WIDTH, WJ_WALL_X, WJ_JUMP = 20, 10, 3          # walljump.py:14,34-35
WJ_DELTAS = np.array([0, 1, -1, 1], dtype=np.int32)   # walljump.py:18 (jump also moves forward)
BASIC_LUT = basic_reward_lut()


# ---- observations --------------------------------------------------------------------
def observe(task, st):
    n = st.shape[0]
    if task == "basic":          # envs.py:24-27
        obs = np.zeros((n, 21), np.float32)
        obs[np.arange(n), np.clip(st["pos"], 0, 20)] = 1.0
        return obs
    if task == "ball3d":         # ball3d.py:61-72 : f32 of (rot f64|f32, pos, vel)
        return np.concatenate([st["rot"].astype(np.float32), st["pos"], st["vel"]], axis=1)
    if task == "gridworld":      # gridworld.py:55-64 ; quarters are exact in f32
        goal = np.where(st["goal_type"][:, None] == 0, st["green"], st["red"])
        d = (goal - st["agent"]).astype(np.float64) / 4.0
        oh = np.stack([st["goal_type"] == 0, st["goal_type"] == 1], axis=1).astype(np.float64)
        return np.concatenate([d, oh], axis=1).astype(np.float32)
    if task == "push":           # push.py:53-59 ; f32(k/5.0) == f32(k)/f32(5) for |k|<=5
        goal = np.stack([st["goal_x"], np.full(n, 5, np.int32)], axis=1)
        ab = (st["box"] - st["agent"]).astype(np.float64) / 5.0
        bg = (goal - st["box"]).astype(np.float64) / 5.0
        return np.concatenate([ab, bg], axis=1).astype(np.float32)
    if task == "walljump":       # walljump.py:48-53 (Python double division, then f32)
        x = st["agent_x"].astype(np.float64)
        return np.stack([(WJ_WIDTH - 1 - x) / (WJ_WIDTH - 1), (WJ_WALL_X - x) / (WJ_WIDTH - 1), st["wall"].astype(np.float64),
                         (st["in_air"] == 0).astype(np.float64)], axis=1).astype(np.float32)
    if task == "glider":         # glider.py:241-265 (the target direction uses the waypoint index AFTER a switch)
        out = np.zeros((n, 16), np.float64)
        for i in range(n):
            pos, vel, rot, av = st["pos"][i], st["vel"][i], st["rot"][i], st["ang_vel"][i]
            vec = GL_WAYPOINTS[st["waypoint"][i]] - pos
            dist = np.linalg.norm(vec)
            out[i] = np.concatenate([np.array([vel[2] / 10.0, (pos[2] - GL_C3) / 50.0, rot[0], rot[1], np.sin(rot[2]), np.cos(rot[2]),
                                               av[0], av[1], av[2]]), vel / 20.0, vec / (dist + 1e-8), [dist / 100.0]])
        return out.astype(np.float32)
    if task == "bicycle":        # bicycle.py:128-145: the goal direction is re-derived from the state; the adapter casts to f32
        vec = st["goal"] - np.stack([st["x"], st["z"]], 1)
        dist = np.sqrt(dot2(vec, vec))
        nv = np.where(dist[:, None] > 0, vec / np.where(dist > 0, dist, 1.0)[:, None], 0.0)
        return np.stack([st["phi"], st["phi_dot"], st["delta"], np.cos(st["theta"]), np.sin(st["theta"]), nv[:, 0], nv[:, 1]],
                        axis=1).astype(np.float32)
    if task == "brickbreak":     # brick_break.py:118-126: f64 concatenate, the adapter casts to f32 (envs.py:150)
        return np.concatenate([st["pos"] / np.array([40.0, 40.0]), st["vel"], (st["paddle"] / 40.0)[:, None],
                               st["bricks"].astype(np.float64)], axis=1).astype(np.float32)
    raise KeyError(task)


# ---- deterministic sin/cos on [-pi/4, pi/4] for the brickbreak serve angle: Horner in plain f64 mul/add (no fma), the
# ---- same operation order as csrc/envs.cuh so that oracle and device resets agree bit for bit
_SIN_C = (-1.0 / 6, 1.0 / 120, -1.0 / 5040, 1.0 / 362880, -1.0 / 39916800, 1.0 / 6227020800, -1.0 / 1307674368000)
_COS_C = (-1.0 / 2, 1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600, -1.0 / 87178291200, 1.0 / 20922789888000)


def sin_cos_quarter(y):
    y = np.asarray(y, dtype=np.float64)
    z = y * y
    ps = np.full_like(z, _SIN_C[-1])
    for c in _SIN_C[-2::-1]:
        ps = ps * z + c
    pc = np.full_like(z, _COS_C[-1])
    for c in _COS_C[-2::-1]:
        pc = pc * z + c
    return y + (y * z) * ps, 1.0 + z * pc


# ---- transitions (no reset) -------------------------------------------------------------
def transition(task, st, actions):
    """Advance `st` (structured array, modified in place) by one step.
    Returns (obs, reward f32, terminated bool, truncated bool) as the adapter reports
    them (envs.py:139-152; for basic envs.py:60-81) — BEFORE any auto-reset."""
    a = np.asarray(actions).astype(np.int64)
    n = st.shape[0]
    max_steps = TASKS[task][2]
    if task == "basic":
        pos = np.clip(st["pos"] + (a - 1), 0, 20)                  # envs.py:61-62
        st["pos"] = pos
        st["steps"] += 1
        small, large = pos == 7, pos == 17                          # envs.py:67-72
        reward = np.where(small, BASIC_LUT[1], np.where(large, BASIC_LUT[2], BASIC_LUT[0])).astype(np.float32)
        terminated = small | large
        truncated = (st["steps"] >= max_steps) & ~terminated        # envs.py:74
    elif task == "ball3d":
        first = st["steps"] == 0
        rot = st["rot"] + BALL3D_DELTAS[a]                          # ball3d.py:76-77 (f64 add)
        # first step after a reset: rot is still the f32 array, `+=` rounds back to f32
        rot = np.where(first[:, None], rot.astype(np.float32).astype(np.float64), rot)
        rot = np.minimum(np.maximum(rot, -MAX_TILT), MAX_TILT)      # ball3d.py:78 -> f64 from here on
        acc = G * np.sin(rot)                                       # ball3d.py:81-82 (f64)
        vel = (st["vel"].astype(np.float64) + acc * DT).astype(np.float32)   # ball3d.py:83-84
        vel = vel * f32(0.98)                                       # ball3d.py:87 (f32)
        pos = st["pos"] + vel * f32(DT)                             # ball3d.py:90 (f32, no fma)
        st["rot"], st["vel"], st["pos"] = rot, vel, pos
        st["steps"] += 1
        off = (np.abs(pos[:, 0]) > f32(3.0)) | (np.abs(pos[:, 1]) > f32(3.0))   # ball3d.py:96-98
        timeout = st["steps"] >= 200                                # ball3d.py:99
        done = off | timeout
        sq = pos[:, 0] * pos[:, 0] + pos[:, 1] * pos[:, 1]          # np.linalg.norm: sqrt(x.dot(x)), f32
        d = np.sqrt(sq).astype(np.float32)
        reward = f32(1.0) - d / f32(3.0)                            # ball3d.py:104
        reward = np.where(done, np.where(timeout & ~off, f32(1.0), f32(-1.0)), reward)  # :105-108
        reward = (reward + f32(-0.02) * d).astype(np.float32)       # :110-111
        hit = st["steps"] >= max_steps                              # envs.py:141-145
        terminated, truncated = done & ~hit, hit
    elif task == "gridworld":
        agent = np.clip(st["agent"] + GRID_DELTAS[a], 0, 4)         # gridworld.py:68-71
        st["agent"] = agent
        st["steps"] += 1
        on_green = np.all(agent == st["green"], axis=1)             # gridworld.py:79-90
        on_red = np.all(agent == st["red"], axis=1) & ~on_green
        gt = st["goal_type"]
        reward = np.full(n, f32(-0.01), np.float32)
        reward = np.where(on_green, np.where(gt == 0, f32(1.0), f32(-1.0)), reward)
        reward = np.where(on_red, np.where(gt == 1, f32(1.0), f32(-1.0)), reward).astype(np.float32)
        done = on_green | on_red | (st["steps"] >= 100)             # gridworld.py:92-93
        hit = st["steps"] >= max_steps
        terminated, truncated = done & ~hit, hit
    elif task == "push":
        d = GRID_DELTAS[a]
        agent0, box0 = st["agent"].copy(), st["box"].copy()
        goal = np.stack([st["goal_x"], np.full(n, 5, np.int32)], axis=1)
        new_agent = np.clip(agent0 + d, 0, 5)                       # push.py:63-65
        prev_bg = np.abs(goal - box0).sum(1)                        # push.py:70-75
        prev_ab = np.abs(box0 - agent0).sum(1)
        into_box = np.all(new_agent == box0, axis=1)                # push.py:83
        tent = box0 + d
        inb = np.all((tent >= 0) & (tent < 6), axis=1)              # push.py:87-90
        new_box = np.where((into_box & inb)[:, None], tent, box0)
        invalid = into_box & ~inb                                   # push.py:92-95
        new_agent = np.where(invalid[:, None], agent0, new_agent)
        st["agent"], st["box"] = new_agent, new_box
        st["steps"] += 1
        dist_bg = np.abs(goal - new_box).sum(1)                     # push.py:103-108
        dist_ab = np.abs(new_box - new_agent).sum(1)
        idx = (prev_ab - dist_ab + 1) * 6 + (prev_bg - dist_bg + 1) * 2 + invalid.astype(np.int64)
        reward = PUSH_LUT[idx]
        top = new_box[:, 1] == 5                                    # push.py:118-120
        reward = np.where(top, f32(1.0), reward).astype(np.float32)
        done = top | (st["steps"] >= 120)                           # push.py:122-123
        hit = st["steps"] >= max_steps
        terminated, truncated = done & ~hit, hit
    elif task == "walljump":
        x0, air = st["agent_x"].copy(), st["in_air"].copy()
        just_jumped = (a == 3) & (air == 0)                          # walljump.py:62-65
        air = np.where(just_jumped, WJ_JUMP, air)
        px_ = np.clip(x0 + WJ_DELTAS[a], 0, WJ_WIDTH - 1)            # walljump.py:68-69
        crossing = ((x0 < WJ_WALL_X) & (WJ_WALL_X <= px_)) | ((px_ < WJ_WALL_X) & (WJ_WALL_X <= x0))   # :72-74
        blocked = crossing & (st["wall"] == 1) & (air == 0)          # walljump.py:75-77
        px_ = np.where(blocked, x0, px_)
        ridx = np.where(blocked, 1, 0)
        ridx = np.where(just_jumped & ~crossing & (np.abs(WJ_WALL_X - x0) > 1), 2, ridx)   # walljump.py:80-81 (never both)
        air = np.where(air > 0, air - 1, air)                        # walljump.py:86-87
        goal = px_ == WJ_WIDTH - 1                                   # walljump.py:90-92
        ridx = np.where(goal, 3, ridx)
        st["agent_x"], st["in_air"] = px_, air
        st["steps"] += 1
        reward = WALLJUMP_LUT[ridx]
        done = goal | (st["steps"] >= 150)                           # walljump.py:94-96
        hit = st["steps"] >= max_steps
        terminated, truncated = done & ~hit, hit
    elif task == "brickbreak":
        paddle = st["paddle"] + np.where(a == 0, -3.0, np.where(a == 2, 3.0, 0.0))       # brick_break.py:51-54
        paddle = np.minimum(np.maximum(paddle, 4.0), 36.0)                                # :56-58
        pos = st["pos"] + st["vel"]                                                        # :61
        vel = st["vel"].copy()
        wall_x = (pos[:, 0] <= 1.0) | (pos[:, 0] >= 39.0)                                  # :67-73
        vel[:, 0] = np.where(wall_x, -vel[:, 0], vel[:, 0])
        vel[:, 1] = np.where(pos[:, 1] >= 39.0, -vel[:, 1], vel[:, 1])
        reward = np.zeros(n, np.float64)
        hit = (vel[:, 1] < 0) & (pos[:, 1] - 1.0 <= 2.0) & (pos[:, 0] >= paddle - 4.0) & (pos[:, 0] <= paddle + 4.0)   # :76-81
        vel[:, 1] = np.where(hit, -vel[:, 1], vel[:, 1])
        offset = (pos[:, 0] - paddle) / 4.0                                                # :83
        vel[:, 0] = np.where(hit, vel[:, 0] + offset * 0.5, vel[:, 0])
        reward = np.where(hit, 0.1, reward)
        bricks = st["bricks"].copy()
        found = np.zeros(n, bool)
        for r in range(5):                                                                  # :88-105, first live brick in row-major order
            for c in range(8):
                bx, by = c * 5.0, 20.0 + r * 2.0
                cond = (~found & (bricks[:, r * 8 + c] == 1) & (pos[:, 0] >= bx) & (pos[:, 0] <= bx + 5.0)
                        & (pos[:, 1] >= by) & (pos[:, 1] <= by + 2.0))
                bricks[:, r * 8 + c] = np.where(cond, 0, bricks[:, r * 8 + c])
                vel[:, 1] = np.where(cond, -vel[:, 1], vel[:, 1])
                found |= cond
        reward = np.where(found, 1.0, reward)
        lost = pos[:, 1] < 1.0                                                              # :109-111
        reward = np.where(lost, -1.0, reward)
        cleared = bricks.sum(1) == 0                                                        # :113-115
        reward = np.where(cleared, 10.0, reward)
        st["pos"], st["vel"], st["paddle"], st["bricks"] = pos, vel, paddle, bricks
        st["steps"] += 1
        done = lost | cleared | (st["steps"] > 2000)                                        # :117-118
        reward = reward.astype(np.float32)
        hit_limit = st["steps"] >= max_steps
        terminated, truncated = done & ~hit_limit, hit_limit
    elif task == "glider":
        reward = np.zeros(n, np.float64)
        done = np.zeros(n, bool)
        for i in range(n):
            pos, vel, rot, av, wp, reward[i], done[i] = _glider_step_one(st["pos"][i], st["vel"][i], st["rot"][i], st["ang_vel"][i],
                                                                         int(st["waypoint"][i]), int(a[i]))
            st["pos"][i], st["vel"][i], st["rot"][i], st["ang_vel"][i], st["waypoint"][i] = pos, vel, rot, av, wp
        st["steps"] += 1
        done |= st["steps"] > 4000                                                                 # glider.py:233-234
        reward = reward.astype(np.float32)
        hit_limit = st["steps"] >= max_steps
        terminated, truncated = done & ~hit_limit, hit_limit
    elif task == "bicycle":
        # bicycle.py:59-126 in the reference's operation order (Python doubles).  np.sin/np.cos are libm; np.tan and `** 0.5`
        # (libm pow, not sqrt) are whatever this host's NumPy dispatches to — the CUDA side is compared within a tolerance.
        delta = st["delta"] + np.where(a == 0, -0.05, np.where(a == 2, 0.05, 0.0))                # :63-69
        delta = np.minimum(np.maximum(delta, -BIKE_MAX_DELTA), BIKE_MAX_DELTA)                     # :70
        phi_ddot = BIKE_GH * np.sin(st["phi"]) - (BIKE_VLH * np.tan(delta)) * np.cos(st["phi"])   # :74-76
        phi_dot = st["phi_dot"] + phi_ddot * BIKE_DT                                               # :77
        phi = st["phi"] + phi_dot * BIKE_DT                                                        # :78
        delta = delta * 0.95                                                                       # :81
        theta = st["theta"] + (BIKE_VL * np.tan(delta)) * BIKE_DT                                  # :84
        x = st["x"] + (5.0 * np.cos(theta)) * BIKE_DT                                              # :85
        z = st["z"] + (5.0 * np.sin(theta)) * BIKE_DT                                              # :86
        gv = st["goal"] - np.stack([x, z], 1)
        new_dist = np.sqrt(dot2(gv, gv))                                                           # :91 np.linalg.norm
        progress = (st["dist"] - new_dist) * 10.0                                                  # :94
        upright = (1.0 - np.power(np.abs(phi) / BIKE_MAX_PHI, 0.5)) * 0.2                          # :98
        hv = np.stack([np.cos(theta), np.sin(theta)], 1)                                           # :101
        ngv = gv / np.where(new_dist > 0, new_dist, 1.0)[:, None]                                  # :103-105
        heading = dot2(hv, ngv) * 0.3                                                              # :106
        steering = -(np.abs(delta) / BIKE_MAX_DELTA) * 0.1                                         # :109
        reward = progress + upright + heading + steering                                           # :111
        st["x"], st["z"], st["theta"], st["phi"], st["phi_dot"], st["delta"], st["dist"] = x, z, theta, phi, phi_dot, delta, new_dist
        st["steps"] += 1
        fell = np.abs(phi) > BIKE_MAX_PHI                                                          # :113-115
        reward = np.where(fell, -10.0, reward)
        reached = new_dist < 2.0                                                                   # :120-122
        reward = np.where(reached, 50.0, reward)
        done = fell | reached | (st["steps"] > 2000)                                               # :117-118
        reward = reward.astype(np.float32)
        hit_limit = st["steps"] >= max_steps
        terminated, truncated = done & ~hit_limit, hit_limit
    else:
        raise KeyError(task)
    return observe(task, st), reward, terminated, truncated


# ---- Philox resets (this repo's stream layout; distributions follow the reference) -----------
def draw_reset(task, seed, env_ids, k, tag=px.TAG_RESET, episode=None):
    """Fresh episode state for global env ids `env_ids` at global step index `k`.
    ball3d auto-resets (tag TAG_RESET) are indexed by the env's own episode counter instead of `k`
    (`episode` = index of the episode being started), see csrc/envs.cuh:Ball3DTask::draw."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    n = env_ids.shape[0]
    st = np.zeros(n, STATE_DTYPES[task])
    if task == "basic":
        st["pos"] = 10                                              # envs.py:21,55
    elif task == "ball3d":                                          # ball3d.py:49-57
        # lo + (hi-lo)*u like np.random.uniform, u = (w + 0.5) * 2^-32 carries 32 random bits
        ctr = k
        if tag == px.TAG_RESET:
            ctr = np.asarray(episode, dtype=np.uint64)
            st["episode"] = ctr.astype(np.int32)
        b0 = px.stream_block(seed, env_ids, ctr, tag, 0)
        b1 = px.stream_block(seed, env_ids, ctr, tag, 1)
        lo = -MAX_TILT * 0.5
        rot = np.stack([lo + MAX_TILT * px.u32_unit(b0[0]), lo + MAX_TILT * px.u32_unit(b0[1])], 1)
        pos = np.stack([-1.5 + 3.0 * px.u32_unit(b0[2]), -1.5 + 3.0 * px.u32_unit(b0[3])], 1)
        vel = np.stack([-1.0 + 2.0 * px.u32_unit(b1[0]), -1.0 + 2.0 * px.u32_unit(b1[1])], 1)
        st["rot"] = rot.astype(np.float32).astype(np.float64)       # `.astype(np.float32)` at reset
        st["pos"] = pos.astype(np.float32)
        st["vel"] = vel.astype(np.float32)
    elif task == "gridworld":                                       # gridworld.py:42-50
        b = px.stream_block(seed, env_ids, k, tag, 0)
        a = px.bounded(b[0], 25)
        g = px.bounded(b[1], 24)
        g = g + (g >= a)
        r = px.bounded(b[2], 23)
        lo, hi = np.minimum(a, g), np.maximum(a, g)
        r = r + (r >= lo)
        r = r + (r >= hi)
        st["agent"] = np.stack([a // 5, a % 5], 1)
        st["green"] = np.stack([g // 5, g % 5], 1)
        st["red"] = np.stack([r // 5, r % 5], 1)
        st["goal_type"] = (b[3] >> np.uint32(31)).astype(np.int32)
    elif task == "push":                                            # push.py:40-47
        b = px.stream_block(seed, env_ids, k, tag, 0)
        a = px.bounded(b[0], 36)
        bx = px.bounded(b[1], 35)
        bx = bx + (bx >= a)
        st["agent"] = np.stack([a // 6, a % 6], 1)
        st["box"] = np.stack([bx // 6, bx % 6], 1)
        st["goal_x"] = px.bounded(b[2], 6)
    elif task == "brickbreak":                                      # brick_break.py:39-46
        b = px.stream_block(seed, env_ids, k, tag, 0)
        angle = np.pi / 4 + (np.pi / 2) * px.u32_unit(b[0])         # np.random.uniform(pi/4, 3pi/4) = lo + (hi - lo) * u
        s_, c_ = sin_cos_quarter(angle - np.pi / 2)                 # cos(angle) = -sin(angle - pi/2), sin(angle) = cos(angle - pi/2)
        st["pos"] = np.array([20.0, 10.0])
        st["vel"] = np.stack([-s_ * 1.5, c_ * 1.5], 1)
        st["paddle"] = 20.0
        st["bricks"] = 1
    elif task == "glider":                                          # glider.py:79-86
        b = px.stream_block(seed, env_ids, k, tag, 0)
        st["pos"] = np.array([0.0, 0.0, 60.0])
        st["vel"] = np.array([15.0, 0.0, -1.0])
        st["ang_vel"] = np.stack([-0.1 + 0.2 * px.u32_unit(b[j]) for j in range(3)], 1)   # np.random.uniform(-0.1, 0.1, 3)
        st["waypoint"] = (b[3] >> np.uint32(31)).astype(np.int32)   # np.random.randint(0, 2)
    elif task == "bicycle":                                         # bicycle.py:40-58
        b = px.stream_block(seed, env_ids, k, tag, 0)
        st["phi"] = -0.1 + 0.2 * px.u32_unit(b[0])                  # np.random.uniform(lo, hi) = lo + (hi - lo) * u
        st["phi_dot"] = -0.1 + 0.2 * px.u32_unit(b[1])
        radius = 15.0 + 10.0 * px.u32_unit(b[2])
        s_, c_ = sin_cos_quarter(-np.pi / 4 + (np.pi / 2) * px.u32_unit(b[3]))
        st["goal"] = np.stack([radius * c_, radius * s_], 1)
        st["dist"] = np.sqrt(dot2(st["goal"], st["goal"]))          # np.linalg.norm(goal - [0, 0])
    elif task == "walljump":                                        # walljump.py:39-45: int(np.random.rand() < 0.7)
        b = px.stream_block(seed, env_ids, k, tag, 0)
        u24 = (b[0] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        st["wall"] = (u24 < np.float32(0.7)).astype(np.int32)
    return st


def random_actions(task, seed, env_ids, step_index):
    """Random-policy action of every env at global step `step_index` (TAG_ACTION stream): Philox block
    step>>4, word (step&15)>>2 read as a 32-bit fraction, base-A digit number step&3 (csrc/philox.cuh)."""
    n_act = np.uint64(TASKS[task][1])
    b = px.stream_block(seed, env_ids, step_index >> 4, px.TAG_ACTION, 0)
    frac = b[(step_index & 15) >> 2].astype(np.uint64)
    a = None
    for _ in range((step_index & 3) + 1):
        p = frac * n_act
        a = (p >> np.uint64(32)).astype(np.int32)
        frac = p & np.uint64(0xFFFFFFFF)
    return a


class OracleVecEnv:
    """DummyVecEnv[Monitor[adapter]]-equivalent over the vectorised transitions above
    with on-the-spot auto-reset; mirrors exactly what one `tmla_step` launch does."""

    def __init__(self, task, n_envs, seed=1, env_id_base=0):
        self.task, self.n, self.seed = task, int(n_envs), int(seed)
        self.obs_dim, self.n_actions, self.max_steps = TASKS[task]
        self.env_ids = np.arange(env_id_base, env_id_base + n_envs, dtype=np.uint64)
        self.step_count = 0
        self.state = np.zeros(n_envs, STATE_DTYPES[task])
        self.reset()

    def reset(self):
        self.state = draw_reset(self.task, self.seed, self.env_ids, self.step_count, px.TAG_RESET_ALL)
        return observe(self.task, self.state)

    def step(self, actions):
        st = self.state
        obs, reward, terminated, truncated = transition(self.task, st, actions)
        st["ep_return"] = (st["ep_return"] + reward).astype(np.float32)   # Monitor (f32 accumulator here)
        self.step_count += 1
        done = terminated | truncated
        info = {
            "terminal_obs": obs.copy(),
            "episode_return": st["ep_return"].copy(),
            "episode_length": st["steps"].copy(),
            "terminated": terminated,
        }
        if done.any():
            idx = np.nonzero(done)[0]
            nxt = (st["episode"][idx] + 1) & 0xFFFFFF if self.task == "ball3d" else None
            fresh = draw_reset(self.task, self.seed, self.env_ids[idx], self.step_count, px.TAG_RESET, episode=nxt)
            st[idx] = fresh
            obs[idx] = observe(self.task, fresh)
        # SB3: infos[i]["TimeLimit.truncated"] = truncated and not terminated
        return obs, reward, done, truncated & ~terminated, info
