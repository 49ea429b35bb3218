"""TEST INFRASTRUCTURE ONLY — minimal stand-in for the `gymnasium` package.

gymnasium is not installed in this image and there is no network.  This shim
provides just enough surface (`Env`, `spaces.Space/Box/Discrete`) for the
reference's `backend/mlagents/envs.py` and `registry.py` to be imported
UNCHANGED from /root/reference by `oracle/make_golden.py`.  It is never on the
product path and never imported by the package.
"""
from . import spaces  # noqa: F401

__version__ = "0.0-shim"


class Env:
    metadata: dict = {}
    observation_space = None
    action_space = None

    def reset(self, *, seed=None, options=None):
        # gymnasium.Env.reset only seeds its private np_random; the reference
        # adapters never read it (envs.py:116-119 seeds numpy's global RNG itself).
        return None

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass
