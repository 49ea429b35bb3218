"""Host logic of CudaVecEnv that needs no GPU: the lazily sorted `infos` over the compact episode-end records
({env index, ep_return, ep_length, terminal_obs[D], padding} as the step kernel writes them, in arbitrary order)."""
import numpy as np

from three_mlagents_b200.vec_env import LazyInfos


def _records(idx, d, stride, rng):
    rec = np.zeros((len(idx), stride), np.float32)
    rec.view(np.int32)[:, 0] = idx
    rec[:, 1] = rng.normal(size=len(idx))                       # episode returns
    rec.view(np.int32)[:, 2] = rng.integers(1, 200, len(idx))   # episode lengths
    rec[:, 3:3 + d] = rng.normal(size=(len(idx), d))            # terminal observations
    return rec


def test_lazy_infos_sorts_on_first_access_and_matches_dense_semantics():
    rng = np.random.default_rng(0)
    n, d, stride = 50, 6, 12                                    # ball3d: 3 + 6 words padded to 12
    idx = rng.permutation(n)[:9]                                # unordered, as atomicAdd slots come out
    rec = _records(idx, d, stride, rng)
    done = np.zeros(n, bool); done[idx] = True
    trunc = np.zeros(n, bool); trunc[idx[:3]] = True
    infos = LazyInfos(n, done, trunc, rec[:, :3 + d], 1.25)     # the view CudaVecEnv passes (padding cut off)
    assert len(infos) == n and infos._idx is None               # nothing sorted yet
    i = int(idx[4])
    info = infos[i]
    assert np.array_equal(info["terminal_observation"], rec[4, 3:3 + d])
    assert info["episode"] == {"r": round(float(rec[4, 1]), 6), "l": int(rec.view(np.int32)[4, 2]), "t": 1.25}
    assert info["steps"] == info["episode"]["l"] and info["TimeLimit.truncated"] == bool(trunc[i])
    assert np.array_equal(infos.finished(), np.sort(idx))
    ret, length = infos.episode_stats()
    order = np.argsort(idx)
    assert np.array_equal(ret, rec[order, 1]) and np.array_equal(length, rec.view(np.int32)[order, 2])
    live = int(np.nonzero(~done)[0][0])
    assert infos[live] == {"TimeLimit.truncated": False}
    assert infos[-1] == infos[n - 1] and len(infos[2:5]) == 3
    # the sorted payload is a copy: overwriting the record block (the pinned block going back to the pool) changes nothing
    want = infos[i]["terminal_observation"].copy()
    rec[:] = 0
    assert np.array_equal(infos[i]["terminal_observation"], want)


def test_lazy_infos_without_finished_episodes():
    infos = LazyInfos(4, np.zeros(4, bool), np.zeros(4, bool), None, 0.0)
    assert infos[0] == {"TimeLimit.truncated": False} and len(infos.finished()) == 0
    ret, length = infos.episode_stats()
    assert len(ret) == 0 and len(length) == 0
    try:
        infos[4]
    except IndexError:
        pass
    else:
        raise AssertionError("index past the end must raise")
