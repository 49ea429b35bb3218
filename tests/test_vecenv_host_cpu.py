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
    assert infos[live] == {"TimeLimit.truncated": False}         # (no step bookkeeping passed: `steps` only at episode ends)
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


class _FakeLib:
    """Stands in for libtmla's result-block calls (plain host memory instead of cudaHostAlloc) so the pool logic runs on CPU."""

    def __init__(self, n, d, rec_words):
        import ctypes as C

        self.C, self.n, self.d, self.rw, self.bufs, self.freed = C, n, d, rec_words, [], []

    def tmla_result_block_layout(self, h, off, nbytes):
        n, d = self.n, self.d
        o = [0, 4 * n * d, 4 * n * d + 4 * n, 4 * n * d + 5 * n, 4 * n * d + 6 * n, 4 * n * d + 6 * n + 16]
        for i, v in enumerate(o):
            off[i] = v
        nbytes._obj.value = o[5] + 4 * self.rw * n
        return 0

    def tmla_host_records(self, h, p, w):
        w._obj.value = self.rw
        return 0

    def tmla_result_block_alloc(self, h, q):
        buf = (self.C.c_uint8 * (4 * self.n * self.d + 6 * self.n + 16 + 4 * self.rw * self.n))()
        self.bufs.append(buf)
        q._obj.value = self.C.addressof(buf)
        return 0

    def tmla_result_block_free(self, q):
        self.freed.append(q.value)
        return 0


def test_result_block_pool_reuses_a_block_only_when_nothing_references_it(monkeypatch):
    """DummyVecEnv hands out fresh arrays every step; the pool must never recycle memory a caller can still see — whether the
    caller holds one of the returned arrays, something derived from them, a buffer export, or the record slice.  Every
    hand-out is a lease (a fresh base array + weakref callback): no reference counts are inspected."""
    import sys

    import three_mlagents_b200.vec_env as ve

    fake = _FakeLib(64, 6, 12)
    monkeypatch.setattr(ve, "lib", fake)
    monkeypatch.setattr(ve, "check", lambda rc: None)
    pool = ve._ResultBlocks(None, 64, 6)
    assert len(pool._mem) == 3 and sorted(pool._free) == [1, 2]  # scratch + two pre-allocated blocks
    assert "getrefcount" not in open(ve.__file__).read()

    def busy():
        return {1, 2, 3, 4, 5, 6, 7, 8}.intersection(range(len(pool._mem))) - set(pool._free)

    k = pool.acquire()
    obs, rew, done, trunc, rec = pool.views(k, 5)
    assert obs.shape == (64, 6) and rew.shape == (64,) and done.dtype == np.bool_ and rec.shape == (5, 9)
    assert obs.base is rew.base is done.base is trunc.base       # one lease base per hand-out
    assert busy() == {k}
    k2 = pool.acquire()
    assert k2 != k                                               # block k is held
    pool.release(k2)                                             # (acquired but not handed out: the step failed)
    del obs, rew, done, trunc
    assert busy() == {k}                                         # ... still held, through the record slice
    del rec
    assert busy() == set()
    extra = [sys.getrefcount]                                    # a tracer-style extra reference only DELAYS the reuse
    k = pool.acquire()
    arrs = pool.views(k, 0)
    extra.append(arrs[0])
    del arrs
    assert busy() == {k}
    extra.pop()
    assert busy() == set()
    k = pool.acquire()
    part = pool.views(k, 0)[0][3:5]                              # a derived view keeps the block alive
    assert busy() == {k}
    del part
    assert busy() == set()
    k = pool.acquire()
    export = memoryview(pool.views(k, 0)[2])                     # so does a buffer export (torch.from_numpy, memoryview)
    assert busy() == {k}
    del export
    assert busy() == set()
    held = []
    for _ in range(12):                                          # a caller that never lets go: the pool grows to its cap, then says no
        k = pool.acquire()
        held.append(None if k is None else pool.views(k, 0)[0])
    assert [h is not None for h in held] == [True] * 8 + [False] * 4 and len(pool._mem) == 9
    survivor = held[0]
    del held, k
    assert len(busy()) == 1                                      # one block is still referenced by `survivor`
    pool.close()
    assert len(fake.freed) == 8                                  # everything but the block `survivor` still sees
    assert survivor.shape == (64, 6)


def test_lazy_infos_reports_steps_of_running_envs():
    """The reference's `_info` (envs.py:154-159) carries `steps` on every step, not only at episode ends."""
    last_reset = np.array([0, 3, 5, 5], np.int64)
    infos = LazyInfos(4, np.zeros(4, bool), np.zeros(4, bool), None, 0.0, step_no=7, last_reset=last_reset)
    assert [infos[i]["steps"] for i in range(4)] == [7, 4, 2, 2]
