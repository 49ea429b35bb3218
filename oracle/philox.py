"""TEST INFRASTRUCTURE ONLY — Philox4x32-10 in NumPy (CPU restatement of the
counter-based generator the CUDA kernels use for auto-reset and sampling).

The reference draws resets from NumPy's global MT19937 (examples/ball3d.py:49-57,
gridworld.py:42-50, push.py:40-47 via mlagents/envs.py:116-121), which is
order-dependent across environments and cannot be reproduced by any per-env
stream.  This repo therefore defines its own stream layout
(DESIGN.md "Random streams"); this file is the bit-exact CPU twin of
`csrc/philox.cuh`, pinned by the Random123 known-answer vectors in
tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)

# stream tags (ctr[3] low nibble) — must match csrc/philox.cuh
TAG_RESET = 0       # auto-reset after the step with global index k-1 (k>=1)
TAG_ACTION = 1      # random-policy actions, 4 per block
TAG_SAMPLE = 2      # categorical sampling from the policy
TAG_PERM = 3        # minibatch permutation keys
TAG_RESET_ALL = 4   # VecEnv.reset() of every env at global step k


def philox4x32_10(ctr, key):
    """ctr: 4 arrays (uint32-valued), key: 2 arrays/scalars -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in ctr]
    c = list(np.broadcast_arrays(*c))
    k0 = np.asarray(key[0], dtype=np.uint64) & MASK
    k1 = np.asarray(key[1], dtype=np.uint64) & MASK
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> S32, p0 & MASK
        hi1, lo1 = p1 >> S32, p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return [x.astype(np.uint32) for x in c]


def stream_block(seed, env_id, k, tag, block=0):
    """The repo's counter layout: ctr = (env_lo, env_hi, k_lo, k_hi<<8 | block<<4 | tag),
    key = (seed_lo, seed_hi)."""
    env_id = np.asarray(env_id, dtype=np.uint64)
    k = np.asarray(k, dtype=np.uint64)
    seed = np.uint64(seed)
    c3 = ((k >> S32) << np.uint64(8)) | np.uint64((block << 4) | tag)
    return philox4x32_10(
        (env_id & MASK, env_id >> S32, k & MASK, c3 & MASK),
        (seed & MASK, seed >> S32),
    )


def bounded(x, n):
    """Lemire multiply-shift: floor(x * n / 2^32); bias <= n / 2^32."""
    return ((np.asarray(x, dtype=np.uint64) * np.uint64(n)) >> S32).astype(np.int32)


def u53(a, b):
    """53-bit double in [0,1) from two uint32 words (same recipe as MT19937's
    genrand_res53 that np.random.uniform uses: (a>>5)*2^26 + (b>>6)) / 2^53)."""
    a = (np.asarray(a, dtype=np.uint64) >> np.uint64(5)).astype(np.float64)
    b = (np.asarray(b, dtype=np.uint64) >> np.uint64(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0


def u24(a):
    """float32 in [0,1) from the top 24 bits of one word."""
    return ((np.asarray(a, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def u32_unit(a):
    """double in (0,1) with 32 random bits: (a + 0.5) * 2^-32 (exact)."""
    return (np.asarray(a, dtype=np.uint64).astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
