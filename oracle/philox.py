"""Philox4x32-10 counter-based RNG and the replay index derivation (oracle; test infra only).

The reference draws minibatch indices with ``numpy.random.randint(0, n, B)`` inside a numba
``@njit`` function (``General/Base/replay_buffer.py:77``): numba's private, never-seeded
MT19937 stream, which cannot be reproduced.  What *is* pinned by the reference is the
distribution: ``B`` iid uniform integers in ``[0, size)``, with replacement, int64.  The B200 path
generates them with Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 v1.14 ``philox.h``), and this file is the bit-exact CPU statement of that
generator and of the counter/key convention the kernels use:

    key     = (seed_lo, seed_hi)
    counter = (slot i, step_lo, step_hi, agent)
    x64     = out[0] | out[1] << 32
    index   = floor(x64 * size / 2**64)            (``__umul64hi`` on the device)

Known-answer vectors are Random123's ``kat_vectors`` entries for philox4x32-10.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)

# (counter, key, expected) -- Random123 kat_vectors, "philox4x32 10"
KAT_VECTORS = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  Inputs broadcastable uint32-valued arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0          # 32x32 -> 64, no overflow in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def mulhi64(x, n):
    """floor(x * n / 2**64) for uint64 array x and 0 < n < 2**31, without 128-bit ints."""
    n = int(n)
    assert 0 < n < (1 << 31), "ring sizes are limited to < 2^31 slots"
    x = np.asarray(x, dtype=np.uint64)
    hi, lo = x >> np.uint64(32), x & _MASK32
    n64 = np.uint64(n)
    return ((hi * n64 + ((lo * n64) >> np.uint64(32))) >> np.uint64(32)).astype(np.int64)


def sample_indices(seed, agent, step, batch_size, size):
    """Indices the kernels draw for train step ``step`` of ``agent``: int64[batch_size] in [0,size)."""
    i = np.arange(batch_size, dtype=np.uint64)
    step = int(step)
    o = philox4x32_10(i, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, int(agent) & 0xFFFFFFFF,
                      int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    x64 = o[0].astype(np.uint64) | (o[1].astype(np.uint64) << np.uint64(32))
    return mulhi64(x64, size)
