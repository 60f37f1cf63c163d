"""Counter-based random numbers shared by every place that needs reproducible draws keyed by
(seed, stream, counter): reset noise, random-exploration actions.  Pure integer hashing
(splitmix64 finaliser), so a draw never depends on batch size, GPU count or call order."""
from __future__ import annotations

import numpy as np

_M1, _M2, _G, _C = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB), np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1342543DE82EF95)


def _mix(x):
    x = x ^ (x >> np.uint64(30))
    x = x * _M1
    x = x ^ (x >> np.uint64(27))
    x = x * _M2
    return x ^ (x >> np.uint64(31))


def uniform01(seed, stream, counter, dim):
    """U[0,1) with 53-bit resolution.  All arguments broadcast (numpy integer arrays)."""
    with np.errstate(over="ignore"):
        s = np.asarray(seed, dtype=np.uint64)
        k = np.asarray(stream, dtype=np.uint64)
        c = np.asarray(counter, dtype=np.uint64)
        d = np.asarray(dim, dtype=np.uint64)
        x = _mix(s ^ (k * _G))
        x = _mix(x + ((c << np.uint64(8)) | d) * _C)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed, stream, counter, dim):
    """Standard normal by Box-Muller on two independent uniforms."""
    u1 = uniform01(seed, stream, counter, np.asarray(dim, dtype=np.uint64) * np.uint64(2))
    u2 = uniform01(seed, stream, counter, np.asarray(dim, dtype=np.uint64) * np.uint64(2) + np.uint64(1))
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
