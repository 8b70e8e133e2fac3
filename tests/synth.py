"""Deterministic synthetic data for tests, golden fixtures and the benchmark.

A counter-based generator (splitmix64 hash of the element index and a seed) so that the same
(shape, seed) yields bit-identical arrays on every machine, numpy version and device — the
golden fixtures store only outputs, inputs are regenerated from their seed.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _bits(n: int, seed: int, stream: int) -> np.ndarray:
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        return _splitmix64(idx ^ key)


def uniform(shape, seed: int, dtype=np.float64, stream: int = 0) -> np.ndarray:
    """U(0,1), 53-bit resolution, never exactly 0."""
    n = int(np.prod(shape))
    u = ((_bits(n, seed, stream) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    return u.reshape(shape).astype(dtype)


def normal(shape, seed: int, dtype=np.float32) -> np.ndarray:
    """N(0,1) by Box-Muller on two independent uniform streams."""
    n = int(np.prod(shape))
    u1 = uniform((n,), seed, stream=1)
    u2 = uniform((n,), seed, stream=2)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return z.reshape(shape).astype(dtype)
