"""Shared helpers for the parity tests."""
import numpy as np


def rel_err(got, want, region):
    """max |got - want| over the region, relative to max |want| (the north_star's metric)."""
    d = np.abs(got[region] - want[region]).max()
    s = np.abs(want[region]).max()
    return d / s if s > 0 else d


def seeded_fields(shape, n, seed=20261017):
    """smooth closed form + 0.1 * U(-1,1) noise, fixed seed."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    X = (np.arange(nx)[:, None, None] + 0.5) / nx
    Y = 2 * np.pi * np.arange(ny)[None, :, None] / max(ny, 1)
    Z = 2 * np.pi * np.arange(nz)[None, None, :] / nz
    out = []
    for i in range(n):
        base = np.sin(2 * np.pi * (3 + i) * X) * np.cos((2 + i) * Z) * np.exp(-(Y - np.pi) ** 2 / 4)
        out.append(np.ascontiguousarray(base + 0.1 * rng.uniform(-1, 1, shape)))
    return out


def metric_2d(shape2, seed=3):
    """Non-trivial positive Field2D metric pieces."""
    rng = np.random.default_rng(seed)
    nx, ny = shape2
    X = (np.arange(nx)[:, None] + 0.5) / nx
    dx = 0.2 * (1 + 0.1 * X) + 0.0 * rng.uniform(size=shape2)
    dy = np.full(shape2, 0.7) + 0.05 * np.cos(2 * np.pi * X)
    dz = np.full(shape2, 2 * np.pi / 64)
    return np.ascontiguousarray(dx), np.ascontiguousarray(dy), np.ascontiguousarray(dz)
