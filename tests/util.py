"""Shared helpers for the parity tests."""
import numpy as np


def rel_err(got, want, region, floor=0.0):
    """max |got - want| over the region, relative to max |want| (the north_star's metric).
    `floor` guards results that are analytically zero (pure round-off on both sides)."""
    d = np.abs(got[region] - want[region]).max()
    s = max(np.abs(want[region]).max(), floor)
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


def make_metric(orc, shape2, nz, seed=11, trivial=False):
    """A Metric bundle with every Field2D varying in x and y (values need not be geometrically
    consistent: both sides of a parity test consume the arrays as given).  fake_mesh_fixture.hxx
    style constants (G1=G2=G3=0.1, d1_dx=0.2) when trivial."""
    nx, ny = shape2
    X = ((np.arange(nx)[:, None] + 0.5) / nx) * np.ones(shape2)
    Y = (np.arange(ny)[None, :] / max(ny, 1)) * np.ones(shape2)
    one = np.ones(shape2)
    if trivial:
        vals = dict(dx=0.2 * one, dy=one.copy(), dz=0.2 * one, g11=one.copy(), g22=one.copy(), g33=one.copy(),
                    g12=0 * one, g13=0 * one, g23=0 * one, g_11=one.copy(), g_22=one.copy(), g_33=one.copy(),
                    g_12=0 * one, g_13=0 * one, g_23=0 * one, J=one.copy(), Bxy=one.copy(), G1=0 * one,
                    G2=0 * one, G3=0 * one, d1_dx=0 * one, d1_dy=0 * one)
        m = orc.Metric(**{k: np.ascontiguousarray(v) for k, v in vals.items()})
        m.non_uniform = False
        return m
    vals = dict(
        dx=0.2 * (1 + 0.1 * X + 0.02 * np.cos(2 * np.pi * Y)), dy=0.7 + 0.05 * np.cos(2 * np.pi * X),
        dz=(2 * np.pi / nz) * one, g11=1 + 0.1 * X + 0.03 * Y, g22=1.1 + 0.1 * Y, g33=1.2 + 0.05 * X,
        g12=0.01 * one, g13=0.05 + 0.02 * X, g23=0.02 * one, g_11=1 / (1 + 0.1 * X), g_22=0.9 + 0.1 * X,
        g_33=0.8 + 0.1 * Y, g_12=0.01 * one, g_13=-0.04 * one, g_23=0.01 * one, J=1 + 0.2 * X,
        Bxy=1 + 0.3 * X, G1=0.1 + 0.05 * X, G2=0.1 * one, G3=0.1 - 0.03 * Y, d1_dx=0.2 + 0.1 * X,
        d1_dy=0.2 * one)
    m = orc.Metric(**{k: np.ascontiguousarray(v) for k, v in vals.items()})
    m.non_uniform = True
    return m
