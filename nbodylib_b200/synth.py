"""Synthetic particle distributions for tests and bench.py.  The reference ships no generator
(InitCond.h:34-48 has them commented out), so these are builder-defined and documented in DESIGN.md.
All values are rounded to fp32 so the fp64 reference and the device path see identical coordinates."""
import numpy as np


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)


def uniform_box(n, seed=12345):
    """cfg 1: x, v ~ U[0,1)^3, m = 1."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32)
    vel = rng.random((n, 3), dtype=np.float32)
    return pos.astype(np.float64), vel.astype(np.float64), np.ones(n)


def clustered_small(n, seed=1, nhalo=24, frac=0.5):
    """Small clustered set for exhaustive parity: uniform background + Plummer spheres, unit box, wrapped."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3))
    vel = 0.05 * rng.standard_normal((n, 3))
    nh = int(frac * n)
    centres = rng.random((nhalo, 3))
    which = rng.integers(0, nhalo, nh)
    a = 0.02 * (0.3 + rng.random(nhalo))
    u = rng.random(nh) * 0.99
    r = a[which] / np.sqrt(np.maximum(u, 1e-9) ** (-2.0 / 3.0) - 1.0)
    d = rng.standard_normal((nh, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    pos[:nh] = (centres[which] + r[:, None] * d) % 1.0
    vel[:nh] = 0.2 * rng.standard_normal((nhalo, 3))[which] + 0.03 * rng.standard_normal((nh, 3))
    mass = 1.0 + rng.random(n)
    return _f32(pos), _f32(vel), _f32(mass)
