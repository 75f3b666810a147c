"""Seeded synthetic inputs shared by the parity tests."""
import numpy as np


def plaw(rng, n, xmin, xmax, p=-2.0):
    # inverse CDF, as in the reference's src/tests/synchrotron.py:65-68
    return ((xmax ** (p + 1) - xmin ** (p + 1)) * rng.random(n) + xmin ** (p + 1)) ** (1 / (p + 1))


def isotropic(rng, n):
    mu = 2 * rng.random(n) - 1
    phi = 2 * np.pi * rng.random(n)
    st = np.sqrt(1 - mu**2)
    return st * np.cos(phi), st * np.sin(phi), mu


def config2(n, seed=123, umin=5e-3, umax=2e3):
    """isotropic U with |U| ~ u^-2 on [umin, umax] (about half of it below the first bin)"""
    rng = np.random.default_rng(seed)
    u = plaw(rng, n, umin, umax)
    nx, ny, nz = isotropic(rng, n)
    return [(u * c).astype(np.float32) for c in (nx, ny, nz)]


def config3(n, seed=123):
    """U1 ~ u^-2 on [1, 100], E = 0, B isotropic unit (reference src/tests/synchrotron.py:72-86)"""
    rng = np.random.default_rng(seed)
    b = isotropic(rng, n)
    u1 = plaw(rng, n, 1, 100)
    z = np.zeros(n, np.float32)
    U = [u1.astype(np.float32), z, z]
    E = [z, z, z]
    B = [c.astype(np.float32) for c in b]
    return U, E, B


def full3d(n, seed=321):
    """isotropic U, |B| in [0.5, 2], E = 0.1 B x random: exercises every term of chi_R"""
    rng = np.random.default_rng(seed)
    u = plaw(rng, n, 0.05, 500)
    nu = isotropic(rng, n)
    bn = 0.5 + 1.5 * rng.random(n)
    nb = isotropic(rng, n)
    nr = isotropic(rng, n)
    U = [(u * c).astype(np.float32) for c in nu]
    B = [(bn * c).astype(np.float32) for c in nb]
    E = [
        (0.1 * (B[1] * nr[2] - B[2] * nr[1])).astype(np.float32),
        (0.1 * (B[2] * nr[0] - B[0] * nr[2])).astype(np.float32),
        (0.1 * (B[0] * nr[1] - B[1] * nr[0])).astype(np.float32),
    ]
    return U, E, B


def rel_err(got, want, floor_frac=1e-6):
    """max relative error over bins whose reference value is >= floor_frac * max"""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    keep = (want >= floor_frac * want.max()) & (want > 0)
    if not keep.any():
        return 0.0
    return float(np.max(np.abs(got[keep] - want[keep]) / want[keep]))
