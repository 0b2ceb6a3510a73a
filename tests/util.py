"""Shared helpers for the parity tests (seeded synthetic spectra)."""
import numpy as np


def random_spectrum(o, rng, zero_l0=False, scale=True):
    """Re,Im ~ N(0,1)/(l+1), Im=0 at m=0."""
    s = rng.standard_normal(o.lm_max) + 1j * rng.standard_normal(o.lm_max)
    if scale:
        s = s / (o.lm2l + 1.0)
    s[o.lm2m == 0] = s[o.lm2m == 0].real
    if zero_l0:
        s[o.lm2l == 0] = 0
    return s


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d
