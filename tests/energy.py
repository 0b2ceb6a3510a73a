"""Kinetic energy of a spectral state evaluated in GRID space through a `module sht` implementation.

E = 1/2 int |u|^2 dV with u_r = vr/r^2, u_theta = vt/(r sin), u_phi = vp/(r sin) from torpol_to_spat (Robert form,
sht_native.f90:99-127), Gauss-Legendre quadrature in theta, uniform phi over the minc-fold sector, Chebyshev
(Clenshaw-Curtis) quadrature in r -- the same quantities kinetic_energy.f90 integrates spectrally and prints to e_kin.TAG.
"""
import numpy as np
from numpy.polynomial import chebyshev as C


def cheb_setup(radius):
    ro, ri = radius[0], radius[-1]
    x = (2 * radius - (ro + ri)) / (ro - ri)
    return x, 0.5 * (ro - ri)


def radial_derivative(radius, f):
    """d/dr of f[n_r, ...] on the Gauss-Lobatto grid by exact differentiation of the degree n_r-1 interpolant."""
    x, half = cheb_setup(radius)
    n = len(radius) - 1
    shp = f.shape
    f2 = f.reshape(len(radius), -1)
    out = np.empty_like(f2)
    for part in (0, 1):
        g = f2.real if part == 0 else f2.imag
        c = C.chebfit(x, g, n)
        d = C.chebval(x, C.chebder(c)).T / half
        if part == 0:
            out.real = d
        else:
            out.imag = d
    return out.reshape(shp)


def radial_integral(radius, f):
    x, half = cheb_setup(radius)
    n = len(radius) - 1
    c = C.chebfit(x, f, n)
    k = np.arange(n + 1)
    with np.errstate(divide="ignore"):
        wk = np.where(k % 2 == 0, 2.0 / (1.0 - k.astype(float) ** 2), 0.0)
    return half * np.dot(wk, c)


def kinetic_energy_grid(sht, gauss, sinTheta, minc, radius, w, z, lcut):
    """sht: object with torpol_to_spat(W, dW, Z, lcut) -> (vr, vt, vp) [n_phi, n_theta]; gauss/sinTheta in the same
    theta order as the grids.  Returns (E_pol, E_tor)."""
    dw = radial_derivative(radius, w)
    n_phi = None
    e_pol, e_tor = [], []
    zero = np.zeros_like(w[0])
    for ir, r in enumerate(radius):
        res = []
        for (W, dW, Z) in ((w[ir], dw[ir], zero), (zero, zero, z[ir])):
            vr, vt, vp = sht.torpol_to_spat(W, dW, Z, lcut)
            n_phi = vr.shape[0]
            u2 = vr[:, :len(gauss)] ** 2 / r ** 4 + (vt[:, :len(gauss)] ** 2 + vp[:, :len(gauss)] ** 2) / (r ** 2 * sinTheta[None, :] ** 2)
            dphi = 2 * np.pi / (n_phi * minc)
            res.append(0.5 * minc * dphi * np.sum(u2 * gauss[None, :]))
        e_pol.append(res[0] * r ** 2)
        e_tor.append(res[1] * r ** 2)
    return radial_integral(radius, np.array(e_pol)), radial_integral(radius, np.array(e_tor))
