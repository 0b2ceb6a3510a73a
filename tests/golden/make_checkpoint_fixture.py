"""Builds tests/golden/boussBenchSat_ckpt.npz from the REFERENCE's own test fixture.

Run in the build container only (reads /root/reference, which does not exist on the GPU box); the .npz travels.

Source: /root/reference/samples/boussBenchSat/checkpoint_end.start (version-4 checkpoint, l_max=64, minc=4,
n_r_max=33, Chebyshev, MHD; binary layout documented in python/magic/checkpoint.py:165-335 and
src/storeCheckPoints.f90:45-277) and the first row of samples/boussBenchSat/reference.out (e_kin.TAG columns:
time, e_kin_pol, e_kin_tor, axisymmetric pol/tor, ...), the golden numbers the reference's autotest compares
against (samples/boussBenchSat/unitTest.py, rtol 1e-8).  The saturated benchmark dynamo drifts steadily, so its
energies are constant in time to all printed digits: the energy of the checkpoint state IS the golden value.
"""
import os

import numpy as np

REF = "/root/reference/samples/boussBenchSat"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "boussBenchSat_ckpt.npz")


def read_checkpoint(path):
    f = open(path, "rb")
    version = np.fromfile(f, "i4", 1)[0]
    assert version == 4, version
    time = np.fromfile(f, "f8", 1)[0]
    family = f.read(10).decode()
    nexp, nimp, nold = np.fromfile(f, np.int32, 3)
    multistep = family.startswith("MULTISTEP")
    dt = np.fromfile(f, np.float64, nexp if multistep else 1)
    np.fromfile(f, np.int32, 1)  # n_time_step
    ra, pr, raxi, sc, prmag, ek, stef, radratio, sigma_ratio = np.fromfile(f, np.float64, 9)
    n_r_max, n_theta_max, n_phi_tot, minc, nalias, n_r_ic_max = np.fromfile(f, np.int32, 6)
    l_max, m_min, m_max = np.fromfile(f, np.int32, 3)
    lm_max = sum(l_max - m + 1 for m in range(m_min, m_max + 1, minc))
    rscheme = f.read(72).decode()
    assert rscheme.startswith("cheb")
    np.fromfile(f, np.int32, 2)
    np.fromfile(f, np.float64, 2)
    radius = np.fromfile(f, np.float64, n_r_max)
    if multistep:
        n = nexp + nimp + nold - 3
        np.fromfile(f, np.float64, 2 * n)
    omega = np.fromfile(f, np.float64, 12)
    l_heat, l_chem, l_phase, l_mag, l_press, l_cond_ic = np.fromfile(f, np.int32, 6)
    extra = (nexp + nimp + nold - 3) if multistep else 0

    def field():
        a = np.fromfile(f, np.complex128, n_r_max * lm_max).reshape(n_r_max, lm_max)
        if extra:
            np.fromfile(f, np.complex128, lm_max * n_r_max * extra)
        return a

    out = dict(w=field(), z=field())
    if l_press:
        out["p"] = field()
    if l_heat:
        out["s"] = field()
    assert not l_chem and not l_phase
    if l_mag:
        out["b"] = field()
        out["aj"] = field()
    if l_mag and l_cond_ic:      # inner-core potentials (for tests/test_boussBenchSat.py)
        for nm in ("b_ic", "aj_ic"):
            out[nm] = np.fromfile(f, np.complex128, n_r_ic_max * lm_max).reshape(n_r_ic_max, lm_max)
        assert f.read() == b""
    out.update(omega_ic1=omega[0], dt=dt, reference_out=np.loadtxt(os.path.join(REF, "reference.out")))
    out.update(radius=radius, l_max=l_max, m_max=m_max, minc=minc, n_r_max=n_r_max, n_theta_max=n_theta_max,
               n_phi_tot=n_phi_tot, ek=ek, prmag=prmag, radratio=radratio, time=time)
    return out


if __name__ == "__main__":
    ck = read_checkpoint(os.path.join(REF, "checkpoint_end.start"))
    ref_row0 = np.loadtxt(os.path.join(REF, "reference.out"))[0]
    np.savez_compressed(OUT, reference_out_row0=ref_row0, **ck)
    print(OUT, os.path.getsize(OUT), "bytes; l_max", ck["l_max"], "minc", ck["minc"], "n_r", ck["n_r_max"], "row0", ref_row0)
