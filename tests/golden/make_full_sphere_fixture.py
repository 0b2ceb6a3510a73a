"""Builds tests/golden/full_sphere_reference.npz from the REFERENCE's own test fixture samples/full_sphere.

Run in the build container only (reads /root/reference, which does not exist on the GPU box); the .npz travels.

Sources:
  /root/reference/samples/full_sphere/checkpoint_end.start -- version-2 checkpoint of the saturated Marti et al. (2014)
      full-sphere benchmark (l_max=32, minc=3, n_r_max=96, finite differences, hydro + heat, CNAB2).  Binary layout:
      python/magic/checkpoint.py:165-335 and src/readCheckPoints.f90:840-1060, :1508-1601 -- after every field follows the
      explicit term of the previous step (`d?dt%expl(:,:,2)`, nexp+nimp+nold-3 = 1 extra array for CNAB2).
  /root/reference/samples/full_sphere/reference.out -- e_kin.TAG of the 100-step restart (logged every 10 steps), the
      golden numbers samples/full_sphere/unitTest.py compares at rtol 1e-8.
  /root/reference/samples/full_sphere/input.nml -- the run parameters, copied by hand below.
"""
import os

import numpy as np

REF = "/root/reference/samples/full_sphere"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_sphere_reference.npz")


def read_checkpoint(path):
    f = open(path, "rb")
    version = np.fromfile(f, "i4", 1)[0]
    assert version == 2, version
    time = np.fromfile(f, "f8", 1)[0]
    family = f.read(10).decode()
    assert family.startswith("MULTISTEP")
    nexp, nimp, nold = np.fromfile(f, np.int32, 3)
    assert (nexp, nimp, nold) == (2, 1, 1)          # CNAB2
    dt = np.fromfile(f, np.float64, nexp)
    n_time_step = np.fromfile(f, np.int32, 1)[0]
    ra, pr, raxi, sc, prmag, ek, radratio, sigma_ratio = np.fromfile(f, np.float64, 8)
    n_r_max, n_theta_max, n_phi_tot, minc, nalias, n_r_ic_max = np.fromfile(f, np.int32, 6)
    l_max = nalias * n_phi_tot // 60                # readCheckPoints.f90:931
    m_max = (l_max // minc) * minc
    lm_max = sum(l_max - m + 1 for m in range(0, m_max + 1, minc))
    rscheme = f.read(72).decode()
    assert rscheme.startswith("fd")
    fd_order, fd_order_bound = np.fromfile(f, np.int32, 2)
    fd_stretch, fd_ratio = np.fromfile(f, np.float64, 2)
    radius = np.fromfile(f, np.float64, n_r_max)
    np.fromfile(f, np.float64, 4 * (nexp + nimp + nold - 3))     # domega_ic/ma, lorentz torques: zero here
    np.fromfile(f, np.float64, 12)
    l_heat, l_chem, l_mag, l_press, l_cond_ic = np.fromfile(f, np.int32, 5)
    assert (l_heat, l_chem, l_mag, l_press, l_cond_ic) == (1, 0, 0, 0, 0)

    def field():
        return np.fromfile(f, np.complex128, n_r_max * lm_max).reshape(n_r_max, lm_max)

    out = {}
    for nm in ("w", "z", "s"):
        out[nm] = field()
        out["d%sdt_expl2" % nm] = field()
    assert f.read() == b""
    out.update(radius=radius, time=time, dt=dt, n_time_step=n_time_step, ra=ra, pr=pr, ek=ek, radratio=radratio,
               n_r_max=n_r_max, n_theta_max=n_theta_max, n_phi_tot=n_phi_tot, minc=minc, nalias=nalias, l_max=l_max,
               m_max=m_max, fd_order=fd_order, fd_order_bound=fd_order_bound, fd_stretch=fd_stretch, fd_ratio=fd_ratio)
    return out


if __name__ == "__main__":
    ck = read_checkpoint(os.path.join(REF, "checkpoint_end.start"))
    e_kin = np.loadtxt(os.path.join(REF, "reference.out"))
    # input.nml: &control dtmax, alpha, courfac, alffac, l_correct_AMz; &phys_param epsc0, ktops, ktopv; &grid l_var_l;
    # &output_control n_log_step.  radratio=0 => l_full_sphere => kbotv=1, kbots=2, g0=g2=0 (Namelists.f90:439-468);
    # radial_scheme='FD' => l_double_curl (Namelists.f90:299-304); rcut_l default 0.1 (Namelists.f90).
    nml = dict(dtmax=5.0e-6, alpha=0.5, courfac=2.5, alffac=1.0, epsc0=3.0, ktops=1, ktopv=1, kbotv=1, kbots=2,
               n_log_step=10, n_time_steps=100, rcut_l=0.1)
    np.savez_compressed(OUT, e_kin=e_kin, **ck, **nml)
    print(OUT, os.path.getsize(OUT), "bytes; l_max", ck["l_max"], "minc", ck["minc"], "n_r", ck["n_r_max"], "rows", e_kin.shape)
