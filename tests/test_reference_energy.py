"""Pins the transforms against a GOLDEN VECTOR OF THE REFERENCE: samples/boussBenchSat.

tests/golden/boussBenchSat_ckpt.npz holds the spectra of the reference's own checkpoint fixture
(samples/boussBenchSat/checkpoint_end.start) and the first row of samples/boussBenchSat/reference.out, the numbers
MagIC's autotest compares e_kin.TAG against (rtol 1e-8).  The saturated benchmark dynamo is a steadily drifting
solution, so its energies are constant in time to every printed digit; the kinetic energy of the checkpoint state,
evaluated in GRID space through torpol_to_spat (tests/energy.py), must therefore reproduce the golden columns
e_kin_pol, e_kin_tor and their axisymmetric parts.  This checks normalisation (orthonormal Y_lm, Hermitian factor 2),
the Robert form of the vector synthesis, l(l+1) scaling, the Gauss grid and the minc=4 sector against output of the
real magic.exe.  reference.out prints 9 significant digits, hence rtol 2e-9.
"""
import os

import numpy as np
import pytest

from tests.energy import kinetic_energy_grid

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 2e-9


@pytest.fixture(scope="module")
def ckpt():
    d = np.load(os.path.join(HERE, "golden", "boussBenchSat_ckpt.npz"))
    return {k: d[k] for k in d.files}


def _axi(a, lm2m):
    b = a.copy()
    b[:, lm2m != 0] = 0
    return b


def _check(sht, gauss, sinTheta, lm2m, ck):
    l_max, minc = int(ck["l_max"]), int(ck["minc"])
    ref = ck["reference_out_row0"]
    ep, et = kinetic_energy_grid(sht, gauss, sinTheta, minc, ck["radius"], ck["w"], ck["z"], l_max)
    assert abs(ep / ref[1] - 1) < RTOL and abs(et / ref[2] - 1) < RTOL, (ep, et, ref[1:3])
    epa, eta = kinetic_energy_grid(sht, gauss, sinTheta, minc, ck["radius"], _axi(ck["w"], lm2m), _axi(ck["z"], lm2m), l_max)
    assert abs(epa / ref[3] - 1) < RTOL and abs(eta / ref[4] - 1) < 2e-8, (epa, eta, ref[3:5])  # ref[4] is printed with 7 digits


def test_oracle_reproduces_reference_kinetic_energy(ckpt):
    from oracle.oracle import Oracle
    l_max, minc = int(ckpt["l_max"]), int(ckpt["minc"])
    o = Oracle(l_max, minc=minc, n_theta=int(ckpt["n_theta_max"]), n_phi=int(ckpt["n_phi_tot"]) // minc, m_max=int(ckpt["m_max"]),
               threads=os.cpu_count() or 1)
    assert o.lm_max == ckpt["w"].shape[1]
    _check(o, o.gauss, o.sinTheta, o.lm2m, ckpt)


@pytest.mark.gpu
def test_gpu_reproduces_reference_kinetic_energy(ckpt):
    from magic_b200 import Sht
    l_max, minc = int(ckpt["l_max"]), int(ckpt["minc"])
    n_theta = int(ckpt["n_theta_max"])
    s = Sht(l_max, m_max=int(ckpt["m_max"]), minc=minc, n_theta_max=n_theta, n_phi_max=int(ckpt["n_phi_tot"]) // minc)
    th, g = s.get_grid()
    gauss = np.empty(n_theta)
    sinT = np.empty(n_theta)
    gauss[0::2], gauss[1::2] = g[: n_theta // 2], g[: n_theta // 2]       # N/S interleaved rows (horizontal.f90:180-188)
    sinT[0::2], sinT[1::2] = np.sin(th[: n_theta // 2]), np.sin(th[: n_theta // 2])
    _check(s, gauss, sinT, s.lm2m, ckpt)
    s.finalize_sht()


@pytest.mark.gpu
def test_gpu_radial_loop_on_reference_state(ckpt):
    """The batched radial loop on the real saturated MHD state of the reference fixture (smooth spectra) vs the oracle."""
    from magic_b200 import RadialLoop, Sht
    from magic_b200.workload import make_params, make_radial
    from oracle.oracle import Oracle, Params as OParams
    from tests.energy import radial_derivative
    from tests.util import rel_l2
    l_max, minc, n_r = int(ckpt["l_max"]), int(ckpt["minc"]), int(ckpt["n_r_max"])
    n_theta, n_phi = int(ckpt["n_theta_max"]), int(ckpt["n_phi_tot"]) // minc
    r = ckpt["radius"]
    f = {k: ckpt[k] for k in ("w", "z", "s", "b", "aj")}
    f["dw"] = radial_derivative(r, f["w"]); f["ddw"] = radial_derivative(r, f["dw"]); f["dz"] = radial_derivative(r, f["z"])
    f["db"] = radial_derivative(r, f["b"]); f["ddb"] = radial_derivative(r, f["db"]); f["dj"] = radial_derivative(r, f["aj"])
    o = Oracle(l_max, minc=minc, n_theta=n_theta, n_phi=n_phi, m_max=int(ckpt["m_max"]), threads=os.cpu_count() or 1)
    s = Sht(l_max, m_max=int(ckpt["m_max"]), minc=minc, n_theta_max=n_theta, n_phi_max=n_phi)
    p = make_params("mhd", n_r)
    rad = make_radial(n_r, l_max)
    for k in ("r", "or1", "or2", "or4"):
        rad[k] = {"r": r, "or1": 1 / r, "or2": 1 / r ** 2, "or4": 1 / r ** 4}[k]
    rl = RadialLoop(s, p, rad)
    got = rl.radialLoop(f)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    ref = o.radial_loop(op, rad, f)
    bulk = slice(1, n_r - 1)
    for nm in ("dwdt", "dzdt", "dsdt", "dbdt", "djdt"):
        assert rel_l2(got[nm][bulk], ref[nm][bulk]) < 1e-12, nm
    assert rel_l2(got["dpdt"][bulk][:, 1:], ref["dpdt"][bulk][:, 1:]) < 1e-12
    assert rel_l2(got["dVxBhLM"], ref["dVxBhLM"]) < 1e-12 and rel_l2(got["dVSrLM"], ref["dVSrLM"]) < 1e-12
    assert np.allclose(got["dtrkc"], ref["dtrkc"], rtol=1e-12) and np.allclose(got["dthkc"], ref["dthkc"], rtol=1e-12)
    rl.finalize()
    s.finalize_sht()
