"""Pins the remaining `module sht` procedures of the oracle -- the inner-core and axisymmetric syntheses
-- to procedures that ARE pinned to reference output.

torpol_to_spat_IC / torpol_to_curl_spat_IC (sht_native.f90:143-229) only pre-scale their spectra with powers of r / r_ICB and
call native_qst_to_spat; axi_to_spat / toraxi_to_spat (shtransforms.f90:374-491) are the m = 0 columns of the scalar and the
toroidal synthesis.  scal_to_spat, sphtor_to_spat and torpol_to_spat are pinned to MagIC's golden energies (boussBenchSat e_kin,
the nine golden runs).  So: restate the few scaling lines independently in numpy, push the result through the pinned
procedures, and require the oracle's wrappers to agree.  The CUDA library is compared with these oracle wrappers in
tests/test_sht_gpu.py, which closes the chain reference -> pinned oracle transforms -> oracle wrappers -> GPU.
"""
import numpy as np
import pytest

from tests.util import random_spectrum, rel_l2


@pytest.fixture(scope="module", params=[(21, 1), (32, 3)])
def orc(request):
    from oracle.oracle import Oracle
    l_max, minc = request.param
    return Oracle(l_max, minc=minc)


def _ic_factors(o, r, r_icb):
    l = o.lm2l.astype(float)
    ratio = r / r_icb
    return ratio ** (l + 1.0), ratio ** l / r_icb, l * (l + 1.0)   # rDep(l), rDep2(l), dLh (sht_native.f90:165-171)


def test_torpol_to_spat_IC_is_the_prescaled_qst_synthesis(orc):
    rng = np.random.default_rng(5)
    W, dW, Z = (random_spectrum(orc, rng, zero_l0=True) for _ in range(3))
    r, r_icb = 0.31, 0.5384615384615384
    rDep, rDep2, dLh = _ic_factors(orc, r, r_icb)
    Br, Bt, Bp = orc.torpol_to_spat_IC(r, r_icb, W, dW, Z)
    # sht_native.f90:213-222: Q = rDep dLh W, S = rDep2 ((l+1) W + r dW), T = rDep Z
    want_r = orc.scal_to_spat(rDep * dLh * W, orc.l_max)
    want_t, want_p = orc.sphtor_to_spat(rDep2 * ((orc.lm2l + 1.0) * W + r * dW), rDep * Z, orc.l_max)
    assert rel_l2(Br, want_r) < 1e-13 and rel_l2(Bt, want_t) < 1e-13 and rel_l2(Bp, want_p) < 1e-13


def test_torpol_to_curl_spat_IC_is_the_prescaled_qst_synthesis(orc):
    rng = np.random.default_rng(6)
    dB, ddB, J, dJ = (random_spectrum(orc, rng, zero_l0=True) for _ in range(4))
    r, r_icb = 0.4, 0.5384615384615384
    rDep, rDep2, dLh = _ic_factors(orc, r, r_icb)
    cbr, cbt, cbp = orc.torpol_to_curl_spat_IC(r, r_icb, dB, ddB, J, dJ)
    # sht_native.f90:170-179: Q = rDep dLh J, S = rDep2 ((l+1) J + r dJ), T = -rDep2 (2 (l+1) dB + r ddB)
    want_r = orc.scal_to_spat(rDep * dLh * J, orc.l_max)
    want_t, want_p = orc.sphtor_to_spat(rDep2 * ((orc.lm2l + 1.0) * J + r * dJ), -rDep2 * (2.0 * (orc.lm2l + 1.0) * dB + r * ddB), orc.l_max)
    assert rel_l2(cbr, want_r) < 1e-13 and rel_l2(cbt, want_t) < 1e-13 and rel_l2(cbp, want_p) < 1e-13


def test_axisymmetric_syntheses_are_the_m0_columns(orc):
    rng = np.random.default_rng(7)
    fl = rng.standard_normal(orc.l_max + 1) + 0j
    full = np.zeros(orc.lm_max, dtype=np.complex128)
    full[: orc.l_max + 1] = fl           # st_map: the m = 0 block comes first (blocking.f90:309-317)
    # axi_to_spat (shtransforms.f90:374-403) = scal_to_spat of an m = 0 spectrum at any longitude
    f = orc.axi_to_spat(fl)
    want = orc.scal_to_spat(full, orc.l_max)
    assert rel_l2(f, want[0]) < 1e-13 and np.abs(want - want[0][None, :]).max() < 1e-13 * np.abs(want).max()
    # toraxi_to_spat (shtransforms.f90:405-491) = sphtor_to_spat with S = 0 and an m = 0 toroidal spectrum
    fl[0] = 0
    full[0] = 0
    lcut = orc.l_max - 2
    ft, fp = orc.toraxi_to_spat(fl, lcut)
    wt, wp = orc.sphtor_to_spat(np.zeros_like(full), full, lcut)
    assert np.abs(ft - wt[0]).max() < 1e-13 * np.abs(wp).max() and rel_l2(fp, wp[0]) < 1e-13
