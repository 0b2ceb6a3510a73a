"""End-to-end golden vectors of the reference: samples/doubleDiffusion (Chebyshev stage).

The reference's autotest (`samples/doubleDiffusion/unitTest.py`, rtol 1e-8) restarts saturated double-diffusive convection
-- thermal (Ra = 4.8e4, Pr = 0.3) AND compositional (Ra_xi = 1.2e5, Sc = 3) buoyancy, l_max = 64 with minc = 4, n_r_max = 33 /
n_cheb_max = 31 -- from `checkpoint_end.start` and runs 25 steps of the IMEX Runge-Kutta scheme BPR353 (dt = 3e-4, three radial
loops per step), logging e_kin.TAG every 5 steps.  The solution is a steadily drifting wave: all six rows carry the same
energies.  On the radial-loop side this is the pinned case for l_chemical_conv: the composition field xi goes through the
synthesis, the advection products VXir/VXit/VXip of get_nl (get_nl.f90:318-323), the analysis, get_dxidt (get_td.f90:521-555)
and the dVXirLM output that finish_exp_comp differentiates (updateXI.f90:495-511); and for a loop that is called several times
per step on intermediate stage states.

Host: oracle/lmloop.py DirkShellHost (stage logic of dirk_schemes.f90 / step_time.f90, composition equation of updateXI.f90,
legacy boundary values translated on restart as startFields.f90:257-279 does).  The radial loop is the CPU oracle (CPU test,
5 steps = 15 loops) or the CUDA library through the C ABI (25 steps).  tests/golden/doubleDiffusion_reference.npz holds the
checkpoint fields and reference.out (tests/golden/make_doubleDiffusion_fixture.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-8          # samples/doubleDiffusion/unitTest.py


@pytest.fixture(scope="module")
def golden():
    d = np.load(os.path.join(HERE, "golden", "doubleDiffusion_reference.npz"))
    return {k: d[k] for k in d.files}


def _sizes(golden):
    from magic_b200.sht import grid_sizes
    gs = grid_sizes(n_phi_tot=int(golden["n_phi_tot"]), minc=int(golden["minc"]))
    assert (gs["l_max"], gs["lm_max"], gs["n_phi_max"]) == (64, 561, 48)
    return gs


class _Ckpt:
    """The part of magic_b200.checkpoint.Checkpoint that ShellHost.load_checkpoint reads, filled from the fixture."""

    def __init__(self, golden, lm_max):
        self.r, self.lm_max, self.time = golden["radius"], lm_max, float(golden["time"])
        self.fields = {k: golden[k] for k in ("w", "z", "p", "s", "xi")}
        self.past, self.scalars_past, self.rotation = {}, {}, {"omega_ic1": 0.0}


def _setup(golden, lm2l, lm2m):
    from magic_b200.workload import make_params, make_radial
    from oracle.lmloop import DirkShellHost
    n_r = int(golden["n_r_max"])
    kw = {k: float(golden[k]) for k in ("radratio", "ra", "ek", "pr", "dtmax", "alpha", "raxi", "sc")}
    h = DirkShellHost(lm2l, lm2m, None, n_r_max=n_r, n_cheb_max=int(golden["n_cheb_max"]), init_s1=0, init_b1=0, l_mag=False,
                      time_scheme="BPR353", **kw)
    h.load_checkpoint(_Ckpt(golden, len(lm2l)))
    assert h.l_chem and abs(h.s[0, 0]) < 1e-12 and abs(h.xi[-1, 0].real - np.sqrt(4 * np.pi)) < 1e-12
    p = make_params("hydro", n_r)
    p.l_chemical_conv = 1                                   # raxi /= 0 (Namelists.f90:421-427)
    p.ra, p.opr = kw["ra"], 1.0 / kw["pr"]
    p.courfac, p.alffac = float(golden["courfac"]), float(golden["alffac"])   # the scheme's own factors, dirk_schemes.f90:238-239
    rad = make_radial(n_r, 64)
    assert np.abs(rad["r"] - h.g.r).max() < 1e-15
    return h, p, rad


def _check(golden, h, row):
    got = np.concatenate([[h.time], h.e_kin()])
    np.testing.assert_allclose(got, golden["e_kin"][row], rtol=RTOL, err_msg=f"row {row}")


def _oracle_host(golden, tweak=None):
    from oracle.oracle import Oracle, Params as OParams
    gs = _sizes(golden)
    o = Oracle(gs["l_max"], minc=4, n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"],
               threads=min(4, os.cpu_count() or 1))
    h, p, rad = _setup(golden, o.lm2l, o.lm2m)
    op = OParams()
    for n, _ in p._fields_:
        setattr(op, n, getattr(p, n))
    if tweak:
        tweak(op)
    h.radial_loop = lambda f: o.radial_loop(op, rad, f)
    return h


def test_oracle_radial_loop_reproduces_reference_energies(golden):
    """CPU oracle inside the reference's Runge-Kutta loop: row 0 (restart state) and the first logged row (5 steps, 15 loops)."""
    h = _oracle_host(golden)
    _check(golden, h, 0)
    for _ in range(int(golden["n_log_step"])):
        h.step()
    _check(golden, h, 1)


def test_the_energies_see_the_composition_advection(golden):
    """Negative control: without the composition branch in the loop the steady state is left by 2.5e-3 within five steps."""
    h = _oracle_host(golden, tweak=lambda op: setattr(op, "l_chemical_conv", 0))
    for _ in range(int(golden["n_log_step"])):
        h.step()
    assert np.abs(h.e_kin() / golden["e_kin"][1][1:] - 1.0).max() > 1e-4


@pytest.mark.gpu
def test_gpu_radial_loop_reproduces_reference_energies(golden):
    """The CUDA radial loop (magic_rloop_run, host containers, xi / dxidt / dVXirLM) inside the reference's Runge-Kutta loop:
    all five logged rows (25 steps, 75 loops)."""
    from magic_b200 import RadialLoop, Sht
    gs = _sizes(golden)
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=4, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    h, p, rad = _setup(golden, s.lm2l, s.lm2m)
    rl = RadialLoop(s, p, rad)
    h.radial_loop = lambda f: rl.radialLoop(f)
    _check(golden, h, 0)
    for row in range(1, len(golden["e_kin"])):
        for _ in range(int(golden["n_log_step"])):
            h.step()
        _check(golden, h, row)
    assert rl.launch_count() > 0
    rl.finalize()
    s.finalize_sht()
