"""Dry run of the GPU legs of the golden-run tests WITHOUT a GPU: `Sht` and `RadialLoop` are replaced by stand-ins that
forward to the CPU oracle, so that everything on the test side -- fixtures, parameter blocks, radial functions, the calling
sequence (set_rotation / torques / br_v_bcs), tolerances -- is executed once before the tests meet a device.  It proves
nothing about the CUDA path (the stand-ins ARE the oracle); it only keeps test-side mistakes from costing GPU minutes.

Usage (build container, CPU only):  python tools/dry_run_gpu_tests.py [name ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import magic_b200  # noqa: E402
from oracle.oracle import Oracle, Params as OParams  # noqa: E402


class FakeSht:
    def __init__(self, l_max, m_max=None, minc=1, n_theta_max=None, n_phi_max=None, **kw):
        from magic_b200.sht import grid_sizes
        if n_theta_max is None:
            gs = grid_sizes(l_max=l_max, minc=minc)
            n_theta_max, n_phi_max = gs["n_theta_max"], gs["n_phi_max"]
        if m_max is None:
            m_max = (l_max // minc) * minc
        self.o = Oracle(l_max, minc=minc, n_theta=n_theta_max, n_phi=n_phi_max, m_max=m_max, threads=min(4, os.cpu_count() or 1))
        self.l_max, self.m_max, self.minc = l_max, m_max, minc
        self.lm_max, self.lm2l, self.lm2m = self.o.lm_max, self.o.lm2l, self.o.lm2m
        self.n_theta_max, self.n_phi_max = n_theta_max, n_phi_max

    def get_grid(self):
        g = np.concatenate([self.o.gauss[0::2], self.o.gauss[0::2][::-1]])
        return np.asarray(self.o.theta_ord), g

    def finalize_sht(self):
        pass

    def __getattr__(self, name):          # the 17 procedures
        return getattr(self.o, name)


class FakeRadialLoop:
    def __init__(self, sht, params, radial, level_chunk=0):
        self.o, self.rad = sht.o, radial
        self.op = OParams()
        for n, _ in params._fields_:
            setattr(self.op, n, getattr(params, n))
        self.calls, self.out = 0, None

    def set_rotation(self, omega_ma, omega_ic):
        self.op.omega_ma, self.op.omega_ic = omega_ma, omega_ic

    def radialLoop(self, fields, time=0.0, out=None):
        self.calls += 1
        self.out = self.o.radial_loop(self.op, self.rad, fields, time=time)
        return self.out

    def torques(self):
        return self.out["lorentz_torque_ic"], self.out["lorentz_torque_ma"]

    def br_v_bcs(self, boundary):
        b = boundary.lower()
        return self.out["br_vt_lm_" + b], self.out["br_vp_lm_" + b]

    def launch_count(self):
        return self.calls

    def finalize(self):
        pass


magic_b200.Sht, magic_b200.RadialLoop = FakeSht, FakeRadialLoop


def _golden(mod):
    return mod.golden.__wrapped__() if hasattr(mod.golden, "__wrapped__") else mod.golden.__pytest_wrapped__.obj()


def main(names):
    import importlib
    jobs = {
        "precession": ("tests.test_precession", "test_gpu_radial_loop_reproduces_reference_energies"),
        "full_sphere": ("tests.test_full_sphere", "test_gpu_radial_loop_reproduces_reference_energies"),
        "varCond": ("tests.test_varCond", "test_gpu_radial_loop_reproduces_reference_energies"),
        "varProps": ("tests.test_varProps", "test_gpu_radial_loop_reproduces_reference_energies"),
        "doubleDiffusion": ("tests.test_doubleDiffusion", "test_gpu_radial_loop_reproduces_reference_energies"),
        "boussBenchSat": ("tests.test_boussBenchSat", "test_gpu_radial_loop_reproduces_reference_energies"),
        "condICrotIC": ("tests.test_condICrotIC", "test_gpu_radial_loop_reproduces_reference_energies"),
    }
    full = (not names) or ("full_size" in names)
    names = [n for n in names if n != "full_size"]
    for nm in (names or ([] if full and sys.argv[1:] else list(jobs))):
        mod = importlib.import_module(jobs[nm][0])
        t0 = time.time()
        getattr(mod, jobs[nm][1])(_golden(mod))
        print(f"{nm}: test body ran to the end with the oracle stand-in ({time.time() - t0:.0f} s)", flush=True)
    if full:
        mod = importlib.import_module("tests.test_full_size_gpu")
        for L in (255, 511):
            s = FakeSht(L)
            for fn in ("test_scalar_round_trip_and_parseval", "test_vector_round_trip", "test_lcut_zeros_and_linearity"):
                getattr(mod, fn).__wrapped__(s) if hasattr(getattr(mod, fn), "__wrapped__") else getattr(mod, fn)(s)
            print(f"full_size l_max={L}: property tests ran with the oracle stand-in", flush=True)
        mod.test_radial_loop_against_the_oracle(255, 121, "mhd")
        print("full_size: loop-vs-oracle body ran at l_max=255", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
