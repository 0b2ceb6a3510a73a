"""magic_b200.checkpoint: MagIC checkpoint reader / writer against the reference's own files and reader.

The two checkpoints the reference ships (samples/full_sphere: version 2, finite differences, CNAB2; samples/boussBenchSat:
version 4, Chebyshev, conducting inner core) exist only in the build container; the parts of them that travel are the
committed fixtures tests/golden/full_sphere_reference.npz and boussBenchSat_ckpt.npz, written by independent ad-hoc readers
(tests/golden/make_*_fixture.py).  The round-trip and error-path tests run everywhere.
"""
import os

import numpy as np
import pytest

from magic_b200.checkpoint import Checkpoint, CheckpointError, read_checkpoint, write_checkpoint

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/samples"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")


def _synthetic(family="MULTISTEP", nexp=3, nimp=2, nold=2, mag=True, cond_ic=True, minc=2):
    rng = np.random.default_rng(5)
    ck = Checkpoint()
    ck.time, ck.family, ck.nexp, ck.nimp, ck.nold = 1.25, family, nexp, nimp, nold
    ck.dt = rng.random(nexp if family == "MULTISTEP" else 1)
    ck.n_time_step = 4711
    ck.params.update(ra=1e5, pr=1.0, raxi=0.0, sc=10.0, prmag=5.0, ek=1e-3, stef=0.0, radratio=0.35, sigma_ratio=1.0)
    ck.trunc.update(n_r_max=9, n_theta_max=12, n_phi_tot=24, minc=minc, nalias=20, n_r_ic_max=5)
    ck.l_max, ck.m_min, ck.m_max = 8, 0, 8
    ck.rscheme.update(version="cheb", n=9, n2=0, ratio1=0.0, ratio2=0.0)
    ck.r = np.linspace(1.5, 0.5, 9)
    lm = ck.lm_max
    names = ["w", "z", "p", "s"] + (["b", "aj"] if mag else []) + (["b_ic", "aj_ic"] if mag and cond_ic else [])
    for nm in names:
        rows = 5 if nm.endswith("_ic") else 9
        ck.fields[nm] = rng.standard_normal((rows, lm)) + 1j * rng.standard_normal((rows, lm))
        ck.past[nm] = {k: [rng.standard_normal((rows, lm)) + 1j * rng.standard_normal((rows, lm)) for _ in range(n)]
                       for k, n in ck._levels()}
    for nm in ("domega_ic_dt", "domega_ma_dt"):
        ck.scalars_past[nm] = {k: rng.random(n) for k, n in ck._levels()}
    ck.rotation["omega_ic1"] = 3.5
    return ck


@pytest.mark.parametrize("family,mag,cond_ic", [("MULTISTEP", True, True), ("MULTISTEP", False, False), ("DIRK", True, False)])
def test_write_then_read_is_the_identity(tmp_path, family, mag, cond_ic):
    a = _synthetic(family=family, mag=mag, cond_ic=cond_ic)
    p = str(tmp_path / "checkpoint_end.test")
    write_checkpoint(p, a)
    b = read_checkpoint(p)
    assert (b.version, b.family, b.nexp, b.nimp, b.nold, b.n_time_step) == (5, family, a.nexp, a.nimp, a.nold, 4711)
    assert b.time == a.time and np.array_equal(b.dt, a.dt) and np.array_equal(b.r, a.r)
    assert b.params == a.params and b.trunc == a.trunc and (b.l_max, b.m_min, b.m_max) == (8, 0, 8)
    assert b.rotation == a.rotation and list(b.fields) == list(a.fields)
    for nm in a.fields:
        assert np.array_equal(b.fields[nm], a.fields[nm]), nm
        for k, n in a._levels():
            assert len(b.past[nm][k]) == n
            for x, y in zip(a.past[nm][k], b.past[nm][k]):
                assert np.array_equal(x, y), (nm, k)
    if family == "MULTISTEP":
        for nm in a.scalars_past:
            for k in ("expl", "impl", "old"):
                assert np.array_equal(a.scalars_past[nm][k], b.scalars_past[nm][k])
    # the byte count is what storeCheckPoints.f90 writes
    n_past = (a.nexp + a.nimp + a.nold - 3) if family == "MULTISTEP" else 0
    n_oc = sum(1 for nm in a.fields if not nm.endswith("_ic"))
    n_ic = len(a.fields) - n_oc
    header = 4 + 8 + 10 + 12 + 8 * len(np.atleast_1d(a.dt)) + 4 + 72 + 24 + 12 + 72 + 8 + 16 + 8 * 9 + 2 * 8 * n_past + 96 + 24
    assert os.path.getsize(p) == header + 16 * a.lm_max * (1 + n_past) * (9 * n_oc + 5 * n_ic)


def test_st_map_order_and_lm_max():
    ck = _synthetic(minc=2)
    l, m = ck.lm_maps()
    assert len(l) == ck.lm_max == sum(9 - mm for mm in range(0, 9, 2))
    assert list(m[:9]) == [0] * 9 and list(l[:9]) == list(range(9)) and (l[9], m[9]) == (2, 2)


def test_errors_are_loud(tmp_path):
    a = _synthetic()
    p = str(tmp_path / "c")
    a.fields["s"] = a.fields["s"][:, :-1]
    with pytest.raises(CheckpointError, match="shape"):
        write_checkpoint(p, a)
    a = _synthetic()
    del a.fields["aj"]
    with pytest.raises(CheckpointError, match="pairs"):
        write_checkpoint(p, a)
    a = _synthetic()
    a.past["w"]["expl"] = a.past["w"]["expl"][:1]
    with pytest.raises(CheckpointError, match="past"):
        write_checkpoint(p, a)
    a = _synthetic()
    write_checkpoint(p, a)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-8])
    with pytest.raises(CheckpointError, match="ends early"):
        read_checkpoint(p)
    open(p, "wb").write(raw + b"\0" * 8)
    with pytest.raises(CheckpointError, match="trailing"):
        read_checkpoint(p)
    open(p, "wb").write(np.array([1 << 24], "<i4").tobytes() + raw[4:])     # record marker / wrong endianness
    with pytest.raises(CheckpointError, match="version"):
        read_checkpoint(p)


@needs_reference
def test_reads_the_reference_full_sphere_checkpoint():
    """Version 2 (no stef, no l_max line, no phase-field flag, Lorentz-torque arrays), FD, CNAB2 with one past explicit level."""
    ck = read_checkpoint(os.path.join(REF, "full_sphere", "checkpoint_end.start"))
    g = np.load(os.path.join(HERE, "golden", "full_sphere_reference.npz"))
    assert (ck.version, ck.family, ck.nexp, ck.nimp, ck.nold) == (2, "MULTISTEP", 2, 1, 1)
    assert (ck.l_max, ck.m_max, ck.lm_max, ck.trunc["minc"]) == (32, 30, 198, 3)
    assert ck.rscheme == dict(version="fd", n=4, n2=2, ratio1=0.3, ratio2=0.2)
    assert ck.time == float(g["time"]) and np.array_equal(ck.r, g["radius"]) and np.array_equal(ck.dt, g["dt"])
    assert list(ck.fields) == ["w", "z", "s"] and "lorentz_torque_ic_dt" in ck.scalars_past
    for nm in ("w", "z", "s"):
        assert np.array_equal(ck.fields[nm], g[nm])
        assert np.array_equal(ck.past[nm]["expl"][0], g["d%sdt_expl2" % nm])
        assert ck.past[nm]["impl"] == [] and ck.past[nm]["old"] == []


@needs_reference
def test_reads_the_reference_boussBenchSat_checkpoint_and_rewrites_it(tmp_path):
    """Version 4 (Chebyshev, MHD with a conducting inner core); written back as version 5 and read again."""
    ck = read_checkpoint(os.path.join(REF, "boussBenchSat", "checkpoint_end.start"))
    g = np.load(os.path.join(HERE, "golden", "boussBenchSat_ckpt.npz"))
    assert ck.version == 4 and (ck.l_max, ck.trunc["minc"], ck.trunc["n_r_max"]) == (64, 4, 33)
    assert ck.rscheme["version"] == "cheb" and np.array_equal(ck.r, g["radius"])
    assert list(ck.fields) == ["w", "z", "p", "s", "b", "aj", "b_ic", "aj_ic"]
    for nm in ("w", "z", "s", "b", "aj"):
        assert np.array_equal(ck.fields[nm], g[nm]), nm
    assert ck.fields["b_ic"].shape == (ck.trunc["n_r_ic_max"], ck.lm_max)
    # continuity of the poloidal potential across the ICB (updateB.f90 kbotb = 3 matching condition)
    assert np.abs(ck.fields["b"][-1] - ck.fields["b_ic"][0]).max() < 1e-10 * np.abs(ck.fields["b"][-1]).max()
    p = str(tmp_path / "checkpoint_v5")
    write_checkpoint(p, ck)
    ck5 = read_checkpoint(p)
    assert ck5.version == 5 and ck5.params == ck.params and ck5.rotation == ck.rotation
    for nm in ck.fields:
        assert np.array_equal(ck5.fields[nm], ck.fields[nm])


@needs_reference
@pytest.mark.parametrize("sample", ["boussBenchSat", "full_sphere"])
def test_the_reference_reader_reads_what_the_writer_wrote(tmp_path, sample):
    """python/magic/checkpoint.py (the reference's own reader, run in place with its plotting-package imports stubbed:
    matplotlib is not installed) on a version-5 file written by write_checkpoint from the shipped older-version checkpoint."""
    import sys
    import types
    saved = {k: sys.modules.get(k) for k in ("magic", "magic.libmagic")}
    try:
        pkg = types.ModuleType("magic")
        pkg.__path__ = []
        lib = types.ModuleType("magic.libmagic")
        lib.chebgrid = lambda nr, a, b: 0.5 * ((a + b) / (b - a) + np.cos(np.pi * (1.0 - np.arange(nr + 1.0) / nr))) * (b - a)
        lib.fd_grid = lib.scanDir = lambda *a, **k: None
        sys.modules["magic"], sys.modules["magic.libmagic"] = pkg, lib
        mod = types.ModuleType("reference_checkpoint")
        exec(compile(open("/root/reference/python/magic/checkpoint.py").read(), "reference_checkpoint", "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ck = read_checkpoint(os.path.join(REF, sample, "checkpoint_end.start"))
    p = str(tmp_path / "checkpoint_v5")
    write_checkpoint(p, ck)
    r = mod.MagicCheckpoint(l_read=True, filename=p)
    assert r.version == 5 and r.time == ck.time and (r.l_max, r.m_max, r.lm_max) == (ck.l_max, ck.m_max, ck.lm_max)
    assert np.array_equal(r.radius, ck.r) and r.ra == ck.params["ra"] and r.omega_ic == ck.rotation["omega_ic1"]
    pairs = [("wpol", "w"), ("ztor", "z"), ("entropy", "s"), ("pre", "p"), ("bpol", "b"), ("btor", "aj"), ("bpol_ic", "b_ic"),
             ("btor_ic", "aj_ic")]
    for theirs, ours in pairs:
        if ours in ck.fields:
            assert np.array_equal(getattr(r, theirs), ck.fields[ours]), ours
