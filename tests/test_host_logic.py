"""CPU: host-side logic of the product (grid sizing, decompositions, lo_map) against the oracle's restatement."""
import numpy as np
import pytest

from oracle.oracle import Oracle, get_blocks as orc_get_blocks, grid_sizes as orc_grid_sizes


@pytest.mark.parametrize("kw", [dict(l_max=16), dict(l_max=85), dict(l_max=255), dict(l_max=511), dict(l_max=1023),
                                dict(n_phi_tot=288), dict(n_phi_tot=96, minc=3), dict(n_phi_tot=1024, minc=4),
                                dict(l_max=64, minc=4), dict(l_max=213), dict(l_max=426)])
def test_grid_sizes_match_oracle(kw):
    from magic_b200 import grid_sizes
    assert grid_sizes(**kw) == orc_grid_sizes(**kw)


@pytest.mark.parametrize("n,p", [(257, 8), (121, 4), (33, 2), (33, 1), (97, 7), (12, 12)])
def test_get_blocks_match_oracle(n, p):
    from magic_b200 import get_blocks
    from magic_b200.lib import check, load_library
    from ctypes import c_int
    s, e = get_blocks(n, p)
    so, eo = orc_get_blocks(n, p)
    assert np.array_equal(s, so) and np.array_equal(e, eo)
    lib = load_library()
    cs, ce = (c_int * p)(), (c_int * p)()
    check(lib.magic_get_blocks(c_int(n), c_int(p), cs, ce))
    assert list(cs) == list(so) and list(ce) == list(eo)


@pytest.mark.parametrize("l_max,minc,procs", [(16, 1, (1, 2, 3, 4, 8, 9, 16)), (32, 3, (1, 2, 5, 16, 17)), (96, 1, (2, 8, 48, 49))])
def test_lo_map_matches_oracle(l_max, minc, procs):
    from magic_b200.transpose import lo_map
    m_max = (l_max // minc) * minc
    o = Oracle(l_max, minc=minc, n_theta=4 * ((3 * l_max // 2 + 3) // 4) + 4, n_phi=max(8, 4 * ((2 * (l_max // minc) + 8) // 4)), m_max=m_max)
    for n_procs in procs:
        a, s, e = lo_map(l_max, m_max, minc, n_procs)
        b, so, eo = o.lo_map(n_procs)
        assert np.array_equal(a, b), n_procs
        assert np.array_equal(s, so) and np.array_equal(e, eo), n_procs


def test_workload_shapes_and_seeding():
    from magic_b200.workload import config_sizes, make_fields, make_params, make_radial, seed_for
    gs = config_sizes("dynamo_l1023")
    assert (gs["n_theta_max"], gs["n_phi_max"], gs["lm_max"], gs["n_r_max"]) == (1536, 3072, 524800, 257)
    gs = config_sizes("hydro_bench_anel")
    assert (gs["l_max"], gs["n_theta_max"], gs["lm_max"]) == (96, 144, 4753)
    assert seed_for(2, 3) == 20261017 + 2003
    rad = make_radial(33, 16, nRstart=5, nRstop=9)
    assert list(rad["nR"]) == [5, 6, 7, 8, 9] and np.all(np.diff(rad["r"]) < 0)
    full = make_radial(33, 16)
    assert np.isclose(full["r"][0], 20 / 13) and np.isclose(full["r"][-1], 7 / 13)
    o = Oracle(16)
    f1 = make_fields("mhd", o.lm2l, o.lm2m, 3, 7)
    f2 = make_fields("mhd", o.lm2l, o.lm2m, 3, 7)
    assert all(np.array_equal(f1[k], f2[k]) for k in f1)
    assert np.all(f1["w"][:, o.lm2m == 0].imag == 0) and np.all(f1["w"][:, 0] == 0)
    p = make_params("mhd", 33)
    assert p.l_mag == 1 and p.LFfac == pytest.approx(200.0) and p.n_r_max == 33


def test_level_chunks_partition_every_slab():
    """magic_level_chunks (host only): chunks tile the slab, all but the last have exactly level_chunk levels, a small
    remainder is folded into the last chunk -- and the ranks of a getBlocks decomposition get chunk counts that differ by
    at most one (magic_rloop_run_lm_dev pads the shorter lists with empty parts)."""
    from magic_b200.riter import level_chunks
    from magic_b200.transpose import get_blocks
    for n in list(range(1, 70)) + [128, 129, 257]:
        for lc in (1, 4, 16, 32):
            st, sz = level_chunks(n, lc)
            assert st[0] == 0 and sum(sz) == n and all(a + b == c for a, b, c in zip(st, sz, st[1:] + [n]))
            eff = min(lc, n)
            assert all(z == eff for z in sz[:-1])
            assert 1 <= sz[-1] <= eff + eff // 4
    assert level_chunks(257, 16)[1] == [16] * 15 + [17]
    assert level_chunks(33, 16)[1] == [16, 17] and level_chunks(32, 16)[1] == [16, 16]
    for n_r, n_procs in ((257, 8), (257, 2), (121, 4), (19, 2)):
        rs, re = get_blocks(n_r, n_procs)
        counts = [len(level_chunks(int(e - s + 1), 16 if n_r > 19 else 4)[0]) for s, e in zip(rs, re)]
        assert max(counts) - min(counts) <= 1, counts
