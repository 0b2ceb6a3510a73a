"""GPU parity of the r<->LM redistribution kernels (bit exact: pure data movement).

n_procs ranks are emulated in one process: each rank's pack kernel fills its send buffer, the exchange is done by
slicing with the library's counts/displacements (what ncclSend/ncclRecv move), each rank's unpack kernel applies the
lo->st permutation; result vs the oracle's emulation of type_mpiatoav (mpi_transpose.f90:307-359,444-530).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("l_max,minc,n_procs,n_r_max,n_fields", [(16, 1, 1, 5, 2), (16, 1, 3, 10, 5), (32, 3, 2, 7, 3), (21, 1, 12, 13, 2)])
def test_pack_exchange_unpack_matches_oracle(l_max, minc, n_procs, n_r_max, n_fields):
    import torch
    from magic_b200 import Sht, Transposer, grid_sizes
    from oracle.oracle import Oracle
    gs = grid_sizes(l_max=l_max, minc=minc)
    o = Oracle(gs["l_max"], minc=minc, n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"])
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=minc, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    trs = [Transposer(s, n_r_max, n_fields, rank=p, n_procs=n_procs) for p in range(n_procs)]
    rng = np.random.default_rng(5)
    arr_LM = [rng.standard_normal((n_fields, n_r_max, t.nlm_loc)) + 1j * rng.standard_normal((n_fields, n_r_max, t.nlm_loc)) for t in trs]
    ref_R = o.transp_lm2r(n_procs, n_r_max, arr_LM)
    dev = torch.device("cuda")
    # extents agree with the oracle's decomposition
    _, ls, le = o.lo_map(n_procs)
    for p, t in enumerate(trs):
        assert (t.llm, t.ulm) == (ls[p], le[p])
    if n_procs == 1:
        got = trs[0].transp_lm2r(arr_LM[0])
        assert np.array_equal(got, ref_R[0])
        assert np.array_equal(trs[0].transp_r2lm(got), arr_LM[0])
    else:
        def exchange(direction, send):
            recv = []
            for q, t in enumerate(trs):
                sc, sd, rc, rd = t.counts(direction)
                recv.append(torch.zeros(int(rc.sum()), dtype=torch.complex128, device=dev))
            for p, t in enumerate(trs):
                sc, sd, _, _ = t.counts(direction)
                for q in range(n_procs):
                    _, _, rc, rd = trs[q].counts(direction)
                    assert sc[q] == rc[p]
                    recv[q][rd[p]:rd[p] + rc[p]] = send[p][sd[q]:sd[q] + sc[q]]
            return recv
        d_LM = [torch.from_numpy(a).to(dev) for a in arr_LM]
        send = []
        for p, t in enumerate(trs):
            b = torch.zeros(d_LM[p].numel(), dtype=torch.complex128, device=dev)
            t.pack_lm2r_dev(d_LM[p].data_ptr(), b.data_ptr())
            send.append(b)
        # the library works on its own stream: wait for it before torch touches the buffers
        torch.cuda.synchronize()
        torch.cuda.ExternalStream(s.stream).synchronize()
        recv = exchange(0, send)
        d_R = []
        for q, t in enumerate(trs):
            r = torch.zeros(n_fields, t.nr_loc, s.lm_max, dtype=torch.complex128, device=dev)
            t.unpack_lm2r_dev(recv[q].data_ptr(), r.data_ptr())
            d_R.append(r)
        torch.cuda.ExternalStream(s.stream).synchronize()
        for q in range(n_procs):
            assert np.array_equal(d_R[q].cpu().numpy(), ref_R[q]), q
        # and back
        send = []
        for q, t in enumerate(trs):
            b = torch.zeros(d_R[q].numel(), dtype=torch.complex128, device=dev)
            t.pack_r2lm_dev(d_R[q].data_ptr(), b.data_ptr())
            send.append(b)
        torch.cuda.ExternalStream(s.stream).synchronize()
        recv = exchange(1, send)
        for p, t in enumerate(trs):
            back = torch.zeros_like(d_LM[p])
            t.unpack_r2lm_dev(recv[p].data_ptr(), back.data_ptr())
            torch.cuda.ExternalStream(s.stream).synchronize()
            assert np.array_equal(back.cpu().numpy(), arr_LM[p]), p
    for t in trs:
        t.destroy_comm()
    s.finalize_sht()


@pytest.mark.parametrize("l_max,minc,n_procs,n_r_max,n_fields,nparts", [(16, 1, 3, 10, 5, 2), (32, 3, 2, 7, 3, 3), (21, 1, 4, 13, 2, 2)])
def test_level_parts_compose_to_the_full_transpose(l_max, minc, n_procs, n_r_max, n_fields, nparts):
    """magic_transp_create_part: the level slabs of every rank are cut into `nparts` pieces (uneven, one of them possibly
    empty); running the pieces one after the other -- pack, emulated exchange, unpack -- must reproduce the full lm2r and
    r2lm bit for bit.  This is what magic_rloop_run_lm_dev overlaps with the compute of the level chunks."""
    import torch
    from magic_b200 import Sht, Transposer, grid_sizes
    from oracle.oracle import Oracle
    gs = grid_sizes(l_max=l_max, minc=minc)
    o = Oracle(gs["l_max"], minc=minc, n_theta=gs["n_theta_max"], n_phi=gs["n_phi_max"], m_max=gs["m_max"])
    s = Sht(gs["l_max"], m_max=gs["m_max"], minc=minc, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
    trs = [Transposer(s, n_r_max, n_fields, rank=p, n_procs=n_procs) for p in range(n_procs)]
    rng = np.random.default_rng(9)
    arr_LM = [rng.standard_normal((n_fields, n_r_max, t.nlm_loc)) + 1j * rng.standard_normal((n_fields, n_r_max, t.nlm_loc)) for t in trs]
    ref_R = o.transp_lm2r(n_procs, n_r_max, arr_LM)
    dev = torch.device("cuda")
    # uneven cut of every slab: rank q gets pieces of sizes ~nr/nparts, the last rank's last piece may be empty
    offs, cnts = [], []
    for c in range(nparts):
        off, cnt = [], []
        for q, t in enumerate(trs):
            edges = np.linspace(0, t.nr_loc, nparts + 1).astype(int)
            if q == n_procs - 1:
                edges[-2] = edges[-1]  # last piece of the last rank is empty
            off.append(edges[c]); cnt.append(edges[c + 1] - edges[c])
        offs.append(off); cnts.append(cnt)
    parts = [[t.part(offs[c], cnts[c]) for t in trs] for c in range(nparts)]
    sync = lambda: (torch.cuda.synchronize(), torch.cuda.ExternalStream(s.stream).synchronize())

    def exchange(ts, direction, send):
        recv = [torch.zeros(max(1, int(t.counts(direction)[2].sum())), dtype=torch.complex128, device=dev) for t in ts]
        for p, t in enumerate(ts):
            sc, sd, _, _ = t.counts(direction)
            for q in range(n_procs):
                _, _, rc, rd = ts[q].counts(direction)
                assert sc[q] == rc[p]
                recv[q][rd[p]:rd[p] + rc[p]] = send[p][sd[q]:sd[q] + sc[q]]
        return recv

    d_LM = [torch.from_numpy(a).to(dev) for a in arr_LM]
    d_R = [torch.zeros(n_fields, t.nr_loc, s.lm_max, dtype=torch.complex128, device=dev) for t in trs]
    for c in range(nparts):
        send = []
        for p, t in enumerate(parts[c]):
            b = torch.zeros(d_LM[p].numel(), dtype=torch.complex128, device=dev)
            t.pack_lm2r_dev(d_LM[p].data_ptr(), b.data_ptr())
            send.append(b)
        sync()
        recv = exchange(parts[c], 0, send)
        for q, t in enumerate(parts[c]):
            t.unpack_lm2r_dev(recv[q].data_ptr(), d_R[q].data_ptr())
        sync()
    for q in range(n_procs):
        assert np.array_equal(d_R[q].cpu().numpy(), ref_R[q]), q
    back = [torch.zeros_like(x) for x in d_LM]
    for c in range(nparts):
        send = []
        for q, t in enumerate(parts[c]):
            b = torch.zeros(d_R[q].numel(), dtype=torch.complex128, device=dev)
            t.pack_r2lm_dev(d_R[q].data_ptr(), b.data_ptr())
            send.append(b)
        sync()
        recv = exchange(parts[c], 1, send)
        for p, t in enumerate(parts[c]):
            t.unpack_r2lm_dev(recv[p].data_ptr(), back[p].data_ptr())
        sync()
    for p in range(n_procs):
        assert np.array_equal(back[p].cpu().numpy(), arr_LM[p]), p
    for row in parts:
        for t in row:
            t.destroy_comm()
    for t in trs:
        t.destroy_comm()
    s.finalize_sht()
