"""CPU, world_size 2 over gloo: the N>1 host logic of the r<->LM redistribution.

Each rank packs its LM slab with the library's lo_map / getBlocks (exactly what the CUDA pack kernels index
with), exchanges with isend/irecv over gloo, unpacks with the lo->st permutation, and the
result is compared with the oracle's in-process emulation of type_mpiatoav (mpi_transpose.f90:307-359,444-530).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _alltoallv(recv, send, rank, world):
    """Point-to-point all-to-all (the type_mpiptop flavour, mpi_transpose.f90:752-814); gloo has no alltoall."""
    reqs = []
    for p in range(world):
        if p == rank:
            recv[p].copy_(send[p])
            continue
        reqs.append(dist.isend(torch.view_as_real(send[p]).contiguous(), dst=p))
        reqs.append(dist.irecv(torch.view_as_real(recv[p]), src=p))
    for r in reqs:
        r.wait()


def _worker(rank, world, port, l_max, minc, n_r_max, n_fields, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from magic_b200.transpose import get_blocks, lo_map
    m_max = (l_max // minc) * minc
    lo2st, ls, le = lo_map(l_max, m_max, minc, world)
    rs, re = get_blocks(n_r_max, world)
    lm_max = len(lo2st)
    nlm = le[rank] - ls[rank] + 1
    nr = re[rank] - rs[rank] + 1
    rng = np.random.default_rng(100 + rank)
    arr_LM = rng.standard_normal((n_fields, n_r_max, nlm)) + 1j * rng.standard_normal((n_fields, n_r_max, nlm))
    # pack (mpi_transpose.f90:320-333): segment q = [f][n_r in block q][lm in my slab]
    send = [torch.from_numpy(np.ascontiguousarray(arr_LM[:, rs[q] - 1:re[q], :]).reshape(-1)) for q in range(world)]
    recv = [torch.empty(n_fields * nr * (le[p] - ls[p] + 1), dtype=torch.complex128) for p in range(world)]
    _alltoallv(recv, send, rank, world)
    arr_R = np.zeros((n_fields, nr, lm_max), dtype=np.complex128)
    for p in range(world):  # unpack with the lo->st permutation (:341-357)
        seg = recv[p].numpy().reshape(n_fields, nr, le[p] - ls[p] + 1)
        arr_R[:, :, lo2st[ls[p] - 1:le[p]]] = seg
    # and back (r2lm, :490-528)
    send = [torch.from_numpy(np.ascontiguousarray(arr_R[:, :, lo2st[ls[p] - 1:le[p]]]).reshape(-1)) for p in range(world)]
    recv = [torch.empty(n_fields * (re[q] - rs[q] + 1) * nlm, dtype=torch.complex128) for q in range(world)]
    _alltoallv(recv, send, rank, world)
    back = np.zeros_like(arr_LM)
    for q in range(world):
        back[:, rs[q] - 1:re[q], :] = recv[q].numpy().reshape(n_fields, re[q] - rs[q] + 1, nlm)
    ret[rank] = (arr_LM, arr_R, back)
    dist.destroy_process_group()


@pytest.mark.parametrize("l_max,minc,n_r_max", [(16, 1, 9), (32, 3, 7)])
def test_two_rank_transpose_matches_oracle(l_max, minc, n_r_max):
    from oracle.oracle import Oracle
    world, n_fields = 2, 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), l_max, minc, n_r_max, n_fields, ret), nprocs=world, join=True)
    m_max = (l_max // minc) * minc
    o = Oracle(l_max, minc=minc, n_theta=4 * ((3 * l_max // 2 + 3) // 4) + 4, n_phi=max(8, 4 * ((2 * (l_max // minc) + 8) // 4)), m_max=m_max)
    arr_LM = [ret[p][0] for p in range(world)]
    ref_R = o.transp_lm2r(world, n_r_max, arr_LM)
    for q in range(world):
        assert np.array_equal(ret[q][1], ref_R[q])      # bit exact: pure data movement
        assert np.array_equal(ret[q][2], ret[q][0])     # r2lm(lm2r(x)) == x
    ref_LM = o.transp_r2lm(world, n_r_max, ref_R)
    for p in range(world):
        assert np.array_equal(ref_LM[p], arr_LM[p])
