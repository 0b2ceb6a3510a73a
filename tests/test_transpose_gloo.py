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


def _worker_parts(rank, world, port, l_max, n_r_max, n_fields, level_chunk, ret):
    """The exchange sequence of magic_rloop_run_lm_dev on the host: all inbound parts in chunk order, then all outbound
    parts, every part an all-to-all restricted to the c-th level chunk of each rank (empty for ranks with fewer chunks)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from magic_b200.riter import level_chunks
    from magic_b200.transpose import get_blocks, lo_map
    lo2st, ls, le = lo_map(l_max, l_max, 1, world)
    rs, re = get_blocks(n_r_max, world)
    lm_max = len(lo2st)
    nlm = le[rank] - ls[rank] + 1
    nr = re[rank] - rs[rank] + 1
    chunks = [level_chunks(int(re[q] - rs[q] + 1), level_chunk) for q in range(world)]
    C = max(len(c[0]) for c in chunks)

    def part(q, c):  # (first global level index, count) of rank q's levels in part c
        st, sz = chunks[q]
        return (rs[q] - 1 + st[c], sz[c]) if c < len(st) else (re[q], 0)

    rng = np.random.default_rng(300 + rank)
    arr_LM = rng.standard_normal((n_fields, n_r_max, nlm)) + 1j * rng.standard_normal((n_fields, n_r_max, nlm))
    arr_R = np.zeros((n_fields, nr, lm_max), dtype=np.complex128)
    for c in range(C):  # lm2r parts
        send = [torch.from_numpy(np.ascontiguousarray(arr_LM[:, part(q, c)[0]:part(q, c)[0] + part(q, c)[1], :]).reshape(-1)) for q in range(world)]
        g0, n0 = part(rank, c)
        recv = [torch.empty(n_fields * n0 * (le[p] - ls[p] + 1), dtype=torch.complex128) for p in range(world)]
        _alltoallv(recv, send, rank, world)
        for p in range(world):
            seg = recv[p].numpy().reshape(n_fields, n0, le[p] - ls[p] + 1)
            arr_R[:, g0 - (rs[rank] - 1):g0 - (rs[rank] - 1) + n0, lo2st[ls[p] - 1:le[p]]] = seg
    back = np.zeros_like(arr_LM)
    for c in range(C):  # r2lm parts
        g0, n0 = part(rank, c)
        r0 = g0 - (rs[rank] - 1)
        send = [torch.from_numpy(np.ascontiguousarray(arr_R[:, r0:r0 + n0, lo2st[ls[p] - 1:le[p]]]).reshape(-1)) for p in range(world)]
        recv = [torch.empty(n_fields * part(q, c)[1] * nlm, dtype=torch.complex128) for q in range(world)]
        _alltoallv(recv, send, rank, world)
        for q in range(world):
            gq, nq = part(q, c)
            back[:, gq:gq + nq, :] = recv[q].numpy().reshape(n_fields, nq, nlm)
    ret[rank] = (arr_LM, arr_R, back, C, [len(c[0]) for c in chunks])
    dist.destroy_process_group()


def test_two_rank_chunk_parts_compose_to_the_full_transpose():
    """n_r_max=19 on 2 ranks with level_chunk=4: slabs of 9 and 10 levels -> 2 and 3 chunks, so rank 0 takes part in a third
    exchange in which it owns no levels."""
    from oracle.oracle import Oracle
    world, n_fields, l_max, n_r_max = 2, 2, 16, 19
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_parts, args=(world, _free_port(), l_max, n_r_max, n_fields, 4, ret), nprocs=world, join=True)
    assert ret[0][3] == 3 and ret[0][4] == [2, 3]
    o = Oracle(l_max)
    arr_LM = [ret[p][0] for p in range(world)]
    ref_R = o.transp_lm2r(world, n_r_max, arr_LM)
    for q in range(world):
        assert np.array_equal(ret[q][1], ref_R[q])
        assert np.array_equal(ret[q][2], ret[q][0])
