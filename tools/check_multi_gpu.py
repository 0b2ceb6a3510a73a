"""Run under torchrun: checks the NCCL r<->LM redistribution against a host-side reference built from the
same global array on every rank (bit exact), then a distributed radial loop against the single-rank result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magic_b200 import RadialLoop, Sht, Transposer  # noqa: E402
from magic_b200.transpose import get_blocks, lo_map, unique_id  # noqa: E402
from magic_b200.workload import make_fields, make_params, make_radial  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    box = [unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    l_max, n_r_max, nf = 42, 19, 5
    sht = Sht(l_max, device_id=local)
    tr = Transposer(sht, n_r_max, nf, rank=rank, n_procs=world, nccl_id=box[0])
    lo2st, ls, le = lo_map(l_max, l_max, 1, world)
    rs, re = get_blocks(n_r_max, world)
    rng = np.random.default_rng(77)
    glob = rng.standard_normal((nf, n_r_max, sht.lm_max)) + 1j * rng.standard_normal((nf, n_r_max, sht.lm_max))  # st order
    mine_lm = np.ascontiguousarray(glob[:, :, lo2st[ls[rank] - 1:le[rank]]])
    want_r = np.ascontiguousarray(glob[:, rs[rank] - 1:re[rank], :])
    ext = torch.cuda.ExternalStream(sht.stream, device=dev)
    d_lm = torch.from_numpy(mine_lm).to(dev)
    d_r = torch.zeros(nf, tr.nr_loc, sht.lm_max, dtype=torch.complex128, device=dev)
    torch.cuda.synchronize()
    tr.transp_lm2r_dev(d_lm.data_ptr(), d_r.data_ptr())
    ext.synchronize()
    ok1 = np.array_equal(d_r.cpu().numpy(), want_r)
    back = torch.zeros_like(d_lm)
    tr.transp_r2lm_dev(d_r.data_ptr(), back.data_ptr())
    ext.synchronize()
    ok2 = np.array_equal(back.cpu().numpy(), mine_lm)
    # narrower container through the same communicator
    d_r3 = torch.zeros(3, tr.nr_loc, sht.lm_max, dtype=torch.complex128, device=dev)
    tr.transp_lm2r_dev_n(3, d_lm.data_ptr(), d_r3.data_ptr())
    ext.synchronize()
    ok3 = np.array_equal(d_r3.cpu().numpy(), want_r[:3])

    # distributed radial loop == single-rank radial loop on the same global fields (levels are independent)
    p = make_params("mhd", n_r_max)
    lm2l, lm2m = sht.lm2l, sht.lm2m
    gfields = make_fields("mhd", lm2l, lm2m, n_r_max, 99)
    rad = make_radial(n_r_max, l_max, nRstart=rs[rank], nRstop=re[rank])
    rl = RadialLoop(sht, p, rad)
    got = rl.radialLoop({k: v[rs[rank] - 1:re[rank]] for k, v in gfields.items()})
    rl.finalize()
    ok4 = True
    if rank == 0:
        rl1 = RadialLoop(sht, p, make_radial(n_r_max, l_max))
        ref = rl1.radialLoop(gfields)
        rl1.finalize()
        for nm in ["dwdt", "dzdt", "dsdt", "dbdt", "djdt", "dVxBhLM"]:
            ok4 = ok4 and np.array_equal(got[nm], ref[nm][rs[0] - 1:re[0]])
    # pipelined LM -> LM path (magic_rloop_run_lm_dev): bit-identical to lm2r -> radial loop -> r2lm done one after the other.
    # level_chunk=4 gives 2-3 chunks per rank (and, with n_r_max=19, ranks with different chunk counts at some world sizes).
    def lm_container(names):
        g = np.stack([gfields[n] for n in names])                     # [nf][n_r_max][lm_max] st order
        return torch.from_numpy(np.ascontiguousarray(g[:, :, lo2st[ls[rank] - 1:le[rank]]])).to(dev)
    flow_LM, s_LM, field_LM = lm_container(["w", "dw", "ddw", "z", "dz"]), lm_container(["s", "s"]), lm_container(["b", "db", "ddb", "aj", "dj"])
    rl = RadialLoop(sht, p, rad, level_chunk=4)
    zl = lambda n: torch.zeros(n, n_r_max, tr.nlm_loc, dtype=torch.complex128, device=dev)
    zr = lambda n: torch.zeros(n, tr.nr_loc, sht.lm_max, dtype=torch.complex128, device=dev)
    dtr, dth = torch.zeros(tr.nr_loc, dtype=torch.float64, device=dev), torch.zeros(tr.nr_loc, dtype=torch.float64, device=dev)
    # (a) sequential reference
    fR, sR, bR, dfR, dsR, dbR = zr(5), zr(2), zr(5), zr(3), zr(2), zr(3)
    torch.cuda.synchronize()
    tr.transp_lm2r_dev_n(5, flow_LM.data_ptr(), fR.data_ptr()); tr.transp_lm2r_dev_n(2, s_LM.data_ptr(), sR.data_ptr())
    tr.transp_lm2r_dev_n(5, field_LM.data_ptr(), bR.data_ptr())
    fin = {"w": fR[0], "dw": fR[1], "ddw": fR[2], "z": fR[3], "dz": fR[4], "s": sR[0], "b": bR[0], "db": bR[1], "ddb": bR[2], "aj": bR[3], "dj": bR[4]}
    fout = {"dwdt": dfR[0], "dzdt": dfR[1], "dpdt": dfR[2], "dsdt": dsR[0], "dVSrLM": dsR[1], "dbdt": dbR[0], "djdt": dbR[1], "dVxBhLM": dbR[2]}
    rl.radialLoop_dev({k: v.data_ptr() for k, v in fin.items()}, {k: v.data_ptr() for k, v in fout.items()}, dtr.data_ptr(), dth.data_ptr())
    ref_LM = [zl(3), zl(2), zl(3)]
    tr.transp_r2lm_dev_n(3, dfR.data_ptr(), ref_LM[0].data_ptr()); tr.transp_r2lm_dev_n(2, dsR.data_ptr(), ref_LM[1].data_ptr())
    tr.transp_r2lm_dev_n(3, dbR.data_ptr(), ref_LM[2].data_ptr())
    ext.synchronize()
    dtr_ref = dtr.clone()
    # (b) pipelined, twice (buffers and events are reused from step to step)
    got_LM = [zl(3), zl(2), zl(3)]
    ok5 = True
    for _ in range(2):
        for g in got_LM:
            g.zero_()
        dtr.zero_()
        torch.cuda.synchronize()
        rl.run_lm_dev(tr, flow_LM.data_ptr(), s_LM.data_ptr(), field_LM.data_ptr(), got_LM[0].data_ptr(), got_LM[1].data_ptr(),
                      got_LM[2].data_ptr(), dtr.data_ptr(), dth.data_ptr())
        ext.synchronize()
        torch.cuda.synchronize()
        for a, b in zip(got_LM, ref_LM):
            ok5 = ok5 and bool(torch.equal(a, b))
        ok5 = ok5 and bool(torch.equal(dtr, dtr_ref)) and float(ref_LM[0].abs().sum()) > 0
    # (c) the host-pointer call magic_rloop_run_lm: host LM containers in, host LM explicit terms out (PCIe transfers pipelined
    #     with the transposes and the compute), twice; bit-identical to (a)
    h_in = {"flow": flow_LM.cpu().numpy(), "s": s_LM.cpu().numpy(), "field": field_LM.cpu().numpy()}
    ok6 = True
    for _ in range(2):
        h_out = {"dflowdt": np.zeros((3, n_r_max, tr.nlm_loc), dtype=np.complex128), "dsdt": np.zeros((2, n_r_max, tr.nlm_loc), dtype=np.complex128),
                 "dbdt": np.zeros((3, n_r_max, tr.nlm_loc), dtype=np.complex128)}
        h_dtr, h_dth = np.zeros(tr.nr_loc), np.zeros(tr.nr_loc)
        rl.run_lm(tr, h_in, h_out, h_dtr, h_dth)
        for a, b in zip([h_out["dflowdt"], h_out["dsdt"], h_out["dbdt"]], ref_LM):
            ok6 = ok6 and np.array_equal(a, b.cpu().numpy())
        ok6 = ok6 and np.array_equal(h_dtr, dtr_ref.cpu().numpy())
    rl.finalize()
    # (d) another field set through the same pipeline: Boussinesq hydro in the double-curl formulation (4-field dflowdt, no magnetic
    #     containers), device containers, against the sequential calls
    p2 = make_params("hydro", n_r_max)
    p2.l_double_curl = 1
    rl = RadialLoop(sht, p2, rad, level_chunk=4)
    dfR4, dsR2 = zr(4), zr(2)
    torch.cuda.synchronize()
    fout2 = {"dwdt": dfR4[0], "dzdt": dfR4[1], "dVxVhLM": dfR4[3], "dsdt": dsR2[0], "dVSrLM": dsR2[1]}
    fin2 = {k: fin[k] for k in ["w", "dw", "ddw", "z", "dz", "s"]}
    rl.radialLoop_dev({k: v.data_ptr() for k, v in fin2.items()}, {k: v.data_ptr() for k, v in fout2.items()}, dtr.data_ptr(), dth.data_ptr())
    ref2 = [zl(4), zl(2)]
    tr.transp_r2lm_dev_n(4, dfR4.data_ptr(), ref2[0].data_ptr()); tr.transp_r2lm_dev_n(2, dsR2.data_ptr(), ref2[1].data_ptr())
    ext.synchronize()
    got2 = [zl(4), zl(2)]
    torch.cuda.synchronize()
    rl.run_lm_dev(tr, flow_LM.data_ptr(), s_LM.data_ptr(), 0, got2[0].data_ptr(), got2[1].data_ptr(), 0, dtr.data_ptr(), dth.data_ptr())
    ext.synchronize()
    torch.cuda.synchronize()
    ok7 = all(bool(torch.equal(a, b)) for a, b in zip(got2, ref2)) and float(ref2[0][3].abs().sum()) > 0
    rl.finalize()
    flags = torch.tensor([ok1, ok2, ok3, ok4, ok5, ok6, ok7], dtype=torch.int32, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if bool(flags.min().item()) else "FAIL", flags.tolist(), "world", world, flush=True)
    tr.destroy_comm()
    sht.finalize_sht()
    dist.destroy_process_group()
    sys.exit(0 if bool(flags.min().item()) else 1)


if __name__ == "__main__":
    main()
