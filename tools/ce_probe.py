"""torchrun probe of the copy-engine transposer exchange on a small problem (development tool)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magic_b200 import Sht, Transposer
from magic_b200.transpose import get_blocks, lo_map, unique_id
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
glob = rng.standard_normal((nf, n_r_max, sht.lm_max)) + 1j * rng.standard_normal((nf, n_r_max, sht.lm_max))
mine_lm = np.ascontiguousarray(glob[:, :, lo2st[ls[rank] - 1:le[rank]]])
want_r = np.ascontiguousarray(glob[:, rs[rank] - 1:re[rank], :])
ext = torch.cuda.ExternalStream(sht.stream, device=dev)
d_lm = torch.from_numpy(mine_lm).to(dev)
d_r = torch.zeros(nf, tr.nr_loc, sht.lm_max, dtype=torch.complex128, device=dev)
torch.cuda.synchronize()
for it in range(5):
    d_r.zero_()
    tr.transp_lm2r_dev(d_lm.data_ptr(), d_r.data_ptr())
    ext.synchronize()
    ok1 = np.array_equal(d_r.cpu().numpy(), want_r)
    back = torch.zeros_like(d_lm)
    tr.transp_r2lm_dev(d_r.data_ptr(), back.data_ptr())
    ext.synchronize()
    ok2 = np.array_equal(back.cpu().numpy(), mine_lm)
    print(f"rank {rank} iter {it}: lm2r {ok1} r2lm {ok2}", flush=True)
tr.destroy_comm(); sht.finalize_sht(); dist.destroy_process_group()
