"""Development probe (torchrun): time pack / NCCL exchange / unpack of one 5-field container transpose."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magic_b200 import Sht, Transposer
from magic_b200.transpose import unique_id
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
box = [unique_id() if rank == 0 else None]; dist.broadcast_object_list(box, src=0)
l_max, n_r_max, nf = int(sys.argv[1]), int(sys.argv[2]), 5
sht = Sht(l_max, device_id=local)
tr = Transposer(sht, n_r_max, nf, rank=rank, n_procs=world, nccl_id=box[0])
ext = torch.cuda.ExternalStream(sht.stream, device=dev)
lm = torch.zeros(nf, n_r_max, tr.nlm_loc, dtype=torch.complex128, device=dev)
r = torch.zeros(nf, tr.nr_loc, sht.lm_max, dtype=torch.complex128, device=dev)
buf = torch.zeros(max(lm.numel(), r.numel()), dtype=torch.complex128, device=dev)
def timeit(fn, n=5):
    fn(); ext.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record(ext)
        for _ in range(n): fn()
        e1.record(ext)
    ext.synchronize()
    return e0.elapsed_time(e1) / n
t_full = timeit(lambda: tr.transp_lm2r_dev(lm.data_ptr(), r.data_ptr()))
t_pack = timeit(lambda: tr.pack_lm2r_dev(lm.data_ptr(), buf.data_ptr()))
t_unpack = timeit(lambda: tr.unpack_lm2r_dev(buf.data_ptr(), r.data_ptr()))
t_full2 = timeit(lambda: tr.transp_r2lm_dev(r.data_ptr(), lm.data_ptr()))
gb = lm.numel() * 16 / 1e9
if rank == 0:
    print(f"container {gb:.2f} GB/rank: lm2r full {t_full:.2f} ms (pack {t_pack:.2f}, unpack {t_unpack:.2f}, exchange ~{t_full - t_pack - t_unpack:.2f}); r2lm full {t_full2:.2f} ms; "
          f"off-rank GB {gb * (world - 1) / world:.2f} -> {gb * (world - 1) / world / max(t_full - t_pack - t_unpack, 1e-3) * 1e3:.0f} GB/s exchange")
tr.destroy_comm(); sht.finalize_sht(); dist.destroy_process_group()
