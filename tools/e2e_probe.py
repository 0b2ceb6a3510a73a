"""Development probe: where does the host-pointer radial loop spend its time (copies vs compute)?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from magic_b200 import RadialLoop, Sht
from magic_b200.workload import make_fields, make_params, make_radial
from magic_b200.riter import OUT_NAMES

l_max, n_lev, chunk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
s = Sht(l_max)
p = make_params("mhd", 257)
rad = make_radial(257, l_max, nRstart=2, nRstop=1 + n_lev)
f1 = make_fields("mhd", s.lm2l, s.lm2m, 1, 1)
host_in = {k: torch.from_numpy(np.repeat(v, n_lev, axis=0)).pin_memory() for k, v in f1.items()}
host_out = {k: torch.empty(n_lev, s.lm_max, dtype=torch.complex128).pin_memory() for k in OUT_NAMES}
np_in = {k: v.numpy() for k, v in host_in.items()}
np_out = {k: v.numpy() for k, v in host_out.items()}
np_out["dtrkc"] = np.zeros(n_lev); np_out["dthkc"] = np.zeros(n_lev)
rl = RadialLoop(s, p, rad, level_chunk=chunk)
for it in range(3):
    t0 = time.perf_counter(); rl.radialLoop(np_in, out=np_out); t1 = time.perf_counter()
    print("host call ms", (t1 - t0) * 1e3, "device stages total ms", rl.last_timing()["total"])
nbytes_in = sum(v.nbytes for v in np_in.values()); nbytes_out = 8 * n_lev * s.lm_max * 16
d = torch.empty(nbytes_in // 16, dtype=torch.complex128, device="cuda")
big = torch.empty(nbytes_in // 16, dtype=torch.complex128).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(big, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("pure H2D GB/s", nbytes_in / (t1 - t0) * 1e-9, "ms", (t1 - t0) * 1e3)
t0 = time.perf_counter(); big.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("pure D2H GB/s", nbytes_in / (t1 - t0) * 1e-9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
big2 = torch.empty(nbytes_in // 16, dtype=torch.complex128).pin_memory(); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(big, non_blocking=True)
with torch.cuda.stream(s2): big2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("duplex: each direction GB/s", nbytes_in / (t1 - t0) * 1e-9)
