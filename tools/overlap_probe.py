"""Development probe: does concurrent PCIe traffic slow the radial-loop kernels?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from magic_b200 import RadialLoop, Sht
from magic_b200.workload import make_fields, make_params, make_radial
from magic_b200.riter import OUT_NAMES
l_max, n_lev = 1023, 32
s = Sht(l_max); p = make_params("mhd", 257); rad = make_radial(257, l_max, nRstart=2, nRstop=1 + n_lev)
f1 = make_fields("mhd", s.lm2l, s.lm2m, 1, 1)
dev = {k: torch.from_numpy(np.repeat(v, n_lev, axis=0)).cuda() for k, v in f1.items()}
rl = RadialLoop(s, p, rad, level_chunk=16)
outs = {k: torch.zeros(n_lev, s.lm_max, dtype=torch.complex128, device="cuda") for k in OUT_NAMES}
dtr = torch.zeros(n_lev, dtype=torch.float64, device="cuda"); dth = torch.zeros_like(dtr)
def run():
    rl.radialLoop_dev({k: v.data_ptr() for k, v in dev.items()}, {k: v.data_ptr() for k, v in outs.items()}, dtr.data_ptr(), dth.data_ptr())
    return rl.last_timing()["total"]
run(); print("alone", run(), run())
n = 1 << 27
hb = torch.empty(n, dtype=torch.complex128).pin_memory(); db = torch.empty(n, dtype=torch.complex128, device="cuda")
hb2 = torch.empty(n, dtype=torch.complex128).pin_memory(); db2 = torch.empty(n, dtype=torch.complex128, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    torch.cuda.synchronize()
    for _ in range(3):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): db.copy_(hb, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): hb2.copy_(db2, non_blocking=True)
    t = run()
    torch.cuda.synchronize()
    print("with", mode, "traffic:", t)
