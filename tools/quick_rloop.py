"""Quick device-resident timing of the radial loop (development tool; bench.py is the contract)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from magic_b200 import RadialLoop, Sht, grid_sizes
from magic_b200.workload import make_fields, make_params, make_radial
from magic_b200.riter import IN_NAMES, OUT_NAMES

l_max = int(sys.argv[1]); n_lev = int(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
gs = grid_sizes(l_max=l_max)
s = Sht(l_max)
n_r_max = 257
p = make_params("mhd", n_r_max)
rad = make_radial(n_r_max, l_max, nRstart=2, nRstop=1 + n_lev)
t0 = time.time()
fields = make_fields("mhd", s.lm2l, s.lm2m, 1, 1)
dev = {k: torch.from_numpy(np.repeat(v, n_lev, axis=0)).cuda() for k, v in fields.items()}
rl = RadialLoop(s, p, rad, level_chunk=chunk)
outs = {k: torch.zeros(n_lev, s.lm_max, dtype=torch.complex128, device="cuda") for k in OUT_NAMES}
dtr = torch.zeros(n_lev, dtype=torch.float64, device="cuda"); dth = torch.zeros_like(dtr)
print("setup s", time.time() - t0)
for it in range(4):
    rl.radialLoop_dev({k: v.data_ptr() for k, v in dev.items()}, {k: v.data_ptr() for k, v in outs.items()}, dtr.data_ptr(), dth.data_ptr())
    t = rl.last_timing()
    fl = rl.legendre_flops()
    leg = t["legendre_syn"] + t["legendre_an"]
    print(json.dumps({k: round(v, 3) for k, v in t.items()}), "legendre TF/s", round(fl / leg * 1e-9, 2), "overall TF/s", round(fl / t["total"] * 1e-9, 2))
import hashlib
torch.cuda.synchronize()
h = hashlib.sha256()
for k in sorted(outs):
    h.update(outs[k].cpu().numpy().tobytes())
print("sha256 of outputs", h.hexdigest()[:16], "dtrkc", float(dtr.min()), "dthkc", float(dth.min()))
