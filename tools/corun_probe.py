"""Does the block scheduler overlap the HBM-bound kernels of one level chunk with the Legendre GEMM of another?
Two independent handles (own stream, own workspace) run the same slab from two host threads; compare the wall/device time
of both running together with twice the time of one running alone.  Development probe, not a bench line."""
import sys, time, threading, json
import numpy as np, torch
sys.path.insert(0, ".")
from magic_b200 import RadialLoop, Sht
from magic_b200.workload import make_fields, make_params, make_radial
from magic_b200.riter import OUT_NAMES

l_max = int(sys.argv[1]) if len(sys.argv) > 1 else 1023
n_lev = int(sys.argv[2]) if len(sys.argv) > 2 else 32
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n_r_max = 257
p = make_params("mhd", n_r_max)
rad = make_radial(n_r_max, l_max, nRstart=2, nRstop=1 + n_lev)


class One:
    def __init__(self):
        self.s = Sht(l_max)
        f = make_fields("mhd", self.s.lm2l, self.s.lm2m, 1, 1)
        self.dev = {k: torch.from_numpy(np.repeat(v, n_lev, axis=0)).cuda() for k, v in f.items()}
        self.rl = RadialLoop(self.s, p, rad, level_chunk=chunk)
        self.outs = {k: torch.zeros(n_lev, self.s.lm_max, dtype=torch.complex128, device="cuda") for k in OUT_NAMES}
        self.dtr = torch.zeros(n_lev, dtype=torch.float64, device="cuda")
        self.dth = torch.zeros_like(self.dtr)

    def run(self, reps):
        for _ in range(reps):
            self.rl.radialLoop_dev({k: v.data_ptr() for k, v in self.dev.items()}, {k: v.data_ptr() for k, v in self.outs.items()},
                                   self.dtr.data_ptr(), self.dth.data_ptr())
        self.rl.sync()


a, b = One(), One()
a.run(1); b.run(1)
torch.cuda.synchronize()
reps = 3
t0 = time.time(); a.run(reps); t_alone = (time.time() - t0) / reps
t0 = time.time(); b.run(reps); t_alone_b = (time.time() - t0) / reps
ths = [threading.Thread(target=x.run, args=(reps,)) for x in (a, b)]
t0 = time.time()
for t in ths: t.start()
for t in ths: t.join()
t_both = (time.time() - t0) / reps
print(json.dumps({"l_max": l_max, "levels_each": n_lev, "chunk": chunk, "alone_ms": [round(1e3 * t_alone, 2), round(1e3 * t_alone_b, 2)],
                  "both_ms": round(1e3 * t_both, 2), "gain_vs_serial": round((t_alone + t_alone_b) / t_both, 3),
                  "stages_alone": {k: round(v, 2) for k, v in a.rl.last_timing().items()}}))
