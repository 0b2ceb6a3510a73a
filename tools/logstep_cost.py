"""Cost of the log-step batches next to a regular pass of the radial loop, device-resident inputs (development tool).
usage: python tools/logstep_cost.py <l_max> <n_levels>"""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from magic_b200 import RadialLoop, Sht
from magic_b200.workload import make_fields, make_params, make_radial
from magic_b200.riter import OUT_NAMES, DIAG_HEL, DIAG_HEMI, DIAG_POWER, DIAG_PERPPAR, DIAG_FLUX, DIAG_VISCBC

l_max = int(sys.argv[1]); n_lev = int(sys.argv[2])
s = Sht(l_max)
n_r_max = 257
p = make_params("mhd", n_r_max)
rad = make_radial(n_r_max, l_max, nRstart=2, nRstop=1 + n_lev)
fields = make_fields("mhd", s.lm2l, s.lm2m, 1, 1)
fields["p"] = 0.5 * fields["s"] + 0.1 * fields["w"]
fields["ds"] = 0.7 * fields["s"]
dev = {k: torch.from_numpy(np.repeat(v, n_lev, axis=0)).cuda() for k, v in fields.items()}
ptr = {k: v.data_ptr() for k, v in dev.items()}
rl = RadialLoop(s, p, rad)
outs = {k: torch.zeros(n_lev, s.lm_max, dtype=torch.complex128, device="cuda") for k in OUT_NAMES}
dtr = torch.zeros(n_lev, dtype=torch.float64, device="cuda"); dth = torch.zeros_like(dtr)
free0 = torch.cuda.mem_get_info()[0]


def timed(label, fn, reps=2):
    res = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    out = res if isinstance(res, np.ndarray) else None
    print(json.dumps({"batch": label, "ms_second_call": round(dt, 2), "result_MB": round(out.nbytes / 1e6, 1) if out is not None else 0,
                      "finite": bool(np.isfinite(out).all()) if out is not None else None,
                      "free_GB": round(torch.cuda.mem_get_info()[0] / 1e9, 1)}))
    return res


timed("radial loop", lambda: rl.radialLoop_dev(ptr, {k: v.data_ptr() for k, v in outs.items()}, dtr.data_ptr(), dth.data_ptr()), reps=3)
ALL = DIAG_HEL | DIAG_HEMI | DIAG_POWER | DIAG_PERPPAR | DIAG_FLUX | DIAG_VISCBC
d1 = timed("diagnostics (hel, hemi, power, perpPar, fluxes, nlBLayers)", lambda: rl.diagnostics(ptr, ALL, device=True))
timed("diagnostics (hemi only)", lambda: rl.diagnostics(ptr, DIAG_HEMI, device=True))
timed("getTOnext", lambda: rl.to_next(ptr, device=True))
t1 = timed("getTO", lambda: rl.to(ptr, 1e-4, device=True))
timed("rms_keep", lambda: rl.rms_keep(ptr, device=True))
r1 = timed("get_nl_RMS batch (14 spectra to the host)", lambda: rl.rms(ptr, 1e-4, device=True))
b1 = timed("get_dtBLM batch (11 spectra to the host)", lambda: rl.dtb(ptr, device=True))
pinned = torch.empty((14, n_lev, s.lm_max), dtype=torch.complex128).pin_memory()
rp = timed("get_nl_RMS batch into page-locked memory", lambda: rl.rms(ptr, 1e-4, device=True, out=pinned.numpy()))
print("pinned result equals the pageable one:", bool(np.array_equal(rp, r1)))
d2 = rl.diagnostics(ptr, ALL, device=True)
print("diagnostics bitwise repeatable after the workspace changed hands:", bool(np.array_equal(d1, d2)),
      "| workspace + kept fields GB:", round((free0 - torch.cuda.mem_get_info()[0]) / 1e9, 1))
