"""One short GPU call: (1) the double-curl branch of the CUDA radial loop against the CPU oracle, (2) the CUDA radial loop inside
the samples/full_sphere time loop against reference.out.  Writes progressively to gpurun_out/full_sphere_gpu.log so that a
call cut off by its time limit still leaves what it had.  Usage: python tools/gpu_check_full_sphere.py [n_rows]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "full_sphere_gpu.log"), "w")
T0 = time.time()


def say(*a):
    msg = "%7.2fs " % (time.time() - T0) + " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


from magic_b200 import RadialLoop, Sht  # noqa: E402
from tests.test_full_sphere import _oracle, _oracle_params, _setup, _sizes  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "full_sphere_reference.npz"))
golden = {k: d[k] for k in d.files}
gs = _sizes(golden)
s = Sht(gs["l_max"], m_max=gs["m_max"], minc=3, n_theta_max=gs["n_theta_max"], n_phi_max=gs["n_phi_max"])
say("Sht created")
h, p, rad = _setup(golden, s.lm2l, s.lm2m)
rl = RadialLoop(s, p, rad)
say("RadialLoop created")
# (1) one radial loop on the checkpoint state, GPU vs oracle
o = _oracle(gs)
f = {k: np.ascontiguousarray(v) for k, v in h.fields_Rloc().items()}
got = rl.radialLoop(f)
ref = o.radial_loop(_oracle_params(p), rad, f)
for nm in ("dwdt", "dzdt", "dsdt", "dVSrLM", "dVxVhLM", "dpdt"):
    sel = slice(None) if nm in ("dVSrLM", "dVxVhLM") else slice(1, -1)
    den = np.linalg.norm(ref[nm][sel])
    say(nm, "rel L2 GPU vs oracle", np.linalg.norm(got[nm][sel] - ref[nm][sel]) / den if den else np.linalg.norm(got[nm][sel]), "norm", den)
say("dtrkc", np.abs(got["dtrkc"] / ref["dtrkc"] - 1).max(), "dthkc", np.abs(got["dthkc"] / ref["dthkc"] - 1).max())
# (2) the time loop
h.radial_loop = lambda fl: rl.radialLoop(fl)
n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for row in range(1, n_rows + 1):
    for _ in range(10):
        h.step()
    got = np.concatenate([[h.time], h.e_kin()])
    say("row", row, "max rel dev from reference.out", np.abs(got / golden["e_kin"][row] - 1).max())
say("launches", rl.launch_count())
