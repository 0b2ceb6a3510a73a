"""Builds tests/golden/condICrotIC_state_1000.npz: the state of the restated samples/dynamo_benchmark_condICrotIC run after
its first 1000 steps (what the reference's checkpoint_end.start holds at that point: fields, previous explicit terms,
inner-core rotation), so that the CPU suite can replay the RESTARTED stage (rows 1001..1101 of reference.out) without
repeating the first stage every time.

Unlike the other fixtures this one is NOT reference output: it is produced by oracle/lmloop.py + the CPU oracle.  It is
trustworthy only because (a) every one of the 1000 steps is checked here against reference.out / referenceMag.out at the
autotest tolerance while the state is being produced, and (b) the test that uses it compares the 101 following rows
against the reference again.  The GPU leg of tests/test_condICrotIC.py does not use it (it runs all 1100 steps).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import test_condICrotIC as t  # noqa: E402

d = np.load(os.path.join(HERE, "condICrotIC_reference.npz"))
golden = {k: d[k] for k in d.files}
h = t._oracle_loop(golden)
for row in range(1, t.N_FIRST + 1):
    h.step()
    t._check(golden, h, row)
np.savez_compressed(os.path.join(HERE, "condICrotIC_state_1000.npz"), **h.state_dict())
print("state after", h.n_steps, "steps, omega_ic", h.omega_ic, "time", h.time)
