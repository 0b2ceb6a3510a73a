"""Extracts the golden vectors of samples/testRMSOutputs from the reference tree (run in the build container).

samples/testRMSOutputs/unitTest.py compares `cat dtVrms.start dtBrms.start dtVrms.continue dtBrms.continue dtVrms.FD dtBrms.FD` with
reference.out.  The first run (input.nml, tag "start") restarts the saturated benchmark dynamo of samples/boussBenchSat (its
checkpoint, same physics: tests/golden/boussBenchSat_ckpt.npz) with l_RMS = .true., rCut = 1e-2 and advances it by 50 BPR353 steps,
logging every 10: the first five rows of reference.out are dtVrms.start (RMS.f90:1178-1186):
  time, InerRms, CorRms, LFRms, AdvRms, DifRms, Buo_tempRms, Buo_xiRms, PreRms (ES16.8), then GeoRms/(Cor+Pre), MagRms/(Cor+Pre+LF),
  ArcRms/(Cor+Pre+Buo), ArcMagRms/(Cor+Pre+LF+Buo), CLFRms/(Cor+LF), PLFRms/(Pre+LF), CIARms/(Cor+Pre+Buo+Iner+LF) (ES14.6).
All but DifRms are built on the fourteen spectra that transform_to_lm_RMS returns from the radial loop (RMS.f90:576-610).
The next five rows are dtBrms.start: the dynamo terms (Pdyn, Tdyn, the omega effect, the dipole part) are built on the eleven
spectra of get_dtBLM (dtB.f90:144-223), which l_RMS switches on for the logged steps (step_time.f90:386).
"""
import os

import numpy as np

REF = "/root/reference/samples/testRMSOutputs"
HERE = os.path.dirname(os.path.abspath(__file__))

rows = [np.array(l.split(), dtype=float) for l in open(os.path.join(REF, "reference.out")) if l.strip()]
dtVrms = np.array(rows[:5])
dtBrms = np.array(rows[5:10])      # dtBrms.start (RMS.f90:1407-1411): time, dtBPolRms, dtBTorRms, PdynRms, TdynRms, PdifRms, TdifRms,
assert dtVrms.shape == (5, 16) and dtBrms.shape == (5, 11)   # TomeRms/TdynRms, TomeAsRms/TdynRms, DdynRms, DdynAsRms (ES16.8)
assert np.array_equal(dtVrms[:, 0], dtBrms[:, 0])
np.savez_compressed(os.path.join(HERE, "testRMSOutputs_reference.npz"), dtVrms=dtVrms, dtBrms=dtBrms, n_log_step=10, n_time_steps=50,
                    rCut=1e-2, rDea=0.0)
print(dtVrms[:, :5])
