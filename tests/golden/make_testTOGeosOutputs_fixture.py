"""Extracts the golden vectors of samples/testTOGeosOutputs from the reference tree (run in the build container).

samples/testTOGeosOutputs/unitTest.py compares `cat geos.start Tay.start tmp` with reference.out, tmp being points of the TO movie,
of TOnhs / TOshs and of the geos movies printed with four decimals.  The run restarts the saturated benchmark dynamo of
samples/boussBenchSat (its checkpoint, same physics: tests/golden/boussBenchSat_ckpt.npz) with l_TO = .true., n_TO_step = 5 and
advances it by 25 BPR353 steps.  Rows 7-11 of reference.out are Tay.start (out_TO.f90:552-553):
  time, VpRMS^2, VgRMS^2, TayRMS, TayRRMS, TayVRMS, eKin   (ES16.8)
-- the energy fractions of the axisymmetric / geostrophic azimuthal flow and the Taylorisation measures of the Lorentz, Reynolds and
viscous stresses, z-averaged on a cylindrical grid from the (r, theta) arrays VAS, dzLFAS, dzRstrAS that getTO fills inside the
radial loop (TO.f90:141-307) and dzStrAS of getTOfinish.
"""
import os

import numpy as np

REF = "/root/reference/samples/testTOGeosOutputs"
HERE = os.path.dirname(os.path.abspath(__file__))

rows = [np.array(l.split(), dtype=float) for l in open(os.path.join(REF, "reference.out")) if l.strip()]
tay = np.array([r for r in rows if len(r) == 7 and r[0] > 100.0])
assert tay.shape == (5, 7)
# the line after Tay.start: seven points of the TO movie (unitTest.py:50-53; frames are written on the TO steps, single precision,
# printed with four decimals): to.asVphi[0, 13, 3], to.rey[1, 21, 22], to.adv[1, 52, 11], to.visc[0, 12, 25], to.lorentz[0, 73, 30],
# to.coriolis[1, 33, 3], to.dtVp[1, 88, 7] -- [frame, theta (ordered), r] of VAS, dzRstrAS, dzAstrAS, dzStrAS, LFfac dzLFAS, dzCorAS,
# dzdVpAS (out_TO.f90:564-575, python/magic/TOreaders.py:122-135)
mov = [r for r in rows if len(r) == 7 and r[0] < 100.0]
assert len(mov) == 1
points = np.array([[0, 13, 3], [1, 21, 22], [1, 52, 11], [0, 12, 25], [0, 73, 30], [1, 33, 3], [1, 88, 7]])
# the two lines after that: points of TOnhs.TAG and TOshs.TAG (unitTest.py:56-63; z-averages on the cylindrical grid, [TO step, n_s]):
# north  to.vp[2, 18], to.rstr[3, 11], to.astr[0, 30], to.LF[2, 21];  south  to.dvp[3, 12], to.viscstr[1, 9], to.tay[4, 27], to.vpr[2, 21]
hemi = [r for r in rows if len(r) == 4]
assert len(hemi) == 2
np.savez_compressed(os.path.join(HERE, "testTOGeosOutputs_reference.npz"), Tay=tay, n_TO_step=5, n_time_steps=25, movie_values=mov[0],
                    movie_points=points, nhs_values=hemi[0], shs_values=hemi[1])
print(tay, mov[0], hemi)
