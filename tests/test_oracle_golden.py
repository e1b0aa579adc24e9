"""The oracle against its committed golden fixture (tests/golden/c1_golden.npz, BASELINE configs[0] on the reference's zig-zag
schedule, single thread, mt19937): pins the restatement against accidental change — every edit of oracle/pm_oracle.cpp that
alters the reference-schedule result fails here.  (Parity of the oracle with the reference itself stays unpinned)"""
import hashlib
import os

import numpy as np

from conftest import agreement


def test_oracle_reproduces_c1_golden_fixture():
	from oracle import oracle as O
	from openmvs_b200 import synth
	g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_golden.npz"))
	sc = synth.make_scene(640, 480, 2, step_deg=5.0, cols=2)
	views = [sc.views[0], sc.views[1]]
	# the synthetic scene generator is part of the fixture's identity
	assert np.array_equal(np.frombuffer(hashlib.sha256(views[0].image.tobytes()).digest(), np.uint8), g["image_sha"])
	d, n, c = O.pm_estimate(views, O.default_params(schedule=0, nEstimationIters=3, threads=1, nSubResolutionLevels=0, nEstimationGeometricIters=0), sc.dmin, sc.dmax)
	assert np.array_equal(d, g["zz1_depth"])
	assert np.array_equal(c.astype(np.float16), g["zz1_conf"])
	# and the fixture's own numbers are the ones quoted in DESIGN.md
	assert 0.95 < float(g["agree_zz1_zz8"]) < 1.0 and float(g["iou_zz1_zz8"]) > 0.998


def test_engine_schedule_rule():
	"""b200mvs_get_schedule: one rule for every iteration count (no per-configuration knob)"""
	from openmvs_b200.depth_estimator import OPTDENSE
	saved = (OPTDENSE.nEstimationIters, OPTDENSE.nSweepsPerIter, OPTDENSE.nRandomIters)
	try:
		OPTDENSE.nRandomIters = 6; OPTDENSE.nSweepsPerIter = 0
		got = {}
		for it in (1, 2, 3, 4, 6, 8):
			OPTDENSE.nEstimationIters = it
			got[it] = OPTDENSE.schedule()
			n, r = got[it]
			assert n >= 8 and n*r >= 6*it   # at least the reference's number of refinement tries
		assert got[3] == (8, 3) and got[6] == (9, 4) and got[8] == (12, 4)
		assert OPTDENSE.schedule(True) == (2, 3)  # a geometric pass: two sweeps
		OPTDENSE.nSweepsPerIter = 2; OPTDENSE.nEstimationIters = 6
		assert OPTDENSE.schedule() == (12, 3) and OPTDENSE.schedule(True) == (2, 3)
	finally:
		OPTDENSE.nEstimationIters, OPTDENSE.nSweepsPerIter, OPTDENSE.nRandomIters = saved
