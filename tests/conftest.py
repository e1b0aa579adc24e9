import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_scene():
	"""5 views 320x240, reference = view 2 with its 4 neighbours (seeded, numpy-generated)."""
	from openmvs_b200 import synth
	sc = synth.make_scene(320, 240, 5, step_deg=5.0)
	ref = 2
	nb = sc.neighbors(ref, 4)
	views = [sc.views[ref]]+[sc.views[i] for i in nb]
	return sc, ref, views


@pytest.fixture(scope="session")
def tiny_scene():
	"""3 views 160x120 for the slow sequential oracle runs."""
	from openmvs_b200 import synth
	sc = synth.make_scene(160, 120, 3, step_deg=5.0, cols=3)
	ref = 1
	views = [sc.views[ref], sc.views[0], sc.views[2]]
	return sc, ref, views


def agreement(da, db, rel_tol=1e-3):
	"""(mask IoU, fraction of commonly-confident pixels with |da-db|/da < rel_tol)"""
	ma, mb = da > 0, db > 0
	both = ma & mb
	union = (ma | mb).sum()
	iou = both.sum()/max(1, union)
	rel = np.abs(da-db)[both]/da[both]
	return float(iou), float((rel < rel_tol).mean()) if both.any() else 0.0


def rb_oracle(views, dmin, dmax, geometric_iter=-1, threads=8, mask=None, depth=None, normal=None, depths=None):
	"""The CPU oracle on the engine's red-black schedule for the current OPTDENSE options: the same sweeps, refinement tries,
	propagation candidates and changed-flag rule as b200mvs_estimate* runs (b200mvs_get_schedule), same Philox stream."""
	from oracle import oracle as O
	from openmvs_b200.depth_estimator import OPTDENSE as OPT
	T, nR = OPT.schedule(False)
	S, nRg = OPT.schedule(True)
	geo = geometric_iter >= 0
	prm = O.default_params(schedule=1, propagation=O.rb_propagation(OPT.nPropagation, OPT.nPropagationFar, OPT.bSkipUnchanged, OPT.nEvalCap),
		nRandomIters=nRg if geo else nR, nSubResolutionLevels=OPT.nSubResolutionLevels, nEstimationGeometricIters=OPT.nEstimationGeometricIters,
		fEstimationGeometricWeight=OPT.fEstimationGeometricWeight, fNCCThresholdKeep=OPT.fNCCThresholdKeep, seed=OPT.nSeed, threads=threads)
	b, e = (T+geometric_iter*S, T+(geometric_iter+1)*S) if geo else (0, T)
	return O.pm_estimate_range(views, prm, dmin, dmax, b, e, geometric=geo, mask=mask, depth=depth, normal=normal, depths=depths)
