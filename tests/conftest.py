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
