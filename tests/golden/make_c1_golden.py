"""Generates tests/golden/c1_golden.npz: the oracle's result on BASELINE configs[0] (2-view synthetic pinhole pair 640x480,
PatchMatch 3 iterations, one neighbour) on the REFERENCE schedule (zig-zag order, mt19937), single-threaded and therefore
reproducible bit for bit, plus its agreement with an 8-thread run (the reference's own schedule-dependence, the yardstick the
GPU gate uses).  Run from the repository root:  python tests/golden/make_c1_golden.py
The oracle is a restatement (oracle/pm_oracle.cpp); the reference itself cannot be compiled in this image."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
from openmvs_b200 import synth
from conftest import agreement

sc = synth.make_scene(640, 480, 2, step_deg=5.0, cols=2)
views = [sc.views[0], sc.views[1]]
base = dict(nSubResolutionLevels=0, nEstimationGeometricIters=0)
zz1 = O.pm_estimate(views, O.default_params(schedule=0, nEstimationIters=3, threads=1, **base), sc.dmin, sc.dmax)
zz8 = O.pm_estimate(views, O.default_params(schedule=0, nEstimationIters=3, threads=8, **base), sc.dmin, sc.dmax)
iou, agree = agreement(zz1[0], zz8[0])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_golden.npz")
np.savez_compressed(out, zz1_depth=zz1[0], zz1_conf=zz1[2].astype(np.float16),
	iou_zz1_zz8=np.float32(iou), agree_zz1_zz8=np.float32(agree), image_sha=np.frombuffer(__import__("hashlib").sha256(views[0].image.tobytes()).digest(), np.uint8))
print("wrote %s: valid %.4f, ZZ1 vs ZZ8 iou %.4f agree %.4f, %d bytes" % (out, (zz1[0] > 0).mean(), iou, agree, os.path.getsize(out)))
