"""Small PatchMatch (photometric, multi-scale, geometric) + SGM runs for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200, SemiGlobalMatcher
sc = synth.make_scene(203, 151, 3, step_deg=5.0, cols=3)
views = [sc.views[1], sc.views[0], sc.views[2]]
OPTDENSE.nEstimationIters = 1; OPTDENSE.nEstimationGeometricIters = 1
pm = PatchMatchB200(0)
for levels in (0, 1):
	OPTDENSE.nSubResolutionLevels = levels
	dd = DepthData([ViewData(np.ascontiguousarray(v.image), Camera(v.K, v.R, v.C)) for v in views], sc.dmin, sc.dmax)
	pm.EstimateDepthMap(dd)
	print("pm levels", levels, "valid", (dd.depthMap > 0).mean(), "tma", pm.stats.tma_active)
OPTDENSE.nSubResolutionLevels = 0
g = DepthData([ViewData(np.ascontiguousarray(v.image), Camera(v.K, v.R, v.C), depthMap=(v.depth_gt if i else None), cameraDepthMap=Camera(v.K, v.R, v.C)) for i, v in enumerate(views)],
	sc.dmin, sc.dmax, depthMap=dd.depthMap.copy(), normalMap=dd.normalMap.copy())
pm.Init(True); pm.EstimateDepthMap(g, 0); pm.Release()
print("geo valid", (g.depthMap > 0).mean())
lg, lc, rg, d = synth.make_stereo_pair(131, 77)
rng = np.random.RandomState(0)
lo = rng.randint(-3, 3, (71, 125)); hi = lo+rng.randint(1, 40, (71, 125))
px, n = synth.sgm_pixel_map(131, 77, lo, hi, rng.rand(71, 125) < 0.1)
m = SemiGlobalMatcher(); disp, cost = m.Match(lg, lc, rg, px, n); m.Release()
print("sgm ok", disp.shape, int((disp != 32767).sum()))
