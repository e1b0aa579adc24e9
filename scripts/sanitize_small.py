"""Small PatchMatch (photometric, multi-scale, geometric) + SGM + post-processing runs for compute-sanitizer.
SANITIZE_ONLY=post runs the post-processing part alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200, SemiGlobalMatcher
sc = synth.make_scene(203, 151, 3, step_deg=5.0, cols=3)
views = [sc.views[1], sc.views[0], sc.views[2]]
OPTDENSE.nEstimationIters = 1; OPTDENSE.nEstimationGeometricIters = 1
ONLY = os.environ.get("SANITIZE_ONLY", "")
pm = PatchMatchB200(0)
if ONLY:
	pm.Release()
if ONLY == "":
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
if ONLY in ("", "sgm"):
	# uniform 16-byte aligned ranges: the bulk-copy ring aggregation kernel (mbarrier + cp.async.bulk)
	lg, lc, rg, d = synth.make_stereo_pair(150, 90)
	for num in (32, 144, 64, 128):   # ring kernel; wave-front aggregation + tensor-core cost kernel
		px, n = synth.sgm_pixel_map(150, 90, -8, -8+num)
		m = SemiGlobalMatcher(); disp, cost = m.Match(lg, lc, rg, px, n); m.Release()
		print("sgm ring num", num, "ok", int((disp != 32767).sum()))
if ONLY == "sgm":
	sys.exit(0)
# depth-map post-processing: filter (both branches), speckles, gaps
from openmvs_b200.depth_estimator import DepthMapsData
maps = synth.make_noisy_dmaps(sc)
def _dd(i, conf=True):
	v = sc.views[i]
	return DepthData([ViewData(None, Camera(v.K, v.R, v.C))], sc.dmin, sc.dmax, depthMap=maps[i][0].copy(), confMap=maps[i][1].copy() if conf else None)
dm = DepthMapsData([], 0, nCalibratedImages=3)
for adjust in (True, False):
	fd, fc = dm.FilterDepthMap(_dd(1), [_dd(0, adjust), _dd(2, adjust)], adjust)
	print("filter adjust", adjust, "kept", (fd > 0).mean())
p = DepthData([], 0, 0, fd.copy(), np.ascontiguousarray(sc.views[1].normal_gt*(fd > 0)[..., None]), fc.copy())
dm.RemoveSmallSegments(p); print("segments kept", (p.depthMap > 0).mean())
dm.GapInterpolation(p); print("gaps filled to", (p.depthMap > 0).mean())
dm.pmCUDA.Release()
