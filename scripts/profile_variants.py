"""Times the experimental tap-loop variants of pm_sweep_kernel against the default on one 1080p reference
view with 9 neighbours, and checks that each variant's result is bit-identical to the default's
(B200MVS_LAYOUT / B200MVS_PACK are read when a context is created)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
sc = synth.make_scene(1920, 1080, 12, step_deg=4.0, device=dev)
r = 5
views = [sc.views[r]]+[sc.views[i] for i in sc.neighbors(r, 9)]
imgs = [ViewData(torch.from_numpy(v.image).to(dev), Camera(v.K, v.R, v.C)) for v in views]
OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 0; OPTDENSE.nEstimationIters = iters
ref = None
for layout, pack in (("1", "0"), ("1", "1"), ("3", "0"), ("3", "1"), ("1", "0")):
	os.environ["B200MVS_LAYOUT"] = layout; os.environ["B200MVS_PACK"] = pack
	pm = PatchMatchB200(0)
	dd = DepthData(imgs, sc.dmin, sc.dmax)
	pm.EstimateDepthMap(dd)                       # from random initialisation
	first = (dd.depthMap.clone(), dd.normalMap.clone(), dd.confMap.clone())
	ms_random = pm.stats.ms_sweep_kernels/max(1, pm.stats.sweep_launches)
	pm.EstimateDepthMap(dd)                       # continues from the converged state
	ms_conv = pm.stats.ms_sweep_kernels/max(1, pm.stats.sweep_launches)
	same = ""
	if ref is None:
		ref = first
	else:
		same = "identical to default: %s" % all(torch.equal(a, b) for a, b in zip(ref, first))
	print("layout %s pack %s | sweep launch ms: random init %.3f, converged %.3f | valid %.4f | %s" % (layout, pack, ms_random, ms_conv,
		float((dd.depthMap > 0).float().mean()), same), flush=True)
	pm.Release()
