"""Short single-view 1080p run for ncu captures (one reference view, 9 neighbours)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
sc = synth.make_scene(1920, 1080, 12, step_deg=4.0, device=dev)
r = 5
views = [sc.views[r]]+[sc.views[i] for i in sc.neighbors(r, 9)]
imgs = [ViewData(torch.from_numpy(v.image).to(dev), Camera(v.K, v.R, v.C)) for v in views]
OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 0; OPTDENSE.nEstimationIters = iters
pm = PatchMatchB200(0)
dd = DepthData(imgs, sc.dmin, sc.dmax)
pm.EstimateDepthMap(dd)
pm.EstimateDepthMap(dd)
import numpy as np
gd = dd.depthMap.cpu().numpy(); gt = sc.views[r].depth_gt
print("layout", os.environ.get("B200MVS_LAYOUT", "default"), "device ms %.2f" % pm.stats.ms_device, "launches", pm.stats.kernel_launches,
	"sweep launch avg ms %.3f" % (pm.stats.ms_sweep_kernels/max(1, pm.stats.sweep_launches)), "n", pm.stats.sweep_launches,
	"valid %.3f gt<1e-3 %.4f" % ((gd > 0).mean(), (np.abs(gd-gt)[gd > 0]/gt[gd > 0] < 1e-3).mean()))
