"""Developer smoke: GPU kernels vs the RB oracle on a small scene + a 1080p timing."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200
from oracle import oracle as O

def stats(name, a, b, mask=None):
	d = np.abs(a-b)
	if mask is not None: d = d[mask]
	print(f"{name}: max {d.max():.3e} mean {d.mean():.3e} frac>1e-4 {(d>1e-4).mean():.4f} frac>1e-3 {(d>1e-3).mean():.4f}")

W, H, NV = 320, 240, 5
sc = synth.make_scene(W, H, NV, step_deg=5.0)
ref = 2; nb = sc.neighbors(ref, 4)
views = [sc.views[ref]]+[sc.views[i] for i in nb]
dev = torch.device("cuda:0")
imgs = [ViewData(torch.from_numpy(v.image).to(dev), Camera(v.K, v.R, v.C)) for v in views]
pm = PatchMatchB200(0)
OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 0; OPTDENSE.nEstimationIters = 3
nR = 3
prm = O.default_params(schedule=1, propagation=4, nRandomIters=nR, nSubResolutionLevels=0, nEstimationGeometricIters=0, threads=8)
# pass A
d0, n0, c0 = O.pm_score(views, prm, sc.dmin, sc.dmax)
plane = torch.zeros(H, W, 4, device=dev); cost = torch.zeros(H, W, device=dev)
pm.ScoreDepthMap(imgs, sc.dmin, sc.dmax, plane, cost)
torch.cuda.synchronize()
pg = plane.cpu().numpy(); cg = cost.cpu().numpy()
stats("passA depth", pg[..., 3], d0); stats("passA normal", pg[..., :3], n0); stats("passA cost", cg, c0)
# sweeps from the ORACLE state each time (isolates per-sweep parity)
d, n, c = d0, n0, c0
for sweep in range(4):
	plane = torch.from_numpy(np.concatenate([n, d[..., None]], -1)).to(dev).contiguous(); cost = torch.from_numpy(c).to(dev)
	pm.SweepDepthMap(imgs, sc.dmin, sc.dmax, plane, cost, sweep, nRandomIters=nR)
	torch.cuda.synchronize()
	d, n, c = O.pm_iterate(views, prm, sc.dmin, sc.dmax, d, n, c, sweep)
	pg = plane.cpu().numpy(); cg = cost.cpu().numpy()
	rel = np.abs(pg[..., 3]-d)/np.maximum(d, 1e-6)
	m = d > 0
	print(f"sweep {sweep}: depth rel>1e-5 {(rel[m]>1e-5).mean():.5f} rel>1e-3 {(rel[m]>1e-3).mean():.5f}", end=" | ")
	stats("cost", cg, c, m)
# full estimate, GPU chain vs oracle chain
dd = DepthData(imgs, sc.dmin, sc.dmax)
pm.EstimateDepthMap(dd)
gd = dd.depthMap.cpu().numpy(); gc = dd.confMap.cpu().numpy()
prm2 = O.default_params(schedule=1, propagation=4, nRandomIters=nR, nEstimationIters=6, nSubResolutionLevels=0, nEstimationGeometricIters=0, threads=8)
od, on, oc = O.pm_estimate(views, prm2, sc.dmin, sc.dmax)
both = (gd > 0) & (od > 0)
rel = np.abs(gd-od)[both]/od[both]
gt = sc.views[ref].depth_gt
print(f"full: IoU {both.sum()/((gd>0)|(od>0)).sum():.4f} agree<1e-3 {(rel<1e-3).mean():.4f} <1e-5 {(rel<1e-5).mean():.4f}; gpu-vs-gt<1e-3 {(np.abs(gd-gt)[gd>0]/gt[gd>0]<1e-3).mean():.4f} oracle-vs-gt {(np.abs(od-gt)[od>0]/gt[od>0]<1e-3).mean():.4f}")
print("stats ms_device", pm.stats.ms_device, "launches", pm.stats.kernel_launches)
# timing 1080p
W, H = 1920, 1080
t = time.time(); sc = synth.make_scene(W, H, 10, step_deg=4.0); print("scene1080 gen", time.time()-t)
views = [sc.views[4]]+[sc.views[i] for i in sc.neighbors(4, 9)]
imgs = [ViewData(torch.from_numpy(v.image).to(dev), Camera(v.K, v.R, v.C)) for v in views]
OPTDENSE.nEstimationIters = 6
for rep in range(3):
	dd = DepthData(imgs, sc.dmin, sc.dmax)
	pm.EstimateDepthMap(dd)
	print(f"1080p N=9 I=6: device {pm.stats.ms_device:.1f} ms -> {W*H/pm.stats.ms_device/1e3:.1f} Mpix/s, launches {pm.stats.kernel_launches}")
gd = dd.depthMap.cpu().numpy(); gt = sc.views[4].depth_gt
print(f"1080p valid {(gd>0).mean():.3f} gt<1e-3 {(np.abs(gd-gt)[gd>0]/gt[gd>0]<1e-3).mean():.4f}")
