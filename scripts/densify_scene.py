"""Dense depth maps of a synthetic scene over the GPUs of one box: pass 1 for every reference view,
then geometric-consistency passes with a depth all-gather in between (compute_depth_maps).
  python -m torch.distributed.run --nproc-per-node N scripts/densify_scene.py [n_views] [width] [height] [geo_iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from openmvs_b200 import synth, multi_gpu
from openmvs_b200.depth_estimator import OPTDENSE

n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 12
w = int(sys.argv[2]) if len(sys.argv) > 2 else 960
h = int(sys.argv[3]) if len(sys.argv) > 3 else 540
geo = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
	dist.init_process_group("nccl", device_id=dev)
sc = synth.make_scene(w, h, n_views, step_deg=4.0, device=dev)
nbrs = [sc.neighbors(v, min(9, n_views-1)) for v in range(n_views)]
OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationIters = 6; OPTDENSE.nEstimationGeometricIters = geo
est = multi_gpu.SceneEstimator(sc.views, nbrs, sc.dmin, sc.dmax, device=dev)
torch.cuda.synchronize(); t = time.perf_counter()
res = multi_gpu.compute_depth_maps(n_views, est, n_geometric_iters=geo, dst=0)
torch.cuda.synchronize(); dt = time.perf_counter()-t
if rank == 0:
	acc, val = [], []
	for v in range(n_views):
		d = res[v][..., 0].cpu().numpy(); gt = sc.views[v].depth_gt; m = d > 0
		val.append(m.mean()); acc.append((np.abs(d-gt)[m]/gt[m] < 1e-3).mean())
	print("densify: %d views %dx%d on %d GPU(s), pass 1 + %d geometric: %.2f s, %.1f Mpix/s of final maps; valid %.3f, within 1e-3 of ground truth %.4f" % (
		n_views, w, h, world, geo, dt, n_views*w*h/1e6/dt, np.mean(val), np.mean(acc)))
if world > 1:
	dist.destroy_process_group()
