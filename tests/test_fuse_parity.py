"""DepthMapsData::FuseDepthMaps (libs/MVS/SceneDensify.cpp:1372-1646): b200mvs_fuse_depth_maps (host code behind the C-ABI) against
the Python restatement oracle/fuse_oracle.py on small synthetic scenes — same points, views, weights, normals, same zeroed
depths — and properties on the analytic scene: fused points lie on the ground-truth surface."""
import copy

import numpy as np
import pytest

from openmvs_b200 import synth
from openmvs_b200.depth_estimator import FuseDepthMaps


def _scene_views(w, h, n, noise, seed, drop=0.1, with_color=True):
	"""depth-maps of a synthetic scene as an estimator would deliver them: ground truth + relative noise, some pixels invalid,
	a band of gross outliers (they must be rejected or invalidated), confidences, camera-space normals"""
	sc = synth.make_scene(w, h, n, step_deg=5.0)
	rng = np.random.RandomState(seed)
	views = []
	for i, v in enumerate(sc.views):
		d = v.depth_gt.astype(np.float32).copy()
		d *= (1+noise*rng.randn(h, w)).astype(np.float32)
		d[rng.rand(h, w) < drop] = 0
		d[h//3:h//3+2, :] *= np.float32(0.8)        # outliers in front of the surface
		nb = [j for j in sc.neighbors(i, min(n-1, 4))]
		color = (np.clip(v.image, 0, 1)[..., None]*255).astype(np.uint8).repeat(3, -1) if with_color else None
		views.append(dict(depth=d, normal=v.normal_gt.astype(np.float32), conf=rng.rand(h, w).astype(np.float32), color=color,
			K=v.K, R=v.R, C=v.C, neighbors=nb, n_scene_neighbors=len(nb)+(i % 2)))
	return sc, views


@pytest.mark.parametrize("nmin,with_color,seed", [(2, True, 0), (3, True, 1), (1, False, 2)])
def test_fusion_equals_oracle(nmin, with_color, seed):
	from oracle import fuse_oracle as FO
	sc, views = _scene_views(56, 42, 5, 2e-3, seed, with_color=with_color)
	if seed == 2:
		views[3]["depth"] = None            # an image without a depth-map
	va, vb = copy.deepcopy(views), copy.deepcopy(views)
	pc = FuseDepthMaps(va, nMinViewsFuse=nmin)
	ref = FO.fuse_depth_maps(vb, nMinViewsFuse=nmin)
	assert len(pc.points) == len(ref["points"]) and len(pc.points) > 500
	assert [list(v) for v in pc.pointViews] == ref["views"]
	assert all(np.array_equal(a, np.array(b, np.float32)) for a, b in zip(pc.pointWeights, ref["weights"]))
	assert all(np.array_equal(a, np.array(b, np.uint16).reshape(-1, 2)) for a, b in zip(pc.projs, ref["projs"]))
	assert np.array_equal(pc.points, ref["points"])
	assert np.allclose(pc.normals, ref["normals"], atol=2e-7)
	if with_color:
		assert np.abs(pc.colors.astype(int)-ref["colors"].astype(int)).max() <= 1
	else:
		assert pc.colors is None
	# the depth-maps were modified identically (blocking depths zeroed)
	for a, b, o in zip(va, vb, views):
		if a["depth"] is not None:
			assert np.array_equal(a["depth"], b["depth"])
	assert sum(int((o["depth"] != a["depth"]).sum()) for a, o in zip(va, views) if o["depth"] is not None) > 0


def test_fused_points_lie_on_the_surface_and_errors():
	sc, views = _scene_views(160, 120, 6, 1e-3, 5, drop=0.05)
	pc = FuseDepthMaps(views, nMinViewsFuse=3)
	assert len(pc.points) > 5000 and pc.nDepths > len(pc.points)
	assert min(len(v) for v in pc.pointViews) >= 3 and all(np.all(np.diff(v.astype(int)) > 0) for v in pc.pointViews[:2000])
	# distance to the analytic surface through the first view's depth: project every point into view 0 and compare depths
	v0 = sc.views[0]
	Xc = (pc.points.astype(np.float64)-v0.C) @ v0.R.T
	uv = (Xc @ v0.K.T); x = np.rint(uv[:, 0]/uv[:, 2]).astype(int); y = np.rint(uv[:, 1]/uv[:, 2]).astype(int)
	ok = (x >= 0) & (y >= 0) & (x < 160) & (y < 120)
	rel = np.abs(Xc[ok, 2]-v0.depth_gt[y[ok], x[ok]])/Xc[ok, 2]
	assert np.median(rel) < 2e-3 and (rel < 1e-2).mean() > 0.97     # the 20 % outlier rows did not survive a 3-view agreement
	assert np.allclose(np.linalg.norm(pc.normals, axis=1), 1, atol=1e-5)
	from openmvs_b200 import lib
	bad = [dict(depth=np.zeros((4, 4), np.float32), K=np.eye(3), R=np.eye(3), C=np.zeros(3), neighbors=[7])]
	with pytest.raises(lib.B200MVSError):
		FuseDepthMaps(bad)
