"""CPU restatement of DepthMapsData::FuseDepthMaps (libs/MVS/SceneDensify.cpp:1372-1646) — TEST INFRASTRUCTURE ONLY.

Plain Python loops over numpy arrays (small cases only): the sequential greedy fusion of all depth-maps into one point
cloud, processing the best connected images first, every pixel in raster order.  Each function cites the reference lines it
follows.  Only tests/ may import this module; the product path is b200mvs_fuse_depth_maps (openmvs_b200/csrc/fuse_host.cu).
parity unpinned: the reference has no golden vectors for this function and cannot be built here.
"""
from __future__ import annotations

import math

import numpy as np

NO_ID = 0xFFFFFFFF


def conf2weight(conf, depth):
	"""Conf2Weight (SceneDensify.cpp:120-122), float arithmetic"""
	conf = np.float32(conf); depth = np.float32(depth)
	return np.float32(1)/(np.maximum(np.float32(1)-conf, np.float32(0.03))*depth*depth)


def is_depth_similar(d0, d1, th):
	"""IsDepthSimilar (libs/Common/Util.inl:797-809): |d0-d1|/d0 < th, float, not symmetric"""
	d0 = np.float32(d0); d1 = np.float32(d1)
	return bool(np.abs(d0-d1)/d0 < np.float32(th))


def projection_matrix(K, R, C):
	"""AssembleProjectionMatrix (libs/MVS/Camera.cpp:173-180): P = [K R | -K R C], double"""
	M = np.asarray(K, np.float64) @ np.asarray(R, np.float64)
	return np.concatenate([M, (M @ (-np.asarray(C, np.float64))).reshape(3, 1)], 1)


def image_to_world(K, R, C, x, y, depth):
	"""Camera::TransformPointI2W(Point3) (libs/MVS/Camera.h:339-356), double; K's skew is not used, like the reference"""
	z = float(depth)
	Xc = np.array([(float(x)-K[0, 2])*z/K[0, 0], (float(y)-K[1, 2])*z/K[1, 1], z], np.float64)
	return R.T @ Xc + C


def fuse_depth_maps(views, nMinViewsFuse=2, fDepthDiffThreshold=0.01, fNormalDiffThreshold=25.0, estimate_color=True, estimate_normal=True):
	"""views: list of dicts with depth (h, w) float32 [modified: depths that disagree with a fused point are zeroed], normal
	(h, w, 3) camera space or None, conf (h, w) or None, color (h, w, 3) uint8 or None, K R C (float64), neighbors (IDs, best
	first = depthData.neighbors), n_scene_neighbors (scene.images[i].neighbors.size(), the connection score).
	Returns dict(points (n,3) f32, views [lists], weights [lists of f32], colors (n,3) u8 or None, normals (n,3) f32 or None)."""
	n = len(views)
	for v in views:
		v["K"] = np.asarray(v["K"], np.float64); v["R"] = np.asarray(v["R"], np.float64); v["C"] = np.asarray(v["C"], np.float64)
		v["P"] = projection_matrix(v["K"], v["R"], v["C"])
	# best connected images first (SceneDensify.cpp:1392-1449); std::sort is not stable: ties are broken by index here and in the
	# engine (the reference's order among equal scores is unspecified)
	conn = [(i, float(views[i]["n_scene_neighbors"])) for i in range(n) if views[i]["depth"] is not None and (views[i]["depth"] > 0).any()]
	conn = [c for c in conn if c[1] > 0]
	conn.sort(key=lambda c: (-c[1], c[0]))
	bNormalMap = all(views[i]["normal"] is not None for i, _ in conn)
	if estimate_normal and not bNormalMap:
		estimate_normal = False
	nMin = min(int(nMinViewsFuse), n)
	normalError = np.float32(math.cos(math.radians(fNormalDiffThreshold)))
	idx_maps = [None]*n
	points, pviews, pweights, pprojs, colors, normals = [], [], [], [], [], []
	for idxImage, _ in conn:
		A = views[idxImage]
		for nb in A["neighbors"]:
			if idx_maps[nb] is None and views[nb]["depth"] is not None:
				idx_maps[nb] = np.full(views[nb]["depth"].shape, NO_ID, np.uint32)
		if idx_maps[idxImage] is None:
			idx_maps[idxImage] = np.full(A["depth"].shape, NO_ID, np.uint32)
		dA, idxA = A["depth"], idx_maps[idxImage]
		h, w = dA.shape
		for i in range(h):
			for j in range(w):
				depth = dA[i, j]
				if depth == 0 or idxA[i, j] != NO_ID:
					continue
				idxPoint = len(points)
				idxA[i, j] = idxPoint
				point = image_to_world(A["K"], A["R"], A["C"], j, i, depth).astype(np.float32)     # PointCloud::Point is float
				vlist, wlist, plist = [idxImage], [conf2weight(1.0 if A["conf"] is None else A["conf"][i, j], depth)], [(j, i)]
				confidence = float(wlist[0])                                                # REAL
				normal = (A["R"].T @ A["normal"][i, j].astype(np.float64)).astype(np.float32) if bNormalMap else np.array([0, 0, -1], np.float32)
				X = point.astype(np.float64)*confidence
				Cc = (A["color"][i, j].astype(np.float32)*np.float32(confidence)) if (estimate_color and A["color"] is not None) else None
				N = normal*np.float32(confidence)
				invalid = []
				for nb in A["neighbors"]:
					B = views[nb]
					if B["depth"] is None:
						continue
					pt = (B["P"] @ np.append(point.astype(np.float64), 1.0)).astype(np.float32)    # ProjectPointP3<float>: double sums, float result
					if pt[2] <= 0:
						continue
					xb = int(math.floor(np.float32(pt[0]/pt[2])+np.float32(0.5))); yb = int(math.floor(np.float32(pt[1]/pt[2])+np.float32(0.5)))
					hb, wb = B["depth"].shape
					if not (0 <= xb < wb and 0 <= yb < hb):
						continue
					depthB = B["depth"][yb, xb]
					if depthB == 0 or idx_maps[nb][yb, xb] != NO_ID:
						continue
					if is_depth_similar(pt[2], depthB, fDepthDiffThreshold):
						normalB = (B["R"].T @ B["normal"][yb, xb].astype(np.float64)).astype(np.float32) if bNormalMap else np.array([0, 0, -1], np.float32)
						if np.float32(np.dot(normal, normalB)) > normalError:
							confB = conf2weight(1.0 if B["conf"] is None else B["conf"][yb, xb], depthB)
							k = 0                                                               # InsertSort: ascending view id
							while k < len(vlist) and vlist[k] < nb:
								k += 1
							vlist.insert(k, nb); wlist.insert(k, confB); plist.insert(k, (xb, yb))
							idx_maps[nb][yb, xb] = idxPoint
							X = X + image_to_world(B["K"], B["R"], B["C"], xb, yb, depthB)*float(confB)
							if Cc is not None:
								Cc = Cc + B["color"][yb, xb].astype(np.float32)*confB
							if estimate_normal:
								N = N + normalB*confB
							confidence += float(confB)
							continue
					if pt[2] < depthB:
						invalid.append((nb, yb, xb))
				if len(vlist) < nMin:
					for vv, (px, py) in zip(vlist, plist):
						idx_maps[vv][py, px] = NO_ID
				else:
					nrm = 1.0/confidence
					points.append((X*nrm).astype(np.float32)); pviews.append(vlist); pweights.append(wlist); pprojs.append(plist)
					if Cc is not None:
						colors.append((Cc*np.float32(nrm)).astype(np.uint8))
					if estimate_normal:
						Nn = N*np.float32(nrm)
						normals.append(Nn/np.float32(np.sqrt(np.float32(np.dot(Nn, Nn)))))
					for nb, yb, xb in invalid:
						views[nb]["depth"][yb, xb] = 0
	return dict(points=np.array(points, np.float32).reshape(-1, 3), views=pviews, weights=pweights, projs=pprojs,
		colors=np.array(colors, np.uint8).reshape(-1, 3) if colors else None,
		normals=np.array(normals, np.float32).reshape(-1, 3) if normals else None)
