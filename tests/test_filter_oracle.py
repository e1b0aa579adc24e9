"""CPU checks of the post-processing oracle (oracle/filter_oracle.cpp) against independent
transcriptions of DepthMapsData::FilterDepthMap / RemoveSmallSegments / GapInterpolation
(libs/MVS/SceneDensify.cpp:810-1299) and against the analytic scene."""
import numpy as np
import pytest

from openmvs_b200 import synth
from oracle import oracle as O

f32 = np.float32


@pytest.fixture(scope="module")
def scene():
	sc = synth.make_scene(96, 72, 4, step_deg=5.0)
	return sc, synth.make_noisy_dmaps(sc)


def _dm(sc, maps, i, conf=True):
	v = sc.views[i]
	return (maps[i][0], maps[i][1] if conf else None, v.K, v.R, v.C)


# ---- independent numpy transcription of the forward projection ----
def _project_np(ref, nbr):
	dr, _, Kr, Rr, Cr = ref
	dn, cn, Kn, Rn, Cn = nbr
	H, W = dr.shape
	h, w = dn.shape
	ys, xs = np.mgrid[0:h, 0:w]
	idx = (ys*w+xs).ravel()
	z = dn.ravel().astype(np.float64)
	x = xs.ravel().astype(np.float64); y = ys.ravel().astype(np.float64)
	ok = dn.ravel() != 0
	cx = (x-Kn[0, 2])*z/Kn[0, 0]; cy = (y-Kn[1, 2])*z/Kn[1, 1]
	X = [((Rn[0, i]*cx+Rn[1, i]*cy)+Rn[2, i]*z)+Cn[i] for i in range(3)]
	t = [X[i]-Cr[i] for i in range(3)]
	c = [(Rr[i, 0]*t[0]+Rr[i, 1]*t[1])+Rr[i, 2]*t[2] for i in range(3)]
	ok &= c[2] > 0
	with np.errstate(all="ignore"):
		u = Kr[0, 2]+Kr[0, 0]*(c[0]/c[2]); v = Kr[1, 2]+Kr[1, 1]*(c[1]/c[2])
	zf = c[2].astype(np.float32)
	cand = []
	for px, py in ((np.floor(u), np.floor(v)), (np.floor(u), np.ceil(v)), (np.ceil(u), np.floor(v)), (np.ceil(u), np.ceil(v))):
		m = ok & (px >= 0) & (py >= 0) & (px < W) & (py < H)
		cand.append(np.stack([(py[m]*W+px[m]).astype(np.int64), idx[m]], 1))
	cand = np.concatenate(cand)
	# z-buffer: per target the smallest depth; among equal depths the last source pixel
	order = np.lexsort((-cand[:, 1], zf[cand[:, 1]], cand[:, 0]))
	cand = cand[order]
	first = np.ones(len(cand), bool); first[1:] = cand[1:, 0] != cand[:-1, 0]
	pd = np.zeros(H*W, np.float32); pc = np.zeros(H*W, np.float32)
	pd[cand[first, 0]] = zf[cand[first, 1]]
	if cn is not None:
		pc[cand[first, 0]] = cn.ravel()[cand[first, 1]]
	return pd.reshape(H, W), pc.reshape(H, W)


def test_projection_matches_numpy_transcription(scene):
	sc, maps = scene
	for n in (0, 2, 3):
		pd, pc = O.filter_project(_dm(sc, maps, 1), _dm(sc, maps, n))
		qd, qc = _project_np(_dm(sc, maps, 1), _dm(sc, maps, n))
		assert np.array_equal(pd, qd) and np.array_equal(pc, qc)
		assert (pd > 0).mean() > 0.7


def test_projection_of_ground_truth_lands_on_ground_truth(scene):
	"""camera convention pin: a neighbour's exact depth-map, projected, agrees with the reference view's exact depth"""
	sc, _ = scene
	gt = [(v.depth_gt, np.ones_like(v.depth_gt), v.K, v.R, v.C) for v in sc.views]
	pd, _ = O.filter_project(gt[1], gt[2])
	m = pd > 0
	rel = np.abs(pd[m]-gt[1][0][m])/gt[1][0][m]
	assert m.mean() > 0.8 and np.median(rel) < 2e-3 and (rel < 0.01).mean() > 0.97


def _similar(d0, d1, th):
	return f32(abs(f32(d0-d1)))/f32(d0) < th


def _filter_adjust_py(ref, nbrs, proj, pconf, nMinViews, nMinViewsAdjust, th, dmin, dmax):
	"""line-by-line transcription of SceneDensify.cpp:1141-1210 in numpy float32 scalars"""
	dr, cr, Kr, Rr, Cr = ref
	H, W = dr.shape
	N = len(nbrs)
	th = f32(th*f32(1.2)) if False else f32(f32(th)*f32(1.2))
	od = np.zeros((H, W), f32); oc = np.zeros((H, W), f32)
	for i in range(H):
		for j in range(W):
			depth = dr[i, j]
			if depth == 0:
				continue
			posConf = cr[i, j]; negConf = f32(0)
			avg = f32(depth*posConf)
			nPos = nNeg = 0
			n = N
			discard = False
			while True:
				n -= 1
				d = proj[n, i, j]
				if d == 0:
					if nPos+nNeg+n < nMinViews:
						discard = True
						break
				elif _similar(depth, d, th):
					c = pconf[n, i, j]
					avg = f32(avg+f32(d*c)); posConf = f32(posConf+c); nPos += 1
				else:
					if depth > d:
						negConf = f32(negConf+pconf[n, i, j])
					else:
						dn, cn, Kn, Rn, Cn = nbrs[n]
						z = float(depth)
						cam = np.array([(j-Kr[0, 2])*z/Kr[0, 0], (i-Kr[1, 2])*z/Kr[1, 1], z])
						X = np.array([((Rr[0, k]*cam[0]+Rr[1, k]*cam[1])+Rr[2, k]*cam[2])+Cr[k] for k in range(3)])
						t = X-Cn
						c3 = np.array([(Rn[k, 0]*t[0]+Rn[k, 1]*t[1])+Rn[k, 2]*t[2] for k in range(3)])
						u = Kn[0, 2]+Kn[0, 0]*(c3[0]/c3[2]); v = Kn[1, 2]+Kn[1, 1]*(c3[1]/c3[2])
						rx, ry = int(np.floor(u+.5)), int(np.floor(v+.5))
						if 0 <= rx < dn.shape[1] and 0 <= ry < dn.shape[0]:
							c = cn[ry, rx]
							negConf = f32(negConf+(c if c > 0 else pconf[n, i, j]))
						else:
							negConf = f32(negConf+pconf[n, i, j])
					nNeg += 1
				if n == 0:
					break
			if discard:
				continue
			if nPos >= nMinViewsAdjust and posConf > negConf:
				avg = f32(avg/posConf)
				if dmin <= avg < dmax:
					od[i, j] = avg; oc[i, j] = f32(posConf-negConf)
	return od, oc


def _filter_strict_py(ref, proj, nMinViews, th):
	"""transcription of SceneDensify.cpp:1211-1289 (out-of-image neighbours count as empty)"""
	dr, cr = ref[0], ref[1]
	H, W = dr.shape
	N = proj.shape[0]
	thS = f32(f32(th)*f32(0.8)); thD = f32(f32(th)*f32(1.2))
	od = np.zeros((H, W), f32); oc = np.zeros((H, W), f32)
	for i in range(H):
		for j in range(W):
			depth = dr[i, j]
			if depth == 0:
				continue
			good = views = 0
			for n in range(N):
				d = proj[n, i, j]
				if d > 0:
					views += 1; good += bool(_similar(depth, d, thS))
			if good < nMinViews or good < views*75//100:
				continue
			good = views = 0
			for dx, dy in ((-1, 0), (1, 0), (0, -1), (0, 1)):
				x, y = j+dx, i+dy
				if not (0 <= x < W and 0 <= y < H):
					continue
				for n in range(N):
					d = proj[n, y, x]
					if d > 0:
						views += 1; good += bool(_similar(depth, d, thD))
			if good < nMinViews*2 or good < views*65//100:
				continue
			od[i, j] = depth; oc[i, j] = cr[i, j]
	return od, oc


def test_filter_adjust_matches_python_transcription(scene):
	sc, maps = scene
	ref = _dm(sc, maps, 1); nbrs = [_dm(sc, maps, i) for i in (0, 2, 3)]
	ok, od, oc, proj = O.filter_depth_map(ref, nbrs, 2, 1, 0.01, True, sc.dmin, sc.dmax)
	assert ok
	pconf = np.stack([O.filter_project(ref, nb)[1] for nb in nbrs])
	qd, qc = _filter_adjust_py(ref, nbrs, proj, pconf, 2, 1, 0.01, f32(sc.dmin), f32(sc.dmax))
	assert np.array_equal(od, qd) and np.array_equal(oc, qc)
	# behaviour: most inliers survive and move towards the truth, gross outliers go
	gt = sc.views[1].depth_gt
	inl = (ref[0] > 0) & (np.abs(ref[0]-gt)/gt < 0.006)
	out = (ref[0] > 0) & (np.abs(ref[0]-gt)/gt > 0.05)
	assert (od[inl] > 0).mean() > 0.85 and (od[out] > 0).mean() < 0.05
	k = od > 0
	# (at this resolution the nearest-of-4 splat biases the projected depths by about a pixel's depth step,
	# so the average is not closer to the truth than the input; it must stay in the same range)
	assert np.abs(od[k]-gt[k]).mean() < 1.5*np.abs(ref[0][k]-gt[k]).mean()
	assert (np.abs(od[k]-gt[k])/gt[k]).max() < 0.02


def test_filter_strict_matches_python_transcription(scene):
	sc, maps = scene
	ref = _dm(sc, maps, 1); nbrs = [_dm(sc, maps, i, conf=False) for i in (0, 2, 3)]
	ok, od, oc, proj = O.filter_depth_map(ref, nbrs, 2, 1, 0.01, False, sc.dmin, sc.dmax)
	assert ok
	qd, qc = _filter_strict_py(ref, proj, 2, 0.01)
	assert np.array_equal(od, qd) and np.array_equal(oc, qc)
	assert 0.2 < (od > 0).mean() < 0.98
	assert np.array_equal(od[od > 0], ref[0][od > 0])


def test_filter_refuses_too_few_neighbours(scene):
	sc, maps = scene
	ok, _, _, _ = O.filter_depth_map(_dm(sc, maps, 1), [_dm(sc, maps, 0)], 2, 1, 0.01, True, sc.dmin, sc.dmax)
	assert not ok


# ---- RemoveSmallSegments ----
def test_small_segments_known_answers():
	d = np.full((40, 60), 5.0, f32)
	d[5:10, 5:15] = 6.0      # 50-pixel island at another depth: removed (50 < 100)
	d[20:30, 30:40] = 7.0    # 100-pixel island: kept (100 is not < 100)
	d[0:3, 50:53] = 0        # invalid pixels stay invalid
	n = np.zeros((40, 60, 3), f32); n[..., 2] = -1
	c = np.full((40, 60), 0.5, f32)
	od, on, oc = O.remove_small_segments(d, n, c, 0.007, 100)
	assert (od[5:10, 5:15] == 0).all() and (on[5:10, 5:15] == 0).all() and (oc[5:10, 5:15] == 0).all()
	assert (od[20:30, 30:40] == 7.0).all() and (od[0:3, 50:53] == 0).all()
	keep = np.ones_like(d, bool); keep[5:10, 5:15] = False; keep[0:3, 50:53] = False
	assert np.array_equal(od[keep], d[keep]) and (oc[keep] == 0.5).all()
	# invalid pixels are one-pixel segments: their normal / confidence are cleared as well
	assert (on[0:3, 50:53] == 0).all() and (oc[0:3, 50:53] == 0).all()


def component_sizes(d, th, both):
	"""per pixel, the size of its connected component over 4-neighbour edges whose IsDepthSimilar test
	holds in both directions (both=True) or in at least one (both=False)"""
	from scipy.sparse import coo_matrix
	from scipy.sparse.csgraph import connected_components
	H, W = d.shape
	idx = np.arange(H*W).reshape(H, W)
	def sim(a, b):
		with np.errstate(all="ignore"):
			ok = (a > 0) & (b > 0)
			s1 = np.abs(a-b)/a < th; s2 = np.abs(b-a)/b < th
		return ok & ((s1 & s2) if both else (s1 | s2))
	eh = sim(d[:, :-1], d[:, 1:]); ev = sim(d[:-1], d[1:])
	r = np.concatenate([idx[:, :-1][eh], idx[:-1][ev]]); c = np.concatenate([idx[:, 1:][eh], idx[1:][ev]])
	_, lab = connected_components(coo_matrix((np.ones(len(r)), (r, c)), shape=(H*W, H*W)), directed=False)
	return np.bincount(lab)[lab].reshape(H, W)


def segment_test_map(sc, seed=1):
	"""smooth surface cut into segments of many sizes by hole lines and depth steps (no noise: no
	similarity test sits near its threshold, so the segments do not depend on the traversal order)"""
	rng = np.random.RandomState(seed)
	d = sc.views[1].depth_gt.copy()
	H, W = d.shape
	for _ in range(14):
		y, x, h, w = rng.randint(0, H-4), rng.randint(0, W-4), rng.randint(2, 14), rng.randint(2, 14)
		d[y:y+h, x:x+w] *= f32(1.0+0.05*rng.randint(1, 4))
	for _ in range(6):
		if rng.rand() < 0.5:
			d[rng.randint(0, H), :] = 0
		else:
			d[:, rng.randint(0, W)] = 0
	return d


def test_small_segments_match_connected_components(scene):
	"""without direction-dependent edges the reference's breadth-first segments are the connected components"""
	sc, _ = scene
	d = segment_test_map(sc)
	th = f32(0.007)
	assert O.count_asymmetric_edges(d, th) == 0
	size = component_sizes(d, th, True)
	exp = np.where(size < 100, 0, d).astype(f32)
	od, _, _ = O.remove_small_segments(d, None, None, th, 100)
	assert np.array_equal(od, exp)
	assert 0 < (od > 0).sum() < (d > 0).sum()


def test_small_segments_bounds_with_direction_dependent_edges(scene):
	"""noisy map: the breadth-first segments lie between the components over two-way edges and those over
	one-way-or-two-way edges, so a pixel in a two-way component >= speckle is kept and a pixel in a
	one-way component < speckle is removed; only the pixels in between depend on the traversal order"""
	sc, maps = scene
	d = maps[1][0]
	th = f32(0.007)
	assert O.count_asymmetric_edges(d, th) > 0
	strong = component_sizes(d, th, True); weak = component_sizes(d, th, False)
	for speckle in (20, 100):
		od, _, _ = O.remove_small_segments(d, None, None, th, speckle)
		v = d > 0
		assert (od[v & (strong >= speckle)] > 0).all()
		assert (od[v & (weak < speckle)] == 0).all()
		assert (v & (strong < speckle) & (weak >= speckle)).mean() < 0.05


# ---- GapInterpolation ----
def _gap_py(depth, normal, conf, th, gap):
	"""transcription of SceneDensify.cpp:904-1045 in numpy float32 scalars"""
	d = depth.copy(); n = None if normal is None else normal.copy(); c = None if conf is None else conf.copy()
	H, W = d.shape
	def n2d(v):
		return f32(np.arctan2(v[1], v[0])), f32(np.arccos(v[2]))
	def d2n(a, b):
		sy = f32(np.sin(b))
		return np.array([f32(np.cos(a))*sy, f32(np.sin(a))*sy, f32(np.cos(b))], f32)
	for rows in (True, False):
		for l in range(H if rows else W):
			at = (lambda k: (l, k)) if rows else (lambda k: (k, l))
			count = 0
			for u in range(W if rows else H):
				d1 = d[at(u)]
				if d1 <= 0:
					count += 1
					continue
				if count == 0:
					continue
				if count <= gap and u > count:
					uc = u-count; uf = uc-1
					d0 = d[at(uf)]
					if _similar(d0, d1, th):
						diff = f32(f32(d1-d0)/f32(count+1))
						cur = d0
						cc = min(c[at(uf)], c[at(u)]) if c is not None else None
						if n is not None:
							a1, b1 = n2d(n[at(uf)]); a2, b2 = n2d(n[at(u)])
							da = f32(f32(a2-a1)/f32(count+1)); db = f32(f32(b2-b1)/f32(count+1))
						while uc < u:
							cur = f32(cur+diff); d[at(uc)] = cur
							if n is not None:
								a1 = f32(a1+da); b1 = f32(b1+db); n[at(uc)] = d2n(a1, b1)
							if c is not None:
								c[at(uc)] = cc
							uc += 1
				count = 0
	return d, n, c


def test_gap_interpolation_known_answers():
	d = np.zeros((3, 12), f32)
	d[0] = [0, 0, 1.0, 0, 0, 0, 1.02, 5, 0, 0, 0, 0]          # leading / trailing gaps stay, the inner one is filled
	d[1] = [1.0, 0, 0, 0, 1.04, 2, 0, 0, 0, 0, 0, 0]           # ends too different (4 % > 2.5 %): not filled
	d[2] = [1.0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 3, 3]            # 8 invalid pixels > gap size 7: not filled
	c = np.where(d > 0, 0.8, 0).astype(f32); c[0, 6] = 0.3
	od, _, oc = O.gap_interpolation(d, None, c, 0.025, 7)
	diff = f32(f32(f32(1.02)-f32(1.0))/f32(4))
	e1 = f32(f32(1.0)+diff); e2 = f32(e1+diff); e3 = f32(e2+diff)
	assert list(od[0, 3:6]) == [e1, e2, e3] and (oc[0, 3:6] == f32(0.3)).all()
	# the column pass may only have touched columns whose rows 0 and 2 are valid with row 1 empty
	assert (od[0, :2] == 0).all() and (od[0, 8:] == 0).all()
	assert (od[1, 1:4] == 0).all()
	assert (od[2, 1:9] == 0).all()


def test_gap_interpolation_matches_python_transcription(scene):
	sc, maps = scene
	rng = np.random.RandomState(5)
	v = sc.views[1]
	d = v.depth_gt.copy(); n = v.normal_gt.copy(); c = rng.uniform(0.1, 1, d.shape).astype(f32)
	# punch gaps of assorted lengths in rows and columns
	for _ in range(60):
		y, x, l = rng.randint(0, d.shape[0]), rng.randint(0, d.shape[1]), rng.randint(1, 11)
		if rng.rand() < 0.5:
			d[y, x:x+l] = 0
		else:
			d[y:y+l, x] = 0
	n[d == 0] = 0; c[d == 0] = 0
	od, on, oc = O.gap_interpolation(d, n, c, 0.025, 7)
	qd, qn, qc = _gap_py(d, n, c, f32(0.025), 7)
	assert np.array_equal(od, qd) and np.array_equal(oc, qc)
	assert np.abs(on-qn).max() < 1e-6
	filled = (d == 0) & (od > 0)
	assert filled.sum() > 50 and (d == 0).sum() > filled.sum()
	assert np.abs(od[filled]-v.depth_gt[filled]).max()/v.depth_gt.mean() < 5e-3
	assert np.abs(np.linalg.norm(on[filled], axis=1)-1).max() < 1e-5


# ---- toGray (image preparation before the estimation) ----
def test_to_gray_known_answers():
	rng = np.random.RandomState(2)
	img = rng.randint(0, 256, (37, 53, 3)).astype(np.uint8)
	inv = f32(1.0)/f32(255.0)
	c = img.astype(f32)*inv
	exp = (f32(0.114)*c[..., 0]+f32(0.587)*c[..., 1])+f32(0.299)*c[..., 2]   # numpy float32: one rounding per operation
	assert np.array_equal(O.to_gray(img, True), exp)
	assert np.array_equal(O.to_gray(img[..., ::-1], False), (f32(0.299)*c[..., 2]+f32(0.587)*c[..., 1])+f32(0.114)*c[..., 0])
	bgra = np.concatenate([img, rng.randint(0, 256, (37, 53, 1)).astype(np.uint8)], -1)
	assert np.array_equal(O.to_gray(bgra, True), exp)                          # the fourth channel is ignored
	g = O.to_gray(np.array([[[255, 255, 255], [0, 0, 0], [255, 0, 0]]], np.uint8))
	assert abs(g[0, 0]-1.0) < 1e-6 and g[0, 1] == 0 and abs(g[0, 2]-0.114) < 1e-6
	# within float rounding of the double-precision value the fixture generator used
	ref64 = (0.114*img[..., 0].astype(np.float64)+0.587*img[..., 1]+0.299*img[..., 2])/255.0
	assert np.abs(O.to_gray(img)-ref64).max() < 2e-7
