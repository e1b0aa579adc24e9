"""Known-answer and property tests that pin the CPU oracle (no GPU).

The reference ships no golden vectors for this path (apps/Tests/Tests.cpp only asserts a
fused point count), so the oracle is pinned by (a) published constants (Philox KAT),
(b) hand-derived answers of the small deterministic pieces, (c) analytic ground truth of
the synthetic scenes and (d) OpenCV (cv2) for the cv::resize restatements.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from conftest import agreement


def test_philox_known_answers():
	# Random123 kat_vectors, philox4x32-10
	assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
	assert O.philox([0xffffffff]*4, [0xffffffff]*2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
	assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
		[0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_zigzag_order_3x3_and_bands():
	# MapMatrix2ZigzagIdx (libs/MVS/DepthMap.cpp:329-356) walks anti-diagonals top-right to
	# bottom-left: row-major labels 1 2 4 3 5 7 6 8 9 for a 3x3 band
	buf = (C.c_uint16*(9*2))()
	O.lib().oracle_zigzag(3, 3, 16, buf)
	xy = np.array(buf).reshape(-1, 2)
	labels = (xy[:, 1]*3+xy[:, 0]+1).tolist()
	assert labels == [1, 2, 4, 3, 5, 7, 6, 8, 9]
	# bands: 7 rows with rawStride 3 -> bands of 3 and 4 rows (the last band absorbs the rest)
	buf = (C.c_uint16*(5*7*2))()
	O.lib().oracle_zigzag(5, 7, 3, buf)
	xy = np.array(buf).reshape(-1, 2)
	assert len({(int(a), int(b)) for a, b in xy}) == 35           # every pixel exactly once
	assert xy[:15, 1].max() == 2 and xy[15:, 1].min() == 3        # first band rows 0..2
	# causality: left and up neighbours are always visited earlier
	order = {(int(a), int(b)): i for i, (a, b) in enumerate(xy)}
	for (x, y), i in order.items():
		if x > 0: assert order[(x-1, y)] < i
		if y > 0: assert order[(x, y-1)] < i


def test_dir_normal_roundtrip_and_known_values():
	n = (C.c_float*3)()
	O.lib().oracle_dir2normal(C.c_float(0.0), C.c_float(np.pi/2), n)
	assert np.allclose(list(n), [1, 0, 0], atol=1e-6)
	O.lib().oracle_dir2normal(C.c_float(np.pi/2), C.c_float(np.pi), n)
	assert np.allclose(list(n), [0, 0, -1], atol=1e-6)
	rng = np.random.RandomState(0)
	for _ in range(50):
		v = rng.normal(size=3); v /= np.linalg.norm(v)
		a, b = C.c_float(), C.c_float()
		vv = (C.c_float*3)(*v)
		O.lib().oracle_normal2dir(vv, C.byref(a), C.byref(b))
		O.lib().oracle_dir2normal(a, b, n)
		assert np.allclose(list(n), v, atol=2e-6)


def test_correct_normal_faces_camera():
	rng = np.random.RandomState(1)
	for _ in range(100):
		X0 = (C.c_double*3)(rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), 1.0)
		v = rng.normal(size=3); v /= np.linalg.norm(v)
		n = (C.c_float*3)(*v)
		O.lib().oracle_correct_normal(X0, n)
		out = np.array(list(n))
		assert abs(np.linalg.norm(out)-1) < 1e-5
		if np.dot(v, list(X0)) < 0:
			assert np.allclose(out, v, atol=1e-7)       # already facing the camera: untouched
		else:
			assert np.dot(out, list(X0)) <= 1e-6         # rotated to (just past) 90 degrees


def test_interpolate_pixel_is_ray_plane_intersection():
	K = np.array([[500.0, 0, 159.5], [0, 500.0, 119.5], [0, 0, 1]])
	Kc = (C.c_double*9)(*K.ravel())
	n = np.array([0.2, -0.1, -0.97]); n /= np.linalg.norm(n)
	nc = (C.c_float*3)(*n)
	d = 5.0
	for (x0, y0, nx, ny) in ((100, 80, 99, 80), (100, 80, 100, 79), (30, 200, 31, 200), (30, 200, 30, 201)):
		got = O.lib().oracle_interpolate_pixel(Kc, x0, y0, nx, ny, C.c_float(d), nc, C.c_float(1.0), C.c_float(50.0))
		# the reference intersects in the x (or y) plane only: depth' = d (nz + x1 n_a)/(nz + x0' n_a)
		if y0 == ny:
			a0, a1, na = (x0-K[0, 2])/K[0, 0], (nx-K[0, 2])/K[0, 0], n[0]
		else:
			a0, a1, na = (y0-K[1, 2])/K[1, 1], (ny-K[1, 2])/K[1, 1], n[1]
		want = d*(n[2]+a1*na)/(n[2]+a0*na)
		assert abs(got-want) < 1e-5*want
	# out-of-range result keeps the neighbour's depth
	got = O.lib().oracle_interpolate_pixel(Kc, 100, 80, 99, 80, C.c_float(d), nc, C.c_float(5.5), C.c_float(50.0))
	assert got == pytest.approx(d)


def test_score_pixel_at_ground_truth_is_near_zero(small_scene):
	sc, ref, views = small_scene
	prm = O.default_params(nSubResolutionLevels=0)
	gt_d, gt_n = sc.views[ref].depth_gt, sc.views[ref].normal_gt
	rng = np.random.RandomState(3)
	good, bad = [], []
	for _ in range(40):
		x, y = int(rng.randint(30, 290)), int(rng.randint(30, 210))
		s, vs = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, float(gt_d[y, x]), gt_n[y, x])
		good.append(s)
		s2, _ = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, float(gt_d[y, x])*1.05, gt_n[y, x])
		bad.append(s2)
		assert len(vs) == 4 and np.all(vs >= 0) and np.all(vs <= 2)
		# MINMEAN: mean of the two smallest view scores (DepthMap.cpp:595-610)
		two = np.sort(vs)[:2]
		assert s == pytest.approx(two.mean() if two[1] < 1.2 else two[0], abs=1e-6)
	assert np.median(good) < 0.02 and np.median(bad) > 0.3


def test_score_pixel_smoothness_and_outside():
	# a hypothesis whose patch leaves the neighbour image scores thRobust = 0.9*4/3 in that view
	from openmvs_b200 import synth
	sc = synth.make_scene(160, 120, 2, step_deg=30.0, cols=2)
	prm = O.default_params(nSubResolutionLevels=0)
	v0, v1 = sc.views
	found = 0
	for x in (6, 20, 80, 140, 153):
		y, d = 60, float(v0.depth_gt[60, x])
		X = v0.R.T @ (np.array([(x-v0.K[0, 2])/v0.K[0, 0], (y-v0.K[1, 2])/v0.K[1, 1], 1.0])*d) + v0.C
		p = v1.K @ (v1.R @ (X-v1.C)); p = p[:2]/p[2]
		if p[0] < -12 or p[0] > 160+12:
			s, vs = O.pm_score_pixel(sc.views, prm, sc.dmin, sc.dmax, x, y, d, v0.normal_gt[y, x])
			assert vs[0] == pytest.approx(1.2) and s == pytest.approx(1.2)
			found += 1
	assert found > 0
	# smoothness bonus: a neighbour on the same plane multiplies the score by (1-.07)(1-.0672)
	sc = synth.make_scene(160, 120, 2, step_deg=5.0, cols=2)
	x, y = 80, 60
	d, n = float(sc.views[0].depth_gt[y, x])*1.01, sc.views[0].normal_gt[y, x]
	s0, _ = O.pm_score_pixel(sc.views, prm, sc.dmin, sc.dmax, x, y, d, n)
	K = sc.views[0].K
	X0 = np.array([(x-K[0, 2])/K[0, 0], (y-K[1, 2])/K[1, 1], 1.0])
	ray = np.array([(x-1-K[0, 2])/K[0, 0], (y-K[1, 2])/K[1, 1], 1.0])
	dn = d*np.dot(n, X0)/np.dot(n, ray)  # neighbour point on the hypothesis plane
	close = np.concatenate([[dn], n, ray*dn]).astype(np.float32)
	s1, _ = O.pm_score_pixel(sc.views, prm, sc.dmin, sc.dmax, x, y, d, n, close=close)
	assert s1 == pytest.approx(s0*(1-0.07)*(1-0.07*0.96), rel=1e-4)


def test_resize_restatements_match_opencv():
	cv2 = pytest.importorskip("cv2")
	rng = np.random.RandomState(5)
	lib = O.lib()
	def fp(a): return a.ctypes.data_as(C.c_void_p)
	for (sw, sh) in ((64, 48), (479, 321), (101, 77)):
		src = rng.rand(sh, sw).astype(np.float32)
		for scale in (0.5, 0.25):
			dw, dh = int(np.rint(sw*scale)), int(np.rint(sh*scale))
			want = cv2.resize(src, None, fx=scale, fy=scale, interpolation=cv2.INTER_AREA)
			assert want.shape == (dh, dw)
			got = np.zeros((dh, dw), np.float32)
			lib.oracle_resize_area(fp(src), sw, sh, fp(got), dw, dh, C.c_double(1/scale), C.c_double(1/scale))
			assert np.abs(got-want).max() < 2e-6
			# destination-size form (general area weights when the ratio is not an integer)
			want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
			lib.oracle_resize_area(fp(src), sw, sh, fp(got), dw, dh, C.c_double(0), C.c_double(0))
			assert np.abs(got-want).max() < 2e-6
		low = rng.rand(sh//2, sw//2).astype(np.float32)
		want = cv2.resize(low, (sw, sh), interpolation=cv2.INTER_LINEAR)
		got = np.zeros((sh, sw), np.float32)
		lib.oracle_resize_linear(fp(low), sw//2, sh//2, fp(got), sw, sh)
		# OpenCV's dispatch (generic vs SIMD/IPP) rounds the tap fractions differently by a few
		# float ulps of the source coordinate, hence 2e-5 rather than 1 ulp of the value
		assert np.abs(got-want).max() < 2e-5
		low3 = rng.rand(sh//2, sw//2, 3).astype(np.float32)
		want = cv2.resize(low3, (sw, sh), interpolation=cv2.INTER_NEAREST)
		got = np.zeros((sh, sw, 3), np.float32)
		lib.oracle_resize_nearest(fp(low3), sw//2, sh//2, 3, fp(got), sw, sh)
		assert np.array_equal(got, want)


def test_scale_K_half_pixel_convention():
	K = np.array([[600.0, 0, 319.5], [0, 600.0, 239.5], [0, 0, 1]])
	out = (C.c_double*9)()
	O.lib().oracle_scale_K((C.c_double*9)(*K.ravel()), 640, 480, 320, 240, out)
	# Camera::ScaleK (libs/MVS/Camera.h:160-173): c' = (c+0.5) s - 0.5
	assert np.allclose(list(out), [300, 0, 159.5, 0, 300, 119.5, 0, 0, 1])


def test_zz_reference_schedule_converges_to_ground_truth(tiny_scene):
	sc, ref, views = tiny_scene
	prm = O.default_params(schedule=0, nEstimationIters=3, nSubResolutionLevels=0, nEstimationGeometricIters=0)
	d, n, c = O.pm_estimate(views, prm, sc.dmin, sc.dmax)
	gt = sc.views[ref].depth_gt
	m = d > 0
	assert m.mean() > 0.80
	rel = np.abs(d-gt)[m]/gt[m]
	assert (rel < 1e-2).mean() > 0.93 and np.median(rel) < 1e-3
	# border of half a window is never estimated; confidence = 1-cost in (0,1]
	assert not m[:4].any() and not m[:, :4].any() and not m[-4:].any() and not m[:, -4:].any()
	assert c[m].min() > 0 and c.max() <= 1 and np.all(c[~m] == 0) and np.all(n[~m] == 0)
	nn = np.linalg.norm(n[m], axis=-1)
	assert np.allclose(nn, 1, atol=1e-4)


def test_rb_schedule_reaches_the_zz_fixed_point(tiny_scene):
	"""Red-black (12 sweeps of 4 propagations + 3 refinements) against the reference's zig-zag
	schedule from the same random initialisation: same confidence mask and the same depths up
	to the reference's own run-to-run variation (ZZ with another thread count / seed)."""
	sc, ref, views = tiny_scene
	base = dict(nSubResolutionLevels=0, nEstimationGeometricIters=0)
	zz1 = O.pm_estimate(views, O.default_params(schedule=0, nEstimationIters=4, threads=1, **base), sc.dmin, sc.dmax)
	zz4 = O.pm_estimate(views, O.default_params(schedule=0, nEstimationIters=4, threads=4, **base), sc.dmin, sc.dmax)
	rb = O.pm_estimate(views, O.default_params(schedule=1, propagation=4, nEstimationIters=12, nRandomIters=3, threads=4, **base), sc.dmin, sc.dmax)
	iou_ref, agree_ref = agreement(zz1[0], zz4[0])
	iou_rb, agree_rb = agreement(zz1[0], rb[0])
	assert iou_ref > 0.99 and iou_rb > 0.99
	assert agree_ref > 0.85
	assert agree_rb > agree_ref-0.05


def test_rb_is_deterministic_and_thread_independent(tiny_scene):
	sc, ref, views = tiny_scene
	base = dict(schedule=1, propagation=4, nEstimationIters=2, nRandomIters=3, nSubResolutionLevels=0, nEstimationGeometricIters=0)
	a = O.pm_estimate(views, O.default_params(threads=1, **base), sc.dmin, sc.dmax)
	b = O.pm_estimate(views, O.default_params(threads=4, **base), sc.dmin, sc.dmax)
	for x, y in zip(a, b):
		assert np.array_equal(x, y)


def _patch_stats(img, x, y):
	"""FillPixelPatch in numpy (DepthMap.cpp:422-462): weights, weighted mean, normSq0"""
	taps = [(i, j) for i in range(-4, 5, 2) for j in range(-4, 5, 2)]
	I = np.array([img[y+i, x+j] for i, j in taps], np.float64)
	w = np.exp(-(I-img[y, x])**2/(2*0.1**2)-np.array([i*i+j*j for i, j in taps])/(2*3.0**2))
	tm = (I*w).sum()/w.sum()
	return w, I, tm, float((w*(I-tm)**2).sum())


def test_low_resolution_prior_blend_known_answer(small_scene):
	"""score' = (1-f) score + f min(|d0-d|/d0, 0.5) with f = exp(-normSq0/0.02) (DepthMap.cpp:552-561);
	normSq0 is recomputed here in numpy, which also pins FillPixelPatch."""
	sc, ref, views = small_scene
	prm = O.default_params(nSubResolutionLevels=0)
	gt_d, gt_n = sc.views[ref].depth_gt, sc.views[ref].normal_gt
	h, w = gt_d.shape
	rng = np.random.RandomState(9)
	checked = 0
	for _ in range(12):
		x, y = int(rng.randint(30, 290)), int(rng.randint(30, 210))
		_, _, _, nsq = _patch_stats(views[0].image.astype(np.float64), x, y)
		f = np.exp(-nsq/0.02)
		d = float(gt_d[y, x])*1.002
		s0, v0 = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d, gt_n[y, x])
		for d0 in (d, d*1.25, d*3.0):
			prior = np.zeros((h, w), np.float32); prior[y, x] = d0
			s1, v1 = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d, gt_n[y, x], lowres=prior)
			want = (1-f)*v0 + f*min(abs(d0-d)/d0, 0.5)
			assert np.allclose(v1, np.minimum(want, 2.0), atol=2e-5)
			checked += 1
	assert checked == 36


def test_geometric_consistency_term_known_answers(small_scene):
	"""score += 0.1 min(sqrt(dist (dist+2)), 4) (DepthMap.cpp:535-551): consistent neighbour depth-maps add
	almost nothing, depth-maps without a similar depth add the full 0.1*4, a shifted hypothesis adds the
	analytic reprojection distance."""
	sc, ref, views = small_scene
	prm = O.default_params(nSubResolutionLevels=0, nEstimationGeometricIters=2)
	gt_d, gt_n = sc.views[ref].depth_gt, sc.views[ref].normal_gt
	good = [None]+[(v.depth_gt, v.K, v.R, v.C) for v in views[1:]]
	empty = [None]+[(np.zeros_like(v.depth_gt), v.K, v.R, v.C) for v in views[1:]]
	rng = np.random.RandomState(4)
	for _ in range(10):
		x, y = int(rng.randint(40, 280)), int(rng.randint(40, 200))
		d, n = float(gt_d[y, x]), gt_n[y, x]
		s0, v0 = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d, n)
		s1, v1 = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d, n, depths=good)
		s2, v2 = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d, n, depths=empty)
		assert np.all(v1-v0 >= -1e-6) and np.all(v1-v0 < 0.02)       # consistent: distance ~ 0
		assert np.allclose(v2-v0, 0.4, atol=1e-5)                     # no similar depth: consistency = 4
		# hypothesis 2 % in front of the surface: the neighbour's surface point re-projects off the pixel
		d2 = d*0.98
		sb, vb = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d2, n)
		sg, vg = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d2, n, depths=good)
		K0, R0, C0 = views[0].K, views[0].R, views[0].C
		X = R0.T @ (np.array([(x-K0[0, 2])/K0[0, 0], (y-K0[1, 2])/K0[1, 1], 1.0])*d2) + C0
		for k, v in enumerate(views[1:]):
			p1 = v.K @ (v.R @ (X-v.C)); z1 = p1[2]; u1 = p1[:2]/z1
			# depth of the neighbour's surface at u1 (bilinear on its ground truth), similar within 3 %
			lx, ly = int(u1[0]), int(u1[1]); ax, ay = u1[0]-lx, u1[1]-ly
			g = v.depth_gt
			dn = (g[ly, lx]*(1-ax)+g[ly, lx+1]*ax)*(1-ay)+(g[ly+1, lx]*(1-ax)+g[ly+1, lx+1]*ax)*ay
			if abs(z1-dn)/z1 >= 0.03:
				assert abs((vg[k]-vb[k])-0.4) < 1e-4
				continue
			Xb = v.R.T @ (np.linalg.inv(v.K) @ np.array([u1[0]*dn, u1[1]*dn, dn])) + v.C
			pb = K0 @ (R0 @ (Xb-C0)); ub = pb[:2]/pb[2]
			dist = np.hypot(x-ub[0], y-ub[1])
			want = 0.1*min(np.sqrt(dist*(dist+2)), 4.0)
			assert abs((vg[k]-vb[k])-want) < 5e-3


def test_score_pixel_image_against_textbook_homography(small_scene):
	"""ScorePixelImage (DepthMap.cpp:465-564) against an independent float64 evaluation: the textbook
	plane-induced homography K1 (R + t n^T / (n.X)) K0^-1, bilinear taps and the weighted NCC
	1 - num / sqrt(normSq0 normSq1) with the bilateral weights of FillPixelPatch."""
	sc, ref, views = small_scene
	prm = O.default_params(nSubResolutionLevels=0)
	gt_d, gt_n = sc.views[ref].depth_gt, sc.views[ref].normal_gt
	img0 = views[0].image.astype(np.float64)
	K0, R0, C0 = views[0].K, views[0].R, views[0].C
	rng = np.random.RandomState(12)
	n_checked = 0
	for _ in range(15):
		x, y = int(rng.randint(40, 280)), int(rng.randint(40, 200))
		d = float(gt_d[y, x])*(1+0.01*rng.randn())
		n = gt_n[y, x].astype(np.float64)+0.05*rng.randn(3); n /= np.linalg.norm(n)
		s, vs = O.pm_score_pixel(views, prm, sc.dmin, sc.dmax, x, y, d, n.astype(np.float32))
		w, I, tm, nsq0 = _patch_stats(img0, x, y)
		X0 = np.array([(x-K0[0, 2])/K0[0, 0], (y-K0[1, 2])/K0[1, 1], 1.0])
		c = float(n @ X0)*d
		for k, v in enumerate(views[1:]):
			Rrel = v.R @ R0.T; t = v.R @ (C0-v.C)
			Hm = v.K @ (Rrel + np.outer(t, n)/c) @ np.linalg.inv(K0)
			img1 = v.image.astype(np.float64)
			vals, inside = [], True
			for i in range(-4, 5, 2):
				for j in range(-4, 5, 2):
					p = Hm @ np.array([x+j, y+i, 1.0]); px, py = p[0]/p[2], p[1]/p[2]
					if not (1 <= px <= img1.shape[1]-2 and 1 <= py <= img1.shape[0]-2):
						inside = False; break
					lx, ly = int(px), int(py); ax, ay = px-lx, py-ly
					vals.append((img1[ly, lx]*(1-ax)+img1[ly, lx+1]*ax)*(1-ay)+(img1[ly+1, lx]*(1-ax)+img1[ly+1, lx+1]*ax)*ay)
				if not inside: break
			if not inside:
				assert vs[k] == pytest.approx(1.2)
				continue
			f = np.array(vals)
			sumw, sumsq, num = (f*w).sum(), (f*f*w).sum(), (f*w*(I-tm)).sum()
			nsq1 = sumsq-sumw**2/w.sum()
			ncc = np.clip(num/np.sqrt(nsq0*nsq1), -1, 1)
			assert abs(vs[k]-(1-ncc)) < 2e-4
			n_checked += 1
	assert n_checked > 40
