"""SGM oracle pins (no GPU): P2 table known values, and the 8-path aggregation against an
independent pure-Python restatement of pixelAccum on a tiny ragged problem."""
import numpy as np
import pytest

from oracle import oracle as O
from openmvs_b200 import synth


def test_p2_table_known_values():
	import ctypes as C
	out = (C.c_uint16*256)()
	O.lib().oracle_sgm_p2s(C.c_uint16(4), C.c_float(14.0), C.c_float(38.0), out)
	t = list(out)
	# P2s[i] = round(4 (1 + 14 exp(-i^2 / (2 38^2))))  (libs/MVS/SemiGlobalMatcher.cpp:518-524)
	assert t[0] == 60 and t[255] == 4 and t[38] == round(4*(1+14*np.exp(-0.5)))
	assert all(a >= b for a, b in zip(t, t[1:]))


def _py_aggregate(costs, px, vw, vh, lgray, w, P1, P2s):
	"""straight transcription of the formula L(d) = C(d) + min_dp(Lp(dp) + V(d,dp)) - min Lp over the
	intersection of ranges, per scanline, 8 directions (quadratic in D, tiny inputs only)"""
	acc = np.zeros(len(costs), np.int64)
	def walk(x, y, dx, dy):
		Lp, pr = {}, (0, 0)
		Ip = np.float32(0.5)
		while 0 <= x < vw and 0 <= y < vh:
			p = px[y*vw+x]
			lo, hi, idx = int(p["dmin"]), int(p["dmax"]), int(p["idx"])
			if lo < hi:
				I = lgray[y, x]
				P2 = P2s[abs(int(np.floor(np.float32(255)*np.float32(I-Ip)+np.float32(0.5))))]
				Ip = I
				ilo, ihi = max(pr[0], lo), min(pr[1], hi)
				Ls = {}
				if ilo >= ihi:
					for d in range(lo, hi):
						Ls[d] = int(costs[idx+d-lo])+P2
				else:
					m = min(Lp[dp] for dp in range(ilo, ihi))
					for d in range(lo, hi):
						best = min(Lp[dp]+(0 if dp == d else P1 if abs(dp-d) == 1 else P2) for dp in range(ilo, ihi))
						Ls[d] = int(costs[idx+d-lo])+best-m
				for d in range(lo, hi):
					acc[idx+d-lo] += Ls[d]
				Lp, pr = Ls, (lo, hi)
			x += dx; y += dy
	W, H = vw, vh
	for x in range(W): walk(x, 0, 0, 1)
	for y in range(H): walk(0, y, 1, 0)
	for x in range(W): walk(x, H-1, 0, -1)
	for y in range(H): walk(W-1, y, -1, 0)
	for x in range(W): walk(x, 0, 1, 1)
	for y in range(1, H): walk(0, y, 1, 1)
	for x in range(W-1): walk(x, 0, -1, 1)
	for y in range(H): walk(W-1, y, -1, 1)
	for x in range(1, W): walk(x, H-1, 1, -1)
	for y in range(H): walk(0, y, 1, -1)
	for x in range(W): walk(x, H-1, -1, -1)
	for y in range(H-1): walk(W-1, y, -1, -1)
	return acc


def test_aggregation_matches_independent_python_restatement():
	import ctypes as C
	w, h = 22, 17
	rng = np.random.RandomState(3)
	lg, lc, rg, d = synth.make_stereo_pair(w, h, d0=2.0, amp=1.5)
	vw, vh = w-6, h-6
	lo = rng.randint(-3, 2, (vh, vw)); hi = lo+rng.randint(1, 9, (vh, vw))
	invalid = rng.rand(vh, vw) < 0.1
	px, n = synth.sgm_pixel_map(w, h, lo, hi, invalid)
	costs = rng.randint(0, 256, n).astype(np.uint8)
	c, a, disp, cost = O.sgm_match(lg, lc, rg, px, n, costs=costs)
	out = (C.c_uint16*256)()
	O.lib().oracle_sgm_p2s(C.c_uint16(4), C.c_float(14.0), C.c_float(38.0), out)
	want = _py_aggregate(costs, px, vw, vh, lg, w, 3, list(out))
	assert np.array_equal(a.astype(np.int64), want)
	# WTA: first arg-min inside each pixel's range; invalid pixels keep NO_DISP / NO_ACCUMCOST
	for i in rng.choice(vw*vh, 40, replace=False):
		p = px[i]
		if p["dmin"] < p["dmax"]:
			seg = a[int(p["idx"]):int(p["idx"])+int(p["dmax"]-p["dmin"])]
			assert disp.ravel()[i] == p["dmin"]+int(np.argmin(seg)) and cost.ravel()[i] == seg.min()
		else:
			assert disp.ravel()[i] == 32767 and cost.ravel()[i] == 65535


def test_cost_stage_properties_and_disparity_accuracy():
	w, h = 160, 96
	lg, lc, rg, d = synth.make_stereo_pair(w, h)
	px, n = synth.sgm_pixel_map(w, h, 0, 32)
	c, a, disp, cost = O.sgm_match(lg, lc, rg, px, n)
	vol = c.reshape(h-6, w-6, 32)
	# windows that leave the right image cost 255 (SemiGlobalMatcher.cpp:959-963)
	assert np.all(vol[:, -1, 1:] == 255) and np.all(vol[:, w-6-40, :].min(-1) < 128)
	gt = d[3:-3, 3:-3]
	# the cost minimum sits at the true disparity, and SGM recovers it to within a pixel
	inner = np.s_[5:-5, 5:-40]
	assert (np.abs(vol.argmin(-1)-gt)[inner] <= 1).mean() > 0.9
	assert (np.abs(disp-gt)[inner] <= 1).mean() > 0.97
