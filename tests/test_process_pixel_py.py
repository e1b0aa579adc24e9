"""An independent pure-Python (float64) transcription of DepthEstimator::ProcessPixel
(libs/MVS/DepthMap.cpp:630-852) with ScorePixel / ScorePixelImage / InterpolatePixel / CorrectNormal,
run pixel by pixel on the red-black schedule with the same Philox stream, against the C++ oracle.
Red-black makes every pixel of a half-sweep a pure function of the input state, so single pixels
can be compared; float64 vs float32 only differ where an accept test is within rounding."""
import numpy as np
import pytest

from oracle import oracle as O

D2R = np.pi/180


class PyEstimator:
	def __init__(self, views, dmin, dmax, depth, normal, conf, sweep, nR, seed=1234):
		self.v = views; self.dmin, self.dmax = dmin, dmax
		self.depth, self.normal, self.conf = depth, normal, conf
		self.sweep, self.nR, self.seed = sweep, nR, seed
		self.img0 = views[0].image.astype(np.float64)
		self.K0, self.R0, self.C0 = views[0].K, views[0].R, views[0].C
		self.h, self.w = self.img0.shape
		self.keep = 0.9
		self.thSmall, self.thBig, self.thRand, self.thRobust = 0.9*0.66, 0.9*0.9, 0.9*1.1, 0.9*4/3
		self.sbD, self.sbN = 1-0.93, (1-0.93)*0.96
		self.sgD, self.sgN = -1/(2*0.02**2), -1/(2*(13*D2R)**2)

	def ray(self, x, y):
		return np.array([(x-self.K0[0, 2])/self.K0[0, 0], (y-self.K0[1, 2])/self.K0[1, 1], 1.0])

	def patch(self, x, y):
		taps = [(i, j) for i in range(-4, 5, 2) for j in range(-4, 5, 2)]
		I = np.array([self.img0[y+i, x+j] for i, j in taps])
		w = np.exp(-(I-self.img0[y, x])**2/(2*0.1**2)-np.array([i*i+j*j for i, j in taps])/18.0)
		tm = (I*w).sum()/w.sum()
		return w, w*(I-tm), float((w*(I-tm)**2).sum())

	def score_image(self, v, x, y, d, n, close):
		X0 = self.ray(x, y)
		c = float(n @ X0)*d
		H = v.K @ (v.R @ self.R0.T + np.outer(v.R @ (self.C0-v.C), n)/c) @ np.linalg.inv(self.K0)
		img1 = v.image.astype(np.float64)
		vals = []
		for i in range(-4, 5, 2):
			for j in range(-4, 5, 2):
				p = H @ np.array([x+j, y+i, 1.0]); px, py = p[0]/p[2], p[1]/p[2]
				if not (1 <= px <= img1.shape[1]-2 and 1 <= py <= img1.shape[0]-2):
					return self.thRobust
				lx, ly = int(px), int(py); ax, ay = px-lx, py-ly
				vals.append((img1[ly, lx]*(1-ax)+img1[ly, lx+1]*ax)*(1-ay)+(img1[ly+1, lx]*(1-ax)+img1[ly+1, lx+1]*ax)*ay)
		f = np.array(vals)
		nsq1 = (f*f*self.W).sum()-(f*self.W).sum()**2/self.W.sum()
		nrm = self.nsq0*nsq1
		if nrm <= 1e-16:
			return self.thRobust
		score = 1-np.clip((f*self.TW).sum()/np.sqrt(nrm), -1, 1)
		D = -d*float(n @ X0)
		for (cd, cn, cX) in close:
			fd = np.exp(((float(n @ cX)+D)/d)**2*self.sgD)
			ca = np.clip(float(n @ cn)/np.sqrt(float(n @ n)*float(cn @ cn)), -1, 1)
			fn = np.exp(np.arccos(ca)**2*self.sgN)
			score *= (1-self.sbD*fd)*(1-self.sbN*fn)
		return min(2.0, score)

	def score(self, x, y, d, n, close):
		s = sorted(self.score_image(v, x, y, d, n, close) for v in self.v[1:])
		if len(s) == 1 or s[1] >= self.thRobust:
			return s[0]
		return (s[0]+s[1])/2

	def interpolate(self, x, y, nx, ny, d, n):
		if x == nx:
			a0, a1, na = self.ray(x, y)[1], self.ray(nx, ny)[1], n[1]
		else:
			a0, a1, na = self.ray(x, y)[0], self.ray(nx, ny)[0], n[0]
		denom = n[2]+a0*na
		if abs(denom) < 1e-4:
			return d
		dn = d*(n[2]+a1*na)/denom
		return dn if self.dmin <= dn < self.dmax else d

	def correct_normal(self, n, vd):
		cal = float(n @ vd)
		if cal < 0:
			return n
		ang = min((np.arccos(cal/np.linalg.norm(vd))-np.pi/2)*1.01, -0.001)
		w = np.cross(n, vd); w /= np.linalg.norm(w)
		return n*np.cos(ang)+np.cross(w, n)*np.sin(ang)+w*float(w @ n)*(1-np.cos(ang))  # Rodrigues

	def draws(self, x, y, slot):
		r = O.philox([y*self.w+x, 1+self.sweep, slot, 0], [self.seed, 0xB200C0DE])
		return [np.float32(np.float32(u)/np.float32(4294967296.0)) for u in r]

	@staticmethod
	def dir2normal(a, b):
		return np.array([np.cos(a)*np.sin(b), np.sin(a)*np.sin(b), np.cos(b)])

	def random_plane(self, x, y, slot, vd):
		u = self.draws(x, y, slot)
		s = np.sqrt(self.dmin)+(np.sqrt(self.dmax)-np.sqrt(self.dmin))*u[0]
		n = self.dir2normal(np.pi*u[1], np.pi/2+(np.pi/2)*u[2])
		return s*s, (-n if float(n @ vd) > 0 else n)

	def process(self, x, y):
		w, h = self.w, self.h
		d0, n0, c0 = float(self.depth[y, x]), self.normal[y, x].astype(np.float64), float(self.conf[y, x])
		if not (4 <= x < w-4 and 4 <= y < h-4):
			return d0, n0, c0
		self.W, self.TW, self.nsq0 = self.patch(x, y)
		if self.nsq0 < 0.02**2:
			return d0, n0, c0
		vd = self.ray(x, y)
		order = [(-1, 0), (0, -1), (1, 0), (0, 1)] if self.sweep % 2 == 0 else [(1, 0), (0, 1), (-1, 0), (0, -1)]
		close, props = [], []
		for ox, oy in order:
			ok = (x > 4) if ox < 0 else (x < w-4) if ox > 0 else (y > 4) if oy < 0 else (y < h-4)
			if not ok:
				continue
			nx, ny = x+ox, y+oy
			nd = float(self.depth[ny, nx])
			if nd > 0:
				nn = self.normal[ny, nx].astype(np.float64)
				close.append((nd, nn, self.ray(nx, ny)*nd))
				props.append((nx, ny, nd, nn))
		conf, depth, normal = c0, d0, n0
		for nx, ny, nd, nn in props:
			if self.conf[ny, nx] >= self.keep:
				continue
			hd = self.interpolate(x, y, nx, ny, nd, nn)
			hn = self.correct_normal(nn, vd)
			s = self.score(x, y, hd, hn, close)
			if conf > s:
				conf, depth, normal = s, hd, hn
		idx, restarted = 0, False
		while True:
			if conf <= self.thSmall: idx = 2
			elif conf <= self.thBig: idx = 1
			elif conf >= self.thRand and not restarted:
				restarted = True; close = []; again = False
				for it in range(self.nR):
					hd, hn = self.random_plane(x, y, it, vd)
					s = self.score(x, y, hd, hn, close)
					if conf > s:
						conf, depth, normal = s, hd, hn
						if conf < self.thRand:
							again = True; break
				if again:
					continue
				return depth, normal, conf
			break
		scale = 0.5**idx
		drange = depth*0.003
		pa, pb = np.arctan2(normal[1], normal[0]), np.arccos(normal[2])
		for it in range(self.nR):
			u = self.draws(x, y, self.nR+it)
			hd = depth+drange*scale*(2*u[0]-1)
			if not (self.dmin <= hd < self.dmax):
				continue
			na, nb = pa+16*D2R*scale*(2*u[1]-1), pb+10*D2R*scale*(2*u[2]-1)
			hn = self.dir2normal(na, nb)
			if float(hn @ vd) >= 0:
				continue
			s = self.score(x, y, hd, hn, close)
			if conf > s:
				conf, depth, normal, pa, pb = s, hd, hn, na, nb
				idx += 1; scale = 0.5**idx
		return depth, normal, conf


def test_python_process_pixel_matches_oracle(tiny_scene):
	sc, ref, views = tiny_scene
	nR = 3
	prm = O.default_params(schedule=1, propagation=4, nRandomIters=nR, nSubResolutionLevels=0, nEstimationGeometricIters=0, threads=4)
	d, n, c = O.pm_score(views, prm, sc.dmin, sc.dmax)
	rng = np.random.RandomState(21)
	total = close_d = close_c = 0
	for sweep in range(3):
		# colour 0 reads the state before the sweep, colour 1 the state after the first half-sweep
		d1, n1, c1 = O.pm_iterate(views, prm, sc.dmin, sc.dmax, d, n, c, sweep, half=0)
		d2, n2, c2 = O.pm_iterate(views, prm, sc.dmin, sc.dmax, d1, n1, c1, sweep, half=1)
		py = [PyEstimator(views, sc.dmin, sc.dmax, d, n, c, sweep, nR), PyEstimator(views, sc.dmin, sc.dmax, d1, n1, c1, sweep, nR)]
		for _ in range(45):
			x, y = int(rng.randint(3, 157)), int(rng.randint(3, 117))
			pd, pn, pc = py[(x+y) & 1].process(x, y)
			total += 1
			close_d += abs(pd-d2[y, x]) <= 1e-4*max(d2[y, x], 1e-6)+1e-9
			close_c += abs(pc-c2[y, x]) < 2e-4
			if abs(pd-d2[y, x]) <= 1e-5*max(d2[y, x], 1e-6)+1e-9 and d2[y, x] > 0:
				assert np.abs(pn-n2[y, x]).max() < 2e-3
		d, n, c = d2, n2, c2
	# the two transcriptions take the same decisions except where float32/float64 rounding flips an accept test
	print('python-vs-oracle ProcessPixel: depth %d/%d, cost %d/%d' % (close_d, total, close_c, total))
	assert total == 135 and close_d >= 0.95*total and close_c >= 0.95*total
