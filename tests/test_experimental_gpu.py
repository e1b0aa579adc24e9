"""Checks of the prepared, not yet default kernel variants: skipped unless
B200MVS_TEST_EXPERIMENTAL=1, so that the round-end `-m gpu` gate only runs what has been measured.
Run with:  B200MVS_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu -q"""
import os

import numpy as np
import pytest
import torch

from openmvs_b200 import synth

pytestmark = [pytest.mark.gpu,
	pytest.mark.skipif(os.environ.get("B200MVS_TEST_EXPERIMENTAL") != "1", reason="experimental variants: set B200MVS_TEST_EXPERIMENTAL=1")]


def _dev(a):
	return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("env", [dict(B200MVS_SGM_DPX="1"), dict(B200MVS_SGM_CONCURRENT="1"), dict(B200MVS_SGM_DPX="1", B200MVS_SGM_CONCURRENT="1")])
@pytest.mark.parametrize("num", [128, 256, 48])
def test_sgm_aggregation_variants_bit_exact(env, num):
	"""packed u16x2 step (num = 128 only; other sizes fall back) and concurrent directions against the oracle"""
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from oracle import oracle as O
	from openmvs_b200.depth_estimator import SemiGlobalMatcher
	w, h = 150, 90
	rng = np.random.RandomState(num)
	lg, lc, rg, d = synth.make_stereo_pair(w, h)
	px, n = synth.sgm_pixel_map(w, h, -7, -7+num, rng.rand(h-6, w-6) < 0.05)
	costs = rng.randint(0, 256, n).astype(np.uint8)
	c, a, disp, cost = O.sgm_match(lg, lc, rg, px, n, costs=costs)
	saved = {k: os.environ.get(k) for k in env}
	os.environ.update(env)
	try:
		m = SemiGlobalMatcher()
		accums = torch.zeros(n, dtype=torch.int16, device="cuda")
		gd, gc = m.MatchDevice(_dev(lg), _dev(lc), _dev(rg), torch.from_numpy(px.view(np.uint8).reshape(-1, 16).copy()).cuda(), n,
			stages=6, costs=_dev(costs), accums=accums)
		m.Release()
	finally:
		for k, v in saved.items():
			if v is None: os.environ.pop(k, None)
			else: os.environ[k] = v
	assert np.array_equal(accums.cpu().numpy().view(np.uint16), a)
	assert np.array_equal(gd.cpu().numpy(), disp) and np.array_equal(gc.cpu().numpy().view(np.uint16), cost)


def test_packed_taps_in_geometric_pass_bit_identical(small_scene):
	"""B200MVS_PACK=2 (packed taps also in the GEOM instantiations) == B200MVS_PACK=0 on a geometric pass"""
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200
	sc, ref, views = small_scene
	rng = np.random.RandomState(7)
	init_d = sc.views[ref].depth_gt*(1+0.002*rng.randn(*sc.views[ref].depth_gt.shape).astype(np.float32))
	init_d[rng.rand(*init_d.shape) < 0.1] = 0
	outs = []
	saved = {k: getattr(OPTDENSE, k) for k in ("nSubResolutionLevels", "nEstimationGeometricIters", "nEstimationIters")}
	old = os.environ.get("B200MVS_PACK")
	try:
		OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 2; OPTDENSE.nEstimationIters = 3
		for pack in ("0", "2"):
			os.environ["B200MVS_PACK"] = pack
			pm = PatchMatchB200(0)
			imgs = [ViewData(np.ascontiguousarray(views[0].image), Camera(views[0].K, views[0].R, views[0].C))]
			for v in views[1:]:
				imgs.append(ViewData(np.ascontiguousarray(v.image), Camera(v.K, v.R, v.C), depthMap=v.depth_gt.copy(), cameraDepthMap=Camera(v.K, v.R, v.C)))
			dd = DepthData(imgs, sc.dmin, sc.dmax, depthMap=init_d.copy(), normalMap=sc.views[ref].normal_gt.copy())
			pm.Init(True)
			pm.EstimateDepthMap(dd, 0)
			pm.Release()
			outs.append((dd.depthMap.copy(), dd.normalMap.copy(), dd.confMap.copy()))
	finally:
		for k, v in saved.items(): setattr(OPTDENSE, k, v)
		if old is None: os.environ.pop("B200MVS_PACK", None)
		else: os.environ["B200MVS_PACK"] = old
	assert all(np.array_equal(a, b) for a, b in zip(*outs))
	assert (outs[0][0] > 0).mean() > 0.8


def test_to_gray_device_bit_exact():
	"""b200mvs_to_gray_device against the oracle (3 and 4 channels, BGR and RGB, odd size)"""
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from oracle import oracle as O
	from openmvs_b200.depth_estimator import PatchMatchB200
	rng = np.random.RandomState(3)
	pm = PatchMatchB200(0)
	for ch in (3, 4):
		img = rng.randint(0, 256, (241, 323, ch)).astype(np.uint8)
		for bgr in (True, False):
			g = pm.ToGray(_dev(img), bgr)
			torch.cuda.synchronize()
			assert np.array_equal(g.cpu().numpy(), O.to_gray(img, bgr))
	pm.Release()
