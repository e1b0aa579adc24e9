"""Image preparation on the device: toGray against the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
	return torch.from_numpy(np.ascontiguousarray(a)).cuda()


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


def test_scale_image_device_matches_cv2():
	"""b200mvs_scale_image_device (ViewData::ScaleImage, DepthMap.h:193-203) against cv2.resize itself: INTER_AREA for
	scale < 1 (integer and fractional ratios), INTER_CUBIC for scale > 1, nothing for |scale - 1| < 0.15."""
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	import cv2
	from openmvs_b200.depth_estimator import PatchMatchB200
	rng = np.random.RandomState(5)
	img = cv2.GaussianBlur(rng.rand(241, 323).astype(np.float32), (0, 0), 1.2)
	pm = PatchMatchB200(0)
	for scale in (0.5, 0.62, 0.8, 0.84, 1.16, 1.3, 1.7, 2.0):
		got = pm.ScaleImage(_dev(img), scale)
		want = cv2.resize(img, None, fx=float(np.float32(scale)), fy=float(np.float32(scale)), interpolation=cv2.INTER_CUBIC if scale > 1 else cv2.INTER_AREA)
		assert got is not None and tuple(got.shape) == want.shape, (scale, got.shape, want.shape)
		err = np.abs(got.cpu().numpy()-want)
		assert err.max() < 3e-5, (scale, float(err.max()))
	# NeedScaleImage is |scale - 1.f| >= 0.15f in float: 0.85f and 1.15f fall just below it, like in the reference
	for scale in (0.85, 0.9, 1.0, 1.1, 1.15):
		assert pm.ScaleImage(_dev(img), scale) is None
	# a row-strided source (cv::Mat ROI)
	wide = torch.zeros((241, 330), dtype=torch.float32, device="cuda"); wide[:, :323] = _dev(img)
	got = pm.ScaleImage(wide[:, :323], 0.62)
	assert np.abs(got.cpu().numpy()-cv2.resize(img, None, fx=float(np.float32(0.62)), fy=float(np.float32(0.62)), interpolation=cv2.INTER_AREA)).max() < 3e-5
	pm.Release()


def test_estimate_from_8bit_colour_images_equals_float_gray_path():
	"""b200mvs_view.image8: the estimate from 8-bit BGR images (toGray on the device inside the call, 3 B per pixel uploaded) is
	bit-identical to the estimate from the float gray images the oracle's toGray produces; host and device paths."""
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from oracle import oracle as O
	from openmvs_b200 import synth
	from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200
	sc = synth.make_scene(200, 150, 3, step_deg=5.0, cols=3)
	rng = np.random.RandomState(1)
	views = [sc.views[1], sc.views[0], sc.views[2]]
	bgr = []
	for v in views:
		g = np.rint(v.image*255).astype(np.int32)
		c = np.stack([np.clip(g+rng.randint(-6, 7, g.shape), 0, 255) for _ in range(3)], -1).astype(np.uint8)
		bgr.append(np.ascontiguousarray(c))
	gray = [O.to_gray(c, True) for c in bgr]
	saved = (OPTDENSE.nSubResolutionLevels, OPTDENSE.nEstimationGeometricIters, OPTDENSE.nEstimationIters)
	OPTDENSE.nSubResolutionLevels = 1; OPTDENSE.nEstimationGeometricIters = 0; OPTDENSE.nEstimationIters = 2
	pm = PatchMatchB200(0)
	try:
		cams = [Camera(v.K, v.R, v.C) for v in views]
		ref = DepthData([ViewData(g, c) for g, c in zip(gray, cams)], sc.dmin, sc.dmax)
		pm.EstimateDepthMap(ref)
		h8 = DepthData([ViewData(b, c) for b, c in zip(bgr, cams)], sc.dmin, sc.dmax)
		pm.EstimateDepthMap(h8)
		assert pm.stats.bytes_h2d == sum(b.nbytes for b in bgr)+gray[0].size*16
		d8 = DepthData([ViewData(_dev(b), c) for b, c in zip(bgr, cams)], sc.dmin, sc.dmax)
		pm.EstimateDepthMap(d8)
		for other in (h8.depthMap, d8.depthMap.cpu().numpy()):
			assert np.array_equal(ref.depthMap, other)
		assert np.array_equal(ref.confMap, h8.confMap) and (ref.depthMap > 0).mean() > 0.5
	finally:
		OPTDENSE.nSubResolutionLevels, OPTDENSE.nEstimationGeometricIters, OPTDENSE.nEstimationIters = saved
		pm.Release()
