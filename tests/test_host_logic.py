"""Host-side mirror of the reference interface: option snapshot, view marshalling, sharding."""
import ctypes as C

import numpy as np
import pytest

from openmvs_b200 import lib
from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, _make_views


def test_optdense_defaults_match_reference_table():
	# libs/MVS/DepthMap.cpp:69-114
	assert OPTDENSE.nEstimationIters == 3 and OPTDENSE.nEstimationGeometricIters == 2
	assert OPTDENSE.nRandomIters == 6 and OPTDENSE.nSubResolutionLevels == 2
	assert OPTDENSE.fNCCThresholdKeep == pytest.approx(0.9)
	assert OPTDENSE.fRandomDepthRatio == pytest.approx(0.003)
	assert (OPTDENSE.fRandomAngle1Range, OPTDENSE.fRandomAngle2Range) == (16.0, 10.0)
	assert (OPTDENSE.fRandomSmoothDepth, OPTDENSE.fRandomSmoothNormal, OPTDENSE.fRandomSmoothBonus) == (0.02, 13.0, 0.93)
	p = OPTDENSE.snapshot()
	d = lib.Params()
	lib.load().b200mvs_default_params(C.byref(d))
	for name, _ in lib.Params._fields_:
		assert getattr(p, name) == pytest.approx(getattr(d, name)), name


def test_make_views_marshalling_and_errors():
	K = np.array([[100.0, 0, 31.5], [0, 100.0, 23.5], [0, 0, 1]])
	cam = Camera(K, np.eye(3), np.zeros(3))
	img = np.random.RandomState(0).rand(48, 64).astype(np.float32)
	padded = np.zeros((48, 80), np.float32)[:, :64]  # row stride 320 bytes, like a cv::Mat ROI
	arr, keep, dev = _make_views([ViewData(img, cam), ViewData(padded, cam, depthMap=img, cameraDepthMap=cam)])
	assert not dev and arr[0].width == 64 and arr[0].height == 48 and arr[0].stride_bytes == 256
	assert arr[1].stride_bytes == 320 and arr[1].depth and arr[1].dwidth == 64
	assert list(arr[1].Kd) == list(K.ravel())
	with pytest.raises(ValueError):
		_make_views([ViewData(img.astype(np.float64), cam), ViewData(img, cam)])
	with pytest.raises(ValueError):
		_make_views([ViewData(img.T, cam), ViewData(img, cam)])  # non-contiguous rows


def test_depthdata_validity():
	dd = DepthData([], 1.0, 2.0)
	assert not dd.IsValid() and dd.IsEmpty()


def test_shard_assignment_is_a_partition():
	from openmvs_b200.multi_gpu import shard_views
	for n, world in ((12, 1), (12, 8), (200, 8), (7, 4), (3, 8)):
		parts = [shard_views(n, r, world) for r in range(world)]
		flat = sorted(i for p in parts for i in p)
		assert flat == list(range(n))
		assert max(len(p) for p in parts)-min(len(p) for p in parts) <= 1


def test_depthdata_save_load_round_trip(tmp_path):
	K = np.array([[100.0, 0, 31.5], [0, 100.0, 23.5], [0, 0, 1]])
	cam = Camera(K, np.eye(3), np.zeros(3))
	img = np.zeros((48, 64), np.float32)
	rng = np.random.RandomState(2)
	dd = DepthData([ViewData(img, cam), ViewData(img, cam)], 1.0, 9.0, depthMap=rng.rand(48, 64).astype(np.float32),
		normalMap=rng.rand(48, 64, 3).astype(np.float32), confMap=rng.rand(48, 64).astype(np.float32),
		viewsMap=rng.randint(0, 255, (48, 64, 4)).astype(np.uint8))
	p = str(tmp_path/"depth0000.dmap")
	assert dd.Save(p, IDs=[4, 2])
	back = DepthData([ViewData(img, cam), ViewData(img, cam)], 0.0, 0.0)
	assert back.Load(p) and np.array_equal(back.depthMap, dd.depthMap) and np.array_equal(back.viewsMap, dd.viewsMap)
	assert (back.dMin, back.dMax) == (1.0, 9.0)
