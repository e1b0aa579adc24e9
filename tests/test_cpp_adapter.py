"""The C++ adapter (include/PatchMatchB200.hpp) compiles against a stand-in DepthData, links with
the C-ABI library, fails loudly without a GPU, and on a GPU reproduces the Python host path."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
	from openmvs_b200 import build
	lib = build.build_extension()
	exe = os.path.join(tmp, "adapter_main")
	subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "adapter_main.cpp"),
		lib, "-Wl,-rpath," + os.path.dirname(lib)])
	return exe


def _dump_scene(path, views, dmin, dmax):
	h, w = views[0].image.shape
	with open(path, "wb") as f:
		f.write(struct.pack("iii", len(views), w, h))
		f.write(struct.pack("ff", dmin, dmax))
		for v in views:
			f.write(np.asarray(v.K, np.float64).tobytes()); f.write(np.asarray(v.R, np.float64).tobytes()); f.write(np.asarray(v.C, np.float64).tobytes())
			f.write(np.ascontiguousarray(v.image, np.float32).tobytes())


def test_adapter_compiles_links_and_fails_loudly_without_gpu(tmp_path, tiny_scene):
	import torch
	exe = _build(str(tmp_path))
	sc, ref, views = tiny_scene
	scene = str(tmp_path/"scene.bin")
	_dump_scene(scene, views, sc.dmin, sc.dmax)
	if torch.cuda.is_available():
		pytest.skip("GPU present: covered by the gpu test")
	r = subprocess.run([exe, scene, str(tmp_path/"out.bin")], capture_output=True, text=True)
	assert r.returncode == 3 and "status 3" in r.stdout  # B200MVS_ERR_NOGPU, no CPU fallback


@pytest.mark.gpu
def test_adapter_matches_python_host_path(tmp_path, tiny_scene):
	import torch
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200
	exe = _build(str(tmp_path))
	sc, ref, views = tiny_scene
	scene, out = str(tmp_path/"scene.bin"), str(tmp_path/"out.bin")
	_dump_scene(scene, views, sc.dmin, sc.dmax)
	r = subprocess.run([exe, scene, out, "2"], capture_output=True, text=True)
	assert r.returncode == 0, r.stdout+r.stderr
	h, w = views[0].image.shape
	raw = np.fromfile(out, np.uint8)
	depth = raw[:h*w*4].view(np.float32).reshape(h, w)
	normal = raw[h*w*4:h*w*16].view(np.float32).reshape(h, w, 3)
	conf = raw[h*w*16:h*w*20].view(np.float32).reshape(h, w)
	saved = (OPTDENSE.nSubResolutionLevels, OPTDENSE.nEstimationGeometricIters, OPTDENSE.nEstimationIters)
	try:
		OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 0; OPTDENSE.nEstimationIters = 2
		pm = PatchMatchB200(0)
		dd = DepthData([ViewData(np.ascontiguousarray(v.image), Camera(v.K, v.R, v.C)) for v in views], sc.dmin, sc.dmax)
		pm.EstimateDepthMap(dd)
		pm.Release()
	finally:
		OPTDENSE.nSubResolutionLevels, OPTDENSE.nEstimationGeometricIters, OPTDENSE.nEstimationIters = saved
	assert np.array_equal(depth, dd.depthMap) and np.array_equal(normal, dd.normalMap) and np.array_equal(conf, dd.confMap)
	# the reference's own call-site signature, pmCUDA->EstimateDepthMap(depthData) with the OPTDENSE globals: same maps
	out2 = str(tmp_path/"out2.bin")
	r = subprocess.run([exe, scene, out2, "2", "seam"], capture_output=True, text=True)
	assert r.returncode == 0, r.stdout+r.stderr
	assert np.array_equal(np.fromfile(out2, np.uint8), raw)


@pytest.mark.gpu
def test_adapter_post_processing_matches_python_host_path(tmp_path, tiny_scene):
	"""RemoveSmallSegments, GapInterpolation and FilterDepthMap through the C++ adapter == the Python host path"""
	import torch
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, DepthMapsData
	exe = _build(str(tmp_path))
	sc, ref, views = tiny_scene
	scene, out = str(tmp_path/"scene.bin"), str(tmp_path/"out.bin")
	_dump_scene(scene, views, sc.dmin, sc.dmax)
	r = subprocess.run([exe, scene, out, "2", "post"], capture_output=True, text=True)
	assert r.returncode == 0, r.stdout+r.stderr
	h, w = views[0].image.shape
	raw = np.fromfile(out, np.uint8)
	depth = raw[:h*w*4].view(np.float32).reshape(h, w)
	normal = raw[h*w*4:h*w*16].view(np.float32).reshape(h, w, 3)
	conf = raw[h*w*16:h*w*20].view(np.float32).reshape(h, w)
	saved = (OPTDENSE.nSubResolutionLevels, OPTDENSE.nEstimationGeometricIters, OPTDENSE.nEstimationIters)
	try:
		OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 0; OPTDENSE.nEstimationIters = 2
		dd = DepthData([ViewData(np.ascontiguousarray(v.image), Camera(v.K, v.R, v.C)) for v in views], sc.dmin, sc.dmax)
		dm = DepthMapsData([dd], 0, nCalibratedImages=3)
		dm.EstimateDepthMap(0)
		raw_valid = (dd.depthMap > 0).sum()
		dm.RemoveSmallSegments(dd)
		dm.GapInterpolation(dd)
		nd, nc = dm.FilterDepthMap(dd, [dd, dd], True)
		dm.pmCUDA.Release()
	finally:
		OPTDENSE.nSubResolutionLevels, OPTDENSE.nEstimationGeometricIters, OPTDENSE.nEstimationIters = saved
	assert np.array_equal(depth, nd) and np.array_equal(conf, nc) and np.array_equal(normal, dd.normalMap)
	assert 0.5*raw_valid < (nd > 0).sum()
