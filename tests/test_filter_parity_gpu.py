"""GPU parity of the depth-map post-processing (FilterDepthMap, RemoveSmallSegments, GapInterpolation;
libs/MVS/SceneDensify.cpp:810-1299) against the oracle, through the C-ABI (run with -m gpu).

FilterDepthMap and the depth / confidence of GapInterpolation are bit-exact (the kernels pin the
rounding of every operation); interpolated normals go through sin/cos/atan2/acos and are compared
within 5e-6; RemoveSmallSegments is exact wherever the reference's own result does not depend on
its traversal order (see test_small_segments_*)."""
import numpy as np
import pytest
import torch

from openmvs_b200 import synth
from test_filter_oracle import component_sizes, segment_test_map
from test_pm_parity_gpu import _record

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def eng():
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from oracle import oracle as O
	from openmvs_b200.depth_estimator import DepthMapsData
	dm = DepthMapsData([], 0, nCalibratedImages=12)
	yield dm, O
	dm.pmCUDA.Release()


def _depth_data(sc, maps, i, device=None, conf=True):
	from openmvs_b200.depth_estimator import Camera, DepthData, ViewData
	v = sc.views[i]
	dd = DepthData([ViewData(None, Camera(v.K, v.R, v.C))], sc.dmin, sc.dmax)
	d, c = maps[i]
	if device is None:
		dd.depthMap, dd.confMap = d.copy(), (c.copy() if conf else None)
	else:
		dd.depthMap = torch.from_numpy(d).to(device)
		dd.confMap = torch.from_numpy(c).to(device) if conf else None
	return dd


def _odm(sc, maps, i, conf=True):
	v = sc.views[i]
	return (maps[i][0], maps[i][1] if conf else None, v.K, v.R, v.C)


@pytest.fixture(scope="module")
def scene():
	sc = synth.make_scene(200, 150, 6, step_deg=4.0)
	return sc, synth.make_noisy_dmaps(sc)


@pytest.mark.parametrize("adjust", [True, False])
def test_filter_depth_map_bit_exact(eng, scene, adjust):
	dm, O = eng
	sc, maps = scene
	ref = 2
	nb = sc.neighbors(ref, 5)
	ok, od, oc, proj = O.filter_depth_map(_odm(sc, maps, ref), [_odm(sc, maps, i, adjust) for i in nb], 2, 1, 0.01, adjust, sc.dmin, sc.dmax)
	assert ok
	# device path, with the projected maps
	gd, gc, pd, pc = dm.FilterDepthMap(_depth_data(sc, maps, ref, "cuda"), [_depth_data(sc, maps, i, "cuda", adjust) for i in nb], adjust, projected=True)
	torch.cuda.synchronize()
	assert np.array_equal(pd.cpu().numpy(), proj)
	if adjust:
		for k, i in enumerate(nb):
			assert np.array_equal(pc[k].cpu().numpy(), O.filter_project(_odm(sc, maps, ref), _odm(sc, maps, i))[1])
	assert np.array_equal(gd.cpu().numpy(), od) and np.array_equal(gc.cpu().numpy(), oc)
	# host path
	hd, hc = dm.FilterDepthMap(_depth_data(sc, maps, ref), [_depth_data(sc, maps, i, None, adjust) for i in nb], adjust)
	assert np.array_equal(hd, od) and np.array_equal(hc, oc)
	assert 0.2 < (od > 0).mean() < 0.99
	_record("filter_%s_200x150" % ("adjust" if adjust else "strict"), kept=(od > 0).mean(), valid_in=(maps[ref][0] > 0).mean())


def test_filter_neighbours_of_other_sizes(eng):
	"""neighbour depth-maps need not have the reference's size (each DepthData has its own, SceneDensify.cpp:1079)"""
	dm, O = eng
	sc = synth.make_scene(160, 120, 3, step_deg=5.0, cols=3)
	sc2 = synth.make_scene(120, 90, 3, step_deg=5.0, cols=3)  # the same cameras at another resolution
	m, m2 = synth.make_noisy_dmaps(sc), synth.make_noisy_dmaps(sc2, seed=9)
	ok, od, oc, _ = O.filter_depth_map(_odm(sc, m, 1), [_odm(sc2, m2, 0), _odm(sc, m, 2)], 2, 1, 0.01, True, sc.dmin, sc.dmax)
	assert ok
	gd, gc = dm.FilterDepthMap(_depth_data(sc, m, 1, "cuda"), [_depth_data(sc2, m2, 0, "cuda"), _depth_data(sc, m, 2, "cuda")], True)
	torch.cuda.synchronize()
	assert np.array_equal(gd.cpu().numpy(), od) and np.array_equal(gc.cpu().numpy(), oc)
	assert (od > 0).mean() > 0.2


def test_filter_refuses_too_few_neighbours_and_bad_arguments(eng, scene):
	dm, O = eng
	sc, maps = scene
	assert dm.FilterDepthMap(_depth_data(sc, maps, 2), [_depth_data(sc, maps, 1)], True) is None
	from openmvs_b200.lib import B200MVSError
	with pytest.raises((B200MVSError, ValueError)):
		dm.FilterDepthMap(_depth_data(sc, maps, 2), [_depth_data(sc, maps, i, None, False) for i in (0, 1, 3)], True)  # bAdjust needs confidences
	with pytest.raises(B200MVSError):
		dm.FilterDepthMap(_depth_data(sc, maps, 2), [_depth_data(sc, maps, 1)]*17, True)


def test_small_segments_exact_without_direction_dependent_edges(eng, scene):
	dm, O = eng
	sc, _ = scene
	d = segment_test_map(sc)
	assert O.count_asymmetric_edges(d, f32(0.007)) == 0
	n = sc.views[1].normal_gt.copy(); n[d == 0] = 0
	c = np.where(d > 0, 0.6, 0).astype(f32)
	from openmvs_b200.depth_estimator import DepthData
	od, on, oc = O.remove_small_segments(d, n, c, f32(0.01)*f32(0.7), 100)
	# host path (in place on numpy arrays)
	dd = DepthData([], 0, 0, d.copy(), n.copy(), c.copy())
	assert dm.RemoveSmallSegments(dd)
	assert np.array_equal(dd.depthMap, od) and np.array_equal(dd.normalMap, on) and np.array_equal(dd.confMap, oc)
	# device path, depth only
	t = torch.from_numpy(d).cuda()
	dm.RemoveSmallSegments(DepthData([], 0, 0, t))
	torch.cuda.synchronize()
	assert np.array_equal(t.cpu().numpy(), od)
	assert 0 < (od > 0).sum() < (d > 0).sum()


def test_small_segments_with_direction_dependent_edges(eng, scene):
	"""Noisy maps with edges that pass IsDepthSimilar in one direction only: the reference's segments depend on its seed order
	(column-major) there.  The engine labels the two-way components on the GPU and replays the reference's loop on the condensed
	graph of one-way edges: bit-identical to the breadth-first oracle, also on a map built to contain many such edges."""
	dm, O = eng
	sc, maps = scene
	th = f32(f32(0.01)*f32(0.7))
	from openmvs_b200.depth_estimator import DepthData, OPTDENSE
	rng = np.random.RandomState(4)
	# (a) an estimator-like noisy map; (b) a staircase whose steps sit right at the threshold: hundreds of one-way edges
	stairs = np.full((150, 200), 5.0, f32)
	for k in range(1, 40):
		stairs[:, 5*k:] *= f32(1.0+float(th)*(0.9965+0.007*rng.rand()))
	stairs[rng.rand(150, 200) < 0.08] = 0
	blobs = maps[2][0].copy()
	for name, d in (("noisy", maps[2][0]), ("stairs", stairs), ("noisy_sparse", np.where(rng.rand(*blobs.shape) < 0.25, 0, blobs).astype(f32))):
		asym = O.count_asymmetric_edges(d, th)
		for speckle in (20, 100, 1000):
			OPTDENSE.nSpeckleSize = speckle
			try:
				dd = DepthData([], 0, 0, d.copy())
				dm.RemoveSmallSegments(dd)
				t = torch.from_numpy(d.copy()).cuda()
				dm.RemoveSmallSegments(DepthData([], 0, 0, t))
			finally:
				OPTDENSE.nSpeckleSize = 100
			od, _, _ = O.remove_small_segments(d, None, None, th, speckle)
			assert np.array_equal(dd.depthMap, od), (name, speckle, int((dd.depthMap != od).sum()))
			assert np.array_equal(t.cpu().numpy(), od)
			_record("segments_%s_speckle%d" % (name, speckle), asym_edges=asym, removed=((d > 0) & (od == 0)).mean(), differ=float((dd.depthMap != od).mean()))
	assert O.count_asymmetric_edges(stairs, th) > 50


def _gap_inputs(sc, seed=5):
	rng = np.random.RandomState(seed)
	v = sc.views[1]
	d = v.depth_gt.copy(); n = v.normal_gt.copy(); c = rng.uniform(0.1, 1, d.shape).astype(f32)
	for _ in range(200):
		y, x, l = rng.randint(0, d.shape[0]), rng.randint(0, d.shape[1]), rng.randint(1, 11)
		if rng.rand() < 0.5:
			d[y, x:x+l] = 0
		else:
			d[y:y+l, x] = 0
	d[:, 0] = 0; d[0, 5:9] = 0; d[-1, :] = 0
	n[d == 0] = 0; c[d == 0] = 0
	return d, n, c


def test_gap_interpolation_parity(eng, scene):
	dm, O = eng
	sc, _ = scene
	from openmvs_b200.depth_estimator import DepthData
	d, n, c = _gap_inputs(sc)
	od, on, oc = O.gap_interpolation(d, n, c, f32(0.01)*f32(2.5), 7)
	dd = DepthData([], 0, 0, d.copy(), n.copy(), c.copy())
	assert dm.GapInterpolation(dd)
	assert np.array_equal(dd.depthMap, od) and np.array_equal(dd.confMap, oc)
	assert np.abs(dd.normalMap-on).max() < 5e-6
	# device path, without normals / confidences
	t = torch.from_numpy(d).cuda()
	dm.GapInterpolation(DepthData([], 0, 0, t))
	torch.cuda.synchronize()
	assert np.array_equal(t.cpu().numpy(), O.gap_interpolation(d, None, None, f32(0.01)*f32(2.5), 7)[0])
	filled = (d == 0) & (od > 0)
	assert filled.sum() > 100


def test_post_processing_at_1080p(eng):
	"""BASELINE configs[1] size: 8 neighbour maps of 1920x1080 projected into one reference view.  The oracle still
	finishes in seconds here, so the full-size run is compared exactly too; timings are recorded."""
	dm, O = eng
	sc = synth.make_scene(1920, 1080, 9, step_deg=4.0, device=torch.device("cuda"))
	maps = synth.make_noisy_dmaps(sc)
	ref = 4
	nb = sc.neighbors(ref, 8)
	R = _depth_data(sc, maps, ref, "cuda"); N = [_depth_data(sc, maps, i, "cuda") for i in nb]
	gd, gc = dm.FilterDepthMap(R, N, True)
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(5):
		gd, gc = dm.FilterDepthMap(R, N, True)
	e1.record(); torch.cuda.synchronize()
	ms_filter = e0.elapsed_time(e1)/5
	ok, od, oc, _ = O.filter_depth_map(_odm(sc, maps, ref), [_odm(sc, maps, i) for i in nb], 2, 1, 0.01, True, sc.dmin, sc.dmax)
	assert ok and np.array_equal(gd.cpu().numpy(), od) and np.array_equal(gc.cpu().numpy(), oc)
	gt = sc.views[ref].depth_gt
	din = maps[ref][0]
	inl = (din > 0) & (np.abs(din-gt)/gt < 0.006); out = (din > 0) & (np.abs(din-gt)/gt > 0.05)
	assert (od[inl] > 0).mean() > 0.97 and (od[out] > 0).mean() < 0.01
	k = od > 0
	# (the nearest-of-4 z-buffer splat biases projected depths towards the camera, so the average is not closer to the
	# truth than the input; it stays in the same range)
	assert np.abs(od[k]-gt[k]).mean() < 1.3*np.abs(din[k]-gt[k]).mean()
	# speckles and gaps on the filtered map, on the device
	from openmvs_b200.depth_estimator import DepthData
	n = torch.from_numpy(sc.views[ref].normal_gt).cuda()*(gd > 0)[..., None]
	dd = DepthData([], 0, 0, gd.clone(), n.contiguous(), gc.clone())
	warm = DepthData([], 0, 0, gd.clone(), n.contiguous().clone(), gc.clone())  # first launches load the kernels: not timed
	dm.RemoveSmallSegments(warm); dm.GapInterpolation(warm); torch.cuda.synchronize()
	e0.record(); dm.RemoveSmallSegments(dd); e1.record(); torch.cuda.synchronize()
	ms_seg = e0.elapsed_time(e1)
	weak = component_sizes(od, f32(f32(0.01)*f32(0.7)), False)
	assert np.array_equal(dd.depthMap.cpu().numpy(), np.where(weak < 100, 0, od).astype(f32))
	seg = dd.depthMap.cpu().numpy(); segn = dd.normalMap.cpu().numpy(); segc = dd.confMap.cpu().numpy()
	e0.record(); dm.GapInterpolation(dd); e1.record(); torch.cuda.synchronize()
	ms_gap = e0.elapsed_time(e1)
	qd, qn, qc = O.gap_interpolation(seg, segn, segc, f32(0.01)*f32(2.5), 7)
	assert np.array_equal(dd.depthMap.cpu().numpy(), qd) and np.array_equal(dd.confMap.cpu().numpy(), qc)
	assert np.abs(dd.normalMap.cpu().numpy()-qn).max() < 5e-6
	_record("post_processing_1080p_n8", ms_filter=ms_filter, ms_segments=ms_seg, ms_gaps=ms_gap,
		kept=(od > 0).mean(), removed_speckle=((od > 0) & (seg == 0)).mean(), filled=((seg == 0) & (qd > 0)).mean())
