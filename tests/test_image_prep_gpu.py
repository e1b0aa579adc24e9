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
