"""The packed u16x2 form of the SGM scanline step (openmvs_b200/csrc/sgm_step.cuh, the arithmetic of the opt-in
B200MVS_SGM_DPX aggregation kernel) equals the scalar form of the default kernels.  The functions are
__host__ __device__: the check is an nvcc-built host program, no GPU involved."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_step_equals_scalar_step(tmp_path):
	nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
	if not os.path.exists(nvcc):
		pytest.skip("nvcc not available")
	exe = str(tmp_path/"sgm_step")
	subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, os.path.join(ROOT, "tests", "cpp", "sgm_step_main.cu")])
	r = subprocess.run([exe], capture_output=True, text=True)
	assert r.returncode == 0 and "== scalar step on" in r.stdout, r.stdout+r.stderr
