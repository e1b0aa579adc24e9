"""The C-ABI library builds, loads and exports every symbol include/b200mvs.h declares
(no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
	src = open(os.path.join(ROOT, "include", "b200mvs.h")).read()
	src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
	return sorted(set(re.findall(r"\b(b200mvs_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
	from openmvs_b200 import build, lib
	path = build.build_extension()
	assert os.path.exists(path)
	dll = C.CDLL(path)
	names = _declared()
	assert len(names) >= 21
	for n in names:
		assert hasattr(dll, n), "missing export %s" % n
	# the python binding lists exactly the declared symbols
	assert sorted(lib.SYMBOLS) == names


def test_struct_layouts_match_header_sizes():
	from openmvs_b200 import lib
	# b200mvs_view: pointer, 3 ints (+pad), 21 doubles, pointer, 3 ints (+pad), 21 doubles
	assert C.sizeof(lib.View) == 8+16+21*8+8+16+21*8-8+8 or C.sizeof(lib.View) % 8 == 0
	assert C.sizeof(lib.Params) == 19*4
	assert C.sizeof(lib.Stats) == 8+8+8+8+4+4+8+4+4


def test_struct_layouts_match_a_c_compiler(tmp_path):
	"""sizeof / offsetof of every struct of include/b200mvs.h as gcc sees them == the ctypes mirrors in openmvs_b200/lib.py"""
	import subprocess
	from openmvs_b200 import lib
	structs = {"b200mvs_view": lib.View, "b200mvs_params": lib.Params, "b200mvs_stats": lib.Stats, "b200mvs_job": lib.Job,
		"b200mvs_sgm_params": lib.SgmParams, "b200mvs_dmap": lib.DMap, "b200mvs_filter_params": lib.FilterParams,
		"b200mvs_debug": lib.Debug, "b200mvs_sgm_pixel": lib.SgmPixel, "b200mvs_fuse_view": lib.FuseView, "b200mvs_fuse_params": lib.FuseParams}
	lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200mvs.h"', 'int main(void) {']
	for cname, ct in structs.items():
		lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
		for fname, _ in ct._fields_:
			lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
	lines += ['printf("b200mvs_sgm_pixel %zu\\n", sizeof(b200mvs_sgm_pixel));', 'return 0; }']
	src = tmp_path/"layout.c"
	src.write_text("\n".join(lines))
	exe = str(tmp_path/"layout")
	subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)])
	got = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
	for cname, ct in structs.items():
		assert int(got[cname]) == C.sizeof(ct), cname
		for fname, _ in ct._fields_:
			assert int(got["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)
	assert int(got["b200mvs_sgm_pixel"]) == 16  # SemiGlobalMatcher::PixelData


def test_create_without_gpu_fails_loudly():
	import torch
	if torch.cuda.is_available():
		pytest.skip("GPU present")
	from openmvs_b200 import lib
	from openmvs_b200.depth_estimator import PatchMatchB200
	with pytest.raises(lib.B200MVSError):
		PatchMatchB200(0)
	dll = lib.load()
	assert dll.b200mvs_device_count() == 0
	ctx = C.c_void_p()
	assert dll.b200mvs_create(0, C.byref(ctx)) == 3  # B200MVS_ERR_NOGPU: never a CPU fallback
	assert not ctx


def test_sources_do_not_reference_the_oracle():
	"""The product package must not import, link or execute anything under oracle/."""
	pkg = os.path.join(ROOT, "openmvs_b200")
	for dp, _, files in os.walk(pkg):
		for f in files:
			if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
				txt = open(os.path.join(dp, f)).read()
				assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_null_context_and_bad_arguments_return_status_codes():
	"""Error behaviour at the boundary: status codes, never a crash or exit()."""
	from openmvs_b200 import lib
	dll = lib.load()
	assert dll.b200mvs_set_params(None, None) == 1
	assert dll.b200mvs_set_debug(None, None) == 1 and dll.b200mvs_set_ignore_mask(None, None, 0, 0, 0, 0) == 1
	assert dll.b200mvs_get_schedule(None, 0, None, None) == 1
	assert dll.b200mvs_abi_version() == lib.ABI_VERSION
	# the library reports the struct sizes it was compiled with (checked against the bindings by lib.load())
	assert dll.b200mvs_sizeof(1) == C.sizeof(lib.Params) and dll.b200mvs_sizeof(99) == 0
	assert dll.b200mvs_destroy(None) == 1
	assert dll.b200mvs_estimate(None, None, 0, C.c_float(1), C.c_float(2), -1, None, None, None, None, None) == 1
	assert dll.b200mvs_sync(None, None) == 1
	assert dll.b200mvs_estimate_batch(None, 0, None, 0) == 1
	assert dll.b200mvs_sgm_match(None, None, None, None, 0, 0, None, C.c_uint64(0), None, None, None, None) == 1
	assert dll.b200mvs_last_error(None) == b"null context"
	assert dll.b200mvs_filter_depth_map(None, None, None, 0, None, C.c_float(0), C.c_float(1), None, None, None, None) == 1
	assert dll.b200mvs_filter_depth_map_device(None, None, None, 0, None, C.c_float(0), C.c_float(1), None, None, None, None, None, None) == 1
	assert dll.b200mvs_remove_small_segments(None, None, None, None, 0, 0, C.c_float(0.01), 100, None) == 1
	assert dll.b200mvs_gap_interpolation_device(None, None, None, None, 0, 0, C.c_float(0.01), 7, None) == 1
