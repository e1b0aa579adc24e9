#!/usr/bin/env python
"""bench.py — Mpix/s of depth+normal estimation (BASELINE.json metric).

Workload (config.workload "C2"): BASELINE.json configs[1] — a 12-image synthetic scene at
1920x1080, every image used once as reference view with its 9 nearest neighbours,
PatchMatch 6 iterations, single scale, no geometric pass.
One step = the 12 DepthMapsData::EstimateDepthMap calls of that scene on one GPU.
With N GPUs every rank estimates its own 12-view shard (weak scaling; reference views are
independent, no data-path collective) and the maps are gathered on rank 0 over NCCL inside
the timed region.

  value      device-resident throughput: images already in HBM, reference views alternating
             between two contexts on two CUDA streams (joined back into the timing stream), CUDA
             events, barrier + synchronize on both sides, max over ranks.
  e2e        the same work through the reference-facing call with HOST buffers (pinned):
             b200mvs_estimate_async + b200mvs_sync on two contexts used alternately, so that the
             copies of one reference view overlap the kernels of the other; per reference view
             the H2D copy of its 10 images + initial maps and the D2H read of
             depth/normal/conf/views are inside the timed region.
  roofline   dominant kernel (pm_sweep_kernel, one red-black half-sweep), timed live with CUDA
             events; algorithmic bytes per launch = (20 B plane+cost read for every pixel +
             20 B written for the active half + 4(N+1) B of images) per pixel.
  cpu_baseline / --impl reference
             the reference algorithm (oracle, zig-zag schedule, all host threads) on a bounded
             sample: a full-width band of one reference view, same N and iteration count.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

W, H, N_VIEWS, N_NEIGH, ITERS = 1920, 1080, 12, 9, 6
METRIC = "Mpix/sec depth+normal (1920x1080, 9 neighbours)"


def parse():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=3)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
	ap.add_argument("--cpu-band", type=int, default=0, help="rows of the CPU sample band (0: auto)")
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--small", action="store_true", help="developer mode: 640x360 scene")
	return ap.parse_args()


class ClockSampler:
	"""nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
	Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

	def __init__(self, index: int):
		self.index = index
		self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
		self.p = None

	def start(self):
		try:
			self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu="+self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
				stdout=self.f, stderr=subprocess.DEVNULL)
		except Exception:
			self.p = None

	def stop(self):
		if self.p is not None:
			self.p.terminate()
			try:
				self.p.wait(5)
			except Exception:
				self.p.kill()
		self.f.flush(); self.f.seek(0)
		sm, mx, reasons = [], [], set()
		for line in self.f.read().splitlines():
			c = [x.strip() for x in line.split(",")]
			if len(c) < 9:
				continue
			try:
				sm.append(float(c[1])); mx.append(float(c[2]))
			except ValueError:
				continue
			for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
				if val.lower().startswith("active"):
					reasons.add(name)
		try:
			os.unlink(self.f.name)
		except OSError:
			pass
		if not sm:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
		busy = [s for s in sm if s > 0.5*max(sm)] or sm
		return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(device, w, h):
	from openmvs_b200 import synth
	return synth.make_scene(w, h, N_VIEWS, step_deg=4.0, device=device)


def cpu_sample(scene, w, h, band_rows, threads):
	"""The reference algorithm on a bounded sample: reference view 5, rows [y0, y0+band) as a
	cropped pinhole view (principal point shifted), full neighbour images, N and iterations as
	in the GPU workload.  Returns (Mpix/s, seconds, description)."""
	from oracle import oracle as O
	from openmvs_b200 import synth
	ref = 5
	nb = scene.neighbors(ref, N_NEIGH)
	v = scene.views[ref]
	y0 = (h-band_rows)//2
	K = v.K.copy(); K[1, 2] -= y0
	crop = synth.View(np.ascontiguousarray(v.image[y0:y0+band_rows]), K, v.R, v.C, v.depth_gt[y0:y0+band_rows], v.normal_gt[y0:y0+band_rows])
	views = [crop]+[scene.views[i] for i in nb]
	prm = O.default_params(schedule=0, nEstimationIters=ITERS, nSubResolutionLevels=0, nEstimationGeometricIters=0, threads=threads)
	t = time.perf_counter()
	d, n, c = O.pm_estimate(views, prm, scene.dmin, scene.dmax)
	dt = time.perf_counter()-t
	mpix = w*band_rows/1e6/dt
	desc = "oracle ZZ schedule (reference algorithm restated), ref view %d rows %d..%d (%dx%d band), %d neighbours, %d iters, %d threads" % (
		ref, y0, y0+band_rows, w, band_rows, N_NEIGH, ITERS, threads)
	return mpix, dt, desc, float((d > 0).mean())


def main():
	args = parse()
	w, h = (640, 360) if args.small else (W, H)
	rank = int(os.environ.get("RANK", "0"))
	world = int(os.environ.get("WORLD_SIZE", "1"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	try:
		threads = len(os.sched_getaffinity(0)) or 1
	except Exception:
		threads = os.cpu_count() or 1

	if args.impl == "reference":
		# the reference's own CPU implementation of the path: the oracle port (the reference cannot
		# be compiled in this image).  Rank 0 only.
		if rank != 0:
			return 0
		scene = build_scene(None if args.small else _maybe_cuda(), w, h)
		band = args.cpu_band or max(24, min(h, int(round(h*0.25*threads/8.0))))
		for _ in range(args.warmup and 1):
			cpu_sample(scene, w, h, 16, threads)
		t_total, px_total, valid = 0.0, 0.0, 0.0
		for _ in range(args.steps):
			mp, dt, desc, valid = cpu_sample(scene, w, h, band, threads)
			t_total += dt; px_total += w*band/1e6
		val = px_total/t_total
		out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
			"warmup": args.warmup, "ms_per_step": 1e3*t_total/args.steps, "higher_is_better": True, "scaling": "weak",
			"vs_baseline": None, "dtype": "f32", "data": "synthetic",
			"config": {"workload": "C2: 12x1920x1080, 9 neighbours, PatchMatch 6 iters, single scale (bounded CPU sample per step)", "sample": desc},
			"cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": desc},
			"e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
		print(json.dumps(out))
		return 0

	import torch
	import torch.distributed as dist
	from openmvs_b200 import multi_gpu
	from openmvs_b200.depth_estimator import OPTDENSE, Camera, ViewData, DepthData, PatchMatchB200

	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
	torch.cuda.set_device(local_rank)
	dev = torch.device("cuda", local_rank)
	if world > 1:
		dist.init_process_group("nccl", device_id=dev)

	OPTDENSE.nSubResolutionLevels = 0; OPTDENSE.nEstimationGeometricIters = 0
	OPTDENSE.nEstimationIters = ITERS; OPTDENSE.nRandomIters = 6; OPTDENSE.nSweepsPerIter = 0; OPTDENSE.nPropagation = 4
	N_SWEEPS, N_REFINE = OPTDENSE.schedule()
	scene = build_scene(dev, w, h)
	nbrs = [scene.neighbors(r, N_NEIGH) for r in range(N_VIEWS)]
	pm = PatchMatchB200(local_rank)
	cams = [Camera(v.K, v.R, v.C) for v in scene.views]
	# device-resident copies and pinned host copies of the images
	d_imgs = [torch.from_numpy(v.image).to(dev) for v in scene.views]
	h_imgs = []
	for v in scene.views:
		t = torch.empty((h, w), dtype=torch.float32, pin_memory=True)
		t.copy_(torch.from_numpy(v.image))
		h_imgs.append(t.numpy())
	# per-reference-view in/out maps
	d_maps = [dict(depth=torch.zeros(h, w, device=dev), normal=torch.zeros(h, w, 3, device=dev), conf=torch.zeros(h, w, device=dev),
		views=torch.zeros(h, w, 4, dtype=torch.uint8, device=dev)) for _ in range(N_VIEWS)]
	pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
	h_maps = [dict(depth=pin((h, w), torch.float32), normal=pin((h, w, 3), torch.float32), conf=pin((h, w), torch.float32),
		views=pin((h, w, 4), torch.uint8)) for _ in range(N_VIEWS)]
	launches = [0]

	pm2 = PatchMatchB200(local_rank)
	pms = [pm, pm2]
	streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

	def step_resident():
		# reference views alternate between two contexts on two CUDA streams, so that the ramp-down of one
		# view's kernels overlaps the start of the next view's (same two-context scheme as the e2e path)
		main = torch.cuda.current_stream(dev)
		for st in streams:
			st.wait_stream(main)
		for r in range(N_VIEWS):
			k = r & 1
			with torch.cuda.stream(streams[k]):
				m = d_maps[r]
				m["depth"].zero_(); m["normal"].zero_()  # depth 0 => random initialisation on device
				dd = DepthData([ViewData(d_imgs[r], cams[r])]+[ViewData(d_imgs[i], cams[i]) for i in nbrs[r]], scene.dmin, scene.dmax,
					depthMap=m["depth"], normalMap=m["normal"], confMap=m["conf"], viewsMap=m["views"])
				pms[k].EstimateDepthMap(dd, sync=False)
			launches[0] += 1+1+N_SWEEPS*2+1
		for st in streams:
			main.wait_stream(st)
		if world > 1:
			# final gather of depth+normal+conf of this rank's views on rank 0 (NCCL)
			packed = {v: torch.cat([d_maps[k]["depth"][..., None], d_maps[k]["normal"], d_maps[k]["conf"][..., None]], -1)
				for k, v in enumerate(multi_gpu.shard_views(N_VIEWS*world, rank, world))}
			multi_gpu.gather_maps(packed, N_VIEWS*world, dst=0)

	def step_e2e():
		# two contexts used alternately (b200mvs_estimate_async / b200mvs_sync): the H2D/D2H copies of one
		# reference view overlap the kernels of the other, like the reference's two worker threads around the seam
		h2d = d2h = 0
		busy = [False, False]
		for r in range(N_VIEWS):
			k = r & 1
			if busy[k]:
				pms[k].Wait(); h2d += pms[k].stats.bytes_h2d; d2h += pms[k].stats.bytes_d2h
			m = h_maps[r]
			m["depth"][:] = 0  # depth 0 => random initialisation (the normal is ignored then)
			dd = DepthData([ViewData(h_imgs[r], cams[r])]+[ViewData(h_imgs[i], cams[i]) for i in nbrs[r]], scene.dmin, scene.dmax,
				depthMap=m["depth"], normalMap=m["normal"], confMap=m["conf"], viewsMap=m["views"])
			pms[k].EstimateDepthMap(dd, sync=False)
			busy[k] = True
		for k in range(2):
			if busy[k]:
				pms[k].Wait(); h2d += pms[k].stats.bytes_h2d; d2h += pms[k].stats.bytes_d2h
		return h2d, d2h

	def timed(fn, steps, warmup):
		for _ in range(warmup):
			fn()
		torch.cuda.synchronize()
		if world > 1: dist.barrier()
		torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		ret = None
		for _ in range(steps):
			ret = fn()
		e1.record()
		torch.cuda.synchronize()
		if world > 1: dist.barrier()
		ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
		if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
		return float(ms.item()), ret

	if os.environ.get("BENCH_DEBUG"):
		for tag in ("cold", "warm"):
			dd0 = DepthData([ViewData(d_imgs[5], cams[5])]+[ViewData(d_imgs[i], cams[i]) for i in nbrs[5]], scene.dmin, scene.dmax)
			pm.EstimateDepthMap(dd0, sync=True)
			print("debug %s: view 5 device %.2f ms, sweep avg %.3f ms" % (tag, pm.stats.ms_device, pm.stats.ms_sweep_kernels/max(1, pm.stats.sweep_launches)), file=sys.stderr)
	sampler = ClockSampler(local_rank)
	if rank == 0 and not os.environ.get("BENCH_NO_SMI"): sampler.start()
	launches[0] = 0
	ms_res, _ = timed(step_resident, args.steps, args.warmup)
	n_launch = launches[0]*args.steps//(args.steps+args.warmup)
	clocks = sampler.stop() if rank == 0 else None
	ms_e2e, (h2d, d2h) = timed(step_e2e, args.steps, max(1, args.warmup//3))
	mpix_step = N_VIEWS*w*h/1e6*world
	value = mpix_step/(ms_res/args.steps/1e3)
	e2e = mpix_step/(ms_e2e/args.steps/1e3)

	# ---- roofline of the dominant kernel: one red-black half-sweep, timed live --------------------
	# CUDA events are recorded by the engine around every pm_sweep_kernel launch of one more (untimed)
	# resident EstimateDepthMap call, on the stream the kernels are launched on (b200mvs_stats)
	r = 5
	m = d_maps[r]
	m["depth"].zero_(); m["normal"].zero_()
	dd = DepthData([ViewData(d_imgs[r], cams[r])]+[ViewData(d_imgs[i], cams[i]) for i in nbrs[r]], scene.dmin, scene.dmax,
		depthMap=m["depth"], normalMap=m["normal"], confMap=m["conf"], viewsMap=m["views"])
	pm.EstimateDepthMap(dd, sync=True)
	k_ms = pm.stats.ms_sweep_kernels/max(1, pm.stats.sweep_launches)
	sweep_share = pm.stats.ms_sweep_kernels/max(1e-9, pm.stats.ms_device)
	bytes_launch = w*h*(20+10+4*(N_NEIGH+1))
	peaks = {}
	try:
		peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
	except Exception:
		pass
	peak = float(peaks.get("hbm_gbs", 6650.0))
	achieved = bytes_launch/(k_ms*1e-3)/1e9
	nR = N_REFINE
	samples_launch = (w*h/2)*(4+nR)*N_NEIGH*25
	roof = {"kernel": "pm_sweep_kernel<true,false> (one red-black half-sweep; taps evaluated in FMUL2/FFMA2 pairs)", "bound": "hbm", "achieved": achieved, "peak": peak,
		"unit": "GB/s", "frac": achieved/peak, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
		"launch_ms": k_ms, "launches_timed": int(pm.stats.sweep_launches), "share_of_step": sweep_share, "algorithmic_bytes_per_launch": bytes_launch,
		"secondary": {"bound": "issue/L1 (gather stencil, AI ~ 280 flop/B)", "bilinear_samples_per_launch": samples_launch,
			"gsamples_per_s": samples_launch/(k_ms*1e-3)/1e9}}
	# the honest bound of this kernel is the L1 LSU data pipe: quote its utilisation from the committed ncu capture
	# (taken on the scalar-tap variant <1,false,true>; the packed-tap default issues about 10 % fewer instructions per tap)
	try:
		for line in open(os.path.join(ROOT, "profiles", "ncu_pm_sweep_r01.txt")):
			if line.startswith("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"):
				roof["secondary"]["l1_lsu_data_pipe_pct_of_peak_ncu"] = float(line.split()[1])
			if line.startswith("smsp__issue_active.avg.pct_of_peak_sustained_active"):
				roof["secondary"]["issue_slots_busy_pct_ncu"] = float(line.split()[1])
	except Exception:
		pass
	traffic_file = os.path.join(ROOT, "profiles", "sweep_traffic.json")
	if os.path.exists(traffic_file):
		try:
			roof["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
		except Exception:
			pass

	if rank == 0:
		cpu = None
		if not args.no_cpu_baseline:
			band = args.cpu_band or max(24, min(h, int(round(h*0.25*threads/8.0))))
			mp, dt, desc, _ = cpu_sample(scene, w, h, band, threads)
			cpu = {"value": mp, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": desc, "seconds": dt}
		out = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
			"ms_per_step": ms_res/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
			"data": "synthetic",
			"config": {"workload": "C2: 12x1920x1080 synthetic scene, 9 neighbours, PatchMatch 6 iters, single scale" if not args.small else "dev 12x640x360",
				"views_per_gpu_per_step": N_VIEWS, "neighbours": N_NEIGH, "iters": ITERS,
				"schedule": "red-black, %d sweeps for %d reference iterations, <= 4 propagation candidates (best of distances 1/3/5 per direction, unchanged ones skipped) + %d refinements per sweep" % (N_SWEEPS, ITERS, nR),
				"parallelism": "per-reference-view shards, %d GPU(s), NCCL gather of maps" % world,
				"l2": "no flush: a step streams %.2f GB of distinct images+maps (> 126 MB L2)" % ((N_VIEWS*w*h*(4+24+20))/1e9)},
			"clocks": clocks,
			"e2e": {"value": e2e, "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e/args.steps},
			"gpu_launches": int(n_launch), "roofline": roof, "cpu_baseline": cpu}
		print(json.dumps(out))
	if world > 1:
		dist.destroy_process_group()
	return 0


def _maybe_cuda():
	try:
		import torch
		if torch.cuda.is_available():
			return torch.device("cuda", 0)
	except Exception:
		pass
	return None


if __name__ == "__main__":
	sys.exit(main())
