"""world_size-2 gloo test of the N>1 path: sharding of reference views, the final gather, the depth all-gather
that precedes a geometric pass and the depth+confidence exchange of the filter pass (no GPU: the per-view
estimator / filter are stubs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_views_list, q):
	sys.path.insert(0, ROOT)
	from openmvs_b200 import multi_gpu
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		oks = [_check(rank, world, n_views, multi_gpu) for n_views in n_views_list]
		q.put((rank, all(oks)))
	finally:
		dist.destroy_process_group()


def _check(rank, world, n_views, multi_gpu):
	if True:
		def estimate(v):  # stub estimator: a map that identifies its view
			return torch.full((6, 8, 5), float(v))+torch.arange(5, dtype=torch.float32)
		res = multi_gpu.estimate_scene(n_views, estimate, dst=0)
		if rank == 0:
			ok = sorted(res.keys()) == list(range(n_views)) and all(
				torch.equal(res[v], estimate(v)) for v in range(n_views))
		else:
			ok = res is None
		local = {v: torch.full((6, 8), float(v)) for v in multi_gpu.shard_views(n_views, rank, world)}
		allv = multi_gpu.all_gather_depth(local, n_views)
		ok = ok and sorted(allv.keys()) == list(range(n_views)) and all(float(allv[v][0, 0]) == v for v in range(n_views))
		# pass 1 + two geometric passes: every geometric pass must see ALL views' depth-maps of the pass before
		seen = []
		def est(v, g, previous, depths):
			if g >= 0:
				seen.append((v, g, previous is not None, sorted(depths.keys()) == list(range(n_views)),
					all(float(depths[u][0, 0]) == 100*(g)+u for u in range(n_views))))
			val = 100.0*(g+1)+v
			return dict(depth=torch.full((6, 8), val), normal=torch.zeros(6, 8, 3), conf=torch.ones(6, 8))
		res = multi_gpu.compute_depth_maps(n_views, est, n_geometric_iters=2, dst=0)
		mine = multi_gpu.shard_views(n_views, rank, world)
		ok = ok and len(seen) == 2*len(mine) and all(s[2] and s[3] and s[4] for s in seen)
		if rank == 0:
			ok = ok and all(float(res[v][0, 0, 0]) == 200.0+v for v in range(n_views))
		# filter pass: one depth+confidence exchange, every view sees at most 3 valid neighbours, unfiltered inputs
		nb = [[(v+k) % n_views for k in (1, 2, 3, 4)] for v in range(n_views)]
		loc = {v: dict(depth=torch.full((6, 8), 1.0+v) if v != 1 else torch.zeros(6, 8), conf=torch.full((6, 8), 0.5+v)) for v in mine}
		calls = []
		def flt(v, ref, nbrs):
			calls.append((v, [i for i, _, _ in nbrs], all(float(d[0, 0]) == 1.0+i and float(c[0, 0]) == 0.5+i for i, d, c in nbrs)))
			return None if v % 3 == 0 else (ref["depth"]*2, ref["conf"]*3)
		fl = multi_gpu.filter_depth_maps(n_views, loc, nb, flt, max_neighbors=3)
		ok = ok and sorted(fl.keys()) == mine and [c[0] for c in calls] == mine
		for v, ids, vals in calls:
			expect = [i for i in nb[v] if i != 1][:3]   # view 1 has an empty depth-map: skipped like !IsValid()
			ok = ok and ids == expect and vals
		for v in mine:
			f = 1.0 if v % 3 == 0 else 2.0
			ok = ok and float(fl[v]["depth"][0, 0]) == f*float(loc[v]["depth"][0, 0])
		# ViewStack: preallocated stacks, estimates written into slices, collectives on the stacks (bench / scene path)
		st = multi_gpu.ViewStack(n_views, 6, 8, torch.device("cpu"))
		ok = ok and st.mine == mine and st.kmax == (n_views+world-1)//world
		for v in mine:
			m = st.maps(v)
			m["depth"].fill_(10.0+v); m["normal"].fill_(0.25*v); m["conf"].fill_(0.5+v)
		st.all_gather_depth()
		ok = ok and all(float(st.depth_of(v)[0, 0]) == 10.0+v and st.depth_of(v).shape == (6, 8) for v in range(n_views))
		for rep in range(2):   # the receive buffers are reused
			g = st.gather(dst=0)
			if rank == 0:
				ok = ok and sorted(g.keys()) == list(range(n_views)) and all(
					float(g[v]["depth"][0, 0]) == 10.0+v and float(g[v]["normal"][0, 0, 2]) == 0.25*v and float(g[v]["conf"][5, 7]) == 0.5+v for v in range(n_views))
			else:
				ok = ok and g is None
		ok = ok and st.bytes_all_gather() == world*st.kmax*6*8*4
		return bool(ok)


def test_shard_gather_world2():
	n_views = [5, 12]  # odd and even shard sizes, checked inside one pair of spawned processes
	s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
	for p in procs: p.start()
	got = [q.get(timeout=600) for _ in procs]  # a cold `import torch` in the children can take a minute
	for p in procs: p.join(120)
	assert sorted(got) == [(0, True), (1, True)]
