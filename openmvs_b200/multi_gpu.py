"""Sharding of reference views across the GPUs of one box.

The unit of work is one reference view (one DepthMapsData::EstimateDepthMap call,
libs/MVS/SceneDensify.cpp:616-805): it reads its own image and its neighbours' images and
writes its own maps, so views are independent in the photometric pass.  One process per GPU
(torch.distributed), reference views dealt round-robin, no data-path collective during the
estimation.  Two exchanges exist:
  * gather_maps      final gather of depth/normal/confidence to one rank (NCCL gather)
  * all_gather_depth depth-only all-gather before a geometric-consistency pass, which needs
                     the neighbours' pass-1 depth-maps (the reference reloads them from
                     .dmap files, libs/MVS/SceneDensify.cpp:380-394); the same exchange with
                     depth+confidence precedes the filter pass (filter_depth_maps), where the
                     reference reloads the neighbours' .dmap files again (SceneDensify.cpp:2149-2167)
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
	"""reference views estimated by `rank`: round-robin, so neighbouring views land on
	different GPUs and per-rank counts differ by at most one"""
	return list(range(rank, n_views, world))


def owner_of(view: int, world: int) -> int:
	return view % world


def gather_maps(local: Dict[int, torch.Tensor], n_views: int, dst: int = 0):
	"""Gather per-view maps to `dst`.  local: {view index: tensor (H, W, C)} for the views of
	this rank (all views share one shape).  Returns {view: tensor} on dst, None elsewhere."""
	if not dist.is_initialized() or dist.get_world_size() == 1:
		return dict(local)
	rank, world = dist.get_rank(), dist.get_world_size()
	mine = shard_views(n_views, rank, world)
	assert sorted(local.keys()) == mine, "rank %d holds %s, expected %s" % (rank, sorted(local.keys()), mine)
	kmax = (n_views+world-1)//world
	any_t = next(iter(local.values())) if local else None
	shape_t = torch.zeros(4, dtype=torch.int64, device=any_t.device if any_t is not None else _default_device())
	if any_t is not None:
		shape_t[:any_t.dim()] = torch.tensor(any_t.shape, device=shape_t.device)
	dist.all_reduce(shape_t, op=dist.ReduceOp.MAX)
	shape = tuple(int(v) for v in shape_t.tolist() if v > 0)
	dtype = any_t.dtype if any_t is not None else torch.float32
	buf = torch.zeros((kmax,)+shape, dtype=dtype, device=shape_t.device)
	for k, v in enumerate(mine):
		buf[k].copy_(local[v])
	out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
	dist.gather(buf, out, dst=dst)
	if rank != dst:
		return None
	res = {}
	for r in range(world):
		for k, v in enumerate(shard_views(n_views, r, world)):
			res[v] = out[r][k]
	return res


def all_gather_depth(local: Dict[int, torch.Tensor], n_views: int) -> Dict[int, torch.Tensor]:
	"""Every rank receives the depth-map of every view (exchange step before a geometric pass)."""
	if not dist.is_initialized() or dist.get_world_size() == 1:
		return dict(local)
	rank, world = dist.get_rank(), dist.get_world_size()
	mine = shard_views(n_views, rank, world)
	kmax = (n_views+world-1)//world
	any_t = next(iter(local.values()))
	buf = torch.zeros((kmax,)+tuple(any_t.shape), dtype=any_t.dtype, device=any_t.device)
	for k, v in enumerate(mine):
		buf[k].copy_(local[v])
	out = [torch.empty_like(buf) for _ in range(world)]
	dist.all_gather(out, buf)
	res = {}
	for r in range(world):
		for k, v in enumerate(shard_views(n_views, r, world)):
			res[v] = out[r][k]
	return res


class ViewStack:
	"""The maps of this rank's reference views as preallocated stacks — depth [K,H,W], normal [K,H,W,3], conf [K,H,W],
	views [K,H,W,4] (K = views per rank, rounded up) — so that EstimateDepthMap writes straight into slices of them and the
	two exchanges take the stacks as they are: no packing, no staging copies, no per-step allocation.
	  gather(dst)          ncclGather of depth, normal and confidence to `dst` (the final hand-over, 20 B per pixel)
	  all_gather_depth()   ncclAllGather of the depth stack (4 B per pixel) before a geometric-consistency pass"""

	def __init__(self, n_views: int, height: int, width: int, device, rank: int = None, world: int = None):
		self.rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
		self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else world
		self.n_views, self.h, self.w = n_views, height, width
		self.mine = shard_views(n_views, self.rank, self.world)
		self.kmax = (n_views+self.world-1)//self.world
		z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
		self.depth, self.normal, self.conf = z(self.kmax, height, width), z(self.kmax, height, width, 3), z(self.kmax, height, width)
		self.views = z(self.kmax, height, width, 4, dt=torch.uint8)
		self.all_depth = None   # [world*K, H, W] after all_gather_depth()
		self._recv = None       # receive stacks on the gather destination

	def slot(self, view: int) -> int:
		return self.mine.index(view)

	def maps(self, view: int) -> dict:
		k = self.slot(view)
		return dict(depth=self.depth[k], normal=self.normal[k], conf=self.conf[k], views=self.views[k])

	def gather(self, dst: int = 0):
		"""-> {view: dict(depth, normal, conf)} on dst (views of the receive stacks), None elsewhere"""
		if self.world == 1:
			return {v: self.maps(v) for v in self.mine}
		if self.rank == dst and self._recv is None:
			self._recv = [[torch.empty_like(t) for _ in range(self.world)] for t in (self.depth, self.normal, self.conf)]
		for i, t in enumerate((self.depth, self.normal, self.conf)):
			dist.gather(t, self._recv[i] if self.rank == dst else None, dst=dst)
		if self.rank != dst:
			return None
		res = {}
		for r in range(self.world):
			for k, v in enumerate(shard_views(self.n_views, r, self.world)):
				res[v] = dict(depth=self._recv[0][r][k], normal=self._recv[1][r][k], conf=self._recv[2][r][k])
		return res

	def all_gather_depth(self):
		"""every rank receives every view's depth-map; depth_of(view) indexes the result"""
		if self.world == 1:
			self.all_depth = self.depth
			return self.all_depth
		if self.all_depth is None or self.all_depth is self.depth:
			self.all_depth = torch.empty((self.world*self.kmax, self.h, self.w), dtype=torch.float32, device=self.depth.device)
		dist.all_gather_into_tensor(self.all_depth, self.depth)
		return self.all_depth

	def depth_of(self, view: int):
		return self.all_depth[owner_of(view, self.world)*self.kmax + view//self.world]

	def bytes_gather(self) -> int:
		return self.kmax*self.h*self.w*20

	def bytes_all_gather(self) -> int:
		return self.world*self.kmax*self.h*self.w*4


def _default_device():
	if dist.is_initialized() and dist.get_backend() == "nccl":
		return torch.device("cuda", torch.cuda.current_device())
	return torch.device("cpu")


def estimate_scene(n_views: int, estimate_view: Callable[[int], torch.Tensor], dst: int = 0, gather: bool = True):
	"""Run `estimate_view(idx) -> (H, W, C) tensor` for the reference views of this rank and
	gather the results on `dst` (NCCL only for that final gather)."""
	rank = dist.get_rank() if dist.is_initialized() else 0
	world = dist.get_world_size() if dist.is_initialized() else 1
	local = {v: estimate_view(v) for v in shard_views(n_views, rank, world)}
	if not gather:
		return local
	return gather_maps(local, n_views, dst)


def compute_depth_maps(n_views: int, estimate: Callable, n_geometric_iters: int = 0, dst: int = 0, gather: bool = True):
	"""The estimation part of Scene::ComputeDepthMaps (libs/MVS/SceneDensify.cpp:1754-1953) over the ranks:

	  pass 1            every rank estimates its reference views photometrically (nGeometricIter = -1)
	  geometric pass g  one exchange step — all-gather of the depth-maps of the previous pass, because a
	                    reference view needs its neighbours' depth-maps (the reference reloads their .dmap
	                    files, SceneDensify.cpp:380-394) — then every rank re-estimates its views with
	                    nGeometricIter = g, initialised from its own previous result

	estimate(view, nGeometricIter, previous, depths) -> dict(depth=(H,W), normal=(H,W,3), conf=(H,W)) of tensors;
	`previous` is this view's result of the pass before (None in pass 1), `depths` the gathered
	{view: depth} of all views (None in pass 1).  Returns {view: (H,W,5) depth|normal|conf} on `dst`."""
	rank = dist.get_rank() if dist.is_initialized() else 0
	world = dist.get_world_size() if dist.is_initialized() else 1
	mine = shard_views(n_views, rank, world)
	local = {v: estimate(v, -1, None, None) for v in mine}
	for g in range(n_geometric_iters):
		depths = all_gather_depth({v: local[v]["depth"] for v in mine}, n_views) if mine or world > 1 else {}
		local = {v: estimate(v, g, local[v], depths) for v in mine}
	packed = {v: torch.cat([m["depth"][..., None], m["normal"], m["conf"][..., None]], -1) for v, m in local.items()}
	if not gather:
		return packed
	return gather_maps(packed, n_views, dst)


def filter_depth_maps(n_views: int, local: Dict[int, dict], neighbors: Sequence[Sequence[int]], filter_view: Callable,
		max_neighbors: int = 8) -> Dict[int, dict]:
	"""The filter pass of Scene::DenseReconstructionFilter (libs/MVS/SceneDensify.cpp:2141-2181) over the ranks: one
	exchange step (all-gather of depth+confidence, 8 B per pixel and view), then every rank filters its own views against
	at most `max_neighbors` (numMaxNeighbors = 8) valid neighbour maps.  All views are filtered against the UNFILTERED
	maps of the pass before — the reference swaps the filtered maps in only after every view is done (EVT_ADJUSTDEPTHMAP).

	local: {view: dict(depth=(H,W), conf=(H,W))} of this rank; neighbors[v]: neighbour views of v, best first;
	filter_view(v, ref, nbrs) -> (depth, conf) or None, with ref = dict(depth, conf) and nbrs = [(view, depth, conf), ...].
	Returns {view: dict(depth, conf)} for the views of this rank (unfilterable views keep their maps)."""
	rank = dist.get_rank() if dist.is_initialized() else 0
	world = dist.get_world_size() if dist.is_initialized() else 1
	mine = shard_views(n_views, rank, world)
	assert sorted(local.keys()) == mine
	packed = {v: torch.stack([local[v]["depth"], local[v]["conf"]], -1) for v in mine}
	allm = all_gather_depth(packed, n_views) if mine or world > 1 else {}
	out = {}
	for v in mine:
		nbrs = []
		for i in neighbors[v]:
			if i == v or i not in allm:
				continue
			d, c = allm[i][..., 0], allm[i][..., 1]
			if not bool((d > 0).any()):   # !depthDataPair.IsValid()
				continue
			nbrs.append((i, d.contiguous(), c.contiguous()))
			if len(nbrs) == max_neighbors:
				break
		res = filter_view(v, local[v], nbrs)
		out[v] = dict(depth=res[0], conf=res[1]) if res is not None else dict(depth=local[v]["depth"], conf=local[v]["conf"])
	return out


class SceneFilter:
	"""filter_view callable for filter_depth_maps on one GPU (DepthMapsData.FilterDepthMap on device tensors).
	views: objects with .K .R .C; dmin/dmax: the depth range FilterDepthMap clips the adjusted depths to."""

	def __init__(self, views, dmin: float, dmax: float, device=None, bAdjust: bool = True):
		from .depth_estimator import Camera, DepthMapsData
		self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
		self.dm = DepthMapsData([], self.dev.index or 0, nCalibratedImages=len(views))
		self.cams = [Camera(v.K, v.R, v.C) for v in views]
		self.dmin, self.dmax, self.bAdjust = dmin, dmax, bAdjust

	def _dd(self, v, depth, conf):
		from .depth_estimator import DepthData, ViewData
		return DepthData([ViewData(None, self.cams[v])], self.dmin, self.dmax, depthMap=depth.to(self.dev).contiguous(),
			confMap=conf.to(self.dev).contiguous())

	def __call__(self, v, ref, nbrs):
		return self.dm.FilterDepthMap(self._dd(v, ref["depth"], ref["conf"]), [self._dd(i, d, c) for i, d, c in nbrs], self.bAdjust)


class SceneEstimator:
	"""estimate() callable for compute_depth_maps on one GPU: images resident in HBM, one PatchMatchB200.
	views: list of objects with .image (numpy HxW float32) .K .R .C; neighbors: list of index lists."""

	def __init__(self, views, neighbors, dmin: float, dmax: float, device=None, writer=None, path_of=None, image_names=None):
		"""writer / path_of: emit every estimated view as a `.dmap` file without waiting for the disk — writer is a
		dmap_io.AsyncDepthDataWriter, path_of(view, nGeometricIter) the file name (the reference writes depthNNNN.dmap after pass 1
		and depthNNNN.geo.dmap after a geometric pass, SceneDensify.cpp:2113); image_names[v]: the image file of view v."""
		from .depth_estimator import Camera, PatchMatchB200
		self.writer, self.path_of, self.image_names = writer, path_of, image_names
		self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
		self.pm = PatchMatchB200(self.dev.index or 0)
		self.cams = [Camera(v.K, v.R, v.C) for v in views]
		self.imgs = [torch.from_numpy(v.image).to(self.dev) for v in views]
		self.neighbors = neighbors
		self.dmin, self.dmax = dmin, dmax

	def __call__(self, v: int, nGeometricIter: int, previous, depths):
		from .depth_estimator import DepthData, ViewData
		images = [ViewData(self.imgs[v], self.cams[v])]
		for i in self.neighbors[v]:
			vd = ViewData(self.imgs[i], self.cams[i])
			if depths is not None:
				vd.depthMap = depths[i].to(self.dev).contiguous(); vd.cameraDepthMap = self.cams[i]
			images.append(vd)
		dd = DepthData(images, self.dmin, self.dmax)
		if previous is not None:
			dd.depthMap = previous["depth"].clone(); dd.normalMap = previous["normal"].clone()
		self.pm.Init(nGeometricIter >= 0)
		self.pm.EstimateDepthMap(dd, nGeometricIter)
		if self.writer is not None:
			# EVT_SAVEDEPTHMAP (SceneDensify.cpp:2095-2113): the copy to pinned memory is ordered after the estimation on this
			# stream, the file is written by the writer's thread while the next view is estimated
			h, w = dd.depthMap.shape
			cam = self.cams[v]
			self.writer.submit(self.path_of(v, nGeometricIter), self.image_names[v] if self.image_names else "%05d.jpg" % v,
				[v]+list(self.neighbors[v]), (w, h), cam.K, cam.R, cam.C, self.dmin, self.dmax, dd.depthMap, dd.normalMap, dd.confMap, dd.viewsMap)
		return dict(depth=dd.depthMap, normal=dd.normalMap, conf=dd.confMap)
