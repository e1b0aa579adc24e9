""".dmap depth-data files — the wire format either side of the hot path.

ExportDepthDataRaw / ImportDepthDataRaw follow the reference's raw layout
(libs/MVS/DepthMap.cpp:1874-2037, header `HeaderDepthDataRaw` libs/MVS/Interface.h:773-792):

  u16 'DR' | u8 type (HAS_DEPTH=1, HAS_NORMAL=2, HAS_CONF=4, HAS_VIEWS=8) | u8 padding
  u32 imageWidth, imageHeight | u32 depthWidth, depthHeight | f32 dMin, dMax
  u16 nFileNameSize, chars | u32 nIDs, u32 IDs[nIDs] (reference view first)
  f64 K[9], R[9], C[3] | f32 depth[h][w] | f32 normal[h][w][3] | f32 conf[h][w] | u8 views[h][w][4]

Host-side plumbing (numpy); files are written atomically (.tmp + rename) like DepthData::Save
(libs/MVS/DepthMap.cpp:237-251).
"""
from __future__ import annotations

import os
import struct

import numpy as np

HAS_DEPTH, HAS_NORMAL, HAS_CONF, HAS_VIEWS = 1, 2, 4, 8


def ExportDepthDataRaw(fileName, imageFileName, IDs, imageSize, K, R, C, dMin, dMax, depthMap, normalMap=None, confMap=None, viewsMap=None) -> bool:
	depth = np.ascontiguousarray(depthMap, np.float32)
	h, w = depth.shape
	if not (1 < len(IDs) < 256) or w > imageSize[0] or h > imageSize[1]:
		raise ValueError("IDs must hold the reference and at least one neighbour; the depth-map may not exceed the image")
	typ = HAS_DEPTH
	blobs = [depth.tobytes()]
	for arr, flag, shape, dt in ((normalMap, HAS_NORMAL, (h, w, 3), np.float32), (confMap, HAS_CONF, (h, w), np.float32), (viewsMap, HAS_VIEWS, (h, w, 4), np.uint8)):
		if arr is not None:
			a = np.ascontiguousarray(arr, dt)
			if a.shape != shape:
				raise ValueError("map shape %s, expected %s" % (a.shape, shape))
			typ |= flag
			blobs.append(a.tobytes())
	name = os.path.relpath(os.path.abspath(imageFileName), os.path.dirname(os.path.abspath(fileName))).encode()
	tmp = fileName+".tmp"
	with open(tmp, "wb") as f:
		f.write(b"DR"+struct.pack("<BBIIIIff", typ, 0, imageSize[0], imageSize[1], w, h, float(dMin), float(dMax)))
		f.write(struct.pack("<H", len(name))+name)
		f.write(struct.pack("<I", len(IDs))+np.asarray(IDs, np.uint32).tobytes())
		f.write(np.asarray(K, np.float64).reshape(9).tobytes()+np.asarray(R, np.float64).reshape(9).tobytes()+np.asarray(C, np.float64).reshape(3).tobytes())
		for b in blobs:
			f.write(b)
	os.replace(tmp, fileName)
	return True


def ImportDepthDataRaw(fileName, flags: int = 15) -> dict:
	"""-> dict(imageFileName, IDs, imageSize, K, R, C, dMin, dMax, depthMap, normalMap, confMap, viewsMap);
	flags selects which optional maps to return (HAS_* bits), like the reference's `flags` argument."""
	with open(fileName, "rb") as f:
		head = f.read(28)
		if len(head) != 28 or head[:2] != b"DR":
			raise ValueError("%s is not a depth-data file" % fileName)
		typ, _, iw, ih, w, h, dmin, dmax = struct.unpack("<BBIIIIff", head[2:])
		if not (typ & HAS_DEPTH) or w == 0 or h == 0 or iw < w or ih < h:
			raise ValueError("invalid depth-data header")
		n, = struct.unpack("<H", f.read(2))
		name = f.read(n).decode()
		nids, = struct.unpack("<I", f.read(4))
		ids = np.frombuffer(f.read(4*nids), np.uint32).copy()
		K = np.frombuffer(f.read(72), np.float64).reshape(3, 3).copy()
		R = np.frombuffer(f.read(72), np.float64).reshape(3, 3).copy()
		Cc = np.frombuffer(f.read(24), np.float64).copy()
		out = dict(imageFileName=name, IDs=ids, imageSize=(iw, ih), K=K, R=R, C=Cc, dMin=dmin, dMax=dmax,
			normalMap=None, confMap=None, viewsMap=None)
		out["depthMap"] = np.frombuffer(f.read(4*w*h), np.float32).reshape(h, w).copy()
		for key, flag, count, dt, shape in (("normalMap", HAS_NORMAL, 12*w*h, np.float32, (h, w, 3)), ("confMap", HAS_CONF, 4*w*h, np.float32, (h, w)), ("viewsMap", HAS_VIEWS, 4*w*h, np.uint8, (h, w, 4))):
			if typ & flag:
				raw = f.read(count)
				if flags & flag:
					out[key] = np.frombuffer(raw, dt).reshape(shape).copy()
	return out
