"""Generates tests/golden/fixture4.npz from the reference's own 4-image test scene
(apps/Tests/data/scene.mvs + images/0000{0..3}.jpg, the input of PipelineTest, apps/Tests/Tests.cpp:76-113).

Run HERE (the container with /root/reference); the GPU box only reads the committed .npz.
Uses the reference's Python reader scripts/python/MvsUtils.py:loadMVSInterface for the cameras and
sparse points, converts the JPEGs to the estimator's gray float images exactly like
Image::toGray(..., bNormalize=true) does (coefficients .114/.587/.299 on BGR, /255,
libs/Common/Types.inl:2382-2420), and halves the resolution with INTER_AREA + Camera::ScaleK so that
the fixture stays small (4 x 320x240 uint16 PNG-equivalent, stored as uint16 in a compressed npz).
"""
import os
import sys

import cv2
import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "scripts", "python"))
from MvsUtils import loadMVSInterface  # noqa: E402

mvs = loadMVSInterface(os.path.join(REF, "apps", "Tests", "data", "scene.mvs"))
plat = mvs["platforms"][0]
cam = plat["cameras"][0]
Kn = np.array(cam["K"], np.float64)
images, Ks, Rs, Cs = [], [], [], []
for im in mvs["images"]:
	path = os.path.join(REF, "apps", "Tests", "data", "images", os.path.basename(im["name"]))
	bgr = cv2.imread(path, cv2.IMREAD_COLOR)
	h, w = bgr.shape[:2]
	gray = (0.114*bgr[..., 0].astype(np.float32)+0.587*bgr[..., 1].astype(np.float32)+0.299*bgr[..., 2].astype(np.float32))/255.0
	# the interface stores K normalised by max(w,h) when width/height are absent; here K is in pixels of the full image
	K = Kn.copy()
	if cam.get("width", 0) == 0 or K[0, 0] < 10:
		s = max(w, h)
		K = np.array([[K[0, 0]*s, 0, K[0, 2]*s], [0, K[1, 1]*s, K[1, 2]*s], [0, 0, 1]])
	# MvsUtils stores the platform poses under the camera; the camera's own R/C are identity/zero in this scene
	pose = cam["poses"][im["pose_id"]]
	R = np.array(pose["R"], np.float64)
	C = np.array(pose["C"], np.float64)
	dw, dh = int(np.rint(w*0.5)), int(np.rint(h*0.5))
	small = cv2.resize(gray, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
	assert small.shape == (dh, dw)
	sx, sy = dw/w, dh/h
	Ks.append(np.array([[K[0, 0]*sx, K[0, 1]*sx, (K[0, 2]+0.5)*sx-0.5], [0, K[1, 1]*sy, (K[1, 2]+0.5)*sy-0.5], [0, 0, 1]]))
	Rs.append(R); Cs.append(C)
	images.append(np.rint(np.clip(small, 0, 1)*65535).astype(np.uint16))
# depth range of view 0 from the sparse points it sees (InitViews / InitDepthMap: dMin*0.9, dMax*1.1)
pts = np.array([v["X"] for v in mvs["vertices"] if any(view["image_id"] == 0 for view in v["views"])], np.float64)
Xc = Rs[0] @ (pts-Cs[0]).T
z = Xc[2]
uv = (Ks[0] @ Xc)[:2]/z
keep = (z > 0) & (uv[0] >= 6) & (uv[1] >= 6) & (uv[0] <= images[0].shape[1]-7) & (uv[1] <= images[0].shape[0]-7)
sparse = np.stack([uv[0][keep], uv[1][keep], z[keep]], 1).astype(np.float32)  # the reference scene's own SfM points in view 0
z = z[z > 0]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixture4.npz"),
	images=np.stack(images), K=np.stack(Ks), R=np.stack(Rs), C=np.stack(Cs),
	dmin=np.float32(z.min()*0.9), dmax=np.float32(z.max()*1.1), n_points=np.int32(len(z)), sparse=sparse)
print("fixture4.npz:", np.stack(images).shape, "depth range", z.min()*0.9, z.max()*1.1, "points", len(z))
