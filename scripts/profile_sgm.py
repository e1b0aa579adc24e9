"""1080p SGM pair for profiling: per-stage device times of the aggregation variants (b200mvs_debug.sgmAggregation) and of the
wave-front layouts / block sizes, with an identity check against the register-pipelined kernel.
usage: profile_sgm.py [D] [default]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import SemiGlobalMatcher
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w, h = 1920, 1080
lg, lc, rg, d = synth.make_stereo_pair(w, h, d0=40.0, amp=25.0)
px, n = synth.sgm_pixel_map(w, h, 0, D)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
args = (dev(lg), dev(lc), dev(rg), torch.from_numpy(px.view(np.uint8).reshape(-1, 16).copy()).cuda(), n)
m = SemiGlobalMatcher()
costs = torch.zeros(n, dtype=torch.uint8, device="cuda"); accums = torch.zeros(n, dtype=torch.int16, device="cuda")
ref = None
MODES = [("register pipeline", dict(sgmAggregation=2)), ("bulk-copy ring, 8 launches", dict(sgmAggregation=3)),
	("wave fronts (default)", dict())]
for layout, name in ((1, "tilted"), (2, "straight"), (3, "single")):
	for serial in (0, 1):
		MODES.append(("fronts %s %s" % (name, "serial" if serial else "paired"), dict(sgmAggregation=4, frontLayout=layout, frontSerial=serial)))
for layout, name in ((1, "tilted"), (2, "straight")):
	for fbk in (32, 64):
		for lag in (0, 1, 2):
			for ctas in (2, 4):
				MODES.append(("fronts %s paired FB %d lag %d ctas %d" % (name, fbk, lag, ctas), dict(sgmAggregation=4, frontLayout=layout, frontBlock=fbk, frontLag=lag+1, frontCtas=ctas)))
MODES += [("fronts tilted paired FB 16 lag 0", dict(sgmAggregation=4, frontBlock=16)), ("fronts tilted paired FB 128 lag 0", dict(sgmAggregation=4, frontBlock=128)),
	("fronts tilted paired FB 64 lag 0 depth 4", dict(sgmAggregation=4, frontDepth=4)), ("fronts tilted paired FB 64 lag 0 ctas 1", dict(sgmAggregation=4, frontCtas=1)),
	("fronts tilted paired FB 64 lag 0 ctas 3", dict(sgmAggregation=4, frontCtas=3))]
MODES.append(("SIMT cost kernel + wave fronts", dict(sgmCost=1)))
if len(sys.argv) > 2 and sys.argv[2] == "default":
	MODES = MODES[2:3]
if len(sys.argv) > 2 and sys.argv[2] == "lag0":
	MODES = [("fronts FB 32 lag 0", dict(sgmAggregation=4, frontLag=1))]
if len(sys.argv) > 2 and sys.argv[2] == "lag2":
	MODES = [("fronts FB 32 lag 2", dict(sgmAggregation=4, frontLag=3))]
if len(sys.argv) > 2 and sys.argv[2] == "lag1":
	MODES = [("fronts FB 32 lag 1", dict(sgmAggregation=4, frontLag=2))]
if len(sys.argv) > 2 and sys.argv[2] == "fb64":
	MODES = [("fronts FB 64 lag 2", dict(sgmAggregation=4, frontBlock=64))]
if len(sys.argv) > 2 and sys.argv[2] == "order":
	MODES = MODES[:1]+[("fronts default (FB 32 lag 2)", dict())]
	for fbk in (32, 64):
		for lag in (0, 1, 2):
			for sw in (128, 64):
				for ctas in (2, 3):
					MODES.append(("fronts FB %d lag %d subcell %d ctas %d" % (fbk, lag, sw, ctas), dict(sgmAggregation=4, frontBlock=fbk, frontLag=lag+1, frontSubCell=sw, frontCtas=ctas)))
	MODES.append(("fronts FB 32 lag 1 subcell 32", dict(sgmAggregation=4, frontLag=2, frontSubCell=32)))
	MODES.append(("fronts FB 32 lag 1 ctas 1", dict(sgmAggregation=4, frontLag=2, frontCtas=1)))
	MODES.append(("fronts straight FB 32 lag 1", dict(sgmAggregation=4, frontLayout=2, frontLag=2)))
if len(sys.argv) > 2 and sys.argv[2] == "tc":
	MODES = MODES[-1:]
for name, dbg in MODES:
	m.SetDebug(**dbg)
	for rep in range(2):
		t = {}
		for stage_name, st in (("cost", 1), ("aggregate", 2), ("wta", 4), ("all", 7)):
			disp, cost = m.MatchDevice(*args, stages=st, costs=costs, accums=accums)
			t[stage_name] = m.stats.ms_device
	same = ""
	if ref is None:
		ref = (accums.clone(), disp.clone())
	else:
		same = "| identical to the register pipeline: %s" % (torch.equal(ref[0], accums) and torch.equal(ref[1], disp))
		if "sgmCost" in dbg:
			same = "| disparities equal to the tensor-core-cost run on %.4f of the pixels" % float((ref[1] == disp).float().mean())
	print("D=%d %-40s" % (D, name), " ".join("%s %.2f ms" % kv for kv in t.items()), "| %.2f G(px.d)/s" % (n/t["all"]/1e6), same, flush=True)
