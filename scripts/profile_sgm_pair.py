"""1080p SGM pair (right->left match, left->right match, cross-check, sub-pixel refinement): one context after the other
against two contexts on two streams.  usage: profile_sgm_pair.py [D]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openmvs_b200 import synth
from openmvs_b200.depth_estimator import SemiGlobalMatcher
D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w, h = 1920, 1080
lg, lc, rg, d, rc = synth.make_stereo_pair(w, h, d0=40.0, amp=25.0, right_color=True)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
L, LC, R, RC = dev(lg), dev(lc), dev(rg), dev(rc)
m = SemiGlobalMatcher()
res = {}
for overlap in (False, True):
	for rep in range(4):
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		torch.cuda.synchronize()
		e0.record()
		ld, rd = m.MatchPairDevice(L, LC, R, RC, -D, 0, overlap=overlap)
		e1.record(); torch.cuda.synchronize()
		ms = e0.elapsed_time(e1)
	res[overlap] = (ld.clone(), rd.clone())
	print("D=%d pair (2 matches + cross-check + refinement), %s: %.2f ms | %.1f G(px.d)/s" % (D, "two contexts on two streams" if overlap else "one context", ms, 2*(w-6)*(h-6)*D/ms/1e6), flush=True)
print("identical results:", torch.equal(res[False][0], res[True][0]) and torch.equal(res[False][1], res[True][1]))
