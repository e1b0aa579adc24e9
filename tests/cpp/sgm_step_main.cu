// Host check of openmvs_b200/csrc/sgm_step.cuh: the packed u16x2 scanline step against the scalar form the kernels use
// (sgm_kernels.cu: min(min(Lp[j], min(lm, lq) + P1), minLp + P2), integer arithmetic with 0xFFFF sentinels).
// Built with nvcc, runs on the CPU (the functions are __host__ __device__).
#include "../../openmvs_b200/csrc/sgm_step.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

static unsigned rnd(unsigned& s) { s = s*1664525u+1013904223u; return s>>8; }

int main() {
	unsigned seed = 12345u;
	long long checked = 0;
	for (int trial = 0; trial < 2000; ++trial) {
		// a whole "warp" of 32 lanes x 4 disparities walking a scanline of 40 pixels, both ways
		const unsigned P1 = rnd(seed)%4, P2base = 4+rnd(seed)%57;
		unsigned Lp[128]; unsigned minLp = 0xFFFFu; bool havePrev = false;
		SgmLane4 st[32]; uint2 acc[32]; unsigned accRef[128];
		for (int l = 0; l < 32; ++l) { st[l].PA = st[l].PB = 0xFFFFFFFFu; acc[l].x = rnd(seed)&0x0FFF0FFFu; acc[l].y = rnd(seed)&0x0FFF0FFFu; }
		for (int d = 0; d < 128; ++d) { Lp[d] = 0xFFFFu; accRef[d] = d&1 ? ((d&2 ? acc[d/4].y : acc[d/4].x)>>16) : ((d&2 ? acc[d/4].y : acc[d/4].x)&0xFFFFu); }
		for (int step = 0; step < 40; ++step) {
			const unsigned P2 = P2base > 4 ? 4+rnd(seed)%(P2base-3) : 4;
			unsigned char cost[128];
			for (int d = 0; d < 128; ++d) cost[d] = (unsigned char)(trial%3 == 0 ? 255 : rnd(seed)&0xFF);
			// scalar form
			unsigned Ln[128], mn = 0xFFFFu;
			for (int d = 0; d < 128; ++d) {
				if (!havePrev) Ln[d] = cost[d]+P2;
				else {
					const int lm = d > 0 ? (int)Lp[d-1] : 0xFFFF, lq = d < 127 ? (int)Lp[d+1] : 0xFFFF;
					int best = (int)Lp[d]; const int n = (lm < lq ? lm : lq)+(int)P1; if (n < best) best = n;
					const int far = (int)minLp+(int)P2; if (far < best) best = far;
					Ln[d] = (unsigned)((int)cost[d]+best-(int)minLp);
				}
				accRef[d] = (accRef[d]+Ln[d])&0xFFFFu;
				if (Ln[d] < mn) mn = Ln[d];
			}
			// packed form, lane by lane (shuffles replaced by reads of the neighbour lanes' previous state)
			SgmLane4 prev[32]; memcpy(prev, st, sizeof(st));
			unsigned mnP = 0xFFFFu;
			for (int l = 0; l < 32; ++l) {
				unsigned cw; memcpy(&cw, cost+4*l, 4);
				const unsigned below = l > 0 ? prev[l-1].PB>>16 : 0xFFFFu, above = l < 31 ? prev[l+1].PA&0xFFFFu : 0xFFFFu;
				const unsigned m = sgm_step_packed4(cw, below, above, P1|(P1<<16), P2|(P2<<16), minLp|(minLp<<16), havePrev, st[l], acc[l]);
				if (m < mnP) mnP = m;
			}
			for (int l = 0; l < 32; ++l) {
				const unsigned got[4] = {st[l].PA&0xFFFFu, st[l].PA>>16, st[l].PB&0xFFFFu, st[l].PB>>16};
				const unsigned ga[4] = {acc[l].x&0xFFFFu, acc[l].x>>16, acc[l].y&0xFFFFu, acc[l].y>>16};
				for (int j = 0; j < 4; ++j) {
					if (got[j] != Ln[4*l+j] || ga[j] != accRef[4*l+j]) {
						printf("MISMATCH trial %d step %d d %d: L %u vs %u, acc %u vs %u\n", trial, step, 4*l+j, got[j], Ln[4*l+j], ga[j], accRef[4*l+j]);
						return 1;
					}
					++checked;
				}
			}
			if (mnP != mn) { printf("MISMATCH minimum trial %d step %d: %u vs %u\n", trial, step, mnP, mn); return 1; }
			memcpy(Lp, Ln, sizeof(Ln)); minLp = mn; havePrev = true;
		}
	}
	printf("sgm_step_packed4 == scalar step on %lld values\n", checked);
	return 0;
}
