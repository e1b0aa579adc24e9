// sgm_front.cu — SGM path aggregation as wave fronts (sm_100a), the default for uniform disparity ranges.
//
// What it computes: the eight path recursions of SemiGlobalMatcher::Match and their sum
//   L_r(p,d) = C(p,d) + min(L_r(q,d), L_r(q,d+-1)+P1, min_d' L_r(q,d')+P2) - min_d' L_r(q,d'),  S = sum_r L_r
// (pixelAccum + path drivers, libs/MVS/SemiGlobalMatcher.cpp:1003-1269; exact O(D) form under P1 <= P2 as in sgm_kernels.cu).
//
// Why a new organisation (round-1 evidence, profiles/launches_r01_sgm.txt): one launch per direction read-modify-writes the
// u16 sum volume eight times (11.2 GB of DRAM traffic against 2.9 GB algorithmic) and spends about 170 warp instructions per
// pixel and direction on a 128-wide scanline step.  Here
//   * 8 lanes own one pixel (16 disparities per lane as 8 packed u16x2 words): the step is SIMD-in-a-word arithmetic
//     (VIADD.16x2 / VIMNMX.U16x2 / the DPX three-input minimum VIMNMX3.U16x2), a warp advances 4 adjacent paths, the per-step
//     fixed cost (penalty lookup, neighbour shuffles, minimum reduction, loop) is shared by 4 pixels: about 25 warp
//     instructions per pixel and direction;
//   * directions whose step moves a tilted wave front f = x + 2y forward (right, right-down, down, left-down — and their
//     mirror images in a second pass) are processed TOGETHER, front block by front block: a work item is (direction,
//     band of 4 adjacent paths, block of FB consecutive fronts); items are handed out from one queue in front order, so all
//     four directions touch a block's slice of the sum volume while it is resident in the 126 MB L2 — the volume
//     crosses HBM once per pass instead of once per direction (2 passes: 2 x 263 MB of costs + 526 MB written + 526 MB
//     read-modify-written for 1914 x 1074 x 128, instead of 8 x 1.3 GB);
//   * ordering instead of atomics: within a block the directions are phases; an item waits (rarely — its predecessors are
//     thousands of queue positions ahead) until the previous phase of its block is complete and until its own paths'
//     previous segment has been stored (path state: the normalised previous line, 256 B per path, kept in a small L2-resident
//     buffer between the segments).  Phase 0 of the first pass stores the sum, so the volume needs no memset.
// Bit-exact against the oracle (tests/test_sgm_parity_gpu.py); the per-direction kernels of sgm_kernels.cu remain for
// ragged (tSGM) ranges and as debug variants (b200mvs_debug.sgmAggregation).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

struct SGMPixel { unsigned long long idx; short dmin, dmax; int pad; };
struct SGMParams {
	const float* lgray; const uchar3* lbgr; const float* rgray;
	int w, h, vw, vh;
	const SGMPixel* px;
	uint8_t* costs; uint16_t* accums;
	int P1;
	uint16_t P2s[256];
	int maxNumDisp;
};

#include "sgm_front_sched.h"

namespace {

constexpr int FRONT_WARPS = 4;

__device__ __forceinline__ int ld_acquire(const int* p) {
	int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
	asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldcg4(const void* p) { return __ldcg((const uint4*)p); }
__device__ __forceinline__ void stcg4(void* p, uint4 v) { __stcg((uint4*)p, v); }

template <int NW> struct FrontVec;   // NW packed u16x2 words per lane = 2*NW disparities; cost bytes 2*NW, sum bytes 4*NW
template <> struct FrontVec<8> {
	struct C { uint4 v; }; struct S { uint4 a, b; };
	static __device__ __forceinline__ C ldc(const uint8_t* p) { C c; c.v = __ldg((const uint4*)p); return c; }
	static __device__ __forceinline__ S lds(const uint16_t* p) { S s; s.a = ldcg4(p); s.b = ldcg4(p+8); return s; }
	static __device__ __forceinline__ void sts(uint16_t* p, const unsigned* w) { stcg4(p, make_uint4(w[0], w[1], w[2], w[3])); stcg4(p+8, make_uint4(w[4], w[5], w[6], w[7])); }
	static __device__ __forceinline__ void unpackC(const C& c, unsigned* o) {
		o[0] = __byte_perm(c.v.x, 0u, 0x4140); o[1] = __byte_perm(c.v.x, 0u, 0x4342);
		o[2] = __byte_perm(c.v.y, 0u, 0x4140); o[3] = __byte_perm(c.v.y, 0u, 0x4342);
		o[4] = __byte_perm(c.v.z, 0u, 0x4140); o[5] = __byte_perm(c.v.z, 0u, 0x4342);
		o[6] = __byte_perm(c.v.w, 0u, 0x4140); o[7] = __byte_perm(c.v.w, 0u, 0x4342);
	}
	static __device__ __forceinline__ void unpackS(const S& s, unsigned* o) { o[0] = s.a.x; o[1] = s.a.y; o[2] = s.a.z; o[3] = s.a.w; o[4] = s.b.x; o[5] = s.b.y; o[6] = s.b.z; o[7] = s.b.w; }
};
template <> struct FrontVec<4> {
	struct C { uint2 v; }; struct S { uint4 a; };
	static __device__ __forceinline__ C ldc(const uint8_t* p) { C c; c.v = __ldg((const uint2*)p); return c; }
	static __device__ __forceinline__ S lds(const uint16_t* p) { S s; s.a = ldcg4(p); return s; }
	static __device__ __forceinline__ void sts(uint16_t* p, const unsigned* w) { stcg4(p, make_uint4(w[0], w[1], w[2], w[3])); }
	static __device__ __forceinline__ void unpackC(const C& c, unsigned* o) {
		o[0] = __byte_perm(c.v.x, 0u, 0x4140); o[1] = __byte_perm(c.v.x, 0u, 0x4342);
		o[2] = __byte_perm(c.v.y, 0u, 0x4140); o[3] = __byte_perm(c.v.y, 0u, 0x4342);
	}
	static __device__ __forceinline__ void unpackS(const S& s, unsigned* o) { o[0] = s.a.x; o[1] = s.a.y; o[2] = s.a.z; o[3] = s.a.w; }
};
template <> struct FrontVec<16> {
	struct C { uint4 v, u; }; struct S { uint4 a, b, c, d; };
	static __device__ __forceinline__ C ldc(const uint8_t* p) { C c; c.v = __ldg((const uint4*)p); c.u = __ldg((const uint4*)(p+16)); return c; }
	static __device__ __forceinline__ S lds(const uint16_t* p) { S s; s.a = ldcg4(p); s.b = ldcg4(p+8); s.c = ldcg4(p+16); s.d = ldcg4(p+24); return s; }
	static __device__ __forceinline__ void sts(uint16_t* p, const unsigned* w) {
		stcg4(p, make_uint4(w[0], w[1], w[2], w[3])); stcg4(p+8, make_uint4(w[4], w[5], w[6], w[7]));
		stcg4(p+16, make_uint4(w[8], w[9], w[10], w[11])); stcg4(p+24, make_uint4(w[12], w[13], w[14], w[15]));
	}
	static __device__ __forceinline__ void unpackC(const C& c, unsigned* o) {
		const unsigned s[8] = {c.v.x, c.v.y, c.v.z, c.v.w, c.u.x, c.u.y, c.u.z, c.u.w};
		#pragma unroll
		for (int i = 0; i < 8; ++i) { o[2*i] = __byte_perm(s[i], 0u, 0x4140); o[2*i+1] = __byte_perm(s[i], 0u, 0x4342); }
	}
	static __device__ __forceinline__ void unpackS(const S& s, unsigned* o) {
		const uint4 q[4] = {s.a, s.b, s.c, s.d};
		#pragma unroll
		for (int i = 0; i < 4; ++i) { o[4*i] = q[i].x; o[4*i+1] = q[i].y; o[4*i+2] = q[i].z; o[4*i+3] = q[i].w; }
	}
};

// NW words per lane, 8 lanes per pixel, 4 pixels (adjacent paths) per warp; PD = steps whose loads are in flight.
// Dense volumes only: every pixel of the valid region is valid, owns `num` = 16*NW entries at idx = (y*vw + x)*num (checked by
// the caller).  Per-item overhead is kept off the critical path: the ticket of the item after next and the record of the next item
// are requested while the current item runs; the read-only loads of the first steps (costs, intensities) are issued before the
// dependency wait; the wait polls with relaxed loads (no L1 invalidation per poll) and fences once.
template <int NW, int PD>
__global__ void __launch_bounds__(FRONT_WARPS*32)
sgm_front_kernel(const __grid_constant__ SGMParams P, const __grid_constant__ FrontArgs A)
{
	typedef FrontVec<NW> V;
	__shared__ unsigned sP2[256];   // adaptive P2 replicated in both halfwords (GenerateP2s, SemiGlobalMatcher.cpp:518-524)
	for (int i = threadIdx.x; i < 256; i += blockDim.x) sP2[i] = (unsigned)P.P2s[i]*0x10001u;
	__syncthreads();
	const int lane = threadIdx.x&31;
	const int g = lane>>3, sub = lane&7;
	const unsigned P1x2 = (unsigned)P.P1*0x10001u;
	const int vw = P.vw, vh = P.vh, num = A.num;
	// queue: `ticket` is being processed, `next` is already claimed, the one after is requested at the top of the loop
	int ticket = 0, next = 0;
	if (lane == 0) { ticket = atomicAdd(A.ticket, 1); next = atomicAdd(A.ticket, 1); }
	ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0); next = __shfl_sync(0xFFFFFFFFu, next, 0);
	uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
	if (ticket < A.nItems) { r0 = __ldg((const uint4*)(A.items+ticket)); r1 = __ldg((const uint4*)(A.items+ticket)+1); }
	#pragma unroll 1
	while (ticket < A.nItems) {
		int next2 = 0;
		if (lane == 0) next2 = atomicAdd(A.ticket, 1);
		uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
		if (next < A.nItems) { n0 = __ldg((const uint4*)(A.items+next)); n1 = __ldg((const uint4*)(A.items+next)+1); }
		const int k0 = (int)r0.x, dir = (int)(short)(r0.y&0xFFFFu), ph = (int)(short)(r0.y>>16), fblk = (int)r0.z, seq = (int)r0.w;
		const int chain = (int)r1.x, depCell = (int)r1.y, depNeed = (int)r1.z, cell = (int)r1.w;
		// geometry of this lane's path
		int xs = 0, ys = 0, dx = 0, dy = 0;
		const bool pv = front_path_start(dir, k0+g, vw, vh, xs, ys, dx, dy);
		const int n = pv ? front_path_len(xs, ys, dx, dy, vw, vh) : 0;
		const int f0 = A.fa*xs + A.fb*ys + A.fc, df = max(1, A.fa*dx + A.fb*dy);
		const int s0 = min(n, front_first_step(fblk*A.FB, f0, df)), s1 = min(n, front_first_step((fblk+1)*A.FB, f0, df));
		const int cnt = s1-s0;
		const int maxcnt = __reduce_max_sync(0xFFFFFFFFu, cnt);
		const bool add = !(A.storePhase0 && ph == 0);
		// pipeline stages.  Addresses advance by constant strides along the path: the load pointers run PD steps ahead of the
		// store pointer (no per-step 64-bit multiplies).
		typename V::C cs[PD]; typename V::S ss[PD]; float is[PD];
		const long long pstep = (long long)dy*vw + dx;                    // pixel index stride of one step
		const long long pix0 = (long long)(ys+s0*dy)*vw + (xs+s0*dx);
		const uint8_t* cptr = P.costs + (size_t)pix0*(size_t)num + (size_t)sub*(2*NW);
		const uint16_t* sptr = P.accums + (size_t)pix0*(size_t)num + (size_t)sub*(2*NW);
		uint16_t* optr = P.accums + (size_t)pix0*(size_t)num + (size_t)sub*(2*NW);
		const float* iptr = P.lgray + (size_t)(ys+s0*dy)*P.w + (xs+s0*dx);
		const long long cstep = pstep*num, istep = (long long)dy*P.w + dx;
		// read-only inputs of the first PD steps: no dependency, issued before the wait
		#pragma unroll
		for (int j = 0; j < PD; ++j) {
			memset(&cs[j], 0, sizeof(cs[j])); memset(&ss[j], 0, sizeof(ss[j])); is[j] = 0.f;
			if (j < cnt) { cs[j] = V::ldc(cptr + j*cstep); is[j] = __ldg(iptr + j*istep); }
		}
		// wait for the predecessors: the previous segment of this band, the previous phase of this front block
		if (lane == 0) {
			const int* pp = A.progress+chain; const int* pc = A.cellDone+(depCell >= 0 ? depCell : 0);
			const int need = depCell >= 0 ? depNeed : 0;
			unsigned spins = 0;
			for (;;) {
				int a, c;
				asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(a) : "l"(pp) : "memory");
				asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(c) : "l"(pc) : "memory");
				if (a >= seq && c >= need) break;
				__nanosleep(256);
				if (++spins > (1u<<21)) { *A.error = 1; break; }
			}
			__threadfence();   // acquire: the predecessors' stores are visible to the loads below (which go to the L2)
		}
		__syncwarp();
		// path state
		const size_t slot = (size_t)ph*A.maxPaths + (size_t)(k0+g);
		unsigned w[NW];
		float Ip = 0.5f; bool havePrev = false;
		#pragma unroll
		for (int i = 0; i < NW; ++i) w[i] = 0xFFFFFFFFu;
		if (s0 > 0 && cnt > 0) {
			const float2 m = __ldcg(A.meta+slot);
			Ip = m.x; havePrev = m.y != 0.f;
			const uint16_t* st = A.state + slot*(size_t)num + (size_t)sub*(2*NW);
			#pragma unroll
			for (int i = 0; i < NW; i += 4) { const uint4 v = ldcg4(st+2*i); w[i] = v.x; w[i+1] = v.y; w[i+2] = v.z; w[i+3] = v.w; }
		}
		if (add) {
			#pragma unroll
			for (int j = 0; j < PD; ++j) if (j < cnt) ss[j] = V::lds(sptr + j*cstep);
		}
		cptr += (long long)min(cnt, PD)*cstep; sptr += (long long)min(cnt, PD)*cstep; iptr += (long long)min(cnt, PD)*istep;
		auto load = [&](int j) {   // loads of the next step not yet requested, then advance
			cs[j] = V::ldc(cptr);
			if (add) ss[j] = V::lds(sptr);
			is[j] = __ldg(iptr);
			cptr += cstep; sptr += cstep; iptr += istep;
		};
		#pragma unroll 1
		for (int t = 0; t < maxcnt; t += PD) {
			#pragma unroll
			for (int j = 0; j < PD; ++j) {
				const int tt = t+j;
				if (tt >= maxcnt) break;
				const bool act = tt < cnt;
				unsigned C[NW], S[NW];
				V::unpackC(cs[j], C);
				V::unpackS(ss[j], S);
				const float I = is[j];
				if (tt+PD < cnt) load(j);
				// penalty of this step: P2s[|round(255 (I - Ip))|] (SemiGlobalMatcher.cpp:1009, 518-524)
				const int di = min(255, abs((int)floorf(255.f*(I-Ip)+.5f)));
				const unsigned P2x2 = sP2[di];
				unsigned L[NW];
				// neighbours d-1 / d+1 across the lanes of the pixel (0xFFFF beyond the two ends of the range)
				unsigned below = __shfl_up_sync(0xFFFFFFFFu, w[NW-1]>>16, 1), above = __shfl_down_sync(0xFFFFFFFFu, w[0]&0xFFFFu, 1);
				if (sub == 0) below = 0xFFFFu;
				if (sub == 7) above = 0xFFFFu;
				if (havePrev) {
					unsigned q[NW+1];
					q[0] = __byte_perm(below, w[0], 0x5410);               // (L[-1], L[0])
					#pragma unroll
					for (int i = 1; i < NW; ++i) q[i] = __byte_perm(w[i-1], w[i], 0x5432);   // (L[2i-1], L[2i])
					q[NW] = __byte_perm(w[NW-1], above, 0x5432);
					#pragma unroll
					for (int i = 0; i < NW; ++i) {
						const unsigned nb = __vaddus2(__vminu2(q[i], q[i+1]), P1x2);
						L[i] = __vadd2(C[i], __vimin3_u16x2(w[i], nb, P2x2));
					}
				} else {
					#pragma unroll
					for (int i = 0; i < NW; ++i) L[i] = __vadd2(C[i], P2x2);
				}
				// minimum of the new line over the pixel's 8 lanes
				unsigned m = L[0];
				#pragma unroll
				for (int i = 1; i+1 < NW; i += 2) m = __vimin3_u16x2(m, L[i], L[i+1]);
				if ((NW&1) == 0) m = __vminu2(m, L[NW-1]);
				// over the pixel's 8 lanes: three butterfly steps on the packed pair (a sub-warp redux.sync with a per-group mask
				// compiles to a WARPSYNC.COLLECTIVE loop over the groups: measured at about 2000 cycles per step)
				m = __vminu2(m, __shfl_xor_sync(0xFFFFFFFFu, m, 4));
				m = __vminu2(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
				m = __vminu2(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
				m = min(m&0xFFFFu, m>>16);
				const unsigned mx2 = m*0x10001u;
				if (act) {
					unsigned out[NW];
					#pragma unroll
					for (int i = 0; i < NW; ++i) { out[i] = add ? __vadd2(S[i], L[i]) : L[i]; w[i] = __vsub2(L[i], mx2); }
					V::sts(optr, out);
					optr += cstep;
					Ip = I; havePrev = true;
				}
			}
		}
		// store the state for the next segment of these paths, then publish
		if (cnt > 0 && s1 < n) {
			uint16_t* st = A.state + slot*(size_t)num + (size_t)sub*(2*NW);
			#pragma unroll
			for (int i = 0; i < NW; i += 4) stcg4(st+2*i, make_uint4(w[i], w[i+1], w[i+2], w[i+3]));
			if (sub == 0) __stcg(A.meta+slot, make_float2(Ip, havePrev ? 1.f : 0.f));
		}
		__threadfence();
		__syncwarp();
		if (lane == 0) {
			st_release(A.progress+chain, seq+1);
			atomicAdd(A.cellDone+cell, 1);
		}
		ticket = next; next = __shfl_sync(0xFFFFFFFFu, next2, 0);
		r0 = n0; r1 = n1;
	}
}

} // namespace

// pd: software pipeline depth (4 or 6 steps in flight)
cudaError_t sgm_front_launch(const SGMParams& P, const FrontArgs& A, int blocks, int pd, cudaStream_t s) {
	const int NW = A.num/16;
	if (NW == 8) { if (pd == 6) sgm_front_kernel<8, 6><<<blocks, FRONT_WARPS*32, 0, s>>>(P, A); else sgm_front_kernel<8, 4><<<blocks, FRONT_WARPS*32, 0, s>>>(P, A); }
	else if (NW == 4) { if (pd == 6) sgm_front_kernel<4, 6><<<blocks, FRONT_WARPS*32, 0, s>>>(P, A); else sgm_front_kernel<4, 4><<<blocks, FRONT_WARPS*32, 0, s>>>(P, A); }
	else if (NW == 16) sgm_front_kernel<16, 4><<<blocks, FRONT_WARPS*32, 0, s>>>(P, A);
	else return cudaErrorInvalidValue;
	return cudaGetLastError();
}
// resident CTAs of the kernel per SM on the current device
int sgm_front_blocks_per_sm(int num, int pd) {
	int per = 1;
	const int NW = num/16;
	if (NW == 8) { if (pd == 6) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, sgm_front_kernel<8, 6>, FRONT_WARPS*32, 0); else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, sgm_front_kernel<8, 4>, FRONT_WARPS*32, 0); }
	else if (NW == 4) { if (pd == 6) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, sgm_front_kernel<4, 6>, FRONT_WARPS*32, 0); else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, sgm_front_kernel<4, 4>, FRONT_WARPS*32, 0); }
	else if (NW == 16) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, sgm_front_kernel<16, 4>, FRONT_WARPS*32, 0);
	return std::max(1, per);
}
bool sgm_front_supports(int num) { return num == 64 || num == 128 || num == 256; }
