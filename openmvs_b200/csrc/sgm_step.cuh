// sgm_step.cuh — one scanline step of the SGM path recursion on packed u16x2 values.
//
// L(d) = C(d) + min(Lp(d), min(Lp(d-1), Lp(d+1)) + P1, minLp + P2) - minLp   (pixelAccum,
// libs/MVS/SemiGlobalMatcher.cpp:1003-1046, evaluated in O(D) as in sgm_kernels.cu) for a lane that owns four
// consecutive, all valid disparities: two 32-bit registers hold (L0, L1) and (L2, L3) as unsigned halfwords and
// every operation handles two disparities — SIMD-in-a-word adds/mins and the DPX three-input minimum, which are
// single instructions on sm_90+ / sm_100.  Path costs stay below 2^16 by construction (L <= 255 + P2), the
// 0xFFFF "no neighbour" sentinels of the first and last lane survive the saturating add.
// The functions are __host__ __device__: tests/cpp/sgm_step_main.cu checks them on the CPU against the scalar
// form used by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define SGM_VADD2(a, b)   __vadd2(a, b)
#define SGM_VSUB2(a, b)   __vsub2(a, b)
#define SGM_VADDUS2(a, b) __vaddus2(a, b)
#define SGM_VMINU2(a, b)  __vminu2(a, b)
#else
// host stand-ins of the device-only SIMD-in-a-word intrinsics (per unsigned halfword)
static inline unsigned sgm_h2(unsigned a, unsigned b, unsigned (*f)(unsigned, unsigned)) {
	return (f(a&0xFFFFu, b&0xFFFFu)&0xFFFFu) | (f(a>>16, b>>16)<<16);
}
static inline unsigned sgm_fadd(unsigned a, unsigned b) { return (a+b)&0xFFFFu; }
static inline unsigned sgm_fsub(unsigned a, unsigned b) { return (a-b)&0xFFFFu; }
static inline unsigned sgm_fadds(unsigned a, unsigned b) { return a+b > 0xFFFFu ? 0xFFFFu : a+b; }
static inline unsigned sgm_fmin(unsigned a, unsigned b) { return a < b ? a : b; }
#define SGM_VADD2(a, b)   sgm_h2(a, b, sgm_fadd)
#define SGM_VSUB2(a, b)   sgm_h2(a, b, sgm_fsub)
#define SGM_VADDUS2(a, b) sgm_h2(a, b, sgm_fadds)
#define SGM_VMINU2(a, b)  sgm_h2(a, b, sgm_fmin)
#endif

// state of the previous pixel of the scanline for this lane: PA = (Lp0 | Lp1 << 16), PB = (Lp2 | Lp3 << 16)
struct SgmLane4 { unsigned PA, PB; };

// cw: the lane's four cost bytes (little endian: C0 in the low byte); below / above: Lp3 of the lane below and
// Lp0 of the lane above (0xFFFF where there is none); P1x2 / P2x2: the penalties replicated in both halfwords;
// minx2: the minimum of the previous line replicated.  Returns the lane minimum of the new values; acc (the four
// u16 accumulators as two words) += L.
__host__ __device__ __forceinline__ unsigned sgm_step_packed4(unsigned cw, unsigned below, unsigned above, unsigned P1x2, unsigned P2x2,
	unsigned minx2, bool havePrev, SgmLane4& s, uint2& acc)
{
	// costs as halfwords: (C0, C1) and (C2, C3)
#if defined(__CUDA_ARCH__)
	const unsigned CA = __byte_perm(cw, 0u, 0x4140), CB = __byte_perm(cw, 0u, 0x4342);
#else
	const unsigned CA = (cw&0xFFu) | ((cw&0xFF00u)<<8), CB = ((cw>>16)&0xFFu) | ((cw>>24)<<16);
#endif
	unsigned LA, LB;
	if (!havePrev) {
		LA = SGM_VADD2(CA, P2x2); LB = SGM_VADD2(CB, P2x2);
	} else {
		const unsigned far = SGM_VADD2(minx2, P2x2);
		// neighbours d-1 / d+1: (below, L0) (L1, L2) | (L1, L2) (L3, above)
		const unsigned mid = (s.PA>>16) | (s.PB<<16);
		const unsigned lmA = (s.PA<<16) | (below&0xFFFFu);
		const unsigned lqB = (s.PB>>16) | (above<<16);
		const unsigned nA = SGM_VADDUS2(SGM_VMINU2(lmA, mid), P1x2);
		const unsigned nB = SGM_VADDUS2(SGM_VMINU2(mid, lqB), P1x2);
		const unsigned bA = __vimin3_u16x2(s.PA, nA, far);
		const unsigned bB = __vimin3_u16x2(s.PB, nB, far);
		LA = SGM_VADD2(CA, SGM_VSUB2(bA, minx2));
		LB = SGM_VADD2(CB, SGM_VSUB2(bB, minx2));
	}
	acc.x = SGM_VADD2(acc.x, LA); acc.y = SGM_VADD2(acc.y, LB);
	s.PA = LA; s.PB = LB;
	const unsigned m = SGM_VMINU2(LA, LB);
	return (m&0xFFFFu) < (m>>16) ? (m&0xFFFFu) : (m>>16);
}
