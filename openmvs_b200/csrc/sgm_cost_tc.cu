// sgm_cost_tc.cu — the WZNCC 7x7 cost volume of the SGM pair matcher on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// What it computes: per valid pixel x of a row and disparity d, the three weighted sums of the 49-tap window
//   sum = S_n w_x(n) f_n,  sumSq = S_n w_x(n) f_n^2,  nom = S_n tw_x(n) f_n,   f_n = right(y+i, x+d+j), n = (i,j)
// and from them the uint8 cost (SemiGlobalMatcher.cpp:875-985; the SIMT form is sgm_cost_kernel in sgm_kernels.cu).
//
// Why tensor cores fit: for a block of 128 pixels of one row the sums are a banded GEMM.  With A[x, n] = w_x(n) (M = 128 pixels,
// K = 49 taps padded to 64) and the Toeplitz matrix B[n, u] = right(y+i, u+j) (u = x + d, the right-image column), the wanted
// entries are D[x, u] for u - x in [dmin, dmin+D).  The band is cut into 64-column tiles of u; a tile serves two pixel blocks, a
// block needs (128+D)/64 tiles (band utilisation 50 % at D = 128).  Precision: the operands are split into fp16 high and low
// parts (w = wh + wl, error 2^-22) and every sum is three kind::f16 MMAs with fp32 accumulation in TMEM (wh fh + wh fl + wl fh) —
// the dropped wl fl term is 2^-22 relative — so the uint8 cost stays within the +-1 level of the SIMT kernel.
//
// Structure (one CTA of 16 worker warps + 1 issuing warp per SM, persistent over the rows of the valid region):
//   the workers stage the seven image rows of the block in shared memory (left colour packed to a word, left and right gray) and
//              build A (bilateral weights of the block's 128 pixels: one warp-uniform tap quarter per warp, so tap offsets and
//              spatial weights are immediates; colour distance by VABSDIFF4 + DP4A, weight by one EX2; fp16 split; 16-byte stores
//              in the UMMA K-major no-swizzle core-matrix layout) and the two new B tiles (im2col of the staged right rows);
//   the issuing warp waits on a named barrier for the operands and issues the MMAs of a tile (9 products x 4 K-steps of
//              128 x 64 x 16) into one of two TMEM accumulator buffers, committing them to an mbarrier; it re-uses a buffer as
//              soon as the workers have signalled (bar.arrive, non-blocking) that they have read it;
//   the workers read a finished buffer (tcgen05.ld 32x32b), turn sums into costs and scatter them into a shared cost tile,
//              which is finally written to the volume with coalesced 16-byte stores.  No CTA-wide barrier inside a block's
//              tile loop: the uneven share of the diagonal band per warp and tile evens out over the block.
// SASS: UTCHMMA (tcgen05.mma), UTCBAR (commit), LDTM (tcgen05.ld), UTCATOMSWS / UTCALLOC (alloc).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

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

namespace {

constexpr int HW = 3, NT = 49;
constexpr int BM = 128;          // pixels per block (MMA M)
constexpr int BN = 64;           // right-image columns per tile (MMA N)
constexpr int KP = 64;           // taps padded to the MMA K granularity (4 x 16)
constexpr int TC_WORKERS = 512;  // operand builders and epilogue: 16 warps
constexpr int TC_THREADS = TC_WORKERS+32;   // + the warp that issues the MMAs
constexpr int A_ARRAY = BM*KP*2;         // one fp16 operand array of a block: 16 KB
constexpr int B_ARRAY = BN*KP*2;         // one fp16 operand array of a tile: 8 KB
constexpr int B_SLOT = 4*B_ARRAY;        // fh, fl, qh, ql
constexpr int RING = 4;                  // B tiles kept (a block of D <= 128 needs (128+D)/64 <= 4)
constexpr int TILE_PITCH = 132;          // bytes per pixel row of the shared cost tile (bank-conflict-free byte scatter)
constexpr int SP = BM+8;                 // pitch of the staged image rows (128 columns + 6 of the window, padded)
constexpr int SMEM_A = 4*A_ARRAY;        // wh, wl, th, tl
constexpr int SMEM_B = RING*B_SLOT;
constexpr int SMEM_TILE = BM*TILE_PITCH; // the block's costs; while the operands are built: the staged image rows and partial sums
constexpr int SMEM_CONST = 4*BM*4 + BM*4;   // normSq0 partial of each tap quarter, 1/sumW per pixel
constexpr int SMEM_TOTAL = SMEM_A + SMEM_B + SMEM_TILE + SMEM_CONST + 64;
// staging area inside the cost tile (free between the store of a block and the epilogue of the next)
constexpr int ST_LC = 0;                 // left colour rows, packed B | G<<8 | R<<16: 7 x SP u32
constexpr int ST_LG = ST_LC + 7*SP*4;    // left gray rows: 7 x SP float
constexpr int ST_RG = ST_LG + 7*SP*4;    // right gray rows of the two new tiles: 7 x SP float
constexpr int ST_PART = ST_RG + 7*SP*4;  // {sum w g, sum w} of each tap quarter: 4 x BM float2
static_assert(ST_PART + 4*BM*8 <= SMEM_TILE, "staging area exceeds the cost tile");
static_assert(SMEM_TOTAL <= 232448, "shared memory of one SM");
// named barriers (id 0 is __syncthreads)
constexpr int BAR_WORKERS = 1, BAR_OPS = 2, BAR_DRAIN = 3;   // BAR_DRAIN, BAR_DRAIN+1: accumulator buffers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

// K-major, no swizzle (INTERLEAVE): a K-chunk of 8 halves (16 B) of row r sits at chunk*rows*16 + r*16; 8 consecutive rows are one
// 128-byte core matrix: stride between 8-row groups SBO = 128 B, between the two 16-byte K-chunks of one MMA LBO = rows*16 B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr>>4) & 0x3FFFu);
	d |= (uint64_t)((lbo_bytes>>4) & 0x3FFFu) << 16;
	d |= (uint64_t)((128u>>4) & 0x3FFFu) << 32;
	d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
	return d;                 // layout type 0 = no swizzle, base offset 0
}
// instruction descriptor of kind::f16: D = F32, A = B = F16, both K-major, M = 128, N = 64
constexpr uint32_t IDESC = (1u<<4) | ((uint32_t)(BN>>3)<<17) | ((uint32_t)(BM>>4)<<24);

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
		:: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded: a tensor-core batch takes microseconds; a wait of seconds means a malformed descriptor — trap instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
	unsigned ok = 0;
	for (unsigned spins = 0; !ok; ++spins) {
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
			: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (!ok && spins > (1u<<26)) asm volatile("trap;");
	}
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		: "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b)<<16); }
// x = hi + lo with hi = fp16(x): two fp16 numbers carrying 22 bits of x
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) { hi = __float2half_rn(x); lo = __float2half_rn(x-__half2float(hi)); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr float LOG2E = 1.4426950408889634f;
constexpr float SIGMA_COLOR = -1.f/(2.f*(0.3f*255)*(0.3f*255));
constexpr float SIGMA_SPATIAL = -1.f/(2.f*(0.4f*7)*(0.4f*7));

// One tap quarter (taps 16 Q .. 16 Q + 15; Q = 3: tap 48) of the A operands of pixel `row` of the block, from the staged left
// rows.  Q is a template parameter so that the tap geometry (i, j) and the spatial weight are compile-time constants.
// Bilateral weight (SemiGlobalMatcher.cpp:889-905): exp(colour distance^2 sigmaColor + spatial distance^2 sigmaSpatial) as one
// ex2 of a fused multiply-add; the colour distance is |dB|^2 + |dG|^2 + |dR|^2 = dp4a of the packed absolute differences.
template <int Q>
__device__ __forceinline__ void build_a_quarter(unsigned char* sA, const unsigned char* sStage, float* sNorm, float* sInv, int row, bool valid) {
	constexpr int N0 = 16*Q, NN = Q == 3 ? 1 : 16;
	const uint32_t* lc = (const uint32_t*)(sStage + ST_LC);
	const float* lg = (const float*)(sStage + ST_LG);
	float2* part = (float2*)(sStage + ST_PART);
	float wv[NN], gv[NN];
	float sumW = 0.f, acc = 0.f;
	const uint32_t cc = lc[HW*SP + row+HW];
	#pragma unroll
	for (int k = 0; k < NN; ++k) {
		const int n = N0+k, i = n/7, j = n-7*i;   // compile-time after unrolling
		const uint32_t d = __vabsdiffu4(lc[i*SP + row+j], cc);
		const int dist2 = (int)__dp4a(d, d, 0u);
		const float spatial = float((j-HW)*(j-HW)+(i-HW)*(i-HW))*(SIGMA_SPATIAL*LOG2E);
		float wgt = ex2_approx(fmaf((float)dist2, SIGMA_COLOR*LOG2E, spatial));
		if (!valid) wgt = 0.f;
		const float g = lg[i*SP + row+j];
		wv[k] = wgt; gv[k] = g;
		acc = fmaf(g, wgt, acc);
		sumW += wgt;
	}
	part[Q*BM + row] = make_float2(acc, sumW);
	bar_sync(BAR_WORKERS, TC_WORKERS);
	{
		const float2 p0 = part[row], p1 = part[BM+row], p2 = part[2*BM+row], p3 = part[3*BM+row];
		acc = (p0.x+p1.x)+(p2.x+p3.x); sumW = (p0.y+p1.y)+(p2.y+p3.y);
	}
	if (!valid) sumW = 1.f;
	const float tm = acc/sumW;
	float normSq0 = 0.f;
	#pragma unroll
	for (int k = 0; k < NN; ++k) {
		const float t = gv[k]-tm;
		gv[k] = wv[k]*t;          // tempWeight
		normSq0 = fmaf(gv[k], t, normSq0);
	}
	sNorm[Q*BM + row] = normSq0;
	if (Q == 0) sInv[row] = 1.f/sumW;
	#pragma unroll
	for (int c2 = 0; c2 < (Q == 3 ? 1 : 2); ++c2) {
		const int kc = 2*Q+c2;
		__half wh[8], wl[8], th[8], tl[8];
		#pragma unroll
		for (int e = 0; e < 8; ++e) {
			const int k = c2*8+e;
			if (k < NN) { split_h(wv[k], wh[e], wl[e]); split_h(gv[k], th[e], tl[e]); }
			else { wh[e] = wl[e] = th[e] = tl[e] = __float2half_rn(0.f); }
		}
		const size_t off = (size_t)kc*(BM*16) + (size_t)row*16;
		*(uint4*)(sA+0*A_ARRAY+off) = make_uint4(pack_h2(wh[0], wh[1]), pack_h2(wh[2], wh[3]), pack_h2(wh[4], wh[5]), pack_h2(wh[6], wh[7]));
		*(uint4*)(sA+1*A_ARRAY+off) = make_uint4(pack_h2(wl[0], wl[1]), pack_h2(wl[2], wl[3]), pack_h2(wl[4], wl[5]), pack_h2(wl[6], wl[7]));
		*(uint4*)(sA+2*A_ARRAY+off) = make_uint4(pack_h2(th[0], th[1]), pack_h2(th[2], th[3]), pack_h2(th[4], th[5]), pack_h2(th[6], th[7]));
		*(uint4*)(sA+3*A_ARRAY+off) = make_uint4(pack_h2(tl[0], tl[1]), pack_h2(tl[2], tl[3]), pack_h2(tl[4], tl[5]), pack_h2(tl[6], tl[7]));
	}
}
// One tap quarter of column c of a B tile (slot) from the staged right rows; cs = column of the tile's first window in the stage
template <int Q>
__device__ __forceinline__ void build_b_quarter(unsigned char* slot, const unsigned char* sStage, int cs, int c) {
	const float* rg = (const float*)(sStage + ST_RG);
	#pragma unroll
	for (int c2 = 0; c2 < (Q == 3 ? 1 : 2); ++c2) {
		const int kc = 2*Q+c2;
		__half fh[8], fl[8], qh[8], ql[8];
		#pragma unroll
		for (int e = 0; e < 8; ++e) {
			const int n = kc*8+e;
			float f = 0.f;
			if (n < NT) { const int i = n/7, j = n-7*i; f = rg[i*SP + cs+c+j]; }
			split_h(f, fh[e], fl[e]);
			split_h(f*f, qh[e], ql[e]);
		}
		const size_t off = (size_t)kc*(BN*16) + (size_t)c*16;
		*(uint4*)(slot+0*B_ARRAY+off) = make_uint4(pack_h2(fh[0], fh[1]), pack_h2(fh[2], fh[3]), pack_h2(fh[4], fh[5]), pack_h2(fh[6], fh[7]));
		*(uint4*)(slot+1*B_ARRAY+off) = make_uint4(pack_h2(fl[0], fl[1]), pack_h2(fl[2], fl[3]), pack_h2(fl[4], fl[5]), pack_h2(fl[6], fl[7]));
		*(uint4*)(slot+2*B_ARRAY+off) = make_uint4(pack_h2(qh[0], qh[1]), pack_h2(qh[2], qh[3]), pack_h2(qh[4], qh[5]), pack_h2(qh[6], qh[7]));
		*(uint4*)(slot+3*B_ARRAY+off) = make_uint4(pack_h2(ql[0], ql[1]), pack_h2(ql[2], ql[3]), pack_h2(ql[4], ql[5]), pack_h2(ql[6], ql[7]));
	}
}

// Dense volumes only (every pixel valid with one range, idx = pixel index x numAll).  One launch computes the disparities
// [dmin, dmin+num), num in {64, 128}, and stores them at offset dOff of every pixel's numAll-wide slice: wider ranges (192, 256)
// are covered by two launches.
__global__ void __launch_bounds__(TC_THREADS, 1)
sgm_cost_tc_kernel(const __grid_constant__ SGMParams P, int dmin, int num, int numAll, int dOff)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	unsigned char* sA = smem;
	unsigned char* sB = smem + SMEM_A;
	unsigned char* sTile = sB + SMEM_B;
	float* sNorm = (float*)(sTile + SMEM_TILE);           // 4 x BM: normSq0 of each tap quarter
	float* sInv = sNorm + 4*BM;                           // 1/sumW
	uint64_t* bars = (uint64_t*)((unsigned char*)sNorm + SMEM_CONST);   // 2 mbarriers
	uint32_t* sTmem = (uint32_t*)(bars+2);
	const int tid = threadIdx.x, warp = tid>>5, lane = tid&31;
	const int w = P.w, vw = P.vw, vh = P.vh;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sTmem)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 32) {
		mbar_init(bars, 1); mbar_init(bars+1, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	// the K padding (taps 49..63) of every operand array is zero and stays zero: only chunks 0..6 are rewritten (chunk 6 holds
	// taps 48..55, its upper seven halves are written as zeros by the builders)
	for (int i = tid; i < (SMEM_A+SMEM_B)/16; i += TC_THREADS) ((uint4*)smem)[i] = make_uint4(0u, 0u, 0u, 0u);
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *sTmem;
	const int nTiles = (BM+num)/BN;             // u tiles a block needs: 3 (D = 64) or 4 (D = 128)
	const int nBlocks = (vw+BM-1)/BM;

	if (warp == TC_WORKERS/32) {
		// ---- the issuing warp: MMAs of a tile as soon as its operands stand and its accumulator buffer has been read ----
		// the 36 MMAs of tile t into accumulator buffer `buf` (columns buf*256 + {0, 64, 128}: sum, sumSq, nom), then commit
		auto issue_tile = [&](int t, int buf) {
			const uint32_t aBase = smem_u32(sA), bBase = smem_u32(sB + (size_t)(t&(RING-1))*B_SLOT);
			const uint32_t dBase = tmem + (uint32_t)buf*256u;
			// {A array, B array, accumulator}: wh fh, wh fl, wl fh -> sum | wh qh, wh ql, wl qh -> sumSq | th fh, th fl, tl fh -> nom
			const int prod[9][3] = {{0, 0, 0}, {0, 1, 0}, {1, 0, 0}, {0, 2, 1}, {0, 3, 1}, {1, 2, 1}, {2, 0, 2}, {2, 1, 2}, {3, 0, 2}};
			#pragma unroll
			for (int p = 0; p < 9; ++p) {
				#pragma unroll
				for (int kk = 0; kk < KP/16; ++kk) {
					const uint64_t ad = umma_desc(aBase + prod[p][0]*A_ARRAY + kk*2*(BM*16), BM*16);
					const uint64_t bd = umma_desc(bBase + prod[p][1]*B_ARRAY + kk*2*(BN*16), BN*16);
					const bool first = (p == 0 || p == 3 || p == 6) && kk == 0;
					umma_f16(dBase + (uint32_t)prod[p][2]*64u, ad, bd, first ? 0u : 1u);
				}
			}
			umma_commit(bars+buf);
		};
		#pragma unroll 1
		for (int r = blockIdx.x; r < vh; r += gridDim.x) {
			#pragma unroll 1
			for (int b = 0; b < nBlocks; ++b) {
				bar_sync(BAR_OPS, TC_THREADS);                 // the workers have built this block's operands
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (lane == 0) { issue_tile(2*b, 0); if (nTiles > 1) issue_tile(2*b+1, 1); }
				__syncwarp();
				for (int k = 2; k < nTiles; ++k) {
					bar_sync(BAR_DRAIN+(k&1), TC_THREADS);     // the workers have read tile k-2 out of this buffer
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					if (lane == 0) issue_tile(2*b+k, k&1);
					__syncwarp();
				}
			}
		}
	} else {
		// ---- the 16 worker warps: operands, then sums -> costs ----
		unsigned phase[2] = {0u, 0u};
		const int wq = warp&3;                       // tap quarter of this warp (operand builders)
		// The staged pixels of a block's first tile pair are loaded into registers while the block before it is in its epilogue:
		// thread t owns elements t and t + 512 of the 7 x 134 window (left colour packed to a word, left gray, right gray).
		struct Pre { uint32_t c[2]; float lg[2], rg[2]; } pre;
		auto prefetch = [&](int rr, int bb) {
			const int xx0 = BM*bb, tt0 = bb == 0 ? 0 : 2*bb+nTiles-2;
			#pragma unroll
			for (int k = 0; k < 2; ++k) {
				const int i = tid + k*TC_WORKERS;
				pre.c[k] = 0u; pre.lg[k] = 0.f; pre.rg[k] = 0.f;
				if (i < 7*(BM+6)) {
					const int ry = i/(BM+6), cx = i-ry*(BM+6);
					const int col = min(xx0+cx, w-1);
					const uchar3 c3 = P.lbgr[(size_t)(rr+ry)*w + col];
					pre.c[k] = (uint32_t)c3.x | ((uint32_t)c3.y<<8) | ((uint32_t)c3.z<<16);
					pre.lg[k] = __ldg(P.lgray + (size_t)(rr+ry)*w + col);
					const int rc = min(max(BN*tt0 + cx + dmin, 0), w-1);
					pre.rg[k] = __ldg(P.rgray + (size_t)(rr+ry)*w + rc);
				}
			}
		};
		if ((int)blockIdx.x < vh) prefetch((int)blockIdx.x, 0);
		#pragma unroll 1
		for (int r = blockIdx.x; r < vh; r += gridDim.x) {
			#pragma unroll 1
			for (int b = 0; b < nBlocks; ++b) {
				const int x0 = BM*b;
				// B tiles not yet in the ring: the last two of the block (all of them for the first block of a row), two at a time
				for (int t0 = (b == 0 ? 0 : 2*b+nTiles-2); t0 < 2*b+nTiles; t0 += 2) {
					if (t0 != (b == 0 ? 0 : 2*b+nTiles-2)) bar_sync(BAR_WORKERS, TC_WORKERS);   // the stage is read by the previous pair
					// stage the image rows r .. r+6: left colour / gray columns x0 .. x0+133 (first pair only), right gray columns of the pair.
					// The first pair of a block comes out of registers: its global loads were issued a block earlier (see below).
					const bool first = t0 == (b == 0 ? 0 : 2*b+nTiles-2);
					if (first) {
						#pragma unroll
						for (int k = 0; k < 2; ++k) {
							const int i = tid + k*TC_WORKERS;
							if (i < 7*(BM+6)) {
								const int ry = i/(BM+6), cx = i-ry*(BM+6);
								((uint32_t*)(sTile+ST_LC))[ry*SP+cx] = pre.c[k];
								((float*)(sTile+ST_LG))[ry*SP+cx] = pre.lg[k];
								((float*)(sTile+ST_RG))[ry*SP+cx] = pre.rg[k];
							}
						}
					} else {
						for (int i = tid; i < 7*(BM+6); i += TC_WORKERS) {
							const int ry = i/(BM+6), cx = i-ry*(BM+6);
							const int rc = min(max(BN*t0 + cx + dmin, 0), w-1);
							((float*)(sTile+ST_RG))[ry*SP+cx] = __ldg(P.rgray + (size_t)(r+ry)*w + rc);
						}
					}
					bar_sync(BAR_WORKERS, TC_WORKERS);
					if (first) {
						const int row = (warp>>2)*32 + lane;
						const bool valid = x0+row < vw;
						switch (wq) {
						case 0: build_a_quarter<0>(sA, sTile, sNorm, sInv, row, valid); break;
						case 1: build_a_quarter<1>(sA, sTile, sNorm, sInv, row, valid); break;
						case 2: build_a_quarter<2>(sA, sTile, sNorm, sInv, row, valid); break;
						default: build_a_quarter<3>(sA, sTile, sNorm, sInv, row, valid); break;
						}
					}
					{
						const int tsel = warp>>3, q = (warp>>1)&3, c = (warp&1)*32 + lane;
						if (t0+tsel < 2*b+nTiles) {
							unsigned char* slot = sB + (size_t)((t0+tsel)&(RING-1))*B_SLOT;
							switch (q) {
							case 0: build_b_quarter<0>(slot, sTile, BN*tsel, c); break;
							case 1: build_b_quarter<1>(slot, sTile, BN*tsel, c); break;
							case 2: build_b_quarter<2>(slot, sTile, BN*tsel, c); break;
							default: build_b_quarter<3>(slot, sTile, BN*tsel, c); break;
							}
						}
					}
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
				bar_sync(BAR_WORKERS, TC_WORKERS);     // every worker is done with the stage (the cost tile is written next)
				bar_arrive(BAR_OPS, TC_THREADS);
				// the next block's window: requested now, consumed after this block's epilogue
				if (b+1 < nBlocks) prefetch(r, b+1);
				else if (r+(int)gridDim.x < vh) prefetch(r+(int)gridDim.x, 0);
				// sums -> costs, tile by tile, into the shared cost tile
				const int row0 = 32*(warp&3), row = row0 + lane;   // TMEM lane = pixel of the block
				const int c0 = 16*(warp>>2);                        // this warp's 16 of the 64 columns
				const int col = x0 + row;                           // valid-region column of the pixel
				// disparities whose right window lies inside the image: 0 <= col + d + dmin, col + d + dmin + 6 < w
				const int dlo = max(0, -(col+dmin)), dhi = min(num, w-2*HW-col-dmin);
				float normSq0 = 0.f, invW = 0.f;
				for (int k = 0; k < nTiles; ++k) {
					const int buf = k&1;
					mbar_wait(bars+buf, phase[buf]); phase[buf] ^= 1u;
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					if (k == 0) { normSq0 = (sNorm[row]+sNorm[BM+row])+(sNorm[2*BM+row]+sNorm[3*BM+row]); invW = sInv[row]; }
					const int kq = BN*k;                             // disparity index of (row 0, column 0) of this tile
					// the band 0 <= d < num covers about half of a tile: a 32-row x 16-column chunk wholly outside it is skipped (warp-uniform)
					if (!(kq+c0+15-row0 < 0 || kq+c0-(row0+31) >= num)) {
						uint32_t s0[16], s1[16], s2[16];
						const uint32_t ta = tmem + ((uint32_t)row0<<16) + (uint32_t)buf*256u + (uint32_t)c0;
						tmem_ld16(ta, s0); tmem_ld16(ta+64u, s1); tmem_ld16(ta+128u, s2);
						asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
						#pragma unroll
						for (int e = 0; e < 16; ++e) {
							const int d = kq + c0+e - row;            // disparity index of (pixel, column)
							if (d >= 0 && d < num) {
								const float sum = __uint_as_float(s0[e]), sumSq = __uint_as_float(s1[e]), nom = __uint_as_float(s2[e]);
								const float normSq1 = fmaf(-sum*invW, sum, sumSq);
								const float ncc = nom*rsqrtf(fmaf(normSq0, normSq1, 1e-3f));
								// ncc <= 0 ? 255 : floor((1 - min(ncc, 1)) * 255 + .5); 255 for windows that leave the right image
								int cv = ncc <= 0.f ? 255 : __float2int_rd(fmaf(-255.f, fminf(ncc, 1.f), 255.5f));
								if (d < dlo || d >= dhi) cv = 255;
								sTile[row*TILE_PITCH + d] = (uint8_t)cv;
							}
						}
					}
					if (k+2 < nTiles) {
						asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
						bar_arrive(BAR_DRAIN+buf, TC_THREADS);      // the issuing warp may overwrite this buffer
					}
				}
				bar_sync(BAR_WORKERS, TC_WORKERS);
				// the block's costs: num bytes per pixel, coalesced 16-byte stores
				const int chunks = num/16;
				for (int i = tid; i < BM*chunks; i += TC_WORKERS) {
					const int prow = i/chunks, c16 = i-prow*chunks;
					const int pcol = x0+prow;
					if (pcol < vw) {
						const uint32_t* src = (const uint32_t*)(sTile + prow*TILE_PITCH + 16*c16);
						const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
						*(uint4*)(P.costs + ((size_t)r*vw + pcol)*(size_t)numAll + dOff + 16*c16) = v;
					}
				}
				bar_sync(BAR_WORKERS, TC_WORKERS);                  // the tile is free: the next block stages into it
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}

} // namespace

cudaError_t sgm_cost_tc_configure() {
	return cudaFuncSetAttribute(sgm_cost_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
}
bool sgm_cost_tc_supports(int num) { return num == 64 || num == 128 || num == 192 || num == 256; }
cudaError_t sgm_cost_tc_launch(const SGMParams& P, int dmin, int num, cudaStream_t s) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	// slices of at most 128 disparities (the ring holds the four B tiles a block of 128 pixels x 128 disparities needs)
	for (int off = 0; off < num; off += 128)
		sgm_cost_tc_kernel<<<sms, TC_THREADS, SMEM_TOTAL, s>>>(P, dmin+off, num-off >= 128 ? 128 : num-off, num, off);
	return cudaGetLastError();
}
