/*
 * pm_oracle.cpp — CPU oracle of the PatchMatch depth+normal estimation path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED (no reference golden vectors).
 *
 * A from-scratch restatement (no OpenCV/Eigen) of the algorithm in
 *   libs/MVS/DepthMap.h:145-155,175-185,276-468      (types, weights, homography, random planes)
 *   libs/MVS/DepthMap.cpp:329-356                    (MapMatrix2ZigzagIdx)
 *   libs/MVS/DepthMap.cpp:361-626                    (ctor constants, FillPixelPatch, ScorePixelImage, ScorePixel)
 *   libs/MVS/DepthMap.cpp:630-971                    (ProcessPixel, InterpolatePixel, InitPlane)
 *   libs/MVS/SceneDensify.cpp:490-805                (passes A/B/C, scale loop)
 *   libs/Common/Types.inl:2271-2313                  (bilinear / masked bilinear sampling)
 *   libs/Common/Util.inl:754-810                     (Normal2Dir/Dir2Normal, depth similarity)
 *   libs/Common/Random.h:101-158                     (mt19937 helpers)
 *   libs/Common/Rotation.inl:701-729                 (Rodrigues)
 * with the reference's compile-time variant: DENSE_NCC_WEIGHTED, DENSE_AGGNCC_MINMEAN,
 * DENSE_SMOOTHNESS_PLANE, DENSE_REFINE_ITER (libs/MVS/DepthMap.h:43-70).
 *
 * Two schedules:
 *   ZZ — the reference's own: zig-zag pixel order, in-place sequential updates, mt19937
 *        seeded with the default seed (the reference's non-_RELEASE behaviour,
 *        DepthMap.cpp:370-372); `threads`>1 reproduces the shared-counter fan-out
 *        (SceneDensify.cpp:519-526) and is schedule-dependent exactly like the reference.
 *   RB — red-black half-sweeps with a counter-based Philox4x32-10 stream keyed by
 *        (pixel, iteration, hypothesis slot): the same per-pixel ProcessPixel, but every
 *        pixel of one colour is updated from the other colour's previous values.  This
 *        is the schedule the CUDA engine implements, so RB is bit-comparable with it.
 */
#include "oracle.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <random>
#include <thread>
#include <atomic>
#include <algorithm>

namespace {

constexpr int kHalf = 4;   // nSizeHalfWindow (DepthMap.h:277)
constexpr int kStep = 2;   // nSizeStep
constexpr int kTexels = 25;// nTexels
constexpr float kPI = 3.14159265358979323846f;
inline float d2r(float d) { return d*(kPI/180.f); }

struct V3 { float x, y, z; };
inline float dot(const V3& a, const V3& b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
inline V3 cross(const V3& a, const V3& b) { return V3{a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x}; }
inline float norm(const V3& a) { return std::sqrt(dot(a,a)); }

// ---- 3x3 double helpers (row-major) --------------------------------------------------
void mul33(const double* A, const double* B, double* C) {
	double T[9];
	for (int i=0;i<3;++i) for (int j=0;j<3;++j) T[i*3+j] = A[i*3+0]*B[0*3+j] + A[i*3+1]*B[1*3+j] + A[i*3+2]*B[2*3+j];
	memcpy(C, T, sizeof(T));
}
void mul31(const double* A, const double* v, double* r) {
	double t[3];
	for (int i=0;i<3;++i) t[i] = A[i*3+0]*v[0] + A[i*3+1]*v[1] + A[i*3+2]*v[2];
	memcpy(r, t, sizeof(t));
}
void transpose33(const double* A, double* T) {
	double R[9];
	for (int i=0;i<3;++i) for (int j=0;j<3;++j) R[j*3+i] = A[i*3+j];
	memcpy(T, R, sizeof(R));
}
void inv33(const double* A, double* I) {
	const double a=A[0],b=A[1],c=A[2],d=A[3],e=A[4],f=A[5],g=A[6],h=A[7],i=A[8];
	const double det = a*(e*i-f*h) - b*(d*i-f*g) + c*(d*h-e*g);
	const double id = 1.0/det;
	double R[9] = {(e*i-f*h)*id, (c*h-b*i)*id, (b*f-c*e)*id,
	               (f*g-d*i)*id, (a*i-c*g)*id, (c*d-a*f)*id,
	               (d*h-e*g)*id, (b*g-a*h)*id, (a*e-b*d)*id};
	memcpy(I, R, sizeof(R));
}

// ---- per-view constants (ViewData::Init, DepthMap.h:175-185) ---------------------------
struct View {
	const float* img; int w, h;
	double Hl[9], Hm[3], Hr[9];
	const float* dmap; int dw, dh;
	float Tl[9], Tm[3], Tr[9], Tn[3];
};

void initView(const oracle_view& v, const oracle_view& ref, View& o) {
	o.img = v.image; o.w = v.width; o.h = v.height;
	double KR[9], RrT[9], dC[3];
	mul33(v.K, v.R, KR);
	transpose33(ref.R, RrT);
	mul33(KR, RrT, o.Hl);                                   // Hl = K R Rref^T
	for (int i=0;i<3;++i) dC[i] = ref.C[i]-v.C[i];
	mul31(KR, dC, o.Hm);                                    // Hm = K R (Cref - C)
	inv33(ref.K, o.Hr);                                     // Hr = Kref^-1
	o.dmap = v.depth; o.dw = v.dwidth; o.dh = v.dheight;
	if (v.depth) {
		double KdRd[9], T[9], t[3], RdT[9], iKd[9], KrRr[9];
		mul33(v.Kd, v.Rd, KdRd);
		mul33(KdRd, RrT, T);                                // Tl = Kd Rd Rref^T
		for (int i=0;i<9;++i) o.Tl[i] = (float)T[i];
		for (int i=0;i<3;++i) dC[i] = ref.C[i]-v.Cd[i];
		mul31(KdRd, dC, t);                                 // Tm = Kd Rd (Cref - Cd)
		for (int i=0;i<3;++i) o.Tm[i] = (float)t[i];
		mul33(ref.K, ref.R, KrRr);
		transpose33(v.Rd, RdT);
		inv33(v.Kd, iKd);
		mul33(KrRr, RdT, T); mul33(T, iKd, T);              // Tr = Kref Rref Rd^T Kd^-1
		for (int i=0;i<9;++i) o.Tr[i] = (float)T[i];
		for (int i=0;i<3;++i) dC[i] = v.Cd[i]-ref.C[i];
		mul31(KrRr, dC, t);                                 // Tn = Kref Rref (Cd - Cref)
		for (int i=0;i<3;++i) o.Tn[i] = (float)t[i];
	}
}

// ---- sampling (Types.inl:2271-2313, Types.h:1633-1651) --------------------------------
inline bool insideBorder1(float x, float y, int w, int h) {
	return x >= 1.f && y >= 1.f && x <= float(w-2) && y <= float(h-2);
}
inline float sampleBilinear(const float* img, int w, float px, float py) {
	const int lx = (int)px, ly = (int)py;
	const float x = px-lx, x1 = 1.f-x;
	const float y = py-ly, y1 = 1.f-y;
	const float* r0 = img + (size_t)ly*w + lx;
	const float* r1 = r0 + w;
	return (r0[0]*x1 + r0[1]*x)*y1 + (r1[0]*x1 + r1[1]*x)*y;
}
// bilinear using only the taps whose depth is similar to `ref` (IsDepthSimilar 3%)
inline bool depthSimilar(float d0, float d1, float th) { return std::fabs(d0-d1)/d0 < th; }
inline bool sampleDepthMasked(const float* img, int w, float px, float py, float ref, float& v) {
	const int lx = (int)px, ly = (int)py;
	const float x = px-lx, x1 = 1.f-x;
	const float y = py-ly, y1 = 1.f-y;
	const float x0y0 = img[(size_t)ly*w+lx],     x1y0 = img[(size_t)ly*w+lx+1];
	const float x0y1 = img[(size_t)(ly+1)*w+lx], x1y1 = img[(size_t)(ly+1)*w+lx+1];
	const bool b00 = depthSimilar(ref, x0y0, 0.03f), b10 = depthSimilar(ref, x1y0, 0.03f);
	const bool b01 = depthSimilar(ref, x0y1, 0.03f), b11 = depthSimilar(ref, x1y1, 0.03f);
	if (!b00 && !b10 && !b01 && !b11)
		return false;
	v = y1*(x1*(b00 ? x0y0 : (b10 ? x1y0 : (b01 ? x0y1 : x1y1))) + x*(b10 ? x1y0 : (b00 ? x0y0 : (b11 ? x1y1 : x0y1)))) +
	    y *(x1*(b01 ? x0y1 : (b11 ? x1y1 : (b00 ? x0y0 : x1y0))) + x*(b11 ? x1y1 : (b01 ? x0y1 : (b10 ? x1y0 : x0y0))));
	return true;
}

// ---- direction encoding (Util.inl:754-766) --------------------------------------------
inline void normal2dir(const V3& d, float& a, float& b) { a = std::atan2(d.y, d.x); b = std::acos(d.z); }
inline void dir2normal(float a, float b, V3& d) {
	const float siny = std::sin(b);
	d.x = std::cos(a)*siny; d.y = std::sin(a)*siny; d.z = std::cos(b);
}
// Rodrigues rotation of v about axis wa by phi (Rotation.inl:701-729), float
inline V3 rotateAxisAngle(const V3& wa, float phi, const V3& v) {
	if (std::fabs(phi) < 0.0001f)
		return v;
	const float inv = 1.f/norm(wa);
	const float w0 = wa.x*inv, w1 = wa.y*inv, w2 = wa.z*inv;
	const float O[9] = {0,-w2,w1, w2,0,-w0, -w1,w0,0};
	float OO[9];
	for (int i=0;i<3;++i) for (int j=0;j<3;++j) OO[i*3+j] = O[i*3+0]*O[0*3+j] + O[i*3+1]*O[1*3+j] + O[i*3+2]*O[2*3+j];
	const float s = std::sin(phi), c1 = 1.f-std::cos(phi);
	float R[9];
	for (int i=0;i<9;++i) R[i] = ((i%4)==0 ? 1.f : 0.f) + O[i]*s + OO[i]*c1;
	return V3{R[0]*v.x+R[1]*v.y+R[2]*v.z, R[3]*v.x+R[4]*v.y+R[5]*v.z, R[6]*v.x+R[7]*v.y+R[8]*v.z};
}

// ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------
inline void philox(const uint32_t c[4], const uint32_t k[2], uint32_t o[4]) {
	uint32_t c0=c[0], c1=c[1], c2=c[2], c3=c[3], k0=k[0], k1=k[1];
	for (int r=0; r<10; ++r) {
		const uint64_t p0 = (uint64_t)0xD2511F53u*c0, p1 = (uint64_t)0xCD9E8D57u*c2;
		const uint32_t n0 = (uint32_t)(p1>>32)^c1^k0, n1 = (uint32_t)p1;
		const uint32_t n2 = (uint32_t)(p0>>32)^c3^k1, n3 = (uint32_t)p0;
		c0=n0; c1=n1; c2=n2; c3=n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	o[0]=c0; o[1]=c1; o[2]=c2; o[3]=c3;
}
inline float u32ToUnit(uint32_t u) { return (float)u/4294967296.f; } // Random.h:113-115: (float)u/(float)max()

struct WeightedPatch { // WeightedPatchFix<25>, DepthMap.h:145-155
	float weight[kTexels], tempWeight[kTexels];
	float sumWeights, normSq0;
};

struct Close { float depth; V3 normal; V3 X; }; // NeighborEstimate (DepthMap.h:300-306)

// shared, read-only per (scale) job description
struct Job {
	std::vector<View> views;   // neighbours (without the reference)
	const float* img0; int w, h;
	double K0[9];
	float dMin, dMax;
	const float* lowres;       // nullable
	oracle_params prm;
	float* depth; float* normal; float* conf;
	std::vector<WeightedPatch>* weights; // lazily filled cache (weightMap0)
	// RB engine schedule with propagation bit 8 (0x100): a pixel remembers whether its plane changed during its last
	// update; a direction whose candidates are all unchanged is not re-tested (it lost against this pixel last sweep)
	uint8_t* changed;                    // nullable: per-pixel flag, 1 = changed (or never tested)
	// ignore-mask of this level (0 = pixel not in the pixel list, DepthMap.cpp:343) or null
	const uint8_t* mask;
};
std::atomic<long long> g_propTested(0), g_propSkipped(0);

struct Estimator {
	const Job& J;
	const int iter;
	const int dir; // 0 = LT2RB, 1 = RB2LT
	// derived constants (DepthEstimator ctor, DepthMap.cpp:361-412)
	const float dMinSqr, dMaxSqr;
	const float smoothBonusDepth, smoothBonusNormal, smoothSigmaDepth, smoothSigmaNormal;
	const float thMagnitudeSq, angle1Range, angle2Range;
	const float thConfSmall, thConfBig, thConfRand, thRobust;
	// per-pixel state
	int x0, y0;
	double X0[3];
	float normSq0;
	const WeightedPatch* wp;
	Close close[4]; int nClose;
	float planeN[3], planeD;
	std::vector<float> scores;
	// rng
	std::mt19937 mt;
	uint32_t pblock[4]; int pidx; uint32_t pctr[4]; uint32_t pkey[2];

	Estimator(const Job& j, int it)
		: J(j), iter(it), dir(it%2 ? 1 : 0),
		dMinSqr(std::sqrt(j.dMin)), dMaxSqr(std::sqrt(j.dMax)),
		smoothBonusDepth(1.f-j.prm.fRandomSmoothBonus), smoothBonusNormal((1.f-j.prm.fRandomSmoothBonus)*0.96f),
		smoothSigmaDepth(-1.f/(2.f*j.prm.fRandomSmoothDepth*j.prm.fRandomSmoothDepth)),
		smoothSigmaNormal(-1.f/(2.f*d2r(j.prm.fRandomSmoothNormal)*d2r(j.prm.fRandomSmoothNormal))),
		thMagnitudeSq(j.prm.fDescriptorMinMagnitudeThreshold>0 ? j.prm.fDescriptorMinMagnitudeThreshold*j.prm.fDescriptorMinMagnitudeThreshold : -1.f),
		angle1Range(d2r(j.prm.fRandomAngle1Range)), angle2Range(d2r(j.prm.fRandomAngle2Range)),
		thConfSmall(j.prm.fNCCThresholdKeep*0.66f), thConfBig(j.prm.fNCCThresholdKeep*0.9f),
		thConfRand(j.prm.fNCCThresholdKeep*1.1f), thRobust(j.prm.fNCCThresholdKeep*4.f/3.f),
		nClose(0), scores(j.views.size()), mt(std::mt19937::default_seed)
	{
		pkey[0] = j.prm.seed; pkey[1] = 0xB200C0DEu;
	}

	// --- random numbers -----------------------------------------------------------------
	// RB: counter = (pixel index, phase, slot, 0); phase 0 = pass A, 1+iter = sweep `iter`
	void rngBegin(uint32_t phase, uint32_t slot) {
		if (J.prm.schedule == 1) {
			pctr[0] = (uint32_t)(y0*J.w + x0); pctr[1] = phase; pctr[2] = slot; pctr[3] = 0;
			philox(pctr, pkey, pblock);
			pidx = 0;
		}
	}
	float rnd() {
		if (J.prm.schedule == 1)
			return u32ToUnit(pblock[pidx++]);
		return u32ToUnit((uint32_t)mt());
	}
	float randomRange(float a, float b) { return a + (b-a)*rnd(); }
	float randomMeanRange(float mean, float delta) { return mean + delta*(2.f*rnd()-1.f); }
	float RandomDepth() { const float s = randomRange(dMinSqr, dMaxSqr); return s*s; }
	V3 RandomNormal(const V3& viewRay) {
		const float a = randomRange(d2r(0.f), d2r(180.f));
		const float b = randomRange(d2r(90.f), d2r(180.f));
		V3 n; dir2normal(a, b, n);
		return dot(n, viewRay) > 0 ? V3{-n.x,-n.y,-n.z} : n;
	}

	// --- patch ----------------------------------------------------------------------------
	bool PreparePixelPatch(int x, int y) {
		x0 = x; y0 = y;
		return x-kHalf >= 0 && y-kHalf >= 0 && x+kHalf < J.w && y+kHalf < J.h;
	}
	bool FillPixelPatch() {
		WeightedPatch& w = (*J.weights)[(size_t)y0*J.w + x0];
		if (w.normSq0 == 0) {
			w.sumWeights = 0;
			int n = 0;
			const float colCenter = J.img0[(size_t)y0*J.w + x0];
			const float sigmaColor = -1.f/(2.f*(0.1f*0.1f));
			const float sigmaSpatial = -1.f/(2.f*float((kHalf-1)*(kHalf-1)));
			for (int i=-kHalf; i<=kHalf; i+=kStep) {
				for (int j=-kHalf; j<=kHalf; j+=kStep) {
					const float I = J.img0[(size_t)(y0+i)*J.w + x0+j];
					const float dI = I-colCenter;
					const float wColor = dI*dI*sigmaColor;
					const float wSpatial = float(j*j + i*i)*sigmaSpatial;
					const float wgt = std::exp(wColor+wSpatial);
					w.tempWeight[n] = I; w.weight[n] = wgt;
					w.normSq0 += I*wgt;
					w.sumWeights += wgt;
					++n;
				}
			}
			const float tm = w.normSq0/w.sumWeights;
			w.normSq0 = 0;
			for (n=0; n<kTexels; ++n) {
				const float t = w.tempWeight[n]-tm;
				w.tempWeight[n] = w.weight[n]*t;
				w.normSq0 += w.tempWeight[n]*t;
			}
		}
		wp = &w;
		normSq0 = w.normSq0;
		if (normSq0 < thMagnitudeSq && (!J.lowres || J.lowres[(size_t)y0*J.w+x0] <= 0))
			return false;
		X0[0] = ((double)x0-J.K0[2])/J.K0[0];
		X0[1] = ((double)y0-J.K0[5])/J.K0[4];
		X0[2] = 1.0;
		return true;
	}
	V3 viewDir() const { return V3{(float)X0[0], (float)X0[1], (float)X0[2]}; }

	void InitPlane(float depth, const V3& n) {
		planeN[0]=n.x; planeN[1]=n.y; planeN[2]=n.z;
		planeD = -depth*dot(n, viewDir());
	}

	// --- scoring --------------------------------------------------------------------------
	float ScorePixelImage(const View& v, float depth, const V3& normal) {
		// H = (Hl + Hm n^T / (n.X0 depth)) Hr  in double, then float (DepthMap.h:414-423)
		const double n[3] = {normal.x, normal.y, normal.z};
		const double inv = 1.0/((n[0]*X0[0] + n[1]*X0[1] + n[2]*X0[2])*(double)depth);
		double M[9], Hd[9];
		for (int i=0;i<3;++i) for (int j=0;j<3;++j) M[i*3+j] = v.Hl[i*3+j] + v.Hm[i]*(n[j]*inv);
		mul33(M, v.Hr, Hd);
		float H[9];
		for (int i=0;i<9;++i) H[i] = (float)Hd[i];
		const float px = float(x0-kHalf), py = float(y0-kHalf);
		float X[3] = {H[0]*px + H[1]*py + H[2], H[3]*px + H[4]*py + H[5], H[6]*px + H[7]*py + H[8]};
		float baseX[3] = {X[0], X[1], X[2]};
		for (int i=0;i<9;++i) H[i] *= float(kStep);
		int k = 0;
		float sum = 0, sumSq = 0, num = 0;
		const WeightedPatch& w = *wp;
		for (int i=-kHalf; i<=kHalf; i+=kStep) {
			for (int j=-kHalf; j<=kHalf; j+=kStep) {
				const float ptx = X[0]/X[2], pty = X[1]/X[2];
				if (!insideBorder1(ptx, pty, v.w, v.h))
					return thRobust;
				const float val = sampleBilinear(v.img, v.w, ptx, pty);
				const float vw = val*w.weight[k];
				sum += vw;
				sumSq += val*vw;
				num += val*w.tempWeight[k];
				++k;
				X[0] += H[0]; X[1] += H[3]; X[2] += H[6];
			}
			baseX[0] += H[1]; baseX[1] += H[4]; baseX[2] += H[7];
			X[0] = baseX[0]; X[1] = baseX[1]; X[2] = baseX[2];
		}
		const float normSq1 = sumSq - sum*sum/w.sumWeights;
		const float nrmSq = normSq0*normSq1;
		if (nrmSq <= 1e-16f)
			return thRobust;
		float ncc = num/std::sqrt(nrmSq);
		ncc = std::min(std::max(ncc, -1.f), 1.f);
		float score = 1.f-ncc;
		// encourage smoothness (DepthMap.cpp:522-534)
		for (int c=0; c<nClose; ++c) {
			const Close& nb = close[c];
			const float dist = planeN[0]*nb.X.x + planeN[1]*nb.X.y + planeN[2]*nb.X.z + planeD;
			const float rd = dist/depth;
			const float factorDepth = std::exp(rd*rd*smoothSigmaDepth);
			float ca = dot(normal, nb.normal)/std::sqrt(dot(normal,normal)*dot(nb.normal,nb.normal));
			ca = std::min(std::max(ca, -1.f), 1.f);
			const float ang = std::acos(ca);
			const float factorNormal = std::exp(ang*ang*smoothSigmaNormal);
			score *= (1.f-smoothBonusDepth*factorDepth)*(1.f-smoothBonusNormal*factorNormal);
		}
		// geometric consistency (DepthMap.cpp:535-551)
		if (v.dmap) {
			float consistency = 4.f;
			const float Xc[3] = {float(X0[0])*depth, float(X0[1])*depth, depth};
			const float X1[3] = {
				v.Tl[0]*Xc[0]+v.Tl[1]*Xc[1]+v.Tl[2]*Xc[2]+v.Tm[0],
				v.Tl[3]*Xc[0]+v.Tl[4]*Xc[1]+v.Tl[5]*Xc[2]+v.Tm[1],
				v.Tl[6]*Xc[0]+v.Tl[7]*Xc[1]+v.Tl[8]*Xc[2]+v.Tm[2]};
			if (X1[2] > 0) {
				const float x1x = X1[0]/X1[2], x1y = X1[1]/X1[2];
				if (insideBorder1(x1x, x1y, v.dw, v.dh)) {
					float depth1;
					if (sampleDepthMasked(v.dmap, v.dw, x1x, x1y, X1[2], depth1)) {
						const float P[3] = {x1x*depth1, x1y*depth1, depth1};
						const float B[3] = {
							v.Tr[0]*P[0]+v.Tr[1]*P[1]+v.Tr[2]*P[2]+v.Tn[0],
							v.Tr[3]*P[0]+v.Tr[4]*P[1]+v.Tr[5]*P[2]+v.Tn[1],
							v.Tr[6]*P[0]+v.Tr[7]*P[1]+v.Tr[8]*P[2]+v.Tn[2]};
						const float xbx = B[0]/B[2], xby = B[1]/B[2];
						const float ex = float(x0)-xbx, ey = float(y0)-xby;
						const float dist = std::sqrt(ex*ex + ey*ey);
						consistency = std::min(std::sqrt(dist*(dist+2.f)), consistency);
					}
				}
			}
			score += J.prm.fEstimationGeometricWeight*consistency;
		}
		// low-resolution depth prior (DepthMap.cpp:552-561)
		if (J.lowres) {
			const float d0 = J.lowres[(size_t)y0*J.w+x0];
			if (d0 > 0) {
				const float deltaDepth = std::min(std::fabs(d0-depth)/d0, 0.5f);
				const float sigma = -1.f/(1.f*0.02f);
				const float f = std::exp(normSq0*sigma);
				score = (1.f-f)*score + f*deltaDepth;
			}
		}
		return std::min(2.f, score);
	}

	float ScorePixel(float depth, const V3& normal) {
		const size_t N = J.views.size();
		for (size_t i=0; i<N; ++i)
			scores[i] = ScorePixelImage(J.views[i], depth, normal);
		if (N <= 1) // idxScore == 0
			return *std::min_element(scores.begin(), scores.end());
		// MINMEAN (DepthMap.cpp:595-610): mean of the two smallest unless the 2nd >= thRobust
		std::nth_element(scores.begin(), scores.begin()+1, scores.end());
		float score = scores[0];
		if (scores[1] >= thRobust)
			return score;
		return (score+scores[1])/2;
	}

	// --- geometry -------------------------------------------------------------------------
	float InterpolatePixel(int nx, int ny, float depth, const V3& normal) const {
		float depthNew;
		if (x0 == nx) {
			const float nx1 = (float)(((double)y0 - J.K0[5])/J.K0[4]);
			const float denom = normal.z + nx1*normal.y;
			if (std::fabs(denom) < 0.0001f)
				return depth;
			const float x1 = (float)(((double)ny - J.K0[5])/J.K0[4]);
			const float nom = depth*(normal.z + x1*normal.y);
			depthNew = nom/denom;
		} else {
			const float nx1 = (float)(((double)x0 - J.K0[2])/J.K0[0]);
			const float denom = normal.z + nx1*normal.x;
			if (std::fabs(denom) < 0.0001f)
				return depth;
			const float x1 = (float)(((double)nx - J.K0[2])/J.K0[0]);
			const float nom = depth*(normal.z + x1*normal.x);
			depthNew = nom/denom;
		}
		return (J.dMin <= depthNew && depthNew < J.dMax) ? depthNew : depth;
	}
	void CorrectNormal(V3& n) const {
		const V3 vd = viewDir();
		const float cosAngLen = dot(n, vd);
		if (cosAngLen >= 0) {
			const float ang = std::min((std::acos(cosAngLen/norm(vd)) - d2r(90.f))*1.01f, -0.001f);
			n = rotateAxisAngle(cross(n, vd), ang, n);
		}
	}

	void addClose(int nx, int ny, float nd) {
		Close& c = close[nClose++];
		c.depth = nd;
		const float* nn = J.normal + ((size_t)ny*J.w+nx)*3;
		c.normal = V3{nn[0], nn[1], nn[2]};
		c.X = V3{(float)(((double)nx-J.K0[2])*(double)nd/J.K0[0]), (float)(((double)ny-J.K0[5])*(double)nd/J.K0[4]), nd};
	}

	// --- pass A pixel (ScoreDepthMapTmp, SceneDensify.cpp:490-517) ---------------------------
	void ScorePixelInit(int x, int y) {
		const size_t i = (size_t)y*J.w+x;
		if (J.mask && !J.mask[i])
			return; // not in the pixel list; DepthData::ApplyIgnoreMask zeroed the maps (DepthMap.cpp:215-230)
		if (!PreparePixelPatch(x, y) || !FillPixelPatch()) {
			J.depth[i] = 0; J.normal[i*3]=J.normal[i*3+1]=J.normal[i*3+2]=0; J.conf[i] = 2.f;
			return;
		}
		float& depth = J.depth[i];
		V3 normal{J.normal[i*3], J.normal[i*3+1], J.normal[i*3+2]};
		const V3 vd = viewDir();
		rngBegin(0, 0);
		if (!(J.dMin <= depth && depth < J.dMax)) {
			depth = RandomDepth();
			normal = RandomNormal(vd);
		} else if (dot(normal, vd) >= 0) {
			normal = RandomNormal(vd);
		}
		J.normal[i*3]=normal.x; J.normal[i*3+1]=normal.y; J.normal[i*3+2]=normal.z;
		nClose = 0;
		J.conf[i] = ScorePixel(depth, normal);
	}

	// --- pass B pixel (ProcessPixel, DepthMap.cpp:630-852) -----------------------------------
	void ProcessPixel(int x, int y) {
		if (J.mask && !J.mask[(size_t)y*J.w+x])
			return;
		if (!PreparePixelPatch(x, y) || !FillPixelPatch())
			return;
		const int w = J.w, h = J.h;
		int prop[4][2]; int nProp = 0;
		int farXY[4][2]; bool farUse[4] = {false, false, false, false};
		bool dirChanged[4] = {true, true, true, true};
		const bool useFlags = J.prm.schedule == 1 && (J.prm.propagation & 0x100) && J.changed;
		nClose = 0;
		// neighbour order: causal pair first, then the anti-causal pair
		const int offs[2][4][2] = {{{-1,0},{0,-1},{1,0},{0,1}}, {{1,0},{0,1},{-1,0},{0,-1}}};
		for (int k=0; k<4; ++k) {
			const int ox = offs[dir][k][0], oy = offs[dir][k][1];
			bool ok;
			if (ox < 0) ok = x0 > kHalf; else if (ox > 0) ok = x0 < w-kHalf;
			else if (oy < 0) ok = y0 > kHalf; else ok = y0 < h-kHalf;
			if (!ok) continue;
			const int nx = x0+ox, ny = y0+oy;
			const float nd = J.depth[(size_t)ny*w+nx];
			if (useFlags) dirChanged[k] = J.changed[(size_t)ny*w+nx] != 0;
			if (nd > 0) {
				const bool propagate = (k < 2) || (J.prm.schedule == 1 && (J.prm.propagation & 15) == 4);
				if (propagate) { prop[nProp][0] = nClose; prop[nProp][1] = k; ++nProp; }
				addClose(nx, ny, nd);
			}
			// RB far propagation (engine schedule, not in the reference): in direction k the candidate is the
			// pixel of the other colour at distance 1, 3, ... 2F+1 with the lowest stored cost (ties: the nearest)
			const int F = J.prm.schedule == 1 ? ((J.prm.propagation >> 4) & 15) : 0;
			if (F > 0 && ((k < 2) || (J.prm.propagation & 15) == 4)) {
				float bestC = nd > 0 ? J.conf[(size_t)ny*w+nx] : 3.f;
				for (int f = 1; f <= F; ++f) {
					const int fx = x0+ox*(2*f+1), fy = y0+oy*(2*f+1);
					if (fx < kHalf || fy < kHalf || fx >= w-kHalf || fy >= h-kHalf) break;
					const size_t fi = (size_t)fy*w+fx;
					if (useFlags && J.changed[fi]) dirChanged[k] = true;
					if (J.depth[fi] > 0 && J.conf[fi] < bestC) { bestC = J.conf[fi]; farUse[k] = true; farXY[k][0] = fx; farXY[k][1] = fy; }
				}
				if (farUse[k] && !(nd > 0)) { prop[nProp][0] = -1; prop[nProp][1] = k; ++nProp; } // near neighbour invalid: the far one still propagates
			}
		}
		const size_t i0 = (size_t)y0*w+x0;
		float& conf = J.conf[i0];
		float& depth = J.depth[i0];
		float* pn = J.normal + i0*3;
		V3 normal{pn[0], pn[1], pn[2]};
		const V3 vd = viewDir();
		const float depthIn = depth; const V3 normalIn = normal;
		struct FlagSetter { // on every exit path: record whether the plane changed
			const Job& J; size_t i; const float& d; const float* pn; float d0; V3 n0;
			~FlagSetter() { if (J.changed) J.changed[i] = (d != d0 || pn[0] != n0.x || pn[1] != n0.y || pn[2] != n0.z) ? 1 : 0; }
		} flagSetter{J, i0, depth, pn, depthIn, normalIn};
		// propagation
		int nTested = 0;   // candidates scored (engine schedule: bounds the refinement tries, see below)
		for (int p=0; p<nProp; ++p) {
			const int c = prop[p][0], k = prop[p][1];
			if (!dirChanged[k]) { g_propSkipped.fetch_add(1, std::memory_order_relaxed); continue; }
			g_propTested.fetch_add(1, std::memory_order_relaxed);
			int nx = x0+offs[dir][k][0], ny = y0+offs[dir][k][1];
			Close nb;
			if (farUse[k]) {
				nx = farXY[k][0]; ny = farXY[k][1];
				const float* fn = J.normal + ((size_t)ny*w+nx)*3;
				nb.depth = J.depth[(size_t)ny*w+nx]; nb.normal = V3{fn[0], fn[1], fn[2]};
			} else
				nb = close[c];
			if (J.conf[(size_t)ny*w+nx] >= J.prm.fNCCThresholdKeep)
				continue;
			++nTested;
			nb.depth = InterpolatePixel(nx, ny, nb.depth, nb.normal);
			CorrectNormal(nb.normal);
			InitPlane(nb.depth, nb.normal);
			const float nconf = ScorePixel(nb.depth, nb.normal);
			if (conf > nconf) { conf = nconf; depth = nb.depth; normal = nb.normal; }
		}
		// refinement
		// RB engine schedule, propagation bits 12-15 = evaluation cap E > 0 (not in the reference): a pixel that scored c
		// candidates above spends at most max(1, E - c) random / perturbation tries; the Philox slots keep their numbering
		const int evalCap = J.prm.schedule == 1 ? ((J.prm.propagation >> 12) & 15) : 0;
		const int nTries = evalCap > 0 ? std::min(J.prm.nRandomIters, std::max(evalCap-nTested, 1)) : J.prm.nRandomIters;
		const uint32_t phase = 1u+(uint32_t)iter;
		unsigned idxScaleRange = 0;
		static const float scaleRanges[12] = {1.f, 0.5f, 0.25f, 0.125f, 0.0625f, 0.03125f, 0.015625f, 0.0078125f, 0.00390625f, 0.001953125f, 0.0009765625f, 0.00048828125f};
		bool restarted = false;
		for (;;) { // RefineIters label
			if (conf <= thConfSmall)
				idxScaleRange = 2;
			else if (conf <= thConfBig)
				idxScaleRange = 1;
			else if (conf >= thConfRand && !restarted) {
				// completely random hypotheses, no smoothness
				restarted = true;
				nClose = 0;
				bool again = false;
				for (int it=0; it<nTries; ++it) {
					rngBegin(phase, (uint32_t)it);
					const float ndepth = RandomDepth();
					const V3 nnormal = RandomNormal(vd);
					const float nconf = ScorePixel(ndepth, nnormal);
					if (conf > nconf) {
						conf = nconf; depth = ndepth; normal = nnormal;
						if (conf < thConfRand) { again = true; break; }
					}
				}
				if (again) continue;
				pn[0]=normal.x; pn[1]=normal.y; pn[2]=normal.z;
				return;
			}
			break;
		}
		float scaleRange = scaleRanges[idxScaleRange];
		const float depthRange = depth*J.prm.fRandomDepthRatio;
		float pa, pb;
		normal2dir(normal, pa, pb);
		for (int it=0; it<nTries; ++it) {
			rngBegin(phase, (uint32_t)(J.prm.nRandomIters+it));
			const float ndepth = randomMeanRange(depth, depthRange*scaleRange);
			if (!(J.dMin <= ndepth && ndepth < J.dMax))
				continue;
			const float na = randomMeanRange(pa, angle1Range*scaleRange);
			const float nb = randomMeanRange(pb, angle2Range*scaleRange);
			V3 nnormal; dir2normal(na, nb, nnormal);
			if (dot(nnormal, vd) >= 0)
				continue;
			InitPlane(ndepth, nnormal);
			const float nconf = ScorePixel(ndepth, nnormal);
			if (conf > nconf) {
				conf = nconf; depth = ndepth; normal = nnormal;
				pa = na; pb = nb;
				scaleRange = scaleRanges[++idxScaleRange];
			}
		}
		pn[0]=normal.x; pn[1]=normal.y; pn[2]=normal.z;
	}
};

// NB: in the reference a pixel re-entering RefineIters with conf >= thConfRand after a
// restart cannot happen (the goto is taken only when conf < thConfRand), so `restarted`
// above never changes behaviour; it only guards against an endless loop.

void zigzag(int w, int h, int rawStride, std::vector<uint16_t>& coords) {
	// MapMatrix2ZigzagIdx (DepthMap.cpp:329-356), no mask
	coords.clear(); coords.reserve((size_t)w*h*2);
	const int w1 = w-1;
	for (int dy=0, hh=rawStride; dy<h; dy+=hh) {
		if (hh*2 > h-dy)
			hh = h-dy;
		int lastX = 0;
		int x = 0, y = 0;
		for (int i=0, ei=w*hh; i<ei; ++i) {
			coords.push_back((uint16_t)x); coords.push_back((uint16_t)(y+dy));
			if (x-- == 0 || ++y == hh) {
				if (++lastX < w) { x = lastX; y = 0; }
				else { x = w1; y = lastX-w1; }
			}
		}
	}
}

void buildJob(const oracle_view* views, int nViews, const oracle_params* prm, float dMin, float dMax,
	const float* lowres, float* depth, float* normal, float* conf, std::vector<WeightedPatch>& weights, Job& J)
{
	J.views.resize(nViews-1);
	for (int i=1; i<nViews; ++i)
		initView(views[i], views[0], J.views[i-1]);
	J.img0 = views[0].image; J.w = views[0].width; J.h = views[0].height;
	memcpy(J.K0, views[0].K, sizeof(J.K0));
	J.dMin = dMin; J.dMax = dMax; J.lowres = lowres; J.prm = *prm;
	J.depth = depth; J.normal = normal; J.conf = conf;
	if (weights.size() != (size_t)J.w*J.h) {
		weights.assign((size_t)J.w*J.h, WeightedPatch());
		for (auto& w: weights) w.normSq0 = 0;
	}
	J.weights = &weights;
	J.changed = nullptr;
	J.mask = nullptr;
}

template <typename F>
void parallelRows(int threads, int h, F f) { // f(threadIndex, row)
	if (threads <= 1) { for (int y=0; y<h; ++y) f(0, y); return; }
	std::atomic<int> next(0);
	std::vector<std::thread> pool;
	for (int t=0; t<threads; ++t)
		pool.emplace_back([&, t]{ int y; while ((y = next.fetch_add(1)) < h) f(t, y); });
	for (auto& t: pool) t.join();
}

void passA(const Job& J) {
	const int threads = std::max(1, J.prm.threads);
	if (J.prm.schedule == 0 && threads == 1) {
		// sequential in zig-zag order so that the mt19937 stream is the reference's
		std::vector<uint16_t> coords;
		zigzag(J.w, J.h, std::max(64, threads*8), coords);
		Estimator E(J, 0);
		for (size_t i=0; i<coords.size(); i+=2)
			E.ScorePixelInit(coords[i], coords[i+1]);
		return;
	}
	std::vector<Estimator*> est(threads);
	for (int t=0; t<threads; ++t) { est[t] = new Estimator(J, 0); est[t]->mt.seed(std::mt19937::default_seed + 31u*(uint32_t)t); }
	parallelRows(threads, J.h, [&](int t, int y) {
		for (int x=0; x<J.w; ++x) est[t]->ScorePixelInit(x, y);
	});
	for (Estimator* e: est) delete e;
}

void passB(const Job& J, int iter, int half) {
	const int threads = std::max(1, J.prm.threads);
	if (J.prm.schedule == 0) {
		std::vector<uint16_t> coords;
		zigzag(J.w, J.h, std::max(64, threads*8), coords);
		const size_t n = coords.size()/2;
		const bool rev = (iter%2) != 0;
		if (threads == 1) {
			Estimator E(J, iter);
			// the reference re-creates the estimator (and its rng) every pass; advance the
			// stream so successive iterations do not replay identical perturbations
			E.mt.seed(std::mt19937::default_seed + 977u*(uint32_t)(iter+1));
			for (size_t i=0; i<n; ++i) {
				const size_t k = rev ? n-1-i : i;
				E.ProcessPixel(coords[k*2], coords[k*2+1]);
			}
		} else {
			// shared-counter fan-out (SceneDensify.cpp:519-526): schedule dependent by design
			std::atomic<size_t> next(0);
			std::vector<std::thread> pool;
			for (int t=0; t<threads; ++t)
				pool.emplace_back([&, t]{
					Estimator E(J, iter);
					E.mt.seed(std::mt19937::default_seed + 977u*(uint32_t)(iter+1) + 31u*(uint32_t)t);
					size_t i;
					while ((i = next.fetch_add(1)) < n) {
						const size_t k = rev ? n-1-i : i;
						E.ProcessPixel(coords[k*2], coords[k*2+1]);
					}
				});
			for (auto& t: pool) t.join();
		}
		return;
	}
	for (int colour=0; colour<2; ++colour) {
		if (half >= 0 && half != colour) continue;
		parallelRows(threads, J.h, [&](int, int y) {
			Estimator E(J, iter);
			for (int x=((y+colour)&1); x<J.w; x+=2) E.ProcessPixel(x, y);
		});
	}
}

void passC(int w, int h, float keep, float* depth, float* normal, float* conf) {
	// EndDepthMapTmp (SceneDensify.cpp:528-548)
	for (size_t i=0, n=(size_t)w*h; i<n; ++i) {
		if (depth[i] <= 0 || conf[i] >= keep) {
			conf[i] = 0; depth[i] = 0; normal[i*3]=normal[i*3+1]=normal[i*3+2]=0;
		} else {
			conf[i] = conf[i] >= 1.f ? 0.f : 1.f-conf[i];
		}
	}
}

// ---- cv::resize restatements (OpenCV imgproc/resize.cpp; third-party) -----
// scx/scy = source/destination scale: 1/fx when the caller gave a factor (cv::resize(..., Size(), fx, fy)),
// sw/dw when it gave the destination size
void resizeArea(const float* src, int sw, int sh, float* dst, int dw, int dh, double sx, double sy) {
	const int isx = (int)std::nearbyint(sx), isy = (int)std::nearbyint(sy);
	if (std::fabs(sx-isx) < 2.2e-16 && std::fabs(sy-isy) < 2.2e-16) {
		// integer ratio: box mean over the part of the box inside the image (ResizeAreaFast)
		for (int y=0; y<dh; ++y) for (int x=0; x<dw; ++x) {
			float s = 0; int count = 0;
			for (int j=0; j<isy && y*isy+j < sh; ++j)
				for (int i=0; i<isx && x*isx+i < sw; ++i) { s += src[(size_t)(y*isy+j)*sw + x*isx+i]; ++count; }
			dst[(size_t)y*dw+x] = count == isx*isy ? s*(1.f/(isx*isy)) : s/count;
		}
		return;
	}
	struct Tab { int si, di; float alpha; };
	auto computeTab = [](int ssize, int dsize, double scale, std::vector<Tab>& tab) {
		tab.clear();
		for (int dx=0; dx<dsize; ++dx) {
			const double fsx1 = dx*scale, fsx2 = fsx1+scale;
			const double cellWidth = std::min(scale, ssize-fsx1);
			int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
			sx2 = std::min(sx2, ssize-1);
			sx1 = std::min(sx1, sx2);
			if (sx1-fsx1 > 1e-3)
				tab.push_back(Tab{sx1-1, dx, (float)((sx1-fsx1)/cellWidth)});
			for (int s=sx1; s<sx2; ++s)
				tab.push_back(Tab{s, dx, (float)(1.0/cellWidth)});
			if (fsx2-sx2 > 1e-3)
				tab.push_back(Tab{sx2, dx, (float)(std::min(std::min(fsx2-sx2, 1.), cellWidth)/cellWidth)});
		}
	};
	std::vector<Tab> xt, yt;
	computeTab(sw, dw, sx, xt);
	computeTab(sh, dh, sy, yt);
	std::vector<float> buf(dw), sum(dw);
	int prevDy = yt.empty() ? -1 : yt[0].di;
	std::fill(sum.begin(), sum.end(), 0.f);
	for (size_t j=0; j<yt.size(); ++j) {
		const float beta = yt[j].alpha; const int dy = yt[j].di; const int syi = yt[j].si;
		std::fill(buf.begin(), buf.end(), 0.f);
		for (const Tab& t: xt)
			buf[t.di] += src[(size_t)syi*sw + t.si]*t.alpha;
		if (dy != prevDy) {
			for (int x=0; x<dw; ++x) { dst[(size_t)prevDy*dw+x] = sum[x]; sum[x] = beta*buf[x]; }
			prevDy = dy;
		} else {
			for (int x=0; x<dw; ++x) sum[x] += beta*buf[x];
		}
	}
	if (prevDy >= 0)
		for (int x=0; x<dw; ++x) dst[(size_t)prevDy*dw+x] = sum[x];
}
void resizeLinear(const float* src, int sw, int sh, float* dst, int dw, int dh) {
	const double sx = (double)sw/dw, sy = (double)sh/dh;
	std::vector<int> xo(dw), yo(dh); std::vector<float> xa(dw), ya(dh);
	auto tab = [](int ssize, int dsize, double scale, std::vector<int>& o, std::vector<float>& a) {
		for (int d=0; d<dsize; ++d) {
			float f = (float)((d+0.5)*scale-0.5);
			int s = (int)std::floor(f);
			f -= s;
			if (s < 0) { f = 0; s = 0; }
			if (s >= ssize-1) { f = 0; s = ssize-1; }
			o[d] = s; a[d] = f;
		}
	};
	tab(sw, dw, sx, xo, xa);
	tab(sh, dh, sy, yo, ya);
	for (int y=0; y<dh; ++y) {
		const int y0 = yo[y], y1 = std::min(y0+1, sh-1);
		const float fy = ya[y];
		for (int x=0; x<dw; ++x) {
			const int x0 = xo[x], x1 = std::min(x0+1, sw-1);
			const float fx = xa[x];
			const float r0 = src[(size_t)y0*sw+x0]*(1.f-fx) + src[(size_t)y0*sw+x1]*fx;
			const float r1 = src[(size_t)y1*sw+x0]*(1.f-fx) + src[(size_t)y1*sw+x1]*fx;
			dst[(size_t)y*dw+x] = r0*(1.f-fy) + r1*fy;
		}
	}
}
// ifx/ify: source/destination scale (1/fx for the factor form of cv::resize, sw/dw for the destination-size form)
void resizeNearest(const float* src, int sw, int sh, int ch, float* dst, int dw, int dh, double ifx, double ify) {
	for (int y=0; y<dh; ++y) {
		const int sy = std::min((int)std::floor(y*ify), sh-1);
		for (int x=0; x<dw; ++x) {
			const int sx = std::min((int)std::floor(x*ifx), sw-1);
			for (int c=0; c<ch; ++c)
				dst[((size_t)y*dw+x)*ch+c] = src[((size_t)sy*sw+sx)*ch+c];
		}
	}
}
void scaleK(const double K[9], int sw, int sh, int dw, int dh, double Ko[9]) {
	// Camera::ScaleK (Camera.h:160-173)
	const double sx = (double)dw/sw, sy = (double)dh/sh;
	Ko[0] = K[0]*sx; Ko[1] = K[1]*sx; Ko[2] = (K[2]+0.5)*sx-0.5;
	Ko[3] = 0;       Ko[4] = K[4]*sy; Ko[5] = (K[5]+0.5)*sy-0.5;
	Ko[6] = 0; Ko[7] = 0; Ko[8] = 1;
}
inline int cvRoundI(double v) { return (int)std::nearbyint(v); }

} // namespace

extern "C" {

void oracle_default_params(oracle_params* p) {
	p->nEstimationIters = 3; p->nEstimationGeometricIters = 2; p->nRandomIters = 6;
	p->fNCCThresholdKeep = 0.9f; p->fDescriptorMinMagnitudeThreshold = 0.02f;
	p->fRandomDepthRatio = 0.003f; p->fRandomAngle1Range = 16.f; p->fRandomAngle2Range = 10.f;
	p->fRandomSmoothDepth = 0.02f; p->fRandomSmoothNormal = 13.f; p->fRandomSmoothBonus = 0.93f;
	p->fEstimationGeometricWeight = 0.1f; p->nSubResolutionLevels = 2;
	p->schedule = 0; p->propagation = 4; p->seed = 1234u; p->threads = 1;
}

void oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }

int oracle_pm_score(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, const float* lowres, float* depth, float* normal, float* conf)
{
	if (nViews < 2) return 1;
	Job J; std::vector<WeightedPatch> weights;
	buildJob(views, nViews, prm, dMin, dMax, lowres, depth, normal, conf, weights, J);
	passA(J);
	return 0;
}

int oracle_pm_iterate(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, const float* lowres, int iter, int half, float* depth, float* normal, float* conf)
{
	if (nViews < 2) return 1;
	Job J; std::vector<WeightedPatch> weights;
	buildJob(views, nViews, prm, dMin, dMax, lowres, depth, normal, conf, weights, J);
	passB(J, iter, half);
	return 0;
}

/* oracle_pm_iterate with the memory of the RB changed-flag rule carried by the caller: changed = width x height bytes
 * (1 = plane changed in the pixel's last update / never tested), read and updated; NULL = no rule */
int oracle_pm_iterate_flags(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, const float* lowres, int iter, int half, float* depth, float* normal, float* conf, uint8_t* changed)
{
	if (nViews < 2) return 1;
	Job J; std::vector<WeightedPatch> weights;
	buildJob(views, nViews, prm, dMin, dMax, lowres, depth, normal, conf, weights, J);
	J.changed = changed;
	passB(J, iter, half);
	return 0;
}

int oracle_pm_finalize(int width, int height, float keep, float* depth, float* normal, float* conf) {
	passC(width, height, keep, depth, normal, conf);
	return 0;
}

// the scale loop of DepthMapsData::EstimateDepthMap (SceneDensify.cpp:640-805) around passes A, B (iterations
// [iterBegin, iterEnd)) and C; mask: ignore-mask of the reference view at full resolution or null
static int estimateImpl(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, bool geometric, int iterBegin, int iterEnd, const uint8_t* mask, float* depth, float* normal, float* conf)
{
	if (nViews < 2) return 1;
	const int W = views[0].width, H = views[0].height;
	const int totalScale = !geometric ? prm->nSubResolutionLevels : 0;
	std::vector<float> lowD, lowN; int lowW = 0, lowH = 0;
	std::vector<float> prior;
	for (int s = totalScale; s >= 0; --s) {
		const double scale = 1.0/(double)(1<<s);
		// ScaleDepthData (SceneDensify.cpp:578-601)
		std::vector<oracle_view> sv(views, views+nViews);
		std::vector<std::vector<float>> imgs(nViews), dms(nViews);
		if (s > 0) {
			for (int i=0; i<nViews; ++i) {
				const int dw = cvRoundI(views[i].width*scale), dh = cvRoundI(views[i].height*scale);
				imgs[i].resize((size_t)dw*dh);
				resizeArea(views[i].image, views[i].width, views[i].height, imgs[i].data(), dw, dh, 1.0/scale, 1.0/scale);
				sv[i].image = imgs[i].data(); sv[i].width = dw; sv[i].height = dh;
				scaleK(views[i].K, views[i].width, views[i].height, dw, dh, sv[i].K);
				if (views[i].depth) {
					dms[i].resize((size_t)dw*dh);
					resizeArea(views[i].depth, views[i].dwidth, views[i].dheight, dms[i].data(), dw, dh, (double)views[i].dwidth/dw, (double)views[i].dheight/dh);
					sv[i].depth = dms[i].data(); sv[i].dwidth = dw; sv[i].dheight = dh;
					scaleK(views[i].Kd, views[i].dwidth, views[i].dheight, dw, dh, sv[i].Kd);
				}
			}
		}
		const int w = sv[0].width, h = sv[0].height;
		std::vector<float> d((size_t)w*h), n((size_t)w*h*3), c((size_t)w*h);
		const float* lowres = nullptr;
		if (s != totalScale) {
			// depth LINEAR (NEAREST when an ignore-mask is set), normal NEAREST (SceneDensify.cpp:660-664)
			if (mask) resizeNearest(lowD.data(), lowW, lowH, 1, d.data(), w, h, (double)lowW/w, (double)lowH/h);
			else resizeLinear(lowD.data(), lowW, lowH, d.data(), w, h);
			resizeNearest(lowN.data(), lowW, lowH, 3, n.data(), w, h, (double)lowW/w, (double)lowH/h);
			prior = d;
			lowres = prior.data();
		} else if (s == 0) {
			memcpy(d.data(), depth, sizeof(float)*w*h);
			memcpy(n.data(), normal, sizeof(float)*w*h*3);
		} else {
			// coarsest level of a multi-scale run: the initial estimate is the caller's, scaled by nearest
			// neighbour with the factor form cv::resize(..., Size(), scale, scale, INTER_NEAREST) (ScaleDepthData)
			resizeNearest(depth, W, H, 1, d.data(), w, h, 1.0/scale, 1.0/scale);
			resizeNearest(normal, W, H, 3, n.data(), w, h, 1.0/scale, 1.0/scale);
		}
		Job J; std::vector<WeightedPatch> weights;
		buildJob(sv.data(), nViews, prm, dMin, dMax, lowres, d.data(), n.data(), c.data(), weights, J);
		std::vector<uint8_t> changed, lmask;
		if (prm->schedule == 1 && (prm->propagation & 0x100)) { changed.assign((size_t)w*h, 1); J.changed = changed.data(); }
		if (mask) {
			// ImportIgnoreMask: cv::resize(mask, size, INTER_NEAREST) (DepthMap.cpp:309), then ApplyIgnoreMask (DepthMap.cpp:215-230)
			lmask.resize((size_t)w*h);
			for (int y=0; y<h; ++y) for (int x=0; x<w; ++x) {
				const int sx = std::min((int)std::floor(x*((double)W/w)), W-1), sy = std::min((int)std::floor(y*((double)H/h)), H-1);
				lmask[(size_t)y*w+x] = mask[(size_t)sy*W+sx];
			}
			for (size_t i=0; i<(size_t)w*h; ++i)
				if (!lmask[i]) { d[i] = 0; n[i*3] = n[i*3+1] = n[i*3+2] = 0; c[i] = 0; }
			J.mask = lmask.data();
		}
		passA(J);
		for (int it=iterBegin; it<iterEnd; ++it)
			passB(J, it, -1);
		if (s > 0) { lowD = d; lowN = n; lowW = w; lowH = h; }
		else {
			memcpy(depth, d.data(), sizeof(float)*w*h);
			memcpy(normal, n.data(), sizeof(float)*w*h*3);
			memcpy(conf, c.data(), sizeof(float)*w*h);
		}
	}
	float keep = prm->fNCCThresholdKeep;
	if (!geometric && prm->nEstimationGeometricIters)
		keep *= 1.333f;
	passC(W, H, keep, depth, normal, conf);
	return 0;
}

int oracle_pm_estimate(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, int nGeometricIter, float* depth, float* normal, float* conf)
{
	const int iterBegin = nGeometricIter < 0 ? 0 : prm->nEstimationIters+nGeometricIter;
	const int iterEnd = nGeometricIter < 0 ? prm->nEstimationIters : iterBegin+1;
	return estimateImpl(views, nViews, prm, dMin, dMax, nGeometricIter >= 0, iterBegin, iterEnd, nullptr, depth, normal, conf);
}

/* the same with an explicit range of pass-B iterations (RB: sweeps) and an optional ignore-mask: the form the engine's
 * schedule maps to (photometric: sweeps [0, T); geometric pass g: sweeps [T + g*S, T + (g+1)*S), no scale loop) */
int oracle_pm_estimate_range(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, int geometric, int iterBegin, int iterEnd, const uint8_t* mask, float* depth, float* normal, float* conf)
{
	return estimateImpl(views, nViews, prm, dMin, dMax, geometric != 0, iterBegin, iterEnd, mask, depth, normal, conf);
}

/* experiment counters of the RB changed-flag rule: {directions tested, directions skipped} since the last call */
void oracle_pm_counters(long long out[2]) { out[0] = g_propTested.exchange(0); out[1] = g_propSkipped.exchange(0); }

float oracle_pm_score_pixel(const oracle_view* views, int nViews, const oracle_params* prm,
	float dMin, float dMax, const float* lowres, int x, int y,
	float depth, const float normal[3], const float* cl, int nClose, float* viewScores)
{
	Job J; std::vector<WeightedPatch> weights;
	std::vector<float> dummy;
	buildJob(views, nViews, prm, dMin, dMax, lowres, nullptr, nullptr, nullptr, weights, J);
	Estimator E(J, 0);
	if (!E.PreparePixelPatch(x, y) || !E.FillPixelPatch())
		return -1.f;
	E.nClose = nClose;
	for (int c=0; c<nClose; ++c) {
		E.close[c].depth = cl[c*7];
		E.close[c].normal = V3{cl[c*7+1], cl[c*7+2], cl[c*7+3]};
		E.close[c].X = V3{cl[c*7+4], cl[c*7+5], cl[c*7+6]};
	}
	const V3 n{normal[0], normal[1], normal[2]};
	E.InitPlane(depth, n);
	if (viewScores)
		for (size_t i=0; i<J.views.size(); ++i)
			viewScores[i] = E.ScorePixelImage(J.views[i], depth, n);
	return E.ScorePixel(depth, n);
}

void oracle_dir2normal(float a, float b, float n[3]) { V3 d; dir2normal(a, b, d); n[0]=d.x; n[1]=d.y; n[2]=d.z; }
void oracle_normal2dir(const float n[3], float* a, float* b) { normal2dir(V3{n[0],n[1],n[2]}, *a, *b); }
void oracle_correct_normal(const double X0[3], float n[3]) {
	Job J; J.w = J.h = 0;
	oracle_params p; oracle_default_params(&p); J.prm = p; J.dMin = 1; J.dMax = 2;
	Estimator E(J, 0);
	E.X0[0]=X0[0]; E.X0[1]=X0[1]; E.X0[2]=X0[2];
	V3 v{n[0],n[1],n[2]};
	E.CorrectNormal(v);
	n[0]=v.x; n[1]=v.y; n[2]=v.z;
}
float oracle_interpolate_pixel(const double K[9], int x0, int y0, int nx, int ny,
	float depth, const float normal[3], float dMin, float dMax)
{
	Job J; memcpy(J.K0, K, sizeof(J.K0)); J.dMin = dMin; J.dMax = dMax; J.w = J.h = 0;
	oracle_params p; oracle_default_params(&p); J.prm = p;
	Estimator E(J, 0);
	E.x0 = x0; E.y0 = y0;
	return E.InterpolatePixel(nx, ny, depth, V3{normal[0],normal[1],normal[2]});
}
void oracle_zigzag(int width, int height, int rawStride, uint16_t* coordsXY) {
	std::vector<uint16_t> c; zigzag(width, height, rawStride, c);
	memcpy(coordsXY, c.data(), c.size()*sizeof(uint16_t));
}
void oracle_resize_area(const float* src, int sw, int sh, float* dst, int dw, int dh, double scx, double scy) { resizeArea(src, sw, sh, dst, dw, dh, scx > 0 ? scx : (double)sw/dw, scy > 0 ? scy : (double)sh/dh); }
void oracle_resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh) { resizeLinear(src, sw, sh, dst, dw, dh); }
void oracle_resize_nearest(const float* src, int sw, int sh, int ch, float* dst, int dw, int dh) { resizeNearest(src, sw, sh, ch, dst, dw, dh, (double)sw/dw, (double)sh/dh); }
void oracle_scale_K(const double K[9], int sw, int sh, int dw, int dh, double Kout[9]) { scaleK(K, sw, sh, dw, dh, Kout); }

} // extern "C"
