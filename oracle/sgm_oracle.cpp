/*
 * sgm_oracle.cpp — CPU oracle of the SGM pair matcher (TEST INFRASTRUCTURE ONLY, see oracle.h).
 * PARITY UNPINNED (the reference has no SGM test or golden vector).
 *
 * Restates SemiGlobalMatcher::Match(leftImage, rightImage, disparityMap, costMap)
 * (libs/MVS/SemiGlobalMatcher.cpp:863-1302) with SGM_SIMILARITY_WZNCC, numDirs = 4 (8 paths),
 * 7x7 window, P1 = 3 and the adaptive P2 table GenerateP2s (:518-524):
 *   (1) cost     WZNCC over a 7x7 window with bilateral weights from the colour image -> uint8   :875-985
 *   (2) paths    L(d) = C(d) + min(Lp(d), Lp(d+-1)+P1, Lp(d')+P2) - min Lp, restricted to the
 *                intersection of the previous and current disparity ranges                      :1003-1046
 *                in the per-scanline order of the threaded variant                              :1048-1201
 *   (3) WTA      first arg-min of the summed path costs                                         :1272-1301
 * Integer work (2)-(3) is reproduced bit-exactly by the CUDA path; (1) rounds a float to uint8.
 */
#include "oracle.h"
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>

namespace {

constexpr int HW = 3;          // halfWindowSizeX/Y
constexpr int NT = 49;         // numTexels
constexpr int16_t NO_DISP = std::numeric_limits<int16_t>::max();
constexpr uint16_t NO_ACCUMCOST = std::numeric_limits<uint16_t>::max();

inline int round2int(float x) { return (int)std::floor(x+.5f); }

struct Ctx {
	const float* lgray; const uint8_t* lbgr; const float* rgray;
	int w, h, vw, vh;
	const oracle_sgm_pixel* px;
	uint8_t* costs; uint16_t* accums;
	uint16_t P1; uint16_t P2s[256];
	int maxNumDisp;
};

void computeCosts(Ctx& c) {
	const float eps = 1e-3f;
	const float sigmaColor = -1.f/(2.f*(0.3f*255)*(0.3f*255));
	const float sigmaSpatial = -1.f/(2.f*(0.4f*7)*(0.4f*7));
	for (int r=0; r<c.vh; ++r) for (int col=0; col<c.vw; ++col) {
		const oracle_sgm_pixel& p = c.px[(size_t)r*c.vw+col];
		if (!(p.dmin < p.dmax)) continue;
		const int ux = col+HW, uy = r+HW;
		float weight[NT], tempWeight[NT];
		float normSq0 = 0, sumWeights = 0;
		int n = 0;
		const uint8_t* cc = c.lbgr + ((size_t)uy*c.w+ux)*3;
		for (int i=-HW; i<=HW; ++i) for (int j=-HW; j<=HW; ++j) {
			const int x = ux+j, y = uy+i;
			const uint8_t* pc = c.lbgr + ((size_t)y*c.w+x)*3;
			unsigned s = 0;
			for (int k=0; k<3; ++k) { const unsigned d = pc[k] < cc[k] ? cc[k]-pc[k] : pc[k]-cc[k]; s += d*d; }
			const float wgt = std::exp(float(s)*sigmaColor + float(j*j+i*i)*sigmaSpatial);
			const float g = c.lgray[(size_t)y*c.w+x];
			tempWeight[n] = g; weight[n] = wgt;
			normSq0 += g*wgt;
			sumWeights += wgt;
			++n;
		}
		const float tm = normSq0/sumWeights;
		normSq0 = 0;
		for (n=0; n<NT; ++n) {
			const float t = tempWeight[n]-tm;
			tempWeight[n] = weight[n]*t;
			normSq0 += tempWeight[n]*t;
		}
		uint8_t* costs = c.costs + p.idx;
		for (int d=p.dmin; d<p.dmax; ++d) {
			float sum = 0, sumSq = 0, nom = 0;
			bool outside = false;
			n = 0;
			for (int i=-HW; i<=HW && !outside; ++i) for (int j=-HW; j<=HW; ++j) {
				const int x = ux+j+d, y = uy+i;
				if (x < 0 || y < 0 || x >= c.w || y >= c.h) { outside = true; break; }
				const float f = c.rgray[(size_t)y*c.w+x];
				const float fw = f*weight[n];
				sum += fw;
				sumSq += f*fw;
				nom += f*tempWeight[n];
				++n;
			}
			if (outside) { *costs++ = 255; continue; }
			const float normSq1 = sumSq - sum*sum/sumWeights;
			const float ncc = nom/std::sqrt(normSq0*normSq1+eps);
			*costs++ = ncc <= 0 ? (uint8_t)255 : (uint8_t)round2int((1.f-std::min(ncc, 1.f))*255.f);
		}
	}
}

struct Line { std::vector<uint16_t> L; int16_t dmin, dmax; };

// pixelAccum (SemiGlobalMatcher.cpp:1003-1046)
void pixelAccum(const Ctx& c, const uint8_t* costs, const Line& Lp, Line& Ls, uint16_t* accums, float DI) {
	const uint16_t P2 = c.P2s[std::abs(round2int(255.f*DI))];
	const int16_t minDisp = std::max(Lp.dmin, Ls.dmin), maxDisp = std::min(Lp.dmax, Ls.dmax);
	const int num = Ls.dmax-Ls.dmin;
	if (minDisp >= maxDisp) {
		for (int k=0; k<num; ++k)
			accums[k] += (Ls.L[k] = (uint16_t)(costs[k]+P2));
		return;
	}
	uint16_t minLp = std::numeric_limits<uint16_t>::max();
	for (int dp=minDisp; dp<maxDisp; ++dp)
		minLp = std::min(minLp, Lp.L[dp-Lp.dmin]);
	for (int d=Ls.dmin; d<Ls.dmax; ++d) {
		const int k = d-Ls.dmin;
		uint16_t L = std::numeric_limits<uint16_t>::max();
		for (int dp=minDisp; dp<maxDisp; ++dp) {
			const uint16_t lp = Lp.L[dp-Lp.dmin];
			uint16_t v;
			if (dp == d) v = lp;
			else if (dp == d-1 || dp == d+1) v = (uint16_t)(lp+c.P1);
			else v = (uint16_t)(lp+P2);
			if (L > v) L = v;
		}
		accums[k] += (Ls.L[k] = (uint16_t)(costs[k]+L-minLp));
	}
}

// one scanline of the threaded variant: ACCUM_PIXELS (SemiGlobalMatcher.cpp:1065-1083)
void walk(const Ctx& c, int x, int y, int dx, int dy) {
	Line lines[2];
	for (Line& l: lines) { l.L.assign(std::max(1, c.maxNumDisp), 0); l.dmin = l.dmax = 0; }
	int cur = 0;
	float Ip = 0.5f;
	for (; x >= 0 && y >= 0 && x < c.vw && y < c.vh; x += dx, y += dy) {
		const oracle_sgm_pixel& p = c.px[(size_t)y*c.vw+x];
		if (!(p.dmin < p.dmax)) continue;
		const Line& Lp = lines[cur]; Line& Ls = lines[cur^1];
		Ls.dmin = p.dmin; Ls.dmax = p.dmax;
		// NB: the reference indexes the gray image with the valid-region coordinates (no half-window offset)
		const float I = c.lgray[(size_t)y*c.w+x];
		pixelAccum(c, c.costs+p.idx, Lp, Ls, c.accums+p.idx, I-Ip);
		Ip = I;
		cur ^= 1;
	}
}

void aggregate(Ctx& c) {
	const int W = c.vw, H = c.vh;
	for (int x=0; x<W; ++x) walk(c, x, 0, 0, 1);        // width-down
	for (int y=0; y<H; ++y) walk(c, 0, y, 1, 0);        // height-right
	for (int x=0; x<W; ++x) walk(c, x, H-1, 0, -1);     // width-up
	for (int y=0; y<H; ++y) walk(c, W-1, y, -1, 0);     // height-left
	for (int x=0; x<W; ++x) walk(c, x, 0, 1, 1);        // right-down
	for (int y=1; y<H; ++y) walk(c, 0, y, 1, 1);
	for (int x=0; x<W-1; ++x) walk(c, x, 0, -1, 1);     // left-down
	for (int y=0; y<H; ++y) walk(c, W-1, y, -1, 1);
	for (int x=1; x<W; ++x) walk(c, x, H-1, 1, -1);     // right-up
	for (int y=H-1; y>=0; --y) walk(c, 0, y, 1, -1);
	for (int x=W-1; x>=0; --x) walk(c, x, H-1, -1, -1); // left-up
	for (int y=H-2; y>=0; --y) walk(c, W-1, y, -1, -1);
}

} // namespace

extern "C" {

void oracle_sgm_p2s(uint16_t P2, float alpha, float beta, uint16_t out[256]) {
	// GenerateP2s (SemiGlobalMatcher.cpp:518-524)
	for (int i=0; i<256; ++i)
		out[i] = (uint16_t)round2int(P2*(1.f+alpha*std::exp(-float(i)*float(i)/(2.f*beta*beta))));
}

// stage: 1 = costs only, 2 = costs + aggregation, 3 = + WTA.  costs/accums are caller buffers of
// numCosts elements (costs may be supplied pre-filled with stage bit 8 set: skip stage 1).
int oracle_sgm_match(const float* leftGray, const uint8_t* leftBGR, const float* rightGray, int width, int height,
	const oracle_sgm_pixel* pixels, uint64_t numCosts, uint16_t P1, uint16_t P2, float P2alpha, float P2beta, int stage,
	uint8_t* costs, uint16_t* accums, int16_t* disparity, uint16_t* cost)
{
	Ctx c;
	c.lgray = leftGray; c.lbgr = leftBGR; c.rgray = rightGray; c.w = width; c.h = height;
	c.vw = width-2*HW; c.vh = height-2*HW;
	if (c.vw <= 0 || c.vh <= 0) return 1;
	c.px = pixels; c.costs = costs; c.accums = accums; c.P1 = P1;
	oracle_sgm_p2s(P2, P2alpha, P2beta, c.P2s);
	c.maxNumDisp = 0;
	for (size_t i=0, n=(size_t)c.vw*c.vh; i<n; ++i)
		if (pixels[i].dmin < pixels[i].dmax)
			c.maxNumDisp = std::max(c.maxNumDisp, pixels[i].dmax-pixels[i].dmin);
	if (!(stage & 8))
		computeCosts(c);
	if ((stage & 7) >= 2) {
		memset(accums, 0, sizeof(uint16_t)*numCosts);
		aggregate(c);
	}
	if ((stage & 7) >= 3) {
		for (size_t i=0, n=(size_t)c.vw*c.vh; i<n; ++i) {
			const oracle_sgm_pixel& p = pixels[i];
			if (p.dmin < p.dmax) {
				const uint16_t* a = accums+p.idx;
				int best = 0;
				for (int k=1; k<p.dmax-p.dmin; ++k)
					if (a[best] > a[k]) best = k;
				disparity[i] = (int16_t)(p.dmin+best);
				cost[i] = a[best];
			} else {
				disparity[i] = p.dmin;
				cost[i] = NO_ACCUMCOST;
			}
		}
	}
	return 0;
}

// ConsistencyCrossCheck (SemiGlobalMatcher.cpp:1449-1489): l2r is modified in place
void oracle_sgm_cross_check(int16_t* l2r, const int16_t* r2l, int width, int height, int thCross) {
	for (int r=0; r<height; ++r) for (int c=0; c<width; ++c) {
		int16_t& ld = l2r[(size_t)r*width+c];
		if (ld == NO_DISP) continue;
		const int vx = c+ld;
		if (vx < 0 || vx >= width) { ld = NO_DISP; continue; }
		const int16_t rd = r2l[(size_t)r*width+vx];
		if (rd == NO_DISP) { ld = NO_DISP; continue; }
		if (std::abs(ld+rd) > thCross) ld = NO_DISP;
	}
}

// RefineDisparityMap, SUBPIXEL_LC_BLEND (SemiGlobalMatcher.cpp:1693-1811)
void oracle_sgm_refine(const oracle_sgm_pixel* pixels, const uint16_t* accums, int16_t* disparity, int nPixels, int subpixelSteps) {
	if (subpixelSteps <= 1) return;
	typedef float real;
	const real PI = 3.14159265358979323846f;
	auto linear = [](real x) { return x/real(2); };
	auto cosine = [&](real x) { return real(1)-std::cos(x*(real)(3.14159265358979323846/3.0)); };
	auto lcBlend = [&](real x) { const real factor = real(1.195)-std::cos(x*(real)(3.14159265358979323846/2.3)); return cosine(x)*factor + linear(x)*(real(1)-factor); };
	auto semi = [](uint16_t primary, uint16_t other) { return other == 0 ? real(0) : real(0.5)*((real)primary/(real)other); };
	(void)PI;
	for (int i=0; i<nPixels; ++i) {
		const oracle_sgm_pixel& p = pixels[i];
		if (p.dmax-p.dmin < 2) continue;
		int16_t& d = disparity[i];
		if (d == NO_DISP) continue;
		const uint16_t* a = accums+p.idx;
		const int k = d-p.dmin;
		real disp = (real)d;
		if (d == p.dmin) disp += semi(a[k], a[k+1]);
		else if (d+1 == p.dmax) disp -= semi(a[k], a[k-1]);
		else {
			const uint16_t prev = a[k-1], center = a[k], next = a[k+1];
			real off;
			if (prev == center) off = center == next ? real(0) : semi(center, next);
			else if (center == next) off = -semi(center, prev);
			else {
				const uint16_t ld = (uint16_t)(prev-center), rd = (uint16_t)(next-center);
				real x, mult;
				if (ld < rd) { x = (real)ld/(real)rd; mult = real(1); }
				else { x = (real)rd/(real)ld; mult = real(-1); }
				off = (lcBlend(x)-real(0.5))*mult;
			}
			disp += off;
		}
		d = (int16_t)round2int(disp*subpixelSteps);
	}
}

} // extern "C"
