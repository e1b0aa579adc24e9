// filter_oracle.cpp — CPU restatement of the depth-map post-processing that follows the
// estimation path (TEST INFRASTRUCTURE, see oracle.h):
//   DepthMapsData::FilterDepthMap       libs/MVS/SceneDensify.cpp:1050-1299
//   DepthMapsData::RemoveSmallSegments  libs/MVS/SceneDensify.cpp:810-900
//   DepthMapsData::GapInterpolation     libs/MVS/SceneDensify.cpp:904-1045
// Camera arithmetic follows libs/MVS/Camera.h:339-393 in double (Point3 = TPoint3<REAL>), with
// the evaluation order written out below (the reference leaves it to Eigen/OpenCV expression
// templates); rounding helpers are the non-_FAST_FLOAT2INT forms (libs/Common/Types.h:916-963,
// CMakeLists.txt:25 default OFF).  Compiled with -ffp-contract=off (oracle/Makefile).
#include "oracle.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace {

// IsDepthSimilar (libs/Common/Util.inl:797-809): |d0-d1|/d0 < threshold (not symmetric)
inline bool depth_similar(float d0, float d1, float th) { return std::fabs(d0-d1)/d0 < th; }

// Camera::TransformPointI2W(Point3(x,y,depth)) (Camera.h:339-356)
inline void i2w(const oracle_dmap& v, double x, double y, double z, double X[3]) {
	const double cx = (x-v.K[2])*z/v.K[0];
	const double cy = (y-v.K[5])*z/v.K[4];
	for (int i = 0; i < 3; ++i)
		X[i] = ((v.R[i]*cx + v.R[3+i]*cy) + v.R[6+i]*z) + v.C[i];
}
// Camera::TransformPointW2C (Camera.h:388-390)
inline void w2c(const oracle_dmap& v, const double X[3], double c[3]) {
	const double t0 = X[0]-v.C[0], t1 = X[1]-v.C[1], t2 = X[2]-v.C[2];
	for (int i = 0; i < 3; ++i)
		c[i] = (v.R[i*3]*t0 + v.R[i*3+1]*t1) + v.R[i*3+2]*t2;
}
// Camera::TransformPointC2I(Point3) (Camera.h:370-386)
inline void c2i(const oracle_dmap& v, const double c[3], double& u, double& w) {
	u = v.K[2]+v.K[0]*(c[0]/c[2]);
	w = v.K[5]+v.K[4]*(c[1]/c[2]);
}

// Normal2Dir / Dir2Normal (libs/Common/Util.inl:754-766)
inline void normal2dir(const float* n, float& a, float& b) { a = std::atan2(n[1], n[0]); b = std::acos(n[2]); }
inline void dir2normal(float a, float b, float* n) {
	const float siny = std::sin(b);
	n[0] = std::cos(a)*siny; n[1] = std::sin(a)*siny; n[2] = std::cos(b);
}

} // namespace

extern "C" {

// forward projection of one neighbour depth-map into the reference view, z-buffered over the
// 4 pixels around the projection (SceneDensify.cpp:1083-1131).  On equal depth the later
// source pixel (row-major scan) wins, as `depthRef < camX.z` lets equal values overwrite.
void oracle_filter_project(const oracle_dmap* ref, const oracle_dmap* nbr, float* projDepth, float* projConf) {
	const int W = ref->width, H = ref->height;
	memset(projDepth, 0, sizeof(float)*W*H);
	if (projConf) memset(projConf, 0, sizeof(float)*W*H);
	for (int i = 0; i < nbr->height; ++i) {
		for (int j = 0; j < nbr->width; ++j) {
			const float depth = nbr->depth[(size_t)i*nbr->width+j];
			if (depth == 0)
				continue;
			double X[3], c[3];
			i2w(*nbr, j, i, depth, X);
			w2c(*ref, X, c);
			if (c[2] <= 0)
				continue;
			double u, v;
			c2i(*ref, c, u, v);
			const double xs[2] = {std::floor(u), std::ceil(u)}, ys[2] = {std::floor(v), std::ceil(v)};
			const float z = (float)c[2];
			for (int p = 0; p < 4; ++p) { // (fx,fy) (fx,cy) (cx,fy) (cx,cy)
				const double px = xs[p>>1], py = ys[p&1];
				if (!(px >= 0 && py >= 0 && px < W && py < H))
					continue;
				const size_t o = (size_t)py*W+(size_t)px;
				if (projDepth[o] != 0 && projDepth[o] < z)
					continue;
				projDepth[o] = z;
				if (projConf) projConf[o] = nbr->conf[(size_t)i*nbr->width+j];
			}
		}
	}
}

int oracle_filter_depth_map(const oracle_dmap* ref, const oracle_dmap* nbrs, int N, int nMinViews, int nMinViewsAdjust,
	float fDepthDiffThreshold, int bAdjust, float dMin, float dMax, float* outDepth, float* outConf, float* projected)
{
	if (N < nMinViews || N < nMinViewsAdjust)
		return 0; // "depth map can not be filtered" (SceneDensify.cpp:1060-1063)
	const int W = ref->width, H = ref->height;
	const size_t P = (size_t)W*H;
	std::vector<float> own;
	float* depthMaps = projected;
	if (!depthMaps) { own.resize(P*N); depthMaps = own.data(); }
	std::vector<float> confMaps(bAdjust ? P*N : 0);
	for (int n = 0; n < N; ++n)
		oracle_filter_project(ref, nbrs+n, depthMaps+P*n, bAdjust ? confMaps.data()+P*n : nullptr);
	const float thDepthDiff = fDepthDiffThreshold*1.2f;
	if (bAdjust) {
		// average similar depths, lower the confidence where depths disagree (:1141-1210)
		for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) {
			const size_t o = (size_t)i*W+j;
			const float depth = ref->depth[o];
			outDepth[o] = 0; outConf[o] = 0;
			if (depth == 0)
				continue;
			float posConf = ref->conf[o], negConf = 0;
			float avgDepth = depth*posConf;
			unsigned nPosViews = 0, nNegViews = 0;
			unsigned n = (unsigned)N;
			bool discard = false;
			do {
				const float d = depthMaps[P*(--n)+o];
				if (d == 0) {
					if (nPosViews + nNegViews + n < (unsigned)nMinViews) { discard = true; break; }
					continue;
				}
				if (depth_similar(depth, d, thDepthDiff)) {
					const float c = confMaps[P*n+o];
					avgDepth += d*c;
					posConf += c;
					++nPosViews;
				} else {
					if (depth > d) {
						negConf += confMaps[P*n+o]; // occlusion
					} else {
						// free-space violation: confidence of the neighbour at the projection of this point
						const oracle_dmap& nb = nbrs[n];
						double X[3], c3[3], u, v;
						i2w(*ref, j, i, depth, X);
						w2c(nb, X, c3);
						c2i(nb, c3, u, v);
						const double rx = std::floor(u+.5), ry = std::floor(v+.5);
						if (rx >= 0 && ry >= 0 && rx < nb.width && ry < nb.height) {
							const float c = nb.conf[(size_t)ry*nb.width+(size_t)rx];
							negConf += (c > 0 ? c : confMaps[P*n+o]);
						} else
							negConf += confMaps[P*n+o];
					}
					++nNegViews;
				}
			} while (n);
			if (discard)
				continue;
			if (nPosViews >= (unsigned)nMinViewsAdjust && posConf > negConf) {
				avgDepth /= posConf;
				if (dMin <= avgDepth && avgDepth < dMax) {
					outDepth[o] = avgDepth;
					outConf[o] = posConf - negConf;
				}
			}
		}
	} else {
		// keep a depth only if enough neighbours agree with it (:1211-1289); positions outside
		// the image count as "no depth" (the reference reads out of bounds there)
		const float thStrict = fDepthDiffThreshold*0.8f;
		const unsigned nMinViewsDelta = (unsigned)nMinViews*2;
		static const int dx[4] = {-1,1,0,0}, dy[4] = {0,0,-1,1};
		for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) {
			const size_t o = (size_t)i*W+j;
			const float depth = ref->depth[o];
			outDepth[o] = 0; outConf[o] = 0;
			if (depth == 0)
				continue;
			unsigned good = 0, views = 0;
			for (int n = N; n-- > 0; ) {
				const float d = depthMaps[P*n+o];
				if (d > 0) { ++views; if (depth_similar(depth, d, thStrict)) ++good; }
			}
			if (good < (unsigned)nMinViews || good < views*75/100)
				continue;
			good = views = 0;
			for (int k = 0; k < 4; ++k) {
				const int x = j+dx[k], y = i+dy[k];
				if (x < 0 || y < 0 || x >= W || y >= H)
					continue;
				for (int n = N; n-- > 0; ) {
					const float d = depthMaps[P*n+(size_t)y*W+x];
					if (d > 0) { ++views; if (depth_similar(depth, d, thDepthDiff)) ++good; }
				}
			}
			if (good < nMinViewsDelta || good < views*65/100)
				continue;
			outDepth[o] = depth;
			outConf[o] = ref->conf[o];
		}
	}
	return 1;
}

// breadth-first segments over 4-neighbours whose depth is similar to the CURRENT pixel's
// (directed test), seeds scanned column by column; segments smaller than `speckle` are zeroed
void oracle_remove_small_segments(float* depth, float* normal, float* conf, int W, int H, float th, unsigned speckle) {
	std::vector<unsigned char> done((size_t)W*H, 0);
	std::vector<int> seg((size_t)W*H);
	for (int u = 0; u < W; ++u) for (int v = 0; v < H; ++v) {
		if (done[(size_t)v*W+u])
			continue;
		seg[0] = v*W+u;
		unsigned count = 1, curr = 0;
		while (curr < count) {
			const int a = seg[curr];
			const int ax = a%W, ay = a/W;
			const float dc = depth[a];
			if (dc > 0) {
				const int nx[4] = {ax-1, ax+1, ax, ax}, ny[4] = {ay, ay, ay-1, ay+1};
				for (int k = 0; k < 4; ++k) {
					if (nx[k] < 0 || ny[k] < 0 || nx[k] >= W || ny[k] >= H)
						continue;
					const int b = ny[k]*W+nx[k];
					if (done[b])
						continue;
					const float dn = depth[b];
					if (dn > 0 && depth_similar(dc, dn, th)) {
						seg[count++] = b;
						done[b] = 1;
					}
				}
			}
			++curr;
			done[a] = 1;
		}
		if (count < speckle) {
			for (unsigned k = 0; k < count; ++k) {
				const int a = seg[k];
				depth[a] = 0;
				if (normal) normal[a*3] = normal[a*3+1] = normal[a*3+2] = 0;
				if (conf) conf[a] = 0;
			}
		}
	}
}

// number of 4-neighbour pairs whose similarity test differs with the direction it is asked in
// (only those make the segments depend on the traversal order)
int oracle_count_asymmetric_edges(const float* depth, int W, int H, float th) {
	int n = 0;
	for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
		const float a = depth[(size_t)y*W+x];
		if (!(a > 0)) continue;
		if (x+1 < W) { const float b = depth[(size_t)y*W+x+1]; if (b > 0 && depth_similar(a,b,th) != depth_similar(b,a,th)) ++n; }
		if (y+1 < H) { const float b = depth[(size_t)(y+1)*W+x]; if (b > 0 && depth_similar(a,b,th) != depth_similar(b,a,th)) ++n; }
	}
	return n;
}

// linear interpolation of depth (and of the normal's direction angles) across gaps of at most
// `gap` invalid pixels between two similar valid ones; rows first, then columns on the result
void oracle_gap_interpolation(float* depth, float* normal, float* conf, int W, int H, float th, unsigned gap) {
	for (int pass = 0; pass < 2; ++pass) {
		const int nLines = pass == 0 ? H : W, len = pass == 0 ? W : H;
		for (int l = 0; l < nLines; ++l) {
			auto at = [&](int k) -> size_t { return pass == 0 ? (size_t)l*W+k : (size_t)k*W+l; };
			unsigned count = 0;
			for (int u = 0; u < len; ++u) {
				const float d1 = depth[at(u)];
				if (d1 <= 0) { ++count; continue; }
				if (count == 0)
					continue;
				if (count <= gap && (unsigned)u > count) {
					int uc = u-(int)count;
					const int uf = uc-1;
					const float d0 = depth[at(uf)];
					if (depth_similar(d0, d1, th)) {
						const float diff = (d1-d0)/(float)(count+1);
						float d = d0;
						const float c = conf ? (conf[at(uf)] < conf[at(u)] ? conf[at(uf)] : conf[at(u)]) : 0.f; // MINF
						float a1 = 0, b1 = 0, a2 = 0, b2 = 0, da = 0, db = 0;
						if (normal) {
							normal2dir(normal+at(uf)*3, a1, b1);
							normal2dir(normal+at(u)*3, a2, b2);
							da = (a2-a1)/(float)(count+1); db = (b2-b1)/(float)(count+1);
						}
						do {
							depth[at(uc)] = (d += diff);
							if (normal) { a1 += da; b1 += db; dir2normal(a1, b1, normal+at(uc)*3); }
							if (conf) conf[at(uc)] = c;
						} while (++uc < u);
					}
				}
				count = 0;
			}
		}
	}
}

// TImage<Pixel8U>::toGray(out, COLOR_BGR2GRAY / COLOR_RGB2GRAY, bNormalize = true) (libs/Common/Types.inl:2377-2431):
// channels scaled by float(1)/float(255) (NormRGB_t, Types.inl:1610-1615), gray = (cb*c0 + cg*c1) + cr*c2 in float
void oracle_to_gray(const uint8_t* src, int width, int height, int stride, int channels, int bgr, float* dst) {
	const float inv = 1.f/255.f;
	const float k0 = bgr ? 0.114f : 0.299f, k1 = 0.587f, k2 = bgr ? 0.299f : 0.114f;
	for (int y = 0; y < height; ++y) for (int x = 0; x < width; ++x) {
		const uint8_t* p = src + (size_t)y*stride + (size_t)x*channels;
		const float c0 = (float)p[0]*inv, c1 = (float)p[1]*inv, c2 = (float)p[2]*inv;
		dst[(size_t)y*width+x] = (k0*c0 + k1*c1) + k2*c2;
	}
}

} // extern "C"
