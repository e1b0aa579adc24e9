// fuse_host.cu — DepthMapsData::FuseDepthMaps behind the C-ABI.  HOST code.
//
// What it computes (libs/MVS/SceneDensify.cpp:1372-1646): all valid depth-maps are fused into one point cloud, the best
// connected images first and every pixel in raster order; a pixel that is not yet part of a point creates one, projects it
// into the neighbour depth-maps, merges the neighbour pixels whose depth (1 % relative) and normal (25 deg) agree — position,
// colour and normal are confidence-weighted means — drops the point when fewer than nMinViewsFuse views agree, and zeroes the
// neighbour depths that lie behind an accepted point (they would block its view).
//
// Why on the host: the result depends on the order — a neighbour pixel claimed by an earlier point is not available to a later
// one, and a zeroed depth changes what later pixels see — so the reference's single sequential loop IS the specification
// (north_star keeps fusion on the host; §8(f) ranks it last for that reason).  The maps come straight from the estimation
// (b200mvs_estimate* / the filter calls); this is the host step after them, written for throughput: flat arrays, the
// projection matrices composed once, no per-point allocations (views / weights live in one pool).
// The CPU restatement used by the tests is oracle/fuse_oracle.py (independent, plain Python loops).
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <new>
#include "../../include/b200mvs.h"

namespace {

constexpr uint32_t NO_ID = 0xFFFFFFFFu;

struct FuseCam {
	double K[9], R[9], C[3], P[12];
};
// AssembleProjectionMatrix (libs/MVS/Camera.cpp:173-180): P = [K R | -K R C]
void compose(const b200mvs_fuse_view& v, FuseCam& c) {
	memcpy(c.K, v.K, sizeof(c.K)); memcpy(c.R, v.R, sizeof(c.R)); memcpy(c.C, v.C, sizeof(c.C));
	double M[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			M[3*i+j] = c.K[3*i]*c.R[j] + c.K[3*i+1]*c.R[3+j] + c.K[3*i+2]*c.R[6+j];
	for (int i = 0; i < 3; ++i) {
		c.P[4*i] = M[3*i]; c.P[4*i+1] = M[3*i+1]; c.P[4*i+2] = M[3*i+2];
		c.P[4*i+3] = M[3*i]*(-c.C[0]) + M[3*i+1]*(-c.C[1]) + M[3*i+2]*(-c.C[2]);
	}
}
// Camera::TransformPointI2W(Point3(x, y, depth)) (libs/MVS/Camera.h:339-356), double
inline void image_to_world(const FuseCam& c, int x, int y, float depth, double out[3]) {
	const double z = (double)depth;
	const double cx = ((double)x-c.K[2])*z/c.K[0], cy = ((double)y-c.K[5])*z/c.K[4];
	out[0] = c.R[0]*cx + c.R[3]*cy + c.R[6]*z + c.C[0];
	out[1] = c.R[1]*cx + c.R[4]*cy + c.R[7]*z + c.C[1];
	out[2] = c.R[2]*cx + c.R[5]*cy + c.R[8]*z + c.C[2];
}
// R^T n: camera-space normal of the depth-map to world space (SceneDensify.cpp:1526,1554), double sums, float result
inline void normal_to_world(const FuseCam& c, const float* n, float out[3]) {
	const double nx = n[0], ny = n[1], nz = n[2];
	out[0] = (float)(c.R[0]*nx + c.R[3]*ny + c.R[6]*nz);
	out[1] = (float)(c.R[1]*nx + c.R[4]*ny + c.R[7]*nz);
	out[2] = (float)(c.R[2]*nx + c.R[5]*ny + c.R[8]*nz);
}
// Conf2Weight (SceneDensify.cpp:120-122)
inline float conf2weight(float conf, float depth) { return 1.f/(fmaxf(1.f-conf, 0.03f)*depth*depth); }

} // namespace

struct b200mvs_pointcloud {
	std::vector<float> points, normals, weights;
	std::vector<uint8_t> colors;
	std::vector<uint32_t> views, offsets;   // views / weights of point i: [offsets[i], offsets[i+1])
	std::vector<uint16_t> projs;            // pixel (x, y) of every view of every point, parallel to `views`
	uint64_t nDepths = 0;
};

extern "C" {

void b200mvs_fuse_default_params(b200mvs_fuse_params* p) {
	p->nMinViewsFuse = 2; p->fDepthDiffThreshold = 0.01f; p->fNormalDiffThreshold = 25.f; p->bEstimateColor = 1; p->bEstimateNormal = 1;
}

int b200mvs_fuse_depth_maps(b200mvs_fuse_view* views, int nViews, const b200mvs_fuse_params* prm, b200mvs_pointcloud** out) {
	if (!views || nViews <= 0 || !out) return B200MVS_ERR_ARG;
	b200mvs_fuse_params def; b200mvs_fuse_default_params(&def);
	if (!prm) prm = &def;
	for (int i = 0; i < nViews; ++i) {
		const b200mvs_fuse_view& v = views[i];
		if (v.depth && (v.width <= 0 || v.height <= 0 || v.width > 65535 || v.height > 65535)) return B200MVS_ERR_ARG;
		for (int k = 0; k < v.nNeighbors; ++k)
			if (!v.neighbors || v.neighbors[k] >= (uint32_t)nViews || v.neighbors[k] == (uint32_t)i) return B200MVS_ERR_ARG;
	}
	b200mvs_pointcloud* pc = new (std::nothrow) b200mvs_pointcloud;
	if (!pc) return B200MVS_ERR_CUDA;
	std::vector<FuseCam> cams(nViews);
	for (int i = 0; i < nViews; ++i) compose(views[i], cams[i]);
	// best connected images first (SceneDensify.cpp:1392-1449); a view is valid when it has a depth-map with a depth in it.
	// std::sort's order among equal scores is unspecified in the reference: ties go to the lower index here.
	struct Conn { int idx; float score; };
	std::vector<Conn> conns;
	bool bNormalMap = true;
	for (int i = 0; i < nViews; ++i) {
		const b200mvs_fuse_view& v = views[i];
		if (!v.depth || v.nSceneNeighbors <= 0) continue;
		bool any = false;
		for (size_t k = 0, n = (size_t)v.width*v.height; k < n && !any; ++k) any = v.depth[k] > 0;
		if (!any) continue;
		conns.push_back(Conn{i, (float)v.nSceneNeighbors});
		if (!v.normal) bNormalMap = false;
	}
	std::stable_sort(conns.begin(), conns.end(), [](const Conn& a, const Conn& b) { return a.score > b.score; });
	const bool bColor = prm->bEstimateColor != 0;
	const bool bNormal = prm->bEstimateNormal != 0 && bNormalMap;
	const unsigned nMinViewsFuse = (unsigned)std::min(prm->nMinViewsFuse, nViews);
	const float normalError = cosf(prm->fNormalDiffThreshold*0.017453292519943295f);
	const float thDepth = prm->fDepthDiffThreshold;
	std::vector<std::vector<uint32_t>> idxMaps(nViews);
	pc->offsets.push_back(0);
	// scratch of the point under construction (a point has at most 1 + nNeighbors views)
	std::vector<uint32_t> pv; std::vector<float> pw; std::vector<uint16_t> pp; std::vector<float*> invalid;
	for (const Conn& cn: conns) {
		const int a = cn.idx;
		b200mvs_fuse_view& A = views[a];
		const FuseCam& camA = cams[a];
		for (int k = 0; k < A.nNeighbors; ++k) {
			const uint32_t b = A.neighbors[k];
			if (idxMaps[b].empty() && views[b].depth) idxMaps[b].assign((size_t)views[b].width*views[b].height, NO_ID);
		}
		if (idxMaps[a].empty()) idxMaps[a].assign((size_t)A.width*A.height, NO_ID);
		uint32_t* idxA = idxMaps[a].data();
		for (int y = 0; y < A.height; ++y) {
			for (int x = 0; x < A.width; ++x) {
				const size_t ia = (size_t)y*A.width + x;
				const float depth = A.depth[ia];
				if (depth == 0) continue;
				++pc->nDepths;
				if (idxA[ia] != NO_ID) continue;
				const uint32_t idxPoint = (uint32_t)(pc->offsets.size()-1);
				idxA[ia] = idxPoint;
				double Xw[3];
				image_to_world(camA, x, y, depth, Xw);
				const float point[3] = {(float)Xw[0], (float)Xw[1], (float)Xw[2]};
				pv.assign(1, (uint32_t)a); pp.assign(2, 0); pp[0] = (uint16_t)x; pp[1] = (uint16_t)y;
				pw.assign(1, conf2weight(A.conf ? A.conf[ia] : 1.f, depth));
				double confidence = (double)pw[0];
				float normal[3] = {0.f, 0.f, -1.f};
				if (bNormalMap) normal_to_world(camA, A.normal+3*ia, normal);
				double X[3] = {(double)point[0]*confidence, (double)point[1]*confidence, (double)point[2]*confidence};
				float Cc[3] = {0.f, 0.f, 0.f};
				if (bColor && A.color) for (int c = 0; c < 3; ++c) Cc[c] = (float)A.color[3*ia+c]*(float)confidence;
				float N[3] = {normal[0]*(float)confidence, normal[1]*(float)confidence, normal[2]*(float)confidence};
				invalid.clear();
				for (int k = 0; k < A.nNeighbors; ++k) {
					const uint32_t b = A.neighbors[k];
					b200mvs_fuse_view& B = views[b];
					if (!B.depth) continue;
					const FuseCam& camB = cams[b];
					// Camera::ProjectPointP3<float> (Camera.h:308-314): double sums, float components
					const float px = (float)(camB.P[0]*point[0] + camB.P[1]*point[1] + camB.P[2]*point[2] + camB.P[3]);
					const float py = (float)(camB.P[4]*point[0] + camB.P[5]*point[1] + camB.P[6]*point[2] + camB.P[7]);
					const float pz = (float)(camB.P[8]*point[0] + camB.P[9]*point[1] + camB.P[10]*point[2] + camB.P[11]);
					if (pz <= 0) continue;
					const int xb = (int)floorf(px/pz+.5f), yb = (int)floorf(py/pz+.5f);   // ROUND2INT (libs/Common/Types.h:947-953)
					if (xb < 0 || yb < 0 || xb >= B.width || yb >= B.height) continue;
					const size_t ib = (size_t)yb*B.width + xb;
					float& depthB = B.depth[ib];
					if (depthB == 0) continue;
					uint32_t& idxPointB = idxMaps[b][ib];
					if (idxPointB != NO_ID) continue;
					if (fabsf(pz-depthB)/pz < thDepth) {       // IsDepthSimilar(pt.z, depthB) (libs/Common/Util.inl:797-809)
						float normalB[3] = {0.f, 0.f, -1.f};
						if (bNormalMap) normal_to_world(camB, B.normal+3*ib, normalB);
						if (normal[0]*normalB[0] + normal[1]*normalB[1] + normal[2]*normalB[2] > normalError) {
							const float confB = conf2weight(B.conf ? B.conf[ib] : 1.f, depthB);
							size_t pos = 0;                      // views stay sorted by image index (InsertSort)
							while (pos < pv.size() && pv[pos] < b) ++pos;
							pv.insert(pv.begin()+pos, b); pw.insert(pw.begin()+pos, confB);
							pp.insert(pp.begin()+2*pos, (uint16_t)yb); pp.insert(pp.begin()+2*pos, (uint16_t)xb);
							idxPointB = idxPoint;
							double Xb[3];
							image_to_world(camB, xb, yb, depthB, Xb);
							X[0] += Xb[0]*(double)confB; X[1] += Xb[1]*(double)confB; X[2] += Xb[2]*(double)confB;
							if (bColor && B.color) for (int c = 0; c < 3; ++c) Cc[c] += (float)B.color[3*ib+c]*confB;
							if (bNormal) for (int c = 0; c < 3; ++c) N[c] += normalB[c]*confB;
							confidence += (double)confB;
							continue;
						}
					}
					if (pz < depthB) invalid.push_back(&depthB);   // in front of the neighbour's estimate: that depth blocks the view
				}
				if (pv.size() < nMinViewsFuse) {
					for (size_t v = 0; v < pv.size(); ++v)
						idxMaps[pv[v]][(size_t)pp[2*v+1]*views[pv[v]].width + pp[2*v]] = NO_ID;
					continue;
				}
				const double nrm = 1.0/confidence;
				for (int c = 0; c < 3; ++c) pc->points.push_back((float)(X[c]*nrm));
				if (bColor && A.color) for (int c = 0; c < 3; ++c) pc->colors.push_back((uint8_t)(Cc[c]*(float)nrm));
				if (bNormal) {
					const float n0 = N[0]*(float)nrm, n1 = N[1]*(float)nrm, n2 = N[2]*(float)nrm;
					const float len = sqrtf(n0*n0 + n1*n1 + n2*n2);
					pc->normals.push_back(n0/len); pc->normals.push_back(n1/len); pc->normals.push_back(n2/len);
				}
				pc->views.insert(pc->views.end(), pv.begin(), pv.end());
				pc->weights.insert(pc->weights.end(), pw.begin(), pw.end());
				pc->projs.insert(pc->projs.end(), pp.begin(), pp.end());
				pc->offsets.push_back((uint32_t)pc->views.size());
				for (float* d: invalid) *d = 0;
			}
		}
	}
	*out = pc;
	return B200MVS_OK;
}

uint64_t b200mvs_pointcloud_size(const b200mvs_pointcloud* pc) { return pc ? pc->offsets.size()-1 : 0; }
uint64_t b200mvs_pointcloud_depths(const b200mvs_pointcloud* pc) { return pc ? pc->nDepths : 0; }
const float* b200mvs_pointcloud_points(const b200mvs_pointcloud* pc) { return pc && !pc->points.empty() ? pc->points.data() : nullptr; }
const float* b200mvs_pointcloud_normals(const b200mvs_pointcloud* pc) { return pc && !pc->normals.empty() ? pc->normals.data() : nullptr; }
const uint8_t* b200mvs_pointcloud_colors(const b200mvs_pointcloud* pc) { return pc && !pc->colors.empty() ? pc->colors.data() : nullptr; }
const uint32_t* b200mvs_pointcloud_view_offsets(const b200mvs_pointcloud* pc) { return pc ? pc->offsets.data() : nullptr; }
const uint32_t* b200mvs_pointcloud_views(const b200mvs_pointcloud* pc) { return pc && !pc->views.empty() ? pc->views.data() : nullptr; }
const float* b200mvs_pointcloud_weights(const b200mvs_pointcloud* pc) { return pc && !pc->weights.empty() ? pc->weights.data() : nullptr; }
const uint16_t* b200mvs_pointcloud_projs(const b200mvs_pointcloud* pc) { return pc && !pc->projs.empty() ? pc->projs.data() : nullptr; }
void b200mvs_pointcloud_free(b200mvs_pointcloud* pc) { delete pc; }

} // extern "C"
