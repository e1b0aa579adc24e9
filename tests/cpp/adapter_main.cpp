// Drives include/PatchMatchB200.hpp with a minimal stand-in for MVS::DepthData (the reference's
// type needs OpenCV/Eigen, absent here).  Reads a scene dumped by tests/test_cpp_adapter.py,
// runs EstimateDepthMap through the adapter and writes the maps back.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <algorithm>

// the slice of cv::Mat the adapter touches
template <typename T, int CH = 1>
struct Mat {
	std::vector<T> buf; int cols = 0, rows = 0; size_t step[1] = {0};
	bool empty() const { return buf.empty(); }
	void create(int h, int w) { rows = h; cols = w; step[0] = sizeof(T)*CH*w; buf.assign((size_t)h*w*CH, T()); }
	void setTo(T v) { std::fill(buf.begin(), buf.end(), v); }
	template <typename U> U* ptr() { return reinterpret_cast<U*>(buf.data()); }
};
struct Mat3 { double val[9]; };
struct Pt3 { double v[3]; const double* ptr() const { return v; } };
struct Camera { Mat3 K, R; Pt3 C; };
struct ViewData { Mat<float> image; Camera camera; Mat<float> depthMap; Camera cameraDepthMap; };
// the reference's names, so that the adapter's PatchMatchCUDA-signature overload EstimateDepthMap(MVS::DepthData&) compiles as it
// would after `#include "DepthMap.h"`: MVS::DepthData and the MVS::OPTDENSE globals it snapshots (libs/MVS/DepthMap.h:88-142)
#define _MVS_DEPTHMAP_H_
namespace MVS {
struct DepthData {
	std::vector<ViewData> images;
	Mat<float> depthMap; Mat<float, 3> normalMap; Mat<float> confMap; Mat<uint8_t, 4> viewsMap;
	float dMin, dMax;
};
namespace OPTDENSE {
unsigned nEstimationIters = 3, nEstimationGeometricIters = 2, nRandomIters = 6, nSubResolutionLevels = 2;
float fNCCThresholdKeep = 0.9f, fDescriptorMinMagnitudeThreshold = 0.02f, fRandomDepthRatio = 0.003f, fRandomAngle1Range = 16.f,
	fRandomAngle2Range = 10.f, fRandomSmoothDepth = 0.02f, fRandomSmoothNormal = 13.f, fRandomSmoothBonus = 0.93f, fEstimationGeometricWeight = 0.1f;
}
}
using MVS::DepthData;
#include "../../include/PatchMatchB200.hpp"

int main(int argc, char** argv) {
	if (argc < 3) { fprintf(stderr, "usage: adapter_main scene.bin out.bin [iters]\n"); return 64; }
	DepthData dd;
	FILE* f = fopen(argv[1], "rb");
	if (!f) return 65;
	int hdr[3];
	if (fread(hdr, sizeof(int), 3, f) != 3) return 66;
	if (fread(&dd.dMin, sizeof(float), 1, f) != 1 || fread(&dd.dMax, sizeof(float), 1, f) != 1) return 66;
	dd.images.resize(hdr[0]);
	for (auto& v: dd.images) {
		if (fread(v.camera.K.val, 8, 9, f) != 9 || fread(v.camera.R.val, 8, 9, f) != 9 || fread(v.camera.C.v, 8, 3, f) != 3) return 66;
		v.image.create(hdr[2], hdr[1]);
		if (fread(v.image.buf.data(), 4, v.image.buf.size(), f) != v.image.buf.size()) return 66;
	}
	fclose(f);
	try {
		b200mvs::PatchMatchB200 pm(0);
		pm.Init(false);
		b200mvs::OptDense opt;
		opt.nSubResolutionLevels = 0; opt.nEstimationGeometricIters = 0;
		opt.nEstimationIters = argc > 3 ? atoi(argv[3]) : 2;
		b200mvs_stats st;
		if (argc > 4 && !strcmp(argv[4], "seam")) {
			// the reference's call site, unchanged: pmCUDA->EstimateDepthMap(depthData) with the OPTDENSE globals
			MVS::OPTDENSE::nSubResolutionLevels = 0; MVS::OPTDENSE::nEstimationGeometricIters = 0;
			MVS::OPTDENSE::nEstimationIters = (unsigned)opt.nEstimationIters;
			pm.EstimateDepthMap(dd);
			memset(&st, 0, sizeof(st));
		} else
			pm.EstimateDepthMap(dd, opt, 0, &st);
		printf("adapter: %dx%d, %d launches, %.2f ms device\n", hdr[1], hdr[2], st.kernel_launches, st.ms_device);
		if (argc > 4 && !strcmp(argv[4], "post")) {
			// the post-processing members: speckles, gaps, then the filter with the map as its own two neighbours
			pm.RemoveSmallSegments(dd, 0.01f, 100);
			pm.GapInterpolation(dd, 0.01f, 7);
			b200mvs_filter_params fp;
			b200mvs_filter_default_params(&fp);
			std::vector<DepthData*> nb = {&dd, &dd};
			Mat<float> nd, nc;
			if (!pm.FilterDepthMap(dd, nb, fp, nd, nc)) { printf("adapter error: filter refused\n"); return 2; }
			dd.depthMap = nd; dd.confMap = nc;
		}
		pm.Release();
	} catch (const std::exception& e) {
		printf("adapter error: %s\n", e.what());
		return strstr(e.what(), "status 3") ? 3 : 2;
	}
	f = fopen(argv[2], "wb");
	fwrite(dd.depthMap.buf.data(), 4, dd.depthMap.buf.size(), f);
	fwrite(dd.normalMap.buf.data(), 4, dd.normalMap.buf.size(), f);
	fwrite(dd.confMap.buf.data(), 4, dd.confMap.buf.size(), f);
	fwrite(dd.viewsMap.buf.data(), 1, dd.viewsMap.buf.size(), f);
	fclose(f);
	return 0;
}
