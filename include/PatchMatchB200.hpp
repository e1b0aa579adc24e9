// PatchMatchB200.hpp — C++ host-side mirror of the reference's accelerator seam.
//
// The reference owns a `PatchMatchCUDA* pmCUDA` in DepthMapsData (libs/MVS/SceneDensify.h:89-92)
// with four methods (libs/MVS/PatchMatchCUDA.inl:78-131):
//     PatchMatchCUDA(int device);  void Init(bool bGeomConsistency);  void Release();
//     void EstimateDepthMap(DepthData&);
// This header provides the same four methods on top of the C-ABI (include/b200mvs.h).  It is a
// template on the DepthData type so that it compiles both against the reference's
// MVS::DepthData (libs/MVS/DepthMap.h:157-271: images[i].image (cv::Mat-backed Image32F),
// images[i].camera.{K,R,C}, images[i].depthMap, images[i].cameraDepthMap, depthMap, normalMap,
// confMap, viewsMap, dMin, dMax) and against the minimal mock used by tests/test_cpp_adapter.py.
// Nothing here computes: it marshals pointers and OPTDENSE values into the C structs.
#pragma once
#include "b200mvs.h"
#include <stdexcept>
#include <string>
#include <vector>

namespace b200mvs {

// OPTDENSE values are globals in the reference (libs/MVS/DepthMap.cpp:69-114); the caller
// passes a snapshot taken at call time (several are mutated during a run).
struct OptDense : b200mvs_params {
	OptDense() { b200mvs_default_params(this); }
};

class PatchMatchB200 {
public:
	explicit PatchMatchB200(int device = 0) : ctx_(nullptr), geom_(false) {
		const int rc = b200mvs_create(device, &ctx_);
		if (rc != B200MVS_OK) // the reference exit()s on CUDA errors; the adapter throws -> EVT_FAIL
			throw std::runtime_error("b200mvs_create failed with status " + std::to_string(rc));
	}
	~PatchMatchB200() { Release(); }
	PatchMatchB200(const PatchMatchB200&) = delete;
	PatchMatchB200& operator=(const PatchMatchB200&) = delete;

	// PatchMatchCUDA::Init(bool bGeomConsistency) (PatchMatchCUDA.cpp:94-105)
	void Init(bool bGeomConsistency) { geom_ = bGeomConsistency; }
	// PatchMatchCUDA::Release() (PatchMatchCUDA.cpp:107-117)
	void Release() { if (ctx_) { b200mvs_destroy(ctx_); ctx_ = nullptr; } }

	// PatchMatchCUDA::EstimateDepthMap(DepthData&) (PatchMatchCUDA.cpp:174-416).
	// nGeometricIter: the iteration index DepthMapsData::EstimateDepthMap received
	// (SceneDensify.cpp:616); ignored unless Init(true) was called.
	template <typename DEPTHDATA>
	void EstimateDepthMap(DEPTHDATA& depthData, const OptDense& opt, int nGeometricIter = 0, b200mvs_stats* stats = nullptr) {
		if (!ctx_) throw std::runtime_error("PatchMatchB200 used after Release()");
		const int n = (int)depthData.images.size();
		std::vector<b200mvs_view> views(n);
		for (int i = 0; i < n; ++i) {
			auto& v = depthData.images[i];
			b200mvs_view& o = views[i];
			o.image = v.image.template ptr<float>();
			o.width = v.image.cols; o.height = v.image.rows; o.stride_bytes = (int)v.image.step[0];
			copy9(v.camera.K.val, o.K); copy9(v.camera.R.val, o.R); copy3(v.camera.C.ptr(), o.C);
			o.depth = nullptr; o.dwidth = o.dheight = o.dstride_bytes = 0;
			if (geom_ && i > 0 && !v.depthMap.empty()) {
				o.depth = v.depthMap.template ptr<float>();
				o.dwidth = v.depthMap.cols; o.dheight = v.depthMap.rows; o.dstride_bytes = (int)v.depthMap.step[0];
				copy9(v.cameraDepthMap.K.val, o.Kd); copy9(v.cameraDepthMap.R.val, o.Rd); copy3(v.cameraDepthMap.C.ptr(), o.Cd);
			}
		}
		// the engine allocates nothing on the host: maps are created here when empty
		// (PatchMatchCUDA.cpp:226-233), zero depth means "initialise randomly"
		const int w = views[0].width, h = views[0].height;
		if (depthData.depthMap.empty()) { depthData.depthMap.create(h, w); depthData.depthMap.setTo(0); }
		if (depthData.normalMap.empty()) { depthData.normalMap.create(h, w); depthData.normalMap.setTo(0); }
		if (depthData.confMap.empty()) depthData.confMap.create(h, w);
		if (depthData.viewsMap.empty()) depthData.viewsMap.create(h, w);
		check(b200mvs_set_params(ctx_, &opt), "b200mvs_set_params");
		check(b200mvs_estimate(ctx_, views.data(), n, depthData.dMin, depthData.dMax, geom_ ? nGeometricIter : -1,
			depthData.depthMap.template ptr<float>(), depthData.normalMap.template ptr<float>(),
			depthData.confMap.template ptr<float>(), depthData.viewsMap.template ptr<uint8_t>(), stats), "b200mvs_estimate");
	}

private:
	static void copy9(const double* s, double* d) { for (int i = 0; i < 9; ++i) d[i] = s[i]; }
	static void copy3(const double* s, double* d) { for (int i = 0; i < 3; ++i) d[i] = s[i]; }
	void check(int rc, const char* what) const {
		if (rc != B200MVS_OK)
			throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + b200mvs_last_error(ctx_));
	}
	b200mvs_ctx* ctx_;
	bool geom_;
};

} // namespace b200mvs
