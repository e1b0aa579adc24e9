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
// Three more members forward the per-view post-processing that follows the estimation
// (DepthMapsData::RemoveSmallSegments / GapInterpolation / FilterDepthMap,
// libs/MVS/SceneDensify.cpp:810-1299) to the same context.
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
			o.image8 = nullptr; o.channels8 = o.bgr8 = o.stride8_bytes = 0;
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

#ifdef _MVS_DEPTHMAP_H_
	// The reference's own signature, PatchMatchCUDA::EstimateDepthMap(DepthData&) (PatchMatchCUDA.inl:108, called at
	// SceneDensify.cpp:618-623): compiled when this header is included after libs/MVS/DepthMap.h.  It snapshots the
	// MVS::OPTDENSE globals itself, so the reference's call sites compile unchanged with `PatchMatchB200* pmCUDA`.
	// Like the reference's CUDA branch it does not receive the geometric iteration index; Init(true) selects the
	// geometric-consistency pass (iteration 0 of the engine's numbering: only the random stream depends on it).
	static OptDense SnapshotOPTDENSE() {
		OptDense o;
		o.nEstimationIters = (int)MVS::OPTDENSE::nEstimationIters;
		o.nEstimationGeometricIters = (int)MVS::OPTDENSE::nEstimationGeometricIters;
		o.nRandomIters = (int)MVS::OPTDENSE::nRandomIters;
		o.nSubResolutionLevels = (int)MVS::OPTDENSE::nSubResolutionLevels;
		o.fNCCThresholdKeep = MVS::OPTDENSE::fNCCThresholdKeep;
		o.fDescriptorMinMagnitudeThreshold = MVS::OPTDENSE::fDescriptorMinMagnitudeThreshold;
		o.fRandomDepthRatio = MVS::OPTDENSE::fRandomDepthRatio;
		o.fRandomAngle1Range = MVS::OPTDENSE::fRandomAngle1Range;
		o.fRandomAngle2Range = MVS::OPTDENSE::fRandomAngle2Range;
		o.fRandomSmoothDepth = MVS::OPTDENSE::fRandomSmoothDepth;
		o.fRandomSmoothNormal = MVS::OPTDENSE::fRandomSmoothNormal;
		o.fRandomSmoothBonus = MVS::OPTDENSE::fRandomSmoothBonus;
		o.fEstimationGeometricWeight = MVS::OPTDENSE::fEstimationGeometricWeight;
		return o;
	}
	void EstimateDepthMap(MVS::DepthData& depthData) { EstimateDepthMap(depthData, SnapshotOPTDENSE(), 0, nullptr); }
#endif

	// Ignore-mask of the reference view for the following calls (OPTDENSE::nIgnoreMaskLabel >= 0; DepthEstimator::ImportIgnoreMask
	// with pMask, libs/MVS/DepthMap.cpp:300-323, delivers it as an Image8U: 0 = ignored).  An empty mask clears it.
	template <typename MASK>
	void SetIgnoreMask(MASK& mask) {
		if (!ctx_) throw std::runtime_error("PatchMatchB200 used after Release()");
		if (mask.empty()) check(b200mvs_set_ignore_mask(ctx_, nullptr, 0, 0, 0, 0), "b200mvs_set_ignore_mask");
		else check(b200mvs_set_ignore_mask(ctx_, mask.template ptr<uint8_t>(), mask.cols, mask.rows, (int)mask.step[0], 0), "b200mvs_set_ignore_mask");
	}

	// DepthMapsData::RemoveSmallSegments(DepthData&) (SceneDensify.cpp:810-900); the OPTDENSE values are
	// passed raw (fDepthDiffThreshold, nSpeckleSize), maps are processed in place
	template <typename DEPTHDATA>
	void RemoveSmallSegments(DEPTHDATA& depthData, float fDepthDiffThreshold, unsigned nSpeckleSize) {
		if (!ctx_) throw std::runtime_error("PatchMatchB200 used after Release()");
		check(b200mvs_remove_small_segments(ctx_, depthData.depthMap.template ptr<float>(),
			depthData.normalMap.empty() ? nullptr : depthData.normalMap.template ptr<float>(),
			depthData.confMap.empty() ? nullptr : depthData.confMap.template ptr<float>(),
			depthData.depthMap.cols, depthData.depthMap.rows, fDepthDiffThreshold, nSpeckleSize, nullptr), "b200mvs_remove_small_segments");
	}
	// DepthMapsData::GapInterpolation(DepthData&) (SceneDensify.cpp:904-1045)
	template <typename DEPTHDATA>
	void GapInterpolation(DEPTHDATA& depthData, float fDepthDiffThreshold, unsigned nIpolGapSize) {
		if (!ctx_) throw std::runtime_error("PatchMatchB200 used after Release()");
		check(b200mvs_gap_interpolation(ctx_, depthData.depthMap.template ptr<float>(),
			depthData.normalMap.empty() ? nullptr : depthData.normalMap.template ptr<float>(),
			depthData.confMap.empty() ? nullptr : depthData.confMap.template ptr<float>(),
			depthData.depthMap.cols, depthData.depthMap.rows, fDepthDiffThreshold, nIpolGapSize, nullptr), "b200mvs_gap_interpolation");
	}
	// DepthMapsData::FilterDepthMap(depthDataRef, idxNeighbors, bAdjust) (SceneDensify.cpp:1050-1299): `neighbors` are the
	// DepthData of the valid neighbour views (arrDepthData[depthDataRef.neighbors[idx].ID]); newDepthMap / newConfMap are
	// what the reference saves as filtered.dmap / filtered.cmap.  Returns false when the map can not be filtered.
	template <typename DEPTHDATA, typename DMAP, typename CMAP>
	bool FilterDepthMap(DEPTHDATA& depthDataRef, const std::vector<DEPTHDATA*>& neighbors, const b200mvs_filter_params& prm,
		DMAP& newDepthMap, CMAP& newConfMap) {
		if (!ctx_) throw std::runtime_error("PatchMatchB200 used after Release()");
		std::vector<b200mvs_dmap> maps(neighbors.size()+1);
		for (size_t i = 0; i < maps.size(); ++i) {
			DEPTHDATA& d = i ? *neighbors[i-1] : depthDataRef;
			b200mvs_dmap& o = maps[i];
			o.depth = d.depthMap.template ptr<float>();
			o.conf = d.confMap.empty() ? nullptr : d.confMap.template ptr<float>();
			o.width = d.depthMap.cols; o.height = d.depthMap.rows;
			const auto& cam = d.images[0].camera;
			copy9(cam.K.val, o.K); copy9(cam.R.val, o.R); copy3(cam.C.ptr(), o.C);
		}
		newDepthMap.create(maps[0].height, maps[0].width);
		newConfMap.create(maps[0].height, maps[0].width);
		int filtered = 0;
		check(b200mvs_filter_depth_map(ctx_, &maps[0], maps.data()+1, (int)neighbors.size(), &prm, depthDataRef.dMin, depthDataRef.dMax,
			newDepthMap.template ptr<float>(), newConfMap.template ptr<float>(), &filtered, nullptr), "b200mvs_filter_depth_map");
		return filtered != 0;
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
