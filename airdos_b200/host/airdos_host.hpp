// airdos_host.hpp -- header-only C++ host side above the C-ABI (include/airdos_b200.h).
//
// Mirrors the three classes of the reference that own the hot path, with the same names, method
// names, argument meaning and error behaviour, so that the reference's own translation units
// (src/Frame.cc, src/Tracking.cc, src/LocalMapping.cc) keep compiling against them:
//   ORB_SLAM2::ORBextractor   include/ORBextractor.h:46-112
//   ORB_SLAM2::ORBmatcher     include/ORBmatcher.h:37-110  (the Hamming primitives)
//   ORB_SLAM2::Optimizer      include/Optimizer.h:38-69    (LocalBundleAdjustment*)
// OpenCV / Eigen are not available in this image, so cv::Mat / cv::KeyPoint are stood in for by
// airdos::ImageView / adb_keypoint; INTEGRATION.md shows the two-line adapters for the real types.
// No computation happens here: every method forwards to libairdos_b200.so and throws
// std::runtime_error on a non-OK status only where the reference would have asserted.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/airdos_b200.h"

namespace airdos {

struct ImageView {   // stands in for a CV_8UC1 cv::Mat header
    const uint8_t* data = nullptr;
    int cols = 0, rows = 0, step = 0;
    bool empty() const { return data == nullptr || cols == 0 || rows == 0; }
};

inline void check(adb_status s, const char* what) {
    if (s != ADB_OK) throw std::runtime_error(std::string(what) + ": " + adb_last_error());
}

}  // namespace airdos

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // include/ORBextractor.h:51-52, the reference's own signature: source compatible with src/Tracking.cc:160-163.  The device
    // buffers are provisioned by the first operator() from the size of its image (and again if the size ever changes).
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        : ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, 0, 0, 0) {}
    // the same with the image size known up front (nothing is allocated inside operator()) and a device ordinal
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int width, int height, int device = 0)
        : nfeatures_(nfeatures), scaleFactor_(scaleFactor) {
        adb_orb_config cfg{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height, 1, device};
        airdos::check(adb_orb_create(&cfg, &h_), "ORBextractor");
        const int nl = adb_orb_levels(h_);
        mvScaleFactor.resize(nl); mvInvScaleFactor.resize(nl); mvLevelSigma2.resize(nl); mvInvLevelSigma2.resize(nl);
        mnFeaturesPerLevel.resize(nl); levelW_.resize(nl); levelH_.resize(nl);
        mvImagePyramid.resize(nl);
        refresh_levels();
    }
    ~ORBextractor() { adb_orb_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // include/ORBextractor.h:59-61.  Empty image: silent return (src/ORBextractor.cc:1057-1058).
    void operator()(const airdos::ImageView& image, const airdos::ImageView& mask, std::vector<adb_keypoint>& keypoints,
                    std::vector<uint8_t>& descriptors) {
        keypoints.clear(); descriptors.clear();
        if (image.empty()) return;
        const int cap = adb_orb_capacity(h_);
        keypoints.resize(cap); descriptors.resize((size_t)cap * 32);
        int32_t n = 0;
        airdos::check(adb_orb_extract(h_, image.data, image.cols, image.rows, image.step, mask.empty() ? nullptr : mask.data, mask.step,
                                      keypoints.data(), descriptors.data(), cap, &n), "ORBextractor::operator()");
        keypoints.resize(n); descriptors.resize((size_t)n * 32);
        pyramid_valid_ = false;
        if (levelW_[0] != image.cols || levelH_[0] != image.rows) refresh_levels();   // lazily provisioned handle: level sizes are known now
    }

    int GetLevels() const { return (int)mvScaleFactor.size(); }
    float GetScaleFactor() const { return scaleFactor_; }
    std::vector<float> GetScaleFactors() const { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() const { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() const { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() const { return mvInvLevelSigma2; }

    // public member of the reference (include/ORBextractor.h:86), read by Frame::ComputeStereoMatches;
    // fetched from the device on first use after an extraction.
    std::vector<std::vector<uint8_t>> mvImagePyramid;
    const std::vector<std::vector<uint8_t>>& ImagePyramid() {
        if (!pyramid_valid_) {
            for (int l = 0; l < GetLevels(); ++l) {
                mvImagePyramid[l].resize((size_t)levelW_[l] * levelH_[l]);
                airdos::check(adb_orb_get_pyramid(h_, 0, l, 0, mvImagePyramid[l].data(), levelW_[l]), "mvImagePyramid");
            }
            pyramid_valid_ = true;
        }
        return mvImagePyramid;
    }
    int LevelWidth(int l) const { return levelW_[l]; }
    int LevelHeight(int l) const { return levelH_[l]; }
    adb_orb_t handle() const { return h_; }

    std::vector<int> mnFeaturesPerLevel;

protected:
    void refresh_levels() {
        for (int l = 0; l < (int)mvScaleFactor.size(); ++l) {
            int32_t w, h, p, q;
            adb_orb_level_info(h_, l, &w, &h, &p, &mvScaleFactor[l], &mvInvScaleFactor[l], &mvLevelSigma2[l], &mvInvLevelSigma2[l], &q);
            mnFeaturesPerLevel[l] = q; levelW_[l] = w; levelH_[l] = h;
        }
    }
    int nfeatures_;
    float scaleFactor_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> levelW_, levelH_;
    bool pyramid_valid_ = false;
    adb_orb_t h_ = nullptr;
};

class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;   // src/ORBmatcher.cc:37-39

    ORBmatcher(float nnratio = 0.6f, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
        airdos::check(adb_matcher_create(device, &m_), "ORBmatcher");
    }
    ~ORBmatcher() { adb_matcher_destroy(m_); }
    ORBmatcher(const ORBmatcher&) = delete;
    ORBmatcher& operator=(const ORBmatcher&) = delete;

    // include/ORBmatcher.h:44: descriptors are 32-byte rows
    static int DescriptorDistance(const uint8_t* a, const uint8_t* b) { return adb_hamming_distance(a, b); }

    // The scan shared by every Search*/Fuse: best / second-best over candidate lists (CSR), first in list wins ties.
    void BestTwo(const uint8_t* queries, int nq, const uint8_t* targets, int nt, const std::vector<int32_t>* cand_off,
                 const std::vector<int32_t>* cand_idx, std::vector<int32_t>& best_idx, std::vector<int32_t>& best_d, std::vector<int32_t>& second_d) {
        best_idx.assign(nq, -1); best_d.assign(nq, 256); second_d.assign(nq, 256);
        if (nq == 0) return;
        airdos::check(adb_match_best2(m_, queries, nq, targets, nt, cand_off ? cand_off->data() : nullptr, cand_idx ? cand_idx->data() : nullptr,
                                      best_idx.data(), best_d.data(), second_d.data()), "ORBmatcher::BestTwo");
    }

    // Frame::ComputeStereoMatches (src/Frame.cc:829-1003) for the frame resident in the two extractors.
    static void ComputeStereoMatches(ORBextractor& left, ORBextractor& right, int nLeft, float mb, float mbf, std::vector<float>& mvuRight,
                                     std::vector<float>& mvDepth) {
        const int cap = adb_orb_capacity(left.handle());
        mvuRight.assign(cap, -1.f); mvDepth.assign(cap, -1.f);
        airdos::check(adb_stereo_match(left.handle(), right.handle(), 1, mb, mbf, mvuRight.data(), mvDepth.data(), nullptr, nullptr, cap),
                      "ComputeStereoMatches");
        mvuRight.resize(nLeft); mvDepth.resize(nLeft);
    }

    // SearchByProjection(Frame& F, const vector<MapPoint*>&, th) / SearchByProjection(Frame& Current, const Frame& Last, th,
    // bMono) (src/ORBmatcher.cc:45-129, 1328-1470) after the shim has copied the Frame / MapPoint members into an
    // adb_proj_search.  mfNNratio / mbCheckOrientation come from this matcher like in the
    // reference; returns nmatches, the new CurrentFrame.mvpMapPoints are in problem.kp_match.
    int SearchByProjection(adb_proj_search& problem) {
        if (problem.last_xw) { problem.use_ratio = 0; problem.check_orientation = mbCheckOrientation ? 1 : 0; }
        else { problem.use_ratio = 1; problem.nn_ratio = mfNNratio; problem.check_orientation = 0; }
        airdos::check(adb_search_by_projection(m_, &problem, 1), "ORBmatcher::SearchByProjection");
        return problem.n_matches;
    }

    // SearchByBoW(pKF, F, vpMapPointMatches) (mode 0) / SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)
    // (mode 1) (src/ORBmatcher.cc:159-288, 657-823) after the shim's FeatureVector walk has filled the per-node lists.
    int SearchByBoW(adb_bow_search& problem) {
        problem.nn_ratio = mfNNratio; problem.check_orientation = mbCheckOrientation ? 1 : 0;
        airdos::check(adb_search_by_bow(m_, &problem, 1), "ORBmatcher::SearchByBoW");
        return problem.n_matches;
    }

    // SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (src/ORBmatcher.cc:657-823): the mode-1 form of the call above
    // by its reference name; vMatchedPairs = {(i, problem.match12[i]) : match12[i] >= 0}.
    int SearchForTriangulation(adb_bow_search& problem) {
        problem.mode = 1;
        return SearchByBoW(problem);
    }

    // Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:825-975), search half: the shim fills the KeyFrame side and the mp_* arrays of an
    // adb_proj_search; q_best_idx[i] >= 0 && q_best_dist[i] <= TH_LOW says "fuse map point i into key-point q_best_idx[i]"; the Replace /
    // AddObservation bookkeeping of :948-968 stays with the caller.  Returns nFused.
    int Fuse(adb_proj_search& problem, float th = 3.0f) {
        problem.fuse = 1; problem.th = th; problem.use_ratio = 0; problem.check_orientation = 0;
        airdos::check(adb_search_by_projection(m_, &problem, 1), "ORBmatcher::Fuse");
        return problem.n_matches;
    }

    // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:245-310) for a batch of map points: observations' descriptors row-wise in
    // `descriptors`, point p owning rows point_ptr[p] .. point_ptr[p + 1]); best_idx[p] = winning row within the point (-1: no observations,
    // the caller keeps mDescriptor), best_desc = the winning 32 bytes per point.
    void ComputeDistinctiveDescriptors(const uint8_t* descriptors, const std::vector<int32_t>& point_ptr, std::vector<int32_t>& best_idx,
                                       std::vector<uint8_t>& best_desc) {
        const int n_points = point_ptr.empty() ? 0 : (int)point_ptr.size() - 1;
        best_idx.assign(n_points, -1); best_desc.assign((size_t)n_points * 32, 0);
        if (n_points == 0) return;
        airdos::check(adb_distinctive_descriptors(m_, descriptors, point_ptr.data(), n_points, best_idx.data(), best_desc.data()),
                      "MapPoint::ComputeDistinctiveDescriptors");
    }

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    adb_matcher_t m_ = nullptr;
};

class Optimizer {
public:
    // Optimizer::LocalBundleAdjustment(KeyFrame*, bool* pbStopFlag, Map*) after the shim has flattened the
    // local window into an adb_ba_problem.  Returns false when the stop flag was already set
    // (the reference returns early without touching the map, src/Optimizer.cc:620-622).
    static bool LocalBundleAdjustment(adb_ba_problem& problem, bool* pbStopFlag, adb_ba_result& result, const adb_ba_options* options = nullptr,
                                      int device = 0) {
        static_assert(sizeof(bool) == 1, "bool* is passed as the C-ABI's uint8_t stop flag");
        adb_ba_t s = nullptr;
        airdos::check(adb_ba_create(device, &s), "Optimizer");
        adb_ba_options o;
        if (options) o = *options; else adb_ba_default_options(&o);
        const adb_status st = adb_ba_solve(s, &problem, &o, reinterpret_cast<volatile const uint8_t*>(pbStopFlag), &result);
        adb_ba_destroy(s);
        if (st == ADB_ERR_STOPPED) return false;
        airdos::check(st, "Optimizer::LocalBundleAdjustment");
        return true;
    }
    // int Optimizer::PoseOptimization(Frame* pFrame) (src/Optimizer.cc:232-429), batched: the shim fills one adb_pose_problem for the
    // frames it wants optimised (per frame: Tcw as quaternion + translation, the map points' world positions, observations, invSigma2);
    // outlier[] = mvbOutlier, the poses are updated in place.  Returns the reference's return value (nInitialCorrespondences - nBad) of the
    // first frame; problem.n_inliers holds it for every frame.
    static int PoseOptimization(adb_pose_problem& problem, int device = 0) {
        adb_ba_t s = nullptr;
        airdos::check(adb_ba_create(device, &s), "Optimizer");
        const adb_status st = adb_pose_optimize(s, &problem);
        adb_ba_destroy(s);
        airdos::check(st, "Optimizer::PoseOptimization");
        return problem.n_frames > 0 ? problem.n_inliers[0] : 0;
    }
    // Optimizer::GlobalBundleAdjustemnt(pMap, nIterations, pbStopFlag, nLoopKF, bRobust) / BundleAdjustment
    // (src/Optimizer.cc:52-230): all key-frames and map points, one round, no gating.
    static bool GlobalBundleAdjustemnt(adb_ba_problem& problem, int nIterations, bool* pbStopFlag, bool bRobust, adb_ba_result& result,
                                       int device = 0) {
        adb_ba_options o;
        adb_ba_global_options(&o, nIterations, bRobust ? 1 : 0);
        return LocalBundleAdjustment(problem, pbStopFlag, result, &o, device);
    }
    // LocalBundleAdjustmentHumanTrajactory(pKF, pbStopFlag, pMap, SigmaStatic, SigmaHuman, SigmaRigidity, SigmaMotion,
    // thRanSacMotion, thRanSacRigidity): the sigmas are the *_info arrays of the problem, the thresholds the options.
    static bool LocalBundleAdjustmentHumanTrajactory(adb_ba_problem& problem, bool* pbStopFlag, adb_ba_result& result, float thRanSacMotion,
                                                     float thRanSacRigidity, int device = 0) {
        adb_ba_options o;
        adb_ba_default_options(&o);
        o.chi2_motion = thRanSacMotion; o.chi2_rigid = thRanSacRigidity;
        o.huber_rigid = thRanSacRigidity;                                 // src/Optimizer.cc:1882 (not rooted)
        o.huber_motion = (double)(float)std::sqrt(thRanSacMotion);         // src/Optimizer.cc:1506
        return LocalBundleAdjustment(problem, pbStopFlag, result, &o, device);
    }
};

}  // namespace ORB_SLAM2
