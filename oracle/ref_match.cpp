// ref_match.cpp -- TEST INFRASTRUCTURE: the reference's own matcher-side functions, compiled from /root/reference.
//   src/Frame.cc       ComputeStereoMatches, AssignFeaturesToGrid, GetFeaturesInArea, PosInGrid, UpdatePoseMatrices, isInFrustum
//   src/ORBmatcher.cc  DescriptorDistance, RadiusByViewingCos, ComputeThreeMaxima, CheckDistEpipolarLine, SearchByProjection(Frame&,
//                      vector<MapPoint*>&, th), SearchByProjection(Frame&, const Frame&, th, bMono), SearchByBoW(KeyFrame*, Frame&, ...),
//                      SearchForTriangulation, Fuse(KeyFrame*, vector<MapPoint*>&, th)
//   src/MapPoint.cc    GetMinDistanceInvariance, GetMaxDistanceInvariance, PredictScale (both overloads), ComputeDistinctiveDescriptors
//   src/KeyFrame.cc    GetFeaturesInArea, IsInImage
// -- every one the whole function definition, unmodified.  The files cannot be compiled as they are (OpenCV, Eigen, DBoW2, Pangolin
// headers are absent and the rest of each file needs them), so the build step (oracle/Makefile, target _ref/libref_match.so) copies the
// text of exactly these definitions out of the reference tree into oracle/_ref/match_snippets.inc (git-ignored, never committed:
// oracle/extract_ref_fn.py) and this file compiles that text between stand-in declarations of the classes they are members of: the
// members the statements read and write, with the reference's names and types, oracle/ref_shim/cv_shim.h for the handful of cv::Mat /
// cv::KeyPoint operations, std::map for DBoW2::FeatureVector.  The map surgery at the end of Fuse is recorded, not performed.
// oracle/gen_ref_{match,search,bow}_golden.py run the extern "C" wrappers below on seeded inputs and write tests/golden/
// {stereo,search,bow}_ref.npz, which pin oracle/match_oracle.cpp (and through it, and directly, the CUDA kernels) to the literal reference.
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "ref_shim/cv_shim.h"

namespace ORB_SLAM2 {
using namespace std;   // src/Frame.cc and src/ORBmatcher.cc both open with it

class Frame;
class KeyFrame;
class MapPoint {                         // include/MapPoint.h: what the matcher reads of a map point
public:
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }         // src/MapPoint.cc:79-83
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }     // src/MapPoint.cc:312-316
    int Observations() { return nObs; }                         // src/MapPoint.cc:134-138
    bool isBad() { return mbBad; }
    cv::Mat GetNormal() { return mNormalVector.clone(); }       // src/MapPoint.cc:85-89
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, Frame* pF);
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    void ComputeDistinctiveDescriptors();
    // the map surgery at the end of ORBmatcher::Fuse is only RECORDED (it belongs to the host side): which key-point this point was fused with
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }   // src/MapPoint.cc:220-224
    void AddObservation(KeyFrame* pKF, size_t idx) { mObservations[pKF] = idx; fused_with = (int)idx; }
    void Replace(MapPoint* pMP);
    int fused_with = -1;
    // variables used by the tracking (include/MapPoint.h:87-94)
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    // stand-in state
    cv::Mat mWorldPos, mDescriptor, mNormalVector;
    float mfMinDistance = 0, mfMaxDistance = 0;                 // include/MapPoint.h:141-142
    std::mutex mMutexPos, mMutexFeatures;
    std::map<KeyFrame*, size_t> mObservations;                  // include/MapPoint.h:111
    int nObs = 0;
    bool mbBad = false;
    int id = -1;                          // index of the query this point stands for (-100: a point the frame held on entry)
};

class ORBmatcher {                       // include/ORBmatcher.h:38-93
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}   // src/ORBmatcher.cc:41-43
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo);
    int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF);
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};
const int ORBmatcher::TH_HIGH = 100;     // src/ORBmatcher.cc:37
const int ORBmatcher::TH_LOW = 50;       // src/ORBmatcher.cc:38
const int ORBmatcher::HISTO_LENGTH = 30; // src/ORBmatcher.cc:39

class ORBextractor {                     // include/ORBextractor.h:86
public:
    std::vector<cv::Mat> mvImagePyramid;
};

}  // namespace ORB_SLAM2
namespace DBoW2 { typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector; }   // Thirdparty/DBoW2/DBoW2/FeatureVector.h:25-26
namespace ORB_SLAM2 {

#define FRAME_GRID_ROWS 48               // include/Frame.h:41-42
#define FRAME_GRID_COLS 64

class Frame {                            // include/Frame.h: the members the extracted functions touch, with the reference's names
public:
    void ComputeStereoMatches();
    void UpdatePoseMatrices();
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
    void AssignFeaturesToGrid();
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    static float fx, fy, cx, cy;
    float mbf, mb;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mTcw;
    cv::Mat mRcw, mtcw, mRwc, mOw;
    float mfLogScaleFactor;
    int mnScaleLevels;
    vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
class KeyFrame {                         // include/KeyFrame.h: what the two vocabulary-bucket searches read of a key-frame
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }          // src/KeyFrame.cc:258-262
    MapPoint* GetMapPoint(const size_t& idx) { last_asked = (int)idx; return mvpMapPoints[idx]; }   // src/KeyFrame.cc:264-268
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    bool isBad() { return false; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }   // src/KeyFrame.cc:218-222
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const;
    float mbf, mfLogScaleFactor;
    int mnScaleLevels;
    std::vector<float> mvInvLevelSigma2;
    int mnGridCols, mnGridRows;
    float mfGridElementWidthInv, mfGridElementHeightInv;
    int mnMinX, mnMinY, mnMaxX, mnMaxY;                             // include/KeyFrame.h:187-190 (const int there)
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    int last_asked = -1;                                            // stand-in: the key-point index of the last GetMapPoint call
    float fx, fy, cx, cy;
    int N;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvLevelSigma2;
    // stand-in state
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Ow, Rcw, tcw;
};
static KeyFrame* g_fuse_kf = nullptr;   // the key-frame of the running Fuse call (for the surgery record)
// pMP->Replace(pMPinKF) / pMPinKF->Replace(pMP): either way the NEW point (id >= 0) and the key-point just asked for are fused
void MapPoint::Replace(MapPoint* pMP) {
    MapPoint* incoming = id >= 0 && fused_with < 0 ? this : pMP;
    incoming->fused_with = g_fuse_kf->last_asked;
}
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

#include "_ref/match_snippets.inc"

}  // namespace ORB_SLAM2

struct RefKp { float x, y, size, angle, response; int octave; };   // = adb_keypoint

extern "C" {

// same argument layout as match_oracle_stereo (oracle/match_oracle.cpp): flattened pyramids, per-level offsets / sizes
void ref_stereo_match(const RefKp* kl, const uint8_t* dl, int nl_kp, const RefKp* kr, const uint8_t* dr, int nr_kp, const uint8_t* pyr_l,
                      const uint8_t* pyr_r, const long long* off, const int* lw, const int* lh, int nlevels, const float* scale, const float* inv_scale,
                      float mb, float mbf, float* u_right, float* depth) {
    using namespace ORB_SLAM2;
    ORBextractor exl, exr;
    for (int l = 0; l < nlevels; ++l) {
        exl.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, pyr_l + off[l]));
        exr.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, pyr_r + off[l]));
    }
    Frame f;
    f.mpORBextractorLeft = &exl; f.mpORBextractorRight = &exr;
    f.mb = mb; f.mbf = mbf; f.N = nl_kp;
    auto to_cv = [](const RefKp* k, int n) {
        std::vector<cv::KeyPoint> v(n);
        for (int i = 0; i < n; ++i) { v[i].pt.x = k[i].x; v[i].pt.y = k[i].y; v[i].size = k[i].size; v[i].angle = k[i].angle; v[i].response = k[i].response; v[i].octave = k[i].octave; }
        return v;
    };
    f.mvKeys = to_cv(kl, nl_kp); f.mvKeysRight = to_cv(kr, nr_kp);
    f.mDescriptors = cv::Mat(nl_kp, 32, CV_8U, dl); f.mDescriptorsRight = cv::Mat(nr_kp, 32, CV_8U, dr);
    f.mvScaleFactors.assign(scale, scale + nlevels); f.mvInvScaleFactors.assign(inv_scale, inv_scale + nlevels);
    f.ComputeStereoMatches();
    for (int i = 0; i < nl_kp; ++i) { u_right[i] = f.mvuRight[i]; depth[i] = f.mvDepth[i]; }
}

// Both SearchByProjection variants of the tracking thread on the arrays of the oracle's problem dicts (oracle/match_oracle.cpp):
// the current frame = key-points (mvKeys = mvKeysUn), mvuRight, descriptors, `taken` (mvpMapPoints set on entry, Observations() > 0),
// image bounds; the grid is built by the reference's own AssignFeaturesToGrid.  kp_match[i] = index of the query whose map point
// key-point i holds at the end, -1 = none (or only the one it held on entry).
namespace {
struct CurFrame {
    ORB_SLAM2::Frame F;
    std::unique_ptr<ORB_SLAM2::MapPoint[]> held;       // (a MapPoint owns a mutex: not movable)
    CurFrame(const RefKp* kps, const float* u_right, const uint8_t* desc, const uint8_t* taken, int n_kp, float minX, float minY, float maxX,
             float maxY, const float* scale, int nlevels) {
        using namespace ORB_SLAM2;
        F.N = n_kp;
        F.mvKeysUn.resize(n_kp);
        for (int i = 0; i < n_kp; ++i) { F.mvKeysUn[i].pt.x = kps[i].x; F.mvKeysUn[i].pt.y = kps[i].y; F.mvKeysUn[i].angle = kps[i].angle; F.mvKeysUn[i].octave = kps[i].octave; }
        F.mvKeys = F.mvKeysUn;
        F.mvuRight.assign(u_right, u_right + n_kp);
        F.mDescriptors = cv::Mat(n_kp, 32, CV_8U, desc);
        F.mvScaleFactors.assign(scale, scale + nlevels); F.mnScaleLevels = nlevels;
        Frame::mnMinX = minX; Frame::mnMinY = minY; Frame::mnMaxX = maxX; Frame::mnMaxY = maxY;
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);     // src/Frame.cc:113-114
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        held.reset(new MapPoint[n_kp]);
        F.mvpMapPoints.assign(n_kp, static_cast<MapPoint*>(nullptr));
        for (int i = 0; i < n_kp; ++i)
            if (taken && taken[i]) { held[i].nObs = 1; held[i].id = -100; F.mvpMapPoints[i] = &held[i]; }
        F.AssignFeaturesToGrid();
    }
    void result(int32_t* kp_match) const {
        for (int i = 0; i < F.N; ++i) kp_match[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->id >= 0 ? F.mvpMapPoints[i]->id : -1;
    }
};
}  // namespace

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)   src/ORBmatcher.cc:1328-1470
int ref_search_last_frame(const RefKp* kps, const float* u_right, const uint8_t* desc, const uint8_t* taken, int n_kp, float minX, float minY, float maxX,
                          float maxY, const float* scale, int nlevels, const float* tcw_cur16, const float* tcw_last16, int n_q, const float* last_xw,
                          const int32_t* last_octave, const float* last_angle, const uint8_t* last_desc, const uint8_t* last_flags, float fx, float fy,
                          float cx, float cy, float mbf, float mb, float th, int mono, int check_ori, int32_t* kp_match) {
    using namespace ORB_SLAM2;
    CurFrame cur(kps, u_right, desc, taken, n_kp, minX, minY, maxX, maxY, scale, nlevels);
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
    cur.F.mbf = mbf; cur.F.mb = mb; cur.F.mTcw = cv::Mat(4, 4, CV_32F, tcw_cur16);
    Frame last;
    last.N = n_q; last.mTcw = cv::Mat(4, 4, CV_32F, tcw_last16);
    last.mvKeys.resize(n_q); last.mvKeysUn.resize(n_q);
    std::vector<MapPoint> mps(n_q);
    last.mvpMapPoints.assign(n_q, static_cast<MapPoint*>(nullptr)); last.mvbOutlier.assign(n_q, false);
    for (int i = 0; i < n_q; ++i) {
        last.mvKeys[i].octave = last_octave[i]; last.mvKeysUn[i].angle = last_angle[i];
        mps[i].id = i; mps[i].nObs = (last_flags[i] & 2) ? 1 : 0;
        mps[i].mWorldPos = cv::Mat(3, 1, CV_32F, last_xw + 3 * i); mps[i].mDescriptor = cv::Mat(1, 32, CV_8U, last_desc + 32 * i);
        if (last_flags[i] & 1) last.mvpMapPoints[i] = &mps[i];     // bit 0: a map point that is not an outlier
    }
    ORBmatcher matcher(0.9, check_ori != 0);                        // src/Tracking.cc:933
    const int n = matcher.SearchByProjection(cur.F, last, th, mono != 0);
    cur.result(kp_match);
    return n;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th)   src/ORBmatcher.cc:45-129
int ref_search_map_points(const RefKp* kps, const float* u_right, const uint8_t* desc, const uint8_t* taken, int n_kp, float minX, float minY, float maxX,
                          float maxY, const float* scale, int nlevels, int n_q, const float* proj_x, const float* proj_y, const float* proj_xr,
                          const int32_t* level, const float* view_cos, const uint8_t* q_desc, const uint8_t* q_flags, float th, float nn_ratio,
                          int32_t* kp_match) {
    using namespace ORB_SLAM2;
    CurFrame cur(kps, u_right, desc, taken, n_kp, minX, minY, maxX, maxY, scale, nlevels);
    std::vector<MapPoint> mps(n_q);
    std::vector<MapPoint*> vp(n_q);
    for (int i = 0; i < n_q; ++i) {
        mps[i].id = i; mps[i].nObs = (q_flags[i] & 2) ? 1 : 0;
        mps[i].mbTrackInView = (q_flags[i] & 1) != 0;
        mps[i].mTrackProjX = proj_x[i]; mps[i].mTrackProjY = proj_y[i]; mps[i].mTrackProjXR = proj_xr[i];
        mps[i].mnTrackScaleLevel = level[i]; mps[i].mTrackViewCos = view_cos[i];
        mps[i].mDescriptor = cv::Mat(1, 32, CV_8U, q_desc + 32 * i);
        vp[i] = &mps[i];
    }
    ORBmatcher matcher(nn_ratio, true);
    const int n = matcher.SearchByProjection(cur.F, vp, th);
    cur.result(kp_match);
    return n;
}

// Tracking::SearchLocalPoints' hot part (src/Tracking.cc:1319-1343): Frame::isInFrustum(pMP, 0.5) with MapPoint::PredictScale for every
// local map point that reaches the test (flags bit 0), then SearchByProjection(F, vpMapPoints, th).  The camera centre is the one the
// reference's own UpdatePoseMatrices derives from Tcw (returned in ow_out so that the oracle can be given the same value).
int ref_search_local_map(const RefKp* kps, const float* u_right, const uint8_t* desc, const uint8_t* taken, int n_kp, float minX, float minY, float maxX,
                         float maxY, const float* scale, int nlevels, const float* tcw16, int n_q, const float* xw, const float* normal,
                         const float* min_distance, const float* max_distance, const uint8_t* q_desc, const uint8_t* q_flags, float fx, float fy, float cx,
                         float cy, float mbf, float view_cos_limit, float log_scale_factor, float th, float nn_ratio, int32_t* kp_match, uint8_t* in_view,
                         float* track4, int32_t* level, float* ow_out) {
    using namespace ORB_SLAM2;
    CurFrame cur(kps, u_right, desc, taken, n_kp, minX, minY, maxX, maxY, scale, nlevels);
    Frame& F = cur.F;
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
    F.mbf = mbf; F.mfLogScaleFactor = log_scale_factor;
    F.mTcw = cv::Mat(4, 4, CV_32F, tcw16);
    F.UpdatePoseMatrices();
    for (int k = 0; k < 3; ++k) ow_out[k] = F.mOw.at<float>(k);
    std::vector<MapPoint> mps(n_q);
    std::vector<MapPoint*> vp;
    for (int i = 0; i < n_q; ++i) {
        MapPoint& m = mps[i];
        m.id = i; m.nObs = (q_flags[i] & 2) ? 1 : 0;
        m.mWorldPos = cv::Mat(3, 1, CV_32F, xw + 3 * i); m.mNormalVector = cv::Mat(3, 1, CV_32F, normal + 3 * i);
        m.mfMinDistance = min_distance[i]; m.mfMaxDistance = max_distance[i];
        m.mDescriptor = cv::Mat(1, 32, CV_8U, q_desc + 32 * i);
        m.mbTrackInView = false; m.mTrackProjX = m.mTrackProjY = m.mTrackProjXR = m.mTrackViewCos = 0.f; m.mnTrackScaleLevel = -1;
        if (q_flags[i] & 1) F.isInFrustum(&m, view_cos_limit);
        in_view[i] = m.mbTrackInView ? 1 : 0;
        track4[4 * i] = m.mTrackProjX; track4[4 * i + 1] = m.mTrackProjY; track4[4 * i + 2] = m.mTrackProjXR; track4[4 * i + 3] = m.mTrackViewCos;
        level[i] = m.mnTrackScaleLevel;
        vp.push_back(&m);
    }
    ORBmatcher matcher(nn_ratio, true);
    const int n = matcher.SearchByProjection(F, vp, th);
    cur.result(kp_match);
    return n;
}

// The two vocabulary-bucket searches on the problem dicts of airdos_b200/synth.py::make_bow_problem.  fv*_node / fv*_ptr / fv*_idx:
// the FeatureVector of each side as (ascending node ids, CSR of feature indices) -- nodes present on one side only included, so the
// reference's lower_bound walk is exercised.
namespace {
DBoW2::FeatureVector make_fv(const int32_t* node, const int32_t* ptr, const int32_t* idx, int n_nodes) {
    DBoW2::FeatureVector fv;
    for (int j = 0; j < n_nodes; ++j) fv[(unsigned)node[j]] = std::vector<unsigned int>(idx + ptr[j], idx + ptr[j + 1]);
    return fv;
}
void fill_kf(ORB_SLAM2::KeyFrame& kf, const RefKp* kps, const uint8_t* desc, int n) {
    kf.N = n; kf.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) { kf.mvKeysUn[i].pt.x = kps[i].x; kf.mvKeysUn[i].pt.y = kps[i].y; kf.mvKeysUn[i].angle = kps[i].angle; kf.mvKeysUn[i].octave = kps[i].octave; }
    kf.mDescriptors = cv::Mat(n, 32, CV_8U, desc);
}
}  // namespace

// ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, th)   src/ORBmatcher.cc:825-975, with KeyFrame::GetFeaturesInArea /
// IsInImage (src/KeyFrame.cc:589-633) and MapPoint::PredictScale(dist, KeyFrame*) (src/MapPoint.cc:388-403).  The key-frame holds no map
// point on entry; the surgery calls at the end are recorded, not performed: fused_with[i] = key-point that map point i was fused with, or -1.
int ref_fuse(const RefKp* kps, const float* u_right, const uint8_t* desc, int n_kp, float minX, float minY, float maxX, float maxY, const float* scale,
             const float* inv_sigma2, int nlevels, const float* tcw16, int n_q, const float* xw, const float* normal, const float* min_distance,
             const float* max_distance, const uint8_t* q_desc, const uint8_t* q_flags, float fx, float fy, float cx, float cy, float mbf,
             float log_scale_factor, float th, int32_t* fused_with, float* ow_out) {
    using namespace ORB_SLAM2;
    CurFrame cur(kps, u_right, desc, nullptr, n_kp, minX, minY, maxX, maxY, scale, nlevels);      // builds the grid the key-frame copies
    Frame& F = cur.F;
    F.mTcw = cv::Mat(4, 4, CV_32F, tcw16);
    F.UpdatePoseMatrices();
    KeyFrame kf; fill_kf(kf, kps, desc, n_kp);
    kf.mvuRight.assign(u_right, u_right + n_kp);
    kf.fx = fx; kf.fy = fy; kf.cx = cx; kf.cy = cy; kf.mbf = mbf; kf.mfLogScaleFactor = log_scale_factor; kf.mnScaleLevels = nlevels;
    kf.mvScaleFactors.assign(scale, scale + nlevels); kf.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlevels);
    kf.mnGridCols = FRAME_GRID_COLS; kf.mnGridRows = FRAME_GRID_ROWS;                              // src/KeyFrame.cc:33-44
    kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    kf.mnMinX = Frame::mnMinX; kf.mnMinY = Frame::mnMinY; kf.mnMaxX = Frame::mnMaxX; kf.mnMaxY = Frame::mnMaxY;
    kf.mGrid.resize(kf.mnGridCols);                                                               // src/KeyFrame.cc:52-58
    for (int i = 0; i < kf.mnGridCols; i++) {
        kf.mGrid[i].resize(kf.mnGridRows);
        for (int j = 0; j < kf.mnGridRows; j++) kf.mGrid[i][j] = F.mGrid[i][j];
    }
    kf.Rcw = F.mRcw.clone(); kf.tcw = F.mtcw.clone(); kf.Ow = F.mOw.clone();                        // KeyFrame::SetPose, src/KeyFrame.cc:68-84
    for (int k = 0; k < 3; ++k) ow_out[k] = kf.Ow.at<float>(k);
    kf.mvpMapPoints.assign(n_kp, static_cast<MapPoint*>(nullptr));
    std::unique_ptr<MapPoint[]> mps(new MapPoint[n_q]);
    std::vector<MapPoint*> vp(n_q, static_cast<MapPoint*>(nullptr));
    for (int i = 0; i < n_q; ++i) {
        MapPoint& m = mps[i];
        m.id = i; m.nObs = (q_flags[i] & 2) ? 1 : 0;
        m.mWorldPos = cv::Mat(3, 1, CV_32F, xw + 3 * i); m.mNormalVector = cv::Mat(3, 1, CV_32F, normal + 3 * i);
        m.mfMinDistance = min_distance[i]; m.mfMaxDistance = max_distance[i];
        m.mDescriptor = cv::Mat(1, 32, CV_8U, q_desc + 32 * i);
        if (q_flags[i] & 1) vp[i] = &m;
    }
    g_fuse_kf = &kf;
    ORBmatcher matcher(0.6, true);
    const int n = matcher.Fuse(&kf, vp, th);
    for (int i = 0; i < n_q; ++i) fused_with[i] = mps[i].fused_with;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)   src/ORBmatcher.cc:159-288
// side 1 = the key-frame (flags1: key-points that hold a good map point), side 2 = the frame; match21[i2] = key-frame index or -1
int ref_search_by_bow(const RefKp* k1, const uint8_t* d1, const uint8_t* flags1, int n1, const RefKp* k2, const uint8_t* d2, int n2, const int32_t* node1,
                      const int32_t* ptr1, const int32_t* idx1, int nn1, const int32_t* node2, const int32_t* ptr2, const int32_t* idx2, int nn2, float nn_ratio,
                      int check_ori, int32_t* match21) {
    using namespace ORB_SLAM2;
    KeyFrame kf; fill_kf(kf, k1, d1, n1);
    std::vector<MapPoint> mps(n1);
    kf.mvpMapPoints.assign(n1, static_cast<MapPoint*>(nullptr));
    for (int i = 0; i < n1; ++i) { mps[i].id = i; if (flags1[i]) kf.mvpMapPoints[i] = &mps[i]; }
    kf.mFeatVec = make_fv(node1, ptr1, idx1, nn1);
    Frame F;
    F.N = n2; F.mvKeys.resize(n2);
    for (int i = 0; i < n2; ++i) { F.mvKeys[i].angle = k2[i].angle; F.mvKeys[i].octave = k2[i].octave; F.mvKeys[i].pt.x = k2[i].x; F.mvKeys[i].pt.y = k2[i].y; }
    F.mDescriptors = cv::Mat(n2, 32, CV_8U, d2);
    F.mFeatVec = make_fv(node2, ptr2, idx2, nn2);
    std::vector<MapPoint*> matches;
    ORBmatcher matcher(nn_ratio, check_ori != 0);
    const int n = matcher.SearchByBoW(&kf, F, matches);
    for (int i = 0; i < n2; ++i) match21[i] = matches[i] ? matches[i]->id : -1;
    return n;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)   src/ORBmatcher.cc:657-823 (+ CheckDistEpipolarLine :140-157)
// flags: key-points that already hold a map point (skipped); cam2 = fx, fy, cx, cy of key-frame 2; ow1 / r2w / t2w give the epipole
int ref_search_for_triangulation(const RefKp* k1, const float* ur1, const uint8_t* d1, const uint8_t* has_mp1, int n1, const RefKp* k2, const float* ur2,
                                 const uint8_t* d2, const uint8_t* has_mp2, int n2, const int32_t* node1, const int32_t* ptr1, const int32_t* idx1, int nn1,
                                 const int32_t* node2, const int32_t* ptr2, const int32_t* idx2, int nn2, const float* f12_9, const float* ow1_3,
                                 const float* r2w_9, const float* t2w_3, const float* cam2_4, const float* scale2, const float* sigma2_2, int nlevels,
                                 float nn_ratio, int check_ori, int only_stereo, int32_t* match12) {
    using namespace ORB_SLAM2;
    KeyFrame a, b; fill_kf(a, k1, d1, n1); fill_kf(b, k2, d2, n2);
    a.mvuRight.assign(ur1, ur1 + n1); b.mvuRight.assign(ur2, ur2 + n2);
    std::vector<MapPoint> m1(n1), m2(n2);
    a.mvpMapPoints.assign(n1, static_cast<MapPoint*>(nullptr)); b.mvpMapPoints.assign(n2, static_cast<MapPoint*>(nullptr));
    for (int i = 0; i < n1; ++i) if (has_mp1[i]) a.mvpMapPoints[i] = &m1[i];
    for (int i = 0; i < n2; ++i) if (has_mp2[i]) b.mvpMapPoints[i] = &m2[i];
    a.mFeatVec = make_fv(node1, ptr1, idx1, nn1); b.mFeatVec = make_fv(node2, ptr2, idx2, nn2);
    a.Ow = cv::Mat(3, 1, CV_32F, ow1_3);
    b.Rcw = cv::Mat(3, 3, CV_32F, r2w_9); b.tcw = cv::Mat(3, 1, CV_32F, t2w_3);
    b.fx = cam2_4[0]; b.fy = cam2_4[1]; b.cx = cam2_4[2]; b.cy = cam2_4[3];
    b.mvScaleFactors.assign(scale2, scale2 + nlevels); b.mvLevelSigma2.assign(sigma2_2, sigma2_2 + nlevels);
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher matcher(nn_ratio, check_ori != 0);
    const int n = matcher.SearchForTriangulation(&a, &b, cv::Mat(3, 3, CV_32F, f12_9), pairs, only_stereo != 0);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (const auto& pr : pairs) match12[pr.first] = (int32_t)pr.second;
    return n;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:245-310) for n_points map points; the observations of point p are the
// descriptors desc[point_ptr[p] .. point_ptr[p + 1]), one per key-frame.  The reference walks std::map<KeyFrame*, size_t>, i.e. in
// ADDRESS order of the key-frames: they are allocated as one array here, so that order is the order given.
void ref_distinctive(const uint8_t* desc, const int32_t* point_ptr, int n_points, uint8_t* out_desc) {
    using namespace ORB_SLAM2;
    int max_obs = 1;
    for (int p = 0; p < n_points; ++p) max_obs = std::max(max_obs, point_ptr[p + 1] - point_ptr[p]);
    std::unique_ptr<KeyFrame[]> kfs(new KeyFrame[max_obs]);
    for (int p = 0; p < n_points; ++p) {
        const int n = point_ptr[p + 1] - point_ptr[p];
        MapPoint mp;
        mp.mDescriptor = cv::Mat(1, 32, CV_8U, out_desc + 32 * p);          // left as it is when there is no observation
        for (int k = 0; k < n; ++k) {
            kfs[k].mDescriptors = cv::Mat(1, 32, CV_8U, desc + 32 * (size_t)(point_ptr[p] + k));
            mp.mObservations[&kfs[k]] = 0;
        }
        mp.ComputeDistinctiveDescriptors();
        std::memcpy(out_desc + 32 * p, mp.mDescriptor.ptr<uint8_t>(), 32);
    }
}

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORB_SLAM2::ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, a), cv::Mat(1, 32, CV_8U, b));
}

}  // extern "C"
