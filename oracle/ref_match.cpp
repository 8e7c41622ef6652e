// ref_match.cpp -- TEST INFRASTRUCTURE: the reference's own stereo matcher, compiled from /root/reference.
//   src/Frame.cc       void Frame::ComputeStereoMatches()           (the whole function body, unmodified)
//   src/ORBmatcher.cc  int ORBmatcher::DescriptorDistance(...)      (the whole function body, unmodified)
// The two files cannot be compiled as they are (OpenCV, Eigen, DBoW2, Pangolin headers are absent and the rest of each file needs
// them), so the build step (oracle/Makefile, target _ref/libref_match.so) copies the text of exactly these two function
// definitions out of the reference tree into oracle/_ref/match_snippets.inc (git-ignored, never committed:
// oracle/extract_ref_fn.py) and this file compiles that text between stand-in declarations of the classes it is a member of:
// the members the statements read and write, with the reference's names and types, and oracle/ref_shim/cv_shim.h for the
// handful of cv::Mat / cv::KeyPoint operations.  oracle/gen_ref_match_golden.py runs it on seeded stereo pairs and writes
// tests/golden/stereo_ref.npz, which pins oracle/match_oracle.cpp (and through it the CUDA matcher) to the literal reference.
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

#include "ref_shim/cv_shim.h"

namespace ORB_SLAM2 {
using namespace std;   // src/Frame.cc and src/ORBmatcher.cc both open with it

class Frame;
class MapPoint {                         // include/MapPoint.h: what the matcher reads of a map point
public:
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }         // src/MapPoint.cc:79-83
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }     // src/MapPoint.cc:312-316
    int Observations() { return nObs; }                         // src/MapPoint.cc:134-138
    bool isBad() { return mbBad; }
    // variables used by the tracking (include/MapPoint.h:87-94)
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    // stand-in state
    cv::Mat mWorldPos, mDescriptor;
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
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
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

#define FRAME_GRID_ROWS 48               // include/Frame.h:41-42
#define FRAME_GRID_COLS 64

class Frame {                            // include/Frame.h: the members the extracted functions touch, with the reference's names
public:
    void ComputeStereoMatches();
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
    cv::Mat mTcw;
    int mnScaleLevels;
    vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
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
    std::vector<ORB_SLAM2::MapPoint> held;
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
        held.resize(n_kp);
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

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORB_SLAM2::ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, a), cv::Mat(1, 32, CV_8U, b));
}

}  // extern "C"
