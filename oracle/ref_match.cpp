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
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

#include "ref_shim/cv_shim.h"

namespace ORB_SLAM2 {
using namespace std;   // src/Frame.cc and src/ORBmatcher.cc both open with it

class ORBmatcher {                       // include/ORBmatcher.h:38-93
public:
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    static const int TH_LOW;
    static const int TH_HIGH;
};
const int ORBmatcher::TH_HIGH = 100;     // src/ORBmatcher.cc:37
const int ORBmatcher::TH_LOW = 50;       // src/ORBmatcher.cc:38

class ORBextractor {                     // include/ORBextractor.h:86
public:
    std::vector<cv::Mat> mvImagePyramid;
};

class Frame {                            // include/Frame.h: the members ComputeStereoMatches touches
public:
    void ComputeStereoMatches();
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    float mbf, mb;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    vector<float> mvScaleFactors, mvInvScaleFactors;
};

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

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORB_SLAM2::ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, a), cv::Mat(1, 32, CV_8U, b));
}

}  // extern "C"
