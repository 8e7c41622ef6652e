// ref_orb.cpp -- TEST INFRASTRUCTURE: the reference's own key-point distribution, compiled from /root/reference.
//   src/ORBextractor.cc  void ExtractorNode::DivideNode(...)                       (:497-546, the whole definition, unmodified)
//                        vector<cv::KeyPoint> ORBextractor::DistributeOctTree(...)  (:541-765, the whole definition, unmodified)
// Same mechanism as oracle/ref_match.cpp: the build step (oracle/Makefile, target _ref/libref_orb.so) copies the text of these two
// definitions out of the reference tree into oracle/_ref/orb_snippets.inc (git-ignored) and this file compiles it between the class
// declarations of include/ORBextractor.h:36-49 (restated here: the header itself needs OpenCV) and oracle/ref_shim/cv_shim.h.
// Note the reference sorts (size, ExtractorNode*) pairs (:686): nodes with equal key counts are ordered by their heap ADDRESS.  The
// oracle and the CUDA kernel use the creation sequence instead (DESIGN.md convention D.1); tests/test_ref_orb.py compares the two
// on seeded candidate sets and reports how often that tie-break is reached at all.
#include <algorithm>
#include <cmath>
#include <list>
#include <vector>

#include <cstdlib>
#include <new>

#include "ref_shim/cv_shim.h"

// ---- allocator switch (this library is linked with -Bsymbolic: its own operator new / delete calls bind to these definitions).
// DistributeOctTree orders nodes of equal size by their ADDRESS, so its result depends on what the allocator hands out.  In
// "monotonic" mode every allocation gets a higher address than all earlier ones and nothing is reused: address order = creation order,
// which is exactly convention D.1 of the oracle / the CUDA kernel.  In the default mode allocations go to malloc like in the reference binary.
namespace {
char* g_arena = nullptr;
size_t g_arena_size = 0, g_arena_used = 0;
bool g_monotonic = false;
bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + g_arena_size; }
void* arena_or_malloc(size_t n) {
    if (g_monotonic) {
        n = (n + 15) & ~(size_t)15;
        if (g_arena_used + n > g_arena_size) throw std::bad_alloc();
        void* p = g_arena + g_arena_used;
        g_arena_used += n;
        return p;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
}  // namespace
void* operator new(size_t n) { return arena_or_malloc(n); }
void* operator new[](size_t n) { return arena_or_malloc(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void* p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete(void* p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void* p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }

#define CV_PI 3.1415926535897932384626433832795

namespace ORB_SLAM2 {
using namespace std;
using namespace cv;      // src/ORBextractor.cc:64-65

class ExtractorNode {    // include/ORBextractor.h:36-49
public:
    ExtractorNode() : bNoMore(false) {}
    void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4);
    std::vector<cv::KeyPoint> vKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};

class ORBextractor {     // include/ORBextractor.h:51-112: the constructor, DistributeOctTree and the members they touch
public:
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    std::vector<cv::KeyPoint> DistributeOctTree(const std::vector<cv::KeyPoint>& vToDistributeKeys, const int& minX, const int& maxX, const int& minY,
                                                const int& maxY, const int& nFeatures, const int& level);
    std::vector<cv::Mat> mvImagePyramid;
    std::vector<cv::Mat> mvMaskPyramid;     // AirDOS addition (include/ORBextractor.h)
    std::vector<cv::Point> pattern;
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
};

const int PATCH_SIZE = 31;         // src/ORBextractor.cc:72-74
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;
const float factorPI = (float)(CV_PI/180.f);   // src/ORBextractor.cc:108

// static int bit_pattern_31_[256*4] = { ... }   src/ORBextractor.cc:151-409, taken from the reference tree like the functions
#include "_ref/orb_pattern.inc"
;
#include "_ref/orb_snippets.inc"

}  // namespace ORB_SLAM2

extern "C" {

// cand: float [m][3] = (x, y, response) in the reference's vToDistributeKeys order; out: the returned key-points in the list order;
// monotonic != 0: run with the increasing-address allocator (see above)
int ref_distribute(const float* cand, int m, int minX, int maxX, int minY, int maxY, int N, float* out, int cap, int monotonic) {
    if (monotonic && !g_arena) { g_arena_size = (size_t)1 << 30; g_arena = (char*)std::malloc(g_arena_size); }
    g_arena_used = 0;
    g_monotonic = monotonic != 0 && g_arena != nullptr;
    int total = 0;
    {
    std::vector<cv::KeyPoint> v(m);
    for (int i = 0; i < m; ++i) { v[i].pt.x = cand[3 * i]; v[i].pt.y = cand[3 * i + 1]; v[i].response = cand[3 * i + 2]; }
    ORB_SLAM2::ORBextractor ex(2000, 1.2f, 8, 20, 7);
    const std::vector<cv::KeyPoint> r = ex.DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    const int n = std::min((int)r.size(), cap);
    for (int i = 0; i < n; ++i) { out[3 * i] = r[i].pt.x; out[3 * i + 1] = r[i].pt.y; out[3 * i + 2] = r[i].response; }
    total = (int)r.size();
    }
    g_monotonic = false;
    return total;
}

// ORBextractor::ORBextractor (src/ORBextractor.cc:411-472): scale factors, sigma^2, per-level quotas, umax and the pattern as the reference
// builds them.  out arrays: [nlevels] each; umax16: [16]; pattern1024: 512 (x, y) pairs
void ref_extractor_tables(int nfeatures, float scale_factor, int nlevels, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota,
                          int* umax16, int* pattern1024) {
    ORB_SLAM2::ORBextractor ex(nfeatures, scale_factor, nlevels, 20, 7);
    for (int l = 0; l < nlevels; ++l) {
        scale[l] = ex.mvScaleFactor[l]; inv_scale[l] = ex.mvInvScaleFactor[l]; sigma2[l] = ex.mvLevelSigma2[l]; inv_sigma2[l] = ex.mvInvLevelSigma2[l];
        quota[l] = ex.mnFeaturesPerLevel[l];
    }
    for (int v = 0; v < 16; ++v) umax16[v] = ex.umax[v];
    for (int i = 0; i < 512; ++i) { pattern1024[2 * i] = ex.pattern[i].x; pattern1024[2 * i + 1] = ex.pattern[i].y; }
}

// IC_Angle (src/ORBextractor.cc:78-105) and computeOrbDescriptor (:109-148) for n key-points at integer level coordinates (x, y) of one
// level image `img` (orientation) and its blurred copy `blurred` (descriptor).  angle_out: degrees; desc_out: [n][32]
void ref_orient_describe(const uint8_t* img, const uint8_t* blurred, int w, int h, int n, const float* xy, float* angle_out, uint8_t* desc_out) {
    using namespace ORB_SLAM2;
    ORBextractor ex(2000, 1.2f, 8, 20, 7);
    const cv::Mat I(h, w, CV_8U, img), B(h, w, CV_8U, blurred);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint kp; kp.pt.x = xy[2 * i]; kp.pt.y = xy[2 * i + 1];
        kp.angle = IC_Angle(I, kp.pt, ex.umax);
        angle_out[i] = kp.angle;
        computeOrbDescriptor(kp, B, &ex.pattern[0], desc_out + 32 * i);
    }
}

}  // extern "C"
