// ref_orb.cpp -- TEST INFRASTRUCTURE: the reference's own key-point distribution, compiled from /root/reference.
//   src/ORBextractor.cc  void ExtractorNode::DivideNode(...)                       (:497-546, the whole definition, unmodified)
//                        vector<cv::KeyPoint> ORBextractor::DistributeOctTree(...)  (:541-765, the whole definition, unmodified)
// Same mechanism as oracle/ref_match.cpp: the build step (oracle/Makefile, target _ref/libref_orb.so) copies the text of these two
// definitions out of the reference tree into oracle/_ref/orb_snippets.inc (git-ignored) and this file compiles it between the class
// declarations of include/ORBextractor.h:36-49 (restated here: the header itself needs OpenCV) and oracle/ref_shim/cv_shim.h.
// Note the reference sorts (size, ExtractorNode*) pairs (:686): nodes with equal key counts are ordered by their heap ADDRESS.  The
// oracle and the CUDA kernel use the creation sequence instead (DESIGN.md convention D.1); tests/test_ref_orb.py compares the two
// on seeded candidate sets and reports how often that tie-break is reached at all.
// Round 2 additions: the constructor, IC_Angle / computeOrbDescriptor, and the extractor's TOP-LEVEL functions
//   ORBextractor::operator() (:1054-1119), ComputePyramid (:1121-1156), ComputeKeyPointsOctTree (:767-864), computeOrientation,
//   computeDescriptors
// -- whole definitions, unmodified.  They call OpenCV (cv::resize, copyMakeBorder, erode, GaussianBlur, FastFeatureDetector), which is
// not in this image: in the stand-in below those five calls land, through function pointers handed in by the test, in the oracle's
// primitives (oracle/orb_oracle.cpp, themselves held to cv2 4.13 primitive by primitive: oracle/crosscheck_cv2.py).  So a run of
// ref_orb_extract is the reference's own control flow -- pyramid construction with its in-place ROI / border handling, the mask pyramid,
// the cell grid with the ini / min threshold rule, the quad-tree, orientation, the per-level blur + descriptors, the final scaling and
// ordering -- over the oracle's pixel arithmetic; tests/test_ref_orb.py compares it with the oracle's own extract().
#include <algorithm>
#include <cassert>
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
#define CV_8UC1 CV_8U

extern "C" {
struct ref_orb_prims {      // the oracle's pixel primitives (orb_oracle_* of oracle/orb_oracle.cpp)
    void (*border101)(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b);
    void (*erode10)(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
    void (*resize)(const uint8_t* src, int sw, int sh, int sp, uint8_t* dst, int dw, int dh, int dp);
    void (*blur7)(const uint8_t* src, int w, int h, int sp, uint8_t* dst, int dp);
    int (*fast)(const uint8_t* img, int w, int h, int pitch, int threshold, const uint8_t* mask, int mpitch, int32_t* out, int cap);
};
}
static const ref_orb_prims* g_prims = nullptr;

// ---- the OpenCV calls of src/ORBextractor.cc on cv_shim's Mat (8-bit, ROI views share storage like cv::Mat)
namespace cv {
enum { INTER_LINEAR = 1, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
typedef const Mat& InputArray;
typedef Mat& OutputArray;
template <class T> using Ptr = std::shared_ptr<T>;
inline std::vector<uint8_t> tight_copy(const Mat& m) {
    std::vector<uint8_t> t((size_t)m.rows * m.cols);
    for (int y = 0; y < m.rows; ++y) std::memcpy(&t[(size_t)y * m.cols], m.ptr<uint8_t>(y), (size_t)m.cols);
    return t;
}
inline void ensure(Mat& m, int rows, int cols) { if (m.rows != rows || m.cols != cols || m.type() != CV_8U) m.create(rows, cols, CV_8U); }   // Mat::create
inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int) {
    const std::vector<uint8_t> s = tight_copy(src);
    ensure(dst, sz.height, sz.width);                                            // an ROI of the right size is written in place (:1140-1141)
    g_prims->resize(s.data(), src.cols, src.rows, src.cols, dst.ptr_mut(0), sz.width, sz.height, (int)dst.step);
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*BORDER_REFLECT_101 [+ ISOLATED]*/) {
    const std::vector<uint8_t> s = tight_copy(src);                              // src may be the ROI of dst (:1143-1146)
    ensure(dst, src.rows + top + bottom, src.cols + left + right);
    g_prims->border101(s.data(), src.cols, src.rows, src.cols, dst.ptr_mut(0), (int)dst.step, top);
}
inline void erode(const Mat& src, Mat& dst, const Mat& /*ones(10, 10)*/) {
    const std::vector<uint8_t> s = tight_copy(src);
    ensure(dst, src.rows, src.cols);
    g_prims->erode10(s.data(), src.cols, src.rows, src.cols, dst.ptr_mut(0), (int)dst.step);
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size /*7 x 7*/, double /*2*/, double /*2*/, int /*BORDER_REFLECT_101*/) {
    const std::vector<uint8_t> s = tight_copy(src);
    ensure(dst, src.rows, src.cols);
    g_prims->blur7(s.data(), src.cols, src.rows, src.cols, dst.ptr_mut(0), (int)dst.step);
}
class FastFeatureDetector {     // FastFeatureDetector::create(threshold, nonmaxSuppression)->detect(image, keypoints, mask)
    int th_;
public:
    explicit FastFeatureDetector(int th) : th_(th) {}
    static Ptr<FastFeatureDetector> create(int threshold, bool /*nonmaxSuppression = true*/) { return std::make_shared<FastFeatureDetector>(threshold); }
    void detect(const Mat& image, std::vector<KeyPoint>& keypoints, const Mat& mask) {
        keypoints.clear();
        const int cap = image.rows * image.cols;
        std::vector<int32_t> out((size_t)3 * std::max(cap, 1));
        const int n = g_prims->fast(image.ptr<uint8_t>(0), image.cols, image.rows, (int)image.step, th_, mask.empty() ? nullptr : mask.ptr<uint8_t>(0), (int)mask.step,
                                    out.data(), cap);
        for (int i = 0; i < n; ++i) {      // cv::KeyPoint(x, y, 7.f, -1, score)
            KeyPoint k; k.pt.x = (float)out[3 * i]; k.pt.y = (float)out[3 * i + 1]; k.size = 7.f; k.angle = -1.f; k.response = (float)out[3 * i + 2];
            keypoints.push_back(k);
        }
    }
};
}  // namespace cv

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
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);
    void ComputePyramid(cv::Mat image, cv::Mat Mask);
    void ComputeKeyPointsOctTree(std::vector<std::vector<cv::KeyPoint> >& allKeypoints);
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

// ORBextractor::operator() (src/ORBextractor.cc:1054-1119) on one image: kps = 24-byte records (x, y, size, angle, response, octave), desc = n x 32.
// mask == NULL: an all-255 mask (the reference always passes one).  pyr_out (optional): the nlevels ROIs packed back to back.
// Returns the number of key-points (or -1 if cap is too small).
int ref_orb_extract(const ref_orb_prims* prims, const uint8_t* img, int w, int h, const uint8_t* mask, int nfeatures, float scale_factor, int nlevels, int ini_th,
                    int min_th, void* kps, uint8_t* desc, int cap, uint8_t* pyr_out, int monotonic) {
    if (monotonic && !g_arena) { g_arena_size = (size_t)1 << 30; g_arena = (char*)std::malloc(g_arena_size); }
    g_arena_used = 0;
    g_monotonic = monotonic != 0 && g_arena != nullptr;
    g_prims = prims;
    int n = 0;
    {
        ORB_SLAM2::ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
        cv::Mat I(h, w, CV_8U, img), M;
        if (mask) M = cv::Mat(h, w, CV_8U, mask);
        else { M = cv::Mat(h, w, CV_8U); std::memset(M.ptr_mut(0), 255, (size_t)w * h); }
        std::vector<cv::KeyPoint> k;
        cv::Mat D;
        ex(I, M, k, D);
        n = (int)k.size();
        if (n <= cap) {
            struct Rec { float x, y, size, angle, response; int32_t octave; };
            Rec* r = (Rec*)kps;
            for (int i = 0; i < n; ++i) {
                r[i] = Rec{k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave};
                std::memcpy(desc + (size_t)32 * i, D.ptr<uint8_t>(i), 32);
            }
        } else n = -1;
        if (pyr_out) {
            size_t o = 0;
            for (int l = 0; l < nlevels; ++l)
                for (int y = 0; y < ex.mvImagePyramid[l].rows; ++y) { std::memcpy(pyr_out + o, ex.mvImagePyramid[l].ptr<uint8_t>(y), ex.mvImagePyramid[l].cols); o += ex.mvImagePyramid[l].cols; }
        }
    }
    g_monotonic = false;
    g_prims = nullptr;
    return n;
}

}  // extern "C"
