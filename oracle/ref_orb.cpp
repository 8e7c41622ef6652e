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

namespace cv {
struct Point2i { int x = 0, y = 0; Point2i() {} Point2i(int x_, int y_) : x(x_), y(y_) {} };
}

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

class ORBextractor {     // include/ORBextractor.h:51-112: the one member function under test
public:
    std::vector<cv::KeyPoint> DistributeOctTree(const std::vector<cv::KeyPoint>& vToDistributeKeys, const int& minX, const int& maxX, const int& minY,
                                                const int& maxY, const int& nFeatures, const int& level);
    int nfeatures = 2000;   // include/ORBextractor.h:95 (DistributeOctTree only reserves with it)
};

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
    ORB_SLAM2::ORBextractor ex;
    const std::vector<cv::KeyPoint> r = ex.DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    const int n = std::min((int)r.size(), cap);
    for (int i = 0; i < n; ++i) { out[3 * i] = r[i].pt.x; out[3 * i + 1] = r[i].pt.y; out[3 * i + 2] = r[i].response; }
    total = (int)r.size();
    }
    g_monotonic = false;
    return total;
}

}  // extern "C"
