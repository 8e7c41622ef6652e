// oracle/orb_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the AirDOS / ORB-SLAM2 feature front-end, used only as the parity
// checker (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
// The product path (airdos_b200/csrc/*.cu behind include/airdos_b200.h) never calls it.
//
// It follows the reference's *algorithm structure* (per-cell FAST on sub-images, list based
// quad-tree, whole-level blur) so that it is an independent statement of what the CUDA
// kernels -- which are organised very differently -- must reproduce bit for bit.
//
// Parity pinning: the reference ships no tests or golden vectors for this path
// (SURVEY.md section 4), and cannot be compiled here (needs OpenCV C++ headers + Eigen).  The
// un-vendored third-party arithmetic (OpenCV FAST / resize / copyMakeBorder / erode /
// GaussianBlur / fastAtan2) is pinned against cv2 4.13.0 by oracle/crosscheck_cv2.py, which
// also emits tests/golden/*.npz.  Conventions for the reference's ill-defined corners are in
// DESIGN.md (quad-tree tie-break by creation sequence, no FMA contraction, cos/sin in
// double rounded once to float).
//
// Reference lines each function follows are cited at the function.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <thread>
#include <vector>

#include "../include/airdos_orb_pattern.h"

namespace {

constexpr int kPatch = 31;       // PATCH_SIZE        src/ORBextractor.cc:73
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE   src/ORBextractor.cc:74
constexpr int kEdge = 19;        // EDGE_THRESHOLD    src/ORBextractor.cc:75

const signed char kPatX[512] = AIRDOS_ORB_PATTERN_X;
const signed char kPatY[512] = AIRDOS_ORB_PATTERN_Y;

struct KeyPoint {  // mirrors the cv::KeyPoint fields the SLAM code reads; 24 bytes
    float x, y, size, angle, response;
    int32_t octave;
};

// cvRound: round half to even (SSE cvtss2si / rint).
inline int round_even(float v) { return (int)std::nearbyintf(v); }
inline int round_even(double v) { return (int)std::nearbyint(v); }

// ---------------------------------------------------------------------------------------
// BORDER_REFLECT_101 index map: gfedcb|abcdefgh|gfedcba
inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

// copyMakeBorder(src, dst, 19,19,19,19, BORDER_REFLECT_101 [+ISOLATED]);  src/ORBextractor.cc:1142-1152
// dst is the full (w+2b) x (h+2b) buffer; roi content is written too.
void make_border101(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b) {
    for (int y = -b; y < h + b; ++y) {
        const uint8_t* s = src + (size_t)reflect101(y, h) * spitch;
        uint8_t* d = dst + (size_t)(y + b) * dpitch;
        for (int x = -b; x < w + b; ++x) d[x + b] = s[reflect101(x, w)];
    }
}

// cv::erode(mask, out, ones(10,10)) with the default anchor (5,5) and the default border
// (morphologyDefaultBorderValue -> +inf for erode);  src/ORBextractor.cc:1130-1131
void erode10(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch) {
    std::vector<uint8_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int m = 255;
            for (int dx = -5; dx <= 4; ++dx) {
                int xx = x + dx;
                if (xx >= 0 && xx < w) m = std::min<int>(m, src[(size_t)y * spitch + xx]);
            }
            tmp[(size_t)y * w + x] = (uint8_t)m;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int m = 255;
            for (int dy = -5; dy <= 4; ++dy) {
                int yy = y + dy;
                if (yy >= 0 && yy < h) m = std::min<int>(m, tmp[(size_t)yy * w + x]);
            }
            dst[(size_t)y * dpitch + x] = (uint8_t)m;
        }
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for 8UC1: 11-bit fixed point coefficients,
// int32 horizontal pass, the (>>4, *b >>16, +2 >>2) vertical pass.  src/ORBextractor.cc:1139-1140
struct LinCoef { int s; int a0, a1; };
void linear_coeffs(int src, int dst, std::vector<LinCoef>& c) {
    c.resize(dst);
    const double scale = (double)src / dst;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        c[d].s = s;
        c[d].a0 = round_even((1.f - f) * 2048.f);
        c[d].a1 = round_even(f * 2048.f);
    }
}
void resize_linear_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch) {
    std::vector<LinCoef> cx, cy;
    linear_coeffs(sw, dw, cx);
    linear_coeffs(sh, dh, cy);
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; ++y) {
        const int sy0 = cy[y].s, sy1 = std::min(sy0 + 1, sh - 1);
        const uint8_t* s0 = src + (size_t)sy0 * spitch;
        const uint8_t* s1 = src + (size_t)sy1 * spitch;
        for (int x = 0; x < dw; ++x) {
            const int sx0 = cx[x].s, sx1 = std::min(sx0 + 1, sw - 1);
            r0[x] = s0[sx0] * cx[x].a0 + s0[sx1] * cx[x].a1;
            r1[x] = s1[sx0] * cx[x].a0 + s1[sx1] * cx[x].a1;
        }
        const int b0 = cy[y].a0, b1 = cy[y].a1;
        uint8_t* d = dst + (size_t)y * dpitch;
        for (int x = 0; x < dw; ++x)
            d[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) for 8UC1: 8.8 fixed-point
// separable kernel {18,34,48,56,48,34,18}/256, result (v + 2^15) >> 16.  src/ORBextractor.cc:1100
const int kBlurK[7] = {18, 34, 48, 56, 48, 34, 18};
void blur7_u8(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch) {
    // row pass into a buffer with 3 reflected rows above and below, so the column pass needs no index mapping
    std::vector<uint16_t> rows((size_t)w * (h + 6));
    std::vector<uint8_t> line((size_t)w + 6);
    for (int y = -3; y < h + 3; ++y) {
        const uint8_t* s = src + (size_t)reflect101(y, h) * spitch;
        for (int x = -3; x < w + 3; ++x) line[x + 3] = s[reflect101(x, w)];
        uint16_t* r = &rows[(size_t)(y + 3) * w];
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = &line[x];
            r[x] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* r = &rows[(size_t)y * w];
        uint8_t* d = dst + (size_t)y * dpitch;
        for (int x = 0; x < w; ++x) {
            const uint32_t acc = 18u * (r[x] + r[x + 6 * (size_t)w]) + 34u * (r[x + w] + r[x + 5 * (size_t)w]) +
                                 48u * (r[x + 2 * (size_t)w] + r[x + 4 * (size_t)w]) + 56u * r[x + 3 * (size_t)w];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// cv::fastAtan2 (degrees): degree-7 odd polynomial, every operation rounded to float32.
// src/ORBextractor.cc:104
float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)DBL_EPSILON;
    volatile float ax = std::fabs(x), ay = std::fabs(y);  // volatile: forbid contraction/reassociation
    volatile float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        volatile float t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
        a = t * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        volatile float t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
        t = t * c;
        a = 90.f - t;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------
// FAST-9/16 exactly as cv::FastFeatureDetector(threshold, nonmaxSuppression=true, TYPE_9_16)
// behaves on a sub-image (src/ORBextractor.cc:812-824): tests pixels 3 <= x < w-3, 3 <= y < h-3,
// response = (largest t for which the pixel is still a corner) = best - 1, 3x3 strict NMS in
// which non-corners and the untested frame count as 0, row-major output, then the mask post-filter.
const int kCircleDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kCircleDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// best = max over the 16 arcs of 9 contiguous circle pixels of max(min(d), min(-d)), d = centre - circle.
inline int fast_best(const uint8_t* p, const int* off) {
    int d[16];
    const int v = p[0];
    for (int i = 0; i < 16; ++i) d[i] = v - p[off[i]];
    // window minima / maxima of length 9 by doubling (2, 4, 8, +1): same value as scanning all 16 arcs
    int lo2[16], hi2[16], lo4[16], hi4[16];
    for (int i = 0; i < 16; ++i) { lo2[i] = std::min(d[i], d[(i + 1) & 15]); hi2[i] = std::max(d[i], d[(i + 1) & 15]); }
    for (int i = 0; i < 16; ++i) { lo4[i] = std::min(lo2[i], lo2[(i + 2) & 15]); hi4[i] = std::max(hi2[i], hi2[(i + 2) & 15]); }
    int best = -256;
    for (int i = 0; i < 16; ++i) {
        const int lo = std::min(std::min(lo4[i], lo4[(i + 4) & 15]), d[(i + 8) & 15]);
        const int hi = std::max(std::max(hi4[i], hi4[(i + 4) & 15]), d[(i + 8) & 15]);
        best = std::max(best, std::max(lo, -hi));
    }
    return best;
}

struct CellKp { int x, y, score; };

// `best` for a run of n adjacent pixels: sixteen difference rows, then the window minima / maxima by doubling as plain
// element-wise loops over int16 arrays, which the compiler turns into AVX2 code (OpenCV's own FAST is SIMD as well, so this
// is also the fairer CPU baseline).  Same values as fast_best() per pixel.
void fast_best_run(const uint8_t* row, const int* off, int n, int16_t* best /*[n]*/, int16_t* scratch /*[80 * n]*/) {
    int16_t* d = scratch;
    int16_t* lo2 = d + 16 * n; int16_t* hi2 = lo2 + 16 * n; int16_t* lo4 = hi2 + 16 * n; int16_t* hi4 = lo4 + 16 * n;
    for (int i = 0; i < 16; ++i) {
        const uint8_t* q = row + off[i];
        int16_t* di = d + (size_t)i * n;
        for (int x = 0; x < n; ++x) di[x] = (int16_t)((int)row[x] - (int)q[x]);
    }
    for (int i = 0; i < 16; ++i) {
        const int16_t* a = d + (size_t)i * n; const int16_t* c = d + (size_t)((i + 1) & 15) * n;
        int16_t* l = lo2 + (size_t)i * n; int16_t* h = hi2 + (size_t)i * n;
        for (int x = 0; x < n; ++x) { l[x] = std::min(a[x], c[x]); h[x] = std::max(a[x], c[x]); }
    }
    for (int i = 0; i < 16; ++i) {
        const int16_t* la = lo2 + (size_t)i * n; const int16_t* lc = lo2 + (size_t)((i + 2) & 15) * n;
        const int16_t* ha = hi2 + (size_t)i * n; const int16_t* hc = hi2 + (size_t)((i + 2) & 15) * n;
        int16_t* l = lo4 + (size_t)i * n; int16_t* h = hi4 + (size_t)i * n;
        for (int x = 0; x < n; ++x) { l[x] = std::min(la[x], lc[x]); h[x] = std::max(ha[x], hc[x]); }
    }
    for (int x = 0; x < n; ++x) best[x] = -256;
    for (int i = 0; i < 16; ++i) {
        const int16_t* la = lo4 + (size_t)i * n; const int16_t* lc = lo4 + (size_t)((i + 4) & 15) * n;
        const int16_t* ha = hi4 + (size_t)i * n; const int16_t* hc = hi4 + (size_t)((i + 4) & 15) * n;
        const int16_t* e = d + (size_t)((i + 8) & 15) * n;
        for (int x = 0; x < n; ++x) {
            const int16_t lo = std::min(std::min(la[x], lc[x]), e[x]);
            const int16_t hi = std::max(std::max(ha[x], hc[x]), e[x]);
            best[x] = std::max(best[x], std::max(lo, (int16_t)-hi));
        }
    }
}

void fast_best_row(const uint8_t* row, const int* off, int n, int16_t* best /*[n]*/, std::vector<int16_t>& scratch) {
    scratch.resize((size_t)80 * n);
    fast_best_run(row, off, n, best, scratch.data());
}

void fast_detect_cell(const uint8_t* img, int w, int h, int pitch, int threshold,
                      const uint8_t* mask, int mpitch, std::vector<CellKp>& out,
                      std::vector<int>& score /*scratch w*h*/) {
    out.clear();
    if (w < 7 || h < 7) return;
    // padded private copy of the sub-image: rows of a whole number of 16-pixel vectors plus slack, so the row loops below
    // have no scalar remainder and never read outside the copy (the extra columns are computed and thrown away)
    const int n = w - 6, n_pad = (n + 15) & ~15, lp = n_pad + 6 + 2;
    static thread_local std::vector<uint8_t> local;
    static thread_local std::vector<int16_t> scratch, best;
    local.assign((size_t)lp * h, 0);
    for (int y = 0; y < h; ++y) std::memcpy(&local[(size_t)y * lp], img + (size_t)y * pitch, (size_t)w);
    int off[16];
    for (int i = 0; i < 16; ++i) off[i] = kCircleDy[i] * lp + kCircleDx[i];
    score.assign((size_t)w * h, 0);
    best.resize(n_pad);
    for (int y = 3; y < h - 3; ++y) {
        fast_best_row(&local[(size_t)y * lp + 3], off, n_pad, best.data(), scratch);
        int* srow = &score[(size_t)y * w + 3];
        for (int x = 0; x < n; ++x)
            if (best[x] > threshold) srow[x] = best[x] - 1;
    }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = score[(size_t)y * w + x];
            if (s == 0) continue;
            const int* c = &score[(size_t)y * w + x];
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                s > c[w - 1] && s > c[w] && s > c[w + 1]) {
                if (mask && mask[(size_t)y * mpitch + x] == 0) continue;
                out.push_back({x, y, s});
            }
        }
}

// ---------------------------------------------------------------------------------------
// Quad-tree distribution.  src/ORBextractor.cc:483-539 (DivideNode), 541-765 (DistributeOctTree).
struct QNode {
    int ulx, uly, brx, bry;  // UL / BR corners; UR = (brx, uly), BL = (ulx, bry)
    std::vector<KeyPoint> keys;
    bool leaf = false;       // bNoMore
    long seq = 0;            // creation sequence number: replaces the reference's pointer tie-break
    std::list<QNode>::iterator self;
};

void divide(const QNode& n, QNode c[4]) {
    const int halfX = (int)std::ceil((float)(n.brx - n.ulx) / 2);
    const int halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
    const int mx = n.ulx + halfX, my = n.uly + halfY;
    c[0].ulx = n.ulx; c[0].uly = n.uly; c[0].brx = mx;    c[0].bry = my;
    c[1].ulx = mx;    c[1].uly = n.uly; c[1].brx = n.brx; c[1].bry = my;
    c[2].ulx = n.ulx; c[2].uly = my;    c[2].brx = mx;    c[2].bry = n.bry;
    c[3].ulx = mx;    c[3].uly = my;    c[3].brx = n.brx; c[3].bry = n.bry;
    for (const KeyPoint& kp : n.keys) {
        const bool left = kp.x < (float)mx, top = kp.y < (float)my;
        c[left ? (top ? 0 : 2) : (top ? 1 : 3)].keys.push_back(kp);
    }
    for (int i = 0; i < 4; ++i) c[i].leaf = (c[i].keys.size() == 1);
}

// set when a level's region is so tall that nIni = round(width / height) = 0: the reference then divides by zero and indexes an
// empty root vector (src/ORBextractor.cc:545-549, undefined behaviour); oracle and library refuse such image shapes instead
static thread_local bool g_bad_aspect = false;

void distribute_quadtree(const std::vector<KeyPoint>& cand, int minX, int maxX, int minY, int maxY,
                         int N, std::vector<KeyPoint>& out) {
    out.clear();
    const int nIni = (int)std::round((float)(maxX - minX) / (float)(maxY - minY));
    if (nIni < 1) { g_bad_aspect = true; return; }
    const float hX = (float)(maxX - minX) / nIni;
    std::list<QNode> nodes;
    std::vector<QNode*> roots(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; ++i) {
        QNode n;
        n.ulx = (int)(hX * (float)i);
        n.brx = (int)(hX * (float)(i + 1));
        n.uly = 0;
        n.bry = maxY - minY;
        n.seq = seq++;
        nodes.push_back(n);
        roots[i] = &nodes.back();
    }
    for (const KeyPoint& kp : cand) roots[(int)(kp.x / hX)]->keys.push_back(kp);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->leaf = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }

    typedef std::pair<int, long> SizeSeq;  // (count, creation seq) -- deterministic stand-in for (count, pointer)
    std::vector<std::pair<SizeSeq, QNode*>> expandable;
    auto push_children = [&](QNode c[4], int* nToExpand) {
        for (int i = 0; i < 4; ++i) {
            if (c[i].keys.empty()) continue;
            c[i].seq = seq++;
            nodes.push_front(c[i]);
            nodes.front().self = nodes.begin();
            if (c[i].keys.size() > 1) {
                if (nToExpand) ++*nToExpand;
                expandable.push_back({{(int)c[i].keys.size(), nodes.front().seq}, &nodes.front()});
            }
        }
    };

    bool finish = false;
    while (!finish) {
        const int prevSize = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->leaf) { ++it; continue; }
            QNode c[4];
            divide(*it, c);
            push_children(c, &nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = (int)nodes.size();
                std::vector<std::pair<SizeSeq, QNode*>> todo = expandable;
                expandable.clear();
                std::sort(todo.begin(), todo.end(),
                          [](const std::pair<SizeSeq, QNode*>& a, const std::pair<SizeSeq, QNode*>& b) { return a.first < b.first; });
                for (int j = (int)todo.size() - 1; j >= 0; --j) {
                    QNode c[4];
                    divide(*todo[j].second, c);
                    push_children(c, nullptr);
                    nodes.erase(todo[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev2) finish = true;
            }
        }
    }
    out.reserve(nodes.size());
    for (const QNode& n : nodes) {
        const KeyPoint* best = &n.keys[0];
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (n.keys[k].response > best->response) best = &n.keys[k];
        out.push_back(*best);
    }
}

// ---------------------------------------------------------------------------------------
struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactor;
    std::vector<float> scale, invScale;
    std::vector<int> quota;
    int umax[kHalfPatch + 1];

    // src/ORBextractor.cc:411-472
    Extractor(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
        scale.resize(nl); invScale.resize(nl); quota.resize(nl);
        scale[0] = 1.0f;
        for (int i = 1; i < nl; ++i) scale[i] = scale[i - 1] * sf;
        for (int i = 0; i < nl; ++i) invScale[i] = 1.0f / scale[i];
        const float factor = 1.0f / sf;
        float per = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            quota[l] = round_even(per);
            sum += quota[l];
            per *= factor;
        }
        quota[nl - 1] = std::max(nf - sum, 0);
        int v, v0;
        const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
        const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v <= vmax; ++v) umax[v] = round_even(std::sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    void level_size(int w, int h, int l, int* lw, int* lh) const {
        *lw = round_even((float)w * invScale[l]);
        *lh = round_even((float)h * invScale[l]);
    }
};

struct Level {
    int w, h, pitch;            // ROI size; pitch of the bordered buffer
    std::vector<uint8_t> buf;   // (w+38) x (h+38)
    std::vector<uint8_t> mbuf;  // same for the mask (empty if no mask)
    uint8_t* roi() { return buf.data() + (size_t)kEdge * pitch + kEdge; }
    const uint8_t* roi() const { return buf.data() + (size_t)kEdge * pitch + kEdge; }
    uint8_t* mroi() { return mbuf.empty() ? nullptr : mbuf.data() + (size_t)kEdge * pitch + kEdge; }
};

// src/ORBextractor.cc:1121-1156
void build_pyramid(const Extractor& ex, const uint8_t* img, int w, int h, int pitch,
                   const uint8_t* mask, int mpitch, std::vector<Level>& pyr) {
    pyr.resize(ex.nlevels);
    for (int l = 0; l < ex.nlevels; ++l) {
        Level& L = pyr[l];
        ex.level_size(w, h, l, &L.w, &L.h);
        L.pitch = L.w + 2 * kEdge;
        L.buf.assign((size_t)L.pitch * (L.h + 2 * kEdge), 0);
        if (mask) L.mbuf.assign(L.buf.size(), 0); else L.mbuf.clear();
        if (l == 0) {
            make_border101(img, w, h, pitch, L.buf.data(), L.pitch, kEdge);
            if (mask) {
                std::vector<uint8_t> er((size_t)w * h);
                erode10(mask, w, h, mpitch, er.data(), w);
                make_border101(er.data(), w, h, w, L.mbuf.data(), L.pitch, kEdge);
            }
        } else {
            Level& P = pyr[l - 1];
            std::vector<uint8_t> tmp((size_t)L.w * L.h);
            resize_linear_u8(P.roi(), P.w, P.h, P.pitch, tmp.data(), L.w, L.h, L.w);
            make_border101(tmp.data(), L.w, L.h, L.w, L.buf.data(), L.pitch, kEdge);
            if (mask) {
                resize_linear_u8(P.mroi(), P.w, P.h, P.pitch, tmp.data(), L.w, L.h, L.w);
                make_border101(tmp.data(), L.w, L.h, L.w, L.mbuf.data(), L.pitch, kEdge);
            }
        }
    }
}

// src/ORBextractor.cc:78-105
float ic_angle(const uint8_t* roi, int pitch, float px, float py, const int* umax) {
    int m01 = 0, m10 = 0;
    const uint8_t* c = roi + (ptrdiff_t)round_even(py) * pitch + round_even(px);
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int p = c[u + v * pitch], m = c[u - v * pitch];
            vsum += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vsum;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// src/ORBextractor.cc:109-148.  Conventions: cos/sin evaluated in double on the float angle and
// rounded once to float (D.3); products and sums individually rounded, never contracted (D.2).
void rbrief(const uint8_t* blurred, int pitch, float px, float py, float angle_deg, uint8_t desc[32]) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = angle_deg * factorPI;
    const float a = (float)std::cos((double)ang), b = (float)std::sin((double)ang);
    const uint8_t* c = blurred + (ptrdiff_t)round_even(py) * pitch + round_even(px);
    auto sample = [&](int idx) -> int {
        const float x = (float)kPatX[idx], y = (float)kPatY[idx];
        volatile float xb = x * b, ya = y * a, xa = x * a, yb = y * b;
        volatile float r = xb + ya, q = xa - yb;
        return c[(ptrdiff_t)round_even((float)r) * pitch + round_even((float)q)];
    };
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int t0 = sample(16 * i + 2 * k), t1 = sample(16 * i + 2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// src/ORBextractor.cc:767-864 for one level: per-cell FAST with ini/min thresholds + distribution.
void detect_level(const Extractor& ex, Level& L, int level, std::vector<KeyPoint>& cand, std::vector<KeyPoint>& kept) {
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    cand.clear();
    kept.clear();
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<CellKp> cell;
    std::vector<int> scratch;
    const uint8_t* roi = L.roi();
    const uint8_t* mroi = L.mroi();
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            const uint8_t* sub = roi + (ptrdiff_t)y0 * L.pitch + x0;
            const uint8_t* msub = mroi ? mroi + (ptrdiff_t)y0 * L.pitch + x0 : nullptr;
            fast_detect_cell(sub, cw, ch, L.pitch, ex.iniTh, msub, L.pitch, cell, scratch);
            if (cell.empty()) fast_detect_cell(sub, cw, ch, L.pitch, ex.minTh, msub, L.pitch, cell, scratch);
            for (const CellKp& k : cell) {
                KeyPoint kp;
                kp.x = (float)k.x + (float)(j * wCell);
                kp.y = (float)k.y + (float)(i * hCell);
                kp.size = 7.f; kp.angle = -1.f; kp.response = (float)k.score; kp.octave = 0;
                cand.push_back(kp);
            }
        }
    }
    distribute_quadtree(cand, minBX, maxBX, minBY, maxBY, ex.quota[level], kept);
    const int scaledPatch = (int)(kPatch * ex.scale[level]);
    for (KeyPoint& kp : kept) {
        kp.x += minBX; kp.y += minBY;
        kp.octave = level;
        kp.size = (float)scaledPatch;
    }
    for (KeyPoint& kp : kept) kp.angle = ic_angle(roi, L.pitch, kp.x, kp.y, ex.umax);
}

// src/ORBextractor.cc:1054-1119
int extract(const Extractor& ex, const uint8_t* img, int w, int h, int pitch, const uint8_t* mask, int mpitch,
            KeyPoint* kps, uint8_t* desc, int cap, std::vector<Level>* keep_pyr, int32_t* cand_counts) {
    std::vector<Level> local;
    std::vector<Level>& pyr = keep_pyr ? *keep_pyr : local;
    build_pyramid(ex, img, w, h, pitch, mask, mpitch, pyr);
    int n = 0;
    std::vector<KeyPoint> cand, kept;
    std::vector<uint8_t> blurred;
    for (int l = 0; l < ex.nlevels; ++l) {
        Level& L = pyr[l];
        detect_level(ex, L, l, cand, kept);
        if (cand_counts) cand_counts[l] = (int32_t)cand.size();
        if (kept.empty()) continue;
        blurred.resize((size_t)L.w * L.h);
        blur7_u8(L.roi(), L.w, L.h, L.pitch, blurred.data(), L.w);  // blur of the ROI clone: reflect at the ROI edge
        for (KeyPoint& kp : kept) {
            if (n >= cap) return -1;
            rbrief(blurred.data(), L.w, kp.x, kp.y, kp.angle, desc + (size_t)n * 32);
            KeyPoint o = kp;
            if (l != 0) { o.x = kp.x * ex.scale[l]; o.y = kp.y * ex.scale[l]; }
            kps[n++] = o;
        }
    }
    return n;
}

}  // namespace

// =======================================================================================
// C interface for the Python tests (ctypes).
extern "C" {

void orb_oracle_border101(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b) {
    make_border101(src, w, h, spitch, dst, dpitch, b);
}
void orb_oracle_erode10(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch) { erode10(src, w, h, spitch, dst, dpitch); }
void orb_oracle_resize(const uint8_t* src, int sw, int sh, int sp, uint8_t* dst, int dw, int dh, int dp) {
    resize_linear_u8(src, sw, sh, sp, dst, dw, dh, dp);
}
void orb_oracle_blur7(const uint8_t* src, int w, int h, int sp, uint8_t* dst, int dp) { blur7_u8(src, w, h, sp, dst, dp); }
float orb_oracle_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }

// FAST on one sub-image; out = int32 triples (x, y, score); returns count (<= cap).
int orb_oracle_fast(const uint8_t* img, int w, int h, int pitch, int threshold, const uint8_t* mask, int mpitch,
                    int32_t* out, int cap) {
    std::vector<CellKp> k;
    std::vector<int> scratch;
    fast_detect_cell(img, w, h, pitch, threshold, mask, mpitch, k, scratch);
    int n = 0;
    for (const CellKp& c : k) {
        if (n >= cap) break;
        out[3 * n] = c.x; out[3 * n + 1] = c.y; out[3 * n + 2] = c.score; ++n;
    }
    return (int)k.size();
}

// Quad-tree distribution on its own: cand = n records of (x, y, response) floats.
int orb_oracle_distribute(const float* cand, int n, int minX, int maxX, int minY, int maxY, int N, float* out, int cap) {
    std::vector<KeyPoint> c(n), o;
    for (int i = 0; i < n; ++i) { c[i] = KeyPoint{cand[3 * i], cand[3 * i + 1], 7.f, -1.f, cand[3 * i + 2], 0}; }
    distribute_quadtree(c, minX, maxX, minY, maxY, N, o);
    for (size_t i = 0; i < o.size() && (int)i < cap; ++i) { out[3 * i] = o[i].x; out[3 * i + 1] = o[i].y; out[3 * i + 2] = o[i].response; }
    return (int)o.size();
}

void orb_oracle_params(int nfeatures, float sf, int nlevels, int w, int h, int32_t* lw, int32_t* lh, int32_t* quota,
                       float* scale, int32_t* umax16) {
    Extractor ex(nfeatures, sf, nlevels, 20, 7);
    for (int l = 0; l < nlevels; ++l) {
        int a, b; ex.level_size(w, h, l, &a, &b);
        lw[l] = a; lh[l] = b; quota[l] = ex.quota[l]; scale[l] = ex.scale[l];
    }
    for (int v = 0; v <= kHalfPatch; ++v) umax16[v] = ex.umax[v];
}

// Constructor tables as the reference's getters return them (include/ORBextractor.h:63-85, src/ORBextractor.cc:411-472): scale factors,
// their inverses, sigma^2 = scale^2 and its inverse, per-level quotas, umax, and the 512 (x, y) pattern points (:151-409).
void orb_oracle_tables(int nfeatures, float sf, int nlevels, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* quota,
                       int32_t* umax16, int32_t* pattern1024) {
    Extractor ex(nfeatures, sf, nlevels, 20, 7);
    for (int l = 0; l < nlevels; ++l) {
        scale[l] = ex.scale[l]; inv_scale[l] = ex.invScale[l];
        sigma2[l] = ex.scale[l] * ex.scale[l]; inv_sigma2[l] = 1.0f / sigma2[l];
        quota[l] = ex.quota[l];
    }
    for (int v = 0; v <= kHalfPatch; ++v) umax16[v] = ex.umax[v];
    for (int i = 0; i < 512; ++i) { pattern1024[2 * i] = kPatX[i]; pattern1024[2 * i + 1] = kPatY[i]; }
}

// IC_Angle + computeOrbDescriptor (src/ORBextractor.cc:78-148) for n key-points at level coordinates xy of one level image (tight pitch w)
// and its blurred copy; the caller keeps the points 19 pixels inside.
void orb_oracle_orient_describe(const uint8_t* img, const uint8_t* blurred, int w, int h, int n, const float* xy, float* angle_out, uint8_t* desc_out) {
    (void)h;
    Extractor ex(2000, 1.2f, 8, 20, 7);
    for (int i = 0; i < n; ++i) {
        angle_out[i] = ic_angle(img, w, xy[2 * i], xy[2 * i + 1], ex.umax);
        rbrief(blurred, w, xy[2 * i], xy[2 * i + 1], angle_out[i], desc_out + 32 * i);
    }
}

// Full extractor.  kps: cap records of 24 B; desc: cap x 32 B.  Optional outputs:
//   pyr_out   -- the nlevels ROIs packed back to back (tight pitch = level width)
//   cand_counts[nlevels] -- number of FAST candidates per level before distribution
// Returns the number of key-points, or -1 if cap was too small.
int orb_oracle_extract(const uint8_t* img, int w, int h, int pitch, const uint8_t* mask, int mpitch,
                       int nfeatures, float sf, int nlevels, int iniTh, int minTh,
                       void* kps, uint8_t* desc, int cap, uint8_t* pyr_out, int32_t* cand_counts) {
    Extractor ex(nfeatures, sf, nlevels, iniTh, minTh);
    std::vector<Level> pyr;
    g_bad_aspect = false;
    const int n = extract(ex, img, w, h, pitch, mask, mpitch, (KeyPoint*)kps, desc, cap, &pyr, cand_counts);
    if (g_bad_aspect) return -2;   // same contract as adb_orb_create: ADB_ERR_INVALID for portrait shapes with nIni = 0
    if (pyr_out) {
        size_t o = 0;
        for (int l = 0; l < nlevels; ++l)
            for (int y = 0; y < pyr[l].h; ++y) { std::memcpy(pyr_out + o, pyr[l].roi() + (size_t)y * pyr[l].pitch, pyr[l].w); o += pyr[l].w; }
    }
    return n;
}

// Batch version used for CPU-baseline timing: n_frames images (frame_stride apart), `threads`
// worker threads each taking whole frames (the reference runs one extractor per image thread,
// src/Frame.cc:81-84).  counts[f] receives the per-frame key-point count.
int orb_oracle_extract_batch(const uint8_t* imgs, int n_frames, size_t frame_stride, int w, int h, int pitch,
                             int nfeatures, float sf, int nlevels, int iniTh, int minTh,
                             void* kps, uint8_t* desc, int cap_per_frame, int32_t* counts, int threads) {
    Extractor ex(nfeatures, sf, nlevels, iniTh, minTh);
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (int f = t; f < n_frames; f += threads)
                counts[f] = extract(ex, imgs + (size_t)f * frame_stride, w, h, pitch, nullptr, 0,
                                    (KeyPoint*)kps + (size_t)f * cap_per_frame, desc + (size_t)f * cap_per_frame * 32,
                                    cap_per_frame, nullptr, nullptr);
        });
    for (auto& th : pool) th.join();
    return 0;
}

}  // extern "C"
