// oracle/match_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.cpp header).
//
// CPU restatement of the descriptor-matching arithmetic on the AirDOS hot path:
//   * ORBmatcher::DescriptorDistance                      src/ORBmatcher.cc:1647-1663
//   * the best / second-best scan all Search* share       src/ORBmatcher.cc:85-114
//   * Frame::ComputeStereoMatches                         src/Frame.cc:829-1003
// Parity pinning: the reference has no tests for these; the arithmetic is pure integer
// (XOR + popcount) plus small exactly-representable float sums, so the restatement is checked
// against numpy's bit counting in tests/test_oracle_match.py.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace {

struct KeyPoint { float x, y, size, angle, response; int32_t octave; };

// The reference's SWAR bit-count over eight 32-bit words.
inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        std::memcpy(&wa, a + 4 * i, 4);
        std::memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

constexpr int TH_HIGH = 100, TH_LOW = 50;  // src/ORBmatcher.cc:37-38

}  // namespace

extern "C" {

int match_oracle_distance(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

// Best / second-best Hamming scan with the reference's update rule (strict '<': the first
// candidate in list order wins ties; src/ORBmatcher.cc:102-114):
//     if d < best { second = best; best = d; idx = c } else if d < second { second = d }
// starting from best = second = 256, idx = -1.  cand_off == NULL means "all targets, in order".
void match_oracle_best2(const uint8_t* Q, int nq, const uint8_t* T, int nt, const int32_t* cand_off,
                        const int32_t* cand_idx, int32_t* best_idx, int32_t* best_d, int32_t* second_d) {
    for (int q = 0; q < nq; ++q) {
        int best = 256, second = 256, idx = -1;
        const int lo = cand_off ? cand_off[q] : 0, hi = cand_off ? cand_off[q + 1] : nt;
        for (int c = lo; c < hi; ++c) {
            const int t = cand_off ? cand_idx[c] : c;
            const int d = hamming256(Q + (size_t)q * 32, T + (size_t)t * 32);
            if (d < best) { second = best; best = d; idx = t; }
            else if (d < second) second = d;
        }
        best_idx[q] = idx; best_d[q] = best; second_d[q] = second;
    }
}

// Frame::ComputeStereoMatches.  Pyramids are passed as packed ROIs (pitch == level width),
// level l of the left image at pyrL + off[l].  Outputs per left key-point:
//   uRight, depth                 (mvuRight / mvDepth, -1 when unmatched)
//   ham_idx, ham_dist             best right index / distance of the Hamming stage
//                                 (idx -1 and dist TH_HIGH when no candidate passed)
// stage: 0 = everything, 1 = Hamming stage only (uRight/depth left at -1).
void match_oracle_stereo(const void* kl_, const uint8_t* dl, int nl, const void* kr_, const uint8_t* dr, int nr,
                         const uint8_t* pyrL, const uint8_t* pyrR, const int64_t* off, const int32_t* lw,
                         const int32_t* lh, int nlevels, const float* scaleF, const float* invScaleF,
                         float mb, float mbf, int stage, float* uRight, float* depth, int32_t* ham_idx,
                         int32_t* ham_dist) {
    const KeyPoint* kl = (const KeyPoint*)kl_;
    const KeyPoint* kr = (const KeyPoint*)kr_;
    for (int i = 0; i < nl; ++i) { uRight[i] = -1.f; depth[i] = -1.f; ham_idx[i] = -1; ham_dist[i] = TH_HIGH; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = lh[0];
    std::vector<std::vector<int>> rows(nRows);
    for (int iR = 0; iR < nr; ++iR) {
        const float y = kr[iR].y;
        const float r = 2.0f * scaleF[kr[iR].octave];
        const int maxr = std::min((int)std::ceil(y + r), nRows - 1);   // clamp: convention D.9
        const int minr = std::max((int)std::floor(y - r), 0);
        for (int yi = minr; yi <= maxr; ++yi) rows[yi].push_back(iR);
    }
    const float minZ = mb, minD = 0.f, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < nl; ++iL) {
        const KeyPoint& kp = kl[iL];
        const int levelL = kp.octave;
        const float vL = kp.y, uL = kp.x;
        const std::vector<int>& cands = rows[(size_t)vL];
        if (cands.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        bool any = false;
        for (int iR : cands) {
            const KeyPoint& k2 = kr[iR];
            if (k2.octave < levelL - 1 || k2.octave > levelL + 1) continue;
            const float uR = k2.x;
            if (uR >= minU && uR <= maxU) {
                const int d = hamming256(dl + (size_t)iL * 32, dr + (size_t)iR * 32);
                if (d < bestDist) { bestDist = d; bestIdxR = iR; any = true; }
            }
        }
        if (any) { ham_idx[iL] = bestIdxR; ham_dist[iL] = bestDist; }
        if (stage == 1) continue;
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float sf = invScaleF[kp.octave];
            const float scaleduL = std::round(kp.x * sf);
            const float scaledvL = std::round(kp.y * sf);
            const float scaleduR0 = std::round(uR0 * sf);
            const int w = 5, L = 5;
            const int lvlW = lw[kp.octave];
            const uint8_t* IL = pyrL + off[kp.octave];
            const uint8_t* IR = pyrR + off[kp.octave];
            const int cy = (int)scaledvL, cxL = (int)scaleduL;
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= (float)lvlW) continue;
            const float cL = (float)IL[(size_t)cy * lvlW + cxL];
            int bestSad = INT_MAX, bestinc = 0;
            float dists[2 * 5 + 1];
            for (int inc = -L; inc <= L; ++inc) {
                const int cxR = (int)scaleduR0 + inc;
                const float cR = (float)IR[(size_t)cy * lvlW + cxR];
                double acc = 0.0;  // cv::norm(NORM_L1) accumulates float data in double
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        const float a = (float)IL[(size_t)(cy + dy) * lvlW + cxL + dx] - cL;
                        const float b = (float)IR[(size_t)(cy + dy) * lvlW + cxR + dx] - cR;
                        acc += std::fabs((double)(a - b));
                    }
                const float dist = (float)acc;
                if (dist < (float)bestSad) { bestSad = (int)dist; bestinc = inc; }
                dists[L + inc] = dist;
            }
            if (bestinc == -L || bestinc == L) continue;
            const float d1 = dists[L + bestinc - 1], d2 = dists[L + bestinc], d3 = dists[L + bestinc + 1];
            const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scaleF[kp.octave] * ((float)scaleduR0 + (float)bestinc + deltaR);
            float disparity = uL - bestuR;
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = (float)0.01; bestuR = (float)((double)uL - 0.01); }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                distIdx.push_back({bestSad, iL});
            }
        }
    }
    if (stage == 1 || distIdx.empty()) return;  // empty guard: convention D.8
    std::sort(distIdx.begin(), distIdx.end());
    const float median = (float)distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
        if ((float)distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1.f;
        depth[distIdx[i].second] = -1.f;
    }
}

}  // extern "C"

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:245-310), batched: descriptors of all observations in CSR
// form (point_ptr); per point the row of the N x N Hamming matrix with the least median (sorted[0.5 * (N - 1)], self
// distance 0 included) wins, first row on ties.  best_idx = -1 for a point without observations.
extern "C" void match_oracle_distinctive(const uint8_t* desc, const int32_t* point_ptr, int n_points, int32_t* best_idx) {
    std::vector<int> d, row;
    for (int p = 0; p < n_points; ++p) {
        const int a = point_ptr[p], N = point_ptr[p + 1] - a;
        if (N <= 0) { best_idx[p] = -1; continue; }
        d.assign((size_t)N * N, 0);
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) {
                const int v = hamming256(desc + (size_t)(a + i) * 32, desc + (size_t)(a + j) * 32);
                d[(size_t)i * N + j] = v; d[(size_t)j * N + i] = v;
            }
        int bestMedian = INT_MAX, bestIdx = 0;
        for (int i = 0; i < N; ++i) {
            row.assign(d.begin() + (size_t)i * N, d.begin() + (size_t)(i + 1) * N);
            std::sort(row.begin(), row.end());
            const int median = row[(size_t)(0.5 * (N - 1))];
            if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        }
        best_idx[p] = bestIdx;
    }
}

// =========================================================================================
// Guided window searches: ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)
// (src/ORBmatcher.cc:45-129) and ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last,
// th, bMono) (src/ORBmatcher.cc:1328-1470), on the Frame grid of src/Frame.cc:534-549, 645-712.
// Both are the same loop: for every query (a map point with a predicted pixel, a window radius and a
// level range) take the key-points of the 64 x 48 grid cells under the window in the reference's order
// (cell column, cell row, insertion order), skip the ones already held by an observed map point, check
// the right coordinate, keep best / second-best Hamming, accept, assign -- sequentially, so an earlier
// query can take a key-point away from a later one.  The restatement keeps that sequential structure
// (std::vector grid, one query after the other); the CUDA kernel reaches the same fixed point in parallel.
namespace {

constexpr int GRID_COLS = 64, GRID_ROWS = 48, HISTO_LENGTH = 30;   // include/Frame.h:38-39, src/ORBmatcher.cc:39

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
};

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:534-549, 700-712)
void assign_to_grid(const KeyPoint* kps, int n, float minX, float minY, float invW, float invH, Grid& g) {
    for (int i = 0; i < n; ++i) {
        const int px = (int)std::round((kps[i].x - minX) * invW);
        const int py = (int)std::round((kps[i].y - minY) * invH);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        g.cell[px][py].push_back(i);
    }
}

// Frame::GetFeaturesInArea (src/Frame.cc:645-698)
void features_in_area(const Grid& g, const KeyPoint* kps, float minX, float minY, float invW, float invH, float x, float y,
                      float r, int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
            for (int idx : g.cell[ix][iy]) {
                const KeyPoint& kp = kps[idx];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float distx = kp.x - x, disty = kp.y - y;
                if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
            }
}

}  // namespace

extern "C" {

// q_flags: bit 0 = the query takes part (mbTrackInView && !isBad / has a map point && !outlier && in front && inside the
// image), bit 1 = its map point has Observations() > 0 (a key-point it takes is closed to later queries).
// taken[i] != 0: CurrentFrame.mvpMapPoints[i] is set on entry and has Observations() > 0.
// use_ratio != 0: the map-point variant (second-best + same-level ratio rule, nn_ratio = mfNNratio);
// check_ori != 0: the last-frame variant's rotation histogram (ComputeThreeMaxima, src/ORBmatcher.cc:1601-1642).
// kp_match[i]: -1 untouched, -2 cleared by the rotation check, >= 0 the query that holds key-point i at the end.
// Returns nmatches.
int match_oracle_search_projection(const void* kps_, const float* u_right, const uint8_t* desc, const uint8_t* taken, int n_kp,
                                   float minX, float minY, float invW, float invH, int n_q, const float* q_u, const float* q_v,
                                   const float* q_ur, const float* q_radius, const int32_t* q_minl, const int32_t* q_maxl,
                                   const uint8_t* q_desc, const uint8_t* q_flags, const float* q_angle, int use_ratio,
                                   float nn_ratio, int check_ori, int32_t* kp_match, int32_t* q_best_idx, int32_t* q_best_dist) {
    const KeyPoint* kps = (const KeyPoint*)kps_;
    Grid* g = new Grid();
    assign_to_grid(kps, n_kp, minX, minY, invW, invH, *g);
    std::vector<int> holder(n_kp, -1);          // query index written into mvpMapPoints[i] (-1: not written here)
    std::vector<uint8_t> closed(n_kp, 0);       // mvpMapPoints[i] && Observations() > 0 right now
    for (int i = 0; i < n_kp; ++i) closed[i] = taken ? (taken[i] != 0) : 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    std::vector<int> cand;
    for (int q = 0; q < n_q; ++q) {
        q_best_idx[q] = -1; q_best_dist[q] = 256;
        if (!(q_flags[q] & 1)) continue;
        features_in_area(*g, kps, minX, minY, invW, invH, q_u[q], q_v[q], q_radius[q], q_minl[q], q_maxl[q], cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (closed[idx]) continue;
            if (u_right[idx] > 0) {
                const float er = std::fabs(q_ur[q] - u_right[idx]);
                if (er > q_radius[q]) continue;
            }
            const int dist = hamming256(q_desc + (size_t)q * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = kps[idx].octave; bestDist2 = dist; }
        }
        q_best_idx[q] = bestIdx; q_best_dist[q] = bestDist;
        if (bestDist <= TH_HIGH) {
            if (use_ratio && bestLevel == bestLevel2 && (float)bestDist > nn_ratio * (float)bestDist2) continue;
            holder[bestIdx] = q;
            closed[bestIdx] = (q_flags[q] & 2) != 0;
            ++nmatches;
            if (check_ori) {
                float rot = q_angle[q] - kps[bestIdx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx);
            }
        }
    }
    for (int i = 0; i < n_kp; ++i) kp_match[i] = holder[i];
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { kp_match[idx] = -2; --nmatches; }
    }
    delete g;
    return nmatches;
}

// The projection at the head of SearchByProjection(CurrentFrame, LastFrame) (src/ORBmatcher.cc:1338-1393): queries from
// the last frame's map points.  cv::Mat arithmetic convention: Rcw * x + tcw is one cv::gemm, which accumulates float
// products in double and rounds once (OpenCV GEMMSingleMul<float, double>); everything after it is float without
// contraction.  tcw_cur / tcw_last: row-major 4x4 float Tcw.  last_flags bit 0: map point present and not an outlier,
// bit 1: Observations() > 0.  Outputs the query arrays of match_oracle_search_projection.
void match_oracle_project_last(const float* tcw_cur, const float* tcw_last, int n, const float* xw, const int32_t* octave,
                               const uint8_t* last_flags, float fx, float fy, float cx, float cy, float mbf, float mb,
                               float minX, float maxX, float minY, float maxY, const float* scale_factors, float th, int mono,
                               float* q_u, float* q_v, float* q_ur, float* q_radius, int32_t* q_minl, int32_t* q_maxl,
                               uint8_t* q_flags) {
    auto R = [](const float* T, int r, int c) { return (double)T[4 * r + c]; };
    // twc = -Rcw^T * tcw ; tlc = Rlw * twc + tlw
    float twc[3], tlc[3];
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (-R(tcw_cur, k, i)) * R(tcw_cur, k, 3);   // (-Rcw.t()) is exact, then one gemm
        twc[i] = (float)s;
    }
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += R(tcw_last, i, k) * (double)twc[k];
        tlc[i] = (float)(s + R(tcw_last, i, 3));
    }
    const bool bForward = tlc[2] > mb && !mono, bBackward = -tlc[2] > mb && !mono;
    for (int i = 0; i < n; ++i) {
        q_u[i] = q_v[i] = q_ur[i] = q_radius[i] = 0.f; q_minl[i] = 0; q_maxl[i] = -1; q_flags[i] = 0;
        if (!(last_flags[i] & 1)) continue;
        float xc3[3];
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += R(tcw_cur, r, k) * (double)xw[3 * i + k];
            xc3[r] = (float)(s + R(tcw_cur, r, 3));
        }
        const float xc = xc3[0], yc = xc3[1];
        const float invzc = (float)(1.0 / (double)xc3[2]);
        if (invzc < 0) continue;
        const float u = fx * xc * invzc + cx, v = fy * yc * invzc + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const int lo = octave[i];
        q_u[i] = u; q_v[i] = v;
        q_ur[i] = u - mbf * invzc;
        q_radius[i] = th * scale_factors[lo];
        if (bForward) { q_minl[i] = lo; q_maxl[i] = -1; }
        else if (bBackward) { q_minl[i] = 0; q_maxl[i] = lo; }
        else { q_minl[i] = lo - 1; q_maxl[i] = lo + 1; }
        q_flags[i] = (uint8_t)(1 | (last_flags[i] & 2));
    }
}

// Frame::isInFrustum (src/Frame.cc:587-643) + MapPoint::PredictScale (src/MapPoint.cc:405-420) + the window radius of
// SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:55-69) -> the query arrays.  cv::Mat conventions: Rcw * P + tcw is
// one gemm (double accumulation, one rounding); P - Ow is float; cv::norm and Mat::dot accumulate in double.  The log of
// PredictScale is evaluated in double on the float ratio and rounded once (the reference's unqualified log() may bind to
// either overload, and libm / CUDA logf differ in the last bit).  in_flags bit 0: the point reaches the test.
void match_oracle_frustum(const float* tcw, const float* ow, int n, const float* xw, const float* normal, const float* min_distance,
                          const float* max_distance, const uint8_t* in_flags, float fx, float fy, float cx, float cy, float mbf,
                          float minX, float maxX, float minY, float maxY, float view_cos_limit, float log_scale_factor,
                          const float* scale_factors, int n_levels, float th, float* q_u, float* q_v, float* q_ur, float* q_radius,
                          int32_t* q_minl, int32_t* q_maxl, uint8_t* q_flags, float* q_track, int32_t* q_level) {
    for (int i = 0; i < n; ++i) {
        q_u[i] = q_v[i] = q_ur[i] = q_radius[i] = 0.f; q_minl[i] = 0; q_maxl[i] = -1; q_flags[i] = 0;
        q_level[i] = -1; for (int k = 0; k < 4; ++k) q_track[4 * i + k] = 0.f;
        if (!(in_flags[i] & 1)) continue;
        const float* P = xw + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)tcw[4 * r + k] * (double)P[k];
            Pc[r] = (float)(s + (double)tcw[4 * r + 3]);
        }
        if (Pc[2] < 0.0f) continue;
        const float invz = 1.0f / Pc[2];
        const float u = fx * Pc[0] * invz + cx, v = fy * Pc[1] * invz + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const float maxD = 1.2f * max_distance[i], minD = 0.8f * min_distance[i];
        const float PO[3] = {P[0] - ow[0], P[1] - ow[1], P[2] - ow[2]};
        const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        if (dist < minD || dist > maxD) continue;
        const float* Pn = normal + 3 * i;
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];
        const float viewCos = (float)(dot / (double)dist);
        if (viewCos < view_cos_limit) continue;
        const float ratio = max_distance[i] / dist;
        int nScale = (int)std::ceil((float)std::log((double)ratio) / log_scale_factor);
        if (nScale < 0) nScale = 0; else if (nScale >= n_levels) nScale = n_levels - 1;
        float r = (double)viewCos > 0.998 ? 2.5f : 4.0f;   // RadiusByViewingCos compares against a double literal
        if (th != 1.0f) r *= th;
        q_u[i] = u; q_v[i] = v; q_ur[i] = u - mbf * invz;
        q_radius[i] = r * scale_factors[nScale];
        q_minl[i] = nScale - 1; q_maxl[i] = nScale;
        q_flags[i] = (uint8_t)(1 | (in_flags[i] & 2));
        q_track[4 * i] = u; q_track[4 * i + 1] = v; q_track[4 * i + 2] = q_ur[i]; q_track[4 * i + 3] = viewCos;
        q_level[i] = nScale;
    }
}

// The search half of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:825-975) with
// KeyFrame::GetFeaturesInArea / IsInImage (src/KeyFrame.cc:589-633): per map point the key-point it would be fused with.
// Same cv::Mat conventions as match_oracle_frustum.  Returns nFused (queries with bestDist <= TH_LOW).
int match_oracle_fuse_search(const void* kps_, const float* u_right, const uint8_t* desc, int n_kp, float minX, float minY, float maxX,
                             float maxY, float invW, float invH, const float* tcw, const float* ow, int n, const float* xw,
                             const float* normal, const float* min_distance, const float* max_distance, const uint8_t* in_flags,
                             const uint8_t* q_desc, float fx, float fy, float cx, float cy, float mbf, float log_scale_factor,
                             const float* scale_factors, const float* inv_level_sigma2, int n_levels, float th, int32_t* kp_match,
                             int32_t* q_best_idx, int32_t* q_best_dist) {
    const KeyPoint* kps = (const KeyPoint*)kps_;
    Grid* g = new Grid();
    assign_to_grid(kps, n_kp, minX, minY, invW, invH, *g);
    for (int i = 0; i < n_kp; ++i) kp_match[i] = -1;
    int nFused = 0;
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        q_best_idx[i] = -1; q_best_dist[i] = 256;
        if (!(in_flags[i] & 1)) continue;
        const float* P = xw + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)tcw[4 * r + k] * (double)P[k];
            Pc[r] = (float)(s + (double)tcw[4 * r + 3]);
        }
        if (Pc[2] < 0.0f) continue;
        const float invz = 1 / Pc[2];
        const float x = Pc[0] * invz, y = Pc[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!(u >= minX && u < maxX && v >= minY && v < maxY)) continue;
        const float ur = u - mbf * invz;
        const float maxD = 1.2f * max_distance[i], minD = 0.8f * min_distance[i];
        const float PO[3] = {P[0] - ow[0], P[1] - ow[1], P[2] - ow[2]};
        const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        if (dist3D < minD || dist3D > maxD) continue;
        const float* Pn = normal + 3 * i;
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];
        if (dot < 0.5 * (double)dist3D) continue;
        const float ratio = max_distance[i] / dist3D;
        int lvl = (int)std::ceil((float)std::log((double)ratio) / log_scale_factor);
        if (lvl < 0) lvl = 0; else if (lvl >= n_levels) lvl = n_levels - 1;
        const float radius = th * scale_factors[lvl];
        features_in_area(*g, kps, minX, minY, invW, invH, u, v, radius, -1, -1, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const KeyPoint& kp = kps[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            if (u_right[idx] >= 0) {
                const float ex = u - kp.x, ey = v - kp.y, er = ur - u_right[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if ((double)(e2 * inv_level_sigma2[kpLevel]) > 7.8) continue;
            } else {
                const float ex = u - kp.x, ey = v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if ((double)(e2 * inv_level_sigma2[kpLevel]) > 5.99) continue;
            }
            const int dist = hamming256(q_desc + (size_t)i * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        q_best_idx[i] = bestIdx; q_best_dist[i] = bestDist;
        if (bestDist <= TH_LOW) { kp_match[bestIdx] = i; ++nFused; }
    }
    delete g;
    return nFused;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:159-288) and ORBmatcher::SearchForTriangulation
// (src/ORBmatcher.cc:657-823, CheckDistEpipolarLine :131-157) on the bucket lists the FeatureVector walk produces.  Sequential
// like the reference.  Note SearchForTriangulation never sets vbMatched2 (:680, :729): two key-points of frame 1 may end up
// with the same partner, exactly as in the reference.  Its update rule "dist > bestDist -> skip" lets a LATER candidate
// with an equal distance replace the best one.
int match_oracle_bow_search(int mode, const void* kps1_, const float* ur1, const uint8_t* d1, const uint8_t* fl1, int n1,
                            const void* kps2_, const float* ur2, const uint8_t* d2, const uint8_t* fl2, int n2, int n_buckets,
                            const int32_t* p1, const int32_t* i1, const int32_t* p2, const int32_t* i2, float nn_ratio, int check_ori,
                            const float* F12, float ex, float ey, const float* scale_factors2, const float* level_sigma2_2,
                            int32_t* match21, int32_t* match12) {
    const KeyPoint* k1 = (const KeyPoint*)kps1_;
    const KeyPoint* k2 = (const KeyPoint*)kps2_;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    if (mode == 0) for (int i = 0; i < n2; ++i) match21[i] = -1;
    else for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (int b = 0; b < n_buckets; ++b) {
        for (int a = p1[b]; a < p1[b + 1]; ++a) {
            const int idx1 = i1[a];
            if (!(fl1[idx1] & 1)) continue;
            if (mode == 0) {
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int c = p2[b]; c < p2[b + 1]; ++c) {
                    const int idxF = i2[c];
                    if (match21[idxF] >= 0) continue;
                    const int dist = hamming256(d1 + (size_t)idx1 * 32, d2 + (size_t)idxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = idxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW && (float)bestDist1 < nn_ratio * (float)bestDist2) {
                    match21[bestIdxF] = idx1;
                    if (check_ori) {
                        float rot = k1[idx1].angle - k2[bestIdxF].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(bestIdxF);
                    }
                    ++nmatches;
                }
            } else {
                const bool bStereo1 = ur1[idx1] >= 0;
                const KeyPoint& kp1 = k1[idx1];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int c = p2[b]; c < p2[b + 1]; ++c) {
                    const int idx2 = i2[c];
                    if (!(fl2[idx2] & 1)) continue;
                    const bool bStereo2 = ur2[idx2] >= 0;
                    const int dist = hamming256(d1 + (size_t)idx1 * 32, d2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const KeyPoint& kp2 = k2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * scale_factors2[kp2.octave]) continue;
                    }
                    // CheckDistEpipolarLine
                    const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                    const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                    const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                    const float num = la * kp2.x + lb * kp2.y + lc;
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if ((double)dsqr < 3.84 * (double)level_sigma2_2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    ++nmatches;
                    if (check_ori) {
                        float rot = kp1.angle - k2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                }
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rotHist[i]) { if (mode == 0) match21[idx] = -1; else match12[idx] = -1; --nmatches; }
    }
    return nmatches;
}

}  // extern "C"
