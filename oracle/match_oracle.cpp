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
