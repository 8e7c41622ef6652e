// search.cu -- guided searches of ORBmatcher with the candidate generation on the device.
//
//   project_last_kernel    head of SearchByProjection(Current, Last)     src/ORBmatcher.cc:1338-1393
//   frustum_kernel         Frame::isInFrustum + MapPoint::PredictScale   src/Frame.cc:587-643, src/MapPoint.cc:405-420
//                          (and the projection head of ORBmatcher::Fuse  src/ORBmatcher.cc:845-889)
//   bow_search_kernel      SearchByBoW(KF, Frame) / SearchForTriangulation + CheckDistEpipolarLine
//                                                                        src/ORBmatcher.cc:159-288, 657-823, 131-157
//   proj_search_kernel     Frame::AssignFeaturesToGrid / PosInGrid       src/Frame.cc:534-549, 700-712
//                          Frame::GetFeaturesInArea                      src/Frame.cc:645-698
//                          SearchByProjection(Frame, vpMapPoints, th)    src/ORBmatcher.cc:45-129
//                          SearchByProjection(Current, Last, th, bMono)  src/ORBmatcher.cc:1395-1467
//                          ComputeThreeMaxima                            src/ORBmatcher.cc:1601-1642
//
// The reference walks the queries one after the other because a query whose map point is already observed closes the
// key-point it takes to every later query.  That dependence only runs from lower to higher query indices, so the
// sequential result is the unique fixed point of "every query searches with the closures of the previous round"; one
// thread block per frame iterates rounds (all queries in parallel, one warp per query) until the closures stop
// changing -- query 0 is final after round 1, query 1 after round 2 at the latest, in practice two or three rounds.
// Candidate order (cell column, cell row, insertion order) and the strict '<' updates make best / second-best the two
// smallest (distance, position) pairs, which lanes compute on disjoint cells and merge with shuffles.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "matcher.cuh"

namespace adb {

constexpr int kGridCols = 64, kGridRows = 48, kGridCells = kGridCols * kGridRows;   // include/Frame.h:38-39
constexpr int kHisto = 30;                                                          // HISTO_LENGTH src/ORBmatcher.cc:39
constexpr int kThHigh = 100, kThLow = 50;                                           // TH_HIGH / TH_LOW src/ORBmatcher.cc:37-38
constexpr int kSearchThreads = 1024;
constexpr uint32_t kNoBlock = 0x7FFFFFFFu;

// Device view of one problem: every pointer is device memory inside the call's scratch block.
struct SearchDev {
    int n_kp, n_q;
    const adb_keypoint* kps;
    const float* u_right;
    const uint8_t* desc;
    const uint8_t* taken;
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    float* q_u; float* q_v; float* q_ur; float* q_radius;
    int32_t* q_minl; int32_t* q_maxl;
    uint8_t* q_flags;
    const uint8_t* q_desc;
    const float* q_angle;
    int use_ratio; float nn_ratio; int check_ori;
    // projection (variant 2)
    const float* last_xw; const int32_t* last_octave; const uint8_t* last_flags;
    float Rcw[9], tcw[3];
    float fx, fy, cx, cy, mbf, th;
    const float* scale_factors;
    int forward, backward;
    // visibility test (variant 1 with Frame::isInFrustum on the device)
    const float* mp_xw; const float* mp_normal; const float* mp_min_distance; const float* mp_max_distance; const uint8_t* mp_flags;
    float ow[3], Tcw[12];
    float view_cos_limit, log_scale_factor;
    int n_levels;
    float* q_track; int32_t* q_level;
    int fuse; const float* inv_sigma2;     // ORBmatcher::Fuse candidate search (src/ORBmatcher.cc:825-975)
    // results
    int32_t* kp_match; int32_t* q_best_idx; int32_t* q_best_dist; int32_t* q_choice; int32_t* n_matches;
};

// Rcw * x + tcw as cv::gemm does it for float matrices: products and sums in double, one rounding (oracle:
// match_oracle_project_last).  Everything after it is float, no contraction.
__global__ void project_last_kernel(const SearchDev* __restrict__ probs) {
    const SearchDev& P = probs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!P.last_xw || i >= P.n_q) return;
    float u = 0.f, v = 0.f, ur = 0.f, rad = 0.f;
    int minl = 0, maxl = -1;
    uint8_t fl = 0;
    const uint8_t lf = P.last_flags[i];
    if (lf & 1) {
        const double x = P.last_xw[3 * i], y = P.last_xw[3 * i + 1], z = P.last_xw[3 * i + 2];
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double s = __dmul_rn((double)P.Rcw[3 * r], x);
            s = __dadd_rn(s, __dmul_rn((double)P.Rcw[3 * r + 1], y));
            s = __dadd_rn(s, __dmul_rn((double)P.Rcw[3 * r + 2], z));
            c[r] = (float)__dadd_rn(s, (double)P.tcw[r]);
        }
        const float invz = (float)__ddiv_rn(1.0, (double)c[2]);
        if (!(invz < 0)) {
            u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c[0]), invz), P.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c[1]), invz), P.cy);
            if (!(u < P.min_x || u > P.max_x) && !(v < P.min_y || v > P.max_y)) {
                const int lo = P.last_octave[i];
                ur = __fsub_rn(u, __fmul_rn(P.mbf, invz));
                rad = __fmul_rn(P.th, P.scale_factors[lo]);
                if (P.forward) { minl = lo; maxl = -1; }
                else if (P.backward) { minl = 0; maxl = lo; }
                else { minl = lo - 1; maxl = lo + 1; }
                fl = (uint8_t)(1 | (lf & 2));
            } else { u = 0.f; v = 0.f; }
        }
    }
    P.q_u[i] = u; P.q_v[i] = v; P.q_ur[i] = ur; P.q_radius[i] = rad; P.q_minl[i] = minl; P.q_maxl[i] = maxl; P.q_flags[i] = fl;
}

// Frame::isInFrustum + MapPoint::PredictScale + the window radius of SearchByProjection(F, vpMapPoints, th); conventions as in
// oracle/match_oracle.cpp (match_oracle_frustum): gemm / norm / dot in double with one rounding, log in double rounded once.
__global__ void frustum_kernel(const SearchDev* __restrict__ probs) {
    const SearchDev& P = probs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (!P.mp_xw || i >= P.n_q) return;
    float u = 0.f, v = 0.f, ur = 0.f, rad = 0.f, vc = 0.f;
    int minl = 0, maxl = -1, level = -1;
    uint8_t fl = 0;
    const uint8_t inf = P.mp_flags[i];
    if (inf & 1) {
        const float px = P.mp_xw[3 * i], py = P.mp_xw[3 * i + 1], pz = P.mp_xw[3 * i + 2];
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double s = __dmul_rn((double)P.Tcw[4 * r], (double)px);
            s = __dadd_rn(s, __dmul_rn((double)P.Tcw[4 * r + 1], (double)py));
            s = __dadd_rn(s, __dmul_rn((double)P.Tcw[4 * r + 2], (double)pz));
            c[r] = (float)__dadd_rn(s, (double)P.Tcw[4 * r + 3]);
        }
        bool ok = !(c[2] < 0.0f);
        const float invz = __fdiv_rn(1.0f, c[2]);
        float uu, vv;
        if (P.fuse) {   // Fuse: x = Xc * invz first, IsInImage is half open (src/ORBmatcher.cc:859-868, src/KeyFrame.cc:630-633)
            uu = __fadd_rn(__fmul_rn(P.fx, __fmul_rn(c[0], invz)), P.cx);
            vv = __fadd_rn(__fmul_rn(P.fy, __fmul_rn(c[1], invz)), P.cy);
            ok = ok && (uu >= P.min_x && uu < P.max_x && vv >= P.min_y && vv < P.max_y);
        } else {
            uu = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c[0]), invz), P.cx);
            vv = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c[1]), invz), P.cy);
            ok = ok && !(uu < P.min_x || uu > P.max_x) && !(vv < P.min_y || vv > P.max_y);
        }
        const float maxD = __fmul_rn(1.2f, P.mp_max_distance[i]), minD = __fmul_rn(0.8f, P.mp_min_distance[i]);
        const float ox = __fsub_rn(px, P.ow[0]), oy = __fsub_rn(py, P.ow[1]), oz = __fsub_rn(pz, P.ow[2]);
        const double n2 = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)ox), __dmul_rn((double)oy, (double)oy)), __dmul_rn((double)oz, (double)oz));
        const float dist = (float)__dsqrt_rn(n2);
        ok = ok && !(dist < minD || dist > maxD);
        const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)P.mp_normal[3 * i]), __dmul_rn((double)oy, (double)P.mp_normal[3 * i + 1])),
                                     __dmul_rn((double)oz, (double)P.mp_normal[3 * i + 2]));
        const float viewCos = (float)__ddiv_rn(dot, (double)dist);
        if (P.fuse) ok = ok && !(dot < __dmul_rn(0.5, (double)dist));   // PO.dot(Pn) < 0.5 * dist3D (src/ORBmatcher.cc:883)
        else ok = ok && !(viewCos < P.view_cos_limit);
        if (ok) {
            const float ratio = __fdiv_rn(P.mp_max_distance[i], dist);
            int ns = (int)ceilf(__fdiv_rn((float)log((double)ratio), P.log_scale_factor));
            if (ns < 0) ns = 0; else if (ns >= P.n_levels) ns = P.n_levels - 1;
            float r = (double)viewCos > 0.998 ? 2.5f : 4.0f;
            if (P.th != 1.0f) r = __fmul_rn(r, P.th);
            if (P.fuse) r = P.th;                                       // radius = th * mvScaleFactors[level] (src/ORBmatcher.cc:889)
            u = uu; v = vv; ur = __fsub_rn(uu, __fmul_rn(P.mbf, invz)); vc = viewCos;
            rad = __fmul_rn(r, P.scale_factors[ns]);
            minl = ns - 1; maxl = ns; level = ns;
            fl = (uint8_t)(P.fuse ? 1 : (1 | (inf & 2)));
        }
    }
    P.q_u[i] = u; P.q_v[i] = v; P.q_ur[i] = ur; P.q_radius[i] = rad; P.q_minl[i] = minl; P.q_maxl[i] = maxl; P.q_flags[i] = fl;
    P.q_track[4 * i] = u; P.q_track[4 * i + 1] = v; P.q_track[4 * i + 2] = ur; P.q_track[4 * i + 3] = vc;
    P.q_level[i] = level;
}

// 64-bit search key: distance << 43 | cell sequence << 31 | position in cell << 18 | key-point index << 5 | octave.
// Unique per candidate and ordered by (distance, position in the reference's candidate list).
__device__ __forceinline__ void key_insert(uint64_t k, uint64_t& best, uint64_t& second) {
    if (k < best) { second = best; best = k; }
    else if (k < second) second = k;
}

__global__ void __launch_bounds__(kSearchThreads) proj_search_kernel(const SearchDev* __restrict__ probs) {
    extern __shared__ __align__(16) uint8_t search_smem[];
    const SearchDev& P = probs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSearchThreads / 32;
    const int nk = P.n_kp, nq = P.n_q;
    // shared layout: cell_start[kGridCells + 1] | fill[kGridCells] | blk[2][nk] | items[nk] (u16) | cellid[nk] (u16)
    uint32_t* cell_start = reinterpret_cast<uint32_t*>(search_smem);
    uint32_t* fill = cell_start + kGridCells + 1;
    uint32_t* blk0 = fill + kGridCells;
    uint32_t* blk1 = blk0 + nk;
    uint16_t* items = reinterpret_cast<uint16_t*>(blk1 + nk);
    uint16_t* cellid = items + ((nk + 1) & ~1);
    __shared__ int s_hist[kHisto];
    __shared__ int s_keep[3];
    __shared__ int s_acc, s_removed;
    __shared__ uint32_t s_scan[kSearchThreads / 32];

    // ---- Frame::AssignFeaturesToGrid: counting sort of the key-points by cell, insertion (= index) order inside a cell
    for (int c = tid; c < kGridCells; c += kSearchThreads) { cell_start[c] = 0; fill[c] = 0; }
    if (tid < kHisto) s_hist[tid] = 0;
    if (tid == 0) { s_acc = 0; s_removed = 0; cell_start[kGridCells] = 0; }
    __syncthreads();
    for (int i = tid; i < nk; i += kSearchThreads) {
        const adb_keypoint kp = P.kps[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, P.min_x), P.inv_w));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, P.min_y), P.inv_h));
        int c = 0xFFFF;
        if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) { c = px * kGridRows + py; atomicAdd(&cell_start[c], 1u); }
        cellid[i] = (uint16_t)c;
    }
    __syncthreads();
    {   // exclusive scan of the 3072 counts: 3 per thread
        const int base = tid * 3;
        const uint32_t a = cell_start[base], b = cell_start[base + 1], c = cell_start[base + 2];
        const uint32_t mine = a + b + c;
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_scan[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += t; }
            s_scan[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t ex = s_scan[warp] + incl - mine;
        cell_start[base] = ex; cell_start[base + 1] = ex + a; cell_start[base + 2] = ex + a + b;
        if (tid == kSearchThreads - 1) cell_start[kGridCells] = ex + mine;
    }
    __syncthreads();
    for (int i = tid; i < nk; i += kSearchThreads) {
        const int c = cellid[i];
        if (c != 0xFFFF) items[cell_start[c] + atomicAdd(&fill[c], 1u)] = (uint16_t)i;
    }
    __syncthreads();
    for (int c = tid; c < kGridCells; c += kSearchThreads) {   // restore index order inside each cell (a handful of items)
        const int a = cell_start[c], b = cell_start[c + 1];
        for (int i = a + 1; i < b; ++i) {
            const uint16_t v = items[i];
            int j = i - 1;
            while (j >= a && items[j] > v) { items[j + 1] = items[j]; --j; }
            items[j + 1] = v;
        }
    }
    for (int i = tid; i < nk; i += kSearchThreads) blk0[i] = (P.taken && P.taken[i]) ? 0u : kNoBlock;
    __syncthreads();

    // ---- rounds.  blk[c] = 1 + the lowest blocking query that holds key-point c (0: closed on entry): query q may not
    //      take c when blk[c] <= q.
    uint32_t* prev = blk0;
    uint32_t* cur = blk1;
    for (int round = 0; round <= nq; ++round) {
        for (int i = tid; i < nk; i += kSearchThreads) cur[i] = (P.taken && P.taken[i]) ? 0u : kNoBlock;
        __syncthreads();
        for (int q = warp; q < nq; q += nwarps) {
            const uint8_t fl = P.q_flags[q];
            uint64_t best = ~0ull, second = ~0ull;
            if (fl & 1) {
                const float x = P.q_u[q], y = P.q_v[q], r = P.q_radius[q], urq = P.q_ur[q];
                const int minl = P.q_minl[q], maxl = P.q_maxl[q];
                // Frame::GetFeaturesInArea cell range, float arithmetic in the reference's order
                const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, P.min_x), r), P.inv_w)));
                const int cx1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, P.min_x), r), P.inv_w)));
                const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, P.min_y), r), P.inv_h)));
                const int cy1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, P.min_y), r), P.inv_h)));
                if (cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0 && cx1 >= cx0 && cy1 >= cy0) {
                    const bool check_levels = (minl > 0) || (maxl >= 0);
                    const int ncy = cy1 - cy0 + 1, ncell = (cx1 - cx0 + 1) * ncy;
                    uint32_t qd[8];
                    {
                        const uint4* qp = reinterpret_cast<const uint4*>(P.q_desc + (size_t)q * 32);
                        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
                        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
                    }
                    for (int k = lane; k < ncell; k += 32) {
                        const int ix = cx0 + k / ncy, iy = cy0 + k % ncy;
                        const int c = ix * kGridRows + iy;
                        const int a = cell_start[c], b = cell_start[c + 1];
                        for (int j = a; j < b; ++j) {
                            const int idx = items[j];
                            const adb_keypoint kp = P.kps[idx];
                            if (check_levels) {
                                if (kp.octave < minl) continue;
                                if (maxl >= 0 && kp.octave > maxl) continue;
                            }
                            if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
                            const float ur = P.u_right[idx];
                            if (P.fuse) {   // chi2 gate on the reprojection error (src/ORBmatcher.cc:913-936)
                                const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
                                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                                double lim = 5.99;
                                if (ur >= 0) { const float er = __fsub_rn(urq, ur); e2 = __fadd_rn(e2, __fmul_rn(er, er)); lim = 7.8; }
                                if ((double)__fmul_rn(e2, P.inv_sigma2[kp.octave]) > lim) continue;
                            } else {
                                if (prev[idx] <= (uint32_t)q) continue;              // held by an observed map point
                                if (ur > 0 && fabsf(__fsub_rn(urq, ur)) > r) continue;
                            }
                            const uint4* tp = reinterpret_cast<const uint4*>(P.desc + (size_t)idx * 32);
                            const uint4 ta = __ldg(tp), tb = __ldg(tp + 1);
                            const int d = __popc(qd[0] ^ ta.x) + __popc(qd[1] ^ ta.y) + __popc(qd[2] ^ ta.z) + __popc(qd[3] ^ ta.w) +
                                          __popc(qd[4] ^ tb.x) + __popc(qd[5] ^ tb.y) + __popc(qd[6] ^ tb.z) + __popc(qd[7] ^ tb.w);
                            const uint64_t key = ((uint64_t)d << 43) | ((uint64_t)k << 31) | ((uint64_t)(j - a) << 18) |
                                                 ((uint64_t)idx << 5) | (uint64_t)(kp.octave & 31);
                            key_insert(key, best, second);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint64_t ob = __shfl_xor_sync(0xFFFFFFFFu, best, o), os = __shfl_xor_sync(0xFFFFFFFFu, second, o);
                const uint64_t lo = min(best, ob), hi = max(best, ob);
                second = min(hi, min(second, os));
                best = lo;
            }
            if (lane == 0) {
                int choice = -1, bidx = -1, bdist = 256;
                if (best != ~0ull) {
                    bdist = (int)(best >> 43); bidx = (int)((best >> 5) & 0x1FFF);
                    if (bdist <= (P.fuse ? kThLow : kThHigh)) {
                        bool ok = true;
                        if (P.use_ratio) {
                            const int lvl = (int)(best & 31), lvl2 = second != ~0ull ? (int)(second & 31) : -1;
                            const int d2 = second != ~0ull ? (int)(second >> 43) : 256;
                            if (lvl == lvl2 && (float)bdist > __fmul_rn(P.nn_ratio, (float)d2)) ok = false;
                        }
                        if (ok) {
                            choice = bidx;
                            if ((fl & 2) && !P.fuse) atomicMin(&cur[bidx], (uint32_t)q + 1u);
                        }
                    }
                }
                P.q_choice[q] = choice;
                if (P.q_best_idx) P.q_best_idx[q] = bidx;
                if (P.q_best_dist) P.q_best_dist[q] = bdist;
            }
        }
        __syncthreads();
        int changed = 0;
        for (int i = tid; i < nk; i += kSearchThreads) changed |= (cur[i] != prev[i]);
        uint32_t* t = prev; prev = cur; cur = t;
        if (!__syncthreads_or(changed)) break;
    }

    // ---- write-back: mvpMapPoints[c] = the last query that took c; rotation histogram on the accepted queries
    int32_t* owner = reinterpret_cast<int32_t*>(cur);      // both closure arrays are free now
    uint32_t* cleared = prev;
    for (int i = tid; i < nk; i += kSearchThreads) { owner[i] = -1; cleared[i] = 0; }
    __syncthreads();
    int my_acc = 0;
    for (int q = tid; q < nq; q += kSearchThreads) {
        const int c = P.q_choice[q];
        if (c < 0) continue;
        ++my_acc;
        atomicMax(&owner[c], q);
        if (P.check_ori) {
            float rot = __fsub_rn(P.q_angle[q], P.kps[c].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHisto));
            if (bin == kHisto) bin = 0;
            atomicAdd(&s_hist[bin], 1);
        }
    }
    if (my_acc) atomicAdd(&s_acc, my_acc);
    __syncthreads();
    if (P.check_ori) {
        if (tid == 0) {   // ORBmatcher::ComputeThreeMaxima
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int i = 0; i < kHisto; ++i) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
            s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
        }
        __syncthreads();
        int my_rm = 0;
        for (int q = tid; q < nq; q += kSearchThreads) {
            const int c = P.q_choice[q];
            if (c < 0) continue;
            float rot = __fsub_rn(P.q_angle[q], P.kps[c].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHisto));
            if (bin == kHisto) bin = 0;
            if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { cleared[c] = 1; ++my_rm; }
        }
        if (my_rm) atomicAdd(&s_removed, my_rm);
        __syncthreads();
    }
    for (int i = tid; i < nk; i += kSearchThreads) P.kp_match[i] = cleared[i] ? -2 : owner[i];
    if (tid == 0) *P.n_matches = s_acc - s_removed;
}

// =========================================================================================
// Vocabulary-bucket searches (SearchByBoW, SearchForTriangulation).  One CTA per problem, one warp per query (a side-1
// feature in FeatureVector order), lanes stride the side-2 features of the same vocabulary node.
struct BowDev {
    int mode, n1, n2, n_buckets, nq;
    const adb_keypoint* kps1; const float* ur1; const uint8_t* d1; const uint8_t* fl1;
    const adb_keypoint* kps2; const float* ur2; const uint8_t* d2; const uint8_t* fl2;
    const int32_t* p2; const int32_t* i1; const int32_t* i2; const int32_t* q_bucket;
    float nn_ratio; int check_ori;
    float F12[9], ex, ey;
    const float* scale2; const float* sigma2_2;
    int32_t* choice;      // [nq] accepted partner of every query position (-1: none)
    int32_t* match21; int32_t* match12; int32_t* n_matches;
};

__device__ __forceinline__ int rot_bin(float a1, float a2) {   // src/ORBmatcher.cc:228-234
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHisto));
    return bin == kHisto ? 0 : bin;
}

__device__ __forceinline__ void three_maxima(const int* hist, int* keep) {   // ORBmatcher::ComputeThreeMaxima
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < kHisto; ++i) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
    keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
}

__global__ void __launch_bounds__(kSearchThreads) bow_search_kernel(const BowDev* __restrict__ probs) {
    extern __shared__ __align__(16) uint8_t search_smem[];
    const BowDev& P = probs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSearchThreads / 32;
    uint32_t* blk0 = reinterpret_cast<uint32_t*>(search_smem);
    uint32_t* blk1 = blk0 + P.n2;
    __shared__ int s_hist[kHisto];
    __shared__ int s_keep[3];
    __shared__ int s_acc, s_removed;
    if (tid < kHisto) s_hist[tid] = 0;
    if (tid == 0) { s_acc = 0; s_removed = 0; }
    for (int i = tid; i < P.n2; i += kSearchThreads) blk0[i] = kNoBlock;
    __syncthreads();
    uint32_t* prev = blk0;
    uint32_t* cur = blk1;
    const int max_rounds = P.mode == 0 ? P.nq + 1 : 1;
    for (int round = 0; round < max_rounds; ++round) {
        for (int i = tid; i < P.n2; i += kSearchThreads) cur[i] = kNoBlock;
        __syncthreads();
        for (int a = warp; a < P.nq; a += nwarps) {
            const int idx1 = P.i1[a];
            uint64_t best = ~0ull, second = ~0ull;
            if (P.fl1[idx1] & 1) {
                const int b = P.q_bucket[a];
                const int c0 = P.p2[b], c1 = P.p2[b + 1];
                uint32_t qd[8];
                {
                    const uint4* qp = reinterpret_cast<const uint4*>(P.d1 + (size_t)idx1 * 32);
                    const uint4 u = __ldg(qp), w = __ldg(qp + 1);
                    qd[0] = u.x; qd[1] = u.y; qd[2] = u.z; qd[3] = u.w; qd[4] = w.x; qd[5] = w.y; qd[6] = w.z; qd[7] = w.w;
                }
                float la = 0.f, lb = 0.f, lc = 0.f, den = 0.f;
                bool stereo1 = false;
                if (P.mode == 1) {   // epipolar line of key-point 1 in image 2 (CheckDistEpipolarLine, src/ORBmatcher.cc:131-157)
                    const adb_keypoint k1 = P.kps1[idx1];
                    la = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, P.F12[0]), __fmul_rn(k1.y, P.F12[3])), P.F12[6]);
                    lb = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, P.F12[1]), __fmul_rn(k1.y, P.F12[4])), P.F12[7]);
                    lc = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, P.F12[2]), __fmul_rn(k1.y, P.F12[5])), P.F12[8]);
                    den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
                    stereo1 = P.ur1[idx1] >= 0;
                }
                for (int c = c0 + lane; c < c1; c += 32) {
                    const int idx2 = P.i2[c];
                    if (P.mode == 0) { if (prev[idx2] <= (uint32_t)a) continue; }           // matched by an earlier query
                    else if (!(P.fl2[idx2] & 1)) continue;
                    const uint4* tp = reinterpret_cast<const uint4*>(P.d2 + (size_t)idx2 * 32);
                    const uint4 ta = __ldg(tp), tb = __ldg(tp + 1);
                    const int d = __popc(qd[0] ^ ta.x) + __popc(qd[1] ^ ta.y) + __popc(qd[2] ^ ta.z) + __popc(qd[3] ^ ta.w) +
                                  __popc(qd[4] ^ tb.x) + __popc(qd[5] ^ tb.y) + __popc(qd[6] ^ tb.z) + __popc(qd[7] ^ tb.w);
                    uint32_t pos = (uint32_t)(c - c0);
                    if (P.mode == 1) {
                        if (d > kThLow) continue;
                        const adb_keypoint k2 = P.kps2[idx2];
                        if (!stereo1 && !(P.ur2[idx2] >= 0)) {   // both monocular: not too close to the epipole (:744-750)
                            const float dx = __fsub_rn(P.ex, k2.x), dy = __fsub_rn(P.ey, k2.y);
                            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, P.scale2[k2.octave])) continue;
                        }
                        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, k2.x), __fmul_rn(lb, k2.y)), lc);
                        if (den == 0.f) continue;
                        const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                        if (!((double)dsqr < __dmul_rn(3.84, (double)P.sigma2_2[k2.octave]))) continue;
                        pos = 8191u - pos;                                                  // equal distance: the later one wins
                    }
                    key_insert(((uint64_t)d << 26) | ((uint64_t)pos << 13) | (uint64_t)idx2, best, second);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint64_t ob = __shfl_xor_sync(0xFFFFFFFFu, best, o), os = __shfl_xor_sync(0xFFFFFFFFu, second, o);
                const uint64_t lo = min(best, ob), hi = max(best, ob);
                second = min(hi, min(second, os));
                best = lo;
            }
            if (lane == 0) {
                int choice = -1;
                if (best != ~0ull) {
                    const int bd = (int)(best >> 26), bi = (int)(best & 0x1FFF);
                    if (P.mode == 1) choice = bi;                                           // gates already applied
                    else if (bd <= kThLow) {
                        const int d2 = second != ~0ull ? (int)(second >> 26) : 256;
                        if ((float)bd < __fmul_rn(P.nn_ratio, (float)d2)) { choice = bi; atomicMin(&cur[bi], (uint32_t)a + 1u); }
                    }
                }
                P.choice[a] = choice;
            }
        }
        __syncthreads();
        int changed = 0;
        for (int i = tid; i < P.n2; i += kSearchThreads) changed |= (cur[i] != prev[i]);
        uint32_t* t = prev; prev = cur; cur = t;
        if (!__syncthreads_or(changed)) break;
    }
    // ---- write-back + rotation consistency
    if (P.mode == 0) for (int i = tid; i < P.n2; i += kSearchThreads) P.match21[i] = -1;
    else for (int i = tid; i < P.n1; i += kSearchThreads) P.match12[i] = -1;
    __syncthreads();
    int my_acc = 0;
    for (int a = tid; a < P.nq; a += kSearchThreads) {
        const int c = P.choice[a];
        if (c < 0) continue;
        const int idx1 = P.i1[a];
        ++my_acc;
        if (P.mode == 0) P.match21[c] = idx1; else P.match12[idx1] = c;   // mode 0: one accepted query per key-point (closure)
        if (P.check_ori) atomicAdd(&s_hist[rot_bin(P.kps1[idx1].angle, P.kps2[c].angle)], 1);
    }
    if (my_acc) atomicAdd(&s_acc, my_acc);
    __syncthreads();
    if (P.check_ori) {
        if (tid == 0) three_maxima(s_hist, s_keep);
        __syncthreads();
        int my_rm = 0;
        for (int a = tid; a < P.nq; a += kSearchThreads) {
            const int c = P.choice[a];
            if (c < 0) continue;
            const int idx1 = P.i1[a];
            const int bin = rot_bin(P.kps1[idx1].angle, P.kps2[c].angle);
            if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) {
                if (P.mode == 0) P.match21[c] = -1; else P.match12[idx1] = -1;
                ++my_rm;
            }
        }
        if (my_rm) atomicAdd(&s_removed, my_rm);
        __syncthreads();
    }
    if (tid == 0) *P.n_matches = s_acc - s_removed;
}

static size_t search_smem_bytes(int nk) {
    return (size_t)(2 * kGridCells + 1) * 4 + (size_t)2 * nk * 4 + (size_t)((nk + 1) & ~1) * 2 + (size_t)nk * 2 + 16;
}

}  // namespace adb

using namespace adb;

namespace {

struct Packer {   // lays host arrays out in one staging block; device pointers are base + offset
    uint8_t* h = nullptr;
    uint8_t* d = nullptr;
    size_t off = 0;
    struct Job { uint8_t* dst; const void* src; size_t bytes; };
    std::vector<Job> jobs;
    size_t job_bytes = 0;
    size_t reserve(size_t bytes) { off = (off + 15) & ~(size_t)15; const size_t o = off; off += bytes; return o; }
    template <typename T>
    T* put(const T* src, size_t n) {   // second pass (h set): queues the copy into the pinned block; returns the device address
        const size_t o = reserve(n * sizeof(T));
        if (h && src && n) { jobs.push_back(Job{h + o, src, n * sizeof(T)}); job_bytes += n * sizeof(T); }
        return reinterpret_cast<T*>(d + o);
    }
    // The staging copy from the caller's (pageable) arrays is what a batched call costs on the host: 25 MB per 128 tracking frames.
    // Above 2 MB it is spread over a few threads (the copies are independent ranges of one pinned block).
    void run() {
        const int nt = job_bytes < (2u << 20) ? 1 : (int)std::min<size_t>(8, std::max<size_t>(1, std::thread::hardware_concurrency() / 2));
        if (nt <= 1) { for (const Job& j : jobs) memcpy(j.dst, j.src, j.bytes); jobs.clear(); job_bytes = 0; return; }
        std::vector<std::thread> th;
        const size_t per = (job_bytes + nt - 1) / nt;
        size_t first = 0;
        for (int t = 0; t < nt && first < jobs.size(); ++t) {
            size_t last = first, acc = 0;
            while (last < jobs.size() && (acc < per || t == nt - 1)) acc += jobs[last++].bytes;
            th.emplace_back([this, first, last] { for (size_t k = first; k < last; ++k) memcpy(jobs[k].dst, jobs[k].src, jobs[k].bytes); });
            first = last;
        }
        for (auto& x : th) x.join();
        jobs.clear(); job_bytes = 0;
    }
};

}  // namespace

extern "C" {

adb_status adb_search_last_ms(adb_matcher_t m, float* ms) {
    ADB_CHECK(m && ms, ADB_ERR_INVALID, "null argument");
    *ms = m->last_ms;
    return ADB_OK;
}

adb_status adb_search_by_projection(adb_matcher_t m, adb_proj_search* probs, int32_t n) {
    ADB_CHECK(m && (probs || n == 0) && n >= 0, ADB_ERR_INVALID, "null argument");
    if (n == 0) return ADB_OK;
    ADB_CUDA(cudaSetDevice(m->device));
    int max_nk = 0, max_nq = 0;
    for (int p = 0; p < n; ++p) {
        const adb_proj_search& s = probs[p];
        ADB_CHECK(s.n_kp >= 0 && s.n_kp <= ADB_SEARCH_MAX && s.n_q >= 0 && s.n_q <= ADB_SEARCH_MAX, ADB_ERR_INVALID,
                  "problem %d: sizes out of range (max %d)", p, ADB_SEARCH_MAX);
        ADB_CHECK(s.kp_match || s.n_kp == 0, ADB_ERR_INVALID, "problem %d: kp_match is NULL", p);
        ADB_CHECK(s.n_kp == 0 || (s.kps && s.u_right && s.desc), ADB_ERR_INVALID, "problem %d: frame arrays missing", p);
        ADB_CHECK(s.n_q == 0 || (s.q_flags && s.q_desc), ADB_ERR_INVALID, "problem %d: query arrays missing", p);
        ADB_CHECK(s.n_q == 0 || s.last_xw || s.mp_xw || (s.q_u && s.q_v && s.q_ur && s.q_radius && s.q_min_level && s.q_max_level),
                  ADB_ERR_INVALID, "problem %d: neither projected queries nor last-frame / map points given", p);
        ADB_CHECK(!(s.last_xw && s.mp_xw), ADB_ERR_INVALID, "problem %d: last-frame and map-point projection are exclusive", p);
        ADB_CHECK(!s.mp_xw || (s.mp_normal && s.mp_min_distance && s.mp_max_distance && s.ow && s.tcw_cur && s.scale_factors && s.n_levels > 0),
                  ADB_ERR_INVALID, "problem %d: map-point visibility inputs missing", p);
        ADB_CHECK(!s.last_xw || (s.last_octave && s.tcw_cur && s.tcw_last && s.scale_factors && s.n_levels > 0), ADB_ERR_INVALID,
                  "problem %d: last-frame projection inputs missing", p);
        ADB_CHECK(!s.check_orientation || s.q_angle || s.fuse, ADB_ERR_INVALID, "problem %d: q_angle missing", p);
        ADB_CHECK(!s.fuse || (s.mp_xw && s.inv_level_sigma2), ADB_ERR_INVALID, "problem %d: fuse needs the map-point inputs and inv_level_sigma2", p);
        // level numbers that index the per-level tables on the device (a bad one from the shim must not read past them)
        if (s.last_xw)
            for (int i = 0; i < s.n_q; ++i)
                ADB_CHECK(s.last_octave[i] >= 0 && s.last_octave[i] < s.n_levels, ADB_ERR_INVALID, "problem %d: last_octave[%d] = %d outside [0, %d)", p, i,
                          s.last_octave[i], s.n_levels);
        if (s.fuse)
            for (int i = 0; i < s.n_kp; ++i)
                ADB_CHECK(s.kps[i].octave >= 0 && s.kps[i].octave < s.n_levels, ADB_ERR_INVALID, "problem %d: key-point %d has octave %d outside [0, %d)", p, i,
                          s.kps[i].octave, s.n_levels);
        max_nk = std::max(max_nk, s.n_kp); max_nq = std::max(max_nq, s.n_q);
    }
    // two passes over the same layout code: size, then copy
    std::vector<SearchDev> dev(n);
    size_t out_begin = 0, total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        Packer pk;
        if (pass == 1) { pk.h = m->h_scratch; pk.d = m->d_scratch; }
        pk.reserve((size_t)n * sizeof(SearchDev));
        for (int p = 0; p < n; ++p) {
            const adb_proj_search& s = probs[p];
            SearchDev& D = dev[p];
            memset(&D, 0, sizeof(D));
            D.n_kp = s.n_kp; D.n_q = s.n_q;
            D.kps = pk.put(s.kps, s.n_kp); D.u_right = pk.put(s.u_right, s.n_kp); D.desc = pk.put(s.desc, (size_t)s.n_kp * 32);
            D.taken = s.taken ? pk.put(s.taken, s.n_kp) : nullptr;
            D.min_x = s.min_x; D.min_y = s.min_y; D.max_x = s.max_x; D.max_y = s.max_y; D.inv_w = s.grid_inv_w; D.inv_h = s.grid_inv_h;
            const bool proj = s.last_xw != nullptr || s.mp_xw != nullptr;
            D.q_u = pk.put(proj ? nullptr : s.q_u, s.n_q); D.q_v = pk.put(proj ? nullptr : s.q_v, s.n_q);
            D.q_ur = pk.put(proj ? nullptr : s.q_ur, s.n_q); D.q_radius = pk.put(proj ? nullptr : s.q_radius, s.n_q);
            D.q_minl = pk.put(proj ? nullptr : s.q_min_level, s.n_q); D.q_maxl = pk.put(proj ? nullptr : s.q_max_level, s.n_q);
            D.q_flags = pk.put(proj ? nullptr : s.q_flags, s.n_q);
            D.q_desc = pk.put(s.q_desc, (size_t)s.n_q * 32);
            D.q_angle = s.check_orientation ? pk.put(s.q_angle, s.n_q) : nullptr;
            D.use_ratio = s.use_ratio; D.nn_ratio = s.nn_ratio; D.check_ori = s.check_orientation;
            if (s.mp_xw) {
                D.mp_xw = pk.put(s.mp_xw, (size_t)s.n_q * 3); D.mp_normal = pk.put(s.mp_normal, (size_t)s.n_q * 3);
                D.mp_min_distance = pk.put(s.mp_min_distance, s.n_q); D.mp_max_distance = pk.put(s.mp_max_distance, s.n_q);
                D.mp_flags = pk.put(s.q_flags, s.n_q);
                D.scale_factors = pk.put(s.scale_factors, s.n_levels);
                for (int k = 0; k < 12; ++k) D.Tcw[k] = s.tcw_cur[k];
                for (int k = 0; k < 3; ++k) D.ow[k] = s.ow[k];
                D.fx = s.fx; D.fy = s.fy; D.cx = s.cx; D.cy = s.cy; D.mbf = s.mbf; D.th = s.th;
                D.view_cos_limit = s.view_cos_limit; D.log_scale_factor = s.log_scale_factor; D.n_levels = s.n_levels;
                D.fuse = s.fuse ? 1 : 0;
                if (s.fuse) { D.inv_sigma2 = pk.put(s.inv_level_sigma2, s.n_levels); D.use_ratio = 0; D.check_ori = 0; }
            } else if (proj) {
                D.last_xw = pk.put(s.last_xw, (size_t)s.n_q * 3); D.last_octave = pk.put(s.last_octave, s.n_q);
                D.last_flags = pk.put(s.q_flags, s.n_q);
                D.scale_factors = pk.put(s.scale_factors, s.n_levels);
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) D.Rcw[3 * r + c] = s.tcw_cur[4 * r + c];
                    D.tcw[r] = s.tcw_cur[4 * r + 3];
                }
                D.fx = s.fx; D.fy = s.fy; D.cx = s.cx; D.cy = s.cy; D.mbf = s.mbf; D.th = s.th;
                // twc = -Rcw^T tcw ; tlc = Rlw twc + tlw (cv::gemm: double accumulation, one rounding) -> bForward / bBackward
                float twc[3], tlc[3];
                for (int i = 0; i < 3; ++i) {
                    double a = 0;
                    for (int k = 0; k < 3; ++k) a += (-(double)s.tcw_cur[4 * k + i]) * (double)s.tcw_cur[4 * k + 3];
                    twc[i] = (float)a;
                }
                for (int i = 0; i < 3; ++i) {
                    double a = 0;
                    for (int k = 0; k < 3; ++k) a += (double)s.tcw_last[4 * i + k] * (double)twc[k];
                    tlc[i] = (float)(a + (double)s.tcw_last[4 * i + 3]);
                }
                D.forward = (tlc[2] > s.mb && !s.mono) ? 1 : 0;
                D.backward = (-tlc[2] > s.mb && !s.mono) ? 1 : 0;
            }
        }
        out_begin = pk.reserve(0);
        for (int p = 0; p < n; ++p) {   // results: contiguous tail of the block, copied back in one piece
            const adb_proj_search& s = probs[p];
            SearchDev& D = dev[p];
            D.kp_match = pk.put((const int32_t*)nullptr, s.n_kp);
            D.q_best_idx = pk.put((const int32_t*)nullptr, s.n_q);
            D.q_best_dist = pk.put((const int32_t*)nullptr, s.n_q);
            D.q_choice = pk.put((const int32_t*)nullptr, s.n_q);
            D.n_matches = pk.put((const int32_t*)nullptr, 1);
            D.q_track = pk.put((const float*)nullptr, (size_t)s.n_q * 4);
            D.q_level = pk.put((const int32_t*)nullptr, s.n_q);
        }
        if (pass == 1) pk.run();
        total = pk.reserve(0);
        if (pass == 0 && total > m->scratch_bytes) {
            cudaFree(m->d_scratch); m->d_scratch = nullptr;
            if (m->h_scratch) { cudaFreeHost(m->h_scratch); m->h_scratch = nullptr; }
            m->scratch_bytes = 0;
            const size_t want = total + total / 4;
            ADB_CUDA(cudaMalloc(&m->d_scratch, want));
            ADB_CUDA(cudaMemsetAsync(m->d_scratch, 0, want, m->stream));   // alignment gaps and device-only work arrays sit inside the downloaded range
            ADB_CUDA(cudaMallocHost(&m->h_scratch, want));
            m->scratch_bytes = want;
        }
    }
    memcpy(m->h_scratch, dev.data(), (size_t)n * sizeof(SearchDev));
    if (!m->ev[0]) { ADB_CUDA(cudaEventCreate(&m->ev[0])); ADB_CUDA(cudaEventCreate(&m->ev[1])); }
    const size_t smem = search_smem_bytes(max_nk);
    static bool attr_set = false;
    if (!attr_set) {
        ADB_CUDA(cudaFuncSetAttribute(proj_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_smem_bytes(ADB_SEARCH_MAX)));
        attr_set = true;
    }
    ADB_CUDA(cudaMemcpyAsync(m->d_scratch, m->h_scratch, out_begin, cudaMemcpyHostToDevice, m->stream));
    const SearchDev* dprobs = reinterpret_cast<const SearchDev*>(m->d_scratch);
    ADB_CUDA(cudaEventRecord(m->ev[0], m->stream));
    bool any_proj = false;
    for (int p = 0; p < n; ++p) any_proj |= probs[p].last_xw != nullptr;
    if (any_proj && max_nq > 0) {
        dim3 grid((max_nq + 255) / 256, n);
        project_last_kernel<<<grid, 256, 0, m->stream>>>(dprobs);
    }
    bool any_mp = false;
    for (int p = 0; p < n; ++p) any_mp |= probs[p].mp_xw != nullptr;
    if (any_mp && max_nq > 0) {
        dim3 grid((max_nq + 255) / 256, n);
        frustum_kernel<<<grid, 256, 0, m->stream>>>(dprobs);
    }
    proj_search_kernel<<<n, kSearchThreads, smem, m->stream>>>(dprobs);
    ADB_CUDA(cudaGetLastError());
    ADB_CUDA(cudaEventRecord(m->ev[1], m->stream));
    ADB_CUDA(cudaMemcpyAsync(m->h_scratch + out_begin, m->d_scratch + out_begin, total - out_begin, cudaMemcpyDeviceToHost, m->stream));
    ADB_CUDA(cudaStreamSynchronize(m->stream));
    ADB_CUDA(cudaEventElapsedTime(&m->last_ms, m->ev[0], m->ev[1]));
    for (int p = 0; p < n; ++p) {
        adb_proj_search& s = probs[p];
        const SearchDev& D = dev[p];
        auto host = [&](const void* dptr) { return m->h_scratch + ((const uint8_t*)dptr - m->d_scratch); };
        if (s.n_kp) memcpy(s.kp_match, host(D.kp_match), (size_t)s.n_kp * 4);
        if (s.q_best_idx && s.n_q) memcpy(s.q_best_idx, host(D.q_best_idx), (size_t)s.n_q * 4);
        if (s.q_best_dist && s.n_q) memcpy(s.q_best_dist, host(D.q_best_dist), (size_t)s.n_q * 4);
        memcpy(&s.n_matches, host(D.n_matches), 4);
        if (s.mp_xw && s.q_track && s.n_q) memcpy(s.q_track, host(D.q_track), (size_t)s.n_q * 16);
        if (s.mp_xw && s.q_level && s.n_q) memcpy(s.q_level, host(D.q_level), (size_t)s.n_q * 4);
    }
    return ADB_OK;
}

adb_status adb_search_by_bow(adb_matcher_t m, adb_bow_search* probs, int32_t n) {
    ADB_CHECK(m && (probs || n == 0) && n >= 0, ADB_ERR_INVALID, "null argument");
    if (n == 0) return ADB_OK;
    ADB_CUDA(cudaSetDevice(m->device));
    int max_n2 = 0;
    std::vector<std::vector<int32_t>> qb(n);
    for (int p = 0; p < n; ++p) {
        const adb_bow_search& s = probs[p];
        ADB_CHECK((s.mode == 0 || s.mode == 1) && s.n1 >= 0 && s.n1 <= ADB_SEARCH_MAX && s.n2 >= 0 && s.n2 <= ADB_SEARCH_MAX && s.n_buckets >= 0,
                  ADB_ERR_INVALID, "problem %d: bad mode or sizes (max %d key-points)", p, ADB_SEARCH_MAX);
        ADB_CHECK(s.n1 == 0 || (s.kps1 && s.desc1 && s.flags1), ADB_ERR_INVALID, "problem %d: side-1 arrays missing", p);
        ADB_CHECK(s.n2 == 0 || (s.kps2 && s.desc2), ADB_ERR_INVALID, "problem %d: side-2 arrays missing", p);
        ADB_CHECK(s.n_buckets == 0 || (s.b_ptr1 && s.b_idx1 && s.b_ptr2 && s.b_idx2), ADB_ERR_INVALID, "problem %d: bucket lists missing", p);
        ADB_CHECK(s.mode == 0 ? s.match21 != nullptr || s.n2 == 0 : s.match12 != nullptr || s.n1 == 0, ADB_ERR_INVALID, "problem %d: result array missing", p);
        ADB_CHECK(s.mode == 0 || (s.f12 && s.scale_factors2 && s.level_sigma2_2 && s.n_levels > 0 && (s.n1 == 0 || s.u_right1) &&
                                  (s.n2 == 0 || (s.u_right2 && s.flags2))), ADB_ERR_INVALID, "problem %d: triangulation inputs missing", p);
        const int nq = s.n_buckets ? s.b_ptr1[s.n_buckets] : 0, nt = s.n_buckets ? s.b_ptr2[s.n_buckets] : 0;
        ADB_CHECK(nq >= 0 && nt >= 0 && nq <= ADB_SEARCH_MAX && nt <= ADB_SEARCH_MAX, ADB_ERR_INVALID, "problem %d: bucket lists longer than the frames", p);
        ADB_CHECK(s.n_buckets == 0 || (s.b_ptr1[0] == 0 && s.b_ptr2[0] == 0), ADB_ERR_INVALID, "problem %d: bucket pointers must start at 0", p);
        for (int b = 0; b < s.n_buckets; ++b)
            ADB_CHECK(s.b_ptr1[b + 1] >= s.b_ptr1[b] && s.b_ptr2[b + 1] >= s.b_ptr2[b], ADB_ERR_INVALID, "problem %d: bucket pointers are not monotone at %d", p, b);
        for (int a = 0; a < nq; ++a)
            ADB_CHECK(s.b_idx1[a] >= 0 && s.b_idx1[a] < s.n1, ADB_ERR_INVALID, "problem %d: b_idx1[%d] = %d outside [0, %d)", p, a, s.b_idx1[a], s.n1);
        for (int a = 0; a < nt; ++a)
            ADB_CHECK(s.b_idx2[a] >= 0 && s.b_idx2[a] < s.n2, ADB_ERR_INVALID, "problem %d: b_idx2[%d] = %d outside [0, %d)", p, a, s.b_idx2[a], s.n2);
        if (s.mode == 1)
            for (int i = 0; i < s.n2; ++i)
                ADB_CHECK(s.kps2[i].octave >= 0 && s.kps2[i].octave < s.n_levels, ADB_ERR_INVALID, "problem %d: key-point %d of side 2 has octave %d outside [0, %d)",
                          p, i, s.kps2[i].octave, s.n_levels);
        for (int b = 0; b < s.n_buckets; ++b) {
            ADB_CHECK(s.b_ptr2[b + 1] - s.b_ptr2[b] <= 8191, ADB_ERR_INVALID, "problem %d: vocabulary node with more than 8191 features", p);
            for (int a = s.b_ptr1[b]; a < s.b_ptr1[b + 1]; ++a) qb[p].push_back(b);
        }
        max_n2 = std::max(max_n2, s.n2);
    }
    std::vector<BowDev> dev(n);
    size_t out_begin = 0, total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        Packer pk;
        if (pass == 1) { pk.h = m->h_scratch; pk.d = m->d_scratch; }
        pk.reserve((size_t)n * sizeof(BowDev));
        for (int p = 0; p < n; ++p) {
            const adb_bow_search& s = probs[p];
            BowDev& D = dev[p];
            memset(&D, 0, sizeof(D));
            const int nq = (int)qb[p].size(), nt = s.n_buckets ? s.b_ptr2[s.n_buckets] : 0;
            D.mode = s.mode; D.n1 = s.n1; D.n2 = s.n2; D.n_buckets = s.n_buckets; D.nq = nq;
            D.kps1 = pk.put(s.kps1, s.n1); D.d1 = pk.put(s.desc1, (size_t)s.n1 * 32); D.fl1 = pk.put(s.flags1, s.n1);
            D.kps2 = pk.put(s.kps2, s.n2); D.d2 = pk.put(s.desc2, (size_t)s.n2 * 32);
            D.p2 = pk.put(s.b_ptr2, s.n_buckets + 1); D.i1 = pk.put(s.b_idx1, nq); D.i2 = pk.put(s.b_idx2, nt);
            D.q_bucket = pk.put(qb[p].data(), nq);
            D.nn_ratio = s.nn_ratio; D.check_ori = s.check_orientation;
            if (s.mode == 1) {
                D.ur1 = pk.put(s.u_right1, s.n1); D.ur2 = pk.put(s.u_right2, s.n2); D.fl2 = pk.put(s.flags2, s.n2);
                D.scale2 = pk.put(s.scale_factors2, s.n_levels); D.sigma2_2 = pk.put(s.level_sigma2_2, s.n_levels);
                for (int k = 0; k < 9; ++k) D.F12[k] = s.f12[k];
                D.ex = s.ex; D.ey = s.ey;
            }
        }
        out_begin = pk.reserve(0);
        for (int p = 0; p < n; ++p) {
            const adb_bow_search& s = probs[p];
            BowDev& D = dev[p];
            D.match21 = pk.put((const int32_t*)nullptr, s.n2);
            D.match12 = pk.put((const int32_t*)nullptr, s.n1);
            D.choice = pk.put((const int32_t*)nullptr, qb[p].size());
            D.n_matches = pk.put((const int32_t*)nullptr, 1);
        }
        if (pass == 1) pk.run();
        total = pk.reserve(0);
        if (pass == 0 && total > m->scratch_bytes) {
            cudaFree(m->d_scratch); m->d_scratch = nullptr;
            if (m->h_scratch) { cudaFreeHost(m->h_scratch); m->h_scratch = nullptr; }
            m->scratch_bytes = 0;
            const size_t want = total + total / 4;
            ADB_CUDA(cudaMalloc(&m->d_scratch, want));
            ADB_CUDA(cudaMemsetAsync(m->d_scratch, 0, want, m->stream));   // alignment gaps and device-only work arrays sit inside the downloaded range
            ADB_CUDA(cudaMallocHost(&m->h_scratch, want));
            m->scratch_bytes = want;
        }
    }
    memcpy(m->h_scratch, dev.data(), (size_t)n * sizeof(BowDev));
    if (!m->ev[0]) { ADB_CUDA(cudaEventCreate(&m->ev[0])); ADB_CUDA(cudaEventCreate(&m->ev[1])); }
    static bool attr_set = false;
    if (!attr_set) {
        ADB_CUDA(cudaFuncSetAttribute(bow_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * ADB_SEARCH_MAX * 4 + 16));
        attr_set = true;
    }
    ADB_CUDA(cudaMemcpyAsync(m->d_scratch, m->h_scratch, out_begin, cudaMemcpyHostToDevice, m->stream));
    ADB_CUDA(cudaEventRecord(m->ev[0], m->stream));
    bow_search_kernel<<<n, kSearchThreads, (size_t)2 * max_n2 * 4 + 16, m->stream>>>(reinterpret_cast<const BowDev*>(m->d_scratch));
    ADB_CUDA(cudaGetLastError());
    ADB_CUDA(cudaEventRecord(m->ev[1], m->stream));
    ADB_CUDA(cudaMemcpyAsync(m->h_scratch + out_begin, m->d_scratch + out_begin, total - out_begin, cudaMemcpyDeviceToHost, m->stream));
    ADB_CUDA(cudaStreamSynchronize(m->stream));
    ADB_CUDA(cudaEventElapsedTime(&m->last_ms, m->ev[0], m->ev[1]));
    for (int p = 0; p < n; ++p) {
        adb_bow_search& s = probs[p];
        const BowDev& D = dev[p];
        auto host = [&](const void* dptr) { return m->h_scratch + ((const uint8_t*)dptr - m->d_scratch); };
        if (s.mode == 0 && s.n2) memcpy(s.match21, host(D.match21), (size_t)s.n2 * 4);
        if (s.mode == 1 && s.n1) memcpy(s.match12, host(D.match12), (size_t)s.n1 * 4);
        memcpy(&s.n_matches, host(D.n_matches), 4);
    }
    return ADB_OK;
}

}  // extern "C"
