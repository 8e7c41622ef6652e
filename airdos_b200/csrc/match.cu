// match.cu -- Hamming matching kernels behind the C-ABI.
//
//   hamming_best2_kernel   best / second-best scan of ORBmatcher::Search*   src/ORBmatcher.cc:85-114 (and 10 more call sites)
//   stereo_match_kernel    Frame::ComputeStereoMatches, Hamming + SAD stage src/Frame.cc:858-986
//   stereo_median_kernel   median-distance cut                              src/Frame.cc:989-1002
//
// The 256-bit distance is XOR + POPC on 8 words per pair (the POPC pipe is the limiter; 1-bit data
// has no use for DP4A).  The reference's "first candidate in list order wins ties" rule makes
// every scan an arg-min over (distance, list position), which is associative, so lanes scan
// disjoint candidate subsets and merge with shuffles.
#include <algorithm>
#include <climits>
#include <cmath>

#include "orb.cuh"
#include "matcher.cuh"

namespace adb {

__device__ __forceinline__ int hamming8(const uint32_t (&q)[8], const uint4* __restrict__ t) {
    const uint4 a = __ldg(t), b = __ldg(t + 1);
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) + __popc(q[4] ^ b.x) +
           __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

constexpr int kMatchWarps = 8;

// One warp per query.  cand_off == nullptr: candidates are all targets in order.
__global__ void __launch_bounds__(kMatchWarps * 32) hamming_best2_kernel(const uint8_t* __restrict__ Q, int nq,
                                                                        const uint8_t* __restrict__ T, int nt,
                                                                        const int32_t* __restrict__ cand_off,
                                                                        const int32_t* __restrict__ cand_idx,
                                                                        int32_t* __restrict__ best_idx,
                                                                        int32_t* __restrict__ best_d,
                                                                        int32_t* __restrict__ second_d) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * kMatchWarps + (threadIdx.x >> 5);
    if (q >= nq) return;
    uint32_t qd[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(Q + (size_t)q * 32);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
    }
    const int lo = cand_off ? cand_off[q] : 0, hi = cand_off ? cand_off[q + 1] : nt;
    int best = 256, second = 256, pos = INT_MAX;
    for (int c = lo + lane; c < hi; c += 32) {
        const int t = cand_off ? __ldg(&cand_idx[c]) : c;
        const int d = hamming8(qd, reinterpret_cast<const uint4*>(T + (size_t)t * 32));
        if (d < best) { second = best; best = d; pos = c; }
        else if (d < second) second = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xFFFFFFFFu, best, o), os = __shfl_xor_sync(0xFFFFFFFFu, second, o);
        const int op = __shfl_xor_sync(0xFFFFFFFFu, pos, o);
        const int nsec = min(max(best, ob), min(second, os));
        if (ob < best || (ob == best && op < pos)) { best = ob; pos = op; }
        second = nsec;
    }
    if (lane == 0) {
        best_idx[q] = pos == INT_MAX ? -1 : (cand_off ? cand_idx[pos] : pos);
        best_d[q] = best;
        second_d[q] = second;
    }
}

// -----------------------------------------------------------------------------------------
struct StereoLevel {
    const uint8_t* L;   // left pyramid level (frame 0)
    const uint8_t* R;
    int w, pitchL, pitchR;
    unsigned fstrideL, fstrideR;
    float scale, inv_scale;
};
struct StereoLevels {
    StereoLevel l[kMaxLevels];
};

constexpr int kStereoWarps = 8;
constexpr int kStereoPerWarp = 8;   // left key-points per warp: amortises the per-CTA band table of the right key-points
constexpr int TH_HIGH = 100, TH_LOW = 50;   // src/ORBmatcher.cc:37-38

// Row buckets of the right key-points (src/Frame.cc:846-856): CSR over image rows, one CTA per frame.  The order inside a
// bucket is arbitrary: the consumer takes the arg-min over (distance, index), which is what "first candidate wins" means
// for the reference's ascending-index buckets.
constexpr int kBucketThreads = 512;
__global__ void __launch_bounds__(kBucketThreads) stereo_bucket_kernel(const __grid_constant__ StereoLevels lv, int n_rows,
                                                                      const adb_keypoint* __restrict__ kpsR, const int32_t* __restrict__ cntR,
                                                                      int cap, int maxband, int32_t* __restrict__ row_ptr,
                                                                      uint16_t* __restrict__ row_items, float2* __restrict__ rinfo, int f0) {
    extern __shared__ int bk_smem[];     // [n_rows + 1] counts -> offsets -> cursors
    __shared__ int s_run;
    const int f = blockIdx.x + f0, tid = threadIdx.x, lane = tid & 31;
    const int nR = min(cntR[f], cap);
    const adb_keypoint* kR = kpsR + (size_t)f * cap;
    for (int i = tid; i <= n_rows; i += kBucketThreads) bk_smem[i] = 0;
    __syncthreads();
    for (int i = tid; i < nR; i += kBucketThreads) {
        const adb_keypoint k = kR[i];
        const float r = __fmul_rn(2.0f, lv.l[k.octave].scale);
        const int maxr = min((int)ceilf(__fadd_rn(k.y, r)), n_rows - 1);      // clamped: DESIGN.md convention D.9
        const int minr = max((int)floorf(__fsub_rn(k.y, r)), 0);
        for (int y = minr; y <= maxr; ++y) atomicAdd(&bk_smem[y], 1);
        rinfo[(size_t)f * cap + i] = make_float2(k.x, __int_as_float(k.octave));
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan over the rows
        int run = 0;
        for (int y0 = 0; y0 <= n_rows; y0 += 32) {
            const int y = y0 + lane;
            const int v = y <= n_rows ? bk_smem[y] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
            if (y <= n_rows) bk_smem[y] = run + inc - v;
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) s_run = run;
    }
    __syncthreads();
    int32_t* rp = row_ptr + (size_t)f * (n_rows + 1);
    for (int i = tid; i <= n_rows; i += kBucketThreads) rp[i] = bk_smem[i];
    __syncthreads();
    uint16_t* items = row_items + (size_t)f * cap * maxband;
    for (int i = tid; i < nR; i += kBucketThreads) {
        const adb_keypoint k = kR[i];
        const float r = __fmul_rn(2.0f, lv.l[k.octave].scale);
        const int maxr = min((int)ceilf(__fadd_rn(k.y, r)), n_rows - 1);
        const int minr = max((int)floorf(__fsub_rn(k.y, r)), 0);
        for (int y = minr; y <= maxr; ++y) {
            const int pos = atomicAdd(&bk_smem[y], 1);
            if (pos < cap * maxband) items[pos] = (uint16_t)i;
        }
    }
}

__global__ void __launch_bounds__(kStereoWarps * 32) stereo_match_kernel(
    const __grid_constant__ StereoLevels lv, int n_rows, const adb_keypoint* __restrict__ kpsL,
    const uint8_t* __restrict__ descL, const int32_t* __restrict__ cntL, const uint8_t* __restrict__ descR, int cap, int maxband,
    const int32_t* __restrict__ row_ptr, const uint16_t* __restrict__ row_items, const float2* __restrict__ rinfo, float mbf, float maxD,
    float* __restrict__ uRight, float* __restrict__ depth, int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist,
    int32_t* __restrict__ sad_out, int f0) {
    const int f = blockIdx.y + f0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nL = min(cntL[f], cap);
    const int32_t* rp = row_ptr + (size_t)f * (n_rows + 1);
    const uint16_t* items = row_items + (size_t)f * cap * maxband;
    const float2* ri = rinfo + (size_t)f * cap;
    for (int sub = 0; sub < kStereoPerWarp; ++sub) {
    const int iL = (blockIdx.x * kStereoWarps + warp) * kStereoPerWarp + sub;
    if (iL >= cap) return;
    const size_t o = (size_t)f * cap + iL;
    if (iL >= nL) {
        if (lane == 0) { uRight[o] = -1.f; depth[o] = -1.f; best_idx[o] = -1; best_dist[o] = TH_HIGH; sad_out[o] = -1; }
        continue;
    }
    const adb_keypoint kp = kpsL[o];
    const int levelL = kp.octave, vL = (int)kp.y;
    const float uL = kp.x, minU = __fsub_rn(uL, maxD), maxU = uL;
    uint32_t qd[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(descL + o * 32);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        qd[0] = a.x; qd[1] = a.y; qd[2] = a.z; qd[3] = a.w; qd[4] = b.x; qd[5] = b.y; qd[6] = b.z; qd[7] = b.w;
    }
    int best = TH_HIGH, idx = INT_MAX;
    if (!(maxU < 0)) {
        const uint8_t* dR = descR + (size_t)f * cap * 32;
        const int c_end = rp[vL + 1];
        for (int c = rp[vL] + lane; c < c_end; c += 32) {
            const int iR = items[c];
            const float2 info = ri[iR];
            const int oct = __float_as_int(info.y);
            if (oct < levelL - 1 || oct > levelL + 1) continue;
            const float uR = info.x;
            if (uR >= minU && uR <= maxU) {
                const int d = hamming8(qd, reinterpret_cast<const uint4*>(dR + (size_t)iR * 32));
                if (d < best || (d == best && idx != INT_MAX && iR < idx)) { best = d; idx = iR; }   // buckets are unordered
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const int ob = __shfl_xor_sync(0xFFFFFFFFu, best, s), oi = __shfl_xor_sync(0xFFFFFFFFu, idx, s);
        if (ob < best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    float ur = -1.f, dp = -1.f;
    int sad = -1;
    if (idx != INT_MAX && best < (TH_HIGH + TH_LOW) / 2) {
        // sub-pixel refinement by 11x11 SAD on the key-point's pyramid level: src/Frame.cc:915-986
        const StereoLevel& S = lv.l[levelL];
        const float uR0 = ri[idx].x;
        const float sf = S.inv_scale;
        const float scaleduL = roundf(__fmul_rn(kp.x, sf)), scaledvL = roundf(__fmul_rn(kp.y, sf));
        const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int w = 5, Ls = 5;
        const float iniu = scaleduR0 + Ls - w, endu = scaleduR0 + Ls + w + 1;
        if (!(iniu < 0 || endu >= (float)S.w)) {
            const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR0 = (int)scaleduR0;
            const uint8_t* IL = S.L + (size_t)f * S.fstrideL;
            const uint8_t* IR = S.R + (size_t)f * S.fstrideR;
            const int cL = IL[(size_t)cy * S.pitchL + cxL];
            // lane owns window pixels p = lane + 32 k (121 = 11 x 11): its left value once, and ONE pointer per pixel into the right image
            // at shift 0 -- the 11 shifts are then immediate byte offsets of the same pointer
            int pl[4];
            const uint8_t* pr[4];
            const uint8_t* rc = IR + (size_t)cy * S.pitchR + cxR0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = lane + 32 * k;
                const int py = p / 11 - w, px = p % 11 - w;
                const bool in = p < 121;
                pl[k] = in ? (int)IL[(ptrdiff_t)(cy + py) * S.pitchL + cxL + px] - cL : 0;
                pr[k] = in ? rc + (ptrdiff_t)py * S.pitchR + px : rc;   // outside the window: the centre itself, |0 - (cR - cR)| = 0
            }
            // per-lane partial SADs of the 11 shifts, two per register (a lane's share is <= 4 x 510, the total <= 121 x 510 < 2^16)
            uint32_t pk[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int inc = -5; inc <= 5; ++inc) {
                const int cR = rc[inc];
                int acc = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc += abs(pl[k] - ((int)pr[k][inc] - cR));
                pk[(inc + 5) >> 1] += (uint32_t)acc << (((inc + 5) & 1) * 16);
            }
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) pk[j] += __shfl_xor_sync(0xFFFFFFFFu, pk[j], s);
            int bestSad = INT_MAX, bestinc = 0, mine = 0;        // lane l < 11 also keeps the SAD of shift l - 5 for the parabola
#pragma unroll
            for (int inc = -5; inc <= 5; ++inc) {
                const int acc = (int)((pk[(inc + 5) >> 1] >> (((inc + 5) & 1) * 16)) & 0xFFFFu);
                if (acc < bestSad) { bestSad = acc; bestinc = inc; }
                if (lane == inc + 5) mine = acc;
            }
            if (bestinc != -Ls && bestinc != Ls) {
                const float d1 = (float)__shfl_sync(0xFFFFFFFFu, mine, bestinc + 4), d2 = (float)bestSad,
                            d3 = (float)__shfl_sync(0xFFFFFFFFu, mine, bestinc + 6);
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(S.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= 0.f && disparity < maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        dp = __fdiv_rn(mbf, disparity);
                        ur = bestuR;
                        sad = bestSad;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        uRight[o] = ur; depth[o] = dp; sad_out[o] = sad;
        best_idx[o] = idx == INT_MAX ? -1 : idx;
        best_dist[o] = best;
    }
    }
}

// Median cut: thDist = 1.5 * 1.4 * median(SAD of the matched); matches at or above it are dropped.
// One CTA per frame; k-th smallest by a two-pass 256-bin radix select (SAD <= 121 * 510 < 2^16).
__global__ void __launch_bounds__(256) stereo_median_kernel(const int32_t* __restrict__ cntL, int cap,
                                                            const int32_t* __restrict__ sad, float* __restrict__ uRight,
                                                            float* __restrict__ depth, int f0) {
    __shared__ int hist[256];
    __shared__ int s_n, s_bin, s_rank;
    const int f = blockIdx.x + f0, tid = threadIdx.x;
    const int nL = min(cntL[f], cap);
    const int32_t* s = sad + (size_t)f * cap;
    hist[tid] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nL; i += 256) {
        const int v = s[i];
        if (v >= 0) { atomicAdd(&hist[(v >> 8) & 0xFF], 1); ++mine; }
    }
    atomicAdd(&s_n, mine);
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;   // src/Frame.cc:989-990 dereferences an empty vector here; convention D.8: no-op
    if (tid == 0) {
        int k = n / 2, b = 0;
        while (k >= hist[b]) { k -= hist[b]; ++b; }
        s_bin = b; s_rank = k;
    }
    __syncthreads();
    const int hb = s_bin;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nL; i += 256) {
        const int v = s[i];
        if (v >= 0 && ((v >> 8) & 0xFF) == hb) atomicAdd(&hist[v & 0xFF], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int k = s_rank, b = 0;
        while (k >= hist[b]) { k -= hist[b]; ++b; }
        s_bin = (hb << 8) | b;
    }
    __syncthreads();
    const float median = (float)s_bin;
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = tid; i < nL; i += 256) {
        const int v = s[i];
        if (v >= 0 && !((float)v < thDist)) {
            uRight[(size_t)f * cap + i] = -1.f;
            depth[(size_t)f * cap + i] = -1.f;
        }
    }
}


// -----------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors, one CTA (4 warps) per map point: N x N Hamming matrix in shared memory,
// per-row median by a counting binary search over the 257 possible distances, arg-min over (median, row).
constexpr int kDistinctThreads = 128;
__global__ void __launch_bounds__(kDistinctThreads) distinctive_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ point_ptr,
                                                                      int32_t* __restrict__ best_idx, uint8_t* __restrict__ best_desc) {
    extern __shared__ uint16_t dmat[];   // [N][N]
    __shared__ int s_best;               // median << 8 | row, reduced with atomicMin
    const int p = blockIdx.x, tid = threadIdx.x;
    const int a = point_ptr[p], N = point_ptr[p + 1] - a;
    if (N <= 0) {   // a point nobody observes (src/MapPoint.cc:259 returns before touching mDescriptor): index -1, the caller keeps its descriptor; the slot is zeroed
        if (tid == 0) best_idx[p] = -1;
        if (best_desc && tid < 32) best_desc[(size_t)p * 32 + tid] = 0;
        return;
    }
    if (tid == 0) s_best = INT_MAX;
    for (int e = tid; e < N * N; e += kDistinctThreads) {
        const int i = e / N, j = e - i * N;
        if (j < i) continue;
        int d = 0;
        if (j > i) {
            uint32_t q[8];
            const uint4* qp = reinterpret_cast<const uint4*>(desc + (size_t)(a + i) * 32);
            const uint4 x = __ldg(qp), y = __ldg(qp + 1);
            q[0] = x.x; q[1] = x.y; q[2] = x.z; q[3] = x.w; q[4] = y.x; q[5] = y.y; q[6] = y.z; q[7] = y.w;
            d = hamming8(q, reinterpret_cast<const uint4*>(desc + (size_t)(a + j) * 32));
        }
        dmat[i * N + j] = (uint16_t)d;
        dmat[j * N + i] = (uint16_t)d;
    }
    __syncthreads();
    const int k = (int)(0.5 * (N - 1));   // index of the median in the sorted row
    for (int i = tid; i < N; i += kDistinctThreads) {
        const uint16_t* row = dmat + i * N;
        int lo = 0, hi = 256;             // smallest v with #{d <= v} >= k + 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j = 0; j < N; ++j) c += row[j] <= mid;
            if (c >= k + 1) hi = mid; else lo = mid + 1;
        }
        atomicMin(&s_best, (lo << 8) | i);   // N <= 128 < 256: the row index breaks ties towards the first row
    }
    __syncthreads();
    const int b = s_best & 0xFF;
    if (tid == 0) best_idx[p] = b;
    if (best_desc && tid < 32) best_desc[(size_t)p * 32 + tid] = desc[(size_t)(a + b) * 32 + tid];
}

}  // namespace adb

using namespace adb;


extern "C" {

int32_t adb_hamming_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t x, y;
        memcpy(&x, a + 8 * i, 8);
        memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

adb_status adb_matcher_create(int32_t device, adb_matcher_t* out) {
    ADB_CHECK(out, ADB_ERR_INVALID, "null argument");
    *out = nullptr;
    adb_status s = select_device(device);
    if (s != ADB_OK) return s;
    adb_matcher* m = new adb_matcher();
    m->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    *out = m;
    return ADB_OK;
}

adb_status adb_matcher_destroy(adb_matcher_t m) {
    if (!m) return ADB_OK;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    cudaStreamDestroy(m->stream);
    cudaFree(m->d_scratch);
    if (m->h_scratch) cudaFreeHost(m->h_scratch);
    if (m->ev[0]) { cudaEventDestroy(m->ev[0]); cudaEventDestroy(m->ev[1]); }
    delete m;
    return ADB_OK;
}

adb_status adb_match_best2_device(adb_matcher_t m, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt,
                                  const int32_t* cand_off, const int32_t* cand_idx, int32_t* best_idx, int32_t* best_d,
                                  int32_t* second_d, void* stream) {
    ADB_CHECK(m && best_idx && best_d && second_d, ADB_ERR_INVALID, "null argument");
    ADB_CHECK(nq >= 0 && nt >= 0 && (!cand_off || cand_idx), ADB_ERR_INVALID, "bad sizes");
    if (nq == 0) return ADB_OK;
    ADB_CHECK(q && (t || nt == 0), ADB_ERR_INVALID, "null descriptors");
    ADB_CHECK((((uintptr_t)q | (uintptr_t)t) & 15) == 0, ADB_ERR_INVALID, "descriptor arrays must be 16-byte aligned");
    ADB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    hamming_best2_kernel<<<(nq + kMatchWarps - 1) / kMatchWarps, kMatchWarps * 32, 0, st>>>(q, nq, t, nt, cand_off, cand_idx,
                                                                                           best_idx, best_d, second_d);
    ADB_CUDA(cudaGetLastError());
    return ADB_OK;
}

adb_status adb_match_best2(adb_matcher_t m, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, const int32_t* cand_off,
                           const int32_t* cand_idx, int32_t* best_idx, int32_t* best_d, int32_t* second_d) {
    ADB_CHECK(m && best_idx && best_d && second_d, ADB_ERR_INVALID, "null argument");
    if (nq <= 0) return ADB_OK;
    ADB_CHECK(nt >= 0 && (!cand_off || cand_idx), ADB_ERR_INVALID, "bad sizes");
    if (cand_off) {   // candidate lists index the train descriptors on the device: offsets monotone from 0, indices inside [0, nt)
        ADB_CHECK(cand_off[0] == 0, ADB_ERR_INVALID, "cand_off must start at 0");
        for (int i = 0; i < nq; ++i) ADB_CHECK(cand_off[i + 1] >= cand_off[i], ADB_ERR_INVALID, "cand_off is not monotone at %d", i);
        for (int c = 0; c < cand_off[nq]; ++c)
            ADB_CHECK(cand_idx[c] >= 0 && cand_idx[c] < nt, ADB_ERR_INVALID, "cand_idx[%d] = %d outside [0, %d)", c, cand_idx[c], nt);
    }
    ADB_CUDA(cudaSetDevice(m->device));
    const int ncand = cand_off ? cand_off[nq] : 0;
    uint8_t *dq = nullptr, *dt = nullptr;
    int32_t *doff = nullptr, *didx = nullptr, *dout = nullptr;
    adb_status s = ADB_OK;
    auto body = [&]() -> adb_status {
        ADB_CUDA(cudaMalloc(&dq, (size_t)nq * 32));
        ADB_CUDA(cudaMalloc(&dt, std::max<size_t>((size_t)nt * 32, 32)));
        ADB_CUDA(cudaMalloc(&dout, (size_t)nq * 12));
        ADB_CUDA(cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
        if (nt > 0) ADB_CUDA(cudaMemcpyAsync(dt, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
        if (cand_off) {
            ADB_CUDA(cudaMalloc(&doff, (size_t)(nq + 1) * 4));
            ADB_CUDA(cudaMalloc(&didx, std::max<size_t>((size_t)ncand * 4, 4)));
            ADB_CUDA(cudaMemcpyAsync(doff, cand_off, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, m->stream));
            if (ncand > 0) ADB_CUDA(cudaMemcpyAsync(didx, cand_idx, (size_t)ncand * 4, cudaMemcpyHostToDevice, m->stream));
        }
        adb_status r = adb_match_best2_device(m, dq, nq, dt, nt, doff, didx, dout, dout + nq, dout + 2 * nq, m->stream);
        if (r != ADB_OK) return r;
        ADB_CUDA(cudaMemcpyAsync(best_idx, dout, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
        ADB_CUDA(cudaMemcpyAsync(best_d, dout + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
        ADB_CUDA(cudaMemcpyAsync(second_d, dout + 2 * nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
        ADB_CUDA(cudaStreamSynchronize(m->stream));
        return ADB_OK;
    };
    s = body();
    cudaFree(dq); cudaFree(dt); cudaFree(doff); cudaFree(didx); cudaFree(dout);
    return s;
}


adb_status adb_distinctive_descriptors(adb_matcher_t m, const uint8_t* desc, const int32_t* point_ptr, int32_t n_points, int32_t* best_idx,
                                       uint8_t* best_desc) {
    ADB_CHECK(m && point_ptr && best_idx, ADB_ERR_INVALID, "null argument");
    if (n_points <= 0) return ADB_OK;
    const int n = point_ptr[n_points];
    ADB_CHECK(point_ptr[0] >= 0, ADB_ERR_INVALID, "point_ptr starts below 0");
    int maxN = 0;
    for (int p = 0; p < n_points; ++p) {
        ADB_CHECK(point_ptr[p + 1] >= point_ptr[p], ADB_ERR_INVALID, "point_ptr is not monotone at %d", p);
        maxN = std::max(maxN, point_ptr[p + 1] - point_ptr[p]);
    }
    ADB_CHECK(maxN <= ADB_MAX_OBSERVATIONS, ADB_ERR_CAPACITY, "a map point has %d observations (max %d)", maxN, ADB_MAX_OBSERVATIONS);
    ADB_CHECK(n == 0 || desc, ADB_ERR_INVALID, "null descriptors");
    ADB_CUDA(cudaSetDevice(m->device));
    uint8_t *dd = nullptr, *dbd = nullptr;
    int32_t *dp = nullptr, *dbi = nullptr;
    auto body = [&]() -> adb_status {
        ADB_CUDA(cudaMalloc(&dd, std::max<size_t>((size_t)n * 32, 32)));
        ADB_CUDA(cudaMalloc(&dp, (size_t)(n_points + 1) * 4));
        ADB_CUDA(cudaMalloc(&dbi, (size_t)n_points * 4));
        if (best_desc) ADB_CUDA(cudaMalloc(&dbd, (size_t)n_points * 32));
        if (n > 0) ADB_CUDA(cudaMemcpyAsync(dd, desc, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
        ADB_CUDA(cudaMemcpyAsync(dp, point_ptr, (size_t)(n_points + 1) * 4, cudaMemcpyHostToDevice, m->stream));
        distinctive_kernel<<<n_points, kDistinctThreads, (size_t)std::max(maxN * maxN, 1) * 2, m->stream>>>(dd, dp, dbi, dbd);
        ADB_CUDA(cudaGetLastError());
        ADB_CUDA(cudaMemcpyAsync(best_idx, dbi, (size_t)n_points * 4, cudaMemcpyDeviceToHost, m->stream));
        if (best_desc) ADB_CUDA(cudaMemcpyAsync(best_desc, dbd, (size_t)n_points * 32, cudaMemcpyDeviceToHost, m->stream));
        ADB_CUDA(cudaStreamSynchronize(m->stream));
        return ADB_OK;
    };
    const adb_status st = body();
    cudaFree(dd); cudaFree(dp); cudaFree(dbi); cudaFree(dbd);
    return st;
}

}  // extern "C"

// Stereo matching of frames [f0, f0 + n) of two handles whose results are resident, on stream `st` (the caller has ordered `st`
// behind both extractions).  Shared by adb_stereo_match_device and the chunk pipeline of adb_stereo_frames_batch (orb.cu).
adb_status adb_stereo_match_range(adb_orb* L, adb_orb* R, int f0, int n, float mb, float mbf, cudaStream_t st) {
    ADB_CHECK(L->cfg.device == R->cfg.device && L->nlevels == R->nlevels && L->capacity == R->capacity &&
                  L->cfg.width == R->cfg.width && L->cfg.height == R->cfg.height,
              ADB_ERR_INVALID, "left / right extractors differ in configuration");
    ADB_CHECK(mb > 0.f, ADB_ERR_INVALID, "baseline must be positive");
    const size_t per = (size_t)L->cfg.max_batch * L->capacity;
    if (!L->d_uright) {
        ADB_CUDA(cudaMalloc(&L->d_uright, per * 4));
        ADB_CUDA(cudaMemset(L->d_uright, 0, per * 4));   // slots past a frame's count travel to the host with the row
        ADB_CUDA(cudaMalloc(&L->d_depth, per * 4));
        ADB_CUDA(cudaMemset(L->d_depth, 0, per * 4));
        ADB_CUDA(cudaMalloc(&L->d_best_idx, per * 4));
        ADB_CUDA(cudaMemset(L->d_best_idx, 0, per * 4));
        ADB_CUDA(cudaMalloc(&L->d_best_dist, per * 4));
        ADB_CUDA(cudaMemset(L->d_best_dist, 0, per * 4));
        ADB_CUDA(cudaMalloc(&L->d_sad, per * 4));
        ADB_CUDA(cudaDeviceSynchronize());   // the fills above run on the legacy stream; the handles' streams do not wait for it
    }
    StereoLevels sl;
    memset(&sl, 0, sizeof(sl));
    for (int l = 0; l < L->nlevels; ++l) {
        const LevelDev& a = L->lv[l].d;
        const LevelDev& b = R->lv[l].d;
        StereoLevel& s = sl.l[l];
        s.L = l == 0 ? L->l0_base : L->lv[l].img;
        s.R = l == 0 ? R->l0_base : R->lv[l].img;
        s.w = a.w;
        s.pitchL = l == 0 ? L->l0_pitch : a.pitch;
        s.pitchR = l == 0 ? R->l0_pitch : b.pitch;
        s.fstrideL = l == 0 ? (unsigned)L->l0_fstride : a.frame_stride;
        s.fstrideR = l == 0 ? (unsigned)R->l0_fstride : b.frame_stride;
        s.scale = a.scale; s.inv_scale = a.inv_scale;
    }
    const float maxD = mbf / mb;   // src/Frame.cc:859-861: minZ = mb, minD = 0, maxD = mbf / minZ
    const int cap = L->capacity, n_rows = L->cfg.height;
    const int maxband = 2 * (int)std::ceil(2.0f * L->lv[L->nlevels - 1].d.scale) + 3;   // rows a right key-point's band can cover
    ADB_CHECK(cap < 65536 && n_rows < 4096, ADB_ERR_INVALID, "capacity %d / height %d too large for the stereo matcher", cap, n_rows);
    if (!L->d_row_ptr) {
        const size_t B = (size_t)L->cfg.max_batch;
        ADB_CUDA(cudaMalloc(&L->d_row_ptr, B * (n_rows + 1) * 4));
        ADB_CUDA(cudaMalloc(&L->d_row_items, B * cap * maxband * 2));
        ADB_CUDA(cudaMalloc(&L->d_rinfo, B * cap * 8));
    }
    stereo_bucket_kernel<<<n, kBucketThreads, (size_t)(n_rows + 1) * 4, st>>>(sl, n_rows, R->d_kps, R->d_counts, cap, maxband, L->d_row_ptr,
                                                                             L->d_row_items, (float2*)L->d_rinfo, f0);
    ADB_CUDA(cudaGetLastError());
    dim3 grid((cap + kStereoWarps * kStereoPerWarp - 1) / (kStereoWarps * kStereoPerWarp), n);
    stereo_match_kernel<<<grid, kStereoWarps * 32, 0, st>>>(sl, n_rows, L->d_kps, L->d_desc, L->d_counts, R->d_desc, cap, maxband,
                                                           L->d_row_ptr, L->d_row_items, (const float2*)L->d_rinfo, mbf, maxD, L->d_uright,
                                                           L->d_depth, L->d_best_idx, L->d_best_dist, L->d_sad, f0);
    ADB_CUDA(cudaGetLastError());
    stereo_median_kernel<<<n, 256, 0, st>>>(L->d_counts, cap, L->d_sad, L->d_uright, L->d_depth, f0);
    ADB_CUDA(cudaGetLastError());
    L->launches += 3;
    return ADB_OK;
}

// Asynchronous copy of the stereo outputs of frames [f0, f0 + n) to host arrays laid out [frame][cap] (pointers at frame 0).
adb_status adb_stereo_download_range(adb_orb* L, int f0, int n, float* ur, float* dp, int32_t* bi, int32_t* bd, int cap, cudaStream_t st) {
    const int rows = std::min(cap, L->capacity);
    const size_t sp = (size_t)L->capacity * 4, dpitch = (size_t)cap * 4, so = (size_t)f0 * L->capacity, ho = (size_t)f0 * cap;
    if (ur) ADB_CUDA(cudaMemcpy2DAsync(ur + ho, dpitch, L->d_uright + so, sp, (size_t)rows * 4, n, cudaMemcpyDeviceToHost, st));
    if (dp) ADB_CUDA(cudaMemcpy2DAsync(dp + ho, dpitch, L->d_depth + so, sp, (size_t)rows * 4, n, cudaMemcpyDeviceToHost, st));
    if (bi) ADB_CUDA(cudaMemcpy2DAsync(bi + ho, dpitch, L->d_best_idx + so, sp, (size_t)rows * 4, n, cudaMemcpyDeviceToHost, st));
    if (bd) ADB_CUDA(cudaMemcpy2DAsync(bd + ho, dpitch, L->d_best_dist + so, sp, (size_t)rows * 4, n, cudaMemcpyDeviceToHost, st));
    return ADB_OK;
}

extern "C" {

adb_status adb_stereo_match_device(adb_orb_t L, adb_orb_t R, int32_t n, float mb, float mbf) {
    ADB_CHECK(L && R, ADB_ERR_INVALID, "null handle");
    ADB_CHECK(n >= 1 && n <= L->last_frames && n <= R->last_frames, ADB_ERR_INVALID, "n_frames %d exceeds the resident frames", n);
    ADB_CUDA(cudaSetDevice(L->cfg.device));
    // the right extractor's stream must have finished before the left stream reads its results
    ADB_CUDA(cudaEventRecord(R->ev, R->stream));
    ADB_CUDA(cudaStreamWaitEvent(L->stream, R->ev, 0));
    return adb_stereo_match_range(L, R, 0, n, mb, mbf, L->stream);
}

adb_status adb_stereo_results_device(adb_orb_t L, const float** ur, const float** dp, const int32_t** bi, const int32_t** bd) {
    ADB_CHECK(L && L->d_uright, ADB_ERR_INVALID, "no stereo results resident");
    if (ur) *ur = L->d_uright;
    if (dp) *dp = L->d_depth;
    if (bi) *bi = L->d_best_idx;
    if (bd) *bd = L->d_best_dist;
    return ADB_OK;
}

adb_status adb_stereo_match(adb_orb_t L, adb_orb_t R, int32_t n, float mb, float mbf, float* ur, float* dp, int32_t* bi, int32_t* bd,
                            int32_t cap) {
    ADB_CHECK(ur && dp, ADB_ERR_INVALID, "null output");
    adb_status s = adb_stereo_match_device(L, R, n, mb, mbf);
    if (s != ADB_OK) return s;
    s = adb_stereo_download_range(L, 0, n, ur, dp, bi, bd, cap, L->stream);
    if (s != ADB_OK) return s;
    ADB_CUDA(cudaMemcpyAsync(L->h_counts, L->d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, L->stream));
    ADB_CUDA(cudaStreamSynchronize(L->stream));
    for (int i = 0; i < n; ++i)
        ADB_CHECK(L->h_counts[i] <= cap, ADB_ERR_CAPACITY, "frame %d holds %d key-points, caller capacity %d", i, L->h_counts[i], cap);
    return ADB_OK;
}

}  // extern "C"
