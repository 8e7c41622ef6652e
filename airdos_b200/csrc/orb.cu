// orb.cu -- batched ORB extractor for sm_100a behind the C-ABI of include/airdos_b200.h.
//
// Replaces ORB_SLAM2::ORBextractor (src/ORBextractor.cc) for a batch of F frames:
//
//   pyr_resize_strip_kernel   ComputePyramid (+ the mask pyramid)   src/ORBextractor.cc:1121-1156
//   erode10_tile_kernel       cv::erode of the mask                 src/ORBextractor.cc:1130-1131
//   fast_cells_warp_kernel    per-cell FAST + ini/min rule          src/ORBextractor.cc:791-840   (fast_cells_kernel: fallback)
//   quadtree_kernel           DistributeOctTree                     src/ORBextractor.cc:541-765
//   blur7_level_kernel        GaussianBlur 7x7 of every level       src/ORBextractor.cc:1099-1100
//   orient_describe_kernel    IC_Angle + computeOrbDescriptor       src/ORBextractor.cc:78-148, 1101-1104
//
// None of these is a translation of the reference's loops: the per-cell detector objects become one WARP per cell fed by a TMA box
// load (column walk, scores and NMS in registers, ordered ballot emission), the std::list quad-tree becomes a level-synchronous pass
// over a flat key array with node ids equal to the reference's creation order (which is what fixes its output order), the level blur
// is a shuffle-based register pipeline without shared memory, and a key-point's descriptor reads two TMA boxes (level, blurred
// level).  The arithmetic (fixed-point resize and blur, FAST score, float32 atan polynomial without contraction) is the one pinned in
// oracle/orb_oracle.cpp.
//
// Data layout in HBM: level l of all frames is one [max_batch][h_l][pitch_l] u8 array, pitch a multiple of 16 B so each level is a
// 3-D TMA tensor {w, h, frame}; the blurred pyramid has the same layout.  No 19-px border is stored: nothing on this path reads it
// (FAST starts at pixel 19; IC-angle / rBRIEF stay inside the ROI; the blur reflects at the ROI edge because the reference blurs a
// clone of the ROI).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "orb.cuh"
#include "../../include/airdos_orb_pattern.h"

namespace adb {

// =========================================================================================
// Pyramid: cv::resize INTER_LINEAR u8 (11-bit fixed point), oracle/orb_oracle.cpp:94-130.
// One thread makes 4 horizontally adjacent output pixels (one 32-bit store).
__global__ void __launch_bounds__(128) pyr_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                         size_t sfstride, uint8_t* __restrict__ dst, int dw, int dh,
                                                         int dpitch, size_t dfstride, const int2* __restrict__ xtab,
                                                         const int2* __restrict__ ytab, int f0) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= dw) return;
    const int f = blockIdx.z + f0;
    const int2 cy = __ldg(&ytab[y]);
    const int sy0 = cy.x, sy1 = min(sy0 + 1, sh - 1);
    const int b0 = cy.y & 0xFFFF, b1 = cy.y >> 16;
    const uint8_t* s0 = src + (size_t)f * sfstride + (size_t)sy0 * spitch;
    const uint8_t* s1 = src + (size_t)f * sfstride + (size_t)sy1 * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x < dw) {
            const int2 cx = __ldg(&xtab[x]);
            const int sx0 = cx.x, sx1 = min(sx0 + 1, sw - 1);
            const int a0 = cx.y & 0xFFFF, a1 = cx.y >> 16;
            const int r0 = (int)__ldg(s0 + sx0) * a0 + (int)__ldg(s0 + sx1) * a1;
            const int r1 = (int)__ldg(s1 + sx0) * a0 + (int)__ldg(s1 + sx1) * a1;
            const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xFF) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)f * dfstride + (size_t)y * dpitch + x4) = out;
}

// The same resize as a strip walk (the hot variant; pyr_resize_kernel above stays as the fallback for scale factors
// above 2).  A thread owns 4 adjacent output columns and walks `rows` output rows downwards.  Per source row it loads
// three aligned words, funnel-shifts them into an 8-byte window that starts at its first source pixel, and picks each
// output's two neighbours with one PRMT (selectors are loop invariant); the horizontally interpolated row is kept in
// registers and reused by the next output row (consecutive output rows share a source row), so a source row is
// interpolated ~1.2 times per output row instead of 2 and no byte-wide loads are issued.
// (Measured dead ends: 8 columns per thread and prefetching the next source row one output row ahead both left the stage at
// 0.17 ms per 128 frames -- the seven dependent launches, not the per-pixel work, set the time of the small levels.)
// kMask: the mask pyramid (src/ORBextractor.cc:1146-1147: the same resize of the eroded mask) rides along in the same threads: tables,
// selectors, row bookkeeping and the loop are shared, only the loads and the interpolation are done twice.
template <bool kMask>
__global__ void __launch_bounds__(128) pyr_resize_strip_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                               size_t sfstride, uint8_t* __restrict__ dst, int dw, int dh,
                                                               int dpitch, size_t dfstride, const int2* __restrict__ xtab,
                                                               const int2* __restrict__ ytab, int nq, int rows, int nchunks, int f0,
                                                               const uint8_t* __restrict__ msrc, int mspitch, size_t msfstride,
                                                               uint8_t* __restrict__ mdst, int mdpitch, size_t mdfstride) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= nq * nchunks) return;
    const int chunk = t / nq, q = t - chunk * nq, f = blockIdx.y + f0;
    uint32_t sel[4], ac[4];
    const int base = __ldg(&xtab[4 * q]).x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int2 cx = __ldg(&xtab[min(4 * q + k, dw - 1)]);
        const int o0 = cx.x - base, o1 = min(cx.x + 1, sw - 1) - base;   // <= 7 (checked on the host)
        sel[k] = (uint32_t)(o0 | (o1 << 4));
        ac[k] = (uint32_t)cx.y;                                          // a0 | a1 << 16: the two 16-bit operands of one DP2A
    }
    const int maxw = (spitch >> 2) - 1;                                  // (the mask levels have the pitch of the image levels' own buffers or more)
    const int w0i = base >> 2, w1i = min(w0i + 1, maxw), w2i = min(w0i + 2, maxw);
    const uint32_t sft = (uint32_t)(base & 3) * 8u;
    const uint8_t* sf = src + (size_t)f * sfstride;
    const uint8_t* mf = kMask ? msrc + (size_t)f * msfstride : nullptr;
    auto hrow = [&](const uint8_t* rowp, uint32_t (&H)[4]) {   // (S[sx0] * a0 + S[sx1] * a1) >> 4 for the 4 columns: PRMT puts the two neighbours in bytes 0, 1
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(rowp);
        const uint32_t w0 = __ldg(sp + w0i), w1 = __ldg(sp + w1i), w2 = __ldg(sp + w2i);
        const uint32_t W0 = __funnelshift_r(w0, w1, sft), W1 = __funnelshift_r(w1, w2, sft);
#pragma unroll
        for (int k = 0; k < 4; ++k) H[k] = __dp2a_lo(ac[k], __byte_perm(W0, W1, sel[k]), 0u) >> 4;
    };
    uint32_t A[4] = {0, 0, 0, 0}, B[4] = {0, 0, 0, 0}, MA[4] = {0, 0, 0, 0}, MB[4] = {0, 0, 0, 0};
    int ia = -1, ib = -1;
    const int y_end = min(dh, (chunk + 1) * rows);
    uint8_t* dp = dst + (size_t)f * dfstride + 4 * q;
    uint8_t* mp = kMask ? mdst + (size_t)f * mdfstride + 4 * q : nullptr;
    for (int y = chunk * rows; y < y_end; ++y) {
        const int2 cy = __ldg(&ytab[y]);
        const int sy0 = cy.x, sy1 = min(sy0 + 1, sh - 1);
        const uint32_t b0s = (uint32_t)cy.y << 16, b1s = (uint32_t)cy.y & 0xFFFF0000u;   // (b * r) >> 16 = umulhi(b << 16, r)
        if (sy0 != ia) {
            if (sy0 == ib) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { A[k] = B[k]; if (kMask) MA[k] = MB[k]; }
            } else {
                hrow(sf + (size_t)sy0 * spitch, A);
                if (kMask) hrow(mf + (size_t)sy0 * mspitch, MA);
            }
            ia = sy0;
        }
        if (sy1 != ib) {
            if (sy1 == ia) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { B[k] = A[k]; if (kMask) MB[k] = MA[k]; }
            } else {
                hrow(sf + (size_t)sy1 * spitch, B);
                if (kMask) hrow(mf + (size_t)sy1 * mspitch, MB);
            }
            ib = sy1;
        }
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (__umulhi(b1s, B[k]) + (__umulhi(b0s, A[k]) + 2u)) >> 2;   // <= 255
        *reinterpret_cast<uint32_t*>(dp + (size_t)y * dpitch) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        if (kMask) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (__umulhi(b1s, MB[k]) + (__umulhi(b0s, MA[k]) + 2u)) >> 2;
            *reinterpret_cast<uint32_t*>(mp + (size_t)y * mdpitch) = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        }
    }
}

// cv::erode(mask, ones(10,10)), anchor (5,5), outside = 255 (src/ORBextractor.cc:1130-1131; oracle/orb_oracle.cpp:70-92): the
// reference passes a mask with EVERY frame (src/Frame.cc:551-571, System.IsMask), so this is on the measured path.  Separable:
// one CTA stages a (16 + 9) x (128 + 9) tile in shared memory, takes the 10-wide row minimum in place, then the 10-high column
// minimum -- each mask byte is read from HBM once instead of 100 times.
constexpr int kErTW = 128, kErTH = 32, kErR0 = 5, kErSpan = 10, kErRun = 16;
// running minimum over 10 of a run of kErRun + 9 values held in registers, by doubling: 2, 4, 8, then 8 + 2 = 10 -> 5 min per output
__device__ __forceinline__ void min10_run(const int (&a)[kErRun + kErSpan - 1], int (&o)[kErRun]) {
    int m2[kErRun + 8], m4[kErRun + 6], m8[kErRun + 2];
#pragma unroll
    for (int i = 0; i < kErRun + 8; ++i) m2[i] = min(a[i], a[i + 1]);
#pragma unroll
    for (int i = 0; i < kErRun + 6; ++i) m4[i] = min(m2[i], m2[i + 2]);
#pragma unroll
    for (int i = 0; i < kErRun + 2; ++i) m8[i] = min(m4[i], m4[i + 4]);
#pragma unroll
    for (int i = 0; i < kErRun; ++i) o[i] = min(m8[i], m2[i + 8]);
}
__global__ void __launch_bounds__(256) erode10_tile_kernel(const uint8_t* __restrict__ src, int w, int h, int spitch, size_t sfstride,
                                                           uint8_t* __restrict__ dst, int dpitch, size_t dfstride, int f0) {
    // input tile: rows y0 - 5 .. y0 + 36, columns x0 - 8 .. x0 + 135 (word aligned: x0 is a multiple of 128), 36 words per row
    constexpr int IH = kErTH + kErSpan - 1, IWW = (kErTW + 16) / 4, OFF = 8 - kErR0;
    __shared__ uint32_t tile[IH][IWW + 1];
    __shared__ uint32_t rmin[IH][kErTW / 4 + 1];
    __shared__ uint32_t outt[kErTH][kErTW / 4 + 1];
    const int x0 = blockIdx.x * kErTW, y0 = blockIdx.y * kErTH, f = blockIdx.z;
    const uint8_t* sp = src + (size_t)f * sfstride;
    const bool aligned = (((uintptr_t)sp | (uintptr_t)spitch) & 3) == 0;
    int mixed = 0;                                                  // bit 0: a word that is not all 255, bit 1: a word that is not all 0
    for (int i = threadIdx.x; i < IWW * IH; i += 256) {
        const int r = i / IWW, cw = i - r * IWW, yy = y0 - kErR0 + r, xx = x0 - 8 + 4 * cw;
        uint32_t v = 0xFFFFFFFFu;                                   // outside the image = 255 (cv::erode's border value)
        if (yy >= 0 && yy < h) {
            const uint8_t* rowp = sp + (size_t)yy * spitch;
            if (aligned && xx >= 0 && xx + 3 < w) v = *reinterpret_cast<const uint32_t*>(rowp + xx);
            else {
                v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) v |= (uint32_t)((xx + k >= 0 && xx + k < w) ? rowp[xx + k] : (uint8_t)255) << (8 * k);
            }
        }
        tile[r][cw] = v;
        mixed |= (v != 0xFFFFFFFFu ? 1 : 0) | (v != 0u ? 2 : 0);
    }
    // A constant input tile (halo included: a superset of every window of the tile) erodes to the same constant.  Segmentation masks
    // are 255 almost everywhere (src/Frame.cc:551-571: 0 only on the person), so most tiles leave here with one load and one store pass.
    const bool not255 = __syncthreads_or(mixed & 1) != 0;           // (the barrier the row pass needs anyway; the vote is a logical OR, one per bit)
    const bool not0 = __syncthreads_or(mixed & 2) != 0;
    const bool constant = !(not255 && not0);
    const uint32_t cval = not255 ? 0u : 0xFFFFFFFFu;
    if (!constant) {                                                // block-uniform: the barriers inside are reached by all or none
    const uint8_t* tb = reinterpret_cast<const uint8_t*>(&tile[0][0]);
    uint8_t* rb = reinterpret_cast<uint8_t*>(&rmin[0][0]);
    // rows: a thread owns a run of 16 outputs of one row (8 runs per row, 41 rows = 328 runs)
    for (int i = threadIdx.x; i < IH * (kErTW / kErRun); i += 256) {
        const int r = i / (kErTW / kErRun), c0 = (i - r * (kErTW / kErRun)) * kErRun;
        int a[kErRun + kErSpan - 1], o[kErRun];
        const uint8_t* tr = tb + (size_t)r * (IWW + 1) * 4 + OFF + c0;
#pragma unroll
        for (int k = 0; k < kErRun + kErSpan - 1; ++k) a[k] = tr[k];
        min10_run(a, o);
#pragma unroll
        for (int k = 0; k < kErRun; k += 4)
            rmin[r][(c0 + k) >> 2] = (uint32_t)o[k] | ((uint32_t)o[k + 1] << 8) | ((uint32_t)o[k + 2] << 16) | ((uint32_t)o[k + 3] << 24);
    }
    __syncthreads();
    // columns: a thread owns a run of 16 outputs of one column (2 runs per column, 128 columns = 256 runs = one per thread)
    {
        const int c = threadIdx.x % kErTW, r0 = (threadIdx.x / kErTW) * kErRun;
        int a[kErRun + kErSpan - 1], o[kErRun];
#pragma unroll
        for (int k = 0; k < kErRun + kErSpan - 1; ++k) a[k] = rb[(size_t)(r0 + k) * (kErTW / 4 + 1) * 4 + c];
        min10_run(a, o);
        uint8_t* ob = reinterpret_cast<uint8_t*>(&outt[0][0]);
#pragma unroll
        for (int k = 0; k < kErRun; ++k) ob[(size_t)(r0 + k) * (kErTW / 4 + 1) * 4 + c] = (uint8_t)o[k];
    }
    __syncthreads();
    }
    // store: 32 x 32 words per tile, coalesced; byte-wise only at a ragged right edge or an unaligned destination
    uint8_t* dp = dst + (size_t)(f0 + f) * dfstride;
    const bool daligned = (((uintptr_t)dp | (uintptr_t)dpitch) & 3) == 0;
    for (int i = threadIdx.x; i < kErTH * (kErTW / 4); i += 256) {
        const int r = i / (kErTW / 4), cw = i - r * (kErTW / 4), y = y0 + r, x = x0 + 4 * cw;
        if (y >= h || x >= w) continue;
        const uint32_t v = constant ? cval : outt[r][cw];
        if (daligned && x + 3 < w) *reinterpret_cast<uint32_t*>(dp + (size_t)y * dpitch + x) = v;
        else for (int k = 0; k < 4 && x + k < w; ++k) dp[(size_t)y * dpitch + x + k] = (uint8_t)(v >> (8 * k));
    }
}

// =========================================================================================
// FAST-9/16 per cell.  One CTA per (cell, frame): the (cell + 6 px) tile arrives through one TMA
// box load; scores only for the cell's own pixels (the 3-px frame of the sub-image is never
// tested and counts as 0 in the NMS, which is what makes the reference's NMS blind across
// cell seams); strict 3x3 NMS; mask post-filter; the ini / min threshold rule; row-major
// ordered compaction into the cell's slot.
struct MaskPtrs {
    const uint8_t* p[kMaxLevels];
};

// The 16 ring pixels q_i of centre v as two 16-bit lanes per word, made by one IMAD each (fma pipe; everything
// after it is one alu-pipe op per three words): x_i = a + q_i * 0xFFFF = (a.lo - q_i) | (a.hi + q_i) << 16.  The
// lanes stay inside (0, 0x10000) for the two choices of `a` below, so nothing carries across.
// OpenCV circle order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
__device__ __forceinline__ void load_ring_packed(const uint8_t* c, int bw, uint32_t a, uint32_t (&x)[16]) {
#define ADB_Q(off) ((uint32_t)c[off] * 0xFFFFu + a)
    x[0] = ADB_Q(3 * bw);      x[1] = ADB_Q(3 * bw + 1);  x[2] = ADB_Q(2 * bw + 2);   x[3] = ADB_Q(bw + 3);
    x[4] = ADB_Q(3);           x[5] = ADB_Q(-bw + 3);     x[6] = ADB_Q(-2 * bw + 2);  x[7] = ADB_Q(-3 * bw + 1);
    x[8] = ADB_Q(-3 * bw);     x[9] = ADB_Q(-3 * bw - 1); x[10] = ADB_Q(-2 * bw - 2); x[11] = ADB_Q(-bw - 3);
    x[12] = ADB_Q(-3);         x[13] = ADB_Q(bw - 3);     x[14] = ADB_Q(2 * bw - 2);  x[15] = ADB_Q(3 * bw - 1);
#undef ADB_Q
}

// The same from a 32-bit shared-memory address (warp-per-cell kernel: one register per pointer); a = (256 + v) | (256 - v) << 16 by one IMAD.
template <int kOff>
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(kOff));
    return v;
}
template <int kOff>
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(kOff));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Ring rows of the single-tile FAST kernel: every pixel of the cell box is converted ONCE to X = (q + 256) | (256 - q) << 16 (one IMAD)
// when its row enters an 8-row ring of 32-bit words, instead of once per (centre, ring position) pair = 16 times.  The packed minimum
// of an arc is then (256 + min q, 256 - max q) and the centre enters only at the very end: with A = max over arcs of the low half and
// H of the high half, 256 + score = max(A - v, H + v) = the halves of the packed sum (A | H << 16) + v * 0xFFFF.
constexpr int kRingPitch = 40, kRingRows = 8;          // words per ring row (cell + 6 <= 40 columns: single-tile cells are <= 32 wide)
template <int J>                                        // J = (inner row) mod 8: the slot of a box row is (row & 7), so all offsets are immediates
__device__ __forceinline__ void load_ring_words(uint32_t rb, uint32_t (&x)[16]) {   // rb: ring address of this lane's centre COLUMN, slot 0
#define ADB_R(ox, oy) lds_u32<(((J + 3 + (oy)) & 7) * kRingPitch + (ox)) * 4>(rb)
    x[0] = ADB_R(0, 3);    x[1] = ADB_R(1, 3);    x[2] = ADB_R(2, 2);    x[3] = ADB_R(3, 1);
    x[4] = ADB_R(3, 0);    x[5] = ADB_R(3, -1);   x[6] = ADB_R(2, -2);   x[7] = ADB_R(1, -3);
    x[8] = ADB_R(0, -3);   x[9] = ADB_R(-1, -3);  x[10] = ADB_R(-2, -2); x[11] = ADB_R(-3, -1);
    x[12] = ADB_R(-3, 0);  x[13] = ADB_R(-3, 1);  x[14] = ADB_R(-2, 2);  x[15] = ADB_R(-1, 3);
#undef ADB_R
}

template <int bw>
__device__ __forceinline__ void load_ring_packed_s(uint32_t c, uint32_t (&x)[16]) {
    const uint32_t a = lds_u8<0>(c) * 0xFFFF0001u + 0x01000100u;
#define ADB_Q(off) (lds_u8<(off)>(c) * 0xFFFFu + a)
    x[0] = ADB_Q(3 * bw);      x[1] = ADB_Q(3 * bw + 1);  x[2] = ADB_Q(2 * bw + 2);   x[3] = ADB_Q(bw + 3);
    x[4] = ADB_Q(3);           x[5] = ADB_Q(-bw + 3);     x[6] = ADB_Q(-2 * bw + 2);  x[7] = ADB_Q(-3 * bw + 1);
    x[8] = ADB_Q(-3 * bw);     x[9] = ADB_Q(-3 * bw - 1); x[10] = ADB_Q(-2 * bw - 2); x[11] = ADB_Q(-bw - 3);
    x[12] = ADB_Q(-3);         x[13] = ADB_Q(bw - 3);     x[14] = ADB_Q(2 * bw - 2);  x[15] = ADB_Q(3 * bw - 1);
#undef ADB_Q
}

// Exact score: max over the 16 arcs of 9 of min(v - q) and of min(q - v) (cv::cornerScore<16>), again both
// polarities at once: a = (256 + v) | (256 - v) << 16 makes the lanes 256 + d and 256 - d, and the arc minima / the
// final maximum are packed 3-input min / max (VIMNMX3.U16x2): 40 of them per pixel.
__device__ __forceinline__ uint32_t fast_score_a(int v) { return (uint32_t)(256 + v) | ((uint32_t)(256 - v) << 16); }
__device__ __forceinline__ uint32_t fast_best_packed(const uint32_t (&x)[16]);
__device__ __forceinline__ int fast_best(const uint32_t (&x)[16]) {
    const uint32_t r0 = fast_best_packed(x);
    return (int)max(r0 & 0xFFFFu, r0 >> 16) - 256;
}
// both polarities, still packed: max over the 16 arcs of the arc minimum, + 256 in each 16-bit lane
__device__ __forceinline__ uint32_t fast_best_packed(const uint32_t (&x)[16]) {
    // (Measured dead end, twice: the arc minima by prefix / suffix minima of the two half rings (van Herk) -- 59 two-input packed
    // min / max instead of these 39 three-input ones -- ran slower in both FAST kernels (8.2 vs 7.8 and 8.3 vs 7.5 ms per 2048 frames),
    // also in a register-lean interleaving without spills: inside this instruction mix a two-input packed min / max costs the alu pipe
    // what a three-input one does, whatever the isolated stream of tools/probe/pipe_probe.cu suggests.)
    uint32_t m3[16], m9[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m3[i] = __vimin3_u16x2(x[i], x[(i + 1) & 15], x[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) m9[i] = __vimin3_u16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
    uint32_t r0 = __vimax3_u16x2(m9[0], m9[1], m9[2]), r1 = __vimax3_u16x2(m9[3], m9[4], m9[5]);
    uint32_t r2 = __vimax3_u16x2(m9[6], m9[7], m9[8]), r3 = __vimax3_u16x2(m9[9], m9[10], m9[11]);
    uint32_t r4 = __vimax3_u16x2(m9[12], m9[13], m9[14]);
    r0 = __vimax3_u16x2(r0, r1, r2);
    r3 = __vimax3_u16x2(r3, r4, m9[15]);
    return __vmaxu2(r0, r3);
}

constexpr int kFastThreads = 256;
constexpr int kMaxKeep = 1152;      // >= ceil(75 / 2) * ceil(60 / 2): strict 3x3 NMS keeps at most one pixel per 2x2 block

// Phases per (cell, frame) CTA, each a dense loop (no divergent heavy branch):
//   1 exact FAST score of every pixel of the cell (packed, branch free); a corner at min(iniTh, minTh) keeps
//     score - 1 in the score tile, everything else 0.  (A packed LOP3 corner test costs the same 40 alu-pipe slots as the
//     score itself, so "test first, score the corners" was measured slower: 0.80 vs 0.75 ms per 128 frames.)
//   3 strict 3x3 NMS + mask post-filter of the corners -> unordered kept list
//   4 ini / min rule, then each kept corner's rank in row-major order places it in the cell's slot
// kBW = pitch of the TMA box in shared memory (one value per handle), a compile-time constant so that the 16 ring
// loads are immediate offsets from one address register.
template <int kBW>
__global__ void __launch_bounds__(kFastThreads) fast_cells_kernel(const __grid_constant__ TmaMaps16 maps,
                                                                 const LevelDev* __restrict__ levels,
                                                                 const uint32_t* __restrict__ cell_table, const __grid_constant__ MaskPtrs masks,
                                                                 int ini_th, int min_th, uint32_t* __restrict__ cand,
                                                                 int cand_total, uint16_t* __restrict__ cellcnt,
                                                                 int ncells_total, int f0) {
    __shared__ __align__(128) uint8_t tile[kCellBoxHMax * kBW];
    __shared__ __align__(16) uint8_t score[kCellBoxHMax * kBW];
    __shared__ uint64_t bar;
    __shared__ uint32_t klist[kMaxKeep];
    __shared__ int n_keep, n_sel;

    const int tid = threadIdx.x, f = blockIdx.y + f0;
    const uint32_t ce = __ldg(&cell_table[blockIdx.x]);
    const int level = ce >> 24, ci = (ce >> 12) & 0xFFF, cj = ce & 0xFFF;
    const LevelDev& L = levels[level];
    const int maxBX = L.w - kMinBorder, maxBY = L.h - kMinBorder;
    const int iniY = kMinBorder + ci * L.hcell, iniX = kMinBorder + cj * L.wcell;
    const int maxY = min(iniY + L.hcell + 6, maxBY), maxX = min(iniX + L.wcell + 6, maxBX);
    const int cw = maxX - iniX, ch = maxY - iniY;
    uint16_t* cnt_out = cellcnt + (size_t)f * ncells_total + blockIdx.x;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6 || cw < 7 || ch < 7) {   // the reference's `continue`s
        if (tid == 0) *cnt_out = 0;
        return;
    }
    constexpr int bw = kBW;
    const int bh = L.box_h;
    if (tid == 0) {
        n_keep = 0; n_sel = 0;
        mbar_init(&bar, 1);
        mbar_fence_init();
        fence_proxy_async();
        mbar_expect_tx(&bar, (uint32_t)(bw * bh));
        tma_load_3d(tile, &maps.m[level], &bar, iniX & ~15, iniY, f);   // box origin must be 16-B aligned in x
    }
    const int dx = iniX & 15;
    // clear the score tile while the box is in flight (its 3-px frame must read 0 in the NMS)
    for (int i = tid; i < bw * bh / 4; i += kFastThreads) reinterpret_cast<uint32_t*>(score)[i] = 0;
    __syncthreads();   // barrier init visible to all waiters
    mbar_wait(&bar, 0);

    const int iw = cw - 6, ih = ch - 6, npx = iw * ih;
    const uint32_t rcp = 0xFFFFFFFFu / (uint32_t)iw + 1u;   // p / iw == umulhi(p, rcp) for p < 65536
    const int t_low = min(ini_th, min_th);

    // ---- phase 1
    for (int p = tid; p < npx; p += kFastThreads) {
        const int y0 = (int)__umulhi((uint32_t)p, rcp);
        const int o = (y0 + 3) * bw + (p - y0 * iw + 3);
        const uint8_t* c = tile + o + dx;
        uint32_t ring[16];
        load_ring_packed(c, bw, fast_score_a(c[0]), ring);
        const int best = fast_best(ring);
        score[o] = (uint8_t)(best > t_low ? best - 1 : 0);   // response = best - 1 (>= 1 for every corner)
    }
    __syncthreads();

    // ---- phase 3: NMS + mask for the corners -> unordered kept list (p | score << 16 | flags << 24;
    //      flag bit 0: kept at minTh, bit 1: kept at iniTh)
    const uint8_t* ml = masks.p[level];
    int mineB = 0;
    for (int p = tid; p < npx; p += kFastThreads) {
        const int y0 = (int)__umulhi((uint32_t)p, rcp);
        const int y = y0 + 3, x = p - y0 * iw + 3;
        const uint8_t* s = score + y * bw + x;
        const int v = s[0];
        if (v == 0) continue;
        bool keep = v > s[-1] && v > s[1] && v > s[-bw - 1] && v > s[-bw] && v > s[-bw + 1] && v > s[bw - 1] && v > s[bw] && v > s[bw + 1];
        if (keep && ml) keep = ml[(size_t)f * L.mframe_stride + (size_t)(iniY + y) * L.mpitch + iniX + x] != 0;
        const int fl = keep ? ((v >= min_th ? 1 : 0) | (v >= ini_th ? 2 : 0)) : 0;
        if (fl) {
            const int k = atomicAdd(&n_keep, 1);
            if (k < kMaxKeep) klist[k] = (uint32_t)p | ((uint32_t)v << 16) | ((uint32_t)fl << 24);
        }
        mineB |= fl & 2;
    }
    const int anyB = __syncthreads_or(mineB);
    const uint32_t selmask = (anyB ? 2u : 1u) << 24;

    // ---- phase 4: the reference's row-major order = rank of p among the selected entries (a few dozen per cell)
    const int nk = min(n_keep, kMaxKeep);
    uint32_t* slot = cand + (size_t)f * cand_total + L.cand_base + (size_t)(blockIdx.x - L.cell_base) * L.slotcap;
    int nsel = 0;
    for (int i = tid; i < nk; i += kFastThreads) {
        const uint32_t me = klist[i];
        if (!(me & selmask)) continue;
        const uint32_t myp = me & 0xFFFFu;
        int rank = 0;
        for (int j = 0; j < nk; ++j) {
            const uint32_t o = klist[j];
            rank += ((o & selmask) != 0) & ((o & 0xFFFFu) < myp);
        }
        const int y0 = (int)__umulhi(myp, rcp);
        const int y = y0 + 3, x = (int)myp - y0 * iw + 3;
        if (rank < L.slotcap)   // cannot trigger: slotcap is the strict-NMS bound
            slot[rank] = (uint32_t)(x + cj * L.wcell) | ((uint32_t)(y + ci * L.hcell) << 12) | (((me >> 16) & 0xFFu) << 24);
        ++nsel;
    }
    if (nsel) atomicAdd(&n_sel, nsel);
    __syncthreads();
    if (tid == 0) *cnt_out = (uint16_t)min(n_sel, L.slotcap);
}

// The hot variant: ONE WARP per (cell, frame), eight independent warps per CTA, no block barrier and no score tile.
//   * lane = column of the cell (cells are 30 .. 33 pixels wide at the usual settings); wider cells are walked as column tiles of
//     30 + 2 halo lanes inside the same row step (kFwTiles of them: up to 92 columns, above that the CTA-per-cell kernel runs);
//   * the warp walks down the rows: address += pitch, no index arithmetic; the scores of rows y - 2, y - 1, y stay in registers, so
//     the strict 3 x 3 NMS of row y - 1 is one packed 3-input maximum per column plus two shuffles -- and the reference's blindness
//     across cell seams is free: whatever lies outside the cell is simply a lane that holds 0;
//   * kept corners leave in row-major order as they are found (ballot + popc): corners at iniTh grow from the front of the cell's
//     slot, corners that only reach minTh from its back; the ini / min rule is then just WHICH end the consumer reads (bit 15 of the
//     cell count = "no corner at iniTh: read the back, reversed").  No rank loop, no atomics.
// Per 32 pixels: 17 LDS + 17 IMAD + 41 packed min / max for the score, ~25 instructions for everything else (the CTA-per-cell
// kernel: ~170 on top of the score).
constexpr int kFwRingBytes = 8 * 40 * 4;   // = kRingRows * kRingPitch words: the ring of converted rows behind every warp's TMA box
constexpr int kFwWarps = 8, kFwTiles = 3, kFwStep = 30;   // column tiles advance by 30: 32 lanes minus the two halo lanes
// NT = column tiles per row step: the cells of levels whose cells fit one tile (<= 32 columns) run the NT = 1 instance, the others the
// NT = kFwTiles one (two launches over the two halves of the cell order list).
template <int kBW, int NT, bool kMasked>
__global__ void __launch_bounds__(kFwWarps * 32, NT == 1 ? 6 : 4) fast_cells_warp_kernel(const __grid_constant__ TmaMaps16 maps,
                                                                        const LevelDev* __restrict__ levels,
                                                                        const uint32_t* __restrict__ cell_table, const uint32_t* __restrict__ cell_order,
                                                                        int order_count, const __grid_constant__ MaskPtrs masks,
                                                                        int ini_th, int min_th, uint32_t* __restrict__ cand,
                                                                        int cand_total, uint16_t* __restrict__ cellcnt,
                                                                        int ncells_total, int tile_bytes, int f0) {
    extern __shared__ __align__(128) uint8_t fw_smem[];
    __shared__ uint64_t bars[kFwWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, f = blockIdx.y + f0;
    const int oi = blockIdx.x * kFwWarps + warp;
    if (oi >= order_count) return;
    const int cell = (int)__ldg(&cell_order[oi]);
    uint8_t* tile = fw_smem + ((128u - (smem_u32(fw_smem) & 127u)) & 127u) + (size_t)warp * (tile_bytes + kFwRingBytes);
    uint64_t* bar = &bars[warp];
    const uint32_t ce = __ldg(&cell_table[cell]);
    const int level = ce >> 24, ci = (ce >> 12) & 0xFFF, cj = ce & 0xFFF;
    const LevelDev& L = levels[level];
    const int maxBX = L.w - kMinBorder, maxBY = L.h - kMinBorder;
    const int iniY = kMinBorder + ci * L.hcell, iniX = kMinBorder + cj * L.wcell;
    const int maxY = min(iniY + L.hcell + 6, maxBY), maxX = min(iniX + L.wcell + 6, maxBX);
    const int cw = maxX - iniX, ch = maxY - iniY;
    uint16_t* cnt_out = cellcnt + (size_t)f * ncells_total + cell;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6 || cw < 7 || ch < 7) {   // the reference's `continue`s
        if (lane == 0) *cnt_out = 0;
        return;
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        fence_proxy_async();
        mbar_expect_tx(bar, (uint32_t)(kBW * L.box_h));
        tma_load_3d(tile, &maps.m[level], bar, iniX & ~15, iniY, f);   // box origin must be 16-B aligned in x
    }
    __syncwarp();
    const int iw = cw - 6, ih = ch - 6;
    const int nt = NT == 1 ? 1 : max(1, (iw + 27) / kFwStep);   // tiles start at columns 0, 30, 60; a tile decides lanes 1 .. 30
                                                                 // (+ lane 0 of the first, lane 31 when it is the last column)
    // Scores live in the "z domain": z = 256 + best, clamped from below to zlow = 256 + min(iniTh, minTh) -- every non-corner, every
    // pixel outside the cell and the rows above / below it hold exactly zlow, so a corner (z > zlow) that beats its 8 neighbours is
    // what the reference keeps, and the response is z - 257.
    const uint32_t zlow = 256u + (uint32_t)min(ini_th, min_th), zini = 257u + (uint32_t)ini_th;
    const bool has_b = min_th < ini_th;                          // else every corner already passes iniTh
    const uint8_t* ml = masks.p[level];
    if (kMasked) ml += (size_t)f * L.mframe_stride + (size_t)(iniY + 3) * L.mpitch + iniX + 3 + lane;
    const int mpitch = L.mpitch;
    const uint32_t slot0 = (uint32_t)((size_t)f * cand_total) + (uint32_t)L.cand_base + (uint32_t)(cell - L.cell_base) * (uint32_t)L.slotcap;   // < 2^32: checked at creation
    uint32_t idxA = slot0, idxB = slot0 + (uint32_t)L.slotcap - 1u;   // next free entry of the front (iniTh) / back (minTh only) list
    const uint32_t rec0 = ((uint32_t)(3 + cj * L.wcell + lane) | ((uint32_t)(3 + ci * L.hcell) << 12)) - (257u << 24);   // + (row << 12) + (z << 24)
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t zA[NT], zB[NT], zcap[NT];        // scores of rows y - 2, y - 1 of this lane's column, per tile
    bool decide[NT];
    uint32_t cp[NT];                          // shared-memory address of the centre pixel
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        zA[t] = zlow; zB[t] = zlow;
        const int col = t * kFwStep + lane;
        decide[t] = t < nt && col < iw && (lane > 0 || t == 0) && (lane < 31 || col == iw - 1);
        zcap[t] = col < iw ? 0xFFFFu : zlow;
        cp[t] = smem_u32(tile) + 3 * kBW + (iniX & 15) + 3 + min(col, iw - 1);
    }
    // NMS of the row above the one just scored (z = its scores; zlow for the flush step) and ordered emission of the kept corners
    // (mv: single-tile instance only -- this lane's mask byte of row yr, loaded at the top of the row step; the wide-cell instance reads it on demand)
    auto nms_emit = [&](const int t, const uint32_t z, const int yr, const uint32_t mv) {
        const uint32_t v = zB[t], vert = max(zA[t], z);
        const uint32_t c3 = max(vert, v);
        uint32_t cl = __shfl_up_sync(0xFFFFFFFFu, c3, 1), cr = __shfl_down_sync(0xFFFFFFFFu, c3, 1);
        if (lane == 0) cl = 0;
        if (lane == 31) cr = 0;
        bool keep = decide[t] && v > max(max(cl, cr), vert);
        zA[t] = v; zB[t] = z;
        if (kMasked) {
            if (NT == 1) keep = keep && mv != 0;
            else if (keep) keep = ml[(size_t)yr * mpitch + t * kFwStep] != 0;
        }
        const uint32_t balK = __ballot_sync(0xFFFFFFFFu, keep);
        if (balK) {
            const bool isA = keep && v >= zini;
            const uint32_t balA = __ballot_sync(0xFFFFFFFFu, isA), balB = has_b ? balK & ~balA : 0u;
            const uint32_t idx = isA ? idxA + __popc(balA & lt) : idxB - __popc(balB & lt);
            if (keep && (isA || has_b)) cand[idx] = rec0 + (uint32_t)(t * kFwStep) + ((uint32_t)yr << 12) + (v << 24);
            idxA += __popc(balA); idxB -= __popc(balB);
        }
    };
    mbar_wait(bar, 0);
    if constexpr (NT == 1) {
        // single-tile cells: rows pass through the ring of converted words (see load_ring_words)
        const uint32_t ring0 = smem_u32(tile) + (uint32_t)tile_bytes;                       // this warp's ring, behind its TMA box
        const uint32_t rb = ring0 + 4u * (3u + lane);                                         // centre column of this lane, slot 0
        uint32_t rawp = smem_u32(tile) + (iniX & 15) + lane;                                  // box row 0, sub-image column `lane`
        const bool second = lane + 32 < cw;                                                   // columns 32 .. cw - 1 (cw <= 38)
        auto convert_row = [&](uint32_t slot_addr) {                                          // box row at rawp -> ring row at slot_addr
            sts_u32(slot_addr + 4u * lane, lds_u8<0>(rawp) * 0xFFFF0001u + 0x01000100u);       // (q + 256) | (256 - q) << 16
            if (second) sts_u32(slot_addr + 4u * (lane + 32), lds_u8<32>(rawp) * 0xFFFF0001u + 0x01000100u);
            rawp += kBW;
        };
        for (int r = 0; r < 6; ++r) convert_row(ring0 + (uint32_t)r * (kRingPitch * 4));
        // masked: the mask byte under this lane's column of row y - 1 (the row nms_emit decides) is loaded unconditionally at the top of the
        // step and used ~100 instructions later -- one LDG and a pointer increment per row instead of a divergent branch around an on-demand
        // load in each of the nine inlined emitters (+16 % warp instructions).  Lanes right of the cell read its last column (never decided).
        const uint8_t* mrow = kMasked ? ml - lane + min(lane, iw - 1) - mpitch : nullptr;
        auto step = [&](auto jc, const int y) {
            constexpr int J = decltype(jc)::value;
            uint32_t mv = 1;
            if (kMasked) { mv = *mrow; mrow += (uint32_t)mpitch; }
            convert_row(ring0 + ((J + 6) & 7) * (kRingPitch * 4));                             // box row y + 6, the last one row y needs
            __syncwarp();
            const uint32_t v = lds_u8<((J + 3) & 7) * kRingPitch * 4>(rb);                    // low byte of the centre word
            uint32_t ring[16];
            load_ring_words<J>(rb, ring);
            const uint32_t r = fast_best_packed(ring) + v * 0xFFFFu;                           // halves: A + 256 - v, (255 - B) + 1 + v
            const uint32_t z = min(max(max(r & 0xFFFFu, r >> 16), zlow), zcap[0]);
            nms_emit(0, z, y - 1, mv);
        };
        for (int y0 = 0; y0 < ih; y0 += 8) {
#define ADB_STEP(J) if (y0 + J >= ih) break; step(std::integral_constant<int, J>{}, y0 + J);
            ADB_STEP(0) ADB_STEP(1) ADB_STEP(2) ADB_STEP(3) ADB_STEP(4) ADB_STEP(5) ADB_STEP(6) ADB_STEP(7)
#undef ADB_STEP
        }
    } else {
        for (int y = 0; y < ih; ++y) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (t > 0 && t >= nt) break;
                uint32_t ring[16];
                load_ring_packed_s<kBW>(cp[t], ring);
                const uint32_t r = fast_best_packed(ring);
                const uint32_t z = min(max(max(r & 0xFFFFu, r >> 16), zlow), zcap[t]);
                cp[t] += kBW;
                nms_emit(t, z, y - 1, 1u);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        if (t > 0 && t >= nt) break;
        nms_emit(t, zlow, ih - 1, (kMasked && NT == 1) ? (uint32_t)ml[(ptrdiff_t)(ih - 1) * mpitch + (min(lane, iw - 1) - lane)] : 1u);
    }
    const int cntA = (int)(idxA - slot0), cntB = (int)(slot0 + (uint32_t)L.slotcap - 1u - idxB);
    if (lane == 0) *cnt_out = (uint16_t)(cntA ? cntA : (cntB ? (cntB | 0x8000) : 0));
}

using FwKernel = void (*)(const TmaMaps16, const LevelDev*, const uint32_t*, const uint32_t*, int, const MaskPtrs, int, int, uint32_t*, int, uint16_t*, int, int, int);
static FwKernel fw_kernel(int box_w, int nt, bool masked) {
    if (box_w == 64) {
        if (nt == 1) return masked ? fast_cells_warp_kernel<64, 1, true> : fast_cells_warp_kernel<64, 1, false>;
        return masked ? fast_cells_warp_kernel<64, kFwTiles, true> : fast_cells_warp_kernel<64, kFwTiles, false>;
    }
    if (nt == 1) return masked ? fast_cells_warp_kernel<kCellBoxWMax, 1, true> : fast_cells_warp_kernel<kCellBoxWMax, 1, false>;
    return masked ? fast_cells_warp_kernel<kCellBoxWMax, kFwTiles, true> : fast_cells_warp_kernel<kCellBoxWMax, kFwTiles, false>;
}

// =========================================================================================
// Quad-tree distribution, one CTA per (level, frame).
//
// The reference keeps a std::list of nodes, pushes children to the front and erases parents
// (src/ORBextractor.cc:596-741).  Two facts make a flat restatement possible:
//   (1) the list is always sorted by node creation order, newest first (roots, pushed to the
//       back in ascending order, are the only exception and stay at the tail), so the output
//       order is "creation sequence descending";
//   (2) after every pass the nodes that can still be split are exactly the children created in
//       that pass with more than one key.
// So each pass only needs, for the current set of splittable ("active") nodes, the four child
// counts, the processing order (list order while the tree is far from the quota; (count, seq)
// descending with an early stop once the next pass could overshoot), and a prefix sum that
// hands out creation sequence numbers.  Keys never move: each carries (active slot, node seq).
// Tie-break of the reference's (count, pointer) sort = creation sequence.
constexpr int kQtThreads = 512;
constexpr uint32_t kQtFinal = 0xFFFu;
constexpr int kQtSeqWords = 1024;           // node creation sequence numbers are limited to 32768 per (frame, level)

struct QtShared {
    int na, nf, size, next_seq, mode, first, cur, done, kstar, tot_ch, tot_ex, ncand;
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(kQtThreads) quadtree_kernel(const LevelDev* __restrict__ levels, int nlevels,
                                                             const uint32_t* __restrict__ cand, int cand_total,
                                                             const uint16_t* __restrict__ cellcnt, int ncells_total,
                                                             uint32_t* __restrict__ qkeys, uint32_t* __restrict__ qstate,
                                                             int32_t* __restrict__ qcount, uint32_t* __restrict__ list,
                                                             int list_total, int32_t* __restrict__ listcnt, int maxa,
                                                             int32_t* __restrict__ status, int f0) {
    extern __shared__ __align__(16) uint8_t qt_smem[];
    // carve dynamic shared memory
    short4* bnd0 = reinterpret_cast<short4*>(qt_smem);            // [2][maxa] node bounds ulx,uly,brx,bry
    uint32_t* nseq0 = reinterpret_cast<uint32_t*>(bnd0 + 2 * maxa);   // [2][maxa]
    uint32_t* ncnt0 = nseq0 + 2 * maxa;                                // [2][maxa]
    uint32_t* child = ncnt0 + 2 * maxa;                                // [4*maxa] child counts, then child (slot,seq) map
    uint32_t* sc_ch = child + 4 * maxa;                                // [maxa] exclusive scans in processing order
    uint32_t* sc_ex = sc_ch + maxa;                                    // [maxa]
    uint32_t* fseq = sc_ex + maxa;                                     // [maxa] final node seqs / sorted
    uint32_t* best = fseq + maxa;                                      // [maxa]
    uint32_t* fsorted = best + maxa;                                   // [maxa]
    uint16_t* rank = reinterpret_cast<uint16_t*>(fsorted + maxa);     // [maxa] processing position of slot
    uint16_t* r_slot = rank + maxa;                                    // [maxa] inverse
    uint8_t* r_nch = reinterpret_cast<uint8_t*>(r_slot + maxa);        // [maxa] nonempty children, by rank
    uint8_t* r_nex = r_nch + maxa;                                     // [maxa] children with > 1 key, by rank
    __shared__ QtShared S;
    __shared__ uint32_t bm[kQtSeqWords];     // surviving node keys (final ordering)
    __shared__ uint16_t suf[kQtSeqWords];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int level = blockIdx.x, f = blockIdx.y + f0;
    const LevelDev& L = levels[level];
    const uint32_t* cslots = cand + (size_t)f * cand_total + L.cand_base;
    const uint16_t* ccnt = cellcnt + (size_t)f * ncells_total + L.cell_base;
    uint32_t* keys = qkeys + (size_t)f * cand_total + L.cand_base;
    uint32_t* state = qstate + (size_t)f * cand_total + L.cand_base;
    uint32_t* out = list + (size_t)f * list_total + L.list_base;
    const int N = L.quota, nIni = L.n_ini;

    // ---- gather the cells' candidates in the reference's order (cell row-major, then pixel row-major)
    // exclusive scan of the cell counts by warp 0 into `state` scratch (reused afterwards)
    if (warp == 0) {
        int run = 0;
        for (int c0 = 0; c0 < L.ncells; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < L.ncells ? (int)(ccnt[c] & 0x7FFF) : 0;
            const int inc = warp_incl_scan(v, lane);
            if (c < L.ncells) state[c] = (uint32_t)(run + inc - v);   // cell offsets live in state[0..ncells)
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) S.ncand = run;
    }
    __syncthreads();
    const int ncand = S.ncand;
    if (tid == 0) qcount[f * nlevels + level] = ncand;
    if (ncand == 0 || nIni <= 0) {
        if (tid == 0) listcnt[f * nlevels + level] = 0;
        return;
    }
    // cell offsets live in state[0..ncells) (ncells <= cand_cap); the gather writes keys[] only
    for (int c = warp; c < L.ncells; c += kQtThreads / 32) {
        const int cc = ccnt[c], n = cc & 0x7FFF;          // bit 15: no corner at iniTh, the minTh corners sit at the back of the slot, reversed
        const uint32_t o = state[c];
        const uint32_t* src = cslots + (size_t)c * L.slotcap;
        if (cc & 0x8000) for (int s = lane; s < n; s += 32) keys[o + s] = src[L.slotcap - 1 - s];
        else for (int s = lane; s < n; s += 32) keys[o + s] = src[s];
    }
    __syncthreads();

    // ---- roots (src/ORBextractor.cc:547-590)
    auto bnd = [&](int b) { return bnd0 + b * maxa; };
    auto nseq = [&](int b) { return nseq0 + b * maxa; };
    auto ncnt = [&](int b) { return ncnt0 + b * maxa; };
    const float hx = L.hx;
    const int bry_root = (L.h - kMinBorder) - kMinBorder;
    for (int i = tid; i < nIni; i += kQtThreads) child[i] = 0;
    __syncthreads();
    if (nIni == 1) {
        if (tid == 0) child[0] = (uint32_t)ncand;
        for (int k = tid; k < ncand; k += kQtThreads) state[k] = 0;   // root index, fixed up below
    } else {
        for (int k = tid; k < ncand; k += kQtThreads) {
            const int x = keys[k] & 0xFFF;
            int r = (int)__fdiv_rn((float)x, hx);
            r = min(r, nIni - 1);
            atomicAdd(&child[r], 1u);
            state[k] = (uint32_t)r;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int na = 0, nf = 0, size = 0;
        for (int i = 0; i < nIni; ++i) {
            const uint32_t c = child[i];
            if (c == 0) { child[i] = 0xFFFFFFFFu; continue; }
            ++size;
            if (c == 1) {
                fseq[nf++] = (uint32_t)i;
                child[i] = (kQtFinal << 20) | (uint32_t)i;
            } else {
                bnd(0)[na] = make_short4((short)(int)__fmul_rn(hx, (float)i), 0, (short)(int)__fmul_rn(hx, (float)(i + 1)), (short)bry_root);
                nseq(0)[na] = (uint32_t)i;
                ncnt(0)[na] = c;
                child[i] = ((uint32_t)na << 20) | (uint32_t)i;
                ++na;
            }
        }
        S.na = na; S.nf = nf; S.size = size; S.next_seq = nIni; S.mode = 0; S.first = 1; S.cur = 0; S.done = 0;
    }
    __syncthreads();
    for (int k = tid; k < ncand; k += kQtThreads) state[k] = child[state[k]];
    __syncthreads();

    // ---- passes
    int guard = 0;
    while (true) {
        const int na = S.na, cur = S.cur, mode = S.mode, first = S.first, size = S.size;
        if (na == 0 || S.done) break;
        if (++guard > 64) {   // cannot happen for distinct integer pixels; never loop forever on the device
            if (tid == 0) atomicExch(status, 2);
            break;
        }
        const short4* B = bnd(cur);
        for (int i = tid; i < 4 * na; i += kQtThreads) child[i] = 0;
        __syncthreads();
        // child counts
        for (int k = tid; k < ncand; k += kQtThreads) {
            const uint32_t st = state[k];
            const uint32_t s = st >> 20;
            if (s == kQtFinal) continue;
            const uint32_t key = keys[k];
            const int x = key & 0xFFF, y = (key >> 12) & 0xFFF;
            const short4 b = B[s];
            const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
            const int q = (x < mx) ? (y < my ? 0 : 2) : (y < my ? 1 : 3);
            atomicAdd(&child[4 * s + q], 1u);
        }
        __syncthreads();
        // processing order
        for (int s = tid; s < na; s += kQtThreads) {
            int r;
            if (mode == 0) {
                r = first ? s : na - 1 - s;
            } else {
                const uint32_t c = ncnt(cur)[s], q = nseq(cur)[s];
                r = 0;
                for (int t = 0; t < na; ++t) {
                    const uint32_t ct = ncnt(cur)[t], qt = nseq(cur)[t];
                    r += (ct > c) || (ct == c && qt > q);
                }
            }
            rank[s] = (uint16_t)r;
            r_slot[r] = (uint16_t)s;
            int nch = 0, nex = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { const uint32_t c = child[4 * s + q]; nch += c > 0; nex += c > 1; }
            r_nch[r] = (uint8_t)nch;
            r_nex[r] = (uint8_t)nex;
        }
        __syncthreads();
        // scans in processing order + stop position (warp 0)
        if (warp == 0) {
            int run_ch = 0, run_ex = 0, kstar = na - 1, found = 0;
            for (int r0 = 0; r0 < na; r0 += 32) {
                const int r = r0 + lane;
                const int vch = r < na ? (int)r_nch[r] : 0, vex = r < na ? (int)r_nex[r] : 0;
                const int ich = warp_incl_scan(vch, lane), iex = warp_incl_scan(vex, lane);
                if (r < na) { sc_ch[r] = (uint32_t)(run_ch + ich - vch); sc_ex[r] = (uint32_t)(run_ex + iex - vex); }
                if (mode == 1 && !found) {
                    const bool hit = r < na && (size + run_ch + ich - (r + 1) >= N);
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
                    if (bal) { kstar = r0 + __ffs(bal) - 1; found = 1; }
                }
                run_ch += __shfl_sync(0xFFFFFFFFu, ich, 31);
                run_ex += __shfl_sync(0xFFFFFFFFu, iex, 31);
            }
            if (lane == 0) S.kstar = kstar;
        }
        __syncthreads();
        const int kstar = S.kstar;
        if (tid == 0) {
            S.tot_ch = (int)sc_ch[kstar] + r_nch[kstar];
            S.tot_ex = (int)sc_ex[kstar] + r_nex[kstar];
        }
        const int next_seq = S.next_seq;
        // create children
        const int nxt = cur ^ 1;
        for (int s = tid; s < na; s += kQtThreads) {
            const int r = rank[s];
            if (r > kstar) {   // left unsplit by the early stop: stays in the list as it is
                const int fi = atomicAdd(&S.nf, 1);
                fseq[fi] = nseq(cur)[s];
                continue;
            }
            const short4 b = B[s];
            const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
            uint32_t cseq = (uint32_t)next_seq + sc_ch[r];
            uint32_t nslot = sc_ex[r];
            uint32_t c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = child[4 * s + q];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (c[q] == 0) { child[4 * s + q] = 0; continue; }
                if (c[q] > 1) {
                    short4 nb;
                    nb.x = (q & 1) ? (short)mx : b.x;
                    nb.z = (q & 1) ? b.z : (short)mx;
                    nb.y = (q & 2) ? (short)my : b.y;
                    nb.w = (q & 2) ? b.w : (short)my;
                    bnd(nxt)[nslot] = nb;
                    nseq(nxt)[nslot] = cseq;
                    ncnt(nxt)[nslot] = c[q];
                    child[4 * s + q] = (nslot << 20) | cseq;
                    ++nslot;
                } else {
                    const int fi = atomicAdd(&S.nf, 1);
                    fseq[fi] = cseq;
                    child[4 * s + q] = (kQtFinal << 20) | cseq;
                }
                ++cseq;
            }
        }
        __syncthreads();
        // move the keys
        for (int k = tid; k < ncand; k += kQtThreads) {
            const uint32_t st = state[k];
            const uint32_t s = st >> 20;
            if (s == kQtFinal) continue;
            if ((int)rank[s] > kstar) {
                state[k] = (kQtFinal << 20) | (st & 0xFFFFFu);
                continue;
            }
            const uint32_t key = keys[k];
            const int x = key & 0xFFF, y = (key >> 12) & 0xFFF;
            const short4 b = B[s];
            const int mx = b.x + ((b.z - b.x + 1) >> 1), my = b.y + ((b.w - b.y + 1) >> 1);
            const int q = (x < mx) ? (y < my ? 0 : 2) : (y < my ? 1 : 3);
            state[k] = child[4 * s + q];
        }
        __syncthreads();
        if (tid == 0) {
            const int nsplit = kstar + 1;
            const int new_size = size - nsplit + S.tot_ch;
            const int n_exp = S.tot_ex;
            S.next_seq = next_seq + S.tot_ch;
            S.size = new_size;
            S.na = n_exp;
            S.cur = nxt;
            S.first = 0;
            if (new_size >= N || new_size == size) S.done = 1;
            else if (mode == 0 && new_size + n_exp * 3 > N) S.mode = 1;
            if (S.nf + n_exp > maxa || S.next_seq >= 32 * kQtSeqWords) { atomicExch(status, 3); S.done = 1; S.na = 0; }
        }
        __syncthreads();
    }
    __syncthreads();
    // remaining splittable nodes stay in the list
    {
        const int na = S.na, cur = S.cur;
        const int nf = S.nf + na;
        for (int s = tid; s < na; s += kQtThreads) fseq[nf - na + s] = nseq(cur)[s];
        __syncthreads();
        // ---- order = list order: creation sequence descending, roots (ascending) at the tail.  Rank of a node =
        // number of surviving nodes with a larger key: a bitmap over the key space + suffix popcounts make it O(1) per
        // node and per candidate (no sort, no search).
        for (int w = tid; w < kQtSeqWords; w += kQtThreads) bm[w] = 0;
        for (int r = tid; r < nf; r += kQtThreads) best[r] = 0;
        __syncthreads();
        for (int i = tid; i < nf; i += kQtThreads) {
            const uint32_t qi = fseq[i];
            const uint32_t ki = qi >= (uint32_t)nIni ? qi : (uint32_t)(nIni - 1) - qi;
            atomicOr(&bm[ki >> 5], 1u << (ki & 31));
        }
        __syncthreads();
        if (warp == 0) {   // suf[w] = set bits in words above w
            int run = 0;
            for (int w0 = kQtSeqWords - 32; w0 >= 0; w0 -= 32) {
                const int w = w0 + (31 - lane);              // lane 0 takes the highest word of the chunk
                const int v = __popc(bm[w]);
                const int inc = warp_incl_scan(v, lane);
                suf[w] = (uint16_t)(run + inc - v);
                run += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
        }
        __syncthreads();
        // max response per node, first key in candidate order wins ties (src/ORBextractor.cc:745-762)
        for (int k = tid; k < ncand; k += kQtThreads) {
            const uint32_t q = state[k] & 0xFFFFFu;
            const uint32_t kq = q >= (uint32_t)nIni ? q : (uint32_t)(nIni - 1) - q;
            const uint32_t word = bm[kq >> 5], bit = kq & 31;
            const int r = suf[kq >> 5] + (bit == 31 ? 0 : __popc(word >> (bit + 1)));
            atomicMax(&best[r], ((keys[k] >> 24) << 24) | (0xFFFFFFu - (uint32_t)k));
        }
        __syncthreads();
        const int nout = min(nf, L.list_cap);
        for (int r = tid; r < nout; r += kQtThreads) {
            const uint32_t k = 0xFFFFFFu - (best[r] & 0xFFFFFFu);
            const uint32_t key = keys[k];
            const uint32_t x = (key & 0xFFF) + kMinBorder, y = ((key >> 12) & 0xFFF) + kMinBorder;
            out[r] = x | (y << 12) | (key & 0xFF000000u);
        }
        if (tid == 0) {
            listcnt[f * nlevels + level] = nout;
            if (nf > L.list_cap) atomicExch(status, 4);
        }
    }
}

// =========================================================================================
// Orientation + descriptor, one warp per key-point (orient_describe_kernel, below the level blur it reads).  Two 64-byte-wide boxes
// around the key-point arrive by TMA on one mbarrier: 31 rows of the level (the r = 15 disc of IC_Angle) and 37 rows of the blurred
// level (the 256 rotated tests reach +-18 px); both lie inside the level because key-points sit >= 19 px from its edges.
// =========================================================================================
// GaussianBlur(7 x 7, sigma 2, BORDER_REFLECT_101) of a whole pyramid level (src/ORBextractor.cc:1099-1100: the reference blurs a clone
// of the level ROI once per level).  Fixed point {18,34,48,56,48,34,18} / 256 on both axes, (v + 2^15) >> 16: integer and exact, so
// the pass order is free.  Round 1 blurred a 37 x 37 core per key-point inside the descriptor kernel: 1200 of its 2065
// warp-instructions per key-point, 2.4 M per frame; the whole level costs 0.4 M per frame and the descriptor kernel fetches the
// blurred patch with a second TMA box load.
// One WARP per (120-pixel column group, 32-row strip), no shared memory and no block barrier: lane = one aligned word (4 pixels) of the
// row, lanes 0 and 31 are the halo words of the group.  A lane walks down its column: one coalesced word load per row slides through
// a 7-row register window (even / odd pixels as two 16-bit lanes each: two pixels per multiply, 7 x 255 x 56 < 2^16; the kernel is
// symmetric: 3 adds + 4 multiply-adds per lane pair), the four 16-bit column sums of the word go to the neighbours by shuffles, the
// row pass is DP2A on the packed sums, one word store per lane and row.  BORDER_REFLECT_101: rows by the row index; the left edge on
// the column SUMS (the column pass is per column, so S(-k) = S(k)); words that reach over the right edge are assembled byte-wise.
constexpr int kBlRows = 42, kBlGroupW = 30;   // 42 output rows = 6 groups of 7 (the window rotates through 7 register slots)
__device__ __forceinline__ int reflect101_any(int p, int len) {   // halo lanes may hang far over the right edge of a small level
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}
// kInterior: every input row of the strip lies inside the level and the lane's word is a full aligned word -> rows are pitch steps.
// The row loop runs in groups of 7 (the window rotates through 7 register slots, so inside a group every slot index is static);
// the 7 loads of the next group are in flight while a group is computed.
template <bool kInterior>
__device__ __forceinline__ void blur7_strip(const uint8_t* __restrict__ sp, int spitch, int w, int h, int y0, int nrows, int wi, int lw,
                                            bool bytewise, uint8_t* __restrict__ dp, int dpitch, bool store) {
    const uint32_t kw0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), kw1 = 48u | (34u << 8) | (18u << 16);
    int bx[4];                                             // byte-wise path: the four (reflected) columns of this lane's word
#pragma unroll
    for (int k = 0; k < 4; ++k) bx[k] = kInterior ? 0 : reflect101_any(min(4 * lw + k, w + 6), w);
    const int last_in = nrows + 5;                         // last input row the strip needs
    auto load_row = [&](int i) -> uint32_t {               // input row i of the strip (row y0 - 3 + i of the level, reflected)
        int yy = y0 - 3 + min(i, last_in);
        if (!kInterior) {
            yy = yy < 0 ? -yy : (yy >= h ? 2 * (h - 1) - yy : yy);
            yy = min(max(yy, 0), h - 1);                   // (levels of fewer than 4 rows cannot occur; keeps the load in bounds)
        }
        const uint8_t* rowp = sp + (size_t)yy * spitch;
        if (kInterior || !bytewise) return __ldg(reinterpret_cast<const uint32_t*>(rowp) + lw);
        return (uint32_t)__ldg(rowp + bx[0]) | ((uint32_t)__ldg(rowp + bx[1]) << 8) | ((uint32_t)__ldg(rowp + bx[2]) << 16) | ((uint32_t)__ldg(rowp + bx[3]) << 24);
    };
    uint32_t we[7], wo[7], pre[7];
#pragma unroll
    for (int j = 0; j < 6; ++j) {                          // rows 0 .. 5 fill slots 0 .. 5
        const uint32_t wv = load_row(j);
        we[j] = wv & 0x00FF00FFu; wo[j] = (wv >> 8) & 0x00FF00FFu;
    }
    we[6] = 0; wo[6] = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) pre[j] = load_row(6 + j);
#pragma unroll 1
    for (int g0 = 0; g0 < nrows; g0 += 7) {                // output rows g0 .. g0 + 6 from input rows g0 + 6 .. g0 + 12
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            constexpr int kNone = 0;
            const int sl = (6 + j) % 7 + kNone;            // slot of the newest row; slot (sl - k) mod 7 = the row k above it
            const uint32_t wv = pre[j];
            pre[j] = load_row(g0 + 13 + j);
            we[sl] = wv & 0x00FF00FFu; wo[sl] = (wv >> 8) & 0x00FF00FFu;
            const uint32_t lo = 18u * (we[(sl + 1) % 7] + we[sl]) + 34u * (we[(sl + 2) % 7] + we[(sl + 6) % 7]) + 48u * (we[(sl + 3) % 7] + we[(sl + 5) % 7]) +
                                56u * we[(sl + 4) % 7];   // column sums of pixels 0, 2
            const uint32_t hi = 18u * (wo[(sl + 1) % 7] + wo[sl]) + 34u * (wo[(sl + 2) % 7] + wo[(sl + 6) % 7]) + 48u * (wo[(sl + 3) % 7] + wo[(sl + 5) % 7]) +
                                56u * wo[(sl + 4) % 7];   // pixels 1, 3
            const uint32_t w2 = __byte_perm(lo, hi, 0x5410), w3 = __byte_perm(lo, hi, 0x7632);   // [S0, S1], [S2, S3]
            uint32_t w0 = __shfl_up_sync(0xFFFFFFFFu, w2, 1), w1 = __shfl_up_sync(0xFFFFFFFFu, w3, 1);
            const uint32_t w4 = __shfl_down_sync(0xFFFFFFFFu, w2, 1), w5 = __shfl_down_sync(0xFFFFFFFFu, w3, 1);
            if (wi == 0) { w0 = w3; w1 = __byte_perm(w2, w3, 0x3254); }      // S(-3) = S3 (upper half), [S(-2), S(-1)] = [S2, S1]
            const uint32_t s0 = __funnelshift_r(w0, w1, 16), s1 = __funnelshift_r(w1, w2, 16), s2 = __funnelshift_r(w2, w3, 16),
                           s3 = __funnelshift_r(w3, w4, 16), s4 = __funnelshift_r(w4, w5, 16);
            const uint32_t v0 = __dp2a_hi(s3, kw1, __dp2a_lo(s2, kw1, __dp2a_hi(s1, kw0, __dp2a_lo(s0, kw0, 32768u))));
            const uint32_t v1 = __dp2a_hi(w4, kw1, __dp2a_lo(w3, kw1, __dp2a_hi(w2, kw0, __dp2a_lo(w1, kw0, 32768u))));
            const uint32_t v2 = __dp2a_hi(s4, kw1, __dp2a_lo(s3, kw1, __dp2a_hi(s2, kw0, __dp2a_lo(s1, kw0, 32768u))));
            const uint32_t v3 = __dp2a_hi(w5, kw1, __dp2a_lo(w4, kw1, __dp2a_hi(w3, kw0, __dp2a_lo(w2, kw0, 32768u))));
            // (v + 2^15) >> 16 = byte 2 of each sum (< 2^24): two PRMTs pick them
            const uint32_t packed = __byte_perm(__byte_perm(v0, v1, 0x0062), __byte_perm(v2, v3, 0x0062), 0x5410);
            const int yo = g0 + j;
            if (store && yo < nrows) *reinterpret_cast<uint32_t*>(dp + (size_t)yo * dpitch) = packed;   // the pitch is padded to 16 B: a ragged last word fits
        }
    }
}

__global__ void __launch_bounds__(128, 8) blur7_level_kernel(const __grid_constant__ BlurLevels B, int f0) {
    const int lane = threadIdx.x & 31;
    int task = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (task >= B.ntasks) return;
    int l = 0;
    while (task >= B.task_end[l]) ++l;
    if (l) task -= B.task_end[l - 1];
    const int w = B.w[l], h = B.h[l], spitch = B.spitch[l], dpitch = B.dpitch[l], ncg = B.ncg[l];
    const int strip = task / ncg, g = task - strip * ncg, f = blockIdx.y + f0;
    const int wi = g * kBlGroupW - 1 + lane;               // this lane's word: pixels 4 wi .. 4 wi + 3
    const int y0 = strip * kBlRows, nrows = min(kBlRows, h - y0);
    const uint8_t* sp = B.src[l] + (size_t)f * B.sfstride[l];
    const bool aligned = (((uintptr_t)sp | (uintptr_t)spitch) & 3) == 0;
    const bool bytewise = wi >= (w >> 2) || !aligned;      // right edge (reflected bytes) or a caller buffer without word alignment
    const int lw = max(wi, 0);
    uint8_t* dp = B.dst[l] + (size_t)f * B.dfstride[l] + (size_t)y0 * dpitch + 4 * lw;   // destination: the handle's own buffer, word aligned
    const bool store = lane >= 1 && lane <= kBlGroupW && 4 * wi < w;
    const bool interior = y0 >= 3 && y0 + kBlRows + 3 <= h && aligned && (g + 1) * kBlGroupW + 1 <= (w >> 2);   // warp-uniform
    if (interior) blur7_strip<true>(sp, spitch, w, h, y0, nrows, wi, lw, false, dp, dpitch, store);
    else blur7_strip<false>(sp, spitch, w, h, y0, nrows, wi, lw, bytewise, dp, dpitch, store);
}

constexpr int kDescWarps = 8;
constexpr int kIcBoxH = 31, kIcR = 15;   // level box for IC_Angle: the r = 15 disc, 31 rows x 64 B
constexpr int kBlBoxH = 37, kBlR = 18;   // blurred box for rBRIEF: samples reach +-18 px, 37 rows x 64 B
constexpr int kRawBytes = 2048;          // 31 x 64 B padded to a multiple of 128 B
constexpr int kBlrBytes = 2432;          // 37 x 64 B padded to a multiple of 128 B
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2 in degrees, every operation rounded to float32, no contraction (oracle/orb_oracle.cpp:153-177)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
    const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
    const float eps = (float)DBL_EPSILON;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        float t = __fmul_rn(p7, c2); t = __fadd_rn(t, p5); t = __fmul_rn(t, c2); t = __fadd_rn(t, p3);
        t = __fmul_rn(t, c2); t = __fadd_rn(t, p1);
        a = __fmul_rn(t, c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        float t = __fmul_rn(p7, c2); t = __fadd_rn(t, p5); t = __fmul_rn(t, c2); t = __fadd_rn(t, p3);
        t = __fmul_rn(t, c2); t = __fadd_rn(t, p1);
        t = __fmul_rn(t, c);
        a = __fsub_rn(90.f, t);
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ int reflect101(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
    return p;
}

struct DescSmem {
    uint8_t raw[kDescWarps][kRawBytes];               // 31 x 64 B box of the level around the key-point: the r = 15 disc of IC_Angle
    uint8_t blr[kDescWarps][kBlrBytes];               // 37 x 64 B box of the BLURRED level (blur7_level_kernel): rBRIEF samples +-18 px
    uint32_t ic_wu[16][8];                            // IC_Angle row tables by |v|: signed weights u of word k (0 outside the disc)
    uint32_t ic_m[16][8];                             // ... and the disc mask as 0 / 1 bytes
    uint64_t bar[kDescWarps];
    uint4 gdesc[kDescWarps * 2];                      // fused gather: the CTA's 8 descriptors (256 B) ...
    uint4 gkp[kDescWarps * 6 / 4];                    // ... and 8 key-point records (192 B), sent as 16-byte stores by warp 0
};

__device__ __forceinline__ int dp4a_u8_s8(uint32_t a, uint32_t b, int c) {   // sum of unsigned bytes of a times signed bytes of b
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// one 32-bit store of the fused all-gather: a plain store to a peer-mapped address, or multimem.st to an NVLS multicast address
__device__ __forceinline__ void gather_store_u128(void* p, uint4 v, int multicast) {
    if (multicast)
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                     "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
    else *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void gather_store_u32(uint32_t* p, uint32_t v, int multicast) {
    if (multicast) asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else *p = v;
}

__global__ void __launch_bounds__(kDescWarps * 32) orient_describe_kernel(const __grid_constant__ TmaMaps16 maps, const __grid_constant__ TmaMaps16 bmaps,
                                                                         const LevelDev* __restrict__ levels, int nlevels,
                                                                         const uint32_t* __restrict__ list, int list_total,
                                                                         const int32_t* __restrict__ listcnt,
                                                                         const float4* __restrict__ pattern,
                                                                         adb_keypoint* __restrict__ kps,
                                                                         uint8_t* __restrict__ desc,
                                                                         int32_t* __restrict__ counts, int cap,
                                                                         const __grid_constant__ adb_gather_targets gather, int f0) {
    extern __shared__ __align__(128) uint8_t desc_smem_raw[];
    // 128-B alignment for the TMA destinations; the offset is added to the shared array itself (not to a uintptr_t) so the
    // compiler keeps the shared address space: LDS / STS with 32-bit addresses instead of generic LD / ST
    DescSmem& sm = *reinterpret_cast<DescSmem*>(desc_smem_raw + ((128u - (smem_u32(desc_smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, f = blockIdx.y + f0;
    if (tid < 128) {   // one table entry per thread: word k of rows +-av covers u = -15 + 4k .. -12 + 4k
        const int av = tid >> 3, k = tid & 7;
        const int lim = c_umax[av];
        uint32_t wu = 0, m = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int u = -15 + 4 * k + b;
            if (u <= 15 && abs(u) <= lim) { wu |= (uint32_t)(u & 0xFF) << (8 * b); m |= 1u << (8 * b); }
        }
        sm.ic_wu[av][k] = wu; sm.ic_m[av][k] = m;
    }
    if (lane == 0) mbar_init(&sm.bar[warp], 1);
    if (tid == 0) mbar_fence_init();
    __syncthreads();

    const int i = blockIdx.x * kDescWarps + warp;
    // level of key-point i: the per-level counts of the frame, one per lane, prefix-summed by shuffles
    int base, lvl, total;
    {
        const int c = lane < nlevels ? listcnt[f * nlevels + lane] : 0;
        const int incl = warp_incl_scan(c, lane);
        const uint32_t below = __ballot_sync(0xFFFFFFFFu, lane < nlevels && i >= incl);   // levels that end at or before i (a prefix of the lanes)
        lvl = __popc(below);
        base = lvl ? __shfl_sync(0xFFFFFFFFu, incl, (lvl - 1) & 31) : 0;
        total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (lvl >= nlevels) lvl = -1;
    }
    if (i == 0 && lane == 0) {
        counts[f] = total;
        for (int g = 0; g < gather.n; ++g) gather_store_u32(reinterpret_cast<uint32_t*>(gather.counts[g] + f), (uint32_t)total, gather.multicast);   // over NVLink
    }
    // fused gather in 16-byte stores: the CTA's 8 consecutive records are staged in shared memory and leave from warp 0 (rows of the
    // result arrays must then start 16-byte aligned: capacity a multiple of 8; else the word-granular path below)
    const bool gvec = gather.n > 0 && (cap & 7) == 0;
    if (lvl < 0 || i >= cap) {
        if (gvec) asm volatile("bar.sync 1, %0;" ::"n"(kDescWarps * 32));   // meets the staging barrier of the CTA's working warps
        return;
    }
    const LevelDev& L = levels[lvl];
    const uint32_t e = list[(size_t)f * list_total + L.list_base + (i - base)];
    const int cx = e & 0xFFF, cy = (e >> 12) & 0xFFF, resp = e >> 24;

    uint8_t* raw0 = sm.raw[warp];
    uint8_t* blr0 = sm.blr[warp];
    uint64_t* bar = &sm.bar[warp];
    const int x0 = cx - kPatchR, y0 = cy - kPatchR;
    if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, kPatchBoxW * (kIcBoxH + kBlBoxH));
        tma_load_3d(raw0, &maps.m[lvl], bar, x0 & ~15, cy - kIcR, f);    // 16-B aligned box origin; both boxes lie inside the level:
        tma_load_3d(blr0, &bmaps.m[lvl], bar, x0 & ~15, cy - kBlR, f);   // key-points sit >= 19 px from its edges
    }
    mbar_wait(bar, 0);
    const int dxp = x0 & 15;

    // ---- IC_Angle (src/ORBextractor.cc:78-105): m10 = sum u*I, m01 = sum v*I over the r=15 disc
    // a lane owns word k of four rows at a time (8 lanes per row: the 64-B row pitch would put all rows of a lane-per-row
    // mapping on two banks); the funnel shift lines the word up with u = -15 + 4k, and two DP4A give its share of
    // sum(u * I) (signed weights, zero outside the disc) and of sum(I) (0 / 1 mask); m01 = sum over rows of v * sum(I).
    int m10 = 0, m01 = 0;
    {
        const int k = lane & 7, rsub = lane >> 3;
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw0) + ((dxp + 6) >> 2) + k;   // row 0 of the box is v = -15
        const uint32_t s8 = (uint32_t)((dxp + 6) & 3) * 8u;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int row = 4 * it + rsub;          // v + 15
            if (row < 31) {
                const uint32_t d = __funnelshift_r(rw[row * (kPatchBoxW / 4)], rw[row * (kPatchBoxW / 4) + 1], s8);
                const int av = abs(row - 15);
                m10 = dp4a_u8_s8(d, sm.ic_wu[av][k], m10);
                m01 += (row - 15) * (int)__dp4a(d, sm.ic_m[av][k], 0u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
        m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered rBRIEF (src/ORBextractor.cc:109-148); lane = descriptor byte
    const float ang = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    double sd, cd;
    sincos((double)ang, &sd, &cd);
    const float a = (float)cd, b = (float)sd;
    // Rounding without the conversion unit: x + 1.5 * 2^23 has ulp 1, so the float addition IS cvRound (nearest, ties to even) and the
    // integer sits in the low mantissa bits: bits = 0x4B400000 + round(x), negative x included.  Row * 64 + column is then formed on the
    // raw bit patterns and the two biases are folded into the base address (F2I runs at a quarter of the FADD rate).
    constexpr float kMagic = 12582912.f;
    const uint32_t cbase = smem_u32(blr0) + (uint32_t)(dxp + kBlR * kPatchBoxW + kPatchR) - 65u * 0x4B400000u;   // the key-point inside the blurred box
    uint32_t byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 pt = __ldg(pattern + k * 32 + lane);   // 4 KB, L1 resident
        const float x0f = pt.x, y0f = pt.y, x1f = pt.z, y1f = pt.w;
        const uint32_t r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(x0f, b), __fmul_rn(y0f, a)), kMagic));
        const uint32_t q0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(x0f, a), __fmul_rn(y0f, b)), kMagic));
        const uint32_t r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(x1f, b), __fmul_rn(y1f, a)), kMagic));
        const uint32_t q1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(x1f, a), __fmul_rn(y1f, b)), kMagic));
        const uint32_t t0 = lds_u8<0>(cbase + r0 * (uint32_t)kPatchBoxW + q0), t1 = lds_u8<0>(cbase + r1 * (uint32_t)kPatchBoxW + q1);
        byte |= (uint32_t)(t0 < t1) << k;
    }
    desc[((size_t)f * cap + i) * 32 + lane] = (uint8_t)byte;
    adb_keypoint kp;
    kp.x = lvl ? __fmul_rn((float)cx, L.scale) : (float)cx;
    kp.y = lvl ? __fmul_rn((float)cy, L.scale) : (float)cy;
    kp.size = (float)L.patch_size;
    kp.angle = angle;
    kp.response = (float)resp;
    kp.octave = lvl;
    if (lane == 0) kps[(size_t)f * cap + i] = kp;
    // fused all-gather: the record also goes to every peer's buffer (or once to the NVLS multicast mapping, which the switch
    // replicates).  Fallback for odd capacities: 32-bit words -- lanes 0..7 one word of the descriptor each, lanes 0..5 one word of
    // the key-point (multimem.st has no byte form)
    if (gvec) {
        reinterpret_cast<uint8_t*>(sm.gdesc)[warp * 32 + lane] = (uint8_t)byte;
        if (lane < 6)
            reinterpret_cast<uint32_t*>(sm.gkp)[warp * 6 + lane] = lane == 0 ? __float_as_uint(kp.x) : lane == 1 ? __float_as_uint(kp.y)
                                                                   : lane == 2 ? __float_as_uint(kp.size) : lane == 3 ? __float_as_uint(kp.angle)
                                                                   : lane == 4 ? __float_as_uint(kp.response) : (uint32_t)kp.octave;
        asm volatile("bar.sync 1, %0;" ::"n"(kDescWarps * 32));
        if (warp == 0 && lane < 28) {            // warp 0 holds the CTA's first record, so it is here whenever any warp of the CTA is
            const int nrec = min(min(total, cap) - (int)blockIdx.x * kDescWarps, kDescWarps);   // records of this CTA that exist
            const size_t rec0 = (size_t)f * cap + (size_t)blockIdx.x * kDescWarps;
            const bool isd = lane < 16;
            const int chunk = isd ? lane : lane - 16;                                         // 16-byte chunk of the 256 / 192 bytes
            const bool live = (isd ? chunk / 2 : chunk * 16 / 24) < nrec;                      // the record its first byte belongs to
            const uint4 v = isd ? sm.gdesc[chunk] : sm.gkp[chunk];
            if (live)
                for (int g = 0; g < gather.n; ++g)
                    gather_store_u128(isd ? (void*)(gather.desc[g] + rec0 * 32 + chunk * 16) : (void*)((uint8_t*)gather.kps[g] + rec0 * 24 + chunk * 16), v,
                                      gather.multicast);
        }
    } else if (gather.n > 0) {
        uint32_t w = byte | (__shfl_down_sync(0xFFFFFFFFu, byte, 1) << 8) | (__shfl_down_sync(0xFFFFFFFFu, byte, 2) << 16) |
                     (__shfl_down_sync(0xFFFFFFFFu, byte, 3) << 24);                // valid in lanes 0, 4, 8, ...
        w = __shfl_sync(0xFFFFFFFFu, w, (lane & 7) * 4);                           // lanes 0..7: word `lane` of the descriptor
        const uint32_t kw = lane == 0 ? __float_as_uint(kp.x) : lane == 1 ? __float_as_uint(kp.y) : lane == 2 ? __float_as_uint(kp.size)
                          : lane == 3 ? __float_as_uint(kp.angle) : lane == 4 ? __float_as_uint(kp.response) : (uint32_t)kp.octave;
        const size_t rec = (size_t)f * cap + i;
        for (int g = 0; g < gather.n; ++g) {
            if (lane < 8) gather_store_u32(reinterpret_cast<uint32_t*>(gather.desc[g]) + rec * 8 + lane, w, gather.multicast);
            if (lane < 6) gather_store_u32(reinterpret_cast<uint32_t*>(gather.kps[g]) + rec * 6 + lane, kw, gather.multicast);
        }
    }
}

// Debug: candidates of (frame, level) in reference order -> int32 triples.
__global__ void unpack_candidates_kernel(const uint32_t* __restrict__ keys, int n, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    out[3 * i] = k & 0xFFF; out[3 * i + 1] = (k >> 12) & 0xFFF; out[3 * i + 2] = k >> 24;
}

// =========================================================================================
// Host side.
static inline int round_even_f(float v) { return (int)nearbyintf(v); }

// oracle/orb_oracle.cpp:96-110 (cv::resize coefficient rule)
static void linear_coeffs(int src, int dst, std::vector<int2>& c) {
    c.resize(dst);
    const double scale = (double)src / dst;
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        const int a0 = round_even_f((1.f - f) * 2048.f), a1 = round_even_f(f * 2048.f);
        c[d] = make_int2(s, a0 | (a1 << 16));
    }
}

static size_t qt_smem_bytes(int maxa) {
    return (size_t)maxa * (2 * sizeof(short4) + 2 * 4 + 2 * 4 + 4 * 4 + 4 + 4 + 4 + 4 + 4 + 2 + 2 + 1 + 1) + 64;
}

static void release_resources(adb_orb* h) {
    for (auto& l : h->lv) {
        cudaFree(l.img); cudaFree(l.mask); cudaFree(l.blur); cudaFree(l.xtab); cudaFree(l.ytab);
    }
    cudaFree(h->mask_stage);
    if (h->copy_stream) {
        cudaStreamDestroy(h->copy_stream); cudaStreamDestroy(h->d2h_stream);
        for (auto& e : h->cev) if (e) cudaEventDestroy(e);
        for (auto& e : h->sev) if (e) cudaEventDestroy(e);
        if (h->stereo_stream) cudaStreamDestroy(h->stereo_stream);
    }
    cudaFree(h->d_levels); cudaFree(h->d_cell_table); cudaFree(h->d_pattern); cudaFree(h->d_cand); cudaFree(h->d_cellcnt);
    cudaFree(h->d_qkeys); cudaFree(h->d_qstate); cudaFree(h->d_qcount); cudaFree(h->d_list); cudaFree(h->d_listcnt);
    cudaFree(h->d_status); cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_counts);
    cudaFree(h->d_uright); cudaFree(h->d_depth); cudaFree(h->d_best_idx); cudaFree(h->d_best_dist); cudaFree(h->d_sad);
    cudaFree(h->d_row_ptr); cudaFree(h->d_row_items); cudaFree(h->d_rinfo);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->ev) cudaEventDestroy(h->ev);
    for (auto& e : h->pev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
}
static void free_handle(adb_orb* h) {
    if (!h) return;
    release_resources(h);
    delete h;
}

static adb_status create_impl(const adb_orb_config* cfg, adb_orb* h) {
    h->cfg = *cfg;
    const int nl = cfg->nlevels, W = cfg->width, H = cfg->height, B = cfg->max_batch;
    h->nlevels = nl;
    h->lv.resize(nl);
    // ---- scale tables and quotas: src/ORBextractor.cc:411-445
    std::vector<float> scale(nl), inv(nl);
    h->sigma2.resize(nl); h->inv_sigma2.resize(nl);
    scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nl; ++i) { scale[i] = scale[i - 1] * cfg->scale_factor; h->sigma2[i] = scale[i] * scale[i]; }
    for (int i = 0; i < nl; ++i) { inv[i] = 1.0f / scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    std::vector<int> quota(nl);
    {
        const float factor = 1.0f / cfg->scale_factor;
        float per = cfg->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) { quota[l] = round_even_f(per); sum += quota[l]; per *= factor; }
        quota[nl - 1] = std::max(cfg->nfeatures - sum, 0);
    }
    int cell_base = 0, cand_base = 0, list_base = 0, maxq = 0;
    std::vector<uint32_t> cell_table;
    int max_box_h = 1, max_wcell = 1;
    for (int l = 0; l < nl; ++l) {
        LevelDev& d = h->lv[l].d;
        d.w = round_even_f((float)W * inv[l]);
        d.h = round_even_f((float)H * inv[l]);
        ADB_CHECK(d.w >= 1 && d.h >= 1, ADB_ERR_INVALID, "level %d is empty (%dx%d)", l, d.w, d.h);
        d.pitch = (d.w + 15) & ~15;
        d.frame_stride = (unsigned)(d.pitch * d.h);
        d.mpitch = d.pitch; d.mframe_stride = d.frame_stride;
        d.scale = scale[l]; d.inv_scale = inv[l];
        d.patch_size = (int)(31 * scale[l]);
        d.quota = quota[l];
        // cell grid: src/ORBextractor.cc:775-789
        const int maxBX = d.w - kMinBorder, maxBY = d.h - kMinBorder;
        const float width = (float)(maxBX - kMinBorder), height = (float)(maxBY - kMinBorder);
        d.ncols = (int)(width / 30.f); d.nrows = (int)(height / 30.f);
        if (d.ncols <= 0 || d.nrows <= 0) { d.ncols = d.nrows = 0; d.wcell = d.hcell = 1; }
        else { d.wcell = (int)ceilf(width / d.ncols); d.hcell = (int)ceilf(height / d.nrows); }
        d.ncells = d.ncols * d.nrows;
        d.cell_base = cell_base;
        d.box_w = (15 + d.wcell + 6 + 15) & ~15; d.box_h = d.hcell + 6;
        ADB_CHECK(d.box_w <= kCellBoxWMax && d.box_h <= kCellBoxHMax, ADB_ERR_INVALID, "level %d: cell box %dx%d too large", l, d.box_w, d.box_h);
        h->cell_box_w = std::max(h->cell_box_w, d.box_w <= 64 ? 64 : kCellBoxWMax);
        max_box_h = std::max(max_box_h, d.box_h); max_wcell = std::max(max_wcell, d.wcell);
        ADB_CHECK(d.nrows < 4096 && d.ncols < 4096 && d.w < 4096 && d.h < 4096, ADB_ERR_INVALID, "image too large (max 4095 px per side)");
        d.slotcap = ((d.wcell + 1) / 2) * ((d.hcell + 1) / 2);
        d.cand_base = cand_base; d.cand_cap = d.ncells * d.slotcap;
        // quad-tree roots: src/ORBextractor.cc:545-549
        d.n_ini = d.ncells ? (int)roundf(width / height) : 0;
        // the reference divides by nIni and indexes an empty root vector when it is 0 (src/ORBextractor.cc:545-549): refuse the shape
        ADB_CHECK(d.ncells == 0 || d.n_ini >= 1, ADB_ERR_INVALID, "level %d (%dx%d): region taller than twice its width, nIni = round(w / h) = 0 is undefined in the reference",
                  l, d.w, d.h);
        d.hx = d.n_ini > 0 ? width / d.n_ini : 1.f;
        d.list_base = list_base;
        d.list_cap = std::max(d.quota + 3, 4 * std::max(d.n_ini, 1));
        for (int i = 0; i < d.nrows; ++i)
            for (int j = 0; j < d.ncols; ++j) cell_table.push_back(((uint32_t)l << 24) | ((uint32_t)i << 12) | (uint32_t)j);
        cell_base += d.ncells; cand_base += d.cand_cap; list_base += d.list_cap;
        maxq = std::max(maxq, d.list_cap);
    }
    h->ncells_total = cell_base; h->cand_total = std::max(cand_base, 1); h->list_total = list_base;
    h->capacity = list_base;
    h->qt_maxa = ((maxq + 16) + 31) & ~31;
    ADB_CHECK(h->qt_maxa < (int)kQtFinal, ADB_ERR_INVALID, "nfeatures too large for the quad-tree node table (%d)", h->qt_maxa);
    h->qt_smem = qt_smem_bytes(h->qt_maxa);
    ADB_CHECK(h->qt_smem <= 200 * 1024, ADB_ERR_INVALID, "quad-tree shared memory %zu too large", h->qt_smem);

    ADB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    ADB_CUDA(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
    for (int l = 0; l < nl; ++l) {
        LevelHost& lh = h->lv[l];
        ADB_CUDA(cudaMalloc(&lh.img, (size_t)B * lh.d.frame_stride));
        // the kernels load whole aligned words, so they touch the pitch padding right of column w - 1 (never selected into a
        // result): give those bytes a defined value once, at provisioning time (compute-sanitizer initcheck)
        ADB_CUDA(cudaMemset(lh.img, 0, (size_t)B * lh.d.frame_stride));
        if (l > 0) {
            std::vector<int2> cx, cy;
            linear_coeffs(h->lv[l - 1].d.w, lh.d.w, cx);
            linear_coeffs(h->lv[l - 1].d.h, lh.d.h, cy);
            // the strip kernel's 8-byte source window must cover the neighbours of 4 adjacent outputs
            lh.strip_ok = getenv("ADB_PYR_SIMPLE") == nullptr;
            for (int q4 = 0; q4 < lh.d.w; q4 += 4) {
                const int last = std::min(q4 + 3, lh.d.w - 1);
                if (std::min(cx[last].x + 1, h->lv[l - 1].d.w - 1) - cx[q4].x > 7) lh.strip_ok = false;
            }
            ADB_CUDA(cudaMalloc(&lh.xtab, cx.size() * sizeof(int2)));
            ADB_CUDA(cudaMalloc(&lh.ytab, cy.size() * sizeof(int2)));
            ADB_CUDA(cudaMemcpy(lh.xtab, cx.data(), cx.size() * sizeof(int2), cudaMemcpyHostToDevice));
            ADB_CUDA(cudaMemcpy(lh.ytab, cy.data(), cy.size() * sizeof(int2), cudaMemcpyHostToDevice));
        }
    }
    std::vector<LevelDev> ld(nl);
    for (int l = 0; l < nl; ++l) ld[l] = h->lv[l].d;
    ADB_CUDA(cudaMalloc(&h->d_levels, nl * sizeof(LevelDev)));
    ADB_CUDA(cudaMemcpy(h->d_levels, ld.data(), nl * sizeof(LevelDev), cudaMemcpyHostToDevice));
    // cell order of the warp-per-cell FAST kernel: the cells of single-tile levels (<= 32 columns) first, then the wide ones
    {
        std::vector<uint32_t> narrow, wide;
        for (size_t c = 0; c < cell_table.size(); ++c) (h->lv[cell_table[c] >> 24].d.wcell <= 32 ? narrow : wide).push_back((uint32_t)c);
        h->n_narrow_cells = (int)narrow.size();
        cell_table.insert(cell_table.end(), narrow.begin(), narrow.end());
        cell_table.insert(cell_table.end(), wide.begin(), wide.end());
    }
    ADB_CUDA(cudaMalloc(&h->d_cell_table, std::max<size_t>(cell_table.size(), 1) * 4));
    if (!cell_table.empty()) ADB_CUDA(cudaMemcpy(h->d_cell_table, cell_table.data(), cell_table.size() * 4, cudaMemcpyHostToDevice));
    {
        static const signed char px[512] = AIRDOS_ORB_PATTERN_X;
        static const signed char py[512] = AIRDOS_ORB_PATTERN_Y;
        std::vector<float> pat(8 * 32 * 4);     // test k of descriptor byte `lane`: {x0, y0, x1, y1} as floats (one 16-byte load per test)
        for (int k = 0; k < 8; ++k)
            for (int lane = 0; lane < 32; ++lane)
                for (int e = 0; e < 2; ++e) {
                    pat[4 * (k * 32 + lane) + 2 * e] = (float)px[16 * lane + 2 * k + e];
                    pat[4 * (k * 32 + lane) + 2 * e + 1] = (float)py[16 * lane + 2 * k + e];
                }
        ADB_CUDA(cudaMalloc(&h->d_pattern, pat.size() * 4));
        ADB_CUDA(cudaMemcpy(h->d_pattern, pat.data(), pat.size() * 4, cudaMemcpyHostToDevice));
    }
    ADB_CHECK((unsigned long long)B * h->cand_total < (1ull << 32), ADB_ERR_INVALID, "max_batch x candidate slots exceeds 2^32 entries (%d x %d)", B, h->cand_total);
    const size_t cb = (size_t)B * h->cand_total * 4;
    ADB_CUDA(cudaMalloc(&h->d_cand, cb));
    ADB_CUDA(cudaMalloc(&h->d_qkeys, cb));
    ADB_CUDA(cudaMalloc(&h->d_qstate, cb));
    ADB_CUDA(cudaMalloc(&h->d_cellcnt, (size_t)B * std::max(h->ncells_total, 1) * 2));
    ADB_CUDA(cudaMalloc(&h->d_qcount, (size_t)B * nl * 4));
    ADB_CUDA(cudaMalloc(&h->d_list, (size_t)B * h->list_total * 4));
    ADB_CUDA(cudaMalloc(&h->d_listcnt, (size_t)B * nl * 4));
    ADB_CUDA(cudaMalloc(&h->d_status, 4));
    ADB_CUDA(cudaMemset(h->d_status, 0, 4));
    ADB_CUDA(cudaMalloc(&h->d_kps, (size_t)B * h->capacity * sizeof(adb_keypoint)));
    ADB_CUDA(cudaMalloc(&h->d_desc, (size_t)B * h->capacity * 32));
    // the downloads move whole capacity rows without waiting for the counts: slots past a frame's count must not carry stale device memory
    ADB_CUDA(cudaMemset(h->d_kps, 0, (size_t)B * h->capacity * sizeof(adb_keypoint)));
    ADB_CUDA(cudaMemset(h->d_desc, 0, (size_t)B * h->capacity * 32));
    ADB_CUDA(cudaMalloc(&h->d_counts, (size_t)B * 4));
    ADB_CUDA(cudaMemset(h->d_counts, 0, (size_t)B * 4));
    ADB_CUDA(cudaMallocHost(&h->h_counts, (size_t)B * 4));
    ADB_CUDA(cudaMallocHost(&h->h_status, 4));
    // blurred pyramid (always the handle's own buffers, level 0 included) and its descriptor-patch boxes
    for (int l = 0; l < nl; ++l) {
        const LevelDev& d = h->lv[l].d;
        const size_t bp = (size_t)((d.w + 15) & ~15);
        ADB_CUDA(cudaMalloc(&h->lv[l].blur, (size_t)B * bp * d.h));
        ADB_CUDA(cudaMemset(h->lv[l].blur, 0, (size_t)B * bp * d.h));
        adb_status s = encode_tma_u8_3d(&h->blur_maps.m[l], h->lv[l].blur, d.w, d.h, B, bp, bp * d.h, kPatchBoxW, kBlBoxH);
        if (s != ADB_OK) return s;
        BlurLevels& bl = h->blur_levels;
        bl.src[l] = h->lv[l].img; bl.dst[l] = h->lv[l].blur; bl.w[l] = d.w; bl.h[l] = d.h; bl.spitch[l] = d.pitch; bl.dpitch[l] = (int)bp;
        bl.sfstride[l] = d.frame_stride; bl.dfstride[l] = (unsigned)(bp * d.h);
        bl.ncg[l] = ((d.w + 3) / 4 + kBlGroupW - 1) / kBlGroupW;
        bl.task_end[l] = (l ? bl.task_end[l - 1] : 0) + bl.ncg[l] * ((d.h + kBlRows - 1) / kBlRows);
        bl.nlevels = nl; bl.ntasks = bl.task_end[l];
    }
    // TMA descriptors for levels >= 1 (level 0 is encoded per call: it may alias the caller's buffer)
    for (int l = 1; l < nl; ++l) {
        const LevelDev& d = h->lv[l].d;
        adb_status s = encode_tma_u8_3d(&h->cell_maps.m[l], h->lv[l].img, d.w, d.h, B, d.pitch, d.frame_stride, h->cell_box_w, d.box_h);
        if (s != ADB_OK) return s;
        s = encode_tma_u8_3d(&h->patch_maps.m[l], h->lv[l].img, d.w, d.h, B, d.pitch, d.frame_stride, kPatchBoxW, kIcBoxH);
        if (s != ADB_OK) return s;
    }
    ADB_CUDA(cudaFuncSetAttribute(quadtree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->qt_smem));
    // warp-per-cell FAST: one TMA box per warp (the CTA-per-cell kernel stays for cells wider than kFwTiles column tiles, or ADB_FAST_CTA=1)
    h->fast_tile_bytes = (max_box_h * h->cell_box_w + 127) & ~127;
    {
        const char* e = getenv("ADB_FAST_CTA");
        h->fast_warp_ok = max_wcell <= kFwTiles * kFwStep + 2 && !(e && *e == '1');
        const int bytes = (h->fast_tile_bytes + kFwRingBytes) * kFwWarps + 128;
        for (auto k : {fw_kernel(h->cell_box_w, 1, false), fw_kernel(h->cell_box_w, 1, true), fw_kernel(h->cell_box_w, kFwTiles, false),
                       fw_kernel(h->cell_box_w, kFwTiles, true)})
            ADB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
    ADB_CUDA(cudaFuncSetAttribute(orient_describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DescSmem) + 128));
    ADB_CUDA(cudaDeviceSynchronize());   // the provisioning fills ran on the legacy stream; the handle's own stream does not wait for it
    return ADB_OK;
}

static adb_status ensure_mask_buffers(adb_orb* h) {
    for (auto& l : h->lv)
        if (!l.mask) {
            ADB_CUDA(cudaMalloc(&l.mask, (size_t)h->cfg.max_batch * l.d.mframe_stride));
            ADB_CUDA(cudaMemset(l.mask, 0, (size_t)h->cfg.max_batch * l.d.mframe_stride));   // pitch padding, as for the image levels
        }
    if (!h->mask_stage) {   // host-buffer calls: the caller's masks land here before the erosion (kept for the life of the handle)
        const size_t p0 = (size_t)((h->cfg.width + 15) & ~15);
        ADB_CUDA(cudaMalloc(&h->mask_stage, (size_t)h->cfg.max_batch * p0 * h->cfg.height));
        ADB_CUDA(cudaMemset(h->mask_stage, 0, (size_t)h->cfg.max_batch * p0 * h->cfg.height));
        ADB_CUDA(cudaDeviceSynchronize());
    }
    return ADB_OK;
}

// Copy n frames of w x h bytes between arbitrary pitched layouts (kind decides the direction).
static adb_status copy_frames(void* dst, size_t dpitch, size_t dfstride, const void* src, size_t spitch, size_t sfstride,
                              int w, int h, int n, cudaMemcpyKind kind, cudaStream_t st) {
    if (dpitch == (size_t)w && spitch == (size_t)w && dfstride == (size_t)w * h && sfstride == (size_t)w * h) {
        ADB_CUDA(cudaMemcpyAsync(dst, src, (size_t)w * h * n, kind, st));
    } else if (dfstride == dpitch * h && sfstride == spitch * h) {
        ADB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, w, (size_t)h * n, kind, st));
    } else {
        for (int f = 0; f < n; ++f)
            ADB_CUDA(cudaMemcpy2DAsync((uint8_t*)dst + f * dfstride, dpitch, (const uint8_t*)src + f * sfstride, spitch, w, h, kind, st));
    }
    return ADB_OK;
}

// Launch the whole pipeline on frames whose level 0 is at (l0, pitch, fstride) in device memory.
static bool debug_sync() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ADB_DEBUG_SYNC"); v = (e && *e == '1') ? 1 : 0; }
    return v == 1;
}
#define ADB_STAGE(name)                                                                      \
    do {                                                                                     \
        ADB_CUDA(cudaGetLastError());                                                        \
        if (h->profiling && h->pev_n < 8) ADB_CUDA(cudaEventRecord(h->pev[h->pev_n++], st)); \
        if (debug_sync()) {                                                                  \
            cudaError_t _e = cudaStreamSynchronize(st);                                      \
            if (_e != cudaSuccess) { set_error("stage %s failed: %s", name, cudaGetErrorString(_e)); return ADB_ERR_CUDA; } \
        }                                                                                    \
    } while (0)

// Per call: level-0 view and TMA maps over all `n_total` frames.  Then run_range launches the kernels for frames
// [f0, f0 + n) -- the whole batch at once, or chunk by chunk behind the chunk's host-to-device copy.
static adb_status run_range(adb_orb* h, int f0, int n, bool masked);
static adb_status run_pipeline(adb_orb* h, int n, const uint8_t* l0, int l0_pitch, size_t l0_fstride, bool masked, bool launch = true) {
    cudaStream_t st = h->stream;
    const int nl = h->nlevels;
    h->l0_base = l0; h->l0_pitch = l0_pitch; h->l0_fstride = l0_fstride; h->have_mask = masked; h->last_frames = n;
    h->pev_n = 0;
    if (h->profiling) ADB_CUDA(cudaEventRecord(h->pev[h->pev_n++], st));
    {
        const LevelDev& d = h->lv[0].d;
        adb_status s = encode_tma_u8_3d(&h->cell_maps.m[0], l0, d.w, d.h, n, l0_pitch, l0_fstride, h->cell_box_w, d.box_h);
        if (s != ADB_OK) return s;
        s = encode_tma_u8_3d(&h->patch_maps.m[0], l0, d.w, d.h, n, l0_pitch, l0_fstride, kPatchBoxW, kIcBoxH);
        if (s != ADB_OK) return s;
    }
    // level table with the per-call level-0 geometry
    if (h->lv[0].d.pitch != l0_pitch || h->lv[0].d.frame_stride != (unsigned)l0_fstride) {
        LevelDev d0 = h->lv[0].d;
        d0.pitch = l0_pitch; d0.frame_stride = (unsigned)l0_fstride;
        ADB_CUDA(cudaMemcpyAsync(h->d_levels, &d0, sizeof(LevelDev), cudaMemcpyHostToDevice, st));
        h->lv[0].d = d0;
    }
    return launch ? run_range(h, 0, n, masked) : ADB_OK;
}

static adb_status run_range(adb_orb* h, int f0, int n, bool masked) {
    cudaStream_t st = h->stream;
    const int nl = h->nlevels;
    const uint8_t* l0 = h->l0_base;
    h->launches += (nl - 1) + (h->ncells_total > 0 ? 1 : 0) + 2;   // + the mask levels that need their own launch (counted below)
    // ---- pyramid
    for (int l = 1; l < nl; ++l) {
        const LevelDev& s = h->lv[l - 1].d;
        const LevelDev& d = h->lv[l].d;
        const uint8_t* src = l == 1 ? l0 : h->lv[l - 1].img;
        if (h->lv[l].strip_ok) {
            // rows per thread: enough threads to fill the machine, long enough strips to amortise the first source row
            const int nq = (d.w + 3) / 4;
            const int rows = std::max(4, std::min(16, (int)((long long)nq * d.h * n / 300000)));
            const int nchunks = (d.h + rows - 1) / rows;
            dim3 grid((nq * nchunks + 127) / 128, n);
            // the mask level shares the launch when its source rows are word-addressable like the image's (the handle's own buffers are)
            const bool fuse = masked && s.mpitch >= s.pitch && (s.mpitch & 3) == 0;
            if (fuse)
                pyr_resize_strip_kernel<true><<<grid, 128, 0, st>>>(src, s.w, s.h, s.pitch, s.frame_stride, h->lv[l].img, d.w, d.h, d.pitch, d.frame_stride,
                                                                   h->lv[l].xtab, h->lv[l].ytab, nq, rows, nchunks, f0, h->lv[l - 1].mask, s.mpitch,
                                                                   s.mframe_stride, h->lv[l].mask, d.mpitch, d.mframe_stride);
            else
                pyr_resize_strip_kernel<false><<<grid, 128, 0, st>>>(src, s.w, s.h, s.pitch, s.frame_stride, h->lv[l].img, d.w, d.h, d.pitch, d.frame_stride,
                                                                    h->lv[l].xtab, h->lv[l].ytab, nq, rows, nchunks, f0, nullptr, 0, 0, nullptr, 0, 0);
            if (masked && !fuse) ++h->launches;
            if (masked && !fuse)
                pyr_resize_strip_kernel<false><<<grid, 128, 0, st>>>(h->lv[l - 1].mask, s.w, s.h, s.mpitch, s.mframe_stride, h->lv[l].mask, d.w, d.h, d.mpitch,
                                                                    d.mframe_stride, h->lv[l].xtab, h->lv[l].ytab, nq, rows, nchunks, f0, nullptr, 0, 0, nullptr, 0, 0);
            continue;
        }
        dim3 grid((d.pitch / 4 + 127) / 128, d.h, n);
        pyr_resize_kernel<<<grid, 128, 0, st>>>(src, s.w, s.h, s.pitch, s.frame_stride, h->lv[l].img, d.w, d.h, d.pitch,
                                               d.frame_stride, h->lv[l].xtab, h->lv[l].ytab, f0);
        if (masked) ++h->launches;
        if (masked)
            pyr_resize_kernel<<<grid, 128, 0, st>>>(h->lv[l - 1].mask, s.w, s.h, s.mpitch, s.mframe_stride, h->lv[l].mask, d.w, d.h,
                                                   d.mpitch, d.mframe_stride, h->lv[l].xtab, h->lv[l].ytab, f0);
    }
    ADB_STAGE("pyramid");
    // ---- FAST per cell
    MaskPtrs mp;
    for (int l = 0; l < kMaxLevels; ++l) mp.p[l] = (masked && l < nl) ? h->lv[l].mask : nullptr;
    if (h->ncells_total > 0) {
        dim3 grid(h->ncells_total, n);
        if (h->fast_warp_ok) {
            const int nn = h->n_narrow_cells, nw = h->ncells_total - nn;
            const uint32_t* order = h->d_cell_table + h->ncells_total;
            const size_t smem = (size_t)(h->fast_tile_bytes + kFwRingBytes) * kFwWarps + 128;
            if (nn) {
                dim3 wgrid((nn + kFwWarps - 1) / kFwWarps, n);
                auto kern = fw_kernel(h->cell_box_w, 1, masked);
                kern<<<wgrid, kFwWarps * 32, smem, st>>>(h->cell_maps, h->d_levels, h->d_cell_table, order, nn, mp, h->cfg.ini_th_fast, h->cfg.min_th_fast,
                                                        h->d_cand, h->cand_total, h->d_cellcnt, h->ncells_total, h->fast_tile_bytes, f0);
            }
            if (nw) {
                dim3 wgrid((nw + kFwWarps - 1) / kFwWarps, n);
                auto kern = fw_kernel(h->cell_box_w, kFwTiles, masked);
                kern<<<wgrid, kFwWarps * 32, smem, st>>>(h->cell_maps, h->d_levels, h->d_cell_table, order + nn, nw, mp, h->cfg.ini_th_fast, h->cfg.min_th_fast,
                                                        h->d_cand, h->cand_total, h->d_cellcnt, h->ncells_total, h->fast_tile_bytes, f0);
                h->launches += nn ? 1 : 0;
            }
        } else {
            auto kern = h->cell_box_w == 64 ? fast_cells_kernel<64> : fast_cells_kernel<kCellBoxWMax>;
            kern<<<grid, kFastThreads, 0, st>>>(h->cell_maps, h->d_levels, h->d_cell_table, mp, h->cfg.ini_th_fast, h->cfg.min_th_fast,
                                                h->d_cand, h->cand_total, h->d_cellcnt, h->ncells_total, f0);
        }
        ADB_STAGE("fast_cells");
    }
    // ---- quad-tree
    {
        dim3 grid(nl, n);
        quadtree_kernel<<<grid, kQtThreads, h->qt_smem, st>>>(h->d_levels, nl, h->d_cand, h->cand_total, h->d_cellcnt, h->ncells_total,
                                                             h->d_qkeys, h->d_qstate, h->d_qcount, h->d_list, h->list_total,
                                                             h->d_listcnt, h->qt_maxa, h->d_status, f0);
        ADB_STAGE("quadtree");
    }
    // ---- GaussianBlur of every level (the reference: once per level before the descriptors, src/ORBextractor.cc:1099-1100), one launch
    {
        BlurLevels bl = h->blur_levels;
        bl.src[0] = l0; bl.spitch[0] = h->lv[0].d.pitch; bl.sfstride[0] = h->lv[0].d.frame_stride;
        dim3 grid((bl.ntasks + 3) / 4, n);
        blur7_level_kernel<<<grid, 128, 0, st>>>(bl, f0);
        ADB_STAGE("blur7_level");
    }
    h->launches += 1;
    // ---- orientation + descriptors
    {
        dim3 grid((h->capacity + kDescWarps - 1) / kDescWarps, n);
        orient_describe_kernel<<<grid, kDescWarps * 32, sizeof(DescSmem) + 128, st>>>(h->patch_maps, h->blur_maps, h->d_levels, nl, h->d_list, h->list_total,
                                                                                     h->d_listcnt, h->d_pattern, h->d_kps, h->d_desc,
                                                                                     h->d_counts, h->capacity, h->gather, f0);
        ADB_STAGE("orient_describe");
    }
    return ADB_OK;
}

static adb_status check_device_status(adb_orb* h) {
    ADB_CUDA(cudaMemcpyAsync(h->h_status, h->d_status, 4, cudaMemcpyDeviceToHost, h->stream));
    ADB_CUDA(cudaStreamSynchronize(h->stream));
    if (*h->h_status != 0) {
        const int s = *h->h_status;
        cudaMemsetAsync(h->d_status, 0, 4, h->stream);
        set_error("extractor device status %d (2 = quad-tree pass guard, 3 = node table overflow, 4 = level list overflow)", s);
        return ADB_ERR_CAPACITY;
    }
    return ADB_OK;
}

}  // namespace adb

using namespace adb;

extern "C" {

adb_status adb_orb_create(const adb_orb_config* cfg, adb_orb_t* out) {
    ADB_CHECK(cfg && out, ADB_ERR_INVALID, "null argument");
    *out = nullptr;
    ADB_CHECK(cfg->nlevels >= 1 && cfg->nlevels <= kMaxLevels, ADB_ERR_INVALID, "nlevels %d out of range 1..%d", cfg->nlevels, kMaxLevels);
    const bool lazy = cfg->width == 0 && cfg->height == 0;   // ORBextractor's own 5-argument constructor: the image size comes with the first frame
    ADB_CHECK(cfg->nfeatures >= 1 && cfg->scale_factor > 1.0f && (lazy || (cfg->width >= 1 && cfg->height >= 1)) && cfg->max_batch >= 1 &&
                  cfg->ini_th_fast >= 0 && cfg->min_th_fast >= 0 && cfg->ini_th_fast < 255 && cfg->min_th_fast < 255,
              ADB_ERR_INVALID, "bad extractor configuration");
    adb_status s = select_device(cfg->device);
    if (s != ADB_OK) return s;
    adb_orb* h = new adb_orb();
    if (lazy) {
        // scale tables for the getters (GetScaleFactors ... src/ORBextractor.cc:411-430) are size independent; everything else is
        // provisioned by the first operator() from the size of its image (ensure_size)
        h->cfg = *cfg; h->nlevels = cfg->nlevels; h->lazy = true;
        h->capacity = cfg->nfeatures + 16 * cfg->nlevels;   // upper bound of any provisioned capacity (sum of quotas + per-level slack)
        h->sigma2.resize(cfg->nlevels); h->inv_sigma2.resize(cfg->nlevels); h->lazy_scale.resize(cfg->nlevels);
        h->lazy_scale[0] = 1.0f; h->sigma2[0] = 1.0f;
        for (int i = 1; i < cfg->nlevels; ++i) { h->lazy_scale[i] = h->lazy_scale[i - 1] * cfg->scale_factor; h->sigma2[i] = h->lazy_scale[i] * h->lazy_scale[i]; }
        for (int i = 0; i < cfg->nlevels; ++i) h->inv_sigma2[i] = 1.0f / h->sigma2[i];
        *out = h;
        return ADB_OK;
    }
    s = create_impl(cfg, h);
    if (s != ADB_OK) { free_handle(h); return s; }
    *out = h;
    return ADB_OK;
}

// lazy handles: (re)provision for a w x hh image, keeping the handle's identity, gather targets and profiling switch
static adb_status ensure_size(adb_orb* h, int w, int hh) {
    if (!h->lazy || (h->provisioned && w == h->cfg.width && hh == h->cfg.height)) return ADB_OK;
    ADB_CHECK(w >= 1 && hh >= 1, ADB_ERR_INVALID, "bad image size %dx%d", w, hh);
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    adb_orb_config cfg = h->cfg;
    cfg.width = w; cfg.height = hh;
    adb_orb* fresh = new adb_orb();
    adb_status s = create_impl(&cfg, fresh);
    if (s != ADB_OK) { free_handle(fresh); return s; }
    if (h->stream) cudaStreamSynchronize(h->stream);
    const adb_gather_targets g = h->gather; const bool prof = h->profiling; const long long launches = h->launches;
    release_resources(h);
    *h = std::move(*fresh);
    delete fresh;                 // its resources now belong to *h (no user destructor: only the moved-from vectors are destroyed)
    h->lazy = true; h->provisioned = true; h->gather = g; h->profiling = prof; h->launches = launches;
    return ADB_OK;
}

adb_status adb_orb_destroy(adb_orb_t h) {
    if (!h) return ADB_OK;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    free_handle(h);
    return ADB_OK;
}

int32_t adb_orb_levels(adb_orb_t h) { return h ? h->nlevels : 0; }
int32_t adb_orb_capacity(adb_orb_t h) { return h ? h->capacity : 0; }

adb_status adb_orb_level_info(adb_orb_t h, int32_t level, int32_t* w, int32_t* h_, int32_t* pitch, float* scale,
                              float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* quota) {
    ADB_CHECK(h && level >= 0 && level < h->nlevels, ADB_ERR_INVALID, "bad level");
    if (h->lazy && !h->provisioned) {   // before the first frame only the size-independent tables exist
        if (w) *w = 0;
        if (h_) *h_ = 0;
        if (pitch) *pitch = 0;
        if (scale) *scale = h->lazy_scale[level];
        if (inv_scale) *inv_scale = 1.0f / h->lazy_scale[level];
        if (sigma2) *sigma2 = h->sigma2[level];
        if (inv_sigma2) *inv_sigma2 = h->inv_sigma2[level];
        if (quota) {   // mnFeaturesPerLevel is size independent (src/ORBextractor.cc:432-445)
            const int nl = h->nlevels;
            const float factor = 1.0f / h->cfg.scale_factor;
            float per = h->cfg.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
            int sum = 0, ql = 0;
            for (int l = 0; l < nl - 1; ++l) { const int v = round_even_f(per); if (l == level) ql = v; sum += v; per *= factor; }
            if (level == nl - 1) ql = std::max(h->cfg.nfeatures - sum, 0);
            *quota = ql;
        }
        return ADB_OK;
    }
    const LevelDev& d = h->lv[level].d;
    if (w) *w = d.w;
    if (h_) *h_ = d.h;
    if (pitch) *pitch = d.pitch;
    if (scale) *scale = d.scale;
    if (inv_scale) *inv_scale = d.inv_scale;
    if (sigma2) *sigma2 = h->sigma2[level];
    if (inv_sigma2) *inv_sigma2 = h->inv_sigma2[level];
    if (quota) *quota = d.quota;
    return ADB_OK;
}

void* adb_orb_stream(adb_orb_t h) { return h ? (void*)h->stream : nullptr; }

adb_status adb_orb_set_gather(adb_orb_t h, const adb_gather_targets* t) {
    ADB_CHECK(h, ADB_ERR_INVALID, "null handle");
    if (!t || t->n == 0) { h->gather.n = 0; return ADB_OK; }
    ADB_CHECK(t->n > 0 && t->n <= ADB_MAX_GATHER, ADB_ERR_INVALID, "gather target count %d out of range", t->n);
    for (int g = 0; g < t->n; ++g) ADB_CHECK(t->kps[g] && t->desc[g] && t->counts[g], ADB_ERR_INVALID, "null gather target %d", g);
    h->gather = *t;
    return ADB_OK;
}

adb_status adb_orb_profile(adb_orb_t h, int32_t enable) {
    ADB_CHECK(h, ADB_ERR_INVALID, "null handle");
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    if (enable && !h->pev[0])
        for (auto& e : h->pev) ADB_CUDA(cudaEventCreate(&e));
    h->profiling = enable != 0;
    h->pev_n = 0;
    return ADB_OK;
}

adb_status adb_orb_stage_ms(adb_orb_t h, float* ms5) {
    ADB_CHECK(h && ms5, ADB_ERR_INVALID, "null argument");
    ADB_CHECK(h->profiling && h->pev_n == 6, ADB_ERR_INVALID, "no profiled call recorded (enable adb_orb_profile, then extract)");
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    ADB_CUDA(cudaEventSynchronize(h->pev[5]));
    for (int i = 0; i < 5; ++i) ADB_CUDA(cudaEventElapsedTime(&ms5[i], h->pev[i], h->pev[i + 1]));
    return ADB_OK;
}

int64_t adb_orb_launch_count(adb_orb_t h) { return h ? h->launches : 0; }

adb_status adb_orb_sync(adb_orb_t h) {
    ADB_CHECK(h, ADB_ERR_INVALID, "null handle");
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    return check_device_status(h);
}

// level-0 mask of frames [f0, f0 + n) = erode(caller's mask); d_masks points at the first of the n masks
static adb_status prepare_masks(adb_orb* h, int f0, int n, const uint8_t* d_masks, size_t mfstride, int mpitch) {
    adb_status s = ensure_mask_buffers(h);
    if (s != ADB_OK) return s;
    const LevelDev& d = h->lv[0].d;
    dim3 grid((d.w + kErTW - 1) / kErTW, (d.h + kErTH - 1) / kErTH, n);
    erode10_tile_kernel<<<grid, 256, 0, h->stream>>>(d_masks, d.w, d.h, mpitch, mfstride, h->lv[0].mask, d.mpitch, d.mframe_stride, f0);
    ++h->launches;
    ADB_CUDA(cudaGetLastError());
    return ADB_OK;
}

adb_status adb_orb_extract_batch_device(adb_orb_t h, int32_t n, const uint8_t* d_images, size_t fstride, int32_t w, int32_t hh,
                                        int32_t pitch, const uint8_t* d_masks, size_t mfstride, int32_t mpitch) {
    ADB_CHECK(h && d_images, ADB_ERR_INVALID, "null argument");
    ADB_CHECK(n >= 1 && n <= h->cfg.max_batch, ADB_ERR_INVALID, "n_frames %d exceeds max_batch %d", n, h->cfg.max_batch);
    { const adb_status es = ensure_size(h, w, hh); if (es != ADB_OK) return es; }
    ADB_CHECK(w == h->cfg.width && hh == h->cfg.height && pitch >= w, ADB_ERR_INVALID, "image %dx%d does not match the handle (%dx%d)", w, hh, h->cfg.width, h->cfg.height);
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    const uint8_t* l0 = d_images; int l0p = pitch; size_t l0s = fstride;
    if (((uintptr_t)d_images & 15) || (pitch & 15) || (fstride & 15)) {   // not TMA-addressable: stage a copy
        const int p0 = (w + 15) & ~15;
        adb_status s = copy_frames(h->lv[0].img, p0, (size_t)p0 * hh, d_images, pitch, fstride, w, hh, n, cudaMemcpyDeviceToDevice, h->stream);
        if (s != ADB_OK) return s;
        l0 = h->lv[0].img; l0p = p0; l0s = (size_t)p0 * hh;
    }
    if (d_masks) {
        adb_status s = prepare_masks(h, 0, n, d_masks, mfstride, mpitch);
        if (s != ADB_OK) return s;
    }
    return run_pipeline(h, n, l0, l0p, l0s, d_masks != nullptr);
}

adb_status adb_orb_results_device(adb_orb_t h, const adb_keypoint** d_kps, const uint8_t** d_desc, const int32_t** d_counts, int32_t* capacity) {
    ADB_CHECK(h, ADB_ERR_INVALID, "null handle");
    if (d_kps) *d_kps = h->d_kps;
    if (d_desc) *d_desc = h->d_desc;
    if (d_counts) *d_counts = h->d_counts;
    if (capacity) *capacity = h->capacity;
    return ADB_OK;
}

// device -> host copies of the results of frames [first, first + n) on stream `st` (asynchronous)
static adb_status download_async(adb_orb* h, int first, int n, adb_keypoint* kps, uint8_t* desc, int cap, int32_t* h_counts, cudaStream_t st) {
    ADB_CUDA(cudaMemcpyAsync(h_counts, h->d_counts + first, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    const int rows = std::min(cap, h->capacity);
    if (kps && rows > 0)
        ADB_CUDA(cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(adb_keypoint), h->d_kps + (size_t)first * h->capacity,
                                   (size_t)h->capacity * sizeof(adb_keypoint), (size_t)rows * sizeof(adb_keypoint), n, cudaMemcpyDeviceToHost, st));
    if (desc && rows > 0)
        ADB_CUDA(cudaMemcpy2DAsync(desc, (size_t)cap * 32, h->d_desc + (size_t)first * h->capacity * 32, (size_t)h->capacity * 32,
                                   (size_t)rows * 32, n, cudaMemcpyDeviceToHost, st));
    return ADB_OK;
}

adb_status adb_orb_download(adb_orb_t h, int32_t first, int32_t n, adb_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* counts) {
    ADB_CHECK(h && counts, ADB_ERR_INVALID, "null argument");
    ADB_CHECK(first >= 0 && n >= 0 && first + n <= h->cfg.max_batch, ADB_ERR_INVALID, "bad frame range");
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    adb_status s = download_async(h, first, n, kps, desc, cap, h->h_counts, h->stream);
    if (s != ADB_OK) return s;
    s = check_device_status(h);
    if (s != ADB_OK) return s;
    for (int i = 0; i < n; ++i) {
        counts[i] = h->h_counts[i];
        ADB_CHECK(counts[i] <= cap, ADB_ERR_CAPACITY, "frame %d holds %d key-points, caller capacity %d", first + i, counts[i], cap);
    }
    return ADB_OK;
}

// Large host batches run as a pipeline of chunks: the host-to-device copy of chunk c + 1 (copy stream), the kernels of chunk c
// (handle stream) and the device-to-host copy of chunk c - 1's results (download stream) overlap, so the call costs about
// max(upload, compute, download) instead of their sum.  Results are identical: chunking only changes the launch ranges.
constexpr int kChunks = 32, kMinChunkedFrames = 32;   // 8 chunks once the batch is chunked at all, more (of >= 256 frames, at most kChunks) for very large batches

static int chunk_count(int n) { return std::min(kChunks, std::max(8, n / 256)); }   // fill / drain = one chunk of upload + one of download

// streams / events of the pipeline, entry fence, per-call level-0 view (no launch yet)
static adb_status chunk_begin(adb_orb* h, int n, int w, int hh, bool masked) {
    const int p0 = (w + 15) & ~15;
    if (!h->copy_stream) {
        ADB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        ADB_CUDA(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2 * kChunks + 1; ++i) ADB_CUDA(cudaEventCreateWithFlags(&h->cev[i], cudaEventDisableTiming));
    }
    adb_status s = ADB_OK;
    if (masked && (s = ensure_mask_buffers(h)) != ADB_OK) return s;
    // the upload must not overtake kernels of an earlier asynchronous call that still read the staging buffer
    ADB_CUDA(cudaEventRecord(h->cev[2 * kChunks], h->stream));
    ADB_CUDA(cudaStreamWaitEvent(h->copy_stream, h->cev[2 * kChunks], 0));
    return run_pipeline(h, n, h->lv[0].img, p0, (size_t)p0 * hh, masked, /*launch=*/false);
}

// chunk c = frames [f0, f0 + nc): upload on the copy stream, kernels on the handle's stream behind it; cev[kChunks + c] = chunk computed
static adb_status chunk_issue(adb_orb* h, int c, int f0, int nc, const uint8_t* images, size_t fstride, int w, int hh, int pitch,
                              const uint8_t* masks, size_t mfstride, int mpitch) {
    const int p0 = (w + 15) & ~15;
    const size_t dfs = (size_t)p0 * hh;
    adb_status s = copy_frames(h->lv[0].img + f0 * dfs, p0, dfs, images + f0 * fstride, pitch, fstride, w, hh, nc, cudaMemcpyHostToDevice, h->copy_stream);
    if (s != ADB_OK) return s;
    if (masks) {   // the chunk's masks ride on the same copy stream; erosion + mask pyramid run with the chunk's kernels
        s = copy_frames(h->mask_stage + f0 * dfs, p0, dfs, masks + f0 * mfstride, mpitch, mfstride, w, hh, nc, cudaMemcpyHostToDevice, h->copy_stream);
        if (s != ADB_OK) return s;
    }
    ADB_CUDA(cudaEventRecord(h->cev[c], h->copy_stream));
    ADB_CUDA(cudaStreamWaitEvent(h->stream, h->cev[c], 0));
    if (masks && (s = prepare_masks(h, f0, nc, h->mask_stage + f0 * dfs, dfs, p0)) != ADB_OK) return s;
    s = run_range(h, f0, nc, masks != nullptr);
    if (s != ADB_OK) return s;
    ADB_CUDA(cudaEventRecord(h->cev[kChunks + c], h->stream));
    return ADB_OK;
}

static adb_status chunk_finish(adb_orb* h, int n, int cap, int32_t* counts) {
    ADB_CUDA(cudaStreamSynchronize(h->d2h_stream));
    adb_status s = check_device_status(h);
    if (s != ADB_OK) return s;
    for (int i = 0; i < n; ++i) {
        counts[i] = h->h_counts[i];
        ADB_CHECK(counts[i] <= cap, ADB_ERR_CAPACITY, "frame %d holds %d key-points, caller capacity %d", i, counts[i], cap);
    }
    return ADB_OK;
}

static adb_status extract_batch_chunked(adb_orb* h, int n, const uint8_t* images, size_t fstride, int w, int hh, int pitch, const uint8_t* masks,
                                        size_t mfstride, int mpitch, adb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts) {
    adb_status s = chunk_begin(h, n, w, hh, masks != nullptr);
    if (s != ADB_OK) return s;
    const int per = (n + chunk_count(n) - 1) / chunk_count(n);
    int c = 0;
    for (int f0 = 0; f0 < n; f0 += per, ++c) {
        const int nc = std::min(per, n - f0);
        s = chunk_issue(h, c, f0, nc, images, fstride, w, hh, pitch, masks, mfstride, mpitch);
        if (s != ADB_OK) return s;
        ADB_CUDA(cudaStreamWaitEvent(h->d2h_stream, h->cev[kChunks + c], 0));
        s = download_async(h, f0, nc, kps ? kps + (size_t)f0 * cap : nullptr, desc ? desc + (size_t)f0 * cap * 32 : nullptr, cap,
                           h->h_counts + f0, h->d2h_stream);
        if (s != ADB_OK) return s;
    }
    return chunk_finish(h, n, cap, counts);
}

adb_status adb_orb_extract_batch(adb_orb_t h, int32_t n, const uint8_t* images, size_t fstride, int32_t w, int32_t hh, int32_t pitch,
                                 const uint8_t* masks, size_t mfstride, int32_t mpitch, adb_keypoint* kps, uint8_t* desc,
                                 int32_t cap, int32_t* counts) {
    ADB_CHECK(h && counts, ADB_ERR_INVALID, "null argument");
    if (w == 0 || hh == 0 || !images) {   // src/ORBextractor.cc:1057-1058: empty image -> silent return
        for (int i = 0; i < n; ++i) counts[i] = 0;
        return ADB_OK;
    }
    ADB_CHECK(n >= 1 && n <= h->cfg.max_batch, ADB_ERR_INVALID, "n_frames %d exceeds max_batch %d", n, h->cfg.max_batch);
    { const adb_status es = ensure_size(h, w, hh); if (es != ADB_OK) return es; }
    ADB_CHECK(w == h->cfg.width && hh == h->cfg.height && pitch >= w, ADB_ERR_INVALID, "image %dx%d does not match the handle (%dx%d)", w, hh, h->cfg.width, h->cfg.height);
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    static const bool no_chunks = getenv("ADB_NO_CHUNKS") != nullptr;   // measurement switch
    if (n >= kMinChunkedFrames && !h->profiling && !no_chunks)
        return extract_batch_chunked(h, n, images, fstride, w, hh, pitch, masks, mfstride, mpitch, kps, desc, cap, counts);
    const int p0 = (w + 15) & ~15;
    adb_status s = copy_frames(h->lv[0].img, p0, (size_t)p0 * hh, images, pitch, fstride, w, hh, n, cudaMemcpyHostToDevice, h->stream);
    if (s != ADB_OK) return s;
    if (masks) {
        s = ensure_mask_buffers(h);
        if (s != ADB_OK) return s;
        s = copy_frames(h->mask_stage, p0, (size_t)p0 * hh, masks, mpitch, mfstride, w, hh, n, cudaMemcpyHostToDevice, h->stream);
        if (s == ADB_OK) s = prepare_masks(h, 0, n, h->mask_stage, (size_t)p0 * hh, p0);
        if (s != ADB_OK) return s;
    }
    s = run_pipeline(h, n, h->lv[0].img, p0, (size_t)p0 * hh, masks != nullptr);
    if (s != ADB_OK) return s;
    return adb_orb_download(h, 0, n, kps, desc, cap, counts);
}

adb_status adb_orb_extract(adb_orb_t h, const uint8_t* image, int32_t w, int32_t hh, int32_t pitch, const uint8_t* mask,
                           int32_t mpitch, adb_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* n_out) {
    ADB_CHECK(n_out, ADB_ERR_INVALID, "null argument");
    return adb_orb_extract_batch(h, 1, image, (size_t)pitch * hh, w, hh, pitch, mask, (size_t)mpitch * hh, mpitch, kps, desc, cap, n_out);
}

// The hot part of the stereo Frame constructor (src/Frame.cc:80-100: ExtractORB left / right on two threads, then ComputeStereoMatches) for
// n stereo pairs in host memory, as ONE pipeline: per chunk of frames the two uploads, the two extractions (one stream per handle), the
// stereo matcher of the chunk behind both, and the downloads of everything the chunk produced -- so the matcher and its results overlap
// the next chunks' copies instead of running after the last one.  Results are identical to adb_orb_extract_batch x 2 + adb_stereo_match.
adb_status adb_stereo_frames_batch(adb_orb_t L, adb_orb_t R, int32_t n, const uint8_t* imagesL, const uint8_t* imagesR, size_t fstride, int32_t w,
                                   int32_t hh, int32_t pitch, const uint8_t* masksL, const uint8_t* masksR, size_t mfstride, int32_t mpitch,
                                   adb_keypoint* kpsL, uint8_t* descL, int32_t* countsL, adb_keypoint* kpsR, uint8_t* descR, int32_t* countsR,
                                   int32_t cap, float mb, float mbf, float* ur, float* dp, int32_t* bi, int32_t* bd) {
    ADB_CHECK(L && R && L != R && imagesL && imagesR && countsL && countsR && ur && dp, ADB_ERR_INVALID, "null argument");
    ADB_CHECK((masksL == nullptr) == (masksR == nullptr), ADB_ERR_INVALID, "masks must be given for both images or for neither");
    ADB_CHECK(n >= 1 && n <= L->cfg.max_batch && n <= R->cfg.max_batch, ADB_ERR_INVALID, "n_frames %d exceeds max_batch", n);
    ADB_CHECK(w > 0 && hh > 0 && pitch >= w, ADB_ERR_INVALID, "bad image geometry");
    static const bool no_chunks = getenv("ADB_NO_CHUNKS") != nullptr;   // measurement switch
    if (n < kMinChunkedFrames || L->profiling || R->profiling || no_chunks) {   // small batches: the three calls it stands for
        adb_status s = adb_orb_extract_batch(L, n, imagesL, fstride, w, hh, pitch, masksL, mfstride, mpitch, kpsL, descL, cap, countsL);
        if (s == ADB_OK) s = adb_orb_extract_batch(R, n, imagesR, fstride, w, hh, pitch, masksR, mfstride, mpitch, kpsR, descR, cap, countsR);
        if (s == ADB_OK) s = adb_stereo_match(L, R, n, mb, mbf, ur, dp, bi, bd, cap);
        return s;
    }
    for (adb_orb* h : {L, R}) {
        const adb_status es = ensure_size(h, w, hh);
        if (es != ADB_OK) return es;
        ADB_CHECK(w == h->cfg.width && hh == h->cfg.height, ADB_ERR_INVALID, "image %dx%d does not match the handle (%dx%d)", w, hh, h->cfg.width, h->cfg.height);
    }
    ADB_CHECK(L->cfg.device == R->cfg.device, ADB_ERR_INVALID, "left / right extractors live on different devices");
    ADB_CUDA(cudaSetDevice(L->cfg.device));
    adb_status s = chunk_begin(L, n, w, hh, masksL != nullptr);
    if (s == ADB_OK) s = chunk_begin(R, n, w, hh, masksR != nullptr);
    if (s != ADB_OK) return s;
    if (!L->sev[0]) {
        for (auto& e : L->sev) ADB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ADB_CUDA(cudaStreamCreateWithFlags(&L->stereo_stream, cudaStreamNonBlocking));
    }
    const int per = (n + chunk_count(n) - 1) / chunk_count(n);
    int c = 0;
    for (int f0 = 0; f0 < n; f0 += per, ++c) {
        const int nc = std::min(per, n - f0);
        s = chunk_issue(L, c, f0, nc, imagesL, fstride, w, hh, pitch, masksL, mfstride, mpitch);
        if (s == ADB_OK) s = chunk_issue(R, c, f0, nc, imagesR, fstride, w, hh, pitch, masksR, mfstride, mpitch);
        if (s != ADB_OK) return s;
        // the right handle's results leave as soon as its chunk is done
        ADB_CUDA(cudaStreamWaitEvent(R->d2h_stream, R->cev[kChunks + c], 0));
        s = download_async(R, f0, nc, kpsR ? kpsR + (size_t)f0 * cap : nullptr, descR ? descR + (size_t)f0 * cap * 32 : nullptr, cap, R->h_counts + f0,
                           R->d2h_stream);
        if (s != ADB_OK) return s;
        // the left handle's own results as soon as its chunk is done
        ADB_CUDA(cudaStreamWaitEvent(L->d2h_stream, L->cev[kChunks + c], 0));
        s = download_async(L, f0, nc, kpsL ? kpsL + (size_t)f0 * cap : nullptr, descL ? descL + (size_t)f0 * cap * 32 : nullptr, cap, L->h_counts + f0,
                           L->d2h_stream);
        if (s != ADB_OK) return s;
        // the matcher of the chunk runs on its own stream behind both extractions of the chunk (on either handle's stream it would hold
        // up that handle's next chunk: measured 49 instead of 37 ms per 2048 pairs); its results follow on the left download stream
        ADB_CUDA(cudaStreamWaitEvent(L->stereo_stream, L->cev[kChunks + c], 0));
        ADB_CUDA(cudaStreamWaitEvent(L->stereo_stream, R->cev[kChunks + c], 0));
        s = adb_stereo_match_range(L, R, f0, nc, mb, mbf, L->stereo_stream);
        if (s != ADB_OK) return s;
        ADB_CUDA(cudaEventRecord(L->sev[c], L->stereo_stream));
        ADB_CUDA(cudaStreamWaitEvent(L->d2h_stream, L->sev[c], 0));
        s = adb_stereo_download_range(L, f0, nc, ur, dp, bi, bd, cap, L->d2h_stream);
        if (s != ADB_OK) return s;
    }
    // later calls on the left handle's stream (and adb_orb_sync) must see the matcher's results too
    ADB_CUDA(cudaEventRecord(L->sev[kChunks - 1], L->stereo_stream));
    ADB_CUDA(cudaStreamWaitEvent(L->stream, L->sev[kChunks - 1], 0));
    s = chunk_finish(R, n, cap, countsR);
    if (s != ADB_OK) return s;
    return chunk_finish(L, n, cap, countsL);
}

adb_status adb_orb_get_pyramid(adb_orb_t h, int32_t frame, int32_t level, int32_t which, uint8_t* dst, int32_t dpitch) {
    ADB_CHECK(h && dst && level >= 0 && level < h->nlevels && frame >= 0 && frame < h->last_frames, ADB_ERR_INVALID, "bad argument");
    ADB_CHECK(which == 0 || (which == 1 && h->have_mask), ADB_ERR_INVALID, "no mask pyramid resident");
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    const LevelDev& d = h->lv[level].d;
    const uint8_t* src; size_t sp, fs;
    if (which == 0) {
        if (level == 0) { src = h->l0_base; sp = h->l0_pitch; fs = h->l0_fstride; }
        else { src = h->lv[level].img; sp = d.pitch; fs = d.frame_stride; }
    } else {
        src = h->lv[level].mask; sp = d.mpitch; fs = d.mframe_stride;
    }
    ADB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src + frame * fs, sp, d.w, d.h, cudaMemcpyDeviceToHost, h->stream));
    ADB_CUDA(cudaStreamSynchronize(h->stream));
    return ADB_OK;
}

adb_status adb_orb_debug_candidates(adb_orb_t h, int32_t frame, int32_t level, int32_t* xys, int32_t cap, int32_t* n) {
    ADB_CHECK(h && n && level >= 0 && level < h->nlevels && frame >= 0 && frame < h->last_frames, ADB_ERR_INVALID, "bad argument");
    ADB_CUDA(cudaSetDevice(h->cfg.device));
    int32_t cnt = 0;
    ADB_CUDA(cudaMemcpyAsync(&cnt, h->d_qcount + frame * h->nlevels + level, 4, cudaMemcpyDeviceToHost, h->stream));
    ADB_CUDA(cudaStreamSynchronize(h->stream));
    *n = cnt;
    const int m = std::min(cnt, cap);
    if (m > 0 && xys) {
        int32_t* tmp = nullptr;
        ADB_CUDA(cudaMalloc(&tmp, (size_t)m * 12));
        unpack_candidates_kernel<<<(m + 255) / 256, 256, 0, h->stream>>>(h->d_qkeys + (size_t)frame * h->cand_total + h->lv[level].d.cand_base, m, tmp);
        cudaError_t e = cudaMemcpyAsync(xys, tmp, (size_t)m * 12, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(tmp);
        ADB_CUDA(e);
    }
    return ADB_OK;
}

}  // extern "C"
