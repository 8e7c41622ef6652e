// chol.cu -- dense FP64 Cholesky solve of the reduced (Schur) system on sm_100a: ONE launch, ONE thread-block cluster.
//
// Replaces g2o's LinearSolverDense::solve (Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:64-113, Eigen LDLT of the
// whole non-marginalised block in LocalBundleAdjustmentHumanTrajactory) and LinearSolverEigen::solve
// (linear_solver_eigen.h:92-115, SimplicialLDLT of the 6K x 6K camera block in LocalBundleAdjustment).  Orders are
// n = 294 (50 key-frames) ... 1226 (80 key-frames + 16 skeletons): far too small to fill 148 SMs, so the solve is a chain
// of n dependent pivots and what matters is the length of that chain, not FLOP/s.  Design:
//
//   * right-looking blocked Cholesky, 32 x 32 tiles, the matrix stays in L2 (12 MB at n = 1226);
//   * one cluster of 8 (portable) or 16 CTAs runs all ceil(n/32) steps inside one kernel; the two dependencies of a step
//     (panel solved -> trailing update, trailing update -> next diagonal tile) are hardware cluster barriers
//     (barrier.cluster arrive.release / wait.acquire, ~0.2 us) instead of kernel boundaries (~3 us) or a grid-wide
//     software barrier;
//   * step k: every CTA factors the diagonal tile redundantly (one warp, rows in registers) -- no barrier between the
//     factorisation and the panel solve; the row blocks below are solved by substitution, one warp per two rows;
//   * the trailing update C_ij -= L_ik L_jk^T runs on the FP64 tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA), one
//     8 x 8 sub-tile per warp, fragments loaded straight from the row-major panel (the k order inside a fragment is free
//     as long as A and B agree, so each lane loads 16-byte pairs);
//   * the right-hand side rides along as row block `nblk` of the array (forward substitution for free); the inverse of
//     every diagonal tile is produced by one more substitution pass on the identity, so the backward substitution is
//     ceil(n/32) mat-vec steps instead of n serial pivots.
//
// Array layout (device): S = (nblk + 1) * 32 rows x ld columns, ld = nblk * 32, row-major, lower triangle used; rows / columns
// n..ld-1 are an identity pad; row ld holds b^T, rows ld+1.. are zero.  Not positive definite -> *info = failing column + 1
// (the LM loop rejects the step, like g2o's `!ldlt.isPositive()`), arithmetic continues with a unit pivot.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "chol.cuh"

namespace adb {

namespace {

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the two halves of the barrier: CTA 0 arrives as soon as its panel rows are stored and waits only after it has factored the next
// diagonal tile, so that nobody waits for that factorisation at the panel barrier
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// D(8x8) += A(8x4) B(4x8) on the FP64 tensor pipe.  Lane l: a = A[l/4][l%4], b = B[l%4][l/4], d = D[l/4][2(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int kCholSB = 8;                                  // sub-block order inside a tile
constexpr int kCholNSB = kCholNB / kCholSB;                 // 4 sub-blocks
constexpr int kTbPitch = kCholNB + 1;
// published diagonal factor: L tile (row-major 32 x 32) + the inverses of its four 8 x 8 diagonal sub-blocks
constexpr int kCholLd = kCholNB * kCholNB + kCholNSB * kCholSB * kCholSB;

struct TileSmem {
    double Tb[kCholNB][kTbPitch];             // diagonal tile -> its factor L (lower; zeros above)
    double Vb[2][kCholNB][kCholSB + 1];       // the 8 columns of the sub-block just factored (double buffered)
    double Dinv[kCholNSB][kCholSB][kCholSB];  // inverses of the 8 x 8 diagonal sub-blocks of L
    double isd[kCholNB];                      // 1 / L[j][j]
};

// One 8-wide sub-block of the tile factorisation, warp 0 only.  The chain of pivots is what costs, so it runs WITHOUT any
// cross-lane traffic: every lane factors the 8 x 8 diagonal sub-block redundantly in registers (left-looking, so that the last
// term added before each rsqrt is the only one that depends on the previous pivot: chain = mul + fma + rsqrt per pivot) and
// solves its own row of the sub-panel against it.  Returns true if a pivot was not positive / finite (then replaced by 1).
__device__ __forceinline__ bool factor_chain(TileSmem& T, int jb, int lane) {
    const int j0 = jb * kCholSB;
    bool bad = false;
    double L[kCholSB][kCholSB], r[kCholSB];
#pragma unroll
    for (int p = 0; p < kCholSB; ++p) {
        r[p] = T.Tb[lane][j0 + p];
#pragma unroll
        for (int q = 0; q <= p; ++q) L[p][q] = T.Tb[j0 + p][j0 + q];   // broadcast reads
    }
    // right-looking inside the sub-block: after pivot p every remaining entry gets ONE independent fma, and the two operations
    // the next pivot waits for (its sub-diagonal entry and its diagonal) are issued first -- the rest fills the rsqrt latency
#pragma unroll
    for (int p = 0; p < kCholSB; ++p) {
        double d = L[p][p];
        if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
        const double isd = rsqrt(d);
        L[p][p] = d * isd;
        if (p + 1 < kCholSB) {
            L[p + 1][p] *= isd;
            L[p + 1][p + 1] = fma(-L[p + 1][p], L[p + 1][p], L[p + 1][p + 1]);
        }
#pragma unroll
        for (int q = p + 2; q < kCholSB; ++q) L[q][p] *= isd;
#pragma unroll
        for (int q = p + 2; q < kCholSB; ++q)
#pragma unroll
            for (int u = p + 1; u <= q; ++u) L[q][u] = fma(-L[q][p], L[u][p], L[q][u]);
        double t = r[p] * isd;
        if (lane == j0 + p) t = L[p][p];     // the pivot row: exactly sqrt(d) (also when the pivot was replaced)
        if (lane < j0 + p) t = 0.0;          // above the diagonal
        r[p] = t;
#pragma unroll
        for (int u = p + 1; u < kCholSB; ++u) r[u] = fma(-t, L[u][p], r[u]);
        if (lane == 0) T.isd[j0 + p] = isd;
    }
    __syncwarp();   // lanes j0 .. j0 + 7 overwrite the rows every lane read above (racecheck: write-after-read inside the warp)
#pragma unroll
    for (int p = 0; p < kCholSB; ++p) { T.Tb[lane][j0 + p] = r[p]; T.Vb[jb & 1][lane][p] = r[p]; }
    return bad;
}

// rank-8 update of columns [c0, c1) of the tile with sub-block jb: Tb[i][q] -= sum_p V[i][p] V[q][p], i >= q; threads
// `first` .. `first + count - 1` of the CTA share the elements
__device__ __forceinline__ void factor_update(TileSmem& T, int jb, int c0, int c1, int idx, int count) {
    const double (*V)[kCholSB + 1] = T.Vb[jb & 1];
    for (int e = idx; e < (c1 - c0) * kCholNB; e += count) {
        const int q = c0 + (e >> 5), i = e & 31;
        if (i >= q) {
            double sacc = T.Tb[i][q];
#pragma unroll
            for (int p = 0; p < kCholSB; ++p) sacc = fma(-V[i][p], V[q][p], sacc);
            T.Tb[i][q] = sacc;
        }
    }
}

// Factor the 32 x 32 tile in T.Tb (lower triangle) in place and invert its four diagonal sub-blocks; all threads of the CTA call
// it.  Look-ahead inside the tile: after sub-block jb only the 8 columns of sub-block jb + 1 are updated by everybody (one
// element per thread); then warp 0 runs the next chain of pivots while warps 1-7 update the columns further right.
__device__ __forceinline__ bool factor_tile(TileSmem& T, int tid, long long* fp = nullptr) {
    const int warp = tid >> 5, lane = tid & 31;
    bool bad = false;
    long long tl = fp ? clock64() : 0;
#define FMARK(slot) do { if (fp) { const long long n_ = clock64(); fp[slot] += n_ - tl; tl = n_; } } while (0)
    if (warp == 0) bad = factor_chain(T, 0, lane);
    FMARK(0);
    __syncthreads();
    FMARK(1);
#pragma unroll 1
    for (int jb = 0; jb + 1 < kCholNSB; ++jb) {
        const int next0 = (jb + 1) * kCholSB;
        factor_update(T, jb, next0, next0 + kCholSB, tid, kCholThreads);
        FMARK(2);
        __syncthreads();
        FMARK(1);
        if (warp == 0) bad |= factor_chain(T, jb + 1, lane);
        else factor_update(T, jb, next0 + kCholSB, kCholNB, tid - 32, kCholThreads - 32);
        FMARK(0);
        __syncthreads();
        FMARK(1);
    }
    // inverses of the 8 x 8 diagonal sub-blocks (the panel solve multiplies by them on the tensor pipe): warp w < 4 takes
    // sub-block w, lane p < 8 the column p of M = L_sub^-1:  M[p][p] = 1 / L[p][p],  M[q][p] = -(sum_{u=p}^{q-1} L[q][u] M[u][p]) / L[q][q]
    if (warp < kCholNSB && lane < kCholSB) {
        const int j0 = warp * kCholSB, p = lane;
        double M[kCholSB];
#pragma unroll
        for (int q = 0; q < kCholSB; ++q) {
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < q; ++u) acc = fma(T.Tb[j0 + q][j0 + u], (u >= p ? M[u] : 0.0), acc);
            M[q] = q < p ? 0.0 : (q == p ? T.isd[j0 + q] : -acc * T.isd[j0 + q]);
        }
#pragma unroll
        for (int q = 0; q < kCholSB; ++q) T.Dinv[warp][q][p] = M[q];
    }
    FMARK(3);
    __syncthreads();
    FMARK(1);
#undef FMARK
    return bad;
}

// X L^T = U for one group of 8 rows on the FP64 tensor pipe (blocked forward substitution over the four sub-blocks):
//     X_j = (U_j - sum_{p<j} X_p L_jp^T) Dinv_j^T.
// x[j] holds sub-tile j of the row group in the DMMA accumulator layout (lane (g, q): row g, columns 8 j + 2 q, + 1).  The same
// registers serve as A fragments of the next products: with the k order {0,2,4,6 | 1,3,5,7} (free, as long as A and B agree)
// the two A fragments of an 8 x 8 block are exactly its accumulator registers -- no shuffles anywhere.
__device__ __forceinline__ void trsm_group(const TileSmem& T, int g, int q, double2 (&x)[kCholNSB]) {
#pragma unroll
    for (int j = 0; j < kCholNSB; ++j) {
#pragma unroll
        for (int p = 0; p < j; ++p) {
            const double b0 = T.Tb[8 * j + g][8 * p + 2 * q], b1 = T.Tb[8 * j + g][8 * p + 2 * q + 1];
            dmma884(x[j].x, x[j].y, -x[p].x, b0);
            dmma884(x[j].x, x[j].y, -x[p].y, b1);
        }
        const double2 d = *reinterpret_cast<const double2*>(&T.Dinv[j][g][2 * q]);
        double2 nx = make_double2(0.0, 0.0);
        dmma884(nx.x, nx.y, x[j].x, d.x);
        dmma884(nx.x, nx.y, x[j].y, d.y);
        x[j] = nx;
    }
}

// tile t of the trailing update of a step with m trailing column blocks, row-major over the lower triangle (a >= b), then the
// m tiles of the right-hand-side row block (a = m)
__device__ __forceinline__ void tile_of(int t, int m, int& a, int& b) {
    const int ntri = m * (m + 1) / 2;
    if (t >= ntri) { a = m; b = t - ntri; return; }
    a = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((a + 1) * (a + 2) / 2 <= t) ++a;
    while (a * (a + 1) / 2 > t) --a;
    b = t - a * (a + 1) / 2;
}

// C(32 x 32) -= A(32 x 32) B(32 x 32)^T by ONE warp: 16 accumulator sub-tiles (independent DMMA chains), operands streamed in
// four chunks of 8 columns straight from the row-major panel (lane (g, q) loads the 16-byte pair at columns 8 s + 2 q).
//   pa / pb: this lane's row g of the two panel tiles at column 2 q;  pc: its C elements (row g, column 2 q)
//   rg_count: row groups to compute (1 for the right-hand-side row block);  diag: skip sub-tiles above the diagonal
//   to_smem != nullptr: write the result there (lower triangle, zeros above) instead of back to pc
__device__ __forceinline__ void update_tile_warp(const double* pa, const double* pb, double* pc, size_t ld, int rg_count, bool diag,
                                                 double (*to_smem)[kTbPitch], int g, int q) {
    double2 acc[4][4];
#pragma unroll
    for (int rg = 0; rg < 4; ++rg)
#pragma unroll
        for (int cg = 0; cg < 4; ++cg)
            if (rg < rg_count && !(diag && cg > rg)) acc[rg][cg] = *reinterpret_cast<const double2*>(pc + (size_t)(8 * rg) * ld + 8 * cg);
    double2 av[4], bv[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        if (x < rg_count) av[x] = *reinterpret_cast<const double2*>(pa + (size_t)(8 * x) * ld);
        bv[x] = *reinterpret_cast<const double2*>(pb + (size_t)(8 * x) * ld);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        double2 an[4], bn[4];
        if (s < 3) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                if (x < rg_count) an[x] = *reinterpret_cast<const double2*>(pa + (size_t)(8 * x) * ld + 8 * (s + 1));
                bn[x] = *reinterpret_cast<const double2*>(pb + (size_t)(8 * x) * ld + 8 * (s + 1));
            }
        }
#pragma unroll
        for (int rg = 0; rg < 4; ++rg)
#pragma unroll
            for (int cg = 0; cg < 4; ++cg)
                if (rg < rg_count && !(diag && cg > rg)) {
                    dmma884(acc[rg][cg].x, acc[rg][cg].y, -av[rg].x, bv[cg].x);
                    dmma884(acc[rg][cg].x, acc[rg][cg].y, -av[rg].y, bv[cg].y);
                }
        if (s < 3) {
#pragma unroll
            for (int x = 0; x < 4; ++x) { av[x] = an[x]; bv[x] = bn[x]; }
        }
    }
#pragma unroll
    for (int rg = 0; rg < 4; ++rg)
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
            if (to_smem) {
                const int r = 8 * rg + g, c = 8 * cg + 2 * q;
                const bool have = !(diag && cg > rg);
                to_smem[r][c] = have && c <= r ? acc[rg][cg].x : 0.0;
                to_smem[r][c + 1] = have && c + 1 <= r ? acc[rg][cg].y : 0.0;
            } else if (rg < rg_count && !(diag && cg > rg)) {
                *reinterpret_cast<double2*>(pc + (size_t)(8 * rg) * ld + 8 * cg) = acc[rg][cg];
            }
        }
}

constexpr int kCholFactorTiles = 12;   // the look-ahead factorisation costs CTA 0 about this many tile updates of its share

}  // namespace

size_t chol_scratch_elems(int n) { const size_t nb = (size_t)chol_nblk(n); return nb * kCholNB * kCholNB + nb * kCholLd; }

__global__ void __launch_bounds__(kCholThreads, 1) chol_cluster_kernel(double* S, int ld, int nblk, double* scratch, double* xout, int* info, const int* skip, long long* prof) {
    if (skip != nullptr && *skip != 0) return;   // device-side LM control: the round is already over (uniform over the cluster)
    __shared__ TileSmem T;
    __shared__ double Pn[kCholNB][kTbPitch];   // CTA 0: its freshly solved panel tile L(k + 1, k), operand of the next diagonal tile's update
    __shared__ double xb[kCholNB];
    extern __shared__ double yb[];   // [ld] (back-substitution, CTA 0)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int rank = (int)cluster_rank(), C = (int)cluster_size();
    const int nrb = nblk + 1;   // row blocks incl. the right-hand-side block
    double* LinvT = scratch;                                         // [nblk][32][32] inverse-transposed diagonal factors
    double* Ld = scratch + (size_t)nblk * kCholNB * kCholNB;         // [nblk][kCholLd] diagonal factors, published by CTA 0
    // optional phase profile (thread 0 of every CTA, SM clock): load, factor, panel, inverse, barrier 1, update, barrier 2, back-substitution
    __shared__ long long pacc[9];   // [8] = last time stamp
    __shared__ long long facc[4];   // inside the tile factorisation: pivot chains, barriers, look-ahead column update, inverses
    const bool profiling = prof != nullptr && tid == 0;
#define CHOL_MARK(slot) do { if (profiling) { const long long now_ = clock64(); pacc[slot] += now_ - pacc[8]; pacc[8] = now_; } } while (0)
    if (profiling) { for (int i = 0; i < 8; ++i) pacc[i] = 0; for (int i = 0; i < 4; ++i) facc[i] = 0; pacc[8] = clock64(); }

    // factor the tile in T.Tb and publish it as diagonal factor k (CTA 0 only)
    auto factor_and_publish = [&](int k) {
        const bool bad = factor_tile(T, tid, profiling ? facc : nullptr);
        if (bad && tid == 0 && *info == 0) *info = k * kCholNB + 1;
        double* dst = Ld + (size_t)k * kCholLd;
#pragma unroll
        for (int h = 0; h < 4; ++h) dst[(warp + 8 * h) * kCholNB + lane] = T.Tb[warp + 8 * h][lane];
        dst[kCholNB * kCholNB + tid] = (&T.Dinv[0][0][0])[tid];   // 4 * 8 * 8 = 256 = kCholThreads
    };

    if (rank == 0) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int rr = warp + 8 * h;
            T.Tb[rr][lane] = lane <= rr ? S[(size_t)rr * ld + lane] : 0.0;
        }
        __syncthreads();
        factor_and_publish(0);
    }
    CHOL_MARK(1);
    cluster_barrier();
    CHOL_MARK(6);

    for (int k = 0; k < nblk; ++k) {
        const size_t kc = (size_t)k * kCholNB;
        // ---- phase A: panel k.  Diagonal factor -> shared memory (CTA 0 still holds it from its own factorisation)
        if (rank != 0) {
            const double* src = Ld + (size_t)k * kCholLd;
#pragma unroll
            for (int h = 0; h < 4; ++h) T.Tb[warp + 8 * h][lane] = src[(warp + 8 * h) * kCholNB + lane];
            (&T.Dinv[0][0][0])[tid] = src[kCholNB * kCholNB + tid];
            __syncthreads();
        }
        CHOL_MARK(0);
        const int m = nblk - (k + 1);   // trailing column blocks
        // CTA 0, look-ahead: the next diagonal tile C(k + 1, k + 1) has been final since the previous step's barrier -- fetch this warp's
        // part now (warp w: row group w / 2, column groups 2 (w % 2), + 1), its latency hides behind the panel solve
        const int drg = warp >> 1, dcg0 = 2 * (warp & 1);
        const bool don0 = dcg0 <= drg, don1 = dcg0 + 1 <= drg;
        double2 dacc[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        if (rank == 0 && m > 0) {
            const size_t r0 = (size_t)(k + 1) * kCholNB;
            const double* pc = S + (r0 + 8 * drg + g) * ld + r0 + 2 * q;
            if (don0) dacc[0] = *reinterpret_cast<const double2*>(pc + 8 * dcg0);
            if (don1) dacc[1] = *reinterpret_cast<const double2*>(pc + 8 * dcg0 + 8);
        }
        // row blocks i = k + 1 + rank, + C, ... (the right-hand side is row block nblk); a pass takes two of them: warp w solves
        // the 8 rows of group w % 4 of row block (w / 4)
        for (int i = k + 1 + rank + C * (warp >> 2); i < nrb; i += 2 * C) {
            double* rowp = S + ((size_t)i * kCholNB + 8 * (warp & 3) + g) * ld + kc + 2 * q;
            double2 x[kCholNSB];
#pragma unroll
            for (int j = 0; j < kCholNSB; ++j) x[j] = *reinterpret_cast<const double2*>(rowp + 8 * j);
            trsm_group(T, g, q, x);
#pragma unroll
            for (int j = 0; j < kCholNSB; ++j) *reinterpret_cast<double2*>(rowp + 8 * j) = x[j];
            if (rank == 0 && i == k + 1 && m > 0) {   // keep L(k + 1, k) on chip for the diagonal update
#pragma unroll
                for (int j = 0; j < kCholNSB; ++j) { Pn[8 * (warp & 3) + g][8 * j + 2 * q] = x[j].x; Pn[8 * (warp & 3) + g][8 * j + 2 * q + 1] = x[j].y; }
            }
        }
        CHOL_MARK(2);
        // the last CTA (fewest row blocks) also inverts the diagonal tile for the back-substitution: X L^T = I -> X = L^-T
        if (rank == C - 1 && warp >= 4) {
            const int grp = warp & 3;
            double2 x[kCholNSB];
#pragma unroll
            for (int j = 0; j < kCholNSB; ++j) x[j] = make_double2(j == grp && 2 * q == g ? 1.0 : 0.0, j == grp && 2 * q + 1 == g ? 1.0 : 0.0);
            trsm_group(T, g, q, x);
            double* rowp = LinvT + kc * kCholNB + (size_t)(8 * grp + g) * kCholNB + 2 * q;
#pragma unroll
            for (int j = 0; j < kCholNSB; ++j) *reinterpret_cast<double2*>(rowp + 8 * j) = x[j];
        }
        CHOL_MARK(3);
        if (rank == 0 && m > 0) {
            // ---- CTA 0: arrive at the panel barrier (its rows are stored), then -- without waiting for the others' panel rows -- update
            //      the next diagonal tile from the on-chip panel tile, factor it and publish the factor; only then wait.
            cluster_arrive();
            __syncthreads();   // Pn complete; everybody is done reading T (the current factor)
            if (don0) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double ax = -Pn[8 * drg + g][8 * s + 2 * q], ay = -Pn[8 * drg + g][8 * s + 2 * q + 1];
                    dmma884(dacc[0].x, dacc[0].y, ax, Pn[8 * dcg0 + g][8 * s + 2 * q]);
                    dmma884(dacc[0].x, dacc[0].y, ay, Pn[8 * dcg0 + g][8 * s + 2 * q + 1]);
                    if (don1) {
                        dmma884(dacc[1].x, dacc[1].y, ax, Pn[8 * dcg0 + 8 + g][8 * s + 2 * q]);
                        dmma884(dacc[1].x, dacc[1].y, ay, Pn[8 * dcg0 + 8 + g][8 * s + 2 * q + 1]);
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 8 * drg + g, c = 8 * (dcg0 + h) + 2 * q;
                T.Tb[r][c] = c <= r ? dacc[h].x : 0.0;
                T.Tb[r][c + 1] = c + 1 <= r ? dacc[h].y : 0.0;
            }
            __syncthreads();
            factor_and_publish(k + 1);
            CHOL_MARK(1);
            cluster_wait();    // now the whole panel k is visible
        } else {
            cluster_barrier();   // panel k complete and visible to the whole cluster
        }
        CHOL_MARK(4);
        if (m > 0) {
            // ---- phase B: trailing update on the FP64 tensor cores.  Tile 0 of the row-major order is the next diagonal tile, which
            //      CTA 0 has already taken care of (look-ahead above); CTA 0 takes a reduced share of the rest.
            const int nt = m * (m + 1) / 2 + m;
            const double* panel = S + kc + 2 * q;
            // contiguous ranges of equal COST per CTA (bytes moved rather than products: the update runs at about half the tensor-pipe
            // rate, bound by L1 / L2 traffic): full tile 16, diagonal tile 14, right-hand-side tile (one row group) 8
            const int ntri = m * (m + 1) / 2;
            const int ctri = 8 * m * (m - 1) + 14 * m;             // cost of the triangle; all costs fit 32 bits (m <= 160)
            auto tile_at_cost = [&](int c) -> int {                 // smallest t with cost of tiles [0, t) >= c: closed form + fix-up
                if (c >= ctri) return min(nt, ntri + ((c - ctri + 7) >> 3));
                if (c <= 0) return 0;
                int a = (int)((sqrtf(36.0f + 32.0f * (float)c) - 6.0f) * 0.0625f);   // largest a with 8 a (a - 1) + 14 a <= c
                while (8 * (a + 1) * a + 14 * (a + 1) <= c) ++a;
                while (a > 0 && 8 * a * (a - 1) + 14 * a > c) --a;
                const int bq = (c - (8 * a * (a - 1) + 14 * a) + 15) >> 4;           // full tiles of row a before the boundary
                return bq > a ? (a + 1) * (a + 2) / 2 : a * (a + 1) / 2 + bq;
            };
            int t0, t1;
            {
                const int total = ctri + 8 * m - 14;   // without tile 0 (the next diagonal tile, CTA 0's look-ahead)
                const int share0 = max(0, total / C - 16 * kCholFactorTiles);
                const int others = total - share0, per = others / (C - 1);
                if (rank == 0) { t0 = 1; t1 = max(1, tile_at_cost(14 + share0)); }
                else {
                    t0 = max(1, tile_at_cost(14 + share0 + per * (rank - 1)));
                    t1 = rank == C - 1 ? nt : max(1, tile_at_cost(14 + share0 + per * rank));
                }
            }
            // one warp per tile: 16 independent accumulator chains keep the tensor pipe busy without any block-level barrier
            for (int t = t0 + warp; t < t1; t += kCholThreads / 32) {
                int a, b;
                tile_of(t, m, a, b);
                const size_t ri = (size_t)(k + 1 + a) * kCholNB + g, rj = (size_t)(k + 1 + b) * kCholNB + g;
                update_tile_warp(panel + ri * ld, panel + rj * ld, S + ri * ld + (size_t)(k + 1 + b) * kCholNB + 2 * q, (size_t)ld, a == m ? 1 : 4, a == b,
                                 nullptr, g, q);
            }
            CHOL_MARK(5);
            cluster_barrier();   // trailing matrix updated, next diagonal factor published
            CHOL_MARK(6);
        }
    }
    if (rank != 0) {
        if (profiling) for (int i = 0; i < 8; ++i) prof[rank * 8 + i] = pacc[i];
        return;
    }
    // ---- backward substitution L^T x = y (CTA 0): y = row ld of S; x_k = L_kk^-T (y_k - sum_{i>k} L_ik^T x_i), right-looking:
    //      after x_k is known every y[m], m < 32 k, gets its contribution from row block k of L
    for (int m = tid; m < ld; m += kCholThreads) yb[m] = S[(size_t)ld * ld + m];
    __syncthreads();
    for (int k = nblk - 1; k >= 0; --k) {
        const size_t kc = (size_t)k * kCholNB;
        {   // x_k[r] = sum_c X[r][c] y_k[c], 8 threads per output
            const int r = tid >> 3, part = tid & 7;
            const double* X = LinvT + kc * kCholNB + r * kCholNB;
            double s = fma(X[part], yb[kc + part], X[part + 8] * yb[kc + part + 8]);
            s += fma(X[part + 16], yb[kc + part + 16], X[part + 24] * yb[kc + part + 24]);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            if (part == 0) { xb[r] = s; xout[kc + r] = s; }
        }
        __syncthreads();
        for (int m = tid; m < (int)kc; m += kCholThreads) {
            const double* Lc = S + kc * ld + m;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // 16 independent loads in flight, twice
                double v[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = Lc[(size_t)(16 * h + r) * ld];
#pragma unroll
                for (int r = 0; r < 16; r += 4) {
                    s0 = fma(v[r], xb[16 * h + r], s0); s1 = fma(v[r + 1], xb[16 * h + r + 1], s1);
                    s2 = fma(v[r + 2], xb[16 * h + r + 2], s2); s3 = fma(v[r + 3], xb[16 * h + r + 3], s3);
                }
            }
            yb[m] -= (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
    }
    CHOL_MARK(7);
    if (profiling) { for (int i = 0; i < 8; ++i) prof[i] = pacc[i]; for (int i = 0; i < 4; ++i) prof[15 * 8 + 8 + i] = facc[i]; }
#undef CHOL_MARK
}

int chol_max_cluster() {
    static int cached = -1;
    if (cached >= 0) return cached;
    cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int best = 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(kCholThreads); cfg.dynamicSmemBytes = 16 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n16 = 0;
    if (cudaOccupancyMaxActiveClusters(&n16, chol_cluster_kernel, &cfg) == cudaSuccess && n16 >= 1) best = 16;
    cudaGetLastError();
    cached = best;
    return cached;
}

adb_status chol_solve_launch(cudaStream_t st, double* S, int ld, int nblk, double* scratch, double* x, int* info, int cluster, const int* skip, long long* prof) {
    ADB_CHECK(nblk >= 1 && ld == nblk * kCholNB, ADB_ERR_INVALID, "chol: bad padded order");
    if (cluster <= 0) cluster = nblk >= 16 ? chol_max_cluster() : 8;   // small systems: fewer CTAs, cheaper barriers
    ADB_CHECK(cluster == 2 || cluster == 4 || cluster == 8 || cluster == 16, ADB_ERR_INVALID, "chol: cluster size %d", cluster);
    if (cluster == 16) ADB_CHECK(chol_max_cluster() == 16, ADB_ERR_INVALID, "chol: a 16-CTA cluster cannot be scheduled on this device");
    const size_t smem = (size_t)ld * sizeof(double);
    ADB_CHECK(smem <= 160 * 1024, ADB_ERR_CAPACITY, "chol: order %d exceeds the back-substitution buffer", ld);
    static size_t smem_set = 0;
    if (smem > 40 * 1024 && smem > smem_set) {
        ADB_CUDA(cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        smem_set = 160 * 1024;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(kCholThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ADB_CUDA(cudaLaunchKernelEx(&cfg, chol_cluster_kernel, S, ld, nblk, scratch, x, info, skip, prof));
    return ADB_OK;
}

}  // namespace adb

using namespace adb;

extern "C" adb_status adb_dense_solve(int32_t device, int32_t n, const double* A, const double* b, double* x, int32_t* info, int32_t cluster,
                                      int32_t reps, float* ms_per_solve) {
    ADB_CHECK(n >= 1 && A && b && x && info, ADB_ERR_INVALID, "null argument / empty system");
    adb_status st = select_device(device);
    if (st != ADB_OK) return st;
    const int nblk = (n + kCholNB - 1) / kCholNB, ld = nblk * kCholNB;
    const size_t rows = (size_t)ld + kCholNB, elems = rows * ld;
    std::vector<double> h(elems, 0.0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c <= r; ++c) h[(size_t)r * ld + c] = A[(size_t)r * n + c];
    for (int r = n; r < ld; ++r) h[(size_t)r * ld + r] = 1.0;
    for (int c = 0; c < n; ++c) h[(size_t)ld * ld + c] = b[c];
    double *d0 = nullptr, *d1 = nullptr, *dinv = nullptr, *dx = nullptr;
    int* dinfo = nullptr;
    long long* dprof = nullptr;
    const bool want_prof = getenv("ADB_CHOL_PROFILE") != nullptr;
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&] {
        cudaFree(d0); cudaFree(d1); cudaFree(dinv); cudaFree(dx); cudaFree(dinfo); cudaFree(dprof);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (s) cudaStreamDestroy(s);
    };
#define TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } } while (0)
    TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    TRY(cudaEventCreate(&e0)); TRY(cudaEventCreate(&e1));
    TRY(cudaMalloc(&d0, elems * 8)); TRY(cudaMalloc(&d1, elems * 8)); TRY(cudaMalloc(&dinv, chol_scratch_elems(n) * 8));
    TRY(cudaMalloc(&dx, (size_t)ld * 8)); TRY(cudaMalloc(&dinfo, 4));
    if (want_prof) { TRY(cudaMalloc(&dprof, (16 * 8 + 4) * sizeof(long long))); TRY(cudaMemset(dprof, 0, (16 * 8 + 4) * sizeof(long long))); }
    TRY(cudaMemcpyAsync(d0, h.data(), elems * 8, cudaMemcpyHostToDevice, s));
    float total = 0.f;
    const int R = std::max(1, (int)reps);
    for (int it = 0; it < R; ++it) {
        TRY(cudaMemcpyAsync(d1, d0, elems * 8, cudaMemcpyDeviceToDevice, s));
        TRY(cudaMemsetAsync(dinfo, 0, 4, s));
        TRY(cudaEventRecord(e0, s));
        st = chol_solve_launch(s, d1, ld, nblk, dinv, dx, dinfo, cluster, nullptr, dprof);
        if (st != ADB_OK) { cleanup(); return st; }
        TRY(cudaEventRecord(e1, s));
        TRY(cudaStreamSynchronize(s));
        float ms = 0.f;
        TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 || R == 1) total += ms;   // the first repetition warms up
    }
    std::vector<double> hx(ld);
    int hinfo = 0;
    TRY(cudaMemcpy(hx.data(), dx, (size_t)ld * 8, cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(&hinfo, dinfo, 4, cudaMemcpyDeviceToHost));
    if (want_prof) {   // developer aid: per-CTA phase cycles of the last repetition on stderr
        long long hp[16 * 8 + 4];
        TRY(cudaMemcpy(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost));
        static const char* names[8] = {"load", "factor", "panel", "inverse", "barrier1", "update", "barrier2", "backsub"};
        for (int r : {0, 1, 7, 15}) {
            fprintf(stderr, "[chol n=%d] cta %2d:", n, r);
            for (int i = 0; i < 8; ++i) fprintf(stderr, " %s %.1f", names[i], hp[r * 8 + i] / 1000.0);
            fprintf(stderr, " kcycles\n");
        }
        fprintf(stderr, "[chol n=%d] inside factor_tile (cta 0): chains %.1f barriers %.1f column-update %.1f inverses %.1f kcycles\n", n, hp[128] / 1000.0,
                hp[129] / 1000.0, hp[130] / 1000.0, hp[131] / 1000.0);
    }
    for (int c = 0; c < n; ++c) x[c] = hx[c];
    *info = hinfo;
    if (ms_per_solve) *ms_per_solve = total / (float)std::max(1, R - 1 + (R == 1));
#undef TRY
    cleanup();
    return ADB_OK;
}
