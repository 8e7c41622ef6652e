// ba.cu -- sparse bundle adjustment on sm_100a behind adb_ba_solve (include/airdos_b200.h).
//
// Replaces Optimizer::LocalBundleAdjustment / LocalBundleAdjustmentHumanTrajactory + the g2o
// machinery under them (SparseOptimizer, BlockSolver, OptimizationAlgorithmLevenberg, the edge
// and vertex types) with flat arrays and a handful of kernels per LM trial:
//
//   ba_linearize_kernel     computeError + linearizeOplus + constructQuadraticForm for the
//                           reprojection edges (edges sorted by map point): Hll / bl, one 6x3
//                           Hpl block per edge, Hpp / bp, robust chi2
//   ba_dyn_linearize_kernel the same for joint reprojection, rigidity and motion edges, all of
//                           which live in the dense (non-marginalised) block
//   ba_dinv_kernel          (Hll + lambda I)^-1 per point, bschur -= Hpl Dinv bl
//   ba_schur_kernel         Hschur -= Hpl_i Dinv Hpl_j^T over the per-point edge pairs
//   chol_cluster_kernel     (chol.cu) dense FP64 Cholesky solve of the reduced system in ONE launch of one thread-block
//                           cluster: DMMA trailing update / panel solve, look-ahead factorisation, substitutions
//   ba_pose_update_kernel   exp(dx) * T, additive / right-multiplicative updates of the rest
//   ba_backsub_kernel       xl = Dinv (bl - Hpl^T xp), trial points, landmark part of the gain ratio
//   ba_eval_kernel          residuals + robust chi2 of the trial state
//
// State is double buffered (current / trial): accepting an LM step swaps two pointers, rejecting
// it costs nothing (g2o's push / pop copies every vertex).  The LM decision itself is taken on
// the host from one 48-byte read-back per trial.
// The landmark and pose blocks are accumulated without atomics (fixed order); the Schur chunks, the chi2 sums and the few
// dense-block edges use FP64 atomics in L2 (RED.ADD.F64): those sums are order-dependent at the 1e-16 level, far below
// the 1e-4 parity bar.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <vector>

#include "chol.cuh"

namespace adb {

// ----------------------------------------------------------------------------------------
// small device math (same formulas as oracle/ba_oracle.cpp; Eigen conventions, q = x,y,z,w)
__host__ __device__ inline void quat_to_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__host__ __device__ inline void rot_to_quat(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        double qq[3];
        qq[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        qq[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        qq[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2];
    }
}
__host__ __device__ inline void quat_normalize_pos(double* q) {
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
__host__ __device__ inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// VertexSE3Expmap::oplusImpl: T <- exp(d) T  (se3quat.h:217-257, 104-110)
__device__ inline void pose_oplus(const double* q, const double* t, const double* d, double* qo, double* to) {
    const double wx = d[0], wy = d[1], wz = d[2];
    const double theta = sqrt(wx * wx + wy * wy + wz * wz);
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3.0);
        for (int i = 0; i < 9; ++i) {
            R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
            V[i] = (i % 4 == 0 ? 1.0 : 0.0) + b * O[i] + c * O2[i];
        }
    }
    double qe[4], te[3], Re[9];
    rot_to_quat(R, qe);
    quat_normalize_pos(qe);
    for (int i = 0; i < 3; ++i) te[i] = V[i * 3] * d[3] + V[i * 3 + 1] * d[4] + V[i * 3 + 2] * d[5];
    quat_to_rot(qe, Re);
    for (int i = 0; i < 3; ++i) to[i] = te[i] + Re[i * 3] * t[0] + Re[i * 3 + 1] * t[1] + Re[i * 3 + 2] * t[2];
    const double ax = qe[0], ay = qe[1], az = qe[2], aw = qe[3], bx = q[0], by = q[1], bz = q[2], bw = q[3];
    qo[3] = aw * bw - ax * bx - ay * by - az * bz;
    qo[0] = aw * bx + ax * bw + ay * bz - az * by;
    qo[1] = aw * by + ay * bw + az * bx - ax * bz;
    qo[2] = aw * bz + az * bw + ax * by - ay * bx;
    quat_normalize_pos(qo);
}
// VertexSE3::oplusImpl (include/g2o_vertex_se3.h:113-122): H <- H * (t = d[0:3], compact quaternion d[3:6])
__device__ inline void motion_oplus(const double* q, const double* t, const double* d, double* qo, double* to) {
    double R[9], Ri[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Rn[9];
    quat_to_rot(q, R);
    const double w = 1 - (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
    if (!(w < 0)) {
        const double qi[4] = {d[3], d[4], d[5], sqrt(w)};
        quat_to_rot(qi, Ri);
    }
    mat3_mul(R, Ri, Rn);
    for (int i = 0; i < 3; ++i) to[i] = t[i] + R[i * 3] * d[0] + R[i * 3 + 1] * d[1] + R[i * 3 + 2] * d[2];
    rot_to_quat(Rn, qo);
    const double n = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
    qo[0] /= n; qo[1] /= n; qo[2] /= n; qo[3] /= n;
}

struct Cam { double fx, fy, cx, cy, bf; };

// residual of Edge(Stereo)SE3ProjectXYZ (types_six_dof_expmap.cpp:141-157): float invz and float bf in the stereo model
__device__ inline int reproj_error(const Cam& C, const double* R, const double* t, const double* X, const double* obs, double* e, double* Xc) {
    for (int i = 0; i < 3; ++i) Xc[i] = R[i * 3] * X[0] + R[i * 3 + 1] * X[1] + R[i * 3 + 2] * X[2] + t[i];
    if (obs[2] >= 0) {
        const float invz = (float)(1.0 / Xc[2]);
        const float bf = (float)C.bf;
        const double u = Xc[0] * invz * C.fx + C.cx, v = Xc[1] * invz * C.fy + C.cy;
        const double ur = u - (double)__fmul_rn(bf, invz);
        e[0] = obs[0] - u; e[1] = obs[1] - v; e[2] = obs[2] - ur;
        return 3;
    }
    e[0] = obs[0] - (Xc[0] / Xc[2] * C.fx + C.cx);
    e[1] = obs[1] - (Xc[1] / Xc[2] * C.fy + C.cy);
    e[2] = 0;
    return 2;
}
// linearizeOplus of Edge(Stereo)SE3ProjectXYZ (types_six_dof_expmap.cpp:188-234, 159-185).  The reference divides by z / z^2 in
// every term (28 FP64 divisions per edge); here 1/z and 1/z^2 are formed once and multiplied in: each term differs from the
// quotient form by at most one rounding (1e-16 relative), far below the 1e-4 parity bar, and the kernels lose a third of
// their instructions.
__device__ inline void reproj_jacobians(const Cam& C, const double* R, const double* Xc, int dim, double* Ji, double* Jj) {
    const double x = Xc[0], y = Xc[1], z = Xc[2], iz = 1.0 / z, iz2 = iz * iz, fx = C.fx, fy = C.fy, bf = C.bf;
    for (int c = 0; c < 3; ++c) {
        Ji[c] = -fx * R[c] * iz + fx * x * R[6 + c] * iz2;
        Ji[3 + c] = -fy * R[3 + c] * iz + fy * y * R[6 + c] * iz2;
        Ji[6 + c] = dim == 3 ? Ji[c] - bf * R[6 + c] * iz2 : 0.0;
    }
    Jj[0] = x * y * iz2 * fx; Jj[1] = -(1 + (x * x * iz2)) * fx; Jj[2] = y * iz * fx; Jj[3] = -iz * fx; Jj[4] = 0; Jj[5] = x * iz2 * fx;
    Jj[6] = (1 + y * y * iz2) * fy; Jj[7] = -x * y * iz2 * fy; Jj[8] = -x * iz * fy; Jj[9] = 0; Jj[10] = -iz * fy; Jj[11] = y * iz2 * fy;
    if (dim == 3) {
        Jj[12] = Jj[0] - bf * y * iz2; Jj[13] = Jj[1] + bf * x * iz2; Jj[14] = Jj[2]; Jj[15] = Jj[3]; Jj[16] = 0; Jj[17] = Jj[5] - bf * iz2;
    } else {
        for (int i = 12; i < 18; ++i) Jj[i] = 0;
    }
}
// RobustKernelHuber::robustify (Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-92).  The reference keeps delta^2 in a FLOAT member
// (core/robust_kernel_impl.h:84 `float dsqr;`, set by setDelta :65-69 from the double product): both the inlier test and rho(e) use the
// rounded value, pinned by tests/test_ref_lm.py against the compiled reference function.
__device__ inline void huber(double delta, bool robust, double e2, double* rho0, double* rho1) {
    const double dsqr = (double)(float)(delta * delta);
    if (!robust || e2 <= dsqr) { *rho0 = e2; *rho1 = 1.0; }
    else { const double s = sqrt(e2); *rho0 = 2 * s * delta - dsqr; *rho1 = delta / s; }
}

// sums the kernels of one LM step produce for the device-side LM controller
struct Scalars {
    double chi_cur, chi_trial, scale, maxdiag;
    int info, pad;
};

// OptimizationAlgorithmLevenberg::solve + SparseOptimizer::optimize as a device-resident state machine
// (optimization_algorithm_levenberg.cpp:61-189, sparse_optimizer.cpp:optimize): every kernel of a step reads its lambda, its
// state buffers (cur = accepted state, cur ^ 1 = trial) and whether it has anything to do from here, so the host can enqueue
// whole rounds without waiting for a single accept / reject decision.
struct Lm {
    double lambda, ni, current, ini, tau, chi2_initial;
    int cur;              // state buffers holding the accepted state
    int chi_last;         // chi2 buffers written by the last evaluation (the gates read these: stale-error semantics)
    int need_lin;         // 1: the next step opens an outer iteration (buildSystem first)
    int done;             // 1: the round is over, every later kernel of the batch returns at once
    int first;            // 1: first iteration of the round (computeLambdaInit)
    int it, it_max, q, max_trials, nbad;
    int iterations_run, trials_run, trace_len, trace_cap, round, pad;
};

struct Opt {
    double huber_mono, huber_stereo, huber_rigid, huber_motion;
    int robust;
};

__device__ inline double block_sum_to(double v, double* target) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0 && v != 0.0) atomicAdd(target, v);
    }
    return v;
}

// ----------------------------------------------------------------------------------------
// Reprojection edges pose <-> marginalised point, one thread per edge.
struct StaticEdges {
    int n;
    const int* pose; const int* point; const double* obs; const double* info; const uint8_t* level;
};
struct State {
    const double* pq; const double* pt; const double* X; const double* J; const double* D; const double* mq; const double* mt;
};
struct StatePair { State s[2]; };
struct ChiPair { double* e[2]; double* j[2]; double* r[2]; double* m[2]; };   // chi2 per edge family, double buffered like the state

// sums `nv` doubles per thread over the block in a fixed order; result in out[0..nv) (shared), valid after the call
template <int NV>
__device__ inline void block_reduce_fixed(double (&v)[NV], double* scratch /*[8][NV]*/, double* out /*[NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
        if (lane == 0) scratch[warp * NV + k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += scratch[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

constexpr int kBaThreads = 128;

// buildSystem, landmark side: kPtLanes lanes per map point, one edge per lane (a point's edges are contiguous): residual,
// Jacobians, Huber weight, one 6x3 Hpl block stored per edge; Hll / bl are summed over the lane group by shuffles in a fixed
// order and stored once (no atomics).  Thread-per-point left 93 % of the warp slots idle (157 CTAs, 0.27 waves).
constexpr int kPtLanes = 8;
__global__ void __launch_bounds__(kBaThreads) ba_point_linearize_kernel(Cam C, Opt O, StaticEdges E, StatePair SP, const Lm* __restrict__ lm,
                                                                       const int* __restrict__ point_ptr, int np, const int* __restrict__ off_pose,
                                                                       double* __restrict__ Hll, double* __restrict__ bl, double* __restrict__ W,
                                                                       ChiPair chi, Scalars* sc) {
    if (lm->done || !lm->need_lin) return;
    const State S = SP.s[lm->cur];
    double* __restrict__ chi_e = chi.e[lm->cur];
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) / kPtLanes, sub = threadIdx.x % kPtLanes;
    double rho_sum = 0;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // Hll (6 unique) | bl (3)
    if (l < np) {
        const double X[3] = {S.X[3 * (size_t)l], S.X[3 * (size_t)l + 1], S.X[3 * (size_t)l + 2]};
        for (int e = point_ptr[l] + sub; e < point_ptr[l + 1]; e += kPtLanes) {
            if (E.level[e]) continue;
            const int ip = E.pose[e];
            double R[9], er[3], Xc[3], Ji[9], Jj[18], rho0, rho1;
            quat_to_rot(S.pq + 4 * ip, R);
            const int dim = reproj_error(C, R, S.pt + 3 * ip, X, E.obs + 3 * (size_t)e, er, Xc);
            reproj_jacobians(C, R, Xc, dim, Ji, Jj);
            const double w0 = E.info[e];
            const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
            chi_e[e] = c;
            huber(dim == 3 ? O.huber_stereo : O.huber_mono, O.robust, c, &rho0, &rho1);
            rho_sum += rho0;
            const double w = rho1 * w0;
            int u = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = i; j < 3; ++j, ++u) {
                    double s = 0;
                    for (int k = 0; k < dim; ++k) s += Ji[k * 3 + i] * w * Ji[k * 3 + j];
                    acc[u] += s;
                }
                double s = 0;
                for (int k = 0; k < dim; ++k) s += Ji[k * 3 + i] * (-w0 * er[k] * rho1);
                acc[6 + i] += s;
            }
            if (off_pose[ip] >= 0) {
                double* We = W + 18 * (size_t)e;
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        double s2 = 0;
                        for (int k = 0; k < dim; ++k) s2 += Jj[k * 6 + i] * w * Ji[k * 3 + j];
                        We[i * 3 + j] = s2;
                    }
            }
        }
    }
    // fixed-order tree over the lane group (all 32 lanes take part: groups are aligned)
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int o = 1; o < kPtLanes; o <<= 1) acc[k] += __shfl_xor_sync(0xFFFFFFFFu, acc[k], o);
    if (l < np && sub == 0) {
#pragma unroll
        for (int u = 0; u < 6; ++u) Hll[6 * (size_t)l + u] = acc[u];
#pragma unroll
        for (int i = 0; i < 3; ++i) bl[3 * (size_t)l + i] = acc[6 + i];
    }
    block_sum_to(rho_sum, &sc->chi_cur);
}

// buildSystem, pose side: kPoseChunks CTAs per free pose, each over a contiguous slice of the pose's edge list (built once per solve):
// J_pose^T (w Omega) J_pose and -J_pose^T w Omega e summed with a fixed-order block reduction into a partial record; the reduce
// kernel adds the partials in a fixed order and stores the 6x6 block: no atomics, deterministic.  (One CTA per pose was 49 CTAs
// on 148 SMs with ~10 edges per thread in sequence.)
// Independent of ba_point_linearize_kernel (it re-evaluates the edge's chi2 itself): the two run concurrently on two streams and
// join before the dense-edge kernel adds to H atomically.
constexpr int kPoseChunks = 8;
__global__ void __launch_bounds__(256) ba_pose_linearize_kernel(Cam C, Opt O, StaticEdges E, StatePair SP, const Lm* __restrict__ lm,
                                                               const int* __restrict__ pose_ptr, const int* __restrict__ pose_edges,
                                                               const int* __restrict__ free_pose, double* __restrict__ partial) {
    __shared__ double scratch[8 * 27], red[27];
    if (lm->done || !lm->need_lin) return;
    const State S = SP.s[lm->cur];
    const int fp = blockIdx.x / kPoseChunks, ch = blockIdx.x % kPoseChunks;
    const int ip = free_pose[fp];
    double v[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) v[k] = 0;
    double R[9];
    quat_to_rot(S.pq + 4 * ip, R);
    const double t[3] = {S.pt[3 * ip], S.pt[3 * ip + 1], S.pt[3 * ip + 2]};
    const int lo = pose_ptr[ip], n = pose_ptr[ip + 1] - lo, per = (n + kPoseChunks - 1) / kPoseChunks;
    const int q0 = lo + ch * per, q1 = min(lo + n, q0 + per);
    for (int q = q0 + threadIdx.x; q < q1; q += 256) {
        const int e = pose_edges[q];
        if (E.level[e]) continue;
        double er[3], Xc[3], Ji[9], Jj[18], rho0, rho1;
        const int dim = reproj_error(C, R, t, S.X + 3 * (size_t)E.point[e], E.obs + 3 * (size_t)e, er, Xc);
        reproj_jacobians(C, R, Xc, dim, Ji, Jj);
        const double w0 = E.info[e];
        // the same expression as the landmark-side kernel evaluates (bit for bit): the two kernels run concurrently on two streams
        const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
        huber(dim == 3 ? O.huber_stereo : O.huber_mono, O.robust, c, &rho0, &rho1);
        const double w = rho1 * w0;
        int u = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int c = 0; c <= r; ++c, ++u) {
                double s = 0;
                for (int k = 0; k < dim; ++k) s += Jj[k * 6 + r] * w * Jj[k * 6 + c];
                v[u] += s;
            }
            double s = 0;
            for (int k = 0; k < dim; ++k) s += Jj[k * 6 + r] * (-w0 * er[k] * rho1);
            v[21 + r] += s;
        }
    }
    block_reduce_fixed<27>(v, scratch, red);
    if (threadIdx.x < 27) partial[(size_t)blockIdx.x * 27 + threadIdx.x] = red[threadIdx.x];
}
__global__ void ba_pose_reduce_kernel(const double* __restrict__ partial, const int* __restrict__ free_pose, const int* __restrict__ off_pose, int nd,
                                      const Lm* __restrict__ lm, double* __restrict__ H, double* __restrict__ b) {
    if (lm->done || !lm->need_lin) return;
    const int fp = blockIdx.x, k = threadIdx.x;   // 32 threads, 27 active
    if (k >= 27) return;
    double sum = 0;
    for (int c = 0; c < kPoseChunks; ++c) sum += partial[((size_t)fp * kPoseChunks + c) * 27 + k];
    const int op = off_pose[free_pose[fp]];
    if (k < 21) {
        int r = 0, acc = 0;
        while (acc + r + 1 <= k) { acc += r + 1; ++r; }
        H[(size_t)(op + r) * nd + op + (k - acc)] = sum;
    } else {
        b[op + k - 21] = sum;
    }
}

// Dense-block edges: joint reprojection (type 0), rigidity (1), motion (2); one thread per edge.
struct DynEdges {
    int nj, nr, nm;
    const int* j_pose; const int* j_joint; const double* j_obs; const double* j_info; const uint8_t* j_level;
    const int* r_i; const int* r_j; const int* r_d; const double* r_info; const uint8_t* r_level;
    const int* m_p1; const int* m_p2; const int* m_m; const double* m_dt; const double* m_info; const uint8_t* m_level;
};
struct DynOff { const int* pose; const int* joint; const int* dist; const int* motion; };

// H(oa.., ob..) += A^T w B restricted to the lower triangle (row >= col); blocks are dim x da / dim x db row-major
__device__ inline void add_block_lower(double* H, int nd, int oa, int da, const double* A, int ob, int db, const double* B, int dim, double w) {
    if (oa < 0 || ob < 0) return;
    for (int i = 0; i < da; ++i)
        for (int j = 0; j < db; ++j) {
            if (oa + i < ob + j) continue;
            double s = 0;
            for (int k = 0; k < dim; ++k) s += A[k * da + i] * w * B[k * db + j];
            if (s != 0.0) atomicAdd(H + (size_t)(oa + i) * nd + ob + j, s);
        }
}
__device__ inline void add_rhs(double* b, int oa, int da, const double* A, int dim, const double* wr) {
    if (oa < 0) return;
    for (int i = 0; i < da; ++i) {
        double s = 0;
        for (int k = 0; k < dim; ++k) s += A[k * da + i] * wr[k];
        if (s != 0.0) atomicAdd(b + oa + i, s);
    }
}
// LandmarkMotionTernaryEdge::computeError, zero measurement (include/g2o_dyn_slam3d.h:65-76): e = p1 - M^-1 p2, M = (R, dt t)
__device__ inline void motion_error_qt(const double* mq, const double* mt, double dt, const double* p1, const double* p2, double* er, double* Rm) {
    quat_to_rot(mq, Rm);
    const double d[3] = {p2[0] - dt * mt[0], p2[1] - dt * mt[1], p2[2] - dt * mt[2]};
    for (int i = 0; i < 3; ++i) er[i] = p1[i] - (Rm[i] * d[0] + Rm[3 + i] * d[1] + Rm[6 + i] * d[2]);
}
__device__ inline void motion_error(const State& S, int m, double dt, const double* p1, const double* p2, double* er, double* Rm) {
    motion_error_qt(S.mq + 4 * m, S.mt + 3 * m, dt, p1, p2, er, Rm);
}
// EdgeRigidBodyDouble::computeError (include/g2o_edge_rigidbody.h:139-149): |a - b| - d; d3 / n return the difference and its norm
__device__ inline double rigid_error(const double* a, const double* b, double dist, double* d3, double* n) {
    d3[0] = a[0] - b[0]; d3[1] = a[1] - b[1]; d3[2] = a[2] - b[2];
    *n = sqrt(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
    return *n - dist;
}

// mode 0: linearise into H / b and accumulate chi_cur; mode 1: evaluate only, accumulate chi_trial
__global__ void __launch_bounds__(kBaThreads) ba_dyn_kernel(Cam C, Opt O, DynEdges E, StatePair SP, const Lm* __restrict__ lm, DynOff off, int nd,
                                                           double* __restrict__ H, double* __restrict__ b, ChiPair chi, Scalars* sc, int mode) {
    if (lm->done || (mode == 0 && !lm->need_lin)) return;
    const int which = mode == 0 ? lm->cur : lm->cur ^ 1;
    const State S = SP.s[which];
    double* __restrict__ chi_j = chi.j[which]; double* __restrict__ chi_r = chi.r[which]; double* __restrict__ chi_m = chi.m[which];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double rho0 = 0, rho1 = 1;
    if (mode == 1) {   // level-1 edges keep their last chi2 (the trial buffers inherit it)
        if (t < E.nj) { if (E.j_level[t]) chi_j[t] = chi.j[which ^ 1][t]; }
        else if (t < E.nj + E.nr) { if (E.r_level[t - E.nj]) chi_r[t - E.nj] = chi.r[which ^ 1][t - E.nj]; }
        else if (t < E.nj + E.nr + E.nm) { if (E.m_level[t - E.nj - E.nr]) chi_m[t - E.nj - E.nr] = chi.m[which ^ 1][t - E.nj - E.nr]; }
    }
    if (t < E.nj) {
        const int e = t;
        if (!E.j_level[e]) {
            const int ip = E.j_pose[e], ij = E.j_joint[e];
            double R[9], er[3], Xc[3];
            quat_to_rot(S.pq + 4 * ip, R);
            const int dim = reproj_error(C, R, S.pt + 3 * ip, S.J + 3 * ij, E.j_obs + 3 * e, er, Xc);
            const double w0 = E.j_info[e];
            const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
            chi_j[e] = c;
            huber(dim == 3 ? O.huber_stereo : O.huber_mono, O.robust, c, &rho0, &rho1);
            if (mode == 0) {
                double Ji[9], Jj[18];
                reproj_jacobians(C, R, Xc, dim, Ji, Jj);
                const double w = rho1 * w0;
                const double wr[3] = {-w0 * er[0] * rho1, -w0 * er[1] * rho1, -w0 * er[2] * rho1};
                const int op = off.pose[ip], oj = off.joint[ij];
                add_block_lower(H, nd, oj, 3, Ji, oj, 3, Ji, dim, w); add_rhs(b, oj, 3, Ji, dim, wr);
                add_block_lower(H, nd, op, 6, Jj, op, 6, Jj, dim, w); add_rhs(b, op, 6, Jj, dim, wr);
                add_block_lower(H, nd, op, 6, Jj, oj, 3, Ji, dim, w); add_block_lower(H, nd, oj, 3, Ji, op, 6, Jj, dim, w);
            }
        }
    } else if (t < E.nj + E.nr) {
        const int e = t - E.nj;
        if (!E.r_level[e]) {
            const int i1 = E.r_i[e], i2 = E.r_j[e], id = E.r_d[e];
            double d[3], n;
            const double er = rigid_error(S.J + 3 * i1, S.J + 3 * i2, S.D[id], d, &n);
            const double w0 = E.r_info[e];
            const double c = er * (w0 * er);
            chi_r[e] = c;
            huber(O.huber_rigid, O.robust, c, &rho0, &rho1);
            if (mode == 0) {
                const double w = rho1 * w0;
                const double wr[1] = {-w0 * er * rho1};
                double Ja[3] = {0, 0, 0}, Jb[3] = {0, 0, 0};
                if (n >= 1e-12) for (int k = 0; k < 3; ++k) { Ja[k] = d[k] / n; Jb[k] = -d[k] / n; }
                const double Jd[1] = {-1.0};
                const int offs[3] = {off.joint[i1], off.joint[i2], off.dist[id]};
                const int dims[3] = {3, 3, 1};
                const double* Js[3] = {Ja, Jb, Jd};
                for (int u = 0; u < 3; ++u) {
                    add_rhs(b, offs[u], dims[u], Js[u], 1, wr);
                    for (int v = 0; v < 3; ++v) add_block_lower(H, nd, offs[u], dims[u], Js[u], offs[v], dims[v], Js[v], 1, w);
                }
            }
        }
    } else if (t < E.nj + E.nr + E.nm) {
        const int e = t - E.nj - E.nr;
        if (!E.m_level[e]) {
            double er[3], Rm[9];
            const int m = E.m_m[e];
            motion_error(S, m, E.m_dt[e], S.J + 3 * E.m_p1[e], S.J + 3 * E.m_p2[e], er, Rm);
            const double w0 = E.m_info[e];
            const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
            chi_m[e] = c;
            huber(O.huber_motion, O.robust, c, &rho0, &rho1);
            if (mode == 0) {
                const double w = rho1 * w0;
                const double wr[3] = {-w0 * er[0] * rho1, -w0 * er[1] * rho1, -w0 * er[2] * rho1};
                const double J1[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                double J2[9], Jm[18];
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J2[i * 3 + j] = -Rm[j * 3 + i];
                for (int i = 0; i < 18; ++i) Jm[i] = 0;
                Jm[0] = E.m_dt[e]; Jm[7] = E.m_dt[e]; Jm[14] = E.m_dt[e];   // convention D.6
                const int offs[3] = {off.joint[E.m_p1[e]], off.joint[E.m_p2[e]], off.motion[m]};
                const int dims[3] = {3, 3, 6};
                const double* Js[3] = {J1, J2, Jm};
                for (int u = 0; u < 3; ++u) {
                    add_rhs(b, offs[u], dims[u], Js[u], 3, wr);
                    for (int v = 0; v < 3; ++v) add_block_lower(H, nd, offs[u], dims[u], Js[u], offs[v], dims[v], Js[v], 3, w);
                }
            }
        }
    }
    block_sum_to(rho0, mode == 0 ? &sc->chi_cur : &sc->chi_trial);
}

// max |diag| over the dense block and the landmark blocks (computeLambdaInit, levenberg.cpp:166-180)
__global__ void ba_maxdiag_kernel(const double* __restrict__ H, int nd, const double* __restrict__ Hll, const uint8_t* __restrict__ act_point,
                                  int np, const Lm* __restrict__ lm, Scalars* sc) {
    if (lm->done || !lm->need_lin || !lm->first) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0;
    if (i < nd) m = fabs(H[(size_t)i * nd + i]);
    else if (i - nd < np && act_point[i - nd]) {
        const double* h = Hll + 6 * (size_t)(i - nd);
        m = fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5])));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0)
        atomicMax(reinterpret_cast<unsigned long long*>(&sc->maxdiag), (unsigned long long)__double_as_longlong(m));   // m >= 0: order preserved
}

// S = H (lower) + lambda I in the padded layout of the cluster Cholesky (chol.cuh: pitch ld = 32 ceil(nd / 32), identity pad,
// right-hand side in row ld), bs = b
__global__ void ba_prepare_kernel(const double* __restrict__ H, const double* __restrict__ b, int nd, int ld, const Lm* __restrict__ lm,
                                  double* __restrict__ Sm, double* __restrict__ bs) {
    if (lm->done) return;
    const double lambda = lm->lambda;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)ld * ld) {
        const int r = (int)(i / ld), c = (int)(i - (size_t)r * ld);
        Sm[i] = (r < nd && c < nd) ? H[(size_t)r * nd + c] + (r == c ? lambda : 0.0) : (r == c ? 1.0 : 0.0);
    }
    if (i < (size_t)ld) { const double v = i < (size_t)nd ? b[i] : 0.0; bs[i] = v; Sm[(size_t)ld * ld + i] = v; }   // row ld carries the right-hand side through the factorisation
    if (i < (size_t)(kCholNB - 1) * ld) Sm[(size_t)(ld + 1) * ld + i] = 0.0;                                          // spare rows of the right-hand-side row block
}

// per point: Dinv = (Hll + lambda I)^-1 by cofactors (Eigen fixed-size inverse), db = Dinv bl
__global__ void ba_dinv_kernel(const double* __restrict__ Hll, const double* __restrict__ bl, const uint8_t* __restrict__ act_point, int np,
                               const Lm* __restrict__ lm, double* __restrict__ Dinv, double* __restrict__ db) {
    if (lm->done) return;
    const double lambda = lm->lambda;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= np || !act_point[l]) return;
    const double* h = Hll + 6 * (size_t)l;
    const double A[9] = {h[0] + lambda, h[1], h[2], h[1], h[3] + lambda, h[4], h[2], h[4], h[5] + lambda};
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    double B[9];
    B[0] = c00 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = c01 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = c02 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    for (int i = 0; i < 9; ++i) Dinv[9 * (size_t)l + i] = B[i];
    for (int i = 0; i < 3; ++i) db[3 * (size_t)l + i] = B[i * 3] * bl[3 * l] + B[i * 3 + 1] * bl[3 * l + 1] + B[i * 3 + 2] * bl[3 * l + 2];
}

// Schur complement by destination block: the pair list is sorted by (pose block i, pose block j) on the host and cut into
// chunks of <= 128 pairs; one warp sums a chunk in registers, reduces with shuffles and issues 36 (+6) atomics per chunk
// instead of 36 per pair.
constexpr int kSchurChunk = 128;
__global__ void __launch_bounds__(kBaThreads) ba_schur_block_kernel(const int2* __restrict__ pairs, const int2* __restrict__ chunks, const int* __restrict__ totals,
                                                                   const int* __restrict__ e_pose, const int* __restrict__ e_point,
                                                                   const int* __restrict__ off_pose, const double* __restrict__ W,
                                                                   const double* __restrict__ Dinv, const double* __restrict__ db, int nd,
                                                                   const Lm* __restrict__ lm, double* __restrict__ Sm, double* __restrict__ bs) {
    if (lm->done) return;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (kBaThreads / 32) + (threadIdx.x >> 5);
    if (c >= totals[1]) return;   // the grid is sized for the host's upper bound of the chunk count
    const int2 ch = chunks[c];
    double acc[36], rhs[6];
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) rhs[i] = 0.0;
    for (int p = ch.x + lane; p < ch.x + ch.y; p += 32) {
        const int2 pr = pairs[p];
        const int l = e_point[pr.x];
        const double* W1 = W + 18 * (size_t)pr.x;
        const double* W2 = W + 18 * (size_t)pr.y;
        const double* Di = Dinv + 9 * (size_t)l;
        double BD[18], w2[18];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) BD[i * 3 + j] = W1[i * 3] * Di[j] + W1[i * 3 + 1] * Di[3 + j] + W1[i * 3 + 2] * Di[6 + j];
#pragma unroll
        for (int i = 0; i < 18; ++i) w2[i] = W2[i];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[i * 6 + j] += BD[i * 3] * w2[j * 3] + BD[i * 3 + 1] * w2[j * 3 + 1] + BD[i * 3 + 2] * w2[j * 3 + 2];
        if (pr.x == pr.y) {
            const double* d = db + 3 * (size_t)l;
#pragma unroll
            for (int i = 0; i < 6; ++i) rhs[i] += W1[i * 3] * d[0] + W1[i * 3 + 1] * d[1] + W1[i * 3 + 2] * d[2];
        }
    }
#pragma unroll
    for (int i = 0; i < 36; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xFFFFFFFFu, acc[i], o);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rhs[i] += __shfl_xor_sync(0xFFFFFFFFu, rhs[i], o);
    const int2 p0 = pairs[ch.x];
    const int o1 = off_pose[e_pose[p0.x]], o2 = off_pose[e_pose[p0.y]];
    // lanes 0..35 own one entry of the 6 x 6 block each (static register indexing via the unrolled select)
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < 36; ++i) if (lane == i) mine = acc[i];
    if (lane < 32) {
        const int i = lane / 6, j = lane % 6;
        if (!(o1 == o2 && j > i)) atomicAdd(Sm + (size_t)(o1 + i) * nd + o2 + j, -mine);
    }
    if (lane < 4) {   // entries 32..35
        const int e = 32 + lane, i = e / 6, j = e % 6;
        double v = 0.0;
#pragma unroll
        for (int k = 32; k < 36; ++k) if (e == k) v = acc[k];
        if (!(o1 == o2 && j > i)) atomicAdd(Sm + (size_t)(o1 + i) * nd + o2 + j, -v);
    }
    if (lane >= 8 && lane < 14) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (lane - 8 == k) v = rhs[k];
        if (v != 0.0) atomicAdd(bs + o1 + (lane - 8), -v);
    }
}

// dense updates: poses (exp), bone lengths (+), motions (right multiply), joints (+); also the dense
// part of the gain-ratio denominator sum x (lambda x + b)
struct DenseSizes { int n_poses, n_dists, n_motions, n_joints; };
__device__ __forceinline__ void dense_update_block(DenseSizes N, DynOff off, const double* __restrict__ x, const double* __restrict__ b, double lambda, int nd,
                                                   const State& cur, const State& trs, int i, Scalars* sc) {
    double* pq = const_cast<double*>(trs.pq); double* pt = const_cast<double*>(trs.pt); double* D = const_cast<double*>(trs.D);
    double* mq = const_cast<double*>(trs.mq); double* mt = const_cast<double*>(trs.mt); double* J = const_cast<double*>(trs.J);
    if (i < N.n_poses) {
        const int o = off.pose[i];
        if (o >= 0) pose_oplus(cur.pq + 4 * i, cur.pt + 3 * i, x + o, pq + 4 * i, pt + 3 * i);
        else { for (int k = 0; k < 4; ++k) pq[4 * i + k] = cur.pq[4 * i + k]; for (int k = 0; k < 3; ++k) pt[3 * i + k] = cur.pt[3 * i + k]; }
    } else if (i < N.n_poses + N.n_dists) {
        const int j = i - N.n_poses, o = off.dist[j];
        D[j] = cur.D[j] + (o >= 0 ? x[o] : 0.0);
    } else if (i < N.n_poses + N.n_dists + N.n_motions) {
        const int j = i - N.n_poses - N.n_dists, o = off.motion[j];
        if (o >= 0) motion_oplus(cur.mq + 4 * j, cur.mt + 3 * j, x + o, mq + 4 * j, mt + 3 * j);
        else { for (int k = 0; k < 4; ++k) mq[4 * j + k] = cur.mq[4 * j + k]; for (int k = 0; k < 3; ++k) mt[3 * j + k] = cur.mt[3 * j + k]; }
    } else if (i < N.n_poses + N.n_dists + N.n_motions + N.n_joints) {
        const int j = i - N.n_poses - N.n_dists - N.n_motions, o = off.joint[j];
        for (int k = 0; k < 3; ++k) J[3 * j + k] = cur.J[3 * j + k] + (o >= 0 ? x[o + k] : 0.0);
    }
    double s = 0;
    if (i < nd) s = x[i] * (lambda * x[i] + b[i]);
    block_sum_to(s, &sc->scale);
}

// xl = Dinv (bl - W^T xp) per point (edges of a point are contiguous), trial point, landmark part of the scale; kPtLanes lanes per
// point share its edges, the three partial sums meet by shuffles
__global__ void __launch_bounds__(kBaThreads) ba_backsub_kernel(const int* __restrict__ point_ptr, const int* __restrict__ e_pose,
                                                               const uint8_t* __restrict__ e_level, const int* __restrict__ off_pose,
                                                               const uint8_t* __restrict__ act_point, int np, const double* __restrict__ W,
                                                               const double* __restrict__ Dinv, const double* __restrict__ bl,
                                                               const double* __restrict__ x, const Lm* __restrict__ lm, StatePair SP, Scalars* sc,
                                                               DenseSizes N, DynOff off, const double* __restrict__ b, int nd, int dense_blocks) {
    if (lm->done) return;
    const double lambda = lm->lambda;
    if ((int)blockIdx.x < dense_blocks) {   // SparseOptimizer::update for the dense vertices rides in the first blocks of the same launch
        dense_update_block(N, off, x, b, lambda, nd, SP.s[lm->cur], SP.s[lm->cur ^ 1], blockIdx.x * blockDim.x + threadIdx.x, sc);
        return;
    }
    const double* __restrict__ Xcur = SP.s[lm->cur].X;
    double* __restrict__ Xtrial = const_cast<double*>(SP.s[lm->cur ^ 1].X);
    const int l = ((blockIdx.x - dense_blocks) * blockDim.x + threadIdx.x) / kPtLanes, sub = threadIdx.x % kPtLanes;
    double s = 0;
    const bool active = l < np && act_point[l];
    double c[3] = {0, 0, 0};
    if (active) {
        for (int e = point_ptr[l] + sub; e < point_ptr[l + 1]; e += kPtLanes) {
            if (e_level[e]) continue;
            const int o = off_pose[e_pose[e]];
            if (o < 0) continue;
            const double* We = W + 18 * (size_t)e;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double xi = x[o + i];
                c[0] -= We[i * 3] * xi; c[1] -= We[i * 3 + 1] * xi; c[2] -= We[i * 3 + 2] * xi;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 1; o < kPtLanes; o <<= 1) c[k] += __shfl_xor_sync(0xFFFFFFFFu, c[k], o);
    if (l < np && sub == 0) {
        if (!active) {
            for (int k = 0; k < 3; ++k) Xtrial[3 * (size_t)l + k] = Xcur[3 * (size_t)l + k];
        } else {
            for (int k = 0; k < 3; ++k) c[k] += bl[3 * (size_t)l + k];
            const double* Di = Dinv + 9 * (size_t)l;
            for (int i = 0; i < 3; ++i) {
                const double xl = Di[i * 3] * c[0] + Di[i * 3 + 1] * c[1] + Di[i * 3 + 2] * c[2];
                Xtrial[3 * (size_t)l + i] = Xcur[3 * (size_t)l + i] + xl;
                s += xl * (lambda * xl + bl[3 * (size_t)l + i]);
            }
        }
    }
    block_sum_to(s, &sc->scale);
}

// computeActiveErrors + activeRobustChi2 on the trial state (static edges)
__global__ void __launch_bounds__(kBaThreads) ba_eval_kernel(Cam C, Opt O, StaticEdges E, StatePair SP, const Lm* __restrict__ lm, ChiPair chi, Scalars* sc) {
    if (lm->done) return;
    const State S = SP.s[lm->cur ^ 1];
    double* __restrict__ chi_e = chi.e[lm->cur ^ 1];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double rho0 = 0, rho1;
    if (e < E.n && E.level[e]) chi_e[e] = chi.e[lm->cur][e];   // level-1 edges keep their last chi2
    if (e < E.n && !E.level[e]) {
        const int ip = E.pose[e];
        double R[9], er[3], Xc[3];
        quat_to_rot(S.pq + 4 * ip, R);
        const int dim = reproj_error(C, R, S.pt + 3 * ip, S.X + 3 * E.point[e], E.obs + 3 * e, er, Xc);
        const double w0 = E.info[e];
        const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
        chi_e[e] = c;
        huber(dim == 3 ? O.huber_stereo : O.huber_mono, O.robust, c, &rho0, &rho1);
    }
    block_sum_to(rho0, &sc->chi_trial);
}

// chi2 gates + live depth test (src/Optimizer.cc:633-662, 671-699): flag = chi2 > gate || z <= 0
__global__ void ba_gate_kernel(StaticEdges E, StatePair SP, const Lm* __restrict__ lm, ChiPair chi, double gate_mono, double gate_stereo,
                               uint8_t* __restrict__ flag) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E.n) return;
    const State S = SP.s[lm->cur];
    const double* __restrict__ chi_e = chi.e[lm->chi_last];
    const int ip = E.pose[e];
    double R[9];
    quat_to_rot(S.pq + 4 * ip, R);
    const double* X = S.X + 3 * E.point[e];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + S.pt[3 * ip + 2];
    const bool stereo = E.obs[3 * e + 2] >= 0;
    flag[e] = (chi_e[e] > (stereo ? gate_stereo : gate_mono)) || !(z > 0);
}

// ---- Schur pair list on the device (was a host counting sort: 1.9 ms per layout at config 4, twice per solve).
// Pairs (e1, e2) of one point with block(e1) >= block(e2), grouped by destination block key = b1 (b1 + 1) / 2 + b2 and cut into
// chunks of <= kSchurChunk pairs.  Order inside a key is the order of the atomic cursors; the chunk sums meet in L2 atomics anyway.
__device__ __forceinline__ int pose_block_of(int e, const int* __restrict__ e_pose, const uint8_t* __restrict__ e_level, const int* __restrict__ off_pose) {
    if (e_level[e]) return -1;
    const int o = off_pose[e_pose[e]];
    return o < 0 ? -1 : o / 6;
}
// pass 0: count per key; pass 1: place (cursor starts at the key's first slot)
__global__ void __launch_bounds__(128) schur_pairs_kernel(const int* __restrict__ point_ptr, int np, const int* __restrict__ e_pose,
                                                         const uint8_t* __restrict__ e_level, const int* __restrict__ off_pose, int* __restrict__ counter,
                                                         int2* __restrict__ pairs, int pass) {
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) / kPtLanes, sub = threadIdx.x % kPtLanes;
    if (l >= np) return;
    const int lo = point_ptr[l], hi = point_ptr[l + 1];
    for (int a = lo + sub; a < hi; a += kPtLanes) {
        const int b1 = pose_block_of(a, e_pose, e_level, off_pose);
        if (b1 < 0) continue;
        const int base = b1 * (b1 + 1) / 2;
        for (int c = lo; c < hi; ++c) {
            const int b2 = pose_block_of(c, e_pose, e_level, off_pose);
            if (b2 < 0 || b2 > b1) continue;
            const int pos = atomicAdd(counter + base + b2, 1);
            if (pass == 1) pairs[pos] = make_int2(a, c);
        }
    }
}
// one block: exclusive scans of the counts (-> cursor = first slot of every key) and of the chunk counts; writes the chunk list
__global__ void __launch_bounds__(1024) schur_scan_kernel(const int* __restrict__ count, int nkeys, int* __restrict__ cursor, int2* __restrict__ chunks,
                                                         int* __restrict__ totals /* [0] pairs, [1] chunks */) {
    __shared__ int warp_p[32], warp_c[32];
    __shared__ int carry_p, carry_c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_p = 0; carry_c = 0; }
    __syncthreads();
    for (int base = 0; base < nkeys; base += 1024) {
        const int k = base + tid;
        const int cnt = k < nkeys ? count[k] : 0, nch = (cnt + kSchurChunk - 1) / kSchurChunk;
        int ip = cnt, ic = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tp = __shfl_up_sync(0xFFFFFFFFu, ip, o), tc = __shfl_up_sync(0xFFFFFFFFu, ic, o);
            if (lane >= o) { ip += tp; ic += tc; }
        }
        if (lane == 31) { warp_p[warp] = ip; warp_c[warp] = ic; }
        __syncthreads();
        if (warp == 0) {
            int wp = warp_p[lane], wc = warp_c[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tp = __shfl_up_sync(0xFFFFFFFFu, wp, o), tc = __shfl_up_sync(0xFFFFFFFFu, wc, o);
                if (lane >= o) { wp += tp; wc += tc; }
            }
            warp_p[lane] = wp; warp_c[lane] = wc;
        }
        __syncthreads();
        const int start_p = carry_p + (warp ? warp_p[warp - 1] : 0) + ip - cnt;
        const int start_c = carry_c + (warp ? warp_c[warp - 1] : 0) + ic - nch;
        if (k < nkeys) {
            cursor[k] = start_p;
            for (int i = 0; i < nch; ++i) chunks[start_c + i] = make_int2(start_p + i * kSchurChunk, min(kSchurChunk, cnt - i * kSchurChunk));
        }
        __syncthreads();
        if (tid == 1023) { carry_p += warp_p[31]; carry_c += warp_c[31]; }
        __syncthreads();
    }
    if (tid == 0) { totals[0] = carry_p; totals[1] = carry_c; }
}

// H = 0, b = 0 and the buildSystem sums, only when the step opens an outer iteration
__global__ void ba_clear_kernel(double* __restrict__ H, size_t n2, double* __restrict__ b, int nd, const Lm* __restrict__ lm, Scalars* sc) {
    if (lm->done || !lm->need_lin) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) H[i] = 0.0;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (size_t)nd) b[t] = 0.0;
    if (t == 0) { sc->chi_cur = 0.0; sc->maxdiag = 0.0; }
}

// between buildSystem and the first trial of an outer iteration: computeLambdaInit on the first one (lambda = tau max diag),
// chi2 bookkeeping (levenberg.cpp:77-92); always: reset the sums of the coming trial
__global__ void lm_iter_kernel(Lm* lm, Scalars* sc) {
    if (lm->done) return;
    if (lm->need_lin) {
        lm->current = sc->chi_cur;
        if (lm->first) {
            lm->lambda = lm->tau * sc->maxdiag; lm->ni = 2; lm->nbad = 0; lm->first = 0;
            if (lm->round == 0) lm->chi2_initial = lm->current;
        }
        lm->ini = lm->current;
        lm->q = 0;
        lm->need_lin = 0;
        lm->chi_last = lm->cur;
    }
    sc->chi_trial = 0.0; sc->scale = 0.0; sc->info = 0;
}

// after the trial state has been evaluated: the accept / reject decision and the loop control of
// OptimizationAlgorithmLevenberg::solve (levenberg.cpp:95-147) and SparseOptimizer::optimize (the nbad early exit of
// src/Optimizer.cc's g2o copy); appends one row to the trace
__global__ void lm_decide_kernel(Lm* lm, const Scalars* sc, double* __restrict__ trace) {
    if (lm->done) return;
    const bool ok = sc->info == 0;
    const double temp = ok ? sc->chi_trial : DBL_MAX;
    double rho = lm->current - temp;
    const double scale = (ok ? sc->scale : 0.0) + 1e-3;
    rho /= scale;
    const bool good = rho > 0 && isfinite(temp);
    if (trace && lm->trace_len < lm->trace_cap) {
        double* t = trace + (size_t)ADB_BA_TRACE_COLS * lm->trace_len++;
        t[0] = lm->lambda; t[1] = lm->current; t[2] = temp; t[3] = rho; t[4] = good ? 1 : 0;
    }
    lm->trials_run++;
    lm->chi_last = lm->cur ^ 1;            // the trial buffers hold the errors of the last evaluation, accepted or not
    if (good) {
        double alpha = 1. - pow(2 * rho - 1, 3.0);
        alpha = fmin(alpha, 2. / 3.);
        lm->lambda *= fmax(1. / 3., alpha);
        lm->ni = 2;
        lm->current = temp;
        lm->cur ^= 1;                      // accept: the trial buffers become the state (g2o: discardTop)
    } else {
        lm->lambda *= lm->ni;
        lm->ni *= 2;                       // reject: nothing to restore (g2o: pop)
    }
    ++lm->q;
    if (rho < 0 && lm->q < lm->max_trials) return;          // another trial on the same linearisation
    ++lm->iterations_run;
    if (lm->q == lm->max_trials || rho == 0) { lm->done = 1; return; }
    if ((lm->ini - lm->current) * 1e3 < lm->ini) ++lm->nbad; else lm->nbad = 0;
    if (lm->nbad >= 3) { lm->done = 1; return; }
    if (++lm->it >= lm->it_max) { lm->done = 1; return; }
    lm->need_lin = 1;
}

// ----------------------------------------------------------------------------------------
// Optimizer::PoseOptimization (src/Optimizer.cc:232-429), one CTA per frame: the whole 4 x 10 LM
// schedule runs inside the kernel (6x6 system, deterministic block reductions, no host round trip).
struct PoseArgs {
    Cam C;
    const int* frame_ptr; double* pose_q; double* pose_t;
    const float* xw; const float* obs; const float* inv_sigma2;
    uint8_t* outlier; int* n_inliers;
};
constexpr int kPoseThreads = 256;

__device__ inline void pose_edge_error(const Cam& C, const double* R, const double* t, const float* xw, const float* ob, bool stereo, double* er,
                                       double* Xc) {
    const double X0 = (double)xw[0], X1 = (double)xw[1], X2 = (double)xw[2];
    for (int i = 0; i < 3; ++i) Xc[i] = R[i * 3] * X0 + R[i * 3 + 1] * X1 + R[i * 3 + 2] * X2 + t[i];
    if (stereo) {
        const float invz = (float)(1.0 / Xc[2]);
        const double u = Xc[0] * invz * C.fx + C.cx, v = Xc[1] * invz * C.fy + C.cy;
        er[0] = (double)ob[0] - u; er[1] = (double)ob[1] - v; er[2] = (double)ob[2] - (u - C.bf * invz);   // double bf in the OnlyPose edge
    } else {
        er[0] = (double)ob[0] - (Xc[0] / Xc[2] * C.fx + C.cx); er[1] = (double)ob[1] - (Xc[1] / Xc[2] * C.fy + C.cy); er[2] = 0;
    }
}

// linearizeOplus of Edge(Stereo)SE3ProjectXYZOnlyPose (types_six_dof_expmap.cpp:300-364): the reference's own reciprocal form
__device__ inline void pose_edge_jacobian(const Cam& C, const double* Xc, bool stereo, double* J) {
    const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz, fx = C.fx, fy = C.fy, bf = C.bf;
    J[0] = x * y * invz_2 * fx; J[1] = -(1 + (x * x * invz_2)) * fx; J[2] = y * invz * fx; J[3] = -invz * fx; J[4] = 0; J[5] = x * invz_2 * fx;
    J[6] = (1 + y * y * invz_2) * fy; J[7] = -x * y * invz_2 * fy; J[8] = -x * invz * fy; J[9] = 0; J[10] = -invz * fy; J[11] = y * invz_2 * fy;
    if (stereo) { J[12] = J[0] - bf * y * invz_2; J[13] = J[1] + bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf * invz_2; }
    else { for (int k = 12; k < 18; ++k) J[k] = 0; }
}

__device__ inline bool chol6_solve(const double* H, double lambda, const double* b, double* x) {
    double L[36];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) L[i * 6 + j] = H[i * 6 + j] + (i == j ? lambda : 0.0);
    for (int j = 0; j < 6; ++j) {
        double d = L[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0) || !isfinite(d)) return false;
        d = sqrt(d);
        L[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = L[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / d;
        }
    }
    for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * x[k]; x[i] = s / L[i * 6 + i]; }
    for (int i = 5; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k]; x[i] = s / L[i * 6 + i]; }
    return true;
}

__global__ void __launch_bounds__(kPoseThreads) pose_optimize_kernel(PoseArgs A, double* __restrict__ chi_scratch, uint8_t* __restrict__ lvl_scratch) {
    __shared__ double scratch[(kPoseThreads / 32) * 28], red[28];
    __shared__ double sq[4], st[3], tq[4], tt[3], sH[36], sb[6];
    __shared__ double s_lambda, s_ni, s_current, s_rho, s_den;
    __shared__ int s_flag, s_nbad_it, s_q;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int a = A.frame_ptr[f], n = A.frame_ptr[f + 1] - a;
    double* chi = chi_scratch + a;
    uint8_t* level = lvl_scratch + a;
    for (int i = tid; i < n; i += kPoseThreads) { A.outlier[a + i] = 0; level[i] = 0; chi[i] = 0; }
    if (n < 3) { if (tid == 0) A.n_inliers[f] = 0; return; }
    const double dm = (double)(float)sqrt(5.991), ds = (double)(float)sqrt(7.815);
    double q0[4], t0[3];
    for (int k = 0; k < 4; ++k) q0[k] = A.pose_q[4 * f + k];
    for (int k = 0; k < 3; ++k) t0[k] = A.pose_t[3 * f + k];
    bool robust = true;
    int nBad = 0;
    __syncthreads();

    auto evaluate = [&](const double* qq, const double* tt_) -> double {   // robust chi2 of the active edges, result broadcast
        double R[9], v[1] = {0.0}, r0, r1;
        quat_to_rot(qq, R);
        for (int i = tid; i < n; i += kPoseThreads) {
            if (level[i]) continue;
            const bool stereo = !(A.obs[3 * (size_t)(a + i) + 2] < 0);
            double er[3], Xc[3];
            pose_edge_error(A.C, R, tt_, A.xw + 3 * (size_t)(a + i), A.obs + 3 * (size_t)(a + i), stereo, er, Xc);
            const double w = (double)A.inv_sigma2[a + i];
            const double c = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
            chi[i] = c;
            huber(stereo ? ds : dm, robust, c, &r0, &r1);
            v[0] += r0;
        }
        block_reduce_fixed<1>(v, scratch, red);
        return red[0];
    };

    for (int round = 0; round < 4; ++round) {
        if (tid == 0) { for (int k = 0; k < 4; ++k) sq[k] = q0[k]; for (int k = 0; k < 3; ++k) st[k] = t0[k]; }
        int mine = 0;
        for (int i = tid; i < n; i += kPoseThreads) mine += !level[i];
        const int nact = __syncthreads_count(mine > 0) ? 1 : 0;   // any active edge?
        if (nact) {
            if (tid == 0) { s_lambda = 0; s_ni = 2; s_nbad_it = 0; }
            __syncthreads();
            for (int it = 0; it < 10; ++it) {
                double q[4] = {sq[0], sq[1], sq[2], sq[3]}, t[3] = {st[0], st[1], st[2]};
                const double current0 = evaluate(q, t);
                // buildSystem
                double v[27];
#pragma unroll
                for (int k = 0; k < 27; ++k) v[k] = 0;
                double R[9], r0, r1;
                quat_to_rot(q, R);
                for (int i = tid; i < n; i += kPoseThreads) {
                    if (level[i]) continue;
                    const bool stereo = !(A.obs[3 * (size_t)(a + i) + 2] < 0);
                    double er[3], Xc[3], J[18];
                    pose_edge_error(A.C, R, t, A.xw + 3 * (size_t)(a + i), A.obs + 3 * (size_t)(a + i), stereo, er, Xc);
                    pose_edge_jacobian(A.C, Xc, stereo, J);
                    const int dim = stereo ? 3 : 2;
                    const double w0 = (double)A.inv_sigma2[a + i];
                    huber(stereo ? ds : dm, robust, chi[i], &r0, &r1);
                    const double w = r1 * w0;
                    int u = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
#pragma unroll
                        for (int c = 0; c <= r; ++c, ++u) { double s = 0; for (int k = 0; k < dim; ++k) s += J[k * 6 + r] * w * J[k * 6 + c]; v[u] += s; }
                        double s = 0;
                        for (int k = 0; k < dim; ++k) s += J[k * 6 + r] * (-w0 * er[k] * r1);
                        v[21 + r] += s;
                    }
                }
                block_reduce_fixed<27>(v, scratch, red);
                if (tid == 0) {
                    int u = 0;
                    for (int r = 0; r < 6; ++r) for (int c = 0; c <= r; ++c, ++u) { sH[r * 6 + c] = red[u]; sH[c * 6 + r] = red[u]; }
                    for (int r = 0; r < 6; ++r) sb[r] = red[21 + r];
                    if (it == 0) { double m = 0; for (int r = 0; r < 6; ++r) m = fmax(m, fabs(sH[r * 7])); s_lambda = 1e-5 * m; s_ni = 2; s_nbad_it = 0; }
                    s_current = current0; s_q = 0;
                }
                __syncthreads();
                const double ini = current0;
                while (true) {   // LM trials
                    if (tid == 0) {
                        double x[6];
                        const bool ok = chol6_solve(sH, s_lambda, sb, x);
                        if (ok) pose_oplus(sq, st, x, tq, tt);
                        else { for (int k = 0; k < 4; ++k) tq[k] = sq[k]; for (int k = 0; k < 3; ++k) tt[k] = st[k]; }
                        double scale = 0;
                        if (ok) for (int k = 0; k < 6; ++k) scale += x[k] * (s_lambda * x[k] + sb[k]);
                        s_den = scale + 1e-3;      // denominator of rho (its own word: the other threads may still be reading s_rho of the previous trial)
                        s_flag = ok ? 1 : 0;
                    }
                    __syncthreads();
                    const double qq[4] = {tq[0], tq[1], tq[2], tq[3]}, tt2[3] = {tt[0], tt[1], tt[2]};
                    double temp = evaluate(qq, tt2);
                    if (tid == 0) {
                        if (!s_flag) temp = DBL_MAX;
                        const double rho = (s_current - temp) / s_den;
                        if (rho > 0 && isfinite(temp)) {
                            double alpha = 1. - pow((2 * rho - 1), 3.0);
                            alpha = fmin(alpha, 2. / 3.);
                            s_lambda *= fmax(1. / 3., alpha); s_ni = 2; s_current = temp;
                            for (int k = 0; k < 4; ++k) sq[k] = tq[k];
                            for (int k = 0; k < 3; ++k) st[k] = tt[k];
                        } else { s_lambda *= s_ni; s_ni *= 2; }
                        s_rho = rho;
                        s_q += 1;
                    }
                    __syncthreads();
                    if (!(s_rho < 0 && s_q < 10)) break;
                }
                bool stop_it = (s_q == 10 || s_rho == 0);
                if (!stop_it) {
                    if (tid == 0) { if ((ini - s_current) * 1e3 < ini) s_nbad_it++; else s_nbad_it = 0; }
                    __syncthreads();
                    stop_it = s_nbad_it >= 3;
                }
                __syncthreads();
                if (stop_it) break;
            }
        }
        __syncthreads();
        // classification at the round's final pose
        {
            const double q[4] = {sq[0], sq[1], sq[2], sq[3]}, t[3] = {st[0], st[1], st[2]};
            double R[9];
            quat_to_rot(q, R);
            int bad = 0;
            for (int i = tid; i < n; i += kPoseThreads) {
                const bool stereo = !(A.obs[3 * (size_t)(a + i) + 2] < 0);
                if (A.outlier[a + i]) {
                    double er[3], Xc[3];
                    pose_edge_error(A.C, R, t, A.xw + 3 * (size_t)(a + i), A.obs + 3 * (size_t)(a + i), stereo, er, Xc);
                    const double w = (double)A.inv_sigma2[a + i];
                    chi[i] = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
                }
                const float c = (float)chi[i];
                const bool out = c > (stereo ? 7.815f : 5.991f);
                A.outlier[a + i] = out; level[i] = out; bad += out;
            }
            nBad = __syncthreads_count(0);   // barrier
            double v[1] = {(double)bad};
            block_reduce_fixed<1>(v, scratch, red);
            nBad = (int)red[0];
        }
        if (round == 2) robust = false;
        if (n < 10) break;
        __syncthreads();
    }
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) A.pose_q[4 * f + k] = sq[k];
        for (int k = 0; k < 3; ++k) A.pose_t[3 * f + k] = st[k];
        A.n_inliers[f] = n - nBad;
    }
}


// ----------------------------------------------------------------------------------------
// Pinning hook (adb_ba_leaf_eval): the device functions above on arrays of inputs, one thread per item, so that the tests can hold
// them against values computed by the reference's own sources (tests/golden/ba_leaf_ref.npz).  Output record per item, doubles:
//   [0,3) stereo error  [3,12) d e / d point  [12,30) d e / d pose      (Edge(Stereo)SE3ProjectXYZ; mono: rows 0-1 of the same at +30)
//   [30,33) [33,42) [42,60) the same for the monocular edge            [60,63) [63,81) stereo OnlyPose error / Jacobian
//   [81,84) [84,102) mono OnlyPose   [102,106) [106,109) pose oplus q, t   [109] rigidity error   [110,113) motion error
//   [113,117) [117,120) motion oplus q, t   [120] [121] Huber rho(e2), rho'(e2) for (huber_delta, huber_e2) -- zeros when not given
constexpr int kLeafRecord = 122;
static_assert(kLeafRecord == ADB_BA_LEAF_RECORD, "record width");
__global__ void ba_leaf_kernel(int n, Cam C, const double* __restrict__ pose_q, const double* __restrict__ pose_t, const double* __restrict__ X,
                               const double* __restrict__ obs, const double* __restrict__ pose_update, const double* __restrict__ joint_a,
                               const double* __restrict__ joint_b, const double* __restrict__ bone, const double* __restrict__ motion_q,
                               const double* __restrict__ motion_t, const double* __restrict__ motion_dt, const double* __restrict__ motion_update,
                               const double* __restrict__ huber_delta, const double* __restrict__ huber_e2, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* o = out + (size_t)kLeafRecord * i;
    double R[9], Xc[3], er[3], Ji[9], Jj[18];
    quat_to_rot(pose_q + 4 * i, R);
    const double ob[3] = {obs[3 * i], obs[3 * i + 1], obs[3 * i + 2]};
    int dim = reproj_error(C, R, pose_t + 3 * i, X + 3 * i, ob, er, Xc);
    reproj_jacobians(C, R, Xc, dim, Ji, Jj);
    for (int k = 0; k < 3; ++k) o[k] = er[k];
    for (int k = 0; k < 9; ++k) o[3 + k] = Ji[k];
    for (int k = 0; k < 18; ++k) o[12 + k] = Jj[k];
    const double obm[3] = {ob[0], ob[1], -1.0};
    dim = reproj_error(C, R, pose_t + 3 * i, X + 3 * i, obm, er, Xc);
    reproj_jacobians(C, R, Xc, dim, Ji, Jj);
    for (int k = 0; k < 3; ++k) o[30 + k] = er[k];
    for (int k = 0; k < 9; ++k) o[33 + k] = Ji[k];
    for (int k = 0; k < 18; ++k) o[42 + k] = Jj[k];
    const float xw[3] = {(float)X[3 * i], (float)X[3 * i + 1], (float)X[3 * i + 2]}, of[3] = {(float)ob[0], (float)ob[1], (float)ob[2]};
    for (int stereo = 1; stereo >= 0; --stereo) {
        pose_edge_error(C, R, pose_t + 3 * i, xw, of, stereo != 0, er, Xc);
        pose_edge_jacobian(C, Xc, stereo != 0, Jj);
        double* p = o + (stereo ? 60 : 81);
        for (int k = 0; k < 3; ++k) p[k] = er[k];
        for (int k = 0; k < 18; ++k) p[3 + k] = Jj[k];
    }
    pose_oplus(pose_q + 4 * i, pose_t + 3 * i, pose_update + 6 * i, o + 102, o + 106);
    double d3[3], nrm;
    o[109] = rigid_error(joint_a + 3 * i, joint_b + 3 * i, bone[i], d3, &nrm);
    double Rm[9];
    motion_error_qt(motion_q + 4 * i, motion_t + 3 * i, motion_dt[i], joint_a + 3 * i, joint_b + 3 * i, o + 110, Rm);
    motion_oplus(motion_q + 4 * i, motion_t + 3 * i, motion_update + 6 * i, o + 113, o + 117);
    o[120] = o[121] = 0.0;
    if (huber_delta) huber(huber_delta[i], true, huber_e2[i], o + 120, o + 121);
}

// ----------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    adb_status ensure(size_t bytes) {
        if (bytes <= cap) return ADB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        ADB_CUDA(cudaMalloc(&p, want));
        cap = want;
        return ADB_OK;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace adb

using namespace adb;

struct adb_ba {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: the pose side of buildSystem, forked / joined by events
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // device buffers
    DevBuf pq[2], pt[2], X[2], Jt[2], Dd[2], mq[2], mt[2];                         // double-buffered state
    DevBuf e_pose, e_point, e_obs, e_info, e_level, point_ptr, pairs, chunks, pair_count, pair_cursor, pair_totals, off_pose, act_point, pose_ptr, pose_edges, free_pose;
    DevBuf j_pose, j_joint, j_obs, j_info, j_level, r_i, r_j, r_d, r_info, r_level, m_p1, m_p2, m_m, m_dt, m_info, m_level;
    DevBuf off_joint, off_dist, off_motion;
    DevBuf H, b, Sm, bs, Hll, bl, W, Dinv, db, chi_e[2], chi_j[2], chi_r[2], chi_m[2], flag, scal, work, lm, trace, pose_partial;
    Lm* h_lm = nullptr;          // pinned mirror of the device LM controller
    char* h_stage = nullptr;     // pinned staging arena for the sorted edge arrays (true asynchronous H2D at PCIe rate)
    size_t h_stage_cap = 0;
    float stage_ms[6] = {0, 0, 0, 0, 0, 0};
    long long launches = 0;
    std::vector<cudaEvent_t> tev;   // per-stage timing events
};

namespace {

struct Timer {   // accumulates device time per stage with event pairs (only when profiling is on)
    adb_ba* s;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
    size_t next = 0;
    explicit Timer(adb_ba* s_) : s(s_) {}
    cudaEvent_t get() {
        if (next == s->tev.size()) { cudaEvent_t e; cudaEventCreate(&e); s->tev.push_back(e); }
        return s->tev[next++];
    }
    void begin(int stage) { cudaEvent_t a = get(); cudaEventRecord(a, s->stream); spans.push_back({stage, {a, nullptr}}); }
    void end() { cudaEvent_t b = get(); cudaEventRecord(b, s->stream); spans.back().second.second = b; }
    void collect() {
        for (int i = 0; i < 6; ++i) s->stage_ms[i] = 0;
        for (auto& sp : spans) { float ms = 0; cudaEventElapsedTime(&ms, sp.second.first, sp.second.second); s->stage_ms[sp.first] += ms; }
        if (!spans.empty()) cudaEventElapsedTime(&s->stage_ms[5], spans.front().second.first, spans.back().second.second);
    }
};

template <typename T>
adb_status upload(DevBuf& d, const T* src, size_t n, cudaStream_t st) {
    adb_status s = d.ensure(std::max<size_t>(n, 1) * sizeof(T));
    if (s != ADB_OK) return s;
    if (n) ADB_CUDA(cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return ADB_OK;
}

inline int grid_for(size_t n, int threads) { return (int)std::max<size_t>(1, (n + threads - 1) / threads); }

struct Ctx {
    adb_ba* s;
    adb_ba_problem* P;
    const adb_ba_options* O;
    adb_ba_result* R;
    volatile const uint8_t* stop;
    Timer tm;
    // host-side derived structure
    std::vector<int> perm;            // sorted position -> original edge index
    int *se_pose = nullptr, *se_point = nullptr, *ptr = nullptr;   // in the pinned arena of the handle
    double *se_obs = nullptr, *se_info = nullptr;
    std::vector<uint8_t> lvl_e, lvl_j, lvl_r, lvl_m, act_point;
    std::vector<int> off_pose, off_dist, off_motion, off_joint;
    int chunks_ub = 0;     // upper bound of the Schur chunk count (the exact one lives on the device)
    int *pose_ptr = nullptr, *pose_edges = nullptr;
    std::vector<int> free_pose;
    int nd = 0, ld = 32;
    Lm hl = {};            // host copy of the controller as of the last read-back
    Ctx(adb_ba* s_, adb_ba_problem* p, const adb_ba_options* o, adb_ba_result* r, volatile const uint8_t* st) : s(s_), P(p), O(o), R(r), stop(st), tm(s_) {}

    bool stopped() const { return stop && *stop; }

    State state(int i) const {
        return State{s->pq[i].as<double>(), s->pt[i].as<double>(), s->X[i].as<double>(), s->Jt[i].as<double>(), s->Dd[i].as<double>(),
                     s->mq[i].as<double>(), s->mt[i].as<double>()};
    }
    StatePair states() const { return StatePair{{state(0), state(1)}}; }
    ChiPair chis() const {
        return ChiPair{{s->chi_e[0].as<double>(), s->chi_e[1].as<double>()}, {s->chi_j[0].as<double>(), s->chi_j[1].as<double>()},
                       {s->chi_r[0].as<double>(), s->chi_r[1].as<double>()}, {s->chi_m[0].as<double>(), s->chi_m[1].as<double>()}};
    }
    Lm* dlm() const { return s->lm.as<Lm>(); }
    StaticEdges sedges() const {
        return StaticEdges{P->n_edges, s->e_pose.as<int>(), s->e_point.as<int>(), s->e_obs.as<double>(), s->e_info.as<double>(), s->e_level.as<uint8_t>()};
    }
    DynEdges dedges() const {
        return DynEdges{P->n_joint_edges, P->n_rigid_edges, P->n_motion_edges,
                        s->j_pose.as<int>(), s->j_joint.as<int>(), s->j_obs.as<double>(), s->j_info.as<double>(), s->j_level.as<uint8_t>(),
                        s->r_i.as<int>(), s->r_j.as<int>(), s->r_d.as<int>(), s->r_info.as<double>(), s->r_level.as<uint8_t>(),
                        s->m_p1.as<int>(), s->m_p2.as<int>(), s->m_m.as<int>(), s->m_dt.as<double>(), s->m_info.as<double>(), s->m_level.as<uint8_t>()};
    }
    DynOff doff() const { return DynOff{s->off_pose.as<int>(), s->off_joint.as<int>(), s->off_dist.as<int>(), s->off_motion.as<int>()}; }
    Cam cam() const { return Cam{P->fx, P->fy, P->cx, P->cy, P->bf}; }
    Opt opt(bool robust) const { return Opt{O->huber_mono, O->huber_stereo, O->huber_rigid, O->huber_motion, robust ? 1 : 0}; }
    int n_dyn() const { return P->n_joint_edges + P->n_rigid_edges + P->n_motion_edges; }

    // every index array of the articulated-human part against its vertex count, NULL arrays with a non-zero count
    adb_status validate_dynamic() const {
        ADB_CHECK(P->n_joints >= 0 && P->n_dists >= 0 && P->n_motions >= 0 && P->n_joint_edges >= 0 && P->n_rigid_edges >= 0 && P->n_motion_edges >= 0,
                  ADB_ERR_INVALID, "negative count in the articulated part");
        ADB_CHECK(P->pose_q && P->pose_t && P->pose_fixed, ADB_ERR_INVALID, "null pose arrays");
        ADB_CHECK(P->n_points == 0 || P->points, ADB_ERR_INVALID, "null point array");
        ADB_CHECK(P->n_edges == 0 || (P->edge_pose && P->edge_point && P->edge_obs && P->edge_info), ADB_ERR_INVALID, "null edge arrays");
        ADB_CHECK(P->n_joints == 0 || P->joints, ADB_ERR_INVALID, "null joint array");
        ADB_CHECK(P->n_dists == 0 || P->dists, ADB_ERR_INVALID, "null bone-length array");
        ADB_CHECK(P->n_motions == 0 || (P->motion_q && P->motion_t), ADB_ERR_INVALID, "null motion arrays");
        ADB_CHECK(P->n_joint_edges == 0 || (P->jedge_pose && P->jedge_joint && P->jedge_obs && P->jedge_info), ADB_ERR_INVALID, "null joint-edge arrays");
        ADB_CHECK(P->n_rigid_edges == 0 || (P->redge_i && P->redge_j && P->redge_dist && P->redge_info), ADB_ERR_INVALID, "null rigidity-edge arrays");
        ADB_CHECK(P->n_motion_edges == 0 || (P->medge_p1 && P->medge_p2 && P->medge_motion && P->medge_dt && P->medge_info), ADB_ERR_INVALID,
                  "null motion-edge arrays");
        auto in = [](int v, int n) { return v >= 0 && v < n; };
        for (int e = 0; e < P->n_joint_edges; ++e)
            ADB_CHECK(in(P->jedge_pose[e], P->n_poses) && in(P->jedge_joint[e], P->n_joints), ADB_ERR_INVALID,
                      "joint edge %d references pose %d / joint %d out of range", e, P->jedge_pose[e], P->jedge_joint[e]);
        for (int e = 0; e < P->n_rigid_edges; ++e)
            ADB_CHECK(in(P->redge_i[e], P->n_joints) && in(P->redge_j[e], P->n_joints) && in(P->redge_dist[e], P->n_dists), ADB_ERR_INVALID,
                      "rigidity edge %d references joints %d, %d / bone length %d out of range", e, P->redge_i[e], P->redge_j[e], P->redge_dist[e]);
        for (int e = 0; e < P->n_motion_edges; ++e)
            ADB_CHECK(in(P->medge_p1[e], P->n_joints) && in(P->medge_p2[e], P->n_joints) && in(P->medge_motion[e], P->n_motions), ADB_ERR_INVALID,
                      "motion edge %d references joints %d, %d / motion %d out of range", e, P->medge_p1[e], P->medge_p2[e], P->medge_motion[e]);
        return ADB_OK;
    }

    // ---- one-time upload: edges sorted by point (stable), state
    adb_status upload_problem() {
        const int E = P->n_edges, NP = P->n_points;
        {
            const adb_status v = validate_dynamic();
            if (v != ADB_OK) return v;
        }
        {   // carve the sorted arrays out of the pinned arena (grown on demand, kept across solves)
            const size_t need = ((size_t)E * (4 + 4 + 24 + 8 + 4) + (size_t)(NP + 1 + P->n_poses + 1) * 4 + 256 + 7 * 16);
            if (need > s->h_stage_cap) {
                ADB_CUDA(cudaStreamSynchronize(s->stream));
                if (s->h_stage) cudaFreeHost(s->h_stage);
                s->h_stage = nullptr; s->h_stage_cap = 0;
                ADB_CUDA(cudaMallocHost(&s->h_stage, need + need / 4));
                s->h_stage_cap = need + need / 4;
            }
            char* cursor = s->h_stage;
            auto take = [&](size_t bytes) { char* r0 = cursor; cursor += (bytes + 15) & ~(size_t)15; return r0; };
            se_obs = reinterpret_cast<double*>(take((size_t)E * 24)); se_info = reinterpret_cast<double*>(take((size_t)E * 8));
            se_pose = reinterpret_cast<int*>(take((size_t)E * 4)); se_point = reinterpret_cast<int*>(take((size_t)E * 4));
            pose_edges = reinterpret_cast<int*>(take((size_t)E * 4));
            ptr = reinterpret_cast<int*>(take((size_t)(NP + 1) * 4)); pose_ptr = reinterpret_cast<int*>(take((size_t)(P->n_poses + 1) * 4));
        }
        std::fill(ptr, ptr + NP + 1, 0);
        for (int e = 0; e < E; ++e) {
            ADB_CHECK(P->edge_point[e] >= 0 && P->edge_point[e] < NP && P->edge_pose[e] >= 0 && P->edge_pose[e] < P->n_poses, ADB_ERR_INVALID,
                      "edge %d references pose %d / point %d out of range", e, P->edge_pose[e], P->edge_point[e]);
            ptr[P->edge_point[e] + 1]++;
        }
        for (int l = 0; l < NP; ++l) ptr[l + 1] += ptr[l];
        perm.assign(E, 0);
        {
            std::vector<int> fill(ptr, ptr + NP);
            for (int e = 0; e < E; ++e) perm[fill[P->edge_point[e]]++] = e;
        }
        for (int k = 0; k < E; ++k) {
            const int e = perm[k];
            se_pose[k] = P->edge_pose[e]; se_point[k] = P->edge_point[e]; se_info[k] = P->edge_info[e];
            for (int c = 0; c < 3; ++c) se_obs[(size_t)3 * k + c] = P->edge_obs[(size_t)3 * e + c];
        }
        // CSR of the (sorted) edges by pose, for the atomic-free pose-block accumulation
        std::fill(pose_ptr, pose_ptr + P->n_poses + 1, 0);
        for (int k = 0; k < E; ++k) pose_ptr[se_pose[k] + 1]++;
        for (int i = 0; i < P->n_poses; ++i) pose_ptr[i + 1] += pose_ptr[i];
        {
            std::vector<int> fill(pose_ptr, pose_ptr + P->n_poses);
            for (int k = 0; k < E; ++k) pose_edges[fill[se_pose[k]]++] = k;
        }
        cudaStream_t st = s->stream;
        adb_status r;
#define UP(buf, ptr_, n) if ((r = upload(buf, ptr_, (size_t)(n), st)) != ADB_OK) return r
        UP(s->e_pose, se_pose, E); UP(s->e_point, se_point, E); UP(s->e_obs, se_obs, 3 * (size_t)E);
        UP(s->e_info, se_info, E); UP(s->point_ptr, ptr, NP + 1);
        UP(s->pose_ptr, pose_ptr, P->n_poses + 1); UP(s->pose_edges, pose_edges, E);
        UP(s->pq[0], P->pose_q, 4 * (size_t)P->n_poses); UP(s->pt[0], P->pose_t, 3 * (size_t)P->n_poses); UP(s->X[0], P->points, 3 * (size_t)NP);
        UP(s->Jt[0], P->joints, 3 * (size_t)P->n_joints); UP(s->Dd[0], P->dists, P->n_dists);
        UP(s->mq[0], P->motion_q, 4 * (size_t)P->n_motions); UP(s->mt[0], P->motion_t, 3 * (size_t)P->n_motions);
        UP(s->j_pose, P->jedge_pose, P->n_joint_edges); UP(s->j_joint, P->jedge_joint, P->n_joint_edges);
        UP(s->j_obs, P->jedge_obs, 3 * (size_t)P->n_joint_edges); UP(s->j_info, P->jedge_info, P->n_joint_edges);
        UP(s->r_i, P->redge_i, P->n_rigid_edges); UP(s->r_j, P->redge_j, P->n_rigid_edges); UP(s->r_d, P->redge_dist, P->n_rigid_edges);
        UP(s->r_info, P->redge_info, P->n_rigid_edges);
        UP(s->m_p1, P->medge_p1, P->n_motion_edges); UP(s->m_p2, P->medge_p2, P->n_motion_edges); UP(s->m_m, P->medge_motion, P->n_motion_edges);
        UP(s->m_dt, P->medge_dt, P->n_motion_edges); UP(s->m_info, P->medge_info, P->n_motion_edges);
#undef UP
        // trial buffers + work arrays
        DevBuf* pairs2[] = {&s->pq[1], &s->pt[1], &s->X[1], &s->Jt[1], &s->Dd[1], &s->mq[1], &s->mt[1]};
        const size_t sz2[] = {4 * (size_t)P->n_poses, 3 * (size_t)P->n_poses, 3 * (size_t)NP, 3 * (size_t)P->n_joints, (size_t)P->n_dists,
                              4 * (size_t)P->n_motions, 3 * (size_t)P->n_motions};
        for (int i = 0; i < 7; ++i) if ((r = pairs2[i]->ensure(std::max<size_t>(sz2[i], 1) * 8)) != ADB_OK) return r;
        if ((r = s->Hll.ensure(std::max<size_t>(NP, 1) * 48)) != ADB_OK) return r;
        if ((r = s->bl.ensure(std::max<size_t>(NP, 1) * 24)) != ADB_OK) return r;
        if ((r = s->W.ensure(std::max<size_t>(E, 1) * 144)) != ADB_OK) return r;
        if ((r = s->Dinv.ensure(std::max<size_t>(NP, 1) * 72)) != ADB_OK) return r;
        if ((r = s->db.ensure(std::max<size_t>(NP, 1) * 24)) != ADB_OK) return r;
        for (int i = 0; i < 2; ++i) {
            if ((r = s->chi_e[i].ensure(std::max<size_t>(E, 1) * 8)) != ADB_OK) return r;
            if ((r = s->chi_j[i].ensure(std::max<size_t>(P->n_joint_edges, 1) * 8)) != ADB_OK) return r;
            if ((r = s->chi_r[i].ensure(std::max<size_t>(P->n_rigid_edges, 1) * 8)) != ADB_OK) return r;
            if ((r = s->chi_m[i].ensure(std::max<size_t>(P->n_motion_edges, 1) * 8)) != ADB_OK) return r;
            ADB_CUDA(cudaMemsetAsync(s->chi_e[i].p, 0, std::max<size_t>(E, 1) * 8, st));
            ADB_CUDA(cudaMemsetAsync(s->chi_j[i].p, 0, std::max<size_t>(P->n_joint_edges, 1) * 8, st));
            ADB_CUDA(cudaMemsetAsync(s->chi_r[i].p, 0, std::max<size_t>(P->n_rigid_edges, 1) * 8, st));
            ADB_CUDA(cudaMemsetAsync(s->chi_m[i].p, 0, std::max<size_t>(P->n_motion_edges, 1) * 8, st));
        }
        if ((r = s->flag.ensure(std::max<size_t>(std::max(E, n_dyn()), 1))) != ADB_OK) return r;
        if ((r = s->scal.ensure(sizeof(Scalars))) != ADB_OK) return r;
        if ((r = s->lm.ensure(sizeof(Lm))) != ADB_OK) return r;
        if ((r = s->pose_partial.ensure(std::max<size_t>(P->n_poses, 1) * kPoseChunks * 27 * sizeof(double))) != ADB_OK) return r;
        if ((r = s->trace.ensure(std::max<size_t>(R->trace ? R->trace_cap : 0, 1) * ADB_BA_TRACE_COLS * sizeof(double))) != ADB_OK) return r;
        ADB_CUDA(cudaMemsetAsync(s->scal.p, 0, sizeof(Scalars), st));
        lvl_e.assign(E, 0); lvl_j.assign(P->n_joint_edges, 0); lvl_r.assign(P->n_rigid_edges, 0); lvl_m.assign(P->n_motion_edges, 0);
        return ADB_OK;
    }

    // SparseOptimizer::initializeOptimization(0): active sets, dense layout, Schur pair list; uploads them
    adb_status build_layout() {
        const int E = P->n_edges, NP = P->n_points;
        std::vector<uint8_t> ap(P->n_poses, 0), aj(P->n_joints, 0), ad(P->n_dists, 0), am(P->n_motions, 0);
        act_point.assign(NP, 0);
        for (int k = 0; k < E; ++k) if (!lvl_e[k]) { ap[se_pose[k]] = 1; act_point[se_point[k]] = 1; }
        for (int e = 0; e < P->n_joint_edges; ++e) if (!lvl_j[e]) { ap[P->jedge_pose[e]] = 1; aj[P->jedge_joint[e]] = 1; }
        for (int e = 0; e < P->n_rigid_edges; ++e) if (!lvl_r[e]) { aj[P->redge_i[e]] = 1; aj[P->redge_j[e]] = 1; ad[P->redge_dist[e]] = 1; }
        for (int e = 0; e < P->n_motion_edges; ++e) if (!lvl_m[e]) { aj[P->medge_p1[e]] = 1; aj[P->medge_p2[e]] = 1; am[P->medge_motion[e]] = 1; }
        int o = 0;
        off_pose.assign(P->n_poses, -1); off_dist.assign(P->n_dists, -1); off_motion.assign(P->n_motions, -1); off_joint.assign(P->n_joints, -1);
        for (int i = 0; i < P->n_poses; ++i) if (ap[i] && !P->pose_fixed[i]) { off_pose[i] = o; o += 6; }
        for (int i = 0; i < P->n_dists; ++i) if (ad[i]) { off_dist[i] = o; o += 1; }
        for (int i = 0; i < P->n_motions; ++i) if (am[i]) { off_motion[i] = o; o += 6; }
        for (int i = 0; i < P->n_joints; ++i) if (aj[i]) { off_joint[i] = o; o += 3; }
        nd = o;
        free_pose.clear();
        for (int i = 0; i < P->n_poses; ++i) if (off_pose[i] >= 0) free_pose.push_back(i);
        cudaStream_t st = s->stream;
        adb_status r;
#define UP(buf, v) if ((r = upload(buf, (v).data(), (v).size(), st)) != ADB_OK) return r
        UP(s->e_level, lvl_e); UP(s->j_level, lvl_j); UP(s->r_level, lvl_r); UP(s->m_level, lvl_m); UP(s->act_point, act_point);
        UP(s->off_pose, off_pose); UP(s->off_dist, off_dist); UP(s->off_motion, off_motion); UP(s->off_joint, off_joint); UP(s->free_pose, free_pose);
#undef UP
        // Schur pair list on the device (schur_pairs_kernel / schur_scan_kernel); the host only needs upper bounds for the buffers and the grid
        {
            int nblk = 0;
            for (int i = 0; i < P->n_poses; ++i) if (off_pose[i] >= 0) nblk = std::max(nblk, off_pose[i] / 6 + 1);
            const int nkeys = nblk * (nblk + 1) / 2;
            size_t pairs_ub = 0;
            for (int l = 0; l < NP; ++l) { const size_t n = (size_t)(ptr[l + 1] - ptr[l]); pairs_ub += n * (n + 1) / 2; }
            chunks_ub = nkeys > 0 && E > 0 ? (int)(pairs_ub / kSchurChunk) + nkeys : 0;
            if ((r = s->pair_totals.ensure(2 * sizeof(int))) != ADB_OK) return r;
            ADB_CUDA(cudaMemsetAsync(s->pair_totals.p, 0, 2 * sizeof(int), st));
            if (chunks_ub > 0) {
                if ((r = s->pairs.ensure(std::max<size_t>(pairs_ub, 1) * sizeof(int2))) != ADB_OK) return r;
                if ((r = s->chunks.ensure((size_t)chunks_ub * sizeof(int2))) != ADB_OK) return r;
                if ((r = s->pair_count.ensure((size_t)nkeys * sizeof(int))) != ADB_OK) return r;
                if ((r = s->pair_cursor.ensure((size_t)nkeys * sizeof(int))) != ADB_OK) return r;
                ADB_CUDA(cudaMemsetAsync(s->pair_count.p, 0, (size_t)nkeys * sizeof(int), st));
                const int grid = grid_for((size_t)NP * kPtLanes, 128);
                schur_pairs_kernel<<<grid, 128, 0, st>>>(s->point_ptr.as<int>(), NP, s->e_pose.as<int>(), s->e_level.as<uint8_t>(), s->off_pose.as<int>(),
                                                         s->pair_count.as<int>(), nullptr, 0);
                schur_scan_kernel<<<1, 1024, 0, st>>>(s->pair_count.as<int>(), nkeys, s->pair_cursor.as<int>(), s->chunks.as<int2>(), s->pair_totals.as<int>());
                schur_pairs_kernel<<<grid, 128, 0, st>>>(s->point_ptr.as<int>(), NP, s->e_pose.as<int>(), s->e_level.as<uint8_t>(), s->off_pose.as<int>(),
                                                         s->pair_cursor.as<int>(), s->pairs.as<int2>(), 1);
                s->launches += 3;
                ADB_CUDA(cudaGetLastError());
            }
        }
        const size_t n2 = std::max<size_t>((size_t)nd * nd, 1);
        ld = chol_nblk(std::max(nd, 1)) * kCholNB;
        if ((r = s->H.ensure(n2 * 8)) != ADB_OK) return r;
        if ((r = s->b.ensure(std::max(nd, 1) * 8)) != ADB_OK) return r;
        if ((r = s->bs.ensure((size_t)ld * 8)) != ADB_OK) return r;
        if (nd > 0) {
            if ((r = s->Sm.ensure(chol_elems(nd) * 8)) != ADB_OK) return r;
            if ((r = s->work.ensure(chol_scratch_elems(nd) * 8)) != ADB_OK) return r;
        }
        return ADB_OK;
    }

    adb_status read_lm() {
        ADB_CUDA(cudaMemcpyAsync(s->h_lm, s->lm.p, sizeof(Lm), cudaMemcpyDeviceToHost, s->stream));
        ADB_CUDA(cudaStreamSynchronize(s->stream));
        hl = *s->h_lm;
        return ADB_OK;
    }

    // One LM step, enqueued without knowing what the previous one decided: buildSystem at the accepted state (skipped on the
    // device unless the step opens an outer iteration), one trial with the controller's lambda into the other state buffers,
    // evaluation, decision.  Every kernel returns at once when the round is already over.
    adb_status enqueue_step(bool robust, bool first_step) {
        cudaStream_t st = s->stream;
        const int E = P->n_edges, NP = P->n_points;
        Scalars* sc = s->scal.as<Scalars>();
        Lm* lm = dlm();
        const StatePair SP = states();
        const ChiPair CH = chis();
        const size_t n2 = (size_t)nd * nd;
        tm.begin(0);
        ba_clear_kernel<<<std::min(grid_for(std::max<size_t>(n2, 1), 256), 592), 256, 0, st>>>(s->H.as<double>(), n2, s->b.as<double>(), nd, lm, sc);
        ++s->launches;
        const bool pose_side = E > 0 && !free_pose.empty();
        if (pose_side) {   // fork: the pose side (J_pose^T W J_pose per free pose) runs on the second stream next to the landmark side
            ADB_CUDA(cudaEventRecord(s->fork_ev, st));
            ADB_CUDA(cudaStreamWaitEvent(s->stream2, s->fork_ev, 0));
            const int nfp = (int)free_pose.size();
            ba_pose_linearize_kernel<<<nfp * kPoseChunks, 256, 0, s->stream2>>>(cam(), opt(robust), sedges(), SP, lm, s->pose_ptr.as<int>(), s->pose_edges.as<int>(),
                                                                               s->free_pose.as<int>(), s->pose_partial.as<double>());
            ba_pose_reduce_kernel<<<nfp, 32, 0, s->stream2>>>(s->pose_partial.as<double>(), s->free_pose.as<int>(), s->off_pose.as<int>(), nd, lm, s->H.as<double>(),
                                                              s->b.as<double>());
            ADB_CUDA(cudaEventRecord(s->join_ev, s->stream2));
            s->launches += 2;
        }
        if (NP > 0) {
            ba_point_linearize_kernel<<<grid_for((size_t)NP * kPtLanes, kBaThreads), kBaThreads, 0, st>>>(cam(), opt(robust), sedges(), SP, lm, s->point_ptr.as<int>(), NP,
                                                                                      s->off_pose.as<int>(), s->Hll.as<double>(), s->bl.as<double>(),
                                                                                      s->W.as<double>(), CH, sc);
            ++s->launches;
        }
        if (pose_side) ADB_CUDA(cudaStreamWaitEvent(st, s->join_ev, 0));   // join
        if (n_dyn() > 0) {
            ba_dyn_kernel<<<grid_for(n_dyn(), kBaThreads), kBaThreads, 0, st>>>(cam(), opt(robust), dedges(), SP, lm, doff(), nd, s->H.as<double>(),
                                                                               s->b.as<double>(), CH, sc, 0);
            ++s->launches;
        }
        tm.end();
        tm.begin(4);
        if (first_step) {   // computeLambdaInit: only the first iteration of a round reads it
            ba_maxdiag_kernel<<<grid_for(nd + NP, 256), 256, 0, st>>>(s->H.as<double>(), nd, s->Hll.as<double>(), s->act_point.as<uint8_t>(), NP, lm, sc);
            ++s->launches;
        }
        lm_iter_kernel<<<1, 1, 0, st>>>(lm, sc);
        ++s->launches;
        tm.end();
        tm.begin(1);
        if (nd > 0) {
            ba_prepare_kernel<<<grid_for((size_t)ld * ld, 256), 256, 0, st>>>(s->H.as<double>(), s->b.as<double>(), nd, ld, lm, s->Sm.as<double>(), s->bs.as<double>());
            ++s->launches;
        }
        if (NP > 0) {
            ba_dinv_kernel<<<grid_for(NP, 128), 128, 0, st>>>(s->Hll.as<double>(), s->bl.as<double>(), s->act_point.as<uint8_t>(), NP, lm, s->Dinv.as<double>(),
                                                              s->db.as<double>());
            ++s->launches;
        }
        if (chunks_ub > 0) {
            ba_schur_block_kernel<<<grid_for(chunks_ub, kBaThreads / 32), kBaThreads, 0, st>>>(s->pairs.as<int2>(), s->chunks.as<int2>(), s->pair_totals.as<int>(), s->e_pose.as<int>(),
                                                                                        s->e_point.as<int>(), s->off_pose.as<int>(), s->W.as<double>(),
                                                                                        s->Dinv.as<double>(), s->db.as<double>(), ld, lm, s->Sm.as<double>(),
                                                                                        s->Sm.as<double>() + (size_t)ld * ld);   // rhs row
            ++s->launches;
        }
        ADB_CUDA(cudaGetLastError());
        tm.end();
        tm.begin(2);
        if (nd > 0) {
            // LinearSolverDense / LinearSolverEigen::solve: one cluster launch factors, substitutes and writes x to bs
            const adb_status cs = chol_solve_launch(st, s->Sm.as<double>(), ld, ld / kCholNB, s->work.as<double>(), s->bs.as<double>(), &sc->info, 0, &lm->done);
            if (cs != ADB_OK) return cs;
            ++s->launches;
        }
        tm.end();
        tm.begin(3);
        {
            const DenseSizes N{P->n_poses, P->n_dists, P->n_motions, P->n_joints};
            const int nthreads = std::max(nd, P->n_poses + P->n_dists + P->n_motions + P->n_joints);
            const int dense_blocks = grid_for(nthreads, kBaThreads), point_blocks = NP > 0 ? grid_for((size_t)NP * kPtLanes, kBaThreads) : 0;
            ba_backsub_kernel<<<dense_blocks + point_blocks, kBaThreads, 0, st>>>(s->point_ptr.as<int>(), s->e_pose.as<int>(), s->e_level.as<uint8_t>(),
                                                                                 s->off_pose.as<int>(), s->act_point.as<uint8_t>(), NP, s->W.as<double>(),
                                                                                 s->Dinv.as<double>(), s->bl.as<double>(), s->bs.as<double>(), lm, SP, sc, N, doff(),
                                                                                 s->b.as<double>(), nd, dense_blocks);
            ++s->launches;
        }
        if (E > 0) {
            ba_eval_kernel<<<grid_for(E, kBaThreads), kBaThreads, 0, st>>>(cam(), opt(robust), sedges(), SP, lm, CH, sc);
            ++s->launches;
        }
        if (n_dyn() > 0) {
            ba_dyn_kernel<<<grid_for(n_dyn(), kBaThreads), kBaThreads, 0, st>>>(cam(), opt(robust), dedges(), SP, lm, doff(), nd, nullptr, nullptr, CH, sc, 1);
            ++s->launches;
        }
        tm.end();
        tm.begin(4);
        lm_decide_kernel<<<1, 1, 0, st>>>(lm, sc, R->trace ? s->trace.as<double>() : nullptr);
        ++s->launches;
        ADB_CUDA(cudaGetLastError());
        tm.end();
        return ADB_OK;
    }

    // SparseOptimizer::optimize(iterations) + OptimizationAlgorithmLevenberg::solve, controlled on the device (struct Lm).  The host
    // enqueues `iterations` steps -- enough when every trial is accepted, the common case -- reads the controller back once, and
    // tops up while rejected trials have used up steps.  The stop flag is polled between those batches (the reference polls it
    // per iteration and per trial; a flag raised mid-batch takes effect at the next read-back).
    adb_status optimize(int iterations, bool robust, int round, double* chi_out) {
        adb_status r;
        hl.need_lin = 1; hl.done = iterations <= 0 ? 1 : 0; hl.first = 1; hl.it = 0; hl.it_max = iterations; hl.q = 0; hl.max_trials = O->max_trials;
        hl.nbad = 0; hl.iterations_run = 0; hl.round = round; hl.tau = O->tau; hl.trace_cap = R->trace ? R->trace_cap : 0;
        if (round == 0) { hl.cur = 0; hl.chi_last = 0; hl.trials_run = 0; hl.trace_len = 0; hl.lambda = 0; hl.ni = 2; hl.current = 0; hl.ini = 0; hl.chi2_initial = 0; }
        *s->h_lm = hl;
        ADB_CUDA(cudaMemcpyAsync(s->lm.p, s->h_lm, sizeof(Lm), cudaMemcpyHostToDevice, s->stream));
        int batch = iterations;
        bool first_step = true;
        while (!hl.done && !stopped()) {
            for (int k = 0; k < batch; ++k) { if ((r = enqueue_step(robust, first_step)) != ADB_OK) return r; first_step = false; }
            if ((r = read_lm()) != ADB_OK) return r;
            batch = 2;
        }
        R->iterations_run[round] = hl.iterations_run;
        R->trials_run = hl.trials_run;
        *chi_out = hl.current;
        return ADB_OK;
    }

    // gates of all four edge families from the chi2 of the last evaluation; result in host vectors (sorted order for static edges)
    adb_status gates(std::vector<uint8_t>& fe, std::vector<uint8_t>& fj, std::vector<uint8_t>& fr, std::vector<uint8_t>& fm, std::vector<double>* chi_out) {
        cudaStream_t st = s->stream;
        const int E = P->n_edges;
        fe.assign(E, 0);
        if (E > 0) {
            ba_gate_kernel<<<grid_for(E, 256), 256, 0, st>>>(sedges(), states(), dlm(), chis(), O->chi2_mono, O->chi2_stereo, s->flag.as<uint8_t>());
            ++s->launches;
            ADB_CUDA(cudaGetLastError());
            ADB_CUDA(cudaMemcpyAsync(fe.data(), s->flag.p, E, cudaMemcpyDeviceToHost, st));
            if (chi_out) { chi_out->assign(E, 0); ADB_CUDA(cudaMemcpyAsync(chi_out->data(), s->chi_e[hl.chi_last].p, (size_t)E * 8, cudaMemcpyDeviceToHost, st)); }
        }
        // the few hundred dense-block edges are gated on the host from their chi2 and the current state
        const int cur = hl.cur, chi_last = hl.chi_last;   // as of the read-back that ended the round
        std::vector<double> cj(P->n_joint_edges), cr(P->n_rigid_edges), cm(P->n_motion_edges), hq, ht, hj;
        if (P->n_joint_edges) {
            ADB_CUDA(cudaMemcpyAsync(cj.data(), s->chi_j[chi_last].p, cj.size() * 8, cudaMemcpyDeviceToHost, st));
            hq.resize(4 * (size_t)P->n_poses); ht.resize(3 * (size_t)P->n_poses); hj.resize(3 * (size_t)P->n_joints);
            ADB_CUDA(cudaMemcpyAsync(hq.data(), s->pq[cur].p, hq.size() * 8, cudaMemcpyDeviceToHost, st));
            ADB_CUDA(cudaMemcpyAsync(ht.data(), s->pt[cur].p, ht.size() * 8, cudaMemcpyDeviceToHost, st));
            ADB_CUDA(cudaMemcpyAsync(hj.data(), s->Jt[cur].p, hj.size() * 8, cudaMemcpyDeviceToHost, st));
        }
        if (P->n_rigid_edges) ADB_CUDA(cudaMemcpyAsync(cr.data(), s->chi_r[chi_last].p, cr.size() * 8, cudaMemcpyDeviceToHost, st));
        if (P->n_motion_edges) ADB_CUDA(cudaMemcpyAsync(cm.data(), s->chi_m[chi_last].p, cm.size() * 8, cudaMemcpyDeviceToHost, st));
        ADB_CUDA(cudaStreamSynchronize(st));
        fj.assign(P->n_joint_edges, 0); fr.assign(P->n_rigid_edges, 0); fm.assign(P->n_motion_edges, 0);
        for (int e = 0; e < P->n_joint_edges; ++e) {
            double Rm[9];
            const int ip = P->jedge_pose[e];
            quat_to_rot(&hq[4 * ip], Rm);
            const double* X = &hj[3 * (size_t)P->jedge_joint[e]];
            const double z = Rm[6] * X[0] + Rm[7] * X[1] + Rm[8] * X[2] + ht[3 * ip + 2];
            fj[e] = cj[e] > O->chi2_stereo || !(z > 0);
        }
        for (int e = 0; e < P->n_rigid_edges; ++e) fr[e] = cr[e] > O->chi2_rigid;
        for (int e = 0; e < P->n_motion_edges; ++e) fm[e] = cm[e] > O->chi2_motion;
        return ADB_OK;
    }

    adb_status download_state() {
        cudaStream_t st = s->stream;
        const int cur = hl.cur;
#define DN(dst, buf, n) if ((n) > 0) ADB_CUDA(cudaMemcpyAsync(dst, buf.p, (size_t)(n) * 8, cudaMemcpyDeviceToHost, st))
        DN(P->pose_q, s->pq[cur], 4 * (size_t)P->n_poses); DN(P->pose_t, s->pt[cur], 3 * (size_t)P->n_poses); DN(P->points, s->X[cur], 3 * (size_t)P->n_points);
        DN(P->joints, s->Jt[cur], 3 * (size_t)P->n_joints); DN(P->dists, s->Dd[cur], P->n_dists);
        DN(P->motion_q, s->mq[cur], 4 * (size_t)P->n_motions); DN(P->motion_t, s->mt[cur], 3 * (size_t)P->n_motions);
#undef DN
        ADB_CUDA(cudaStreamSynchronize(st));
        return ADB_OK;
    }
};

}  // namespace

extern "C" {

void adb_ba_default_options(adb_ba_options* o) {
    if (!o) return;
    o->iterations[0] = 5; o->iterations[1] = 10; o->max_trials = 10; o->tau = 1e-5;
    o->chi2_mono = 5.991; o->chi2_stereo = 7.815; o->chi2_rigid = 1.0; o->chi2_motion = 4.0;
    o->huber_mono = (double)(float)std::sqrt(5.991); o->huber_stereo = (double)(float)std::sqrt(7.815);
    o->huber_rigid = 1.0; o->huber_motion = (double)(float)std::sqrt(4.0);
    o->robust[0] = 1; o->robust[1] = 0;
}

void adb_ba_global_options(adb_ba_options* o, int32_t n_iterations, int32_t robust) {
    adb_ba_default_options(o);
    o->iterations[0] = n_iterations; o->iterations[1] = 0;
    o->huber_mono = (double)(float)std::sqrt(5.99);   // thHuber2D of BundleAdjustment (src/Optimizer.cc:84): 5.99, not 5.991
    o->robust[0] = robust ? 1 : 0;
}

void adb_ba_pose_from_tcw(const float* T, double* q, double* t) {
    double R[9];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = (double)T[i * 4 + j]; t[i] = (double)T[i * 4 + 3]; }
    rot_to_quat(R, q);
    quat_normalize_pos(q);
}

void adb_ba_pose_to_tcw(const double* q, const double* t, float* T) {
    double R[9];
    quat_to_rot(q, R);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float)R[i * 3 + j]; T[i * 4 + 3] = (float)t[i]; }
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

adb_status adb_ba_create(int32_t device, adb_ba_t* out) {
    ADB_CHECK(out, ADB_ERR_INVALID, "null argument");
    *out = nullptr;
    adb_status st = select_device(device);
    if (st != ADB_OK) return st;
    adb_ba* s = new adb_ba();
    s->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->fork_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->join_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost(&s->h_lm, sizeof(Lm));
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "ba create", __FILE__, __LINE__); }
    *out = s;
    return ADB_OK;
}

adb_status adb_ba_destroy(adb_ba_t s) {
    if (!s) return ADB_OK;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    DevBuf* all[] = {&s->pq[0], &s->pq[1], &s->pt[0], &s->pt[1], &s->X[0], &s->X[1], &s->Jt[0], &s->Jt[1], &s->Dd[0], &s->Dd[1], &s->mq[0], &s->mq[1],
                     &s->mt[0], &s->mt[1], &s->e_pose, &s->e_point, &s->e_obs, &s->e_info, &s->e_level, &s->point_ptr, &s->pairs, &s->chunks, &s->pair_count, &s->pair_cursor, &s->pair_totals, &s->pose_ptr, &s->pose_edges, &s->free_pose, &s->off_pose,
                     &s->act_point, &s->j_pose, &s->j_joint, &s->j_obs, &s->j_info, &s->j_level, &s->r_i, &s->r_j, &s->r_d, &s->r_info, &s->r_level,
                     &s->m_p1, &s->m_p2, &s->m_m, &s->m_dt, &s->m_info, &s->m_level, &s->off_joint, &s->off_dist, &s->off_motion, &s->H, &s->b,
                     &s->Sm, &s->bs, &s->Hll, &s->bl, &s->W, &s->Dinv, &s->db, &s->chi_e[0], &s->chi_e[1], &s->chi_j[0], &s->chi_j[1], &s->chi_r[0],
                     &s->chi_r[1], &s->chi_m[0], &s->chi_m[1], &s->flag, &s->scal, &s->work, &s->lm, &s->trace, &s->pose_partial};
    for (DevBuf* b : all) b->release();
    for (cudaEvent_t e : s->tev) cudaEventDestroy(e);
    if (s->h_lm) cudaFreeHost(s->h_lm);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    cudaStreamDestroy(s->stream);
    if (s->stream2) { cudaStreamSynchronize(s->stream2); cudaStreamDestroy(s->stream2); }
    if (s->fork_ev) cudaEventDestroy(s->fork_ev);
    if (s->join_ev) cudaEventDestroy(s->join_ev);
    cudaGetLastError();
    delete s;
    return ADB_OK;
}

adb_status adb_ba_solve(adb_ba_t s, adb_ba_problem* P, const adb_ba_options* O, volatile const uint8_t* stop, adb_ba_result* R) {
    ADB_CHECK(s && P && O && R, ADB_ERR_INVALID, "null argument");
    ADB_CHECK(P->n_poses >= 1 && P->n_points >= 0 && P->n_edges >= 0, ADB_ERR_INVALID, "empty problem");
    if (stop && *stop) { set_error("stop flag set before optimisation"); return ADB_ERR_STOPPED; }
    ADB_CUDA(cudaSetDevice(s->device));
    Ctx c(s, P, O, R, stop);
    R->iterations_run[0] = R->iterations_run[1] = 0; R->trials_run = 0; R->stopped = 0; R->trace_len = 0;
    R->chi2_initial = 0; R->chi2_round[0] = R->chi2_round[1] = 0; R->lambda_final = 0;
    adb_status r;
    const bool timing = getenv("ADB_BA_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto t0 = now();
    if ((r = c.upload_problem()) != ADB_OK) return r;
    auto t1 = now();
    if ((r = c.build_layout()) != ADB_OK) return r;
    auto t2 = now();
    double chi = 0;
    if ((r = c.optimize(O->iterations[0], O->robust[0] != 0, 0, &chi)) != ADB_OK) return r;
    auto t3 = now();
    if (timing) fprintf(stderr, "[adb_ba] upload %.2f ms, layout %.2f ms, round-1 optimise %.2f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
    R->chi2_round[0] = chi;
    std::vector<uint8_t> fe, fj, fr, fm;
    const bool more = !c.stopped() && O->iterations[1] > 0;
    if (c.stopped()) R->stopped = 1;
    if (more) {
        if ((r = c.gates(fe, fj, fr, fm, nullptr)) != ADB_OK) return r;
        for (size_t k = 0; k < fe.size(); ++k) if (fe[k]) c.lvl_e[k] = 1;
        for (size_t k = 0; k < fj.size(); ++k) if (fj[k]) c.lvl_j[k] = 1;
        for (size_t k = 0; k < fr.size(); ++k) if (fr[k]) c.lvl_r[k] = 1;
        for (size_t k = 0; k < fm.size(); ++k) if (fm[k]) c.lvl_m[k] = 1;
        // the dense layout may shrink: H offsets change, but the state buffers and chi2 arrays stay
        if ((r = c.build_layout()) != ADB_OK) return r;
        if ((r = c.optimize(O->iterations[1], O->robust[1] != 0, 1, &chi)) != ADB_OK) return r;
        R->chi2_round[1] = chi;
        if (c.stopped()) R->stopped = 1;
    }
    auto t4 = now();
    R->lambda_final = c.hl.lambda;
    R->chi2_initial = c.hl.chi2_initial;
    R->trace_len = c.hl.trace_len;
    if (R->trace && c.hl.trace_len > 0)
        ADB_CUDA(cudaMemcpyAsync(R->trace, s->trace.p, (size_t)c.hl.trace_len * ADB_BA_TRACE_COLS * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    std::vector<double> chi_sorted;
    if ((r = c.gates(fe, fj, fr, fm, R->edge_chi2 ? &chi_sorted : nullptr)) != ADB_OK) return r;
    for (int k = 0; k < P->n_edges; ++k) {
        if (R->edge_outlier) R->edge_outlier[c.perm[k]] = fe[k];
        if (R->edge_chi2) R->edge_chi2[c.perm[k]] = chi_sorted[k];
    }
    if (R->jedge_outlier) for (int e = 0; e < P->n_joint_edges; ++e) R->jedge_outlier[e] = fj[e];
    if (R->redge_outlier) for (int e = 0; e < P->n_rigid_edges; ++e) R->redge_outlier[e] = fr[e];
    if (R->medge_outlier) for (int e = 0; e < P->n_motion_edges; ++e) R->medge_outlier[e] = fm[e];
    if ((r = c.download_state()) != ADB_OK) return r;
    c.tm.collect();
    if (timing) fprintf(stderr, "[adb_ba] gates + round 2 %.2f ms, final gates + write-back %.2f ms\n", ms(t3, t4), ms(t4, now()));
    return ADB_OK;
}


adb_status adb_pose_optimize(adb_ba_t s, adb_pose_problem* P) {
    ADB_CHECK(s && P && P->frame_ptr && P->pose_q && P->pose_t && P->outlier && P->n_inliers, ADB_ERR_INVALID, "null argument");
    ADB_CHECK(P->n_frames >= 1, ADB_ERR_INVALID, "no frames");
    ADB_CUDA(cudaSetDevice(s->device));
    const int F = P->n_frames, n = P->frame_ptr[F];
    ADB_CHECK(n >= 0 && P->frame_ptr[0] == 0, ADB_ERR_INVALID, "bad frame_ptr");
    for (int f = 0; f < F; ++f) ADB_CHECK(P->frame_ptr[f + 1] >= P->frame_ptr[f], ADB_ERR_INVALID, "frame_ptr is not monotone at %d", f);
    cudaStream_t st = s->stream;
    // reuse BA buffers as scratch: e_pose = frame_ptr, pq/pt[0] = poses, e_obs = xw | obs | inv_sigma2 (float), chi_e[0], flag, e_level, off_pose = n_inliers
    adb_status r;
    if ((r = upload(s->e_pose, P->frame_ptr, (size_t)F + 1, st)) != ADB_OK) return r;
    if ((r = upload(s->pq[0], P->pose_q, 4 * (size_t)F, st)) != ADB_OK) return r;
    if ((r = upload(s->pt[0], P->pose_t, 3 * (size_t)F, st)) != ADB_OK) return r;
    if ((r = s->e_obs.ensure(std::max<size_t>(n, 1) * 7 * sizeof(float))) != ADB_OK) return r;
    float* d_xw = s->e_obs.as<float>();
    float* d_obs = d_xw + 3 * (size_t)std::max(n, 1);
    float* d_w = d_obs + 3 * (size_t)std::max(n, 1);
    if (n > 0) {
        ADB_CHECK(P->xw && P->obs && P->inv_sigma2, ADB_ERR_INVALID, "null correspondence arrays");
        ADB_CUDA(cudaMemcpyAsync(d_xw, P->xw, 3 * (size_t)n * 4, cudaMemcpyHostToDevice, st));
        ADB_CUDA(cudaMemcpyAsync(d_obs, P->obs, 3 * (size_t)n * 4, cudaMemcpyHostToDevice, st));
        ADB_CUDA(cudaMemcpyAsync(d_w, P->inv_sigma2, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    if ((r = s->chi_e[0].ensure(std::max<size_t>(n, 1) * 8)) != ADB_OK) return r;
    if ((r = s->flag.ensure(std::max<size_t>(n, 1))) != ADB_OK) return r;
    if ((r = s->e_level.ensure(std::max<size_t>(n, 1))) != ADB_OK) return r;
    if ((r = s->off_pose.ensure((size_t)F * 4)) != ADB_OK) return r;
    PoseArgs A{Cam{P->fx, P->fy, P->cx, P->cy, P->bf}, s->e_pose.as<int>(), s->pq[0].as<double>(), s->pt[0].as<double>(), d_xw, d_obs, d_w,
               s->flag.as<uint8_t>(), s->off_pose.as<int>()};
    pose_optimize_kernel<<<F, kPoseThreads, 0, st>>>(A, s->chi_e[0].as<double>(), s->e_level.as<uint8_t>());
    ++s->launches;
    ADB_CUDA(cudaGetLastError());
    ADB_CUDA(cudaMemcpyAsync(P->pose_q, s->pq[0].p, 4 * (size_t)F * 8, cudaMemcpyDeviceToHost, st));
    ADB_CUDA(cudaMemcpyAsync(P->pose_t, s->pt[0].p, 3 * (size_t)F * 8, cudaMemcpyDeviceToHost, st));
    if (n > 0) ADB_CUDA(cudaMemcpyAsync(P->outlier, s->flag.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    ADB_CUDA(cudaMemcpyAsync(P->n_inliers, s->off_pose.p, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
    ADB_CUDA(cudaStreamSynchronize(st));
    return ADB_OK;
}

adb_status adb_ba_leaf_eval(adb_ba_t s, const adb_ba_leaf_io* io) {
    ADB_CHECK(s && io && io->n >= 1 && io->out, ADB_ERR_INVALID, "null argument");
    ADB_CUDA(cudaSetDevice(s->device));
    const int n = io->n;
    const double* src[14] = {io->pose_q, io->pose_t, io->x, io->obs, io->pose_update, io->joint_a, io->joint_b, io->bone, io->motion_q, io->motion_t, io->motion_dt,
                             io->motion_update, io->huber_delta, io->huber_e2};
    const int width[14] = {4, 3, 3, 3, 6, 3, 3, 1, 4, 3, 1, 6, 1, 1};
    ADB_CHECK((io->huber_delta == nullptr) == (io->huber_e2 == nullptr), ADB_ERR_INVALID, "huber_delta and huber_e2 go together");
    const int n_in = io->huber_delta ? 14 : 12;
    size_t total = 0;
    for (int k = 0; k < n_in; ++k) { ADB_CHECK(src[k], ADB_ERR_INVALID, "null input array %d", k); total += (size_t)width[k] * n; }
    adb_status r;
    if ((r = s->work.ensure((total + (size_t)kLeafRecord * n) * sizeof(double))) != ADB_OK) return r;
    double* d = s->work.as<double>();
    const double* dev[14] = {};
    for (int k = 0; k < n_in; ++k) {
        ADB_CUDA(cudaMemcpyAsync(d, src[k], (size_t)width[k] * n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        dev[k] = d; d += (size_t)width[k] * n;
    }
    ba_leaf_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(n, Cam{io->fx, io->fy, io->cx, io->cy, io->bf}, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7],
                                                          dev[8], dev[9], dev[10], dev[11], dev[12], dev[13], d);
    ++s->launches;
    ADB_CUDA(cudaGetLastError());
    ADB_CUDA(cudaMemcpyAsync(io->out, d, (size_t)kLeafRecord * n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    ADB_CUDA(cudaStreamSynchronize(s->stream));
    return ADB_OK;
}

adb_status adb_ba_stage_ms(adb_ba_t s, float* ms6) {
    ADB_CHECK(s && ms6, ADB_ERR_INVALID, "null argument");
    for (int i = 0; i < 6; ++i) ms6[i] = s->stage_ms[i];
    return ADB_OK;
}

int64_t adb_ba_launch_count(adb_ba_t s) { return s ? s->launches : 0; }

}  // extern "C"
