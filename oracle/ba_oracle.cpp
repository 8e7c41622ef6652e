// oracle/ba_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the bundle-adjustment arithmetic on the AirDOS hot path, used only as the
// parity checker (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
// It restates, without g2o or Eigen:
//   Optimizer::LocalBundleAdjustment                         src/Optimizer.cc:431-731
//   Optimizer::LocalBundleAdjustmentHumanTrajactory          src/Optimizer.cc:1496-2222 (solver part)
//   OptimizationAlgorithmLevenberg::solve                    Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
//   BlockSolver::buildSystem / setLambda / solve             Thirdparty/g2o/g2o/core/block_solver.hpp:354-604
//   Edge(Stereo)SE3ProjectXYZ error + Jacobians              Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:103-234
//   SE3Quat::exp / operator* / normalizeRotation             Thirdparty/g2o/g2o/types/se3quat.h:104-110, 217-285
//   RobustKernelHuber::robustify, robustInformation          core/robust_kernel_impl.cpp:78-92, core/base_edge.h:96-102
//   VertexSE3 / VertexDistanceDouble / EdgeRigidBodyDouble / LandmarkMotionTernaryEdge
//                                                            include/g2o_vertex_se3.h, g2o_vertex_distance.h,
//                                                            g2o_edge_rigidbody.h, g2o_dyn_slam3d.h
// Parity pinning.  The reference ships no tests or golden vectors for BA; it is pinned by the reference's own
// functions compiled here from /root/reference against stand-ins (Eigen / OpenCV proper are absent):
//   leaf arithmetic      g2o / AirDOS type sources unmodified (oracle/ref_leaf.cpp)            -> tests/golden/ba_leaf_ref.npz, 1e-12
//   LM control           optimization_algorithm_levenberg.cpp whole + SparseOptimizer::optimize (oracle/ref_lm.cpp), run over the steps of
//                        Solver / PoseSolver below through function pointers                   -> lm_ref.npz: trials, lambda, state bit for bit
//   Huber kernel         RobustKernelHuber::setDelta / robustify (float dsqr member)           -> lm_ref.npz, bit for bit
//   quadratic form       BaseBinaryEdge / BaseUnaryEdge::constructQuadraticForm                -> lm_ref.npz, 1e-12
//   schedules and gates  Optimizer::LocalBundleAdjustment, ::BundleAdjustment, ::PoseOptimization, ::LocalBundleAdjustmentHumanTrajactory whole, with the reference's edge types,
//                        Converter and LM control over this file's solver steps (oracle/ref_lba.cpp) -> lba_ref.npz, pose_ref.npz: every
//                        trial, final state, erase list / mvbOutlier / return value bit for bit
//   Schur complement     BlockSolver<Traits>::solve() whole (core/block_solver.hpp:353-483: marginalisation of the landmarks, reduced
//                        right-hand side, landmark back-substitution), compiled between stand-in block containers (oracle/ref_schur.cpp)
//                        on the systems Solver::build_system assembles, as BlockSolver_6_3 (static windows) and as BlockSolverX
//                        (articulated windows: run-time block widths 6 / 3 / 1)                 -> schur_ref.npz: the whole update x within
//                        1e-12 (measured 2.5e-14), the oracle's pose update satisfies the reference's reduced system to 3e-15
// What stays "parity unpinned" against the literal reference: the linear solver behind the reduced system -- it needs Eigen proper; the
// reduced system is solved by a dense Cholesky instead of Eigen's SimplicialLDLT / LDLT, which agrees to rounding (the residual check of
// schur_ref.npz involves no factorisation) -- and the rigidity / motion Jacobians (undefined in the reference, D.4 / D.6).  Those are
// pinned mathematically: finite-difference Jacobians, an independent numpy normal-equation solve and an independent numpy LM trajectory
// (tests/test_oracle_ba.py).
// Conventions for the reference's ill-defined corners: D.4 analytic
// rigidity Jacobian, D.5 motion prior = identity (the caller passes it), D.6 d(error)/d(motion
// translation) = delta_t * I with a zero rotation block.
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/airdos_b200.h"

namespace {

typedef double M3[9];   // row-major 3x3

inline void quat_to_rot(const double* q, double* R) {   // Eigen::Quaterniond::toRotationMatrix, q = x,y,z,w
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

inline void rot_to_quat(const double* m, double* q) {   // Eigen::Quaterniond(Matrix3d)
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}

inline void quat_normalize_pos(double* q) {   // SE3Quat::normalizeRotation
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
}

inline void quat_mul(const double* a, const double* b, double* r) {   // Eigen: a * b, x,y,z,w storage
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    r[3] = aw * bw - ax * bx - ay * by - az * bz;
    r[0] = aw * bx + ax * bw + ay * bz - az * by;
    r[1] = aw * by + ay * bw + az * bx - ax * bz;
    r[2] = aw * bz + az * bw + ax * by - ay * bx;
}

inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// T <- exp(delta) * T  (VertexSE3Expmap::oplusImpl), delta = (omega, upsilon)
void pose_oplus(double* q, double* t, const double* d) {
    const double wx = d[0], wy = d[1], wz = d[2];
    const double theta = std::sqrt(wx * wx + wy * wy + wz * wz);
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9];
    mat3_mul(O, O, O2);
    double R[9], V[9];
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i];
        std::memcpy(V, R, sizeof(R));
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
        const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 9; ++i) {
            R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * O[i] + b * O2[i];
            V[i] = (i % 4 == 0 ? 1.0 : 0.0) + b * O[i] + c * O2[i];
        }
    }
    double qe[4], te[3];
    rot_to_quat(R, qe);
    quat_normalize_pos(qe);
    for (int i = 0; i < 3; ++i) te[i] = V[i * 3] * d[3] + V[i * 3 + 1] * d[4] + V[i * 3 + 2] * d[5];
    // SE3Quat::operator*: t = te + Re * t ; r = qe * q ; normalizeRotation
    double Re[9];
    quat_to_rot(qe, Re);
    double tn[3];
    for (int i = 0; i < 3; ++i) tn[i] = te[i] + Re[i * 3] * t[0] + Re[i * 3 + 1] * t[1] + Re[i * 3 + 2] * t[2];
    double qn[4];
    quat_mul(qe, q, qn);
    quat_normalize_pos(qn);
    std::memcpy(q, qn, sizeof(qn));
    std::memcpy(t, tn, sizeof(tn));
}

// VertexSE3::oplusImpl (include/g2o_vertex_se3.h:113-122): H <- H * inc, inc = (t = d[0:3], q = compact d[3:6])
void motion_oplus(double* q, double* t, const double* d) {
    double R[9];
    quat_to_rot(q, R);
    double w = 1 - (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
    double Ri[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!(w < 0)) {
        const double qi[4] = {d[3], d[4], d[5], std::sqrt(w)};
        quat_to_rot(qi, Ri);
    }
    double Rn[9];
    mat3_mul(R, Ri, Rn);
    for (int i = 0; i < 3; ++i) t[i] += R[i * 3] * d[0] + R[i * 3 + 1] * d[1] + R[i * 3 + 2] * d[2];
    rot_to_quat(Rn, q);
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
}

// EdgeRigidBodyDouble::computeError (include/g2o_edge_rigidbody.h:139-149): |p_from - p_to| - d; `n` returns the norm
inline double rigid_error(const double* a, const double* c2, double dist, double* d3, double* n) {
    d3[0] = a[0] - c2[0]; d3[1] = a[1] - c2[1]; d3[2] = a[2] - c2[2];
    *n = std::sqrt(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
    return *n - dist;
}
// LandmarkMotionTernaryEdge::computeError with zero measurement (include/g2o_dyn_slam3d.h:65-76): e = p1 - M^-1 p2, M = (R, dt * t)
inline void motion_edge_error(const double* mq, const double* mt, double dt, const double* p1, const double* p2, double* er, double* Rm) {
    quat_to_rot(mq, Rm);
    const double d[3] = {p2[0] - dt * mt[0], p2[1] - dt * mt[1], p2[2] - dt * mt[2]};
    for (int i = 0; i < 3; ++i) er[i] = p1[i] - (Rm[i] * d[0] + Rm[3 + i] * d[1] + Rm[6 + i] * d[2]);   // R^T d
}
// its linearizeOplus (g2o_dyn_slam3d.h:78-101), first call on a fresh edge: I, -R^T, (dt I | 0)   (D.6)
inline void motion_edge_jacobians(const double* Rm, double dt, double* J1, double* J2, double* Jm) {
    for (int i = 0; i < 9; ++i) J1[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J2[i * 3 + j] = -Rm[j * 3 + i];
    std::memset(Jm, 0, 18 * sizeof(double));
    Jm[0] = dt; Jm[7] = dt; Jm[14] = dt;
}

// RobustKernelHuber keeps delta^2 in a FLOAT member (core/robust_kernel_impl.h:84 `float dsqr;`, setDelta :65-69)
struct Huber { double delta, dsqr; };
inline Huber make_huber(double d) { return Huber{d, (double)(float)(d * d)}; }
inline void robustify(const Huber& h, bool robust, double e2, double* rho0, double* rho1) {
    if (!robust || e2 <= h.dsqr) { *rho0 = e2; *rho1 = 1.0; }
    else { const double s = std::sqrt(e2); *rho0 = 2 * s * h.delta - h.dsqr; *rho1 = h.delta / s; }
}

// Reprojection residual (types_six_dof_expmap.cpp:141-157, .h:122-127).  Returns the dimension (2 or 3).
inline int reproj_error(const adb_ba_problem& P, const double* R, const double* t, const double* X, const double* obs, double* e,
                        double* Xc) {
    for (int i = 0; i < 3; ++i) Xc[i] = R[i * 3] * X[0] + R[i * 3 + 1] * X[1] + R[i * 3 + 2] * X[2] + t[i];
    if (obs[2] >= 0) {
        const float invz = (float)(1.0 / Xc[2]);              // 1.0f / double -> float
        const float bf = (float)P.bf;                          // const float& bf parameter
        const double u = Xc[0] * invz * P.fx + P.cx, v = Xc[1] * invz * P.fy + P.cy;
        const double ur = u - (double)(bf * invz);             // float product
        e[0] = obs[0] - u; e[1] = obs[1] - v; e[2] = obs[2] - ur;
        return 3;
    }
    e[0] = obs[0] - (Xc[0] / Xc[2] * P.fx + P.cx);
    e[1] = obs[1] - (Xc[1] / Xc[2] * P.fy + P.cy);
    e[2] = 0;
    return 2;
}

// Jacobians wrt point (Ji, dim x 3) and pose (Jj, dim x 6); rows beyond dim are zero.
inline void reproj_jacobians(const adb_ba_problem& P, const double* R, const double* Xc, int dim, double* Ji, double* Jj) {
    const double x = Xc[0], y = Xc[1], z = Xc[2], z2 = z * z, fx = P.fx, fy = P.fy, bf = P.bf;
    for (int c = 0; c < 3; ++c) {
        Ji[c] = -fx * R[c] / z + fx * x * R[6 + c] / z2;
        Ji[3 + c] = -fy * R[3 + c] / z + fy * y * R[6 + c] / z2;
        Ji[6 + c] = dim == 3 ? Ji[c] - bf * R[6 + c] / z2 : 0.0;
    }
    Jj[0] = x * y / z2 * fx; Jj[1] = -(1 + (x * x / z2)) * fx; Jj[2] = y / z * fx; Jj[3] = -1. / z * fx; Jj[4] = 0; Jj[5] = x / z2 * fx;
    Jj[6] = (1 + y * y / z2) * fy; Jj[7] = -x * y / z2 * fy; Jj[8] = -x / z * fy; Jj[9] = 0; Jj[10] = -1. / z * fy; Jj[11] = y / z2 * fy;
    if (dim == 3) {
        Jj[12] = Jj[0] - bf * y / z2; Jj[13] = Jj[1] + bf * x / z2; Jj[14] = Jj[2]; Jj[15] = Jj[3]; Jj[16] = 0; Jj[17] = Jj[5] - bf / z2;
    } else {
        for (int i = 12; i < 18; ++i) Jj[i] = 0;
    }
}

// BaseBinaryEdge::constructQuadraticForm (Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-117) for one reprojection edge with
// information w0 * I: vertex 0 = the point (Ji, dim x 3), vertex 1 = the pose (Jj, dim x 6), rows beyond dim zero.  The robust
// branch scales the information and the right-hand side by rho'(chi2) (core/base_edge.h:96-102: the second-order term is commented
// out in the reference).  Outputs are this edge's CONTRIBUTIONS: hl = Ji^T w Ji, gl = -Ji^T w e, hp = Jj^T w Jj, gp = -Jj^T w e,
// w63 = Jj^T w Ji (the pose x point block; g2o stores its transpose).  hp == nullptr: the pose is fixed (hp / gp / w63 untouched).
inline void edge_quadratic_form(int dim, const double* Ji, const double* Jj, const double* er, double w0, const Huber& hub, bool robust,
                                double* hl, double* gl, double* hp, double* gp, double* w63) {
    double r0, r1;
    const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
    robustify(hub, robust, c, &r0, &r1);
    const double w = r1 * w0;
    const double wr[3] = {-w0 * er[0] * r1, -w0 * er[1] * r1, -w0 * er[2] * r1};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < dim; ++k) s += Ji[k * 3 + i] * w * Ji[k * 3 + j];
            hl[i * 3 + j] = s;
        }
        double s = 0;
        for (int k = 0; k < dim; ++k) s += Ji[k * 3 + i] * wr[k];
        gl[i] = s;
    }
    if (!hp) return;
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int k = 0; k < dim; ++k) s += Jj[k * 6 + i] * w * Jj[k * 6 + j];
            hp[i * 6 + j] = s;
        }
        double s = 0;
        for (int k = 0; k < dim; ++k) s += Jj[k * 6 + i] * wr[k];
        gp[i] = s;
        for (int j = 0; j < 3; ++j) {
            double t = 0;
            for (int k = 0; k < dim; ++k) t += Jj[k * 6 + i] * w * Ji[k * 3 + j];
            w63[i * 3 + j] = t;
        }
    }
}

// one block of BaseMultiEdge::computeQuadraticForm (Thirdparty/g2o/g2o/core/base_multi_edge.hpp:170-222): A^T (w I) B for Jacobians given
// row-major dim x da / dim x db, and A^T wr for the right-hand side
inline double block_entry(int da, const double* A, int i, int db, const double* B, int j, int dim, double w) {
    double s = 0;
    for (int k = 0; k < dim; ++k) s += A[k * da + i] * w * B[k * db + j];
    return s;
}
inline double rhs_entry(int da, const double* A, int i, int dim, const double* wr) {
    double s = 0;
    for (int k = 0; k < dim; ++k) s += A[k * da + i] * wr[k];
    return s;
}

inline bool inv3(const double* A, double* B) {   // Eigen fixed-size inverse (cofactors)
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    B[0] = c00 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = c01 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = c02 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return std::isfinite(id);
}

// In-place dense Cholesky A = L L^T (lower), returns false if not positive definite.
bool cholesky(std::vector<double>& A, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0) || !std::isfinite(d)) return false;
        d = std::sqrt(d);
        A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            const double* ri = &A[(size_t)i * n];
            const double* rj = &A[(size_t)j * n];
            for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    return true;
}
void chol_solve(const std::vector<double>& L, int n, double* x) {
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * x[k];
        x[i] = s / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * x[k];
        x[i] = s / L[(size_t)i * n + i];
    }
}

struct Solver {
    adb_ba_problem& P;
    const adb_ba_options& O;
    // state
    std::vector<double> pq, pt, X, J, D, mq, mt;
    // per-edge bookkeeping
    std::vector<uint8_t> lvl_e, lvl_j, lvl_r, lvl_m;       // level (0 active, 1 excluded)
    std::vector<double> chi_e, chi_j, chi_r, chi_m;        // chi2 of the last evaluation that touched the edge
    bool robust = true;
    // dense layout of the current round
    std::vector<int> off_pose, off_dist, off_motion, off_joint;   // dense offset or -1
    std::vector<uint8_t> act_point;
    int n_dense = 0;
    // system
    std::vector<double> H, b, Hll, bl, W;   // W: per static edge 6x3 block (pose x point)
    std::vector<double> x_d, x_l;
    double lambda = 0, ni = 2;
    int trace_len = 0;

    Solver(adb_ba_problem& p, const adb_ba_options& o) : P(p), O(o) {
        pq.assign(p.pose_q, p.pose_q + 4 * p.n_poses); pt.assign(p.pose_t, p.pose_t + 3 * p.n_poses);
        X.assign(p.points, p.points + 3 * p.n_points);
        if (p.n_joints) J.assign(p.joints, p.joints + 3 * p.n_joints);
        if (p.n_dists) D.assign(p.dists, p.dists + p.n_dists);
        if (p.n_motions) { mq.assign(p.motion_q, p.motion_q + 4 * p.n_motions); mt.assign(p.motion_t, p.motion_t + 3 * p.n_motions); }
        lvl_e.assign(p.n_edges, 0); lvl_j.assign(p.n_joint_edges, 0); lvl_r.assign(p.n_rigid_edges, 0); lvl_m.assign(p.n_motion_edges, 0);
        chi_e.assign(p.n_edges, 0); chi_j.assign(p.n_joint_edges, 0); chi_r.assign(p.n_rigid_edges, 0); chi_m.assign(p.n_motion_edges, 0);
    }

    Huber huber(double d) const { return make_huber(d); }

    // SparseOptimizer::initializeOptimization(level 0): active vertices = non-fixed vertices with an active edge
    void build_layout() {
        std::vector<uint8_t> ap(P.n_poses, 0), aj(P.n_joints, 0), ad(P.n_dists, 0), am(P.n_motions, 0);
        act_point.assign(P.n_points, 0);
        for (int e = 0; e < P.n_edges; ++e) if (!lvl_e[e]) { ap[P.edge_pose[e]] = 1; act_point[P.edge_point[e]] = 1; }
        for (int e = 0; e < P.n_joint_edges; ++e) if (!lvl_j[e]) { ap[P.jedge_pose[e]] = 1; aj[P.jedge_joint[e]] = 1; }
        for (int e = 0; e < P.n_rigid_edges; ++e) if (!lvl_r[e]) { aj[P.redge_i[e]] = 1; aj[P.redge_j[e]] = 1; ad[P.redge_dist[e]] = 1; }
        for (int e = 0; e < P.n_motion_edges; ++e) if (!lvl_m[e]) { aj[P.medge_p1[e]] = 1; aj[P.medge_p2[e]] = 1; am[P.medge_motion[e]] = 1; }
        int o = 0;
        off_pose.assign(P.n_poses, -1); off_dist.assign(P.n_dists, -1); off_motion.assign(P.n_motions, -1); off_joint.assign(P.n_joints, -1);
        // g2o orders non-marginalised vertices by id: key-frames, bone lengths, motions, joints (src/Optimizer.cc:1740,1760,1786)
        for (int i = 0; i < P.n_poses; ++i) if (ap[i] && !P.pose_fixed[i]) { off_pose[i] = o; o += 6; }
        for (int i = 0; i < P.n_dists; ++i) if (ad[i]) { off_dist[i] = o; o += 1; }
        for (int i = 0; i < P.n_motions; ++i) if (am[i]) { off_motion[i] = o; o += 6; }
        for (int i = 0; i < P.n_joints; ++i) if (aj[i]) { off_joint[i] = o; o += 3; }
        n_dense = o;
    }

    // dense accumulation helpers: H(oa.., ob..) += A^T (w) B for blocks given as row-major dim x da / dim x db
    void add_block(int oa, int da, const double* A, int ob, int db, const double* B, int dim, double w) {
        if (oa < 0 || ob < 0) return;
        for (int i = 0; i < da; ++i)
            for (int j = 0; j < db; ++j) H[(size_t)(oa + i) * n_dense + ob + j] += block_entry(da, A, i, db, B, j, dim, w);
    }
    void add_rhs(int oa, int da, const double* A, int dim, const double* wr) {   // b += A^T * wr
        if (oa < 0) return;
        for (int i = 0; i < da; ++i) b[oa + i] += rhs_entry(da, A, i, dim, wr);
    }

    // computeActiveErrors + activeRobustChi2 on an arbitrary state (the current trial state is *this)
    double evaluate() {
        double chi = 0, r0, r1;
        std::vector<double> R((size_t)9 * P.n_poses);
        for (int i = 0; i < P.n_poses; ++i) quat_to_rot(&pq[4 * i], &R[9 * i]);
        for (int e = 0; e < P.n_edges; ++e) {
            if (lvl_e[e]) continue;
            double er[3], Xc[3];
            const int dim = reproj_error(P, &R[9 * P.edge_pose[e]], &pt[3 * P.edge_pose[e]], &X[3 * P.edge_point[e]], &P.edge_obs[3 * e], er, Xc);
            const double w = P.edge_info[e];
            const double c = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
            chi_e[e] = c;
            robustify(huber(dim == 3 ? O.huber_stereo : O.huber_mono), robust, c, &r0, &r1);
            chi += r0;
        }
        for (int e = 0; e < P.n_joint_edges; ++e) {
            if (lvl_j[e]) continue;
            double er[3], Xc[3];
            const int dim = reproj_error(P, &R[9 * P.jedge_pose[e]], &pt[3 * P.jedge_pose[e]], &J[3 * P.jedge_joint[e]], &P.jedge_obs[3 * e], er, Xc);
            const double w = P.jedge_info[e];
            const double c = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
            chi_j[e] = c;
            robustify(huber(dim == 3 ? O.huber_stereo : O.huber_mono), robust, c, &r0, &r1);
            chi += r0;
        }
        for (int e = 0; e < P.n_rigid_edges; ++e) {
            if (lvl_r[e]) continue;
            double d3[3], nrm;
            const double er = rigid_error(&J[3 * P.redge_i[e]], &J[3 * P.redge_j[e]], D[P.redge_dist[e]], d3, &nrm);
            const double c = er * (P.redge_info[e] * er);
            chi_r[e] = c;
            robustify(huber(O.huber_rigid), robust, c, &r0, &r1);
            chi += r0;
        }
        for (int e = 0; e < P.n_motion_edges; ++e) {
            if (lvl_m[e]) continue;
            double er[3], Rm[9];
            motion_error(e, er, Rm);
            const double w = P.medge_info[e];
            const double c = er[0] * (w * er[0]) + er[1] * (w * er[1]) + er[2] * (w * er[2]);
            chi_m[e] = c;
            robustify(huber(O.huber_motion), robust, c, &r0, &r1);
            chi += r0;
        }
        return chi;
    }

    // e = p1 - M^-1 p2, M = (R, dt * t)   (include/g2o_dyn_slam3d.h:65-76)
    void motion_error(int e, double* er, double* Rm) const {
        const int m = P.medge_motion[e];
        motion_edge_error(&mq[4 * m], &mt[3 * m], P.medge_dt[e], &J[3 * P.medge_p1[e]], &J[3 * P.medge_p2[e]], er, Rm);
    }

    // BlockSolver::buildSystem at the current state (errors of the current state are recomputed inside)
    void build_system() {
        H.assign((size_t)n_dense * n_dense, 0.0); b.assign(n_dense, 0.0);
        Hll.assign((size_t)9 * P.n_points, 0.0); bl.assign((size_t)3 * P.n_points, 0.0);
        W.assign((size_t)18 * P.n_edges, 0.0);
        std::vector<double> R((size_t)9 * P.n_poses);
        for (int i = 0; i < P.n_poses; ++i) quat_to_rot(&pq[4 * i], &R[9 * i]);
        double r0, r1;
        for (int e = 0; e < P.n_edges; ++e) {
            if (lvl_e[e]) continue;
            const int ip = P.edge_pose[e], il = P.edge_point[e];
            double er[3], Xc[3], Ji[9], Jj[18];
            const int dim = reproj_error(P, &R[9 * ip], &pt[3 * ip], &X[3 * il], &P.edge_obs[3 * e], er, Xc);
            reproj_jacobians(P, &R[9 * ip], Xc, dim, Ji, Jj);
            const int op = off_pose[ip];
            double hl[9], gl[3], hp[36], gp[6];
            edge_quadratic_form(dim, Ji, Jj, er, P.edge_info[e], huber(dim == 3 ? O.huber_stereo : O.huber_mono), robust, hl, gl,
                                op >= 0 ? hp : nullptr, gp, &W[(size_t)18 * e]);
            for (int i = 0; i < 9; ++i) Hll[(size_t)9 * il + i] += hl[i];
            for (int i = 0; i < 3; ++i) bl[(size_t)3 * il + i] += gl[i];
            if (op >= 0) {
                for (int i = 0; i < 6; ++i) {
                    for (int j = 0; j < 6; ++j) H[(size_t)(op + i) * n_dense + op + j] += hp[i * 6 + j];
                    b[op + i] += gp[i];
                }
            }
        }
        for (int e = 0; e < P.n_joint_edges; ++e) {
            if (lvl_j[e]) continue;
            const int ip = P.jedge_pose[e], ij = P.jedge_joint[e];
            double er[3], Xc[3], Ji[9], Jj[18];
            const int dim = reproj_error(P, &R[9 * ip], &pt[3 * ip], &J[3 * ij], &P.jedge_obs[3 * e], er, Xc);
            reproj_jacobians(P, &R[9 * ip], Xc, dim, Ji, Jj);
            const double w0 = P.jedge_info[e];
            const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
            robustify(huber(dim == 3 ? O.huber_stereo : O.huber_mono), robust, c, &r0, &r1);
            const double w = r1 * w0;
            const double wr[3] = {-w0 * er[0] * r1, -w0 * er[1] * r1, -w0 * er[2] * r1};
            const int op = off_pose[ip], oj = off_joint[ij];
            add_block(oj, 3, Ji, oj, 3, Ji, dim, w); add_rhs(oj, 3, Ji, dim, wr);
            add_block(op, 6, Jj, op, 6, Jj, dim, w); add_rhs(op, 6, Jj, dim, wr);
            add_block(op, 6, Jj, oj, 3, Ji, dim, w); add_block(oj, 3, Ji, op, 6, Jj, dim, w);
        }
        for (int e = 0; e < P.n_rigid_edges; ++e) {
            if (lvl_r[e]) continue;
            const int i1 = P.redge_i[e], i2 = P.redge_j[e], id = P.redge_dist[e];
            double d[3], n;
            const double er = rigid_error(&J[3 * i1], &J[3 * i2], D[id], d, &n);
            const double w0 = P.redge_info[e];
            const double c = er * (w0 * er);
            robustify(huber(O.huber_rigid), robust, c, &r0, &r1);
            const double w = r1 * w0;
            const double wr[1] = {-w0 * er * r1};
            double Ja[3] = {0, 0, 0}, Jb[3] = {0, 0, 0};   // D.4: analytic Jacobian; zero rows for coincident joints
            if (n >= 1e-12) for (int k = 0; k < 3; ++k) { Ja[k] = d[k] / n; Jb[k] = -d[k] / n; }
            const double Jd[1] = {-1.0};
            const int o1 = off_joint[i1], o2 = off_joint[i2], od = off_dist[id];
            const int offs[3] = {o1, o2, od}; const int dims[3] = {3, 3, 1}; const double* Js[3] = {Ja, Jb, Jd};
            for (int u = 0; u < 3; ++u) {
                add_rhs(offs[u], dims[u], Js[u], 1, wr);
                for (int v = 0; v < 3; ++v) add_block(offs[u], dims[u], Js[u], offs[v], dims[v], Js[v], 1, w);
            }
        }
        for (int e = 0; e < P.n_motion_edges; ++e) {
            if (lvl_m[e]) continue;
            double er[3], Rm[9];
            motion_error(e, er, Rm);
            const double w0 = P.medge_info[e];
            const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
            robustify(huber(O.huber_motion), robust, c, &r0, &r1);
            const double w = r1 * w0;
            const double wr[3] = {-w0 * er[0] * r1, -w0 * er[1] * r1, -w0 * er[2] * r1};
            double J1[9], J2[9], Jm[18];
            motion_edge_jacobians(Rm, P.medge_dt[e], J1, J2, Jm);                                  // I, -R^T, (dt I | 0): D.6
            const int offs[3] = {off_joint[P.medge_p1[e]], off_joint[P.medge_p2[e]], off_motion[P.medge_motion[e]]};
            const int dims[3] = {3, 3, 6}; const double* Js[3] = {J1, J2, Jm};
            for (int u = 0; u < 3; ++u) {
                add_rhs(offs[u], dims[u], Js[u], 3, wr);
                for (int v = 0; v < 3; ++v) add_block(offs[u], dims[u], Js[u], offs[v], dims[v], Js[v], 3, w);
            }
        }
    }

    double max_diag() const {
        double m = 0;
        for (int i = 0; i < n_dense; ++i) m = std::max(m, std::fabs(H[(size_t)i * n_dense + i]));
        for (int l = 0; l < P.n_points; ++l)
            if (act_point[l]) for (int i = 0; i < 3; ++i) m = std::max(m, std::fabs(Hll[(size_t)9 * l + 4 * i]));
        return m;
    }

    // BlockSolver::solve with lambda on every diagonal; returns false if the reduced system is not positive definite
    bool solve_trial() {
        std::vector<double> S(H), bs(b);
        for (int i = 0; i < n_dense; ++i) S[(size_t)i * n_dense + i] += lambda;
        std::vector<double> Dinv((size_t)9 * P.n_points, 0.0);
        // per point: edges (active, free pose) grouped
        std::vector<std::vector<int>> pe(P.n_points);
        for (int e = 0; e < P.n_edges; ++e) if (!lvl_e[e] && off_pose[P.edge_pose[e]] >= 0) pe[P.edge_point[e]].push_back(e);
        for (int l = 0; l < P.n_points; ++l) {
            if (!act_point[l]) continue;
            double Dl[9];
            std::memcpy(Dl, &Hll[(size_t)9 * l], sizeof(Dl));
            Dl[0] += lambda; Dl[4] += lambda; Dl[8] += lambda;
            inv3(Dl, &Dinv[(size_t)9 * l]);
            const double* Di = &Dinv[(size_t)9 * l];
            double db[3];
            for (int i = 0; i < 3; ++i) db[i] = Di[i * 3] * bl[3 * l] + Di[i * 3 + 1] * bl[3 * l + 1] + Di[i * 3 + 2] * bl[3 * l + 2];
            for (int e1 : pe[l]) {
                const double* B1 = &W[(size_t)18 * e1];
                const int o1 = off_pose[P.edge_pose[e1]];
                double BD[18];
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 3; ++j) BD[i * 3 + j] = B1[i * 3] * Di[j] + B1[i * 3 + 1] * Di[3 + j] + B1[i * 3 + 2] * Di[6 + j];
                for (int i = 0; i < 6; ++i) bs[o1 + i] -= B1[i * 3] * db[0] + B1[i * 3 + 1] * db[1] + B1[i * 3 + 2] * db[2];
                for (int e2 : pe[l]) {
                    const double* B2 = &W[(size_t)18 * e2];
                    const int o2 = off_pose[P.edge_pose[e2]];
                    for (int i = 0; i < 6; ++i)
                        for (int j = 0; j < 6; ++j)
                            S[(size_t)(o1 + i) * n_dense + o2 + j] -= BD[i * 3] * B2[j * 3] + BD[i * 3 + 1] * B2[j * 3 + 1] + BD[i * 3 + 2] * B2[j * 3 + 2];
                }
            }
        }
        x_d.assign(bs.begin(), bs.end());
        x_l.assign((size_t)3 * P.n_points, 0.0);
        if (n_dense > 0) {
            if (!cholesky(S, n_dense)) return false;
            chol_solve(S, n_dense, x_d.data());
        }
        // xl = Dinv (bl - B^T xp)
        for (int l = 0; l < P.n_points; ++l) {
            if (!act_point[l]) continue;
            double c[3] = {bl[3 * l], bl[3 * l + 1], bl[3 * l + 2]};
            for (int e : pe[l]) {
                const double* B = &W[(size_t)18 * e];
                const int o = off_pose[P.edge_pose[e]];
                for (int j = 0; j < 3; ++j)
                    for (int i = 0; i < 6; ++i) c[j] -= B[i * 3 + j] * x_d[o + i];
            }
            const double* Di = &Dinv[(size_t)9 * l];
            for (int i = 0; i < 3; ++i) x_l[3 * l + i] = Di[i * 3] * c[0] + Di[i * 3 + 1] * c[1] + Di[i * 3 + 2] * c[2];
        }
        return true;
    }

    void apply_update() {
        for (int i = 0; i < P.n_poses; ++i) if (off_pose[i] >= 0) pose_oplus(&pq[4 * i], &pt[3 * i], &x_d[off_pose[i]]);
        for (int i = 0; i < P.n_dists; ++i) if (off_dist[i] >= 0) D[i] += x_d[off_dist[i]];
        for (int i = 0; i < P.n_motions; ++i) if (off_motion[i] >= 0) motion_oplus(&mq[4 * i], &mt[3 * i], &x_d[off_motion[i]]);
        for (int i = 0; i < P.n_joints; ++i) if (off_joint[i] >= 0) for (int k = 0; k < 3; ++k) J[3 * i + k] += x_d[off_joint[i] + k];
        for (int l = 0; l < P.n_points; ++l) if (act_point[l]) for (int k = 0; k < 3; ++k) X[3 * l + k] += x_l[3 * l + k];
    }

    double scale_term() const {   // sum x (lambda x + b) over the whole solution vector
        double s = 0;
        for (int i = 0; i < n_dense; ++i) s += x_d[i] * (lambda * x_d[i] + b[i]);
        for (int l = 0; l < P.n_points; ++l)
            if (act_point[l]) for (int k = 0; k < 3; ++k) s += x_l[3 * l + k] * (lambda * x_l[3 * l + k] + bl[3 * l + k]);
        return s;
    }

    // SparseOptimizer::optimize(iterations); returns the number of iterations run
    int optimize(int iterations, volatile const uint8_t* stop, adb_ba_result* res, double* chi_out) {
        int nbad = 0, it_run = 0;
        double current = 0;
        for (int it = 0; it < iterations && !(stop && *stop); ++it) {
            current = evaluate();
            if (it == 0 && res && res->iterations_run[0] == 0 && res->iterations_run[1] == 0 && res->trials_run == 0) res->chi2_initial = current;
            const double ini = current;
            double temp = current;
            build_system();
            if (it == 0) { lambda = O.tau * max_diag(); ni = 2; nbad = 0; }
            double rho = 0;
            int q = 0;
            do {
                std::vector<double> bq(pq), bt(pt), bX(X), bJ(J), bD(D), bmq(mq), bmt(mt);   // push
                const bool ok = solve_trial();
                if (ok) apply_update();
                temp = evaluate();
                if (!ok) temp = std::numeric_limits<double>::max();
                rho = (current - temp);
                double scale = ok ? scale_term() : 0.0;
                scale += 1e-3;
                rho /= scale;
                const double lam_used = lambda;
                const bool good = rho > 0 && std::isfinite(temp);
                if (res && res->trace && trace_len < res->trace_cap) {
                    double* tr = res->trace + (size_t)ADB_BA_TRACE_COLS * trace_len++;
                    tr[0] = lam_used; tr[1] = current; tr[2] = temp; tr[3] = rho; tr[4] = good ? 1 : 0;
                }
                if (res) res->trials_run++;
                if (good) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    lambda *= std::max(1. / 3., alpha);
                    ni = 2;
                    current = temp;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    pq = bq; pt = bt; X = bX; J = bJ; D = bD; mq = bmq; mt = bmt;   // pop (errors stay stale, like the reference)
                }
                ++q;
            } while (rho < 0 && q < O.max_trials && !(stop && *stop));
            ++it_run;
            if (q == O.max_trials || rho == 0) break;
            if ((ini - current) * 1e3 < ini) ++nbad; else nbad = 0;
            if (nbad >= 3) break;
        }
        *chi_out = current;
        return it_run;
    }
};

inline bool depth_positive(const double* q, const double* t, const double* X) {
    double R[9];
    quat_to_rot(q, R);
    return R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2] > 0;
}

}  // namespace

extern "C" {

void ba_oracle_default_options(adb_ba_options* o) {
    o->iterations[0] = 5; o->iterations[1] = 10; o->max_trials = 10; o->tau = 1e-5;
    o->chi2_mono = 5.991; o->chi2_stereo = 7.815; o->chi2_rigid = 1.0; o->chi2_motion = 4.0;
    o->huber_mono = (double)(float)std::sqrt(5.991); o->huber_stereo = (double)(float)std::sqrt(7.815);
    o->huber_rigid = 1.0; o->huber_motion = (double)(float)std::sqrt(4.0);
    o->robust[0] = 1; o->robust[1] = 0;
}

// Optimizer::BundleAdjustment (src/Optimizer.cc:52-230): one round, thHuber2D = sqrt(5.99), robust kernel by bRobust.
void ba_oracle_global_options(adb_ba_options* o, int32_t n_iterations, int32_t robust) {
    ba_oracle_default_options(o);
    o->iterations[0] = n_iterations; o->iterations[1] = 0;
    o->huber_mono = (double)(float)std::sqrt(5.99);
    o->robust[0] = robust ? 1 : 0;
}

// Same contract as adb_ba_solve (include/airdos_b200.h).
int ba_oracle_solve(adb_ba_problem* prob, const adb_ba_options* opt, volatile const uint8_t* stop, adb_ba_result* res) {
    if (stop && *stop) return ADB_ERR_STOPPED;
    {   // same range checks as adb_ba_solve: a bad index is ADB_ERR_INVALID, never a wild write
        auto in = [](int v, int n) { return v >= 0 && v < n; };
        for (int e = 0; e < prob->n_edges; ++e) if (!in(prob->edge_pose[e], prob->n_poses) || !in(prob->edge_point[e], prob->n_points)) return ADB_ERR_INVALID;
        for (int e = 0; e < prob->n_joint_edges; ++e) if (!in(prob->jedge_pose[e], prob->n_poses) || !in(prob->jedge_joint[e], prob->n_joints)) return ADB_ERR_INVALID;
        for (int e = 0; e < prob->n_rigid_edges; ++e)
            if (!in(prob->redge_i[e], prob->n_joints) || !in(prob->redge_j[e], prob->n_joints) || !in(prob->redge_dist[e], prob->n_dists)) return ADB_ERR_INVALID;
        for (int e = 0; e < prob->n_motion_edges; ++e)
            if (!in(prob->medge_p1[e], prob->n_joints) || !in(prob->medge_p2[e], prob->n_joints) || !in(prob->medge_motion[e], prob->n_motions)) return ADB_ERR_INVALID;
    }
    Solver S(*prob, *opt);
    res->iterations_run[0] = res->iterations_run[1] = 0; res->trials_run = 0; res->stopped = 0; res->trace_len = 0;
    res->chi2_initial = 0; res->chi2_round[0] = res->chi2_round[1] = 0;
    S.robust = opt->robust[0] != 0;
    S.build_layout();
    double chi = 0;
    res->iterations_run[0] = S.optimize(opt->iterations[0], stop, res, &chi);
    res->chi2_round[0] = chi;
    const bool more = !(stop && *stop) && opt->iterations[1] > 0;
    if (stop && *stop) res->stopped = 1;
    if (more) {
        // chi2 gates with the errors of the last evaluated state; depth test on the current estimates
        for (int e = 0; e < prob->n_edges; ++e) {
            const bool stereo = prob->edge_obs[3 * e + 2] >= 0;
            const int ip = prob->edge_pose[e];
            if (S.chi_e[e] > (stereo ? opt->chi2_stereo : opt->chi2_mono) || !depth_positive(&S.pq[4 * ip], &S.pt[3 * ip], &S.X[3 * prob->edge_point[e]]))
                S.lvl_e[e] = 1;
        }
        for (int e = 0; e < prob->n_joint_edges; ++e) {
            const int ip = prob->jedge_pose[e];
            if (S.chi_j[e] > opt->chi2_stereo || !depth_positive(&S.pq[4 * ip], &S.pt[3 * ip], &S.J[3 * prob->jedge_joint[e]])) S.lvl_j[e] = 1;
        }
        for (int e = 0; e < prob->n_rigid_edges; ++e) if (S.chi_r[e] > opt->chi2_rigid) S.lvl_r[e] = 1;
        for (int e = 0; e < prob->n_motion_edges; ++e) if (S.chi_m[e] > opt->chi2_motion) S.lvl_m[e] = 1;
        S.robust = opt->robust[1] != 0;
        S.build_layout();
        res->iterations_run[1] = S.optimize(opt->iterations[1], stop, res, &chi);
        res->chi2_round[1] = chi;
        if (stop && *stop) res->stopped = 1;
    }
    res->lambda_final = S.lambda;
    res->trace_len = S.trace_len;
    for (int e = 0; e < prob->n_edges; ++e) {
        const bool stereo = prob->edge_obs[3 * e + 2] >= 0;
        const int ip = prob->edge_pose[e];
        const bool out = S.chi_e[e] > (stereo ? opt->chi2_stereo : opt->chi2_mono) || !depth_positive(&S.pq[4 * ip], &S.pt[3 * ip], &S.X[3 * prob->edge_point[e]]);
        if (res->edge_outlier) res->edge_outlier[e] = out;
        if (res->edge_chi2) res->edge_chi2[e] = S.chi_e[e];
    }
    for (int e = 0; e < prob->n_joint_edges; ++e) {
        const int ip = prob->jedge_pose[e];
        if (res->jedge_outlier)
            res->jedge_outlier[e] = S.chi_j[e] > opt->chi2_stereo || !depth_positive(&S.pq[4 * ip], &S.pt[3 * ip], &S.J[3 * prob->jedge_joint[e]]);
    }
    for (int e = 0; e < prob->n_rigid_edges; ++e) if (res->redge_outlier) res->redge_outlier[e] = S.chi_r[e] > opt->chi2_rigid;
    for (int e = 0; e < prob->n_motion_edges; ++e) if (res->medge_outlier) res->medge_outlier[e] = S.chi_m[e] > opt->chi2_motion;
    std::memcpy(prob->pose_q, S.pq.data(), S.pq.size() * 8); std::memcpy(prob->pose_t, S.pt.data(), S.pt.size() * 8);
    std::memcpy(prob->points, S.X.data(), S.X.size() * 8);
    if (prob->n_joints) std::memcpy(prob->joints, S.J.data(), S.J.size() * 8);
    if (prob->n_dists) std::memcpy(prob->dists, S.D.data(), S.D.size() * 8);
    if (prob->n_motions) { std::memcpy(prob->motion_q, S.mq.data(), S.mq.size() * 8); std::memcpy(prob->motion_t, S.mt.data(), S.mt.size() * 8); }
    return ADB_OK;
}

// ---- pieces exposed for the pinning tests ----
// residual + Jacobians of one reprojection edge; returns dim
int ba_oracle_reproj(const adb_ba_problem* P, const double* q, const double* t, const double* X, const double* obs, double* e, double* Ji, double* Jj) {
    double R[9], Xc[3];
    quat_to_rot(q, R);
    const int dim = reproj_error(*P, R, t, X, obs, e, Xc);
    reproj_jacobians(*P, R, Xc, dim, Ji, Jj);
    return dim;
}
double ba_oracle_rigid_error(const double* a, const double* c2, double dist) { double d3[3], n; return rigid_error(a, c2, dist, d3, &n); }
void ba_oracle_motion_edge(const double* p1, const double* p2, const double* mq, const double* mt, double dt, double* er, double* J1, double* J2, double* Jm) {
    double Rm[9];
    motion_edge_error(mq, mt, dt, p1, p2, er, Rm);
    motion_edge_jacobians(Rm, dt, J1, J2, Jm);
}
void ba_oracle_pose_oplus(double* q, double* t, const double* d) { pose_oplus(q, t, d); }
void ba_oracle_motion_oplus(double* q, double* t, const double* d) { motion_oplus(q, t, d); }

// Converter::toSE3Quat: float 4x4 -> double R -> Quaterniond(R) -> normalised (SE3Quat ctor)
void ba_oracle_pose_from_tcw(const float* T, double* q, double* t) {
    double R[9];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = (double)T[i * 4 + j]; t[i] = (double)T[i * 4 + 3]; }
    rot_to_quat(R, q);
    quat_normalize_pos(q);
}
void ba_oracle_pose_to_tcw(const double* q, const double* t, float* T) {
    double R[9];
    quat_to_rot(q, R);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float)R[i * 3 + j]; T[i * 4 + 3] = (float)t[i]; }
    T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}
// one LM trial system at the initial state, for the independent numpy cross-check:
// returns the dense step x_d (n_dense) and x_l (3 * n_points) for a given lambda.
int ba_oracle_first_step(adb_ba_problem* prob, const adb_ba_options* opt, double lambda, int robust, double* x_d, int cap_d, double* x_l) {
    Solver S(*prob, *opt);
    S.robust = robust != 0;
    S.build_layout();
    S.evaluate();
    S.build_system();
    S.lambda = lambda;
    if (!S.solve_trial()) return -1;
    for (int i = 0; i < S.n_dense && i < cap_d; ++i) x_d[i] = S.x_d[i];
    std::memcpy(x_l, S.x_l.data(), S.x_l.size() * 8);
    return S.n_dense;
}

// a multi-vertex edge's quadratic form the way Solver::build_system accumulates rigidity / motion edges: nv vertices of dimensions dims[],
// Jacobians Js[v] row-major dim x dims[v], information w0 * I, Huber(delta) if robust.  H: dense (sum dims)^2 row-major, b: sum dims
void ba_oracle_multi_quadratic_form(int dim, int nv, const int32_t* dims, const double* const* Js, const double* er, double w0, double delta, int robust,
                                    double* H, double* b) {
    double r0, r1, c = 0;
    for (int k = 0; k < dim; ++k) c += er[k] * (w0 * er[k]);
    robustify(make_huber(delta), robust != 0, c, &r0, &r1);
    const double w = r1 * w0;
    double wr[3] = {0, 0, 0};
    for (int k = 0; k < dim; ++k) wr[k] = -w0 * er[k] * r1;
    int n = 0; std::vector<int> off(nv);
    for (int v = 0; v < nv; ++v) { off[v] = n; n += dims[v]; }
    for (int u = 0; u < nv; ++u) {
        for (int i = 0; i < dims[u]; ++i) b[off[u] + i] = rhs_entry(dims[u], Js[u], i, dim, wr);
        for (int v = 0; v < nv; ++v)
            for (int i = 0; i < dims[u]; ++i)
                for (int j = 0; j < dims[v]; ++j) H[(size_t)(off[u] + i) * n + off[v] + j] = block_entry(dims[u], Js[u], i, dims[v], Js[v], j, dim, w);
    }
}
void ba_oracle_huber(double delta, double e2, double* rho2) { robustify(make_huber(delta), true, e2, rho2, rho2 + 1); }
// one reprojection edge's quadratic form (what Solver::build_system adds for it); hp may be NULL for a fixed pose
void ba_oracle_edge_quadratic_form(int dim, const double* Ji, const double* Jj, const double* er, double w0, double delta, int robust,
                                   double* hl, double* gl, double* hp, double* gp, double* w63) {
    edge_quadratic_form(dim, Ji, Jj, er, w0, make_huber(delta), robust != 0, hl, gl, hp, gp, w63);
}

// ---- the solver's steps one by one, for oracle/ref_lm.cpp: the reference's OWN OptimizationAlgorithmLevenberg::solve and
// SparseOptimizer::optimize (compiled from /root/reference) drive these through function pointers, so that the control flow of
// Solver::optimize above can be held against the literal reference on the same arithmetic.
struct LmSession {
    Solver S;
    struct Snap { std::vector<double> pq, pt, X, J, D, mq, mt; };
    std::vector<Snap> stack;
    double chi = 0;
    LmSession(adb_ba_problem& p, const adb_ba_options& o) : S(p, o) {}
};
// the same for all four edge kinds (any pointer may be NULL = all active)
void ba_oracle_lm_set_levels4(void* h, const uint8_t* lvl_e, const uint8_t* lvl_j, const uint8_t* lvl_r, const uint8_t* lvl_m) {
    Solver& S = ((LmSession*)h)->S;
    for (int e = 0; e < S.P.n_edges; ++e) S.lvl_e[e] = lvl_e ? lvl_e[e] : 0;
    for (int e = 0; e < S.P.n_joint_edges; ++e) S.lvl_j[e] = lvl_j ? lvl_j[e] : 0;
    for (int e = 0; e < S.P.n_rigid_edges; ++e) S.lvl_r[e] = lvl_r ? lvl_r[e] : 0;
    for (int e = 0; e < S.P.n_motion_edges; ++e) S.lvl_m[e] = lvl_m ? lvl_m[e] : 0;
    S.build_layout();
}
void* ba_oracle_lm_open(adb_ba_problem* prob, const adb_ba_options* opt, int robust) {
    LmSession* h = new LmSession(*prob, *opt);
    h->S.robust = robust != 0;
    h->S.build_layout();
    return h;
}
void ba_oracle_lm_close(void* h) { delete (LmSession*)h; }
// levels of the static edges (1 = excluded, like setLevel(1) + initializeOptimization(0)); the layout is rebuilt
void ba_oracle_lm_set_levels(void* h, const uint8_t* lvl_e) {
    Solver& S = ((LmSession*)h)->S;
    for (int e = 0; e < S.P.n_edges; ++e) S.lvl_e[e] = lvl_e[e];
    S.build_layout();
}
void ba_oracle_lm_compute_errors(void* h) { LmSession* s = (LmSession*)h; s->chi = s->S.evaluate(); }   // computeActiveErrors
double ba_oracle_lm_chi2(void* h) { return ((LmSession*)h)->chi; }                                       // activeRobustChi2
void ba_oracle_lm_build(void* h) { ((LmSession*)h)->S.build_system(); }                                  // Solver::buildSystem
// indexMapping(): the dimension of every active vertex in solver order (key-frames, bone lengths, motions, joints, map points)
int ba_oracle_lm_layout(void* h, int32_t* dims, int cap) {
    const Solver& S = ((LmSession*)h)->S;
    int n = 0;
    auto put = [&](int d) { if (n < cap) dims[n] = d; ++n; };
    for (int i = 0; i < S.P.n_poses; ++i) if (S.off_pose[i] >= 0) put(6);
    for (int i = 0; i < S.P.n_dists; ++i) if (S.off_dist[i] >= 0) put(1);
    for (int i = 0; i < S.P.n_motions; ++i) if (S.off_motion[i] >= 0) put(6);
    for (int i = 0; i < S.P.n_joints; ++i) if (S.off_joint[i] >= 0) put(3);
    for (int l = 0; l < S.P.n_points; ++l) if (S.act_point[l]) put(3);
    return n;
}
// the whole solution vector x, right-hand side b and Hessian diagonal in that order; returns the vector size
int ba_oracle_lm_vectors(void* h, double* x, double* b, double* diag, int cap) {
    const Solver& S = ((LmSession*)h)->S;
    int n = 0;
    for (int i = 0; i < S.n_dense; ++i, ++n)
        if (n < cap) { if (x) x[n] = S.x_d.empty() ? 0.0 : S.x_d[i]; if (b) b[n] = S.b[i]; if (diag) diag[n] = S.H[(size_t)i * S.n_dense + i]; }
    for (int l = 0; l < S.P.n_points; ++l) {
        if (!S.act_point[l]) continue;
        for (int k = 0; k < 3; ++k, ++n)
            if (n < cap) { if (x) x[n] = S.x_l.empty() ? 0.0 : S.x_l[3 * l + k]; if (b) b[n] = S.bl[3 * l + k]; if (diag) diag[n] = S.Hll[(size_t)9 * l + 4 * k]; }
    }
    return n;
}
// The assembled normal equations of a STATIC window (no articulated vertices) in the block form BlockSolver_6_3 holds them, for the
// pin of the Schur complement against the reference's own BlockSolver::solve (oracle/ref_schur.cpp): free poses and active points
// renumbered densely in solver order; per active edge on a free pose its (pose, point) pair and 6 x 3 block W (row major); the diagonal
// 6 x 6 pose blocks, the 3 x 3 point blocks, b = [poses | points].  Any output pointer may be NULL; sizes[3] = {free poses, active
// points, edges with a free pose}.  Returns 0 when the window has articulated vertices.
int ba_oracle_lm_system(void* h, int32_t* sizes, int32_t* edge_pose, int32_t* edge_point, double* W, double* Hpp, double* Hll, double* b) {
    const Solver& S = ((LmSession*)h)->S;
    if (S.P.n_joints || S.P.n_dists || S.P.n_motions) return 0;
    std::vector<int> pt_idx(S.P.n_points, -1);
    int np = 0, nl = 0, ne = 0;
    for (int i = 0; i < S.P.n_poses; ++i) if (S.off_pose[i] >= 0) ++np;
    for (int l = 0; l < S.P.n_points; ++l) if (S.act_point[l]) pt_idx[l] = nl++;
    for (int e = 0; e < S.P.n_edges; ++e) {
        if (S.lvl_e[e] || S.off_pose[S.P.edge_pose[e]] < 0) continue;
        if (edge_pose) edge_pose[ne] = S.off_pose[S.P.edge_pose[e]] / 6;
        if (edge_point) edge_point[ne] = pt_idx[S.P.edge_point[e]];
        if (W) std::memcpy(W + (size_t)18 * ne, &S.W[(size_t)18 * e], 18 * sizeof(double));
        ++ne;
    }
    if (Hpp)
        for (int i = 0; i < np; ++i)
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) Hpp[(size_t)36 * i + 6 * r + c] = S.H[(size_t)(6 * i + r) * S.n_dense + 6 * i + c];
    for (int l = 0; l < S.P.n_points; ++l) {
        if (!S.act_point[l]) continue;
        if (Hll) std::memcpy(Hll + (size_t)9 * pt_idx[l], &S.Hll[(size_t)9 * l], 9 * sizeof(double));
        if (b) for (int k = 0; k < 3; ++k) b[S.n_dense + 3 * pt_idx[l] + k] = S.bl[3 * l + k];
    }
    if (b) for (int i = 0; i < S.n_dense; ++i) b[i] = S.b[i];
    if (sizes) { sizes[0] = np; sizes[1] = nl; sizes[2] = ne; }
    return 1;
}
// The same for ANY window, articulated ones included, in the form BlockSolverX holds it: the non-marginalised vertices (key-frames 6,
// bone lengths 1, motions 6, joints 3, in solver order) with the whole dense block H over them (n_dense^2, symmetric), the map points
// as the marginalised landmarks.  sizes[4] = {non-marginalised vertices, active points, edges with a free pose, n_dense}.
int ba_oracle_lm_system_x(void* h, int32_t* sizes, int32_t* dims, int32_t* edge_block, int32_t* edge_point, double* W, double* H, double* Hll,
                          double* b) {
    const Solver& S = ((LmSession*)h)->S;
    std::vector<int> pt_idx(S.P.n_points, -1);
    int nb = 0, nl = 0, ne = 0;
    auto put = [&](int d) { if (dims) dims[nb] = d; ++nb; };
    for (int i = 0; i < S.P.n_poses; ++i) if (S.off_pose[i] >= 0) put(6);
    for (int i = 0; i < S.P.n_dists; ++i) if (S.off_dist[i] >= 0) put(1);
    for (int i = 0; i < S.P.n_motions; ++i) if (S.off_motion[i] >= 0) put(6);
    for (int i = 0; i < S.P.n_joints; ++i) if (S.off_joint[i] >= 0) put(3);
    for (int l = 0; l < S.P.n_points; ++l) if (S.act_point[l]) pt_idx[l] = nl++;
    for (int e = 0; e < S.P.n_edges; ++e) {
        if (S.lvl_e[e] || S.off_pose[S.P.edge_pose[e]] < 0) continue;
        if (edge_block) edge_block[ne] = S.off_pose[S.P.edge_pose[e]] / 6;      // the key-frames come first and are all 6 wide
        if (edge_point) edge_point[ne] = pt_idx[S.P.edge_point[e]];
        if (W) std::memcpy(W + (size_t)18 * ne, &S.W[(size_t)18 * e], 18 * sizeof(double));
        ++ne;
    }
    if (H) std::memcpy(H, S.H.data(), (size_t)S.n_dense * S.n_dense * sizeof(double));
    for (int l = 0; l < S.P.n_points; ++l) {
        if (!S.act_point[l]) continue;
        if (Hll) std::memcpy(Hll + (size_t)9 * pt_idx[l], &S.Hll[(size_t)9 * l], 9 * sizeof(double));
        if (b) for (int k = 0; k < 3; ++k) b[S.n_dense + 3 * pt_idx[l] + k] = S.bl[3 * l + k];
    }
    if (b) for (int i = 0; i < S.n_dense; ++i) b[i] = S.b[i];
    if (sizes) { sizes[0] = nb; sizes[1] = nl; sizes[2] = ne; sizes[3] = S.n_dense; }
    return 1;
}
void ba_oracle_lm_set_lambda(void* h, double lambda) { ((LmSession*)h)->S.lambda = lambda; }             // Solver::setLambda
int ba_oracle_lm_solve(void* h) { return ((LmSession*)h)->S.solve_trial() ? 1 : 0; }                     // Solver::solve
void ba_oracle_lm_update(void* h) { ((LmSession*)h)->S.apply_update(); }                                 // SparseOptimizer::update(x)
void ba_oracle_lm_push(void* h) {
    LmSession* s = (LmSession*)h;
    s->stack.push_back(LmSession::Snap{s->S.pq, s->S.pt, s->S.X, s->S.J, s->S.D, s->S.mq, s->S.mt});
}
void ba_oracle_lm_pop(void* h) {
    LmSession* s = (LmSession*)h;
    LmSession::Snap& t = s->stack.back();
    s->S.pq = t.pq; s->S.pt = t.pt; s->S.X = t.X; s->S.J = t.J; s->S.D = t.D; s->S.mq = t.mq; s->S.mt = t.mt;
    s->stack.pop_back();
}
void ba_oracle_lm_discard_top(void* h) { ((LmSession*)h)->stack.pop_back(); }
// the oracle's own loop on the session (Solver::optimize): trace rows of ADB_BA_TRACE_COLS; returns iterations run
int ba_oracle_lm_optimize(void* h, int iterations, double* trace, int trace_cap, int* trace_len, double* lambda_final) {
    LmSession* s = (LmSession*)h;
    adb_ba_result res{};
    res.trace = trace; res.trace_cap = trace_cap;
    double chi = 0;
    const int it = s->S.optimize(iterations, nullptr, &res, &chi);
    *trace_len = s->S.trace_len; *lambda_final = s->S.lambda;
    return it;
}
// current estimates, concatenated (poses q | t, points, joints, bone lengths, motions q | t); returns the length
int ba_oracle_lm_state(void* h, double* out, int cap) {
    const Solver& S = ((LmSession*)h)->S;
    int n = 0;
    for (const std::vector<double>* v : {&S.pq, &S.pt, &S.X, &S.J, &S.D, &S.mq, &S.mt})
        for (double d : *v) { if (n < cap) out[n] = d; ++n; }
    return n;
}
}

// ---------------------------------------------------------------------------------------
// Optimizer::PoseOptimization (src/Optimizer.cc:232-429): one free pose, unary OnlyPose edges
// (types_six_dof_expmap.cpp:266-364, .h:150-202), 4 rounds x 10 LM iterations, each round restarted
// from the initial pose with the inlier set of the previous one, robust kernel off in round 4.
namespace {
struct PoseEdge { double X[3], obs[3], w; bool stereo; };

inline void pose_edge_error(const adb_pose_problem& P, const double* R, const double* t, const PoseEdge& e, double* er, double* Xc) {
    for (int i = 0; i < 3; ++i) Xc[i] = R[i * 3] * e.X[0] + R[i * 3 + 1] * e.X[1] + R[i * 3 + 2] * e.X[2] + t[i];
    if (e.stereo) {
        const float invz = (float)(1.0 / Xc[2]);
        const double u = Xc[0] * invz * P.fx + P.cx, v = Xc[1] * invz * P.fy + P.cy;
        er[0] = e.obs[0] - u; er[1] = e.obs[1] - v; er[2] = e.obs[2] - (u - P.bf * invz);   // member bf is double here
    } else {
        er[0] = e.obs[0] - (Xc[0] / Xc[2] * P.fx + P.cx); er[1] = e.obs[1] - (Xc[1] / Xc[2] * P.fy + P.cy); er[2] = 0;
    }
}
inline void pose_edge_jac(const adb_pose_problem& P, const double* Xc, bool stereo, double* J) {
    const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz, fx = P.fx, fy = P.fy;
    J[0] = x * y * invz_2 * fx; J[1] = -(1 + (x * x * invz_2)) * fx; J[2] = y * invz * fx; J[3] = -invz * fx; J[4] = 0; J[5] = x * invz_2 * fx;
    J[6] = (1 + y * y * invz_2) * fy; J[7] = -x * y * invz_2 * fy; J[8] = -x * invz * fy; J[9] = 0; J[10] = -invz * fy; J[11] = y * invz_2 * fy;
    if (stereo) { J[12] = J[0] - P.bf * y * invz_2; J[13] = J[1] + P.bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - P.bf * invz_2; }
    else for (int i = 12; i < 18; ++i) J[i] = 0;
}

// BaseUnaryEdge::constructQuadraticForm (Thirdparty/g2o/g2o/core/base_unary_edge.hpp:40-69) for one OnlyPose edge, information w0 * I:
// h = J^T (rho' w0) J, g = -rho' J^T w0 e: this edge's contribution to the 6 x 6 system
inline void pose_edge_quadratic_form(int dim, const double* J, const double* er, double w0, const Huber& hub, bool robust, double* h, double* g) {
    double r0, r1;
    const double c = er[0] * (w0 * er[0]) + er[1] * (w0 * er[1]) + er[2] * (w0 * er[2]);
    robustify(hub, robust, c, &r0, &r1);
    const double w = r1 * w0;
    for (int u = 0; u < 6; ++u) {
        for (int v = 0; v < 6; ++v) { double s = 0; for (int k = 0; k < dim; ++k) s += J[k * 6 + u] * w * J[k * 6 + v]; h[u * 6 + v] = s; }
        double s = 0; for (int k = 0; k < dim; ++k) s += J[k * 6 + u] * (-w0 * er[k] * r1); g[u] = s;
    }
}

// One frame of PoseOptimization as solver steps (the same split as Solver above: evaluate / build / solve / update / push / pop), so that
// both the loop below and the reference's own control flow (oracle/ref_lm.cpp, ref_lba.cpp) can drive it.
struct PoseSolver {
    const adb_pose_problem& P;
    int f, a, n;
    std::vector<PoseEdge> E;
    std::vector<uint8_t> level;
    std::vector<double> chi;
    bool robust = true;
    Huber hmono = make_huber((double)(float)std::sqrt(5.991)), hstereo = make_huber((double)(float)std::sqrt(7.815));
    double q[4], t[3];
    double H[36], b[6], x[6], lambda = 0, ni = 2;
    std::vector<std::array<double, 7>> stack;
    std::vector<double> trace;   // (lambda, chi2 before, chi2 after, accepted) per trial

    PoseSolver(const adb_pose_problem& p, int frame) : P(p), f(frame), a(p.frame_ptr[frame]), n(p.frame_ptr[frame + 1] - p.frame_ptr[frame]) {
        E.resize(n);
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < 3; ++k) { E[i].X[k] = (double)P.xw[3 * (size_t)(a + i) + k]; E[i].obs[k] = (double)P.obs[3 * (size_t)(a + i) + k]; }
            E[i].w = (double)P.inv_sigma2[a + i];
            E[i].stereo = !(P.obs[3 * (size_t)(a + i) + 2] < 0);
        }
        level.assign(n, 0); chi.assign(n, 0.0);
        std::memcpy(q, P.pose_q + 4 * f, sizeof(q)); std::memcpy(t, P.pose_t + 3 * f, sizeof(t));
        std::memset(H, 0, sizeof(H)); std::memset(b, 0, sizeof(b)); std::memset(x, 0, sizeof(x));
    }
    int n_active() const { int c = 0; for (int i = 0; i < n; ++i) c += !level[i]; return c; }
    double evaluate() {          // computeActiveErrors + activeRobustChi2 at the current pose
        double R[9], s = 0, r0, r1;
        quat_to_rot(q, R);
        for (int i = 0; i < n; ++i) {
            if (level[i]) continue;
            double er[3], Xc[3];
            pose_edge_error(P, R, t, E[i], er, Xc);
            const double c = er[0] * (E[i].w * er[0]) + er[1] * (E[i].w * er[1]) + er[2] * (E[i].w * er[2]);
            chi[i] = c;
            robustify(E[i].stereo ? hstereo : hmono, robust, c, &r0, &r1);
            s += r0;
        }
        return s;
    }
    void build_system() {
        std::memset(H, 0, sizeof(H)); std::memset(b, 0, sizeof(b));
        double R[9];
        quat_to_rot(q, R);
        for (int i = 0; i < n; ++i) {
            if (level[i]) continue;
            double er[3], Xc[3], J[18];
            pose_edge_error(P, R, t, E[i], er, Xc);
            pose_edge_jac(P, Xc, E[i].stereo, J);
            double h[36], g[6];
            pose_edge_quadratic_form(E[i].stereo ? 3 : 2, J, er, E[i].w, E[i].stereo ? hstereo : hmono, robust, h, g);
            for (int u = 0; u < 36; ++u) H[u] += h[u];
            for (int u = 0; u < 6; ++u) b[u] += g[u];
        }
    }
    double max_diag() const { double m = 0; for (int u = 0; u < 6; ++u) m = std::max(m, std::fabs(H[u * 7])); return m; }
    bool solve_trial() {         // (H + lambda I) x = b by dense Cholesky (g2o: LinearSolverDense)
        std::vector<double> S(H, H + 36);
        for (int u = 0; u < 6; ++u) S[u * 7] += lambda;
        std::memcpy(x, b, sizeof(x));
        if (!cholesky(S, 6)) return false;
        chol_solve(S, 6, x);
        return true;
    }
    void apply_update() { pose_oplus(q, t, x); }
    double scale_term() const { double s = 0; for (int u = 0; u < 6; ++u) s += x[u] * (lambda * x[u] + b[u]); return s; }
    void push() { stack.push_back({q[0], q[1], q[2], q[3], t[0], t[1], t[2]}); }
    void pop() { const auto& v = stack.back(); for (int k = 0; k < 4; ++k) q[k] = v[k]; for (int k = 0; k < 3; ++k) t[k] = v[4 + k]; stack.pop_back(); }
    void discard_top() { stack.pop_back(); }

    // SparseOptimizer::optimize(iterations) + OptimizationAlgorithmLevenberg::solve, as Solver::optimize
    int optimize(int iterations) {
        int nbad_it = 0, it_run = 0;
        for (int it = 0; it < iterations; ++it) {
            double current = evaluate();
            const double ini = current;
            build_system();
            if (it == 0) { lambda = 1e-5 * max_diag(); ni = 2; nbad_it = 0; }
            double rho = 0; int qn = 0;
            do {
                push();
                const bool ok = solve_trial();
                if (ok) apply_update();
                double temp = evaluate();
                if (!ok) temp = std::numeric_limits<double>::max();
                rho = current - temp;
                double scale = ok ? scale_term() : 0.0;
                scale += 1e-3;
                rho /= scale;
                const bool good = rho > 0 && std::isfinite(temp);
                trace.insert(trace.end(), {lambda, current, temp, good ? 1.0 : 0.0});
                if (good) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    lambda *= std::max(1. / 3., alpha); ni = 2; current = temp;
                    discard_top();
                } else { lambda *= ni; ni *= 2; pop(); }
                ++qn;
            } while (rho < 0 && qn < 10);
            ++it_run;
            if (qn == 10 || rho == 0) break;
            if ((ini - current) * 1e3 < ini) ++nbad_it; else nbad_it = 0;
            if (nbad_it >= 3) break;
        }
        return it_run;
    }
};

int pose_optimize_one(const adb_pose_problem& P, int f, std::vector<double>* trace_out = nullptr) {
    PoseSolver S(P, f);
    const int a = S.a, n = S.n;
    for (int i = 0; i < n; ++i) P.outlier[a + i] = 0;
    if (n < 3) return 0;
    double q0[4], t0[3];
    std::memcpy(q0, S.q, sizeof(q0)); std::memcpy(t0, S.t, sizeof(t0));
    int nBad = 0;
    for (int round = 0; round < 4; ++round) {
        std::memcpy(S.q, q0, sizeof(q0)); std::memcpy(S.t, t0, sizeof(t0));   // vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw)) (:341)
        if (S.n_active() > 0) S.optimize(10);   // optimizer.optimize(10) with 0 active vertices returns immediately
        // classification (src/Optimizer.cc:361-416): outlier edges are re-evaluated at the round's final pose,
        // inlier edges keep the error of the last evaluated trial; float comparison
        double R[9];
        quat_to_rot(S.q, R);
        nBad = 0;
        for (int i = 0; i < n; ++i) {
            if (P.outlier[a + i]) {
                double er[3], Xc[3];
                pose_edge_error(P, R, S.t, S.E[i], er, Xc);
                S.chi[i] = er[0] * (S.E[i].w * er[0]) + er[1] * (S.E[i].w * er[1]) + er[2] * (S.E[i].w * er[2]);
            }
            const float c = (float)S.chi[i];
            if (c > (S.E[i].stereo ? 7.815f : 5.991f)) { P.outlier[a + i] = 1; S.level[i] = 1; ++nBad; }
            else { P.outlier[a + i] = 0; S.level[i] = 0; }
        }
        if (round == 2) S.robust = false;
        if (n < 10) break;
    }
    std::memcpy(P.pose_q + 4 * f, S.q, sizeof(S.q)); std::memcpy(P.pose_t + 3 * f, S.t, sizeof(S.t));
    if (trace_out) *trace_out = S.trace;
    return n - nBad;
}
}  // namespace

extern "C" void ba_oracle_pose_quadratic_form(int dim, const double* J, const double* er, double w0, double delta, int robust, double* h, double* g) {
    pose_edge_quadratic_form(dim, J, er, w0, make_huber(delta), robust != 0, h, g);
}
// ---- PoseSolver's steps one by one for oracle/ref_lm.cpp / ref_lba.cpp (same contract as ba_oracle_lm_*)
extern "C" {
struct PoseSession { PoseSolver S; double chi = 0; PoseSession(const adb_pose_problem& p, int f) : S(p, f) {} };
// level: [n] 1 = excluded (setLevel(1)); robust: Huber kernel on the active edges
void* ba_oracle_pose_lm_open(adb_pose_problem* P, int frame, const uint8_t* level, int robust) {
    PoseSession* h = new PoseSession(*P, frame);
    for (int i = 0; i < h->S.n; ++i) h->S.level[i] = level ? level[i] : 0;
    h->S.robust = robust != 0;
    return h;
}
void ba_oracle_pose_lm_close(void* h) { delete (PoseSession*)h; }
void ba_oracle_pose_lm_compute_errors(void* h) { PoseSession* s = (PoseSession*)h; s->chi = s->S.evaluate(); }
double ba_oracle_pose_lm_chi2(void* h) { return ((PoseSession*)h)->chi; }
void ba_oracle_pose_lm_build(void* h) { ((PoseSession*)h)->S.build_system(); }
int ba_oracle_pose_lm_layout(void*, int32_t* dims, int cap) { if (cap > 0) dims[0] = 6; return 1; }
int ba_oracle_pose_lm_vectors(void* h, double* x, double* b, double* diag, int cap) {
    const PoseSolver& S = ((PoseSession*)h)->S;
    for (int u = 0; u < 6 && u < cap; ++u) { if (x) x[u] = S.x[u]; if (b) b[u] = S.b[u]; if (diag) diag[u] = S.H[u * 7]; }
    return 6;
}
void ba_oracle_pose_lm_set_lambda(void* h, double lambda) { ((PoseSession*)h)->S.lambda = lambda; }
int ba_oracle_pose_lm_solve(void* h) { return ((PoseSession*)h)->S.solve_trial() ? 1 : 0; }
void ba_oracle_pose_lm_update(void* h) { ((PoseSession*)h)->S.apply_update(); }
void ba_oracle_pose_lm_push(void* h) { ((PoseSession*)h)->S.push(); }
void ba_oracle_pose_lm_pop(void* h) { ((PoseSession*)h)->S.pop(); }
void ba_oracle_pose_lm_discard_top(void* h) { ((PoseSession*)h)->S.discard_top(); }
int ba_oracle_pose_lm_state(void* h, double* out, int cap) {       // q (x y z w), t
    const PoseSolver& S = ((PoseSession*)h)->S;
    for (int k = 0; k < 7 && k < cap; ++k) out[k] = k < 4 ? S.q[k] : S.t[k - 4];
    return 7;
}
// the oracle's whole four-round schedule on one frame with the LM trials of all rounds: rows of (lambda, chi2 before, chi2 after, accepted)
int ba_oracle_pose_optimize_traced(adb_pose_problem* P, int frame, double* rows, int row_cap, int* n_rows) {
    std::vector<double> tr;
    const int inl = pose_optimize_one(*P, frame, &tr);
    *n_rows = (int)tr.size() / 4;
    for (size_t i = 0; i < tr.size() && (int)i < 4 * row_cap; ++i) rows[i] = tr[i];
    P->n_inliers[frame] = inl;
    return inl;
}
}
extern "C" int ba_oracle_pose_optimize(adb_pose_problem* P) {
    for (int f = 0; f < P->n_frames; ++f) P->n_inliers[f] = pose_optimize_one(*P, f);
    return ADB_OK;
}

// Edge(Stereo)SE3ProjectXYZOnlyPose: error and d e / d pose of one correspondence (pinning tests); cam5 = fx fy cx cy bf
extern "C" void ba_oracle_pose_edge(const double* q, const double* t, const double* Xw, const double* obs, int stereo, const double* cam5, double* er,
                                    double* J18) {
    adb_pose_problem P{};
    P.fx = cam5[0]; P.fy = cam5[1]; P.cx = cam5[2]; P.cy = cam5[3]; P.bf = cam5[4];
    PoseEdge e{};
    for (int i = 0; i < 3; ++i) { e.X[i] = Xw[i]; e.obs[i] = obs[i]; }
    e.stereo = stereo != 0;
    double R[9], Xc[3];
    quat_to_rot(q, R);
    pose_edge_error(P, R, t, e, er, Xc);
    pose_edge_jac(P, Xc, e.stereo, J18);
}
