// ref_leaf.cpp -- TEST INFRASTRUCTURE: the reference's own leaf arithmetic of the BA path, compiled from the UNMODIFIED sources
// under /root/reference (recipe: oracle/Makefile, target _ref/libref_leaf.so; built only where /root/reference exists):
//   Thirdparty/g2o/g2o/types/se3_ops.hpp, se3quat.h             SE3Quat::exp / map / operator* / normalizeRotation
//   Thirdparty/g2o/g2o/types/types_sba.{h,cpp}                  VertexSBAPointXYZ::oplusImpl
//   Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp}       VertexSE3Expmap::oplusImpl, Edge(Stereo)SE3ProjectXYZ[OnlyPose]
//                                                               computeError / linearizeOplus / cam_project
//   include/g2o_vertex_distance.h, g2o_vertex_se3.h             VertexDistanceDouble, VertexSE3::oplusImpl
//   include/g2o_edge_rigidbody.h, g2o_dyn_slam3d.h              EdgeRigidBodyDouble::computeError, LandmarkMotionTernaryEdge
// against oracle/ref_shim (an Eigen stand-in and stubs of the g2o base classes; Eigen and g2o's graph core cannot be built here).
// The extern "C" functions below only move numbers in and out of those classes.  oracle/gen_ref_leaf_golden.py turns their
// outputs into tests/golden/ba_leaf_ref.npz, which pins oracle/ba_oracle.cpp and the CUDA kernels to the literal reference.
#include "ref_shim/g2o_core_stub.h"

#include "Thirdparty/g2o/g2o/types/types_sba.cpp"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp"
#include "include/g2o_edge_rigidbody.h"
#include "include/g2o_dyn_slam3d.h"

namespace {
g2o::SE3Quat make_pose(const double* q /*x y z w*/, const double* t) {
    return g2o::SE3Quat(Eigen::Quaterniond(q[3], q[0], q[1], q[2]), Eigen::Vector3d(t[0], t[1], t[2]));
}
void put_pose(const g2o::SE3Quat& T, double* q, double* t) {
    q[0] = T.rotation().x(); q[1] = T.rotation().y(); q[2] = T.rotation().z(); q[3] = T.rotation().w();
    for (int i = 0; i < 3; ++i) t[i] = T.translation()[i];
}
template <typename M> void put_rowmajor(const M& m, int rows, int cols, double* out) {
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) out[r * cols + c] = m(r, c);
}
}  // namespace

extern "C" {

// VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:73-76): T <- SE3Quat::exp(update) * T
void ref_pose_oplus(const double* q, const double* t, const double* update6, double* q_out, double* t_out) {
    g2o::VertexSE3Expmap v;
    v.setEstimate(make_pose(q, t));
    v.oplusImpl(update6);
    put_pose(v.estimate(), q_out, t_out);
}
// SE3Quat(R, t) as Converter::toSE3Quat builds it (src/Converter.cc:37-47) from a rotation matrix (row-major 3 x 3)
void ref_pose_from_rt(const double* R9, const double* t3, double* q_out, double* t_out) {
    Eigen::Matrix3d R;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = R9[r * 3 + c];
    put_pose(g2o::SE3Quat(R, Eigen::Vector3d(t3[0], t3[1], t3[2])), q_out, t_out);
}
// SE3Quat::to_homogeneous_matrix (se3quat.h) -> row-major 4 x 4, the double half of Converter::toCvMat(SE3Quat)
void ref_pose_to_matrix(const double* q, const double* t, double* T16) { put_rowmajor(make_pose(q, t).to_homogeneous_matrix(), 4, 4, T16); }

// EdgeStereoSE3ProjectXYZ / EdgeSE3ProjectXYZ: error (obs - projection), d e / d point (dim x 3), d e / d pose (dim x 6), row-major
void ref_edge_stereo(const double* q, const double* t, const double* X, const double* obs3, const double* cam5, double* err3, double* Ji9, double* Jj18,
                     int* depth_positive) {
    g2o::VertexSE3Expmap pose; pose.setEstimate(make_pose(q, t));
    g2o::VertexSBAPointXYZ pt; pt.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
    g2o::EdgeStereoSE3ProjectXYZ e;
    e.setVertex(0, &pt); e.setVertex(1, &pose);
    e.fx = cam5[0]; e.fy = cam5[1]; e.cx = cam5[2]; e.cy = cam5[3]; e.bf = cam5[4];
    e.setMeasurement(Eigen::Vector3d(obs3[0], obs3[1], obs3[2]));
    e.computeError();
    e.linearizeOplus();
    for (int i = 0; i < 3; ++i) err3[i] = e.error()[i];
    put_rowmajor(e._jacobianOplusXi, 3, 3, Ji9); put_rowmajor(e._jacobianOplusXj, 3, 6, Jj18);
    *depth_positive = e.isDepthPositive() ? 1 : 0;
}
void ref_edge_mono(const double* q, const double* t, const double* X, const double* obs2, const double* cam5, double* err2, double* Ji6, double* Jj12,
                   int* depth_positive) {
    g2o::VertexSE3Expmap pose; pose.setEstimate(make_pose(q, t));
    g2o::VertexSBAPointXYZ pt; pt.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
    g2o::EdgeSE3ProjectXYZ e;
    e.setVertex(0, &pt); e.setVertex(1, &pose);
    e.fx = cam5[0]; e.fy = cam5[1]; e.cx = cam5[2]; e.cy = cam5[3];
    e.setMeasurement(Eigen::Vector2d(obs2[0], obs2[1]));
    e.computeError();
    e.linearizeOplus();
    for (int i = 0; i < 2; ++i) err2[i] = e.error()[i];
    put_rowmajor(e._jacobianOplusXi, 2, 3, Ji6); put_rowmajor(e._jacobianOplusXj, 2, 6, Jj12);
    *depth_positive = e.isDepthPositive() ? 1 : 0;
}
// Edge(Stereo)SE3ProjectXYZOnlyPose (PoseOptimization): error and d e / d pose
void ref_edge_stereo_onlypose(const double* q, const double* t, const double* Xw, const double* obs3, const double* cam5, double* err3, double* J18) {
    g2o::VertexSE3Expmap pose; pose.setEstimate(make_pose(q, t));
    g2o::EdgeStereoSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &pose);
    e.fx = cam5[0]; e.fy = cam5[1]; e.cx = cam5[2]; e.cy = cam5[3]; e.bf = cam5[4];
    e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.setMeasurement(Eigen::Vector3d(obs3[0], obs3[1], obs3[2]));
    e.computeError();
    e.linearizeOplus();
    for (int i = 0; i < 3; ++i) err3[i] = e.error()[i];
    put_rowmajor(e._jacobianOplusXi, 3, 6, J18);
}
void ref_edge_mono_onlypose(const double* q, const double* t, const double* Xw, const double* obs2, const double* cam5, double* err2, double* J12) {
    g2o::VertexSE3Expmap pose; pose.setEstimate(make_pose(q, t));
    g2o::EdgeSE3ProjectXYZOnlyPose e;
    e.setVertex(0, &pose);
    e.fx = cam5[0]; e.fy = cam5[1]; e.cx = cam5[2]; e.cy = cam5[3];
    e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
    e.setMeasurement(Eigen::Vector2d(obs2[0], obs2[1]));
    e.computeError();
    e.linearizeOplus();
    for (int i = 0; i < 2; ++i) err2[i] = e.error()[i];
    put_rowmajor(e._jacobianOplusXi, 2, 6, J12);
}
// EdgeRigidBodyDouble::computeError (include/g2o_edge_rigidbody.h:139-149): |p_from - p_to| - d.  Its linearizeOplus reads members
// that computeError never assigns (it shadows them with locals: undefined behaviour) and is therefore not exposed.
double ref_rigid_error(const double* p_from, const double* p_to, double d) {
    g2o::VertexSBAPointXYZ a, b; a.setEstimate(Eigen::Vector3d(p_from[0], p_from[1], p_from[2])); b.setEstimate(Eigen::Vector3d(p_to[0], p_to[1], p_to[2]));
    VertexDistanceDouble dist; dist.setEstimate(d);
    EdgeRigidBodyDouble e;
    e.setVertex(0, &a); e.setVertex(1, &b); e.setVertex(2, &dist);
    e.computeError();
    return e.error()[0];
}
// LandmarkMotionTernaryEdge (include/g2o_dyn_slam3d.h:48-101) with zero measurement: error, and the three Jacobians of the FIRST
// linearizeOplus call on a fresh edge (the member J is rescaled by delta_t on every call: SURVEY.md D.6)
void ref_motion_edge(const double* p1, const double* p2, const double* mq, const double* mt, double dt, double* err3, double* J1_9, double* J2_9, double* J3_18) {
    g2o::VertexSBAPointXYZ a, b; a.setEstimate(Eigen::Vector3d(p1[0], p1[1], p1[2])); b.setEstimate(Eigen::Vector3d(p2[0], p2[1], p2[2]));
    VertexSE3 H;
    Isometry3 M;
    M = Eigen::Quaterniond(mq[3], mq[0], mq[1], mq[2]).toRotationMatrix();
    M.translation() = Eigen::Vector3d(mt[0], mt[1], mt[2]);
    H.setEstimate(M);
    LandmarkMotionTernaryEdge e;
    e.setVertex(0, &a); e.setVertex(1, &b); e.setVertex(2, &H);
    e.delta_t = dt;
    e.setMeasurement(Eigen::Vector3d(0, 0, 0));
    e.computeError();
    e.linearizeOplus();
    for (int i = 0; i < 3; ++i) err3[i] = e.error()[i];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { J1_9[r * 3 + c] = e._jacobianOplus[0](r, c); J2_9[r * 3 + c] = e._jacobianOplus[1](r, c); }
        for (int c = 0; c < 6; ++c) J3_18[r * 6 + c] = e._jacobianOplus[2](r, c);
    }
}
// VertexSE3::oplusImpl (include/g2o_vertex_se3.h:113-122): H <- H * fromVectorMQT(update); in / out as rotation matrix (row-major) + translation
void ref_motion_oplus(const double* R9, const double* t3, const double* update6, double* R9_out, double* t3_out) {
    VertexSE3 H;
    Isometry3 M;
    Matrix3 R;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = R9[r * 3 + c];
    M = R;
    M.translation() = Eigen::Vector3d(t3[0], t3[1], t3[2]);
    H.setEstimate(M);
    H.oplusImpl(update6);
    const Isometry3& o = H.estimate();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R9_out[r * 3 + c] = o.matrix()(r, c); t3_out[r] = o.matrix()(r, 3); }
}
// VertexSBAPointXYZ::oplusImpl / VertexDistanceDouble::oplusImpl
void ref_point_oplus(const double* X, const double* d3, double* out3) {
    g2o::VertexSBAPointXYZ v; v.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
    v.oplusImpl(d3);
    for (int i = 0; i < 3; ++i) out3[i] = v.estimate()[i];
}
double ref_dist_oplus(double d, double update) { VertexDistanceDouble v; v.setEstimate(d); v.oplusImpl(&update); return v.estimate(); }

}  // extern "C"
