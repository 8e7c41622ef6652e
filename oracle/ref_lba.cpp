// oracle/ref_lba.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Built only where /root/reference exists (make -C oracle ref ->
// oracle/_ref/libref_lba.so, git-ignored).
//
// The reference's OWN Optimizer::LocalBundleAdjustment (src/Optimizer.cc:431-731), the whole function definition, compiled from
// /root/reference: window selection (local / fixed key-frames from the covisibility list and the observations), vertex ids and the
// fixed flag, one Edge(Stereo)SE3ProjectXYZ per observation with information mvInvLevelSigma2[octave] * I and a Huber kernel
// (deltas sqrt(5.991) / sqrt(7.815) held in floats), optimize(5), the chi2 / depth gate (setLevel(1)) and setRobustKernel(0),
// initializeOptimization(0) + optimize(10), the final gate -> erase list, and the recovery of poses / points through Converter.
// The function text is taken out of the reference tree at build time (oracle/extract_ref_fn.py -> oracle/_ref/lba_snippets.inc) and
// compiled between stand-in declarations of KeyFrame / MapPoint / Map (the members the statements touch, with the reference's names
// and types).  The g2o vertex and edge TYPES are the reference's own sources (types_sba.cpp, types_six_dof_expmap.cpp, compiled as in
// ref_leaf.cpp), so chi2() / isDepthPositive() / computeError() in the gates are the literal ones, on errors the literal
// computeError() left behind.  Converter::toSE3Quat / toCvMat / toVector3d are the reference's (src/Converter.cc).
// What is NOT the reference: g2o::SparseOptimizer is a stand-in that keeps the graph and, on optimize(n), hands the active part to
// the oracle's solver session (oracle/ba_oracle.cpp, ba_oracle_lm_*) driven by the reference's own Levenberg-Marquardt control
// (oracle/ref_lm.cpp); after every step of that control it copies the session's estimates into the real vertices and lets the real
// edges recompute their errors, as g2o's computeActiveErrors would.  So a run is: the reference's schedule and gates, the reference's
// LM control, the reference's leaf types for the gate -- over the oracle's linear algebra.  tests/test_ref_lba.py compares it with
// ba_oracle_solve on the problem the stand-in recorded.
#define G2O_STUB_WITH_GRAPH_MEMBERS
#include "ref_shim/g2o_core_stub.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "Thirdparty/g2o/g2o/types/types_sba.cpp"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp"
#include "include/g2o_edge_rigidbody.h"      // VertexDistanceDouble, EdgeRigidBodyDouble            (AirDOS, unmodified)
#include "include/g2o_dyn_slam3d.h"          // VertexSE3 (Isometry3), LandmarkMotionTernaryEdge      (AirDOS, unmodified)

#include <cstdio>
#include <fstream>

#include "ref_shim/cv_shim.h"
#include "../include/airdos_b200.h"      // adb_ba_problem / adb_ba_options: what the oracle's session takes

extern "C" {
struct ref_lm_hooks {          // as in oracle/ref_lm.cpp
    void* ctx;
    void (*compute_errors)(void*);
    double (*chi2)(void*);
    void (*build)(void*);
    int (*layout)(void*, int32_t*, int);
    int (*vectors)(void*, double*, double*, double*, int);
    void (*set_lambda)(void*, double);
    int (*solve)(void*);
    void (*update)(void*);
    void (*push)(void*);
    void (*pop)(void*);
    void (*discard_top)(void*);
};
struct ref_lba_backend {       // the oracle's session functions + the reference LM driver, handed in by the test
    ref_lm_hooks steps;        // ctx unused: filled per session
    void* (*open)(adb_ba_problem*, const adb_ba_options*, int robust);
    void (*close)(void*);
    void (*set_levels)(void*, const uint8_t*);
    int (*state)(void*, double*, int);
    int (*lm_optimize)(const ref_lm_hooks*, int iterations, int max_trials, double* rows, int row_cap, int* n_rows, double* lambda_final,
                       int* n_error_evaluations, double* tau);
    void (*default_options)(adb_ba_options*);
    // PoseOptimization: the oracle's one-frame pose session (ba_oracle_pose_lm_*)
    ref_lm_hooks pose_steps;
    void* (*pose_open)(adb_pose_problem*, int frame, const uint8_t* level, int robust);
    void (*pose_close)(void*);
    int (*pose_state)(void*, double*, int);
    void (*set_levels4)(void*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*);   // static / joint / rigidity / motion edges
};
}

namespace std {
struct lba_null_stream_t {
    template <class T> lba_null_stream_t& operator<<(const T&) { return *this; }
    lba_null_stream_t& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
static lba_null_stream_t lba_null_stream;
}
namespace g2o {
class RobustKernelHuber : public RobustKernel {     // core/robust_kernel_impl.h:76-85
public:
    virtual void setDelta(double delta);
    virtual void setDeltaSqr(const double& delta, const double& deltaSqr) { dsqr = deltaSqr; _delta = delta; }
    virtual void robustify(double e2, Eigen::Vector3d& rho) const;
private:
    float dsqr;
};
#include "_ref/lm_huber.inc"

// solver construction as Optimizer.cc writes it: the objects only have to exist
template <typename T> struct LinearSolver {};
template <typename T> struct LinearSolverEigen : LinearSolver<T> {};
template <typename T> struct LinearSolverDense : LinearSolver<T> {};
struct BlockSolverX {
    typedef Eigen::MatrixXd PoseMatrixType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
    explicit BlockSolverX(LinearSolverType*) {}
};
struct BlockSolver_6_3 {
    typedef Eigen::Matrix<double, 6, 6> PoseMatrixType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
    explicit BlockSolver_6_3(LinearSolverType*) {}
};
struct OptimizationAlgorithmLevenberg {
    explicit OptimizationAlgorithmLevenberg(BlockSolver_6_3*) {}
    explicit OptimizationAlgorithmLevenberg(BlockSolverX*) {}
};

struct LbaRecord {             // what the stand-in saw: the first round's problem and every LM trial
    std::vector<double> pose_q, pose_t, points, edge_obs, edge_info, rows;
    std::vector<uint8_t> pose_fixed;
    std::vector<int32_t> edge_pose, edge_point, pose_id, point_id;
    double cam[5] = {0, 0, 0, 0, 0}, huber_mono = 0, huber_stereo = 0;
    std::vector<int32_t> round_iterations, round_robust;
    std::vector<double> final_state;
    // PoseOptimization: the frame's correspondences as the function turned them into edges, and the active count of every round
    std::vector<float> xw, obs, inv_sigma2;
    std::vector<int32_t> round_active;
    // articulated part (LocalBundleAdjustmentHumanTrajactory) and per-edge facts at the end of the function, by name
    std::map<std::string, std::vector<double>> named;
};

class SparseOptimizer {
public:
    const ref_lba_backend* be = nullptr;
    LbaRecord* rec = nullptr;
    std::map<int, OptimizableGraph::Vertex*> vmap;
    struct EdgeRef {
        EdgeSE3ProjectXYZ* mono; EdgeStereoSE3ProjectXYZ* stereo; EdgeRigidBodyDouble* rigid; LandmarkMotionTernaryEdge* motion;
        int level() const { return mono ? mono->level() : stereo ? stereo->level() : rigid ? rigid->level() : motion->level(); }
        RobustKernel* kernel() const { return mono ? mono->robustKernel() : stereo ? stereo->robustKernel() : rigid ? rigid->robustKernel() : motion->robustKernel(); }
        void compute_error() const { if (mono) mono->computeError(); else if (stereo) stereo->computeError(); else if (rigid) rigid->computeError(); else motion->computeError(); }
        double chi2() const { return mono ? mono->chi2() : stereo ? stereo->chi2() : rigid ? rigid->chi2() : motion->chi2(); }
        HyperGraph::Vertex* v(int i) const { return mono ? mono->_vertices[i] : stereo ? stereo->_vertices[i] : rigid ? rigid->_vertices[i] : motion->_vertices[i]; }
    };
    std::vector<EdgeRef> ba_edges;       // insertion order = g2o's internal edge ids = the order of _activeEdges
    struct PoseEdgeRef { EdgeSE3ProjectXYZOnlyPose* mono; EdgeStereoSE3ProjectXYZOnlyPose* stereo; };
    std::vector<PoseEdgeRef> pose_edges;
    struct EdgeCount { size_t n; size_t size() const { return n; } };
    EdgeCount edges() const { return EdgeCount{ba_edges.size() + pose_edges.size()}; }   // optimizer.edges().size() (:418)
    int active_level = 0;
    bool* stop = nullptr;

    void setAlgorithm(OptimizationAlgorithmLevenberg*) {}
    void setForceStopFlag(bool* f) { stop = f; }
    bool addVertex(OptimizableGraph::Vertex* v) { vmap[v->id()] = v; return true; }
    bool removeVertex(OptimizableGraph::Vertex* v) { vmap.erase(v->id()); return true; }     // a point without edges (src/Optimizer.cc:178-180)
    bool addEdge(EdgeSE3ProjectXYZ* e) { ba_edges.push_back(EdgeRef{e, nullptr, nullptr, nullptr}); return true; }
    bool addEdge(EdgeStereoSE3ProjectXYZ* e) { ba_edges.push_back(EdgeRef{nullptr, e, nullptr, nullptr}); return true; }
    bool addEdge(EdgeRigidBodyDouble* e) { ba_edges.push_back(EdgeRef{nullptr, nullptr, e, nullptr}); return true; }
    bool addEdge(LandmarkMotionTernaryEdge* e) { ba_edges.push_back(EdgeRef{nullptr, nullptr, nullptr, e}); return true; }
    ~SparseOptimizer();
    bool addEdge(EdgeSE3ProjectXYZOnlyPose* e) { pose_edges.push_back(PoseEdgeRef{e, nullptr}); return true; }
    bool addEdge(EdgeStereoSE3ProjectXYZOnlyPose* e) { pose_edges.push_back(PoseEdgeRef{nullptr, e}); return true; }
    OptimizableGraph::Vertex* vertex(int id) { auto it = vmap.find(id); return it == vmap.end() ? nullptr : it->second; }
    bool initializeOptimization(int level = 0) { active_level = level; return true; }
    int optimize(int iterations) { return pose_edges.empty() ? optimize_ba(iterations) : optimize_pose(iterations); }
    int optimize_ba(int iterations);
    int optimize_pose(int iterations);

    // ---- session plumbing
    struct Run {
        SparseOptimizer* self; void* session;
        std::vector<VertexSE3Expmap*> poses; std::vector<VertexSBAPointXYZ*> points, joints;
        std::vector<VertexDistanceDouble*> dists; std::vector<VertexSE3*> motions;
        std::vector<uint8_t> lvl;          // per edge of ba_edges, insertion order
        void sync() {          // the session's estimates -> the real vertices (layout of ba_oracle_lm_state)
            const int n = self->be->state(session, nullptr, 0);
            std::vector<double> s(n);
            self->be->state(session, s.data(), n);
            const size_t np = poses.size(), nx = points.size(), nj = joints.size(), nd = dists.size(), nm = motions.size();
            const double* pq = s.data(); const double* pt = pq + 4 * np; const double* X = pt + 3 * np; const double* J = X + 3 * nx;
            const double* D = J + 3 * nj; const double* mq = D + nd; const double* mt = mq + 4 * nm;
            for (size_t i = 0; i < np; ++i) {
                SE3Quat T = poses[i]->estimate();
                T.setRotation(Eigen::Quaterniond(pq[4 * i + 3], pq[4 * i], pq[4 * i + 1], pq[4 * i + 2]));
                T.setTranslation(Eigen::Vector3d(pt[3 * i], pt[3 * i + 1], pt[3 * i + 2]));
                poses[i]->setEstimate(T);
            }
            for (size_t l = 0; l < nx; ++l) points[l]->setEstimate(Eigen::Vector3d(X[3 * l], X[3 * l + 1], X[3 * l + 2]));
            for (size_t l = 0; l < nj; ++l) joints[l]->setEstimate(Eigen::Vector3d(J[3 * l], J[3 * l + 1], J[3 * l + 2]));
            for (size_t l = 0; l < nd; ++l) dists[l]->setEstimate(D[l]);
            for (size_t l = 0; l < nm; ++l) {
                Isometry3 T;
                T = Eigen::Quaterniond(mq[4 * l + 3], mq[4 * l], mq[4 * l + 1], mq[4 * l + 2]).toRotationMatrix();
                T.translation() = Eigen::Vector3d(mt[3 * l], mt[3 * l + 1], mt[3 * l + 2]);
                motions[l]->setEstimate(T);
            }
        }
        void real_errors() {   // g2o's computeActiveErrors on the real edges (core/sparse_optimizer.cpp:70-97): active edges only
            for (size_t e = 0; e < self->ba_edges.size(); ++e) if (!lvl[e]) self->ba_edges[e].compute_error();
        }
    };
    static void h_compute(void* c) { Run* r = (Run*)c; r->self->be->steps.compute_errors(r->session); r->real_errors(); }
    static double h_chi2(void* c) { Run* r = (Run*)c; return r->self->be->steps.chi2(r->session); }
    static void h_build(void* c) { Run* r = (Run*)c; r->self->be->steps.build(r->session); }
    static int h_layout(void* c, int32_t* d, int n) { Run* r = (Run*)c; return r->self->be->steps.layout(r->session, d, n); }
    static int h_vectors(void* c, double* x, double* b, double* dg, int n) { Run* r = (Run*)c; return r->self->be->steps.vectors(r->session, x, b, dg, n); }
    static void h_lambda(void* c, double l) { Run* r = (Run*)c; r->self->be->steps.set_lambda(r->session, l); }
    static int h_solve(void* c) { Run* r = (Run*)c; return r->self->be->steps.solve(r->session); }
    static void h_update(void* c) { Run* r = (Run*)c; r->self->be->steps.update(r->session); r->sync(); }
    static void h_push(void* c) { Run* r = (Run*)c; r->self->be->steps.push(r->session); }
    static void h_pop(void* c) { Run* r = (Run*)c; r->self->be->steps.pop(r->session); r->sync(); }
    static void h_discard(void* c) { Run* r = (Run*)c; r->self->be->steps.discard_top(r->session); }
};

int SparseOptimizer::optimize_ba(int iterations) {
    const std::vector<EdgeRef>& edges = ba_edges;
    Run run; run.self = this;
    std::map<HyperGraph::Vertex*, int> index;      // vertex -> index within its own kind
    for (auto& kv : vmap) {      // g2o orders the vertices of the index mapping by id (std::map: ascending)
        if (VertexSE3Expmap* p = dynamic_cast<VertexSE3Expmap*>(kv.second)) { index[p] = (int)run.poses.size(); run.poses.push_back(p); }
        else if (VertexSBAPointXYZ* x = dynamic_cast<VertexSBAPointXYZ*>(kv.second)) {
            if (x->marginalized()) { index[x] = (int)run.points.size(); run.points.push_back(x); }      // map points: setMarginalized(true)
            else { index[x] = (int)run.joints.size(); run.joints.push_back(x); }                       // human keys stay in the reduced system
        }
        else if (VertexDistanceDouble* d = dynamic_cast<VertexDistanceDouble*>(kv.second)) { index[d] = (int)run.dists.size(); run.dists.push_back(d); }
        else if (VertexSE3* m = dynamic_cast<VertexSE3*>(kv.second)) { index[m] = (int)run.motions.size(); run.motions.push_back(m); }
    }
    const int np = (int)run.poses.size(), nx = (int)run.points.size(), nj = (int)run.joints.size(), nd = (int)run.dists.size(), nm = (int)run.motions.size();
    std::vector<double> pq(4 * np), pt(3 * np), X(3 * nx), J(3 * nj), D(nd), mq(4 * nm), mt(3 * nm);
    std::vector<uint8_t> fixed(np);
    std::vector<int32_t> pid(np), xid(nx), jid(nj), did(nd), mid(nm);
    for (int i = 0; i < np; ++i) {
        const SE3Quat& T = run.poses[i]->estimate();
        pq[4 * i] = T.rotation().x(); pq[4 * i + 1] = T.rotation().y(); pq[4 * i + 2] = T.rotation().z(); pq[4 * i + 3] = T.rotation().w();
        for (int k = 0; k < 3; ++k) pt[3 * i + k] = T.translation()[k];
        fixed[i] = run.poses[i]->fixed(); pid[i] = run.poses[i]->id();
    }
    for (int l = 0; l < nx; ++l) { for (int k = 0; k < 3; ++k) X[3 * l + k] = run.points[l]->estimate()[k]; xid[l] = run.points[l]->id(); }
    for (int l = 0; l < nj; ++l) { for (int k = 0; k < 3; ++k) J[3 * l + k] = run.joints[l]->estimate()[k]; jid[l] = run.joints[l]->id(); }
    for (int l = 0; l < nd; ++l) { D[l] = run.dists[l]->estimate(); did[l] = run.dists[l]->id(); }
    for (int l = 0; l < nm; ++l) {
        const Isometry3& T = run.motions[l]->estimate();
        const Eigen::Quaterniond q(T.rotation());
        mq[4 * l] = q.x(); mq[4 * l + 1] = q.y(); mq[4 * l + 2] = q.z(); mq[4 * l + 3] = q.w();
        for (int k = 0; k < 3; ++k) mt[3 * l + k] = T.matrix()(k, 3);
        mid[l] = run.motions[l]->id();
    }
    adb_ba_problem P{};
    adb_ba_options O{};
    be->default_options(&O);
    // edges by kind, each kind in insertion order
    std::vector<double> obs, info, jobs, jinfo, rinfo, minfo, mdt;
    std::vector<int32_t> ep, ex, jp, jj, ri, rj, rd, m1, m2, mm, kind;
    std::vector<uint8_t> le, lj, lr, lm;
    int n_kernel = 0, n_active = 0;
    run.lvl.assign(edges.size(), 0);
    for (size_t e = 0; e < edges.size(); ++e) {
        const EdgeRef& r = edges[e];
        const uint8_t lv = r.level() != active_level;
        run.lvl[e] = lv;
        RobustKernel* k = r.kernel();
        if (!lv) { ++n_active; n_kernel += k != nullptr; }
        if (r.mono) {
            kind.push_back(0);
            ex.push_back(index.at(r.v(0))); ep.push_back(index.at(r.v(1)));
            obs.insert(obs.end(), {r.mono->measurement()[0], r.mono->measurement()[1], -1.0});
            info.push_back(r.mono->information()(0, 0)); le.push_back(lv);
            P.fx = r.mono->fx; P.fy = r.mono->fy; P.cx = r.mono->cx; P.cy = r.mono->cy;
            if (!lv && k) O.huber_mono = k->delta();
        } else if (r.stereo) {
            const bool joint = !static_cast<VertexSBAPointXYZ*>(r.v(0))->marginalized();
            kind.push_back(joint ? 2 : 1);
            P.fx = r.stereo->fx; P.fy = r.stereo->fy; P.cx = r.stereo->cx; P.cy = r.stereo->cy; P.bf = r.stereo->bf;
            if (!lv && k) O.huber_stereo = k->delta();
            if (joint) {
                jj.push_back(index.at(r.v(0))); jp.push_back(index.at(r.v(1)));
                for (int c = 0; c < 3; ++c) jobs.push_back(r.stereo->measurement()[c]);
                jinfo.push_back(r.stereo->information()(0, 0)); lj.push_back(lv);
            } else {
                ex.push_back(index.at(r.v(0))); ep.push_back(index.at(r.v(1)));
                for (int c = 0; c < 3; ++c) obs.push_back(r.stereo->measurement()[c]);
                info.push_back(r.stereo->information()(0, 0)); le.push_back(lv);
            }
        } else if (r.rigid) {
            kind.push_back(3);
            ri.push_back(index.at(r.v(0))); rj.push_back(index.at(r.v(1))); rd.push_back(index.at(r.v(2)));
            rinfo.push_back(r.rigid->information()(0, 0)); lr.push_back(lv);
            if (!lv && k) O.huber_rigid = k->delta();
        } else {
            kind.push_back(4);
            m1.push_back(index.at(r.v(0))); m2.push_back(index.at(r.v(1))); mm.push_back(index.at(r.v(2)));
            minfo.push_back(r.motion->information()(0, 0)); mdt.push_back(r.motion->delta_t); lm.push_back(lv);
            if (!lv && k) O.huber_motion = k->delta();
        }
    }
    if (n_kernel != 0 && n_kernel != n_active) return -2;     // mixed kernels: only with bad map points, which the stand-ins never report
    const int robust = n_kernel != 0;
    P.n_poses = np; P.n_points = nx; P.n_edges = (int)ep.size();
    P.pose_q = pq.data(); P.pose_t = pt.data(); P.pose_fixed = fixed.data(); P.points = X.data();
    P.edge_pose = ep.data(); P.edge_point = ex.data(); P.edge_obs = obs.data(); P.edge_info = info.data();
    P.n_joints = nj; P.joints = J.data(); P.n_joint_edges = (int)jp.size(); P.jedge_pose = jp.data(); P.jedge_joint = jj.data(); P.jedge_obs = jobs.data();
    P.jedge_info = jinfo.data();
    P.n_dists = nd; P.dists = D.data(); P.n_rigid_edges = (int)ri.size(); P.redge_i = ri.data(); P.redge_j = rj.data(); P.redge_dist = rd.data();
    P.redge_info = rinfo.data();
    P.n_motions = nm; P.motion_q = mq.data(); P.motion_t = mt.data(); P.n_motion_edges = (int)m1.size(); P.medge_p1 = m1.data(); P.medge_p2 = m2.data();
    P.medge_motion = mm.data(); P.medge_dt = mdt.data(); P.medge_info = minfo.data();
    if (rec && rec->round_iterations.empty()) {
        rec->pose_q = pq; rec->pose_t = pt; rec->points = X; rec->edge_obs = obs; rec->edge_info = info; rec->pose_fixed = fixed;
        rec->edge_pose = ep; rec->edge_point = ex; rec->pose_id = pid; rec->point_id = xid;
        rec->cam[0] = P.fx; rec->cam[1] = P.fy; rec->cam[2] = P.cx; rec->cam[3] = P.cy; rec->cam[4] = P.bf;
        rec->huber_mono = O.huber_mono; rec->huber_stereo = O.huber_stereo;
        auto put = [&](const char* name, const auto& v) { rec->named[name].assign(v.begin(), v.end()); };
        put("joints", J); put("dists", D); put("motion_q", mq); put("motion_t", mt);
        put("jedge_pose", jp); put("jedge_joint", jj); put("jedge_obs", jobs); put("jedge_info", jinfo);
        put("redge_i", ri); put("redge_j", rj); put("redge_dist", rd); put("redge_info", rinfo);
        put("medge_p1", m1); put("medge_p2", m2); put("medge_motion", mm); put("medge_dt", mdt); put("medge_info", minfo);
        put("joint_id", jid); put("dist_id", did); put("motion_id", mid); put("edge_kind", kind);
        rec->named["huber_rigid"] = {O.huber_rigid}; rec->named["huber_motion"] = {O.huber_motion};
    }
    run.session = be->open(&P, &O, robust);
    be->set_levels4(run.session, le.data(), lj.data(), lr.data(), lm.data());
    ref_lm_hooks hk{&run, h_compute, h_chi2, h_build, h_layout, h_vectors, h_lambda, h_solve, h_update, h_push, h_pop, h_discard};
    std::vector<double> rows(4 * 512);
    int n_rows = 0, n_eval = 0; double lam = 0, tau = 0;
    const int it = (stop && *stop) ? 0 : be->lm_optimize(&hk, iterations, 10, rows.data(), 512, &n_rows, &lam, &n_eval, &tau);
    run.sync();
    if (rec) {
        rec->rows.insert(rec->rows.end(), rows.begin(), rows.begin() + 4 * (size_t)std::min(n_rows, 512));
        rec->round_iterations.push_back(it); rec->round_robust.push_back(robust);
        const int n = be->state(run.session, nullptr, 0);
        rec->final_state.resize(n);
        be->state(run.session, rec->final_state.data(), n);
    }
    be->close(run.session);
    return it;
}

// at the end of the Optimizer function: what the reference's own edges say (chi2() of the error they hold, the depth test at the final
// estimates), per edge in insertion order -- the facts behind every gate of the function's epilogue
SparseOptimizer::~SparseOptimizer() {
    if (!rec || ba_edges.empty()) return;
    std::vector<double>& c = rec->named["edge_final_chi2"]; std::vector<double>& d = rec->named["edge_final_depth_positive"];
    for (const EdgeRef& r : ba_edges) {
        c.push_back(r.chi2());
        d.push_back(r.mono ? (r.mono->isDepthPositive() ? 1 : 0) : r.stereo ? (r.stereo->isDepthPositive() ? 1 : 0) : 1);
    }
}

// ---- PoseOptimization: one pose vertex (id 0), unary OnlyPose edges
struct PoseRun {
    SparseOptimizer* self; void* session; VertexSE3Expmap* pose; std::vector<uint8_t> lvl;
    void sync() {
        double s[7];
        self->be->pose_state(session, s, 7);
        SE3Quat T = pose->estimate();
        T.setRotation(Eigen::Quaterniond(s[3], s[0], s[1], s[2]));
        T.setTranslation(Eigen::Vector3d(s[4], s[5], s[6]));
        pose->setEstimate(T);
    }
    void real_errors() {
        for (size_t e = 0; e < self->pose_edges.size(); ++e) {
            if (lvl[e]) continue;
            if (self->pose_edges[e].mono) self->pose_edges[e].mono->computeError(); else self->pose_edges[e].stereo->computeError();
        }
    }
    static void h_compute(void* c) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.compute_errors(r->session); r->real_errors(); }
    static double h_chi2(void* c) { PoseRun* r = (PoseRun*)c; return r->self->be->pose_steps.chi2(r->session); }
    static void h_build(void* c) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.build(r->session); }
    static int h_layout(void* c, int32_t* d, int n) { PoseRun* r = (PoseRun*)c; return r->self->be->pose_steps.layout(r->session, d, n); }
    static int h_vectors(void* c, double* x, double* b, double* dg, int n) { PoseRun* r = (PoseRun*)c; return r->self->be->pose_steps.vectors(r->session, x, b, dg, n); }
    static void h_lambda(void* c, double l) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.set_lambda(r->session, l); }
    static int h_solve(void* c) { PoseRun* r = (PoseRun*)c; return r->self->be->pose_steps.solve(r->session); }
    static void h_update(void* c) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.update(r->session); r->sync(); }
    static void h_push(void* c) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.push(r->session); }
    static void h_pop(void* c) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.pop(r->session); r->sync(); }
    static void h_discard(void* c) { PoseRun* r = (PoseRun*)c; r->self->be->pose_steps.discard_top(r->session); }
};

int SparseOptimizer::optimize_pose(int iterations) {
    PoseRun run; run.self = this;
    run.pose = dynamic_cast<VertexSE3Expmap*>(vertex(0));
    const int n = (int)pose_edges.size();
    std::vector<float> xw(3 * n), obs(3 * n), w(n);
    std::vector<uint8_t> outl(std::max(n, 1), 0);
    run.lvl.assign(n, 0);
    adb_pose_problem P{};
    int n_active = 0, n_kernel = 0;
    for (int e = 0; e < n; ++e) {
        const PoseEdgeRef& r = pose_edges[e];
        if (r.mono) {
            for (int k = 0; k < 3; ++k) xw[3 * e + k] = (float)r.mono->Xw[k];
            obs[3 * e] = (float)r.mono->measurement()[0]; obs[3 * e + 1] = (float)r.mono->measurement()[1]; obs[3 * e + 2] = -1.f;
            w[e] = (float)r.mono->information()(0, 0);
            P.fx = r.mono->fx; P.fy = r.mono->fy; P.cx = r.mono->cx; P.cy = r.mono->cy;
        } else {
            for (int k = 0; k < 3; ++k) { xw[3 * e + k] = (float)r.stereo->Xw[k]; obs[3 * e + k] = (float)r.stereo->measurement()[k]; }
            w[e] = (float)r.stereo->information()(0, 0);
            P.fx = r.stereo->fx; P.fy = r.stereo->fy; P.cx = r.stereo->cx; P.cy = r.stereo->cy; P.bf = r.stereo->bf;
        }
        const int level = r.mono ? r.mono->level() : r.stereo->level();
        RobustKernel* k = r.mono ? r.mono->robustKernel() : r.stereo->robustKernel();
        run.lvl[e] = level != active_level;
        if (!run.lvl[e]) { ++n_active; n_kernel += k != nullptr; }
    }
    if (n_kernel != 0 && n_kernel != n_active) return -2;
    if (rec) {
        if (rec->round_active.empty()) { rec->xw = xw; rec->obs = obs; rec->inv_sigma2 = w; rec->cam[0] = P.fx; rec->cam[1] = P.fy; rec->cam[2] = P.cx; rec->cam[3] = P.cy; rec->cam[4] = P.bf; }
        rec->round_active.push_back(n_active);
    }
    if (n_active == 0) {         // initializeOptimization leaves no active vertex: optimize() returns -1 (core/sparse_optimizer.cpp:356-359)
        if (rec) { rec->round_iterations.push_back(-1); rec->round_robust.push_back(0); }
        return -1;
    }
    const SE3Quat& T = run.pose->estimate();
    double pq[4] = {T.rotation().x(), T.rotation().y(), T.rotation().z(), T.rotation().w()}, pt[3] = {T.translation()[0], T.translation()[1], T.translation()[2]};
    if (rec && rec->pose_q.empty()) { rec->pose_q.assign(pq, pq + 4); rec->pose_t.assign(pt, pt + 3); }
    int32_t fptr[2] = {0, n}, inl = 0;
    P.n_frames = 1; P.frame_ptr = fptr; P.pose_q = pq; P.pose_t = pt; P.xw = xw.data(); P.obs = obs.data(); P.inv_sigma2 = w.data();
    P.outlier = outl.data(); P.n_inliers = &inl;
    run.session = be->pose_open(&P, 0, run.lvl.data(), n_kernel != 0);
    ref_lm_hooks hk{&run, PoseRun::h_compute, PoseRun::h_chi2, PoseRun::h_build, PoseRun::h_layout, PoseRun::h_vectors, PoseRun::h_lambda, PoseRun::h_solve,
                    PoseRun::h_update, PoseRun::h_push, PoseRun::h_pop, PoseRun::h_discard};
    std::vector<double> rows(4 * 512);
    int n_rows = 0, n_eval = 0; double lam = 0, tau = 0;
    const int it = be->lm_optimize(&hk, iterations, 10, rows.data(), 512, &n_rows, &lam, &n_eval, &tau);
    run.sync();
    if (rec) {
        rec->rows.insert(rec->rows.end(), rows.begin(), rows.begin() + 4 * (size_t)std::min(n_rows, 512));
        rec->round_iterations.push_back(it); rec->round_robust.push_back(n_kernel != 0);
        rec->final_state.resize(7);
        be->pose_state(run.session, rec->final_state.data(), 7);
    }
    be->pose_close(run.session);
    return it;
}
}  // namespace g2o

namespace ORB_SLAM2 {
using namespace std;
class MapPoint;
class KeyFrame;
class MapHumanTrajectory;
// include/MapHumanPose.h:22-52, MapHumanTrajectory.h:19-24: the plain structs as the reference declares them
struct human_pose {
    int human_idx;
    std::vector<cv::KeyPoint> vHumanKeyPoints;
    std::vector<cv::KeyPoint> vHumanKeyPointsRight;
    std::vector<float> vKeysConfidence;
    std::vector<float> vKeysConfidenceRight;
};
struct MapHumanKey {
    int mnId;
    cv::Mat WorldPos;
    cv::Mat RelativePose;
    bool bIsBad = false;
    bool bIsLost = false;
    bool bOptimized = false;
    bool bIsFirstBad = false;
    bool bIsSecondBad = false;
};
struct HumanKeyPair {
    int idFirstKey;
    int idSecondKey;
    int idDistance;
    bool bIsBad = false;
    bool bOptimized = false;
};
struct Rigidbody {
    float mnDistance;
    int mnId;
    bool isOptimized = false;
    bool isBad = false;
};
class MapHumanPose {             // include/MapHumanPose.h:56-105: the members LocalBundleAdjustmentHumanTrajactory touches
public:
    long unsigned int mnId = 0;
    long unsigned int mnTrackId = -1;
    std::vector<HumanKeyPair> mvHumanKeyPair;
    std::vector<MapHumanKey*> mvHumanKeyPos;
    double mTimeStamp = 0;
    std::pair<KeyFrame*, size_t> mObservations;
    bool mbIsInKeyFrame = false;
    bool isLost = false;
    bool isEarsed = false;
    MapHumanTrajectory* mpRefHMT = nullptr;
    KeyFrame* mpRefKF = nullptr;
    std::vector<int> n_set;      // SetHumanKeyPos calls per key
    void SetHumanKeyPos(int nkey, cv::Mat HumanKeyPos, bool isOptimized) {   // src/MapHumanPose.cc: stores the position and the flag
        mvHumanKeyPos[nkey]->WorldPos = HumanKeyPos.clone(); mvHumanKeyPos[nkey]->bOptimized = isOptimized; ++n_set[nkey];
    }
};
class MapHumanTrajectory {       // include/MapHumanTrajectory.h:29-77
public:
    long unsigned int mnId = 0;
    long unsigned int mnTrackID = 0;
    cv::Mat mTMotion;
    std::vector<MapHumanPose*> mvHumanTrajactory;
    std::vector<Rigidbody> mvRigidBodys;
    long unsigned int mnBALocalForHM = 0;
    long unsigned int mnBAFixedForHM = 0;
    bool isOptimized = false;
    int mnHumanPoses = 0;
    int mnBadTrack = 0;
    std::vector<MapHumanPose*> GetMapHumanTrajectory() { return mvHumanTrajactory; }
};
class KeyFrame {                 // include/KeyFrame.h: the members LocalBundleAdjustment touches
public:
    long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAFixedForKF = 0, mnBAGlobalForKF = 0, mnBALocalForHM = 0, mnBAFixedForHM = 0;
    cv::Mat mTcwGBA;
    std::vector<human_pose> mvHumanPoses;              // include/KeyFrame.h:203-205
    std::vector<MapHumanPose*> mvpMapHumanPoses;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    std::vector<float> mvInvLevelSigma2;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    std::vector<KeyFrame*> covisible;
    std::vector<MapPoint*> matches;
    cv::Mat Tcw;
    std::vector<std::pair<int, MapPoint*>> erased;      // (global sequence number, point)
    static int erase_seq;
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return covisible; }
    std::vector<MapPoint*> GetMapPointMatches() { return matches; }
    bool isBad() { return false; }
    cv::Mat GetPose() { return Tcw.clone(); }
    void SetPose(const cv::Mat& T) { Tcw = T.clone(); }
    void EraseMapPointMatch(MapPoint* p) { erased.push_back(std::make_pair(erase_seq++, p)); }
};
class MapPoint {                 // include/MapPoint.h
public:
    long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAGlobalForKF = 0, mnBALocalForHM = 0;
    cv::Mat mPosGBA;
    std::map<KeyFrame*, size_t> observations;
    cv::Mat pos;
    int n_updates = 0;
    bool isBad() { return false; }
    std::map<KeyFrame*, size_t> GetObservations() { return observations; }
    cv::Mat GetWorldPos() { return pos.clone(); }
    void SetWorldPos(const cv::Mat& p) { pos = p.clone(); }
    void UpdateNormalAndDepth() { ++n_updates; }
    void EraseObservation(KeyFrame*) {}
    static std::mutex mGlobalMutex;
};
std::mutex MapPoint::mGlobalMutex;
class Frame {                    // include/Frame.h: the members PoseOptimization touches
public:
    cv::Mat mTcw;
    int N = 0;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<float> mvuRight;
    std::vector<bool> mvbOutlier;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvInvLevelSigma2;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); }
};
int KeyFrame::erase_seq = 0;
class Map {                      // include/Map.h:49-56, 89-100, 119
public:
    std::mutex mMutexMapUpdate;
    int mainskleton[5] = {1, 2, 5, 11, 8};
    int mimainskleton = 5;
    int mnbodyparts = 14;
    int body1[14] = {1, 1, 8, 2, 5, 2, 3, 5, 6, 8, 9, 11, 12, 1};
    int body2[14] = {2, 5, 11, 8, 11, 3, 4, 6, 7, 9, 10, 12, 13, 0};
    std::set<int> msetOptimizedTrackID;
    bool mbVOOnlyFlag = false;
    float thLongTrajectory = 3;
    std::map<size_t, MapHumanTrajectory*> trajectories;
    MapHumanTrajectory* GetMapHumanTrajectory(const size_t& idx) { auto it = trajectories.find(idx); return it == trajectories.end() ? nullptr : it->second; }
};
class Converter {                // include/Converter.h
public:
    static g2o::SE3Quat toSE3Quat(const cv::Mat& cvT);
    static cv::Mat toCvMat(const g2o::SE3Quat& SE3);
    static cv::Mat toCvMat(const Eigen::Matrix<double, 4, 4>& m);
    static cv::Mat toCvMat(const Eigen::Matrix<double, 3, 1>& m);
    static Eigen::Matrix<double, 3, 1> toVector3d(const cv::Mat& cvVector);
    static Eigen::Matrix<double, 4, 4> toEigenMatrix(const cv::Mat& cvT);
};
class Optimizer {
public:
    static void LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap);
};

const ref_lba_backend* g_backend = nullptr;     // reach the optimizer the literal function declares itself (see lba_scope::g2o below)
g2o::LbaRecord* g_record = nullptr;
}  // namespace ORB_SLAM2

namespace ORB_SLAM2 {
#include "_ref/lba_converter.inc"
}
// The function declares `g2o::SparseOptimizer optimizer;` itself.  It is compiled inside a scope whose `g2o` is the real namespace
// plus a SparseOptimizer that picks up the backend and the record in its constructor.
namespace ORB_SLAM2 {
namespace lba_scope {
namespace g2o {
using namespace ::g2o;
struct SparseOptimizer : ::g2o::SparseOptimizer { SparseOptimizer() { be = g_backend; rec = g_record; } };
}
using ORB_SLAM2::KeyFrame; using ORB_SLAM2::MapPoint; using ORB_SLAM2::Map; using ORB_SLAM2::Converter; using ORB_SLAM2::Frame;
using ORB_SLAM2::MapHumanPose; using ORB_SLAM2::MapHumanTrajectory; using ORB_SLAM2::MapHumanKey; using ORB_SLAM2::HumanKeyPair; using ORB_SLAM2::Rigidbody;
using ORB_SLAM2::human_pose;
class Optimizer {
public:
    static void LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap);
    static int PoseOptimization(Frame* pFrame);
    static void BundleAdjustment(const std::vector<KeyFrame*>& vpKFs, const std::vector<MapPoint*>& vpMP, int nIterations, bool* pbStopFlag,
                                 const unsigned long nLoopKF, const bool bRobust);
    static void LocalBundleAdjustmentHumanTrajactory(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, float SigmaStatic, float SigmaHuman, float SigmaRigidity,
                                                     float SigmaMotion, float thRanSacMotion, float thRanSacRigidity);
};
// The functions report progress on std::cerr.  Every shared library built by this toolchain carries its own static libstdc++, whose
// number-formatting facets do not survive several such copies in one process, so inside the function text `std::cerr` is a sink.
#define cerr lba_null_stream
#include "_ref/lba_snippets.inc"
#undef cerr
}  // namespace lba_scope
}  // namespace ORB_SLAM2

extern "C" {

struct ref_lba_io {
    // window: key-frame 0 is the current one (pKF); `covisible` lists the indices of its covisible key-frames in order
    int32_t n_kf; const int32_t* kf_id; const float* kf_tcw;        /* [n_kf][16] row-major 4 x 4 */
    int32_t n_covisible; const int32_t* covisible;
    float fx, fy, cx, cy, bf;
    int32_t n_levels; const float* inv_level_sigma2;
    int32_t n_mp; const int32_t* mp_id; const float* mp_pos;        /* [n_mp][3] */
    int32_t n_obs; const int32_t* obs_kf; const int32_t* obs_mp; const float* obs_uvr; const int32_t* obs_octave;   /* uvr: [n_obs][3], ur < 0 = mono */
    // results
    float* kf_tcw_out;                 /* [n_kf][16] */
    float* mp_pos_out;                 /* [n_mp][3] */
    int32_t* erased; int32_t erased_cap; int32_t n_erased;          /* (kf index, mp index) pairs in erase order */
    int32_t* mp_updates;               /* [n_mp] UpdateNormalAndDepth calls */
};

namespace {
struct Window {
    std::vector<ORB_SLAM2::KeyFrame> kfs;        // one block: addresses ascend with the index (std::map<KeyFrame*, size_t> iterates by address)
    std::vector<ORB_SLAM2::MapPoint> mps;
    explicit Window(const ref_lba_io* io) : kfs(io->n_kf), mps(io->n_mp) {
        using namespace ORB_SLAM2;
        for (int k = 0; k < io->n_kf; ++k) {
            KeyFrame& K = kfs[k];
            K.mnId = io->kf_id[k]; K.mnBALocalForKF = K.mnBAFixedForKF = K.mnBALocalForHM = K.mnBAFixedForHM = (unsigned long)-1;
            K.fx = io->fx; K.fy = io->fy; K.cx = io->cx; K.cy = io->cy; K.mbf = io->bf;
            K.mvInvLevelSigma2.assign(io->inv_level_sigma2, io->inv_level_sigma2 + io->n_levels);
            K.Tcw = cv::Mat(4, 4, CV_32F, io->kf_tcw + 16 * k);
        }
        for (int i = 0; i < io->n_covisible; ++i) kfs[0].covisible.push_back(&kfs[io->covisible[i]]);
        for (int m = 0; m < io->n_mp; ++m) { mps[m].mnId = io->mp_id[m]; mps[m].mnBALocalForKF = mps[m].mnBALocalForHM = (unsigned long)-1; mps[m].pos = cv::Mat(3, 1, CV_32F, io->mp_pos + 3 * m); }
        for (int o = 0; o < io->n_obs; ++o) {
            KeyFrame& K = kfs[io->obs_kf[o]];
            cv::KeyPoint kp; kp.pt.x = io->obs_uvr[3 * o]; kp.pt.y = io->obs_uvr[3 * o + 1]; kp.octave = io->obs_octave[o];
            const size_t idx = K.mvKeysUn.size();
            K.mvKeysUn.push_back(kp); K.mvuRight.push_back(io->obs_uvr[3 * o + 2]); K.matches.push_back(&mps[io->obs_mp[o]]);
            mps[io->obs_mp[o]].observations[&K] = idx;
        }
    }
    void results(ref_lba_io* io, bool gba) {
        for (int k = 0; k < io->n_kf; ++k) {
            const cv::Mat& T = gba ? kfs[k].mTcwGBA : kfs[k].Tcw;
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) io->kf_tcw_out[16 * k + 4 * i + j] = T.empty() ? 0.f : T.at<float>(i, j);
        }
        for (int m = 0; m < io->n_mp; ++m) {
            const cv::Mat& X = gba ? mps[m].mPosGBA : mps[m].pos;
            for (int i = 0; i < 3; ++i) io->mp_pos_out[3 * m + i] = X.empty() ? 0.f : X.at<float>(i);
            io->mp_updates[m] = gba ? (int)mps[m].mnBAGlobalForKF : mps[m].n_updates;
        }
        // the erase list in the function's own order (vToErase: mono edges first, then stereo, :672-700)
        std::vector<std::pair<int, std::pair<int, int>>> er;
        for (int k = 0; k < io->n_kf; ++k) for (auto& p : kfs[k].erased) er.push_back({p.first, {k, (int)(p.second - mps.data())}});
        std::sort(er.begin(), er.end());
        io->n_erased = (int)er.size();
        for (int i = 0; i < io->n_erased && i < io->erased_cap; ++i) { io->erased[2 * i] = er[i].second.first; io->erased[2 * i + 1] = er[i].second.second; }
    }
};
}  // namespace

// runs the literal Optimizer::LocalBundleAdjustment on the window; *rec_out receives a handle for ref_lba_record_* (free with ref_lba_record_free)
int ref_lba_run(const ref_lba_backend* be, ref_lba_io* io, void** rec_out) {
    using namespace ORB_SLAM2;
    Window W(io);
    g2o::LbaRecord* rec = new g2o::LbaRecord;
    g_backend = be; g_record = rec; KeyFrame::erase_seq = 0;
    Map map;
    lba_scope::Optimizer::LocalBundleAdjustment(&W.kfs[0], nullptr, &map);
    g_backend = nullptr; g_record = nullptr;
    W.results(io, false);
    *rec_out = rec;
    return 0;
}
// the literal Optimizer::BundleAdjustment (src/Optimizer.cc:60-230; what GlobalBundleAdjustemnt :52-58 calls) on all key-frames / points of
// the window: nLoopKF == 0 writes poses / positions back, otherwise into mTcwGBA / mPosGBA with mnBAGlobalForKF = nLoopKF (mp_updates then
// returns mnBAGlobalForKF per point)
int ref_gba_run(const ref_lba_backend* be, ref_lba_io* io, int n_iterations, int n_loop_kf, int robust, void** rec_out) {
    using namespace ORB_SLAM2;
    Window W(io);
    g2o::LbaRecord* rec = new g2o::LbaRecord;
    g_backend = be; g_record = rec; KeyFrame::erase_seq = 0;
    std::vector<KeyFrame*> vk; std::vector<MapPoint*> vm;
    for (auto& k : W.kfs) vk.push_back(&k);
    for (auto& m : W.mps) vm.push_back(&m);
    lba_scope::Optimizer::BundleAdjustment(vk, vm, n_iterations, nullptr, (unsigned long)n_loop_kf, robust != 0);
    g_backend = nullptr; g_record = nullptr;
    W.results(io, n_loop_kf != 0);
    *rec_out = rec;
    return 0;
}

struct ref_human_io {
    // trajectories: 14 bone lengths each (Rigidbody ids = rigid_id); human poses in trajectory order; 14 keys per pose
    int32_t n_traj; const int32_t* traj_id; const int32_t* traj_track_id; const int32_t* traj_n_poses;
    const int32_t* rigid_id; const float* rigid_dist;                    /* [n_traj][14] */
    int32_t n_hp; const int32_t* hp_id; const int32_t* hp_traj; const int32_t* hp_ref_kf; const double* hp_time;
    const int32_t* key_id; const float* key_pos; const float* key_uvr;  /* [n_hp][14] (, [3]) */
    int32_t n_current; const int32_t* current_hp;                       /* pKF->mvpMapHumanPoses: indices into the human poses */
    float sigma_static, sigma_human, sigma_rigidity, sigma_motion, th_motion, th_rigidity;
    // results
    float* key_pos_out;                /* [n_hp][14][3] */
    uint8_t* key_flags;                /* [n_hp][14][5]: bIsBad, bIsLost, bIsFirstBad, bIsSecondBad, bOptimized */
    uint8_t* pair_flags;               /* [n_hp][14][2]: bIsBad, bOptimized */
    int32_t* traj_out;                 /* [n_traj][2]: mnBadTrack, isOptimized */
    float* traj_motion;                /* [n_traj][16] mTMotion */
    int32_t n_optimized_tracks;        /* pMap->msetOptimizedTrackID.size() */
};
// the literal Optimizer::LocalBundleAdjustmentHumanTrajactory (src/Optimizer.cc:1496-2222) on the window + the human trajectories
int ref_hba_run(const ref_lba_backend* be, ref_lba_io* io, ref_human_io* hio, void** rec_out) {
    using namespace ORB_SLAM2;

    Window W(io);
    Map map;
    std::vector<MapHumanTrajectory> trajs(hio->n_traj);
    std::vector<MapHumanPose> hps(hio->n_hp);
    std::vector<MapHumanKey> keys((size_t)hio->n_hp * 14);
    for (int t = 0; t < hio->n_traj; ++t) {
        MapHumanTrajectory& T = trajs[t];
        T.mnId = hio->traj_id[t]; T.mnTrackID = hio->traj_track_id[t]; T.mnHumanPoses = hio->traj_n_poses[t]; T.mnBALocalForHM = T.mnBAFixedForHM = (unsigned long)-1;
        for (int b = 0; b < 14; ++b) { Rigidbody r; r.mnDistance = hio->rigid_dist[14 * t + b]; r.mnId = hio->rigid_id[14 * t + b]; T.mvRigidBodys.push_back(r); }
        map.trajectories[T.mnTrackID] = &T;
    }
    for (int h = 0; h < hio->n_hp; ++h) {
        MapHumanPose& H = hps[h];
        MapHumanTrajectory& T = trajs[hio->hp_traj[h]];
        KeyFrame& K = W.kfs[hio->hp_ref_kf[h]];
        H.mnId = hio->hp_id[h]; H.mnTrackId = T.mnTrackID; H.mTimeStamp = hio->hp_time[h]; H.mpRefHMT = &T; H.mpRefKF = &K; H.mbIsInKeyFrame = true;
        H.n_set.assign(14, 0);
        human_pose hp; hp.human_idx = (int)K.mvHumanPoses.size();
        for (int j = 0; j < 14; ++j) {
            MapHumanKey& key = keys[(size_t)14 * h + j];
            key.mnId = hio->key_id[14 * h + j];
            key.WorldPos = cv::Mat(3, 1, CV_32F, hio->key_pos + 3 * (14 * h + j));
            H.mvHumanKeyPos.push_back(&key);
            cv::KeyPoint l, r; l.pt.x = hio->key_uvr[3 * (14 * h + j)]; l.pt.y = hio->key_uvr[3 * (14 * h + j) + 1]; r.pt.x = hio->key_uvr[3 * (14 * h + j) + 2]; r.pt.y = l.pt.y;
            hp.vHumanKeyPoints.push_back(l); hp.vHumanKeyPointsRight.push_back(r);
        }
        for (int b = 0; b < 14; ++b) {      // the segment table of the map (include/Map.h:55-56)
            HumanKeyPair p; p.idFirstKey = hio->key_id[14 * h + map.body1[b]]; p.idSecondKey = hio->key_id[14 * h + map.body2[b]]; p.idDistance = T.mvRigidBodys[b].mnId;
            H.mvHumanKeyPair.push_back(p);
        }
        H.mObservations = std::make_pair(&K, K.mvHumanPoses.size());
        K.mvHumanPoses.push_back(hp);
        T.mvHumanTrajactory.push_back(&H);
    }
    for (int i = 0; i < hio->n_current; ++i) W.kfs[0].mvpMapHumanPoses.push_back(&hps[hio->current_hp[i]]);
    g2o::LbaRecord* rec = new g2o::LbaRecord;
    g_backend = be; g_record = rec; KeyFrame::erase_seq = 0;
    lba_scope::Optimizer::LocalBundleAdjustmentHumanTrajactory(&W.kfs[0], nullptr, &map, hio->sigma_static, hio->sigma_human, hio->sigma_rigidity, hio->sigma_motion,
                                                               hio->th_motion, hio->th_rigidity);
    g_backend = nullptr; g_record = nullptr;
    W.results(io, false);
    for (int h = 0; h < hio->n_hp; ++h)
        for (int j = 0; j < 14; ++j) {
            const MapHumanKey& key = keys[(size_t)14 * h + j];
            for (int c = 0; c < 3; ++c) hio->key_pos_out[3 * (14 * h + j) + c] = key.WorldPos.at<float>(c);
            uint8_t* f = hio->key_flags + 5 * (14 * h + j);
            f[0] = key.bIsBad; f[1] = key.bIsLost; f[2] = key.bIsFirstBad; f[3] = key.bIsSecondBad; f[4] = key.bOptimized;
            hio->pair_flags[2 * (14 * h + j)] = hps[h].mvHumanKeyPair[j].bIsBad; hio->pair_flags[2 * (14 * h + j) + 1] = hps[h].mvHumanKeyPair[j].bOptimized;
        }
    for (int t = 0; t < hio->n_traj; ++t) {
        hio->traj_out[2 * t] = trajs[t].mnBadTrack; hio->traj_out[2 * t + 1] = trajs[t].isOptimized;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) hio->traj_motion[16 * t + 4 * i + j] = trajs[t].mTMotion.empty() ? 0.f : trajs[t].mTMotion.at<float>(i, j);
    }
    hio->n_optimized_tracks = (int)map.msetOptimizedTrackID.size();
    *rec_out = rec;
    return 0;
}
// any recorded array by name, as doubles (see LbaRecord::named); returns its length, -1 if absent
int ref_lba_record_named(void* h, const char* name, double* out, int cap) {
    g2o::LbaRecord* r = (g2o::LbaRecord*)h;
    auto it = r->named.find(name);
    if (it == r->named.end()) return -1;
    for (size_t i = 0; i < it->second.size() && (int)i < cap; ++i) out[i] = it->second[i];
    return (int)it->second.size();
}

struct ref_pose_io {
    const float* tcw;                  /* [16] Frame::mTcw */
    float fx, fy, cx, cy, bf;
    int32_t n_levels; const float* inv_level_sigma2;
    int32_t n; const float* uvr; const int32_t* octave; const float* xw; const uint8_t* has_point;   /* key-points; xw / has_point: mvpMapPoints */
    uint8_t* outlier;                  /* [n] mvbOutlier after the call */
    float* tcw_out;                    /* [16] */
    int32_t n_inliers;                 /* the function's return value */
};
// the literal Optimizer::PoseOptimization (src/Optimizer.cc:232-429) on one frame
int ref_pose_run(const ref_lba_backend* be, ref_pose_io* io, void** rec_out) {
    using namespace ORB_SLAM2;
    Frame F;
    F.mTcw = cv::Mat(4, 4, CV_32F, io->tcw);
    F.N = io->n; F.fx = io->fx; F.fy = io->fy; F.cx = io->cx; F.cy = io->cy; F.mbf = io->bf;
    F.mvInvLevelSigma2.assign(io->inv_level_sigma2, io->inv_level_sigma2 + io->n_levels);
    std::vector<MapPoint> mps(io->n);
    for (int i = 0; i < io->n; ++i) {
        cv::KeyPoint kp; kp.pt.x = io->uvr[3 * i]; kp.pt.y = io->uvr[3 * i + 1]; kp.octave = io->octave[i];
        F.mvKeysUn.push_back(kp); F.mvuRight.push_back(io->uvr[3 * i + 2]); F.mvbOutlier.push_back(true);
        mps[i].pos = cv::Mat(3, 1, CV_32F, io->xw + 3 * i);
        F.mvpMapPoints.push_back(io->has_point[i] ? &mps[i] : nullptr);
    }
    g2o::LbaRecord* rec = new g2o::LbaRecord;
    g_backend = be; g_record = rec;
    io->n_inliers = lba_scope::Optimizer::PoseOptimization(&F);
    g_backend = nullptr; g_record = nullptr;
    for (int i = 0; i < io->n; ++i) io->outlier[i] = F.mvbOutlier[i] ? 1 : 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) io->tcw_out[4 * i + j] = F.mTcw.at<float>(i, j);
    *rec_out = rec;
    return 0;
}
// which = 0 xw, 1 obs, 2 inv_sigma2
int ref_lba_record_f32(void* h, int which, float* out, int cap) {
    g2o::LbaRecord* r = (g2o::LbaRecord*)h;
    const std::vector<float>* v[3] = {&r->xw, &r->obs, &r->inv_sigma2};
    const std::vector<float>& a = *v[which];
    for (size_t i = 0; i < a.size() && (int)i < cap; ++i) out[i] = a[i];
    return (int)a.size();
}

// record accessors: which = 0 pose_q, 1 pose_t, 2 points, 3 edge_obs, 4 edge_info, 5 rows, 6 final_state, 7 cam + huber (7 doubles)
int ref_lba_record_f64(void* h, int which, double* out, int cap) {
    g2o::LbaRecord* r = (g2o::LbaRecord*)h;
    std::vector<double> camv(r->cam, r->cam + 5); camv.push_back(r->huber_mono); camv.push_back(r->huber_stereo);
    const std::vector<double>* v[8] = {&r->pose_q, &r->pose_t, &r->points, &r->edge_obs, &r->edge_info, &r->rows, &r->final_state, &camv};
    const std::vector<double>& a = *v[which];
    for (size_t i = 0; i < a.size() && (int)i < cap; ++i) out[i] = a[i];
    return (int)a.size();
}
// which = 0 edge_pose, 1 edge_point, 2 pose_id, 3 point_id, 4 round_iterations, 5 round_robust, 6 pose_fixed, 7 round_active
int ref_lba_record_i32(void* h, int which, int32_t* out, int cap) {
    g2o::LbaRecord* r = (g2o::LbaRecord*)h;
    std::vector<int32_t> fx(r->pose_fixed.begin(), r->pose_fixed.end());
    const std::vector<int32_t>* v[8] = {&r->edge_pose, &r->edge_point, &r->pose_id, &r->point_id, &r->round_iterations, &r->round_robust, &fx, &r->round_active};
    const std::vector<int32_t>& a = *v[which];
    for (size_t i = 0; i < a.size() && (int)i < cap; ++i) out[i] = a[i];
    return (int)a.size();
}
void ref_lba_record_free(void* h) { delete (g2o::LbaRecord*)h; }

}  // extern "C"
