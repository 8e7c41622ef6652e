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
};
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
struct BlockSolver_6_3 {
    typedef Eigen::Matrix<double, 6, 6> PoseMatrixType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
    explicit BlockSolver_6_3(LinearSolverType*) {}
};
struct OptimizationAlgorithmLevenberg { explicit OptimizationAlgorithmLevenberg(BlockSolver_6_3*) {} };

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
};

class SparseOptimizer {
public:
    const ref_lba_backend* be = nullptr;
    LbaRecord* rec = nullptr;
    std::map<int, HyperGraph::Vertex*> vmap;
    struct EdgeRef { EdgeSE3ProjectXYZ* mono; EdgeStereoSE3ProjectXYZ* stereo; };
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
    bool addEdge(EdgeSE3ProjectXYZ* e) { ba_edges.push_back(EdgeRef{e, nullptr}); return true; }
    bool addEdge(EdgeStereoSE3ProjectXYZ* e) { ba_edges.push_back(EdgeRef{nullptr, e}); return true; }
    bool addEdge(EdgeSE3ProjectXYZOnlyPose* e) { pose_edges.push_back(PoseEdgeRef{e, nullptr}); return true; }
    bool addEdge(EdgeStereoSE3ProjectXYZOnlyPose* e) { pose_edges.push_back(PoseEdgeRef{nullptr, e}); return true; }
    HyperGraph::Vertex* vertex(int id) { auto it = vmap.find(id); return it == vmap.end() ? nullptr : it->second; }
    bool initializeOptimization(int level = 0) { active_level = level; return true; }
    int optimize(int iterations) { return pose_edges.empty() ? optimize_ba(iterations) : optimize_pose(iterations); }
    int optimize_ba(int iterations);
    int optimize_pose(int iterations);

    // ---- session plumbing
    struct Run {
        SparseOptimizer* self; void* session;
        std::vector<VertexSE3Expmap*> poses; std::vector<VertexSBAPointXYZ*> points;
        std::vector<uint8_t> lvl;
        void sync() {          // the session's estimates -> the real vertices
            const int n = self->be->state(session, nullptr, 0);
            std::vector<double> s(n);
            self->be->state(session, s.data(), n);
            const size_t np = poses.size(), nx = points.size();
            for (size_t i = 0; i < np; ++i) {
                SE3Quat T = poses[i]->estimate();
                T.setRotation(Eigen::Quaterniond(s[4 * i + 3], s[4 * i], s[4 * i + 1], s[4 * i + 2]));
                T.setTranslation(Eigen::Vector3d(s[4 * np + 3 * i], s[4 * np + 3 * i + 1], s[4 * np + 3 * i + 2]));
                poses[i]->setEstimate(T);
            }
            for (size_t l = 0; l < nx; ++l) points[l]->setEstimate(Eigen::Vector3d(s[7 * np + 3 * l], s[7 * np + 3 * l + 1], s[7 * np + 3 * l + 2]));
        }
        void real_errors() {   // g2o's computeActiveErrors on the real edges (core/sparse_optimizer.cpp:70-97): active edges only
            for (size_t e = 0; e < self->ba_edges.size(); ++e) {
                if (lvl[e]) continue;
                if (self->ba_edges[e].mono) self->ba_edges[e].mono->computeError(); else self->ba_edges[e].stereo->computeError();
            }
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
    std::map<HyperGraph::Vertex*, int> pose_index, point_index;
    for (auto& kv : vmap) {      // g2o orders the vertices of the index mapping by id (std::map: ascending)
        if (VertexSE3Expmap* p = dynamic_cast<VertexSE3Expmap*>(kv.second)) { pose_index[p] = (int)run.poses.size(); run.poses.push_back(p); }
        else if (VertexSBAPointXYZ* x = dynamic_cast<VertexSBAPointXYZ*>(kv.second)) { point_index[x] = (int)run.points.size(); run.points.push_back(x); }
    }
    const int np = (int)run.poses.size(), nx = (int)run.points.size(), ne = (int)edges.size();
    std::vector<double> pq(4 * np), pt(3 * np), X(3 * nx), obs(3 * ne), info(ne);
    std::vector<uint8_t> fixed(np);
    std::vector<int32_t> ep(ne), ex(ne), pid(np), xid(nx);
    for (int i = 0; i < np; ++i) {
        const SE3Quat& T = run.poses[i]->estimate();
        pq[4 * i] = T.rotation().x(); pq[4 * i + 1] = T.rotation().y(); pq[4 * i + 2] = T.rotation().z(); pq[4 * i + 3] = T.rotation().w();
        for (int k = 0; k < 3; ++k) pt[3 * i + k] = T.translation()[k];
        fixed[i] = run.poses[i]->fixed(); pid[i] = run.poses[i]->id();
    }
    for (int l = 0; l < nx; ++l) { for (int k = 0; k < 3; ++k) X[3 * l + k] = run.points[l]->estimate()[k]; xid[l] = run.points[l]->id(); }
    adb_ba_problem P{};
    adb_ba_options O{};
    be->default_options(&O);
    int n_kernel = 0, n_active = 0;
    run.lvl.assign(ne, 0);
    for (int e = 0; e < ne; ++e) {
        const EdgeRef& r = edges[e];
        HyperGraph::Vertex* v0 = r.mono ? r.mono->_vertices[0] : r.stereo->_vertices[0];
        HyperGraph::Vertex* v1 = r.mono ? r.mono->_vertices[1] : r.stereo->_vertices[1];
        ex[e] = point_index.at(v0); ep[e] = pose_index.at(v1);
        if (r.mono) {
            obs[3 * e] = r.mono->measurement()[0]; obs[3 * e + 1] = r.mono->measurement()[1]; obs[3 * e + 2] = -1.0;
            info[e] = r.mono->information()(0, 0);
            P.fx = r.mono->fx; P.fy = r.mono->fy; P.cx = r.mono->cx; P.cy = r.mono->cy;
        } else {
            for (int k = 0; k < 3; ++k) obs[3 * e + k] = r.stereo->measurement()[k];
            info[e] = r.stereo->information()(0, 0);
            P.fx = r.stereo->fx; P.fy = r.stereo->fy; P.cx = r.stereo->cx; P.cy = r.stereo->cy; P.bf = r.stereo->bf;
        }
        const int level = r.mono ? r.mono->level() : r.stereo->level();
        RobustKernel* k = r.mono ? r.mono->robustKernel() : r.stereo->robustKernel();
        run.lvl[e] = level != active_level;
        if (!run.lvl[e]) { ++n_active; if (k) { ++n_kernel; (r.mono ? O.huber_mono : O.huber_stereo) = k->delta(); } }
    }
    if (n_kernel != 0 && n_kernel != n_active) return -2;     // mixed kernels: only with bad map points, which the stand-ins never report
    const int robust = n_kernel != 0;
    P.n_poses = np; P.n_points = nx; P.n_edges = ne;
    P.pose_q = pq.data(); P.pose_t = pt.data(); P.pose_fixed = fixed.data(); P.points = X.data();
    P.edge_pose = ep.data(); P.edge_point = ex.data(); P.edge_obs = obs.data(); P.edge_info = info.data();
    if (rec && rec->round_iterations.empty()) {
        rec->pose_q = pq; rec->pose_t = pt; rec->points = X; rec->edge_obs = obs; rec->edge_info = info; rec->pose_fixed = fixed;
        rec->edge_pose = ep; rec->edge_point = ex; rec->pose_id = pid; rec->point_id = xid;
        rec->cam[0] = P.fx; rec->cam[1] = P.fy; rec->cam[2] = P.cx; rec->cam[3] = P.cy; rec->cam[4] = P.bf;
        rec->huber_mono = O.huber_mono; rec->huber_stereo = O.huber_stereo;
    }
    run.session = be->open(&P, &O, robust);
    be->set_levels(run.session, run.lvl.data());
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
class KeyFrame {                 // include/KeyFrame.h: the members LocalBundleAdjustment touches
public:
    long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAFixedForKF = 0, mnBAGlobalForKF = 0;
    cv::Mat mTcwGBA;
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
    long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAGlobalForKF = 0;
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
class Map { public: std::mutex mMutexMapUpdate; };
class Converter {                // include/Converter.h
public:
    static g2o::SE3Quat toSE3Quat(const cv::Mat& cvT);
    static cv::Mat toCvMat(const g2o::SE3Quat& SE3);
    static cv::Mat toCvMat(const Eigen::Matrix<double, 4, 4>& m);
    static cv::Mat toCvMat(const Eigen::Matrix<double, 3, 1>& m);
    static Eigen::Matrix<double, 3, 1> toVector3d(const cv::Mat& cvVector);
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
class Optimizer {
public:
    static void LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap);
    static int PoseOptimization(Frame* pFrame);
    static void BundleAdjustment(const std::vector<KeyFrame*>& vpKFs, const std::vector<MapPoint*>& vpMP, int nIterations, bool* pbStopFlag,
                                 const unsigned long nLoopKF, const bool bRobust);
};
#include "_ref/lba_snippets.inc"
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
            K.mnId = io->kf_id[k]; K.mnBALocalForKF = K.mnBAFixedForKF = (unsigned long)-1;
            K.fx = io->fx; K.fy = io->fy; K.cx = io->cx; K.cy = io->cy; K.mbf = io->bf;
            K.mvInvLevelSigma2.assign(io->inv_level_sigma2, io->inv_level_sigma2 + io->n_levels);
            K.Tcw = cv::Mat(4, 4, CV_32F, io->kf_tcw + 16 * k);
        }
        for (int i = 0; i < io->n_covisible; ++i) kfs[0].covisible.push_back(&kfs[io->covisible[i]]);
        for (int m = 0; m < io->n_mp; ++m) { mps[m].mnId = io->mp_id[m]; mps[m].mnBALocalForKF = (unsigned long)-1; mps[m].pos = cv::Mat(3, 1, CV_32F, io->mp_pos + 3 * m); }
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
