// oracle/ref_lm.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Built only where /root/reference exists (make -C oracle ref ->
// oracle/_ref/libref_lm.so, git-ignored).
//
// The reference's OWN Levenberg-Marquardt control and per-edge quadratic form, compiled from /root/reference:
//   Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp   the WHOLE file, unmodified (#include below): constructor defaults,
//                                                                  solve (:61-161), computeLambdaInit (:163-177), computeScale (:179-186)
//   SparseOptimizer::optimize             core/sparse_optimizer.cpp:354-419     the iteration loop around solve()
//   RobustKernelHuber::setDelta/robustify core/robust_kernel_impl.cpp:65-92     (dsqr is a FLOAT member: core/robust_kernel_impl.h:84)
//   BaseEdge::chi2 / robustInformation    core/base_edge.h:58-61, 96-102
//   BaseBinaryEdge::constructQuadraticForm  core/base_binary_edge.hpp:54-117
//   BaseUnaryEdge::constructQuadraticForm   core/base_unary_edge.hpp:40-69
//   BaseMultiEdge::constructQuadraticForm / computeQuadraticForm   core/base_multi_edge.hpp:35-49, 170-222   (rigidity / motion edges)
// The function texts are taken out of the reference tree at build time (oracle/extract_ref_fn.py -> oracle/_ref/lm_*.inc) and
// compiled between stand-in declarations of the classes they are members of (same member names and types as the reference's
// headers, which themselves need Eigen proper and all of g2o core).  What the stand-ins do NOT restate is the arithmetic
// underneath: computeActiveErrors / buildSystem / solve / update / push / pop are function pointers handed in by the test, which
// point at the oracle's own steps (oracle/ba_oracle.cpp, ba_oracle_lm_*).  So a run of ref_lm_optimize is the reference's control
// flow over the oracle's arithmetic, and any difference from the oracle's own loop (Solver::optimize) is a control-flow difference.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#include "ref_shim/eigen_shim.h"

// the real headers of these are skipped (they pull in Eigen proper / the whole graph machinery); the stand-ins follow
#define G2O_SOLVER_LEVENBERG_H
#define G2O_GRAPH_OPTIMIZER_CHOL_H_
#define G2O_SOLVER_H
#define G2O_BATCH_STATS_H_
#define G2O_TIMEUTIL_H
#include "Thirdparty/g2o/g2o/stuff/macros.h"   // g2o_isfinite, FIXED: the reference's own (no dependencies)

extern "C" {
struct ref_lm_hooks {          // the arithmetic under the control flow (ctx = the oracle's LM session)
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
}

namespace g2o {
using namespace std;

inline double get_monotonic_time() { return 0.0; }                       // stuff/timeutil.h
struct G2OBatchStatistics {                                              // core/batch_stats.h: never active here
    int iteration = 0, numEdges = 0, numVertices = 0, levenbergIterations = 0;
    double chi2 = 0, timeResiduals = 0, timeQuadraticForm = 0, timeLinearSolution = 0, timeUpdate = 0, timeIteration = 0;
    static G2OBatchStatistics* globalStats() { return nullptr; }
    static void setGlobalStats(G2OBatchStatistics*) {}
};
typedef std::vector<G2OBatchStatistics> BatchStatisticsContainer;

template <typename T> class Property {                                   // stuff/property.h
    T v_;
public:
    Property(const std::string&, const T& d) : v_(d) {}
    const T& value() const { return v_; }
    void setValue(const T& v) { v_ = v; }
};
struct PropertyMap {
    template <typename P, typename T> P* makeProperty(const std::string& name, const T& d) { return new P(name, d); }
};

struct LmLog {     // what the stand-ins observe of a run: one row per trial = (lambda, chi2 before, chi2 after, accepted)
    std::vector<double> rows;
    double lambda = 0, before = 0, after = 0;
};

struct QVertexBase { virtual ~QVertexBase() {} };
class OptimizableGraph {
public:
    struct Vertex : QVertexBase {
        int dim = 0; const double* diag = nullptr;
        int dimension() const { return dim; }
        double hessian(int i, int j) const { return i == j ? diag[i] : 0.0; }
        // the multi-edge writes through raw pointers (core/optimizable_graph.h: hessianData / bData)
        bool fixed_ = false; std::vector<double> H, b;
        bool fixed() const { return fixed_; }
        double* hessianData() { return H.data(); }
        double* bData() { return b.data(); }
    };
    typedef std::vector<Vertex*> VertexContainer;
};

class OptimizationAlgorithm;
class SparseOptimizer {                    // core/sparse_optimizer.h: the members optimize() and the Levenberg solver touch
public:
    ref_lm_hooks hk; LmLog* log; class Solver* solver_ = nullptr;
    OptimizableGraph::VertexContainer _ivMap, _activeVertices;
    std::vector<int> _activeEdges;
    OptimizationAlgorithm* _algorithm = nullptr;
    BatchStatisticsContainer _batchStatistics;
    bool _computeBatchStatistics = false;
    std::vector<OptimizableGraph::Vertex> verts; std::vector<double> diag;
    int n_eval = 0;
    bool iter_start = true;

    int optimize(int iterations, bool online = false);
    void computeActiveErrors() { hk.compute_errors(hk.ctx); ++n_eval; if (iter_start) { log->before = hk.chi2(hk.ctx); iter_start = false; } }
    double activeRobustChi2() const { return hk.chi2(hk.ctx); }
    void push() { hk.push(hk.ctx); }
    void pop() { hk.pop(hk.ctx); row(0); }                                  // rejected: chi2 "before" stays (errors are left stale, like :142)
    void discardTop() { hk.discard_top(hk.ctx); row(1); log->before = hk.chi2(hk.ctx); }
    void update(const double*);           // defined after Solver
    const OptimizableGraph::VertexContainer& indexMapping() const { return _ivMap; }
    bool terminate() const { return false; }
    bool verbose() const { return false; }
    void preIteration(int) { iter_start = true; }
    void postIteration(int) {}
    void row(int accepted) { log->rows.insert(log->rows.end(), {log->lambda, log->before, hk.chi2(hk.ctx), (double)accepted}); }
    void refresh_index_mapping() {         // after buildSystem: vertex dimensions and Hessian diagonals for computeLambdaInit
        std::vector<int32_t> dims(hk.layout(hk.ctx, nullptr, 0));
        hk.layout(hk.ctx, dims.data(), (int)dims.size());
        diag.assign(hk.vectors(hk.ctx, nullptr, nullptr, nullptr, 0), 0.0);
        hk.vectors(hk.ctx, nullptr, nullptr, diag.data(), (int)diag.size());
        verts.resize(dims.size()); _ivMap.resize(dims.size());
        size_t o = 0;
        for (size_t k = 0; k < dims.size(); ++k) { verts[k].dim = dims[k]; verts[k].diag = diag.data() + o; o += dims[k]; _ivMap[k] = &verts[k]; }
    }
};

class Solver {                             // core/solver.h
public:
    SparseOptimizer* opt = nullptr;
    std::vector<double> xv, bv;
    bool solved_ok = true;
    bool buildStructure(bool = false) { return true; }
    bool buildSystem() { opt->hk.build(opt->hk.ctx); opt->refresh_index_mapping(); fetch(); return true; }
    bool setLambda(double lambda, bool = false) { opt->log->lambda = lambda; opt->hk.set_lambda(opt->hk.ctx, lambda); return true; }
    void restoreDiagonal() {}              // the oracle's solve works on a copy of the diagonal
    bool solve() { solved_ok = opt->hk.solve(opt->hk.ctx) != 0; fetch(); return solved_ok; }
    double* x() { return xv.data(); }
    double* b() { return bv.data(); }
    size_t vectorSize() const { return xv.size(); }
    bool schur() const { return true; }
    SparseOptimizer* optimizer() const { return opt; }
    void fetch() {
        const int n = opt->hk.vectors(opt->hk.ctx, nullptr, nullptr, nullptr, 0);
        xv.assign(n, 0.0); bv.assign(n, 0.0);
        opt->hk.vectors(opt->hk.ctx, xv.data(), bv.data(), nullptr, n);
    }
};

// the reference applies x even after a failed solve and then restores the state (pop); the oracle's failed solve leaves no x to apply
inline void SparseOptimizer::update(const double*) { if (solver_->solved_ok) hk.update(hk.ctx); }

class OptimizationAlgorithm {              // core/optimization_algorithm.h
public:
    enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
    virtual ~OptimizationAlgorithm() {}
    virtual bool init(bool online = false) = 0;
    virtual SolverResult solve(int iteration, bool online = false) = 0;
    virtual void printVerbose(std::ostream&) const {}
    SparseOptimizer* _optimizer = nullptr;
    PropertyMap _properties;
};
class OptimizationAlgorithmWithHessian : public OptimizationAlgorithm {   // core/optimization_algorithm_with_hessian.h
public:
    explicit OptimizationAlgorithmWithHessian(Solver* solver) : _solver(solver) {}
    virtual bool init(bool = false) { return true; }
    Solver* _solver;
};

// core/optimization_algorithm_levenberg.h:37-87, declarations as in the reference
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithmWithHessian {
public:
    explicit OptimizationAlgorithmLevenberg(Solver* solver);
    virtual ~OptimizationAlgorithmLevenberg();
    virtual SolverResult solve(int iteration, bool online = false);
    virtual void printVerbose(std::ostream& os) const;
    double currentLambda() const { return _currentLambda; }
    void setMaxTrialsAfterFailure(int max_trials);
    int maxTrialsAfterFailure() const { return _maxTrialsAfterFailure->value(); }
    double userLambdaInit() { return _userLambdaInit->value(); }
    void setUserLambdaInit(double lambda);
    int levenbergIteration() { return _levenbergIterations; }
    double tau() const { return _tau; }
protected:
    Property<int>* _maxTrialsAfterFailure;
    Property<double>* _userLambdaInit;
    double _currentLambda;
    double _tau;
    double _goodStepLowerScale;
    double _goodStepUpperScale;
    double _ni;
    int _levenbergIterations;
    int _nBad;
    double computeLambdaInit() const;
    double computeScale() const;
};

// ---- robust kernel (core/robust_kernel.h:50-70, core/robust_kernel_impl.h:76-85)
class RobustKernel {
public:
    virtual ~RobustKernel() {}
    virtual void robustify(double squaredError, Eigen::Vector3d& rho) const = 0;
    virtual void setDelta(double delta) { _delta = delta; }
    double delta() const { return _delta; }
protected:
    double _delta = 1.0;
};
class RobustKernelHuber : public RobustKernel {
public:
    virtual void setDelta(double delta);
    virtual void robustify(double e2, Eigen::Vector3d& rho) const;
private:
    float dsqr;
};

// ---- edges: the members constructQuadraticForm touches (core/base_edge.h, base_binary_edge.h, base_unary_edge.h)
template <int D> struct QVertex : QVertexBase {
    static const int Dimension = D;
    bool fixed_ = false;
    Eigen::Matrix<double, D, D> A_;
    Eigen::Matrix<double, D, 1> b_;
    bool fixed() const { return fixed_; }
    Eigen::Matrix<double, D, D>& A() { return A_; }
    Eigen::Matrix<double, D, 1>& b() { return b_; }
};
template <int D, typename E> class BaseEdge {
public:
    typedef E Measurement;
    typedef Eigen::Matrix<double, D, 1> ErrorVector;
    typedef Eigen::Matrix<double, D, D> InformationType;
    virtual ~BaseEdge() {}
    const InformationType& information() const { return _information; }
    RobustKernel* robustKernel() const { return _robustKernel; }
    std::vector<QVertexBase*> _vertices;
    RobustKernel* _robustKernel = nullptr;
#include "_ref/lm_edge_inline.inc"          // chi2(), robustInformation(): the reference's inline bodies
    Measurement _measurement;
    InformationType _information;
    ErrorVector _error;
};
template <int D, typename E, typename VertexXiType, typename VertexXjType> class BaseBinaryEdge : public BaseEdge<D, E> {
public:
    typedef typename BaseEdge<D, E>::InformationType InformationType;
    typedef Eigen::Matrix<double, D, VertexXiType::Dimension> JacobianXiOplusType;
    typedef Eigen::Matrix<double, D, VertexXjType::Dimension> JacobianXjOplusType;
    typedef Eigen::Matrix<double, VertexXiType::Dimension, VertexXjType::Dimension> HessianBlockType;
    typedef Eigen::Matrix<double, VertexXjType::Dimension, VertexXiType::Dimension> HessianBlockTransposedType;
    using BaseEdge<D, E>::_vertices; using BaseEdge<D, E>::_information; using BaseEdge<D, E>::_error;
    const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
    const JacobianXjOplusType& jacobianOplusXj() const { return _jacobianOplusXj; }
    void constructQuadraticForm();
    bool _hessianRowMajor = false;
    HessianBlockType _hessian;
    HessianBlockTransposedType _hessianTransposed;
    JacobianXiOplusType _jacobianOplusXi;
    JacobianXjOplusType _jacobianOplusXj;
};
template <int D, typename E, typename VertexXiType> class BaseUnaryEdge : public BaseEdge<D, E> {
public:
    typedef typename BaseEdge<D, E>::InformationType InformationType;
    typedef Eigen::Matrix<double, D, VertexXiType::Dimension> JacobianXiOplusType;
    using BaseEdge<D, E>::_vertices; using BaseEdge<D, E>::_information; using BaseEdge<D, E>::_error;
    const JacobianXiOplusType& jacobianOplusXi() const { return _jacobianOplusXi; }
    void constructQuadraticForm();
    JacobianXiOplusType _jacobianOplusXi;
};

template <int D, typename E> class BaseMultiEdge : public BaseEdge<D, E> {      // core/base_multi_edge.h:53-108
public:
    typedef typename BaseEdge<D, E>::InformationType InformationType;
    typedef typename BaseEdge<D, E>::ErrorVector ErrorVector;
    typedef Eigen::MatrixXd JacobianType;
    typedef Eigen::Map<Eigen::MatrixXd> HessianBlockType;
    struct HessianHelper { Eigen::Map<Eigen::MatrixXd> matrix; bool transposed; HessianHelper() : matrix(nullptr, 0, 0), transposed(false) {} };
    using BaseEdge<D, E>::_vertices; using BaseEdge<D, E>::_information; using BaseEdge<D, E>::_error;
    void constructQuadraticForm();
    void computeQuadraticForm(const InformationType& omega, const ErrorVector& weightedError);
    std::vector<HessianHelper> _hessian;
    std::vector<JacobianType> _jacobianOplus;
};

using namespace Eigen;
#include "_ref/lm_huber.inc"
namespace internal { inline int computeUpperTriangleIndex(int i, int j) { int elemsUpToCol = ((j - 1) * j) / 2; return elemsUpToCol + i; } }   // base_multi_edge.hpp:27-33
template <int D, typename E>
#include "_ref/lm_multi_cqf.inc"
template <int D, typename E>
#include "_ref/lm_multi_compute.inc"
template <int D, typename E, typename VertexXiType, typename VertexXjType>
#include "_ref/lm_binary_cqf.inc"
template <int D, typename E, typename VertexXiType>
#include "_ref/lm_unary_cqf.inc"
#include "_ref/lm_optimize.inc"
}  // namespace g2o

#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp"

namespace {
template <int D> void run_binary(const double* Ji, const double* Jj, const double* er, double w0, double delta, int robust, int fixed_point, int fixed_pose,
                                 double* hl, double* gl, double* hp, double* gp, double* w63) {
    using namespace g2o;
    QVertex<3> point; QVertex<6> pose;
    point.fixed_ = fixed_point != 0; pose.fixed_ = fixed_pose != 0;
    BaseBinaryEdge<D, Eigen::Matrix<double, D, 1>, QVertex<3>, QVertex<6>> e;
    e._vertices = {&point, &pose};
    RobustKernelHuber rk;
    if (robust) { rk.setDelta(delta); e._robustKernel = &rk; }
    for (int k = 0; k < D; ++k) {
        e._error(k) = er[k];
        for (int j = 0; j < 3; ++j) e._jacobianOplusXi(k, j) = Ji[k * 3 + j];
        for (int j = 0; j < 6; ++j) e._jacobianOplusXj(k, j) = Jj[k * 6 + j];
        e._information(k, k) = w0;
    }
    e.constructQuadraticForm();
    for (int i = 0; i < 3; ++i) { gl[i] = point.b_(i); for (int j = 0; j < 3; ++j) hl[i * 3 + j] = point.A_(i, j); }
    for (int i = 0; i < 6; ++i) { gp[i] = pose.b_(i); for (int j = 0; j < 6; ++j) hp[i * 6 + j] = pose.A_(i, j); }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 3; ++j) w63[i * 3 + j] = e._hessian(j, i);   // g2o keeps point x pose
}
template <int D> void run_unary(const double* J, const double* er, double w0, double delta, int robust, double* h, double* g) {
    using namespace g2o;
    QVertex<6> pose;
    BaseUnaryEdge<D, Eigen::Matrix<double, D, 1>, QVertex<6>> e;
    e._vertices = {&pose};
    RobustKernelHuber rk;
    if (robust) { rk.setDelta(delta); e._robustKernel = &rk; }
    for (int k = 0; k < D; ++k) {
        e._error(k) = er[k];
        for (int j = 0; j < 6; ++j) e._jacobianOplusXi(k, j) = J[k * 6 + j];
        e._information(k, k) = w0;
    }
    e.constructQuadraticForm();
    for (int i = 0; i < 6; ++i) { g[i] = pose.b_(i); for (int j = 0; j < 6; ++j) h[i * 6 + j] = pose.A_(i, j); }
}
template <int D> void run_multi(int nv, const int32_t* dims, const double* const* Js, const double* er, double w0, double delta, int robust, const uint8_t* fixed,
                                double* H, double* b) {
    using namespace g2o;
    std::vector<OptimizableGraph::Vertex> vs(nv);
    BaseMultiEdge<D, Eigen::Matrix<double, D, 1>> e;
    int n = 0; std::vector<int> off(nv);
    for (int v = 0; v < nv; ++v) { off[v] = n; n += dims[v]; }
    for (int v = 0; v < nv; ++v) {
        vs[v].dim = dims[v]; vs[v].fixed_ = fixed && fixed[v]; vs[v].H.assign((size_t)dims[v] * dims[v], 0.0); vs[v].b.assign(dims[v], 0.0);
        e._vertices.push_back(&vs[v]);
        Eigen::MatrixXd J(D, dims[v]);
        for (int k = 0; k < D; ++k) for (int j = 0; j < dims[v]; ++j) J(k, j) = Js[v][k * dims[v] + j];
        e._jacobianOplus.push_back(J);
    }
    // off-diagonal blocks: upper triangle, column-major maps into scratch (g2o maps them into the solver's block matrix: mapHessianMemory)
    std::vector<std::vector<double>> blocks((size_t)nv * (nv - 1) / 2);
    e._hessian.resize(blocks.size());
    for (int j = 1; j < nv; ++j)
        for (int i = 0; i < j; ++i) {
            const int idx = internal::computeUpperTriangleIndex(i, j);
            blocks[idx].assign((size_t)dims[i] * dims[j], 0.0);
            e._hessian[idx].matrix = Eigen::Map<Eigen::MatrixXd>(blocks[idx].data(), dims[i], dims[j]);
        }
    RobustKernelHuber rk;
    if (robust) { rk.setDelta(delta); e._robustKernel = &rk; }
    for (int k = 0; k < D; ++k) { e._error(k) = er[k]; e._information(k, k) = w0; }
    e.constructQuadraticForm();
    for (int i = 0; i < n * n; ++i) H[i] = 0;
    for (int v = 0; v < nv; ++v) {
        for (int i = 0; i < dims[v]; ++i) {
            b[off[v] + i] = vs[v].b[i];
            for (int j = 0; j < dims[v]; ++j) H[(size_t)(off[v] + i) * n + off[v] + j] = vs[v].H[(size_t)j * dims[v] + i];
        }
    }
    for (int j = 1; j < nv; ++j)
        for (int i = 0; i < j; ++i) {
            const std::vector<double>& B = blocks[internal::computeUpperTriangleIndex(i, j)];
            for (int r = 0; r < dims[i]; ++r)
                for (int c = 0; c < dims[j]; ++c) {
                    H[(size_t)(off[i] + r) * n + off[j] + c] = B[(size_t)c * dims[i] + r];
                    H[(size_t)(off[j] + c) * n + off[i] + r] = B[(size_t)c * dims[i] + r];      // the solver reads the block and its transpose
                }
        }
}
}  // namespace

extern "C" {

// BaseMultiEdge<dim>::constructQuadraticForm for nv vertices (dims[], fixed[]), Jacobians Js[v] row-major dim x dims[v]: H dense (sum dims)^2
// row-major with the diagonal blocks the vertices received and both triangles of the off-diagonal blocks, b the right-hand sides
void ref_multi_quadratic_form(int dim, int nv, const int32_t* dims, const double* const* Js, const double* er, double w0, double delta, int robust,
                              const uint8_t* fixed, double* H, double* b) {
    if (dim == 1) run_multi<1>(nv, dims, Js, er, w0, delta, robust, fixed, H, b);
    else run_multi<3>(nv, dims, Js, er, w0, delta, robust, fixed, H, b);
}

// RobustKernelHuber::robustify after setDelta(delta): rho[3]
void ref_huber(double delta, double e2, double* rho) {
    g2o::RobustKernelHuber rk;
    rk.setDelta(delta);
    Eigen::Vector3d r;
    rk.robustify(e2, r);
    rho[0] = r[0]; rho[1] = r[1]; rho[2] = r[2];
}

// BaseBinaryEdge<dim, ., point(3), pose(6)>::constructQuadraticForm on fresh (zero) vertex blocks: Ji dim x 3, Jj dim x 6 row-major,
// information w0 * I, Huber(delta) if robust.  hl/gl: the point's A / b, hp/gp: the pose's, w63: the off-diagonal block as pose x point
void ref_binary_quadratic_form(int dim, const double* Ji, const double* Jj, const double* er, double w0, double delta, int robust, int fixed_point,
                               int fixed_pose, double* hl, double* gl, double* hp, double* gp, double* w63) {
    if (dim == 2) run_binary<2>(Ji, Jj, er, w0, delta, robust, fixed_point, fixed_pose, hl, gl, hp, gp, w63);
    else run_binary<3>(Ji, Jj, er, w0, delta, robust, fixed_point, fixed_pose, hl, gl, hp, gp, w63);
}
// BaseUnaryEdge<dim, ., pose(6)>::constructQuadraticForm (the OnlyPose edges of PoseOptimization)
void ref_unary_quadratic_form(int dim, const double* J, const double* er, double w0, double delta, int robust, double* h, double* g) {
    if (dim == 2) run_unary<2>(J, er, w0, delta, robust, h, g); else run_unary<3>(J, er, w0, delta, robust, h, g);
}

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg on the hooks.  rows: one (lambda, chi2 before, chi2 after,
// accepted) per trial as the stand-ins saw them.  Returns optimize()'s return value (iterations run).
int ref_lm_optimize(const ref_lm_hooks* hk, int iterations, int max_trials, double* rows, int row_cap, int* n_rows, double* lambda_final,
                    int* n_error_evaluations, double* tau) {
    g2o::LmLog log;
    g2o::SparseOptimizer opt; opt.hk = *hk; opt.log = &log;
    g2o::Solver solver; solver.opt = &opt; opt.solver_ = &solver;
    g2o::OptimizationAlgorithmLevenberg lm(&solver);
    lm._optimizer = &opt;
    if (max_trials > 0) lm.setMaxTrialsAfterFailure(max_trials);
    opt._algorithm = &lm;
    opt._ivMap.resize(1);                 // "initializeOptimization was called"; rebuilt after the first buildSystem
    const int it = opt.optimize(iterations);
    const int n = (int)log.rows.size() / 4;
    for (int i = 0; i < n && i < row_cap; ++i) std::memcpy(rows + 4 * i, &log.rows[4 * i], 32);
    *n_rows = n; *lambda_final = lm.currentLambda(); *n_error_evaluations = opt.n_eval; *tau = lm.tau();
    return it;
}

}  // extern "C"
