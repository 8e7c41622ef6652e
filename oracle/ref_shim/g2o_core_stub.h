// g2o_core_stub.h -- TEST INFRASTRUCTURE.  Stand-ins for the g2o base classes the reference's leaf types derive from
// (Thirdparty/g2o/g2o/core/base_vertex.h, base_unary_edge.h, base_binary_edge.h, base_multi_edge.h): just the data members the leaf
// code touches (_estimate, _measurement, _information, _error, _vertices, _jacobianOplus*).  The real headers pull in all of
// g2o's graph machinery and unshimmed Eigen; defining their include guards below makes the reference's own
// `#include "../core/base_vertex.h"` lines open the real files and skip their bodies, so the leaf sources compile UNMODIFIED
// from /root/reference against these stand-ins.
#pragma once
#define G2O_BASE_VERTEX_H
#define G2O_BASE_UNARY_EDGE_H
#define G2O_BASE_BINARY_EDGE_H
#define G2O_BASE_MULTI_EDGE_H
#define G2O_FACTORY_H
#define G2O_MACROS_H
#include <iostream>
#include <vector>

#include "eigen_shim.h"

namespace g2o {
using namespace Eigen;

struct HyperGraph {
    struct Vertex { virtual ~Vertex() {} };
};
struct OptimizableGraph {
    struct Vertex : HyperGraph::Vertex {
        virtual void oplusImpl(const double* update) = 0;
        virtual void setToOriginImpl() = 0;
        void updateCache() {}
        // graph bookkeeping the Optimizer functions set (core/optimizable_graph.h: id, fixed, marginalized); used by oracle/ref_lba.cpp
        int id() const { return _id; }
        void setId(int i) { _id = i; }
        bool fixed() const { return _fixed; }
        void setFixed(bool f) { _fixed = f; }
        bool marginalized() const { return _marginalized; }
        void setMarginalized(bool m) { _marginalized = m; }
    protected:
        int _id = -1;
        bool _fixed = false, _marginalized = false;
    };
};
#ifdef G2O_STUB_WITH_GRAPH_MEMBERS
class RobustKernel {                        // core/robust_kernel.h:50-70
public:
    virtual ~RobustKernel() {}
    virtual void robustify(double squaredError, Eigen::Vector3d& rho) const = 0;
    virtual void setDelta(double delta) { _delta = delta; }
    double delta() const { return _delta; }
protected:
    double _delta = 1.0;
};
#endif

template <int D, typename T> class BaseVertex : public OptimizableGraph::Vertex {
public:
    typedef T EstimateType;
    static const int Dimension = D;
    const EstimateType& estimate() const { return _estimate; }
    void setEstimate(const EstimateType& et) { _estimate = et; }
protected:
    EstimateType _estimate;
};

template <int D, typename E> class BaseEdgeStub {
public:
    typedef E Measurement;
    typedef Matrix<double, D, 1> ErrorVector;
    typedef Matrix<double, D, D> InformationType;
    virtual ~BaseEdgeStub() {}
    const ErrorVector& error() const { return _error; }
    const Measurement& measurement() const { return _measurement; }
    virtual void setMeasurement(const Measurement& m) { _measurement = m; }
    InformationType& information() { return _information; }
    const InformationType& information() const { return _information; }
    void setVertex(size_t i, HyperGraph::Vertex* v) { if (_vertices.size() <= i) _vertices.resize(i + 1, nullptr); _vertices[i] = v; }
    std::vector<HyperGraph::Vertex*> _vertices;
#ifdef G2O_STUB_WITH_GRAPH_MEMBERS              // what Optimizer::LocalBundleAdjustment touches on an edge (oracle/ref_lba.cpp)
    void setInformation(const InformationType& i) { _information = i; }
    int level() const { return _level; }
    void setLevel(int l) { _level = l; }
    RobustKernel* robustKernel() const { return _robustKernel; }
    void setRobustKernel(RobustKernel* k) { _robustKernel = k; }
#include "_ref/lm_edge_inline.inc"             // chi2(), robustInformation(): the reference's inline bodies (core/base_edge.h:58-61, 96-102)
    int _level = 0;
    RobustKernel* _robustKernel = nullptr;
#endif
protected:
    Measurement _measurement;
    InformationType _information;
    ErrorVector _error;
};

template <int D, typename E, typename VertexXi> class BaseUnaryEdge : public BaseEdgeStub<D, E> {
public:
    BaseUnaryEdge() { this->_vertices.resize(1, nullptr); }
    Matrix<double, D, VertexXi::Dimension> _jacobianOplusXi;
    virtual void linearizeOplus() {}
};
template <int D, typename E, typename VertexXi, typename VertexXj> class BaseBinaryEdge : public BaseEdgeStub<D, E> {
public:
    BaseBinaryEdge() { this->_vertices.resize(2, nullptr); }
    Matrix<double, D, VertexXi::Dimension> _jacobianOplusXi;
    Matrix<double, D, VertexXj::Dimension> _jacobianOplusXj;
    virtual void linearizeOplus() {}
};

// Jacobian of a multi-edge w.r.t. one vertex: D rows, run-time columns (g2o maps these into a workspace)
template <int D> class DynJacobian {
    std::vector<double> v_;
    int cols_ = 0;
public:
    void setCols(int c) { cols_ = c; v_.assign((size_t)D * c, 0.0); }
    int cols() const { return cols_; }
    double& operator()(int r, int c) { return v_[(size_t)c * D + r]; }
    double operator()(int r, int c) const { return v_[(size_t)c * D + r]; }
    template <int C, int O> DynJacobian& operator=(const Matrix<double, D, C, O>& m) {
        setCols(C);
        for (int r = 0; r < D; ++r) for (int c = 0; c < C; ++c) (*this)(r, c) = m(r, c);
        return *this;
    }
};
template <int D, typename E> class BaseMultiEdge : public BaseEdgeStub<D, E> {
public:
    void resize(size_t n) { this->_vertices.resize(n, nullptr); _jacobianOplus.resize(n); }
    std::vector<DynJacobian<D>> _jacobianOplus;
    virtual void linearizeOplus() {}
};
// g2o's scalar-measurement edges (BaseMultiEdge<1, double>) index _error[0]
}  // namespace g2o
