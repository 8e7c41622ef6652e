// eigen_shim.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the parts of Eigen 3 that the reference's g2o LEAF types use
// (fixed-size Matrix, Map, Quaternion, Isometry Transform), so that the UNMODIFIED reference sources
//   Thirdparty/g2o/g2o/types/{se3_ops.h, se3_ops.hpp, se3quat.h, types_sba.{h,cpp}, types_six_dof_expmap.{h,cpp}}
//   include/{g2o_vertex_distance.h, g2o_vertex_se3.h, g2o_edge_rigidbody.h, g2o_dyn_slam3d.h}
// compile in this image (Eigen itself is absent: SURVEY.md 8c).  Only oracle/ref_leaf.cpp includes it; nothing of the product
// does.  The operations follow Eigen's own formulas and evaluation order where the result depends on it (quaternion from a
// rotation matrix, toRotationMatrix, quaternion * vector, coefficient-wise matrix products); values may still differ from a
// real Eigen build in the last bit where Eigen vectorises a reduction, which is why the tests built on it use 1e-12, not
// bit equality.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1, Dynamic = -1, Aligned = 16, Unaligned = 0, Isometry = 1, Affine = 2 };

template <typename Derived> struct MatrixBase {
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    Derived& derived() { return *static_cast<Derived*>(this); }
    int size() const { return derived().rows() * derived().cols(); }
    double operator[](int i) const { return derived().coeff(i); }
    double operator()(int i) const { return derived().coeff(i); }
};

template <typename S, int R, int C, int Opt = ColMajor> class Matrix;

// assignable view of a fixed-size block of a matrix
template <typename M, int BR, int BC> class BlockRef {
    M& m_; int r0_, c0_;
public:
    BlockRef(M& m, int r0, int c0) : m_(m), r0_(r0), c0_(c0) {}
    template <int O> BlockRef& operator=(const Matrix<double, BR, BC, O>& v) {
        for (int r = 0; r < BR; ++r) for (int c = 0; c < BC; ++c) m_(r0_ + r, c0_ + c) = v(r, c);
        return *this;
    }
    operator Matrix<double, BR, BC>() const;
};
// run-time sized block (se3quat.h adj / to_homogeneous_matrix): assign only
template <typename M> class DynBlockRef {
    M& m_; int r0_, c0_, nr_, nc_;
public:
    DynBlockRef(M& m, int r0, int c0, int nr, int nc) : m_(m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
    template <int BR, int BC, int O> DynBlockRef& operator=(const Matrix<double, BR, BC, O>& v) {
        assert(BR == nr_ && BC == nc_);
        for (int r = 0; r < BR; ++r) for (int c = 0; c < BC; ++c) m_(r0_ + r, c0_ + c) = v(r, c);
        return *this;
    }
    DynBlockRef head(int n) { return DynBlockRef(m_, r0_, c0_, n, nc_); }
};
template <typename M> class CommaInit {
    M& m_; int k_;
public:
    CommaInit(M& m, double first) : m_(m), k_(0) { put(first); }
    template <int BR, int O> CommaInit(M& m, const Matrix<double, BR, 1, O>& v) : m_(m), k_(0) { for (int i = 0; i < BR; ++i) put(v(i)); }
    void put(double v) { const int r = k_ / m_.cols(), c = k_ % m_.cols(); m_(r, c) = v; ++k_; }   // row by row, like Eigen
    CommaInit& operator,(double v) { put(v); return *this; }
};

template <typename S, int R, int C, int Opt> class Matrix : public MatrixBase<Matrix<S, R, C, Opt>> {
    S m_[R * C];   // column major
public:
    typedef S Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
    Matrix() { for (int i = 0; i < R * C; ++i) m_[i] = S(0); }
    Matrix(S x, S y) { static_assert(R * C == 2, "size"); m_[0] = x; m_[1] = y; }
    Matrix(S x, S y, S z) { static_assert(R * C == 3, "size"); m_[0] = x; m_[1] = y; m_[2] = z; }
    Matrix(S x, S y, S z, S w) { static_assert(R * C == 4, "size"); m_[0] = x; m_[1] = y; m_[2] = z; m_[3] = w; }
    template <int O2> Matrix(const Matrix<S, R, C, O2>& o) { for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) (*this)(r, c) = o(r, c); }
    template <typename D> explicit Matrix(const MatrixBase<D>& o) { for (int i = 0; i < R * C; ++i) m_[i] = o.derived().coeff(i); }
    template <typename D> Matrix& operator=(const MatrixBase<D>& o) { for (int i = 0; i < R * C; ++i) m_[i] = o.derived().coeff(i); return *this; }
    int rows() const { return R; }
    int cols() const { return C; }
    S coeff(int i) const { return m_[i]; }
    S& operator()(int r, int c) { return m_[c * R + r]; }
    S operator()(int r, int c) const { return m_[c * R + r]; }
    S& operator()(int i) { return m_[i]; }
    S operator()(int i) const { return m_[i]; }
    S& operator[](int i) { return m_[i]; }
    S operator[](int i) const { return m_[i]; }
    S* data() { return m_; }
    const S* data() const { return m_; }
    S& x() { return m_[0]; } S& y() { return m_[1]; } S& z() { return m_[2]; }
    S x() const { return m_[0]; } S y() const { return m_[1]; } S z() const { return m_[2]; }
    void fill(S v) { for (int i = 0; i < R * C; ++i) m_[i] = v; }
    void setZero() { fill(S(0)); }
    Matrix& noalias() { return *this; }
    void setIdentity() { setZero(); for (int i = 0; i < (R < C ? R : C); ++i) (*this)(i, i) = S(1); }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int, int) { return Matrix(); }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    S squaredNorm() const { S s = S(0); for (int i = 0; i < R * C; ++i) s += m_[i] * m_[i]; return s; }
    S norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const S n = norm(); for (int i = 0; i < R * C; ++i) m_[i] /= n; }
    S dot(const Matrix& o) const { S s = S(0); for (int i = 0; i < R * C; ++i) s += m_[i] * o.m_[i]; return s; }
    Matrix cross(const Matrix& o) const {
        static_assert(R * C == 3, "cross");
        return Matrix(m_[1] * o.m_[2] - m_[2] * o.m_[1], m_[2] * o.m_[0] - m_[0] * o.m_[2], m_[0] * o.m_[1] - m_[1] * o.m_[0]);
    }
    Matrix<S, C, R> transpose() const { Matrix<S, C, R> t; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c); return t; }
    Matrix operator-() const { Matrix o; for (int i = 0; i < R * C; ++i) o.m_[i] = -m_[i]; return o; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m_[i] += o.m_[i]; return *this; }
    template <typename D> Matrix& operator+=(const MatrixBase<D>& o) { for (int i = 0; i < R * C; ++i) m_[i] += o.derived().coeff(i); return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m_[i] -= o.m_[i]; return *this; }
    Matrix& operator*=(S k) { for (int i = 0; i < R * C; ++i) m_[i] *= k; return *this; }
    template <int N> Matrix<S, N, 1> head() const { Matrix<S, N, 1> h; for (int i = 0; i < N; ++i) h(i) = m_[i]; return h; }
    template <int BR, int BC> BlockRef<Matrix, BR, BC> block(int r0, int c0) { return BlockRef<Matrix, BR, BC>(*this, r0, c0); }
    template <int BR, int BC> Matrix<S, BR, BC> block(int r0, int c0) const {
        Matrix<S, BR, BC> b; for (int r = 0; r < BR; ++r) for (int c = 0; c < BC; ++c) b(r, c) = (*this)(r0 + r, c0 + c); return b;
    }
    DynBlockRef<Matrix> block(int r0, int c0, int nr, int nc) { return DynBlockRef<Matrix>(*this, r0, c0, nr, nc); }
    DynBlockRef<Matrix> col(int c) { return DynBlockRef<Matrix>(*this, 0, c, R, 1); }
    template <int BR, int BC> Matrix<S, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    CommaInit<Matrix> operator<<(S v) { return CommaInit<Matrix>(*this, v); }
    Matrix inverse() const;   // 3 x 3 only
};
template <typename M, int BR, int BC> BlockRef<M, BR, BC>::operator Matrix<double, BR, BC>() const {
    Matrix<double, BR, BC> b; for (int r = 0; r < BR; ++r) for (int c = 0; c < BC; ++c) b(r, c) = m_(r0_ + r, c0_ + c); return b;
}

template <typename S, int R, int C, int O, int O2> Matrix<S, R, C> operator+(const Matrix<S, R, C, O>& a, const Matrix<S, R, C, O2>& b) {
    Matrix<S, R, C> o; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) o(r, c) = a(r, c) + b(r, c); return o;
}
template <typename S, int R, int C, int O, int O2> Matrix<S, R, C> operator-(const Matrix<S, R, C, O>& a, const Matrix<S, R, C, O2>& b) {
    Matrix<S, R, C> o; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) o(r, c) = a(r, c) - b(r, c); return o;
}
template <typename S, int R, int C, int O> Matrix<S, R, C> operator*(const Matrix<S, R, C, O>& a, double k) {
    Matrix<S, R, C> o; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) o(r, c) = a(r, c) * k; return o;
}
template <typename S, int R, int C, int O> Matrix<S, R, C> operator*(double k, const Matrix<S, R, C, O>& a) {
    Matrix<S, R, C> o; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) o(r, c) = k * a(r, c); return o;
}
template <typename S, int R, int C, int O> Matrix<S, R, C> operator/(const Matrix<S, R, C, O>& a, double k) {
    Matrix<S, R, C> o; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) o(r, c) = a(r, c) / k; return o;
}
template <typename S, int R, int K, int C, int O, int O2> Matrix<S, R, C> operator*(const Matrix<S, R, K, O>& a, const Matrix<S, K, C, O2>& b) {
    Matrix<S, R, C> o;
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) { S s = a(r, 0) * b(0, c); for (int k = 1; k < K; ++k) s += a(r, k) * b(k, c); o(r, c) = s; }
    return o;
}
template <typename S, int R, int C, int O> std::ostream& operator<<(std::ostream& os, const Matrix<S, R, C, O>& m) {
    for (int r = 0; r < R; ++r) { for (int c = 0; c < C; ++c) os << m(r, c) << (c + 1 < C ? " " : ""); if (r + 1 < R) os << "\n"; }
    return os;
}
template <typename S, int R, int C, int Opt> Matrix<S, R, C, Opt> Matrix<S, R, C, Opt>::inverse() const {
    static_assert(R == 3 && C == 3, "only the 3 x 3 inverse is provided");
    const Matrix& a = *this; Matrix o;
    const S c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1), c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2), c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
    const S id = S(1) / (a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02);
    o(0, 0) = c00 * id; o(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id; o(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
    o(1, 0) = c01 * id; o(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id; o(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
    o(2, 0) = c02 * id; o(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id; o(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
    return o;
}

// run-time sized matrix: only what Optimizer.cc does with its `Eigen::MatrixXd Info` objects (Identity, element access, scaling, and the
// conversion to the fixed-size information matrix of the edge it is handed to)
class MatrixXd {
    std::vector<double> v_; int r_ = 0, c_ = 0;
public:
    MatrixXd() {}
    MatrixXd(int r, int c) : v_((size_t)r * c, 0.0), r_(r), c_(c) {}
    static MatrixXd Identity(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0; return m; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    double& operator()(int r, int c) { return v_[(size_t)c * r_ + r]; }
    double operator()(int r, int c) const { return v_[(size_t)c * r_ + r]; }
    template <int R, int C, int O> operator Matrix<double, R, C, O>() const {
        assert(R == r_ && C == c_);
        Matrix<double, R, C, O> m; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) m(r, c) = (*this)(r, c); return m;
    }
    // what BaseMultiEdge::computeQuadraticForm does with its run-time sized Jacobians (core/base_multi_edge.hpp:170-222)
    template <int R, int C, int O> MatrixXd(const Matrix<double, R, C, O>& m) : v_((size_t)R * C), r_(R), c_(C) { for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) (*this)(r, c) = m(r, c); }
    MatrixXd transpose() const { MatrixXd t(c_, r_); for (int r = 0; r < r_; ++r) for (int c = 0; c < c_; ++c) t(c, r) = (*this)(r, c); return t; }
    MatrixXd operator*(const MatrixXd& b) const {
        assert(c_ == b.r_);
        MatrixXd o(r_, b.c_);
        for (int r = 0; r < r_; ++r) for (int c = 0; c < b.c_; ++c) { double s = 0; for (int k = 0; k < c_; ++k) s += (*this)(r, k) * b(k, c); o(r, c) = s; }
        return o;
    }
    template <int R, int C, int O> MatrixXd operator*(const Matrix<double, R, C, O>& b) const { return *this * MatrixXd(b); }
};
typedef MatrixXd VectorXd;
inline MatrixXd operator*(double k, const MatrixXd& m) { MatrixXd o(m.rows(), m.cols()); for (int r = 0; r < m.rows(); ++r) for (int c = 0; c < m.cols(); ++c) o(r, c) = k * m(r, c); return o; }
inline MatrixXd operator*(const MatrixXd& m, double k) { return k * m; }

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

// Map<const VectorN> / Map<VectorN> over a raw double array
template <typename T> struct MapPlain { typedef T type; typedef double* ptr; };
template <typename T> struct MapPlain<const T> { typedef T type; typedef const double* ptr; };
template <typename T, int MapOptions = Unaligned> class Map : public MatrixBase<Map<T, MapOptions>> {
    typedef typename MapPlain<T>::type Plain;
    typename MapPlain<T>::ptr p_;
public:
    explicit Map(typename MapPlain<T>::ptr p) : p_(p) {}
    int rows() const { return Plain::RowsAtCompileTime; }
    int cols() const { return Plain::ColsAtCompileTime; }
    double coeff(int i) const { return p_[i]; }
    operator Plain() const { Plain m; for (int i = 0; i < rows() * cols(); ++i) m(i) = p_[i]; return m; }
    Map& operator=(const Plain& v) { for (int i = 0; i < rows() * cols(); ++i) const_cast<double*>(p_)[i] = v(i); return *this; }
    template <int BR, int BC> Matrix<double, BR, BC> block(int r0, int c0) const {
        Matrix<double, BR, BC> b; for (int r = 0; r < BR; ++r) for (int c = 0; c < BC; ++c) b(r, c) = p_[(c0 + c) * rows() + r0 + r]; return b;
    }
    template <int N> Matrix<double, N, 1> head() const { Matrix<double, N, 1> h; for (int i = 0; i < N; ++i) h(i) = p_[i]; return h; }
};

// Map<MatrixXd>(ptr, rows, cols) / Map<VectorXd>(ptr, n): column-major view that the multi-edge accumulates into
template <int MapOptions> class Map<MatrixXd, MapOptions> {
    double* p_; int r_, c_;
public:
    Map(double* p, int r, int c) : p_(p), r_(r), c_(c) {}
    Map(double* p, int n) : p_(p), r_(n), c_(1) {}
    Map& noalias() { return *this; }
    Map& operator+=(const MatrixXd& m) { assert(m.rows() == r_ && m.cols() == c_); for (int c = 0; c < c_; ++c) for (int r = 0; r < r_; ++r) p_[(size_t)c * r_ + r] += m(r, c); return *this; }
    double operator()(int r, int c) const { return p_[(size_t)c * r_ + r]; }
};

// Eigen::Quaternion: coefficients stored x, y, z, w; constructor order w, x, y, z
template <typename S> class Quaternion {
    Matrix<S, 4, 1> c_;
public:
    Quaternion() {}
    Quaternion(S w, S x, S y, S z) : c_(x, y, z, w) {}
    explicit Quaternion(const Matrix<S, 3, 3>& m) {   // Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other, 3, 3>
        S t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > S(0)) {
            t = std::sqrt(t + S(1.0));
            w() = S(0.5) * t;
            t = S(0.5) / t;
            x() = (m(2, 1) - m(1, 2)) * t; y() = (m(0, 2) - m(2, 0)) * t; z() = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + S(1.0));
            c_(i) = S(0.5) * t;
            t = S(0.5) / t;
            w() = (m(k, j) - m(j, k)) * t;
            c_(j) = (m(j, i) + m(i, j)) * t;
            c_(k) = (m(k, i) + m(i, k)) * t;
        }
    }
    Matrix<S, 4, 1>& coeffs() { return c_; }
    const Matrix<S, 4, 1>& coeffs() const { return c_; }
    S& x() { return c_(0); } S& y() { return c_(1); } S& z() { return c_(2); } S& w() { return c_(3); }
    S x() const { return c_(0); } S y() const { return c_(1); } S z() const { return c_(2); } S w() const { return c_(3); }
    void setIdentity() { c_ = Matrix<S, 4, 1>(0, 0, 0, 1); }
    S squaredNorm() const { return c_.squaredNorm(); }
    S norm() const { return c_.norm(); }
    void normalize() { c_.normalize(); }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Matrix<S, 3, 3> toRotationMatrix() const {      // QuaternionBase::toRotationMatrix
        Matrix<S, 3, 3> r;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w(), txx = tx * x(), txy = ty * x(), txz = tz * x(), tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        r(0, 0) = S(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = S(1) - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = S(1) - (txx + tyy);
        return r;
    }
    Quaternion operator*(const Quaternion& b) const {   // internal::quat_product
        const Quaternion& a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(), a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(), a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
    Matrix<S, 3, 1> operator*(const Matrix<S, 3, 1>& v) const {   // QuaternionBase::_transformVector
        const Matrix<S, 3, 1> q(x(), y(), z());
        Matrix<S, 3, 1> uv = q.cross(v);
        uv += uv;
        return v + w() * uv + q.cross(uv);
    }
};
typedef Quaternion<double> Quaterniond;

// Transform<double, 3, Isometry>: 4 x 4 homogeneous matrix with the Isometry fast paths of Eigen (inverse = transpose)
template <typename S, int Dim, int Mode, int Opt = ColMajor> class Transform {
    static_assert(Dim == 3 && Mode == Isometry, "only Isometry3 is provided");
    Matrix<S, 4, 4> m_;
    class TranslationRef {
        Matrix<S, 4, 4>& m_;
    public:
        explicit TranslationRef(Matrix<S, 4, 4>& m) : m_(m) {}
        TranslationRef& operator=(const Matrix<S, 3, 1>& v) { for (int i = 0; i < 3; ++i) m_(i, 3) = v(i); return *this; }
        operator Matrix<S, 3, 1>() const { return Matrix<S, 3, 1>(m_(0, 3), m_(1, 3), m_(2, 3)); }
        Matrix<S, 3, 1> operator*(S k) const { return Matrix<S, 3, 1>(m_(0, 3) * k, m_(1, 3) * k, m_(2, 3) * k); }
        TranslationRef& operator<<(const Matrix<S, 3, 1>& v) { return *this = v; }   // `t << expr`: comma initialiser with one block
    };
public:
    typedef Matrix<S, 3, 3> ConstLinearPart;
    Transform() { m_.setIdentity(); }
    explicit Transform(const Quaternion<S>& q) { m_.setIdentity(); *this = q.toRotationMatrix(); }
    explicit Transform(const Matrix<S, 4, 4>& m) : m_(m) {}
    static Transform Identity() { return Transform(); }
    Transform& operator=(const Matrix<S, 3, 3>& r) {   // Transform = linear part: translation zero, last row (0 0 0 1)
        m_.setIdentity();
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m_(i, j) = r(i, j);
        return *this;
    }
    const Matrix<S, 4, 4>& matrix() const { return m_; }
    Matrix<S, 4, 4>& matrix() { return m_; }
    TranslationRef translation() { return TranslationRef(m_); }
    Matrix<S, 3, 1> translation() const { return Matrix<S, 3, 1>(m_(0, 3), m_(1, 3), m_(2, 3)); }
    Matrix<S, 3, 3> linear() const { return m_.template block<3, 3>(0, 0); }
    Matrix<S, 3, 3> rotation() const { return linear(); }   // Isometry mode: rotation() == linear()
    Transform inverse() const {   // Transform::inverse(Isometry): R^T, -R^T t
        Transform o;
        const Matrix<S, 3, 3> rt = linear().transpose();
        o = rt;
        o.translation() = -(rt * translation());
        return o;
    }
    Transform operator*(const Transform& b) const {
        Transform o;
        const Matrix<S, 4, 4> p = m_ * b.m_;
        o.m_ = p;
        return o;
    }
    Matrix<S, 3, 1> operator*(const Matrix<S, 3, 1>& v) const { return linear() * v + translation(); }
};
typedef Transform<double, 3, Isometry> Isometry3d;

}  // namespace Eigen
