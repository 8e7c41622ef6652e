// TEST INFRASTRUCTURE (oracle/_ref): the reference's own BlockSolver<Traits>::solve()
// (Thirdparty/g2o/g2o/core/block_solver.hpp:353-483 -- the Schur complement over the landmarks, the reduced right-hand side and
// the landmark back-substitution) compiled from /root/reference.  The function's text is taken out of the reference tree at build
// time (oracle/_ref/schur_solve.inc, git-ignored; oracle/extract_ref_fn.py) and compiled here, unmodified, as a member of a
// stand-in BlockSolver whose containers offer exactly the interface that text uses:
//   SparseBlockMatrix<M>          core/sparse_block_matrix.h      (blockCols() as vector<map<int, M*>>, add(), clear(), block bases)
//   SparseBlockMatrixCCS<M>       core/sparse_block_matrix_ccs.h  (RowBlock / SparseColumn, rightMultiply: dest += A^T src)
//   SparseBlockMatrixDiagonal<M>  core/sparse_block_matrix_diagonal.h (diagonal(), multiply: dest += A src)
//   fixed-size matrices           stand-ins of Eigen::Matrix<double, R, C>: products accumulate over the inner index in ascending
//                                 order, the 3 x 3 inverse is the cofactor formula (Eigen's compute_inverse for size 3)
// NOT the reference's: the linear solver behind _linearSolver->solve() (Eigen's LDLT / SimplicialLDLT in the reference: any
// backward-stable FP64 factorisation) -- a plain dense Cholesky below.
// Only tests/ use this library; nothing of the product links or loads it.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <vector>

namespace Eigen {}
namespace g2o {
using namespace std;
using namespace Eigen;

inline double get_monotonic_time() { return 0.0; }
struct G2OBatchStatistics {
    double timeSchurComplement, timeLinearSolver;
    int hessianPoseDimension, hessianLandmarkDimension, hessianDimension;
    static G2OBatchStatistics* globalStats() { return nullptr; }
};

// ---- fixed-size column-major matrix with the operations solve() applies to its blocks -------------------------------------------
template <int R, int C> struct Mat {
    double m[R * C];
    Mat() { for (int i = 0; i < R * C; ++i) m[i] = 0.0; }
    explicit Mat(int) { for (int i = 0; i < R * C; ++i) m[i] = 0.0; }      // Eigen: vector of a (fixed) size
    int rows() const { return R; }
    int cols() const { return C; }
    double& operator()(int r, int c) { return m[c * R + r]; }
    double operator()(int r, int c) const { return m[c * R + r]; }
    double& operator[](int i) { return m[i]; }
    double operator[](int i) const { return m[i]; }
    Mat& noalias() { return *this; }
    Mat<C, R> transpose() const { Mat<C, R> t; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c); return t; }
    Mat& operator-=(const Mat& o) { for (int i = 0; i < R * C; ++i) m[i] -= o.m[i]; return *this; }
    Mat& operator+=(const Mat& o) { for (int i = 0; i < R * C; ++i) m[i] += o.m[i]; return *this; }
    Mat inverse() const {
        static_assert(R == 3 && C == 3, "only the landmark block is inverted");
        const Mat& a = *this;
        Mat cof;   // cofactor(i, j) stored transposed, as Eigen's compute_inverse_size3_helper does
        cof(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
        cof(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
        cof(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
        const double det = a(0, 0) * cof(0, 0) + a(0, 1) * cof(1, 0) + a(0, 2) * cof(2, 0);
        const double invdet = 1.0 / det;
        Mat r;
        r(0, 0) = cof(0, 0) * invdet; r(1, 0) = cof(1, 0) * invdet; r(2, 0) = cof(2, 0) * invdet;
        r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * invdet;
        r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * invdet;
        r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * invdet;
        r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * invdet;
        r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * invdet;
        r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * invdet;
        return r;
    }
    struct MapType {   // Eigen::Map over a raw array: Bb.noalias() += (*Bi) * db
        double* p; int n;
        MapType(double* p_, int n_) : p(p_), n(n_) {}
        MapType& noalias() { return *this; }
        MapType& operator+=(const Mat<R, C>& o) { static_assert(C == 1, "vector map"); for (int i = 0; i < R; ++i) p[i] += o.m[i]; return *this; }
    };
};
template <int R, int K, int C> Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
    Mat<R, C> o;
    for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) {
            double s = a(r, 0) * b(0, c);
            for (int k = 1; k < K; ++k) s += a(r, k) * b(k, c);
            o(r, c) = s;
        }
    return o;
}

// ---- run-time sized matrix (Eigen::MatrixXd / VectorXd of BlockSolverX) with the same operations -----------------------------------
struct MatX {
    int r = 0, c = 0;
    std::vector<double> m;
    MatX() {}
    explicit MatX(int n) : r(n), c(1), m((size_t)n, 0.0) {}
    MatX(int r_, int c_) : r(r_), c(c_), m((size_t)r_ * c_, 0.0) {}
    int rows() const { return r; }
    int cols() const { return c; }
    double& operator()(int i, int j) { return m[(size_t)j * r + i]; }
    double operator()(int i, int j) const { return m[(size_t)j * r + i]; }
    double& operator[](int i) { return m[i]; }
    double operator[](int i) const { return m[i]; }
    MatX& noalias() { return *this; }
    MatX transpose() const { MatX t(c, r); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j); return t; }
    MatX& operator-=(const MatX& o) { assert(o.r == r && o.c == c); for (size_t i = 0; i < m.size(); ++i) m[i] -= o.m[i]; return *this; }
    MatX& operator+=(const MatX& o) { assert(o.r == r && o.c == c); for (size_t i = 0; i < m.size(); ++i) m[i] += o.m[i]; return *this; }
    MatX inverse() const {
        assert(r == 3 && c == 3 && "only the landmark block is inverted");
        Mat<3, 3> a;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a(i, j) = (*this)(i, j);
        const Mat<3, 3> inv = a.inverse();
        MatX o(3, 3);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o(i, j) = inv(i, j);
        return o;
    }
    struct MapType {
        double* p; int n;
        MapType(double* p_, int n_) : p(p_), n(n_) {}
        MapType& noalias() { return *this; }
        MapType& operator+=(const MatX& o) { assert(o.c == 1 && o.r == n); for (int i = 0; i < n; ++i) p[i] += o.m[i]; return *this; }
    };
};
inline MatX operator*(const MatX& a, const MatX& b) {
    assert(a.c == b.r);
    MatX o(a.r, b.c);
    for (int c = 0; c < b.c; ++c)
        for (int r = 0; r < a.r; ++r) {
            double s = a(r, 0) * b(0, c);
            for (int k = 1; k < a.c; ++k) s += a(r, k) * b(k, c);
            o(r, c) = s;
        }
    return o;
}
template <int R, int C> inline Mat<R, C>* new_block(const Mat<R, C>*, int, int) { return new Mat<R, C>(); }
inline MatX* new_block(const MatX*, int r, int c) { return new MatX(r, c); }
template <int R, int C> inline void zero_block(Mat<R, C>& b) { b = Mat<R, C>(); }
inline void zero_block(MatX& b) { std::fill(b.m.begin(), b.m.end(), 0.0); }

// ---- containers ------------------------------------------------------------------------------------------------------------------
template <class M> class SparseBlockMatrix {
public:
    typedef std::map<int, M*> IntBlockMap;
    SparseBlockMatrix(const std::vector<int>& rbi, const std::vector<int>& cbi) : _rowBlockIndices(rbi), _colBlockIndices(cbi), _blockCols(cbi.size()) {}
    ~SparseBlockMatrix() { for (auto& col : _blockCols) for (auto& kv : col) delete kv.second; }
    int rows() const { return _rowBlockIndices.size() ? _rowBlockIndices.back() : 0; }
    int cols() const { return _colBlockIndices.size() ? _colBlockIndices.back() : 0; }
    int rowBaseOfBlock(int r) const { return r ? _rowBlockIndices[r - 1] : 0; }      // (r = number of blocks gives the total, used for block sizes)
    int colBaseOfBlock(int c) const { return c ? _colBlockIndices[c - 1] : 0; }
    const std::vector<IntBlockMap>& blockCols() const { return _blockCols; }
    std::vector<IntBlockMap>& blockCols() { return _blockCols; }
    M* block(int r, int c, bool alloc = false) {
        auto it = _blockCols[c].find(r);
        if (it != _blockCols[c].end()) return it->second;
        if (!alloc) return nullptr;
        M* b = new_block((const M*)nullptr, rowBaseOfBlock(r + 1) - rowBaseOfBlock(r), colBaseOfBlock(c + 1) - colBaseOfBlock(c));
        _blockCols[c][r] = b;
        return b;
    }
    void clear() { for (auto& col : _blockCols) for (auto& kv : col) zero_block(*kv.second); }     // keeps the pattern (sparse_block_matrix.hpp:38-57, dealloc = false)
    bool add(SparseBlockMatrix* dest) const {                                                     // sparse_block_matrix.hpp:204-236: dest's blocks += ours
        for (size_t c = 0; c < _blockCols.size(); ++c)
            for (const auto& kv : _blockCols[c]) *dest->block(kv.first, (int)c, true) += *kv.second;
        return true;
    }
private:
    std::vector<int> _rowBlockIndices, _colBlockIndices;
    std::vector<IntBlockMap> _blockCols;
};

template <class M> class SparseBlockMatrixCCS {
public:
    struct RowBlock {
        int row; M* block;
        RowBlock() : row(-1), block(0) {}
        RowBlock(int r, M* b) : row(r), block(b) {}
        bool operator<(const RowBlock& other) const { return row < other.row; }
    };
    typedef std::vector<RowBlock> SparseColumn;
    SparseBlockMatrixCCS(const std::vector<int>& rbi, const std::vector<int>& cbi) : _rowBlockIndices(rbi), _colBlockIndices(cbi), _blockCols(cbi.size()) {}
    int rows() const { return _rowBlockIndices.size() ? _rowBlockIndices.back() : 0; }
    int cols() const { return _colBlockIndices.size() ? _colBlockIndices.back() : 0; }
    int rowBaseOfBlock(int r) const { return r ? _rowBlockIndices[r - 1] : 0; }
    int colBaseOfBlock(int c) const { return c ? _colBlockIndices[c - 1] : 0; }
    const std::vector<SparseColumn>& blockCols() const { return _blockCols; }
    std::vector<SparseColumn>& blockCols() { return _blockCols; }
    // sparse_block_matrix_ccs.h:103-128: dest += A^T * src, block column by block column, blocks in row order
    void rightMultiply(double*& dest, const double* src) const {
        for (int i = 0; i < (int)_blockCols.size(); ++i) {
            const int destOffset = colBaseOfBlock(i);
            for (const RowBlock& rb : _blockCols[i]) {
                const M& a = *rb.block;
                const int srcOffset = rowBaseOfBlock(rb.row);
                for (int c = 0; c < a.cols(); ++c) {          // y.segment(yoff) += A.transpose() * x.segment(xoff)
                    double s = a(0, c) * src[srcOffset];
                    for (int r = 1; r < a.rows(); ++r) s += a(r, c) * src[srcOffset + r];
                    dest[destOffset + c] += s;
                }
            }
        }
    }
private:
    std::vector<int> _rowBlockIndices, _colBlockIndices;
    std::vector<SparseColumn> _blockCols;
};

template <class M> class SparseBlockMatrixDiagonal {
public:
    explicit SparseBlockMatrixDiagonal(const std::vector<int>& bi) : _blockIndices(bi), _diagonal(bi.size()) {}
    int cols() const { return _blockIndices.size() ? _blockIndices.back() : 0; }
    int baseOfBlock(int r) const { return r ? _blockIndices[r - 1] : 0; }
    std::vector<M>& diagonal() { return _diagonal; }
    // sparse_block_matrix_diagonal.h:77-99: dest += A * src per diagonal block
    void multiply(double*& dest, const double* src) const {
        for (int i = 0; i < (int)_diagonal.size(); ++i) {
            const int off = baseOfBlock(i);
            const M& a = _diagonal[i];
            for (int r = 0; r < a.rows(); ++r) {
                double s = a(r, 0) * src[off];
                for (int c = 1; c < a.cols(); ++c) s += a(r, c) * src[off + c];
                dest[off + r] += s;
            }
        }
    }
private:
    std::vector<int> _blockIndices;
    std::vector<M> _diagonal;
};

// the solver behind the reduced system: symmetric from the stored upper block triangle, dense Cholesky (NOT the reference's LDLT)
template <class M> struct LinearSolver {
    bool solve(const SparseBlockMatrix<M>& A, double* x, double* b) {
        const int n = A.cols();
        std::vector<double> a((size_t)n * n, 0.0);
        for (size_t c = 0; c < A.blockCols().size(); ++c)
            for (const auto& kv : A.blockCols()[c]) {
                const int r0 = A.rowBaseOfBlock(kv.first), c0 = A.colBaseOfBlock((int)c);
                const M& blk = *kv.second;
                for (int i = 0; i < blk.rows(); ++i)
                    for (int j = 0; j < blk.cols(); ++j) {
                        a[(size_t)(r0 + i) * n + c0 + j] = blk(i, j);
                        a[(size_t)(c0 + j) * n + r0 + i] = blk(i, j);
                    }
            }
        for (int j = 0; j < n; ++j) {
            double d = a[(size_t)j * n + j];
            for (int k = 0; k < j; ++k) d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
            if (!(d > 0.0)) return false;
            d = std::sqrt(d);
            a[(size_t)j * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double s = a[(size_t)i * n + j];
                for (int k = 0; k < j; ++k) s -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
                a[(size_t)i * n + j] = s / d;
            }
        }
        for (int i = 0; i < n; ++i) {
            double s = b[i];
            for (int k = 0; k < i; ++k) s -= a[(size_t)i * n + k] * x[k];
            x[i] = s / a[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = x[i];
            for (int k = i + 1; k < n; ++k) s -= a[(size_t)k * n + i] * x[k];
            x[i] = s / a[(size_t)i * n + i];
        }
        return true;
    }
};

template <int p, int l> struct BlockSolverTraits {
    typedef Mat<p, p> PoseMatrixType;
    typedef Mat<l, l> LandmarkMatrixType;
    typedef Mat<p, l> PoseLandmarkMatrixType;
    typedef Mat<p, 1> PoseVectorType;
    typedef Mat<l, 1> LandmarkVectorType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
};

template <> struct BlockSolverTraits<-1, -1> {    // BlockSolverX (core/block_solver.h:74-92: Eigen::Dynamic everywhere)
    typedef MatX PoseMatrixType;
    typedef MatX LandmarkMatrixType;
    typedef MatX PoseLandmarkMatrixType;
    typedef MatX PoseVectorType;
    typedef MatX LandmarkVectorType;
    typedef LinearSolver<PoseMatrixType> LinearSolverType;
};

// the members solve() touches, named and typed as core/block_solver.h:95-182 declares them
template <typename Traits> class BlockSolver {
public:
    typedef typename Traits::PoseMatrixType PoseMatrixType;
    typedef typename Traits::LandmarkMatrixType LandmarkMatrixType;
    typedef typename Traits::PoseLandmarkMatrixType PoseLandmarkMatrixType;
    typedef typename Traits::PoseVectorType PoseVectorType;
    typedef typename Traits::LandmarkVectorType LandmarkVectorType;
    typedef typename Traits::LinearSolverType LinearSolverType;
    bool solve();
    SparseBlockMatrix<PoseMatrixType>* _Hpp = nullptr;
    SparseBlockMatrix<LandmarkMatrixType>* _Hll = nullptr;
    SparseBlockMatrix<PoseMatrixType>* _Hschur = nullptr;
    SparseBlockMatrixDiagonal<LandmarkMatrixType>* _DInvSchur = nullptr;
    SparseBlockMatrixCCS<PoseLandmarkMatrixType>* _HplCCS = nullptr;
    SparseBlockMatrixCCS<PoseMatrixType>* _HschurTransposedCCS = nullptr;
    LinearSolverType* _linearSolver = nullptr;
    bool _doSchur = true;
    double* _coefficients = nullptr;
    double* _bschur = nullptr;
    double* _x = nullptr;
    double* _b = nullptr;
    int _sizePoses = 0, _sizeLandmarks = 0;
};

template <typename Traits>
#include "_ref/schur_solve.inc"

}  // namespace g2o

// One BlockSolver_6_3::solve() on a system handed over as flat arrays (the layout oracle/ba_oracle.cpp dumps through ba_oracle_lm_system):
//   n_poses free poses (block i at 6 i), n_points landmarks (block l at 3 l behind the poses), edges (pose, point) with a 6 x 3 block W
//   (row major) each, sorted by nothing in particular; Hpp: the n_poses diagonal 6 x 6 blocks (row major, symmetric), Hll: 3 x 3 per
//   landmark (row major), b: right-hand side of the whole system, lambda: added to both diagonals as BlockSolver::setLambda does
//   (block_solver.hpp:494-518).  Out: x (whole solution), hschur (dense (6 n_poses)^2, upper triangle as the reference stores it, mirrored),
//   bschur.  Returns 1 when the reduced system was solved.
extern "C" int ref_schur_solve(int n_poses, int n_points, int n_edges, const int* edge_pose, const int* edge_point, const double* W,
                               const double* Hpp, const double* Hll, const double* b, double lambda, double* x, double* hschur, double* bschur) {
    using namespace g2o;
    typedef BlockSolverTraits<6, 3> T;
    std::vector<int> pidx(n_poses), lidx(n_points);
    for (int i = 0; i < n_poses; ++i) pidx[i] = 6 * (i + 1);
    for (int l = 0; l < n_points; ++l) lidx[l] = 3 * (l + 1);
    BlockSolver<T> S;
    SparseBlockMatrix<T::PoseMatrixType> hpp(pidx, pidx), hs(pidx, pidx);
    SparseBlockMatrix<T::LandmarkMatrixType> hll(lidx, lidx);
    SparseBlockMatrixDiagonal<T::LandmarkMatrixType> dinv(lidx);
    SparseBlockMatrixCCS<T::PoseLandmarkMatrixType> hpl(pidx, lidx);
    SparseBlockMatrixCCS<T::PoseMatrixType> hst(pidx, pidx);
    std::vector<T::PoseLandmarkMatrixType> wblocks(n_edges);
    for (int i = 0; i < n_poses; ++i) {
        T::PoseMatrixType* m = hpp.block(i, i, true);
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) (*m)(r, c) = Hpp[(size_t)36 * i + 6 * r + c];
        for (int r = 0; r < 6; ++r) (*m)(r, r) += lambda;
    }
    for (int l = 0; l < n_points; ++l) {
        T::LandmarkMatrixType* m = hll.block(l, l, true);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) (*m)(r, c) = Hll[(size_t)9 * l + 3 * r + c];
        for (int r = 0; r < 3; ++r) (*m)(r, r) += lambda;
    }
    // Hpl in CCS form: per landmark column the (pose row, block) pairs sorted by row (sparse_block_matrix.hpp:600-620 fillSparseBlockMatrixCCS)
    for (int e = 0; e < n_edges; ++e) {
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 3; ++c) wblocks[e](r, c) = W[(size_t)18 * e + 3 * r + c];
        hpl.blockCols()[edge_point[e]].push_back(SparseBlockMatrixCCS<T::PoseLandmarkMatrixType>::RowBlock(edge_pose[e], &wblocks[e]));
    }
    for (auto& col : hpl.blockCols()) std::sort(col.begin(), col.end());
    // the pattern of Hschur (block_solver.hpp:200-240): the diagonal plus (i1, i2), i1 <= i2, for every pair of poses that share a landmark
    for (int i = 0; i < n_poses; ++i) hs.block(i, i, true);
    for (const auto& col : hpl.blockCols())
        for (size_t a = 0; a < col.size(); ++a)
            for (size_t c2 = a; c2 < col.size(); ++c2) hs.block(col[a].row, col[c2].row, true);
    // its transposed CCS view: column i1 lists the blocks (i1, i2) by ascending i2 (sparse_block_matrix.hpp:622-650 fillSparseBlockMatrixCCSTransposed)
    for (size_t c = 0; c < hs.blockCols().size(); ++c)
        for (auto& kv : hs.blockCols()[c]) hst.blockCols()[kv.first].push_back(SparseBlockMatrixCCS<T::PoseMatrixType>::RowBlock((int)c, kv.second));
    for (auto& col : hst.blockCols()) std::sort(col.begin(), col.end());
    const int np = 6 * n_poses, nl = 3 * n_points;
    std::vector<double> coeff((size_t)np + nl, 0.0), bcopy(b, b + np + nl), bs((size_t)std::max(np, 1), 0.0);
    LinearSolver<T::PoseMatrixType> ls;
    S._Hpp = &hpp; S._Hll = &hll; S._Hschur = &hs; S._DInvSchur = &dinv; S._HplCCS = &hpl; S._HschurTransposedCCS = &hst; S._linearSolver = &ls;
    S._coefficients = coeff.data(); S._bschur = bs.data(); S._x = x; S._b = bcopy.data(); S._sizePoses = np; S._sizeLandmarks = nl;
    for (int i = 0; i < np + nl; ++i) x[i] = 0.0;
    const bool ok = S.solve();
    if (hschur) {
        for (size_t i = 0; i < (size_t)np * np; ++i) hschur[i] = 0.0;
        for (size_t c = 0; c < hs.blockCols().size(); ++c)
            for (const auto& kv : hs.blockCols()[c])
                for (int r = 0; r < 6; ++r)
                    for (int cc = 0; cc < 6; ++cc) {
                        hschur[(size_t)(6 * kv.first + r) * np + 6 * c + cc] = (*kv.second)(r, cc);
                        hschur[(size_t)(6 * c + cc) * np + 6 * kv.first + r] = (*kv.second)(r, cc);
                    }
    }
    if (bschur) for (int i = 0; i < np; ++i) bschur[i] = bs[i];
    return ok ? 1 : 0;
}

// The same through BlockSolverX = BlockSolver<BlockSolverTraits<Dynamic, Dynamic>> (what the AirDOS window uses: key-frames 6, bone lengths 1,
// motions 6 and joints 3 wide stay in the reduced system, the map points are marginalised; src/Optimizer.cc:1508-1516): n_blocks
// non-marginalised vertices of widths dims[], their dense symmetric H (n_dense^2: diagonal blocks and the pose-joint / joint-joint / ...
// couplings of the articulated edges), edges (block of the key-frame, point) with their 6 x 3 W, Hll, b, lambda.
extern "C" int ref_schur_solve_x(int n_blocks, const int* dims, int n_points, int n_edges, const int* edge_block, const int* edge_point, const double* W,
                                 const double* H, const double* Hll, const double* b, double lambda, double* x, double* hschur, double* bschur) {
    using namespace g2o;
    typedef BlockSolverTraits<-1, -1> T;
    std::vector<int> pidx(n_blocks), lidx(n_points);
    int nd = 0;
    for (int i = 0; i < n_blocks; ++i) { nd += dims[i]; pidx[i] = nd; }
    for (int l = 0; l < n_points; ++l) lidx[l] = 3 * (l + 1);
    BlockSolver<T> S;
    SparseBlockMatrix<MatX> hpp(pidx, pidx), hs(pidx, pidx), hll(lidx, lidx);
    SparseBlockMatrixDiagonal<MatX> dinv(lidx);
    SparseBlockMatrixCCS<MatX> hpl(pidx, lidx), hst(pidx, pidx);
    std::vector<MatX> wblocks(n_edges, MatX(6, 3));
    // Hpp: the upper block triangle; a coupling block exists where an edge joins the two vertices = where H is not identically zero
    for (int i = 0; i < n_blocks; ++i)
        for (int j = i; j < n_blocks; ++j) {
            const int r0 = hpp.rowBaseOfBlock(i), c0 = hpp.colBaseOfBlock(j);
            bool any = i == j;
            for (int r = 0; r < dims[i] && !any; ++r) for (int c = 0; c < dims[j]; ++c) any |= H[(size_t)(r0 + r) * nd + c0 + c] != 0.0;
            if (!any) continue;
            MatX* m = hpp.block(i, j, true);
            for (int r = 0; r < dims[i]; ++r) for (int c = 0; c < dims[j]; ++c) (*m)(r, c) = H[(size_t)(r0 + r) * nd + c0 + c];
            if (i == j) for (int r = 0; r < dims[i]; ++r) (*m)(r, r) += lambda;
        }
    for (int l = 0; l < n_points; ++l) {
        MatX* m = hll.block(l, l, true);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) (*m)(r, c) = Hll[(size_t)9 * l + 3 * r + c];
        for (int r = 0; r < 3; ++r) (*m)(r, r) += lambda;
    }
    for (int e = 0; e < n_edges; ++e) {
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 3; ++c) wblocks[e](r, c) = W[(size_t)18 * e + 3 * r + c];
        hpl.blockCols()[edge_point[e]].push_back(SparseBlockMatrixCCS<MatX>::RowBlock(edge_block[e], &wblocks[e]));
    }
    for (auto& col : hpl.blockCols()) std::sort(col.begin(), col.end());
    for (size_t c = 0; c < hpp.blockCols().size(); ++c) for (auto& kv : hpp.blockCols()[c]) hs.block(kv.first, (int)c, true);
    for (const auto& col : hpl.blockCols())
        for (size_t a = 0; a < col.size(); ++a)
            for (size_t c2 = a; c2 < col.size(); ++c2) hs.block(col[a].row, col[c2].row, true);
    for (size_t c = 0; c < hs.blockCols().size(); ++c)
        for (auto& kv : hs.blockCols()[c]) hst.blockCols()[kv.first].push_back(SparseBlockMatrixCCS<MatX>::RowBlock((int)c, kv.second));
    for (auto& col : hst.blockCols()) std::sort(col.begin(), col.end());
    const int nl = 3 * n_points;
    std::vector<double> coeff((size_t)nd + nl, 0.0), bcopy(b, b + nd + nl), bs((size_t)std::max(nd, 1), 0.0);
    LinearSolver<MatX> ls;
    S._Hpp = &hpp; S._Hll = &hll; S._Hschur = &hs; S._DInvSchur = &dinv; S._HplCCS = &hpl; S._HschurTransposedCCS = &hst; S._linearSolver = &ls;
    S._coefficients = coeff.data(); S._bschur = bs.data(); S._x = x; S._b = bcopy.data(); S._sizePoses = nd; S._sizeLandmarks = nl;
    for (int i = 0; i < nd + nl; ++i) x[i] = 0.0;
    const bool ok = S.solve();
    if (hschur) {
        for (size_t i = 0; i < (size_t)nd * nd; ++i) hschur[i] = 0.0;
        for (size_t c = 0; c < hs.blockCols().size(); ++c)
            for (const auto& kv : hs.blockCols()[c]) {
                const int r0 = hs.rowBaseOfBlock(kv.first), c0 = hs.colBaseOfBlock((int)c);
                for (int r = 0; r < kv.second->rows(); ++r)
                    for (int cc = 0; cc < kv.second->cols(); ++cc) {
                        hschur[(size_t)(r0 + r) * nd + c0 + cc] = (*kv.second)(r, cc);
                        hschur[(size_t)(c0 + cc) * nd + r0 + r] = (*kv.second)(r, cc);
                    }
            }
    }
    if (bschur) for (int i = 0; i < nd; ++i) bschur[i] = bs[i];
    return ok ? 1 : 0;
}
