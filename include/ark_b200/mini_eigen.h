// mini_eigen.h -- the few Eigen / OpenCV types that appear in the reference's public signatures
// (include/Avatar.h, include/AvatarOptimizer.h), for build environments WITHOUT Eigen and OpenCV (this
// container).  When <Eigen/Core> exists, ark_b200/Avatar.h includes the real headers instead and this file is
// unused.  Column-major storage and Quaternion coefficient order (x,y,z,w) follow Eigen.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

namespace Eigen {
constexpr int Dynamic = -1;
constexpr int ColMajor = 0, RowMajor = 1;
template <class T> using aligned_allocator = std::allocator<T>;

template <class S, int R, int C, int Opt = ColMajor>
class Matrix {
   public:
    typedef S Scalar;
    Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)r_ * c_, S(0)) {}
    Matrix(long r, long c) : r_(r), c_(c), d_((size_t)r * c, S(0)) {}
    explicit Matrix(long n) : r_(C == 1 ? n : (R == Dynamic ? n : R)), c_(C == 1 ? 1 : n), d_((size_t)r_ * c_, S(0)) {}
    Matrix(S x, S y, S z) : r_(R == Dynamic ? 3 : R), c_(C == Dynamic ? 1 : C), d_((size_t)r_ * c_, S(0)) { d_[0] = x; d_[1] = y; d_[2] = z; }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    S* data() { return d_.data(); }
    const S* data() const { return d_.data(); }
    void resize(long r, long c) { r_ = r; c_ = c; d_.assign((size_t)r * c, S(0)); }
    void resize(long n) { if (C == 1) resize(n, 1); else resize(R == Dynamic ? 1 : R, n); }
    S& operator()(long i, long j) { return d_[Opt == RowMajor ? i * c_ + j : j * r_ + i]; }
    const S& operator()(long i, long j) const { return d_[Opt == RowMajor ? i * c_ + j : j * r_ + i]; }
    S& operator()(long i) { return d_[i]; }
    const S& operator()(long i) const { return d_[i]; }
    S& operator[](long i) { return d_[i]; }
    const S& operator[](long i) const { return d_[i]; }
    S& x() { return d_[0]; } S& y() { return d_[1]; } S& z() { return d_[2]; }
    const S& x() const { return d_[0]; } const S& y() const { return d_[1]; } const S& z() const { return d_[2]; }
    void setZero() { d_.assign(d_.size(), S(0)); }
    void setIdentity() { setZero(); for (long i = 0; i < (r_ < c_ ? r_ : c_); ++i) (*this)(i, i) = S(1); }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Zero() { return Matrix(); }
    struct Rowwise {   // dataCloud.rowwise().mean() (demo.cpp:254)
        const Matrix& m;
        Matrix<S, R, 1> mean() const {
            Matrix<S, R, 1> out(m.rows());
            for (long i = 0; i < m.rows(); ++i) {
                S acc = S(0);
                for (long j = 0; j < m.cols(); ++j) acc += m(i, j);
                out(i) = acc / (S)m.cols();
            }
            return out;
        }
    };
    Rowwise rowwise() const { return Rowwise{*this}; }
    void setConstant(S v) { d_.assign(d_.size(), v); }
   private:
    long r_, c_;
    std::vector<S> d_;
};
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;

class Quaterniond {
   public:
    Quaterniond() { c_(3) = 1.0; }
    Quaterniond(double w, double x, double y, double z) { c_(0) = x; c_(1) = y; c_(2) = z; c_(3) = w; }  // ctor is (w,x,y,z)
    Vector4d& coeffs() { return c_; }                 // storage is (x,y,z,w)
    const Vector4d& coeffs() const { return c_; }
    double x() const { return c_(0); } double y() const { return c_(1); }
    double z() const { return c_(2); } double w() const { return c_(3); }
   private:
    Vector4d c_;
};

/** Eigen::AngleAxisd(angle, axis).toRotationMatrix() (Rodrigues), used by demo.cpp:259-261 */
class AngleAxisd {
   public:
    AngleAxisd(double angle, const Vector3d& axis) : angle_(angle), axis_(axis) {}
    Matrix3d toRotationMatrix() const;
   private:
    double angle_;
    Vector3d axis_;
};

template <class S>
struct SparseMatrix {  // placeholder for AvatarModel::jointRegressor / weights (CSC, column = vertex)
    long rows_ = 0, cols_ = 0;
    std::vector<int> outer, inner;
    std::vector<S> values;
    long rows() const { return rows_; }
    long cols() const { return cols_; }
};
}  // namespace Eigen

inline Eigen::Matrix3d Eigen::AngleAxisd::toRotationMatrix() const {
    Matrix3d R;
    const double c = __builtin_cos(angle_), s = __builtin_sin(angle_), t = 1.0 - c;
    const double x = axis_(0), y = axis_(1), z = axis_(2);
    R(0, 0) = t * x * x + c;     R(0, 1) = t * x * y - s * z; R(0, 2) = t * x * z + s * y;
    R(1, 0) = t * x * y + s * z; R(1, 1) = t * y * y + c;     R(1, 2) = t * y * z - s * x;
    R(2, 0) = t * x * z - s * y; R(2, 1) = t * y * z + s * x; R(2, 2) = t * z * z + c;
    return R;
}

#include "mini_cv.h"

