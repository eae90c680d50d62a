// oracle/shim/Eigen/src/shim_ops.h -- TEST INFRASTRUCTURE (see ../Core): operators, decompositions, sparse, geometry.
#pragma once

namespace Eigen {

// ---------------------------------------------------------------------------------------------------------------
// element-wise operators and products (eager, sums in index order)
// ---------------------------------------------------------------------------------------------------------------
template <class A, class B, class T, class F>
inline Matrix<T, Dynamic, Dynamic, ColMajor> zip(const Base<A, T>& a, const Base<B, T>& b, F f) {
    const A& x = a.d();
    const B& y = b.d();
    Matrix<T, Dynamic, Dynamic, ColMajor> o(x.rows(), x.cols());
    if (x.rows() == y.rows() && x.cols() == y.cols()) {
        for (Index j = 0; j < x.cols(); ++j)
            for (Index i = 0; i < x.rows(); ++i) o.ref(i, j) = f(x.coeff(i, j), y.coeff(i, j));
    } else {   // a column vector against a row vector
        assert(x.rows() == y.cols() && x.cols() == y.rows() && (x.rows() == 1 || x.cols() == 1));
        for (Index j = 0; j < x.cols(); ++j)
            for (Index i = 0; i < x.rows(); ++i) o.ref(i, j) = f(x.coeff(i, j), y.coeff(j, i));
    }
    return o;
}
template <class A, class B, class T>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator+(const Base<A, T>& a, const Base<B, T>& b) {
    return zip(a, b, [](T p, T q) { return p + q; });
}
template <class A, class B, class T>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator-(const Base<A, T>& a, const Base<B, T>& b) {
    return zip(a, b, [](T p, T q) { return p - q; });
}
template <class A, class T>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator-(const Base<A, T>& a) {
    Matrix<T, Dynamic, Dynamic, ColMajor> o(a.d().rows(), a.d().cols());
    for (Index j = 0; j < o.cols(); ++j)
        for (Index i = 0; i < o.rows(); ++i) o.ref(i, j) = -a.d().coeff(i, j);
    return o;
}
template <class A, class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator*(const Base<A, T>& a, S s) {
    Matrix<T, Dynamic, Dynamic, ColMajor> o(a.d().rows(), a.d().cols());
    for (Index j = 0; j < o.cols(); ++j)
        for (Index i = 0; i < o.rows(); ++i) o.ref(i, j) = a.d().coeff(i, j) * (T)s;
    return o;
}
template <class A, class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator*(S s, const Base<A, T>& a) {
    Matrix<T, Dynamic, Dynamic, ColMajor> o(a.d().rows(), a.d().cols());
    for (Index j = 0; j < o.cols(); ++j)
        for (Index i = 0; i < o.rows(); ++i) o.ref(i, j) = (T)s * a.d().coeff(i, j);
    return o;
}
template <class A, class T, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator/(const Base<A, T>& a, S s) {
    Matrix<T, Dynamic, Dynamic, ColMajor> o(a.d().rows(), a.d().cols());
    for (Index j = 0; j < o.cols(); ++j)
        for (Index i = 0; i < o.rows(); ++i) o.ref(i, j) = a.d().coeff(i, j) / (T)s;
    return o;
}
template <class A, class B, class T>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator*(const Base<A, T>& a, const Base<B, T>& b) {
    const A& x = a.d();
    const B& y = b.d();
    assert(x.cols() == y.rows());
    Matrix<T, Dynamic, Dynamic, ColMajor> o(x.rows(), y.cols());
    for (Index j = 0; j < y.cols(); ++j)
        for (Index i = 0; i < x.rows(); ++i) {
            T s = 0;
            for (Index k = 0; k < x.cols(); ++k) s += x.coeff(i, k) * y.coeff(k, j);
            o.ref(i, j) = s;
        }
    return o;
}
// v *= M (matrix on the right), as GaussianMixture::sample writes it
template <class D, class T, class B>
inline D& operator*=(WBase<D, T>& a, const Base<B, T>& b) {
    Matrix<T, Dynamic, Dynamic, ColMajor> t = a * b;
    a.w() = t;
    return a.w();
}

template <class D, class T> typename Base<D, T>::Plain Base<D, T>::eval() const { return Plain(*this); }
template <class D, class T> typename Base<D, T>::Plain Base<D, T>::transpose() const {
    Plain o(d().cols(), d().rows());
    for (Index j = 0; j < d().cols(); ++j)
        for (Index i = 0; i < d().rows(); ++i) o.ref(j, i) = d().coeff(i, j);
    return o;
}
template <class D, class T> typename Base<D, T>::Plain Base<D, T>::homogeneous() const {
    const Index n = size();
    Plain o(d().cols() == 1 ? n + 1 : 1, d().cols() == 1 ? 1 : n + 1);
    for (Index k = 0; k < n; ++k) o.wlin(k) = lin(k);
    o.wlin(n) = T(1);
    return o;
}
template <class D, class T> typename Base<D, T>::Plain Base<D, T>::normalized() const {
    const T n = norm();
    return n > T(0) ? Plain(*this / n) : Plain(*this);
}
template <class D, class T> T Base<D, T>::dot(const Plain& o) const {
    T s = 0;
    for (Index k = 0; k < size(); ++k) s += lin(k) * o.lin(k);
    return s;
}
template <class D, class T> typename Base<D, T>::Plain Base<D, T>::cross(const Plain& o) const {
    Plain r(3, 1);
    r.ref(0, 0) = lin(1) * o.lin(2) - lin(2) * o.lin(1);
    r.ref(1, 0) = lin(2) * o.lin(0) - lin(0) * o.lin(2);
    r.ref(2, 0) = lin(0) * o.lin(1) - lin(1) * o.lin(0);
    return r;
}

// LU with partial pivoting (Eigen's default for inverse() / determinant() of dynamic matrices)
template <class T>
struct PartialLU {
    Matrix<T, Dynamic, Dynamic, ColMajor> lu;
    std::vector<Index> perm;
    int sign = 1;
    template <class M> explicit PartialLU(const M& m) : lu(m), perm((size_t)m.rows()) {
        const Index n = lu.rows();
        for (Index i = 0; i < n; ++i) perm[(size_t)i] = i;
        for (Index k = 0; k < n; ++k) {
            Index piv = k;
            for (Index i = k + 1; i < n; ++i)
                if (std::fabs(lu.coeff(i, k)) > std::fabs(lu.coeff(piv, k))) piv = i;
            if (piv != k) {
                for (Index j = 0; j < n; ++j) std::swap(lu.ref(k, j), lu.ref(piv, j));
                std::swap(perm[(size_t)k], perm[(size_t)piv]);
                sign = -sign;
            }
            for (Index i = k + 1; i < n; ++i) {
                lu.ref(i, k) /= lu.coeff(k, k);
                for (Index j = k + 1; j < n; ++j) lu.ref(i, j) -= lu.coeff(i, k) * lu.coeff(k, j);
            }
        }
    }
};
template <class D, class T> T Base<D, T>::determinant() const {
    PartialLU<T> f(d());
    T det = (T)f.sign;
    for (Index i = 0; i < f.lu.rows(); ++i) det *= f.lu.coeff(i, i);
    return det;
}
template <class D, class T> typename Base<D, T>::Plain Base<D, T>::inverse() const {
    PartialLU<T> f(d());
    const Index n = f.lu.rows();
    Plain inv(n, n);
    for (Index c = 0; c < n; ++c) {
        std::vector<T> y((size_t)n);
        for (Index i = 0; i < n; ++i) {   // L y = P e_c
            T s = (f.perm[(size_t)i] == c) ? T(1) : T(0);
            for (Index k = 0; k < i; ++k) s -= f.lu.coeff(i, k) * y[(size_t)k];
            y[(size_t)i] = s;
        }
        for (Index i = n - 1; i >= 0; --i) {   // U x = y
            T s = y[(size_t)i];
            for (Index k = i + 1; k < n; ++k) s -= f.lu.coeff(i, k) * inv.coeff(k, c);
            inv.ref(i, c) = s / f.lu.coeff(i, i);
        }
    }
    return inv;
}

// ---------------------------------------------------------------------------------------------------------------
// LLT and TriangularView
// ---------------------------------------------------------------------------------------------------------------
template <class M, int UpLo = Lower>
class LLT {
    typedef typename M::Scalar T;
    Matrix<T, Dynamic, Dynamic, ColMajor> L;
    ComputationInfo status = Success;
public:
    template <class O, class S> explicit LLT(const Base<O, S>& a) : L(a.d().rows(), a.d().cols()) {
        const Index n = L.rows();
        for (Index j = 0; j < n; ++j) {
            T s = a.d().coeff(j, j);
            for (Index k = 0; k < j; ++k) s -= L.coeff(j, k) * L.coeff(j, k);
            if (!(s > T(0))) { status = NumericalIssue; return; }
            const T ljj = std::sqrt(s);
            L.ref(j, j) = ljj;
            for (Index i = j + 1; i < n; ++i) {
                T t = a.d().coeff(i, j);
                for (Index k = 0; k < j; ++k) t -= L.coeff(i, k) * L.coeff(j, k);
                L.ref(i, j) = t / ljj;
            }
        }
    }
    ComputationInfo info() const { return status; }
    Matrix<T, Dynamic, Dynamic, ColMajor> matrixL() const { return L; }
};

template <class M, int Mode>
class TriangularView {
    typedef typename M::Scalar T;
    const M& m;
    bool transposed;
public:
    explicit TriangularView(const M& m_, bool tr = false) : m(m_), transposed(tr) {}
    TriangularView transpose() const { return TriangularView(m, !transposed); }
    T at(Index i, Index j) const {
        const Index a = transposed ? j : i, b = transposed ? i : j;
        const bool in = (Mode == Lower) ? (b <= a) : (b >= a);
        return in ? m.coeff(a, b) : T(0);
    }
    Index rows() const { return transposed ? m.cols() : m.rows(); }
    Index cols() const { return transposed ? m.rows() : m.cols(); }
    template <class B> Matrix<T, Dynamic, Dynamic, ColMajor> operator*(const Base<B, T>& b) const {
        Matrix<T, Dynamic, Dynamic, ColMajor> o(rows(), b.d().cols());
        for (Index j = 0; j < b.d().cols(); ++j)
            for (Index i = 0; i < rows(); ++i) {
                T s = 0;
                for (Index k = 0; k < cols(); ++k) s += at(i, k) * b.d().coeff(k, j);
                o.ref(i, j) = s;
            }
        return o;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// sparse (compressed columns); dense x sparse products
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct SparseCol;
template <class T>
struct Triplet {
    Index r, c;
    T v;
    Triplet(Index r_, Index c_, T v_) : r(r_), c(c_), v(v_) {}
    Index row() const { return r; }
    Index col() const { return c; }
    T value() const { return v; }
};
template <class T, int Opt>
class SparseMatrix {
public:
    typedef T Scalar;
    Index r = 0, c = 0;
    std::vector<Index> start, idx;   // start[c + 1], row index per entry
    std::vector<T> val;
    SparseMatrix() : start(1, 0) {}
    SparseMatrix(Index r_, Index c_) : r(r_), c(c_), start((size_t)c_ + 1, 0) {}
    Index rows() const { return r; }
    Index cols() const { return c; }
    Index nonZeros() const { return (Index)val.size(); }
    void resize(Index r_, Index c_) { *this = SparseMatrix(r_, c_); }
    template <class It> void setFromTriplets(It b, It e) {   // entries of one column keep their order; duplicates are summed by the product
        std::vector<Index> cnt((size_t)c + 1, 0);
        for (It i = b; i != e; ++i) ++cnt[(size_t)i->col() + 1];
        for (Index j = 0; j < c; ++j) cnt[(size_t)j + 1] += cnt[(size_t)j];
        start = cnt;
        idx.assign((size_t)cnt[(size_t)c], 0);
        val.assign((size_t)cnt[(size_t)c], T(0));
        std::vector<Index> fill(cnt.begin(), cnt.end() - 1);
        for (It i = b; i != e; ++i) {
            const Index k = fill[(size_t)i->col()]++;
            idx[(size_t)k] = i->row();
            val[(size_t)k] = i->value();
        }
    }
    void makeCompressed() {}
    Index outerSize() const { return c; }
    template <class X> void reserve(const X&) {}
    T& insert(Index i, Index j) {   // keeps the row indices of a column ascending
        Index k = start[(size_t)j];
        while (k < start[(size_t)j + 1] && idx[(size_t)k] < i) ++k;
        idx.insert(idx.begin() + k, i);
        val.insert(val.begin() + k, T(0));
        for (Index q = j + 1; q <= c; ++q) ++start[(size_t)q];
        return val[(size_t)k];
    }
    class InnerIterator {
        const SparseMatrix& m;
        Index k, e, outer;
    public:
        InnerIterator(const SparseMatrix& m_, Index j) : m(m_), k(m_.start[(size_t)j]), e(m_.start[(size_t)j + 1]), outer(j) {}
        operator bool() const { return k < e; }
        InnerIterator& operator++() { ++k; return *this; }
        Index row() const { return m.idx[(size_t)k]; }
        Index col() const { return outer; }
        Index index() const { return m.idx[(size_t)k]; }
        T value() const { return m.val[(size_t)k]; }
    };
    SparseCol<T> col(Index j) const { return SparseCol<T>{&idx, &val, start[(size_t)j], start[(size_t)j + 1], r}; }
    T coeff(Index i, Index j) const {
        T s = 0;
        for (Index k = start[(size_t)j]; k < start[(size_t)j + 1]; ++k)
            if (idx[(size_t)k] == i) s += val[(size_t)k];
        return s;
    }
};
template <class T>
struct SparseCol {
    const std::vector<Index>* idx;
    const std::vector<T>* val;
    Index b, e, n;
};
template <class A, class T>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator*(const Base<A, T>& a, const SparseCol<T>& s) {
    assert(a.d().cols() == s.n);
    Matrix<T, Dynamic, Dynamic, ColMajor> o(a.d().rows(), 1);
    for (Index k = s.b; k < s.e; ++k)
        for (Index i = 0; i < o.rows(); ++i) o.ref(i, 0) += a.d().coeff(i, (*s.idx)[(size_t)k]) * (*s.val)[(size_t)k];
    return o;
}
template <class D, class T> SparseMatrix<T, ColMajor> Base<D, T>::sparseView() const {
    SparseMatrix<T, ColMajor> o(d().rows(), d().cols());
    for (Index j = 0; j < d().cols(); ++j) {
        for (Index i = 0; i < d().rows(); ++i)
            if (d().coeff(i, j) != T(0)) {
                o.idx.push_back(i);
                o.val.push_back(d().coeff(i, j));
            }
        o.start[(size_t)j + 1] = (Index)o.idx.size();
    }
    return o;
}
template <class A, class T, int Opt>
inline Matrix<T, Dynamic, Dynamic, ColMajor> operator*(const Base<A, T>& a, const SparseMatrix<T, Opt>& s) {
    assert(a.d().cols() == s.rows());
    Matrix<T, Dynamic, Dynamic, ColMajor> o(a.d().rows(), s.cols());
    for (Index j = 0; j < s.cols(); ++j)
        for (Index k = s.start[(size_t)j]; k < s.start[(size_t)j + 1]; ++k) {
            const Index row = s.idx[(size_t)k];
            const T v = s.val[(size_t)k];
            for (Index i = 0; i < o.rows(); ++i) o.ref(i, j) += a.d().coeff(i, row) * v;
        }
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// geometry: Quaternion, AngleAxis (Eigen's formulas: Quaternion.h, AngleAxis.h)
// ---------------------------------------------------------------------------------------------------------------
template <class T> class AngleAxis;
template <class T>
class Quaternion {
public:
    typedef T Scalar;
    struct Coeffs {                       // x, y, z, w (Eigen's coefficient order)
        T v[4] = {0, 0, 0, 1};
        T& operator()(Index i) { return v[i]; }
        T operator()(Index i) const { return v[i]; }
        T& operator[](Index i) { return v[i]; }
        T operator[](Index i) const { return v[i]; }
        T* data() { return v; }
        const T* data() const { return v; }
    } cf;
    T (&c)[4] = cf.v;
    Coeffs& coeffs() { return cf; }
    const Coeffs& coeffs() const { return cf; }
    Quaternion() {}
    Quaternion(const Quaternion& o) : cf(o.cf) {}
    Quaternion& operator=(const Quaternion& o) { cf = o.cf; return *this; }
    Quaternion& operator=(const AngleAxis<T>& aa) { return *this = Quaternion(aa); }
    Quaternion(T w_, T x_, T y_, T z_) { c[0] = x_; c[1] = y_; c[2] = z_; c[3] = w_; }
    template <class O> explicit Quaternion(const Base<O, T>& m) { *this = m; }
    explicit Quaternion(const AngleAxis<T>& aa);
    T x() const { return c[0]; }
    T y() const { return c[1]; }
    T z() const { return c[2]; }
    T w() const { return c[3]; }
    T& x() { return c[0]; }
    T& y() { return c[1]; }
    T& z() { return c[2]; }
    T& w() { return c[3]; }
    Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(c[0], c[1], c[2]); }
    Quaternion operator*(const Quaternion& b) const {
        const Quaternion& a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    T norm() const { return std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]); }
    void normalize() { const T n = norm(); for (T& v : c) v /= n; }
    Quaternion normalized() const { Quaternion q = *this; q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(c[3], -c[0], -c[1], -c[2]); }
    Quaternion inverse() const { return conjugate(); }
    Matrix<T, 3, 3> toRotationMatrix() const {
        Matrix<T, 3, 3> res;
        const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
        const T twx = tx * w(), twy = ty * w(), twz = tz * w();
        const T txx = tx * x(), txy = ty * x(), txz = tz * x();
        const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res.ref(0, 0) = T(1) - (tyy + tzz); res.ref(0, 1) = txy - twz; res.ref(0, 2) = txz + twy;
        res.ref(1, 0) = txy + twz; res.ref(1, 1) = T(1) - (txx + tzz); res.ref(1, 2) = tyz - twx;
        res.ref(2, 0) = txz - twy; res.ref(2, 1) = tyz + twx; res.ref(2, 2) = T(1) - (txx + tyy);
        return res;
    }
    // rotation matrix -> quaternion (Eigen: quaternionbase_assign_impl<Other, 3, 3>)
    template <class O> Quaternion& operator=(const Base<O, T>& mb) {
        const O& m = mb.d();
        T t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
        if (t > T(0)) {
            t = std::sqrt(t + T(1));
            w() = T(0.5) * t;
            t = T(0.5) / t;
            x() = (m.coeff(2, 1) - m.coeff(1, 2)) * t;
            y() = (m.coeff(0, 2) - m.coeff(2, 0)) * t;
            z() = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
        } else {
            Index i = 0;
            if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
            if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
            const Index j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + T(1));
            c[i] = T(0.5) * t;
            t = T(0.5) / t;
            w() = (m.coeff(k, j) - m.coeff(j, k)) * t;
            c[j] = (m.coeff(j, i) + m.coeff(i, j)) * t;
            c[k] = (m.coeff(k, i) + m.coeff(i, k)) * t;
        }
        return *this;
    }
    // Eigen: QuaternionBase::setFromTwoVectors
    template <class A, class B> static Quaternion FromTwoVectors(const Base<A, T>& a, const Base<B, T>& b) {
        Matrix<T, 3, 1> v0 = a.normalized(), v1 = b.normalized();
        const T cth = v1.dot(v0);
        Quaternion q;
        if (cth < T(-1) + T(1e-12)) {   // nearly opposite: any axis orthogonal to v0 (Eigen takes it from an SVD)
            Matrix<T, 3, 1> e(T(1), T(0), T(0));
            if (std::fabs(v0.lin(0)) > T(0.9)) e = Matrix<T, 3, 1>(T(0), T(1), T(0));
            Matrix<T, 3, 1> axis = Matrix<T, 3, 1>(v0.cross(e)).normalized();
            const T w2 = (T(1) + cth) * T(0.5);
            q.w() = std::sqrt(w2);
            const T s = std::sqrt(T(1) - w2);
            q.x() = axis.lin(0) * s; q.y() = axis.lin(1) * s; q.z() = axis.lin(2) * s;
            return q;
        }
        Matrix<T, 3, 1> axis = v0.cross(v1);
        const T s = std::sqrt((T(1) + cth) * T(2)), invs = T(1) / s;
        q.x() = axis.lin(0) * invs; q.y() = axis.lin(1) * invs; q.z() = axis.lin(2) * invs;
        q.w() = s * T(0.5);
        return q;
    }
};
// Map<Quaternion>, Map<const Quaternion>: four doubles in x, y, z, w order
template <class T>
class Map<Quaternion<T>> {
    T* p;
public:
    explicit Map(T* ptr) : p(ptr) {}
    operator Quaternion<T>() const { return Quaternion<T>(p[3], p[0], p[1], p[2]); }
    Map& operator=(const Quaternion<T>& q) { for (int i = 0; i < 4; ++i) p[i] = q.c[i]; return *this; }
    Map& operator=(const Map& o) { for (int i = 0; i < 4; ++i) p[i] = o.p[i]; return *this; }
    Map& operator=(const Map<const Quaternion<T>>& o);
    T x() const { return p[0]; }
    T y() const { return p[1]; }
    T z() const { return p[2]; }
    T w() const { return p[3]; }
    Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(p[0], p[1], p[2]); }
    Matrix<T, 3, 3> toRotationMatrix() const { return Quaternion<T>(*this).toRotationMatrix(); }
};
template <class T>
class Map<const Quaternion<T>> {
    const T* p;
public:
    explicit Map(const T* ptr) : p(ptr) {}
    operator Quaternion<T>() const { return Quaternion<T>(p[3], p[0], p[1], p[2]); }
    const T* raw() const { return p; }
    T x() const { return p[0]; }
    T y() const { return p[1]; }
    T z() const { return p[2]; }
    T w() const { return p[3]; }
    Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(p[0], p[1], p[2]); }
    Matrix<T, 3, 3> toRotationMatrix() const { return Quaternion<T>(*this).toRotationMatrix(); }
};
template <class T>
Map<Quaternion<T>>& Map<Quaternion<T>>::operator=(const Map<const Quaternion<T>>& o) {
    for (int i = 0; i < 4; ++i) p[i] = o.raw()[i];
    return *this;
}
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class T>
class AngleAxis {
    Matrix<T, 3, 1> ax;
    T ang = 0;
public:
    AngleAxis() {}
    template <class O> AngleAxis(T angle, const Base<O, T>& axis) : ax(axis), ang(angle) {}
    explicit AngleAxis(const Quaternion<T>& q) { *this = q; }
    template <class O> explicit AngleAxis(const Base<O, T>& m) { fromRotationMatrix(m); }
    T angle() const { return ang; }
    T& angle() { return ang; }
    const Matrix<T, 3, 1>& axis() const { return ax; }
    Matrix<T, 3, 1>& axis() { return ax; }
    // Eigen: AngleAxis::operator=(QuaternionBase)
    AngleAxis& operator=(const Quaternion<T>& q) {
        T n = q.vec().norm();
        if (n < std::numeric_limits<T>::epsilon()) n = q.vec().stableNorm();
        if (n != T(0)) {
            ang = T(2) * std::atan2(n, std::fabs(q.w()));
            if (q.w() < T(0)) n = -n;
            ax = q.vec() / n;
        } else {
            ang = T(0);
            ax = Matrix<T, 3, 1>(T(1), T(0), T(0));
        }
        return *this;
    }
    template <class O> AngleAxis& fromRotationMatrix(const Base<O, T>& m) {
        Quaternion<T> q;
        q = m;
        return *this = q;
    }
    Matrix<T, 3, 3> toRotationMatrix() const {   // Eigen: AngleAxis::toRotationMatrix
        Matrix<T, 3, 3> res;
        const T s = std::sin(ang), c = std::cos(ang);
        const Matrix<T, 3, 1> sa = ax * s, ca = ax * (T(1) - c);
        T tmp = ca.lin(0) * ax.lin(1);
        res.ref(0, 1) = tmp - sa.lin(2);
        res.ref(1, 0) = tmp + sa.lin(2);
        tmp = ca.lin(0) * ax.lin(2);
        res.ref(0, 2) = tmp + sa.lin(1);
        res.ref(2, 0) = tmp - sa.lin(1);
        tmp = ca.lin(1) * ax.lin(2);
        res.ref(1, 2) = tmp - sa.lin(0);
        res.ref(2, 1) = tmp + sa.lin(0);
        res.ref(0, 0) = ca.lin(0) * ax.lin(0) + c;
        res.ref(1, 1) = ca.lin(1) * ax.lin(1) + c;
        res.ref(2, 2) = ca.lin(2) * ax.lin(2) + c;
        return res;
    }
    Quaternion<T> operator*(const AngleAxis& o) const { return Quaternion<T>(*this) * Quaternion<T>(o); }
};
typedef AngleAxis<double> AngleAxisd;

template <class T>
Quaternion<T>::Quaternion(const AngleAxis<T>& aa) {
    const T ha = T(0.5) * aa.angle(), s = std::sin(ha);
    c[3] = std::cos(ha);
    c[0] = s * aa.axis().lin(0);
    c[1] = s * aa.axis().lin(1);
    c[2] = s * aa.axis().lin(2);
}

}  // namespace Eigen
