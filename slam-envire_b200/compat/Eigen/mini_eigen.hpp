// mini_eigen.hpp -- the small slice of the Eigen3 API that envire's icp/ interface touches.
//
// Used only when the real Eigen3 is not installed (this image has none): the host mirror
// (icp/icp_gpu.hpp) and the envire-mini boundary types compile against it, and oracle/_ref uses it
// to compile the reference's own icp.cpp/icp.hpp.  Fixed-size, eager evaluation, fp64 evaluation
// order as documented at each operator (it mirrors Eigen's coefficient-based products).
// Not a general linear-algebra library.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <complex>
#include <cstddef>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define MINI_EIGEN 1

namespace Eigen {

enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x12, Projective = 0x20 };

template <typename T> struct real_of { typedef T type; };
template <typename T> struct real_of<std::complex<T> > { typedef T type; };
template <typename T> inline T real_part(const T &v) { return v; }
template <typename T> inline T real_part(const std::complex<T> &v) { return v.real(); }

template <typename T, int R, int C> class Matrix;

template <typename T, int R, int C> class CommaInit {
public:
    CommaInit(Matrix<T, R, C> &m, const T &s) : m_(m), row_(0), col_(1), block_rows_(1) { m_(0, 0) = s; }
    template <int R2, int C2> CommaInit(Matrix<T, R, C> &m, const Matrix<T, R2, C2> &b) : m_(m), row_(0), col_(0), block_rows_(R2)
    {
        place(b);
    }
    CommaInit &operator,(const T &s)
    {
        if (col_ == C) { row_ += block_rows_; col_ = 0; block_rows_ = 1; }
        m_(row_, col_++) = s;
        return *this;
    }
    template <int R2, int C2> CommaInit &operator,(const Matrix<T, R2, C2> &b)
    {
        if (col_ == C) { row_ += block_rows_; col_ = 0; block_rows_ = R2; }
        place(b);
        return *this;
    }
private:
    template <int R2, int C2> void place(const Matrix<T, R2, C2> &b)
    {
        for (int r = 0; r < R2; ++r)
            for (int c = 0; c < C2; ++c) m_(row_ + r, col_ + c) = b(r, c);
        col_ += C2;
    }
    Matrix<T, R, C> &m_;
    int row_, col_, block_rows_;
};

template <typename T, int R, int C> class Matrix {
public:
    typedef T Scalar;
    typedef typename real_of<T>::type Real;
    T d[R * C]; // column-major, like Eigen's default

    Matrix() { for (int i = 0; i < R * C; ++i) d[i] = T(); }
    Matrix(const T &x, const T &y, const T &z) { static_assert(R * C == 3, "3-vector"); d[0] = x; d[1] = y; d[2] = z; }
    Matrix(const T &x, const T &y, const T &z, const T &w) { static_assert(R * C == 4, "4-vector"); d[0] = x; d[1] = y; d[2] = z; d[3] = w; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Ones() { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = T(1); return m; }
    static Matrix Random() { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = T(2.0 * rand() / (double)RAND_MAX - 1.0); return m; } // uniform in [-1, 1]
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    static Matrix UnitX() { Matrix m; m.d[0] = T(1); return m; }
    static Matrix UnitY() { Matrix m; m.d[1] = T(1); return m; }
    static Matrix UnitZ() { Matrix m; m.d[2] = T(1); return m; }
    void setZero() { *this = Matrix(); }

    int rows() const { return R; }
    int cols() const { return C; }
    T &operator()(int r, int c) { return d[c * R + r]; }
    const T &operator()(int r, int c) const { return d[c * R + r]; }
    T &operator()(int i) { return d[i]; }
    const T &operator()(int i) const { return d[i]; }
    T &operator[](size_t i) { return d[i]; }
    const T &operator[](size_t i) const { return d[i]; }
    T &x() { return d[0]; } T &y() { return d[1]; } T &z() { return d[2]; }
    const T &x() const { return d[0]; } const T &y() const { return d[1]; } const T &z() const { return d[2]; }
    const T *data() const { return d; }
    T *data() { return d; }

    CommaInit<T, R, C> operator<<(const T &s) { return CommaInit<T, R, C>(*this, s); }
    template <int R2, int C2> CommaInit<T, R, C> operator<<(const Matrix<T, R2, C2> &b) { return CommaInit<T, R, C>(*this, b); }

    Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c); return t; }
    T trace() const { T s = (*this)(0, 0); for (int i = 1; i < (R < C ? R : C); ++i) s += (*this)(i, i); return s; }
    Matrix<T, R, 1> col(int c) const { Matrix<T, R, 1> v; for (int r = 0; r < R; ++r) v(r) = (*this)(r, c); return v; }
    Matrix<T, 1, C> row(int r) const { Matrix<T, 1, C> v; for (int c = 0; c < C; ++c) v(c) = (*this)(r, c); return v; }
    Matrix<Real, R, C> real() const { Matrix<Real, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = real_part(d[i]); return m; }
    T maxCoeff() const { T m = d[0]; for (int i = 1; i < R * C; ++i) if (d[i] > m) m = d[i]; return m; }
    T dot(const Matrix &o) const { T s = d[0] * o.d[0]; for (int i = 1; i < R * C; ++i) s += d[i] * o.d[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    Matrix normalized() const { return *this / norm(); }
    void normalize() { *this = normalized(); }
    Matrix cross(const Matrix &o) const
    {
        static_assert(R * C == 3, "cross needs 3-vectors");
        return Matrix(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]);
    }

    Matrix operator+(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] + o.d[i]; return m; }
    Matrix operator-(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
    Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }
    Matrix operator*(const T &s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
    Matrix operator/(const T &s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] / s; return m; }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; ++i) d[i] += o.d[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; ++i) d[i] -= o.d[i]; return *this; }
    Matrix &operator*=(const T &s) { for (int i = 0; i < R * C; ++i) d[i] *= s; return *this; }
    // coefficient-based product: res(r,c) = ((a(r,0)*b(0,c) + a(r,1)*b(1,c)) + a(r,2)*b(2,c)) + ...
    template <int C2> Matrix<T, R, C2> operator*(const Matrix<T, C, C2> &o) const
    {
        Matrix<T, R, C2> m;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C2; ++c) {
                T s = (*this)(r, 0) * o(0, c);
                for (int k = 1; k < C; ++k) s += (*this)(r, k) * o(k, c);
                m(r, c) = s;
            }
        return m;
    }
    bool operator==(const Matrix &o) const { for (int i = 0; i < R * C; ++i) if (!(d[i] == o.d[i])) return false; return true; }
};
template <typename T, int R, int C> inline Matrix<T, R, C> operator*(const T &s, const Matrix<T, R, C> &m) { return m * s; }

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 1, 3> RowVector3d;

inline double det3(const Matrix3d &m)
{
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
}
inline Matrix3d inverse3(const Matrix3d &m)
{
    const double d = det3(m);
    Matrix3d r;
    r(0, 0) = (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) / d;
    r(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) / d;
    r(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) / d;
    r(1, 0) = (m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2)) / d;
    r(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) / d;
    r(1, 2) = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) / d;
    r(2, 0) = (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0)) / d;
    r(2, 1) = (m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1)) / d;
    r(2, 2) = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) / d;
    return r;
}

class AngleAxisd;
class Affine3d;

class Quaterniond {
public:
    Quaterniond() : w_(1), v_(0, 0, 0) {}
    Quaterniond(double w, double x, double y, double z) : w_(w), v_(x, y, z) {}
    static Quaterniond Identity() { return Quaterniond(); }
    double w() const { return w_; } double x() const { return v_[0]; } double y() const { return v_[1]; } double z() const { return v_[2]; }
    const Vector3d &vec() const { return v_; }
    // Eigen's _transformVector: v + w*(2 u x v) + u x (2 u x v)
    Vector3d operator*(const Vector3d &v) const
    {
        Vector3d uv = v_.cross(v);
        uv += uv;
        return v + uv * w_ + v_.cross(uv);
    }
    Matrix3d toRotationMatrix() const
    {
        Matrix3d res;
        const double tx = 2.0 * x(), ty = 2.0 * y(), tz = 2.0 * z();
        const double twx = tx * w(), twy = ty * w(), twz = tz * w();
        const double txx = tx * x(), txy = ty * x(), txz = tz * x();
        const double tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res(0, 0) = 1.0 - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
        res(1, 0) = txy + twz; res(1, 1) = 1.0 - (txx + tzz); res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = 1.0 - (txx + tyy);
        return res;
    }
private:
    double w_;
    Vector3d v_;
};

class AngleAxisd {
public:
    AngleAxisd(double angle, const Vector3d &axis) : angle_(angle), axis_(axis) {}
    explicit AngleAxisd(const Matrix3d &R)
    {
        const double c = (R.trace() - 1.0) / 2.0;
        angle_ = std::acos(c > 1 ? 1 : (c < -1 ? -1 : c));
        Vector3d a(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
        axis_ = a.norm() > 0 ? a.normalized() : Vector3d::UnitX();
    }
    double angle() const { return angle_; }
    const Vector3d &axis() const { return axis_; }
    Matrix3d toRotationMatrix() const
    {
        Matrix3d res;
        const double s = std::sin(angle_), c = std::cos(angle_);
        const Vector3d sin_axis = axis_ * s, cos1_axis = axis_ * (1.0 - c);
        double tmp;
        tmp = cos1_axis.x() * axis_.y(); res(0, 1) = tmp - sin_axis.z(); res(1, 0) = tmp + sin_axis.z();
        tmp = cos1_axis.x() * axis_.z(); res(0, 2) = tmp + sin_axis.y(); res(2, 0) = tmp - sin_axis.y();
        tmp = cos1_axis.y() * axis_.z(); res(1, 2) = tmp - sin_axis.x(); res(2, 1) = tmp + sin_axis.x();
        res(0, 0) = cos1_axis.x() * axis_.x() + c; res(1, 1) = cos1_axis.y() * axis_.y() + c; res(2, 2) = cos1_axis.z() * axis_.z() + c;
        return res;
    }
    Vector3d operator*(const Vector3d &v) const { return toRotationMatrix() * v; }
private:
    double angle_;
    Vector3d axis_;
};

class Translation3d {
public:
    Translation3d(double x, double y, double z) : t_(x, y, z) {}
    explicit Translation3d(const Vector3d &t) : t_(t) {}
    const Vector3d &vector() const { return t_; }
private:
    Vector3d t_;
};

// Transform<double,3,Affine>: [linear | translation], last row (0 0 0 1)
class Affine3d {
public:
    Affine3d() {}
    Affine3d(const Matrix3d &L, const Vector3d &t) : L_(L), t_(t) {}
    explicit Affine3d(const Matrix4d &m)
    {
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) L_(r, c) = m(r, c); t_(r) = m(r, 3); }
    }
    Affine3d(const Translation3d &t) : L_(Matrix3d::Identity()), t_(t.vector()) {}
    Affine3d(const AngleAxisd &a) : L_(a.toRotationMatrix()) {}
    Affine3d(const Quaterniond &q) : L_(q.toRotationMatrix()) {}
    static Affine3d Identity() { return Affine3d(Matrix3d::Identity(), Vector3d::Zero()); }
    void setIdentity() { *this = Identity(); }
    Matrix3d &linear() { return L_; }
    const Matrix3d &linear() const { return L_; }
    Vector3d &translation() { return t_; }
    const Vector3d &translation() const { return t_; }
    Matrix4d matrix() const
    {
        Matrix4d m;
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m(r, c) = L_(r, c); m(r, 3) = t_(r); }
        m(3, 3) = 1.0;
        return m;
    }
    // p = linear*v + translation
    Vector3d operator*(const Vector3d &v) const { return L_ * v + t_; }
    // res.linear = a.linear*b.linear; res.translation = a.linear*b.translation + a.translation
    Affine3d operator*(const Affine3d &o) const { return Affine3d(L_ * o.L_, L_ * o.t_ + t_); }
    Affine3d operator*(const Translation3d &t) const { return *this * Affine3d(t); }
    Affine3d operator*(const AngleAxisd &a) const { return *this * Affine3d(a); }
    Affine3d inverse(TransformTraits traits = Affine) const
    {
        const Matrix3d Li = (traits == Isometry) ? L_.transpose() : inverse3(L_);
        return Affine3d(Li, -(Li * t_));
    }
    // orthogonal polar factor of linear() (Eigen: U*V^T of a JacobiSVD, sign-fixed); Newton iteration X <- (X + X^-T)/2
    Matrix3d rotation() const
    {
        Matrix3d X = L_;
        for (int it = 0; it < 100; ++it) {
            const Matrix3d Y = inverse3(X).transpose();
            double delta = 0;
            for (int i = 0; i < 9; ++i) {
                const double nx = 0.5 * (X.d[i] + Y.d[i]);
                delta = std::fmax(delta, std::fabs(nx - X.d[i]));
                X.d[i] = nx;
            }
            if (delta < 1e-16) break;
        }
        return X;
    }
private:
    Matrix3d L_ = Matrix3d::Identity();
    Vector3d t_;
};
typedef Affine3d Isometry3d;

inline Affine3d operator*(const Translation3d &t, const Quaterniond &q) { return Affine3d(q.toRotationMatrix(), t.vector()); }
inline Affine3d operator*(const Translation3d &t, const AngleAxisd &a) { return Affine3d(a.toRotationMatrix(), t.vector()); }
inline Affine3d operator*(const Translation3d &t, const Affine3d &a) { return Affine3d(t) * a; }
inline Affine3d operator*(const AngleAxisd &a, const Affine3d &b) { return Affine3d(a) * b; }

template <typename T> using aligned_allocator = std::allocator<T>; // plain new is aligned enough for these doubles

} // namespace Eigen
