// TEST INFRASTRUCTURE — the iDynTree value types the MPC sources name (rotation, position, transform, twist, fixed vectors)
// as plain FP64 containers with iDynTree's documented conventions: row-major 3x3 rotation, RPY = R_z(y) R_y(p) R_x(r),
// adjoint transform [R, S(p) R; 0, R].  The kinematics engine itself (KinDynComputations) is an empty class: the MPC path
// never calls it, it only reads Robot's getters, whose values the checker supplies as data (oracle/ref_mpc_shim.cpp).
#pragma once
#include <Eigen/Dense>
#include <cmath>
#include <string>
#include <vector>

namespace iDynTree {
typedef std::ptrdiff_t FrameIndex;

template <int N> class VectorFixSize
{
public:
    VectorFixSize() { for (int i = 0; i < N; ++i) m_d[i] = 0.0; }
    double& operator()(int i) { return m_d[i]; }
    double operator()(int i) const { return m_d[i]; }
    double* data() { return m_d; }
    const double* data() const { return m_d; }
    void zero() { for (int i = 0; i < N; ++i) m_d[i] = 0.0; }
    double m_d[N];
};
typedef VectorFixSize<3> Vector3;
typedef VectorFixSize<6> Vector6;
class Position : public Vector3 {};
class Direction : public Vector3 {};
class AngularMotionVector3 : public Vector3 {};
class LinearMotionVector3 : public Vector3 {};

template <int R, int C> class MatrixFixSize
{
public:
    MatrixFixSize() { for (int i = 0; i < R * C; ++i) m_d[i] = 0.0; }
    double& operator()(int i, int j) { return m_d[i * C + j]; }
    double operator()(int i, int j) const { return m_d[i * C + j]; }
    double m_d[R * C];       // row-major, as in iDynTree
};
typedef MatrixFixSize<3, 3> Matrix3x3;
typedef MatrixFixSize<6, 6> Matrix6x6;

class Rotation : public Matrix3x3
{
public:
    Rotation() { (*this)(0, 0) = (*this)(1, 1) = (*this)(2, 2) = 1.0; }
    static Rotation Identity() { return Rotation(); }
    static Rotation RPY(double r, double p, double y)
    {
        Rotation R;
        const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
        R(0, 0) = cp * cy; R(0, 1) = cy * sp * sr - sy * cr; R(0, 2) = cy * sp * cr + sy * sr;
        R(1, 0) = cp * sy; R(1, 1) = sy * sp * sr + cy * cr; R(1, 2) = sy * sp * cr - cy * sr;
        R(2, 0) = -sp;     R(2, 1) = cp * sr;                R(2, 2) = cp * cr;
        return R;
    }
    Rotation inverse() const
    {
        Rotation T;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T(i, j) = (*this)(j, i);
        return T;
    }
    Vector3 asRPY() const
    {
        Vector3 v;
        const Rotation& R = *this;
        if (R(2, 0) < 1.0) {
            if (R(2, 0) > -1.0) {
                v(0) = std::atan2(R(2, 1), R(2, 2)); v(1) = std::asin(-R(2, 0)); v(2) = std::atan2(R(1, 0), R(0, 0));
            } else {
                v(0) = 0.0; v(1) = M_PI / 2.0; v(2) = -std::atan2(-R(1, 2), R(1, 1));
            }
        } else {
            v(0) = 0.0; v(1) = -M_PI / 2.0; v(2) = std::atan2(-R(1, 2), R(1, 1));
        }
        return v;
    }
};

class Transform
{
public:
    const Rotation& getRotation() const { return m_R; }
    const Position& getPosition() const { return m_p; }
    void setRotation(const Rotation& R) { m_R = R; }
    void setPosition(const Position& p) { m_p = p; }
    Matrix6x6 asAdjointTransform() const
    {
        Matrix6x6 X;
        const double S[3][3] = {{0.0, -m_p(2), m_p(1)}, {m_p(2), 0.0, -m_p(0)}, {-m_p(1), m_p(0), 0.0}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                X(i, j) = m_R(i, j);
                X(3 + i, 3 + j) = m_R(i, j);
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += S[i][k] * m_R(k, j);
                X(i, 3 + j) = s;
            }
        return X;
    }
private:
    Rotation m_R;
    Position m_p;
};

class Twist
{
public:
    const LinearMotionVector3& getLinearVec3() const { return m_lin; }
    const AngularMotionVector3& getAngularVec3() const { return m_ang; }
    LinearMotionVector3& getLinearVec3() { return m_lin; }
    AngularMotionVector3& getAngularVec3() { return m_ang; }
private:
    LinearMotionVector3 m_lin;
    AngularMotionVector3 m_ang;
};

class VectorDynSize
{
public:
    VectorDynSize() {}
    explicit VectorDynSize(size_t n) : m_d(n, 0.0) {}
    void resize(size_t n) { m_d.resize(n, 0.0); }
    size_t size() const { return m_d.size(); }
    double& operator()(size_t i) { return m_d[i]; }
    double operator()(size_t i) const { return m_d[i]; }
    std::vector<double> m_d;
};
class MatrixDynSize {};
class Model {};
class ModelLoader {};
class KinDynComputations {};

template <int N> inline Eigen::View toEigen(VectorFixSize<N>& v) { return Eigen::View(v.m_d, N, 1, 1, N); }
template <int N> inline Eigen::View toEigen(const VectorFixSize<N>& v) { return Eigen::View(const_cast<double*>(v.m_d), N, 1, 1, N); }
template <int R, int C> inline Eigen::View toEigen(MatrixFixSize<R, C>& m) { return Eigen::View(m.m_d, R, C, C, 1); }
template <int R, int C> inline Eigen::View toEigen(const MatrixFixSize<R, C>& m) { return Eigen::View(const_cast<double*>(m.m_d), R, C, C, 1); }
inline Eigen::View toEigen(VectorDynSize& v) { return Eigen::View(v.m_d.data(), Eigen::Index(v.m_d.size()), 1, 1, Eigen::Index(v.m_d.size())); }
inline Eigen::View toEigen(const VectorDynSize& v) { return Eigen::View(const_cast<double*>(v.m_d.data()), Eigen::Index(v.m_d.size()), 1, 1, Eigen::Index(v.m_d.size())); }
} // namespace iDynTree
