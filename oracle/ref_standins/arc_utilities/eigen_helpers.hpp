/*
 * Stand-in for arc_utilities/eigen_helpers.hpp - TEST INFRASTRUCTURE ONLY (see arc_helpers.hpp next to this file).
 *
 * UM-ARM-Lab/arc_utilities is neither vendored by the reference nor present here.  This header provides the EigenHelpers
 * names the reference's simple_robot_models.hpp calls, written from their textbook definitions (Murray-Li-Sastry closed forms
 * for the SE(3) exponential / logarithm, shortest-arc angle arithmetic, circular mean) with libm - deliberately independent of
 * include/upc_math.h and of the oracle's own closed forms, so that comparing the oracle with the reference header compiled on
 * top of this layer checks both the reference's composition of these pieces and the oracle's formulas, to rounding.
 * simple_uncertainty_models.hpp includes this file too and uses nothing from it.
 */
#ifndef EIGEN_HELPERS_STANDIN_HPP
#define EIGEN_HELPERS_STANDIN_HPP

#include <Eigen/Geometry>
#include <cmath>
#include <vector>

namespace EigenHelpers
{
typedef std::vector<Eigen::Vector4d, Eigen::aligned_allocator<Eigen::Vector4d>> VectorVector4d;
typedef std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> VectorVector3d;
typedef std::vector<Eigen::Isometry3d, Eigen::aligned_allocator<Eigen::Isometry3d>> VectorIsometry3d;

/* angle in (-pi, pi] */
inline double EnforceContinuousRevoluteBounds(const double value)
{
    if ((value <= -M_PI) || (value > M_PI))
    {
        const double remainder = std::fmod(value, 2.0 * M_PI);
        if (remainder <= -M_PI) return remainder + 2.0 * M_PI;
        if (remainder > M_PI) return remainder - 2.0 * M_PI;
        return remainder;
    }
    return value;
}

/* shortest signed rotation that takes start to end */
inline double ContinuousRevoluteSignedDistance(const double start, const double end)
{
    const double s = EnforceContinuousRevoluteBounds(start);
    const double e = EnforceContinuousRevoluteBounds(end);
    return EnforceContinuousRevoluteBounds(e - s);
}

inline double ContinuousRevoluteDistance(const double start, const double end) { return std::abs(ContinuousRevoluteSignedDistance(start, end)); }

inline double Interpolate(const double p1, const double p2, const double ratio) { return p1 + ((p2 - p1) * ratio); }

inline double InterpolateContinuousRevolute(const double p1, const double p2, const double ratio)
{
    return EnforceContinuousRevoluteBounds(p1 + ContinuousRevoluteSignedDistance(p1, p2) * ratio);
}

inline std::vector<double> Abs(const std::vector<double>& values)
{
    std::vector<double> out(values.size());
    for (size_t i = 0; i < values.size(); i++) out[i] = std::abs(values[i]);
    return out;
}

inline double AverageStdVectorDouble(const std::vector<double>& values)
{
    double sum = 0.0;
    for (size_t i = 0; i < values.size(); i++) sum += values[i];
    return sum / (double)values.size();
}

/* circular mean */
inline double AverageContinuousRevolute(const std::vector<double>& angles)
{
    double s = 0.0, c = 0.0;
    for (size_t i = 0; i < angles.size(); i++)
    {
        s += std::sin(angles[i]);
        c += std::cos(angles[i]);
    }
    return std::atan2(s, c);
}

inline Eigen::VectorXd AverageEigenVectorXd(const std::vector<Eigen::VectorXd>& vectors)
{
    Eigen::VectorXd sum = Eigen::VectorXd::Zero(vectors.front().size());
    for (size_t i = 0; i < vectors.size(); i++) sum = sum + vectors[i];
    return sum / (double)vectors.size();
}

inline Eigen::Matrix3d Skew(const Eigen::Vector3d& v)
{
    Eigen::Matrix3d m = Eigen::Matrix3d::Zero();
    m(0, 1) = -v(2);
    m(0, 2) = v(1);
    m(1, 0) = v(2);
    m(1, 2) = -v(0);
    m(2, 0) = -v(1);
    m(2, 1) = v(0);
    return m;
}

/* exp of the twist [v; w] * delta_t (Murray, Li, Sastry, A Mathematical Introduction to Robotic Manipulation, eq. 2.36) */
inline Eigen::Isometry3d ExpTwist(const Eigen::Matrix<double, 6, 1>& twist, const double delta_t)
{
    const Eigen::Vector3d v = twist.block<3, 1>(0, 0);
    const Eigen::Vector3d w = twist.block<3, 1>(3, 0);
    const double wn = w.norm();
    if (wn < 1e-12)
    {
        return Eigen::Isometry3d::FromParts(Eigen::Matrix3d::Identity(), v * delta_t);
    }
    const Eigen::Vector3d wh = w / wn;
    const Eigen::Vector3d vh = v / wn;
    const double theta = wn * delta_t;
    const Eigen::Matrix3d K = Skew(wh);
    const Eigen::Matrix3d R = Eigen::Matrix3d::Identity() + (K * std::sin(theta)) + ((K * K) * (1.0 - std::cos(theta)));
    const Eigen::Vector3d t = ((Eigen::Matrix3d::Identity() - R) * wh.cross(vh)) + (wh * (wh.dot(vh) * theta));
    return Eigen::Isometry3d::FromParts(R, t);
}

/* log of a rigid transform as the twist [v; w] (same reference, proposition 2.9) */
inline Eigen::Matrix<double, 6, 1> TwistUnhat(const Eigen::Isometry3d& T)
{
    const Eigen::Matrix3d R = T.rotation();
    const Eigen::Vector3d p = T.translation();
    Eigen::Matrix<double, 6, 1> twist = Eigen::Matrix<double, 6, 1>::Zero();
    const Eigen::Vector3d s((R(2, 1) - R(1, 2)) * 0.5, (R(0, 2) - R(2, 0)) * 0.5, (R(1, 0) - R(0, 1)) * 0.5);
    const double sn = s.norm();
    const double cs = ((R(0, 0) + R(1, 1) + R(2, 2)) - 1.0) * 0.5;
    const double theta = std::atan2(sn, cs);
    if (theta < 1e-9)
    {
        twist.block<3, 1>(0, 0) = p;
        return twist;
    }
    Eigen::Vector3d wh;
    if (sn > 1e-6)
    {
        wh = s / sn;
    }
    else
    {
        /* rotation by (nearly) pi: axis from the symmetric part, R + I = 2 wh wh^T */
        int k = 0;
        if (R(1, 1) > R(k, k)) k = 1;
        if (R(2, 2) > R(k, k)) k = 2;
        Eigen::Vector3d col((R(0, k) + (k == 0 ? 1.0 : 0.0)), (R(1, k) + (k == 1 ? 1.0 : 0.0)), (R(2, k) + (k == 2 ? 1.0 : 0.0)));
        wh = col / col.norm();
        if (wh.dot(s) < 0.0) wh = -wh;
    }
    /* p = [(I - R) K + wh wh^T theta] v  =>  solve the 3x3 system */
    const Eigen::Matrix3d K = Skew(wh);
    Eigen::Matrix3d A = (Eigen::Matrix3d::Identity() - R) * K;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A(i, j) += wh(i) * wh(j) * theta;
    const double det = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
                       A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    Eigen::Vector3d v;
    for (int c = 0; c < 3; c++)
    {
        Eigen::Matrix3d B = A;
        for (int r = 0; r < 3; r++) B(r, c) = p(r);
        const double d = B(0, 0) * (B(1, 1) * B(2, 2) - B(1, 2) * B(2, 1)) - B(0, 1) * (B(1, 0) * B(2, 2) - B(1, 2) * B(2, 0)) +
                         B(0, 2) * (B(1, 0) * B(2, 1) - B(1, 1) * B(2, 0));
        v(c) = d / det;
    }
    twist.block<3, 1>(0, 0) = v * theta;
    twist.block<3, 1>(3, 0) = wh * theta;
    return twist;
}

/* twist (in t1's frame) that takes t1 to t2 */
inline Eigen::Matrix<double, 6, 1> TwistBetweenTransforms(const Eigen::Isometry3d& t1, const Eigen::Isometry3d& t2)
{
    return TwistUnhat(t1.inverse() * t2);
}

/* rotation angle between two orientations */
inline double Distance(const Eigen::Quaterniond& q1, const Eigen::Quaterniond& q2)
{
    const double dq = std::abs(q1.dot(q2));
    if (dq < (1.0 - std::numeric_limits<double>::epsilon()))
    {
        return std::acos(2.0 * (dq * dq) - 1.0);
    }
    return 0.0;
}

inline double Distance(const Eigen::Vector3d& v1, const Eigen::Vector3d& v2) { return (v2 - v1).norm(); }

inline double Distance(const Eigen::Isometry3d& t1, const Eigen::Isometry3d& t2, const double alpha)
{
    const double vdist = Distance(t1.translation(), t2.translation()) * (1.0 - alpha);
    const double qdist = Distance(Eigen::Quaterniond(t1.rotation()), Eigen::Quaterniond(t2.rotation())) * alpha;
    return vdist + qdist;
}

/* normalised sum of sign-aligned quaternions + mean translation */
inline Eigen::Isometry3d AverageEigenIsometry3d(const VectorIsometry3d& transforms)
{
    Eigen::Vector3d t = Eigen::Vector3d::Zero();
    const Eigen::Quaterniond q0(transforms.front().rotation());
    double w = 0.0, x = 0.0, y = 0.0, z = 0.0;
    for (size_t i = 0; i < transforms.size(); i++)
    {
        t = t + transforms[i].translation();
        const Eigen::Quaterniond q(transforms[i].rotation());
        const double sign = (q.dot(q0) < 0.0) ? -1.0 : 1.0;
        w += sign * q.w();
        x += sign * q.x();
        y += sign * q.y();
        z += sign * q.z();
    }
    const Eigen::Quaterniond mean = Eigen::Quaterniond(w, x, y, z).normalized();
    return Eigen::Isometry3d::FromParts(mean.toRotationMatrix(), t / (double)transforms.size());
}

/* translation linearly, rotation along the shortest arc (slerp) */
inline Eigen::Isometry3d Interpolate(const Eigen::Isometry3d& t1, const Eigen::Isometry3d& t2, const double ratio)
{
    const Eigen::Vector3d t = t1.translation() + ((t2.translation() - t1.translation()) * ratio);
    const Eigen::Quaterniond q1(t1.rotation());
    Eigen::Quaterniond q2(t2.rotation());
    double d = q1.dot(q2);
    if (d < 0.0)
    {
        q2 = Eigen::Quaterniond(-q2.w(), -q2.x(), -q2.y(), -q2.z());
        d = -d;
    }
    double a = 1.0 - ratio, b = ratio;
    if (d < 1.0 - 1e-9)
    {
        const double th = std::acos(d);
        a = std::sin((1.0 - ratio) * th) / std::sin(th);
        b = std::sin(ratio * th) / std::sin(th);
    }
    const Eigen::Quaterniond q = Eigen::Quaterniond(a * q1.w() + b * q2.w(), a * q1.x() + b * q2.x(), a * q1.y() + b * q2.y(), a * q1.z() + b * q2.z()).normalized();
    return Eigen::Isometry3d::FromParts(q.toRotationMatrix(), t);
}
}

#endif
