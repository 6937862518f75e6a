/*
 * upc_oracle.cpp - CPU oracle for the particle hot path.  TEST INFRASTRUCTURE ONLY (see upc_oracle.hpp).
 *
 * Reference files restated here (paths under include/uncertainty_planning_core/ of the reference):
 *   simple_pid_controller.hpp:98-135            Pid
 *   simple_uncertainty_models.hpp:29,38-45,59,68-88   SensorValue / Actuator
 *   simple_robot_models.hpp:110-119,157-168,213-295,334-371,410-430        Se2Robot
 *   simple_robot_models.hpp:533-537,654-763,813-874                         Se3Robot
 *   simple_robot_models.hpp:881-1151,1299-1369,1784-1911,2011-2084,2176-2217 LinkedRobot
 *   simple_simulator_interface.hpp:613-623      ForwardSimulateRobots / CheckConfigCollision contract
 *   uncertainty_contact_planning.hpp:1558-1578,1583-1686,1698-1799,1801-1863,1915-1984,1986-2034  clustering
 * Pieces with no source in the reference follow DESIGN.md (sections 3-5).
 */
#include "upc_oracle.hpp"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <iterator>
#include <functional>
#include <memory>
#include <stdexcept>
#include <omp.h>

namespace upc_oracle
{
static thread_local std::string g_last_error;

/* ---------------------------------------------------------------- small linear algebra */

Mat34 Identity34()
{
    Mat34 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.0 : 0.0;
    return r;
}

/* c = a o b with the fused-multiply-add chains spelled out */
Mat34 Compose(const Mat34& a, const Mat34& b)
{
    Mat34 c;
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++)
        {
            c.m[i][j] = upc_fma(a.m[i][2], b.m[2][j], upc_fma(a.m[i][1], b.m[1][j], a.m[i][0] * b.m[0][j]));
        }
        c.m[i][3] = upc_fma(a.m[i][2], b.m[2][3], upc_fma(a.m[i][1], b.m[1][3], upc_fma(a.m[i][0], b.m[0][3], a.m[i][3])));
    }
    return c;
}

void TransformPoint(const Mat34& a, const double p[3], double out[3])
{
    for (int i = 0; i < 3; i++)
    {
        out[i] = upc_fma(a.m[i][2], p[2], upc_fma(a.m[i][1], p[1], upc_fma(a.m[i][0], p[0], a.m[i][3])));
    }
}

static Mat34 FromRows12(const double* v)
{
    /* boundary layout of an SE(3) configuration / a 3x4 transform given as R row-major then t */
    Mat34 r;
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++) r.m[i][j] = v[i * 3 + j];
        r.m[i][3] = v[9 + i];
    }
    return r;
}

static void ToRows12(const Mat34& t, double* v)
{
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++) v[i * 3 + j] = t.m[i][j];
        v[9 + i] = t.m[i][3];
    }
}

static Mat34 FromRowMajor34(const double* v)
{
    Mat34 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) r.m[i][j] = v[i * 4 + j];
    return r;
}

/* ---------------------------------------------------------------- PID, sensor, actuator */

void Pid::Initialize(double in_kp, double in_ki, double in_kd, double in_clamp)
{
    kp = fabs(in_kp);
    ki = fabs(in_ki);
    kd = fabs(in_kd);
    integral_clamp = fabs(in_clamp);
    error_integral = 0.0;
    last_error = 0.0;
}

void Pid::Zero()
{
    last_error = 0.0;
    error_integral = 0.0;
}

double Pid::ComputeFeedbackTerm(double current_error, double timestep)
{
    const double timestep_error_integral = ((current_error * 0.5) + (last_error * 0.5)) * timestep;
    const double new_error_integral = error_integral + timestep_error_integral;
    error_integral = std::max(-integral_clamp, std::min(integral_clamp, new_error_integral));
    const double error_derivative = (current_error - last_error) / timestep;
    last_error = current_error;
    const double correction = (current_error * kp) + (error_integral * ki) + (error_derivative * kd);
    return correction;
}

void Actuator::Initialize(double in_noise_bound, double in_limit)
{
    actuator_limit = fabs(in_limit);
    noise_bound = fabs(in_noise_bound);
}

double Actuator::GetControlValue(double control_input) const
{
    double real_control_input = std::min(actuator_limit, control_input);
    real_control_input = std::max(-actuator_limit, real_control_input);
    return real_control_input;
}

double Actuator::GetControlValue(double control_input, double draw) const
{
    const double real_control_input = GetControlValue(control_input);
    const double noise = draw * (noise_bound * real_control_input);
    const double noisy_control_input = real_control_input + noise;
    return noisy_control_input;
}

/* ---------------------------------------------------------------- SE(3) closed forms */

static inline void Cross(const double a[3], const double b[3], double out[3])
{
    out[0] = upc_fma(a[1], b[2], -(a[2] * b[1]));
    out[1] = upc_fma(a[2], b[0], -(a[0] * b[2]));
    out[2] = upc_fma(a[0], b[1], -(a[1] * b[0]));
}

static inline double Dot3(const double a[3], const double b[3])
{
    return upc_fma(a[2], b[2], upc_fma(a[1], b[1], a[0] * b[0]));
}

void Se3Log(const Mat34& t, double twist[6])
{
    const double(*R)[4] = t.m;
    const double trace = (R[0][0] + R[1][1]) + R[2][2];
    double s[3];
    s[0] = 0.5 * (R[2][1] - R[1][2]);
    s[1] = 0.5 * (R[0][2] - R[2][0]);
    s[2] = 0.5 * (R[1][0] - R[0][1]);
    const double sn = sqrt(Dot3(s, s));
    const double ct = 0.5 * (trace - 1.0);
    const double theta = upc_atan2(sn, ct);
    double w[3];
    if (theta < 1.0e-8)
    {
        w[0] = s[0]; w[1] = s[1]; w[2] = s[2];
    }
    else if ((UPC_PI - theta) > 1.0e-3)
    {
        const double k = theta / sn;
        w[0] = s[0] * k; w[1] = s[1] * k; w[2] = s[2] * k;
    }
    else
    {
        /* near pi: axis from the symmetric part, (R + R^T)/2 = ct*I + (1-ct) a a^T */
        const double omc = 1.0 - ct;
        double a2[3];
        for (int i = 0; i < 3; i++) a2[i] = (R[i][i] - ct) / omc;
        int k = 0;
        if (a2[1] > a2[k]) k = 1;
        if (a2[2] > a2[k]) k = 2;
        double a[3];
        const double ak = sqrt(upc_max(a2[k], 0.0));
        a[k] = (s[k] < 0.0) ? -ak : ak;
        const double denom = (2.0 * omc) * a[k];
        for (int j = 0; j < 3; j++)
        {
            if (j != k) a[j] = (R[k][j] + R[j][k]) / denom;
        }
        w[0] = a[0] * theta; w[1] = a[1] * theta; w[2] = a[2] * theta;
    }
    /* v = t - 0.5 w x t + k2 w x (w x t) */
    double k2;
    const double th2 = theta * theta;
    if (theta < 0.05)
    {
        k2 = upc_fma(th2, 1.0 / 47900160.0, 1.0 / 1209600.0);
        k2 = upc_fma(th2, k2, 1.0 / 30240.0);
        k2 = upc_fma(th2, k2, 1.0 / 720.0);
        k2 = upc_fma(th2, k2, 1.0 / 12.0);
    }
    else
    {
        const double two_omc = 2.0 * (1.0 - ct);
        k2 = (two_omc - theta * sn) / (two_omc * th2);
    }
    const double tr[3] = {R[0][3], R[1][3], R[2][3]};
    double wxt[3], wxwxt[3];
    Cross(w, tr, wxt);
    Cross(w, wxt, wxwxt);
    for (int i = 0; i < 3; i++)
    {
        twist[i] = upc_fma(k2, wxwxt[i], upc_fma(-0.5, wxt[i], tr[i]));
        twist[3 + i] = w[i];
    }
}

Mat34 Se3Exp(const double twist[6])
{
    const double* v = twist;
    const double* w = twist + 3;
    const double th2 = Dot3(w, w);
    const double theta = sqrt(th2);
    double A, B, C;
    if (theta < 1.0e-3)
    {
        A = upc_fma(th2, -1.0 / 6.0, 1.0);
        B = upc_fma(th2, -1.0 / 24.0, 0.5);
        C = upc_fma(th2, -1.0 / 120.0, 1.0 / 6.0);
    }
    else
    {
        double sn, cs;
        upc_sincos(theta, &sn, &cs);
        const double inv = 1.0 / theta;
        A = sn * inv;
        B = ((1.0 - cs) * inv) * inv;
        C = ((1.0 - A) * inv) * inv;
    }
    Mat34 out;
    /* R = I + A [w]x + B [w]x^2 ;  [w]x^2 = w w^T - |w|^2 I */
    const double wx = w[0], wy = w[1], wz = w[2];
    out.m[0][0] = upc_fma(B, upc_fma(wx, wx, -th2), 1.0);
    out.m[1][1] = upc_fma(B, upc_fma(wy, wy, -th2), 1.0);
    out.m[2][2] = upc_fma(B, upc_fma(wz, wz, -th2), 1.0);
    out.m[0][1] = upc_fma(B, wx * wy, -(A * wz));
    out.m[1][0] = upc_fma(B, wx * wy, A * wz);
    out.m[0][2] = upc_fma(B, wx * wz, A * wy);
    out.m[2][0] = upc_fma(B, wx * wz, -(A * wy));
    out.m[1][2] = upc_fma(B, wy * wz, -(A * wx));
    out.m[2][1] = upc_fma(B, wy * wz, A * wx);
    /* p = v + B w x v + C w x (w x v) */
    double wxv[3], wxwxv[3];
    Cross(w, v, wxv);
    Cross(w, wxv, wxwxv);
    for (int i = 0; i < 3; i++)
    {
        out.m[i][3] = upc_fma(C, wxwxv[i], upc_fma(B, wxv[i], v[i]));
    }
    return out;
}

Mat34 Se3Between(const Mat34& a, const Mat34& b)
{
    Mat34 r;
    const double d[3] = {b.m[0][3] - a.m[0][3], b.m[1][3] - a.m[1][3], b.m[2][3] - a.m[2][3]};
    for (int i = 0; i < 3; i++)
    {
        for (int j = 0; j < 3; j++)
        {
            r.m[i][j] = upc_fma(a.m[2][i], b.m[2][j], upc_fma(a.m[1][i], b.m[1][j], a.m[0][i] * b.m[0][j]));
        }
        r.m[i][3] = upc_fma(a.m[2][i], d[2], upc_fma(a.m[1][i], d[1], a.m[0][i] * d[0]));
    }
    return r;
}

/* rotation matrix -> quaternion, the branch structure Eigen::Quaterniond(Matrix3d) uses
 * (needed by simple_robot_models.hpp:871-874 through EigenHelpers::Distance) */
void QuaternionFromRotation(const Mat34& tf, double q[4])
{
    const double(*R)[4] = tf.m;
    double t = (R[0][0] + R[1][1]) + R[2][2];
    if (t > 0.0)
    {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[2][1] - R[1][2]) * t;
        q[1] = (R[0][2] - R[2][0]) * t;
        q[2] = (R[1][0] - R[0][1]) * t;
    }
    else
    {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        const int j = (i + 1) % 3;
        const int k = (j + 1) % 3;
        t = sqrt(((R[i][i] - R[j][j]) - R[k][k]) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[k][j] - R[j][k]) * t;
        q[j] = (R[j][i] + R[i][j]) * t;
        q[k] = (R[k][i] + R[i][k]) * t;
    }
}

/* ---------------------------------------------------------------- environment */

Environment::Environment(const upc_env_desc& d)
{
    nx = d.nx; ny = d.ny; nz = d.nz;
    resolution = d.resolution;
    inv_resolution = 1.0 / d.resolution;
    grid_from_world = FromRowMajor34(d.grid_from_world);
    sdf = d.sdf;
    convex_segment = d.convex_segment;
    occupancy = d.occupancy;
    sdf_oob_value = d.sdf_oob_value;
    occupancy_oob_value = d.occupancy_oob_value;
}

/* voxel-space transform of a link: V = inv_res * (grid_from_world o link)   */
static Mat34 VoxelFromLink(const Environment& env, const Mat34& link)
{
    Mat34 v = Compose(env.grid_from_world, link);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) v.m[i][j] = v.m[i][j] * env.inv_resolution;
    return v;
}

static inline bool VoxelInBounds(const Environment& env, const double u[3])
{
    return (u[0] >= 0.0) && (u[0] < (double)env.nx) && (u[1] >= 0.0) && (u[1] < (double)env.ny) && (u[2] >= 0.0) &&
           (u[2] < (double)env.nz);
}

static inline int64_t CellIndex(const Environment& env, int64_t ix, int64_t iy, int64_t iz)
{
    return (ix * (int64_t)env.ny + iy) * (int64_t)env.nz + iz;
}

bool Environment::LocationToGridIndex(const double world[3], int64_t idx[3]) const
{
    const Mat34 v = VoxelFromLink(*this, Identity34());
    double u[3];
    TransformPoint(v, world, u);
    if (!VoxelInBounds(*this, u)) return false;
    idx[0] = (int64_t)floor(u[0]);
    idx[1] = (int64_t)floor(u[1]);
    idx[2] = (int64_t)floor(u[2]);
    return true;
}

/*
 * Trilinear signed distance and its analytic gradient at voxel coordinate u.
 * Returns false (dist = oob value, zero gradient) when u is out of bounds.
 * grad_world is d(dist)/d(world point), metres per metre.
 */
static bool SdfSample(const Environment& env, const double u[3], double* dist, double grad_world[3])
{
    if (!VoxelInBounds(env, u))
    {
        *dist = (double)env.sdf_oob_value;
        if (grad_world) { grad_world[0] = 0.0; grad_world[1] = 0.0; grad_world[2] = 0.0; }
        return false;
    }
    const int64_t n[3] = {env.nx, env.ny, env.nz};
    int64_t lo[3], hi[3];
    double f[3];
    for (int a = 0; a < 3; a++)
    {
        const double c = u[a] - 0.5;
        const double fl = floor(c);
        f[a] = c - fl;
        const int64_t i0 = (int64_t)fl;
        lo[a] = std::max<int64_t>(i0, 0);
        hi[a] = std::min<int64_t>(i0 + 1, n[a] - 1);
    }
    double v[2][2][2];
    for (int ix = 0; ix < 2; ix++)
        for (int iy = 0; iy < 2; iy++)
            for (int iz = 0; iz < 2; iz++)
            {
                v[ix][iy][iz] = (double)env.sdf[CellIndex(env, ix ? hi[0] : lo[0], iy ? hi[1] : lo[1], iz ? hi[2] : lo[2])];
            }
    double cz[2][2], dz[2][2];
    for (int ix = 0; ix < 2; ix++)
        for (int iy = 0; iy < 2; iy++)
        {
            dz[ix][iy] = v[ix][iy][1] - v[ix][iy][0];
            cz[ix][iy] = upc_fma(f[2], dz[ix][iy], v[ix][iy][0]);
        }
    double cy[2], dyv[2], dzv[2];
    for (int ix = 0; ix < 2; ix++)
    {
        dyv[ix] = cz[ix][1] - cz[ix][0];
        cy[ix] = upc_fma(f[1], dyv[ix], cz[ix][0]);
        dzv[ix] = upc_fma(f[1], dz[ix][1] - dz[ix][0], dz[ix][0]);
    }
    const double ddx = cy[1] - cy[0];
    *dist = upc_fma(f[0], ddx, cy[0]);
    if (grad_world)
    {
        const double ddy = upc_fma(f[0], dyv[1] - dyv[0], dyv[0]);
        const double ddz = upc_fma(f[0], dzv[1] - dzv[0], dzv[0]);
        const double gg[3] = {ddx * env.inv_resolution, ddy * env.inv_resolution, ddz * env.inv_resolution};
        const Mat34& G = env.grid_from_world;
        for (int j = 0; j < 3; j++)
        {
            grad_world[j] = upc_fma(G.m[2][j], gg[2], upc_fma(G.m[1][j], gg[1], G.m[0][j] * gg[0]));
        }
    }
    return true;
}

/* ---------------------------------------------------------------- robot models */

struct PointSet
{
    /* per link: list of (x, y, z, radius) */
    std::vector<std::vector<std::array<double, 4>>> links;
};

static std::shared_ptr<PointSet> MakePointSet(const upc_robot_desc& d)
{
    auto ps = std::make_shared<PointSet>();
    ps->links.resize((size_t)d.num_links);
    for (int l = 0; l < d.num_links; l++)
    {
        for (int p = d.link_point_offset[l]; p < d.link_point_offset[l + 1]; p++)
        {
            ps->links[(size_t)l].push_back({d.points[4 * p + 0], d.points[4 * p + 1], d.points[4 * p + 2], d.points[4 * p + 3]});
        }
    }
    return ps;
}

/* SimpleSE2Robot, simple_robot_models.hpp:92-431 */
class Se2Robot
{
public:
    static const int kDof = 3;
    static const int kWidth = 3;
    double config_[3];
    Mat34 pose_;
    Pid controllers_[3];
    Actuator actuators_[3];
    std::shared_ptr<PointSet> points_;
    bool sensor_noise_free_;
    bool SensorNoiseFree() const { return sensor_noise_free_; }

    explicit Se2Robot(const upc_robot_desc& d) : points_(MakePointSet(d))
    {
        sensor_noise_free_ = true;
        for (int i = 0; i < 3; i++)
            if (d.axis[i].max_sensor_noise != 0.0) sensor_noise_free_ = false;
        for (int i = 0; i < 3; i++)
        {
            controllers_[i].Initialize(d.axis[i].kp, d.axis[i].ki, d.axis[i].kd, d.axis[i].integral_clamp);
            actuators_[i].Initialize(d.axis[i].max_actuator_noise, d.axis[i].velocity_limit);
        }
        const double zero[3] = {0.0, 0.0, 0.0};
        ResetPosition(zero);
    }
    int Dof() const { return 3; }
    int Kind() const { return UPC_ROBOT_SE2; }
    int NumLinks() const { return 1; }
    const Mat34& LinkTransform(int) const { return pose_; }
    const PointSet& Points() const { return *points_; }

    void SetConfig(const double* c) /* :110-119, :363-371 */
    {
        config_[0] = c[0];
        config_[1] = c[1];
        config_[2] = upc_wrap_angle(c[2]);
        double s, cs;
        upc_sincos(config_[2], &s, &cs);
        pose_.m[0][0] = cs;  pose_.m[0][1] = -s;  pose_.m[0][2] = 0.0; pose_.m[0][3] = config_[0];
        pose_.m[1][0] = s;   pose_.m[1][1] = cs;  pose_.m[1][2] = 0.0; pose_.m[1][3] = config_[1];
        pose_.m[2][0] = 0.0; pose_.m[2][1] = 0.0; pose_.m[2][2] = 1.0; pose_.m[2][3] = 0.0;
    }
    void UpdatePosition(const double* c) { SetConfig(c); }
    void ResetPosition(const double* c) /* :157-168 */
    {
        SetConfig(c);
        for (int i = 0; i < 3; i++) controllers_[i].Zero();
    }
    void GetPosition(double* c) const { c[0] = config_[0]; c[1] = config_[1]; c[2] = config_[2]; }

    static void RawDistance(const double* c1, const double* c2, double* d) /* :410-417 */
    {
        d[0] = c2[0] - c1[0];
        d[1] = c2[1] - c1[1];
        d[2] = upc_angle_signed_distance(c1[2], c2[2]);
    }
    double ComputeConfigurationDistance(const double* c1, const double* c2) const /* :424-430 */
    {
        double d[3];
        RawDistance(c1, c2, d);
        const double trans_dist = sqrt((d[0] * d[0]) + (d[1] * d[1]));
        const double rots_dist = fabs(d[2]);
        return trans_dist + rots_dist;
    }
    void GenerateControlAction(const double* target, double interval, const double* sensor_draws, double* u) /* :239-259 */
    {
        double current[3];
        for (int i = 0; i < 3; i++) current[i] = SensorValue(config_[i], sensor_draws[i]); /* :213-222, no re-wrap */
        double err[3];
        RawDistance(current, target, err);
        for (int i = 0; i < 3; i++)
        {
            const double term = controllers_[i].ComputeFeedbackTerm(err[i], interval);
            u[i] = actuators_[i].GetControlValue(term);
        }
    }
    void ApplyControlInput(const double* input, const double* draws) /* :282-308 */
    {
        double next[3];
        for (int i = 0; i < 3; i++)
        {
            const double n = draws ? actuators_[i].GetControlValue(input[i], draws[i]) : actuators_[i].GetControlValue(input[i]);
            next[i] = config_[i] + n;
        }
        SetConfig(next);
    }
    /* :334-361 */
    void PointJacobian(int, const double*, const double* world, double J[3][UPC_MAX_DOF]) const
    {
        J[0][0] = 1.0; J[1][0] = 0.0; J[2][0] = 0.0;
        J[0][1] = 0.0; J[1][1] = 1.0; J[2][1] = 0.0;
        const double rx = world[0] - config_[0];
        const double ry = world[1] - config_[1];
        J[0][2] = -ry; J[1][2] = rx; J[2][2] = 0.0;
    }
};

/* SimpleSE3Robot, simple_robot_models.hpp:507-875 */
class Se3Robot
{
public:
    static const int kDof = 6;
    static const int kWidth = 12;
    Mat34 config_;
    Pid controllers_[6];
    Actuator actuators_[6];
    std::shared_ptr<PointSet> points_;
    bool sensor_noise_free_;

    explicit Se3Robot(const upc_robot_desc& d) : points_(MakePointSet(d))
    {
        sensor_noise_free_ = true;
        for (int i = 0; i < 6; i++)
        {
            controllers_[i].Initialize(d.axis[i].kp, d.axis[i].ki, d.axis[i].kd, d.axis[i].integral_clamp);
            actuators_[i].Initialize(d.axis[i].max_actuator_noise, d.axis[i].velocity_limit);
            if (d.axis[i].max_sensor_noise != 0.0) sensor_noise_free_ = false;
        }
        config_ = Identity34();
    }
    int Dof() const { return 6; }
    int Kind() const { return UPC_ROBOT_SE3; }
    bool SensorNoiseFree() const { return sensor_noise_free_; }
    int NumLinks() const { return 1; }
    const Mat34& LinkTransform(int) const { return config_; }
    const PointSet& Points() const { return *points_; }

    void UpdatePosition(const double* c) { config_ = FromRows12(c); }
    void ResetPosition(const double* c)
    {
        config_ = FromRows12(c);
        for (int i = 0; i < 6; i++) controllers_[i].Zero();
    }
    void GetPosition(double* c) const { ToRows12(config_, c); }

    double ComputeConfigurationDistance(const double* c1, const double* c2) const /* :871-874 */
    {
        const Mat34 t1 = FromRows12(c1);
        const Mat34 t2 = FromRows12(c2);
        const double dx = t1.m[0][3] - t2.m[0][3];
        const double dy = t1.m[1][3] - t2.m[1][3];
        const double dz = t1.m[2][3] - t2.m[2][3];
        const double tdist = sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
        double q1[4], q2[4];
        QuaternionFromRotation(t1, q1);
        QuaternionFromRotation(t2, q2);
        const double dq = fabs((((q1[0] * q2[0]) + (q1[1] * q2[1])) + (q1[2] * q2[2])) + (q1[3] * q2[3]));
        double rdist = 0.0;
        if (dq < (1.0 - 2.220446049250313e-16))
        {
            rdist = upc_acos(((2.0 * dq) * dq) - 1.0);
        }
        return (((1.0 - 0.5) * tdist) + (0.5 * rdist)) * 2.0;
    }
    void GenerateControlAction(const double* target, double interval, const double* sensor_draws, double* u) /* :685-714 */
    {
        /* GetPosition(rng), :654-668.  With every sensor noise bound zero the draws are exact zeros and
         * Exp(Log(T) + 0) is T up to rounding: the round trip is skipped. */
        Mat34 current = config_;
        if (!sensor_noise_free_)
        {
            double twist[6];
            Se3Log(config_, twist);
            double noisy[6];
            for (int i = 0; i < 6; i++) noisy[i] = SensorValue(twist[i], sensor_draws[i]);
            current = Se3Exp(noisy);
        }
        const Mat34 tgt = FromRows12(target);
        double err[6];
        Se3Log(Se3Between(current, tgt), err);
        for (int i = 0; i < 6; i++)
        {
            const double term = controllers_[i].ComputeFeedbackTerm(err[i], interval);
            u[i] = actuators_[i].GetControlValue(term);
        }
    }
    void ApplyControlInput(const double* input, const double* draws) /* :746-789 */
    {
        double tw[6];
        for (int i = 0; i < 6; i++)
        {
            tw[i] = draws ? actuators_[i].GetControlValue(input[i], draws[i]) : actuators_[i].GetControlValue(input[i]);
        }
        config_ = Compose(config_, Se3Exp(tw));
    }
    /* :813-832  J = R [ I | -skew(p) ] */
    void PointJacobian(int, const double* local, const double*, double J[3][UPC_MAX_DOF]) const
    {
        const double(*R)[4] = config_.m;
        const double px = local[0], py = local[1], pz = local[2];
        for (int i = 0; i < 3; i++)
        {
            J[i][0] = R[i][0];
            J[i][1] = R[i][1];
            J[i][2] = R[i][2];
            J[i][3] = upc_fma(R[i][2], py, -(R[i][1] * pz));
            J[i][4] = upc_fma(R[i][0], pz, -(R[i][2] * px));
            J[i][5] = upc_fma(R[i][1], px, -(R[i][0] * py));
        }
    }
};

/* SimpleJointModel + SimpleLinkedRobot, simple_robot_models.hpp:877-2218 */
class LinkedRobot
{
public:
    struct Joint
    {
        int type, parent, child, active_index;
        double axis[3], lower, upper, value;
        Mat34 transform;
        Pid controller;
        Actuator actuator;
    };
    int dof_;
    int width_;
    std::vector<Joint> joints_;
    std::vector<Mat34> link_transforms_;
    std::vector<int> parent_joint_of_link_;
    std::vector<double> weights_;
    Mat34 base_;
    std::shared_ptr<PointSet> points_;

    explicit LinkedRobot(const upc_robot_desc& d) : points_(MakePointSet(d))
    {
        dof_ = d.dof;
        sensor_noise_free_ = true;
        for (int i = 0; i < d.dof; i++)
            if (d.axis[i].max_sensor_noise != 0.0) sensor_noise_free_ = false;
        width_ = d.dof;
        base_ = FromRowMajor34(d.base_transform);
        link_transforms_.assign((size_t)d.num_links, Identity34());
        parent_joint_of_link_.assign((size_t)d.num_links, -1);
        int active = 0;
        for (int j = 0; j < d.num_joints; j++)
        {
            const upc_joint_desc& jd = d.joints[j];
            Joint jt;
            jt.type = jd.type;
            jt.parent = jd.parent_link;
            jt.child = jd.child_link;
            for (int a = 0; a < 3; a++) jt.axis[a] = jd.axis[a];
            if (jd.type == UPC_JOINT_CONTINUOUS)
            {
                jt.lower = -INFINITY; jt.upper = INFINITY; /* :906-909 */
            }
            else
            {
                jt.lower = jd.lower_limit; jt.upper = jd.upper_limit;
            }
            jt.transform = FromRowMajor34(jd.transform);
            jt.value = 0.0;
            jt.active_index = -1;
            if (jd.type != UPC_JOINT_FIXED)
            {
                jt.active_index = active;
                const upc_axis_params& ap = d.axis[active];
                jt.controller.Initialize(ap.kp, ap.ki, ap.kd, ap.integral_clamp);
                jt.actuator.Initialize(ap.max_actuator_noise, ap.velocity_limit);
                weights_.push_back(fabs(d.distance_weights[active])); /* :1609 */
                active++;
            }
            parent_joint_of_link_[(size_t)jd.child_link] = j;
            joints_.push_back(jt);
        }
        UpdateTransforms();
    }
    int Dof() const { return dof_; }
    int Kind() const { return UPC_ROBOT_LINKED; }
    bool sensor_noise_free_;
    bool SensorNoiseFree() const { return sensor_noise_free_; }
    int NumLinks() const { return (int)link_transforms_.size(); }
    const Mat34& LinkTransform(int l) const { return link_transforms_[(size_t)l]; }
    const PointSet& Points() const { return *points_; }

    static double EnforceLimits(const Joint& j, double value) /* :1077-1100 */
    {
        if (j.type == UPC_JOINT_CONTINUOUS) return upc_wrap_angle(value);
        if (value < j.lower) return j.lower;
        if (value > j.upper) return j.upper;
        return value;
    }
    static double SignedDistance(const Joint& j, double v1, double v2) /* :1108-1118 */
    {
        if (j.type == UPC_JOINT_CONTINUOUS) return upc_angle_signed_distance(v1, v2);
        return (v2 - v1);
    }
    void UpdateTransforms() /* :1299-1343 */
    {
        link_transforms_[0] = base_;
        for (size_t idx = 0; idx < joints_.size(); idx++)
        {
            const Joint& j = joints_[idx];
            const Mat34 complete = Compose(link_transforms_[(size_t)j.parent], j.transform);
            Mat34 child = complete;
            if ((j.type == UPC_JOINT_REVOLUTE) || (j.type == UPC_JOINT_CONTINUOUS))
            {
                double s, c;
                upc_sincos(j.value, &s, &c);
                const double t = 1.0 - c;
                const double ax = j.axis[0], ay = j.axis[1], az = j.axis[2];
                double Rj[3][3];
                Rj[0][0] = upc_fma(t * ax, ax, c);
                Rj[0][1] = upc_fma(t * ax, ay, -(s * az));
                Rj[0][2] = upc_fma(t * ax, az, s * ay);
                Rj[1][0] = upc_fma(t * ay, ax, s * az);
                Rj[1][1] = upc_fma(t * ay, ay, c);
                Rj[1][2] = upc_fma(t * ay, az, -(s * ax));
                Rj[2][0] = upc_fma(t * az, ax, -(s * ay));
                Rj[2][1] = upc_fma(t * az, ay, s * ax);
                Rj[2][2] = upc_fma(t * az, az, c);
                for (int r = 0; r < 3; r++)
                    for (int cidx = 0; cidx < 3; cidx++)
                    {
                        child.m[r][cidx] = upc_fma(complete.m[r][2], Rj[2][cidx],
                                                   upc_fma(complete.m[r][1], Rj[1][cidx], complete.m[r][0] * Rj[0][cidx]));
                    }
            }
            else if (j.type == UPC_JOINT_PRISMATIC)
            {
                const double d[3] = {j.axis[0] * j.value, j.axis[1] * j.value, j.axis[2] * j.value};
                for (int r = 0; r < 3; r++)
                {
                    child.m[r][3] = upc_fma(complete.m[r][2], d[2],
                                            upc_fma(complete.m[r][1], d[1], upc_fma(complete.m[r][0], d[0], complete.m[r][3])));
                }
            }
            link_transforms_[(size_t)j.child] = child;
        }
    }
    void SetConfig(const double* c) /* :1345-1369 */
    {
        for (auto& j : joints_)
        {
            if (j.type == UPC_JOINT_FIXED) continue;
            j.value = EnforceLimits(j, c[j.active_index]);
        }
        UpdateTransforms();
    }
    void UpdatePosition(const double* c) { SetConfig(c); }
    void ResetPosition(const double* c)
    {
        SetConfig(c);
        for (auto& j : joints_) j.controller.Zero();
    }
    void GetPosition(double* c) const
    {
        for (const auto& j : joints_)
        {
            if (j.type != UPC_JOINT_FIXED) c[j.active_index] = j.value;
        }
    }
    double ComputeConfigurationDistance(const double* c1, const double* c2) const /* :2191-2217 */
    {
        double sum = 0.0;
        for (const auto& j : joints_)
        {
            if (j.type == UPC_JOINT_FIXED) continue;
            const int a = j.active_index;
            const double raw = SignedDistance(j, c1[a], c2[a]);
            const double wd = raw * weights_[(size_t)a];
            sum = sum + (wd * wd);
        }
        return sqrt(sum);
    }
    void GenerateControlAction(const double* target, double interval, const double* sensor_draws, double* u) /* :1823-1851 */
    {
        for (auto& j : joints_)
        {
            if (j.type == UPC_JOINT_FIXED) continue;
            const int a = j.active_index;
            /* GetPosition(rng) :1784-1806 - sensed value goes through CopyWithNewValue => EnforceLimits */
            const double sensed = EnforceLimits(j, SensorValue(j.value, sensor_draws[a]));
            const double err = SignedDistance(j, sensed, target[a]); /* :2176-2189 */
            const double term = j.controller.ComputeFeedbackTerm(err, interval);
            u[a] = j.actuator.GetControlValue(term);
        }
    }
    void ApplyControlInput(const double* input, const double* draws) /* :1882-1940 */
    {
        double next[UPC_MAX_DOF];
        for (auto& j : joints_)
        {
            if (j.type == UPC_JOINT_FIXED) continue;
            const int a = j.active_index;
            const double n = draws ? j.actuator.GetControlValue(input[a], draws[a]) : j.actuator.GetControlValue(input[a]);
            next[a] = EnforceLimits(j, j.value + n);
        }
        SetConfig(next);
    }
    /* :2011-2084, translational rows */
    void PointJacobian(int link, const double*, const double* world, double J[3][UPC_MAX_DOF]) const
    {
        for (int r = 0; r < 3; r++)
            for (int a = 0; a < dof_; a++) J[r][a] = 0.0;
        int current = link;
        while (current > 0)
        {
            const int pj = parent_joint_of_link_[(size_t)current];
            const Joint& j = joints_[(size_t)pj];
            const Mat34& T = link_transforms_[(size_t)j.child];
            if ((j.type == UPC_JOINT_REVOLUTE) || (j.type == UPC_JOINT_CONTINUOUS))
            {
                double aw[3];
                for (int r = 0; r < 3; r++)
                {
                    aw[r] = upc_fma(T.m[r][2], j.axis[2], upc_fma(T.m[r][1], j.axis[1], T.m[r][0] * j.axis[0]));
                }
                const double rel[3] = {world[0] - T.m[0][3], world[1] - T.m[1][3], world[2] - T.m[2][3]};
                double c[3];
                Cross(aw, rel, c);
                for (int r = 0; r < 3; r++) J[r][j.active_index] = c[r];
            }
            else if (j.type == UPC_JOINT_PRISMATIC)
            {
                /* :2064 applies the full isometry (translation included) to the axis; reproduced */
                double aw[3];
                TransformPoint(T, j.axis, aw);
                for (int r = 0; r < 3; r++) J[r][j.active_index] = aw[r];
            }
            current = j.parent;
        }
    }
};

/* ---------------------------------------------------------------- the simulator */

template <typename Robot>
static bool CheckCollision(const Environment& env, const Robot& robot, double threshold)
{
    const PointSet& ps = robot.Points();
    for (int l = 0; l < robot.NumLinks(); l++)
    {
        const Mat34 V = VoxelFromLink(env, robot.LinkTransform(l));
        for (const auto& pt : ps.links[(size_t)l])
        {
            double u[3], d;
            TransformPoint(V, pt.data(), u);
            SdfSample(env, u, &d, nullptr);
            if ((d - pt[3]) < threshold) return true;
        }
    }
    return false;
}

template <typename Robot>
static double EstimateMaxWorkspaceMotion(const Robot& robot, const double* input)
{
    Robot moved = robot;
    moved.ApplyControlInput(input, nullptr);
    const PointSet& ps = robot.Points();
    double max_sq = 0.0;
    for (int l = 0; l < robot.NumLinks(); l++)
    {
        const Mat34& a = robot.LinkTransform(l);
        const Mat34& b = moved.LinkTransform(l);
        Mat34 D;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++) D.m[i][j] = b.m[i][j] - a.m[i][j];
        for (const auto& pt : ps.links[(size_t)l])
        {
            double disp[3];
            TransformPoint(D, pt.data(), disp);
            const double sq = Dot3(disp, disp);
            max_sq = upc_max(max_sq, sq);
        }
    }
    return sqrt(max_sq);
}

/* damped normal equations (H + lambda I) x = g by LDL^T with reciprocal pivots, DESIGN.md 3.5 */
static void SolveDamped(int n, double H[UPC_MAX_DOF][UPC_MAX_DOF], const double* g, double lambda, double* x)
{
    double L[UPC_MAX_DOF][UPC_MAX_DOF];
    double D[UPC_MAX_DOF], R[UPC_MAX_DOF];
    for (int j = 0; j < n; j++)
    {
        double d = H[j][j] + lambda;
        for (int k = 0; k < j; k++) d = upc_fma(-(L[j][k] * L[j][k]), D[k], d);
        if (d < lambda) d = lambda;
        const double r = 1.0 / d;
        for (int i = j + 1; i < n; i++)
        {
            double t = H[i][j];
            for (int k = 0; k < j; k++) t = upc_fma(-(L[i][k] * L[j][k]), D[k], t);
            L[i][j] = t * r;
        }
        D[j] = d;
        R[j] = r;
    }
    double y[UPC_MAX_DOF];
    for (int i = 0; i < n; i++)
    {
        double s = g[i];
        for (int k = 0; k < i; k++) s = upc_fma(-L[i][k], y[k], s);
        y[i] = s;
    }
    for (int i = n - 1; i >= 0; i--)
    {
        double s = y[i] * R[i];
        for (int k = i + 1; k < n; k++) s = upc_fma(-L[k][i], x[k], s);
        x[i] = s;
    }
}

/* one pass over the body points: accumulate corrections into x (raw correction step). Returns number of contributing points. */
template <typename Robot>
static int ComputeResolverCorrection(const Environment& env, const Robot& robot, const upc_sim_params& sp, double* x)
{
    const int n = robot.Dof();
    double H[UPC_MAX_DOF][UPC_MAX_DOF];
    double g[UPC_MAX_DOF];
    double xs[UPC_MAX_DOF];
    for (int a = 0; a < n; a++)
    {
        g[a] = 0.0; xs[a] = 0.0;
        for (int b = 0; b < n; b++) H[a][b] = 0.0;
    }
    int count = 0;
    const PointSet& ps = robot.Points();
    for (int l = 0; l < robot.NumLinks(); l++)
    {
        const Mat34& T = robot.LinkTransform(l);
        const Mat34 V = VoxelFromLink(env, T);
        for (const auto& pt : ps.links[(size_t)l])
        {
            double u[3], d, grad[3];
            TransformPoint(V, pt.data(), u);
            SdfSample(env, u, &d, grad);
            const double de = d - pt[3];
            if (!(de < sp.resolve_distance)) continue;
            const double gn2 = Dot3(grad, grad);
            if (!(gn2 > 1.0e-24)) continue;
            const double scale = (sp.resolve_distance - de) / sqrt(gn2);
            const double c[3] = {grad[0] * scale, grad[1] * scale, grad[2] * scale};
            double world[3];
            TransformPoint(T, pt.data(), world);
            double J[3][UPC_MAX_DOF];
            robot.PointJacobian(l, pt.data(), world, J);
            count++;
            if (sp.use_individual_jacobians)
            {
                /* per-point damped pseudo-inverse: x_i = J^T (J J^T + lambda I)^-1 c, summed then averaged */
                double A[UPC_MAX_DOF][UPC_MAX_DOF];
                for (int r = 0; r < 3; r++)
                    for (int q = 0; q < 3; q++)
                    {
                        double s = 0.0;
                        for (int a = 0; a < n; a++) s = upc_fma(J[r][a], J[q][a], s);
                        A[r][q] = s;
                    }
                double y[UPC_MAX_DOF];
                SolveDamped(3, A, c, sp.resolver_damping, y);
                for (int a = 0; a < n; a++)
                {
                    xs[a] = upc_fma(J[2][a], y[2], upc_fma(J[1][a], y[1], upc_fma(J[0][a], y[0], xs[a])));
                }
            }
            else
            {
                for (int a = 0; a < n; a++)
                {
                    for (int b = a; b < n; b++)
                    {
                        H[a][b] = upc_fma(J[2][a], J[2][b], upc_fma(J[1][a], J[1][b], upc_fma(J[0][a], J[0][b], H[a][b])));
                    }
                    g[a] = upc_fma(J[2][a], c[2], upc_fma(J[1][a], c[1], upc_fma(J[0][a], c[0], g[a])));
                }
            }
        }
    }
    if (count == 0) return 0;
    if (sp.use_individual_jacobians)
    {
        const double inv = 1.0 / (double)count;
        for (int a = 0; a < n; a++) x[a] = xs[a] * inv;
    }
    else
    {
        for (int a = 0; a < n; a++)
            for (int b = 0; b < a; b++) H[a][b] = H[b][a];
        SolveDamped(n, H, g, sp.resolver_damping, x);
    }
    return count;
}

/* per-interval record of one particle, the content of ForwardSimulationStepTrace (simple_simulator_interface.hpp:40-63): the
 * resolved configuration at the end of every controller interval, its control input and per-microstep control step */
struct ParticleTrace
{
    std::vector<std::vector<double>> configs, control_input, control_input_step;
    int32_t last_microsteps = 0;
};

template <typename Robot>
static bool SimulateParticle(const Environment& env, Robot& robot, const upc_sim_params& sp, const double* target,
                             const double* tape, int64_t tape_stride, SimStats& stats, ParticleTrace* trace = nullptr)
{
    /* tape element (step, slot, dof d) of this particle is tape[((step*(1+M)+slot)*dof + d) * tape_stride] */
    const int dof = robot.Dof();
    const int slots = 1 + sp.max_microsteps;
    bool collided = false;
    double cfg[UPC_MAX_DOF > 12 ? UPC_MAX_DOF : 12];
    for (int step = 0; step < sp.num_steps; step++)
    {
        stats.particle_steps++;
        double draws[UPC_MAX_DOF];
        /* with every sensor noise bound zero the sensor slot of the tape holds exact zeros by construction and is not read */
        for (int d = 0; d < dof; d++) draws[d] = robot.SensorNoiseFree() ? 0.0 : tape[(((int64_t)step * slots + 0) * dof + d) * tape_stride];
        double u[UPC_MAX_DOF];
        robot.GenerateControlAction(target, sp.controller_interval, draws, u);
        double real[UPC_MAX_DOF];
        for (int d = 0; d < dof; d++) real[d] = u[d] * sp.controller_interval;
        const double motion = EstimateMaxWorkspaceMotion(robot, real);
        uint32_t microsteps = std::max(1u, (uint32_t)ceil(motion / sp.microstep_distance));
        microsteps = std::min(microsteps, (uint32_t)sp.max_microsteps);
        const double inv_micro = 1.0 / (double)microsteps;
        double ustep[UPC_MAX_DOF];
        for (int d = 0; d < dof; d++) ustep[d] = real[d] * inv_micro;
        for (uint32_t m = 0; m < microsteps; m++)
        {
            double previous[UPC_MAX_DOF > 12 ? UPC_MAX_DOF : 12];
            robot.GetPosition(previous);
            for (int d = 0; d < dof; d++) draws[d] = tape[(((int64_t)step * slots + 1 + m) * dof + d) * tape_stride];
            robot.ApplyControlInput(ustep, draws);
            bool in_collision = CheckCollision(env, robot, sp.contact_distance);
            if (in_collision) collided = true;
            if (in_collision && sp.allow_contacts)
            {
                int iterations = 0;
                double scaling = sp.resolver_initial_scale;
                while (in_collision)
                {
                    double raw[UPC_MAX_DOF];
                    const int contributing = ComputeResolverCorrection(env, robot, sp, raw);
                    if (contributing == 0) break;
                    const double est = EstimateMaxWorkspaceMotion(robot, raw);
                    const double limit_scale = (est > sp.microstep_distance) ? (sp.microstep_distance / est) : 1.0;
                    const double total_scale = limit_scale * fabs(scaling);
                    double corr[UPC_MAX_DOF];
                    for (int d = 0; d < dof; d++) corr[d] = raw[d] * total_scale;
                    robot.ApplyControlInput(corr, nullptr);
                    in_collision = CheckCollision(env, robot, sp.contact_distance);
                    iterations++;
                    stats.resolver_iterations++;
                    if (iterations >= sp.max_resolver_iterations) break;
                    if ((iterations % sp.resolver_decay_iterations) == 0) scaling = scaling * sp.resolver_decay_rate;
                }
                if (in_collision)
                {
                    robot.UpdatePosition(previous);
                    stats.failed_resolves++;
                    break;
                }
            }
            else if (in_collision)
            {
                robot.UpdatePosition(previous);
                break;
            }
        }
        robot.GetPosition(cfg);
        if (trace)
        {
            trace->configs.push_back(std::vector<double>(cfg, cfg + upc_config_width(robot.Kind(), dof)));
            trace->control_input.push_back(std::vector<double>(u, u + dof));
            trace->control_input_step.push_back(std::vector<double>(ustep, ustep + dof));
            trace->last_microsteps = (int32_t)microsteps;
        }
        if (robot.ComputeConfigurationDistance(cfg, target) < sp.simulation_shortcut_distance) break;
    }
    return collided;
}

template <typename Robot>
static void ForwardSimulateRobots(const Environment& env, const upc_robot_desc& rd, const upc_sim_params& sp, int64_t n,
                                  const double* starts, int64_t n_targets, const double* targets, const double* tape,
                                  double* out_configs, uint8_t* out_collided, upc_stats* stats)
{
    const Robot prototype(rd);
    const int width = upc_config_width(rd.kind, rd.dof);
    int64_t steps = 0, coll = 0, iters = 0, fails = 0;
#pragma omp parallel for schedule(static) reduction(+ : steps, coll, iters, fails)
    for (int64_t i = 0; i < n; i++)
    {
        Robot robot = prototype; /* every user of a robot copies it first, e.g. uncertainty_contact_planning.hpp:1599 */
        double start[16], target[16], result[16] = {0.0};
        for (int c = 0; c < width; c++)
        {
            start[c] = starts[(int64_t)c * n + i];
            target[c] = targets[(int64_t)c * n_targets + ((n_targets == 1) ? 0 : i)];
        }
        robot.ResetPosition(start);
        SimStats st = {0, 0, 0, 0};
        const bool collided = SimulateParticle(env, robot, sp, target, tape + i, n, st);
        robot.GetPosition(result);
        for (int c = 0; c < width; c++) out_configs[(int64_t)c * n + i] = result[c];
        out_collided[i] = collided ? 1 : 0;
        steps += st.particle_steps;
        coll += collided ? 1 : 0;
        iters += st.resolver_iterations;
        fails += st.failed_resolves;
    }
    if (stats)
    {
        stats->particles_simulated += n;
        stats->particle_steps += steps;
        stats->collided_particles += coll;
        stats->resolver_iterations += iters;
        stats->failed_resolves += fails;
    }
}

/* ---------------------------------------------------------------- clustering */

std::vector<std::vector<int64_t>> CompleteLinkCluster(const std::vector<double>& D, int64_t n, double threshold)
{
    /* clusters are kept sorted by representative (smallest member); members ascending */
    std::vector<std::vector<int64_t>> clusters((size_t)n);
    for (int64_t i = 0; i < n; i++) clusters[(size_t)i] = {i};
    while (clusters.size() > 1)
    {
        double best = INFINITY;
        size_t ba = 0, bb = 0;
        bool found = false;
        for (size_t a = 0; a < clusters.size(); a++)
        {
            for (size_t b = a + 1; b < clusters.size(); b++)
            {
                double link = 0.0;
                for (int64_t ia : clusters[a])
                    for (int64_t ib : clusters[b]) link = std::max(link, D[(size_t)(ia * n + ib)]);
                if (!found || (link < best))
                {
                    best = link; ba = a; bb = b; found = true;
                }
            }
        }
        if (!(best <= threshold)) break;
        std::vector<int64_t> merged;
        std::merge(clusters[ba].begin(), clusters[ba].end(), clusters[bb].begin(), clusters[bb].end(), std::back_inserter(merged));
        clusters[ba] = merged;
        clusters.erase(clusters.begin() + (ptrdiff_t)bb);
    }
    return clusters;
}

/* arc_helpers::BuildDistanceMatrix as used at uncertainty_contact_planning.hpp:1674,1960: symmetric, zero diagonal */
static std::vector<double> BuildDistanceMatrix(int64_t n, const std::function<double(int64_t, int64_t)>& fn)
{
    std::vector<double> D((size_t)(n * n), 0.0);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++)
    {
        for (int64_t j = i; j < n; j++)
        {
            if (i != j)
            {
                const double d = fn(i, j);
                D[(size_t)(i * n + j)] = d;
                D[(size_t)(j * n + i)] = d;
            }
        }
    }
    return D;
}

static double MaxCoeff(const std::vector<double>& D)
{
    double m = D.empty() ? 0.0 : D[0];
    for (double v : D) m = std::max(m, v);
    return m;
}

/* threshold test + hierarchical clustering, the shape shared by :1677-1685, :1789-1798, :1853-1862, :1962-1980 */
static std::vector<std::vector<int64_t>> ThresholdOrCluster(const std::vector<int64_t>& items, const std::vector<double>& D,
                                                            double threshold, bool* used_hierarchical)
{
    const int64_t n = (int64_t)items.size();
    if (used_hierarchical) *used_hierarchical = false;
    if (MaxCoeff(D) <= threshold)
    {
        return {items};
    }
    if (used_hierarchical) *used_hierarchical = true;
    std::vector<std::vector<int64_t>> out;
    for (const auto& c : CompleteLinkCluster(D, n, threshold))
    {
        std::vector<int64_t> mapped;
        for (int64_t k : c) mapped.push_back(items[(size_t)k]);
        out.push_back(mapped);
    }
    return out;
}

template <typename Robot>
static void RegionSignature(const Environment& env, Robot& robot, const double* cfg, uint32_t* sig) /* :1597-1635 */
{
    robot.UpdatePosition(cfg);
    const PointSet& ps = robot.Points();
    size_t k = 0;
    for (int l = 0; l < robot.NumLinks(); l++)
    {
        const Mat34 V = VoxelFromLink(env, robot.LinkTransform(l));
        for (const auto& pt : ps.links[(size_t)l])
        {
            double u[3];
            TransformPoint(V, pt.data(), u);
            uint32_t s = 0u;
            if (VoxelInBounds(env, u))
            {
                s = env.convex_segment[CellIndex(env, (int64_t)floor(u[0]), (int64_t)floor(u[1]), (int64_t)floor(u[2]))];
            }
            sig[k++] = s;
        }
    }
}

static double SignatureDistance(const uint32_t* a, const uint32_t* b, int64_t total_points) /* :1637-1661 */
{
    size_t non_matching = 0u;
    for (int64_t k = 0; k < total_points; k++)
    {
        if ((a[k] != 0u) && (b[k] != 0u) && ((a[k] & b[k]) == 0u)) non_matching++;
    }
    return (double)non_matching / (double)total_points;
}

/* CheckStraightLine3DPath :1698-1716 */
static bool StraightLineFree(const Environment& env, const double a[3], const double b[3])
{
    const double d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const double distance = sqrt(((d[0] * d[0]) + (d[1] * d[1])) + (d[2] * d[2]));
    const uint32_t steps = std::max(1u, (uint32_t)ceil(distance / env.resolution));
    const Mat34 V = VoxelFromLink(env, Identity34());
    for (uint32_t step = 1u; step <= steps; step++)
    {
        const double percent = (double)step / (double)steps;
        const double q[3] = {upc_fma(d[0], percent, a[0]), upc_fma(d[1], percent, a[1]), upc_fma(d[2], percent, a[2])};
        double u[3];
        TransformPoint(V, q, u);
        float occupancy = env.occupancy_oob_value; /* in-bounds flag ignored at :1709: the default cell is used */
        if (VoxelInBounds(env, u))
        {
            occupancy = env.occupancy[CellIndex(env, (int64_t)floor(u[0]), (int64_t)floor(u[1]), (int64_t)floor(u[2]))];
        }
        if (occupancy > 0.5f) return false;
    }
    return true;
}

template <typename Robot>
static void ClusterOneSet(const Environment& env, const upc_robot_desc& rd, const upc_cluster_params& cp, int64_t n,
                          int64_t total, const double* configs /* [C][total], offset applied */, int32_t* labels,
                          int32_t* n_clusters, upc_stats* stats)
{
    const int width = upc_config_width(rd.kind, rd.dof);
    const Robot prototype(rd);
    std::vector<std::vector<double>> particles((size_t)n, std::vector<double>((size_t)width));
    for (int64_t i = 0; i < n; i++)
        for (int c = 0; c < width; c++) particles[(size_t)i][(size_t)c] = configs[(int64_t)c * total + i];
    if (n == 0)
    {
        *n_clusters = 0;
        return;
    }
    std::vector<int64_t> all((size_t)n);
    for (int64_t i = 0; i < n; i++) all[(size_t)i] = i;
    std::vector<std::vector<int64_t>> initial_clusters;
    if (n == 1)
    {
        initial_clusters = {all}; /* :1993-1996 */
        labels[0] = 0;
        *n_clusters = 1;
        return;
    }
    if (stats) stats->cluster_calls++;
    /* first pass: spatial features (:1915-1945) */
    if (cp.feature_type == UPC_FEATURE_NONE)
    {
        initial_clusters = {all};
    }
    else if (cp.feature_type == UPC_FEATURE_CONVEX_REGION_SIGNATURE)
    {
        if (!env.convex_segment) throw std::invalid_argument("convex_segment grid required for CRS clustering");
        const int64_t P = rd.num_points;
        std::vector<uint32_t> sigs((size_t)(n * P));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++)
        {
            Robot robot = prototype;
            RegionSignature(env, robot, particles[(size_t)i].data(), &sigs[(size_t)(i * P)]);
        }
        const std::vector<double> D = BuildDistanceMatrix(n, [&](int64_t a, int64_t b) {
            return SignatureDistance(&sigs[(size_t)(a * P)], &sigs[(size_t)(b * P)], P);
        });
        initial_clusters = ThresholdOrCluster(all, D, cp.signature_matching_threshold, nullptr);
    }
    else if (cp.feature_type == UPC_FEATURE_ACTUATION_CENTER_CONNECTIVITY)
    {
        if (!env.occupancy) throw std::invalid_argument("occupancy grid required for AC clustering");
        const int L = rd.num_links;
        std::vector<double> centers((size_t)(n * L * 3));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++)
        {
            Robot robot = prototype;
            robot.UpdatePosition(particles[(size_t)i].data());
            for (int l = 0; l < L; l++)
                for (int a = 0; a < 3; a++) centers[(size_t)((i * L + l) * 3 + a)] = robot.LinkTransform(l).m[a][3];
        }
        const std::vector<double> D = BuildDistanceMatrix(n, [&](int64_t a, int64_t b) {
            for (int l = 0; l < L; l++)
            {
                if (!StraightLineFree(env, &centers[(size_t)((a * L + l) * 3)], &centers[(size_t)((b * L + l) * 3)])) return 1.0;
            }
            return 0.0;
        });
        initial_clusters = ThresholdOrCluster(all, D, 0.0, nullptr);
    }
    else if (cp.feature_type == UPC_FEATURE_POINT_TO_POINT_MOVEMENT)
    {
        /* :1801-1863: N^2-N simulations, (i->j) then (j->i) for i<j, allow_contacts = true */
        if (!cp.ptpm_sim) throw std::invalid_argument("PTPM clustering needs sim params (and a tape or a seed)");
        const int64_t sims = n * n - n;
        /* no tape: pair simulation k is particle k of the counter-based stream of cp.ptpm_seed */
        std::vector<double> seeded_tape;
        const double* pair_tape = cp.ptpm_tape;
        if (!pair_tape && sims > 0)
        {
            seeded_tape.resize((size_t)(upc_tape_doubles_per_particle(cp.ptpm_sim, rd.dof) * sims));
            oracle_fill_noise_tape_counter(cp.ptpm_seed, &rd, cp.ptpm_sim, 0, sims, sims, seeded_tape.data());
            pair_tape = seeded_tape.data();
        }
        std::vector<double> starts((size_t)(width * sims)), targets((size_t)(width * sims)), results((size_t)(width * sims));
        std::vector<uint8_t> coll((size_t)sims);
        int64_t cur = 0;
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = i + 1; j < n; j++)
            {
                for (int c = 0; c < width; c++)
                {
                    starts[(size_t)(c * sims + cur)] = particles[(size_t)i][(size_t)c];
                    targets[(size_t)(c * sims + cur)] = particles[(size_t)j][(size_t)c];
                    starts[(size_t)(c * sims + cur + 1)] = particles[(size_t)j][(size_t)c];
                    targets[(size_t)(c * sims + cur + 1)] = particles[(size_t)i][(size_t)c];
                }
                cur += 2;
            }
        upc_sim_params sp = *cp.ptpm_sim;
        sp.allow_contacts = 1;
        ForwardSimulateRobots<Robot>(env, rd, sp, sims, starts.data(), sims, targets.data(), pair_tape, results.data(),
                                     coll.data(), nullptr);
        std::vector<double> D((size_t)(n * n), 0.0);
        cur = 0;
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = i + 1; j < n; j++)
            {
                double fr[16], br[16];
                for (int c = 0; c < width; c++)
                {
                    fr[c] = results[(size_t)(c * sims + cur)];
                    br[c] = results[(size_t)(c * sims + cur + 1)];
                }
                const bool ff = prototype.ComputeConfigurationDistance(fr, particles[(size_t)j].data()) <= cp.goal_distance_threshold;
                const bool bf = prototype.ComputeConfigurationDistance(br, particles[(size_t)i].data()) <= cp.goal_distance_threshold;
                const double d = (ff && bf) ? 0.0 : 1.0;
                D[(size_t)(i * n + j)] = d;
                D[(size_t)(j * n + i)] = d;
                cur += 2;
            }
        initial_clusters = ThresholdOrCluster(all, D, 0.0, nullptr);
    }
    else
    {
        throw std::invalid_argument("unknown feature_type");
    }
    /* second pass: distance threshold per initial cluster (:1947-1984) */
    std::vector<std::vector<int64_t>> final_clusters;
    for (const auto& c : initial_clusters)
    {
        const int64_t m = (int64_t)c.size();
        const std::vector<double> D = BuildDistanceMatrix(m, [&](int64_t a, int64_t b) {
            return prototype.ComputeConfigurationDistance(particles[(size_t)c[(size_t)a]].data(), particles[(size_t)c[(size_t)b]].data());
        });
        bool hierarchical = false;
        for (const auto& sub : ThresholdOrCluster(c, D, cp.distance_clustering_threshold, &hierarchical)) final_clusters.push_back(sub);
        if (hierarchical && stats) stats->cluster_fallback_calls++;
    }
    /* canonical numbering: clusters by increasing smallest member */
    std::sort(final_clusters.begin(), final_clusters.end(),
              [](const std::vector<int64_t>& a, const std::vector<int64_t>& b) { return *std::min_element(a.begin(), a.end()) < *std::min_element(b.begin(), b.end()); });
    for (size_t k = 0; k < final_clusters.size(); k++)
        for (int64_t i : final_clusters[k]) labels[i] = (int32_t)k;
    *n_clusters = (int32_t)final_clusters.size();
}

template <typename Robot>
static void ClusterSets(const Environment& env, const upc_robot_desc& rd, const upc_cluster_params& cp, int64_t n_sets,
                        int64_t set_size, const double* configs, int32_t* labels, int32_t* n_clusters, upc_stats* stats)
{
    const int64_t total = n_sets * set_size;
    /* Particle sets are independent (one per planner edge).  The reference clusters one edge at a time and parallelises inside
     * (the distance matrix, uncertainty_contact_planning.hpp:1674 with ENABLE_PARALLEL_DISTANCE_MATRIX); with many sets in hand
     * the CPU path uses its threads better by taking whole sets per thread, so that is what a batch of sets does here - the
     * per-set work (and its result) is the same either way. */
    const bool over_sets = (n_sets >= 2 * (int64_t)omp_get_max_threads()) && !omp_in_parallel();
    int64_t calls = 0, fallbacks = 0;
    std::string failure; /* exceptions must not leave the parallel region */
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : calls, fallbacks) if (over_sets)
    for (int64_t s = 0; s < n_sets; s++)
    {
        upc_cluster_params local = cp;
        upc_stats st;
        memset(&st, 0, sizeof(st));
        try
        {
            ClusterOneSet<Robot>(env, rd, local, set_size, total, configs + s * set_size, labels + s * set_size, n_clusters + s, &st);
        }
        catch (const std::exception& e)
        {
#pragma omp critical
            failure = e.what();
        }
        calls += st.cluster_calls;
        fallbacks += st.cluster_fallback_calls;
    }
    if (!failure.empty()) throw std::invalid_argument(failure);
    if (stats)
    {
        stats->cluster_calls += calls;
        stats->cluster_fallback_calls += fallbacks;
    }
}

/* ---------------------------------------------------------------- per-cluster statistics */

static void QuaternionToRotation(const double q[4], Mat34& t)
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    t.m[0][0] = 1.0 - 2.0 * ((y * y) + (z * z));
    t.m[0][1] = 2.0 * ((x * y) - (z * w));
    t.m[0][2] = 2.0 * ((x * z) + (y * w));
    t.m[1][0] = 2.0 * ((x * y) + (z * w));
    t.m[1][1] = 1.0 - 2.0 * ((x * x) + (z * z));
    t.m[1][2] = 2.0 * ((y * z) - (x * w));
    t.m[2][0] = 2.0 * ((x * z) - (y * w));
    t.m[2][1] = 2.0 * ((y * z) + (x * w));
    t.m[2][2] = 1.0 - 2.0 * ((x * x) + (y * y));
}

/* AverageConfigurations: simple_robot_models.hpp:373-399 (SE2), :834-844 (SE3), :2093-2148 (linked) */
static void AverageConfigurations(const upc_robot_desc& rd, const std::vector<const double*>& members, double* out)
{
    const double inv_guard = (double)members.size();
    if (rd.kind == UPC_ROBOT_SE2)
    {
        double sx = 0.0, sy = 0.0, ss = 0.0, sc = 0.0;
        for (const double* c : members)
        {
            sx = sx + c[0];
            sy = sy + c[1];
            double s, k;
            upc_sincos(c[2], &s, &k);
            ss = ss + s;
            sc = sc + k;
        }
        out[0] = sx / inv_guard;
        out[1] = sy / inv_guard;
        out[2] = upc_atan2(ss, sc); /* circular mean (AverageContinuousRevolute) */
    }
    else if (rd.kind == UPC_ROBOT_SE3)
    {
        double st[3] = {0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0}, qref[4] = {0.0, 0.0, 0.0, 1.0};
        bool first = true;
        for (const double* c : members)
        {
            const Mat34 t = FromRows12(c);
            for (int a = 0; a < 3; a++) st[a] = st[a] + t.m[a][3];
            double q[4];
            QuaternionFromRotation(t, q);
            if (first) { for (int a = 0; a < 4; a++) qref[a] = q[a]; first = false; }
            const double dot = (((q[0] * qref[0]) + (q[1] * qref[1])) + (q[2] * qref[2])) + (q[3] * qref[3]);
            const double sign = (dot < 0.0) ? -1.0 : 1.0;
            for (int a = 0; a < 4; a++) sq[a] = sq[a] + (sign * q[a]);
        }
        const double norm = sqrt((((sq[0] * sq[0]) + (sq[1] * sq[1])) + (sq[2] * sq[2])) + (sq[3] * sq[3]));
        double q[4];
        for (int a = 0; a < 4; a++) q[a] = sq[a] / norm;
        Mat34 e;
        QuaternionToRotation(q, e);
        for (int a = 0; a < 3; a++) e.m[a][3] = st[a] / inv_guard;
        ToRows12(e, out);
    }
    else
    {
        int active = 0;
        for (int j = 0; j < rd.num_joints; j++)
        {
            if (rd.joints[j].type == UPC_JOINT_FIXED) continue;
            if (rd.joints[j].type == UPC_JOINT_CONTINUOUS)
            {
                double ss = 0.0, sc = 0.0;
                for (const double* c : members)
                {
                    double s, k;
                    upc_sincos(c[active], &s, &k);
                    ss = ss + s;
                    sc = sc + k;
                }
                out[active] = upc_atan2(ss, sc);
            }
            else
            {
                double sum = 0.0;
                for (const double* c : members) sum = sum + c[active];
                out[active] = sum / inv_guard;
            }
            active++;
        }
    }
}

/* ComputePerDimensionConfigurationDistance: simple_robot_models.hpp:419-422, :866-869, :2206-2209 (absolute values) */
static void PerDimensionDistance(const upc_robot_desc& rd, const double* c1, const double* c2, double* out)
{
    if (rd.kind == UPC_ROBOT_SE2)
    {
        out[0] = fabs(c2[0] - c1[0]);
        out[1] = fabs(c2[1] - c1[1]);
        out[2] = fabs(upc_angle_signed_distance(c1[2], c2[2]));
    }
    else if (rd.kind == UPC_ROBOT_SE3)
    {
        const Mat34 t1 = FromRows12(c1), t2 = FromRows12(c2);
        for (int a = 0; a < 3; a++) out[a] = fabs(t2.m[a][3] - t1.m[a][3]);
        double q1[4], q2[4];
        QuaternionFromRotation(t1, q1);
        QuaternionFromRotation(t2, q2);
        const double dq = fabs((((q1[0] * q2[0]) + (q1[1] * q2[1])) + (q1[2] * q2[2])) + (q1[3] * q2[3]));
        out[3] = (dq < (1.0 - 2.220446049250313e-16)) ? fabs(upc_acos(((2.0 * dq) * dq) - 1.0)) : 0.0;
    }
    else
    {
        int active = 0;
        for (int j = 0; j < rd.num_joints; j++)
        {
            if (rd.joints[j].type == UPC_JOINT_FIXED) continue;
            const double raw = (rd.joints[j].type == UPC_JOINT_CONTINUOUS) ? upc_angle_signed_distance(c1[active], c2[active]) : (c2[active] - c1[active]);
            out[active] = fabs(raw * fabs(rd.distance_weights[active]));
            active++;
        }
    }
}

template <typename Robot>
static void ClusterStatistics(const upc_robot_desc& rd, int64_t n, const double* configs, const int32_t* labels, const uint8_t* collided,
                              int32_t allow_contacts, double step_size, int32_t n_clusters, int64_t* counts, uint8_t* any_collided,
                              double* expectations, double* variance, double* si_variance, double* variances, double* si_variances)
{
    const Robot robot(rd);
    const int width = upc_config_width(rd.kind, rd.dof);
    const int vw = upc_variance_width(rd.kind, rd.dof);
    std::vector<double> aos((size_t)(n * width));
    for (int64_t i = 0; i < n; i++)
        for (int c = 0; c < width; c++) aos[(size_t)(i * width + c)] = configs[(int64_t)c * n + i];
    for (int32_t k = 0; k < n_clusters; k++)
    {
        std::vector<const double*> members;
        bool any = false;
        for (int64_t i = 0; i < n; i++)
        {
            if (labels[i] != k) continue;
            if ((collided[i] == 0) || allow_contacts) /* uncertainty_contact_planning.hpp:2019 */
            {
                members.push_back(&aos[(size_t)(i * width)]);
                if (collided[i]) any = true; /* :2132 */
            }
        }
        counts[k] = (int64_t)members.size();
        any_collided[k] = any ? 1 : 0;
        double* e = expectations + (int64_t)k * width;
        for (int c = 0; c < width; c++) e[c] = 0.0;
        variance[k] = 0.0;
        si_variance[k] = 0.0;
        for (int v = 0; v < vw; v++) { variances[(int64_t)k * vw + v] = 0.0; si_variances[(int64_t)k * vw + v] = 0.0; }
        if (members.empty()) continue;
        /* ComputeExpectation, uncertainty_planner_state.hpp:591-607 */
        if (members.size() == 1)
        {
            for (int c = 0; c < width; c++) e[c] = members[0][c];
        }
        else
        {
            AverageConfigurations(rd, members, e);
        }
        if (members.size() == 1)
        {
            /* :608-714: zero variances; the directional ones are dim_distance(p, p) */
            double dd[UPC_MAX_DOF];
            PerDimensionDistance(rd, members[0], members[0], dd);
            for (int v = 0; v < vw; v++) { variances[(int64_t)k * vw + v] = dd[v]; si_variances[(int64_t)k * vw + v] = dd[v]; }
            continue;
        }
        const double weight = 1.0 / (double)members.size();
        double var_sum = 0.0, si_sum = 0.0;
        for (const double* p : members)
        {
            const double raw = robot.ComputeConfigurationDistance(e, p);
            var_sum = var_sum + ((raw * raw) * weight);
            const double si = raw / step_size;
            si_sum = si_sum + ((si * si) * weight);
            double dd[UPC_MAX_DOF];
            PerDimensionDistance(rd, e, p, dd);
            for (int v = 0; v < vw; v++)
            {
                variances[(int64_t)k * vw + v] = variances[(int64_t)k * vw + v] + ((dd[v] * dd[v]) * weight);
                const double sd = dd[v] / step_size;
                si_variances[(int64_t)k * vw + v] = si_variances[(int64_t)k * vw + v] + ((sd * sd) * weight);
            }
        }
        variance[k] = var_sum;
        si_variance[k] = si_sum;
    }
}

static void ValidateRobot(const upc_robot_desc* r)
{
    if (!r) throw std::invalid_argument("null robot");
    if ((r->kind == UPC_ROBOT_SE2 && r->dof != 3) || (r->kind == UPC_ROBOT_SE3 && r->dof != 6)) throw std::invalid_argument("bad dof");
    if (r->dof < 1 || r->dof > UPC_MAX_DOF) throw std::invalid_argument("dof out of range");
    if (r->num_links < 1 || r->num_links > UPC_MAX_LINKS) throw std::invalid_argument("num_links out of range");
}
} /* namespace upc_oracle */

/* ---------------------------------------------------------------- extern "C" surface */

using namespace upc_oracle;

#define ORACLE_TRY try {
#define ORACLE_CATCH                                   \
    }                                                  \
    catch (const std::invalid_argument& e)             \
    {                                                  \
        g_last_error = e.what();                       \
        return UPC_ERR_INVALID_ARGUMENT;               \
    }                                                  \
    catch (const std::exception& e)                    \
    {                                                  \
        g_last_error = e.what();                       \
        return UPC_ERR_UNSUPPORTED;                    \
    }

extern "C" {

#ifndef UPC_ORACLE_NO_ABI_HELPERS
/* the oracle library is self-contained: it carries its own copies of the two size helpers */
int32_t upc_config_width(int32_t kind, int32_t dof)
{
    if (kind == UPC_ROBOT_SE2) return 3;
    if (kind == UPC_ROBOT_SE3) return 12;
    return dof;
}
int64_t upc_tape_doubles_per_particle(const upc_sim_params* p, int32_t dof)
{
    return (int64_t)p->num_steps * (int64_t)(1 + p->max_microsteps) * (int64_t)dof;
}
int32_t upc_variance_width(int32_t kind, int32_t dof)
{
    if (kind == UPC_ROBOT_SE2) return 3;
    if (kind == UPC_ROBOT_SE3) return 4;
    return dof;
}
#endif

const char* oracle_last_error(void) { return g_last_error.c_str(); }
int32_t oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int32_t n) { omp_set_num_threads(n); }

double oracle_pid_sequence(double kp, double ki, double kd, double clamp, const double* errors, const double* dts,
                           int32_t n, double* out_terms)
{
    Pid pid;
    pid.Initialize(kp, ki, kd, clamp);
    double last = 0.0;
    for (int32_t i = 0; i < n; i++)
    {
        last = pid.ComputeFeedbackTerm(errors[i], dts[i]);
        if (out_terms) out_terms[i] = last;
    }
    return last;
}

void oracle_math(int32_t which, int64_t n, const double* a, const double* b, double* out)
{
    for (int64_t i = 0; i < n; i++)
    {
        switch (which)
        {
            case 0: out[i] = upc_sin(a[i]); break;
            case 1: out[i] = upc_cos(a[i]); break;
            case 2: out[i] = upc_atan2(a[i], b[i]); break;
            case 3: out[i] = upc_acos(a[i]); break;
            case 4: out[i] = upc_wrap_angle(a[i]); break;
            default: out[i] = upc_angle_signed_distance(a[i], b[i]); break;
        }
    }
}

/* sensor / actuator arithmetic on replayed draws (pinned against the reference's own header by tests/test_oracle_unc.py) */
void oracle_sensor_values(int64_t n, const double* process_values, const double* draws, double* out)
{
    for (int64_t i = 0; i < n; i++) out[i] = SensorValue(process_values[i], draws[i]);
}

void oracle_actuator_values(double noise_bound, double actuator_limit, int64_t n, const double* inputs, const double* draws, double* out, double* noiseless)
{
    Actuator a;
    a.Initialize(noise_bound, actuator_limit);
    for (int64_t i = 0; i < n; i++)
    {
        out[i] = a.GetControlValue(inputs[i], draws[i]);
        noiseless[i] = a.GetControlValue(inputs[i]);
    }
}

void oracle_se3_log(const double* cfg12, double* twist6) { Se3Log(FromRows12(cfg12), twist6); }
void oracle_quaternion_from_rotation(const double* cfg12, double* q4) { QuaternionFromRotation(FromRows12(cfg12), q4); }
void oracle_se3_exp(const double* twist6, double* cfg12) { ToRows12(Se3Exp(twist6), cfg12); }

int32_t oracle_forward_simulate(const upc_env_desc* env_desc, const upc_robot_desc* robot, const upc_sim_params* params,
                                int64_t n, const double* starts, int64_t n_targets, const double* targets,
                                const double* tape, double* out_configs, uint8_t* out_collided, upc_stats* stats)
{
    ORACLE_TRY
    ValidateRobot(robot);
    if (!env_desc || !params) throw std::invalid_argument("null argument");
    if (!(n_targets == 1 || n_targets == n)) throw std::invalid_argument("targets must have size 1 or n");
    if (params->max_microsteps < 1 || params->num_steps < 1) throw std::invalid_argument("bad step counts");
    const Environment env(*env_desc);
    if (n == 0) return UPC_OK;
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: ForwardSimulateRobots<Se2Robot>(env, *robot, *params, n, starts, n_targets, targets, tape, out_configs, out_collided, stats); break;
        case UPC_ROBOT_SE3: ForwardSimulateRobots<Se3Robot>(env, *robot, *params, n, starts, n_targets, targets, tape, out_configs, out_collided, stats); break;
        case UPC_ROBOT_LINKED: ForwardSimulateRobots<LinkedRobot>(env, *robot, *params, n, starts, n_targets, targets, tape, out_configs, out_collided, stats); break;
        default: throw std::invalid_argument("unknown robot kind");
    }
    return UPC_OK;
    ORACLE_CATCH
}

extern "C++" {
template <typename Robot>
static void TracedImpl(const Environment& env, const upc_robot_desc& rd, const upc_sim_params& sp, int64_t n, const double* starts, int64_t n_targets,
                       const double* targets, const double* tape, double* out_configs, uint8_t* out_collided, double* trace_configs,
                       double* trace_controls, int32_t* trace_steps)
{
    const Robot prototype(rd);
    const int width = upc_config_width(rd.kind, rd.dof), dof = rd.dof;
    for (int64_t i = 0; i < n; i++)
    {
        Robot robot = prototype;
        double start[16], target[16], result[16] = {0.0};
        for (int c = 0; c < width; c++)
        {
            start[c] = starts[(int64_t)c * n + i];
            target[c] = targets[(int64_t)c * n_targets + ((n_targets == 1) ? 0 : i)];
        }
        robot.ResetPosition(start);
        SimStats st = {0, 0, 0, 0};
        ParticleTrace tr;
        const bool collided = SimulateParticle(env, robot, sp, target, tape + i, n, st, &tr);
        robot.GetPosition(result);
        for (int c = 0; c < width; c++) out_configs[(int64_t)c * n + i] = result[c];
        out_collided[i] = collided ? 1 : 0;
        for (size_t s = 0; s < tr.configs.size(); s++)
        {
            for (int c = 0; c < width; c++) trace_configs[((int64_t)s * width + c) * n + i] = tr.configs[s][(size_t)c];
            for (int d = 0; d < dof; d++)
            {
                trace_controls[((int64_t)s * 2 * dof + d) * n + i] = tr.control_input[s][(size_t)d];
                trace_controls[((int64_t)s * 2 * dof + dof + d) * n + i] = tr.control_input_step[s][(size_t)d];
            }
        }
        trace_steps[i] = (int32_t)tr.configs.size();
        trace_steps[n + i] = tr.last_microsteps;
    }
}
}

/* same outputs as upc_forward_simulate_traced: trace_configs [S][C][n], trace_controls [S][2*dof][n], trace_steps [2][n]; rows
 * of intervals a particle did not execute are left untouched (pass zeroed buffers) */
int32_t oracle_forward_simulate_traced(const upc_env_desc* env_desc, const upc_robot_desc* robot, const upc_sim_params* params, int64_t n,
                                       const double* starts, int64_t n_targets, const double* targets, const double* tape, double* out_configs,
                                       uint8_t* out_collided, double* trace_configs, double* trace_controls, int32_t* trace_steps)
{
    ORACLE_TRY
    ValidateRobot(robot);
    if (!env_desc || !params) throw std::invalid_argument("null argument");
    if (!(n_targets == 1 || n_targets == n)) throw std::invalid_argument("targets must have size 1 or n");
    const Environment env(*env_desc);
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: TracedImpl<Se2Robot>(env, *robot, *params, n, starts, n_targets, targets, tape, out_configs, out_collided, trace_configs, trace_controls, trace_steps); break;
        case UPC_ROBOT_SE3: TracedImpl<Se3Robot>(env, *robot, *params, n, starts, n_targets, targets, tape, out_configs, out_collided, trace_configs, trace_controls, trace_steps); break;
        default: TracedImpl<LinkedRobot>(env, *robot, *params, n, starts, n_targets, targets, tape, out_configs, out_collided, trace_configs, trace_controls, trace_steps); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_check_config_collision(const upc_env_desc* env_desc, const upc_robot_desc* robot, const double* config,
                                      double contact_distance, double inflation, int32_t* out)
{
    ORACLE_TRY
    ValidateRobot(robot);
    const Environment env(*env_desc);
    const double thr = contact_distance + inflation;
    bool hit = false;
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: { Se2Robot r(*robot); r.UpdatePosition(config); hit = CheckCollision(env, r, thr); break; }
        case UPC_ROBOT_SE3: { Se3Robot r(*robot); r.UpdatePosition(config); hit = CheckCollision(env, r, thr); break; }
        default: { LinkedRobot r(*robot); r.UpdatePosition(config); hit = CheckCollision(env, r, thr); break; }
    }
    *out = hit ? 1 : 0;
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_cluster_particles(const upc_env_desc* env_desc, const upc_robot_desc* robot, const upc_cluster_params* params,
                                 int64_t n_sets, int64_t set_size, const double* configs, int32_t* labels,
                                 int32_t* n_clusters, upc_stats* stats)
{
    ORACLE_TRY
    ValidateRobot(robot);
    const Environment env(*env_desc);
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: ClusterSets<Se2Robot>(env, *robot, *params, n_sets, set_size, configs, labels, n_clusters, stats); break;
        case UPC_ROBOT_SE3: ClusterSets<Se3Robot>(env, *robot, *params, n_sets, set_size, configs, labels, n_clusters, stats); break;
        default: ClusterSets<LinkedRobot>(env, *robot, *params, n_sets, set_size, configs, labels, n_clusters, stats); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

extern "C++" {
template <typename Robot>
static void PairDistancesImpl(const upc_robot_desc& rd, int64_t n, const double* configs, double* matrix)
{
    const Robot robot(rd);
    const int width = upc_config_width(rd.kind, rd.dof);
    std::vector<double> aos((size_t)(n * width));
    for (int64_t i = 0; i < n; i++)
        for (int c = 0; c < width; c++) aos[(size_t)(i * width + c)] = configs[(int64_t)c * n + i];
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++)
    {
        matrix[i * n + i] = 0.0;
        for (int64_t j = i + 1; j < n; j++)
        {
            const double d = robot.ComputeConfigurationDistance(&aos[(size_t)(i * width)], &aos[(size_t)(j * width)]);
            matrix[i * n + j] = d;
            matrix[j * n + i] = d;
        }
    }
}
}

int32_t oracle_pair_distances(const upc_robot_desc* robot, int64_t n, const double* configs, double* matrix)
{
    ORACLE_TRY
    ValidateRobot(robot);
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: PairDistancesImpl<Se2Robot>(*robot, n, configs, matrix); break;
        case UPC_ROBOT_SE3: PairDistancesImpl<Se3Robot>(*robot, n, configs, matrix); break;
        default: PairDistancesImpl<LinkedRobot>(*robot, n, configs, matrix); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

extern "C++" {
template <typename Robot>
static void PairDistancesReferenceStyle(const upc_robot_desc& rd, int64_t n, const double* configs, double* matrix)
{
    /* mirrors the reference's cost structure: std::function indirection per pair and a heap-allocated
     * per-dimension delta vector per pair (Eigen::VectorXd at simple_robot_models.hpp:412 / :2195) */
    const Robot robot(rd);
    const int width = upc_config_width(rd.kind, rd.dof);
    std::vector<std::vector<double>> particles((size_t)n, std::vector<double>((size_t)width));
    for (int64_t i = 0; i < n; i++)
        for (int c = 0; c < width; c++) particles[(size_t)i][(size_t)c] = configs[(int64_t)c * n + i];
    std::function<double(const int64_t&, const int64_t&)> distance_fn = [&](const int64_t& a, const int64_t& b) {
        std::unique_ptr<std::vector<double>> scratch(new std::vector<double>((size_t)width));
        (*scratch)[0] = particles[(size_t)a][0];
        return robot.ComputeConfigurationDistance(particles[(size_t)a].data(), particles[(size_t)b].data()) + 0.0 * (*scratch)[0];
    };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        for (int64_t j = i; j < n; j++)
        {
            if (i != j)
            {
                const double d = distance_fn(i, j);
                matrix[i * n + j] = d;
                matrix[j * n + i] = d;
            }
            else
            {
                matrix[i * n + j] = 0.0;
            }
        }
    }
}
}

int32_t oracle_pair_distances_reference_style(const upc_robot_desc* robot, int64_t n, const double* configs, double* matrix)
{
    ORACLE_TRY
    ValidateRobot(robot);
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: PairDistancesReferenceStyle<Se2Robot>(*robot, n, configs, matrix); break;
        case UPC_ROBOT_SE3: PairDistancesReferenceStyle<Se3Robot>(*robot, n, configs, matrix); break;
        default: PairDistancesReferenceStyle<LinkedRobot>(*robot, n, configs, matrix); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

extern "C++" {
template <typename Robot>
static void SignaturesImpl(const Environment& env, const upc_robot_desc& rd, int64_t n, const double* configs, uint32_t* sigs)
{
    const Robot prototype(rd);
    const int width = upc_config_width(rd.kind, rd.dof);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        Robot robot = prototype;
        double cfg[16];
        for (int c = 0; c < width; c++) cfg[c] = configs[(int64_t)c * n + i];
        RegionSignature(env, robot, cfg, sigs + i * rd.num_points);
    }
}
}

int32_t oracle_region_signatures(const upc_env_desc* env_desc, const upc_robot_desc* robot, int64_t n, const double* configs,
                                 uint32_t* signatures)
{
    ORACLE_TRY
    ValidateRobot(robot);
    const Environment env(*env_desc);
    if (!env.convex_segment) throw std::invalid_argument("convex_segment grid required");
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: SignaturesImpl<Se2Robot>(env, *robot, n, configs, signatures); break;
        case UPC_ROBOT_SE3: SignaturesImpl<Se3Robot>(env, *robot, n, configs, signatures); break;
        default: SignaturesImpl<LinkedRobot>(env, *robot, n, configs, signatures); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_sdf_lookup(const upc_env_desc* env_desc, int64_t n, const double* world_points, double* dist, double* grad)
{
    ORACLE_TRY
    const Environment env(*env_desc);
    const Mat34 V = VoxelFromLink(env, Identity34());
    for (int64_t i = 0; i < n; i++)
    {
        double u[3];
        TransformPoint(V, world_points + 3 * i, u);
        SdfSample(env, u, dist + i, grad ? (grad + 3 * i) : nullptr);
    }
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_cluster_statistics(const upc_robot_desc* robot, int64_t n, const double* configs, const int32_t* labels, const uint8_t* collided,
                                  int32_t allow_contacts, double step_size, int32_t n_clusters, int64_t* counts, uint8_t* any_collided,
                                  double* expectations, double* variance, double* si_variance, double* variances, double* si_variances)
{
    ORACLE_TRY
    ValidateRobot(robot);
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: ClusterStatistics<Se2Robot>(*robot, n, configs, labels, collided, allow_contacts, step_size, n_clusters, counts, any_collided, expectations, variance, si_variance, variances, si_variances); break;
        case UPC_ROBOT_SE3: ClusterStatistics<Se3Robot>(*robot, n, configs, labels, collided, allow_contacts, step_size, n_clusters, counts, any_collided, expectations, variance, si_variance, variances, si_variances); break;
        default: ClusterStatistics<LinkedRobot>(*robot, n, configs, labels, collided, allow_contacts, step_size, n_clusters, counts, any_collided, expectations, variance, si_variance, variances, si_variances); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

extern "C++" {
template <typename Robot>
static void NearestNeighbors(const upc_robot_desc& rd, int64_t n_states, const double* expectations, const double* fw, const double* vw,
                             const uint8_t* use, int64_t n_queries, const double* queries, double step_size, int64_t* out_index, double* out_distance)
{
    const Robot robot(rd);
    const int width = upc_config_width(rd.kind, rd.dof);
    std::vector<double> e((size_t)width), q((size_t)width);
    for (int64_t qi = 0; qi < n_queries; qi++)
    {
        for (int c = 0; c < width; c++) q[(size_t)c] = queries[(int64_t)c * n_queries + qi];
        int64_t best_index = -1;
        double best_distance = INFINITY;
        for (int64_t i = 0; i < n_states; i++) /* uncertainty_contact_planning.hpp:1513-1530, one thread's view */
        {
            if (use != nullptr && use[i] == 0) continue;
            for (int c = 0; c < width; c++) e[(size_t)c] = expectations[(int64_t)c * n_states + i];
            const double expectation_distance = robot.ComputeConfigurationDistance(e.data(), q.data()) / step_size; /* :1491 */
            const double distance = (fw[i] * expectation_distance) * vw[i];                                      /* :1500 */
            if (distance < best_distance)
            {
                best_index = i;
                best_distance = distance;
            }
        }
        out_index[qi] = best_index;
        if (out_distance) out_distance[qi] = best_distance;
    }
}
}

int32_t oracle_nearest_neighbors(const upc_robot_desc* robot, int64_t n_states, const double* expectations, const double* feasibility_weight,
                                 const double* variance_weight, const uint8_t* use_for_nn, int64_t n_queries, const double* queries,
                                 double step_size, int64_t* out_index, double* out_distance)
{
    ORACLE_TRY
    ValidateRobot(robot);
    switch (robot->kind)
    {
        case UPC_ROBOT_SE2: NearestNeighbors<Se2Robot>(*robot, n_states, expectations, feasibility_weight, variance_weight, use_for_nn, n_queries, queries, step_size, out_index, out_distance); break;
        case UPC_ROBOT_SE3: NearestNeighbors<Se3Robot>(*robot, n_states, expectations, feasibility_weight, variance_weight, use_for_nn, n_queries, queries, step_size, out_index, out_distance); break;
        default: NearestNeighbors<LinkedRobot>(*robot, n_states, expectations, feasibility_weight, variance_weight, use_for_nn, n_queries, queries, step_size, out_index, out_distance); break;
    }
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_complete_link(const double* matrix, int64_t n, double threshold, int32_t* labels, int32_t* n_clusters)
{
    ORACLE_TRY
    const std::vector<double> D(matrix, matrix + n * n);
    const auto clusters = CompleteLinkCluster(D, n, threshold);
    for (size_t k = 0; k < clusters.size(); k++)
        for (int64_t i : clusters[k]) labels[i] = (int32_t)k;
    *n_clusters = (int32_t)clusters.size();
    return UPC_OK;
    ORACLE_CATCH
}

/* ---- robot-model pieces exposed one call at a time, for tests/test_oracle_rob.py (the reference's simple_robot_models.hpp,
 * compiled unmodified into oracle/_ref/libref_rob.so, answers the same questions through oracle/ref_rob_shim.cpp) */
extern "C++" {
template <typename Robot>
static void RobSequence(const upc_robot_desc& rd, int32_t n_seq, int32_t n_steps, double dt, int32_t noisy, const double* starts, const double* targets,
                        const double* sensor_draws, const double* actuator_draws, double* controls, double* configs)
{
    const int width = upc_config_width(rd.kind, rd.dof);
    Robot robot(rd);
    const double zeros[UPC_MAX_DOF] = {0.0};
    for (int32_t s = 0; s < n_seq; s++)
    {
        robot.ResetPosition(starts + (size_t)s * width);
        for (int32_t k = 0; k < n_steps; k++)
        {
            const size_t at = ((size_t)s * n_steps + k) * rd.dof;
            double* u = controls + at;
            robot.GenerateControlAction(targets + (size_t)s * width, dt, noisy ? sensor_draws + at : zeros, u);
            double input[UPC_MAX_DOF];
            for (int a = 0; a < rd.dof; a++) input[a] = u[a] * dt;
            robot.ApplyControlInput(input, noisy ? actuator_draws + at : nullptr);
            robot.GetPosition(configs + ((size_t)s * n_steps + k) * width);
        }
    }
}

template <typename Robot>
static void RobKinematics(const upc_robot_desc& rd, int32_t n, const double* configs, double* link_transforms, double* jacobians)
{
    const int width = upc_config_width(rd.kind, rd.dof);
    Robot robot(rd);
    for (int32_t i = 0; i < n; i++)
    {
        robot.UpdatePosition(configs + (size_t)i * width);
        for (int l = 0; l < rd.num_links; l++)
        {
            const Mat34& T = robot.LinkTransform(l);
            double* out = link_transforms + ((size_t)i * rd.num_links + l) * 12;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++) out[4 * r + c] = T.m[r][c];
            for (int p = rd.link_point_offset[l]; p < rd.link_point_offset[l + 1]; p++)
            {
                const double local[3] = {rd.points[4 * p + 0], rd.points[4 * p + 1], rd.points[4 * p + 2]};
                double world[3];
                TransformPoint(T, local, world);
                double J[3][UPC_MAX_DOF];
                robot.PointJacobian(l, local, world, J);
                for (int r = 0; r < 3; r++)
                    for (int a = 0; a < rd.dof; a++) jacobians[(((size_t)i * rd.num_points + p) * 3 + r) * rd.dof + a] = J[r][a];
            }
        }
    }
}

template <typename Robot>
static void RobDistances(const upc_robot_desc& rd, int32_t n, const double* a, const double* b, double* dist, double* dim_dist)
{
    const int width = upc_config_width(rd.kind, rd.dof);
    const int dim_width = upc_variance_width(rd.kind, rd.dof);
    const Robot robot(rd);
    for (int32_t i = 0; i < n; i++)
    {
        dist[i] = robot.ComputeConfigurationDistance(a + (size_t)i * width, b + (size_t)i * width);
        PerDimensionDistance(rd, a + (size_t)i * width, b + (size_t)i * width, dim_dist + (size_t)i * dim_width);
    }
}
}

#define ROB_DISPATCH(fn, ...)                                         \
    switch (robot->kind)                                              \
    {                                                                 \
        case UPC_ROBOT_SE2: fn<Se2Robot>(*robot, __VA_ARGS__); break; \
        case UPC_ROBOT_SE3: fn<Se3Robot>(*robot, __VA_ARGS__); break; \
        default: fn<LinkedRobot>(*robot, __VA_ARGS__); break;         \
    }

/* n_seq independent edges: ResetPosition(start); n_steps x { u = GenerateControlAction(target, dt, sensor draws); ApplyControlInput(u * dt, actuator draws) };
 * draws [n_seq][n_steps][dof] each, already scaled as on a tape; noisy = 0 uses the noise-free forms */
int32_t oracle_rob_sequence(const upc_robot_desc* robot, int32_t n_seq, int32_t n_steps, double dt, int32_t noisy, const double* starts, const double* targets,
                            const double* sensor_draws, const double* actuator_draws, double* controls, double* configs)
{
    ORACLE_TRY
    ValidateRobot(robot);
    ROB_DISPATCH(RobSequence, n_seq, n_steps, dt, noisy, starts, targets, sensor_draws, actuator_draws, controls, configs)
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_rob_kinematics(const upc_robot_desc* robot, int32_t n, const double* configs, double* link_transforms, double* jacobians)
{
    ORACLE_TRY
    ValidateRobot(robot);
    ROB_DISPATCH(RobKinematics, n, configs, link_transforms, jacobians)
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_rob_distances(const upc_robot_desc* robot, int32_t n, const double* a, const double* b, double* dist, double* dim_dist)
{
    ORACLE_TRY
    ValidateRobot(robot);
    ROB_DISPATCH(RobDistances, n, a, b, dist, dim_dist)
    return UPC_OK;
    ORACLE_CATCH
}

int32_t oracle_rob_average(const upc_robot_desc* robot, int32_t n, const double* configs, double* out)
{
    ORACLE_TRY
    ValidateRobot(robot);
    const int width = upc_config_width(robot->kind, robot->dof);
    std::vector<const double*> members;
    for (int32_t i = 0; i < n; i++) members.push_back(configs + (size_t)i * width);
    AverageConfigurations(*robot, members, out);
    return UPC_OK;
    ORACLE_CATCH
}
}
