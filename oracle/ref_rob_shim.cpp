/*
 * ref_rob_shim.cpp - extern "C" wrapper around the reference's OWN, UNMODIFIED simple_robot_models.hpp
 * (SimpleSE2Robot :92-431, SimpleSE3Robot :507-875, SimpleJointModel :877-1151, SimpleLinkedRobot :1274-2218), compiled against
 * the stand-in Eigen / arc_utilities headers of oracle/ref_standins (neither library is vendored by the reference or installed
 * here; the stand-ins follow the libraries' documented semantics with libm, independently of the oracle's arithmetic).
 * Built by oracle/Makefile into oracle/_ref/libref_rob.so from the header where it lies under $(REFERENCE)/include; used by
 * tests/test_oracle_rob.py to pin the oracle's robot models (sense -> PID -> clamp -> noisy apply sequences, forward
 * kinematics, point Jacobians, configuration distances, averages) against the reference's code, and - through the reference's
 * own, UNMODIFIED uncertainty_planner_state.hpp (UncertaintyPlannerState::UpdateStatistics :339-351, ComputeExpectation /
 * ComputeVariance / ComputeDirectionalVariance / the space-independent forms :591-714) - the per-cluster statistics of the child
 * states (oracle_cluster_statistics, the product's upc_cluster_statistics).  TEST INFRASTRUCTURE ONLY.
 *
 * Robots are described by the product's POD upc_robot_desc (include/upc_b200.h) so that the same Python fixtures drive both
 * sides.  The reference's SE(2) / SE(3) robots take ONE parameter set for the translational axes and one for the rotational
 * axes (SE2_ROBOT_CONFIG :38-90, SE3_ROBOT_CONFIG :453-505); this shim reads them from axis 0 and from the last axis.
 */
#include <uncertainty_planning_core/simple_robot_models.hpp>
#include <uncertainty_planning_core/uncertainty_planner_state.hpp>
#include "upc_b200.h"

#include <string>
#include <vector>

namespace
{
typedef simple_uncertainty_models::TruncatedNormalUncertainVelocityActuator Actuator;
typedef simple_robot_models::SimpleLinkedRobot<Actuator> LinkedRobot;
typedef arc_helpers::StandardDrawTape Tape;

std::shared_ptr<EigenHelpers::VectorVector4d> LinkPoints(const upc_robot_desc& d, int link)
{
    std::shared_ptr<EigenHelpers::VectorVector4d> points(new EigenHelpers::VectorVector4d());
    for (int p = d.link_point_offset[link]; p < d.link_point_offset[link + 1]; p++)
    {
        points->push_back(Eigen::Vector4d(d.points[4 * p + 0], d.points[4 * p + 1], d.points[4 * p + 2], 1.0));
    }
    return points;
}

Eigen::Isometry3d FromRowMajor34(const double* v)
{
    Eigen::Isometry3d t = Eigen::Isometry3d::Identity();
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) t.matrix()(r, c) = v[4 * r + c];
    return t;
}

/* SE(3) configurations cross as 12 doubles: R row-major, then t (the product's layout) */
Eigen::Isometry3d FromRows12(const double* v)
{
    Eigen::Isometry3d t = Eigen::Isometry3d::Identity();
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++) t.matrix()(r, c) = v[3 * r + c];
        t.matrix()(r, 3) = v[9 + r];
    }
    return t;
}

void ToRows12(const Eigen::Isometry3d& t, double* v)
{
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++) v[3 * r + c] = t(r, c);
        v[9 + r] = t(r, 3);
    }
}

void ToRowMajor34(const Eigen::Isometry3d& t, double* v)
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) v[4 * r + c] = t(r, c);
}

simple_robot_models::SimpleSE2Robot MakeSe2(const upc_robot_desc& d)
{
    const upc_axis_params& a = d.axis[0];
    const upc_axis_params& r = d.axis[2];
    const simple_robot_models::SE2_ROBOT_CONFIG config(a.kp, a.ki, a.kd, a.integral_clamp, a.velocity_limit, a.max_sensor_noise, a.max_actuator_noise,
                                                       r.kp, r.ki, r.kd, r.integral_clamp, r.velocity_limit, r.max_sensor_noise, r.max_actuator_noise);
    return simple_robot_models::SimpleSE2Robot(LinkPoints(d, 0), Eigen::Matrix<double, 3, 1>::Zero(), config);
}

simple_robot_models::SimpleSE3Robot MakeSe3(const upc_robot_desc& d)
{
    const upc_axis_params& a = d.axis[0];
    const upc_axis_params& r = d.axis[5];
    const simple_robot_models::SE3_ROBOT_CONFIG config(a.kp, a.ki, a.kd, a.integral_clamp, a.velocity_limit, a.max_sensor_noise, a.max_actuator_noise,
                                                       r.kp, r.ki, r.kd, r.integral_clamp, r.velocity_limit, r.max_sensor_noise, r.max_actuator_noise);
    return simple_robot_models::SimpleSE3Robot(LinkPoints(d, 0), Eigen::Isometry3d::Identity(), config);
}

std::string LinkName(int l) { return "link" + std::to_string(l); }

simple_robot_models::SimpleLinkedConfiguration MakeLinkedConfig(const upc_robot_desc& d, const double* values)
{
    simple_robot_models::SimpleLinkedConfiguration config;
    int active = 0;
    for (int j = 0; j < d.num_joints; j++)
    {
        const upc_joint_desc& jd = d.joints[j];
        if (jd.type == UPC_JOINT_FIXED) continue;
        config.push_back(simple_robot_models::SimpleJointModel(std::make_pair(jd.lower_limit, jd.upper_limit), values ? values[active] : 0.0,
                                                               (simple_robot_models::SimpleJointModel::JOINT_TYPE)jd.type));
        active++;
    }
    return config;
}

LinkedRobot MakeLinked(const upc_robot_desc& d)
{
    std::vector<simple_robot_models::RobotLink> links;
    for (int l = 0; l < d.num_links; l++) links.push_back(simple_robot_models::RobotLink(LinkPoints(d, l), LinkName(l)));
    std::vector<simple_robot_models::RobotJoint<Actuator>> joints;
    std::vector<double> weights;
    int active = 0;
    for (int j = 0; j < d.num_joints; j++)
    {
        const upc_joint_desc& jd = d.joints[j];
        simple_robot_models::RobotJoint<Actuator> joint;
        joint.name = "joint" + std::to_string(j);
        joint.parent_link_index = jd.parent_link;
        joint.child_link_index = jd.child_link;
        joint.joint_axis = Eigen::Vector3d(jd.axis[0], jd.axis[1], jd.axis[2]);
        joint.joint_transform = FromRowMajor34(jd.transform);
        if (jd.type == UPC_JOINT_FIXED)
        {
            joint.joint_model = simple_robot_models::SimpleJointModel();
        }
        else
        {
            const upc_axis_params& a = d.axis[active];
            joint.joint_model = simple_robot_models::SimpleJointModel(std::make_pair(jd.lower_limit, jd.upper_limit), 0.0, (simple_robot_models::SimpleJointModel::JOINT_TYPE)jd.type);
            const simple_robot_models::LINKED_ROBOT_CONFIG config(a.kp, a.ki, a.kd, a.integral_clamp, a.velocity_limit, a.max_sensor_noise, a.max_actuator_noise);
            joint.joint_controller = simple_robot_models::JointControllerGroup<Actuator>(config, Actuator(a.max_actuator_noise, a.velocity_limit));
            weights.push_back(d.distance_weights[active]);
            active++;
        }
        joints.push_back(joint);
    }
    return LinkedRobot(FromRowMajor34(d.base_transform), links, joints, std::vector<std::pair<size_t, size_t>>(), MakeLinkedConfig(d, nullptr), weights);
}

Eigen::Matrix<double, 3, 1> Se2Config(const double* v)
{
    Eigen::Matrix<double, 3, 1> c;
    c(0) = v[0];
    c(1) = v[1];
    c(2) = v[2];
    return c;
}

/* one adapter per robot type: configurations as flat doubles in the product's layout */
struct Se2Adapter
{
    typedef simple_robot_models::SimpleSE2Robot Robot;
    typedef Eigen::Matrix<double, 3, 1> Config;
    static const int kWidth = 3;
    static Robot Make(const upc_robot_desc& d) { return MakeSe2(d); }
    static Config In(const upc_robot_desc&, const double* v) { return Se2Config(v); }
    static void Out(const Config& c, double* v)
    {
        for (int i = 0; i < 3; i++) v[i] = c(i);
    }
    static std::string Link(int) { return "robot"; }
};

struct Se3Adapter
{
    typedef simple_robot_models::SimpleSE3Robot Robot;
    typedef Eigen::Isometry3d Config;
    static const int kWidth = 12;
    static Robot Make(const upc_robot_desc& d) { return MakeSe3(d); }
    static Config In(const upc_robot_desc&, const double* v) { return FromRows12(v); }
    static void Out(const Config& c, double* v) { ToRows12(c, v); }
    static std::string Link(int) { return "robot"; }
};

struct LinkedAdapter
{
    typedef LinkedRobot Robot;
    typedef simple_robot_models::SimpleLinkedConfiguration Config;
    static Robot Make(const upc_robot_desc& d) { return MakeLinked(d); }
    static Config In(const upc_robot_desc& d, const double* v) { return MakeLinkedConfig(d, v); }
    static void Out(const Config& c, double* v)
    {
        for (size_t i = 0; i < c.size(); i++) v[i] = c[i].GetValue();
    }
    static std::string Link(int l) { return LinkName(l); }
};

int Width(const upc_robot_desc& d) { return d.kind == UPC_ROBOT_SE2 ? 3 : (d.kind == UPC_ROBOT_SE3 ? 12 : d.dof); }

template <typename A>
void Sequence(const upc_robot_desc& d, int n_seq, int n_steps, double dt, int noisy, const double* starts, const double* targets, const double* z,
              double* controls, double* configs)
{
    const int width = Width(d);
    typename A::Robot robot = A::Make(d);
    for (int s = 0; s < n_seq; s++)
    {
        robot.ResetPosition(A::In(d, starts + (size_t)s * width));
        const typename A::Config target = A::In(d, targets + (size_t)s * width);
        Tape tape = {z + (size_t)s * n_steps * 2 * d.dof, (size_t)n_steps * 2 * d.dof, 0};
        for (int k = 0; k < n_steps; k++)
        {
            /* standard draws are laid out [sequence][step][sensor | actuator][axis]; the noise-free overloads skip their slot */
            tape.next = (size_t)k * 2 * d.dof;
            const Eigen::VectorXd control = noisy ? robot.GenerateControlAction(target, dt, tape) : robot.GenerateControlAction(target, dt);
            assert(control.size() == d.dof);
            for (int a = 0; a < d.dof; a++) controls[((size_t)s * n_steps + k) * d.dof + a] = control(a);
            tape.next = (size_t)(k * 2 + 1) * d.dof;
            const Eigen::VectorXd input = control * dt;
            if (noisy)
                robot.ApplyControlInput(input, tape);
            else
                robot.ApplyControlInput(input);
            A::Out(robot.GetPosition(), configs + ((size_t)s * n_steps + k) * width);
        }
    }
}

template <typename A>
void Kinematics(const upc_robot_desc& d, int n, const double* configs, double* link_transforms, double* jacobians)
{
    const int width = Width(d);
    typename A::Robot robot = A::Make(d);
    for (int i = 0; i < n; i++)
    {
        robot.UpdatePosition(A::In(d, configs + (size_t)i * width));
        for (int l = 0; l < d.num_links; l++)
        {
            ToRowMajor34(robot.GetLinkTransform(A::Link(l)), link_transforms + ((size_t)i * d.num_links + l) * 12);
            for (int p = d.link_point_offset[l]; p < d.link_point_offset[l + 1]; p++)
            {
                const Eigen::Vector4d point(d.points[4 * p + 0], d.points[4 * p + 1], d.points[4 * p + 2], 1.0);
                const Eigen::Matrix<double, 3, Eigen::Dynamic> J = robot.ComputeLinkPointJacobian(A::Link(l), point);
                assert(J.cols() == d.dof);
                for (int r = 0; r < 3; r++)
                    for (int a = 0; a < d.dof; a++) jacobians[(((size_t)i * d.num_points + p) * 3 + r) * d.dof + a] = J(r, a);
            }
        }
    }
}

template <typename A>
void Distances(const upc_robot_desc& d, int n, const double* a, const double* b, double* dist, double* dim_dist, int dim_width)
{
    const int width = Width(d);
    const typename A::Robot robot = A::Make(d);
    for (int i = 0; i < n; i++)
    {
        const typename A::Config c1 = A::In(d, a + (size_t)i * width);
        const typename A::Config c2 = A::In(d, b + (size_t)i * width);
        dist[i] = robot.ComputeConfigurationDistance(c1, c2);
        const Eigen::VectorXd dims = robot.ComputePerDimensionConfigurationDistance(c1, c2);
        assert(dims.size() == dim_width);
        for (int k = 0; k < dim_width; k++) dim_dist[(size_t)i * dim_width + k] = dims(k);
    }
}
}

extern "C" {

/* n_seq independent edges: ResetPosition(start); n_steps x { u = GenerateControlAction(target, dt[, rng]); ApplyControlInput(u * dt[, rng]) };
 * controls [n_seq][n_steps][dof], configs [n_seq][n_steps][width], standard draws z [n_seq][n_steps][2][dof] */
int ref_rob_sequence(const upc_robot_desc* d, int n_seq, int n_steps, double dt, int noisy, const double* starts, const double* targets, const double* z,
                     double* controls, double* configs)
{
    if (d->kind == UPC_ROBOT_SE2) Sequence<Se2Adapter>(*d, n_seq, n_steps, dt, noisy, starts, targets, z, controls, configs);
    else if (d->kind == UPC_ROBOT_SE3) Sequence<Se3Adapter>(*d, n_seq, n_steps, dt, noisy, starts, targets, z, controls, configs);
    else Sequence<LinkedAdapter>(*d, n_seq, n_steps, dt, noisy, starts, targets, z, controls, configs);
    return 0;
}

/* per configuration: every link's transform (row-major 3x4) and ComputeLinkPointJacobian of every body point ([3][dof]) */
int ref_rob_kinematics(const upc_robot_desc* d, int n, const double* configs, double* link_transforms, double* jacobians)
{
    if (d->kind == UPC_ROBOT_SE2) Kinematics<Se2Adapter>(*d, n, configs, link_transforms, jacobians);
    else if (d->kind == UPC_ROBOT_SE3) Kinematics<Se3Adapter>(*d, n, configs, link_transforms, jacobians);
    else Kinematics<LinkedAdapter>(*d, n, configs, link_transforms, jacobians);
    return 0;
}

/* ComputeConfigurationDistance and ComputePerDimensionConfigurationDistance of n pairs (dim_width = 3 / 4 / dof) */
int ref_rob_distances(const upc_robot_desc* d, int n, const double* a, const double* b, double* dist, double* dim_dist, int dim_width)
{
    if (d->kind == UPC_ROBOT_SE2) Distances<Se2Adapter>(*d, n, a, b, dist, dim_dist, dim_width);
    else if (d->kind == UPC_ROBOT_SE3) Distances<Se3Adapter>(*d, n, a, b, dist, dim_dist, dim_width);
    else Distances<LinkedAdapter>(*d, n, a, b, dist, dim_dist, dim_width);
    return 0;
}

/*
 * UncertaintyPlannerState(state_id, particles, ...)::UpdateStatistics(robot) (uncertainty_planner_state.hpp:339-351): expectation,
 * variance, space-independent variance and the two per-dimension variance vectors of one particle set (a child state's cluster)
 */
extern "C++" {
namespace
{
template <typename A, typename Serializer, typename Alloc>
void StateStatistics(const upc_robot_desc& d, int n, const double* configs, double step_size, int dim_width, double* expectation, double* scalars, double* variances,
                     double* si_variances)
{
    const int width = Width(d);
    const typename A::Robot robot = A::Make(d);
    std::vector<typename A::Config, Alloc> particles;
    for (int i = 0; i < n; i++) particles.push_back(A::In(d, configs + (size_t)i * width));
    uncertainty_planning_tools::UncertaintyPlannerState<typename A::Config, Serializer, Alloc> state(1u, particles, 1u, 1u, 1.0, 1u, 1u, 1.0, step_size,
                                                                                                  particles.front(), 0u, 0u, 0u);
    state.UpdateStatistics(robot);
    A::Out(state.GetExpectation(), expectation);
    scalars[0] = state.GetVariance();
    scalars[1] = state.GetSpaceIndependentVariance();
    const Eigen::VectorXd v = state.GetVariances(), sv = state.GetSpaceIndependentVariances();
    assert(v.size() == dim_width && sv.size() == dim_width);
    for (int k = 0; k < dim_width; k++)
    {
        variances[k] = v(k);
        si_variances[k] = sv(k);
    }
}
}
}

int ref_state_statistics(const upc_robot_desc* d, int n, const double* configs, double step_size, int dim_width, double* expectation, double* scalars,
                         double* variances, double* si_variances)
{
    if (d->kind == UPC_ROBOT_SE2)
        StateStatistics<Se2Adapter, simple_robot_models::EigenMatrixD31Serializer, std::allocator<Eigen::Matrix<double, 3, 1>>>(*d, n, configs, step_size, dim_width, expectation,
                                                                                                                              scalars, variances, si_variances);
    else if (d->kind == UPC_ROBOT_SE3)
        StateStatistics<Se3Adapter, simple_robot_models::EigenIsometry3dSerializer, Eigen::aligned_allocator<Eigen::Isometry3d>>(*d, n, configs, step_size, dim_width, expectation,
                                                                                                                               scalars, variances, si_variances);
    else
        StateStatistics<LinkedAdapter, simple_robot_models::SimpleLinkedConfigurationSerializer, std::allocator<simple_robot_models::SimpleLinkedConfiguration>>(
            *d, n, configs, step_size, dim_width, expectation, scalars, variances, si_variances);
    return 0;
}

/* AverageConfigurations of n configurations */
int ref_rob_average(const upc_robot_desc* d, int n, const double* configs, double* out)
{
    const int width = Width(*d);
    if (d->kind == UPC_ROBOT_SE2)
    {
        std::vector<Eigen::Matrix<double, 3, 1>> all;
        for (int i = 0; i < n; i++) all.push_back(Se2Config(configs + (size_t)i * width));
        Se2Adapter::Out(MakeSe2(*d).AverageConfigurations(all), out);
    }
    else if (d->kind == UPC_ROBOT_SE3)
    {
        EigenHelpers::VectorIsometry3d all;
        for (int i = 0; i < n; i++) all.push_back(FromRows12(configs + (size_t)i * width));
        ToRows12(MakeSe3(*d).AverageConfigurations(all), out);
    }
    else
    {
        std::vector<simple_robot_models::SimpleLinkedConfiguration> all;
        for (int i = 0; i < n; i++) all.push_back(MakeLinkedConfig(*d, configs + (size_t)i * width));
        LinkedAdapter::Out(MakeLinked(*d).AverageConfigurations(all), out);
    }
    return 0;
}
}
