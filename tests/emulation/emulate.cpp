/*
 * emulate.cpp - runs the product's per-particle device program (uncertainty_planning_core_b200/csrc/upc_sim.cuh,
 * one lane per particle, G = 1) on the CPU.  DEVELOPMENT / TEST AID ONLY: it lets the CPU test suite compare the
 * very code the CUDA kernels execute against the independent oracle in a container without a GPU.  It is built
 * into tests/emulation/libupc_emulation.so by tests/conftest.py and never loaded by the product package.
 */
#include <string.h>
#include <math.h>
#include <cmath>
#include <vector>
#include "upc_sim.cuh"
#include "upc_host_prep.h"

using namespace upc;

static void fill_env(const upc_env_desc* e, DevEnv* d)
{
    d->nx = e->nx; d->ny = e->ny; d->nz = e->nz;
    d->planar = 0;
    if (e->nz == 1)
    {
        d->planar = 1;
        for (int64_t k = 0; k < (int64_t)e->nx * e->ny; k++)
            if (!std::isfinite(e->sdf[k])) d->planar = 0;
    }
    d->sdf_oob = e->sdf_oob_value; d->occ_oob = e->occupancy_oob_value;
    d->res = e->resolution; d->inv_res = 1.0 / e->resolution;
    d->dnx = (double)e->nx; d->dny = (double)e->ny; d->dnz = (double)e->nz;
    for (int k = 0; k < 12; k++) d->G[k] = e->grid_from_world[k];
    d->sdf = e->sdf; d->seg = e->convex_segment; d->occ = e->occupancy; d->cellmin = nullptr; d->cells = nullptr;
}

/* same conversion as upc_set_robot in upc_capi.cu (validation omitted) */
static void fill_robot(const upc_robot_desc* r, DevRobot* d)
{
    memset(d, 0, sizeof(*d));
    d->kind = r->kind; d->dof = r->dof; d->num_links = r->num_links;
    d->num_joints = (r->kind == UPC_ROBOT_LINKED) ? r->num_joints : 0;
    d->num_points = r->num_points;
    d->sensor_noise_free = 1;
    for (int a = 0; a < r->dof; a++)
        if (r->axis[a].max_sensor_noise != 0.0) d->sensor_noise_free = 0;
    for (int l = 0; l <= r->num_links; l++) d->link_off[l] = r->link_point_offset[l];
    for (int l = 0; l < UPC_MAX_LINKS; l++) d->parent_joint_of_link[l] = -1;
    for (int a = 0; a < r->dof; a++)
    {
        d->axis[a].kp = fabs(r->axis[a].kp); d->axis[a].ki = fabs(r->axis[a].ki); d->axis[a].kd = fabs(r->axis[a].kd);
        d->axis[a].clamp = fabs(r->axis[a].integral_clamp); d->axis[a].vlim = fabs(r->axis[a].velocity_limit);
        d->axis[a].nb = fabs(r->axis[a].max_actuator_noise); d->weights[a] = fabs(r->distance_weights[a]);
        d->axis[a].sensor_scale = sensor_noise_scale(r->axis[a]); d->axis[a].actuator_scale = actuator_noise_scale(r->axis[a]);
    }
    compute_reach(r, d->reach);
    compute_link_flat_z(r, d->link_flat_z, d->link_z);
    for (int l = 0; l < r->num_links; l++)
    {
        d->link_radius[l] = 0.0;
        for (int k = r->link_point_offset[l]; k < r->link_point_offset[l + 1]; k++) d->link_radius[l] = fmax(d->link_radius[l], r->points[4 * k + 3]);
    }
    if (r->kind == UPC_ROBOT_LINKED)
    {
        for (int k = 0; k < 12; k++) d->base[k] = r->base_transform[k];
        int active = 0;
        for (int j = 0; j < r->num_joints; j++)
        {
            const upc_joint_desc& jd = r->joints[j];
            DevJoint& dj = d->joints[j];
            dj.type = jd.type; dj.parent = jd.parent_link; dj.child = jd.child_link; dj.active = -1;
            for (int k = 0; k < 3; k++) dj.axis[k] = jd.axis[k];
            for (int k = 0; k < 12; k++) dj.T[k] = jd.transform[k];
            if (jd.type == UPC_JOINT_CONTINUOUS) { dj.lo = -INFINITY; dj.hi = INFINITY; }
            else { dj.lo = jd.lower_limit; dj.hi = jd.upper_limit; }
            if (jd.type != UPC_JOINT_FIXED) dj.active = active++;
            d->parent_joint_of_link[jd.child_link] = j;
        }
    }
}

static int g_use_shared_masks = 1;
extern "C" void emulate_set_shared_masks(int on) { g_use_shared_masks = on; }

template <class M>
static void run(const DevEnv& env, const DevRobot& rob, const DevSim& sp, int64_t n, const double* starts, int64_t n_targets,
                const double* targets, const double* tape, double* out, uint8_t* coll, int64_t* counters)
{
    const bool seeded = (tape == nullptr);
    Group<1> g; g.mask = 1u; g.base = 0; g.sub = 0;
    std::vector<double> f0((size_t)rob.num_links * 12), f1((size_t)rob.num_links * 12), park((size_t)2 * M::WIDTH + 2 * M::DOF);
    for (int64_t i = 0; i < n; i++)
    {
        LocalCounters lc = {};
        bool c = false;
        unsigned masks[kMaxMaskChunks]; /* shared pre-cull masks of the collision check and the resolver pass behind it */
        unsigned* mp = g_use_shared_masks ? masks : nullptr;
        if (seeded) simulate_particle<M, 1, true>(env, rob, sp, g, i, n, starts, n_targets, targets, tape, out, coll, rob.pts, f0.data(), f1.data(), park.data(), lc, c, true, mp, 1);
        else simulate_particle<M, 1, false>(env, rob, sp, g, i, n, starts, n_targets, targets, tape, out, coll, rob.pts, f0.data(), f1.data(), park.data(), lc, c, true, mp, 1);
        counters[0] += lc.particle_steps; counters[1] += c ? 1 : 0; counters[2] += lc.resolver_iterations; counters[3] += lc.failed_resolves;
    }
}

static int g_use_cells = 1;
static int g_use_fp32_cull = 1;
static int g_use_link_cull = 1;
static int g_use_group_cull = 1;
extern "C" void emulate_set_corner_cells(int on) { g_use_cells = on; }
extern "C" void emulate_set_fp32_cull(int on) { g_use_fp32_cull = on; }
extern "C" void emulate_set_link_cull(int on) { g_use_link_cull = on; }
extern "C" void emulate_set_group_cull(int on) { g_use_group_cull = on; }

/* link boxes + dilated cell minima, as upc_set_robot prepares them */
static void fill_link_cull(const upc_env_desc* e, const upc_robot_desc* r, const std::vector<float>& cellmin, DevRobot* rob, std::vector<float>* dil)
{
    rob->dil_radius = g_use_link_cull ? compute_link_boxes(r, e->resolution, e->grid_from_world, rob->link_center, rob->link_half) : 0;
    rob->dilmin = nullptr;
    if (rob->dil_radius > 0)
    {
        dil->resize(cellmin.size());
        build_dilated_cellmin(cellmin.data(), e->nx, e->ny, e->nz, rob->dil_radius, dil->data());
        rob->dilmin = dil->data();
    }
}

/* point groups + the cell minima dilated by the largest group, as upc_set_robot prepares them */
static void fill_group_cull(const upc_env_desc* e, const upc_robot_desc* r, const std::vector<float>& cellmin, DevRobot* rob, std::vector<float>* dil,
                            std::vector<HostPointGroup>* groups)
{
    rob->grp_radius = g_use_group_cull ? compute_point_groups(r, e->resolution, e->grid_from_world, rob->link_chunk_off, groups) : 0;
    rob->grpmin = nullptr;
    rob->groups = nullptr;
    if (rob->grp_radius > 0)
    {
        dil->resize(cellmin.size());
        build_dilated_cellmin(cellmin.data(), e->nx, e->ny, e->nz, rob->grp_radius, dil->data());
        rob->grpmin = dil->data();
        rob->groups = reinterpret_cast<const DevPointGroup*>(groups->data());
    }
}

/*
 * Conservativeness of the group-level cull, tested directly: for n_frames link poses every point of a group the dilated lookup
 * declares clear must be clear for the exact FP64 program at `threshold`.  counts[0] = groups tested, [1] = groups cleared,
 * [2] = cleared groups holding a point the exact program flags (must be 0), [3] = points in cleared groups.
 */
extern "C" int emulate_group_cull_check(const upc_env_desc* e, const upc_robot_desc* r, int64_t n_frames, const double* frames, double threshold, int64_t* counts)
{
    DevEnv env; DevRobot rob;
    fill_env(e, &env); fill_robot(r, &rob);
    std::vector<float> cellmin((size_t)(e->nx + 1) * (size_t)(e->ny + 1) * (size_t)(e->nz + 1));
    build_cellmin(e->sdf, e->nx, e->ny, e->nz, cellmin.data());
    env.cellmin = cellmin.data();
    const std::vector<double> both = points_with_float_copy(r->points, r->num_points);
    rob.pts = both.data();
    std::vector<float> dil;
    std::vector<HostPointGroup> groups;
    const int saved = g_use_group_cull;
    g_use_group_cull = 1;
    fill_group_cull(e, r, cellmin, &rob, &dil, &groups);
    g_use_group_cull = saved;
    for (int k = 0; k < 4; k++) counts[k] = 0;
    if (rob.grp_radius == 0) return 1;
    for (int64_t f = 0; f < n_frames; f++)
    {
        double V[12];
        voxel_from_link(env, frames + 12 * f, V);
        for (int l = 0; l < rob.num_links; l++)
        {
            const float clear_above = float_round_up((rob.link_radius[l] + threshold) + UPC_CULL_MARGIN);
            for (int chunk = rob.link_off[l]; chunk < rob.link_off[l + 1]; chunk += 32)
            {
                const int chunk_index = (chunk - rob.link_off[l]) >> 5;
                const unsigned candidates = group_candidates(env, rob, l, chunk_index, V, clear_above);
                const DevPointGroup* grp = rob.groups + (int64_t)(rob.link_chunk_off[l] + chunk_index) * 4;
                for (int q = 0; q < 4 && grp[q].mask != 0u; q++)
                {
                    counts[0]++;
                    if ((candidates & grp[q].mask) != 0u) continue;
                    counts[1]++;
                    bool any = false;
                    for (int j = 0; j < 32; j++)
                    {
                        if (!((grp[q].mask >> j) & 1u)) continue;
                        const int k = chunk + j;
                        double u[3];
                        xf_point(V, rob.pts[4 * k + 0], rob.pts[4 * k + 1], rob.pts[4 * k + 2], u);
                        any |= (sdf_distance(env, u) - rob.pts[4 * k + 3]) < threshold;
                        counts[3]++;
                    }
                    counts[2] += any ? 1 : 0;
                }
            }
        }
    }
    return 0;
}

/*
 * Conservativeness of the single-precision pre-cull, tested directly: for n_frames link poses (rows of 12 doubles)
 * every body point the cull declares clear must be clear for the exact FP64 program at `threshold`.
 * counts[0] = points tested, [1] = culled, [2] = violations (must be 0), [3] = exact hits, [4] = links cleared by the link-level
 * cull, [5] = violations of the link-level cull (must be 0).
 */
extern "C" int emulate_cull_check(const upc_env_desc* e, const upc_robot_desc* r, int64_t n_frames, const double* frames, double threshold, int64_t* counts)
{
    DevEnv env; DevRobot rob;
    fill_env(e, &env); fill_robot(r, &rob);
    std::vector<float> cellmin((size_t)(e->nx + 1) * (size_t)(e->ny + 1) * (size_t)(e->nz + 1));
    build_cellmin(e->sdf, e->nx, e->ny, e->nz, cellmin.data());
    env.cellmin = cellmin.data();
    const std::vector<double> both = points_with_float_copy(r->points, r->num_points);
    rob.pts = both.data();
    rob.cull_delta = compute_cull_delta(e->nx, e->ny, e->nz, e->resolution, e->grid_from_world, r->points, r->num_points);
    std::vector<float> dil;
    {
        const int saved = g_use_link_cull; /* this check always looks at the link-level cull, whatever the simulation switch says */
        g_use_link_cull = 1;
        fill_link_cull(e, r, cellmin, &rob, &dil);
        g_use_link_cull = saved;
    }
    for (int k = 0; k < 6; k++) counts[k] = 0;
    if (!(rob.cull_delta > 0.0f)) return 1;
    const CullPoint* ptsf = reinterpret_cast<const CullPoint*>(rob.pts + 4 * rob.num_points);
    for (int64_t f = 0; f < n_frames; f++)
    {
        double V[12];
        voxel_from_link(env, frames + 12 * f, V);
        for (int l = 0; l < rob.num_links; l++)
        {
            const float clear_above = float_round_up((rob.link_radius[l] + threshold) + UPC_CULL_MARGIN);
            /* link-level cull: counts[4] = links declared clear, counts[5] = of those, links with a point the exact program hits (must be 0) */
            if (link_certainly_clear(env, rob, l, V, clear_above))
            {
                counts[4]++;
                bool any = false;
                for (int k = rob.link_off[l]; k < rob.link_off[l + 1]; k++)
                {
                    double u[3];
                    xf_point(V, rob.pts[4 * k + 0], rob.pts[4 * k + 1], rob.pts[4 * k + 2], u);
                    any |= (sdf_distance(env, u) - rob.pts[4 * k + 3]) < threshold;
                }
                counts[5] += any ? 1 : 0;
            }
            /* both forms of the point test: the packed planar one (taken only when the link's z is uniform) and the scalar 3-D one */
            const CullFrame cf = cull_frame<true>(env, rob, l, V, rob.cull_delta, clear_above);
            const CullFrame cf3 = cull_frame<false>(env, rob, l, V, rob.cull_delta, clear_above);
            for (int k = rob.link_off[l]; k < rob.link_off[l + 1]; k++)
            {
                double u[3];
                xf_point(V, rob.pts[4 * k + 0], rob.pts[4 * k + 1], rob.pts[4 * k + 2], u);
                const bool exact_hit = (sdf_distance(env, u) - rob.pts[4 * k + 3]) < threshold;
                const bool culled_planar = cull_certainly_clear<true>(env, cf, ptsf[k]);
                const bool culled_3d = cull_certainly_clear<false>(env, cf3, ptsf[k]);
                const bool culled = culled_planar || culled_3d;
                counts[0]++;
                counts[1] += culled ? 1 : 0;
                counts[2] += (culled && exact_hit) ? 1 : 0;
                counts[3] += exact_hit ? 1 : 0;
            }
        }
    }
    return 0;
}

static int emulate_impl(const upc_env_desc* e, const upc_robot_desc* r, const upc_sim_params* p, int64_t n, int64_t n_starts,
                        const double* starts, int64_t n_targets, const double* targets, const double* tape, uint64_t seed, int64_t first_particle,
                        double* out, uint8_t* coll, int64_t* counters, double* trace_configs, double* trace_controls, int32_t* trace_steps);

extern "C" int emulate_forward_simulate(const upc_env_desc* e, const upc_robot_desc* r, const upc_sim_params* p, int64_t n,
                                        const double* starts, int64_t n_targets, const double* targets, const double* tape,
                                        double* out, uint8_t* coll, int64_t* counters)
{
    return emulate_impl(e, r, p, n, n, starts, n_targets, targets, tape, 0, 0, out, coll, counters, nullptr, nullptr, nullptr);
}

/* seeded noise (tape == NULL inside), block-broadcast starts / targets, optional per-interval trace */
extern "C" int emulate_forward_simulate_seeded(const upc_env_desc* e, const upc_robot_desc* r, const upc_sim_params* p, uint64_t seed,
                                               int64_t first_particle, int64_t n, int64_t n_starts, const double* starts, int64_t n_targets,
                                               const double* targets, double* out, uint8_t* coll, int64_t* counters, double* trace_configs,
                                               double* trace_controls, int32_t* trace_steps)
{
    return emulate_impl(e, r, p, n, n_starts, starts, n_targets, targets, nullptr, seed, first_particle, out, coll, counters, trace_configs,
                        trace_controls, trace_steps);
}

static int emulate_impl(const upc_env_desc* e, const upc_robot_desc* r, const upc_sim_params* p, int64_t n, int64_t n_starts,
                        const double* starts, int64_t n_targets, const double* targets, const double* tape, uint64_t seed, int64_t first_particle,
                        double* out, uint8_t* coll, int64_t* counters, double* trace_configs, double* trace_controls, int32_t* trace_steps)
{
    DevEnv env; DevRobot rob; DevSim sp;
    fill_env(e, &env); fill_robot(r, &rob);
    const std::vector<double> both = points_with_float_copy(r->points, r->num_points);
    rob.pts = both.data();
    rob.cull_delta = g_use_fp32_cull ? compute_cull_delta(e->nx, e->ny, e->nz, e->resolution, e->grid_from_world, r->points, r->num_points) : 0.0f;
    std::vector<float> cellmin((size_t)(e->nx + 1) * (size_t)(e->ny + 1) * (size_t)(e->nz + 1));
    build_cellmin(e->sdf, e->nx, e->ny, e->nz, cellmin.data());
    env.cellmin = cellmin.data();
    std::vector<float> dil;
    fill_link_cull(e, r, cellmin, &rob, &dil);
    std::vector<float> group_dil;
    std::vector<HostPointGroup> groups;
    fill_group_cull(e, r, cellmin, &rob, &group_dil, &groups);
    std::vector<float> cells;
    if (g_use_cells)
    {
        cells.resize(cellmin.size() * 8 + 4);
        float* aligned = cells.data();
        while (((uintptr_t)aligned) & 15u) aligned++;
        build_corner_cells(e->sdf, e->nx, e->ny, e->nz, aligned);
        env.cells = aligned;
    }
    sp.num_steps = p->num_steps; sp.max_microsteps = p->max_microsteps; sp.allow_contacts = p->allow_contacts;
    sp.individual_jacobians = p->use_individual_jacobians; sp.max_resolver_iterations = p->max_resolver_iterations;
    sp.decay_iterations = p->resolver_decay_iterations; sp.dt = p->controller_interval; sp.shortcut = p->simulation_shortcut_distance;
    sp.contact_distance = p->contact_distance; sp.resolve_distance = p->resolve_distance; sp.microstep_distance = p->microstep_distance;
    sp.initial_scale = p->resolver_initial_scale; sp.decay_rate = p->resolver_decay_rate; sp.damping = p->resolver_damping;
    sp.step_begin = 0; sp.step_end = p->num_steps; sp.carry_pid = nullptr; sp.carry_flags = nullptr;
    sp.n_starts = n_starts; sp.per_start = n / n_starts; sp.per_target = n / n_targets;
    sp.seed = seed; sp.first_particle = first_particle;
    sp.trace_configs = trace_configs; sp.trace_controls = trace_controls; sp.trace_steps = trace_steps;
    for (int k = 0; k < 4; k++) counters[k] = 0;
    if (r->kind == UPC_ROBOT_SE2) run<Se2Model>(env, rob, sp, n, starts, n_targets, targets, tape, out, coll, counters);
    else if (r->kind == UPC_ROBOT_SE3) run<Se3Model>(env, rob, sp, n, starts, n_targets, targets, tape, out, coll, counters);
    else if (rob.dof <= LinkedModelSmall::DOF && rob.num_links <= LinkedModelSmall::LINKS) run<LinkedModelSmall>(env, rob, sp, n, starts, n_targets, targets, tape, out, coll, counters); /* same dispatch as launch_forward_simulate */
    else run<LinkedModel>(env, rob, sp, n, starts, n_targets, targets, tape, out, coll, counters);
    return 0;
}
