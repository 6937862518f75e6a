/*
 * upc_capi.cu - the C ABI declared in include/upc_b200.h: handle life cycle, uploads, the host- and
 * device-pointer entry points, statistics and the host-side noise-tape generator.
 */
#include "upc_internal.h"
#include "upc_host_prep.h"
#include "upc_noise.cuh"

#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <memory>
#include <random>

namespace upc
{
static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

cudaError_t DeviceBuffer::reserve(size_t want)
{
    if (want <= bytes && ptr) return cudaSuccess;
    if (ptr)
    {
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    if (want == 0) want = 256;
    /* round up so that slowly growing requests do not reallocate every call */
    size_t cap = 256;
    while (cap < want) cap *= 2;
    if (cap > want + (size_t(1) << 30)) cap = want;
    cudaError_t e = cudaMalloc(&ptr, cap);
    if (e != cudaSuccess)
    {
        ptr = nullptr;
        return e;
    }
    bytes = cap;
    return cudaSuccess;
}

void DeviceBuffer::release()
{
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
}

static bool finite_all(const double* v, int n)
{
    for (int i = 0; i < n; i++)
        if (!isfinite(v[i])) return false;
    return true;
}

static int32_t fail_arg(const std::string& msg)
{
    set_error(msg);
    return UPC_ERR_INVALID_ARGUMENT;
}

static int32_t to_dev_sim(const upc_sim_params* p, DevSim* out)
{
    if (!p) return fail_arg("null sim params");
    if (p->num_steps < 1) return fail_arg("num_steps must be >= 1");
    if (p->max_microsteps < 1) return fail_arg("max_microsteps must be >= 1");
    if (!(p->controller_interval > 0.0)) return fail_arg("controller_interval must be > 0");
    if (!(p->microstep_distance > 0.0)) return fail_arg("microstep_distance must be > 0");
    if (p->max_resolver_iterations < 1 || p->resolver_decay_iterations < 1) return fail_arg("resolver iteration counts must be >= 1");
    if (!(p->resolve_distance >= p->contact_distance)) return fail_arg("resolve_distance must be >= contact_distance");
    if (!(p->resolver_damping > 0.0)) return fail_arg("resolver_damping must be > 0");
    out->num_steps = p->num_steps;
    out->max_microsteps = p->max_microsteps;
    out->allow_contacts = p->allow_contacts;
    out->individual_jacobians = p->use_individual_jacobians;
    out->max_resolver_iterations = p->max_resolver_iterations;
    out->decay_iterations = p->resolver_decay_iterations;
    out->dt = p->controller_interval;
    out->shortcut = p->simulation_shortcut_distance;
    out->contact_distance = p->contact_distance;
    out->resolve_distance = p->resolve_distance;
    out->microstep_distance = p->microstep_distance;
    out->initial_scale = p->resolver_initial_scale;
    out->decay_rate = p->resolver_decay_rate;
    out->damping = p->resolver_damping;
    out->step_begin = 0;
    out->step_end = p->num_steps;
    out->carry_pid = nullptr;
    out->carry_flags = nullptr;
    out->n_starts = 0; /* filled by set_broadcast */
    out->per_start = 1;
    out->per_target = 1;
    out->seed = 0;
    out->first_particle = 0;
    out->trace_configs = nullptr;
    out->trace_controls = nullptr;
    out->trace_steps = nullptr;
    return UPC_OK;
}

/* starts / targets given per particle, per block of particles (edge batches) or once for all: n_x must divide n */
static int32_t set_broadcast(DevSim* sp, int64_t n, int64_t n_starts, int64_t n_targets)
{
    if (n_starts < 1 || (n % n_starts) != 0) return fail_arg("start_positions must have size N, or a size that divides N (one start per block of particles)");
    if (n_targets < 1 || (n % n_targets) != 0)
        return fail_arg("target_positions must have size 1 or size start_positions.size() (or a size that divides N: one target per block of particles)");
    sp->n_starts = n_starts;
    sp->per_start = n / n_starts;
    sp->per_target = n / n_targets;
    return UPC_OK;
}

/* fold the device-side counters into the host-side totals */
static int32_t pull_counters(upc_handle* h)
{
    DevCounters c;
    UPC_CUDA_TRY(cudaMemcpyAsync(&c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(c), h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.particle_steps += (int64_t)c.particle_steps;
    h->stats.collided_particles += (int64_t)c.collided;
    h->stats.resolver_iterations += (int64_t)c.resolver_iterations;
    h->stats.failed_resolves += (int64_t)c.failed_resolves;
    for (int k = 0; k < 12; k++) h->phase_cycles[k] += (int64_t)c.phase[k];
    return UPC_OK;
}
} /* namespace upc */

using namespace upc;

extern "C" {

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

const char* upc_last_error(void) { return g_last_error.c_str(); }
const char* upc_version(void) { return "upc_b200 0.1 (sm_100a)"; }

int32_t upc_create(const upc_env_desc* env, int32_t device, upc_handle** out)
{
    if (!env || !out) return fail_arg("null argument");
    if (env->nx < 1 || env->ny < 1 || env->nz < 1) return fail_arg("grid dimensions must be >= 1");
    if (!(env->resolution > 0.0) || !isfinite(env->resolution)) return fail_arg("resolution must be > 0");
    if (!env->sdf) return fail_arg("sdf grid is required");
    if ((int64_t)env->nx * (int64_t)env->ny * (int64_t)env->nz >= ((int64_t)1 << 31)) return fail_arg("grids are limited to 2^31 cells");
    if (!finite_all(env->grid_from_world, 12)) return fail_arg("grid_from_world must be finite");
    int count = 0;
    UPC_CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail_arg("no such CUDA device");
    UPC_CUDA_TRY(cudaSetDevice(device));
    /* owned until the end: every early return below releases the stream and the allocations made so far */
    std::unique_ptr<upc_handle, int32_t (*)(upc_handle*)> owner(new upc_handle(), upc_destroy);
    upc_handle* h = owner.get();
    h->device = device;
    cudaDeviceProp prop;
    UPC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    UPC_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t cells = (size_t)env->nx * (size_t)env->ny * (size_t)env->nz;
    UPC_CUDA_TRY(cudaMalloc(&h->d_sdf, cells * sizeof(float)));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->d_sdf, env->sdf, cells * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (int64_t)(cells * sizeof(float));
    if (env->convex_segment)
    {
        UPC_CUDA_TRY(cudaMalloc(&h->d_seg, cells * sizeof(uint32_t)));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->d_seg, env->convex_segment, cells * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (int64_t)(cells * sizeof(uint32_t));
    }
    if (env->occupancy)
    {
        UPC_CUDA_TRY(cudaMalloc(&h->d_occ, cells * sizeof(float)));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->d_occ, env->occupancy, cells * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (int64_t)(cells * sizeof(float));
    }
    {
        /* cell-minimum grid for the conservative SDF cull */
        const size_t mcells = (size_t)(env->nx + 1) * (size_t)(env->ny + 1) * (size_t)(env->nz + 1);
        std::vector<float>& cellmin = h->host_cellmin; /* kept: upc_set_robot dilates it by the robot's link size (link-level cull) */
        cellmin.resize(mcells);
        build_cellmin(env->sdf, env->nx, env->ny, env->nz, cellmin.data());
        UPC_CUDA_TRY(cudaMalloc(&h->d_cellmin, mcells * sizeof(float)));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->d_cellmin, cellmin.data(), mcells * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->stats.h2d_bytes += (int64_t)(mcells * sizeof(float));
    }
    {
        /* corner-cell table: 8x the SDF, kept only while it fits comfortably in the 126 MB L2 */
        const size_t mcells = (size_t)(env->nx + 1) * (size_t)(env->ny + 1) * (size_t)(env->nz + 1);
        const size_t bytes = mcells * 8 * sizeof(float);
        /* UPC_B200_NO_CORNER_CELLS=1 forces the plain-grid path (tests exercise both) */
        const char* off = getenv("UPC_B200_NO_CORNER_CELLS");
        if (bytes <= (size_t)80 * 1024 * 1024 && !(off && off[0] == '1'))
        {
            std::vector<float> cells(mcells * 8);
            build_corner_cells(env->sdf, env->nx, env->ny, env->nz, cells.data());
            UPC_CUDA_TRY(cudaMalloc(&h->d_cells, bytes));
            UPC_CUDA_TRY(cudaMemcpyAsync(h->d_cells, cells.data(), bytes, cudaMemcpyHostToDevice, h->stream));
            UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
            h->stats.h2d_bytes += (int64_t)bytes;
        }
    }
    UPC_CUDA_TRY(cudaMalloc(&h->d_counters, sizeof(DevCounters)));
    UPC_CUDA_TRY(cudaMemsetAsync(h->d_counters, 0, sizeof(DevCounters), h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    DevEnv& e = h->env;
    e.nx = env->nx; e.ny = env->ny; e.nz = env->nz;
    e.planar = 0;
    if (env->nz == 1)
    {
        e.planar = 1;
        for (size_t k = 0; k < cells; k++)
            if (!isfinite(env->sdf[k])) e.planar = 0;
    }
    e.sdf_oob = env->sdf_oob_value;
    e.occ_oob = env->occupancy_oob_value;
    e.res = env->resolution;
    e.inv_res = 1.0 / env->resolution;
    e.dnx = (double)env->nx; e.dny = (double)env->ny; e.dnz = (double)env->nz;
    for (int k = 0; k < 12; k++) e.G[k] = env->grid_from_world[k];
    e.sdf = h->d_sdf;
    e.seg = h->d_seg;
    e.occ = h->d_occ;
    e.cellmin = h->d_cellmin;
    e.cells = h->d_cells;
    *out = owner.release();
    return UPC_OK;
}

int32_t upc_destroy(upc_handle* h)
{
    if (!h) return UPC_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_sdf);
    cudaFree(h->d_seg);
    cudaFree(h->d_occ);
    cudaFree(h->d_cellmin);
    cudaFree(h->d_dilmin);
    cudaFree(h->d_grpmin);
    cudaFree(h->d_groups);
    cudaFree(h->d_cells);
    cudaFree(h->d_points);
    cudaFree(h->d_counters);
    DeviceBuffer* bufs[] = {&h->user[0], &h->user[1], &h->user[2], &h->user[3], &h->st_trace, &h->st_records, &h->st_order, &h->eb_starts[0], &h->eb_starts[1], &h->eb_targets[0], &h->eb_targets[1], &h->eb_out[0], &h->eb_out[1], &h->eb_coll[0], &h->eb_coll[1], &h->st_starts, &h->st_targets, &h->st_tape, &h->st_out, &h->st_coll, &h->st_labels, &h->st_ncl, &h->st_carry,
                            &h->cw.features, &h->cw.adjacency, &h->cw.roots, &h->cw.comp, &h->cw.flags, &h->cw.signatures,
                            &h->cw.frames, &h->cw.dist_matrix, &h->cw.members, &h->cw.link_scratch, &h->cw.misc, &h->cw.hash, &h->cw.dedupe, &h->cw.usig, &h->cw.uroot, &h->cw.comp_list, &h->cw.comp_feat};
    for (DeviceBuffer* b : bufs) b->release();
    if (h->aux_stream)
    {
        cudaStreamDestroy(h->aux_stream);
        for (int c = 0; c < 2; c++) cudaEventDestroy(h->eb_sim_done[c]);
    }
    if (h->copy_stream)
    {
        cudaStreamDestroy(h->copy_stream);
        for (int c = 0; c < 8; c++) cudaEventDestroy(h->copy_done[c]);
        cudaEventDestroy(h->compute_done);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return UPC_OK;
}

int32_t upc_set_robot(upc_handle* h, const upc_robot_desc* r)
{
    if (!h || !r) return fail_arg("null argument");
    if (r->kind != UPC_ROBOT_SE2 && r->kind != UPC_ROBOT_SE3 && r->kind != UPC_ROBOT_LINKED) return fail_arg("unknown robot kind");
    if (r->kind == UPC_ROBOT_SE2 && r->dof != 3) return fail_arg("SE2 robots have 3 degrees of freedom");
    if (r->kind == UPC_ROBOT_SE3 && r->dof != 6) return fail_arg("SE3 robots have 6 degrees of freedom");
    if (r->dof < 1 || r->dof > UPC_MAX_DOF) return fail_arg("dof out of range");
    if (r->num_links < 1 || r->num_links > UPC_MAX_LINKS) return fail_arg("num_links out of range");
    if (r->kind != UPC_ROBOT_LINKED && r->num_links != 1) return fail_arg("SE2/SE3 robots have exactly one link");
    if (r->num_points < 1 || !r->points) return fail_arg("robot needs body points");
    if (r->link_point_offset[0] != 0 || r->link_point_offset[r->num_links] != r->num_points) return fail_arg("link_point_offset must span all points");
    for (int l = 0; l < r->num_links; l++)
        if (r->link_point_offset[l + 1] < r->link_point_offset[l]) return fail_arg("link_point_offset must be non-decreasing");
    DevRobot d;
    memset(&d, 0, sizeof(d));
    d.kind = r->kind;
    d.dof = r->dof;
    d.num_links = r->num_links;
    d.num_joints = (r->kind == UPC_ROBOT_LINKED) ? r->num_joints : 0;
    d.num_points = r->num_points;
    d.sensor_noise_free = 1;
    for (int a = 0; a < r->dof; a++)
        if (r->axis[a].max_sensor_noise != 0.0) d.sensor_noise_free = 0;
    for (int l = 0; l <= r->num_links; l++) d.link_off[l] = r->link_point_offset[l];
    for (int l = 0; l < UPC_MAX_LINKS; l++) d.parent_joint_of_link[l] = -1;
    for (int a = 0; a < r->dof; a++)
    {
        d.axis[a].kp = fabs(r->axis[a].kp);
        d.axis[a].ki = fabs(r->axis[a].ki);
        d.axis[a].kd = fabs(r->axis[a].kd);
        d.axis[a].clamp = fabs(r->axis[a].integral_clamp);
        d.axis[a].vlim = fabs(r->axis[a].velocity_limit);
        d.axis[a].nb = fabs(r->axis[a].max_actuator_noise);
        d.axis[a].sensor_scale = sensor_noise_scale(r->axis[a]);
        d.axis[a].actuator_scale = actuator_noise_scale(r->axis[a]);
        d.weights[a] = fabs(r->distance_weights[a]);
    }
    if (r->kind == UPC_ROBOT_LINKED)
    {
        if (r->num_joints < 1 || r->num_joints > UPC_MAX_JOINTS) return fail_arg("num_joints out of range");
        for (int k = 0; k < 12; k++) d.base[k] = r->base_transform[k];
        int active = 0;
        std::vector<bool> placed((size_t)r->num_links, false);
        placed[0] = true;
        for (int j = 0; j < r->num_joints; j++)
        {
            const upc_joint_desc& jd = r->joints[j];
            if (jd.parent_link < 0 || jd.parent_link >= r->num_links || jd.child_link < 1 || jd.child_link >= r->num_links)
                return fail_arg("joint refers to a link that does not exist");
            /* UpdateTransforms (simple_robot_models.hpp:1299-1343) relies on parents being placed first */
            if (!placed[(size_t)jd.parent_link]) return fail_arg("joints must be ordered parents before children");
            if (placed[(size_t)jd.child_link]) return fail_arg("link has more than one parent joint");
            placed[(size_t)jd.child_link] = true;
            DevJoint& dj = d.joints[j];
            dj.type = jd.type;
            dj.parent = jd.parent_link;
            dj.child = jd.child_link;
            dj.active = -1;
            for (int k = 0; k < 3; k++) dj.axis[k] = jd.axis[k];
            for (int k = 0; k < 12; k++) dj.T[k] = jd.transform[k];
            if (jd.type == UPC_JOINT_CONTINUOUS)
            {
                dj.lo = -INFINITY; dj.hi = INFINITY;
            }
            else if (jd.type == UPC_JOINT_REVOLUTE || jd.type == UPC_JOINT_PRISMATIC)
            {
                if (!(jd.lower_limit <= jd.upper_limit)) return fail_arg("joint limits must satisfy lower <= upper");
                dj.lo = jd.lower_limit; dj.hi = jd.upper_limit;
            }
            else if (jd.type != UPC_JOINT_FIXED)
            {
                return fail_arg("unknown joint type");
            }
            if (jd.type != UPC_JOINT_FIXED)
            {
                d.axis_continuous[active] = (jd.type == UPC_JOINT_CONTINUOUS) ? 1 : 0;
                dj.active = active++;
            }
            d.parent_joint_of_link[jd.child_link] = j;
        }
        if (active != r->dof) return fail_arg("dof must equal the number of non-fixed joints");
        for (int l = 1; l < r->num_links; l++)
            if (!placed[(size_t)l]) return fail_arg("every link except the root needs a parent joint");
    }
    compute_reach(r, d.reach);
    compute_link_flat_z(r, d.link_flat_z, d.link_z);
    for (int l = 0; l < r->num_links; l++)
    {
        d.link_radius[l] = 0.0;
        for (int k = r->link_point_offset[l]; k < r->link_point_offset[l + 1]; k++) d.link_radius[l] = fmax(d.link_radius[l], r->points[4 * k + 3]);
    }
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    if (h->d_points) cudaFree(h->d_points);
    h->d_points = nullptr;
    {
        const std::vector<double> both = points_with_float_copy(r->points, r->num_points);
        UPC_CUDA_TRY(cudaMalloc(&h->d_points, both.size() * sizeof(double)));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->d_points, both.data(), both.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    {
        /* UPC_B200_NO_FP32_CULL=1 forces the FP64-only point passes (tests exercise both) */
        const char* off = getenv("UPC_B200_NO_FP32_CULL");
        d.cull_delta = (off && off[0] == '1') ? 0.0f : compute_cull_delta(h->env.nx, h->env.ny, h->env.nz, h->env.res, h->env.G, r->points, r->num_points);
    }
    d.pts = h->d_points;
    {
        /* link-level cull: bounding boxes of the links and the cell-minimum grid dilated by the largest link;
         * UPC_B200_NO_LINK_CULL=1 switches it off (tests exercise both) */
        const char* off = getenv("UPC_B200_NO_LINK_CULL");
        d.dil_radius = (off && off[0] == '1') ? 0 : compute_link_boxes(r, h->env.res, h->env.G, d.link_center, d.link_half);
        if (h->d_dilmin) cudaFree(h->d_dilmin);
        h->d_dilmin = nullptr;
        d.dilmin = nullptr;
        if (d.dil_radius > 0)
        {
            std::vector<float> dil(h->host_cellmin.size());
            build_dilated_cellmin(h->host_cellmin.data(), h->env.nx, h->env.ny, h->env.nz, d.dil_radius, dil.data());
            UPC_CUDA_TRY(cudaMalloc(&h->d_dilmin, dil.size() * sizeof(float)));
            UPC_CUDA_TRY(cudaMemcpyAsync(h->d_dilmin, dil.data(), dil.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
            h->stats.h2d_bytes += (int64_t)(dil.size() * sizeof(float));
            d.dilmin = h->d_dilmin;
        }
    }
    {
        /* group-level cull of the one-lane programs: point groups per 32-point chunk and the cell-minimum grid dilated by the largest
         * group; UPC_B200_NO_GROUP_CULL=1 switches it off (tests exercise both) */
        const char* off = getenv("UPC_B200_NO_GROUP_CULL");
        std::vector<HostPointGroup> groups;
        d.grp_radius = (off && off[0] == '1') ? 0 : compute_point_groups(r, h->env.res, h->env.G, d.link_chunk_off, &groups);
        if (d.dil_radius > 0 && d.grp_radius >= d.dil_radius) d.grp_radius = 0; /* no finer than the link-level test: nothing to gain */
        if (h->d_grpmin) cudaFree(h->d_grpmin);
        if (h->d_groups) cudaFree(h->d_groups);
        h->d_grpmin = nullptr;
        h->d_groups = nullptr;
        d.grpmin = nullptr;
        d.groups = nullptr;
        if (d.grp_radius > 0)
        {
            static_assert(sizeof(HostPointGroup) == sizeof(DevPointGroup), "group record layout");
            std::vector<float> dil(h->host_cellmin.size());
            build_dilated_cellmin(h->host_cellmin.data(), h->env.nx, h->env.ny, h->env.nz, d.grp_radius, dil.data());
            UPC_CUDA_TRY(cudaMalloc(&h->d_grpmin, dil.size() * sizeof(float)));
            UPC_CUDA_TRY(cudaMalloc(&h->d_groups, groups.size() * sizeof(HostPointGroup)));
            UPC_CUDA_TRY(cudaMemcpyAsync(h->d_grpmin, dil.data(), dil.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            UPC_CUDA_TRY(cudaMemcpyAsync(h->d_groups, groups.data(), groups.size() * sizeof(HostPointGroup), cudaMemcpyHostToDevice, h->stream));
            UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
            h->stats.h2d_bytes += (int64_t)(dil.size() * sizeof(float) + groups.size() * sizeof(HostPointGroup));
            d.grpmin = h->d_grpmin;
            d.groups = reinterpret_cast<const DevPointGroup*>(h->d_groups);
        }
    }
    h->robot = d;
    h->has_robot = true;
    return UPC_OK;
}

int32_t upc_get_stats(upc_handle* h, upc_stats* out)
{
    if (!h || !out) return fail_arg("null argument");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int32_t rc = pull_counters(h);
    if (rc != UPC_OK) return rc;
    *out = h->stats;
    return UPC_OK;
}

/* development aid (UPC_PHASE_TIMING builds): cycles per phase of the simulation program, summed over particles */
int32_t upc_debug_phase_cycles(upc_handle* h, int64_t* out8)
{
    if (!h || !out8) return fail_arg("null argument");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int32_t rc = pull_counters(h);
    if (rc != UPC_OK) return rc;
    for (int k = 0; k < 12; k++) { out8[k] = h->phase_cycles[k]; h->phase_cycles[k] = 0; }
    return UPC_OK;
}

int32_t upc_reset_stats(upc_handle* h)
{
    if (!h) return fail_arg("null argument");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int32_t rc = pull_counters(h);
    if (rc != UPC_OK) return rc;
    memset(&h->stats, 0, sizeof(h->stats));
    return UPC_OK;
}

int32_t upc_set_lanes_per_particle(upc_handle* h, int32_t lanes)
{
    if (!h) return fail_arg("null argument");
    if (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 3 && lanes != 4 && lanes != 5 && lanes != 6 && lanes != 8 && lanes != 16 && lanes != 32)
        return fail_arg("lanes must be 0, 1, 2, 3, 4, 5, 6, 8, 16 or 32");
    h->lanes_override = lanes;
    return UPC_OK;
}

int32_t upc_host_alloc(void** ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail_arg("bad argument");
    UPC_CUDA_TRY(cudaMallocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return UPC_OK;
}

int32_t upc_host_free(void* ptr)
{
    if (ptr) UPC_CUDA_TRY(cudaFreeHost(ptr));
    return UPC_OK;
}

/* device memory for hosts that do not link the CUDA runtime themselves (the C++ shim): four grow-only scratch buffers owned by
 * the handle, and copies ordered on the handle's stream (both copies return when the data has arrived) */
int32_t upc_device_scratch(upc_handle* h, int32_t slot, int64_t bytes, void** out)
{
    if (!h || !out || slot < 0 || slot >= 4 || bytes < 0) return fail_arg("bad argument");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    UPC_CUDA_TRY(h->user[slot].reserve((size_t)bytes));
    *out = h->user[slot].ptr;
    return UPC_OK;
}

int32_t upc_memcpy_to_device(upc_handle* h, void* d_dst, const void* src, int64_t bytes)
{
    if (!h || !d_dst || !src || bytes < 0) return fail_arg("bad argument");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    UPC_CUDA_TRY(cudaMemcpyAsync(d_dst, src, (size_t)bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.h2d_bytes += bytes;
    return UPC_OK;
}

int32_t upc_memcpy_to_host(upc_handle* h, void* dst, const void* d_src, int64_t bytes)
{
    if (!h || !dst || !d_src || bytes < 0) return fail_arg("bad argument");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    UPC_CUDA_TRY(cudaMemcpyAsync(dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += bytes;
    return UPC_OK;
}

int32_t upc_synchronize(upc_handle* h)
{
    if (!h) return fail_arg("null argument");
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return UPC_OK;
}

int32_t upc_forward_simulate_device(upc_handle* h, const upc_sim_params* params, int64_t n, const double* d_starts,
                                    int64_t n_targets, const double* d_targets, const double* d_tape,
                                    double* d_out_configs, uint8_t* d_out_collided, void* stream)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    DevSim sp;
    const int32_t rc = to_dev_sim(params, &sp);
    if (rc != UPC_OK) return rc;
    if (n < 0) return fail_arg("n must be >= 0");
    if (n == 0) return UPC_OK;
    if (!d_starts || !d_targets || !d_tape || !d_out_configs || !d_out_collided) return fail_arg("null buffer");
    {
        const int32_t brc = set_broadcast(&sp, n, n, n_targets);
        if (brc != UPC_OK) return brc;
    }
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    UPC_CUDA_TRY(launch_forward_simulate(h, sp, n, d_starts, n_targets, d_targets, d_tape, d_out_configs, d_out_collided, s));
    h->stats.kernel_launches++;
    h->stats.particles_simulated += n;
    return UPC_OK;
}

int32_t upc_forward_simulate_seeded_device(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n,
                                           int64_t n_starts, const double* d_starts, int64_t n_targets, const double* d_targets,
                                           double* d_out_configs, uint8_t* d_out_collided, void* stream)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    DevSim sp;
    const int32_t rc = to_dev_sim(params, &sp);
    if (rc != UPC_OK) return rc;
    if (n < 0 || first_particle < 0) return fail_arg("n and first_particle must be >= 0");
    if (n == 0) return UPC_OK;
    if (params->max_microsteps > 65534) return fail_arg("seeded noise supports at most 65534 microsteps per controller interval");
    if (!d_starts || !d_targets || !d_out_configs || !d_out_collided) return fail_arg("null buffer");
    {
        const int32_t brc = set_broadcast(&sp, n, n_starts, n_targets);
        if (brc != UPC_OK) return brc;
    }
    sp.seed = seed;
    sp.first_particle = first_particle;
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    UPC_CUDA_TRY(launch_forward_simulate(h, sp, n, d_starts, n_targets, d_targets, nullptr, d_out_configs, d_out_collided, s));
    h->stats.kernel_launches++;
    h->stats.particles_simulated += n;
    return UPC_OK;
}

/* shared by the host-pointer seeded entry points: uploads the (small) inputs, runs, leaves the results in st_out / st_coll */
static int32_t seeded_from_host(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n, int64_t n_starts,
                                const double* starts, int64_t n_targets, const double* targets)
{
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    const size_t start_bytes = (size_t)width * (size_t)n_starts * sizeof(double);
    const size_t tgt_bytes = (size_t)width * (size_t)n_targets * sizeof(double);
    UPC_CUDA_TRY(h->st_starts.reserve(start_bytes));
    UPC_CUDA_TRY(h->st_targets.reserve(tgt_bytes));
    UPC_CUDA_TRY(h->st_out.reserve((size_t)width * (size_t)n * sizeof(double)));
    UPC_CUDA_TRY(h->st_coll.reserve((size_t)n));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, starts, start_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_targets.ptr, targets, tgt_bytes, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (int64_t)(start_bytes + tgt_bytes);
    return upc_forward_simulate_seeded_device(h, params, seed, first_particle, n, n_starts, (const double*)h->st_starts.ptr, n_targets,
                                              (const double*)h->st_targets.ptr, (double*)h->st_out.ptr, (uint8_t*)h->st_coll.ptr, h->stream);
}

int32_t upc_forward_simulate_seeded(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n,
                                    int64_t n_starts, const double* starts, int64_t n_targets, const double* targets, double* out_configs,
                                    uint8_t* out_collided)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n < 0) return fail_arg("n must be >= 0");
    if (n == 0) return UPC_OK;
    if (n_starts < 1 || n_targets < 1) return fail_arg("start_positions / target_positions must not be empty");
    if (!starts || !targets || !out_configs || !out_collided) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int32_t rc = seeded_from_host(h, params, seed, first_particle, n, n_starts, starts, n_targets, targets);
    if (rc != UPC_OK) return rc;
    const size_t cfg_bytes = (size_t)upc_config_width(h->robot.kind, h->robot.dof) * (size_t)n * sizeof(double);
    UPC_CUDA_TRY(cudaMemcpyAsync(out_configs, h->st_out.ptr, cfg_bytes, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_collided, h->st_coll.ptr, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (int64_t)(cfg_bytes + (size_t)n);
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return UPC_OK;
}

/*
 * The planner-batch step in one call (BASELINE config 5; uncertainty_contact_planning.hpp:2039-2071 followed by :1986-2034 for
 * every candidate edge): n_edges x particles_per_edge particles simulated with seeded noise, then the outcome of every edge
 * clustered, with the outcomes staying in device memory between the two halves.
 */
int32_t upc_simulate_and_cluster_edges(upc_handle* h, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed,
                                       int64_t first_particle, int64_t n_edges, int64_t particles_per_edge, int64_t n_starts, const double* edge_starts,
                                       const double* edge_targets, double* out_configs, uint8_t* out_collided, int32_t* out_labels,
                                       int32_t* out_n_clusters)
{
    if (!h || !cparams) return fail_arg("null argument");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n_edges < 0 || particles_per_edge < 0) return fail_arg("negative size");
    if (n_edges == 0 || particles_per_edge == 0) return UPC_OK;
    if (n_edges > 65535) return fail_arg("at most 65535 edges per call");
    if (particles_per_edge > (int64_t)1 << 20) return fail_arg("at most 2^20 particles per edge");
    if (!edge_starts || !edge_targets || !out_configs || !out_collided || !out_labels || !out_n_clusters) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int64_t n = n_edges * particles_per_edge;
    if (n_starts != n_edges && n_starts != n) return fail_arg("starts must hold one configuration per edge or one per particle");
    int32_t rc = seeded_from_host(h, params, seed, first_particle, n, n_starts, edge_starts, n_edges, edge_targets);
    if (rc != UPC_OK) return rc;
    /* the configurations and collision flags (five sixths of what goes back to the host) are final once the simulation has run:
     * they travel on the copy stream while the outcome is clustered */
    if (!h->copy_stream)
    {
        UPC_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int c = 0; c < 8; c++) UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->copy_done[c], cudaEventDisableTiming));
        UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->compute_done, cudaEventDisableTiming));
    }
    const size_t cfg_bytes = (size_t)upc_config_width(h->robot.kind, h->robot.dof) * (size_t)n * sizeof(double);
    UPC_CUDA_TRY(cudaEventRecord(h->compute_done, h->stream));
    UPC_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->compute_done, 0));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_configs, h->st_out.ptr, cfg_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_collided, h->st_coll.ptr, (size_t)n, cudaMemcpyDeviceToHost, h->copy_stream));
    UPC_CUDA_TRY(cudaEventRecord(h->copy_done[0], h->copy_stream));
    UPC_CUDA_TRY(h->st_labels.reserve((size_t)n * sizeof(int32_t)));
    UPC_CUDA_TRY(h->st_ncl.reserve((size_t)n_edges * sizeof(int32_t) + 64));
    rc = upc_cluster_particles_device(h, cparams, n_edges, particles_per_edge, (const double*)h->st_out.ptr, (int32_t*)h->st_labels.ptr,
                                      (int32_t*)h->st_ncl.ptr, h->stream);
    if (rc != UPC_OK)
    {
        cudaStreamSynchronize(h->copy_stream); /* the caller's buffers must not be written after the call returns */
        return rc;
    }
    UPC_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->copy_done[0], 0));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_labels, h->st_labels.ptr, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_n_clusters, h->st_ncl.ptr, (size_t)n_edges * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (int64_t)(cfg_bytes + (size_t)n * 5 + (size_t)n_edges * 4);
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return UPC_OK;
}

/*
 * The planner-batch step as a pair of calls, so that a caller can keep two batches in flight: upc_edge_batch_submit uploads the
 * edge end points of a batch and queues its simulation (asynchronous: returns as soon as the work is queued);
 * upc_edge_batch_collect clusters the batch submitted to `slot` on the handle's second, high-priority stream and brings its
 * results to the host (blocks until they are there).  submit(k + 1) before collect(k): the simulation of the next batch fills the
 * machine while this batch's clustering - small latency-bound launches with host decisions in between - and its 119 MB of
 * results run underneath.  Same results as upc_simulate_and_cluster_edges for the same arguments.
 */
static int32_t ensure_edge_batch_streams(upc_handle* h)
{
    if (!h->aux_stream)
    {
        int least = 0, greatest = 0;
        UPC_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        UPC_CUDA_TRY(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, greatest));
        for (int c = 0; c < 2; c++) UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->eb_sim_done[c], cudaEventDisableTiming));
    }
    if (!h->copy_stream)
    {
        UPC_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int c = 0; c < 8; c++) UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->copy_done[c], cudaEventDisableTiming));
        UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->compute_done, cudaEventDisableTiming));
    }
    return UPC_OK;
}

int32_t upc_edge_batch_submit(upc_handle* h, const upc_sim_params* params, uint64_t seed, int64_t first_particle, int64_t n_edges,
                              int64_t particles_per_edge, int64_t n_starts, const double* edge_starts, const double* edge_targets, int32_t slot)
{
    if (!h || !params) return fail_arg("null argument");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (slot < 0 || slot > 1) return fail_arg("slot must be 0 or 1");
    if (n_edges < 1 || particles_per_edge < 1) return fail_arg("n_edges and particles_per_edge must be >= 1");
    if (n_edges > 65535) return fail_arg("at most 65535 edges per call");
    if (particles_per_edge > (int64_t)1 << 20) return fail_arg("at most 2^20 particles per edge");
    if (!edge_starts || !edge_targets) return fail_arg("null buffer");
    const int64_t n = n_edges * particles_per_edge;
    if (n_starts != n_edges && n_starts != n) return fail_arg("starts must hold one configuration per edge or one per particle");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int32_t src = ensure_edge_batch_streams(h);
    if (src != UPC_OK) return src;
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    const size_t start_bytes = (size_t)width * (size_t)n_starts * sizeof(double);
    const size_t tgt_bytes = (size_t)width * (size_t)n_edges * sizeof(double);
    UPC_CUDA_TRY(h->eb_starts[slot].reserve(start_bytes));
    UPC_CUDA_TRY(h->eb_targets[slot].reserve(tgt_bytes));
    UPC_CUDA_TRY(h->eb_out[slot].reserve((size_t)width * (size_t)n * sizeof(double)));
    UPC_CUDA_TRY(h->eb_coll[slot].reserve((size_t)n));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->eb_starts[slot].ptr, edge_starts, start_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->eb_targets[slot].ptr, edge_targets, tgt_bytes, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (int64_t)(start_bytes + tgt_bytes);
    const int32_t rc = upc_forward_simulate_seeded_device(h, params, seed, first_particle, n, n_starts, (const double*)h->eb_starts[slot].ptr, n_edges,
                                                          (const double*)h->eb_targets[slot].ptr, (double*)h->eb_out[slot].ptr, (uint8_t*)h->eb_coll[slot].ptr,
                                                          h->stream);
    if (rc != UPC_OK) return rc;
    UPC_CUDA_TRY(cudaEventRecord(h->eb_sim_done[slot], h->stream));
    h->eb_edges[slot] = n_edges;
    h->eb_particles[slot] = particles_per_edge;
    return UPC_OK;
}

int32_t upc_edge_batch_collect(upc_handle* h, const upc_cluster_params* cparams, int32_t slot, double* out_configs, uint8_t* out_collided,
                               int32_t* out_labels, int32_t* out_n_clusters)
{
    if (!h || !cparams) return fail_arg("null argument");
    if (slot < 0 || slot > 1) return fail_arg("slot must be 0 or 1");
    if (h->eb_edges[slot] < 1) return fail_arg("nothing was submitted to this slot");
    if (!out_configs || !out_collided || !out_labels || !out_n_clusters) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int64_t n_edges = h->eb_edges[slot], particles_per_edge = h->eb_particles[slot], n = n_edges * particles_per_edge;
    h->eb_edges[slot] = 0;
    const size_t cfg_bytes = (size_t)upc_config_width(h->robot.kind, h->robot.dof) * (size_t)n * sizeof(double);
    /* configurations and flags are final once the simulation has run: they travel on the copy stream while the outcome is clustered */
    UPC_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->eb_sim_done[slot], 0));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_configs, h->eb_out[slot].ptr, cfg_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_collided, h->eb_coll[slot].ptr, (size_t)n, cudaMemcpyDeviceToHost, h->copy_stream));
    UPC_CUDA_TRY(cudaStreamWaitEvent(h->aux_stream, h->eb_sim_done[slot], 0));
    UPC_CUDA_TRY(h->st_labels.reserve((size_t)n * sizeof(int32_t)));
    UPC_CUDA_TRY(h->st_ncl.reserve((size_t)n_edges * sizeof(int32_t) + 64));
    const int32_t rc = upc_cluster_particles_device(h, cparams, n_edges, particles_per_edge, (const double*)h->eb_out[slot].ptr, (int32_t*)h->st_labels.ptr,
                                                    (int32_t*)h->st_ncl.ptr, h->aux_stream);
    if (rc != UPC_OK)
    {
        cudaStreamSynchronize(h->copy_stream); /* the caller's buffers must not be written after the call returns */
        return rc;
    }
    UPC_CUDA_TRY(cudaMemcpyAsync(out_labels, h->st_labels.ptr, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->aux_stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_n_clusters, h->st_ncl.ptr, (size_t)n_edges * sizeof(int32_t), cudaMemcpyDeviceToHost, h->aux_stream));
    h->stats.d2h_bytes += (int64_t)(cfg_bytes + (size_t)n * 5 + (size_t)n_edges * 4);
    UPC_CUDA_TRY(cudaStreamSynchronize(h->aux_stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    return UPC_OK;
}

int32_t upc_forward_simulate(upc_handle* h, const upc_sim_params* params, int64_t n, const double* starts,
                             int64_t n_targets, const double* targets, const double* tape, double* out_configs,
                             uint8_t* out_collided)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n < 0) return fail_arg("n must be >= 0");
    if (n == 0) return UPC_OK;
    if (!params) return fail_arg("null sim params");
    if (n_targets < 1 || (n % n_targets) != 0)
        return fail_arg("target_positions must have size 1 or size start_positions.size() (or a size that divides N: one target per block of particles)");
    if (!starts || !targets || !tape || !out_configs || !out_collided) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    const size_t cfg_bytes = (size_t)width * (size_t)n * sizeof(double);
    const size_t tgt_bytes = (size_t)width * (size_t)n_targets * sizeof(double);
    if (params->num_steps < 1 || params->max_microsteps < 1) return fail_arg("bad step counts");
    const size_t tape_bytes = (size_t)upc_tape_doubles_per_particle(params, h->robot.dof) * (size_t)n * sizeof(double);
    UPC_CUDA_TRY(h->st_starts.reserve(cfg_bytes));
    UPC_CUDA_TRY(h->st_targets.reserve(tgt_bytes));
    UPC_CUDA_TRY(h->st_tape.reserve(tape_bytes));
    UPC_CUDA_TRY(h->st_out.reserve(cfg_bytes));
    UPC_CUDA_TRY(h->st_coll.reserve((size_t)n));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, starts, cfg_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_targets.ptr, targets, tgt_bytes, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (int64_t)(cfg_bytes + tgt_bytes);
    DevSim sp;
    {
        int32_t prc = to_dev_sim(params, &sp);
        if (prc != UPC_OK) return prc;
        prc = set_broadcast(&sp, n, n, n_targets);
        if (prc != UPC_OK) return prc;
    }
    /* The tape dominates the upload (C2: 61 MB of 63 MB).  It is step-major, so the edge is simulated as several
     * launches over consecutive step ranges while the next slices are still in flight on a second stream.  Only the
     * first slice's upload is exposed, so the slices start small: 1/32, 1/32, 1/16, 1/8, 1/4, 1/2 of the steps (doubling
     * keeps up as long as uploading a step takes at most half the time of simulating it; a slower link only delays the
     * later launches).  When every sensor noise bound is zero the kernels do not read the sensor slot and it
     * is not uploaded either: a pitched copy moves the actuator slots only - half of the tape for max_microsteps = 1. */
    int chunks = 1;
    int bounds[8] = {0, params->num_steps, 0, 0, 0, 0, 0, 0};
    if (tape_bytes >= ((size_t)8 << 20) && params->num_steps >= 8)
    {
        int len = params->num_steps / 32;
        if (len < 1) len = 1;
        int at = len;
        chunks = 1;
        bounds[1] = at;
        while (at < params->num_steps && chunks < 6)
        {
            at = (params->num_steps - (at + len) < len) ? params->num_steps : at + len; /* a short tail joins the last slice */
            bounds[++chunks] = at;
            if (chunks >= 2) len *= 2;
        }
        bounds[chunks] = params->num_steps;
    }
    const size_t step_bytes = tape_bytes / (size_t)params->num_steps;
    const size_t slot_bytes = (size_t)h->robot.dof * (size_t)n * sizeof(double);
    const bool skip_sensor_slot = (h->robot.sensor_noise_free != 0);
    if (chunks > 1)
    {
        UPC_CUDA_TRY(h->st_carry.reserve((size_t)2 * h->robot.dof * (size_t)n * sizeof(double) + (size_t)2 * n + 64));
        sp.carry_pid = (double*)h->st_carry.ptr;
        sp.carry_flags = (uint8_t*)h->st_carry.ptr + (size_t)2 * h->robot.dof * (size_t)n * sizeof(double);
        if (!h->copy_stream)
        {
            UPC_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            for (int c = 0; c < 8; c++) UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->copy_done[c], cudaEventDisableTiming));
            UPC_CUDA_TRY(cudaEventCreateWithFlags(&h->compute_done, cudaEventDisableTiming));
        }
        /* the previous call's kernels may still be reading the staging tape */
        UPC_CUDA_TRY(cudaEventRecord(h->compute_done, h->stream));
        UPC_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->compute_done, 0));
    }
    for (int c = 0; c < chunks; c++)
    {
        const int s0 = bounds[c], s1 = bounds[c + 1];
        const size_t off = (size_t)s0 * step_bytes, len = (size_t)(s1 - s0) * step_bytes;
        cudaStream_t cs = (chunks > 1) ? h->copy_stream : h->stream;
        if (skip_sensor_slot)
        {
            UPC_CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)h->st_tape.ptr + off + slot_bytes, step_bytes, (const uint8_t*)tape + off + slot_bytes, step_bytes,
                                           step_bytes - slot_bytes, (size_t)(s1 - s0), cudaMemcpyHostToDevice, cs));
            h->stats.h2d_bytes += (int64_t)((step_bytes - slot_bytes) * (size_t)(s1 - s0));
        }
        else
        {
            UPC_CUDA_TRY(cudaMemcpyAsync((uint8_t*)h->st_tape.ptr + off, (const uint8_t*)tape + off, len, cudaMemcpyHostToDevice, cs));
            h->stats.h2d_bytes += (int64_t)len;
        }
        if (chunks > 1)
        {
            UPC_CUDA_TRY(cudaEventRecord(h->copy_done[c], h->copy_stream));
            UPC_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->copy_done[c], 0));
        }
        sp.step_begin = s0;
        sp.step_end = s1;
        UPC_CUDA_TRY(launch_forward_simulate(h, sp, n, (const double*)h->st_starts.ptr, n_targets, (const double*)h->st_targets.ptr,
                                             (const double*)h->st_tape.ptr, (double*)h->st_out.ptr, (uint8_t*)h->st_coll.ptr, h->stream));
        h->stats.kernel_launches++;
    }
    h->stats.particles_simulated += n;
    UPC_CUDA_TRY(cudaMemcpyAsync(out_configs, h->st_out.ptr, cfg_bytes, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_collided, h->st_coll.ptr, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (int64_t)(cfg_bytes + (size_t)n);
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return UPC_OK;
}

int32_t upc_check_config_collision(upc_handle* h, const double* config, double contact_distance, double inflation,
                                   int32_t* out_in_collision)
{
    if (!h || !config || !out_in_collision) return fail_arg("null argument");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    UPC_CUDA_TRY(h->st_starts.reserve((size_t)width * sizeof(double) + 64));
    UPC_CUDA_TRY(h->st_ncl.reserve(64));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, config, (size_t)width * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(launch_check_collision(h, (const double*)h->st_starts.ptr, contact_distance + inflation, (int*)h->st_ncl.ptr, h->stream));
    h->stats.kernel_launches++;
    int result = 0;
    UPC_CUDA_TRY(cudaMemcpyAsync(&result, h->st_ncl.ptr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out_in_collision = result;
    return UPC_OK;
}

int32_t upc_cluster_statistics(upc_handle* h, int64_t n, const double* configs, const int32_t* labels, const uint8_t* collided,
                               int32_t allow_contacts, double step_size, int32_t n_clusters, int64_t* counts, uint8_t* any_collided,
                               double* expectations, double* variance, double* si_variance, double* variances, double* si_variances)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n < 0 || n_clusters < 0) return fail_arg("negative size");
    if (n_clusters == 0) return UPC_OK;
    if (!configs || !labels || !collided || !counts || !any_collided || !expectations || !variance || !si_variance || !variances || !si_variances)
        return fail_arg("null buffer");
    if (!(step_size > 0.0)) return fail_arg("step_size must be > 0");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    const int vw = upc_variance_width(h->robot.kind, h->robot.dof);
    const size_t cfg_bytes = (size_t)width * (size_t)n * sizeof(double);
    UPC_CUDA_TRY(h->st_starts.reserve(cfg_bytes + 64));
    UPC_CUDA_TRY(h->st_labels.reserve((size_t)n * sizeof(int32_t) + 64));
    UPC_CUDA_TRY(h->st_coll.reserve((size_t)n + 64));
    /* outputs: counts i64 | expectations | variance | si_variance | variances | si_variances (doubles) | any_collided u8 */
    const size_t K = (size_t)n_clusters;
    const size_t out_doubles = K * (size_t)(width + 2 + 2 * vw);
    UPC_CUDA_TRY(h->st_out.reserve(K * 8 + out_doubles * 8 + K + 64));
    uint8_t* base = (uint8_t*)h->st_out.ptr;
    int64_t* d_counts = (int64_t*)base;
    double* d_exp = (double*)(base + K * 8);
    double* d_var = d_exp + K * width;
    double* d_si = d_var + K;
    double* d_vars = d_si + K;
    double* d_sivars = d_vars + K * vw;
    uint8_t* d_any = (uint8_t*)(d_sivars + K * vw);
    if (n > 0)
    {
        UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, configs, cfg_bytes, cudaMemcpyHostToDevice, h->stream));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->st_labels.ptr, labels, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->st_coll.ptr, collided, (size_t)n, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (int64_t)(cfg_bytes + (size_t)n * 5);
    }
    UPC_CUDA_TRY(launch_cluster_stats(h, n, (const double*)h->st_starts.ptr, (const int32_t*)h->st_labels.ptr, (const uint8_t*)h->st_coll.ptr,
                                      allow_contacts, step_size, n_clusters, d_counts, d_any, d_exp, d_var, d_si, d_vars, d_sivars, h->stream));
    h->stats.kernel_launches++;
    UPC_CUDA_TRY(cudaMemcpyAsync(counts, d_counts, K * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(expectations, d_exp, K * width * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(variance, d_var, K * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(si_variance, d_si, K * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(variances, d_vars, K * vw * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(si_variances, d_sivars, K * vw * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(any_collided, d_any, K, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += (int64_t)(K * 8 + out_doubles * 8 + K);
    return UPC_OK;
}

int32_t upc_nearest_neighbors(upc_handle* h, int64_t n_states, const double* expectations, const double* feasibility_weight,
                              const double* variance_weight, const uint8_t* use_for_nn, int64_t n_queries, const double* queries,
                              double step_size, int64_t* out_index, double* out_distance)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n_states < 0 || n_queries < 0) return fail_arg("negative size");
    if (n_queries == 0) return UPC_OK;
    if (n_queries > 2147483647) return fail_arg("at most 2^31-1 queries per call");
    if (!queries || !out_index) return fail_arg("null buffer");
    if (!(step_size > 0.0)) return fail_arg("step_size must be > 0");
    if (n_states == 0)
    {
        for (int64_t q = 0; q < n_queries; q++)
        {
            out_index[q] = -1;
            if (out_distance) out_distance[q] = INFINITY;
        }
        return UPC_OK;
    }
    if (!expectations || !feasibility_weight || !variance_weight) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const size_t width = (size_t)upc_config_width(h->robot.kind, h->robot.dof);
    const size_t ns = (size_t)n_states, nq = (size_t)n_queries;
    /* st_starts: expectations | fw | vw ; st_targets: queries ; st_coll: use ; st_out: index | distance */
    UPC_CUDA_TRY(h->st_starts.reserve((width + 2) * ns * sizeof(double)));
    UPC_CUDA_TRY(h->st_targets.reserve(width * nq * sizeof(double)));
    UPC_CUDA_TRY(h->st_coll.reserve(ns + 64));
    UPC_CUDA_TRY(h->st_out.reserve(nq * 16));
    double* d_exp = (double*)h->st_starts.ptr;
    double* d_fw = d_exp + width * ns;
    double* d_vw = d_fw + ns;
    UPC_CUDA_TRY(cudaMemcpyAsync(d_exp, expectations, width * ns * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(d_fw, feasibility_weight, ns * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(d_vw, variance_weight, ns * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_targets.ptr, queries, width * nq * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (use_for_nn) UPC_CUDA_TRY(cudaMemcpyAsync(h->st_coll.ptr, use_for_nn, ns, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (int64_t)((width + 2) * ns * 8 + width * nq * 8 + (use_for_nn ? ns : 0));
    long long* d_index = (long long*)h->st_out.ptr;
    double* d_dist = (double*)((uint8_t*)h->st_out.ptr + nq * 8);
    UPC_CUDA_TRY(launch_nearest_states(h, n_states, d_exp, d_fw, d_vw, use_for_nn ? (const uint8_t*)h->st_coll.ptr : nullptr, n_queries,
                                       (const double*)h->st_targets.ptr, step_size, d_index, d_dist, h->stream));
    h->stats.kernel_launches++;
    static_assert(sizeof(long long) == sizeof(int64_t), "index width");
    UPC_CUDA_TRY(cudaMemcpyAsync(out_index, d_index, nq * 8, cudaMemcpyDeviceToHost, h->stream));
    if (out_distance) UPC_CUDA_TRY(cudaMemcpyAsync(out_distance, d_dist, nq * 8, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += (int64_t)(nq * (out_distance ? 16 : 8));
    return UPC_OK;
}

int32_t upc_count_within(upc_handle* h, int64_t n, const double* configs, int64_t n_targets, const double* targets, double threshold,
                         int64_t* out_count, uint8_t* within)
{
    if (!h || !out_count) return fail_arg("null argument");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n < 0) return fail_arg("n must be >= 0");
    *out_count = 0;
    if (n == 0) return UPC_OK;
    if (!(n_targets == 1 || n_targets == n)) return fail_arg("targets must have size 1 or n");
    if (!configs || !targets) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    const size_t cfg_bytes = (size_t)width * (size_t)n * sizeof(double), tgt_bytes = (size_t)width * (size_t)n_targets * sizeof(double);
    UPC_CUDA_TRY(h->st_starts.reserve(cfg_bytes));
    UPC_CUDA_TRY(h->st_targets.reserve(tgt_bytes));
    UPC_CUDA_TRY(h->st_coll.reserve((size_t)n));
    UPC_CUDA_TRY(h->st_ncl.reserve(64));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, configs, cfg_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_targets.ptr, targets, tgt_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemsetAsync(h->st_ncl.ptr, 0, sizeof(unsigned long long), h->stream));
    h->stats.h2d_bytes += (int64_t)(cfg_bytes + tgt_bytes);
    UPC_CUDA_TRY(launch_count_within(h, n, (const double*)h->st_starts.ptr, n_targets, (const double*)h->st_targets.ptr, threshold,
                                     (unsigned long long*)h->st_ncl.ptr, within ? (uint8_t*)h->st_coll.ptr : nullptr, h->stream));
    h->stats.kernel_launches++;
    unsigned long long count = 0;
    UPC_CUDA_TRY(cudaMemcpyAsync(&count, h->st_ncl.ptr, sizeof(count), cudaMemcpyDeviceToHost, h->stream));
    if (within) UPC_CUDA_TRY(cudaMemcpyAsync(within, h->st_coll.ptr, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->stats.d2h_bytes += 8 + (within ? n : 0);
    *out_count = (int64_t)count;
    return UPC_OK;
}

int32_t upc_cluster_particles_device(upc_handle* h, const upc_cluster_params* params, int64_t n_sets, int64_t set_size,
                                     const double* d_configs, int32_t* d_labels, int32_t* d_n_clusters, void* stream)
{
    if (!h || !params) return fail_arg("null argument");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n_sets < 0 || set_size < 0) return fail_arg("negative size");
    if (n_sets > 65535) return fail_arg("at most 65535 particle sets per call");
    if (set_size > (int64_t)1 << 20) return fail_arg("at most 2^20 particles per set");
    if (n_sets == 0) return UPC_OK;
    if (!d_n_clusters) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    if (set_size == 0)
    {
        UPC_CUDA_TRY(cudaMemsetAsync(d_n_clusters, 0, (size_t)n_sets * sizeof(int32_t), s));
        return UPC_OK;
    }
    if (!d_configs || !d_labels) return fail_arg("null buffer");
    if (!isfinite(params->distance_clustering_threshold)) return fail_arg("distance_clustering_threshold must be finite");
    return run_cluster(h, *params, n_sets, set_size, d_configs, d_labels, d_n_clusters, s);
}

int32_t upc_cluster_particles(upc_handle* h, const upc_cluster_params* params, int64_t n_sets, int64_t set_size,
                              const double* configs, int32_t* labels, int32_t* n_clusters)
{
    if (!h || !params) return fail_arg("null argument");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    if (n_sets < 0 || set_size < 0) return fail_arg("negative size");
    if (n_sets > 65535) return fail_arg("at most 65535 particle sets per call");
    if (set_size > (int64_t)1 << 20) return fail_arg("at most 2^20 particles per set");
    if (n_sets == 0) return UPC_OK;
    if (!n_clusters) return fail_arg("null buffer");
    if (set_size == 0)
    {
        for (int64_t s = 0; s < n_sets; s++) n_clusters[s] = 0;
        return UPC_OK;
    }
    if (!configs || !labels) return fail_arg("null buffer");
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int64_t total = n_sets * set_size;
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    const size_t cfg_bytes = (size_t)width * (size_t)total * sizeof(double);
    UPC_CUDA_TRY(h->st_starts.reserve(cfg_bytes));
    UPC_CUDA_TRY(h->st_labels.reserve((size_t)total * sizeof(int32_t)));
    UPC_CUDA_TRY(h->st_ncl.reserve((size_t)n_sets * sizeof(int32_t) + 64));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, configs, cfg_bytes, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (int64_t)cfg_bytes;
    upc_cluster_params local = *params;
    if (params->feature_type == UPC_FEATURE_POINT_TO_POINT_MOVEMENT && params->ptpm_sim && params->ptpm_tape && set_size > 1)
    {
        /* the pair simulations' tape arrives in host memory here */
        const int64_t sims = set_size * set_size - set_size;
        if (params->ptpm_sim->num_steps < 1 || params->ptpm_sim->max_microsteps < 1) return fail_arg("bad step counts");
        const size_t tape_bytes = (size_t)upc_tape_doubles_per_particle(params->ptpm_sim, h->robot.dof) * (size_t)sims * sizeof(double);
        UPC_CUDA_TRY(h->st_tape.reserve(tape_bytes));
        UPC_CUDA_TRY(cudaMemcpyAsync(h->st_tape.ptr, params->ptpm_tape, tape_bytes, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (int64_t)tape_bytes;
        local.ptpm_tape = (const double*)h->st_tape.ptr;
    }
    const int32_t rc = upc_cluster_particles_device(h, &local, n_sets, set_size, (const double*)h->st_starts.ptr, (int32_t*)h->st_labels.ptr,
                                                    (int32_t*)h->st_ncl.ptr, h->stream);
    if (rc != UPC_OK) return rc;
    UPC_CUDA_TRY(cudaMemcpyAsync(labels, h->st_labels.ptr, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(n_clusters, h->st_ncl.ptr, (size_t)n_sets * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (int64_t)((size_t)total * sizeof(int32_t) + (size_t)n_sets * sizeof(int32_t));
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return UPC_OK;
}

/*
 * Noise tape: replaces the per-thread PRNG fan-out of
 * uncertainty_contact_planning.hpp:343-356 + the truncated-normal draws of simple_uncertainty_models.hpp:38-45, 77-88.
 * Particle i draws from its own std::mt19937_64 seeded with SplitMix64(seed + i), so the tape - and therefore
 * every result - is independent of thread counts and of how particles are sharded across GPUs.
 * Standard normals come from the Marsaglia polar method on 53-bit uniforms, rejected outside [-2, 2]
 * (the reference's distributions are all mean 0, bounds +-2 sigma): sensor draws are z * max_sensor_noise / 2,
 * actuator draws z / 2.  A zero noise bound yields exact zeros without consuming random numbers.
 */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline double truncated_standard_normal(std::mt19937_64& rng)
{
    while (true)
    {
        const double u = ((double)(rng() >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
        const double v = ((double)(rng() >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
        const double s = u * u + v * v;
        if (s >= 1.0 || s == 0.0) continue;
        const double z = u * sqrt(-2.0 * log(s) / s);
        if (z >= -2.0 && z <= 2.0) return z;
    }
}

int32_t upc_fill_noise_tape(uint64_t seed, const upc_robot_desc* robot, const upc_sim_params* params, int64_t first_particle,
                            int64_t n, int64_t tape_stride, double* tape)
{
    if (!robot || !params || !tape) return fail_arg("null argument");
    if (robot->dof < 1 || robot->dof > UPC_MAX_DOF) return fail_arg("dof out of range");
    if (params->num_steps < 1 || params->max_microsteps < 1) return fail_arg("bad step counts");
    if (tape_stride < n) return fail_arg("tape_stride must be >= n");
    const int dof = robot->dof;
    const int slots = 1 + params->max_microsteps;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        std::mt19937_64 rng(splitmix64(seed + (uint64_t)(first_particle + i)));
        for (int step = 0; step < params->num_steps; step++)
        {
            for (int slot = 0; slot < slots; slot++)
            {
                for (int d = 0; d < dof; d++)
                {
                    double value;
                    if (slot == 0)
                    {
                        const double bound = fabs(robot->axis[d].max_sensor_noise);
                        value = (bound > 0.0) ? truncated_standard_normal(rng) * (bound * 0.5) : 0.0;
                    }
                    else
                    {
                        const double bound = fabs(robot->axis[d].max_actuator_noise);
                        value = (bound > 0.0) ? truncated_standard_normal(rng) * 0.5 : 0.0;
                    }
                    tape[(((int64_t)step * slots + slot) * dof + d) * tape_stride + i] = value;
                }
            }
        }
    }
    return UPC_OK;
}

/* The draws of the seeded entry points written out as a tape: replaying it through upc_forward_simulate gives
 * the results of upc_forward_simulate_seeded bit for bit. */
int32_t upc_fill_noise_tape_seeded(uint64_t seed, const upc_robot_desc* robot, const upc_sim_params* params, int64_t first_particle,
                                   int64_t n, int64_t tape_stride, double* tape)
{
    if (!robot || !params || !tape) return fail_arg("null argument");
    if (robot->dof < 1 || robot->dof > UPC_MAX_DOF) return fail_arg("dof out of range");
    if (params->num_steps < 1 || params->max_microsteps < 1 || params->max_microsteps > 65534) return fail_arg("bad step counts");
    if (tape_stride < n || first_particle < 0) return fail_arg("tape_stride must be >= n and first_particle >= 0");
    const int dof = robot->dof;
    const int slots = 1 + params->max_microsteps;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        for (int step = 0; step < params->num_steps; step++)
            for (int slot = 0; slot < slots; slot++)
                for (int pair = 0; 2 * pair < dof; pair++)
                {
                    const int d0 = 2 * pair, d1 = 2 * pair + 1;
                    const double s0 = (slot == 0) ? sensor_noise_scale(robot->axis[d0]) : actuator_noise_scale(robot->axis[d0]);
                    const double s1 = (d1 < dof) ? ((slot == 0) ? sensor_noise_scale(robot->axis[d1]) : actuator_noise_scale(robot->axis[d1])) : 0.0;
                    double z0 = 0.0, z1 = 0.0;
                    if (s0 != 0.0 || s1 != 0.0) noise_pair(seed, (uint64_t)(first_particle + i), (uint32_t)step, (uint32_t)slot, (uint32_t)pair, &z0, &z1);
                    tape[(((int64_t)step * slots + slot) * dof + d0) * tape_stride + i] = (s0 != 0.0) ? (z0 * s0) : 0.0;
                    if (d1 < dof) tape[(((int64_t)step * slots + slot) * dof + d1) * tape_stride + i] = (s1 != 0.0) ? (z1 * s1) : 0.0;
                }
    }
    return UPC_OK;
}

/*
 * ForwardSimulateRobot with enable_tracing (simple_simulator_interface.hpp:621, trace types :40-63): a (small) batch simulated
 * with a record of every controller interval - the resolved configuration at its end (what ExtractTrajectoryFromTrace :66-86
 * reads), the control input and the per-microstep control step.  tape == NULL selects seeded noise.
 */
int32_t upc_forward_simulate_traced(upc_handle* h, const upc_sim_params* params, int64_t n, const double* starts, int64_t n_targets,
                                    const double* targets, const double* tape, uint64_t seed, int64_t first_particle, double* out_configs,
                                    uint8_t* out_collided, double* trace_configs, double* trace_controls, int32_t* trace_steps)
{
    if (!h) return fail_arg("null handle");
    if (!h->has_robot)
    {
        set_error("upc_set_robot has not been called");
        return UPC_ERR_NO_ROBOT;
    }
    DevSim sp;
    int32_t rc = to_dev_sim(params, &sp);
    if (rc != UPC_OK) return rc;
    if (n < 0 || first_particle < 0) return fail_arg("n and first_particle must be >= 0");
    if (n == 0) return UPC_OK;
    if (!starts || !targets || !out_configs || !out_collided || !trace_configs || !trace_controls || !trace_steps) return fail_arg("null buffer");
    rc = set_broadcast(&sp, n, n, n_targets);
    if (rc != UPC_OK) return rc;
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    const int width = upc_config_width(h->robot.kind, h->robot.dof), dof = h->robot.dof;
    const size_t S = (size_t)params->num_steps;
    const size_t cfg_bytes = (size_t)width * (size_t)n * sizeof(double), tgt_bytes = (size_t)width * (size_t)n_targets * sizeof(double);
    const size_t tcfg_bytes = S * cfg_bytes, tctl_bytes = S * 2 * (size_t)dof * (size_t)n * sizeof(double), tsteps_bytes = 2 * (size_t)n * sizeof(int32_t);
    const size_t tape_bytes = tape ? (size_t)upc_tape_doubles_per_particle(params, dof) * (size_t)n * sizeof(double) : 0;
    UPC_CUDA_TRY(h->st_starts.reserve(cfg_bytes));
    UPC_CUDA_TRY(h->st_targets.reserve(tgt_bytes));
    UPC_CUDA_TRY(h->st_out.reserve(cfg_bytes));
    UPC_CUDA_TRY(h->st_coll.reserve((size_t)n));
    UPC_CUDA_TRY(h->st_trace.reserve(tcfg_bytes + tctl_bytes + tsteps_bytes + 64));
    if (tape) UPC_CUDA_TRY(h->st_tape.reserve(tape_bytes));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_starts.ptr, starts, cfg_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h->st_targets.ptr, targets, tgt_bytes, cudaMemcpyHostToDevice, h->stream));
    if (tape) UPC_CUDA_TRY(cudaMemcpyAsync(h->st_tape.ptr, tape, tape_bytes, cudaMemcpyHostToDevice, h->stream));
    UPC_CUDA_TRY(cudaMemsetAsync(h->st_trace.ptr, 0, tcfg_bytes + tctl_bytes + tsteps_bytes, h->stream));
    h->stats.h2d_bytes += (int64_t)(cfg_bytes + tgt_bytes + tape_bytes);
    sp.seed = seed;
    sp.first_particle = first_particle;
    sp.trace_configs = (double*)h->st_trace.ptr;
    sp.trace_controls = (double*)((uint8_t*)h->st_trace.ptr + tcfg_bytes);
    sp.trace_steps = (int32_t*)((uint8_t*)h->st_trace.ptr + tcfg_bytes + tctl_bytes);
    UPC_CUDA_TRY(launch_forward_simulate(h, sp, n, (const double*)h->st_starts.ptr, n_targets, (const double*)h->st_targets.ptr,
                                         tape ? (const double*)h->st_tape.ptr : nullptr, (double*)h->st_out.ptr, (uint8_t*)h->st_coll.ptr, h->stream));
    h->stats.kernel_launches++;
    h->stats.particles_simulated += n;
    UPC_CUDA_TRY(cudaMemcpyAsync(out_configs, h->st_out.ptr, cfg_bytes, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(out_collided, h->st_coll.ptr, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(trace_configs, sp.trace_configs, tcfg_bytes, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(trace_controls, sp.trace_controls, tctl_bytes, cudaMemcpyDeviceToHost, h->stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(trace_steps, sp.trace_steps, tsteps_bytes, cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (int64_t)(cfg_bytes + (size_t)n + tcfg_bytes + tctl_bytes + tsteps_bytes);
    UPC_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return UPC_OK;
}

} /* extern "C" */
