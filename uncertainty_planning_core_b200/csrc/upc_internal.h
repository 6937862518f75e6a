/*
 * upc_internal.h - host-side state behind the C ABI (include/upc_b200.h) and the launcher prototypes.
 */
#ifndef UPC_INTERNAL_H
#define UPC_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "upc_device.cuh"

namespace upc
{
void set_error(const std::string& msg);

struct DeviceBuffer
{
    void* ptr = nullptr;
    size_t bytes = 0;
    /* grow-only scratch; contents are not preserved */
    cudaError_t reserve(size_t want);
    void release();
};

struct ClusterWorkspace
{
    DeviceBuffer features;     /* per-particle distance features */
    DeviceBuffer adjacency;    /* bit matrix rows */
    DeviceBuffer roots;        /* first-pass roots | final roots | rank scratch */
    DeviceBuffer comp;         /* connected-component labels (exact path) */
    DeviceBuffer flags;        /* per-row ok flags + per-set bad flags */
    DeviceBuffer signatures;   /* convex-region signatures, point-major */
    DeviceBuffer frames;       /* link frames for linked-robot signatures */
    DeviceBuffer dist_matrix;  /* FP64 sub-matrices of non-clique components */
    DeviceBuffer members;      /* member lists + offsets of those components */
    DeviceBuffer link_scratch; /* per-member state of the complete-link kernel */
    DeviceBuffer misc;
    DeviceBuffer hash;         /* per-particle signature hashes */
    DeviceBuffer dedupe;       /* hash table + representative bookkeeping */
    DeviceBuffer usig;         /* compact signatures of the distinct vectors */
    DeviceBuffer uroot;        /* roots of the distinct vectors */
    DeviceBuffer comp_list;    /* sets the pivot pass left unsettled: list, packed groups, packed roots */
    DeviceBuffer comp_feat;    /* their packed features */
};

} /* namespace upc */

struct upc_handle
{
    int device = 0;
    cudaStream_t stream = nullptr;
    upc::DevEnv env{};
    upc::DevRobot robot{};
    bool has_robot = false;
    int lanes_override = 0;
    int sm_count = 148;
    float* d_sdf = nullptr;
    uint32_t* d_seg = nullptr;
    float* d_occ = nullptr;
    float* d_cellmin = nullptr;
    float* d_dilmin = nullptr;          /* cell minima dilated by the robot's largest link (link-level cull) */
    float* d_grpmin = nullptr;          /* cell minima dilated by the robot's largest point group (group-level cull) */
    void* d_groups = nullptr;           /* point-group records (DevPointGroup), four per 32-point chunk */
    std::vector<float> host_cellmin;    /* host copy of the cell-minimum grid, dilated again whenever the robot changes */
    float* d_cells = nullptr;
    double* d_points = nullptr;
    upc::DevCounters* d_counters = nullptr; /* device-side running totals */
    upc_stats stats{};
    int64_t phase_cycles[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    /* staging for the host-pointer entry points */
    upc::DeviceBuffer st_starts, st_targets, st_tape, st_out, st_coll, st_labels, st_ncl, st_carry, st_trace, st_records, st_order;
    /* upc_edge_batch_submit / _collect: two batches in flight (device-side staging per slot), clustering on its own stream */
    upc::DeviceBuffer eb_starts[2], eb_targets[2], eb_out[2], eb_coll[2];
    cudaEvent_t eb_sim_done[2] = {nullptr, nullptr};
    int64_t eb_edges[2] = {0, 0}, eb_particles[2] = {0, 0};
    cudaStream_t aux_stream = nullptr; /* highest priority: its small CTAs take the slots retiring simulation CTAs free */
    upc::DeviceBuffer user[4]; /* upc_device_scratch */
    cudaStream_t copy_stream = nullptr; /* tape slices of the host-pointer entry point */
    cudaEvent_t copy_done[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t compute_done = nullptr;
    upc::ClusterWorkspace cw;
};

namespace upc
{
/* launchers; all asynchronous on `stream`, return cudaGetLastError() */
int choose_lanes(const upc_handle* h, int64_t n);
cudaError_t launch_forward_simulate(upc_handle* h, const DevSim& sp, int64_t n, const double* starts, int64_t n_targets,
                                    const double* targets, const double* tape, double* out_configs, uint8_t* out_collided,
                                    cudaStream_t stream);
cudaError_t launch_nearest_states(upc_handle* h, int64_t n_states, const double* d_expectations, const double* d_fw, const double* d_vw,
                                  const uint8_t* d_use, int64_t n_queries, const double* d_queries, double step_size, long long* d_index,
                                  double* d_distance, cudaStream_t stream);
cudaError_t launch_count_within(upc_handle* h, int64_t n, const double* d_configs, int64_t n_targets, const double* d_targets, double threshold,
                                unsigned long long* d_count, uint8_t* d_within, cudaStream_t stream);
cudaError_t launch_check_collision(upc_handle* h, const double* d_config, double threshold, int* d_out, cudaStream_t stream);
cudaError_t launch_cluster_stats(upc_handle* h, int64_t n, const double* cfg, const int32_t* labels, const uint8_t* collided, int allow_contacts,
                                 double step_size, int n_clusters, int64_t* counts, uint8_t* any_collided, double* expectations, double* variance,
                                 double* si_variance, double* variances, double* si_variances, cudaStream_t stream);
int32_t run_cluster(upc_handle* h, const upc_cluster_params& cp, int64_t n_sets, int64_t set_size, const double* d_configs,
                    int32_t* d_labels, int32_t* d_n_clusters, cudaStream_t stream);
}

#define UPC_CUDA_TRY(expr)                                                                    \
    do                                                                                        \
    {                                                                                         \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
        {                                                                                     \
            upc::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            return UPC_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#endif
