/*
 * upc_comm.cu - the multi-GPU half of the C ABI: edges of a planner batch sharded over the GPUs of one box, SDF and robot
 * tables replicated (every rank owns a complete upc_handle), and ONE NCCL all-gather of outcome records per batch
 * (BASELINE.json north_star).  The path has no other exchange: particles never interact
 * (SimulatorInterface::ForwardSimulateRobots contract, simple_simulator_interface.hpp:623) and clustering is per edge
 * (uncertainty_contact_planning.hpp:1986-2034), so there is nothing to fuse a collective into - every rank simulates and
 * clusters its contiguous block of edges straight into its slot of the gather buffer and the all-gather runs in place.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2): the single-GPU product does not need it, and inside a process that
 * already loaded a NCCL (torch) the same library instance is used.
 */
#include "upc_internal.h"

#include <dlfcn.h>
#include <string.h>
#include <mutex>
#include <thread>

namespace upc
{
typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id
{
    char internal[128];
};
enum { kNcclSuccess = 0, kNcclUint8 = 1 };

struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string error;
};

static NcclApi* nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names)
        {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib)
        {
            api.error = std::string("NCCL not found: ") + dlerror();
            return;
        }
#define UPC_NCCL_SYM(field, name)                                                    \
    *(void**)(&api.field) = dlsym(api.lib, name);                                    \
    if (!api.field) api.error = std::string("NCCL symbol missing: ") + name;
        UPC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        UPC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        UPC_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        UPC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        UPC_NCCL_SYM(AllGather, "ncclAllGather")
        UPC_NCCL_SYM(GroupStart, "ncclGroupStart")
        UPC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        UPC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        UPC_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef UPC_NCCL_SYM
    });
    return &api;
}

static int32_t nccl_ready()
{
    NcclApi* a = nccl();
    if (!a->error.empty())
    {
        set_error(a->error);
        return UPC_ERR_UNSUPPORTED;
    }
    return UPC_OK;
}

#define UPC_NCCL_TRY(expr)                                                                      \
    do                                                                                          \
    {                                                                                           \
        const int _r = (expr);                                                                  \
        if (_r != kNcclSuccess)                                                                 \
        {                                                                                       \
            upc::set_error(std::string(#expr) + ": " + upc::nccl()->GetErrorString(_r));        \
            return UPC_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)
} /* namespace upc */

struct upc_comm
{
    upc_handle* h = nullptr;
    upc::nccl_comm_t comm = nullptr;
    int world = 1, rank = 0;
    cudaStream_t stream = nullptr;   /* the exchange runs on its own stream, ordered after the compute stream by events */
    cudaEvent_t ready = nullptr, done = nullptr;
    bool exchanged = false;
};

using namespace upc;

static int32_t fail_comm(const char* msg)
{
    set_error(msg);
    return UPC_ERR_INVALID_ARGUMENT;
}

static int32_t finish_comm(upc_comm* c)
{
    UPC_CUDA_TRY(cudaSetDevice(c->h->device));
    /* highest priority: the exchange of one batch overlaps the simulation of the next, whose kernel holds every register of every
     * SM - NCCL's few CTAs must get the slots that free up as simulation CTAs retire instead of queueing behind the whole launch */
    int least = 0, greatest = 0;
    UPC_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    UPC_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, greatest));
    UPC_CUDA_TRY(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
    UPC_CUDA_TRY(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    return UPC_OK;
}

extern "C" {

void upc_shard_bounds(int64_t total, int32_t rank, int32_t world, int64_t* lo, int64_t* hi)
{
    if (world < 1) world = 1;
    const int64_t base = total / world, extra = total % world;
    const int64_t a = (int64_t)rank * base + (rank < extra ? rank : extra);
    if (lo) *lo = a;
    if (hi) *hi = a + base + (rank < extra ? 1 : 0);
}

int32_t upc_outcome_layout_for(int32_t width, int64_t particles, int64_t sets, upc_outcome_layout* out)
{
    if (!out || width < 1 || particles < 0 || sets < 0) return fail_comm("bad outcome layout arguments");
    auto pad = [](int64_t v) { return (v + 255) / 256 * 256; };
    out->particles = particles;
    out->sets = sets;
    out->labels_offset = 0;
    out->n_clusters_offset = pad(4 * particles);
    out->collided_offset = out->n_clusters_offset + pad(4 * sets);
    out->configs_offset = out->collided_offset + pad(particles);
    out->block_bytes = out->configs_offset + pad(8 * (int64_t)width * particles);
    return UPC_OK;
}

int32_t upc_comm_get_unique_id(uint8_t* id128)
{
    if (!id128) return fail_comm("null id buffer");
    const int32_t rc = nccl_ready();
    if (rc != UPC_OK) return rc;
    nccl_unique_id id;
    UPC_NCCL_TRY(nccl()->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return UPC_OK;
}

int32_t upc_comm_create(upc_handle* h, int32_t world, int32_t rank, const uint8_t* id128, upc_comm** out)
{
    if (!h || !out || !id128) return fail_comm("null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail_comm("rank must lie in [0, world)");
    const int32_t rc = nccl_ready();
    if (rc != UPC_OK) return rc;
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    upc_comm* c = new upc_comm();
    c->h = h; c->world = world; c->rank = rank;
    nccl_unique_id id;
    memcpy(id.internal, id128, 128);
    const int r = nccl()->CommInitRank(&c->comm, world, id, rank);
    if (r != kNcclSuccess)
    {
        set_error(std::string("ncclCommInitRank: ") + nccl()->GetErrorString(r));
        delete c;
        return UPC_ERR_CUDA;
    }
    const int32_t frc = finish_comm(c);
    if (frc != UPC_OK)
    {
        upc_comm_destroy(c);
        return frc;
    }
    *out = c;
    return UPC_OK;
}

int32_t upc_comm_create_all(int32_t n, upc_handle* const* handles, upc_comm** out)
{
    if (n < 1 || n > 64 || !handles || !out) return fail_comm("bad argument");
    const int32_t rc = nccl_ready();
    if (rc != UPC_OK) return rc;
    int devs[64];
    nccl_comm_t comms[64];
    for (int i = 0; i < n; i++)
    {
        if (!handles[i]) return fail_comm("null handle");
        devs[i] = handles[i]->device;
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) return fail_comm("one handle per device");
    }
    UPC_NCCL_TRY(nccl()->CommInitAll(comms, n, devs));
    for (int i = 0; i < n; i++)
    {
        upc_comm* c = new upc_comm();
        c->h = handles[i]; c->world = n; c->rank = i; c->comm = comms[i];
        out[i] = c;
    }
    for (int i = 0; i < n; i++)
    {
        const int32_t frc = finish_comm(out[i]);
        if (frc != UPC_OK) return frc;
    }
    return UPC_OK;
}

int32_t upc_comm_destroy(upc_comm* c)
{
    if (!c) return UPC_OK;
    cudaSetDevice(c->h->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) nccl()->CommDestroy(c->comm);
    if (c->ready) cudaEventDestroy(c->ready);
    if (c->done) cudaEventDestroy(c->done);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return UPC_OK;
}

int32_t upc_comm_world(const upc_comm* c) { return c ? c->world : 0; }
int32_t upc_comm_rank(const upc_comm* c) { return c ? c->rank : -1; }

int32_t upc_comm_nccl_version(int32_t* out)
{
    const int32_t rc = nccl_ready();
    if (rc != UPC_OK) return rc;
    int v = 0;
    UPC_NCCL_TRY(nccl()->GetVersion(&v));
    if (out) *out = v;
    return UPC_OK;
}

/* the exchange of n communicators driven by this thread (n = 1: one process per GPU; n > 1: one process, several GPUs - the
 * NCCL calls of the ranks must then sit in one group).  Each exchange is ordered after the work queued on its rank's compute
 * stream and runs on the communicator's own stream. */
static int32_t all_gather_blocks(int n, upc_comm* const* cs, void* const* d_blocks, int64_t block_bytes, void* const* streams)
{
    for (int i = 0; i < n; i++)
    {
        upc_comm* c = cs[i];
        UPC_CUDA_TRY(cudaSetDevice(c->h->device));
        cudaStream_t s = (streams && streams[i]) ? (cudaStream_t)streams[i] : c->h->stream;
        UPC_CUDA_TRY(cudaEventRecord(c->ready, s));
        UPC_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ready, 0));
        /* work queued on the compute stream from here on waits for the PREVIOUS exchange: a caller alternating between two
         * gather buffers can issue step k + 1 while the exchange of step k is in flight, and step k + 1 will not overwrite the
         * buffer the exchange of step k - 1 was still using */
        if (c->exchanged) UPC_CUDA_TRY(cudaStreamWaitEvent(s, c->done, 0));
    }
    if (cs[0]->world > 1)
    {
        if (n > 1) UPC_NCCL_TRY(nccl()->GroupStart());
        for (int i = 0; i < n; i++)
        {
            upc_comm* c = cs[i];
            const uint8_t* mine = (const uint8_t*)d_blocks[i] + (size_t)c->rank * (size_t)block_bytes;
            UPC_NCCL_TRY(nccl()->AllGather(mine, d_blocks[i], (size_t)block_bytes, kNcclUint8, c->comm, c->stream));
        }
        if (n > 1) UPC_NCCL_TRY(nccl()->GroupEnd());
    }
    for (int i = 0; i < n; i++)
    {
        UPC_CUDA_TRY(cudaSetDevice(cs[i]->h->device));
        UPC_CUDA_TRY(cudaEventRecord(cs[i]->done, cs[i]->stream));
        cs[i]->exchanged = true;
    }
    return UPC_OK;
}

/* in-place all-gather of the outcome blocks: rank r's block sits at d_blocks + r * block_bytes on every rank.  Ordered after
 * everything queued on `stream` (the compute stream; NULL = the handle's), runs on the communicator's own stream, and `stream`
 * does NOT wait for it: call upc_comm_wait (or upc_comm_synchronize) before reading other ranks' blocks. */
int32_t upc_comm_all_gather_outcomes(upc_comm* c, void* d_blocks, int64_t block_bytes, void* stream)
{
    if (!c || !d_blocks || block_bytes <= 0) return fail_comm("bad argument");
    return all_gather_blocks(1, &c, &d_blocks, block_bytes, &stream);
}

/* make `stream` (NULL = the handle's compute stream) wait for the last exchange */
int32_t upc_comm_wait(upc_comm* c, void* stream)
{
    if (!c) return fail_comm("null communicator");
    UPC_CUDA_TRY(cudaSetDevice(c->h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->h->stream;
    UPC_CUDA_TRY(cudaStreamWaitEvent(s, c->done, 0));
    return UPC_OK;
}

int32_t upc_comm_synchronize(upc_comm* c)
{
    if (!c) return fail_comm("null communicator");
    UPC_CUDA_TRY(cudaSetDevice(c->h->device));
    UPC_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return UPC_OK;
}

/*
 * One sharded planner-batch step (BASELINE config 5): the batch's n_edges candidate edges are split into contiguous blocks
 * (upc_shard_bounds), this rank simulates its edges x particles_per_edge particles with seeded noise keyed by the GLOBAL
 * particle index (so results do not depend on the number of GPUs), clusters the outcome of each of its edges, and the
 * outcome blocks of all ranks are exchanged by one in-place all-gather.  d_edge_starts / d_edge_targets: [C][n_edges]
 * (replicated on every rank, device memory).  d_blocks: world * layout.block_bytes bytes with
 * layout = upc_outcome_layout_for(width, max shard particles, max shard edges).  Asynchronous; see upc_comm_all_gather_outcomes.
 */
static int32_t sharded_simulate(upc_comm* c, const upc_sim_params* params, uint64_t seed, int64_t n_edges, int64_t particles_per_edge,
                                const double* d_edge_starts, const double* d_edge_targets, void* d_blocks, void* stream, upc_outcome_layout* lay)
{
    if (!c || !params || !d_edge_starts || !d_edge_targets || !d_blocks) return fail_comm("null argument");
    if (n_edges < 1 || particles_per_edge < 1) return fail_comm("n_edges and particles_per_edge must be >= 1");
    upc_handle* h = c->h;
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    int64_t lo, hi;
    upc_shard_bounds(n_edges, c->rank, c->world, &lo, &hi);
    const int64_t cap_edges = (n_edges + c->world - 1) / c->world;
    upc_outcome_layout_for(width, cap_edges * particles_per_edge, cap_edges, lay);
    uint8_t* block = (uint8_t*)d_blocks + (size_t)c->rank * (size_t)lay->block_bytes;
    const int64_t e = hi - lo, n = e * particles_per_edge;
    if (e == 0) return UPC_OK;
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    /* this rank's columns [lo, hi) of the replicated edge arrays, compacted to [C][e] (tiny: e * C doubles) */
    UPC_CUDA_TRY(h->st_starts.reserve((size_t)width * (size_t)e * sizeof(double)));
    UPC_CUDA_TRY(h->st_targets.reserve((size_t)width * (size_t)e * sizeof(double)));
    UPC_CUDA_TRY(cudaMemcpy2DAsync(h->st_starts.ptr, (size_t)e * sizeof(double), d_edge_starts + lo, (size_t)n_edges * sizeof(double),
                                   (size_t)e * sizeof(double), (size_t)width, cudaMemcpyDeviceToDevice, s));
    UPC_CUDA_TRY(cudaMemcpy2DAsync(h->st_targets.ptr, (size_t)e * sizeof(double), d_edge_targets + lo, (size_t)n_edges * sizeof(double),
                                   (size_t)e * sizeof(double), (size_t)width, cudaMemcpyDeviceToDevice, s));
    return upc_forward_simulate_seeded_device(h, params, seed, lo * particles_per_edge, n, e, (const double*)h->st_starts.ptr, e,
                                              (const double*)h->st_targets.ptr, (double*)(block + lay->configs_offset),
                                              block + lay->collided_offset, s);
}

static int32_t sharded_cluster(upc_comm* c, const upc_cluster_params* cparams, int64_t n_edges, int64_t particles_per_edge, void* d_blocks, void* stream,
                               upc_outcome_layout* lay)
{
    if (!c || !cparams || !d_blocks) return fail_comm("null argument");
    if (n_edges < 1 || particles_per_edge < 1) return fail_comm("n_edges and particles_per_edge must be >= 1");
    upc_handle* h = c->h;
    const int width = upc_config_width(h->robot.kind, h->robot.dof);
    int64_t lo, hi;
    upc_shard_bounds(n_edges, c->rank, c->world, &lo, &hi);
    const int64_t cap_edges = (n_edges + c->world - 1) / c->world;
    upc_outcome_layout_for(width, cap_edges * particles_per_edge, cap_edges, lay);
    uint8_t* block = (uint8_t*)d_blocks + (size_t)c->rank * (size_t)lay->block_bytes;
    if (hi == lo) return UPC_OK;
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    return upc_cluster_particles_device(h, cparams, hi - lo, particles_per_edge, (const double*)(block + lay->configs_offset),
                                        (int32_t*)(block + lay->labels_offset), (int32_t*)(block + lay->n_clusters_offset), s);
}

static int32_t sharded_local(upc_comm* c, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed, int64_t n_edges,
                             int64_t particles_per_edge, const double* d_edge_starts, const double* d_edge_targets, void* d_blocks, void* stream,
                             upc_outcome_layout* lay)
{
    const int32_t rc = sharded_simulate(c, params, seed, n_edges, particles_per_edge, d_edge_starts, d_edge_targets, d_blocks, stream, lay);
    if (rc != UPC_OK) return rc;
    return sharded_cluster(c, cparams, n_edges, particles_per_edge, d_blocks, stream, lay);
}

int32_t upc_sharded_edge_batch(upc_comm* c, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed,
                               int64_t n_edges, int64_t particles_per_edge, const double* d_edge_starts, const double* d_edge_targets,
                               void* d_blocks, void* stream)
{
    upc_outcome_layout lay;
    const int32_t rc = sharded_local(c, params, cparams, seed, n_edges, particles_per_edge, d_edge_starts, d_edge_targets, d_blocks, stream, &lay);
    if (rc != UPC_OK) return rc;
    return all_gather_blocks(1, &c, &d_blocks, lay.block_bytes, &stream);
}

/*
 * The two halves of upc_sharded_edge_batch as separate calls, for callers that pipeline batches: the simulation of batch k + 1
 * (stream A, fills the machine) overlaps the clustering and the exchange of batch k (stream B: a handful of small, latency-bound
 * launches and the host decisions between them).  upc_sharded_edge_batch_simulate is asynchronous and never waits on the host;
 * upc_sharded_edge_batch_finish clusters this rank's block on `stream` - the caller orders it after the simulation of the same
 * batch with an event - and issues the all-gather behind it.  The two calls of one handle touch disjoint device state (simulation:
 * edge staging + counters; clustering: its workspace), so batch k + 1 may be simulated while batch k is finished, from ONE host
 * thread.  d_blocks of batches in flight must be distinct buffers.
 */
int32_t upc_sharded_edge_batch_simulate(upc_comm* c, const upc_sim_params* params, uint64_t seed, int64_t n_edges, int64_t particles_per_edge,
                                        const double* d_edge_starts, const double* d_edge_targets, void* d_blocks, void* stream)
{
    upc_outcome_layout lay;
    return sharded_simulate(c, params, seed, n_edges, particles_per_edge, d_edge_starts, d_edge_targets, d_blocks, stream, &lay);
}

int32_t upc_sharded_edge_batch_finish(upc_comm* c, const upc_cluster_params* cparams, int64_t n_edges, int64_t particles_per_edge, void* d_blocks, void* stream)
{
    upc_outcome_layout lay;
    const int32_t rc = sharded_cluster(c, cparams, n_edges, particles_per_edge, d_blocks, stream, &lay);
    if (rc != UPC_OK) return rc;
    return all_gather_blocks(1, &c, &d_blocks, lay.block_bytes, &stream);
}

/* the same step for all n ranks of a single-process communicator set (upc_comm_create_all): every array argument has one entry
 * per rank (device pointers on that rank's GPU); streams may be NULL */
int32_t upc_sharded_edge_batch_all(int32_t n, upc_comm* const* comms, const upc_sim_params* params, const upc_cluster_params* cparams, uint64_t seed,
                                   int64_t n_edges, int64_t particles_per_edge, const double* const* d_edge_starts,
                                   const double* const* d_edge_targets, void* const* d_blocks, void* const* streams)
{
    if (n < 1 || n > 64 || !comms || !d_edge_starts || !d_edge_targets || !d_blocks) return fail_comm("bad argument");
    /* one host thread per rank: the clustering half reads small decision flags back from its device, which would otherwise
     * serialise the ranks behind one another */
    upc_outcome_layout lays[64];
    int32_t codes[64];
    std::string messages[64];
    std::vector<std::thread> workers;
    for (int i = 0; i < n; i++)
    {
        workers.emplace_back([&, i]() {
            codes[i] = sharded_local(comms[i], params, cparams, seed, n_edges, particles_per_edge, d_edge_starts[i], d_edge_targets[i], d_blocks[i],
                                     streams ? streams[i] : nullptr, &lays[i]);
            if (codes[i] != UPC_OK) messages[i] = upc_last_error(); /* the error text is thread-local */
        });
    }
    for (std::thread& w : workers) w.join();
    for (int i = 0; i < n; i++)
        if (codes[i] != UPC_OK)
        {
            set_error(messages[i]);
            return codes[i];
        }
    return all_gather_blocks(n, comms, d_blocks, lays[0].block_bytes, streams);
}

} /* extern "C" */
