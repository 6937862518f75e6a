/*
 * upc_cluster.cu - outcome clustering on the GPU.
 *
 * Replaces the index-clustering half of UncertaintyPlanningSpace::ClusterParticles
 * (uncertainty_contact_planning.hpp:1986-2034): PerformSpatialFeatureClustering (:1915) followed by
 * RunDistanceClusteringOnInitialClusters (:1947), whose heavy callees are arc_helpers::BuildDistanceMatrix
 * (:1674, :1960) and SimpleHierarchicalClustering::Cluster (:1684, :1974).
 *
 * The reference materialises an N x N FP64 matrix per pass.  Here each pass is fused into a 1-bit
 * threshold-adjacency matrix (N^2/8 bytes) built by a tiled pair kernel; the partition is then read off it:
 *   - if every row equals the row of its smallest neighbour, the threshold graph is a disjoint union of
 *     cliques and complete-link clustering cut at the threshold returns exactly those cliques
 *      - this also covers the reference's own "max <= threshold" shortcut;
 *   - otherwise the affected connected components are clustered exactly (Lance-Williams complete link with
 *     the deterministic tie-break of DESIGN.md 5.2) from their FP64 distance sub-matrices, on the GPU.
 * Labels are canonical: clusters numbered by increasing smallest member.
 */
#include "upc_internal.h"
#include "upc_sim.cuh"
#include <cooperative_groups.h>

#include <algorithm>
#include <map>

namespace cg = cooperative_groups;

namespace upc
{
constexpr int kTileRows = 64;
constexpr int kTileCols = 256;

/* ------------------------------------------------------------------ pair distances on per-particle features */

/* features: SE2 (x, y, theta); SE3 (tx, ty, tz, qx, qy, qz, qw); linked (joint values) */
template <int KIND>
struct PairDistance;

template <>
struct PairDistance<UPC_ROBOT_SE2>
{
    static constexpr int MAXF = 3;
    __device__ __forceinline__ static double eval(const DevRobot& r, const double* a, const double* b)
    {
        return Se2Model::distance(r, a, b);
    }
    /* eval(a, b) <= thr; the square root is only taken inside a 1e-14-wide band around the threshold */
    __device__ __forceinline__ static bool within(const DevRobot& r, const double* a, const double* b, double thr)
    {
        const double dx = b[0] - a[0];
        const double dy = b[1] - a[1];
        const double rot = fabs(upc_angle_signed_distance(a[2], b[2]));
        if (rot > thr) return false; /* the translation part is >= 0 */
        const double t2 = (dx * dx) + (dy * dy);
        const double rem = thr - rot;
        const double r2 = rem * rem;
        if (rem > 0.1 * thr)
        {
            /* margins of 1e-14 relative to rem are then also > 4 ulp relative to thr */
            if (t2 > r2 * (1.0 + 1.0e-14)) return false;
            if (t2 <= r2 * (1.0 - 1.0e-14)) return true;
        }
        return (sqrt(t2) + rot) <= thr;
    }
};

template <>
struct PairDistance<UPC_ROBOT_SE3>
{
    static constexpr int MAXF = 7;
    __device__ __forceinline__ static double eval(const DevRobot&, const double* a, const double* b)
    {
        /* simple_robot_models.hpp:871-874 with the quaternions extracted once per particle */
        const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const double tdist = sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
        const double dq = fabs((((a[3] * b[3]) + (a[4] * b[4])) + (a[5] * b[5])) + (a[6] * b[6]));
        double rdist = 0.0;
        if (dq < (1.0 - 2.220446049250313e-16))
        {
            rdist = upc_acos(((2.0 * dq) * dq) - 1.0);
        }
        return (((1.0 - 0.5) * tdist) + (0.5 * rdist)) * 2.0;
    }
    /* eval(a, b) <= thr, decided without the arccosine whenever rigorous bounds allow:
     * the distance is |dt| + angle with angle = acos(c) >= 0, and sqrt(2(1-c)) <= acos(c) <= pi*sqrt((1-c)/2). */
    __device__ __forceinline__ static bool within(const DevRobot& r, const double* a, const double* b, double thr)
    {
        const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const double tdist = sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
        if (tdist > thr) return false;
        const double dq = fabs((((a[3] * b[3]) + (a[4] * b[4])) + (a[5] * b[5])) + (a[6] * b[6]));
        if (!(dq < (1.0 - 2.220446049250313e-16))) return tdist <= thr;
        const double c = ((2.0 * dq) * dq) - 1.0;
        const double sq = sqrt(1.0 - c);
        if ((tdist + (sq * 1.4142135623730951) * (1.0 - 1.0e-9)) > thr * (1.0 + 1.0e-9)) return false;
        if ((tdist + (sq * 2.2214414690791831) * (1.0 + 1.0e-9)) <= thr * (1.0 - 1.0e-9)) return true;
        return eval(r, a, b) <= thr;
    }
};

template <>
struct PairDistance<UPC_ROBOT_LINKED>
{
    static constexpr int MAXF = UPC_MAX_DOF;
    __device__ __forceinline__ static double eval(const DevRobot& r, const double* a, const double* b)
    {
        return LinkedModel::distance(r, a, b);
    }
    __device__ __forceinline__ static bool within(const DevRobot& r, const double* a, const double* b, double thr) { return eval(r, a, b) <= thr; }
    /* same test with the degree-of-freedom count known at compile time (fully unrolled, features stay in registers)
     * and the square root only taken inside a 1e-14-wide band around the threshold */
    template <int NF>
    __device__ __forceinline__ static bool within_n(const DevRobot& r, const double* a, const double* b, double thr)
    {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < NF; k++)
        {
            const double raw = r.axis_continuous[k] ? upc_angle_signed_distance(a[k], b[k]) : (b[k] - a[k]);
            const double wd = raw * r.weights[k];
            sum = sum + (wd * wd);
        }
        const double t2 = thr * thr;
        if (thr < 0.0 || sum > t2 * (1.0 + 1.0e-14)) return false;
        if (sum <= t2 * (1.0 - 1.0e-14)) return true;
        return sqrt(sum) <= thr;
    }
};

__global__ void se3_features_kernel(const double* __restrict__ cfg, double* __restrict__ feat, int64_t total)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double T[12];
#pragma unroll
    for (int r = 0; r < 3; r++)
    {
#pragma unroll
        for (int k = 0; k < 3; k++) T[r * 4 + k] = cfg[(int64_t)(r * 3 + k) * total + i];
        T[r * 4 + 3] = cfg[(int64_t)(9 + r) * total + i];
    }
    double q[4];
    quat_from_xf(T, q);
    feat[0 * total + i] = T[3];
    feat[1 * total + i] = T[7];
    feat[2 * total + i] = T[11];
    feat[3 * total + i] = q[0];
    feat[4 * total + i] = q[1];
    feat[5 * total + i] = q[2];
    feat[6 * total + i] = q[3];
}

struct AdjArgs
{
    DevRobot robot;
    const double* feat; /* [F][total] */
    int nfeat;
    int64_t total, set_size;
    int wpr; /* 32-bit words per adjacency row */
    const int32_t* group; /* optional first-pass cluster id per particle (pairs in different groups are never adjacent) */
    uint32_t* adj;        /* [total][wpr] */
    double threshold;
};

/*
 * Threshold adjacency, one CTA per 64-row x 256-column tile of one particle set.
 * Thread = column; the 64 row feature vectors sit in shared memory; one warp ballot per row yields a packed
 * 32-column word; the 64x8 words of the tile are staged in shared memory and written as full 32-byte sectors.
 * Distances are symmetric (bit for bit), so tiles lying entirely below the diagonal are skipped and filled in by
 * mirror_lower_kernel.  NF > 0 fixes the feature count at compile time (linked robots: the degrees of freedom).
 */
template <int KIND, int NF>
__global__ void __launch_bounds__(kTileCols) adjacency_kernel(const __grid_constant__ AdjArgs a)
{
    constexpr int MAXF = (NF > 0) ? NF : PairDistance<KIND>::MAXF;
    const int64_t row0 = (int64_t)blockIdx.y * kTileRows;
    const int64_t col_first = (int64_t)blockIdx.x * kTileCols;
    if (col_first + kTileCols - 1 < row0) return; /* whole tile below the diagonal: mirrored later */
    __shared__ double s_row[MAXF][kTileRows];
    __shared__ int32_t s_grp[kTileRows];
    __shared__ uint32_t s_words[kTileRows][kTileCols / 32];
    const int64_t set_base = (int64_t)blockIdx.z * a.set_size;
    const int64_t col = col_first + threadIdx.x;
    const int F = (NF > 0) ? NF : a.nfeat;
    for (int k = threadIdx.x; k < F * kTileRows; k += kTileCols)
    {
        const int f = k / kTileRows, r = k % kTileRows;
        const int64_t i = row0 + r;
        s_row[f][r] = (i < a.set_size) ? a.feat[(int64_t)f * a.total + set_base + i] : 0.0;
    }
    if (threadIdx.x < kTileRows)
    {
        const int64_t i = row0 + threadIdx.x;
        s_grp[threadIdx.x] = (a.group && i < a.set_size) ? a.group[set_base + i] : 0;
    }
    double cf[MAXF];
    const bool col_ok = col < a.set_size;
#pragma unroll
    for (int f = 0; f < MAXF; f++) cf[f] = (col_ok && f < F) ? a.feat[(int64_t)f * a.total + set_base + col] : 0.0;
    const int32_t cgrp = (a.group && col_ok) ? a.group[set_base + col] : 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = 0; r < kTileRows; r++)
    {
        bool adjacent = false;
        if (col_ok && (row0 + r) < a.set_size && s_grp[r] == cgrp)
        {
            double rf[MAXF];
#pragma unroll
            for (int f = 0; f < MAXF; f++) rf[f] = s_row[f][r];
            if constexpr (KIND == UPC_ROBOT_LINKED && NF > 0) adjacent = PairDistance<KIND>::template within_n<NF>(a.robot, rf, cf, a.threshold);
            else adjacent = PairDistance<KIND>::within(a.robot, rf, cf, a.threshold);
        }
        const unsigned word = __ballot_sync(0xffffffffu, adjacent);
        if (lane == 0) s_words[r][warp] = word;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kTileRows * (kTileCols / 32); k += kTileCols)
    {
        const int r = k / (kTileCols / 32), w = k % (kTileCols / 32);
        const int64_t i = row0 + r;
        const int64_t word_index = (int64_t)blockIdx.x * (kTileCols / 32) + w;
        if (i < a.set_size && word_index < a.wpr) a.adj[(set_base + i) * a.wpr + word_index] = s_words[r][w];
    }
}

/* fills the adjacency words of the tiles adjacency_kernel skipped from their transposes: one warp per 32x32 bit block */
__global__ void __launch_bounds__(256) mirror_lower_kernel(uint32_t* __restrict__ adj, int64_t set_size, int wpr, int64_t n_sets)
{
    const int64_t blocks_per_row = wpr; /* 32-column blocks */
    const int64_t row_blocks = (set_size + 31) / 32;
    const int64_t per_set = row_blocks * blocks_per_row;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= per_set * n_sets) return;
    const int64_t set = w / per_set;
    const int64_t rb = (w - set * per_set) / blocks_per_row; /* rows rb*32 .. */
    const int64_t cb = (w - set * per_set) % blocks_per_row; /* columns cb*32 .. */
    /* was the tile containing this block skipped?  (tile = 64 rows x 256 columns) */
    const int64_t tile_row0 = (rb * 32 / kTileRows) * kTileRows;
    const int64_t tile_col_last = (cb * 32 / kTileCols) * kTileCols + kTileCols - 1;
    if (!(tile_col_last < tile_row0)) return;
    uint32_t* base = adj + set * set_size * wpr;
    /* transposed source block: rows cb*32 + lane, word rb */
    const int64_t src_row = cb * 32 + lane;
    const uint32_t v = (src_row < set_size) ? base[src_row * wpr + rb] : 0u;
    uint32_t mine = 0u;
#pragma unroll
    for (int b = 0; b < 32; b++)
    {
        const uint32_t t = __ballot_sync(0xffffffffu, (v >> b) & 1u);
        if (lane == b) mine = t;
    }
    const int64_t dst_row = rb * 32 + lane;
    if (dst_row < set_size) base[dst_row * wpr + cb] = mine;
}

/* ------------------------------------------------------------------ convex-region signatures (CRS first pass) */

struct SigArgs
{
    DevEnv env;
    DevRobot robot;
    const double* cfg;
    int64_t total;
    uint32_t* sig; /* [P][total] */
    unsigned long long* hash; /* [total] 64-bit mix of the particle's signature vector */
    double* frames; /* linked only: total * num_links * 12 scratch */
};

/* uncertainty_contact_planning.hpp:1597-1635: one thread per particle, signatures stored point-major */
template <class M>
__global__ void __launch_bounds__(128) signature_kernel(const __grid_constant__ SigArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total) return;
    M m;
    if constexpr (M::HAS_FRAMES)
    {
        m.frames = a.frames + i * (int64_t)a.robot.num_links * 12;
    }
    m.load(a.robot, a.cfg, a.total, i);
    if constexpr (M::HAS_FRAMES) m.fk(a.robot, m.frames);
    unsigned long long h = 0x243F6A8885A308D3ull;
    for (int l = 0; l < a.robot.num_links; l++)
    {
        const int beg = a.robot.link_off[l], end = a.robot.link_off[l + 1];
        if (beg == end) continue;
        double T[12], V[12];
        m.frame(a.robot, l, T);
        voxel_from_link(a.env, T, V);
        int k = beg;
        for (; k + 4 <= end; k += 4) /* four independent lookups in flight; stores and the hash keep the point order */
        {
            int64_t cell[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                double u[3];
                xf_point(V, a.robot.pts[4 * (k + q) + 0], a.robot.pts[4 * (k + q) + 1], a.robot.pts[4 * (k + q) + 2], u);
                cell[q] = voxel_in_bounds(a.env, u) ? (int64_t)cell_index(a.env, floor_to_int(u[0]), floor_to_int(u[1]), floor_to_int(u[2])) : (int64_t)-1;
            }
            uint32_t sv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) sv[q] = (cell[q] >= 0) ? ldg(a.env.seg + cell[q]) : 0u;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                a.sig[(int64_t)(k + q) * a.total + i] = sv[q];
                h = (h ^ (unsigned long long)sv[q]) * 0x9E3779B97F4A7C15ull;
                h ^= h >> 29;
            }
        }
        for (; k < end; k++)
        {
            double u[3];
            xf_point(V, a.robot.pts[4 * k + 0], a.robot.pts[4 * k + 1], a.robot.pts[4 * k + 2], u);
            uint32_t s = 0u;
            if (voxel_in_bounds(a.env, u))
            {
                s = ldg(a.env.seg + cell_index(a.env, floor_to_int(u[0]), floor_to_int(u[1]), floor_to_int(u[2])));
            }
            a.sig[(int64_t)k * a.total + i] = s;
            h = (h ^ (unsigned long long)s) * 0x9E3779B97F4A7C15ull;
            h ^= h >> 29;
        }
    }
    if (h == ~0ull) h = 0ull; /* ~0 marks an empty hash-table slot */
    a.hash[i] = h;
}

/* ---- duplicate-signature elimination: particles of one outcome usually share a handful of
 * distinct signature vectors, and the signature pass only needs one row/column per distinct vector ---- */

struct DedupeArgs
{
    const uint32_t* sig;
    const unsigned long long* hash;
    int num_points;
    int64_t total, set_size;
    int table_size; /* per set, power of two >= 2*set_size */
    unsigned long long* keys; /* [n_sets][table_size], ~0 = empty */
    int32_t* reps;            /* [n_sets][table_size], smallest particle (index within the set) seen with the key */
    int32_t* urep;            /* [total] representative (index within the set) with an identical signature vector */
    int32_t* ucount;          /* [n_sets] number of distinct vectors */
    int32_t* ulist;           /* [n_sets][set_size] representatives, in arrival order */
    int32_t* uid;             /* [total] position of the particle's representative in ulist */
};

__global__ void dedupe_insert_kernel(const __grid_constant__ DedupeArgs a)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.total) return;
    const int64_t set = p / a.set_size;
    const int i = (int)(p - set * a.set_size);
    const unsigned long long h = a.hash[p];
    unsigned long long* keys = a.keys + set * a.table_size;
    int32_t* reps = a.reps + set * a.table_size;
    unsigned slot = (unsigned)(h >> 17) & (unsigned)(a.table_size - 1);
    while (true)
    {
        const unsigned long long old = atomicCAS(&keys[slot], ~0ull, h);
        if (old == ~0ull || old == h)
        {
            atomicMin(&reps[slot], i);
            break;
        }
        slot = (slot + 1u) & (unsigned)(a.table_size - 1);
    }
}

__global__ void dedupe_resolve_kernel(const __grid_constant__ DedupeArgs a)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.total) return;
    const int64_t set = p / a.set_size;
    const int64_t base = set * a.set_size;
    const int i = (int)(p - base);
    const unsigned long long h = a.hash[p];
    const unsigned long long* keys = a.keys + set * a.table_size;
    const int32_t* reps = a.reps + set * a.table_size;
    unsigned slot = (unsigned)(h >> 17) & (unsigned)(a.table_size - 1);
    while (keys[slot] != h) slot = (slot + 1u) & (unsigned)(a.table_size - 1);
    int r = reps[slot];
    /* equal hashes are only a hint: the vectors themselves decide */
    bool same = true;
    int k = 0;
    for (; k + 8 <= a.num_points && same; k += 8) /* eight independent load pairs in flight per trip */
    {
        uint32_t x[8], y[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            x[q] = a.sig[(int64_t)(k + q) * a.total + base + i];
            y[q] = a.sig[(int64_t)(k + q) * a.total + base + r];
        }
#pragma unroll
        for (int q = 0; q < 8; q++) same = same && (x[q] == y[q]);
    }
    for (; k < a.num_points && same; k++) same = a.sig[(int64_t)k * a.total + base + i] == a.sig[(int64_t)k * a.total + base + r];
    if (!same) r = i;
    a.urep[p] = r;
    if (r == i)
    {
        const int pos = atomicAdd(&a.ucount[set], 1);
        a.ulist[base + pos] = i;
        a.uid[p] = pos;
    }
}

__global__ void dedupe_uid_kernel(const __grid_constant__ DedupeArgs a)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.total) return;
    const int64_t base = (p / a.set_size) * a.set_size;
    const int r = a.urep[p];
    if (r != (int)(p - base)) a.uid[p] = a.uid[base + r];
}

/* compact signature matrix of the distinct vectors: usig[k][set*umax + u] */
__global__ void dedupe_gather_kernel(const uint32_t* __restrict__ sig, int num_points, int64_t total, int64_t set_size,
                                     const int32_t* __restrict__ ucount, const int32_t* __restrict__ ulist, int umax, uint32_t* __restrict__ usig)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_sets = total / set_size;
    const int64_t utotal = n_sets * umax;
    if (e >= utotal * num_points) return;
    const int k = (int)(e / utotal);
    const int64_t q = e - (int64_t)k * utotal;
    const int64_t set = q / umax;
    const int u = (int)(q - set * umax);
    uint32_t v = 0u;
    if (u < ucount[set]) v = sig[(int64_t)k * total + set * set_size + ulist[set * set_size + u]];
    usig[e] = v;
}

__global__ void dedupe_expand_kernel(const int32_t* __restrict__ uroot, const int32_t* __restrict__ uid, int64_t total, int64_t set_size, int umax,
                                     int32_t* __restrict__ group)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int64_t set = p / set_size;
    group[p] = uroot[set * umax + uid[p]];
}

struct SigAdjArgs
{
    const uint32_t* sig; /* [P][total] */
    int num_points;
    int64_t total, set_size;
    int wpr;
    uint32_t* adj;
    int max_mismatch; /* largest count m with (double)m / (double)P <= signature_matching_threshold */
    const int32_t* set_count; /* optional: valid rows/columns per set (padded compact matrices) */
};

/*
 * Signature-mismatch adjacency (uncertainty_contact_planning.hpp:1637-1661 + the threshold at :1677-1685):
 * a 64x64 tile of pairs per CTA, 4x4 pairs per thread, 32 body points per shared-memory stage.
 */
__global__ void __launch_bounds__(256) signature_adjacency_kernel(const __grid_constant__ SigAdjArgs a)
{
    __shared__ uint32_t s_a[32][64];
    __shared__ uint32_t s_b[32][64];
    __shared__ uint32_t s_bits[64][2];
    const int64_t set_base = (int64_t)blockIdx.z * a.set_size;
    const int64_t row0 = (int64_t)blockIdx.y * 64;
    const int64_t col0 = (int64_t)blockIdx.x * 64;
    const int64_t valid = a.set_count ? (int64_t)a.set_count[blockIdx.z] : a.set_size;
    if (row0 >= valid && col0 >= valid)
    {
        /* tile entirely in the padding: rows stay empty */
        if (threadIdx.x < 128)
        {
            const int rr = threadIdx.x >> 1, w = threadIdx.x & 1;
            const int64_t word_index = (int64_t)blockIdx.x * 2 + w;
            if ((row0 + rr) < a.set_size && word_index < a.wpr) a.adj[(set_base + row0 + rr) * a.wpr + word_index] = 0u;
        }
        return;
    }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0;
    if (threadIdx.x < 128) s_bits[threadIdx.x >> 1][threadIdx.x & 1] = 0u;
    for (int k0 = 0; k0 < a.num_points; k0 += 32)
    {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * 64; e += 256)
        {
            const int kk = e >> 6, r = e & 63;
            const int k = k0 + kk;
            const int64_t ri = row0 + r, ci = col0 + r;
            s_a[kk][r] = (k < a.num_points && ri < a.set_size) ? a.sig[(int64_t)k * a.total + set_base + ri] : 0u;
            s_b[kk][r] = (k < a.num_points && ci < a.set_size) ? a.sig[(int64_t)k * a.total + set_base + ci] : 0u;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 32; kk++)
        {
            uint32_t av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; r++) av[r] = s_a[kk][ty + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; c++) bv[c] = s_b[kk][tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    acc[r][c] += ((av[r] != 0u) && (bv[c] != 0u) && ((av[r] & bv[c]) == 0u)) ? 1 : 0;
                }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            const int rr = ty + 16 * r, cc = tx + 16 * c;
            if ((row0 + rr) < valid && (col0 + cc) < valid && acc[r][c] <= a.max_mismatch)
            {
                atomicOr(&s_bits[rr][cc >> 5], 1u << (cc & 31));
            }
        }
    __syncthreads();
    if (threadIdx.x < 128)
    {
        const int rr = threadIdx.x >> 1, w = threadIdx.x & 1;
        const int64_t word_index = (int64_t)blockIdx.x * 2 + w;
        if ((row0 + rr) < a.set_size && word_index < a.wpr) a.adj[(set_base + row0 + rr) * a.wpr + word_index] = s_bits[rr][w];
    }
}

/* ------------------------------------------------------------------ actuation-centre connectivity (AC first pass) */

struct CenterArgs
{
    DevRobot robot;
    const double* cfg;
    int64_t total;
    double* centers; /* [num_links*3][total]: link frame origins */
    double* frames;  /* linked only */
};

/* link origins of every particle (uncertainty_contact_planning.hpp:1718-1748) */
template <class M>
__global__ void __launch_bounds__(128) link_centers_kernel(const __grid_constant__ CenterArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total) return;
    M m;
    if constexpr (M::HAS_FRAMES) m.frames = a.frames + i * (int64_t)a.robot.num_links * 12;
    m.pid_i = nullptr;
    m.pid_e = nullptr;
    m.load(a.robot, a.cfg, a.total, i);
    if constexpr (M::HAS_FRAMES) m.fk(a.robot, m.frames);
    for (int l = 0; l < a.robot.num_links; l++)
    {
        double T[12];
        m.frame(a.robot, l, T);
        a.centers[(int64_t)(l * 3 + 0) * a.total + i] = T[3];
        a.centers[(int64_t)(l * 3 + 1) * a.total + i] = T[7];
        a.centers[(int64_t)(l * 3 + 2) * a.total + i] = T[11];
    }
}

/* CheckStraightLine3DPath (uncertainty_contact_planning.hpp:1698-1716): occupancy sampled at `steps` points from a to b */
__device__ __forceinline__ bool straight_line_free(const DevEnv& env, const double* V, const double* a, const double* b)
{
    const double d0 = b[0] - a[0], d1 = b[1] - a[1], d2 = b[2] - a[2];
    const double distance = sqrt(((d0 * d0) + (d1 * d1)) + (d2 * d2));
    unsigned steps = (unsigned)ceil(distance / env.res);
    if (steps < 1u) steps = 1u;
    for (unsigned step = 1u; step <= steps; step++)
    {
        const double percent = (double)step / (double)steps;
        const double q0 = upc_fma(d0, percent, a[0]), q1 = upc_fma(d1, percent, a[1]), q2 = upc_fma(d2, percent, a[2]);
        double u[3];
        xf_point(V, q0, q1, q2, u);
        float occupancy = env.occ_oob; /* the in-bounds flag is ignored at :1709: the default cell is used */
        if (voxel_in_bounds(env, u)) occupancy = ldg(env.occ + cell_index(env, floor_to_int(u[0]), floor_to_int(u[1]), floor_to_int(u[2])));
        if (occupancy > 0.5f) return false;
    }
    return true;
}

struct AcAdjArgs
{
    DevEnv env;
    const double* centers; /* [L*3][total] */
    int num_links;
    int64_t total, set_size;
    int wpr;
    uint32_t* adj;
};

/* adjacency bit = every link origin of the lower-indexed particle sees the same link's origin of the other one */
__global__ void __launch_bounds__(256) ac_adjacency_kernel(const __grid_constant__ AcAdjArgs a)
{
    __shared__ uint32_t s_words[kTileRows][kTileCols / 32];
    const int64_t set_base = (int64_t)blockIdx.z * a.set_size;
    const int64_t row0 = (int64_t)blockIdx.y * kTileRows;
    const int64_t col = (int64_t)blockIdx.x * kTileCols + threadIdx.x;
    const bool col_ok = col < a.set_size;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double V[12];
    {
        const double I[12] = {1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0};
        voxel_from_link(a.env, I, V);
    }
    for (int r = 0; r < kTileRows; r++)
    {
        const int64_t row = row0 + r;
        bool adjacent = false;
        if (col_ok && row < a.set_size)
        {
            adjacent = true;
            if (row != col)
            {
                const int64_t lo = (row < col) ? row : col, hi = (row < col) ? col : row;
                for (int l = 0; l < a.num_links && adjacent; l++)
                {
                    double pa[3], pb[3];
                    for (int c = 0; c < 3; c++)
                    {
                        pa[c] = a.centers[(int64_t)(l * 3 + c) * a.total + set_base + lo];
                        pb[c] = a.centers[(int64_t)(l * 3 + c) * a.total + set_base + hi];
                    }
                    adjacent = straight_line_free(a.env, V, pa, pb);
                }
            }
        }
        const unsigned word = __ballot_sync(0xffffffffu, adjacent);
        if (lane == 0) s_words[r][warp] = word;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kTileRows * (kTileCols / 32); k += kTileCols)
    {
        const int r = k / (kTileCols / 32), w = k % (kTileCols / 32);
        const int64_t i = row0 + r;
        const int64_t word_index = (int64_t)blockIdx.x * (kTileCols / 32) + w;
        if (i < a.set_size && word_index < a.wpr) a.adj[(set_base + i) * a.wpr + word_index] = s_words[r][w];
    }
}

/* ------------------------------------------------------------------ point-to-point movement (PTPM first pass) */

/* start/target pairs of uncertainty_contact_planning.hpp:1804-1825: for i < j the simulations (i -> j) then (j -> i) */
__global__ void __launch_bounds__(256) ptpm_pairs_kernel(const double* __restrict__ cfg, int width, int64_t n, double* __restrict__ starts,
                                                         double* __restrict__ targets)
{
    const int64_t sims = n * n - n;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (n * (n - 1)) / 2) return;
    /* unrank e -> (i, j), i < j, row-major over the strict upper triangle */
    int64_t i = (int64_t)floor(((double)(2 * n - 1) - sqrt((double)(2 * n - 1) * (double)(2 * n - 1) - 8.0 * (double)e)) * 0.5);
    while (i * n - (i * (i + 1)) / 2 > e) i--;
    while ((i + 1) * n - ((i + 1) * (i + 2)) / 2 <= e) i++;
    const int64_t j = e - (i * n - (i * (i + 1)) / 2) + i + 1;
    for (int c = 0; c < width; c++)
    {
        const double ci = cfg[(int64_t)c * n + i], cj = cfg[(int64_t)c * n + j];
        starts[(int64_t)c * sims + 2 * e] = ci;
        targets[(int64_t)c * sims + 2 * e] = cj;
        starts[(int64_t)c * sims + 2 * e + 1] = cj;
        targets[(int64_t)c * sims + 2 * e + 1] = ci;
    }
}

struct PtpmAdjArgs
{
    DevRobot robot;
    const double* cfg;     /* [C][n] */
    const double* results; /* [C][n*n-n] */
    int width;
    int64_t n;
    int wpr;
    uint32_t* adj; /* zeroed; diagonal and feasible pairs are set */
    double goal_distance_threshold;
};

template <int KIND>
__global__ void __launch_bounds__(256) ptpm_adjacency_kernel(const __grid_constant__ PtpmAdjArgs a)
{
    const int64_t n = a.n;
    const int64_t sims = n * n - n;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) atomicOr(&a.adj[e * a.wpr + (e >> 5)], 1u << (e & 31));
    if (e >= (n * (n - 1)) / 2) return;
    int64_t i = (int64_t)floor(((double)(2 * n - 1) - sqrt((double)(2 * n - 1) * (double)(2 * n - 1) - 8.0 * (double)e)) * 0.5);
    while (i * n - (i * (i + 1)) / 2 > e) i--;
    while ((i + 1) * n - ((i + 1) * (i + 2)) / 2 <= e) i++;
    const int64_t j = e - (i * n - (i * (i + 1)) / 2) + i + 1;
    double fr[16], br[16], ci[16], cj[16];
    for (int c = 0; c < a.width; c++)
    {
        fr[c] = a.results[(int64_t)c * sims + 2 * e];
        br[c] = a.results[(int64_t)c * sims + 2 * e + 1];
        ci[c] = a.cfg[(int64_t)c * n + i];
        cj[c] = a.cfg[(int64_t)c * n + j];
    }
    double df, db;
    if (KIND == UPC_ROBOT_SE2) { df = Se2Model::distance(a.robot, fr, cj); db = Se2Model::distance(a.robot, br, ci); }
    else if (KIND == UPC_ROBOT_SE3) { df = Se3Model::distance(a.robot, fr, cj); db = Se3Model::distance(a.robot, br, ci); }
    else { df = LinkedModel::distance(a.robot, fr, cj); db = LinkedModel::distance(a.robot, br, ci); }
    if ((df <= a.goal_distance_threshold) && (db <= a.goal_distance_threshold))
    {
        atomicOr(&a.adj[i * a.wpr + (j >> 5)], 1u << (j & 31));
        atomicOr(&a.adj[j * a.wpr + (i >> 5)], 1u << (i & 31));
    }
}

/* ------------------------------------------------------------------ reading the partition off the adjacency */

/* One warp per row: root = smallest neighbour; row_ok = (row == row of root). */
__global__ void __launch_bounds__(256) root_check_kernel(const uint32_t* __restrict__ adj, int64_t total, int64_t set_size, int wpr,
                                                         int32_t* __restrict__ root, uint8_t* __restrict__ row_ok,
                                                         int32_t* __restrict__ set_bad)
{
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= total) return;
    const int64_t set = p / set_size;
    const uint32_t* row = adj + p * wpr;
    /* first set bit */
    int first = -1;
    for (int w0 = 0; w0 < wpr && first < 0; w0 += 32)
    {
        const int w = w0 + lane;
        const uint32_t v = (w < wpr) ? row[w] : 0u;
        const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
        if (nz)
        {
            const int src = __ffs((int)nz) - 1;
            const uint32_t word = __shfl_sync(0xffffffffu, v, src);
            first = (w0 + src) * 32 + (__ffs((int)word) - 1);
        }
    }
    /* the diagonal guarantees a set bit; keep a defined value anyway */
    if (first < 0) first = (int)(p - set * set_size);
    const uint32_t* rrow = adj + (set * set_size + first) * wpr;
    bool same = true;
    for (int w = lane; w < wpr; w += 32) same = same && (row[w] == rrow[w]);
    same = __all_sync(0xffffffffu, same);
    if (lane == 0)
    {
        root[p] = first;
        row_ok[p] = same ? 1 : 0;
        if (!same) atomicOr(&set_bad[set], 1);
    }
}

__global__ void iota_in_set_kernel(int32_t* out, int64_t total, int64_t set_size)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < total) out[p] = (int32_t)(p % set_size);
}

/* canonical labels: one CTA per set; clusters numbered by increasing representative */
__global__ void __launch_bounds__(1024) label_kernel(const int32_t* __restrict__ root, int64_t set_size, int32_t* __restrict__ rank_scratch,
                                                     int32_t* __restrict__ labels, int32_t* __restrict__ n_clusters)
{
    __shared__ int s_warp[32];
    __shared__ int s_running;
    const int64_t base = (int64_t)blockIdx.x * set_size;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < set_size; c0 += 1024)
    {
        const int64_t i = c0 + threadIdx.x;
        const bool is_root = (i < set_size) && (root[base + i] == (int32_t)i);
        const unsigned b = __ballot_sync(0xffffffffu, is_root);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        int before = s_running;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        if (is_root) rank_scratch[base + i] = before + __popc(b & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x == 0)
        {
            int tot = 0;
            for (int w = 0; w < 32; w++) tot += s_warp[w];
            s_running += tot;
        }
        __syncthreads();
    }
    for (int64_t i = threadIdx.x; i < set_size; i += 1024) labels[base + i] = rank_scratch[base + root[base + i]];
    if (threadIdx.x == 0) n_clusters[blockIdx.x] = s_running;
}

/* ------------------------------------------------------------------ pivot pass: blob-shaped outcomes in O(n * clusters) */

struct PivotArgs
{
    DevRobot robot;
    const double* feat; /* [F][total] */
    int nfeat;
    int64_t total, set_size;
    const int32_t* group; /* optional first-pass roots (in-set indices) */
    double threshold;
    double radius;        /* threshold / 2 - margin */
    double separation;    /* 2 * threshold + margin */
    double* scratch;      /* [2][total]: distances to the current pivot; largest distance to the extreme points */
    int32_t* root;        /* out: in-set index of the smallest member of the particle's cluster */
    uint8_t* ok;          /* out, per set: 1 = the partition written to root is the complete-link result */
};

constexpr int kMaxPivots = 64;
constexpr int kPivotCtas = 8; /* thread blocks per cluster when a cluster shares one set (portable maximum) */

/* CTA-wide minimum of (key, index) pairs, smaller index on equal keys; every thread receives the result */
__device__ __forceinline__ void block_min_pair(double& key, int& idx, double* s_key, int* s_idx)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        const double ok = __shfl_xor_sync(0xffffffffu, key, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (ok < key || (ok == key && oi < idx)) { key = ok; idx = oi; }
    }
    __syncthreads(); /* previous users of the scratch are done */
    if (lane == 0) { s_key[warp] = key; s_idx[warp] = idx; }
    __syncthreads();
    key = s_key[0]; idx = s_idx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
        if (s_key[w] < key || (s_key[w] == key && s_idx[w] < idx)) { key = s_key[w]; idx = s_idx[w]; }
}

/*
 * The same minimum over the CTAS thread blocks of a cluster (distributed shared memory): every CTA reduces its own threads, publishes
 * the result in its slot `parity` and reads the slots of all ranks after ONE cluster barrier.  Slots alternate between calls: a CTA
 * can only overwrite slot `parity` two calls later, i.e. after the barrier of the call in between, which every CTA passes only after
 * it has finished reading.  Every thread of every CTA receives the same result.
 */
template <int CTAS>
__device__ __forceinline__ void team_min_pair(double& key, int& idx, double* s_key, int* s_idx, double* s_ckey, int* s_cidx, int& parity)
{
    block_min_pair(key, idx, s_key, s_idx);
    if constexpr (CTAS > 1)
    {
        cg::cluster_group cluster = cg::this_cluster();
        if (threadIdx.x == 0) { s_ckey[parity] = key; s_cidx[parity] = idx; }
        cluster.sync();
        double best = INFINITY;
        int best_idx = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < CTAS; r++)
        {
            const double* rk = cluster.map_shared_rank(s_ckey, r);
            const int* ri = cluster.map_shared_rank(s_cidx, r);
            const double k2 = rk[parity];
            const int i2 = ri[parity];
            if (r == 0 || k2 < best || (k2 == best && i2 < best_idx)) { best = k2; best_idx = i2; }
        }
        key = best; idx = best_idx;
        parity ^= 1;
    }
}

/*
 * One CTA per set, or - for a few large sets - one cluster of CTAS thread blocks per set (thread-block clusters: the blocks
 * share the sweeps over the set and exchange their partial minima through distributed shared memory).  Repeat: the smallest unassigned particle p opens a group.  If every unassigned particle of its
 * first-pass group that lies within the threshold of p lies within threshold/2 - margin, p is the group's centre;
 * otherwise a more central member is looked for (the candidate minimising its largest distance to p, to the candidate
 * farthest from p and to the candidates farthest from the centres tried so far).  The group takes every unassigned particle of the first-pass group within
 * threshold/2 - margin of the centre, and must contain p.  The configuration distances are metrics, so all pairs inside
 * a group are within the threshold (a clique); if, in addition, centres of the same first-pass group are further than
 * 2 * threshold + margin apart, no pair across groups is, the threshold graph is exactly the disjoint union of these
 * cliques and complete link returns them, numbered by their smallest members p.  The margin (1e-6)
 * covers the rounding of the evaluated distances (<= 2e-8: the arccosine of the SE(3) rotation part near zero angle).
 * Anything else - more than kMaxPivots groups, centres too close, p outside its group - reports ok = 0 and the
 * adjacency pass decides.  The choice of the centre is a heuristic: it only affects whether the set is settled here.
 */
template <int KIND, int CTAS>
__global__ void __launch_bounds__(1024) pivot_pass_kernel(const __grid_constant__ PivotArgs a)
{
    constexpr int MAXF = PairDistance<KIND>::MAXF;
    __shared__ double s_key[32];
    __shared__ int s_idx[32];
    __shared__ int s_centres[kMaxPivots];
    __shared__ int s_fail;
    __shared__ double s_ckey[2];
    __shared__ int s_cidx[2];
    int parity = 0;
    /* the set's thread team: one CTA, or the CTAS blocks of a cluster; particle i always belongs to the same thread of the team, so
     * root / scratch entries are private to their owner between the team-wide reductions */
    const int set = (int)(blockIdx.x / CTAS);
    const int crank = (CTAS > 1) ? (int)(blockIdx.x % CTAS) : 0;
    const int tid = crank * (int)blockDim.x + (int)threadIdx.x;
    const int tstride = CTAS * (int)blockDim.x;
    const int64_t base = (int64_t)set * a.set_size;
    const int n = (int)a.set_size;
    for (int i = tid; i < n; i += tstride) a.root[base + i] = -1;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    int next = (n > 0) ? 0 : 0x7fffffff;
    for (int k = 0; k < kMaxPivots && next != 0x7fffffff; k++)
    {
        const int p = next;
        double fp[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; f++) fp[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + p] : 0.0;
        const int gp = a.group ? a.group[base + p] : 0;
        /* sweep 1: distances of the candidates to p; the farthest one within the threshold */
        double far_key = 1.0; /* keys are negated distances: block_min_pair returns the largest distance */
        int far_idx = 0x7fffffff;
        for (int i = tid; i < n; i += tstride)
        {
            double d = -1.0; /* not a candidate */
            if (a.root[base + i] < 0 && (!a.group || a.group[base + i] == gp))
            {
                if (i == p)
                {
                    d = 0.0;
                }
                else
                {
                    double fi[MAXF];
#pragma unroll
                    for (int f = 0; f < MAXF; f++) fi[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + i] : 0.0;
                    d = PairDistance<KIND>::eval(a.robot, fp, fi);
                    if (!(d >= 0.0)) d = INFINITY; /* NaN: never a member, the separation test then rejects the set */
                }
                if (d <= a.threshold && (-d < far_key || (-d == far_key && i < far_idx))) { far_key = -d; far_idx = i; }
            }
            a.scratch[base + i] = d;
        }
        team_min_pair<CTAS>(far_key, far_idx, s_key, s_idx, s_ckey, s_cidx, parity);
        int centre = p;
        if (-far_key > a.radius)
        {
            /* a more central candidate: minimise the largest distance to a growing set of extreme points (p, the candidate
             * farthest from p, then the candidates farthest from the successive centres) - up to six rounds */
            double* runmax = a.scratch + a.total; /* per candidate: largest distance to the extremes so far */
            int ext = far_idx;
            for (int round = 0; round < 6; round++)
            {
                double fe[MAXF];
#pragma unroll
                for (int f = 0; f < MAXF; f++) fe[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + ext] : 0.0;
                double best_key = INFINITY;
                int best_idx = 0x7fffffff;
                for (int i = tid; i < n; i += tstride)
                {
                    const double dp = a.scratch[base + i];
                    if (dp < 0.0 || dp > a.threshold) continue;
                    double fi[MAXF];
#pragma unroll
                    for (int f = 0; f < MAXF; f++) fi[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + i] : 0.0;
                    const double de = (i == ext) ? 0.0 : PairDistance<KIND>::eval(a.robot, fe, fi);
                    const double before = (round == 0) ? dp : runmax[base + i];
                    const double key = (before > de) ? before : de;
                    runmax[base + i] = key;
                    if (key < best_key || (key == best_key && i < best_idx)) { best_key = key; best_idx = i; }
                }
                team_min_pair<CTAS>(best_key, best_idx, s_key, s_idx, s_ckey, s_cidx, parity);
                if (best_idx == 0x7fffffff) break;
                centre = best_idx;
                /* how far is the farthest candidate from this centre? */
                double fcand[MAXF];
#pragma unroll
                for (int f = 0; f < MAXF; f++) fcand[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + centre] : 0.0;
                double worst_key = 1.0;
                int worst_idx = 0x7fffffff;
                for (int i = tid; i < n; i += tstride)
                {
                    const double dp = a.scratch[base + i];
                    if (dp < 0.0 || dp > a.threshold) continue;
                    double fi[MAXF];
#pragma unroll
                    for (int f = 0; f < MAXF; f++) fi[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + i] : 0.0;
                    const double dc = (i == centre) ? 0.0 : PairDistance<KIND>::eval(a.robot, fcand, fi);
                    if (-dc < worst_key || (-dc == worst_key && i < worst_idx)) { worst_key = -dc; worst_idx = i; }
                }
                team_min_pair<CTAS>(worst_key, worst_idx, s_key, s_idx, s_ckey, s_cidx, parity);
                if (-worst_key <= a.radius || worst_idx == 0x7fffffff) break;
                ext = worst_idx;
            }
        }
        double fc[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; f++) fc[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + centre] : 0.0;
        /* fail fast: a new centre must be clearly separated from the earlier centres of its first-pass group */
        if ((int)threadIdx.x < k)
        {
            const int cb = s_centres[threadIdx.x];
            if (!a.group || a.group[base + cb] == gp)
            {
                double fb[MAXF];
#pragma unroll
                for (int f = 0; f < MAXF; f++) fb[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + cb] : 0.0;
                if (!(PairDistance<KIND>::eval(a.robot, fb, fc) > a.separation)) s_fail = 1;
            }
        }
        /* sweep 3: the group */
        double min_key = (double)0x7fffffff;
        int min_idx = 0x7fffffff;
        for (int i = tid; i < n; i += tstride)
        {
            if (a.root[base + i] >= 0) continue;
            const double dp = a.scratch[base + i];
            bool take = false;
            if (dp >= 0.0)
            {
                if (centre == p)
                {
                    take = dp <= a.radius;
                }
                else if (i == centre)
                {
                    take = true;
                }
                else
                {
                    double fi[MAXF];
#pragma unroll
                    for (int f = 0; f < MAXF; f++) fi[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + base + i] : 0.0;
                    take = PairDistance<KIND>::eval(a.robot, fc, fi) <= a.radius;
                }
            }
            if (take) a.root[base + i] = p;
            else if (i < min_idx) { min_idx = i; min_key = (double)i; }
        }
        team_min_pair<CTAS>(min_key, min_idx, s_key, s_idx, s_ckey, s_cidx, parity); /* also orders the root writes before the test below */
        if (threadIdx.x == 0)
        {
            s_centres[k] = centre;
            if (a.root[base + p] != p) s_fail = 1; /* p fell outside its own group */
        }
        __syncthreads();
        if (s_fail) break;
        next = min_idx;
    }
    if (threadIdx.x == 0 && crank == 0) a.ok[set] = (s_fail || next != 0x7fffffff) ? 0 : 1; /* also: more clusters than pivots */
    if constexpr (CTAS > 1) cg::this_cluster().sync(); /* no block may exit while another still reads its slots */
}

/* ------------------------------------------------------------------ exact complete link for sets that are not unions of cliques */

/*
 * Complete-link agglomeration cut at the threshold, parallel form.  Order all cluster pairs by the
 * key (distance, smaller representative, larger representative).  Complete linkage is reducible
 * (d(k, a U b) = max(d(k,a), d(k,b)) never undercuts d(k,a) or d(k,b)), so a pair whose members are each other's
 * nearest neighbour under that order stays so until it is merged: all such reciprocal pairs within the threshold
 * can be merged in the same round, and the partition reached when no pair is left within the threshold is the one
 * the sequential rule produces.  Each round is a handful of fully parallel passes over the FP64 distance matrices
 * of the affected sets; the number of rounds is logarithmic for blob-like data.
 */
struct ExactArgs
{
    DevRobot robot;
    const double* feat;   /* config features [F][total] or null */
    const uint32_t* sig;  /* signatures [P][total] or null */
    const int32_t* group; /* optional first-pass groups: pairs across groups never merge */
    const uint32_t* binary_adj; /* AC / PTPM: distances are 0 (bit set) or 1, read from the adjacency bits */
    int wpr;
    int nfeat, num_points;
    int64_t total, set_size;
    /* the non-clique connected components, concatenated: component c owns flat rows [comp_off[c], comp_off[c+1]) */
    int64_t rows;               /* total number of flat rows */
    const int64_t* members;     /* [rows] global particle index of every row */
    const int32_t* comp_of_row; /* [rows] */
    const int64_t* comp_off;    /* [ncomp+1] */
    const int64_t* matrix_off;  /* [ncomp+1] offsets of the m_c x m_c FP64 matrices */
    double* D;
    uint8_t* active;  /* [rows] */
    int32_t* nn;      /* [rows] nearest active neighbour (local index) under the key order */
    int32_t* partner; /* [rows] local index b if this row absorbs b this round, else -1 */
    int32_t* into;    /* [rows] representative chain (local indices) */
    double threshold;
    int* merged_flag;
};

struct RowRef
{
    int64_t base; /* first flat row of the component */
    int i;        /* local index of the row */
    int m;        /* component size */
    double* D;    /* component matrix */
};

__device__ __forceinline__ RowRef row_ref(const ExactArgs& a, int64_t r)
{
    const int c = a.comp_of_row[r];
    RowRef q;
    q.base = a.comp_off[c];
    q.i = (int)(r - q.base);
    q.m = (int)(a.comp_off[c + 1] - q.base);
    q.D = a.D + a.matrix_off[c];
    return q;
}

/* one CTA per flat row: distances of the row's particle to every member of its component */
template <int KIND>
__global__ void __launch_bounds__(256) exact_matrix_kernel(const __grid_constant__ ExactArgs a)
{
    constexpr int MAXF = PairDistance<KIND>::MAXF;
    for (int64_t r = blockIdx.x; r < a.rows; r += gridDim.x)
    {
        const RowRef q = row_ref(a, r);
        const int64_t pi = a.members[r];
        double fa[MAXF];
#pragma unroll
        for (int f = 0; f < MAXF; f++) fa[f] = (a.feat && f < a.nfeat) ? a.feat[(int64_t)f * a.total + pi] : 0.0;
        for (int j = threadIdx.x; j < q.m; j += blockDim.x)
        {
            double d = 0.0;
            if (j != q.i)
            {
                const int64_t pj = a.members[q.base + j];
                if (a.group && a.group[pi] != a.group[pj])
                {
                    d = INFINITY;
                }
                else if (a.binary_adj)
                {
                    const int64_t cj = pj % a.set_size;
                    d = ((a.binary_adj[pi * a.wpr + (cj >> 5)] >> (cj & 31)) & 1u) ? 0.0 : 1.0;
                }
                else if (a.sig)
                {
                    int mism = 0;
                    for (int k = 0; k < a.num_points; k++)
                    {
                        const uint32_t x = a.sig[(int64_t)k * a.total + pi], y = a.sig[(int64_t)k * a.total + pj];
                        mism += ((x != 0u) && (y != 0u) && ((x & y) == 0u)) ? 1 : 0;
                    }
                    d = (double)mism / (double)a.num_points;
                }
                else
                {
                    double fb[MAXF];
#pragma unroll
                    for (int f = 0; f < MAXF; f++) fb[f] = (f < a.nfeat) ? a.feat[(int64_t)f * a.total + pj] : 0.0;
                    /* smaller index first, like BuildDistanceMatrix evaluates the upper triangle */
                    d = (q.i < j) ? PairDistance<KIND>::eval(a.robot, fa, fb) : PairDistance<KIND>::eval(a.robot, fb, fa);
                }
            }
            q.D[(int64_t)q.i * q.m + j] = d;
        }
        if (threadIdx.x == 0)
        {
            a.active[r] = 1;
            a.into[r] = q.i;
            a.partner[r] = -1;
        }
    }
}

/* nearest active neighbour of every active row: key (distance, min(row, col), max(row, col)); one warp per row */
__global__ void __launch_bounds__(256) exact_nearest_kernel(const __grid_constant__ ExactArgs a)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.rows) return;
    if (!a.active[r])
    {
        if (lane == 0) a.nn[r] = -1;
        return;
    }
    const RowRef q = row_ref(a, r);
    const uint8_t* active = a.active + q.base;
    const double* Drow = q.D + (int64_t)q.i * q.m;
    const int row = q.i;
    double bv = INFINITY;
    int bi = -1;
    for (int c = lane; c < q.m; c += 32)
    {
        if (c == row || !active[c]) continue;
        const double d = Drow[c];
        bool better;
        if (bi < 0) better = true;
        else if (d != bv) better = d < bv;
        else
        {
            const int lo1 = min(row, c), hi1 = max(row, c), lo0 = min(row, bi), hi0 = max(row, bi);
            better = (lo1 < lo0) || (lo1 == lo0 && hi1 < hi0);
        }
        if (better) { bv = d; bi = c; }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        bool better = false;
        if (oi >= 0)
        {
            if (bi < 0) better = true;
            else if (ov != bv) better = ov < bv;
            else
            {
                const int lo1 = min(row, oi), hi1 = max(row, oi), lo0 = min(row, bi), hi0 = max(row, bi);
                better = (lo1 < lo0) || (lo1 == lo0 && hi1 < hi0);
            }
        }
        if (better) { bv = ov; bi = oi; }
    }
    if (lane == 0) a.nn[r] = bi;
}

/* reciprocal nearest pairs within the threshold: the smaller representative absorbs the larger */
__global__ void __launch_bounds__(256) exact_mark_kernel(const __grid_constant__ ExactArgs a)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    int partner = -1;
    const int b = a.nn[r];
    if (b >= 0)
    {
        const RowRef q = row_ref(a, r);
        if (b > q.i && a.nn[q.base + b] == q.i)
        {
            const double d = q.D[(int64_t)q.i * q.m + b];
            if (d <= a.threshold)
            {
                partner = b;
                *a.merged_flag = 1;
            }
        }
    }
    a.partner[r] = partner;
}

/* Lance-Williams, rows: D[a][k] = max(D[a][k], D[b][k]) for every absorbing row a; one CTA per flat row */
__global__ void __launch_bounds__(256) exact_row_merge_kernel(const __grid_constant__ ExactArgs a)
{
    for (int64_t r = blockIdx.x; r < a.rows; r += gridDim.x)
    {
        const int b = a.partner[r];
        if (b < 0) continue;
        const RowRef q = row_ref(a, r);
        double* Da = q.D + (int64_t)q.i * q.m;
        const double* Db = q.D + (int64_t)b * q.m;
        for (int k = threadIdx.x; k < q.m; k += blockDim.x)
        {
            const double x = Da[k], y = Db[k];
            Da[k] = (x < y) ? y : x;
        }
    }
}

/* Lance-Williams, columns: D[k][a] = max(D[k][a], D[k][b]) for every active row k */
__global__ void __launch_bounds__(256) exact_col_merge_kernel(const __grid_constant__ ExactArgs a)
{
    for (int64_t r = blockIdx.x; r < a.rows; r += gridDim.x)
    {
        if (!a.active[r]) continue;
        const RowRef q = row_ref(a, r);
        double* Dk = q.D + (int64_t)q.i * q.m;
        const int32_t* partner = a.partner + q.base;
        for (int c = threadIdx.x; c < q.m; c += blockDim.x)
        {
            const int b = partner[c];
            if (b >= 0 && c != q.i)
            {
                const double x = Dk[c], y = Dk[b];
                Dk[c] = (x < y) ? y : x;
            }
        }
    }
}

__global__ void __launch_bounds__(256) exact_retire_kernel(const __grid_constant__ ExactArgs a)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    const int b = a.partner[r];
    if (b >= 0)
    {
        const RowRef q = row_ref(a, r);
        a.active[q.base + b] = 0;
        a.into[q.base + b] = q.i;
    }
}

__global__ void __launch_bounds__(256) exact_roots_kernel(const __grid_constant__ ExactArgs a, int32_t* root)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.rows) return;
    const RowRef q = row_ref(a, r);
    const int32_t* into = a.into + q.base;
    int k = q.i;
    while (into[k] != k) k = into[k];
    root[a.members[r]] = (int32_t)(a.members[q.base + k] % a.set_size);
}

/*
 * The reciprocal-nearest-neighbour rounds of one component inside ONE CTA (components of at most kFusedExactMax members): the same
 * phases as the kernels above - nearest, mark, row merge, column merge, retire - separated by CTA barriers instead of kernel
 * boundaries, repeated until a round merges nothing.  Components are independent, so every CTA runs its own number of rounds and
 * the host does not take part (the multi-kernel loop with its device-to-host flag per round remains for bigger components).
 *
 * Work per round follows the merges, not the component size: the pairs merged in a round are listed in shared memory, the
 * column update D[k][a] = max(D[k][a], D[k][b]) visits those pairs only, and a row's nearest neighbour is recomputed only when it
 * can have changed - the row absorbed another one, or its nearest neighbour took part in a merge.  (For any other row k every
 * entry D[k][j] either stayed or grew, so the minimum of its keys (distance, min(k, j), max(k, j)) is where it was.)
 */
constexpr int kFusedExactMax = 4096;

__global__ void __launch_bounds__(1024) exact_rounds_fused_kernel(const __grid_constant__ ExactArgs a, int ncomp)
{
    __shared__ int s_npairs;
    __shared__ short2 s_pairs[kFusedExactMax / 2];
    __shared__ uint8_t s_dirty[kFusedExactMax];
    for (int c = blockIdx.x; c < ncomp; c += gridDim.x)
    {
        const int64_t base = a.comp_off[c];
        const int m = (int)(a.comp_off[c + 1] - base);
        if (m > kFusedExactMax) continue;
        double* D = a.D + a.matrix_off[c];
        uint8_t* active = a.active + base;
        int32_t* nn = a.nn + base;
        int32_t* partner = a.partner + base;
        int32_t* into = a.into + base;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
        __syncthreads(); /* separates the components handled by this CTA */
        for (int row = threadIdx.x; row < m; row += blockDim.x) s_dirty[row] = 1;
        __syncthreads();
        while (true)
        {
            /* nearest active neighbour of the rows that may have changed: key (distance, min(row, col), max(row, col)) */
            for (int row = warp; row < m; row += warps)
            {
                if (!s_dirty[row]) continue;
                if (!active[row])
                {
                    if (lane == 0) nn[row] = -1;
                    continue;
                }
                const double* Drow = D + (int64_t)row * m;
                double bv = INFINITY;
                int bi = -1;
                for (int col = lane; col < m; col += 32)
                {
                    if (col == row || !active[col]) continue;
                    const double d = Drow[col];
                    bool better;
                    if (bi < 0) better = true;
                    else if (d != bv) better = d < bv;
                    else
                    {
                        const int lo1 = min(row, col), hi1 = max(row, col), lo0 = min(row, bi), hi0 = max(row, bi);
                        better = (lo1 < lo0) || (lo1 == lo0 && hi1 < hi0);
                    }
                    if (better) { bv = d; bi = col; }
                }
                for (int o = 16; o > 0; o >>= 1)
                {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    bool better = false;
                    if (oi >= 0)
                    {
                        if (bi < 0) better = true;
                        else if (ov != bv) better = ov < bv;
                        else
                        {
                            const int lo1 = min(row, oi), hi1 = max(row, oi), lo0 = min(row, bi), hi0 = max(row, bi);
                            better = (lo1 < lo0) || (lo1 == lo0 && hi1 < hi0);
                        }
                    }
                    if (better) { bv = ov; bi = oi; }
                }
                if (lane == 0) nn[row] = bi;
            }
            if (threadIdx.x == 0) s_npairs = 0;
            __syncthreads();
            /* reciprocal nearest pairs within the threshold: the smaller representative absorbs the larger */
            for (int row = threadIdx.x; row < m; row += blockDim.x)
            {
                s_dirty[row] = 0;
                int pr = -1;
                const int b = nn[row];
                if (b > row && nn[b] == row && D[(int64_t)row * m + b] <= a.threshold)
                {
                    pr = b;
                    const int slot = atomicAdd(&s_npairs, 1);
                    s_pairs[slot] = make_short2((short)row, (short)b);
                }
                partner[row] = pr;
            }
            __syncthreads();
            const int npairs = s_npairs;
            if (npairs == 0) break;
            /* Lance-Williams, rows: D[a][k] = max(D[a][k], D[b][k]) for every merged pair */
            for (int p = warp; p < npairs; p += warps)
            {
                const int ra = s_pairs[p].x, rb = s_pairs[p].y;
                double* Da = D + (int64_t)ra * m;
                const double* Db = D + (int64_t)rb * m;
                for (int k = lane; k < m; k += 32)
                {
                    const double x = Da[k], y = Db[k];
                    Da[k] = (x < y) ? y : x;
                }
            }
            __syncthreads();
            /* columns: D[k][a] = max(D[k][a], D[k][b]) for every active row k; rows whose nearest neighbour merged are marked */
            for (int row = warp; row < m; row += warps)
            {
                if (!active[row]) continue;
                double* Dk = D + (int64_t)row * m;
                const int my_nn = nn[row];
                bool touched = false;
                for (int p = lane; p < npairs; p += 32)
                {
                    const int ca = s_pairs[p].x, cb = s_pairs[p].y;
                    if (ca != row)
                    {
                        const double x = Dk[ca], y = Dk[cb];
                        Dk[ca] = (x < y) ? y : x;
                    }
                    touched |= (my_nn == ca) | (my_nn == cb) | (row == ca);
                }
                if (__any_sync(0xffffffffu, touched) && lane == 0) s_dirty[row] = 1;
            }
            __syncthreads();
            for (int p = threadIdx.x; p < npairs; p += blockDim.x)
            {
                const int rb = s_pairs[p].y;
                active[rb] = 0;
                into[rb] = s_pairs[p].x;
            }
            __syncthreads();
        }
    }
}

/* connected components of the threshold graph by min-label propagation + pointer jumping (exact path only) */
__global__ void __launch_bounds__(256) propagate_kernel(const uint32_t* __restrict__ adj, int64_t total, int64_t set_size, int wpr,
                                                        const int32_t* __restrict__ set_bad, int32_t* __restrict__ comp,
                                                        int* __restrict__ changed)
{
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= total) return;
    const int64_t set = p / set_size;
    if (!set_bad[set]) return;
    const uint32_t* row = adj + p * wpr;
    int32_t* sc = comp + set * set_size;
    int32_t best = sc[p - set * set_size];
    for (int w = lane; w < wpr; w += 32)
    {
        uint32_t v = row[w];
        while (v)
        {
            const int b = __ffs((int)v) - 1;
            v &= v - 1u;
            best = min(best, sc[w * 32 + b]);
        }
    }
    for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0)
    {
        int32_t rr = best;
        while (sc[rr] < rr) rr = sc[rr];
        if (rr < comp[p])
        {
            comp[p] = rr;
            *changed = 1;
        }
    }
}

/* the same propagation for sets of at most kFusedPropagateMax particles, iterated to its fixed point inside one CTA per set (no
 * host round trip per sweep): the fixed point - every particle labelled with the smallest index of its component - does not depend
 * on the order in which the labels improve */
constexpr int64_t kFusedPropagateMax = 8192;

__global__ void __launch_bounds__(1024) propagate_fused_kernel(const uint32_t* __restrict__ adj, int64_t set_size, int wpr,
                                                               const int32_t* __restrict__ set_bad, int32_t* __restrict__ comp)
{
    /* the labels of the set live in shared memory while they are iterated (the sweeps chase them through dependent loads: ~ 30
     * cycles each here against an L2 round trip per hop in global memory) and are written back once */
    extern __shared__ int32_t s_labels[];
    __shared__ int s_changed;
    const int64_t set = blockIdx.x;
    if (!set_bad[set]) return;
    int32_t* gc = comp + set * set_size;
    volatile int32_t* sc = s_labels;
    for (int64_t i = threadIdx.x; i < set_size; i += blockDim.x) s_labels[i] = gc[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    while (true)
    {
        __syncthreads();
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        for (int64_t i = warp; i < set_size; i += warps)
        {
            const uint32_t* row = adj + (set * set_size + i) * wpr;
            int32_t best = sc[i];
            for (int w = lane; w < wpr; w += 32)
            {
                uint32_t v = row[w];
                while (v)
                {
                    const int b = __ffs((int)v) - 1;
                    v &= v - 1u;
                    best = min(best, sc[w * 32 + b]);
                }
            }
            for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
            if (lane == 0)
            {
                int32_t rr = best;
                while (sc[rr] < rr) rr = sc[rr];
                if (rr < sc[i])
                {
                    sc[i] = rr;
                    s_changed = 1;
                }
            }
        }
        __syncthreads();
        if (!s_changed) break;
    }
    for (int64_t i = threadIdx.x; i < set_size; i += blockDim.x) gc[i] = s_labels[i];
}

/* ------------------------------------------------------------------ per-cluster statistics */

struct StatsArgs
{
    DevRobot robot;
    int64_t n;
    const double* cfg; /* [C][n] */
    const int32_t* labels;
    const uint8_t* collided;
    int allow_contacts;
    double step_size;
    int64_t* counts;
    uint8_t* any_collided;
    double* expectations; /* [K][C] */
    double* variance;
    double* si_variance;
    double* variances;    /* [K][V] */
    double* si_variances;
};

__device__ __forceinline__ void quat_to_rotation(const double* q, double* T)
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    T[0] = 1.0 - 2.0 * ((y * y) + (z * z));
    T[1] = 2.0 * ((x * y) - (z * w));
    T[2] = 2.0 * ((x * z) + (y * w));
    T[4] = 2.0 * ((x * y) + (z * w));
    T[5] = 1.0 - 2.0 * ((x * x) + (z * z));
    T[6] = 2.0 * ((y * z) - (x * w));
    T[8] = 2.0 * ((x * z) - (y * w));
    T[9] = 2.0 * ((y * z) + (x * w));
    T[10] = 1.0 - 2.0 * ((x * x) + (y * y));
}

template <int KIND>
__device__ __forceinline__ void per_dimension_distance(const DevRobot& r, const double* c1, const double* c2, double* out)
{
    if (KIND == UPC_ROBOT_SE2)
    {
        out[0] = fabs(c2[0] - c1[0]);
        out[1] = fabs(c2[1] - c1[1]);
        out[2] = fabs(upc_angle_signed_distance(c1[2], c2[2]));
    }
    else if (KIND == UPC_ROBOT_SE3)
    {
        double t1[12], t2[12];
        for (int rr = 0; rr < 3; rr++)
        {
            for (int k = 0; k < 3; k++) { t1[rr * 4 + k] = c1[rr * 3 + k]; t2[rr * 4 + k] = c2[rr * 3 + k]; }
            t1[rr * 4 + 3] = c1[9 + rr]; t2[rr * 4 + 3] = c2[9 + rr];
        }
        out[0] = fabs(t2[3] - t1[3]);
        out[1] = fabs(t2[7] - t1[7]);
        out[2] = fabs(t2[11] - t1[11]);
        double q1[4], q2[4];
        quat_from_xf(t1, q1);
        quat_from_xf(t2, q2);
        const double dq = fabs((((q1[0] * q2[0]) + (q1[1] * q2[1])) + (q1[2] * q2[2])) + (q1[3] * q2[3]));
        out[3] = (dq < (1.0 - 2.220446049250313e-16)) ? fabs(upc_acos(((2.0 * dq) * dq) - 1.0)) : 0.0;
    }
    else
    {
        for (int a = 0; a < r.dof; a++)
        {
            const double raw = r.axis_continuous[a] ? upc_angle_signed_distance(c1[a], c2[a]) : (c2[a] - c1[a]);
            out[a] = fabs(raw * r.weights[a]);
        }
    }
}

/*
 * One warp per cluster.  Members are visited in ascending particle order (32 candidates per ballot, then the set
 * bits in order), every lane accumulating the same sequential sums, so the results are those of the serial loops of
 * uncertainty_planner_state.hpp:591-714 bit for bit.
 */
template <int KIND>
__global__ void __launch_bounds__(32) cluster_stats_kernel(const __grid_constant__ StatsArgs a)
{
    constexpr int W = (KIND == UPC_ROBOT_SE2) ? 3 : ((KIND == UPC_ROBOT_SE3) ? 12 : UPC_MAX_DOF);
    constexpr int VW = (KIND == UPC_ROBOT_SE2) ? 3 : ((KIND == UPC_ROBOT_SE3) ? 4 : UPC_MAX_DOF);
    const int k = blockIdx.x;
    const int lane = threadIdx.x;
    const int width = (KIND == UPC_ROBOT_LINKED) ? a.robot.dof : W;
    const int vw = (KIND == UPC_ROBOT_LINKED) ? a.robot.dof : VW;
    /* pass 1: count and sums */
    int64_t count = 0;
    bool any = false;
    double acc[2 * UPC_MAX_DOF]; /* SE2: sx sy ss sc; SE3: st[3] sq[4]; linked: per joint (sum | sin, cos) */
    for (int j = 0; j < 2 * UPC_MAX_DOF; j++) acc[j] = 0.0;
    double qref[4] = {0.0, 0.0, 0.0, 1.0};
    double first_cfg[W];
    for (int c = 0; c < W; c++) first_cfg[c] = 0.0;
    for (int64_t base = 0; base < a.n; base += 32)
    {
        const int64_t i = base + lane;
        const bool member = (i < a.n) && (a.labels[i] == k) && ((a.collided[i] == 0) || a.allow_contacts);
        unsigned todo = __ballot_sync(0xffffffffu, member);
        while (todo)
        {
            const int src = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            const int64_t p = base + src;
            double c[W];
            for (int x = 0; x < width; x++) c[x] = a.cfg[(int64_t)x * a.n + p];
            if (count == 0)
                for (int x = 0; x < width; x++) first_cfg[x] = c[x];
            if (a.collided[p]) any = true;
            if (KIND == UPC_ROBOT_SE2)
            {
                acc[0] = acc[0] + c[0];
                acc[1] = acc[1] + c[1];
                double sn, cs;
                upc_sincos(c[2], &sn, &cs);
                acc[2] = acc[2] + sn;
                acc[3] = acc[3] + cs;
            }
            else if (KIND == UPC_ROBOT_SE3)
            {
                double T[12], q[4];
                for (int rr = 0; rr < 3; rr++)
                {
                    for (int x = 0; x < 3; x++) T[rr * 4 + x] = c[rr * 3 + x];
                    T[rr * 4 + 3] = c[9 + rr];
                }
                acc[0] = acc[0] + T[3];
                acc[1] = acc[1] + T[7];
                acc[2] = acc[2] + T[11];
                quat_from_xf(T, q);
                if (count == 0)
                    for (int x = 0; x < 4; x++) qref[x] = q[x];
                const double dot = (((q[0] * qref[0]) + (q[1] * qref[1])) + (q[2] * qref[2])) + (q[3] * qref[3]);
                const double sign = (dot < 0.0) ? -1.0 : 1.0;
                for (int x = 0; x < 4; x++) acc[3 + x] = acc[3 + x] + (sign * q[x]);
            }
            else
            {
                for (int x = 0; x < width; x++)
                {
                    if (a.robot.axis_continuous[x])
                    {
                        double sn, cs;
                        upc_sincos(c[x], &sn, &cs);
                        acc[2 * x] = acc[2 * x] + sn;
                        acc[2 * x + 1] = acc[2 * x + 1] + cs;
                    }
                    else
                    {
                        acc[2 * x] = acc[2 * x] + c[x];
                    }
                }
            }
            count++;
        }
    }
    /* expectation (ComputeExpectation: a single particle is returned as is) */
    double e[W];
    for (int c = 0; c < W; c++) e[c] = 0.0;
    const double nd = (double)count;
    if (count == 1)
    {
        for (int c = 0; c < width; c++) e[c] = first_cfg[c];
    }
    else if (count > 1)
    {
        if (KIND == UPC_ROBOT_SE2)
        {
            e[0] = acc[0] / nd;
            e[1] = acc[1] / nd;
            e[2] = upc_atan2(acc[2], acc[3]);
        }
        else if (KIND == UPC_ROBOT_SE3)
        {
            const double norm = sqrt((((acc[3] * acc[3]) + (acc[4] * acc[4])) + (acc[5] * acc[5])) + (acc[6] * acc[6]));
            double q[4], T[12];
            for (int x = 0; x < 4; x++) q[x] = acc[3 + x] / norm;
            quat_to_rotation(q, T);
            for (int rr = 0; rr < 3; rr++)
                for (int x = 0; x < 3; x++) e[rr * 3 + x] = T[rr * 4 + x];
            e[9] = acc[0] / nd;
            e[10] = acc[1] / nd;
            e[11] = acc[2] / nd;
        }
        else
        {
            for (int x = 0; x < width; x++) e[x] = a.robot.axis_continuous[x] ? upc_atan2(acc[2 * x], acc[2 * x + 1]) : (acc[2 * x] / nd);
        }
    }
    /* pass 2: variances */
    double var_sum = 0.0, si_sum = 0.0;
    double dv[VW], dsi[VW];
    for (int v = 0; v < VW; v++) { dv[v] = 0.0; dsi[v] = 0.0; }
    if (count == 1)
    {
        double dd[VW];
        per_dimension_distance<KIND>(a.robot, first_cfg, first_cfg, dd);
        for (int v = 0; v < vw; v++) { dv[v] = dd[v]; dsi[v] = dd[v]; }
    }
    else if (count > 1)
    {
        const double weight = 1.0 / nd;
        for (int64_t base = 0; base < a.n; base += 32)
        {
            const int64_t i = base + lane;
            const bool member = (i < a.n) && (a.labels[i] == k) && ((a.collided[i] == 0) || a.allow_contacts);
            unsigned todo = __ballot_sync(0xffffffffu, member);
            while (todo)
            {
                const int src = __ffs((int)todo) - 1;
                todo &= todo - 1u;
                const int64_t p = base + src;
                double c[W];
                for (int x = 0; x < width; x++) c[x] = a.cfg[(int64_t)x * a.n + p];
                double raw;
                if (KIND == UPC_ROBOT_SE2) raw = Se2Model::distance(a.robot, e, c);
                else if (KIND == UPC_ROBOT_SE3) raw = Se3Model::distance(a.robot, e, c);
                else raw = LinkedModel::distance(a.robot, e, c);
                var_sum = var_sum + ((raw * raw) * weight);
                const double si = raw / a.step_size;
                si_sum = si_sum + ((si * si) * weight);
                double dd[VW];
                per_dimension_distance<KIND>(a.robot, e, c, dd);
                for (int v = 0; v < vw; v++)
                {
                    dv[v] = dv[v] + ((dd[v] * dd[v]) * weight);
                    const double sd = dd[v] / a.step_size;
                    dsi[v] = dsi[v] + ((sd * sd) * weight);
                }
            }
        }
    }
    if (lane == 0)
    {
        a.counts[k] = count;
        a.any_collided[k] = any ? 1 : 0;
        for (int c = 0; c < width; c++) a.expectations[(int64_t)k * width + c] = e[c];
        a.variance[k] = var_sum;
        a.si_variance[k] = si_sum;
        for (int v = 0; v < vw; v++)
        {
            a.variances[(int64_t)k * vw + v] = dv[v];
            a.si_variances[(int64_t)k * vw + v] = dsi[v];
        }
    }
}

cudaError_t launch_cluster_stats(upc_handle* h, int64_t n, const double* cfg, const int32_t* labels, const uint8_t* collided, int allow_contacts,
                                 double step_size, int n_clusters, int64_t* counts, uint8_t* any_collided, double* expectations, double* variance,
                                 double* si_variance, double* variances, double* si_variances, cudaStream_t stream)
{
    StatsArgs a;
    a.robot = h->robot; a.n = n; a.cfg = cfg; a.labels = labels; a.collided = collided; a.allow_contacts = allow_contacts; a.step_size = step_size;
    a.counts = counts; a.any_collided = any_collided; a.expectations = expectations; a.variance = variance; a.si_variance = si_variance;
    a.variances = variances; a.si_variances = si_variances;
    if (h->robot.kind == UPC_ROBOT_SE2) cluster_stats_kernel<UPC_ROBOT_SE2><<<(unsigned)n_clusters, 32, 0, stream>>>(a);
    else if (h->robot.kind == UPC_ROBOT_SE3) cluster_stats_kernel<UPC_ROBOT_SE3><<<(unsigned)n_clusters, 32, 0, stream>>>(a);
    else cluster_stats_kernel<UPC_ROBOT_LINKED><<<(unsigned)n_clusters, 32, 0, stream>>>(a);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ compaction of the sets the pivot pass left unsettled */

/* features (and first-pass groups) of the listed sets, packed as a batch of n_list sets */
__global__ void __launch_bounds__(256) gather_sets_kernel(const double* __restrict__ feat, int nfeat, int64_t total, int64_t set_size,
                                                          const int32_t* __restrict__ set_list, int64_t n_list, double* __restrict__ cfeat,
                                                          const int32_t* __restrict__ group, int32_t* __restrict__ cgroup)
{
    const int64_t ctotal = n_list * set_size;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ctotal) return;
    const int64_t k = p / set_size;
    const int64_t src = (int64_t)set_list[k] * set_size + (p - k * set_size);
    for (int f = 0; f < nfeat; f++) cfeat[(int64_t)f * ctotal + p] = feat[(int64_t)f * total + src];
    if (group) cgroup[p] = group[src];
}

__global__ void __launch_bounds__(256) scatter_roots_kernel(const int32_t* __restrict__ croot, const int32_t* __restrict__ set_list, int64_t n_list,
                                                            int64_t set_size, int32_t* __restrict__ root)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_list * set_size) return;
    const int64_t k = p / set_size;
    root[(int64_t)set_list[k] * set_size + (p - k * set_size)] = croot[p];
}

/* ------------------------------------------------------------------ host orchestration */

struct PassInput
{
    const double* feat = nullptr;
    int nfeat = 0;
    const uint32_t* sig = nullptr;
    int max_mismatch = 0;
    const int32_t* group = nullptr;
    double threshold = 0.0;
    const int32_t* set_count = nullptr; /* valid items per set when the sets are padded to a common size */
    bool allow_exact = true;            /* false: only report whether every set is a union of cliques */
    bool binary = false;                /* the adjacency bits were already written by the caller (AC / PTPM) */
};

template <typename T>
static T* buf(DeviceBuffer& b, size_t count)
{
    if (b.reserve(count * sizeof(T)) != cudaSuccess) return nullptr;
    return (T*)b.ptr;
}

#define UPC_NULLCHECK(p)                            \
    if (!(p))                                       \
    {                                               \
        set_error("device allocation failed");      \
        return UPC_ERR_CUDA;                        \
    }

/* one clustering pass: adjacency -> roots (+ exact fallback). roots[p] = representative index within the set. */
static int32_t cluster_pass(upc_handle* h, const PassInput& in, int64_t n_sets, int64_t set_size, int32_t* d_root, cudaStream_t stream,
                            int64_t* fallback_sets)
{
    const int64_t total = n_sets * set_size;
    const int wpr = (int)((set_size + 31) / 32);
    uint32_t* adj = buf<uint32_t>(h->cw.adjacency, (size_t)total * wpr);
    UPC_NULLCHECK(adj);
    uint8_t* row_ok = buf<uint8_t>(h->cw.flags, (size_t)total + (size_t)n_sets * sizeof(int32_t) + 64);
    UPC_NULLCHECK(row_ok);
    int32_t* set_bad = (int32_t*)(row_ok + ((total + 15) / 16) * 16);
    UPC_CUDA_TRY(cudaMemsetAsync(set_bad, 0, (size_t)n_sets * sizeof(int32_t), stream));
    const int kind = h->robot.kind;
    if (in.binary)
    {
        /* adjacency already in place */
    }
    else if (in.sig)
    {
        SigAdjArgs sa;
        sa.sig = in.sig; sa.num_points = h->robot.num_points; sa.total = total; sa.set_size = set_size; sa.wpr = wpr; sa.adj = adj;
        sa.max_mismatch = in.max_mismatch;
        sa.set_count = in.set_count;
        dim3 grid((unsigned)((set_size + 63) / 64), (unsigned)((set_size + 63) / 64), (unsigned)n_sets);
        signature_adjacency_kernel<<<grid, 256, 0, stream>>>(sa);
    }
    else
    {
        AdjArgs aa;
        aa.robot = h->robot; aa.feat = in.feat; aa.nfeat = in.nfeat; aa.total = total; aa.set_size = set_size; aa.wpr = wpr;
        aa.group = in.group; aa.adj = adj; aa.threshold = in.threshold;
        dim3 grid((unsigned)((set_size + kTileCols - 1) / kTileCols), (unsigned)((set_size + kTileRows - 1) / kTileRows), (unsigned)n_sets);
        if (kind == UPC_ROBOT_SE2) adjacency_kernel<UPC_ROBOT_SE2, 0><<<grid, kTileCols, 0, stream>>>(aa);
        else if (kind == UPC_ROBOT_SE3) adjacency_kernel<UPC_ROBOT_SE3, 0><<<grid, kTileCols, 0, stream>>>(aa);
        else
        {
            switch (in.nfeat)
            {
                case 1: adjacency_kernel<UPC_ROBOT_LINKED, 1><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 2: adjacency_kernel<UPC_ROBOT_LINKED, 2><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 3: adjacency_kernel<UPC_ROBOT_LINKED, 3><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 4: adjacency_kernel<UPC_ROBOT_LINKED, 4><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 5: adjacency_kernel<UPC_ROBOT_LINKED, 5><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 6: adjacency_kernel<UPC_ROBOT_LINKED, 6><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 7: adjacency_kernel<UPC_ROBOT_LINKED, 7><<<grid, kTileCols, 0, stream>>>(aa); break;
                case 8: adjacency_kernel<UPC_ROBOT_LINKED, 8><<<grid, kTileCols, 0, stream>>>(aa); break;
                default: adjacency_kernel<UPC_ROBOT_LINKED, 0><<<grid, kTileCols, 0, stream>>>(aa); break;
            }
        }
        {
            const int64_t warps = ((set_size + 31) / 32) * (int64_t)wpr * n_sets;
            mirror_lower_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(adj, set_size, wpr, n_sets);
            h->stats.kernel_launches++;
        }
    }
    if (!in.binary) h->stats.kernel_launches++;
    UPC_CUDA_TRY(cudaGetLastError());
    {
        const int64_t threads = total * 32;
        root_check_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(adj, total, set_size, wpr, d_root, row_ok, set_bad);
        h->stats.kernel_launches++;
        UPC_CUDA_TRY(cudaGetLastError());
    }
    /* did every set come out as a union of cliques? */
    std::vector<int32_t> bad((size_t)n_sets);
    UPC_CUDA_TRY(cudaMemcpyAsync(bad.data(), set_bad, (size_t)n_sets * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(stream));
    h->stats.d2h_bytes += n_sets * (int64_t)sizeof(int32_t);
    int64_t nbad = 0;
    for (int32_t b : bad) nbad += b ? 1 : 0;
    if (fallback_sets) *fallback_sets = nbad;
    if (nbad == 0 || !in.allow_exact) return UPC_OK;

    /* ---- exact path: the connected components that are not cliques ---- */
    int32_t* comp = buf<int32_t>(h->cw.comp, (size_t)total);
    UPC_NULLCHECK(comp);
    iota_in_set_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(comp, total, set_size);
    int* d_flag = (int*)buf<int>(h->cw.misc, 16);
    UPC_NULLCHECK(d_flag);
    if (set_size <= kFusedPropagateMax)
    {
        propagate_fused_kernel<<<(unsigned)n_sets, 1024, (size_t)set_size * sizeof(int32_t), stream>>>(adj, set_size, wpr, set_bad, comp);
        h->stats.kernel_launches++;
        UPC_CUDA_TRY(cudaGetLastError());
    }
    for (int iter = 0; set_size > kFusedPropagateMax && iter < 1 << 20; iter++)
    {
        UPC_CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
        propagate_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, stream>>>(adj, total, set_size, wpr, set_bad, comp, d_flag);
        h->stats.kernel_launches++;
        int changed = 0;
        UPC_CUDA_TRY(cudaMemcpyAsync(&changed, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
        UPC_CUDA_TRY(cudaStreamSynchronize(stream));
        if (!changed) break;
    }
    std::vector<int32_t> h_comp((size_t)total);
    std::vector<uint8_t> h_ok((size_t)total);
    UPC_CUDA_TRY(cudaMemcpyAsync(h_comp.data(), comp, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(h_ok.data(), row_ok, (size_t)total, cudaMemcpyDeviceToHost, stream));
    UPC_CUDA_TRY(cudaStreamSynchronize(stream));
    h->stats.d2h_bytes += total * 5;
    /* components containing a row that differs from its root's row, members in ascending order */
    std::vector<int64_t> members, comp_off(1, 0), matrix_off(1, 0);
    std::vector<int32_t> comp_of_row;
    for (int64_t s = 0; s < n_sets; s++)
    {
        if (!bad[(size_t)s]) continue;
        const int64_t base = s * set_size;
        std::vector<uint8_t> dirty((size_t)set_size, 0);
        std::vector<int32_t> count((size_t)set_size, 0);
        for (int64_t i = 0; i < set_size; i++)
        {
            const int32_t c = h_comp[(size_t)(base + i)];
            count[(size_t)c]++;
            if (!h_ok[(size_t)(base + i)]) dirty[(size_t)c] = 1;
        }
        std::vector<int64_t> slot((size_t)set_size, -1);
        for (int64_t c = 0; c < set_size; c++)
        {
            if (!dirty[(size_t)c]) continue;
            slot[(size_t)c] = (int64_t)members.size();
            members.resize(members.size() + (size_t)count[(size_t)c]);
            comp_of_row.insert(comp_of_row.end(), (size_t)count[(size_t)c], (int32_t)(comp_off.size() - 1));
            comp_off.push_back((int64_t)members.size());
            matrix_off.push_back(matrix_off.back() + (int64_t)count[(size_t)c] * (int64_t)count[(size_t)c]);
        }
        std::vector<int64_t> fill(slot);
        for (int64_t i = 0; i < set_size; i++)
        {
            const int32_t c = h_comp[(size_t)(base + i)];
            if (slot[(size_t)c] >= 0) members[(size_t)(fill[(size_t)c]++)] = base + i;
        }
    }
    const int ncomp = (int)comp_off.size() - 1;
    if (ncomp == 0) return UPC_OK;
    const int64_t rows = (int64_t)members.size();
    if ((double)matrix_off.back() * 8.0 > 100.0e9)
    {
        set_error("exact complete-link path: the non-clique components need more than 100 GB of distance matrices");
        return UPC_ERR_UNSUPPORTED;
    }
    ExactArgs ea;
    ea.robot = h->robot; ea.feat = in.feat; ea.sig = in.sig; ea.group = in.group; ea.nfeat = in.nfeat; ea.num_points = h->robot.num_points;
    ea.binary_adj = in.binary ? adj : nullptr; ea.wpr = wpr;
    ea.total = total; ea.set_size = set_size; ea.rows = rows; ea.threshold = in.threshold;
    double* d_mat = buf<double>(h->cw.dist_matrix, (size_t)matrix_off.back());
    UPC_NULLCHECK(d_mat);
    ea.D = d_mat;
    /* [members i64 rows | comp_off i64 ncomp+1 | matrix_off i64 ncomp+1] */
    int64_t* d_members = buf<int64_t>(h->cw.members, (size_t)rows + 2 * (size_t)(ncomp + 1));
    UPC_NULLCHECK(d_members);
    int64_t* d_comp_off = d_members + rows;
    int64_t* d_matrix_off = d_comp_off + (ncomp + 1);
    UPC_CUDA_TRY(cudaMemcpyAsync(d_members, members.data(), (size_t)rows * 8, cudaMemcpyHostToDevice, stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(d_comp_off, comp_off.data(), (size_t)(ncomp + 1) * 8, cudaMemcpyHostToDevice, stream));
    UPC_CUDA_TRY(cudaMemcpyAsync(d_matrix_off, matrix_off.data(), (size_t)(ncomp + 1) * 8, cudaMemcpyHostToDevice, stream));
    /* [comp_of_row, nn, partner, into i32 rows each | active u8 rows] */
    uint8_t* scratch = buf<uint8_t>(h->cw.link_scratch, (size_t)rows * 17 + 64);
    UPC_NULLCHECK(scratch);
    int32_t* d_comp_of_row = (int32_t*)scratch;
    UPC_CUDA_TRY(cudaMemcpyAsync(d_comp_of_row, comp_of_row.data(), (size_t)rows * 4, cudaMemcpyHostToDevice, stream));
    h->stats.h2d_bytes += rows * 12 + (ncomp + 1) * 16;
    ea.members = d_members; ea.comp_off = d_comp_off; ea.matrix_off = d_matrix_off; ea.comp_of_row = d_comp_of_row;
    ea.nn = d_comp_of_row + rows;
    ea.partner = ea.nn + rows;
    ea.into = ea.partner + rows;
    ea.active = (uint8_t*)(ea.into + rows);
    ea.merged_flag = d_flag;
    const unsigned cta_rows = (unsigned)std::min<int64_t>(rows, 1 << 20);
    if (kind == UPC_ROBOT_SE2) exact_matrix_kernel<UPC_ROBOT_SE2><<<cta_rows, 256, 0, stream>>>(ea);
    else if (kind == UPC_ROBOT_SE3) exact_matrix_kernel<UPC_ROBOT_SE3><<<cta_rows, 256, 0, stream>>>(ea);
    else exact_matrix_kernel<UPC_ROBOT_LINKED><<<cta_rows, 256, 0, stream>>>(ea);
    h->stats.kernel_launches++;
    UPC_CUDA_TRY(cudaGetLastError());
    const unsigned row_blocks = (unsigned)((rows + 255) / 256);
    const unsigned warp_blocks = (unsigned)((rows * 32 + 255) / 256);
    int64_t largest = 0;
    for (int c = 0; c < ncomp; c++) largest = std::max(largest, comp_off[(size_t)c + 1] - comp_off[(size_t)c]);
    {
        /* components of at most kFusedExactMax members: all rounds inside one CTA each, no host round trips */
        exact_rounds_fused_kernel<<<(unsigned)std::min(ncomp, 65535), 1024, 0, stream>>>(ea, ncomp);
        h->stats.kernel_launches++;
        UPC_CUDA_TRY(cudaGetLastError());
    }
    for (int round = 0; largest > kFusedExactMax && round < 1 << 20; round++)
    {
        UPC_CUDA_TRY(cudaMemsetAsync(ea.merged_flag, 0, sizeof(int), stream));
        exact_nearest_kernel<<<warp_blocks, 256, 0, stream>>>(ea);
        exact_mark_kernel<<<row_blocks, 256, 0, stream>>>(ea);
        h->stats.kernel_launches += 2;
        int merged = 0;
        UPC_CUDA_TRY(cudaMemcpyAsync(&merged, ea.merged_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
        UPC_CUDA_TRY(cudaStreamSynchronize(stream));
        if (!merged) break;
        exact_row_merge_kernel<<<cta_rows, 256, 0, stream>>>(ea);
        exact_col_merge_kernel<<<cta_rows, 256, 0, stream>>>(ea);
        exact_retire_kernel<<<row_blocks, 256, 0, stream>>>(ea);
        h->stats.kernel_launches += 3;
        UPC_CUDA_TRY(cudaGetLastError());
    }
    exact_roots_kernel<<<row_blocks, 256, 0, stream>>>(ea, d_root);
    h->stats.kernel_launches++;
    UPC_CUDA_TRY(cudaGetLastError());
    return UPC_OK;
}

template <class M>
static cudaError_t launch_signatures(upc_handle* h, const double* cfg, int64_t total, uint32_t* sig, unsigned long long* hash, double* frames,
                                     cudaStream_t stream)
{
    SigArgs sa;
    sa.env = h->env; sa.robot = h->robot; sa.cfg = cfg; sa.total = total; sa.sig = sig; sa.hash = hash; sa.frames = frames;
    signature_kernel<M><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(sa);
    return cudaGetLastError();
}

int32_t run_cluster(upc_handle* h, const upc_cluster_params& cp, int64_t n_sets, int64_t set_size, const double* d_configs,
                    int32_t* d_labels, int32_t* d_n_clusters, cudaStream_t stream)
{
    const int64_t total = n_sets * set_size;
    if (total == 0) return UPC_OK;
    const int kind = h->robot.kind;
    /* distance features */
    const double* feat = d_configs;
    int nfeat = (kind == UPC_ROBOT_SE2) ? 3 : h->robot.dof;
    if (kind == UPC_ROBOT_SE3)
    {
        double* f = buf<double>(h->cw.features, (size_t)total * 7);
        UPC_NULLCHECK(f);
        se3_features_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_configs, f, total);
        h->stats.kernel_launches++;
        UPC_CUDA_TRY(cudaGetLastError());
        feat = f;
        nfeat = 7;
    }
    /* [first-pass roots | final roots | rank scratch] */
    int32_t* roots = buf<int32_t>(h->cw.roots, (size_t)total * 3);
    UPC_NULLCHECK(roots);
    int32_t* root1 = roots;
    int32_t* root2 = roots + total;
    int32_t* rank = roots + 2 * total;
    const int32_t* group = nullptr;
    if (set_size > 1) h->stats.cluster_calls += n_sets;
    if (cp.feature_type == UPC_FEATURE_CONVEX_REGION_SIGNATURE)
    {
        if (!h->d_seg)
        {
            set_error("convex_segment grid required for CRS clustering");
            return UPC_ERR_INVALID_ARGUMENT;
        }
        const int P = h->robot.num_points;
        uint32_t* sig = buf<uint32_t>(h->cw.signatures, (size_t)total * P);
        UPC_NULLCHECK(sig);
        double* frames = nullptr;
        if (kind == UPC_ROBOT_LINKED)
        {
            frames = buf<double>(h->cw.frames, (size_t)total * h->robot.num_links * 12);
            UPC_NULLCHECK(frames);
        }
        unsigned long long* hash = buf<unsigned long long>(h->cw.hash, (size_t)total);
        UPC_NULLCHECK(hash);
        cudaError_t e;
        if (kind == UPC_ROBOT_SE2) e = launch_signatures<Se2Model>(h, d_configs, total, sig, hash, frames, stream);
        else if (kind == UPC_ROBOT_SE3) e = launch_signatures<Se3Model>(h, d_configs, total, sig, hash, frames, stream);
        else e = launch_signatures<LinkedModel>(h, d_configs, total, sig, hash, frames, stream);
        h->stats.kernel_launches++;
        UPC_CUDA_TRY(e);
        PassInput in;
        in.threshold = cp.signature_matching_threshold;
        int mm = -1;
        for (int m = 0; m <= P; m++)
        {
            if (((double)m / (double)P) <= cp.signature_matching_threshold) mm = m;
            else break;
        }
        in.max_mismatch = mm;
        bool done = false;
        if (mm >= 0 && set_size >= 64)
        {
            /* ---- fast path: cluster the distinct signature vectors only (exact whenever their threshold graph is a
             * union of cliques, which is then also true of the full graph; otherwise fall through to the full pass) ---- */
            int table = 64;
            while (table < 2 * set_size) table *= 2;
            const size_t tcells = (size_t)n_sets * (size_t)table;
            /* [keys u64 | reps i32 | urep | uid | ulist (i32 each, total) | ucount (n_sets)] */
            const size_t bytes = tcells * 12 + (size_t)total * 12 + (size_t)n_sets * 4 + 64;
            uint8_t* base = buf<uint8_t>(h->cw.dedupe, bytes);
            UPC_NULLCHECK(base);
            DedupeArgs da;
            da.sig = sig; da.hash = hash; da.num_points = P; da.total = total; da.set_size = set_size; da.table_size = table;
            da.keys = (unsigned long long*)base;
            da.reps = (int32_t*)(base + tcells * 8);
            da.urep = da.reps + tcells;
            da.uid = da.urep + total;
            da.ulist = da.uid + total;
            da.ucount = da.ulist + total;
            UPC_CUDA_TRY(cudaMemsetAsync(da.keys, 0xFF, tcells * 8, stream));
            UPC_CUDA_TRY(cudaMemsetAsync(da.reps, 0x7F, tcells * 4, stream));
            UPC_CUDA_TRY(cudaMemsetAsync(da.ucount, 0, (size_t)n_sets * 4, stream));
            const unsigned blocks = (unsigned)((total + 255) / 256);
            dedupe_insert_kernel<<<blocks, 256, 0, stream>>>(da);
            dedupe_resolve_kernel<<<blocks, 256, 0, stream>>>(da);
            dedupe_uid_kernel<<<blocks, 256, 0, stream>>>(da);
            h->stats.kernel_launches += 3;
            UPC_CUDA_TRY(cudaGetLastError());
            std::vector<int32_t> ucount((size_t)n_sets);
            UPC_CUDA_TRY(cudaMemcpyAsync(ucount.data(), da.ucount, (size_t)n_sets * 4, cudaMemcpyDeviceToHost, stream));
            UPC_CUDA_TRY(cudaStreamSynchronize(stream));
            h->stats.d2h_bytes += n_sets * 4;
            int umax = 1;
            for (int32_t c : ucount) umax = std::max(umax, (int)c);
            if ((int64_t)umax * 2 <= set_size)
            {
                const int64_t utotal = n_sets * (int64_t)umax;
                uint32_t* usig = buf<uint32_t>(h->cw.usig, (size_t)utotal * P);
                UPC_NULLCHECK(usig);
                int32_t* uroot = buf<int32_t>(h->cw.uroot, (size_t)utotal);
                UPC_NULLCHECK(uroot);
                dedupe_gather_kernel<<<(unsigned)((utotal * P + 255) / 256), 256, 0, stream>>>(sig, P, total, set_size, da.ucount, da.ulist, umax, usig);
                h->stats.kernel_launches++;
                UPC_CUDA_TRY(cudaGetLastError());
                PassInput uin = in;
                uin.sig = usig;
                uin.set_count = da.ucount;
                uin.allow_exact = false;
                int64_t ubad = 0;
                const int32_t urc = cluster_pass(h, uin, n_sets, umax, uroot, stream, &ubad);
                if (urc != UPC_OK) return urc;
                if (ubad == 0)
                {
                    dedupe_expand_kernel<<<blocks, 256, 0, stream>>>(uroot, da.uid, total, set_size, umax, root1);
                    h->stats.kernel_launches++;
                    UPC_CUDA_TRY(cudaGetLastError());
                    done = true;
                }
            }
        }
        if (!done)
        {
            in.sig = sig;
            const int32_t rc = cluster_pass(h, in, n_sets, set_size, root1, stream, nullptr);
            if (rc != UPC_OK) return rc;
        }
        group = root1;
    }
    else if (cp.feature_type == UPC_FEATURE_ACTUATION_CENTER_CONNECTIVITY)
    {
        if (!h->d_occ)
        {
            set_error("occupancy grid required for actuation-centre clustering");
            return UPC_ERR_INVALID_ARGUMENT;
        }
        const int L = h->robot.num_links;
        double* centers = buf<double>(h->cw.usig, (size_t)total * L * 3);
        UPC_NULLCHECK(centers);
        double* frames = nullptr;
        if (kind == UPC_ROBOT_LINKED)
        {
            frames = buf<double>(h->cw.frames, (size_t)total * L * 12);
            UPC_NULLCHECK(frames);
        }
        CenterArgs ca;
        ca.robot = h->robot; ca.cfg = d_configs; ca.total = total; ca.centers = centers; ca.frames = frames;
        const unsigned cblocks = (unsigned)((total + 127) / 128);
        if (kind == UPC_ROBOT_SE2) link_centers_kernel<Se2Model><<<cblocks, 128, 0, stream>>>(ca);
        else if (kind == UPC_ROBOT_SE3) link_centers_kernel<Se3Model><<<cblocks, 128, 0, stream>>>(ca);
        else link_centers_kernel<LinkedModel><<<cblocks, 128, 0, stream>>>(ca);
        const int wpr = (int)((set_size + 31) / 32);
        uint32_t* adj = buf<uint32_t>(h->cw.adjacency, (size_t)total * wpr);
        UPC_NULLCHECK(adj);
        AcAdjArgs aa;
        aa.env = h->env; aa.centers = centers; aa.num_links = L; aa.total = total; aa.set_size = set_size; aa.wpr = wpr; aa.adj = adj;
        dim3 grid((unsigned)((set_size + kTileCols - 1) / kTileCols), (unsigned)((set_size + kTileRows - 1) / kTileRows), (unsigned)n_sets);
        ac_adjacency_kernel<<<grid, kTileCols, 0, stream>>>(aa);
        h->stats.kernel_launches += 2;
        UPC_CUDA_TRY(cudaGetLastError());
        PassInput ain;
        ain.binary = true;
        ain.threshold = 0.0;
        const int32_t rc = cluster_pass(h, ain, n_sets, set_size, root1, stream, nullptr);
        if (rc != UPC_OK) return rc;
        group = root1;
    }
    else if (cp.feature_type == UPC_FEATURE_POINT_TO_POINT_MOVEMENT)
    {
        if (n_sets != 1)
        {
            set_error("point-to-point-movement clustering handles one particle set per call");
            return UPC_ERR_UNSUPPORTED;
        }
        if (!cp.ptpm_sim)
        {
            set_error("point-to-point-movement clustering needs simulation parameters (and a noise tape or a seed)");
            return UPC_ERR_INVALID_ARGUMENT;
        }
        const int64_t n = set_size;
        const int wpr = (int)((n + 31) / 32);
        uint32_t* adj = buf<uint32_t>(h->cw.adjacency, (size_t)n * wpr);
        UPC_NULLCHECK(adj);
        UPC_CUDA_TRY(cudaMemsetAsync(adj, 0, (size_t)n * wpr * sizeof(uint32_t), stream));
        const int width = upc_config_width(kind, h->robot.dof);
        const int64_t sims = n * n - n;
        if (sims > 0)
        {
            /* [starts | targets | results] and the collided flags */
            double* work = buf<double>(h->cw.dist_matrix, (size_t)width * (size_t)sims * 3);
            UPC_NULLCHECK(work);
            uint8_t* coll = buf<uint8_t>(h->cw.link_scratch, (size_t)sims);
            UPC_NULLCHECK(coll);
            double* starts = work;
            double* targets = work + (size_t)width * sims;
            double* results = targets + (size_t)width * sims;
            const unsigned pblocks = (unsigned)((sims / 2 + 255) / 256);
            ptpm_pairs_kernel<<<pblocks, 256, 0, stream>>>(d_configs, width, n, starts, targets);
            h->stats.kernel_launches++;
            UPC_CUDA_TRY(cudaGetLastError());
            upc_sim_params sp = *cp.ptpm_sim;
            sp.allow_contacts = 1; /* :1826 */
            /* pair simulation k replays column k of the tape, or is particle k of the seeded stream */
            const int32_t src = cp.ptpm_tape ? upc_forward_simulate_device(h, &sp, sims, starts, sims, targets, cp.ptpm_tape, results, coll, stream)
                                             : upc_forward_simulate_seeded_device(h, &sp, cp.ptpm_seed, 0, sims, sims, starts, sims, targets, results, coll, stream);
            if (src != UPC_OK) return src;
            PtpmAdjArgs pa;
            pa.robot = h->robot; pa.cfg = d_configs; pa.results = results; pa.width = width; pa.n = n; pa.wpr = wpr; pa.adj = adj;
            pa.goal_distance_threshold = cp.goal_distance_threshold;
            const unsigned ablocks = (unsigned)((std::max<int64_t>(sims / 2, n) + 255) / 256);
            if (kind == UPC_ROBOT_SE2) ptpm_adjacency_kernel<UPC_ROBOT_SE2><<<ablocks, 256, 0, stream>>>(pa);
            else if (kind == UPC_ROBOT_SE3) ptpm_adjacency_kernel<UPC_ROBOT_SE3><<<ablocks, 256, 0, stream>>>(pa);
            else ptpm_adjacency_kernel<UPC_ROBOT_LINKED><<<ablocks, 256, 0, stream>>>(pa);
            h->stats.kernel_launches++;
            UPC_CUDA_TRY(cudaGetLastError());
        }
        PassInput pin;
        pin.binary = true;
        pin.threshold = 0.0;
        const int32_t rc = cluster_pass(h, pin, n_sets, set_size, root1, stream, nullptr);
        if (rc != UPC_OK) return rc;
        group = root1;
    }
    else if (cp.feature_type != UPC_FEATURE_NONE)
    {
        set_error("unknown feature_type");
        return UPC_ERR_INVALID_ARGUMENT;
    }
    PassInput in;
    in.feat = feat;
    in.nfeat = nfeat;
    in.group = group;
    in.threshold = cp.distance_clustering_threshold;
    /* blob-shaped outcomes (the common case after a contact edge): the pivot pass settles the distance clustering in
     * O(n * clusters); UPC_B200_NO_PIVOT_PASS=1 forces the adjacency pass, UPC_B200_FORCE_PIVOT_PASS=1 tries the pivot pass on
     * small inputs too (tests exercise all three) */
    bool settled = false;
    {
        const char* off = getenv("UPC_B200_NO_PIVOT_PASS");
        /* worth it once the pair pass would cost more than a few sweeps: >= 3e7 pair tests over all sets */
        const double pair_tests = (double)n_sets * (double)set_size * (double)set_size;
        const char* force = getenv("UPC_B200_FORCE_PIVOT_PASS"); /* tests: try it whatever the size */
        const bool big_enough = (pair_tests >= 3.0e7) || (force && force[0] == '1');
        if (!(off && off[0] == '1') && set_size >= 2 && set_size < ((int64_t)1 << 31) && big_enough && cp.distance_clustering_threshold > 1.0e-5)
        {
            uint8_t* ok = buf<uint8_t>(h->cw.link_scratch, (size_t)n_sets);
            UPC_NULLCHECK(ok);
            PivotArgs pa;
            pa.robot = h->robot; pa.feat = feat; pa.nfeat = nfeat; pa.total = total; pa.set_size = set_size; pa.group = group;
            pa.threshold = cp.distance_clustering_threshold;
            pa.radius = 0.5 * cp.distance_clustering_threshold - 1.0e-6;
            pa.separation = 2.0 * cp.distance_clustering_threshold + 1.0e-6;
            pa.scratch = buf<double>(h->cw.dist_matrix, (size_t)total * 2);
            UPC_NULLCHECK(pa.scratch);
            pa.root = root2; pa.ok = ok;
            const int threads = (set_size >= 1024) ? 1024 : ((set_size >= 256) ? 256 : 64);
            /* a few large sets (the single 10^4 .. 10^5-particle outcome of one edge): one CTA per set leaves 147 SMs idle, so a
             * cluster of eight thread blocks shares each set (UPC_B200_NO_PIVOT_CLUSTER=1: one CTA per set, tests compare both) */
            const char* no_cluster = getenv("UPC_B200_NO_PIVOT_CLUSTER");
            const bool clustered = set_size >= 8192 && n_sets * kPivotCtas <= 2 * (int64_t)h->sm_count && !(no_cluster && no_cluster[0] == '1');
            if (clustered)
            {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)(n_sets * kPivotCtas), 1, 1);
                cfg.blockDim = dim3(1024, 1, 1);
                cfg.dynamicSmemBytes = 0;
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = kPivotCtas;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                if (kind == UPC_ROBOT_SE2) UPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, pivot_pass_kernel<UPC_ROBOT_SE2, kPivotCtas>, pa));
                else if (kind == UPC_ROBOT_SE3) UPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, pivot_pass_kernel<UPC_ROBOT_SE3, kPivotCtas>, pa));
                else UPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, pivot_pass_kernel<UPC_ROBOT_LINKED, kPivotCtas>, pa));
            }
            else if (kind == UPC_ROBOT_SE2) pivot_pass_kernel<UPC_ROBOT_SE2, 1><<<(unsigned)n_sets, threads, 0, stream>>>(pa);
            else if (kind == UPC_ROBOT_SE3) pivot_pass_kernel<UPC_ROBOT_SE3, 1><<<(unsigned)n_sets, threads, 0, stream>>>(pa);
            else pivot_pass_kernel<UPC_ROBOT_LINKED, 1><<<(unsigned)n_sets, threads, 0, stream>>>(pa);
            h->stats.kernel_launches++;
            UPC_CUDA_TRY(cudaGetLastError());
            std::vector<uint8_t> flags((size_t)n_sets);
            UPC_CUDA_TRY(cudaMemcpyAsync(flags.data(), ok, (size_t)n_sets, cudaMemcpyDeviceToHost, stream));
            UPC_CUDA_TRY(cudaStreamSynchronize(stream));
            h->stats.d2h_bytes += n_sets;
            settled = true;
            std::vector<int32_t> unsettled;
            for (int64_t k = 0; k < n_sets; k++)
                if (!flags[(size_t)k])
                {
                    settled = false;
                    unsettled.push_back((int32_t)k);
                }
            if (!settled && (int64_t)unsettled.size() * 2 <= n_sets)
            {
                /* a minority of the sets is left: run the pair pass on those only, packed as their own batch (the settled sets'
                 * roots stay as the pivot pass wrote them) */
                const int64_t nl = (int64_t)unsettled.size(), ctotal = nl * set_size;
                /* [set list | packed groups | packed roots] and the packed features */
                int32_t* ibuf = buf<int32_t>(h->cw.comp_list, (size_t)nl + 2 * (size_t)ctotal);
                UPC_NULLCHECK(ibuf);
                double* cfeat = buf<double>(h->cw.comp_feat, (size_t)nfeat * (size_t)ctotal);
                UPC_NULLCHECK(cfeat);
                int32_t* d_list = ibuf;
                int32_t* cgroup = ibuf + nl;
                int32_t* croot = cgroup + ctotal;
                UPC_CUDA_TRY(cudaMemcpyAsync(d_list, unsettled.data(), (size_t)nl * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
                h->stats.h2d_bytes += nl * (int64_t)sizeof(int32_t);
                const unsigned gblocks = (unsigned)((ctotal + 255) / 256);
                gather_sets_kernel<<<gblocks, 256, 0, stream>>>(feat, nfeat, total, set_size, d_list, nl, cfeat, group, group ? cgroup : nullptr);
                h->stats.kernel_launches++;
                UPC_CUDA_TRY(cudaGetLastError());
                PassInput cin = in;
                cin.feat = cfeat;
                cin.group = group ? cgroup : nullptr;
                int64_t fallback = 0;
                const int32_t rc = cluster_pass(h, cin, nl, set_size, croot, stream, &fallback);
                if (rc != UPC_OK) return rc;
                h->stats.cluster_fallback_calls += fallback;
                scatter_roots_kernel<<<gblocks, 256, 0, stream>>>(croot, d_list, nl, set_size, root2);
                h->stats.kernel_launches++;
                UPC_CUDA_TRY(cudaGetLastError());
                settled = true;
            }
        }
    }
    if (!settled)
    {
        int64_t fallback = 0;
        const int32_t rc = cluster_pass(h, in, n_sets, set_size, root2, stream, &fallback);
        if (rc != UPC_OK) return rc;
        h->stats.cluster_fallback_calls += fallback;
    }
    label_kernel<<<(unsigned)n_sets, 1024, 0, stream>>>(root2, set_size, rank, d_labels, d_n_clusters);
    h->stats.kernel_launches++;
    UPC_CUDA_TRY(cudaGetLastError());
    return UPC_OK;
}

} /* namespace upc */
