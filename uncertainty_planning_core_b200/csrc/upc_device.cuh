/*
 * upc_device.cuh - per-particle device code of the B200 particle hot path.
 *
 * Everything here is written once as __host__ __device__ so that tests/emulation can run the exact per-thread
 * program on the CPU (one lane per particle) while developing without a GPU; the product only ever launches
 * it through the __global__ kernels in upc_kernels.cu.  Arithmetic follows DESIGN.md operation by operation:
 * reference formulas (PID, actuator, SE2 distance ...) keep the reference's separate multiplies and adds,
 * the self-specified geometry (transform composition, trilinear SDF, normal equations) uses explicit fma.
 * Compile with -fmad=false: contraction must never be left to the compiler.
 */
#ifndef UPC_DEVICE_CUH
#define UPC_DEVICE_CUH

#include <stdint.h>
#if !defined(__CUDACC__)
#include <algorithm>
using std::max;
using std::min;
#endif
#include "upc_b200.h"
#include "upc_math.h"

#if defined(__CUDACC__)
#define UPC_D __host__ __device__ __forceinline__
#ifdef UPC_ALL_INLINE
#define UPC_DNI __host__ __device__ __forceinline__
#else
#define UPC_DNI inline __host__ __device__ __noinline__
#endif
#else
#define UPC_D inline
#define UPC_DNI inline
#endif
#ifndef UPC_SOLVE_OUT_OF_LINE
#define UPC_SOLVE_LINKAGE UPC_D
#else
#define UPC_SOLVE_LINKAGE UPC_DNI
#endif

namespace upc
{
struct DevEnv
{
    int nx, ny, nz;
    int planar; /* nz == 1 and every SDF value finite: the z interpolations are exact identities and are skipped */
    float sdf_oob, occ_oob;
    double res, inv_res;
    double dnx, dny, dnz; /* grid dimensions as doubles (bounds tests) */
    double G[12]; /* grid_from_world, row-major 3x4 */
    const float* sdf;
    const uint32_t* seg;
    const float* occ;
    /* (nx+1)(ny+1)(nz+1) floats: minimum of the 8 (border-clamped) corners of every trilinear interpolation cell,
     * indexed by the cell's low corner + 1.  A conservative lower bound of the interpolated distance used to skip
     * the 8-corner fetch for points that are provably clear; may be NULL (no culling). */
    const float* cellmin;
    /* optional corner-cell table: 8 floats (one 32-byte sector) per interpolation cell, same indexing as cellmin */
    const float* cells;
};

struct DevAxis
{
    double kp, ki, kd, clamp; /* already abs()'d, simple_pid_controller.hpp:104-112 */
    double vlim, nb;          /* |actuator limit|, |noise bound|, simple_uncertainty_models.hpp:59 */
    /* factors applied to a standard truncated-normal draw of this axis by the counter-based generator:
     * sensor = |max_sensor_noise| / 2, actuator = 1/2 when the axis has actuator noise; 0 = noise-free axis (exact 0.0 draws) */
    double sensor_scale, actuator_scale;
};

struct DevJoint
{
    int type, parent, child, active;
    double axis[3];
    double lo, hi;
    double T[12];
};

/* bounding box (link frame) and chunk-relative point mask of one point group; layout shared with upc_host_prep.h: HostPointGroup */
struct DevPointGroup
{
    double c[3];
    double h[3];
    unsigned mask;
    unsigned pad;
};

struct DevRobot
{
    int kind, dof, num_links, num_joints, num_points;
    float cull_delta;      /* single-precision pre-cull: bound of the float voxel-coordinate error, 0 = pre-cull off */
    int sensor_noise_free; /* every max_sensor_noise is zero: the sensed configuration is the configuration */
    int link_off[UPC_MAX_LINKS + 1];
    int parent_joint_of_link[UPC_MAX_LINKS];
    const double* pts; /* num_points * 4 doubles: x, y, z, radius; followed by num_points * 8 floats: (float) x, x, y, y, z, z, 0, 0 */
    int link_flat_z[UPC_MAX_LINKS];   /* 1: every body point of the link has the same single-precision z (planar robots) */
    float link_z[UPC_MAX_LINKS];      /* that z */
    DevAxis axis[UPC_MAX_DOF];
    double weights[UPC_MAX_DOF];
    double base[12];
    /* static bounds on how far a body point can be from the axis that moves it:
     * reach[a] for active axis a of a linked robot; reach[0] = largest |p| of the single link of SE2/SE3 robots */
    double reach[UPC_MAX_DOF];
    double link_radius[UPC_MAX_LINKS]; /* largest body-sphere radius of each link (conservative single-precision cull) */
    int axis_continuous[UPC_MAX_DOF];  /* linked robots: 1 when the active joint wraps (SimpleJointModel::IsContinuous) */
    /* link-level cull: bounding box of every link's body points in its frame, and the cell-minimum grid dilated by
     * dil_radius cells (same indexing as DevEnv::cellmin); dil_radius == 0 switches the cull off */
    int dil_radius;
    const float* dilmin;
    double link_center[UPC_MAX_LINKS][3];
    double link_half[UPC_MAX_LINKS][3];
    /* group-level cull of the one-lane programs: every chunk of 32 consecutive points of a link is split into up to
     * four compact groups (upc_host_prep.h: compute_point_groups); grpmin = cell minima dilated by grp_radius cells, the radius that
     * covers the largest group; chunk j of link l is record link_chunk_off[l] + j; grp_radius == 0 switches the cull off */
    int grp_radius;
    const float* grpmin;
    const DevPointGroup* groups; /* 4 records per chunk, unused ones have an empty mask */
    int link_chunk_off[UPC_MAX_LINKS + 1];
    DevJoint joints[UPC_MAX_JOINTS];
};

struct DevSim
{
    int num_steps, max_microsteps, allow_contacts, individual_jacobians, max_resolver_iterations, decay_iterations;
    double dt, shortcut, contact_distance, resolve_distance, microstep_distance, initial_scale, decay_rate, damping;
    /* An edge may be simulated as several launches over consecutive step ranges (the host-pointer entry point does so
     * to overlap the upload of later tape slices with the simulation of earlier ones).  The per-particle state that is
     * not part of the configuration travels between launches in carry_pid ([2*dof][n]: integral, last error) and
     * carry_flags ([2][n]: collided, finished early); both NULL for a single launch covering [0, num_steps). */
    int step_begin, step_end;
    double* carry_pid;
    uint8_t* carry_flags;
    /* starts [C][n_starts] and targets [C][n_targets] may hold fewer entries than particles: particle i reads entry
     * i / per_start (i / per_target), per_x = n / n_x - one entry per particle, one per edge of an edge batch, or one for all */
    int64_t n_starts, per_start, per_target;
    /* seeded launches (no tape): draws are a function of (seed, first_particle + i, step, slot, axis), DESIGN.md 4.2 */
    uint64_t seed;
    int64_t first_particle;
    /* optional per-interval trace of a (small) batch: configuration after every controller interval, [num_steps][C][n], and the
     * the control inputs behind it (SimulatorInterface::ForwardSimulateRobot with enable_tracing,
     * simple_simulator_interface.hpp:40-86, 621); NULL = no trace */
    double* trace_configs;
    double* trace_controls; /* [num_steps][2 * dof][n]: control input (velocity command), then the control step applied per microstep */
    int32_t* trace_steps;   /* [2][n]: controller intervals executed, microsteps of the last one */
};

/* per-launch counters, reduced per block then added to the handle's totals */
struct DevCounters
{
    unsigned long long particle_steps, collided, resolver_iterations, failed_resolves;
    unsigned long long phase[12]; /* UPC_PHASE_TIMING builds only: cycles per phase, summed over particles */
};

/* floor(x) as a 32-bit integer in one conversion instruction (|x| < 2^31 by the grid-size limit) */
UPC_D int floor_to_int(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2int_rd(x);
#else
    return (int)floor(x);
#endif
}

/*
 * floor(x) as a double and as a 32-bit integer WITHOUT the conversion unit (FP64 <-> integer conversions have a 45-cycle
 * dependent latency and a quarter of the DFMA rate on sm_100a): adding 1.5 * 2^52 leaves round-to-nearest(x) in the low
 * mantissa bits; one compare fixes the round-up cases.  Exact for |x| < 2^31, identical on host and device.
 */
UPC_D int floor_parts(double x, double* floor_x)
{
    const double magic = 6755399441055744.0;
    const double t = x + magic;
    double r = t - magic;
#if defined(__CUDA_ARCH__)
    int lo = __double2loint(t);
#else
    long long bits;
    __builtin_memcpy(&bits, &t, sizeof(bits));
    int lo = (int)(unsigned int)(bits & 0xffffffffLL);
#endif
    if (r > x)
    {
        r -= 1.0;
        lo -= 1;
    }
    *floor_x = r;
    return lo;
}

template <typename T>
UPC_D T ldg(const T* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

/* ------------------------------------------------------------------ rigid transforms, row-major 3x4 in 12 doubles */

UPC_D void xf_compose(const double* a, const double* b, double* c)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            c[i * 4 + j] = upc_fma(a[i * 4 + 2], b[8 + j], upc_fma(a[i * 4 + 1], b[4 + j], a[i * 4 + 0] * b[j]));
        }
        c[i * 4 + 3] = upc_fma(a[i * 4 + 2], b[11], upc_fma(a[i * 4 + 1], b[7], upc_fma(a[i * 4 + 0], b[3], a[i * 4 + 3])));
    }
}

UPC_D void xf_point(const double* a, double px, double py, double pz, double* out)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        out[i] = upc_fma(a[i * 4 + 2], pz, upc_fma(a[i * 4 + 1], py, upc_fma(a[i * 4 + 0], px, a[i * 4 + 3])));
    }
}

UPC_D void cross3(const double* a, const double* b, double* out)
{
    out[0] = upc_fma(a[1], b[2], -(a[2] * b[1]));
    out[1] = upc_fma(a[2], b[0], -(a[0] * b[2]));
    out[2] = upc_fma(a[0], b[1], -(a[1] * b[0]));
}

UPC_D double dot3(const double* a, const double* b) { return upc_fma(a[2], b[2], upc_fma(a[1], b[1], a[0] * b[0])); }

/* V = inv_res * (grid_from_world o link) : link frame -> continuous voxel coordinates */
UPC_D void voxel_from_link(const DevEnv& env, const double* link, double* V)
{
    xf_compose(env.G, link, V);
#pragma unroll
    for (int k = 0; k < 12; k++) V[k] = V[k] * env.inv_res;
}

/* ------------------------------------------------------------------ SE(3) closed forms */

UPC_D void se3_log(const double* T, double* twist)
{
    const double trace = (T[0] + T[5]) + T[10];
    double s[3];
    s[0] = 0.5 * (T[9] - T[6]);
    s[1] = 0.5 * (T[2] - T[8]);
    s[2] = 0.5 * (T[4] - T[1]);
    const double sn = sqrt(dot3(s, s));
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
        const double omc = 1.0 - ct;
        const double a20 = (T[0] - ct) / omc;
        const double a21 = (T[5] - ct) / omc;
        const double a22 = (T[10] - ct) / omc;
        int k = 0;
        double best = a20;
        if (a21 > best) { k = 1; best = a21; }
        if (a22 > best) { k = 2; best = a22; }
        const double akm = sqrt(upc_max(best, 0.0));
        const double sk = (k == 0) ? s[0] : ((k == 1) ? s[1] : s[2]);
        const double ak = (sk < 0.0) ? -akm : akm;
        const double denom = (2.0 * omc) * ak;
        double a[3];
        /* (R[k][j] + R[j][k]) / denom for j != k */
        a[0] = (k == 0) ? ak : ((T[k * 4 + 0] + T[0 * 4 + k]) / denom);
        a[1] = (k == 1) ? ak : ((T[k * 4 + 1] + T[1 * 4 + k]) / denom);
        a[2] = (k == 2) ? ak : ((T[k * 4 + 2] + T[2 * 4 + k]) / denom);
        w[0] = a[0] * theta; w[1] = a[1] * theta; w[2] = a[2] * theta;
    }
    double k2;
    const double th2 = theta * theta;
    if (theta < 0.05)
    {
        /* 1/12 + th^2/720 + th^4/30240 + th^6/1209600 + th^8/47900160 */
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
    const double tr[3] = {T[3], T[7], T[11]};
    double wxt[3], wxwxt[3];
    cross3(w, tr, wxt);
    cross3(w, wxt, wxwxt);
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        twist[i] = upc_fma(k2, wxwxt[i], upc_fma(-0.5, wxt[i], tr[i]));
        twist[3 + i] = w[i];
    }
}

UPC_D void se3_exp(const double* twist, double* T)
{
    const double* v = twist;
    const double* w = twist + 3;
    const double th2 = dot3(w, w);
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
        const double inv = 1.0 / theta; /* the only division of the exponential */
        A = sn * inv;
        B = ((1.0 - cs) * inv) * inv;
        C = ((1.0 - A) * inv) * inv;
    }
    const double wx = w[0], wy = w[1], wz = w[2];
    T[0] = upc_fma(B, upc_fma(wx, wx, -th2), 1.0);
    T[5] = upc_fma(B, upc_fma(wy, wy, -th2), 1.0);
    T[10] = upc_fma(B, upc_fma(wz, wz, -th2), 1.0);
    T[1] = upc_fma(B, wx * wy, -(A * wz));
    T[4] = upc_fma(B, wx * wy, A * wz);
    T[2] = upc_fma(B, wx * wz, A * wy);
    T[8] = upc_fma(B, wx * wz, -(A * wy));
    T[6] = upc_fma(B, wy * wz, -(A * wx));
    T[9] = upc_fma(B, wy * wz, A * wx);
    double wxv[3], wxwxv[3];
    cross3(w, v, wxv);
    cross3(w, wxv, wxwxv);
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        T[i * 4 + 3] = upc_fma(C, wxwxv[i], upc_fma(B, wxv[i], v[i]));
    }
}

/* a^-1 o b */
UPC_D void se3_between(const double* a, const double* b, double* r)
{
    const double d0 = b[3] - a[3], d1 = b[7] - a[7], d2 = b[11] - a[11];
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            r[i * 4 + j] = upc_fma(a[8 + i], b[8 + j], upc_fma(a[4 + i], b[4 + j], a[i] * b[j]));
        }
        r[i * 4 + 3] = upc_fma(a[8 + i], d2, upc_fma(a[4 + i], d1, a[i] * d0));
    }
}

/* rotation -> quaternion (x, y, z, w) with Eigen's branch structure */
UPC_D void quat_from_xf(const double* T, double* q)
{
    double t = (T[0] + T[5]) + T[10];
    if (t > 0.0)
    {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (T[9] - T[6]) * t;
        q[1] = (T[2] - T[8]) * t;
        q[2] = (T[4] - T[1]) * t;
    }
    else
    {
        int i = 0;
        if (T[5] > T[0]) i = 1;
        if (T[10] > T[i * 5]) i = 2;
        const int j = (i + 1) % 3;
        const int k = (j + 1) % 3;
        t = sqrt(((T[i * 5] - T[j * 5]) - T[k * 5]) + 1.0);
        const double qi = 0.5 * t;
        t = 0.5 / t;
        const double qw = (T[k * 4 + j] - T[j * 4 + k]) * t;
        const double qj = (T[j * 4 + i] + T[i * 4 + j]) * t;
        const double qk = (T[k * 4 + i] + T[i * 4 + k]) * t;
        q[3] = qw;
        q[0] = (i == 0) ? qi : ((j == 0) ? qj : qk);
        q[1] = (i == 1) ? qi : ((j == 1) ? qj : qk);
        q[2] = (i == 2) ? qi : ((j == 2) ? qj : qk);
    }
}

/* ------------------------------------------------------------------ PID / actuator (reference arithmetic, no fma) */

UPC_D double pid_step(const DevAxis& ax, double& integral, double& last_error, double e, double dt)
{
    const double step_integral = ((e * 0.5) + (last_error * 0.5)) * dt;
    const double new_integral = integral + step_integral;
    integral = upc_max(-ax.clamp, upc_min(ax.clamp, new_integral));
    const double derivative = (e - last_error) / dt;
    last_error = e;
    return ((e * ax.kp) + (integral * ax.ki)) + (derivative * ax.kd);
}

UPC_D double actuator_clamp(const DevAxis& ax, double u)
{
    double r = upc_min(ax.vlim, u);
    r = upc_max(-ax.vlim, r);
    return r;
}

UPC_D double actuator_noisy(const DevAxis& ax, double u, double draw)
{
    const double r = actuator_clamp(ax, u);
    const double noise = draw * (ax.nb * r);
    return r + noise;
}

/* ------------------------------------------------------------------ SDF sampling */

UPC_D bool voxel_in_bounds(const DevEnv& env, const double* u)
{
    return (u[0] >= 0.0) && (u[0] < env.dnx) && (u[1] >= 0.0) && (u[1] < env.dny) && (u[2] >= 0.0) && (u[2] < env.dnz);
}

UPC_D int64_t cell_index(const DevEnv& env, int ix, int iy, int iz)
{
    return ((int64_t)ix * (int64_t)env.ny + (int64_t)iy) * (int64_t)env.nz + (int64_t)iz;
}

/* distance only */
UPC_D double sdf_distance(const DevEnv& env, const double* u)
{
    if (!voxel_in_bounds(env, u)) return (double)env.sdf_oob;
    const double cx = u[0] - 0.5, cy = u[1] - 0.5, cz = u[2] - 0.5;
    const int ix = floor_to_int(cx), iy = floor_to_int(cy), iz = floor_to_int(cz);
    const double fx = cx - (double)ix, fy = cy - (double)iy, fz = cz - (double)iz;
    const int x0 = (ix < 0) ? 0 : ix, x1 = (ix + 1 > env.nx - 1) ? env.nx - 1 : ix + 1;
    const int y0 = (iy < 0) ? 0 : iy, y1 = (iy + 1 > env.ny - 1) ? env.ny - 1 : iy + 1;
    const int z0 = (iz < 0) ? 0 : iz, z1 = (iz + 1 > env.nz - 1) ? env.nz - 1 : iz + 1;
    /* grids are limited to 2^31 cells (checked in upc_create): 32-bit index arithmetic */
    const int r00 = (x0 * env.ny + y0) * env.nz;
    const int r01 = (x0 * env.ny + y1) * env.nz;
    const int r10 = (x1 * env.ny + y0) * env.nz;
    const int r11 = (x1 * env.ny + y1) * env.nz;
    const double v000 = (double)ldg(env.sdf + r00 + z0), v001 = (double)ldg(env.sdf + r00 + z1);
    const double v010 = (double)ldg(env.sdf + r01 + z0), v011 = (double)ldg(env.sdf + r01 + z1);
    const double v100 = (double)ldg(env.sdf + r10 + z0), v101 = (double)ldg(env.sdf + r10 + z1);
    const double v110 = (double)ldg(env.sdf + r11 + z0), v111 = (double)ldg(env.sdf + r11 + z1);
    const double c00 = upc_fma(fz, v001 - v000, v000);
    const double c01 = upc_fma(fz, v011 - v010, v010);
    const double c10 = upc_fma(fz, v101 - v100, v100);
    const double c11 = upc_fma(fz, v111 - v110, v110);
    const double c0 = upc_fma(fy, c01 - c00, c00);
    const double c1 = upc_fma(fy, c11 - c10, c10);
    return upc_fma(fx, c1 - c0, c0);
}

/* margin (metres) by which the cell-minimum bound must clear a threshold before the exact interpolation is skipped;
 * the interpolant can undershoot the smallest corner only by a few ulps of the corner values */
#define UPC_CULL_MARGIN 1.0e-9

/* true when the point at voxel coordinate u is certainly at distance >= limit + radius (so the exact test
 * "distance - radius < limit" is false); false means "unknown, evaluate exactly". */
UPC_D bool sdf_certainly_clear(const DevEnv& env, const double* u, double radius, double limit)
{
    if (!voxel_in_bounds(env, u)) return (((double)env.sdf_oob - radius) - limit) > UPC_CULL_MARGIN;
    if (env.cellmin == nullptr) return false;
    const int ix = floor_to_int(u[0] - 0.5) + 1, iy = floor_to_int(u[1] - 0.5) + 1, iz = floor_to_int(u[2] - 0.5) + 1;
    const double m = (double)ldg(env.cellmin + ((ix * (env.ny + 1) + iy) * (env.nz + 1) + iz));
    return ((m - radius) - limit) > UPC_CULL_MARGIN;
}

/*
 * Link-level cull: true when EVERY body point of link l is certainly at distance >= limit + its radius, decided from ONE lookup.
 * V is the link's voxel-from-link transform.  The link's points lie in the box centre +- half of the link frame, hence - axis by
 * axis - within c_a +- e_a in voxel coordinates, c = V centre, e_a = sum_j |V_aj| half_j.  When that interval is inside the grid
 * on every axis all points are in bounds, and the cell every point interpolates in lies within dil_radius cells of the centre's
 * cell (dil_radius >= half diagonal in cells + 1.5), so the dilated cell minimum at the centre's cell bounds all their
 * interpolated distances from below (to the few ulps the cull margin covers).  Conservative: anything else returns false and the
 * per-point passes decide.  clear_above is the rounded-up single-precision threshold the per-point cull uses.
 */
UPC_D bool link_certainly_clear(const DevEnv& env, const DevRobot& r, int l, const double* V, float clear_above)
{
    if (r.dil_radius == 0) return false;
    double c[3];
    xf_point(V, r.link_center[l][0], r.link_center[l][1], r.link_center[l][2], c);
    const double hx = r.link_half[l][0], hy = r.link_half[l][1], hz = r.link_half[l][2];
    /* 1e-6 cells of slack absorb the rounding of V, c and e (coordinates are below 2^31 cells) */
    const double ex = (fabs(V[0]) * hx + fabs(V[1]) * hy) + (fabs(V[2]) * hz + 1.0e-6);
    const double ey = (fabs(V[4]) * hx + fabs(V[5]) * hy) + (fabs(V[6]) * hz + 1.0e-6);
    const double ez = (fabs(V[8]) * hx + fabs(V[9]) * hy) + (fabs(V[10]) * hz + 1.0e-6);
    const bool inside = (c[0] - ex >= 0.0) & (c[0] + ex < env.dnx) & (c[1] - ey >= 0.0) & (c[1] + ey < env.dny) & (c[2] - ez >= 0.0) & (c[2] + ez < env.dnz);
    if (!inside) return false;
    const int ix = floor_to_int(c[0] - 0.5) + 1, iy = floor_to_int(c[1] - 0.5) + 1, iz = floor_to_int(c[2] - 0.5) + 1;
    const float m = ldg(r.dilmin + ((ix * (env.ny + 1) + iy) * (env.nz + 1) + iz));
    return m > clear_above;
}

/*
 * Group-level cull (one-lane programs): the points of chunk `chunk_index` of link l that are NOT proven clear by one dilated
 * lookup per point group - same test and same proof as link_certainly_clear, with the group's box and the grid dilated by the
 * largest group's radius.  Bit j = point chunk + j still needs the per-point passes.  All ones when the cull is off.
 */
/* compiled out by default: measured on B200 the group tests cost more than they save (C5 25.4 -> 26.3 ms, C3 +-0: when a link is
 * not clear as a whole, most of its groups are not either at these robot / cell sizes); -DUPC_GROUP_CULL=1 builds it in, and the
 * host emulation of the test suite always does */
#ifndef UPC_GROUP_CULL
#define UPC_GROUP_CULL 0
#endif
UPC_D unsigned group_candidates(const DevEnv& env, const DevRobot& r, int l, int chunk_index, const double* V, float clear_above)
{
    if (!UPC_GROUP_CULL || r.grp_radius == 0) return 0xffffffffu;
    const DevPointGroup* grp = r.groups + (int64_t)(r.link_chunk_off[l] + chunk_index) * 4;
    const double a00 = fabs(V[0]), a01 = fabs(V[1]), a02 = fabs(V[2]);
    const double a10 = fabs(V[4]), a11 = fabs(V[5]), a12 = fabs(V[6]);
    const double a20 = fabs(V[8]), a21 = fabs(V[9]), a22 = fabs(V[10]);
    unsigned candidates = 0u;
#pragma unroll 1
    for (int q = 0; q < 4; q++)
    {
        const unsigned mask = ldg(&grp[q].mask);
        if (mask == 0u) break;
        const double hx = ldg(&grp[q].h[0]), hy = ldg(&grp[q].h[1]), hz = ldg(&grp[q].h[2]);
        double c[3];
        xf_point(V, ldg(&grp[q].c[0]), ldg(&grp[q].c[1]), ldg(&grp[q].c[2]), c);
        const double ex = (a00 * hx + a01 * hy) + (a02 * hz + 1.0e-6);
        const double ey = (a10 * hx + a11 * hy) + (a12 * hz + 1.0e-6);
        const double ez = (a20 * hx + a21 * hy) + (a22 * hz + 1.0e-6);
        const bool inside = (c[0] - ex >= 0.0) & (c[0] + ex < env.dnx) & (c[1] - ey >= 0.0) & (c[1] + ey < env.dny) & (c[2] - ez >= 0.0) & (c[2] + ez < env.dnz);
        bool clear = false;
        if (inside)
        {
            const int ix = floor_to_int(c[0] - 0.5) + 1, iy = floor_to_int(c[1] - 0.5) + 1, iz = floor_to_int(c[2] - 0.5) + 1;
            clear = ldg(r.grpmin + ((ix * (env.ny + 1) + iy) * (env.nz + 1) + iz)) > clear_above;
        }
        if (!clear) candidates |= mask;
    }
    return candidates;
}

/* smallest float >= v (v finite): a single-precision value may be compared against it instead of against v */
UPC_D float float_round_up(double v)
{
    float f = (float)v;
    if ((double)f < v)
    {
#if defined(__CUDA_ARCH__)
        f = __int_as_float(__float_as_int(f) + ((f >= 0.0f) ? 1 : -1));
#else
        f = __builtin_nextafterf(f, __builtin_inff());
#endif
    }
    return f;
}

struct alignas(16) Corner4
{
    float x, y, z, w;
};

UPC_D Corner4 ldg4(const float* p)
{
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    Corner4 c;
    c.x = v.x; c.y = v.y; c.z = v.z; c.w = v.w;
    return c;
#else
    Corner4 c;
    c.x = p[0]; c.y = p[1]; c.z = p[2]; c.w = p[3];
    return c;
#endif
}

UPC_D float fmin4(const Corner4& a, const Corner4& b)
{
    const float m0 = (a.y < a.x) ? a.y : a.x, m1 = (a.w < a.z) ? a.w : a.z;
    const float m2 = (b.y < b.x) ? b.y : b.x, m3 = (b.w < b.z) ? b.w : b.z;
    const float m01 = (m1 < m0) ? m1 : m0, m23 = (m3 < m2) ? m3 : m2;
    return (m23 < m01) ? m23 : m01;
}

/*
 * Single-precision pre-cull.  The voxel coordinate of a body point is recomputed in float with an
 * absolute error below r.cull_delta (bound derived in upc_host_prep.h: compute_cull_delta).  A point is declared clear
 * only when (a) it is inside the grid by more than that error, (b) its cell fractions are further than that error
 * from 0 and 1, so the float cell is the cell the FP64 program would pick, and (c) the minimum corner of that cell
 * exceeds the rounded-up threshold - the same bound the FP64 corner test uses.  Anything else is "undecided" and
 * goes through the exact FP64 program, so no result ever depends on this code.
 */
/* a body point as the pre-cull reads it: every coordinate twice, so that one packed multiply-add (FFMA2, sm_100) serves the x and
 * the y voxel axis */
struct alignas(16) CullPoint
{
    float xx[2], yy[2], zz[2], pad[2];
};

struct Float2
{
    float x, y;
};

/* two single-precision fused multiply-adds / additions in one instruction on sm_100 (FFMA2 / FADD2: crt/sm_100_rt.h); each half
 * rounds exactly like the scalar operation, which is what the host build executes */
#ifndef UPC_CULL_PACKED
#define UPC_CULL_PACKED 1
#endif
UPC_D Float2 fma2(const Float2& a, const Float2& b, const Float2& c)
{
    Float2 out;
#if defined(__CUDA_ARCH__) && UPC_CULL_PACKED
    const float2 v = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    out.x = v.x; out.y = v.y;
#elif defined(__CUDA_ARCH__)
    out.x = __fmaf_rn(a.x, b.x, c.x);
    out.y = __fmaf_rn(a.y, b.y, c.y);
#else
    out.x = __builtin_fmaf(a.x, b.x, c.x);
    out.y = __builtin_fmaf(a.y, b.y, c.y);
#endif
    return out;
}

UPC_D Float2 add2(const Float2& a, const Float2& b)
{
    Float2 out;
#if defined(__CUDA_ARCH__) && UPC_CULL_PACKED
    const float2 v = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    out.x = v.x; out.y = v.y;
#elif defined(__CUDA_ARCH__)
    out.x = __fadd_rn(a.x, b.x);
    out.y = __fadd_rn(a.y, b.y);
#else
    out.x = a.x + b.x;
    out.y = a.y + b.y;
#endif
    return out;
}

#ifndef UPC_CULL_PIN_FRAME
#define UPC_CULL_PIN_FRAME 1
#endif

struct CullFrame
{
    Float2 vxy[3], txy; /* rows x and y of voxel_from_link, column by column, with the cell offset folded into the translation */
    float vz[3], tz;    /* row z */
    float gx, gy, gz;   /* largest |rounding residue| per axis that still pins the cell */
    int wx, wy, wz;     /* the rounded coordinate must lie in [0, w] */
    float clear_above;
    int sy, sx;         /* cell-minimum table strides of the x and y index: index = ix * sx + iy * sy + iz + base */
    int base;
    /* z handled once per link: every point of the link has the same z and the z row of the transform ignores x and y (planar
     * robots in axis-aligned grids).  z_index already holds iz; z_fail = the z axis cannot be pinned, nothing can be culled */
    bool z_uniform, z_fail;
};

UPC_D float upc_fmaf(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

UPC_D int float_bits(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(x);
#else
    int b;
    __builtin_memcpy(&b, &x, sizeof(b));
    return b;
#endif
}

/* round-to-nearest of |q| < 2^22 without the conversion unit: the integer sits in the low mantissa bits of q + 1.5 * 2^23 */
UPC_D int cull_round(float q, float* residue)
{
    const float magic = 12582912.0f;
    const float t = q + magic;
    *residue = q - (t - magic);
    return float_bits(t) - 0x4B400000;
}

/*
 * Per axis the float coordinate q is arranged so that round-to-nearest(q) identifies the cell:
 *   n >= 2: q = u - 1 = (cell coordinate) - 0.5; cell = round(q) when |q - round(q)| <= 0.5 - delta; only interior
 *           cells 0..n-2 are accepted (border cells need the exact in-bounds test: undecided);
 *   n == 1: q = u - 0.5 = cell coordinate; any in-bounds point has |q| < 0.5 and both candidate cells (-1, 0) hold the
 *           same clamped corners, so cell 0 stands for both: accepted when round(q) = 0 and |q| <= 0.5 - 2 delta.
 * The cell-minimum table is indexed by cell + 1.
 */
template <bool PLANAR>
UPC_D CullFrame cull_frame(const DevEnv& env, const DevRobot& r, int l, const double* V, float delta, float clear_above)
{
    CullFrame f;
    f.vxy[0].x = (float)V[0]; f.vxy[1].x = (float)V[1]; f.vxy[2].x = (float)V[2];
    f.vxy[0].y = (float)V[4]; f.vxy[1].y = (float)V[5]; f.vxy[2].y = (float)V[6];
    f.vz[0] = (float)V[8]; f.vz[1] = (float)V[9]; f.vz[2] = (float)V[10];
    f.txy.x = (float)(V[3] - ((env.nx == 1) ? 0.5 : 1.0));
    f.txy.y = (float)(V[7] - ((env.ny == 1) ? 0.5 : 1.0));
    f.tz = (float)(V[11] - ((env.nz == 1) ? 0.5 : 1.0));
    const float two = delta + delta;
    f.gx = 0.5f - ((env.nx == 1) ? two : delta);
    f.gy = 0.5f - ((env.ny == 1) ? two : delta);
    f.gz = 0.5f - ((env.nz == 1) ? two : delta);
    f.wx = (env.nx == 1) ? 0 : env.nx - 2;
    f.wy = (env.ny == 1) ? 0 : env.ny - 2;
    f.wz = (env.nz == 1) ? 0 : env.nz - 2;
    f.clear_above = clear_above;
    f.sy = env.nz + 1;
    f.sx = (env.ny + 1) * (env.nz + 1);
    f.base = f.sx + f.sy + 1; /* the table is indexed by cell + 1 on every axis */
    /* the z coordinate of every point of the link is the same number when the points share their z and the row has no x / y part:
     * fma(0, p, t) = t exactly, so the value below is the value the per-point chain would produce */
    f.z_uniform = PLANAR && (r.link_flat_z[l] != 0) && (f.vz[0] == 0.0f) && (f.vz[1] == 0.0f);
    f.z_fail = false;
    if (f.z_uniform)
    {
        const float qz = upc_fmaf(f.vz[2], r.link_z[l], f.tz);
        float rz;
        const int iz = cull_round(qz, &rz);
        f.z_fail = !((fabsf(rz) <= f.gz) & ((unsigned)iz <= (unsigned)f.wz));
        f.base += iz;
    }
#if defined(__CUDA_ARCH__) && UPC_CULL_PIN_FRAME
    /* planar programs keep the frame in registers: left alone ptxas rematerialises these conversions (F2F, constant loads, selects)
     * for every point of the pass instead of holding 16 values across the loop (C5 21.9 -> 21.3 ms).  The 3-D programs (SE(3),
     * linked: 168 registers, already spilling) are better off without: C2 1.80 ms against 1.95 ms with the frame pinned. */
    if constexpr (PLANAR)
    {
        asm volatile("" : "+f"(f.vxy[0].x), "+f"(f.vxy[0].y), "+f"(f.vxy[1].x), "+f"(f.vxy[1].y), "+f"(f.vxy[2].x), "+f"(f.vxy[2].y), "+f"(f.txy.x), "+f"(f.txy.y));
        asm volatile("" : "+f"(f.gx), "+f"(f.gy), "+r"(f.wx), "+r"(f.wy), "+r"(f.sx), "+r"(f.sy), "+r"(f.base), "+f"(f.clear_above));
    }
#endif
    return f;
}

/* index into the cell-minimum table of the cell holding the point, or -1 when the float arithmetic cannot pin it.
 * PLANAR programs (SE(2)): x and y axes in packed instructions, z already folded into the frame when the link's z is uniform;
 * the 3-D programs: three scalar chains (the form they were tuned with: packing them cost C2 8 %) */
template <bool PLANAR>
UPC_D int cull_locate(const CullFrame& f, const CullPoint& p)
{
    if constexpr (PLANAR)
    {
        const Float2 px = {p.xx[0], p.xx[1]}, py = {p.yy[0], p.yy[1]}, pz = {p.zz[0], p.zz[1]};
        const Float2 q = fma2(f.vxy[2], pz, fma2(f.vxy[1], py, fma2(f.vxy[0], px, f.txy)));
        const Float2 magic = {12582912.0f, 12582912.0f}, minus_magic = {-12582912.0f, -12582912.0f}, minus_one = {-1.0f, -1.0f};
        const Float2 t = add2(q, magic);
        const Float2 res = fma2(add2(t, minus_magic), minus_one, q); /* q - (t - magic): the product with -1 is exact */
        const int ix = float_bits(t.x) - 0x4B400000, iy = float_bits(t.y) - 0x4B400000;
        bool ok = (fabsf(res.x) <= f.gx) & (fabsf(res.y) <= f.gy) & ((unsigned)ix <= (unsigned)f.wx) & ((unsigned)iy <= (unsigned)f.wy);
        int idx = ix * f.sx + iy * f.sy + f.base;
        if (!f.z_uniform)
        {
            const float qz = upc_fmaf(f.vz[2], p.zz[0], upc_fmaf(f.vz[1], p.yy[0], upc_fmaf(f.vz[0], p.xx[0], f.tz)));
            float rz;
            const int iz = cull_round(qz, &rz);
            ok = ok & (fabsf(rz) <= f.gz) & ((unsigned)iz <= (unsigned)f.wz);
            idx += iz;
        }
        return ok ? idx : -1;
    }
    else
    {
        const float x = p.xx[0], y = p.yy[0], z = p.zz[0];
        const float qx = upc_fmaf(f.vxy[2].x, z, upc_fmaf(f.vxy[1].x, y, upc_fmaf(f.vxy[0].x, x, f.txy.x)));
        const float qy = upc_fmaf(f.vxy[2].y, z, upc_fmaf(f.vxy[1].y, y, upc_fmaf(f.vxy[0].y, x, f.txy.y)));
        const float qz = upc_fmaf(f.vz[2], z, upc_fmaf(f.vz[1], y, upc_fmaf(f.vz[0], x, f.tz)));
        float rx, ry, rz;
        const int ix = cull_round(qx, &rx), iy = cull_round(qy, &ry), iz = cull_round(qz, &rz);
        const bool ok = (fabsf(rx) <= f.gx) & (fabsf(ry) <= f.gy) & (fabsf(rz) <= f.gz) & ((unsigned)ix <= (unsigned)f.wx) &
                        ((unsigned)iy <= (unsigned)f.wy) & ((unsigned)iz <= (unsigned)f.wz);
        const int idx = ix * f.sx + iy * f.sy + iz + f.base;
        return ok ? idx : -1;
    }
}

/* phase A of a point pass: bit j set = point first + j * stride (< lim) needs the exact FP64 evaluation.
 * UPC_CULL_INFLIGHT independent lookups are put in flight per trip (branch-free locate, then the loads, then the tests);
 * four measured best (C2 2.10 ms, 10^6-particle batch 2.1e9 particle-steps/s; eight: 2.14 ms / 1.8e9; one: 2.40 ms / 1.6e9). */
#ifndef UPC_CULL_INFLIGHT
#define UPC_CULL_INFLIGHT 4
#endif
template <bool PLANAR>
UPC_D unsigned cull_undecided(const DevEnv& env, const CullFrame& f, const CullPoint* ptsf, int first, int lim, int stride, unsigned candidates = 0xffffffffu)
{
    unsigned undecided = 0u;
    if (f.z_fail)
    {
        /* the link's common z cannot be pinned to a cell layer: every point goes through the exact program */
        for (int j = 0; j < 32; j++)
            if (first + j * stride < lim) undecided |= (1u << j);
        return undecided & candidates;
    }
    constexpr unsigned kTrip = (UPC_CULL_INFLIGHT >= 32) ? 0xffffffffu : ((1u << UPC_CULL_INFLIGHT) - 1u);
    for (int j0 = 0; j0 < 32; j0 += UPC_CULL_INFLIGHT)
    {
        if (first + j0 * stride >= lim) break;
        if (((candidates >> j0) & kTrip) == 0u) continue; /* every point of this trip was cleared by the group-level cull */
        int idx[UPC_CULL_INFLIGHT];
#pragma unroll
        for (int q = 0; q < UPC_CULL_INFLIGHT; q++)
        {
            const int k = first + (j0 + q) * stride;
            idx[q] = cull_locate<PLANAR>(f, ptsf[(k < lim) ? k : first]);
        }
        float m[UPC_CULL_INFLIGHT];
#pragma unroll
        for (int q = 0; q < UPC_CULL_INFLIGHT; q++) m[q] = ldg(env.cellmin + max(idx[q], 0));
#pragma unroll
        for (int q = 0; q < UPC_CULL_INFLIGHT; q++)
        {
            const int k = first + (j0 + q) * stride;
            const bool clear = (idx[q] >= 0) & (m[q] > f.clear_above);
            if ((k < lim) & !clear) undecided |= (1u << (j0 + q));
        }
    }
    return undecided & candidates;
}

/* single-point form (tests) */
template <bool PLANAR>
UPC_D bool cull_certainly_clear(const DevEnv& env, const CullFrame& f, const CullPoint& p)
{
    if (f.z_fail) return false;
    const int idx = cull_locate<PLANAR>(f, p);
    if (idx < 0) return false;
    return ldg(env.cellmin + idx) > f.clear_above;
}

UPC_D int lowest_bit(unsigned m)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

/*
 * Corner-cell path, split so that callers can put several independent lookups in flight before the first use
 * (the loads are the long pole: an L2 hit is ~300 cycles, and the compiler will not hoist them across the
 * per-point branches on its own):
 *   cell_locate  voxel coordinate -> address of the 32-byte corner record (always legal: clamped), fractions, in-bounds
 *   cell_below   "sdf - radius < limit" from the loaded corners: single-precision minimum against a rounded-up
 *                threshold decides the clear case, otherwise the exact trilinear value (same operations as sdf_distance)
 */
struct CellRef
{
    const float* cell;
    double fx, fy, fz;
    bool inb;
};

UPC_D CellRef cell_locate(const DevEnv& env, const double* u)
{
    CellRef c;
    c.inb = voxel_in_bounds(env, u);
    const double cx = u[0] - 0.5, cy = u[1] - 0.5, cz = u[2] - 0.5;
    double flx, fly, flz;
    int ix = floor_parts(cx, &flx), iy = floor_parts(cy, &fly), iz = floor_parts(cz, &flz);
    c.fx = cx - flx; c.fy = cy - fly; c.fz = cz - flz;
    ix = max(-1, min(ix, env.nx - 1));
    iy = max(-1, min(iy, env.ny - 1));
    iz = max(-1, min(iz, env.nz - 1));
    c.cell = env.cells + (int64_t)(((ix + 1) * (env.ny + 1) + (iy + 1)) * (env.nz + 1) + (iz + 1)) * 8;
    return c;
}

/* Single-layer grids (nz == 1, the 2-D worlds of the SE(2) robot) with finite values: both z corners of every cell are the same
 * clamped voxel, so the four z interpolations fma(fz, v - v, v) return v exactly and the z derivative is exactly +0 - skipped, bit
 * for bit the same result (DevEnv::planar is set by the host only when every SDF value is finite: inf - inf would be NaN) */
#ifndef UPC_PLANAR_SHORTCUT
#define UPC_PLANAR_SHORTCUT 1
#endif

UPC_D bool cell_below(const DevEnv& env, const CellRef& c, const Corner4& a, const Corner4& b, double radius, double limit, float clear_above)
{
    if (!c.inb) return ((double)env.sdf_oob - radius) < limit;
    if (fmin4(a, b) > clear_above) return false;
    const double v000 = (double)a.x, v010 = (double)a.z;
    const double v100 = (double)b.x, v110 = (double)b.z;
    double c00 = v000, c01 = v010, c10 = v100, c11 = v110;
    if (!(UPC_PLANAR_SHORTCUT && env.planar))
    {
        const double v001 = (double)a.y, v011 = (double)a.w, v101 = (double)b.y, v111 = (double)b.w;
        c00 = upc_fma(c.fz, v001 - v000, v000);
        c01 = upc_fma(c.fz, v011 - v010, v010);
        c10 = upc_fma(c.fz, v101 - v100, v100);
        c11 = upc_fma(c.fz, v111 - v110, v110);
    }
    const double c0 = upc_fma(c.fy, c01 - c00, c00);
    const double c1 = upc_fma(c.fy, c11 - c10, c10);
    const double d = upc_fma(c.fx, c1 - c0, c0);
    return (d - radius) < limit;
}

UPC_D bool sdf_below_cells(const DevEnv& env, const double* u, double radius, double limit, float clear_above)
{
    const CellRef c = cell_locate(env, u);
    const Corner4 a = ldg4(c.cell), b = ldg4(c.cell + 4);
    return cell_below(env, c, a, b, radius, limit, clear_above);
}

/* Corner-cell path of the resolver: distance and world gradient when the point is (possibly) closer than `limit`;
 * returns false - nothing to do - when the corner minimum proves it is not. */
UPC_D bool cell_gradient(const DevEnv& env, const CellRef& c, const Corner4& a, const Corner4& b, double radius, double limit, float clear_above,
                         double* dist, double* grad)
{
    if (!c.inb)
    {
        *dist = (double)env.sdf_oob;
        grad[0] = 0.0; grad[1] = 0.0; grad[2] = 0.0;
        return !((((double)env.sdf_oob - radius) - limit) > UPC_CULL_MARGIN);
    }
    if (fmin4(a, b) > clear_above) return false;
    const double fx = c.fx, fy = c.fy, fz = c.fz;
    const double v000 = (double)a.x, v010 = (double)a.z;
    const double v100 = (double)b.x, v110 = (double)b.z;
    if (UPC_PLANAR_SHORTCUT && env.planar)
    {
        /* z corners equal: c00 = v000 etc. and ddz = 0 exactly; fma(G, 0, x) = x, so the gradient keeps its two in-plane terms */
        const double dy0 = v010 - v000, dy1 = v110 - v100;
        const double c0 = upc_fma(fy, dy0, v000);
        const double c1 = upc_fma(fy, dy1, v100);
        const double ddx = c1 - c0;
        *dist = upc_fma(fx, ddx, c0);
        const double ddy = upc_fma(fx, dy1 - dy0, dy0);
        const double g0 = ddx * env.inv_res, g1 = ddy * env.inv_res, g2 = 0.0 * env.inv_res;
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            grad[j] = upc_fma(env.G[8 + j], g2, upc_fma(env.G[4 + j], g1, env.G[j] * g0));
        }
        return true;
    }
    const double v001 = (double)a.y, v011 = (double)a.w, v101 = (double)b.y, v111 = (double)b.w;
    const double dz00 = v001 - v000, dz01 = v011 - v010, dz10 = v101 - v100, dz11 = v111 - v110;
    const double c00 = upc_fma(fz, dz00, v000);
    const double c01 = upc_fma(fz, dz01, v010);
    const double c10 = upc_fma(fz, dz10, v100);
    const double c11 = upc_fma(fz, dz11, v110);
    const double dy0 = c01 - c00, dy1 = c11 - c10;
    const double c0 = upc_fma(fy, dy0, c00);
    const double c1 = upc_fma(fy, dy1, c10);
    const double dzv0 = upc_fma(fy, dz01 - dz00, dz00);
    const double dzv1 = upc_fma(fy, dz11 - dz10, dz10);
    const double ddx = c1 - c0;
    *dist = upc_fma(fx, ddx, c0);
    const double ddy = upc_fma(fx, dy1 - dy0, dy0);
    const double ddz = upc_fma(fx, dzv1 - dzv0, dzv0);
    const double g0 = ddx * env.inv_res, g1 = ddy * env.inv_res, g2 = ddz * env.inv_res;
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        grad[j] = upc_fma(env.G[8 + j], g2, upc_fma(env.G[4 + j], g1, env.G[j] * g0));
    }
    return true;
}

/* distance and world-frame gradient */
UPC_D double sdf_distance_gradient(const DevEnv& env, const double* u, double* grad)
{
    if (!voxel_in_bounds(env, u))
    {
        grad[0] = 0.0; grad[1] = 0.0; grad[2] = 0.0;
        return (double)env.sdf_oob;
    }
    const double cx = u[0] - 0.5, cy = u[1] - 0.5, cz = u[2] - 0.5;
    const int ix = floor_to_int(cx), iy = floor_to_int(cy), iz = floor_to_int(cz);
    const double fx = cx - (double)ix, fy = cy - (double)iy, fz = cz - (double)iz;
    const int x0 = (ix < 0) ? 0 : ix, x1 = (ix + 1 > env.nx - 1) ? env.nx - 1 : ix + 1;
    const int y0 = (iy < 0) ? 0 : iy, y1 = (iy + 1 > env.ny - 1) ? env.ny - 1 : iy + 1;
    const int z0 = (iz < 0) ? 0 : iz, z1 = (iz + 1 > env.nz - 1) ? env.nz - 1 : iz + 1;
    /* grids are limited to 2^31 cells (checked in upc_create): 32-bit index arithmetic */
    const int r00 = (x0 * env.ny + y0) * env.nz;
    const int r01 = (x0 * env.ny + y1) * env.nz;
    const int r10 = (x1 * env.ny + y0) * env.nz;
    const int r11 = (x1 * env.ny + y1) * env.nz;
    const double v000 = (double)ldg(env.sdf + r00 + z0), v001 = (double)ldg(env.sdf + r00 + z1);
    const double v010 = (double)ldg(env.sdf + r01 + z0), v011 = (double)ldg(env.sdf + r01 + z1);
    const double v100 = (double)ldg(env.sdf + r10 + z0), v101 = (double)ldg(env.sdf + r10 + z1);
    const double v110 = (double)ldg(env.sdf + r11 + z0), v111 = (double)ldg(env.sdf + r11 + z1);
    const double dz00 = v001 - v000, dz01 = v011 - v010, dz10 = v101 - v100, dz11 = v111 - v110;
    const double c00 = upc_fma(fz, dz00, v000);
    const double c01 = upc_fma(fz, dz01, v010);
    const double c10 = upc_fma(fz, dz10, v100);
    const double c11 = upc_fma(fz, dz11, v110);
    const double dy0 = c01 - c00, dy1 = c11 - c10;
    const double c0 = upc_fma(fy, dy0, c00);
    const double c1 = upc_fma(fy, dy1, c10);
    const double dzv0 = upc_fma(fy, dz01 - dz00, dz00);
    const double dzv1 = upc_fma(fy, dz11 - dz10, dz10);
    const double ddx = c1 - c0;
    const double dist = upc_fma(fx, ddx, c0);
    const double ddy = upc_fma(fx, dy1 - dy0, dy0);
    const double ddz = upc_fma(fx, dzv1 - dzv0, dzv0);
    const double g0 = ddx * env.inv_res, g1 = ddy * env.inv_res, g2 = ddz * env.inv_res;
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        grad[j] = upc_fma(env.G[8 + j], g2, upc_fma(env.G[4 + j], g1, env.G[j] * g0));
    }
    return dist;
}

/* ------------------------------------------------------------------ damped normal equations */

template <int N>
UPC_SOLVE_LINKAGE void solve_damped(double (&H)[N][N], const double* g, double lambda, double* x, int n)
{
    /* (H + lambda I) x = g by an LDL^T factorisation without square roots: unit-lower L in the strict lower
     * triangle of H (the upper triangle holds the input), reciprocal pivots in rd[] - one division per column. */
    double rd[N];
#pragma unroll
    for (int j = 0; j < N; j++)
    {
        if (j < n)
        {
            double d = H[j][j] + lambda;
#pragma unroll
            for (int k = 0; k < N; k++)
                if (k < j) d = upc_fma(-(H[j][k] * H[j][k]), H[k][k], d); /* H[k][k] holds pivot d_k once column k is done */
            if (d < lambda) d = lambda;
            const double r = 1.0 / d;
#pragma unroll
            for (int i = 0; i < N; i++)
            {
                if (i > j && i < n)
                {
                    double t = H[j][i]; /* input H[i][j], read from the symmetric upper triangle */
#pragma unroll
                    for (int k = 0; k < N; k++)
                        if (k < j) t = upc_fma(-(H[i][k] * H[j][k]), H[k][k], t);
                    H[i][j] = t * r;
                }
            }
            H[j][j] = d;
            rd[j] = r;
        }
    }
    double y[N];
#pragma unroll
    for (int i = 0; i < N; i++)
    {
        if (i < n)
        {
            double s = g[i];
#pragma unroll
            for (int k = 0; k < N; k++)
                if (k < i) s = upc_fma(-H[i][k], y[k], s);
            y[i] = s;
        }
    }
#pragma unroll
    for (int ii = 0; ii < N; ii++)
    {
        const int i = N - 1 - ii;
        if (i < n)
        {
            double s = y[i] * rd[i];
#pragma unroll
            for (int k = 0; k < N; k++)
                if (k > i && k < n) s = upc_fma(-H[k][i], x[k], s);
            x[i] = s;
        }
    }
}

/* ------------------------------------------------------------------ robot models */

/*
 * Each model keeps the particle's configuration, the PID state of every axis and whatever pose cache it needs.
 * Common interface used by the simulation program:
 *   load(cfg_soa, stride, i) / store(...)   boundary layout <-> registers
 *   reset_controllers()
 *   control(robot, target, dt, draws, u)    GenerateControlAction(target, dt, rng)
 *   apply(robot, input, draws|nullptr)      ApplyControlInput(input[, rng])
 *   frame(robot, l, T)                      current transform of link l
 *   jacobian(robot, l, local, world, J)     3 x dof translational point Jacobian
 *   distance_to(robot, target)              ComputeConfigurationDistance(config, target)
 */
struct Se2Model
{
    static constexpr int DOF = 3;
    static constexpr int WIDTH = 3;
    double x, y, th, sn, cs;
    double* pid_i; /* per-particle PID state (integral, last error), parked in shared memory on the GPU */
    double* pid_e;

    UPC_D int dof(const DevRobot&) const { return 3; }
    UPC_D void set_config(double nx, double ny, double nth) /* simple_robot_models.hpp:110-119 */
    {
        x = nx;
        y = ny;
        th = upc_wrap_angle(nth);
        upc_sincos(th, &sn, &cs);
    }
    UPC_D void load(const DevRobot&, const double* c, int64_t stride, int64_t i) { set_config(c[i], c[stride + i], c[2 * stride + i]); }
    UPC_D void store(const DevRobot&, double* c, int64_t stride, int64_t i) const
    {
        c[i] = x; c[stride + i] = y; c[2 * stride + i] = th;
    }
    UPC_D void get(const DevRobot&, double* c) const { c[0] = x; c[1] = y; c[2] = th; }
    UPC_D void set(const DevRobot&, const double* c) { set_config(c[0], c[1], c[2]); }
    UPC_D void reset_controllers()
    {
        for (int i = 0; i < 3; i++) { pid_i[i] = 0.0; pid_e[i] = 0.0; }
    }
    UPC_D static double distance(const DevRobot&, const double* a, const double* b) /* :424-430 */
    {
        const double dx = b[0] - a[0];
        const double dy = b[1] - a[1];
        const double dth = upc_angle_signed_distance(a[2], b[2]);
        const double trans = sqrt((dx * dx) + (dy * dy));
        return trans + fabs(dth);
    }
    UPC_D double distance_to(const DevRobot& r, const double* target) const
    {
        const double c[3] = {x, y, th};
        return distance(r, c, target);
    }
    /* ComputeConfigurationDistance(config, target) < shortcut */
    UPC_D bool reached(const DevRobot& r, const double* target, double shortcut) const { return distance_to(r, target) < shortcut; }
    /* upper bound of the body-point displacement caused by a noise-free input */
    UPC_D double motion_bound(const DevRobot& r, const double* in) const
    {
        const double c0 = actuator_clamp(r.axis[0], in[0]), c1 = actuator_clamp(r.axis[1], in[1]), c2 = actuator_clamp(r.axis[2], in[2]);
        return sqrt((c0 * c0) + (c1 * c1)) + fabs(c2) * r.reach[0];
    }
    UPC_D void control(const DevRobot& r, const double* target, double dt, const double* draws, double* u) /* :239-259 */
    {
        const double sx = x + draws[0];
        const double sy = y + draws[1];
        const double sth = th + draws[2];
        const double e0 = target[0] - sx;
        const double e1 = target[1] - sy;
        const double e2 = upc_angle_signed_distance(sth, target[2]);
        u[0] = actuator_clamp(r.axis[0], pid_step(r.axis[0], pid_i[0], pid_e[0], e0, dt));
        u[1] = actuator_clamp(r.axis[1], pid_step(r.axis[1], pid_i[1], pid_e[1], e1, dt));
        u[2] = actuator_clamp(r.axis[2], pid_step(r.axis[2], pid_i[2], pid_e[2], e2, dt));
    }
    UPC_D void apply(const DevRobot& r, const double* in, const double* draws) /* :282-308 */
    {
        double n0, n1, n2;
        if (draws)
        {
            n0 = actuator_noisy(r.axis[0], in[0], draws[0]);
            n1 = actuator_noisy(r.axis[1], in[1], draws[1]);
            n2 = actuator_noisy(r.axis[2], in[2], draws[2]);
        }
        else
        {
            n0 = actuator_clamp(r.axis[0], in[0]);
            n1 = actuator_clamp(r.axis[1], in[1]);
            n2 = actuator_clamp(r.axis[2], in[2]);
        }
        set_config(x + n0, y + n1, th + n2);
    }
    static constexpr bool HAS_FRAMES = false;
    static constexpr bool PLANAR = true; /* body points usually share one z: the pre-cull may locate the z axis once per link */
    UPC_D void frame(const DevRobot&, int, double* T) const /* :363-371 */
    {
        T[0] = cs;  T[1] = -sn; T[2] = 0.0;  T[3] = x;
        T[4] = sn;  T[5] = cs;  T[6] = 0.0;  T[7] = y;
        T[8] = 0.0; T[9] = 0.0; T[10] = 1.0; T[11] = 0.0;
    }
    UPC_D void jacobian(const DevRobot&, int, const double*, const double* world, double (&J)[3][DOF]) const /* :334-361 */
    {
        J[0][0] = 1.0; J[1][0] = 0.0; J[2][0] = 0.0;
        J[0][1] = 0.0; J[1][1] = 1.0; J[2][1] = 0.0;
        const double rx = world[0] - x;
        const double ry = world[1] - y;
        J[0][2] = -ry; J[1][2] = rx; J[2][2] = 0.0;
    }
};

struct Se3Model
{
    static constexpr int DOF = 6;
    static constexpr int WIDTH = 12;
    double T[12];
    double* pid_i; /* per-particle PID state (integral, last error), parked in shared memory on the GPU */
    double* pid_e;

    UPC_D int dof(const DevRobot&) const { return 6; }
    UPC_D void load(const DevRobot&, const double* c, int64_t stride, int64_t i)
    {
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) T[r * 4 + k] = c[(int64_t)(r * 3 + k) * stride + i];
            T[r * 4 + 3] = c[(int64_t)(9 + r) * stride + i];
        }
    }
    UPC_D void store(const DevRobot&, double* c, int64_t stride, int64_t i) const
    {
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) c[(int64_t)(r * 3 + k) * stride + i] = T[r * 4 + k];
            c[(int64_t)(9 + r) * stride + i] = T[r * 4 + 3];
        }
    }
    UPC_D void get(const DevRobot&, double* c) const
    {
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) c[r * 3 + k] = T[r * 4 + k];
            c[9 + r] = T[r * 4 + 3];
        }
    }
    UPC_D void set(const DevRobot&, const double* c)
    {
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) T[r * 4 + k] = c[r * 3 + k];
            T[r * 4 + 3] = c[9 + r];
        }
    }
    UPC_D void reset_controllers()
    {
        for (int i = 0; i < 6; i++) { pid_i[i] = 0.0; pid_e[i] = 0.0; }
    }
    /* a, b in the boundary layout (R row-major, then t) */
    UPC_D static double distance(const DevRobot&, const double* a, const double* b) /* :871-874 */
    {
        double ta[12], tb[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) { ta[r * 4 + k] = a[r * 3 + k]; tb[r * 4 + k] = b[r * 3 + k]; }
            ta[r * 4 + 3] = a[9 + r]; tb[r * 4 + 3] = b[9 + r];
        }
        return distance_xf(ta, tb);
    }
    UPC_DNI static double distance_xf(const double* ta, const double* tb)
    {
        const double dx = ta[3] - tb[3], dy = ta[7] - tb[7], dz = ta[11] - tb[11];
        const double tdist = sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
        double q1[4], q2[4];
        quat_from_xf(ta, q1);
        quat_from_xf(tb, q2);
        const double dq = fabs((((q1[0] * q2[0]) + (q1[1] * q2[1])) + (q1[2] * q2[2])) + (q1[3] * q2[3]));
        double rdist = 0.0;
        if (dq < (1.0 - 2.220446049250313e-16))
        {
            rdist = upc_acos(((2.0 * dq) * dq) - 1.0);
        }
        return (((1.0 - 0.5) * tdist) + (0.5 * rdist)) * 2.0;
    }
    UPC_D double distance_to(const DevRobot&, const double* target) const
    {
        double tb[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) tb[r * 4 + k] = target[r * 3 + k];
            tb[r * 4 + 3] = target[9 + r];
        }
        return distance_xf(T, tb);
    }
    /* ComputeConfigurationDistance(config, target) < shortcut.  The distance is |dt| + angle with angle >= 0 and
     * monotone rounding, so |dt| >= shortcut already decides "no" without the quaternion / acos work. */
    UPC_D bool reached(const DevRobot& r, const double* target, double shortcut) const
    {
        const double dx = T[3] - target[9], dy = T[7] - target[10], dz = T[11] - target[11];
        const double tdist = sqrt(((dx * dx) + (dy * dy)) + (dz * dz));
        if (!(tdist < shortcut)) return false;
        return distance_to(r, target) < shortcut;
    }
    UPC_D double motion_bound(const DevRobot& r, const double* in) const
    {
        double c[6];
#pragma unroll
        for (int i = 0; i < 6; i++) c[i] = actuator_clamp(r.axis[i], in[i]);
        const double v = sqrt(((c[0] * c[0]) + (c[1] * c[1])) + (c[2] * c[2]));
        const double w = sqrt(((c[3] * c[3]) + (c[4] * c[4])) + (c[5] * c[5]));
        return v + w * r.reach[0];
    }
    /* GetPosition(rng) :654-668; out of line: benchmarks and most planner setups run without sensor noise */
    UPC_DNI static void sense_noisy(const double* T, const double* draws, double* sensed)
    {
        double tw[6];
        se3_log(T, tw);
#pragma unroll
        for (int i = 0; i < 6; i++) tw[i] = tw[i] + draws[i];
        se3_exp(tw, sensed);
    }
    UPC_D void control(const DevRobot& r, const double* target, double dt, const double* draws, double* u) /* :685-714 */
    {
        double sensed[12];
#pragma unroll
        for (int k = 0; k < 12; k++) sensed[k] = T[k];
        /* Exp(Log(T) + 0) is T up to rounding: with no sensor noise the round trip is skipped */
        if (!r.sensor_noise_free) sense_noisy(T, draws, sensed);
        double tgt[12];
#pragma unroll
        for (int rr = 0; rr < 3; rr++)
        {
#pragma unroll
            for (int k = 0; k < 3; k++) tgt[rr * 4 + k] = target[rr * 3 + k];
            tgt[rr * 4 + 3] = target[9 + rr];
        }
        double rel[12];
        se3_between(sensed, tgt, rel);
        double err[6];
        se3_log(rel, err);
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            u[i] = actuator_clamp(r.axis[i], pid_step(r.axis[i], pid_i[i], pid_e[i], err[i], dt));
        }
    }
    UPC_D void apply(const DevRobot& r, const double* in, const double* draws) /* :746-789 */
    {
        double tw[6];
#pragma unroll
        for (int i = 0; i < 6; i++)
        {
            tw[i] = draws ? actuator_noisy(r.axis[i], in[i], draws[i]) : actuator_clamp(r.axis[i], in[i]);
        }
        double motion[12], next[12];
        se3_exp(tw, motion);
        xf_compose(T, motion, next);
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = next[k];
    }
    static constexpr bool HAS_FRAMES = false;
    static constexpr bool PLANAR = false;
    UPC_D void frame(const DevRobot&, int, double* out) const
    {
#pragma unroll
        for (int k = 0; k < 12; k++) out[k] = T[k];
    }
    UPC_D void jacobian(const DevRobot&, int, const double* local, const double*, double (&J)[3][DOF]) const /* :813-832 */
    {
        const double px = local[0], py = local[1], pz = local[2];
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            J[i][0] = T[i * 4 + 0];
            J[i][1] = T[i * 4 + 1];
            J[i][2] = T[i * 4 + 2];
            J[i][3] = upc_fma(T[i * 4 + 2], py, -(T[i * 4 + 1] * pz));
            J[i][4] = upc_fma(T[i * 4 + 0], pz, -(T[i * 4 + 2] * px));
            J[i][5] = upc_fma(T[i * 4 + 1], px, -(T[i * 4 + 0] * py));
        }
    }
};

/*
 * Linked robot: joint values and PID state live in registers/local memory; link transforms are recomputed
 * into a caller-provided frame buffer (shared memory on the GPU) by fk().
 */
/* CAP = capacity in degrees of freedom: arrays sized by it (joint values here, the resolver's normal equations in
 * upc_sim.cuh) are what the per-thread stack of the simulation program consists of; robots with dof <= 8 run the CAP = 8
 * instantiation (C3: 8.7 KB -> about a third of it per thread) */
template <int CAP>
struct LinkedModelT
{
    static constexpr int DOF = CAP;
    static constexpr int WIDTH = CAP;
    static constexpr int LINKS = (CAP <= 8) ? 9 : UPC_MAX_LINKS; /* link-frame capacity of the per-thread frame buffers */
    double q[CAP];
    double* pid_i;
    double* pid_e;
    double* frames; /* num_links * 12 doubles, current link transforms */
    static constexpr bool HAS_FRAMES = true;
    static constexpr bool PLANAR = false;
    UPC_D void frame(const DevRobot&, int l, double* out) const
    {
#pragma unroll
        for (int k = 0; k < 12; k++) out[k] = frames[l * 12 + k];
    }

    UPC_D int dof(const DevRobot& r) const { return r.dof; }
    UPC_D static double enforce_limits(const DevJoint& j, double v) /* :1077-1100 */
    {
        if (j.type == UPC_JOINT_CONTINUOUS) return upc_wrap_angle(v);
        if (v < j.lo) return j.lo;
        if (v > j.hi) return j.hi;
        return v;
    }
    UPC_D static double signed_distance(const DevJoint& j, double v1, double v2) /* :1108-1118 */
    {
        if (j.type == UPC_JOINT_CONTINUOUS) return upc_angle_signed_distance(v1, v2);
        return (v2 - v1);
    }
    /* forward kinematics into out (num_links*12), :1299-1343 */
    UPC_D void fk(const DevRobot& r, double* out) const
    {
        for (int k = 0; k < 12; k++) out[k] = r.base[k];
        for (int idx = 0; idx < r.num_joints; idx++)
        {
            const DevJoint& j = r.joints[idx];
            double parent[12], complete[12];
            for (int k = 0; k < 12; k++) parent[k] = out[j.parent * 12 + k];
            xf_compose(parent, j.T, complete);
            double* child = out + j.child * 12;
            if ((j.type == UPC_JOINT_REVOLUTE) || (j.type == UPC_JOINT_CONTINUOUS))
            {
                double s, c;
                upc_sincos(q[j.active], &s, &c);
                const double t = 1.0 - c;
                const double ax = j.axis[0], ay = j.axis[1], az = j.axis[2];
                double Rj[9];
                Rj[0] = upc_fma(t * ax, ax, c);
                Rj[1] = upc_fma(t * ax, ay, -(s * az));
                Rj[2] = upc_fma(t * ax, az, s * ay);
                Rj[3] = upc_fma(t * ay, ax, s * az);
                Rj[4] = upc_fma(t * ay, ay, c);
                Rj[5] = upc_fma(t * ay, az, -(s * ax));
                Rj[6] = upc_fma(t * az, ax, -(s * ay));
                Rj[7] = upc_fma(t * az, ay, s * ax);
                Rj[8] = upc_fma(t * az, az, c);
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
                {
#pragma unroll
                    for (int cc = 0; cc < 3; cc++)
                    {
                        child[rr * 4 + cc] = upc_fma(complete[rr * 4 + 2], Rj[6 + cc],
                                                     upc_fma(complete[rr * 4 + 1], Rj[3 + cc], complete[rr * 4 + 0] * Rj[cc]));
                    }
                    child[rr * 4 + 3] = complete[rr * 4 + 3];
                }
            }
            else if (j.type == UPC_JOINT_PRISMATIC)
            {
                const double v = q[j.active];
                const double d0 = j.axis[0] * v, d1 = j.axis[1] * v, d2 = j.axis[2] * v;
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
                {
                    child[rr * 4 + 0] = complete[rr * 4 + 0];
                    child[rr * 4 + 1] = complete[rr * 4 + 1];
                    child[rr * 4 + 2] = complete[rr * 4 + 2];
                    child[rr * 4 + 3] = upc_fma(complete[rr * 4 + 2], d2,
                                                upc_fma(complete[rr * 4 + 1], d1, upc_fma(complete[rr * 4 + 0], d0, complete[rr * 4 + 3])));
                }
            }
            else
            {
                for (int k = 0; k < 12; k++) child[k] = complete[k];
            }
        }
    }
    UPC_D void set_values(const DevRobot& r, const double* c) /* :1345-1369 */
    {
        for (int idx = 0; idx < r.num_joints; idx++)
        {
            const DevJoint& j = r.joints[idx];
            if (j.type == UPC_JOINT_FIXED) continue;
            q[j.active] = enforce_limits(j, c[j.active]);
        }
    }
    UPC_D void load(const DevRobot& r, const double* c, int64_t stride, int64_t i)
    {
        double v[CAP];
        for (int a = 0; a < r.dof; a++) v[a] = c[(int64_t)a * stride + i];
        set_values(r, v);
    }
    UPC_D void store(const DevRobot& r, double* c, int64_t stride, int64_t i) const
    {
        for (int a = 0; a < r.dof; a++) c[(int64_t)a * stride + i] = q[a];
    }
    UPC_D void get(const DevRobot& r, double* c) const
    {
        for (int a = 0; a < r.dof; a++) c[a] = q[a];
    }
    UPC_D void set(const DevRobot& r, const double* c) { set_values(r, c); }
    UPC_D void reset_controllers()
    {
        for (int i = 0; i < CAP; i++) { pid_i[i] = 0.0; pid_e[i] = 0.0; }
    }
    UPC_D static double distance(const DevRobot& r, const double* a, const double* b) /* :2191-2217 */
    {
        double sum = 0.0;
        for (int idx = 0; idx < r.num_joints; idx++)
        {
            const DevJoint& j = r.joints[idx];
            if (j.type == UPC_JOINT_FIXED) continue;
            const double raw = signed_distance(j, a[j.active], b[j.active]);
            const double wd = raw * r.weights[j.active];
            sum = sum + (wd * wd);
        }
        return sqrt(sum);
    }
    UPC_D double distance_to(const DevRobot& r, const double* target) const { return distance(r, q, target); }
    UPC_D bool reached(const DevRobot& r, const double* target, double shortcut) const { return distance_to(r, target) < shortcut; }
    UPC_D double motion_bound(const DevRobot& r, const double* in) const
    {
        double b = 0.0;
        for (int a = 0; a < r.dof; a++) b = b + fabs(actuator_clamp(r.axis[a], in[a])) * r.reach[a];
        return b;
    }
    UPC_D void control(const DevRobot& r, const double* target, double dt, const double* draws, double* u) /* :1823-1851 */
    {
        for (int idx = 0; idx < r.num_joints; idx++)
        {
            const DevJoint& j = r.joints[idx];
            if (j.type == UPC_JOINT_FIXED) continue;
            const int a = j.active;
            const double sensed = enforce_limits(j, q[a] + draws[a]); /* :1784-1806 */
            const double err = signed_distance(j, sensed, target[a]);
            u[a] = actuator_clamp(r.axis[a], pid_step(r.axis[a], pid_i[a], pid_e[a], err, dt));
        }
    }
    UPC_D void apply(const DevRobot& r, const double* in, const double* draws) /* :1882-1940 */
    {
        for (int idx = 0; idx < r.num_joints; idx++)
        {
            const DevJoint& j = r.joints[idx];
            if (j.type == UPC_JOINT_FIXED) continue;
            const int a = j.active;
            const double n = draws ? actuator_noisy(r.axis[a], in[a], draws[a]) : actuator_clamp(r.axis[a], in[a]);
            q[a] = enforce_limits(j, q[a] + n);
        }
    }
    /* :2011-2084 using the frames computed by the last fk() into `frames` */
    UPC_D void jacobian(const DevRobot& r, int link, const double*, const double* world, double (&J)[3][DOF]) const
    {
        for (int rr = 0; rr < 3; rr++)
            for (int a = 0; a < r.dof; a++) J[rr][a] = 0.0;
        int current = link;
        while (current > 0)
        {
            const DevJoint& j = r.joints[r.parent_joint_of_link[current]];
            const double* Tc = frames + j.child * 12;
            if ((j.type == UPC_JOINT_REVOLUTE) || (j.type == UPC_JOINT_CONTINUOUS))
            {
                double aw[3];
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
                {
                    aw[rr] = upc_fma(Tc[rr * 4 + 2], j.axis[2], upc_fma(Tc[rr * 4 + 1], j.axis[1], Tc[rr * 4 + 0] * j.axis[0]));
                }
                const double rel[3] = {world[0] - Tc[3], world[1] - Tc[7], world[2] - Tc[11]};
                double c[3];
                cross3(aw, rel, c);
                J[0][j.active] = c[0]; J[1][j.active] = c[1]; J[2][j.active] = c[2];
            }
            else if (j.type == UPC_JOINT_PRISMATIC)
            {
                double aw[3];
                xf_point(Tc, j.axis[0], j.axis[1], j.axis[2], aw); /* full isometry, :2064 */
                J[0][j.active] = aw[0]; J[1][j.active] = aw[1]; J[2][j.active] = aw[2];
            }
            current = j.parent;
        }
    }
};

using LinkedModel = LinkedModelT<UPC_MAX_DOF>;
using LinkedModelSmall = LinkedModelT<8>;

} /* namespace upc */

#endif /* UPC_DEVICE_CUH */
