/*
 * upc_host_prep.h - host-side precomputation shared by the C ABI (upc_capi.cu) and the CPU emulation of the
 * device program (tests/emulation): the cell-minimum grid behind the SDF cull and the static reach bounds
 * behind the motion bound.  Both only ever make the kernels skip work whose outcome is
 * already certain; they never change a result.
 */
#ifndef UPC_HOST_PREP_H
#define UPC_HOST_PREP_H

#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "upc_b200.h"

namespace upc
{
/* factors of the counter-based noise: what a standard truncated-normal draw of the axis is multiplied by;
 * the same scaling the tape generator applies (sensor: z * max_sensor_noise / 2, actuator: z / 2), 0 for a noise-free axis */
inline double sensor_noise_scale(const upc_axis_params& a) { const double b = fabs(a.max_sensor_noise); return (b > 0.0) ? (b * 0.5) : 0.0; }
inline double actuator_noise_scale(const upc_axis_params& a) { return (fabs(a.max_actuator_noise) > 0.0) ? 0.5 : 0.0; }

/* out[(ix*(ny+1)+iy)*(nz+1)+iz] = min of the 8 border-clamped SDF corners of the interpolation cell whose low
 * corner is (ix-1, iy-1, iz-1) */
inline void build_cellmin(const float* sdf, int nx, int ny, int nz, float* out)
{
#pragma omp parallel for schedule(static)
    for (int ix = 0; ix <= nx; ix++)
    {
        const int x0 = std::max(ix - 1, 0), x1 = std::min(ix, nx - 1);
        for (int iy = 0; iy <= ny; iy++)
        {
            const int y0 = std::max(iy - 1, 0), y1 = std::min(iy, ny - 1);
            for (int iz = 0; iz <= nz; iz++)
            {
                const int z0 = std::max(iz - 1, 0), z1 = std::min(iz, nz - 1);
                float m = sdf[((int64_t)x0 * ny + y0) * nz + z0];
                m = std::min(m, sdf[((int64_t)x0 * ny + y0) * nz + z1]);
                m = std::min(m, sdf[((int64_t)x0 * ny + y1) * nz + z0]);
                m = std::min(m, sdf[((int64_t)x0 * ny + y1) * nz + z1]);
                m = std::min(m, sdf[((int64_t)x1 * ny + y0) * nz + z0]);
                m = std::min(m, sdf[((int64_t)x1 * ny + y0) * nz + z1]);
                m = std::min(m, sdf[((int64_t)x1 * ny + y1) * nz + z0]);
                m = std::min(m, sdf[((int64_t)x1 * ny + y1) * nz + z1]);
                out[((int64_t)ix * (ny + 1) + iy) * (nz + 1) + iz] = m;
            }
        }
    }
}

/*
 * Corner-cell table: for every trilinear interpolation cell (low corner i0 in [-1, n-1] per axis, stored at
 * i0 + 1) the 8 border-clamped SDF corners as 8 consecutive floats v000 v001 v010 v011 v100 v101 v110 v111 (x y z bits):
 * one aligned 32-byte sector per lookup instead of 8 scattered loads.  8x the SDF's size, so only built while it stays
 * L2-resident (upc_create decides).
 */
inline void build_corner_cells(const float* sdf, int nx, int ny, int nz, float* out)
{
#pragma omp parallel for schedule(static)
    for (int ix = 0; ix <= nx; ix++)
    {
        const int x0 = std::max(ix - 1, 0), x1 = std::min(ix, nx - 1);
        for (int iy = 0; iy <= ny; iy++)
        {
            const int y0 = std::max(iy - 1, 0), y1 = std::min(iy, ny - 1);
            for (int iz = 0; iz <= nz; iz++)
            {
                const int z0 = std::max(iz - 1, 0), z1 = std::min(iz, nz - 1);
                float* o = out + (((int64_t)ix * (ny + 1) + iy) * (nz + 1) + iz) * 8;
                o[0] = sdf[((int64_t)x0 * ny + y0) * nz + z0];
                o[1] = sdf[((int64_t)x0 * ny + y0) * nz + z1];
                o[2] = sdf[((int64_t)x0 * ny + y1) * nz + z0];
                o[3] = sdf[((int64_t)x0 * ny + y1) * nz + z1];
                o[4] = sdf[((int64_t)x1 * ny + y0) * nz + z0];
                o[5] = sdf[((int64_t)x1 * ny + y0) * nz + z1];
                o[6] = sdf[((int64_t)x1 * ny + y1) * nz + z0];
                o[7] = sdf[((int64_t)x1 * ny + y1) * nz + z1];
            }
        }
    }
}

/*
 * Single-precision pre-cull: bound of |float cell coordinate - FP64 cell coordinate|.
 * c_i = t_i + sum_j V_ij p_j with V = (grid_from_world o link) / res.  Rounding V, p, t to float and three fused
 * multiply-adds perturb c_i by at most eps * (4 |t_i| + 5 sum_j |V_ij p_j|), eps = 2^-24.  With rigid transforms
 * sum_j |V_ij p_j| <= |p| / res, and a point the cull accepts lies inside the grid, so |t_i| <= n_i + |p| / res:
 * the error is below 5 eps S, S = max(n) + 1 + 2 max|p| / res.  The value returned is 16 eps S (3x head-room, which
 * also absorbs the ~1e-13 rounding of the FP64 coordinate itself and transforms that are rigid only to 1e-6), or 0 -
 * pre-cull off - when grid_from_world is not rigid or the guard would eat a noticeable part of a cell.
 */
inline float compute_cull_delta(int nx, int ny, int nz, double res, const double* grid_from_world, const double* points, int num_points)
{
    for (int i = 0; i < 3; i++)
    {
        const double* row = grid_from_world + 4 * i;
        const double norm = sqrt(row[0] * row[0] + row[1] * row[1] + row[2] * row[2]);
        if (!(fabs(norm - 1.0) <= 1.0e-6)) return 0.0f;
    }
    double pmax = 0.0;
    for (int k = 0; k < num_points; k++)
    {
        const double* p = points + 4 * k;
        pmax = std::max(pmax, sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
    }
    const double S = (double)std::max(nx, std::max(ny, nz)) + 1.0 + 2.0 * pmax / res;
    const double delta = S * (16.0 / 16777216.0);
    if (!(delta <= 1.0 / 64.0)) return 0.0f;
    return (float)(delta * 1.000001);
}

/*
 * Link-level cull: per link the centre and half extents of the bounding box of its body points in the link frame,
 * and a cell-minimum grid dilated by `radius` cells - dilated[c] = min of cellmin over the cube of cells within `radius` of c
 * (clamped at the grid border) - so that ONE lookup at the cell of a link's box centre bounds the SDF at every point of the
 * link from below.  radius must cover the farthest a body point's cell can lie from the centre's cell: ceil(half diagonal /
 * resolution + 1.5), maximised over the links.  Returns 0 (cull off) when the window would be larger than 32 cells or the grid
 * transform is not rigid.
 */
inline int compute_link_boxes(const upc_robot_desc* r, double res, const double* grid_from_world, double (*center)[3], double (*half)[3])
{
    for (int i = 0; i < 3; i++)
    {
        const double* row = grid_from_world + 4 * i;
        const double norm = sqrt(row[0] * row[0] + row[1] * row[1] + row[2] * row[2]);
        if (!(fabs(norm - 1.0) <= 1.0e-6)) return 0;
    }
    double worst = 0.0;
    for (int l = 0; l < r->num_links; l++)
    {
        double lo[3] = {0.0, 0.0, 0.0}, hi[3] = {0.0, 0.0, 0.0};
        for (int k = r->link_point_offset[l]; k < r->link_point_offset[l + 1]; k++)
        {
            const double* p = r->points + 4 * k;
            for (int a = 0; a < 3; a++)
            {
                if (k == r->link_point_offset[l] || p[a] < lo[a]) lo[a] = p[a];
                if (k == r->link_point_offset[l] || p[a] > hi[a]) hi[a] = p[a];
            }
        }
        double diag2 = 0.0;
        for (int a = 0; a < 3; a++)
        {
            center[l][a] = 0.5 * (lo[a] + hi[a]);
            /* a hair more than half the span, so that rounding of the centre never leaves a point outside the box */
            half[l][a] = (0.5 * (hi[a] - lo[a])) * (1.0 + 1.0e-12) + 1.0e-12;
            diag2 += half[l][a] * half[l][a];
        }
        worst = std::max(worst, sqrt(diag2));
    }
    const double cells = ceil(worst / res + 1.5 + 1.0e-3);
    if (!(cells <= 32.0)) return 0;
    return (int)cells;
}

/*
 * Group-level cull: between the link-level test and the per-point passes.  The one-lane programs walk a link's points
 * in chunks of 32 consecutive points; every chunk is split here into up to four spatially compact groups (two rounds of bisection
 * along the longest axis of the bounding box, at the median), each with the centre / half extents of its bounding box in the link
 * frame and the mask of its points inside the chunk.  A second cell-minimum grid, dilated by the radius that covers the LARGEST
 * GROUP instead of the largest link, lets one lookup at a group's centre clear all of its points - by the inequality of the
 * link-level cull (compute_link_boxes).  Returns the dilation radius in cells, 0 = off (non-rigid grid transform, window above 32
 * cells, or no link with more than one group).
 */
struct HostPointGroup
{
    double c[3];
    double h[3];
    unsigned mask;
    unsigned pad;
};
constexpr int kGroupsPerChunk = 4;

inline int compute_point_groups(const upc_robot_desc* r, double res, const double* grid_from_world, int* link_chunk_off, std::vector<HostPointGroup>* groups)
{
    groups->clear();
    for (int l = 0; l <= r->num_links; l++) link_chunk_off[l] = 0;
    for (int i = 0; i < 3; i++)
    {
        const double* row = grid_from_world + 4 * i;
        const double norm = sqrt(row[0] * row[0] + row[1] * row[1] + row[2] * row[2]);
        if (!(fabs(norm - 1.0) <= 1.0e-6)) return 0;
    }
    double worst = 0.0;
    bool any_split = false;
    int chunks = 0;
    for (int l = 0; l < r->num_links; l++)
    {
        link_chunk_off[l] = chunks;
        const int beg = r->link_point_offset[l], end = r->link_point_offset[l + 1];
        for (int chunk = beg; chunk < end; chunk += 32)
        {
            const int lim = std::min(end, chunk + 32);
            std::vector<std::vector<int>> parts(1);
            for (int k = chunk; k < lim; k++) parts[0].push_back(k);
            for (int round = 0; round < 2; round++)
            {
                std::vector<std::vector<int>> next;
                for (auto& part : parts)
                {
                    if (part.size() <= 4)
                    {
                        next.push_back(part);
                        continue;
                    }
                    double lo[3], hi[3];
                    for (int a = 0; a < 3; a++) lo[a] = hi[a] = r->points[4 * part[0] + a];
                    for (int k : part)
                        for (int a = 0; a < 3; a++)
                        {
                            lo[a] = std::min(lo[a], r->points[4 * k + a]);
                            hi[a] = std::max(hi[a], r->points[4 * k + a]);
                        }
                    int axis = 0;
                    for (int a = 1; a < 3; a++)
                        if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
                    std::stable_sort(part.begin(), part.end(), [&](int x, int y) { return r->points[4 * x + axis] < r->points[4 * y + axis]; });
                    const size_t half = part.size() / 2;
                    next.emplace_back(part.begin(), part.begin() + (long)half);
                    next.emplace_back(part.begin() + (long)half, part.end());
                }
                parts.swap(next);
            }
            if (parts.size() > 1) any_split = true;
            for (int gidx = 0; gidx < kGroupsPerChunk; gidx++)
            {
                HostPointGroup g = {};
                if (gidx < (int)parts.size())
                {
                    const std::vector<int>& part = parts[(size_t)gidx];
                    double lo[3], hi[3];
                    for (int a = 0; a < 3; a++) lo[a] = hi[a] = r->points[4 * part[0] + a];
                    for (int k : part)
                    {
                        g.mask |= 1u << (k - chunk);
                        for (int a = 0; a < 3; a++)
                        {
                            lo[a] = std::min(lo[a], r->points[4 * k + a]);
                            hi[a] = std::max(hi[a], r->points[4 * k + a]);
                        }
                    }
                    double diag2 = 0.0;
                    for (int a = 0; a < 3; a++)
                    {
                        g.c[a] = 0.5 * (lo[a] + hi[a]);
                        g.h[a] = (0.5 * (hi[a] - lo[a])) * (1.0 + 1.0e-12) + 1.0e-12;
                        diag2 += g.h[a] * g.h[a];
                    }
                    worst = std::max(worst, sqrt(diag2));
                }
                groups->push_back(g);
            }
            chunks++;
        }
    }
    link_chunk_off[r->num_links] = chunks;
    if (!any_split) return 0;
    const double cells = ceil(worst / res + 1.5 + 1.0e-3);
    if (!(cells <= 32.0)) return 0;
    return (int)cells;
}

inline void build_dilated_cellmin(const float* cellmin, int nx, int ny, int nz, int radius, float* out)
{
    /* separable minimum filter over the (nx+1)(ny+1)(nz+1) cell-minimum grid: z, then y, then x */
    const int64_t X = nx + 1, Y = ny + 1, Z = nz + 1;
    std::vector<float> tmp((size_t)(X * Y * Z));
#pragma omp parallel for schedule(static)
    for (int64_t xy = 0; xy < X * Y; xy++)
    {
        const float* src = cellmin + xy * Z;
        float* dst = out + xy * Z;
        for (int64_t z = 0; z < Z; z++)
        {
            float m = src[z];
            for (int64_t d = std::max<int64_t>(0, z - radius); d <= std::min<int64_t>(Z - 1, z + radius); d++) m = std::min(m, src[d]);
            dst[z] = m;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t x = 0; x < X; x++)
        for (int64_t y = 0; y < Y; y++)
            for (int64_t z = 0; z < Z; z++)
            {
                float m = out[(x * Y + y) * Z + z];
                for (int64_t d = std::max<int64_t>(0, y - radius); d <= std::min<int64_t>(Y - 1, y + radius); d++) m = std::min(m, out[(x * Y + d) * Z + z]);
                tmp[(size_t)((x * Y + y) * Z + z)] = m;
            }
#pragma omp parallel for schedule(static)
    for (int64_t x = 0; x < X; x++)
        for (int64_t yz = 0; yz < Y * Z; yz++)
        {
            float m = tmp[(size_t)(x * Y * Z + yz)];
            for (int64_t d = std::max<int64_t>(0, x - radius); d <= std::min<int64_t>(X - 1, x + radius); d++) m = std::min(m, tmp[(size_t)(d * Y * Z + yz)]);
            out[x * Y * Z + yz] = m;
        }
}

/* device layout of the body points: num_points * 4 doubles (x, y, z, radius) followed by num_points * 8 floats
 * (x, x, y, y, z, z, 0, 0): every coordinate twice, the operand shape of the packed single-precision pre-cull (upc_device.cuh: CullPoint) */
constexpr int kPointDoubles = 8; /* doubles of device storage per body point */
inline std::vector<double> points_with_float_copy(const double* points, int num_points)
{
    std::vector<double> out((size_t)num_points * kPointDoubles, 0.0);
    memcpy(out.data(), points, (size_t)num_points * 4 * sizeof(double));
    float* f = reinterpret_cast<float*>(out.data() + (size_t)num_points * 4);
    for (int k = 0; k < num_points; k++)
    {
        for (int a = 0; a < 3; a++)
        {
            f[8 * k + 2 * a + 0] = (float)points[4 * k + a];
            f[8 * k + 2 * a + 1] = (float)points[4 * k + a];
        }
        f[8 * k + 6] = 0.0f;
        f[8 * k + 7] = 0.0f;
    }
    return out;
}

/* links whose body points all share one (single-precision) z coordinate - the planar robots: the pre-cull then handles the z axis
 * once per link instead of once per point (upc_device.cuh: cull_frame) */
inline void compute_link_flat_z(const upc_robot_desc* r, int* flat, float* z)
{
    for (int l = 0; l < r->num_links; l++)
    {
        const int beg = r->link_point_offset[l], end = r->link_point_offset[l + 1];
        flat[l] = (end > beg) ? 1 : 0;
        z[l] = (end > beg) ? (float)r->points[4 * beg + 2] : 0.0f;
        for (int k = beg; k < end; k++)
            if ((float)r->points[4 * k + 2] != z[l]) flat[l] = 0;
    }
}

/*
 * reach[a]: an upper bound of |d(point)/d(axis a)| over all body points and all configurations.
 * SE2 / SE3: reach[0] = largest |p| of the body points (rotation about the body origin).
 * Linked: for active joint a, the largest distance any point of a descendant link can have from the joint's
 * frame origin: sum over the joints further down the chain of (|translation of the joint transform| + the
 * largest prismatic extension) plus the largest |p| of the link; prismatic joints move points by |axis| per unit.
 */
inline void compute_reach(const upc_robot_desc* r, double* reach)
{
    for (int a = 0; a < UPC_MAX_DOF; a++) reach[a] = 0.0;
    std::vector<double> link_radius((size_t)r->num_links, 0.0);
    for (int l = 0; l < r->num_links; l++)
    {
        for (int k = r->link_point_offset[l]; k < r->link_point_offset[l + 1]; k++)
        {
            const double* p = r->points + 4 * k;
            link_radius[(size_t)l] = std::max(link_radius[(size_t)l], sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
        }
    }
    if (r->kind != UPC_ROBOT_LINKED)
    {
        reach[0] = link_radius[0];
        return;
    }
    /* offset_below[j]: how far the child-link frame origin of joint j can be from joint j's own frame origin is 0
     * (revolute) or its prismatic extension; distances accumulate down the tree */
    const int nj = r->num_joints;
    std::vector<double> extent((size_t)r->num_links, 0.0); /* largest distance of any descendant point from the link's frame origin */
    for (int l = 0; l < r->num_links; l++) extent[(size_t)l] = link_radius[(size_t)l];
    for (int j = nj - 1; j >= 0; j--) /* parents precede children, so walking backwards sees children first */
    {
        const upc_joint_desc& jd = r->joints[j];
        const double* t = jd.transform;
        double hop = sqrt(t[3] * t[3] + t[7] * t[7] + t[11] * t[11]);
        if (jd.type == UPC_JOINT_PRISMATIC)
        {
            const double an = sqrt(jd.axis[0] * jd.axis[0] + jd.axis[1] * jd.axis[1] + jd.axis[2] * jd.axis[2]);
            hop += an * std::max(fabs(jd.lower_limit), fabs(jd.upper_limit));
        }
        extent[(size_t)jd.parent_link] = std::max(extent[(size_t)jd.parent_link], hop + extent[(size_t)jd.child_link]);
    }
    int active = 0;
    for (int j = 0; j < nj; j++)
    {
        const upc_joint_desc& jd = r->joints[j];
        if (jd.type == UPC_JOINT_FIXED) continue;
        if (jd.type == UPC_JOINT_PRISMATIC)
        {
            reach[active] = sqrt(jd.axis[0] * jd.axis[0] + jd.axis[1] * jd.axis[1] + jd.axis[2] * jd.axis[2]);
        }
        else
        {
            const double an = sqrt(jd.axis[0] * jd.axis[0] + jd.axis[1] * jd.axis[1] + jd.axis[2] * jd.axis[2]);
            reach[active] = std::max(an, 1.0) * extent[(size_t)jd.child_link];
        }
        active++;
    }
}
} /* namespace upc */

#endif
