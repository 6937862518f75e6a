/*
 * upc_sim.cuh - the per-particle forward-simulation program, shared by the CUDA kernel
 * (G cooperating lanes per particle) and the CPU emulation used by tests (G = 1).
 *
 * Replaces the body of a concrete SimulatorInterface::ForwardSimulateRobots
 * (include/uncertainty_planning_core/simple_simulator_interface.hpp:623; per-step building blocks at
 * simple_robot_models.hpp:239-259, 282-295 etc.).
 *
 * Lane cooperation: every lane of a group holds an identical copy of the particle state and executes the
 * state update redundantly (a warp instruction costs the same for 1 or 32 active lanes); the body points of
 * each link are dealt round-robin to the lanes, which is where the SDF gathers - the dominant cost - are
 * parallelised.  Reductions over points are either order-free (any / max) or replayed in ascending point
 * order by every lane from ballot masks and shuffles, so results do not depend on G and equal the
 * sequential oracle bit for bit.
 */
#ifndef UPC_SIM_CUH
#define UPC_SIM_CUH

#include "upc_device.cuh"
#include "upc_noise.cuh"

namespace upc
{
template <int G>
struct Group
{
    unsigned mask; /* lanes of this group inside the warp */
    int base;      /* first lane of the group */
    int sub;       /* this lane's rank in the group */

    UPC_D bool any(bool p) const
    {
#if defined(__CUDA_ARCH__)
        if (G > 1) return (__ballot_sync(mask, p) & mask) != 0u;
#endif
        return p;
    }
    /* group-relative ballot: bit k = predicate of lane k of the group */
    UPC_D unsigned ballot(bool p) const
    {
#if defined(__CUDA_ARCH__)
        if (G > 1) return (__ballot_sync(mask, p) & mask) >> base;
#endif
        return p ? 1u : 0u;
    }
    UPC_D double bcast(double v, int src) const
    {
#if defined(__CUDA_ARCH__)
        if (G > 1) return __shfl_sync(mask, v, base + src);
#endif
        (void)src;
        return v;
    }
    UPC_D double max(double v) const
    {
#if defined(__CUDA_ARCH__)
        if (G > 1)
        {
            if constexpr ((G & (G - 1)) == 0)
            {
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1)
                {
                    const double o = __shfl_xor_sync(mask, v, off);
                    v = upc_max(v, o);
                }
            }
            else
            {
                /* group sizes that are not powers of two: every lane reads every lane (max is order-free) */
                double best = v;
#pragma unroll
                for (int t = 0; t < G; t++) best = upc_max(best, __shfl_sync(mask, v, base + t));
                v = best;
            }
        }
#endif
        return v;
    }
    UPC_D void sync() const
    {
#if defined(__CUDA_ARCH__)
        if (G > 1) __syncwarp(mask);
#endif
    }
};

/*
 * Independent SDF lookups a lane keeps in flight per trip of its point loops.  More lookups hide more load latency but
 * cost registers, and this program is register-bound (168 registers with three CTAs per SM and still spilling), while
 * 91 % of the lookups hit L1: measured on C2 (4 lanes), 4 / 2 / 8 (exact check / resolver points / pre-cull) ran 2.43 ms,
 * 1 / 1 / 8 2.15 ms, 1 / 1 / 4 2.10 ms, 4 / 4 / 8 3.22 ms (profiles/README.md).
 */
#ifndef UPC_CHECK_INFLIGHT
#define UPC_CHECK_INFLIGHT 1
#endif
#ifndef UPC_RES_INFLIGHT
#define UPC_RES_INFLIGHT 1
#endif
/* unroll factor of the conservative-test loops of the plain-grid path (grids too big for the corner-cell table) */
#ifndef UPC_PLAIN_UNROLL
#define UPC_PLAIN_UNROLL 1
#endif
#if UPC_PLAIN_UNROLL == 1
#define UPC_PLAIN_UNROLL_PRAGMA _Pragma("unroll 1")
#elif UPC_PLAIN_UNROLL == 2
#define UPC_PLAIN_UNROLL_PRAGMA _Pragma("unroll 2")
#else
#define UPC_PLAIN_UNROLL_PRAGMA _Pragma("unroll 4")
#endif
/* 1: the resolver's point pass runs the single-precision pre-cull first (2..8 lanes, links with >= 8 points per lane) */
#ifndef UPC_RES_PRECULL
#define UPC_RES_PRECULL 1
#endif
#ifndef UPC_RES_PRECULL_MIN_LANES
#define UPC_RES_PRECULL_MIN_LANES 2
#endif
/*
 * Linkage of the big helpers.  They used to be out-of-line calls (one copy of the code, smaller loop body); every call then
 * saves and restores the caller's registers through local memory and passes the model through memory.  Inlined, the
 * SE(3) program drops from 508 to ~ 250 B of spills and a 1.8 to a 1.2 KB stack frame: C2 2.03 -> 1.85 ms, C1 0.416 -> 0.360 ms,
 * C3 23.0 -> 19.3 ms (profiles/README.md).  -DUPC_RESOLVER_OUT_OF_LINE / -DUPC_ESTIMATE_OUT_OF_LINE / -DUPC_SOLVE_OUT_OF_LINE
 * bring the calls back.
 */
#ifndef UPC_RESOLVER_OUT_OF_LINE
#define UPC_RESOLVER_LINKAGE UPC_D
#else
#define UPC_RESOLVER_LINKAGE UPC_DNI
#endif
#ifndef UPC_ESTIMATE_OUT_OF_LINE
#define UPC_ESTIMATE_LINKAGE UPC_D
#else
#define UPC_ESTIMATE_LINKAGE UPC_DNI
#endif
/* the resolver's point pass runs the single-precision pre-cull from this many lanes per particle on: 2 for the SE(3) / linked
 * programs (the one-lane big-batch program lost 3 % carrying it), 1 for the small SE(2) program, whose one-lane form is the planner-batch
 * workhorse (BASELINE config 5: a third of its time was the exact evaluation of all 64 points per resolver pass) */
#ifndef UPC_RES_PRECULL_SE2_ONE_LANE
#define UPC_RES_PRECULL_SE2_ONE_LANE 1
#endif
template <class M>
UPC_D constexpr int resolver_precull_min_lanes()
{
    return (UPC_RES_PRECULL_SE2_ONE_LANE && !M::HAS_FRAMES && M::DOF == 3) ? 1 : UPC_RES_PRECULL_MIN_LANES;
}
constexpr int kCheckInFlight = UPC_CHECK_INFLIGHT;
constexpr int kResolverInFlight = UPC_RES_INFLIGHT;

/* 1: the warps of a CTA enter every controller interval together (one __syncthreads per interval).  Measured on B200
 * (profiles/README.md): C5 25.4 -> 21.7 ms (20.9 ms with 256-thread CTAs for the one-lane SE(2) program), C2 1.98 -> 1.80 ms,
 * C3 14.5 -> 13.1 ms; a second barrier per apply / check / resolve action inside the interval: 21.3 ms, not kept. */
#ifndef UPC_STEP_BARRIER
#define UPC_STEP_BARRIER 1
#endif
/* one-lane programs accumulate a contributing point as soon as it has been evaluated (ascending order either way) instead of
 * parking its correction vector in a per-thread buffer for the replay phase: no dynamically indexed local array */
#ifndef UPC_STEP_BARRIER_EVERY
#define UPC_STEP_BARRIER_EVERY 1
#endif
#ifndef UPC_ONE_LANE_IMMEDIATE
#define UPC_ONE_LANE_IMMEDIATE 1
#endif

struct LocalCounters
{
    unsigned int particle_steps, resolver_iterations, failed_resolves;
#ifdef UPC_PHASE_TIMING
    long long cyc[12]; /* control, estimate, apply, check, resolver pass, reached, tape loads, -, resolver: points, replay, solve, - */
#endif
};

#if defined(UPC_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define UPC_T0 long long _t0 = clock64();
#define UPC_T1(slot) { const long long _t1 = clock64(); cnt.cyc[slot] += _t1 - _t0; _t0 = _t1; }
#else
#define UPC_T0
#define UPC_T1(slot)
#endif

/* recompute cached link frames after a configuration change (linked robots only) */
template <class M, int G>
UPC_D void refresh_frames(const DevRobot& r, M& m, const Group<G>& g)
{
    if constexpr (M::HAS_FRAMES)
    {
        if constexpr (G >= 4)
        {
            /* frames shared by the group: one lane writes them */
            g.sync();
            if (g.sub == 0) m.fk(r, m.frames);
            g.sync();
        }
        else
        {
            m.fk(r, m.frames); /* per-lane frames */
        }
    }
}

/*
 * Collision check of the current configuration.  Per link, each lane first runs the cheap conservative test
 * (one cell-minimum load) over all of its points - independent iterations, so the loads of several points are in
 * flight together - and records the undecided ones in a bit mask; only those get the exact trilinear evaluation,
 * kCheckInFlight at a time.  The result (an OR over points) does not depend on the order.
 */
/*
 * Sharing the pre-cull between the collision check and the resolver pass that follows it.  When a check finds a collision the
 * resolver's point pass runs on the SAME configuration, and its pre-cull differs from the check's only in the threshold
 * (resolve_distance >= contact_distance).  With share_threshold > 0 the check therefore runs its link-level and per-point culls at
 * that larger threshold - a superset of the points it needs itself, the exact evaluation still decides at `threshold` - and leaves
 * the undecided masks (one word per chunk of 32 * G points, links in order) in `masks`; they are exactly the masks the resolver's
 * own cull would compute, so the resolver skips it.  *masks_state = number of words written, -1 = not recorded (a link took a path
 * without the pre-cull, or more than kMaxMaskChunks words).
 */
constexpr int kMaxMaskChunks = 8;
#ifndef UPC_SHARE_MASKS
#define UPC_SHARE_MASKS 1
#endif

/* the programs whose resolver pass runs the single-precision pre-cull (and can therefore take it from the check before it) */
template <class M, int G>
UPC_D constexpr bool shares_cull_masks()
{
    return UPC_SHARE_MASKS && UPC_RES_PRECULL && (G >= resolver_precull_min_lanes<M>()) && (G <= 8);
}

template <class M, int G>
UPC_D bool check_collision(const DevEnv& env, const DevRobot& r, const M& m, const Group<G>& g, const double* pts, double threshold,
                           double share_threshold = -1.0, unsigned* masks = nullptr, int mask_stride = 0, int* masks_state = nullptr)
{
    bool hit = false;
    const bool rec = (masks != nullptr) && (share_threshold >= threshold);
    bool rec_ok = rec;
    int rec_n = 0;
    for (int l = 0; l < r.num_links; l++)
    {
        const int beg = r.link_off[l], end = r.link_off[l + 1];
        if (beg == end) continue;
        double T[12], V[12];
        m.frame(r, l, T);
        voxel_from_link(env, T, V);
        /* single-precision clear threshold: (largest radius of the link + threshold + margin) rounded up - conservative for
         * every point of the link, so it never changes a result */
        const float clear_above = float_round_up((r.link_radius[l] + threshold) + UPC_CULL_MARGIN);
        /* the culls run at the shared (larger) threshold when their masks are recorded for the resolver */
        const float clear_cull = rec ? float_round_up((r.link_radius[l] + share_threshold) + UPC_CULL_MARGIN) : clear_above;
        const int link_chunks = (end - beg + 32 * G - 1) / (32 * G);
        /* link-level cull: one lookup of the dilated cell minimum clears the whole link when it is far from every obstacle */
        if (link_certainly_clear(env, r, l, V, clear_cull))
        {
            if (rec_ok)
            {
                if (rec_n + link_chunks > kMaxMaskChunks) rec_ok = false;
                else
                    for (int c = 0; c < link_chunks; c++) masks[(rec_n++) * mask_stride] = 0u;
            }
            continue;
        }
        if (env.cells != nullptr)
        {
            /* corner-cell table: one sector per point decides (and, if needed, evaluates) it */
            if ((r.cull_delta > 0.0f) && (end - beg >= 8 * G))
            {
                /* enough points per lane for full batches: single-precision pre-cull against the 4-byte cell minimum
                 * first (UPC_CULL_INFLIGHT lookups in flight), exact FP64 evaluation of the undecided points only */
                const CullFrame cf = cull_frame<M::PLANAR>(env, r, l, V, r.cull_delta, clear_cull);
                const CullPoint* ptsf = reinterpret_cast<const CullPoint*>(pts + 4 * r.num_points);
                if (rec_ok && rec_n + link_chunks > kMaxMaskChunks) rec_ok = false;
                for (int chunk = beg; chunk < end; chunk += 32 * G)
                {
                    const int first = chunk + g.sub;
                    const int lim = (end < chunk + 32 * G) ? end : (chunk + 32 * G);
                    /* group-level cull (one-lane programs): only the points of groups that are not provably clear go on */
                    unsigned group_mask = 0xffffffffu;
                    if constexpr (G == 1) group_mask = group_candidates(env, r, l, (chunk - beg) >> 5, V, clear_cull);
                    unsigned undecided = cull_undecided<M::PLANAR>(env, cf, ptsf, first, lim, G, group_mask);
                    if (rec_ok) masks[(rec_n++) * mask_stride] = undecided;
                    while (undecided)
                    {
                        int kk[kCheckInFlight];
                        int valid = 0;
#pragma unroll
                        for (int q = 0; q < kCheckInFlight; q++)
                        {
                            if (undecided)
                            {
                                kk[q] = first + lowest_bit(undecided) * G;
                                undecided &= undecided - 1u;
                                valid++;
                            }
                            else
                            {
                                kk[q] = kk[0]; /* padding repeats a valid point and is ignored below */
                            }
                        }
                        CellRef c[kCheckInFlight];
                        Corner4 ca[kCheckInFlight], cb[kCheckInFlight];
#pragma unroll
                        for (int q = 0; q < kCheckInFlight; q++)
                        {
                            double u[3];
                            xf_point(V, pts[4 * kk[q] + 0], pts[4 * kk[q] + 1], pts[4 * kk[q] + 2], u);
                            c[q] = cell_locate(env, u);
                        }
#pragma unroll
                        for (int q = 0; q < kCheckInFlight; q++)
                        {
                            ca[q] = ldg4(c[q].cell);
                            cb[q] = ldg4(c[q].cell + 4);
                        }
#pragma unroll
                        for (int q = 0; q < kCheckInFlight; q++)
                            if (q < valid) hit |= cell_below(env, c[q], ca[q], cb[q], pts[4 * kk[q] + 3], threshold, clear_above);
                    }
                }
                continue;
            }
            rec_ok = false; /* this link is checked without the pre-cull: nothing to share */
            int k = beg + g.sub;
            /* four lookups in flight per trip: locate all, load all, then evaluate */
            for (; k + 3 * G < end; k += 4 * G)
            {
                CellRef c[4];
                Corner4 ca[4], cb[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    const int kj = k + j * G;
                    double u[3];
                    xf_point(V, pts[4 * kj + 0], pts[4 * kj + 1], pts[4 * kj + 2], u);
                    c[j] = cell_locate(env, u);
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    ca[j] = ldg4(c[j].cell);
                    cb[j] = ldg4(c[j].cell + 4);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) hit |= cell_below(env, c[j], ca[j], cb[j], pts[4 * (k + j * G) + 3], threshold, clear_above);
            }
            for (; k < end; k += G)
            {
                double u[3];
                xf_point(V, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], u);
                hit |= sdf_below_cells(env, u, pts[4 * k + 3], threshold, clear_above);
            }
            continue;
        }
        rec_ok = false; /* plain-grid path: its conservative test is not the resolver's pre-cull */
        for (int chunk = beg; chunk < end; chunk += 32 * G)
        {
            const int first = chunk + g.sub;
            const int lim = (end < chunk + 32 * G) ? end : (chunk + 32 * G);
            unsigned undecided = 0u;
            unsigned group_mask = 0xffffffffu;
            if constexpr (G == 1) group_mask = group_candidates(env, r, l, (chunk - beg) >> 5, V, clear_above);
UPC_PLAIN_UNROLL_PRAGMA
            for (int j = 0; j < 32; j++)
            {
                const int k = first + j * G;
                if ((k < lim) && ((group_mask >> j) & 1u))
                {
                    double u[3];
                    xf_point(V, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], u);
                    if (!sdf_certainly_clear(env, u, pts[4 * k + 3], threshold)) undecided |= (1u << j);
                }
            }
            while (undecided)
            {
#if defined(__CUDA_ARCH__)
                const int ja = __ffs((int)undecided) - 1;
#else
                const int ja = __builtin_ctz(undecided);
#endif
                undecided &= undecided - 1u;
                const int ka = first + ja * G;
                double ua[3];
                xf_point(V, pts[4 * ka + 0], pts[4 * ka + 1], pts[4 * ka + 2], ua);
                if (undecided)
                {
#if defined(__CUDA_ARCH__)
                    const int jb = __ffs((int)undecided) - 1;
#else
                    const int jb = __builtin_ctz(undecided);
#endif
                    undecided &= undecided - 1u;
                    const int kb = first + jb * G;
                    double ub[3];
                    xf_point(V, pts[4 * kb + 0], pts[4 * kb + 1], pts[4 * kb + 2], ub);
                    const double da = sdf_distance(env, ua);
                    const double db = sdf_distance(env, ub);
                    hit |= ((da - pts[4 * ka + 3]) < threshold) | ((db - pts[4 * kb + 3]) < threshold);
                }
                else
                {
                    hit |= ((sdf_distance(env, ua) - pts[4 * ka + 3]) < threshold);
                }
            }
        }
    }
    if (masks_state != nullptr) *masks_state = rec_ok ? rec_n : -1;
    return g.any(hit);
}

/* EstimateMaxControlInputWorkspaceMotion: largest body-point displacement caused by a noise-free input */
template <class M, int G>
UPC_ESTIMATE_LINKAGE double estimate_motion(const DevRobot& r, const M& m, const Group<G>& g, const double* pts, const double* input, double* scratch_frames)
{
    M moved = m;
    if constexpr (M::HAS_FRAMES) moved.frames = scratch_frames;
    moved.apply(r, input, nullptr);
    refresh_frames(r, moved, g);
    double max_sq = 0.0;
    for (int l = 0; l < r.num_links; l++)
    {
        const int beg = r.link_off[l], end = r.link_off[l + 1];
        if (beg == end) continue;
        double Ta[12], Tb[12];
        m.frame(r, l, Ta);
        moved.frame(r, l, Tb);
#pragma unroll
        for (int k = 0; k < 12; k++) Tb[k] = Tb[k] - Ta[k];
        for (int k = beg + g.sub; k < end; k += G)
        {
            double disp[3];
            xf_point(Tb, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], disp);
            max_sq = upc_max(max_sq, dot3(disp, disp));
        }
    }
    max_sq = g.max(max_sq);
    return sqrt(max_sq);
}

/*
 * One resolver pass: raw correction step x from the body points closer than resolve_distance.  Returns the number
 * of contributing points.
 *   phase A  per lane, all its points of the link: conservative test, undecided points into a bit mask;
 *   phase B  per lane, its undecided points: exact distance + gradient -> correction vector, kept in a small
 *            per-thread buffer (independent evaluations, so their latencies overlap);
 *   phase C  the contributing points are replayed in ascending point order by every lane of the group
 *            (ballot + shuffles), accumulating J^T J and J^T c in registers - the same sequence of operations as
 *            the sequential oracle, whatever the number of lanes.
 */
template <class M, int G, bool INDIVIDUAL>
UPC_RESOLVER_LINKAGE int resolver_correction(const DevEnv& env, const DevRobot& r, const DevSim& sp, const M& m_in, const Group<G>& g,
                              const double* pts, double* x, LocalCounters& cnt, const unsigned* shared_masks = nullptr, int mask_stride = 0)
{
    int shared_at = 0; /* next word of the masks the preceding collision check recorded (check_collision: share_threshold) */
    const double resolve_distance = sp.resolve_distance;
    const M m = m_in; /* private copy (when the function is compiled out of line the caller's model sits in memory) */
    UPC_T0
    constexpr int N = M::DOF;
    const int n = m.dof(r);
    double H[N][N]; /* accumulators: never passed by address, so they can live in registers */
    double gv[N], xs[N];
#pragma unroll
    for (int a = 0; a < N; a++)
    {
        gv[a] = 0.0; xs[a] = 0.0;
#pragma unroll
        for (int b = 0; b < N; b++) H[a][b] = 0.0;
    }
    int count = 0;
    double cbuf[32][3];
    for (int l = 0; l < r.num_links; l++)
    {
        const int beg = r.link_off[l], end = r.link_off[l + 1];
        if (beg == end) continue;
        double T[12], V[12];
        m.frame(r, l, T);
        voxel_from_link(env, T, V);
        const float clear_above = float_round_up((r.link_radius[l] + resolve_distance) + UPC_CULL_MARGIN);
        if (shared_masks != nullptr)
        {
            /* the check before this pass already culled this configuration at resolve_distance: a link without candidates is done */
            const int link_chunks = (end - beg + 32 * G - 1) / (32 * G);
            unsigned any = 0u;
            for (int c = 0; c < link_chunks; c++) any |= shared_masks[(shared_at + c) * mask_stride];
            if (G > 1) any = g.any(any != 0u) ? 1u : 0u; /* the accumulation phase below is a group operation: all lanes or none */
            if (any == 0u)
            {
                shared_at += link_chunks;
                continue;
            }
        }
        /* link-level cull: a link farther than resolve_distance from every obstacle contributes nothing */
        else if (link_certainly_clear(env, r, l, V, clear_above)) continue;
        /* accumulate the contribution (b0, b1, b2) of body point kk of this link; called in ascending point order */
        auto accumulate = [&](int kk, double b0, double b1, double b2)
        {
            const double local[3] = {pts[4 * kk + 0], pts[4 * kk + 1], pts[4 * kk + 2]};
            double world[3];
            xf_point(T, local[0], local[1], local[2], world);
            double J[3][N];
            m.jacobian(r, l, local, world, J);
            count++;
            if (INDIVIDUAL)
            {
                double A[3][3];
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
#pragma unroll
                    for (int qq = 0; qq < 3; qq++)
                    {
                        double s = 0.0;
#pragma unroll
                        for (int a = 0; a < N; a++)
                            if (a < n) s = upc_fma(J[rr][a], J[qq][a], s);
                        A[rr][qq] = s;
                    }
                const double cv[3] = {b0, b1, b2};
                double yv[3];
                solve_damped<3>(A, cv, sp.damping, yv, 3);
#pragma unroll
                for (int a = 0; a < N; a++)
                    if (a < n) xs[a] = upc_fma(J[2][a], yv[2], upc_fma(J[1][a], yv[1], upc_fma(J[0][a], yv[0], xs[a])));
            }
            else
            {
#pragma unroll
                for (int a = 0; a < N; a++)
                {
                    if (a < n)
                    {
#pragma unroll
                        for (int b = 0; b < N; b++)
                        {
                            if (b >= a && b < n)
                                H[a][b] = upc_fma(J[2][a], J[2][b], upc_fma(J[1][a], J[1][b], upc_fma(J[0][a], J[0][b], H[a][b])));
                        }
                        gv[a] = upc_fma(J[2][a], b2, upc_fma(J[1][a], b1, upc_fma(J[0][a], b0, gv[a])));
                    }
                }
            }
        };
        for (int chunk = beg; chunk < end; chunk += 32 * G)
        {
            const int first = chunk + g.sub;
            const int lim = (end < chunk + 32 * G) ? end : (chunk + 32 * G);
            unsigned contributing = 0u;
            unsigned undecided = 0u;
            unsigned group_mask = 0xffffffffu;
            if constexpr (G == 1) group_mask = group_candidates(env, r, l, (chunk - beg) >> 5, V, clear_above);
            if (env.cells != nullptr)
            {
                /* corner-cell table: phases A and B in one pass over the lane's points, kResolverInFlight lookups in flight per trip */
                bool culled_pass = false;
                if constexpr (UPC_RES_PRECULL && G >= resolver_precull_min_lanes<M>() && G <= 8)
                if ((r.cull_delta > 0.0f) && (end - beg >= 8 * G))
                {
                    culled_pass = true;
                    /* mask-then-evaluate, as in the collision check: the single-precision pre-cull (4-byte cell minimum) drops
                     * the points that are provably farther than resolve_distance, and only the rest - a few per lane - take
                     * the exact gradient evaluation.  A warp then runs as many expensive trips as its busiest lane has
                     * candidates instead of one per point slot (some lane of the warp is near a surface at almost every slot).
                     * Compiled for 2..8 lanes only: measured C2 (4 lanes) 2.11 -> 2.03 ms, but the one-lane big-batch program
                     * (-3 %) and the 32-lane SE(2) program (+5 % on C1, where the pass never runs) lose by carrying it. */
                    unsigned candidates;
                    if (shared_masks != nullptr)
                    {
                        candidates = shared_masks[(shared_at++) * mask_stride];
                    }
                    else
                    {
                        const CullFrame cf = cull_frame<M::PLANAR>(env, r, l, V, r.cull_delta, clear_above);
                        const CullPoint* ptsf = reinterpret_cast<const CullPoint*>(pts + 4 * r.num_points);
                        candidates = cull_undecided<M::PLANAR>(env, cf, ptsf, first, lim, G, group_mask);
                    }
                    while (candidates)
                    {
                        const int j = lowest_bit(candidates);
                        candidates &= candidates - 1u;
                        const int k = first + j * G;
                        double u[3];
                        xf_point(V, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], u);
                        const CellRef c = cell_locate(env, u);
                        const Corner4 ca = ldg4(c.cell);
                        const Corner4 cb = ldg4(c.cell + 4);
                        double grad[3], dist;
                        if (cell_gradient(env, c, ca, cb, pts[4 * k + 3], resolve_distance, clear_above, &dist, grad))
                        {
                            const double de = dist - pts[4 * k + 3];
                            if (de < resolve_distance)
                            {
                                const double gn2 = dot3(grad, grad);
                                if (gn2 > 1.0e-24)
                                {
                                    const double scale = (resolve_distance - de) / sqrt(gn2);
                                    if constexpr (G == 1 && UPC_ONE_LANE_IMMEDIATE)
                                    {
                                        accumulate(k, grad[0] * scale, grad[1] * scale, grad[2] * scale);
                                    }
                                    else
                                    {
                                        cbuf[j][0] = grad[0] * scale;
                                        cbuf[j][1] = grad[1] * scale;
                                        cbuf[j][2] = grad[2] * scale;
                                        contributing |= (1u << j);
                                    }
                                }
                            }
                        }
                    }
                }
                if (!culled_pass)
                for (int j0 = 0; j0 < 32 && (first + j0 * G) < lim; j0 += kResolverInFlight)
                {
                    if (((group_mask >> j0) & ((1u << kResolverInFlight) - 1u)) == 0u) continue; /* cleared by the group-level cull */
                    CellRef c[kResolverInFlight];
                    Corner4 ca[kResolverInFlight], cb[kResolverInFlight];
#pragma unroll
                    for (int jj = 0; jj < kResolverInFlight; jj++)
                    {
                        int k = first + (j0 + jj) * G;
                        if (k >= lim) k = first; /* padding lanes of the last trip repeat a valid point and are ignored below */
                        double u[3];
                        xf_point(V, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], u);
                        c[jj] = cell_locate(env, u);
                    }
#pragma unroll
                    for (int jj = 0; jj < kResolverInFlight; jj++)
                    {
                        ca[jj] = ldg4(c[jj].cell);
                        cb[jj] = ldg4(c[jj].cell + 4);
                    }
#pragma unroll
                    for (int jj = 0; jj < kResolverInFlight; jj++)
                    {
                        const int j = j0 + jj;
                        const int k = first + j * G;
                        if ((k < lim) && ((group_mask >> j) & 1u))
                        {
                            double grad[3], dist;
                            if (cell_gradient(env, c[jj], ca[jj], cb[jj], pts[4 * k + 3], resolve_distance, clear_above, &dist, grad))
                            {
                                const double de = dist - pts[4 * k + 3];
                                if (de < resolve_distance)
                                {
                                    const double gn2 = dot3(grad, grad);
                                    if (gn2 > 1.0e-24)
                                    {
                                        const double scale = (resolve_distance - de) / sqrt(gn2);
                                        if constexpr (G == 1 && UPC_ONE_LANE_IMMEDIATE)
                                        {
                                            accumulate(k, grad[0] * scale, grad[1] * scale, grad[2] * scale);
                                        }
                                        else
                                        {
                                            cbuf[j][0] = grad[0] * scale;
                                            cbuf[j][1] = grad[1] * scale;
                                            cbuf[j][2] = grad[2] * scale;
                                            contributing |= (1u << j);
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            else
            {
                /* phase A */
UPC_PLAIN_UNROLL_PRAGMA
                for (int j = 0; j < 32; j++)
                {
                    const int k = first + j * G;
                    if ((k < lim) && ((group_mask >> j) & 1u))
                    {
                        double u[3];
                        xf_point(V, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], u);
                        if (!sdf_certainly_clear(env, u, pts[4 * k + 3], resolve_distance)) undecided |= (1u << j);
                    }
                }
            }
            /* phase B */
            unsigned todo_b = undecided;
            while (todo_b)
            {
#if defined(__CUDA_ARCH__)
                const int j = __ffs((int)todo_b) - 1;
#else
                const int j = __builtin_ctz(todo_b);
#endif
                todo_b &= todo_b - 1u;
                const int k = first + j * G;
                double u[3], grad[3];
                xf_point(V, pts[4 * k + 0], pts[4 * k + 1], pts[4 * k + 2], u);
                const double de = sdf_distance_gradient(env, u, grad) - pts[4 * k + 3];
                if (de < resolve_distance)
                {
                    const double gn2 = dot3(grad, grad);
                    if (gn2 > 1.0e-24)
                    {
                        const double scale = (resolve_distance - de) / sqrt(gn2);
                        if constexpr (G == 1 && UPC_ONE_LANE_IMMEDIATE)
                        {
                            accumulate(k, grad[0] * scale, grad[1] * scale, grad[2] * scale);
                        }
                        else
                        {
                            cbuf[j][0] = grad[0] * scale;
                            cbuf[j][1] = grad[1] * scale;
                            cbuf[j][2] = grad[2] * scale;
                            contributing |= (1u << j);
                        }
                    }
                }
            }
            UPC_T1(8)
            /* phase C: rounds in ascending j, lanes in ascending rank = ascending point index */
            unsigned rounds = contributing;
#if defined(__CUDA_ARCH__)
            if constexpr (G > 1 && (G & (G - 1)) == 0)
            {
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) rounds |= __shfl_xor_sync(g.mask, rounds, off);
            }
            else if constexpr (G > 1)
            {
                unsigned all = 0u;
#pragma unroll
                for (int t = 0; t < G; t++) all |= __shfl_sync(g.mask, contributing, g.base + t);
                rounds = all;
            }
#endif
            while (rounds)
            {
#if defined(__CUDA_ARCH__)
                const int j = __ffs((int)rounds) - 1;
#else
                const int j = __builtin_ctz(rounds);
#endif
                rounds &= rounds - 1u;
                const bool mine = ((contributing >> j) & 1u) != 0u;
                double c0 = 0.0, c1 = 0.0, c2 = 0.0;
                if (mine) { c0 = cbuf[j][0]; c1 = cbuf[j][1]; c2 = cbuf[j][2]; }
                unsigned todo = g.ballot(mine);
                while (todo)
                {
#if defined(__CUDA_ARCH__)
                    const int src = __ffs((int)todo) - 1;
#else
                    const int src = 0;
#endif
                    todo &= todo - 1u;
                    const double b0 = g.bcast(c0, src), b1 = g.bcast(c1, src), b2 = g.bcast(c2, src);
                    const int kk = chunk + src + j * G;
                    accumulate(kk, b0, b1, b2);
                }
            }
            UPC_T1(9)
        }
    }
#if defined(UPC_PHASE_TIMING) && defined(__CUDA_ARCH__)
    cnt.cyc[11] += count; /* contributing points per resolver pass (a count, not cycles) */
#endif
    if (count == 0) return 0;
    if (INDIVIDUAL)
    {
        const double inv = 1.0 / (double)count;
#pragma unroll
        for (int a = 0; a < N; a++)
            if (a < n) x[a] = xs[a] * inv;
    }
    else
    {
        double Hs[N][N]; /* the out-of-line solver works on its own copy */
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) Hs[a][b] = H[a][b];
        solve_damped<N>(Hs, gv, sp.damping, x, n);
    }
    UPC_T1(10)
    return count;
}

/*
 * Counter-based draws of one (particle, controller interval, slot) for all axes: slot 0 = sensor draws
 * (standard draw * max_sensor_noise / 2), slots 1.. = actuator draws (standard draw / 2); a noise-free axis gets an exact
 * 0.0 and, when both axes of a pair are noise-free, no Philox block is computed.
 */
template <int N>
UPC_D void seeded_draws(const DevRobot& r, uint64_t seed, uint64_t particle, int step, int slot, double* out)
{
    constexpr int kPairs = (N + 1) / 2;
    constexpr int kStride = UPC_NOISE_QUAD ? 2 : 1; /* pairs per generator call */
#pragma unroll
    for (int p = 0; p < kPairs; p += kStride)
    {
        if (2 * p < r.dof)
        {
            double s[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int d = 2 * p + k;
                const bool used = (k < 2 * kStride) && (d < N) && (d < r.dof);
                s[k] = used ? ((slot == 0) ? r.axis[(d < N) ? d : 0].sensor_scale : r.axis[(d < N) ? d : 0].actuator_scale) : 0.0;
            }
            const bool first = (s[0] != 0.0) || (s[1] != 0.0), second = (s[2] != 0.0) || (s[3] != 0.0);
            NoiseQuad v;
            v.z[0] = 0.0; v.z[1] = 0.0; v.z[2] = 0.0; v.z[3] = 0.0;
            if (first || second) v = noise_quad_values(seed, particle, (uint32_t)step, (uint32_t)slot, (uint32_t)p, second);
#pragma unroll
            for (int k = 0; k < 2 * kStride; k++)
            {
                const int d = 2 * p + k;
                if ((d < N) && (d < r.dof)) out[(d < N) ? d : 0] = (s[k] != 0.0) ? (v.z[k] * s[k]) : 0.0;
            }
        }
    }
}

/*
 * Simulate particle i for one planner edge. `pts` are the robot's body points (shared memory on the GPU),
 * frames0/frames1 are per-particle scratch for link transforms (linked robots only), `park` is per-particle
 * scratch of 2*WIDTH + 2*DOF doubles (shared memory on the GPU) holding the target, the pre-microstep
 * configuration and the PID state - values touched once per controller interval (or only when a microstep is
 * reverted) that would otherwise occupy registers for the whole edge.  Every lane of a group writes identical
 * values there.
 * Writes the final configuration and the collided flag from lane 0 of the group.
 *
 * Control flow: one loop body = "apply an input, then collision-check", used both for noisy microsteps and for
 * noise-free resolver corrections, so that the model's apply / the motion estimate / the collision check are
 * each instantiated once (code size is instruction-cache pressure here).  The sequence of operations per
 * particle is exactly the sequential program of DESIGN.md section 3.
 */
template <class M, int G, bool SEEDED>
UPC_D void simulate_particle(const DevEnv& env, const DevRobot& r, const DevSim& sp, const Group<G>& g, int64_t i, int64_t n,
                             const double* starts, int64_t n_targets, const double* targets, const double* tape,
                             double* out_configs, uint8_t* out_collided, const double* pts, double* frames0,
                             double* frames1, double* park, LocalCounters& cnt, bool& collided_out, const bool exists = true,
                             unsigned* masks = nullptr, const int mask_stride = 0)
{
    /* masks: kMaxMaskChunks words per thread (stride mask_stride) for the pre-cull shared by a collision check and the resolver
     * pass behind it; NULL = every pass culls for itself */
    constexpr bool kShare = shares_cull_masks<M, G>();
    int masks_state = -1;
    /* exists == false: a lane without a particle (tail of the batch, lanes a group size leaves over).  It runs the same control flow
     * on particle i (a valid index) with `finished` set from the start and writes nothing to global memory, so that every thread of
     * a CTA executes the SAME per-interval barrier instruction - an aligned barrier must not be reached divergently inside a warp. */
    constexpr int N = M::DOF;
    const int dof = r.dof;
    const int width = (M::HAS_FRAMES) ? dof : M::WIDTH;
    M m;
    if constexpr (M::HAS_FRAMES) m.frames = frames0;
    double* target = park;
    double* previous = park + M::WIDTH;
    m.pid_i = park + 2 * M::WIDTH;
    m.pid_e = park + 2 * M::WIDTH + M::DOF;
    const bool resumed = sp.step_begin > 0;
    if (resumed) m.load(r, out_configs, n, i);
    else m.load(r, starts, sp.n_starts, (sp.per_start == 1) ? i : ((sp.per_start == n) ? 0 : (i / sp.per_start)));
    m.reset_controllers();
    bool collided = false;
    bool finished = !exists;
    if (resumed)
    {
        for (int d = 0; d < dof; d++)
        {
            m.pid_i[d] = sp.carry_pid[(int64_t)d * n + i];
            m.pid_e[d] = sp.carry_pid[(int64_t)(dof + d) * n + i];
        }
        collided = sp.carry_flags[i] != 0;
        finished = (sp.carry_flags[n + i] != 0) || !exists;
    }
    refresh_frames(r, m, g);
    {
        const int64_t ti = (sp.per_target == 1) ? i : ((sp.per_target == n) ? 0 : (i / sp.per_target));
        if (g.sub == 0)
        {
            for (int c = 0; c < width; c++) target[c] = targets[(int64_t)c * n_targets + ti];
        }
        g.sync();
    }
    const int slots = 1 + sp.max_microsteps;
#if UPC_STEP_BARRIER && defined(__CUDA_ARCH__)
    /* the warps of a CTA enter every controller interval together, so that they run through the same stretch of the program at
     * about the same time and share its instruction-cache lines: without the barrier the SM instruction cache hits 78 % and its
     * refills keep the GPC instruction cache at 95 % of its request rate (profiles/r02g_sim_c5_metrics.txt) - the per-particle
     * program is 140 KB of SASS against 32 KB of SM instruction cache.  Every thread of the CTA passes the barrier once per
     * interval, finished particles included. */
    for (int step = sp.step_begin; step < sp.step_end; step++)
    {
        if (((step - sp.step_begin) % UPC_STEP_BARRIER_EVERY) == 0) __syncthreads();
        if (finished) continue;
        cnt.particle_steps++;
#else
    for (int step = sp.step_begin; step < sp.step_end && !finished; step++)
    {
        cnt.particle_steps++;
#endif
        UPC_T0
        double vec[N];   /* the input about to be estimated / applied: control step, then resolver corrections */
        double ustep[N]; /* control input per microstep */
        {
            double draws[N];
            if (SEEDED && !r.sensor_noise_free)
            {
                seeded_draws<N>(r, sp.seed, (uint64_t)(sp.first_particle + i), step, 0, draws);
            }
            else
            {
#pragma unroll
                for (int d = 0; d < N; d++)
                    if (d < dof) draws[d] = (SEEDED || r.sensor_noise_free) ? 0.0 : tape[(((int64_t)step * slots) * dof + d) * n + i]; /* zeros by construction: not read */
            }
#ifdef UPC_PHASE_TIMING
            if (draws[0] == 123.456) cnt.particle_steps++; /* force the loads to complete inside this phase */
#endif
            UPC_T1(6)
            double u[N];
            m.control(r, target, sp.dt, draws, u);
#pragma unroll
            for (int d = 0; d < N; d++)
                if (d < dof) vec[d] = u[d] * sp.dt;
            if (sp.trace_configs != nullptr && g.sub == 0)
            {
#pragma unroll
                for (int d = 0; d < N; d++)
                    if (d < dof) sp.trace_controls[((int64_t)step * 2 * dof + d) * n + i] = u[d];
            }
            UPC_T1(0)
        }
        unsigned microsteps = 1u, ms = 0u;
        bool resolving = false;    /* false: next action is a (noisy) microstep; true: a resolver correction */
        bool need_estimate = true; /* the control step and every resolver correction are measured first */
        int iterations = 0;
        double scaling = sp.initial_scale;
        bool step_done = false;
        while (!step_done)
        {
            double apply_in[N];
            const double* noise = nullptr;
            double draws[N];
            if (need_estimate)
            {
                /* a static upper bound decides the common case (motion <= microstep_distance => one microstep, no
                 * down-scaling) without touching the body points; otherwise the exact estimate runs */
                /* with a single microstep slot the clamp below yields one microstep whatever the estimate says (BASELINE configs: M = 1) */
                const bool small = (!resolving && sp.max_microsteps == 1) || ((m.motion_bound(r, vec) * 1.000001) <= sp.microstep_distance);
                double est = 0.0;
                if (!small) est = estimate_motion(r, m, g, pts, vec, frames1);
                if (!resolving)
                {
                    microsteps = small ? 1u : (unsigned)ceil(est / sp.microstep_distance);
                    if (microsteps < 1u) microsteps = 1u;
                    if (microsteps > (unsigned)sp.max_microsteps) microsteps = (unsigned)sp.max_microsteps;
                    const double inv_micro = 1.0 / (double)microsteps;
#pragma unroll
                    for (int d = 0; d < N; d++)
                        if (d < dof) ustep[d] = vec[d] * inv_micro;
                    if (sp.trace_configs != nullptr && g.sub == 0)
                    {
#pragma unroll
                        for (int d = 0; d < N; d++)
                            if (d < dof) sp.trace_controls[((int64_t)step * 2 * dof + dof + d) * n + i] = ustep[d];
                    }
                }
                else
                {
                    const double limit_scale = (!small && est > sp.microstep_distance) ? (sp.microstep_distance / est) : 1.0;
                    const double total_scale = limit_scale * fabs(scaling);
#pragma unroll
                    for (int d = 0; d < N; d++)
                        if (d < dof) apply_in[d] = vec[d] * total_scale;
                }
                need_estimate = false;
                UPC_T1(1)
            }
            if (!resolving)
            {
                /* noisy microstep `ms` */
                g.sync();
                if (g.sub == 0) m.get(r, previous);
                g.sync();
                if (SEEDED) seeded_draws<N>(r, sp.seed, (uint64_t)(sp.first_particle + i), step, 1 + (int)ms, draws);
#pragma unroll
                for (int d = 0; d < N; d++)
                    if (d < dof)
                    {
                        if (!SEEDED) draws[d] = tape[(((int64_t)step * slots + 1 + ms) * dof + d) * n + i];
                        apply_in[d] = ustep[d];
                    }
                noise = draws;
            }
            UPC_T1(6)
            m.apply(r, apply_in, noise);
            refresh_frames(r, m, g);
            UPC_T1(2)
            /* one call site (the check is inlined: a second one would be a second copy of it in a program bound by instruction fetch) */
            const bool share = kShare && (masks != nullptr) && (sp.allow_contacts != 0);
            bool in_collision = check_collision(env, r, m, g, pts, sp.contact_distance, share ? sp.resolve_distance : -1.0, share ? masks : nullptr,
                                                mask_stride, &masks_state);
            UPC_T1(3)
            bool revert = false;
            if (!resolving)
            {
                if (in_collision)
                {
                    collided = true;
                    if (sp.allow_contacts)
                    {
                        resolving = true;
                        iterations = 0;
                        scaling = sp.initial_scale;
                    }
                    else
                    {
                        revert = true;
                    }
                }
            }
            else
            {
                iterations++;
                cnt.resolver_iterations++;
                if (in_collision)
                {
                    if (iterations >= sp.max_resolver_iterations)
                    {
                        revert = true;
                        cnt.failed_resolves++;
                    }
                    else if ((iterations % sp.decay_iterations) == 0)
                    {
                        scaling = scaling * sp.decay_rate;
                    }
                }
            }
            if (resolving && in_collision && !revert)
            {
                /* next action: a resolver correction from the current configuration */
                const unsigned* shared = (kShare && masks_state >= 0) ? masks : nullptr;
                const int contributing = sp.individual_jacobians ? resolver_correction<M, G, true>(env, r, sp, m, g, pts, vec, cnt, shared, mask_stride)
                                                                 : resolver_correction<M, G, false>(env, r, sp, m, g, pts, vec, cnt, shared, mask_stride);
                UPC_T1(4)
                if (contributing == 0)
                {
                    revert = true;
                    cnt.failed_resolves++;
                }
                else
                {
                    need_estimate = true;
                }
            }
            if (revert)
            {
                double prev[M::WIDTH];
#pragma unroll
                for (int c = 0; c < M::WIDTH; c++)
                    if (c < width) prev[c] = previous[c];
                m.set(r, prev);
                refresh_frames(r, m, g);
                step_done = true; /* a reverted microstep ends the controller interval */
            }
            else if (!in_collision)
            {
                /* microstep finished (directly, or after a successful resolve) */
                resolving = false;
                ms++;
                if (ms >= microsteps) step_done = true;
            }
        }
        const bool arrived = m.reached(r, target, sp.shortcut);
        UPC_T1(5)
        if (arrived) finished = true;
        if (sp.trace_configs != nullptr && g.sub == 0)
        {
            /* tracing (single-particle path): the resolved configuration of this interval and the inputs behind it */
            m.store(r, sp.trace_configs + (int64_t)step * width * n, n, i);
            sp.trace_steps[i] = step + 1;
            sp.trace_steps[n + i] = (int32_t)microsteps;
        }
    }
    const bool last_launch = sp.step_end >= sp.num_steps;
    if (g.sub == 0 && exists)
    {
        m.store(r, out_configs, n, i);
        out_collided[i] = collided ? 1 : 0;
        if (!last_launch)
        {
            for (int d = 0; d < dof; d++)
            {
                sp.carry_pid[(int64_t)d * n + i] = m.pid_i[d];
                sp.carry_pid[(int64_t)(dof + d) * n + i] = m.pid_e[d];
            }
            sp.carry_flags[i] = collided ? 1 : 0;
            sp.carry_flags[n + i] = finished ? 1 : 0;
        }
    }
    collided_out = collided && last_launch; /* the collided-particle counter is bumped once per edge */
}

} /* namespace upc */

#endif /* UPC_SIM_CUH */
