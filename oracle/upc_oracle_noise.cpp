/*
 * upc_oracle_noise.cpp - CPU restatement of the counter-based noise of DESIGN.md section 4.2.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows the published algorithm and the specification, not the product's source: Philox-4x32-10 (Salmon, Moraes, Dror,
 * Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; pinned by the Random123 known-answer vectors in
 * tests/test_noise.py) and the rational approximation of the truncated normal quantile specified in DESIGN.md 4.2 (pinned
 * against scipy.special.ndtri).  The role of the
 * draws is that of simple_uncertainty_models.hpp:38-45 (sensor: value + draw) and :77-88 (actuator: draw in [-1, 1] scaled
 * by the commanded value), with the truncation bounds of :29 and :59 (mean 0, +-2 sigma).
 *
 * oracle_fill_noise_tape_counter writes the draws of particles first..first+n-1 as a noise tape in the layout of
 * upc_forward_simulate ([step][slot][axis][particle]); oracle_forward_simulate then replays that tape.  The seeded GPU
 * entry points, which evaluate the same function inside the kernel, must reproduce those results bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "upc_b200.h"
#include "upc_math.h"

namespace
{
struct Block128
{
    uint32_t v[4];
};

/* one Philox-4x32 round: two 32x32 -> 64 multiplies, the high halves are mixed with the other counter words and the key */
Block128 PhiloxRound(const Block128& ctr, uint32_t key0, uint32_t key1)
{
    const uint64_t product0 = (uint64_t)0xD2511F53u * (uint64_t)ctr.v[0];
    const uint64_t product1 = (uint64_t)0xCD9E8D57u * (uint64_t)ctr.v[2];
    Block128 next;
    next.v[0] = (uint32_t)(product1 >> 32) ^ ctr.v[1] ^ key0;
    next.v[1] = (uint32_t)product1;
    next.v[2] = (uint32_t)(product0 >> 32) ^ ctr.v[3] ^ key1;
    next.v[3] = (uint32_t)product0;
    return next;
}

Block128 Philox4x32_10(Block128 ctr, uint32_t key0, uint32_t key1)
{
    for (int round = 0; round < 10; round++)
    {
        if (round > 0)
        {
            key0 += 0x9E3779B9u; /* Weyl sequence: golden ratio */
            key1 += 0xBB67AE85u; /* sqrt(3) - 1 */
        }
        ctr = PhiloxRound(ctr, key0, key1);
    }
    return ctr;
}

/*
 * Quantile of the standard normal truncated to +-2 sigma at the centred uniform t in (-1/2, 1/2): Phi^-1(1/2 + q) with
 * q = t (Phi(2) - Phi(-2)), by the degree-8 / degree-8 rational approximation q P(r) / Q(r), r = 0.228 - q^2, that DESIGN.md 4.2
 * specifies (coefficient table produced by scripts/fit_truncated_quantile.py; Horner steps fused, as the specification says).
 * tests/test_noise.py pins it against scipy.special.ndtri to 1e-13.
 */
const double kQuantileP[9] = {4.19803057674432534e+00, 4.98914924494953766e+02, 2.23256080577635075e+04, 4.73334579451213649e+05, 4.94985450103831850e+06,
                              2.44404971329663284e+07, 5.03468522432988584e+07, 3.29915794812198505e+07, 3.04619464759342372e+06};
const double kQuantileQ[9] = {1.0, 1.26415601550449168e+02, 6.10854034951519316e+03, 1.42817358525897551e+05, 1.69745287590994616e+06,
                              9.97100045574528910e+06, 2.63619050062270910e+07, 2.58257609111968540e+07, 5.87769839876705315e+06};

double TruncatedNormalQuantile(double t)
{
    const double q = t * 9.544997361036415856e-01;
    const double r = upc_fma(-q, q, 0.228);
    double numerator = kQuantileP[8], denominator = kQuantileQ[8];
    for (int i = 7; i >= 0; i--)
    {
        numerator = upc_fma(numerator, r, kQuantileP[i]);
        denominator = upc_fma(denominator, r, kQuantileQ[i]);
    }
    const double z = (q * numerator) / denominator;
    if (z < -2.0) return -2.0;
    if (z > 2.0) return 2.0;
    return z;
}

/* 52 random bits -> (k + 1/2) 2^-52 - 1/2, exactly (k + 1/2 - 2^51 is a half-integer below 2^51: 52 significant bits) */
double CenteredUniform52(uint32_t hi, uint32_t lo)
{
    const uint64_t k = ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
    const int64_t twice = 2 * (int64_t)k + 1 - ((int64_t)1 << 52); /* 2 (k + 1/2 - 2^51) */
    return ldexp((double)twice, -53);
}

/* standard draw of (seed, particle, step, slot, axis) */
double StandardDraw(uint64_t seed, uint64_t particle, uint32_t step, uint32_t slot, uint32_t axis)
{
    Block128 ctr;
    ctr.v[0] = (uint32_t)particle;
    ctr.v[1] = (uint32_t)(particle >> 32);
    ctr.v[2] = step;
    ctr.v[3] = (slot << 16) | (axis >> 1);
    const Block128 bits = Philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double t = (axis & 1u) ? CenteredUniform52(bits.v[2], bits.v[3]) : CenteredUniform52(bits.v[0], bits.v[1]);
    return TruncatedNormalQuantile(t);
}
} /* namespace */

extern "C" {

void oracle_philox4x32_10(const uint32_t* key2, const uint32_t* counter4, uint32_t* out4)
{
    Block128 c;
    memcpy(c.v, counter4, sizeof(c.v));
    const Block128 r = Philox4x32_10(c, key2[0], key2[1]);
    memcpy(out4, r.v, sizeof(r.v));
}

/* z(t) for centred uniforms t in (-1/2, 1/2) */
void oracle_truncated_normal_quantile(int64_t n, const double* t, double* out)
{
    for (int64_t i = 0; i < n; i++) out[i] = TruncatedNormalQuantile(t[i]);
}

void oracle_centered_uniform52(int64_t n, const uint32_t* hi, const uint32_t* lo, double* out)
{
    for (int64_t i = 0; i < n; i++) out[i] = CenteredUniform52(hi[i], lo[i]);
}

void oracle_log(int64_t n, const double* x, double* out)
{
    for (int64_t i = 0; i < n; i++) out[i] = upc_log(x[i]);
}

/* truncated standard normals of axis `axis` for particles first..first+n-1 at (step, slot) */
void oracle_standard_draws(uint64_t seed, int64_t first, int64_t n, uint32_t step, uint32_t slot, uint32_t axis, double* out)
{
    for (int64_t i = 0; i < n; i++) out[i] = StandardDraw(seed, (uint64_t)(first + i), step, slot, axis);
}

int32_t oracle_fill_noise_tape_counter(uint64_t seed, const upc_robot_desc* robot, const upc_sim_params* params, int64_t first_particle,
                                       int64_t n, int64_t tape_stride, double* tape)
{
    if (!robot || !params || !tape || tape_stride < n) return 1;
    const int dof = robot->dof;
    const int slots = 1 + params->max_microsteps;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
    {
        const uint64_t particle = (uint64_t)(first_particle + i);
        for (int step = 0; step < params->num_steps; step++)
            for (int slot = 0; slot < slots; slot++)
                for (int d = 0; d < dof; d++)
                {
                    double value = 0.0;
                    if (slot == 0)
                    {
                        const double bound = fabs(robot->axis[d].max_sensor_noise);
                        if (bound > 0.0) value = StandardDraw(seed, particle, (uint32_t)step, 0u, (uint32_t)d) * (bound * 0.5);
                    }
                    else
                    {
                        const double bound = fabs(robot->axis[d].max_actuator_noise);
                        if (bound > 0.0) value = StandardDraw(seed, particle, (uint32_t)step, (uint32_t)slot, (uint32_t)d) * 0.5;
                    }
                    tape[(((int64_t)step * slots + slot) * dof + d) * tape_stride + i] = value;
                }
    }
    return 0;
}

} /* extern "C" */
