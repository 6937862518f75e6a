/*
 * upc_oracle_noise.cpp - CPU restatement of the counter-based noise of DESIGN.md section 4.2.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows the published algorithms, not the product's source: Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random
 * numbers: as easy as 1, 2, 3", SC'11; pinned by the Random123 known-answer vectors in tests/test_noise.py) and the inverse
 * normal CDF AS 241 / PPND16 (Wichura, Applied Statistics 37, 1988; pinned against scipy.special.ndtri).  The role of the
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

double Horner8(const double* c, double r)
{
    double acc = c[7];
    for (int i = 6; i >= 0; i--) acc = acc * r + c[i];
    return acc;
}

/* AS 241 PPND16 restricted to 0.02 < p < 0.98 (its third branch, r > 5, needs p < 1.4e-11) */
double NormalQuantile(double p)
{
    static const double A[8] = {3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3, 1.3731693765509461125e+4,
                                4.5921953931549871457e+4, 6.7265770927008700853e+4, 3.3430575583588128105e+4, 2.5090809287301226727e+3};
    static const double B[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
                                2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4, 5.2264952788528545610e+3};
    static const double C[8] = {1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0, 3.64784832476320460504e0,
                                1.27045825245236838258e0, 2.41780725177450611770e-1, 2.27238449892691845833e-2, 7.74545014278341407640e-4};
    static const double D[8] = {1.0, 2.05319162663775882187e0, 1.67638483018380384940e0, 6.89767334985100004550e-1,
                                1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4, 1.05075007164441684324e-9};
    const double q = p - 0.5;
    if (fabs(q) <= 0.425)
    {
        const double r = 0.180625 - q * q;
        return (q * Horner8(A, r)) / Horner8(B, r);
    }
    double r = (q < 0.0) ? p : (1.0 - p);
    r = sqrt(-upc_log(r)) - 1.6;
    const double val = Horner8(C, r) / Horner8(D, r);
    return (q < 0.0) ? -val : val;
}

double UniformOpen52(uint32_t hi, uint32_t lo)
{
    const uint64_t k = ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
    return ((double)k + 0.5) * ldexp(1.0, -52);
}

double TruncatedStandardNormal(double u)
{
    const double p = 2.275013194817920720e-02 + u * 9.544997361036415856e-01;
    const double z = NormalQuantile(p);
    if (z < -2.0) return -2.0;
    if (z > 2.0) return 2.0;
    return z;
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
    const double u = (axis & 1u) ? UniformOpen52(bits.v[2], bits.v[3]) : UniformOpen52(bits.v[0], bits.v[1]);
    return TruncatedStandardNormal(u);
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

void oracle_normal_quantile(int64_t n, const double* p, double* out)
{
    for (int64_t i = 0; i < n; i++) out[i] = NormalQuantile(p[i]);
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
