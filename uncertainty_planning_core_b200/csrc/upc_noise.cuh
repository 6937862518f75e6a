/*
 * upc_noise.cuh - counter-based truncated-normal noise.
 *
 * Replaces, on the seeded entry points, the generator fan-out of uncertainty_contact_planning.hpp:343-356 and the
 * truncated-normal draws of simple_uncertainty_models.hpp:38-45 (sensor) and :77-88 (actuator): the reference threads
 * one std::mt19937_64 per OpenMP thread through every draw, so its stream depends on the thread count; here a draw is a
 * pure function of (seed, particle, controller interval, slot, axis) and can be evaluated anywhere - inside the
 * simulation kernel, on any GPU of a sharded batch, or on the host when a tape of the same draws is wanted.
 *
 *   bits    Philox-4x32-10 (Salmon et al., SC'11): key = the 64-bit seed, counter = (particle lo, particle hi,
 *           controller interval, slot << 16 | axis pair); one block of 128 bits serves axes 2j and 2j + 1
 *   uniform u = (k + 1/2) 2^-52 from 52 of the 64 bits of an axis: strictly inside (0, 1), exactly representable
 *   normal  z = Phi^-1(Phi(-2) + u (Phi(2) - Phi(-2))): the inverse-CDF transform of the normal distribution truncated
 *           to +-2 sigma (the reference's distributions are mean 0, bounds +-2 sigma: simple_uncertainty_models.hpp:29, 59),
 *           Phi^-1 by Wichura's AS 241 (PPND16) rational approximations - central branch for |p - 1/2| <= 0.425, the
 *           first tail branch otherwise (its r = sqrt(-log p) stays below 1.95 here; the far-tail branch cannot occur);
 *           no rejection loop: one Philox block per axis pair, always
 *   draw    sensor slot: z * (max_sensor_noise / 2); actuator slots: z / 2  (the scaling of the tape generator, section 4.1);
 *           a zero noise bound yields an exact 0.0
 *
 * Written once as __host__ __device__; only exact IEEE operations and upc_log (include/upc_math.h), so host and device
 * produce identical bits.
 */
#ifndef UPC_NOISE_CUH
#define UPC_NOISE_CUH

#include <stdint.h>
#include "upc_math.h"

#if defined(__CUDACC__)
#define UPC_NOISE_HD __host__ __device__ __forceinline__
#else
#define UPC_NOISE_HD inline
#endif

namespace upc
{
UPC_NOISE_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* Philox-4x32 with 10 rounds */
UPC_NOISE_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; round++)
    {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* (k + 1/2) 2^-52 from the top 52 bits of (hi, lo) */
UPC_NOISE_HD double uniform_open52(uint32_t hi, uint32_t lo)
{
    const uint64_t k = ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
    return ((double)k + 0.5) * 2.220446049250313080847e-16; /* 2^-52 */
}

/* inverse of the standard normal CDF for 0.02 < p < 0.98 (AS 241, PPND16, branches 1 and 2) */
UPC_NOISE_HD double normal_quantile_central(double p)
{
    const double q = p - 0.5;
    if (fabs(q) <= 0.425)
    {
        const double r = 0.180625 - q * q;
        double num = 2.5090809287301226727e+3;
        num = num * r + 3.3430575583588128105e+4;
        num = num * r + 6.7265770927008700853e+4;
        num = num * r + 4.5921953931549871457e+4;
        num = num * r + 1.3731693765509461125e+4;
        num = num * r + 1.9715909503065514427e+3;
        num = num * r + 1.3314166789178437745e+2;
        num = num * r + 3.3871328727963666080e+0;
        double den = 5.2264952788528545610e+3;
        den = den * r + 2.8729085735721942674e+4;
        den = den * r + 3.9307895800092710610e+4;
        den = den * r + 2.1213794301586595867e+4;
        den = den * r + 5.3941960214247511077e+3;
        den = den * r + 6.8718700749205790830e+2;
        den = den * r + 4.2313330701600911252e+1;
        den = den * r + 1.0;
        return (q * num) / den;
    }
    const double tail = (q < 0.0) ? p : (1.0 - p);
    const double r = sqrt(-upc_log(tail)) - 1.6;
    double num = 7.74545014278341407640e-4;
    num = num * r + 2.27238449892691845833e-2;
    num = num * r + 2.41780725177450611770e-1;
    num = num * r + 1.27045825245236838258e+0;
    num = num * r + 3.64784832476320460504e+0;
    num = num * r + 5.76949722146069140550e+0;
    num = num * r + 4.63033784615654529590e+0;
    num = num * r + 1.42343711074968357734e+0;
    double den = 1.05075007164441684324e-9;
    den = den * r + 5.47593808499534494600e-4;
    den = den * r + 1.51986665636164571966e-2;
    den = den * r + 1.48103976427480074590e-1;
    den = den * r + 6.89767334985100004550e-1;
    den = den * r + 1.67638483018380384940e+0;
    den = den * r + 2.05319162663775882187e+0;
    den = den * r + 1.0;
    const double z = num / den;
    return (q < 0.0) ? -z : z;
}

/* standard normal truncated to [-2, 2] from one uniform in (0, 1) */
UPC_NOISE_HD double truncated_normal_from_uniform(double u)
{
    const double P_LO = 2.275013194817920720e-02;  /* Phi(-2) */
    const double P_W = 9.544997361036415856e-01;   /* Phi(2) - Phi(-2) */
    const double z = normal_quantile_central(P_LO + u * P_W);
    return upc_max(-2.0, upc_min(2.0, z));
}

/* the two standard truncated normals of axis pair `pair` (axes 2 pair, 2 pair + 1) */
UPC_NOISE_HD void noise_pair(uint64_t seed, uint64_t particle, uint32_t step, uint32_t slot, uint32_t pair, double* z0, double* z1)
{
    uint32_t r[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)particle, (uint32_t)(particle >> 32), step, (slot << 16) | pair, r);
    *z0 = truncated_normal_from_uniform(uniform_open52(r[0], r[1]));
    *z1 = truncated_normal_from_uniform(uniform_open52(r[2], r[3]));
}

/*
 * All draws of one (particle, controller interval, slot): out[d] for d < dof.  scale[d] is the factor applied to the
 * standard draw of axis d (sensor: max_sensor_noise / 2, actuator: 1/2 when the axis has actuator noise) and 0 for a
 * noise-free axis, which yields an exact 0.0.  N is the compile-time capacity of `out`.
 */
template <int N>
UPC_NOISE_HD void noise_draws(uint64_t seed, uint64_t particle, uint32_t step, uint32_t slot, int dof, const double* scale, double* out)
{
#pragma unroll
    for (int p = 0; p < (N + 1) / 2; p++)
    {
        const int d0 = 2 * p, d1 = 2 * p + 1;
        if (d0 < dof)
        {
            const double s0 = scale[d0];
            const double s1 = (d1 < N && d1 < dof) ? scale[d1] : 0.0;
            double z0 = 0.0, z1 = 0.0;
            if ((s0 != 0.0) || (s1 != 0.0)) noise_pair(seed, particle, step, slot, (uint32_t)p, &z0, &z1);
            out[d0] = (s0 != 0.0) ? z0 * s0 : 0.0;
            if (d1 < N && d1 < dof) out[d1] = (s1 != 0.0) ? z1 * s1 : 0.0;
        }
    }
}

} /* namespace upc */

#endif /* UPC_NOISE_CUH */
