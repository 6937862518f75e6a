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
 *   uniform t = (k + 1/2) 2^-52 - 1/2 from 52 of the 64 bits of an axis: strictly inside (-1/2, 1/2), never 0, exact
 *   normal  z = Phi^-1(1/2 + t (Phi(2) - Phi(-2))): the inverse-CDF transform of the normal distribution truncated to +-2 sigma
 *           (the reference's distributions are mean 0, bounds +-2 sigma: simple_uncertainty_models.hpp:29, 59), evaluated by one
 *           rational approximation of degree 8 / 8 valid on the whole truncated range (|error| <= 4e-14; Wichura's AS 241 needs
 *           a logarithm and a square root beyond |p - 1/2| = 0.425, a second branch that nearly every warp would execute);
 *           no rejection loop: one Philox block per axis pair, always
 *   draw    sensor slot: z * (max_sensor_noise / 2); actuator slots: z / 2  (the scaling of the tape generator, section 4.1);
 *           a zero noise bound yields an exact 0.0
 *
 * Written once as __host__ __device__; only exact IEEE operations (explicit fused multiply-adds, one division), so host and
 * device produce identical bits.
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

/* (k + 1/2) 2^-52 - 1/2 for the 52-bit integer k = top 52 bits of (hi, lo): a uniform variate in (-1/2, 1/2), never 0, exact.
 * k is placed in the mantissa of a double in [1, 2) - no integer-to-double conversion; both subtractions are exact. */
UPC_NOISE_HD double uniform_centered52(uint32_t hi, uint32_t lo)
{
    const uint64_t k = ((uint64_t)hi << 20) | (uint64_t)(lo >> 12);
    const double one_to_two = upc_bits_double(0x3ff0000000000000ull | k); /* 1 + k 2^-52 */
    return (one_to_two - 1.5) + 1.1102230246251565404e-16;                  /* + 2^-53 */
}

/*
 * Standard normal truncated to [-2, 2] from a centred uniform t in (-1/2, 1/2): z = Phi^-1(1/2 + q), q = t (Phi(2) - Phi(-2)).
 * On |q| <= 0.47725 the quantile is evaluated by ONE rational approximation z = q P(r) / Q(r), r = 0.228 - q^2, of degree 8 / 8
 * (coefficients fitted in 60-digit arithmetic by scripts/fit_truncated_quantile.py; all positive, so the fused Horner steps do
 * not cancel): maximum absolute error 4e-14 against the exact quantile.  No branch, no logarithm, one division.
 */
UPC_NOISE_HD double truncated_normal_from_centered(double t)
{
    const double q = t * 9.544997361036415856e-01;
    const double r = upc_fma(-q, q, 0.228);
    double num = 3.04619464759342372e+06;
    num = upc_fma(num, r, 3.29915794812198505e+07);
    num = upc_fma(num, r, 5.03468522432988584e+07);
    num = upc_fma(num, r, 2.44404971329663284e+07);
    num = upc_fma(num, r, 4.94985450103831850e+06);
    num = upc_fma(num, r, 4.73334579451213649e+05);
    num = upc_fma(num, r, 2.23256080577635075e+04);
    num = upc_fma(num, r, 4.98914924494953766e+02);
    num = upc_fma(num, r, 4.19803057674432534e+00);
    double den = 5.87769839876705315e+06;
    den = upc_fma(den, r, 2.58257609111968540e+07);
    den = upc_fma(den, r, 2.63619050062270910e+07);
    den = upc_fma(den, r, 9.97100045574528910e+06);
    den = upc_fma(den, r, 1.69745287590994616e+06);
    den = upc_fma(den, r, 1.42817358525897551e+05);
    den = upc_fma(den, r, 6.10854034951519316e+03);
    den = upc_fma(den, r, 1.26415601550449168e+02);
    den = upc_fma(den, r, 1.0);
    const double z = (q * num) / den;
    return upc_max(-2.0, upc_min(2.0, z));
}

struct NoiseQuad
{
    double z[4];
};

/*
 * The standard truncated normals of axis pairs `pair` and `pair + 1` of one (particle, step, slot): z[0], z[1] = axes 2 pair,
 * 2 pair + 1; z[2], z[3] = axes 2 pair + 2, 2 pair + 3 (zeros when `two_blocks` is false).  The two Philox blocks and the four
 * quantiles are independent chains evaluated side by side, which is what hides their latency (a warp runs alone on its
 * scheduler for long stretches of this program).  One out-of-line copy per kernel: inlined at every use the generator made the
 * simulation program 30 - 35 % longer, past the instruction cache (profiles/README.md).
 */
#ifndef UPC_NOISE_QUAD
#define UPC_NOISE_QUAD 1
#endif
#if defined(__CUDACC__)
inline __host__ __device__ __noinline__
#else
inline
#endif
NoiseQuad noise_quad_values(uint64_t seed, uint64_t particle, uint32_t step, uint32_t slot, uint32_t pair, bool two_blocks)
{
    uint32_t a[4], b[4] = {0u, 0u, 0u, 0u};
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32);
    philox4x32_10(k0, k1, c0, c1, step, (slot << 16) | pair, a);
    if (two_blocks) philox4x32_10(k0, k1, c0, c1, step, (slot << 16) | (pair + 1u), b);
    NoiseQuad out;
    out.z[0] = truncated_normal_from_centered(uniform_centered52(a[0], a[1]));
    out.z[1] = truncated_normal_from_centered(uniform_centered52(a[2], a[3]));
    out.z[2] = 0.0;
    out.z[3] = 0.0;
    if (two_blocks)
    {
        out.z[2] = truncated_normal_from_centered(uniform_centered52(b[0], b[1]));
        out.z[3] = truncated_normal_from_centered(uniform_centered52(b[2], b[3]));
    }
    return out;
}

/* one pair (host-side tape generation and tests) */
UPC_NOISE_HD void noise_pair(uint64_t seed, uint64_t particle, uint32_t step, uint32_t slot, uint32_t pair, double* z0, double* z1)
{
    const NoiseQuad v = noise_quad_values(seed, particle, step, slot, pair, false);
    *z0 = v.z[0];
    *z1 = v.z[1];
}

} /* namespace upc */

#endif /* UPC_NOISE_CUH */
