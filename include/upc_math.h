/*
 * upc_math.h - deterministic elementary functions shared by the CUDA kernels and the CPU oracle.
 *
 * Why this exists: the reference computes angle wraps, SE(3) log/exp and quaternion distances through
 * libm / Eigen / arc_utilities (EnforceContinuousRevoluteBounds at simple_robot_models.hpp:115,1083;
 * ContinuousRevoluteSignedDistance :415,1112; TwistBetweenTransforms / ExpTwist :657-666,759;
 * EigenHelpers::Distance :873).  glibc and the CUDA math library do not round sin/cos/atan2 identically,
 * so cluster memberships and collision flags could not be bit-exact between a CPU oracle and the GPU.
 * Everything here is built only from IEEE-754 exact operations (+ - * / sqrt fma floor rint fmod), which
 * round identically on x86-64 and sm_100a when contraction is disabled (nvcc -fmad=false,
 * g++ -ffp-contract=off) and fused multiply-adds are written explicitly.
 *
 * Polynomial kernels follow the classic published minimax forms for sin/cos on [-pi/4,pi/4] and atan with
 * four break points; coefficients are checked against mpmath in tests/test_math.py.
 */
#ifndef UPC_MATH_H
#define UPC_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define UPC_HD __host__ __device__ __forceinline__
/* larger leaf functions: one out-of-line copy per kernel keeps code size (instruction-cache pressure) and registers down */
#define UPC_HD_CALL inline __host__ __device__ __noinline__
#else
#define UPC_HD inline
#define UPC_HD_CALL inline
#endif

#define UPC_PI 3.14159265358979323846
#define UPC_TWO_PI 6.28318530717958647692
#define UPC_HALF_PI 1.57079632679489661923

UPC_HD double upc_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

UPC_HD double upc_rint(double x)
{
#if defined(__CUDA_ARCH__)
    return rint(x);
#else
    return __builtin_rint(x);
#endif
}

UPC_HD double upc_abs(double x) { return fabs(x); }

/* min/max written as selects so that host and device agree on every input (no NaN inputs by contract). */
UPC_HD double upc_min(double a, double b) { return (b < a) ? b : a; }
UPC_HD double upc_max(double a, double b) { return (a < b) ? b : a; }

/* sin and cos kernels on |r| <= pi/4 */
UPC_HD double upc_ksin(double r)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = r * r;
    double p = upc_fma(z, S6, S5);
    p = upc_fma(z, p, S4);
    p = upc_fma(z, p, S3);
    p = upc_fma(z, p, S2);
    p = upc_fma(z, p, S1);
    return upc_fma(r * z, p, r);
}

UPC_HD double upc_kcos(double r)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = r * r;
    double p = upc_fma(z, C6, C5);
    p = upc_fma(z, p, C4);
    p = upc_fma(z, p, C3);
    p = upc_fma(z, p, C2);
    p = upc_fma(z, p, C1);
    /* 1 - z/2 + z^2 * p */
    return upc_fma(z * z, p, upc_fma(-0.5, z, 1.0));
}

/* Simultaneous sine and cosine. Accurate (about 1 ulp) for |x| up to about 1e5; beyond that the two-term
 * Cody-Waite reduction degrades gracefully. Angles on this path are bounded by a few pi. */
UPC_HD void upc_sincos(double x, double* s, double* c)
{
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_HI = 1.57079632673412561417e+00; /* first 33 bits of pi/2 */
    const double PIO2_LO = 6.07710050650619224932e-11; /* pi/2 - PIO2_HI */
    const double k = upc_rint(x * TWO_OVER_PI);
    double r = upc_fma(-k, PIO2_HI, x);
    r = upc_fma(-k, PIO2_LO, r);
    const long long q = (long long)k;
    const double sr = upc_ksin(r);
    const double cr = upc_kcos(r);
    switch ((int)(q & 3LL))
    {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}

UPC_HD double upc_sin(double x) { double s, c; upc_sincos(x, &s, &c); return s; }
UPC_HD double upc_cos(double x) { double s, c; upc_sincos(x, &s, &c); return c; }

/*
 * atan2 for finite inputs, atan2(0,0) = 0.  The classic four-break-point reduction of atan is applied to the
 * pair (|y|, |x|) directly, so the whole function costs ONE division (an FP64 division is ~130 dependent
 * cycles on sm_100a): t = (|y| - c|x|) / (|x| + c|y|) for c in {0, 1/2, 1, 3/2, inf}, then the odd/even split
 * minimax polynomial in t.
 */
UPC_HD double upc_atan2(double y, double x)
{
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
                 aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
                 aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
                 aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
                 aT10 = 1.62858201153657823623e-02;
    const double ay = fabs(y);
    const double ax = fabs(x);
    double a = 0.0;
    if (!((ax == 0.0) && (ay == 0.0)))
    {
        double num, den, hi = 0.0, lo = 0.0;
        bool reduced = true;
        if (ay < 0.4375 * ax)
        {
            num = ay; den = ax; reduced = false;
        }
        else if (ay < 0.6875 * ax)
        {
            hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17;
            num = 2.0 * ay - ax; den = 2.0 * ax + ay;
        }
        else if (ay < 1.1875 * ax)
        {
            hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17;
            num = ay - ax; den = ay + ax;
        }
        else if (ay < 2.4375 * ax)
        {
            hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17;
            num = upc_fma(-1.5, ax, ay); den = upc_fma(1.5, ay, ax);
        }
        else
        {
            hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17;
            num = -ax; den = ay;
        }
        const double t = num / den;
        const double z = t * t;
        const double w = z * z;
        double s1 = upc_fma(w, aT10, aT8);
        s1 = upc_fma(w, s1, aT6);
        s1 = upc_fma(w, s1, aT4);
        s1 = upc_fma(w, s1, aT2);
        s1 = upc_fma(w, s1, aT0);
        s1 = z * s1;
        double s2 = upc_fma(w, aT9, aT7);
        s2 = upc_fma(w, s2, aT5);
        s2 = upc_fma(w, s2, aT3);
        s2 = upc_fma(w, s2, aT1);
        s2 = w * s2;
        if (!reduced)
        {
            a = t - t * (s1 + s2);
        }
        else
        {
            a = hi - ((t * (s1 + s2) - lo) - t);
        }
        if (x < 0.0)
        {
            a = UPC_PI - a;
        }
    }
    return (y < 0.0) ? -a : a;
}

UPC_HD double upc_atan(double x) { return upc_atan2(x, 1.0); }

/* arccosine for |x| <= 1 through atan2(sqrt((1-x)(1+x)), x) */
UPC_HD double upc_acos(double x)
{
    const double sn = sqrt((1.0 - x) * (1.0 + x));
    return upc_atan2(sn, x);
}

/* raw IEEE-754 bits of a double and back (memcpy semantics on the host, register moves on the device) */
UPC_HD uint64_t upc_double_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b;
    __builtin_memcpy(&b, &x, sizeof(b));
    return b;
#endif
}

UPC_HD double upc_bits_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double x;
    __builtin_memcpy(&x, &b, sizeof(x));
    return x;
#endif
}

/*
 * Natural logarithm of a positive, finite, normal double.  The classic argument reduction x = 2^k (1 + f) with
 * sqrt(2)/2 < 1 + f < sqrt(2), s = f / (2 + f), and the degree-14 minimax polynomial in s (the published fdlibm form);
 * only exact IEEE operations, one division, no fused operations - identical bits on host and device.
 * Used by the counter-based noise generator, whose arguments lie in [0.0227, 0.075].
 */
UPC_HD double upc_log(double x)
{
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double LG1 = 6.666666666666735130e-01, LG2 = 3.999999999940941908e-01, LG3 = 2.857142874366239149e-01,
                 LG4 = 2.222219843214978396e-01, LG5 = 1.818357216161805012e-01, LG6 = 1.531383769920937332e-01,
                 LG7 = 1.479819860511658591e-01;
    uint64_t bits = upc_double_bits(x);
    int k = (int)((bits >> 52) & 0x7ffu) - 1023;
    uint64_t mant = bits & 0x000fffffffffffffull;
    /* mantissa above sqrt(2): halve it and bump the exponent */
    if (mant >= 0x6a09e667f3bcdull)
    {
        k += 1;
        bits = mant | 0x3fe0000000000000ull;
    }
    else
    {
        bits = mant | 0x3ff0000000000000ull;
    }
    const double f = upc_bits_double(bits) - 1.0;
    const double dk = (double)k;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (LG2 + w * (LG4 + w * LG6));
    const double t2 = z * (LG1 + w * (LG3 + w * (LG5 + w * LG7)));
    const double r = t2 + t1;
    const double hfsq = (0.5 * f) * f;
    return dk * LN2_HI - ((hfsq - (s * (hfsq + r) + dk * LN2_LO)) - f);
}

/* Angle wrap into (-pi, pi]; restates EigenHelpers::EnforceContinuousRevoluteBounds as used at
 * simple_robot_models.hpp:115 and :1083 (source un-vendored; this is the normative form here). */
UPC_HD double upc_wrap_angle(double v)
{
    if ((v <= -UPC_PI) || (v > UPC_PI))
    {
        const double r = fmod(v, UPC_TWO_PI);
        if (r <= -UPC_PI)
        {
            return r + UPC_TWO_PI;
        }
        else if (r > UPC_PI)
        {
            return r - UPC_TWO_PI;
        }
        return r;
    }
    return v;
}

/* Shortest signed rotation from a to b, in (-pi, pi]; restates ContinuousRevoluteSignedDistance as used
 * at simple_robot_models.hpp:415 and :1112. */
UPC_HD double upc_angle_signed_distance(double a, double b)
{
    const double raw = upc_wrap_angle(b) - upc_wrap_angle(a);
    if (raw <= -UPC_PI)
    {
        return raw + UPC_TWO_PI;
    }
    else if (raw > UPC_PI)
    {
        return raw - UPC_TWO_PI;
    }
    return raw;
}

#endif /* UPC_MATH_H */
