"""Fits the single-branch rational approximation used by the counter-based noise generator:
    z(q) = Phi^-1(1/2 + q) ~= q * P(r) / Q(r),  r = 0.228 - q^2 >= 0,  |q| <= (Phi(2) - Phi(-2)) / 2 = 0.47725
(the shifted variable keeps every coefficient positive, so Horner's scheme does not cancel - the device of AS 241)
(the inverse normal CDF restricted to the +-2 sigma the truncated distributions of simple_uncertainty_models.hpp:29, 59 use).
Iteratively reweighted linear least squares (Loeb) in 60-digit arithmetic on Chebyshev nodes; prints the coefficients as
C literals and the maximum absolute error of the double-precision evaluation against mpmath on a dense grid.
Run once; the coefficients are pasted into csrc/upc_noise.cuh and oracle/upc_oracle_noise.cpp."""
import sys

import mpmath as mp

mp.mp.dps = 60
M = int(sys.argv[1]) if len(sys.argv) > 1 else 7
HALF_W = (mp.ncdf(2) - mp.ncdf(-2)) / 2
SMAX = HALF_W ** 2
SHIFT = mp.mpf("0.228")


def target(s):
    """Phi^-1(1/2 + q) / q at s = q^2"""
    if s <= 0:
        return mp.sqrt(2 * mp.pi)
    q = mp.sqrt(s)
    return mp.sqrt(2) * mp.erfinv(2 * q) / q


def fit(m, n=400, iters=25):
    lo = SHIFT - SMAX
    nodes = [lo + (SHIFT - lo) * (1 - mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n)) / 2 for k in range(n)]   # values of r
    f = [target(SHIFT - r) for r in nodes]
    w = [mp.mpf(1)] * n
    for _ in range(iters):
        # unknowns p0..pm, q1..qm: sum p_i s^i - f * sum q_j s^j = f
        rows = []
        for s, fv, wv in zip(nodes, f, w):
            pw = [s ** i for i in range(m + 1)]
            rows.append([wv * x for x in pw] + [-wv * fv * x for x in pw[1:]])
        A = mp.matrix(rows)
        b = mp.matrix([wv * fv for fv, wv in zip(f, w)])
        sol = mp.lu_solve(A.T * A, A.T * b)
        p = [sol[i] for i in range(m + 1)]
        qd = [mp.mpf(1)] + [sol[m + 1 + j] for j in range(m)]
        Q = [sum(qd[j] * s ** j for j in range(m + 1)) for s in nodes]
        w = [1 / abs(x) for x in Q]
    return p, qd


def main():
    p, qd = fit(M)
    pf, qf = [float(x) for x in p], [float(x) for x in qd]
    worst = 0.0
    N = 20001
    for k in range(N):
        q = float(HALF_W) * (2.0 * k / (N - 1) - 1.0)
        if q == 0.0:
            continue
        s = 0.228 - q * q
        num, den = pf[-1], qf[-1]
        for i in range(M - 1, -1, -1):
            num = num * s + pf[i]       # (the C code uses fma: at most one rounding less per step)
            den = den * s + qf[i]
        z = q * num / den
        want = mp.sqrt(2) * mp.erfinv(2 * mp.mpf(q))
        worst = max(worst, abs(float(z - want)))
    print("degree %d/%d, max abs error of the double evaluation: %.3e" % (M, M, worst))
    print("P:", ", ".join("%.17e" % x for x in pf))
    print("Q:", ", ".join("%.17e" % x for x in qf))


if __name__ == "__main__":
    main()
