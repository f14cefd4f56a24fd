// Sturm-count core of the bisection solver for bidiagonal singular values (shared by the CUDA kernel
// in bidiag_bisect.cu and by a host-side unit test; header-only, no CUDA dependencies).
//
// The singular values of the upper bidiagonal B (diagonal d_0..d_{n-1}, superdiagonal e_0..e_{n-2}) are
// the positive eigenvalues of the Golub-Kahan matrix: the 2n x 2n symmetric tridiagonal with zero
// diagonal and off-diagonal z = (d_0, e_0, d_1, e_1, ..., e_{n-2}, d_{n-1}).  For mu > 0 the number of
// eigenvalues below mu is n + #{sigma_i < mu}; it equals the number of sign changes of the Sturm
// sequence p_0 = 1, p_1 = -mu, p_i = -mu p_{i-1} - z_{i-2}^2 p_{i-2}.  The three-term form has one FMA
// on the dependent chain per step (the pivot form q_i = -mu - z^2/q_{i-1} has a division); the pair
// (p_{i-1}, p_i) is rescaled by a power of two every 4 steps.  Inputs are pre-scaled so that |z| <= 1
// and floored at 2^-50 (a 2^-100 floor on z^2 perturbs sigma by < 1e-15 sigma_max).
#pragma once
#if defined(__CUDACC__)
#define SVDB_HD __host__ __device__ __forceinline__
#else
#define SVDB_HD inline
#endif

namespace svdb200 {

constexpr double kBisZ2Floor = 7.888609052210118e-31;      // 2^-100
constexpr double kBisBig = 2.037035976334486e+90;           // 2^300
constexpr double kBisSmall = 4.909093465297727e-91;         // 2^-300

// number of singular values < mu (mu > 0); z2[0 .. 2n-2] = squared, scaled, floored off-diagonals
SVDB_HD int bisect_count(const double* z2, int n, double mu) {
    const int len = 2 * n - 1;
    double p0 = 1.0, p1 = -mu;
    bool prev_neg = true;        // sign of p1 (mu > 0)
    int changes = 1;             // p_0 -> p_1
    int i = 0;
    for (; i + 4 <= len; i += 4) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 4; ++u) {
            const double t = z2[i + u] * p0;
            const double p2 = -mu * p1 - t;
            const bool neg = (p2 < 0.0) || (p2 == 0.0 && !prev_neg);
            changes += (neg != prev_neg);
            prev_neg = neg;
            p0 = p1;
            p1 = p2;
        }
        const double a0 = p0 < 0 ? -p0 : p0, a1 = p1 < 0 ? -p1 : p1;
        const double mx = a0 > a1 ? a0 : a1;
        if (mx > kBisBig) { p0 *= kBisSmall; p1 *= kBisSmall; }
        else if (mx < kBisSmall) { p0 *= kBisBig; p1 *= kBisBig; }
    }
    for (; i < len; ++i) {
        const double t = z2[i] * p0;
        const double p2 = -mu * p1 - t;
        const bool neg = (p2 < 0.0) || (p2 == 0.0 && !prev_neg);
        changes += (neg != prev_neg);
        prev_neg = neg;
        p0 = p1;
        p1 = p2;
    }
    return changes - n;
}

// k-th smallest singular value of the scaled problem (all sigma in [0, 1]); k in [0, n)
SVDB_HD double bisect_kth(const double* z2, int n, int k, int max_iter, double rel_tol = 4.440892098500626e-16) {
    double lo = 0.0, hi = 1.0000000000000004;
    for (int it = 0; it < max_iter; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo && mid < hi)) break;
        if (bisect_count(z2, n, mid) <= k) lo = mid; else hi = mid;
        if (hi - lo <= rel_tol * hi || hi < 1e-18) break;
    }
    return 0.5 * (lo + hi);
}

}  // namespace svdb200
