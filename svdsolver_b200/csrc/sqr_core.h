// Implicit SHIFTED QR for the singular values of a bidiagonal (the reference's serial::qrd is zero-shift
// only, svd_serial.h:314-333): core shared by the CUDA kernel (bidiag_sqr.cu) and a host-side unit test; header-only.
//
// One sweep = one Golub-Kahan SVD step with shift mu on the unreduced window d[lo..hi], e[lo..hi-1]: the bulge created by
// the first right rotation (built from (d_lo^2 - mu^2)/d_lo and e_lo) is chased to the bottom by alternating right and
// left Givens rotations.  A sweep touches position i (entries d_i, e_i, d_{i+1}, e_{i+1}, and e_{i-1}) once, top to
// bottom, carrying two scalars (f, g): sweep s+1 may work on position i as soon as sweep s has finished position i+2.
// The kernel therefore pipelines P sweeps (lane l at position t - 3 l), which is the multishift QR iteration with P
// small bulges: its P shifts are the singular values of the trailing P x P block of the window (bisection, bisect_core.h),
// the generalisation of the Wilkinson shift (P = 1: the trailing 2 x 2).
#pragma once
#include <math.h>
#include "bisect_core.h"

namespace svdb200 {

struct SqrCarry { double f, g; };

SVDB_HD void sqr_lartg(double f, double g, double& c, double& s, double& r) {
    if (g == 0.0) { c = 1.0; s = 0.0; r = f; return; }
    if (f == 0.0) { c = 0.0; s = 1.0; r = g; return; }
    const double af = fabs(f), ag = fabs(g);
    const double sc = af > ag ? af : ag;
    const double fs = f / sc, gs = g / sc;
    const double h = sc * sqrt(fs * fs + gs * gs);
    r = h; c = f / h; s = g / h;
}

// start of a sweep on the window whose first diagonal / off-diagonal entries are d0, e0
SVDB_HD SqrCarry sqr_start(double d0, double e0, double mu) {
    SqrCarry k;
    if (d0 != 0.0) k.f = (fabs(d0) - mu) * ((d0 >= 0.0 ? 1.0 : -1.0) + mu / d0);
    else k.f = -mu * mu;
    k.g = e0;
    return k;
}

// position i of a sweep (lo <= i < hi); d, e are the full arrays.  LAPACK dbdsqr's shifted top-to-bottom chase.
template <typename T>
SVDB_HD void sqr_position(T* d, T* e, int i, int lo, int hi, SqrCarry& k) {
    double cr, sr, cl, sl, r;
    const double di = (double)d[i], ei = (double)e[i], dn = (double)d[i + 1];
    sqr_lartg(k.f, k.g, cr, sr, r);
    if (i > lo) e[i - 1] = (T)r;
    double f = cr * di + sr * ei;
    const double ei2 = cr * ei - sr * di;
    double g = sr * dn;
    const double dn2 = cr * dn;
    sqr_lartg(f, g, cl, sl, r);
    d[i] = (T)r;
    f = cl * ei2 + sl * dn2;
    d[i + 1] = (T)(cl * dn2 - sl * ei2);
    if (i < hi - 1) {
        const double en = (double)e[i + 1];
        g = sl * en;
        e[i + 1] = (T)(cl * en);
        e[i] = (T)f;                 // overwritten by the next position's first rotation (kept consistent for readers)
    } else {
        e[i] = (T)f;
        g = 0.0;
    }
    k.f = f; k.g = g;
}

// A (numerically) zero diagonal entry d[i], lo <= i < hi, splits the window: the entry e[i] of its row is rotated away to
// the right with left rotations (Golub & Van Loan 8.6.2), after which row i is zero and e[i] = 0.  i == hi (zero last
// diagonal entry): e[hi-1] is rotated away upwards with right rotations.
template <typename T>
SVDB_HD void sqr_chase_zero_row(T* d, T* e, int i, int hi) {
    double f = (double)e[i];
    e[i] = (T)0;
    for (int k = i + 1; k <= hi && f != 0.0; ++k) {
        double c, s, r;
        sqr_lartg((double)d[k], f, c, s, r);
        d[k] = (T)r;
        if (k < hi) {
            const double ek = (double)e[k];
            f = -s * ek;
            e[k] = (T)(c * ek);
        }
    }
}
template <typename T>
SVDB_HD void sqr_chase_zero_col(T* d, T* e, int lo, int hi) {
    double f = (double)e[hi - 1];
    e[hi - 1] = (T)0;
    for (int k = hi - 1; k >= lo && f != 0.0; --k) {
        double c, s, r;
        sqr_lartg((double)d[k], f, c, s, r);
        d[k] = (T)r;
        if (k > lo) {
            const double ek = (double)e[k - 1];
            f = -s * ek;
            e[k - 1] = (T)(c * ek);
        }
    }
}

// negligible off-diagonal entry (absolute-accuracy criterion of the shifted iteration)
template <typename T>
SVDB_HD bool sqr_negligible(T ei, T di, T dn, double tol) {
    return fabs((double)ei) <= tol * (fabs((double)di) + fabs((double)dn));
}

// shifts of a pass: the `cnt` singular values of the trailing cnt x cnt block d[hi-cnt+1..hi], e[hi-cnt+1..hi-1] of the
// window, k-th smallest for k = 0..cnt-1 (z2: scratch of 2*cnt doubles, filled by the caller through sqr_fill_z2)
template <typename T>
SVDB_HD double sqr_fill_z2(const T* d, const T* e, int first, int cnt, double* z2) {
    double bound = 0.0;
    for (int i = 0; i < cnt; ++i) {
        const double di = fabs((double)d[first + i]);
        const double er = i < cnt - 1 ? fabs((double)e[first + i]) : 0.0;
        const double el = i > 0 ? fabs((double)e[first + i - 1]) : 0.0;
        const double m = di + (er > el ? er : el);
        bound = m > bound ? m : bound;
    }
    if (bound == 0.0) return 0.0;
    const double inv = 1.0 / bound;
    for (int i = 0; i < cnt; ++i) {
        const double a = (double)d[first + i] * inv;
        z2[2 * i] = a * a > kBisZ2Floor ? a * a : kBisZ2Floor;
        if (i < cnt - 1) {
            const double b = (double)e[first + i] * inv;
            z2[2 * i + 1] = b * b > kBisZ2Floor ? b * b : kBisZ2Floor;
        }
    }
    return bound;
}

}  // namespace svdb200
