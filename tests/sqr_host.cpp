// Host-side harness around svdsolver_b200/csrc/sqr_core.h (the implicit shifted QR core shared with the CUDA kernel):
// runs the multishift iteration with P sweeps per pass SEQUENTIALLY -- mathematically what the kernel's pipeline of P
// lagged sweeps computes -- so the CPU test-suite can check convergence and accuracy against LAPACK without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>
#include "sqr_core.h"
using namespace svdb200;

extern "C" long long sqr_all(const double* d_in, const double* e_in, int n, double* sigma_desc, int P, long long* passes_out) {
    std::vector<double> d(d_in, d_in + n), e(e_in, e_in + std::max(n - 1, 0)), z2(4 * (size_t)std::max(P, 2)), mu((size_t)std::max(P, 1));
    const double tol = 8.881784197001252e-16;   // 8 eps: a sweep leaves O(eps) noise in e, a tighter test never settles
    long long sweeps = 0, passes = 0;
    const long long max_sweeps = 60LL * n + 1000;
    int hi = n - 1;
    while (hi > 0 && sweeps < max_sweeps) {
        for (int i = 0; i < hi; ++i)
            if (e[i] != 0.0 && sqr_negligible(e[i], d[i], d[i + 1], tol)) e[i] = 0.0;
        while (hi > 0 && e[hi - 1] == 0.0) --hi;
        if (hi == 0) break;
        int lo = hi - 1;
        while (lo > 0 && e[lo - 1] != 0.0) --lo;
        const int nd = hi - lo + 1;
        {   // zero diagonal entries split the window before any sweep
            double dmax = 0.0;
            for (int i = lo; i <= hi; ++i) dmax = std::max(dmax, std::fabs(d[i]));
            bool split = false;
            for (int i = lo; i <= hi; ++i)
                if (std::fabs(d[i]) <= tol * dmax) {
                    d[i] = 0.0;
                    if (i < hi) sqr_chase_zero_row(d.data(), e.data(), i, hi);
                    else sqr_chase_zero_col(d.data(), e.data(), lo, hi);
                    split = true;
                }
            if (split) continue;
        }
        const int cnt = std::min(P, std::max(1, nd / 2));          // never the whole spectrum of the window: prod (B^T B - mu_k^2) must not vanish
        const int blk = std::min(nd, 2 * cnt);                       // shifts: the cnt smallest singular values of the trailing blk x blk block
        const double bound = sqr_fill_z2(d.data(), e.data(), hi - blk + 1, blk, z2.data());
        for (int k = 0; k < cnt; ++k) mu[k] = bound > 0 ? bisect_kth(z2.data(), blk, k, 60, 4.5e-16) * bound : 0.0;
        // sweep k carries the k-th smallest shift; a sweep that makes the last off-diagonal entry negligible deflates it at
        // once, so the sweeps behind it end one position earlier (the kernel publishes the new bottom to the pipeline)
        int bot = hi;
        for (int k = 0; k < cnt && bot > lo; ++k) {
            SqrCarry c = sqr_start(d[lo], e[lo], mu[k]);
            for (int i = lo; i < bot; ++i) sqr_position(d.data(), e.data(), i, lo, bot, c);
            ++sweeps;
            while (bot > lo && sqr_negligible(e[bot - 1], d[bot - 1], d[bot], tol)) { e[bot - 1] = 0.0; --bot; }
        }
        ++passes;
    }
    for (int i = 0; i < n; ++i) sigma_desc[i] = std::fabs(d[i]);
    std::sort(sigma_desc, sigma_desc + n, [](double a, double b) { return a > b; });
    if (passes_out) *passes_out = passes;
    return hi > 0 ? -sweeps : sweeps;
}
