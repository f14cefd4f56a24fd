// Host-side harness around svdsolver_b200/csrc/bisect_core.h (header-only Sturm-count core shared with the CUDA kernel):
// lets the CPU test-suite check the numerics of the bisection solver against LAPACK without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "bisect_core.h"
extern "C" void bis_all(const double* d, const double* e, int n, double* sigma_desc) {
    double bound = 0;
    for (int i = 0; i < n; ++i) { double r = std::fabs(d[i]) + (i < n - 1 ? std::fabs(e[i]) : 0.0); double c = std::fabs(d[i]) + (i > 0 ? std::fabs(e[i-1]) : 0.0); if (r > bound) bound = r; if (c > bound) bound = c; }
    if (bound == 0) { for (int i = 0; i < n; ++i) sigma_desc[i] = 0; return; }
    std::vector<double> z2(2 * n);
    for (int i = 0; i < n; ++i) {
        double a = d[i] / bound; a = a * a; if (a < svdb200::kBisZ2Floor) a = svdb200::kBisZ2Floor; z2[2 * i] = a;
        if (i < n - 1) { double b = e[i] / bound; b = b * b; if (b < svdb200::kBisZ2Floor) b = svdb200::kBisZ2Floor; z2[2 * i + 1] = b; }
    }
    for (int k = 0; k < n; ++k) sigma_desc[n - 1 - k] = svdb200::bisect_kth(z2.data(), n, k, 200) * bound;
}
