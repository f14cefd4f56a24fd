/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for the SVDSolver hot path (see svd_oracle_impl.h).
 * Never linked into, imported by, or called from the product library (svdsolver_b200/).
 * Parity status: float and double stage 1 / stage 2 and float QR diagonalisation are PINNED
 * bit-for-bit against the reference's fixtures and against oracle/_ref; the DOUBLE QR
 * diagonalisation is "parity unpinned" (the reference's serial::qrd does not compile for double,
 * svd_serial.h:60-65,321) and is cross-checked against LAPACK instead. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include "svd_oracle.h"

#define T float
#define FN(name) name##_f32
#define SQRT_T sqrtf
#define FABS_T fabsf
#include "svd_oracle_impl.h"
#undef T
#undef FN
#undef SQRT_T
#undef FABS_T

#define T double
#define FN(name) name##_f64
#define SQRT_T sqrt
#define FABS_T fabs
#include "svd_oracle_impl.h"
#undef T
#undef FN
#undef SQRT_T
#undef FABS_T

/* gpu::Matrix::mse, matrix_gpu.h:438-453: float accumulators, sign-insensitive, `band` diagonals
 * starting at the main one, divided by band*nrows. */
float svdo_mse_f32(const float* a, const float* b, size_t n, size_t band) {
    float error = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        size_t je = i + band < n ? i + band : n;
        for (size_t j = i; j < je; ++j) {
            float d = fabsf(a[i * n + j]) - fabsf(b[i * n + j]);
            error += (float)sqrt(pow((double)d, 2));
        }
    }
    return error / (float)(band * n);
}
double svdo_mse_f64(const double* a, const double* b, size_t n, size_t band) {
    float error = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        size_t je = i + band < n ? i + band : n;
        for (size_t j = i; j < je; ++j) {
            double d = fabs(a[i * n + j]) - fabs(b[i * n + j]);
            error += sqrt(pow(d, 2));
        }
    }
    return error / (band * n);
}
