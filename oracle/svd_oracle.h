/* TEST INFRASTRUCTURE ONLY -- C interface of the CPU oracle (oracle/libsvd_oracle.so).
 * All matrices are square n x n, dense row-major, updated in place.  Each entry point names
 * the reference function it restates (file:line in /root/reference). */
#ifndef SVD_ORACLE_H
#define SVD_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define SVDO_DECL(T, S)                                                                              \
    /* csc586::parallel::brd_p1<T>  svd_parallel.h:411-533 (tile flat-tree dense->band) */           \
    int svdo_brd_p1_##S(T* A, size_t n, size_t t);                                                   \
    /* csc586::gpu::brd_p1  svd_cpu.h:370-425 (full-height panel order; what cuda_brd_p1 mirrors) */ \
    int svdo_brd_p1_panel_##S(T* A, size_t n, size_t b);                                             \
    /* csc586::parallel::brd_p2<T>  svd_parallel.h:640-695 (band->bidiagonal; d,e optional) */       \
    int svdo_brd_p2_##S(T* A, size_t n, size_t band, T* d, T* e);                                    \
    size_t svdo_brd_p2_schedule_##S(size_t n, size_t band, long long* out, size_t cap);              \
    /* NOT the reference: same windows and arithmetic, every bulge chased to the end (checker for the complete schedule) */ \
    int svdo_brd_p2_complete_##S(T* A, size_t n, size_t band, T* d, T* e);                           \
    /* csc586::serial::brd<T>  svd_serial.h:233-266 (one-stage Golub-Kahan bidiagonalisation) */     \
    int svdo_brd_serial_##S(T* A, size_t n, T* d, T* e);                                             \
    /* csc586::serial::householder<T>  svd_serial.h:189-216 */                                       \
    int svdo_householder_##S(const T* x, size_t len, T* w, T* H, T* tau);                            \
    /* parallel::qr / lq  svd_parallel.h:133-226; qr_apply / lq_apply 243-281 */                      \
    int svdo_panel_qr_##S(T* A, size_t m, size_t n, T* S_, T* V);                                    \
    int svdo_panel_lq_##S(T* A, size_t m, size_t n, T* S_, T* U);                                    \
    int svdo_qr_apply_##S(T* A, size_t rows, size_t cols, const T* S_, const T* V, size_t t);        \
    int svdo_lq_apply_##S(T* A, size_t rows, size_t cols, const T* S_, const T* U, size_t t);        \
    /* serial::impl_zero_shift svd_serial.h:314-333; serial::qrd 368-422 (+Criteria 138-166) */      \
    void svdo_zero_shift_##S(T* d, T* e, size_t nd);                                                 \
    long long svdo_qrd_##S(T* d, T* e, size_t n, T* threshold_out, unsigned long long* max_iter_out);\
    /* gpu::Matrix<T>::mse  matrix_gpu.h:438-453 */                                                  \
    T svdo_mse_##S(const T* a, const T* b, size_t n, size_t band);
SVDO_DECL(float, f32)
SVDO_DECL(double, f64)
#undef SVDO_DECL
#ifdef __cplusplus
}
#endif
#endif
