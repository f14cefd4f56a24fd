// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" harness around the UNMODIFIED reference headers, compiled where they lie
// (-I/root/reference; nothing is copied into this repository).  It exposes the reference's own
// CPU implementation of the hot path on flat row-major buffers so that
//   (1) the C restatement in oracle/svd_oracle.c can be pinned bit-for-bit against it,
//   (2) golden vectors (tests/golden/) can be (re)generated,
//   (3) bench.py's `--impl reference` / `cpu_baseline` leg can time the reference on the host.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may load the
// resulting oracle/_ref/libsvdref*.so.
//
// Entry points follow the reference symbols:
//   csc586::parallel::brd_p1<T>   svd_parallel.h:411     (tile flat-tree dense -> band)
//   csc586::parallel::brd_p2<T>   svd_parallel.h:640     (band -> bidiagonal bulge chasing)
//   csc586::serial::qrd<float>    svd_serial.h:368       (zero-shift QR diagonalisation)
//   csc586::serial::householder   svd_serial.h:189
//   csc586::serial::brd<T>        svd_serial.h:233       (one-stage Golub-Kahan bidiagonalisation)
//   csc586::gpu::brd_p1           svd_cpu.h:370          (full-height panel dense -> band, float)
//   csc586::gpu::brd_p2<T>        svd_cpu.h:631
#include <cstring>
#include <chrono>
#include <vector>
#include "matrix.h"
#include "svd_serial.h"
#include "svd_parallel.h"
#include "svd_cpu.h"

namespace {
template <typename T>
void flatten_into(csc586::Matrix<T>& A, T* out) {
    for (size_t i = 0; i < A.nrows; ++i) std::memcpy(out + i * A.ncols, A[i].data(), A.ncols * sizeof(T));
}
template <typename T>
void flatten_into(csc586::gpu::Matrix<T>& A, T* out) {
    for (size_t i = 0; i < A.nrows; ++i) std::memcpy(out + i * A.ncols, A[i].data(), A.ncols * sizeof(T));
}
template <typename T>
int p1(T* a, size_t n, size_t t) {
    csc586::Matrix<T> A(a, n, n);
    A.parallel = false;  // the array ctor (matrix.h:110) leaves it uninitialised
    csc586::parallel::brd_p1<T>(A, t);
    flatten_into(A, a);
    return 0;
}
template <typename T>
int p2(T* a, size_t n, size_t b, T* d, T* e) {
    csc586::Matrix<T> A(a, n, n);
    A.parallel = false;
    auto B = csc586::parallel::brd_p2<T>(A, b);
    flatten_into(A, a);
    if (d) std::memcpy(d, B.d.data(), B.d.size() * sizeof(T));
    if (e) std::memcpy(e, B.e.data(), B.e.size() * sizeof(T));
    return 0;
}
// csc586::serial::brd<T> (svd_serial.h:233-266): the one-stage Golub-Kahan Householder bidiagonalisation
template <typename T>
int serial_brd(T* a, size_t n, T* d, T* e) {
    csc586::Matrix<T> A(a, n, n);
    A.parallel = false;
    auto B = csc586::serial::brd<T>(A);
    flatten_into(A, a);
    if (d) std::memcpy(d, B.d.data(), B.d.size() * sizeof(T));
    if (e) std::memcpy(e, B.e.data(), B.e.size() * sizeof(T));
    return 0;
}
template <typename T>
int hh(const T* x, size_t len, T* w, T* H, T* tau) {
    csc586::Matrix<T> X(x, len, 1);
    X.parallel = false;  // uninitialised otherwise -> operator*= would walk out of bounds
    auto R = csc586::serial::householder<T>(X);
    for (size_t i = 0; i < len; ++i) w[i] = R.w[i][0];
    flatten_into(R.transform, H);
    *tau = R.tau;
    return 0;
}
}  // namespace

extern "C" {

int svdref_brd_p1_f32(float* a, size_t n, size_t t) { return p1<float>(a, n, t); }
int svdref_brd_p1_f64(double* a, size_t n, size_t t) { return p1<double>(a, n, t); }
int svdref_brd_p2_f32(float* a, size_t n, size_t b, float* d, float* e) { return p2<float>(a, n, b, d, e); }
int svdref_brd_p2_f64(double* a, size_t n, size_t b, double* d, double* e) { return p2<double>(a, n, b, d, e); }
int svdref_serial_brd_f32(float* a, size_t n, float* d, float* e) { return serial_brd<float>(a, n, d, e); }
int svdref_serial_brd_f64(double* a, size_t n, double* d, double* e) { return serial_brd<double>(a, n, d, e); }
int svdref_householder_f32(const float* x, size_t len, float* w, float* H, float* tau) { return hh<float>(x, len, w, H, tau); }
int svdref_householder_f64(const double* x, size_t len, double* w, double* H, double* tau) { return hh<double>(x, len, w, H, tau); }

// serial::qrd is float-only in the reference (Rotation hard-codes float, svd_serial.h:60-65).
int svdref_qrd_f32(const float* d, const float* e, size_t n, float* d_out, float* e_out) {
    csc586::serial::Bidiagonal<float> B;
    B.d.assign(d, d + n);
    B.e.assign(e, e + n - 1);
    auto R = csc586::serial::qrd<float>(B);
    std::memcpy(d_out, R.d.data(), n * sizeof(float));
    if (e_out) std::memcpy(e_out, R.e.data(), (n - 1) * sizeof(float));
    return 0;
}

int svdref_zero_shift_f32(float* d, float* e, size_t n) {
    csc586::serial::Bidiagonal<float> B;
    B.d.assign(d, d + n);
    B.e.assign(e, e + n - 1);
    csc586::serial::impl_zero_shift<float>(B);
    std::memcpy(d, B.d.data(), n * sizeof(float));
    std::memcpy(e, B.e.data(), (n - 1) * sizeof(float));
    return 0;
}

// Panel (full-height) dense -> band, the algorithm the reference's CUDA files mirror. float only.
int svdref_gpu_brd_p1_f32(float* a, size_t n, size_t b) {
    csc586::gpu::Matrix<float> A(a, n, n);
    csc586::gpu::brd_p1(A, b);
    flatten_into(A, a);
    return 0;
}

// Same timer convention as timing.h:78-83: steady_clock around f(copy, b); copy excluded.
// Returns seconds for stage 1 and stage 2.  `chain` != 0 feeds stage 2 the stage-1 output (the
// meaningful chain); chain == 0 times stage 2 on a fresh dense copy like svd_cpu.cpp:234-238.
int svdref_time_multicore_f32(const float* a, size_t n, size_t b, int chain, double* t1, double* t2) {
    csc586::Matrix<float> A(a, n, n);
    A.parallel = false;
    auto x = A;
    auto s = std::chrono::steady_clock::now();
    csc586::parallel::brd_p1<float>(x, b);
    auto m = std::chrono::steady_clock::now();
    auto y = chain ? x : A;
    auto s2 = std::chrono::steady_clock::now();
    csc586::parallel::brd_p2<float>(y, b);
    auto e = std::chrono::steady_clock::now();
    *t1 = std::chrono::duration<double>(m - s).count();
    *t2 = std::chrono::duration<double>(e - s2).count();
    return 0;
}
int svdref_time_multicore_f64(const double* a, size_t n, size_t b, int chain, double* t1, double* t2) {
    csc586::Matrix<double> A(a, n, n);
    A.parallel = false;
    auto x = A;
    auto s = std::chrono::steady_clock::now();
    csc586::parallel::brd_p1<double>(x, b);
    auto m = std::chrono::steady_clock::now();
    auto y = chain ? x : A;
    auto s2 = std::chrono::steady_clock::now();
    csc586::parallel::brd_p2<double>(y, b);
    auto e = std::chrono::steady_clock::now();
    *t1 = std::chrono::duration<double>(m - s).count();
    *t2 = std::chrono::duration<double>(e - s2).count();
    return 0;
}

int svdref_omp_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
