// svdb200_matrix.hpp -- source-compatible C++ host surface of the SVDSolver hot path on top of the
// C ABI (svdb200.h).  A caller of the reference that includes matrix.h / matrix_gpu.h / timing.h and
// calls
//     csc586::gpu::cuda_brd_p1(A, band)            (svd_cuda_1.cu:750, svd_cuda_2.cu:1117)
//     csc586::parallel::brd_p1<T>(A, band)          (svd_parallel.h:411)
//     csc586::parallel::brd_p2<T>(A, band)          (svd_parallel.h:640)
//     csc586::gpu::brd_p2<T>(A, band + 1)           (svd_cpu.h:631)
//     csc586::serial::qrd<T>(B)                     (svd_serial.h:368)
//     csc586::benchmark::benchmark(f, instances, b) (timing.h:55)
// can include this header instead and link libsvdb200.so.
//
// This is NOT the reference's container: the reference stores a matrix as vector<vector<T>> and
// flatten()s it before every upload (matrix.h:82,258; svd_cuda_2.cu:549).  Here a matrix is ONE
// contiguous row-major buffer (what the device consumes), rows are views, and the same class
// serves both namespaces.  Same member names, argument meaning and assert()-style error behaviour
// for the members the drivers of the path use (matrix.h / matrix_gpu.h line numbers in comments).
#ifndef SVDB200_MATRIX_HPP
#define SVDB200_MATRIX_HPP

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "svdb200.h"

namespace csc586 {

// half-open index ranges, matrix.h:41-51
struct Slice {
    size_t i1, i2, j1, j2;
    bool contains(const Slice s) const { return (s.i2 - s.i1 <= i2 - i1) && (s.j2 - s.j1 <= j2 - j1); }
};

// matrix.h:59-62 (accumulated in T, index order)
template <typename T>
T norm(const std::vector<T>& v) {
    T acc = 0;
    for (const T& x : v) acc += x * x;
    return std::sqrt(acc);
}

namespace detail {
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
}  // namespace detail

template <typename T>
class Matrix {
    std::vector<T> data_;   // row-major, contiguous: the layout the device consumes

public:
    size_t nrows = 0, ncols = 0;
    bool parallel = false;   // kept for source compatibility (matrix.h:86); unused

    // bounds-checked row view: A[i][j], A[i].at(j), iteration (matrix.h:123 returns the row vector)
    struct Row {
        T* p; size_t n;
        T& operator[](size_t j) { assert(j < n && "column index out of range"); return p[j]; }
        T& at(size_t j) { if (j >= n) throw std::out_of_range("Matrix column"); return p[j]; }
        T* begin() { return p; }
        T* end() { return p + n; }
        T* data() { return p; }
        size_t size() const { return n; }
    };
    struct ConstRow {
        const T* p; size_t n;
        const T& operator[](size_t j) const { assert(j < n && "column index out of range"); return p[j]; }
        const T* begin() const { return p; }
        const T* end() const { return p + n; }
        size_t size() const { return n; }
    };

    Matrix() {}
    Matrix(const size_t& row_dim, const size_t& col_dim, const bool par = false)   // matrix.h:91
        : data_(row_dim * col_dim, T(0)), nrows(row_dim), ncols(col_dim), parallel(par) {}
    Matrix(const T* arr, size_t rows, size_t cols) : data_(arr, arr + rows * cols), nrows(rows), ncols(cols) {}   // matrix.h:110

    Row operator[](size_t i) { if (i >= nrows) throw std::out_of_range("Matrix row"); return Row{data_.data() + i * ncols, ncols}; }
    ConstRow operator[](size_t i) const { if (i >= nrows) throw std::out_of_range("Matrix row"); return ConstRow{data_.data() + i * ncols, ncols}; }
    T* data() { return data_.data(); }               // the flat buffer handed to the C ABI
    const T* data() const { return data_.data(); }
    size_t size() const { return nrows * ncols; }     // matrix.h:203

    Matrix<T>& operator+=(const Matrix<T>& m) {       // matrix.h:126
        assert(nrows == m.nrows && ncols == m.ncols && "dimension mismatch");
        for (size_t i = 0; i < data_.size(); ++i) data_[i] += m.data_[i];
        return *this;
    }
    Matrix<T>& operator-=(const Matrix<T>& m) {       // matrix.h:152
        assert(nrows == m.nrows && ncols == m.ncols && "dimension mismatch");
        for (size_t i = 0; i < data_.size(); ++i) data_[i] -= m.data_[i];
        return *this;
    }
    Matrix<T>& operator*=(const T alpha) {            // matrix.h:177
        for (T& x : data_) x *= alpha;
        return *this;
    }

    Matrix<T> transpose(const bool = false) const {   // matrix.h:209
        Matrix<T> t(ncols, nrows);
        for (size_t i = 0; i < nrows; ++i)
            for (size_t j = 0; j < ncols; ++j) t.data_[j * nrows + i] = data_[i * ncols + j];
        return t;
    }
    // A*B with k-ascending accumulation in T starting from 0 (matrix.h:234-248)
    Matrix<T> mm(const Matrix<T>& M, const bool = false) const {
        assert(ncols == M.nrows && "Matrix 1 col dim must match Matrix 2 row dim.");
        Matrix<T> r(nrows, M.ncols);
        for (size_t i = 0; i < nrows; ++i)
            for (size_t j = 0; j < M.ncols; ++j) {
                T acc = 0;
                for (size_t k = 0; k < ncols; ++k) acc += data_[i * ncols + k] * M.data_[k * M.ncols + j];
                r.data_[i * M.ncols + j] = acc;
            }
        return r;
    }
    Matrix<T> flatten(const bool transposed = false) const {   // matrix.h:258
        Matrix<T> f(1, size());
        if (!transposed) f.data_ = data_;
        else
            for (size_t i = 0; i < nrows; ++i)
                for (size_t j = 0; j < ncols; ++j) f.data_[j * nrows + i] = data_[i * ncols + j];
        return f;
    }
    Matrix<T> reshape(const size_t& m, const size_t& n) const {   // matrix_gpu.h:245
        assert(m * n == size() && "Reshape dimensions must match matrix size.");
        Matrix<T> r(m, n);
        r.data_ = data_;
        return r;
    }
    // copy(src, s, t): slice s of src -> slice t of this (matrix.h:277)
    void copy(const Matrix<T>& src, Slice s, Slice t) {
        assert(t.contains(s) && "Slice range from source outside target range.");
        for (size_t i = s.i1, it = t.i1; i < s.i2; ++i, ++it)
            std::copy(src.data_.begin() + i * src.ncols + s.j1, src.data_.begin() + i * src.ncols + s.j2,
                      data_.begin() + it * ncols + t.j1);
    }
    void copy(const Matrix<T>& src, Slice t) {                    // matrix.h:290
        assert(t.i2 - t.i1 <= nrows && t.j2 - t.j1 <= ncols && "Copy range from source outside target size.");
        for (size_t i = 0; i < src.nrows; ++i)
            std::copy(src.data_.begin() + i * src.ncols, src.data_.begin() + (i + 1) * src.ncols, data_.begin() + (t.i1 + i) * ncols + t.j1);
    }
    void copy(const Matrix<T>& src) { copy(src, Slice{0, src.nrows, 0, src.ncols}); }   // matrix.h:307
    void row_concat(const Matrix<T>& B) {                         // matrix.h:316
        assert(B.ncols == ncols && "Column dimensions must match for row concatenation.");
        data_.insert(data_.end(), B.data_.begin(), B.data_.end());
        nrows += B.nrows;
    }
    void col_concat(const Matrix<T>& B) {                         // matrix.h:324
        assert(B.nrows == nrows && "Row dimensions must match for column concatenation.");
        Matrix<T> r(nrows, ncols + B.ncols);
        r.copy(*this, Slice{0, nrows, 0, ncols});
        r.copy(B, Slice{0, nrows, ncols, ncols + B.ncols});
        *this = r;
    }
    void fill(const T value, const Slice t) {                     // matrix.h:334 (with the intended column start)
        for (size_t i = t.i1; i < t.i2; ++i) std::fill(data_.begin() + i * ncols + t.j1, data_.begin() + i * ncols + t.j2, value);
    }
    // U[min,max) fill (matrix.h:350, matrix_gpu.h:336).  The reference seeds a fresh mt19937 from
    // random_device per element and is not reproducible; this uses the documented splitmix64 stream
    // (svdsolver_b200/synth.py) so host, device and oracle agree.  `seed` selects the stream.
    void fill(const T& min_val, const T& max_val, uint64_t seed = 586) {
        for (size_t i = 0; i < data_.size(); ++i) {
            double u = (double)(detail::splitmix64(seed + i) >> 11) * (1.0 / 9007199254740992.0);
            data_[i] = (T)((double)min_val + ((double)max_val - (double)min_val) * u);
        }
    }
    std::vector<T> diag(size_t offset = 0) const {                // matrix.h:366
        std::vector<T> d(ncols - offset, T(0));
        for (size_t i = 0; i + offset < ncols && i < nrows; ++i) d[i] = data_[i * ncols + i + offset];
        return d;
    }
    Matrix<T> slice(const size_t r0, const size_t r1, const size_t c0, const size_t c1) const {   // matrix.h:376
        Matrix<T> s(r1 - r0, c1 - c0);
        for (size_t i = r0; i < r1; ++i)
            std::copy(data_.begin() + i * ncols + c0, data_.begin() + i * ncols + c1, s.data_.begin() + (i - r0) * s.ncols);
        return s;
    }
    Matrix<T> slice(const Slice& s) const { return slice(s.i1, s.i2, s.j1, s.j2); }               // matrix.h:391
    Matrix<T> get_tile(const size_t i, const size_t j, const size_t nbt) const {                  // matrix.h:406
        size_t t = nrows / nbt;
        assert((i + 1) * t <= nrows && (j + 1) * t <= ncols && "Tile out of range of matrix.");
        return slice(i * t, (i + 1) * t, j * t, (j + 1) * t);
    }
    void set_tile(const Matrix<T>& tile, const size_t i, const size_t j, const size_t nbt) {      // matrix.h:418
        size_t t = nrows / nbt;
        assert((i + 1) * t <= nrows && (j + 1) * t <= ncols && "Tile out of range of matrix.");
        copy(tile, Slice{i * t, (i + 1) * t, j * t, (j + 1) * t});
    }
    std::vector<T> col_slice(size_t j, size_t r0, size_t r1) const {                              // matrix.h:441
        assert(r1 > r0 && "Slice start must be less than slice end.");
        std::vector<T> v(r1 - r0);
        for (size_t i = r0; i < r1; ++i) v[i - r0] = data_[i * ncols + j];
        return v;
    }
    // reference error metric (matrix_gpu.h:438-453): sum over `band_size` diagonals of | |a|-|b| |,
    // float accumulators, divided by band_size*nrows; sign-insensitive.
    T mse(const Matrix<T>& B, size_t const band_size) const {
        assert(nrows == B.nrows && ncols == B.ncols && "Matrices must have identical dimensions.");
        float error = 0.0f;
        for (size_t i = 0; i < nrows; ++i)
            for (size_t j = i; j < std::min(i + band_size, ncols); ++j)
                error += (float)std::sqrt(std::pow(std::abs(data_[i * ncols + j]) - std::abs(B.data_[i * ncols + j]), 2));
        return (T)(error / (band_size * nrows));
    }
    // raw row-major binary I/O.  write truncates (matrix_gpu.h:463).  read uses sizeof(T) per element:
    // the reference reads sizeof(float) regardless of T (matrix.h:484, matrix_gpu.h:489), which makes
    // its double fixtures unreadable through read(); that defect is not reproduced.
    bool write(std::string const& filepath) const {
        std::ofstream f(filepath, std::ios::out | std::ios::binary | std::ios::trunc);
        if (!f) { std::cout << "File does not exist" << std::endl; return false; }
        f.write(reinterpret_cast<const char*>(data_.data()), (std::streamsize)(data_.size() * sizeof(T)));
        return (bool)f;
    }
    bool read(std::string const& filepath) {
        std::ifstream f(filepath, std::ios::in | std::ios::binary);
        if (!f) { std::cout << "File does not exist" << std::endl; return false; }
        f.read(reinterpret_cast<char*>(data_.data()), (std::streamsize)(data_.size() * sizeof(T)));
        return f.gcount() == (std::streamsize)(data_.size() * sizeof(T));
    }
    void print(const uint32_t& truc = 16u) const {                // matrix.h:493
        std::cout << std::fixed << std::setprecision(6);
        std::cout << "\n-------\nMatrix [" << size() << " elements; m = " << nrows << ", n = " << ncols << "]" << std::endl;
        std::cout << "Size of Payload: " << sizeof(T) * size() << 'b' << std::endl << std::endl;
        for (size_t i = 0; i <= truc && i < nrows; ++i) {
            if (i == truc) { std::cout << " ... " << std::endl; i = nrows - 1u; }
            for (size_t j = 0; j <= truc && j < ncols; ++j) {
                if (j == truc) { std::cout << "... "; j = ncols - 1u; }
                std::cout << ' ' << data_[i * ncols + j] << ' ';
            }
            std::cout << std::endl;
        }
    }
};

namespace gpu {
template <typename T> using Matrix = ::csc586::Matrix<T>;   // matrix_gpu.h:79: same class, no OpenMP flag
using ::csc586::Slice;
template <typename T> struct Reflection { Matrix<T> w, w_T; T tau; };   // matrix_gpu.h:538
}  // namespace gpu

namespace serial {
template <typename T>
struct Bidiagonal {          // svd_serial.h:80
    std::vector<T> d, e;
    Bidiagonal<T> slice(const size_t d0, const size_t d1, const size_t e0, const size_t e1) const {
        Bidiagonal<T> t;
        t.d.assign(d.begin() + d0, d.begin() + d1 + 1);
        t.e.assign(e.begin() + e0, e.begin() + e1 + 1);
        return t;
    }
};
}  // namespace serial

// ---------------------------------------------------------------------------------------------------
// C-ABI plumbing: one lazily created handle per (dtype) and host thread, grown on demand.
// ---------------------------------------------------------------------------------------------------
namespace b200 {
template <typename T> constexpr int dtype_code() { return std::is_same<T, float>::value ? SVDB200_F32 : SVDB200_F64; }

struct Session {
    svdb200_handle h = nullptr;
    size_t max_n = 0, band = 0;
    int dtype = -1, device = 0;
    ~Session() { if (h) svdb200_destroy(h); }
    svdb200_handle get(size_t n, size_t b, int dt) {
        if (!h || n > max_n || b > band || dt != dtype) {
            if (h) svdb200_destroy(h);
            h = nullptr;
            int st = svdb200_create(&h, device, std::max(n, max_n), std::max(b, band), dt);
            if (st != 0) {
                std::fprintf(stderr, "svdb200_create failed: %s\n", svdb200_strerror(st));
                assert(false && "svdb200: no usable CUDA device (there is no CPU fallback)");
                std::abort();
            }
            max_n = std::max(n, max_n); band = std::max(b, band); dtype = dt;
        }
        return h;
    }
};
template <typename T> inline Session& session() { static thread_local Session s; return s; }

inline void check(int st, const char* what) {
    if (st != 0) {
        std::fprintf(stderr, "%s: svdb200 status %d (%s)\n", what, st, svdb200_strerror(st));
        assert(false && "svdb200 call failed");      // the reference assert()s on shape errors
        std::abort();
    }
}
template <typename T> int dense_to_band(svdb200_handle h, T* a, size_t m, size_t n, size_t b, int order);
template <> inline int dense_to_band<float>(svdb200_handle h, float* a, size_t m, size_t n, size_t b, int o) { return svdb200_dense_to_band_f32(h, a, m, n, b, o); }
template <> inline int dense_to_band<double>(svdb200_handle h, double* a, size_t m, size_t n, size_t b, int o) { return svdb200_dense_to_band_f64(h, a, m, n, b, o); }
template <typename T> int band_to_bidiag(svdb200_handle h, T* a, size_t m, size_t n, size_t b, T* d, T* e);
template <> inline int band_to_bidiag<float>(svdb200_handle h, float* a, size_t m, size_t n, size_t b, float* d, float* e) { return svdb200_band_to_bidiag_f32(h, a, m, n, b, d, e); }
template <> inline int band_to_bidiag<double>(svdb200_handle h, double* a, size_t m, size_t n, size_t b, double* d, double* e) { return svdb200_band_to_bidiag_f64(h, a, m, n, b, d, e); }
template <typename T> int bidiagonalize_many(svdb200_handle h, size_t count, T* const* a, const size_t* n, size_t b, int order, T* const* d, T* const* e);
template <> inline int bidiagonalize_many<float>(svdb200_handle h, size_t c, float* const* a, const size_t* n, size_t b, int o, float* const* d, float* const* e) { return svdb200_bidiagonalize_many_f32(h, c, a, n, b, o, d, e); }
template <> inline int bidiagonalize_many<double>(svdb200_handle h, size_t c, double* const* a, const size_t* n, size_t b, int o, double* const* d, double* const* e) { return svdb200_bidiagonalize_many_f64(h, c, a, n, b, o, d, e); }
template <typename T> int bidiag_qr(svdb200_handle h, const T* d, const T* e, size_t n, T* s, long long* sw);
template <> inline int bidiag_qr<float>(svdb200_handle h, const float* d, const float* e, size_t n, float* s, long long* sw) { return svdb200_bidiag_qr_f32(h, d, e, n, s, sw); }
template <> inline int bidiag_qr<double>(svdb200_handle h, const double* d, const double* e, size_t n, double* s, long long* sw) { return svdb200_bidiag_qr_f64(h, d, e, n, s, sw); }
}  // namespace b200

namespace gpu {
// Panel-order dense -> band on the GPU; mutates A in place AND returns it by value, like
// cuda_brd_p1 (svd_cuda_1.cu:750, svd_cuda_2.cu:1117).  The reference is float-only; the
// template also serves double.
template <typename T>
Matrix<T> cuda_brd_p1(Matrix<T>& A, size_t const b_size) {
    assert(A.nrows == A.ncols && b_size > 0 && A.nrows % b_size == 0 && "square matrix with band | n expected");
    auto h = b200::session<T>().get(A.nrows, b_size, b200::dtype_code<T>());
    b200::check(b200::dense_to_band<T>(h, A.data(), A.nrows, A.ncols, b_size, SVDB200_ORDER_PANEL), "cuda_brd_p1");
    return A;
}
// gpu::brd_p2(A, w) takes the internal width w = band + 1 (svd_cpu.h:631); returns the matrix.
template <typename T>
Matrix<T> brd_p2(Matrix<T>& A, size_t const w) {
    assert(w >= 2 && "gpu::brd_p2 takes band + 1");
    auto h = b200::session<T>().get(A.nrows, w - 1, b200::dtype_code<T>());
    b200::check(b200::band_to_bidiag<T>(h, A.data(), A.nrows, A.ncols, w - 1, nullptr, nullptr), "gpu::brd_p2");
    return A;
}
// A list of independent instances, dense -> band -> bidiagonal, handed over together: what the reference's benchmark
// loop over test instances (timing.h:55-91, svd_cuda_2.cu:1370-1384) becomes when the device may overlap stage 2 of
// one instance with stage 1 of the next.  Every matrix is overwritten like cuda_brd_p1 + brd_p2 would; the
// results equal one call per instance.
template <typename T>
std::vector<serial::Bidiagonal<T>> cuda_bidiagonalize_many(std::vector<Matrix<T>>& As, size_t const b_size) {
    std::vector<serial::Bidiagonal<T>> out(As.size());
    if (As.empty()) return out;
    std::vector<T*> a(As.size()), d(As.size()), e(As.size());
    std::vector<size_t> n(As.size());
    size_t max_n = 0;
    for (size_t i = 0; i < As.size(); ++i) {
        assert(As[i].nrows == As[i].ncols && b_size > 0 && As[i].nrows % b_size == 0 && "square matrices with band | n expected");
        n[i] = As[i].nrows;
        max_n = std::max(max_n, n[i]);
        out[i].d.resize(n[i]);
        out[i].e.resize(n[i] > 0 ? n[i] - 1 : 0);
        a[i] = As[i].data(); d[i] = out[i].d.data(); e[i] = out[i].e.data();
    }
    auto h = b200::session<T>().get(max_n, b_size, b200::dtype_code<T>());
    b200::check(b200::bidiagonalize_many<T>(h, As.size(), a.data(), n.data(), b_size, SVDB200_ORDER_PANEL, d.data(), e.data()), "cuda_bidiagonalize_many");
    return out;
}
}  // namespace gpu

namespace parallel {
// Tile-order (flat-tree) dense -> band: same task order, arithmetic order and therefore the same
// bits as the reference's OpenMP path (svd_parallel.h:411-533).
template <typename T>
Matrix<T> brd_p1(Matrix<T>& A, size_t const t_size) {
    assert(A.nrows == A.ncols && t_size > 0 && A.nrows % t_size == 0 && "square matrix with tile | n expected");
    auto h = b200::session<T>().get(A.nrows, t_size, b200::dtype_code<T>());
    b200::check(b200::dense_to_band<T>(h, A.data(), A.nrows, A.ncols, t_size, SVDB200_ORDER_TILE), "parallel::brd_p1");
    return A;
}
// band -> bidiagonal; leaves the full matrix updated in A and returns {diag(A), diag(A,1)} (svd_parallel.h:640-695)
template <typename T>
serial::Bidiagonal<T> brd_p2(Matrix<T>& A, size_t b_size = 0u) {
    assert(b_size > 0 && "band size required");
    auto h = b200::session<T>().get(A.nrows, b_size, b200::dtype_code<T>());
    serial::Bidiagonal<T> B;
    B.d.resize(A.ncols);
    B.e.resize(A.ncols > 0 ? A.ncols - 1 : 0);
    b200::check(b200::band_to_bidiag<T>(h, A.data(), A.nrows, A.ncols, b_size, B.d.data(), B.e.data()), "parallel::brd_p2");
    return B;
}
}  // namespace parallel

namespace serial {
// zero-shift QR diagonalisation; returns |sigma| sorted descending in .d (svd_serial.h:368-422).
// On non-convergence the reference prints an error and returns; so does this.
template <typename T>
Bidiagonal<T> qrd(Bidiagonal<T>& B) {
    auto h = b200::session<T>().get(B.d.size(), 1, b200::dtype_code<T>());
    std::vector<T> sigma(B.d.size());
    int st = b200::bidiag_qr<T>(h, B.d.data(), B.e.data(), B.d.size(), sigma.data(), nullptr);
    if (st == SVDB200_E_NOCONV) std::cout << "Error: Maximum iterations reached without convergence." << std::endl;
    else b200::check(st, "serial::qrd");
    B.d = sigma;
    return B;
}
}  // namespace serial

namespace benchmark {
using duration = float;
// mean wall-clock microseconds of f(copy_of_instance, b_size); the copy is excluded (timing.h:55-91)
template <typename Callable, typename Container>
duration benchmark(Callable f, Container test_instances, size_t const b_size) {
    auto elapsed = std::chrono::steady_clock::duration::zero();
    for (const auto& inst : test_instances) {
        auto x = inst;
        auto const t0 = std::chrono::steady_clock::now();
        auto out = f(x, b_size);
        auto const t1 = std::chrono::steady_clock::now();
        (void)out;
        elapsed += t1 - t0;
    }
    return std::chrono::duration_cast<std::chrono::microseconds>(elapsed).count() / static_cast<duration>(test_instances.size());
}
// the same mean for a callable that takes the whole list of instances at once (gpu::cuda_bidiagonalize_many)
template <typename Callable, typename Container>
duration benchmark_many(Callable f, Container test_instances, size_t const b_size) {
    auto x = test_instances;
    auto const t0 = std::chrono::steady_clock::now();
    auto out = f(x, b_size);
    auto const t1 = std::chrono::steady_clock::now();
    (void)out;
    return std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() / static_cast<duration>(test_instances.size());
}
}  // namespace benchmark

}  // namespace csc586
#endif  // SVDB200_MATRIX_HPP
