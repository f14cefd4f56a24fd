// Stage 2: band -> bidiagonal by bulge chasing.  Replaces csc586::parallel::brd_p2<T>
// (svd_parallel.h:640-695; kernels band_rd_top 569-589, band_rd_right 600-608, band_rd_left 617-624).
//
// B200 design
//   * one CTA per in-flight sweep; sweeps are pipelined across SMs: pair p of sweep i+1 may start
//     once sweep i has completed pair p+2 (window-overlap analysis, DESIGN.md / SURVEY 8a''),
//     enforced with one per-sweep progress counter in global memory (st.release / ld.acquire);
//   * the band region (n*(3b) elements) is L2-resident; windows are staged in shared memory and all
//     inter-CTA data goes through L2 (ld.global.cg), never the non-coherent L1;
//   * arithmetic is BIT-FAITHFUL to the reference: explicit H = I - tau w w^T, window*H / H*window
//     with k-ascending sums from 0, separate (never fused) multiply and add, reflector scalars in
//     double -- so the kernel reproduces data/bidiagonal_* exactly when fed data/band_* (SURVEY 0.7).
//   * the reference's window schedule is reproduced including its boundary behaviour.
#include <climits>
#include "common.cuh"

namespace svdb200 {

namespace {

constexpr int kEptMax = 8;   // outputs per thread (2*c*c / blockDim) upper bound

// One window operation. kind 0: A_t <- A_t * H(first row); kind 1: A_t <- H(first column) * A_t.
template <typename T>
__device__ __forceinline__ void window_op(T* __restrict__ A, size_t n, int kind, int i1, int i2, int j1, int j2,
                                          T* Win, T* H, T* wv, T* sc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nr = i2 - i1, nc = j2 - j1;
    const int tot = nr * nc;
    for (int e = tid; e < tot; e += nt) {
        int r = e / nc, cc = e - r * nc;
        Win[e] = ld_cg(&A[(size_t)(i1 + r) * n + (j1 + cc)]);
    }
    __syncthreads();
    const int L = kind == 0 ? nc : nr;        // reflector length
    const int xs = kind == 0 ? 1 : nc;        // stride of x inside Win
    if (tid == 0) {
        T acc = (T)0;
        for (int i = 0; i < L; ++i) {          // matrix.h:59-62: index-order, unfused
            T x = Win[i * xs];
            acc = RN<T>::add(acc, RN<T>::mul(x, x));
        }
        T alpha, tau;
        householder_scalars<T>(Win[0], RN<T>::sqrt(acc), alpha, tau);
        sc[0] = alpha;
        sc[1] = tau;
    }
    __syncthreads();
    const T alpha = sc[0];
    const T mtau = -sc[1];
    for (int i = tid; i < L; i += nt) wv[i] = (i == 0) ? (T)1 : RN<T>::mul(Win[i * xs], alpha);
    __syncthreads();
    for (int e = tid; e < L * L; e += nt) {    // svd_serial.h:204-211
        int i = e / L, j = e - i * L;
        T h = RN<T>::mul(RN<T>::add((T)0, RN<T>::mul(wv[i], wv[j])), mtau);
        if (i == j) h = RN<T>::add((T)1, h);
        H[e] = h;
    }
    __syncthreads();
    T acc[kEptMax];
    int rr[kEptMax], cc_[kEptMax];
#pragma unroll
    for (int q = 0; q < kEptMax; ++q) {
        int e = tid + q * nt;
        acc[q] = (T)0;
        rr[q] = (e < tot) ? e / nc : 0;
        cc_[q] = (e < tot) ? e - rr[q] * nc : 0;
    }
    if (kind == 0) {
        for (int k = 0; k < L; ++k) {
#pragma unroll
            for (int q = 0; q < kEptMax; ++q)
                acc[q] = RN<T>::add(acc[q], RN<T>::mul(Win[rr[q] * nc + k], H[k * L + cc_[q]]));
        }
    } else {
        for (int k = 0; k < L; ++k) {
#pragma unroll
            for (int q = 0; q < kEptMax; ++q)
                acc[q] = RN<T>::add(acc[q], RN<T>::mul(H[rr[q] * L + k], Win[k * nc + cc_[q]]));
        }
    }
#pragma unroll
    for (int q = 0; q < kEptMax; ++q) {
        int e = tid + q * nt;
        if (e < tot) st_cg(&A[(size_t)(i1 + rr[q]) * n + (j1 + cc_[q])], acc[q]);
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) stage2_chase_kernel(T* __restrict__ A, int n, int band, int* __restrict__ prog) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Win = reinterpret_cast<T*>(smem_raw);
    const int c = band, w = band + 1;
    T* H = Win + 2 * c * c + c;
    T* wv = H + c * c;
    T* sc = wv + c;
    const int tid = threadIdx.x;
    for (int i = blockIdx.x; i < n - 1; i += gridDim.x) {
        const int top_j2 = min(i + 2 * w - 1, n);
        const int npairs = 1 + (n - top_j2) / c + 1;
        for (int p = 0; p < npairs; ++p) {
            if (i > 0) {
                if (tid == 0) {
                    while (ld_acquire(&prog[i - 1]) < p + 3) __nanosleep(40);
                }
                __syncthreads();
            }
            if (p == 0) {
                window_op<T>(A, (size_t)n, 0, i, min(i + w, n), i + 1, min(i + w, n), Win, H, wv, sc);
                window_op<T>(A, (size_t)n, 1, i + 1, min(i + w, n), i + 1, top_j2, Win, H, wv, sc);
            } else {
                const int k = p - 1;
                const int r0 = min(i + 1 + k * c, n), r1 = min(i + 1 + (k + 1) * c, n);
                const int r2 = min(i + 1 + (k + 2) * c, n), c3 = min(i + 1 + (k + 3) * c, n);
                if (r2 > r1) window_op<T>(A, (size_t)n, 0, r0, r2, r1, r2, Win, H, wv, sc);
                if (c3 > r1) window_op<T>(A, (size_t)n, 1, r1, r2, r1, c3, Win, H, wv, sc);
            }
            if (tid == 0) {
                __threadfence();
                st_release(&prog[i], p + 1 == npairs ? INT_MAX : p + 1);
            }
        }
    }
}

template <typename T>
__global__ void extract_bidiagonal_kernel(const T* __restrict__ A, int n, T* __restrict__ d, T* __restrict__ e) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (d) d[i] = A[(size_t)i * n + i];
        if (e && i + 1 < n) e[i] = A[(size_t)i * n + i + 1];
    }
}

}  // namespace

template <typename T>
int stage2_chase(Ctx* c, T* a, size_t n, size_t band, T* d, T* e) {
    if (n < 2 || band < 1) return SVDB200_E_SHAPE;
    if (n > (size_t)INT_MAX / 4) return SVDB200_E_CAPACITY;
    const int cb = (int)band;
    ProfScope ps(c, 4, 4.0 * (double)band * (double)n * (double)n * sizeof(T));
    size_t smem = (size_t)(3 * cb * cb + 2 * cb + 8) * sizeof(T);
    int want = 2 * cb * cb;
    int nt = ((want + 31) / 32) * 32;
    if (nt < 32) nt = 32;
    if (nt > 1024) nt = 1024;
    if ((2 * cb * cb + nt - 1) / nt > kEptMax) return SVDB200_E_CAPACITY;
    if (smem > 227 * 1024) return SVDB200_E_CAPACITY;
    auto kern = stage2_chase_kernel<T>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SVDB_CHECK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
    if (per_sm < 1) return SVDB200_E_CAPACITY;
    // No more sweeps can be in flight than the pipeline admits (one every 3 pairs, SURVEY 8a'').
    long long inflight = (long long)(n / band) / 3 + 2;
    long long grid = (long long)per_sm * c->num_sms;
    if (grid > inflight) grid = inflight;
    if (grid > (long long)n - 1) grid = (long long)n - 1;
    if (grid < 1) grid = 1;
    SVDB_CHECK(c, cudaMemsetAsync(c->prog, 0, sizeof(int) * n, c->stream));
    int ni = (int)n, bi = cb;
    int* prog = c->prog;
    void* args[] = {&a, &ni, &bi, &prog};
    SVDB_CHECK(c, cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)grid), dim3(nt), args, smem, c->stream));
    c->launches++;
    if (d || e) {
        extract_bidiagonal_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a, ni, d, e);
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

template int stage2_chase<float>(Ctx*, float*, size_t, size_t, float*, float*);
template int stage2_chase<double>(Ctx*, double*, size_t, size_t, double*, double*);

}  // namespace svdb200
