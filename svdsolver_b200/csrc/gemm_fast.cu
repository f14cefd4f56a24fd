// Pipelined variants of the three trailing-update GEMMs (see gemm.cu for the shapes and for the
// generic fallbacks).  Used when the band is 32 or 64 and every operand is 16-byte aligned:
//   * operands travel global -> shared with 16-byte cp.async (LDGSTS, L2-only .cg path) in a
//     multi-stage ring (3-4 stages in flight), no integer division on the copy path, zero-fill for
//     the tails through the src-size operand;
//   * FP64 on the tensor cores via mma.sync.m8n8k4 (DMMA); FP32 as 3xTF32 mma.sync.m16n8k8;
//     warp tiles of 32x32 (32x16 for band 32) cut the shared-memory fragment traffic per MMA;
//   * padded leading dimensions (ld = 4 mod 16 doubles / 8 mod 32 floats) make every fragment
//     load conflict-free; accumulator pairs are read/written as 16-byte (8-byte) vectors.
#include "common.cuh"

namespace svdb200 {
namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct Frag;

// FP64: m8n8k4.  Warp tile = (MT*8) x (NT*8); acc[mt][nt][2].
template <> struct Frag<double> {
    static constexpr int kStep = 4, kRowsPerTile = 8, kAccPerTile = 2, kEpc = 2;   // elements per 16-byte chunk
    template <int MT, int NT>
    __device__ static __forceinline__ void chunk(double (&acc)[MT][NT][2], const double* As, int sai, int sak, const double* Bs,
                                                 int ldb, int kc, int lane) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll 2
        for (int k0 = 0; k0 < kc; k0 += 4) {
            double a[MT], b[NT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) a[mt] = As[(mt * 8 + g) * sai + (k0 + q) * sak];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) b[nt] = Bs[(k0 + q) * ldb + nt * 8 + g];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1]) : "d"(a[mt]), "d"(b[nt]));
        }
    }
    // (row, col) of accumulator pair p (0: the only pair) of tile (mt, nt); the pair is 2 adjacent columns
    __device__ static __forceinline__ void pair_coord(int mt, int nt, int p, int lane, int& r, int& c) {
        r = mt * 8 + (lane >> 2);
        c = nt * 8 + 2 * (lane & 3);
        (void)p;
    }
    static constexpr int kPairs = 1;
};

// FP32: 3xTF32 m16n8k8.  Warp tile = (MT*16) x (NT*8); acc[mt][nt][4].
template <> struct Frag<float> {
    static constexpr int kStep = 8, kRowsPerTile = 16, kAccPerTile = 4, kEpc = 4;
    __device__ static __forceinline__ uint32_t tf32(float x) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
        return r;
    }
    __device__ static __forceinline__ void mma(float* c, const uint32_t* a, const uint32_t* b) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    template <int MT, int NT>
    __device__ static __forceinline__ void chunk(float (&acc)[MT][NT][4], const float* As, int sai, int sak, const float* Bs, int ldb,
                                                 int kc, int lane) {
        const int g = lane >> 2, q = lane & 3;
        for (int k0 = 0; k0 < kc; k0 += 8) {
            uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float v[4];
                v[0] = As[(mt * 16 + g) * sai + (k0 + q) * sak];
                v[1] = As[(mt * 16 + g + 8) * sai + (k0 + q) * sak];
                v[2] = As[(mt * 16 + g) * sai + (k0 + q + 4) * sak];
                v[3] = As[(mt * 16 + g + 8) * sai + (k0 + q + 4) * sak];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ah[mt][i] = tf32(v[i]);
                    al[mt][i] = tf32(v[i] - __uint_as_float(ah[mt][i]));
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float v0 = Bs[(k0 + q) * ldb + nt * 8 + g], v1 = Bs[(k0 + q + 4) * ldb + nt * 8 + g];
                bh[nt][0] = tf32(v0); bl[nt][0] = tf32(v0 - __uint_as_float(bh[nt][0]));
                bh[nt][1] = tf32(v1); bl[nt][1] = tf32(v1 - __uint_as_float(bh[nt][1]));
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma(acc[mt][nt], al[mt], bh[nt]);   // small terms first
                    mma(acc[mt][nt], ah[mt], bl[nt]);
                    mma(acc[mt][nt], ah[mt], bh[nt]);
                }
        }
    }
    __device__ static __forceinline__ void pair_coord(int mt, int nt, int p, int lane, int& r, int& c) {
        r = mt * 16 + (lane >> 2) + (p ? 8 : 0);
        c = nt * 8 + 2 * (lane & 3);
    }
    static constexpr int kPairs = 2;
};

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename T> __host__ __device__ constexpr int pad_mn(int cols) { return sizeof(T) == 8 ? cols + 4 : cols + 8; }
template <typename T> __host__ __device__ constexpr int pad_k(int cols) { return cols + 4; }

// Copy ROWS x COLS (COLS a multiple of the 16-byte chunk) from global (row stride ldg elements) to
// shared (row stride lds) with cp.async; rows >= row_lim and columns >= col_lim are zero-filled.
template <typename T, int ROWS, int COLS, int NTHREADS>
__device__ __forceinline__ void async_tile(T* dst, int lds, const T* src, size_t ldg, int row_lim, int col_lim) {
    constexpr int EPC = Frag<T>::kEpc, CPR = COLS / EPC, TOTAL = ROWS * CPR;
    static_assert(COLS % EPC == 0, "tile width must be a multiple of the 16-byte chunk");
#pragma unroll
    for (int u = 0; u < (TOTAL + NTHREADS - 1) / NTHREADS; ++u) {
        const int idx = threadIdx.x + u * NTHREADS;
        if (TOTAL % NTHREADS != 0 && idx >= TOTAL) break;
        const int r = idx / CPR, cc = (idx % CPR) * EPC;
        const bool ok = r < row_lim && cc < col_lim;
        cp_async16(dst + r * lds + cc, ok ? src + (size_t)r * ldg + cc : src, ok ? 16 : 0);
    }
}

// ---- Wpart[split](B x N) = V(rows x B)^T * C(rows x N) ------------------------------------------------
template <typename T, int B>
__global__ void __launch_bounds__(256, 2)
gemm_tn_fast_kernel(const T* __restrict__ V, const T* __restrict__ C, size_t ldc, int M, int N, T* __restrict__ Wpart, int rows_per_split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BN = 128, KC = 16, NS = 4, WM = B / 32, WN = 8 / WM, WCOLS = BN / WN;
    constexpr int MT = 32 / Frag<T>::kRowsPerTile, NT = WCOLS / 8;
    constexpr int LDV = pad_mn<T>(B), LDC = pad_mn<T>(BN), STAGE = KC * (LDV + LDC);
    T* sm = reinterpret_cast<T*>(smem_raw);
    const int n0 = blockIdx.x * BN, split = blockIdx.y;
    const int r_begin = split * rows_per_split, r_end = min(M, r_begin + rows_per_split);
    const int niter = (r_end - r_begin + KC - 1) / KC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp / WN, wn = warp % WN;
    T acc[MT][NT][Frag<T>::kAccPerTile];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NT; ++b2)
#pragma unroll
            for (int i = 0; i < Frag<T>::kAccPerTile; ++i) acc[a][b2][i] = (T)0;
    auto load_stage = [&](int it) {
        T* vs = sm + (it % NS) * STAGE;
        T* cs = vs + KC * LDV;
        const int k0 = r_begin + it * KC;
        async_tile<T, KC, B, 256>(vs, LDV, V + (size_t)k0 * B, (size_t)B, r_end - k0, B);
        async_tile<T, KC, BN, 256>(cs, LDC, C + (size_t)k0 * ldc + n0, ldc, r_end - k0, N - n0);
    };
#pragma unroll
    for (int s = 0; s < NS - 1; ++s) {
        if (s < niter) load_stage(s);
        cp_async_commit();
    }
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<NS - 2>();
        __syncthreads();
        if (it + NS - 1 < niter) load_stage(it + NS - 1);
        cp_async_commit();
        const T* vs = sm + (it % NS) * STAGE;
        const T* cs = vs + KC * LDV;
        Frag<T>::template chunk<MT, NT>(acc, vs + wm * 32, 1, LDV, cs + wn * WCOLS, LDC, KC, lane);
    }
    T* out = Wpart + (size_t)split * B * N;
    using V2 = typename Vec2<T>::type;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NT; ++b2)
#pragma unroll
            for (int p = 0; p < Frag<T>::kPairs; ++p) {
                int r, cc;
                Frag<T>::pair_coord(a, b2, p, lane, r, cc);
                const int gi = wm * 32 + r, gc = n0 + wn * WCOLS + cc;
                if (gc < N) {
                    V2 v;
                    v.x = acc[a][b2][2 * p];
                    v.y = acc[a][b2][2 * p + 1];
                    *reinterpret_cast<V2*>(out + (size_t)gi * N + gc) = v;
                }
            }
}

// ---- Wpart[split](M x B) = C(M x cols) * Ut(cols x B) ----------------------------------------------------
template <typename T, int B>
__global__ void __launch_bounds__(256, 1)
gemm_nn_fast_kernel(const T* __restrict__ C, size_t ldc, int M, int N, const T* __restrict__ Ut, T* __restrict__ Wpart, int cols_per_split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BM = 128, KC = 32, NS = 3, WCOLS = B / 2;
    constexpr int MT = 32 / Frag<T>::kRowsPerTile, NT = WCOLS / 8;
    constexpr int LDK = pad_k<T>(KC), LDU = pad_mn<T>(B), STAGE = BM * LDK + KC * LDU;
    T* sm = reinterpret_cast<T*>(smem_raw);
    const int m0 = blockIdx.x * BM, split = blockIdx.y;
    const int c_begin = split * cols_per_split, c_end = min(N, c_begin + cols_per_split);
    const int niter = (c_end - c_begin + KC - 1) / KC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1;
    T acc[MT][NT][Frag<T>::kAccPerTile];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NT; ++b2)
#pragma unroll
            for (int i = 0; i < Frag<T>::kAccPerTile; ++i) acc[a][b2][i] = (T)0;
    auto load_stage = [&](int it) {
        T* cs = sm + (it % NS) * STAGE;
        T* us = cs + BM * LDK;
        const int k0 = c_begin + it * KC;
        async_tile<T, BM, KC, 256>(cs, LDK, C + (size_t)m0 * ldc + k0, ldc, M - m0, c_end - k0);
        async_tile<T, KC, B, 256>(us, LDU, Ut + (size_t)k0 * B, (size_t)B, c_end - k0, B);
    };
#pragma unroll
    for (int s = 0; s < NS - 1; ++s) {
        if (s < niter) load_stage(s);
        cp_async_commit();
    }
    for (int it = 0; it < niter; ++it) {
        cp_async_wait<NS - 2>();
        __syncthreads();
        if (it + NS - 1 < niter) load_stage(it + NS - 1);
        cp_async_commit();
        const T* cs = sm + (it % NS) * STAGE;
        const T* us = cs + BM * LDK;
        Frag<T>::template chunk<MT, NT>(acc, cs + wm * 32 * LDK, LDK, 1, us + wn * WCOLS, LDU, KC, lane);
    }
    T* out = Wpart + (size_t)split * M * B;
    using V2 = typename Vec2<T>::type;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NT; ++b2)
#pragma unroll
            for (int p = 0; p < Frag<T>::kPairs; ++p) {
                int r, cc;
                Frag<T>::pair_coord(a, b2, p, lane, r, cc);
                const int gr = m0 + wm * 32 + r, gj = wn * WCOLS + cc;
                if (gr < M) {
                    V2 v;
                    v.x = acc[a][b2][2 * p];
                    v.y = acc[a][b2][2 * p + 1];
                    *reinterpret_cast<V2*>(out + (size_t)gr * B + gj) = v;
                }
            }
}

// ---- C(M x N) += P(M x KB) * Q(KB x N) --------------------------------------------------------------------
template <typename T, int KB>
__global__ void __launch_bounds__(256, 2)
rank_update_fast_kernel(T* __restrict__ C, size_t ldc, int M, int N, const T* __restrict__ P, const T* __restrict__ Q, size_t ldq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BM = 128, BN = 64, WCOLS = 32;
    constexpr int MT = 32 / Frag<T>::kRowsPerTile, NT = WCOLS / 8;
    constexpr int LDP = pad_k<T>(KB), LDQ = pad_mn<T>(BN);
    T* ps = reinterpret_cast<T*>(smem_raw);
    T* qs = ps + BM * LDP;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    async_tile<T, BM, KB, 256>(ps, LDP, P + (size_t)m0 * KB, (size_t)KB, M - m0, KB);
    async_tile<T, KB, BN, 256>(qs, LDQ, Q + n0, ldq, KB, N - n0);
    cp_async_commit();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp >> 1, wn = warp & 1;
    T acc[MT][NT][Frag<T>::kAccPerTile];
    using V2 = typename Vec2<T>::type;
    // accumulators start from C: these loads overlap the asynchronous tile fill
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NT; ++b2)
#pragma unroll
            for (int p = 0; p < Frag<T>::kPairs; ++p) {
                int r, cc;
                Frag<T>::pair_coord(a, b2, p, lane, r, cc);
                const int gr = m0 + wm * 32 + r, gc = n0 + wn * WCOLS + cc;
                V2 v;
                v.x = (T)0; v.y = (T)0;
                if (gr < M && gc < N) v = *reinterpret_cast<const V2*>(C + (size_t)gr * ldc + gc);
                acc[a][b2][2 * p] = v.x;
                acc[a][b2][2 * p + 1] = v.y;
            }
    cp_async_wait<0>();
    __syncthreads();
    Frag<T>::template chunk<MT, NT>(acc, ps + wm * 32 * LDP, LDP, 1, qs + wn * WCOLS, LDQ, KB, lane);
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b2 = 0; b2 < NT; ++b2)
#pragma unroll
            for (int p = 0; p < Frag<T>::kPairs; ++p) {
                int r, cc;
                Frag<T>::pair_coord(a, b2, p, lane, r, cc);
                const int gr = m0 + wm * 32 + r, gc = n0 + wn * WCOLS + cc;
                if (gr < M && gc < N) {
                    V2 v;
                    v.x = acc[a][b2][2 * p];
                    v.y = acc[a][b2][2 * p + 1];
                    *reinterpret_cast<V2*>(C + (size_t)gr * ldc + gc) = v;
                }
            }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// Each *_fast returns 0 when it ran, 1 when its preconditions do not hold (caller falls back), or an
// error status.
template <typename T>
int rank_update_fast(Ctx* c, T* cm, size_t ldc, int M, int N, int K, const T* p, const T* q, size_t ldq) {
    constexpr int EPC = 16 / sizeof(T);
    if ((K != 32 && K != 64) || N % 2 != 0 || N % EPC != 0) return 1;
    if (!aligned16(cm) || !aligned16(p) || !aligned16(q) || (ldc * sizeof(T)) % 16 || (ldq * sizeof(T)) % 16) return 1;
    dim3 grid((N + 63) / 64, (M + 127) / 128);
#define SVDB_RU(KBv)                                                                                               \
    {                                                                                                              \
        size_t smem = ((size_t)128 * pad_k<T>(KBv) + (size_t)KBv * pad_mn<T>(64)) * sizeof(T);                     \
        auto kern = rank_update_fast_kernel<T, KBv>;                                                               \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kern<<<grid, 256, smem, c->stream>>>(cm, ldc, M, N, p, q, ldq);                                            \
    }
    if (K == 32) SVDB_RU(32) else SVDB_RU(64)
#undef SVDB_RU
    SVDB_CHECK(c, cudaGetLastError());
    c->launches++;
    return 0;
}

template <typename T>
int gemm_tn_fast(Ctx* c, const T* v, const T* cm, size_t ldc, int M, int N, int B, T* out, int splits, int rows_per_split) {
    constexpr int EPC = 16 / sizeof(T);
    if ((B != 32 && B != 64) || N % EPC != 0 || rows_per_split % 16 != 0) return 1;
    if (!aligned16(cm) || !aligned16(v) || !aligned16(out) || (ldc * sizeof(T)) % 16) return 1;
    dim3 grid((N + 127) / 128, splits);
#define SVDB_TN(Bv)                                                                                                \
    {                                                                                                              \
        size_t smem = (size_t)4 * 16 * (pad_mn<T>(Bv) + pad_mn<T>(128)) * sizeof(T);                               \
        auto kern = gemm_tn_fast_kernel<T, Bv>;                                                                    \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kern<<<grid, 256, smem, c->stream>>>(v, cm, ldc, M, N, out, rows_per_split);                               \
    }
    if (B == 32) SVDB_TN(32) else SVDB_TN(64)
#undef SVDB_TN
    SVDB_CHECK(c, cudaGetLastError());
    c->launches++;
    return 0;
}

template <typename T>
int gemm_nn_fast(Ctx* c, const T* cm, size_t ldc, int M, int N, int B, const T* ut, T* out, int splits, int cols_per_split) {
    constexpr int EPC = 16 / sizeof(T);
    if ((B != 32 && B != 64) || N % EPC != 0 || cols_per_split % 32 != 0) return 1;
    if (!aligned16(cm) || !aligned16(ut) || !aligned16(out) || (ldc * sizeof(T)) % 16) return 1;
    dim3 grid((M + 127) / 128, splits);
#define SVDB_NN(Bv)                                                                                                \
    {                                                                                                              \
        size_t smem = (size_t)3 * (128 * pad_k<T>(32) + 32 * pad_mn<T>(Bv)) * sizeof(T);                           \
        auto kern = gemm_nn_fast_kernel<T, Bv>;                                                                    \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kern<<<grid, 256, smem, c->stream>>>(cm, ldc, M, N, ut, out, cols_per_split);                              \
    }
    if (B == 32) SVDB_NN(32) else SVDB_NN(64)
#undef SVDB_NN
    SVDB_CHECK(c, cudaGetLastError());
    c->launches++;
    return 0;
}

template int rank_update_fast<float>(Ctx*, float*, size_t, int, int, int, const float*, const float*, size_t);
template int rank_update_fast<double>(Ctx*, double*, size_t, int, int, int, const double*, const double*, size_t);
template int gemm_tn_fast<float>(Ctx*, const float*, const float*, size_t, int, int, int, float*, int, int);
template int gemm_tn_fast<double>(Ctx*, const double*, const double*, size_t, int, int, int, double*, int, int);
template int gemm_nn_fast<float>(Ctx*, const float*, size_t, int, int, int, const float*, float*, int, int);
template int gemm_nn_fast<double>(Ctx*, const double*, size_t, int, int, int, const double*, double*, int, int);

}  // namespace svdb200
