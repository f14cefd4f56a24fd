// Compact-WY trailing update of stage 1 -- the only dense contraction on the path.
// Replaces qr_apply / lq_apply (svd_parallel.h:243-281; GPU: svd_cuda_2.cu:1039-1110), which form
// the square Q = V(SV^T) explicitly (O(m^2 n) work).  Here the update is three GEMM shapes that
// never form Q and touch the trailing matrix twice per half-step:
//     W  = V^T C            (b x N, reduction over the M rows)            gemm_tn
//     W  = C Ut             (M x b, reduction over the N columns)         gemm_nn
//     C += P Q              (rank-b update, K = b)                         rank_update
// FP64 runs on the tensor cores through mma.sync.m8n8k4.f64 (DMMA): tcgen05.mma has no f64 kind
// (SURVEY 0.6).  FP32 runs as 3xTF32 error-compensated mma.sync.m16n8k8 (hi*hi + hi*lo + lo*hi,
// fp32 accumulate), which keeps ~fp32 accuracy through the n/b chained updates.
#include <algorithm>
#include "common.cuh"

namespace svdb200 {

// pipelined variants (gemm_fast.cu): return 0 when they ran, 1 when their preconditions do not hold
template <typename T> int rank_update_fast(Ctx*, T*, size_t, int, int, int, const T*, const T*, size_t);
template <typename T> int gemm_tn_fast(Ctx*, const T*, const T*, size_t, int, int, int, T*, int, int);
template <typename T> int gemm_nn_fast(Ctx*, const T*, size_t, int, int, int, const T*, T*, int, int);
// tcgen05 / TMEM / TMA kernels for FP32 (gemm_tc05.cu): same convention
template <typename T> int rank_update_tc05(Ctx*, T*, size_t, int, int, int, const T*, const T*, size_t);
template <typename T> int gemm_tn_tc05(Ctx*, const T*, const T*, size_t, int, int, int, T*);
template <typename T> int gemm_nn_tc05(Ctx*, const T*, size_t, int, int, int, const T*, T*);

namespace {

// ------------------------------------------------------------------------------------------------
// Warp-level MMA tile: each warp owns a 32 x 16 tile of the output.
// A(i,k) is read from smem at As[i*sai + k*sak], B(k,n) at Bs[k*ldb + n].
// ------------------------------------------------------------------------------------------------
template <typename T> struct WarpMma;

template <> struct WarpMma<double> {
    static constexpr int kStep = 4;
    static constexpr int kAcc = 16;   // 4 x 2 m8n8 tiles x 2
    __device__ static __forceinline__ void mma(double& c0, double& c1, double a, double b) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    // kc must be a multiple of 4
    __device__ static __forceinline__ void chunk(double* acc, const double* As, int sai, int sak, const double* Bs, int ldb,
                                                 int kc, int lane) {
        const int g = lane >> 2, q = lane & 3;
        for (int k0 = 0; k0 < kc; k0 += 4) {
            double a[4], b[2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = As[(mt * 8 + g) * sai + (k0 + q) * sak];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) b[nt] = Bs[(k0 + q) * ldb + nt * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma(acc[(mt * 2 + nt) * 2], acc[(mt * 2 + nt) * 2 + 1], a[mt], b[nt]);
        }
    }
    // element (row, col) of accumulator slot s (0..15) inside the 32x16 warp tile
    __device__ static __forceinline__ void coord(int s, int lane, int& r, int& c) {
        int t = s >> 1, mt = t >> 1, nt = t & 1;
        r = mt * 8 + (lane >> 2);
        c = nt * 8 + 2 * (lane & 3) + (s & 1);
    }
};

template <> struct WarpMma<float> {
    static constexpr int kStep = 8;
    static constexpr int kAcc = 16;   // 2 x 2 m16n8 tiles x 4
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
    // kc must be a multiple of 8
    __device__ static __forceinline__ void chunk(float* acc, const float* As, int sai, int sak, const float* Bs, int ldb,
                                                 int kc, int lane) {
        const int g = lane >> 2, q = lane & 3;
        for (int k0 = 0; k0 < kc; k0 += 8) {
            uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
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
            for (int nt = 0; nt < 2; ++nt) {
                float v0 = Bs[(k0 + q) * ldb + nt * 8 + g];
                float v1 = Bs[(k0 + q + 4) * ldb + nt * 8 + g];
                bh[nt][0] = tf32(v0); bl[nt][0] = tf32(v0 - __uint_as_float(bh[nt][0]));
                bh[nt][1] = tf32(v1); bl[nt][1] = tf32(v1 - __uint_as_float(bh[nt][1]));
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float* c = acc + (mt * 2 + nt) * 4;
                    mma(c, al[mt], bh[nt]);   // small terms first
                    mma(c, ah[mt], bl[nt]);
                    mma(c, ah[mt], bh[nt]);
                }
        }
    }
    __device__ static __forceinline__ void coord(int s, int lane, int& r, int& c) {
        int t = s >> 2, mt = t >> 1, nt = t & 1, i = s & 3;
        r = mt * 16 + (lane >> 2) + ((i >> 1) ? 8 : 0);
        c = nt * 8 + 2 * (lane & 3) + (i & 1);
    }
};

template <typename T> __host__ __device__ constexpr int pad_kcontig(int cols) {
    // rows are indexed by the MMA row/col index, k is contiguous
    return sizeof(T) == 8 ? ((cols + 15) / 16) * 16 + 4 : ((cols + 31) / 32) * 32 + 4;
}
template <typename T> __host__ __device__ constexpr int pad_mncontig(int cols) {
    // rows are indexed by k, the MMA row/col index is contiguous
    return sizeof(T) == 8 ? ((cols + 15) / 16) * 16 + 4 : ((cols + 31) / 32) * 32 + 8;
}

// Copy a rows x cols block (global row-major, leading dimension ldg) into smem (leading dim lds),
// zero-filling everything outside [0,row_lim) x [0,col_lim).
template <typename T>
__device__ __forceinline__ void load_tile(T* __restrict__ dst, int lds, const T* __restrict__ src, size_t ldg, int rows,
                                          int cols, int row_lim, int col_lim) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < rows * cols; e += nt) {
        int r = e / cols, c = e - r * cols;
        T v = (T)0;
        if (r < row_lim && c < col_lim) v = src[(size_t)r * ldg + c];
        dst[r * lds + c] = v;
    }
}

// ---- C(M x N) += P(M x K) * Q(K x N) --------------------------------------------------------------
template <typename T, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32)
rank_update_kernel(T* __restrict__ C, size_t ldc, int M, int N, int K, const T* __restrict__ P, const T* __restrict__ Q,
                   size_t ldq, size_t sC, size_t sP, size_t sQ) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BM = WM * 32, BN = WN * 16;
    C += (size_t)blockIdx.z * sC; P += (size_t)blockIdx.z * sP; Q += (size_t)blockIdx.z * sQ;   // batched: one matrix per z
    const int kp = ((K + WarpMma<T>::kStep - 1) / WarpMma<T>::kStep) * WarpMma<T>::kStep;
    const int lda = pad_kcontig<T>(kp), ldb = pad_mncontig<T>(BN);
    T* Ps = reinterpret_cast<T*>(smem_raw);
    T* Qs = Ps + BM * lda;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    load_tile<T>(Ps, lda, P + (size_t)m0 * K, (size_t)K, BM, kp, M - m0, K);
    load_tile<T>(Qs, ldb, Q + n0, ldq, kp, BN, K, N - n0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    T acc[WarpMma<T>::kAcc];
    // accumulators start from C (issued before the barrier so the loads overlap the tile fill)
#pragma unroll
    for (int s = 0; s < WarpMma<T>::kAcc; ++s) {
        int r, c;
        WarpMma<T>::coord(s, lane, r, c);
        int gr = m0 + wm * 32 + r, gc = n0 + wn * 16 + c;
        acc[s] = (gr < M && gc < N) ? C[(size_t)gr * ldc + gc] : (T)0;
    }
    __syncthreads();
    WarpMma<T>::chunk(acc, Ps + wm * 32 * lda, lda, 1, Qs + wn * 16, ldb, kp, lane);
#pragma unroll
    for (int s = 0; s < WarpMma<T>::kAcc; ++s) {
        int r, c;
        WarpMma<T>::coord(s, lane, r, c);
        int gr = m0 + wm * 32 + r, gc = n0 + wn * 16 + c;
        if (gr < M && gc < N) C[(size_t)gr * ldc + gc] = acc[s];
    }
}

// ---- Wpart[split](b x N) = V(rows of this split x b)^T * C(rows x N) -------------------------------
template <typename T, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32)
gemm_tn_kernel(const T* __restrict__ V, const T* __restrict__ C, size_t ldc, int M, int N, int B, T* __restrict__ Wpart,
               int rows_per_split, int nsplit, size_t sV, size_t sC, size_t sW) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BMT = WM * 32, BN = WN * 16, KC = 32;
    const int lda = pad_mncontig<T>(BMT), ldb = pad_mncontig<T>(BN);
    T* Vs = reinterpret_cast<T*>(smem_raw);
    T* Cs = Vs + KC * lda;
    int split = blockIdx.y;
    if (nsplit > 0) {                                     // batched: blockIdx.y = matrix * nsplit + split
        const int bat = blockIdx.y / nsplit;
        split = blockIdx.y - bat * nsplit;
        V += (size_t)bat * sV; C += (size_t)bat * sC; Wpart += (size_t)bat * sW;
    }
    const int n0 = blockIdx.x * BN, i0 = blockIdx.z * BMT;
    const int r_begin = split * rows_per_split;
    const int r_end = min(M, r_begin + rows_per_split);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    T acc[WarpMma<T>::kAcc];
#pragma unroll
    for (int s = 0; s < WarpMma<T>::kAcc; ++s) acc[s] = (T)0;
    for (int k0 = r_begin; k0 < r_end; k0 += KC) {
        __syncthreads();
        load_tile<T>(Vs, lda, V + (size_t)k0 * B + i0, (size_t)B, KC, BMT, r_end - k0, B - i0);
        load_tile<T>(Cs, ldb, C + (size_t)k0 * ldc + n0, ldc, KC, BN, r_end - k0, N - n0);
        __syncthreads();
        WarpMma<T>::chunk(acc, Vs + wm * 32, 1, lda, Cs + wn * 16, ldb, KC, lane);
    }
    T* out = Wpart + (size_t)split * B * N;
#pragma unroll
    for (int s = 0; s < WarpMma<T>::kAcc; ++s) {
        int r, c;
        WarpMma<T>::coord(s, lane, r, c);
        int gi = i0 + wm * 32 + r, gc = n0 + wn * 16 + c;
        if (gi < B && gc < N) out[(size_t)gi * N + gc] = acc[s];
    }
}

// ---- Wpart[split](M x b) = C(M x cols of this split) * Ut(cols x b) ---------------------------------
template <typename T, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32)
gemm_nn_kernel(const T* __restrict__ C, size_t ldc, int M, int N, int B, const T* __restrict__ Ut, T* __restrict__ Wpart,
               int cols_per_split, int nsplit, size_t sC, size_t sU, size_t sW) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BM = WM * 32, BNB = WN * 16, KC = 32;
    const int lda = pad_kcontig<T>(KC), ldb = pad_mncontig<T>(BNB);
    T* Cs = reinterpret_cast<T*>(smem_raw);
    T* Us = Cs + BM * lda;
    int split = blockIdx.y;
    if (nsplit > 0) {                                     // batched: blockIdx.y = matrix * nsplit + split
        const int bat = blockIdx.y / nsplit;
        split = blockIdx.y - bat * nsplit;
        C += (size_t)bat * sC; Ut += (size_t)bat * sU; Wpart += (size_t)bat * sW;
    }
    const int m0 = blockIdx.x * BM, j0 = blockIdx.z * BNB;
    const int c_begin = split * cols_per_split;
    const int c_end = min(N, c_begin + cols_per_split);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    T acc[WarpMma<T>::kAcc];
#pragma unroll
    for (int s = 0; s < WarpMma<T>::kAcc; ++s) acc[s] = (T)0;
    for (int k0 = c_begin; k0 < c_end; k0 += KC) {
        __syncthreads();
        load_tile<T>(Cs, lda, C + (size_t)m0 * ldc + k0, ldc, BM, KC, M - m0, c_end - k0);
        load_tile<T>(Us, ldb, Ut + (size_t)k0 * B + j0, (size_t)B, KC, BNB, c_end - k0, B - j0);
        __syncthreads();
        WarpMma<T>::chunk(acc, Cs + wm * 32 * lda, lda, 1, Us + wn * 16, ldb, KC, lane);
    }
    T* out = Wpart + (size_t)split * M * B;
#pragma unroll
    for (int s = 0; s < WarpMma<T>::kAcc; ++s) {
        int r, c;
        WarpMma<T>::coord(s, lane, r, c);
        int gr = m0 + wm * 32 + r, gj = j0 + wn * 16 + c;
        if (gr < M && gj < B) out[(size_t)gr * B + gj] = acc[s];
    }
}

// W[i] = sum_{s ascending} Wpart[s][i]   (fixed order => deterministic)
template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ Wpart, T* __restrict__ W, size_t count, int splits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    T acc = Wpart[i];
    for (int s = 1; s < splits; ++s) acc += Wpart[(size_t)s * count + i];
    W[i] = acc;
}

template <typename T, int WM, int WN>
int launch_rank_update(Ctx* c, T* cm, size_t ldc, int M, int N, int K, const T* p, const T* q, size_t ldq) {
    constexpr int BM = WM * 32, BN = WN * 16;
    const int kp = ((K + WarpMma<T>::kStep - 1) / WarpMma<T>::kStep) * WarpMma<T>::kStep;
    size_t smem = ((size_t)BM * pad_kcontig<T>(kp) + (size_t)kp * pad_mncontig<T>(BN)) * sizeof(T);
    auto kern = rank_update_kernel<T, WM, WN>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    kern<<<grid, WM * WN * 32, smem, c->stream>>>(cm, ldc, M, N, K, p, q, ldq, (size_t)0, (size_t)0, (size_t)0);
    SVDB_CHECK(c, cudaGetLastError());
    c->launches++;
    return 0;
}

inline int pick_splits(long long tiles, long long kiters, int num_sms, size_t cap_elems, size_t out_elems) {
    // enough CTAs for ~2 waves, at least 4 k-iterations per split, bounded by the partial buffer
    long long want = (2LL * num_sms + tiles - 1) / tiles;
    long long maxs = kiters / 4;
    if (maxs < 1) maxs = 1;
    if (want > maxs) want = maxs;
    long long capn = (long long)(cap_elems / (out_elems ? out_elems : 1));
    if (want > capn) want = capn;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace

template <typename T>
int rank_update(Ctx* c, T* cm, size_t ldc, size_t mrows, size_t ncols, size_t b, const T* p, const T* q, size_t ldq) {
    if (mrows == 0 || ncols == 0) return 0;
    if (b == 0 || b > (size_t)kMaxBand) return SVDB200_E_CAPACITY;
    ProfScope ps(c, 3, 2.0 * (double)mrows * (double)ncols * (double)b);
    {
        int st = rank_update_tc05<T>(c, cm, ldc, (int)mrows, (int)ncols, (int)b, p, q, ldq);
        if (st != 1) return st;
    }
    {
        int st = rank_update_fast<T>(c, cm, ldc, (int)mrows, (int)ncols, (int)b, p, q, ldq);
        if (st != 1) return st;
    }
    return launch_rank_update<T, 4, 4>(c, cm, ldc, (int)mrows, (int)ncols, (int)b, p, q, ldq);
}

template <typename T>
int gemm_tn(Ctx* c, const T* v, const T* cm, size_t ldc, size_t mrows, size_t ncols, size_t b, T* w) {
    if (mrows == 0 || ncols == 0) return 0;
    if (b == 0 || b > (size_t)kMaxBand) return SVDB200_E_CAPACITY;
    const int M = (int)mrows, N = (int)ncols, B = (int)b, KC = 32;
    ProfScope ps(c, 1, 2.0 * (double)mrows * (double)ncols * (double)b);
    {
        int st = gemm_tn_tc05<T>(c, v, cm, ldc, M, N, B, w);
        if (st != 1) return st;
    }
    if (B == 32 || B == 64) {          // pipelined kernel: 128-column tiles, 16-row stages
        long long tiles_f = (N + 127) / 128;
        int splits_f = pick_splits(tiles_f, (M + KC - 1) / KC, c->num_sms, c->wpart_elems, (size_t)B * N);
        int rps = (((M + splits_f - 1) / splits_f + KC - 1) / KC) * KC;
        splits_f = (M + rps - 1) / rps;
        T* out_f = splits_f == 1 ? w : reinterpret_cast<T*>(c->wpart);
        int st = gemm_tn_fast<T>(c, v, cm, ldc, M, N, B, out_f, splits_f, rps);
        if (st == 0 && splits_f > 1) {
            size_t count = (size_t)B * N;
            reduce_partials_kernel<T><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(out_f, w, count, splits_f);
            SVDB_CHECK(c, cudaGetLastError());
            c->launches++;
        }
        if (st != 1) return st;
    }
    int wm, wn;
    if (B <= 32) { wm = 1; wn = 8; } else if (B <= 64) { wm = 2; wn = 4; } else { wm = 4; wn = 2; }
    const int BMT = wm * 32, BN = wn * 16;
    long long tiles = (long long)((N + BN - 1) / BN) * ((B + BMT - 1) / BMT);
    int splits = pick_splits(tiles, (M + KC - 1) / KC, c->num_sms, c->wpart_elems, (size_t)B * N);
    int rows_per_split = (((M + splits - 1) / splits + KC - 1) / KC) * KC;
    splits = (M + rows_per_split - 1) / rows_per_split;
    T* out = splits == 1 ? w : reinterpret_cast<T*>(c->wpart);
    dim3 grid((N + BN - 1) / BN, splits, (B + BMT - 1) / BMT);
    size_t smem = ((size_t)KC * pad_mncontig<T>(BMT) + (size_t)KC * pad_mncontig<T>(BN)) * sizeof(T);
#define SVDB_LAUNCH_TN(WMv, WNv)                                                                              \
    {                                                                                                         \
        auto kern = gemm_tn_kernel<T, WMv, WNv>;                                                              \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        kern<<<grid, 256, smem, c->stream>>>(v, cm, ldc, M, N, B, out, rows_per_split, 0, (size_t)0, (size_t)0, (size_t)0);                       \
    }
    if (wm == 1) SVDB_LAUNCH_TN(1, 8) else if (wm == 2) SVDB_LAUNCH_TN(2, 4) else SVDB_LAUNCH_TN(4, 2)
#undef SVDB_LAUNCH_TN
    SVDB_CHECK(c, cudaGetLastError());
    c->launches++;
    if (splits > 1) {
        size_t count = (size_t)B * N;
        reduce_partials_kernel<T><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(out, w, count, splits);
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

template <typename T>
int gemm_nn(Ctx* c, const T* cm, size_t ldc, size_t mrows, size_t ncols, size_t b, const T* ut, T* w) {
    if (mrows == 0 || ncols == 0) return 0;
    if (b == 0 || b > (size_t)kMaxBand) return SVDB200_E_CAPACITY;
    const int M = (int)mrows, N = (int)ncols, B = (int)b, KC = 32;
    ProfScope ps(c, 2, 2.0 * (double)mrows * (double)ncols * (double)b);
    {
        int st = gemm_nn_tc05<T>(c, cm, ldc, M, N, B, ut, w);
        if (st != 1) return st;
    }
    if (B == 32 || B == 64) {          // pipelined kernel: 128-row tiles, 32-column stages
        long long tiles_f = (M + 127) / 128;
        int splits_f = pick_splits(tiles_f, (N + KC - 1) / KC, c->num_sms, c->wpart_elems, (size_t)M * B);
        int cps = (((N + splits_f - 1) / splits_f + KC - 1) / KC) * KC;
        splits_f = (N + cps - 1) / cps;
        T* out_f = splits_f == 1 ? w : reinterpret_cast<T*>(c->wpart);
        int st = gemm_nn_fast<T>(c, cm, ldc, M, N, B, ut, out_f, splits_f, cps);
        if (st == 0 && splits_f > 1) {
            size_t count = (size_t)M * B;
            reduce_partials_kernel<T><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(out_f, w, count, splits_f);
            SVDB_CHECK(c, cudaGetLastError());
            c->launches++;
        }
        if (st != 1) return st;
    }
    int wm, wn;
    if (B <= 16) { wm = 8; wn = 1; } else if (B <= 32) { wm = 4; wn = 2; } else if (B <= 64) { wm = 2; wn = 4; } else { wm = 1; wn = 8; }
    const int BM = wm * 32, BNB = wn * 16;
    long long tiles = (long long)((M + BM - 1) / BM) * ((B + BNB - 1) / BNB);
    int splits = pick_splits(tiles, (N + KC - 1) / KC, c->num_sms, c->wpart_elems, (size_t)M * B);
    int cols_per_split = (((N + splits - 1) / splits + KC - 1) / KC) * KC;
    splits = (N + cols_per_split - 1) / cols_per_split;
    T* out = splits == 1 ? w : reinterpret_cast<T*>(c->wpart);
    dim3 grid((M + BM - 1) / BM, splits, (B + BNB - 1) / BNB);
    size_t smem = ((size_t)BM * pad_kcontig<T>(KC) + (size_t)KC * pad_mncontig<T>(BNB)) * sizeof(T);
#define SVDB_LAUNCH_NN(WMv, WNv)                                                                              \
    {                                                                                                         \
        auto kern = gemm_nn_kernel<T, WMv, WNv>;                                                              \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        kern<<<grid, 256, smem, c->stream>>>(cm, ldc, M, N, B, ut, out, cols_per_split, 0, (size_t)0, (size_t)0, (size_t)0);                      \
    }
    if (wm == 8) SVDB_LAUNCH_NN(8, 1) else if (wm == 4) SVDB_LAUNCH_NN(4, 2) else if (wm == 2) SVDB_LAUNCH_NN(2, 4) else SVDB_LAUNCH_NN(1, 8)
#undef SVDB_LAUNCH_NN
    SVDB_CHECK(c, cudaGetLastError());
    c->launches++;
    if (splits > 1) {
        size_t count = (size_t)M * B;
        reduce_partials_kernel<T><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(out, w, count, splits);
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

// ---- batched variants (uniform shapes, one matrix per grid slice): the small-matrix driver of capi.cu ----------------
template <typename T>
int rank_update_batched(Ctx* c, T* cm, size_t ldc, size_t sC, int M, int N, int K, const T* p, size_t sP, const T* q, size_t ldq,
                        size_t sQ, int count) {
    if (M <= 0 || N <= 0 || count <= 0) return 0;
    if (K <= 0 || K > kMaxBand) return SVDB200_E_CAPACITY;
    constexpr int WM = 4, WN = 4, BM = WM * 32, BN = WN * 16;
    const int kp = ((K + WarpMma<T>::kStep - 1) / WarpMma<T>::kStep) * WarpMma<T>::kStep;
    size_t smem = ((size_t)BM * pad_kcontig<T>(kp) + (size_t)kp * pad_mncontig<T>(BN)) * sizeof(T);
    auto kern = rank_update_kernel<T, WM, WN>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int z0 = 0; z0 < count; z0 += 32768) {
        const int zc = std::min(32768, count - z0);
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, zc);
        kern<<<grid, WM * WN * 32, smem, c->stream>>>(cm + (size_t)z0 * sC, ldc, M, N, K, p + (size_t)z0 * sP, q + (size_t)z0 * sQ, ldq, sC, sP, sQ);
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

template <typename T>
int gemm_tn_batched(Ctx* c, const T* v, size_t sV, const T* cm, size_t ldc, size_t sC, int M, int N, int B, T* w, size_t sW, int count) {
    if (M <= 0 || N <= 0 || count <= 0) return 0;
    if (B <= 0 || B > 64) return SVDB200_E_CAPACITY;
    constexpr int KC = 32;
    for (int z0 = 0; z0 < count; z0 += 32768) {
        const int zc = std::min(32768, count - z0);
#define SVDB_LAUNCH_TNB(WMv, WNv)                                                                              \
    {                                                                                                          \
        constexpr int BMT = WMv * 32, BN = WNv * 16;                                                           \
        size_t smem = ((size_t)KC * pad_mncontig<T>(BMT) + (size_t)KC * pad_mncontig<T>(BN)) * sizeof(T);     \
        auto kern = gemm_tn_kernel<T, WMv, WNv>;                                                               \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        dim3 grid((N + BN - 1) / BN, zc, (B + BMT - 1) / BMT);                                                 \
        kern<<<grid, 256, smem, c->stream>>>(v + (size_t)z0 * sV, cm + (size_t)z0 * sC, ldc, M, N, B, w + (size_t)z0 * sW, M, 1, sV, sC, sW); \
    }
        if (B <= 32) SVDB_LAUNCH_TNB(1, 8) else SVDB_LAUNCH_TNB(2, 4)
#undef SVDB_LAUNCH_TNB
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

template <typename T>
int gemm_nn_batched(Ctx* c, const T* cm, size_t ldc, size_t sC, int M, int N, int B, const T* ut, size_t sU, T* w, size_t sW, int count) {
    if (M <= 0 || N <= 0 || count <= 0) return 0;
    if (B <= 0 || B > 64) return SVDB200_E_CAPACITY;
    constexpr int KC = 32;
    for (int z0 = 0; z0 < count; z0 += 32768) {
        const int zc = std::min(32768, count - z0);
#define SVDB_LAUNCH_NNB(WMv, WNv)                                                                              \
    {                                                                                                          \
        constexpr int BM = WMv * 32, BNB = WNv * 16;                                                           \
        size_t smem = ((size_t)BM * pad_kcontig<T>(KC) + (size_t)KC * pad_mncontig<T>(BNB)) * sizeof(T);      \
        auto kern = gemm_nn_kernel<T, WMv, WNv>;                                                               \
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        dim3 grid((M + BM - 1) / BM, zc, (B + BNB - 1) / BNB);                                                 \
        kern<<<grid, 256, smem, c->stream>>>(cm + (size_t)z0 * sC, ldc, M, N, B, ut + (size_t)z0 * sU, w + (size_t)z0 * sW, N, 1, sC, sU, sW); \
    }
        if (B <= 16) SVDB_LAUNCH_NNB(8, 1) else if (B <= 32) SVDB_LAUNCH_NNB(4, 2) else SVDB_LAUNCH_NNB(2, 4)
#undef SVDB_LAUNCH_NNB
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

#define SVDB_INST_BATCHED(T)                                                                                                        \
    template int rank_update_batched<T>(Ctx*, T*, size_t, size_t, int, int, int, const T*, size_t, const T*, size_t, size_t, int); \
    template int gemm_tn_batched<T>(Ctx*, const T*, size_t, const T*, size_t, size_t, int, int, int, T*, size_t, int);             \
    template int gemm_nn_batched<T>(Ctx*, const T*, size_t, size_t, int, int, int, const T*, size_t, T*, size_t, int);
SVDB_INST_BATCHED(float)
SVDB_INST_BATCHED(double)
#undef SVDB_INST_BATCHED

template int rank_update<float>(Ctx*, float*, size_t, size_t, size_t, size_t, const float*, const float*, size_t);
template int rank_update<double>(Ctx*, double*, size_t, size_t, size_t, size_t, const double*, const double*, size_t);
template int gemm_tn<float>(Ctx*, const float*, const float*, size_t, size_t, size_t, size_t, float*);
template int gemm_tn<double>(Ctx*, const double*, const double*, size_t, size_t, size_t, size_t, double*);
template int gemm_nn<float>(Ctx*, const float*, size_t, size_t, size_t, size_t, const float*, float*);
template int gemm_nn<double>(Ctx*, const double*, size_t, size_t, size_t, size_t, const double*, double*);

}  // namespace svdb200
