// Stage 1, tile order: flat-tree tile QR/LQ dense -> band in the exact task order AND arithmetic
// order of csc586::parallel::brd_p1<T> (svd_parallel.h:411-533) with its kernels qr (133-169),
// lq (189-226), hholder_compact (97-113), qr_apply (243-254), lq_apply (271-281) and
// serial::householder (svd_serial.h:189-216).  This is the parity path: it reproduces
// data/band_* bit-for-bit, including the sign of every band entry, which a full-height panel
// reduction can only match modulo D1*B*D2 (SURVEY 8a').
//
// Per tile column k the reference runs two dependent chains (TSQRT down the column, TSLQT along
// the row) each followed by an embarrassingly parallel update of tile pairs.  On the GPU:
//   chain kernel  : one CTA walks the chain with the stacked panel resident in shared memory and
//                   emits, per chain step, the explicit block reflector (Q or P, exactly as the
//                   reference forms it) into an L2-resident array;
//   apply kernel  : every thread group owns a few columns (QR) / rows (LQ) of the trailing matrix,
//                   keeps the tile-row-k (tile-column-k+1) part on chip across the whole chain and
//                   streams the other tile of each pair through HBM exactly once per half-step.
// All sums are k-ascending from +0 with separate multiply and add (no FMA contraction), scalars of
// the reflector in double: see RN<T> / householder_scalars in common.cuh.  Skipped terms are
// exactly those whose product is a structural +-0 (adding +-0 to a sum that started at +0 is the
// identity in IEEE arithmetic), so skipping them is bit-exact.
#include "common.cuh"

namespace svdb200 {
namespace {

// ---- bit-faithful panel QR of an m x t panel held in smem (A0, ld = t) ------------------------------
// On exit A0 <- A0 - V*Y (svd_parallel.h:167), V (m x t, ld t) and S (t x t, global) are valid.
template <typename T>
__device__ void qr_faithful(T* A0, T* Pm, T* V, T* Y, T* S, int m, int t, T* w, T* z, T* z2, T* sc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < m * t; e += nt) { Pm[e] = (T)0; V[e] = (T)0; }
    for (int e = tid; e < t * t; e += nt) { Y[e] = (T)0; S[e] = (T)0; }
    __syncthreads();
    const int kmax = min(t, m);
    for (int j = 0; j < kmax; ++j) {
        const int len = m - j;
        if (tid == 0) {
            T acc = (T)0;
            for (int i = 0; i < len; ++i) {
                T x = RN<T>::sub(A0[(j + i) * t + j], Pm[(j + i) * t + j]);
                acc = RN<T>::add(acc, RN<T>::mul(x, x));
            }
            T x0 = RN<T>::sub(A0[j * t + j], Pm[j * t + j]);
            T alpha, tau;
            householder_scalars<T>(x0, RN<T>::sqrt(acc), alpha, tau);
            sc[0] = alpha; sc[1] = tau;
        }
        __syncthreads();
        const T alpha = sc[0], tau = sc[1], mtau = -sc[1];
        for (int i = tid; i < len; i += nt)
            w[i] = (i == 0) ? (T)1 : RN<T>::mul(RN<T>::sub(A0[(j + i) * t + j], Pm[(j + i) * t + j]), alpha);
        __syncthreads();
        for (int c = j + tid; c < t; c += nt) {           // y = tau * R[j:,j:]^T w   (153-155)
            T acc = (T)0;
            for (int r = 0; r < len; ++r)
                acc = RN<T>::add(acc, RN<T>::mul(RN<T>::sub(A0[(j + r) * t + c], Pm[(j + r) * t + c]), w[r]));
            Y[j * t + c] = RN<T>::mul(acc, tau);
        }
        for (int i = tid; i < len; i += nt) V[(j + i) * t + j] = w[i];
        __syncthreads();
        if (j == 0) {
            if (tid == 0) S[0] = mtau;
        } else {                                          // hholder_compact (97-113)
            for (int r = tid; r < j; r += nt) {
                T acc = (T)0;
                for (int i = j; i < m; ++i) acc = RN<T>::add(acc, RN<T>::mul(V[i * t + r], V[i * t + j]));
                z[r] = acc;
            }
            __syncthreads();
            for (int r = tid; r < j; r += nt) {
                T acc = (T)0;
                for (int c = r; c < j; ++c) acc = RN<T>::add(acc, RN<T>::mul(S[r * t + c], z[c]));
                z2[r] = acc;
            }
            __syncthreads();
            for (int r = tid; r < j; r += nt) S[r * t + j] = RN<T>::mul(z2[r], mtau);
            if (tid == 0) S[j * t + j] = mtau;
        }
        __syncthreads();
        const int nc = t - j;                             // next partial sum of V*Y (k = j term)
        for (int e = tid; e < len * nc; e += nt) {
            int i = j + e / nc, c = j + e % nc;
            Pm[i * t + c] = RN<T>::add(Pm[i * t + c], RN<T>::mul(V[i * t + j], Y[j * t + c]));
        }
        __syncthreads();
    }
    for (int e = tid; e < m * t; e += nt) A0[e] = RN<T>::sub(A0[e], Pm[e]);
    __syncthreads();
}

// ---- bit-faithful panel LQ of a t x nn panel held in smem (A0, ld = nn) ----------------------------
// On exit A0 <- A0 - X*U (224), U (t x nn, ld nn) and S (t x t, global) are valid.
template <typename T>
__device__ void lq_faithful(T* A0, T* Pm, T* U, T* X, T* S, int t, int nn, T* w, T* z, T* z2, T* sc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < t * nn; e += nt) { Pm[e] = (T)0; U[e] = (T)0; }
    for (int e = tid; e < t * t; e += nt) { X[e] = (T)0; S[e] = (T)0; }
    __syncthreads();
    const int kmax = min(t, nn);
    for (int i = 0; i < kmax; ++i) {
        const int len = nn - i;
        if (tid == 0) {
            T acc = (T)0;
            for (int c = 0; c < len; ++c) {
                T x = RN<T>::sub(A0[i * nn + i + c], Pm[i * nn + i + c]);
                acc = RN<T>::add(acc, RN<T>::mul(x, x));
            }
            T x0 = RN<T>::sub(A0[i * nn + i], Pm[i * nn + i]);
            T alpha, tau;
            householder_scalars<T>(x0, RN<T>::sqrt(acc), alpha, tau);
            sc[0] = alpha; sc[1] = tau;
        }
        __syncthreads();
        const T alpha = sc[0], tau = sc[1], mtau = -sc[1];
        for (int c = tid; c < len; c += nt)
            w[c] = (c == 0) ? (T)1 : RN<T>::mul(RN<T>::sub(A0[i * nn + i + c], Pm[i * nn + i + c]), alpha);
        __syncthreads();
        for (int r = i + tid; r < t; r += nt) {           // x = tau * L[i:,i:] w   (210-212)
            T acc = (T)0;
            for (int c = 0; c < len; ++c)
                acc = RN<T>::add(acc, RN<T>::mul(RN<T>::sub(A0[r * nn + i + c], Pm[r * nn + i + c]), w[c]));
            X[r * t + i] = RN<T>::mul(acc, tau);
        }
        for (int c = tid; c < len; c += nt) U[i * nn + i + c] = w[c];
        __syncthreads();
        if (i == 0) {
            if (tid == 0) S[0] = mtau;
        } else {
            for (int r = tid; r < i; r += nt) {           // z = U_k u_i over all columns; u_i[c] = 0 for c < i
                T acc = (T)0;
                for (int c = i; c < nn; ++c) acc = RN<T>::add(acc, RN<T>::mul(U[r * nn + c], U[i * nn + c]));
                z[r] = acc;
            }
            __syncthreads();
            for (int r = tid; r < i; r += nt) {
                T acc = (T)0;
                for (int c = r; c < i; ++c) acc = RN<T>::add(acc, RN<T>::mul(S[r * t + c], z[c]));
                z2[r] = acc;
            }
            __syncthreads();
            for (int r = tid; r < i; r += nt) S[r * t + i] = RN<T>::mul(z2[r], mtau);
            if (tid == 0) S[i * t + i] = mtau;
        }
        __syncthreads();
        const int nr = t - i;
        for (int e = tid; e < nr * len; e += nt) {
            int r = i + e / len, c = i + e % len;
            Pm[r * nn + c] = RN<T>::add(Pm[r * nn + c], RN<T>::mul(X[r * t + i], U[i * nn + c]));
        }
        __syncthreads();
    }
    for (int e = tid; e < t * nn; e += nt) A0[e] = RN<T>::sub(A0[e], Pm[e]);
    __syncthreads();
}

// Q2 = V (S V^T)  (m x m, row-major, to global); Q1 (t x m) is staged in `tmp` (smem).  243-249
template <typename T>
__device__ void form_q(T* Q2, const T* S, const T* V, int m, int t, T* tmp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < t * m; e += nt) {
        int r = e / m, c = e - r * m;
        T acc = (T)0;
        for (int k = 0; k < t; ++k) acc = RN<T>::add(acc, RN<T>::mul(S[r * t + k], V[c * t + k]));
        tmp[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += nt) {
        int i = e / m, c = e - i * m;
        T acc = (T)0;
        for (int k = 0; k < t; ++k) acc = RN<T>::add(acc, RN<T>::mul(V[i * t + k], tmp[k * m + c]));
        Q2[e] = acc;
    }
    __syncthreads();
}
// P = U^T (S U)  (nn x nn, row-major, to global); P1 (t x nn) staged in tmp.  275-277
template <typename T>
__device__ void form_p(T* P, const T* S, const T* U, int t, int nn, T* tmp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < t * nn; e += nt) {
        int r = e / nn, c = e - r * nn;
        T acc = (T)0;
        for (int k = 0; k < t; ++k) acc = RN<T>::add(acc, RN<T>::mul(S[r * t + k], U[k * nn + c]));
        tmp[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < nn * nn; e += nt) {
        int i = e / nn, c = e - i * nn;
        T acc = (T)0;
        for (int k = 0; k < t; ++k) acc = RN<T>::add(acc, RN<T>::mul(U[k * nn + i], tmp[k * nn + c]));
        P[e] = acc;
    }
    __syncthreads();
}

// ---- chain kernels (one CTA) ----------------------------------------------------------------------------
// QR chain of tile column k: GEQRT(k,k) [factor_1tile, 296-308] then TSQRT([R_kk; A_ik]) for
// i = k+1..nbt-1 [factor_2tile, 311-339].  Qkk: t x t; Qts[i]: 2t x 2t.
template <typename T>
__global__ void tile_chain_qr_kernel(T* __restrict__ A, int n, int t, int k, int nbt, T* __restrict__ Qkk,
                                     T* __restrict__ Qts, T* __restrict__ Yg, T* __restrict__ Sg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    T* A0 = reinterpret_cast<T*>(smem_raw);   // 2t x t
    T* Pm = A0 + 2 * t * t;
    T* V = Pm + 2 * t * t;
    T* w = V + 2 * t * t;                      // 2t
    T* z = w + 2 * t;
    T* z2 = z + t;
    T* sc = z2 + t;
    T* Akk = A + (size_t)(k * t) * n + k * t;
    for (int e = tid; e < t * t; e += nt) A0[e] = Akk[(size_t)(e / t) * n + e % t];
    __syncthreads();
    qr_faithful<T>(A0, Pm, V, Yg, Sg, t, t, w, z, z2, sc);
    form_q<T>(Qkk, Sg, V, t, t, Pm);
    for (int i = k + 1; i < nbt; ++i) {
        T* Aik = A + (size_t)(i * t) * n + k * t;
        for (int e = tid; e < t * t; e += nt) A0[t * t + e] = Aik[(size_t)(e / t) * n + e % t];
        __syncthreads();
        qr_faithful<T>(A0, Pm, V, Yg, Sg, 2 * t, t, w, z, z2, sc);
        for (int e = tid; e < t * t; e += nt) Aik[(size_t)(e / t) * n + e % t] = A0[t * t + e];
        form_q<T>(Qts + (size_t)i * 4 * t * t, Sg, V, 2 * t, t, Pm);
    }
    for (int e = tid; e < t * t; e += nt) Akk[(size_t)(e / t) * n + e % t] = A0[e];
}

// LQ chain of tile row k: GELQT(k,k+1) then TSLQT([L | A_k,i]) for i = k+2..nbt-1.
template <typename T>
__global__ void tile_chain_lq_kernel(T* __restrict__ A, int n, int t, int k, int nbt, T* __restrict__ Pkk,
                                     T* __restrict__ Pts, T* __restrict__ Xg, T* __restrict__ Sg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    T* A0 = reinterpret_cast<T*>(smem_raw);   // t x 2t
    T* Pm = A0 + 2 * t * t;
    T* U = Pm + 2 * t * t;
    T* w = U + 2 * t * t;
    T* z = w + 2 * t;
    T* z2 = z + t;
    T* sc = z2 + t;
    T* Ak1 = A + (size_t)(k * t) * n + (k + 1) * t;
    // single tile first: panel t x t with ld = t
    for (int e = tid; e < t * t; e += nt) A0[e] = Ak1[(size_t)(e / t) * n + e % t];
    __syncthreads();
    lq_faithful<T>(A0, Pm, U, Xg, Sg, t, t, w, z, z2, sc);
    form_p<T>(Pkk, Sg, U, t, t, Pm);
    if (k + 2 < nbt) {
        // re-lay the left tile with ld = 2t (back to front so that nothing is overwritten early)
        __syncthreads();
        if (tid == 0) {
            for (int r = t - 1; r >= 1; --r)
                for (int c = t - 1; c >= 0; --c) A0[r * 2 * t + c] = A0[r * t + c];
        }
        __syncthreads();
    }
    for (int i = k + 2; i < nbt; ++i) {
        T* Aki = A + (size_t)(k * t) * n + i * t;
        for (int e = tid; e < t * t; e += nt) A0[(e / t) * 2 * t + t + e % t] = Aki[(size_t)(e / t) * n + e % t];
        __syncthreads();
        lq_faithful<T>(A0, Pm, U, Xg, Sg, t, 2 * t, w, z, z2, sc);
        for (int e = tid; e < t * t; e += nt) Aki[(size_t)(e / t) * n + e % t] = A0[(e / t) * 2 * t + t + e % t];
        form_p<T>(Pts + (size_t)i * 4 * t * t, Sg, U, t, 2 * t, Pm);
    }
    const int ldl = (k + 2 < nbt) ? 2 * t : t;
    for (int e = tid; e < t * t; e += nt) Ak1[(size_t)(e / t) * n + e % t] = A0[(e / t) * ldl + e % t];
}

// ---- apply kernel ----------------------------------------------------------------------------------------
// A "vector" is a column (QR side: elements strided by n) or a row (LQ side: contiguous) of the
// trailing matrix restricted to two tiles.  For vector v and chain step i:
//     x = [keep-part(t) ; stream-part_i(t)],  x <- x + M_i^T-like product:  out[o] = sum_kk M_i[kk][o] * x[kk]
// (QR: Q2[k][i]*A[k][c], svd_parallel.h:250-253 via mm 243-246; LQ: A[r][k]*P[k][c], 280.)
constexpr int kVpt = 4;   // vectors per thread

template <typename T, bool kRowVectors>
__global__ void tile_apply_kernel(T* __restrict__ A, int n, int t, int keep_tile, int first_stream_tile, int nbt,
                                  int vec_begin, int vec_end, const T* __restrict__ M1, const T* __restrict__ Ms_g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int L = 2 * t;
    const int groups = nt / L > 0 ? nt / L : 1;           // thread groups of L threads
    const int NV = groups * kVpt;
    T* Msm = reinterpret_cast<T*>(smem_raw);              // L x L
    T* xs = Msm + L * L;                                  // NV x L
    const int v0 = vec_begin + blockIdx.x * NV;
    const int nv = min(NV, vec_end - v0);
    if (nv <= 0) return;
    // element (vector v, position p in tile `tile`) in global memory
    auto gaddr = [&](int v, int tile, int p) -> size_t {
        return kRowVectors ? (size_t)(v0 + v) * n + (size_t)tile * t + p
                           : ((size_t)tile * t + p) * n + (size_t)(v0 + v);
    };
    // keep-part
    for (int e = tid; e < nv * t; e += nt) {
        int v, p;
        if (kRowVectors) { v = e / t; p = e - v * t; } else { p = e / nv; v = e - p * nv; }
        xs[v * L + p] = A[gaddr(v, keep_tile, p)];
    }
    // apply_1tile with the t x t matrix M1 (347-359)
    for (int e = tid; e < t * t; e += nt) Msm[e] = M1[e];
    __syncthreads();
    {
        T outv[kVpt];
        const int o = tid % L, grp = tid / L;
        const bool act = (grp < groups) && (o < t);
#pragma unroll
        for (int q = 0; q < kVpt; ++q) outv[q] = (T)0;
        if (act) {
            for (int kk = 0; kk < t; ++kk) {
                T mv = Msm[kk * t + o];
#pragma unroll
                for (int q = 0; q < kVpt; ++q) outv[q] = RN<T>::add(outv[q], RN<T>::mul(mv, xs[(grp * kVpt + q) * L + kk]));
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int q = 0; q < kVpt; ++q) {
                int v = grp * kVpt + q;
                if (v < nv) xs[v * L + o] = RN<T>::add(xs[v * L + o], outv[q]);
            }
        }
        __syncthreads();
    }
    // chain steps (apply_2tile, 363-391)
    for (int i = first_stream_tile; i < nbt; ++i) {
        const T* Mi = Ms_g + (size_t)i * L * L;
        for (int e = tid; e < L * L; e += nt) Msm[e] = Mi[e];
        for (int e = tid; e < nv * t; e += nt) {
            int v, p;
            if (kRowVectors) { v = e / t; p = e - v * t; } else { p = e / nv; v = e - p * nv; }
            xs[v * L + t + p] = A[gaddr(v, i, p)];
        }
        __syncthreads();
        T outv[kVpt];
        const int o = tid % L, grp = tid / L;
        const bool act = grp < groups;
#pragma unroll
        for (int q = 0; q < kVpt; ++q) outv[q] = (T)0;
        if (act) {
            for (int kk = 0; kk < L; ++kk) {
                T mv = Msm[kk * L + o];
#pragma unroll
                for (int q = 0; q < kVpt; ++q) outv[q] = RN<T>::add(outv[q], RN<T>::mul(mv, xs[(grp * kVpt + q) * L + kk]));
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int q = 0; q < kVpt; ++q) {
                int v = grp * kVpt + q;
                if (v < nv) xs[v * L + o] = RN<T>::add(xs[v * L + o], outv[q]);
            }
        }
        __syncthreads();
        for (int e = tid; e < nv * t; e += nt) {
            int v, p;
            if (kRowVectors) { v = e / t; p = e - v * t; } else { p = e / nv; v = e - p * nv; }
            A[gaddr(v, i, p)] = xs[v * L + t + p];
        }
        __syncthreads();
    }
    for (int e = tid; e < nv * t; e += nt) {
        int v, p;
        if (kRowVectors) { v = e / t; p = e - v * t; } else { p = e / nv; v = e - p * nv; }
        A[gaddr(v, keep_tile, p)] = xs[v * L + p];
    }
}

}  // namespace

template <typename T>
int stage1_tile_order(Ctx* c, T* a, size_t n, size_t band) {
    if (band == 0 || n == 0 || n % band != 0) return SVDB200_E_SHAPE;
    if (n > c->max_n || band > c->band) return SVDB200_E_CAPACITY;
    const int t = (int)band, nbt = (int)(n / band), ni = (int)n;
    const int L = 2 * t;
    size_t smem_chain = ((size_t)6 * t * t + 4 * (size_t)t + 8) * sizeof(T);
    if (smem_chain > 227 * 1024) return SVDB200_E_CAPACITY;
    int nt_chain = ((2 * t * t + 31) / 32) * 32;
    if (nt_chain > 512) nt_chain = 512;
    if (nt_chain < 32) nt_chain = 32;
    int nt_apply = L <= 256 ? 256 : L;
    if (nt_apply < L) nt_apply = L;
    const int groups = nt_apply / L;
    const int NV = groups * kVpt;
    size_t smem_apply = ((size_t)L * L + (size_t)NV * L) * sizeof(T);
    if (smem_apply > 227 * 1024) return SVDB200_E_CAPACITY;
    auto kq = tile_chain_qr_kernel<T>;
    auto kl = tile_chain_lq_kernel<T>;
    auto kaq = tile_apply_kernel<T, false>;
    auto kal = tile_apply_kernel<T, true>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain));
    SVDB_CHECK(c, cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chain));
    SVDB_CHECK(c, cudaFuncSetAttribute(kaq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply));
    SVDB_CHECK(c, cudaFuncSetAttribute(kal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply));
    T* st = reinterpret_cast<T*>(c->tilestate);
    T* Qkk = st;                       // t x t
    T* Yg = Qkk + (size_t)t * t;       // t x t  (Y of qr / X of lq)
    T* Sg = Yg + (size_t)t * t;        // t x t
    T* Qts = reinterpret_cast<T*>(c->tileq);
    for (int k = 0; k < nbt; ++k) {
        kq<<<1, nt_chain, smem_chain, c->stream>>>(a, ni, t, k, nbt, Qkk, Qts, Yg, Sg);
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
        if (k + 1 >= nbt) break;
        {
            int vb = (k + 1) * t, ve = ni;
            int blocks = (ve - vb + NV - 1) / NV;
            kaq<<<blocks, nt_apply, smem_apply, c->stream>>>(a, ni, t, k, k + 1, nbt, vb, ve, Qkk, Qts);
            SVDB_CHECK(c, cudaGetLastError());
            c->launches++;
        }
        kl<<<1, nt_chain, smem_chain, c->stream>>>(a, ni, t, k, nbt, Qkk, Qts, Yg, Sg);
        SVDB_CHECK(c, cudaGetLastError());
        c->launches++;
        {
            int vb = (k + 1) * t, ve = ni;
            int blocks = (ve - vb + NV - 1) / NV;
            kal<<<blocks, nt_apply, smem_apply, c->stream>>>(a, ni, t, k + 1, k + 2, nbt, vb, ve, Qkk, Qts);
            SVDB_CHECK(c, cudaGetLastError());
            c->launches++;
        }
    }
    return 0;
}

template int stage1_tile_order<float>(Ctx*, float*, size_t, size_t);
template int stage1_tile_order<double>(Ctx*, double*, size_t, size_t);

}  // namespace svdb200
