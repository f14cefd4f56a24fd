// Register-resident panel factorisation (band <= 64, panels up to 148 * 256 rows): the fast path of
// the stage-1 panel QR/LQ, same algorithm and outputs as panel_factor_kernel in stage1_panel.cu
// (which stays as the general fallback), but
//   * each CTA keeps its slice of the panel in REGISTERS for the whole column loop: a warp owns the
//     rows w, w+8, w+16, ... of the slice, a lane owns column `lane` (and `lane+32` for band 64);
//   * the pivot column is broadcast inside the warp with shuffles, so the rank-1 update of column j
//     and the dot products needed for column j+1 are ONE fused pass over the registers
//     (the shared-memory version makes two passes per column and was bound by their latency);
//   * per column: one cross-warp reduction, one all-reduce across CTAs (DSMEM + barrier.cluster for
//     clusters of <= 16 CTAs, L2 + software grid barrier otherwise), scalars recomputed per thread.
// Shared memory is only used for the reductions, the transposed load/store of LQ row panels and the
// epilogue (V, V2 = V S^T by back substitution with T^-1 = D + striu(V^T V), R/L write-back).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace svdb200 {
namespace {

constexpr int kThreads = 256, kWarps = 8;

template <typename T, int CPL>
__device__ __forceinline__ T pick(const T (&v)[CPL], int u) {
    T r = v[0];
#pragma unroll
    for (int q = 1; q < CPL; ++q) r = (u == q) ? v[q] : r;
    return r;
}

template <typename T, bool kTrans, bool kCluster, int RPT, int CPL>
__global__ void __launch_bounds__(kThreads)
panel_reg_kernel(T* __restrict__ A, size_t lda, int m, int b, T* __restrict__ V, T* __restrict__ V2, T* __restrict__ red,
                 unsigned* __restrict__ bar) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ROWS = RPT * kWarps;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5;
    const int G = gridDim.x, g = blockIdx.x;
    const int r0 = g * ROWS;
    const int R = max(0, min(ROWS, m - r0));
    const int ld = b + 1, slot = 2 * b;
    T* lred = reinterpret_cast<T*>(smem_raw);            // 2 * slot
    T* Ps = lred + 2 * slot;                             // ROWS x ld (transposed staging + epilogue)
    T* Gm = Ps + (size_t)ROWS * ld;                      // b x b
    T* zs = Gm + b * b;                                  // b
    T* piv = zs + b;                                     // b
    T* taus = piv + b;                                   // b
    T* psum = taus + b;                                  // kWarps * b (>= blockDim)
    unsigned gen = 0;
    int cu[CPL];
    bool valid[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) { cu[u] = lane + 32 * u; valid[u] = cu[u] < b; }
    const int tx = tid % b, tyy = tid / b, rgroups = max(1, nt / b);
    const bool in2d = tyy < rgroups;

    // ---- load ----------------------------------------------------------------------------------------
    T a[RPT][CPL];
    if (!kTrans) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int rl = w + kWarps * i;
#pragma unroll
            for (int u = 0; u < CPL; ++u) a[i][u] = (rl < R && valid[u]) ? A[(size_t)(r0 + rl) * lda + cu[u]] : (T)0;
        }
    } else {
        for (int c = w; c < b; c += kWarps)
            for (int rl = lane; rl < R; rl += 32) Ps[rl * ld + c] = A[(size_t)c * lda + (r0 + rl)];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int rl = w + kWarps * i;
#pragma unroll
            for (int u = 0; u < CPL; ++u) a[i][u] = (rl < R && valid[u]) ? Ps[rl * ld + cu[u]] : (T)0;
        }
    }
    for (int e = tid; e < b * b; e += nt) Gm[e] = (T)0;

    const int kmax = min(b, m);
    // dots of column 0 (rows > 0) with every column
    T acc[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) acc[u] = (T)0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int rl = w + kWarps * i;
        const T p0 = __shfl_sync(0xffffffffu, a[i][0], 0);
        if (rl < R && r0 + rl > 0) {
#pragma unroll
            for (int u = 0; u < CPL; ++u) acc[u] += a[i][u] * p0;
        }
    }
    __syncthreads();

    for (int j = 0; j < kmax; ++j) {
        // ---- cross-warp reduction of the local dots, publication -----------------------------------------
#pragma unroll
        for (int u = 0; u < CPL; ++u) if (valid[u]) psum[w * b + cu[u]] = acc[u];
        T* mine = kCluster ? lred + (j & 1) * slot : red + ((size_t)(j & 1) * (G + 1) + g) * slot;
        T* pivslot = kCluster ? lred + (j & 1) * slot + b : red + ((size_t)(j & 1) * (G + 1) + G) * slot;
        {   // the warp that holds global row j publishes it (compile-time register indices only)
            const int jl = j - r0;
            if (jl >= 0 && jl < R && (jl % kWarps) == w) {
                const int ij = jl / kWarps;
#pragma unroll
                for (int i = 0; i < RPT; ++i)
                    if (i == ij) {
#pragma unroll
                        for (int u = 0; u < CPL; ++u)
                            if (valid[u]) { if (kCluster) pivslot[cu[u]] = a[i][u]; else st_cg(&pivslot[cu[u]], a[i][u]); }
                    }
            }
        }
        __syncthreads();
        for (int c = tid; c < b; c += nt) {
            T s = psum[c];
#pragma unroll
            for (int q = 1; q < kWarps; ++q) s += psum[q * b + c];
            if (kCluster) mine[c] = s; else st_cg(&mine[c], s);
        }
        if (kCluster) cg::this_cluster().sync(); else grid_barrier(bar, (unsigned)G, gen);
        // ---- all-reduce across CTAs in a fixed association -------------------------------------------------
        {
            const int chunk = (G + rgroups - 1) / rgroups;
            const int jowner = j / ROWS;
            if (in2d) {
                const int c = tx, part = tyy;
                const int q0 = part * chunk, q1 = min(G, q0 + chunk);
                T s = (T)0;
                if (kCluster) {
                    cg::cluster_group cl = cg::this_cluster();
#pragma unroll 4
                    for (int q = q0; q < q1; ++q) s += cl.map_shared_rank(lred, q)[(j & 1) * slot + c];
                    if (part == 0) piv[c] = cl.map_shared_rank(lred, jowner)[(j & 1) * slot + b + c];
                } else {
                    const T* buf = red + (size_t)(j & 1) * (G + 1) * slot;
#pragma unroll 8
                    for (int q = q0; q < q1; ++q) s += ld_cg(&buf[(size_t)q * slot + c]);
                    if (part == 0) piv[c] = ld_cg(&buf[(size_t)G * slot + c]);
                }
                psum[part * b + c] = s;
            }
            __syncthreads();
            for (int c = tid; c < b; c += nt) {
                T s = psum[c];
                for (int part = 1; part < rgroups; ++part) s += psum[part * b + c];
                zs[c] = s;
            }
        }
        __syncthreads();
        // ---- scalars (every thread), Gram column ---------------------------------------------------------------
        const T x0 = piv[j];
        const T nrm = sqrt(zs[j] + x0 * x0);
        const double sgn = -copysign(1.0, (double)x0);
        const double u1 = (double)x0 - sgn * (double)nrm;
        const T alpha = (T)(1.0 / u1);
        const T tau = (T)(-sgn * u1 / (double)nrm);
        const T beta = (T)(sgn * (double)nrm);           // R_jj = -sign(x0) ||x||
        T fsr[CPL];
#pragma unroll
        for (int u = 0; u < CPL; ++u) fsr[u] = valid[u] ? tau * (piv[cu[u]] + alpha * zs[cu[u]]) : (T)0;
        if (w == 0) {
#pragma unroll
            for (int u = 0; u < CPL; ++u)
                if (valid[u] && cu[u] < j) Gm[cu[u] * b + j] = piv[cu[u]] + alpha * zs[cu[u]];
            if (lane == 0) taus[j] = tau;
        }
        // ---- fused pass: rank-1 update of column j, dots for column j+1 ---------------------------------------
        const int lj = j & 31, uj = j >> 5, ln = (j + 1) & 31, un = (j + 1) >> 5;
        const bool more = (j + 1 < kmax);
#pragma unroll
        for (int u = 0; u < CPL; ++u) acc[u] = (T)0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int rl = w + kWarps * i;
            const int grow = r0 + rl;
            const T xj = __shfl_sync(0xffffffffu, pick<T, CPL>(a[i], uj), lj);
            if (rl < R && grow >= j) {
                const T wv = (grow == j) ? (T)1 : xj * alpha;
#pragma unroll
                for (int u = 0; u < CPL; ++u) {
                    if (cu[u] > j) a[i][u] -= wv * fsr[u];
                    else if (cu[u] == j) a[i][u] = (grow == j) ? beta : wv;
                }
            }
            if (more) {
                const T pn = __shfl_sync(0xffffffffu, pick<T, CPL>(a[i], un), ln);
                if (rl < R && grow > j + 1) {
#pragma unroll
                    for (int u = 0; u < CPL; ++u) acc[u] += a[i][u] * pn;
                }
            }
        }
        // psum / zs / piv are rewritten only after the next __syncthreads-protected phases
        __syncthreads();
    }

    // ---- epilogue ------------------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int rl = w + kWarps * i;
        if (rl < R) {
#pragma unroll
            for (int u = 0; u < CPL; ++u) if (valid[u]) Ps[rl * ld + cu[u]] = a[i][u];
        }
    }
    __syncthreads();
    const int nwarps = kWarps;
    if (in2d)
        for (int rl = tyy; rl < R; rl += rgroups) {
            const int row = r0 + rl, c = tx;
            T vv = (row == c) ? (T)1 : (row > c ? Ps[rl * ld + c] : (T)0);
            if (c >= kmax) vv = (T)0;
            V[(size_t)row * b + c] = vv;
            if (!kTrans) A[(size_t)row * lda + c] = (c >= row) ? Ps[rl * ld + c] : (T)0;
        }
    if (kTrans)
        for (int c = w; c < b; c += nwarps)
            for (int rl = lane; rl < R; rl += 32) {
                const int row = r0 + rl;
                A[(size_t)c * lda + row] = (c >= row) ? Ps[rl * ld + c] : (T)0;
            }
    __syncthreads();
    // every row x of V2 solves x (D + U)^T = -v by back substitution (see stage1_panel.cu)
    for (int rl = tid; rl < R; rl += nt) {
        const int row = r0 + rl;
        T* x = Ps + rl * ld;
        const int khi = min(kmax - 1, row);
        for (int c = b - 1; c > khi; --c) x[c] = (T)0;
        for (int c = khi; c >= 0; --c) {
            T s = (row == c) ? (T)-1 : -x[c];
            for (int k = c + 1; k <= khi; ++k) s -= x[k] * Gm[c * b + k];
            x[c] = s * taus[c];
        }
    }
    __syncthreads();
    if (!kTrans) {
        if (in2d)
            for (int rl = tyy; rl < R; rl += rgroups) V2[(size_t)(r0 + rl) * b + tx] = Ps[rl * ld + tx];
    } else {
        for (int c = w; c < b; c += nwarps)
            for (int rl = lane; rl < R; rl += 32) V2[(size_t)c * m + (r0 + rl)] = Ps[rl * ld + c];
    }
    if (kCluster) cg::this_cluster().sync();
}

inline size_t reg_smem_bytes(int rows, int b, size_t esz) {
    return ((size_t)4 * b + (size_t)rows * (b + 1) + (size_t)b * b + 3 * (size_t)b + (size_t)kWarps * b + kThreads + 8) * esz;
}

template <typename T, bool kTrans, int RPT, int CPL>
int launch_reg(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream, bool cooperative) {
    constexpr int ROWS = RPT * kWarps;
    const int G = (m + ROWS - 1) / ROWS;
    if (G > c->num_sms || G > kMaxPanelCtas) return 1;
    const size_t smem = reg_smem_bytes(ROWS, b, sizeof(T));
    T* red = reinterpret_cast<T*>(c->red);
    unsigned* bar = c->bar;
    if (c->cluster_ok && G <= c->cluster_ok) {
        auto kern = panel_reg_kernel<T, kTrans, true, RPT, CPL>;
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (G > 8) SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(G);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = G;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, lda, m, b, V, V2, red, bar);
        if (e == cudaSuccess) { c->launches++; return 0; }
        cudaGetLastError();
        c->cluster_ok = c->cluster_ok > 8 ? 8 : 0;
        return launch_reg<T, kTrans, RPT, CPL>(c, a, lda, m, b, V, V2, stream, cooperative);
    }
    auto kern = panel_reg_kernel<T, kTrans, false, RPT, CPL>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVDB_CHECK(c, cudaMemsetAsync(c->bar, 0, 2 * sizeof(unsigned), stream));
    if (cooperative) {
        void* args[] = {&a, &lda, &m, &b, &V, &V2, &red, &bar};
        SVDB_CHECK(c, cudaLaunchCooperativeKernel((void*)kern, dim3(G), dim3(kThreads), args, smem, stream));
    } else {
        // look-ahead panels run beside the trailing update: a cooperative launch is gang-scheduled
        // and would wait for the update to drain; G <= #SMs CTAs of this size always become
        // co-resident once the update's CTAs retire, so the software barrier still completes.
        kern<<<G, kThreads, smem, stream>>>(a, lda, m, b, V, V2, red, bar);
        SVDB_CHECK(c, cudaGetLastError());
    }
    c->launches++;
    return 0;
}

}  // namespace

// returns 0 when it ran, 1 when the shape is outside this kernel's range (caller falls back)
template <typename T, bool kTrans>
int launch_panel_reg(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream, bool cooperative) {
    if (b <= 32) return launch_reg<T, kTrans, 32, 1>(c, a, lda, m, b, V, V2, stream, cooperative);
    if (b <= 64) return launch_reg<T, kTrans, 16, 2>(c, a, lda, m, b, V, V2, stream, cooperative);
    return 1;
}
template int launch_panel_reg<float, false>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t, bool);
template int launch_panel_reg<float, true>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t, bool);
template int launch_panel_reg<double, false>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t, bool);
template int launch_panel_reg<double, true>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t, bool);

}  // namespace svdb200
