// Blocked register-resident panel factorisation: ONE exchange per sub-panel of C = 8 columns instead of one per column.
// Same algorithm, sign convention (svd_serial.h:194-201: H x = -sign(x0) ||x|| e1) and outputs as panel_reg_kernel /
// panel_factor_kernel (R or L in A with exact zeros below the diagonal, V with explicit unit diagonal, V2 = V S^T with
// S = -T of the compact-WY form, svd_parallel.h:97-113) -- those kernels exchange the dot products of ONE pivot column
// per all-reduce (3.6 us per column inside a cluster, 6-7.5 us across clusters: the exposed part of stage 1).
//
// Per sub-panel J = {j .. j+C-1} the CTAs exchange, in one deterministic all-reduce,
//     D = A_lo[:, J]^T A_lo          (C x b)   "lo" = rows >= j + C of the CURRENT panel (finished columns hold v)
//     Top = rows j .. j+C-1 of A     (C x b)   (their owner contributes them, everybody else zeros)
// and every CTA then runs the C Householder steps redundantly on that small data: all later states of the lo part are
// linear combinations  A_lo[:, k] - A_lo[:, J] * Cm[:, k]  of the exchanged columns, so norms and dot products follow from
// D (x_lo^T y_lo = m^T (D[:,k] - D[:,J] Cm[:,k]) with x_lo = A_lo[:,J] m), while the C top rows are carried explicitly.
// The local rows are then updated once with the accumulated coefficients (a rank-C pass over the registers), fused with
// the D of the next sub-panel.  With C = 1 this is exactly the existing kernels' recurrence.
//
// Numerical safety: norms obtained through D lose log2(d_ii / ||x_lo||^2) bits when a column of the sub-panel is nearly
// dependent on its predecessors INSIDE the sub-panel (e.g. the common mean of U[0,5) inputs in the very first sub-panel).
// Every CTA checks that ratio on identical numbers before each step; when it falls below `guard` the sub-panel ends
// there (C_eff < C: progress >= 1 column per exchange, since the first step needs no subtraction) and the next exchange
// starts from freshly computed dot products.  Worst case = one column per exchange = the previous kernels.
//
// Transport: all-reduce of L = 2*C*b values as reduce-scatter + all-gather over DSMEM with flag-stamped words pushed into
// the peers' shared memory (ll_words.cuh; no barrier.cluster in the loop); panels taller than one 16-CTA cluster run as
// several clusters whose slice owners exchange their slice through L2 (flag-stamped, fixed summation order).
#include <cooperative_groups.h>
#include "common.cuh"
#include "ll_words.cuh"

namespace cg = cooperative_groups;

namespace svdb200 {
namespace {

constexpr int kThreads = 256, kWarps = 8;
constexpr int kC = 8;                  // sub-panel width (= kWarps: every warp owns exactly one of the C top rows)
constexpr int kMaxCS = 16, kMaxNC = 16;

#ifndef SVDB_PANEL_TIMING
#define SVDB_PANEL_TIMING 0
#endif
__device__ long long g_blk_dbg[16];
#define BLK_TICK(k)                                                          \
    do {                                                                     \
        if (SVDB_PANEL_TIMING && blockIdx.x == 0 && threadIdx.x == 0) {      \
            long long _t = clock64();                                        \
            g_blk_dbg[k] += _t - tick;                                       \
            tick = _t;                                                       \
        }                                                                    \
    } while (0)

template <typename T, int CPL>
__device__ __forceinline__ T pick(const T (&v)[CPL], int u) {
    T r = v[0];
#pragma unroll
    for (int q = 1; q < CPL; ++q) r = (u == q) ? v[q] : r;
    return r;
}
template <typename T> __device__ __forceinline__ T bcast(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <typename T> __device__ __forceinline__ T guard_ratio();
template <> __device__ __forceinline__ float guard_ratio<float>() { return 0.25f; }
template <> __device__ __forceinline__ double guard_ratio<double>() { return 1.0 / 64.0; }

struct BlkShape {
    int CS, NC, SL, L;
    size_t exch_bytes;      // 2 parities x (CS x SL in-words + L out-words) x 16 B
};
__host__ __device__ inline BlkShape blk_shape(int b, int CS, int NC) {
    BlkShape s;
    s.CS = CS; s.NC = NC; s.L = 2 * kC * b; s.SL = (s.L + CS - 1) / CS;
    s.exch_bytes = (size_t)2 * ((size_t)CS * s.SL + s.L) * 16;
    return s;
}

template <typename T, bool kTrans, int RPT, int CPL>
__global__ void __launch_bounds__(kThreads, 1)
panel_blk_kernel(T* __restrict__ A, size_t lda, int m, int b, T* __restrict__ V, T* __restrict__ V2, char* __restrict__ gbuf, int NC,
                 unsigned epoch, size_t stage_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ROWS = RPT * kWarps, C = kC;
    const int tid = threadIdx.x, nt = kThreads, lane = tid & 31, w = tid >> 5;
    const int G = gridDim.x, g = blockIdx.x;
    const int CS = G / NC, cl = g / CS, crank = g - cl * CS;
    const BlkShape sh = blk_shape(b, CS, NC);
    const int L = sh.L, SL = sh.SL, CB = C * b;
    const int r0 = g * ROWS;
    const int R = max(0, min(ROWS, m - r0));
    const int ld = b + 1;
    // ---- shared memory ----------------------------------------------------------------------------------
    unsigned char* stage = smem_raw;                      // Ps (ROWS x ld: transposed load, epilogue) / exchange words (loop)
    T* Ps = reinterpret_cast<T*>(stage);
    T* Gm = reinterpret_cast<T*>(stage + stage_bytes);    // b x b : Gm[c][j] = v_c^T v_j (c < j)
    T* taus = Gm + b * b;                                 // b
    T* psum = taus + b;                                   // kWarps x CB : per-warp dot products (reused for slice partials)
    T* topS = psum + kWarps * CB;                         // CB : this CTA's rows among the C top rows (zeros elsewhere)
    T* red = topS + CB;                                   // L  : reduced D (CB) followed by Top (CB)
    T* Kc = red + L;                                      // CB : pass coefficients per column
    T* TopF = Kc + CB;                                    // CB : final top rows of the sub-panel
    int* ctl = reinterpret_cast<int*>(TopF + CB);         // [0] = C_eff
    const unsigned l1_base = smem_addr(stage);
    const unsigned in_bytes = (unsigned)(2 * CS * SL * 16);
    auto in_off = [&](int par, int src, int e) { return (unsigned)(((par * CS + src) * SL + e) * 16); };
    auto out_off = [&](int par, int idx) { return in_bytes + (unsigned)((par * L + idx) * 16); };

    int cu[CPL];
    bool valid[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) { cu[u] = lane + 32 * u; valid[u] = cu[u] < b; }
    const int tx = tid % b, tyy = tid / b, rgroups = max(1, nt / b);
    const bool in2d = tyy < rgroups;

    // ---- load --------------------------------------------------------------------------------------------
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
    for (int e = tid; e < CB; e += nt) topS[e] = (T)0;
    __syncthreads();                                      // the transposed load is done with Ps
    if (G > 1) {
        unsigned* z = reinterpret_cast<unsigned*>(stage);
        for (int e = tid; e < (int)(sh.exch_bytes / 4); e += nt) z[e] = 0u;
        cg::this_cluster().sync();                        // every peer's word buffers are cleared before the first push
    }

    const int kmax = b;                                   // m >= b (checked by the launcher)
    T acc[C][CPL];
    // One pass over the local rows.  apply: update with the coefficients of the sub-panel that started at column j and
    // processed ceff columns; then (always) dot products and top rows for the sub-panel starting at jn.
    auto pass = [&](bool apply, int j, int ceff, int jn) {
        T K[C][CPL];
        bool fin[CPL];
#pragma unroll
        for (int u = 0; u < CPL; ++u) {
            fin[u] = apply && cu[u] >= j && cu[u] < j + ceff;
#pragma unroll
            for (int p = 0; p < C; ++p) K[p][u] = (apply && valid[u]) ? Kc[p * b + cu[u]] : (T)0;
        }
#pragma unroll
        for (int p = 0; p < C; ++p)
#pragma unroll
            for (int u = 0; u < CPL; ++u) acc[p][u] = (T)0;
        const bool more = jn < kmax;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int rl = w + kWarps * i;
            const int grow = r0 + rl;
            const bool inb = rl < R;                      // warp-uniform: a warp works on one row at a time
            if (apply && inb && grow >= j) {
                if (grow < j + C) {
                    const int t = grow - j;
#pragma unroll
                    for (int u = 0; u < CPL; ++u) if (valid[u]) a[i][u] = TopF[t * b + cu[u]];
                } else {
                    T x[C];
#pragma unroll
                    for (int p = 0; p < C; ++p) x[p] = bcast(pick<T, CPL>(a[i], (j + p) >> 5), (j + p) & 31);
#pragma unroll
                    for (int u = 0; u < CPL; ++u) {
                        T v = fin[u] ? (T)0 : a[i][u];
#pragma unroll
                        for (int p = 0; p < C; ++p) v += x[p] * K[p][u];
                        a[i][u] = v;
                    }
                }
            }
            if (more && inb && grow >= jn) {
                if (grow < jn + C) {
                    const int t = grow - jn;
#pragma unroll
                    for (int u = 0; u < CPL; ++u) if (valid[u]) topS[t * b + cu[u]] = a[i][u];
                } else {
                    T y[C];
#pragma unroll
                    for (int p = 0; p < C; ++p) {
                        const T yv = bcast(pick<T, CPL>(a[i], (jn + p) >> 5), (jn + p) & 31);
                        y[p] = (jn + p < kmax) ? yv : (T)0;
                    }
#pragma unroll
                    for (int p = 0; p < C; ++p)
#pragma unroll
                        for (int u = 0; u < CPL; ++u) acc[p][u] += y[p] * a[i][u];
                }
            }
        }
    };

    long long tick = SVDB_PANEL_TIMING ? clock64() : 0;
    (void)tick;
    pass(false, 0, 0, 0);
    BLK_TICK(0);
    int j = 0;
    unsigned round = 0;
    while (j < kmax) {
        // ---- publish the per-warp dot products ----------------------------------------------------------------
#pragma unroll
        for (int p = 0; p < C; ++p)
#pragma unroll
            for (int u = 0; u < CPL; ++u)
                if (valid[u]) psum[w * CB + p * b + cu[u]] = acc[p][u];
        __syncthreads();
        BLK_TICK(1);
        // ---- deterministic all-reduce of [D | Top] -----------------------------------------------------------------
        auto local_val = [&](int idx) -> T {
            if (idx >= CB) return topS[idx - CB];
            T s = psum[idx];
#pragma unroll
            for (int q = 1; q < kWarps; ++q) s += psum[q * CB + idx];
            return s;
        };
        if (G == 1) {
            for (int idx = tid; idx < L; idx += nt) red[idx] = local_val(idx);
            __syncthreads();
        } else {
            const unsigned seq = epoch * 128u + round + 1u;
            const int par = (int)(round & 1u);
            // hop 1 (reduce-scatter): slice q of the vector goes to the CTA of rank q in the cluster
            for (int idx = tid; idx < L; idx += nt) {
                const T v = local_val(idx);
                const int q = idx / SL, e = idx - q * SL;
                LLSmem<T>::push(map_to_rank(l1_base + in_off(par, crank, e), (unsigned)q), v, seq);
            }
            __syncthreads();                              // psum / topS have been read: psum is reused for the partials
            const int mine = max(0, min(SL, L - crank * SL));      // entries of my slice
            // the CS contributions of an entry are summed in chunks by `parts` threads, then in part order: one fixed
            // association on every CTA
            const int parts = SL <= nt ? max(1, min(CS, nt / SL)) : 1;
            const int chunk = (CS + parts - 1) / parts;
            SpinGuard sg;
            for (int item = tid; item < parts * SL; item += nt) {
                const int part = item / SL, e = item - part * SL;
                T s = (T)0;
                if (e < mine) {
                    const int q0 = part * chunk, q1 = min(CS, q0 + chunk);
                    for (int q = q0; q < q1; ++q) {
                        T v;
                        while (!LLSmem<T>::try_load(l1_base + in_off(par, q, e), seq, v)) sg.tick();
                        s += v;
                    }
                }
                psum[part * SL + e] = s;
            }
            __syncthreads();
            for (int e = tid; e < mine; e += nt) {
                T s = psum[e];
                for (int part = 1; part < parts; ++part) s += psum[part * SL + e];
                const int idx = crank * SL + e;
                if (NC > 1) {
                    // level 2: the owners of the same slice in the other clusters exchange their sums through L2
                    char* gs = gbuf + (size_t)par * NC * L * 16;
                    LLWord<T>::store(gs + ((size_t)cl * L + idx) * 16, s, seq);
                    T v[kMaxNC];
                    unsigned pending = (NC >= 32) ? 0xffffffffu : ((1u << NC) - 1u);
                    while (pending != 0u) {
#pragma unroll
                        for (int q = 0; q < kMaxNC; ++q)
                            if ((pending >> q) & 1u) {
                                if (LLWord<T>::try_load(gs + ((size_t)q * L + idx) * 16, seq, v[q])) pending &= ~(1u << q);
                            }
                        if (pending != 0u) { sg.tick(); __nanosleep(40); }
                    }
                    s = v[0];
#pragma unroll
                    for (int q = 1; q < kMaxNC; ++q) if (q < NC) s += v[q];
                }
                // hop 2 (all-gather): the finished entry goes to every CTA of the cluster
                for (int q = 0; q < CS; ++q) LLSmem<T>::push(map_to_rank(l1_base + out_off(par, idx), (unsigned)q), s, seq);
            }
            for (int idx = tid; idx < L; idx += nt) {
                T v;
                while (!LLSmem<T>::try_load(l1_base + out_off(par, idx), seq, v)) sg.tick();
                red[idx] = v;
            }
            __syncthreads();
        }
        BLK_TICK(2);
        for (int e = tid; e < CB; e += nt) topS[e] = (T)0;         // next round's top rows: only their owner writes them
        // ---- the C Householder steps on the exchanged data (warp 0; lane = column) ----------------------------------
        if (w == 0) {
            T Tk[CPL][C], Ck[CPL][C];
#pragma unroll
            for (int u = 0; u < CPL; ++u)
#pragma unroll
                for (int t = 0; t < C; ++t) {
                    Tk[u][t] = valid[u] ? red[CB + t * b + cu[u]] : (T)0;
                    Ck[u][t] = (T)0;
                }
            const T* D = red;
            int done = 0;
            bool active = true;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const int ji = j + i;
                active = active && (ji < kmax);
                if (active) {                              // warp-uniform
                    const int lj = ji & 31, uj = ji >> 5;
                    T mv[C], tj[C];
#pragma unroll
                    for (int p = 0; p < C; ++p) {
                        T sel = Ck[0][p];
#pragma unroll
                        for (int u = 1; u < CPL; ++u) sel = (uj == u) ? Ck[u][p] : sel;
                        const T cv = bcast(sel, lj);
                        mv[p] = (p <= i) ? (((p == i) ? (T)1 : (T)0) - cv) : (T)0;
                        T selt = Tk[0][p];
#pragma unroll
                        for (int u = 1; u < CPL; ++u) selt = (uj == u) ? Tk[u][p] : selt;
                        tj[p] = bcast(selt, lj);
                    }
                    T gl[CPL];
#pragma unroll
                    for (int u = 0; u < CPL; ++u) {
                        T s = (T)0;
                        if (valid[u]) {
#pragma unroll
                            for (int q = 0; q < C; ++q) if (q <= i) s += D[q * b + cu[u]] * mv[q];
                        }
                        gl[u] = s;
                    }
                    T gv[C];
#pragma unroll
                    for (int p = 0; p < C; ++p) {
                        const int col = min(j + p, kmax - 1);
                        gv[p] = (p <= i) ? bcast(pick<T, CPL>(gl, col >> 5), col & 31) : (T)0;
                    }
                    T s[CPL], dT[CPL];
#pragma unroll
                    for (int u = 0; u < CPL; ++u) {
                        T sv = gl[u], dv = (T)0;
#pragma unroll
                        for (int p = 0; p < C; ++p) {
                            if (p <= i) sv -= gv[p] * Ck[u][p];
                            if (p > i) dv += tj[p] * Tk[u][p];
                        }
                        s[u] = sv; dT[u] = dv;
                    }
                    const T sji = bcast(pick<T, CPL>(s, uj), lj);
                    const T dii = D[i * b + ji];
                    if (i > 0 && !(sji >= guard_ratio<T>() * dii)) {
                        active = false;                    // too much cancellation: end the sub-panel here
                    } else {
                        T top2 = (T)0;
#pragma unroll
                        for (int t = 0; t < C; ++t) if (t >= i) top2 += tj[t] * tj[t];
                        const T x0 = tj[i];
                        const T nrm = sqrt(sji + top2);
                        const double sgn = -copysign(1.0, (double)x0);
                        const double u1 = (double)x0 - sgn * (double)nrm;
                        const T alpha = (T)(1.0 / u1);
                        const T tau = (T)(-sgn * u1 / (double)nrm);
                        const T beta = (T)(sgn * (double)nrm);       // R_jj = -sign(x0) ||x||
#pragma unroll
                        for (int u = 0; u < CPL; ++u) {
                            const int col = cu[u];
                            const T dot = Tk[u][i] + alpha * (dT[u] + s[u]);
                            if (col > ji) {
                                const T f = tau * dot, fa = f * alpha;
                                Tk[u][i] -= f;
#pragma unroll
                                for (int p = 0; p < C; ++p) {
                                    if (p > i) Tk[u][p] -= fa * tj[p];
                                    if (p <= i) Ck[u][p] += fa * mv[p];
                                }
                            } else if (col == ji) {
                                Tk[u][i] = beta;
#pragma unroll
                                for (int p = 0; p < C; ++p) {
                                    if (p > i) Tk[u][p] = alpha * tj[p];
                                    if (p <= i) Ck[u][p] = ((p == i) ? (T)1 : (T)0) - alpha * mv[p];
                                }
                                taus[ji] = tau;
                            } else if (valid[u]) {
                                Gm[col * b + ji] = dot;    // v_col^T v_ji
                            }
                        }
                        done = i + 1;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < CPL; ++u)
                if (valid[u]) {
                    const int col = cu[u];
                    const bool f = col >= j && col < j + done;
#pragma unroll
                    for (int p = 0; p < C; ++p) {
                        Kc[p * b + col] = f ? (((p == col - j) ? (T)1 : (T)0) - Ck[u][p]) : -Ck[u][p];
                        TopF[p * b + col] = Tk[u][p];
                    }
                }
            if (lane == 0) ctl[0] = done;
        }
        __syncthreads();
        BLK_TICK(3);
        const int ceff = ctl[0];
        pass(true, j, ceff, j + ceff);
        j += ceff;
        round += 1;
        BLK_TICK(4);
        if (SVDB_PANEL_TIMING && blockIdx.x == 0 && threadIdx.x == 0) g_blk_dbg[8] += 1;
    }

    // ---- epilogue (as panel_reg_kernel) --------------------------------------------------------------------------
    if (G > 1) cg::this_cluster().sync();                 // nobody pushes words into the staging area any more
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int rl = w + kWarps * i;
        if (rl < R) {
#pragma unroll
            for (int u = 0; u < CPL; ++u) if (valid[u]) Ps[rl * ld + cu[u]] = a[i][u];
        }
    }
    __syncthreads();
    if (in2d)
        for (int rl = tyy; rl < R; rl += rgroups) {
            const int row = r0 + rl, c = tx;
            const T vv = (row == c) ? (T)1 : (row > c ? Ps[rl * ld + c] : (T)0);
            V[(size_t)row * b + c] = vv;
            if (!kTrans) A[(size_t)row * lda + c] = (c >= row) ? Ps[rl * ld + c] : (T)0;
        }
    if (kTrans)
        for (int c = w; c < b; c += kWarps)
            for (int rl = lane; rl < R; rl += 32) {
                const int row = r0 + rl;
                A[(size_t)c * lda + row] = (c >= row) ? Ps[rl * ld + c] : (T)0;
            }
    __syncthreads();
    // every row x of V2 solves x (D + U)^T = -v by back substitution (T^-1 = diag(1/tau) + striu(V^T V)); two threads
    // per row split the inner sums (even / odd k) to shorten the dependent chains.  The loops are warp-uniform (shuffles).
    for (int base = 0; base < 2 * R; base += nt) {
        const int item = base + tid;
        const bool act = item < 2 * R;
        const int rl = act ? (item >> 1) : 0, half = item & 1;
        const int row = r0 + rl;
        T* x = Ps + rl * ld;
        const int khi = min(kmax - 1, row);
        for (int c = kmax - 1; c >= 0; --c) {
            T s0 = (T)0, s1 = (T)0;
            if (act && c <= khi) {
                int k = c + 1 + half;
                for (; k + 2 <= khi; k += 4) { s0 += x[k] * Gm[c * b + k]; s1 += x[k + 2] * Gm[c * b + k + 2]; }
                for (; k <= khi; k += 2) s0 += x[k] * Gm[c * b + k];
            }
            T s = s0 + s1;
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (act && half == 0) {
                const T vc = (row == c) ? (T)1 : x[c];
                x[c] = (c <= khi) ? (-vc - s) * taus[c] : (T)0;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (!kTrans) {
        if (in2d)
            for (int rl = tyy; rl < R; rl += rgroups) V2[(size_t)(r0 + rl) * b + tx] = Ps[rl * ld + tx];
    } else {
        for (int c = w; c < b; c += kWarps)
            for (int rl = lane; rl < R; rl += 32) V2[(size_t)c * m + (r0 + rl)] = Ps[rl * ld + c];
    }
    BLK_TICK(5);
    if (SVDB_PANEL_TIMING && blockIdx.x == 0 && threadIdx.x == 0) g_blk_dbg[9] += 1;
}

inline size_t blk_smem_bytes(int rows, int b, size_t esz, const BlkShape& sh, size_t* stage_out) {
    size_t stage = (size_t)rows * (b + 1) * esz;
    if (sh.CS * sh.NC > 1 && sh.exch_bytes > stage) stage = sh.exch_bytes;
    stage = (stage + 15) & ~(size_t)15;
    *stage_out = stage;
    const size_t CB = (size_t)kC * b;
    return stage + ((size_t)b * b + b + kWarps * CB + CB + 2 * CB + CB + CB) * esz + 64;
}

template <typename T, bool kTrans, int RPT, int CPL>
int launch_blk(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream) {
    constexpr int ROWS = RPT * kWarps;
    const int G0 = (m + ROWS - 1) / ROWS;
    int CS = 1, NC = 1;
    if (G0 > 1) {
        if (!c->cluster_ok) return 1;
        CS = G0 <= c->cluster_ok ? G0 : c->cluster_ok;
        NC = (G0 + CS - 1) / CS;
        if (NC > kMaxNC || NC * CS > c->num_sms) return 1;
        // beside a running stage-2 kernel only single-cluster launches are allowed (co-residency, stage1_panel_reg.cu)
        if (NC > 1 && c->overlap_safe) return 1;
    }
    const BlkShape sh = blk_shape(b, CS, NC);
    size_t stage = 0;
    const size_t smem = blk_smem_bytes(ROWS, b, sizeof(T), sh, &stage);
    if (smem > 227 * 1024) return 1;
    auto kern = panel_blk_kernel<T, kTrans, RPT, CPL>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * CS);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (NC > 1) {
        // all clusters of a launch wait for each other's words: they must be co-resident
        static int cache_key[16][3], cache_val[16], cache_n = 0;
        const int key[3] = {CS, (int)smem, (int)(sizeof(T) * 1000 + RPT * 10 + CPL + (kTrans ? 5 : 0))};
        int max_clusters = -1;
        for (int q = 0; q < cache_n; ++q)
            if (cache_key[q][0] == key[0] && cache_key[q][1] == key[1] && cache_key[q][2] == key[2]) max_clusters = cache_val[q];
        if (max_clusters < 0) {
            max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
            if (cache_n < 16) { cache_key[cache_n][0] = key[0]; cache_key[cache_n][1] = key[1]; cache_key[cache_n][2] = key[2]; cache_val[cache_n] = max_clusters; ++cache_n; }
        }
        if (NC > max_clusters) return 1;
    }
    char* gbuf = reinterpret_cast<char*>(c->red2);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, lda, m, b, V, V2, gbuf, NC, ++c->panel_epoch, stage);
    if (e != cudaSuccess) { cudaGetLastError(); return 1; }
    c->launches++;
    return 0;
}

}  // namespace

int panel_blk_debug_read(long long* out16) {
    long long z[16] = {};
    if (cudaMemcpyFromSymbol(out16, g_blk_dbg, sizeof(z)) != cudaSuccess) return 1;
    cudaMemcpyToSymbol(g_blk_dbg, z, sizeof(z));
    return 0;
}

// returns 0 when it ran, 1 when the shape is outside this kernel's range (caller falls back to the per-column kernels)
template <typename T, bool kTrans>
int launch_panel_blk(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream) {
    if (b < kC || b > 64 || b % kC != 0 || m < b || !c->red2) return 1;
    const int maxcs = c->cluster_ok > 0 ? c->cluster_ok : 1;
    // rows per CTA = 8 * RPT: one CTA up to 128 rows, then the smallest slice that keeps the panel inside one cluster
    auto fits = [&](int rpt) { return (m + 8 * rpt - 1) / (8 * rpt) <= (m <= 128 ? 1 : maxcs); };
    if (b <= 32) {
        if (fits(8)) return launch_blk<T, kTrans, 8, 1>(c, a, lda, m, b, V, V2, stream);
        if (fits(16)) return launch_blk<T, kTrans, 16, 1>(c, a, lda, m, b, V, V2, stream);
        if (sizeof(T) == 8 || fits(32)) return launch_blk<T, kTrans, 32, 1>(c, a, lda, m, b, V, V2, stream);
        return launch_blk<float, kTrans, 64, 1>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream);
    }
    if (fits(8)) return launch_blk<T, kTrans, 8, 2>(c, a, lda, m, b, V, V2, stream);
    if (sizeof(T) == 8 || fits(16)) return launch_blk<T, kTrans, 16, 2>(c, a, lda, m, b, V, V2, stream);
    if (fits(32)) return launch_blk<float, kTrans, 32, 2>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream);
    return launch_blk<float, kTrans, 64, 2>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream);
}
template int launch_panel_blk<float, false>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t);
template int launch_panel_blk<float, true>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t);
template int launch_panel_blk<double, false>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t);
template int launch_panel_blk<double, true>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t);

}  // namespace svdb200
