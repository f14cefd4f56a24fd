// Blocked panel factorisation, slice resident in shared memory: ONE exchange per sub-panel of C = 8 columns instead of one
// per column.
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

__device__ __forceinline__ double guard_ratio(bool is_float) { return is_float ? 0.25 : 1.0 / 64.0; }

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

// Shared-memory plan (bytes).  The slice of the panel (rows x (b+1)) stays in shared memory for the whole factorisation
// (an earlier version kept it in registers: its fully unrolled row loops made the kernel instruction-fetch bound --
// ncu: 4 "no instruction" stall cycles per issued instruction -- and needed one instantiation per slice height).
struct BlkSmem {
    size_t ps, exch_off, psum_off, rest_off, total;
};
__host__ __device__ inline BlkSmem blk_smem(int rows, int b, size_t esz, const BlkShape& sh, int dbl) {
    const size_t CB = (size_t)kC * b;
    BlkSmem m;
    m.ps = (((size_t)rows * (b + 1) * esz) + 15) & ~(size_t)15;
    const size_t exch = (sh.CS * sh.NC > 1) ? (dbl ? sh.exch_bytes : sh.exch_bytes / 2) : 0;
    m.exch_off = m.ps;
    m.psum_off = m.exch_off + ((exch + 15) & ~(size_t)15);
    m.rest_off = m.psum_off + 4 * CB * 8;                  // 4 buffers: warps w and w+4 share one (two write phases)
    // doubles: red (L) + topS (CB) + Cs (CB) + srow (b) + fas (b);
    // elements: Tt (b*b) + Gp (2*b*8) + Zs (b*8) + T22s (64) + taus (b) + Kc + TopF (CB each); ctl
    m.total = m.rest_off + (2 * CB + CB + CB + 2 * (size_t)b) * 8 + ((size_t)b * b + 24 * (size_t)b + 64 + b + 2 * CB) * esz + 64;
    return m;
}

template <typename T, bool kTrans, int CPL>
__global__ void __launch_bounds__(kThreads, 1)
panel_blk_kernel(T* __restrict__ A, size_t lda, int m, int b, T* __restrict__ V, T* __restrict__ V2, char* __restrict__ gbuf, int NC,
                 unsigned epoch, int ROWS, int dbl, const int* __restrict__ run_if) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (run_if != nullptr && *run_if == 0) return;       // fallback launch behind the Cholesky-QR panel: only when it gave up
    constexpr bool kFloat = sizeof(T) == 4;
    constexpr int C = kC;
    const int tid = threadIdx.x, nt = kThreads, lane = tid & 31, w = tid >> 5;
    const int G = gridDim.x, g = blockIdx.x;
    const int CS = G / NC, cl = g / CS, crank = g - cl * CS;
    const BlkShape sh = blk_shape(b, CS, NC);
    const BlkSmem plan = blk_smem(ROWS, b, sizeof(T), sh, dbl);
    const int L = sh.L, SL = sh.SL, CB = C * b;
    const int r0 = g * ROWS;
    const int R = max(0, min(ROWS, m - r0));
    const int ld = b + 1;
    // ---- shared memory ----------------------------------------------------------------------------------
    T* Ps = reinterpret_cast<T*>(smem_raw);               // ROWS x ld : the slice, resident for the whole factorisation
    unsigned char* exch = smem_raw + plan.exch_off;       // exchange words
    double* psum = reinterpret_cast<double*>(smem_raw + plan.psum_off);   // 4 x CB per-warp-pair dot products (reused for slice partials)
    double* red = reinterpret_cast<double*>(smem_raw + plan.rest_off);    // L : reduced D (CB), then Top (CB) -- the algebra works on Top in place
    double* topS = red + L;                               // CB : this CTA's rows among the C top rows (zeros elsewhere)
    double* Cs = topS + CB;                               // CB : coefficients Cm[p][col] of the algebra
    double* srow = Cs + CB;                               // b  : row i of S (dots of the pivot column's lo part with every column)
    double* fas = srow + b;                               // b  : f_k * alpha of the current step
    T* Tt = reinterpret_cast<T*>(fas + b);                // b x b : Tt[k*b + c] = T[c][k] (compact-WY T, upper triangular)
    T* Gp = Tt + b * b;                                   // 2 x (b x 8): v_col^T v_(j+q) of the current / previous sub-panel
    T* Zs = Gp + 2 * b * 8;                               // b x 8
    T* T22s = Zs + b * 8;                                 // 8 x 8
    T* taus = T22s + 64;                                  // b
    T* Kc = taus + b;                                     // CB : pass coefficients per column
    T* TopF = Kc + CB;                                    // CB : final top rows of the sub-panel
    int* ctl = reinterpret_cast<int*>(TopF + CB);         // [0] = C_eff
    const unsigned l1_base = smem_addr(exch);
    const int npar = dbl ? 2 : 1;
    const unsigned in_bytes = (unsigned)(npar * CS * SL * 16);
    auto in_off = [&](int par, int src, int e) { return (unsigned)(((par * CS + src) * SL + e) * 16); };
    auto out_off = [&](int par, int idx) { return in_bytes + (unsigned)((par * L + idx) * 16); };

    int cu[CPL];
    bool valid[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) { cu[u] = lane + 32 * u; valid[u] = cu[u] < b; }
    const int tx = tid % b, tyy = tid / b, rgroups = max(1, nt / b);
    const bool in2d = tyy < rgroups;

    // ---- load the slice ------------------------------------------------------------------------------------------
    if (!kTrans) {
        if (in2d)
            for (int rl = tyy; rl < R; rl += rgroups) Ps[rl * ld + tx] = A[(size_t)(r0 + rl) * lda + tx];
    } else {
        for (int c = w; c < b; c += kWarps)
            for (int rl = lane; rl < R; rl += 32) Ps[rl * ld + c] = A[(size_t)c * lda + (r0 + rl)];
    }
    for (int e = tid; e < b * b; e += nt) Tt[e] = (T)0;
    for (int e = tid; e < CB; e += nt) topS[e] = 0.0;
    if (G > 1) {
        unsigned* z = reinterpret_cast<unsigned*>(exch);
        for (int e = tid; e < (int)((npar * ((size_t)CS * SL + L) * 16) / 4); e += nt) z[e] = 0u;
        cg::this_cluster().sync();                        // every peer's word buffers are cleared before the first push
    } else {
        __syncthreads();
    }

    const int kmax = b;                                   // m >= b (checked by the launcher)
    double acc[C][CPL];
    // One pass over the local rows (warp w: rows w, w+8, ...; lane = column).  apply: update with the coefficients of the
    // sub-panel that started at column j and processed ceff columns; then (always) dot products and top rows for the
    // sub-panel starting at jn.  The values of a row that every lane needs (the sub-panel's C columns) are broadcast loads
    // from the row itself.  Float: the dot products of 8 rows are summed in float, the chunks in double (the norms of
    // later columns are differences of these sums).
    auto pass = [&](bool apply, int j, int ceff, int jn) {
        T K[C][CPL];
        bool fin[CPL];
#pragma unroll
        for (int u = 0; u < CPL; ++u) {
            fin[u] = apply && cu[u] >= j && cu[u] < j + ceff;
#pragma unroll
            for (int p = 0; p < C; ++p) K[p][u] = (apply && valid[u]) ? Kc[p * b + cu[u]] : (T)0;
        }
        T accf[C][CPL];
#pragma unroll
        for (int p = 0; p < C; ++p)
#pragma unroll
            for (int u = 0; u < CPL; ++u) { acc[p][u] = 0.0; accf[p][u] = (T)0; }
        const bool more = jn < kmax;
        int cnt = 0;
#pragma unroll 1
        for (int rl = w; rl < R; rl += kWarps) {
            T* row = Ps + rl * ld;
            const int grow = r0 + rl;
            T av[CPL];
#pragma unroll
            for (int u = 0; u < CPL; ++u) av[u] = valid[u] ? row[cu[u]] : (T)0;
            if (apply && grow >= j) {                      // warp-uniform
                if (grow < j + C) {                        // one of the C top rows: final values from the algebra
#pragma unroll
                    for (int u = 0; u < CPL; ++u) if (valid[u]) av[u] = TopF[(grow - j) * b + cu[u]];
                } else {
                    T x[C];
#pragma unroll
                    for (int p = 0; p < C; ++p) x[p] = row[min(j + p, b - 1)];
#pragma unroll
                    for (int u = 0; u < CPL; ++u) {
                        T v0 = fin[u] ? (T)0 : av[u], v1 = (T)0;
#pragma unroll
                        for (int p = 0; p < C; p += 2) { v0 += x[p] * K[p][u]; v1 += x[p + 1] * K[p + 1][u]; }
                        av[u] = v0 + v1;
                    }
                }
                __syncwarp();                              // every lane has read the row before it is rewritten
#pragma unroll
                for (int u = 0; u < CPL; ++u) if (valid[u]) row[cu[u]] = av[u];
                __syncwarp();
            }
            if (more && grow >= jn) {
                if (grow < jn + C) {
#pragma unroll
                    for (int u = 0; u < CPL; ++u) if (valid[u]) topS[(grow - jn) * b + cu[u]] = (double)av[u];
                } else {
#pragma unroll
                    for (int p = 0; p < C; ++p) {
                        const T yv = (jn + p < kmax) ? row[jn + p] : (T)0;
#pragma unroll
                        for (int u = 0; u < CPL; ++u) accf[p][u] += yv * av[u];
                    }
                    if (!kFloat || ((++cnt) & 7) == 0) {
#pragma unroll
                        for (int p = 0; p < C; ++p)
#pragma unroll
                            for (int u = 0; u < CPL; ++u) { acc[p][u] += (double)accf[p][u]; accf[p][u] = (T)0; }
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < C; ++p)
#pragma unroll
            for (int u = 0; u < CPL; ++u) acc[p][u] += (double)accf[p][u];
    };

    // compact-WY T of the sub-panel [jp, jp + ce) from its block of Gram entries Gq (b x 8) and the T of the columns left of
    // it:  T22 = (D22 + striu(G22))^-1 by the column recurrence,  T12 = -T11 (G12 T22).  Run by the `nthr` threads with
    // index t (whole warps) that are synchronised through named barrier 2.
    auto t_update = [&](int jp, int ce, const T* Gq, int t, int nthr) {
        if (ce <= 0) return;
        if (t < 32) {                                      // first warp of the group: lane r = row r of T22 (lane-local recurrence)
            if (t < ce) {
                T row[C];
#pragma unroll
                for (int q = 0; q < C; ++q) {
                    T v = (T)0;
                    if (q < ce && q >= t) {
                        const T tq = taus[jp + q];
                        if (q == t) v = tq;
                        else {
                            T sacc = (T)0;
#pragma unroll
                            for (int cc = 0; cc < C; ++cc) if (cc >= t && cc < q) sacc += row[cc] * Gq[(jp + cc) * 8 + q];
                            v = -tq * sacc;
                        }
                    }
                    row[q] = v;
                }
#pragma unroll
                for (int q = 0; q < C; ++q) {
                    T22s[t * 8 + q] = row[q];
                    if (q < ce) Tt[(jp + q) * b + (jp + t)] = row[q];
                }
            } else if (t < C) {
#pragma unroll
                for (int q = 0; q < C; ++q) T22s[t * 8 + q] = (T)0;
            }
        }
        asm volatile("bar.sync 2, %0;" ::"r"(nthr) : "memory");
        for (int o = t; o < jp * ce; o += nthr) {          // Z = G12 T22
            const int cc = o / ce, q = o - cc * ce;
            T z = (T)0;
            for (int p = 0; p <= q; ++p) z += Gq[cc * 8 + p] * T22s[p * 8 + q];
            Zs[cc * 8 + q] = z;
        }
        asm volatile("bar.sync 2, %0;" ::"r"(nthr) : "memory");
        for (int o = t; o < jp * ce; o += nthr) {          // T12 = -T11 Z   (T11 upper triangular; Tt is T transposed)
            const int q = o / jp, r = o - q * jp;
            T z0 = (T)0, z1 = (T)0;
            int cc = r;
            for (; cc + 1 < jp; cc += 2) { z0 += Tt[cc * b + r] * Zs[cc * 8 + q]; z1 += Tt[(cc + 1) * b + r] * Zs[(cc + 1) * 8 + q]; }
            if (cc < jp) z0 += Tt[cc * b + r] * Zs[cc * 8 + q];
            Tt[(jp + q) * b + r] = -(z0 + z1);
        }
    };

    long long tick = SVDB_PANEL_TIMING ? clock64() : 0;
    (void)tick;
    int j = 0, jprev = 0, ceprev = 0, ceff = 0;
    unsigned round = 0;
    bool first = true;
    while (true) {
        // update with the finished sub-panel (none the first time) + dot products of the next one: ONE call site, so the
        // row loop exists once in the instruction stream
        pass(!first, j, ceff, j + ceff);
        if (first) { BLK_TICK(0); } else { BLK_TICK(4); if (SVDB_PANEL_TIMING && blockIdx.x == 0 && threadIdx.x == 0) g_blk_dbg[8] += 1; }
        if (!first) { jprev = j; ceprev = ceff; j += ceff; round += 1; }
        first = false;
        if (j >= kmax) break;
        // ---- publish the per-warp dot products: warps 0-3 write, warps 4-7 add (4 buffers instead of 8) ----------------
        if (w < 4) {
#pragma unroll
            for (int p = 0; p < C; ++p)
#pragma unroll
                for (int u = 0; u < CPL; ++u)
                    if (valid[u]) psum[w * CB + p * b + cu[u]] = acc[p][u];
        }
        __syncthreads();
        if (w >= 4) {
#pragma unroll
            for (int p = 0; p < C; ++p)
#pragma unroll
                for (int u = 0; u < CPL; ++u)
                    if (valid[u]) psum[(w - 4) * CB + p * b + cu[u]] += acc[p][u];
        }
        __syncthreads();
        BLK_TICK(1);
        // ---- deterministic all-reduce of [D | Top] -----------------------------------------------------------------
        auto local_val = [&](int idx) -> double {
            if (idx >= CB) return topS[idx - CB];
            return (psum[idx] + psum[CB + idx]) + (psum[2 * CB + idx] + psum[3 * CB + idx]);
        };
        if (G == 1) {
            for (int idx = tid; idx < L; idx += nt) red[idx] = local_val(idx);
            __syncthreads();
        } else {
            const unsigned seq = epoch * 128u + round + 1u;
            const int par = dbl ? (int)(round & 1u) : 0;
            // hop 1 (reduce-scatter): slice q of the vector goes to the CTA of rank q in the cluster
            for (int idx = tid; idx < L; idx += nt) {
                const double v = local_val(idx);
                const int q = idx / SL, e = idx - q * SL;
                LLSmem<double>::push(map_to_rank(l1_base + in_off(par, crank, e), (unsigned)q), v, seq);
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
                double s = 0.0;
                if (e < mine) {
                    const int q0 = part * chunk, q1 = min(CS, q0 + chunk);
                    for (int q = q0; q < q1; ++q) {
                        double v;
                        while (!LLSmem<double>::try_load(l1_base + in_off(par, q, e), seq, v)) sg.tick();
                        s += v;
                    }
                }
                psum[part * SL + e] = s;
            }
            __syncthreads();
            for (int e = tid; e < mine; e += nt) {
                double s = psum[e];
                for (int part = 1; part < parts; ++part) s += psum[part * SL + e];
                const int idx = crank * SL + e;
                if (NC > 1) {
                    // level 2: the owners of the same slice in the other clusters exchange their sums through L2
                    char* gs = gbuf + (size_t)(round & 1u) * NC * L * 16;
                    LLWord<double>::store(gs + ((size_t)cl * L + idx) * 16, s, seq);
                    double v[kMaxNC];
                    unsigned pending = (1u << NC) - 1u;
                    while (pending != 0u) {
#pragma unroll
                        for (int q = 0; q < kMaxNC; ++q)
                            if ((pending >> q) & 1u) {
                                if (LLWord<double>::try_load(gs + ((size_t)q * L + idx) * 16, seq, v[q])) pending &= ~(1u << q);
                            }
                        if (pending != 0u) { sg.tick(); __nanosleep(40); }
                    }
                    s = v[0];
#pragma unroll
                    for (int q = 1; q < kMaxNC; ++q) if (q < NC) s += v[q];
                }
                // hop 2 (all-gather): the finished entry goes to every CTA of the cluster
                for (int q = 0; q < CS; ++q) LLSmem<double>::push(map_to_rank(l1_base + out_off(par, idx), (unsigned)q), s, seq);
            }
            for (int idx = tid; idx < L; idx += nt) {
                double v;
                while (!LLSmem<double>::try_load(l1_base + out_off(par, idx), seq, v)) sg.tick();
                red[idx] = v;
            }
            __syncthreads();
        }
        BLK_TICK(2);
        for (int e = tid; e < CB; e += nt) topS[e] = 0.0;          // next round's top rows: only their owner writes them
        T* Gcur = Gp + (round & 1u) * (b * 8);
        // ---- the C Householder steps on the exchanged data (warp 0; lane = column, state in shared memory) ------------
        if (w == 0) {
            // Downdating form: S[p][k] = (current column j+p, lo part)^T (current column k, lo part) is carried along and
            // corrected after every reflector (rank-one formulas, no dependent chains); the norm and the dot products of
            // the next pivot column are then single entries of S.  Top rows and coefficients live in shared memory.
            double* Ts = red + CB;                         // Top rows, updated in place
            const double* D = red;
            double S[C][CPL], d0[C];
#pragma unroll
            for (int p = 0; p < C; ++p) {
                d0[p] = D[p * b + min(j + p, kmax - 1)];   // original ||lo part||^2 of the p-th column of the sub-panel
#pragma unroll
                for (int u = 0; u < CPL; ++u) {
                    S[p][u] = valid[u] ? D[p * b + cu[u]] : 0.0;
                    if (valid[u]) Cs[p * b + cu[u]] = 0.0;
                }
            }
            for (int e = lane; e < b * 8; e += 32) Gcur[e] = (T)0;
#pragma unroll
            for (int u = 0; u < CPL; ++u) if (valid[u]) srow[cu[u]] = S[0][u];
            __syncwarp();
            int done = 0;
            bool active = true;
            const double guard = guard_ratio(kFloat);
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const int ji = j + i;
                active = active && (ji < kmax);
                if (active) {                              // warp-uniform
                    const double sji = srow[ji];
                    if (i > 0 && !(sji >= guard * d0[i])) {
                        active = false;                    // too much cancellation: end the sub-panel here
                    } else {
                        double tj[C], mv[C];
#pragma unroll
                        for (int p = 0; p < C; ++p) {
                            tj[p] = (p >= i) ? Ts[p * b + ji] : 0.0;
                            mv[p] = (p <= i) ? (((p == i) ? 1.0 : 0.0) - Cs[p * b + ji]) : 0.0;
                        }
                        double t2a = 0.0, t2b = 0.0;
#pragma unroll
                        for (int t = i; t < C; t += 2) { t2a += tj[t] * tj[t]; if (t + 1 < C) t2b += tj[t + 1] * tj[t + 1]; }
                        const double x0 = tj[i];
                        const double normsq = sji + (t2a + t2b);
                        const double rn = rsqrt(normsq);
                        const double nrm = normsq * rn;
                        const double sgn = -copysign(1.0, x0);
                        const double u1 = x0 - sgn * nrm;
                        const double alpha = __drcp_rn(u1);
                        const double tau = -sgn * u1 * rn;
                        const double beta = sgn * nrm;                // R_jj = -sign(x0) ||x||
                        double fav[CPL], xav[CPL];
#pragma unroll
                        for (int u = 0; u < CPL; ++u) {
                            fav[u] = 0.0; xav[u] = S[i][u];
                            if (!valid[u]) continue;
                            const int col = cu[u];
                            double dva = 0.0, dvb = 0.0;
#pragma unroll
                            for (int t = i + 1; t < C; t += 2) {
                                dva += tj[t] * Ts[t * b + col];
                                if (t + 1 < C) dvb += tj[t + 1] * Ts[(t + 1) * b + col];
                            }
                            const double dot = Ts[i * b + col] + alpha * ((dva + dvb) + xav[u]);
                            if (col > ji) {
                                const double f = tau * dot, fa = f * alpha;
                                fav[u] = fa;
                                Ts[i * b + col] -= f;
#pragma unroll
                                for (int p = 0; p < C; ++p) {
                                    if (p > i) Ts[p * b + col] -= fa * tj[p];
                                    else Cs[p * b + col] += fa * mv[p];
                                }
                            } else if (col == ji) {
                                Ts[i * b + col] = beta;
#pragma unroll
                                for (int p = 0; p < C; ++p) {
                                    if (p > i) Ts[p * b + col] = alpha * tj[p];
                                    else Cs[p * b + col] = ((p == i) ? 1.0 : 0.0) - alpha * mv[p];
                                }
                                taus[ji] = (T)tau;
                            } else {
                                Gcur[col * 8 + i] = (T)dot;   // v_col^T v_ji
                            }
                            fas[col] = fav[u];
                        }
                        __syncwarp();
                        if (i + 1 < C) {
#pragma unroll
                            for (int p = i + 1; p < C; ++p) {
                                const int q = min(j + p, kmax - 1);
                                const double faq = fas[q], xq = srow[q];
#pragma unroll
                                for (int u = 0; u < CPL; ++u) {
                                    if (cu[u] == ji) S[p][u] = alpha * (xq - faq * sji);
                                    else S[p][u] = S[p][u] - fav[u] * xq - faq * (xav[u] - fav[u] * sji);
                                }
                            }
                            __syncwarp();                  // everybody has read row i of S and the f's
#pragma unroll
                            for (int u = 0; u < CPL; ++u) if (valid[u]) srow[cu[u]] = S[i + 1][u];
                            __syncwarp();
                        }
                        done = i + 1;
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < CPL; ++u)
                if (valid[u]) {
                    const int col = cu[u];
                    const bool f = col >= j && col < j + done;
#pragma unroll
                    for (int p = 0; p < C; ++p) {
                        const double cv = Cs[p * b + col];
                        Kc[p * b + col] = (T)(f ? (((p == col - j) ? 1.0 : 0.0) - cv) : -cv);
                        TopF[p * b + col] = (T)Ts[p * b + col];
                    }
                }
            if (lane == 0) ctl[0] = done;
        } else {
            // meanwhile warps 1-7: the block column of T that belongs to the PREVIOUS sub-panel
            t_update(jprev, ceprev, Gp + ((round + 1u) & 1u) * (b * 8), tid - 32, nt - 32);
        }
        __syncthreads();
        BLK_TICK(3);
        ceff = ctl[0];
        if (!dbl && G > 1) cg::this_cluster().sync();     // single-parity exchange buffers: everybody is done with this round's words
    }
    __syncthreads();
    if (w != 0) t_update(jprev, ceprev, Gp + ((round + 1u) & 1u) * (b * 8), tid - 32, nt - 32);   // T of the last sub-panel

    // ---- epilogue --------------------------------------------------------------------------------------------------
    if (G > 1) cg::this_cluster().sync();                 // nobody pushes words into the staging area any more
    else __syncthreads();
    BLK_TICK(10);
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
    // V2 = V S^T with S = -T:  V2[r][c] = -sum_{k = c .. min(r, b-1)} V[r][k] T[c][k].  One thread per row: the row of V
    // sits in Ps (odd leading dimension: conflict-free), every T entry is a broadcast load shared by the 32 rows of a warp;
    // no dependent chain (the per-column kernels solve a triangular system per row here).
    __syncthreads();
    BLK_TICK(11);
    for (int rl = tid; rl < R; rl += nt) {                // explicit unit diagonal, zeros above it (rows of the top block only)
        const int row = r0 + rl;
        if (row < b) {
            T* x = Ps + rl * ld;
            x[row] = (T)1;
            for (int k = row + 1; k < b; ++k) x[k] = (T)0;
        }
    }
    __syncthreads();
    {
        // tpr threads per row share its b/8 column chunks (short slices leave most threads idle otherwise); all outputs of
        // a thread stay in registers until every thread of the row has read the row
        const int tpr = (4 * R <= nt) ? 4 : ((2 * R <= nt) ? 2 : 1);
        const int nch = b / 8;
        for (int base = 0; base < R * tpr; base += nt) {
            const int item = base + tid;
            const bool act = item < R * tpr;
            const int rl = act ? item / tpr : 0, sub = item % tpr;
            const int row = r0 + rl, khi = min(b - 1, row);
            T* x = Ps + rl * ld;
            T o[8][8];
#pragma unroll
            for (int qc = 0; qc < 8; ++qc) {
                const int c0 = (sub + qc * tpr) * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) o[qc][q] = (T)0;
                if (act && sub + qc * tpr < nch && c0 <= khi) {
                    for (int k = c0; k <= khi; ++k) {
                        const T xv = x[k];
                        const T* tk = Tt + k * b + c0;         // T[c0 + q][k], zero for c0 + q > k  (16-byte aligned)
                        T tv[8];
                        if constexpr (kFloat) {
                            const float4 v0 = *reinterpret_cast<const float4*>(tk), v1 = *reinterpret_cast<const float4*>(tk + 4);
                            tv[0] = v0.x; tv[1] = v0.y; tv[2] = v0.z; tv[3] = v0.w; tv[4] = v1.x; tv[5] = v1.y; tv[6] = v1.z; tv[7] = v1.w;
                        } else {
#pragma unroll
                            for (int p2 = 0; p2 < 4; ++p2) {
                                const double2 v = *reinterpret_cast<const double2*>(tk + 2 * p2);
                                tv[2 * p2] = v.x; tv[2 * p2 + 1] = v.y;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) o[qc][q] += xv * tv[q];
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int qc = 0; qc < 8; ++qc) {
                const int c0 = (sub + qc * tpr) * 8;
                if (act && sub + qc * tpr < nch) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[c0 + q] = -o[qc][q];
                }
            }
        }
    }
    __syncthreads();
    BLK_TICK(12);
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

// rows: rows per CTA (multiple of 8).  returns 0 when launched, 1 when this shape cannot run (caller tries another)
template <typename T, bool kTrans, int CPL>
int launch_blk(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream, int rows) {
    const int G0 = (m + rows - 1) / rows;
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
    int dbl = 1;
    size_t smem = blk_smem(rows, b, sizeof(T), sh, dbl).total;
    if (smem > 227 * 1024) { dbl = 0; smem = blk_smem(rows, b, sizeof(T), sh, dbl).total; }   // single-parity exchange words + a cluster barrier per round
    if (smem > 227 * 1024) return 1;
    auto kern = panel_blk_kernel<T, kTrans, CPL>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
        static int cache_key[32][3], cache_val[32], cache_n = 0;
        const int key[3] = {CS, (int)smem, (int)(sizeof(T) * 100 + CPL * 10 + (kTrans ? 5 : 0))};
        int max_clusters = -1;
        for (int q = 0; q < cache_n; ++q)
            if (cache_key[q][0] == key[0] && cache_key[q][1] == key[1] && cache_key[q][2] == key[2]) max_clusters = cache_val[q];
        if (max_clusters < 0) {
            max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
            if (cache_n < 32) { cache_key[cache_n][0] = key[0]; cache_key[cache_n][1] = key[1]; cache_key[cache_n][2] = key[2]; cache_val[cache_n] = max_clusters; ++cache_n; }
        }
        if (NC > max_clusters) return 1;
    }
    char* gbuf = reinterpret_cast<char*>(c->red2);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, lda, m, b, V, V2, gbuf, NC, ++c->panel_epoch, rows, dbl, c->panel_run_if);
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
    // Rows per CTA: one CTA up to 128 rows; then 64 .. 256 rows with the panel inside ONE cluster (<= 16 CTAs) when that
    // is possible; taller panels run as several clusters of 256-row CTAs, the tallest (more CTAs than the GPU can keep
    // co-resident) with up to 512 rows per CTA.  Shapes that fit neither shared memory nor the GPU return 1.
    const int maxcs = c->cluster_ok > 0 ? c->cluster_ok : 1;
    auto run = [&](int rows) -> int {
        if (b <= 32) return launch_blk<T, kTrans, 1>(c, a, lda, m, b, V, V2, stream, rows);
        return launch_blk<T, kTrans, 2>(c, a, lda, m, b, V, V2, stream, rows);
    };
    if (m <= 128) return run((m + 7) / 8 * 8);
    for (int rows = 64; rows <= 512; rows += 32)
        if ((m + rows - 1) / rows <= maxcs) { const int st = run(rows); if (st != 1) return st; }
    for (int rows = 64; rows <= 512; rows += 32) {
        const int G0 = (m + rows - 1) / rows;
        if (G0 <= 8 * maxcs && G0 <= c->num_sms - 20) { const int st = run(rows); if (st != 1) return st; }
    }
    return 1;
}
template int launch_panel_blk<float, false>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t);
template int launch_panel_blk<float, true>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t);
template int launch_panel_blk<double, false>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t);
template int launch_panel_blk<double, true>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t);

}  // namespace svdb200
