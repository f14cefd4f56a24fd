// Register-resident panel factorisation (band <= 64, panels up to 148 * 256 rows): the fast path of
// the stage-1 panel QR/LQ, same algorithm and outputs as panel_factor_kernel in stage1_panel.cu
// (which stays as the general fallback), but
//   * each CTA keeps its slice of the panel in REGISTERS for the whole column loop: a warp owns the
//     rows w, w+8, w+16, ... of the slice, a lane owns column `lane` (and `lane+32` for band 64);
//   * the pivot column is broadcast inside the warp with shuffles, so the rank-1 update of column j
//     and the dot products needed for column j+1 are ONE fused pass over the registers
//     (the shared-memory version makes two passes per column and was bound by their latency);
//   * per column: one cross-warp reduction, one all-reduce across CTAs, scalars recomputed per thread.  The
//     all-reduce is two-level: DSMEM + barrier.cluster inside a cluster of <= 16 CTAs; panels taller than one
//     cluster run as several clusters whose per-cluster vectors cross L2 once, stamped with a release flag
//     (no grid barrier, 8 x b values to read instead of G x b).  The L2-only grid transport remains as fallback.
// Shared memory is only used for the reductions, the transposed load/store of LQ row panels and the
// epilogue (V, V2 = V S^T by back substitution with T^-1 = D + striu(V^T V), R/L write-back).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace svdb200 {
namespace {

constexpr int kThreads = 256, kWarps = 8;     // two CTAs per SM: eight 16-CTA clusters (128 CTAs) can be co-resident
constexpr int kMaxClusters = 18;      // clusters per panel launch (two-level all-reduce)

#ifndef SVDB_PANEL_TIMING
#define SVDB_PANEL_TIMING 0
#endif
// per-phase cycle counters of CTA 0 / thread 0 (debug builds only: -DSVDB_PANEL_TIMING=1)
__device__ long long g_panel_dbg[16];
#define PANEL_TICK(k)                                                        \
    do {                                                                     \
        if (SVDB_PANEL_TIMING && blockIdx.x == 0 && threadIdx.x == 0) {      \
            long long _t = clock64();                                        \
            g_panel_dbg[k] += _t - tick;                                     \
            tick = _t;                                                       \
        }                                                                    \
    } while (0)

// Flag-stamped words (the "LL" idea): every 8-byte half carries 32 data bits and the 32-bit sequence number of the
// column it belongs to, so a reader that sees the expected number in both halves has the data -- one L2 round trip,
// no separate flag, no fence.  Entries are 16 bytes for both element types (float uses the first half).
template <typename T> struct LLWord;
template <> struct LLWord<float> {
    static __device__ __forceinline__ void store(void* p, float v, unsigned seq) {
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
    }
    static __device__ __forceinline__ bool try_load(const void* p, unsigned seq, float& v) {
        unsigned a, b;
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
        v = __uint_as_float(a);
        return b == seq;
    }
};
template <> struct LLWord<double> {
    static __device__ __forceinline__ void store(void* p, double v, unsigned seq) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)u), "r"(seq), "r"((unsigned)(u >> 32)), "r"(seq) : "memory");
    }
    static __device__ __forceinline__ bool try_load(const void* p, unsigned seq, double& v) {
        unsigned a, b, c2, d;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d) : "l"(p) : "memory");
        v = __longlong_as_double((long long)(((unsigned long long)c2 << 32) | a));
        return b == seq && d == seq;
    }
};

// The same words in shared memory: a CTA pushes them into a PEER's shared memory (st.shared::cluster on the mapa
// address) and the peer polls its own copy -- the per-column all-reduce inside a cluster needs no barrier.cluster.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned map_to_rank(unsigned local_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
template <typename T> struct LLSmem;
template <> struct LLSmem<float> {
    static __device__ __forceinline__ void push(unsigned remote, float v, unsigned seq) {
        asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(remote), "r"(__float_as_uint(v)), "r"(seq) : "memory");
    }
    static __device__ __forceinline__ bool try_load(unsigned local, unsigned seq, float& v) {
        unsigned a, b;
        asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(local) : "memory");
        v = __uint_as_float(a);
        return b == seq;
    }
};
template <> struct LLSmem<double> {
    static __device__ __forceinline__ void push(unsigned remote, double v, unsigned seq) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"((unsigned)u), "r"(seq), "r"((unsigned)(u >> 32)), "r"(seq) : "memory");
    }
    static __device__ __forceinline__ bool try_load(unsigned local, unsigned seq, double& v) {
        unsigned a, b, c2, d;
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d) : "r"(local) : "memory");
        v = __longlong_as_double((long long)(((unsigned long long)c2 << 32) | a));
        return b == seq && d == seq;
    }
};

template <typename T, int CPL>
__device__ __forceinline__ T pick(const T (&v)[CPL], int u) {
    T r = v[0];
#pragma unroll
    for (int q = 1; q < CPL; ++q) r = (u == q) ? v[q] : r;
    return r;
}

// elements of the staging area: max(ROWS x (b+1), the l1 word buffers: 2 parities x (CS + 1) vectors x b entries x 16 B)
template <typename T>
__host__ __device__ inline size_t stage_elems(int rows, int b, int max_cs) {
    const size_t ps = (size_t)rows * (b + 1);
    const size_t l1 = ((size_t)2 * (max_cs + 1) * b * 16 + sizeof(T) - 1) / sizeof(T) + 4;
    return ps > l1 ? ps : l1;
}

template <typename T, bool kTrans, bool kCluster, int RPT, int CPL>
__global__ void __launch_bounds__(kThreads, (sizeof(T) == 4 && RPT * CPL <= 32) ? 2 : 1)
panel_reg_kernel(T* __restrict__ A, size_t lda, int m, int b, T* __restrict__ V, T* __restrict__ V2, T* __restrict__ red,
                 unsigned* __restrict__ bar, int NC, unsigned epoch, const int* __restrict__ run_if) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (run_if != nullptr && *run_if == 0) return;       // fallback launch behind the Cholesky-QR panel: only when it gave up
    constexpr int ROWS = RPT * kWarps;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5;
    const int G = gridDim.x, g = blockIdx.x;
    // kCluster: NC clusters of CS = G / NC CTAs.  The all-reduce of a column is two-level: DSMEM inside a cluster
    // (barrier.cluster), then -- only when NC > 1 -- one flag-stamped vector per cluster through L2.
    const int CS = kCluster ? G / max(NC, 1) : G;
    const int cl = kCluster ? g / CS : 0, crank = kCluster ? g - cl * CS : g;
    (void)cl; (void)crank;
    const int r0 = g * ROWS;
    const int R = max(0, min(ROWS, m - r0));
    const int ld = b + 1, slot = 2 * b;
    T* lred = reinterpret_cast<T*>(smem_raw);            // 2 * slot
    T* Ps = lred + 2 * slot;                             // ROWS x ld (transposed staging + epilogue); during the column
                                                         // loop the same bytes hold the pushed all-reduce words (l1)
    constexpr int kMaxCS = 16;
    const size_t ps_elems = stage_elems<T>(ROWS, b, kMaxCS);
    T* Gm = Ps + ps_elems;                               // b x b
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
    // l1: pushed all-reduce words, [parity][sender rank 0..CS-1, CS = pivot row][column] x 16 B, aliasing Ps
    const unsigned l1_base = (smem_addr(Ps) + 15u) & ~15u;
    if (kCluster) {
        __syncthreads();                                  // the transposed load above is done with Ps
        unsigned* z = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(Ps) + (l1_base - smem_addr(Ps)));
        for (int e = tid; e < 2 * (kMaxCS + 1) * b * 4; e += nt) z[e] = 0u;
        cg::this_cluster().sync();                        // every peer's buffers are cleared before the first push
    }

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

    long long tick = SVDB_PANEL_TIMING ? clock64() : 0;
    (void)tick;
    for (int j = 0; j < kmax; ++j) {
        PANEL_TICK(0);
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
        PANEL_TICK(1);
        for (int c = tid; c < b; c += nt) {
            T s = psum[c];
#pragma unroll
            for (int q = 1; q < kWarps; ++q) s += psum[q * b + c];
            if (kCluster) mine[c] = s; else st_cg(&mine[c], s);
        }
        if (kCluster) __syncthreads(); else grid_barrier(bar, (unsigned)G, gen);
        PANEL_TICK(2);
        // ---- all-reduce across CTAs in a fixed association -------------------------------------------------
        {
            const int jowner = j / ROWS;                 // CTA that holds global row j
            if (kCluster) {
                // level 1 (inside the cluster), push model: every CTA writes its b partial sums -- and the owner of
                // row j the pivot row -- as flag-stamped words into the shared memory of all CS peers (itself included),
                // then polls its OWN shared memory until the CS + 1 vectors of this column have arrived.  No
                // barrier.cluster, one one-way DSMEM latency.  Parity double buffering is safe: a peer can push column
                // j + 2 only after it has received my column j + 1, which I send after I am done with column j.
                const unsigned seq1 = epoch * 128u + (unsigned)(j + 1);
                const unsigned par_off = (unsigned)(j & 1) * (unsigned)((kMaxCS + 1) * b * 16);
                const bool own_cta = (jowner == g);
                if (in2d) {
                    const int c = tx;
                    const T sv = mine[c];
                    const T pvv = own_cta ? pivslot[c] : (T)0;
                    for (int q = tyy; q < CS; q += rgroups) {
                        const unsigned rbase = map_to_rank(l1_base + par_off, (unsigned)q);
                        LLSmem<T>::push(rbase + (unsigned)((crank * b + c) * 16), sv, seq1);
                        if (own_cta) LLSmem<T>::push(rbase + (unsigned)((kMaxCS * b + c) * 16), pvv, seq1);
                    }
                }
                const int chunk = (CS + rgroups - 1) / rgroups;
                if (in2d) {
                    const int c = tx, part = tyy;
                    const int q0 = part * chunk, q1 = min(CS, q0 + chunk);
                    T s = (T)0;
                    unsigned polls = 0;
                    unsigned long long t0 = 0;
                    for (int q = q0; q < q1; ++q) {
                        T v;
                        while (!LLSmem<T>::try_load(l1_base + par_off + (unsigned)((q * b + c) * 16), seq1, v)) {
                            if ((++polls & 0xffffu) == 0u) {
                                unsigned long long now;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) __trap();
                            }
                        }
                        s += v;
                    }
                    if (part == 0 && jowner / CS == cl) {
                        T v;
                        while (!LLSmem<T>::try_load(l1_base + par_off + (unsigned)((kMaxCS * b + c) * 16), seq1, v)) {
                            if ((++polls & 0xffffu) == 0u) {
                                unsigned long long now;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                if (t0 == 0) t0 = now; else if (now - t0 > 4000000000ull) __trap();
                            }
                        }
                        piv[c] = v;
                    }
                    psum[part * b + c] = s;
                }
                __syncthreads();
                PANEL_TICK(3);
                if (tid < b) {                            // nt >= b: thread c owns column c of the reduced vectors
                    const int c = tid;
                    T s = psum[c];
                    for (int part = 1; part < rgroups; ++part) s += psum[part * b + c];
                    if (NC > 1) {
                        // level 2: rank 0 of every cluster publishes the cluster's vector (and the pivot row when the
                        // cluster owns it) as flag-stamped words; thread c of every CTA collects the NC words of its
                        // column and adds them in cluster order.  Slots alternate with the column parity; a cluster can
                        // only be one column ahead of the slowest one, so a slot is never overwritten while still read.
                        const unsigned seq = epoch * 128u + (unsigned)(j + 1);
                        char* gs = reinterpret_cast<char*>(red) + (size_t)(j & 1) * NC * slot * 16;
                        const int ocl = jowner / CS;
                        if (crank == 0) {
                            LLWord<T>::store(gs + ((size_t)cl * slot + c) * 16, s, seq);
                            if (ocl == cl) LLWord<T>::store(gs + ((size_t)cl * slot + b + c) * 16, piv[c], seq);
                        }
                        PANEL_TICK(4);
                        T v[kMaxClusters], pv = (T)0;
                        unsigned pending = (1u << NC) - 1u;
                        bool pdone = false;
                        unsigned polls = 0;
                        unsigned long long t0 = 0;
                        while (pending != 0u || !pdone) {
                            if ((++polls & 4095u) == 0u) {       // a cluster that never becomes resident must not hang the GPU
                                unsigned long long now;
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                if (t0 == 0) t0 = now;
                                else if (now - t0 > 4000000000ull) __trap();
                            }
#pragma unroll
                            for (int q = 0; q < kMaxClusters; ++q)
                                if ((pending >> q) & 1u) {
                                    if (LLWord<T>::try_load(gs + ((size_t)q * slot + c) * 16, seq, v[q])) pending &= ~(1u << q);
                                }
                            if (!pdone) pdone = LLWord<T>::try_load(gs + ((size_t)ocl * slot + b + c) * 16, seq, pv);
                            if (pending != 0u || !pdone) __nanosleep(100);   // back off: the update kernels beside us are L2 / HBM bound
                        }
                        PANEL_TICK(5);
                        s = v[0];
#pragma unroll
                        for (int q = 1; q < kMaxClusters; ++q) if (q < NC) s += v[q];
                        piv[c] = pv;
                    }
                    zs[c] = s;
                }
            } else {
                const int chunk = (G + rgroups - 1) / rgroups;
                if (in2d) {
                    const int c = tx, part = tyy;
                    const int q0 = part * chunk, q1 = min(G, q0 + chunk);
                    T s = (T)0;
                    const T* buf = red + (size_t)(j & 1) * (G + 1) * slot;
#pragma unroll 8
                    for (int q = q0; q < q1; ++q) s += ld_cg(&buf[(size_t)q * slot + c]);
                    if (part == 0) piv[c] = ld_cg(&buf[(size_t)G * slot + c]);
                    psum[part * b + c] = s;
                }
                __syncthreads();
                for (int c = tid; c < b; c += nt) {
                    T s = psum[c];
                    for (int part = 1; part < rgroups; ++part) s += psum[part * b + c];
                    zs[c] = s;
                }
            }
        }
        __syncthreads();
        PANEL_TICK(6);
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
        // Branch-free: per-lane column predicates become factors (fm = 0 for finished columns), per-row predicates
        // become a zero reflector entry, so every (row, column) is one FMA for the update, one select for the
        // column that becomes v_j, and one FMA for the dot products.
        const int lj = j & 31, uj = j >> 5, ln = (j + 1) & 31, un = (j + 1) >> 5;
        const bool more = (j + 1 < kmax);
        T fm[CPL];
        bool isj[CPL];
#pragma unroll
        for (int u = 0; u < CPL; ++u) { fm[u] = (cu[u] > j) ? fsr[u] : (T)0; isj[u] = (cu[u] == j); acc[u] = (T)0; }
        // The two broadcasts per row (pivot column before the update, next pivot column after it) are issued for ALL
        // rows of the slice before their results are consumed: a shuffle has ~30 cycles of latency and with two warps
        // per scheduler a row-by-row order leaves the pass waiting on it (short-scoreboard stalls, prof_r1_panel4).
        // (double: chunks of 8 rows; it runs one CTA per SM with the full register file)
        if constexpr (true) {
        constexpr int CH = sizeof(T) == 8 ? (RPT < 8 ? RPT : 8) : (RPT < 32 ? RPT : 32);
        T bc[CH];
#pragma unroll
        for (int i0 = 0; i0 < RPT; i0 += CH) {
#pragma unroll
            for (int ii = 0; ii < CH; ++ii) bc[ii] = __shfl_sync(0xffffffffu, pick<T, CPL>(a[i0 + ii], uj), lj);
#pragma unroll
            for (int ii = 0; ii < CH; ++ii) {
                const int i = i0 + ii;
                const int rl = w + kWarps * i;
                const int grow = r0 + rl;
                const bool act = (rl < R) && (grow >= j);
                const T wv = act ? ((grow == j) ? (T)1 : bc[ii] * alpha) : (T)0;
                const T cj = (grow == j) ? beta : wv;
#pragma unroll
                for (int u = 0; u < CPL; ++u) {
                    const T nv = a[i][u] - wv * fm[u];
                    a[i][u] = (isj[u] && act) ? cj : nv;
                }
            }
        }
        if (more) {
#pragma unroll
            for (int i0 = 0; i0 < RPT; i0 += CH) {
#pragma unroll
                for (int ii = 0; ii < CH; ++ii) bc[ii] = __shfl_sync(0xffffffffu, pick<T, CPL>(a[i0 + ii], un), ln);
#pragma unroll
                for (int ii = 0; ii < CH; ++ii) {
                    const int i = i0 + ii;
                    const int rl = w + kWarps * i;
                    const int grow = r0 + rl;
                    const T pe = ((rl < R) && (grow > j + 1)) ? bc[ii] : (T)0;
#pragma unroll
                    for (int u = 0; u < CPL; ++u) acc[u] += a[i][u] * pe;
                }
            }
        }
        } else {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int rl = w + kWarps * i;
            const int grow = r0 + rl;
            const T xj = __shfl_sync(0xffffffffu, pick<T, CPL>(a[i], uj), lj);
            const bool act = (rl < R) && (grow >= j);
            const T wv = act ? ((grow == j) ? (T)1 : xj * alpha) : (T)0;
            const T cj = (grow == j) ? beta : wv;
#pragma unroll
            for (int u = 0; u < CPL; ++u) {
                const T nv = a[i][u] - wv * fm[u];
                a[i][u] = (isj[u] && act) ? cj : nv;
            }
            if (more) {
                const T pn = __shfl_sync(0xffffffffu, pick<T, CPL>(a[i], un), ln);
                const T pe = ((rl < R) && (grow > j + 1)) ? pn : (T)0;
#pragma unroll
                for (int u = 0; u < CPL; ++u) acc[u] += a[i][u] * pe;
            }
        }
        }
        // psum / zs / piv are rewritten only after the next __syncthreads-protected phases
        __syncthreads();
        PANEL_TICK(7);
    }
    PANEL_TICK(8);

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
    PANEL_TICK(9);
}

inline size_t reg_smem_bytes(int rows, int b, size_t esz) {
    const size_t stage = esz == 8 ? stage_elems<double>(rows, b, 16) : stage_elems<float>(rows, b, 16);
    return ((size_t)4 * b + stage + (size_t)b * b + 3 * (size_t)b + (size_t)kWarps * b + kThreads + 8) * esz;
}

template <typename T, bool kTrans, int RPT, int CPL>
int launch_reg(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream, bool cooperative) {
    constexpr int ROWS = RPT * kWarps;
    const int G = (m + ROWS - 1) / ROWS;
    if (G > c->num_sms || G > kMaxPanelCtas) return 1;
    // beside a running stage-2 kernel (bidiagonalize_many) only single-cluster launches are allowed: kernels whose
    // clusters / CTAs wait for each other need all of them co-resident, which a concurrent cooperative kernel can prevent
    if (c->overlap_safe && (!c->cluster_ok || G > c->cluster_ok)) return 1;
    const size_t smem = reg_smem_bytes(ROWS, b, sizeof(T));
    T* red = reinterpret_cast<T*>(c->red);
    unsigned* bar = c->bar;
    if (c->cluster_ok) {
        // clusters of CS CTAs; more than one cluster when the panel is taller than CS * ROWS rows (two-level all-reduce).
        // All clusters of a launch must be co-resident (they wait for each other's words): ask the occupancy API.
        auto kern = panel_reg_kernel<T, kTrans, true, RPT, CPL>;
        SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->cluster_ok > 8) SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        const int cands[2] = {c->cluster_ok, c->cluster_ok > 8 ? 8 : 0};
        for (int t = 0; t < 2; ++t) {
            const int cs_try = cands[t];
            if (cs_try <= 0) continue;
            const int CS = G <= cs_try ? G : cs_try;
            const int NC = (G + CS - 1) / CS;
            if (NC > kMaxClusters) continue;
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
                // the occupancy query costs ~0.1 ms of host time: once per (cluster size, shared-memory size)
                static int cache_cs[8], cache_smem[8], cache_val[8], cache_n = 0;
                int max_clusters = -1;
                for (int q = 0; q < cache_n; ++q)
                    if (cache_cs[q] == CS && cache_smem[q] == (int)smem) max_clusters = cache_val[q];
                if (max_clusters < 0) {
                    max_clusters = 0;
                    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
                    if (cache_n < 8) { cache_cs[cache_n] = CS; cache_smem[cache_n] = (int)smem; cache_val[cache_n] = max_clusters; ++cache_n; }
                }
                if (NC > max_clusters) continue;
            }
            cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, lda, m, b, V, V2, red, bar, NC, ++c->panel_epoch, c->panel_run_if);
            if (e == cudaSuccess) { c->launches++; return 0; }
            cudaGetLastError();
            if (NC == 1) {                     // cluster shape not schedulable here: smaller clusters, then the grid transport
                c->cluster_ok = c->cluster_ok > 8 ? 8 : 0;
                return launch_reg<T, kTrans, RPT, CPL>(c, a, lda, m, b, V, V2, stream, cooperative);
            }
        }
    }
    auto kern = panel_reg_kernel<T, kTrans, false, RPT, CPL>;
    SVDB_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SVDB_CHECK(c, cudaMemsetAsync(c->bar, 0, 2 * sizeof(unsigned), stream));
    int nc0 = 0;
    if (cooperative) {
        unsigned ep0 = 0;
        const int* run_if = c->panel_run_if;
        void* args[] = {&a, &lda, &m, &b, &V, &V2, &red, &bar, &nc0, &ep0, &run_if};
        SVDB_CHECK(c, cudaLaunchCooperativeKernel((void*)kern, dim3(G), dim3(kThreads), args, smem, stream));
    } else {
        // look-ahead panels run beside the trailing update: a cooperative launch is gang-scheduled
        // and would wait for the update to drain; G <= #SMs CTAs of this size always become
        // co-resident once the update's CTAs retire, so the software barrier still completes.
        kern<<<G, kThreads, smem, stream>>>(a, lda, m, b, V, V2, red, bar, nc0, 0u, c->panel_run_if);
        SVDB_CHECK(c, cudaGetLastError());
    }
    c->launches++;
    return 0;
}

}  // namespace

// debug: read (and clear) the per-phase cycle counters of the timing build
int panel_reg_debug_read(long long* out16) {
    long long z[16] = {};
    if (cudaMemcpyFromSymbol(out16, g_panel_dbg, sizeof(z)) != cudaSuccess) return 1;
    cudaMemcpyToSymbol(g_panel_dbg, z, sizeof(z));
    return 0;
}

// returns 0 when it ran, 1 when the shape is outside this kernel's range (caller falls back)
template <typename T, bool kTrans>
int launch_panel_reg(Ctx* c, T* a, size_t lda, int m, int b, T* V, T* V2, cudaStream_t stream, bool cooperative) {
    // rows per CTA = 8 * RPT: the smallest slice that still fits the panel into 128 co-resident CTAs (eight 16-CTA
    // clusters), so that the per-column register pass shrinks with the panel height
    // (double only: in float the column loop is bound by the exchange, and more CTAs only add participants)
    const int cap = (sizeof(T) == 8 && !c->overlap_safe) ? 128 : 0;
    if (b <= 32) {
        if (m <= cap * 64) return launch_reg<T, kTrans, 8, 1>(c, a, lda, m, b, V, V2, stream, cooperative);
        if (m <= cap * 128) return launch_reg<T, kTrans, 16, 1>(c, a, lda, m, b, V, V2, stream, cooperative);
        if (sizeof(T) == 8 || m <= 128 * 256) return launch_reg<T, kTrans, 32, 1>(c, a, lda, m, b, V, V2, stream, cooperative);
        // very tall float panels (multi-GPU stage 1, n = 65536): 512 / 1024 rows per CTA keep the launch within 128 CTAs
        if (m <= 128 * 512) return launch_reg<float, kTrans, 64, 1>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream, cooperative);
        return launch_reg<float, kTrans, 128, 1>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream, cooperative);
    }
    if (b <= 64) {
        if (m <= cap * 32) return launch_reg<T, kTrans, 4, 2>(c, a, lda, m, b, V, V2, stream, cooperative);
        if (m <= cap * 64) return launch_reg<T, kTrans, 8, 2>(c, a, lda, m, b, V, V2, stream, cooperative);
        if (sizeof(T) == 8 || m <= 128 * 128) return launch_reg<T, kTrans, 16, 2>(c, a, lda, m, b, V, V2, stream, cooperative);
        if (m <= 128 * 256) return launch_reg<float, kTrans, 32, 2>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream, cooperative);
        return launch_reg<float, kTrans, 64, 2>(c, (float*)a, lda, m, b, (float*)V, (float*)V2, stream, cooperative);
    }
    return 1;
}
template int launch_panel_reg<float, false>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t, bool);
template int launch_panel_reg<float, true>(Ctx*, float*, size_t, int, int, float*, float*, cudaStream_t, bool);
template int launch_panel_reg<double, false>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t, bool);
template int launch_panel_reg<double, true>(Ctx*, double*, size_t, int, int, double*, double*, cudaStream_t, bool);

}  // namespace svdb200
