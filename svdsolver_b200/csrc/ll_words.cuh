// Flag-stamped words (the "LL" idea) for latency-bound exchanges between CTAs: every 8-byte half carries 32 data bits and
// the 32-bit sequence number of the exchange round it belongs to, so a reader that sees the expected number in both halves
// has the data -- one round trip, no separate flag, no fence.  Entries are 16 bytes for both element types (float uses
// the first half).  LLWord: through L2 (between clusters); LLSmem: pushed into a PEER CTA's shared memory over DSMEM
// (st.shared::cluster on the mapa address), the peer polls its own copy -- no barrier.cluster per round.
#pragma once
#include <cuda_runtime.h>

namespace svdb200 {

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

// a spin wait that can never hang the GPU: trap after ~4 s without progress (e.g. a cluster that never became resident)
struct SpinGuard {
    unsigned polls = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ void tick() {
        if ((++polls & 0xffffu) == 0u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
};

}  // namespace svdb200
