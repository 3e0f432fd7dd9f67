// In-kernel sum all-reduce of a small table over the GPUs of one node (fused multi-GPU mode).
//
// Every rank owns an exchange buffer (cudaMalloc + cudaIpc handle, mapped by all peers):
//     [flag of source 0 | flag of source 1 | ...]  (HMY_XFLAG_STRIDE apart)
//     payload[parity][source][xslot bytes]          (from HMY_XPAYLOAD_OFF)
// One-shot all-to-all: the calling CTA (the last to arrive at the round kernel's grid barrier,
// so the local table is complete) stores its table into slot [parity][my rank] of EVERY rank
// over NVLink, fences at system scope, raises flag[my rank] = seq everywhere, waits until all
// its own flags reach seq, and sums the copies in rank order -- every rank gets bit-identical
// totals.  Payloads are 8-500 KB and the step is latency-bound: ~2 NVLink hops instead of an
// NCCL launch + host round trip per block.  Slot reuse is safe with two parities: a rank can
// only start exchange s+2 after exchange s+1 completed, which needs every peer's flag s+1,
// which a peer raises only after it finished reading the slots of exchange s.
#pragma once
#include "hmy_common.cuh"

__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 16-byte vector loads/stores, 8 per thread in flight (a .cg load may not be hoisted over a
// store by the compiler, so the batching is explicit: one round trip per batch, not per element)
__device__ __forceinline__ void xchg_copy16(uint4* dst, const uint4* src, int n16) {
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int i0 = tid; i0 < n16; i0 += 8 * nth) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * nth; if (i < n16) v[u] = __ldcg(src + i); }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * nth; if (i < n16) dst[i] = v[u]; }
    }
}

// T = float or double; `count` elements at `table` (global, complete on this rank, 16-byte
// aligned, count * sizeof(T) a multiple of 16 is NOT required: the tail is handled scalar);
// all threads of the CTA.
template <class T>
__device__ __noinline__ void xchg_allreduce(const HmyDev& st, T* table, int count, unsigned int seq) {
    constexpr int PER16 = 16 / (int)sizeof(T);
    const int W = st.xworld, me = st.xrank, tid = threadIdx.x, nth = blockDim.x;
    // vector path only for 16-byte aligned tables (odd B*K shifts the per-block tables by 4 bytes)
    const int n16 = ((reinterpret_cast<unsigned long long>(table) & 15ull) == 0ull) ? count / PER16 : 0;
    const int tail0 = n16 * PER16;
    const size_t slot_off = HMY_XPAYLOAD_OFF + ((size_t)(seq & 1u) * HMY_MAX_WORLD + me) * st.xslot;
    // 1. my table -> slot [parity][me] on every rank
    for (int r = 0; r < W; ++r) {
        T* dst = reinterpret_cast<T*>(st.xpeer[r] + slot_off);
        xchg_copy16(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(table), n16);
        for (int i = tail0 + tid; i < count; i += nth) dst[i] = __ldcg(&table[i]);
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, 3. wait for every source
    if (tid < W) {
        st_release_sys_u32(reinterpret_cast<unsigned int*>(st.xpeer[tid] + (size_t)me * HMY_XFLAG_STRIDE), seq);
        const unsigned int* f = reinterpret_cast<const unsigned int*>(st.xpeer[me] + (size_t)tid * HMY_XFLAG_STRIDE);
        while ((int)(ld_acquire_sys_u32(f) - seq) < 0) { __nanosleep(40); }
    }
    __syncthreads();
    __threadfence_system();
    // 4. sum the copies in rank order (loads of a batch first, then the adds)
    const unsigned char* mine = st.xpeer[me] + HMY_XPAYLOAD_OFF + (size_t)(seq & 1u) * HMY_MAX_WORLD * st.xslot;
    for (int i0 = tid; i0 < n16; i0 += 2 * nth) {
        uint4 v[2][HMY_MAX_WORLD];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * nth;
#pragma unroll
            for (int r = 0; r < HMY_MAX_WORLD; ++r)
                if (r < W && i < n16) v[u][r] = __ldcg(reinterpret_cast<const uint4*>(mine + (size_t)r * st.xslot) + i);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * nth;
            if (i < n16) {
                T acc[PER16];
#pragma unroll
                for (int e = 0; e < PER16; ++e) acc[e] = (T)0;
#pragma unroll
                for (int r = 0; r < HMY_MAX_WORLD; ++r) {
                    if (r < W) {
                        const T* pv = reinterpret_cast<const T*>(&v[u][r]);
#pragma unroll
                        for (int e = 0; e < PER16; ++e) acc[e] += pv[e];
                    }
                }
                __stcg(reinterpret_cast<uint4*>(table) + i, *reinterpret_cast<const uint4*>(acc));
            }
        }
    }
    for (int i = tail0 + tid; i < count; i += nth) {
        T s = (T)0;
        for (int r = 0; r < W; ++r) s += __ldcg(reinterpret_cast<const T*>(mine + (size_t)r * st.xslot) + i);
        __stcg(&table[i], s);
    }
    __syncthreads();
}

// ---- low-latency variant for the per-block K x B float tables ------------------------------------
// LL protocol (as in NCCL's LL): every element travels as one 8-byte store {value, seq}; an 8-byte
// store is atomic, so the receiver polls the payload itself -- no system-scope fence and no separate
// flag round trip between "data written" and "data visible".  Twice the bytes, which is irrelevant
// for 8-27 KB tables.  Slot layout: payload_ll[parity][source][count] of uint2, placed after the
// fenced-protocol slots (see hmy_comm_export).  The two parities make a slot reusable: a rank can be
// at most one exchange ahead of the slowest reader of the previous use of that parity.
__device__ __forceinline__ void st_volatile_v2(uint2* p, unsigned int a, unsigned int b) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

static __device__ __noinline__ void xchg_allreduce_ll_f32(const HmyDev& st, float* table, int count, unsigned int seq) {
    const int W = st.xworld, me = st.xrank, tid = threadIdx.x, nth = blockDim.x;
    const size_t ll_base = HMY_XPAYLOAD_OFF + 2 * (size_t)HMY_MAX_WORLD * st.xslot;
    const size_t slot = (size_t)st.xll_count * sizeof(uint2);
    const size_t my_off = ll_base + ((size_t)(seq & 1u) * HMY_MAX_WORLD + me) * slot;
    // push {value, seq} to every rank (own copy included: the reduction reads all sources alike)
    for (int i0 = tid; i0 < count; i0 += 8 * nth) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * nth; v[u] = (i < count) ? __ldcg(&table[i]) : 0.f; }
        for (int r = 0; r < W; ++r) {
            uint2* dst = reinterpret_cast<uint2*>(st.xpeer[r] + my_off);
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = i0 + u * nth; if (i < count) st_volatile_v2(dst + i, __float_as_uint(v[u]), seq); }
        }
    }
    // receive: poll each element of each source until it carries this sequence number
    const unsigned char* mine = st.xpeer[me] + ll_base + (size_t)(seq & 1u) * HMY_MAX_WORLD * slot;
    for (int i = tid; i < count; i += nth) {
        float s = 0.f;
        for (int r = 0; r < W; ++r) {
            const uint2* src = reinterpret_cast<const uint2*>(mine + (size_t)r * slot) + i;
            uint2 v = ld_volatile_v2(src);
            while (v.y != seq) { __nanosleep(20); v = ld_volatile_v2(src); }
            s += __uint_as_float(v.x);
        }
        __stcg(&table[i], s);
    }
    __syncthreads();
}
