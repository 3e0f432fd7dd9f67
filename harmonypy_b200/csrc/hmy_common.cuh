// Shared device-side definitions for the Harmony engine (sm_100a).
//
// HBM layout (all cell-major, cells physically sorted by covariate combination so that
// every contiguous run of cells shares its batch levels -- see DESIGN.md "Data layout"):
//   Zorig, Zcos, Zcorr : [N][dp]  fp32, dp = round_up(d, 4), zero padded (16-byte rows)
//   R                  : [N][Kp]  fp32, Kp = round_up(K, 4)
//   combo              : [N]      int32 id of the covariate combination of the cell
//   combo_lev          : [ncombo][V] int32 global one-hot row of each covariate's level
//   blk                : [N]      uint8 update block of the cell in the current round
// Small tables (L2 / shared-memory resident):
//   Yhat [K][dp] fp32 unit centroids, Yacc [K][dp] fp64 centroid sums,
//   Told/Dnew [nblk][B][K] fp32 per-block removed / re-added batch sums,
//   P [B][K] fp32 diversity penalty, O/Orun [B][K] fp64, obj[4] fp64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HMY_THREADS 256
#define HMY_WARPS 8
#define HMY_TILE 64          // cells per tile in every streaming kernel
#define HMY_CPW 8            // cells per warp inside a tile (HMY_TILE / HMY_WARPS)
#define HMY_MAX_V 8
#define HMY_MAX_NBLK 250
#define HMY_TRACE_SLOTS 256
#define HMY_MAX_WORLD 8
#define HMY_XFLAG_STRIDE 128     // bytes between the per-source flags of an exchange buffer
#define HMY_XPAYLOAD_OFF 4096

struct HmyDev {
    long long N;             // cells on this rank
    long long Nglobal;
    long long cell_offset;
    long long cpb;           // int(Nglobal * block_size), harmony.py:475
    int d, dp, K, Kp, KS;    // KS = 32 * KPT, shared-memory stride over clusters
    int V, B, nblk, ncombo;
    int lev0;                // levels of covariate 0 (their one-hot rows are [0, lev0))
    int nseg;                // ridge work items
    int lambda_estimation;
    float alpha;
    // cell-major state
    float* Zorig; float* Zcos; float* Zcorr; float* R;
    int* combo; int* combo_lev; long long* combo_start;   // [ncombo+1] first cell of each combo
    int* order;              // order[pos] = caller's local index of the cell stored at pos
    int* pos_of;             // inverse of order
    unsigned char* blk;
    int* list; long long* blk_start;     // cells of block b: list[blk_start[b] .. blk_start[b+1]) (ascending positions)
    int* seg;                            // ridge work items: [nseg][3] = start, count, combo
    // tables
    float* Yhat; float* Ynext; double* Yacc;   // Yhat: centroids this stage reads; Ynext: the ones it produces
    float* Told; float* Dnew; float* P;
    double* O; double* Orun;             // [B][K]
    double* Ofresh;                      // [B][K] sum over the round's blocks of the re-added batch sums = the new O
    double* obj;                         // [0..2] objective sums, [3] spare
    double* obj_out;                     // [3] finished objective of the last stage
    float* Pr_b; float* theta; float* sigma; float* lamb;   // lamb[B+1]
    double* Gram;            // [K][B+1][B+1]
    double* Mom;             // [B+1][K][dp]
    float* W;                // [B][K][dp]
    float* wmax;             // max |W| of the last solve
    double* solve_scratch;   // [K][(B+1)(B+1+d) + B+1] work area of the ridge solve when it does not fit in shared memory (else nullptr)
    // grid barrier
    unsigned int* bar_count; unsigned int* bar_gen;
    // optional per-CTA timeline (globaltimer ns), [grid][HMY_TRACE_SLOTS]; nullptr = off
    unsigned long long* trace;
    // fused multi-GPU exchange: peer-mapped exchange buffers (own one included), see hmy_xchg.cuh
    unsigned char* xpeer[HMY_MAX_WORLD];
    int xrank, xworld;
    int xrelaxed;                    // 1: exchange only at the end of a round (blocks see local updates only)
    unsigned int xseq_base;          // sequence number of the last exchange before this launch
    unsigned long long xslot;        // bytes of one payload slot (fenced protocol)
    int xll_count;                   // elements of one low-latency slot (K x B)
    // ---- tensor-memory round kernel (hmy_round_tc5.cuh)
    unsigned short* Zs16;            // [N][32 dt] fp16: hi[16 dt] | lo[16 dt] of 1024 * Z_cos, dt = ceil(d / 16)
    int2* list2;                     // annotated block lists of the round: {cell, combo << 8 | block of the cell in the NEXT round}
    unsigned char* blk_next;         // block of every cell in the next round
    float* Told_next;                // [nblk][B][K] sums the NEXT round's blocks remove (accumulated by this round)
    double* Rsum; double* Rsum_next; // [K] running sum_n R[n][k] of this round / start value of the next one
    unsigned long long* bar64;       // monotone grid-barrier counter
    int write_R;                     // 1: the round stores the new R rows to HBM
    unsigned long long x5_off_d, x5_off_t, x5_off_y;   // byte offsets of the LL regions of that kernel in every rank's exchange buffer
    unsigned int x5_seq;             // sequence number of this launch (LL packets carry it)
    int dbg;                         // timing experiments only (option "dbg", results are WRONG when set): see hmy_round_tc5.cuh
    int sigma_uniform; float sigma_u;   // every cluster has the same sigma (the usual case): no per-cluster constants to load
};

__device__ __forceinline__ void hmy_trace(const HmyDev& st, int slot) {
    if (st.trace != nullptr && threadIdx.x == 0 && slot < HMY_TRACE_SLOTS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        st.trace[(size_t)blockIdx.x * HMY_TRACE_SLOTS + slot] = t;
    }
}

// the same from any single thread (warp-specialised kernels: the issuing thread of a role)
__device__ __forceinline__ void hmy_trace_any(const HmyDev& st, int slot) {
    if (st.trace != nullptr && slot < HMY_TRACE_SLOTS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        st.trace[(size_t)blockIdx.x * HMY_TRACE_SLOTS + slot] = t;
    }
}

// Longest gap between successive events of one role (one thread calls it): slot `base` holds the last event's time,
// base + 1 + site the longest gap that ended at `site`, base + 1 + nsites + site where (caller's tag) it happened.
__device__ __forceinline__ void hmy_trace_gap(const HmyDev& st, int base, int nsites, int site, unsigned long long tag) {
    if (st.trace != nullptr) {
        volatile unsigned long long* T = st.trace + (size_t)blockIdx.x * HMY_TRACE_SLOTS;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        const unsigned long long prev = T[base];
        if (prev != 0 && now - prev > T[base + 1 + site]) { T[base + 1 + site] = now - prev; T[base + 1 + nsites + site] = tag; }
        T[base] = now;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier with a serial section: every CTA arrives; the LAST one to arrive runs
// `serial()` (all of its threads) and then releases generation `gen`; the others wait.
// Requires all CTAs of the grid to be co-resident (cooperative launch).
// Memory ordering follows the cooperative-groups grid.sync() pattern: bar.sync makes the
// CTA's prior writes (incl. its REDs) happen-before thread 0's gpu-scope fence, which is
// cumulative; the ticket atomic / the release store then publish them.
template <class F>
__device__ __forceinline__ void grid_barrier_serial(const HmyDev& st, unsigned int nctas,
                                                    unsigned int gen, int* s_flag, F serial) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(st.bar_count, 1u);
        *s_flag = (t == nctas - 1u);
    }
    __syncthreads();
    if (*s_flag) {
#define HMY_SER_STAMP(i_)                                                                        \
        if (st.trace != nullptr && threadIdx.x == 0) {                                             \
            unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));          \
            st.trace[(size_t)nctas * HMY_TRACE_SLOTS + 4 * (gen & 31u) + (i_)] = t_;               \
        }
        HMY_SER_STAMP(0)
        __threadfence();                 // acquire side: see every other CTA's contributions
        HMY_SER_STAMP(1)
        serial();
        __syncthreads();
        HMY_SER_STAMP(2)
        if (threadIdx.x == 0) {
            *st.bar_count = 0u;
            __threadfence();
            HMY_SER_STAMP(3)
            st_release_u32(st.bar_gen, gen);
        }
    } else {
        if (threadIdx.x == 0) {
            while ((int)(ld_acquire_u32(st.bar_gen) - gen) < 0) { __nanosleep(20); }
        }
    }
    __syncthreads();
}

// Plain grid barrier (no serial section).
__device__ __forceinline__ void grid_barrier(const HmyDev& st, unsigned int nctas, unsigned int gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(st.bar_count, 1u);
        if (t == nctas - 1u) {
            *st.bar_count = 0u;
            __threadfence();
            st_release_u32(st.bar_gen, gen);
        } else {
            while ((int)(ld_acquire_u32(st.bar_gen) - gen) < 0) { __nanosleep(20); }
        }
        __threadfence();
    }
    __syncthreads();
}
