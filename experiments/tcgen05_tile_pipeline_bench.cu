// EXPERIMENT (not product code): throughput of the tensor-memory tile step of the clustering round under four
// ways of feeding and overlapping it -- the measurement the next version of hmy_round_tc5.cuh is designed from.
//
// One persistent CTA per SM walks its share of the cells block by block (20 blocks, cells of a block = a 5 %
// sample of the CTA's contiguous share, ascending positions: what the engine's block lists look like), 128 cells
// per tile, and does per tile exactly what experiments/tcgen05_tile_step_probe.cu validated on a B200, with the
// engine's epilogue math (softmax over 100 clusters, penalty, clamp, objective terms, R row to HBM):
//
//   MODE 0  "v1"        fp32 Z_cos rows gathered through registers, split to fp16 hi/lo by the CTA, everything
//                       synchronous (what hmy_round_tc5.cuh does today)
//   MODE 1  "presplit"  Z_cos kept a second time as fp16 hi|lo rows (256 B/cell): a tile is 16 cp.async of 16 B
//                       per thread straight into the core-matrix layout; still synchronous
//   MODE 2  "pipelined" MODE 1 with three Z buffers and two score accumulators: the gather of tile t+2 and the
//                       scoring of tile t+1 run under the epilogue of tile t; the R row is stored to HBM after
//                       the accumulation MMAs have been issued
//   MODE 2 / 256 threads the same with the clusters of a cell split over two warps (warps w and w+4 read the same
//                       tensor-memory lanes): 56 scores per thread, two warps per scheduler
//
// No grid barriers, no phase 0, no penalty-table update: this isolates the tile step (the rest of the round is
// the same code as today).  Output: microseconds per pass over all cells and per tile, for each mode; all modes
// are checked against a double-precision CPU evaluation at a smaller N first.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench experiments/tcgen05_tile_pipeline_bench.cu
//   ./pipe_bench --hostcheck            # CPU: consistency of the generated work lists
//   timeout 120 ./pipe_bench            # verify at N = 131072, then time at N = 1048576
//
// Compile-checked in the build container; NOT yet run on hardware.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef ABL
#define ABL 0      // ablation bit mask (timing runs only): 1 no HBM R store, 2 no R operand (smem) store, 4 no lg2 term,
#endif             // 8 no ex2, 16 no accumulation MMAs, 32 no Z gather, 64 no scoring MMAs
constexpr int TILE = 128, DP = 64, DPF = 52, D = 50, K = 100, KP = 112, KM = 128, NB = 32, NLEV = 8, NBLK = 20;
constexpr float OPSCALE = 1024.f, ACCSCALE = 1.f / 1048576.f;
constexpr int TMEM_COLS = 512;                 // D1[0]: [0,112)  D1[1]: [128,240)  D2y: [256,320)  D2o: [320,352)
constexpr int COL_D1B = 128, COL_Y = 256, COL_O = 320;

constexpr int Z_LBO_K = 128, Z_SBO_K = 1024, Z_LBO_MN = 1024, Z_SBO_MN = 128;
constexpr int Y_LBO = 128, Y_SBO = 1024, R_LBO = 128, R_SBO = 2048, O_LBO = 128, O_SBO = 2048;
constexpr int SZ = TILE * DP * 2, SY = KP * DP * 2, SR = KM * TILE * 2, SO = NB * TILE * 2;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---- device helpers (same as the tile-step probe) ---------------------------------------------------------------
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ inline uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ inline void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ inline void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ inline void wait_parity(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && (++spins & 1023u) == 0u) {             // a lost completion must be an error, not a hang
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t; else if (t - t0 > 2000000000ull) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ inline void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
}
__device__ inline void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t u[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(u[j]);
}
__device__ inline void publish() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ inline float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ inline float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ inline void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 H = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(H);
    const __half2 L = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&H);
    lo = *reinterpret_cast<const uint32_t*>(&L);
}
__device__ inline void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Args {
    const float* Zf;         // [N][DPF] fp32 unit rows
    const __half* Zs;        // [N][2*DP] fp16: hi[64] | lo[64] of 1024 * Z
    const float* Y;          // [KP][DP]
    const float* sigma;      // [KP]
    const float* P;          // [NLEV][KP] penalty rows
    const int* lev;          // [N] level (= one-hot column) of each cell
    const int* list;         // concatenated cell lists
    const int* tile_first;   // [G+1] first tile of each CTA
    const int* tile_off;     // [tiles] offset of the tile's first cell in `list`
    const int* tile_nt;      // [tiles] cells in the tile | (last tile of its block ? 1 << 30 : 0)
    float* R;                // [N][KP]
    float* Yslab;            // [G][KM][DP]
    float* Oslab;            // [G][KM][NB]
    float* obj;              // [G][256][2]
};

// ---- one CTA ------------------------------------------------------------------------------------------------------
template <int MODE, int NTHR>
__global__ void __launch_bounds__(NTHR, 1) tile_pipeline(Args a) {
    constexpr int NZ = (MODE == 2) ? 3 : 1;          // Z / one-hot buffers
    constexpr int H = NTHR / 128;                    // threads per cell (each owns CW clusters)
    constexpr int CW = KP / H;
    static_assert(H == 1 || (H == 2 && MODE == 2), "256 threads: pipelined mode only");
    static_assert(CW % 8 == 0, "column split");
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sZ = smem;                              // NZ x (hi | lo)
    unsigned char* sO = sZ + NZ * 2 * SZ;                  // NZ x one-hot
    unsigned char* sRh = sO + NZ * SO;
    unsigned char* sRl = sRh + SR;
    unsigned char* sYh = sRl + SR;
    unsigned char* sYl = sYh + SY;
    float* sP = reinterpret_cast<float*>(sYl + SY);        // [NLEV][KP]
    float* sc1 = sP + NLEV * KP;
    float* sc3 = sc1 + KP;
    int* sCell = reinterpret_cast<int*>(sc3 + KP);         // [128] (MODE 0 gather)
    float* sXch = reinterpret_cast<float*>(sCell + 128);   // [2][128][2] partial row sums (H = 2)
    __shared__ __align__(8) uint64_t bar_s[2], bar_a;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ct = tid & 127, half = tid >> 7, cb = half * CW;     // this thread's cell of the tile / first cluster

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_s[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_s[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_a)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < KP * (DP / 2); i += NTHR) {
        const int k = i / (DP / 2), j = 2 * (i % (DP / 2));
        uint32_t hi, lo;
        split2(a.Y[k * DP + j] * OPSCALE, a.Y[k * DP + j + 1] * OPSCALE, hi, lo);
        const int off = (k >> 3) * Y_SBO + (j >> 3) * Y_LBO + (k & 7) * 16 + (j & 7) * 2;
        *reinterpret_cast<uint32_t*>(sYh + off) = hi;
        *reinterpret_cast<uint32_t*>(sYl + off) = lo;
    }
    for (int i = tid; i < NLEV * KP; i += NTHR) sP[i] = a.P[i];
    for (int k = tid; k < KP; k += NTHR) {
        const float sg = a.sigma[k];
        sc1[k] = (k < K) ? (2.0f * 1.4426950408889634f / sg) * ACCSCALE : 0.f;
        sc3[k] = (k < K) ? sg * 0.6931471805599453f : 0.f;
    }
    for (int i = tid; i < (NZ * (2 * SZ + SO) + 2 * SR) / 16; i += NTHR) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base, lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const uint32_t sb = smem_u32(smem);
    const uint32_t id_score = make_idesc(TILE, KP, 0, 0), id_y = make_idesc(KM, DP, 1, 1), id_o = make_idesc(KM, NB, 1, 1);

    const int t_begin = a.tile_first[blockIdx.x], T = a.tile_first[blockIdx.x + 1] - t_begin;
    float objd = 0.f, obje = 0.f;
    bool y_started = false, o_started = false;
    float* Yslab = a.Yslab + (size_t)blockIdx.x * KM * DP;
    float* Oslab = a.Oslab + (size_t)blockIdx.x * KM * NB;

    // ---- building blocks -----------------------------------------------------------------------------------------
    // gather tile t into buffer b (MODE 0: through registers, blocking; MODE 1/2: cp.async, one commit group)
    auto load_tile = [&](int t, int b) {
        const int off = a.tile_off[t_begin + t], nt = a.tile_nt[t_begin + t] & 0xFFFF;
        unsigned char* Zh = sZ + b * 2 * SZ;
        unsigned char* Zl = Zh + SZ;
        // one-hot level row of this thread's cell
        if (tid < 128) {
            const int lv = (tid < nt) ? a.lev[a.list[off + tid]] : -1;
#pragma unroll
            for (int j = 0; j < NB / 8; ++j) {
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (lv >= 8 * j && lv < 8 * j + 8) {
                    const uint32_t h = (lv & 1) ? 0x3C000000u : 0x3C00u;
                    const int q = (lv & 7) >> 1;
                    w.x = q == 0 ? h : 0u; w.y = q == 1 ? h : 0u; w.z = q == 2 ? h : 0u; w.w = q == 3 ? h : 0u;
                }
                *reinterpret_cast<uint4*>(sO + b * SO + j * O_SBO + (tid >> 3) * O_LBO + (tid & 7) * 16) = w;
            }
        }
        if (MODE == 0) {
            sCell[tid] = (tid < nt) ? a.list[off + tid] : 0;
            __syncthreads();
            constexpr int dp4 = DPF / 4, ZU = 8;
            const int total = nt * dp4;
            for (int base = 0; base < total; base += ZU * 128) {
                float4 zr[ZU];
#pragma unroll
                for (int u = 0; u < ZU; ++u) {
                    const int i = base + tid + u * 128;
                    if (i < total) { const int row = i / dp4, c4 = i - row * dp4; zr[u] = __ldg(reinterpret_cast<const float4*>(a.Zf + (size_t)sCell[row] * DPF) + c4); }
                }
#pragma unroll
                for (int u = 0; u < ZU; ++u) {
                    const int i = base + tid + u * 128;
                    if (i < total) {
                        const int row = i / dp4, c4 = i - row * dp4;
                        uint2 hi, lo;
                        split2(zr[u].x * OPSCALE, zr[u].y * OPSCALE, hi.x, lo.x);
                        split2(zr[u].z * OPSCALE, zr[u].w * OPSCALE, hi.y, lo.y);
                        const int o = (row >> 3) * Z_SBO_K + (c4 >> 1) * Z_LBO_K + (row & 7) * 16 + (c4 & 1) * 8;
                        *reinterpret_cast<uint2*>(Zh + o) = hi;
                        *reinterpret_cast<uint2*>(Zl + o) = lo;
                    }
                }
            }
        } else {
            // 16 chunks of 16 B per row: chunks 0..7 = hi PCs 8c..8c+7, 8..15 = lo; 16 consecutive threads = one row
#pragma unroll
            for (int u = 0; u < 2048 / NTHR; ++u) {
                const int q = tid + u * NTHR, row = q >> 4, ch = q & 15;
                if (row < nt && !(ABL & 32)) {
                    const int cell = a.list[off + row];
                    const uint32_t dst = smem_u32((ch < 8 ? Zh : Zl) + (row >> 3) * Z_SBO_K + (ch & 7) * Z_LBO_K + (row & 7) * 16);
                    cp_async16(dst, a.Zs + (size_t)cell * (2 * DP) + ch * 8);
                }
            }
            cp_async_commit();
        }
    };
    auto issue_score = [&](int b, int d1) {       // buffer b -> D1[d1]
        if (tid == 0) {
            const uint32_t zh = sb + b * 2 * SZ, zl = zh + SZ;
            for (int ks = 0; ks < ((ABL & 64) ? 0 : DP / 16); ++ks) {
                const uint32_t o = ks * 2 * Z_LBO_K;
                const uint64_t dzh = make_desc(zh + o, Z_LBO_K, Z_SBO_K), dzl = make_desc(zl + o, Z_LBO_K, Z_SBO_K);
                const uint64_t dyh = make_desc(smem_u32(sYh) + o, Y_LBO, Y_SBO), dyl = make_desc(smem_u32(sYl) + o, Y_LBO, Y_SBO);
                mma(tbase + d1 * COL_D1B, dzl, dyh, id_score, ks > 0 ? 1u : 0u);
                mma(tbase + d1 * COL_D1B, dzh, dyl, id_score, 1u);
                mma(tbase + d1 * COL_D1B, dzh, dyh, id_score, 1u);
            }
            commit(&bar_s[d1]);
        }
    };
    auto issue_acc = [&](int b, int nt) {
        if (tid == 0) {
            const uint32_t zh = sb + b * 2 * SZ, zl = zh + SZ, ot = smem_u32(sO) + b * SO;
            const int ksteps = (nt + 15) >> 4;
            for (int ks = 0; ks < ((ABL & 16) ? 0 : ksteps); ++ks) {
                const uint64_t rh = make_desc(smem_u32(sRh) + ks * 2 * R_LBO, R_LBO, R_SBO), rl = make_desc(smem_u32(sRl) + ks * 2 * R_LBO, R_LBO, R_SBO);
                const uint64_t dzh = make_desc(zh + ks * 2 * Z_LBO_MN, Z_LBO_MN, Z_SBO_MN), dzl = make_desc(zl + ks * 2 * Z_LBO_MN, Z_LBO_MN, Z_SBO_MN);
                const uint64_t dot = make_desc(ot + ks * 2 * O_LBO, O_LBO, O_SBO);
                mma(tbase + COL_Y, rl, dzh, id_y, (y_started || ks > 0) ? 1u : 0u);
                mma(tbase + COL_Y, rh, dzl, id_y, 1u);
                mma(tbase + COL_Y, rh, dzh, id_y, 1u);
                mma(tbase + COL_O, rl, dot, id_o, (o_started || ks > 0) ? 1u : 0u);
                mma(tbase + COL_O, rh, dot, id_o, 1u);
            }
            commit(&bar_a);
        }
        y_started = true; o_started = true;
    };
    // thread = cluster: level sums of the finished block -> this CTA's slab
    auto flush_block = [&]() {
        if (tid < 128) {                              // whole warps: tcgen05.ld is warp-collective
            float v0[16], v1[16];
            tmem_ld16(tbase + lane_base + COL_O, v0);
            tmem_ld16(tbase + lane_base + COL_O + 16, v1);
#pragma unroll
            for (int j = 0; j < 16; ++j) { Oslab[tid * NB + j] += v0[j] * (1.f / OPSCALE); Oslab[tid * NB + 16 + j] += v1[j] * (1.f / OPSCALE); }
        }
        o_started = false;
    };
    // epilogue of one tile, thread = cell.  E stays in registers; store_global = false defers the HBM row
    float E[CW];
    float sc = 0.f;
    auto epilogue = [&](int d1, int lv, bool valid) {
        float ss = 0.f, sp = 0.f, sd = 0.f;
        const float* Pr = sP + (valid ? lv : 0) * KP;
        auto cols4 = [&](const float* acc, int col, int e0) {      // 4 clusters starting at `col` (a multiple of 4)
            const float4 k1 = *reinterpret_cast<const float4*>(sc1 + col), k3 = *reinterpret_cast<const float4*>(sc3 + col), pn = *reinterpret_cast<const float4*>(Pr + col);
            const float k1v[4] = {k1.x, k1.y, k1.z, k1.w}, k3v[4] = {k3.x, k3.y, k3.z, k3.w}, pv[4] = {pn.x, pn.y, pn.z, pn.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float t = fmaf(-acc[e], k1v[e], k1v[e] * 1048576.0f);
                const float s = (col + e < K) ? ((ABL & 8) ? fmaf(t, -1e-3f, 1.0f) : ex2_approx(-t)) : 0.f;
                ss += s;
                const float ev = s * pv[e];
                sp += ev;
                sd = fmaf(k3v[e], ev * t, sd);
                E[e0 + e] = ev;
            }
        };
#pragma unroll
        for (int ch = 0; ch < CW / 16; ++ch) {
            float acc[16];
            tmem_ld16(tbase + lane_base + (uint32_t)(d1 * COL_D1B + cb + 16 * ch), acc);
#pragma unroll
            for (int q = 0; q < 4; ++q) cols4(acc + 4 * q, cb + 16 * ch + 4 * q, 16 * ch + 4 * q);
        }
        if (CW % 16 == 8) {
            float acc[8];
            tmem_ld8(tbase + lane_base + (uint32_t)(d1 * COL_D1B + cb + CW - 8), acc);
#pragma unroll
            for (int q = 0; q < 2; ++q) cols4(acc + 4 * q, cb + CW - 8 + 4 * q, CW - 8 + 4 * q);
        }
        if (H == 2) {                                  // the two halves of a cell exchange their row sums
            sXch[(half * 128 + ct) * 2] = ss; sXch[(half * 128 + ct) * 2 + 1] = sp;
            __syncthreads();
            ss += sXch[((half ^ 1) * 128 + ct) * 2]; sp += sXch[((half ^ 1) * 128 + ct) * 2 + 1];
        }
        const float is = 1.f / ss;
        sc = valid ? is / fmaxf(sp * is, 1e-8f) : 0.f;
        float oe = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 8) {
            const float4 k3a = *reinterpret_cast<const float4*>(sc3 + cb + c0), k3b = *reinterpret_cast<const float4*>(sc3 + cb + c0 + 4);
            const float k3v[8] = {k3a.x, k3a.y, k3a.z, k3a.w, k3b.x, k3b.y, k3b.z, k3b.w};
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                r[j] = E[c0 + j] * sc;
                if (!(ABL & 4)) oe = fmaf(k3v[j], (r[j] > 0.f ? r[j] * lg2_approx(r[j]) : 0.f), oe);
                else oe = fmaf(k3v[j], r[j], oe);
            }
            uint4 hi, lo;
            split2(r[0] * OPSCALE, r[1] * OPSCALE, hi.x, lo.x);
            split2(r[2] * OPSCALE, r[3] * OPSCALE, hi.y, lo.y);
            split2(r[4] * OPSCALE, r[5] * OPSCALE, hi.z, lo.z);
            split2(r[6] * OPSCALE, r[7] * OPSCALE, hi.w, lo.w);
            const int off = ((cb + c0) >> 3) * R_SBO + (ct >> 3) * R_LBO + (ct & 7) * 16;
            if (!(ABL & 2)) {
                *reinterpret_cast<uint4*>(sRh + off) = hi;
                *reinterpret_cast<uint4*>(sRl + off) = lo;
            } else if (hi.x == 0x12345u && lo.y == 0x54321u) objd += 1.f;
        }
        if (valid) { objd += sc * sd; obje += oe; }
    };
    auto store_row = [&](int cell, bool valid) {
        if (!valid || (ABL & 1)) return;
        float* Rg = a.R + (size_t)cell * KP + cb;
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 4)
            *reinterpret_cast<float4*>(Rg + c0) = make_float4(E[c0] * sc, E[c0 + 1] * sc, E[c0 + 2] * sc, E[c0 + 3] * sc);
    };

    // ---- the tile loop ---------------------------------------------------------------------------------------------
    uint32_t ph_a = 0;                               // accumulation phases consumed
    if (MODE != 2) {
        uint32_t ph_s = 0;
        for (int t = 0; t < T; ++t) {
            const int off = a.tile_off[t_begin + t], ntf = a.tile_nt[t_begin + t], nt = ntf & 0xFFFF;
            if (t > 0) { wait_parity(&bar_a, ph_a & 1u); ph_a++; if (a.tile_nt[t_begin + t - 1] >> 30) flush_block(); }
            load_tile(t, 0);
            if (MODE == 1) cp_async_wait<0>();
            publish();
            issue_score(0, 0);
            const bool valid = ct < nt;
            const int cell = valid ? a.list[off + ct] : 0, lv = valid ? a.lev[cell] : 0;
            wait_parity(&bar_s[0], ph_s & 1u); ph_s++;
            epilogue(0, lv, valid);
            store_row(cell, valid);
            publish();
            issue_acc(0, nt);
        }
    } else {
        // three Z buffers (tile t lives in t % 3 until its accumulation is done), two score accumulators (t & 1)
        if (T > 0) load_tile(0, 0);
        if (T > 1) load_tile(1, 1);
        if (T > 0) {
            if (T > 1) cp_async_wait<1>(); else cp_async_wait<0>();
            publish();
            issue_score(0, 0);
        }
        for (int t = 0; t < T; ++t) {
            const int off = a.tile_off[t_begin + t], nt = a.tile_nt[t_begin + t] & 0xFFFF;
            // scoring of the NEXT tile first: it runs on the tensor pipe under this tile's epilogue
            if (t + 1 < T) {
                cp_async_wait<0>();                  // the only group in flight is tile t+1 (t+2 is issued below)
                publish();                           // also orders the epilogue(t-1) reads of D1[(t+1)&1] before the overwrite
                issue_score((t + 1) % 3, (t + 1) & 1);
            }
            if (t > 0) {                             // accumulation of tile t-1: frees its Z buffer and the R tile
                wait_parity(&bar_a, ph_a & 1u); ph_a++;
                if (a.tile_nt[t_begin + t - 1] >> 30) flush_block();
            }
            if (t + 2 < T) load_tile(t + 2, (t + 2) % 3);          // = buffer of tile t-1
            const bool valid = ct < nt;
            const int cell = valid ? a.list[off + ct] : 0, lv = valid ? a.lev[cell] : 0;
            wait_parity(&bar_s[t & 1], (uint32_t)(t >> 1) & 1u);
            epilogue(t & 1, lv, valid);
            publish();
            issue_acc(t % 3, nt);
            store_row(cell, valid);                  // HBM stores overlap the accumulation MMAs
        }
    }
    if (T > 0) { wait_parity(&bar_a, ph_a & 1u); ph_a++; flush_block(); }
    if (y_started && tid < 128) {
#pragma unroll
        for (int ch = 0; ch < DP / 16; ++ch) {
            float v[16];
            tmem_ld16(tbase + lane_base + (uint32_t)(COL_Y + 16 * ch), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) Yslab[tid * DP + 16 * ch + j] = v[j] * ACCSCALE;
        }
    }
    a.obj[((size_t)blockIdx.x * 256 + tid) * 2] = objd;
    a.obj[((size_t)blockIdx.x * 256 + tid) * 2 + 1] = obje;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(TMEM_COLS));
}

// ---- host ----------------------------------------------------------------------------------------------------------
struct Host {
    int N, G;
    std::vector<float> Zf, Y, sigma, P;
    std::vector<__half> Zs;
    std::vector<int> lev, list, tile_first, tile_off, tile_nt;
};

static uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

static Host make_host(int N, int G) {
    Host h; h.N = N; h.G = G;
    h.Zf.assign((size_t)N * DPF, 0.f); h.Zs.assign((size_t)N * 2 * DP, __float2half(0.f));
    h.Y.assign((size_t)KP * DP, 0.f); h.sigma.assign(KP, 1.f); h.P.assign((size_t)NLEV * KP, 0.f); h.lev.assign(N, 0);
    srand(5);
    auto unit = [&](float* r, int n) { double ss = 0; for (int j = 0; j < n; ++j) { r[j] = (rand() % 2001 - 1000) / 1000.f; ss += (double)r[j] * r[j]; } for (int j = 0; j < n; ++j) r[j] = (float)(r[j] / sqrt(ss)); };
    for (int k = 0; k < K; ++k) { unit(&h.Y[(size_t)k * DP], D); h.sigma[k] = 0.1f; }
    for (int l = 0; l < NLEV; ++l) for (int k = 0; k < K; ++k) h.P[(size_t)l * KP + k] = 0.3f + (rand() % 1000) / 1000.f;
    for (int i = 0; i < N; ++i) {
        float* z = &h.Zf[(size_t)i * DPF];
        unit(z, D);
        const float* y = &h.Y[(size_t)(hash32(i) % K) * DP];
        double ss = 0;
        for (int j = 0; j < D; ++j) { z[j] = 0.6f * y[j] + 0.4f * z[j]; ss += (double)z[j] * z[j]; }
        for (int j = 0; j < D; ++j) z[j] = (float)(z[j] / sqrt(ss));
        for (int j = 0; j < D; ++j) {
            const float x = z[j] * OPSCALE;
            const __half hi = __float2half_rn(x);
            h.Zs[(size_t)i * 2 * DP + j] = hi;
            h.Zs[(size_t)i * 2 * DP + DP + j] = __float2half_rn(x - __half2float(hi));
        }
        h.lev[i] = (int)(((long long)i * NLEV) / N);            // cells sorted by level, like the engine's layout
    }
    // block lists: CTA g owns [g N / G, (g+1) N / G); block b = cells with hash % NBLK == b, ascending
    h.tile_first.assign(G + 1, 0);
    for (int g = 0; g < G; ++g) {
        const int c0 = (int)((long long)g * N / G), c1 = (int)((long long)(g + 1) * N / G);
        for (int b = 0; b < NBLK; ++b) {
            const int start = (int)h.list.size();
            for (int c = c0; c < c1; ++c) if ((int)(hash32(c * 2654435761u + 17u) % NBLK) == b) h.list.push_back(c);
            const int n = (int)h.list.size() - start;
            for (int o = 0; o < n; o += TILE) {
                h.tile_off.push_back(start + o);
                const int nt = std::min(TILE, n - o);
                h.tile_nt.push_back(nt | ((o + TILE >= n) ? (1 << 30) : 0));
            }
        }
        h.tile_first[g + 1] = (int)h.tile_off.size();
    }
    return h;
}

template <class T> static T* upload(const std::vector<T>& v) { T* d; CK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T))); CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

struct Dev { Args a; float *R, *Ys, *Os, *obj; int G; size_t N; };

static Dev to_device(const Host& h) {
    Dev d{}; d.G = h.G; d.N = h.N;
    d.a.Zf = upload(h.Zf); d.a.Zs = upload(h.Zs); d.a.Y = upload(h.Y); d.a.sigma = upload(h.sigma); d.a.P = upload(h.P);
    d.a.lev = upload(h.lev); d.a.list = upload(h.list); d.a.tile_first = upload(h.tile_first); d.a.tile_off = upload(h.tile_off); d.a.tile_nt = upload(h.tile_nt);
    CK(cudaMalloc(&d.R, (size_t)h.N * KP * 4)); CK(cudaMalloc(&d.Ys, (size_t)h.G * KM * DP * 4)); CK(cudaMalloc(&d.Os, (size_t)h.G * KM * NB * 4)); CK(cudaMalloc(&d.obj, (size_t)h.G * 256 * 2 * 4));
    d.a.R = d.R; d.a.Yslab = d.Ys; d.a.Oslab = d.Os; d.a.obj = d.obj;
    return d;
}
static void free_device(Dev& d) {
    cudaFree((void*)d.a.Zf); cudaFree((void*)d.a.Zs); cudaFree((void*)d.a.Y); cudaFree((void*)d.a.sigma); cudaFree((void*)d.a.P); cudaFree((void*)d.a.lev);
    cudaFree((void*)d.a.list); cudaFree((void*)d.a.tile_first); cudaFree((void*)d.a.tile_off); cudaFree((void*)d.a.tile_nt);
    cudaFree(d.R); cudaFree(d.Ys); cudaFree(d.Os); cudaFree(d.obj);
}

static size_t smem_bytes(int mode) {
    const int NZ = mode == 2 ? 3 : 1;
    return (size_t)NZ * (2 * SZ + SO) + 2 * SR + 2 * SY + (NLEV * KP + 2 * KP) * 4 + 128 * 4 + 2 * 128 * 2 * 4 + 1024;
}
template <int MODE, int NTHR> static void launch(Dev& d) {
    static bool attr = false;
    if (!attr) { CK(cudaFuncSetAttribute(tile_pipeline<MODE, NTHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MODE))); attr = true; }
    tile_pipeline<MODE, NTHR><<<d.G, NTHR, smem_bytes(MODE)>>>(d.a);
}
// variant 0..3: the three modes with 128 threads, then the pipelined mode with 256
static void launch_variant(int v, Dev& d) {
    if (v == 0) launch<0, 128>(d); else if (v == 1) launch<1, 128>(d); else if (v == 2) launch<2, 128>(d); else launch<2, 256>(d);
}
static void zero_slabs(Dev& d) { CK(cudaMemset(d.Ys, 0, (size_t)d.G * KM * DP * 4)); CK(cudaMemset(d.Os, 0, (size_t)d.G * KM * NB * 4)); }

// CPU-only consistency check of the generated work lists (runs in the build container): every cell is in exactly one
// tile, tiles are <= 128 cells of ONE CTA's share in ascending order, the last tile of every block is flagged
static int hostcheck() {
    const int N = 131072, G = 148;
    Host h = make_host(N, G);
    std::vector<int> seen(N, 0);
    int bad = 0, flagged = 0;
    for (int g = 0; g < G; ++g) {
        const int c0 = (int)((long long)g * N / G), c1 = (int)((long long)(g + 1) * N / G);
        for (int t = h.tile_first[g]; t < h.tile_first[g + 1]; ++t) {
            const int nt = h.tile_nt[t] & 0xFFFF, off = h.tile_off[t];
            if (nt < 1 || nt > TILE) ++bad;
            flagged += (h.tile_nt[t] >> 30) & 1;
            for (int i = 0; i < nt; ++i) {
                const int c = h.list[off + i];
                if (c < c0 || c >= c1) ++bad;
                if (i > 0 && c <= h.list[off + i - 1]) ++bad;
                seen[c]++;
            }
        }
    }
    for (int c = 0; c < N; ++c) if (seen[c] != 1) ++bad;
    for (int i = 0; i < N; ++i) if (h.lev[i] < 0 || h.lev[i] >= NLEV) ++bad;
    printf("hostcheck: %d cells, %d tiles, %d block-final tiles (<= %d), %s\n", N, (int)h.tile_off.size(), flagged, G * NBLK, bad ? "FAIL" : "PASS");
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "--hostcheck")) return hostcheck();
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    const char* names[4] = {"v1 (fp32 gather, synchronous)", "presplit cp.async, synchronous", "presplit cp.async, pipelined", "pipelined, 256 threads"};
    int bad = 0;
    const bool time_only = (argc > 1 && !strcmp(argv[1], "--time-only"));
    const int only_mode = (time_only && argc > 2) ? atoi(argv[2]) : -1;
    if (!time_only)
    {   // ---- correctness at a small N (every tile ragged) and at a medium N against fp64
        const int N = (argc > 1) ? atoi(argv[1]) : 131072;
        Host h = make_host(N, G);
        std::vector<double> Rref((size_t)N * KP, 0.0), Yref((size_t)KM * DP, 0.0), Oref((size_t)KM * NB, 0.0);
        for (int i = 0; i < N; ++i) {
            double s[K], ss = 0, sp = 0;
            for (int k = 0; k < K; ++k) {
                double dot = 0;
                for (int j = 0; j < D; ++j) dot += (double)h.Zf[(size_t)i * DPF + j] * h.Y[(size_t)k * DP + j];
                s[k] = exp(-2.0 * (1.0 - dot) / h.sigma[k]); ss += s[k];
            }
            for (int k = 0; k < K; ++k) { s[k] = s[k] / ss * h.P[(size_t)h.lev[i] * KP + k]; sp += s[k]; }
            for (int k = 0; k < K; ++k) {
                const double r = s[k] / std::max(sp, 1e-8);
                Rref[(size_t)i * KP + k] = r;
                for (int j = 0; j < D; ++j) Yref[(size_t)k * DP + j] += r * h.Zf[(size_t)i * DPF + j];
                Oref[(size_t)k * NB + h.lev[i]] += r;
            }
        }
        Dev d = to_device(h);
        for (int mode = 0; mode < 4; ++mode) {
            CK(cudaMemset(d.R, 0xff, (size_t)N * KP * 4));
            zero_slabs(d);
            launch_variant(mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: CUDA error %s\nFAIL\n", mode, cudaGetErrorString(e)); return 1; }
            std::vector<float> R((size_t)N * KP), Ys((size_t)G * KM * DP), Os((size_t)G * KM * NB);
            CK(cudaMemcpy(R.data(), d.R, R.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(Ys.data(), d.Ys, Ys.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(Os.data(), d.Os, Os.size() * 4, cudaMemcpyDeviceToHost));
            double eR = 0, eY = 0, eO = 0, yscale = 0, oscale = 0;
            for (size_t i = 0; i < R.size(); ++i) eR = fmax(eR, fabs((double)R[i] - Rref[i]));
            for (int q = 0; q < KM * DP; ++q) { double s = 0; for (int g = 0; g < G; ++g) s += Ys[(size_t)g * KM * DP + q]; eY = fmax(eY, fabs(s - Yref[q])); yscale = fmax(yscale, fabs(Yref[q])); }
            for (int q = 0; q < KM * NB; ++q) { double s = 0; for (int g = 0; g < G; ++g) s += Os[(size_t)g * KM * NB + q]; eO = fmax(eO, fabs(s - Oref[q])); oscale = fmax(oscale, fabs(Oref[q])); }
            const bool ok = eR < 2e-5 && eY / yscale < 2e-5 && eO / oscale < 2e-5;
            printf("N=%d  %-34s R max|err| %.2e   Y rel %.2e   O rel %.2e   %s\n", N, names[mode], eR, eY / yscale, eO / oscale, ok ? "ok" : "BAD");
            bad += !ok;
        }
        free_device(d);
    }
    if (bad) { printf("FAIL (timings not taken)\n"); return 1; }
    {   // ---- timing at the bench size
        const int N = 1 << 20;
        Host h = make_host(N, G);
        Dev d = to_device(h);
        const int tiles = (int)h.tile_off.size();
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int mode = 0; mode < 4; ++mode) {
            if (only_mode >= 0 && mode != only_mode) continue;
            zero_slabs(d); launch_variant(mode, d); CK(cudaDeviceSynchronize());
            float best = 1e30f, sum = 0;
            for (int it = 0; it < 5; ++it) {
                zero_slabs(d);
                CK(cudaEventRecord(e0));
                launch_variant(mode, d);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms); sum += ms;
            }
            const double bytes = (double)N * (KP * 4 + (mode == 0 ? DPF * 4 : 2 * DP * 2) + 8);
            printf("ABL=%d N=%d  %-34s %.1f us per pass (best %.1f), %.2f us per tile and SM, %.0f GB/s of R + Z traffic\n", ABL, N, names[mode],
                   1e3 * sum / 5, 1e3 * best, 1e3 * (sum / 5) / ((double)tiles / G), bytes / (sum / 5 * 1e-3) / 1e9);
        }
        free_device(d);
    }
    printf("PASS\n");
    return 0;
}
