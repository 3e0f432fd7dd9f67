// Tensor-core version of the k-means round (same algorithm and same grid-level structure as
// hmy_round.cuh; only the per-block cell processing differs).
//
// The two contractions of a round
//     scores  S[cell][k]  = z_cell . y_k                 (harmony.py:447, K x d x N)
//     sums    Y^T[j][k]  += z_cell[j] * R[cell][k]       (harmony.py:443, d x N x K)
// run on the tensor cores as fp16 mma.sync m16n8k16 with a two-way operand split: every fp32
// operand x (scaled by 2^10) is written as hi + lo with hi = fp16(x), lo = fp16(x - hi), and
//     a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo      (3 MMAs, fp32 accumulate)
// which keeps 22 mantissa bits per operand -- the accuracy of an fp32 FMA chain (plain
// fp16/tf32 would put a 1e-3 error into exp(-dist/sigma), see DESIGN.md "Precision").
// The softmax / penalty / objective epilogue works on the accumulator fragments in registers
// with ex2/lg2.approx; the new R rows go to HBM straight from the fragments and, as fp16
// hi/lo, to shared memory where they are the B operand of the second contraction.
//
// CTA = 4*WN warps, tile = 64 cells.  Warp (mw, nh): scoring rows [16 mw, 16 mw + 16) of the
// tile x its n-range of clusters; second contraction: PCs [16 mw, 16 mw + 16) x its n-range.
#pragma once
#include <cuda_fp16.h>
#include "hmy_common.cuh"
#include "hmy_round.cuh"
#include "hmy_xchg.cuh"

#define HMY_MT 64                 // cells per tile
#define HMY_OPSCALE 1024.0f       // operand scale before the fp16 split (2^10)
#define HMY_ACCSCALE (1.0f / 1048576.0f)

struct MmaSmem {
    int ZSH, RSH;                 // row strides (in halves) of the Z/Y and R tiles: (stride/8) odd
    int NTW;                      // n-tiles (8 clusters) per warp
    int KT2;                      // cluster rows of the Y tile (multiple of 16)
    int off_Yh, off_Yl, off_c1, off_c3, off_Ps, off_Os, off_rs, off_prb, off_part, off_rc, off_union, off_misc;
    int off_Zh, off_Zl, off_Rh, off_Rl, off_cell, off_combo, off_lev, off_xch;
    int off_T, off_cnt, off_btot;
    int total;
};

__host__ __device__ inline int hmy_odd8(int halves) { return ((halves / 8) & 1) ? halves : halves + 8; }

__host__ __device__ inline MmaSmem mma_smem_plan(int d, int K, int KS, int B, int V, int nblk, int WN, int NT) {
    MmaSmem s;
    const int dp16 = (d + 15) & ~15;
    // every warp owns exactly NT n-tiles (8 clusters each); clusters beyond K are zero padding
    s.NTW = NT;
    s.KT2 = 8 * NT * WN;
    (void)K;
    s.ZSH = hmy_odd8(dp16);
    s.RSH = hmy_odd8(s.KT2);
    const int NTHR = 128 * WN;
    int o = 0;
    s.off_Yh = o; o += s.KT2 * s.ZSH * 2;
    s.off_Yl = o; o += s.KT2 * s.ZSH * 2;
    s.off_c1 = o; o += s.KT2 * 4;
    s.off_c3 = o; o += s.KT2 * 4;
    s.off_Ps = o; o += B * s.KT2 * 4;
    s.off_Os = o; o += B * s.KT2 * 4;          // running O of the round (every CTA keeps its own copy)
    s.off_rs = o; o += s.KT2 * 4;              // sum_n R[n][k] = sum of O over covariate 0
    s.off_prb = o; o += 2 * B * 4;             // Pr_b | theta
    s.off_part = o; o += 8 * s.KT2 * 4;        // finished runs' column sums: 2 slots per row group
    s.off_rc = o; o += 16 * 4;
    o = (o + 15) & ~15;
    s.off_union = o;
    int a = o;
    s.off_Zh = a; a += HMY_MT * s.ZSH * 2;
    s.off_Zl = a; a += HMY_MT * s.ZSH * 2;
    s.off_Rh = a; a += HMY_MT * s.RSH * 2;
    s.off_Rl = a; a += HMY_MT * s.RSH * 2;
    s.off_cell = a; a += HMY_MT * 4;
    s.off_combo = a; a += HMY_MT * 4;
    s.off_lev = a; a += HMY_MT * V * 4;
    s.off_xch = a; a += 2 * 2 * HMY_MT * 4;
    int b = o;
    s.off_T = b; if (nblk > 32) b += phase0_tables(KS) * nblk * KS * 4;
    s.off_cnt = b; s.off_btot = b;
    (void)NTHR;
    o = (a > b ? a : b);
    o = (o + 15) & ~15;
    s.off_misc = o; o += 8 * 256 + 128;
    s.total = o;
    return s;
}

// n / d for n < 2^24 and 1 <= d <= 256 with m = ceil(2^32 / d): one IMAD.HI instead of a division
// (d = 1 has no 32-bit magic: encoded as 0 and handled by the select)
__host__ __device__ inline unsigned int hmy_magic(int d) { return d <= 1 ? 0u : (unsigned int)((0x100000000ull + (unsigned long long)d - 1ull) / (unsigned long long)d); }
__device__ __forceinline__ int hmy_div(int n, unsigned int magic) { return magic ? (int)__umulhi((unsigned int)n, magic) : n; }

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(unsigned int (&r)[4], unsigned int addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(unsigned int (&r)[4], unsigned int addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned int (&a)[4], unsigned int b0, unsigned int b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// x (already scaled) -> packed fp16 pairs of the hi and lo parts
__device__ __forceinline__ void split2(float x0, float x1, unsigned int& hi, unsigned int& lo) {
    const __half2 H = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(H);
    const __half2 L = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const unsigned int*>(&H);
    lo = *reinterpret_cast<const unsigned int*>(&L);
}

// ---- per-CTA context --------------------------------------------------------------------------
template <int NT, int WN>
struct MmaCtx {
    __half *Yh, *Yl, *Zh, *Zl, *Rh, *Rl;
    float *c1, *c3, *Ps, *xch, *Os, *sRs, *sPrb, *sTheta, *sPart;
    int* sRc;
    int *sCell, *sCombo, *sLev;
    double *sRow, *sRed; int* sFlag;
    int ZSH, RSH, NTW, KT2, dt;         // dt = k16 steps over the PCs
    unsigned int mg_dp4, mg_kp4, mg_K;  // magic divisors
    int n0;                             // first n-tile of this warp
    int ntw;                            // n-tiles this warp really has
    int run_combo;
    int slot_used;                      // this row group already parked one finished run in its first slot
    float colacc[NT][2];                // running batch sums of the warp's rows (this thread's columns)
    float yacc[NT][4];                  // centroid sums Y^T[PC m-tile][cluster n-tiles], whole round
    double objd, obje;
};

template <int NT, int WN>
__device__ __forceinline__ void mma_ctx_init(MmaCtx<NT, WN>& c, const HmyDev& st, unsigned char* smem) {
    const MmaSmem p = mma_smem_plan(st.d, st.K, st.KS, st.B, st.V, st.nblk, WN, NT);
    c.Yh = (__half*)(smem + p.off_Yh); c.Yl = (__half*)(smem + p.off_Yl);
    c.Zh = (__half*)(smem + p.off_Zh); c.Zl = (__half*)(smem + p.off_Zl);
    c.Rh = (__half*)(smem + p.off_Rh); c.Rl = (__half*)(smem + p.off_Rl);
    c.c1 = (float*)(smem + p.off_c1); c.c3 = (float*)(smem + p.off_c3); c.Ps = (float*)(smem + p.off_Ps);
    c.xch = (float*)(smem + p.off_xch);
    c.Os = (float*)(smem + p.off_Os); c.sRs = (float*)(smem + p.off_rs);
    c.sPrb = (float*)(smem + p.off_prb); c.sTheta = c.sPrb + st.B;
    c.sPart = (float*)(smem + p.off_part); c.sRc = (int*)(smem + p.off_rc);
    c.sCell = (int*)(smem + p.off_cell); c.sCombo = (int*)(smem + p.off_combo); c.sLev = (int*)(smem + p.off_lev);
    c.sRow = (double*)(smem + p.off_misc); c.sRed = c.sRow + 256; c.sFlag = (int*)(c.sRed + 8);
    c.ZSH = p.ZSH; c.RSH = p.RSH; c.NTW = p.NTW; c.KT2 = p.KT2;
    c.dt = (st.d + 15) >> 4;
    c.mg_dp4 = hmy_magic(st.dp >> 2); c.mg_kp4 = hmy_magic(st.Kp >> 2); c.mg_K = hmy_magic(st.K);
    const int warp = threadIdx.x >> 5, nh = warp >> 2;
    c.n0 = nh * NT;
    c.ntw = NT;
    c.run_combo = -1; c.slot_used = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        c.colacc[i][0] = c.colacc[i][1] = 0.f;
        c.yacc[i][0] = c.yacc[i][1] = c.yacc[i][2] = c.yacc[i][3] = 0.f;
    }
    c.objd = 0.0; c.obje = 0.0;
}

// centroids as fp16 hi/lo [cluster][PC], per-cluster constants, zeroed tiles
template <int NT, int WN>
__device__ void mma_load_centroids(MmaCtx<NT, WN>& c, const HmyDev& st) {
    const int NTHR = 128 * WN;
    for (int i = threadIdx.x; i < c.KT2 * (c.ZSH / 2); i += NTHR) {
        const int k = i / (c.ZSH / 2), j = 2 * (i - k * (c.ZSH / 2));
        float y0 = 0.f, y1 = 0.f;
        if (k < st.K) {
            if (j < st.dp) y0 = st.Yhat[(size_t)k * st.dp + j] * HMY_OPSCALE;
            if (j + 1 < st.dp) y1 = st.Yhat[(size_t)k * st.dp + j + 1] * HMY_OPSCALE;
        }
        unsigned int hi, lo;
        split2(y0, y1, hi, lo);
        *reinterpret_cast<unsigned int*>(c.Yh + k * c.ZSH + j) = hi;
        *reinterpret_cast<unsigned int*>(c.Yl + k * c.ZSH + j) = lo;
    }
    for (int i = threadIdx.x; i < st.B * c.KT2; i += NTHR) { c.Ps[i] = 0.f; c.Os[i] = 0.f; }
    for (int b = threadIdx.x; b < st.B; b += NTHR) { c.sPrb[b] = st.Pr_b[b]; c.sTheta[b] = st.theta[b]; }
    // t = c2 - acc * c1 is (dist / sigma) * log2(e); c2 = c1 * 2^20; dist = t * c3; sigma ln r = c3 lg2 r
    for (int k = threadIdx.x; k < c.KT2; k += NTHR) {
        const float sg = (k < st.K) ? st.sigma[k] : 1.f;
        c.c1[k] = (k < st.K) ? (2.0f * 1.4426950408889634f / sg) * HMY_ACCSCALE : 0.f;
        c.c3[k] = (k < st.K) ? sg * 0.6931471805599453f : 0.f;
    }
}

template <int NT, int WN>
__device__ void mma_zero_tiles(MmaCtx<NT, WN>& c) {
    const int NTHR = 128 * WN;
    unsigned int* z = reinterpret_cast<unsigned int*>(c.Zh);
    for (int i = threadIdx.x; i < HMY_MT * c.ZSH; i += NTHR) z[i] = 0u;            // Zh and Zl are adjacent
    unsigned int* r = reinterpret_cast<unsigned int*>(c.Rh);
    for (int i = threadIdx.x; i < HMY_MT * c.RSH; i += NTHR) r[i] = 0u;            // Rh and Rl are adjacent
}

// batch sums of a finished run -> Dnew[blk][level][cluster]  (harmony.py:506-507)
template <int NT, int WN>
__device__ __forceinline__ void mma_flush_run(MmaCtx<NT, WN>& c, const HmyDev& st, int blk) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (c.run_combo >= 0) {
        // levels first: a load placed after the first atomic could not be hoisted by the compiler
        const int* lvp = st.combo_lev + c.run_combo * st.V;
        const int lev0 = lvp[0], lev1 = (st.V > 1) ? lvp[1] : 0;
        float* base = st.Dnew + (size_t)blk * st.B * st.K;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v = c.colacc[i][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    const int col = 8 * (c.n0 + i) + 2 * t + e;
                    if (g == 0 && col < st.K && v != 0.f) {
                        atomicAdd(base + (size_t)lev0 * st.K + col, v);
                        atomicAdd(&st.Ofresh[(size_t)lev0 * st.K + col], (double)v);
                        if (st.V > 1) {
                            atomicAdd(base + (size_t)lev1 * st.K + col, v);
                            atomicAdd(&st.Ofresh[(size_t)lev1 * st.K + col], (double)v);
                        }
                        for (int vv = 2; vv < st.V; ++vv) {
                            atomicAdd(base + (size_t)lvp[vv] * st.K + col, v);
                            atomicAdd(&st.Ofresh[(size_t)lvp[vv] * st.K + col], (double)v);
                        }
                    }
                    c.colacc[i][e] = 0.f;
                }
            }
        }
    }
    c.run_combo = -1;
}

// The running column sums of a row group go to shared-memory slot `slot` (2 per row group)
template <int NT, int WN>
__device__ __forceinline__ void mma_park_run(MmaCtx<NT, WN>& c, int slot) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float v = c.colacc[i][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (g == 0) c.sPart[slot * c.KT2 + 8 * (c.n0 + i) + 2 * t + e] = v;
                c.colacc[i][e] = 0.f;
            }
        }
    }
    if (lane == 0 && (threadIdx.x >> 5) < 4) c.sRc[slot] = c.run_combo;   // WN = 2: both column halves share the rows
}

// A run of equal-combination rows ended inside a block: park it (first time) or, if the row
// group already used its spare slot, send it out with per-warp atomics (slow, rare).
template <int NT, int WN>
__device__ __forceinline__ void mma_end_run(MmaCtx<NT, WN>& c, const HmyDev& st, int blk) {
    if (c.run_combo < 0) return;
    const int mw = (threadIdx.x >> 5) & 3;
    if (!c.slot_used) { mma_park_run(c, 2 * mw); c.slot_used = 1; c.run_combo = -1; }
    else mma_flush_run(c, st, blk);
}

// End of a block: every parked run leaves through shared memory so that each thread issues ONE
// atomic per covariate and run (a warp's atomics complete one after the other, ~0.35 us each:
// 26 per warp per block were 10 us on the block's critical path).
template <int NT, int WN>
__device__ void mma_flush_block(MmaCtx<NT, WN>& c, const HmyDev& st, int blk) {
    constexpr int NTHR = 128 * WN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, mw = warp & 3;
    if (!c.slot_used && lane == 0 && warp < 4) c.sRc[2 * mw] = -1;
    mma_park_run(c, 2 * mw + 1);
    c.run_combo = -1; c.slot_used = 0;
    __syncthreads();
    for (int col = tid; col < st.K; col += NTHR) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int cb = c.sRc[w];
            if (cb < 0) continue;
            a += c.sPart[w * c.KT2 + col];
            int nxt = -2;
#pragma unroll
            for (int w2 = 7; w2 > w; --w2) if (c.sRc[w2] >= 0) nxt = c.sRc[w2];   // next occupied slot
            if (nxt != cb) {
                if (a != 0.f) {
                    for (int vv = 0; vv < st.V; ++vv) {
                        const int lev = st.combo_lev[cb * st.V + vv];
                        atomicAdd(&st.Dnew[((size_t)blk * st.B + lev) * st.K + col], a);
                        atomicAdd(&st.Ofresh[(size_t)lev * st.K + col], (double)a);
                    }
                }
                a = 0.f;
            }
        }
    }
    __syncthreads();
}

// Every CTA keeps the running O of the round in shared memory and derives the penalty table
// from it itself (no single-CTA serial section between two blocks): put block blk-1 back
// (harmony.py:506-507), take block blk out (:491-492), E = rowsum x Pr_b (:388/:491),
// P = clamp(E / clamp(O+E))^theta (:495-499).  Identical inputs and order on every CTA.
template <int NT, int WN>
__device__ void mma_update_tables(MmaCtx<NT, WN>& c, const HmyDev& st, int blk, bool combined) {
    constexpr int NTHR = 128 * WN;
    const int n = st.B * st.K, K = st.K, B = st.B, tid = threadIdx.x;
    const float* dn = st.Dnew + (size_t)(blk > 0 ? blk - 1 : 0) * n;
    const float* to = st.Told + (size_t)blk * n;
    const bool use_told = !(combined && blk > 0), use_dn = blk > 0;
    // thread k owns cluster k: one round trip for the whole column, row sum and penalty on the fly
    for (int k = tid; k < K; k += NTHR) {
        float rs = 0.f;
        for (int b0 = 0; b0 < B; b0 += 24) {
            float a[24], r[24];
#pragma unroll
            for (int u = 0; u < 24; ++u) {
                const int b = b0 + u;
                a[u] = (b < B && use_dn) ? __ldcg(&dn[b * K + k]) : 0.f;
                r[u] = (b < B && use_told) ? __ldcg(&to[b * K + k]) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 24; ++u) {
                const int b = b0 + u;
                if (b < B) {
                    const float o = c.Os[b * c.KT2 + k] + (a[u] - r[u]);
                    c.Os[b * c.KT2 + k] = o;
                    if (b < st.lev0) rs += o;
                }
            }
        }
        for (int b = 0; b < B; ++b) {
            const float o = c.Os[b * c.KT2 + k];
            const float e = rs * c.sPrb[b];
            const float ratio = fminf(fmaxf(e / fmaxf(o + e, 1e-8f), 1e-8f), 1.0f);
            const float th = c.sTheta[b];
            c.Ps[b * c.KT2 + k] = (th == 2.0f) ? ratio * ratio : powf(ratio, th);
        }
    }
    __syncthreads();
}

// O of the previous stage (fp64, global) -> this CTA's running copy
template <int NT, int WN>
__device__ void mma_load_O(MmaCtx<NT, WN>& c, const HmyDev& st) {
    constexpr int NTHR = 128 * WN;
    const int n = st.B * st.K, K = st.K;
    for (int i0 = threadIdx.x; i0 < n; i0 += NTHR * 8) {
        double o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * NTHR; o[u] = (i < n) ? __ldcg(&st.O[i]) : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NTHR;
            if (i < n) { const int b = hmy_div(i, c.mg_K), k = i - b * K; c.Os[b * c.KT2 + k] = (float)o[u]; }
        }
    }
}

// ---- phase 0 on the tensor cores -------------------------------------------------------------
// Told[blk][level][k] = sum over this CTA's cells of [cell in blk] * R_old[cell][k]
// (the R_block.sum / R_block @ Phi_block.T of harmony.py:491-492 for all blocks at once) is the
// product onehot(blk)^T (nblk x cells) . R_old (cells x K): A is built in registers from the
// block ids (exact in fp16), B is the fp16 hi/lo split of the R rows, which stream in
// contiguously (cells are stored sorted, a combination segment is one HBM range).
// Needs nblk <= 32 (two m-tiles); larger block counts use the scalar phase0<> above.
template <int NT, int WN>
__device__ void mma_phase0(MmaCtx<NT, WN>& c, const HmyDev& st, long long c0, long long c1) {
    constexpr int NTHR = 128 * WN, W = 4 * WN, NP = NT * WN / 2, PP = (NP + W - 1) / W;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int lj = lane >> 3, lr = lane & 7;
    const int Kp = st.Kp, Kp4 = Kp >> 2, RSH = c.RSH;
    unsigned char* sB = reinterpret_cast<unsigned char*>(c.sCell);       // 64 block ids of the tile
    long long s0 = c0;
    while (s0 < c1) {
        const int combo = st.combo[s0];
        const long long s1 = min(c1, st.combo_start[combo + 1]);
        float acc[2][2 * PP][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 2 * PP; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
        // R rows of a tile: all loads of a thread in one batch, issued one tile ahead of their use
        constexpr int RU = 16;                                         // 64 rows x (Kp4 <= 32 WN) float4 / NTHR
        float4 v[RU];
        unsigned char bnext = 255;
        auto issue = [&](long long tb) {
            const int nt = (int)min((long long)HMY_MT, s1 - tb);
            const float4* src = reinterpret_cast<const float4*>(st.R + (size_t)tb * Kp);
            const int total = nt * Kp4;
#pragma unroll
            for (int u = 0; u < RU; ++u) { const int i = tid + u * NTHR; if (i < total) v[u] = __ldg(src + i); }
            if (tid < HMY_MT) bnext = (tid < nt) ? st.blk[tb + tid] : (unsigned char)255;
        };
        issue(s0);
        for (long long tb = s0; tb < s1; tb += HMY_MT) {
            const int nt = (int)min((long long)HMY_MT, s1 - tb);
            {
                const int total = nt * Kp4;
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int i = tid + u * NTHR;
                    if (i < total) {
                        const int row = hmy_div(i, c.mg_kp4), c4 = i - row * Kp4;
                        uint2 hi, lo;
                        split2(v[u].x * HMY_OPSCALE, v[u].y * HMY_OPSCALE, hi.x, lo.x);
                        split2(v[u].z * HMY_OPSCALE, v[u].w * HMY_OPSCALE, hi.y, lo.y);
                        *reinterpret_cast<uint2*>(c.Rh + row * RSH + 4 * c4) = hi;
                        *reinterpret_cast<uint2*>(c.Rl + row * RSH + 4 * c4) = lo;
                    }
                }
                if (tid < HMY_MT) sB[tid] = bnext;
            }
            if (tb + HMY_MT < s1) issue(tb + HMY_MT);
            __syncthreads();
            const int ksteps = (nt + 15) >> 4;
            for (int ks = 0; ks < ksteps; ++ks) {
                // A[m = block][k = cell]: 1 where the cell belongs to the block
                const unsigned int b01 = *reinterpret_cast<const unsigned short*>(sB + 16 * ks + 2 * t);
                const unsigned int b23 = *reinterpret_cast<const unsigned short*>(sB + 16 * ks + 2 * t + 8);
                const unsigned int k0 = b01 & 255u, k1 = b01 >> 8, k2 = b23 & 255u, k3 = b23 >> 8;
                unsigned int a[2][4];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const unsigned int rl = 16 * m + g, rh = rl + 8;
                    a[m][0] = (k0 == rl ? 0x3C00u : 0u) | (k1 == rl ? 0x3C000000u : 0u);
                    a[m][1] = (k0 == rh ? 0x3C00u : 0u) | (k1 == rh ? 0x3C000000u : 0u);
                    a[m][2] = (k2 == rl ? 0x3C00u : 0u) | (k3 == rl ? 0x3C000000u : 0u);
                    a[m][3] = (k2 == rh ? 0x3C00u : 0u) | (k3 == rh ? 0x3C000000u : 0u);
                }
#pragma unroll
                for (int jp = 0; jp < PP; ++jp) {
                    const int p = warp + W * jp;
                    if (p < NP) {
                        unsigned int bh[4], bl[4];
                        const int brow = 16 * ks + lr + 8 * (lj & 1), bcol = 16 * p + 8 * (lj >> 1);
                        ldsm_x4_t(bh, smem_u32(c.Rh + brow * RSH + bcol));
                        ldsm_x4_t(bl, smem_u32(c.Rl + brow * RSH + bcol));
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            mma_f16(acc[m][2 * jp], a[m], bh[0], bh[1]);
                            mma_f16(acc[m][2 * jp + 1], a[m], bh[2], bh[3]);
                            mma_f16(acc[m][2 * jp], a[m], bl[0], bl[1]);
                            mma_f16(acc[m][2 * jp + 1], a[m], bl[2], bl[3]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // this segment's sums -> Told (rows = blocks, fragment layout)
        int lev[HMY_MAX_V];
#pragma unroll
        for (int vv = 0; vv < HMY_MAX_V; ++vv) lev[vv] = (vv < st.V) ? st.combo_lev[combo * st.V + vv] : 0;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 2 * PP; ++j) {
                const int p = warp + W * (j >> 1);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = 16 * m + g + 8 * (e >> 1), col = 16 * p + 8 * (j & 1) + 2 * t + (e & 1);
                    const float v = acc[m][j][e] * (1.0f / HMY_OPSCALE);
                    if (p < NP && row < st.nblk && col < st.K && v != 0.f) {
#pragma unroll
                        for (int vv = 0; vv < HMY_MAX_V; ++vv)
                            if (vv < st.V) atomicAdd(&st.Told[((size_t)row * st.B + lev[vv]) * st.K + col], v);
                    }
                }
            }
        s0 = s1;
    }
}

// Stage one tile: ids / combination / levels of its cells and their Z_cos rows as fp16 hi/lo.
// Independent of the penalty table, so the first tile of the NEXT block is staged before the
// grid barrier is waited on (its HBM latency overlaps the barrier and the table update).
template <int NT, int WN>
__device__ void mma_stage_tile(MmaCtx<NT, WN>& c, const HmyDev& st, const int* list, long long tb, int nt) {
    constexpr int NTHR = 128 * WN;
    const int tid = threadIdx.x;
    const int dp = st.dp, dp4 = dp >> 2, V = st.V;
    const int ZSH = c.ZSH;
    {
        // ---- stage: cell ids / levels, Z_cos rows -> fp16 hi/lo tile
        if (tid < HMY_MT) {
            int cell = 0, combo = -1;
            if (tid < nt) {
                cell = list ? list[tb + tid] : (int)(tb + tid);
                combo = st.combo[cell];
                for (int v = 0; v < V; ++v) c.sLev[tid * V + v] = st.combo_lev[combo * V + v];
            }
            c.sCell[tid] = cell; c.sCombo[tid] = combo;
        }
        __syncthreads();
        {
            // gather: a batch of loads per thread is issued before its first conversion / store
            constexpr int ZU = 4;
            const int total = nt * dp4;
            for (int base = 0; base < total; base += ZU * NTHR) {
                float4 zr[ZU];
#pragma unroll
                for (int u = 0; u < ZU; ++u) {
                    const int i = base + tid + u * NTHR;
                    if (i < total) {
                        const int row = hmy_div(i, c.mg_dp4), c4 = i - row * dp4;
                        zr[u] = __ldg(reinterpret_cast<const float4*>(st.Zcos + (size_t)c.sCell[row] * dp) + c4);
                    }
                }
#pragma unroll
                for (int u = 0; u < ZU; ++u) {
                    const int i = base + tid + u * NTHR;
                    if (i < total) {
                        const int row = hmy_div(i, c.mg_dp4), c4 = i - row * dp4;
                        uint2 hi, lo;
                        split2(zr[u].x * HMY_OPSCALE, zr[u].y * HMY_OPSCALE, hi.x, lo.x);
                        split2(zr[u].z * HMY_OPSCALE, zr[u].w * HMY_OPSCALE, hi.y, lo.y);
                        *reinterpret_cast<uint2*>(c.Zh + row * ZSH + 4 * c4) = hi;
                        *reinterpret_cast<uint2*>(c.Zl + row * ZSH + 4 * c4) = lo;
                    }
                }
            }
        }
    }
    __syncthreads();
}

// One block of update_R (harmony.py:495-509) for this CTA's cells, or the init assignment
// (harmony.py:380-389) when init = true.  staged_tb: first cell index of a tile that
// mma_stage_tile already prepared (-1: none).
template <int NT, int WN>
__device__ void mma_process_block(MmaCtx<NT, WN>& c, const HmyDev& st, int blk, const int* list,
                                  long long lbeg, long long lend, bool init, long long staged_tb) {
    constexpr int NTHR = 128 * WN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mw = warp & 3, nh = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    const int Kp = st.Kp, K = st.K, V = st.V;
    const int ZSH = c.ZSH, RSH = c.RSH;
    const int lj = lane >> 3, lr = lane & 7;            // ldmatrix: matrix index / row inside it
    (void)NTHR; (void)tid;
    int tslot = 64;
#define HMY_TILE_STAMP() do { if (blk == 5 && !init) hmy_trace(st, tslot < 124 ? tslot++ : 124); } while (0)
    for (long long tb = lbeg; tb < lend; tb += HMY_MT) {
        const int nt = (int)min((long long)HMY_MT, lend - tb);
        HMY_TILE_STAMP();
        if (tb != staged_tb) mma_stage_tile(c, st, list, tb, nt);
        HMY_TILE_STAMP();
        HMY_TILE_STAMP();
        const int row0 = 16 * mw;
        const bool have_rows = row0 < nt;                // warp-uniform
        float acc[NT][4];
        float sp0 = 0.f, sp1 = 0.f;                      // sum S*pen of rows g, g+8 (this warp's columns)
        float ss0 = 0.f, ss1 = 0.f;                      // sum S
        float sd0 = 0.f, sd1 = 0.f;                      // sum S*pen*dist (this thread's columns)
        if (have_rows) {
            // ---- scores: acc[i] = (Z tile rows) x (Y rows of n-tile n0+i)^T, K-dim = PCs
#pragma unroll
            for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            for (int ks = 0; ks < c.dt; ++ks) {
                unsigned int ah[4], al[4];
                const int arow = row0 + lr + 8 * (lj & 1), acol = 16 * ks + 8 * (lj >> 1);
                ldsm_x4(ah, smem_u32(c.Zh + arow * ZSH + acol));
                ldsm_x4(al, smem_u32(c.Zl + arow * ZSH + acol));
#pragma unroll
                for (int ip = 0; ip < NT; ip += 2) {
                    {
                        unsigned int bh[4], bl[4];
                        const int brow = 8 * (c.n0 + ip) + lr + 8 * (lj >> 1), bcol = 16 * ks + 8 * (lj & 1);
                        ldsm_x4(bh, smem_u32(c.Yh + brow * ZSH + bcol));
                        ldsm_x4(bl, smem_u32(c.Yl + brow * ZSH + bcol));
                        constexpr bool two = true;
                        mma_f16(acc[ip], al, bh[0], bh[1]);
                        if (two) mma_f16(acc[ip + 1], al, bh[2], bh[3]);
                        mma_f16(acc[ip], ah, bl[0], bl[1]);
                        if (two) mma_f16(acc[ip + 1], ah, bl[2], bl[3]);
                        mma_f16(acc[ip], ah, bh[0], bh[1]);
                        if (two) mma_f16(acc[ip + 1], ah, bh[2], bh[3]);
                    }
                }
            }
            HMY_TILE_STAMP();
            // ---- S = exp(-dist/sigma) (harmony.py:466-467), times the penalty (harmony.py:500)
            const bool v0 = (row0 + g) < nt, v1 = (row0 + g + 8) < nt;
            const int* lv0 = c.sLev + (v0 ? row0 + g : 0) * V;
            const int* lv1 = c.sLev + (v1 ? row0 + g + 8 : 0) * V;
            const float* pr0 = c.Ps + lv0[0] * c.KT2;
            const float* pr1 = c.Ps + lv1[0] * c.KT2;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                {
                    const int col = 8 * (c.n0 + i) + 2 * t;
                    const float2 k1 = *reinterpret_cast<const float2*>(c.c1 + col);
                    const float2 k3 = *reinterpret_cast<const float2*>(c.c3 + col);
                    float pa0 = 1.f, pb0 = 1.f, pa1 = 1.f, pb1 = 1.f;
                    if (!init) {
                        const float2 q0 = *reinterpret_cast<const float2*>(pr0 + col);
                        const float2 q1 = *reinterpret_cast<const float2*>(pr1 + col);
                        pa0 = q0.x; pb0 = q0.y; pa1 = q1.x; pb1 = q1.y;
                        for (int v = 1; v < V; ++v) {          // more covariates: the factors add (harmony.py:500)
                            const float2 u0 = *reinterpret_cast<const float2*>(c.Ps + lv0[v] * c.KT2 + col);
                            const float2 u1 = *reinterpret_cast<const float2*>(c.Ps + lv1[v] * c.KT2 + col);
                            pa0 += u0.x; pb0 += u0.y; pa1 += u1.x; pb1 += u1.y;
                        }
                    }
                    // t = (dist / sigma) log2 e  >= 0 ; columns beyond K give s = 0
                    const float ta0 = fmaf(-acc[i][0], k1.x, k1.x * 1048576.0f), tb0 = fmaf(-acc[i][1], k1.y, k1.y * 1048576.0f);
                    const float ta1 = fmaf(-acc[i][2], k1.x, k1.x * 1048576.0f), tb1 = fmaf(-acc[i][3], k1.y, k1.y * 1048576.0f);
                    const float sa0 = (col < K) ? ex2_approx(-ta0) : 0.f, sb0 = (col + 1 < K) ? ex2_approx(-tb0) : 0.f;
                    const float sa1 = (col < K) ? ex2_approx(-ta1) : 0.f, sb1 = (col + 1 < K) ? ex2_approx(-tb1) : 0.f;
                    ss0 += sa0 + sb0; ss1 += sa1 + sb1;
                    acc[i][0] = sa0 * pa0; acc[i][1] = sb0 * pb0; acc[i][2] = sa1 * pa1; acc[i][3] = sb1 * pb1;
                    sp0 += acc[i][0] + acc[i][1]; sp1 += acc[i][2] + acc[i][3];
                    // dist = t * c3 (c3 = sigma ln 2): partial sums of S*pen*dist for the objective (harmony.py:399)
                    sd0 += k3.x * (acc[i][0] * ta0) + k3.y * (acc[i][1] * tb0);
                    sd1 += k3.x * (acc[i][2] * ta1) + k3.y * (acc[i][3] * tb1);
                }
            }
            ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
            ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
            sp0 += __shfl_xor_sync(0xffffffffu, sp0, 1); sp0 += __shfl_xor_sync(0xffffffffu, sp0, 2);
            sp1 += __shfl_xor_sync(0xffffffffu, sp1, 1); sp1 += __shfl_xor_sync(0xffffffffu, sp1, 2);
        }
        if (WN == 2) {
            // the two warps that share a row each hold half of the clusters: exchange row sums
            if (have_rows && t == 0) {
                c.xch[(0 * 2 + nh) * HMY_MT + row0 + g] = ss0; c.xch[(0 * 2 + nh) * HMY_MT + row0 + g + 8] = ss1;
                c.xch[(1 * 2 + nh) * HMY_MT + row0 + g] = sp0; c.xch[(1 * 2 + nh) * HMY_MT + row0 + g + 8] = sp1;
            }
            __syncthreads();
            if (have_rows) {
                ss0 = c.xch[0 * HMY_MT + row0 + g] + c.xch[1 * HMY_MT + row0 + g];
                ss1 = c.xch[0 * HMY_MT + row0 + g + 8] + c.xch[1 * HMY_MT + row0 + g + 8];
                sp0 = c.xch[2 * HMY_MT + row0 + g] + c.xch[3 * HMY_MT + row0 + g];
                sp1 = c.xch[2 * HMY_MT + row0 + g + 8] + c.xch[3 * HMY_MT + row0 + g + 8];
            }
        }
        if (have_rows) {
            // R = (S/sumS) pen / max(sum (S/sumS) pen, 1e-8)   (harmony.py:468, :500-503)
            const bool v0 = (row0 + g) < nt, v1 = (row0 + g + 8) < nt;
            const float is0 = 1.f / ss0, is1 = 1.f / ss1;
            const float sc0 = v0 ? is0 / fmaxf(sp0 * is0, 1e-8f) : 0.f;
            const float sc1 = v1 ? is1 / fmaxf(sp1 * is1, 1e-8f) : 0.f;
            float* Rg0 = st.R + (size_t)c.sCell[v0 ? row0 + g : 0] * Kp;
            float* Rg1 = st.R + (size_t)c.sCell[v1 ? row0 + g + 8 : 0] * Kp;
            float oe = 0.f;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                {
                    const int col = 8 * (c.n0 + i) + 2 * t;
                    const float2 k3 = *reinterpret_cast<const float2*>(c.c3 + col);
                    const float ra0 = acc[i][0] * sc0, rb0 = acc[i][1] * sc0;
                    const float ra1 = acc[i][2] * sc1, rb1 = acc[i][3] * sc1;
                    acc[i][0] = ra0; acc[i][1] = rb0; acc[i][2] = ra1; acc[i][3] = rb1;
                    // entropy term sigma * r * ln r  (harmony.py:402, :572-576)
                    oe += k3.x * ((ra0 > 0.f ? ra0 * lg2_approx(ra0) : 0.f) + (ra1 > 0.f ? ra1 * lg2_approx(ra1) : 0.f))
                        + k3.y * ((rb0 > 0.f ? rb0 * lg2_approx(rb0) : 0.f) + (rb1 > 0.f ? rb1 * lg2_approx(rb1) : 0.f));
                    if (col < Kp) {
                        if (v0) *reinterpret_cast<float2*>(Rg0 + col) = make_float2(ra0, rb0);
                        if (v1) *reinterpret_cast<float2*>(Rg1 + col) = make_float2(ra1, rb1);
                    }
                    unsigned int h0, l0, h1, l1;
                    split2(ra0 * HMY_OPSCALE, rb0 * HMY_OPSCALE, h0, l0);
                    split2(ra1 * HMY_OPSCALE, rb1 * HMY_OPSCALE, h1, l1);
                    *reinterpret_cast<unsigned int*>(c.Rh + (row0 + g) * RSH + col) = h0;
                    *reinterpret_cast<unsigned int*>(c.Rl + (row0 + g) * RSH + col) = l0;
                    *reinterpret_cast<unsigned int*>(c.Rh + (row0 + g + 8) * RSH + col) = h1;
                    *reinterpret_cast<unsigned int*>(c.Rl + (row0 + g + 8) * RSH + col) = l1;
                }
            }
            c.objd += (double)(sc0 * sd0 + sc1 * sd1);          // sum R * dist over this thread's entries
            c.obje += (double)oe;
            // ---- batch sums of the new assignments, run-length over combos (rows are sorted)
            const int rhi = min(nt, row0 + 16);
            int r = row0;
            while (r < rhi) {
                const int cb = c.sCombo[r];
                int e = rhi;
                if (c.sCombo[rhi - 1] != cb) { e = r + 1; while (e < rhi && c.sCombo[e] == cb) ++e; }
                if (cb != c.run_combo) { mma_end_run(c, st, blk); c.run_combo = cb; }
                const bool in0 = (row0 + g) >= r && (row0 + g) < e, in1 = (row0 + g + 8) >= r && (row0 + g + 8) < e;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    {
                        c.colacc[i][0] += (in0 ? acc[i][0] : 0.f) + (in1 ? acc[i][2] : 0.f);
                        c.colacc[i][1] += (in0 ? acc[i][1] : 0.f) + (in1 ? acc[i][3] : 0.f);
                    }
                }
                r = e;
            }
        } else {
            // rows of this warp are beyond the tile: their R must read as zero in the second contraction
            for (int i = lane; i < 16 * (RSH / 2); i += 32) {
                const int rr = i / (RSH / 2), cc = 2 * (i - rr * (RSH / 2));
                if (cc >= 8 * c.n0 && cc < 8 * (c.n0 + NT)) {
                    *reinterpret_cast<unsigned int*>(c.Rh + (row0 + rr) * RSH + cc) = 0u;
                    *reinterpret_cast<unsigned int*>(c.Rl + (row0 + rr) * RSH + cc) = 0u;
                }
            }
        }
        HMY_TILE_STAMP();
        __syncthreads();
        HMY_TILE_STAMP();
        // ---- Y^T[PC][cluster] += Z^T R over the cells of the tile (harmony.py:443), K-dim = cells
        if (mw < c.dt) {
            const int ksteps = (nt + 15) >> 4;
            for (int ks = 0; ks < ksteps; ++ks) {
                unsigned int ah[4], al[4];
                // A = Z^T: A[m = PC][k = cell] read transposed from the [cell][PC] tile
                const int arow = 16 * ks + lr + 8 * (lj >> 1), acol = 16 * mw + 8 * (lj & 1);
                ldsm_x4_t(ah, smem_u32(c.Zh + arow * ZSH + acol));
                ldsm_x4_t(al, smem_u32(c.Zl + arow * ZSH + acol));
#pragma unroll
                for (int ip = 0; ip < NT; ip += 2) {
                    {
                        unsigned int bh[4], bl[4];
                        // B[k = cell][n = cluster] read transposed from the [cell][cluster] tile
                        const int brow = 16 * ks + lr + 8 * (lj & 1), bcol = 8 * (c.n0 + ip) + 8 * (lj >> 1);
                        ldsm_x4_t(bh, smem_u32(c.Rh + brow * RSH + bcol));
                        ldsm_x4_t(bl, smem_u32(c.Rl + brow * RSH + bcol));
                        constexpr bool two = true;
                        mma_f16(c.yacc[ip], al, bh[0], bh[1]);
                        if (two) mma_f16(c.yacc[ip + 1], al, bh[2], bh[3]);
                        mma_f16(c.yacc[ip], ah, bl[0], bl[1]);
                        if (two) mma_f16(c.yacc[ip + 1], ah, bl[2], bl[3]);
                        mma_f16(c.yacc[ip], ah, bh[0], bh[1]);
                        if (two) mma_f16(c.yacc[ip + 1], ah, bh[2], bh[3]);
                    }
                }
            }
        }
        HMY_TILE_STAMP();
        __syncthreads();
    }
    HMY_TILE_STAMP();
    mma_flush_block(c, st, blk);
    HMY_TILE_STAMP();
}

// centroid sums (fragment layout: rows = PCs 16 mw + g (+8), cols = clusters) and objective sums
template <int NT, int WN>
__device__ void mma_flush_round_sums(MmaCtx<NT, WN>& c, const HmyDev& st) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, mw = warp & 3, g = lane >> 2, t = lane & 3;
    const double sc = (double)HMY_ACCSCALE;
    if (mw < c.dt) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 16 * mw + g + 8 * (e >> 1), k = 8 * (c.n0 + i) + 2 * t + (e & 1);
                    if (j < st.d && k < st.K && c.yacc[i][e] != 0.f)
                        atomicAdd(&st.Yacc[(size_t)k * st.dp + j], (double)c.yacc[i][e] * sc);
                    c.yacc[i][e] = 0.f;
                }
            }
        }
    }
    double a = warp_sum_d(c.objd), b = warp_sum_d(c.obje);
    if (lane == 0) { atomicAdd(&st.obj[0], a); atomicAdd(&st.obj[1], b); }
    c.objd = 0.0; c.obje = 0.0;
}

template <int NT, int WN>
__device__ __forceinline__ void mma_load_penalty(MmaCtx<NT, WN>& c, const HmyDev& st) {
    // loads first, shared-memory stores after: the compiler may not hoist .cg loads over stores
    constexpr int NTHR = 128 * WN;
    const int n = st.B * st.K;
    for (int i0 = threadIdx.x; i0 < n; i0 += NTHR * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * NTHR; v[u] = (i < n) ? __ldcg(&st.P[i]) : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NTHR;
            if (i < n) { const int b = i / st.K, k = i - b * st.K; c.Ps[b * c.KT2 + k] = v[u]; }
        }
    }
}

// ---- kernels ---------------------------------------------------------------------------------
// ---- fused multi-GPU mode: a dedicated communication CTA ------------------------------------------
// With more than one GPU the LAST CTA of the grid does no cell work: it waits until all worker
// CTAs have arrived at a barrier, exchanges the K x B table with the other GPUs (hmy_xchg.cuh)
// and releases the workers.  Keeping that code out of the worker path matters: inlined into the
// worker's barrier it cost the 255-register kernel 40 % of its speed in spills.
__device__ __forceinline__ void worker_barrier(const HmyDev& st, unsigned int gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(st.bar_count, 1u);
        while ((int)(ld_acquire_u32(st.bar_gen) - gen) < 0) { __nanosleep(20); }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void comm_wait_workers(const HmyDev& st, unsigned int nworkers) {
    if (threadIdx.x == 0) {
        while (ld_acquire_u32(st.bar_count) < nworkers) { __nanosleep(20); }
        *st.bar_count = 0u;
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void comm_release(const HmyDev& st, unsigned int gen) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); st_release_u32(st.bar_gen, gen); }
}

static __device__ __noinline__ void comm_cta_main(const HmyDev& st, int mode, unsigned int gen_base, unsigned int GW,
                                           double* sRow, double* sRed) {
    const int n = st.B * st.K, nend = 4 + n + st.K * st.dp;
    unsigned int gen = gen_base + 1u, xs = st.xseq_base;
    if (mode == 1) {
        comm_wait_workers(st, GW);
        xchg_allreduce<double>(st, st.obj, nend, ++xs);          // objective sums | Ofresh | Yacc
        serial_finalize(st, 1, sRow, sRed);
        comm_release(st, gen);
        return;
    }
    const bool exact = !st.xrelaxed;
    comm_wait_workers(st, GW);                                   // all local Told sums are in
    if (exact) xchg_allreduce_ll_f32(st, st.Told, n, ++xs);      // block 0's removed sums
    comm_release(st, gen++);
    for (int blk = 0; blk < st.nblk; ++blk) {
        comm_wait_workers(st, GW);
        if (blk + 1 < st.nblk) {
            if (exact) {
                // one K x B table per block: (re-added sums of blk) - (removed sums of blk + 1)
                float* dnew = st.Dnew + (size_t)blk * n;
                const float* told = st.Told + (size_t)(blk + 1) * n;
                for (int i0 = threadIdx.x; i0 < n; i0 += 8 * blockDim.x) {
                    float a[8], r[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; a[u] = (i < n) ? __ldcg(&dnew[i]) : 0.f; r[u] = (i < n) ? __ldcg(&told[i]) : 0.f; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; if (i < n) __stcg(&dnew[i], a[u] - r[u]); }
                }
                __syncthreads();
                xchg_allreduce_ll_f32(st, dnew, n, ++xs);
            }
        } else {
            xchg_allreduce<double>(st, st.obj, nend, ++xs);
            serial_finalize(st, 0, sRow, sRed);
        }
        comm_release(st, gen++);
    }
}

// FUSED = true is the multi-GPU instantiation (the single-GPU kernel carries none of that code).
template <int NT, int WN, bool FUSED>
__global__ void __launch_bounds__(128 * WN, (WN == 2 && NT <= 8) ? 2 : 1) k_round_mma(HmyDev st, int mode, unsigned int gen_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NTHR = 128 * WN;
    const MmaSmem p = mma_smem_plan(st.d, st.K, st.KS, st.B, st.V, st.nblk, WN, NT);
    const bool multi_any = FUSED && st.xworld > 1;
    const unsigned int G = multi_any ? gridDim.x - 1u : gridDim.x;          // worker CTAs
    if (multi_any && blockIdx.x == G) {
        double* sRow = (double*)(smem + p.off_misc);
        comm_cta_main(st, mode, gen_base, G, sRow, sRow + 256);
        return;
    }
    MmaCtx<NT, WN> c;
    mma_ctx_init(c, st, smem);
    const long long c0 = (long long)blockIdx.x * st.N / G, c1 = (long long)(blockIdx.x + 1) * st.N / G;
    hmy_trace(st, 0);
    mma_load_centroids(c, st);
    __syncthreads();
    if (mode == 1) {
        mma_zero_tiles(c);
        __syncthreads();
        mma_process_block(c, st, 0, nullptr, c0, c1, true, -1);
        mma_flush_round_sums(c, st);
        if (multi_any) worker_barrier(st, gen_base + 1u);
        else grid_barrier_serial(st, G, gen_base + 1u, c.sFlag, [&]() { serial_finalize(st, 1, c.sRow, c.sRed); });
        return;
    }
    mma_zero_tiles(c);
    __syncthreads();
    if (st.nblk <= 32) {
        mma_phase0(c, st, c0, c1);
    } else {
        phase0<NTHR>(Phase0Mem{(float*)(smem + p.off_T), (unsigned int*)(smem + p.off_cnt), (int*)(smem + p.off_btot), st.KS},
                     st, c0, c1);
        mma_zero_tiles(c);
        __syncthreads();
    }
    hmy_trace(st, 1);
    unsigned int gen = gen_base + 1u;
    mma_load_O(c, st);                       // O is only written by the finalize of the previous launch
    long long staged = -1;
    {
        long long nb, ne;
        block_share(st, 0, blockIdx.x, G, nb, ne);
        if (nb < ne) { mma_stage_tile(c, st, st.list, nb, (int)min((long long)HMY_MT, ne - nb)); staged = nb; }
    }
    const bool multi = multi_any && !st.xrelaxed;      // exact mode: one table exchange per block
    if (multi_any) worker_barrier(st, gen++);          // comm CTA: all Told sums are in (+ exchange)
    else grid_barrier(st, G, gen++);
    hmy_trace(st, 2);
    for (int blk = 0; blk < st.nblk; ++blk) {
        mma_update_tables(c, st, blk, multi);
        hmy_trace(st, 3 + 3 * blk);
        long long lb, le;
        block_share(st, blk, blockIdx.x, G, lb, le);
        mma_process_block(c, st, blk, st.list, lb, le, false, staged);
        hmy_trace(st, 4 + 3 * blk);
        staged = -1;
        if (blk + 1 < st.nblk) {        // next block's first tile: stage before waiting at the barrier
            long long nb, ne;
            block_share(st, blk + 1, blockIdx.x, G, nb, ne);
            if (nb < ne) { mma_stage_tile(c, st, st.list, nb, (int)min((long long)HMY_MT, ne - nb)); staged = nb; }
            if (multi_any) worker_barrier(st, gen++);
            else grid_barrier(st, G, gen++);
        } else {
            mma_flush_round_sums(c, st);
            if (multi_any) worker_barrier(st, gen++);
            else grid_barrier_serial(st, G, gen++, c.sFlag, [&]() { serial_finalize(st, 0, c.sRow, c.sRed); });
        }
        hmy_trace(st, 5 + 3 * blk);
    }
}

template <int NT, int WN>
__global__ void __launch_bounds__(128 * WN, (WN == 2 && NT <= 8) ? 2 : 1) k_round_mma_stage(HmyDev st, int what, int blk) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NTHR = 128 * WN;
    MmaCtx<NT, WN> c;
    mma_ctx_init(c, st, smem);
    const MmaSmem p = mma_smem_plan(st.d, st.K, st.KS, st.B, st.V, st.nblk, WN, NT);
    const unsigned int G = gridDim.x;
    const long long c0 = (long long)blockIdx.x * st.N / G, c1 = (long long)(blockIdx.x + 1) * st.N / G;
    if (what == 0) {
        if (st.nblk <= 32) {
            mma_zero_tiles(c);
            __syncthreads();
            mma_phase0(c, st, c0, c1);
        } else {
            phase0<NTHR>(Phase0Mem{(float*)(smem + p.off_T), (unsigned int*)(smem + p.off_cnt), (int*)(smem + p.off_btot), st.KS},
                         st, c0, c1);
        }
        return;
    }
    mma_load_centroids(c, st);
    mma_zero_tiles(c);
    if (what == 1) mma_load_penalty(c, st);
    __syncthreads();
    if (what == 1) {
        long long lb, le;
        block_share(st, blk, blockIdx.x, G, lb, le);
        mma_process_block(c, st, blk, st.list, lb, le, false, -1);
    } else {
        mma_process_block(c, st, 0, nullptr, c0, c1, true, -1);
    }
    mma_flush_round_sums(c, st);
}
