// moe_correct_ridge (harmony.py:535-569) on the tensor cores: the two streaming passes of
// hmy_ridge.cuh with their contractions as fp16 two-way-split mma.sync (see hmy_round_mma.cuh
// for the split and the fragment conventions).  Used for d <= 63 and K <= 256.
//
//   k_ridge_moments_mma : per combination segment (contiguous cells)
//                            Mom^T[j][k] += z[cell][j] * R[cell][k]   (harmony.py:556-563)
//                            G[k]        += R[cell][k]                (harmony.py:547-550)
//                         one GEMM: Z gets an extra all-ones column, so G is row d of Mom^T.
//   k_ridge_apply_mma   : corr = R . Wc  (Wc = sum over the combination's levels of W),
//                         Z_corr = Z_orig - corr (harmony.py:566), Z_cos = unit rows (:569),
//                         centroid sums Z_cos^T R for the next cluster() (harmony.py:443).
// Operand scales: R x 2^10, Z_orig x zscale, Wc x wscale (powers of two chosen from max |.| so
// that the fp16 hi part stays below 2^14), Z_cos x 2^10.
#pragma once
#include "hmy_round_mma.cuh"
#include "hmy_ridge.cuh"

struct RidgeMmaSmem { int ZSH, RSH, KT2, off_Zh, off_Zl, off_Rh, off_Rl, off_Wh, off_Wl, total; };

__host__ __device__ inline RidgeMmaSmem ridge_mma_plan(int d, int WN, int NT, bool apply) {
    RidgeMmaSmem s;
    const int dpa = (d + 1 + 15) & ~15;               // PCs + the ones column, in k16 / m16 units
    s.KT2 = 8 * NT * WN;
    s.ZSH = hmy_odd8(dpa);
    s.RSH = hmy_odd8(s.KT2);
    int o = 0;
    s.off_Zh = o; o += HMY_MT * s.ZSH * 2;
    s.off_Zl = o; o += HMY_MT * s.ZSH * 2;
    s.off_Rh = o; o += HMY_MT * s.RSH * 2;
    s.off_Rl = o; o += HMY_MT * s.RSH * 2;
    s.off_Wh = o; if (apply) o += dpa * s.RSH * 2;    // Wc^T as [PC][cluster]
    s.off_Wl = o; if (apply) o += dpa * s.RSH * 2;
    s.total = o;
    return s;
}

// contiguous rows [base, base+nt): fp32 rows (stride `ld` floats, `n4` float4 per row) -> scaled fp16 hi/lo tile
template <int NTHR>
__device__ __forceinline__ void ridge_tile_to_half(const float* src, int ld4, int n4, int nt, float scale,
                                                   __half* Th, __half* Tl, int SH) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const int total = nt * n4;
    const unsigned int mg = hmy_magic(n4);
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * NTHR) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NTHR;
            if (i < total) { const int row = hmy_div(i, mg), c4 = i - row * n4; v[u] = __ldg(s4 + (size_t)row * ld4 + c4); }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NTHR;
            if (i < total) {
                const int row = hmy_div(i, mg), c4 = i - row * n4;
                uint2 hi, lo;
                split2(v[u].x * scale, v[u].y * scale, hi.x, lo.x);
                split2(v[u].z * scale, v[u].w * scale, hi.y, lo.y);
                *reinterpret_cast<uint2*>(Th + row * SH + 4 * c4) = hi;
                *reinterpret_cast<uint2*>(Tl + row * SH + 4 * c4) = lo;
            }
        }
    }
}

// The two passes work tile by tile (load, split, multiply): nothing of a CTA's next tile is in flight while it multiplies.
// Asking L2 for the NEXT work item's rows (<= HMY_SEG_MAX of R and of Z_orig, contiguous) when an item starts keeps HBM
// busy during those phases and turns the next tiles' loads into L2 hits.  One bulk prefetch per HMY_MT rows, issued by
// the first threads of the CTA; addresses and sizes are multiples of 16 bytes (rows are whole float4s).
__device__ __forceinline__ void ridge_prefetch_item(const HmyDev& st, int it, int i1) {
    if (it >= i1) return;
    const long long start = st.seg[3 * it]; const int count = st.seg[3 * it + 1];
    const int tiles = (count + HMY_MT - 1) / HMY_MT;
    if ((int)threadIdx.x < 2 * tiles) {
        const int tl = threadIdx.x >> 1, rows = min(HMY_MT, count - tl * HMY_MT);
        const float* p; unsigned int bytes;
        if (threadIdx.x & 1) { p = st.Zorig + (size_t)(start + tl * HMY_MT) * st.dp; bytes = (unsigned int)(rows * st.dp * 4); }
        else { p = st.R + (size_t)(start + tl * HMY_MT) * st.Kp; bytes = (unsigned int)(rows * st.Kp * 4); }
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
    }
}

// yacc[i] += (Z tile)^T (R tile) for PC m-tile `mw`, cluster n-tiles [n0, n0 + NT)  (K-dim = cells)
template <int NT>
__device__ __forceinline__ void ridge_ztr(float (&yacc)[NT][4], const __half* Zh, const __half* Zl, int ZSH,
                                          const __half* Rh, const __half* Rl, int RSH, int mw, int n0, int nt) {
    const int lane = threadIdx.x & 31, lj = lane >> 3, lr = lane & 7;
    const int ksteps = (nt + 15) >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        unsigned int ah[4], al[4];
        const int arow = 16 * ks + lr + 8 * (lj >> 1), acol = 16 * mw + 8 * (lj & 1);
        ldsm_x4_t(ah, smem_u32(Zh + arow * ZSH + acol));
        ldsm_x4_t(al, smem_u32(Zl + arow * ZSH + acol));
#pragma unroll
        for (int ip = 0; ip < NT; ip += 2) {
            unsigned int bh[4], bl[4];
            const int brow = 16 * ks + lr + 8 * (lj & 1), bcol = 8 * (n0 + ip) + 8 * (lj >> 1);
            ldsm_x4_t(bh, smem_u32(Rh + brow * RSH + bcol));
            ldsm_x4_t(bl, smem_u32(Rl + brow * RSH + bcol));
            mma_f16(yacc[ip], al, bh[0], bh[1]);
            mma_f16(yacc[ip + 1], al, bh[2], bh[3]);
            mma_f16(yacc[ip], ah, bl[0], bl[1]);
            mma_f16(yacc[ip + 1], ah, bl[2], bl[3]);
            mma_f16(yacc[ip], ah, bh[0], bh[1]);
            mma_f16(yacc[ip + 1], ah, bh[2], bh[3]);
        }
    }
}

// ---- pass 1 ------------------------------------------------------------------------------------
template <int NT, int WN>
__device__ void ridge_mma_flush_moments(const HmyDev& st, int combo, float (&yacc)[NT][4], int mw, int n0, double inv) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int V = st.V, n1 = st.B + 1;
    int lev[HMY_MAX_V];
#pragma unroll
    for (int v = 0; v < HMY_MAX_V; ++v) lev[v] = (v < V) ? st.combo_lev[combo * V + v] : 0;
    const double ginv = 1.0 / (double)HMY_OPSCALE;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 16 * mw + g + 8 * (e >> 1), k = 8 * (n0 + i) + 2 * t + (e & 1);
            const float a = yacc[i][e];
            if (k < st.K && a != 0.f) {
                if (j < st.d) {
                    const double m = (double)a * inv;
                    atomicAdd(&st.Mom[((size_t)0 * st.K + k) * st.dp + j], m);
#pragma unroll
                    for (int v = 0; v < HMY_MAX_V; ++v)
                        if (v < V) atomicAdd(&st.Mom[((size_t)(1 + lev[v]) * st.K + k) * st.dp + j], m);
                } else if (j == st.d) {          // the ones column: sum_n R[n][k] over the segment
                    const double gk = (double)a * ginv;
                    double* A = st.Gram + (size_t)k * n1 * n1;
                    atomicAdd(&A[0], gk);
#pragma unroll
                    for (int v = 0; v < HMY_MAX_V; ++v) {
                        if (v < V) {
                            atomicAdd(&A[1 + lev[v]], gk);
                            atomicAdd(&A[(size_t)(1 + lev[v]) * n1], gk);
#pragma unroll
                            for (int u = 0; u < HMY_MAX_V; ++u)
                                if (u < V) atomicAdd(&A[(size_t)(1 + lev[v]) * n1 + 1 + lev[u]], gk);
                        }
                    }
                }
            }
            yacc[i][e] = 0.f;
        }
    }
}

template <int NT, int WN>
__global__ void __launch_bounds__(128 * WN) k_ridge_moments_mma(HmyDev st, float zscale) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NTHR = 128 * WN;
    const RidgeMmaSmem p = ridge_mma_plan(st.d, WN, NT, false);
    __half* Zh = (__half*)(smem + p.off_Zh); __half* Zl = (__half*)(smem + p.off_Zl);
    __half* Rh = (__half*)(smem + p.off_Rh); __half* Rl = (__half*)(smem + p.off_Rl);
    const int tid = threadIdx.x, warp = tid >> 5, mw = warp & 3, nh = warp >> 2, n0 = nh * NT;
    const int mtiles = (st.d + 1 + 15) >> 4;
    for (int i = tid; i < HMY_MT * p.ZSH; i += NTHR) reinterpret_cast<unsigned int*>(Zh)[i] = 0u;   // Zh and Zl
    for (int i = tid; i < HMY_MT * p.RSH; i += NTHR) reinterpret_cast<unsigned int*>(Rh)[i] = 0u;   // Rh and Rl
    __syncthreads();
    float yacc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) yacc[i][0] = yacc[i][1] = yacc[i][2] = yacc[i][3] = 0.f;
    const double inv = 1.0 / ((double)HMY_OPSCALE * (double)zscale);
    int cur = -1;
    const int i0 = (int)((long long)blockIdx.x * st.nseg / gridDim.x), i1 = (int)((long long)(blockIdx.x + 1) * st.nseg / gridDim.x);
    ridge_prefetch_item(st, i0, i1);
    for (int it = i0; it < i1; ++it) {
        const long long start = st.seg[3 * it]; const int count = st.seg[3 * it + 1], combo = st.seg[3 * it + 2];
        ridge_prefetch_item(st, it + 1, i1);
        if (combo != cur) { if (cur >= 0 && mw < mtiles) ridge_mma_flush_moments<NT, WN>(st, cur, yacc, mw, n0, inv); cur = combo; }
        for (int tb = 0; tb < count; tb += HMY_MT) {
            const int nt = min(HMY_MT, count - tb);
            const long long base = start + tb;
            ridge_tile_to_half<NTHR>(st.Zorig + (size_t)base * st.dp, st.dp >> 2, st.dp >> 2, nt, zscale, Zh, Zl, p.ZSH);
            __syncthreads();
            // the ones column (PC index d, inside the padded row): 1.0 in the hi part, 0 in lo
            for (int r = tid; r < HMY_MT; r += NTHR) { Zh[r * p.ZSH + st.d] = __float2half_rn(1.0f); Zl[r * p.ZSH + st.d] = __float2half_rn(0.0f); }
            ridge_tile_to_half<NTHR>(st.R + (size_t)base * st.Kp, st.Kp >> 2, st.Kp >> 2, nt, HMY_OPSCALE, Rh, Rl, p.RSH);
            if (nt < HMY_MT) {        // rows beyond the tile must not contribute: zero their R
                for (int i = tid; i < (HMY_MT - nt) * (p.RSH / 2); i += NTHR) {
                    const int rr = nt + i / (p.RSH / 2), cc = 2 * (i % (p.RSH / 2));
                    *reinterpret_cast<unsigned int*>(Rh + rr * p.RSH + cc) = 0u;
                    *reinterpret_cast<unsigned int*>(Rl + rr * p.RSH + cc) = 0u;
                }
            }
            __syncthreads();
            if (mw < mtiles) ridge_ztr<NT>(yacc, Zh, Zl, p.ZSH, Rh, Rl, p.RSH, mw, n0, nt);
            __syncthreads();
        }
    }
    if (cur >= 0 && mw < mtiles) ridge_mma_flush_moments<NT, WN>(st, cur, yacc, mw, n0, inv);
}

// ---- pass 2 ------------------------------------------------------------------------------------
template <int NT, int WN>
__global__ void __launch_bounds__(128 * WN) k_ridge_apply_mma(HmyDev st, const float* wmax_ptr) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NTHR = 128 * WN;
    constexpr int ND = 8;                         // PC n-tiles of the correction (d <= 63 -> 64 columns)
    const RidgeMmaSmem p = ridge_mma_plan(st.d, WN, NT, true);
    __half* Zh = (__half*)(smem + p.off_Zh); __half* Zl = (__half*)(smem + p.off_Zl);
    __half* Rh = (__half*)(smem + p.off_Rh); __half* Rl = (__half*)(smem + p.off_Rl);
    __half* Wh = (__half*)(smem + p.off_Wh); __half* Wl = (__half*)(smem + p.off_Wl);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, mw = warp & 3, nh = warp >> 2, n0 = nh * NT;
    const int g = lane >> 2, t = lane & 3, lj = lane >> 3, lr = lane & 7;
    const int ZSH = p.ZSH, RSH = p.RSH, dp = st.dp;
    const int dpa = (st.d + 1 + 15) & ~15, mtiles = dpa >> 4;
    const int kst = p.KT2 >> 4;                   // k16 steps over the clusters
    // scale of the coefficient tile: largest |W| of the solve -> hi part below 2^14
    float wscale;
    { int e; const float wm = fmaxf(__ldg(wmax_ptr), 1e-30f); frexpf(wm, &e); wscale = ldexpf(1.0f, 13 - e); }
    const float cinv = 1.0f / (HMY_OPSCALE * wscale);
    const int zs_pairs = 8 * ((st.d + 15) >> 4);                  // u32 (= PC pairs) per hi / lo part of a Zs16 row
    unsigned int* const zs_rows = reinterpret_cast<unsigned int*>(st.Zs16);
    for (int i = tid; i < HMY_MT * ZSH; i += NTHR) reinterpret_cast<unsigned int*>(Zh)[i] = 0u;
    for (int i = tid; i < HMY_MT * RSH; i += NTHR) reinterpret_cast<unsigned int*>(Rh)[i] = 0u;
    for (int i = tid; i < dpa * RSH; i += NTHR) reinterpret_cast<unsigned int*>(Wh)[i] = 0u;       // Wh and Wl
    __syncthreads();
    float yacc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) yacc[i][0] = yacc[i][1] = yacc[i][2] = yacc[i][3] = 0.f;
    int cur = -1;
    const int i0 = (int)((long long)blockIdx.x * st.nseg / gridDim.x), i1 = (int)((long long)(blockIdx.x + 1) * st.nseg / gridDim.x);
    ridge_prefetch_item(st, i0, i1);
    for (int it = i0; it < i1; ++it) {
        const long long start = st.seg[3 * it]; const int count = st.seg[3 * it + 1], combo = st.seg[3 * it + 2];
        ridge_prefetch_item(st, it + 1, i1);
        if (combo != cur) {
            // Wc^T[j][k] = sum_v W[level_v][k][j]: what W.T @ Phi_Rk selects for this combination
            __syncthreads();
            for (int i = tid; i < st.K * (dp >> 1); i += NTHR) {
                const int k = i / (dp >> 1), j = 2 * (i - k * (dp >> 1));
                float w0 = 0.f, w1 = 0.f;
                for (int v = 0; v < st.V; ++v) {
                    const float2 w = *reinterpret_cast<const float2*>(st.W + ((size_t)st.combo_lev[combo * st.V + v] * st.K + k) * dp + j);
                    w0 += w.x; w1 += w.y;
                }
                const __half h0 = __float2half_rn(w0 * wscale), h1 = __float2half_rn(w1 * wscale);
                Wh[j * RSH + k] = h0; Wh[(j + 1) * RSH + k] = h1;
                Wl[j * RSH + k] = __float2half_rn(w0 * wscale - __half2float(h0));
                Wl[(j + 1) * RSH + k] = __float2half_rn(w1 * wscale - __half2float(h1));
            }
            cur = combo;
            __syncthreads();
        }
        for (int tb = 0; tb < count; tb += HMY_MT) {
            const int nt = min(HMY_MT, count - tb);
            const long long base = start + tb;
            ridge_tile_to_half<NTHR>(st.R + (size_t)base * st.Kp, st.Kp >> 2, st.Kp >> 2, nt, HMY_OPSCALE, Rh, Rl, RSH);
            if (nt < HMY_MT) {
                for (int i = tid; i < (HMY_MT - nt) * (RSH / 2); i += NTHR) {
                    const int rr = nt + i / (RSH / 2), cc = 2 * (i % (RSH / 2));
                    *reinterpret_cast<unsigned int*>(Rh + rr * RSH + cc) = 0u;
                    *reinterpret_cast<unsigned int*>(Rl + rr * RSH + cc) = 0u;
                }
            }
            __syncthreads();
            // corr[cell][j] = sum_k R[cell][k] Wc[k][j]: rows 16 mw .. 16 mw + 15 (WN = 2: the nh = 0 warps)
            const int row0 = 16 * mw;
            float acc[ND][4];
#pragma unroll
            for (int i = 0; i < ND; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            if (row0 < nt && nh == 0) {
                for (int ks = 0; ks < kst; ++ks) {
                    unsigned int ah[4], al[4];
                    const int arow = row0 + lr + 8 * (lj & 1), acol = 16 * ks + 8 * (lj >> 1);
                    ldsm_x4(ah, smem_u32(Rh + arow * RSH + acol));
                    ldsm_x4(al, smem_u32(Rl + arow * RSH + acol));
#pragma unroll
                    for (int ip = 0; ip < ND; ip += 2) {
                        if (8 * ip < dpa) {
                            unsigned int bh[4], bl[4];
                            const int brow = 8 * ip + lr + 8 * (lj >> 1), bcol = 16 * ks + 8 * (lj & 1);
                            ldsm_x4(bh, smem_u32(Wh + brow * RSH + bcol));
                            ldsm_x4(bl, smem_u32(Wl + brow * RSH + bcol));
                            mma_f16(acc[ip], al, bh[0], bh[1]);
                            mma_f16(acc[ip + 1], al, bh[2], bh[3]);
                            mma_f16(acc[ip], ah, bl[0], bl[1]);
                            mma_f16(acc[ip + 1], ah, bl[2], bl[3]);
                            mma_f16(acc[ip], ah, bh[0], bh[1]);
                            mma_f16(acc[ip + 1], ah, bh[2], bh[3]);
                        }
                    }
                }
            }
            // ---- Z_corr, Z_cos  (harmony.py:566, :569); Z_cos also as fp16 hi/lo for the centroid sums
            if (row0 < nt && nh == 0) {
                const bool v0 = (row0 + g) < nt, v1 = (row0 + g + 8) < nt;
                const size_t r0 = (size_t)(base + row0 + g) * dp, r1 = (size_t)(base + row0 + g + 8) * dp;
                float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
                for (int i = 0; i < ND; ++i) {
                    const int j = 8 * i + 2 * t;
                    float2 z0 = make_float2(0.f, 0.f), z1 = make_float2(0.f, 0.f);
                    if (j < dp) {
                        if (v0) z0 = __ldg(reinterpret_cast<const float2*>(st.Zorig + r0 + j));
                        if (v1) z1 = __ldg(reinterpret_cast<const float2*>(st.Zorig + r1 + j));
                    }
                    acc[i][0] = z0.x - acc[i][0] * cinv; acc[i][1] = z0.y - acc[i][1] * cinv;
                    acc[i][2] = z1.x - acc[i][2] * cinv; acc[i][3] = z1.y - acc[i][3] * cinv;
                    if (j >= st.d) { acc[i][0] = 0.f; acc[i][2] = 0.f; }
                    if (j + 1 >= st.d) { acc[i][1] = 0.f; acc[i][3] = 0.f; }
                    ss0 += acc[i][0] * acc[i][0] + acc[i][1] * acc[i][1];
                    ss1 += acc[i][2] * acc[i][2] + acc[i][3] * acc[i][3];
                }
                ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
                ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
                const float in0 = 1.0f / sqrtf(ss0), in1 = 1.0f / sqrtf(ss1);
#pragma unroll
                for (int i = 0; i < ND; ++i) {
                    const int j = 8 * i + 2 * t;
                    const float c00 = acc[i][0] * in0, c01 = acc[i][1] * in0, c10 = acc[i][2] * in1, c11 = acc[i][3] * in1;
                    if (j < dp) {
                        if (v0) {
                            *reinterpret_cast<float2*>(st.Zcorr + r0 + j) = make_float2(acc[i][0], acc[i][1]);
                            *reinterpret_cast<float2*>(st.Zcos + r0 + j) = make_float2(c00, c01);
                        }
                        if (v1) {
                            *reinterpret_cast<float2*>(st.Zcorr + r1 + j) = make_float2(acc[i][2], acc[i][3]);
                            *reinterpret_cast<float2*>(st.Zcos + r1 + j) = make_float2(c10, c11);
                        }
                    }
                    if (8 * i < dpa) {
                        unsigned int h0, l0, h1, l1;
                        split2(v0 ? c00 * HMY_OPSCALE : 0.f, v0 ? c01 * HMY_OPSCALE : 0.f, h0, l0);
                        split2(v1 ? c10 * HMY_OPSCALE : 0.f, v1 ? c11 * HMY_OPSCALE : 0.f, h1, l1);
                        *reinterpret_cast<unsigned int*>(Zh + (row0 + g) * ZSH + j) = h0;
                        *reinterpret_cast<unsigned int*>(Zl + (row0 + g) * ZSH + j) = l0;
                        *reinterpret_cast<unsigned int*>(Zh + (row0 + g + 8) * ZSH + j) = h1;
                        *reinterpret_cast<unsigned int*>(Zl + (row0 + g + 8) * ZSH + j) = l1;
                        // the same split values are the operand rows of the tensor-memory round kernel (hmy_common.cuh: Zs16)
                        if (st.Zs16 != nullptr && j < 2 * zs_pairs) {
                            if (v0) { zs_rows[(size_t)(base + row0 + g) * (2 * zs_pairs) + (j >> 1)] = h0; zs_rows[(size_t)(base + row0 + g) * (2 * zs_pairs) + zs_pairs + (j >> 1)] = l0; }
                            if (v1) { zs_rows[(size_t)(base + row0 + g + 8) * (2 * zs_pairs) + (j >> 1)] = h1; zs_rows[(size_t)(base + row0 + g + 8) * (2 * zs_pairs) + zs_pairs + (j >> 1)] = l1; }
                        }
                    }
                }
            }
            __syncthreads();
            if (mw < mtiles) ridge_ztr<NT>(yacc, Zh, Zl, ZSH, Rh, Rl, RSH, mw, n0, nt);
            __syncthreads();
        }
    }
    // centroid sums of the next cluster() (scaled 2^-20 like the round kernel's)
    const double sc = (double)HMY_ACCSCALE;
    if (mw < mtiles) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 16 * mw + g + 8 * (e >> 1), k = 8 * (n0 + i) + 2 * t + (e & 1);
                if (j < st.d && k < st.K && yacc[i][e] != 0.f) atomicAdd(&st.Yacc[(size_t)k * dp + j], (double)yacc[i][e] * sc);
            }
    }
}
