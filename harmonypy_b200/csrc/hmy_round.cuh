// k-means round of Harmony.cluster() as one streaming pass over the cells.
//
// Replaces harmony.py:443-453 (centroid update, cosine distances, update_R, objective) and
// the post-sklearn tail of init_cluster (harmony.py:377-392).  SURVEY.md Appendix A is the
// per-cell statement this implements.
//
// Work decomposition: CTA c owns the contiguous cell range [c*N/G, (c+1)*N/G) for the whole
// round.  Phase 0 streams the range once (old R rows) to (a) bucket its cells by update
// block and (b) sum the assignments each block is about to remove (Told).  Then the blocks
// are processed in order; after each block the per-block batch sums (Dnew) of all CTAs must
// be complete before the next penalty table exists, which is a grid barrier whose LAST
// arriving CTA does the K x B table math (O, E, penalty) and publishes it.
#pragma once
#include "hmy_common.cuh"

// ------------------------------------------------------------------------------------------
// shared-memory plan (identical on host and device)
__host__ __device__ inline int phase0_tables(int KS) { return (KS <= 128) ? 4 : 2; }

struct RoundSmem {
    int ZS, RS;
    int off_YsT, off_sigma, off_Ps, off_union, off_misc;
    int off_Zs, off_Rs, off_cell, off_ccombo, off_clev;      // step view of the union
    int off_T, off_cnt, off_btot;                            // phase-0 view of the union
    int total;
};

__host__ __device__ inline RoundSmem round_smem_plan(int dp, int KS, int B, int V, int nblk, int JPW) {
    RoundSmem s;
    s.ZS = 8 * JPW + 4;
    s.RS = KS + 4;
    int o = 0;
    s.off_YsT = o;    o += dp * KS * 4;
    s.off_sigma = o;  o += KS * 4;
    s.off_Ps = o;     o += B * KS * 4;
    s.off_union = o;
    int a = o;
    s.off_Zs = a;     a += HMY_TILE * s.ZS * 4;
    s.off_Rs = a;     a += HMY_TILE * s.RS * 4;
    s.off_cell = a;   a += HMY_TILE * 4;
    s.off_ccombo = a; a += HMY_TILE * 4;
    s.off_clev = a;   a += HMY_TILE * V * 4;
    int b = o;
    s.off_T = b;      b += phase0_tables(KS) * nblk * KS * 4;
    s.off_cnt = b;    s.off_btot = b;
    o = (a > b ? a : b);
    o = (o + 15) & ~15;
    s.off_misc = o;   o += 8 * 256 + 128;      // K doubles of scratch for the serial sections
    s.total = o;
    return s;
}

// ------------------------------------------------------------------------------------------
// K x B table math, executed by all threads of ONE CTA (the last to arrive at a barrier, or a
// single-CTA kernel in staged mode).  Tables are [b][k].  .cg accesses everywhere: successive
// serial sections run on different SMs, so nothing may be served from a stale L1 line.

// Orun <- O: the maintained O of harmony.py:389 / :507 as of the end of the previous stage.
__device__ inline void serial_copy_O(const HmyDev& st) {
    for (int i = threadIdx.x; i < st.B * st.K; i += blockDim.x) __stcg(&st.Orun[i], __ldcg(&st.O[i]));
    __syncthreads();
}

// sum_n R[n][k]: every cell has exactly one level of covariate 0, so the row sum of R is the
// sum of O over covariate 0's one-hot rows (this is what E is an outer product of, :388).
__device__ inline void serial_rowsum(const HmyDev& st, const double* Otab, double* s_row) {
    for (int k = threadIdx.x; k < st.K; k += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < st.lev0; ++b) s += __ldcg(&Otab[b * st.K + k]);
        s_row[k] = s;
    }
    __syncthreads();
}

// Put block blk-1 back (harmony.py:506-507), take block blk out (:491-492), then the penalty
// (E / (O + E))^theta with the reference's clamps (:495-499, :579-584).
// Thread k owns cluster k and issues ALL loads of its column before using any of them: this
// runs on one CTA between two blocks, so its latency sits on every block's critical path.
#define HMY_SER_B 24
__device__ inline void serial_prepare_block(const HmyDev& st, int blk, double* /*s_row*/) {
    const int BK = st.B * st.K, K = st.K, B = st.B;
    const float* dn = st.Dnew + (size_t)(blk > 0 ? blk - 1 : 0) * BK;
    const float* to = st.Told + (size_t)blk * BK;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double rs = 0.0;
        if (B <= HMY_SER_B) {
            double o[HMY_SER_B]; float a[HMY_SER_B], r[HMY_SER_B], pr[HMY_SER_B], th[HMY_SER_B];
#pragma unroll
            for (int b = 0; b < HMY_SER_B; ++b) {
                pr[b] = (b < B) ? st.Pr_b[b] : 0.f;
                th[b] = (b < B) ? st.theta[b] : 2.f;
                o[b] = (b < B) ? __ldcg(&st.Orun[b * K + k]) : 0.0;
                a[b] = (b < B && blk > 0) ? __ldcg(&dn[b * K + k]) : 0.f;
                r[b] = (b < B) ? __ldcg(&to[b * K + k]) : 0.f;
            }
#pragma unroll
            for (int b = 0; b < HMY_SER_B; ++b) {
                o[b] += (double)a[b] - (double)r[b];
                if (b < st.lev0) rs += o[b];
            }
#pragma unroll
            for (int b = 0; b < HMY_SER_B; ++b) {
                if (b < B) {
                    __stcg(&st.Orun[b * K + k], o[b]);
                    const float of = (float)o[b];
                    const float e = (float)(rs * (double)pr[b]);
                    const float ratio = fminf(fmaxf(e / fmaxf(of + e, 1e-8f), 1e-8f), 1.0f);
                    __stcg(&st.P[b * K + k], (th[b] == 2.0f) ? ratio * ratio : powf(ratio, th[b]));
                }
            }
        } else {
            for (int b0 = 0; b0 < B; b0 += 8) {
                double o[8]; float a[8], r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int b = b0 + u;
                    o[u] = (b < B) ? __ldcg(&st.Orun[b * K + k]) : 0.0;
                    a[u] = (b < B && blk > 0) ? __ldcg(&dn[b * K + k]) : 0.f;
                    r[u] = (b < B) ? __ldcg(&to[b * K + k]) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int b = b0 + u;
                    if (b < B) {
                        const double v = o[u] + (double)a[u] - (double)r[u];
                        __stcg(&st.Orun[b * K + k], v);
                        if (b < st.lev0) rs += v;
                    }
                }
            }
            for (int b0 = 0; b0 < B; b0 += 8) {
                double o[8]; float pr[8], th[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    o[u] = (b0 + u < B) ? __ldcg(&st.Orun[(b0 + u) * K + k]) : 0.0;
                    pr[u] = (b0 + u < B) ? st.Pr_b[b0 + u] : 0.f;
                    th[u] = (b0 + u < B) ? st.theta[b0 + u] : 2.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int b = b0 + u;
                    if (b < B) {
                        const float of = (float)o[u];
                        const float e = (float)(rs * (double)pr[u]);
                        const float ratio = fminf(fmaxf(e / fmaxf(of + e, 1e-8f), 1e-8f), 1.0f);
                        __stcg(&st.P[b * K + k], (th[u] == 2.0f) ? ratio * ratio : powf(ratio, th[u]));
                    }
                }
            }
        }
    }
    __syncthreads();
}

// End of a stage: fold the last block back in, publish O, the cross-entropy term of the
// objective (harmony.py:404-411, collapsed to a K x B sum because sum_n R[n][k] Phi[b][n] is
// O[k][b]), and the unit centroids for the next round (harmony.py:443-444).
// mode 0 = k-means round, 1 = init (Orun was zero and Dnew[0] holds all of O), 2 = centroids only.
__device__ inline void serial_finalize(const HmyDev& st, int mode, double* s_row, double* s_red) {
    const int BK = st.B * st.K;
    if (mode != 2) {
        // the maintained O after a full pass over the blocks is the sum of what the blocks put
        // back (harmony.py:506-507 over all blocks; for init: harmony.py:389)
        for (int i = threadIdx.x; i < BK; i += blockDim.x) {
            const double o = __ldcg(&st.Ofresh[i]);
            __stcg(&st.Orun[i], o);
            __stcg(&st.O[i], o);
        }
        __syncthreads();
        serial_rowsum(st, st.Orun, s_row);
        double part = 0.0;
        for (int i0 = threadIdx.x; i0 < BK; i0 += blockDim.x * 8) {
            double o8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; o8[u] = (i < BK) ? __ldcg(&st.Orun[i]) : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < BK) {
                    const int b = i / st.K, k = i - b * st.K;
                    const double o = o8[u];
                    const float oc = fmaxf((float)o, 1e-8f);
                    const float ec = fmaxf((float)(s_row[k] * (double)st.Pr_b[b]), 1e-8f);
                    part += (double)st.sigma[k] * (double)st.theta[b] * (double)logf((oc + ec) / ec) * o;
                }
            }
        }
        part = warp_sum_d(part);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += s_red[w];
            st.obj_out[0] = __ldcg(&st.obj[0]);
            st.obj_out[1] = __ldcg(&st.obj[1]);
            st.obj_out[2] = c;
        }
        __syncthreads();
    }
    // unit centroids from the accumulated sums (each warp: 4 clusters at a time, loads up front)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k0 = 4 * warp; k0 < st.K; k0 += 4 * nw) {
        double y[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int k = k0 + q, j = lane + 32 * m;
                y[q][m] = (k < st.K && j < st.d) ? __ldcg(&st.Yacc[(size_t)k * st.dp + j]) : 0.0;
            }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + q;
            double ss = y[q][0] * y[q][0] + y[q][1] * y[q][1] + y[q][2] * y[q][2] + y[q][3] * y[q][3];
            ss = warp_sum_d(ss);
            const double inv = 1.0 / sqrt(ss);
            if (k < st.K) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int j = lane + 32 * m;
                    if (j < st.dp) st.Ynext[(size_t)k * st.dp + j] = (j < st.d) ? (float)(y[q][m] * inv) : 0.f;
                }
            }
        }
    }
    __syncthreads();
}

// CTA `cta` of `G` takes an equal share of block blk's cell list (global, position-sorted)
__device__ __forceinline__ void block_share(const HmyDev& st, int blk, unsigned int cta, unsigned int G,
                                            long long& lb, long long& le) {
    const long long b0 = st.blk_start[blk], n = st.blk_start[blk + 1] - b0;
    lb = b0 + n * cta / G;
    le = b0 + n * (cta + 1) / G;
}

// ------------------------------------------------------------------------------------------
// per-CTA context
template <int KPT, int JPW>
struct RoundCtx {
    float* YsT; float* sSigma; float* Ps; float* Zs; float* Rs;
    int* sCell; int* sCombo; int* sLev;
    float* T; unsigned int* cnt; int* btot;
    double* sRow; double* sRed; int* sFlag;
    int ZS, RS, KS;
    // running batch-sum (thread k < K): combo of the current run and its sum
    int run_combo; float run_sum;
    float acc2[KPT][JPW];     // centroid partial sums, persistent over the whole round
    double objd, obje;
};

template <int KPT, int JPW>
__device__ __forceinline__ void round_ctx_init(RoundCtx<KPT, JPW>& c, const HmyDev& st, unsigned char* smem) {
    const RoundSmem p = round_smem_plan(st.dp, st.KS, st.B, st.V, st.nblk, JPW);
    c.ZS = p.ZS; c.RS = p.RS; c.KS = st.KS;
    c.YsT = (float*)(smem + p.off_YsT); c.sSigma = (float*)(smem + p.off_sigma);
    c.Ps = (float*)(smem + p.off_Ps);
    c.Zs = (float*)(smem + p.off_Zs); c.Rs = (float*)(smem + p.off_Rs);
    c.sCell = (int*)(smem + p.off_cell); c.sCombo = (int*)(smem + p.off_ccombo); c.sLev = (int*)(smem + p.off_clev);
    c.T = (float*)(smem + p.off_T); c.cnt = (unsigned int*)(smem + p.off_cnt); c.btot = (int*)(smem + p.off_btot);
    c.sRow = (double*)(smem + p.off_misc); c.sRed = c.sRow + 256; c.sFlag = (int*)(c.sRed + 8);
    c.run_combo = -1; c.run_sum = 0.f;
#pragma unroll
    for (int q = 0; q < KPT; ++q)
#pragma unroll
        for (int j = 0; j < JPW; ++j) c.acc2[q][j] = 0.f;
    c.objd = 0.0; c.obje = 0.0;
}

// centroids (transposed, zero padded to KS) and sigma into shared memory
template <int KPT, int JPW>
__device__ __forceinline__ void load_centroids(RoundCtx<KPT, JPW>& c, const HmyDev& st) {
    for (int i = threadIdx.x; i < st.dp * c.KS; i += blockDim.x) {
        const int j = i / c.KS, k = i - j * c.KS;
        c.YsT[i] = (k < st.K) ? st.Yhat[(size_t)k * st.dp + j] : 0.f;
    }
    for (int k = threadIdx.x; k < c.KS; k += blockDim.x) c.sSigma[k] = (k < st.K) ? st.sigma[k] : 1.f;
}

template <int KPT, int JPW>
__device__ __forceinline__ void zero_step_buffers(RoundCtx<KPT, JPW>& c) {
    for (int i = threadIdx.x; i < HMY_TILE * c.ZS; i += blockDim.x) c.Zs[i] = 0.f;
    for (int i = threadIdx.x; i < HMY_TILE * c.RS; i += blockDim.x) c.Rs[i] = 0.f;
}

// ------------------------------------------------------------------------------------------
// Phase 0: sum what each block will remove: Told[blk][b][k] = sum_{n in blk, level b} R_old[n][k]
// (the R_block @ Phi_block.T and R_block.sum of harmony.py:491-492, for all blocks at once).
struct Phase0Mem { float* T; unsigned int* cnt; int* btot; int KS; };

template <int NTHR>
__device__ void phase0(const Phase0Mem c, const HmyDev& st, long long c0, long long c1) {
    constexpr int HMY_THREADS_L = NTHR;
    const int tid = threadIdx.x;
    const int n = (int)(c1 - c0);
    const int nblk = st.nblk;
    // ---- (b) Told.  GS threads cover one R row; NG groups x U tables keep the shared-memory
    // read-modify-write chains of consecutive cells independent.
    const int GS = (c.KS <= 128) ? 128 : 256;
    const int NG = HMY_THREADS_L / GS;
    const int NTAB = phase0_tables(c.KS);
    const int U = NTAB / NG;
    const int g = tid / GS, kk = tid - g * GS;
    const bool active = kk < st.K;
    const int TS = nblk * c.KS;                 // one table
    long long s0 = c0;
    while (s0 < c1) {
        const int combo = st.combo[s0];
        const long long s1 = min(c1, st.combo_start[combo + 1]);
        for (int i = tid; i < NTAB * TS; i += HMY_THREADS_L) c.T[i] = 0.f;
        __syncthreads();
        if (kk < c.KS) {
            float* Tg = c.T + (size_t)g * U * TS + kk;
            constexpr int BATCH = 16;
            for (long long base = s0 + (long long)g * BATCH; base < s1; base += (long long)NG * BATCH) {
                float r[BATCH]; int b[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const long long cell = base + u;
                    const bool ok = cell < s1;
                    b[u] = ok ? (int)st.blk[cell] : 0;
                    r[u] = (ok && active) ? __ldg(&st.R[(size_t)cell * st.Kp + kk]) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < BATCH; u += 4) {
                    // up to 4 independent tables per group (U = 2 or 4): table index u % U
                    float* p0 = Tg + (size_t)((u + 0) % U) * TS + b[u + 0] * c.KS;
                    float* p1 = Tg + (size_t)((u + 1) % U) * TS + b[u + 1] * c.KS;
                    float* p2 = Tg + (size_t)((u + 2) % U) * TS + b[u + 2] * c.KS;
                    float* p3 = Tg + (size_t)((u + 3) % U) * TS + b[u + 3] * c.KS;
                    if (U == 4) {
                        const float t0v = *p0, t1v = *p1, t2v = *p2, t3v = *p3;
                        *p0 = t0v + r[u + 0]; *p1 = t1v + r[u + 1]; *p2 = t2v + r[u + 2]; *p3 = t3v + r[u + 3];
                    } else {
                        const float t0v = *p0, t1v = *p1;
                        *p0 = t0v + r[u + 0]; *p1 = t1v + r[u + 1];
                        const float t2v = *p2, t3v = *p3;
                        *p2 = t2v + r[u + 2]; *p3 = t3v + r[u + 3];
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < nblk * st.K; i += HMY_THREADS_L) {
            const int b = i / st.K, k = i - b * st.K;
            float s = 0.f;
            for (int u = 0; u < NTAB; ++u) s += c.T[u * TS + b * c.KS + k];
            if (s != 0.f) {
                for (int v = 0; v < st.V; ++v) {
                    const int lev = st.combo_lev[combo * st.V + v];
                    atomicAdd(&st.Told[((size_t)b * st.B + lev) * st.K + k], s);
                }
            }
        }
        __syncthreads();
        s0 = s1;
    }
}

// ------------------------------------------------------------------------------------------
// batch-sum run handling (thread k < K owns column k of the CTA's contribution to Dnew[blk])
template <int KPT, int JPW>
__device__ __forceinline__ void flush_run(RoundCtx<KPT, JPW>& c, const HmyDev& st, int blk) {
    if (c.run_combo >= 0 && c.run_sum != 0.f) {
        for (int v = 0; v < st.V; ++v) {
            const int lev = st.combo_lev[c.run_combo * st.V + v];
            atomicAdd(&st.Dnew[((size_t)blk * st.B + lev) * st.K + threadIdx.x], c.run_sum);
            atomicAdd(&st.Ofresh[(size_t)lev * st.K + threadIdx.x], (double)c.run_sum);
        }
    }
    c.run_combo = -1; c.run_sum = 0.f;
}

// One block of update_R for this CTA's cells (harmony.py:495-509) -- or, with init = true, the
// un-penalised assignment of init_cluster (harmony.py:380-389) over all of its cells.
//   cells: list[lbeg .. lend) (global positions) or the identity range when list == nullptr.
template <int KPT, int JPW>
__device__ void process_block(RoundCtx<KPT, JPW>& c, const HmyDev& st, int blk, const int* list,
                              long long lbeg, long long lend, bool init) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dp = st.dp, dp4 = dp >> 2, Kp = st.Kp, Kp4 = Kp >> 2, K = st.K, V = st.V;
    const int KS = c.KS, ZS = c.ZS, RS = c.RS;
    for (long long tb = lbeg; tb < lend; tb += HMY_TILE) {
        const int nt = (int)min((long long)HMY_TILE, lend - tb);
        // ---- stage the tile: cell ids, levels, Z_cos rows
        if (tid < nt) {
            const int cell = list ? list[tb + tid] : (int)(tb + tid);
            const int combo = st.combo[cell];
            c.sCell[tid] = cell; c.sCombo[tid] = combo;
            for (int v = 0; v < V; ++v) c.sLev[tid * V + v] = st.combo_lev[combo * V + v];
        }
        __syncthreads();
        for (int i = tid; i < nt * dp4; i += HMY_THREADS) {
            const int row = i / dp4, c4 = i - row * dp4;
            const float4 z = __ldg(reinterpret_cast<const float4*>(st.Zcos + (size_t)c.sCell[row] * dp) + c4);
            *reinterpret_cast<float4*>(c.Zs + row * ZS + 4 * c4) = z;
        }
        __syncthreads();
        // ---- scores: warp handles HMY_CPW cells, lane handles clusters lane + 32 q
        float acc[HMY_CPW][KPT];
#pragma unroll
        for (int i = 0; i < HMY_CPW; ++i)
#pragma unroll
            for (int q = 0; q < KPT; ++q) acc[i][q] = 0.f;
        const float* zbase = c.Zs + warp * HMY_CPW * ZS;
        for (int j = 0; j < dp; j += 4) {
            float4 zv[HMY_CPW];
#pragma unroll
            for (int i = 0; i < HMY_CPW; ++i) zv[i] = *reinterpret_cast<const float4*>(zbase + i * ZS + j);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float y[KPT];
#pragma unroll
                for (int q = 0; q < KPT; ++q) y[q] = c.YsT[(j + jj) * KS + lane + 32 * q];
#pragma unroll
                for (int i = 0; i < HMY_CPW; ++i) {
                    const float zz = (jj == 0) ? zv[i].x : (jj == 1) ? zv[i].y : (jj == 2) ? zv[i].z : zv[i].w;
#pragma unroll
                    for (int q = 0; q < KPT; ++q) acc[i][q] = fmaf(zz, y[q], acc[i][q]);
                }
            }
        }
        // ---- soft assignment per cell (harmony.py:466-468, :500-503)
        float od = 0.f, oe = 0.f;
#pragma unroll
        for (int i = 0; i < HMY_CPW; ++i) {
            const int cl = warp * HMY_CPW + i;
            const bool valid = cl < nt;
            float s[KPT], dist[KPT];
            float ssum = 0.f;
#pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const int k = lane + 32 * q;
                dist[q] = 2.f * (1.f - acc[i][q]);
                s[q] = (k < K) ? expf(-dist[q] / c.sSigma[k]) : 0.f;
                ssum += s[q];
            }
            ssum = warp_sum(ssum);
            float rsum = 0.f;
#pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const int k = lane + 32 * q;
                float pen = 1.f;
                if (!init) {
                    pen = 0.f;
                    if (valid)
                        for (int v = 0; v < V; ++v) pen += c.Ps[c.sLev[cl * V + v] * KS + k];
                }
                s[q] = (s[q] / ssum) * pen;
                rsum += s[q];
            }
            rsum = warp_sum(rsum);
            const float den = fmaxf(rsum, 1e-8f);
#pragma unroll
            for (int q = 0; q < KPT; ++q) {
                const int k = lane + 32 * q;
                const float r = s[q] / den;
                if (valid) {
                    c.Rs[cl * RS + k] = r;
                    if (k < K) {
                        od += r * dist[q];
                        if (r > 0.f) oe += c.sSigma[k] * r * logf(r);
                    }
                }
            }
        }
        c.objd += (double)od; c.obje += (double)oe;
        __syncthreads();
        // ---- new R rows to HBM (harmony.py:509)
        for (int i = tid; i < nt * Kp4; i += HMY_THREADS) {
            const int row = i / Kp4, c4 = i - row * Kp4;
            const float4 r = *reinterpret_cast<const float4*>(c.Rs + row * RS + 4 * c4);
            *(reinterpret_cast<float4*>(st.R + (size_t)c.sCell[row] * Kp) + c4) = r;
        }
        // ---- batch sums of the new assignments (harmony.py:506-507), run-length over combos
        if (tid < K) {
            for (int n = 0; n < nt; ++n) {
                const int cb = c.sCombo[n];
                if (cb != c.run_combo) { flush_run(c, st, blk); c.run_combo = cb; }
                c.run_sum += c.Rs[n * RS + tid];
            }
        }
        // ---- centroid sums for the next round: Yacc[k][j] += R[n][k] * z[n][j] (harmony.py:443)
        for (int n = 0; n < nt; ++n) {
            float r[KPT];
#pragma unroll
            for (int q = 0; q < KPT; ++q) r[q] = c.Rs[n * RS + lane + 32 * q];
            float z[JPW];
#pragma unroll
            for (int m = 0; m < JPW; m += 4) {
                const float4 t = *reinterpret_cast<const float4*>(c.Zs + n * ZS + warp * JPW + m);
                z[m] = t.x; z[m + 1] = t.y; z[m + 2] = t.z; z[m + 3] = t.w;
            }
#pragma unroll
            for (int q = 0; q < KPT; ++q)
#pragma unroll
                for (int m = 0; m < JPW; ++m) c.acc2[q][m] = fmaf(r[q], z[m], c.acc2[q][m]);
        }
        __syncthreads();
    }
    if (tid < K) flush_run(c, st, blk);
}

// centroid partial sums and objective partial sums of this CTA -> global accumulators
template <int KPT, int JPW>
__device__ void flush_round_sums(RoundCtx<KPT, JPW>& c, const HmyDev& st) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
        const int k = lane + 32 * q;
#pragma unroll
        for (int m = 0; m < JPW; ++m) {
            const int j = warp * JPW + m;
            if (k < st.K && j < st.d && c.acc2[q][m] != 0.f)
                atomicAdd(&st.Yacc[(size_t)k * st.dp + j], (double)c.acc2[q][m]);
            c.acc2[q][m] = 0.f;
        }
    }
    double a = warp_sum_d(c.objd), b = warp_sum_d(c.obje);
    if (lane == 0) { atomicAdd(&st.obj[0], a); atomicAdd(&st.obj[1], b); }
    c.objd = 0.0; c.obje = 0.0;
}

// ------------------------------------------------------------------------------------------
// Persistent kernel: one launch per round (mode 0) or for the init assignment (mode 1).
template <int KPT, int JPW>
__global__ void __launch_bounds__(HMY_THREADS) k_round(HmyDev st, int mode, unsigned int gen_base) {
    extern __shared__ __align__(16) unsigned char smem[];
    RoundCtx<KPT, JPW> c;
    round_ctx_init(c, st, smem);
    const unsigned int G = gridDim.x;
    const long long c0 = (long long)blockIdx.x * st.N / G, c1 = (long long)(blockIdx.x + 1) * st.N / G;
    load_centroids(c, st);
    __syncthreads();
    if (mode == 1) {
        zero_step_buffers(c);
        __syncthreads();
        process_block(c, st, 0, nullptr, c0, c1, true);
        flush_round_sums(c, st);
        grid_barrier_serial(st, G, gen_base + 1u, c.sFlag, [&]() { serial_finalize(st, 1, c.sRow, c.sRed); });
        return;
    }
    phase0<HMY_THREADS>(Phase0Mem{c.T, c.cnt, c.btot, c.KS}, st, c0, c1);
    unsigned int gen = gen_base + 1u;
    grid_barrier_serial(st, G, gen++, c.sFlag, [&]() { serial_copy_O(st); serial_prepare_block(st, 0, c.sRow); });
    zero_step_buffers(c);
    for (int blk = 0; blk < st.nblk; ++blk) {
        for (int i = threadIdx.x; i < st.B * c.KS; i += HMY_THREADS) {
            const int b = i / c.KS, k = i - b * c.KS;
            c.Ps[i] = (k < st.K) ? __ldcg(&st.P[b * st.K + k]) : 0.f;
        }
        __syncthreads();
        long long lb, le;
        block_share(st, blk, blockIdx.x, G, lb, le);
        process_block(c, st, blk, st.list, lb, le, false);
        if (blk + 1 < st.nblk) {
            grid_barrier_serial(st, G, gen++, c.sFlag, [&]() { serial_prepare_block(st, blk + 1, c.sRow); });
        } else {
            flush_round_sums(c, st);
            grid_barrier_serial(st, G, gen++, c.sFlag, [&]() { serial_finalize(st, 0, c.sRow, c.sRed); });
        }
    }
}

// Staged variants (one launch per phase; used when a host-side all-reduce sits between the
// phases in multi-GPU staged mode, and as a debugging cross-check of the persistent kernel).
//   what 0 = phase 0, 1 = block step `blk`, 2 = init assignment
template <int KPT, int JPW>
__global__ void __launch_bounds__(HMY_THREADS) k_round_stage(HmyDev st, int what, int blk) {
    extern __shared__ __align__(16) unsigned char smem[];
    RoundCtx<KPT, JPW> c;
    round_ctx_init(c, st, smem);
    const unsigned int G = gridDim.x;
    const long long c0 = (long long)blockIdx.x * st.N / G, c1 = (long long)(blockIdx.x + 1) * st.N / G;
    if (what == 0) { phase0<HMY_THREADS>(Phase0Mem{c.T, c.cnt, c.btot, c.KS}, st, c0, c1); return; }
    load_centroids(c, st);
    zero_step_buffers(c);
    if (what == 1) {
        for (int i = threadIdx.x; i < st.B * c.KS; i += HMY_THREADS) {
            const int b = i / c.KS, k = i - b * c.KS;
            c.Ps[i] = (k < st.K) ? __ldcg(&st.P[b * st.K + k]) : 0.f;
        }
    }
    __syncthreads();
    if (what == 1) {
        long long lb, le;
        block_share(st, blk, blockIdx.x, G, lb, le);
        process_block(c, st, blk, st.list, lb, le, false);
    } else {
        process_block(c, st, 0, nullptr, c0, c1, true);
    }
    flush_round_sums(c, st);
}

#ifdef HMY_NONTEMPLATE_KERNELS
// ------------------------------------------------------------------------------------------
// Per-block cell lists of the round: list[blk_start[b] .. blk_start[b+1]) = positions of the
// cells of block b, ascending (stable counting sort over chunks of the position range).
#define HMY_LIST_THREADS 256
// pass 1 (count = 1): cnt[chunk][b]; pass 2 (count = 0): scatter using base[chunk][b]
__global__ void __launch_bounds__(HMY_LIST_THREADS) k_block_lists(HmyDev st, int* cnt, int count) {
    extern __shared__ unsigned int s_cnt[];            // [nblk][256] + [nblk]
    const int tid = threadIdx.x, nblk = st.nblk;
    unsigned int* s_tot = s_cnt + nblk * HMY_LIST_THREADS;
    const long long c0 = (long long)blockIdx.x * st.N / gridDim.x, c1 = (long long)(blockIdx.x + 1) * st.N / gridDim.x;
    const int n = (int)(c1 - c0);
    const int per = (n + HMY_LIST_THREADS - 1) / HMY_LIST_THREADS;
    const int t0 = min(n, tid * per), t1 = min(n, t0 + per);
    for (int i = tid; i < nblk * HMY_LIST_THREADS; i += HMY_LIST_THREADS) s_cnt[i] = 0u;
    __syncthreads();
    for (int i = t0; i < t1; ++i) s_cnt[(int)st.blk[c0 + i] * HMY_LIST_THREADS + tid]++;
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int b = warp; b < nblk; b += HMY_LIST_THREADS / 32) {
        unsigned int loc[8]; unsigned int s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { loc[i] = s_cnt[b * HMY_LIST_THREADS + 8 * lane + i]; s += loc[i]; }
        unsigned int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        unsigned int run = inc - s;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s_cnt[b * HMY_LIST_THREADS + 8 * lane + i] = run; run += loc[i]; }
        if (lane == 31) s_tot[b] = inc;
    }
    __syncthreads();
    if (count) {
        for (int b = tid; b < nblk; b += HMY_LIST_THREADS) cnt[(size_t)blockIdx.x * nblk + b] = (int)s_tot[b];
        return;
    }
    for (int i = t0; i < t1; ++i) {
        const int b = st.blk[c0 + i];
        const long long pos = st.blk_start[b] + cnt[(size_t)blockIdx.x * nblk + b] + s_cnt[b * HMY_LIST_THREADS + tid]++;
        if (st.list2 != nullptr) st.list2[pos] = make_int2((int)(c0 + i), (st.combo[c0 + i] << 8) | (int)st.blk_next[c0 + i]);
        else st.list[pos] = (int)(c0 + i);
    }
}

// cnt[chunk][b] -> exclusive prefix over chunks (in place) and blk_start[b] (exclusive over blocks).
// One warp per block id; a lane owns a contiguous run of <= 32 chunks whose counts it loads in one
// batch (a load-add-store chain over all chunks was 130 us per round on the critical path).
#define HMY_SCAN_PER_LANE 32
__global__ void __launch_bounds__(1024) k_block_scan(HmyDev st, int* cnt, int nchunk) {
    __shared__ long long tot[HMY_MAX_NBLK + 1];
    const int nblk = st.nblk, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int per = (nchunk + 31) / 32;                  // <= HMY_SCAN_PER_LANE (host checks)
    for (int b = warp; b < nblk; b += nw) {
        int v[HMY_SCAN_PER_LANE];
        int s = 0;
#pragma unroll
        for (int u = 0; u < HMY_SCAN_PER_LANE; ++u) {
            const int c = lane * per + u;
            v[u] = (u < per && c < nchunk) ? cnt[(size_t)c * nblk + b] : 0;
        }
#pragma unroll
        for (int u = 0; u < HMY_SCAN_PER_LANE; ++u) s += v[u];
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int run = inc - s;
#pragma unroll
        for (int u = 0; u < HMY_SCAN_PER_LANE; ++u) {
            const int c = lane * per + u;
            if (u < per && c < nchunk) { cnt[(size_t)c * nblk + b] = run; run += v[u]; }
        }
        if (lane == 31) tot[b] = inc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int b = 0; b < nblk; ++b) { st.blk_start[b] = run; run += tot[b]; }
        st.blk_start[nblk] = run;
    }
}

// single-CTA table kernel for staged mode: what 0 = copy O + prepare block 0,
// 1 = prepare block `blk`, 2 = finalize (mode in `blk`)
__global__ void __launch_bounds__(HMY_THREADS) k_tables(HmyDev st, int what, int blk) {
    __shared__ double sRow[256];
    __shared__ double sRed[8];
    if (what == 0) { serial_copy_O(st); serial_prepare_block(st, 0, sRow); }
    else if (what == 1) serial_prepare_block(st, blk, sRow);
    else serial_finalize(st, blk, sRow, sRed);
}

// ------------------------------------------------------------------------------------------
// block assignment of the round (harmony.py:471-475, :483-484)

// from the reference's host permutation: position i of the permutation -> block i / cpb
__global__ void k_assign_from_perm(HmyDev st, const long long* perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.Nglobal) return;
    const long long cell = perm[i] - st.cell_offset;
    if (cell < 0 || cell >= st.N) return;
    long long b = (st.cpb > 0) ? i / st.cpb : (long long)(st.nblk - 1);
    if (b > st.nblk - 1) b = st.nblk - 1;
    st.blk[st.pos_of[cell]] = (unsigned char)b;
}

__device__ __forceinline__ unsigned int hmy_mix(unsigned int x, unsigned int key) {
    x ^= key; x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}

// device-side pseudo-random permutation: 4-round Feistel network on 2*hb bits with cycle
// walking, i.e. a bijection of [0, Nglobal) keyed by (seed, round) -- every rank computes the
// same position for a cell without any memory traffic.
__global__ void k_assign_feistel(HmyDev st, unsigned long long seed, unsigned int round, int hb) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= st.N) return;
    const unsigned long long mask = (1ull << hb) - 1ull;
    unsigned long long x = (unsigned long long)(st.order[p] + st.cell_offset);
    const unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32) ^ (round * 0x632BE5ABu);
    do {
        unsigned long long l = x >> hb, r = x & mask;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned long long f = hmy_mix((unsigned int)r, k0 + 0x9E3779B9u * (unsigned int)(i + 1) + k1) & mask;
            const unsigned long long nl = r; r = l ^ f; l = nl;
        }
        x = (l << hb) | r;
    } while (x >= (unsigned long long)st.Nglobal);
    long long b = (st.cpb > 0) ? (long long)x / st.cpb : (long long)(st.nblk - 1);
    if (b > st.nblk - 1) b = st.nblk - 1;
    st.blk[p] = (unsigned char)b;
}
#endif  // HMY_NONTEMPLATE_KERNELS
