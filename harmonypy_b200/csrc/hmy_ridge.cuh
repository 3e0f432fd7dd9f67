// moe_correct_ridge (harmony.py:535-569) as two streaming passes and K small dense solves.
//
//   pass 1  k_ridge_moments : Gram[k] = Phi_moe diag(R_k) Phi_moe^T  (harmony.py:547-550) and
//                             Mom[k]  = Phi_moe diag(R_k) Z_orig^T   (the sums :556-563 gather)
//   solve   k_ridge_solve   : W_k = (Gram[k] + diag(lambda))^-1 Mom[k], row 0 dropped (:553-565)
//   pass 2  k_ridge_apply   : Z_corr = Z_orig - sum_k R_k (sum_v W_k[level_v]) (:566),
//                             Z_cos = unit(Z_corr) (:569), and the centroid sums Z_cos R^T that
//                             the next cluster() starts from (harmony.py:443).
//
// Cells are stored sorted by covariate combination, so each work item (<= HMY_SEG_MAX
// consecutive cells of ONE combination) turns the per-batch gathers of the reference into a
// plain (K x n)(n x d) contraction over a contiguous range of HBM.
#pragma once
#include <cuda_fp16.h>
#include "hmy_common.cuh"

#define HMY_SEG_MAX 256

struct RidgeSmem { int ZS, RS, WS; int off_Zs, off_Rs, off_Wc, total; };

__host__ __device__ inline RidgeSmem ridge_smem_plan(int K, int KS, int JPW, bool apply) {
    RidgeSmem s;
    s.ZS = 8 * JPW + 4; s.RS = KS + 4; s.WS = 8 * JPW;     // WS = 32 * (JPW / 4)
    int o = 0;
    s.off_Zs = o; o += HMY_TILE * s.ZS * 4;
    s.off_Rs = o; o += HMY_TILE * s.RS * 4;
    s.off_Wc = o; if (apply) o += K * s.WS * 4;
    s.total = o;
    return s;
}

// contiguous tile of cells [base, base+nt): Z rows (from `Zsrc`) and R rows into shared memory
__device__ __forceinline__ void ridge_load_tile(const HmyDev& st, const float* Zsrc, long long base, int nt,
                                                float* Zs, int ZS, float* Rs, int RS) {
    const int dp4 = st.dp >> 2, Kp4 = st.Kp >> 2;
    const float4* zsrc = reinterpret_cast<const float4*>(Zsrc + (size_t)base * st.dp);
    for (int i = threadIdx.x; i < nt * dp4; i += HMY_THREADS) {
        const int row = i / dp4, c4 = i - row * dp4;
        *reinterpret_cast<float4*>(Zs + row * ZS + 4 * c4) = __ldg(zsrc + i);
    }
    const float4* rsrc = reinterpret_cast<const float4*>(st.R + (size_t)base * st.Kp);
    for (int i = threadIdx.x; i < nt * Kp4; i += HMY_THREADS) {
        const int row = i / Kp4, c4 = i - row * Kp4;
        *reinterpret_cast<float4*>(Rs + row * RS + 4 * c4) = __ldg(rsrc + i);
    }
}

// acc2[q][m] += sum_n Rs[n][lane + 32 q] * Zs[n][warp * JPW + m]
template <int KPT, int JPW>
__device__ __forceinline__ void rtz_accumulate(float (&acc2)[KPT][JPW], const float* Rs, int RS,
                                               const float* Zs, int ZS, int nt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = 0; n < nt; ++n) {
        float r[KPT];
#pragma unroll
        for (int q = 0; q < KPT; ++q) r[q] = Rs[n * RS + lane + 32 * q];
        float z[JPW];
#pragma unroll
        for (int m = 0; m < JPW; m += 4) {
            const float4 t = *reinterpret_cast<const float4*>(Zs + n * ZS + warp * JPW + m);
            z[m] = t.x; z[m + 1] = t.y; z[m + 2] = t.z; z[m + 3] = t.w;
        }
#pragma unroll
        for (int q = 0; q < KPT; ++q)
#pragma unroll
            for (int m = 0; m < JPW; ++m) acc2[q][m] = fmaf(r[q], z[m], acc2[q][m]);
    }
}

// ---- pass 1 -------------------------------------------------------------------------------
template <int KPT, int JPW>
__device__ void ridge_flush_moments(const HmyDev& st, int combo, float (&acc2)[KPT][JPW], float& gsum) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = st.V, n1 = st.B + 1;
    int lev[HMY_MAX_V];
    for (int v = 0; v < V; ++v) lev[v] = st.combo_lev[combo * V + v];
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
        const int k = lane + 32 * q;
#pragma unroll
        for (int m = 0; m < JPW; ++m) {
            const int j = warp * JPW + m;
            const float a = acc2[q][m];
            if (k < st.K && j < st.d && a != 0.f) {
                atomicAdd(&st.Mom[((size_t)0 * st.K + k) * st.dp + j], (double)a);
                for (int v = 0; v < V; ++v) atomicAdd(&st.Mom[((size_t)(1 + lev[v]) * st.K + k) * st.dp + j], (double)a);
            }
            acc2[q][m] = 0.f;
        }
    }
    if ((int)threadIdx.x < st.K && gsum != 0.f) {
        double* A = st.Gram + (size_t)threadIdx.x * n1 * n1;
        const double g = (double)gsum;
        atomicAdd(&A[0], g);
        for (int v = 0; v < V; ++v) {
            atomicAdd(&A[1 + lev[v]], g);
            atomicAdd(&A[(size_t)(1 + lev[v]) * n1], g);
            for (int u = 0; u < V; ++u) atomicAdd(&A[(size_t)(1 + lev[v]) * n1 + 1 + lev[u]], g);
        }
    }
    gsum = 0.f;
}

template <int KPT, int JPW>
__global__ void __launch_bounds__(HMY_THREADS) k_ridge_moments(HmyDev st) {
    extern __shared__ __align__(16) unsigned char smem[];
    const RidgeSmem p = ridge_smem_plan(st.K, st.KS, JPW, false);
    float* Zs = (float*)(smem + p.off_Zs); float* Rs = (float*)(smem + p.off_Rs);
    for (int i = threadIdx.x; i < HMY_TILE * p.ZS; i += HMY_THREADS) Zs[i] = 0.f;
    for (int i = threadIdx.x; i < HMY_TILE * p.RS; i += HMY_THREADS) Rs[i] = 0.f;
    __syncthreads();
    float acc2[KPT][JPW];
#pragma unroll
    for (int q = 0; q < KPT; ++q)
#pragma unroll
        for (int m = 0; m < JPW; ++m) acc2[q][m] = 0.f;
    float gsum = 0.f;
    int cur = -1;
    const int i0 = (int)((long long)blockIdx.x * st.nseg / gridDim.x), i1 = (int)((long long)(blockIdx.x + 1) * st.nseg / gridDim.x);
    for (int it = i0; it < i1; ++it) {
        const long long start = st.seg[3 * it]; const int count = st.seg[3 * it + 1], combo = st.seg[3 * it + 2];
        if (combo != cur) { if (cur >= 0) ridge_flush_moments<KPT, JPW>(st, cur, acc2, gsum); cur = combo; }
        for (int t = 0; t < count; t += HMY_TILE) {
            const int nt = min(HMY_TILE, count - t);
            ridge_load_tile(st, st.Zorig, start + t, nt, Zs, p.ZS, Rs, p.RS);
            __syncthreads();
            rtz_accumulate<KPT, JPW>(acc2, Rs, p.RS, Zs, p.ZS, nt);
            if ((int)threadIdx.x < st.K) {
                float g = 0.f;
                for (int n = 0; n < nt; ++n) g += Rs[n * p.RS + threadIdx.x];
                gsum += g;
            }
            __syncthreads();
        }
    }
    if (cur >= 0) ridge_flush_moments<KPT, JPW>(st, cur, acc2, gsum);
}

#ifdef HMY_NONTEMPLATE_KERNELS
// ---- solve --------------------------------------------------------------------------------
// One CTA per cluster: Gauss-Jordan with partial pivoting on [Gram + diag(lambda) | Mom] in
// fp64 (hard part 1c: the Gram has cond ~1e3 and fp32 serial sums are not
// accurate enough).  Equivalent to inv(cov) @ [sums] of harmony.py:553-563.
__global__ void __launch_bounds__(128) k_ridge_solve(HmyDev st) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int k = blockIdx.x, n = st.B + 1, d = st.d, m = n + d;
    // the augmented system lives in shared memory, or -- hundreds of batch levels -- in this cluster's slice of a global
    // work area (same algorithm; the solve is off the hot path: K small systems per Harmony iteration)
    double* A = st.solve_scratch ? st.solve_scratch + (size_t)k * ((size_t)n * m + n) : (double*)smem;     // [n][m]
    double* fac = A + (size_t)n * m;         // [n]
    __shared__ int s_piv;
    __shared__ double s_rowsum;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < st.lev0; ++b) s += st.O[b * st.K + k];
        s_rowsum = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * m; i += blockDim.x) {
        const int r = i / m, c = i - r * m;
        double v;
        if (c < n) {
            v = st.Gram[((size_t)k * n + r) * n + c];
            if (r == c) {
                double lam;
                if (st.lambda_estimation) lam = (r == 0) ? 0.0 : (double)((float)(s_rowsum * (double)st.Pr_b[r - 1]) * st.alpha);
                else lam = (double)st.lamb[r];
                v += lam;
            }
        } else {
            v = st.Mom[((size_t)r * st.K + k) * st.dp + (c - n)];
        }
        A[i] = v;
    }
    __syncthreads();
    for (int col = 0; col < n; ++col) {
        if (threadIdx.x == 0) {
            int best = col; double bv = fabs(A[(size_t)col * m + col]);
            for (int r = col + 1; r < n; ++r) { const double v = fabs(A[(size_t)r * m + col]); if (v > bv) { bv = v; best = r; } }
            s_piv = best;
        }
        __syncthreads();
        const int pr = s_piv;
        if (pr != col)
            for (int c = threadIdx.x; c < m; c += blockDim.x) { const double t = A[(size_t)col * m + c]; A[(size_t)col * m + c] = A[(size_t)pr * m + c]; A[(size_t)pr * m + c] = t; }
        __syncthreads();
        const double inv = 1.0 / A[(size_t)col * m + col];
        __syncthreads();
        for (int c = threadIdx.x; c < m; c += blockDim.x) A[(size_t)col * m + c] *= inv;
        for (int r = threadIdx.x; r < n; r += blockDim.x) fac[r] = A[(size_t)r * m + col];
        __syncthreads();
        for (int i = threadIdx.x; i < n * (m - col - 1); i += blockDim.x) {
            const int r = i / (m - col - 1), c = col + 1 + (i - r * (m - col - 1));
            if (r != col) A[(size_t)r * m + c] -= fac[r] * A[(size_t)col * m + c];
        }
        __syncthreads();
    }
    // W[b][k][j]; the intercept row is dropped (harmony.py:565)
    float wm = 0.f;
    for (int i = threadIdx.x; i < st.B * st.dp; i += blockDim.x) {
        const int b = i / st.dp, j = i - b * st.dp;
        const float w = (j < d) ? (float)A[(size_t)(1 + b) * m + n + j] : 0.f;
        st.W[((size_t)b * st.K + k) * st.dp + j] = w;
        if (isfinite(w)) wm = fmaxf(wm, fabsf(w));
    }
    // largest |coefficient| (scale of the fp16 coefficient tile of the tensor-core apply pass)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    if ((threadIdx.x & 31) == 0 && st.wmax != nullptr) atomicMax(reinterpret_cast<unsigned int*>(st.wmax), __float_as_uint(wm));
}

#endif  // HMY_NONTEMPLATE_KERNELS

// ---- pass 2 -------------------------------------------------------------------------------
template <int KPT, int JPW>
__global__ void __launch_bounds__(HMY_THREADS) k_ridge_apply(HmyDev st) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int JPL = JPW / 4;
    const RidgeSmem p = ridge_smem_plan(st.K, st.KS, JPW, true);
    float* Zs = (float*)(smem + p.off_Zs); float* Rs = (float*)(smem + p.off_Rs); float* Wc = (float*)(smem + p.off_Wc);
    const int ZS = p.ZS, RS = p.RS, WS = p.WS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < HMY_TILE * ZS; i += HMY_THREADS) Zs[i] = 0.f;
    for (int i = threadIdx.x; i < HMY_TILE * RS; i += HMY_THREADS) Rs[i] = 0.f;
    __syncthreads();
    float acc2[KPT][JPW];
#pragma unroll
    for (int q = 0; q < KPT; ++q)
#pragma unroll
        for (int m = 0; m < JPW; ++m) acc2[q][m] = 0.f;
    int cur = -1;
    const int i0 = (int)((long long)blockIdx.x * st.nseg / gridDim.x), i1 = (int)((long long)(blockIdx.x + 1) * st.nseg / gridDim.x);
    const int K4 = (st.K + 3) & ~3;
    for (int it = i0; it < i1; ++it) {
        const long long start = st.seg[3 * it]; const int count = st.seg[3 * it + 1], combo = st.seg[3 * it + 2];
        if (combo != cur) {
            // Wc[k][j] = sum_v W[level_v][k][j]: what W.T @ Phi_Rk picks for this combination
            __syncthreads();
            for (int i = threadIdx.x; i < st.K * WS; i += HMY_THREADS) {
                const int k = i / WS, j = i - k * WS;
                float w = 0.f;
                if (j < st.dp)
                    for (int v = 0; v < st.V; ++v) w += st.W[((size_t)st.combo_lev[combo * st.V + v] * st.K + k) * st.dp + j];
                Wc[i] = w;
            }
            cur = combo;
            __syncthreads();
        }
        for (int t = 0; t < count; t += HMY_TILE) {
            const int nt = min(HMY_TILE, count - t);
            const long long base = start + t;
            ridge_load_tile(st, st.Zorig, base, nt, Zs, ZS, Rs, RS);
            __syncthreads();
            // corr[cell][j] = sum_k R[cell][k] Wc[k][j]; warp: HMY_CPW cells, lane: j = lane + 32 m
            float acc[HMY_CPW][JPL];
#pragma unroll
            for (int i = 0; i < HMY_CPW; ++i)
#pragma unroll
                for (int mm = 0; mm < JPL; ++mm) acc[i][mm] = 0.f;
            const float* rbase = Rs + warp * HMY_CPW * RS;
            for (int k = 0; k < K4; k += 4) {
                float4 rv[HMY_CPW];
#pragma unroll
                for (int i = 0; i < HMY_CPW; ++i) rv[i] = *reinterpret_cast<const float4*>(rbase + i * RS + k);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (k + kk < st.K) {
                        float w[JPL];
#pragma unroll
                        for (int mm = 0; mm < JPL; ++mm) w[mm] = Wc[(k + kk) * WS + lane + 32 * mm];
#pragma unroll
                        for (int i = 0; i < HMY_CPW; ++i) {
                            const float rr = (kk == 0) ? rv[i].x : (kk == 1) ? rv[i].y : (kk == 2) ? rv[i].z : rv[i].w;
#pragma unroll
                            for (int mm = 0; mm < JPL; ++mm) acc[i][mm] = fmaf(rr, w[mm], acc[i][mm]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < HMY_CPW; ++i) {
                const int cl = warp * HMY_CPW + i;
                float zc[JPL]; float ss = 0.f;
#pragma unroll
                for (int mm = 0; mm < JPL; ++mm) {
                    const int j = lane + 32 * mm;
                    zc[mm] = (j < st.dp) ? Zs[cl * ZS + j] - acc[i][mm] : 0.f;
                    ss += zc[mm] * zc[mm];
                }
                ss = warp_sum(ss);
                const float nrm = sqrtf(ss);
                if (cl < nt) {
#pragma unroll
                    for (int mm = 0; mm < JPL; ++mm) {
                        const int j = lane + 32 * mm;
                        if (j < st.dp) {
                            const float zn = zc[mm] / nrm;
                            st.Zcorr[(size_t)(base + cl) * st.dp + j] = zc[mm];
                            st.Zcos[(size_t)(base + cl) * st.dp + j] = zn;
                            Zs[cl * ZS + j] = zn;
                        }
                    }
                }
            }
            __syncthreads();
            rtz_accumulate<KPT, JPW>(acc2, Rs, RS, Zs, ZS, nt);
            __syncthreads();
        }
    }
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
        const int k = lane + 32 * q;
#pragma unroll
        for (int m = 0; m < JPW; ++m) {
            const int j = warp * JPW + m;
            if (k < st.K && j < st.d && acc2[q][m] != 0.f) atomicAdd(&st.Yacc[(size_t)k * st.dp + j], (double)acc2[q][m]);
        }
    }
}

#ifdef HMY_NONTEMPLATE_KERNELS
// ---- ingest / egress ----------------------------------------------------------------------
// one warp per cell: gather the caller's row into the sorted, padded layout; Z_cos (harmony.py:238)
__global__ void k_ingest(HmyDev st, const float* Zraw, float* zmax) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= st.N) return;
    const float* src = Zraw + (size_t)st.order[p] * st.d;
    float ss = 0.f, zm = 0.f;
    for (int j = lane; j < st.d; j += 32) { const float z = src[j]; ss += z * z; if (isfinite(z)) zm = fmaxf(zm, fabsf(z)); }
    // largest |z| of the upload (scale of the fp16 split in the tensor-core ridge passes)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
    if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(zmax), __float_as_uint(zm));
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    for (int j = lane; j < st.dp; j += 32) {
        const float z = (j < st.d) ? src[j] : 0.f;
        st.Zorig[(size_t)p * st.dp + j] = z;
        st.Zcorr[(size_t)p * st.dp + j] = z;
        st.Zcos[(size_t)p * st.dp + j] = z / nrm;
    }
}

// back to the state right after hmy_set_data: Z_corr = Z_orig, Z_cos = unit(Z_orig) (harmony.py:234-238)
__global__ void k_reset(HmyDev st) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= st.N) return;
    const float* src = st.Zorig + (size_t)p * st.dp;
    float ss = 0.f;
    for (int j = lane; j < st.dp; j += 32) { const float z = src[j]; ss += z * z; }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    for (int j = lane; j < st.dp; j += 32) {
        const float z = src[j];
        st.Zcorr[(size_t)p * st.dp + j] = z;
        st.Zcos[(size_t)p * st.dp + j] = z / nrm;
    }
}

// Z_cos -> the pre-split operand rows of the tensor-memory round kernel (hmy_round_tc5.cuh): per cell
// hi[16 dt] | lo[16 dt] halves of 1024 * z (zero padded), one thread per pair of PCs
__global__ void k_split_zcos(HmyDev st) {
    const int dt = (st.d + 15) >> 4, pairs = 8 * dt;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.N * pairs) return;
    const long long p = i / pairs;
    const int j2 = (int)(i - p * pairs), j = 2 * j2;
    const float* z = st.Zcos + (size_t)p * st.dp;
    const float x0 = (j < st.dp) ? z[j] * 1024.0f : 0.f, x1 = (j + 1 < st.dp) ? z[j + 1] * 1024.0f : 0.f;
    const __half2 H = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(H);
    const __half2 L = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    unsigned int* row = reinterpret_cast<unsigned int*>(st.Zs16) + (size_t)p * (2 * pairs);
    row[j2] = *reinterpret_cast<const unsigned int*>(&H);
    row[pairs + j2] = *reinterpret_cast<const unsigned int*>(&L);
}

// dst[order[p]][0..w) = src[p][0..w)   (src row stride sp)
__global__ void k_unsort_rows(const float* src, int sp, float* dst, int w, const int* order, long long N) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= N) return;
    const float* s = src + (size_t)p * sp;
    float* o = dst + (size_t)order[p] * w;
    for (int j = lane; j < w; j += 32) o[j] = s[j];
}
#endif  // HMY_NONTEMPLATE_KERNELS
