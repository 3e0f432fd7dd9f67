// compute_lisi on the GPU (reference: harmonypy/lisi.py).
//
// STATUS: written at the end of round 1 without GPU time left; the CPU oracle it is to be compared with
// (oracle/lisi_oracle.py) is pinned to the reference's known-answer test, this file has only been compiled.
// The Python entry point (harmonypy_b200/lisi.py) works, its GPU tests run with HMY_TEST_LISI=1.
//
//   k_lisi_knn      exact k nearest neighbours of every cell among all cells (Euclidean, fp64 like the
//                   reference's sklearn kd_tree on a float64 matrix, lisi.py:53-54), one warp per query,
//                   candidates staged through shared memory, the k best kept as a sorted list per warp
//   k_lisi_simpson  one warp per cell: bisection on beta until the entropy of exp(-beta * dist) equals
//                   log(perplexity) (lisi.py:79-122, same control flow, fp64) and, per label column,
//                   sum over categories of (sum of the weights of the neighbours in the category)^2
//                   (lisi.py:127-132); LISI = 1 / that (lisi.py:64)
//
// Brute force is O(n^2 d): meant for the sizes the metric is used at (1e3 .. 2e5 cells).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/harmony_b200.h"

extern "C" void harmony_b200_set_global_error(const char* msg);     // hmy_api.cu: what hmy_last_error(NULL) returns

#define LISI_MAXK 128          // neighbours kept per cell, self included (3 * perplexity <= 128)
#define LISI_TILE 128          // candidate rows staged per step
#define LISI_WARPS 8
#define LISI_MAXD 128

__device__ __forceinline__ double lisi_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dist2 / index lists are sorted ascending by (dist2, index)
__global__ void __launch_bounds__(32 * LISI_WARPS) k_lisi_knn(const double* __restrict__ X, long long n, int d, int k,
                                                              double* __restrict__ dist_out, int* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ds = d | 1;                                        // odd row stride (in doubles): fewer bank conflicts
    double* sC = reinterpret_cast<double*>(smem_raw);            // [LISI_TILE][ds] candidates
    double* sQ = sC + (size_t)LISI_TILE * ds;                    // [LISI_WARPS][ds] queries
    double* sLd = sQ + (size_t)LISI_WARPS * ds;                  // [LISI_WARPS][LISI_MAXK]
    int* sLi = reinterpret_cast<int*>(sLd + LISI_WARPS * LISI_MAXK);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long q = (long long)blockIdx.x * LISI_WARPS + warp;
    const bool have_q = q < n;
    double* Ld = sLd + warp * LISI_MAXK;
    int* Li = sLi + warp * LISI_MAXK;
    if (have_q) for (int j = lane; j < d; j += 32) sQ[warp * ds + j] = X[(size_t)q * d + j];
    int cnt = 0;
    double thr = INFINITY;
    __syncthreads();
    for (long long t0 = 0; t0 < n; t0 += LISI_TILE) {
        const int tn = (int)min((long long)LISI_TILE, n - t0);
        for (int i = tid; i < tn * d; i += 32 * LISI_WARPS) { const int r = i / d, j = i - r * d; sC[r * ds + j] = X[(size_t)(t0 + r) * d + j]; }
        __syncthreads();
        if (have_q) {
            for (int c0 = 0; c0 < tn; c0 += 32) {
                const int c = c0 + lane;
                double d2 = INFINITY;
                if (c < tn) {
                    const double* xc = sC + c * ds;
                    const double* xq = sQ + warp * ds;
                    double s = 0.0;
                    for (int j = 0; j < d; ++j) { const double u = xq[j] - xc[j]; s = fma(u, u, s); }
                    d2 = s;
                }
                unsigned int mask = __ballot_sync(0xffffffffu, c < tn && (cnt < k || d2 < thr));
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    const double val = __shfl_sync(0xffffffffu, d2, src);
                    const int id = (int)(t0 + c0 + src);
                    if (cnt == k && !(val < thr)) continue;          // the threshold moved since the ballot
                    // position = number of entries that sort before (val, id)
                    int before = 0;
#pragma unroll
                    for (int u = 0; u < LISI_MAXK / 32; ++u) {
                        const int e = lane + 32 * u;
                        if (e < cnt) before += (Ld[e] < val || (Ld[e] == val && Li[e] < id)) ? 1 : 0;
                    }
                    const int p = __reduce_add_sync(0xffffffffu, before);
                    const int newlen = min(cnt + 1, k);
                    // shift [p, newlen - 1) one to the right: read first, then write
                    double mv_d[LISI_MAXK / 32]; int mv_i[LISI_MAXK / 32];
#pragma unroll
                    for (int u = 0; u < LISI_MAXK / 32; ++u) {
                        const int e = lane + 32 * u;                 // target slot
                        if (e > p && e < newlen) { mv_d[u] = Ld[e - 1]; mv_i[u] = Li[e - 1]; }
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < LISI_MAXK / 32; ++u) {
                        const int e = lane + 32 * u;
                        if (e > p && e < newlen) { Ld[e] = mv_d[u]; Li[e] = mv_i[u]; }
                    }
                    if (lane == 0 && p < newlen) { Ld[p] = val; Li[p] = id; }
                    __syncwarp();
                    cnt = newlen;
                    if (cnt == k) thr = Ld[k - 1];
                }
            }
        }
        __syncthreads();
    }
    if (have_q)
        for (int e = lane; e < k; e += 32) { dist_out[(size_t)q * k + e] = sqrt(Ld[e]); idx_out[(size_t)q * k + e] = Li[e]; }
}

// dist / idx: [n][k] from k_lisi_knn (entry 0 is dropped like lisi.py:56-57); codes: [n_labels][n]; out: [n][n_labels]
__global__ void __launch_bounds__(32 * LISI_WARPS) k_lisi_simpson(const double* __restrict__ dist, const int* __restrict__ idx, long long n, int k,
                                                                  const int* __restrict__ codes, int n_labels, double perplexity,
                                                                  double* __restrict__ out) {
    __shared__ double sP[LISI_WARPS][LISI_MAXK];
    __shared__ int sLab[LISI_WARPS][LISI_MAXK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * LISI_WARPS + warp;
    if (cell >= n) return;                                       // whole warps leave together; no CTA-wide sync below
    const int m = k - 1;                                         // neighbours without the cell itself
    constexpr int PER = LISI_MAXK / 32;
    double D[PER], P[PER];
    int nb[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int j = lane + 32 * u;
        D[u] = (j < m) ? dist[(size_t)cell * k + 1 + j] : 0.0;
        nb[u] = (j < m) ? idx[(size_t)cell * k + 1 + j] : 0;
        P[u] = 0.0;
    }
    const double logU = log(perplexity), tol = 1e-5;
    double beta = 1.0, betamin = -INFINITY, betamax = INFINITY, H = 0.0;
    auto evaluate = [&]() {                                      // lisi.py:85-93 / :113-121
        double ps = 0.0, dp = 0.0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int j = lane + 32 * u;
            P[u] = (j < m) ? exp(-D[u] * beta) : 0.0;
            ps += P[u]; dp += D[u] * P[u];
        }
        ps = lisi_warp_sum(ps); dp = lisi_warp_sum(dp);
        if (ps == 0.0) {
            H = 0.0;
#pragma unroll
            for (int u = 0; u < PER; ++u) P[u] = 0.0;
        } else {
            H = log(ps) + beta * dp / ps;
#pragma unroll
            for (int u = 0; u < PER; ++u) P[u] /= ps;
        }
    };
    evaluate();
    double Hdiff = H - logU;
    for (int t = 0; t < 50; ++t) {                               // lisi.py:95-122 (all lanes hold identical scalars)
        if (fabs(Hdiff) < tol) break;
        if (Hdiff > 0) {
            betamin = beta;
            beta = isfinite(betamax) ? (beta + betamax) / 2 : beta * 2;
        } else {
            betamax = beta;
            beta = isfinite(betamin) ? (beta + betamin) / 2 : beta / 2;
        }
        evaluate();
        Hdiff = H - logU;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) { const int j = lane + 32 * u; if (j < m) sP[warp][j] = P[u]; }
    for (int l = 0; l < n_labels; ++l) {
        __syncwarp();
#pragma unroll
        for (int u = 0; u < PER; ++u) { const int j = lane + 32 * u; if (j < m) sLab[warp][j] = codes[(size_t)l * n + nb[u]]; }
        __syncwarp();
        // sum_c (sum_{i in c} P_i)^2 = sum_j P_j * (sum of P_i over the neighbours with j's label)
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int j = lane + 32 * u;
            if (j < m) {
                const int mine = sLab[warp][j];
                double s = 0.0;
                for (int i = 0; i < m; ++i) s += (sLab[warp][i] == mine) ? sP[warp][i] : 0.0;
                acc += P[u] * s;
            }
        }
        acc = lisi_warp_sum(acc);
        const double simpson = ((H == 0.0) ? -1.0 : 0.0) + acc;  // lisi.py:123-132
        if (lane == 0) out[(size_t)cell * n_labels + l] = 1.0 / simpson;
    }
}

#define LCK(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            char b_[512];                                                                                      \
            snprintf(b_, sizeof b_, "hmy_lisi_compute: %s failed: %s", #call, cudaGetErrorString(e_));          \
            harmony_b200_set_global_error(b_);                                                                 \
            for (void* p_ : allocs) cudaFree(p_);                                                              \
            return 1;                                                                                          \
        }                                                                                                      \
    } while (0)
#define LFAIL(msg) do { harmony_b200_set_global_error("hmy_lisi_compute: " msg); for (void* p_ : allocs) cudaFree(p_); return 1; } while (0)

extern "C" int hmy_lisi_compute(int device, int64_t n, int d, const double* X_host, int n_labels,
                                const int32_t* codes_host, double perplexity, double* out_host) {
    std::vector<void*> allocs;
    const int k = (int)(perplexity * 3.0);                       // lisi.py:53
    if (!X_host || !codes_host || !out_host) LFAIL("NULL argument");
    if (n < 2 || n > 2000000000LL) LFAIL("n out of range");
    if (d < 1 || d > LISI_MAXD) LFAIL("d must be in 1..128");
    if (n_labels < 1) LFAIL("needs at least one label column");
    if (k < 2 || k > LISI_MAXK) LFAIL("3 * perplexity must be in 2..128");
    if (n < k) LFAIL("fewer cells than 3 * perplexity neighbours");    // sklearn raises here too
    LCK(cudaSetDevice(device));
    double *dX = nullptr, *dDist = nullptr, *dOut = nullptr; int *dIdx = nullptr, *dCodes = nullptr;
    LCK(cudaMalloc((void**)&dX, (size_t)n * d * sizeof(double))); allocs.push_back(dX);
    LCK(cudaMalloc((void**)&dDist, (size_t)n * k * sizeof(double))); allocs.push_back(dDist);
    LCK(cudaMalloc((void**)&dIdx, (size_t)n * k * sizeof(int))); allocs.push_back(dIdx);
    LCK(cudaMalloc((void**)&dCodes, (size_t)n * n_labels * sizeof(int))); allocs.push_back(dCodes);
    LCK(cudaMalloc((void**)&dOut, (size_t)n * n_labels * sizeof(double))); allocs.push_back(dOut);
    LCK(cudaMemcpy(dX, X_host, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice));
    LCK(cudaMemcpy(dCodes, codes_host, (size_t)n * n_labels * sizeof(int), cudaMemcpyHostToDevice));
    const int ds = d | 1;
    const size_t smem = ((size_t)LISI_TILE * ds + (size_t)LISI_WARPS * ds + (size_t)LISI_WARPS * LISI_MAXK) * sizeof(double)
                        + (size_t)LISI_WARPS * LISI_MAXK * sizeof(int);
    LCK(cudaFuncSetAttribute((const void*)k_lisi_knn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned int grid = (unsigned int)((n + LISI_WARPS - 1) / LISI_WARPS);
    k_lisi_knn<<<grid, 32 * LISI_WARPS, smem>>>(dX, (long long)n, d, k, dDist, dIdx);
    LCK(cudaGetLastError());
    k_lisi_simpson<<<grid, 32 * LISI_WARPS>>>(dDist, dIdx, (long long)n, k, dCodes, n_labels, perplexity, dOut);
    LCK(cudaGetLastError());
    LCK(cudaMemcpy(out_host, dOut, (size_t)n * n_labels * sizeof(double), cudaMemcpyDeviceToHost));
    for (void* p : allocs) cudaFree(p);
    return 0;
}
