// k-means++ / Lloyd initialisation of the centroids on the device: the optional
// replacement of the reference's sklearn call (harmony.py:369-373: KMeans(init="k-means++", n_init=1, max_iter=25)
// on the unit-length cells) for sizes where that call dominates the wall time.  It cannot reproduce sklearn's random
// stream, so the parity configurations keep sklearn; what it is compared with is oracle/kmeans_init_oracle.py, the
// same algorithm in NumPy with the same counter-based random numbers.
//
// STATUS: written at the end of round 1 without GPU time left; compiled only, GPU tests opt-in (HMY_TEST_KMINIT=1).
//
// Seeding (k-means++, Arthur & Vassilvitskii): centre 0 uniformly; centre c with probability proportional to the
// squared distance to the nearest centre so far.  The draw is an exponential race, argmin_i E_i / D2_i with
// E_i = -log(u(seed, c, cell id)): one min-reduction, no prefix sums, and independent of how the cells are laid out
// in memory (the engine stores them sorted by covariate combination).  Lloyd: one warp per cell, lanes over clusters,
// per-CTA shared-memory sums, fp64 global sums; stops at sklearn's criterion (squared centre shift <=
// tol * mean feature variance) or after max_iter iterations.
#pragma once
#include "hmy_common.cuh"

#define HMY_KMI_THREADS 256

__host__ __device__ inline unsigned long long hmy_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform in (0, 1] from (seed, step, cell id)
__host__ __device__ inline double hmy_kmi_uniform(unsigned long long seed, int step, unsigned long long id) {
    const unsigned long long h = hmy_splitmix64(hmy_splitmix64(hmy_splitmix64(seed) + (unsigned long long)step) ^ id);
    return (double)((h >> 11) + 1ull) * (1.0 / 9007199254740992.0);
}

#ifdef HMY_NONTEMPLATE_KERNELS

// One seeding step: D2 to centre `c - 1` (row `prev_pos` of Z_cos), running minimum, race for centre c.
// best[c] must be ~0ull on entry; afterwards its low 32 bits are the winner's position.
__global__ void __launch_bounds__(HMY_KMI_THREADS) k_kmpp_pass(HmyDev st, const unsigned long long* best, long long first_pos, int c,
                                                               unsigned long long seed, float* mind2, unsigned long long* best_out) {
    extern __shared__ __align__(16) unsigned char kmi_smem[];
    float* sy = reinterpret_cast<float*>(kmi_smem);                       // [dp] the newest centre
    __shared__ unsigned long long s_best[HMY_KMI_THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, dp = st.dp;
    const long long prev_pos = (c == 1) ? first_pos : (long long)(best[c - 1] & 0xFFFFFFFFull);
    if (prev_pos >= st.N) return;          // the previous step found no cell at a positive distance (host reports it)
    for (int j = tid; j < dp; j += HMY_KMI_THREADS) sy[j] = st.Zcos[(size_t)prev_pos * dp + j];
    __syncthreads();
    unsigned long long mine = ~0ull;
    const long long nw = (long long)gridDim.x * (HMY_KMI_THREADS / 32);
    for (long long pos = (long long)blockIdx.x * (HMY_KMI_THREADS / 32) + warp; pos < st.N; pos += nw) {
        const float* z = st.Zcos + (size_t)pos * dp;
        float s = 0.f;
        for (int j = lane; j < dp; j += 32) { const float u = z[j] - sy[j]; s = fmaf(u, u, s); }
        s = warp_sum(s);
        if (lane == 0) {
            const float m = (c == 1) ? s : fminf(mind2[pos], s);
            mind2[pos] = m;
            if (m > 0.f) {
                const unsigned long long id = (unsigned long long)(st.cell_offset + st.order[pos]);
                const float key = (float)(-log(hmy_kmi_uniform(seed, c, id)) / (double)m);
                const unsigned long long packed = ((unsigned long long)__float_as_uint(key) << 32) | (unsigned long long)(unsigned int)pos;
                mine = packed < mine ? packed : mine;
            }
        }
    }
    if (lane == 0) s_best[warp] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = s_best[0];
        for (int w = 1; w < HMY_KMI_THREADS / 32; ++w) b = s_best[w] < b ? s_best[w] : b;
        if (b != ~0ull) atomicMin(&best_out[c], b);
    }
}

// rows of the chosen cells -> C [K][dp]
__global__ void k_kmpp_gather(HmyDev st, const unsigned long long* best, long long first_pos, float* C) {
    const int k = blockIdx.x;
    const long long pos = (k == 0) ? first_pos : (long long)(best[k] & 0xFFFFFFFFull);
    if (pos >= st.N) return;
    for (int j = threadIdx.x; j < st.dp; j += blockDim.x) C[(size_t)k * st.dp + j] = st.Zcos[(size_t)pos * st.dp + j];
}

// One Lloyd assignment pass.  sums [K][dp] fp64, counts [K], inertia [1] must be zero on entry.
__global__ void __launch_bounds__(HMY_KMI_THREADS) k_lloyd_assign(HmyDev st, const float* __restrict__ C, double* sums, unsigned int* counts, double* inertia) {
    extern __shared__ __align__(16) unsigned char kmi_smem[];
    const int K = st.K, dp = st.dp, dps = dp | 1;
    float* sC = reinterpret_cast<float*>(kmi_smem);                       // [K][dps]
    float* sN = sC + (size_t)K * dps;                                     // [K] squared norms
    float* sSum = sN + K;                                                 // [K][dp]
    float* sZ = sSum + (size_t)K * dp;                                    // [warps][dp]
    unsigned int* sCnt = reinterpret_cast<unsigned int*>(sZ + (HMY_KMI_THREADS / 32) * dp);   // [K]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < K * dp; i += HMY_KMI_THREADS) { const int k = i / dp, j = i - k * dp; sC[k * dps + j] = C[i]; sSum[i] = 0.f; }
    for (int k = tid; k < K; k += HMY_KMI_THREADS) sCnt[k] = 0u;
    __syncthreads();
    for (int k = tid; k < K; k += HMY_KMI_THREADS) { float s = 0.f; for (int j = 0; j < dp; ++j) s = fmaf(sC[k * dps + j], sC[k * dps + j], s); sN[k] = s; }
    __syncthreads();
    float* mz = sZ + warp * dp;
    double my_inertia = 0.0;
    const long long nw = (long long)gridDim.x * (HMY_KMI_THREADS / 32);
    for (long long pos = (long long)blockIdx.x * (HMY_KMI_THREADS / 32) + warp; pos < st.N; pos += nw) {
        const float* z = st.Zcos + (size_t)pos * dp;
        float zn = 0.f;
        for (int j = lane; j < dp; j += 32) { const float v = z[j]; mz[j] = v; zn = fmaf(v, v, zn); }
        zn = warp_sum(zn);
        __syncwarp();
        float bestv = INFINITY; int bestk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            const float* ck = sC + k * dps;
            float dot = 0.f;
            for (int j = 0; j < dp; ++j) dot = fmaf(mz[j], ck[j], dot);
            const float v = sN[k] - 2.f * dot;                            // ||z - c||^2 - ||z||^2
            if (v < bestv) { bestv = v; bestk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
            if (ov < bestv || (ov == bestv && ok < bestk)) { bestv = ov; bestk = ok; }
        }
        for (int j = lane; j < dp; j += 32) atomicAdd(&sSum[bestk * dp + j], mz[j]);
        if (lane == 0) { atomicAdd(&sCnt[bestk], 1u); my_inertia += (double)(bestv + zn); }
        __syncwarp();
    }
    __syncthreads();
    for (int i = tid; i < K * dp; i += HMY_KMI_THREADS) if (sSum[i] != 0.f) atomicAdd(&sums[i], (double)sSum[i]);
    for (int k = tid; k < K; k += HMY_KMI_THREADS) if (sCnt[k]) atomicAdd(&counts[k], sCnt[k]);
    if (lane == 0 && my_inertia != 0.0) atomicAdd(inertia, my_inertia);
}

// new centres = means (an empty cluster keeps its centre); info[0] = squared shift, info[1] = mean feature variance
__global__ void k_lloyd_update(HmyDev st, float* C, const double* sums, const unsigned int* counts, double* info) {
    __shared__ double s_shift[32], s_mean2[32];
    const int K = st.K, dp = st.dp, d = st.d, tid = threadIdx.x;
    double shift = 0.0, mean2 = 0.0;
    for (int j = tid; j < d; j += blockDim.x) {                           // ||mean of all cells||^2, one feature per thread
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += sums[(size_t)k * dp + j];
        s /= (double)st.N;
        mean2 += s * s;
    }
    for (int i = tid; i < K * dp; i += blockDim.x) {
        const int k = i / dp, j = i - k * dp;
        if (j < d && counts[k] > 0u) {
            const float nc = (float)(sums[i] / (double)counts[k]);
            const double df = (double)nc - (double)C[i];
            shift += df * df;
            C[i] = nc;
        }
    }
    shift = warp_sum_d(shift); mean2 = warp_sum_d(mean2);
    if ((tid & 31) == 0) { s_shift[tid >> 5] = shift; s_mean2[tid >> 5] = mean2; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_shift[w]; b += s_mean2[w]; }
        info[0] = a;
        info[1] = (1.0 - b) / (double)d;                                  // rows of Z_cos have unit length: E|z|^2 = 1
    }
}

#endif  // HMY_NONTEMPLATE_KERNELS
