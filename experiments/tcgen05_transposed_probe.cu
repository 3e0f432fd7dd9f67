// EXPERIMENT (not product code): second probe for the next round's TMEM pipeline -- the centroid GEMM.
//
//   D[128 x 112] (fp32, TMEM, accumulated over TWO tiles of 128 cells) += A^T . B   with
//   A^T: M = 128 rows (64 PCs + 32 one-hot block rows + 32 zero rows) x K = 128 cells, MN-major
//   B  : K = 128 cells x N = 112 clusters, N-major ("MN-major" B)
// and both operands given as fp16 hi + lo parts: D += Ah.Bh + Al.Bh + Ah.Bl (3 MMAs per K = 16 step).
// The shared-memory tiles use the same 8 x 16-byte core matrices as the K-major probe; only the
// descriptors (a_major = b_major = 1, LBO / SBO) change -- i.e. ONE copy of the Z tile in shared memory
// can feed the scoring GEMM (K-major A) and this GEMM (MN-major A).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_t experiments/tcgen05_transposed_probe.cu && timeout 60 ./probe_t
//
// Run on a B200 at the very end of round 1: max |err| / sum|terms| = 3.545e-07 (PASS).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 112, KT = 128;          // KT cells per tile
constexpr int TILES = 2;
constexpr int TMEM_COLS = 128;

// MN-major no-swizzle: element (mn, k) lives in core matrix (mn/8, k/8); inside it k%8 selects the 16-byte
// row and mn%8 the element.  Core matrices of one mn-block are contiguous over k-blocks.
__host__ __device__ inline int canon_mn_off_bytes(int mn, int k) { return ((mn >> 3) * (KT / 8) + (k >> 3)) * 128 + (k & 7) * 16 + (mn & 7) * 2; }
constexpr uint32_t LBO = 128, SBO = (KT / 8) * 128;   // between k-blocks / between mn-blocks

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ inline uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 1u << 15;                      // a_major = MN
    d |= 1u << 16;                      // b_major = MN
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ inline void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ inline void wait_parity(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// At/Bt: [TILES][KT][M] resp. [TILES][KT][N] row-major fp32 (cell-major, like the engine's Z and R rows)
__global__ void __launch_bounds__(128) probe_t(const float* At, const float* Bt, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sAh = smem;                       // M x KT halves = 32 KB each
    unsigned char* sAl = sAh + M * KT * 2;
    unsigned char* sBh = sAl + M * KT * 2;           // N x KT halves = 28 KB each
    unsigned char* sBl = sBh + N * KT * 2;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    const uint32_t idesc = make_idesc();

    for (int tile = 0; tile < TILES; ++tile) {
        // the previous tile's MMAs must have finished reading shared memory before it is overwritten
        if (tile > 0) { wait_parity(&bar, (tile - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        const float* a = At + (size_t)tile * KT * M;
        const float* b = Bt + (size_t)tile * KT * N;
        for (int i = tid; i < KT * M; i += 128) {
            const int k = i / M, mn = i % M;
            const __half h = __float2half_rn(a[i]);
            *reinterpret_cast<__half*>(sAh + canon_mn_off_bytes(mn, k)) = h;
            *reinterpret_cast<__half*>(sAl + canon_mn_off_bytes(mn, k)) = __float2half_rn(a[i] - __half2float(h));
        }
        for (int i = tid; i < KT * N; i += 128) {
            const int k = i / N, mn = i % N;
            const __half h = __float2half_rn(b[i]);
            *reinterpret_cast<__half*>(sBh + canon_mn_off_bytes(mn, k)) = h;
            *reinterpret_cast<__half*>(sBl + canon_mn_off_bytes(mn, k)) = __float2half_rn(b[i] - __half2float(h));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            for (int ks = 0; ks < KT / 16; ++ks) {
                const uint32_t o = ks * 2 * LBO;
                const uint64_t ah = make_desc(smem_u32(sAh) + o), al = make_desc(smem_u32(sAl) + o);
                const uint64_t bh = make_desc(smem_u32(sBh) + o), bl = make_desc(smem_u32(sBl) + o);
                mma(tbase, al, bh, idesc, (tile > 0 || ks > 0) ? 1u : 0u);
                mma(tbase, ah, bl, idesc, 1u);
                mma(tbase, ah, bh, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
    }
    wait_parity(&bar, (TILES - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = 32 * warp + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tbase + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(TMEM_COLS));
}

int main() {
    std::vector<float> hA((size_t)TILES * KT * M), hB((size_t)TILES * KT * N);
    srand(2);
    for (auto& x : hA) x = (rand() % 20001 - 10000) / 10.0f;          // |x| <= 1000 (scaled-Z-like)
    for (auto& x : hB) x = (rand() % 10001) / 10.0f;                  // 0..1000 (scaled-R-like)
    for (int t = 0; t < TILES; ++t)                                   // rows 64..95 of A^T: one-hot block ids, 96..127: zero
        for (int k = 0; k < KT; ++k)
            for (int m = 64; m < M; ++m) hA[((size_t)t * KT + k) * M + m] = (m < 96 && (k * 7 + t) % 32 == m - 64) ? 1.f : 0.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * M + 2 * N) * KT * 2 + 1024;
    cudaFuncSetAttribute(probe_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_t<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\nFAIL\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(M * N);
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxrel = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0, sa = 0;
            for (int t = 0; t < TILES; ++t)
                for (int k = 0; k < KT; ++k) { const double p = (double)hA[((size_t)t * KT + k) * M + m] * hB[((size_t)t * KT + k) * N + n]; s += p; sa += fabs(p); }
            if (sa > 0) maxrel = fmax(maxrel, fabs(s - hD[m * N + n]) / sa);
        }
    printf("max |err| / sum|terms| = %.3e  (%s; fp16 two-way split should give ~1e-7)\n", maxrel, maxrel < 2e-6 ? "PASS" : "FAIL");
    return maxrel < 2e-6 ? 0 : 1;
}
