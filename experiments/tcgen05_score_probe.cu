// EXPERIMENT (not built into libharmony_b200.so, not on any product path): first-contact probe for the
// tcgen05/TMEM version of the scoring GEMM planned for the next round.
//
//   D[128 x 112] (fp32, TMEM) = A[128 x 64] (fp16, smem) . B[112 x 64]^T (fp16, smem)
//
// with both operands in the canonical K-major NO-SWIZZLE layout (8 x 16-byte core matrices; see the
// SmemDescriptor / InstrDescriptor bit layouts in CUTLASS cute/arch/mma_sm100_desc.hpp), issued as four
// tcgen05.mma.cta_group::1.kind::f16 (K = 16 each) by one thread, completion through
// tcgen05.commit -> mbarrier, read back with tcgen05.ld.32x32b (thread = row) and compared with a CPU
// product.  It exists to validate descriptors and layouts on a B200 before the real kernel is written:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tcgen05_score_probe experiments/tcgen05_score_probe.cu
//   timeout 60 ./tcgen05_score_probe        # prints max |err| and PASS/FAIL
//
// Run on a B200 at the end of round 1: max |err| = 2.161e-06 (PASS).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 112, K = 64;
constexpr int TMEM_COLS = 128;                    // power of two >= N

// element (r, k) of a K-major no-swizzle operand tile with K = 64: core matrix (r/8, k/8), 128 bytes each,
// k-chunks of one row block contiguous: LBO (between k-chunks) = 128 B, SBO (between row blocks) = 8 * 128 B
__host__ __device__ inline int canon_off_bytes(int r, int k) { return ((r >> 3) * (K / 8) + (k >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2; }
constexpr uint32_t LBO = 128, SBO = (K / 8) * 128;

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ inline uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);               // start address  [0,14)
    d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;           // leading byte offset [16,30)
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;           // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                               // version = 1 (Blackwell)
    // base_offset = 0, lbo_mode = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
    return d;
}

__device__ inline uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32           [4,6)
    // a_format [7,10) = 0 (F16), b_format [10,13) = 0 (F16), no negate, a_major = b_major = 0 (K-major)
    d |= (uint32_t)(N >> 3) << 17;      // n_dim                    [17,23)
    d |= (uint32_t)(M >> 4) << 24;      // m_dim                    [24,29)
    return d;
}

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                            // 128 x 64 halves = 16 KB
    unsigned char* sB = smem + M * K * 2;                // 112 x 64 halves = 14 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < M * K; i += 128) { const int r = i / K, k = i % K; *reinterpret_cast<__half*>(sA + canon_off_bytes(r, k)) = A[i]; }
    for (int i = tid; i < N * K; i += 128) { const int r = i / K, k = i % K; *reinterpret_cast<__half*>(sB + canon_off_bytes(r, k)) = B[i]; }
    // generic-proxy writes -> visible to the async proxy that tcgen05.mma reads shared memory through
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

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

    if (tid == 0) {
        const uint32_t idesc = make_idesc();
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO);      // two k-chunks per K = 16 step
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO);
            const uint32_t accumulate = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
        }
        // arrives on the mbarrier when all MMAs above have completed (implies fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for phase 0 of the barrier
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes 32w .. 32w+31; thread = row
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
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { fA[i] = (rand() % 2001 - 1000) / 1000.0f; hA[i] = __float2half(fA[i]); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { fB[i] = (rand() % 2001 - 1000) / 1000.0f; hB[i] = __float2half(fB[i]); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    const size_t smem = (size_t)(M + N) * K * 2 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\nFAIL\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(M * N);
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int r = 0; r < M; ++r)
        for (int c = 0; c < N; ++c) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[r * K + k] * fB[c * K + k];
            maxerr = fmax(maxerr, fabs(s - hD[r * N + c]));
        }
    printf("max |err| = %.3e  (%s)\n", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
    return maxerr < 1e-3 ? 0 : 1;
}
