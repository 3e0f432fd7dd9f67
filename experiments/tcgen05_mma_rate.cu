// EXPERIMENT (not product code): issue + execution rate of tcgen05.mma.kind::f16 (cta_group::1, M = 128, K = 16) for the
// operand layouts and widths the round kernel uses -- no-swizzle core-matrix tiles, K-major vs MN-major A and B,
// N in {16, 32, 48, 64, 112, 128}.  One thread of one CTA per SM issues L MMAs back to back into one accumulator,
// commits, waits; cycles per MMA = (clock after the wait - clock before the first issue) / L.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate experiments/tcgen05_mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

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

__device__ inline bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
struct Cfg { int n, a_mn, b_mn, nacc, ksteps, style, swz; };   // swz: descriptor layout type (0 none, 2 128B, 4 64B, 6 32B)      // style 0: one thread in a divergent branch; 1: converged warp + elect.sync

// A tile: 128 (M) x 128 (K) halves = 32 KB; B tile: 128 (N) x 128 (K) halves = 32 KB.  K-major: LBO 128 / SBO 1024 over
// 64-wide K rows would need several tiles; here every MMA reads K step (i % ksteps) of a [128][16 ksteps] arrangement:
//   K-major operand:  element (r, k): (r >> 3) * SBOK + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2, SBOK = 2048 (16 chunks)
//   MN-major operand: element (r, k): (r >> 3) * 2048 + (k >> 3) * 128 + (k & 7) * 16 + (r & 7) * 2
__global__ void __launch_bounds__(128, 1) rate(Cfg c, int L, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;     // all ones (fp16)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base, sa = smem_u32(smem), sbb = sa + 32768;
    // descriptors of the 8 K steps up front: the issue loop itself is 8 MMAs, nothing else
    uint64_t da[8], db[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        if (c.swz == 0) { da[ks] = make_desc(sa + ks * 256, 128, 2048); db[ks] = make_desc(sbb + ks * 256, 128, 2048); }
        else {
            da[ks] = (c.a_mn ? make_desc(sa + ks * 2048, 16384, 1024) : make_desc(sa + (ks & 3) * 32, 16, 1024)) | ((uint64_t)c.swz << 61);
            db[ks] = (c.b_mn ? make_desc(sbb + ks * 2048, 16384, 1024) : make_desc(sbb + (ks & 3) * 32, 16, 1024)) | ((uint64_t)c.swz << 61);
        }
    }
    const uint32_t idesc = make_idesc(128, c.n, c.a_mn, c.b_mn);
    if (c.style == 1 && warp == 0) {
        long long t0 = clock64();
        for (int i = 0; i < L; i += 8) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                if (elect_one()) mma(tb, da[ks], db[ks], idesc, (i | ks) ? 1u : 0u);
        }
        long long t1 = clock64();
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    if (c.style == 0 && tid == 0) {
        long long t0 = clock64();
        for (int i = 0; i < L; i += 8) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) mma(tb, da[ks], db[ks], idesc, (i | ks) ? 1u : 0u);
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512));
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    long long* d; CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
    const int L = 512;
    const int ns[] = {16, 32, 48, 64, 112, 128};
    printf("cycles per tcgen05.mma (M = 128, K = 16, fp16, no-swizzle tiles), L = %d back to back, %d CTAs:\n", L, prop.multiProcessorCount);
    printf("%5s  %-22s %10s %12s\n", "N", "layout (A, B)", "issue/MMA", "complete/MMA");
    for (int swz = 0; swz <= 2; swz += 2)
    for (int style = 0; style < 2; ++style)
    for (int nacc = 1; nacc <= 1; ++nacc)
        for (int n : ns)
            for (int lay = 0; lay < 4; lay += 3) {
                if (nacc == 2 && n > 64) continue;
                Cfg c{n, lay >> 1, lay & 1, nacc, 8, style, swz};
                rate<<<prop.multiProcessorCount, 128, 65536 + 1024>>>(c, L, d);
                CK(cudaDeviceSynchronize());
                rate<<<prop.multiProcessorCount, 128, 65536 + 1024>>>(c, L, d);
                CK(cudaDeviceSynchronize());
                long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                printf("swz %d style %d %5d  A %-8s B %-8s %s %8.1f %12.1f\n", swz, style, n, c.a_mn ? "MN-major" : "K-major", c.b_mn ? "MN-major" : "K-major",
                       nacc == 1 ? "1 acc " : "2 accs", (double)h[0] / L, (double)h[1] / L);
            }
    return 0;
}
