// EXPERIMENT (not product code): cost of one grid-wide barrier of the persistent round kernel, for the variants
// that could replace grid_barrier() in harmonypy_b200/csrc/hmy_common.cuh.  A round has 21 of them on its critical
// path; the round-1 timeline (profiles/r1_final_timeline_syn1m.txt) shows >= 5-7 us of pure hand-shake per block.
//
//   A  ticket + flag (today):  __threadfence; atomicAdd(count); last arriver resets count, fences, st.release(gen);
//                              the others poll ld.acquire(gen) with __nanosleep(20)
//   B  A without the nanosleep
//   C  cooperative_groups grid.sync()
//   D  A with the fences folded into the atomics (atom.acq_rel.gpu / ld.acquire.gpu), no nanosleep
//   E  one monotonically increasing counter, no reset and no second flag: red.release.gpu.add(count, 1), then poll
//      ld.acquire.gpu(count) until it reaches barrier_index * grid -- the last arrival itself releases everybody
//
// Each variant runs 2000 barriers back to back in grids of 148 x 128 threads (the tc5 kernel's shape) and
// 296 x 256 (today's kernel) and reports microseconds per barrier.  A correctness counter checks that no CTA ever
// ran ahead (every CTA writes its barrier index before arriving and reads all others' after leaving).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -o gbar experiments/grid_barrier_bench.cu && timeout 60 ./gbar
//
// Compile-checked in the build container; NOT yet run on hardware.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ inline unsigned int ld_acquire(const unsigned int* p) { unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ inline void st_release(unsigned int* p, unsigned int v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ inline unsigned int atom_add_acq_rel(unsigned int* p, unsigned int v) { unsigned int o; asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }
__device__ inline void red_add_release(unsigned int* p, unsigned int v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct Bar { unsigned int* count; unsigned int* gen; };

template <int V>
__device__ inline void barrier(const Bar& b, unsigned int idx /* 1-based index of this barrier */, cg::grid_group& grid) {
    if (V == 2) { grid.sync(); return; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int G = gridDim.x;
        if (V == 0 || V == 1) {
            __threadfence();
            const unsigned int t = atomicAdd(b.count, 1u);
            if (t == G - 1u) { *b.count = 0u; __threadfence(); st_release(b.gen, idx); }
            else { while ((int)(ld_acquire(b.gen) - idx) < 0) { if (V == 0) __nanosleep(20); } }
            __threadfence();
        } else if (V == 3) {
            const unsigned int t = atom_add_acq_rel(b.count, 1u);
            if (t == G - 1u) { *b.count = 0u; st_release(b.gen, idx); }
            else { while ((int)(ld_acquire(b.gen) - idx) < 0) { } }
        } else {
            red_add_release(b.count, 1u);
            const unsigned int target = idx * G;
            while ((int)(ld_acquire(b.count) - target) < 0) { }
        }
    }
    __syncthreads();
}

template <int V>
__global__ void bench(Bar b, int iters, unsigned int* progress, unsigned int* errors) {
    cg::grid_group grid = cg::this_grid();
    for (int i = 1; i <= iters; ++i) {
        if (threadIdx.x == 0) progress[blockIdx.x] = (unsigned int)i;      // plain store, published by the barrier
        barrier<V>(b, (unsigned int)i, grid);
        // after barrier i every CTA must have announced i (or already i+1)
        if ((i & 255) == 0) {
            for (unsigned int c = threadIdx.x; c < gridDim.x; c += blockDim.x) {
                const unsigned int p = *(volatile unsigned int*)&progress[c];
                if (p < (unsigned int)i) atomicAdd(errors, 1u);
            }
        }
    }
}

template <int V>
static void run(const char* name, int G, int T, int iters) {
    Bar b; unsigned int *progress, *errors;
    CK(cudaMalloc(&b.count, 8)); b.gen = b.count + 1;
    CK(cudaMalloc(&progress, G * 4)); CK(cudaMalloc(&errors, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    unsigned int herr = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(b.count, 0, 8)); CK(cudaMemset(progress, 0, G * 4)); CK(cudaMemset(errors, 0, 4));
        void* args[] = {&b, &iters, &progress, &errors};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void*)bench<V>, dim3(G), dim3(T), args, 0, 0));
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("%-44s %3d x %3d: CUDA error %s\n", name, G, T, cudaGetErrorString(e)); exit(1); }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
        unsigned int h; CK(cudaMemcpy(&h, errors, 4, cudaMemcpyDeviceToHost)); herr += h;
    }
    printf("%-44s %3d x %3d threads: %6.2f us per barrier   %s\n", name, G, T, 1e3 * best / iters, herr ? "ORDER VIOLATIONS" : "ok");
    cudaFree(b.count); cudaFree(progress); cudaFree(errors);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, iters = 2000;
    const int shapes[2][2] = {{sms, 128}, {2 * sms, 256}};
    for (int s = 0; s < 2; ++s) {
        const int G = shapes[s][0], T = shapes[s][1];
        run<0>("A ticket + flag, nanosleep(20) (today)", G, T, iters);
        run<1>("B ticket + flag, busy poll", G, T, iters);
        run<2>("C cooperative_groups grid.sync()", G, T, iters);
        run<3>("D ticket + flag, acq_rel atomics", G, T, iters);
        run<4>("E single counter, red.release + poll", G, T, iters);
    }
    return 0;
}
