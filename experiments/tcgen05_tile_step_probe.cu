// EXPERIMENT (not product code): third probe for the next round's TMEM pipeline -- one CTA runs the
// complete per-tile step of the clustering round on tensor memory, at the engine's real shapes:
//
//   (1) scoring      D1[128 cells x 112 clusters] = Zs . Ys^T          (K-major A and B, fp16 hi/lo split,
//                                                                        3 MMAs per 16 PCs)
//   (2) epilogue     thread = cell: tcgen05.ld of its 112 scores, r_k = ex2((s-1) c_k) P[combo][k] / sum,
//                    R row -> global (fp32) and -> shared memory as the fp16 hi/lo MN-major A operand of (3)
//   (3) accumulate   D2a[128 clusters x 64 PCs]   += R^T . Zs            (MN-major A = R, MN-major B = the SAME
//                    D2b[128 clusters x 32 levels] += R^T . onehot        shared-memory Z tile the scoring read K-major)
//   (4) final        thread = cluster: tcgen05.ld of its D2a / D2b rows -> global
//
// over TILES tiles (the last one ragged), D2 persistent in TMEM across tiles, mbarrier phase hand-offs between
// the issuing thread and the epilogue threads.  Everything is compared with a double-precision CPU evaluation.
//
// Two things make it self-diagnosing:
//   * `--selfcheck` (CPU only, runs in the build container) checks the index arithmetic against the canonical
//     no-swizzle layouts documented in CUTLASS cute/atom/mma_traits_sm100.hpp (make_umma_desc): the K-major and
//     MN-major views of the ONE shared-memory Z tile address identical bytes, every layout is a bijection onto
//     its buffer, a K = 16 step is a start-address shift of two k-blocks, the epilogue's 16-byte stores cover 8
//     consecutive clusters, and all descriptor fields fit;
//   * on the GPU it runs twice: MN-major descriptors as CUTLASS documents them (LBO = stride between 8-wide
//     k-blocks, SBO = stride between 8-wide mn-blocks) and with the two swapped, and prints which one passes.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_s experiments/tcgen05_tile_step_probe.cu
//   ./probe_s --selfcheck            # CPU
//   timeout 60 ./probe_s             # B200
//
// Run on a B200 at the very end of round 1 (profiles/r1_tcgen05_probes_b200.txt):
//   documented convention: R max|err| = 8.056e-07, Yacc rel = 2.220e-06, Oacc rel = 3.466e-06 -> PASS
//   LBO/SBO exchanged:     R ok (scoring is K-major), Yacc / Oacc wrong                     -> as expected
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int TILE = 128;            // cells per tile = TMEM lanes of D1
constexpr int DP = 64;               // PCs, padded
constexpr int KP = 112;              // clusters, padded to a multiple of 16 (N of the scoring MMA)
constexpr int KM = 128;              // clusters as M of the accumulation MMA (rows 112..127 stay zero)
constexpr int NB = 32;               // one-hot level columns, padded
constexpr int TILES = 3;
constexpr int NCOMBO = 4;
constexpr int TMEM_COLS = 256;       // D1: cols [0,112)   D2a: [128,192)   D2b: [192,224)
constexpr int COL_D2A = 128, COL_D2B = 192;
constexpr float SCALE = 1024.f;      // operands are scaled by 2^10 before the fp16 hi/lo split

// ---- canonical no-swizzle layouts (CUTLASS make_umma_desc, SWIZZLE_NONE / "INTERLEAVE") -----------------------
// 8 x 16-byte core matrices.  K-major: a core matrix holds 8 mn-rows x 8 k;  MN-major: 8 k-rows x 8 mn.
// In both, LBO is the byte stride between core matrices along k and SBO the one along mn.
__host__ __device__ inline int off_kmajor(int mn, int k, int lbo, int sbo) { return (mn >> 3) * sbo + (k >> 3) * lbo + (mn & 7) * 16 + (k & 7) * 2; }
__host__ __device__ inline int off_mnmajor(int mn, int k, int lbo, int sbo) { return (mn >> 3) * sbo + (k >> 3) * lbo + (k & 7) * 16 + (mn & 7) * 2; }

// Z tile: staged once as the K-major A of the scoring (mn = cell, k = PC) ...
constexpr int Z_LBO_K = 128, Z_SBO_K = (DP / 8) * 128;            // = 1024
// ... and read again as the MN-major B of the accumulation (mn = PC, k = cell): the same bytes with
constexpr int Z_LBO_MN = Z_SBO_K, Z_SBO_MN = Z_LBO_K;
// centroids: K-major B of the scoring (mn = cluster, k = PC)
constexpr int Y_LBO = 128, Y_SBO = (DP / 8) * 128;
// R tile: MN-major A of the accumulation (mn = cluster, k = cell); cluster blocks outermost so that one
// warp's 16-byte stores (thread = cell) land on 512 contiguous bytes
constexpr int R_LBO = 128, R_SBO = (TILE / 8) * 128;              // = 2048
// one-hot tile: MN-major B (mn = level, k = cell), same arrangement
constexpr int O_LBO = 128, O_SBO = (TILE / 8) * 128;

constexpr int SZ_BYTES = TILE * DP * 2, SY_BYTES = KP * DP * 2, SR_BYTES = KM * TILE * 2, SO_BYTES = NB * TILE * 2;

// ---- device helpers ------------------------------------------------------------------------------------------
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);               // start address        [0,14)
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;           // leading byte offset  [16,30)
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;           // stride byte offset   [32,46)
    d |= (uint64_t)1 << 46;                               // version = 1; layout_type [61,64) = 0: no swizzle
    return d;
}
__device__ inline uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                                         // c_format = F32; a/b_format = 0 (F16)
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}
__device__ inline void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ inline void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ inline void wait_parity(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ inline void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
}
// generic-proxy shared-memory writes -> async proxy (tcgen05.mma), then the CTA-wide hand-off to the issuer
__device__ inline void publish_smem_and_sync() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ inline float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ inline void split_store(unsigned char* hi, unsigned char* lo, int off, float x) {
    const __half h = __float2half_rn(x);
    *reinterpret_cast<__half*>(hi + off) = h;
    *reinterpret_cast<__half*>(lo + off) = __float2half_rn(x - __half2float(h));
}

struct Args {
    const float* Z;        // [TILES*TILE][DP]   unit rows (zero rows for the ragged tail)
    const float* Y;        // [KP][DP]           unit rows, rows >= K are zero
    const float* c;        // [KP]               2 log2(e) / sigma_k
    const float* P;        // [NCOMBO][KP]       penalty product, 0 for padded clusters
    const int* combo;      // [TILES*TILE]
    const int* lev;        // [NCOMBO][2]        the two one-hot columns of a combination
    int n_valid;           // cells < n_valid are real
    int swap_mn;           // 0: MN-major descriptors as documented; 1: LBO and SBO exchanged
    float* R;              // [TILES*TILE][KP]
    float* Yacc;           // [KM][DP]
    float* Oacc;           // [KM][NB]
};

__global__ void __launch_bounds__(128) tile_step(Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sZh = smem;
    unsigned char* sZl = sZh + SZ_BYTES;
    unsigned char* sYh = sZl + SZ_BYTES;
    unsigned char* sYl = sYh + SY_BYTES;
    unsigned char* sRh = sYl + SY_BYTES;
    unsigned char* sRl = sRh + SR_BYTES;
    unsigned char* sO  = sRl + SR_BYTES;
    float* sP = reinterpret_cast<float*>(sO + SO_BYTES);          // [NCOMBO][KP]
    float* sc = sP + NCOMBO * KP;                                 // [KP]
    __shared__ __align__(8) uint64_t bar_score, bar_acc;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_score)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_acc)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // round-constant operands: centroids (scaled, split), penalty table, exponents; R pad rows = 0
    for (int i = tid; i < KP * DP; i += 128) { const int k = i / DP, p = i % DP; split_store(sYh, sYl, off_kmajor(k, p, Y_LBO, Y_SBO), a.Y[i] * SCALE); }
    for (int i = tid; i < NCOMBO * KP; i += 128) sP[i] = a.P[i];
    for (int i = tid; i < KP; i += 128) sc[i] = a.c[i];
    for (int i = tid; i < SR_BYTES / 4; i += 128) { reinterpret_cast<uint32_t*>(sRh)[i] = 0; reinterpret_cast<uint32_t*>(sRl)[i] = 0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    const uint32_t id_score = make_idesc(TILE, KP, 0, 0);
    const uint32_t id_acc_y = make_idesc(KM, DP, 1, 1);
    const uint32_t id_acc_o = make_idesc(KM, NB, 1, 1);
    // MN-major (LBO, SBO) pairs, optionally exchanged to let the hardware arbitrate the documentation
    const uint32_t r_lbo = a.swap_mn ? R_SBO : R_LBO, r_sbo = a.swap_mn ? R_LBO : R_SBO;
    const uint32_t z_lbo = a.swap_mn ? Z_SBO_MN : Z_LBO_MN, z_sbo = a.swap_mn ? Z_LBO_MN : Z_SBO_MN;
    const uint32_t o_lbo = a.swap_mn ? O_SBO : O_LBO, o_sbo = a.swap_mn ? O_LBO : O_SBO;

    for (int tile = 0; tile < TILES; ++tile) {
        const int cell = tile * TILE + tid;                       // thread = cell from here on
        const bool valid = cell < a.n_valid;
        // the accumulation MMAs of the previous tile still read sZ / sR / sO
        if (tile > 0) wait_parity(&bar_acc, (tile - 1) & 1);

        // ---- stage the Z tile (coalesced global reads, thread i handles element i of the tile) and the one-hot tile
        for (int i = tid; i < TILE * DP; i += 128) {
            const int r = i / DP, p = i % DP;
            split_store(sZh, sZl, off_kmajor(r, p, Z_LBO_K, Z_SBO_K), a.Z[(size_t)tile * TILE * DP + i] * SCALE);
        }
        const int cb = valid ? a.combo[cell] : 0;
        {
            const int l0 = a.lev[cb * 2], l1 = a.lev[cb * 2 + 1];
#pragma unroll
            for (int j = 0; j < NB / 8; ++j) {
                __align__(16) __half h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = __float2half_rn((valid && (8 * j + e == l0 || 8 * j + e == l1)) ? 1.f : 0.f);
                *reinterpret_cast<uint4*>(sO + off_mnmajor(8 * j, tid, O_LBO, O_SBO)) = *reinterpret_cast<const uint4*>(h);
            }
        }
        publish_smem_and_sync();

        // ---- (1) scoring: 4 k-steps x 3 split products
        if (tid == 0) {
#pragma unroll
            for (int ks = 0; ks < DP / 16; ++ks) {
                const uint32_t o = ks * 2 * Z_LBO_K;              // two 8-wide PC chunks per K = 16
                const uint64_t zh = make_desc(smem_u32(sZh) + o, Z_LBO_K, Z_SBO_K), zl = make_desc(smem_u32(sZl) + o, Z_LBO_K, Z_SBO_K);
                const uint64_t yh = make_desc(smem_u32(sYh) + o, Y_LBO, Y_SBO), yl = make_desc(smem_u32(sYl) + o, Y_LBO, Y_SBO);
                mma(tbase, zl, yh, id_score, ks > 0 ? 1u : 0u);
                mma(tbase, zh, yl, id_score, 1u);
                mma(tbase, zh, yh, id_score, 1u);
            }
            commit(&bar_score);
        }
        wait_parity(&bar_score, tile & 1);

        // ---- (2) epilogue, thread = cell
        {
            float e[KP];
            float sum = 0.f;
            const float* Pc = sP + cb * KP;
#pragma unroll
            for (int c0 = 0; c0 < KP; c0 += 16) {
                float s[16];
                tmem_ld16(tbase + lane_base + (uint32_t)c0, s);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float sim = s[j] * (1.f / (SCALE * SCALE));
                    const float v = ex2_approx((sim - 1.f) * sc[c0 + j]) * Pc[c0 + j];
                    e[c0 + j] = v;
                    sum += v;
                }
            }
            const float inv = valid ? 1.f / sum : 0.f;
            float* Rrow = a.R + (size_t)cell * KP;
#pragma unroll
            for (int c0 = 0; c0 < KP; c0 += 8) {
                __align__(16) __half hi[8], lo[8];
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    r[j] = e[c0 + j] * inv;
                    const float x = r[j] * SCALE;
                    hi[j] = __float2half_rn(x);
                    lo[j] = __float2half_rn(x - __half2float(hi[j]));
                }
                *reinterpret_cast<float4*>(Rrow + c0) = make_float4(r[0], r[1], r[2], r[3]);
                *reinterpret_cast<float4*>(Rrow + c0 + 4) = make_float4(r[4], r[5], r[6], r[7]);
                const int off = off_mnmajor(c0, tid, R_LBO, R_SBO);           // 8 clusters = one 16-byte core row
                *reinterpret_cast<uint4*>(sRh + off) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(sRl + off) = *reinterpret_cast<const uint4*>(lo);
            }
        }
        publish_smem_and_sync();                                   // also orders the tcgen05.ld of D1 before the next scoring

        // ---- (3) accumulation: 8 k-steps (16 cells each)
        if (tid == 0) {
#pragma unroll
            for (int ks = 0; ks < TILE / 16; ++ks) {
                const uint32_t first = (tile > 0 || ks > 0) ? 1u : 0u;
                const uint32_t ro = ks * 2 * R_LBO, zo = ks * 2 * Z_LBO_MN, oo = ks * 2 * O_LBO;   // two 8-cell blocks per K = 16
                const uint64_t rh = make_desc(smem_u32(sRh) + ro, r_lbo, r_sbo), rl = make_desc(smem_u32(sRl) + ro, r_lbo, r_sbo);
                const uint64_t zh = make_desc(smem_u32(sZh) + zo, z_lbo, z_sbo), zl = make_desc(smem_u32(sZl) + zo, z_lbo, z_sbo);
                const uint64_t oh = make_desc(smem_u32(sO) + oo, o_lbo, o_sbo);
                mma(tbase + COL_D2A, rl, zh, id_acc_y, first);
                mma(tbase + COL_D2A, rh, zl, id_acc_y, 1u);
                mma(tbase + COL_D2A, rh, zh, id_acc_y, 1u);
                mma(tbase + COL_D2B, rl, oh, id_acc_o, first);
                mma(tbase + COL_D2B, rh, oh, id_acc_o, 1u);
            }
            commit(&bar_acc);
        }
    }
    wait_parity(&bar_acc, (TILES - 1) & 1);

    // ---- (4) final read-out, thread = cluster
    {
        float v[16];
        for (int c0 = 0; c0 < DP; c0 += 16) {
            tmem_ld16(tbase + lane_base + (uint32_t)(COL_D2A + c0), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) a.Yacc[tid * DP + c0 + j] = v[j] * (1.f / (SCALE * SCALE));
        }
        for (int c0 = 0; c0 < NB; c0 += 16) {
            tmem_ld16(tbase + lane_base + (uint32_t)(COL_D2B + c0), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) a.Oacc[tid * NB + c0 + j] = v[j] * (1.f / SCALE);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(TMEM_COLS));
}

// ---- host ------------------------------------------------------------------------------------------------------
struct Problem {
    int K = 100, d = 50, n_valid = TILES * TILE - 28;
    std::vector<float> Z, Y, c, P;
    std::vector<int> combo, lev;
};

static Problem make_problem() {
    Problem p;
    const int n = TILES * TILE;
    p.Z.assign((size_t)n * DP, 0.f); p.Y.assign((size_t)KP * DP, 0.f); p.c.assign(KP, 0.f); p.P.assign((size_t)NCOMBO * KP, 0.f);
    p.combo.assign(n, 0); p.lev = {0, 3, 1, 4, 2, 7, 1, 5};                   // covariate 1: levels 0..2, covariate 2: columns 3..7
    srand(3);
    auto unit_row = [&](float* row) {
        double ss = 0;
        for (int j = 0; j < p.d; ++j) { row[j] = (rand() % 2001 - 1000) / 1000.0f; ss += (double)row[j] * row[j]; }
        for (int j = 0; j < p.d; ++j) row[j] = (float)(row[j] / sqrt(ss));
    };
    for (int k = 0; k < p.K; ++k) unit_row(&p.Y[(size_t)k * DP]);
    for (int i = 0; i < p.n_valid; ++i) {
        // cells near a centroid so that R is not uniform
        float* z = &p.Z[(size_t)i * DP];
        unit_row(z);
        const float* y = &p.Y[(size_t)(rand() % p.K) * DP];
        double ss = 0;
        for (int j = 0; j < p.d; ++j) { z[j] = 0.6f * y[j] + 0.4f * z[j]; ss += (double)z[j] * z[j]; }
        for (int j = 0; j < p.d; ++j) z[j] = (float)(z[j] / sqrt(ss));
        p.combo[i] = (i / 37) % NCOMBO;
    }
    for (int k = 0; k < p.K; ++k) p.c[k] = (float)(2.0 * 1.4426950408889634 / 0.1);
    for (int cb = 0; cb < NCOMBO; ++cb)
        for (int k = 0; k < p.K; ++k) p.P[(size_t)cb * KP + k] = 0.5f + (rand() % 1000) / 1000.0f;
    return p;
}

// CPU restatement of what the kernel stages, viewed through the canonical layouts
static int selfcheck(const Problem& p) {
    int bad = 0;
    // (a) Z tile written K-major (cell, pc), read MN-major (pc, cell): identical byte for every element
    for (int cell = 0; cell < TILE; ++cell)
        for (int pc = 0; pc < DP; ++pc)
            if (off_kmajor(cell, pc, Z_LBO_K, Z_SBO_K) != off_mnmajor(pc, cell, Z_LBO_MN, Z_SBO_MN)) ++bad;
    // (b) every layout is a bijection onto its buffer (no two elements collide, nothing outside)
    auto bijective = [&](int mn_n, int k_n, int lbo, int sbo, bool mnmajor, int bytes) {
        std::vector<char> seen(bytes / 2, 0);
        for (int mn = 0; mn < mn_n; ++mn)
            for (int k = 0; k < k_n; ++k) {
                const int o = mnmajor ? off_mnmajor(mn, k, lbo, sbo) : off_kmajor(mn, k, lbo, sbo);
                if (o < 0 || o + 2 > bytes || (o & 1) || seen[o / 2]) return 1;
                seen[o / 2] = 1;
            }
        return 0;
    };
    bad += bijective(TILE, DP, Z_LBO_K, Z_SBO_K, false, SZ_BYTES);
    bad += bijective(KP, DP, Y_LBO, Y_SBO, false, SY_BYTES);
    bad += bijective(KM, TILE, R_LBO, R_SBO, true, SR_BYTES);
    bad += bijective(NB, TILE, O_LBO, O_SBO, true, SO_BYTES);
    // (c) a K = 16 step advances the start address by two k-blocks: element (mn, 16 ks + kk) through the shifted
    //     descriptor equals element (mn, kk) + ks * 2 * LBO
    for (int ks = 0; ks < TILE / 16; ++ks)
        for (int mn = 0; mn < KM; mn += 13)
            for (int kk = 0; kk < 16; ++kk)
                if (off_mnmajor(mn, 16 * ks + kk, R_LBO, R_SBO) != ks * 2 * R_LBO + off_mnmajor(mn, kk, R_LBO, R_SBO)) ++bad;
    for (int ks = 0; ks < TILE / 16; ++ks)
        for (int pc = 0; pc < DP; pc += 5)
            for (int kk = 0; kk < 16; ++kk)
                if (off_mnmajor(pc, 16 * ks + kk, Z_LBO_MN, Z_SBO_MN) != ks * 2 * Z_LBO_MN + off_mnmajor(pc, kk, Z_LBO_MN, Z_SBO_MN)) ++bad;
    // (d) the epilogue's 16-byte R store covers clusters c0..c0+7 of its cell
    for (int cell = 0; cell < TILE; cell += 9)
        for (int c0 = 0; c0 < KP; c0 += 8)
            for (int j = 0; j < 8; ++j)
                if (off_mnmajor(c0 + j, cell, R_LBO, R_SBO) != off_mnmajor(c0, cell, R_LBO, R_SBO) + 2 * j) ++bad;
    // (e) descriptor fields fit: 14-bit fields in 16-byte units
    const int fields[] = {Z_LBO_K, Z_SBO_K, Z_LBO_MN, Z_SBO_MN, Y_LBO, Y_SBO, R_LBO, R_SBO, O_LBO, O_SBO};
    for (int f : fields) if ((f & 15) || (f >> 4) > 0x3FFF) ++bad;
    (void)p;
    printf("selfcheck: %s (%d problems)\n", bad ? "FAIL" : "PASS", bad);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    Problem p = make_problem();
    if (argc > 1 && !strcmp(argv[1], "--selfcheck")) return selfcheck(p);
    const int n = TILES * TILE;

    // double-precision evaluation of the same step
    std::vector<double> Rref((size_t)n * KP, 0.0), Yref((size_t)KM * DP, 0.0), Oref((size_t)KM * NB, 0.0), Yabs((size_t)KM * DP, 0.0);
    for (int i = 0; i < p.n_valid; ++i) {
        double sum = 0; std::vector<double> e(KP, 0.0);
        for (int k = 0; k < p.K; ++k) {
            double s = 0;
            for (int j = 0; j < DP; ++j) s += (double)p.Z[(size_t)i * DP + j] * p.Y[(size_t)k * DP + j];
            e[k] = exp2((s - 1.0) * p.c[k]) * p.P[(size_t)p.combo[i] * KP + k];
            sum += e[k];
        }
        for (int k = 0; k < p.K; ++k) {
            const double r = e[k] / sum;
            Rref[(size_t)i * KP + k] = r;
            for (int j = 0; j < DP; ++j) { Yref[(size_t)k * DP + j] += r * p.Z[(size_t)i * DP + j]; Yabs[(size_t)k * DP + j] += fabs(r * p.Z[(size_t)i * DP + j]); }
            Oref[(size_t)k * NB + p.lev[p.combo[i] * 2]] += r;
            Oref[(size_t)k * NB + p.lev[p.combo[i] * 2 + 1]] += r;
        }
    }

    Args a{};
    float *dZ, *dY, *dc, *dP, *dR, *dYacc, *dOacc; int *dcombo, *dlev;
    cudaMalloc(&dZ, p.Z.size() * 4); cudaMalloc(&dY, p.Y.size() * 4); cudaMalloc(&dc, KP * 4); cudaMalloc(&dP, p.P.size() * 4);
    cudaMalloc(&dR, (size_t)n * KP * 4); cudaMalloc(&dYacc, KM * DP * 4); cudaMalloc(&dOacc, KM * NB * 4);
    cudaMalloc(&dcombo, n * 4); cudaMalloc(&dlev, p.lev.size() * 4);
    cudaMemcpy(dZ, p.Z.data(), p.Z.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dY, p.Y.data(), p.Y.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, p.c.data(), KP * 4, cudaMemcpyHostToDevice); cudaMemcpy(dP, p.P.data(), p.P.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dcombo, p.combo.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dlev, p.lev.data(), p.lev.size() * 4, cudaMemcpyHostToDevice);
    a.Z = dZ; a.Y = dY; a.c = dc; a.P = dP; a.combo = dcombo; a.lev = dlev; a.n_valid = p.n_valid; a.R = dR; a.Yacc = dYacc; a.Oacc = dOacc;
    const size_t smem = 2 * SZ_BYTES + 2 * SY_BYTES + 2 * SR_BYTES + SO_BYTES + (NCOMBO * KP + KP) * 4 + 1024;
    cudaFuncSetAttribute(tile_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

    int pass_any = 0;
    for (int swap = 0; swap < 2; ++swap) {
        a.swap_mn = swap;
        cudaMemset(dR, 0xff, (size_t)n * KP * 4); cudaMemset(dYacc, 0xff, KM * DP * 4); cudaMemset(dOacc, 0xff, KM * NB * 4);
        tile_step<<<1, 128, smem>>>(a);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("swap_mn=%d: CUDA error: %s\n", swap, cudaGetErrorString(err)); return 1; }
        std::vector<float> R((size_t)n * KP), Ya(KM * DP), Oa(KM * NB);
        cudaMemcpy(R.data(), dR, R.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(Ya.data(), dYacc, Ya.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(Oa.data(), dOacc, Oa.size() * 4, cudaMemcpyDeviceToHost);
        double eR = 0, eY = 0, eO = 0;
        for (size_t i = 0; i < R.size(); ++i) eR = fmax(eR, fabs(R[i] - Rref[i]));
        for (int k = 0; k < KM; ++k)
            for (int j = 0; j < DP; ++j) eY = fmax(eY, fabs(Ya[k * DP + j] - Yref[(size_t)k * DP + j]) / fmax(Yabs[(size_t)k * DP + j], 1e-3));
        for (int k = 0; k < KM; ++k)
            for (int j = 0; j < NB; ++j) eO = fmax(eO, fabs(Oa[k * NB + j] - Oref[(size_t)k * NB + j]) / fmax(Oref[(size_t)k * NB + j], 1e-3));
        const bool okR = eR < 2e-5, okY = eY < 1e-4, okO = eO < 1e-4;
        printf("MN-major descriptors %s:  R max|err| = %.3e (%s)   Yacc rel = %.3e (%s)   Oacc rel = %.3e (%s)\n",
               swap ? "with LBO/SBO exchanged" : "as documented (LBO: k-blocks, SBO: mn-blocks)",
               eR, okR ? "ok" : "BAD", eY, okY ? "ok" : "BAD", eO, okO ? "ok" : "BAD");
        if (okR && okY && okO) pass_any |= 1 << swap;
    }
    printf("%s\n", pass_any == 1 ? "PASS (documented convention)" : pass_any ? "PASS only with exchanged LBO/SBO -- fix the constants before building on this" : "FAIL");
    return pass_any == 1 ? 0 : 1;
}
