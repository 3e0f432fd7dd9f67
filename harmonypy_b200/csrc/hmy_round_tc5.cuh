// tcgen05 / tensor-memory version of the k-means round (opt-in: engine option "tc5").
//
// STATUS: written at the end of round 1 after the GPU budget was used up.  Its inner loop -- the tile step
// (1)-(3) below with exactly these layouts, descriptors and fences -- ran stand-alone on a B200
// (experiments/tcgen05_tile_step_probe.cu: R 8e-7, sums 2-3e-6 against fp64), but THIS kernel, with the
// block lists, tables and barriers around it, has not run yet: it is NOT selected by default and the parity
// tests only run it when HMY_TEST_TC5=1.
//
// Same algorithm and the same grid-level structure as k_round_mma (hmy_round_mma.cuh): phase 0, the
// per-CTA K x B tables, the block lists, the barriers and the multi-GPU communication CTA are reused
// unchanged.  Only the per-block cell processing differs:
//
//   tile = 128 cells = the 128 lanes of tensor memory; CTA = 128 threads; THREAD = CELL.
//   (1) scoring     D1[cell][cluster]  = Zs . Ys^T      tcgen05.mma kind::f16, M = 128, N = 16 NC, fp16
//                   hi/lo split (3 MMAs per 16 PCs), operands in shared memory in the canonical no-swizzle
//                   K-major layout, one thread issues, completion through tcgen05.commit -> mbarrier
//   (2) epilogue    every thread reads ITS cell's scores with tcgen05.ld.32x32b: the softmax sums, the
//                   penalty product, the objective terms are thread-local (no shuffles, no fragment
//                   bookkeeping); the new R row goes to HBM and, as fp16 hi/lo, to shared memory in the
//                   MN-major layout (harmony.py:466-468, :495-503)
//   (3) sums        D2y[cluster][PC]    += R^T . Zs     (harmony.py:443; the Z tile of (1) read MN-major)
//                   D2o[cluster][level] += R^T . onehot (harmony.py:506-507; multi-hot level rows)
//                   accumulated IN TENSOR MEMORY: D2o over a block (then thread = cluster adds its row to
//                   Dnew / Ofresh), D2y over the whole round (then thread = cluster adds its row to Yacc).
//
// Limits: K <= 128, d <= 64, B <= 32, nblk <= 32; everything else stays on k_round_mma.
#pragma once
#include "hmy_round_mma.cuh"

#define TC5_TILE 128          // cells per tile
#define TC5_DP 64             // PCs (padded)
#define TC5_NB 32             // one-hot columns (padded)
#define TC5_KM 128            // clusters as the M of the accumulation
#define TC5_TMEM_COLS 256     // D1: [0, 16 NC)   D2y: [128, 192)   D2o: [192, 224)
#define TC5_COL_Y 128
#define TC5_COL_O 192

// canonical no-swizzle operand layouts (8 x 16-byte core matrices; CUTLASS make_umma_desc):
// LBO = byte stride between core matrices along k, SBO = along m/n
#define TC5_Z_LBO_K 128       // Z tile as K-major A of the scoring: (cell, PC)
#define TC5_Z_SBO_K 1024
#define TC5_Z_LBO_MN 1024     // the same bytes as MN-major B of the accumulation: (PC, cell)
#define TC5_Z_SBO_MN 128
#define TC5_Y_LBO 128         // centroids, K-major B of the scoring: (cluster, PC)
#define TC5_Y_SBO 1024
#define TC5_R_LBO 128         // R tile, MN-major A of the accumulation: (cluster, cell), cluster blocks outermost
#define TC5_R_SBO 2048
#define TC5_O_LBO 128         // one-hot tile, MN-major B: (level, cell)
#define TC5_O_SBO 2048

#define TC5_OFF_ZH 0
#define TC5_OFF_ZL (TC5_OFF_ZH + TC5_TILE * TC5_DP * 2)
#define TC5_OFF_RH (TC5_OFF_ZL + TC5_TILE * TC5_DP * 2)
#define TC5_OFF_RL (TC5_OFF_RH + TC5_KM * TC5_TILE * 2)
#define TC5_OFF_OT (TC5_OFF_RL + TC5_KM * TC5_TILE * 2)
#define TC5_OFF_YH (TC5_OFF_OT + TC5_NB * TC5_TILE * 2)
__host__ __device__ constexpr int tc5_off_yl(int NC) { return TC5_OFF_YH + 16 * NC * TC5_DP * 2; }
__host__ __device__ constexpr int tc5_off_c1(int NC) { return tc5_off_yl(NC) + 16 * NC * TC5_DP * 2; }
__host__ __device__ constexpr int tc5_off_c3(int NC) { return tc5_off_c1(NC) + 16 * NC * 4; }
__host__ __device__ constexpr int tc5_off_ps(int NC) { return tc5_off_c3(NC) + 16 * NC * 4; }

struct Tc5Smem {
    int KT2, RSH;
    int off_Zh, off_Zl, off_Rh, off_Rl, off_Ot, off_Yh, off_Yl, off_c1, off_c3, off_Ps, off_Os, off_prb, off_cell, off_misc;
    int total;
};

__host__ __device__ inline Tc5Smem tc5_smem_plan(int B, int NC) {
    Tc5Smem s;
    s.KT2 = 16 * NC;
    s.RSH = hmy_odd8(s.KT2);                       // row stride of the phase-0 R tile (aliases the R tiles below)
    // everything up to the penalty table sits at an offset that only depends on NC: the kernel addresses it
    // relative to ONE base register (TC5_OFF_* below)
    int o = 0;
    s.off_Zh = TC5_OFF_ZH; s.off_Zl = TC5_OFF_ZL; s.off_Rh = TC5_OFF_RH; s.off_Rl = TC5_OFF_RL; s.off_Ot = TC5_OFF_OT;
    s.off_Yh = TC5_OFF_YH;
    s.off_Yl = tc5_off_yl(NC); s.off_c1 = tc5_off_c1(NC); s.off_c3 = tc5_off_c3(NC); s.off_Ps = tc5_off_ps(NC);
    o = s.off_Ps + B * s.KT2 * 4;
    s.off_Os = o; o += B * s.KT2 * 4;
    s.off_prb = o; o += 2 * B * 4;
    o = (o + 15) & ~15;
    s.off_cell = o; o += TC5_TILE * 4;
    o = (o + 15) & ~15;
    s.off_misc = o; o += 8 * 256 + 128;
    s.total = o;
    return s;
}

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long tc5_desc(unsigned int saddr, unsigned int lbo, unsigned int sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFFu);              // start address        [0,14)
    d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;          // leading byte offset  [16,30)
    d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;          // stride byte offset   [32,46)
    d |= 1ull << 46;                                                // version 1; layout_type 0 = no swizzle
    return d;
}
__device__ __forceinline__ unsigned int tc5_idesc(int m, int n, int a_mn_major, int b_mn_major) {
    unsigned int d = 0;
    d |= 1u << 4;                                                   // D = f32; A, B = f16
    d |= (unsigned int)a_mn_major << 15;
    d |= (unsigned int)b_mn_major << 16;
    d |= (unsigned int)(n >> 3) << 17;
    d |= (unsigned int)(m >> 4) << 24;
    return d;
}
__device__ __forceinline__ void tc5_mma(unsigned int tmem, unsigned long long da, unsigned long long db, unsigned int idesc, unsigned int accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc5_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// An MMA that never completes (bad descriptor, lost commit) must become an error, not a hung cooperative
// grid: after ~2 s of waiting the kernel traps and the host sees a launch failure.
__device__ __forceinline__ void tc5_wait(unsigned long long* bar, unsigned int parity) {
    unsigned int done = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && (++spins & 1023u) == 0u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc5_ld16(unsigned int taddr, float (&v)[16]) {
    unsigned int u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
}
// generic-proxy shared-memory writes -> async proxy (the MMA reads shared memory through it), the CTA's
// tensor-memory loads -> before the next MMA, then the hand-off to the issuing thread
__device__ __forceinline__ void tc5_publish() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- per-CTA context ---------------------------------------------------------------------------------
template <int NC>
struct Tc5Ctx {
    unsigned char* base;                 // dynamic shared memory; tiles / centroids / constants at TC5_OFF_*
    int* sCell;
    unsigned long long* bar;             // [0] scoring done, [1] accumulation done
    unsigned int tmem, lane_base;
    unsigned int ph_score, ph_acc;       // phases consumed so far (parity = & 1); identical in all threads
    bool acc_pending;                    // an accumulation is in flight: it still reads the Z / R / one-hot tiles
    bool y_started, o_started;           // D2y / D2o hold sums (first MMA into an empty accumulator overwrites)
    int dt;                              // k-steps over the PCs
    unsigned int mg_dp4;
    __device__ __forceinline__ unsigned char* Zh() const { return base + TC5_OFF_ZH; }
    __device__ __forceinline__ unsigned char* Zl() const { return base + TC5_OFF_ZL; }
    __device__ __forceinline__ unsigned char* Rh() const { return base + TC5_OFF_RH; }
    __device__ __forceinline__ unsigned char* Rl() const { return base + TC5_OFF_RL; }
    __device__ __forceinline__ unsigned char* Ot() const { return base + TC5_OFF_OT; }
    __device__ __forceinline__ unsigned char* Yh() const { return base + TC5_OFF_YH; }
    __device__ __forceinline__ unsigned char* Yl() const { return base + tc5_off_yl(NC); }
    __device__ __forceinline__ float* c1() const { return reinterpret_cast<float*>(base + tc5_off_c1(NC)); }
    __device__ __forceinline__ float* c3() const { return reinterpret_cast<float*>(base + tc5_off_c3(NC)); }
    __device__ __forceinline__ float* Ps() const { return reinterpret_cast<float*>(base + tc5_off_ps(NC)); }
    // the staged tile, thread = cell
    int cell; bool valid;
    int lev[HMY_MAX_V];
    double objd, obje;
};

template <int NC>
__device__ __forceinline__ void tc5_wait_acc(Tc5Ctx<NC>& c) {
    if (c.acc_pending) { tc5_wait(&c.bar[1], c.ph_acc & 1u); c.ph_acc++; c.acc_pending = false; }
}

// centroids as fp16 hi/lo in the canonical K-major layout, per-cluster constants, empty tables
template <int NC>
__device__ void tc5_load_round_constants(Tc5Ctx<NC>& c, const HmyDev& st, float* Os, float* sPrb, float* sTheta) {
    constexpr int KT2 = 16 * NC;
    const int tid = threadIdx.x;
    for (int i = tid; i < KT2 * (TC5_DP / 2); i += TC5_TILE) {
        const int k = i / (TC5_DP / 2), j = 2 * (i - k * (TC5_DP / 2));
        float y0 = 0.f, y1 = 0.f;
        if (k < st.K) {
            if (j < st.dp) y0 = st.Yhat[(size_t)k * st.dp + j] * HMY_OPSCALE;
            if (j + 1 < st.dp) y1 = st.Yhat[(size_t)k * st.dp + j + 1] * HMY_OPSCALE;
        }
        unsigned int hi, lo;
        split2(y0, y1, hi, lo);
        const int off = (k >> 3) * TC5_Y_SBO + (j >> 3) * TC5_Y_LBO + (k & 7) * 16 + (j & 7) * 2;
        *reinterpret_cast<unsigned int*>(c.Yh() + off) = hi;
        *reinterpret_cast<unsigned int*>(c.Yl() + off) = lo;
    }
    for (int i = tid; i < st.B * KT2; i += TC5_TILE) { c.Ps()[i] = 0.f; Os[i] = 0.f; }
    for (int b = tid; b < st.B; b += TC5_TILE) { sPrb[b] = st.Pr_b[b]; sTheta[b] = st.theta[b]; }
    // t = c2 - acc * c1 is (dist / sigma) * log2(e); c2 = c1 * 2^20; dist = t * c3; sigma ln r = c3 lg2 r
    for (int k = tid; k < KT2; k += TC5_TILE) {
        const float sg = (k < st.K) ? st.sigma[k] : 1.f;
        c.c1()[k] = (k < st.K) ? (2.0f * 1.4426950408889634f / sg) * HMY_ACCSCALE : 0.f;
        c.c3()[k] = (k < st.K) ? sg * 0.6931471805599453f : 0.f;
    }
    // Z tiles: the PCs beyond dp are never written again and must read as zero; R tiles: phase 0 aliases them
    uint4* z = reinterpret_cast<uint4*>(c.Zh());
    for (int i = tid; i < 2 * TC5_TILE * TC5_DP * 2 / 16; i += TC5_TILE) z[i] = make_uint4(0u, 0u, 0u, 0u);                 // Zh | Zl
    uint4* r = reinterpret_cast<uint4*>(c.Rh());
    for (int i = tid; i < (2 * TC5_KM * TC5_TILE * 2 + TC5_NB * TC5_TILE * 2) / 16; i += TC5_TILE) r[i] = make_uint4(0u, 0u, 0u, 0u);   // Rh | Rl | Ot
}

// Stage one tile -- ids / levels of its cells (registers: thread = cell), the one-hot level rows, the Z_cos
// rows as fp16 hi/lo -- and issue its scoring.  Independent of the penalty table, so the first tile of the
// NEXT block is staged and scored before the grid barrier is waited on.
// trb: first of 5 timeline slots of this tile (>= HMY_TRACE_SLOTS: none): stage begin, operands published,
// scores ready, epilogue done, R tile published.
template <int NC>
__device__ void tc5_stage_tile(Tc5Ctx<NC>& c, const HmyDev& st, const int* list, long long tb, int nt, int trb) {
    const int tid = threadIdx.x;
    hmy_trace(st, trb);
    tc5_wait_acc(c);
    c.valid = tid < nt;
    int cell = 0, combo = 0;
    if (c.valid) { cell = list ? list[tb + tid] : (int)(tb + tid); combo = st.combo[cell]; }
    c.cell = cell;
    c.sCell[tid] = cell;
    unsigned int mask = 0u;
#pragma unroll
    for (int v = 0; v < HMY_MAX_V; ++v) {
        c.lev[v] = (v < st.V) ? st.combo_lev[combo * st.V + v] : 0;
        if (v < st.V && c.valid) mask |= 1u << c.lev[v];
    }
    // level rows of this cell: element (level, cell) of the MN-major one-hot tile, 8 levels = one 16-byte core row
#pragma unroll
    for (int j = 0; j < TC5_NB / 8; ++j) {
        const unsigned int m8 = (mask >> (8 * j)) & 0xFFu;
        uint4 w;
        w.x = ((m8 & 1u) ? 0x3C00u : 0u) | ((m8 & 2u) ? 0x3C000000u : 0u);
        w.y = ((m8 & 4u) ? 0x3C00u : 0u) | ((m8 & 8u) ? 0x3C000000u : 0u);
        w.z = ((m8 & 16u) ? 0x3C00u : 0u) | ((m8 & 32u) ? 0x3C000000u : 0u);
        w.w = ((m8 & 64u) ? 0x3C00u : 0u) | ((m8 & 128u) ? 0x3C000000u : 0u);
        *reinterpret_cast<uint4*>(c.Ot() + j * TC5_O_SBO + (tid >> 3) * TC5_O_LBO + (tid & 7) * 16) = w;
    }
    __syncthreads();
    {
        // gather: a batch of loads per thread is issued before its first conversion / store
        constexpr int ZU = 8;
        const int dp = st.dp, dp4 = dp >> 2, total = nt * dp4;
        for (int base = 0; base < total; base += ZU * TC5_TILE) {
            float4 zr[ZU];
#pragma unroll
            for (int u = 0; u < ZU; ++u) {
                const int i = base + tid + u * TC5_TILE;
                if (i < total) {
                    const int row = hmy_div(i, c.mg_dp4), c4 = i - row * dp4;
                    zr[u] = __ldg(reinterpret_cast<const float4*>(st.Zcos + (size_t)c.sCell[row] * dp) + c4);
                }
            }
#pragma unroll
            for (int u = 0; u < ZU; ++u) {
                const int i = base + tid + u * TC5_TILE;
                if (i < total) {
                    const int row = hmy_div(i, c.mg_dp4), c4 = i - row * dp4;
                    uint2 hi, lo;
                    split2(zr[u].x * HMY_OPSCALE, zr[u].y * HMY_OPSCALE, hi.x, lo.x);
                    split2(zr[u].z * HMY_OPSCALE, zr[u].w * HMY_OPSCALE, hi.y, lo.y);
                    const int off = (row >> 3) * TC5_Z_SBO_K + (c4 >> 1) * TC5_Z_LBO_K + (row & 7) * 16 + (c4 & 1) * 8;
                    *reinterpret_cast<uint2*>(c.Zh() + off) = hi;
                    *reinterpret_cast<uint2*>(c.Zl() + off) = lo;
                }
            }
        }
    }
    tc5_publish();
    hmy_trace(st, trb + 1);
    if (tid == 0) {
        const unsigned int sb = smem_u32(c.base), id_score = tc5_idesc(TC5_TILE, 16 * NC, 0, 0);
        const unsigned long long dZh = tc5_desc(sb + TC5_OFF_ZH, TC5_Z_LBO_K, TC5_Z_SBO_K), dZl = tc5_desc(sb + TC5_OFF_ZL, TC5_Z_LBO_K, TC5_Z_SBO_K);
        const unsigned long long dYh = tc5_desc(sb + TC5_OFF_YH, TC5_Y_LBO, TC5_Y_SBO), dYl = tc5_desc(sb + tc5_off_yl(NC), TC5_Y_LBO, TC5_Y_SBO);
        for (int ks = 0; ks < c.dt; ++ks) {
            const unsigned long long o = (unsigned long long)((ks * 2 * TC5_Z_LBO_K) >> 4);     // two 8-wide PC chunks per K = 16
            tc5_mma(c.tmem, dZl + o, dYh + o, id_score, ks > 0 ? 1u : 0u);
            tc5_mma(c.tmem, dZh + o, dYl + o, id_score, 1u);
            tc5_mma(c.tmem, dZh + o, dYh + o, id_score, 1u);
        }
        tc5_commit(&c.bar[0]);
    }
}

// Epilogue of the staged tile (thread = cell) and its contribution to the sums.
template <int NC>
__device__ void tc5_finish_tile(Tc5Ctx<NC>& c, const HmyDev& st, bool init, int nt, int trb) {
    constexpr int KT2 = 16 * NC;
    const int tid = threadIdx.x, K = st.K, Kp = st.Kp, V = st.V;
    tc5_wait(&c.bar[0], c.ph_score & 1u);
    c.ph_score++;
    hmy_trace(st, trb + 2);
    float E[KT2];
    float ss = 0.f, sp = 0.f, sd = 0.f;
    // ---- S = exp(-dist/sigma) (harmony.py:466-467), times the penalty (harmony.py:500)
#pragma unroll
    for (int ch = 0; ch < NC; ++ch) {
        float a[16];
        tc5_ld16(c.tmem + c.lane_base + (unsigned int)(16 * ch), a);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = 16 * ch + 4 * q;
            const float4 k1 = *reinterpret_cast<const float4*>(c.c1() + col);
            const float4 k3 = *reinterpret_cast<const float4*>(c.c3() + col);
            float4 pen = make_float4(1.f, 1.f, 1.f, 1.f);
            if (!init) {
                pen = *reinterpret_cast<const float4*>(c.Ps() + c.lev[0] * KT2 + col);
#pragma unroll
                for (int v = 1; v < HMY_MAX_V; ++v)            // more covariates: the factors add (harmony.py:500)
                    if (v < V) {
                        const float4 u = *reinterpret_cast<const float4*>(c.Ps() + c.lev[v] * KT2 + col);
                        pen.x += u.x; pen.y += u.y; pen.z += u.z; pen.w += u.w;
                    }
            }
            const float k1v[4] = {k1.x, k1.y, k1.z, k1.w}, k3v[4] = {k3.x, k3.y, k3.z, k3.w}, pv[4] = {pen.x, pen.y, pen.z, pen.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // t = (dist / sigma) log2 e >= 0; columns beyond K give s = 0
                const float t = fmaf(-a[4 * q + e], k1v[e], k1v[e] * 1048576.0f);
                const float s = (col + e < K) ? ex2_approx(-t) : 0.f;
                ss += s;
                const float ev = s * pv[e];
                sp += ev;
                sd = fmaf(k3v[e], ev * t, sd);                  // dist = t * c3: sum S*pen*dist for the objective (harmony.py:399)
                E[col + e] = ev;
            }
        }
    }
    // R = (S/sumS) pen / max(sum (S/sumS) pen, 1e-8)   (harmony.py:468, :500-503)
    const float is = 1.f / ss;
    const float sc = c.valid ? is / fmaxf(sp * is, 1e-8f) : 0.f;
    float* Rg = st.R + (size_t)c.cell * Kp;
    float oe = 0.f;
#pragma unroll
    for (int c0 = 0; c0 < KT2; c0 += 8) {
        const float4 k3a = *reinterpret_cast<const float4*>(c.c3() + c0), k3b = *reinterpret_cast<const float4*>(c.c3() + c0 + 4);
        const float k3v[8] = {k3a.x, k3a.y, k3a.z, k3a.w, k3b.x, k3b.y, k3b.z, k3b.w};
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[j] = E[c0 + j] * sc;
            oe = fmaf(k3v[j], (r[j] > 0.f ? r[j] * lg2_approx(r[j]) : 0.f), oe);   // sigma r ln r (harmony.py:402)
        }
        if (c.valid) {
            if (c0 < Kp) *reinterpret_cast<float4*>(Rg + c0) = make_float4(r[0], r[1], r[2], r[3]);
            if (c0 + 4 < Kp) *reinterpret_cast<float4*>(Rg + c0 + 4) = make_float4(r[4], r[5], r[6], r[7]);
        }
        uint4 hi, lo;
        split2(r[0] * HMY_OPSCALE, r[1] * HMY_OPSCALE, hi.x, lo.x);
        split2(r[2] * HMY_OPSCALE, r[3] * HMY_OPSCALE, hi.y, lo.y);
        split2(r[4] * HMY_OPSCALE, r[5] * HMY_OPSCALE, hi.z, lo.z);
        split2(r[6] * HMY_OPSCALE, r[7] * HMY_OPSCALE, hi.w, lo.w);
        const int off = (c0 >> 3) * TC5_R_SBO + (tid >> 3) * TC5_R_LBO + (tid & 7) * 16;      // element (cluster c0.., cell tid)
        *reinterpret_cast<uint4*>(c.Rh() + off) = hi;
        *reinterpret_cast<uint4*>(c.Rl() + off) = lo;
    }
    if (c.valid) { c.objd += (double)(sc * sd); c.obje += (double)oe; }
    hmy_trace(st, trb + 3);
    tc5_publish();
    hmy_trace(st, trb + 4);
    // ---- sums over the cells of the tile (K-dim = cells, 16 per step)
    if (tid == 0) {
        const unsigned int sb = smem_u32(c.base), id_y = tc5_idesc(TC5_KM, TC5_DP, 1, 1), id_o = tc5_idesc(TC5_KM, TC5_NB, 1, 1);
        const unsigned long long dRh = tc5_desc(sb + TC5_OFF_RH, TC5_R_LBO, TC5_R_SBO), dRl = tc5_desc(sb + TC5_OFF_RL, TC5_R_LBO, TC5_R_SBO);
        const unsigned long long dZh = tc5_desc(sb + TC5_OFF_ZH, TC5_Z_LBO_MN, TC5_Z_SBO_MN), dZl = tc5_desc(sb + TC5_OFF_ZL, TC5_Z_LBO_MN, TC5_Z_SBO_MN);
        const unsigned long long dOt = tc5_desc(sb + TC5_OFF_OT, TC5_O_LBO, TC5_O_SBO);
        const int ksteps = (nt + 15) >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {
            const unsigned long long ro = (unsigned long long)((ks * 2 * TC5_R_LBO) >> 4);
            const unsigned long long zo = (unsigned long long)((ks * 2 * TC5_Z_LBO_MN) >> 4);
            const unsigned long long oo = (unsigned long long)((ks * 2 * TC5_O_LBO) >> 4);
            tc5_mma(c.tmem + TC5_COL_Y, dRl + ro, dZh + zo, id_y, (c.y_started || ks > 0) ? 1u : 0u);
            tc5_mma(c.tmem + TC5_COL_Y, dRh + ro, dZl + zo, id_y, 1u);
            tc5_mma(c.tmem + TC5_COL_Y, dRh + ro, dZh + zo, id_y, 1u);
            tc5_mma(c.tmem + TC5_COL_O, dRl + ro, dOt + oo, id_o, (c.o_started || ks > 0) ? 1u : 0u);
            tc5_mma(c.tmem + TC5_COL_O, dRh + ro, dOt + oo, id_o, 1u);
        }
        tc5_commit(&c.bar[1]);
    }
    c.acc_pending = true; c.y_started = true; c.o_started = true;
}

// End of a block: thread = cluster adds its row of the level sums to Dnew[blk] / Ofresh (harmony.py:506-507)
template <int NC>
__device__ void tc5_flush_block(Tc5Ctx<NC>& c, const HmyDev& st, int blk) {
    if (!c.o_started) return;
    tc5_wait_acc(c);
    const int tid = threadIdx.x;
    float v0[16], v1[16];
    tc5_ld16(c.tmem + c.lane_base + (unsigned int)TC5_COL_O, v0);
    tc5_ld16(c.tmem + c.lane_base + (unsigned int)(TC5_COL_O + 16), v1);
    if (tid < st.K) {
        float* dn = st.Dnew + (size_t)blk * st.B * st.K + tid;
        double* of = st.Ofresh + tid;
#pragma unroll
        for (int b = 0; b < TC5_NB; ++b) {
            const float x = (b < 16 ? v0[b & 15] : v1[b & 15]) * (1.0f / HMY_OPSCALE);
            if (b < st.B && x != 0.f) {
                atomicAdd(dn + (size_t)b * st.K, x);
                atomicAdd(of + (size_t)b * st.K, (double)x);
            }
        }
    }
    c.o_started = false;
}

// One block of update_R (harmony.py:495-509) for this CTA's cells, or the init assignment
// (harmony.py:380-389) when init = true.  staged_tb: first list index of the tile tc5_stage_tile already
// prepared and scored (-1: none).
template <int NC>
__device__ void tc5_process_block(Tc5Ctx<NC>& c, const HmyDev& st, int blk, const int* list,
                                  long long lbeg, long long lend, bool init, long long staged_tb) {
    int trb = (blk == 5 && !init) ? 64 : HMY_TRACE_SLOTS;          // per-tile timeline of block 5 (option "trace")
    for (long long tb = lbeg; tb < lend; tb += TC5_TILE) {
        const int nt = (int)min((long long)TC5_TILE, lend - tb);
        if (tb != staged_tb) tc5_stage_tile(c, st, list, tb, nt, trb);
        tc5_finish_tile(c, st, init, nt, trb);
        if (trb < HMY_TRACE_SLOTS - 10) trb += 5; else trb = HMY_TRACE_SLOTS;
    }
    tc5_flush_block(c, st, blk);
}

// End of a round: centroid sums (thread = cluster) and objective sums
template <int NC>
__device__ void tc5_flush_round(Tc5Ctx<NC>& c, const HmyDev& st) {
    const int tid = threadIdx.x;
    if (c.y_started) {
        tc5_wait_acc(c);
#pragma unroll
        for (int ch = 0; ch < TC5_DP / 16; ++ch) {
            float v[16];
            tc5_ld16(c.tmem + c.lane_base + (unsigned int)(TC5_COL_Y + 16 * ch), v);
            if (tid < st.K) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int jj = 16 * ch + j;
                    if (jj < st.d && v[j] != 0.f) atomicAdd(&st.Yacc[(size_t)tid * st.dp + jj], (double)v[j] * (double)HMY_ACCSCALE);
                }
            }
        }
        c.y_started = false;
    }
    const double a = warp_sum_d(c.objd), b = warp_sum_d(c.obje);
    if ((tid & 31) == 0) { atomicAdd(&st.obj[0], a); atomicAdd(&st.obj[1], b); }
    c.objd = 0.0; c.obje = 0.0;
}

// ---- kernel -----------------------------------------------------------------------------------------
template <int NC, bool FUSED>
__global__ void __launch_bounds__(TC5_TILE, 1) k_round_tc5(HmyDev st, int mode, unsigned int gen_base) {
    extern __shared__ __align__(1024) unsigned char smem_tc5[];
    unsigned char* const smem = smem_tc5;
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ unsigned int s_tmem;
    const Tc5Smem p = tc5_smem_plan(st.B, NC);
    const bool multi_any = FUSED && st.xworld > 1;
    const unsigned int G = multi_any ? gridDim.x - 1u : gridDim.x;          // worker CTAs
    double* sRow = (double*)(smem + p.off_misc);
    double* sRed = sRow + 256;
    int* sFlag = (int*)(sRed + 8);
    if (multi_any && blockIdx.x == G) {
        comm_cta_main(st, mode, gen_base, G, sRow, sRed);
        return;
    }
    const int tid = threadIdx.x, warp = tid >> 5;

    // the parts of the mma.sync kernel that are reused unchanged (phase 0, per-CTA tables) see this context
    MmaCtx<2 * NC, 1> m;
    m.Rh = (__half*)(smem + p.off_Rh); m.Rl = (__half*)(smem + p.off_Rl);
    m.sCell = (int*)(smem + p.off_cell);
    m.Ps = (float*)(smem + p.off_Ps); m.Os = (float*)(smem + p.off_Os);
    m.sPrb = (float*)(smem + p.off_prb); m.sTheta = m.sPrb + st.B;
    m.RSH = p.RSH; m.KT2 = p.KT2;
    m.mg_kp4 = hmy_magic(st.Kp >> 2); m.mg_K = hmy_magic(st.K);

    Tc5Ctx<NC> c;
    c.base = smem;
    c.sCell = m.sCell;
    c.bar = s_bar;
    c.ph_score = 0u; c.ph_acc = 0u; c.acc_pending = false; c.y_started = false; c.o_started = false;
    c.dt = (st.d + 15) >> 4;
    c.mg_dp4 = hmy_magic(st.dp >> 2);
    c.cell = 0; c.valid = false;
#pragma unroll
    for (int v = 0; v < HMY_MAX_V; ++v) c.lev[v] = 0;
    c.objd = 0.0; c.obje = 0.0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TC5_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const long long c0 = (long long)blockIdx.x * st.N / G, c1 = (long long)(blockIdx.x + 1) * st.N / G;
    hmy_trace(st, 0);
    tc5_load_round_constants(c, st, m.Os, m.sPrb, m.sTheta);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.tmem = s_tmem;
    c.lane_base = (unsigned int)(32 * warp) << 16;

    if (mode == 1) {
        tc5_process_block(c, st, 0, nullptr, c0, c1, true, -1);
        tc5_flush_round(c, st);
        if (multi_any) worker_barrier(st, gen_base + 1u);
        else grid_barrier_serial(st, G, gen_base + 1u, sFlag, [&]() { serial_finalize(st, 1, sRow, sRed); });
    } else {
        if (st.nblk <= 32) mma_phase0(m, st, c0, c1);
        hmy_trace(st, 1);
        unsigned int gen = gen_base + 1u;
        mma_load_O(m, st);                       // O is only written by the finalize of the previous launch
        long long staged = -1;
        {
            long long nb, ne;
            block_share(st, 0, blockIdx.x, G, nb, ne);
            if (nb < ne) { tc5_stage_tile(c, st, st.list, nb, (int)min((long long)TC5_TILE, ne - nb), HMY_TRACE_SLOTS); staged = nb; }
        }
        const bool multi = multi_any && !st.xrelaxed;      // exact mode: one table exchange per block
        if (multi_any) worker_barrier(st, gen++);          // comm CTA: all Told sums are in (+ exchange)
        else grid_barrier(st, G, gen++);
        hmy_trace(st, 2);
        for (int blk = 0; blk < st.nblk; ++blk) {
            mma_update_tables(m, st, blk, multi);
            hmy_trace(st, 3 + 3 * blk);
            long long lb, le;
            block_share(st, blk, blockIdx.x, G, lb, le);
            tc5_process_block(c, st, blk, st.list, lb, le, false, staged);
            hmy_trace(st, 4 + 3 * blk);
            staged = -1;
            if (blk + 1 < st.nblk) {        // next block's first tile: stage and score before waiting at the barrier
                long long nb, ne;
                block_share(st, blk + 1, blockIdx.x, G, nb, ne);
                if (nb < ne) { tc5_stage_tile(c, st, st.list, nb, (int)min((long long)TC5_TILE, ne - nb), blk + 1 == 5 ? 64 : HMY_TRACE_SLOTS); staged = nb; }
                if (multi_any) worker_barrier(st, gen++);
                else grid_barrier(st, G, gen++);
            } else {
                tc5_flush_round(c, st);
                if (multi_any) worker_barrier(st, gen++);
                else grid_barrier_serial(st, G, gen++, sFlag, [&]() { serial_finalize(st, 0, sRow, sRed); });
            }
            hmy_trace(st, 5 + 3 * blk);
        }
    }
    // every MMA has been waited for (flush_block / flush_round); give the tensor memory back
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "n"(TC5_TMEM_COLS));
}
