// The k-means round of Harmony (harmony.py:437-513) on tcgen05 / tensor memory, warp-specialised.
//
// One persistent cooperative launch per round, ONE CTA per SM, 18 warps:
//
//   warp 16  PRODUCER   walks the CTA's share of every block's annotated cell list (int2 {cell, combo << 8 | block
//                       of the cell in the NEXT round}), builds 128-row tiles (a change of covariate combination
//                       starts on a 16-row boundary, so every 16-cell K step of the accumulation has ONE combination)
//                       and gathers the cells' Z_cos rows -- kept pre-split as fp16 hi | lo (hmy_common.cuh: Zs16) --
//                       with 16-byte cp.async straight into the K-major core-matrix layout of a 3-stage ring.
//   warp 17  MMA        one thread issues every tcgen05.mma:  scoring  D1[cell][cluster] = Z . Y^T  (fp16 two-way
//                       split, fp32 accumulate, two D1 buffers) as soon as a stage has landed, and the accumulation
//                       of a finished tile  D2y[cluster][PC] += R^T Z,  D2o[slot][cluster] += R^T 1,
//                       D2t[slot][cluster][next block] += R^T onehot(next block)  in tensor memory.
//   warps 0-15 EPILOGUE thread = (cell, quarter of the clusters): tcgen05.ld of the scores, exp2, penalty, the
//                       softmax / normalisation sums (four threads per cell meet through shared memory), objective
//                       terms in closed form (no log per entry), the new R as fp16 hi/lo into the MN-major operand
//                       tile.  Between blocks: flush D2o into the running O (fp64 atomics), grid barrier, penalty
//                       rows of the combinations this CTA meets (read from the running O: O(K) per CTA and block,
//                       independent of B).
//
// What is NOT done any more compared with k_round_mma (hmy_round_mma.cuh):
//   * no phase 0: the sums a block removes from O before its update (harmony.py:491-492) are the sums of the
//     assignments as of the END of the previous round grouped by THIS round's blocks; the previous round
//     accumulated them in tensor memory (D2t) because it already knew this round's block of every cell
//     (the host / the device permutation runs one round ahead) -- 400 B/cell of R re-reads gone;
//   * R is not written to HBM unless the host asks (write_R): nothing on the device reads it between rounds;
//     the host asks in the rounds after which cluster() can stop (harmony.py:455-458) because the ridge
//     correction and the R property need it -- another 400 B/cell gone in the other rounds;
//   * no K x B tables in shared memory and no O(B K) table update per CTA and block.
//
// Limits: K <= 128, d <= 64, <= 32 blocks, one GPU per launch domain (multi-GPU runs stay on k_round_mma).
#pragma once
#include "hmy_round_mma.cuh"

#define T5_TILE 128
#define T5_NZ 3
#define T5_SLOTS 4
#define T5_EPI_WARPS 16
#define T5_EPI_THREADS (32 * T5_EPI_WARPS)
#define T5_WARP_PROD 16
#define T5_WARP_MMA 17
#define T5_THREADS (32 * (T5_EPI_WARPS + 2))
// tensor-memory columns (512 allocated)
#define T5_COL_D1 0              // + 128 * buffer
#define T5_COL_Y 256
#define T5_COL_OT 320            // + 48 * slot: 16 columns that all hold the slot's column sums of R, then 32 columns
                                 // of those sums split by the cell's block in the next round (one MMA pair feeds both)
// shared memory (dynamic, 1024-byte aligned)
#define T5_SZ_ZPART 16384
#define T5_SZ_STAGE 32768
#define T5_SZ_RPART 32768
#define T5_OFF_Z 0
#define T5_OFF_RH (T5_NZ * T5_SZ_STAGE)
#define T5_OFF_RL (T5_OFF_RH + T5_SZ_RPART)
#define T5_OFF_NB (T5_OFF_RL + T5_SZ_RPART)          // MN-major B tile of 48 columns x 128 cells: 16 columns of 1.0 (written once),
#define T5_SZ_NB (6 * 2048)                           // then one-hot(block of the cell in the next round)
#define T5_OFF_YH (T5_OFF_NB + T5_SZ_NB)
__host__ __device__ constexpr int t5_off_yl(int NC) { return T5_OFF_YH + NC * 2048; }
__host__ __device__ constexpr int t5_off_ps(int NC) { return t5_off_yl(NC) + NC * 2048; }      // float2 [4][KT2] {pen, sigma ln pen}
__host__ __device__ constexpr int t5_off_ck(int NC) { return t5_off_ps(NC) + T5_SLOTS * 16 * NC * 8; }   // float2 [KT2] {-log2e / sigma, sigma}
__host__ __device__ constexpr int t5_off_xch(int NC) { return t5_off_ck(NC) + 16 * NC * 8; }   // float2 [2][4][128]
__host__ __device__ constexpr int t5_off_meta(int NC) { return t5_off_xch(NC) + 2 * 4 * 128 * 8; }
__host__ __device__ constexpr int t5_off_bar(int NC) { return t5_off_meta(NC) + T5_NZ * 1024; }
__host__ __device__ constexpr int t5_smem_bytes(int NC) { return t5_off_bar(NC) + 256; }
// per-stage tile description written by the producer: int cell[128] | u16 slotnb[128] | header
#define T5_META_SLOTNB 512
#define T5_META_HDR 768
#define T5_H_FLAGS 0
#define T5_H_BLK 1
#define T5_H_MASK 2              // slots used by this tile
#define T5_H_ALLOC 3             // slots that have a combination
#define T5_H_COMBO 4             // [4] combination of every slot (-1: free)
#define T5_H_KSLOT 8             // [2] slot of every 16-row K step, one byte each (0xFF: no cells)
#define T5_H_LEV 12              // u16 [4][8] one-hot rows of every slot's combination (16-byte aligned)
#define T5_F_EMPTY 1
#define T5_F_LAST_BLOCK 2
#define T5_F_LAST_ROUND 4
#define T5_F_EVICT 8
// mbarriers
#define T5_B_ZFULL 0             // [3] producer -> MMA / epilogue
#define T5_B_ACC 3               // [3] accumulation of the stage's tile done (tcgen05.commit)
#define T5_B_SFULL 6             // [2] scores in D1[d]
#define T5_B_SFREE 8             // [2] D1[d] read by every epilogue warp
#define T5_B_RFULL 10            // operand tiles of the accumulation written
#define T5_B_ODONE 11            // D2o of the block's last tile complete

// canonical no-swizzle layouts (8 rows x 16 bytes core matrices)
#define T5_LBO_K 128             // K-major: next 8-wide k chunk
#define T5_SBO_K 1024            //          next 8 rows
#define T5_Z_LBO_MN 1024         // the Z tile read MN-major (PC, cell): next 8 cells
#define T5_Z_SBO_MN 128          //                                      next 8 PCs
#define T5_R_LBO 128             // MN-major (cluster, cell): next 8 cells
#define T5_R_SBO 2048            //                           next 8 clusters

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long t5_desc(unsigned int saddr, unsigned int lbo, unsigned int sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFFu);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ unsigned int t5_idesc(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((unsigned int)a_mn_major << 15) | ((unsigned int)b_mn_major << 16) | ((unsigned int)(n >> 3) << 17) | ((unsigned int)(m >> 4) << 24);
}
__device__ __forceinline__ void t5_mma(unsigned int tmem, unsigned long long da, unsigned long long db, unsigned int idesc, unsigned int accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void t5_commit(unsigned int bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t5_arrive(unsigned int bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
// try_wait may suspend the thread for a system-dependent time before it reports "not yet": right for a thread that
// waits for ONE barrier (t5_wait), wrong for a role that polls several (it would sit out the time limit on the first
// one while the second has long completed): t5_poll is the non-blocking test
__device__ __forceinline__ bool t5_test(unsigned int bar, unsigned int parity) {
    unsigned int done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ bool t5_poll(unsigned int bar, unsigned int parity) {
    unsigned int done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
// A completion that never comes (bad descriptor, lost arrival) must become an error, not a hung cooperative grid
__device__ __forceinline__ void t5_wait(unsigned int bar, unsigned int parity) {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (!t5_test(bar, parity)) {
        if ((++spins & 255u) == 0u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void t5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t5_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void t5_ld16(unsigned int taddr, float* v) {
    unsigned int u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
}
__device__ __forceinline__ void t5_ld8(unsigned int taddr, float* v) {
    unsigned int u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(u[j]);
}
__device__ __forceinline__ float t5_ld1(unsigned int taddr) {
    unsigned int u;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(u) : "r"(taddr));
    return __uint_as_float(u);
}
__device__ __forceinline__ void t5_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void t5_cp16(unsigned int dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void t5_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// first cluster / number of 8-cluster groups of epilogue quarter q (the 16 NC columns split in whole groups of 8)
template <int NC> __device__ __forceinline__ int t5_g0(int q) { return (NC == 7) ? ((q < 2) ? 4 * q : 8 + 3 * (q - 2)) : (NC / 2) * q; }
template <int NC> __device__ __forceinline__ int t5_ng(int q) { return (NC == 7) ? ((q < 2) ? 4 : 3) : NC / 2; }

// ---- producer warp ----------------------------------------------------------------------------------------------
// Tile t of this CTA lives in stage t % 3.  Rows are filled from the block's share of the list; a new combination
// starts on a 16-row boundary; a combination without a slot when all four are taken closes the tile (the next one
// starts with an eviction: the epilogue flushes the per-slot accumulators before it goes on).
template <int NC>
__device__ void t5_producer(const HmyDev& st, int mode, unsigned char* smem, unsigned int bar0, unsigned int G) {
    const int lane = threadIdx.x & 31;
    const int dt = (st.d + 15) >> 4;
    const size_t zrow = (size_t)64 * dt;                       // bytes of one Zs16 row: hi[16 dt] | lo[16 dt] halves
    const unsigned char* Zs = reinterpret_cast<const unsigned char*>(st.Zs16);
    int sc0 = -1, sc1 = -1, sc2 = -1, sc3 = -1, nslots = 0;   // slot table (warp-uniform)
    unsigned short mylev = 0;                                  // lane = 8 slot + covariate: one-hot row of that slot's combination
    unsigned int t = 0;
    bool pending = false;                                      // tile t-1's copies are in flight, its zfull not yet signalled
    const int nblocks = (mode == 1) ? 1 : st.nblk;
    const unsigned int sbase = smem_u32(smem);
    if (mode != 1) {
        // Every combination this CTA meets in the round gets its slot up front, so that the epilogue follows its running
        // O rows from block 0: a slot first met in the middle of a round has to catch up on every finished block (measured
        // 1.1 us per block, 8 us with 8 GPUs -- and the whole grid, on every GPU, waits for that one CTA).  The lists are
        // sorted by position and positions by combination, so the CTA's share of block b covers the combinations from its
        // first entry's to its last entry's; lane b looks those two up.
        int cf = 0x7fffffff, cl = -1;
        for (int b = lane; b < nblocks; b += 32) {
            long long lb, le;
            block_share(st, b, blockIdx.x, G, lb, le);
            if (le > lb) { cf = min(cf, __ldg(&st.list2[lb]).y >> 8); cl = max(cl, __ldg(&st.list2[le - 1]).y >> 8); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { cf = min(cf, __shfl_xor_sync(0xffffffffu, cf, o)); cl = max(cl, __shfl_xor_sync(0xffffffffu, cl, o)); }
        for (int c = cf; c <= cl && nslots < T5_SLOTS; ++c) {       // more than four: the tiles evict as they go
            const int slot = nslots++;
            if (slot == 0) sc0 = c; else if (slot == 1) sc1 = c; else if (slot == 2) sc2 = c; else sc3 = c;
            if ((lane >> 3) == slot) mylev = (unsigned short)(((lane & 7) < st.V) ? st.combo_lev[c * st.V + (lane & 7)] : 0);
        }
    }
    for (int blk = 0; blk < nblocks; ++blk) {
        long long lb, le;
        if (mode == 1) { lb = (long long)blockIdx.x * st.N / G; le = (long long)(blockIdx.x + 1) * st.N / G; }
        else block_share(st, blk, blockIdx.x, G, lb, le);
        long long cur = lb;
        int pr_in_blk = 0;
        do {
            const unsigned int s = t % T5_NZ, use = t / T5_NZ;
            const int psl = (blk == 5 && pr_in_blk < 4 && lane == 0) ? 144 + 4 * pr_in_blk : HMY_TRACE_SLOTS;
            ++pr_in_blk;
            const unsigned int bacc = bar0 + 8u * (T5_B_ACC + s);
            if (!t5_poll(bacc, (use & 1u) ^ 1u)) {
                // the stage is still in use: hand over the previous tile first, then wait
                if (pending) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    t5_fence_async();
                    __syncwarp();
                    if (lane == 0) t5_arrive(bar0 + 8u * (T5_B_ZFULL + (t - 1u) % T5_NZ));
                    pending = false;
                }
                t5_wait(bacc, (use & 1u) ^ 1u);
            }
            hmy_trace_any(st, psl);
            if (lane == 0) hmy_trace_gap(st, 184, 3, 0, (unsigned long long)t);
            unsigned char* meta = smem + t5_off_meta(NC) + s * 1024;
            int* mcell = reinterpret_cast<int*>(meta);
            unsigned short* mslot = reinterpret_cast<unsigned short*>(meta + T5_META_SLOTNB);
            int* hdr = reinterpret_cast<int*>(meta + T5_META_HDR);
            // ---- candidate entries: i = lane + 32 j
            const int navail = (int)min((long long)T5_TILE, le - cur);
            int ecell[4], ecombo[4], enb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = lane + 32 * j;
                ecell[j] = 0; ecombo[j] = -1; enb[j] = 0;
                if (i < navail) {
                    if (mode == 1) {
                        const long long pos = cur + i;
                        ecell[j] = (int)pos; ecombo[j] = st.combo[pos]; enb[j] = st.blk_next[pos];
                    } else {
                        const int2 e = __ldg(&st.list2[cur + i]);
                        ecell[j] = e.x; ecombo[j] = e.y >> 8; enb[j] = e.y & 255;
                    }
                }
                mslot[i] = 0xFFFFu; mcell[i] = 0;
            }
            __syncwarp();
            int consumed = 0, row_base = 0, nrows = 0;
            unsigned int flags = 0u, mask = 0u;
            unsigned int ks_lo = 0xFFFFFFFFu, ks_hi = 0xFFFFFFFFu;
            bool evicted = false;
            while (consumed < navail && row_base < T5_TILE) {
                const int jj = consumed >> 5;
                const int mine_c = (jj == 0) ? ecombo[0] : (jj == 1) ? ecombo[1] : (jj == 2) ? ecombo[2] : ecombo[3];
                const int c = __shfl_sync(0xffffffffu, mine_c, consumed & 31);
                int mine = navail;
#pragma unroll
                for (int j = 3; j >= 0; --j) { const int i = lane + 32 * j; if (i >= consumed && i < navail && ecombo[j] != c) mine = i; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
                const int run_end = mine;
                int slot = (c == sc0) ? 0 : (c == sc1) ? 1 : (c == sc2) ? 2 : (c == sc3) ? 3 : -1;
                if (slot < 0) {
                    if (nslots == T5_SLOTS) {
                        if (row_base != 0 || evicted) break;                 // close the tile; the next one evicts
                        evicted = true; flags |= T5_F_EVICT;
                        sc0 = sc1 = sc2 = sc3 = -1; nslots = 0;
                    }
                    slot = nslots++;
                    if (slot == 0) sc0 = c; else if (slot == 1) sc1 = c; else if (slot == 2) sc2 = c; else sc3 = c;
                    // the combination's one-hot rows travel with every tile (the epilogue derives penalties from them)
                    if ((lane >> 3) == slot) mylev = (unsigned short)(((lane & 7) < st.V) ? st.combo_lev[c * st.V + (lane & 7)] : 0);
                }
                const int take = min(run_end - consumed, T5_TILE - row_base);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = lane + 32 * j;
                    if (i >= consumed && i < consumed + take) {
                        const int row = row_base + (i - consumed);
                        mcell[row] = ecell[j];
                        mslot[row] = (unsigned short)((slot << 8) | enb[j]);
                    }
                }
                for (int ks = row_base >> 4; ks <= (row_base + take - 1) >> 4; ++ks) {
                    if (ks < 4) ks_lo = (ks_lo & ~(0xFFu << (8 * ks))) | ((unsigned int)slot << (8 * ks));
                    else ks_hi = (ks_hi & ~(0xFFu << (8 * (ks - 4)))) | ((unsigned int)slot << (8 * (ks - 4)));
                }
                mask |= 1u << slot;
                consumed += take;
                nrows = row_base + take;
                row_base = (nrows + 15) & ~15;
            }
            cur += consumed;
            if (consumed == 0 && navail > 0) { /* cannot happen: an eviction always frees a slot */ }
            if (nrows == 0) flags |= T5_F_EMPTY;
            if (cur >= le) flags |= T5_F_LAST_BLOCK | ((blk + 1 == nblocks) ? T5_F_LAST_ROUND : 0u);
            reinterpret_cast<unsigned short*>(hdr + T5_H_LEV)[lane] = mylev;
            __syncwarp();
            if (lane == 0) {
                hdr[T5_H_FLAGS] = (int)flags; hdr[T5_H_BLK] = blk; hdr[T5_H_MASK] = (int)mask; hdr[T5_H_ALLOC] = (1 << nslots) - 1;
                hdr[T5_H_COMBO + 0] = sc0; hdr[T5_H_COMBO + 1] = sc1; hdr[T5_H_COMBO + 2] = sc2; hdr[T5_H_COMBO + 3] = sc3;
                hdr[T5_H_KSLOT] = (int)ks_lo; hdr[T5_H_KSLOT + 1] = (int)ks_hi;
            }
            __syncwarp();
            hmy_trace_any(st, psl + 1);
            if (lane == 0) hmy_trace_gap(st, 184, 3, 1, (unsigned long long)t);
            // ---- gather: 8 rows x 4 chunks of 16 bytes per instruction (conflict-free core-matrix writes).  Straight-line
            // code: the cell ids of this lane's 16 rows first (one batch of shared-memory loads), then nothing but copies
            {
                const int r8 = lane & 7, cs = lane >> 3;
                const unsigned int zs = sbase + T5_OFF_Z + s * T5_SZ_STAGE + (unsigned int)r8 * 16u;
                const int ngroups = (nrows + 7) >> 3;
                int cg[16]; unsigned int vmask = 0u;
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    cg[g] = mcell[8 * g + r8];
                    if (g < ngroups && mslot[8 * g + r8] != 0xFFFFu) vmask |= 1u << g;
                }
                unsigned int doff[4]; int soff[4];
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    const int ch = 4 * cq + cs;                    // chunk of the row: [0, 2 dt) hi, [2 dt, 4 dt) lo
                    const int part = (ch >= 2 * dt) ? 1 : 0, cc = ch - part * 2 * dt;
                    doff[cq] = (unsigned int)part * T5_SZ_ZPART + (unsigned int)cc * T5_LBO_K; soff[cq] = ch * 16;
                }
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    if ((vmask >> g) & 1u) {
                        const unsigned char* src = Zs + (size_t)cg[g] * zrow;
                        const unsigned int drow = zs + (unsigned int)g * T5_SBO_K;
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq)
                            if (cq < dt) t5_cp16(drow + doff[cq], src + soff[cq]);
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            hmy_trace_any(st, psl + 2);
            if (lane == 0) hmy_trace_gap(st, 184, 3, 2, (unsigned long long)t);
            if (pending) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                t5_fence_async();
                __syncwarp();
                if (lane == 0) t5_arrive(bar0 + 8u * (T5_B_ZFULL + (t - 1u) % T5_NZ));
            }
            pending = true;
            ++t;
        } while (cur < le);
    }
    if (pending) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        t5_fence_async();
        __syncwarp();
        if (lane == 0) t5_arrive(bar0 + 8u * (T5_B_ZFULL + (t - 1u) % T5_NZ));
    }
}

// ---- MMA warp -----------------------------------------------------------------------------------------------------
// The whole warp runs the loop with uniform control flow and one elected lane issues (a tcgen05.mma in divergent
// code is wrapped in a per-instruction election loop by the compiler; at N <= 64 that issue overhead, not the tensor
// pipe, set the pace: experiments/tcgen05_mma_rate.cu).  Everything the loop reads from shared memory passes through
// a warp reduction so that the compiler knows it is warp-uniform (descriptor arithmetic on the uniform datapath).
__device__ __forceinline__ bool t5_elect() {
    unsigned int pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ unsigned int t5_uni(unsigned int x) { return __reduce_or_sync(0xffffffffu, x); }

template <int NC>
__device__ void t5_mma_warp(const HmyDev& st, unsigned char* smem, unsigned int bar0, unsigned int tmem) {
    const int lane = threadIdx.x & 31;
    const int dt = (st.d + 15) >> 4;
    const unsigned int sb = smem_u32(smem);
    const unsigned int id_score = t5_idesc(T5_TILE, 16 * NC, 0, 0), id_y = t5_idesc(128, 16 * dt, 1, 1), id_ot = t5_idesc(128, 48, 1, 1);
    const unsigned long long dYh = t5_desc(sb + T5_OFF_YH, T5_LBO_K, T5_SBO_K), dYl = t5_desc(sb + t5_off_yl(NC), T5_LBO_K, T5_SBO_K);
    const unsigned long long dRh = t5_desc(sb + T5_OFF_RH, T5_R_LBO, T5_R_SBO), dRl = t5_desc(sb + T5_OFF_RL, T5_R_LBO, T5_R_SBO);
    const unsigned long long dNb = t5_desc(sb + T5_OFF_NB, T5_R_LBO, T5_R_SBO);
    const unsigned long long dZK = t5_desc(sb + T5_OFF_Z, T5_LBO_K, T5_SBO_K);            // stage 0, hi part, K-major view
    const unsigned long long dZM = t5_desc(sb + T5_OFF_Z, T5_Z_LBO_MN, T5_Z_SBO_MN);      // the same bytes MN-major
    constexpr unsigned long long STAGE16 = T5_SZ_STAGE >> 4, PART16 = T5_SZ_ZPART >> 4;
    unsigned int ts = 0, ta = 0;                       // next tile to score / to accumulate
    bool score_end = false, y_started = false;
    unsigned int ot_started = 0u;                      // per slot: its accumulator columns hold sums
    unsigned int idle = 0; unsigned long long idle_t0 = 0;
    int sc_in_blk = 0, ac_in_blk = 0;                  // timeline (option "trace"): tiles of block 5, slots 128 + 4 i + {0..3}
    for (;;) {
        bool progressed = false;
        if (!score_end) {
            const unsigned int s = ts % T5_NZ, d = ts & 1u;
            const bool ready = t5_poll(bar0 + 8u * (T5_B_ZFULL + s), (ts / T5_NZ) & 1u) && t5_poll(bar0 + 8u * (T5_B_SFREE + d), ((ts >> 1) & 1u) ^ 1u);
            if (__all_sync(0xffffffffu, ready)) {
                t5_fence_after();
                const int* hdr = reinterpret_cast<const int*>(smem + t5_off_meta(NC) + s * 1024 + T5_META_HDR);
                const unsigned int flags_v = (unsigned int)hdr[T5_H_FLAGS];
                const bool f_empty = __any_sync(0xffffffffu, (flags_v & T5_F_EMPTY) != 0u);
                const bool f_lastb = __any_sync(0xffffffffu, (flags_v & T5_F_LAST_BLOCK) != 0u);
                const bool f_lastr = __any_sync(0xffffffffu, (flags_v & T5_F_LAST_ROUND) != 0u);
                const int tsl = (hdr[T5_H_BLK] == 5 && sc_in_blk < 4 && lane == 0) ? 128 + 4 * sc_in_blk : HMY_TRACE_SLOTS;
                hmy_trace_any(st, tsl);
                if (lane == 0) hmy_trace_gap(st, 192, 4, 0, ((unsigned long long)hdr[T5_H_BLK] << 8) | (unsigned long long)sc_in_blk);
                if (!f_empty) {
                    const unsigned long long dZh = dZK + (unsigned long long)s * STAGE16, dZl = dZh + PART16;
                    const unsigned int dcol = tmem + T5_COL_D1 + 128u * d;
                    // straight-line code on purpose (here and below): this warp shares its instruction cache with four
                    // epilogue warps that stream through far more code than it holds -- a loop's backward branch
                    // refetches its lines every iteration (measured: ~190 cycles per iteration of an EMPTY loop)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (ks < dt) {
                            const unsigned long long o = (unsigned long long)((ks * 2 * T5_LBO_K) >> 4);
                            if (t5_elect()) {
                                t5_mma(dcol, dZl + o, dYh + o, id_score, ks > 0 ? 1u : 0u);
                                t5_mma(dcol, dZh + o, dYl + o, id_score, 1u);
                                t5_mma(dcol, dZh + o, dYh + o, id_score, 1u);
                            }
                        }
                    }
                }
                if (t5_elect()) t5_commit(bar0 + 8u * (T5_B_SFULL + d));
                hmy_trace_any(st, tsl + 1);
                if (lane == 0) hmy_trace_gap(st, 192, 4, 1, ((unsigned long long)hdr[T5_H_BLK] << 8) | (unsigned long long)sc_in_blk);
                sc_in_blk = f_lastb ? 0 : sc_in_blk + 1;
                if (f_lastr) score_end = true;
                ++ts;
                progressed = true;
            }
        }
        if (ta < ts && __all_sync(0xffffffffu, t5_poll(bar0 + 8u * T5_B_RFULL, ta & 1u))) {
            t5_fence_after();
            const unsigned int s = ta % T5_NZ;
            const int* hdr = reinterpret_cast<const int*>(smem + t5_off_meta(NC) + s * 1024 + T5_META_HDR);
            // Every decision below goes through a warp vote (a uniform predicate the compiler can branch on without
            // treating the region as divergent -- a branch on a value loaded from shared memory makes it move every
            // descriptor into uniform registers again for every MMA: ~15 R2UR per K step, measured 120 cycles per step).
            const unsigned int flags_v = (unsigned int)hdr[T5_H_FLAGS];
            const bool f_empty = __any_sync(0xffffffffu, (flags_v & T5_F_EMPTY) != 0u);
            const bool f_lastb = __any_sync(0xffffffffu, (flags_v & T5_F_LAST_BLOCK) != 0u);
            const bool f_lastr = __any_sync(0xffffffffu, (flags_v & T5_F_LAST_ROUND) != 0u);
            const bool f_evict = __any_sync(0xffffffffu, (flags_v & T5_F_EVICT) != 0u);
            const int blk_v = hdr[T5_H_BLK];
            const int tsl = (blk_v == 5 && ac_in_blk < 4 && lane == 0) ? 130 + 4 * ac_in_blk : HMY_TRACE_SLOTS;
            hmy_trace_any(st, tsl);
            if (lane == 0) hmy_trace_gap(st, 192, 4, 2, ((unsigned long long)blk_v << 8) | (unsigned long long)ac_in_blk);
            if (f_evict) ot_started = 0u;
            const bool cyc = st.trace != nullptr && blk_v == 5 && ac_in_blk == 0 && lane == 0;
#define T5_CYC(i_) do { if (cyc) st.trace[(size_t)blockIdx.x * HMY_TRACE_SLOTS + 160 + (i_)] = (unsigned long long)clock64(); } while (0)
            T5_CYC(0);
            const unsigned int kslo = (unsigned int)hdr[T5_H_KSLOT], kshi = (unsigned int)hdr[T5_H_KSLOT + 1];
            T5_CYC(1);
            if (!f_empty) {
                // the per-slot column sums first: they are what the end of a block waits for.  K steps without cells
                // (slot byte 0xFF) are not issued; their predicate is a vote as well.
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const unsigned int sl = ((ks < 4 ? kslo >> (8 * ks) : kshi >> (8 * (ks - 4))) & 0xFFu);
                    const bool on = __any_sync(0xffffffffu, sl != 0xFFu);
                    const unsigned int slot = sl & 3u;
                    const unsigned long long ro = (unsigned long long)((ks * 2 * T5_R_LBO) >> 4);
                    const unsigned int dcol = tmem + T5_COL_OT + 48u * slot;
                    if (on && !(st.dbg & 2)) {
                        if (t5_elect()) {
                            t5_mma(dcol, dRh + ro, dNb + ro, id_ot, (ot_started >> slot) & 1u);
                            t5_mma(dcol, dRl + ro, dNb + ro, id_ot, 1u);
                        }
                    }
                    if (on) ot_started |= 1u << slot;
                }
            }
            T5_CYC(2);
            if (f_lastb) { if (t5_elect()) t5_commit(bar0 + 8u * T5_B_ODONE); }
            if (!f_empty) {
                const unsigned long long dZh = dZM + (unsigned long long)s * STAGE16, dZl = dZh + PART16;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const unsigned int sl = ((ks < 4 ? kslo >> (8 * ks) : kshi >> (8 * (ks - 4))) & 0xFFu);
                    const bool on = __any_sync(0xffffffffu, sl != 0xFFu);
                    const unsigned long long ro = (unsigned long long)((ks * 2 * T5_R_LBO) >> 4);
                    const unsigned long long zo = (unsigned long long)((ks * 2 * T5_Z_LBO_MN) >> 4);
                    if (on && !(st.dbg & 1)) {
                        if (t5_elect()) {
                            t5_mma(tmem + T5_COL_Y, dRl + ro, dZh + zo, id_y, y_started ? 1u : 0u);
                            t5_mma(tmem + T5_COL_Y, dRh + ro, dZl + zo, id_y, 1u);
                            t5_mma(tmem + T5_COL_Y, dRh + ro, dZh + zo, id_y, 1u);
                        }
                    }
                    if (on) y_started = true;
                }
            }
            T5_CYC(3);
            if (t5_elect()) t5_commit(bar0 + 8u * (T5_B_ACC + s));
            T5_CYC(4);
            hmy_trace_any(st, tsl + 1);
            if (lane == 0) hmy_trace_gap(st, 192, 4, 3, ((unsigned long long)blk_v << 8) | (unsigned long long)ac_in_blk);
            T5_CYC(5);
            ac_in_blk = f_lastb ? 0 : ac_in_blk + 1;
            ++ta;
            progressed = true;
            if (f_lastr) break;
        }
        if (progressed) { idle = 0; idle_t0 = 0; }
        else {
            // nothing to issue: sleep on the barrier that normally completes next (try_wait suspends the warp until
            // the phase completes or a time limit passes) instead of spinning -- a busy poll loop here took issue slots
            // from the four epilogue warps of this scheduler, i.e. from a quarter of every tile's rows
            if (ta < ts) (void)t5_test(bar0 + 8u * T5_B_RFULL, ta & 1u);
            else if (!score_end) (void)t5_test(bar0 + 8u * (T5_B_ZFULL + ts % T5_NZ), (ts / T5_NZ) & 1u);
        }
        if (!progressed && (++idle & 255u) == 0u) {        // nothing to issue for seconds: an error, not a hang
            unsigned long long tn; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tn));
            if (idle_t0 == 0) idle_t0 = tn; else if (tn - idle_t0 > 4000000000ull) __trap();
        }
    }
}

// ---- small tables and their cross-GPU form --------------------------------------------------------------------------
// Per block the round needs two tables of nD = B K + K floats: [level][cluster] sums and, behind them, the per-cluster
// row sums (every cell has exactly one level of covariate 0, so the row sum of R is a column sum over those levels):
//   Told[blk]  what block blk takes out of O (accumulated by the PREVIOUS launch, complete and already summed over ranks)
//   Dnew[blk]  what block blk put back (atomics of this launch; complete on this rank once its barrier has been passed)
// With the cells sharded over GPUs (MULTI) nothing waits at a cross-GPU barrier: after the rank's own barrier every CTA
// pushes its slice of Dnew[blk] to every rank as LL packets {value, sequence number of the launch} (8-byte stores are
// atomic, the reader polls the packet itself) into slot [round parity][blk][source rank] of the peer-mapped exchange
// buffer, and whoever needs an element sums the W sources in rank order -- identical totals on every rank.  The end of
// a round moves the next round's Told table, the centroid sums and the objective sums the same way.
__device__ __forceinline__ size_t t5_nD(const HmyDev& st) { return (size_t)st.B * st.K + st.K; }

__device__ __forceinline__ float t5_ll_read(const uint2* p, unsigned int seq) {
    uint2 v = ld_volatile_v2(p);
    unsigned int spins = 0; unsigned long long t0 = 0;
    while (v.y != seq) {
        if ((++spins & 1023u) == 0u) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t; else if (t - t0 > 8000000000ull) __trap();
        }
        v = ld_volatile_v2(p);
    }
    return __uint_as_float(v.x);
}
// sum over the W sources of one LL element (p: source 0's packet, stride: packets between sources), in rank order.
// All W loads are issued before the first sequence number is looked at (one round trip when the data is there, which
// is the usual case: a dependent poll per source cost W round trips -- measured 8 us per block at 8 GPUs).
static __device__ __noinline__ float t5_ll_sum(const uint2* p, size_t stride, int W, unsigned int seq) {
    uint2 v[HMY_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < HMY_MAX_WORLD; ++r) if (r < W) v[r] = ld_volatile_v2(p + (size_t)r * stride);
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < HMY_MAX_WORLD; ++r)
        if (r < W) s += (v[r].y == seq) ? __uint_as_float(v[r].x) : t5_ll_read(p + (size_t)r * stride, seq);
    return s;
}
// the same for a double that travels as two packets (low word, high word)
static __device__ __noinline__ double t5_ll_sum_d(const uint2* p, size_t stride, int W, unsigned int seq) {
    uint2 lo[HMY_MAX_WORLD], hi[HMY_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < HMY_MAX_WORLD; ++r) if (r < W) { lo[r] = ld_volatile_v2(p + (size_t)r * stride); hi[r] = ld_volatile_v2(p + (size_t)r * stride + 1); }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < HMY_MAX_WORLD; ++r)
        if (r < W) {
            const unsigned int l = (lo[r].y == seq) ? lo[r].x : __float_as_uint(t5_ll_read(p + (size_t)r * stride, seq));
            const unsigned int h = (hi[r].y == seq) ? hi[r].x : __float_as_uint(t5_ll_read(p + (size_t)r * stride + 1, seq));
            s += __longlong_as_double((long long)(((unsigned long long)h << 32) | l));
        }
    return s;
}
// element e of this launch's Dnew[i], summed over ranks
template <bool MULTI>
__device__ __forceinline__ float t5_dnew(const HmyDev& st, int i, size_t e) {
    const size_t nD = t5_nD(st);
    if (!MULTI) return __ldcg(&st.Dnew[(size_t)i * nD + e]);
    const uint2* base = reinterpret_cast<const uint2*>(st.xpeer[st.xrank] + st.x5_off_d) + ((size_t)(st.x5_seq & 1u) * st.nblk + i) * HMY_MAX_WORLD * nD + e;
    return t5_ll_sum(base, nD, st.xworld, st.x5_seq);
}
// after the rank's barrier: this CTA's slice of Dnew[blk] goes to every rank
__device__ __forceinline__ void t5_push_block(const HmyDev& st, int blk, unsigned int G) {
    const size_t nD = t5_nD(st);
    const int e0 = (int)((long long)blockIdx.x * (long long)nD / G), e1 = (int)((long long)(blockIdx.x + 1) * (long long)nD / G), len = e1 - e0;
    const size_t slot = (((size_t)(st.x5_seq & 1u) * st.nblk + blk) * HMY_MAX_WORLD + st.xrank) * nD;
    for (int idx = threadIdx.x; idx < len * st.xworld; idx += T5_EPI_THREADS) {
        const int r = idx / len, e = e0 + (idx - r * len);
        const float v = __ldcg(&st.Dnew[(size_t)blk * nD + e]);
        st_volatile_v2(reinterpret_cast<uint2*>(st.xpeer[r] + st.x5_off_d) + slot + e, __float_as_uint(v), st.x5_seq);
    }
}

// ---- epilogue warps ---------------------------------------------------------------------------------------------------
// one-hot rows of slot q's combination: 8 x u16 held in registers by every thread of quarter q (the tile header
// that carried them may be recycled by the producer before a flush reads them)
struct T5Lev { uint4 w; __device__ __forceinline__ int get(int v) const { const unsigned int x = (v < 2) ? w.x : (v < 4) ? w.y : (v < 6) ? w.z : w.w; return (int)((v & 1) ? (x >> 16) : (x & 0xFFFFu)); } };

// The O row of a (slot, cluster) pair as block `blk` sees it (harmony.py:491-492, :506-507): O of the previous stage
// plus what the finished blocks put back (Dnew[i], complete once barrier i + 1 has been passed and never touched again)
// minus what the blocks up to blk took out (Told[i], accumulated by the previous launch).  Every thread of quarter q
// keeps the rows of slot q's levels (and the cluster's row sum) as running values; `upto` is the block they are at.
struct T5Run { float o[HMY_MAX_V]; float rs; int upto; };

// penalty rows {pen, sigma ln pen} of the slots in `need` (harmony.py:495-499); all 512 threads, quarter q = slot q,
// thread-in-quarter = cluster
template <int NC, bool MULTI>
__device__ void t5_penalty_rows(const HmyDev& st, int mode, unsigned char* smem, const T5Lev& lev, T5Run& run,
                                unsigned int need, unsigned int live, int blk) {
    constexpr int KT2 = 16 * NC;
    const int tid = threadIdx.x, j = tid >> 7, k = tid & 127;
    if (((need >> j) & 1u) && k < KT2) {
        float2 out = make_float2(0.f, 0.f);
        if (k < st.K) {
            if (mode == 1) out = make_float2(1.f, 0.f);
            else {
                const size_t nBK = (size_t)st.B * st.K, nD = nBK + st.K;
                if (!((live >> j) & 1u)) {
                    // a slot this CTA has not met in this round: start from the previous stage's O
#pragma unroll
                    for (int v = 0; v < HMY_MAX_V; ++v) run.o[v] = (v < st.V) ? (float)__ldcg(&st.O[(size_t)lev.get(v) * st.K + k]) : 0.f;
                    run.rs = (float)__ldcg(&st.Rsum[k]);
                    run.upto = -1;
                }
                while (run.upto < blk) {
                    const int i0 = run.upto + 1;
                    if (i0 == blk) {
                        // the usual single step: block blk-1's re-added sums in, block blk's removed sums out
                        float a[HMY_MAX_V], r[HMY_MAX_V];
#pragma unroll
                        for (int v = 0; v < HMY_MAX_V; ++v) {
                            const size_t e = (size_t)lev.get(v) * st.K + k;
                            a[v] = (v < st.V && i0 > 0) ? t5_dnew<MULTI>(st, i0 - 1, e) : 0.f;
                            r[v] = (v < st.V) ? st.Told[(size_t)i0 * nD + e] : 0.f;      // read-only in this launch: may sit in L1 (prefetched below)
                        }
                        const float ar = (i0 > 0) ? t5_dnew<MULTI>(st, i0 - 1, nBK + k) : 0.f, rr = st.Told[(size_t)i0 * nD + nBK + k];
#pragma unroll
                        for (int v = 0; v < HMY_MAX_V; ++v) run.o[v] += a[v] - r[v];
                        run.rs += ar - rr;
                        run.upto = blk;
                    } else {
                        // a slot met for the first time in the middle of a round catches up on the finished blocks, four
                        // steps per batch with all loads of a batch in flight (one round trip per batch, not per block:
                        // the whole grid waits for the slowest CTA at the next barrier)
                        const int i1 = min(blk - 1, i0 + 3);
                        float acc[HMY_MAX_V];
#pragma unroll
                        for (int v = 0; v < HMY_MAX_V; ++v) acc[v] = 0.f;
                        float accr = 0.f;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + u;
                            if (i <= i1) {
#pragma unroll
                                for (int v = 0; v < HMY_MAX_V; ++v)
                                    if (v < st.V) {
                                        const size_t e = (size_t)lev.get(v) * st.K + k;
                                        acc[v] += ((i > 0) ? t5_dnew<MULTI>(st, i - 1, e) : 0.f) - __ldcg(&st.Told[(size_t)i * nD + e]);
                                    }
                                accr += ((i > 0) ? t5_dnew<MULTI>(st, i - 1, nBK + k) : 0.f) - __ldcg(&st.Told[(size_t)i * nD + nBK + k]);
                            }
                        }
#pragma unroll
                        for (int v = 0; v < HMY_MAX_V; ++v) run.o[v] += acc[v];
                        run.rs += accr;
                        run.upto = i1;
                    }
                }
                if (blk + 1 < st.nblk) {
                    // the next block's removed sums are known already: pull them into L1 under this block's tiles
#pragma unroll
                    for (int v = 0; v < HMY_MAX_V; ++v)
                        if (v < st.V) asm volatile("prefetch.global.L1 [%0];" ::"l"(&st.Told[(size_t)(blk + 1) * nD + (size_t)lev.get(v) * st.K + k]));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(&st.Told[(size_t)(blk + 1) * nD + nBK + k]));
                }
                float pen = 0.f;
#pragma unroll
                for (int v = 0; v < HMY_MAX_V; ++v)
                    if (v < st.V) {
                        const int b = lev.get(v);
                        const float e = run.rs * __ldg(&st.Pr_b[b]);
                        const float ratio = fminf(fmaxf(e / fmaxf(run.o[v] + e, 1e-8f), 1e-8f), 1.0f);
                        const float th = __ldg(&st.theta[b]);
                        pen += (th == 2.0f) ? ratio * ratio : powf(ratio, th);
                    }
                out.x = pen;
                out.y = (pen > 0.f) ? __ldg(&st.sigma[k]) * logf(pen) : 0.f;
            }
        }
        reinterpret_cast<float2*>(smem + t5_off_ps(NC))[j * KT2 + k] = out;
    }
    t5_bar_sync(1, T5_EPI_THREADS);
}

// per-slot column sums of R since the previous flush -> the block's re-added sums (harmony.py:506-507).  The
// accumulator runs over the whole round (it shares its MMAs with the per-next-block sums): `prev` is what it held at
// the previous flush.
__device__ __forceinline__ void t5_flush_o(const HmyDev& st, unsigned int tmem, const T5Lev& lev, unsigned int mask, int blk, float& prev) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp >> 2, k = 32 * (warp & 3) + lane;
    if ((mask >> q) & 1u) {                                  // warps of quarter q take slot q (lane = cluster)
        const float v = t5_ld1(tmem + ((unsigned int)(32 * (warp & 3)) << 16) + T5_COL_OT + 48u * q);
        t5_ld_wait();
        const float x = (v - prev) * (1.0f / HMY_OPSCALE);
        prev = v;
        if (k < st.K && x != 0.f) {
            float* dn = st.Dnew + (size_t)blk * t5_nD(st);
            for (int vv = 0; vv < st.V; ++vv) atomicAdd(&dn[(size_t)lev.get(vv) * st.K + k], x);
            atomicAdd(&dn[(size_t)st.B * st.K + k], x);          // row sums
        }
    }
}
// per-slot sums by NEXT round's block -> the table that round removes block by block (harmony.py:491-492)
__device__ __forceinline__ void t5_flush_t(const HmyDev& st, unsigned int tmem, const T5Lev& lev, unsigned int mask) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp >> 2, k = 32 * (warp & 3) + lane;
    if ((mask >> q) & 1u) {
        float v[32];
        const unsigned int ta = tmem + ((unsigned int)(32 * (warp & 3)) << 16) + T5_COL_OT + 48u * q + 16u;
        t5_ld16(ta, v); t5_ld16(ta + 16u, v + 16);
        t5_ld_wait();
        if (k < st.K) {
#pragma unroll
            for (int nb = 0; nb < 32; ++nb) {
                const float x = v[nb] * (1.0f / HMY_OPSCALE);
                if (nb < st.nblk && x != 0.f) {
                    float* tn = st.Told_next + (size_t)nb * t5_nD(st);
                    for (int vv = 0; vv < st.V; ++vv) atomicAdd(&tn[(size_t)lev.get(vv) * st.K + k], x);
                    atomicAdd(&tn[(size_t)st.B * st.K + k], x);
                }
            }
        }
    }
}

__device__ __forceinline__ void t5_grid_barrier(const HmyDev& st, unsigned long long target) {
    t5_bar_sync(1, T5_EPI_THREADS);
    if (threadIdx.x == 0) {
        __threadfence();
        red_release_u64(st.bar64, 1ull);
        unsigned int spins = 0; unsigned long long t0 = 0;
        while (ld_acquire_u64(st.bar64) < target) {
            if ((++spins & 1023u) == 0u) {
                unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t; else if (t - t0 > 8000000000ull) __trap();
            }
        }
        __threadfence();
    }
    t5_bar_sync(1, T5_EPI_THREADS);
}

// Timeline stamps (option "trace", thread 0): slot 0 start, 1 prologue done, per block b: 2 + 5 b + {0 first tile
// staged, 1 penalty rows ready, 2 last tile handed to the accumulation, 3 block sums flushed, 4 grid barrier passed};
// per tile i of block 5: 102 + 6 i + {0 scores ready, 1 scores in registers, 2 pass 1 done, 3 row sums met,
// 4 operand tile free, 5 operand tile written}
#define T5_STAMP(slot_) hmy_trace(st, (slot_))
// The slowest tile of every CTA (thread 0, all traced launches): slots 166 + {0 tile described, 1 scores ready, 2 scores in
// registers, 3 pass 1 done, 4 row sums met, 5 operand tile free} are the running tile's stamps; when a tile ends later
// after its predecessor (or after the block's penalty rows) than any before, they are copied to 176.., with the gap in 174,
// block << 8 | tile in 175, the end in 182 and the predecessor's end in 183.
#define T5_WT(i_) hmy_trace(st, 166 + (i_))
__device__ __forceinline__ void t5_trace_tile_end(const HmyDev& st, int blk, int tile, bool restart) {
    if (st.trace != nullptr && threadIdx.x == 0) {
        volatile unsigned long long* T = st.trace + (size_t)blockIdx.x * HMY_TRACE_SLOTS;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        const unsigned long long prev = T[173];
        if (!restart && prev != 0 && now - prev > T[174]) {
            T[174] = now - prev; T[175] = ((unsigned long long)blk << 8) | (unsigned long long)tile;
#pragma unroll
            for (int i = 0; i < 6; ++i) T[176 + i] = T[166 + i];
            T[182] = now; T[183] = prev;
        }
        T[173] = now;
    }
}
template <int NC, bool MULTI>
__device__ void t5_epilogue(const HmyDev& st, int mode, unsigned char* smem, unsigned int bar0, unsigned int tmem,
                            unsigned int G, unsigned long long bar_base) {
    constexpr int KT2 = 16 * NC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp >> 2, rq = warp & 3, row = 32 * rq + lane;
    const int g0 = t5_g0<NC>(q), ng = t5_ng<NC>(q), col0 = 8 * g0;
    const unsigned int lane_base = (unsigned int)(32 * rq) << 16;
    const float2* cK = reinterpret_cast<const float2*>(smem + t5_off_ck(NC));
    const float2* pS = reinterpret_cast<const float2*>(smem + t5_off_ps(NC));
    float2* xch = reinterpret_cast<float2*>(smem + t5_off_xch(NC));
    unsigned int bmask = 0u, rmask = 0u, computed = 0u;
    double objd = 0.0, obje = 0.0;
    unsigned int t = 0, nblocks_done = 0;
    unsigned long long bar_next = bar_base;
    T5Lev lev; lev.w = make_uint4(0u, 0u, 0u, 0u);
    T5Run run; run.rs = 0.f; run.upto = -1;
#pragma unroll
    for (int v = 0; v < HMY_MAX_V; ++v) run.o[v] = 0.f;
    unsigned int live = 0u;                       // slots whose running rows are being followed
    bool had_cells = false;                       // D2y holds sums (tensor memory is not cleared by the allocation)
    float o_prev = 0.f;                           // slot q's running column sum (this thread's cluster) at the last flush
    bool first_of_block = true; int tile_in_block = 0;
    T5_STAMP(1);
    for (;;) {
        const unsigned int s = t % T5_NZ, d = t & 1u;
        t5_wait(bar0 + 8u * (T5_B_ZFULL + s), (t / T5_NZ) & 1u);
        T5_WT(0);
        const unsigned char* meta = smem + t5_off_meta(NC) + s * 1024;
        const int* hdr = reinterpret_cast<const int*>(meta + T5_META_HDR);
        const unsigned int flags = (unsigned int)hdr[T5_H_FLAGS], tmask = (unsigned int)hdr[T5_H_MASK];
        const int blk = hdr[T5_H_BLK];
        const bool empty = (flags & T5_F_EMPTY) != 0u;
        bool waited_acc = false;
        if (first_of_block) { T5_STAMP(2 + 5 * blk); }
        const int tslot = (blk == 5 && tile_in_block < 4) ? 102 + 6 * tile_in_block : HMY_TRACE_SLOTS;
        if (flags & T5_F_EVICT) {
            // the slot table starts over with this tile: everything accumulated under the old table leaves now
            if (t > 0) { t5_wait(bar0 + 8u * (T5_B_ACC + (t - 1u) % T5_NZ), ((t - 1u) / T5_NZ) & 1u); t5_fence_after(); }
            waited_acc = true;
            t5_flush_o(st, tmem, lev, bmask, blk, o_prev);
            t5_flush_t(st, tmem, lev, rmask);
            t5_fence_before();
            bmask = 0u; rmask = 0u; computed = 0u; live = 0u; o_prev = 0.f;
        }
        lev.w = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(hdr + T5_H_LEV) + 16 * q);
        {
            // the first tile of a block brings the rows of ALL slots with a combination up to date (slot q = quarter q:
            // it costs no more time than one slot, and no slot ever falls behind); later tiles only add new slots
            const unsigned int want = (first_of_block ? (unsigned int)hdr[T5_H_ALLOC] : tmask) & ~computed;
            if (want) {
                t5_penalty_rows<NC, MULTI>(st, mode, smem, lev, run, want, live, blk);
                computed |= want; live |= want;
            }
        }
        if (first_of_block) { T5_STAMP(3 + 5 * blk); t5_trace_tile_end(st, blk, 0, true); first_of_block = false; }
        const unsigned int slotnb = reinterpret_cast<const unsigned short*>(meta + T5_META_SLOTNB)[row];
        const bool valid = slotnb != 0xFFFFu;
        const int slot = valid ? (int)(slotnb >> 8) : 0, nb = (int)(slotnb & 255u);
        const int cell = reinterpret_cast<const int*>(meta)[row];
        float E[32];
        // ---- scores of this thread's clusters (an empty tile still hands the accumulator back)
        t5_wait(bar0 + 8u * (T5_B_SFULL + d), (t >> 1) & 1u);
        t5_fence_after();
        T5_STAMP(tslot);
        T5_WT(1);
        if (!empty) {
            const unsigned int ta = tmem + lane_base + T5_COL_D1 + 128u * d + (unsigned int)col0;
            t5_ld16(ta, E);
            if (ng == 4) t5_ld16(ta + 16u, E + 16);
            else if (ng == 3) t5_ld8(ta + 16u, E + 16);
            t5_ld_wait();
        }
        t5_fence_before();
        __syncwarp();
        if (lane == 0) t5_arrive(bar0 + 8u * (T5_B_SFREE + d));
        T5_STAMP(tslot + 1);
        T5_WT(2);
        if (!empty) {
            // ---- S = exp(-dist / sigma) (harmony.py:466-467) times the penalty (harmony.py:500)
            float ss = 0.f, sp = 0.f, sd = 0.f, se = 0.f, sg = 0.f;
            const float2* pr = pS + slot * KT2 + col0;
            if (col0 + 8 * ng > st.K) {
                // padding clusters (this quarter's tail): a score of -1e30 makes dist huge and S exactly 0
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j < 8 * ng && col0 + j >= st.K) E[j] = -1.0e30f;
            }
            if (st.dbg & 4) { ss = 1.f; sp = 1.f; }
            else if (st.sigma_uniform) {
                const float c1u = -1.4426950408889634f / st.sigma_u;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < ng) {
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const float4 pp = *reinterpret_cast<const float4*>(pr + 8 * g + 2 * e2);
                            {
                                const float dist = fmaf(E[8 * g + 2 * e2], -1.9073486328125e-6f, 2.0f);      // 2 (1 - z.y); scores carry 2^20
                                const float sv = ex2_approx(dist * c1u);
                                const float ev = sv * pp.x;
                                ss += sv; sp += ev;
                                sd = fmaf(ev, dist, sd); se = fmaf(ev, pp.y, se);
                                E[8 * g + 2 * e2] = ev;
                            }
                            {
                                const float dist = fmaf(E[8 * g + 2 * e2 + 1], -1.9073486328125e-6f, 2.0f);
                                const float sv = ex2_approx(dist * c1u);
                                const float ev = sv * pp.z;
                                ss += sv; sp += ev;
                                sd = fmaf(ev, dist, sd); se = fmaf(ev, pp.w, se);
                                E[8 * g + 2 * e2 + 1] = ev;
                            }
                        }
                    }
                }
                sg = st.sigma_u * sp;                   // sum_k sigma_k ev_k
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g < ng) {
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const float4 ck = *reinterpret_cast<const float4*>(cK + col0 + 8 * g + 2 * e2);
                            const float4 pp = *reinterpret_cast<const float4*>(pr + 8 * g + 2 * e2);
                            {
                                const float dist = fmaf(E[8 * g + 2 * e2], -1.9073486328125e-6f, 2.0f);
                                const float sv = ex2_approx(dist * ck.x);
                                const float ev = sv * pp.x;
                                ss += sv; sp += ev;
                                sd = fmaf(ev, dist, sd); se = fmaf(ev, pp.y, se); sg = fmaf(ev, ck.y, sg);
                                E[8 * g + 2 * e2] = ev;
                            }
                            {
                                const float dist = fmaf(E[8 * g + 2 * e2 + 1], -1.9073486328125e-6f, 2.0f);
                                const float sv = ex2_approx(dist * ck.z);
                                const float ev = sv * pp.z;
                                ss += sv; sp += ev;
                                sd = fmaf(ev, dist, sd); se = fmaf(ev, pp.w, se); sg = fmaf(ev, ck.w, sg);
                                E[8 * g + 2 * e2 + 1] = ev;
                            }
                        }
                    }
                }
            }
            // ---- the four threads of a cell meet: sums over all clusters
            float2* xs = xch + (t & 1u) * 512;
            T5_STAMP(tslot + 2);
            T5_WT(3);
            xs[q * 128 + row] = make_float2(ss, sp);
            t5_bar_sync(2 + rq, 128);
            T5_STAMP(tslot + 3);
            T5_WT(4);
            float sst = 0.f, spt = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) { const float2 u = xs[qq * 128 + row]; sst += u.x; spt += u.y; }
            // R = (S / sum S) pen / max(sum (S / sum S) pen, 1e-8)   (harmony.py:468, :500-503)
            const float is = 1.f / sst;
            const float sc = valid ? is / fmaxf(spt * is, 1e-8f) : 0.f;
            if (valid) {
                // sum_k r dist and sum_k sigma r ln r with ln r = -dist / sigma + ln pen + ln sc (harmony.py:399-402)
                objd += (double)(sc * sd);
                obje += (double)(sc * (se - sd + (lg2_approx(sc) * 0.6931471805599453f) * sg));
            }
            // ---- operand tiles of the accumulation: wait until the previous tile's MMAs have read them
            if (t > 0 && !waited_acc) t5_wait(bar0 + 8u * (T5_B_ACC + (t - 1u) % T5_NZ), ((t - 1u) / T5_NZ) & 1u);
            T5_STAMP(tslot + 4);
            T5_WT(5);
            const float sc1024 = sc * HMY_OPSCALE;
            unsigned char* Rh = smem + T5_OFF_RH + (row >> 3) * T5_R_LBO + (row & 7) * 16;
            float* Rg = st.R + (size_t)cell * st.Kp;
            const bool wr = st.write_R && valid;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g < ng) {
                    uint4 hi, lo;
                    split2(E[8 * g] * sc1024, E[8 * g + 1] * sc1024, hi.x, lo.x);
                    split2(E[8 * g + 2] * sc1024, E[8 * g + 3] * sc1024, hi.y, lo.y);
                    split2(E[8 * g + 4] * sc1024, E[8 * g + 5] * sc1024, hi.z, lo.z);
                    split2(E[8 * g + 6] * sc1024, E[8 * g + 7] * sc1024, hi.w, lo.w);
                    if (!(st.dbg & 8)) {
                        *reinterpret_cast<uint4*>(Rh + (g0 + g) * T5_R_SBO) = hi;
                        *reinterpret_cast<uint4*>(Rh + T5_SZ_RPART + (g0 + g) * T5_R_SBO) = lo;
                    }
                    if (wr) {
                        const int c = col0 + 8 * g;
                        if (c < st.Kp) *reinterpret_cast<float4*>(Rg + c) = make_float4(E[8 * g] * sc, E[8 * g + 1] * sc, E[8 * g + 2] * sc, E[8 * g + 3] * sc);
                        if (c + 4 < st.Kp) *reinterpret_cast<float4*>(Rg + c + 4) = make_float4(E[8 * g + 4] * sc, E[8 * g + 5] * sc, E[8 * g + 6] * sc, E[8 * g + 7] * sc);
                    }
                }
            }
            {
                // one-hot of the cell's block in the next round: quarter q writes columns [8 q, 8 q + 8)
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (valid && (nb >> 3) == q) {
                    const unsigned int h = (nb & 1) ? 0x3C000000u : 0x3C00u;
                    const int p = (nb & 7) >> 1;
                    w.x = p == 0 ? h : 0u; w.y = p == 1 ? h : 0u; w.z = p == 2 ? h : 0u; w.w = p == 3 ? h : 0u;
                }
                *reinterpret_cast<uint4*>(smem + T5_OFF_NB + (2 + q) * T5_R_SBO + (row >> 3) * T5_R_LBO + (row & 7) * 16) = w;
            }
            bmask |= tmask; rmask |= tmask; had_cells = true;
        } else {
            if (t > 0 && !waited_acc) t5_wait(bar0 + 8u * (T5_B_ACC + (t - 1u) % T5_NZ), ((t - 1u) / T5_NZ) & 1u);
        }
        t5_fence_async();
        __syncwarp();
        if (lane == 0) t5_arrive(bar0 + 8u * T5_B_RFULL);
        T5_STAMP(tslot + 5);
        t5_trace_tile_end(st, blk, tile_in_block, false);
        ++t; ++tile_in_block;
        if (flags & T5_F_LAST_BLOCK) {
            T5_STAMP(4 + 5 * blk);
            first_of_block = true; tile_in_block = 0;
            // ---- end of the block: its sums join the running O, the next block's removed sums leave it
            t5_wait(bar0 + 8u * T5_B_ODONE, nblocks_done & 1u);
            t5_fence_after();
            ++nblocks_done;
            t5_flush_o(st, tmem, lev, bmask, blk, o_prev);
            bmask = 0u; computed = 0u;
            T5_STAMP(5 + 5 * blk);
            const bool last = (flags & T5_F_LAST_ROUND) != 0u;
            if (!last) {
                t5_fence_before();
                t5_grid_barrier(st, bar_next += G);
                if (MULTI) t5_push_block(st, blk, G);          // this rank's part of the block's sums goes to every rank
                T5_STAMP(6 + 5 * blk);
            } else {
                // ---- end of the round: centroid sums, next round's removed sums, objective sums
                t5_wait(bar0 + 8u * (T5_B_ACC + (t - 1u) % T5_NZ), ((t - 1u) / T5_NZ) & 1u);
                t5_fence_after();
                t5_flush_t(st, tmem, lev, rmask);
                const int dt = (st.d + 15) >> 4;
                if (q < dt && had_cells) {
                    float v[16];
                    t5_ld16(tmem + lane_base + T5_COL_Y + 16u * q, v);
                    t5_ld_wait();
                    if (row < st.K) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int jj = 16 * q + j;
                            if (jj < st.d && v[j] != 0.f) atomicAdd(&st.Yacc[(size_t)row * st.dp + jj], (double)v[j] * (double)HMY_ACCSCALE);
                        }
                    }
                }
                const double a = warp_sum_d(objd), b = warp_sum_d(obje);
                if (lane == 0) { atomicAdd(&st.obj[0], a); atomicAdd(&st.obj[1], b); }
                t5_fence_before();
                t5_grid_barrier(st, bar_next += G);
                T5_STAMP(6 + 5 * blk);
                if (MULTI) {
                    // ---- this rank's sums are complete: its slices of every small table go to every rank
                    t5_push_block(st, blk, G);
                    const unsigned int seq = st.x5_seq, par = seq & 1u;
                    const int W = st.xworld, me = st.xrank;
                    const size_t nD = t5_nD(st), nT = (size_t)st.nblk * nD, nY = (size_t)st.K * st.dp + 2;
                    {
                        const long long e0 = (long long)blockIdx.x * (long long)nT / G, e1 = (long long)(blockIdx.x + 1) * (long long)nT / G, len = e1 - e0;
                        for (long long idx = tid; idx < len * W; idx += T5_EPI_THREADS) {
                            const int r = (int)(idx / len); const size_t e = (size_t)(e0 + (idx - (long long)r * len));
                            st_volatile_v2(reinterpret_cast<uint2*>(st.xpeer[r] + st.x5_off_t) + ((size_t)par * HMY_MAX_WORLD + me) * nT + e,
                                           __float_as_uint(__ldcg(&st.Told_next[e])), seq);
                        }
                    }
                    {
                        const long long e0 = (long long)blockIdx.x * (long long)nY / G, e1 = (long long)(blockIdx.x + 1) * (long long)nY / G, len = e1 - e0;
                        for (long long idx = tid; idx < len * W; idx += T5_EPI_THREADS) {
                            const int r = (int)(idx / len); const size_t e = (size_t)(e0 + (idx - (long long)r * len));
                            const double v = (e < nY - 2) ? __ldcg(&st.Yacc[e]) : __ldcg(&st.obj[e - (nY - 2)]);
                            const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                            uint2* dst = reinterpret_cast<uint2*>(st.xpeer[r] + st.x5_off_y) + ((size_t)par * HMY_MAX_WORLD + me) * 2 * nY + 2 * e;
                            st_volatile_v2(dst, (unsigned int)bits, seq);
                            st_volatile_v2(dst + 1, (unsigned int)(bits >> 32), seq);
                        }
                    }
                    {
                        // the next round's removed sums, summed over ranks (this CTA's slice; rank order: identical everywhere)
                        const long long e0 = (long long)blockIdx.x * (long long)nT / G, e1 = (long long)(blockIdx.x + 1) * (long long)nT / G;
                        const uint2* src = reinterpret_cast<const uint2*>(st.xpeer[me] + st.x5_off_t) + (size_t)par * HMY_MAX_WORLD * nT;
                        for (long long e = e0 + tid; e < e1; e += T5_EPI_THREADS) st.Told_next[e] = t5_ll_sum(src + (size_t)e, nT, W, seq);
                    }
                }
                break;
            }
        }
    }
    // ---- after the last barrier: every CTA finishes its slice of the small tables
    {
        // O of the finished stage = what its blocks put back (harmony.py:506-507 over all blocks; drift-free);
        // cross-entropy term of the objective (harmony.py:404-411 collapsed to K x B)
        const int nblocks = (mode == 1) ? 1 : st.nblk;
        const size_t nBK = (size_t)st.B * st.K;
        const int n = (int)nBK;
        const int e0 = (int)((long long)blockIdx.x * n / G), e1 = (int)((long long)(blockIdx.x + 1) * n / G);
        double part = 0.0;
        // one warp per element, lane = block: the per-block terms (with several GPUs: W polled packets each) are read
        // side by side instead of one after the other
        for (int e = e0 + warp; e < e1; e += T5_EPI_WARPS) {
            const int b = e / st.K, k = e - b * st.K;
            double o = (lane < nblocks) ? (double)t5_dnew<MULTI>(st, lane, (size_t)e) : 0.0;
            double rs = (lane < nblocks) ? (double)t5_dnew<MULTI>(st, lane, nBK + k) : 0.0;
            o = warp_sum_d(o); rs = warp_sum_d(rs);
            if (lane == 0) {
                st.O[e] = o;
                const float oc = fmaxf((float)o, 1e-8f), ec = fmaxf((float)(rs * (double)st.Pr_b[b]), 1e-8f);
                part += (double)st.sigma[k] * (double)st.theta[b] * (double)logf((oc + ec) / ec) * o;
            }
        }
        if (!MULTI) {
            part = warp_sum_d(part);
            if (lane == 0 && part != 0.0) atomicAdd(&st.obj[2], part);
        } else if (part != 0.0) {
            // every rank computes this term from the same tables; summed in 2^-30 fixed point so that the order of
            // the atomics cannot make the ranks' objectives (and with them their convergence decisions) differ
            atomicAdd(reinterpret_cast<unsigned long long*>(&st.obj[3]), (unsigned long long)__double2ll_rn(part * 1073741824.0));
        }
        // row sums the next round starts from, unit centroids of the next round (harmony.py:443-444)
        for (int k = blockIdx.x * T5_EPI_WARPS + warp; k < st.K; k += G * T5_EPI_WARPS) {
            double rs = (lane < nblocks) ? (double)t5_dnew<MULTI>(st, lane, nBK + k) : 0.0;
            rs = warp_sum_d(rs);
            if (lane == 0) st.Rsum_next[k] = rs;
            double y0 = 0.0, y1 = 0.0;
            if (!MULTI) {
                y0 = (lane < st.d) ? __ldcg(&st.Yacc[(size_t)k * st.dp + lane]) : 0.0;
                y1 = (lane + 32 < st.d) ? __ldcg(&st.Yacc[(size_t)k * st.dp + lane + 32]) : 0.0;
            } else {
                const size_t nY = (size_t)st.K * st.dp + 2;
                const uint2* src = reinterpret_cast<const uint2*>(st.xpeer[st.xrank] + st.x5_off_y) + (size_t)(st.x5_seq & 1u) * HMY_MAX_WORLD * 2 * nY;
                if (lane < st.d) y0 = t5_ll_sum_d(src + 2 * ((size_t)k * st.dp + lane), 2 * nY, st.xworld, st.x5_seq);
                if (lane + 32 < st.d) y1 = t5_ll_sum_d(src + 2 * ((size_t)k * st.dp + lane + 32), 2 * nY, st.xworld, st.x5_seq);
            }
            const double inv = 1.0 / sqrt(warp_sum_d(y0 * y0 + y1 * y1));
            if (lane < st.dp) st.Ynext[(size_t)k * st.dp + lane] = (lane < st.d) ? (float)(y0 * inv) : 0.f;
            if (lane + 32 < st.dp) st.Ynext[(size_t)k * st.dp + lane + 32] = (lane + 32 < st.d) ? (float)(y1 * inv) : 0.f;
        }
        if (MULTI && blockIdx.x == 0 && tid < 2) {
            // objective sums over all ranks (rank order) for the host
            const size_t nY = (size_t)st.K * st.dp + 2;
            const uint2* src = reinterpret_cast<const uint2*>(st.xpeer[st.xrank] + st.x5_off_y) + (size_t)(st.x5_seq & 1u) * HMY_MAX_WORLD * 2 * nY;
            const double sum = t5_ll_sum_d(src + 2 * (nY - 2 + tid), 2 * nY, st.xworld, st.x5_seq);
            st.obj_out[tid] = sum;
        }
    }
}

// ---- kernel -----------------------------------------------------------------------------------------------------------
// mode 0: one k-means round; mode 1: the assignment of init_cluster (harmony.py:377-392: no penalty, all cells,
// storage order).  bar_base: value of the grid-barrier counter when the launch starts.
template <int NC, bool MULTI>
__global__ void __launch_bounds__(T5_THREADS, 1) k_round_tc5(HmyDev st, int mode, unsigned long long bar_base) {
    extern __shared__ __align__(1024) unsigned char smem_t5[];
    unsigned char* const smem = smem_t5;
    constexpr int KT2 = 16 * NC;
    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: the compiler then knows the role dispatch below is warp-uniform (without that the
    // MMA warp's code counts as divergent and every descriptor is moved into uniform registers again for every MMA)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const unsigned int G = gridDim.x;
    const unsigned int bar0 = smem_u32(smem + t5_off_bar(NC));
    unsigned int* s_tmem = reinterpret_cast<unsigned int*>(smem + t5_off_bar(NC) + 128);
    if (warp == T5_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * (T5_B_ZFULL + i)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * (T5_B_ACC + i)));
        }
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * (T5_B_SFULL + i)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (T5_B_SFREE + i)), "r"(T5_EPI_WARPS));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * T5_B_RFULL), "r"(T5_EPI_WARPS));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * T5_B_ODONE));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- round constants: centroids as fp16 hi/lo (K-major core matrices), per-cluster constants, zeroed tiles
    for (int i = tid; i < KT2 * 32; i += T5_THREADS) {
        const int k = i >> 5, j = 2 * (i & 31);
        float y0 = 0.f, y1 = 0.f;
        if (k < st.K) {
            if (j < st.dp) y0 = st.Yhat[(size_t)k * st.dp + j] * HMY_OPSCALE;
            if (j + 1 < st.dp) y1 = st.Yhat[(size_t)k * st.dp + j + 1] * HMY_OPSCALE;
        }
        unsigned int hi, lo;
        split2(y0, y1, hi, lo);
        const int off = (k >> 3) * T5_SBO_K + (j >> 3) * T5_LBO_K + (k & 7) * 16 + (j & 7) * 2;
        *reinterpret_cast<unsigned int*>(smem + T5_OFF_YH + off) = hi;
        *reinterpret_cast<unsigned int*>(smem + t5_off_yl(NC) + off) = lo;
    }
    for (int k = tid; k < KT2; k += T5_THREADS) {
        // clusters beyond K: exp2(dist * -1e4) = 0 keeps them out of every sum
        const float sg = (k < st.K) ? st.sigma[k] : 0.f;
        reinterpret_cast<float2*>(smem + t5_off_ck(NC))[k] = make_float2((k < st.K) ? -1.4426950408889634f / sg : -1.0e4f, sg);
    }
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < T5_OFF_YH / 16; i += T5_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);          // Z stages | R tiles | one-hot tile
        __syncthreads();
        unsigned int* ones = reinterpret_cast<unsigned int*>(smem + T5_OFF_NB);
        for (int i = tid; i < 2 * 2048 / 4; i += T5_THREADS) ones[i] = 0x3C003C00u;                         // 16 columns of 1.0
        float2* ps = reinterpret_cast<float2*>(smem + t5_off_ps(NC));
        for (int i = tid; i < T5_SLOTS * KT2; i += T5_THREADS) ps[i] = make_float2(0.f, 0.f);
    }
    t5_fence_async();
    t5_fence_before();
    __syncthreads();
    t5_fence_after();
    const unsigned int tmem = *s_tmem;
    hmy_trace(st, 0);
    if (st.trace != nullptr && lane == 0 && (warp == T5_WARP_PROD || warp == T5_WARP_MMA))       // gaps do not span launches
        st.trace[(size_t)blockIdx.x * HMY_TRACE_SLOTS + (warp == T5_WARP_PROD ? 184 : 192)] = 0ull;
    if (warp == T5_WARP_PROD) t5_producer<NC>(st, mode, smem, bar0, G);
    else if (warp == T5_WARP_MMA) t5_mma_warp<NC>(st, smem, bar0, tmem);
    else t5_epilogue<NC, MULTI>(st, mode, smem, bar0, tmem, G, bar_base);
    t5_fence_before();
    __syncthreads();
    if (warp == T5_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
