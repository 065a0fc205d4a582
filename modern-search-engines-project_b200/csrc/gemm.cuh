// K4 — exhaustive dense scan for large query batches on the 5th-gen tensor cores (tcgen05).
//
// S^T[B x rows] = Q[B x 768] * E^T, bf16 operands, fp32 accumulators in TMEM; the epilogue is the
// same per-document max + running-bound candidate emission as the GEMV kernel (dense.cuh), so the
// score matrix never exists in HBM.
//
// Roles of the operands: the QUERIES are the M dimension (TMEM lanes), the chunk rows the N dimension (TMEM
// columns).  A tcgen05.ld (32x32b) therefore hands every epilogue thread 32 consecutive chunk scores of ITS
// query: the per-document max is a running max over registers with warp-uniform document boundaries — no
// shuffles — and the query's running bound is one register.
//
// The query panel lives in TENSOR MEMORY for the whole kernel (tcgen05.mma with the A operand in TMEM): 128 queries
// x 768 bf16 = 384 of the 512 TMEM columns (lane = query, two K elements per 32-bit column), written once with
// tcgen05.st.  Only the chunk rows stream through shared memory, so the whole 200 KB ring buffers E tiles (24 stages
// of 8 KB: ~190 KB of TMA loads in flight per SM) and the L2->SM path carries E only.  (Round 1 kept Q in shared
// memory and re-fetched the 384 KB panel from L2 for every 128-row tile: 46 GB through L2->SM per 15.4 GB of E,
// which pinned the kernel at the ~6300 B/cycle L2 limit, 4.7-5.0 ms at B=256.)
// A CTA owns ONE panel of <= 128 queries; a batch of 129..256 queries runs as two interleaved sets of CTAs (even /
// odd blockIdx) that walk the same tiles in the same order, so the second reader of an E tile hits L2.
//
// Tiling: UMMA M=128 queries x N=64 chunk rows x K=16, 12 k-blocks of 64 elements (one 128-byte SWIZZLE_128B row
// per k-block), accumulators 2 x 64 TMEM columns (double-buffered: the epilogue of tile i overlaps the MMAs of
// tile i+1).  A 64-row tile is assembled from TWO doc-aligned groups of <= 32 rows (two {64 x 32} TMA boxes per
// k-block, each starting at the first row of a document), so a 32-column tcgen05.ld covers whole documents only.
// (Rows past the end of a group belong to the next group; they are computed twice and masked, ~8 % redundant L2
// reads.)  Documents with more than 32 chunks do not fit a group: the caller falls back to the GEMV kernel.
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2-9 = epilogue (warp w reads TMEM lane quarter w%4 =
// queries 32*(w%4).. of the panel; the two warps of a quarter take one group each).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dense.cuh"
#include "topk.cuh"

namespace mse {

constexpr int kGemmEpiWarps = 8;                      // two per TMEM lane quarter, one doc-aligned group each
constexpr int kGemmThreads = (2 + kGemmEpiWarps) * 32;
constexpr int kGemmBlockK = 64;                        // elements per k-block (128 B of bf16)
constexpr int kGemmKBlocks = kDim / kGemmBlockK;       // 12
constexpr int kGemmGroupRows = 32;
constexpr int kGemmTileGroups = 2;
constexpr int kGemmTileRows = kGemmTileGroups * kGemmGroupRows;   // 64 chunk rows = N of the MMA
constexpr int kGemmStageBytes = kGemmTileRows * 128;   // 8 KB per stage (one k-block of a tile)
constexpr int kGemmMaxStages = 24;
constexpr int kGemmPanel = 128;                        // queries per CTA
constexpr int kGemmACols = kDim / 2;                   // TMEM columns of the query panel (2 bf16 per column)
constexpr int kGemmStage = 96;                         // emissions staged per epilogue warp between flushes

struct GemmWork {
    const int64_t* group_row;    // [n_groups + 1] first row of every doc-aligned group (<= 32 rows each)
    int64_t n_groups;
    int64_t n_tiles;             // ceil(n_groups / kGemmTileGroups)
    const __nv_bfloat16* qb16;   // [n_panels * 128][768] bf16 queries, zero padded
    int32_t n_panels;            // 1 or 2 panels of 128 queries
    int32_t n_real;              // real queries in this launch
    int32_t q0;                  // first query (index into cand / tau arrays)
    int32_t stages;
    int32_t debug;               // bit0: skip epilogue math, bit1: skip MMA issue (timing experiments only)
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarrier_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarrier_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbarrier_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbarrier_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; !mbarrier_try_wait(bar, parity); ++it)
        if (it > (1u << 26)) __trap();                 // a lost completion becomes an error, not a hang
}
// for the single-lane producer / MMA roles: back off so the spin does not steal issue slots from the
// epilogue warps that share the scheduler
__device__ __forceinline__ void mbarrier_wait_backoff(uint64_t* bar, uint32_t parity) {
    for (uint32_t it = 0; !mbarrier_try_wait(bar, parity); ++it) {
        __nanosleep(64);
        if (it > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand in tensor memory (lane = row of A, two 16-bit K elements per 32-bit column)
__device__ __forceinline__ void tcgen05_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// K-major, SWIZZLE_128B, 128-byte rows: LBO = 1 (16 B units, unused for swizzled K-major), SBO = 1024 B
// (8-row atom), descriptor version 1 (sm_100), layout type 2.  (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_byte_addr) {
    return uint64_t((smem_byte_addr >> 4) & 0x3fffu) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
           (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__device__ __forceinline__ uint32_t umma_idesc_bf16_f32(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_e, DenseDev dx, DenseWork w, GemmWork g) {
    extern __shared__ __align__(1024) unsigned char gemm_smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kGemmMaxStages], s_empty[kGemmMaxStages], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ uint64_t s_stage_key[kGemmEpiWarps][kGemmStage];
    __shared__ uint16_t s_stage_q[kGemmEpiWarps][kGemmStage];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr uint32_t tmem_cols = 512;                                   // 384 query panel + 2 x 64 accumulator
    // CTAs of panel p are blockIdx p, p + n_panels, ...: both panels walk the same tiles in the same order
    const int panel = int(blockIdx.x) % g.n_panels;
    const int64_t cta = int64_t(blockIdx.x) / g.n_panels, n_cta = int64_t(gridDim.x) / g.n_panels;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbarrier_init(&s_full[s], 1); mbarrier_init(&s_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbarrier_init(&s_tfull[b], 1); mbarrier_init(&s_tempty[b], kGemmEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t tmem_acc = tmem_base + uint32_t(kGemmACols);

    if (warp >= 2) {
        // query panel -> TMEM: thread = one query (TMEM lane), 32 columns (64 K elements) per tcgen05.st; the two warps
        // of a lane quarter write one half of the columns each
        const int quarter = warp & 3;
        const int qrow = panel * kGemmPanel + quarter * 32 + lane;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g.qb16) + int64_t(qrow) * kGemmACols;
        const int c_lo = ((warp - 2) >> 2) * (kGemmACols / 2);
        for (int c0 = c_lo; c0 < c_lo + kGemmACols / 2; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(src + c0 + 4 * i);
                r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
            }
            tmem_st_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c0), r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = cta; tile < g.n_tiles; tile += n_cta) {
                int rows[kGemmTileGroups];
#pragma unroll
                for (int gi = 0; gi < kGemmTileGroups; ++gi) {
                    const int64_t grp = tile * kGemmTileGroups + gi;
                    rows[gi] = grp < g.n_groups ? int(g.group_row[grp]) : int(dx.n_chunks);     // past the end -> zero fill
                }
                for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                    mbarrier_wait_backoff(&s_empty[stage], phase ^ 1u);
                    unsigned char* sa = smem + size_t(stage) * kGemmStageBytes;
                    mbarrier_expect_tx(&s_full[stage], uint32_t(kGemmStageBytes));
#pragma unroll
                    for (int gi = 0; gi < kGemmTileGroups; ++gi)
                        tma_load_2d(sa + gi * (kGemmGroupRows * 128), &map_e, kb * kGemmBlockK, rows[gi], &s_full[stage]);
                    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_f32(kGemmPanel, kGemmTileRows);   // M = 128 queries, N = 64 chunk rows
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int64_t tile = cta; tile < g.n_tiles; tile += n_cta, ++it) {
                const int buf = it & 1;
                const uint32_t use = uint32_t(it >> 1);
                mbarrier_wait_backoff(&s_tempty[buf], (use & 1u) ^ 1u);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_acc + uint32_t(buf) * kGemmTileRows;
                for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                    mbarrier_wait(&s_full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t e_addr = smem_addr(smem + size_t(stage) * kGemmStageBytes);   // 64 chunk rows x 64
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        if (g.debug & 2) break;
                        const uint64_t bd = umma_desc_sw128(e_addr + k * 32);
                        // A: 16 K elements = 8 TMEM columns per MMA
                        tcgen05_mma_bf16_ts(tmem_d, tmem_base + uint32_t((kb * (kGemmBlockK / 16) + k) * 8), bd, idesc, (kb | k) ? 1u : 0u);
                    }
                    tcgen05_commit(&s_empty[stage]);                      // frees the smem slot when the MMAs retire
                    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
                }
                tcgen05_commit(&s_tfull[buf]);                            // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue (warp w -> TMEM lane quarter w % 4 = 32 queries of the panel) ==========
        const int quarter = warp & 3;
        const int ew = warp - 2;                                          // 0..7
        const int gi = ew >> 2;                                           // this warp's doc-aligned group of the tile
        const unsigned lt_mask = (1u << lane) - 1u;
        uint64_t* st_key = s_stage_key[ew];
        uint16_t* st_q = s_stage_q[ew];
        int staged = 0;                                                   // uniform
        int rr = int((blockIdx.x * kGemmEpiWarps + ew) % g.n_real);       // round-robin cursor for bound refreshes
        const int my_q = panel * kGemmPanel + quarter * 32 + lane;        // this thread's query (may be padding)
        const bool q_real = my_q < g.n_real;
        const bool warp_idle = panel * kGemmPanel + quarter * 32 >= g.n_real;   // no real query on these lanes
        auto flush = [&]() {
            if (staged == 0) return;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(w.log_count, (unsigned long long)staged);
            base = __shfl_sync(0xffffffffu, base, 0);
            __syncwarp();
            for (int e = lane; e < staged; e += 32) {
                const int64_t pos = int64_t(base) + e;
                if (pos < w.log_cap) { w.log_key[pos] = st_key[e]; w.log_q[pos] = st_q[e]; }
            }
            __syncwarp();
            staged = 0;
        };
        int it = 0;
        for (int64_t tile = cta; tile < g.n_tiles; tile += n_cta, ++it) {
            const int buf = it & 1;
            const uint32_t use = uint32_t(it >> 1);
            if (warp_idle || (g.debug & 1)) {
                mbarrier_wait_backoff(&s_tfull[buf], use & 1u);
                if (lane == 0) mbarrier_arrive(&s_tempty[buf]);
                continue;
            }
            // this thread's running bound (padding lanes never pass)
            float tau = INFINITY;
            if (q_real) {
                const uint32_t tk = w.use_tau ? ld_relaxed_u32(&w.ts.tau[g.q0 + my_q]) : 0u;
                tau = tk ? key_to_float(tk) : -INFINITY;
            }
            // document layout of the warp's group (lane i <-> row i of the group; warp-uniform masks)
            int my_doc;
            unsigned tl;
            {
                const int64_t grp = tile * kGemmTileGroups + gi;
                int64_t r0 = dx.n_chunks, r1 = dx.n_chunks;
                if (grp < g.n_groups) { r0 = g.group_row[grp]; r1 = g.group_row[grp + 1]; }
                const bool valid = lane < int(r1 - r0);
                const int d = valid ? dx.row_doc[r0 + lane] : (-2 - lane);
                const int nd = __shfl_down_sync(0xffffffffu, d, 1);
                my_doc = d;
                tl = __ballot_sync(0xffffffffu, valid && (lane == 31 || nd != d));   // last row of every document
            }
            mbarrier_wait(&s_tfull[buf], use & 1u);
            tcgen05_fence_after();
            if (tl != 0u) {                                               // uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_acc + (uint32_t(quarter * 32) << 16) + uint32_t(buf * kGemmTileRows + gi * kGemmGroupRows), r);
                // straight-line pass over the 32 columns: running max of the current document (boundaries are
                // warp-uniform bits), bit i of passmask = "column i closes a document whose max reaches my bound"
                float mm[32];
                unsigned passmask = 0u;
                float m = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    m = fmaxf(m, __uint_as_float(r[i]));
                    mm[i] = m;
                    if (m >= tau) passmask |= 1u << i;                    // masked with the document ends below
                    m = ((tl >> i) & 1u) ? -INFINITY : m;
                }
                passmask &= tl;
                unsigned colmask = __reduce_or_sync(0xffffffffu, passmask);
                while (colmask) {                                         // stage (query, key) entries: no global round trip
                    const int i = __ffs(colmask) - 1;
                    colmask &= colmask - 1;
                    float vi = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) vi = (j == i) ? mm[j] : vi;
                    vi += 0.0f;
                    const bool pass = (passmask >> i) & 1u;
                    const unsigned pm = __ballot_sync(0xffffffffu, pass);
                    const int n = __popc(pm);
                    if (staged + n > kGemmStage) flush();
                    const int doc = __shfl_sync(0xffffffffu, my_doc, i);
                    if (pass) {
                        const int q = g.q0 + my_q;
                        const uint32_t key = float_to_key(vi);
                        const int e = staged + __popc(pm & lt_mask);
                        st_key[e] = make_key64(key, dx.doc_base + uint32_t(doc));
                        st_q[e] = uint16_t(q);
                        if (w.use_tau) tau_count(w.ts, q, key);           // fire-and-forget histogram update
                    }
                    staged += n;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbarrier_arrive(&s_tempty[buf]);               // accumulator may be overwritten
            if (staged > kGemmStage / 2) flush();
            // refresh the bound of ONE query per tile, round-robin over warps and tiles
            if (w.use_tau) { tau_raise(w.ts, g.q0 + rr); rr = (rr + 1 == g.n_real) ? 0 : rr + 1; }
        }
        flush();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// emission log -> per-query candidate lists
__global__ void gemm_bucket_kernel(DenseWork w) {
    const unsigned long long n_raw = *w.log_count;
    const int64_t n = n_raw < (unsigned long long)w.log_cap ? int64_t(n_raw) : w.log_cap;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int q = w.log_q[i];
        const int slot = atomicAdd(&w.cand_count[q], 1);
        if (slot < w.cap) w.cand[int64_t(q) * w.cap + slot] = w.log_key[i];
        else w.overflow[q] = 1;
    }
    if (n_raw > (unsigned long long)w.log_cap && blockIdx.x == 0)         // log overflow: every query of the group is suspect
        for (int q = threadIdx.x; q < w.n_log_queries; q += blockDim.x) w.overflow[q] = 1;
}

// Q fp32 [n_real][768] -> bf16 [n_pad][768], zero padded
__global__ void gemm_pack_q_kernel(const float* __restrict__ q, __nv_bfloat16* __restrict__ out, int n_real, int n_pad) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= int64_t(n_pad) * kDim) return;
    const int row = int(i / kDim);
    out[i] = row < n_real ? __float2bfloat16_rn(q[i]) : __float2bfloat16_rn(0.f);
}

}  // namespace mse
