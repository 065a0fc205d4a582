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
// Measured on B200 (tools/microbench/mma_rate.cu): a 128x256x16 tcgen05.mma runs at its arithmetic floor of 128 cycles
// (2.2 PFLOP/s over 148 SMs on zero operands) ONLY when the issuing warp is converged; the same instruction issued
// from inside `if (lane == 0)` costs ~300 cycles whatever its shape, because the compiler wraps every tcgen05 / TMA
// instruction of a divergent region in an elect-broadcast-branch loop.  Issue overhead of the converged form is
// ~110 cycles per MMA, so N = 256 chunk rows per instruction is the shape that keeps the tensor pipe busy.
//
// Operand placement: the query panel (128 queries x 768 bf16) stays on chip for the whole kernel — the first 512 K
// elements in TENSOR MEMORY (256 columns, lane = query, two K elements per 32-bit column; tcgen05.mma with the A
// operand in TMEM), the last 256 in shared memory (64 KB, K-major SWIZZLE_128B, loaded once by TMA) — because TMEM
// also has to hold the 256-column accumulator.  Only the chunk rows stream: 4 stages of 32 KB (256 rows x one 64-element
// k-block), so the L2->SM path carries E only.  (Round 1 re-fetched the 384 KB query panel from L2 for every 128-row
// tile and issued 128x128x16 MMAs from a divergent lane: 4.7-5.0 ms at B=256.)
// A CTA owns ONE panel of <= 128 queries; a batch of 129..256 queries runs as two interleaved sets of CTAs (even /
// odd blockIdx) that walk the same tiles in the same order, so the second reader of an E tile hits L2.
//
// Tiling: UMMA M=128 queries x N=256 chunk rows x K=16, 12 k-blocks of 64 elements (one 128-byte SWIZZLE_128B row
// per k-block), one 256-column accumulator (the loads of the next tile run under the epilogue).  A 256-row tile is
// assembled from EIGHT doc-aligned groups of <= 32 rows (eight {64 x 32} TMA boxes per k-block, each starting at the
// first row of a document), so a 32-column tcgen05.ld covers whole documents only.  (Rows past the end of a group
// belong to the next group; they are computed twice and masked, ~8 % redundant L2 reads.)  Documents with more
// than 32 chunks do not fit a group: the caller falls back to the GEMV kernel.
//
// Warp roles (352 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer (converged, one elected lane issues),
// warp 1 = TMEM allocator + MMA issuer (the same), warps 2-9 = epilogue, warp 10 = bound refresher (warp w reads TMEM lane quarter w%4 =
// queries 32*(w%4).. of the panel; the two warps of a quarter take four groups each).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dense.cuh"
#include "topk.cuh"

namespace mse {

constexpr int kGemmEpiWarps = 8;                      // two per TMEM lane quarter, four doc-aligned groups each
constexpr int kGemmThreads = (3 + kGemmEpiWarps) * 32;   // + TMA producer, MMA issuer, bound warp
constexpr int kGemmBlockK = 64;                        // elements per k-block (128 B of bf16)
constexpr int kGemmKBlocks = kDim / kGemmBlockK;       // 12
constexpr int kGemmGroupRows = 32;
constexpr int kGemmTileGroups = 8;
constexpr int kGemmTileRows = kGemmTileGroups * kGemmGroupRows;   // 256 chunk rows = N of the MMA
constexpr int kGemmStageBytes = kGemmTileRows * 128;   // 32 KB per stage (one k-block of a tile)
constexpr int kGemmMaxStages = 4;                     // 12 k-blocks per tile = 3 turns of the ring: stage = kb & 3
constexpr int kGemmPanel = 128;                        // queries per CTA
constexpr int kGemmKBlocksTmem = 8;                    // k-blocks of the query panel kept in TMEM (the rest in shared memory)
constexpr int kGemmACols = kGemmKBlocksTmem * kGemmBlockK / 2;    // 256 TMEM columns (2 bf16 per column)
constexpr int kGemmASmemBytes = (kGemmKBlocks - kGemmKBlocksTmem) * kGemmPanel * 128;   // 64 KB
constexpr int kGemmDCol = 256;                         // first accumulator column
constexpr int kGemmSmemBytes = kGemmASmemBytes + kGemmMaxStages * kGemmStageBytes + 1024;
constexpr int kGemmCg2Stages = 8;                      // MODE 2: stages of 16 KB per CTA (half a k-block of the tile each)
constexpr int kGemmStage = 96;                         // emissions staged per epilogue warp between flushes

struct GemmWork {
    const int64_t* group_row;    // [n_groups + 1] first row of every doc-aligned group (<= 32 rows each)
    int64_t n_groups;
    int64_t n_tiles;             // ceil(n_groups / kGemmTileGroups)
    const __nv_bfloat16* qb16;   // [n_panels * 128][768] bf16 queries, zero padded (map_q describes the same array)
    int32_t n_panels;            // 1 or 2 panels of 128 queries
    int32_t n_real;              // real queries in this launch
    int32_t q0;                  // first query (index into cand / tau arrays)
    int32_t stages;
    int32_t debug;               // bit0: skip epilogue math (timing experiments only)
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
// ---- warp-converged issue -------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit / cp.async.bulk.tensor take their operands from UNIFORM registers.  Issued from inside
// `if (lane == 0)` the compiler cannot prove uniformity and wraps every instruction in an elect / broadcast / branch
// loop: ~300 cycles per MMA on B200 whatever its shape (tools/microbench/mma_rate.cu), more than twice the 128 cycles a
// 128x256x16 MMA takes on the tensor pipe.  The producer and MMA warps therefore run CONVERGED (all 32 lanes execute
// the loop and the barrier waits) and one elected lane issues a whole k-block per asm block.
//
// One k-block of MMAs (4 x K=16) + the commit that frees the shared-memory stage.  A from tensor memory.
template <bool PAIR>
__device__ __forceinline__ void umma_kblock_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate_first, uint64_t* empty_bar) {
    if constexpr (PAIR) {
    asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 b1, b2, b3;\n\t.reg .b16 m;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "mov.b16 m, 3;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
                 "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%5], m;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_addr(empty_bar)) : "memory");
    return;
    }
    asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 b1, b2, b3;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
                 "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_addr(empty_bar)) : "memory");
}
// the same with A from shared memory (descriptor; +32 bytes = +2 in the 16-byte address field per K=16 step)
template <bool PAIR>
__device__ __forceinline__ void umma_kblock_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate_first, uint64_t* empty_bar) {
    if constexpr (PAIR) {
    asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t.reg .b16 m;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "mov.b16 m, 3;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
                 "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%5], m;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_addr(empty_bar)) : "memory");
    return;
    }
    asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
                 "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_addr(empty_bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
// one stage of the E ring: expect_tx + eight {64 x 32-row} boxes (box i lands 4 KB after box i - 1), one elected lane
__device__ __forceinline__ void tma_stage_elect(uint32_t dst, const CUtensorMap* map, int c0, const int (&rows)[8],
                                                uint64_t* full_bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 d1, d2, d3, d4, d5, d6, d7;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "add.u32 d1, %0, 4096;\n\tadd.u32 d2, %0, 8192;\n\tadd.u32 d3, %0, 12288;\n\tadd.u32 d4, %0, 16384;\n\t"
                 "add.u32 d5, %0, 20480;\n\tadd.u32 d6, %0, 24576;\n\tadd.u32 d7, %0, 28672;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %12;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %4}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d1], [%1, {%2, %5}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d2], [%1, {%2, %6}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d3], [%1, {%2, %7}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d4], [%1, {%2, %8}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d5], [%1, {%2, %9}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d6], [%1, {%2, %10}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [d7], [%1, {%2, %11}], [%3];\n\t}"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(smem_addr(full_bar)),
                   "r"(rows[0]), "r"(rows[1]), "r"(rows[2]), "r"(rows[3]), "r"(rows[4]), "r"(rows[5]), "r"(rows[6]), "r"(rows[7]),
                   "r"(bytes) : "memory");
}
// ---- CTA pair (cluster of 2): the two query panels of a 129..256-query batch share every E tile ----------------------
// Each CTA of the pair loads HALF of a stage (four of the eight boxes) and multicasts it into the shared memory of
// both, so a chunk row leaves HBM / L2 once per pair instead of once per panel.  A stage is free when BOTH CTAs'
// MMAs have read it: the commit that frees it is multicast to the empty barrier of both CTAs (arrival count 2).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_half_stage_multicast_elect(uint32_t dst, const CUtensorMap* map, int c0, const int (&rows)[4],
                                                               uint64_t* full_bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 d1, d2, d3;\n\t.reg .b16 m;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "mov.b16 m, 3;\n\t"
                 "add.u32 d1, %0, 4096;\n\tadd.u32 d2, %0, 8192;\n\tadd.u32 d3, %0, 12288;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %8;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %4}], [%3], m;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [d1], [%1, {%2, %5}], [%3], m;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [d2], [%1, {%2, %6}], [%3], m;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [d3], [%1, {%2, %7}], [%3], m;\n\t}"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(smem_addr(full_bar)),
                   "r"(rows[0]), "r"(rows[1]), "r"(rows[2]), "r"(rows[3]), "r"(bytes) : "memory");
}
// ---- CTA pair with cta_group::2 MMAs (kernel MODE 2) ---------------------------------------------------------------
// One tcgen05.mma.cta_group::2 (M = 256: 128 queries per CTA, N = 256 chunk rows) is issued by the leader CTA (rank 0) for
// the pair.  Each CTA keeps its own query panel (tensor + shared memory) and its own 128 x 256 accumulator; the E tile
// is SPLIT: CTA r holds rows [128 r, 128 r + 128) of the tile, so a ring stage is 16 KB per CTA instead of 32 KB and the
// same shared memory holds twice as many k-blocks in flight.  Barriers: every TMA load of the pair completes on the
// LEADER's full barrier; the leader's commits are multicast to the empty / accumulator-full barriers of both CTAs; the
// epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t local_addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_addr));
    return r;
}
__device__ __forceinline__ void mbarrier_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// four {64 x 32-row} boxes into this CTA's half stage; completion counted on the leader's barrier (cluster address)
__device__ __forceinline__ void tma_half_stage_cg2_elect(uint32_t dst, const CUtensorMap* map, int c0, const int (&rows)[4],
                                                         uint32_t leader_full_bar) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 d1, d2, d3;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "add.u32 d1, %0, 4096;\n\tadd.u32 d2, %0, 8192;\n\tadd.u32 d3, %0, 12288;\n\t"
                 "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %4}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [d1], [%1, {%2, %5}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [d2], [%1, {%2, %6}], [%3];\n\t"
                 "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [d3], [%1, {%2, %7}], [%3];\n\t}"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(leader_full_bar),
                   "r"(rows[0]), "r"(rows[1]), "r"(rows[2]), "r"(rows[3]) : "memory");
}
__device__ __forceinline__ void mbarrier_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// one k-block of cta_group::2 MMAs + the multicast commit that frees the stage in both CTAs
__device__ __forceinline__ void umma_kblock_ts_cg2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                   uint32_t accumulate_first, uint64_t* empty_bar) {
    asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 b1, b2, b3;\n\t.reg .b16 m;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "mov.b16 m, 3;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
                 "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a1], b1, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], b2, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], b3, %3, t;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%5], m;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_addr(empty_bar)) : "memory");
}
__device__ __forceinline__ void umma_kblock_ss_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                   uint32_t accumulate_first, uint64_t* empty_bar) {
    asm volatile("{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t.reg .b16 m;\n\t"
                 "elect.sync _|q, 0xffffffff;\n\t"
                 "mov.b16 m, 3;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
                 "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
                 "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, t;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, t;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%5], m;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first), "r"(smem_addr(empty_bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_cg2_multicast_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 m;\n\telect.sync _|q, 0xffffffff;\n\tmov.b16 m, 3;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
                 ::"r"(smem_addr(bar)) : "memory");
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
}
// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: one panel per CTA, no cluster.  MODE 1: cluster of two CTAs = two panels, cta_group::1 MMAs, every E stage
// multicast into both CTAs.  MODE 2: cluster of two CTAs = two panels, cta_group::2 MMAs issued by rank 0, E stages split.
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_q,
                  DenseDev dx, DenseWork w, GemmWork g) {
    constexpr bool PAIR = MODE != 0;
    constexpr bool CG2 = MODE == 2;
    constexpr int kStages = CG2 ? kGemmCg2Stages : kGemmMaxStages;
    constexpr int kStageBytes = CG2 ? kGemmStageBytes / 2 : kGemmStageBytes;
    extern __shared__ __align__(1024) unsigned char gemm_smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kGemmCg2Stages], s_empty[kGemmCg2Stages], s_tfull, s_tempty, s_aready;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_epi_done;                                            // epilogue warps that have finished (stops the bound warp)
    __shared__ uint64_t s_stage_key[kGemmEpiWarps][kGemmStage];
    __shared__ uint16_t s_stage_q[kGemmEpiWarps][kGemmStage];
    __shared__ __align__(16) uint32_t s_transpose[kGemmEpiWarps][32];     // one query's 32 scores of a group, lane <-> row

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* a_ss = smem;                                           // query panel, k-blocks 8..11: 4 x {128 rows x 128 B}
    unsigned char* ring = smem + kGemmASmemBytes;                         // E stages
    constexpr uint32_t tmem_cols = 512;                                   // 256 query panel + 256 accumulator
    // CTAs of panel p are blockIdx p, p + n_panels, ...: both panels walk the same tiles in the same order
    // (PAIR: the two CTAs of a cluster are the two panels and share the E loads; otherwise a single panel)
    const int panel = PAIR ? int(cluster_ctarank()) : int(blockIdx.x) % g.n_panels;
    const int64_t cta = int64_t(blockIdx.x) / g.n_panels, n_cta = int64_t(gridDim.x) / g.n_panels;

    if (threadIdx.x == 0) {
        s_epi_done = 0;
        // MODE 1: a stage is free when the MMAs of both CTAs have read it (two multicast commits); MODE 2: one multicast
        // commit of the leader per CTA, and the leader's accumulator-empty barrier collects the epilogue warps of both CTAs
        for (int s = 0; s < kStages; ++s) { mbarrier_init(&s_full[s], 1); mbarrier_init(&s_empty[s], MODE == 1 ? 2 : 1); }
        mbarrier_init(&s_tfull, 1); mbarrier_init(&s_tempty, CG2 ? 2 * kGemmEpiWarps : kGemmEpiWarps); mbarrier_init(&s_aready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem_base)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();                               // the peer's barriers are initialised too
    tcgen05_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t tmem_acc = tmem_base + uint32_t(kGemmDCol);

    if (warp == 0 && lane == 0) {
        // second half of the query panel -> shared memory (resident): one TMA box per k-block
        mbarrier_expect_tx(&s_aready, uint32_t(kGemmASmemBytes));
        for (int kb = kGemmKBlocksTmem; kb < kGemmKBlocks; ++kb)
            tma_load_2d(a_ss + (kb - kGemmKBlocksTmem) * (kGemmPanel * 128), &map_q, kb * kGemmBlockK, panel * kGemmPanel, &s_aready);
    }
    if (warp >= 2 && warp < 2 + kGemmEpiWarps) {
        // first half of the query panel -> TMEM: thread = one query (TMEM lane), 32 columns (64 K elements) per tcgen05.st;
        // the two warps of a lane quarter write one half of the columns each
        const int quarter = warp & 3;
        const int qrow = panel * kGemmPanel + quarter * 32 + lane;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g.qb16) + int64_t(qrow) * (kDim / 2);
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
    if constexpr (CG2) mbarrier_wait(&s_aready, 0u);                      // this CTA's shared-memory part of the panel has landed
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CG2) cluster_sync_all();                                // ... and so has the peer's: the leader's MMAs read both
    tcgen05_fence_after();

    if (warp == 0) {
        // ===================== TMA producer (converged warp, one elected lane issues) =====================
        static_assert(kGemmTileGroups == 8 && kGemmGroupRows * 128 == 4096, "tma_stage_elect is written for 8 boxes of 4 KB");
        static_assert(kGemmKBlocks % kGemmMaxStages == 0, "stage = kb % stages needs whole turns of the ring per tile");
        const uint32_t ring_addr = smem_addr(ring);
        int it = 0;
        int rstage = 0;                                                   // MODE 2: running stage / phase (12 k-blocks over 8 stages)
        uint32_t rphase = 0;
        for (int64_t tile = cta; tile < g.n_tiles; tile += n_cta, ++it) {
            constexpr int n_rows = PAIR ? kGemmTileGroups / 2 : kGemmTileGroups;   // boxes this CTA loads per stage
            const int gi0 = PAIR ? panel * n_rows : 0;
            int rows[n_rows];
#pragma unroll
            for (int gi = 0; gi < n_rows; ++gi) {
                const int64_t grp = tile * kGemmTileGroups + gi0 + gi;
                rows[gi] = grp < g.n_groups ? int(g.group_row[grp]) : int(dx.n_chunks);     // past the end -> zero fill
            }
#pragma unroll
            for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                if constexpr (CG2) {
                    // this CTA's four groups = its half of the tile's rows; the pair's bytes are counted on the leader's barrier
                    mbarrier_wait(&s_empty[rstage], rphase ^ 1u);         // the leader's MMAs have read the stage (in both CTAs)
                    if (panel == 0) mbarrier_expect_tx_elect(&s_full[rstage], uint32_t(kGemmStageBytes));
                    tma_half_stage_cg2_elect(ring_addr + uint32_t(rstage) * kStageBytes, &map_e, kb * kGemmBlockK, rows,
                                             mapa_rank0(smem_addr(&s_full[rstage])));
                    if (++rstage == kStages) { rstage = 0; rphase ^= 1u; }
                } else {
                    constexpr int turns = kGemmKBlocks / kGemmMaxStages;  // turns of the ring per tile
                    const int stage = kb % kGemmMaxStages;
                    const uint32_t phase = uint32_t(it * turns + kb / kGemmMaxStages) & 1u;
                    mbarrier_wait(&s_empty[stage], phase ^ 1u);           // MODE 1: both CTAs have consumed the stage
                    const uint32_t dst = ring_addr + uint32_t(stage) * kGemmStageBytes + uint32_t(gi0) * (kGemmGroupRows * 128);
                    if constexpr (PAIR)                                   // the other half arrives from the peer's multicast
                        tma_half_stage_multicast_elect(dst, &map_e, kb * kGemmBlockK, rows, &s_full[stage], uint32_t(kGemmStageBytes));
                    else
                        tma_stage_elect(dst, &map_e, kb * kGemmBlockK, rows, &s_full[stage], uint32_t(kGemmStageBytes));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, one elected lane issues a k-block per asm block) ==========
        // M = 128 queries (MODE 2: 256 over the pair), N = 256 chunk rows
        const uint32_t idesc = umma_idesc_bf16_f32(CG2 ? 2 * kGemmPanel : kGemmPanel, kGemmTileRows);
        const uint64_t e_desc0 = umma_desc_sw128(smem_addr(ring));
        const uint64_t a_desc0 = umma_desc_sw128(smem_addr(a_ss));
        if constexpr (!CG2) mbarrier_wait(&s_aready, 0u);                 // shared-memory part of the query panel has landed
        int it = 0;
        int rstage = 0;
        uint32_t rphase = 0;
        const bool issuer = !CG2 || panel == 0;                           // MODE 2: the leader CTA issues for the pair
        for (int64_t tile = cta; issuer && tile < g.n_tiles; tile += n_cta, ++it) {
            mbarrier_wait(&s_tempty, (uint32_t(it) & 1u) ^ 1u);          // epilogue has drained the accumulator (MODE 2: of both CTAs)
            tcgen05_fence_after();
#pragma unroll
            for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                if constexpr (CG2) {
                    mbarrier_wait(&s_full[rstage], rphase);               // both halves of the stage have landed
                    tcgen05_fence_after();
                    const uint64_t bd = e_desc0 + uint64_t(rstage * (kStageBytes >> 4));       // this CTA's 128 rows x 64; the peer's at the same offset
                    if (kb < kGemmKBlocksTmem)
                        umma_kblock_ts_cg2(tmem_acc, tmem_base + uint32_t(kb * (kGemmBlockK / 16) * 8), bd, idesc, kb ? 1u : 0u, &s_empty[rstage]);
                    else
                        umma_kblock_ss_cg2(tmem_acc, a_desc0 + uint64_t((kb - kGemmKBlocksTmem) * ((kGemmPanel * 128) >> 4)), bd, idesc, 1u,
                                           &s_empty[rstage]);
                    if (++rstage == kStages) { rstage = 0; rphase ^= 1u; }
                } else {
                    constexpr int turns = kGemmKBlocks / kGemmMaxStages;
                    const int stage = kb % kGemmMaxStages;
                    const uint32_t phase = uint32_t(it * turns + kb / kGemmMaxStages) & 1u;
                    mbarrier_wait(&s_full[stage], phase);
                    tcgen05_fence_after();
                    const uint64_t bd = e_desc0 + uint64_t(stage * (kGemmStageBytes >> 4));      // 256 chunk rows x 64
                    if (kb < kGemmKBlocksTmem)                            // A: 16 K elements = 8 TMEM columns per MMA
                        umma_kblock_ts<MODE == 1>(tmem_acc, tmem_base + uint32_t(kb * (kGemmBlockK / 16) * 8), bd, idesc, kb ? 1u : 0u, &s_empty[stage]);
                    else
                        umma_kblock_ss<MODE == 1>(tmem_acc, a_desc0 + uint64_t((kb - kGemmKBlocksTmem) * ((kGemmPanel * 128) >> 4)), bd, idesc, 1u,
                                                  &s_empty[stage]);
                }
            }
            if constexpr (CG2) tcgen05_commit_cg2_multicast_elect(&s_tfull);     // accumulators of both CTAs ready
            else tcgen05_commit_elect(&s_tfull);                          // accumulator ready for the epilogue
        }
    } else if (warp == 2 + kGemmEpiWarps) {
        // ===================== bound warp: raises the running bounds of the queries, four at a time, for as long as the
        // epilogue runs (kept off the epilogue warps: a refresh is three dependent L2 round trips, ~2.5k cycles that
        // would sit between two tiles of every epilogue warp) =====================
        if (w.use_tau && !(g.debug & 1)) {
            int q = int((int64_t(blockIdx.x) * 4) % g.n_real);
            while (*reinterpret_cast<volatile int*>(&s_epi_done) < kGemmEpiWarps) {
                int qs[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { qs[i] = g.q0 + q; q = (q + 1 == g.n_real) ? 0 : q + 1; }
                tau_raise_multi<4>(w.ts, qs);
            }
        }
    } else {
        // ===================== epilogue (warp w -> TMEM lane quarter w % 4 = 32 queries of the panel) ==========
        const int quarter = warp & 3;
        const int ew = warp - 2;                                          // 0..7
        const int g_lo = (ew >> 2) * (kGemmTileGroups / 2);               // this warp's doc-aligned groups of the tile
        const unsigned lt_mask = (1u << lane) - 1u;
        uint64_t* st_key = s_stage_key[ew];
        uint16_t* st_q = s_stage_q[ew];
        uint32_t* tr = s_transpose[ew];
        int staged = 0;                                                   // uniform
        int rr2 = int((blockIdx.x * kGemmEpiWarps + ew) % g.n_real);
        const int my_q = panel * kGemmPanel + quarter * 32 + lane;        // this thread's query (may be padding)
        const bool q_real = my_q < g.n_real;
        const bool warp_idle = panel * kGemmPanel + quarter * 32 >= g.n_real;   // no real query on these lanes
        const uint32_t tempty_leader = CG2 ? mapa_rank0(smem_addr(&s_tempty)) : 0u;
        auto tempty_arrive = [&]() {                                      // MODE 2: the leader CTA's barrier collects both CTAs
            if constexpr (CG2) mbarrier_arrive_cluster(tempty_leader); else mbarrier_arrive(&s_tempty);
        };
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
            const uint32_t par = uint32_t(it) & 1u;
            if (warp_idle || (g.debug & 1)) {
                mbarrier_wait_backoff(&s_tfull, par);
                if (lane == 0) tempty_arrive();
                continue;
            }
            // this thread's running bound (padding lanes never pass)
            float tau = INFINITY;
            if (q_real) {
                const uint32_t tk = w.use_tau ? ld_relaxed_u32(&w.ts.tau[g.q0 + my_q]) : 0u;
                tau = tk ? key_to_float(tk) : -INFINITY;
            }
            // document layout of the warp's groups (lane i <-> row i of the group; warp-uniform masks)
            int my_doc[kGemmTileGroups / 2];
            unsigned tails[kGemmTileGroups / 2];
#pragma unroll
            for (int j = 0; j < kGemmTileGroups / 2; ++j) {
                const int64_t grp = tile * kGemmTileGroups + g_lo + j;
                int64_t r0 = dx.n_chunks, r1 = dx.n_chunks;
                if (grp < g.n_groups) { r0 = g.group_row[grp]; r1 = g.group_row[grp + 1]; }
                const bool valid = lane < int(r1 - r0);
                const int d = valid ? dx.row_doc[r0 + lane] : (-2 - lane);
                const int nd = __shfl_down_sync(0xffffffffu, d, 1);
                my_doc[j] = d;
                tails[j] = __ballot_sync(0xffffffffu, valid && (lane == 31 || nd != d));   // last row of every document
            }
            mbarrier_wait(&s_tfull, par);
            tcgen05_fence_after();
            // The accumulator is single-buffered (TMEM also holds half of the query panel): copy this warp's 4 x 32 columns
            // to registers first and hand the accumulator back at once, so that the MMAs of the next tile run under the
            // per-document pass below.
            uint32_t racc[kGemmTileGroups / 2][32];
#pragma unroll
            for (int j = 0; j < kGemmTileGroups / 2; ++j)
                if (tails[j] != 0u)                                       // uniform
                    tmem_ld_32x32b_x32(tmem_acc + (uint32_t(quarter * 32) << 16) + uint32_t((g_lo + j) * kGemmGroupRows), racc[j]);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tempty_arrive();                    // accumulator may be overwritten
#pragma unroll
            for (int j = 0; j < kGemmTileGroups / 2; ++j) {
                if (tails[j] == 0u) continue;                             // uniform
                uint32_t (&r)[32] = racc[j];
                // Pre-filter: a document reaches the bound of a query iff one of its chunk scores does, so only lanes
                // (queries) whose largest score of the group reaches their bound can emit: 16 three-input max + one
                // vote per group.  (Columns past the end of the group only make the test conservative.)  Streaming a
                // corpus of N documents for the best k emits ~k ln(N / k) candidates per query, i.e. for C3 less than
                // one lane per group on average.
                float gm = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
#pragma unroll
                for (int i = 3; i + 1 < 32; i += 2) gm = fmax3(gm, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                gm = fmaxf(gm, __uint_as_float(r[31]));
                unsigned hm = __ballot_sync(0xffffffffu, gm >= tau);
                if (hm == 0u) continue;                                   // uniform
                // Exact pass, one hit lane at a time: the lane's 32 scores go through shared memory so that lane i holds
                // the score of row i of the group, a segmented max over the lanes of each document (rows of a document
                // are adjacent) leaves the document's max on its last row, and those rows test it against the bound.
                const unsigned tl = tails[j];
                const unsigned heads = (tl << 1) | 1u;                    // first row of every document
                const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                const bool is_tail = (tl >> lane) & 1u;
                while (hm) {                                              // uniform
                    const int hl = __ffs(hm) - 1;
                    hm &= hm - 1;
                    __syncwarp();
                    if (lane == hl) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(&tr[i]) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
                    }
                    __syncwarp();
                    float v = __uint_as_float(tr[lane]);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const float t = __shfl_up_sync(0xffffffffu, v, d);
                        if (lane - d >= seg_start) v = fmaxf(v, t);
                    }
                    const float tau_hl = __shfl_sync(0xffffffffu, tau, hl);
                    const bool pass = is_tail && v >= tau_hl;
                    const unsigned pm = __ballot_sync(0xffffffffu, pass);
                    if (pm == 0u) continue;                               // uniform
                    const int n = __popc(pm);
                    if (staged + n > kGemmStage) flush();
                    if (pass) {
                        const int q = g.q0 + panel * kGemmPanel + quarter * 32 + hl;
                        const uint32_t key = float_to_key(v + 0.0f);
                        const int e = staged + __popc(pm & lt_mask);
                        st_key[e] = make_key64(key, dx.doc_base + uint32_t(my_doc[j]));
                        st_q[e] = uint16_t(q);
                        if (w.use_tau) tau_count(w.ts, q, key);           // fire-and-forget histogram update
                    }
                    staged += n;
                }
            }
            if (staged > kGemmStage / 2) flush();
            // two panels (129..256 queries): the bound warp alone does not keep 256 bounds fresh enough; every epilogue warp adds
            // one refresh per tile (A/B on one board: 4.65 -> 4.40 ms at B=256, but 2.96 -> 3.00 ms at B=128, so one panel skips it)
            if (g.n_panels == 2 && w.use_tau) { tau_raise(w.ts, g.q0 + rr2); rr2 = (rr2 + 1 == g.n_real) ? 0 : rr2 + 1; }
        }
        flush();
        if (lane == 0) atomicAdd(&s_epi_done, 1);
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();                               // no exit while the peer may still signal this CTA's barriers
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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
