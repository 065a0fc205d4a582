// K4 — exhaustive dense scan for large query batches on the 5th-gen tensor cores (tcgen05).
//
// S^T[B x rows] = Q[B x 768] * E^T, bf16 operands, fp32 accumulators in TMEM; the epilogue is the
// same per-document max + running-bound candidate emission as the GEMV kernel (dense.cuh), so the
// score matrix never exists in HBM.
//
// Roles of the operands: the QUERIES are the M dimension (TMEM lanes), the chunk rows the N dimension (TMEM
// columns).  A tcgen05.ld (32x32b) therefore hands every epilogue thread 32 consecutive chunk scores of ITS
// query: the per-document max is a running max over registers with warp-uniform document boundaries — no
// shuffles — and the query's running bound is one register.  (The first version of this kernel had the chunk
// rows on the lanes and spent 20k warp-instructions per tile on segmented max-scans across lanes.)
//
// Tiling: UMMA M=128 queries (one or two M tiles for B <= 128 / <= 256) x N=128 chunk rows x K=16, 12 k-blocks
// of 64 elements (one 128-byte SWIZZLE_128B row per k-block).  A 128-row tile is assembled from FOUR
// doc-aligned groups of <= 32 rows (four {64 x 32} TMA boxes per k-block, each starting at the first row of a
// document), so a 32-column tcgen05.ld covers whole documents only.  (Rows past the end of a group belong to
// the next group; they are computed twice and masked, ~8 % redundant L2 reads.)  Documents with more than 32
// chunks do not fit a group: the caller falls back to the GEMV kernel for such corpora.
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2-9 = epilogue (warp w reads TMEM lane quarter w%4 =
// queries 32*(w%4).. of an M tile; the two warps of a quarter take one M tile each, or two groups each when
// there is a single M tile).
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA, freed by tcgen05.commit) and a
// double-buffered TMEM accumulator full/empty pair (MMA <-> epilogue), so the epilogue of tile i
// overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "dense.cuh"
#include "topk.cuh"

namespace mse {

constexpr int kGemmEpiWarps = 8;                      // two per TMEM lane quarter, alternating 32-column chunks
constexpr int kGemmThreads = (2 + kGemmEpiWarps) * 32;
constexpr int kGemmBlockK = 64;                        // elements per k-block (128 B of bf16)
constexpr int kGemmKBlocks = kDim / kGemmBlockK;       // 12
constexpr int kGemmTileRows = 128;
constexpr int kGemmGroupRows = 32;
constexpr int kGemmATileBytes = kGemmTileRows * 128;   // 16 KB per stage
constexpr int kGemmStage = 96;                         // emissions staged per epilogue warp between flushes

// shared-memory bytes of one pipeline stage: 128 chunk rows + one or two M tiles of 128 query rows, 128 B each (the
// MMA addresses whole M tiles; rows past the padded batch are never loaded and only feed accumulator lanes nobody reads)
__host__ __device__ inline uint32_t gemm_stage_bytes(int n_pad) {
    return uint32_t(kGemmATileBytes) + uint32_t(n_pad > 128 ? 2 : 1) * 128u * 128u;
}

struct GemmWork {
    const int64_t* group_row;    // [n_groups + 1] first row of every doc-aligned group (<= 32 rows each)
    int64_t n_groups;
    int64_t n_tiles;             // ceil(n_groups / 4)
    int32_t n_pad;               // padded batch (multiple of 32)
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
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_q,
                  DenseDev dx, DenseWork w, GemmWork g) {
    extern __shared__ __align__(1024) unsigned char gemm_smem_raw[];
    constexpr int kMaxStages = 8;
    __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ uint64_t s_stage_key[kGemmEpiWarps][kGemmStage];
    __shared__ uint16_t s_stage_q[kGemmEpiWarps][kGemmStage];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_pad = g.n_pad;
    const uint32_t tx_bytes = uint32_t(kGemmATileBytes) + uint32_t(n_pad) * 128u;        // bytes TMA delivers per stage
    const uint32_t stage_bytes = gemm_stage_bytes(n_pad);                                // a whole M tile of query rows is addressable
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~uintptr_t(1023));
    const int MT = n_pad > 128 ? 2 : 1;                                   // M tiles of 128 queries
    const uint32_t acc_cols = uint32_t(MT) * kGemmTileRows;               // accumulator columns per buffer
    const uint32_t tmem_cols = 2 * acc_cols;                              // 256 or 512: a power of two

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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                int rows[4];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    const int64_t grp = tile * 4 + gi;
                    rows[gi] = grp < g.n_groups ? int(g.group_row[grp]) : int(dx.n_chunks);     // past the end -> zero fill
                }
                for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                    mbarrier_wait_backoff(&s_empty[stage], phase ^ 1u);
                    unsigned char* sa = smem + size_t(stage) * stage_bytes;
                    mbarrier_expect_tx(&s_full[stage], tx_bytes);
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi)
                        tma_load_2d(sa + gi * (kGemmGroupRows * 128), &map_e, kb * kGemmBlockK, rows[gi], &s_full[stage]);
                    tma_load_2d(sa + kGemmATileBytes, &map_q, kb * kGemmBlockK, 0, &s_full[stage]);
                    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_f32(128, kGemmTileRows);   // M = 128 queries, N = 128 chunk rows
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t use = uint32_t(it >> 1);
                mbarrier_wait_backoff(&s_tempty[buf], (use & 1u) ^ 1u);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(buf) * acc_cols;
                for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                    mbarrier_wait(&s_full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t e_addr = smem_addr(smem + size_t(stage) * stage_bytes);   // 128 chunk rows x 64
                    const uint32_t q_addr = e_addr + kGemmATileBytes;                        // n_pad queries x 64
                    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                        for (int k = 0; k < kGemmBlockK / 16; ++k) {
                            if (g.debug & 2) break;
                            const uint64_t ad = umma_desc_sw128(q_addr + uint32_t(mt) * (128u * 128u) + k * 32);
                            const uint64_t bd = umma_desc_sw128(e_addr + k * 32);
                            tcgen05_mma_bf16(tmem_d + uint32_t(mt) * kGemmTileRows, ad, bd, idesc, (kb | k) ? 1u : 0u);
                        }
                    }
                    tcgen05_commit(&s_empty[stage]);                      // frees the smem slot when the MMAs retire
                    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
                }
                tcgen05_commit(&s_tfull[buf]);                            // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue (warp w -> TMEM lane quarter w % 4 = 32 queries of an M tile) ==========
        const int quarter = warp & 3;
        const int ew = warp - 2;                                          // 0..7
        const int half = ew >> 2;                                         // which of the quarter's two warps
        const unsigned lt_mask = (1u << lane) - 1u;
        uint64_t* st_key = s_stage_key[ew];
        uint16_t* st_q = s_stage_q[ew];
        int staged = 0;                                                   // uniform
        int rr = int((blockIdx.x * kGemmEpiWarps + ew) % g.n_real);       // round-robin cursor for bound refreshes
        const int mt = MT == 2 ? half : 0;                                // this warp's M tile
        const int g_lo = MT == 2 ? 0 : 2 * half, g_hi = MT == 2 ? 4 : 2 * half + 2;   // and its doc-aligned groups
        const int my_q = mt * 128 + quarter * 32 + lane;                  // this thread's query (may be padding)
        const bool q_real = my_q < g.n_real;
        const bool warp_idle = mt * 128 + quarter * 32 >= g.n_real;       // no real query on these lanes
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
        for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
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
            // document layout of the warp's groups (lane i <-> row i of the group; warp-uniform masks)
            int my_doc[4];
            unsigned tails[4];
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
                my_doc[gi] = -1; tails[gi] = 0u;
                if (gi >= g_lo && gi < g_hi) {
                    const int64_t grp = tile * 4 + gi;
                    int64_t r0 = dx.n_chunks, r1 = dx.n_chunks;
                    if (grp < g.n_groups) { r0 = g.group_row[grp]; r1 = g.group_row[grp + 1]; }
                    const bool valid = lane < int(r1 - r0);
                    const int d = valid ? dx.row_doc[r0 + lane] : (-2 - lane);
                    const int nd = __shfl_down_sync(0xffffffffu, d, 1);
                    my_doc[gi] = d;
                    tails[gi] = __ballot_sync(0xffffffffu, valid && (lane == 31 || nd != d));   // last row of every document
                }
            }
            mbarrier_wait(&s_tfull[buf], use & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
                if (gi < g_lo || gi >= g_hi || tails[gi] == 0u) continue;  // uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf) * acc_cols +
                                   uint32_t(mt * kGemmTileRows + gi * kGemmGroupRows), r);
                const unsigned tl = tails[gi];
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
                    const int doc = __shfl_sync(0xffffffffu, my_doc[gi], i);
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
