// K4 — exhaustive dense scan for large query batches on the 5th-gen tensor cores (tcgen05).
//
// S[rows x B] = E[rows x 768] * Q^T, bf16 operands, fp32 accumulators in TMEM; the epilogue is the
// same per-document max + running-bound candidate emission as the GEMV kernel (dense.cuh), so the
// score matrix never exists in HBM.
//
// Tiling: UMMA M=128 (cta_group::1) x N=B padded to a multiple of 32 (<= 256) x K=16, 12 k-blocks
// of 64 elements (one 128-byte SWIZZLE_128B row per k-block).  A 128-row MMA tile is assembled from
// FOUR doc-aligned groups of <= 32 rows (four {64 x 32} TMA boxes per k-block, each starting at the
// first row of a document): TMEM lane quarter w then holds whole documents only, and epilogue warp
// w — the only warp allowed to read that quarter — needs no cross-warp exchange for the per-doc max.
// (Rows past the end of a group belong to the next group; they are computed twice and masked, ~8 %
// redundant L2 reads.)  Documents with more than 32 chunks do not fit a group: the caller falls back
// to the GEMV kernel for such corpora.
//
// Warp roles (320 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2-9 = epilogue (tcgen05.ld 32x32b; warp w reads
// TMEM lane quarter w%4, the two warps of a quarter take alternate 32-column chunks).
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
    __shared__ __align__(16) float s_tau[kGemmEpiWarps][256];
    __shared__ uint64_t s_stage_key[kGemmEpiWarps][kGemmStage];
    __shared__ uint16_t s_stage_q[kGemmEpiWarps][kGemmStage];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_pad = g.n_pad;
    const uint32_t b_bytes = uint32_t(n_pad) * 128u;
    const uint32_t stage_bytes = uint32_t(kGemmATileBytes) + b_bytes;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~uintptr_t(1023));
    uint32_t tmem_cols = 32;
    while (tmem_cols < uint32_t(2 * n_pad)) tmem_cols <<= 1;

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
                    mbarrier_expect_tx(&s_full[stage], stage_bytes);
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
            const uint32_t idesc = umma_idesc_bf16_f32(kGemmTileRows, n_pad);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t use = uint32_t(it >> 1);
                mbarrier_wait_backoff(&s_tempty[buf], (use & 1u) ^ 1u);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(buf * n_pad);
                for (int kb = 0; kb < kGemmKBlocks; ++kb) {
                    mbarrier_wait(&s_full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_addr(smem + size_t(stage) * stage_bytes);
                    const uint32_t b_addr = a_addr + kGemmATileBytes;
#pragma unroll
                    for (int k = 0; k < kGemmBlockK / 16; ++k) {
                        if (g.debug & 2) break;
                        const uint64_t ad = umma_desc_sw128(a_addr + k * 32);
                        const uint64_t bd = umma_desc_sw128(b_addr + k * 32);
                        tcgen05_mma_bf16(tmem_d, ad, bd, idesc, (kb | k) ? 1u : 0u);
                    }
                    tcgen05_commit(&s_empty[stage]);                      // frees the smem slot when the MMAs retire
                    if (++stage == g.stages) { stage = 0; phase ^= 1u; }
                }
                tcgen05_commit(&s_tfull[buf]);                            // accumulator ready for the epilogue
            }
        }
    } else {
        // ===================== epilogue (warp w -> TMEM lane quarter w % 4) =====================
        const int quarter = warp & 3;
        const int ew = warp - 2;                                          // 0..7
        const int half = ew >> 2;                                         // which of the quarter's two warps
        const unsigned lt_mask = (1u << lane) - 1u;
        float* my_tau = s_tau[ew];
        uint64_t* st_key = s_stage_key[ew];
        uint16_t* st_q = s_stage_q[ew];
        int staged = 0;                                                   // uniform
        int rr = int((blockIdx.x * kGemmEpiWarps + ew) % g.n_real);       // round-robin cursor for bound refreshes
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
            if (half * 32 >= g.n_real) {                                  // small batch: this warp has no column chunk
                mbarrier_wait_backoff(&s_tfull[buf], use & 1u);
                if (lane == 0) mbarrier_arrive(&s_tempty[buf]);
                continue;
            }
            // this warp's doc-aligned group
            const int64_t grp = tile * 4 + quarter;
            int64_t r0 = dx.n_chunks, r1 = dx.n_chunks;
            if (grp < g.n_groups) { r0 = g.group_row[grp]; r1 = g.group_row[grp + 1]; }
            const int nrow = int(r1 - r0);
            const bool valid = lane < nrow;
            const int my_doc = valid ? dx.row_doc[r0 + lane] : (-2 - lane);
            unsigned same = 0;                                            // bit s: lane-(1<<s) holds the same document
            int src[5];                                                   // scan source lane per step (self when not the same doc)
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const int od = __shfl_up_sync(0xffffffffu, my_doc, 1 << s);
                const bool sm = lane >= (1 << s) && od == my_doc;
                if (sm) same |= 1u << s;
                src[s] = sm ? lane - (1 << s) : lane;
            }
            const int nd = __shfl_down_sync(0xffffffffu, my_doc, 1);
            const bool is_tail = valid && (lane == 31 || nd != my_doc);   // groups hold whole documents
            unsigned step_mask = 0;                                       // scan steps some lane needs (uniform)
#pragma unroll
            for (int s = 0; s < 5; ++s)
                if (__any_sync(0xffffffffu, (same >> s) & 1u)) step_mask |= 1u << s;
            for (int c = lane; c < n_pad; c += 32) {                      // padded columns never pass
                float t = INFINITY;
                if (c < g.n_real) {
                    const uint32_t tk = w.use_tau ? ld_relaxed_u32(&w.ts.tau[g.q0 + c]) : 0u;
                    t = tk ? key_to_float(tk) : -INFINITY;
                }
                my_tau[c] = t;
            }
            __syncwarp();
            mbarrier_wait(&s_tfull[buf], use & 1u);
            tcgen05_fence_after();
            for (int c0 = half * 32; c0 < ((g.debug & 1) ? 0 : g.n_real); c0 += 64) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * n_pad + c0), r);
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                // per-document max: segmented inclusive max-scan over the lanes, 32 columns in lock-step
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    if (step_mask & (1u << s)) {                          // uniform
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], __shfl_sync(0xffffffffu, v[i], src[s]));
                    }
                }
                uint32_t pmask = 0;                                       // bit i: this lane's value reaches tau[c0+i]
                const float4* tau4 = reinterpret_cast<const float4*>(my_tau + c0);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 t = tau4[i4];
                    if (v[4 * i4 + 0] >= t.x) pmask |= 1u << (4 * i4 + 0);
                    if (v[4 * i4 + 1] >= t.y) pmask |= 1u << (4 * i4 + 1);
                    if (v[4 * i4 + 2] >= t.z) pmask |= 1u << (4 * i4 + 2);
                    if (v[4 * i4 + 3] >= t.w) pmask |= 1u << (4 * i4 + 3);
                }
                pmask = is_tail ? pmask : 0u;                             // only the last row of a document emits
                uint32_t colmask = __reduce_or_sync(0xffffffffu, pmask);
                while (colmask) {                                         // stage (query, key) entries: no global round trip
                    const int i = __ffs(colmask) - 1;
                    colmask &= colmask - 1;
                    float vi = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) vi = (j == i) ? v[j] : vi;
                    vi += 0.0f;
                    const bool pass = (pmask >> i) & 1u;
                    const unsigned pm = __ballot_sync(0xffffffffu, pass);
                    const int n = __popc(pm);
                    if (staged + n > kGemmStage) flush();
                    if (pass) {
                        const int q = g.q0 + c0 + i;
                        const uint32_t key = float_to_key(vi);
                        const int e = staged + __popc(pm & lt_mask);
                        st_key[e] = make_key64(key, dx.doc_base + uint32_t(my_doc));
                        st_q[e] = uint16_t(q);
                        if (w.use_tau) tau_count(w.ts, q, key);           // fire-and-forget histogram update
                    }
                    staged += n;
                }
                __syncwarp();
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
