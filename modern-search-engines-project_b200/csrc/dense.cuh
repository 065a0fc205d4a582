// K3 — exhaustive dense scan for small query batches (GEMV regime, HBM-bound).
//
// Reconstructs the removed Retriever.quick_search (call sites search_api.py:60,87; the stored
// vectors are L2-normalised, indexer/indexer.py:165): score(chunk) = <q, e_chunk>, document
// score = max over its chunks, then top-k documents (selection: topk.cuh).
//
// Layout in HBM: emb bf16 [n_chunks][768] row-major (1536 B per row); doc_chunk_off int64[D+1];
// row_doc int32[n_chunks] (document of every row, 0.26 % extra traffic); tile_row int64[T+1]
// (doc-aligned tiles of ~128 rows, so a document never spans two warps).
//
// Each warp owns a tile.  A row is read as 3 x 128-bit loads per lane (coalesced 512 B per warp
// instruction, L1::no_allocate: every byte is used once), multiplied against query fragments
// held in registers (fp32 accumulate), reduced with shuffles, and lane (row % 32) keeps the row's
// score.  Every 32 rows the warp does a segmented max over the lanes (rows of a document are
// adjacent), carries a document that continues into the next 32 rows, and the lane holding the
// last row of a finished document emits (score, doc) — but only when the score reaches tau[q],
// the running lower bound of the final k-th best (same scheme as the BM25 kernel).  There is no
// per-document array in HBM and no atomic per row: the selection stage reads a short candidate
// list instead of every document.
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace mse {

constexpr int kDim = MSE_EMB_DIM;
constexpr int kRowBytes = kDim * 2;
constexpr int kScanThreads = 256;
constexpr int kScanTileRows = 128;      // target rows per warp tile (tiles are doc-aligned)
constexpr int kScanRowsInFlight = 4;

struct DenseDev {
    const __nv_bfloat16* emb;
    const int64_t* doc_chunk_off;
    const int32_t* row_doc;
    const float* row_sq;         // [n_chunks] sum of squares of every stored (bf16) row, formed at load by dense_row_sq_kernel in
                                 // exactly the order the rerank kernels used to accumulate it per fetch (same bits)
    const int64_t* tile_row;
    int64_t n_chunks, n_docs, n_tiles;
    uint32_t doc_base;
    int64_t chunk_base;
};

struct DenseWork {
    const float* q;              // [B][768]
    uint64_t* cand;              // [B * cap]
    int32_t* cand_count;         // [B]
    int32_t* overflow;           // [B]
    TauState ts;
    int32_t cap, use_tau;
    // batch-wide emission log of the tensor-core kernel (gemm.cuh): entries (query, key) appended with one
    // atomic per flush, split into the per-query lists above by gemm_bucket_kernel
    uint64_t* log_key;
    uint16_t* log_q;
    unsigned long long* log_count;
    int64_t log_cap;
    int32_t n_log_queries;
};

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

template <int QB>
__global__ void __launch_bounds__(kScanThreads)
dense_scan_kernel(DenseDev dx, DenseWork w, int q0) {
    const int lane = lane_id();
    const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    float qf[QB][24];                     // lane's 3 x 8 query elements per query
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) qf[b][j * 8 + e] = w.q[int64_t(q0 + b) * kDim + j * 256 + lane * 8 + e];

    for (int64_t tile = warp_global; tile < dx.n_tiles; tile += n_warps) {
        const int64_t r0 = dx.tile_row[tile], r1 = dx.tile_row[tile + 1];
        float tau_f[QB];
#pragma unroll
        for (int b = 0; b < QB; ++b) {
            const uint32_t tk = w.use_tau ? ld_relaxed_u32(&w.ts.tau[q0 + b]) : 0u;      // key 0 == no bound yet
            tau_f[b] = tk ? key_to_float(tk) : -INFINITY;
        }
        int carry_doc = -1;
        float carry_val[QB];
#pragma unroll
        for (int b = 0; b < QB; ++b) carry_val[b] = -INFINITY;
        int emitted = 0;

        for (int64_t g = r0; g < r1; g += 32) {
            const int nrow = (r1 - g) < 32 ? int(r1 - g) : 32;
            const int my_doc = lane < nrow ? dx.row_doc[g + lane] : (-2 - lane);
            const int next_doc = (g + 32 < r1) ? dx.row_doc[g + 32] : -1;        // uniform
            float my_score[QB];
#pragma unroll
            for (int b = 0; b < QB; ++b) my_score[b] = -INFINITY;
#pragma unroll 1
            for (int i = 0; i < nrow; i += kScanRowsInFlight) {
                uint4 v[kScanRowsInFlight][3];
#pragma unroll
                for (int k = 0; k < kScanRowsInFlight; ++k) {
                    const int64_t row = (g + i + k) < r1 ? (g + i + k) : (r1 - 1);
                    const uint4* p = reinterpret_cast<const uint4*>(dx.emb + row * kDim) + lane;
#pragma unroll
                    for (int j = 0; j < 3; ++j) v[k][j] = ldg_stream(p + j * 32);
                }
#pragma unroll
                for (int k = 0; k < kScanRowsInFlight; ++k) {
                    float acc[QB];
#pragma unroll
                    for (int b = 0; b < QB; ++b) acc[b] = 0.f;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        float f[8];
                        bf16x8_to_float(v[k][j], f);
#pragma unroll
                        for (int e = 0; e < 8; ++e)
#pragma unroll
                            for (int b = 0; b < QB; ++b) acc[b] = fmaf(f[e], qf[b][j * 8 + e], acc[b]);
                    }
#pragma unroll
                    for (int b = 0; b < QB; ++b) {
                        const float s = warp_sum(acc[b]);
                        if (lane == i + k) my_score[b] = s;          // rows beyond nrow land in lanes >= nrow (ignored)
                    }
                }
            }
            // ---- per-document max over adjacent lanes (segmented inclusive max-scan) --------------
#pragma unroll
            for (int b = 0; b < QB; ++b)
                if (my_doc == carry_doc) my_score[b] = fmaxf(my_score[b], carry_val[b]);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int od = __shfl_up_sync(0xffffffffu, my_doc, o);
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const float ov = __shfl_up_sync(0xffffffffu, my_score[b], o);
                    if (lane >= o && od == my_doc) my_score[b] = fmaxf(my_score[b], ov);
                }
            }
            const int nd = __shfl_down_sync(0xffffffffu, my_doc, 1);
            const bool is_tail = lane < nrow && (lane == nrow - 1 ? (my_doc != next_doc) : (nd != my_doc));
            carry_doc = __shfl_sync(0xffffffffu, my_doc, nrow - 1);
#pragma unroll
            for (int b = 0; b < QB; ++b) carry_val[b] = __shfl_sync(0xffffffffu, my_score[b], nrow - 1);
            // ---- emit finished documents whose score reaches the running bound -----------------------
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                const float v = my_score[b] + 0.0f;
                const bool pass = is_tail && v >= tau_f[b];
                const unsigned pm = __ballot_sync(0xffffffffu, pass);
                if (pm) {
                    const int q = q0 + b;
                    const int total = __popc(pm);
                    int slot0 = 0;
                    if (lane == 0) slot0 = atomicAdd(&w.cand_count[q], total);
                    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                    if (pass) {
                        const int slot = slot0 + __popc(pm & lt_mask);
                        if (slot < w.cap) {
                            const uint32_t key = float_to_key(v);
                            w.cand[int64_t(q) * w.cap + slot] = make_key64(key, dx.doc_base + uint32_t(my_doc));
                            if (w.use_tau) tau_count(w.ts, q, key);
                        } else {
                            w.overflow[q] = 1;
                        }
                    }
                    emitted += total;
                }
            }
        }
        if (w.use_tau && emitted > 0) {
#pragma unroll
            for (int b = 0; b < QB; ++b) tau_raise(w.ts, q0 + b);
        }
    }
}

// Squared norm of every row (load time): lane l holds elements (l + 32 j) * 8 .. + 8 of the row, j = 0..2, accumulates them
// in that order with FMAs and the warp adds the 32 partial sums by butterfly — the arithmetic the rerank kernels did per
// fetched row (24 FMAs + 5 shuffles per lane and row: a third of their instructions), so the cosine keeps its bits.
__global__ void __launch_bounds__(256)
dense_row_sq_kernel(const __nv_bfloat16* __restrict__ emb, float* __restrict__ row_sq, int64_t n_chunks) {
    const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
    for (int64_t r = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_chunks; r += warps) {
        const uint4* p = reinterpret_cast<const uint4*>(emb + r * kDim) + lane_id();
        float ee = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float f[8];
            bf16x8_to_float(ldg_stream(p + j * 32), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) ee = fmaf(f[e], f[e], ee);
        }
        ee = warp_sum(ee);
        if (lane_id() == 0) row_sq[r] = ee;
    }
}

// row -> document map, built once at load
__global__ void dense_row_doc_kernel(const int64_t* __restrict__ off, int32_t* __restrict__ row_doc, int64_t n_docs) {
    const int64_t d = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (d >= n_docs) return;
    for (int64_t r = off[d]; r < off[d + 1]; ++r) row_doc[r] = int32_t(d);
}

// fp32 -> bf16 conversion of the embedding table at load time
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(in + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<__nv_bfloat162*>(out + i) = a;
        *reinterpret_cast<__nv_bfloat162*>(out + i + 2) = b;
    } else {
        for (; i < n; ++i) out[i] = __float2bfloat16_rn(in[i]);
    }
}

}  // namespace mse
