// K3 — exhaustive dense scan for small query batches (GEMV regime, HBM-bound).
//
// Reconstructs the removed Retriever.quick_search (call sites search_api.py:60,87; the stored
// vectors are L2-normalised, indexer/indexer.py:165): score(chunk) = <q, e_chunk>, document
// score = max over its chunks, then top-k documents (selection: topk.cuh).
//
// Layout in HBM: emb bf16 [n_chunks][768] row-major (1536 B per row), doc_chunk_off int64[D+1];
// chunks of a document are contiguous rows.  Per query a uint32 array best[D] of
// order-preserving score keys (0 == no chunk) receives the per-document maxima.
//
// Each warp owns a contiguous tile of rows.  A row is read as 3 x 128-bit loads per lane
// (coalesced 512 B per warp instruction, L1::no_allocate: every byte is used once), multiplied
// against query fragments held in registers (fp32), reduced with shuffles; lane 0 walks the
// document boundaries of its tile and issues one atomicMax per (document, query) — documents
// straddling two tiles are merged by that atomic.
#pragma once
#include "common.cuh"

namespace mse {

constexpr int kDim = MSE_EMB_DIM;
constexpr int kRowBytes = kDim * 2;
constexpr int kScanThreads = 256;
constexpr int kScanTileRows = 64;       // rows per warp tile
constexpr int kScanRowsInFlight = 4;

struct DenseDev {
    const __nv_bfloat16* emb;
    const int64_t* doc_chunk_off;
    int64_t n_chunks, n_docs;
    uint32_t doc_base;
    int64_t chunk_base;
};

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// first doc whose chunk range ends after `row` (docs without chunks are skipped naturally)
__device__ __forceinline__ int64_t doc_of_row(const int64_t* __restrict__ off, int64_t n_docs, int64_t row) {
    int64_t lo = 0, hi = n_docs;          // find smallest d with off[d+1] > row
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid + 1] > row) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int QB>
__global__ void __launch_bounds__(kScanThreads)
dense_scan_kernel(DenseDev dx, const float* __restrict__ q, int q0, uint32_t* __restrict__ best) {
    const int lane = lane_id();
    const int64_t warp_global = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t n_tiles = (dx.n_chunks + kScanTileRows - 1) / kScanTileRows;

    float qf[QB][24];                     // lane's 3 x 8 query elements per query
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) qf[b][j * 8 + e] = q[int64_t(q0 + b) * kDim + j * 256 + lane * 8 + e];

    for (int64_t tile = warp_global; tile < n_tiles; tile += n_warps) {
        const int64_t r0 = tile * kScanTileRows;
        const int64_t r1 = (r0 + kScanTileRows) < dx.n_chunks ? (r0 + kScanTileRows) : dx.n_chunks;
        int64_t doc = 0, doc_end = 0;     // lane 0 only
        float run[QB];
        if (lane == 0) {
            doc = doc_of_row(dx.doc_chunk_off, dx.n_docs, r0);
            doc_end = dx.doc_chunk_off[doc + 1];
#pragma unroll
            for (int b = 0; b < QB; ++b) run[b] = -INFINITY;
        }
        for (int64_t r = r0; r < r1; r += kScanRowsInFlight) {
            uint4 v[kScanRowsInFlight][3];
#pragma unroll
            for (int i = 0; i < kScanRowsInFlight; ++i) {
                const int64_t row = (r + i) < r1 ? (r + i) : (r1 - 1);
                const uint4* p = reinterpret_cast<const uint4*>(dx.emb + row * kDim) + lane;
#pragma unroll
                for (int j = 0; j < 3; ++j) v[i][j] = ldg_stream(p + j * 32);
            }
#pragma unroll
            for (int i = 0; i < kScanRowsInFlight; ++i) {
                float acc[QB];
#pragma unroll
                for (int b = 0; b < QB; ++b) acc[b] = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float f[8];
                    bf16x8_to_float(v[i][j], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e)
#pragma unroll
                        for (int b = 0; b < QB; ++b) acc[b] = fmaf(f[e], qf[b][j * 8 + e], acc[b]);
                }
#pragma unroll
                for (int b = 0; b < QB; ++b) acc[b] = warp_sum(acc[b]);
                if (lane == 0 && (r + i) < r1) {
                    const int64_t row = r + i;
                    if (row >= doc_end) {                 // document finished: publish its maximum
#pragma unroll
                        for (int b = 0; b < QB; ++b) {
                            atomicMax(&best[int64_t(q0 + b) * dx.n_docs + doc], float_to_key(run[b] + 0.0f));
                            run[b] = -INFINITY;
                        }
                        do { ++doc; doc_end = dx.doc_chunk_off[doc + 1]; } while (row >= doc_end);
                    }
#pragma unroll
                    for (int b = 0; b < QB; ++b) run[b] = fmaxf(run[b], acc[b]);
                }
            }
        }
        if (lane == 0 && r1 > r0) {
#pragma unroll
            for (int b = 0; b < QB; ++b)
                atomicMax(&best[int64_t(q0 + b) * dx.n_docs + doc], float_to_key(run[b] + 0.0f));
        }
    }
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
