// N4 — BM25 index aggregation on the device.
//
// Replaces the aggregation half of BM25.build_index (indexer/bm25_indexer.py:181-211, 283-343): per-document term
// frequencies, document frequencies and total frequencies from TOKENISED documents.  Tokenisation (spaCy in the
// reference) and the term dictionary stay on the host; the float32 corpus statistics and the float32 log10 IDF
// (:130-147, :346-369) are V- and N-sized and are formed on the host from these outputs so that they stay
// bit-identical to the reference's.
//
// Input: documents as a CSR of term ids (doc_tok_off[n_docs+1], tok_term[T]).  One 64-bit key (term << 32 | doc) per
// token, radix-sorted (CUB), run-length encoded: a run is a posting (term, doc) with tf = run length, already in the
// (term, doc) order the CSR index wants.  term_off / total_freq are lower bounds of term boundaries in the unique /
// full key arrays.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace mse {

__global__ void build_keys_kernel(const int64_t* __restrict__ doc_tok_off, const int32_t* __restrict__ tok_term, int64_t n_docs,
                                  int64_t n_terms, uint64_t* __restrict__ keys, int32_t* __restrict__ bad) {
    const int64_t d = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per document
    if (d >= n_docs) return;
    const int64_t a = doc_tok_off[d], e = doc_tok_off[d + 1];
    if (e < a) { *bad = 1; return; }
    for (int64_t i = a + lane_id(); i < e; i += 32) {
        const int32_t t = tok_term[i];
        if (t < 0 || t >= n_terms) { *bad = 2; keys[i] = ~0ull; }
        else keys[i] = (uint64_t(uint32_t(t)) << 32) | uint64_t(uint32_t(d));
    }
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* __restrict__ a, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void build_split_kernel(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ counts, const int32_t* __restrict__ n_runs,
                                   int32_t* __restrict__ post_doc, int32_t* __restrict__ post_tf) {
    const int64_t n = *n_runs;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        post_doc[i] = int32_t(uint32_t(uniq[i]));
        post_tf[i] = counts[i];
    }
}

__global__ void build_offsets_kernel(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ n_runs, const uint64_t* __restrict__ keys,
                                     int64_t n_tokens, int64_t n_terms, int64_t* __restrict__ term_off, int64_t* __restrict__ total_freq) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t > n_terms) return;
    const uint64_t key = uint64_t(t) << 32;
    term_off[t] = lower_bound_u64(uniq, *n_runs, key);
    if (t < n_terms) total_freq[t] = lower_bound_u64(keys, n_tokens, key + (1ull << 32)) - lower_bound_u64(keys, n_tokens, key);
}

}  // namespace mse
