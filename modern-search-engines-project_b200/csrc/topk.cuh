// K2 — batched exact top-k selection (one CTA per query).
//
// Replaces the full sort + slice of the reference (indexer/bm25_indexer.py:484-485) and the
// sort_values of the reranker (reranker/reranker_api.py:372).  Elements are 64-bit keys
// (order-preserving score bits << 32 | ~doc), all distinct, so "largest k keys, descending" is
// exactly "score descending, ties by ascending doc id" — the order Python's stable sort yields
// over rows that arrive in ascending doc id.
//
// Algorithm per query: (1) one pass computes the count and the common bit prefix of the keys,
// (2) MSB-first radix select (<= 8-bit digits, shared-memory histogram) starting at the first
// bit where the keys differ, with early exit as soon as a digit bucket is taken whole,
// (3) the k survivors are compacted into shared memory and bitonic-sorted.  All traffic after
// the first pass hits L2 (the candidate lists were just written by the scoring kernel).
#pragma once
#include "common.cuh"

namespace mse {

#ifndef MSE_BM25_HIST_BITS
#define MSE_BM25_HIST_BITS 13
#endif
constexpr int kHistBits = MSE_BM25_HIST_BITS;    // sign + exponent + 4 mantissa bits of the score key: bins 6 % wide in score (32 KB per
                                                 // query).  Measured at C5: 12 bits 5789 candidates per query / 14.33 ms score kernel,
                                                 // 13 and 14 bits 5209 / 14.16 ms (the refresh cadence, not the bin width, limits it then)
constexpr int kHistBins = 1 << kHistBits;
constexpr int kHistShift = 32 - kHistBits;
// The dense scans bin finer (sign + exponent + 6 mantissa bits: bins 1.6 % wide in score, 128 KB per query): their
// scores crowd just above the k-th best (cosines of a large corpus), where 12.5 %-wide bins leave the bound so far below
// the true k-th score that most 32-row groups of the GEMM epilogue still hold a passing document.
constexpr int kDenseHistBits = 15;

// Running lower bound of the final k-th best score of every query (see bm25.cuh / dense.cuh):
// producers count each emitted candidate in hist[q][score bin]; tau[q] is the largest bin edge with
// >= top_k emitted candidates at or above it.  Any value tau ever took is a valid bound.
struct TauState {
    uint32_t* tau;      // [B] score key
    uint32_t* hist;     // [B << (32 - shift)]
    uint32_t* maxbin;   // [B]
    int32_t top_k;
    int32_t shift;      // score-key bits dropped by a bin (kHistShift for BM25)
};

__device__ __forceinline__ void tau_count(const TauState& ts, int q, uint32_t key) {
    const uint32_t bin = key >> ts.shift;
    atomicAdd(&ts.hist[(int64_t(q) << (32 - ts.shift)) + bin], 1u);
    atomicMax(&ts.maxbin[q], bin);
}

// one full warp; continues the scan of query q's histogram at bin b with `acc` candidates already counted above it
__device__ __forceinline__ void tau_raise_from(const TauState& ts, int q, int b, int acc, int cur_bin) {
    const int lane = lane_id();
    const uint32_t* h = ts.hist + (int64_t(q) << (32 - ts.shift));
    while (b >= cur_bin) {
        const int bin = b - lane;
        const int c = (bin >= cur_bin && bin >= 0) ? int(ld_relaxed_u32(&h[bin])) : 0;
        const int incl = warp_incl_scan(c);
        const unsigned hit = __ballot_sync(0xffffffffu, acc + incl >= ts.top_k);
        if (hit) {
            const int tb = b - (__ffs(hit) - 1);
            if (lane == 0 && tb > cur_bin) atomicMax(&ts.tau[q], uint32_t(tb) << ts.shift);
            return;
        }
        acc += __shfl_sync(0xffffffffu, incl, 31);
        b -= 32;
    }
}
__device__ __forceinline__ void tau_raise(const TauState& ts, int q) {
    const int cur_bin = int(ld_relaxed_u32(&ts.tau[q]) >> ts.shift);
    tau_raise_from(ts, q, int(ld_relaxed_u32(&ts.maxbin[q])), 0, cur_bin);
}
// NQ queries at once: the dependent loads (bound and top bin, then the first 32 bins) of all of them are in flight
// together; a query whose k-th candidate lies further down continues alone
template <int NQ>
__device__ __forceinline__ void tau_raise_multi(const TauState& ts, const int (&q)[NQ]) {
    const int lane = lane_id();
    int cur_bin[NQ], b[NQ], c[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) { cur_bin[i] = int(ld_relaxed_u32(&ts.tau[q[i]]) >> ts.shift); b[i] = int(ld_relaxed_u32(&ts.maxbin[q[i]])); }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int bin = b[i] - lane;
        c[i] = (bin >= cur_bin[i] && bin >= 0) ? int(ld_relaxed_u32(&ts.hist[(int64_t(q[i]) << (32 - ts.shift)) + bin])) : 0;
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        if (b[i] < cur_bin[i]) continue;                                  // uniform
        const int incl = warp_incl_scan(c[i]);
        const unsigned hit = __ballot_sync(0xffffffffu, incl >= ts.top_k);
        if (hit) {
            const int tb = b[i] - (__ffs(hit) - 1);
            if (lane == 0 && tb > cur_bin[i]) atomicMax(&ts.tau[q[i]], uint32_t(tb) << ts.shift);
        } else {
            tau_raise_from(ts, q[i], b[i] - 32, __shfl_sync(0xffffffffu, incl, 31), cur_bin[i]);
        }
    }
}

constexpr int kSelectThreads = 512;
constexpr int kSortCap = MSE_MAX_TOPK;   // survivors sorted in shared memory

struct ListLoader {          // candidate lists written by the scoring kernels
    const uint64_t* keys;
    const int32_t* count;
    int64_t stride;
    int32_t cap;
    __device__ __forceinline__ int64_t n(int q) const { int c = count[q]; return c < cap ? c : cap; }
    __device__ __forceinline__ uint64_t get(int q, int64_t i) const { return keys[int64_t(q) * stride + i]; }
};

struct DenseLoader {         // per-doc best score keys (0 == document has no chunk)
    const uint32_t* best;
    int64_t n_docs;
    uint32_t doc_base;
    __device__ __forceinline__ int64_t n(int) const { return n_docs; }
    __device__ __forceinline__ uint64_t get(int q, int64_t i) const {
        uint32_t k = best[int64_t(q) * n_docs + i];
        return k ? make_key64(k, doc_base + uint32_t(i)) : 0ull;
    }
};

struct MergeLoader {         // all-gathered per-rank top-k lists, [n_lists][n_queries][list_k]
    const int32_t* doc;
    const float* score;
    const int32_t* count;
    int32_t n_lists, n_queries, list_k;
    __device__ __forceinline__ int64_t n(int) const { return int64_t(n_lists) * list_k; }
    __device__ __forceinline__ uint64_t get(int q, int64_t i) const {
        int l = int(i / list_k), j = int(i % list_k);
        int64_t row = int64_t(l) * n_queries + q;
        if (j >= count[row]) return 0ull;
        float s = score[row * list_k + j] + 0.0f;
        return make_key64(float_to_key(s), uint32_t(doc[row * list_k + j]));
    }
};

struct KeyListLoader {       // per-shard lists of 64-bit keys exchanged between ranks, [n_lists][n_queries][list_k], 0 = empty slot
    const uint64_t* keys;
    int32_t n_lists, n_queries, list_k;
    __device__ __forceinline__ int64_t n(int) const { return int64_t(n_lists) * list_k; }
    __device__ __forceinline__ uint64_t get(int q, int64_t i) const {
        const int l = int(i / list_k), j = int(i % list_k);
        return keys[(int64_t(l) * n_queries + q) * list_k + j];
    }
};

// Outputs: (out_doc, out_score, out_count) as the C ABI returns them and / or out_key, the same list as 64-bit keys
// (0-padded) — the form the shard exchange moves.  Any of the two may be null.
template <class Loader>
__global__ void __launch_bounds__(kSelectThreads, 2)
topk_select_kernel(Loader ld, int32_t top_k, int32_t* __restrict__ out_doc, float* __restrict__ out_score,
                   int32_t* __restrict__ out_count, const int32_t* __restrict__ skip_flag,
                   unsigned long long* __restrict__ summary = nullptr,   // [0] += elements selected from, [1] += skipped queries
                   uint64_t* __restrict__ out_key = nullptr) {
    constexpr int NT = kSelectThreads;
    __shared__ uint64_t s_buf[kSortCap];
    __shared__ int s_hist[256];
    __shared__ uint64_t s_or[NT / 32], s_and[NT / 32];
    __shared__ int s_cnt[NT / 32];
    __shared__ int s_digit, s_above, s_cntd, s_n;

    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    if (skip_flag && skip_flag[q]) {            // candidate list overflowed: the query is reported (count -1, no entries)
        if (tid == 0) {
            if (out_count) out_count[q] = -1;
            if (summary) atomicAdd(summary + 1, 1ull);
        }
        for (int i = tid; i < top_k; i += NT) {
            if (out_doc) { out_doc[int64_t(q) * top_k + i] = -1; out_score[int64_t(q) * top_k + i] = 0.f; }
            if (out_key) out_key[int64_t(q) * top_k + i] = 0ull;
        }
        return;
    }
    const int64_t n = ld.n(q);
    if (summary && tid == 0) atomicAdd(summary, (unsigned long long)n);
    // Lists that fit the sort buffer (the common case once the scoring kernels filter by a running bound) are
    // copied to shared memory once; every later pass reads them from there instead of from L2.
    const bool in_smem = n <= kSortCap;
    auto key_at = [&](int64_t i) -> uint64_t { return in_smem ? s_buf[i] : ld.get(q, i); };

    // ---- pass A: count valid keys, common prefix ------------------------------------------
    uint64_t vor = 0, vand = ~0ull;
    int cnt = 0;
    for (int64_t i = tid; i < n; i += NT) {
        uint64_t k = ld.get(q, i);
        if (in_smem) s_buf[i] = k;
        if (k) { vor |= k; vand &= k; ++cnt; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vor |= __shfl_xor_sync(0xffffffffu, vor, o);
        vand &= __shfl_xor_sync(0xffffffffu, vand, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane_id() == 0) { s_or[warp_id()] = vor; s_and[warp_id()] = vand; s_cnt[warp_id()] = cnt; }
    if (tid == 0) s_n = 0;
    __syncthreads();
    vor = 0; vand = ~0ull; int n_valid = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { vor |= s_or[w]; vand &= s_and[w]; n_valid += s_cnt[w]; }

    const int kk = n_valid < top_k ? n_valid : top_k;
    if (kk == 0) {
        if (tid == 0 && out_count) out_count[q] = 0;
        for (int i = tid; i < top_k; i += NT) {
            if (out_doc) { out_doc[int64_t(q) * top_k + i] = -1; out_score[int64_t(q) * top_k + i] = 0.f; }
            if (out_key) out_key[int64_t(q) * top_k + i] = 0ull;
        }
        return;
    }

    // ---- radix select of the kk-th largest key ---------------------------------------------
    int sel_shift = 0;            // selection criterion: (key >> sel_shift) >= (prefix >> sel_shift)
    uint64_t prefix = 0;
    bool select_all = (n_valid <= kk);
    if (!select_all) {
        const uint64_t diff = vor ^ vand;                 // != 0: at least two distinct keys
        int rem = 64 - __clzll((long long)diff);          // number of low bits still undecided
        prefix = (rem >= 64) ? 0ull : (vand >> rem) << rem;
        int k_rem = kk;
        while (true) {
            const int width = rem < 8 ? rem : 8;
            const int shift = rem - width;
            const int hi = shift + width;                 // bits >= hi are decided (== prefix)
            if (tid < 256) s_hist[tid] = 0;
            __syncthreads();
            for (int64_t i = tid; i < n; i += NT) {
                uint64_t k = key_at(i);
                if (!k) continue;
                bool match = (hi >= 64) ? true : (((k ^ prefix) >> hi) == 0);
                if (match) atomicAdd(&s_hist[int(k >> shift) & ((1 << width) - 1)], 1);
            }
            __syncthreads();
            if (tid < 32) {                               // descending suffix scan over 256 bins
                int c[8], lane_sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = s_hist[255 - (8 * tid + j)]; lane_sum += c[j]; }
                int incl = warp_incl_scan(lane_sum);
                int run = incl - lane_sum;
                if (run < k_rem && k_rem <= incl) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (run < k_rem && k_rem <= run + c[j]) { s_digit = 255 - (8 * tid + j); s_above = run; s_cntd = c[j]; }
                        run += c[j];
                    }
                }
            }
            __syncthreads();
            const int d = s_digit, above = s_above, cntd = s_cntd;
            __syncthreads();
            k_rem -= above;
            prefix |= uint64_t(d) << shift;
            rem = shift;
            if (cntd == k_rem || rem == 0) { sel_shift = shift; break; }
        }
    }

    // ---- compaction of the survivors ---------------------------------------------------------
    if (in_smem) {
        // in place: every thread keeps the survivors of its contiguous slice in registers, a block-wide scan of
        // the counts gives their destination, and the writes start only after everybody has read
        constexpr int C = (kSortCap + NT - 1) / NT;
        uint64_t keep[C];
        int mine = 0;
        const int per = int((n + NT - 1) / NT);
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const int i = tid * per + j;
            uint64_t k = (j < per && i < n) ? s_buf[i] : 0ull;
            const bool sel = k != 0ull && (select_all || (k >> sel_shift) >= (prefix >> sel_shift));
            keep[j] = sel ? k : 0ull;
            mine += sel ? 1 : 0;
        }
        const int incl = warp_incl_scan(mine);
        if (lane_id() == 31) s_cnt[warp_id()] = incl;
        __syncthreads();
        int base = incl - mine;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) if (w < warp_id()) base += s_cnt[w];
        if (tid == NT - 1) s_n = base + mine;
#pragma unroll
        for (int j = 0; j < C; ++j) if (keep[j] != 0ull) { if (base < kSortCap) s_buf[base] = keep[j]; ++base; }
    } else {
        for (int64_t i = tid; i < n; i += NT) {
            uint64_t k = ld.get(q, i);
            if (!k) continue;
            if (select_all || (k >> sel_shift) >= (prefix >> sel_shift)) {
                int slot = atomicAdd(&s_n, 1);
                if (slot < kSortCap) s_buf[slot] = k;
            }
        }
    }
    __syncthreads();
    int m = s_n < kSortCap ? s_n : kSortCap;
    int P = 1;
    while (P < m) P <<= 1;
    for (int i = m + tid; i < P; i += NT) s_buf[i] = 0ull;
    __syncthreads();

    // ---- bitonic sort, descending --------------------------------------------------------------
    // With compare distance j <= 32 the pairs a warp handles stay inside that warp's own 64-key segments
    // (pair index t -> positions in [64 * (t / 32), 64 * (t / 32) + 63]), so those steps only need __syncwarp;
    // a block barrier is needed after the steps with j >= 64 and between merges.
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int lj = 31 - __clz(j);                                  // j is a power of two
            for (int t = tid; t < (P >> 1); t += NT) {
                const int i = ((t >> lj) << (lj + 1)) | (t & (j - 1));
                const int p = i + j;
                const uint64_t a = s_buf[i], b = s_buf[p];
                const bool up = ((i & k) == 0);
                if ((a < b) == up) { s_buf[i] = b; s_buf[p] = a; }
            }
            if (j > 32) __syncthreads(); else __syncwarp();
        }
        __syncthreads();
    }
    const int out_n = m < kk ? m : kk;
    for (int i = tid; i < top_k; i += NT) {
        int64_t o = int64_t(q) * top_k + i;
        const uint64_t k = i < out_n ? s_buf[i] : 0ull;
        if (out_key) out_key[o] = k;
        if (out_doc) {
            out_doc[o] = i < out_n ? int32_t(key64_doc(k)) : -1;
            out_score[o] = i < out_n ? key_to_float(key64_score_key(k)) : 0.f;
        }
    }
    if (tid == 0 && out_count) out_count[q] = out_n;
}

// Exactness check of a shard exchange that moved only the first list_k entries of every shard list (see sharding notes
// in api.cu): a list that came back full (its last slot is used) may have been cut; if its last score is still >= the
// merged k-th score, or the merge is short, it could hide a result of the exact top-k.  One thread per (list, query).
__global__ void shard_cut_check_kernel(const uint64_t* __restrict__ keys, int32_t n_lists, int32_t n_queries, int32_t list_k,
                                       const uint64_t* __restrict__ merged_key, int32_t top_k, int32_t* __restrict__ status) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= int64_t(n_lists) * n_queries) return;
    const int q = int(i % n_queries);
    const uint64_t last = keys[i * list_k + (list_k - 1)];
    if (last == 0ull) return;                                   // the shard sent everything it had
    const uint64_t kth = merged_key[int64_t(q) * top_k + (top_k - 1)];
    if (kth == 0ull || key64_score_key(last) >= key64_score_key(kth)) atomicOr(status + 2, 1);
}

}  // namespace mse
