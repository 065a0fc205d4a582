// K1 — batched BM25 scoring over a CSR inverted index resident in HBM.
//
// Replaces the candidate SQL, the dict grouping and the float64 scoring loop of
// BM25.search (indexer/bm25_indexer.py:435-481):
//     score(d) = sum_t idf_t * tf*(k1+1) / (tf + k1*(1 - b + b*dl_d/avgdl)) * qtf_t
// over the documents that hold at least one posting of a valid query term, kept when
// score >= min_score.
//
// Layout in HBM: term_off int64[V+1]; postings as two int32 arrays (doc, tf), ascending doc
// inside a term; doc_norm float32[N] = k1*(1-b+b*dl/avgdl) (computed in double from the
// float32 avgdl); idf float32[V] verbatim from bm25_term_stats.
//
// Execution model (doc-range tiling, no HBM accumulators):
//   * the doc space is cut into ranges of R docs; a (range, query-chunk) pair is one work item;
//     persistent CTAs pull items from an atomic counter in range-major order, so consecutive
//     items reuse the range's doc_norm slice already staged in shared memory;
//   * bm25_prepare_kernel binary-searches, once per (query term, range boundary), where each
//     posting list crosses each range boundary; the scoring kernel then streams exactly the
//     postings of its range with coalesced loads (no search on the critical path);
//   * accumulators live in shared memory (fp32[R]); doc ids are unique inside a term, so a term
//     is applied with plain read-modify-write and terms are separated by a CTA barrier —
//     no atomics, deterministic summation order (= the reference's term order);
//   * "touched" is tracked without a bitmap: accumulators start as -0.0f (x + -0.0 == x,
//     +0.0 + -0.0 == +0.0), so a touched document whose score is exactly zero (idf == 0, kept by
//     the reference) is distinguishable from an untouched one (never returned);
//   * the scan that reads out a range re-arms the accumulators and emits only candidates whose
//     score key is >= tau[q], a per-query lower bound of the final k-th best score.  tau is
//     raised on the fly from a per-query histogram of the candidates emitted so far (any value
//     tau ever took is a valid bound, so no inter-CTA synchronisation is needed); with
//     range-major scheduling the ranges of one query are visited roughly in sequence and the
//     emitted volume is ~k*ln(candidates/k) instead of every candidate.
#pragma once
#include "common.cuh"

namespace mse {

constexpr int kBm25Threads = 256;
constexpr int kBm25CtasPerSm = 4;
constexpr int kHistBits = 12;                    // sign + exponent + 3 mantissa bits
constexpr int kHistBins = 1 << kHistBits;
constexpr int kHistShift = 32 - kHistBits;

struct Bm25Dev {                                 // device-resident index of one shard
    const int64_t* term_off;
    const int32_t* post_doc;
    const int32_t* post_tf;
    const float* doc_norm;
    const float* idf;
    int64_t n_terms, n_docs, n_postings;
    uint32_t doc_base;
    float k1;
};

struct Bm25Work {                                // per-call workspace
    const int32_t* q_off;
    const int32_t* q_term;
    const int32_t* q_tf;
    float* slot_w;               // [S]  idf * qtf * (k1+1)
    int64_t* slot_base;          // [S]  term_off[term]
    uint32_t* seg;               // [S * (n_ranges+1)] posting offset (relative to slot_base) of each range boundary
    uint32_t* tau;               // [B]  lower bound (score key) of the final k-th best
    uint32_t* hist;              // [B * kHistBins] emitted candidates per score bin
    uint32_t* maxbin;            // [B]
    uint64_t* cand;              // [B * cap]
    int32_t* cand_count;         // [B]
    int32_t* overflow;           // [B] 1 when more than cap candidates were emitted
    int32_t* item_counter;       // [1]
    unsigned long long* stats;   // [0] postings traversed
    int32_t n_queries, n_slots, n_ranges, range_docs, queries_per_item, cap, top_k;
    uint32_t min_key;
    int32_t use_tau;
};

// ---- prepare: slot weights, list bases, range boundaries, tau init ---------------------------
__global__ void bm25_prepare_kernel(Bm25Dev ix, Bm25Work w) {
    const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int nb = w.n_ranges + 1;
    if (gid < w.n_queries) w.tau[gid] = w.min_key;
    if (gid >= int64_t(w.n_slots) * nb) return;
    const int s = int(gid / nb), j = int(gid % nb);
    const int t = w.q_term[s];
    int64_t a = 0, e = 0;
    if (t >= 0 && t < ix.n_terms) { a = ix.term_off[t]; e = ix.term_off[t + 1]; }
    if (j == 0) {
        float idf = (t >= 0 && t < ix.n_terms) ? ix.idf[t] : 0.f;
        // idf * qtf * (k1+1) formed in double, rounded once (reference: float64 throughout)
        w.slot_w[s] = float(double(idf) * double(w.q_tf[s]) * (double(ix.k1) + 1.0)) + 0.0f;
        w.slot_base[s] = a;
        if (e > a) atomicAdd(w.stats, (unsigned long long)(e - a));
    }
    // first posting with doc >= j * R
    const int64_t target = int64_t(j) * w.range_docs;
    int64_t lo = a, hi = e;
    if (j == w.n_ranges) lo = e;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (int64_t(ix.post_doc[mid]) < target) lo = mid + 1; else hi = mid;
    }
    w.seg[gid] = uint32_t(lo - a);
}

// ---- tau update: largest bin edge with >= k emitted candidates at or above it ------------------
__device__ __forceinline__ void bm25_raise_tau(const Bm25Work& w, int q) {
    const int lane = lane_id();
    const uint32_t cur = ld_relaxed_u32(&w.tau[q]);
    const int cur_bin = int(cur >> kHistShift);
    int b = int(ld_relaxed_u32(&w.maxbin[q]));
    int acc = 0;
    const uint32_t* h = w.hist + int64_t(q) * kHistBins;
    while (b >= cur_bin) {
        const int bin = b - lane;
        int c = (bin >= cur_bin && bin >= 0) ? int(ld_relaxed_u32(&h[bin])) : 0;
        int incl = warp_incl_scan(c);
        unsigned hit = __ballot_sync(0xffffffffu, acc + incl >= w.top_k);
        if (hit) {
            const int first = __ffs(hit) - 1;
            const int tb = b - first;
            if (lane == 0 && tb > cur_bin) atomicMax(&w.tau[q], uint32_t(tb) << kHistShift);
            return;
        }
        acc += __shfl_sync(0xffffffffu, incl, 31);
        b -= 32;
    }
}

// ---- scoring ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBm25Threads, kBm25CtasPerSm)
bm25_score_kernel(Bm25Dev ix, Bm25Work w) {
    constexpr int NT = kBm25Threads;
    extern __shared__ __align__(16) float smem[];
    float* s_norm = smem;
    float* s_acc = smem + w.range_docs;
    __shared__ int s_item;
    __shared__ int s_emit;

    const int tid = threadIdx.x;
    const int R = w.range_docs;
    const int nb = w.n_ranges + 1;
    const int chunks = (w.n_queries + w.queries_per_item - 1) / w.queries_per_item;
    const int n_items = w.n_ranges * chunks;
    const float neg0 = __uint_as_float(kUntouchedBits);

    for (int i = tid; i < R; i += NT) s_acc[i] = neg0;
    int cur_r = -1;

    while (true) {
        __syncthreads();                                   // previous item fully retired (s_item reuse)
        if (tid == 0) s_item = atomicAdd(w.item_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        const int r = item / chunks, c = item % chunks;
        const int lo = r * R;
        const int nd = (ix.n_docs - lo) < R ? int(ix.n_docs - lo) : R;
        if (r != cur_r) {                                  // stage this range's doc norms
            for (int i = tid; i < nd; i += NT) s_norm[i] = ix.doc_norm[lo + i];
            cur_r = r;
        }
        const int q0 = c * w.queries_per_item;
        const int q1 = (q0 + w.queries_per_item) < w.n_queries ? (q0 + w.queries_per_item) : w.n_queries;
        for (int q = q0; q < q1; ++q) {
            const int s0 = w.q_off[q], s1 = w.q_off[q + 1];
            if (tid == 0) s_emit = 0;
            __syncthreads();                               // norms staged / previous scan done
            bool any = false;
            for (int s = s0; s < s1; ++s) {
                const uint32_t sb = w.seg[int64_t(s) * nb + r], se = w.seg[int64_t(s) * nb + r + 1];
                if (se > sb) {
                    any = true;
                    const float wt = w.slot_w[s];
                    const int64_t base = w.slot_base[s];
                    const int32_t* __restrict__ pd = ix.post_doc + base;
                    const int32_t* __restrict__ pt = ix.post_tf + base;
                    for (uint32_t i = sb + tid; i < se; i += NT) {
                        const int d = ldg_stream_i32(pd + i) - lo;
                        const float tf = float(ldg_stream_i32(pt + i));
                        // idf*qtf*(k1+1) * tf / (tf + k1*(1-b+b*dl/avgdl))
                        const float contrib = (wt * tf) / (tf + s_norm[d]);
                        s_acc[d] += contrib;               // docs are unique inside a term: no race
                    }
                    __syncthreads();                       // next term may touch the same docs
                }
            }
            if (!any) continue;                            // uniform: nothing touched in this range
            // ---- read-out scan: emit candidates >= tau, re-arm accumulators --------------------
            const uint32_t tau = w.use_tau ? ld_relaxed_u32(&w.tau[q]) : w.min_key;
            float4* acc4 = reinterpret_cast<float4*>(s_acc);
            const float4 rearm = make_float4(neg0, neg0, neg0, neg0);
            const int n4 = (nd + 3) >> 2;
            for (int j4 = tid; j4 < ((n4 + NT - 1) / NT) * NT; j4 += NT) {
                uint32_t key[4];
                int nem = 0;
                if (j4 < n4) {
                    float4 v = acc4[j4];
                    acc4[j4] = rearm;
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool touched = __float_as_uint(vv[u]) != kUntouchedBits;
                        const uint32_t k = float_to_key(vv[u] + 0.0f);
                        key[u] = (touched && k >= tau) ? k : 0u;
                        nem += key[u] != 0u;
                    }
                }
                if (__ballot_sync(0xffffffffu, nem > 0) == 0u) continue;
                const int incl = warp_incl_scan(nem);
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                int base_slot = 0;
                if (lane_id() == 31) {
                    base_slot = atomicAdd(&w.cand_count[q], total);
                    atomicAdd(&s_emit, total);
                }
                base_slot = __shfl_sync(0xffffffffu, base_slot, 31);
                int slot = base_slot + incl - nem;
                if (nem) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (key[u]) {
                            if (slot < w.cap) {
                                w.cand[int64_t(q) * w.cap + slot] =
                                    make_key64(key[u], ix.doc_base + uint32_t(lo + 4 * j4 + u));
                                if (w.use_tau) {
                                    const uint32_t bin = key[u] >> kHistShift;
                                    atomicAdd(&w.hist[int64_t(q) * kHistBins + bin], 1u);
                                    atomicMax(&w.maxbin[q], bin);
                                }
                            } else {
                                w.overflow[q] = 1;
                            }
                            ++slot;
                        }
                    }
                }
            }
            __syncthreads();                               // accumulators re-armed; s_emit final
            if (w.use_tau && warp_id() == 0 && s_emit > 0) {
                __threadfence();                           // histogram updates of this CTA are visible
                bm25_raise_tau(w, q);
            }
        }
    }
}

// ---- load-time kernels --------------------------------------------------------------------------
__global__ void bm25_norm_kernel(const int32_t* __restrict__ doc_len, float* __restrict__ norm, int64_t n,
                                 double k1, double b, double avgdl) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) norm[i] = float(k1 * (1.0 - b + b * double(doc_len[i]) / avgdl));
}

// postings must be strictly ascending inside a term and inside [0, n_docs); tf >= 1
__global__ void bm25_validate_kernel(const int64_t* __restrict__ term_off, const int32_t* __restrict__ post_doc,
                                     const int32_t* __restrict__ post_tf, int64_t n_terms, int64_t n_docs,
                                     int32_t* __restrict__ bad) {
    const int64_t t = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_terms) return;
    const int64_t a = term_off[t], e = term_off[t + 1];
    if (e < a) { *bad = 1; return; }
    for (int64_t i = a + lane_id(); i < e; i += 32) {
        const int d = post_doc[i];
        if (d < 0 || d >= n_docs || post_tf[i] < 1) *bad = 2;
        if (i > a && post_doc[i - 1] >= d) *bad = 3;
    }
}

}  // namespace mse
