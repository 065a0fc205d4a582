// K5s — gathered rerank when the chunk table is sharded by document range over several GPUs.
//
// The reference normalises cosines with the min / max over the WHOLE candidate pool
// (reranker/reranker_api.py:289-296,360-361), so a rank that owns only part of the candidates cannot
// finish alone.  The work is split in two kernels around one exchange step:
//   rerank_shard_cos_kernel   every rank, same replicated candidate list: sort + URL-group dedupe (identical
//                             on all ranks), then cosine of the <= max_chunks rows of the candidates THIS rank
//                             owns, written into zero-initialised arrays indexed by (query, survivor slot, row)
//   -- all-reduce(sum) of those arrays over the ranks (each slot is owned by exactly one rank, so the sum
//      is a gather; B x 1024 x 10 floats) --
//   rerank_shard_fuse_kernel  pool-wide min-max, 0.85/0.15 fusion, positional weighting, per-doc max, sort
//                             (:299-372) from the gathered cosines — no embedding access, same result on
//                             every rank.
// On one GPU the fused rerank_kernel (rerank.cuh) does all of this in one pass.
#pragma once
#include "common.cuh"
#include "dense.cuh"
#include "rerank.cuh"

namespace mse {

struct RerankShardArgs {
    const int32_t* cand_off;
    const int32_t* cand_doc;        // GLOBAL dense doc index, BM25 order (replicated on all ranks)
    const float* cand_bm25;
    const int32_t* url_group;       // [n_docs_global] or null (replicated)
    const float* q;                 // [B][768]
    int32_t max_chunks;
    int64_t n_docs_global;
    // exchange arrays (zero-initialised by the caller before the cos kernel)
    float* cos;                     // [B][kRerankMaxCand][kRerankMaxChunks]
    int32_t* rows;                  // [B][kRerankMaxCand]
    int64_t* chunk0;                // [B][kRerankMaxCand] global row of the doc's first chunk
    // replicated survivor description (every rank writes identical values)
    int32_t* surv_doc;              // [B][kRerankMaxCand]
    float* surv_bm25;               // [B][kRerankMaxCand]
    int32_t* surv_count;            // [B]
};

__global__ void __launch_bounds__(kRerankThreads)
rerank_shard_cos_kernel(DenseDev dx, RerankShardArgs a) {
    constexpr int NT = kRerankThreads;
    __shared__ uint64_t s_key[kRerankMaxCand];
    __shared__ uint64_t s_key2[kRerankMaxCand];
    __shared__ uint8_t s_dup[kRerankMaxCand];
    __shared__ int32_t s_doc[kRerankMaxCand];
    __shared__ int s_ns;
    const int qi = blockIdx.x;
    const int tid = threadIdx.x;
    const int c0 = a.cand_off[qi];
    int nc = a.cand_off[qi + 1] - c0;
    if (nc > kRerankMaxCand) nc = kRerankMaxCand;
    int P = 1;
    while (P < nc) P <<= 1;
    // ascending (doc, input slot); docs outside [0, n_docs_global) are dropped (-1 padding of BM25 results)
    for (int i = tid; i < P; i += NT) {
        uint64_t k = 0;
        if (i < nc) {
            const int64_t g = a.cand_doc[c0 + i];
            if (g >= 0 && g < a.n_docs_global)
                k = (1ull << 63) | (uint64_t(0x7fffffffu - uint32_t(g)) << 10) | uint64_t(0x3ff - i);
        }
        s_key[i] = k;
    }
    if (tid == 0) s_ns = 0;
    __syncthreads();
    block_bitonic_desc<NT>(s_key, P);
    for (int i = tid; i < P; i += NT) {
        const uint64_t k = s_key[i];
        uint64_t k2 = 0;
        s_dup[i] = 0;
        if (k >> 63) {
            const uint32_t g = 0x7fffffffu - uint32_t((k >> 10) & 0x7fffffffu);
            const uint32_t grp = a.url_group ? uint32_t(a.url_group[g]) & 0x7fffffffu : g;
            k2 = (1ull << 63) | (uint64_t(0x7fffffffu - grp) << 10) | uint64_t(0x3ff - i);
        }
        s_key2[i] = k2;
    }
    __syncthreads();
    block_bitonic_desc<NT>(s_key2, P);
    for (int j = tid; j < P; j += NT) {
        const uint64_t k = s_key2[j];
        if ((k >> 63) && j > 0) {
            const uint64_t p = s_key2[j - 1];
            if ((p >> 63) && ((p >> 10) == (k >> 10))) s_dup[0x3ff - int(k & 0x3ffull)] = 1;
        }
    }
    __syncthreads();
    if (tid < 32) {                                        // stable compaction (one warp)
        int carry = 0;
        for (int base = 0; base < P; base += 32) {
            const int i = base + tid;
            const uint64_t k = i < P ? s_key[i] : 0ull;
            const int keep = ((k >> 63) && !s_dup[i]) ? 1 : 0;
            const int incl = warp_incl_scan(keep);
            __syncwarp();
            if (keep) {
                const int o = carry + incl - 1;
                const int32_t g = int32_t(0x7fffffffu - uint32_t((k >> 10) & 0x7fffffffu));
                s_doc[o] = g;
                if (blockIdx.y == 0) {                     // (every slice of a query computes the same list)
                    a.surv_doc[int64_t(qi) * kRerankMaxCand + o] = g;
                    a.surv_bm25[int64_t(qi) * kRerankMaxCand + o] = a.cand_bm25[c0 + (0x3ff - int(k & 0x3ffull))];
                }
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) { s_ns = carry; if (blockIdx.y == 0) a.surv_count[qi] = carry; }
    }
    __syncthreads();
    const int ns = s_ns;
    const int max_chunks = a.max_chunks < kRerankMaxChunks ? a.max_chunks : kRerankMaxChunks;

    // cosines of the rows this rank owns: one warp per (survivor, chunk) row
    const float* qv = a.q + int64_t(qi) * kDim;
    float qf[24];
    float qq = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            qf[j * 8 + e] = qv[j * 256 + lane_id() * 8 + e];
            qq = fmaf(qf[j * 8 + e], qf[j * 8 + e], qq);
        }
    qq = warp_sum(qq);
    const float qn = sqrtf(qq);
    // gridDim.y > 1 (small batches on one GPU): the survivors of a query are dealt out to gridDim.y CTAs
    for (int i = int(blockIdx.y) * (NT / 32) + warp_id(); i < ns; i += int(gridDim.y) * (NT / 32)) {
        const int64_t d = int64_t(s_doc[i]) - int64_t(dx.doc_base);
        if (d < 0 || d >= dx.n_docs) continue;             // another rank owns this document
        const int64_t ra = dx.doc_chunk_off[d], re = dx.doc_chunk_off[d + 1];
        const int n = int(re - ra) < max_chunks ? int(re - ra) : max_chunks;
        const int64_t slot = int64_t(qi) * kRerankMaxCand + i;
        if (lane_id() == 0) { a.rows[slot] = n; a.chunk0[slot] = dx.chunk_base + ra; }
        for (int j = 0; j < n; ++j) {
            const uint4* p = reinterpret_cast<const uint4*>(dx.emb + (ra + j) * kDim) + lane_id();
            float dot = 0.f, ee = 0.f;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                float f[8];
                bf16x8_to_float(ldg_stream(p + t * 32), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) { dot = fmaf(f[e], qf[t * 8 + e], dot); ee = fmaf(f[e], f[e], ee); }
            }
            dot = warp_sum(dot);
            ee = warp_sum(ee);
            if (lane_id() == 0) a.cos[slot * kRerankMaxChunks + j] = dot / (sqrtf(ee) * qn);
        }
    }
}

struct RerankFuseArgs {
    const float* cos;               // gathered [B][kRerankMaxCand][kRerankMaxChunks]
    const int32_t* rows;
    const int64_t* chunk0;
    const int32_t* surv_doc;
    const float* surv_bm25;
    const int32_t* surv_count;
    float smoothing;
    int32_t max_out;
    int32_t* out_doc;
    float* out_score;
    float* out_orig;
    int64_t* out_chunk;
    int32_t* out_count;
    int32_t* out_rows;
};

__global__ void __launch_bounds__(kRerankThreads)
rerank_shard_fuse_kernel(RerankFuseArgs a) {
    constexpr int NT = kRerankThreads;
    __shared__ uint64_t s_key[kRerankMaxCand];
    __shared__ float s_score[kRerankMaxCand], s_orig[kRerankMaxCand];
    __shared__ uint8_t s_best[kRerankMaxCand];
    __shared__ float s_tmp[NT / 32];
    const int qi = blockIdx.x;
    const int tid = threadIdx.x;
    const int ns = a.surv_count[qi];
    const int64_t base = int64_t(qi) * kRerankMaxCand;
    float cmin = INFINITY, cmax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
    int T = 0;
    for (int i = tid; i < ns; i += NT) {
        const int n = a.rows[base + i];
        if (n > 0) {
            T += n;
            const float bm = a.surv_bm25[base + i];
            bmin = fminf(bmin, bm); bmax = fmaxf(bmax, bm);
            for (int j = 0; j < n; ++j) {
                const float c = a.cos[(base + i) * kRerankMaxChunks + j];
                cmin = fminf(cmin, c); cmax = fmaxf(cmax, c);
            }
        }
    }
    cmin = block_reduce<NT>(cmin, s_tmp, 2);
    cmax = block_reduce<NT>(cmax, s_tmp, 1);
    bmin = block_reduce<NT>(bmin, s_tmp, 2);
    bmax = block_reduce<NT>(bmax, s_tmp, 1);
    T = int(block_reduce<NT>(float(T), s_tmp, 0) + 0.5f);
    int P = 1;
    while (P < ns) P <<= 1;
    if (T == 0) {                                          // reference: HTTP 401 "No documents found"
        if (tid == 0) { a.out_count[qi] = 0; a.out_rows[qi] = 0; }
        for (int o = tid; o < a.max_out; o += NT) {
            const int64_t dst = int64_t(qi) * a.max_out + o;
            a.out_doc[dst] = -1; a.out_score[dst] = 0.f; a.out_orig[dst] = 0.f; a.out_chunk[dst] = -1;
        }
        return;
    }
    const double crange = double(cmax) - double(cmin), brange = double(bmax) - double(bmin);
    const double sm = double(a.smoothing);
    for (int i = tid; i < P; i += NT) {
        uint64_t key = 0;
        const int n = i < ns ? a.rows[base + i] : 0;
        if (n > 0) {
            const double oldn = brange == 0.0 ? 0.0 : (double(a.surv_bm25[base + i]) - double(bmin)) / brange;
            double vals[kRerankMaxChunks];
            double best = -1.0;
            int bi = 0;
            for (int j = 0; j < n; ++j) {
                const double cn = crange == 0.0 ? 0.0 : (double(a.cos[(base + i) * kRerankMaxChunks + j]) - double(cmin)) / crange;
                vals[j] = cn * (1.0 - sm) + oldn * sm;
                if (vals[j] > best) { best = vals[j]; bi = j; }
            }
            if (n > 1) {
                const double adj = 0.1 - (0.1 + 0.05) * (double(bi) / double(n - 1));
                double v = vals[bi] + adj;
                vals[bi] = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
                best = -1.0;
                for (int j = 0; j < n; ++j) if (vals[j] > best) { best = vals[j]; bi = j; }
            }
            const float sc = float(best) + 0.0f;
            key = (uint64_t(float_to_key(sc)) << 32) | (uint64_t(0x3fffffu - uint32_t(i)) << 10) | uint64_t(i);
            s_score[i] = sc; s_orig[i] = float(oldn); s_best[i] = uint8_t(bi);
        }
        s_key[i] = key;
    }
    __syncthreads();
    block_bitonic_desc<NT>(s_key, P);
    int nd = 0;
    for (int i = tid; i < P; i += NT) nd += (s_key[i] != 0);
    nd = int(block_reduce<NT>(float(nd), s_tmp, 0) + 0.5f);
    const int n_out = nd < a.max_out ? nd : a.max_out;
    for (int o = tid; o < a.max_out; o += NT) {
        const int64_t dst = int64_t(qi) * a.max_out + o;
        if (o < n_out) {
            const int i = int(s_key[o] & 0x3ffull);
            a.out_doc[dst] = a.surv_doc[base + i];
            a.out_score[dst] = s_score[i];
            a.out_orig[dst] = s_orig[i];
            a.out_chunk[dst] = a.chunk0[base + i] + int64_t(s_best[i]);
        } else {
            a.out_doc[dst] = -1; a.out_score[dst] = 0.f; a.out_orig[dst] = 0.f; a.out_chunk[dst] = -1;
        }
    }
    if (tid == 0) { a.out_count[qi] = n_out; a.out_rows[qi] = T; }
}

}  // namespace mse
