// K1, two-phase form — the BM25 scoring loop of BM25.search (indexer/bm25_indexer.py:459-481) for batches whose
// min_score is >= 0 (the reference's 0.0).
//
// Why: the fp32 kernel of bm25.cuh is bound by the shared-memory data pipe, and a third of its wavefronts re-arm the
// 6 KB of fp32 accumulators of every (sub-range, query) task (`profiles/README.md`, round 2: without the re-arm the
// same kernel takes 10.1 instead of 18.1 ms; a single `st.bulk` zero fill costs the same as the 12 STS.128 it replaces —
// it passes through the same pipe).  Here phase 1 accumulates a 16-bit UPPER BOUND of every document's score, so the
// same 6 KB cover 3072 documents: half the tasks, half the re-arm bytes and half the per-task control per document.
// Phase 2 forms the exact fp32 score of the few documents whose bound reaches the running k-th best score: a warp
// queues them (query, sub-range, document) and, 32 at a time, every lane looks its document up in the posting slices of
// its query's terms (binary search, the slices are L2-resident) and adds the contributions in the reference's term
// order with the same round-down FMAs as the fp32 kernel — the scores are bit-identical to it.
//
//   phase 1, per posting:  q = ceil(w_t * invU * impact) + 1   (one round-up FMA onto 2^23 + 1; w_t < 0 counts as 0)
//                          acc16[doc] += q                      (LDS.U16 / IADD3 / STS.U16, docs unique inside a term)
//                          hit iff acc16[doc] - class * pen16 >= tau16   (class: 4 bits of the posting, bm25.cuh)
//   invU = 60000 / (sum of the query's positive weights), so an accumulator cannot overflow 16 bits; the "+ 1" keeps a
//   touched document apart from an untouched one (score exactly 0.0 is kept by the reference, bm25_indexer.py:480).
//   With U = 1 / invU:  exact score <= (acc16 - class * pen16) * U  and  tau16 <= tau / U, i.e. a document that
//   reaches the bound is always a hit (no false negative); false positives fall out at the exact test of phase 2.  (The
//   "+ 1" of every touching term also absorbs the fp32 rounding of the exact score, a fraction of a unit; the arithmetic
//   is restated in oracle/bm25_oracle.py::two_phase_upper_bound and checked on CPU by tests/test_oracle_bm25.py.)
//   An accumulator only grows, so a document reaches the bound iff its LAST update does: every lane remembers its last
//   two hit events.  A task with a third event on one lane, or with more than kExactModeEvents in all (the first tasks
//   of a query, while its bound is still low), is rescored in EXACT MODE: the 6 KB become fp32 accumulators of half the
//   sub-range and the task's postings are applied to each half in turn, exactly as bm25_score_kernel does.
#pragma once
#include "bm25.cuh"

namespace mse {

#ifndef MSE_BM25_RANGE16
#define MSE_BM25_RANGE16 3072
#endif
constexpr int kBm25Range16 = MSE_BM25_RANGE16;   // docs per sub-range: 6 KB of 16-bit accumulators per warp
constexpr int kBm25Ctas16 = kBm25Range16 <= 3072 ? 4 : 3;
constexpr int kQueueDocBits = 13;                // doc within the sub-range
constexpr int kQueueMax = 32;                    // hit documents a warp collects before it forms their exact scores
#ifndef MSE_BM25_EXACT_EVENTS
#define MSE_BM25_EXACT_EVENTS 12
#endif
constexpr int kExactModeEvents = MSE_BM25_EXACT_EVENTS;   // more hit events than this in one task: exact mode
static_assert(kQueueMax * 8 == 2 * kEmitStage * 8, "the queue takes the place of the fp32 kernel's emission stage");

__global__ void __launch_bounds__(kBm25Threads, kBm25Ctas16)
bm25_score16_kernel(Bm25Dev ix, Bm25Work w) {
    extern __shared__ __align__(16) unsigned char bm25_smem[];
    constexpr int RS = kBm25Range16;
    int lane, wid;                                   // (volatile: the compiler otherwise re-reads the special registers all over the kernel)
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("shr.u32 %0, %1, 5;" : "=r"(wid) : "r"(threadIdx.x));
    unsigned char* my = bm25_smem + bm25_score_warp_bytes(RS / 2) * wid;
    uint4* s_meta = reinterpret_cast<uint4*>(my);                                  // {begin, count, weight bits, 1 + dense row}
    uint64_t* s_queue = reinterpret_cast<uint64_t*>(my + kMetaSlots * 16);         // [kQueueMax] query << 32 | sub-range << 13 | doc in it
    uint4* s_qinfo = reinterpret_cast<uint4*>(my + kMetaSlots * 16 + size_t(kQueueMax) * 8 + 16);
    uint16_t* s_acc = reinterpret_cast<uint16_t*>(my + kMetaSlots * 16 + size_t(kQueueMax) * 8 + 16 + size_t(kBm25MaxQueriesPerItem) * 16);

    const int QC = w.queries_per_item;
    const int chunks = (w.n_queries + QC - 1) / QC;
    const int n_items = w.n_sub * chunks;
    const uint32_t ulane = uint32_t(lane);
    const unsigned lt_mask = (1u << lane) - 1u;
    const int2* __restrict__ g_post = ix.post2;
    constexpr int MP = kPrefetchSlots;
    constexpr int kSweep = RS * 2 / 16 / 32;         // 16-byte stores per lane that re-arm a sub-range (12)
    constexpr uint32_t kNoDoc = 0x1fffffffu;         // matches no posting (doc fields are 28 bits)

    uint4* const a4 = reinterpret_cast<uint4*>(s_acc) + lane;
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int it = 0; it < kSweep; ++it) a4[it * 32] = z4;
    __syncwarp();
    int lo = 0;
    int qcount = 0;                                  // queued hit documents (warp-uniform)

    // ---- emission of exact candidates ----------------------------------------------------------------------------
    // warp-uniform calls; `bits` are the bits of the NEGATED exact score, `ldoc` the shard-local document
    auto refresh_bounds = [&](unsigned need, int q) {                  // lanes in `need`: their query's count crossed a multiple of 64
        while (need) {
            const int l = __ffs(int(need)) - 1;
            need &= need - 1u;
            tau_raise(w.ts, __shfl_sync(0xffffffffu, q, l));
        }
    };
    auto emit_mixed = [&](bool on, int q, uint32_t bits, uint32_t ldoc) {     // a different query on every lane
        bool cross = false;
        if (on) {
            const uint32_t key = float_to_key(0.0f - __uint_as_float(bits));
            const int g = atomicAdd(&w.cand_count[q], 1);
            if (g < w.cap) w.cand[int64_t(q) * w.cap + g] = make_key64(key, ix.doc_base + ldoc);
            else w.overflow[q] = 1;
            if (w.use_tau) { atomicAdd(&w.ts.hist[int64_t(q) * kHistBins + (key >> kHistShift)], 1u); cross = (((g + 1) ^ g) >> 6) != 0; }
        }
        refresh_bounds(__ballot_sync(0xffffffffu, cross), q);
    };
    auto emit_same = [&](bool on, int q, uint32_t bits, uint32_t ldoc) {      // one query: a single slot reservation
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (m == 0u) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(&w.cand_count[q], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (on) {
            const uint32_t key = float_to_key(0.0f - __uint_as_float(bits));
            const int g = base + __popc(m & lt_mask);
            if (g < w.cap) w.cand[int64_t(q) * w.cap + g] = make_key64(key, ix.doc_base + ldoc);
            else w.overflow[q] = 1;
            if (w.use_tau) atomicAdd(&w.ts.hist[int64_t(q) * kHistBins + (key >> kHistShift)], 1u);
        }
        if (w.use_tau && (((base + __popc(m)) ^ base) >> 6)) tau_raise(w.ts, q);
    };

    // ---- phase 2: exact scores of the queued documents, one per lane -----------------------------------------------
    auto flush = [&]() {
        if (qcount == 0) return;
        __syncwarp();
        const bool on = lane < qcount;
        const uint64_t e = on ? s_queue[lane] : 0ull;
        const int q = int(e >> 32);
        const uint32_t sub = uint32_t(e) >> kQueueDocBits;
        const uint32_t ldoc = sub * uint32_t(RS) + (uint32_t(e) & ((1u << kQueueDocBits) - 1u));
        int s0 = 0, ns = 0;
        uint4 qi = make_uint4(0u, 0u, 0u, 0u);
        if (on) { s0 = w.q_off[q]; ns = w.q_off[q + 1] - s0; qi = w.qinfo[q]; }
        const int maxs = __reduce_max_sync(0xffffffffu, ns);
        const uint2* __restrict__ rec = w.rec + int64_t(sub) * w.n_slots + s0;
        float ex = 0.f;
#pragma unroll 1
        for (int k = 0; k < maxs; k += 2) {                            // the query's terms in order, two searches in flight
            uint32_t bA = 0u, bB = 0u;
            int hA = 0, hB = 0;
            float wA = 0.f, wB = 0.f;
            if (on && k < ns) { const uint2 r = rec[k]; bA = r.x; hA = int(r.y & 0xffffu); wA = w.slot_w[s0 + k]; }
            if (on && k + 1 < ns) { const uint2 r = rec[k + 1]; bB = r.x; hB = int(r.y & 0xffffu); wB = w.slot_w[s0 + k + 1]; }
            int lA = 0, lB = 0, iA = 0, iB = 0;
            uint32_t cA = kNoDoc, cB = kNoDoc;
            while (__any_sync(0xffffffffu, (lA < hA) | (lB < hB))) {   // lower bound of the document in each slice
                const bool a = lA < hA, b = lB < hB;
                const int mA = (lA + hA) >> 1, mB = (lB + hB) >> 1;
                int2 pA = make_int2(0, 0), pB = make_int2(0, 0);
                if (a) pA = __ldg(g_post + (bA + uint32_t(mA)));
                if (b) pB = __ldg(g_post + (bB + uint32_t(mB)));
                if (a) { const uint32_t dm = uint32_t(pA.x) & kDocMask; if (dm < ldoc) lA = mA + 1; else { hA = mA; cA = dm; iA = pA.y; } }
                if (b) { const uint32_t dm = uint32_t(pB.x) & kDocMask; if (dm < ldoc) lB = mB + 1; else { hB = mB; cB = dm; iB = pB.y; } }
            }
            if (cA == ldoc) ex = __fmaf_rd(-wA, __int_as_float(iA), ex);
            if (cB == ldoc) ex = __fmaf_rd(-wB, __int_as_float(iB), ex);
        }
        // looked-up negative terms: after the streamed ones, in slot order (as bm25_score_kernel)
        const int look = int(qi.x & 0xffu);
        if (look == 1) ex = __fmaf_rd(-__uint_as_float(qi.y), __ldg(ix.neg_imp + int64_t(qi.z - 1u) * ix.neg_stride + ldoc), ex);
        if (__any_sync(0xffffffffu, look > 1)) {
#pragma unroll 1
            for (int k = 0; k < maxs; ++k) {
                if (look > 1 && k < ns) {
                    const uint32_t row = (rec[k].y >> kRecRowShift) & 0x7fffu;
                    if (row) ex = __fmaf_rd(-w.slot_w[s0 + k], __ldg(ix.neg_imp + int64_t(row - 1u) * ix.neg_stride + ldoc), ex);
                }
            }
        }
        // the exact test, against the query's bound as it stands now (any value it ever took is a valid bound)
        uint32_t tau_key = w.min_key;
        if (on && w.use_tau) tau_key = ld_relaxed_u32(&w.ts.tau[q]);
        const uint32_t tau_u = __float_as_uint(key_to_float(tau_key)) | 0x80000000u;
        emit_mixed(on && __float_as_uint(ex) >= tau_u, q, __float_as_uint(ex), ldoc);
        qcount = 0;
        __syncwarp();
    };

    int neg_lo = 0, neg_hi = 0, look_n = 0;
    uint32_t look_w = 0u, look_row = 0u;
    auto neg_adjust = [&](uint32_t bits, int d) -> uint32_t {          // looked-up negative terms of the CURRENT query (exact mode)
        float a = __uint_as_float(bits);
        if (look_n == 1)
            return __float_as_uint(__fmaf_rd(-__uint_as_float(look_w), __ldg(ix.neg_imp + int64_t(look_row - 1u) * ix.neg_stride + (lo + d)), a));
#pragma unroll 1
        for (int sl = neg_lo; sl < neg_hi; ++sl) {
            const uint4 m = s_meta[sl];
            if (m.w) a = __fmaf_rd(-__uint_as_float(m.z), __ldg(ix.neg_imp + int64_t(m.w - 1u) * ix.neg_stride + (lo + d)), a);
        }
        return __float_as_uint(a);
    };

    // ---- phase 1: one warp-round (<= 32 postings of one term) -------------------------------------------------
    int hc = 0;                                       // this lane's hit events in the current task
    uint32_t h1 = 0u, h2 = 0u;                        // shared-memory addresses of the accumulators of the last two
    int tau16 = 0, npen = 0;                          // floor(tau * invU); minus floor(class penalty unit * invU)
    const uint32_t acc_addr = uint32_t(__cvta_generic_to_shared(s_acc));
    uint32_t acc_lo = acc_addr;                       // acc_addr - 2 * lo
#define MSE_U16_HEAD(D, V, A)      "setp.ne.b32 " V ", " D ", -1;\n\t" "and.b32 " A ", " D ", 0x0fffffff;\n\t" "mad.lo.u32 " A ", " A ", 2, %5;\n\t"
#define MSE_U16_LOAD(V, A, O)      "mov.b32 " O ", 0;\n\t" "@" V " ld.shared.u16 " O ", [" A "];\n\t"
#define MSE_U16_ADD(T, O, N)       "fma.rp.f32 y, " T ", %6, 0f4B000001;\n\t" "mov.b32 " N ", y;\n\t" "add.s32 " N ", " N ", " O ";\n\t" "sub.u32 " N ", " N ", 0x4B000000;\n\t"
#define MSE_U16_STORE(V, A, N)     "@" V " st.shared.u16 [" A "], " N ";\n\t"
#define MSE_U16_TEST(D, V, A, N)   "shr.u32 c, " D ", 28;\n\t" "mad.lo.s32 c, c, %7, " N ";\n\t" "setp.ge.and.s32 h, c, %8, " V ";\n\t" \
                                   "@h add.s32 %0, %0, 1;\n\t" "@h mov.b32 %2, %1;\n\t" "@h mov.b32 %1, " A ";\n\t"
    auto apply = [&](int dd, int tfi, float ws) {      // dd == -1: no posting on this lane
        asm volatile("{\n\t.reg .pred v, h;\n\t.reg .b32 a, o, n, c;\n\t.reg .f32 y;\n\t"
                     MSE_U16_HEAD("%3", "v", "a") MSE_U16_LOAD("v", "a", "o") MSE_U16_ADD("%4", "o", "n") MSE_U16_STORE("v", "a", "n")
                     MSE_U16_TEST("%3", "v", "a", "n") "}"
                     : "+r"(hc), "+r"(h1), "+r"(h2)
                     : "r"(dd), "f"(__int_as_float(tfi)), "r"(acc_lo), "f"(ws), "r"(npen), "r"(tau16));
    };
#undef MSE_U16_HEAD
#undef MSE_U16_TEST
#undef MSE_U16_ADD
    // four rounds of ONE term (distinct documents): loads, adds and stores of the four are interleaved, so a round does
    // not wait for the shared-memory round trip of the one before it; the hit tests follow in round order
#define MSE_U16_HEAD(D, V, A)      "setp.ne.b32 " V ", " D ", -1;\n\t" "and.b32 " A ", " D ", 0x0fffffff;\n\t" "mad.lo.u32 " A ", " A ", 2, %11;\n\t"
#define MSE_U16_ADD(T, O, N)       "fma.rp.f32 y, " T ", %12, 0f4B000001;\n\t" "mov.b32 " N ", y;\n\t" "add.s32 " N ", " N ", " O ";\n\t" "sub.u32 " N ", " N ", 0x4B000000;\n\t"
#define MSE_U16_TEST(D, V, A, N)   "shr.u32 c, " D ", 28;\n\t" "mad.lo.s32 c, c, %13, " N ";\n\t" "setp.ge.and.s32 h, c, %14, " V ";\n\t" \
                                   "@h add.s32 %0, %0, 1;\n\t" "@h mov.b32 %2, %1;\n\t" "@h mov.b32 %1, " A ";\n\t"
    auto apply4 = [&](const int (&dd)[4], const int (&tt)[4], float ws) {
        asm volatile("{\n\t.reg .pred v0, v1, v2, v3, h;\n\t.reg .b32 a0, a1, a2, a3, o0, o1, o2, o3, n0, n1, n2, n3, c;\n\t.reg .f32 y;\n\t"
                     MSE_U16_HEAD("%3", "v0", "a0") MSE_U16_HEAD("%4", "v1", "a1") MSE_U16_HEAD("%5", "v2", "a2") MSE_U16_HEAD("%6", "v3", "a3")
                     MSE_U16_LOAD("v0", "a0", "o0") MSE_U16_LOAD("v1", "a1", "o1") MSE_U16_LOAD("v2", "a2", "o2") MSE_U16_LOAD("v3", "a3", "o3")
                     MSE_U16_ADD("%7", "o0", "n0") MSE_U16_ADD("%8", "o1", "n1") MSE_U16_ADD("%9", "o2", "n2") MSE_U16_ADD("%10", "o3", "n3")
                     MSE_U16_STORE("v0", "a0", "n0") MSE_U16_STORE("v1", "a1", "n1") MSE_U16_STORE("v2", "a2", "n2") MSE_U16_STORE("v3", "a3", "n3")
                     MSE_U16_TEST("%3", "v0", "a0", "n0") MSE_U16_TEST("%4", "v1", "a1", "n1") MSE_U16_TEST("%5", "v2", "a2", "n2")
                     MSE_U16_TEST("%6", "v3", "a3", "n3") "}"
                     : "+r"(hc), "+r"(h1), "+r"(h2)
                     : "r"(dd[0]), "r"(dd[1]), "r"(dd[2]), "r"(dd[3]), "f"(__int_as_float(tt[0])), "f"(__int_as_float(tt[1])),
                       "f"(__int_as_float(tt[2])), "f"(__int_as_float(tt[3])), "r"(acc_lo), "f"(ws), "r"(npen), "r"(tau16));
    };
#undef MSE_U16_HEAD
#undef MSE_U16_LOAD
#undef MSE_U16_ADD
#undef MSE_U16_STORE
#undef MSE_U16_TEST
    // postings 32.. of a slice: four rounds at a time (their loads are issued before the first is applied); a last
    // group of at most 32 postings runs one round
    auto load4 = [&](const int2* __restrict__ p, int rem, int (&dd)[4], int (&tt)[4]) {     // p: this lane's posting of the first round
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dd[u] = -1; tt[u] = 0;
            if (lane + 32 * u < rem) { const int2 v = ldg_stream_i2(p + 32 * u); dd[u] = v.x; tt[u] = v.y; }
        }
    };
    auto apply_rest = [&](uint32_t begin, int n, float ws) {
        const int2* __restrict__ p = g_post + (begin + 32u + ulane);    // (posting indices fit 32 bits: one widening add)
#pragma unroll 1
        for (int rem = n - 32; rem > 0; rem -= 128, p += 128) {
            if (rem <= 32) {
                int d1 = -1, t1 = 0;
                if (lane < rem) { const int2 v = ldg_stream_i2(p); d1 = v.x; t1 = v.y; }
                apply(d1, t1, ws);
                break;
            }
            int dd[4], tt[4];
            load4(p, rem, dd, tt);
            apply4(dd, tt, ws);
        }
    };

    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(w.item_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int j = item / chunks, c = item - j * chunks;
        lo = j * RS;
        acc_lo = acc_addr - 2u * uint32_t(lo);
        const int q0 = c * QC;
        const int q1 = (q0 + QC) < w.n_queries ? (q0 + QC) : w.n_queries;
        const int nq = q1 - q0;
        const int qo_reg = (lane <= nq) ? w.q_off[q0 + lane] : 0;
        const uint32_t tau_reg = (lane < nq && w.use_tau) ? ld_relaxed_u32(&w.ts.tau[q0 + lane]) : w.min_key;
        const float inv_reg = (lane < nq) ? w.inv_unit[q0 + lane] : 0.f;
        __syncwarp();
        if (lane < nq) s_qinfo[lane] = w.qinfo[q0 + lane];
        __syncwarp();
        const uint2* __restrict__ rec = w.rec + int64_t(j) * w.n_slots;

        int qa = 0;
        while (qa < nq) {
            // largest qb in (qa, nq] whose slots still fit the staging area (a query has <= 32 slots)
            const int sa = __shfl_sync(0xffffffffu, qo_reg, qa);
            const unsigned okm = __ballot_sync(0xffffffffu, lane > qa && lane <= nq && (qo_reg - sa) <= kMetaSlots);
            const int qb = 31 - __clz(okm);
            const int ns = __shfl_sync(0xffffffffu, qo_reg, qb) - sa;
            __syncwarp();
            for (int i = lane; i < ns; i += 32) {
                const uint2 r = rec[sa + i];
                s_meta[i] = make_uint4(r.x, r.y & 0xffffu, __float_as_uint(w.slot_w[sa + i]), (r.y >> kRecRowShift) & 0x7fffu);
            }
            __syncwarp();

            // first round (<= 32 postings) of the first MP terms of a query: loaded one query ahead, into the registers the
            // current query has just consumed; a longer slice is announced to L2 at the same time (one 128-byte line per lane)
            int pd[MP], pt[MP];
            int o_nxt = 0;                                                  // first staged slot of the next query
            int e_nxt = __shfl_sync(0xffffffffu, qo_reg, qa + 1) - sa;
            auto fetch_first = [&](int t) {                                // term t of the query whose staged slots are [o_nxt, e_nxt)
                pd[t] = -1; pt[t] = 0;
                if (o_nxt + t < e_nxt) {
                    const uint4 m = s_meta[o_nxt + t];
                    if (lane < int(m.y)) { const int2 p = ldg_stream_i2(g_post + (m.x + ulane)); pd[t] = p.x; pt[t] = p.y; }
#ifndef MSE_BM25_NO_L2_PREFETCH
                    if (int(m.y) > 32 + 16 * lane) asm volatile("prefetch.global.L2 [%0];" :: "l"(g_post + (m.x + 32u + 16u * ulane)));
#endif
                }
            };
#pragma unroll
            for (int t = 0; t < MP; ++t) fetch_first(t);                    // the first query of the group
#pragma unroll 1
            for (int qr = qa; qr < qb; ++qr) {
                const int o_cur = o_nxt, e_cur = e_nxt;
                const int q = q0 + qr;
                const uint32_t tau_key = __shfl_sync(0xffffffffu, tau_reg, qr);
                const float invU = __shfl_sync(0xffffffffu, inv_reg, qr);
                o_nxt = e_cur;
                e_nxt = (qr + 1 < qb) ? __shfl_sync(0xffffffffu, qo_reg, qr + 2) - sa : e_cur;      // last query of the group: nothing to fetch
                // ---- phase 1 -----------------------------------------------------------------------------------
                const float tau_f = key_to_float(tau_key);                      // >= +0.0: the host sends min_score >= 0 here
                const uint32_t tau_u = __float_as_uint(tau_f) | 0x80000000u;    // bits of -tau
                const uint4 qi = s_qinfo[qr];
                look_n = int(qi.x & 0xffu); look_w = qi.y; look_row = qi.z;
                tau16 = __float2int_rd(__fmul_rd(tau_f, invU));
                npen = -__float2int_rd(__fmul_rd(__uint_as_float(qi.w), invU));
                hc = 0;
                int touched = 0;
                // no bound yet (the first tasks of a query): every touched document would be a hit, phase 1 is skipped
                const bool direct = tau_key == float_to_key(0.0f);
#pragma unroll
                for (int t = 0; t < MP; ++t) {
                    if (o_cur + t < e_cur) {
                        const uint4 m = s_meta[o_cur + t];
                        const int n = int(m.y);
                        if (n > 0 && direct) touched = 1;
                        else if (n > 0) {
                            const float ws = __fmul_ru(fmaxf(__uint_as_float(m.z), 0.f), invU);
                            touched = 1;
                            apply(pd[t], pt[t], ws);
                            if (n > 32) apply_rest(m.x, n, ws);
                            __syncwarp();                                   // next term may touch the same docs
                        }
                    }
                    fetch_first(t);                                         // the same term of the next query
                }
#pragma unroll 1
                for (int sl = o_cur + MP; sl < e_cur; ++sl) {               // queries with more than MP terms
                    const uint4 m = s_meta[sl];
                    const int n = int(m.y);
                    if (n == 0) continue;
                    touched = 1;
                    if (direct) continue;
                    const float ws = __fmul_ru(fmaxf(__uint_as_float(m.z), 0.f), invU);
                    int dd = -1, tfi = 0;
                    if (lane < n) { const int2 p = ldg_stream_i2(g_post + (m.x + ulane)); dd = p.x; tfi = p.y; }
                    apply(dd, tfi, ws);
                    if (n > 32) apply_rest(m.x, n, ws);
                    __syncwarp();
                }
                if (!touched) continue;
                // ---- hit documents -----------------------------------------------------------------------------
                const unsigned ev1 = __ballot_sync(0xffffffffu, hc > 0);
                if (ev1 | unsigned(direct)) {
                    const int n_events = __popc(ev1) + __popc(__ballot_sync(0xffffffffu, hc > 1));
                    if (!direct && n_events <= kExactModeEvents && !__any_sync(0xffffffffu, hc > 2)) {
                        // the remembered events name every document that can reach the bound: queue them.  A document's
                        // accumulator is cleared when it is taken, so that a second event of the same document finds it done
                        if (qcount + n_events > kQueueMax) flush();
#pragma unroll 1
                        for (int pass = 0; pass < 2; ++pass) {
                            const bool have = hc > pass;
                            if (!__any_sync(0xffffffffu, have)) break;
                            int d = -1 - lane;                               // idle lanes: distinct keys
                            bool act = false;
                            if (have) { d = int((pass == 0 ? h1 : h2) - acc_addr) >> 1; act = s_acc[d] != 0; }
                            const unsigned same = __match_any_sync(0xffffffffu, d);
                            act = act && (same & lt_mask) == 0u;
                            __syncwarp();
                            if (act) s_acc[d] = 0;
                            const unsigned tm = __ballot_sync(0xffffffffu, act);
                            if (act) s_queue[qcount + __popc(tm & lt_mask)] = (uint64_t(uint32_t(q)) << 32) | uint64_t((uint32_t(j) << kQueueDocBits) | uint32_t(d));
                            qcount += __popc(tm);
                            __syncwarp();
                        }
                    } else {
                        // exact mode: fp32 accumulators of HALF the sub-range, the task's postings applied to each half in
                        // turn (NEGATED score, round-down FMAs, -0.0 = touched with score zero); every accumulator that
                        // reaches the bound is adjusted for the looked-up terms and emitted
                        float* acc32 = reinterpret_cast<float*>(s_acc);
                        if (lane == 0) atomicAdd(w.stats + 4, 1ull);
                        neg_lo = o_cur; neg_hi = look_n ? e_cur : 0;
                        int split = 0;                                       // lane k: where slot o_cur + k's postings of the second half begin
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {
                            __syncwarp();
#pragma unroll
                            for (int it = 0; it < kSweep; ++it) a4[it * 32] = z4;
                            __syncwarp();
                            const uint32_t base = uint32_t(lo + half * (RS / 2));
                            const uint32_t mid_doc = uint32_t(lo + RS / 2);
                            auto apply32 = [&](int dd, int tfi, float wt) {
                                const uint32_t l = (uint32_t(dd) & kDocMask) - base;
                                if (dd != -1 && l < uint32_t(RS / 2)) acc32[l] = __fmaf_rd(-wt, __int_as_float(tfi), acc32[l]);
                            };
#pragma unroll 1
                            for (int sl = o_cur; sl < e_cur; ++sl) {
                                const uint4 m = s_meta[sl];
                                const int n = int(m.y);
                                if (n == 0) continue;
                                const float wt = __uint_as_float(m.z);
                                // the first half stops at the group of 128 postings that reaches the middle of the sub-range
                                // (a slice is sorted by document), the second half starts there
                                const int start = half ? __shfl_sync(0xffffffffu, split, sl - o_cur) : 0;
                                const int2* __restrict__ p = g_post + (m.x + uint32_t(start) + ulane);
                                int done = n;
#pragma unroll 1
                                for (int rem = n - start; rem > 0; rem -= 128, p += 128) {
                                    int dd[4], tt[4];
                                    load4(p, rem, dd, tt);
                                    bool ge = false;
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        apply32(dd[u], tt[u], wt);
                                        ge = ge || (dd[u] != -1 && (uint32_t(dd[u]) & kDocMask) >= mid_doc);
                                    }
                                    if (half == 0 && __any_sync(0xffffffffu, ge)) { done = n - rem; break; }
                                }
                                if (half == 0 && lane == sl - o_cur) split = done;
                                __syncwarp();
                            }
                            // read-out: lane's accumulators 4 * (it * 32 + lane) + u; bit (it * 4 + u) of `fl` = reaches the bound
                            // (one unsigned compare: tau >= +0.0, so untouched (+0.0) and negative scores fail it)
                            uint64_t fl = 0ull;
#pragma unroll
                            for (int it = 0; it < kSweep; ++it) {
                                const uint4 v = a4[it * 32];
                                const uint32_t f4 = uint32_t(v.x >= tau_u) | (uint32_t(v.y >= tau_u) << 1) | (uint32_t(v.z >= tau_u) << 2) |
                                                    (uint32_t(v.w >= tau_u) << 3);
                                fl |= uint64_t(f4) << (it * 4);
                            }
                            while (__any_sync(0xffffffffu, fl != 0ull)) {      // warp-uniform; one candidate per lane and round
                                const bool act = fl != 0ull;
                                int d = 0;
                                uint32_t bits = 0u;
                                if (act) {
                                    const int bit = __ffsll((long long)fl) - 1;
                                    fl &= fl - 1ull;
                                    const int l = ((bit >> 2) * 32 + lane) * 4 + (bit & 3);
                                    d = half * (RS / 2) + l;
                                    bits = reinterpret_cast<const uint32_t*>(acc32)[l];
                                    if (neg_hi) bits = neg_adjust(bits, d);
                                }
                                emit_same(act && bits >= tau_u, q, bits, uint32_t(lo + d));
                            }
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < kSweep; ++it) a4[it * 32] = z4;      // re-arm
                __syncwarp();
            }
            qa = qb;
        }
    }
    flush();
}

}  // namespace mse
