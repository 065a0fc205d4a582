// K1 (second generation) — BM25 scoring with bulk-copy staged posting slices.
//
// Same contract as bm25_score_kernel (bm25.cuh): replaces the candidate SQL, the dict grouping
// and the float64 scoring loop of BM25.search (indexer/bm25_indexer.py:435-481).  Same work
// decomposition too — a (sub-range of RS docs, query) pair is a task owned by one warp, with
// fp32 accumulators in shared memory resting at -0.0f — but the task's posting slices are no
// longer fetched by the lanes:
//
//   * stage: the lanes that hold the task's {first posting, count} records (one per query term,
//     written by bm25_prepare_kernel) each issue one cp.async.bulk copy of their 16-byte-aligned
//     slice of the interleaved {doc, tf} posting array into a per-warp ring of staging buffers;
//     one mbarrier per buffer collects the bytes (expect_tx).  Tasks are staged NBUF-1 ahead of the task being scored and
//     their records are loaded one task earlier still, so the copy engine — not warp
//     occupancy — covers HBM latency, and the scoring loop reads postings with LDS;
//   * score: terms in the reference's order, 32 postings per warp round, plain read-modify-write
//     (doc ids are unique inside a term; __syncwarp between terms) — deterministic summation.
//     A slice that does not fit the staging capacity is staged as a prefix and its tail is read
//     with ordinary loads;
//   * read-out: either a scan of the RS accumulators, 128 per warp round, or (fully staged tasks,
//     read-out mode 1) a second walk over the staged doc ids, term by term: a document touched by
//     two terms is taken by the first and found re-armed by the second, so the cost follows the
//     postings, not the range.
//     Both emit only scores >= tau[q] (running lower bound of the k-th best, topk.cuh) and re-arm
//     the accumulators to -0.0f.  Candidates are staged in shared memory, their list slot is
//     reserved with one atomicAdd per task whose result is consumed one task later.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "topk.cuh"
#include "bm25.cuh"

namespace mse {

constexpr int kStSlotsMax = 32;                  // distinct terms per query (checked by the caller)
constexpr int kStEmit = 32;                      // candidates of one task staged in shared memory
constexpr int kStHdrBytes = 32;                  // {j, q, T, padded postings} + mbarrier
constexpr int kStTailSlack = 1024;               // bytes after the last warp's area: unrolled rounds may read past a slice

__host__ __device__ inline size_t bm25_staged_buf_bytes(int cap, int slots) {
    return size_t(kStHdrBytes) + size_t(slots) * 16 + size_t(cap) * 8;
}
__host__ __device__ inline size_t bm25_staged_warp_bytes(int rs, int cap, int slots, int nbuf, bool len16) {
    // accumulators and lengths carry one extra 16-byte unit: slot RS is the dummy target of idle lanes
    return (size_t(rs) * 4 + 16) + (size_t(rs) * (len16 ? 2 : 4) + 16) + size_t(nbuf) * bm25_staged_buf_bytes(cap, slots) +
           size_t(2 * kStEmit) * 8;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1 << 24)) __trap();       // a copy that never lands must not hang the device
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}


// Running bound of the staged kernel: the k-th best score key among the first n entries of a query's candidate
// list (every entry is a document whose FINAL score is known, so the k-th best of any subset is a lower
// bound of the final k-th best; unwritten entries are zero and only make the bound more conservative).
// MSB-first radix select with 6-bit digits on the score key; one warp, 64-bin histogram in shared memory.
// The result keeps its undecided low bits at zero (still a valid bound).
__device__ __noinline__ void staged_tau_refresh(uint32_t* tau_q, const uint64_t* __restrict__ list, int n, int k, int* hist) {
    const int lane = lane_id();
    constexpr int U = 8;                                                  // independent L2 loads in flight per lane
    uint32_t vor = 0u, vand = 0xffffffffu;
    int valid = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += 32 * U) {
        uint32_t key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + 32 * u + lane;
            key[u] = i < n ? uint32_t(__ldcg(list + i) >> 32) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) if (key[u]) { vor |= key[u]; vand &= key[u]; ++valid; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vor |= __shfl_xor_sync(0xffffffffu, vor, o);
        vand &= __shfl_xor_sync(0xffffffffu, vand, o);
        valid += __shfl_xor_sync(0xffffffffu, valid, o);
    }
    const uint32_t diff = vor ^ vand;
    if (valid < k) return;                                                // fewer than k entries written: no bound yet
    int top = diff ? 32 - __clz(int(diff)) : 0;                           // bits >= top are common to every key
    uint32_t prefix = top >= 32 ? 0u : (vand >> top) << top;
    uint32_t mask = top >= 32 ? 0u : (0xffffffffu >> top) << top;
    int krem = k;
    int passes = 0;
#pragma unroll 1
    while (top > 6 && passes < 2) {
        const int shift = top - 6;
        hist[lane] = 0; hist[lane + 32] = 0;
        __syncwarp();
#pragma unroll 1
        for (int i0 = 0; i0 < n; i0 += 32 * U) {
            uint32_t key[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + 32 * u + lane;
                key[u] = i < n ? uint32_t(__ldcg(list + i) >> 32) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (key[u] != 0u && (key[u] & mask) == prefix) atomicAdd(&hist[(key[u] >> shift) & 63u], 1);
        }
        __syncwarp();
        const int c_hi = hist[63 - 2 * lane], c_lo = hist[62 - 2 * lane];   // descending digits across lanes
        const int s2 = c_hi + c_lo;
        const int incl = warp_incl_scan(s2);
        const int excl = incl - s2;
        const bool hit_hi = excl < krem && krem <= excl + c_hi;
        const bool hit_lo = !hit_hi && excl + c_hi < krem && krem <= incl;
        const unsigned bal = __ballot_sync(0xffffffffu, hit_hi || hit_lo);
        if (bal == 0u) return;
        const int src = __ffs(int(bal)) - 1;
        const int digit = __shfl_sync(0xffffffffu, hit_hi ? 63 - 2 * lane : 62 - 2 * lane, src);
        const int above = __shfl_sync(0xffffffffu, hit_hi ? excl : excl + c_hi, src);
        const int inbin = __shfl_sync(0xffffffffu, hit_hi ? c_hi : c_lo, src);
        prefix |= uint32_t(digit) << shift;
        mask |= 63u << shift;
        krem -= above;
        top = shift;
        ++passes;
        __syncwarp();
        if (inbin == krem) break;                                         // the bin is taken whole: its lower edge is exact enough
    }
    if (lane == 0) atomicMax(tau_q, prefix);                              // undecided low bits stay zero: still a bound
}

template <bool LEN16, int NBUF, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
bm25_score_staged_kernel(Bm25Dev ix, Bm25Work w) {
    using LenT = typename std::conditional<LEN16, uint16_t, float>::type;
    extern __shared__ __align__(128) unsigned char st_smem[];
    const int RS = w.sub_docs;                   // multiple of 128, <= 4096
    const int CAP = w.stage_cap;                 // postings per staging buffer, multiple of 32, <= 16384
    const int lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const int SLOTS = w.stage_slots;             // slice-table entries per buffer (>= distinct terms of any query)
    const size_t buf_bytes = bm25_staged_buf_bytes(CAP, SLOTS);
    unsigned char* my = st_smem + bm25_staged_warp_bytes(RS, CAP, SLOTS, NBUF, LEN16) * warp_id();
    float* s_acc = reinterpret_cast<float*>(my);
    LenT* s_len = reinterpret_cast<LenT*>(my + size_t(RS) * 4 + 16);
    unsigned char* s_bufs = my + size_t(RS) * (4 + sizeof(LenT)) + 32;
    uint64_t* s_emit = reinterpret_cast<uint64_t*>(s_bufs + size_t(NBUF) * buf_bytes);

    const float neg0 = __uint_as_float(kUntouchedBits);
    const float c0 = ix.norm_c0, c1 = ix.norm_c1;
    const int2* __restrict__ g_post = ix.post2;
    const int S = w.n_slots;

    for (int i = lane; i < RS; i += 32) s_acc[i] = neg0;
    if (lane == 0) {
        s_acc[RS] = 0.f;
        s_len[RS] = LenT(1);
        for (int b = 0; b < NBUF; ++b) mbar_init(smem_u32(s_bufs + b * buf_bytes + 16), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    // ---- cursor over the (sub-range, query) tasks this warp will own --------------------------------
    const int QC = w.queries_per_item;           // <= 31: lane l holds q_off[q0 + l]
    const int chunks = (w.n_queries + QC - 1) / QC;
    const int n_items = w.n_sub * chunks;
    int c_j = 0, c_q0 = 0, c_nq = 0, c_qi = 0, c_qo = 0;
    bool c_ok = true;
    int nxt_raw = 0;                             // lane 0: look-ahead item index (consumed at the next item change)
    auto enter_item = [&](int item) {
        c_ok = item < n_items;
        if (!c_ok) return;
        c_j = item / chunks;
        const int c = item - c_j * chunks;
        c_q0 = c * QC;
        c_nq = (c_q0 + QC) < w.n_queries ? QC : (w.n_queries - c_q0);
        c_qi = 0;
        c_qo = (lane <= c_nq) ? w.q_off[c_q0 + lane] : 0;
        if (lane == 0) nxt_raw = atomicAdd(w.item_counter, 1);
    };
    {
        int it = 0;
        if (lane == 0) it = atomicAdd(w.item_counter, 1);
        enter_item(__shfl_sync(0xffffffffu, it, 0));
    }

    // ---- R: records of the task under the cursor -> registers (lane s holds slot s) -----------------
    uint2 r_rec = make_uint2(0u, 0u);
    float r_w = 0.f;
    int r_j = 0, r_q = 0, r_T = -1;
    auto load_records = [&]() {
        if (!c_ok) { r_T = -1; return; }
        const int sa = __shfl_sync(0xffffffffu, c_qo, c_qi);
        const int se = __shfl_sync(0xffffffffu, c_qo, c_qi + 1);
        r_T = se - sa; r_j = c_j; r_q = c_q0 + c_qi;
        r_rec = make_uint2(0u, 0u); r_w = 0.f;
        if (lane < r_T) {
            r_rec = __ldg(w.rec + int64_t(c_j) * S + sa + lane);
            r_w = __ldg(w.slot_w + sa + lane);
        }
        if (++c_qi >= c_nq) enter_item(__shfl_sync(0xffffffffu, nxt_raw, 0));
    };

    // ---- S: issue the bulk copies of the task held in the registers into buffer b ---------------------
    // Header {j, q, number of non-empty slices, padded postings}; slice table entries {first posting,
    // count, staged offset | staged count << 16, weight bits} for the NON-EMPTY slices only, in term order.
    auto stage = [&](int b) {
        unsigned char* B = s_bufs + size_t(b) * buf_bytes;
        if (r_T < 0) {                                             // end of the stream
            if (lane == 0) *reinterpret_cast<int4*>(B) = make_int4(0, 0, -1, 0);
            return;
        }
        const uint32_t first = r_rec.x, cnt = r_rec.y;
        const unsigned live = __ballot_sync(0xffffffffu, cnt != 0u);
        if (live == 0u) {                                          // no posting of this query in this sub-range
            if (lane == 0) *reinterpret_cast<int4*>(B) = make_int4(r_j, r_q, 0, 0);
            return;
        }
        const uint32_t skew = first & 1u;                          // 16-byte units hold two postings
        const int L = cnt ? int((skew + cnt + 1u) & ~1u) : 0;
        int incl = L;
        for (int o = 1; o < r_T; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - L;
        int Lc = CAP - excl;
        Lc = Lc < L ? Lc : L;
        Lc = Lc < 0 ? 0 : Lc;
        int ns = Lc - int(skew);
        ns = ns < 0 ? 0 : (ns > int(cnt) ? int(cnt) : ns);
        const int totL = __shfl_sync(0xffffffffu, incl, 31 - __clz(live));
        const uint32_t bar = smem_u32(B + 16);
        if (lane == 0) {
            *reinterpret_cast<int4*>(B) = make_int4(r_j, r_q, __popc(live), totL);
            if (!(w.debug_skip & 4)) mbar_arrive_expect_tx(bar, uint32_t(totL < CAP ? totL : CAP) * 8u);
        }
        __syncwarp();
        if (cnt != 0u) {
            if (Lc > 0 && !(w.debug_skip & 4)) bulk_g2s(smem_u32(B + kStHdrBytes + SLOTS * 16) + uint32_t(excl) * 8u, g_post + (first - skew), uint32_t(Lc) * 8u, bar);
            reinterpret_cast<uint4*>(B + kStHdrBytes)[__popc(live & lt_mask)] =
                make_uint4(first, cnt, uint32_t(excl + int(skew)) | (uint32_t(ns) << 16), __float_as_uint(r_w));
        }
    };

    // ---- candidate write-out: staged in shared memory, list slot reserved with one atomicAdd per task whose
    //      result is consumed one task later (as in bm25_score_kernel) ------------------------------------
    // The bound of a query is refreshed when its candidate count crosses top_k + i * step (i = 0, 1, ...):
    // between two refreshes the seen fraction of the corpus grows by a constant factor, so a query needs
    // O(log(docs / top_k)) refreshes and no per-candidate bookkeeping.
    const int ms_k = w.ts.top_k;
    const int ms_step = ms_k / 2 > 64 ? ms_k / 2 : 64;
    auto crossed = [&](int before, int after) {
        if (after < ms_k) return false;
        if (before < ms_k) return true;
        return (after - ms_k) / ms_step != (before - ms_k) / ms_step;
    };
    int pend_n = 0, pend_q = 0, pend_base = 0, stage_sel = 0;
    auto complete_pending = [&]() {
        if (pend_n == 0) return;
        const int base = __shfl_sync(0xffffffffu, pend_base, 0);
        const uint64_t* st = s_emit + (stage_sel ^ 1) * kStEmit;
        if (lane < pend_n) {
            const int slot = base + lane;
            if (slot < w.cap) w.cand[int64_t(pend_q) * w.cap + slot] = st[lane];
            else w.overflow[pend_q] = 1;
        }
        if (w.use_tau && crossed(base, base + pend_n)) {
            __syncwarp();
            const int n = (base + pend_n) < w.cap ? (base + pend_n) : w.cap;
            staged_tau_refresh(&w.ts.tau[pend_q], w.cand + int64_t(pend_q) * w.cap, n, w.ts.top_k,
                               reinterpret_cast<int*>(s_emit + (stage_sel ^ 1) * kStEmit));
        }
        pend_n = 0;
    };
    int cur_j = -1, lo = 0;
    uint32_t phase = 0;                                            // bit b: parity buffer b completes next
    int staged = 0;                                                // candidates of the current task in the stage
    // all lanes call; `pass` lanes hold a candidate (score bits vb, local doc d)
    auto emit_round = [&](int q, bool pass, int vb, int d) {
        const unsigned pm = __ballot_sync(0xffffffffu, pass);
        if (pm == 0u || (w.debug_skip & 8)) return;
        const int total = __popc(pm);
        uint64_t* st = s_emit + stage_sel * kStEmit;
        if (staged + total > kStEmit) {                            // stage full (ramp-up): write it through
            int base = 0;
            if (lane == 0) base = atomicAdd(&w.cand_count[q], staged);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < staged) {
                if (base + lane < w.cap) w.cand[int64_t(q) * w.cap + base + lane] = st[lane];
                else w.overflow[q] = 1;
            }
            __syncwarp();
            if (w.use_tau && crossed(base, base + staged)) {
                const int n = (base + staged) < w.cap ? (base + staged) : w.cap;
                staged_tau_refresh(&w.ts.tau[q], w.cand + int64_t(q) * w.cap, n, w.ts.top_k, reinterpret_cast<int*>(st));
            }
            staged = 0;
            __syncwarp();
        }
        if (pass) {
            const uint32_t key = float_to_key(__int_as_float(vb) + 0.0f);
            st[staged + __popc(pm & lt_mask)] = make_key64(key, ix.doc_base + uint32_t(lo + d));
        }
        staged += total;
    };

    // ---- C: score + read out the task staged in buffer b; false at the end of the stream -------------
    auto consume = [&](int b) -> bool {
        unsigned char* B = s_bufs + size_t(b) * buf_bytes;
        const int4 h = *reinterpret_cast<const int4*>(B);
        if (h.z < 0) return false;
        if (h.z == 0) return true;
        const int j = h.x, q = h.y, T = h.z;
        const uint32_t tau_key = w.use_tau ? ld_relaxed_u32(&w.ts.tau[q]) : w.min_key;
        if (j != cur_j) {                                           // stage this sub-range's doc lengths
            lo = j * RS;
            const int nd = (ix.n_docs - lo) < RS ? int(ix.n_docs - lo) : RS;
            // 16-byte loads (lo is a multiple of 128; the arrays are padded past n_docs)
            const uint4* src = LEN16 ? reinterpret_cast<const uint4*>(ix.doc_len16 + lo) : reinterpret_cast<const uint4*>(ix.doc_norm + lo);
            const int n16 = (nd * int(sizeof(LenT)) + 15) >> 4;
            for (int i = lane; i < n16; i += 32) reinterpret_cast<uint4*>(s_len)[i] = __ldg(src + i);
            cur_j = j;
            __syncwarp();
        }
        if (!(w.debug_skip & 4)) {
            mbar_wait(smem_u32(B + 16), (phase >> b) & 1u);
            phase ^= 1u << b;
        }

        const uint4* slots = reinterpret_cast<const uint4*>(B + kStHdrBytes);
        const int2* spost = reinterpret_cast<const int2*>(B + kStHdrBytes + SLOTS * 16);
        float* acc_lo = s_acc - lo;                                 // indexed by absolute (shard-local) doc id
        const LenT* len_lo = s_len - lo;
        const int dummy = lo + RS;                                  // idle lanes go to a slot nobody reads

        // U independent rounds of one term (doc ids are unique inside a term): all loads first, then the
        // arithmetic, then the stores — the warp has U dependency chains in flight instead of one.
        auto contrib = [&](int tfi, LenT l, float wt, float a) {
            const float tf = float(tfi);
            const float norm = LEN16 ? fmaf(float(l), c1, c0) : float(l);
            // idf*qtf*(k1+1) * tf / (tf + k1*(1-b+b*dl/avgdl)); tf + norm >= 1, rcp.approx: <= 1 ulp
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(tf + norm));
            return fmaf(wt * tf, r, a);
        };
#define MSE_ST_BATCH(U, LOADP)                                                                       \
        {                                                                                            \
            int2 p_[U]; float a_[U]; LenT l_[U];                                                     \
            _Pragma("unroll") for (int u = 0; u < U; ++u) { p_[u] = LOADP(u); }                      \
            _Pragma("unroll") for (int u = 0; u < U; ++u) { a_[u] = acc_lo[p_[u].x]; l_[u] = len_lo[p_[u].x]; } \
            _Pragma("unroll") for (int u = 0; u < U; ++u) { a_[u] = contrib(p_[u].y, l_[u], wt, a_[u]); }       \
            _Pragma("unroll") for (int u = 0; u < U; ++u) { acc_lo[p_[u].x] = a_[u]; }               \
        }
#pragma unroll 1
        for (int s = 0; s < ((w.debug_skip & 1) ? 0 : T); ++s) {
            const uint4 m = slots[s];
            const int n = int(m.y);
            const int ns = int(m.z >> 16);
            const float wt = __uint_as_float(m.w);
            const int2* ps = spost + int(m.z & 0xffffu) + lane;
            int base = 0;
#define MSE_ST_LOAD_S(u) [&]() { int2 p = ps[base + 32 * (u)]; if (base + 32 * (u) + lane >= ns) p.x = dummy; return p; }()
#pragma unroll 1
            for (; base + 64 < ns; base += 128) MSE_ST_BATCH(4, MSE_ST_LOAD_S)
            if (base + 32 < ns) MSE_ST_BATCH(2, MSE_ST_LOAD_S)
            else if (base < ns) MSE_ST_BATCH(1, MSE_ST_LOAD_S)
            if (ns < n) {                                           // tail of a slice larger than the staging buffer
                const int2* gp = g_post + m.x + lane;
                base = ns;
#define MSE_ST_LOAD_G(u) ((base + 32 * (u) + lane < n) ? __ldg(gp + base + 32 * (u)) : make_int2(dummy, 1))
#pragma unroll 1
                for (; base + 64 < n; base += 128) MSE_ST_BATCH(4, MSE_ST_LOAD_G)
                if (base + 32 < n) MSE_ST_BATCH(2, MSE_ST_LOAD_G)
                else if (base < n) MSE_ST_BATCH(1, MSE_ST_LOAD_G)
            }
            __syncwarp();                                           // next term may touch the same docs
        }

        // ---- read-out -------------------------------------------------------------------------------------
        const float tau_f = key_to_float(tau_key);
        const int tau_i = __float_as_int(tau_f);
        const bool fast = tau_i >= 0;                               // tau is +0.0 or positive: one signed compare also
                                                                    // rejects -0.0 (untouched) and every negative score
        staged = 0;
        if (w.debug_skip & 2) return true;
        if (h.w <= CAP && w.readout_mode != 0) {
            // every slice is staged whole: walk the staged doc ids, term by term
            int* acc_i = reinterpret_cast<int*>(acc_lo);
#pragma unroll 1
            for (int s = 0; s < T; ++s) {
                const uint4 m = slots[s];
                const int n = int(m.y);
                const int2* ps = spost + int(m.z & 0xffffu) + lane;
                int base = 0;
#define MSE_ST_READ(U)                                                                               \
                {                                                                                    \
                    int d_[U], v_[U]; bool ok_[U], any_ = false;                                     \
                    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                  \
                        ok_[u] = base + 32 * u + lane < n;                                           \
                        const int x = ps[base + 32 * u].x;                                           \
                        d_[u] = ok_[u] ? x : dummy;                                                  \
                    }                                                                                \
                    _Pragma("unroll") for (int u = 0; u < U; ++u) v_[u] = acc_i[d_[u]];              \
                    _Pragma("unroll") for (int u = 0; u < U; ++u) acc_i[d_[u]] = int(kUntouchedBits); \
                    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                  \
                        ok_[u] = ok_[u] && (fast ? (v_[u] >= tau_i) : (uint32_t(v_[u]) != kUntouchedBits && __int_as_float(v_[u]) >= tau_f)); \
                        any_ = any_ || ok_[u];                                                       \
                    }                                                                                \
                    if (__any_sync(0xffffffffu, any_)) {                                             \
                        _Pragma("unroll") for (int u = 0; u < U; ++u) emit_round(q, ok_[u], v_[u], d_[u] - lo); \
                    }                                                                                \
                }
#pragma unroll 1
                for (; base + 64 < n; base += 128) MSE_ST_READ(4)
                if (base + 32 < n) MSE_ST_READ(2)
                else if (base < n) MSE_ST_READ(1)
                __syncwarp();
            }
        } else {
            const int iters = RS >> 7;
            const int4* a4 = reinterpret_cast<const int4*>(s_acc);
            uint32_t flag = 0;
            if (fast) {
#pragma unroll 4
                for (int it = 0; it < iters; ++it) {
                    const int4 v = a4[it * 32 + lane];
                    flag |= uint32_t(max(max(v.x, v.y), max(v.z, v.w)) >= tau_i) << it;
                }
            } else {
#pragma unroll 2
                for (int it = 0; it < iters; ++it) {
                    const int4 v = a4[it * 32 + lane];
                    const bool p = (uint32_t(v.x) != kUntouchedBits && __int_as_float(v.x) >= tau_f) ||
                                   (uint32_t(v.y) != kUntouchedBits && __int_as_float(v.y) >= tau_f) ||
                                   (uint32_t(v.z) != kUntouchedBits && __int_as_float(v.z) >= tau_f) ||
                                   (uint32_t(v.w) != kUntouchedBits && __int_as_float(v.w) >= tau_f);
                    flag |= uint32_t(p) << it;
                }
            }
            while (__any_sync(0xffffffffu, flag != 0u)) {           // rare after the bound has ramped up
                int4 v = make_int4(0, 0, 0, 0);
                int d0 = 0;
                const bool have = flag != 0u;
                if (have) {
                    const int it = __ffs(int(flag)) - 1;
                    flag &= flag - 1u;
                    v = a4[it * 32 + lane];
                    d0 = (it * 32 + lane) * 4;
                }
                const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool pass = have && (fast ? (vv[u] >= tau_i)
                                                    : (uint32_t(vv[u]) != kUntouchedBits && __int_as_float(vv[u]) >= tau_f));
                    emit_round(q, pass, vv[u], d0 + u);
                }
            }
            const float4 z4 = make_float4(neg0, neg0, neg0, neg0);
#pragma unroll 4
            for (int it = 0; it < iters; ++it) reinterpret_cast<float4*>(s_acc)[it * 32 + lane] = z4;
        }
        __syncwarp();
        if (staged > 0) {
            complete_pending();                                     // previous task's atomic has returned by now
            if (lane == 0) pend_base = atomicAdd(&w.cand_count[q], staged);   // result consumed one task later
            pend_n = staged; pend_q = q; stage_sel ^= 1;
        }
        return true;
    };

    // ---- pipeline: records one task ahead of the copies, copies NBUF-1 tasks ahead of the scoring ----
    load_records();
#pragma unroll
    for (int d = 0; d < NBUF - 1; ++d) { stage(d); load_records(); }
    int bs = NBUF - 1, bc = 0;
    while (true) {
        stage(bs);
        load_records();
        __syncwarp();
        if (!consume(bc)) break;
        bs = (bs + 1 == NBUF) ? 0 : bs + 1;
        bc = (bc + 1 == NBUF) ? 0 : bc + 1;
    }
    complete_pending();
}

#undef MSE_ST_BATCH
#undef MSE_ST_LOAD_S
#undef MSE_ST_LOAD_G
#undef MSE_ST_READ

}  // namespace mse
