// How fast does ONE thread issue tcgen05.mma (kind::f16, bf16 -> f32, M = 128, cta_group::1) on a B200, by operand
// placement (A in shared memory / A in tensor memory) and N, with nothing else running on the SM?  And how fast do
// the epilogue warps drain TMEM (tcgen05.ld 32x32b.x32)?  The operands are zero-filled shared / tensor memory; no
// loads.  Cycles are SM clocks (clock64), ns from globaltimer: their ratio is the clock the SM really ran at.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I modern-search-engines-project_b200/csrc \
//        -o tools/microbench/mma_rate tools/microbench/mma_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"

using namespace mse;

struct Out { long long cycles; long long ns; };

// warp-converged issue: every lane executes the instruction stream, one elected lane issues
__device__ __forceinline__ void mma_ss_elect(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts_elect(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bd), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}

// mode 0: SS (A and B descriptors), mode 1: TS (A in TMEM).  commit_every: MMAs per tcgen05.commit (0 = one at the end)
__global__ void __launch_bounds__(192, 1)
mma_rate_kernel(int mode, int n, int iters, int commit_every, int ld_warps, Out* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_done, s_step;
    __shared__ uint32_t s_tmem;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) { mbarrier_init(&s_done, 1); mbarrier_init(&s_step, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = s_tmem;
    if (warp >= 2) {                                    // zero the A region of TMEM (columns 0..255)
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = 0u;
        for (int c = 0; c < 256; c += 32) tmem_st_32x32b_x32(tm + (uint32_t((warp & 3) * 32) << 16) + c, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    long long c0 = 0, t0 = 0;
    if (warp == 1 && lane == 0 && mode < 2) {
        const uint32_t idesc = umma_idesc_bf16_f32(128, n);
        const uint32_t a_addr = smem_addr(smem), b_addr = smem_addr(smem + 32 * 1024);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            const int k = it & 3;
            const uint64_t bd = umma_desc_sw128(b_addr + k * 32);
            if (mode == 0) tcgen05_mma_bf16(tm + 256, umma_desc_sw128(a_addr + k * 32), bd, idesc, 1u);
            else tcgen05_mma_bf16_ts(tm + 256, tm + uint32_t((it & 15) * 8), bd, idesc, 1u);
            const int ce = commit_every < 0 ? -commit_every : commit_every;
            if (ce && (it % ce) == ce - 1) {
                tcgen05_commit(&s_step);
                if (commit_every < 0) { mbarrier_wait(&s_step, ph); ph ^= 1u; }     // negative: wait for completion (latency)
            }
        }
        tcgen05_commit(&s_done);
        mbarrier_wait(&s_done, 0u);
        const long long c1 = clock64();
        long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (blockIdx.x == 0) { out->cycles = c1 - c0; out->ns = t1 - t0; }
    } else if (warp == 1 && (mode == 3 || mode == 4)) {
        // converged warp, four MMAs per k-block unrolled, commit per k-block
        const uint32_t idesc = umma_idesc_bf16_f32(128, n);
        const uint64_t ad0 = umma_desc_sw128(smem_addr(smem)), bd0 = umma_desc_sw128(smem_addr(smem + 32 * 1024));
        if (lane == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); c0 = clock64(); }
        for (int it = 0; it < iters; it += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 3) mma_ss_elect(tm + 256, ad0 + uint64_t(k * 2), bd0 + uint64_t(k * 2), idesc);
                else mma_ts_elect(tm + 256, tm + uint32_t(((it & 12) + k) * 8), bd0 + uint64_t(k * 2), idesc);
            }
            if (commit_every) commit_elect(&s_step);
        }
        commit_elect(&s_done);
        mbarrier_wait(&s_done, 0u);
        if (lane == 0) {
            const long long c1 = clock64();
            long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (blockIdx.x == 0) { out->cycles = c1 - c0; out->ns = t1 - t0; }
        }
    } else if (warp >= 2 && warp < 2 + ld_warps && mode == 2) {
        // TMEM drain rate: each warp reads its lane quarter, 32 columns per instruction
        uint32_t acc = 0;
        if (threadIdx.x == 64) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); c0 = clock64(); }
        for (int it = 0; it < iters; ++it) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tm + (uint32_t((warp & 3) * 32) << 16) + uint32_t((it & 7) * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= r[i];
        }
        if (acc == 0x12345u) out[1].cycles = acc;
        if (threadIdx.x == 64) {
            const long long c1 = clock64();
            long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (blockIdx.x == 0) { out->cycles = c1 - c0; out->ns = t1 - t0; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

int main() {
    Out* d_out;
    cudaMalloc(&d_out, 2 * sizeof(Out));
    const int smem = 97 * 1024 + 1024;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 48 * 400;
    auto run = [&](const char* what, int grid, int mode, int n, int commit_every, int ld_warps, int its) {
        Out h{};
        for (int rep = 0; rep < 2; ++rep) {
            mma_rate_kernel<<<grid, 192, smem>>>(mode, n, its, commit_every, ld_warps, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); exit(1); }
        }
        cudaMemcpy(&h, d_out, sizeof(Out), cudaMemcpyDeviceToHost);
        const double flops = mode == 2 ? 0.0 : 2.0 * 128 * n * 16 * double(its) * grid;
        printf("%-44s grid %3d  %8.1f cycles/op  %8.1f ns/op  clock %.0f MHz  %7.1f TFLOP/s\n", what, grid,
               double(h.cycles) / its, double(h.ns) / its, 1e3 * double(h.cycles) / double(h.ns), flops / (double(h.ns) * 1e-9) * 1e-12);
    };
    for (int grid : {1, 148}) {
        for (int n : {64, 128, 256}) {
            char nm[64];
            snprintf(nm, sizeof nm, "SS M128 N%d K16, commit/4", n);      run(nm, grid, 0, n, 4, 0, iters);
            snprintf(nm, sizeof nm, "TS M128 N%d K16, commit/4", n);      run(nm, grid, 1, n, 4, 0, iters);
        }
            for (int n : {64, 128, 256}) {
            char nm[64];
            snprintf(nm, sizeof nm, "SS M128 N%d K16, converged, commit/4", n);      run(nm, grid, 3, n, 4, 0, iters);
            snprintf(nm, sizeof nm, "TS M128 N%d K16, converged, commit/4", n);      run(nm, grid, 4, n, 4, 0, iters);
        }
        run("SS M128 N256, one commit at the end", grid, 0, 256, 0, 0, iters);
        run("TS M128 N256, one commit at the end", grid, 1, 256, 0, 0, iters);
        run("TS M128 N64, one commit at the end", grid, 1, 64, 0, 0, iters);
        run("TS M128 N256, commit+wait every MMA", grid, 1, 256, -1, 0, 2000);
        run("TS M128 N256, commit+wait every 4 MMAs", grid, 1, 256, -4, 0, 4000);
        run("tcgen05.ld x32, 4 warps", grid, 2, 64, 0, 4, 20000);
    }
    return 0;
}
