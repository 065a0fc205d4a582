// Microbenchmark: how fast can per-warp SMALL global->shared copies run on a B200?
//   mode 0: cp.async.bulk (UBLKCP) per slice, mbarrier completion, double buffered per warp
//   mode 1: cp.async 16 B per lane (LDGSTS), commit/wait groups, double buffered per warp
//   mode 2: ld.global.nc 16 B per lane + st.shared (synchronous)
// Each warp repeatedly fetches `slices` slices of `bytes` bytes from pseudo-random 16-byte aligned
// offsets of a large buffer (no reuse), mimicking the BM25 staged kernel.  Prints GB/s.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) copy_kernel(const uint4* __restrict__ src, uint64_t n16, int slices, int units /*16B units per slice*/,
                                                      int iters, unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_buf = slices * units * 16;
    unsigned char* my = smem + size_t(warp) * (2 * per_buf + 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(my + 2 * per_buf);
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    if (MODE == 0 && lane == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + b)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    unsigned acc = 0;
    uint32_t phase = 0;
    auto issue = [&](int it, int b) {
        unsigned char* dst = my + b * per_buf;
        if (MODE == 0) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + b)), "r"(per_buf) : "memory");
            __syncwarp();
            if (lane < slices) {
                const uint64_t off = (uint64_t(hash32(gw * 7919u + it * 131u + lane)) * 2654435761ull) % (n16 - units);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(dst + lane * units * 16)), "l"(src + off), "r"(units * 16), "r"(smem_u32(bars + b)) : "memory");
            }
        } else {
            const int total = slices * units;
            for (int u = lane; u < total; u += 32) {
                const int s = u / units, k = u - s * units;
                const uint64_t off = (uint64_t(hash32(gw * 7919u + it * 131u + s)) * 2654435761ull) % (n16 - units);
                if (MODE == 1) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + u * 16)), "l"(src + off + k) : "memory");
                } else {
                    uint4 v;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + off + k));
                    *reinterpret_cast<uint4*>(dst + u * 16) = v;
                }
            }
            if (MODE == 1) asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    issue(0, 0);
    for (int it = 0; it < iters; ++it) {
        const int b = it & 1;
        if (it + 1 < iters) issue(it + 1, b ^ 1);
        if (MODE == 0) {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(bars + b)), "r"((phase >> b) & 1u) : "memory");
            }
            phase ^= 1u << b;
        } else if (MODE == 1) {
            if (it + 1 < iters) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        } else {
            __syncwarp();
        }
        // consume: read one word per lane per 128 B (cheap), keeps the data dependency
        const unsigned char* buf = my + b * per_buf;
        for (int o = lane * 4; o < per_buf; o += 128) acc += *reinterpret_cast<const unsigned*>(buf + o);
        __syncwarp();
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int MODE>
void run(const uint4* src, uint64_t n16, int warps, int slices, int units, int iters, unsigned long long* sink, int sms) {
    const size_t smem = size_t(warps) * (2 * slices * units * 16 + 32);
    cudaFuncSetAttribute(copy_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    copy_kernel<MODE><<<sms, warps * 32, smem>>>(src, n16, slices, units, 8, sink);
    cudaEventRecord(e0);
    copy_kernel<MODE><<<sms, warps * 32, smem>>>(src, n16, slices, units, iters, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = double(sms) * warps * iters * slices * units * 16.0;
    printf("mode %d warps/SM %2d slices %d bytes/slice %5d : %8.1f GB/s  (%.3f ms, %s)\n", MODE, warps, slices, units * 16, bytes / ms / 1e6, ms,
           cudaGetErrorString(e));
}

int main() {
    const uint64_t bytes = 3ull << 30;
    uint4* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    unsigned long long* sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t n16 = bytes / 16;
    for (int warps : {8, 16}) {
        for (int units : {4, 11, 22, 44, 88, 176}) {           // 64 B ... 2816 B per slice
            const int slices = 4;
            const int iters = 2000 * 44 / units > 200 ? 2000 * 44 / units : 200;
            run<0>(src, n16, warps, slices, units, iters, sink, sms);
            run<1>(src, n16, warps, slices, units, iters, sink, sms);
            run<2>(src, n16, warps, slices, units, iters, sink, sms);
        }
    }
    return 0;
}
