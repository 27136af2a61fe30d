// Microbenchmark: issue rate / dependent-chain latency of tcgen05.mma (kind::f16, bf16, cta_group::1, M=128) for
// several N, with all MMAs accumulating into ONE TMEM tile vs alternating between two tiles. No loads: smem is
// whatever it is (values irrelevant). One CTA per SM. Prints clocks per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench tools/umma_microbench.cu && ./umma_microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t) ((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24); }

__global__ void __launch_bounds__(64, 1) bench(int N, int iters, int two_acc, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero operands so nothing overflows
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint64_t a = make_desc(base), b = make_desc(base + 16384);
        const uint32_t idesc = make_idesc(128, N);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + ((two_acc && (i & 1)) ? 256u : 0u);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a + 2u * (i & 3)),
                         "l"(b + 2u * (i & 3)), "r"(idesc), "r"(i > 1 ? 1u : 0u)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// `groups` groups of `per_group` MMAs; optionally a tcgen05.commit (to a rotating mbarrier nobody waits on) after each
// group, as the conv kernels do per K block to release the smem stage. Does the commit serialise the groups?
__global__ void __launch_bounds__(64, 1) bench_groups(int N, int groups, int per_group, int do_commit, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[9];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint64_t a = make_desc(base), b = make_desc(base + 16384);
        const uint32_t idesc = make_idesc(128, N);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            for (int i = 0; i < per_group; ++i)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a + 2u * (i & 3)),
                             "l"(b + 2u * (i & 3)), "r"(idesc), "r"(1u)
                             : "memory");
            if (do_commit) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[g & 7])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[8])) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[8])) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// Clean issue (elect.sync, 12 MMAs unrolled, descriptors in uniform registers), a commit per group, and `delay` clocks of
// unrelated scalar work between groups: is the gap hidden behind queued MMAs or does it add to the cadence?
// fill != 0: operands hold pseudo-random bf16 instead of zeros.
__global__ void __launch_bounds__(64, 1) bench_clean(int N, int groups, int delay, int fill, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[9];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) {
        uint32_t v = 0;
        if (fill) {
            uint32_t h = (uint32_t) i * 2654435761u;
            v = 0x3c003c00u | (h & 0x007f007fu) | ((h >> 3) & 0x80008000u); // bf16 pairs around +-0.01
        }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(v));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1) {
        const uint64_t a = make_desc(base), b = make_desc(base + 16384);
        const uint32_t idesc = make_idesc(128, N);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
#pragma unroll
                for (int i = 0; i < 12; ++i)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a + 2u * (i & 3)),
                                 "l"(b + 2u * (i & 3)), "r"(idesc), "r"((g | i) ? 1u : 0u)
                                 : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[g & 7])) : "memory");
            }
            __syncwarp();
            if (delay) {
                const long long s = clock64();
                while (clock64() - s < delay) {}
            }
        }
        if (threadIdx.x == 32) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[8])) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[8])) : "memory");
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(bench_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    cudaFuncSetAttribute(bench_clean, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    for (int N : {64, 128, 256})
        for (int fill : {0, 1})
            for (int delay : {0, 100, 200, 400}) {
                bench_clean<<<148, 64, 60 * 1024>>>(N, 256, delay, fill, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("error %s\n", cudaGetErrorString(e));
                    return 1;
                }
                long long c;
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("clean issue: 256 groups x 12 MMAs + commit, N %3d, %s operands, %3d clk gap between groups: %.1f clk per group\n", N, fill ? "random" : "zero  ", delay,
                       (double) c / 256);
            }
    for (int rep = 0; rep < 1; ++rep)
        for (int N : {64, 128})
            for (int groups : {512})
                for (int commit : {0, 1}) {
                    bench_groups<<<148, 64, 60 * 1024>>>(N, groups, 12, commit, d);
                    cudaDeviceSynchronize();
                    long long c;
                    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("%d groups of 12 MMAs, N %3d, commit after each group: %d -> %lld clk, %.1f clk per group\n", groups, N, commit, c, (double) c / groups);
                }
    for (int it : {768, 6144}) {
        bench<<<148, 64, 60 * 1024>>>(64, it, 0, d);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("tight loop N 64, %d MMAs: %lld clk, %.1f clk/MMA\n", it, c, (double) c / it);
    }
    // fixed latency of a short burst: issue `iters` MMAs, commit, wait (what one K block of the conv kernels does)
    for (int N : {64, 128})
        for (int it : {1, 4, 12, 24, 48}) {
            bench<<<148, 64, 60 * 1024>>>(N, it, 0, d);
            cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("burst N %3d  %2d MMAs + commit + wait: %lld clk\n", N, it, c);
        }
    const int iters = 4096;
    for (int grid : {148})
        for (int two : {0, 1})
            for (int N : {16, 32, 64, 96, 128, 192, 256}) {
                bench<<<grid, 64, 60 * 1024>>>(N, iters, two, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("error %s\n", cudaGetErrorString(e));
                    return 1;
                }
                long long c;
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("grid %3d  N %3d  %s  %7.1f clk/MMA  (%.0f MAC/clk/SM)\n", grid, N, two ? "two accumulators" : "one accumulator ", (double) c / iters,
                       128.0 * N * 16 * iters / (double) c);
            }
    return 0;
}
