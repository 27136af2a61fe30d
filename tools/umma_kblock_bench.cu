// Tensor-pipe floor of ONE K block (64 channels x one tap) of conv_umma_kernel's issue patterns, with the A operand read
// either through the canonical descriptor (8-row groups 1024-aligned, SBO 1024) or through the halo descriptor of the 3x3
// halo mode (start shifted by (ky*10+kx) pixel rows, SBO 1280): does the unaligned operand cost tensor-pipe time?
// No loads, no epilogue: smem holds zeros. One CTA per SM, 148 CTAs. Prints clocks per K block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_kblock_bench tools/umma_kblock_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
    return (uint64_t) ((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// terms: 3 = A_hi x [B_hi;B_lo] (N = 2n) + A_lo x B_hi (N = n); 2 = A_hi x B, A_lo x B (N = n); 1 = A_hi x B
__global__ void __launch_bounds__(64, 1) bench(int terms, int n, int halo, int kblocks, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[9];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0));
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
        // A planes: hi at base, lo at base + 24 KB (halo tile 10 x 18 pixels x 128 B = 23 040 B); B at base + 48 KB (up to 64 KB)
        const uint64_t a_hi0 = make_desc(base, halo ? 1280 : 1024), a_lo0 = make_desc(base + 24576, halo ? 1280 : 1024);
        const uint64_t b = make_desc(base + 49152, 1024);
        const uint32_t id_cat = make_idesc(128, terms == 3 ? 2 * n : n), id = make_idesc(128, n);
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
            const long long t0 = clock64();
            int tap = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const uint32_t off = halo ? (uint32_t) (((tap / 3) * 10 + tap % 3) * 128) >> 4 : 0u;
                const uint64_t a_hi = a_hi0 + off, a_lo = a_lo0 + off;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma(tmem, a_hi + 2u * j, b + 2u * j, id_cat, 1u);
                    if (terms >= 2) mma(tmem, a_lo + 2u * j, b + 2u * j, id, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[kb & 7])) : "memory");
                tap = tap == 8 ? 0 : tap + 1;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[8])) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[8])) : "memory");
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// The same loop with what conv_umma_kernel's issue thread does besides the MMAs, as compile-time variants (scalar work between
// K blocks is NOT hidden behind queued MMAs, so the variants must not add runtime branching of their own):
//   ORDER 0: A_hi x cat, A_lo x B_hi alternating per K step      ORDER 1: the kernel's order, cat x4 then lo x4
//   FENCE: tcgen05.fence::after_thread_sync per K block
//   WAIT 0 none | 1 mbarrier.test_wait (completed phase) after the first MMA | 2 mbarrier.try_wait (completed phase) before the first MMA
template <int ORDER, int FENCE, int WAIT>
__global__ void __launch_bounds__(64, 1) bench3(int n, int kblocks, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[10];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 10; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[9])) : "memory"); // bars[9]: phase 0 complete
    }
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0));
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
        const uint64_t a_hi0 = make_desc(base, 1280), a_lo0 = make_desc(base + 24576, 1280), b = make_desc(base + 49152, 1024);
        const uint32_t id_cat = make_idesc(128, 2 * n), id = make_idesc(128, n), done = smem_u32(&bars[9]);
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
            const long long t0 = clock64();
            uint32_t off = 0, bad = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                const uint64_t a_hi = a_hi0 + off, a_lo = a_lo0 + off;
                if (WAIT == 2) {
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(done) : "memory");
                    bad += ok ^ 1u;
                }
                if (FENCE) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (ORDER == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        mma(tmem, a_hi + 2u * j, b + 2u * j, id_cat, 1u);
                        mma(tmem, a_lo + 2u * j, b + 2u * j, id, 1u);
                        if (j == 0 && WAIT == 1) {
                            uint32_t ok;
                            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(done) : "memory");
                            bad += ok ^ 1u;
                        }
                    }
                } else {
                    mma(tmem, a_hi, b, id_cat, 1u);
                    if (WAIT == 1) {
                        uint32_t ok;
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(done) : "memory");
                        bad += ok ^ 1u;
                    }
#pragma unroll
                    for (int j = 1; j < 4; ++j) mma(tmem, a_hi + 2u * j, b + 2u * j, id_cat, 1u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma(tmem, a_lo + 2u * j, b + 2u * j, id, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[kb & 7])) : "memory");
                off = off == 176u ? 0u : off + 8u; // walks (and wraps) the halo offsets
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[8])) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[8])) : "memory");
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0 + (bad ? 1000000000LL : 0);
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}
// The row-window kernel's stage (stems: IC <= 8): the A operand is read through a SWIZZLE_NONE K-major descriptor whose K-chunk stride
// (LBO) is 16 B - chunk j of row m is pixel m + j of a dense pixel row. `ksteps` K steps per stage (7x7 stride 2: 4), each
// A_hi x [B_hi;B_lo] (N = 2n) + A_lo x B_hi (N = n); B panels are SWIZZLE_128B as in the kernel. Is this operand form as fast as the swizzled one?
__global__ void __launch_bounds__(64, 1) bench_window(int n, int ksteps, int stages, int swizzled_a, long long* out, int commit_every = 1) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[9];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (200 * 1024) / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4u * i), "r"(0));
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
        // window descriptor: LBO 16 B, SBO 128 B, version 1, layout type 0 (no swizzle); or the canonical swizzled one for comparison
        const uint64_t a_hi0 = swizzled_a ? make_desc(base, 1024) : ((uint64_t) ((base >> 4) & 0x3FFFu) | (1ull << 16) | (8ull << 32) | (1ull << 46));
        const uint64_t a_lo0 = a_hi0 + (4608u >> 4);
        const uint64_t b     = make_desc(base + 49152, 1024);
        const uint32_t id_cat = make_idesc(128, 2 * n), id = make_idesc(128, n);
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
            const long long t0 = clock64();
            for (int st = 0; st < stages; ++st) {
                for (int q = 0; q < ksteps; ++q) {
                    const uint32_t ao = swizzled_a ? 2u * (q & 3) : 2u * q; // window: next K step = 2 pixels further; swizzled: +32 B
                    mma(tmem, a_hi0 + ao, b + 2u * (q & 3), id_cat, 1u);
                    mma(tmem, a_lo0 + ao, b + 2u * (q & 3), id, 1u);
                }
                if ((st + 1) % commit_every == 0)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[st & 7])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[8])) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[8])) : "memory");
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int ORDER, int FENCE, int WAIT> static void run3(long long* d) {
    cudaFuncSetAttribute(bench3<ORDER, FENCE, WAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    for (int n : {64, 128}) {
        bench3<ORDER, FENCE, WAIT><<<148, 64, 202 * 1024>>>(n, 900, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = -1;
        if (e == cudaSuccess) cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("3-term n_blk %3d halo | order %s | fence.after %d | %s: %7.1f clk per K block%s\n", n, ORDER ? "cat x4, lo x4 (kernel)" : "hi/lo alternating     ", FENCE,
               WAIT == 0 ? "no barrier op          " : (WAIT == 1 ? "test_wait after 1st MMA" : "try_wait before MMAs   "), (double) c / 900, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    const int kblocks = 900;
    printf("# clocks per K block (64 channels): 4 K steps x {terms} MMAs, M = 128; ideal = tensor math at 4096 MAC/clk/SM\n");
    for (int terms : {3, 2, 1})
        for (int n : {64, 128, 256}) {
            if (terms == 3 && n > 128) continue;
            for (int halo : {0, 1}) {
                bench<<<148, 64, 202 * 1024>>>(terms, n, halo, kblocks, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("error %s\n", cudaGetErrorString(e));
                    return 1;
                }
                long long c;
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                const double ideal = 4.0 * terms * 128.0 * n * 16 / 4096.0;
                printf("terms %d n_blk %3d %-6s: %7.1f clk per K block   (math floor %5.0f, %.0f%% of it)\n", terms, n, halo ? "halo" : "plain", (double) c / kblocks, ideal,
                       100.0 * ideal * kblocks / (double) c);
            }
        }
    cudaFuncSetAttribute(bench_window, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    for (int sw : {1, 0})
        for (int n : {64, 32, 16}) {
            bench_window<<<148, 64, 202 * 1024>>>(n, 4, 900, sw, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = -1;
            if (e == cudaSuccess) cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("row-window stage, 4 K steps, 3-term, n_blk %2d, A operand %s: %7.1f clk per stage (math floor %4.0f)%s\n", n,
                   sw ? "SWIZZLE_128B (canonical)       " : "SWIZZLE_NONE window (LBO 16 B) ", (double) c / 900, 4.0 * 3 * 128 * n * 16 / 4096.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    // how much issue-thread time does a tcgen05.commit cost? stages of `ksteps` K steps (2 MMAs each, n_blk 64), one commit every `ce` stages
    for (int ks : {0, 1, 2, 4})
        for (int ce : {1, 2, 7, 900}) {
            bench_window<<<148, 64, 202 * 1024>>>(64, ks, 900, 0, d, ce);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = -1;
            if (e == cudaSuccess) cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("commit cost: %d K steps per stage (math floor %3d clk), commit every %3d stages: %7.1f clk per stage%s\n", ks, ks * 112, ce, (double) c / 900, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    run3<0, 0, 0>(d);
    run3<1, 0, 0>(d);
    run3<1, 1, 0>(d);
    run3<1, 0, 1>(d);
    run3<1, 0, 2>(d);
    run3<1, 1, 1>(d);
    run3<0, 1, 1>(d);
    return 0;
}
