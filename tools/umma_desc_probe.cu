// Probe of tcgen05 shared-memory descriptor semantics on B200 (no documentation reachable offline):
//   (1) SWIZZLE_128B K-major operand whose start address is shifted by r0 rows (128 B each), with / without the
//       descriptor's base_offset field: which physical 16-byte pieces does the tensor core read?
//   (2) SWIZZLE_NONE K-major operand with LBO = 16 B (K chunks OVERLAP the next row): is a sliding window legal?
// Every 16-byte piece of the A region holds its own index (element 0 = idx & 255, element 1 = idx >> 8, exact in bf16);
// B selects k = 8n / 8n+1 so that D[m][2n] / D[m][2n+1] return the piece index that logical (row m, chunk n) resolved to.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_desc_probe tools/umma_desc_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(uint64_t desc_hi_bits, uint32_t a_byte_off, int b_swizzled, float* out, int b_fp16 = 0) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen        = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A region: 48 KB of pieces; piece i at byte 16*i
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gen);
    for (int i = threadIdx.x; i < 3072; i += blockDim.x) {
        for (int e = 0; e < 8; ++e) A[i * 8 + e] = __float2bfloat16(0.0f);
        A[i * 8 + 0] = __float2bfloat16((float) (i & 255));
        A[i * 8 + 1] = __float2bfloat16((float) (i >> 8));
    }
    // B region at +48 KB: N = 16 rows x K = 16 (we only issue ONE K=16 MMA): B[n][k] = 1 where k == 8*(n/2) + (n&1)
    //   n = 0 -> k0 (chunk0 elem0), n = 1 -> k1 (chunk0 elem1), n = 2 -> k8 (chunk1 elem0), n = 3 -> k9 (chunk1 elem1)
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(gen + 49152);
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) B[i] = __float2bfloat16(0.0f);
    __syncthreads();
    if (threadIdx.x < 4) {
        const int n = threadIdx.x, k = 8 * (n / 2) + (n & 1);
        // B stored as SW128 K-major rows of 128 B: row n, chunk c = k/8 at physical chunk c ^ (n & 7)
        const int chunk = k / 8, phys = b_swizzled ? (chunk ^ (n & 7)) : chunk;
        B[n * 64 + phys * 8 + (k & 7)] = __float2bfloat16(1.0f);
        // (3) mixed operand formats: B holds the FP16 value 1 + 2^-10 (0x3C01; read as bf16 it would be ~0.0079)
        if (b_fp16) reinterpret_cast<uint16_t*>(B)[n * 64 + phys * 8 + (k & 7)] = 0x3C01;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint64_t a_desc = desc_hi_bits | (uint64_t) (((base + a_byte_off) >> 4) & 0x3FFFu);
        const uint64_t b_desc = (uint64_t) (((base + 49152) >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        // a_format bits [7,10), b_format bits [10,13): 0 = F16, 1 = BF16
        const uint32_t idesc  = (1u << 4) | (1u << 7) | ((b_fp16 ? 0u : 1u) << 10) | ((uint32_t) (16 >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a_desc), "l"(b_desc),
                     "r"(idesc), "r"(0u)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t) (warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int j = 0; j < 4; ++j) out[m * 4 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

static void run(const char* what, uint64_t hi_bits, uint32_t a_off, float* d) {
    probe<<<1, 128, 60 * 1024>>>(hi_bits, a_off, 1, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: ERROR %s\n", what, cudaGetErrorString(e));
        exit(1);
    }
    float h[512];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s\n  row: (chunk0 piece, chunk1 piece) as [row*8+chunk]:", what);
    for (int m = 0; m < 20; ++m) {
        const int p0 = (int) h[m * 4 + 0] + 256 * (int) h[m * 4 + 1], p1 = (int) h[m * 4 + 2] + 256 * (int) h[m * 4 + 3];
        printf(" m%d:(%d.%d,%d.%d)", m, p0 / 8, p0 % 8, p1 / 8, p1 % 8);
    }
    printf("\n");
}

static void run_mixed(float* d) {
    const uint64_t SW128 = (1ull << 16) | (1ull << 46) | (2ull << 61);
    probe<<<1, 128, 60 * 1024>>>(SW128 | (64ull << 32), 0, 1, d, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("MIXED bf16(A) x fp16(B): ERROR %s\n", cudaGetErrorString(e));
        exit(1);
    }
    float h[512];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int ok = 1;
    printf("MIXED A = bf16, B = fp16 (idesc a_format 1, b_format 0): D[m][0] vs (piece & 255) * (1 + 2^-10):");
    for (int m = 0; m < 128; ++m) {
        const float want = (float) ((m * 8 + (0 ^ (m & 7))) & 255) * (1.0f + 1.0f / 1024.0f);
        if (m < 6) printf(" m%d: %.6f (want %.6f)", m, h[m * 4 + 0], want);
        if (h[m * 4 + 0] != want) ok = 0;
    }
    printf("\n  => mixed-format MMA %s\n", ok ? "EXACT: legal and computed as fp16 x bf16" : "MISMATCH");
}

int main() {
    float* d;
    cudaMalloc(&d, 512 * sizeof(float));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    const uint64_t SW128 = (1ull << 16) | (1ull << 46) | (2ull << 61);
    // (1) SWIZZLE_128B: aligned reference, then start shifted by r0 rows with and without base_offset, SBO = 1024 and 2048
    run("SW128 start+0 rows, SBO 1024, base_offset 0 (reference: row m chunk c -> piece m*8 + (c ^ (m&7)))", SW128 | (64ull << 32), 0, d);
    for (int r0 = 1; r0 <= 3; ++r0) {
        char buf[128];
        snprintf(buf, sizeof buf, "SW128 start+%d rows, SBO 1024, base_offset 0", r0);
        run(buf, SW128 | (64ull << 32), r0 * 128, d);
        snprintf(buf, sizeof buf, "SW128 start+%d rows, SBO 1024, base_offset %d", r0, r0);
        run(buf, SW128 | (64ull << 32) | ((uint64_t) r0 << 49), r0 * 128, d);
    }
    run("SW128 start+1 rows, SBO 2048 (pitch 16 rows), base_offset 1", SW128 | (128ull << 32) | (1ull << 49), 128, d);
    run("SW128 start+1 rows, SBO 1280 (pitch 10 rows), base_offset 1", SW128 | (80ull << 32) | (1ull << 49), 128, d);
    run("SW128 start+1 rows, SBO 1280 (pitch 10 rows), base_offset 0", SW128 | (80ull << 32), 128, d);
    // (2) SWIZZLE_NONE K-major: canonical (LBO = stride between K chunks, SBO = stride between 8-row groups)
    const uint64_t NOSW = (1ull << 46);
    run("NOSW LBO 2048 B, SBO 128 B (chunk-major planes of 128 rows x 16 B): expect piece m + 128*c", NOSW | (128ull << 16) | (8ull << 32), 0, d);
    run("NOSW LBO 16 B, SBO 128 B (OVERLAPPING sliding window): expect piece m + c", NOSW | (1ull << 16) | (8ull << 32), 0, d);
    run("NOSW LBO 16 B, SBO 256 B (stride-2 rows?): expect piece 2*(m/8)*8.. ", NOSW | (1ull << 16) | (16ull << 32), 0, d);
    // halo mode of conv_umma_kernel: start shifted by (ky * 10 + kx) pixel rows, SBO = 10 rows; expect piece (m/8*10 + m%8 + r0)*8 + (c ^ (row & 7))
    for (int r0 : {1, 2, 10, 11, 12, 21, 22}) {
        char buf[160];
        snprintf(buf, sizeof buf, "HALO SW128 start+%d rows, SBO 1280, base_offset 0: row m -> smem row %d + (m/8)*10 + m%%8", r0, r0);
        run(buf, SW128 | (80ull << 32), r0 * 128, d);
    }
    run_mixed(d); // LAST: a refused operand-format combination is a sticky 'illegal instruction' error
    return 0;
}
