// Does tcgen05.mma take its A operand from tensor memory the way we think?  D[128 x 64] = A[128 x 64] * B[64 x 64]^T with
// A written to TMEM by tcgen05.st (lane = row m, 32-bit column c holds bf16 elements k = 2c (low half), 2c + 1 (high half)),
// B in shared memory (K-major, SWIZZLE_128B, written by hand like a TMA would), cta_group::1, kind::f16.
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tmem_a_probe tmem_a_probe.cu && ./tmem_a_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int K) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
    const int N = 64;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // B: [N rows][K] bf16 K-major, k-blocks of 64, SW128: row n at n*128 within a block, chunk j at j ^ (n & 7)
    for (int i = threadIdx.x; i < N * K / 8; i += 128) {
        const int n = i / (K / 8), j8 = i % (K / 8), kb = j8 / 8, j = j8 % 8;
        *(uint4*)(smem + kb * (N * 128) + n * 128 + ((j ^ (n & 7)) << 4)) = *(const uint4*)(B + (size_t)n * K + 8 * j8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // A: thread r = lane (row) r; K/2 columns starting at column 64 (D occupies columns 0..63)
    const uint32_t a_col0 = 64;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = *(const uint32_t*)(A + (size_t)r * K + 2 * (c0 + i));
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + a_col0 + c0;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < K / 16; ++k) {
            const int kb = k / 4, kk = k % 4;
            const uint64_t bd = desc_sw128(smem_u32(smem + kb * (N * 128)) + kk * 32);
            const uint32_t a_addr = tmem + a_col0 + k * 8;                  // 16 bf16 = 8 columns per k-step
            const uint32_t acc = k ? 1u : 0u;
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tmem),
                         "r"(a_addr), "l"(bd), "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra DN; bra W; DN: }" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[(size_t)r * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
    const int M = 128, N = 64, K = 128;
    __nv_bfloat16 *hA = new __nv_bfloat16[M * K], *hB = new __nv_bfloat16[N * K];
    float* ref = new float[M * N];
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16(rnd());
    for (int i = 0; i < N * K; ++i) hB[i] = __float2bfloat16(rnd());
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double a = 0;
            for (int k = 0; k < K; ++k) a += (double)__bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
            ref[m * N + n] = (float)a;
        }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    const int smem = 1024 + N * K * 2 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, K);
    cudaError_t e = cudaDeviceSynchronize();
    float* out = new float[M * N];
    cudaMemcpy(out, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < M * N; ++i) worst = fmax(worst, fabs(out[i] - ref[i]));
    printf("{\"probe\": \"tmem_a\", \"cuda\": \"%s\", \"max_abs_err\": %.3e, \"out0\": %.5f, \"ref0\": %.5f, \"out_last\": %.5f, \"ref_last\": %.5f}\n",
           cudaGetErrorString(e), worst, out[0], ref[0], out[M * N - 1], ref[M * N - 1]);
    return 0;
}
