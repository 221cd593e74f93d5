// Probe: cp.reduce.async.bulk.tensor.3d .add on fp32, in-bounds / negative / partially OOB boxes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, int mode) {
    extern __shared__ __align__(1024) float sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = (unsigned)__cvta_generic_to_shared(sm);
        if (mode == 0)
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(&m), "r"(s), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(&m), "r"(s), "r"(c0), "r"(c1), "r"(c2) : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
}
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int Wp2 = 256, HP2 = 28, Q = 64;                  // [query][row pair][2*Wp floats]
    float* d; cudaMalloc(&d, sizeof(float) * Wp2 * HP2 * Q); cudaMemset(d, 0, sizeof(float) * Wp2 * HP2 * Q);
    cuuint64_t dims[3] = {Wp2, HP2, Q}; cuuint64_t str[2] = {Wp2 * 4, (cuuint64_t)Wp2 * HP2 * 4};
    cuuint32_t box[3] = {32, 5, 1}, es[3] = {1, 1, 1};
    for (int variant = 0; variant < 2; ++variant) {
        CUtensorMap m;
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, variant ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode variant %d -> %d\n", variant, (int)r);
        int cases[5][3] = {{32, 3, 1}, {240, 26, 4}, {0, 0, 63}, {32, -2, 3}, {-16, 3, 2}};
        for (int mode = 0; mode < 2; ++mode)
            for (int c = 0; c < 5; ++c) {
                cudaMemset(d, 0, sizeof(float) * Wp2 * HP2 * Q);
                k<<<1, 128, 8192>>>(m, cases[c][0], cases[c][1], cases[c][2], mode);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<float> h(Wp2 * HP2 * Q);
                cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
                double sum = 0; for (float v : h) sum += v;
                printf("variant %d mode %s case (%d,%d,%d): %s sum=%.0f\n", variant, mode ? "store" : "reduce", cases[c][0], cases[c][1], cases[c][2],
                       cudaGetErrorString(e), sum);
                if (e != cudaSuccess) return 1;
            }
    }
    return 0;
}
