// What does HBM3e give the lookup's READ pattern?  A footprint = ROWS runs of RUN bytes, 1 KB apart (row pairs of a
// level-0 map), at a random 64-byte-aligned position of a 2 GB buffer.  8 lanes x 16 B read one 128-byte run; every
// thread keeps ROWS loads in flight and 2048 threads/SM are resident, so latency is hidden and what remains is the
// memory system's rate for scattered 128/192-byte runs.   One JSON line per (ROWS, RUN).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_pattern gather_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int ROWS, int LANES>   // LANES x 16 B per run: 8 -> 128 B, 12 -> 192 B
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ buf, size_t n_granules, int n_foot, float* sink) {
    const int per_block = 256 / 16;                                  // 16 lanes reserved per footprint slot (8 or 12 used)
    const int slot = threadIdx.x / 16, l = threadIdx.x % 16;
    float acc = 0.f;
    for (int f = blockIdx.x * per_block + slot; f < n_foot; f += gridDim.x * per_block) {
        const size_t g = ((size_t)hash32((uint32_t)f) * 2654435761ull) % (n_granules - 16 * ROWS - 4);   // 64-byte granule index
        if (l < LANES) {
            uint4 v[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) v[r] = __ldg(buf + g * 4 + (size_t)r * 64 + l);               // rows 1 KB apart
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc += __uint_as_float(v[r].x ^ v[r].w);
        }
    }
    if (acc == 1.2345f) *sink = acc;
}

template <int ROWS, int LANES>
void run(const uint4* buf, size_t n_granules, float* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n_foot = 1 << 21;
    float sum = 0.f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        gather<ROWS, LANES><<<148 * 8, 256>>>(buf, n_granules, n_foot, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) sum += ms;
    }
    const double bytes = (double)n_foot * ROWS * LANES * 16;
    // 192-byte runs at 64-byte alignment touch 3 granules; 128-byte runs 2
    printf("{\"probe\": \"gather_pattern\", \"rows\": %d, \"run_bytes\": %d, \"MB\": %.1f, \"us_mean\": %.1f, \"GBps\": %.0f}\n", ROWS,
           LANES * 16, bytes / 1e6, 1e3 * sum / 5, bytes / (sum / 5 * 1e-3) / 1e9);
}

int main() {
    const size_t bytes = 2ull << 30;
    uint4* buf; float* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    const size_t n_granules = bytes / 64;
    run<6, 8>(buf, n_granules, sink);
    run<5, 8>(buf, n_granules, sink);
    run<6, 12>(buf, n_granules, sink);
    run<1, 8>(buf, n_granules, sink);
    run<12, 4>(buf, n_granules, sink);     // 64-byte runs (single granules)
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
