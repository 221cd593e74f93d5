// Which store pattern does HBM3e take at full rate?  The build epilogue writes, per CTA tile,
// 128 query rows x 1 KB where rows are 28 KB apart (query-major pyramid); tools/probe_bounds.py
// showed the epilogue alone (no MMAs) needs 637 us for 2.07 GB while fill_ writes the same
// bytes in 276 us.  This probe times the bare patterns (no tensor work) over a level-0-sized
// buffer (56 320 rows x 7168 floats):
//   0 seq        : tile-interleaved layout, 128 KB contiguous per unit, coalesced 512 B per warp instruction
//   1 rowstride  : query-major layout, thread <-> row, 8 chunks of 128 B per row and tile (the epilogue today)
//   2 rowrun     : query-major layout, warp writes 1 KB contiguous per row, row after row
//   3 ilv-thread : tile-interleaved layout, thread <-> row (1 KB pitch), 8 chunks of 128 B
//   4 rowstride, tiles visited 4 at a time per row (4 KB per row before moving on)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROWS = 56320, NP = 7168, TILE = 256, NT = NP / TILE, RB = 128;

__device__ __forceinline__ void st128(float* p, float v) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
        asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(p + 8 * g), "f"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) pattern(float* out, int units) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u0 = (int)((long long)units * blockIdx.x / gridDim.x), u1 = (int)((long long)units * (blockIdx.x + 1) / gridDim.x);
    for (int u = u0; u < u1; ++u) {
        const int rb = u / NT, t = u - rb * NT;
        const float v = (float)u;
        if (MODE == 0) {
            float* base = out + ((long long)u * RB + warp * 32) * TILE;           // this warp's 32 KB
#pragma unroll 4
            for (int i = 0; i < 64; ++i)
                *reinterpret_cast<float4*>(base + (i * 32 + lane) * 4) = make_float4(v, v, v, v);
        } else if (MODE == 1 || MODE == 4) {
            const int row = rb * RB + warp * 32 + lane;
            int tt = t;
            if (MODE == 4) tt = t;                                              // same order: units already walk t fastest
            float* base = out + (long long)row * NP + tt * TILE;
#pragma unroll
            for (int c = 0; c < 8; ++c) st128(base + c * 32, v);
        } else if (MODE == 2) {
            for (int r = 0; r < 32; ++r) {
                const int row = rb * RB + warp * 32 + r;
                float* base = out + (long long)row * NP + t * TILE;
                asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(base + lane * 8), "f"(v) : "memory");
            }
        } else if (MODE == 3) {
            float* base = out + ((long long)u * RB + warp * 32 + lane) * TILE;
#pragma unroll
            for (int c = 0; c < 8; ++c) st128(base + c * 32, v);
        }
    }
}

// MODE 5: query-major, but the CTA owns ONE row block and a warp owns one row at a time for the
// whole map: 28 KB contiguous per warp before it moves to the next row (the limit case)
__global__ void __launch_bounds__(128, 1) rowwhole(float* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 4 + warp; row < ROWS; row += gridDim.x * 4) {
        float* base = out + (long long)row * NP;
        for (int i = 0; i < NP / 256; ++i)
            asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(base + i * 256 + lane * 8), "f"(1.f) : "memory");
    }
}

// The epilogue's real mix: level 0 as in mode 1 plus the pooled levels.  POOLED 1 = today (every tile writes
// 32-byte half patches of level 1; level 2 every 2nd, level 3 every 4th tile), 2 = stashed (whole 64-byte
// patches: level 1 every 2nd tile as 512 contiguous bytes, level 2 every 4th, level 3 every 8th).
template <int L0, int POOLED>
__global__ void __launch_bounds__(128, 1) epilogue_mix(float* l0, float* l1, float* l2, float* l3, int units) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u0 = (int)((long long)units * blockIdx.x / gridDim.x), u1 = (int)((long long)units * (blockIdx.x + 1) / gridDim.x);
    auto v8 = [](float* p, float v) {
        asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(p), "f"(v) : "memory");
    };
    for (int u = u0; u < u1; ++u) {
        const int rb = u / NT, t = u - rb * NT;
        const float v = (float)u;
        const long long row = rb * RB + warp * 32 + lane;
        if (L0) {
            float* base = l0 + row * NP + t * TILE;
#pragma unroll
            for (int c = 0; c < 8; ++c) st128(base + c * 32, v);
        }
        float* r1 = l1 + row * 1792; float* r2 = l2 + row * 448; float* r3 = l3 + row * 96;
        if (POOLED == 1) {
            if (t < 27)
#pragma unroll
                for (int c = 0; c < 8; ++c) v8(r1 + (t >> 1) * 128 + c * 16 + (t & 1) * 8, v);
            if ((t & 1) && (t >> 1) < 13)
#pragma unroll
                for (int c = 0; c < 4; ++c) v8(r2 + (t >> 2) * 64 + c * 16 + ((t >> 1) & 1) * 8, v);
            if ((t & 3) == 3 && (t >> 2) < 6)
#pragma unroll
                for (int c = 0; c < 2; ++c) v8(r3 + (t >> 3) * 32 + c * 16 + ((t >> 2) & 1) * 8, v);
        } else if (POOLED == 2) {
            if (t & 1)
#pragma unroll
                for (int c = 0; c < 16; ++c) v8(r1 + (t >> 1) * 128 + c * 8, v);
            if ((t & 3) == 3)
#pragma unroll
                for (int c = 0; c < 8; ++c) v8(r2 + (t >> 2) * 64 + c * 8, v);
            if ((t & 7) == 7 || t == NT - 1)
#pragma unroll
                for (int c = 0; c < 4; ++c) v8(r3 + (t >> 3) * 32 + c * 8, v);
        }
    }
}

// MODE 1 with a padded row pitch: does the 28 KB (= 7 x 4096 B) pitch camp on L2 slices / DRAM channels?
__global__ void __launch_bounds__(128, 1) rowstride_pitch(float* out, int units, int pitch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u0 = (int)((long long)units * blockIdx.x / gridDim.x), u1 = (int)((long long)units * (blockIdx.x + 1) / gridDim.x);
    for (int u = u0; u < u1; ++u) {
        const int rb = u / NT, t = u - rb * NT;
        const int row = rb * RB + warp * 32 + lane;
        float* base = out + (long long)row * pitch + t * TILE;
#pragma unroll
        for (int c = 0; c < 8; ++c) st128(base + c * 32, (float)u);
    }
}

int main() {
    float* buf;
    const size_t bytes = (size_t)ROWS * NP * 4;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    const int units = (ROWS / RB) * NT;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"seq (interleaved layout, coalesced)", "rowstride (today: thread<->row, 8 x 128 B)",
                           "rowrun (warp writes 1 KB per row)", "interleaved layout, thread<->row 1 KB pitch",
                           "rowstride again", "rowwhole (28 KB per warp-row)"};
    for (int grid : {148, 296, 592}) {
        for (int mode = 0; mode < 6; ++mode) {
            float best = 1e9f, sum = 0.f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                switch (mode) {
                    case 0: pattern<0><<<grid, 128>>>(buf, units); break;
                    case 1: pattern<1><<<grid, 128>>>(buf, units); break;
                    case 2: pattern<2><<<grid, 128>>>(buf, units); break;
                    case 3: pattern<3><<<grid, 128>>>(buf, units); break;
                    case 4: pattern<4><<<grid, 128>>>(buf, units); break;
                    case 5: rowwhole<<<grid, 128>>>(buf); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) { sum += ms; best = ms < best ? ms : best; }
            }
            printf("{\"probe\": \"store_pattern\", \"grid\": %d, \"mode\": %d, \"what\": \"%s\", \"GB\": %.3f, \"us_mean\": %.1f, \"us_best\": %.1f, \"GBps_mean\": %.0f}\n",
                   grid, mode, names[mode], bytes / 1e9, 1e3 * sum / 5, 1e3 * best, bytes / (sum / 5 * 1e-3) / 1e9);
        }
    }
    {
        float *l1, *l2, *l3;
        cudaMalloc(&l1, (size_t)ROWS * 1792 * 4); cudaMalloc(&l2, (size_t)ROWS * 448 * 4); cudaMalloc(&l3, (size_t)ROWS * 96 * 4);
        const char* mn[] = {"level 0 + pooled today (32 B half patches)", "level 0 + pooled stashed (64 B patches, contiguous)",
                            "level 0 only", "pooled today only", "pooled stashed only"};
        for (int mode = 0; mode < 5; ++mode) {
            float sum = 0.f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                switch (mode) {
                    case 0: epilogue_mix<1, 1><<<148, 128>>>(buf, l1, l2, l3, units); break;
                    case 1: epilogue_mix<1, 2><<<148, 128>>>(buf, l1, l2, l3, units); break;
                    case 2: epilogue_mix<1, 0><<<148, 128>>>(buf, l1, l2, l3, units); break;
                    case 3: epilogue_mix<0, 1><<<148, 128>>>(buf, l1, l2, l3, units); break;
                    case 4: epilogue_mix<0, 2><<<148, 128>>>(buf, l1, l2, l3, units); break;
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) sum += ms;
            }
            printf("{\"probe\": \"epilogue_mix\", \"what\": \"%s\", \"us_mean\": %.1f}\n", mn[mode], 1e3 * sum / 5);
        }
    }
    {
        float* big;
        cudaMalloc(&big, (size_t)ROWS * (NP + 1024) * 4);
        for (int pad : {0, 32, 64, 96, 160, 288, 544, 1024, 0}) {
            float sum = 0.f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                rowstride_pitch<<<148, 128>>>(big, units, NP + pad);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) sum += ms;
            }
            printf("{\"probe\": \"rowstride_pitch\", \"pad_floats\": %d, \"us_mean\": %.1f, \"GBps\": %.0f}\n", pad, 1e3 * sum / 5,
                   bytes / (sum / 5 * 1e-3) / 1e9);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
