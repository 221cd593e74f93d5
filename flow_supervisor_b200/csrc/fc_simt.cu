// FC_MATH_FP32 arithmetic: the all-pairs contraction and its two transposed gradients
// on the CUDA cores (fp32 FMA, sequential-k accumulation -- the arithmetic of the
// reference's SGEMM, corr.py:58), plus the stand-alone pooling / gradient-fold kernels
// this mode uses.  The tensor-core modes live in fc_build_tc.cu; this file is the exact
// mode and the numerical yardstick the tcgen05 kernels are checked against on the GPU.
//
// One kernel template, three operand mappings:
//   OP_FWD : vol0[b,p,q'] = sum_d f1[b,d,p] * f2pad[b,d,q'] / sqrt(D)
//   OP_DF1 : dF1[b,d,p]   = sum_q' f2pad[b,d,q'] * G0[b,p,q'] / sqrt(D)
//   OP_DF2 : dF2[b,d,q]   = sum_p  f1[b,d,p]    * G0[b,p,q'(q)] / sqrt(D)
// where q' = tile_off(y, x) runs over the PADDED, patch-ordered target index space of a
// level-0 map (fc_common.cuh): pad rows/columns read as 0 and are written as 0, so the
// pyramid's pad invariant holds by construction and q' IS the memory offset.
#include "fc_common.cuh"

namespace fc {

constexpr int BM = 128, BN = 128, BK = 16, PITCH = 132, GEMM_THREADS = 256;

enum { OP_FWD = 0, OP_DF1 = 1, OP_DF2 = 2 };

struct GemmParams {
    const float* f1;     // (B, D, N)
    const float* f2;     // (B, D, N)
    float* vol;          // level 0 of the (gradient) pyramid: (B*N, NP)
    float* dout;         // dF1 or dF2 (B, D, N)
    int D, N, H, W, Wp, NP;   // NP = Hp * Wp
    float sqrt_d;
};

// padded target index -> real target index, or -1 on a pad column
__device__ __forceinline__ int unpad(int qp, int H, int W, int Wp) {
    int y, x;
    tile_inv(qp, Wp, y, x);
    return (x < W && y < H) ? y * W + x : -1;
}

template <int OP>
__global__ void __launch_bounds__(GEMM_THREADS)
simt_gemm_kernel(const GemmParams P) {
    __shared__ __align__(16) float As[BK][PITCH];
    __shared__ __align__(16) float Bs[BK][PITCH];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int M = (OP == OP_FWD) ? P.N : P.D;
    const int Nn = (OP == OP_DF1) ? P.N : P.NP;
    const int K = (OP == OP_FWD) ? P.D : (OP == OP_DF1 ? P.NP : P.N);

    const float* f1 = P.f1 + (long long)b * P.D * P.N;
    const float* f2 = P.f2 + (long long)b * P.D * P.N;
    const float* G = P.vol + (long long)b * P.N * P.NP;

    // element e of a BK x 128 tile handled by this thread in pass i: e = tid + i*256
    //   "row-of-k" sources  (src[k][m]): k = e / 128, m = e % 128   (coalesced along m)
    //   "k-contiguous" sources (src[m][k]): m = e / 16,  k = e % 16 (64-byte runs along k)
    float ra[8], rb[8];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * GEMM_THREADS;
            if (OP == OP_FWD) {
                const int k = e >> 7, m = e & 127;
                const int kk = k0 + k, mm = m0 + m, nn = n0 + m;
                ra[i] = (kk < K && mm < M) ? __ldg(f1 + (long long)kk * P.N + mm) : 0.f;
                int src = (kk < K && nn < Nn) ? unpad(nn, P.H, P.W, P.Wp) : -1;
                rb[i] = src >= 0 ? __ldg(f2 + (long long)kk * P.N + src) : 0.f;
            } else if (OP == OP_DF1) {
                const int m = e >> 4, k = e & 15;
                const int kk = k0 + k, mm = m0 + m, nn = n0 + m;
                int src = (kk < K && mm < M) ? unpad(kk, P.H, P.W, P.Wp) : -1;
                ra[i] = src >= 0 ? __ldg(f2 + (long long)mm * P.N + src) : 0.f;
                rb[i] = (kk < K && nn < Nn) ? __ldg(G + (long long)nn * P.NP + kk) : 0.f;
            } else {
                const int m = e >> 4, k = e & 15;
                ra[i] = (k0 + k < K && m0 + m < M) ? __ldg(f1 + (long long)(m0 + m) * P.N + k0 + k) : 0.f;
                const int kb = e >> 7, n = e & 127;
                rb[i] = (k0 + kb < K && n0 + n < Nn) ? __ldg(G + (long long)(k0 + kb) * P.NP + n0 + n) : 0.f;
            }
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * GEMM_THREADS;
            if (OP == OP_FWD) {
                As[e >> 7][e & 127] = ra[i];
                Bs[e >> 7][e & 127] = rb[i];
            } else if (OP == OP_DF1) {
                As[e & 15][e >> 4] = ra[i];
                Bs[e & 15][e >> 4] = rb[i];
            } else {
                As[e & 15][e >> 4] = ra[i];
                Bs[e >> 7][e & 127] = rb[i];
            }
        }
    };

    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: the reference divides the finished sum by sqrt(D) (corr.py:60)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * 64 + tx * 4;
            float4 v = make_float4(__fdiv_rn(acc[i][h * 4 + 0], P.sqrt_d), __fdiv_rn(acc[i][h * 4 + 1], P.sqrt_d),
                                   __fdiv_rn(acc[i][h * 4 + 2], P.sqrt_d), __fdiv_rn(acc[i][h * 4 + 3], P.sqrt_d));
            if (OP == OP_FWD) {
                if (n < Nn)   // NP is a multiple of 8: whole float4 in range
                    *reinterpret_cast<float4*>(P.vol + ((long long)b * P.N + m) * P.NP + n) = v;
            } else {
                float* dst = P.dout + ((long long)b * P.D + m) * P.N;
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j >= Nn) continue;
                    if (OP == OP_DF1) {
                        dst[n + j] = vv[j];
                    } else {
                        int q = unpad(n + j, P.H, P.W, P.Wp);
                        if (q >= 0) dst[q] = vv[j];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- pooling / fold
// level l -> l+1 over the padded output index space (pads written as 0).
// (a + b + c + d) * 0.25 in this order reproduces ATen's avg_pool2d bit for bit
// (oracle/corr_spec.py::pool_pyramid).
__global__ void pool_level_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                  long long Q, int Hps, int Wps, int Hd, int Wd, int Hpd, int Wpd) {
    const int msrc = Hps * Wps, mdst = Hpd * Wpd;
    const long long total = Q * mdst;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / mdst;
        int y, x;
        tile_inv((int)(i - q * mdst), Wpd, y, x);
        float v = 0.f;
        if (x < Wd && y < Hd) {
            const float* s = src + q * msrc + tile_off(2 * y, 2 * x, Wps);   // rows 2y, 2y+1 share a patch
            const float2 r0 = *reinterpret_cast<const float2*>(s);
            const float2 r1 = *reinterpret_cast<const float2*>(s + 8);
            v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(r0.x, r0.y), r1.x), r1.y), 0.25f);
        }
        dst[i] = v;
    }
}

// avg_pool2d backward, in place: every parent gives a quarter to its 4 children.
__global__ void fold_level_kernel(const float* __restrict__ parent, float* __restrict__ child,
                                  long long Q, int Hpc, int Wpc, int Hp, int Wp_, int Hpp, int Wpp) {
    const int mpar = Hpp * Wpp, mch = Hpc * Wpc;
    const long long total = Q * Hp * Wp_;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wp_);
        const long long t = i / Wp_;
        const int y = (int)(t % Hp);
        const long long q = t / Hp;
        const float g = 0.25f * parent[q * mpar + tile_off(y, x, Wpp)];
        float* c = child + q * mch + tile_off(2 * y, 2 * x, Wpc);
        float2 r0 = *reinterpret_cast<float2*>(c);
        float2 r1 = *reinterpret_cast<float2*>(c + 8);
        r0.x += g; r0.y += g; r1.x += g; r1.y += g;
        *reinterpret_cast<float2*>(c) = r0;
        *reinterpret_cast<float2*>(c + 8) = r1;
    }
}

static inline unsigned grid_for(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count_cached() * 32;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

// Pad row (y = H when H is odd) of every query map of levels [first, last]: the lookup's TMA
// boxes read whole row pairs, so the row must hold zeros (the pyramid invariant).
__global__ void zero_pad_row_kernel(float* __restrict__ lvl, long long Q, int H, int Wp, int msize) {
    const int per = Wp / 4;                                    // float4 per pad row
    const long long total = Q * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / per;
        const int x = (int)(i - q * per) * 4;
        *reinterpret_cast<float4*>(lvl + q * msize + tile_off(H, x, Wp)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

int simt_zero_pad_rows(float* pyramid, const Pyramid& pyr, int first, int last, cudaStream_t s) {
    const long long Q = (long long)pyr.B * pyr.N;
    for (int l = first; l <= last && l < pyr.L; ++l) {
        const Level& a = pyr.lv[l];
        if (a.Hp == a.H) continue;
        zero_pad_row_kernel<<<grid_for(Q * (a.Wp / 4), 256), 256, 0, s>>>(pyramid + a.offset, Q, a.H, a.Wp, a.Hp * a.Wp);
        FC_LAUNCH_CHECK("zero_pad_row_kernel");
    }
    return FC_OK;
}

// computes levels [first_level, L) from the level below each
int simt_pool_levels(float* pyramid, const Pyramid& pyr, int first_level, cudaStream_t s) {
    const long long Q = (long long)pyr.B * pyr.N;
    for (int l = (first_level < 1 ? 1 : first_level) - 1; l + 1 < pyr.L; ++l) {
        const Level &a = pyr.lv[l], &c = pyr.lv[l + 1];
        pool_level_kernel<<<grid_for(Q * c.Hp * c.Wp, 256), 256, 0, s>>>(
            pyramid + a.offset, pyramid + c.offset, Q, a.Hp, a.Wp, c.H, c.W, c.Hp, c.Wp);
        FC_LAUNCH_CHECK("pool_level_kernel");
    }
    return FC_OK;
}

int simt_build(const float* f1, const float* f2, float* pyramid, const Pyramid& pyr,
               int D, int H, int W, cudaStream_t s) {
    GemmParams P{};
    P.f1 = f1; P.f2 = f2; P.vol = pyramid + pyr.lv[0].offset; P.dout = nullptr;
    P.D = D; P.N = pyr.N; P.H = H; P.W = W; P.Wp = pyr.lv[0].Wp; P.NP = pyr.lv[0].Hp * P.Wp;
    P.sqrt_d = sqrtf((float)D);
    dim3 grid((P.NP + BN - 1) / BN, (P.N + BM - 1) / BM, pyr.B);
    simt_gemm_kernel<OP_FWD><<<grid, GEMM_THREADS, 0, s>>>(P);
    FC_LAUNCH_CHECK("simt_gemm_kernel<FWD>");
    return simt_pool_levels(pyramid, pyr, 1, s);
}

int simt_fold(float* gpyr, const Pyramid& pyr, cudaStream_t s) {
    const long long Q = (long long)pyr.B * pyr.N;
    for (int l = pyr.L - 1; l >= 1; --l) {
        const Level &p = pyr.lv[l], &c = pyr.lv[l - 1];
        fold_level_kernel<<<grid_for(Q * p.H * p.W, 256), 256, 0, s>>>(
            gpyr + p.offset, gpyr + c.offset, Q, c.Hp, c.Wp, p.H, p.W, p.Hp, p.Wp);
        FC_LAUNCH_CHECK("fold_level_kernel");
    }
    return FC_OK;
}

int simt_build_bwd(float* gpyr, const float* f1, const float* f2, float* d1, float* d2,
                   const Pyramid& pyr, int D, int H, int W, cudaStream_t s) {
    GemmParams P{};
    P.f1 = f1; P.f2 = f2; P.vol = gpyr + pyr.lv[0].offset;
    P.D = D; P.N = pyr.N; P.H = H; P.W = W; P.Wp = pyr.lv[0].Wp; P.NP = pyr.lv[0].Hp * P.Wp;
    P.sqrt_d = sqrtf((float)D);
    if (d1) {
        P.dout = d1;
        dim3 grid((P.N + BN - 1) / BN, (D + BM - 1) / BM, pyr.B);
        simt_gemm_kernel<OP_DF1><<<grid, GEMM_THREADS, 0, s>>>(P);
        FC_LAUNCH_CHECK("simt_gemm_kernel<DF1>");
    }
    if (d2) {
        P.dout = d2;
        dim3 grid((P.NP + BN - 1) / BN, (D + BM - 1) / BM, pyr.B);
        simt_gemm_kernel<OP_DF2><<<grid, GEMM_THREADS, 0, s>>>(P);
        FC_LAUNCH_CHECK("simt_gemm_kernel<DF2>");
    }
    return FC_OK;
}

}  // namespace fc
