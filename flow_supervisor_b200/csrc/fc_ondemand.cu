// On-demand correlation (AlternateCorrBlock, corr.py:63-91; kernels of
// pytorch/alt_cuda_corr/correlation_kernel.cu:18-256): no volume is materialised.
// For each query and level the (2r+2)^2 dot products f1[q] . pool^l(f2)[pos] around
// floor(coords/2^l) are computed and bilinearly combined into the (2r+1)^2 taps.
//
// Layout: channels-last copies (f1t: (B,N,D); f2t_l: (B,Hl,Wl,D)) are made ONCE per
// block by fc_ondemand_prepare (the reference re-permutes both maps on every call,
// corr.py:82-83).  A warp owns a query: 8 lanes x 16 B cover one 128-byte line of a
// target's channel vector, 4 targets per load instruction; CTA = 32 consecutive
// queries so the (B, K, H, W) output tile leaves through shared memory as 128-byte rows.
#include "fc_common.cuh"

namespace fc {

constexpr int OD_THREADS = 256;      // 8 warps
constexpr int OD_QT = 32;            // queries per CTA
constexpr int OD_MAXD = 256;
constexpr int OD_MAXP = (2 * FC_MAX_RADIUS + 2) * (2 * FC_MAX_RADIUS + 2);  // 100
constexpr int OD_MAXT = (2 * FC_MAX_RADIUS + 1) * (2 * FC_MAX_RADIUS + 1);  // 81

struct OdLevel {
    const float* f2t;    // (B, Hl, Wl, D)
    int H, W;
    float inv_scale;     // 1 / 2^l applied to coords
};

struct OdParams {
    const float* f1t;    // (B, N, D)
    const float* coords;
    long long c_sb, c_sq, c_sxy;   // coords strides: sample, query, x->y
    float* out;          // (B, K, N): K = L * R * R
    float* d1;           // backward: f1 grad (B, N, D)
    float* d2;           // backward: f2 grad (B, H2, W2, D)
    const float* gout;   // backward: corr grad (B, K, N)
    int B, N, D, L, K;
    float sqrt_d;        // 0 => unscaled (alt_cuda_corr.forward semantics)
    OdLevel lv[FC_MAX_LEVELS];
};

// (B, D, N) -> (B, N, D)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int D, int N) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const float* s = src + (long long)b * D * N;
    float* t = dst + (long long)b * D * N;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int d = d0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (d < D && n < N) ? s[(long long)d * N + n] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, d = d0 + threadIdx.x;
        if (n < N && d < D) t[(long long)n * D + d] = tile[threadIdx.x][i];
    }
}

// channels-last 2x2 mean (corr.py:70-71 on fmap2), same summation order as avg_pool2d
__global__ void pool_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                 int B, int Hs, int Ws, int Hd, int Wd, int D) {
    const long long total = (long long)B * Hd * Wd * (D / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (D / 4));
        long long t = i / (D / 4);
        const int x = (int)(t % Wd); t /= Wd;
        const int y = (int)(t % Hd);
        const int b = (int)(t / Hd);
        const float4* s = reinterpret_cast<const float4*>(src) +
                          (((long long)b * Hs + 2 * y) * Ws + 2 * x) * (D / 4) + c4;
        const float4 a = s[0], bb = s[D / 4], c = s[(long long)Ws * (D / 4)], d = s[(long long)(Ws + 1) * (D / 4)];
        float4 r;
        r.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.x, bb.x), c.x), d.x), 0.25f);
        r.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.y, bb.y), c.y), d.y), 0.25f);
        r.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.z, bb.z), c.z), d.z), 0.25f);
        r.w = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.w, bb.w), c.w), d.w), 0.25f);
        reinterpret_cast<float4*>(dst)[i] = r;
    }
}

template <int RADIUS>
__global__ void __launch_bounds__(OD_THREADS)
ondemand_fwd_kernel(const OdParams P) {
    constexpr int R = 2 * RADIUS + 1, RP = R + 1, NPOS = RP * RP, NTAP = R * R;
    __shared__ float S[OD_THREADS / 32][OD_MAXP + 4];
    __shared__ float tile[OD_MAXT][OD_QT + 1];

    const int level = blockIdx.y;
    const OdLevel lv = P.lv[level];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 3, chunk = lane & 7;
    const long long Q = (long long)P.B * P.N;
    const long long gq0 = (long long)blockIdx.x * OD_QT;
    const int nj = P.D / 32;

    for (int qi = warp; qi < OD_QT; qi += OD_THREADS / 32) {
        const long long gq = gq0 + qi;
        if (gq >= Q) break;
        const long long b = gq / P.N, p = gq - b * P.N;
        const float* cptr = P.coords + b * P.c_sb + p * P.c_sq;
        const float xl = __fmul_rn(__ldg(cptr), lv.inv_scale);
        const float yl = __fmul_rn(__ldg(cptr + P.c_sxy), lv.inv_scale);
        const float xf = floorf(xl), yf = floorf(yl);
        const float dx = __fsub_rn(xl, xf), dy = __fsub_rn(yl, yf);
        const bool near_ = fabsf(xl) < 1048576.f && fabsf(yl) < 1048576.f;
        const int x0 = near_ ? (int)xf : -(1 << 24), y0 = near_ ? (int)yf : -(1 << 24);

        float4 f1r[OD_MAXD / 32];
        const float4* f1p = reinterpret_cast<const float4*>(P.f1t + gq * P.D) + chunk;
#pragma unroll
        for (int j = 0; j < OD_MAXD / 32; ++j)
            if (j < nj) f1r[j] = __ldg(f1p + j * 8);

        const float* f2b = lv.f2t + b * (long long)lv.H * lv.W * P.D;
        for (int t0 = 0; t0 < NPOS; t0 += 4) {
            const int t = t0 + sub;
            const int iy = t / RP, ix = t - iy * RP;
            const int h2 = y0 - RADIUS + iy, w2 = x0 - RADIUS + ix;
            float s = 0.f;
            if (t < NPOS && h2 >= 0 && h2 < lv.H && w2 >= 0 && w2 < lv.W) {
                const float4* f2p = reinterpret_cast<const float4*>(f2b + ((long long)h2 * lv.W + w2) * P.D) + chunk;
#pragma unroll
                for (int j = 0; j < OD_MAXD / 32; ++j)
                    if (j < nj) {
                        const float4 v = __ldg(f2p + j * 8);
                        s = fmaf(f1r[j].x, v.x, s); s = fmaf(f1r[j].y, v.y, s);
                        s = fmaf(f1r[j].z, v.z, s); s = fmaf(f1r[j].w, v.w, s);
                    }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (chunk == 0 && t < NPOS) S[warp][t] = s;
        }
        __syncwarp();
        // bilinear combination (correlation_kernel.cu:91-113 in gather form)
        const float w00 = (1.f - dy) * (1.f - dx), w01 = (1.f - dy) * dx, w10 = dy * (1.f - dx), w11 = dy * dx;
        for (int k = lane; k < NTAP; k += 32) {
            const int tx = k / R, ty = k - tx * R;           // channel = ix * R + iy (x-major)
            const float* s0 = &S[warp][ty * RP + tx];
            float v = s0[0] * w00 + s0[1] * w01 + s0[RP] * w10 + s0[RP + 1] * w11;
            if (P.sqrt_d > 0.f) v = __fdiv_rn(v, P.sqrt_d);
            tile[k][qi] = v;
        }
        __syncwarp();
    }
    __syncthreads();
    // coalesced tile store: lane <-> query
    for (int k = warp; k < NTAP; k += OD_THREADS / 32) {
        const long long gq = gq0 + lane;
        if (gq < Q) {
            const long long b = gq / P.N, p = gq - b * P.N;
            P.out[(b * P.K + (long long)level * NTAP + k) * P.N + p] = tile[k][lane];
        }
    }
}

// Backward of ONE level (alt_cuda_corr.backward, correlation_kernel.cu:122-256):
// d1[q] = sum_pos ds(pos) f2[pos];  d2[pos] += ds(pos) f1[q]  (vector reductions).
template <int RADIUS>
__global__ void __launch_bounds__(OD_THREADS)
ondemand_bwd_kernel(const OdParams P) {
    constexpr int R = 2 * RADIUS + 1, RP = R + 1, NPOS = RP * RP, NTAP = R * R;
    __shared__ float Gs[OD_THREADS / 32][OD_MAXT + 3];
    const OdLevel lv = P.lv[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 3, chunk = lane & 7;
    const long long Q = (long long)P.B * P.N;
    const int nj = P.D / 32;
    const long long gq = (long long)blockIdx.x * (OD_THREADS / 32) + warp;
    if (gq >= Q) return;
    const long long b = gq / P.N, p = gq - b * P.N;
    const float* cptr = P.coords + b * P.c_sb + p * P.c_sq;
    const float xl = __fmul_rn(__ldg(cptr), lv.inv_scale);
    const float yl = __fmul_rn(__ldg(cptr + P.c_sxy), lv.inv_scale);
    const float xf = floorf(xl), yf = floorf(yl);
    const float dx = __fsub_rn(xl, xf), dy = __fsub_rn(yl, yf);
    const bool near_ = fabsf(xl) < 1048576.f && fabsf(yl) < 1048576.f;
    const int x0 = near_ ? (int)xf : -(1 << 24), y0 = near_ ? (int)yf : -(1 << 24);

    for (int k = lane; k < NTAP; k += 32) Gs[warp][k] = __ldg(P.gout + (b * P.K + k) * P.N + p);
    __syncwarp();

    float4 f1r[OD_MAXD / 32], acc[OD_MAXD / 32];
    const float4* f1p = reinterpret_cast<const float4*>(P.f1t + gq * P.D) + chunk;
#pragma unroll
    for (int j = 0; j < OD_MAXD / 32; ++j) {
        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nj) f1r[j] = __ldg(f1p + j * 8);
    }
    const float w00 = (1.f - dy) * (1.f - dx), w01 = (1.f - dy) * dx, w10 = dy * (1.f - dx), w11 = dy * dx;
    const float* f2b = lv.f2t + b * (long long)lv.H * lv.W * P.D;
    float* d2b = P.d2 + b * (long long)lv.H * lv.W * P.D;
    for (int t0 = 0; t0 < NPOS; t0 += 4) {
        const int t = t0 + sub;
        const int iy = t / RP, ix = t - iy * RP;
        const int h2 = y0 - RADIUS + iy, w2 = x0 - RADIUS + ix;
        if (t < NPOS && h2 >= 0 && h2 < lv.H && w2 >= 0 && w2 < lv.W) {
            // position (iy, ix) feeds taps (iy, ix) w00, (iy, ix-1) w01, (iy-1, ix) w10, (iy-1, ix-1) w11
            float ds = 0.f;
            if (iy < R && ix < R) ds += w00 * Gs[warp][ix * R + iy];
            if (iy < R && ix > 0) ds += w01 * Gs[warp][(ix - 1) * R + iy];
            if (iy > 0 && ix < R) ds += w10 * Gs[warp][ix * R + iy - 1];
            if (iy > 0 && ix > 0) ds += w11 * Gs[warp][(ix - 1) * R + iy - 1];
            const long long off = ((long long)h2 * lv.W + w2) * P.D;
            const float4* f2p = reinterpret_cast<const float4*>(f2b + off) + chunk;
            float* d2p = d2b + off + chunk * 4;
#pragma unroll
            for (int j = 0; j < OD_MAXD / 32; ++j)
                if (j < nj) {
                    const float4 v = __ldg(f2p + j * 8);
                    acc[j].x = fmaf(ds, v.x, acc[j].x); acc[j].y = fmaf(ds, v.y, acc[j].y);
                    acc[j].z = fmaf(ds, v.z, acc[j].z); acc[j].w = fmaf(ds, v.w, acc[j].w);
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(d2p + j * 32),
                                 "f"(ds * f1r[j].x), "f"(ds * f1r[j].y), "f"(ds * f1r[j].z), "f"(ds * f1r[j].w)
                                 : "memory");
                }
        }
    }
    // the 4 lane groups hold partial sums for the same channels
#pragma unroll
    for (int j = 0; j < OD_MAXD / 32; ++j)
        if (j < nj) {
            float4 a = acc[j];
#pragma unroll
            for (int m = 8; m <= 16; m <<= 1) {
                a.x += __shfl_xor_sync(0xffffffffu, a.x, m); a.y += __shfl_xor_sync(0xffffffffu, a.y, m);
                a.z += __shfl_xor_sync(0xffffffffu, a.z, m); a.w += __shfl_xor_sync(0xffffffffu, a.w, m);
            }
            if (sub == 0) reinterpret_cast<float4*>(P.d1 + gq * P.D)[chunk + j * 8] = a;
        }
}

static inline unsigned grid1d(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count_cached() * 32;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

struct OdLayout {
    size_t f1t;                     // element offsets into the workspace
    size_t f2t[FC_MAX_LEVELS];
    int H[FC_MAX_LEVELS], W[FC_MAX_LEVELS];
    size_t total;
};

static bool od_layout(OdLayout& Lo, int B, int D, int H, int W, int L) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || L < 1 || L > FC_MAX_LEVELS) return false;
    size_t off = 0;
    Lo.f1t = off; off += (size_t)B * H * W * D;
    for (int l = 0; l < L; ++l) {
        Lo.H[l] = H >> l; Lo.W[l] = W >> l;
        if (Lo.H[l] < 1 || Lo.W[l] < 1) return false;
        Lo.f2t[l] = off; off += (size_t)B * Lo.H[l] * Lo.W[l] * D;
    }
    Lo.total = off;
    return true;
}

template <int RADIUS>
static void launch_od_fwd(const OdParams& P, cudaStream_t s) {
    const long long Q = (long long)P.B * P.N;
    dim3 grid((unsigned)((Q + OD_QT - 1) / OD_QT), (unsigned)P.L);
    ondemand_fwd_kernel<RADIUS><<<grid, OD_THREADS, 0, s>>>(P);
}
template <int RADIUS>
static void launch_od_bwd(const OdParams& P, cudaStream_t s) {
    const long long Q = (long long)P.B * P.N;
    ondemand_bwd_kernel<RADIUS><<<(unsigned)((Q + 7) / 8), OD_THREADS, 0, s>>>(P);
}

}  // namespace fc

using namespace fc;

extern "C" size_t fc_ondemand_workspace_bytes(int B, int D, int H, int W, int num_levels) {
    OdLayout Lo;
    if (!od_layout(Lo, B, D, H, W, num_levels)) {
        set_error("fc_ondemand_workspace_bytes: bad geometry");
        return 0;
    }
    return Lo.total * sizeof(float);
}

extern "C" int fc_ondemand_prepare(const float* fmap1, const float* fmap2,
                                   int B, int D, int H, int W, int num_levels,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    FC_REQUIRE(fmap1 && fmap2 && workspace, "fc_ondemand_prepare: null pointer");
    FC_REQUIRE(D % 32 == 0 && D <= OD_MAXD, "fc_ondemand: D=%d must be a multiple of 32 and <= %d", D, OD_MAXD);
    OdLayout Lo;
    FC_REQUIRE(od_layout(Lo, B, D, H, W, num_levels), "fc_ondemand_prepare: bad geometry");
    if (workspace_bytes < Lo.total * sizeof(float)) {
        set_error("fc_ondemand_prepare: workspace %zu < %zu bytes", workspace_bytes, Lo.total * sizeof(float));
        return FC_EWORKSPACE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    const int N = H * W;
    dim3 tb(32, 8), tg((N + 31) / 32, (D + 31) / 32, B);
    nchw_to_nhwc_kernel<<<tg, tb, 0, s>>>(fmap1, ws + Lo.f1t, D, N);
    nchw_to_nhwc_kernel<<<tg, tb, 0, s>>>(fmap2, ws + Lo.f2t[0], D, N);
    FC_LAUNCH_CHECK("nchw_to_nhwc_kernel");
    for (int l = 1; l < num_levels; ++l) {
        const long long total = (long long)B * Lo.H[l] * Lo.W[l] * (D / 4);
        pool_nhwc_kernel<<<grid1d(total, 256), 256, 0, s>>>(ws + Lo.f2t[l - 1], ws + Lo.f2t[l], B,
                                                            Lo.H[l - 1], Lo.W[l - 1], Lo.H[l], Lo.W[l], D);
        FC_LAUNCH_CHECK("pool_nhwc_kernel");
    }
    return FC_OK;
}

extern "C" int fc_ondemand_fwd(const void* workspace, const float* coords, float* out,
                               int B, int D, int H, int W, int num_levels, int radius, void* stream) {
    FC_REQUIRE(workspace && coords && out, "fc_ondemand_fwd: null pointer");
    FC_REQUIRE(D % 32 == 0 && D <= OD_MAXD, "fc_ondemand: D=%d must be a multiple of 32 and <= %d", D, OD_MAXD);
    FC_REQUIRE(radius >= 1 && radius <= FC_MAX_RADIUS, "radius %d unsupported", radius);
    OdLayout Lo;
    FC_REQUIRE(od_layout(Lo, B, D, H, W, num_levels), "fc_ondemand_fwd: bad geometry");
    const float* ws = static_cast<const float*>(workspace);
    OdParams P{};
    const int R = 2 * radius + 1;
    P.f1t = ws + Lo.f1t; P.coords = coords;
    P.c_sb = 2LL * H * W; P.c_sq = 1; P.c_sxy = (long long)H * W;
    P.out = out; P.B = B; P.N = H * W; P.D = D; P.L = num_levels; P.K = num_levels * R * R;
    P.sqrt_d = sqrtf((float)D);
    for (int l = 0; l < num_levels; ++l) {
        P.lv[l].f2t = ws + Lo.f2t[l]; P.lv[l].H = Lo.H[l]; P.lv[l].W = Lo.W[l];
        P.lv[l].inv_scale = 1.0f / (float)(1 << l);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (radius) {
        case 1: launch_od_fwd<1>(P, s); break;
        case 2: launch_od_fwd<2>(P, s); break;
        case 3: launch_od_fwd<3>(P, s); break;
        default: launch_od_fwd<4>(P, s); break;
    }
    FC_LAUNCH_CHECK("ondemand_fwd_kernel");
    return FC_OK;
}

static int altcorr_params(OdParams& P, const float* fmap1, const float* fmap2, const float* coords,
                          int B, int H1, int W1, int H2, int W2, int C, int radius) {
    FC_REQUIRE(fmap1 && fmap2 && coords, "fc_altcorr: null pointer");
    FC_REQUIRE(C % 32 == 0 && C <= OD_MAXD, "fc_altcorr: C=%d must be a multiple of 32 and <= %d", C, OD_MAXD);
    FC_REQUIRE(radius >= 1 && radius <= FC_MAX_RADIUS, "radius %d unsupported", radius);
    FC_REQUIRE(B > 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0, "fc_altcorr: bad geometry");
    const int R = 2 * radius + 1;
    P.f1t = fmap1; P.coords = coords;
    P.c_sb = 2LL * H1 * W1; P.c_sq = 2; P.c_sxy = 1;     // (B, 1, H1, W1, 2)
    P.B = B; P.N = H1 * W1; P.D = C; P.L = 1; P.K = R * R;
    P.sqrt_d = 0.f;                                      // unscaled (corr.py:91 scales afterwards)
    P.lv[0].f2t = fmap2; P.lv[0].H = H2; P.lv[0].W = W2; P.lv[0].inv_scale = 1.0f;
    return FC_OK;
}

extern "C" int fc_altcorr_fwd(const float* fmap1, const float* fmap2, const float* coords, float* corr,
                              int B, int H1, int W1, int H2, int W2, int C, int radius, void* stream) {
    OdParams P{};
    if (int e = altcorr_params(P, fmap1, fmap2, coords, B, H1, W1, H2, W2, C, radius)) return e;
    FC_REQUIRE(corr, "fc_altcorr_fwd: null output");
    P.out = corr;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (radius) {
        case 1: launch_od_fwd<1>(P, s); break;
        case 2: launch_od_fwd<2>(P, s); break;
        case 3: launch_od_fwd<3>(P, s); break;
        default: launch_od_fwd<4>(P, s); break;
    }
    FC_LAUNCH_CHECK("ondemand_fwd_kernel");
    return FC_OK;
}

extern "C" int fc_altcorr_bwd(const float* fmap1, const float* fmap2, const float* coords,
                              const float* corr_grad, float* fmap1_grad, float* fmap2_grad,
                              int B, int H1, int W1, int H2, int W2, int C, int radius, void* stream) {
    OdParams P{};
    if (int e = altcorr_params(P, fmap1, fmap2, coords, B, H1, W1, H2, W2, C, radius)) return e;
    FC_REQUIRE(corr_grad && fmap1_grad && fmap2_grad, "fc_altcorr_bwd: null pointer");
    P.gout = corr_grad; P.d1 = fmap1_grad; P.d2 = fmap2_grad;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    FC_CUDA(cudaMemsetAsync(fmap2_grad, 0, (size_t)B * H2 * W2 * C * sizeof(float), s));
    switch (radius) {
        case 1: launch_od_bwd<1>(P, s); break;
        case 2: launch_od_bwd<2>(P, s); break;
        case 3: launch_od_bwd<3>(P, s); break;
        default: launch_od_bwd<4>(P, s); break;
    }
    FC_LAUNCH_CHECK("ondemand_bwd_kernel");
    return FC_OK;
}
