// Fused tail of the feature encoder: the 1x1 output convolution of `fnet`
// (/root/reference/pytorch/core/extractor.py:145,184 `conv2`, 128 -> 256 channels) computed on the tensor cores with an
// epilogue that writes the result DIRECTLY as the build's packed operands -- K-major bf16 hi/lo rows, fmap1 rows scaled by
// 1/sqrt(D), fmap2 rows in patch order with zero pad rows (fc_build_tc.cu) -- so neither the fp32 feature maps
// (raft.py:102-103 `.float()`) nor the pack pre-pass exist.
//
//   out[token][co] = bias[co] + sum_c x[c][token] * W[co][c]          (a GEMM: M = tokens, N = D = 256, K = C = 128)
//
// One CTA per SM, persistent over tiles of 128 tokens of one frame (cta_group::1, UMMA M = 128, N = D):
//   warp 0      TMA producer: the weights once (K-major bf16 hi/lo, SWIZZLE_128B, resident), then per tile and k-block a
//               [64 channels][128 tokens] fp32 box of the activations (NCHW is token-contiguous: no transpose pass);
//   warp 1      TMEM allocator + single-thread MMA issuer: hi*hi + lo*hi + hi*lo into one fp32 accumulator
//               (FC_MATH_TC_3XBF16; hi*hi only in FC_MATH_TC_BF16), two 256-column accumulators;
//   warps 2-5   thread = token: convert the fp32 column of the staged box to bf16 hi/lo and write the row of the K-major
//               A operand (16-byte chunks, XOR-swizzled like the TMA would), one tile AHEAD of the epilogue they also run:
//               TMEM -> + bias -> scale -> split -> 2 x 512-byte rows of the packed operands.
#include "fc_build.cuh"

namespace fc {

constexpr int FT_TOK = 128;                       // tokens per tile (UMMA M)
constexpr int FT_BK = 64;                         // channels per k-block (one 128-byte swizzle row of bf16)
constexpr int FT_THREADS = 192;                   // warp 0 TMA, warp 1 MMA, warps 2-5 workers

__device__ __forceinline__ void umma1_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma1_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 v(__float2bfloat16_rn(a), __float2bfloat16_rn(b));
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

struct FeatParams {
    __nv_bfloat16* a_hi; __nv_bfloat16* a_lo; __nv_bfloat16* b_hi; __nv_bfloat16* b_lo;
    const float* bias;         // [D]
    int B, C, D, N, NP, H, W, Wp;
    int tiles_per_frame, n_tiles;
    int three_pass;
    float prescale;
};

template <int KBN>   // k-blocks: C / 64
__global__ void __launch_bounds__(FT_THREADS, 1)
fnet_tail_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w_hi,
                 const __grid_constant__ CUtensorMap map_w_lo, const FeatParams P) {
    extern __shared__ __align__(1024) uint8_t ft_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ft_raw) + 1023) & ~uintptr_t(1023));
    const int wblk = P.D * FT_BK * 2;                                  // bytes of one weight k-block plane: [D rows][128 B]
    constexpr int ABLK = FT_TOK * FT_BK * 2;                           // 16 KB: one A k-block plane
    uint8_t* w_hi = smem;                                              // [KBN][D][64] bf16
    uint8_t* w_lo = w_hi + KBN * wblk;
    uint8_t* a_hi = w_lo + KBN * wblk;                                 // [KBN][128][64] bf16
    uint8_t* a_lo = a_hi + KBN * ABLK;
    float* xs = reinterpret_cast<float*>(a_lo + KBN * ABLK);           // [64 channels][128 tokens] fp32
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + FT_BK * FT_TOK * 4);
    uint64_t* w_full = bars;                    // 1
    uint64_t* x_full = bars + 1;                // 1
    uint64_t* x_empty = bars + 2;               // 1
    uint64_t* a_full = bars + 3;                // KBN (<= 2)
    uint64_t* a_empty = bars + 5;               // KBN
    uint64_t* t_full = bars + 7;                // 2
    uint64_t* t_empty = bars + 9;               // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        mbar_init(x_full, 1);
        mbar_init(x_empty, 4);
        for (int i = 0; i < 2; ++i) { mbar_init(a_full + i, 4); mbar_init(a_empty + i, 1); mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int first = blockIdx.x, stride = gridDim.x;
    const int n_local = first < P.n_tiles ? (P.n_tiles - first + stride - 1) / stride : 0;
    const int n_parts = P.three_pass ? 2 : 1;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            mbar_expect_tx(w_full, (uint32_t)(n_parts * KBN * wblk));
            for (int kb = 0; kb < KBN; ++kb) {
                tma_load_2d(w_hi + kb * wblk, &map_w_hi, w_full, kb * FT_BK, 0);
                if (P.three_pass) tma_load_2d(w_lo + kb * wblk, &map_w_lo, w_full, kb * FT_BK, 0);
            }
            int xc = 0;
            for (int j = 0; j < n_local; ++j) {
                const int tile = first + j * stride;
                const int f = tile / P.tiles_per_frame, p0 = (tile - f * P.tiles_per_frame) * FT_TOK;
                for (int kb = 0; kb < KBN; ++kb, ++xc) {
                    if (xc > 0) mbar_wait(x_empty, (uint32_t)(xc - 1) & 1u);
                    mbar_expect_tx(x_full, (uint32_t)(FT_BK * FT_TOK * 4));
                    tma_load_2d(xs, &map_x, x_full, p0, f * P.C + kb * FT_BK);      // tokens beyond N: zero fill
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(FT_TOK, P.D);
            mbar_wait(w_full, 0);
            for (int j = 0; j < n_local; ++j) {
                const int buf = j & 1;
                mbar_wait(t_empty + buf, ((uint32_t)(j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256);
                for (int kb = 0; kb < KBN; ++kb) {
                    mbar_wait(a_full + kb, (uint32_t)j & 1u);
                    tc_fence_after();
                    const uint32_t ah = smem_u32(a_hi + kb * ABLK), al = smem_u32(a_lo + kb * ABLK);
                    const uint32_t bh = smem_u32(w_hi + kb * wblk), bl = smem_u32(w_lo + kb * wblk);
#pragma unroll
                    for (int k = 0; k < FT_BK / 16; ++k) {
                        umma1_bf16(d_addr, umma_desc_sw128(ah + k * 32), umma_desc_sw128(bh + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                        if (P.three_pass) {
                            umma1_bf16(d_addr, umma_desc_sw128(al + k * 32), umma_desc_sw128(bh + k * 32), idesc, 1u);
                            umma1_bf16(d_addr, umma_desc_sw128(ah + k * 32), umma_desc_sw128(bl + k * 32), idesc, 1u);
                        }
                    }
                    umma1_commit(a_empty + kb);                // this k-block of A may be overwritten once these MMAs retire
                }
                umma1_commit(t_full + buf);
            }
        }
    } else {
        // ================= workers: thread = token =================
        const int quarter = warp & 3;                              // TMEM lane quarter this warp may read
        const int r = quarter * 32 + lane;                         // token row inside the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        int xc = 0;

        auto convert = [&](int j) {
            for (int kb = 0; kb < KBN; ++kb, ++xc) {
                mbar_wait(x_full, (uint32_t)xc & 1u);
                uint32_t hi[FT_BK / 2], lo[FT_BK / 2];
#pragma unroll
                for (int c2 = 0; c2 < FT_BK / 2; ++c2) {
                    const float v0 = xs[(2 * c2) * FT_TOK + r], v1 = xs[(2 * c2 + 1) * FT_TOK + r];
                    const float h0 = bf16_round(v0), h1 = bf16_round(v1);
                    hi[c2] = pack2(h0, h1);
                    lo[c2] = pack2(v0 - h0, v1 - h1);
                }
                if (j > 0) mbar_wait(a_empty + kb, (uint32_t)(j - 1) & 1u);   // the previous tile's MMAs have read this block
                uint4* rh = reinterpret_cast<uint4*>(a_hi + kb * ABLK + r * 128);
                uint4* rl = reinterpret_cast<uint4*>(a_lo + kb * ABLK + r * 128);
#pragma unroll
                for (int c = 0; c < 8; ++c) {                      // 16-byte chunk c (channels 8c .. 8c+7) at position c ^ (r % 8)
                    rh[c ^ (r & 7)] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    if (P.three_pass) rl[c ^ (r & 7)] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                }
                fence_proxy_async_smem();                          // generic-proxy writes -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) {
                    // (in program order after the stores above, which consume every value loaded from the staged box:
                    // the next box cannot overtake those loads)
                    mbar_arrive(x_empty);
                    mbar_arrive(a_full + kb);
                }
            }
        };

        auto epilogue = [&](int j) {
            const int tile = first + j * stride;
            const int f = tile / P.tiles_per_frame, ti = tile - f * P.tiles_per_frame;
            const int p = ti * FT_TOK + r;
            const bool valid = p < P.N;
            const bool second = f >= P.B;                          // frames of image 2: target operand, patch order
            const int b = second ? f - P.B : f;
            long long row = 0;
            if (valid) row = second ? (long long)b * P.NP + tile_off(p / P.W, p % P.W, P.Wp) : (long long)b * P.N + p;
            __nv_bfloat16* oh = (second ? P.b_hi : P.a_hi) + row * P.D;
            __nv_bfloat16* ol = (second ? P.b_lo : P.a_lo) + row * P.D;
            const float scale = second ? 1.0f : P.prescale;
            if (second && ti == 0 && P.NP > P.N) {
                // pad rows of this sample's target operand (y >= H or x >= W in patch order) hold zeros
                const int wt = (warp - 2) * 32 + lane;
                for (int q = wt; q < P.NP; q += 128) {
                    int y, x;
                    tile_inv(q, P.Wp, y, x);
                    if (y >= P.H || x >= P.W) {
                        uint4* zh = reinterpret_cast<uint4*>(P.b_hi + ((long long)b * P.NP + q) * P.D);
                        uint4* zl = reinterpret_cast<uint4*>(P.b_lo + ((long long)b * P.NP + q) * P.D);
                        for (int i = 0; i < P.D / 8; ++i) {
                            zh[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (P.three_pass) zl[i] = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                }
            }
            const int buf = j & 1;
            mbar_wait(t_full + buf, (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            for (int c = 0; c < P.D / 32; ++c) {
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)(buf * 256 + c * 32), v);
                tmem_ld_wait();
                uint32_t h[16], l[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float v0 = (v[2 * i] + __ldg(P.bias + c * 32 + 2 * i)) * scale;
                    const float v1 = (v[2 * i + 1] + __ldg(P.bias + c * 32 + 2 * i + 1)) * scale;
                    const float h0 = bf16_round(v0), h1 = bf16_round(v1);
                    h[i] = pack2(h0, h1);
                    l[i] = pack2(v0 - h0, v1 - h1);
                }
                if (valid) {
                    uint4* dh = reinterpret_cast<uint4*>(oh + c * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dh[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                    if (P.three_pass) {
                        uint4* dl = reinterpret_cast<uint4*>(ol + c * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) dl[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + buf);
        };

        if (n_local > 0) convert(0);
        for (int j = 0; j < n_local; ++j) {
            if (j + 1 < n_local) convert(j + 1);                   // the next tile's operand while this tile's MMAs run
            epilogue(j);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// weights (D, C) fp32 -> [w_hi D*C][w_lo D*C] bf16 K-major + [bias D] fp32
__global__ void fnet_tail_prepare_kernel(const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* hi,
                                         __nv_bfloat16* lo, float* bias_out, int n, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = w[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    if (i < D) bias_out[i] = bias ? bias[i] : 0.f;
}

static size_t packed_w_bytes(int C, int D) { return (size_t)D * C * 2 * 2 + (size_t)D * 4; }

static bool fnet_tail_shape_ok(int C, int D, int N) {
    return (C == 64 || C == 128) && D % 64 == 0 && D >= 64 && D <= 256 && N % 4 == 0;
}

int fnet_tail_pack(const FeatSource& src, const TcPacked& dst, cudaStream_t s) {
    const int C = src.C, D = dst.D, B = dst.B, N = dst.N;
    FC_REQUIRE(fnet_tail_shape_ok(C, D, N), "fc_build_from_fnet_tail: needs C in {64, 128}, D %% 64 == 0, D <= 256 and H*W %% 4 == 0 "
               "(got C=%d D=%d H*W=%d)", C, D, N);
    FC_REQUIRE((reinterpret_cast<uintptr_t>(src.x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(src.packed_w) & 127u) == 0,
               "fc_build_from_fnet_tail: x must be 16-byte and the packed weights 128-byte aligned");
    const uint8_t* pw = static_cast<const uint8_t*>(src.packed_w);
    const __nv_bfloat16* w_hi = reinterpret_cast<const __nv_bfloat16*>(pw);
    const __nv_bfloat16* w_lo = w_hi + (size_t)D * C;
    const float* bias = reinterpret_cast<const float*>(pw + (size_t)D * C * 4);
    const bool three = dst.a_lo != nullptr;

    CUtensorMap map_x, map_wh, map_wl;
    {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)2 * B * C};
        cuuint64_t str[1] = {(cuuint64_t)N * 4};
        cuuint32_t box[2] = {(cuuint32_t)FT_TOK, (cuuint32_t)FT_BK};
        if (int e = encode_tiled_cached(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src.x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE)) return e;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)D};
        cuuint64_t str[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {(cuuint32_t)FT_BK, (cuuint32_t)D};
        if (int e = encode_tiled_cached(&map_wh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_hi, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
        if (int e = encode_tiled_cached(&map_wl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_lo, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return e;
    }
    FeatParams P{};
    P.a_hi = dst.a_hi; P.a_lo = dst.a_lo; P.b_hi = dst.b_hi; P.b_lo = dst.b_lo; P.bias = bias;
    P.B = B; P.C = C; P.D = D; P.N = N; P.NP = dst.NP; P.H = dst.H; P.W = dst.W; P.Wp = dst.Wp;
    P.tiles_per_frame = (N + FT_TOK - 1) / FT_TOK;
    P.n_tiles = 2 * B * P.tiles_per_frame;
    P.three_pass = three ? 1 : 0;
    P.prescale = dst.prescale;
    const int kbn = C / FT_BK;
    const size_t smem = 1024 + (size_t)2 * kbn * D * FT_BK * 2 + (size_t)2 * kbn * FT_TOK * FT_BK * 2 + (size_t)FT_BK * FT_TOK * 4 + 256;
    const int n_sm = sm_count_cached();
    const int grid = P.n_tiles < n_sm ? P.n_tiles : n_sm;
    if (kbn == 1) {
        FC_SMEM_ATTR_GROW((fnet_tail_kernel<1>), smem);
        fnet_tail_kernel<1><<<grid, FT_THREADS, smem, s>>>(map_x, map_wh, map_wl, P);
    } else {
        FC_SMEM_ATTR_GROW((fnet_tail_kernel<2>), smem);
        fnet_tail_kernel<2><<<grid, FT_THREADS, smem, s>>>(map_x, map_wh, map_wl, P);
    }
    FC_LAUNCH_CHECK("fnet_tail_kernel");
    return FC_OK;
}

}  // namespace fc

using namespace fc;

extern "C" size_t fc_fnet_tail_weights_bytes(int C, int D) {
    if (C <= 0 || D <= 0) return 0;
    return packed_w_bytes(C, D);
}

extern "C" int fc_fnet_tail_supported(int C, int D, int H, int W) {
    return (H > 0 && W > 0 && fnet_tail_shape_ok(C, D, H * W)) ? 1 : 0;
}

extern "C" int fc_fnet_tail_prepare(const float* weight, const float* bias, int C, int D, void* packed, size_t packed_bytes,
                                    void* stream) {
    FC_REQUIRE(weight && packed, "fc_fnet_tail_prepare: null pointer");
    FC_REQUIRE(C > 0 && D > 0 && packed_bytes >= packed_w_bytes(C, D), "fc_fnet_tail_prepare: buffer %zu < %zu bytes", packed_bytes,
               packed_w_bytes(C, D));
    __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(packed);
    __nv_bfloat16* lo = hi + (size_t)D * C;
    float* b = reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + (size_t)D * C * 4);
    const int n = D * C;
    fnet_tail_prepare_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, bias, hi, lo, b, n, D);
    FC_LAUNCH_CHECK("fnet_tail_prepare_kernel");
    return FC_OK;
}
