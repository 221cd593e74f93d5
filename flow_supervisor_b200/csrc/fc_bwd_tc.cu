// Tensor-core backward of CorrBlock.__init__ (autograd of corr.py:21-27,52-60; driven by
// pytorch/train.py:273,277) for sm_100a:
//
//   dC = fold(G) / sqrt(D);   dfmap1[b,:,p] = sum_q dC[b,p,q] fmap2[b,:,q];
//                             dfmap2[b,:,q] = sum_p dC[b,p,q] fmap1[b,:,p]
//
//   fold + pack   : ONE pass over the gradient pyramid.  A CTA owns one query row: it folds the
//                   coarse levels into level 0 on the fly (avg_pool2d backward: every parent
//                   gives a quarter to its 4 children, coarsest level first -- same arithmetic
//                   as fold_level_kernel) and rewrites the row IN PLACE as two bf16 planes
//                   hi = bf16(g), lo = bf16(g - hi): bytes [0, 2 NP) and [2 NP, 4 NP) of the
//                   row's 4 NP bytes.  No extra volume-sized buffer exists.
//   feature pack  : fmap1 -> bf16 hi/lo [b][d][p], fmap2 -> bf16 hi/lo [b][d][q'] with q' the
//                   padded patch-ordered target index (zeros on pads), both contraction-major.
//   two GEMMs     : persistent CTA pairs, tcgen05.mma.cta_group::2 (M = 256 rows per pair,
//                   N = D, fp32 accumulators in TMEM), 3-stage TMA ring, 3 MMAs per k-step in
//                   FC_MATH_TC_3XBF16 (hi*hi + lo*hi + hi*lo), split-K over work units with
//                   red.global.add.f32 into the zero-initialised outputs.
//       dF1: rows = queries p,  k = q' : the G planes are K-major operands.
//       dF2: rows = targets q', k = p  : the SAME G planes read as MN-major operands (the TMA
//            box is 64 q' x 64 p, the UMMA descriptor says "M-contiguous") -- no transposed copy.
//
// Warp roles as in fc_build_tc.cu: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer (leader CTA), warps 2-5 = epilogue.
#include "fc_umma.cuh"

namespace fc {

constexpr int BW_THREADS = 192;
constexpr int BW_BM = 128;                                  // rows per CTA (UMMA M = 256 per pair)
constexpr int BW_BK = 64;                                   // bf16 per 128-byte swizzle row
constexpr int BW_STAGES = 3;
constexpr int BW_PART_BYTES = 128 * BW_BK * 2;              // 16 KB: one operand part (<= 128 rows x 64 k)
constexpr int BW_STAGE_BYTES = 4 * BW_PART_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int BW_MAX_NP = 16384;                            // fold+pack stages a whole query row (all levels + both planes) in shared memory

enum { BW_DF1 = 0, BW_DF2 = 1 };

// ---------------------------------------------------------------- PTX
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// MN-major, SWIZZLE_128B operand (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units): 64 M-contiguous bf16 per 128-byte row, 8 k-rows per 1 KB atom, atoms along k
// SBO = 1 KB apart, the next 64 M elements LBO = 8 KB apart (a second 64 x 64 TMA box).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void lds128(uint32_t addr, float* v) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t* v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(p), "f"(v) : "memory");
}

// ---------------------------------------------------------------- fold + pack (in place)
// One CTA per query row.  The row's maps of every level arrive in shared memory as 1-D bulk
// copies (one instruction each, no registers in flight), the fold + split runs out of shared
// memory, and the two bf16 planes leave as ONE bulk store over the row's own level-0 bytes
// (hi plane = bytes [0, 2 NP), lo plane = [2 NP, 4 NP)): in place is safe because the store
// is issued after this CTA's loads have landed and no other CTA touches the row.
struct FoldParams {
    float* lvl[FC_MAX_LEVELS];          // level base pointers of the gradient pyramid
    int H[FC_MAX_LEVELS], W[FC_MAX_LEVELS], Wp[FC_MAX_LEVELS], msize[FC_MAX_LEVELS];
    int soff[FC_MAX_LEVELS];            // float offset of level l inside the shared staging area
    int L, NP, two_planes, in_floats;   // in_floats = sum of msize
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}

template <int L>
__global__ void __launch_bounds__(256) bwd_fold_pack_kernel(const FoldParams P) {
    extern __shared__ __align__(128) uint8_t fp_smem[];
    float* in = reinterpret_cast<float*>(fp_smem);                                   // [in_floats]
    uint4* out_hi = reinterpret_cast<uint4*>(fp_smem + (size_t)P.in_floats * 4);      // NP bf16
    uint4* out_lo = reinterpret_cast<uint4*>(fp_smem + (size_t)P.in_floats * 4 + (size_t)P.NP * 2);
    __shared__ uint64_t bar;
    const long long row = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)P.in_floats * 4u);
#pragma unroll
        for (int l = 0; l < L; ++l)
            bulk_load_1d(smem_u32(in + P.soff[l]), P.lvl[l] + row * P.msize[l], (uint32_t)P.msize[l] * 4u, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    const int n_patches = P.NP >> 4, ppr = P.Wp[0] >> 3;      // a patch = rows (y, y + 1) x 8 columns
    for (int pt = threadIdx.x; pt < n_patches; pt += 256) {
        const int y = 2 * (pt / ppr), x0 = 8 * (pt % ppr);
        // v: the (already folded) cells of the level above over this patch, coarsest level first
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int l = L - 1; l >= 1; --l) {
            const int n = (8 >> l) > 0 ? (8 >> l) : 1;        // cells of level l under 8 columns (rows y, y+1 share them)
            const int yl = y >> l, xl = x0 >> l;
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            if (yl < P.H[l] && xl < P.W[l]) {
                const float* src = in + P.soff[l] + tile_off(yl, xl, P.Wp[l]);
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                if (n == 4) { const float4 t = *reinterpret_cast<const float4*>(src); c[0] = t.x; c[1] = t.y; c[2] = t.z; c[3] = t.w; }
                else if (n == 2) { const float2 t = *reinterpret_cast<const float2*>(src); c[0] = t.x; c[1] = t.y; }
                else c[0] = *src;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < n && xl + j < P.W[l]) w[j] = c[j] + 0.25f * v[j >> 1];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = w[j];
        }
        const float4* g4 = reinterpret_cast<const float4*>(in + 16 * pt);
        float g[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float4 t = g4[i]; g[4 * i] = t.x; g[4 * i + 1] = t.y; g[4 * i + 2] = t.z; g[4 * i + 3] = t.w; }
        if (L > 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] += 0.25f * v[(j & 7) >> 1];
        }
        uint32_t h[8], lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(g[2 * j]), h1 = __float2bfloat16_rn(g[2 * j + 1]);
            const __nv_bfloat162 hh(h0, h1);
            const __nv_bfloat162 ll(__float2bfloat16_rn(g[2 * j] - __bfloat162float(h0)),
                                    __float2bfloat16_rn(g[2 * j + 1] - __bfloat162float(h1)));
            h[j] = *reinterpret_cast<const uint32_t*>(&hh);
            lw[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        out_hi[2 * pt] = make_uint4(h[0], h[1], h[2], h[3]);
        out_hi[2 * pt + 1] = make_uint4(h[4], h[5], h[6], h[7]);
        if (P.two_planes) {
            out_lo[2 * pt] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            out_lo[2 * pt + 1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store_1d(P.lvl[0] + row * P.NP, smem_u32(out_hi), (uint32_t)P.NP * (P.two_planes ? 4u : 2u));
        tma_commit_group();
        tma_wait_group_read<0>();                              // shared memory may be released
    }
}

// ---------------------------------------------------------------- feature pack
// blockIdx.z = 2 * b + which;  which 0: fmap1 -> [b][d][p] (row pitch N8), which 1: fmap2 ->
// [b][d][q'] (row pitch NPk, zeros on pad targets).  Each thread writes 8 consecutive elements.
struct BwdPackParams {
    const float* src[2];
    __nv_bfloat16* hi[2];
    __nv_bfloat16* lo[2];
    int D, N, N8, NP, NPk, H, W, Wp, two_planes;
};

__global__ void __launch_bounds__(256) bwd_pack_kernel(const BwdPackParams P) {
    const int which = blockIdx.z & 1, b = blockIdx.z >> 1, d = blockIdx.y;
    const int pitch = which ? P.NPk : P.N8;
    const int c = blockIdx.x * 256 + threadIdx.x;            // chunk of 8 output elements
    if (8 * c >= pitch) return;
    const float* s = P.src[which] + ((long long)b * P.D + d) * P.N;
    float x[8];
    if (which == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (8 * c + j < P.N) ? __ldg(s + 8 * c + j) : 0.f;
    } else {
        int y, x0;
        tile_inv(8 * c, P.Wp, y, x0);                        // a chunk is one patch row: same y, x0 .. x0 + 7
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (8 * c < P.NP && y < P.H && x0 + j < P.W) ? __ldg(s + y * P.W + x0 + j) : 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * j]), h1 = __float2bfloat16_rn(x[2 * j + 1]);
        const __nv_bfloat162 hh(h0, h1);
        const __nv_bfloat162 ll(__float2bfloat16_rn(x[2 * j] - __bfloat162float(h0)),
                                __float2bfloat16_rn(x[2 * j + 1] - __bfloat162float(h1)));
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const long long o = ((long long)b * P.D + d) * pitch + 8 * c;
    *reinterpret_cast<uint4*>(P.hi[which] + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (P.two_planes) *reinterpret_cast<uint4*>(P.lo[which] + o) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------- GEMM
struct BwdParams {
    float* out;            // dfmap1 or dfmap2: (B, D, N) fp32, zero-initialised (split-K accumulates)
    int D, N, NP, H, W, Wp;
    int M;                 // rows of this GEMM: N (dF1) or NP (dF2)
    int kb_total;          // ceil(K / 64), K = NP (dF1) or N (dF2)
    int mp;                // pair-tiles (256 rows) per sample
    int ksplit;            // work units per (sample, pair-tile)
    int units;             // B * mp * ksplit
    int three_pass;
    float scale;           // 1 / sqrt(D)
};

template <int OP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BW_THREADS, 1)
tc_bwd_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
              const BwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + BW_STAGES * BW_STAGE_BYTES);
    uint64_t* full = bars;                          // BW_STAGES (leader's are used)
    uint64_t* empty = full + BW_STAGES;             // BW_STAGES
    uint64_t* t_full = empty + BW_STAGES;           // 2
    uint64_t* t_empty = t_full + 2;                 // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_parts = P.three_pass ? 2 : 1;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int half_n = P.D / 2;                     // rows of the N-side operand held by each CTA

    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int u_begin = (int)((long long)P.units * cluster_id / n_clusters);
    const int u_end = (int)((long long)P.units * (cluster_id + 1) / n_clusters);
    auto decode = [&](int u, int& b, int& m0, int& kb0, int& kb1) {
        const int am = u / P.ksplit, ks = u - am * P.ksplit;
        b = am / P.mp;
        m0 = ((am - b * P.mp) * 2 + (int)rank) * BW_BM;
        kb0 = (int)((long long)P.kb_total * ks / P.ksplit);
        kb1 = (int)((long long)P.kb_total * (ks + 1) / P.ksplit);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < BW_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }   // 4 epilogue warps x 2 CTAs
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();          // CTA-wide: orders tcgen05.alloc's write of the TMEM address before every read of it
    cluster_sync_all();       // cluster-wide: the peer's barriers are initialised before anyone arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            const uint32_t stage_tx = (uint32_t)(2 * n_parts * (BW_PART_BYTES + half_n * BW_BK * 2));
            int it = 0;
            for (int u = u_begin; u < u_end; ++u) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % BW_STAGES;
                    mbar_wait(empty + s, ((uint32_t)(it / BW_STAGES) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(full + s, stage_tx);
                    uint8_t* st = ring + s * BW_STAGE_BYTES;
                    for (int part = 0; part < n_parts; ++part) {
                        const CUtensorMap* ma = part ? &map_a_lo : &map_a_hi;
                        const CUtensorMap* mb = part ? &map_b_lo : &map_b_hi;
                        uint8_t* a_dst = st + part * BW_PART_BYTES;
                        if (OP == BW_DF1) {
                            tma2_load_3d(a_dst, ma, full + s, kb * BW_BK, m0, b);               // [128 p][64 q']
                        } else {
                            tma2_load_3d(a_dst, ma, full + s, m0, kb * BW_BK, b);               // [64 p][64 q'] x 2
                            tma2_load_3d(a_dst + BW_PART_BYTES / 2, ma, full + s, m0 + 64, kb * BW_BK, b);
                        }
                        tma2_load_2d(st + (2 + part) * BW_PART_BYTES, mb, full + s, kb * BW_BK,
                                     b * P.D + (int)rank * half_n);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(2 * BW_BM, P.D) | (OP == BW_DF2 ? (1u << 15) : 0u);
            int it = 0, uc = 0;
            for (int u = u_begin; u < u_end; ++u, ++uc) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                const int buf = uc & 1;
                mbar_wait(t_empty + buf, ((uint32_t)(uc >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % BW_STAGES;
                    mbar_wait(full + s, (uint32_t)(it / BW_STAGES) & 1u);
                    tc_fence_after();
                    const uint32_t st = smem_u32(ring + s * BW_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BW_BK / 16; ++k) {
                        uint64_t ah, al;
                        if (OP == BW_DF1) {
                            ah = umma_desc_sw128(st + k * 32);
                            al = umma_desc_sw128(st + BW_PART_BYTES + k * 32);
                        } else {
                            ah = umma_desc_mn_sw128(st + k * 2048);
                            al = umma_desc_mn_sw128(st + BW_PART_BYTES + k * 2048);
                        }
                        const uint64_t bh = umma_desc_sw128(st + 2 * BW_PART_BYTES + k * 32);
                        const uint64_t bl = umma_desc_sw128(st + 3 * BW_PART_BYTES + k * 32);
                        umma2_bf16(d_addr, ah, bh, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        if (P.three_pass) {
                            umma2_bf16(d_addr, al, bh, idesc, 1u);
                            umma2_bf16(d_addr, ah, bl, idesc, 1u);
                        }
                    }
                    umma2_commit(empty + s);
                }
                umma2_commit(t_full + buf);
            }
        }
    } else {
        // ================= epilogue: TMEM -> scale -> red.add into (B, D, N) =================
        const int quarter = warp & 3;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        int uc = 0;
        for (int u = u_begin; u < u_end; ++u, ++uc) {
            int b, m0, kb0, kb1;
            decode(u, b, m0, kb0, kb1);
            const int buf = uc & 1;
            const int r = m0 + quarter * 32 + lane;            // row of this GEMM held by this thread
            int col = -1;                                      // position inside the (.., N) output row
            if (r < P.M) {
                if (OP == BW_DF1) {
                    col = r;
                } else {
                    int y, x;
                    tile_inv(r, P.Wp, y, x);
                    if (y < P.H && x < P.W) col = y * P.W + x;
                }
            }
            float* dst = P.out + (long long)b * P.D * P.N + col;
            mbar_wait(t_full + buf, (uint32_t)(uc >> 1) & 1u);
            tc_fence_after();
            for (int c0 = 0; c0 < P.D; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)(buf * 256 + c0), v);
                tmem_ld_wait();
                if (col >= 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) red_add_f32(dst + (long long)(c0 + j) * P.N, v[j] * P.scale);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(t_empty + buf, 0);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------- GEMM straight from the fp32 gradient pyramid
// The default backward.  Same CTA-pair GEMMs as tc_bwd_kernel, but the G operand comes from the fp32 gradient pyramid
// itself: the fold (avg_pool2d backward of corr.py:24-27), the bf16 hi/lo split and the swizzled operand layout happen in
// shared memory on the way to the tensor core, so the fold + pack pass over the volume (a read AND a write of 4 N NP
// bytes per sample) and the in-place planes do not exist, the gradient pyramid is left untouched and there is no
// "whole query row in shared memory" size limit.
//
//   warp 0      TMA: per k-block one fp32 box of G level 0 into this CTA's `raw` buffer ([128 p][64 q'] for dF1,
//               [64 p][128 q'] for dF2; rows past the tensor arrive as zeros) + the feature operand's bf16 hi/lo boxes
//   warps 6-13  converters: thread = 4 consecutive targets q' of one row: += the folded coarse levels (read through the
//               read-only path: 32-byte pieces of the level-1..L-1 maps of the same query), split into bf16 hi/lo and
//               store where a SWIZZLE_128B TMA box would have put them (row r, 16-byte chunk j at j ^ (r & 7)); the
//               same bytes serve as K-major operand (dF1: rows = p) and MN-major operand (dF2: rows = k = p)
//   warp 1      leader: waits for both CTAs' converted halves + the feature boxes, issues the MMAs; peer: relays its
//               CTA's "converted" barrier to the leader with ONE cluster-scope release per k-block
//   warps 2-5   epilogue as in tc_bwd_kernel
constexpr int BF_THREADS = 448;
constexpr int BF_CONV0 = 6, BF_CONV_WARPS = 8;                // warps 2-5: epilogue, 6-13: converters
constexpr int BF_STAGES = 2;
constexpr int BF_RAW_BYTES = 128 * BW_BK * 4;                // 32 KB of fp32
constexpr int BF_STAGE_BYTES = BF_RAW_BYTES + 4 * BW_PART_BYTES;   // raw | A_hi | A_lo | B_hi | B_lo = 96 KB

// Timeline probe (FC_PROBES builds only; tools/probe_bwd_trace.py): CTA 0 stamps clock64() at the hand-offs of its first
// 256 k-blocks.  producer: 0 box buffer free, 1 operand stage free; converter thread 0: 2 ready for the k-block, 3 box
// landed, 4 operand stage free, 5 operand stored; MMA thread: 6 own half converted, 7 peer's half, 8 feature boxes,
// 9 MMAs issued + committed
#ifdef FC_PROBES
constexpr int BF_TRACE_SLOTS = 10;
__device__ unsigned long long fc_bwd_trace_buf[2][256 * BF_TRACE_SLOTS];
#define BF_TRACE(it, k) do { if (blockIdx.x == 0 && (it) < 256) fc_bwd_trace_buf[OP][(it) * BF_TRACE_SLOTS + (k)] = clock64(); } while (0)
#else
#define BF_TRACE(it, k) do {} while (0)
#endif

struct FoldSrc {
    const float* lvl[FC_MAX_LEVELS];
    int H[FC_MAX_LEVELS], W[FC_MAX_LEVELS], Wp[FC_MAX_LEVELS], msize[FC_MAX_LEVELS];
    int L;
};

template <int OP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BF_THREADS, 1)
tc_bwd_fold_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_b_hi,
                   const __grid_constant__ CUtensorMap map_b_lo, const BwdParams P, const FoldSrc F) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + BF_STAGES * BF_STAGE_BYTES);
    uint64_t* raw_full = bars;                      // per stage: the fp32 box landed (this CTA)
    uint64_t* raw_empty = raw_full + BF_STAGES;     // the converters have read it (8 warps)
    uint64_t* b_full = raw_empty + BF_STAGES;       // leader's: feature boxes of BOTH CTAs landed
    uint64_t* a_part = b_full + BF_STAGES;          // this CTA's converted operand is written (8 warps)
    uint64_t* a_peer = a_part + BF_STAGES;          // leader's: the peer's relay
    uint64_t* empty = a_peer + BF_STAGES;           // multicast commit: the MMAs reading this stage retired
    uint64_t* t_full = empty + BF_STAGES;           // 2
    uint64_t* t_empty = t_full + 2;                 // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_parts = P.three_pass ? 2 : 1;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int half_n = P.D / 2;

    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    // units are dealt round-robin: at any time the CTA pairs work on neighbouring row tiles and all k-splits of ONE
    // sample, so its feature planes (7 MB) are shared out of L2 and the k-splits of a tile reduce into L2-resident lines
    // (DRAM reads per GEMM at B = 6, 54x128: 2.07 / 2.59 GB with contiguous ranges -> 1.63 / 1.64 GB; the pyramid is 1.58)
    const int u_begin = cluster_id, u_end = P.units, u_step = n_clusters;
    auto decode = [&](int u, int& b, int& m0, int& kb0, int& kb1) {
        const int am = u / P.ksplit, ks = u - am * P.ksplit;
        b = am / P.mp;
        m0 = ((am - b * P.mp) * 2 + (int)rank) * BW_BM;
        kb0 = (int)((long long)P.kb_total * ks / P.ksplit);
        kb1 = (int)((long long)P.kb_total * (ks + 1) / P.ksplit);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < BF_STAGES; ++i) {
            mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, BF_CONV_WARPS); mbar_init(b_full + i, 1);
            mbar_init(a_part + i, BF_CONV_WARPS); mbar_init(a_peer + i, 1); mbar_init(empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            const uint32_t b_tx = (uint32_t)(2 * n_parts * half_n * BW_BK * 2);
            // the gradient volume streams through L2 once; the feature planes are swept again by every row tile
            const uint64_t stream = l2_policy_evict_first(), keep = l2_policy_evict_last();
            int it = 0;
            for (int u = u_begin; u < u_end; u += u_step) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % BF_STAGES;
                    const uint32_t free_parity = ((uint32_t)(it / BF_STAGES) & 1u) ^ 1u;
                    uint8_t* st = ring + s * BF_STAGE_BYTES;
                    mbar_wait(raw_empty + s, free_parity);
                    BF_TRACE(it, 0);
                    mbar_expect_tx(raw_full + s, (uint32_t)BF_RAW_BYTES);
                    if (OP == BW_DF1) tma_load_3d_hint(smem_u32(st), &map_g, smem_u32(raw_full + s), kb * BW_BK, m0, b, stream);   // [128 p][64 q']
                    else tma_load_3d_hint(smem_u32(st), &map_g, smem_u32(raw_full + s), m0, kb * BW_BK, b, stream);                // [64 p][128 q']
                    mbar_wait(empty + s, free_parity);
                    BF_TRACE(it, 1);
                    if (leader) mbar_expect_tx(b_full + s, b_tx);
                    for (int part = 0; part < n_parts; ++part)
                        tma2_load_2d_hint(st + BF_RAW_BYTES + (2 + part) * BW_PART_BYTES, part ? &map_b_lo : &map_b_hi, b_full + s,
                                          kb * BW_BK, b * P.D + (int)rank * half_n, keep);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader) / relay (peer) =================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(2 * BW_BM, P.D) | (OP == BW_DF2 ? (1u << 15) : 0u);
            int it = 0, uc = 0;
            for (int u = u_begin; u < u_end; u += u_step, ++uc) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                const int buf = uc & 1;
                if (leader) {
                    mbar_wait_cluster(t_empty + buf, ((uint32_t)(uc >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
                const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % BF_STAGES;
                    const uint32_t parity = (uint32_t)(it / BF_STAGES) & 1u;
                    mbar_wait(a_part + s, parity);                       // this CTA's half of the G operand
                    BF_TRACE(it, 6);
                    if (!leader) { mbar_arrive_remote_default(a_peer + s, 0); continue; }
                    mbar_wait(a_peer + s, parity);                       // the peer's half
                    BF_TRACE(it, 7);
                    mbar_wait(b_full + s, parity);
                    BF_TRACE(it, 8);
                    tc_fence_after();
                    const uint32_t st = smem_u32(ring + s * BF_STAGE_BYTES) + BF_RAW_BYTES;
#pragma unroll
                    for (int k = 0; k < BW_BK / 16; ++k) {
                        uint64_t ah, al;
                        if (OP == BW_DF1) {
                            ah = umma_desc_sw128(st + k * 32);
                            al = umma_desc_sw128(st + BW_PART_BYTES + k * 32);
                        } else {
                            ah = umma_desc_mn_sw128(st + k * 2048);
                            al = umma_desc_mn_sw128(st + BW_PART_BYTES + k * 2048);
                        }
                        const uint64_t bh = umma_desc_sw128(st + 2 * BW_PART_BYTES + k * 32);
                        const uint64_t bl = umma_desc_sw128(st + 3 * BW_PART_BYTES + k * 32);
                        umma2_bf16(d_addr, ah, bh, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        if (P.three_pass) {
                            umma2_bf16(d_addr, al, bh, idesc, 1u);
                            umma2_bf16(d_addr, ah, bl, idesc, 1u);
                        }
                    }
                    umma2_commit(empty + s);
                    BF_TRACE(it, 9);
                }
                if (leader) umma2_commit(t_full + buf);
            }
        }
    } else if (warp < BF_CONV0) {
        // ================= epilogue: TMEM -> scale -> red.add into (B, D, N) =================
        const int quarter = warp & 3;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        int uc = 0;
        for (int u = u_begin; u < u_end; u += u_step, ++uc) {
            int b, m0, kb0, kb1;
            decode(u, b, m0, kb0, kb1);
            const int buf = uc & 1;
            const int r = m0 + quarter * 32 + lane;
            int col = -1;
            if (r < P.M) {
                if (OP == BW_DF1) {
                    col = r;
                } else {
                    int y, x;
                    tile_inv(r, P.Wp, y, x);
                    if (y < P.H && x < P.W) col = y * P.W + x;
                }
            }
            float* dst = P.out + (long long)b * P.D * P.N + col;
            mbar_wait(t_full + buf, (uint32_t)(uc >> 1) & 1u);
            tc_fence_after();
            for (int c0 = 0; c0 < P.D; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)(buf * 256 + c0), v);
                tmem_ld_wait();
                if (col >= 0) {
                    float* d = dst + (long long)c0 * P.N;
#pragma unroll
                    for (int j = 0; j < 32; ++j, d += P.N) red_add_f32(d, v[j] * P.scale);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(t_empty + buf, 0);
        }
    } else {
        // ================= converters: fp32 box (+ folded coarse levels) -> bf16 hi/lo swizzled operand =================
        // thread = one patch row (8 consecutive targets of one map row) of one query; 4 sweeps cover the box.  The coarse
        // cells over a patch row are 4 + 2 + 1 (+ 1 + 1) values: they are fetched ONE K-BLOCK AHEAD into registers (global
        // latency would otherwise sit between every two k-blocks of a warp).
        constexpr int C = (OP == BW_DF1) ? 64 : 128;        // targets per row of the box
        constexpr int R = 128 * BW_BK / C;                  // rows (queries) of the box
        constexpr int CPR = C / 8;                          // patch rows per box row
        constexpr int RPS = 256 / CPR;                      // box rows per sweep of the 256 converter threads
        constexpr int SWEEPS = R / RPS;                     // 4
        const int ctid = threadIdx.x - BF_CONV0 * 32;
        const int c = (ctid % CPR) * 8, r0 = ctid / CPR;
        const uint32_t ring_s = smem_u32(ring);
        const uint32_t raw_off = (uint32_t)(r0 * C + c) * 4u;
        const uint32_t a_off = (uint32_t)(BF_RAW_BYTES + (c >> 6) * (R * 128) + r0 * 128 + (((((c & 63) >> 3)) ^ (r0 & 7)) << 4));
        const int L = F.L;
        // element strides between this thread's rows of consecutive sweeps, per coarse level
        const int st1 = L > 1 ? RPS * F.msize[1] : 0, st2 = L > 2 ? RPS * F.msize[2] : 0, st3 = L > 3 ? RPS * F.msize[3] : 0;

        struct Pos { int u, kb, kb1, b, m0; };
        auto first = [&](Pos& o, int u) {
            o.u = u;
            if (u < u_end) { int kb0; decode(u, o.b, o.m0, kb0, o.kb1); o.kb = kb0; }
        };
        struct Coarse {                 // the coarse cells over this thread's patch row, per sweep (= per query row)
            float4 c1[SWEEPS];          // level 1: 4 cells
            float2 c2[SWEEPS];          // level 2: 2 cells
            float c3[SWEEPS];           // level 3: 1 cell
        };
        // validity of the coarse cells over the patch row at target q: bits 0-3 level-1 cells, 4-5 level-2 cells, 6.. levels
        // 3, 4, 5; element offsets of the first cell inside the query's level-l map
        uint32_t ok_n = 0;
        int off1 = 0, off2 = 0, off3 = 0, off4 = 0, off5 = 0, q_taps = -1;
        // the patch (row pair py, patch px of the row) this thread's 8 targets lie in.  dF1 walks the targets 4 patches per
        // k-block: the position is carried along (one division per work unit instead of one per k-block)
        const int ppr = P.Wp >> 3;
        int py = 0, px = 0;
        auto taps = [&](int q) {
            if (L > 1 && q < P.NP) {
                if (OP == BW_DF1 && q == q_taps + BW_BK) {
                    px += BW_BK / 16;
                    while (px >= ppr) { px -= ppr; ++py; }
                } else {
                    const int pt = q >> 4;
                    py = pt / ppr; px = pt - py * ppr;
                }
            }
            q_taps = q; ok_n = 0;
            if (L <= 1 || q >= P.NP) return;
            const int y = 2 * py + ((q >> 3) & 1), x = 8 * px;
#pragma unroll
            for (int l = 1; l < FC_MAX_LEVELS; ++l) {
                if (l < L) {
                    const int yl = y >> l, xl = x >> l, n = l == 1 ? 4 : (l == 2 ? 2 : 1);
                    const int o_l = tile_off(yl, xl, F.Wp[l]);
                    const int bit0 = l == 1 ? 0 : (l == 2 ? 4 : 3 + l);
                    if (yl < F.H[l]) {
#pragma unroll
                        for (int j = 0; j < n; ++j)
                            if (xl + j < F.W[l]) ok_n |= 1u << (bit0 + j);
                    }
                    if (l == 1) off1 = o_l; else if (l == 2) off2 = o_l; else if (l == 3) off3 = o_l;
                    else if (l == 4) off4 = o_l; else off5 = o_l;
                }
            }
        };
        auto fetch = [&](const Pos& o, Coarse& f) {
            const int q = (OP == BW_DF1 ? o.kb * BW_BK : o.m0) + c;
            if (q != q_taps) taps(q);
            const int p0 = (OP == BW_DF1 ? o.m0 : o.kb * BW_BK) + r0;
            const long long row = (long long)o.b * P.N + p0;
            const float* g1 = (ok_n & 0x1u) ? F.lvl[1] + row * F.msize[1] + off1 : nullptr;
            const float* g2 = (ok_n & 0x10u) ? F.lvl[2] + row * F.msize[2] + off2 : nullptr;
            const float* g3 = (ok_n & 0x40u) ? F.lvl[3] + row * F.msize[3] + off3 : nullptr;
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                const bool pv = p0 + i * RPS < P.N;
                f.c1[i] = (pv && g1) ? __ldg(reinterpret_cast<const float4*>(g1 + i * st1)) : make_float4(0.f, 0.f, 0.f, 0.f);
                f.c2[i] = (pv && g2) ? __ldg(reinterpret_cast<const float2*>(g2 + i * st2)) : make_float2(0.f, 0.f);
                f.c3[i] = (pv && g3) ? __ldg(g3 + i * st3) : 0.f;
            }
        };

        Pos cur;
        first(cur, u_begin);
        Coarse nxt_f;
        if (cur.u < u_end) fetch(cur, nxt_f);
        for (int it = 0; cur.u < u_end; ++it) {
            const int s = it % BF_STAGES;
            const uint32_t parity = (uint32_t)(it / BF_STAGES) & 1u;
            // fold the prefetched coarse cells into one addend per pair of targets (frees their registers for the next fetch)
            float w1[SWEEPS][4];
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                float v = 0.f;                               // levels >= 3: one cell over the patch row, coarsest first
                if (L > 4) {
                    const int p = (OP == BW_DF1 ? cur.m0 : cur.kb * BW_BK) + r0 + i * RPS;
                    if (p < P.N) {
                        const long long row = (long long)cur.b * P.N + p;
                        if (L > 5 && (ok_n >> 8 & 1u)) v = __ldg(F.lvl[5] + row * F.msize[5] + off5);
                        v = (ok_n >> 7 & 1u) ? __ldg(F.lvl[4] + row * F.msize[4] + off4) + 0.25f * v : 0.f;
                    }
                }
                v = (ok_n >> 6 & 1u) ? nxt_f.c3[i] + 0.25f * v : 0.f;
                const float w2a = (ok_n >> 4 & 1u) ? nxt_f.c2[i].x + 0.25f * v : 0.f;
                const float w2b = (ok_n >> 5 & 1u) ? nxt_f.c2[i].y + 0.25f * v : 0.f;
                w1[i][0] = (ok_n & 1u) ? 0.25f * (nxt_f.c1[i].x + 0.25f * w2a) : 0.f;
                w1[i][1] = (ok_n & 2u) ? 0.25f * (nxt_f.c1[i].y + 0.25f * w2a) : 0.f;
                w1[i][2] = (ok_n & 4u) ? 0.25f * (nxt_f.c1[i].z + 0.25f * w2b) : 0.f;
                w1[i][3] = (ok_n & 8u) ? 0.25f * (nxt_f.c1[i].w + 0.25f * w2b) : 0.f;
            }
            Pos nx = cur;
            if (++nx.kb >= nx.kb1) first(nx, nx.u + u_step);
            if (nx.u < u_end) fetch(nx, nxt_f);             // in flight while this k-block is converted
            const uint32_t st = ring_s + (uint32_t)(s * BF_STAGE_BYTES);
            if (ctid == 0) BF_TRACE(it, 2);
            mbar_wait(raw_full + s, parity);
            if (ctid == 0) BF_TRACE(it, 3);
            mbar_wait(empty + s, parity ^ 1u);              // the MMAs that read this stage's operand two k-blocks ago retired
            if (ctid == 0) BF_TRACE(it, 4);
            float xs[SWEEPS][8];                             // all shared loads first: the sweeps below then overlap
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                lds128(st + raw_off + i * (RPS * C * 4), xs[i]);
                lds128(st + raw_off + i * (RPS * C * 4) + 16, xs[i] + 4);
            }
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                float* x = xs[i];
                if (L > 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] += w1[i][j >> 1];
                }
                uint32_t h[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
                    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * j] - __uint_as_float(h[j] << 16),
                                                                    x[2 * j + 1] - __uint_as_float(h[j] & 0xffff0000u));
                    lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                sts128(st + a_off + i * (RPS * 128), h);
                if (P.three_pass) sts128(st + a_off + BW_PART_BYTES + i * (RPS * 128), lo);
            }
            fence_proxy_async_smem();                       // generic-proxy writes -> the tensor core's reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(raw_empty + s); mbar_arrive(a_part + s); }
            if (ctid == 0) BF_TRACE(it, 5);
            cur = nx;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------- the same GEMMs for patch-row aligned maps
// When a row of the map holds a multiple of 4 patches (Wp % 32 == 0: 128, 96, 64, 160 ... -- every size the reference's
// configurations use) a k-block of dF1 (64 targets = 4 patches) and each half of a dF2 row tile (128 targets = 2 x 4
// patches) lies inside ONE patch row, so the coarse cells over it are three small boxes per level-1..3 map --
// [rows][2 patches x 8], [rows][8], [rows][4] cells -- that the TMA unit can fetch: no index arithmetic, no global loads
// and no validity logic in the converter threads (pad cells of the gradient pyramid are zeros and an invalid cell's
// parents are pads too, so `c1 + (c2 + c3/4)/4` needs no masks), ~270 instead of ~600 dependent instructions per thread
// and k-block.  What the timeline of the generic kernel asked for (tools/probe_bwd_trace.py) is built in:
//   * the fp32 boxes have their own 2-deep ring, filled by their own producer warp as soon as the converters hold the
//     previous box in registers -- their DRAM latency (~2200 cycles under load) is outside the loop MMAs retire ->
//     operand stage free -> converted -> MMAs;
//   * a converter thread reads its part of the box, folds and splits it BEFORE it waits for the operand stage, so only
//     the 8 shared stores sit between "the MMAs two k-blocks back retired" and "this k-block's operand is ready";
//   * the peer CTA's relay arrives on the leader's barrier with the instruction's default (CTA-scope release) semantics:
//     a release at cluster scope cost a MEMBAR of ~700 cycles per k-block (see mbar_arrive_remote_default).
// Measured (B = 6, 54x128 / 46x96): 0.82 / 0.36 ms against 0.97 / 0.43 for the generic kernel and 1.20 / 0.50 for the
// round-1 pipeline; ~1950 cycles per k-block, the 12 MMAs of a k-block retire in ~1850.
constexpr int BA_STAGES = 2;                                             // operand stages: A_hi | A_lo | B_hi | B_lo = 64 KB
constexpr int BA_STAGE_BYTES = 4 * BW_PART_BYTES;
constexpr int BA_RSTAGES = 2;                                            // fp32 boxes in flight (32 KB each)
constexpr int BA_RAW_BYTES = 128 * BW_BK * 4;
constexpr int BA_CSTAGES = 2;
constexpr int BA_THREADS = BF_THREADS + 32;                              // + warp 14: the box / coarse-box producer
constexpr int BA_C1 = 0, BA_C2 = 8192, BA_C3 = 12288, BA_CBYTES = 14336; // coarse boxes of one k-block: level 1 | 2 | 3

__device__ __forceinline__ void tma_load_5d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
            smem_dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void lds128f(uint32_t addr, float* v) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

struct CoarseMaps { CUtensorMap m[3]; };            // levels 1..3 as {8 cells, 2 sub-rows, patches, N, B}
struct CoarseGeo { int ppr[4]; int L; };            // patches per row of levels 0..3; number of levels (<= 4)

template <int OP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BA_THREADS, 1)
tc_bwd_fold_aligned_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_b_hi,
                           const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CoarseMaps CM,
                           const BwdParams P, const CoarseGeo G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint8_t* rring = ring + BA_STAGES * BA_STAGE_BYTES;
    uint8_t* cring = rring + BA_RSTAGES * BA_RAW_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(cring + BA_CSTAGES * BA_CBYTES);
    uint64_t* raw_full = bars;                      // per box buffer: the fp32 box landed (this CTA)
    uint64_t* raw_empty = raw_full + BA_RSTAGES;    // ... and is in the converters' registers (8 warps)
    uint64_t* b_full = raw_empty + BA_RSTAGES;      // leader's: feature boxes of BOTH CTAs landed
    uint64_t* a_part = b_full + BA_STAGES;          // this CTA's converted operand is written (8 warps)
    uint64_t* a_peer = a_part + BA_STAGES;          // leader's: the peer's relay
    uint64_t* empty = a_peer + BA_STAGES;           // multicast commit: the MMAs reading this stage retired
    uint64_t* c_full = empty + BA_STAGES;           // coarse boxes landed
    uint64_t* c_empty = c_full + BA_CSTAGES;        // ... and read (8 warps)
    uint64_t* t_full = c_empty + BA_CSTAGES;        // 2
    uint64_t* t_empty = t_full + 2;                 // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_parts = P.three_pass ? 2 : 1;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int half_n = P.D / 2;

    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int u_begin = cluster_id, u_end = P.units, u_step = n_clusters;      // round robin (see tc_bwd_fold_kernel)
    auto decode = [&](int u, int& b, int& m0, int& kb0, int& kb1) {
        const int am = u / P.ksplit, ks = u - am * P.ksplit;
        b = am / P.mp;
        m0 = ((am - b * P.mp) * 2 + (int)rank) * BW_BM;
        kb0 = (int)((long long)P.kb_total * ks / P.ksplit);
        kb1 = (int)((long long)P.kb_total * (ks + 1) / P.ksplit);
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < BA_STAGES; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(a_part + i, BF_CONV_WARPS); mbar_init(a_peer + i, 1); mbar_init(empty + i, 1);
        }
        for (int i = 0; i < BA_RSTAGES; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_empty + i, BF_CONV_WARPS); }
        for (int i = 0; i < BA_CSTAGES; ++i) { mbar_init(c_full + i, 1); mbar_init(c_empty + i, BF_CONV_WARPS); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int ROWS = OP == BW_DF1 ? 128 : 64;   // query rows of a k-block's box
    const int nlev = G.L;                            // 1 .. 4
    const uint32_t c_tx = (uint32_t)(ROWS * ((nlev > 1 ? 64 : 0) + (nlev > 2 ? 32 : 0) + (nlev > 3 ? 16 : 0)) * (OP == BW_DF1 ? 1 : 2));

    if (warp == 0) {
        // ================= TMA producer of the feature operand (both CTAs) =================
        if (elect_one()) {
            const uint32_t b_tx = (uint32_t)(2 * n_parts * half_n * BW_BK * 2);
            const uint64_t keep = l2_policy_evict_last();
            int it = 0;
            for (int u = u_begin; u < u_end; u += u_step) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % BA_STAGES;
                    uint8_t* st = ring + s * BA_STAGE_BYTES;
                    mbar_wait(empty + s, ((uint32_t)(it / BA_STAGES) & 1u) ^ 1u);   // the MMAs that read this stage retired (both CTAs)
                    BF_TRACE(it, 1);
                    if (leader) mbar_expect_tx(b_full + s, b_tx);
                    for (int part = 0; part < n_parts; ++part)
                        tma2_load_2d_hint(st + (2 + part) * BW_PART_BYTES, part ? &map_b_lo : &map_b_hi, b_full + s,
                                          kb * BW_BK, b * P.D + (int)rank * half_n, keep);
                }
            }
        }
    } else if (warp == BF_CONV0 + BF_CONV_WARPS) {
        // ================= TMA producer of the fp32 boxes and the coarse boxes (both CTAs) =================
        // its own warp: these loads wait for the CONVERTERS (box buffer read), not for the MMAs, and run ahead of them
        if (elect_one()) {
            const uint64_t stream = l2_policy_evict_first();
            // the coarse boxes over 4 patches starting at patch px0 of patch row py (levels 1..3), `rows` queries from p0
            auto coarse = [&](uint32_t dst, uint32_t bar, int py, int px0, int p0, int b, int rows_off) {
                if (nlev > 1) tma_load_5d(dst + BA_C1 + rows_off * 64, &CM.m[0], bar, 0, py & 1, (py >> 1) * G.ppr[1] + (px0 >> 1), p0, b);
                if (nlev > 2) tma_load_5d(dst + BA_C2 + rows_off * 32, &CM.m[1], bar, 0, (py >> 1) & 1, (py >> 2) * G.ppr[2] + (px0 >> 2), p0, b);
                if (nlev > 3) tma_load_5d(dst + BA_C3 + rows_off * 16, &CM.m[2], bar, px0 & 4, (py >> 2) & 1, (py >> 3) * G.ppr[3] + (px0 >> 3), p0, b);
            };
            int it = 0;
            for (int u = u_begin; u < u_end; u += u_step) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    if (nlev > 1) {
                        const int cs = it % BA_CSTAGES;
                        mbar_wait(c_empty + cs, ((uint32_t)(it / BA_CSTAGES) & 1u) ^ 1u);
                        mbar_expect_tx(c_full + cs, c_tx);
                        const uint32_t dst = smem_u32(cring + cs * BA_CBYTES), bar = smem_u32(c_full + cs);
                        if (OP == BW_DF1) {
                            const int pt = kb * (BW_BK / 16);                 // first patch of the k-block
                            const int py = pt / G.ppr[0];
                            coarse(dst, bar, py, pt - py * G.ppr[0], m0, b, 0);
                        } else {
#pragma unroll
                            for (int g = 0; g < 2; ++g) {                     // the two 4-patch halves of this CTA's row tile
                                const int pt = (m0 >> 4) + 4 * g;
                                const int py = pt / G.ppr[0];
                                coarse(dst, bar, py, pt - py * G.ppr[0], kb * BW_BK, b, g * ROWS);
                            }
                        }
                    }
                    const int rs = it % BA_RSTAGES;
                    mbar_wait(raw_empty + rs, ((uint32_t)(it / BA_RSTAGES) & 1u) ^ 1u);
                    BF_TRACE(it, 0);
                    mbar_expect_tx(raw_full + rs, (uint32_t)BA_RAW_BYTES);
                    const uint32_t dst = smem_u32(rring + rs * BA_RAW_BYTES);
                    if (OP == BW_DF1) tma_load_3d_hint(dst, &map_g, smem_u32(raw_full + rs), kb * BW_BK, m0, b, stream);   // [128 p][64 q']
                    else tma_load_3d_hint(dst, &map_g, smem_u32(raw_full + rs), m0, kb * BW_BK, b, stream);                // [64 p][128 q']
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader) / relay (peer) =================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(2 * BW_BM, P.D) | (OP == BW_DF2 ? (1u << 15) : 0u);
            int it = 0, uc = 0;
            for (int u = u_begin; u < u_end; u += u_step, ++uc) {
                int b, m0, kb0, kb1;
                decode(u, b, m0, kb0, kb1);
                const int buf = uc & 1;
                if (leader) {
                    mbar_wait_cluster(t_empty + buf, ((uint32_t)(uc >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                }
                const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % BA_STAGES;
                    const uint32_t parity = (uint32_t)(it / BA_STAGES) & 1u;
                    mbar_wait(a_part + s, parity);                       // this CTA's half of the G operand
                    BF_TRACE(it, 6);
                    if (!leader) { mbar_arrive_remote_default(a_peer + s, 0); continue; }
                    mbar_wait(a_peer + s, parity);                       // the peer's half
                    BF_TRACE(it, 7);
                    mbar_wait(b_full + s, parity);
                    BF_TRACE(it, 8);
                    tc_fence_after();
                    const uint32_t st = smem_u32(ring + s * BA_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BW_BK / 16; ++k) {
                        uint64_t ah, al;
                        if (OP == BW_DF1) {
                            ah = umma_desc_sw128(st + k * 32);
                            al = umma_desc_sw128(st + BW_PART_BYTES + k * 32);
                        } else {
                            ah = umma_desc_mn_sw128(st + k * 2048);
                            al = umma_desc_mn_sw128(st + BW_PART_BYTES + k * 2048);
                        }
                        const uint64_t bh = umma_desc_sw128(st + 2 * BW_PART_BYTES + k * 32);
                        const uint64_t bl = umma_desc_sw128(st + 3 * BW_PART_BYTES + k * 32);
                        umma2_bf16(d_addr, ah, bh, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        if (P.three_pass) {
                            umma2_bf16(d_addr, al, bh, idesc, 1u);
                            umma2_bf16(d_addr, ah, bl, idesc, 1u);
                        }
                    }
                    umma2_commit(empty + s);
                    BF_TRACE(it, 9);
                }
                if (leader) umma2_commit(t_full + buf);
            }
        }
    } else if (warp < BF_CONV0) {
        // ================= epilogue: TMEM -> scale -> red.add into (B, D, N) =================
        const int quarter = warp & 3;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        int uc = 0;
        for (int u = u_begin; u < u_end; u += u_step, ++uc) {
            int b, m0, kb0, kb1;
            decode(u, b, m0, kb0, kb1);
            const int buf = uc & 1;
            const int r = m0 + quarter * 32 + lane;
            int col = -1;
            if (r < P.M) {
                if (OP == BW_DF1) {
                    col = r;
                } else {
                    int y, x;
                    tile_inv(r, P.Wp, y, x);
                    if (y < P.H && x < P.W) col = y * P.W + x;
                }
            }
            float* dst = P.out + (long long)b * P.D * P.N + col;
            mbar_wait(t_full + buf, (uint32_t)(uc >> 1) & 1u);
            tc_fence_after();
            for (int c0 = 0; c0 < P.D; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)(buf * 256 + c0), v);
                tmem_ld_wait();
                if (col >= 0) {
                    float* d = dst + (long long)c0 * P.N;
#pragma unroll
                    for (int j = 0; j < 32; ++j, d += P.N) red_add_f32(d, v[j] * P.scale);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(t_empty + buf, 0);
        }
    } else if (warp < BF_CONV0 + BF_CONV_WARPS) {
        // ================= converters =================
        // thread = one patch row (8 consecutive targets of one map row) of one query; 4 sweeps cover the box
        constexpr int C = (OP == BW_DF1) ? 64 : 128;        // targets per row of the box
        constexpr int CPR = C / 8;                          // patch rows per box row
        constexpr int RPS = 256 / CPR;                      // box rows per sweep
        constexpr int SWEEPS = ROWS / RPS;                  // 4
        const int ctid = threadIdx.x - BF_CONV0 * 32;
        const int j = ctid % CPR, r0 = ctid / CPR, c = j * 8;
        const uint32_t ring_s = smem_u32(ring), rring_s = smem_u32(rring), cring_s = smem_u32(cring);
        const uint32_t raw_off = (uint32_t)(r0 * C + c) * 4u;
        const uint32_t a_off = (uint32_t)((c >> 6) * (ROWS * 128) + r0 * 128 + (((((c & 63) >> 3)) ^ (r0 & 7)) << 4));
        // this thread's cells inside the coarse boxes: patch pg of 4-patch group g (dF1: one group)
        const int g = j >> 3, pg = (j >> 1) & 3;
        const uint32_t c1_off = (uint32_t)(BA_C1 + (g * ROWS + r0) * 64 + pg * 16);
        const uint32_t c2_off = (uint32_t)(BA_C2 + (g * ROWS + r0) * 32 + pg * 8);
        const uint32_t c3_off = (uint32_t)(BA_C3 + (g * ROWS + r0) * 16 + pg * 4);
        int n_it = 0;
        for (int u = u_begin; u < u_end; u += u_step) {
            int b, m0, kb0, kb1;
            decode(u, b, m0, kb0, kb1);
            n_it += kb1 - kb0;
        }
        for (int it = 0; it < n_it; ++it) {
            const int s = it % BA_STAGES, cs = it % BA_CSTAGES, rs = it % BA_RSTAGES;
            float w1[SWEEPS][4];
            if (ctid == 0) BF_TRACE(it, 2);
            if (nlev > 1) {
                mbar_wait(c_full + cs, (uint32_t)(it / BA_CSTAGES) & 1u);
                if (ctid == 0) BF_TRACE(it, 0);
                const uint32_t cb = cring_s + (uint32_t)(cs * BA_CBYTES);
#pragma unroll
                for (int i = 0; i < SWEEPS; ++i) {
                    float c1[4], c2a = 0.f, c2b = 0.f, c3 = 0.f;
                    lds128f(cb + c1_off + i * (RPS * 64), c1);
                    if (nlev > 2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(c2a), "=f"(c2b) : "r"(cb + c2_off + i * (RPS * 32)));
                    if (nlev > 3) asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(c3) : "r"(cb + c3_off + i * (RPS * 16)));
                    const float w2a = c2a + 0.25f * c3, w2b = c2b + 0.25f * c3;
                    w1[i][0] = 0.25f * (c1[0] + 0.25f * w2a);
                    w1[i][1] = 0.25f * (c1[1] + 0.25f * w2a);
                    w1[i][2] = 0.25f * (c1[2] + 0.25f * w2b);
                    w1[i][3] = 0.25f * (c1[3] + 0.25f * w2b);
                }
            }
            const uint32_t st = ring_s + (uint32_t)(s * BA_STAGE_BYTES), rb = rring_s + (uint32_t)(rs * BA_RAW_BYTES);
            mbar_wait(raw_full + rs, (uint32_t)(it / BA_RSTAGES) & 1u);
            if (ctid == 0) BF_TRACE(it, 3);
            float xs[SWEEPS][8];
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                lds128f(rb + raw_off + i * (RPS * C * 4), xs[i]);
                lds128f(rb + raw_off + i * (RPS * C * 4) + 16, xs[i] + 4);
            }
            uint32_t h[SWEEPS][4], lo[SWEEPS][4];
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                float* x = xs[i];
                if (nlev > 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[k] += w1[i][k >> 1];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
                    h[i][k] = *reinterpret_cast<const uint32_t*>(&hh);
                    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * k] - __uint_as_float(h[i][k] << 16),
                                                                    x[2 * k + 1] - __uint_as_float(h[i][k] & 0xffff0000u));
                    lo[i][k] = *reinterpret_cast<const uint32_t*>(&ll);
                }
            }
            // the box is in registers: hand its buffer back (the vote makes the arrival depend on the converted values,
            // i.e. on the shared loads having completed), THEN wait for the operand stage -- the conversion above overlaps
            // the MMAs that still read it
            {
                const bool landed = __any_sync(0xffffffffu, lo[SWEEPS - 1][3] != 0x7fc1dead);
                if (lane == 0 && landed) { mbar_arrive(raw_empty + rs); if (nlev > 1) mbar_arrive(c_empty + cs); }
            }
            mbar_wait(empty + s, ((uint32_t)(it / BA_STAGES) & 1u) ^ 1u);
            if (ctid == 0) BF_TRACE(it, 4);
#pragma unroll
            for (int i = 0; i < SWEEPS; ++i) {
                sts128(st + a_off + i * (RPS * 128), h[i]);
                if (P.three_pass) sts128(st + a_off + BW_PART_BYTES + i * (RPS * 128), lo[i]);
            }
            fence_proxy_async_smem();                       // generic-proxy writes -> the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(a_part + s);
            if (ctid == 0) BF_TRACE(it, 5);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------- host side
static int encode_bf16(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    return encode_tiled_cached(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides_bytes, box,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

struct BwdLayout { size_t f1_hi, f1_lo, f2_hi, f2_lo, total; int N8, NPk; };

static BwdLayout bwd_layout(int B, int D, int N, int NP) {
    BwdLayout L;
    L.N8 = round_up(N, 8); L.NPk = round_up(NP, 8);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
    L.f1_hi = take((size_t)B * D * L.N8 * 2);
    L.f1_lo = take((size_t)B * D * L.N8 * 2);
    L.f2_hi = take((size_t)B * D * L.NPk * 2);
    L.f2_lo = take((size_t)B * D * L.NPk * 2);
    L.total = off;
    return L;
}

// the default (fold-in-the-GEMM) backward takes any map size; FLOWCORR_BWD_FUSED=0 selects the round-1 pipeline
// (fold + pack pass, then GEMMs from the in-place bf16 planes), which stages a whole query row in shared memory
static bool bwd_fold_in_gemm() { return tunables().bwd_fused != 0; }
bool tc_bwd_supported(int D, int H, int W) {
    const long long NP = (long long)round_up(H, 2) * round_up(W, 8);
    return D % 64 == 0 && D <= 256 && (bwd_fold_in_gemm() || NP <= BW_MAX_NP);
}

size_t tc_bwd_workspace_bytes(int B, int D, int H, int W) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || !tc_bwd_supported(D, H, W)) return 0;
    return bwd_layout(B, D, H * W, round_up(H, 2) * round_up(W, 8)).total + 1024;
}

static int pick_ksplit(int m_units, int kb_total, int n_clusters) {
    int best = 1; long long best_cost = -1;
    for (int s = 1; s <= 8 && s <= kb_total; ++s) {
        const long long rounds = ((long long)m_units * s + n_clusters - 1) / n_clusters;
        const long long cost = rounds * ((kb_total + s - 1) / s) + 2 * rounds;      // + per-unit epilogue/fill slack
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    return best;
}

template <int OP>
static int launch_bwd_gemm(const CUtensorMap* maps, BwdParams P, int B, cudaStream_t s) {
    const size_t smem = 1024 + (size_t)BW_STAGES * BW_STAGE_BYTES + 256;
    FC_SMEM_ATTR_ONCE((tc_bwd_kernel<OP>), smem);
    static std::atomic<int> clusters_of[64];                   // occupancy query once per (kernel instantiation, device)
    int dev = 0;
    FC_CUDA(cudaGetDevice(&dev));
    int n_clusters = clusters_of[dev & 63].load(std::memory_order_acquire);
    if (n_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sm_count_cached() & ~1); cfg.blockDim = dim3(BW_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        FC_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, tc_bwd_kernel<OP>, &cfg));
        clusters_of[dev & 63].store(n_clusters, std::memory_order_release);
    }
    if (n_clusters < 1) { set_error("fc_build_bwd: no CTA pair of the tensor-core kernel fits on this device"); return FC_ECUDA; }
    P.ksplit = pick_ksplit(B * P.mp, P.kb_total, n_clusters);
    P.units = B * P.mp * P.ksplit;
    if (n_clusters > P.units) n_clusters = P.units;
    tc_bwd_kernel<OP><<<dim3(2 * n_clusters), BW_THREADS, smem, s>>>(maps[0], maps[1], maps[2], maps[3], P);
    FC_LAUNCH_CHECK("tc_bwd_kernel");
    return FC_OK;
}

template <int OP>
static int launch_bwd_fold_gemm(const CUtensorMap* maps, BwdParams P, const FoldSrc& F, int B, cudaStream_t s) {
    const size_t smem = 1024 + (size_t)BF_STAGES * BF_STAGE_BYTES + 256;
    FC_SMEM_ATTR_ONCE((tc_bwd_fold_kernel<OP>), smem);
    static std::atomic<int> clusters_of[64];
    int dev = 0;
    FC_CUDA(cudaGetDevice(&dev));
    int n_clusters = clusters_of[dev & 63].load(std::memory_order_acquire);
    if (n_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sm_count_cached() & ~1); cfg.blockDim = dim3(BF_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        FC_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, tc_bwd_fold_kernel<OP>, &cfg));
        clusters_of[dev & 63].store(n_clusters, std::memory_order_release);
    }
    if (n_clusters < 1) { set_error("fc_build_bwd: no CTA pair of the tensor-core kernel fits on this device"); return FC_ECUDA; }
    P.ksplit = pick_ksplit(B * P.mp, P.kb_total, n_clusters);
    P.units = B * P.mp * P.ksplit;
    if (n_clusters > P.units) n_clusters = P.units;
    tc_bwd_fold_kernel<OP><<<dim3(2 * n_clusters), BF_THREADS, smem, s>>>(maps[0], maps[2], maps[3], P, F);
    FC_LAUNCH_CHECK("tc_bwd_fold_kernel");
    return FC_OK;
}

template <int OP>
static int launch_bwd_fold_aligned(const CUtensorMap* maps, const CoarseMaps& CMp, BwdParams P, const CoarseGeo& G, int B, cudaStream_t s) {
    const size_t smem = 1024 + (size_t)BA_STAGES * BA_STAGE_BYTES + (size_t)BA_RSTAGES * BA_RAW_BYTES + (size_t)BA_CSTAGES * BA_CBYTES + 256;
    FC_SMEM_ATTR_ONCE((tc_bwd_fold_aligned_kernel<OP>), smem);
    static std::atomic<int> clusters_of[64];
    int dev = 0;
    FC_CUDA(cudaGetDevice(&dev));
    int n_clusters = clusters_of[dev & 63].load(std::memory_order_acquire);
    if (n_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sm_count_cached() & ~1); cfg.blockDim = dim3(BA_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        FC_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, tc_bwd_fold_aligned_kernel<OP>, &cfg));
        clusters_of[dev & 63].store(n_clusters, std::memory_order_release);
    }
    if (n_clusters < 1) { set_error("fc_build_bwd: no CTA pair of the tensor-core kernel fits on this device"); return FC_ECUDA; }
    P.ksplit = pick_ksplit(B * P.mp, P.kb_total, n_clusters);
    P.units = B * P.mp * P.ksplit;
    if (n_clusters > P.units) n_clusters = P.units;
    tc_bwd_fold_aligned_kernel<OP><<<dim3(2 * n_clusters), BA_THREADS, smem, s>>>(maps[0], maps[2], maps[3], CMp, P, G);
    FC_LAUNCH_CHECK("tc_bwd_fold_aligned_kernel");
    return FC_OK;
}

// the coarse levels 1..3 of the gradient pyramid as 5-D tensors {8 cells of a patch sub-row, 2 sub-rows, patches, N, B}
// with the boxes tc_bwd_fold_aligned_kernel fetches per k-block: [rows][2 patches x 8], [rows][8], [rows][4] cells
static int encode_coarse(CoarseMaps& CMp, float* gpyr, const Pyramid& pyr, int rows) {
    for (int l = 1; l < pyr.L && l < 4; ++l) {
        const cuuint64_t ms = (cuuint64_t)pyr.lv[l].Hp * pyr.lv[l].Wp;
        const cuuint64_t dims[5] = {8, 2, ms / 16, (cuuint64_t)pyr.N, (cuuint64_t)pyr.B};
        const cuuint64_t str[4] = {32, 64, ms * 4, (cuuint64_t)pyr.N * ms * 4};
        const cuuint32_t box[5] = {l == 3 ? 4u : 8u, 1, l == 1 ? 2u : 1u, (cuuint32_t)rows, 1};
        if (int e = encode_tiled_cached(&CMp.m[l - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, gpyr + pyr.lv[l].offset, dims, str, box,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE)) return e;
    }
    return FC_OK;
}

int tc_build_bwd(float* gpyr, const float* f1, const float* f2, float* d1, float* d2, const Pyramid& pyr,
                 int D, int H, int W, int math, void* ws, size_t ws_bytes, cudaStream_t s) {
    const int B = pyr.B, N = pyr.N, Wp = pyr.lv[0].Wp, NP = pyr.lv[0].Hp * Wp;
    FC_REQUIRE(tc_bwd_supported(D, H, W), "fc_build_bwd: tensor-core modes need D %% 64 == 0, D <= 256 and a padded map of "
               "at most %d targets (got D=%d, %d targets); use FC_MATH_FP32", BW_MAX_NP, D, NP);
    const BwdLayout L = bwd_layout(B, D, N, NP);
    uint8_t* w8 = static_cast<uint8_t*>(ws);
    const size_t shift = w8 ? ((1024 - (reinterpret_cast<uintptr_t>(w8) & 1023)) & 1023) : 0;
    if (!w8 || ws_bytes < L.total + shift) {
        set_error("fc_build_bwd: workspace %zu < %zu bytes (fc_build_bwd_workspace_bytes)", ws_bytes, L.total + shift);
        return FC_EWORKSPACE;
    }
    w8 += shift;
    const int three = (math == FC_MATH_TC_3XBF16) ? 1 : 0;
    float* g0 = gpyr + pyr.lv[0].offset;

    const bool fused = bwd_fold_in_gemm();
    // patch-row aligned maps (a multiple of 4 patches per map row, at most 4 levels) take the kernel with TMA-fetched coarse
    // boxes; FLOWCORR_BWD_FUSED=2 keeps the generic fold-in-GEMM kernel for them too
    const bool aligned = fused && tunables().bwd_fused == 1 && pyr.L <= 4 && ((Wp >> 3) & 3) == 0;
    CoarseGeo CG{};
    CG.L = pyr.L;
    for (int l = 0; l < pyr.L && l < 4; ++l) CG.ppr[l] = pyr.lv[l].Wp >> 3;
    CoarseMaps CMp{};
    FoldSrc FS{};
    FS.L = pyr.L;
    for (int l = 0; l < pyr.L; ++l) {
        FS.lvl[l] = gpyr + pyr.lv[l].offset;
        FS.H[l] = pyr.lv[l].H; FS.W[l] = pyr.lv[l].W; FS.Wp[l] = pyr.lv[l].Wp; FS.msize[l] = pyr.lv[l].Hp * pyr.lv[l].Wp;
    }
    if (!fused) {   // fold the pyramid into level 0 and split it into bf16 planes, in place
        FoldParams F{};
        F.L = pyr.L; F.NP = NP; F.two_planes = three;
        int so = 0;
        for (int l = 0; l < pyr.L; ++l) {
            F.lvl[l] = gpyr + pyr.lv[l].offset;
            F.H[l] = pyr.lv[l].H; F.W[l] = pyr.lv[l].W; F.Wp[l] = pyr.lv[l].Wp; F.msize[l] = pyr.lv[l].Hp * pyr.lv[l].Wp;
            F.soff[l] = so; so += F.msize[l];
        }
        F.in_floats = so;
        const size_t smem = (size_t)so * 4 + (size_t)NP * 4;
        const dim3 grid((unsigned)((long long)B * N));
#define FC_FOLD_CASE(LV)                                                                                              \
    case LV:                                                                                                          \
        FC_SMEM_ATTR_GROW((bwd_fold_pack_kernel<LV>), smem);   /* smem varies with the geometry */ \
        bwd_fold_pack_kernel<LV><<<grid, 256, smem, s>>>(F);                                                         \
        break;
        switch (pyr.L) {
            FC_FOLD_CASE(1) FC_FOLD_CASE(2) FC_FOLD_CASE(3) FC_FOLD_CASE(4) FC_FOLD_CASE(5) FC_FOLD_CASE(6)
            default: set_error("fc_build_bwd: %d levels", pyr.L); return FC_EINVAL;
        }
#undef FC_FOLD_CASE
        FC_LAUNCH_CHECK("bwd_fold_pack_kernel");
    }
    __nv_bfloat16* f1_hi = reinterpret_cast<__nv_bfloat16*>(w8 + L.f1_hi);
    __nv_bfloat16* f1_lo = reinterpret_cast<__nv_bfloat16*>(w8 + L.f1_lo);
    __nv_bfloat16* f2_hi = reinterpret_cast<__nv_bfloat16*>(w8 + L.f2_hi);
    __nv_bfloat16* f2_lo = reinterpret_cast<__nv_bfloat16*>(w8 + L.f2_lo);
    {
        BwdPackParams K{};
        K.src[0] = f1; K.src[1] = f2;
        K.hi[0] = f1_hi; K.lo[0] = f1_lo; K.hi[1] = f2_hi; K.lo[1] = f2_lo;
        K.D = D; K.N = N; K.N8 = L.N8; K.NP = NP; K.NPk = L.NPk; K.H = H; K.W = W; K.Wp = Wp; K.two_planes = three;
        const int chunks = (L.NPk > L.N8 ? L.NPk : L.N8) / 8;
        bwd_pack_kernel<<<dim3((unsigned)((chunks + 255) / 256), (unsigned)D, (unsigned)(2 * B)), 256, 0, s>>>(K);
        FC_LAUNCH_CHECK("bwd_pack_kernel");
    }

    // the G planes as a 3-D bf16 tensor {q' (contiguous), p, b}: row pitch 4 NP bytes
    const uint8_t* g_hi = reinterpret_cast<const uint8_t*>(g0);
    const uint8_t* g_lo = g_hi + (size_t)NP * 2;
    const cuuint64_t gdims[3] = {(cuuint64_t)NP, (cuuint64_t)N, (cuuint64_t)B};
    const cuuint64_t gstr[2] = {(cuuint64_t)NP * 4, (cuuint64_t)N * NP * 4};
    auto encode_raw = [&](CUtensorMap* map, cuuint32_t box_q, cuuint32_t box_p) {      // fp32 level 0 as {q', p, b}
        const cuuint32_t box[3] = {box_q, box_p, 1};
        return encode_tiled_cached(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g0, gdims, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    };

    BwdParams P{};
    P.D = D; P.N = N; P.NP = NP; P.H = H; P.W = W; P.Wp = Wp;
    P.three_pass = three; P.scale = 1.0f / sqrtf((float)D);
    const cuuint32_t bbox[2] = {(cuuint32_t)BW_BK, (cuuint32_t)(D / 2)};
    CUtensorMap maps[4];
    if (d1) {
        FC_CUDA(cudaMemsetAsync(d1, 0, (size_t)B * D * N * 4, s));
        const cuuint32_t abox[3] = {(cuuint32_t)BW_BK, (cuuint32_t)BW_BM, 1};
        if (fused) {
            if (int e = encode_raw(&maps[0], BW_BK, BW_BM)) return e;
        } else {
            if (int e = encode_bf16(&maps[0], g_hi, 3, gdims, gstr, abox)) return e;
            if (int e = encode_bf16(&maps[1], three ? g_lo : g_hi, 3, gdims, gstr, abox)) return e;
        }
        const cuuint64_t bdims[2] = {(cuuint64_t)NP, (cuuint64_t)B * D};
        const cuuint64_t bstr[1] = {(cuuint64_t)L.NPk * 2};
        if (int e = encode_bf16(&maps[2], f2_hi, 2, bdims, bstr, bbox)) return e;
        if (int e = encode_bf16(&maps[3], three ? f2_lo : f2_hi, 2, bdims, bstr, bbox)) return e;
        P.out = d1; P.M = N; P.kb_total = (NP + BW_BK - 1) / BW_BK; P.mp = (N + 2 * BW_BM - 1) / (2 * BW_BM);
        if (aligned) {
            if (int e = encode_coarse(CMp, gpyr, pyr, BW_BM)) return e;
            if (int e = launch_bwd_fold_aligned<BW_DF1>(maps, CMp, P, CG, B, s)) return e;
        } else if (int e = fused ? launch_bwd_fold_gemm<BW_DF1>(maps, P, FS, B, s) : launch_bwd_gemm<BW_DF1>(maps, P, B, s)) return e;
    }
    if (d2) {
        FC_CUDA(cudaMemsetAsync(d2, 0, (size_t)B * D * N * 4, s));
        const cuuint32_t abox[3] = {64, 64, 1};
        if (fused) {
            if (int e = encode_raw(&maps[0], BW_BM, BW_BK)) return e;
        } else {
            if (int e = encode_bf16(&maps[0], g_hi, 3, gdims, gstr, abox)) return e;
            if (int e = encode_bf16(&maps[1], three ? g_lo : g_hi, 3, gdims, gstr, abox)) return e;
        }
        const cuuint64_t bdims[2] = {(cuuint64_t)N, (cuuint64_t)B * D};
        const cuuint64_t bstr[1] = {(cuuint64_t)L.N8 * 2};
        if (int e = encode_bf16(&maps[2], f1_hi, 2, bdims, bstr, bbox)) return e;
        if (int e = encode_bf16(&maps[3], three ? f1_lo : f1_hi, 2, bdims, bstr, bbox)) return e;
        P.out = d2; P.M = NP; P.kb_total = (N + BW_BK - 1) / BW_BK; P.mp = (NP + 2 * BW_BM - 1) / (2 * BW_BM);
        if (aligned) {
            if (int e = encode_coarse(CMp, gpyr, pyr, BW_BK)) return e;
            if (int e = launch_bwd_fold_aligned<BW_DF2>(maps, CMp, P, CG, B, s)) return e;
        } else if (int e = fused ? launch_bwd_fold_gemm<BW_DF2>(maps, P, FS, B, s) : launch_bwd_gemm<BW_DF2>(maps, P, B, s)) return e;
    }
    return FC_OK;
}

}  // namespace fc

#ifdef FC_PROBES
extern "C" int fc_debug_bwd_trace(unsigned long long* host_out) {      // 2 x 256 x BF_TRACE_SLOTS stamps (see BF_TRACE)
    return cudaMemcpyFromSymbol(host_out, fc::fc_bwd_trace_buf, sizeof(fc::fc_bwd_trace_buf)) == cudaSuccess ? 0 : 1;
}
#endif
