// Tensor-core build of the correlation volume (CorrBlock.__init__, corr.py:13-27,52-60)
// for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
//   pack pre-pass : (B, D, N) fp32 NCHW  ->  K-major bf16 rows [row][D], hi = bf16(x) and
//                   lo = bf16(x - hi).  fmap1 rows = queries p; fmap2 rows = PADDED targets in
//                   patch order q' = tile_off(y, x) (fc_common.cuh; pad rows are zero), so a
//                   GEMM output row IS a query's level-0 map in its final memory layout and
//                   pad entries come out as exact zeros.
//   GEMM          : persistent CTA pairs (cta_group::2, M = 256).  A pair takes whole query pair-tiles
//                   c, c + n_pairs, ... so that all pairs work on two or three samples at a time and the
//                   packed target operand stays hot in L2.  The query operand (hi and lo, all of K)
//                   stays resident in shared memory; target tiles of NT = 2 rows x Wp (two tiles per
//                   row pair, split at a patch boundary, when 2*Wp > 256) stream through a 4-stage TMA
//                   ring in 128-row x 64-k sub-stages.  FC_MATH_TC_3XBF16 issues hi*hi + lo*hi + hi*lo
//                   into the same fp32 TMEM accumulator (error ~4e-6);
//                   FC_MATH_TC_BF16 issues hi*hi only.  Two 256-column accumulators double
//                   buffer the MMA against the epilogue.
//   epilogue      : 4 (default) or 8 warps, tcgen05.ld 32 lanes x 32 columns, swizzled staging box +
//                   one TMA tensor store per 32 queries x 128 bytes of level 0; levels 1-3 are pooled
//                   thread-locally (bit-exact avg_pool2d) and leave as 32-byte runs.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, then the epilogue
// warps (TMEM lane quarter = warp_id % 4): warps 2-5, or warps 4-11 with two warps per quarter
// splitting the tile's columns.
//
// Diagnostic switches (fc::tunables(): read from the environment once per process, none needed in production):
// FLOWCORR_BUILD_EPI_WARPS=4|8, FLOWCORR_BUILD_STAGES=1..4, FLOWCORR_BUILD_SCHED=0|1, FLOWCORR_NO_FUSE; the stage probes
// (FLOWCORR_PROBE, results are garbage) exist only in a library compiled with -DFC_PROBES.
#include <cstdlib>

#include "fc_build.cuh"

namespace fc {

int simt_pool_levels(float* pyramid, const Pyramid& pyr, int first_level, cudaStream_t s);   // fc_simt.cu
int simt_zero_pad_rows(float* pyramid, const Pyramid& pyr, int first, int last, cudaStream_t s);

// Epilogue warps per CTA (template parameter EW): 4 -> 192 threads (warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, one per
// TMEM lane quarter, two staging boxes each), 8 -> 384 threads (warps 2-3 idle, warps 4-11 epilogue, two per quarter
// splitting the tile's chunks, one staging box each).
__host__ __device__ constexpr int tc_threads(int ew) { return ew == 4 ? 192 : 384; }
__host__ __device__ constexpr int tc_first_epi_warp(int ew) { return ew == 4 ? 2 : 4; }
constexpr int TC_BM = 128;           // queries per CTA (UMMA M)
constexpr int TC_BK = 64;            // bf16 elements per 128-byte swizzle row
#ifndef FC_TC_STAGES
#define FC_TC_STAGES 5
#endif
constexpr int TC_STAGES = FC_TC_STAGES;                   // operand ring depth (barrier arrays); the 8-warp epilogue keeps 8
__host__ __device__ constexpr int tc_ring_stages(int ew) { return (ew == 8 && TC_STAGES > 4) ? 4 : TC_STAGES; }   // staging boxes -> 4
constexpr int TC_STAGE_BYTES = 128 * TC_BK * 2;           // 16 KB: this CTA's half (<= 128 rows) of a target tile x 64 k
constexpr int TC_ABLK_BYTES = TC_BM * TC_BK * 2;          // 16 KB per k-block of the query tile
constexpr int TC_STG_FLOATS = 32 * 32;                    // one TMA-store box: [32 queries][32 floats] = 4 KB
#ifndef FC_TC_STG_BOXES
#define FC_TC_STG_BOXES 4
#endif
// staging boxes per CTA: one per epilogue warp.  With EW = 4 that leaves room for a FIFTH ring stage next to the resident
// query operand: same-box round robin (profiles/r02b_build_ring_ab.txt) 4 stages + 8 boxes 0.655 ms, 5 stages + 4 boxes
// 0.628 ms; level-0 stores straight from registers (no staging, 5 or 6 stages) 0.69 ms.
__host__ __device__ constexpr int tc_stg_boxes(int ew) { return ew == 8 ? 8 : FC_TC_STG_BOXES; }
__host__ __device__ constexpr int tc_stg_bytes(int ew) { return tc_stg_boxes(ew) * TC_STG_FLOATS * 4; }


// ---------------------------------------------------------------- pack pre-pass
// One launch for both operands (blockIdx.z = 2*b + which):
//   which 0: fmap1 (B, D, N) fp32 -> rows p = query,              hi/lo [b*N  + p ][D] bf16, scaled by `prescale`
//   which 1: fmap2 (B, D, N) fp32 -> rows q' = PADDED target in patch order (tile_inv),
//                                                                 hi/lo [b*NP + q'][D] bf16; pad rows are zeros
// hi = bf16(x), lo = bf16(x - hi).
struct PackParams {
    const float* src[2];
    __nv_bfloat16* hi[2];
    __nv_bfloat16* lo[2];      // nullptr: single-pass bf16 mode
    int D, N, NP, H, W, Wp;
    float prescale;
};

__global__ void __launch_bounds__(256) pack_bf16_kernel(const PackParams P) {
    // lane <-> token (coalesced 128-byte reads per channel), warp <-> 16 consecutive channels:
    // every thread writes one full 32-byte sector of its row in hi and in lo
    const int which = blockIdx.z & 1, b = blockIdx.z >> 1;
    const int rows = which ? P.NP : P.N;                     // rows per sample of the packed operand
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const int d0 = (blockIdx.y * 8 + warp) * 16;
    const float* __restrict__ s = P.src[which] + (long long)b * P.D * P.N;
    __nv_bfloat16* __restrict__ hi = P.hi[which];
    __nv_bfloat16* __restrict__ lo = P.lo[which];
    const float scale = which ? 1.0f : P.prescale;
    if (n < P.N && d0 < P.D) {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __ldg(s + (long long)(d0 + i) * P.N + n) * scale;
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
            const __nv_bfloat162 hh(h0, h1);
            const __nv_bfloat162 ll(__float2bfloat16_rn(x[2 * i] - __bfloat162float(h0)),
                                    __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1)));
            h[i] = *reinterpret_cast<const uint32_t*>(&hh);
            l[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const int row = which ? tile_off(n / P.W, n % P.W, P.Wp) : n;
        const long long o = ((long long)b * rows + row) * P.D + d0;
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(hi + o), "r"(h[0]), "r"(h[1]),
                     "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
        if (lo != nullptr)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(lo + o), "r"(l[0]), "r"(l[1]),
                         "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
    }
    // pad rows of the target operand (y >= H or x >= W in patch order) hold zeros; block x scans
    // its share [x * span, (x + 1) * span) of the padded index space
    if (which && P.NP > P.N && d0 < P.D) {
        const int span = (P.NP + (int)gridDim.x - 1) / (int)gridDim.x;
        const int q_end = min((int)(blockIdx.x + 1) * span, P.NP);
        for (int q = blockIdx.x * span + lane; q < q_end; q += 32) {
            int y, x;
            tile_inv(q, P.Wp, y, x);
            if (y >= P.H || x >= P.W) {
                const long long o = ((long long)b * rows + q) * P.D + d0;
                asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(hi + o), "r"(0u) : "memory");
                if (lo != nullptr)
                    asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(lo + o), "r"(0u) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------- GEMM
__device__ __forceinline__ void st_v8(float* dst, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// 8 / 4 consecutive volume elements from fp32 registers: fp32 volume = 32 / 16 bytes, bf16 volume = 16 / 8 bytes (RN)
// explicit shared-space 16-byte store: through a generic pointer the staging writes compiled to ST.E.128, which ride the
// global pipeline's long scoreboard -- every chunk's proxy fence then waited for them (ncu: long_sb on the fence / BSYNC)
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v(__float2bfloat16_rn(lo), __float2bfloat16_rn(hi));
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void vol_store8(float* dst, const float* v) { st_v8(dst, v); }
__device__ __forceinline__ void vol_store8(__nv_bfloat16* dst, const float* v) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void vol_store4(float* dst, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void vol_store4(__nv_bfloat16* dst, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
template <int VB> struct VolT { using type = float; };
template <> struct VolT<1> { using type = __nv_bfloat16; };

struct TcStoreMaps {
    CUtensorMap l0_c32, l0_c16;    // level 0 as {NP, N, B}: boxes of 32 queries x 32 / 16 columns
};

struct TcParams {
    void* lvl[4];          // fused pyramid: level base pointers (level 0 == vol0), fp32 or bf16 elements
    int lvH[4], lvW[4], lvWp[4], lvHp[4];
    int n_fused;           // levels written by the epilogue (1 = level 0 only)
    int W;                 // valid target columns of level 0
    int N, NP, H, Wp;      // queries per sample, padded targets per sample
    int halves;            // tiles per row pair: 1 (2 * Wp <= 256) or 2 (tile 0 = first NT targets of the row pair in patch
                           // order = both rows x columns [0, NT / 2); tile 1 = the remaining 2 * Wp - NT)
    int NT, NT2;           // targets (accumulator columns) of tile 0 / tile 1 of a row pair (multiples of 16, <= 256)
    int n_rp;              // row pairs per map (Hp / 2)
    int m_tiles;           // ceil(N / 128)
    int mp;                // CTA pairs (256 query rows) per sample
    int groups;            // groups of 4 consecutive row pairs per sample (the level-3 pooling period)
    int units;             // B * mp * groups work units, split evenly over the resident CTA pairs
    int three_pass;        // 1 = hi*hi + lo*hi + hi*lo, 0 = hi*hi
    float scale;           // 1 / sqrt(D)
    int sched;             // 1: pair-tiles strided over the CTA pairs (default), 0: contiguous unit ranges
    int stages;            // operand ring depth in use (<= TC_STAGES; FLOWCORR_BUILD_STAGES for the ring-depth measurement)
    int probe;             // 0 in production; FLOWCORR_PROBE (tools/probe_bounds.py): 1 = epilogue without
                           // global stores, 2 = no MMAs issued, 3 = epilogue neither reads TMEM nor stores,
                           // 5 = pooled-level stores off, 6 = level-0 stores off, 7 = no target-operand loads
};

// Timeline probe (FC_PROBES builds only; tools/probe_build_trace.py): CTA 0 stamps clock64() per tile (accumulator).
// MMA thread: 0 accumulator free, 1 all MMAs of the tile issued + committed, 2 cycles spent waiting for ring stages;
// first epilogue warp: 3 ready for the tile, 4 accumulator complete, 5 accumulator drained (handed back), 6 tile's stores issued;
// producer: 7 cycles spent waiting for free ring stages
#ifdef FC_PROBES
constexpr int TB_SLOTS = 8;
__device__ unsigned long long fc_build_trace_buf[512 * TB_SLOTS];
#define TB_TRACE(tile, k, v) do { if (blockIdx.x == 0 && (tile) < 512) fc_build_trace_buf[(tile) * TB_SLOTS + (k)] = (v); } while (0)
#define TB_CLOCK() clock64()
// finer: the chunks of tiles 16..23 of the same warp: 0 before the TMEM load, 1 loaded, 2 staging box free, 3 store issued, 4 pooled
__device__ unsigned long long fc_build_chunk_trace[8 * 8 * 8];
#define TB_CHUNK(tile, cc, k) do { if (blockIdx.x == 0 && ew == 0 && lane == 0 && (tile) >= 16 && (tile) < 24) \
    fc_build_chunk_trace[(((tile) - 16) * 8 + (cc)) * 8 + (k)] = clock64(); } while (0)
#else
#define TB_CHUNK(tile, cc, k) do {} while (0)
#define TB_TRACE(tile, k, v) do {} while (0)
#define TB_CLOCK() 0ll
#endif

template <int KB, int EW, int VB>   // KB = D / 64 k-blocks, EW = epilogue warps (4 or 8), VB = 1: bf16 volume
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc_threads(EW), 1)
tc_build_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const __grid_constant__ CUtensorMap map_b2_hi, const __grid_constant__ CUtensorMap map_b2_lo,
                const __grid_constant__ TcStoreMaps SM, const TcParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve (all operand regions 1024-byte aligned for SWIZZLE_128B)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + KB * TC_ABLK_BYTES;
    uint8_t* ring = a_lo + KB * TC_ABLK_BYTES;
    float* stg = reinterpret_cast<float*>(ring + tc_ring_stages(EW) * TC_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg) + tc_stg_bytes(EW));
    uint64_t* a_full = bars;                       // 1
    uint64_t* b_full = bars + 1;                   // TC_STAGES
    uint64_t* b_empty = b_full + TC_STAGES;        // TC_STAGES
    uint64_t* t_full = b_empty + TC_STAGES;        // 2
    uint64_t* t_empty = t_full + 2;                // 2
    uint64_t* a_empty = t_empty + 2;               // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_parts = P.three_pass ? 2 : 1;
    // CTA pair: rank 0 (the leader) issues every MMA for both; each CTA owns 128 query rows and
    // streams its half of every target tile.  Full barriers live in the leader; empty / tmem-full
    // barriers are signalled in both CTAs by multicast commits.
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    // Persistent schedule: work unit u = (sample, query pair-tile, group of 4 row pairs), u = pair-tile * groups + group.
    // A CTA pair takes WHOLE pair-tiles c, c + n_clusters, c + 2 n_clusters, ... (the resident query operand is reloaded
    // only when the pair-tile changes) so that at any time all pairs work on neighbouring pair-tiles, i.e. on two or
    // three samples whose target operand stays hot in L2 instead of all B of them (FLOWCORR_BUILD_SCHED=0: contiguous
    // unit ranges); the pair-tiles left over after the last full round are dealt out unit by unit.  All three roles
    // walk the same sequence.
    const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int n_pt = P.units / P.groups;                               // pair-tiles in all
    const int full_rounds = n_pt / n_clusters;
    const int tail_units = (n_pt - full_rounds * n_clusters) * P.groups;
    int u_begin = 0, u_end = 0;                                        // local unit counter [u_begin, u_end)
    if (P.sched == 0) {
        u_begin = (int)((long long)P.units * cluster_id / n_clusters);
        u_end = (int)((long long)P.units * (cluster_id + 1) / n_clusters);
    } else {
        // the tail's units go round robin: cluster c takes tail units c, c + n_clusters, ...
        u_end = full_rounds * P.groups + (tail_units > cluster_id ? (tail_units - cluster_id + n_clusters - 1) / n_clusters : 0);
    }
    auto unit_of = [&](int j) {                                        // j-th local unit -> global unit
        if (P.sched == 0) return j;
        const int k = j / P.groups;
        if (k < full_rounds) return (cluster_id + k * n_clusters) * P.groups + (j - k * P.groups);
        return full_rounds * n_clusters * P.groups + cluster_id + (j - full_rounds * P.groups) * n_clusters;
    };
    auto decode = [&](int j, int& b, int& m0, int& t0, int& t1) {        // [t0, t1): ROW PAIRS of the unit
        const int u = unit_of(j);
        const int am = u / P.groups, g = u - am * P.groups;
        b = am / P.mp;
        m0 = ((am - b * P.mp) * 2 + (int)rank) * TC_BM;
        t0 = 4 * g;
        t1 = min(t0 + 4, P.n_rp);
    };

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 2 * EW); }   // EW epilogue warps x 2 CTAs
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
            int a_use = 0, cur_am = -1, slot = 0, ptile = 0;
            uint32_t phase = 0;                                // ring position: stage `slot`, use parity `phase`
            for (int u = u_begin; u < u_end; ++u) {
                int b, m0, t0, t1;
                decode(u, b, m0, t0, t1);
                const int am = unit_of(u) / P.groups;
                if (am != cur_am) {
                    // query operand (resident): wait until the MMAs reading the old one have retired
                    if (a_use > 0) mbar_wait(a_empty, (uint32_t)(a_use - 1) & 1u);
                    if (leader) mbar_expect_tx(a_full, (uint32_t)(2 * n_parts * KB * TC_ABLK_BYTES));
                    for (int kb = 0; kb < KB; ++kb) {
                        tma2_load_2d(a_hi + kb * TC_ABLK_BYTES, &map_a_hi, a_full, kb * TC_BK, b * P.N + m0);
                        if (P.three_pass) tma2_load_2d(a_lo + kb * TC_ABLK_BYTES, &map_a_lo, a_full, kb * TC_BK, b * P.N + m0);
                    }
                    ++a_use; cur_am = am;
                }
                // target operand: stages of [NT/2 rows][64 k], hi then lo of each k-block
                for (int h = 0; h < P.halves; ++h)                 // tile order of a unit: half 0 of its 4 row pairs, then half 1
                  for (int rp = t0; rp < t1; ++rp) {
                    long long pw = 0;
                    const int ncols = h ? P.NT2 : P.NT;            // each CTA streams its half of the tile's targets
                    const int row0 = b * P.NP + rp * 2 * P.Wp + h * P.NT + (int)rank * (ncols >> 1);
                    for (int kb = 0; kb < KB; ++kb)
                        for (int part = 0; part < n_parts; ++part) {
                            const int s = slot;
                            const uint32_t ph = phase;
                            if (++slot == P.stages) { slot = 0; phase ^= 1u; }
                            { const long long w0 = TB_CLOCK(); mbar_wait(b_empty + s, ph ^ 1u); pw += TB_CLOCK() - w0; }
                            if (FC_PROBE_VAL(P) == 7) {                // no target loads (stage probe)
                                if (leader) mbar_arrive(b_full + s);
                                continue;
                            }
                            if (leader) mbar_expect_tx(b_full + s, (uint32_t)(ncols * TC_BK * 2));  // both halves
                            const CUtensorMap* mp = part == 0 ? (h ? &map_b2_hi : &map_b_hi) : (h ? &map_b2_lo : &map_b_lo);
                            tma2_load_2d(ring + s * TC_STAGE_BYTES, mp, b_full + s, kb * TC_BK, row0);
                        }
                    TB_TRACE(ptile, 7, (unsigned long long)pw);
                    ++ptile;
                  }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (leader && elect_one()) {
            const uint32_t idesc0 = umma_idesc_bf16(2 * TC_BM, P.NT), idesc1 = umma_idesc_bf16(2 * TC_BM, P.NT2 > 0 ? P.NT2 : P.NT);
            int tc = 0, a_use = 0, cur_am = -1, slot = 0;
            uint32_t phase = 0;
            for (int u = u_begin; u < u_end; ++u) {
                int b, m0, t0, t1;
                decode(u, b, m0, t0, t1);
                const int am = unit_of(u) / P.groups;
                if (am != cur_am) {
                    if (cur_am >= 0) umma2_commit(a_empty);    // old query operand is free once everything issued retires
                    mbar_wait(a_full, (uint32_t)a_use & 1u);
                    tc_fence_after();
                    ++a_use; cur_am = am;
                }
                for (int tt = 0; tt < (t1 - t0) * P.halves; ++tt, ++tc) {
                    const uint32_t idesc = tt >= (t1 - t0) ? idesc1 : idesc0;      // second run of the unit = half 1
                    const int buf = tc & 1;
                    mbar_wait(t_empty + buf, ((uint32_t)(tc >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    TB_TRACE(tc, 0, (unsigned long long)TB_CLOCK());
                    long long rw = 0;
                    const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256);
                    for (int kb = 0; kb < KB; ++kb)
                        for (int part = 0; part < n_parts; ++part) {
                            const int s = slot;
                            const uint32_t ph = phase;
                            if (++slot == P.stages) { slot = 0; phase ^= 1u; }
                            { const long long w0 = TB_CLOCK(); mbar_wait(b_full + s, ph); rw += TB_CLOCK() - w0; }
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(ring + s * TC_STAGE_BYTES);
                            const uint32_t ah_addr = smem_u32(a_hi + kb * TC_ABLK_BYTES);
                            const uint32_t al_addr = smem_u32(a_lo + kb * TC_ABLK_BYTES);
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; ++k) {
                                if (FC_PROBE_VAL(P) == 2) break;
                                const uint64_t bd = umma_desc_sw128(b_addr + k * 32);
                                // part 0: B = hi -> A_hi*B_hi (+ A_lo*B_hi); part 1: B = lo -> A_hi*B_lo
                                umma2_bf16(d_addr, umma_desc_sw128(ah_addr + k * 32), bd, idesc, (kb | part | k) != 0 ? 1u : 0u);
                                if (part == 0 && P.three_pass)
                                    umma2_bf16(d_addr, umma_desc_sw128(al_addr + k * 32), bd, idesc, 1u);
                            }
                            umma2_commit(b_empty + s);         // frees the ring slot in both CTAs when these MMAs retire
                        }
                    umma2_commit(t_full + buf);                // accumulator complete: both epilogues
                    TB_TRACE(tc, 1, (unsigned long long)TB_CLOCK());
                    TB_TRACE(tc, 2, (unsigned long long)rw);
                }
            }
        }
    } else if (warp >= tc_first_epi_warp(EW)) {
        // ================= epilogue (8 warps) =================
        // TMEM -> registers -> (scale, 2x2 pooling) -> shared staging -> TMA tensor store.
        // A thread owns one query row (TMEM lane quarter = warp % 4); the two warps of a quarter
        // split the tile's 32-column chunks (half = first or second run of `CH` chunks), so every
        // scheduler holds two epilogue warps and one warp's dependent-issue and TMA-store waits
        // are covered by the other (one warp per scheduler ran at 0.17 IPC and bounded the
        // kernel: profiles/r01f).  A chunk = 32 accumulator columns = 2 patches = 16 target
        // columns x 2 rows = 128 contiguous bytes of the query's map.  The staging box
        // [32 queries][32 floats] is what the store maps describe: one elected lane writes
        // 32 queries x 128 bytes with a single instruction; rows beyond the sample are clipped
        // by the TMA unit.
        constexpr int NH = EW / 4;                             // warps per TMEM lane quarter
        constexpr int CMAX = 8 / NH;                           // chunk slots per warp
        constexpr int NBUF = tc_stg_boxes(EW) / EW;            // staging boxes per warp
        static_assert((NBUF & (NBUF - 1)) == 0, "staging boxes per epilogue warp: a power of two (0 = stores from registers)");
        using vol_t = typename VolT<VB>::type;
        vol_t* const lv1 = static_cast<vol_t*>(P.lvl[1]);
        vol_t* const lv2 = static_cast<vol_t*>(P.lvl[2]);
        vol_t* const lv3 = static_cast<vol_t*>(P.lvl[3]);
        const int ew = warp - tc_first_epi_warp(EW);
        const int quarter = warp & 3, half_id = ew >> 2;
        float* sbuf0 = stg + ew * NBUF * TC_STG_FLOATS;        // this warp's 4 KB staging box(es)
        uint32_t use = 0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const bool do_scale = P.scale != 1.0f;
        const bool fused = P.n_fused > 1;
        const int W1 = P.lvW[1], W2 = P.lvW[2], W3 = P.lvW[3];
        const long long ms1 = (long long)P.lvHp[1] * P.lvWp[1];
        const long long ms2 = (long long)P.lvHp[2] * P.lvWp[2], ms3 = (long long)P.lvHp[3] * P.lvWp[3];
        int b = 0, tc = 0;
        float s2[CMAX][4], s3[CMAX][2];                        // level-2 / level-3 partial sums across the row pairs of a run
#pragma unroll
        for (int i = 0; i < CMAX; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s2[i][j] = 0.f;
            s3[i][0] = s3[i][1] = 0.f;
        }
        int row0 = 0, rows_valid = 0;                          // first query row of this warp inside the sample
        bool mine = false;
        long long qrow = 0;

        // One tile = half H of row pair rp (the whole row pair when 2 * Wp <= 256).  A unit walks half 0 of its four
        // row pairs, then half 1, so one set of stashes serves both.
        auto tile_body = [&](const int H, int rp) {
            const int buf = tc & 1;
            if (ew == 0 && lane == 0) TB_TRACE(tc, 3, (unsigned long long)TB_CLOCK());
            mbar_wait(t_full + buf, (uint32_t)(tc >> 1) & 1u);
            tc_fence_after();
            if (ew == 0 && lane == 0) TB_TRACE(tc, 4, (unsigned long long)TB_CLOCK());
            const int q0 = rp * 2 * P.Wp + H * P.NT;           // first padded target of the tile
            const int ncols = H ? P.NT2 : P.NT;
            const int n_chunks = (ncols + 31) >> 5;            // 32-column chunks of the tile (the last may be 16 wide)
            // chunks per warp and the chunk of its cc-th slot.  One warp per quarter: all chunks in order.  Two warps:
            // wide tiles alternate PAIRS of chunks ({0,1,4,5} / {2,3,6,7}) so that the four lines of a 512-byte run reach
            // L2 close together; narrow tiles (<= 2 chunks): the second warp only keeps the barriers moving
            const int CH = NH == 1 ? n_chunks : (n_chunks > 4 ? 4 : (n_chunks > 2 ? 2 : n_chunks));
            const bool ilv = NH == 2 && n_chunks > 4;
            const int c_lo = NH == 1 ? 0 : ((n_chunks > 2 || half_id == 0) ? half_id * CH : 8);
            auto chunk_of = [&](int cc) { return ilv ? ((cc >> 1) * 4 + half_id * 2 + (cc & 1)) : (c_lo + cc); };
            float l2[CMAX][4];
            if (FC_PROBE_VAL(P) != 3) {
#pragma unroll
              for (int cc = 0; cc < CMAX; ++cc) {
                const int c = chunk_of(cc);                    // chunk inside the tile
                const int gc = H * (P.NT >> 5) + c;                      // chunk inside the row pair: level-0 columns [16 gc, 16 gc + 16)
                if (cc < CH && c * 32 < ncols) {
                    float v[32];
                    TB_CHUNK(tc, cc, 0);
                    tmem_ld32(lane_addr + (uint32_t)(buf * 256 + c * 32), v);
                    tmem_ld_wait();
                    TB_CHUNK(tc, cc, 1);
                    if (do_scale) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= P.scale;
                    }
                    const int rem = ncols - c * 32;
                    if constexpr (NBUF == 0) {
                        // ---- level 0 straight from registers: a thread owns 128 contiguous bytes (64 with a bf16 volume)
                        // of its query's map; all of shared memory beyond the resident operand goes to the operand ring
                        if (mine && FC_PROBE_VAL(P) != 1 && FC_PROBE_VAL(P) != 6) {
                            vol_t* dst = static_cast<vol_t*>(P.lvl[0]) + qrow * (long long)P.NP + q0 + c * 32;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < 2 || rem >= 32) vol_store8(dst + 8 * k, v + 8 * k);
                        }
                    } else {
                        // ---- level 0: staging (swizzled like the store map) -> TMA store
                        float* sbuf = sbuf0 + (use & (NBUF - 1)) * TC_STG_FLOATS;
                        ++use;
                        if (elect_one()) tma_wait_group_read<NBUF - 1>();   // the store that last read this box is done
                        __syncwarp();
                        TB_CHUNK(tc, cc, 2);
                        const uint32_t sb32 = smem_u32(sbuf);
                        if (VB) {
                            // bf16 volume: rows of 64 bytes (SWIZZLE_64B box of 32 columns) or 32 bytes (plain box of 16)
                            if (rem >= 32) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    st_shared_v4(sb32 + 16u * (uint32_t)(lane * 4 + (k ^ ((lane >> 1) & 3))),
                                                 pack_bf16x2(v[8 * k], v[8 * k + 1]), pack_bf16x2(v[8 * k + 2], v[8 * k + 3]),
                                                 pack_bf16x2(v[8 * k + 4], v[8 * k + 5]), pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
                            } else {
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    st_shared_v4(sb32 + 16u * (uint32_t)(lane * 2 + k),
                                                 pack_bf16x2(v[8 * k], v[8 * k + 1]), pack_bf16x2(v[8 * k + 2], v[8 * k + 3]),
                                                 pack_bf16x2(v[8 * k + 4], v[8 * k + 5]), pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
                            }
                        } else if (rem >= 32) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                st_shared_v4(sb32 + 4u * (uint32_t)(lane * 32 + ((k ^ (lane & 7)) << 2)), __float_as_uint(v[4 * k]),
                                             __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]), __float_as_uint(v[4 * k + 3]));
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                st_shared_v4(sb32 + 4u * (uint32_t)(lane * 16 + ((k ^ ((lane >> 1) & 3)) << 2)), __float_as_uint(v[4 * k]),
                                             __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]), __float_as_uint(v[4 * k + 3]));
                        }
                        TB_CHUNK(tc, cc, 5);
                        fence_proxy_async_smem();
                        __syncwarp();
                        TB_CHUNK(tc, cc, 6);
                        if (elect_one() && rows_valid > 0 && FC_PROBE_VAL(P) != 1 && FC_PROBE_VAL(P) != 6) {
                            tma_store_3d(rem >= 32 ? &SM.l0_c32 : &SM.l0_c16, smem_u32(sbuf), q0 + c * 32, row0, b);
                            tma_commit_group();
                        }
                        TB_CHUNK(tc, cc, 3);
                    }
                    if (fused) {
                        // ---- level 1: row rp, columns [8 gc, 8 gc + 8); ((a + b) + c) + d, then * 0.25:
                        // bit-exact avg_pool2d of the level below (oracle/corr_spec.py::pool_pyramid)
                        float l1[8];
#pragma unroll
                        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float a = __fadd_rn(__fadd_rn(__fadd_rn(v[16 * pp + 2 * j], v[16 * pp + 2 * j + 1]),
                                                                    v[16 * pp + 8 + 2 * j]), v[16 * pp + 8 + 2 * j + 1]);
                                l1[4 * pp + j] = (8 * gc + 4 * pp + j < W1) ? a * 0.25f : 0.f;
                            }
                        if (rp < P.lvH[1] && 8 * gc < P.lvWp[1] && mine && FC_PROBE_VAL(P) != 1 && FC_PROBE_VAL(P) != 5)
                            vol_store8(lv1 + qrow * ms1 + (long long)(rp >> 1) * 2 * P.lvWp[1] + gc * 16 + (rp & 1) * 8, l1);
                        if (P.n_fused > 2) {
                            if ((rp & 1) == 0) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) s2[cc][j] = __fadd_rn(l1[2 * j], l1[2 * j + 1]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float a = __fadd_rn(__fadd_rn(s2[cc][j], l1[2 * j]), l1[2 * j + 1]);
                                    l2[cc][j] = (4 * gc + j < W2) ? a * 0.25f : 0.f;
                                }
                            }
                        }
                    }
                    TB_CHUNK(tc, cc, 4);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) l2[cc][j] = 0.f;
                }
              }
            }
            // all TMEM reads of this tile are done: hand the accumulator back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(t_empty + buf, 0);
            if (ew == 0 && lane == 0) TB_TRACE(tc, 5, (unsigned long long)TB_CLOCK());

            if (P.n_fused > 2 && (rp & 1) && FC_PROBE_VAL(P) != 3) {
                // ---- level 2: row y2 = rp / 2, columns [4 gc, 4 gc + 4) per chunk -> 32-byte runs per chunk pair
                const int y2 = rp >> 1;
                const bool st = mine && FC_PROBE_VAL(P) != 1 && FC_PROBE_VAL(P) != 5;
#pragma unroll
                for (int cp = 0; cp < CMAX / 2; ++cp) {
                    const int gc = H * (P.NT >> 5) + chunk_of(2 * cp);
                    if (2 * cp < CH && y2 < P.lvH[2] && 4 * gc < P.lvWp[2] && st) {
                        const float o[8] = {l2[2 * cp][0], l2[2 * cp][1], l2[2 * cp][2], l2[2 * cp][3],
                                            l2[2 * cp + 1][0], l2[2 * cp + 1][1], l2[2 * cp + 1][2], l2[2 * cp + 1][3]};
                        vol_store8(lv2 + qrow * ms2 + (long long)(y2 >> 1) * 2 * P.lvWp[2] + (gc >> 1) * 16 + (y2 & 1) * 8, o);
                    }
                }
                if (P.n_fused > 3) {
                    // ---- level 3: row y3 = rp / 4, columns [2 gc, 2 gc + 2) per chunk -> 16-byte runs per chunk pair
                    const int y3 = rp >> 2;
                    if ((y2 & 1) == 0) {
#pragma unroll
                        for (int cc = 0; cc < CMAX; ++cc) {
                            s3[cc][0] = __fadd_rn(l2[cc][0], l2[cc][1]);
                            s3[cc][1] = __fadd_rn(l2[cc][2], l2[cc][3]);
                        }
                    } else {
#pragma unroll
                        for (int cp = 0; cp < CMAX / 2; ++cp) {
                            const int gc = H * (P.NT >> 5) + chunk_of(2 * cp);
                            float o[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int cc = 2 * cp + (i >> 1), j = i & 1;
                                const float a = __fadd_rn(__fadd_rn(s3[cc][j], l2[cc][2 * j]), l2[cc][2 * j + 1]);
                                o[i] = (2 * gc + i < W3) ? a * 0.25f : 0.f;
                            }
                            // (narrow tiles: the one active warp also writes the zero pad columns 4..7 of the patch)
                            // (a warp that owns whole level-3 patches completes them: 4 chunks = 8 columns, zeros past the map)
                            if ((2 * cp < (NH == 1 ? ((CH + 3) & ~3) : CH) || n_chunks <= 2) && y3 < P.lvH[3] && 2 * gc < P.lvWp[3] && st)
                                vol_store4(lv3 + qrow * ms3 + (long long)(y3 >> 1) * 2 * P.lvWp[3] +
                                               (gc >> 2) * 16 + (y3 & 1) * 8 + ((gc >> 1) & 1) * 4, o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            }
            if (fused && rp == P.n_rp - 1 && H == P.halves - 1 && mine && FC_PROBE_VAL(P) != 1 && FC_PROBE_VAL(P) != 5) {
                // pad row (y = Hl, Hl odd) of every pooled level: the lookup's TMA boxes read whole
                // row pairs, so it must hold zeros (the tile loop itself never produces it)
                const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int l = 1; l < P.n_fused; ++l)
                    if (P.lvHp[l] > P.lvH[l]) {
                        const int y = P.lvH[l], wp = P.lvWp[l];
                        vol_t* rowp = static_cast<vol_t*>(P.lvl[l]) + qrow * ((long long)P.lvHp[l] * wp) + (long long)(y >> 1) * 2 * wp + (y & 1) * 8;
                        for (int g = half_id; g * 8 < wp; g += NH) vol_store8(rowp + g * 16, z);
                    }
            }
            if (ew == 0 && lane == 0) TB_TRACE(tc, 6, (unsigned long long)TB_CLOCK());
            ++tc;
        };

        for (int u = u_begin; u < u_end; ++u) {
            int m0, t0, t1;
            decode(u, b, m0, t0, t1);
            row0 = m0 + quarter * 32;
            rows_valid = P.N - row0;
            mine = rows_valid > lane;
            qrow = (long long)b * P.N + row0 + lane;
            for (int h = 0; h < P.halves; ++h)
                for (int rp = t0; rp < t1; ++rp) tile_body(h, rp);
        }
        if (elect_one()) tma_wait_group<0>();                  // staging is read and the stores have landed
        __syncwarp();
    }

    // neither CTA may leave while its peer can still touch its barriers / tensor memory
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------- host side
// 2-D bf16 row-major [rows][cols] tensor, box {box_cols, box_rows}; box_cols = 64 -> 128-byte
// swizzle (query operand), 32 -> 64-byte swizzle (target operand stages)
static int make_map(CUtensorMap* map, const void* base, long long rows, int cols, int box_rows, int box_cols) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    return encode_tiled_cached(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box,
                               box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

// volume tensor (fp32 or bf16 elements) of `rank` dims (dim 0 contiguous): the epilogue's store boxes
static int make_vol_map(CUtensorMap* map, void* base, bool bf16, int rank, const cuuint64_t* dims,
                        const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
    return encode_tiled_cached(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, base, dims,
                               strides_bytes, box, swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE);
}

struct TcLayout { size_t a_hi, a_lo, b_hi, b_lo, total; long long NP; };

static TcLayout tc_layout(int B, int D, int H, int W) {
    TcLayout L;
    const long long N = (long long)H * W, NP = (long long)round_up(H, 2) * round_up(W, 8);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
    L.a_hi = take((size_t)B * N * D * 2);
    L.a_lo = take((size_t)B * N * D * 2);
    L.b_hi = take((size_t)B * NP * D * 2);
    L.b_lo = take((size_t)B * NP * D * 2);
    L.total = off; L.NP = NP;
    return L;
}

size_t tc_build_workspace_bytes(int B, int D, int H, int W, int, int) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return tc_layout(B, D, H, W).total + 1024;   // slack to align the base to 1 KB
}

template <int KB, int EW, int VB>
static int launch_tc(const CUtensorMap* maps, const TcStoreMaps& SM, const TcParams& P_in, int B, cudaStream_t s) {
    const size_t smem = 1024 + 2 * KB * TC_ABLK_BYTES + tc_ring_stages(EW) * TC_STAGE_BYTES + tc_stg_bytes(EW) + 256;
    TcParams P = P_in;
    if (P.stages > tc_ring_stages(EW)) P.stages = tc_ring_stages(EW);
    FC_SMEM_ATTR_ONCE((tc_build_kernel<KB, EW, VB>), smem);
    // persistent: one CTA pair (cluster 2x1x1) per co-resident SM pair; the occupancy query runs once per
    // (kernel instantiation, device)
    static std::atomic<int> clusters_of[64];
    int dev = 0, n_clusters = 0;
    FC_CUDA(cudaGetDevice(&dev));
    n_clusters = clusters_of[dev & 63].load(std::memory_order_acquire);
    if (n_clusters == 0) {
        int n_sm = 0;
        FC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_sm & ~1); cfg.blockDim = dim3(tc_threads(EW)); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        FC_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, tc_build_kernel<KB, EW, VB>, &cfg));
        clusters_of[dev & 63].store(n_clusters, std::memory_order_release);
    }
    if (n_clusters < 1) { set_error("fc_build: no CTA pair of the tensor-core kernel fits on this device"); return FC_ECUDA; }
    if (n_clusters > P.units) n_clusters = P.units;
    dim3 grid(2 * n_clusters);
    tc_build_kernel<KB, EW, VB><<<grid, tc_threads(EW), smem, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], SM, P);
    FC_LAUNCH_CHECK("tc_build_kernel");
    return FC_OK;
}

// feat != nullptr: the packed operands come from the fused fnet tail (fc_feat.cu) instead of the pack pre-pass
int tc_build(const float* f1, const float* f2, void* pyramid, const Pyramid& pyr, int D, int H, int W,
             int vol_dtype, int math, void* ws, size_t ws_bytes, cudaStream_t s, const FeatSource* feat) {
    FC_REQUIRE(vol_dtype == FC_VOL_F32 || vol_dtype == FC_VOL_BF16, "fc_build: unknown vol_dtype %d", vol_dtype);
    const bool vb = vol_dtype == FC_VOL_BF16;
    const size_t es = vb ? 2 : 4;
    FC_REQUIRE(!vb || (pyr.L <= 4 && !tunables().no_fuse),
               "fc_build: the bf16 volume is written by the fused epilogue only (num_levels <= 4, got %d)", pyr.L);
    FC_REQUIRE(D % 64 == 0 && D <= 256, "fc_build: tensor-core modes need D %% 64 == 0 and D <= 256 (got %d); use FC_MATH_FP32", D);
    const int Wp = pyr.lv[0].Wp;
    FC_REQUIRE(Wp >= 16 && Wp <= 256, "fc_build: tensor-core modes need 9 <= W <= 256 tokens (got %d); use FC_MATH_FP32", W);
    const int B = pyr.B;
    const TcLayout L = tc_layout(B, D, H, W);
    if (!ws || ws_bytes < L.total) { set_error("fc_build: workspace %zu < %zu bytes", ws_bytes, L.total); return FC_EWORKSPACE; }
    uint8_t* w8 = static_cast<uint8_t*>(ws);
    // align the workspace base to 1 KB ourselves (torch allocations are 512-byte aligned)
    const size_t shift = (1024 - (reinterpret_cast<uintptr_t>(w8) & 1023)) & 1023;
    if (ws_bytes < L.total + shift) { set_error("fc_build: workspace %zu < %zu bytes (after alignment)", ws_bytes, L.total + shift); return FC_EWORKSPACE; }
    w8 += shift;
    const bool three = (math == FC_MATH_TC_3XBF16);
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(w8 + L.a_hi);
    __nv_bfloat16* a_lo = reinterpret_cast<__nv_bfloat16*>(w8 + L.a_lo);
    __nv_bfloat16* b_hi = reinterpret_cast<__nv_bfloat16*>(w8 + L.b_hi);
    __nv_bfloat16* b_lo = reinterpret_cast<__nv_bfloat16*>(w8 + L.b_lo);
    const int N = pyr.N;
    const long long NP = L.NP;

    // 1/sqrt(D) is folded into the query operand when it is a power of two (D = 64, 256:
    // exact, bit-identical to scaling the product); otherwise the epilogue multiplies
    const float inv_sqrt_d = 1.0f / sqrtf((float)D);
    const bool fold_scale = (D == 4 || D == 16 || D == 64 || D == 256);
    if (feat != nullptr) {
        TcPacked dst{a_hi, three ? a_lo : nullptr, b_hi, three ? b_lo : nullptr, B, D, N, (int)NP, H, W, Wp,
                     fold_scale ? inv_sqrt_d : 1.0f};
        if (int e = fnet_tail_pack(*feat, dst, s)) return e;
    } else {
        PackParams K{};
        K.src[0] = f1; K.src[1] = f2;
        K.hi[0] = a_hi; K.hi[1] = b_hi;
        K.lo[0] = three ? a_lo : nullptr; K.lo[1] = three ? b_lo : nullptr;
        K.D = D; K.N = N; K.NP = (int)NP; K.H = H; K.W = W; K.Wp = Wp;
        K.prescale = fold_scale ? inv_sqrt_d : 1.0f;
        // (D % 64 == 0 is required above, so channel chunks of 16 never straddle D)
        dim3 pb(256), pg((unsigned)((N + 31) / 32), (unsigned)((D + 127) / 128), 2 * B);
        pack_bf16_kernel<<<pg, pb, 0, s>>>(K);
        FC_LAUNCH_CHECK("pack_bf16_kernel");
    }

    TcParams P{};
    P.W = W;
    // the epilogue produces the pyramid itself when a tile holds two whole target rows
    const Tunables& T = tunables();
    const bool fuse = pyr.L >= 2 && !T.no_fuse;
    P.n_fused = fuse ? (pyr.L < 4 ? pyr.L : 4) : 1;
    for (int l = 0; l < 4 && l < pyr.L; ++l) {
        P.lvl[l] = static_cast<uint8_t*>(pyramid) + (size_t)pyr.lv[l].offset * es;
        P.lvH[l] = pyr.lv[l].H; P.lvW[l] = pyr.lv[l].W; P.lvWp[l] = pyr.lv[l].Wp; P.lvHp[l] = pyr.lv[l].Hp;
    }
    P.N = N; P.NP = (int)NP; P.H = H; P.Wp = Wp;
    // a tile = a row pair in patch order (both rows of every 8-column patch), split in two when it exceeds the 256
    // accumulator columns of an MMA; the split keeps whole patches together, so pooling stays thread-local
    P.halves = (2 * Wp <= 256) ? 1 : 2;
    P.NT = (P.halves == 1) ? 2 * Wp : (2 * Wp - 256 >= 32 ? 256 : 192);   // (no 16-column MMA: N >= 32 for M = 256)
    P.NT2 = (P.halves == 1) ? 0 : 2 * Wp - P.NT;
    P.n_rp = pyr.lv[0].Hp / 2;
    P.m_tiles = (N + TC_BM - 1) / TC_BM;
    P.mp = (P.m_tiles + 1) / 2;
    P.groups = (P.n_rp + 3) / 4;
    P.units = B * P.mp * P.groups;
    P.three_pass = three ? 1 : 0;
    P.stages = TC_STAGES;
    P.sched = T.build_sched;
    if (T.build_stages >= 1 && T.build_stages <= TC_STAGES) P.stages = T.build_stages;
    P.scale = fold_scale ? 1.0f : inv_sqrt_d;
    P.probe = T.probe;

    CUtensorMap maps[6];
    if (int e = make_map(&maps[0], a_hi, (long long)B * N, D, TC_BM, TC_BK)) return e;
    if (int e = make_map(&maps[1], three ? a_lo : a_hi, (long long)B * N, D, TC_BM, TC_BK)) return e;
    if (int e = make_map(&maps[2], b_hi, (long long)B * NP, D, P.NT / 2, TC_BK)) return e;
    if (int e = make_map(&maps[3], three ? b_lo : b_hi, (long long)B * NP, D, P.NT / 2, TC_BK)) return e;
    const int box2 = (P.halves == 2 ? P.NT2 : P.NT) / 2;               // second tile of a row pair: its own (narrower) box
    if (int e = make_map(&maps[4], b_hi, (long long)B * NP, D, box2, TC_BK)) return e;
    if (int e = make_map(&maps[5], three ? b_lo : b_hi, (long long)B * NP, D, box2, TC_BK)) return e;

    // store maps (see the epilogue)
    TcStoreMaps SM;
    {
        void* l0 = P.lvl[0];
        cuuint64_t dims[3] = {(cuuint64_t)NP, (cuuint64_t)N, (cuuint64_t)B};
        cuuint64_t str[2] = {(cuuint64_t)NP * es, (cuuint64_t)N * NP * es};
        cuuint32_t box32[3] = {32, 32, 1}, box16[3] = {16, 32, 1};
        // staging rows: fp32 128 / 64 bytes, bf16 64 / 32 bytes (see the epilogue)
        if (int e = make_vol_map(&SM.l0_c32, l0, vb, 3, dims, str, box32, vb ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)) return e;
        if (int e = make_vol_map(&SM.l0_c16, l0, vb, 3, dims, str, box16, vb ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    }

    int e;
    // 4 epilogue warps measured 2.5 % faster than 8 in the three-pass mode at cfg 2 (same box, round robin:
    // profiles/r01j_build_epilogue_warps_ab.jsonl) and equal in single-pass mode; FLOWCORR_BUILD_EPI_WARPS=8 selects the other
    const int ewarps = T.build_epi_warps;
    const int kbs = D / 64;
#define FC_TC_LAUNCH(EWV, VBV)                                                                           \
    (kbs == 1 ? launch_tc<1, EWV, VBV>(maps, SM, P, B, s) : kbs == 2 ? launch_tc<2, EWV, VBV>(maps, SM, P, B, s) \
   : kbs == 3 ? launch_tc<3, EWV, VBV>(maps, SM, P, B, s) : launch_tc<4, EWV, VBV>(maps, SM, P, B, s))
    if (vb) e = FC_TC_LAUNCH(4, 1);                          // bf16 volume: half the store bytes, the 4-warp epilogue
    else if (ewarps == 8) e = FC_TC_LAUNCH(8, 0);
    else e = FC_TC_LAUNCH(4, 0);
#undef FC_TC_LAUNCH
    if (e) return e;
    if (vb) return FC_OK;                                    // (all levels fused: checked above)
    return simt_pool_levels(static_cast<float*>(pyramid), pyr, P.n_fused, s);
}

}  // namespace fc

#ifdef FC_PROBES
extern "C" int fc_debug_build_chunk_trace(unsigned long long* host_out) {    // 8 tiles x 8 chunks x 5 stamps (see TB_CHUNK)
    return cudaMemcpyFromSymbol(host_out, fc::fc_build_chunk_trace, sizeof(fc::fc_build_chunk_trace)) == cudaSuccess ? 0 : 1;
}
extern "C" int fc_debug_build_trace(unsigned long long* host_out) {     // 512 x TB_SLOTS stamps (see TB_TRACE)
    return cudaMemcpyFromSymbol(host_out, fc::fc_build_trace_buf, sizeof(fc::fc_build_trace_buf)) == cudaSuccess ? 0 : 1;
}
#endif
