// Pyramid lookup fused into the motion encoder's first convolution:
//
//   out[b, co, p] = relu(bias[co] + sum_k W[co, k] * lookup(pyramid, coords)[b, k, p])
//
// i.e. `F.relu(self.convc1(corr))` of /root/reference/pytorch/core/update.py:83,90 applied to `corr = corr_fn(coords1)`
// (raft.py:124) without the 324-channel tensor ever reaching HBM (the reference writes 73 MB per iteration at config 2
// and reads them straight back in a 1x1 convolution wrapped in two layout conversions).
//
// Shape of the contraction: per query 324 looked-up values x 256 output channels -- a GEMM with K = 324 whose
// "activation" operand is produced by the lookup consumers.  The WEIGHTS are the stationary operand and live in TENSOR
// MEMORY for the whole (persistent) kernel: tcgen05.mma takes A from TMEM, so M = output channels (256 over a CTA
// pair, cta_group::2, 128 TMEM lanes per CTA), N = queries (64 per pair-tile, 32 from each CTA's lookup tile),
// K = 4 levels x 96 slots (81 taps + padding; slot = 10 * x_offset + y_offset so that a consumer thread stores pairs).
// TMEM columns per CTA: 2 x 64 accumulator + 192 (weights hi) + 192 (weights lo) = 512.  Nothing of the weights
// touches shared memory, which stays with the footprint ring (3 stages) and the double-buffered B operand (2 x 48 KB:
// 32 queries x 384 slots, bf16 hi and lo, K-major SWIZZLE_128B written by the consumers themselves).
// fp32 parity through the same three-pass split as the volume build: W_hi*V_hi + W_lo*V_hi + W_hi*V_lo, fp32 accumulate.
//
// Warps (512 threads): 0-2 = footprint producers (TMA), 3-6 = weights -> TMEM once, then epilogue (TMEM -> + bias -> ReLU
// -> (B, 256, H, W)); lane 0 of warp 3 also allocates TMEM and issues the MMAs (leader CTA) / relays the peer's
// completion, 7-15 = three lookup consumer groups of three warps.
#include "fc_lookup_fwd.cuh"
#include "fc_umma.cuh"

namespace fc {

constexpr int LC_RADIUS = 4, LC_R = 2 * LC_RADIUS + 1, LC_L = 4;
#ifndef FC_LC_STAGES
#define FC_LC_STAGES 3
#endif
#ifndef FC_LC_GROUPS
#define FC_LC_GROUPS 3
#endif
#ifndef FC_LC_BBUFS
#define FC_LC_BBUFS 2
#endif
constexpr int LC_STAGES = FC_LC_STAGES;                   // footprint ring depth (LfShared holds 6 barriers)
constexpr int LC_GROUPS = FC_LC_GROUPS;                   // consumer groups of LF_GWARPS warps
#ifndef FC_LC_PRODUCERS
#define FC_LC_PRODUCERS 3
#endif
constexpr int LC_PRODUCERS = FC_LC_PRODUCERS;             // footprint producer warps: warp 0 and the last LC_PRODUCERS - 1 warps
constexpr int LC_NB = FC_LC_BBUFS;
// a ring stage must always be filled by the same producer and drained by the same consumer group: an mbarrier parity wait
// may run at most one phase behind the barrier it waits on (a waiter two phases late sees a stale "completed")
static_assert(LC_STAGES % LC_GROUPS == 0 && LC_STAGES % LC_PRODUCERS == 0, "stage ownership");                        // B-operand buffers (1: a pair-tile's values wait in registers for the previous MMAs)
// warps: [0, P) footprint producers, [P, P + 4) epilogue (the first of them also issues the MMAs), then the consumer groups
constexpr int LC_EW0 = LC_PRODUCERS;                      // first epilogue warp
constexpr int LC_CW0 = LC_PRODUCERS + 4;                  // first consumer warp
constexpr int LC_THREADS = 32 * (LC_PRODUCERS + 4 + LC_GROUPS * LF_GWARPS);
constexpr int LC_KL = 96;                                 // K slots per level: 9 x 10 = 90 used
constexpr int LC_K = LC_L * LC_KL;                        // 384
constexpr int LC_COUT = 256;
constexpr int LC_QT = 64;                                 // queries per pair-tile
constexpr int LC_BPLANE = (LC_K / 64) * QT * 128;         // 24 KB: [6 k-blocks][32 rows][128 B]
constexpr int LC_BBUF = 2 * LC_BPLANE;                    // hi + lo
constexpr uint32_t LC_D0 = 0, LC_D1 = 64, LC_AHI = 128, LC_ALO = LC_AHI + LC_K / 2;    // TMEM columns
constexpr int LC_WP = LC_K + 8;                           // packed weight row pitch in elements: 784 B, rows 16 B apart modulo 128 B
constexpr int LC_WPLANE_BYTES = 128 * LC_WP * 2;          // one CTA's half of one plane: 100 352 B

struct ConvParams {
    const __nv_bfloat16* w_hi;      // [256][392] slot-ordered (384 slots + 8 pad elements per row), zero at pad slots
    const __nv_bfloat16* w_lo;
    const float* bias;              // [256]
    float* out;                     // (B, 256, N)
    int B, tiles_per_sample, n_pair_tiles;
};

// D[tmem, 256 x N over the pair] (+)= A[TMEM, 128 lanes per CTA] * B[smem, N/2 rows per CTA]^T
__device__ __forceinline__ void umma2_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ uint32_t lc_pack2(float a, float b) {
    const __nv_bfloat162 v(__float2bfloat16_rn(a), __float2bfloat16_rn(b));
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float lc_round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

// A consumer warp's 3 x 9 lookup values of one level -> bf16 hi / lo pairs in the K-major SWIZZLE_128B B operand.
// Slot kp = 96 l + 10 (3 W + aa) + j.  With the warp W and the level's parity LODD as template parameters, kp = 64 m + c with
// m = (96 l) >> 6 a run-time row block and c a compile-time constant: the 128-byte row block, the 16-byte chunk before the
// swizzle and the position inside the chunk are immediates, and the lane's eight swizzled chunk offsets sw[] are computed
// once per kernel (the run-time slot arithmetic was ~8 integer instructions in front of every store).
template <int W, int LODD, int APW, int NR>
__device__ __forceinline__ void lc_store_b(uint32_t hi_row, int m, const uint32_t (&sw)[8], const float (&o)[APW][NR]) {
#pragma unroll
    for (int aa = 0; aa < APW; ++aa) {
#pragma unroll
        for (int jj = 0; jj < (NR + 1) / 2; ++jj) {
            constexpr int dummy = 0; (void)dummy;
            const int c = 32 * LODD + (W * APW + aa) * (NR + 1) + 2 * jj;          // compile-time after unrolling
            const float v0 = o[aa][2 * jj], v1 = (2 * jj + 1 < NR) ? o[aa][2 * jj + 1 < NR ? 2 * jj + 1 : 0] : 0.f;
            const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xffff0000u));
            const uint32_t addr = hi_row + (uint32_t)((m + (c >> 6)) * (QT * 128)) + sw[(c & 63) >> 3] + (uint32_t)((c & 7) * 2);
            asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(addr), "r"(hb) : "memory");
            asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(addr + LC_BPLANE), "r"(*reinterpret_cast<const uint32_t*>(&ll)) : "memory");
        }
    }
}

// the query of lane `lane` in lookup tile (pair-tile T, level l) of CTA `rank`
__device__ __forceinline__ LfQuery lc_query(const LookupParams& P, const ConvParams& C, int T, int level, int rank, int lane) {
    LfQuery q;
    q.level = level;
    q.b = T / C.tiles_per_sample;
    q.p = (T - q.b * C.tiles_per_sample) * LC_QT + rank * QT + lane;
    q.live = q.p < P.N;
    q.gq = q.live ? q.b * P.N + q.p : 0;
    q.cx = 0.f; q.cy = 0.f; q.near_ = false;
    if (q.live) {
        const float* c = P.coords + (long long)q.b * 2 * P.N + q.p;
        q.cx = __ldg(c);                     // raw: scaled by lf_finish_query, so a prefetch does not stall on the load
        q.cy = __ldg(c + P.N);
    }
    return q;
}

// Timeline probe (FC_PROBES builds only; tools/probe_lookup_convc1_trace.py): per CTA, thread 0: 0 kernel start, 1 barriers /
// TMEM ready, 2 weights in tensor memory + B zeroed (main loop starts), 3 main loop done
#ifdef FC_PROBES
__device__ unsigned long long fc_lc_trace_buf[148 * 4];
#define LC_TRACE(k) do { if (threadIdx.x == 0 && blockIdx.x < 148) fc_lc_trace_buf[blockIdx.x * 4 + (k)] = clock64(); } while (0)
// main loop of CTA 0: consumer group 0 / warp 0 per lookup tile k (0 top, 1 interpolated, 2 B buffer free, 3 B written);
// epilogue warp 0 per pair-tile i (4 issue start, 5 issue end, 6 accumulator complete, 7 drained + stored)
__device__ unsigned long long fc_lc_loop_trace[64 * 8];
#define LC_LOOP(idx, k) do { if (blockIdx.x == 0 && lane == 0 && (idx) < 64) fc_lc_loop_trace[(idx) * 8 + (k)] = clock64(); } while (0)
#else
#define LC_TRACE(k) do {} while (0)
#define LC_LOOP(idx, k) do {} while (0)
#endif

__device__ __forceinline__ void lc_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int CM, int VB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LC_THREADS, 1)
lookup_convc1_kernel(const __grid_constant__ LookupMaps M, const LookupParams P, const ConvParams C) {
    constexpr int LF_STAGE_BYTES = lf_stage_bytes(VB);
    // both weight planes fit the staging area at once with the fp32 volume's ring; with the bf16 volume's (half the bytes) they
    // are staged one after the other
    constexpr bool W_BOTH = 2 * LC_WPLANE_BYTES <= LC_NB * LC_BBUF + LC_STAGES * LF_STAGE_BYTES;
    LC_TRACE(0);
    extern __shared__ __align__(1024) uint8_t lc_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lc_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* bbuf = smem;                                               // [LC_NB][hi plane | lo plane]
    const uint32_t win = smem_u32(smem + LC_NB * LC_BBUF);              // [stage][query][window]
    LfShared& sh = *reinterpret_cast<LfShared*>(smem + LC_NB * LC_BBUF + LC_STAGES * LF_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(&sh + 1);
    uint64_t* b_part = bars;             // 2, per CTA: its 4 levels x 3 consumer warps = 12 arrivals (CTA scope: cheap)
    uint64_t* b_peer = bars + 2;         // 2, used in the leader: ONE cluster-scope release per pair-tile from the peer's relay
    uint64_t* b_empty = bars + 4;        // 2, one multicast commit
    uint64_t* t_full = bars + 6;         // 2, one multicast commit
    uint64_t* t_empty = bars + 8;        // 2, used in the leader: 2 CTAs x 4 epilogue warps
    uint64_t* w_full = bars + 10;        // 1: both weight planes of this CTA landed in the staging area
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int n_mine = pair < C.n_pair_tiles ? (C.n_pair_tiles - pair + n_pairs - 1) / n_pairs : 0;    // pair-tiles of this pair

    if (threadIdx.x == 0) {
        for (int i = 0; i < LC_STAGES; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], LF_GWARPS); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_part + i, LC_L * LF_GWARPS); mbar_init(b_peer + i, 1); mbar_init(b_empty + i, 1);
            mbar_init(t_full + i, 1); mbar_init(t_empty + i, 2 * 4);
        }
        mbar_init(w_full, 1);
        mbar_fence_init();
        // this CTA's 128 channels of both weight planes: two bulk copies into the (still idle) B buffers + footprint ring, issued
        // before anything else (the per-thread copy loop + two block barriers they replace were most of a 12 400-cycle prologue)
        static_assert(LC_WPLANE_BYTES <= LC_NB * LC_BBUF + LC_STAGES * lf_stage_bytes(1), "weight staging fits B buffers + ring");
        mbar_expect_tx(w_full, (W_BOTH ? 2u : 1u) * LC_WPLANE_BYTES);
        lc_bulk_load(smem_u32(smem), C.w_hi + (long long)rank * 128 * LC_WP, LC_WPLANE_BYTES, w_full);
        if (W_BOTH) lc_bulk_load(smem_u32(smem) + LC_WPLANE_BYTES, C.w_lo + (long long)rank * 128 * LC_WP, LC_WPLANE_BYTES, w_full);
    }
    if (warp == LC_EW0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    LC_TRACE(1);

    // weights of this CTA's 128 output channels -> tensor memory (lane = channel, 32-bit column = two consecutive slots).
    // A thread needs its channel's whole row (768 B per plane); the packed rows have a pitch of 784 B, so the 32 lanes of a
    // warp read their rows out of the staging area without bank conflicts.
    for (int plane = 0; plane < 2; ++plane) {
        if (!W_BOTH && plane == 1) {
            __syncthreads();                                       // plane 0 is in tensor memory: its staging bytes are free
            if (threadIdx.x == 0) {
                mbar_expect_tx(w_full, (uint32_t)LC_WPLANE_BYTES);
                lc_bulk_load(smem_u32(smem), C.w_lo + (long long)rank * 128 * LC_WP, LC_WPLANE_BYTES, w_full);
            }
        }
        if (warp >= LC_EW0 && warp < LC_CW0) {
            if (plane == 0 || !W_BOTH) mbar_wait(w_full, (uint32_t)(W_BOTH ? 0 : plane));
            const int quarter = warp & 3;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            const uint4* rowp = reinterpret_cast<const uint4*>(smem + (W_BOTH ? plane : 0) * LC_WPLANE_BYTES + (quarter * 32 + lane) * (LC_WP * 2));
            for (int part = 0; part < LC_K / 64; ++part) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 t = rowp[part * 8 + i];
                    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
                }
                tmem_st32(lane_addr + (plane ? LC_ALO : LC_AHI) + (uint32_t)(part * 32), v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        }
    }
    __syncthreads();
    // the B operand starts out as zeros: pad slots (y-offset 9 of every x-offset, slots 90..95 of every level) are never
    // written again and must not hold NaN patterns (their weights are zero)
    for (int i = threadIdx.x; i < LC_NB * LC_BBUF / 16; i += LC_THREADS) reinterpret_cast<uint4*>(bbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // both CTAs: barriers initialised, weights in place
    tc_fence_after();
    LC_TRACE(2);

    if (warp < LC_PRODUCERS) {
        // ================= footprint producers: lookup tiles pw, pw + LC_PRODUCERS, ... =================
        // (one warp issues a tile's 32 footprint loads one by one through the uniform datapath, ~2.5 us per tile:
        // a single producer warp bounded the whole kernel at 120 us; three bring it to the consumers' pace)
        const int n_k = n_mine * LC_L, pw = warp;
        LfQuery q{};
        if (pw < n_k) q = lc_query(P, C, pair + (pw / LC_L) * n_pairs, pw % LC_L, (int)rank, lane);
        for (int k = pw; k < n_k; k += LC_PRODUCERS) {
            LfQuery qn = q;
            const int k1 = k + LC_PRODUCERS;
            if (k1 < n_k) qn = lc_query(P, C, pair + (k1 / LC_L) * n_pairs, k1 % LC_L, (int)rank, lane);   // coordinates one turn ahead
            lf_finish_query(P, q);
            lf_produce<LC_RADIUS, CM, VB>(P, M, sh, win, q, k % LC_STAGES, lane, 0, k >= LC_STAGES,
                                          ((uint32_t)(k / LC_STAGES) & 1u) ^ 1u);
            q = qn;
        }
    } else if (warp < LC_CW0) {
        // ================= epilogue: TMEM -> + bias -> ReLU -> (B, 256, H, W); its first warp also issues the MMAs ======
        const int quarter = warp & 3;
        const int co = (int)rank * 128 + quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float bias = __ldg(C.bias + co);
        const bool vec = (P.N & 3) == 0;
        const uint32_t idesc = umma_idesc_bf16(2 * 128, LC_QT);

        // pair-tile i: (leader) wait for both halves of the B operand and a free accumulator, issue 3 x 24 MMAs;
        // (peer) ONE release at cluster scope per pair-tile tells the leader that this CTA's half is written -- its
        // consumers arrive on a local barrier (a cluster-scope release per consumer warp and lookup tile cost the
        // consumers more than the lookups themselves)
        auto issue = [&](int i) {
            const int bb = i % LC_NB, db = i & 1;
            mbar_wait(b_part + bb, (uint32_t)(i / LC_NB) & 1u);                 // this CTA's half of the B operand
            if (!leader) { mbar_arrive_remote_default(b_peer + bb, 0); return; }
            mbar_wait(b_peer + bb, (uint32_t)(i / LC_NB) & 1u);                 // the peer's half
            mbar_wait_cluster(t_empty + db, ((uint32_t)(i >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (db ? LC_D1 : LC_D0);
            const uint32_t bh = smem_u32(bbuf + bb * LC_BBUF), bl = bh + LC_BPLANE;
#pragma unroll 4
            for (int ks = 0; ks < LC_K / 16; ++ks) {
                const uint32_t boff = (uint32_t)((ks >> 2) * (QT * 128) + (ks & 3) * 32);
                const uint64_t dh = umma_desc_sw128(bh + boff), dl = umma_desc_sw128(bl + boff);
                const uint32_t ah = tmem_base + LC_AHI + (uint32_t)(ks * 8), al = tmem_base + LC_ALO + (uint32_t)(ks * 8);
                umma2_ts_bf16(d_addr, ah, dh, idesc, ks != 0 ? 1u : 0u);
                umma2_ts_bf16(d_addr, al, dh, idesc, 1u);
                umma2_ts_bf16(d_addr, ah, dl, idesc, 1u);
            }
            umma2_commit(b_empty + bb);              // both CTAs: this B buffer may be overwritten
            umma2_commit(t_full + db);               // both CTAs: accumulator complete
        };
        if (warp == LC_EW0 && n_mine > 0) {
            if (elect_one()) issue(0);
            __syncwarp();
        }
        for (int i = 0; i < n_mine; ++i) {
            if (warp == LC_EW0) LC_LOOP(i, 4);
            if (warp == LC_EW0 && i + 1 < n_mine) {      // the next pair-tile's MMAs before this one's epilogue
                if (elect_one()) issue(i + 1);
                __syncwarp();
            }
            if (warp == LC_EW0) LC_LOOP(i, 5);
            const int T = pair + i * n_pairs, db = i & 1;
            const int b = T / C.tiles_per_sample, p0 = (T - b * C.tiles_per_sample) * LC_QT;
            mbar_wait(t_full + db, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            if (warp == LC_EW0) LC_LOOP(i, 6);
            float v[LC_QT];
            tmem_ld32(lane_addr + (db ? LC_D1 : LC_D0), v);
            tmem_ld32(lane_addr + (db ? LC_D1 : LC_D0) + 32u, v + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(t_empty + db, 0);          // the accumulator is in registers
            float* dst = C.out + ((long long)b * LC_COUT + co) * P.N + p0;
            const int n_valid = min(LC_QT, P.N - p0);
            if (vec && n_valid == LC_QT) {
#pragma unroll
                for (int j = 0; j < LC_QT / 4; ++j)
                    reinterpret_cast<float4*>(dst)[j] = make_float4(fmaxf(v[4 * j] + bias, 0.f), fmaxf(v[4 * j + 1] + bias, 0.f),
                                                                    fmaxf(v[4 * j + 2] + bias, 0.f), fmaxf(v[4 * j + 3] + bias, 0.f));
            } else {
#pragma unroll
                for (int j = 0; j < LC_QT; ++j)
                    if (j < n_valid) dst[j] = fmaxf(v[j] + bias, 0.f);
            }
            if (warp == LC_EW0) LC_LOOP(i, 7);
        }
    } else {
        // ================= lookup consumers: interpolate, split, write the B operand =================
        const int cw = warp - LC_CW0, g = cw / LF_GWARPS, w = cw - g * LF_GWARPS;
        constexpr int APW = (LC_R + LF_GWARPS - 1) / LF_GWARPS;              // 3 x-offsets per warp
        const int n_k = n_mine * LC_L;
        uint32_t sw[8];                                                 // this lane's row: chunk ch sits at (ch ^ (lane & 7)) * 16
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) sw[ch] = (uint32_t)((ch ^ (lane & 7)) << 4);
        LfQuery q{};
        if (g < n_k) q = lc_query(P, C, pair + (g / LC_L) * n_pairs, g % LC_L, (int)rank, lane);
        for (int k = g; k < n_k; k += LC_GROUPS) {
            const int i = k / LC_L, l = k - i * LC_L, bb = i % LC_NB;
            if (cw == 0) LC_LOOP(k / LC_GROUPS, 0);
            LfQuery qn = q;
            const int kn = k + LC_GROUPS;
            if (kn < n_k) qn = lc_query(P, C, pair + (kn / LC_L) * n_pairs, kn % LC_L, (int)rank, lane);   // coordinates one turn ahead
            lf_finish_query(P, q);
            RegSink<APW, LC_R> sink;
#pragma unroll
            for (int aa = 0; aa < APW; ++aa)
#pragma unroll
                for (int j = 0; j < LC_R; ++j) sink.o[aa][j] = 0.f;
            lf_consume<LC_RADIUS, CM, false, VB>(P, sh, win, q, k % LC_STAGES, (uint32_t)(k / LC_STAGES) & 1u, lane, w, sink);
            if (cw == 0) LC_LOOP(k / LC_GROUPS, 1);
            // the MMAs that read this buffer LC_NB pair-tiles ago have retired
            if (i >= LC_NB) mbar_wait(b_empty + bb, (uint32_t)(i / LC_NB - 1) & 1u);
            if (cw == 0) LC_LOOP(k / LC_GROUPS, 2);
            const uint32_t hi_row = smem_u32(bbuf + bb * LC_BBUF + lane * 128);
            const int m = (l * LC_KL) >> 6;
            if (l & 1) {
                if (w == 0) lc_store_b<0, 1, APW, LC_R>(hi_row, m, sw, sink.o);
                else if (w == 1) lc_store_b<1, 1, APW, LC_R>(hi_row, m, sw, sink.o);
                else lc_store_b<2, 1, APW, LC_R>(hi_row, m, sw, sink.o);
            } else {
                if (w == 0) lc_store_b<0, 0, APW, LC_R>(hi_row, m, sw, sink.o);
                else if (w == 1) lc_store_b<1, 0, APW, LC_R>(hi_row, m, sw, sink.o);
                else lc_store_b<2, 0, APW, LC_R>(hi_row, m, sw, sink.o);
            }
            fence_proxy_async_smem();                                  // generic-proxy writes -> the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(b_part + bb);
            if (cw == 0) LC_LOOP(k / LC_GROUPS, 3);
            q = qn;
        }
    }

    LC_TRACE(3);
    // neither CTA may leave while its peer can still touch its barriers / tensor memory
    tc_fence_before();
    cluster_sync_all();
    if (warp == LC_EW0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// convc1 weights (256, 324) fp32 -> slot-ordered bf16 hi/lo [256][384 slots + 8 pad] + bias
__global__ void convc1_prepare_kernel(const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                      float* bias_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < LC_COUT * LC_WP) {
        const int co = i / LC_WP, kp = i - co * LC_WP;                   // kp >= LC_K: the row's 8 pad elements
        const int l = kp / LC_KL, s = kp - l * LC_KL, a = s / (LC_R + 1), j = s - a * (LC_R + 1);
        float v = 0.f;
        if (kp < LC_K && a < LC_R && j < LC_R) v = w[(long long)co * (LC_L * LC_R * LC_R) + l * LC_R * LC_R + a * LC_R + j];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    if (i < LC_COUT) bias_out[i] = bias ? bias[i] : 0.f;
}

static size_t convc1_packed_bytes() { return (size_t)LC_COUT * LC_WP * 2 * 2 + (size_t)LC_COUT * 4; }

template <int CM>
static int launch_lc(const LookupMaps& M, const LookupParams& P, const ConvParams& C, int vb, cudaStream_t s) {
    const int n_sm = sm_count_cached();
    int n_pairs = n_sm / 2;
    if (n_pairs > C.n_pair_tiles) n_pairs = C.n_pair_tiles;
    if (vb) {
        const size_t smem = 1024 + LC_NB * LC_BBUF + (size_t)LC_STAGES * lf_stage_bytes(1) + sizeof(LfShared) + 128;
        FC_SMEM_ATTR_ONCE((lookup_convc1_kernel<CM, 1>), smem);
        lookup_convc1_kernel<CM, 1><<<2 * n_pairs, LC_THREADS, smem, s>>>(M, P, C);
    } else {
        const size_t smem = 1024 + LC_NB * LC_BBUF + (size_t)LC_STAGES * lf_stage_bytes(0) + sizeof(LfShared) + 128;
        FC_SMEM_ATTR_ONCE((lookup_convc1_kernel<CM, 0>), smem);
        lookup_convc1_kernel<CM, 0><<<2 * n_pairs, LC_THREADS, smem, s>>>(M, P, C);
    }
    FC_LAUNCH_CHECK("lookup_convc1_kernel");
    return FC_OK;
}

}  // namespace fc

using namespace fc;

extern "C" size_t fc_convc1_weights_bytes(void) { return convc1_packed_bytes(); }

extern "C" int fc_lookup_convc1_supported(int num_levels, int radius, int out_channels) {
    return (num_levels == LC_L && radius == LC_RADIUS && out_channels == LC_COUT) ? 1 : 0;
}

extern "C" int fc_convc1_prepare(const float* weight, const float* bias, void* packed, size_t packed_bytes, void* stream) {
    FC_REQUIRE(weight && packed, "fc_convc1_prepare: null pointer");
    FC_REQUIRE(packed_bytes >= convc1_packed_bytes(), "fc_convc1_prepare: buffer %zu < %zu bytes", packed_bytes, convc1_packed_bytes());
    __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(packed);
    __nv_bfloat16* lo = hi + (size_t)LC_COUT * LC_WP;
    float* b = reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + (size_t)LC_COUT * LC_WP * 4);
    const int n = LC_COUT * LC_WP;
    convc1_prepare_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, bias, hi, lo, b);
    FC_LAUNCH_CHECK("convc1_prepare_kernel");
    return FC_OK;
}

extern "C" int fc_lookup_convc1_fwd(const void* pyramid, const float* coords, const void* packed_weights, float* out,
                                    int B, int H, int W, int num_levels, int radius, int vol_dtype, int coord_mode, void* stream) {
    FC_REQUIRE(pyramid && coords && packed_weights && out, "fc_lookup_convc1_fwd: null pointer");
    FC_REQUIRE(num_levels == LC_L && radius == LC_RADIUS, "fc_lookup_convc1_fwd: built for num_levels = 4, radius = 4 (got %d, %d)",
               num_levels, radius);
    FC_REQUIRE(vol_dtype == FC_VOL_F32 || vol_dtype == FC_VOL_BF16, "fc_lookup_convc1_fwd: unknown vol_dtype %d", vol_dtype);
    FC_REQUIRE(coord_mode == FC_COORD_CUDA || coord_mode == FC_COORD_CPU, "fc_lookup_convc1_fwd: bad coord_mode %d", coord_mode);
    const int vb = vol_dtype == FC_VOL_BF16 ? 1 : 0;
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_lookup_convc1_fwd: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    if (int e = check_lookup_common(pyr, radius, coord_mode)) return e;
    FC_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
               (reinterpret_cast<uintptr_t>(packed_weights) & 15u) == 0, "fc_lookup_convc1_fwd: pointers must be 16-byte aligned");
    LookupParams P{};
    fill_params(P, pyr, radius);
    P.pyr = static_cast<const float*>(pyramid);
    P.coords = coords; P.io = nullptr; P.gpyr = nullptr;
    LookupMaps M;
    if (int e = get_level_maps(M, pyramid, pyr, H, W, vb)) return e;
    ConvParams C{};
    const uint8_t* pw = static_cast<const uint8_t*>(packed_weights);
    C.w_hi = reinterpret_cast<const __nv_bfloat16*>(pw);
    C.w_lo = C.w_hi + (size_t)LC_COUT * LC_WP;
    C.bias = reinterpret_cast<const float*>(pw + (size_t)LC_COUT * LC_WP * 4);
    C.out = out; C.B = B;
    C.tiles_per_sample = (pyr.N + LC_QT - 1) / LC_QT;
    C.n_pair_tiles = B * C.tiles_per_sample;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return coord_mode == FC_COORD_CUDA ? launch_lc<FC_COORD_CUDA>(M, P, C, vb, s) : launch_lc<FC_COORD_CPU>(M, P, C, vb, s);
}

#ifdef FC_PROBES
extern "C" int fc_debug_lookup_convc1_loop_trace(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, fc::fc_lc_loop_trace, sizeof(fc::fc_lc_loop_trace)) == cudaSuccess ? 0 : 1;
}
extern "C" int fc_debug_lookup_convc1_trace(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, fc::fc_lc_trace_buf, sizeof(fc::fc_lc_trace_buf)) == cudaSuccess ? 0 : 1;
}
#endif
