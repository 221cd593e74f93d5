// Pyramid lookup BACKWARD (autograd of CorrBlock.__call__, corr.py:29-50; the forward lives in
// fc_lookup_fwd.cu).  HBM-bound scatter into the block's ONE gradient pyramid.
//
// Mirror image of the forward: persistent CTAs walk tiles = (32 consecutive queries) x (one
// level).  A group of three warps (lane <-> query, warp <-> a third of the window columns)
// builds every query's (2r+2)^2 gradient window in shared memory in GATHER form --
//     W[i][j] = wy0[i] (wx0[j] g[j][i] + wx1[j-1] g[j-1][i]) + wy1[i-1] (wx0[j] g[j][i-1] + wx1[j-1] g[j-1][i-1])
// with g[a][b] the output gradient of window tap (x-offset a, y-offset b) -- laid out exactly as
// the TMA box of the forward ({2|3 patches, 5|6 row pairs} of the 2x8-patch layout), and ONE
// `cp.reduce.async.bulk.tensor` per query adds the box into the gradient pyramid: the adds
// happen in L2 at sector granularity, nothing goes through the SM's REDG path (the previous
// kernel issued ~30 red.global.add.v4 per query-level and was bound by that issue rate), and
// parts of the box outside the padded map are dropped by the TMA unit.  Taps on pad columns /
// the pad row are masked here, so the pads of the gradient pyramid stay zero.
// Every (query, level) map is touched by exactly one lane per launch; launches of one
// CorrBlock are stream-ordered, so the only concurrency is inside L2's adder.
//
// Lattice coordinates (GRU iteration 0) can flip the floor of single taps
// (fc::axis_tap): such queries take a per-tap path (warp 0 of the group, plain shared-memory
// read-modify-write on the lane's private window).
#include "fc_lookup.cuh"

namespace fc {

#ifndef FC_LB_GROUPS
#define FC_LB_GROUPS 3
#endif
#ifndef FC_LB_STAGES
#define FC_LB_STAGES 1
#endif
constexpr int LB_GROUPS = FC_LB_GROUPS;                       // independent warp groups per CTA
constexpr int LB_GWARPS = 3;                                  // warps per group (window column thirds)
constexpr int LB_THREADS = 32 * LB_GROUPS * LB_GWARPS;        // 288
constexpr int LB_STAGES = FC_LB_STAGES;                       // window buffers per group
constexpr int LB_WIN_BYTES = 6 * 3 * 64;                      // 6 row pairs x 3 patches x 64 B = 1152
constexpr int LB_STAGE_BYTES = QT * LB_WIN_BYTES;             // 36 864
// the output gradients of a tile ((2r+1)^2 channels x 32 queries) arrive as ONE TMA box per tile, one tile ahead: issued as
// 45 predicated loads per lane they took a third of a tile's time (2200 of 6600 cycles: the load/store unit's queue of
// outstanding requests, not DRAM, set the pace -- tools/probe_lookup_bwd_trace.py)
constexpr int LB_GBOX_FLOATS = 81 * QT;                       // room for radius 4: [(2r+1)^2 channels][32 queries]
constexpr int LB_GBOX_BYTES = (LB_GBOX_FLOATS * 4 + 127) / 128 * 128;

__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(g + 1), "n"(32 * LB_GWARPS) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32b(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// Raw per-lane inputs of one tile: coordinates and the output gradients of the taps this warp
// needs.  Loaded one tile AHEAD (nothing here depends on the coordinates' values).
// The three warps of a group run the SAME code on their own third of the window columns (c_lo is a run-time value):
// column-specialised copies of the loop made the kernel 4160 instructions = 66 KB, and ncu showed 27 % of the stall
// samples waiting for instruction fetch with a time that did not move when the warps per SM were doubled.
template <int RADIUS>
struct LbTile {
    static constexpr int R = 2 * RADIUS + 1;
    static constexpr int NCOL = R + 1;
    static constexpr int CPW = (NCOL + LB_GWARPS - 1) / LB_GWARPS;      // window columns per warp (4 for r = 4)
    static constexpr int NT = CPW + 1;                                  // x-taps a = c_lo - 1 + k, k < NT
    int level, gq;
    bool live;
    float cx, cy;                                                        // raw (unscaled) coordinates
    float gv[NT][R];
    const float* gptr;

    // coordinates + bookkeeping only (the gradients come from the staged box, see take())
    __device__ __forceinline__ void head(const LookupParams& P, const TileIt& it, int lane) {
        level = it.level(P.L);
        gq = it.qt * QT + lane;
        live = gq < P.Q;
        cx = 0.f; cy = 0.f;
        int b = 0, p = 0;
        if (live) {
            split_query(P, gq, b, p);
            const float* c = P.coords + (long long)b * 2 * P.N + p;
            cx = __ldg(c);
            cy = __ldg(c + P.N);
        }
        gptr = P.io + ((long long)b * P.K + level * R * R) * P.N + p;
    }
    // this warp's taps out of the staged box [channel a * R + i][query]
    __device__ __forceinline__ void take(uint32_t box, int lane, int c_lo) {
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const int a = c_lo - 1 + k;
            const bool ok = live && a >= 0 && a < R;
            const uint32_t src = box + 4u * (uint32_t)((ok ? a * R : 0) * QT + lane);
#pragma unroll
            for (int i = 0; i < R; ++i) gv[k][i] = ok ? lds_f32b(src + 4u * (uint32_t)(i * QT)) : 0.f;
        }
    }
    // the same taps straight from global memory (tiles whose 32 queries straddle two samples, unaligned tensors)
    __device__ __forceinline__ void direct(const LookupParams& P, int c_lo) {
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const int a = c_lo - 1 + k;
            const bool ok = live && a >= 0 && a < R;
            const float* g = gptr + (long long)(ok ? a * R : 0) * P.N;
#pragma unroll
            for (int i = 0; i < R; ++i) gv[k][i] = ok ? __ldg(g + (long long)i * P.N) : 0.f;
        }
    }
};

// Timeline probe (FC_PROBES builds only; tools/probe_lookup_bwd_trace.py): warp 0 of group 0 of CTA 0 stamps clock64() per tile:
// 0 loop top, 1 taps done, 2 previous reduce-adds read + first group barrier, 3 window zeroed + barrier, 4 window written,
// 5 fence + barrier, 6 reduce-adds issued
#ifdef FC_PROBES
__device__ unsigned long long fc_lb_trace_buf[64 * 8];
#define LB_TRACE(it, k) do { if (blockIdx.x == 0 && g == 0 && WI == 0 && lane == 0 && (it) < 64) fc_lb_trace_buf[(it) * 8 + (k)] = clock64(); } while (0)
#else
#define LB_TRACE(it, k) do {} while (0)
#endif

template <int RADIUS, int CM>
__device__ __forceinline__ void lb_warp(const LookupMaps& M, const CUtensorMap* gmap, const LookupParams& P, int n_tiles, int g,
                                        int WI, int lane, uint32_t gbase, uint32_t gbox, uint64_t* gfull, volatile uint8_t* reg_flags) {
    using T = LbTile<RADIUS>;
    constexpr int R = T::R;
    const int C_LO = WI * T::CPW;                                        // this warp's window columns [C_LO, C_LO + CPW)
    const int tg = WI * 32 + lane;                                       // thread index inside the group
    const int n_groups = gridDim.x * LB_GROUPS;
    int tile = blockIdx.x * LB_GROUPS + g;
    if (tile >= n_tiles) return;                                         // whole group leaves together
    // tile = qt * L + slot, level = (slot + qt) % L: a group's tiles rotate through the levels (their cost differs)
    const int L = P.L, hop_q = n_groups / L, hop_l = n_groups - hop_q * L, hop_qm = hop_q % L;
    TileIt ti;
    ti.qt = tile / L; ti.slot = tile - ti.qt * L; ti.qm = ti.qt % L;
    // a tile's gradients come as one TMA box [(2r+1)^2 channels][32 queries] when its queries lie in one sample
    auto boxable = [&](int qt, int& b0, int& p0) {
        if (gmap == nullptr) return false;
        split_query(P, qt * QT, b0, p0);
        return qt * QT < P.Q && p0 + QT <= P.N;
    };
    auto issue_box = [&](const TileIt& t, int b0, int p0, int buf) {
        if (WI == 0 && lane == 0) {
            mbar_expect_tx(gfull + buf, (uint32_t)(R * R * QT * 4));
            tma_load_3d(gbox + (uint32_t)(buf * LB_GBOX_BYTES), gmap, smem_u32(gfull + buf), p0, t.level(P.L) * R * R, b0);
        }
    };
    T cur;
    cur.head(P, ti, lane);
    int b0 = 0, p0 = 0;
    bool cur_box = boxable(ti.qt, b0, p0);
    if (cur_box) issue_box(ti, b0, p0, 0);
    uint32_t uses[2] = {0u, 0u};                                           // boxable tiles each buffer has served (barrier parity)
    for (int it = 0;; ++it) {
        const int next = tile + n_groups;
        // next tile: coordinates by two loads, gradients by a TMA box into the other buffer (all warps read the tile before
        // last out of it before that tile's first group barrier)
        T nxt = cur;
        bool nxt_box = false;
        if (next < n_tiles) {
            ti.advance(hop_q, hop_l, hop_qm, L);
            nxt.head(P, ti, lane);
            nxt_box = boxable(ti.qt, b0, p0);
            if (nxt_box) issue_box(ti, b0, p0, (it + 1) & 1);
        }
        if (cur_box) {
            const int buf = it & 1;
            mbar_wait(gfull + buf, uses[buf] & 1u);
            ++uses[buf];
            cur.take(gbox + (uint32_t)(buf * LB_GBOX_BYTES), lane, C_LO);
        } else {
            cur.direct(P, C_LO);
        }

        LB_TRACE(it, 0);
        const int level = cur.level;
        const float cx = __fmul_rn(cur.cx, P.inv_scale[level]), cy = __fmul_rn(cur.cy, P.inv_scale[level]);
        const bool near_ = cur.live && (fabsf(cx) < 1048576.f) && (fabsf(cy) < 1048576.f);

        // ---- tap arithmetic (bit-exact integer part, shared with the forward)
        int y0[R]; float wy0[R], wy1[R];
#pragma unroll
        for (int j = 0; j < R; ++j) axis_tap<CM>(cy, j - RADIUS, P.ay[level], y0[j], wy0[j], wy1[j]);
        // x: this warp's taps a = C_LO - 1 + k (weights 0 outside [0, R)), plus the first and the last tap for the box
        int x0o[T::NT]; float wx0o[T::NT], wx1o[T::NT];
        int x_first, x_last;
        { float d0, d1; axis_tap<CM>(cx, -RADIUS, P.ax[level], x_first, d0, d1); axis_tap<CM>(cx, RADIUS, P.ax[level], x_last, d0, d1); }
        bool regular = true;
#pragma unroll
        for (int k = 0; k < T::NT; ++k) {
            const int a = C_LO - 1 + k;
            axis_tap<CM>(cx, a - RADIUS, P.ax[level], x0o[k], wx0o[k], wx1o[k]);
            if (a < 0 || a >= R) { wx0o[k] = 0.f; wx1o[k] = 0.f; }
            else regular = regular && (x0o[k] == x_first + a);
        }
#pragma unroll
        for (int j = 1; j < R; ++j) regular = regular && (y0[j] == y0[0] + j);
        // every x-tap is checked by one of the three warps: the query is regular if all of them say so
        reg_flags[tg] = regular ? 1 : 0;

        // footprint box as the forward sizes it, but clamped to non-negative tensor coordinates:
        // TMA loads zero-fill at negative coordinates, TMA stores / reductions TRAP there
        // (tools/probes/tma_reduce_probe.cu), while boxes overhanging the far edge are clipped
        // ... and, like the forward's, clipped to the padded map at the far edge as well
        const int rp_lo = max(y0[0] >> 1, 0), pc_lo = max(x_first >> 3, 0);
        const int n_rp = min(min((y0[R - 1] + 1) >> 1, P.nrp[level] - 1) - rp_lo + 1, 6);
        const int n_pc = min(min((x_last + 1) >> 3, P.npc[level] - 1) - pc_lo + 1, 3);
        const int sel = (n_rp > 0 && n_pc > 0) ? lk_shape(n_rp, n_pc) : 0;
        const int pitch = 64 * n_pc;
        const int ybase = 2 * rp_lo, xbase = 8 * pc_lo;
        const int Hl = P.H[level], Wl = P.W[level];
        // some part of the window lies inside the map
        const bool touches = near_ && n_rp > 0 && n_pc > 0 && ybase < Hl && xbase < Wl;

        LB_TRACE(it, 1);
        // ---- the buffer used two tiles ago must have been read by its reduce-adds
        const uint32_t stage = gbase + (uint32_t)((it % LB_STAGES) * LB_STAGE_BYTES);
        tma_wait_group_read<LB_STAGES - 1>();                            // every lane waits for its own reduce-adds
        group_sync(g);
        LB_TRACE(it, 2);
        regular = reg_flags[lane] && reg_flags[32 + lane] && reg_flags[64 + lane];
#pragma unroll
        for (int i = 0; i < LB_STAGE_BYTES / 16 / (32 * LB_GWARPS); ++i)
            asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(stage + 16u * (uint32_t)(tg + i * 32 * LB_GWARPS)), "f"(0.f) : "memory");
        group_sync(g);

        LB_TRACE(it, 3);
        const uint32_t wq = stage + (uint32_t)(lane * LB_WIN_BYTES);
        if (touches && regular) {
            // window row n (relative to ybase) sits at (n >> 1) * pitch + (n & 1) * 32 bytes
            const int n0 = y0[0] - ybase;                                  // 0 or 1; negative above the map
#pragma unroll
            for (int k = 0; k < T::CPW; ++k) {
                const int j = C_LO + k;                                    // window column
                if (j >= T::NCOL) break;
                const int x = x_first + j, xr = x - xbase;                 // 0 <= xr <= 16 when x >= 0
                const bool xin = (x >= 0) && (x < Wl);
                const uint32_t col = wq + 4u * (uint32_t)(xr + (xr & ~7));
                // horizontal weights: own tap a = j (left corner, index k + 1), tap a = j - 1 (right corner, index k);
                // taps outside [0, R) carry zero weights and zero gradients
                const float wl = wx0o[k + 1];
                const float wr = wx1o[k];
                uint32_t rofs = (uint32_t)((n0 >> 1) * pitch + (n0 & 1) * 32);     // wraps for n0 < 0: never stored
                uint32_t step = (n0 & 1) ? (uint32_t)pitch - 32u : 32u;
                float hprev = 0.f;
#pragma unroll
                for (int i = 0; i <= R; ++i) {
                    float h = 0.f, wt = 0.f, wb = 0.f;
                    if (i < R) { h = fmaf(wl, cur.gv[k + 1][i < R ? i : 0], wr * cur.gv[k][i < R ? i : 0]); wt = wy0[i < R ? i : 0]; }
                    if (i > 0) wb = wy1[i > 0 ? i - 1 : 0];
                    const float v = fmaf(wt, h, wb * hprev);
                    const int y = y0[0] + i;
                    if (xin && y >= 0 && y < Hl) sts_f32(col + rofs, v);
                    hprev = h;
                    rofs += step;
                    step = (uint32_t)pitch - step;
                }
            }
        } else if (touches && WI == 0) {
            // floor flips among the taps: every tap splats its four corners on its own
            for (int a = 0; a < R; ++a) {
                int xa; float wxa0, wxa1;
                axis_tap<CM>(cx, a - RADIUS, P.ax[level], xa, wxa0, wxa1);
                for (int j = 0; j < R; ++j) {
                    int ya; float wya0, wya1;
                    axis_tap<CM>(cy, j - RADIUS, P.ay[level], ya, wya0, wya1);
                    const float gg = __ldg(cur.gptr + (long long)(a * R + j) * P.N);
#pragma unroll
                    for (int cyy = 0; cyy < 2; ++cyy)
#pragma unroll
                        for (int cxx = 0; cxx < 2; ++cxx) {
                            const int x = xa + cxx, y = ya + cyy;
                            const int xr = x - xbase, n = y - ybase;
                            if (x >= 0 && x < Wl && y >= 0 && y < Hl && xr >= 0 && xr < 24 && n >= 0 && n < 12) {
                                const uint32_t ad = wq + 4u * (uint32_t)(xr + (xr & ~7)) + (uint32_t)((n >> 1) * pitch + (n & 1) * 32);
                                const float wgt = (cyy ? wya1 : wya0) * (cxx ? wxa1 : wxa0);
                                sts_f32(ad, fmaf(gg, wgt, lds_f32b(ad)));
                            }
                        }
                }
            }
        }
        LB_TRACE(it, 4);
        fence_proxy_async_smem();
        group_sync(g);
        LB_TRACE(it, 5);
        // every warp holds every query's box geometry: the 32 reduce-adds of the tile (one per lane, serialised
        // by the hardware's uniform-operand issue) are dealt out over the group's three warps
        if (touches && FC_PROBE_VAL(P) != 1 && (lane % LB_GWARPS) == WI) {
            if (FC_PROBE_VAL(P) == 2) tma_store_3d(&M.m[level][sel], wq, 16 * pc_lo, rp_lo, cur.gq);
            else tma_reduce_add_3d(&M.m[level][FC_PROBE_VAL(P) == 3 ? lk_shape(1, 1) : sel], wq, 16 * pc_lo, rp_lo, cur.gq);   // probe 3: one box shape
        }
        tma_commit_group();
        LB_TRACE(it, 6);
        if (next >= n_tiles) break;
        tile = next;
        cur.level = nxt.level; cur.gq = nxt.gq; cur.live = nxt.live; cur.cx = nxt.cx; cur.cy = nxt.cy; cur.gptr = nxt.gptr;
        cur_box = nxt_box;
    }
    tma_wait_group<0>();
}

template <int RADIUS, int CM>
__global__ void __launch_bounds__(LB_THREADS, 1)
lookup_bwd_kernel(const __grid_constant__ LookupMaps M, const __grid_constant__ CUtensorMap gmap, const LookupParams P, int n_tiles,
                  int staged) {
    extern __shared__ __align__(1024) uint8_t lb_smem[];
    __shared__ __align__(8) uint64_t gfull[LB_GROUPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp / LB_GWARPS, w = warp - g * LB_GWARPS;
    const uint32_t gbase = smem_u32(lb_smem) + (uint32_t)(g * LB_STAGES * LB_STAGE_BYTES);
    const uint32_t gbox = smem_u32(lb_smem) + (uint32_t)(LB_GROUPS * LB_STAGES * LB_STAGE_BYTES + g * 2 * LB_GBOX_BYTES);
    if (threadIdx.x == 0) {
        for (int i = 0; i < LB_GROUPS; ++i) { mbar_init(&gfull[i][0], 1); mbar_init(&gfull[i][1], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    __shared__ uint8_t reg_flags[LB_GROUPS][32 * LB_GWARPS];
    lb_warp<RADIUS, CM>(M, staged ? &gmap : nullptr, P, n_tiles, g, w, lane, gbase, gbox, gfull[g], reg_flags[g]);
}

template <int RADIUS>
static int launch_bwd(const LookupMaps& M, const CUtensorMap& gmap, int staged, const LookupParams& P, int n_tiles, int n_sm, int coord_mode,
                      cudaStream_t s) {
    const size_t smem = (size_t)LB_GROUPS * (LB_STAGES * LB_STAGE_BYTES + 2 * LB_GBOX_BYTES);
    const int want = (n_tiles + LB_GROUPS - 1) / LB_GROUPS;
    const int grid = want < n_sm ? want : n_sm;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(LB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = tunables().pdl ? 1 : 0;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (coord_mode == FC_COORD_CUDA) {
        FC_SMEM_ATTR_ONCE((lookup_bwd_kernel<RADIUS, FC_COORD_CUDA>), smem);
        FC_CUDA(cudaLaunchKernelEx(&cfg, lookup_bwd_kernel<RADIUS, FC_COORD_CUDA>, M, gmap, P, n_tiles, staged));
    } else {
        FC_SMEM_ATTR_ONCE((lookup_bwd_kernel<RADIUS, FC_COORD_CPU>), smem);
        FC_CUDA(cudaLaunchKernelEx(&cfg, lookup_bwd_kernel<RADIUS, FC_COORD_CPU>, M, gmap, P, n_tiles, staged));
    }
    FC_LAUNCH_CHECK("lookup_bwd_kernel");
    return FC_OK;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_lookup_bwd(const float* grad_out, const float* coords, float* grad_pyramid,
                             int B, int H, int W, int num_levels, int radius,
                             int coord_mode, void* stream) {
    FC_REQUIRE(grad_out && coords && grad_pyramid, "fc_lookup_bwd: null pointer");
    FC_REQUIRE((reinterpret_cast<uintptr_t>(grad_pyramid) & 15u) == 0, "fc_lookup_bwd: grad_pyramid must be 16-byte aligned");
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_lookup_bwd: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    FC_REQUIRE(coord_mode == FC_COORD_CUDA || coord_mode == FC_COORD_CPU,
               "fc_lookup_bwd: coord_mode %d has no backward (AlternateCorrBlock is inference-only, corr.py:86)", coord_mode);
    if (int e = check_lookup_common(pyr, radius, coord_mode)) return e;
    LookupParams P{};
    fill_params(P, pyr, radius);
    P.pyr = nullptr; P.coords = coords;
    P.io = const_cast<float*>(grad_out); P.gpyr = grad_pyramid;
    LookupMaps M;
    if (int e = get_level_maps(M, grad_pyramid, pyr, H, W, 0)) return e;
    int n_sm = 0;
    if (int e = sm_count(n_sm)) return e;
    const int n_tiles = ((P.Q + QT - 1) / QT) * pyr.L;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the output gradient as {N, K, B} with boxes of 32 queries x (2r+1)^2 channels (needs 16-byte aligned rows)
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    const int RR = (2 * radius + 1) * (2 * radius + 1);
    int staged = (pyr.N % 4 == 0) && (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0 && pyr.N >= QT;
    if (staged) {
        const cuuint64_t dims[3] = {(cuuint64_t)pyr.N, (cuuint64_t)P.K, (cuuint64_t)B};
        const cuuint64_t str[2] = {(cuuint64_t)pyr.N * 4, (cuuint64_t)pyr.N * P.K * 4};
        const cuuint32_t box[3] = {(cuuint32_t)QT, (cuuint32_t)RR, 1};
        if (int e = encode_tiled_cached(&gmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, grad_out, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE)) return e;
    }
    switch (radius) {
        case 1: return launch_bwd<1>(M, gmap, staged, P, n_tiles, n_sm, coord_mode, s);
        case 2: return launch_bwd<2>(M, gmap, staged, P, n_tiles, n_sm, coord_mode, s);
        case 3: return launch_bwd<3>(M, gmap, staged, P, n_tiles, n_sm, coord_mode, s);
        default: return launch_bwd<4>(M, gmap, staged, P, n_tiles, n_sm, coord_mode, s);
    }
}

#ifdef FC_PROBES
extern "C" int fc_debug_lookup_bwd_trace(unsigned long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, fc::fc_lb_trace_buf, sizeof(fc::fc_lb_trace_buf)) == cudaSuccess ? 0 : 1;
}
#endif
