// Pyramid lookup FORWARD (CorrBlock.__call__, corr.py:29-50 + utils.py:57-65 + ATen
// grid_sample(bilinear, zeros, align_corners=True)): the per-GRU-iteration HBM-bound gather.
//
// Persistent CTAs (one per SM) walk over tiles = (32 consecutive queries) x (one pyramid level);
// the level of a tile rotates with its query block so that every CTA sees all levels.
// A query's footprint is ONE TMA tensor load: the level is described to the TMA unit as a
// 3-D tensor [query][row pair][2*Wp floats] over the 2x8-patch layout (include/flowcorr.h),
// and a box of {2|3 patches, 5|6 row pairs, 1 query} at signed coordinates lands the (2r+2)^2
// footprint (+1 guard row/column for floor flips of the normalise/un-normalise round trip) in
// shared memory.  Everything outside the padded map is zero-filled by the TMA unit itself --
// that IS the reference's padding_mode='zeros'; pad rows/columns inside the map hold zeros by the
// pyramid invariant.  Four box shapes per level follow the patches the footprint touches.
// (Measured, profiles/r02a: the TMA unit requests every byte of a box from L2, also the part it then
// zero-fills -- 168.4 MB per launch = exactly the unclipped boxes, 117.6 MB when the boxes are clipped
// to the maps in software -- but L2 fills from DRAM in whole 128-byte lines either way: 166 MB read per
// launch with and without clipping, and the clipped variant's extra masking made the kernel ~2 % slower.)
//
// A 6-stage mbarrier ring decouples three producer warps (tile k -> producer k % 3: coordinate
// prefetch, footprint arithmetic, one TMA issue per lane) from three consumer groups of
// three warps (tile k -> group k % 3).  Interpolation: lane <-> query, warp <-> three
// x-offsets, so every store of the (B, K, H, W) output is a coalesced
// 128-byte row; one horizontally interpolated column of R+1 rows serves all R outputs of an
// x-offset.  The integer part (fc::axis_tap) is bit-exact to the reference's op sequence.
#include <cstring>
#include <mutex>

#include "fc_lookup.cuh"
#include "fc_tma.cuh"

namespace fc {

constexpr int LF_PTEAMS = 3;                                  // producer teams (tile k -> team k % 3)
constexpr int LF_PSPLIT = 1;                                  // warps per team: each issues 32 / PSPLIT of a tile's loads
constexpr int LF_PRODUCERS = LF_PTEAMS * LF_PSPLIT;           // warps issuing TMA loads
constexpr int LF_GROUPS = 3;                                  // consumer groups (tile k -> group k % 3)
constexpr int LF_GWARPS = 3;                                  // warps per group (x-offset thirds)
constexpr int LF_THREADS = 32 * (LF_PRODUCERS + LF_GROUPS * LF_GWARPS);   // 352
#ifndef FC_LF_STAGES
#define FC_LF_STAGES 6
#endif
constexpr int LF_STAGES = FC_LF_STAGES;
// a ring stage must always be filled by the same producer and drained by the same group:
// an mbarrier parity wait may run at most one phase ahead of the barrier
static_assert(LF_STAGES % LF_PTEAMS == 0 && LF_STAGES % LF_GROUPS == 0, "stage ownership");
// VB = 0: fp32 volume, VB = 1: bf16 volume (same element indexing, half the bytes)
__host__ __device__ constexpr int lf_es(int vb) { return vb ? 2 : 4; }                       // element size
// window = 6 row pairs x 3 patches x 16 elements, rounded up to the 128-byte alignment of a TMA destination
__host__ __device__ constexpr int lf_win_bytes(int vb) { return (6 * 3 * 16 * lf_es(vb) + 127) / 128 * 128; }   // 1152 / 640
__host__ __device__ constexpr int lf_stage_bytes(int vb) { return QT * lf_win_bytes(vb); }   // 36 864 / 20 480 B per stage

// The footprint box of one query on one level, clipped to the padded map.  Producer and consumers derive it from the
// same coordinates with the same arithmetic (fc::axis_tap is deterministic), so nothing but the footprints themselves
// travels through shared memory (an earlier version handed a descriptor over per lane: 14 racecheck hazards, waived
// by a release/acquire argument -- now there is no such write).
struct LfBox {
    int ybase;                            // first window row    (2 * first row pair)
    int xbase;                            // first window column (8 * first patch)
    int n_rp, n_pc;                       // row pairs (<= 6) / patches (<= 3) in the box; n_rp == 0: nothing inside the map
};

struct LfShared {
    uint64_t full[LF_STAGES];
    uint64_t empty[LF_STAGES];
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}
// one volume element from shared memory as fp32 (bf16 -> fp32 is a 16-bit shift)
template <int VB>
__device__ __forceinline__ float lds_vol(uint32_t addr) {
    if (VB) {
        uint16_t h;
        asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(h) : "r"(addr));
        return __uint_as_float((uint32_t)h << 16);
    }
    return lds_f32(addr);
}

struct LfQuery { int level, gq, b, p; bool live, near_; float cx, cy; };

__device__ __forceinline__ LfQuery lf_load_query(const LookupParams& P, const TileIt& it, int lane) {
    LfQuery q;
    q.level = it.level(P.L);
    q.gq = it.qt * QT + lane;
    q.live = q.gq < P.Q;
    q.cx = 0.f; q.cy = 0.f; q.b = 0; q.p = 0;
    if (q.live) {
        split_query(P, q.gq, q.b, q.p);
        const float* c = P.coords + (long long)q.b * 2 * P.N + q.p;
        q.cx = __ldg(c);                                             // raw: scaled by lf_finish_query,
        q.cy = __ldg(c + P.N);                                       // so a prefetch does not stall on the load
    }
    q.near_ = false;
    return q;
}

__device__ __forceinline__ void lf_finish_query(const LookupParams& P, LfQuery& q) {
    q.cx = __fmul_rn(q.cx, P.inv_scale[q.level]);
    q.cy = __fmul_rn(q.cy, P.inv_scale[q.level]);
    // beyond 2^20 every tap is out of bounds for any map this library accepts and the
    // +-1 flip bound used to size the window no longer holds; NaN compares false.
    q.near_ = q.live && (fabsf(q.cx) < 1048576.f) && (fabsf(q.cy) < 1048576.f);
}

// xl / xh, yl / yh: floor indices of the first and last tap on each axis (taps are monotone, span <= R + 1)
__device__ __forceinline__ LfBox lf_box(bool near_, int xl, int xh, int yl, int yh) {
    LfBox bx{0, 0, 0, 0};
    if (near_) {
        const int rp0 = yl >> 1, pc0 = xl >> 3;                       // arithmetic shifts: floor
        const int n_rp = ((yh + 1) >> 1) - rp0 + 1;                   // <= 6
        const int n_pc = ((xh + 1) >> 3) - pc0 + 1;                   // <= 3
        bx.ybase = 2 * rp0; bx.xbase = 8 * pc0; bx.n_rp = n_rp > 5 ? 6 : 5; bx.n_pc = n_pc > 2 ? 3 : 2;
    }
    return bx;
}

// Producer: one warp issues the 32 footprint loads of a tile into ring stage `stage`.
template <int RADIUS, int CM, int VB>
__device__ __forceinline__ void lf_produce(const LookupParams& P, const LookupMaps& M, LfShared& sh,
                                           uint32_t win, const LfQuery& q, int stage, int lane, int member,
                                           bool wait_empty, uint32_t empty_parity) {
    constexpr int R = 2 * RADIUS + 1;
    constexpr int ES = lf_es(VB), LF_WIN_BYTES = lf_win_bytes(VB), LF_STAGE_BYTES = lf_stage_bytes(VB);
    uint32_t bytes = 0;
    int sel = 0, c0 = 0, c1 = 0;
    const bool mine = (lane / (32 / LF_PSPLIT)) == member;          // this warp's share of the tile's queries
    if (q.near_ && mine) {
        const int level = q.level;
        int xl, xh, yl, yh; float t0, t1;
        axis_tap<CM>(q.cx, -RADIUS, P.ax[level], xl, t0, t1);
        axis_tap<CM>(q.cx, R - 1 - RADIUS, P.ax[level], xh, t0, t1);
        axis_tap<CM>(q.cy, -RADIUS, P.ay[level], yl, t0, t1);
        axis_tap<CM>(q.cy, R - 1 - RADIUS, P.ay[level], yh, t0, t1);
        const LfBox bx = lf_box(true, xl, xh, yl, yh);
        sel = lk_shape(bx.n_rp, bx.n_pc);
        bytes = (uint32_t)(bx.n_rp * bx.n_pc * 16 * ES);              // (the TMA unit counts zero-filled bytes too)
        c0 = 2 * bx.xbase; c1 = bx.ybase >> 1;                        // signed tensor coordinates (arithmetic shift)
    }
    // the footprint arithmetic above ran while the stage was still being drained
    if (wait_empty) mbar_wait(&sh.empty[stage], empty_parity);
    if (FC_PROBE_VAL(P) & 2) bytes = 0;                                      // stage probe: no loads
    if (bytes)
        tma_load_3d(win + stage * LF_STAGE_BYTES + lane * LF_WIN_BYTES, &M.m[q.level][sel], smem_u32(&sh.full[stage]),
                    c0, c1, q.gq);
    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&sh.full[stage], total);
}

// Debug outputs of one consumer thread: floor indices per tap and the 4 corner in-bounds predicates per sample.
template <int RADIUS, int APW, bool DBG>
__device__ __forceinline__ void lf_debug_out(const LookupParams& P, const LfQuery& q, int w, const int* x0, const int* y0) {
    constexpr int R = 2 * RADIUS + 1;
    if (!DBG || !q.live) return;
    const int level = q.level, gq = q.gq;
    const int Hl = P.H[level], Wl = P.W[level];
    if (P.dbg_y0 != nullptr && w == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) P.dbg_y0[((long long)gq * P.L + level) * R + j] = y0[j];
    }
#pragma unroll
    for (int aa = 0; aa < APW; ++aa) {
        const int a = w * APW + aa;
        if (a >= R) break;
        if (P.dbg_x0 != nullptr) P.dbg_x0[((long long)gq * P.L + level) * R + a] = x0[aa];
        if (P.dbg_mask != nullptr) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const bool xa = (x0[aa] >= 0 && x0[aa] < Wl), xb = (x0[aa] + 1 >= 0 && x0[aa] + 1 < Wl);
                const bool ya = (y0[j] >= 0 && y0[j] < Hl), yb = (y0[j] + 1 >= 0 && y0[j] + 1 < Hl);
                uint8_t m = (uint8_t)((ya && xa) | ((ya && xb) << 1) | ((yb && xa) << 2) | ((yb && xb) << 3));
                if (!q.near_) m = 0;
                P.dbg_mask[(((long long)gq * P.L + level) * R + a) * R + j] = m;
            }
        }
    }
}

// Consumer: warp `w` of a group interpolates x-offsets [w*APW, w*APW + APW) of the tile.
template <int RADIUS, int CM, bool DBG, int VB>
__device__ __forceinline__ void lf_consume(const LookupParams& P, LfShared& sh, uint32_t win, const LfQuery& q,
                                           int stage, uint32_t parity, int lane, int w) {
    constexpr int R = 2 * RADIUS + 1;
    constexpr int APW = (R + LF_GWARPS - 1) / LF_GWARPS;             // x-offsets per warp
    constexpr bool EVEN = (R % APW) == 0;                            // every warp owns APW valid x-offsets
    constexpr int ES = lf_es(VB), LF_WIN_BYTES = lf_win_bytes(VB), LF_STAGE_BYTES = lf_stage_bytes(VB);
    const int level = q.level;

    // tap arithmetic overlaps the loads in flight
    int y0[R]; float wy0[R], wy1[R];
#pragma unroll
    for (int j = 0; j < R; ++j) axis_tap<CM>(q.cy, j - RADIUS, P.ay[level], y0[j], wy0[j], wy1[j]);
    int x0[APW]; float wx0[APW], wx1[APW];
#pragma unroll
    for (int aa = 0; aa < APW; ++aa) {
        const int a = EVEN ? w * APW + aa : min(w * APW + aa, R - 1);
        axis_tap<CM>(q.cx, a - RADIUS, P.ax[level], x0[aa], wx0[aa], wx1[aa]);
    }
    bool regular = true;
#pragma unroll
    for (int j = 1; j < R; ++j) regular = regular && (y0[j] == y0[0] + j);
#pragma unroll
    for (int aa = 1; aa < APW; ++aa) regular = regular && (x0[aa] == x0[0] + aa);
    // the box the producer loaded for this lane (same arithmetic on the same coordinates)
    int xl, xh;
    {
        float t0, t1;
        axis_tap<CM>(q.cx, -RADIUS, P.ax[level], xl, t0, t1);
        axis_tap<CM>(q.cx, R - 1 - RADIUS, P.ax[level], xh, t0, t1);
    }
    const LfBox d = lf_box(q.near_, xl, xh, y0[0], y0[R - 1]);
    const bool valid = d.n_rp > 0;
    const int pitch = 16 * ES * (d.n_pc > 2 ? 3 : 2);                // bytes per window row pair

    // outputs of this thread: out[b][level*R*R + (w*APW + aa)*R + j][p] = outq[(aa*R + j) * N]
    float* outq = P.io + ((long long)q.b * P.K + level * R * R + w * APW * R) * P.N + q.p;
    const int N = P.N;

    mbar_wait(&sh.full[stage], parity);
    const uint32_t wq = win + stage * LF_STAGE_BYTES + lane * LF_WIN_BYTES;
    const bool fast = q.live && valid && regular;                    // this lane's outputs come from the fast path
    // when no lane needs the per-tap slow path, the ring stage is handed back as soon as the
    // sub-windows sit in registers: a stage is then busy for the loads only, not for the
    // arithmetic and the stores
    const bool early = !__any_sync(0xffffffffu, q.live && valid && !regular) && !(FC_PROBE_VAL(P) & 1);   // FLOWCORR_PROBE=1: late release (stage probe)

    // horizontally interpolated (R + 1) x APW sub-window.  EVERY lane executes the loads (addresses are clamped into
    // the lane's own window, so lanes without a footprint read stale bytes they never use): the stage release below
    // can then depend on the last load through any lane's register.
    float h[R + 1][APW];
    {
        // byte addresses of window columns x0[0] .. x0[0]+APW (a patch jump every 8 columns)
        uint32_t col[APW + 1];
#pragma unroll
        for (int i = 0; i <= APW; ++i) {
            const int xr = min(max(x0[0] + i - d.xbase, 0), 23);
            col[i] = wq + (uint32_t)ES * (uint32_t)(xr + (xr & ~7));
        }
        // footprint row n = y0[0] - ybase + r sits at (n >> 1) * pitch + (n & 1) * 8 elements
        const int n0 = min(max(y0[0] - d.ybase, 0), 1);
        uint32_t rofs = (uint32_t)(8 * ES * n0);
        uint32_t step = n0 ? (uint32_t)(pitch - 8 * ES) : (uint32_t)(8 * ES);   // n even -> +8 elements, n odd -> +pitch - 8 elements
        float v[R + 1][APW + 1];
#pragma unroll
        for (int n = 0; n <= R; ++n) {
#pragma unroll
            for (int i = 0; i <= APW; ++i) v[n][i] = lds_vol<VB>(col[i] + rofs);
            rofs += step;
            step = (uint32_t)pitch - step;
        }
#pragma unroll
        for (int n = 0; n <= R; ++n)
#pragma unroll
            for (int aa = 0; aa < APW; ++aa) h[n][aa] = fmaf(wx1[aa], v[n][aa + 1], wx0[aa] * v[n][aa]);
    }
    if (early) {
        // The arrive must not overtake the shared loads: they drain through the LSU (bank conflicts make that take a
        // while) whereas the barrier unit answers at once, and a refill racing the LAST loads of the burst was observed
        // (wrong values in the last rows / columns of the low lanes; tests at ring-reuse sizes and the debug variant caught
        // it).  A dependency that exists only in the asm operand list is not enough -- ptxas sees no consumer of the
        // register -- so every lane's last interpolated value (it depends on the last load issued; a warp's shared loads
        // return in order) goes through a warp vote, and the arrive is predicated on the vote.  The compared pattern is a
        // NaN payload no FFMA produces, so the vote is always true; the hardware cannot know that.
        const bool landed = __any_sync(0xffffffffu, __float_as_uint(h[R][APW - 1]) != 0xffffffffu);
        if (lane == 0 && landed) mbar_arrive(&sh.empty[stage]);
    }

    if (q.live) {
        if (!valid) {
#pragma unroll
            for (int aa = 0; aa < APW; ++aa)
                if (EVEN || w * APW + aa < R) {
#pragma unroll
                    for (int j = 0; j < R; ++j) outq[(aa * R + j) * N] = 0.f;
                }
        } else if (regular) {
            float* oj = outq;                                        // row j of every x-offset: oj[aa * R * N]
#pragma unroll
            for (int j = 0; j < R; ++j) {
#pragma unroll
                for (int aa = 0; aa < APW; ++aa)
                    if (EVEN || w * APW + aa < R) oj[aa * R * N] = fmaf(wy1[j], h[j + 1][aa], wy0[j] * h[j][aa]);
                oj += N;
            }
        } else {
            // floor flips among the taps (lattice coordinates): every tap addressed on its own
#pragma unroll
            for (int aa = 0; aa < APW; ++aa) {
                if (w * APW + aa >= R) break;
                const int xa = min(max(x0[aa] - d.xbase, 0), 22), xb = xa + 1;
                const uint32_t ca = wq + (uint32_t)ES * (uint32_t)(xa + (xa & ~7)), cb = wq + (uint32_t)ES * (uint32_t)(xb + (xb & ~7));
                float* oa = outq + aa * R * N;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int ya = min(max(y0[j] - d.ybase, 0), 10), yb = ya + 1;
                    const uint32_t ra = (uint32_t)((ya >> 1) * pitch + (ya & 1) * (8 * ES));
                    const uint32_t rb = (uint32_t)((yb >> 1) * pitch + (yb & 1) * (8 * ES));
                    const float top = fmaf(wx1[aa], lds_vol<VB>(cb + ra), wx0[aa] * lds_vol<VB>(ca + ra));
                    const float bot = fmaf(wx1[aa], lds_vol<VB>(cb + rb), wx0[aa] * lds_vol<VB>(ca + rb));
                    *oa = fmaf(wy1[j], bot, wy0[j] * top);
                    oa += N;
                }
            }
        }
    }
    lf_debug_out<RADIUS, APW, DBG>(P, q, w, x0, y0);
    if (!early) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty[stage]);
    }
}

template <int RADIUS, int CM, bool DBG, int VB>
__global__ void __launch_bounds__(LF_THREADS, 1)
lookup_fwd_kernel(const __grid_constant__ LookupMaps M, const LookupParams P, int n_tiles) {
    constexpr int LF_STAGE_BYTES = lf_stage_bytes(VB);
    extern __shared__ __align__(1024) uint8_t lf_smem[];
    const uint32_t win = smem_u32(lf_smem);                          // [stage][query][6 x 3 x 16 floats]
    LfShared& sh = *reinterpret_cast<LfShared*>(lf_smem + LF_STAGES * LF_STAGE_BYTES);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < LF_STAGES; ++i) { mbar_init(&sh.full[i], LF_PSPLIT); mbar_init(&sh.empty[i], LF_GWARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    // tiles of this CTA: blockIdx.x, + gridDim.x, ...   (k-th local tile lives in stage k % LF_STAGES)
    const int first = blockIdx.x, stride = gridDim.x, L = P.L;
    const int n_local = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;
    static_assert(LF_PTEAMS == LF_GROUPS, "producers and consumers step by the same number of tiles");
    const int hop = LF_GROUPS * stride, hop_q = hop / L, hop_l = hop - hop_q * L, hop_qm = hop_q % L;   // tiles between two turns of a role

    // role index r in [0, 3): local tiles r, r + 3, ...
    const bool producer = warp < LF_PRODUCERS;
    const int cw = warp - LF_PRODUCERS;
    const int r = producer ? warp / LF_PSPLIT : cw / LF_GWARPS;
    const int sub = producer ? warp - r * LF_PSPLIT : cw - r * LF_GWARPS;
    int k = r;
    if (k >= n_local) return;
    TileIt it;
    { const int t0 = first + k * stride; it.qt = t0 / L; it.slot = t0 - it.qt * L; it.qm = it.qt % L; }
    LfQuery q = lf_load_query(P, it, lane);
    while (true) {
        const int kn = k + LF_GROUPS;
        LfQuery qn = q;
        if (kn < n_local) { it.advance(hop_q, hop_l, hop_qm, L); qn = lf_load_query(P, it, lane); }   // prefetch coords
        const int s = k % LF_STAGES;
        const uint32_t round = (uint32_t)(k / LF_STAGES);
        lf_finish_query(P, q);
        if (producer) {
            lf_produce<RADIUS, CM, VB>(P, M, sh, win, q, s, lane, sub, k >= LF_STAGES, (round & 1u) ^ 1u);
        } else {
            lf_consume<RADIUS, CM, DBG, VB>(P, sh, win, q, s, round & 1u, lane, sub);
        }
        if (kn >= n_local) break;
        k = kn; q = qn;
    }
}

// ---------------------------------------------------------------- host side
// Tensor maps are a pure function of (pyramid pointer, geometry): memoised so that the
// per-iteration call does not pay 4 * L driver encodes.
struct MapKey {
    const void* ptr; int B, H, W, L, vb;
    bool operator==(const MapKey& o) const { return ptr == o.ptr && B == o.B && H == o.H && W == o.W && L == o.L && vb == o.vb; }
};
static std::mutex g_map_mutex;
static constexpr int MAP_CACHE = 16;
static MapKey g_map_keys[MAP_CACHE];
static LookupMaps g_map_vals[MAP_CACHE];
static int g_map_next = 0, g_map_used = 0;

static int encode_level_maps(LookupMaps& M, const void* pyramid, const Pyramid& pyr, int vb) {
    const size_t es = (size_t)lf_es(vb);
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FC_ECUDA; }
    const long long Q = (long long)pyr.B * pyr.N;
    memset(&M, 0, sizeof(M));
    for (int l = 0; l < pyr.L; ++l) {
        const Level& lv = pyr.lv[l];
        cuuint64_t dims[3] = {(cuuint64_t)(2 * lv.Wp), (cuuint64_t)(lv.Hp / 2), (cuuint64_t)Q};
        cuuint64_t strides[2] = {(cuuint64_t)(2 * lv.Wp) * es, (cuuint64_t)lv.Hp * lv.Wp * es};
        cuuint32_t estr[3] = {1, 1, 1};
        void* base = const_cast<uint8_t*>(static_cast<const uint8_t*>(pyramid)) + (size_t)lv.offset * es;
        // the forward uses {5|6 row pairs} x {2|3 patches} (boxes may overhang the map: zero fill), the backward's
        // reduce-adds are clipped to the map and take any of the 18 shapes
        for (int n_rp = 1; n_rp <= 6; ++n_rp)
            for (int n_pc = 1; n_pc <= 3; ++n_pc) {
                cuuint32_t box[3] = {(cuuint32_t)(16 * n_pc), (cuuint32_t)n_rp, 1};
                CUresult r = enc(&M.m[l][lk_shape(n_rp, n_pc)],
                                 vb ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    set_error("cuTensorMapEncodeTiled failed (%d) for level %d (%dx%d, pitch %d), box %dx%d", (int)r, l, lv.H,
                              lv.W, lv.Wp, n_rp, n_pc);
                    return FC_ECUDA;
                }
            }
    }
    return FC_OK;
}

int get_level_maps(LookupMaps& M, const void* pyramid, const Pyramid& pyr, int H, int W, int vb) {
    const MapKey key{pyramid, pyr.B, H, W, pyr.L, vb};
    {
        std::lock_guard<std::mutex> g(g_map_mutex);
        for (int i = 0; i < g_map_used; ++i)
            if (g_map_keys[i] == key) { M = g_map_vals[i]; return FC_OK; }
    }
    if (int e = encode_level_maps(M, pyramid, pyr, vb)) return e;
    std::lock_guard<std::mutex> g(g_map_mutex);
    g_map_keys[g_map_next] = key; g_map_vals[g_map_next] = M;
    g_map_next = (g_map_next + 1) % MAP_CACHE;
    if (g_map_used < MAP_CACHE) ++g_map_used;
    return FC_OK;
}

template <int RADIUS, int CM, bool DBG, int VB>
static int launch_fwd4(const LookupMaps& M, const LookupParams& P, int n_tiles, int n_sm, cudaStream_t s) {
    const size_t smem = (size_t)LF_STAGES * lf_stage_bytes(VB) + sizeof(LfShared);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    FC_SMEM_ATTR_ONCE((lookup_fwd_kernel<RADIUS, CM, DBG, VB>), smem);
    lookup_fwd_kernel<RADIUS, CM, DBG, VB><<<grid, LF_THREADS, smem, s>>>(M, P, n_tiles);
    FC_LAUNCH_CHECK("lookup_fwd_kernel");
    return FC_OK;
}
template <int RADIUS, int CM, bool DBG>
static int launch_fwd3(const LookupMaps& M, const LookupParams& P, int n_tiles, int n_sm, int vb, cudaStream_t s) {
    return vb ? launch_fwd4<RADIUS, CM, DBG, 1>(M, P, n_tiles, n_sm, s) : launch_fwd4<RADIUS, CM, DBG, 0>(M, P, n_tiles, n_sm, s);
}
template <int RADIUS>
static int launch_fwd2(const LookupMaps& M, const LookupParams& P, int n_tiles, int n_sm, int coord_mode, bool dbg, int vb, cudaStream_t s) {
    if (coord_mode == FC_COORD_CUDA)
        return dbg ? launch_fwd3<RADIUS, FC_COORD_CUDA, true>(M, P, n_tiles, n_sm, vb, s)
                   : launch_fwd3<RADIUS, FC_COORD_CUDA, false>(M, P, n_tiles, n_sm, vb, s);
    if (coord_mode == FC_COORD_RAW)
        return dbg ? launch_fwd3<RADIUS, FC_COORD_RAW, true>(M, P, n_tiles, n_sm, vb, s)
                   : launch_fwd3<RADIUS, FC_COORD_RAW, false>(M, P, n_tiles, n_sm, vb, s);
    return dbg ? launch_fwd3<RADIUS, FC_COORD_CPU, true>(M, P, n_tiles, n_sm, vb, s)
               : launch_fwd3<RADIUS, FC_COORD_CPU, false>(M, P, n_tiles, n_sm, vb, s);
}

int sm_count(int& n_sm) {
    static thread_local int cached_dev = -1, cached_sm = 0;
    int dev = 0;
    FC_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        FC_CUDA(cudaDeviceGetAttribute(&cached_sm, cudaDevAttrMultiProcessorCount, dev));
        cached_dev = dev;
    }
    n_sm = cached_sm;
    return FC_OK;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_lookup_fwd(const void* pyramid, const float* coords, float* out,
                             int B, int H, int W, int num_levels, int radius,
                             int vol_dtype, int coord_mode,
                             int32_t* dbg_x0, int32_t* dbg_y0, uint8_t* dbg_mask, void* stream) {
    FC_REQUIRE(pyramid && coords && out, "fc_lookup_fwd: null pointer");
    FC_REQUIRE(vol_dtype == FC_VOL_F32 || vol_dtype == FC_VOL_BF16, "fc_lookup_fwd: unknown vol_dtype %d", vol_dtype);
    const int vb = vol_dtype == FC_VOL_BF16 ? 1 : 0;
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_lookup_fwd: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    if (int e = check_lookup_common(pyr, radius, coord_mode)) return e;
    FC_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 15u) == 0, "fc_lookup_fwd: pyramid must be 16-byte aligned");
    LookupParams P{};
    fill_params(P, pyr, radius);
    P.pyr = static_cast<const float*>(pyramid);
    P.coords = coords; P.io = out; P.gpyr = nullptr;
    P.dbg_x0 = dbg_x0; P.dbg_y0 = dbg_y0; P.dbg_mask = dbg_mask;
    if (tunables().l2_fetch) {                               // experiment switch (profiles/r02): DRAM fetch granularity of L2 misses
        static std::atomic<int> applied{0};
        if (!applied.exchange(1)) {
            size_t before = 0;
            cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)tunables().l2_fetch);
            size_t after = 0;
            cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
            note_once("l2_fetch", "cudaLimitMaxL2FetchGranularity %zu -> %zu (%s)", before, after, cudaGetErrorName(e));
            (void)cudaGetLastError();
        }
    }
    LookupMaps M;
    if (int e = get_level_maps(M, pyramid, pyr, H, W, vb)) return e;
    int n_sm = 0;
    if (int e = sm_count(n_sm)) return e;
    const int n_tiles = ((P.Q + QT - 1) / QT) * pyr.L;
    const bool dbg = dbg_x0 || dbg_y0 || dbg_mask;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (radius) {
        case 1: return launch_fwd2<1>(M, P, n_tiles, n_sm, coord_mode, dbg, vb, s);
        case 2: return launch_fwd2<2>(M, P, n_tiles, n_sm, coord_mode, dbg, vb, s);
        case 3: return launch_fwd2<3>(M, P, n_tiles, n_sm, coord_mode, dbg, vb, s);
        default: return launch_fwd2<4>(M, P, n_tiles, n_sm, coord_mode, dbg, vb, s);
    }
}
