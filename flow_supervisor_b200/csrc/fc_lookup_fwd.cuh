// Device side of the pyramid lookup forward: ring, producer and consumer templates, shared by lookup_fwd_kernel
// (fc_lookup_fwd.cu) and the lookup fused into the motion encoder's first convolution (fc_lookup_conv.cu).
#pragma once

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

// Where a consumer thread's outputs go: tap (aa, j) = x-offset w * APW + aa, y-offset j of the tile's level.
// GlobalSink: the (B, K, H, W) output tensor, one coalesced 128-byte row per warp store.
struct GlobalSink {
    float* outq;            // out[b][level*R*R + w*APW*R][p]
    long long N;
    __device__ __forceinline__ void put(int aa, int j, int R, float v) { outq[(long long)(aa * R + j) * N] = v; }
};
// RegSink: kept in registers for a fused consumer (fc_lookup_conv.cu)
template <int APW, int R>
struct RegSink {
    float o[APW][R];
    __device__ __forceinline__ void put(int aa, int j, int, float v) { o[aa][j] = v; }
};

// Consumer: warp `w` of a group interpolates x-offsets [w*APW, w*APW + APW) of the tile.
template <int RADIUS, int CM, bool DBG, int VB, typename Sink>
__device__ __forceinline__ void lf_consume(const LookupParams& P, LfShared& sh, uint32_t win, const LfQuery& q,
                                           int stage, uint32_t parity, int lane, int w, Sink& sink) {
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


    mbar_wait(&sh.full[stage], parity);
    const uint32_t wq = win + stage * LF_STAGE_BYTES + lane * LF_WIN_BYTES;
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
                    for (int j = 0; j < R; ++j) sink.put(aa, j, R, 0.f);
                }
        } else if (regular) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
#pragma unroll
                for (int aa = 0; aa < APW; ++aa)
                    if (EVEN || w * APW + aa < R) sink.put(aa, j, R, fmaf(wy1[j], h[j + 1][aa], wy0[j] * h[j][aa]));
            }
        } else {
            // floor flips among the taps (lattice coordinates): every tap addressed on its own
#pragma unroll
            for (int aa = 0; aa < APW; ++aa) {
                if (w * APW + aa >= R) break;
                const int xa = min(max(x0[aa] - d.xbase, 0), 22), xb = xa + 1;
                const uint32_t ca = wq + (uint32_t)ES * (uint32_t)(xa + (xa & ~7)), cb = wq + (uint32_t)ES * (uint32_t)(xb + (xb & ~7));
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int ya = min(max(y0[j] - d.ybase, 0), 10), yb = ya + 1;
                    const uint32_t ra = (uint32_t)((ya >> 1) * pitch + (ya & 1) * (8 * ES));
                    const uint32_t rb = (uint32_t)((yb >> 1) * pitch + (yb & 1) * (8 * ES));
                    const float top = fmaf(wx1[aa], lds_vol<VB>(cb + ra), wx0[aa] * lds_vol<VB>(ca + ra));
                    const float bot = fmaf(wx1[aa], lds_vol<VB>(cb + rb), wx0[aa] * lds_vol<VB>(ca + rb));
                    sink.put(aa, j, R, fmaf(wy1[j], bot, wy0[j] * top));
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
    // programmatic dependent launch (fc_lookup_fwd.cu launches with programmatic stream serialisation): a CTA of this grid
    // may have become resident while the previous kernel of the stream was still draining; everything above (barrier
    // set-up) overlapped that.  From here on the coordinates and the pyramid are read: wait for the previous grid's memory.
    // The NEXT kernel of the stream, if it is launched the same way (the next lookup), may start taking over SMs as soon as
    // CTAs of this grid exit.
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

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
            constexpr int R = 2 * RADIUS + 1, APW = (R + LF_GWARPS - 1) / LF_GWARPS;
            // outputs of this thread: out[b][level*R*R + (sub*APW + aa)*R + j][p]
            GlobalSink sink{P.io + ((long long)q.b * P.K + q.level * R * R + sub * APW * R) * P.N + q.p, P.N};
            lf_consume<RADIUS, CM, DBG, VB>(P, sh, win, q, s, round & 1u, lane, sub, sink);
        }
        if (kn >= n_local) break;
        k = kn; q = qn;
    }
}

}  // namespace fc
