// Shared between the lookup forward (fc_lookup_fwd.cu) and backward (fc_lookup.cu).
#pragma once

#include <stdlib.h>

#include "fc_tma.cuh"

namespace fc {

constexpr int QT = 32;            // queries per tile (one per lane)

struct LookupParams {
    const float* pyr;
    const float* coords;    // (B, 2, H, W)
    float* io;              // forward: out (B, K, H, W); backward: grad_out (read)
    float* gpyr;            // backward only: gradient pyramid (atomically accumulated)
    int Q;                  // B * N (< 2^31)
    int N, L, K;
    long long off[FC_MAX_LEVELS];
    int H[FC_MAX_LEVELS], W[FC_MAX_LEVELS], Wp[FC_MAX_LEVELS];
    int msize[FC_MAX_LEVELS];   // elements per query map (Hp * Wp)
    AxisConst ax[FC_MAX_LEVELS], ay[FC_MAX_LEVELS];
    int nrp[FC_MAX_LEVELS], npc[FC_MAX_LEVELS];   // row pairs (Hp / 2) and 8-column patches (Wp / 8) per map
    float inv_scale[FC_MAX_LEVELS];
    int probe;              // always 0 unless the library is compiled with -DFC_PROBES (stage probes of tools/probe_bounds.py:
                            // FLOWCORR_PROBE=n switches one pipeline stage off, results are then garbage)
    uint32_t div_m, div_s;  // gq / N without a division: (__umulhi(div_m, gq) + gq) >> div_s   (gq < 2^31)
    int32_t* dbg_x0;
    int32_t* dbg_y0;
    uint8_t* dbg_mask;
};

// Per level, 18 TMA views of the (gradient) pyramid as a 3-D tensor [query][row pair][2*Wp floats]
// over the 2x8-patch layout: boxes of {1..3 patches, 1..6 row pairs, 1 query}.  The forward (footprint loads) uses
// {5|6} x {2|3} at signed coordinates and lets the TMA unit zero-fill what lies outside the map; the backward
// (footprint reduce-adds, which trap at negative coordinates) clips its boxes to the map and takes any shape.
// Memoised per (pointer, geometry) in fc_lookup_fwd.cu.
constexpr int LK_SHAPES = 18;
__host__ __device__ __forceinline__ int lk_shape(int n_rp, int n_pc) { return (n_rp - 1) * 3 + (n_pc - 1); }
struct LookupMaps {
    CUtensorMap m[FC_MAX_LEVELS][LK_SHAPES];      // [level][lk_shape(row pairs, patches)]
};
int get_level_maps(LookupMaps& M, const void* pyramid, const Pyramid& pyr, int H, int W, int vb);
int sm_count(int& n_sm);

#ifdef __CUDACC__
// sample index of global query gq (< 2^31) and its token inside the sample
__device__ __forceinline__ void split_query(const LookupParams& P, int gq, int& b, int& p) {
    b = (int)((__umulhi(P.div_m, (uint32_t)gq) + (uint32_t)gq) >> P.div_s);
    p = gq - b * P.N;
}
#endif

#ifdef __CUDACC__
// A role walks tiles t0, t0 + hop, ...; tile = qt * L + slot.  Roles step by a fixed
// number of tiles, so (qt, level) advance by precomputed increments instead of a division per tile.
struct TileIt {
    int qt, slot, qm;       // tile t = qt * L + slot; qm = qt % L; the tile's level is (slot + qm) % L
    // The rotation by qt makes consecutive tiles of a CTA (stride = grid size, usually a multiple of L) walk
    // through the levels: the in-bounds share of a footprint (hence the DRAM bytes of a tile) differs per level,
    // and a CTA pinned to one level would set the pace.
    __device__ __forceinline__ int level(int L) const { const int l = slot + qm; return l >= L ? l - L : l; }
    __device__ __forceinline__ void advance(int dq, int dl, int dqm, int L) {
        qt += dq; slot += dl; qm += dqm;
        if (slot >= L) { slot -= L; ++qt; ++qm; }
        while (qm >= L) qm -= L;
    }
};

#endif

inline void fill_params(LookupParams& P, const Pyramid& pyr, int radius) {
    P.Q = pyr.B * pyr.N;
    P.N = pyr.N; P.L = pyr.L;
#ifdef FC_PROBES
    const char* pr = getenv("FLOWCORR_PROBE");
    P.probe = pr ? atoi(pr) : 0;
#else
    P.probe = 0;
#endif
    // Granlund-Montgomery round-up multiplier for the divisor N
    P.div_s = 0;
    while ((1ull << P.div_s) < (unsigned long long)pyr.N) ++P.div_s;
    P.div_m = (uint32_t)(((1ull << 32) * ((1ull << P.div_s) - (unsigned long long)pyr.N)) / (unsigned long long)pyr.N + 1);
    const int R = 2 * radius + 1;
    P.K = pyr.L * R * R;
    for (int l = 0; l < pyr.L; ++l) {
        P.off[l] = pyr.lv[l].offset;
        P.H[l] = pyr.lv[l].H; P.W[l] = pyr.lv[l].W; P.Wp[l] = pyr.lv[l].Wp;
        P.msize[l] = pyr.lv[l].Hp * pyr.lv[l].Wp;
        P.nrp[l] = pyr.lv[l].Hp / 2; P.npc[l] = pyr.lv[l].Wp / 8;
        P.ax[l] = make_axis(pyr.lv[l].W);
        P.ay[l] = make_axis(pyr.lv[l].H);
        P.inv_scale[l] = 1.0f / (float)(1 << l);
    }
}

inline int check_lookup_common(const Pyramid& pyr, int radius, int coord_mode) {
    FC_REQUIRE((long long)pyr.B * pyr.N < (1LL << 31) - QT, "B*H*W = %lld queries exceed 2^31", (long long)pyr.B * pyr.N);
    FC_REQUIRE(radius >= 1 && radius <= FC_MAX_RADIUS, "radius %d unsupported (1..%d)", radius, FC_MAX_RADIUS);
    FC_REQUIRE(coord_mode == FC_COORD_CUDA || coord_mode == FC_COORD_CPU || coord_mode == FC_COORD_RAW, "bad coord_mode %d", coord_mode);
    // FC_COORD_RAW (AlternateCorrBlock's indexing) never divides by (size - 1): unit dimensions are fine there
    for (int l = 0; l < pyr.L && coord_mode != FC_COORD_RAW; ++l)
        FC_REQUIRE(pyr.lv[l].H >= 2 && pyr.lv[l].W >= 2,
                   "pyramid level %d is %dx%d: a unit dimension makes the reference divide by zero "
                   "(utils.py:61-62); inputs must be at least %d px on a side",
                   l, pyr.lv[l].H, pyr.lv[l].W, 16 << (pyr.L - 1));
    return 0;
}

}  // namespace fc
