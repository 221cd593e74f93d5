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


#include "fc_lookup_fwd.cuh"

namespace fc {

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
    // programmatic stream serialisation: the grid may be scheduled while the previous kernel of the stream drains (its CTAs
    // wait in griddepcontrol.wait before they touch memory), which hides the launch latency between the lookups of an iteration
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(LF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = tunables().pdl ? 1 : 0;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    FC_CUDA(cudaLaunchKernelEx(&cfg, lookup_fwd_kernel<RADIUS, CM, DBG, VB>, M, P, n_tiles));
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
