// C ABI entry points that are not kernels themselves: error plumbing, geometry,
// and the math-mode dispatch of fc_build / fc_build_bwd.  See include/flowcorr.h.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>

#include "fc_tma.cuh"

namespace fc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return FC_ECUDA;
}

// fc_simt.cu
int simt_build(const float* f1, const float* f2, float* pyramid, const Pyramid& pyr,
               int D, int H, int W, cudaStream_t s);
int simt_fold(float* gpyr, const Pyramid& pyr, cudaStream_t s);
int simt_build_bwd(float* gpyr, const float* f1, const float* f2, float* d1, float* d2,
                   const Pyramid& pyr, int D, int H, int W, cudaStream_t s);
// fc_build_tc.cu
size_t tc_build_workspace_bytes(int B, int D, int H, int W, int L, int math);
struct FeatSource { const float* x; const void* packed_w; int C; };     // fc_build.cuh
int tc_build(const float* f1, const float* f2, void* pyramid, const Pyramid& pyr, int D, int H, int W,
             int vol_dtype, int math, void* ws, size_t ws_bytes, cudaStream_t s, const FeatSource* feat);

// fc_bwd_tc.cu
bool tc_bwd_supported(int D, int H, int W);
size_t tc_bwd_workspace_bytes(int B, int D, int H, int W);
int tc_build_bwd(float* gpyr, const float* f1, const float* f2, float* d1, float* d2, const Pyramid& pyr,
                 int D, int H, int W, int math, void* ws, size_t ws_bytes, cudaStream_t s);

EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static Tunables& tunables_mut() {
    static Tunables t = []() {
        Tunables v{};
        auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
#ifdef FC_PROBES
        v.probe = geti("FLOWCORR_PROBE", 0);
#endif
        v.build_sched = geti("FLOWCORR_BUILD_SCHED", 1);
        v.build_stages = geti("FLOWCORR_BUILD_STAGES", 0);
        v.build_epi_warps = geti("FLOWCORR_BUILD_EPI_WARPS", 4) == 8 ? 8 : 4;
        v.no_fuse = getenv("FLOWCORR_NO_FUSE") != nullptr;
        v.l2_fetch = geti("FLOWCORR_L2_FETCH", 0);
        v.bwd_fused = geti("FLOWCORR_BWD_FUSED", 1);
        v.pdl = geti("FLOWCORR_PDL", 1);
        v.verbose = geti("FLOWCORR_VERBOSE", 1);
        return v;
    }();
    return t;
}
const Tunables& tunables() { return tunables_mut(); }

int sm_count_cached() {
    static thread_local int cached_dev = -1, cached_sm = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) { cached_sm = n; cached_dev = dev; }
    }
    return cached_sm;
}

void note_once(const char* key, const char* fmt, ...) {
    if (!tunables().verbose) return;
    static std::mutex mu;
    static std::set<std::string> seen;
    {
        std::lock_guard<std::mutex> g(mu);
        if (!seen.insert(key).second) return;
    }
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[flowcorr] %s\n", buf);
}

// ---- memoised tensor-map encoder
namespace {
struct EncKey {
    const void* base; int dtype, rank, swizzle, l2;
    cuuint64_t dims[5], strides[4]; cuuint32_t box[5];
    bool operator==(const EncKey& o) const { return memcmp(this, &o, sizeof(EncKey)) == 0; }
};
constexpr int ENC_CACHE = 1024;               // direct-mapped by a hash of the key
std::mutex g_enc_mutex;
EncKey g_enc_keys[ENC_CACHE];
CUtensorMap g_enc_vals[ENC_CACHE];
bool g_enc_valid[ENC_CACHE];
unsigned enc_hash(const EncKey& k) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
    unsigned long long h = 1469598103934665603ull;                       // FNV-1a
    for (size_t i = 0; i < sizeof(EncKey); ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return (unsigned)(h ^ (h >> 32)) % ENC_CACHE;
}
}  // namespace

int encode_tiled_cached(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base,
                        const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                        CUtensorMapSwizzle swizzle, CUtensorMapL2promotion l2) {
    EncKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.dtype = (int)dtype; k.rank = rank; k.swizzle = (int)swizzle; k.l2 = (int)l2;
    for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides_bytes[i];
    const unsigned slot = enc_hash(k);
    {
        std::lock_guard<std::mutex> g(g_enc_mutex);
        if (g_enc_valid[slot] && g_enc_keys[slot] == k) { *out = g_enc_vals[slot]; return FC_OK; }
    }
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FC_ECUDA; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d, dims %llu x %llu, box %u x %u", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1u);
        return FC_ECUDA;
    }
    std::lock_guard<std::mutex> g(g_enc_mutex);
    g_enc_keys[slot] = k; g_enc_vals[slot] = *out; g_enc_valid[slot] = true;
    return FC_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fc

using namespace fc;

extern "C" int fc_abi_version(void) { return FC_ABI_VERSION; }

extern "C" int fc_tunable_set(const char* name, int value) {
    FC_REQUIRE(name != nullptr, "fc_tunable_set: null name");
    Tunables& t = tunables_mut();
    const std::string n(name);
    if (n == "build_sched") t.build_sched = value;
    else if (n == "build_stages") t.build_stages = value;
    else if (n == "build_epi_warps") t.build_epi_warps = value == 8 ? 8 : 4;
    else if (n == "no_fuse") t.no_fuse = value != 0;
    else if (n == "verbose") t.verbose = value;
    else if (n == "pdl") t.pdl = value;
    else if (n == "bwd_fused") t.bwd_fused = value;
    else if (n == "l2_fetch") t.l2_fetch = value;
#ifdef FC_PROBES
    else if (n == "probe") t.probe = value;
#endif
    else { set_error("fc_tunable_set: unknown switch '%s'", name); return FC_EINVAL; }
    return FC_OK;
}

extern "C" const char* fc_last_error(void) { return g_err; }

extern "C" unsigned long long fc_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int fc_level_dims(int H, int W, int level, int* Hl, int* Wl, int* Wp) {
    FC_REQUIRE(H > 0 && W > 0 && level >= 0 && level < FC_MAX_LEVELS, "fc_level_dims: bad arguments");
    const int h = H >> level, w = W >> level;
    FC_REQUIRE(h >= 1 && w >= 1, "fc_level_dims: level %d of %dx%d is empty", level, H, W);
    if (Hl) *Hl = h;
    if (Wl) *Wl = w;
    if (Wp) *Wp = round_up(w, 8);
    return FC_OK;
}

extern "C" size_t fc_pyramid_bytes(int B, int H, int W, int num_levels, int vol_dtype,
                                   size_t* level_offsets) {
    Pyramid pyr;
    if (!make_pyramid(pyr, B, H, W, num_levels) || (vol_dtype != FC_VOL_F32 && vol_dtype != FC_VOL_BF16)) {
        set_error("fc_pyramid_bytes: bad geometry B=%d H=%d W=%d L=%d dtype=%d", B, H, W, num_levels, vol_dtype);
        return 0;
    }
    const size_t es = vol_dtype == FC_VOL_F32 ? 4 : 2;
    if (level_offsets)
        for (int l = 0; l < num_levels; ++l) level_offsets[l] = (size_t)pyr.lv[l].offset * es;
    return (size_t)pyr.total * es;
}

extern "C" size_t fc_build_workspace_bytes(int B, int D, int H, int W, int num_levels, int math) {
    if (math == FC_MATH_FP32) return 0;
    return tc_build_workspace_bytes(B, D, H, W, num_levels, math);
}

extern "C" int fc_build(const float* fmap1, const float* fmap2, void* pyramid,
                        int B, int D, int H, int W, int num_levels,
                        int vol_dtype, int math,
                        void* workspace, size_t workspace_bytes, void* stream) {
    FC_REQUIRE(fmap1 && fmap2 && pyramid, "fc_build: null pointer");
    FC_REQUIRE(aligned16(fmap1) && aligned16(fmap2) && aligned16(pyramid), "fc_build: pointers must be 16-byte aligned");
    FC_REQUIRE(D >= 1, "fc_build: D=%d", D);
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_build: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (math == FC_MATH_FP32) {
        FC_REQUIRE(vol_dtype == FC_VOL_F32, "fc_build: FC_MATH_FP32 writes an fp32 volume only");
        return simt_build(fmap1, fmap2, static_cast<float*>(pyramid), pyr, D, H, W, s);
    }
    FC_REQUIRE(math == FC_MATH_TC_3XBF16 || math == FC_MATH_TC_BF16, "fc_build: unknown math mode %d", math);
    return tc_build(fmap1, fmap2, pyramid, pyr, D, H, W, vol_dtype, math, workspace, workspace_bytes, s, nullptr);
}

extern "C" size_t fc_build_bwd_workspace_bytes(int B, int D, int H, int W, int num_levels, int math) {
    (void)num_levels;
    if (math == FC_MATH_FP32) return 0;
    return tc_bwd_workspace_bytes(B, D, H, W);        // 0: this shape runs the fp32 CUDA-core contractions
}

extern "C" int fc_build_bwd(float* grad_pyramid, const float* fmap1, const float* fmap2,
                            float* dfmap1, float* dfmap2,
                            int B, int D, int H, int W, int num_levels, int math,
                            void* workspace, size_t workspace_bytes, void* stream) {
    FC_REQUIRE(grad_pyramid && fmap1 && fmap2, "fc_build_bwd: null pointer");
    FC_REQUIRE(aligned16(grad_pyramid) && aligned16(fmap1) && aligned16(fmap2) && aligned16(dfmap1) && aligned16(dfmap2),
               "fc_build_bwd: pointers must be 16-byte aligned");
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_build_bwd: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    FC_REQUIRE(math == FC_MATH_FP32 || math == FC_MATH_TC_3XBF16 || math == FC_MATH_TC_BF16,
               "fc_build_bwd: unknown math mode %d", math);
    // tensor-core modes: fused fold + bf16 split in place, then two tcgen05 GEMMs; shapes the
    // tensor-core kernel does not take (fc_build_bwd_workspace_bytes == 0) run the fp32 mode
    if (math != FC_MATH_FP32 && tc_bwd_supported(D, H, W))
        return tc_build_bwd(grad_pyramid, fmap1, fmap2, dfmap1, dfmap2, pyr, D, H, W, math, workspace, workspace_bytes, s);
    if (math != FC_MATH_FP32)
        note_once("build_bwd_simt", "fc_build_bwd: D=%d, %dx%d tokens is outside the tensor-core backward's range "
                  "(D %% 64 == 0, D <= 256, padded map <= 16384 targets): running the fp32 CUDA-core contractions", D, H, W);
    if (int e = simt_fold(grad_pyramid, pyr, s)) return e;
    return simt_build_bwd(grad_pyramid, fmap1, fmap2, dfmap1, dfmap2, pyr, D, H, W, s);
}

extern "C" int fc_build_from_fnet_tail(const float* x, const void* packed_weights, void* pyramid,
                                       int B, int C, int D, int H, int W, int num_levels, int vol_dtype, int math,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    FC_REQUIRE(x && packed_weights && pyramid, "fc_build_from_fnet_tail: null pointer");
    FC_REQUIRE(aligned16(x) && aligned16(pyramid), "fc_build_from_fnet_tail: pointers must be 16-byte aligned");
    FC_REQUIRE(math == FC_MATH_TC_3XBF16 || math == FC_MATH_TC_BF16, "fc_build_from_fnet_tail: tensor-core math modes only (got %d)", math);
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_build_from_fnet_tail: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    const FeatSource src{x, packed_weights, C};
    return tc_build(nullptr, nullptr, pyramid, pyr, D, H, W, vol_dtype, math, workspace, workspace_bytes,
                    static_cast<cudaStream_t>(stream), &src);
}
