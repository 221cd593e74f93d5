// Shared device/host helpers for libflowcorr (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>

#include "../../include/flowcorr.h"

namespace fc {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();          // fc_kernel_launches() instrumentation

#define FC_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            fc::set_error(__VA_ARGS__);       \
            return FC_EINVAL;                 \
        }                                     \
    } while (0)

#define FC_CUDA(call)                                         \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) { (void)cudaGetLastError(); return fc::cuda_fail(e__, #call); } \
    } while (0)

#define FC_LAUNCH_CHECK(name)                                 \
    do {                                                      \
        fc::count_launch();                                   \
        cudaError_t e__ = cudaGetLastError();   /* clears a non-sticky error: a later call must not inherit it */ \
        if (e__ != cudaSuccess) return fc::cuda_fail(e__, name); \
    } while (0)

// Stage probes (tools/probe_bounds.py) exist only in a library compiled with -DFC_PROBES; the shipped
// library has no probe branches in its kernels and never reads FLOWCORR_PROBE.
#ifdef FC_PROBES
#define FC_PROBE_VAL(P) ((P).probe)
#else
#define FC_PROBE_VAL(P) 0
#endif

// Run-time switches, read from the environment ONCE per process (fc_api.cu).
struct Tunables {
    int probe;               // FLOWCORR_PROBE (FC_PROBES builds only, else 0)
    int build_sched;         // FLOWCORR_BUILD_SCHED      0 | 1 (default 1: pair-tiles strided over the CTA pairs)
    int build_stages;        // FLOWCORR_BUILD_STAGES     operand ring depth, 0 = kernel default
    int build_epi_warps;     // FLOWCORR_BUILD_EPI_WARPS  4 | 8 (default 4)
    int no_fuse;             // FLOWCORR_NO_FUSE          pyramid by separate pooling launches
    int l2_fetch;            // FLOWCORR_L2_FETCH         32 | 64 | 128: cudaLimitMaxL2FetchGranularity set at the first lookup (0 = leave)
    int bwd_fused;           // FLOWCORR_BWD_FUSED        1 (default): fold + bf16 split inside the backward GEMMs (aligned maps: TMA coarse boxes);
                             //                           2: the generic fold-in-GEMM kernel for every map; 0: separate fold + pack pass
    int pdl;                 // FLOWCORR_PDL              1 (default): lookups launch with programmatic stream serialisation
    int verbose;             // FLOWCORR_VERBOSE          log mode fall-backs (shape not taken by a tensor-core kernel) to stderr
};
const Tunables& tunables();
// SM count of the current device (cached per host thread); 148 (B200) if the query fails
int sm_count_cached();
// one line on stderr the first time `key` is seen (FLOWCORR_VERBOSE=0 silences): fall-backs must not be silent
void note_once(const char* key, const char* fmt, ...);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site = kernel instantiation, device)
#define FC_SMEM_ATTR_ONCE(kernel, bytes)                                                                   \
    do {                                                                                                   \
        static std::atomic<unsigned long long> done__{0};                                                  \
        int dev__ = 0;                                                                                     \
        FC_CUDA(cudaGetDevice(&dev__));                                                                    \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                             \
        if (!(done__.load(std::memory_order_acquire) & bit__)) {                                           \
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            done__.fetch_or(bit__, std::memory_order_release);                                             \
        }                                                                                                  \
    } while (0)

// the same for a kernel whose dynamic shared memory varies with the geometry: the attribute only ever grows
#define FC_SMEM_ATTR_GROW(kernel, bytes)                                                                   \
    do {                                                                                                   \
        static std::atomic<int> cur__[64];                                                                 \
        int dev__ = 0;                                                                                     \
        FC_CUDA(cudaGetDevice(&dev__));                                                                    \
        if ((int)(bytes) > cur__[dev__ & 63].load(std::memory_order_acquire)) {                            \
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            cur__[dev__ & 63].store((int)(bytes), std::memory_order_release);                              \
        }                                                                                                  \
    } while (0)

// ---------------------------------------------------------------- pyramid geometry
struct Level {
    int H, W, Wp, Hp;      // rows, valid columns, row pitch (multiple of 8), padded rows (even)
    long long offset;      // ELEMENT offset of the level inside the pyramid buffer
};

// Patch layout of one query's map: 64-byte patches of 2 rows x 8 columns, patches of a
// row pair stored left to right (DRAM granule = 64 B: a (2r+2)^2 window touches ~25% fewer
// granules than with row-major rows).  Element (y, x) of a map with row pitch Wp:
__host__ __device__ __forceinline__ int tile_off(int y, int x, int Wp) {
    return (y >> 1) * (2 * Wp) + ((x >> 3) << 4) + ((y & 1) << 3) + (x & 7);
}
__host__ __device__ __forceinline__ void tile_inv(int off, int Wp, int& y, int& x) {
    const int pr = off / (2 * Wp), r = off - pr * 2 * Wp;
    y = 2 * pr + ((r >> 3) & 1);
    x = ((r >> 4) << 3) + (r & 7);
}

struct Pyramid {
    int L;
    int B, N;              // samples, queries per sample (H*W of level 0)
    Level lv[FC_MAX_LEVELS];
    long long total;       // total elements
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

inline bool make_pyramid(Pyramid& P, int B, int H, int W, int L) {
    if (B <= 0 || H <= 0 || W <= 0 || L < 1 || L > FC_MAX_LEVELS) return false;
    P.L = L; P.B = B; P.N = H * W;
    long long off = 0;
    for (int l = 0; l < L; ++l) {
        int Hl = H >> l, Wl = W >> l;
        if (Hl < 1 || Wl < 1) return false;
        P.lv[l].H = Hl; P.lv[l].W = Wl; P.lv[l].Wp = round_up(Wl, 8); P.lv[l].Hp = round_up(Hl, 2);
        P.lv[l].offset = off;
        off += (long long)B * P.N * P.lv[l].Hp * P.lv[l].Wp;
    }
    P.total = off;
    return true;
}

// ---------------------------------------------------------------- tap arithmetic
// One window sample on one axis, rounded exactly where the reference rounds
// (corr.py:41-43, utils.py:61-62, ATen GridSampler.cuh unnormalise/floor/weights;
// restated in oracle/corr_spec.py::axis_taps).  The *_rn intrinsics are never
// contracted into FMAs, so the integer index is reproducible bit for bit.
struct AxisConst {
    float den;   // size - 1
    float inv;   // fl(1 / (size - 1))   (ATen CUDA div-by-scalar fast path)
};

__host__ __device__ inline AxisConst make_axis(int size) {
    AxisConst a;
    a.den = (float)(size - 1);
    a.inv = 1.0f / a.den;
    return a;
}

template <int COORD_MODE>
__device__ __forceinline__ void axis_tap(float c_l, int off, AxisConst ax,
                                         int& i0, float& w0, float& w1) {
    if (COORD_MODE == FC_COORD_RAW) {
        // on-demand semantics (corr.py:85, correlation_kernel.cu:67-76): no normalise round trip,
        // one floor and one fraction per axis shared by every tap
        const float f0 = floorf(c_l);
        w1 = __fsub_rn(c_l, f0);
        w0 = __fsub_rn(1.0f, w1);
        i0 = (int)f0 + off;
        return;
    }
    float X = __fadd_rn(c_l, (float)off);
    float t = __fmul_rn(2.0f, X);
    float g = (COORD_MODE == FC_COORD_CUDA) ? __fmul_rn(t, ax.inv) : __fdiv_rn(t, ax.den);
    g = __fadd_rn(g, -1.0f);
    float u = __fmul_rn(__fadd_rn(g, 1.0f), 0.5f);
    float ix = __fmul_rn(u, ax.den);
    float f = floorf(ix);
    w1 = __fsub_rn(ix, f);
    w0 = __fsub_rn(__fadd_rn(f, 1.0f), ix);
    i0 = (int)f;            // saturating; NaN -> 0 (such queries are masked as "far")
}

// ---------------------------------------------------------------- small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}  // namespace fc
