// Pyramid lookup BACKWARD (autograd of CorrBlock.__call__, corr.py:29-50; the forward
// lives in fc_lookup_fwd.cu).  HBM-bound scatter: one CTA = 32 consecutive queries x one
// pyramid level.  The (2r+2)^2 footprint of every query (plus one guard row/column
// on each side for floor flips of the normalise/un-normalise round trip) is staged
// into a skewed shared-memory window with 16-byte cp.async row chunks; interpolation
// then reads the window with lane <-> query so the (B, K, H, W) output stores are
// 128-byte coalesced.
#include "fc_lookup.cuh"

namespace fc {

constexpr int WIN_ROWS = 12;      // (2r+2) + 2 guard rows, r <= 4
constexpr int WIN_PITCH = 16;     // floats per window row: 3 alignment + 12 + 1 spare
constexpr int WIN_STRIDE = 224;   // floats per query window incl. skew room (== 0 mod 32)
constexpr int LOOKUP_THREADS = 96;
constexpr int A_PER_WARP = 3;     // x-offsets handled by one warp

// Bank skew: lanes reading the same (row, column-within-chunk) of their own windows
// land on 32 different banks when the per-query alignment offsets cycle mod 4 (the
// smooth-flow case): 4 banks from x0 mod 4, x4 from (q>>2)&3, x2 from (q>>4)&1.
__device__ __forceinline__ int win_base(int q) {
    return q * WIN_STRIDE + 4 * ((q >> 2) & 3) + 16 * ((q >> 4) & 1);
}

struct WinDesc {            // per-query footprint descriptor (shared memory)
    int y_lo[QT];           // first footprint row (may be negative)
    int x_s[QT];            // first footprint column rounded down to a multiple of 4
    int n_row[QT];          // rows to stage   (0 = nothing: dead or far query)
    int n_chunk[QT];        // 16-byte chunks per row to stage
    long long q_off[QT];    // element offset of the query's map inside the level
};

template <int RADIUS, int CM>
__device__ __forceinline__ void query_setup(const LookupParams& P, int level, int lane,
                                            int gq0, bool live, int& b, int& p,
                                            float& cx, float& cy, bool& near_,
                                            WinDesc& d, bool write_desc) {
    constexpr int R = 2 * RADIUS + 1;
    // 32 consecutive flattened queries: one 32-bit division per thread, then a wrap
    b = gq0 / P.N;
    p = gq0 - b * P.N + lane;
    while (p >= P.N) { p -= P.N; ++b; }
    cx = 0.f; cy = 0.f;
    if (live) {
        const float* c = P.coords + (long long)b * 2 * P.N + p;
        cx = __fmul_rn(__ldg(c), P.inv_scale[level]);
        cy = __fmul_rn(__ldg(c + P.N), P.inv_scale[level]);
    }
    // beyond 2^20 every tap is out of bounds for any map this library accepts and the
    // +-1 flip bound used to size the window no longer holds; NaN compares false.
    near_ = live && (fabsf(cx) < 1048576.f) && (fabsf(cy) < 1048576.f);
    if (write_desc) {
        int nrow = 0, nchunk = 0, ylo = 0, xs = 0;
        if (near_) {
            int xl, xh, yl, yh; float t0, t1;
            axis_tap<CM>(cx, -RADIUS, P.ax[level], xl, t0, t1);
            axis_tap<CM>(cx, R - 1 - RADIUS, P.ax[level], xh, t0, t1);
            axis_tap<CM>(cy, -RADIUS, P.ay[level], yl, t0, t1);
            axis_tap<CM>(cy, R - 1 - RADIUS, P.ay[level], yh, t0, t1);
            ylo = yl;
            xs = xl & ~3;                                  // floor to multiple of 4 (two's complement)
            nrow = min(yh + 1 - yl + 1, WIN_ROWS);
            nchunk = min(((xh + 1 - xs) >> 2) + 1, WIN_PITCH / 4);
        }
        d.y_lo[lane] = ylo; d.x_s[lane] = xs; d.n_row[lane] = nrow; d.n_chunk[lane] = nchunk;
        d.q_off[lane] = (long long)(gq0 + lane) * P.msize[level];
    }
}

// Backward: the same window, used as an accumulator.  Every output gradient is
// splatted into shared memory (4 shared atomics), then the window is flushed with
// 16-byte vector reductions (red.global.add.v4.f32) into the gradient pyramid.
__device__ __forceinline__ void red_add_v4(float* gptr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(gptr), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int RADIUS, int CM>
__global__ void __launch_bounds__(LOOKUP_THREADS)
lookup_bwd_kernel(const LookupParams P) {
    constexpr int R = 2 * RADIUS + 1;
    __shared__ __align__(16) float win[QT * WIN_STRIDE];
    __shared__ WinDesc desc;

    const int level = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq0 = blockIdx.x * QT;
    const int gq = gq0 + lane;
    const bool live = gq < P.Q;

    float cx, cy; bool near_; int b, p;
    query_setup<RADIUS, CM>(P, level, lane, gq0, live, b, p, cx, cy, near_, desc, warp == 0);
    for (int i = threadIdx.x; i < QT * WIN_STRIDE / 4; i += LOOKUP_THREADS)
        reinterpret_cast<float4*>(win)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    if (near_) {
        int y0[R]; float wy0[R], wy1[R];
#pragma unroll
        for (int j = 0; j < R; ++j) axis_tap<CM>(cy, j - RADIUS, P.ay[level], y0[j], wy0[j], wy1[j]);
        const int ylo = desc.y_lo[lane], xs = desc.x_s[lane];
        const float* gq_ptr = P.io + ((long long)b * P.K + level * R * R) * P.N + p;
        float* wq = win + win_base(lane);
#pragma unroll
        for (int aa = 0; aa < A_PER_WARP; ++aa) {
            const int a = warp * A_PER_WARP + aa;
            if (a >= R) break;
            int x0; float wx0, wx1;
            axis_tap<CM>(cx, a - RADIUS, P.ax[level], x0, wx0, wx1);
            const int rx = min(max(x0 - xs, 0), WIN_PITCH - 2);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const float g = __ldg(gq_ptr + (long long)(a * R + j) * P.N);
                const int ry = min(max(y0[j] - ylo, 0), WIN_ROWS - 2);
                float* w = wq + ry * WIN_PITCH + rx;
                const float gt = g * wy0[j], gb = g * wy1[j];
                atomicAdd(w, gt * wx0);
                atomicAdd(w + 1, gt * wx1);
                atomicAdd(w + WIN_PITCH, gb * wx0);
                atomicAdd(w + WIN_PITCH + 1, gb * wx1);
            }
        }
    }
    __syncthreads();

    // flush: only chunks that lie inside the map (out-of-bounds taps carry no gradient)
    const int Hl = P.H[level], Wl = P.W[level], Wp = P.Wp[level];
    float* base = P.gpyr + P.off[level];
    for (int idx = threadIdx.x; idx < QT * WIN_ROWS * 4; idx += LOOKUP_THREADS) {
        int q = idx / (WIN_ROWS * 4);
        int rem = idx - q * (WIN_ROWS * 4);
        int row = rem >> 2, chunk = rem & 3;
        if (row < desc.n_row[q] && chunk < desc.n_chunk[q]) {
            int y = desc.y_lo[q] + row;
            int x = desc.x_s[q] + 4 * chunk;
            if (y >= 0 && y < Hl && x >= 0 && x < Wl) {
                float4 v = *reinterpret_cast<const float4*>(win + win_base(q) + row * WIN_PITCH + 4 * chunk);
                // taps on pad columns [Wl, Wp) are out of bounds: keep the pads zero
                if (x + 1 >= Wl) v.y = 0.f;
                if (x + 2 >= Wl) v.z = 0.f;
                if (x + 3 >= Wl) v.w = 0.f;
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                    red_add_v4(base + desc.q_off[q] + tile_off(y, x, Wp), v);
            }
        }
    }
}

// ---------------------------------------------------------------- host side
template <int RADIUS>
static void launch_bwd(const LookupParams& P, int coord_mode, dim3 grid, cudaStream_t s) {
    if (coord_mode == FC_COORD_CUDA)
        lookup_bwd_kernel<RADIUS, FC_COORD_CUDA><<<grid, LOOKUP_THREADS, 0, s>>>(P);
    else
        lookup_bwd_kernel<RADIUS, FC_COORD_CPU><<<grid, LOOKUP_THREADS, 0, s>>>(P);
}

}  // namespace fc

using namespace fc;

extern "C" int fc_lookup_bwd(const float* grad_out, const float* coords, float* grad_pyramid,
                             int B, int H, int W, int num_levels, int radius,
                             int coord_mode, void* stream) {
    FC_REQUIRE(grad_out && coords && grad_pyramid, "fc_lookup_bwd: null pointer");
    Pyramid pyr;
    FC_REQUIRE(make_pyramid(pyr, B, H, W, num_levels), "fc_lookup_bwd: bad geometry B=%d H=%d W=%d L=%d", B, H, W, num_levels);
    if (int e = check_lookup_common(pyr, radius, coord_mode)) return e;
    LookupParams P{};
    fill_params(P, pyr, radius);
    P.pyr = nullptr; P.coords = coords;
    P.io = const_cast<float*>(grad_out); P.gpyr = grad_pyramid;
    dim3 grid((unsigned)((P.Q + QT - 1) / QT), (unsigned)pyr.L);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (radius) {
        case 1: launch_bwd<1>(P, coord_mode, grid, s); break;
        case 2: launch_bwd<2>(P, coord_mode, grid, s); break;
        case 3: launch_bwd<3>(P, coord_mode, grid, s); break;
        default: launch_bwd<4>(P, coord_mode, grid, s); break;
    }
    FC_LAUNCH_CHECK("lookup_bwd_kernel");
    return FC_OK;
}
