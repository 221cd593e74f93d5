// Convex 8x upsampling of the flow field (RAFT.upsample_flow,
// /root/reference/pytorch/core/raft.py:72-83, twin gma_network.py:60-72):
//
//   mask (B, 9*8*8, H, W) -> softmax over the 9 neighbours;  up_flow = unfold(8 * flow, 3x3, padding 1)
//   out[b, c, 8h + i, 8w + j] = sum_k softmax_k(mask[b, k*64 + i*8 + j, h, w]) * 8 * flow[b, c, h + k/3 - 1, w + k%3 - 1]
//
// The reference materialises the softmax (B x 576 x H x W), the unfolded flow, their product and a
// permuted copy: five full passes over 16 MB per Sintel-size sample.  Here the mask is read once
// (coalesced along w for every channel) and the 2 x 8H x 8W result written once: HBM-bound at
// (576 + 128) * 4 bytes per token.  Thread = (token, sub-row i): 72 mask values, 8 sub-columns.
#include "fc_common.cuh"

namespace fc {

__global__ void __launch_bounds__(256) upsample_flow_kernel(const float* __restrict__ flow, const float* __restrict__ mask,
                                                            float* __restrict__ out, int H, int W) {
    const int w = blockIdx.x * 32 + threadIdx.x, h = blockIdx.y, b = blockIdx.z, i = threadIdx.y;
    if (w >= W) return;
    const long long HW = (long long)H * W;
    const float* m = mask + ((long long)b * 576 + i * 8) * HW + (long long)h * W + w;
    // 3x3 neighbourhood of 8 * flow, zero padded (F.unfold(..., padding=1))
    float f[2][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int hh = h + k / 3 - 1, ww = w + k % 3 - 1;
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
        for (int c = 0; c < 2; ++c)
            f[c][k] = in ? 8.0f * __ldg(flow + ((long long)b * 2 + c) * HW + (long long)hh * W + ww) : 0.f;
    }
    float o[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float x[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) x[k] = __ldg(m + ((long long)k * 64 + j) * HW);
        float mx = x[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) mx = fmaxf(mx, x[k]);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) { x[k] = expf(x[k] - mx); s += x[k]; }
        const float inv = 1.0f / s;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float p = x[k] * inv;
            a0 = fmaf(p, f[0][k], a0);
            a1 = fmaf(p, f[1][k], a1);
        }
        o[0][j] = a0; o[1][j] = a1;
    }
    const long long W8 = 8LL * W;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float* dst = out + (((long long)b * 2 + c) * 8 * H + (8LL * h + i)) * W8 + 8LL * w;
        *reinterpret_cast<float4*>(dst) = make_float4(o[c][0], o[c][1], o[c][2], o[c][3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[c][4], o[c][5], o[c][6], o[c][7]);
    }
}

}  // namespace fc

using namespace fc;

extern "C" int fc_upsample_flow(const float* flow, const float* mask, float* out, int B, int H, int W, void* stream) {
    FC_REQUIRE(flow && mask && out, "fc_upsample_flow: null pointer");
    FC_REQUIRE(B >= 1 && H >= 1 && W >= 1 && B <= 65535 && H <= 65535, "fc_upsample_flow: bad geometry B=%d H=%d W=%d", B, H, W);
    FC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "fc_upsample_flow: out must be 16-byte aligned");
    dim3 grid((unsigned)((W + 31) / 32), (unsigned)H, (unsigned)B), block(32, 8);
    upsample_flow_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(flow, mask, out, H, W);
    FC_LAUNCH_CHECK("upsample_flow_kernel");
    return FC_OK;
}
