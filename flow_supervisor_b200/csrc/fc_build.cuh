// Shared between the tensor-core build (fc_build_tc.cu) and the fused fnet tail (fc_feat.cu): where the packed
// K-major bf16 operands of the build live, and who fills them.
#pragma once

#include "fc_umma.cuh"

namespace fc {

// destination of a packer: fmap1 rows = queries p (scaled by `prescale`), fmap2 rows = PADDED targets in patch
// order (pad rows zero); hi = bf16(x), lo = bf16(x - hi) (lo == nullptr: single-pass bf16 mode)
struct TcPacked {
    __nv_bfloat16* a_hi; __nv_bfloat16* a_lo;       // [B * N ][D]
    __nv_bfloat16* b_hi; __nv_bfloat16* b_lo;       // [B * NP][D]
    int B, D, N, NP, H, W, Wp;
    float prescale;
};

// the source of the fused fnet tail: the activations in front of the encoder's 1x1 output
// convolution and that convolution's pre-packed weights (fc_fnet_tail_prepare)
struct FeatSource {
    const float* x;            // (2B, C, H, W) fp32: frames of image 1, then frames of image 2 (extractor.py:170-172)
    const void* packed_w;      // [w_hi D*C bf16][w_lo D*C bf16][bias D fp32]
    int C;
};
int fnet_tail_pack(const FeatSource& src, const TcPacked& dst, cudaStream_t s);   // fc_feat.cu

}  // namespace fc
