/*
 * flowcorr.h -- C ABI of libflowcorr.so: RAFT's correlation hot path for B200 (sm_100a).
 *
 * This is the drop-in boundary for the one path of iwbn/flow-supervisor this
 * repository replaces: pytorch/core/corr.py (CorrBlock / AlternateCorrBlock), its
 * GMA twin pytorch/core/gma_corr.py, and the reference's only native module
 * pytorch/alt_cuda_corr/.  Every entry point cites the reference code it stands in
 * for (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch / CUDA types in signatures
 *     (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - the caller owns every buffer; device pointers must be 16-byte aligned;
 *   - every call is asynchronous with respect to the host, launches on `stream`,
 *     allocates nothing and never synchronises (CUDA-graph capturable);
 *   - return 0 on success, a negative FC_E* code otherwise; fc_last_error() gives
 *     a thread-local human readable message.  Nothing throws or aborts;
 *   - thread safe: results depend on the arguments only.  The process-wide state is a set of
 *     mutex-guarded memo caches (tensor-map encodings keyed by pointer + geometry, per-device
 *     kernel attributes, the environment switches read once), so concurrent calls from several
 *     host threads / devices are fine (nn.DataParallel replicas, pytorch/train.py:183).
 *
 * Correlation pyramid layout (internal to this library; the reference keeps a list
 * of (B*N, 1, Hl, Wl) tensors, corr.py:14-27, which nothing outside corr.py reads):
 *   one buffer, level l at byte offset level_offset[l], holding for every query
 *   b*N + p (p < N = H*W) a map of Hp_l x Wp_l elements, Hl = H >> l, Wl = W >> l (floor,
 *   like avg_pool2d), Wp_l = round_up(Wl, 8), Hp_l = round_up(Hl, 2), stored in 64-byte
 *   patches of 2 rows x 8 columns (the DRAM access granule), patches of a row pair left
 *   to right:  offset(y, x) = (y/2)*(2*Wp_l) + (x/8)*16 + (y%2)*8 + x%8.
 *   INVARIANT: every pad row (y >= Hl) and pad column (x >= Wl) of EVERY level holds exact zeros.
 *   fc_build establishes it; the lookup kernels read whole patches / row pairs and rely on it for the
 *   reference's padding_mode='zeros' at the right / bottom edge.  A caller filling a pyramid itself must
 *   zero the pads.
 *   Element type: fp32 (FC_VOL_F32) or bf16 (FC_VOL_BF16).
 */
#ifndef FLOWCORR_H_
#define FLOWCORR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FC_ABI_VERSION 1
#define FC_MAX_LEVELS 6
#define FC_MAX_RADIUS 4

/* error codes */
#define FC_OK 0
#define FC_EINVAL (-1)   /* bad argument (shape, alignment, null pointer, unsupported size) */
#define FC_ECUDA (-2)    /* a CUDA runtime / driver call failed (see fc_last_error) */
#define FC_EWORKSPACE (-3) /* workspace too small */

/* volume element type */
#define FC_VOL_F32 0
#define FC_VOL_BF16 1

/* arithmetic of the all-pairs contraction (fc_build / fc_build_bwd) */
#define FC_MATH_FP32 0      /* CUDA-core fp32 FMA: the reference's SGEMM arithmetic        */
#define FC_MATH_TC_3XBF16 1 /* tcgen05 bf16 split hi/lo, 3 MMAs, fp32 accumulate (~4e-6)    */
#define FC_MATH_TC_BF16 2   /* tcgen05 bf16 inputs, fp32 accumulate (~2e-3, stated bf16 mode)*/

/* how utils.py:61-62's `2*x/(W-1)` is rounded (decides the integer tap index):
 * CUDA tensors multiply by fl(1/(W-1)), CPU tensors divide. */
#define FC_COORD_CUDA 0
#define FC_COORD_CPU 1
/* AlternateCorrBlock's indexing (corr.py:85, correlation_kernel.cu:67-76): tap = floor(c) + offset,
 * weights (1 - frac(c), frac(c)), no normalise / un-normalise round trip.  fc_lookup_fwd only:
 * lets a materialised pyramid answer on-demand lookups with the on-demand kernel's indices. */
#define FC_COORD_RAW 2

int fc_abi_version(void);
const char* fc_last_error(void);

/* Diagnostic switches.  They are read from the environment ONCE per process (FLOWCORR_BUILD_SCHED,
 * FLOWCORR_BUILD_STAGES, FLOWCORR_BUILD_EPI_WARPS, FLOWCORR_NO_FUSE, FLOWCORR_VERBOSE, FLOWCORR_L2_FETCH); this call
 * overrides one at run time (names: build_sched, build_stages, build_epi_warps, no_fuse, verbose, l2_fetch).
 * None changes results -- only which kernel variant computes them (tests/test_gpu_tensorcore.py). */
int fc_tunable_set(const char* name, int value);

/* Number of CUDA kernels this library has launched in the calling process so far
 * (monotonic; instrumentation for benchmarks: bench.py's gpu_launches). */
unsigned long long fc_kernel_launches(void);

/* Geometry of level `level` for an H x W (1/8-resolution) token grid.
 * Replaces: the implicit shapes of corr.py:24-27 (F.avg_pool2d(corr, 2, stride=2)). */
int fc_level_dims(int H, int W, int level, int* Hl, int* Wl, int* Wp);

/* Total bytes of the pyramid buffer and (optional, may be NULL) the byte offset of
 * each level inside it.  Returns 0 on bad arguments. */
size_t fc_pyramid_bytes(int B, int H, int W, int num_levels, int vol_dtype,
                        size_t* level_offsets /* [num_levels] or NULL */);

/* Scratch needed by fc_build for the given problem and math mode. */
size_t fc_build_workspace_bytes(int B, int D, int H, int W, int num_levels, int math);

/* CorrBlock.__init__  (corr.py:13-27; CorrBlock.corr corr.py:52-60; gma_corr.py:15-63)
 *   pyramid level 0 [b,p,q] = sum_d fmap1[b,d,p] * fmap2[b,d,q] / sqrt(D), then
 *   num_levels-1 successive 2x2 means, all in one pass (no separate /sqrt(D) or
 *   avg_pool2d kernels).
 * fmap1, fmap2: (B, D, H, W) fp32 contiguous (NCHW, as raft.py:102-107 passes them). */
int fc_build(const float* fmap1, const float* fmap2, void* pyramid,
             int B, int D, int H, int W, int num_levels,
             int vol_dtype, int math,
             void* workspace, size_t workspace_bytes, void* stream);

/* CorrBlock.__call__  (corr.py:29-50 + utils.py:57-65 bilinear_sampler + ATen
 * grid_sample(bilinear, zeros, align_corners=True)), all levels in one launch.
 *   coords: (B, 2, H, W) fp32, channel 0 = x, channel 1 = y (utils.py:74-77)
 *   out:    (B, num_levels*(2r+1)^2, H, W) fp32, channel = l*(2r+1)^2 + a*(2r+1) + b
 *           with a the x-offset index and b the y-offset index (corr.py:37-39).
 * Optional debug outputs (NULL to skip) expose the bit-exact integer part:
 *   dbg_x0, dbg_y0: (B*N, num_levels, 2r+1) int32 floor indices per axis tap;
 *   dbg_mask:       (B*N, num_levels, 2r+1 [a], 2r+1 [b]) uint8, bit0 = (y0,x0),
 *                   bit1 = (y0,x0+1), bit2 = (y0+1,x0), bit3 = (y0+1,x0+1) in bounds. */
int fc_lookup_fwd(const void* pyramid, const float* coords, float* out,
                  int B, int H, int W, int num_levels, int radius,
                  int vol_dtype, int coord_mode,
                  int32_t* dbg_x0, int32_t* dbg_y0, uint8_t* dbg_mask, void* stream);

/* Backward of one lookup w.r.t. the volume (autograd of corr.py:29-50; implied by
 * pytorch/train.py:273,277).  Scatter-ADDS into grad_pyramid (fp32, same layout as
 * an FC_VOL_F32 pyramid), which the caller zeroes once per CorrBlock: all lookups
 * of a block accumulate into the same buffer.  coords get no gradient (they arrive
 * detached, raft.py:123). */
int fc_lookup_bwd(const float* grad_out, const float* coords, float* grad_pyramid,
                  int B, int H, int W, int num_levels, int radius,
                  int coord_mode, void* stream);

/* Scratch needed by fc_build_bwd (0 for FC_MATH_FP32 and for shapes the tensor-core
 * backward does not take: those run the fp32 CUDA-core contractions). */
size_t fc_build_bwd_workspace_bytes(int B, int D, int H, int W, int num_levels, int math);

/* Backward of CorrBlock.__init__ (autograd of corr.py:21-27,52-60): folds the
 * gradient pyramid to level 0 (avg_pool2d backward), scales by 1/sqrt(D) and contracts:
 *   dfmap1[b,:,p] = sum_q dC[b,p,q] fmap2[b,:,q],  dfmap2[b,:,q] = sum_p dC[b,p,q] fmap1[b,:,p].
 * dfmap1 / dfmap2: (B, D, H, W) fp32, overwritten; either may be NULL to skip.
 * grad_pyramid must be treated as consumed: the tensor-core modes' default kernels only read it (the fold
 * and the bf16 split happen inside the GEMMs), but FC_MATH_FP32 and the FLOWCORR_BWD_FUSED=0 pipeline fold
 * it in place.  Its pad cells must be zero (fc_lookup_bwd never writes them). */
int fc_build_bwd(float* grad_pyramid, const float* fmap1, const float* fmap2,
                 float* dfmap1, float* dfmap2,
                 int B, int D, int H, int W, int num_levels, int math,
                 void* workspace, size_t workspace_bytes, void* stream);

/* AlternateCorrBlock.__call__ (corr.py:74-91): on-demand correlation, no volume.
 * Pools fmap2 (corr.py:68-72) into `workspace`, then for every level computes the
 * (2r+2)^2 dot products around floor(coords / 2^l) and their bilinear splat
 * (correlation_kernel.cu:59-116), divides by sqrt(D) (corr.py:91).
 * out: (B, num_levels*(2r+1)^2, H, W) fp32. */
size_t fc_ondemand_workspace_bytes(int B, int D, int H, int W, int num_levels);
int fc_ondemand_prepare(const float* fmap1, const float* fmap2,
                        int B, int D, int H, int W, int num_levels,
                        void* workspace, size_t workspace_bytes, void* stream);
int fc_ondemand_fwd(const void* workspace, const float* coords, float* out,
                    int B, int D, int H, int W, int num_levels, int radius, void* stream);

/* alt_cuda_corr.forward / backward (pytorch/alt_cuda_corr/correlation.cpp:23-54,
 * kernels correlation_kernel.cu:18-119 and :122-256), one pyramid level per call,
 * channels-last inputs exactly as corr.py:82-86 passes them:
 *   fmap1 (B,H1,W1,C), fmap2 (B,H2,W2,C), coords (B,1,H1,W1,2) [x,y],
 *   corr (B,1,(2r+1)^2,H1,W1), unscaled (the caller divides by sqrt(C), corr.py:91).
 * Backward: fmap1_grad / fmap2_grad overwritten; coords get no gradient (the
 * reference allocates coords_grad and never writes it, correlation_kernel.cu:307). */
int fc_altcorr_fwd(const float* fmap1, const float* fmap2, const float* coords, float* corr,
                   int B, int H1, int W1, int H2, int W2, int C, int radius, void* stream);
int fc_altcorr_bwd(const float* fmap1, const float* fmap2, const float* coords,
                   const float* corr_grad, float* fmap1_grad, float* fmap2_grad,
                   int B, int H1, int W1, int H2, int W2, int C, int radius, void* stream);

/* ---- adjacent components: opt-in, the reference scripts run without them ---- */

/* RAFT.upsample_flow (pytorch/core/raft.py:72-83; gma_network.py:60-72): convex 8x upsampling,
 *   out[b,c,8h+i,8w+j] = sum_k softmax_k(mask[b, k*64+i*8+j, h, w]) * 8 * flow[b,c,h+k/3-1,w+k%3-1]  (zero padded)
 * flow (B,2,H,W), mask (B,576,H,W), out (B,2,8H,8W), all fp32 contiguous.  One pass over the mask. */
int fc_upsample_flow(const float* flow, const float* mask, float* out, int B, int H, int W, void* stream);

/* Lookup fused into the motion encoder's first convolution (pytorch/core/update.py:83,90: `F.relu(self.convc1(corr))`
 * applied to `corr = corr_fn(coords1)`, raft.py:124): out[b,co,p] = relu(bias[co] + sum_k W[co,k] * lookup[b,k,p]).
 * The (B, 324, H, W) tensor never reaches HBM; the weights live in tensor memory for the whole kernel.
 *   fc_convc1_prepare     once per model: weight (256, 324) fp32 [+ bias (256), may be NULL] -> fc_convc1_weights_bytes() bytes
 *   fc_lookup_convc1_fwd  pyramid + coords (B, 2, H, W) -> out (B, 256, H, W) fp32
 *   fc_lookup_convc1_supported  num_levels == 4, radius == 4, 256 output channels (RAFT / GMA basic models). */
size_t fc_convc1_weights_bytes(void);
int fc_lookup_convc1_supported(int num_levels, int radius, int out_channels);
int fc_convc1_prepare(const float* weight, const float* bias, void* packed, size_t packed_bytes, void* stream);
int fc_lookup_convc1_fwd(const void* pyramid, const float* coords, const void* packed_weights, float* out,
                         int B, int H, int W, int num_levels, int radius, int vol_dtype, int coord_mode, void* stream);

/* Fused tail of the feature encoder (pytorch/core/extractor.py:145,184 `conv2`, the 1x1 output convolution of `fnet`;
 * raft.py:99-107): the convolution runs on the tensor cores and its epilogue writes the build's packed K-major bf16
 * hi/lo operands directly, so the fp32 feature maps and the pack pre-pass of fc_build do not exist.
 *   fc_fnet_tail_prepare   once per model: weight (D, C) fp32 [+ bias (D), may be NULL] -> packed buffer of
 *                          fc_fnet_tail_weights_bytes(C, D) bytes (128-byte aligned)
 *   fc_build_from_fnet_tail  x (2B, C, H, W) fp32 = the activations in front of conv2, frames of image 1 then image 2
 *                          (extractor.py:170-172 concatenates them) -> the pyramid, exactly as fc_build would produce it
 *                          from fmap = conv2(x).  Tensor-core math modes only; workspace = fc_build_workspace_bytes.
 *   fc_fnet_tail_supported C in {64, 128}, D % 64 == 0, D <= 256, H*W % 4 == 0 (else run conv2 + fc_build). */
size_t fc_fnet_tail_weights_bytes(int C, int D);
int fc_fnet_tail_supported(int C, int D, int H, int W);
int fc_fnet_tail_prepare(const float* weight, const float* bias, int C, int D, void* packed, size_t packed_bytes, void* stream);
int fc_build_from_fnet_tail(const float* x, const void* packed_weights, void* pyramid,
                            int B, int C, int D, int H, int W, int num_levels, int vol_dtype, int math,
                            void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWCORR_H_ */
