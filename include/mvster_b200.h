/*
 * mvster_b200.h - C ABI of libmvster_b200.so: the sm_100a implementation of the
 * MVSTER per-frame forward hot path (homography warp -> epipolar-Transformer
 * aggregation -> cost regularisation -> depth head).
 *
 * The reference (JeffWang987/MVSTER) is pure Python/PyTorch and has no FFI
 * layer, so there is no existing binding to mirror; each entry point below
 * names the reference Python function (file:line under the reference tree)
 * whose arithmetic it replaces.  The Python-side binding a maintainer adds is
 * shown in INTEGRATION.md (ctypes; mvster_b200/_lib.py is the working copy).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *    the caller owns all memory; the library never allocates device memory.
 *  - all tensors are dense fp32.  Feature maps are channels-last:
 *      feature  [B][H][W][C]          ("NHWC")
 *      cost     [B][D][H][W][G]       ("NDHWC")
 *      per-depth maps (hypo, logits, attn, wsum)   [B][D][H][W]
 *      per-pixel maps (depth, conf, inv_min, inv_max) [B][H][W]
 *  - every call is asynchronous on `stream` (a cudaStream_t / CUstream), is
 *    re-entrant, and keeps no state besides a thread-local error string and a
 *    monotonically increasing launch counter.
 *  - return value: 0 = launched, negative = error (see mvster_last_error()).
 */
#ifndef MVSTER_B200_H_
#define MVSTER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mvster_stream_t; /* cudaStream_t */

#define MVSTER_OK 0
#define MVSTER_ERR_ARG (-1)     /* unsupported shape / null pointer */
#define MVSTER_ERR_CUDA (-2)    /* CUDA launch/runtime error */

#define MVSTER_MAX_VIEWS 16     /* source views per mvster_et_fuse_f32 call */

/* flags of mvster_et_fuse_f32 */
#define MVSTER_ET_PARTIAL 1     /* write un-normalised acc + wsum (view-sharded run) */
#define MVSTER_ET_ACCUMULATE 2  /* start from the acc/wsum already in cost/wsum */
#define MVSTER_ET_GENERIC 4     /* force the generic G-lanes-per-pixel kernel (A/B testing) */
#define MVSTER_ET_NO_FUSE_D 8   /* attn_fuse_d=False: per-view scalar weight max_d softmax(sum_c cost) (mvs4net_utils.py:1049-1051) */
#define MVSTER_ET_SQDIFF 16     /* group_cor=False: cost[c] = (ref[c]-warped[c])^2, C cost channels (:1042); pass G == C */
#define MVSTER_ET_WINDOW 32     /* force the window kernel (correlate, then interpolate) where a specialisation exists */
#define MVSTER_ET_NO_WINDOW 64  /* never use the window kernel (A/B testing) */
#define MVSTER_ET_TMA_ON 128    /* force the TMA-staged window kernel (source boxes in shared memory) where a specialisation exists */
#define MVSTER_ET_TMA_OFF 256   /* never use it: window taps gathered from global memory (A/B testing) */
/* ref and src store every aligned block of 8 channels group-interleaved: position p of the block holds channel perm[p] with
 * perm = {0,2,1,3,4,6,5,7} for C/G = 2 and {0,4,1,5,2,6,3,7} for C/G = 4 (identity for C/G = 8), i.e. a 64-bit pair holds the same
 * channel index of two neighbouring groups and the group sums need no horizontal adds.  The producer of the features applies
 * the permutation for free (mvster_b200/fpn_engine.py permutes the output channels of the pyramid's last convolutions when
 * it packs their weights).  Honoured by the window kernels (TMA-staged or not) for (C, G, D) = (8,4,4), (16,4,4), (32,8,8);
 * any other kernel rejects it unless C/G = 8.  The cost volume's G channels keep their natural order. */
#define MVSTER_ET_INTERLEAVED 512

int mvster_version(void);                 /* 10000*major + 100*minor + patch */
const char* mvster_last_error(void);      /* thread-local, never NULL */
uint64_t mvster_launch_count(void);       /* kernels launched by this library so far (process-wide) */

/* ---- hypotheses -------------------------------------------------------- */
/* models/mvs4net_utils.py:71-77 init_inverse_range.  depth_values [B][n_dv]
 * (only column 0 and n_dv-1 are read) -> hypo [B][D][H][W], far -> near. */
int mvster_hypo_init_inverse_f32(const float* depth_values, int n_dv, float* hypo,
                                 int B, int D, int H, int W, mvster_stream_t stream);

/* models/mvs4net_utils.py:79-86 schedule_inverse_range.  inv_min/inv_max
 * [B][H/2][W/2] -> hypo [B][D][H][W] (bilinear x2, align_corners, reciprocal). */
int mvster_hypo_schedule_inverse_f32(const float* inv_min, const float* inv_max, float* hypo,
                                     int B, int D, int H, int W, mvster_stream_t stream);

/* Linear-depth twins (inverse_depth=False): models/mvs4net_utils.py:61-69 init_range and :88-99
 * schedule_range.  `depth` [B][H/2][W/2] is the previous stage's depth; the per-sample interval is
 * ratio * (depth_values[:, -1] - depth_values[:, 0]) / n_dv  (MVS4Net.py:61-63, :97). */
int mvster_hypo_init_linear_f32(const float* depth_values, int n_dv, float* hypo,
                                int B, int D, int H, int W, mvster_stream_t stream);
int mvster_hypo_schedule_linear_f32(const float* depth, const float* depth_values, int n_dv, float ratio, float* hypo,
                                    int B, int D, int H, int W, mvster_stream_t stream);

/* ---- relative pose ------------------------------------------------------- */
/* models/mvs4net_utils.py:1032-1035 (K @ E[:3,:4]) and :24 (src_proj @ inverse(ref_proj)).
 * proj [B][Nv][2][4][4] (slot 0 extrinsic, slot 1 intrinsic; view 0 = reference)
 * -> pose [B][V][12] (row-major R, then t) for source views first_view .. first_view+V-1. */
int mvster_pose_f32(const float* proj, float* pose, int B, int Nv, int first_view, int V,
                    mvster_stream_t stream);

/* ---- fused homography warp + epipolar-Transformer aggregation ----------- */
/* models/mvs4net_utils.py:13-59 (homo_warping) fused with :1037-1060 (group
 * correlation, softmax over D / temp / sqrt(C), running weighted sum over source
 * views, final division).
 *   ref       [B][H][W][C]
 *   src_host  HOST array of V device pointers, each [B][Hs][Ws][C]
 *   pose      [B][V][12]: row-major 3x3 R then t of  src_proj @ inverse(ref_proj)
 *             (the matrix of mvs4net_utils.py:24, computed by the caller)
 *   hypo      [B][D][H][W]
 *   cost      [B][D][H][W][G]   (normalised cost, or acc when MVSTER_ET_PARTIAL)
 *   wsum      [B][D][H][W]      (required with PARTIAL / ACCUMULATE, else may be NULL)
 * Supported: (C,G) in {(64,8),(32,8),(16,4),(8,4)} and any C==G*{1,2,4,8} with
 * G in {4,8}; D in {4,8}; 1 <= V <= MVSTER_MAX_VIEWS. */
int mvster_et_fuse_f32(const float* ref, const float* const* src_host, int V, const float* pose,
                       const float* hypo, float* cost, float* wsum,
                       int B, int C, int G, int D, int H, int W, int Hs, int Ws,
                       float attn_temp, int flags, mvster_stream_t stream);


/* bf16-storage form of the same fused kernel (BASELINE configs[2]): ref / src hold bf16 features, cost [B][D][H][W][G] is
 * written as bf16 (one rounding, on the way out); geometry, correlations, softmax and the weighted sums are the fp32
 * arithmetic of mvster_et_fuse_f32.  Halves the kernel's HBM traffic.  (C,G,D) in {(64,8,8),(32,8,8),(16,4,4),(8,4,4)};
 * flags: MVSTER_ET_WINDOW / NO_WINDOW / INTERLEAVED only (no partial sums: view sharding stays fp32). */
int mvster_et_fuse_bf16(const void* ref, const void* const* src_host, int V, const float* pose,
                        const float* hypo, void* cost, int B, int C, int G, int D, int H, int W, int Hs, int Ws,
                        float attn_temp, int flags, mvster_stream_t stream);
/* fp32 -> bf16, round to nearest even (the feature pyramid's outputs on their way into mvster_et_fuse_bf16); 16-byte aligned. */
int mvster_cast_bf16(const float* in, void* out, long long n, mvster_stream_t stream);

/* cost[b,d,y,x,g] = acc / (1e-8 + wsum[b,d,y,x]) in place: the division of
 * mvs4net_utils.py:1060 applied after the partials were all-reduced. */
int mvster_et_normalize_f32(float* cost, const float* wsum, int B, int G, int D, int H, int W,
                            mvster_stream_t stream);
/* Name of the kernel the last mvster_et_fuse_f32 call of this thread dispatched to (never NULL; "" before the first call):
 * what a benchmark prints next to the time it measured, instead of guessing from environment switches. */
const char* mvster_et_last_kernel(void);

/* Backward of mvster_et_fuse_f32 (group correlation, attn_fuse_d) for training: what autograd derives through
 * mvs4net_utils.py:1037-1060 and homo_warping :13-59.  The sampling grid is built under torch.no_grad() (:23) and the
 * hypotheses are detached (MVS4Net.py:95), so only the features receive gradients.  Inputs: the forward's operands,
 * its normalised output `cost` [B][D][H][W][G], `wsum` [B][D][H][W] = the sum over views of the weights WITHOUT the 1e-8
 * seed (what the forward writes with MVSTER_ET_PARTIAL, before mvster_et_normalize_f32), grad_cost [B][D][H][W][G].
 * Outputs: grad_ref [B][H][W][C] (overwritten); grad_src_host[v] [B][Hs][Ws][C] (HOST array of V device pointers, each
 * accumulated into with atomics - zero it first - or NULL when that view needs no gradient).  Nothing but cost and wsum
 * is saved between the passes: taps are gathered again and the per-view softmax is recomputed. */
int mvster_et_fuse_bwd_f32(const float* ref, const float* const* src_host, int V, const float* pose,
                           const float* hypo, const float* cost, const float* wsum, const float* grad_cost,
                           float* grad_ref, float* const* grad_src_host,
                           int B, int C, int G, int D, int H, int W, int Hs, int Ws,
                           float attn_temp, mvster_stream_t stream);

/* ---- regularisation ------------------------------------------------------ */
/* Generic channels-last 3-D convolution layer with folded BN:
 *   y = [relu](conv(x, w) + bias) [+ skip]
 * x [B][Di][Hi][Wi][Cin], w [kd*kh*kw][Cin][Cout], bias [Cout] (may be NULL),
 * y/skip [B][Do][Ho][Wo][Cout].  kernel (kd,3,3) with kd in {1,3}; pad = k/2;
 * stride (sd,s,s).  transposed != 0: ConvTranspose3d(k, stride, padding=k/2,
 * output_padding=stride-1), i.e. exact x2 up-sampling along strided axes.
 * Replaces ConvBnReLU3D (mvs4net_utils.py:116-123) and the ConvTranspose3d+BN+ReLU
 * sequences (:885-898, :926-940). */
int mvster_conv3d_ndhwc_f32(const float* x, const float* w, const float* bias, const float* skip, float* y,
                            int B, int Di, int Hi, int Wi, int Cin, int Cout,
                            int kd, int stride_d, int stride_hw, int transposed, int relu,
                            mvster_stream_t stream);

/* Generation 2 of the tensor-core layer: each 16x8-pixel tile (+halo) is staged once per depth plane and
 * the in-plane taps are addressed through the UMMA descriptor (no per-tap reload).  Cout may also be 8
 * (padded to N = 16).  w_packed: [hi | lo] x [kd*9][Cin/16][max(Cout,16)][16], packing.pack_tc2_weights. */
int mvster_conv3d_tc2_supported(int Cin, int Cout, int kd, int stride_hw, int transposed);
int mvster_conv3d_tc2_f32(const float* x, const float* w_packed, const float* bias, const float* skip, float* y,
                          int B, int D, int H, int W, int Cin, int Cout, int kd, int relu, int npass,
                          mvster_stream_t stream);

/* Generation 3 of the tensor-core layer (conv_tc3.cu): persistent kernel, error-compensated BF16 (three bf16 terms per
 * fp32 operand, six products in three MMAs per 16 channels: fp32-faithful, dropped terms <= 2^-24 relative), staged halo
 * tiles addressed through a per-layer stage/tap plan.  Replaces ConvBnReLU3D / ConvBnReLU (mvs4net_utils.py:116-123,
 * :224-251 `Conv2d` after BN folding) for:  stride 1, k in {1,3}, kd in {1,3};  stride (1,2,2), k in {3,5}, kd = 1 (pad = k/2).
 * Cin in {4,8,16,32,64}, Cout in {8,16,32,64}.  x [B][D][H][W][Cin] -> y [B][D][Ho][Wo][Cout], Ho = (H-1)/stride + 1.
 * w_packed: one 96*max(Cout,16)-byte slab per MMA slot in the order mvster_conv_tc3_plan reports (slabs[i], 6 ints each =
 * {kz, ky, kx, first input channel, ky2, kx2}; max_slabs counts slabs); slab = [2 K-halves][w1 | w2 | w3 rows of
 * max(Cout,16)][8 bf16].  Its 16 K rows are 16 input channels of tap (ky,kx) (zero padded) or, for Cin <= 8, 8 channels of
 * tap (ky,kx) then 8 channels of tap (ky2,kx2) (two taps per MMA; -1,-1 = none) - packing.pack_tc3_weights.
 * x and w_packed 16-byte aligned.
 * Second arithmetic: OR MVSTER_TC3_FP16X2 into the `relu` argument (bit 0 stays the ReLU switch) when w_packed holds two FP16
 * terms per weight, rows [w1 | w2 | unused] with w1 = fp16(w), w2 = fp16(2^11 (w - w1)) (packing.pack_tc3_weights(split=2)):
 * the activations are split the same way and the layer costs two MMAs per 16 channels instead of three; same fp32-class
 * accuracy (22-bit operands) for |x|, |w| < 65504. */
#define MVSTER_TC3_FP16X2 256
/* Third arithmetic, the bf16-storage configuration (BASELINE configs[2]; the reference itself has no bf16 path, its pixel grid is
 * hard-coded fp32, mvs4net_utils.py:28-29): OR MVSTER_TC3_BF16X1 into `relu` when w_packed holds ONE bf16 term per weight, rows
 * [w1 | unused | unused] (packing.pack_tc3_weights(split=1)).  The layer's fp32 input is rounded to bf16 (nearest even - what a
 * bf16 store applies) on its way into the MMA, products are exact, accumulation fp32: one MMA per 16 channels.  The _scaled
 * entry points multiply the accumulator by a per-output-channel fp32 factor before the bias (the BatchNorm scale, which
 * therefore stays out of the bf16 weights): y = [relu](scale * conv(bf16(x), w) + bias) [+ skip].  scale may be NULL. */
#define MVSTER_TC3_BF16X1 512
/* with MVSTER_TC3_BF16X1: round the stored output (after ReLU and the skip sum) to bf16 values - the layer's output as a bf16
 * build stores it, kept in its fp32 container */
#define MVSTER_TC3_ROUND_OUT 2048
/* The generation-3 kernel is persistent (one CTA per SM for the whole launch): the grid size is how much of the GPU a launch
 * claims.  mvster_set_sm_budget(n) caps the grid of the mvster_*_tc3_* launches that follow FROM THE CALLING THREAD (thread-local
 * like the error string; 0 = all SMs, the default): replicas driven by different host threads (nn.DataParallel) do not see each
 * other's cap.  The SM count is that of the calling thread's current device.  Used by the host to run the small
 * early cascade stages on a second stream next to the feature pyramid's large layers. */
void mvster_set_sm_budget(int n);
/* Range check of the two-fp16-term arithmetic (MVSTER_TC3_FP16X2): while a device word is registered, every such launch of the
 * CALLING THREAD sets it to 1 if one of the layer's input values does not fit the fp16 terms (|x| >= 65504 or non-finite; the
 * conversion saturates there).  NULL switches the check off again (then it costs nothing).  The host (engine.py) keeps it on for
 * the warm-up forwards of a new weight set / input signature and falls back to three bf16 terms (full fp32 range) if it fires. */
void mvster_tc3_set_overflow_flag(unsigned* device_flag);
unsigned* mvster_tc3_overflow_flag(void);   /* the calling thread's registered word (NULL = none) */
int mvster_conv_tc3_supported(int Cin, int Cout, int kd, int k, int stride_hw);
int mvster_conv_tc3_plan(int Cin, int kd, int k, int stride_hw, int* slabs, int max_slabs);
size_t mvster_conv_tc3_packed_bytes(int Cin, int Cout, int kd, int k, int stride_hw);
int mvster_conv_tc3_f32(const float* x, const void* w_packed, const float* bias, const float* skip, float* y,
                        int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu,
                        mvster_stream_t stream);
int mvster_conv_tc3_scaled_f32(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip,
                               float* y, int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int relu,
                               mvster_stream_t stream);

/* Transposed convolution on the same kernel: kernel (1,3,3), stride (1,2,2), padding 1, output padding 1 (Deconv3d after BN
 * folding: the ConvTranspose3d + BN + ReLU sequences of reg2d, mvs4net_utils.py:885-898) computed as a 2x2 stride-1 convolution on the input grid whose N columns are the
 * output parity classes [class (py,px)][Cout], scattered depth-to-space by the epilogue.  x [B][D][H][W][Cin] ->
 * y [B][D][2H][2W][Cout] (+ skip, same shape).  rows = -1: all four classes in one launch (needs 4*Cout <= 64); rows = 0 / 1:
 * only output rows of that parity (2*Cout <= 64), two launches cover the layer.  Cin in {16,32,64}, Cout in {8,16,32}.
 * w_packed: packing.pack_tc3_deconv_weights(w [9][Cin][Cout], rows), mvster_deconv_tc3_packed_bytes bytes. */
int mvster_deconv_tc3_supported(int Cin, int Cout, int rows);
size_t mvster_deconv_tc3_packed_bytes(int Cin, int Cout, int rows);
int mvster_deconv_tc3_f32(const float* x, const void* w_packed, const float* bias, const float* skip, float* y,
                          int B, int D, int H, int W, int Cin, int Cout, int rows, int relu, mvster_stream_t stream);
int mvster_deconv_tc3_scaled_f32(const float* x, const void* w_packed, const float* scale, const float* bias, const float* skip,
                                 float* y, int B, int D, int H, int W, int Cin, int Cout, int rows, int relu, mvster_stream_t stream);

/* reg2d U-Net (mvs4net_utils.py:870-912) up to, not including, the 1x1x1 `prob`
 * layer: cost [B][D][H][W][G] -> feat8 [B][D][H][W][8].  `blob` holds the folded
 * weights of conv0..conv11 in the layout reported by mvster_reg2d_layer_info;
 * `workspace` must hold mvster_reg2d_workspace_floats() floats. */
#define MVSTER_REG2D_LAYERS 10
size_t mvster_reg2d_blob_floats(int G);
size_t mvster_reg2d_workspace_floats(int B, int D, int H, int W);
/* info_host[8] = {Cin, Cout, kd, stride_hw, transposed, w_offset, bias_offset, n_taps} */
int mvster_reg2d_layer_info(int G, int layer, int64_t* info_host);
int mvster_reg2d_f32(const float* blob, const float* cost, float* feat8, float* workspace,
                     int B, int G, int D, int H, int W, mvster_stream_t stream);

/* Same network with conv2/conv4/conv6 on the tensor cores in TF32 arithmetic (mvster_conv3d_tc2_f32; npass = 3: 3xTF32,
 * 1: plain TF32): tc_blob = pack_tc2_weights(conv2) | (conv4) | (conv6) in the npass = 3 [hi|lo] layout
 * (mvster_reg2d_tc_blob_floats() floats); biases are read from `blob`.  kernel_gen must be 2 (the first-generation per-tap
 * kernel was removed in round 2). */
size_t mvster_reg2d_tc_blob_floats(void);
int mvster_reg2d_tc_f32(const float* blob, const float* tc_blob, const float* cost, float* feat8, float* workspace,
                        int B, int G, int D, int H, int W, int npass, int kernel_gen, mvster_stream_t stream);

/* Same network with conv1..conv6 (stride 1 and stride (1,2,2)) on the generation-3 tensor-core kernel
 * (mvster_conv_tc3_f32, 3 x bf16, fp32-faithful); conv0 (G -> 8 channels: K and N would be mostly padding) and the three
 * transposed layers stay on the CUDA cores.  tc3_blob = packing.pack_tc3_weights of conv0..conv6 back to back
 * (mvster_reg2d_tc3_blob_bytes(G) bytes, 16-byte aligned; conv0's slabs are present but unused); biases and the CUDA-core
 * layers' weights are read from `blob`. */
size_t mvster_reg2d_tc3_blob_bytes(int G);
/* flags: 0 or MVSTER_TC3_FP16X2 (tc3_blob packed with split=2); with it, MVSTER_REG2D_PACKED keeps the activations between the
 * layers as packed fp16 pairs (mvster_conv_tc3_pb16 with MVSTER_TC3_FP16X2: conv0 and every epilogue write the operand terms of
 * the next layer; no conversion pass) - same products, the skip sums see their addend to 22 bits (G in {4,8}, else ignored). */
#define MVSTER_REG2D_PACKED 4096
int mvster_reg2d_tc3_ex_f32(const float* blob, const void* tc3_blob, const float* cost, float* feat8, float* workspace,
                            int B, int G, int D, int H, int W, int flags, mvster_stream_t stream);
int mvster_reg2d_tc3_f32(const float* blob, const void* tc3_blob, const float* cost, float* feat8, float* workspace,
                         int B, int G, int D, int H, int W, mvster_stream_t stream);

/* Packed operands: the same two layers with activations that LIVE in HBM as the 16-bit operand terms of the arithmetic, in the
 * operand's own order: octet-planar [B*D][C/8][NT][H][W][8 channels].  NT = 1 (default): bf16 ("PB16", the bf16-storage
 * configuration; for C = 8 this is plain NHWC bf16).  NT = 2 (MVSTER_TC3_FP16X2 in `flags`): the fp16 pair (a1, a2) with
 * a = a1 + 2^-11 a2 of the two-fp16-term arithmetic - 32 bits per element like fp32, split once by the PRODUCING layer instead of
 * by every consumer (w_packed then holds two fp16 terms, split = 2).  A TMA box lands a stage's halo tile directly in the MMA
 * operand ring: no fp32 staging, no conversion pass (and half the activation bytes with bf16).  x and skip (may be NULL; shape of
 * the output) are packed; y is packed with MVSTER_TC3_OUT_PB16 in `flags` (rounded / split after ReLU and the skip sum), else fp32
 * NDHWC.  flags bit 0 = ReLU.  scale as in the _scaled entry points.  Cin in {8,16,32,64}, Cout in {8,16,32,64}. */
#define MVSTER_TC3_OUT_PB16 1024
int mvster_conv_tc3_pb16(const void* x, const void* w_packed, const float* scale, const float* bias, const void* skip, void* y,
                         int B, int D, int H, int W, int Cin, int Cout, int kd, int k, int stride_hw, int flags, mvster_stream_t stream);
int mvster_deconv_tc3_pb16(const void* x, const void* w_packed, const float* scale, const float* bias, const void* skip, void* y,
                           int B, int D, int H, int W, int Cin, int Cout, int rows, int flags, mvster_stream_t stream);

/* bf16-storage reg2d (BASELINE configs[2]): cost_bf16 [B][D][H][W][G] as bf16 (what mvster_et_fuse_bf16 writes) -> feat8 fp32.
 * Every convolution computes conv(bf16(x), bf16(w)) with fp32 accumulation, then the BatchNorm scale and shift in fp32, ReLU and
 * the skip sum in fp32 - i.e. conv -> BN -> ReLU of mvs4net_utils.py:116-123 with bf16 rounding at the convolution's operands
 * only.  `blob_q`: the layout of mvster_reg2d_layer_info with the UNFOLDED weights rounded to bf16 (stored as fp32) and the
 * BatchNorm shift in the bias slots; `scales`: the BatchNorm factors of conv0..conv11 back to back (sum of Cout = 288 floats);
 * `tc3_blob`: the slab streams of conv0..conv11 in the order of mvster_reg2d_tc3_ex_f32 packed with split = 1.  conv0 runs on
 * the CUDA cores (bf16 loads, fp32 FMA), conv1..conv11 on the tcgen05 kernel with MVSTER_TC3_BF16X1. */
#define MVSTER_REG2D_SCALE_FLOATS 288
/* flags: MVSTER_REG2D_BF16_PACKED keeps the activations between the layers as packed bf16 operands (mvster_conv_tc3_pb16 /
 * mvster_deconv_tc3_pb16: no conversion pass, half the bytes); 0 keeps them as bf16-rounded values in fp32 containers
 * (mvster_conv_tc3_scaled_f32 with MVSTER_TC3_BF16X1 | MVSTER_TC3_ROUND_OUT).  Same arithmetic either way. */
#define MVSTER_REG2D_BF16_PACKED 1
int mvster_reg2d_bf16(const float* blob_q, const void* tc3_blob, const float* scales, const void* cost_bf16, float* feat8,
                      float* workspace, int B, int G, int D, int H, int W, int flags, mvster_stream_t stream);

/* reg3d U-Net (mvs4net_utils.py:914-965; optional `--reg_mode reg3d`): 3x3x3 kernels, stride 2 along D too,
 * down_size in {1,2,3} (MVS4Net.py:48), prob = 3x3x3 conv 8->1 without bias.  cost [B][D][H][W][G] ->
 * logits [B][D][H][W].  Blob layout: the existing layers in network order (mvster_reg3d_layer_info:
 * info_host[8] = {Cin, Cout, stride, transposed, w_offset, bias_offset or -1, 27, layer id}); weights
 * [27][Cin][Cout] with BN folded.  CUDA-core fp32 kernels. */
int mvster_reg3d_num_layers(int down_size);
int mvster_reg3d_layer_info(int G, int down_size, int layer, int64_t* info_host);
size_t mvster_reg3d_blob_floats(int G, int down_size);
size_t mvster_reg3d_workspace_floats(int B, int D, int H, int W);
int mvster_reg3d_f32(const float* blob, const float* cost, float* logits, float* workspace,
                     int B, int G, int D, int H, int W, int down_size, mvster_stream_t stream);

/* ---- feature pyramid (FPN4, mvs4net_utils.py:419-502; outside the named hot path, it feeds it) ------- */
/* NHWC 2-D convolution, k in {1,3,5}, stride 1/2, pad k/2, optional bias (folded BN) and ReLU, CUDA cores.
 * x [N][H][W][Cin], w [k*k][Cin][Cout] -> y [N][Ho][Wo][Cout].  Conv2d blocks of mvs4net_utils.py:224-251. */
int mvster_conv2d_nhwc_f32(const float* x, const float* w, const float* bias, float* y,
                           int N, int H, int W, int Cin, int Cout, int k, int stride, int relu, mvster_stream_t stream);
/* First layer: 3-channel NCHW image [N][3][H][W] -> NHWC [N][H][W][8]; 3x3, bias, ReLU (conv0.0, :424). w [9][3][8]. */
int mvster_conv_first_f32(const float* img_nchw, const float* w, const float* bias, float* y,
                          int N, int H, int W, mvster_stream_t stream);
/* Top-down merge (:479-486): out = bilinear_x2(top, align_corners=True) + conv1x1(lateral) + bias.
 * top [N][H/2][W/2][64], lateral [N][H][W][Clat], w [Clat][64], out [N][H][W][64]. */
int mvster_fpn_merge_f32(const float* top, const float* lateral, const float* w, const float* bias, float* out,
                         int N, int H, int W, int Clat, mvster_stream_t stream);

/* Point-wise (1x1) convolution as a tensor-core GEMM [pixels x Cin] x [Cin x Cout] (conv_tc2.cu, centre tap only).
 * Cin in {16,32,64}; Cout in {8,16,32,64} or 68..80 (N padded to 80).  w_packed = pack_tc2_weights([1][Cin][Cout]). */
int mvster_pointwise_tc2_f32(const float* x, const float* w_packed, const float* bias, float* y,
                             int N, int H, int W, int Cin, int Cout, int relu, int npass, mvster_stream_t stream);
/* Point-wise convolution on the generation-3 kernel whose output channel blocks go to separate tensors: channels
 * [j*block, (j+1)*block) -> [N][H][W][block] at y + j*block_stride_floats (no bias / relu).  Cin in {16,32,64}, Cout in
 * {8,16,32,64}; w_packed = packing.pack_tc3_weights([1][Cin][Cout], 1, 1, 1).  Used to produce the per-tap planes of U below. */
int mvster_pointwise_tc3_blocks_ex_f32(const float* x, const void* w_packed, float* y, int N, int H, int W, int Cin, int Cout,
                                       int block, long long block_stride_floats, int flags /* 0 or MVSTER_TC3_FP16X2 */,
                                       mvster_stream_t stream);
int mvster_pointwise_tc3_blocks_f32(const float* x, const void* w_packed, float* y, int N, int H, int W, int Cin, int Cout,
                                    int block, long long block_stride_floats, mvster_stream_t stream);
/* Last pyramid level fused (mvs4net_utils.py:485-486, stage4 = out4(up2(top2) + inner3(c0))) without forming the
 * 64-channel full-resolution map: U = per-tap 1x1 conv of top2, either interleaved [N][H/2][W/2][u_channels >= 72] (channels
 * tap*8+o) or, with u_channels == 8, planar [9][N][H/2][W/2][8];
 * c0 [N][H][W][8], w_comp [9][8][8] = W4[tap] Wi3, b_tap [9][8] = W4[tap] bi3  ->  out [N][H][W][8]. */
int mvster_fpn_out4_gather_f32(const float* U, int u_channels, const float* c0, const float* w_comp, const float* b_tap,
                               float* out, int N, int H, int W, mvster_stream_t stream);

/* ---- head ---------------------------------------------------------------- */
/* mvs4net_utils.py:1066-1088.  Either `logits` [B][D][H][W] is given, or
 * (feat8 [B][D][H][W][8], prob_w[8], prob_b[1]) and the 1x1x1 `prob` conv of
 * mvs4net_utils.py:900 is applied on the fly.  Outputs (any may be NULL):
 *   attn [B][D][H][W] softmax over D;  depth [B][H][W] = hypo[argmax] (first max);
 *   conf [B][H][W] = max prob (LOW resolution; see mvster_upsample_bilinear_f32);
 *   inv_min/inv_max = 1/depth +- split_itv*(1/hypo[2]-1/hypo[1]);
 *   soft_depth = sum_d attn*hypo (models/module.py:935-941, optional extra). */
int mvster_head_f32(const float* logits, const float* feat8, const float* prob_w, const float* prob_b,
                    const float* hypo, float* attn, float* depth, float* conf,
                    float* inv_min, float* inv_max, float* soft_depth,
                    int B, int D, int H, int W, float split_itv, mvster_stream_t stream);

/* Same with flags: MVSTER_HEAD_BF16_INPUT rounds feat8 to bf16 on the way into the `prob` layer (bf16-storage configuration;
 * prob_w then holds bf16 values). */
#define MVSTER_HEAD_BF16_INPUT 1
int mvster_head_ex_f32(const float* logits, const float* feat8, const float* prob_w, const float* prob_b,
                       const float* hypo, float* attn, float* depth, float* conf,
                       float* inv_min, float* inv_max, float* soft_depth,
                       int B, int D, int H, int W, float split_itv, int flags, mvster_stream_t stream);

/* F.interpolate(mode='bilinear', align_corners=True) on [B][H][W] -> [B][H*f][W*f]
 * (confidence up-sampling of mvs4net_utils.py:1076-1077). */
int mvster_upsample_bilinear_f32(const float* in, float* out, int B, int H, int W, int factor,
                                 mvster_stream_t stream);

/* NCHW [B][C][H][W] -> NHWC [B][H][W][C] (feeds reference-layout features to the path). */
int mvster_nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W,
                            mvster_stream_t stream);

/* ---- training loss ----------------------------------------------------------- */
/* models/mvs4net_utils.py:1096-1142 `sinkhorn` (the optimal-transport term of MVS4net_loss, MVS4Net.py:149), forward and backward
 * fused: one thread per pixel runs the `iters` log-domain Sinkhorn iterations on its own D x D (continuous != 0: D x (D+1)) ground
 * cost, writes the pixel's transport cost sum_ij T_ij cost_ij into loss_px [B][H][W] (0 where mask == 0) and, if grad_attn is not
 * NULL, d loss_px / d attn into grad_attn [B][D][H][W] by reverse-mode differentiation of its iteration history (0 where
 * mask == 0).  The loss of the reference is the mean of loss_px over the masked pixels.  gt_depth [B][H][W], hypo / attn
 * [B][D][H][W], mask [B][H][W] as bytes (torch.bool).  D in {4, 8}; 0 <= iters <= 32. */
int mvster_sinkhorn_f32(const float* gt_depth, const float* hypo, const float* attn, const unsigned char* mask,
                        float* loss_px, float* grad_attn, int B, int D, int H, int W, int iters, float eps, int continuous,
                        mvster_stream_t stream);

/* ---- geometric-consistency filter (after the forward) ---------------------- */
/* test_mvs4.py:271-328 (reproject_with_depth + check_geometric_consistency) for one (reference, source) pair, and the
 * accumulation of filter_depth :362-378.  depth_ref [H][W], depth_src [Hs][Ws] (device, fp32).  mats: HOST array of 60 doubles =
 * inv(K_ref) 3x3 | (E_src inv(E_ref)) rows 0-2 (3x4) | K_src | inv(K_src) | (E_ref inv(E_src)) rows 0-2 | K_ref, row-major,
 * computed by the caller the way the reference computes them (float32 numpy).  Outputs (device): depth_reproj [H][W]
 * (zero where |p_reproj - p| >= dist_thres pixels or |d_reproj - d| / d >= rel_thres); optional x_src, y_src [H][W] (the
 * float32 source-pixel coordinates), mask [H][W] (uint8); optional accumulators mask_sum (int32, += mask) and depth_sum
 * (fp32, += depth_reproj).  The source depth is sampled like cv2.remap(INTER_LINEAR): coordinates rounded to 1/32 pixel,
 * zeros outside the image. */
int mvster_geo_consistency_f32(const float* depth_ref, const float* depth_src, const double* mats, float* depth_reproj,
                               float* x_src, float* y_src, unsigned char* mask, int* mask_sum, float* depth_sum,
                               int H, int W, int Hs, int Ws, float dist_thres, float rel_thres, mvster_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVSTER_B200_H_ */
