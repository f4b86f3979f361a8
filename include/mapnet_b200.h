/* mapnet_b200.h -- C ABI of the B200-native MapNet/PoseNet training hot path.
 *
 * Drop-in boundary.  The reference (NVlabs/geomapnet) is pure
 * Python on PyTorch: it has NO native FFI for this path, so these entry points
 * are the contract a maintainer binds (ctypes stub in INTEGRATION.md) behind the
 * reference's own Python surface:
 *
 *   reference call site                               entry point that replaces it
 *   -----------------------------------------------   ------------------------------
 *   models/posenet.py:65-73  PoseNet.forward          mapnet_forward
 *   models/posenet.py:93-97  MapNet.forward           mapnet_forward (T folded into B)
 *   common/train.py:356      loss.backward() (trunk)  mapnet_backward
 *   common/criterion.py:42-52,76-109,137-184          mapnet_loss_fwd_bwd
 *   common/pose_utils.py:234-260 calc_vos*            mapnet_loss_fwd_bwd (folded in)
 *   common/optimizer.py:21-23 + train.py:357-359      mapnet_adam_step (+ mapnet_sqnorm)
 *   models/posenet.py:37-63  parameter layout         mapnet_param_count / mapnet_param_info
 *
 * Conventions: every pointer is a DEVICE pointer unless named host_*; all
 * tensors are contiguous fp32 in the reference's layouts (images NCHW
 * [B,3,H,W]; poses [..,6] = xyz + log-quaternion).  Functions enqueue work on
 * `stream` (a cudaStream_t passed as void*) and return 0 on success; non-zero
 * means failure and mapnet_last_error() (thread-local) describes it.  No
 * exception crosses this boundary; the library never frees caller memory.
 * Handles are not thread-safe (one Python thread drives one GPU, like nn.Module).
 * There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef MAPNET_B200_H_
#define MAPNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mapnet_trunk mapnet_trunk_t;

/* precision of the conv engines / stored activations */
#define MAPNET_PREC_FP32 0      /* fp32 CUDA-core implicit GEMM: strict 1e-4 parity mode            */
#define MAPNET_PREC_BF16 1      /* tcgen05 tensor-core implicit GEMM, bf16 operands, fp32 accumulate */
#define MAPNET_PREC_BF16_SIMT 2 /* bf16 storage + CUDA-core fp32 math: cross-check of mode 1         */
#define MAPNET_PREC_TC_SPLIT 3  /* tcgen05 strict mode: fp32 conv outputs, conv operands as split fp16 hi/lo planes,\
                                   4 MMAs per product, fp32 accumulate -- meets the 1e-4 parity bar on tensor cores */

/* criterion modes (common/criterion.py) */
#define MAPNET_LOSS_POSENET 0    /* PoseNetCriterion      pred [N,6],    targ [N,6]        */
#define MAPNET_LOSS_MAPNET 1     /* MapNetCriterion       pred [N,T,6],  targ [N,T,6]      */
#define MAPNET_LOSS_ONLINE 2     /* MapNetOnlineCriterion pred [N,2T,6], targ [N,2T-1,6]   */
#define MAPNET_LOSS_ONLINE_GPS 3 /* ... gps_mode=True     pred [N,2T,6], targ [N,2T,6]     */

const char* mapnet_last_error(void);
int mapnet_abi_version(void);

/* ---- trunk handle: owns only scratch (activation arena, packed weights, BN partials).
 * max_B == 0 creates a spec-only handle (parameter table queries, no device needed). */
int mapnet_trunk_create(mapnet_trunk_t** out, int max_B, int H, int W, int feat_dim, int precision);
int mapnet_trunk_destroy(mapnet_trunk_t* h);

/* ---- parameter table (state_dict order of the reference PoseNet, 222 entries) */
/* kind: 0 = trainable fp32 in params_flat, 1 = fp32 buffer (BN running stats) in
 * bufs_flat, 2 = int64 num_batches_tracked (offset = index in a separate int64 vector) */
int mapnet_param_count(mapnet_trunk_t* h);
int mapnet_param_info(mapnet_trunk_t* h, int i, char* host_name, int name_cap, int* host_kind, int* host_ndim,
                      int64_t* host_shape4, int64_t* host_offset);
/* Element order of entry i inside its flat buffer.  0: torch's contiguous order of the entry's shape.  1: a conv weight
 * of shape [Co,Ci,KH,KW] stored as [Co][KH][KW][Ci] -- the K-major order the tcgen05 engines read (fprop operand) and
 * accumulate (wgrad), so neither direction needs a transposing pass per step.  To PyTorch that memory is the same
 * [Co,Ci,KH,KW] tensor in torch.channels_last strides: flat[off:off+n].view(Co,KH,KW,Ci).permute(0,3,1,2) -- logical
 * indexing, state_dict()/load_state_dict() and the reference's optimizer code are unaffected.  Every conv except the stem
 * (7x7, 3 input channels) has layout 1; -1 on a bad index.  (ABI version 2.) */
int mapnet_param_layout(mapnet_trunk_t* h, int i);
int64_t mapnet_params_numel(mapnet_trunk_t* h); /* floats in params_flat / grads_flat (padded) */
int64_t mapnet_bufs_numel(mapnet_trunk_t* h);   /* floats in bufs_flat                          */

/* ---- forward: x [B,3,H,W] -> pred [B,6].  training!=0: BN batch statistics + running-
 * stat update in bufs_flat; activations are kept for mapnet_backward.  droprate>0
 * applies dropout with a counter-based mask from (seed, step); step == UINT64_MAX uses and
 * advances a device-side counter owned by the handle (CUDA-graph capturable). */
int mapnet_forward(mapnet_trunk_t* h, const float* x, const float* params_flat, float* bufs_flat, int B,
                   int training, float droprate, uint64_t seed, uint64_t step, float* pred, void* stream);

/* ---- backward of the last training forward: dpred [B,6] -> every parameter gradient,
 * written (not accumulated) into grads_flat at the offsets of params_flat.
 * filter_nans: the NaN filter of the models/posenet.py:28-34 hook on fc_wpqr (NaN entries of that Linear's bias /
 * input / weight gradients become 0: a NaN in dpred[b,3+j] wipes row j of d W_wpqr and sample b's rotation
 * half of the gradient into the trunk). */
int mapnet_backward(mapnet_trunk_t* h, const float* dpred, const float* params_flat, float* grads_flat,
                    int filter_nans, void* stream);

/* The same backward pass in three parts, called in order 0, 1, 2 with the same arguments: after part k returns
 * (asynchronously, on `stream`) the gradients in grads_flat[lo_k, hi_k) (mapnet_grad_part_range) are final, so a
 * data-parallel caller can enqueue their allreduce on another stream while the next part computes.
 * Part 0 = head + layer4 (64 % of the parameters after ~20 % of the backward FLOPs), 1 = layer3, 2 = layer2, layer1, stem.
 * (The reference is single-GPU, common/train.py:193-196; this is the overlap hook of SURVEY.md section 8e.) */
int mapnet_backward_part(mapnet_trunk_t* h, int part, const float* dpred, const float* params_flat, float* grads_flat,
                         int filter_nans, void* stream);
int mapnet_grad_part_range(mapnet_trunk_t* h, int part, int64_t* host_lo, int64_t* host_hi);

/* ---- fused criterion forward+backward.  s4 = device (sax,saq,srx,srq).
 * Outputs: loss[1], dpred (same shape as pred), ds4[4] = d loss / d s4. */
int mapnet_loss_fwd_bwd(int mode, const float* pred, const float* targ, int N, int T_pred, int T_targ,
                        const float* s4, float* loss, float* dpred, float* ds4, void* stream);

/* ---- optimizer step on flat buffers (torch.optim.Adam semantics, L2 weight decay).
 * grad_scale multiplies the gradient first (1/world_size after a sum-allreduce).
 * sqnorm: device scalar holding the squared global grad norm (mapnet_sqnorm) to
 * apply clip_grad_norm_(max_norm), or NULL for no clipping. */
int mapnet_sqnorm(const float* g, int64_t n, float* scratch1024, float* out_sq, void* stream);
int mapnet_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                     const float* sqnorm, float max_norm, void* stream);

/* CUDA-graph capturable variant: the step count lives in device memory (*step_counter is
 * incremented, then used for the bias corrections), so a captured launch stays correct
 * on every replay. */
int mapnet_adam_step_dev(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, int32_t* step_counter,
                         float grad_scale, const float* sqnorm, float max_norm, void* stream);

/* ---- GPU-side input pipeline (geomapnet_b200/csrc/preprocess.cu).
 * Replaces the per-image CPU transform stack of /root/reference/scripts/train.py:119-128 (and eval.py:97-101):
 * transforms.Resize(size) [Pillow bilinear on 8-bit images, bit-exact] [-> ColorJitter, train.py:121-126, bit-exact]
 * -> ToTensor -> Normalize(mean, std), and the frame gather of the MF / MFOnline tuple datasets
 * (dataset_loaders/composite.py:60-97,117-126).
 * One plan per input size; frames are uint8 [n][Hin][Win][3] in device memory, the result float32 [n][3][Hout][Wout]
 * (what PoseNet.forward takes).  mean3 / std3 are HOST pointers to 3 floats (stats.txt row 0, sqrt of row 1). */
typedef struct mapnet_preprocess mapnet_preprocess_t;
int mapnet_preprocess_create(mapnet_preprocess_t** out, int Hin, int Win, int size, int max_images);
int mapnet_preprocess_output_size(const mapnet_preprocess_t* h, int* Hout, int* Wout);
int mapnet_preprocess_run(mapnet_preprocess_t* h, const void* img_nhwc_u8, int n, const float* mean3, const float* std3,
                          float* out_nchw, void* out_u8_or_null /* the resized uint8 image, tests */, void* stream);
/* frame_index_dev: NULL, or n int32 in DEVICE memory -- output image i is cut from frame frame_index[i] of the sequence
 * at img_nhwc_u8 (the tuple gather).  jitter_dev: NULL, or n records {int32 order[4]; float factor[4]} in DEVICE memory:
 * the adjustments (0 brightness, 1 contrast, 2 saturation, 3 hue -- torchvision's fn_idx) in the order they are applied
 * and their factors indexed by adjustment id, i.e. one ColorJitter.get_params() draw per image. */
int mapnet_preprocess_run_ex(mapnet_preprocess_t* h, const void* img_nhwc_u8, int n, const int32_t* frame_index_dev,
                             const void* jitter_dev, const float* mean3, const float* std3, float* out_nchw,
                             void* out_u8_or_null, void* stream);
int mapnet_preprocess_destroy(mapnet_preprocess_t* h);

/* ---- batched inference post-processing + pose-graph optimisation (csrc/pgo.cu).
 * mapnet_pose_post: scripts/eval.py:163-181 for n poses at once -- pred6 = (t, log q) float32 -> out7 = (t * pose_s +
 * pose_m, qexp(log q)) float64; pose_m3 / pose_s3 are HOST pointers (pose_stats.txt) or NULL (keep the translations).
 * mapnet_pgo_optimize: common/pose_utils.py:458-804 (PoseGraph / PoseGraphFC.optimize behind optimize_poses) for
 * n_windows windows of N <= 16 poses in one launch: poses [n_windows][N][7], vos [n_windows][E][7] (E = N-1, or N(N-1)/2
 * pairs i<j row-major when fc_vos), quaternions (w,x,y,z), fp64, device pointers; sax..srq are the covariances
 * eval.py passes; n_iters = 10 in the reference.  flags bit 0: 0 = the reference's linear solve taken literally
 * (pose_utils.py:605-608 calls solve_triangular(R.T, -b) with scipy's default lower=False, which reads only the diagonal
 * of R': x = R^-1 diag(R)^-1 (-b)), 1 = the Gauss-Newton step H^-1 (-b).  *status_dev (device int, may be NULL) is set
 * to 1 when a window's normal matrix is not positive definite (scipy's cholesky raises there). */
int mapnet_pose_post(const float* pred6, double* out7, int64_t n, const double* pose_m3, const double* pose_s3, void* stream);
int mapnet_pgo_optimize(const double* poses, const double* vos, double* out, int n_windows, int N, int fc_vos,
                        double sax, double saq, double srx, double srq, int n_iters, int flags, int* status_dev, void* stream);

/* ---- measurement support (bench.py): number of kernels this library has launched so
 * far in this process, and per-class conv timing with CUDA events on the launching
 * stream (class 0 fprop, 1 dgrad, 2 wgrad; algorithmic FLOPs of each launch summed). */
unsigned long long mapnet_launch_count(void);
int mapnet_profile(mapnet_trunk_t* h, int enable);
int mapnet_profile_read(mapnet_trunk_t* h, double* host_ms3, double* host_flops3, int* host_launches3);

/* ---- unit entry points used by the parity tests (tests/test_gpu_kernels.py) */
int mapnet_test_conv(int precision, int kind /*0 fprop 1 dgrad 2 wgrad*/, int B, int Hi, int Wi, int Ci, int Co,
                     int k, int stride, const void* in0, const void* in1, const void* wmat, void* out, void* stream);

/* host-only: the tile / parity-class / filter-tap plan of the tensor-core engines for one conv, as JSON
 * (tests/test_tc_plan.py replays it on the CPU against torch's convolutions; kind 0 fprop, 1 dgrad, 2 wgrad) */
int mapnet_test_plan_describe(int kind, int B, int Hi, int Wi, int Ci, int Co, int k, int stride, int with_shortcut,
                              char* buf, int cap);

/* the tensor-core dgrad of a downsampling BasicBlock's two input-side convs in ONE launch, as the trunk runs it:
 * dx = dgrad(conv1 3x3/s2/p1; dy1, w1_dg [Ci][3][3][Co]) + dgrad(downsample 1x1/s2; dy2, w2_dg [Ci][Co]), all bf16 NHWC
 * (torchvision BasicBlock.forward's two uses of the block input, /root/reference/models/posenet.py:66). */
int mapnet_test_dgrad_shortcut(int B, int Hi, int Wi, int Ci, int Co, const void* dy1, const void* dy2,
                               const void* w1_dg, const void* w2_dg, void* dx, void* stream);

/* one tcgen05 conv launch with the epilogue the training step fuses into it (bf16 mode; tensors bf16 NHWC):
 * kind 0 fprop -> out + the BatchNorm batch statistics (sum, sum of squares per output channel) of the stored values;
 * kind 1 dgrad -> out = the gradient gated by the consumer BatchNorm's ReLU (+ residual) and that BatchNorm's backward
 * reductions (sum g, sum g*y [, sum g*yd]).  host_sums: [3][C] doubles.  Replaces cuDNN BN forward-training statistics
 * and the reduction half of cuDNN BN backward behind torchvision BasicBlock (/root/reference/models/posenet.py:66). */
int mapnet_test_conv_epilogue(int kind, int B, int Hi, int Wi, int Ci, int Co, int k, int stride, const void* in0,
                              const void* wmat, const void* residual, const void* y, const void* zmask, const void* yd,
                              const float* mscale, const float* mshift, void* out, double* host_sums, void* stream);

/* the tensor-core stem (7x7/s2/p3, 3 -> 64) alone, as the trunk runs it: space-to-depth image, packed weights,
 * fprop into y_out (bf16 NHWC [B,Hc,Wc,64]) and, when dy / dw_oihw are given, the weight gradient in the
 * .grad layout [64,3,7,7].  Replaces torchvision ResNet.conv1 (/root/reference/models/posenet.py:66). */
int mapnet_test_stem(int B, int H, int W, const float* x_nchw, const float* w_oihw, void* y_out, const void* dy,
                     float* dw_oihw, void* stream);

/* micro-benchmark of one tensor-core conv configuration (tools/bench_conv.py): average ms per launch */
int mapnet_bench_conv(int kind, int B, int Hi, int Wi, int Ci, int Co, int k, int stride, const void* in0,
                      const void* in1, const void* wmat, void* out, int iters, float* host_ms);

#ifdef __cplusplus
}
#endif
#endif /* MAPNET_B200_H_ */
