/* rightlane.h — C ABI of librightlane.so: the B200-native (sm_100a) replacement for the arithmetic behind
 * timurlenk07/sim2real_lane_segment's segmentation step.  The reference has no native code and therefore no FFI;
 * each entry point cites the reference Python it stands in for (paths relative to rightLaneNetwork/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise; the library allocates nothing
 *     on the device: the caller provides one workspace of rl_plan_workspace_bytes() bytes (256-byte aligned)
 *   - every call is asynchronous on the given cudaStream_t (passed as void*), performs no host synchronisation and no
 *     host callback, so a whole step can be captured in a CUDA graph
 *   - return value: 0 = ok, <0 = error (RL_ERR_*), message via rl_last_error() (thread-local); nothing throws or aborts
 *   - parameters / gradients / optimizer state are flat fp32 arenas in the reference's nn.Module registration order
 *     (state_dict order: featureExtractor.* then classifier.*); BN running statistics are a second flat fp32 arena
 *     ([running_mean, running_var] per BatchNorm, registration order)
 *   - activations are internal (NHWC bf16 concat arenas inside the workspace); images enter as fp32 NCHW exactly as the
 *     reference's DataLoader delivers them, probabilities leave as fp32 NCHW, masks as uint8 [B,H,W]
 */
#ifndef RIGHTLANE_H_
#define RIGHTLANE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RL_OK = 0, RL_ERR_ARG = -1, RL_ERR_SHAPE = -2, RL_ERR_CUDA = -3, RL_ERR_STATE = -4 };

typedef struct rl_plan rl_plan;

int rl_abi_version(void);
const char* rl_last_error(void);

/* FC-DenseNet67 + classifier for a fixed batch geometry.  Stands in for the module construction at
 * trainingModules/TrainingBase.py:32-33 (FCDenseNet67Base + FCDenseNet67Classifier, models/FCDenseNet/tiramisu.py:178-194).
 * training != 0 additionally reserves gradient arenas.  H, W >= 32; 2 <= num_cls <= 4. */
int rl_plan_create(rl_plan** out, int B, int H, int W, int num_cls, int training);
void rl_plan_destroy(rl_plan* p);
long rl_plan_workspace_bytes(const rl_plan* p);
long rl_plan_num_params(const rl_plan* p);     /* 3 461 220 for num_cls = 4 */
long rl_plan_num_bn_buffers(const rl_plan* p); /* floats in the BN running-statistics arena */
int rl_plan_bind(rl_plan* p, void* workspace, void* stream);

/* introspection used by tests and by the Python mirror (no device work) */
enum {
  RL_Q_NUM_LAYERS = 0,
  RL_Q_LEVEL_ACT_OFF = 1,   /* index = level 0..5 -> byte offset of the NHWC bf16 arena */
  RL_Q_LEVEL_DACT_OFF = 2,
  RL_Q_LEVEL_CT = 3,
  RL_Q_LEVEL_H = 4,
  RL_Q_LEVEL_W = 5,
  RL_Q_LEVEL_CSKIP = 6,
  RL_Q_SCALARS_OFF = 7,     /* struct: int hist[8]; int cm[64]; double loss_num, loss_den; int bad_label */
  RL_Q_LAYER_TYPE = 8,      /* index = layer in forward order: 0 first conv, 1 dense, 2 transition down, 3 transition up */
  RL_Q_LAYER_MASK_OFF = 9,  /* byte offset of float[B][cout] dropout keep-scale (dense / transition-down layers) */
  RL_Q_LAYER_COUT = 10,
  RL_Q_LAYER_PARAM_W = 11,  /* float offset of the layer's conv weight in the parameter arena */
  RL_Q_LEVEL_STATS_OFF = 12, /* byte offset of double[2][Ct]: per-channel sum, sum of squares */
  RL_Q_GSTAGE_OFF = 13,      /* byte offset of float[num_params]: raw gradients of the current backward pass */
  RL_Q_LEVEL_XP_OFF = 14,    /* byte offset of the planar mirror [Ct/16][pixels][16] bf16 of a level's activations */
  RL_Q_LEVEL_ACC_OFF = 15,   /* byte offset of double[2][Ct]: BatchNorm-backward mean terms A_c, B_c (sums, not yet / N) */
  RL_Q_LAYER_LEVEL_IN = 16,
  RL_Q_LAYER_LEVEL_OUT = 17,
  RL_Q_LAYER_CIN = 18,
  RL_Q_LAYER_IN_OFF = 19,    /* first input channel in the input level's arena (TransitionUp) */
  RL_Q_LAYER_OUT_OFF = 20,   /* first output channel in the output level's arena */
  RL_Q_LAYER_CS_PERM = 21,   /* > 0: up-path layer, C_skip of its level (reference [up|skip|new] <-> physical [skip|up|new]) */
  RL_Q_LAYER_AMAX_OFF = 22,  /* TransitionDown: byte offset of the uint8 arg-max [pooled pixels][C] */
  RL_Q_LAYER_RMW_FROM = 23,
  RL_Q_LAYER_PARAM_GAMMA = 24, /* float offsets in the parameter arena (-1: absent): BatchNorm weight (bias follows, then conv weight) */
  RL_Q_LAYER_PARAM_B = 25,     /* conv bias */
  RL_Q_LAYER_BN_RM = 26,       /* float offset of running_mean in the BN-buffer arena (running_var follows) */
  RL_Q_DL_OFF = 27             /* byte offset of bf16 [pixels][16]: d(logits)/||f|| written by rl_head(backward) */
};
int rl_plan_query(const rl_plan* p, int what, int index, long* out);

/* bf16 repack of all convolution weights into tensor-core operand layout; call after every parameter change
 * (load_state_dict, optimizer step). */
int rl_pack_weights(rl_plan* p, const float* params, void* stream);

/* nn.Dropout2d(0.2) keep-scales (models/FCDenseNet/layers.py:12,51) for all layers: p_drop = 0 writes 1.0 everywhere
 * (eval / parity runs); tests may overwrite the per-layer float[B][cout] blocks (RL_Q_LAYER_MASK_OFF). */
int rl_set_dropout(rl_plan* p, float p_drop, unsigned long long seed, unsigned long long step, void* stream);
/* the same with the step counter read from device memory (step + *step_dev): the launch holds no host state that changes
 * between steps, so it can be part of a captured CUDA graph of the whole training step */
int rl_set_dropout_dev(rl_plan* p, float p_drop, unsigned long long seed, unsigned long long step, const long long* step_dev,
                       void* stream);

/* FCDenseNetFeatureExtractor.forward without the final F.normalize (tiramisu.py:89-103): x fp32 [B,3,H,W] ->
 * level-0 arena (288 channels).  train != 0: BatchNorm uses batch statistics (and they are retained for
 * rl_bn_running_update / rl_backward); else running statistics from bn_buffers. */
int rl_forward(rl_plan* p, const float* params, const float* bn_buffers, const float* x, int train, void* stream);
/* the same for layers layer_lo..layer_hi of the plan's forward order only (0 = first conv; RL_Q_LAYER_TYPE), on whatever
 * the arenas hold: the per-layer entry point the parity tests use (models/FCDenseNet/layers.py:5-69, one module at a
 * time).  Does not clear the batch statistics. */
int rl_forward_range(rl_plan* p, const float* params, const float* bn_buffers, const float* x, int train, int layer_lo,
                     int layer_hi, void* stream);

/* F.normalize (tiramisu.py:105) + FCDenseNetClassifier.forward (tiramisu.py:120-125) + cross_entropy on the
 * probabilities + argmax + confusion counts (SimpleTrain.py:15-20, TrainingBase.py:84-93, test.py:93-100).
 *   y: labels uint8 or int64 [B,H,W] or NULL; weighted: class weights 1/count as getClassWeight (TrainingBase.py:12-23)
 *   probs (fp32 [B,C,H,W]) / mask (uint8 [B,H,W]) may be NULL.  Loss sums and counts land in the scalars block.
 *   backward: 0 = forward only; 1 = the same launch also starts the backward pass with the fused CE gradient
 *   d(loss)/d(features) for loss grad_output 1 (rl_backward applies the real grad_output); 2 = backward from an external
 *   d(loss)/d(probs) (dprobs fp32 [B,C,H,W]).  feat_grad_sign = -1 realises GradReverse (tiramisu.py:7-18).
 *   backward != 0 clears the plan's gradient stage (RL_Q_GSTAGE_OFF) and adds the classifier-bias gradient to it. */
int rl_head(rl_plan* p, const float* params, const void* y, int y_is_i64, int weighted, float* probs, uint8_t* mask,
            int backward, const float* dprobs, float feat_grad_sign, void* stream);

/* loss.backward() for everything below the head (SimpleTrain.py:25 + autograd).  The pass's raw gradients are staged in
 * the workspace and added to grads (fp32 arena, same layout as params) at the end:
 *   grads += grad_scale * (grad_scale_dev ? *grad_scale_dev : 1) * stage
 * (grad_scale_dev: the loss's grad_output as a device scalar, read without a host sync).
 * Requires rl_forward(train=1) and rl_head(backward != 0) on the same plan. */
int rl_backward(rl_plan* p, const float* params, float* grads, float grad_scale, const float* grad_scale_dev, void* stream);
/* The three pieces rl_backward is made of, for callers that interleave the pass with other work (data-parallel training
 * fires the all-reduce of a parameter range as soon as the layers that own it are done, ddp.py):
 *   rl_backward_begin: clears the BatchNorm-backward accumulators;  rl_backward_range: see below;
 *   rl_grad_accum: grads[lo:hi) += grad_scale * (grad_scale_dev ? *grad_scale_dev : 1) * stage[lo:hi)  (float offsets). */
int rl_backward_begin(rl_plan* p, void* stream);
int rl_grad_accum(rl_plan* p, float* grads, long lo, long hi, float grad_scale, const float* grad_scale_dev, void* stream);
/* backward of layers layer_hi..layer_lo only (descending; layer_hi = RL_Q_NUM_LAYERS adds the classifier weight
 * gradient), from whatever the gradient arenas hold, into the gradient stage; no clearing, no final scaling: the
 * per-layer / per-DenseBlock entry point the parity tests use (autograd of layers.py:5-69 one module at a time). */
int rl_backward_range(rl_plan* p, const float* params, int layer_hi, int layer_lo, void* stream);

/* running_mean / running_var momentum update of all 60 BatchNorm2d layers from the last training forward */
int rl_bn_running_update(rl_plan* p, float* bn_buffers, float momentum, void* stream);

/* torch.optim.AdamW step (SimpleTrain.py:28) over a flat arena; step >= 1; grad_scale multiplies g first.  The
 * hyper-parameters are doubles: 1 - beta, the bias corrections and lr / bias_correction1 are formed in double, as torch
 * forms them in Python floats, before the fp32 update. */
int rl_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, long step, float grad_scale, void* stream);
/* the same update with the hyper-parameters in device memory, hyper_dev = double[6] {lr, beta1, beta2, eps, weight_decay, step}:
 * graph-capturable (the learning-rate schedule and the step counter change between replays without re-capturing) */
int rl_adamw_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, const double* hyper_dev,
                      float grad_scale, void* stream);
/* torch.optim.SGD(momentum, nesterov=True, weight_decay) step (MMETrainingModule.py:16-19) */
int rl_sgd_nesterov_step(float* params, const float* grads, float* momentum_buf, long n, float lr, float momentum,
                         float weight_decay, int first_step, float grad_scale, void* stream);

/* Input pipeline, deterministic part: MyTransform(width, height, gray, augment=False) of
 * dataManagement/myTransforms.py:8-31 applied to a batch of cv2.imread frames (myDatasets.py:49-53) in one kernel.
 * img_bgr: uint8 [B][Hs][Ws][3] (device), label: uint8 [B][Hs][Ws] or null.  x_out: float32 [B][3][H][W] =
 * Normalize(cv2.resize(img, (W,H), INTER_LINEAR)) -- the resize is cv2's uint8 fixed-point algorithm, bit-exact --
 * with ToGray applied in between when gray != 0; y_out: int64 [B][H][W] = cv2.resize(label, INTER_NEAREST), or null.
 * The random augmentations are rl_preprocess_augment below. */
int rl_preprocess(const unsigned char* img_bgr, const unsigned char* label, int B, int Hs, int Ws, int H, int W, int gray,
                  float* x_out, long long* y_out, void* stream);

/* Input pipeline with the random augmentations: MyTransform(width, height, gray, augment=True) of
 * dataManagement/myTransforms.py:8-20 for caller-drawn parameters, one entry per frame:
 *   HueSaturationValue -> lut_h/lut_s/lut_v: the three cv2.LUT tables albumentations builds from its hue / sat / val shifts
 *                         (hsv = 0: all shifts were 0, albumentations returns the frame untouched);
 *   RandomSizedCrop    -> the crop window (crop_x, crop_y, crop_w, crop_h) in source pixels, resized to (W, H);
 *   MotionBlur         -> ksize (3, 5 or 7; 0: not chosen) and the normalised ksize x ksize float32 kernel, row major;
 *   GaussNoise         -> noise != 0: noise_dev[b] (float64 [H][W][3], what RandomState.normal drew) is added.
 * Every uint8 stage is bit-exact with what albumentations 0.5.2 makes cv2 compute (RGB2HSV / LUT / HSV2RGB, crop +
 * resize INTER_LINEAR / INTER_NEAREST, filter2D BORDER_REFLECT_101, float32 + float64 noise clipped and truncated). */
typedef struct rl_augment_params {
  unsigned char lut_h[256], lut_s[256], lut_v[256];
  int hsv;
  int crop_x, crop_y, crop_w, crop_h;
  int ksize;
  float kernel[49];
  int noise;
} rl_augment_params;
/* img_bgr uint8 [B][Hs][Ws][3], label uint8 [B][Hs][Ws] or null, params_dev [B], noise_dev float64 [B][H][W][3] or null,
 * scratch_u8 uint8 [B][H][W][3] (receives the augmented uint8 frames), x_out float32 [B][3][H][W], y_out int64 [B][H][W]
 * or null: all device pointers. */
int rl_preprocess_augment(const unsigned char* img_bgr, const unsigned char* label, int B, int Hs, int Ws, int H, int W, int gray,
                          const rl_augment_params* params_dev, const double* noise_dev, unsigned char* scratch_u8, float* x_out,
                          long long* y_out, void* stream);

/* Inference consumer (makeDemoVideo.py:36-42, comparison.py:44-58): out uint8 [B][H][W][3] = cv2.resize(frame, (W,H))
 * (INTER_LINEAR, bit-exact) with the predicted classes painted over it in BGR: 1 -> (0,255,0), 2 -> (255,0,0),
 * 3 -> (0,0,255).  mask: uint8 (rl_head's mask output) or int64 [B][H][W]. */
int rl_overlay(const unsigned char* frames_bgr, int B, int Hs, int Ws, int H, int W, const void* mask, int mask_is_i64,
               unsigned char* out, void* stream);

/* launch accounting: number of kernels this library has launched (kernel_class < 0: all classes), and optional
 * per-class CUDA-event timing for bench.py's roofline (classes: 0 dense_fwd, 1 gemm_px, 2 dgrad_dense, 3 dgrad_td,
 * 4 wgrad, 5 head, 6 dz_prepare, 7 channel_stats, 8 first_fwd, 9 tu_col2im, 10 pack, 11 optimizer, 12 misc,
 * 13 dgrad_bulk, 14 wgrad_dense).  rl_profile_end synchronises the device and fills ms[RL_NUM_KERNEL_CLASSES],
 * launches[RL_NUM_KERNEL_CLASSES] (host pointers). */
enum { RL_NUM_KERNEL_CLASSES = 15 };
long rl_launch_count(int kernel_class);
int rl_profile_begin(void);
int rl_profile_end(double* ms, long* launches);
int rl_profile_launches(float* ms, int* kernel_class, int max); /* per-launch times of the last profiled region */

/* per-plan switches (all default 1): "overlap": weight-gradient kernels of rl_backward run on a plan-owned side stream
 * concurrently with the input-gradient chain (joined before rl_backward returns control of the stream); "fwd_tma": the
 * DenseLayer forward goes through the tensor-map TMA kernel (0: the register-staged kernel; same GEMMs, the fp32
 * shift-add order differs); "xp": training forward passes also write the planar mirror of the activation arenas that the
 * backward kernels read (0: they read the NHWC arenas; set it before rl_forward); "bwd_fused": DenseBlocks run their
 * backward through the block-fused kernels (0: five per-layer input-gradient and weight-gradient launches). */
int rl_plan_set_option(rl_plan* p, const char* name, int value);

/* test hook: one tcgen05 GEMM (M = 128) from caller-supplied shared-memory images of A and B and the upper bits of both
 * matrix descriptors (start address filled in by the kernel, advanced by *_kstep_bytes per K = 16 step) */
int rl_selftest_umma_raw(const void* a_img, int a_bytes, unsigned long long a_desc_tmpl, int a_kstep_bytes,
                         const void* b_img, int b_bytes, unsigned long long b_desc_tmpl, int b_kstep_bytes, unsigned idesc,
                         int nk, int N, float* D, void* stream);

/* measurement hook: stream the channel window [c0, c0+nch) of an NHWC bf16 arena [npx][Ct] with 32-byte accesses
 * (mode 0 read, 1 read+write back) -- what the memory system delivers for the access pattern of the concat arenas */
/* fragment layout probe of tcgen05.ld.16x256b.x4: out[128 threads][2 halves][16] receives the (lane << 8 | column) tags
 * each thread read from a 128 x 32 TMEM block (tests/test_umma_selftest.py) */
int rl_selftest_tmem_shapes(unsigned int* out, void* stream);
/* the tensor map with overlapping dimensions that delivers the im2col tile of a DenseLayer's dz planes in one TMA box
 * (csrc/rl_bwd_fused.cu): planes = bf16 [2][npos][8] (device), box start coordinate c1 along the position dimension;
 * out receives the 18 x 64 x 16-byte tile as it landed in shared memory */
int rl_selftest_im2col_tmap(const void* planes, long npos, long positions, int W, int c1, void* out, void* stream);
int rl_selftest_stream(void* arena, long npx, int Ct, int c0, int nch, int mode, int grid, void* sink, void* stream);

/* test hook: one 128 x N x K tcgen05 GEMM through the shared-memory operand conventions of csrc/rl_umma.cuh */
int rl_selftest_umma(const void* A, int a_rows, int a_cols, int a_mn_major, int a_row_shift, const void* B, int b_rows,
                     int b_cols, int b_mn_major, int b_row_shift, int N, int K, float* D, int bulk_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIGHTLANE_H_ */
