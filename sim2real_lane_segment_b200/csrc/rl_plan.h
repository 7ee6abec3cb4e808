// rl_plan.h — host-side description of one FC-DenseNet67 instance at a fixed (B, H, W, num_cls):
// arena geometry, the layer schedule and every offset the kernels need.  Internal to librightlane.so.
//
// Memory model (all in one caller-owned workspace, see rl_plan_workspace_bytes):
//   * one NHWC bf16 *concat arena* per resolution level; producers write channel slices, consumers read
//     channel prefixes, so torch.cat (reference layers.py:33,35,39,68) never materialises.
//     Physical channel order of level l (l = 0..4):  [ skip C_skip(l) | up 80 | new 80 ];  level 5
//     (bottleneck) is [ in 448 | new 80 ].  The reference's up-path order is [up | skip | new]
//     (layers.py:68 `cat([out, skip])`), so parameters indexed by input channel are permuted when they
//     are packed / when their gradients are scattered (ref_channel()).
//   * a second arena of the same geometry for activation gradients (training only)
//   * per-channel fp64 sums (Σx, Σx²) per level: BatchNorm batch statistics are computed once by the
//     kernel that produces a channel and shared by every BatchNorm that consumes it
//   * per-channel fp64 backward accumulators (A_c, B_c): the mean-subtraction terms of BatchNorm
//     backward are affine in x̂ with per-channel coefficients, so they are summed over all consumer
//     layers and applied once, when the producing layer reads its output gradient (dz_prepare kernels).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace rl {

enum LayerType { L_FIRST = 0, L_DENSE = 1, L_TD = 2, L_TU = 3 };

constexpr int kGrowth = 16;
constexpr int kFirst = 48;
constexpr int kBlockLayers = 5;
constexpr int kLevels = 6;  // 5 resolution levels + bottleneck
constexpr int kUpCh = 80;   // growth * layers carried upward by every TransitionUp
constexpr int kTaps = 9;
constexpr int kDzBufs = 5;      // padded dz16 plane pairs per level: one per DenseLayer of a block
constexpr int kBulkCb = 96;     // channel block of the fused block input-gradient kernel (two halves of <= 48)
constexpr int kBulkHalf = 48;

struct Level {
  int H, W, Ct;
  int Cskip;  // channels of the down-path block output at this level (level 5: input channels 448)
  size_t act_off, dact_off;      // bytes into workspace (act: NHWC [px][Ct]; dact: planar [Ct/16][px][16], rl_dev.cuh)
  size_t xp_off;                 // training: planar mirror of act (same layout as dact) for the backward's lane = pixel reads
  size_t stats_off, acc_off;     // bytes into workspace: double[2][Ct] each
  long px;                       // B*H*W
  // zero-padded planes of the 16-channel output gradient of a DenseLayer (training):
  // two planes [npos][8] bf16; position(b,y,x) = guard + (b*(H+1)+y)*(W+2) + x+1; pads are never written
  size_t dzp_off;                // kDzBufs buffers of 2 planes each: DenseLayer j of a block uses buffer j, so the block's
  long dzp_npos, dzp_guard;      // fused input-gradient kernel and its weight gradients can read all five after the chain
};

struct Layer {
  int type;
  int lvl_in, lvl_out;
  int cin, cout;
  int in_off, out_off;  // channel offsets inside the arenas
  int cs_perm;          // >0: up-block layer, C_skip of its level (ref<->phys permutation); 0: identity
  long p_gamma, p_beta, p_w, p_b;  // float offsets into the parameter / gradient arenas (-1: absent)
  long b_rm, b_rv;                 // float offsets into the BN-buffer arena
  size_t wpk_fwd, wpk_bwd;         // byte offsets into workspace (packed bf16 weights)
  size_t wpk_chain, wpk_bulk;      // dense layers: [18][128][8] rows = the block's earlier new channels; [ncb][18][96][8] block inputs
  int blk, jblk;                   // dense layers: index into Plan::blocks and position 0..4 inside the block (else -1)
  size_t mask_off;                 // byte offset into workspace: float[B][cout] dropout scale (0 or 1/(1-p))
  size_t amax_off;                 // TD only: byte offset of uint8 argmax [pooled px][C]
  int rmw_from;                    // backward: input-gradient channels below this are overwritten, not accumulated
  int bn_index;                    // 0..59 or -1
};

// A DenseBlock (reference layers.py:18-40): five growth-16 layers on one level.  Channels [0, C0) are the block's
// input, [C0, C0 + 80) the layers' outputs.  Backward runs the five layers as a chain over the new channels only and
// then ONE kernel adds the input gradient of all five layers to the block input (dgrad_bulk_kernel).
struct Block {
  int first;     // index of layer 0 of the block in Plan::layers
  int lvl;
  int C0;
  int rmw;       // 1: the block-input gradient accumulates into earlier contributions, 0: it is the first writer
  int ncb, cbs;  // channel blocks of the fused kernel: ncb blocks of cbs channels (the last one may be shorter)
};

struct PackEntry {
  int kind;      // 0 dense fwd, 1 dense bwd, 2 TD fwd, 3 TD bwd, 4 (retired), 5 TU bwd, 6 dense chain bwd, 7 dense bulk bwd, 8 TU fwd
  int cin, cs_perm, np0;
  int c_lo, nch;  // kind 6: channel range [c_lo, c_lo + nch); kind 7: c_lo = C0, nch = cbs
  long dst_off;  // bytes into workspace
  long n16;      // number of 16-byte chunks to produce
  long p_w;      // float offset of the weight tensor in the parameter arena
};

struct Plan {
  int B, H, W, num_cls;
  int sm_count;
  Level lv[kLevels];
  std::vector<Layer> layers;
  std::vector<Block> blocks;
  long n_params;      // floats in the parameter arena (3 461 220 for 4 classes)
  long n_bnbuf;       // floats in the BN-buffer arena
  long p_cls_w, p_cls_b;
  size_t wpk_off, wpk_bytes;
  size_t mask_off_all, mask_bytes;
  size_t stats_off_all, stats_bytes;  // Σx, Σx² for all levels (zeroed every training forward)
  size_t acc_off_all, acc_bytes;      // A_c, B_c for all levels (zeroed every backward)
  size_t tmp_off, tmp_bytes;          // shared scratch: dz buffers, convT column matrices
  size_t dl_off, dl_bytes;            // head: d(logits)/n as bf16 [px0][16]
  size_t dzp_off_all, dzp_bytes;      // all levels' padded dz16 planes (zeroed once at bind)
  size_t scal_off;                    // small scalars: hist[8] int, loss sums, counters (256 B)
  size_t bntab_off;                   // device copy of the per-BN table for rl_bn_running_update
  size_t packtab_off;                 // device copy of the weight-packing table
  int n_pack, n_bn;
  size_t xcol_off;                    // first conv: bf16 im2col rows [px0][32] kept for its weight gradient
  size_t ws_bytes;
  void* ws;                           // bound workspace (device)
  alignas(64) unsigned char tmap_act[6][128];  // CUtensorMap per level arena: [pixels][Ct] bf16, box 128 px x 64 ch, 128-byte swizzle
  alignas(64) unsigned char tmap_dz[kLevels][kDzBufs][128];  // CUtensorMap per padded dz16 buffer (dgrad_bulk_kernel halo tiles)
  alignas(64) unsigned char tmap_dz_im2col[kLevels][kDzBufs][128];  // im2col map of the same buffer (wgrad_dense2_kernel)
  bool tmap_ok;
  size_t gstage_off;                  // training: fp32 [n_params] raw gradients of one backward pass (scaled into `grads` at its end)
  void* side_stream;                  // cudaStream_t: weight gradients run here, concurrently with the input-gradient chain
  void* ev_dz[4];                     // cudaEvent_t: dz buffer q written (main stream); 0,1: per-layer path, 2,3: scratch
  void* ev_wg[4];                     // cudaEvent_t: weight gradient that read buffer q finished (side stream)
  bool ev_wg_used[4];
  void* ev_blk[kLevels];              // cudaEvent_t: the batched weight gradients of the last DenseBlock of a level finished
  bool ev_blk_used[kLevels];
  // per-plan switches (rl_plan_set_option): side-stream overlap, planar activation mirror, tensor-map forward,
  // block-fused backward
  int opt_overlap, opt_xp, opt_fwd_tma, opt_bwd_fused;
  bool training;                      // arenas for gradients present
  char err[256];
};

// physical channel -> reference channel index for parameters indexed by input channel
__host__ __device__ inline int ref_channel(int c, int cs_perm) {
  if (cs_perm <= 0) return c;
  if (c < cs_perm) return c + kUpCh;
  if (c < cs_perm + kUpCh) return c - cs_perm;
  return c;
}

int build_plan(Plan* p, int B, int H, int W, int num_cls, int training);

// scalar block layout (ints / floats / doubles at scal_off)
struct Scalars {
  int hist[8];          // label histogram
  int cm[64];           // confusion matrix counts (label * C + pred), up to 8x8
  double loss_num;      // Σ w[y]·nll
  double loss_den;      // Σ w[y]
  int bad_label;        // count of labels outside [0, num_cls)
  int pad[3];
};

}  // namespace rl
