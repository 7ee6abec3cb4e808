// rl_kernels.h — argument blocks and launchers of the device kernels (internal to librightlane.so).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "rl_dev.cuh"
#include "rl_plan.h"

namespace rl {

// ---------------------------------------------------------------- forward, tensor core
struct DenseFwdArgs {
  const __nv_bfloat16* act;  // level arena (input prefix [0,Cin))
  __nv_bfloat16* act_out;    // same arena (output slice [Cin, Cin+16))
  int Ct, Cin, H, W, rows_total;
  int rows_per_cta, nstrips, strip_w;
  const uint8_t* wpk;        // packed bf16 weights, K chunks of 64: [jc][144][8]
  const float* bias;         // [16]
  const float* mask;         // [B][16] dropout scale or null
  BnSrc bn;
  double* stats_out;         // level stats base (Σx at [c], Σx² at [Ct+c]) or null
  __nv_bfloat16* x16;        // training: the output's 16-channel group plane [px][16] of the planar mirror (or null)
};
int launch_dense_fwd(const DenseFwdArgs& a, int grid, cudaStream_t st);
// TMA-fed variant (levels up to 160 wide: full rows; wider levels: strips of at most 126 columns): tmap = the level arena's 128-byte CUtensorMap blob
int launch_dense_fwd_tma(const DenseFwdArgs& a, const void* tmap, int grid, cudaStream_t st);
// host: encode a [pixels][Ct] bf16 arena as a 2-D tiled tensor map, box 128 px x 64 channels, SWIZZLE_128B
int encode_arena_tmap(void* tmap_out, const void* arena, long npx, int Ct, int W);

// TransitionDown forward: BN -> ReLU -> Conv1x1 (+bias, Dropout2d scale) -> MaxPool2d(2) in one kernel (gemm_px_kernel)
struct GemmPxArgs {
  const __nv_bfloat16* A;
  int lda, a_off, K;
  long ntiles;
  int H, W, Hp, Wp, wblocks;  // input / pooled geometry, 32-window blocks per pooled row
  int transform;
  BnSrc bn;
  const uint8_t* wpk;
  int N;            // total output columns (all parts)
  int part_n0[3], part_np[3];
  const float* bias;
  const float* mask;
  __nv_bfloat16* out;
  int ldo, o_off;
  __nv_bfloat16* out_xp;  // training: planar mirror of the pooled level (written besides out) or null
  long out_xp_npx;
  uint8_t* amax;
};
int launch_gemm_px(const GemmPxArgs& a, int grid_x, int nparts, cudaStream_t st);
// warp-specialised variant with the part's weights resident in shared memory (rl_td.cu); td_fwd_fits: its shared memory suffices
bool td_fwd_fits(int K, const int* part_np, int nparts);
int launch_td_fwd(const GemmPxArgs& a, int grid_x, int nparts, cudaStream_t st);
int configure_td_kernels();

// ---------------------------------------------------------------- backward, tensor core
struct DgradArgs {
  // G^T[c, p] = Σ_k W^T[c, k] · dz[p(-tap shift), k]; epilogue: ReLU mask, BN-backward sums, accumulate into dX
  int dense;                 // 1: 3x3 growth-16 layer (9 taps, halo tile); 0: 1x1 TransitionDown (K = Kdz, streamed)
  const __nv_bfloat16* dz;   // td: [px][Kdz]; dense: zero-padded planes (see Level::dzp_off)
  long dz_npos, dz_guard;    // dense: positions per plane, front guard
  int Kdz;
  const uint8_t* wpk;        // dense: per M-block [18][128][8]; td: per M-block, per K chunk [8][128][8]
  const __nv_bfloat16* act;  // level arena
  __nv_bfloat16* dact;       // gradient arena, planar by 16-channel group (rl_dev.cuh: dact_index)
  int Ct, Cin, H, W, B;      // Cin: number of channels this launch covers, starting at arena channel c_lo
  int c_lo;                  // first arena channel (0 for a whole layer; C0 of the block for the chain launches)
  int rmw_from;
  BnSrc bn;
  float* g_gamma;            // reference-ordered gradient slots; null: this launch leaves dgamma / dbeta to wgrad_dense2
  float* g_beta;
  double* acc;               // level accumulators: A at [c], B at [Ct+c]
  int vrows_per_cta, nstrips, strip_w;  // dense work decomposition (virtual rows: H+1 per image)
  long ntiles;                           // td: 128-pixel tiles
  long npx;                              // pixels of the level (group stride of the planar gradient arena = 16 * npx)
  const __nv_bfloat16* xp;               // planar mirror of the activation arena (same layout as dact) or null
};
int launch_dgrad(const DgradArgs& a, int grid, cudaStream_t st);

enum { WG_DENSE = 0, WG_TD = 1, WG_TU = 2, WG_HEAD = 3 };
struct WgradArgs {
  // dW^T[c, n] = Σ_p A[p, c] · Bm[p, n]   (A optionally relu(bn(.)))
  int kind;
  const __nv_bfloat16* A;    // arena
  const __nv_bfloat16* Axp;  // planar mirror of the arena (rl_dev.cuh dact_index) or null: A is read through it
  int lda, a_off, Ca;        // Ca = valid A channels
  int transform;
  BnSrc bn;
  const __nv_bfloat16* Bm;   // dense: zero-padded dz16 planes; else row-major [px][ldb]
  long dz_npos, dz_guard;
  int ldb;
  int N;                     // total B columns (dense: 144)
  int part_n0[3], part_np[3];
  long npx;
  int H, W;                  // dense: level geometry for the tap gather
  long px_per_cta;
  float* gw;                 // gradient slot of the weight tensor (fp32 atomics)
  int Cin_w;                 // dense/td: Cin of the weight tensor; head: 288
  int cs_perm;
  int num_cls;
};
int launch_wgrad(const WgradArgs& a, int nruns, int nmb, int nparts, cudaStream_t st);

// ---------------------------------------------------------------- backward of a whole DenseBlock (rl_bwd_fused.cu)
// Input gradient of all five DenseLayers of a block with respect to the block's INPUT channels [0, C0):
//   dX[p, c] (+)= Σ_j γ_j r_c · relu'_j(p, c) · (dz_j ⊛ W_j)[p, c]
// one pass over x and dX instead of five; the BatchNorm-backward mean terms of these channels (A_c, B_c) are the
// pixel sums of the total it adds.  dγ_j / dβ_j come from wgrad_dense2_kernel.
struct DgradBulkArgs {
  int nl;                        // layers fused (kBlockLayers)
  const __nv_bfloat16* dz[5];    // zero-padded dz16 planes of layer j
  const uint8_t* wpk[5];         // [ncb][18][96][8] packed weights of layer j for the block-input channels
  const float* gamma[5];         // reference-ordered BatchNorm weights of layer j
  const float* beta[5];
  const double* stats;           // level stats (Σx, Σx²)
  double inv_n;
  int Ct, cs_perm;
  long dz_npos, dz_guard;
  const __nv_bfloat16* xp;       // planar activation mirror
  __nv_bfloat16* dact;           // planar gradient arena
  long npx;
  int C0, cbs;                   // block-input channels, channel-block size
  int H, W, B;
  int TP, ncolt;                 // tile pitch (8 / 16 / 32: TR = 128 / TP rows x TP - 2 columns per tile), tiles per row of tiles
  long ntiles;
  int tiles_per_cta;
  int rmw;
};
// tmaps[j]: 128-byte CUtensorMap blob of layer j's dz planes (encode_dz_tmap)
int launch_dgrad_bulk(const DgradBulkArgs& a, const void* const* tmaps, int nruns, int ncb, cudaStream_t st);
int encode_dz_tmap(void* tmap_out, const void* plane0, long npos, long guard, int W, long nvrows, int TP);
inline int bulk_tile_pitch(int W) { return W + 2 <= 8 ? 8 : (W + 2 <= 16 ? 16 : 32); }

// Weight gradients of the (up to five) DenseLayers of a block in one launch (grid.z = layer), plus dγ / dβ of their
// BatchNorms.  For 64 channels at a time two operands are built from x: m·x̂ and m (m = relu mask), so with
// D2 = (m·x̂)ᵀ·im2col(dz), D1 = mᵀ·im2col(dz):  dW = γ·D2 + β·D1,  dγ = Σ_n W·D2,  dβ = Σ_n W·D1.
struct WgDenseLayer {
  const uint8_t* wpk_bwd;        // [mb][18][128][8] packed weights (contracted with D1 / D2 for dβ / dγ)
  const float* gamma;            // reference-ordered
  const float* beta;
  float* g_gamma;
  float* g_beta;
  float* gw;
  const __nv_bfloat16* dz;       // zero-padded dz16 planes
  int Cin, cs_perm;
};
struct WgDenseArgs {
  WgDenseLayer L[5];
  int nl;
  const __nv_bfloat16* xp;       // planar activation mirror
  long npx;
  int H, W;
  long dz_npos, dz_guard;
  const double* stats;
  int Ct;
  double inv_n;
  long pos_per_cta;              // K range of a CTA in positions of the padded dz planes (multiple of 64)
  int B;
  double* acc;                   // level accumulators (A at [c], B at [Ct + c]); channels below C0 receive the layers' mean terms
  int C0;
};
// tmaps[j]: the 5-D im2col tensor map of layer j's dz planes (encode_dz_im2col_tmap)
int launch_wgrad_dense2(const WgDenseArgs& a, const void* const* tmaps, int nruns, int ngroups, cudaStream_t st);
int encode_dz_im2col_tmap(void* tmap_out, const void* plane0, long npos, long positions, int W);
// once per device / plan bind: opt the kernels into their dynamic shared-memory sizes
int configure_kernels();

// ---------------------------------------------------------------- CUDA-core kernels (rl_misc.cu)
int launch_pack_weights(const Plan* p, const float* params, cudaStream_t st);
int launch_first_fwd(const Plan* p, const float* x, const float* params, int train, cudaStream_t st);
int launch_channel_stats(const __nv_bfloat16* arena, long npx, int Ct, int c_off, int nch, double* stats, int sm,
                         cudaStream_t st);
int launch_label_hist(const void* y, int y_is_i64, long n, int num_cls, Scalars* sc, cudaStream_t st);

struct HeadArgs {
  const __nv_bfloat16* feat;  // level-0 arena (288 channels)
  int Ct;
  long npx;
  int HW;                     // pixels per image
  int num_cls;
  const float* w;             // [C][288] reference order
  const float* b;             // [C]
  const void* y;              // labels (uint8 or int64) or null
  int y_is_i64;
  float* probs;               // [B][C][H][W] fp32 or null
  uint8_t* mask_out;          // argmax [px] or null
  Scalars* sc;                // histogram in, loss / confusion out
  int weighted;               // 1: class-weighted CE (training_step), 0: plain CE (evaluate_batch)
  // backward (null dact = forward only)
  const float* dprobs;        // external d(loss)/d(probs) [B][C][H][W] or null -> fused CE gradient
  float grad_scale;           // multiplies the fused CE gradient (grad_output of the loss)
  const float* grad_scale_dev;  // optional device scalar multiplied in as well
  float feat_grad_sign;       // -1 for GradReverse between extractor and classifier (tiramisu.py:7-18)
  __nv_bfloat16* dact;
  __nv_bfloat16* dl;          // [px][16] bf16: d(logits)/||f||
  float* g_b;                 // classifier bias gradient slot
};
int launch_head(const HeadArgs& a, int sm, cudaStream_t st);

struct DzArgs {
  int kind;                   // 0: same-resolution slice (dense / first conv); 1: TD (max-pool scatter); 2: TU (column matrix)
  const __nv_bfloat16* act;   // producer's output arena
  const __nv_bfloat16* dact;
  const double* stats;        // of the output arena
  const double* acc;
  int Ct, c_off, nch;
  long npx_out;               // pixels of the output arena
  double inv_n_out;
  const float* mask;          // [B][nch] or null
  int HW_out;                 // pixels per image of the output arena
  __nv_bfloat16* dz;          // result
  int planes;                 // kind 0, nch = 16: write the zero-padded two-plane layout (H, W = level geometry)
  long dz_npos, dz_guard;
  float* g_bias;              // Σ dz per channel -> conv bias gradient (fp32 atomics)
  // TD: output arena is the pooled level; dz is written at the input resolution [px_in][nch]
  const __nv_bfloat16* xp;    // planar mirror of act (training; same indexing as dact) or null: read the NHWC arena
  const uint8_t* amax;
  int H, W, Hp, Wp, B;        // TD: input H,W and pooled Hp,Wp.  TU: input (low-res) h=Hp,w=Wp and output H,W
};
int launch_dz_prepare(const DzArgs& a, int sm, cudaStream_t st);

int launch_first_wgrad(const Plan* p, const float* x, const __nv_bfloat16* dz48, float* gw, cudaStream_t st);

struct BnTabEntry {
  long stats_byte_off;  // level stats base in workspace
  long rm, rv;          // float offsets into the BN-buffer arena
  int C, Ct, cs_perm;
  float inv_n, unbias;  // 1/N, N/(N-1)
};
int launch_bn_running(const Plan* p, float* bnbuf, float momentum, cudaStream_t st);
int launch_adamw(float* p, const float* g, float* m, float* v, long n, double lr, double b1, double b2, double eps, double wd,
                 long step, float grad_scale, cudaStream_t st);
int launch_sgd_nesterov(float* p, const float* g, float* buf, long n, float lr, float momentum, float wd, int first,
                        float grad_scale, cudaStream_t st);
int launch_fill_masks(float* masks, long n, float p_drop, unsigned long long seed, unsigned long long step,
                      const long long* step_dev, cudaStream_t st);
int launch_adamw_dev(float* p, const float* g, float* m, float* v, long n, const double* hyper, float grad_scale, cudaStream_t st);
// grads[i] += scale * (*scale_dev) * stage[i]   (end of a backward pass; scale_dev may be null)
int launch_grad_accum(float* grads, const float* stage, long n, float scale, const float* scale_dev, cudaStream_t st);

}  // namespace rl

namespace rl {
// ---------------------------------------------------------------- TransitionUp (rl_tu.cu)
// ConvTranspose2d(80 -> 80, k3, s2, p0) + bias + center_crop (offset 0) written straight into the concat slice
// (reference layers.py:58-69, 82-86) as ONE kernel: the four output parities (Y%2, X%2) of a low-resolution position are
// four accumulators fed by row-shifted views of one input halo tile; no column matrix, no col2im pass.
struct TuFwdArgs {
  const __nv_bfloat16* A;   // input level arena
  int lda, a_off;           // its channel count, first input channel
  int B, h, w;              // input resolution
  int H, W;                 // output resolution (cropped)
  const uint8_t* wpk;       // [tap 9][ci chunk 10][co 80][8 ci] bf16
  const float* bias;        // [80]
  __nv_bfloat16* out;       // output level arena (NHWC)
  int ldo, o_off;
  __nv_bfloat16* out_xp;    // planar mirror of the output level or null
  long npx_out;
  int TP, ncolt;            // tile pitch (8 / 16 / 32), tiles per row of tiles
  long ntiles;
};
int launch_tu_fwd(const TuFwdArgs& a, int grid, cudaStream_t st);
// backward of the transposed convolution from the four parity planes of dz (dz_tu_kernel, rl_misc.cu)
struct TuBwdArgs {
  const __nv_bfloat16* dzp;  // [4][npos][80], npos = B (h+1) (w+1)
  long npos;
  int B, h, w;               // input (low) resolution
  int TP, ncolt;
  long ntiles;
  const uint8_t* wpk;        // input gradient: [tap][co chunk][ci][8 co]
  __nv_bfloat16* dact;       // planar gradient arena of the input level (plain store of the 80 input channels)
  long npx_in;
  int in_off;
  const __nv_bfloat16* x;    // weight gradient: input level arena (NHWC)
  int ldx, x_off;
  float* gw;                 // [ci][co][3][3] fp32, accumulated with atomics
};
int launch_tu_dgrad(const TuBwdArgs& a, int grid, cudaStream_t st);
int launch_tu_wgrad(const TuBwdArgs& a, int grid_x, cudaStream_t st);
inline int tu_tile_pitch(int w) { return w + 2 <= 8 ? 8 : (w + 2 <= 16 ? 16 : 32); }
int configure_tu_kernels();
}  // namespace rl
