// rl_plan.cu — builds the layer schedule / arena geometry of FC-DenseNet67 (reference
// rightLaneNetwork/models/FCDenseNet/tiramisu.py:21-109, factory :178-183) for a fixed (B, H, W, C).
#include "rl_plan.h"

#include <stdio.h>
#include <string.h>

namespace rl {

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int td_parts(int C, int* n0, int* np) {
  if (C <= 256) {
    n0[0] = 0;
    np[0] = C;
    return 1;
  }
  int h = ((C / 2 + 15) / 16) * 16;
  n0[0] = 0;
  np[0] = h;
  n0[1] = h;
  np[1] = C - h;
  return 2;
}

int build_plan(Plan* p, int B, int H, int W, int num_cls, int training) {
  p->B = B;
  p->H = H;
  p->W = W;
  p->num_cls = num_cls;
  p->training = training != 0;
  p->ws = nullptr;
  p->side_stream = nullptr;
  for (int i = 0; i < 4; ++i) {
    p->ev_dz[i] = p->ev_wg[i] = nullptr;
    p->ev_wg_used[i] = false;
  }
  for (int i = 0; i < kLevels; ++i) {
    p->ev_blk[i] = nullptr;
    p->ev_blk_used[i] = false;
  }
  p->opt_overlap = p->opt_xp = p->opt_fwd_tma = p->opt_bwd_fused = 1;
  p->err[0] = 0;
  // B*H*W*16 < 2^31: the elementwise kernels index pixels (and pixel x chunk work items) with 32-bit arithmetic
  if (B < 1 || H < 32 || W < 32 || num_cls < 2 || num_cls > 4 || (long)B * H * W >= (1L << 27)) {
    snprintf(p->err, sizeof(p->err), "unsupported shape B=%d H=%d W=%d num_cls=%d (need H,W>=32, 2<=C<=4, B*H*W < 2^27)", B, H,
             W, num_cls);
    return -1;
  }
  // ---- levels
  int cskip[kLevels] = {128, 208, 288, 368, 448, 448};
  int h = H, w = W;
  for (int l = 0; l < kLevels; ++l) {
    p->lv[l].H = h;
    p->lv[l].W = w;
    p->lv[l].Cskip = cskip[l];
    p->lv[l].Ct = (l < 5) ? cskip[l] + 2 * kUpCh : cskip[l] + kUpCh;
    p->lv[l].px = (long)B * h * w;
    h /= 2;
    w /= 2;
  }
  // ---- layers in forward order
  p->layers.clear();
  auto mk = [&](int type, int li, int lo, int cin, int cout, int in_off, int out_off, int cs_perm) {
    Layer L;
    memset(&L, 0, sizeof(L));
    L.type = type;
    L.lvl_in = li;
    L.lvl_out = lo;
    L.cin = cin;
    L.cout = cout;
    L.in_off = in_off;
    L.out_off = out_off;
    L.cs_perm = cs_perm;
    L.p_gamma = L.p_beta = L.p_w = L.p_b = -1;
    L.b_rm = L.b_rv = -1;
    L.bn_index = -1;
    L.rmw_from = 0;
    L.blk = L.jblk = -1;
    p->layers.push_back(L);
    return (int)p->layers.size() - 1;
  };
  int id_first = mk(L_FIRST, 0, 0, 3, kFirst, 0, 0, 0);
  int id_down[5][5], id_td[5], id_bott[5], id_tu[5], id_up[5][5];
  int cin_l = kFirst;
  for (int l = 0; l < 5; ++l) {
    for (int j = 0; j < kBlockLayers; ++j) id_down[l][j] = mk(L_DENSE, l, l, cin_l + kGrowth * j, kGrowth, 0, cin_l + kGrowth * j, 0);
    id_td[l] = mk(L_TD, l, l + 1, cskip[l], cskip[l], 0, 0, 0);
    cin_l = cskip[l];
  }
  for (int j = 0; j < kBlockLayers; ++j) {
    id_bott[j] = mk(L_DENSE, 5, 5, 448 + kGrowth * j, kGrowth, 0, 448 + kGrowth * j, 0);
    if (j == kBlockLayers - 1) p->layers.back().rmw_from = 448;
  }
  for (int i = 0; i < 5; ++i) {
    int l = 4 - i;
    int src_off = (l + 1 == 5) ? 448 : cskip[l + 1] + kUpCh;
    id_tu[i] = mk(L_TU, l + 1, l, kUpCh, kUpCh, src_off, cskip[l], 0);
    for (int j = 0; j < kBlockLayers; ++j) {
      int cin = cskip[l] + kUpCh + kGrowth * j;
      id_up[i][j] = mk(L_DENSE, l, l, cin, kGrowth, 0, cin, cskip[l]);
      if (j == kBlockLayers - 1 && l != 0) p->layers.back().rmw_from = cskip[l] + kUpCh;
    }
  }
  // ---- DenseBlocks (consecutive runs of five dense layers on one level)
  p->blocks.clear();
  for (int i = 0; i + kBlockLayers <= (int)p->layers.size(); ++i) {
    if (p->layers[i].type != L_DENSE || p->layers[i].blk >= 0) continue;
    Block b;
    b.first = i;
    b.lvl = p->layers[i].lvl_in;
    b.C0 = p->layers[i].cin;
    b.rmw = p->layers[i + kBlockLayers - 1].rmw_from == 0 ? 1 : 0;
    b.ncb = (b.C0 + kBulkCb - 1) / kBulkCb;
    b.cbs = (((b.C0 + b.ncb - 1) / b.ncb) + 15) / 16 * 16;
    for (int j = 0; j < kBlockLayers; ++j) {
      p->layers[i + j].blk = (int)p->blocks.size();
      p->layers[i + j].jblk = j;
    }
    p->blocks.push_back(b);
  }
  // ---- parameter offsets in the reference's registration order (tiramisu.py:33-87, TrainingBase.py:32-33)
  long off = 0, boff = 0;
  int bn = 0;
  auto take = [&](long n) {
    long o = off;
    off += n;
    return o;
  };
  auto reg_bn_conv = [&](int id, int k) {
    Layer& L = p->layers[id];
    L.p_gamma = take(L.cin);
    L.p_beta = take(L.cin);
    L.p_w = take((long)L.cout * L.cin * k * k);
    L.p_b = take(L.cout);
    L.b_rm = boff;
    boff += L.cin;
    L.b_rv = boff;
    boff += L.cin;
    L.bn_index = bn++;
  };
  {
    Layer& L = p->layers[id_first];
    L.p_w = take(kFirst * 3 * 9);
    L.p_b = take(kFirst);
  }
  for (int l = 0; l < 5; ++l)
    for (int j = 0; j < 5; ++j) reg_bn_conv(id_down[l][j], 3);
  for (int l = 0; l < 5; ++l) reg_bn_conv(id_td[l], 1);
  for (int j = 0; j < 5; ++j) reg_bn_conv(id_bott[j], 3);
  for (int i = 0; i < 5; ++i) {
    Layer& L = p->layers[id_tu[i]];
    L.p_w = take(kUpCh * kUpCh * 9);
    L.p_b = take(kUpCh);
  }
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) reg_bn_conv(id_up[i][j], 3);
  p->p_cls_w = take((long)num_cls * 288);
  p->p_cls_b = take(num_cls);
  p->n_params = off;
  p->n_bnbuf = boff;

  // ---- workspace carve-up
  size_t ws = 0;
  auto carve = [&](size_t bytes) {
    size_t o = ws;
    ws = align_up(ws + bytes, 1024);
    return o;
  };
  p->scal_off = carve(1024);
  p->bntab_off = carve(64 * 64);
  p->packtab_off = carve(256 * sizeof(PackEntry));
  // packed weights
  p->wpk_off = ws;
  for (Layer& L : p->layers) {
    if (L.type == L_DENSE) {
      int nkc = (L.cin + 63) / 64, nmb = (L.cin + 127) / 128;
      L.wpk_fwd = carve((size_t)nkc * 18432);
      L.wpk_bwd = carve((size_t)nmb * 36864);
      L.wpk_chain = training ? carve(36864) : 0;
      L.wpk_bulk = training ? carve((size_t)p->blocks[L.blk].ncb * 18 * kBulkCb * 16) : 0;
    } else if (L.type == L_TD) {
      int nkc = (L.cin + 63) / 64, nmb = (L.cin + 127) / 128;
      L.wpk_fwd = carve((size_t)nkc * 8 * L.cin * 16);  // all parts: sum(np) = C rows per chunk-slot
      L.wpk_bwd = carve((size_t)nmb * nkc * 16384);
    } else if (L.type == L_TU) {
      L.wpk_fwd = carve((size_t)3 * 2 * 30720);
      L.wpk_bwd = carve((size_t)12 * 10240);
    }
  }
  p->wpk_bytes = ws - p->wpk_off;
  p->mask_off_all = ws;
  for (Layer& L : p->layers)
    if (L.type == L_DENSE || L.type == L_TD) L.mask_off = carve((size_t)B * L.cout * sizeof(float));
  p->mask_bytes = ws - p->mask_off_all;
  p->stats_off_all = ws;
  for (int l = 0; l < kLevels; ++l) p->lv[l].stats_off = carve((size_t)2 * p->lv[l].Ct * sizeof(double));
  p->stats_bytes = ws - p->stats_off_all;
  p->acc_off_all = ws;
  for (int l = 0; l < kLevels; ++l) p->lv[l].acc_off = carve((size_t)2 * p->lv[l].Ct * sizeof(double));
  p->acc_bytes = ws - p->acc_off_all;
  for (int l = 0; l < kLevels; ++l) p->lv[l].act_off = carve((size_t)p->lv[l].px * p->lv[l].Ct * 2);
  for (int l = 0; l < kLevels; ++l) p->lv[l].dact_off = training ? carve((size_t)p->lv[l].px * p->lv[l].Ct * 2) : 0;
  for (Layer& L : p->layers)
    if (L.type == L_TD) L.amax_off = training ? carve((size_t)p->lv[L.lvl_out].px * L.cout) : 0;
  // scratch (training): TransitionUp dz parity planes [4][B (h+1) (w+1)][80] bf16, TD output gradients [px(l) x C] bf16, dz16/dz48
  size_t tmp = 0;
  for (Layer& L : p->layers) {
    size_t need = 0;
    if (training) {
      if (L.type == L_TU) need = (size_t)4 * p->B * (p->lv[L.lvl_in].H + 1) * (p->lv[L.lvl_in].W + 1) * kUpCh * 2;
      if (L.type == L_TD) need = (size_t)p->lv[L.lvl_in].px * L.cout * 2;
      if (L.type == L_DENSE) need = (size_t)p->lv[L.lvl_in].px * 16 * 2;
      if (L.type == L_FIRST) need = (size_t)p->lv[0].px * kFirst * 2;
    }
    if (need > tmp) tmp = need;
  }
  p->tmp_bytes = tmp;
  p->tmp_off = carve(training ? 2 * align_up(tmp, 1024) : tmp);  // training: two buffers (side-stream weight gradients)
  p->xcol_off = training ? carve((size_t)p->lv[0].px * 32 * 2) : 0;
  p->dzp_off_all = ws;
  for (int l = 0; l < kLevels; ++l) {
    Level& L = p->lv[l];
    L.dzp_guard = L.W + 3;
    L.dzp_npos = L.dzp_guard + (long)B * (L.H + 1) * (L.W + 2) + 128 + 2 * (L.W + 2) + 8;
    L.dzp_off = training ? carve((size_t)L.dzp_npos * 32 * kDzBufs) : 0;
  }
  p->dzp_bytes = ws - p->dzp_off_all;
  // Planar mirror of the activation arenas (training): [Ct/16][px][16] like the gradient arenas.  Every producer writes
  // its channels to both; the backward's lane = pixel readers (input-gradient epilogues, dz_prepare) use the mirror,
  // where a pixel's 16-channel group is 32 contiguous bytes next to its neighbours' (a 16-channel slice of the NHWC
  // arena streams at 1.0 TB/s, profiles/r01_stream_probe.txt, and costs 32 L1 lines per warp access).
  for (int l = 0; l < kLevels; ++l) p->lv[l].xp_off = training ? carve((size_t)p->lv[l].px * p->lv[l].Ct * 2) : 0;
  p->dl_bytes = training ? (size_t)p->lv[0].px * 16 * 2 : 0;
  p->dl_off = carve(p->dl_bytes);
  p->gstage_off = training ? carve((size_t)p->n_params * sizeof(float)) : 0;
  p->ws_bytes = ws;
  return 0;
}

}  // namespace rl
