// rl_api.cu — extern "C" boundary (include/rightlane.h) and the host-side layer schedule.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "rl_common.cuh"
#include "rl_kernels.h"
#include "rl_plan.h"

using namespace rl;

struct rl_plan {
  Plan p;
};

namespace rl {
int td_parts(int C, int* n0, int* np);
int configure_fwd_kernels();
int configure_bwd_kernels();
int configure_fused_kernels();
int configure_kernels() {
  int rc;
  if ((rc = configure_fwd_kernels())) return rc;
  if ((rc = configure_bwd_kernels())) return rc;
  if ((rc = configure_tu_kernels())) return rc;
  if ((rc = configure_td_kernels())) return rc;
  return configure_fused_kernels();
}
}

static inline uint8_t* wsb(const Plan* p) { return (uint8_t*)p->ws; }
static inline __nv_bfloat16* act_of(const Plan* p, int l) { return (__nv_bfloat16*)(wsb(p) + p->lv[l].act_off); }
static inline __nv_bfloat16* dact_of(const Plan* p, int l) { return (__nv_bfloat16*)(wsb(p) + p->lv[l].dact_off); }
// planar mirror of a level's activations (training plans only; option "xp" = 0 keeps every reader on the NHWC arena)
static inline __nv_bfloat16* xp_of(const Plan* p, int l) {
  return (p->training && p->opt_xp) ? (__nv_bfloat16*)(wsb(p) + p->lv[l].xp_off) : nullptr;
}
static inline float* gstage_of(const Plan* p) { return (float*)(wsb(p) + p->gstage_off); }
static inline double* stats_of(const Plan* p, int l) { return (double*)(wsb(p) + p->lv[l].stats_off); }
static inline double* acc_of(const Plan* p, int l) { return (double*)(wsb(p) + p->lv[l].acc_off); }
static inline Scalars* scal_of(const Plan* p) { return (Scalars*)(wsb(p) + p->scal_off); }

static BnSrc make_bn(const Plan* p, const Layer& L, const float* params, const float* bnbuf, int train) {
  BnSrc b;
  b.stats = train ? stats_of(p, L.lvl_in) : nullptr;
  b.Ct = p->lv[L.lvl_in].Ct;
  b.inv_n = 1.0 / (double)p->lv[L.lvl_in].px;
  b.gamma = params + L.p_gamma;
  b.beta = params + L.p_beta;
  b.rmean = bnbuf ? bnbuf + L.b_rm : nullptr;
  b.rvar = bnbuf ? bnbuf + L.b_rv : nullptr;
  b.cs_perm = L.cs_perm;
  b.eval = train ? 0 : 1;
  return b;
}

extern "C" {

int rl_plan_create(rl_plan** out, int B, int H, int W, int num_cls, int training) {
  if (!out) return rl_fail(RL_ERR_ARG, "rl_plan_create: out is null");
  rl_plan* h = new (std::nothrow) rl_plan();
  if (!h) return rl_fail(RL_ERR_STATE, "rl_plan_create: out of host memory");
  if (build_plan(&h->p, B, H, W, num_cls, training) != 0) {
    int rc = rl_fail(RL_ERR_SHAPE, "%s", h->p.err);
    delete h;
    return rc;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
  } else {
    cudaGetLastError();
  }
  h->p.sm_count = sms;
  *out = h;
  return RL_OK;
}

void rl_plan_destroy(rl_plan* h) {
  if (!h) return;
  Plan& p = h->p;
  for (int i = 0; i < 4; ++i) {
    if (p.ev_dz[i]) cudaEventDestroy((cudaEvent_t)p.ev_dz[i]);
    if (p.ev_wg[i]) cudaEventDestroy((cudaEvent_t)p.ev_wg[i]);
  }
  for (int i = 0; i < kLevels; ++i)
    if (p.ev_blk[i]) cudaEventDestroy((cudaEvent_t)p.ev_blk[i]);
  if (p.side_stream) cudaStreamDestroy((cudaStream_t)p.side_stream);
  delete h;
}
long rl_plan_workspace_bytes(const rl_plan* h) { return h ? (long)h->p.ws_bytes : 0; }
long rl_plan_num_params(const rl_plan* h) { return h ? h->p.n_params : 0; }
long rl_plan_num_bn_buffers(const rl_plan* h) { return h ? h->p.n_bnbuf : 0; }

int rl_plan_query(const rl_plan* h, int what, int index, long* out) {
  if (!h || !out) return rl_fail(RL_ERR_ARG, "rl_plan_query: null argument");
  const Plan& p = h->p;
  const bool lvl_ok = index >= 0 && index < kLevels, lay_ok = index >= 0 && index < (int)p.layers.size();
  switch (what) {
    case 0: *out = (long)p.layers.size(); return RL_OK;
    case 1: if (!lvl_ok) break; *out = (long)p.lv[index].act_off; return RL_OK;
    case 2: if (!lvl_ok) break; *out = (long)p.lv[index].dact_off; return RL_OK;
    case 3: if (!lvl_ok) break; *out = p.lv[index].Ct; return RL_OK;
    case 4: if (!lvl_ok) break; *out = p.lv[index].H; return RL_OK;
    case 5: if (!lvl_ok) break; *out = p.lv[index].W; return RL_OK;
    case 6: if (!lvl_ok) break; *out = p.lv[index].Cskip; return RL_OK;
    case 7: *out = (long)p.scal_off; return RL_OK;
    case 8: if (!lay_ok) break; *out = p.layers[index].type; return RL_OK;
    case 9: if (!lay_ok) break; *out = (long)p.layers[index].mask_off; return RL_OK;
    case 10: if (!lay_ok) break; *out = p.layers[index].cout; return RL_OK;
    case 11: if (!lay_ok) break; *out = p.layers[index].p_w; return RL_OK;
    case 12: if (!lvl_ok) break; *out = (long)p.lv[index].stats_off; return RL_OK;
    case 13: *out = (long)p.gstage_off; return RL_OK;
    case 14: if (!lvl_ok) break; *out = (long)p.lv[index].xp_off; return RL_OK;
    case 15: if (!lvl_ok) break; *out = (long)p.lv[index].acc_off; return RL_OK;
    case 16: if (!lay_ok) break; *out = p.layers[index].lvl_in; return RL_OK;
    case 17: if (!lay_ok) break; *out = p.layers[index].lvl_out; return RL_OK;
    case 18: if (!lay_ok) break; *out = p.layers[index].cin; return RL_OK;
    case 19: if (!lay_ok) break; *out = p.layers[index].in_off; return RL_OK;
    case 20: if (!lay_ok) break; *out = p.layers[index].out_off; return RL_OK;
    case 21: if (!lay_ok) break; *out = p.layers[index].cs_perm; return RL_OK;
    case 22: if (!lay_ok) break; *out = (long)p.layers[index].amax_off; return RL_OK;
    case 23: if (!lay_ok) break; *out = p.layers[index].rmw_from; return RL_OK;
    case 24: if (!lay_ok) break; *out = p.layers[index].p_gamma; return RL_OK;
    case 25: if (!lay_ok) break; *out = p.layers[index].p_b; return RL_OK;
    case 26: if (!lay_ok) break; *out = p.layers[index].b_rm; return RL_OK;
    case 27: *out = (long)p.dl_off; return RL_OK;
    default: return rl_fail(RL_ERR_ARG, "rl_plan_query: unknown selector %d", what);
  }
  return rl_fail(RL_ERR_ARG, "rl_plan_query: index %d out of range for selector %d", index, what);
}

int rl_plan_bind(rl_plan* h, void* workspace, void* stream) {
  if (!h || !workspace) return rl_fail(RL_ERR_ARG, "rl_plan_bind: null argument");
  if (((uintptr_t)workspace & 255) != 0) return rl_fail(RL_ERR_ARG, "rl_plan_bind: workspace must be 256-byte aligned");
  Plan& p = h->p;
  p.ws = workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // ---- weight packing table
  std::vector<PackEntry> pk;
  for (const Layer& L : p.layers) {
    PackEntry e;
    memset(&e, 0, sizeof(e));
    e.cin = L.cin;
    e.cs_perm = L.cs_perm;
    e.p_w = L.p_w;
    if (L.type == L_DENSE) {
      int nkc = (L.cin + 63) / 64, nmb = (L.cin + 127) / 128;
      e.kind = 0; e.dst_off = (long)L.wpk_fwd; e.n16 = (long)nkc * 8 * 144; pk.push_back(e);
      e.kind = 1; e.dst_off = (long)L.wpk_bwd; e.n16 = (long)nmb * 18 * 128; pk.push_back(e);
    } else if (L.type == L_TD) {
      int nkc = (L.cin + 63) / 64, nmb = (L.cin + 127) / 128;
      int n0[3], np[3];
      td_parts(L.cin, n0, np);
      e.np0 = np[0];
      e.kind = 2; e.dst_off = (long)L.wpk_fwd; e.n16 = (long)nkc * 8 * L.cin; pk.push_back(e);
      e.kind = 3; e.dst_off = (long)L.wpk_bwd; e.n16 = (long)nmb * nkc * 8 * 128; pk.push_back(e);
    } else if (L.type == L_TU) {
      e.kind = 8; e.dst_off = (long)L.wpk_fwd; e.n16 = 9L * 10 * 80; pk.push_back(e);
      e.kind = 9; e.dst_off = (long)L.wpk_bwd; e.n16 = 9L * 10 * 80; pk.push_back(e);
    }
    if (L.type == L_DENSE && p.training) {
      const Block& b = p.blocks[L.blk];
      if (L.jblk > 0) {
        e.kind = 6; e.dst_off = (long)L.wpk_chain; e.n16 = 18L * 128; e.c_lo = b.C0; e.nch = kGrowth * L.jblk; pk.push_back(e);
      }
      e.kind = 7; e.dst_off = (long)L.wpk_bulk; e.n16 = (long)b.ncb * 18 * kBulkCb; e.c_lo = b.C0; e.nch = b.cbs; pk.push_back(e);
    }
  }
  if (pk.size() > 256) return rl_fail(RL_ERR_STATE, "rl_plan_bind: pack table overflow");
  p.n_pack = (int)pk.size();
  RL_CUDA(cudaMemcpyAsync(wsb(&p) + p.packtab_off, pk.data(), pk.size() * sizeof(PackEntry), cudaMemcpyHostToDevice, st));
  // ---- BN table
  std::vector<BnTabEntry> bt;
  for (const Layer& L : p.layers) {
    if (L.bn_index < 0) continue;
    BnTabEntry e;
    const Level& lv = p.lv[L.lvl_in];
    e.stats_byte_off = (long)lv.stats_off;
    e.rm = L.b_rm;
    e.rv = L.b_rv;
    e.C = L.cin;
    e.Ct = lv.Ct;
    e.cs_perm = L.cs_perm;
    e.inv_n = (float)(1.0 / (double)lv.px);
    e.unbias = lv.px > 1 ? (float)((double)lv.px / (double)(lv.px - 1)) : 1.f;
    bt.push_back(e);
  }
  p.n_bn = (int)bt.size();
  RL_CUDA(cudaMemcpyAsync(wsb(&p) + p.bntab_off, bt.data(), bt.size() * sizeof(BnTabEntry), cudaMemcpyHostToDevice, st));
  if (p.training && p.dzp_bytes) RL_CUDA(cudaMemsetAsync(wsb(&p) + p.dzp_off_all, 0, p.dzp_bytes, st));
  // ---- tensor maps of the level arenas (activation tiles of the DenseLayer forward go through the TMA engine)
  p.tmap_ok = true;
  for (int l = 0; l < kLevels; ++l) {
    int rc = encode_arena_tmap(p.tmap_act[l], act_of(&p, l), p.lv[l].px, p.lv[l].Ct, p.lv[l].W);
    if (rc) return rc;
  }
  {
    int rc = configure_kernels();  // per device: a process may bind plans on several GPUs
    if (rc) return rc;
  }
  if (p.training)
    for (int l = 0; l < kLevels; ++l)
      for (int q = 0; q < kDzBufs; ++q) {
        const Level& lv = p.lv[l];
        int rc = encode_dz_tmap(p.tmap_dz[l][q], wsb(&p) + lv.dzp_off + (size_t)q * lv.dzp_npos * 32, lv.dzp_npos, lv.dzp_guard,
                                lv.W, (long)p.B * (lv.H + 1), bulk_tile_pitch(lv.W));
        if (rc) return rc;
        rc = encode_dz_im2col_tmap(p.tmap_dz_im2col[l][q], wsb(&p) + lv.dzp_off + (size_t)q * lv.dzp_npos * 32, lv.dzp_npos,
                                   lv.dzp_npos - 2 * (lv.W + 2) - 2, lv.W);
        if (rc) return rc;
      }
  if (p.training && !p.side_stream) {
    cudaStream_t s2;
    RL_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    p.side_stream = s2;
    for (int i = 0; i < 4; ++i) {
      cudaEvent_t e;
      RL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p.ev_dz[i] = e;
      RL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p.ev_wg[i] = e;
    }
    for (int i = 0; i < kLevels; ++i) {
      cudaEvent_t e;
      RL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p.ev_blk[i] = e;
    }
  }
  RL_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
  // dropout scales default to 1 (no dropout)
  return rl_set_dropout(h, 0.f, 0, 0, stream);
}

int rl_pack_weights(rl_plan* h, const float* params, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_pack_weights: plan not bound");
  return launch_pack_weights(&h->p, params, (cudaStream_t)stream);
}

int rl_set_dropout(rl_plan* h, float p_drop, unsigned long long seed, unsigned long long step, void* stream) {
  return rl_set_dropout_dev(h, p_drop, seed, step, nullptr, stream);
}

int rl_set_dropout_dev(rl_plan* h, float p_drop, unsigned long long seed, unsigned long long step, const long long* step_dev,
                       void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_set_dropout: plan not bound");
  if (p_drop < 0.f || p_drop >= 1.f) return rl_fail(RL_ERR_ARG, "rl_set_dropout: p_drop %f out of range", p_drop);
  return launch_fill_masks((float*)(wsb(&h->p) + h->p.mask_off_all), (long)(h->p.mask_bytes / sizeof(float)), p_drop, seed,
                           step, step_dev, (cudaStream_t)stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// forward schedule: one call per layer of Plan::layers (forward order)
// ------------------------------------------------------------------------------------------------
static int forward_layer(Plan* p, int li, const float* params, const float* bnbuf, const float* x, int train, cudaStream_t st) {
  const Layer& L = p->layers[li];
  const Level& lin = p->lv[L.lvl_in];
  const Level& lo = p->lv[L.lvl_out];
  const int sm = p->sm_count;
  int rc;
  if (L.type == L_FIRST) {
    if ((rc = launch_first_fwd(p, x, params, train && p->opt_xp, st))) return rc;
    if (train && (rc = launch_channel_stats(act_of(p, 0), lo.px, lo.Ct, 0, kFirst, stats_of(p, 0), sm, st))) return rc;
  } else if (L.type == L_DENSE) {
    DenseFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.act = act_of(p, L.lvl_in);
    a.act_out = act_of(p, L.lvl_in);
    a.Ct = lin.Ct;
    a.Cin = L.cin;
    a.H = lin.H;
    a.W = lin.W;
    a.rows_total = p->B * lin.H;
    const bool use_tma = p->opt_fwd_tma && p->tmap_ok;
    if (use_tma && lin.W > 160) {  // strips whose 128-column input window (strip + 2 halo columns) is one TMA box row
      a.nstrips = (lin.W + 125) / 126;
      a.strip_w = (lin.W + a.nstrips - 1) / a.nstrips;
    } else {
      a.strip_w = 160;
      a.nstrips = (lin.W + a.strip_w - 1) / a.strip_w;
    }
    int nruns = std::max(1, std::min(a.rows_total, ((use_tma ? 1 : 2) * sm) / a.nstrips));
    a.rows_per_cta = (a.rows_total + nruns - 1) / nruns;
    nruns = (a.rows_total + a.rows_per_cta - 1) / a.rows_per_cta;
    a.wpk = wsb(p) + L.wpk_fwd;
    a.bias = params + L.p_b;
    a.mask = train ? (const float*)(wsb(p) + L.mask_off) : nullptr;
    a.bn = make_bn(p, L, params, bnbuf, train);
    a.stats_out = train ? stats_of(p, L.lvl_in) : nullptr;
    a.x16 = (train && xp_of(p, L.lvl_in)) ? xp_of(p, L.lvl_in) + (size_t)(L.cin / 16) * lin.px * 16 : nullptr;
    if (use_tma) {
      if ((rc = launch_dense_fwd_tma(a, p->tmap_act[L.lvl_in], nruns * a.nstrips, st))) return rc;
    } else if ((rc = launch_dense_fwd(a, nruns * a.nstrips, st))) {
      return rc;
    }
  } else if (L.type == L_TD) {
    GemmPxArgs a;
    memset(&a, 0, sizeof(a));
    a.A = act_of(p, L.lvl_in);
    a.lda = lin.Ct;
    a.a_off = 0;
    a.K = L.cin;
    a.H = lin.H;
    a.W = lin.W;
    a.Hp = lo.H;
    a.Wp = lo.W;
    a.wblocks = (lo.W + 31) / 32;
    a.ntiles = (long)p->B * lo.H * a.wblocks;
    a.transform = 1;
    a.bn = make_bn(p, L, params, bnbuf, train);
    a.wpk = wsb(p) + L.wpk_fwd;
    a.N = L.cout;
    int nparts = td_parts(L.cout, a.part_n0, a.part_np);
    a.bias = params + L.p_b;
    a.mask = train ? (const float*)(wsb(p) + L.mask_off) : nullptr;
    a.out = act_of(p, L.lvl_out);
    a.ldo = lo.Ct;
    a.o_off = 0;
    a.amax = (train && p->training) ? wsb(p) + L.amax_off : nullptr;
    a.out_xp = train ? xp_of(p, L.lvl_out) : nullptr;
    a.out_xp_npx = lo.px;
    int npmax = 0;
    for (int q = 0; q < nparts; ++q) npmax = std::max(npmax, a.part_np[q]);
    if (td_fwd_fits(a.K, a.part_np, nparts)) {
      if ((rc = launch_td_fwd(a, (int)std::max(1L, std::min(a.ntiles, (long)sm / nparts)), nparts, st))) return rc;
    } else {
      int gx = (int)std::max(1L, std::min(a.ntiles, (long)((npmax <= 128 ? 3 : 2) * sm) / nparts));
      if ((rc = launch_gemm_px(a, gx, nparts, st))) return rc;
    }
    if (train && (rc = launch_channel_stats(act_of(p, L.lvl_out), lo.px, lo.Ct, 0, L.cout, stats_of(p, L.lvl_out), sm, st)))
      return rc;
  } else {  // L_TU: all four output parities of a tile of low-resolution positions in one kernel (rl_tu.cu)
    TuFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.A = act_of(p, L.lvl_in);
    a.lda = lin.Ct;
    a.a_off = L.in_off;
    a.B = p->B;
    a.h = lin.H;
    a.w = lin.W;
    a.H = lo.H;
    a.W = lo.W;
    a.wpk = wsb(p) + L.wpk_fwd;
    a.bias = params + L.p_b;
    a.out = act_of(p, L.lvl_out);
    a.ldo = lo.Ct;
    a.o_off = L.out_off;
    a.out_xp = train ? xp_of(p, L.lvl_out) : nullptr;
    a.npx_out = lo.px;
    a.TP = tu_tile_pitch(lin.W);
    const int TR = 128 / a.TP;
    a.ncolt = (lin.W + 1 + a.TP - 2) / (a.TP - 1);
    a.ntiles = (long)((p->B * (lin.H + 1) + TR - 1) / TR) * a.ncolt;
    if ((rc = launch_tu_fwd(a, (int)std::max(1L, std::min(a.ntiles, (long)sm)), st))) return rc;
    if (train && (rc = launch_channel_stats(act_of(p, L.lvl_out), lo.px, lo.Ct, L.out_off, kUpCh, stats_of(p, L.lvl_out), sm, st)))
      return rc;
  }
  return RL_OK;
}

extern "C" int rl_forward_range(rl_plan* h, const float* params, const float* bnbuf, const float* x, int train, int layer_lo,
                                int layer_hi, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_forward: plan not bound");
  if (!params || (!train && !bnbuf)) return rl_fail(RL_ERR_ARG, "rl_forward: null argument");
  Plan* p = &h->p;
  const int n = (int)p->layers.size();
  if (layer_lo < 0 || layer_hi >= n || layer_lo > layer_hi) return rl_fail(RL_ERR_ARG, "rl_forward_range: bad layer range");
  if (layer_lo == 0 && !x) return rl_fail(RL_ERR_ARG, "rl_forward: null input");
  cudaStream_t st = (cudaStream_t)stream;
  for (int li = layer_lo; li <= layer_hi; ++li) {
    int rc = forward_layer(p, li, params, bnbuf, x, train, st);
    if (rc) return rc;
  }
  return RL_OK;
}

extern "C" int rl_forward(rl_plan* h, const float* params, const float* bnbuf, const float* x, int train, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_forward: plan not bound");
  Plan* p = &h->p;
  if (train) RL_CUDA(cudaMemsetAsync(wsb(p) + p->stats_off_all, 0, p->stats_bytes, (cudaStream_t)stream));
  return rl_forward_range(h, params, bnbuf, x, train, 0, (int)p->layers.size() - 1, stream);
}

extern "C" int rl_head(rl_plan* h, const float* params, const void* y, int y_is_i64, int weighted, float* probs, uint8_t* mask,
                       int backward, const float* dprobs, float feat_grad_sign, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_head: plan not bound");
  Plan* p = &h->p;
  cudaStream_t st = (cudaStream_t)stream;
  if (!params) return rl_fail(RL_ERR_ARG, "rl_head: null parameters");
  if (backward && !p->training) return rl_fail(RL_ERR_STATE, "rl_head: backward needs a training plan");
  if (backward == 2 && !dprobs) return rl_fail(RL_ERR_ARG, "rl_head: backward=2 needs dprobs");
  if (backward == 1 && !y) return rl_fail(RL_ERR_ARG, "rl_head: fused CE backward needs labels");
  RL_CUDA(cudaMemsetAsync(scal_of(p), 0, sizeof(Scalars), st));
  // a backward pass starts here: its raw gradients are staged and scaled into the caller's arena by rl_backward
  if (backward) RL_CUDA(cudaMemsetAsync(gstage_of(p), 0, (size_t)p->n_params * sizeof(float), st));
  int rc;
  if (y && (rc = launch_label_hist(y, y_is_i64, p->lv[0].px, p->num_cls, scal_of(p), st))) return rc;
  HeadArgs a;
  memset(&a, 0, sizeof(a));
  a.feat = act_of(p, 0);
  a.Ct = p->lv[0].Ct;
  a.npx = p->lv[0].px;
  a.HW = p->H * p->W;
  a.num_cls = p->num_cls;
  a.w = params + p->p_cls_w;
  a.b = params + p->p_cls_b;
  a.y = y;
  a.y_is_i64 = y_is_i64;
  a.probs = probs;
  a.mask_out = mask;
  a.sc = scal_of(p);
  a.weighted = weighted;
  a.dprobs = backward == 2 ? dprobs : nullptr;
  a.grad_scale = 1.f;
  a.grad_scale_dev = nullptr;
  a.feat_grad_sign = feat_grad_sign;
  a.dact = backward ? dact_of(p, 0) : nullptr;
  a.dl = backward ? (__nv_bfloat16*)(wsb(p) + p->dl_off) : nullptr;
  a.g_b = backward ? gstage_of(p) + p->p_cls_b : nullptr;
  return launch_head(a, p->sm_count, st);
}

extern "C" {

int rl_bn_running_update(rl_plan* h, float* bnbuf, float momentum, void* stream) {
  if (!h || !h->p.ws || !bnbuf) return rl_fail(RL_ERR_STATE, "rl_bn_running_update: plan not bound / null buffers");
  return launch_bn_running(&h->p, bnbuf, momentum, (cudaStream_t)stream);
}

int rl_adamw_step(float* params, const float* grads, float* m, float* v, long n, double lr, double b1, double b2, double eps,
                  double wd, long step, float grad_scale, void* stream) {
  if (!params || !grads || !m || !v || n <= 0 || step < 1) return rl_fail(RL_ERR_ARG, "rl_adamw_step: bad argument");
  return launch_adamw(params, grads, m, v, n, lr, b1, b2, eps, wd, step, grad_scale, (cudaStream_t)stream);
}

int rl_adamw_step_dev(float* params, const float* grads, float* m, float* v, long n, const double* hyper_dev, float grad_scale,
                      void* stream) {
  if (!params || !grads || !m || !v || !hyper_dev || n <= 0) return rl_fail(RL_ERR_ARG, "rl_adamw_step_dev: bad argument");
  return launch_adamw_dev(params, grads, m, v, n, hyper_dev, grad_scale, (cudaStream_t)stream);
}

int rl_sgd_nesterov_step(float* params, const float* grads, float* buf, long n, float lr, float momentum, float wd,
                         int first_step, float grad_scale, void* stream) {
  if (!params || !grads || !buf || n <= 0) return rl_fail(RL_ERR_ARG, "rl_sgd_nesterov_step: bad argument");
  return launch_sgd_nesterov(params, grads, buf, n, lr, momentum, wd, first_step, grad_scale, (cudaStream_t)stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// backward schedule: reverse of the forward layer list.  Every gradient of one pass goes to the plan's staging arena;
// rl_backward scales it into the caller's arena at the end (grads += scale * stage).
// ------------------------------------------------------------------------------------------------
namespace {

struct BwdCtx {
  Plan* p;
  const float* params;
  float* gs;  // staging arena
  cudaStream_t st, st2;
  int overlap;
  int dense_count, other_count;
  size_t tmp_stride;
};

// pixel range per CTA of a weight-gradient launch.  Every CTA ends with a full set of atomics for its [channels x N] block, so
// a CTA gets at least kMinChunks 64-pixel chunks: at the small levels the atomics, not the MMAs, are the cost.
long wgrad_px_per_cta(long npx, long want_runs, int* nruns) {
  constexpr long kMinChunks = 8;
  const long chunks = (npx + 63) / 64;
  const long runs = std::max(1L, std::min(want_runs, (chunks + kMinChunks - 1) / kMinChunks));
  const long cpr = (chunks + runs - 1) / runs;
  *nruns = (int)((chunks + cpr - 1) / cpr);
  return cpr * 64;
}

int wgrad_runs(const Plan* p, long npx, int nmb, int nparts, long* px_per_cta, int ctas_per_sm = 2) {
  int nruns;
  *px_per_cta = wgrad_px_per_cta(npx, std::max(1L, (long)(ctas_per_sm * p->sm_count) / std::max(1, nmb * nparts)), &nruns);
  return nruns;
}

// fork / join helpers for a weight gradient that reads scratch buffer q on the side stream
int before_dz(BwdCtx& c, int q) {
  if (c.overlap && c.p->ev_wg_used[q]) RL_CUDA(cudaStreamWaitEvent(c.st, (cudaEvent_t)c.p->ev_wg[q], 0));  // buffer q free again
  return RL_OK;
}
int after_dz(BwdCtx& c, int q) {
  if (c.overlap) {
    RL_CUDA(cudaEventRecord((cudaEvent_t)c.p->ev_dz[q], c.st));
    RL_CUDA(cudaStreamWaitEvent(c.st2, (cudaEvent_t)c.p->ev_dz[q], 0));
  }
  return RL_OK;
}
int after_wg(BwdCtx& c, int q) {
  if (c.overlap) {
    RL_CUDA(cudaEventRecord((cudaEvent_t)c.p->ev_wg[q], c.st2));
    c.p->ev_wg_used[q] = true;
  }
  return RL_OK;
}
// the padded dz16 planes of a level are about to be rewritten: every side-stream reader of them must be done
int before_level_planes(BwdCtx& c, int lvl) {
  if (!c.overlap) return RL_OK;
  if (c.p->ev_blk_used[lvl]) RL_CUDA(cudaStreamWaitEvent(c.st, (cudaEvent_t)c.p->ev_blk[lvl], 0));
  for (int q = 0; q < 2; ++q)
    if (c.p->ev_wg_used[q]) RL_CUDA(cudaStreamWaitEvent(c.st, (cudaEvent_t)c.p->ev_wg[q], 0));
  return RL_OK;
}

void fill_dz_common(const BwdCtx& c, const Layer& L, DzArgs& z) {
  const Plan* p = c.p;
  const Level& lout = p->lv[L.lvl_out];
  memset(&z, 0, sizeof(z));
  z.act = act_of(p, L.lvl_out);
  z.dact = dact_of(p, L.lvl_out);
  z.stats = stats_of(p, L.lvl_out);
  z.acc = acc_of(p, L.lvl_out);
  z.Ct = lout.Ct;
  z.c_off = L.out_off;
  z.nch = L.cout;
  z.npx_out = lout.px;
  z.inv_n_out = 1.0 / (double)lout.px;
  z.HW_out = lout.H * lout.W;
  z.g_bias = c.gs + L.p_b;
  z.B = p->B;
}

// dz of a DenseLayer into plane buffer q of its level
int dense_dz(BwdCtx& c, const Layer& L, int q, DzArgs& z) {
  const Plan* p = c.p;
  const Level& lin = p->lv[L.lvl_in];
  fill_dz_common(c, L, z);
  z.xp = xp_of(p, L.lvl_out);
  z.kind = 0;
  z.mask = (const float*)(wsb(p) + L.mask_off);
  z.planes = 1;
  z.dz = (__nv_bfloat16*)(wsb(p) + lin.dzp_off + (size_t)q * lin.dzp_npos * 32);
  z.dz_npos = lin.dzp_npos;
  z.dz_guard = lin.dzp_guard;
  z.H = lin.H;
  z.W = lin.W;
  return launch_dz_prepare(z, p->sm_count, c.st);
}

// input gradient of a DenseLayer for arena channels [c_lo, c_lo + nch) (per-layer kernel)
int dense_dgrad(BwdCtx& c, const Layer& L, const __nv_bfloat16* dz, const uint8_t* wpk, int c_lo, int nch, int rmw_from,
                bool param_grads) {
  const Plan* p = c.p;
  const Level& lin = p->lv[L.lvl_in];
  const int sm = p->sm_count;
  DgradArgs d;
  memset(&d, 0, sizeof(d));
  d.dense = 1;
  d.dz = dz;
  d.dz_npos = lin.dzp_npos;
  d.dz_guard = lin.dzp_guard;
  d.Kdz = 16;
  d.npx = lin.px;
  d.xp = xp_of(p, L.lvl_in);
  d.wpk = wpk;
  d.act = act_of(p, L.lvl_in);
  d.dact = dact_of(p, L.lvl_in);
  d.Ct = lin.Ct;
  d.Cin = nch;
  d.c_lo = c_lo;
  d.H = lin.H;
  d.W = lin.W;
  d.B = p->B;
  d.rmw_from = rmw_from;
  d.bn = make_bn(p, L, c.params, nullptr, 1);
  d.g_gamma = param_grads ? c.gs + L.p_gamma : nullptr;
  d.g_beta = param_grads ? c.gs + L.p_beta : nullptr;
  d.acc = acc_of(p, L.lvl_in);
  d.strip_w = 160;
  d.nstrips = (lin.W + d.strip_w - 1) / d.strip_w;
  const int vrows = p->B * (lin.H + 1);
  // small levels run their channel blocks in parallel CTAs (launch_dgrad splits over grid.y): keep one wave in total
  const int nmb_d = (nch + 127) / 128;
  const bool split = nmb_d > 1 && (long)vrows * (lin.W + 2) <= 40000;
  int nruns = std::max(1, std::min(vrows, (3 * sm) / (d.nstrips * (split ? nmb_d : 1))));
  d.vrows_per_cta = (vrows + nruns - 1) / nruns;
  nruns = (vrows + d.vrows_per_cta - 1) / d.vrows_per_cta;
  return launch_dgrad(d, nruns * d.nstrips, c.st);
}

int head_wgrad(BwdCtx& c) {  // classifier weight gradient: dW[k, c] = sum_p dl[p, k] * f[p, c]
  Plan* p = c.p;
  WgradArgs w;
  memset(&w, 0, sizeof(w));
  w.kind = WG_HEAD;
  w.A = act_of(p, 0);
  w.Axp = xp_of(p, 0);
  w.lda = p->lv[0].Ct;
  w.a_off = 0;
  w.Ca = 288;
  w.transform = 0;
  w.Bm = (const __nv_bfloat16*)(wsb(p) + p->dl_off);
  w.ldb = 16;
  w.N = 16;
  w.part_n0[0] = 0;
  w.part_np[0] = 16;
  w.npx = p->lv[0].px;
  w.H = p->H;
  w.W = p->W;
  w.gw = c.gs + p->p_cls_w;
  w.Cin_w = 288;
  w.num_cls = p->num_cls;
  int nruns = wgrad_runs(p, w.npx, 3, 1, &w.px_per_cta);
  return launch_wgrad(w, nruns, 3, 1, c.st);
}

// one layer through the per-layer kernels (TransitionDown / TransitionUp / first conv always; DenseLayers when the
// block-fused path is off or not applicable)
int backward_layer(BwdCtx& c, int li) {
  Plan* p = c.p;
  const Layer& L = p->layers[li];
  const Level& lin = p->lv[L.lvl_in];
  const Level& lout = p->lv[L.lvl_out];
  const int sm = p->sm_count;
  int rc;
  if (L.type == L_DENSE) {
    const int q = c.dense_count++ & 1;  // dz plane buffer of this layer
    if (c.overlap && p->ev_blk_used[L.lvl_in]) {
      RL_CUDA(cudaStreamWaitEvent(c.st, (cudaEvent_t)p->ev_blk[L.lvl_in], 0));
      p->ev_blk_used[L.lvl_in] = false;
    }
    if ((rc = before_dz(c, q))) return rc;
    DzArgs z;
    if ((rc = dense_dz(c, L, q, z))) return rc;
    if ((rc = after_dz(c, q))) return rc;
    if ((rc = dense_dgrad(c, L, z.dz, wsb(p) + L.wpk_bwd, 0, L.cin, L.rmw_from, true))) return rc;
    WgradArgs w;
    memset(&w, 0, sizeof(w));
    w.kind = WG_DENSE;
    w.A = act_of(p, L.lvl_in);
    w.Axp = xp_of(p, L.lvl_in);
    w.lda = lin.Ct;
    w.Ca = L.cin;
    w.transform = 1;
    w.bn = make_bn(p, L, c.params, nullptr, 1);
    w.Bm = z.dz;
    w.dz_npos = lin.dzp_npos;
    w.dz_guard = lin.dzp_guard;
    w.ldb = 16;
    w.N = 144;
    w.part_n0[0] = 0;
    w.part_np[0] = 144;
    w.npx = lin.px;
    w.H = lin.H;
    w.W = lin.W;
    w.gw = c.gs + L.p_w;
    w.Cin_w = L.cin;
    w.cs_perm = L.cs_perm;
    const int nmb = (L.cin + 127) / 128;
    int wr = wgrad_runs(p, w.npx, nmb, 1, &w.px_per_cta, 2);
    if ((rc = launch_wgrad(w, wr, nmb, 1, c.st2))) return rc;
    return after_wg(c, q);
  }
  const int tq = 2 + (c.other_count++ & 1);
  __nv_bfloat16* tmp = (__nv_bfloat16*)(wsb(p) + p->tmp_off + (size_t)(tq - 2) * c.tmp_stride);
  DzArgs z;
  fill_dz_common(c, L, z);
  z.dz = tmp;
  if (L.type == L_TD) {
    z.kind = 1;
    z.mask = (const float*)(wsb(p) + L.mask_off);
    z.amax = wsb(p) + L.amax_off;
    z.H = lin.H;
    z.W = lin.W;
    z.Hp = lout.H;
    z.Wp = lout.W;
    if ((rc = before_dz(c, tq))) return rc;
    if ((rc = launch_dz_prepare(z, sm, c.st))) return rc;
    if ((rc = after_dz(c, tq))) return rc;
    DgradArgs d;
    memset(&d, 0, sizeof(d));
    d.dense = 0;
    d.npx = lin.px;
    d.xp = xp_of(p, L.lvl_in);
    d.dz = tmp;
    d.Kdz = L.cout;
    d.wpk = wsb(p) + L.wpk_bwd;
    d.act = act_of(p, L.lvl_in);
    d.dact = dact_of(p, L.lvl_in);
    d.Ct = lin.Ct;
    d.Cin = L.cin;
    d.H = lin.H;
    d.W = lin.W;
    d.B = p->B;
    d.rmw_from = L.rmw_from;
    d.bn = make_bn(p, L, c.params, nullptr, 1);
    d.g_gamma = c.gs + L.p_gamma;
    d.g_beta = c.gs + L.p_beta;
    d.acc = acc_of(p, L.lvl_in);
    d.ntiles = (lin.px + 127) / 128;
    int grid = (int)std::max(1L, std::min(d.ntiles, (long)3 * sm));
    if ((rc = launch_dgrad(d, grid, c.st))) return rc;
    WgradArgs w;
    memset(&w, 0, sizeof(w));
    w.kind = WG_TD;
    w.A = act_of(p, L.lvl_in);
    w.Axp = xp_of(p, L.lvl_in);
    w.lda = lin.Ct;
    w.Ca = L.cin;
    w.transform = 1;
    w.bn = d.bn;
    w.Bm = tmp;
    w.ldb = L.cout;
    w.N = L.cout;
    int nparts = td_parts(L.cout, w.part_n0, w.part_np);
    w.npx = lin.px;
    w.H = lin.H;
    w.W = lin.W;
    w.gw = c.gs + L.p_w;
    w.Cin_w = L.cin;
    const int nmb = (L.cin + 127) / 128;
    int wr = wgrad_runs(p, w.npx, nmb, nparts, &w.px_per_cta);
    if ((rc = launch_wgrad(w, wr, nmb, nparts, c.st2))) return rc;
    return after_wg(c, tq);
  }
  if (L.type == L_TU) {
    z.kind = 2;
    z.H = lout.H;
    z.W = lout.W;
    z.Hp = lin.H;
    z.Wp = lin.W;
    if ((rc = before_dz(c, tq))) return rc;
    if ((rc = launch_dz_prepare(z, sm, c.st))) return rc;
    if ((rc = after_dz(c, tq))) return rc;
    TuBwdArgs t;
    memset(&t, 0, sizeof(t));
    t.dzp = tmp;
    t.npos = (long)p->B * (lin.H + 1) * (lin.W + 1);
    t.B = p->B;
    t.h = lin.H;
    t.w = lin.W;
    t.TP = tu_tile_pitch(lin.W - 1);  // TP - 1 own columns per tile
    const int TR = 128 / t.TP;
    t.ncolt = (lin.W + t.TP - 2) / (t.TP - 1);
    t.ntiles = (long)((p->B * (lin.H + 1) + TR - 1) / TR) * t.ncolt;
    t.wpk = wsb(p) + L.wpk_bwd;
    t.dact = dact_of(p, L.lvl_in);
    t.npx_in = lin.px;
    t.in_off = L.in_off;
    t.x = act_of(p, L.lvl_in);
    t.ldx = lin.Ct;
    t.x_off = L.in_off;
    t.gw = c.gs + L.p_w;
    if ((rc = launch_tu_dgrad(t, (int)std::max(1L, std::min(t.ntiles, (long)sm)), c.st))) return rc;
    if ((rc = launch_tu_wgrad(t, (int)std::max(1L, std::min(t.ntiles, (long)(sm / 3))), c.st2))) return rc;
    return after_wg(c, tq);
  }
  // L_FIRST: weight gradient only (the image needs no gradient)
  z.kind = 0;
  if ((rc = before_dz(c, tq))) return rc;
  if ((rc = launch_dz_prepare(z, sm, c.st))) return rc;
  WgradArgs w;
  memset(&w, 0, sizeof(w));
  w.kind = 4;  // WG_FIRST
  w.A = (const __nv_bfloat16*)(wsb(p) + p->xcol_off);
  w.lda = 32;
  w.Ca = 27;
  w.transform = 0;
  w.Bm = tmp;
  w.ldb = kFirst;
  w.N = kFirst;
  w.part_n0[0] = 0;
  w.part_np[0] = kFirst;
  w.npx = lout.px;
  w.H = lout.H;
  w.W = lout.W;
  w.gw = c.gs + L.p_w;
  int wr = wgrad_runs(p, w.npx, 1, 1, &w.px_per_cta);
  return launch_wgrad(w, wr, 1, 1, c.st);
}

bool block_fusable(const Plan* p, const Block& b) {
  return p->opt_bwd_fused && p->opt_xp && p->lv[b.lvl].W + 2 <= 165;  // (the chain launches use the un-stripped per-layer kernel)
}

// a whole DenseBlock: chain over the block's own new channels, one fused kernel for the block input, one batched
// weight-gradient launch (side stream)
int backward_block(BwdCtx& c, int bi) {
  Plan* p = c.p;
  const Block& b = p->blocks[bi];
  const Level& lv = p->lv[b.lvl];
  const int sm = p->sm_count;
  int rc;
  if ((rc = before_level_planes(c, b.lvl))) return rc;
  const __nv_bfloat16* dzbuf[kBlockLayers];
  for (int j = kBlockLayers - 1; j >= 0; --j) {
    const Layer& L = p->layers[b.first + j];
    DzArgs z;
    if ((rc = dense_dz(c, L, j, z))) return rc;
    dzbuf[j] = z.dz;
    if (j > 0 && (rc = dense_dgrad(c, L, z.dz, wsb(p) + L.wpk_chain, b.C0, kGrowth * j, 0, false))) return rc;
  }
  // ---- weight gradients + BatchNorm parameter gradients of the five layers (side stream; off the critical path)
  if (c.overlap) {
    RL_CUDA(cudaEventRecord((cudaEvent_t)p->ev_dz[0], c.st));
    RL_CUDA(cudaStreamWaitEvent(c.st2, (cudaEvent_t)p->ev_dz[0], 0));
  }
  {
    WgDenseArgs w;
    memset(&w, 0, sizeof(w));
    w.nl = kBlockLayers;
    int groups_total = 0, groups_max = 0;
    for (int j = 0; j < kBlockLayers; ++j) {
      const Layer& L = p->layers[b.first + j];
      WgDenseLayer& wl = w.L[j];
      wl.wpk_bwd = wsb(p) + L.wpk_bwd;
      wl.gamma = c.params + L.p_gamma;
      wl.beta = c.params + L.p_beta;
      wl.g_gamma = c.gs + L.p_gamma;
      wl.g_beta = c.gs + L.p_beta;
      wl.gw = c.gs + L.p_w;
      wl.dz = dzbuf[j];
      wl.Cin = L.cin;
      wl.cs_perm = L.cs_perm;
      const int g = ((L.cin + 63) / 64 + 1) / 2;
      groups_total += g;
      groups_max = std::max(groups_max, g);
    }
    w.xp = xp_of(p, b.lvl);
    w.npx = lv.px;
    w.H = lv.H;
    w.W = lv.W;
    w.dz_npos = lv.dzp_npos;
    w.dz_guard = lv.dzp_guard;
    w.stats = stats_of(p, b.lvl);
    w.Ct = lv.Ct;
    w.inv_n = 1.0 / (double)lv.px;
    w.acc = acc_of(p, b.lvl);
    w.C0 = b.C0;
    w.B = p->B;
    int nruns;
    w.pos_per_cta = wgrad_px_per_cta((long)p->B * (lv.H + 1) * (lv.W + 2), std::max(1, (2 * sm) / std::max(1, groups_total)), &nruns);
    const void* tm5[kBlockLayers];
    for (int j = 0; j < kBlockLayers; ++j) tm5[j] = p->tmap_dz_im2col[b.lvl][j];
    if ((rc = launch_wgrad_dense2(w, tm5, nruns, groups_max, c.st2))) return rc;
  }
  if (c.overlap) {
    RL_CUDA(cudaEventRecord((cudaEvent_t)p->ev_blk[b.lvl], c.st2));
    p->ev_blk_used[b.lvl] = true;
  }
  // ---- block-input gradient of all five layers
  DgradBulkArgs g;
  memset(&g, 0, sizeof(g));
  g.nl = kBlockLayers;
  for (int j = 0; j < kBlockLayers; ++j) {
    const Layer& L = p->layers[b.first + j];
    g.dz[j] = dzbuf[j];
    g.wpk[j] = wsb(p) + L.wpk_bulk;
    g.gamma[j] = c.params + L.p_gamma;
    g.beta[j] = c.params + L.p_beta;
  }
  g.stats = stats_of(p, b.lvl);
  g.inv_n = 1.0 / (double)lv.px;
  g.Ct = lv.Ct;
  g.cs_perm = p->layers[b.first].cs_perm;
  g.dz_npos = lv.dzp_npos;
  g.dz_guard = lv.dzp_guard;
  g.xp = xp_of(p, b.lvl);
  g.dact = dact_of(p, b.lvl);
  g.npx = lv.px;
  g.C0 = b.C0;
  g.cbs = b.cbs;
  g.H = lv.H;
  g.W = lv.W;
  g.B = p->B;
  g.rmw = b.rmw;
  g.TP = bulk_tile_pitch(lv.W);
  const int TR = 128 / g.TP;
  g.ncolt = (lv.W + g.TP - 3) / (g.TP - 2);
  g.ntiles = (long)((p->B * (lv.H + 1) + TR - 1) / TR) * g.ncolt;
  long nruns = std::max(1L, std::min(g.ntiles, (long)(2 * sm) / b.ncb));
  g.tiles_per_cta = (int)((g.ntiles + nruns - 1) / nruns);
  nruns = (g.ntiles + g.tiles_per_cta - 1) / g.tiles_per_cta;
  const void* tmaps[kBlockLayers];
  for (int j = 0; j < kBlockLayers; ++j) tmaps[j] = p->tmap_dz[b.lvl][j];
  if ((rc = launch_dgrad_bulk(g, tmaps, (int)nruns, b.ncb, c.st))) return rc;
  // the block-input channels' mean terms come from the weight-gradient launch: whoever produced those channels runs next
  if (c.overlap) RL_CUDA(cudaStreamWaitEvent(c.st, (cudaEvent_t)p->ev_blk[b.lvl], 0));
  return RL_OK;
}

int backward_range(Plan* p, const float* params, int layer_hi, int layer_lo, cudaStream_t st) {
  const int n = (int)p->layers.size();
  BwdCtx c;
  c.p = p;
  c.params = params;
  c.gs = gstage_of(p);
  c.st = st;
  c.overlap = p->opt_overlap;
  c.st2 = c.overlap ? (cudaStream_t)p->side_stream : st;
  c.dense_count = c.other_count = 0;
  c.tmp_stride = (p->tmp_bytes + 1023) / 1024 * 1024;
  for (int i = 0; i < 4; ++i) p->ev_wg_used[i] = false;
  for (int i = 0; i < kLevels; ++i) p->ev_blk_used[i] = false;
  int rc;
  int li = layer_hi;
  if (li == n) {
    if ((rc = head_wgrad(c))) return rc;
    --li;
  }
  while (li >= layer_lo) {
    const Layer& L = p->layers[li];
    if (L.type == L_DENSE && L.jblk == kBlockLayers - 1 && li - (kBlockLayers - 1) >= layer_lo && block_fusable(p, p->blocks[L.blk])) {
      if ((rc = backward_block(c, L.blk))) return rc;
      li -= kBlockLayers;
    } else {
      if ((rc = backward_layer(c, li))) return rc;
      --li;
    }
  }
  if (c.overlap) {  // join: every weight gradient is complete before the caller's stream continues
    for (int q = 0; q < 4; ++q)
      if (p->ev_wg_used[q]) RL_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)p->ev_wg[q], 0));
    for (int l = 0; l < kLevels; ++l)
      if (p->ev_blk_used[l]) RL_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)p->ev_blk[l], 0));
  }
  return RL_OK;
}

}  // namespace

extern "C" int rl_backward_range(rl_plan* h, const float* params, int layer_hi, int layer_lo, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_backward: plan not bound");
  Plan* p = &h->p;
  if (!p->training) return rl_fail(RL_ERR_STATE, "rl_backward: plan was created without training arenas");
  if (!params) return rl_fail(RL_ERR_ARG, "rl_backward: null argument");
  const int n = (int)p->layers.size();
  if (layer_lo < 0 || layer_hi > n || layer_lo > layer_hi) return rl_fail(RL_ERR_ARG, "rl_backward_range: bad layer range");
  return backward_range(p, params, layer_hi, layer_lo, (cudaStream_t)stream);
}

extern "C" int rl_backward_begin(rl_plan* h, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_backward: plan not bound");
  Plan* p = &h->p;
  if (!p->training) return rl_fail(RL_ERR_STATE, "rl_backward: plan was created without training arenas");
  RL_CUDA(cudaMemsetAsync(wsb(p) + p->acc_off_all, 0, p->acc_bytes, (cudaStream_t)stream));
  return RL_OK;
}

extern "C" int rl_grad_accum(rl_plan* h, float* grads, long lo, long hi, float grad_scale, const float* grad_scale_dev, void* stream) {
  if (!h || !h->p.ws) return rl_fail(RL_ERR_STATE, "rl_grad_accum: plan not bound");
  Plan* p = &h->p;
  if (!p->training || !grads) return rl_fail(RL_ERR_STATE, "rl_grad_accum: needs a training plan and a gradient arena");
  if (lo < 0 || hi > p->n_params || lo >= hi) return rl_fail(RL_ERR_ARG, "rl_grad_accum: bad range [%ld, %ld)", lo, hi);
  return launch_grad_accum(grads + lo, gstage_of(p) + lo, hi - lo, grad_scale, grad_scale_dev, (cudaStream_t)stream);
}

extern "C" int rl_backward(rl_plan* h, const float* params, float* grads, float grad_scale, const float* grad_scale_dev,
                           void* stream) {
  if (!params || !grads) return rl_fail(RL_ERR_ARG, "rl_backward: null argument");
  int rc = rl_backward_begin(h, stream);
  if (rc) return rc;
  Plan* p = &h->p;
  if ((rc = backward_range(p, params, (int)p->layers.size(), 0, (cudaStream_t)stream))) return rc;
  return rl_grad_accum(h, grads, 0, p->n_params, grad_scale, grad_scale_dev, stream);
}

extern "C" int rl_plan_set_option(rl_plan* h, const char* name, int value) {
  if (!h || !name) return rl_fail(RL_ERR_ARG, "rl_plan_set_option: null argument");
  Plan& p = h->p;
  if (strcmp(name, "overlap") == 0) p.opt_overlap = value ? 1 : 0;
  else if (strcmp(name, "xp") == 0) p.opt_xp = value ? 1 : 0;  // takes effect with the next rl_forward (the mirror is written by the forward pass)
  else if (strcmp(name, "fwd_tma") == 0) p.opt_fwd_tma = value ? 1 : 0;
  else if (strcmp(name, "bwd_fused") == 0) p.opt_bwd_fused = value ? 1 : 0;
  else return rl_fail(RL_ERR_ARG, "rl_plan_set_option: unknown option '%s'", name);
  return RL_OK;
}
