"""Host-side driver of one ``rl_plan`` (include/rightlane.h): owns the workspace tensor, forwards torch device
pointers to the C ABI on torch's current stream, and exposes the scalar block (loss sums, confusion counts).

PyTorch is used here only for device memory and streams.  There is no CPU fallback: every method raises if
``librightlane.so`` is missing or the tensors are not CUDA tensors.
"""
import ctypes
from ctypes import POINTER, byref, c_float, c_int, c_long, c_ulonglong, c_void_p

import torch

from . import _lib

_SIGS_DONE = False

Q_NUM_LAYERS, Q_ACT_OFF, Q_DACT_OFF, Q_CT, Q_H, Q_W, Q_CSKIP, Q_SCAL_OFF, Q_LAYER_TYPE, Q_LAYER_MASK_OFF, \
    Q_LAYER_COUT, Q_LAYER_PARAM_W, Q_STATS_OFF, Q_GSTAGE_OFF, Q_XP_OFF, Q_ACC_OFF, Q_LAYER_LEVEL_IN, Q_LAYER_LEVEL_OUT, \
    Q_LAYER_CIN, Q_LAYER_IN_OFF, Q_LAYER_OUT_OFF, Q_LAYER_CS_PERM, Q_LAYER_AMAX_OFF, Q_LAYER_RMW_FROM, \
    Q_LAYER_PARAM_GAMMA, Q_LAYER_PARAM_B, Q_LAYER_BN_RM, Q_DL_OFF = range(28)


def _sigs(L):
    global _SIGS_DONE
    if _SIGS_DONE:
        return
    L.rl_plan_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int]
    L.rl_plan_destroy.argtypes = [c_void_p]
    L.rl_plan_destroy.restype = None
    for f in (L.rl_plan_workspace_bytes, L.rl_plan_num_params, L.rl_plan_num_bn_buffers):
        f.argtypes = [c_void_p]
        f.restype = c_long
    L.rl_plan_bind.argtypes = [c_void_p, c_void_p, c_void_p]
    L.rl_plan_query.argtypes = [c_void_p, c_int, c_int, POINTER(c_long)]
    L.rl_pack_weights.argtypes = [c_void_p, c_void_p, c_void_p]
    L.rl_set_dropout.argtypes = [c_void_p, c_float, c_ulonglong, c_ulonglong, c_void_p]
    L.rl_set_dropout_dev.argtypes = [c_void_p, c_float, c_ulonglong, c_ulonglong, c_void_p, c_void_p]
    L.rl_adamw_step_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_float, c_void_p]
    L.rl_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    L.rl_forward_range.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    L.rl_head.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_float,
                          c_void_p]
    L.rl_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]
    L.rl_backward_range.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p]
    L.rl_backward_begin.argtypes = [c_void_p, c_void_p]
    L.rl_grad_accum.argtypes = [c_void_p, c_void_p, c_long, c_long, c_float, c_void_p, c_void_p]
    L.rl_plan_set_option.argtypes = [c_void_p, ctypes.c_char_p, c_int]
    L.rl_bn_running_update.argtypes = [c_void_p, c_void_p, c_float, c_void_p]
    L.rl_adamw_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_long, ctypes.c_double, ctypes.c_double,
                                ctypes.c_double, ctypes.c_double, ctypes.c_double, c_long, c_float, c_void_p]
    L.rl_sgd_nesterov_step.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_float, c_float, c_float, c_int, c_float,
                                       c_void_p]
    _SIGS_DONE = True


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.RightLaneError("librightlane operates on CUDA tensors only (no CPU fallback)")


class Engine:
    """One FC-DenseNet67 execution plan for a fixed (B, H, W, num_cls)."""

    def __init__(self, B, H, W, num_cls=4, training=True, device=None):
        self.L = _lib.lib()
        _sigs(self.L)
        self.B, self.H, self.W, self.num_cls, self.training = B, H, W, num_cls, bool(training)
        self.device = torch.device(device if device is not None else "cuda")
        h = c_void_p()
        _lib.check(self.L.rl_plan_create(byref(h), B, H, W, num_cls, int(training)))
        self.h = h
        self.n_params = self.L.rl_plan_num_params(h)
        self.n_bnbuf = self.L.rl_plan_num_bn_buffers(h)
        self.ws_bytes = self.L.rl_plan_workspace_bytes(h)
        self.ws = None

    # ------------------------------------------------------------------ plumbing
    def bind(self):
        if self.ws is None:
            with torch.cuda.device(self.device):
                self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
                off = (-self.ws.data_ptr()) % 256
                self._ws_view = self.ws[off:off + self.ws_bytes]
                _lib.check(self.L.rl_plan_bind(self.h, c_void_p(self._ws_view.data_ptr()), _lib.cur_stream()))
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.rl_plan_destroy(self.h)
            self.h = None
        self.ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, what, index=0):
        out = c_long()
        _lib.check(self.L.rl_plan_query(self.h, what, index, byref(out)))
        return out.value

    def level_view(self, level, grad=False):
        """[B, H_l, W_l, Ct_l] bf16 view of a concat arena (physical channel order: skip | up | new); grad=True: a copy of
        the level's gradient arena in the same NHWC form."""
        off = self.query(Q_DACT_OFF if grad else Q_ACT_OFF, level)
        Ct, H, W = self.query(Q_CT, level), self.query(Q_H, level), self.query(Q_W, level)
        n = self.B * H * W * Ct
        flat = self._ws_view[off:off + 2 * n].view(torch.bfloat16)
        if grad:   # gradient arenas are planar by 16-channel group, [Ct/16][px][16]: return an NHWC copy
            return flat.view(Ct // 16, self.B, H, W, 16).permute(1, 2, 3, 0, 4).reshape(self.B, H, W, Ct)
        return flat.view(self.B, H, W, Ct)

    def level_stats(self, level):
        off, Ct = self.query(Q_STATS_OFF, level), self.query(Q_CT, level)
        return self._ws_view[off:off + 16 * Ct].view(torch.float64).view(2, Ct)

    def level_acc(self, level):
        """BatchNorm-backward mean terms of a level: float64 [2, Ct] = (Σ A-terms, Σ B-terms), before the division by N"""
        off, Ct = self.query(Q_ACC_OFF, level), self.query(Q_CT, level)
        return self._ws_view[off:off + 16 * Ct].view(torch.float64).view(2, Ct)

    def level_planar(self, level, which):
        """raw planar view [Ct/16, B, H, W, 16] bf16 of a level's gradient arena ('dact') or activation mirror ('xp')"""
        off = self.query(Q_DACT_OFF if which == "dact" else Q_XP_OFF, level)
        Ct, H, W = self.query(Q_CT, level), self.query(Q_H, level), self.query(Q_W, level)
        n = self.B * H * W * Ct
        return self._ws_view[off:off + 2 * n].view(torch.bfloat16).view(Ct // 16, self.B, H, W, 16)

    def set_level(self, level, nhwc, which="act"):
        """write a [B, H, W, Ct] tensor into a level's activation arena (and its planar mirror) or gradient arena"""
        t = nhwc.to(torch.bfloat16)
        if which == "act":
            self.level_view(level).copy_(t)
            which = "xp"
        Ct = t.shape[-1]
        self.level_planar(level, which).copy_(t.view(self.B, t.shape[1], t.shape[2], Ct // 16, 16).permute(3, 0, 1, 2, 4))

    def gstage(self):
        """float32 [num_params]: raw gradients of the current backward pass (before rl_backward's final scaling)"""
        off = self.query(Q_GSTAGE_OFF)
        return self._ws_view[off:off + 4 * self.n_params].view(torch.float32)

    def set_option(self, name, value):
        _lib.check(self.L.rl_plan_set_option(self.h, name.encode(), int(value)))

    def scalars(self):
        """device views: hist int32[8], cm int32[64], (loss_num, loss_den) float64[2], bad_label int32[1]"""
        off = self.query(Q_SCAL_OFF)
        raw = self._ws_view[off:off + 320]
        ints = raw[:288].view(torch.int32)
        return ints[:8], ints[8:72], raw[288:304].view(torch.float64), raw[304:308].view(torch.int32)

    def layer_mask_view(self, layer):
        off, cout = self.query(Q_LAYER_MASK_OFF, layer), self.query(Q_LAYER_COUT, layer)
        return self._ws_view[off:off + 4 * self.B * cout].view(torch.float32).view(self.B, cout)

    # ------------------------------------------------------------------ ops
    def pack_weights(self, params):
        _need_cuda(params)
        _lib.check(self.L.rl_pack_weights(self.h, _lib.ptr(params), _lib.cur_stream()))

    def set_dropout(self, p_drop, seed=0, step=0, step_dev=None):
        """step_dev: optional int64 device scalar added to `step` inside the kernel (graph-capturable step counter)"""
        _lib.check(self.L.rl_set_dropout_dev(self.h, float(p_drop), seed & 0xFFFFFFFFFFFFFFFF, step, _lib.ptr(step_dev),
                                             _lib.cur_stream()))

    def forward(self, params, bnbuf, x, train):
        _need_cuda(params, bnbuf, x)
        assert x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (self.B, 3, self.H, self.W)
        _lib.check(self.L.rl_forward(self.h, _lib.ptr(params), _lib.ptr(bnbuf), _lib.ptr(x), int(train), _lib.cur_stream()))

    def forward_range(self, params, bnbuf, x, train, lo, hi):
        _need_cuda(params, bnbuf, x)
        _lib.check(self.L.rl_forward_range(self.h, _lib.ptr(params), _lib.ptr(bnbuf), _lib.ptr(x), int(train), lo, hi,
                                           _lib.cur_stream()))

    def head(self, params, y=None, weighted=False, probs=None, mask=None, backward=0, dprobs=None, feat_grad_sign=1.0):
        _need_cuda(params, y, probs, mask, dprobs)
        y64 = 0
        if y is not None:
            assert y.is_contiguous() and y.dtype in (torch.uint8, torch.int64)
            y64 = int(y.dtype == torch.int64)
        _lib.check(self.L.rl_head(self.h, _lib.ptr(params), _lib.ptr(y), y64, int(weighted), _lib.ptr(probs), _lib.ptr(mask),
                                  int(backward), _lib.ptr(dprobs), float(feat_grad_sign), _lib.cur_stream()))

    def backward(self, params, grads, grad_scale=1.0, grad_scale_dev=None):
        _need_cuda(params, grads, grad_scale_dev)
        _lib.check(self.L.rl_backward(self.h, _lib.ptr(params), _lib.ptr(grads), float(grad_scale), _lib.ptr(grad_scale_dev),
                                      _lib.cur_stream()))

    def backward_range(self, params, hi, lo):
        _need_cuda(params)
        _lib.check(self.L.rl_backward_range(self.h, _lib.ptr(params), hi, lo, _lib.cur_stream()))

    def backward_begin(self):
        _lib.check(self.L.rl_backward_begin(self.h, _lib.cur_stream()))

    def grad_accum(self, grads, lo, hi, grad_scale=1.0, grad_scale_dev=None):
        _need_cuda(grads, grad_scale_dev)
        _lib.check(self.L.rl_grad_accum(self.h, _lib.ptr(grads), lo, hi, float(grad_scale), _lib.ptr(grad_scale_dev),
                                        _lib.cur_stream()))

    def backward_segments(self):
        """Three segments of the backward pass in the order they complete, each with the parameter-arena ranges whose
        gradients are final once it has run: [(layer_hi, layer_lo, [(lo, hi), ...]), ...] — up path + classifier first,
        then bottleneck + deep down path, then the shallow down path + first conv."""
        if getattr(self, "_segments", None) is None:
            n = self.query(Q_NUM_LAYERS)
            types = [self.query(Q_LAYER_TYPE, i) for i in range(n)]

            def start(i):   # first float of layer i's parameters (BatchNorm weight where the layer has one, else conv weight)
                g = self.query(Q_LAYER_PARAM_GAMMA, i)
                return g if g >= 0 else self.query(Q_LAYER_PARAM_W, i)
            tds = [i for i, t in enumerate(types) if t == 2]
            tus = [i for i, t in enumerate(types) if t == 3]
            first_tu, bott = tus[0], tds[-1] + 1
            down2 = tds[1] + 1                       # first layer of the level-2 down block
            self._segments = [
                (n, first_tu, [(start(first_tu), self.n_params)]),
                (first_tu - 1, down2, [(start(down2), start(tds[0])), (start(tds[2]), start(first_tu))]),
                (down2 - 1, 0, [(0, start(down2)), (start(tds[0]), start(tds[2]))]),
            ]
            assert bott > down2
        return self._segments

    def bn_running_update(self, bnbuf, momentum=0.1):
        _need_cuda(bnbuf)
        _lib.check(self.L.rl_bn_running_update(self.h, _lib.ptr(bnbuf), float(momentum), _lib.cur_stream()))


def adamw_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    L = _lib.lib()
    _sigs(L)
    _need_cuda(params, grads, exp_avg, exp_avg_sq)
    _lib.check(L.rl_adamw_step(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), params.numel(),
                               lr, beta1, beta2, eps, weight_decay, int(step), grad_scale, _lib.cur_stream()))


def adamw_step_dev(params, grads, exp_avg, exp_avg_sq, hyper_dev, grad_scale=1.0):
    """hyper_dev: float64 device tensor [lr, beta1, beta2, eps, weight_decay, step]"""
    L = _lib.lib()
    _sigs(L)
    _need_cuda(params, grads, exp_avg, exp_avg_sq, hyper_dev)
    assert hyper_dev.dtype == torch.float64 and hyper_dev.numel() == 6
    _lib.check(L.rl_adamw_step_dev(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), params.numel(),
                                   _lib.ptr(hyper_dev), grad_scale, _lib.cur_stream()))


def sgd_nesterov_step(params, grads, buf, lr, momentum, weight_decay, first_step, grad_scale=1.0):
    L = _lib.lib()
    _sigs(L)
    _need_cuda(params, grads, buf)
    _lib.check(L.rl_sgd_nesterov_step(_lib.ptr(params), _lib.ptr(grads), _lib.ptr(buf), params.numel(), lr, momentum,
                                      weight_decay, int(first_step), grad_scale, _lib.cur_stream()))
