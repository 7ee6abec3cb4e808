"""Runtime that puts the reference-shaped ``nn.Module`` tree on top of the flat arenas of ``librightlane``.

* every ``nn.Parameter`` / BN buffer of ``FCDenseNetFeatureExtractor`` + ``FCDenseNetClassifier`` keeps the
  reference's name and shape (state_dict compatible: reference trainingModules/TrainingBase.py:32-33), but its
  storage is a view into one flat fp32 arena (params, grads, BN buffers) in the order of ``spec.entries``;
* ``features = extractor(x)`` runs the CUDA forward and returns a *feature token* (the 1.4 GB normalised feature
  tensor of the reference, tiramisu.py:105, is never materialised); ``classifier(token)`` runs the fused head and,
  under autograd, registers one node whose backward runs head-backward + the whole network backward;
* ``Runtime.fused_loss`` is the SimpleTrainModule.training_step fast path (class-weighted CE computed on device).
"""
import collections
import weakref

import torch
import torch.distributed as dist

from . import _lib, spec
from .engine import Engine

DROP_P = 0.2   # nn.Dropout2d(0.2), reference layers.py:12,51


class FeatureToken:
    """Stands for the output of FCDenseNetFeatureExtractor.forward (normalised features living in the level-0 arena)."""

    def __init__(self, rt, engine, x, train, generation, sign=1.0):
        self.rt, self.engine, self.x, self.train, self.generation, self.sign = rt, engine, x, train, generation, sign

    @property
    def shape(self):
        return torch.Size((self.x.shape[0], spec.FEATURES, self.x.shape[2], self.x.shape[3]))

    def reversed(self):
        return FeatureToken(self.rt, self.engine, self.x, self.train, self.generation, -self.sign)

    def materialize(self):
        """fp32 [B, 288, H, W] normalised features in the reference's channel order (debug / interop, not a hot path)"""
        self.rt._check_generation(self)
        a = self.engine.level_view(0).float()                      # [B,H,W,288] physical order: skip128 | up80 | new80
        a = torch.cat([a[..., 128:208], a[..., :128], a[..., 208:]], -1)
        return torch.nn.functional.normalize(a.permute(0, 3, 1, 2), dim=1)


class Runtime:
    def __init__(self, extractor):
        self.extractor_ref = weakref.ref(extractor)
        self.classifier_ref = None
        self.num_cls = 4
        self.params = self.grads = self.bnbuf = None
        self.engines = collections.OrderedDict()   # LRU: train / val / tail-batch shapes of a fit loop stay alive
        self.max_engines = 4
        self.generation = 0
        self.weights_version = 0           # bumped whenever parameter values may have changed
        self.dropout_p = DROP_P
        self.dropout_seed = None           # None: derived from torch.initial_seed() and the distributed rank at first use
        self.dropout_step = 0              # one monotonically increasing counter for all plans of this network
        self._anchor = None
        self._counts = []
        self.dropout_mask_hook = None      # callable(engine) run after the Dropout2d keep-scales of a step have been drawn
        self.grad_sync = None              # ddp.GradSync: all-reduce buckets fired from inside the backward pass
        self._drop_step_dev = None         # int64 device scalar: dropout step counter (device side: graph-capturable)

    # ------------------------------------------------------------------ flat arenas
    def attach_classifier(self, classifier):
        if self.classifier_ref is None or self.classifier_ref() is not classifier:
            self.classifier_ref = weakref.ref(classifier)
            if classifier.n_classes != self.num_cls:
                self.num_cls = classifier.n_classes
                self.engines.clear()
            self.params = None  # force re-flatten with the classifier included

    def _modules(self):
        ex = self.extractor_ref()
        cl = self.classifier_ref() if self.classifier_ref is not None else None
        return ex, cl

    def _named_tensors(self):
        ex, cl = self._modules()
        out = {}
        for n, p in ex.named_parameters():
            out["featureExtractor." + n] = p
        for n, b in ex.named_buffers():
            out["featureExtractor." + n] = b
        if cl is not None:
            for n, p in cl.named_parameters():
                out["classifier." + n] = p
        return out

    def _aliased(self):
        if self.params is None:
            return False
        ex, cl = self._modules()
        first = ex.firstconv.weight
        ok = first.data_ptr() == self.params.data_ptr() and first.device == self.params.device
        if cl is not None:
            off = self._layout[0]["classifier.finalConv.bias"][0]
            ok = ok and cl.finalConv.bias.data_ptr() == self.params.data_ptr() + 4 * off
        rm = ex.denseBlocksUp[4].layers[4].norm.running_var
        off = self._layout[1]["featureExtractor.denseBlocksUp.4.layers.4.norm.running_var"][0]
        return ok and rm.data_ptr() == self.bnbuf.data_ptr() + 4 * off

    def flatten(self):
        """(re)build the flat arenas from the modules' current tensors and re-point the modules at views of them"""
        if self._aliased():
            return
        ex, cl = self._modules()
        dev = ex.firstconv.weight.device
        if dev.type != "cuda":
            raise _lib.RightLaneError("the CUDA path needs the module on a CUDA device (no CPU fallback): call .cuda()")
        pl, bl, counts, npar, nbn = spec.layout(self.num_cls)
        self._layout = (pl, bl)
        tensors = self._named_tensors()
        params = torch.zeros(npar, dtype=torch.float32, device=dev)
        grads = torch.zeros(npar, dtype=torch.float32, device=dev)
        bnbuf = torch.zeros(nbn, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for name, (off, shape) in pl.items():
                t = tensors.get(name)
                if t is None:
                    continue  # classifier not attached yet
                n = t.numel()
                params[off:off + n].copy_(t.detach().reshape(-1))
                if t.grad is not None:
                    grads[off:off + n].copy_(t.grad.reshape(-1))
                t.data = params[off:off + n].view(shape)
                t.grad = grads[off:off + n].view(shape)
            for name, (off, shape) in bl.items():
                t = tensors[name]
                n = t.numel()
                bnbuf[off:off + n].copy_(t.reshape(-1))
                t.data = bnbuf[off:off + n].view(shape)
        self._counts = [tensors[n] for n in counts]
        self.params, self.grads, self.bnbuf = params, grads, bnbuf
        self.n_extractor = pl["classifier.finalConv.weight"][0]
        self.weights_version += 1
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        for t in tensors.values():
            if isinstance(t, torch.nn.Parameter):
                t._rl_runtime = weakref.ref(self)

    def reattach_grads(self):
        """optimizers that call zero_grad(set_to_none=True) drop the arena views; restore them"""
        tensors = self._named_tensors()
        for name, (off, shape) in self._layout[0].items():
            t = tensors.get(name)
            if t is not None and (t.grad is None or t.grad.data_ptr() != self.grads.data_ptr() + 4 * off):
                if t.grad is not None:
                    self.grads[off:off + t.numel()].add_(t.grad.reshape(-1))
                t.grad = self.grads[off:off + t.numel()].view(shape)

    def mark_weights_changed(self):
        self.weights_version += 1

    # ------------------------------------------------------------------ engines
    def engine_for(self, x, need_grad, train=False):
        B, _, H, W = x.shape
        if x.device != self.params.device:
            raise _lib.RightLaneError(f"input on {x.device} but the network lives on {self.params.device}")
        key = (B, H, W, self.num_cls, bool(need_grad), x.device.index)
        eng = self.engines.get(key)
        if eng is None:
            while len(self.engines) >= self.max_engines:   # a plan owns its activation arenas: evict the least recently used
                self.engines.popitem(last=False)[1].close()
            eng = Engine(B, H, W, self.num_cls, training=need_grad, device=x.device).bind()
            eng._packed_version = -1
            self.engines[key] = eng
        else:
            self.engines.move_to_end(key)
        # Training-mode forwards always repack the bf16 tensor-core operands from the fp32 arena (one 18 us launch), so
        # ANY in-place parameter update (stock torch optimizers, nn.init, EMA swaps) is picked up; eval-mode forwards
        # repack when a tracked change happened (fused optimizers, load_state_dict, mark_weights_changed()).
        if train or eng._packed_version != self.weights_version:
            eng.pack_weights(self.params)
            eng._packed_version = self.weights_version
        return eng

    def _dropout_seed(self):
        if self.dropout_seed is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
            self.dropout_seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03 + 0x5EED) & (2 ** 64 - 1)
        return self.dropout_seed

    def _check_generation(self, tok):
        if tok.generation != self.generation:
            raise _lib.RightLaneError("feature token is stale: another forward pass has overwritten the activation arenas")

    # ------------------------------------------------------------------ forward pieces
    def run_extractor(self, x):
        ex, _ = self._modules()
        self.flatten()
        x = x.contiguous().float()
        train = ex.training
        need_grad = train and torch.is_grad_enabled()
        with torch.cuda.device(x.device):
            return self._run_extractor(x, train, need_grad)

    def _run_extractor(self, x, train, need_grad):
        eng = self.engine_for(x, need_grad, train)
        if train:
            # the step counter lives on the device (one in-place add per step), so a captured CUDA graph of the step draws
            # fresh Dropout2d masks at every replay
            if self._drop_step_dev is None or self._drop_step_dev.device != x.device:
                self._drop_step_dev = torch.full((1,), self.dropout_step, dtype=torch.int64, device=x.device)
            eng.set_dropout(self.dropout_p, self._dropout_seed(), 0, step_dev=self._drop_step_dev)
            self._drop_step_dev += 1
            self.dropout_step += 1
            if self.dropout_mask_hook is not None:     # parity tests inject the keep masks torch drew for the oracle
                self.dropout_mask_hook(eng)
        eng.forward(self.params, self.bnbuf, x, train)
        if train:
            eng.bn_running_update(self.bnbuf, 0.1)
            if self._counts:
                torch._foreach_add_(self._counts, 1)
        self.generation += 1
        return FeatureToken(self, eng, x, train, self.generation)

    def run_head(self, tok, y=None, weighted=False, want_probs=True, want_mask=False, backward=0):
        """fused head; backward=1 (training plans): the same launch also writes d(loss)/d(features) for grad_output 1"""
        self._check_generation(tok)
        eng, x = tok.engine, tok.x
        B, _, H, W = x.shape
        probs = torch.empty(B, self.num_cls, H, W, device=x.device, dtype=torch.float32) if want_probs else None
        mask = torch.empty(B, H, W, device=x.device, dtype=torch.uint8) if want_mask else None
        with torch.cuda.device(x.device):
            eng.head(self.params, y=y, weighted=weighted, probs=probs, mask=mask, backward=backward, feat_grad_sign=tok.sign)
        return probs, mask

    def run_backward(self, tok, y=None, weighted=True, dprobs=None, grad_scale=1.0, grad_scale_dev=None, head_done=False):
        self._check_generation(tok)
        if not tok.engine.training:
            raise _lib.RightLaneError("backward requested for a forward pass that ran without gradient arenas")
        self.reattach_grads()
        eng = tok.engine
        with torch.cuda.device(tok.x.device):
            if dprobs is not None:
                eng.head(self.params, y=y, weighted=weighted, backward=2, dprobs=dprobs.contiguous().float(),
                         feat_grad_sign=tok.sign)
            elif not head_done:
                eng.head(self.params, y=y, weighted=weighted, backward=1, feat_grad_sign=tok.sign)
            if self.grad_sync is None or self.grad_sync.world == 1:
                eng.backward(self.params, self.grads, grad_scale, grad_scale_dev)
            else:   # data parallel: each segment's parameter ranges are final when it returns -> their all-reduce starts now
                eng.backward_begin()
                for hi, lo, ranges in eng.backward_segments():
                    eng.backward_range(self.params, hi, lo)
                    for a, b in ranges:
                        eng.grad_accum(self.grads, a, b, grad_scale, grad_scale_dev)
                    self.grad_sync.reduce_ranges(self.grads, ranges)

    def check_labels(self, eng):
        """labels outside [0, num_cls) are excluded from loss, gradient and metrics by the kernels; the reference's
        cross_entropy raises on them.  Reads the plan's counter (a host sync: call it at epoch ends, not per step)."""
        bad = int(eng.scalars()[3].item())
        if bad:
            raise _lib.RightLaneError(f"{bad} label values outside [0, {self.num_cls}) in the last batch")

    def metrics_from_scalars(self, eng):
        """(loss, confusion matrix int64 [C,C]) as device tensors from the plan's scalar block — no host sync"""
        hist, cm, loss, bad = eng.scalars()
        C = self.num_cls
        return (loss[0] / loss[1]).float(), cm[:C * C].view(C, C).long()


# ---------------------------------------------------------------------- autograd glue
class _ProbsFn(torch.autograd.Function):
    """probs = classifier(token); backward: d(probs) -> every parameter gradient (written into the flat arena)"""

    @staticmethod
    def forward(ctx, anchor, tok):
        ctx.tok = tok
        probs, _ = tok.rt.run_head(tok)
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        ctx.tok.rt.run_backward(ctx.tok, dprobs=dprobs)
        return None, None


class _FusedLossFn(torch.autograd.Function):
    """SimpleTrainModule.training_step (reference SimpleTrain.py:11-25): class-weighted CE on the probabilities.
    On a training plan the head runs ONCE: the forward launch already writes d(loss)/d(features) for grad_output 1, and
    backward() scales the staged gradients by the real grad_output (a device scalar: no host synchronisation)."""

    @staticmethod
    def forward(ctx, anchor, tok, y, weighted):
        ctx.tok, ctx.y, ctx.weighted = tok, y, weighted
        rt = tok.rt
        ctx.head_done = bool(tok.engine.training and tok.train)
        rt.run_head(tok, y=y, weighted=weighted, want_probs=False, backward=1 if ctx.head_done else 0)
        loss, cm = rt.metrics_from_scalars(tok.engine)
        ctx.mark_non_differentiable(cm)
        return loss.clone(), cm.clone()

    @staticmethod
    def backward(ctx, dloss, _dcm):
        gs = dloss.detach().reshape(1).float().contiguous()
        ctx.tok.rt.run_backward(ctx.tok, y=ctx.y, weighted=ctx.weighted, grad_scale=1.0, grad_scale_dev=gs,
                                head_done=ctx.head_done)
        return None, None, None, None
