"""Runtime that puts the reference-shaped ``nn.Module`` tree on top of the flat arenas of ``librightlane``.

* every ``nn.Parameter`` / BN buffer of ``FCDenseNetFeatureExtractor`` + ``FCDenseNetClassifier`` keeps the
  reference's name and shape (state_dict compatible: reference trainingModules/TrainingBase.py:32-33), but its
  storage is a view into one flat fp32 arena (params, grads, BN buffers) in the order of ``spec.entries``;
* ``features = extractor(x)`` runs the CUDA forward and returns a *feature token* (the 1.4 GB normalised feature
  tensor of the reference, tiramisu.py:105, is never materialised); ``classifier(token)`` runs the fused head and,
  under autograd, registers one node whose backward runs head-backward + the whole network backward;
* ``Runtime.fused_loss`` is the SimpleTrainModule.training_step fast path (class-weighted CE computed on device).
"""
import weakref

import torch

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
        self.engines = {}
        self.generation = 0
        self.weights_version = 0           # bumped whenever parameter values may have changed
        self.dropout_p = DROP_P
        self.dropout_seed = 0x5EED
        self._anchor = None
        self._counts = []

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
    def engine_for(self, x, need_grad):
        B, _, H, W = x.shape
        key = (B, H, W, self.num_cls, bool(need_grad), x.device.index)
        eng = self.engines.get(key)
        if eng is None:
            if len(self.engines) >= 3:   # a plan owns its activation arenas: keep only a few shapes alive
                self.engines.pop(next(iter(self.engines))).close()
            eng = Engine(B, H, W, self.num_cls, training=need_grad, device=x.device).bind()
            eng._packed_version = -1
            self.engines[key] = eng
        if eng._packed_version != self.weights_version:
            eng.pack_weights(self.params)
            eng._packed_version = self.weights_version
        return eng

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
        eng = self.engine_for(x, need_grad)
        if train:
            eng.set_dropout(self.dropout_p, self.dropout_seed)
        eng.forward(self.params, self.bnbuf, x, train)
        if train:
            eng.bn_running_update(self.bnbuf, 0.1)
            if self._counts:
                torch._foreach_add_(self._counts, 1)
        self.generation += 1
        return FeatureToken(self, eng, x, train, self.generation)

    def run_head(self, tok, y=None, weighted=False, want_probs=True, want_mask=False):
        self._check_generation(tok)
        eng, x = tok.engine, tok.x
        B, _, H, W = x.shape
        probs = torch.empty(B, self.num_cls, H, W, device=x.device, dtype=torch.float32) if want_probs else None
        mask = torch.empty(B, H, W, device=x.device, dtype=torch.uint8) if want_mask else None
        eng.head(self.params, y=y, weighted=weighted, probs=probs, mask=mask)
        return probs, mask

    def run_backward(self, tok, y=None, weighted=True, dprobs=None, grad_scale=1.0, grad_scale_dev=None):
        self._check_generation(tok)
        if not tok.engine.training:
            raise _lib.RightLaneError("backward requested for a forward pass that ran without gradient arenas")
        self.reattach_grads()
        eng = tok.engine
        if dprobs is not None:
            eng.head(self.params, y=y, weighted=weighted, backward=2, dprobs=dprobs.contiguous().float(),
                     feat_grad_sign=tok.sign, grads=self.grads)
        else:
            eng.head(self.params, y=y, weighted=weighted, backward=1, grad_scale=grad_scale, feat_grad_sign=tok.sign,
                     grads=self.grads, grad_scale_dev=grad_scale_dev)
        eng.backward(self.params, tok.x, self.grads)

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
    """SimpleTrainModule.training_step (reference SimpleTrain.py:11-25): class-weighted CE on the probabilities."""

    @staticmethod
    def forward(ctx, anchor, tok, y, weighted):
        ctx.tok, ctx.y, ctx.weighted = tok, y, weighted
        rt = tok.rt
        rt.run_head(tok, y=y, weighted=weighted, want_probs=False)
        loss, cm = rt.metrics_from_scalars(tok.engine)
        ctx.mark_non_differentiable(cm)
        return loss.clone(), cm.clone()

    @staticmethod
    def backward(ctx, dloss, _dcm):
        # grad_output of the loss (1 unless the caller scales the loss) is read by the head kernel from device
        # memory, so no host synchronisation happens between forward and backward
        gs = dloss.detach().reshape(1).float().contiguous()
        ctx.tok.rt.run_backward(ctx.tok, y=ctx.y, weighted=ctx.weighted, grad_scale=1.0, grad_scale_dev=gs)
        return None, None, None, None
