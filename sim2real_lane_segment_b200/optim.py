"""Fused optimizers over the flat parameter arena (one kernel launch per step instead of 254 per-tensor updates).

FusedAdamW   — torch.optim.AdamW semantics (reference SimpleTrain.py:28, MMETrainingModule.py:15)
FusedSGD     — torch.optim.SGD(momentum, nesterov=True, weight_decay) with per-group learning rates
               (reference MMETrainingModule.py:16-19: extractor lr/3, classifier lr)
Both are ``torch.optim.Optimizer`` subclasses, so ``CosineAnnealingLR`` (SimpleTrain.py:29) drives ``param_groups[i]['lr']``.
"""
import torch

from . import engine


def _runtime_of(params):
    rts = {}
    for p in params:
        ref = getattr(p, "_rl_runtime", None)
        rt = ref() if ref is not None else None
        if rt is None:
            raise ValueError("parameter is not backed by a librightlane runtime (move the model to CUDA and run "
                             "model.prepare() / one forward pass before constructing the optimizer)")
        rts[id(rt)] = rt
    if len(rts) != 1:
        raise ValueError("all parameters of a fused optimizer must belong to one network")
    return next(iter(rts.values()))


def _group_range(rt, group):
    """contiguous [lo, hi) float range of the flat arena covered by a param group"""
    base = rt.params.data_ptr()
    spans = sorted(((p.data_ptr() - base) // 4, (p.data_ptr() - base) // 4 + p.numel()) for p in group["params"])
    lo, hi = spans[0][0], spans[0][1]
    for a, b in spans[1:]:
        if a != hi:
            raise ValueError("a fused optimizer group must cover a contiguous range of the parameter arena")
        hi = b
    return lo, hi


class _FlatOptimizer(torch.optim.Optimizer):
    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self.rt = _runtime_of([p for g in self.param_groups for p in g["params"]])
        self._ranges = None
        self._steps = 0
        self.grad_scale = 1.0   # multiplies the gradients inside the kernel (1/world_size after a summed all-reduce)

    def _ensure(self):
        self.rt.flatten()
        self.rt.reattach_grads()
        if self._ranges is None or self._arena is not self.rt.params:
            self._arena = self.rt.params
            self._ranges = [_group_range(self.rt, g) for g in self.param_groups]
            self._alloc_state()

    def zero_grad(self, set_to_none=True):
        # gradients live in the flat arena and the kernels accumulate into it: one memset, views stay attached
        self.rt.flatten()
        self.rt.reattach_grads()
        self.rt.grads.zero_()


class FusedAdamW(_FlatOptimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _alloc_state(self):
        self.exp_avg = torch.zeros_like(self.rt.params)
        self.exp_avg_sq = torch.zeros_like(self.rt.params)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._ensure()
        self._steps += 1
        for g, (lo, hi) in zip(self.param_groups, self._ranges):
            engine.adamw_step(self.rt.params[lo:hi], self.rt.grads[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                              g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._steps, self.grad_scale)
        self.rt.mark_weights_changed()
        return loss


class FusedSGD(_FlatOptimizer):
    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0, nesterov=True):
        if not nesterov or momentum <= 0:
            raise NotImplementedError("FusedSGD implements the reference's configuration: momentum with nesterov=True")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=True))

    def _alloc_state(self):
        self.momentum_buf = torch.zeros_like(self.rt.params)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._ensure()
        first = self._steps == 0
        self._steps += 1
        for g, (lo, hi) in zip(self.param_groups, self._ranges):
            engine.sgd_nesterov_step(self.rt.params[lo:hi], self.rt.grads[lo:hi], self.momentum_buf[lo:hi], g["lr"],
                                     g["momentum"], g["weight_decay"], first, self.grad_scale)
        self.rt.mark_weights_changed()
        return loss
