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
    """State lives in flat arenas parallel to the parameter arena; ``self.state[p]`` holds per-parameter VIEWS of them (and
    a ``step`` tensor) under torch's own key names, so ``state_dict()`` / ``load_state_dict()`` round-trip through
    Lightning checkpoints and are interchangeable with the stock torch optimizer of the same kind."""
    STATE_KEYS = ()

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self.rt = _runtime_of([p for g in self.param_groups for p in g["params"]])
        self._ranges = None
        self._arena = None
        self._flat = {}
        self._steps = 0
        self.grad_scale = 1.0   # multiplies the gradients inside the kernel (1/world_size after a summed all-reduce)

    def _all_params(self):
        return [p for g in self.param_groups for p in g["params"]]

    def _publish_state(self):
        """(re)create the per-parameter views of the flat state arenas in self.state"""
        base = self.rt.params.data_ptr()
        for p in self._all_params():
            off = (p.data_ptr() - base) // 4
            st = self.state[p]
            st["step"] = torch.tensor(float(self._steps))
            for k in self.STATE_KEYS:
                st[k] = self._flat[k][off:off + p.numel()].view(p.shape)

    def _ensure(self):
        self.rt.flatten()
        self.rt.reattach_grads()
        if self._ranges is None or self._arena is not self.rt.params:
            old = self._flat if self._arena is not None else {}
            self._arena = self.rt.params
            self._ranges = [_group_range(self.rt, g) for g in self.param_groups]
            self._flat = {}
            for k in self.STATE_KEYS:
                t = torch.zeros_like(self.rt.params)
                if k in old and old[k].numel() == t.numel():   # the arena moved (.to(), .cuda()): the state moves with it
                    t.copy_(old[k])
                self._flat[k] = t
            self._publish_state()

    def zero_grad(self, set_to_none=True):
        # gradients live in the flat arena and the kernels accumulate into it: one memset, views stay attached
        self.rt.flatten()
        self.rt.reattach_grads()
        self.rt.grads.zero_()

    def _after_step(self):
        if not torch.cuda.is_current_stream_capturing():
            for p in self._all_params():
                self.state[p]["step"] = torch.tensor(float(self._steps))
        self.rt.mark_weights_changed()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)          # fills self.state[p] with loaded copies (torch's own machinery)
        loaded = {p: dict(self.state[p]) for p in self._all_params() if p in self.state}
        self._ranges = None
        self._arena = None
        self._ensure()                               # fresh flat arenas + views
        base = self.rt.params.data_ptr()
        steps = 0
        for p, st in loaded.items():
            off = (p.data_ptr() - base) // 4
            for k in self.STATE_KEYS:
                if k in st and st[k] is not None:
                    self._flat[k][off:off + p.numel()].copy_(st[k].reshape(-1).to(self._flat[k].device, torch.float32))
            if "step" in st:
                steps = max(steps, int(float(st["step"])))
        self._steps = steps
        self._publish_state()


class FusedAdamW(_FlatOptimizer):
    STATE_KEYS = ("exp_avg", "exp_avg_sq")

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @property
    def exp_avg(self):
        return self._flat["exp_avg"]

    @property
    def exp_avg_sq(self):
        return self._flat["exp_avg_sq"]

    # ---- CUDA-graph mode: lr / betas / eps / weight_decay / step live in device memory, so a captured step stays valid
    # while the scheduler changes lr and the step counter advances (graphs.GraphedTrainStep)
    def enable_device_hyper(self):
        self._ensure()
        dev = self.rt.params.device
        self._hyper_host = [self._hyper_values(g) for g in self.param_groups]
        self._hyper = [torch.tensor(h + [float(self._steps)], dtype=torch.float64, device=dev) for h in self._hyper_host]

    @staticmethod
    def _hyper_values(g):
        return [float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"])]

    def sync_hyper(self):
        """push changed hyper-parameters (learning-rate schedule) to the device copies; call before replaying a graph"""
        for i, g in enumerate(self.param_groups):
            h = self._hyper_values(g)
            if h != self._hyper_host[i]:
                self._hyper[i][:5].copy_(torch.tensor(h, dtype=torch.float64))
                self._hyper_host[i] = h

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._ensure()
        self._steps += 1
        if getattr(self, "_hyper", None) is not None:
            for h, (lo, hi) in zip(self._hyper, self._ranges):
                h[5:6] += 1.0                                   # device-side step counter (captured with the step)
                engine.adamw_step_dev(self.rt.params[lo:hi], self.rt.grads[lo:hi], self.exp_avg[lo:hi],
                                      self.exp_avg_sq[lo:hi], h, self.grad_scale)
        else:
            for g, (lo, hi) in zip(self.param_groups, self._ranges):
                engine.adamw_step(self.rt.params[lo:hi], self.rt.grads[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi],
                                  g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self._steps, self.grad_scale)
        self._after_step()
        return loss


class FusedSGD(_FlatOptimizer):
    STATE_KEYS = ("momentum_buffer",)

    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0, nesterov=True):
        if not nesterov or momentum <= 0:
            raise NotImplementedError("FusedSGD implements the reference's configuration: momentum with nesterov=True")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=True))

    @property
    def momentum_buf(self):
        return self._flat["momentum_buffer"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._ensure()
        first = self._steps == 0
        self._steps += 1
        for g, (lo, hi) in zip(self.param_groups, self._ranges):
            engine.sgd_nesterov_step(self.rt.params[lo:hi], self.rt.grads[lo:hi], self.momentum_buf[lo:hi], g["lr"],
                                     g["momentum"], g["weight_decay"], first, self.grad_scale)
        self._after_step()
        return loss
