"""Whole-step CUDA graph for SimpleTrainModule.

Every launch of librightlane is stream-asynchronous with no host synchronisation (include/rightlane.h), the Dropout2d
step counter and the AdamW hyper-parameters / step counter live in device memory, and the weight-gradient side stream
forks from and re-joins the caller's stream through events, so

    zero_grad -> training_step (forward + fused head + loss) -> loss.backward() -> optimizer.step()

of the reference's fit loop (trainingModules/SimpleTrain.py:11-30 + the Lightning loop around it) is capturable as ONE
graph and replayed with one host call per step: ~250 kernel launches and the Python between them disappear from the
critical path.
"""
import torch


class GraphedTrainStep:
    """Single GPU: ONE graph per step.  Data parallel (the runtime has a ddp.GradSync attached): the step is cut where the
    backward pass hands a finished parameter range to NCCL — graph A (zero_grad, forward, head, backward of the up path),
    graph B (bottleneck + deep down path), graph C (shallow down path + first conv), graph D (optimizer) — and the
    all-reduce of each range is launched between the replays, so it still runs under the rest of backward; NCCL itself
    stays outside the graphs."""

    def __init__(self, module, optimizer, x_example, y_example, warmup=3, include_optimizer=True):
        self.module, self.optimizer, self.include_optimizer = module, optimizer, include_optimizer
        dev = x_example.device
        rt = module.runtime()
        self.rt = rt
        self.sync = rt.grad_sync if (rt.grad_sync is not None and rt.grad_sync.world > 1) else None
        self.x = torch.empty_like(x_example)
        self.y = torch.empty_like(y_example.contiguous())
        self.x.copy_(x_example)
        self.y.copy_(y_example.contiguous())
        warmup = max(1, warmup)     # the plan is created and bound (a synchronising call) by the first eager step
        if include_optimizer:
            optimizer.enable_device_hyper()
        # warm-up on a side stream (allocates the plan, the optimizer state and the workspaces before the capture)
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for i in range(warmup):
                self._eager(i)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        from . import _lib
        L = _lib.lib()
        L.rl_launch_count.restype = __import__("ctypes").c_long
        n0 = L.rl_launch_count(-1)
        self.graphs, self.ranges = [], []
        pool = None
        for piece in self._pieces():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                rng = piece()
            pool = g.pool()
            self.graphs.append(g)
            self.ranges.append(rng)
        self.captured_launches = L.rl_launch_count(-1) - n0   # librightlane kernels per step
        self.replays = 0

    # ---- the step as a list of capturable pieces; each returns the parameter ranges to all-reduce after it (or None)
    def _pieces(self):
        rt = self.rt
        state = {}

        def forward_and_head():
            self.optimizer.zero_grad()
            tok = self.module.featureExtractor(self.x)                   # forward (training mode: batch statistics, dropout)
            if not tok.engine.training:
                raise RuntimeError("GraphedTrainStep needs the module in training mode with gradients enabled")
            rt.run_head(tok, y=self.y, weighted=True, want_probs=False, backward=1)   # weighted CE on probabilities + gradient
            loss, cm = rt.metrics_from_scalars(tok.engine)
            self.loss, self.confusion = loss.clone(), cm.clone()
            state["tok"] = tok
            return tok

        if self.sync is None:
            def whole():
                tok = forward_and_head()
                rt.run_backward(tok, y=self.y, weighted=True, head_done=True)
                if self.include_optimizer:
                    self.optimizer.step()
                return None
            return [whole]
        pieces = []
        segs = None

        def first():
            tok = forward_and_head()
            eng = tok.engine
            rt.reattach_grads()
            eng.backward_begin()
            hi, lo, ranges = eng.backward_segments()[0]
            eng.backward_range(rt.params, hi, lo)
            for a, b in ranges:
                eng.grad_accum(rt.grads, a, b)
            return ranges
        pieces.append(first)
        for k in (1, 2):
            def seg(k=k):
                eng = state["tok"].engine
                hi, lo, ranges = eng.backward_segments()[k]
                eng.backward_range(rt.params, hi, lo)
                for a, b in ranges:
                    eng.grad_accum(rt.grads, a, b)
                return ranges
            pieces.append(seg)
        if self.include_optimizer:
            def opt_step():
                self.optimizer.grad_scale = 1.0 / self.sync.world
                self.optimizer.step()
                return None
            pieces.append(opt_step)
        return pieces

    def _eager(self, i):
        pieces = self._pieces()
        for k, piece in enumerate(pieces):
            if self.sync is not None and self.include_optimizer and k == len(pieces) - 1:
                self.sync.finish()                         # the optimizer needs every bucket
            rng = piece()
            if rng is not None and self.sync is not None:
                self.sync.reduce_ranges(self.rt.grads, rng)
        if self.sync is not None:
            self.sync.finish()

    def __call__(self, x, y):
        """one training step on (x, y) (device or pinned host tensors); returns the loss as a device scalar that the next
        call overwrites"""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.include_optimizer:
            self.optimizer.sync_hyper()
        for g, rng in zip(self.graphs, self.ranges):
            if rng is None and self.sync is not None:
                self.sync.finish()                         # the optimizer graph needs every bucket
            g.replay()
            if rng is not None and self.sync is not None:
                self.sync.reduce_ranges(self.rt.grads, rng)  # NCCL, ordered after the replay, concurrent with the next one
        if self.sync is not None and not self.include_optimizer:
            self.sync.finish()
        self.replays += 1
        rt = self.rt
        rt.dropout_step += 1                     # host mirrors of the device-side counters
        if self.include_optimizer:
            self.optimizer._steps += 1
            rt.mark_weights_changed()
        return self.loss
