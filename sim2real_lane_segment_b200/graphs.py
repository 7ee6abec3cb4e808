"""Whole-step CUDA graph for SimpleTrainModule.

Every launch of librightlane is stream-asynchronous with no host synchronisation (include/rightlane.h), the Dropout2d
step counter and the AdamW hyper-parameters / step counter live in device memory, and the weight-gradient side stream
forks from and re-joins the caller's stream through events, so

    zero_grad -> training_step (forward + fused head + loss) -> loss.backward() -> optimizer.step()

of the reference's fit loop (trainingModules/SimpleTrain.py:11-30 + the Lightning loop around it) is capturable as ONE
graph and replayed with one host call per step: ~250 kernel launches and the Python between them disappear from the
critical path.  Data-parallel runs capture forward + backward and keep the gradient all-reduce and the optimizer outside
the graph (`include_optimizer=False`).
"""
import torch


class GraphedTrainStep:
    def __init__(self, module, optimizer, x_example, y_example, warmup=3, include_optimizer=True):
        self.module, self.optimizer, self.include_optimizer = module, optimizer, include_optimizer
        dev = x_example.device
        self.x = torch.empty_like(x_example)
        self.y = torch.empty_like(y_example)
        self.x.copy_(x_example)
        self.y.copy_(y_example.contiguous())
        warmup = max(1, warmup)     # the plan is created and bound (a synchronising call) by the first eager step
        if include_optimizer:
            optimizer.enable_device_hyper()
        # warm-up on a side stream (allocates the plan, the optimizer state and autograd's buffers before the capture)
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for i in range(warmup):
                self._body(i)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        from . import _lib
        L = _lib.lib()
        L.rl_launch_count.restype = __import__("ctypes").c_long
        n0 = L.rl_launch_count(-1)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body(0)
        self.captured_launches = L.rl_launch_count(-1) - n0   # librightlane kernels per replay
        self.replays = 0

    def _body(self, i):
        """SimpleTrainModule.training_step + loss.backward() (+ optimizer.step()) through the runtime directly: the same
        launches as the autograd path (runtime._FusedLossFn), without the autograd engine's worker thread, whose stream
        bookkeeping is not capture-safe the first time a process captures"""
        rt = self.module.runtime()
        self.optimizer.zero_grad()
        tok = self.module.featureExtractor(self.x)                       # forward (training mode: batch statistics, dropout)
        if not tok.engine.training:
            raise RuntimeError("GraphedTrainStep needs the module in training mode with gradients enabled")
        rt.run_head(tok, y=self.y, weighted=True, want_probs=False, backward=1)   # class-weighted CE on probabilities + its gradient
        rt.run_backward(tok, y=self.y, weighted=True, head_done=True)
        if self.include_optimizer:
            self.optimizer.step()
        loss, cm = rt.metrics_from_scalars(tok.engine)
        self.confusion = cm.clone()
        return loss.clone()

    def __call__(self, x, y):
        """one training step on (x, y) (device or pinned host tensors); returns the loss as a device scalar that the next
        call overwrites"""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.include_optimizer:
            self.optimizer.sync_hyper()
        self.graph.replay()
        self.replays += 1
        rt = self.module.runtime()
        rt.dropout_step += 1                     # host mirrors of the device-side counters
        if self.include_optimizer:
            self.optimizer._steps += 1
            rt.mark_weights_changed()
        return self.loss
