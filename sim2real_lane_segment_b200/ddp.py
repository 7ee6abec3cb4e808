"""Data-parallel gradient exchange for the flat gradient arena.

The reference has no distributed code of its own: ``Trainer(gpus=N)`` of pytorch-lightning 1.2.1 wraps the module in
DDP, i.e. one all-reduce (mean) of the 3 461 220-element gradient vector per step, plain (un-synchronised) BatchNorm
with local batch statistics per rank.  Here the gradients already live in ONE flat fp32 tensor, so the exchange is a
handful of NCCL all-reduces over contiguous byte ranges ("buckets"), launched asynchronously and awaited right before
the optimizer step; the 1/world scaling is folded into the fused optimizer kernel (``grad_scale``).
"""
import torch
import torch.distributed as dist


def bucket_ranges(n, n_buckets):
    """split [0, n) into n_buckets contiguous ranges whose starts are multiples of 256 elements"""
    n_buckets = max(1, min(n_buckets, (n + 255) // 256))
    per = ((n + n_buckets - 1) // n_buckets + 255) // 256 * 256
    out = []
    lo = 0
    while lo < n:
        hi = min(n, lo + per)
        out.append((lo, hi))
        lo = hi
    return out


class GradSync:
    """NCCL all-reduce (SUM) of the flat gradient arena.  Attached to a runtime (`rt.grad_sync = GradSync()`), the backward
    pass fires the all-reduce of a parameter range as soon as the layers that own it are done (`reduce_ranges`, called by
    `Runtime.run_backward` after each of the three backward segments), so the exchange runs under the rest of backward;
    `finish()` is called right before the optimizer step."""
    collective = "ncclAllReduce(sum) over 3 arena buckets, issued from inside backward (up path, deep down path, shallow down path)"

    def __init__(self, group=None, n_buckets=3):
        self.group = group
        self.n_buckets = n_buckets
        self.handles = []
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def reduce_ranges(self, flat_grads, ranges):
        """async all-reduce of the given [lo, hi) float ranges; ordered after everything enqueued on the current stream"""
        if self.world == 1:
            return
        for lo, hi in ranges:
            if hi > lo:
                self.handles.append(dist.all_reduce(flat_grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def start(self, flat_grads):
        """launch the all-reduces (SUM); returns immediately"""
        self.handles = []
        if self.world == 1:
            return
        for lo, hi in bucket_ranges(flat_grads.numel(), self.n_buckets):
            self.handles.append(dist.all_reduce(flat_grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """wait for the exchange; returns the factor the optimizer must apply to the summed gradients"""
        for h in self.handles:
            h.wait()
        self.handles = []
        return 1.0 / self.world


def broadcast_buffers(bnbuf, counts, src=0, group=None):
    """DDP's broadcast_buffers=True semantics: rank-0 BatchNorm running statistics to every rank"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.broadcast(bnbuf, src, group=group)
    for c in counts:
        dist.broadcast(c, src, group=group)
