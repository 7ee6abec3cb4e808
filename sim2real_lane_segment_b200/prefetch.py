"""Double-buffered host -> device staging of training batches on a copy stream.

The reference feeds `training_step` from a `DataLoader(pin_memory=...)` and Lightning moves each batch to the GPU in
the fit loop (train.py:48) on the compute stream, so every step starts with a 25 MB copy nothing
overlaps.  `DevicePrefetcher` wraps any iterable of pinned `(x, y)` host batches and yields device tensors whose copy
was issued one step earlier on a separate stream: the copy of batch i+1 runs under the kernels of step i."""
import torch


class DevicePrefetcher:
    def __init__(self, batches, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePrefetcher: CUDA is required")
        self.device = torch.device(device if device is not None else "cuda")
        self.batches = batches
        self.copy_stream = torch.cuda.Stream(self.device)

    def __iter__(self):
        it = iter(self.batches)
        bufs = [None, None]
        copied = [torch.cuda.Event(), torch.cuda.Event()]     # copy of the batch in slot s has finished
        released = [None, None]                                # compute stream is done with the batch in slot s

        def stage(slot, batch):
            hx, hy = batch
            if bufs[slot] is None or bufs[slot][0].shape != hx.shape or bufs[slot][1].shape != hy.shape:
                bufs[slot] = (torch.empty(hx.shape, dtype=hx.dtype, device=self.device),
                              torch.empty(hy.shape, dtype=hy.dtype, device=self.device))
            with torch.cuda.stream(self.copy_stream):
                if released[slot] is not None:
                    self.copy_stream.wait_event(released[slot])
                bufs[slot][0].copy_(hx, non_blocking=True)
                bufs[slot][1].copy_(hy, non_blocking=True)
                copied[slot].record(self.copy_stream)

        try:
            nxt = next(it)
        except StopIteration:
            return
        stage(0, nxt)
        slot = 0
        while True:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(copied[slot])
            out = bufs[slot]
            try:
                nxt = next(it)
                stage(slot ^ 1, nxt)     # its last reader (two steps ago) recorded released[slot ^ 1]
            except StopIteration:
                nxt = None
            yield out
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))   # everything the consumer enqueued for this batch
            released[slot] = ev
            if nxt is None:
                return
            slot ^= 1
