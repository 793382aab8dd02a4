"""Host -> device batch feed for the training loop.

The reference moves every batch with a synchronous `.to(device)` after a `DataLoader(pin_memory=True)` handed
it over (train_bilinear.py:33-43,72-73).  At batch 64 that copy is 20 KB and rides inside
`BilinearUnit.train_step_host`; at batch 65 536 it is 21 MB -- a third of a millisecond of PCIe time per
~2 ms step if it sits on the compute stream.  `PinnedPrefetcher` keeps the copy of batch i+1 in flight on a
side stream while iteration i runs (two device buffers, CUDA events both ways), so every step still copies its
own inputs from pinned host memory, just not in series with the kernel.
"""
import torch


class PinnedPrefetcher:
    """Iterate over an iterable of (x_host [B,32], y_host [B,48]) pinned float32 CPU tensors, yielding device
    tensors; the copy of the NEXT pair is issued before the current pair is handed out.

    The yielded tensors are views of two rotating device buffers: they stay valid until the next-but-one
    `next()` (the usual contract of a prefetching loader: consume the batch in the iteration that received it).
    """

    def __init__(self, batches, device, depth=2):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None] * self.depth          # (x_dev, y_dev) per slot, allocated on first use
        self.ready = [None] * self.depth         # copy finished (recorded on the copy stream)
        self.freed = [None] * self.depth         # consumer finished (recorded on the compute stream)
        self.head = 0                            # next slot to hand out
        self.tail = 0                            # next slot to fill
        self.issued = 0                          # slots whose copy was issued and that were not released yet
        self.last = None                         # slot handed out by the previous next()
        self.h2d_bytes = 0
        self._fill()

    def _fill(self):
        while self.issued < self.depth:
            try:
                hx, hy = next(self.it)
            except StopIteration:
                return
            s = self.tail
            if self.bufs[s] is None or self.bufs[s][0].shape != hx.shape or self.bufs[s][1].shape != hy.shape:
                self.bufs[s] = (torch.empty(hx.shape, dtype=torch.float32, device=self.device),
                                torch.empty(hy.shape, dtype=torch.float32, device=self.device))
            with torch.cuda.stream(self.copy_stream):
                if self.freed[s] is not None:
                    self.copy_stream.wait_event(self.freed[s])   # the step that read this slot has finished
                self.bufs[s][0].copy_(hx, non_blocking=True)
                self.bufs[s][1].copy_(hy, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self.ready[s] = ev
            self.h2d_bytes += hx.numel() * 4 + hy.numel() * 4
            self.tail = (s + 1) % self.depth
            self.issued += 1

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.last is not None:
            ev = torch.cuda.Event()
            ev.record(cur)                       # everything enqueued so far may have read the previous slot
            self.freed[self.last] = ev
            self.issued -= 1
            self.last = None
        self._fill()                             # the slot just released takes the copy of the batch after next
        if self.issued == 0:
            raise StopIteration
        s = self.head
        cur.wait_event(self.ready[s])
        self.head = (s + 1) % self.depth
        self.last = s
        return self.bufs[s]
