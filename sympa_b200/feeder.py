"""Host -> device feed of the training triplets.

The reference keeps all triplets on the device (train.py:95-96); for graphs whose pair list does
not fit (5e11 pairs, sampled) or that arrive from a host-side sampler, the
batch (index pairs int64 (b, 2), graph distances float64 (b,)) has to cross PCIe every step.
PairFeeder double-buffers that copy on a side stream: while batch k computes, batch k + 1 is on the
wire, so a step costs max(copy, compute) instead of their sum.  Host tensors must be pinned.  The index
pairs may be int32 on the host (node ids fit; half the bytes) and the graph distances small integers (uint8 /
int16: shortest-path lengths of an unweighted graph, preprocess.py:118-126): both are widened on the device -
9 bytes per pair on the wire instead of 24.
"""
import torch


class PairFeeder:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None] * depth            # (idx_d, gd_d) per slot
        self._ready = [torch.cuda.Event() for _ in range(depth)]     # copy of the slot has landed
        self._consumed = [None] * depth        # compute that read the slot has been enqueued/finished
        self._head = 0                         # next slot to fill
        self._tail = 0                         # next slot to hand out
        self._in_flight = 0

    def submit(self, idx_h, gd_h):
        """Enqueue the asynchronous copy of one batch (pinned host tensors) into the next free slot."""
        assert self._in_flight < self.depth, "PairFeeder: all slots in flight (call next()/done() first)"
        assert idx_h.is_pinned() and gd_h.is_pinned(), "PairFeeder needs pinned host memory"
        s = self._head
        narrow = idx_h.dtype == torch.int32     # node ids as int32 on the host: 8 instead of 16 bytes per pair on the wire
        small = gd_h.dtype != torch.float64     # graph distances as small integers (uint8 / int16) on the host
        buf = self._bufs[s]
        if (buf is None or buf[0].shape != idx_h.shape or buf[1].shape != gd_h.shape or (buf[2] is None) == narrow
                or (buf[3] is None) == small or (small and buf[3].dtype != gd_h.dtype)):
            buf = (torch.empty(idx_h.shape, dtype=torch.int64, device=self.device),
                   torch.empty(gd_h.shape, dtype=torch.float64, device=self.device),
                   torch.empty(idx_h.shape, dtype=torch.int32, device=self.device) if narrow else None,
                   torch.empty(gd_h.shape, dtype=gd_h.dtype, device=self.device) if small else None)
            self._bufs[s] = buf
        with torch.cuda.stream(self.copy_stream):
            if self._consumed[s] is not None:
                self.copy_stream.wait_event(self._consumed[s])   # the step that read this slot is done with it
            if narrow:
                buf[2].copy_(idx_h, non_blocking=True)
                buf[0].copy_(buf[2])                             # widened on the device (the kernels take int64 indices)
            else:
                buf[0].copy_(idx_h, non_blocking=True)
            if small:
                buf[3].copy_(gd_h, non_blocking=True)
                buf[1].copy_(buf[3])                             # widened to float64 on the device
            else:
                buf[1].copy_(gd_h, non_blocking=True)
            self._ready[s].record(self.copy_stream)
        self._head = (s + 1) % self.depth
        self._in_flight += 1

    def next(self):
        """Device tensors of the oldest submitted batch; the current stream waits for their copy."""
        assert self._in_flight > 0, "PairFeeder: nothing submitted"
        s = self._tail
        torch.cuda.current_stream(self.device).wait_event(self._ready[s])
        self._cur = s
        return self._bufs[s][0], self._bufs[s][1]

    def done(self):
        """Call after the last kernel that reads the batch handed out by next() has been enqueued."""
        s = self._cur
        ev = self._consumed[s] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._consumed[s] = ev
        self._tail = (s + 1) % self.depth
        self._in_flight -= 1
