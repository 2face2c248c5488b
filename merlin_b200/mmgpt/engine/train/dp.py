"""Data-parallel gradient exchange: bucketed all-reduce over slices of the flat gradient arena.

north_star: "the interpair / interleave multi-image batch shards over the 8 GPUs of one box with a single
NCCL allreduce on the gradient buckets (data-parallel only, no tensor parallel)".  Replaces the FSDP
reduce-scatter / all-gather of the reference.  Buckets are contiguous arena ranges in
backward-completion order; `ready(upto)` launches (asynchronously, on a side stream when on CUDA) the
all-reduce of every bucket that lies entirely below `upto`, so the exchange overlaps the remaining
backward GEMMs; `finish()` waits for all of them.  Gradients are pre-scaled by 1/world_size in the loss
backward, so SUM yields the mean.  Device-agnostic (works on CPU tensors with the gloo backend).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, flat_grads: torch.Tensor, bucket_elems: int = 128 * 1024 * 1024, group=None):
        self.flat = flat_grads
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        n = flat_grads.numel()
        self.buckets: List[Tuple[int, int]] = [(a, min(n, a + bucket_elems)) for a in range(0, n, bucket_elems)]
        self.next_bucket = 0
        self.handles = []
        self.comm_stream = torch.cuda.Stream() if flat_grads.is_cuda else None
        self.bytes_reduced = 0

    def start_step(self):
        self.next_bucket = 0
        self.handles = []

    def ready(self, upto: int):
        """Gradients in arena range [0, upto) are final: reduce every complete bucket below it."""
        if self.world == 1:
            return
        while self.next_bucket < len(self.buckets) and self.buckets[self.next_bucket][1] <= upto:
            a, b = self.buckets[self.next_bucket]
            self._launch(self.flat[a:b])
            self.next_bucket += 1

    def _launch(self, t: torch.Tensor):
        self.bytes_reduced += t.numel() * t.element_size()
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record()                                    # compute stream: this bucket's grads are written
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                self.handles.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.handles.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Flush the remaining buckets and make the current stream wait for every all-reduce."""
        if self.world == 1:
            return
        self.ready(self.flat.numel())
        for h in self.handles:
            h.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.handles = []
