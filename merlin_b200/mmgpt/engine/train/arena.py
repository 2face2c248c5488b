"""Flat parameter / gradient / optimizer-state arenas.

B200-first replacement for FSDP's flat parameters (the reference trains with torch FSDP
`shard_grad_op auto_wrap`, pretrain.sh:32-33): every trainable tensor lives in ONE contiguous bf16 buffer
(weights), with a gradient buffer and two AdamW moment buffers of the same layout.  That makes the
optimizer one fused kernel launch, the gradient norm one reduction, and the data-parallel exchange a
handful of all-reduces over contiguous slices.  Segments are laid out in the order their gradients
become final during backward (lm_head first, embeddings / projector last), so a bucket can be reduced
as soon as backward has walked past it.  Each segment starts on a `chunk` boundary; a chunk carries one
(lr multiplier, weight decay) pair for the LLRD param groups.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch

CHUNK = 8192  # elements per hyper-parameter chunk (16 KB of bf16)


class ParamArena:
    def __init__(self, named_params: List[Tuple[str, torch.nn.Parameter]], hyper: Callable[[str, tuple], Tuple[float, float]],
                 device=None, chunk: int = CHUNK, with_state: bool = True):
        self.chunk = chunk
        self.names: List[str] = []
        self.offsets: Dict[str, Tuple[int, int, tuple]] = {}
        total = 0
        lr_mult, wd = [], []
        for name, p in named_params:
            n = p.numel()
            padded = (n + chunk - 1) // chunk * chunk
            self.offsets[name] = (total, n, tuple(p.shape))
            self.names.append(name)
            m, w = hyper(name, tuple(p.shape))
            lr_mult += [m] * (padded // chunk)
            wd += [w] * (padded // chunk)
            total += padded
        device = device if device is not None else named_params[0][1].device
        dtype = named_params[0][1].dtype
        self.numel = total
        self.params = torch.zeros(total, dtype=dtype, device=device)
        self.grads = torch.zeros(total, dtype=dtype, device=device)
        self.exp_avg = torch.zeros(total, dtype=dtype, device=device) if with_state else None
        self.exp_avg_sq = torch.zeros(total, dtype=dtype, device=device) if with_state else None
        self.lr_mult = torch.tensor(lr_mult, dtype=torch.float32, device=device)
        self.weight_decay = torch.tensor(wd, dtype=torch.float32, device=device)
        # move the parameters into the arena (the model's tensors become views of it)
        with torch.no_grad():
            for name, p in named_params:
                off, n, shape = self.offsets[name]
                view = self.params[off:off + n].view(shape)
                view.copy_(p.data)
                p.data = view

    def grad(self, name: str) -> torch.Tensor:
        off, n, shape = self.offsets[name]
        return self.grads[off:off + n].view(shape)

    def param(self, name: str) -> torch.Tensor:
        off, n, shape = self.offsets[name]
        return self.params[off:off + n].view(shape)

    def span(self, names: List[str]) -> Tuple[int, int]:
        """[start, end) element range covering the given (contiguous in arena order) parameters."""
        starts = [self.offsets[n][0] for n in names]
        ends = []
        for n in names:
            off, cnt, _ = self.offsets[n]
            ends.append(off + (cnt + self.chunk - 1) // self.chunk * self.chunk)
        return min(starts), max(ends)
