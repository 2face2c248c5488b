"""ConvProjector on B200 kernels — mirror of mmgpt/model/projector/conv_projector.py:8-39.

Conv2d(C -> D, 3x3, stride s, pad 1) over the sqrt(P) x sqrt(P) patch grid, computed as one
im2col gather (reading the channel-last ViT features in place, no permute copies) + one tcgen05 GEMM
[n*P', 9C] x [9C, D] with the bias in the epilogue, batched over every image of every sample (the
reference loops over samples in Python).  The weight is stored as [D, 3, 3, C] (k = tap*C + c) to match
the gather; state_dict()/load_state_dict() use the reference layout `projector.weight [D,C,3,3]`.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from .... import ops
from .base_projector import BaseProjector


class ConvProjector(BaseProjector):
    def __init__(self, args, vision_hidden_size, lm_hidden_size, conv_stride=1):
        super().__init__()
        self.conv_stride = conv_stride
        self.freeze_projector = args.freeze_projector
        self.in_channels, self.out_channels = vision_hidden_size, lm_hidden_size
        g = not self.freeze_projector
        self.weight = nn.Parameter(torch.zeros(lm_hidden_size, 9 * vision_hidden_size), requires_grad=g)
        self.bias = nn.Parameter(torch.zeros(lm_hidden_size), requires_grad=g)
        with torch.no_grad():  # nn.Conv2d default init (kaiming_uniform, a=sqrt(5))
            bound = 1.0 / math.sqrt(9 * vision_hidden_size)
            self.weight.uniform_(-bound, bound)
            self.bias.uniform_(-bound, bound)
        self.load_weights(getattr(args, "model_name_or_path", None))

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = destination if destination is not None else OrderedDict()
        w = self.weight.view(self.out_channels, 3, 3, self.in_channels).permute(0, 3, 1, 2)  # [D,C,3,3] view
        out[prefix + "projector.weight"] = w if keep_vars else w.detach()
        out[prefix + "projector.bias"] = self.bias if keep_vars else self.bias.detach()
        return out

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = self.state_dict(keep_vars=True)
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"ConvProjector.load_state_dict: missing {missing} unexpected {unexpected}")
        with torch.no_grad():
            for k, dst in own.items():
                if k in state_dict:
                    dst.copy_(state_dict[k])
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def project(self, feat: torch.Tensor, n_img: int, grid: int, tok_stride: int, tok_off: int) -> torch.Tensor:
        """feat: bf16 buffer whose image i / patch (y,x) row sits at (i*tok_stride + tok_off + y*grid + x)."""
        return ops.projector_forward(feat, self.weight, self.bias, n_img, grid, self.in_channels, self.conv_stride,
                                     tok_stride, tok_off)

    def forward(self, features):
        with torch.set_grad_enabled(False):
            fixed = []
            for f in features:
                if f.dim() == 1:  # dummy feature: conv_projector.py:27-28 tiles it to 256 patches
                    f = f.view(1, 1, -1).repeat(1, 256, 1)
                fixed.append(f)
            sizes = [f.shape[0] for f in fixed]
            P = fixed[0].shape[1]
            grid = int(math.sqrt(P))
            x = torch.cat([f.reshape(-1, f.shape[-1]) for f in fixed], dim=0).to(torch.bfloat16).contiguous()
            y = self.project(x, sum(sizes), grid, P, 0)
            Po = y.shape[0] // sum(sizes)
            return list(torch.split(y.view(sum(sizes), Po, -1), sizes, dim=0))
