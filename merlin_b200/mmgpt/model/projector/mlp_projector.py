"""MLPProjector on B200 kernels — mirror of mmgpt/model/projector/mlp_projector.py:11-23:
one nn.Linear(vision_hidden -> lm_hidden) per patch token, run as a single tcgen05 GEMM with the
bias in the epilogue over all samples at once."""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from .... import ops
from .base_projector import BaseProjector


class MLPProjector(BaseProjector):
    def __init__(self, args, vision_hidden_size, lm_hidden_size):
        super().__init__()
        self.freeze_projector = args.freeze_projector
        g = not self.freeze_projector
        self.weight = nn.Parameter(torch.zeros(lm_hidden_size, vision_hidden_size), requires_grad=g)
        self.bias = nn.Parameter(torch.zeros(lm_hidden_size), requires_grad=g)
        with torch.no_grad():
            bound = 1.0 / math.sqrt(vision_hidden_size)
            self.weight.uniform_(-bound, bound)
            self.bias.uniform_(-bound, bound)
        self.load_weights(getattr(args, "model_name_or_path", None))

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = destination if destination is not None else OrderedDict()
        out[prefix + "projector.weight"] = self.weight if keep_vars else self.weight.detach()
        out[prefix + "projector.bias"] = self.bias if keep_vars else self.bias.detach()
        return out

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = self.state_dict(keep_vars=True)
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"MLPProjector.load_state_dict: missing {missing} unexpected {unexpected}")
        with torch.no_grad():
            for k, dst in own.items():
                if k in state_dict:
                    dst.copy_(state_dict[k])
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def forward(self, features):
        with torch.set_grad_enabled(False):
            sizes = [f.shape[0] for f in features]
            P = features[0].shape[1]
            x = torch.cat([f.reshape(-1, f.shape[-1]) for f in features], dim=0).to(torch.bfloat16).contiguous()
            y = ops.gemm(x, self.weight, bias=self.bias)
            return list(torch.split(y.view(sum(sizes), P, -1), sizes, dim=0))
