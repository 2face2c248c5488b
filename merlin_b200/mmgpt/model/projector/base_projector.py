"""Checkpoint discovery shared by the projectors — behaviour of
mmgpt/model/projector/base_projector.py:12-48 (pull the `model.projector.*` slice out of
pytorch_model.bin[.index.json]; silently keep the init when none is found)."""
import json
import os
from collections import defaultdict

import torch
import torch.nn as nn


class BaseProjector(nn.Module):
    def load_weights(self, model_path):
        if not model_path:
            return
        prefix = "model.projector."
        index = os.path.join(model_path, "pytorch_model.bin.index.json")
        single = os.path.join(model_path, "pytorch_model.bin")
        weights = {}
        if os.path.exists(index):
            by_file = defaultdict(list)
            for k, v in json.load(open(index))["weight_map"].items():
                if "projector" in k:
                    by_file[v].append(k)
            for fname, keys in by_file.items():
                ckpt = torch.load(os.path.join(model_path, fname), map_location="cpu")
                for k in keys:
                    weights[k[len(prefix):]] = ckpt[k]
        elif os.path.exists(single):
            ckpt = torch.load(single, map_location="cpu")
            weights = {k[len(prefix):]: v for k, v in ckpt.items() if "projector" in k}
        if weights:
            self.load_state_dict(weights, strict=True)

    @property
    def dtype(self):
        return self.bias.dtype

    @property
    def device(self):
        return self.bias.device
