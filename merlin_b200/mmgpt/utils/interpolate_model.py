"""Position-embedding interpolation for checkpoints whose vision tower was trained at another resolution.

Mirror of mmgpt/utils/interpolate_model.py:7-31 (the script that turns the 336 px CLIP-L/14 table of
577 rows into the 448 px table of 1025 rows the published Merlin weights use): the class-token row is kept, the
grid rows are reshaped to [1, C, g, g], resized with bicubic interpolation (align_corners=True) and flattened
back.  The reference hard-codes "double the grid"; here the target grid is a parameter, and the in-memory form
`interpolate_position_embedding` is what `CLIPVisionTower.load_state_dict(..., interpolate=True)` calls when a
checkpoint's table does not match the tower's image size.  One-off host-side weight surgery (PyTorch on the
checkpoint's device), not part of the per-step path.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

POS_KEY = "model.vision_tower.vision_tower.vision_model.embeddings.position_embedding.weight"
IDS_KEY = "model.vision_tower.vision_tower.vision_model.embeddings.position_ids"


def interpolate_position_embedding(old_pos_embed: torch.Tensor, grid_size: int, extra_tokens: int = 1) -> torch.Tensor:
    """[extra + g_old^2, C] -> [extra + grid_size^2, C]  (interpolate_model.py:11-28)."""
    n_grid = old_pos_embed.shape[0] - extra_tokens
    old_grid = round(n_grid ** 0.5)
    if old_grid * old_grid != n_grid:
        raise ValueError(f"position table of {old_pos_embed.shape[0]} rows is not {extra_tokens} + a square grid")
    if old_grid == grid_size:
        return old_pos_embed.clone()
    tok, img = old_pos_embed[:extra_tokens], old_pos_embed[extra_tokens:]
    dtype = img.dtype
    img = img.float().reshape(1, old_grid, old_grid, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=(grid_size, grid_size), mode="bicubic", align_corners=True)
    img = img.permute(0, 2, 3, 1).reshape(grid_size * grid_size, -1).to(dtype)
    return torch.cat([tok, img], dim=0)


def interpolate(path: str, shard: str = "pytorch_model-00002-of-00002.bin", scale: int = 2) -> str:
    """File form of the reference script (interpolate_model.py:7-31): rewrites the shard in place with the grid
    scaled by `scale` (2 in the reference: 336 px -> 672/…; the published recipe runs it on 224 -> 448)."""
    f = os.path.join(path, shard)
    model = torch.load(f, map_location=torch.device("cpu"))
    old = model[POS_KEY]
    grid = round((old.shape[0] - 1) ** 0.5) * scale
    new = interpolate_position_embedding(old, grid)
    model[IDS_KEY] = torch.arange(new.shape[0]).expand((1, -1))
    model[POS_KEY] = new
    torch.save(model, f)
    return f
