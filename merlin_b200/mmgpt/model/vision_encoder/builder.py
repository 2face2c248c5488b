"""build_vision_tower — mirror of mmgpt/model/vision_encoder/builder.py:7-17 (CLIP branch only;
the Qwen / SAM / MetaCLIP towers are outside the Merlin hot path)."""
from .clip_encoder import CLIPVisionTower


def build_vision_tower(vision_tower_cfg, config=None):
    vision_tower = getattr(vision_tower_cfg, "vision_tower", None)
    if vision_tower is not None and "clip" in vision_tower.lower():
        return CLIPVisionTower(args=vision_tower_cfg, config=config)
    raise ValueError(f"Unknown vision tower: {vision_tower}")
