"""build_projector — mirror of mmgpt/model/projector/builder.py:8-41 ("mlp" and "conv"; the SAM /
Qwen projectors pair with towers that are outside the Merlin hot path)."""
from .conv_projector import ConvProjector
from .mlp_projector import MLPProjector


def build_projector(projector_cfg, vision_hidden_size, lm_hidden_size):
    projector = getattr(projector_cfg, "projector", None)
    if projector == "mlp":
        return MLPProjector(args=projector_cfg, vision_hidden_size=vision_hidden_size,
                            lm_hidden_size=lm_hidden_size)
    if projector == "conv":
        return ConvProjector(args=projector_cfg, vision_hidden_size=vision_hidden_size,
                             lm_hidden_size=lm_hidden_size, conv_stride=projector_cfg.conv_stride)
    raise ValueError(f"Unknown projector: {projector}")
