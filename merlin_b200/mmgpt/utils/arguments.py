"""ModelArguments / DataArguments field names of the reference (mmgpt/utils/arguments.py:6-36),
as plain dataclasses (no HfArgumentParser / TrainingArguments dependency on the hot path)."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class ModelArguments:
    model_name_or_path: Optional[str] = None
    vision_tower: Optional[str] = "clip-vit-large-patch14-336"
    projector: Optional[str] = "mlp"
    freeze_vision_tower: bool = False
    freeze_projector: bool = False
    freeze_lm_model: bool = False
    vision_select_layer: Optional[int] = -2
    vision_select_feature: Optional[str] = "patch"
    use_im_start_end: bool = True
    conv_stride: int = 1


@dataclass
class DataArguments:
    image_size: int = 224
    num_patches: int = 256
    image_aspect_ratio: str = "square"
    use_beam_search: bool = False
    image_token_len: int = 256
    image_processor: object = None
    use_im_start_end: bool = True
