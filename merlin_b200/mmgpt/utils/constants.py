"""Boundary constants of the hot path (reference: mmgpt/utils/constants.py:7-21).

Only the model constants are mirrored; the dataset registries of the reference (private S3 paths,
constants.py:25-354) are out of scope.
"""
IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IM_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
DEFAULT_PAD_TOKEN = "[PAD]"
DEFAULT_EOS_TOKEN = "</s>"
DEFAULT_BOS_TOKEN = "</s>"
DEFAULT_UNK_TOKEN = "<unk>"

# Vicuna-v1.5 ids after build_vision_tokenizer (mm_utils.py:105 hard-codes 32000)
IM_PATCH_TOKEN_ID = 32000
IM_START_TOKEN_ID = 32001
IM_END_TOKEN_ID = 32002
