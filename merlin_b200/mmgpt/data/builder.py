"""build_dataloader — mirror of mmgpt/data/builder.py:12-81 for the interpair datasets.

Same `multimodal_cfg` dict, same return value (`train_dataset`, `eval_dataset`, `data_collator`).  The
reference resolves dataset NAMES through the PAIR_WEBDATA table of constants.py (S3 globs); here
`data_args.interpair_webdatasets` is a '+'-separated list of entries `path_glob[@merge_round]` that name LOCAL
tar shards.  The other dataset kinds (conversation / pair / pair_token / interleave) are outside SURVEY.md §8f
N2 and raise."""
from __future__ import annotations

import glob

import torch

from .collator import DataCollatorForSupervisedDataset
from .dataset.interpair_webdataset import InterPairWebDataset


def build_dataloader(tokenizer, data_args):
    data_collator = DataCollatorForSupervisedDataset(tokenizer=tokenizer)
    multimodal_cfg = dict(
        image_token_len=data_args.image_token_len,
        image_aspect_ratio=getattr(data_args, "image_aspect_ratio", "resize"),
        use_im_start_end=data_args.use_im_start_end,
        image_processor=getattr(data_args, "image_processor", None),
        box_limit=getattr(data_args, "box_limit", 0),
        image_size=data_args.image_size,
    )
    for kind in ("conversation_datasets", "pair_webdatasets", "pair_token_webdatasets", "interleave_webdatasets"):
        if getattr(data_args, kind, None):
            raise NotImplementedError(f"{kind}: only the interpair datasets are mirrored (SURVEY.md §8f N2)")
    if multimodal_cfg["image_aspect_ratio"] not in ("resize", "pad"):
        raise NotImplementedError("image_aspect_ratio: the GPU image path covers 'resize' and 'pad' ('keep' yields "
                                  "variable-size images that the fixed-resolution CLIP tower cannot take)")
    datasets = []
    for entry in (getattr(data_args, "interpair_webdatasets", None) or "").split("+"):
        if not entry:
            continue
        pattern, _, rounds = entry.partition("@")
        shards = sorted(glob.glob(pattern))
        if not shards:
            raise FileNotFoundError(f"no tar shards match {pattern!r}")
        datasets.append(InterPairWebDataset(shards, tokenizer, multimodal_cfg, merge_round=int(rounds or 1)))
    if not datasets:
        raise ValueError("build_dataloader: data_args.interpair_webdatasets names no dataset")
    train_dataset = datasets[0] if len(datasets) == 1 else torch.utils.data.ConcatDataset(datasets)
    return dict(train_dataset=train_dataset, eval_dataset=None, data_collator=data_collator)
