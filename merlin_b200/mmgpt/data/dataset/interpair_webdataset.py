"""InterPairWebDataset over LOCAL tar shards — mirror of mmgpt/data/dataset/interpair_webdataset.py:16-185.

What is kept: the token grammar (`add_image_token`, :53-61), the prompt / text tokenisation with the prompt
and every image token masked out of the labels (`token_processor`, :63-152), the merge of `merge_round` records
into one interleaved training sample (`__getitem__`, :157-185) and the zero-image fallback (:143-144).  What is
replaced: the webdataset / megfile / S3 pipeline (:30-41) by a plain `tarfile` reader over local shards with the
same record format (one `<key>.json` with `prompt`, `txt`, `image_name_list`, plus `<suffix>.jpg` members) that
keeps the pipeline's data-parallel semantics — an infinite, per-epoch reshuffled shard list (seed 3407), a
shard-level shuffle buffer of world_size, `split_by_node` (rank r takes every world_size-th shard),
`split_by_worker` (DataLoader worker w every num_workers-th of those) and a 1000-sample shuffle buffer (seed 42) —
so ranks and workers read DISJOINT shards instead of identical samples; and
the PIL + CLIPImageProcessor image path (`base_dataset.py:179-198`, mode 'resize') by the GPU preprocessing
kernel (`ops.image_preprocess`: bit-exact PIL bicubic resize on the uint8 pixels + CLIP mean / std normalise).
"""
from __future__ import annotations

import io
import itertools
import json
import os
import random
import tarfile
from typing import Dict, Iterator, List, Optional, Sequence

import torch

from ...utils.constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_PATCH_TOKEN, DEFAULT_IM_START_TOKEN,
                                DEFAULT_IMAGE_TOKEN, IGNORE_INDEX)


def iter_tar_records(shards: Sequence[str]) -> Iterator[dict]:
    """Groups tar members by the basename before the first '.', like webdataset's tarfile_to_samples: yields
    {"json": {...}, "<name>.jpg": bytes, ...} per record."""
    for shard in shards:
        with tarfile.open(shard) as tf:
            cur_key, cur = None, {}
            for m in tf:
                if not m.isfile():
                    continue
                base = m.name.rsplit("/", 1)[-1]
                key, _, ext = base.partition(".")
                if key != cur_key and cur:
                    yield cur
                    cur = {}
                cur_key = key
                data = tf.extractfile(m).read()
                if ext == "json":
                    cur["json"] = json.loads(data)
                else:
                    cur[ext] = data
            if cur:
                yield cur


def shuffle_buffer(src: Iterator, bufsize: int, rng: random.Random) -> Iterator:
    """webdataset's `shuffle(bufsize, rng=...)` (the `_shuffle` filter): keep `bufsize` items, swap each incoming
    item with a random slot and emit what was there; drain the buffer at the end of a finite stream."""
    if bufsize <= 1:
        yield from src
        return
    buf = []
    for item in src:
        if len(buf) < bufsize:
            buf.append(item)
            continue
        k = rng.randint(0, bufsize - 1)
        item, buf[k] = buf[k], item
        yield item
    rng.shuffle(buf)
    yield from buf


def rank_and_world():
    """(rank, world_size) of this process: torch.distributed when initialised, else torchrun's environment."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class InterPairWebDataset:
    """Interleave-format dataset for stage-1 pre-training (reference docstring, :17)."""

    SHARD_SEED, SHUFFLE_SEED, SAMPLE_BUFFER = 3407, 42, 1000      # interpair_webdataset.py:32-37

    def __init__(self, shards: Sequence[str], tokenizer, multimodal_cfg: dict, merge_round: int = 1,
                 data_size: Optional[int] = None, image_loader=None, rank: Optional[int] = None,
                 world_size: Optional[int] = None, shuffle: bool = True):
        self.tokenizer = tokenizer
        self.multimodal_cfg = multimodal_cfg
        self.image_size = multimodal_cfg["image_size"]
        self.merge_round = merge_round
        self.use_im_start_end = multimodal_cfg["use_im_start_end"]
        self.im_patch_token, self.im_start_token, self.im_end_token = tokenizer.convert_tokens_to_ids(
            [DEFAULT_IM_PATCH_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN])
        self.replace_token = DEFAULT_IM_PATCH_TOKEN * multimodal_cfg["image_token_len"]
        if self.use_im_start_end:
            self.replace_token = DEFAULT_IM_START_TOKEN + self.replace_token + DEFAULT_IM_END_TOKEN
        self.shards = list(shards)
        if not self.shards:
            raise RuntimeError("InterPairWebDataset: no shards given")
        self.rank, self.world_size, self.shuffle = rank, world_size, shuffle
        self._inner_iter = None           # built on first use, i.e. inside the DataLoader worker that consumes it
        self.data_size = data_size if data_size is not None else sum(1 for _ in iter_tar_records(self.shards))
        self.image_loader = image_loader or self._load_image

    # ------------------------------------------------------------------ the data pipeline (:30-41)
    def _shard_stream(self) -> Iterator[str]:
        """wds.InfiniteShardList(shards, seed=3407): every epoch walks the shard list in a freshly seeded order."""
        epoch = 0
        while True:
            order = list(self.shards)
            if self.shuffle:
                random.Random(self.SHARD_SEED + epoch).shuffle(order)
            yield from order
            epoch += 1

    def _pipeline(self) -> Iterator[dict]:
        rank, world = rank_and_world()
        rank = self.rank if self.rank is not None else rank
        world = self.world_size if self.world_size is not None else world
        info = torch.utils.data.get_worker_info()
        wid, nworkers = (info.id, info.num_workers) if info is not None else (0, 1)
        stream = self._shard_stream()
        if self.shuffle:
            stream = shuffle_buffer(stream, world, random.Random(self.SHUFFLE_SEED))     # wds.shuffle(world_size)
        stream = itertools.islice(stream, rank, None, world)                             # wds.split_by_node
        stream = itertools.islice(stream, wid, None, nworkers)                           # wds.split_by_worker

        def records():
            empty_in_a_row = 0
            for shard in stream:
                n = 0
                for rec in iter_tar_records([shard]):
                    n += 1
                    yield rec
                empty_in_a_row = 0 if n else empty_in_a_row + 1
                if empty_in_a_row > len(self.shards):
                    raise RuntimeError("InterPairWebDataset: no records in the given shards")

        if self.shuffle:
            return shuffle_buffer(records(), self.SAMPLE_BUFFER, random.Random(self.SHUFFLE_SEED))  # wds.shuffle(1000)
        return records()

    @property
    def inner_iter(self):
        if self._inner_iter is None:
            self._inner_iter = self._pipeline()
        return self._inner_iter

    # ------------------------------------------------------------------ image path (base_dataset.py:179-198)
    def _load_image(self, data: bytes) -> torch.Tensor:
        """'resize' mode: decode -> RGB -> square bicubic resize -> /255 -> CLIP mean / std ('pad' mode: the image
        is first expanded to a square on the mean colour, base_dataset.py:186-188).  The decode is PIL's
        (JPEG entropy decoding is not part of this path); resize + normalise run in merlin_image_preprocess on
        the GPU — there is no CPU fallback: without a CUDA device / the library this raises."""
        import numpy as np
        from PIL import Image
        from .... import ops
        if not torch.cuda.is_available():
            raise RuntimeError("InterPairWebDataset: the image path runs on the GPU (merlin_image_preprocess); "
                               "pass image_loader= to supply pre-processed tensors on a machine without one")
        image = Image.open(io.BytesIO(data)).convert("RGB")
        proc = self.multimodal_cfg.get("image_processor")
        mean = getattr(proc, "image_mean", None) or (0.48145466, 0.4578275, 0.40821073)
        std = getattr(proc, "image_std", None) or (0.26862954, 0.26130258, 0.27577711)
        mode = self.multimodal_cfg.get("image_aspect_ratio", "resize")
        if mode not in ("resize", "pad"):
            raise NotImplementedError(f"image_aspect_ratio={mode!r}: the GPU image path covers 'resize' and 'pad'")
        pixels = torch.from_numpy(np.array(image)).cuda()      # np.array: a writable copy of PIL's buffer
        return ops.image_preprocess(pixels, self.image_size, mean, std, pad_to_square=(mode == "pad")).cpu()

    # ------------------------------------------------------------------ token grammar (:53-152)
    def add_image_token(self, text):
        if DEFAULT_IMAGE_TOKEN + "\n" in text:
            text = text.replace(DEFAULT_IMAGE_TOKEN, self.replace_token)
        elif DEFAULT_IMAGE_TOKEN in text:
            text = text.replace(DEFAULT_IMAGE_TOKEN, self.replace_token + "\n")
        else:
            text = self.replace_token + "\n" + text
        return text

    def token_processor(self, text_list, image_list):
        img_count = 0
        input_ids, targets = [], []
        max_len = self.tokenizer.model_max_length
        for i, (prompt, text) in enumerate(text_list):
            cur_image_count = (prompt.count(DEFAULT_IMAGE_TOKEN) if prompt is not None else 0) + \
                text.count(DEFAULT_IMAGE_TOKEN)
            if cur_image_count == 0:
                cur_image_count += 1
            if prompt is not None:
                prompt = self.add_image_token(prompt)
            else:
                text = self.add_image_token(text)
            prompt_input_ids = self.tokenizer(prompt, padding="longest", max_length=max_len,
                                              truncation=True).input_ids if prompt is not None else []
            text_input_ids = self.tokenizer(text + self.tokenizer.eos_token, padding="longest",
                                            max_length=max_len - len(prompt_input_ids), truncation=True).input_ids
            if prompt is not None and text_input_ids[0] == 1:      # drop the duplicated BOS (:103-104)
                text_input_ids = text_input_ids[1:]
            if len(input_ids) + len(prompt_input_ids) + len(text_input_ids) > max_len:
                image_list = image_list[:img_count]                # clip the images with the dropped pairs (:112)
                break
            input_ids.extend(prompt_input_ids + text_input_ids)
            targets.extend([IGNORE_INDEX] * len(prompt_input_ids) + text_input_ids)
            img_count += cur_image_count
        input_ids = torch.tensor(input_ids, dtype=torch.long)
        targets = torch.tensor(targets, dtype=torch.long)
        image_indices = (targets == self.im_patch_token)
        if self.use_im_start_end:
            image_indices = torch.logical_or(image_indices, torch.logical_or(targets == self.im_start_token,
                                                                             targets == self.im_end_token))
        targets = targets.masked_fill(image_indices, IGNORE_INDEX)
        if len(image_list) == 0:
            image_list = [torch.zeros(3, self.image_size, self.image_size)]
        return dict(image=image_list, input_ids=input_ids, labels=targets)

    def __len__(self):
        return self.data_size // self.merge_round

    def __getitem__(self, i) -> Dict[str, torch.Tensor]:
        text_list: List = []
        image_list: List[torch.Tensor] = []
        for _ in range(self.merge_round):
            sample = next(self.inner_iter)
            try:
                images = [self.image_loader(sample[name.split("-")[-1] + ".jpg"])
                          for name in sample["json"]["image_name_list"]]
            except Exception:                                       # a broken record is skipped (:171-173)
                continue
            image_list.extend(images)
            text_list.append((sample["json"]["prompt"], sample["json"]["txt"]))
        return self.token_processor(text_list, image_list)
