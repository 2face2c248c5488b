"""Mirror of mmgpt/utils/mm_utils.py — the helpers the eval / serve drivers call around generate():
`KeywordsStoppingCriteria` (:62-85), `expand2square` (:10-21), `process_image` (:28-46),
`get_model_name_from_path` (:53-59).  Host-side glue, same names, arguments and behaviour."""
from __future__ import annotations

import torch


def expand2square(pil_img, background_color):
    """Pad to a square canvas with the image pasted at the TOP-LEFT corner (mm_utils.py:10-21 — not centred,
    "for simpler box processing")."""
    from PIL import Image
    width, height = pil_img.size
    if width == height:
        return pil_img
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    result.paste(pil_img)
    return result


def process_image(image, processor, image_size=448, mode="resize"):
    """mm_utils.py:28-46: 'keep' (shortest edge from the aspect ratio, no crop), 'pad' (expand2square with the
    processor's mean colour, then resize), 'resize' (PIL resize to a square, no further resize / crop), else
    the processor's defaults."""
    if mode == "keep":
        max_hw, min_hw = max(image.size), min(image.size)
        aspect_ratio = max_hw / min_hw
        max_len, min_len = image_size * 2, image_size
        shortest_edge = int(min(max_len / aspect_ratio, min_len))
        return processor.preprocess(image, return_tensors="pt", do_center_crop=False,
                                    size={"shortest_edge": shortest_edge})["pixel_values"]
    if mode == "pad":
        image = expand2square(image, tuple(int(x * 255) for x in processor.image_mean))
        return processor.preprocess(image, return_tensors="pt", do_center_crop=False,
                                    size={"shortest_edge": image_size})["pixel_values"]
    if mode == "resize":
        image = image.resize((image_size, image_size))
        return processor.preprocess(image, return_tensors="pt", do_resize=False, do_center_crop=False)["pixel_values"]
    return processor.preprocess(image, return_tensors="pt")["pixel_values"]


def process_images(images, image_processor):
    return image_processor(images, return_tensors="pt")["pixel_values"]


def get_model_name_from_path(model_path):
    model_paths = model_path.strip("/").split("/")
    if model_paths[-1].startswith("checkpoint-"):
        return model_paths[-2] + "_" + model_paths[-1]
    return model_paths[-1]


class KeywordsStoppingCriteria:
    """mm_utils.py:62-85.  Stops when the generated ids END with the token ids of a keyword (the leading BOS a
    tokenizer adds is dropped), or when the text decoded from the last min(generated, 3) tokens of sequence 0
    contains a keyword.  Callable as criterion(output_ids, scores) like an HF StoppingCriteria."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.keyword_ids = []
        for keyword in keywords:
            cur_keyword_ids = tokenizer(keyword).input_ids
            if len(cur_keyword_ids) > 1 and cur_keyword_ids[0] == tokenizer.bos_token_id:
                cur_keyword_ids = cur_keyword_ids[1:]
            self.keyword_ids.append(torch.tensor(cur_keyword_ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def __call__(self, output_ids: torch.LongTensor, scores=None, **kwargs) -> bool:
        offset = min(output_ids.shape[1] - self.start_len, 3)
        self.keyword_ids = [k.to(output_ids.device) for k in self.keyword_ids]
        for keyword_id in self.keyword_ids:
            tail = output_ids[0, -keyword_id.shape[0]:]
            # the reference compares the tensors with `==` inside `if`, which is only well defined for
            # one-token keywords; all-equal is what it means for longer ones
            if tail.shape[0] == keyword_id.shape[0] and bool((tail == keyword_id).all()):
                return True
        if offset <= 0:
            return False
        outputs = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(keyword in outputs for keyword in self.keywords)
