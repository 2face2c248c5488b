"""MMGPTMetaForCausalLM on B200 kernels — mirror of mmgpt/model/mmgpt/base_mmgpt.py:12-165.

Same methods (`encode_images`, `build_vision_tokenizer`, `prepare_inputs_labels_for_multimodal`), same
return tuple, same errors.  What changes is the mechanism: the reference loops over samples and images
in Python and torch.cat-s feature rows into the embedded sequence (base_mmgpt.py:104-160, with host
syncs at :109/:116); here the token grammar is validated once on the host copy of input_ids and a
single gather/scatter kernel (merlin_embed_splice) writes `inputs_embeds` bit-exactly.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional

import torch

from .... import ops
from ...utils.constants import DEFAULT_IM_END_TOKEN, DEFAULT_IM_PATCH_TOKEN, DEFAULT_IM_START_TOKEN
from ..projector.builder import build_projector
from ..vision_encoder.builder import build_vision_tower


class MMGPTMetaForCausalLM(ABC):

    @abstractmethod
    def get_model(self):
        pass

    # ------------------------------------------------------------------ base_mmgpt.py:18-21
    def encode_images(self, images: List[torch.Tensor]):
        image_features = self.get_model().vision_tower(images)
        image_features = self.get_model().projector(image_features)
        return image_features

    def _encode_images_flat(self, images: List[torch.Tensor]):
        """Fast path used by forward: every image of every sample in one ViT pass, the projector reading
        the [n,577,C] hidden states in place (CLS skipped by offset), features returned as one
        contiguous [sum_i n_i * P', D] buffer + per-sample row counts."""
        tower, proj = self.get_model().vision_tower, self.get_model().projector
        concat = torch.cat([im for im in images], dim=0)
        n = concat.shape[0]
        hidden = tower.encode(concat)
        if tower.select_feature != "patch":
            feats = proj([tower.feature_select(hidden, n)])[0]
        elif hasattr(proj, "project"):  # ConvProjector
            grid = tower.config.image_size // tower.config.patch_size
            feats = proj.project(hidden, n, grid, tower.tokens, 1)
        else:
            feats = proj([tower.feature_select(hidden, n)])[0]
        feats = feats.reshape(-1, feats.shape[-1])
        per_image = feats.shape[0] // n
        return feats, [im.shape[0] * per_image for im in images], per_image

    # ------------------------------------------------------------------ base_mmgpt.py:23-79
    def build_vision_tokenizer(self, model_args, data_args, training_args, tokenizer, vision_config=None):
        vision_tower = build_vision_tower(model_args, config=vision_config)
        if hasattr(vision_tower, "hidden_size"):
            vision_hidden_size = vision_tower.hidden_size
        elif hasattr(vision_tower, "output_dim"):
            vision_hidden_size = vision_tower.output_dim
        else:
            vision_hidden_size = 256
        projector = build_projector(model_args, vision_hidden_size, self.config.hidden_size)
        self.get_model().vision_tower = vision_tower
        self.get_model().projector = projector

        data_args.image_token_len = vision_tower.num_patches
        data_args.image_processor = vision_tower.image_processor
        data_args.use_im_start_end = model_args.use_im_start_end
        self.use_im_start_end = model_args.use_im_start_end
        self.use_beam_search = getattr(data_args, "use_beam_search", False)

        tokenizer.add_tokens([DEFAULT_IM_PATCH_TOKEN], special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        self.im_patch_token = tokenizer.convert_tokens_to_ids([DEFAULT_IM_PATCH_TOKEN])[0]
        if self.use_im_start_end:
            self.num_new_tokens = tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN],
                                                       special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
            self.im_start_token, self.im_end_token = tokenizer.convert_tokens_to_ids(
                [DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN])
            if self.num_new_tokens > 0:
                # new rows start at the mean of the old ones (base_mmgpt.py:69-76)
                with torch.no_grad():
                    inp = self.get_input_embeddings().weight.data
                    outp = self.get_output_embeddings().weight.data
                    inp[-self.num_new_tokens:] = inp[:-self.num_new_tokens].mean(dim=0, keepdim=True)
                    outp[-self.num_new_tokens:] = outp[:-self.num_new_tokens].mean(dim=0, keepdim=True)
                if model_args.freeze_lm_model:
                    self.get_model().orig_embeds_params = self.get_input_embeddings().weight.data.clone()

    # ------------------------------------------------------------------ base_mmgpt.py:82-165
    def _validate_image_grammar(self, ids_host: torch.Tensor, images: List[torch.Tensor], per_image: int):
        """The reference's runtime checks (base_mmgpt.py:116-126) on the host copy of input_ids."""
        for b in range(min(ids_host.shape[0], len(images))):
            row = ids_host[b]
            if (row == self.im_patch_token).sum() == 0:
                continue
            if not self.use_im_start_end:
                raise NotImplementedError
            n_start = int((row == self.im_start_token).sum())
            n_end = int((row == self.im_end_token).sum())
            if n_start != n_end:
                raise ValueError(f"The number of image start tokens ({n_start}) and image end tokens "
                                 f"({n_end}) should be the same.")
            starts = torch.where(row == self.im_start_token)[0].tolist()
            blocks = 0
            for pos, _ in zip(starts, range(images[b].shape[0])):
                end = pos + per_image + 1
                if end >= row.shape[0] or int(row[end]) != self.im_end_token:
                    raise ValueError("The image end token should follow the image start token.")
                if not bool((row[pos + 1:end] == self.im_patch_token).all()):
                    raise ValueError("The image patch tokens between <im_start> and <im_end> must be "
                                     "<im_patch> placeholders.")
                blocks += 1
            # The splice kernel places feature row r at the r-th <im_patch> of the sample.  The reference replaces
            # the span after each of the first n <im_start> tokens and leaves any other <im_patch> as a plain
            # embedding (base_mmgpt.py:119-135): a stray placeholder outside the validated blocks would shift every
            # later feature row by one here, so it is an error instead of a silent misalignment.
            n_patch = int((row == self.im_patch_token).sum())
            if n_patch != blocks * per_image:
                raise ValueError(f"{n_patch} <im_patch> tokens for {blocks} image block(s) of {per_image}: "
                                 "<im_patch> may only appear between <im_start> and <im_end>")

    def prepare_inputs_labels_for_multimodal(self, input_ids, attention_mask, past_key_values, labels, images):
        model = self.get_model()
        vision_tower = getattr(model, "vision_tower", None)
        if vision_tower is None or images is None or input_ids.shape[1] == 1:
            if (past_key_values is not None and vision_tower is not None and images is not None
                    and input_ids.shape[1] == 1 and attention_mask is not None):
                # decode step: the mask is rebuilt to cover cache + 1 (base_mmgpt.py:85-89); works with
                # our cache object where the reference indexes a legacy tuple
                attention_mask = torch.ones((attention_mask.shape[0], past_key_values.get_seq_length() + 1),
                                            dtype=attention_mask.dtype, device=attention_mask.device)
            return input_ids, attention_mask, past_key_values, None, labels

        if len(images) < input_ids.shape[0]:
            # the reference zip()s and silently drops the trailing samples (base_mmgpt.py:105), which then
            # fails inside the decoder on a batch-size mismatch; fail early with a clear message instead
            raise ValueError(f"images has {len(images)} entries for a batch of {input_ids.shape[0]} samples")

        device = model.embed_tokens.device
        feats, rows_per_sample, per_image = self._encode_images_flat(images)
        ids_host = input_ids if not input_ids.is_cuda else input_ids.cpu()
        self._validate_image_grammar(ids_host, images, per_image)
        B = input_ids.shape[0]
        offsets, acc = [], 0
        for r in rows_per_sample[:B]:
            offsets.append(acc)
            acc += r
        feat_offset = torch.tensor(offsets, dtype=torch.int32).to(device, non_blocking=True)
        feat_limit = torch.tensor(rows_per_sample[:B], dtype=torch.int32).to(device, non_blocking=True)
        ids_dev = input_ids.to(device, non_blocking=True).contiguous()
        embeds, _counts = ops.embed_splice(ids_dev, model.embed_tokens, feats, self.im_patch_token,
                                           feat_offset, feat_limit)
        inputs_embeds = embeds.view(B, input_ids.shape[1], -1)
        if getattr(self, "use_beam_search", False):
            inputs_embeds = inputs_embeds.repeat_interleave(5, dim=0)  # base_mmgpt.py:162-163
        return None, attention_mask, past_key_values, inputs_embeds, labels
