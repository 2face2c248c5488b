"""build_model_tokenizer — mirror of mmgpt/model/builder.py:18-165 for the Llama / Vicuna branch.

Same signature and side effects: loads the config, the language model and the tokenizer from
`model_args.model_name_or_path`, sets pad = unk, resizes the embeddings, applies the freezing policy,
calls `build_vision_tokenizer` (vision tower + projector + the three image tokens) and moves the model to
bf16 on the training device.  Not carried over: the Baichuan / Phi / OPT branches (other model families,
SURVEY.md §2.1 #11) and LoRA (`training_args.lora_enable` raises NotImplementedError).

`tokenizer=` lets a caller inject an already constructed tokenizer (used offline, where no
sentencepiece model file is available); by default it is loaded exactly like the reference does.
"""
from __future__ import annotations

import torch

from .mmgpt.llama_mmgpt import MMGPTLlamaForCausalLM


def build_model_tokenizer(model_args, data_args, training_args, tokenizer=None, vision_config=None):
    if getattr(training_args, "lora_enable", False):
        raise NotImplementedError("LoRA fine-tuning is outside the B200 hot path")
    name = model_args.model_name_or_path
    if any(k in name for k in ("baichuan", "phi", "opt")):
        raise NotImplementedError(f"only the Llama / Vicuna backbone is built (got {name})")
    model = MMGPTLlamaForCausalLM.from_pretrained(name)

    if tokenizer is None:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(name, model_max_length=training_args.model_max_length,
                                                  padding_side="right", use_fast=False)
    # NOTE (builder.py:94-97): ids beyond the vocab would fault in training, so pad reuses <unk>
    tokenizer.pad_token = tokenizer.unk_token
    model.resize_token_embeddings(len(tokenizer))

    dtype = torch.float32
    if getattr(training_args, "bf16", True):
        dtype = torch.bfloat16
    if getattr(training_args, "fp16", False):
        raise NotImplementedError("the kernels compute in bf16 (pretrain.sh:13 uses --bf16 True)")
    if dtype != torch.bfloat16:
        raise NotImplementedError("merlin_b200 runs in bf16 only")

    if model_args.freeze_lm_model:
        for p in model.parameters():
            p.requires_grad_(False)

    if model_args.vision_tower is not None:
        model.build_vision_tokenizer(model_args=model_args, data_args=data_args, training_args=training_args,
                                     tokenizer=tokenizer, vision_config=vision_config)
        tower, projector = model.get_model().vision_tower, model.get_model().projector
        if model_args.freeze_lm_model:
            # builder.py:139-160: the tower trains unless frozen (its last layer / post_layernorm never get a
            # gradient), the projector unless frozen, the input embeddings when new tokens are learned
            for p in tower.parameters():
                p.requires_grad_(not model_args.freeze_vision_tower)
            last = tower.config.num_hidden_layers - 1
            for lst in (tower.qkv_w, tower.qkv_b, tower.out_w, tower.out_b, tower.ln1_w, tower.ln1_b, tower.ln2_w,
                        tower.ln2_b, tower.fc1_w, tower.fc1_b, tower.fc2_w, tower.fc2_b):
                lst[last].requires_grad_(False)
            tower.post_ln_w.requires_grad_(False)
            tower.post_ln_b.requires_grad_(False)
            for p in projector.parameters():
                p.requires_grad_(not model_args.freeze_projector)
            if model_args.use_im_start_end:
                model.model.embed_tokens.requires_grad_(True)
                model.lm_head_w.requires_grad_(False)

    model.to(dtype=dtype, device=getattr(training_args, "device", "cuda"))
    return model, tokenizer, data_args
