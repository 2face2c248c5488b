"""The REAL reference model, assembled from oracle/_ref — TEST INFRASTRUCTURE ONLY.

`import_reference()` imports the reference's own classes (`MMGPTLlamaForCausalLM`, `CLIPVisionTower`,
`ConvProjector`, `MLPProjector`, `get_param_groups` …) from the patched copy `oracle/build_ref.py` makes;
`build_reference_model()` assembles them the way `build_vision_tokenizer` does (base_mmgpt.py:23-79), without
`from_pretrained` (no weights exist offline), on CPU or on a GPU, in fp32 or bf16.  Used by
`oracle/make_golden.py` (pins the restatement), by the `-m gpu` config-parity tests (the reference itself run on
the GPU box in fp32 is the checker at BASELINE.json's full sizes) and by `bench.py`'s reference arm.
Only `tests/`, `bench.py`'s reference legs and `oracle/` itself may import this module.

The third-party modules underneath are the installed transformers 5.5.0 ones (the reference pins 4.31.0);
SURVEY.md §8c's deltas are neutralised here: eager attention unless asked otherwise, `inv_freq` kept in fp32
whatever dtype the model is cast to.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_cached = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "mmgpt")) or os.path.isdir("/root/reference/mmgpt")


def import_reference():
    global _cached
    if _cached is not None:
        return _cached
    from oracle import build_ref
    build_ref.build()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from mmgpt.model.mmgpt.llama_mmgpt import MMGPTConfig, MMGPTLlamaForCausalLM
    from mmgpt.model.projector.conv_projector import ConvProjector
    from mmgpt.model.projector.mlp_projector import MLPProjector
    from mmgpt.model.vision_encoder.clip_encoder import CLIPVisionTower
    from mmgpt.utils import llrd_utils
    _cached = types.SimpleNamespace(MMGPTConfig=MMGPTConfig, MMGPTLlamaForCausalLM=MMGPTLlamaForCausalLM,
                                    ConvProjector=ConvProjector, MLPProjector=MLPProjector,
                                    CLIPVisionTower=CLIPVisionTower, llrd_utils=llrd_utils)
    return _cached


def _fix_buffers(model, cfg, device):
    """Non-persistent buffers are not in the state dict: rebuild them in fp32 (RoPE inv_freq must never be
    rounded to bf16) after a meta-device construction or a dtype cast."""
    hd = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32, device=device) / hd))
    rot = model.model.rotary_emb
    rot.inv_freq = inv
    if hasattr(rot, "original_inv_freq"):
        rot.original_inv_freq = inv.clone()
    emb = model.get_model().vision_tower.vision_tower.vision_model.embeddings
    emb.position_ids = torch.arange(cfg.vit_tokens, device=device).expand(1, -1)


def build_reference_model(cfg, sd=None, device="cpu", dtype=torch.float32, attn: str = "eager",
                          max_position_embeddings: int = 4096, ref=None):
    """Reference MMGPTLlamaForCausalLM + CLIPVisionTower + projector for an OracleConfig.  With `sd` the
    modules are built on the meta device and the tensors of `sd` are adopted (assign=True: no second copy of a
    7B / 13B model); without it they are randomly initialised by the reference's own constructors."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    ref = ref or import_reference()
    lcfg = ref.MMGPTConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                           intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_layers,
                           num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_heads,
                           rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                           max_position_embeddings=max_position_embeddings, tie_word_embeddings=False,
                           hidden_act="silu", attention_bias=False, mlp_bias=False, attn_implementation=attn)
    vcfg = CLIPVisionConfig(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_intermediate,
                            num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                            image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act="quick_gelu",
                            layer_norm_eps=cfg.vit_ln_eps, attn_implementation=attn)
    args = types.SimpleNamespace(freeze_projector=False, model_name_or_path="/nonexistent")
    ctx = torch.device("meta") if sd is not None else torch.device(device)
    with ctx:
        model = ref.MMGPTLlamaForCausalLM(lcfg)
        vt = ref.CLIPVisionTower.__new__(ref.CLIPVisionTower)  # __init__ calls from_pretrained (no weights offline)
        torch.nn.Module.__init__(vt)
        vt.select_layer, vt.select_feature = cfg.select_layer, cfg.select_feature
        vt.freeze_vision_tower, vt.conv_stride = False, (cfg.conv_stride if cfg.projector == "conv" else 1)
        vt.vision_tower = CLIPVisionModel(vcfg)
        if cfg.projector == "conv":
            proj = ref.ConvProjector(args, cfg.vit_hidden, cfg.hidden_size, conv_stride=cfg.conv_stride)
        else:
            proj = ref.MLPProjector(args, cfg.vit_hidden, cfg.hidden_size)
    model.get_model().vision_tower = vt
    model.get_model().projector = proj
    model.use_im_start_end, model.use_beam_search = True, False
    model.im_patch_token, model.im_start_token, model.im_end_token = \
        cfg.im_patch_token, cfg.im_start_token, cfg.im_end_token
    if sd is not None:
        sd = {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False, assign=True)
        assert not unexpected, unexpected
        # the only keys we do not provide are non-persistent / derived buffers
        assert all("position_ids" in k or "inv_freq" in k for k in missing), missing
    else:
        model.to(dtype=dtype)
    _fix_buffers(model, cfg, device)
    model.eval()
    return model


def param_groups_reference(named_params, lr: float, weight_decay: float, llrd: bool = True, llm_llrd: bool = False,
                           ref=None):
    """The reference's own `get_param_groups` (mmgpt/utils/llrd_utils.py:25-79), called the way
    `MMGPTTrainer.create_optimizer` calls it (mmgpt/engine/train/trainer.py:53-66), on real nn.Parameters."""
    ref = ref or import_reference()
    U = ref.llrd_utils
    fn = U.vit_lr_scale_func if llrd else (U.llm_lr_scale_func if llm_llrd else None)
    model = types.SimpleNamespace(named_parameters=lambda: list(named_params))
    return U.get_param_groups(model, None, fn, lr, weight_decay)
