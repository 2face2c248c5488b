"""Shared test helpers: golden fixtures, oracle <-> product model construction."""
import hashlib
import os
from types import SimpleNamespace

import torch

from oracle import merlin_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    cfg = O.OracleConfig(**g["cfg"])
    return g, cfg


def weights_sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def build_product_model(cfg: O.OracleConfig, sd, device="cuda", max_position_embeddings=4096):
    """Assemble merlin_b200's MMGPTLlamaForCausalLM for an oracle config and load the oracle's
    reference-format state dict through the public load_state_dict (strict).  When the state dict already
    lives on a GPU (7B / 13B config cases) the model is constructed there directly in bf16."""
    import contextlib
    from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MMGPTConfig, MMGPTLlamaForCausalLM
    from merlin_b200.mmgpt.model.projector.builder import build_projector
    from merlin_b200.mmgpt.model.vision_encoder.builder import build_vision_tower

    lcfg = MMGPTConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                       intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_layers,
                       num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_heads,
                       rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                       max_position_embeddings=max_position_embeddings, tie_word_embeddings=False)
    on_gpu = next(iter(sd.values())).is_cuda
    prev = torch.get_default_dtype()
    if on_gpu:
        torch.set_default_dtype(torch.bfloat16)
    try:
        with (torch.device(next(iter(sd.values())).device) if on_gpu else contextlib.nullcontext()):
            model = _assemble(cfg, lcfg, MMGPTLlamaForCausalLM, build_vision_tower, build_projector)
    finally:
        torch.set_default_dtype(prev)
    model.load_state_dict(sd, strict=True)
    if device is not None:
        model.to(device=device, dtype=torch.bfloat16)
    model.eval()
    # inference by default: with trainable parameters and grad mode on, forward() takes the autograd path (like any
    # HF model); the training tests switch requires_grad on for what they train
    model.requires_grad_(False)
    return model


def _assemble(cfg, lcfg, MMGPTLlamaForCausalLM, build_vision_tower, build_projector):
    model = MMGPTLlamaForCausalLM(lcfg)
    args = SimpleNamespace(vision_tower="clip-synthetic", vision_select_layer=cfg.select_layer,
                           vision_select_feature=cfg.select_feature, freeze_vision_tower=True,
                           freeze_projector=True, conv_stride=cfg.conv_stride if cfg.projector == "conv" else 1,
                           projector=cfg.projector, model_name_or_path=None)
    vcfg = SimpleNamespace(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_intermediate,
                           num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                           image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act="quick_gelu",
                           layer_norm_eps=cfg.vit_ln_eps, num_channels=3)
    model.get_model().vision_tower = build_vision_tower(args, config=vcfg)
    model.get_model().projector = build_projector(args, cfg.vit_hidden, cfg.hidden_size)
    model.im_patch_token, model.im_start_token, model.im_end_token = \
        cfg.im_patch_token, cfg.im_start_token, cfg.im_end_token
    model.use_im_start_end, model.use_beam_search = cfg.use_im_start_end, cfg.use_beam_search
    return model


class FakeTokenizer:
    """Duck-typed stand-in for the sentencepiece tokenizer (no tokenizer files exist offline): the
    subset of the HF tokenizer API that build_model_tokenizer / build_vision_tokenizer touch."""

    def __init__(self, base_vocab: int, extra=()):
        self.vocab = {f"tok{i}": i for i in range(base_vocab)}
        self.vocab["<unk>"] = 0
        self.unk_token, self.pad_token = "<unk>", None
        self.unk_token_id = 0
        self.model_max_length = 2048
        for t in extra:
            self.add_tokens([t], special_tokens=True)

    def __len__(self):
        return len(set(self.vocab.values()))

    def add_tokens(self, tokens, special_tokens=False):
        added = 0
        for t in tokens:
            if t not in self.vocab:
                self.vocab[t] = len(self)
                added += 1
        return added

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab[t] for t in tokens]
