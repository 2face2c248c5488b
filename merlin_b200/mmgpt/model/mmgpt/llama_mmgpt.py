"""MMGPTLlamaForCausalLM (alias MerlinForCausalLM) on B200 kernels.

Mirror of mmgpt/model/mmgpt/llama_mmgpt.py:27-138 — `MMGPTConfig`, `MMGPTLlamaModel`,
`MMGPTLlamaForCausalLM.forward(input_ids, attention_mask, past_key_values, inputs_embeds, labels,
use_cache, output_attentions, output_hidden_states, images, return_dict)`,
`prepare_inputs_for_generation`, and a `generate(input_ids, images=..., ...)` with the call shape the
reference's eval/serve code uses (eval_mmvet.py:101-120, model_worker.py:169-180).

The decoder stack (HF LlamaModel + the flash-attn monkey patch) runs per layer as
  rmsnorm -> [fused QKV tcgen05 GEMM with RoPE epilogue] -> causal flash-attention (tcgen05/TMEM)
  -> [o_proj GEMM + residual epilogue] -> rmsnorm -> [gate|up GEMM + SwiGLU epilogue]
  -> [down GEMM + residual epilogue]
through libmerlin_b200.so; there is no PyTorch fallback.  Weights are stored packed for the kernels
(q|k|v fused, gate/up interleaved per 128 rows); state_dict()/load_state_dict() speak the reference's
key names.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import List, Optional

import torch
import torch.nn as nn

from .... import decode as _decode
from .... import ops
from ....lib import EPI_RESIDUAL, EPI_SWIGLU
from ...utils.constants import IGNORE_INDEX
from .base_mmgpt import MMGPTMetaForCausalLM

from transformers import AutoConfig, AutoModelForCausalLM, GenerationMixin, LlamaConfig, PreTrainedModel
from transformers.modeling_outputs import CausalLMOutputWithPast


class MMGPTConfig(LlamaConfig):
    """llama_mmgpt.py:27-28."""
    model_type = "mmgpt"


def vicuna_7b_config(**kw) -> "MMGPTConfig":
    """Vicuna-7B-v1.5 dims; vocab 32003 after the three image tokens."""
    base = dict(vocab_size=32003, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                num_attention_heads=32, num_key_value_heads=32, rms_norm_eps=1e-5, rope_theta=10000.0,
                max_position_embeddings=4096, tie_word_embeddings=False)
    base.update(kw)
    return MMGPTConfig(**base)


def vicuna_13b_config(**kw) -> "MMGPTConfig":
    base = dict(vocab_size=32003, hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                num_attention_heads=40, num_key_value_heads=40, rms_norm_eps=1e-5, rope_theta=10000.0,
                max_position_embeddings=4096, tie_word_embeddings=False)
    base.update(kw)
    return MMGPTConfig(**base)


class MerlinKVCache:
    """Per-layer key/value store for generate(): kv[l] is [B, capacity/32, H, 32, 256] bf16 — blocks of 32
    tokens, head-major inside a block, row (h, t) = [k_t^h (128, already rotated) | v_t^h (128)] — so a
    (head, 32-key) tile is 16 KB contiguous for the decode kernels' bulk copies while a token range stays
    one dense region.  Stands in for the legacy tuple cache the reference indexes (base_mmgpt.py:88) and
    for HF 5.x Cache objects (`get_seq_length`)."""

    def __init__(self, num_layers: int, batch: int, capacity: int, hidden: int, device):
        capacity = (capacity + 31) // 32 * 32
        self.kv = [torch.empty((batch, capacity // 32, hidden // 128, 32, 256), dtype=torch.bfloat16, device=device)
                   for _ in range(num_layers)]
        self.capacity = capacity
        self.seq_len = 0

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.seq_len

    def __len__(self):
        return len(self.kv)

    def __bool__(self):  # an allocated-but-empty cache must read as "no past" 
        return self.seq_len > 0


class _WeightHolder(nn.Module):
    """Gives packed parameters the `.weight` attribute HF-style callers expect
    (get_input_embeddings().weight, get_output_embeddings().weight)."""

    def __init__(self, owner, name):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        self._name = name

    @property
    def weight(self):
        return getattr(self._owner, self._name)


class MMGPTLlamaModel(nn.Module):
    """Decoder stack (reference: MMGPTLlamaModel(LlamaModel), llama_mmgpt.py:31-35)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        D, FF, L, V = config.hidden_size, config.intermediate_size, config.num_hidden_layers, config.vocab_size
        H = config.num_attention_heads
        if D // H != 128 or getattr(config, "num_key_value_heads", H) != H:
            raise ValueError("MMGPTLlamaModel: kernels need head_dim 128 and MHA (Vicuna-7B/13B)")
        if FF % 128 != 0 or D % 8 != 0:
            raise ValueError("MMGPTLlamaModel: intermediate_size must be a multiple of 128")
        std = getattr(config, "initializer_range", 0.02)

        def W(*shape):
            return nn.Parameter(torch.empty(*shape).normal_(0.0, std))

        self.embed_tokens = W(V, D)
        self.qkv_w = nn.ParameterList([W(3 * D, D) for _ in range(L)])
        self.o_w = nn.ParameterList([W(D, D) for _ in range(L)])
        self.gate_up_w = nn.ParameterList([W(2 * FF, D) for _ in range(L)])   # interleaved per 128 rows
        self.down_w = nn.ParameterList([W(D, FF) for _ in range(L)])
        self.ln1_w = nn.ParameterList([nn.Parameter(torch.ones(D)) for _ in range(L)])
        self.ln2_w = nn.ParameterList([nn.Parameter(torch.ones(D)) for _ in range(L)])
        self.norm_w = nn.Parameter(torch.ones(D))
        self._rope = None  # (cos, sin) fp32 tables, built lazily on the parameter device
        self.gradient_checkpointing = False  # set by gradient_checkpointing_enable(): per-layer activation recompute

    def rope_tables(self, need: int):
        if self._rope is None or self._rope[0].shape[0] < need or self._rope[0].device != self.embed_tokens.device:
            size = max(need, getattr(self.config, "max_position_embeddings", 4096))
            self._rope = ops.rope_tables(size, 128, float(getattr(self.config, "rope_theta", 10000.0)),
                                         device=self.embed_tokens.device)
        return self._rope

    def forward_layers(self, x: torch.Tensor, batch: int, seq: int, seqlens: Optional[torch.Tensor],
                       cache: Optional[MerlinKVCache] = None, collect_hidden: bool = False):
        """x [batch*seq, D] bf16 (modified in place) -> final-normed hidden [batch*seq, D]."""
        cfg = self.config
        D, H = cfg.hidden_size, cfg.num_attention_heads
        cos, sin = self.rope_tables(seq)
        xn = torch.empty_like(x)
        hidden = [x.clone()] if collect_hidden else None
        if not collect_hidden and not ops.PROFILE.enabled:
            # one C call per layer (merlin_decoder_layer_forward): same kernels, sequenced in the library
            ws = ops.decoder_layer_workspace(batch * seq, D, cfg.intermediate_size, x.device)
            for i in range(cfg.num_hidden_layers):
                ops.decoder_layer_forward(x, ws, self.qkv_w[i], self.o_w[i], self.gate_up_w[i], self.down_w[i],
                                          self.ln1_w[i], self.ln2_w[i], cos, sin, seqlens,
                                          None if cache is None else cache.kv[i], batch, seq, H,
                                          cfg.intermediate_size, cfg.rms_norm_eps)
            if cache is not None:
                cache.seq_len = seq
            return ops.rmsnorm(x, self.norm_w, cfg.rms_norm_eps, out=xn), None
        for i in range(cfg.num_hidden_layers):
            ops.rmsnorm(x, self.ln1_w[i], cfg.rms_norm_eps, out=xn)
            fuse_kv = cache is not None and (batch == 1 or seq % 128 == 0)
            qkv = ops.gemm_qkv_rope(xn, self.qkv_w[i], cos, sin, seq, kv_cache=cache.kv[i] if fuse_kv else None,
                                    batch=batch, heads=H)
            if cache is not None and not fuse_kv:
                ops.kv_append(qkv, cache.kv[i], batch, seq, D, H)
            att = ops.flash_attn(qkv, batch, seq, H, 128, causal=True, seqlens=seqlens)
            ops.gemm(att, self.o_w[i], residual=x, out=x, epilogue=EPI_RESIDUAL)
            ops.rmsnorm(x, self.ln2_w[i], cfg.rms_norm_eps, out=xn)
            mid = ops.gemm(xn, self.gate_up_w[i], epilogue=EPI_SWIGLU)
            ops.gemm(mid, self.down_w[i], residual=x, out=x, epilogue=EPI_RESIDUAL)
            if collect_hidden:
                hidden.append(x.clone())
        if cache is not None:
            cache.seq_len = seq
        out = ops.rmsnorm(x, self.norm_w, cfg.rms_norm_eps, out=xn)
        return out, hidden


class MMGPTLlamaForCausalLM(PreTrainedModel, GenerationMixin, MMGPTMetaForCausalLM):
    """HF `PreTrainedModel` + `GenerationMixin` subclass registered as model_type "mmgpt", like the reference's
    `MMGPTLlamaForCausalLM(LlamaForCausalLM, MMGPTMetaForCausalLM)` (llama_mmgpt.py:38,137-138): `AutoConfig` /
    `AutoModelForCausalLM` resolve it, `Trainer`-style `model(**batch).loss.backward()` trains it (autograd
    Functions over the library's backward entry points, merlin_b200/autograd.py) and `generate(**hf_kwargs)`
    accepts the HF generation arguments.  The decoder itself is not HF's LlamaModel: weights live packed in
    `MMGPTLlamaModel` and every FLOP runs in libmerlin_b200.so."""
    config_class = MMGPTConfig
    base_model_prefix = "model"
    main_input_name = "input_ids"
    supports_gradient_checkpointing = True
    _supports_flash_attn = True
    _supports_sdpa = True
    _supports_flex_attn = True
    _supports_attention_backend = True
    _no_split_modules = []
    _tied_weights_keys = None

    def __init__(self, config):
        super().__init__(config)
        self.model = MMGPTLlamaModel(config)
        std = getattr(config, "initializer_range", 0.02)
        self.lm_head_w = nn.Parameter(torch.empty(config.vocab_size, config.hidden_size).normal_(0.0, std))
        self.use_im_start_end = True
        self.use_beam_search = False
        self.im_patch_token, self.im_start_token, self.im_end_token = 32000, 32001, 32002
        self.num_new_tokens = 0
        self._decoder = None
        self.post_init()

    def _init_weights(self, module):
        """Parameters are initialised by the constructors above (N(0, initializer_range), norm weights 1)."""
        return

    def tie_weights(self, *args, **kwargs):
        """lm_head and embed_tokens are never tied (Vicuna: tie_word_embeddings=False)."""
        return

    # ------------------------------------------------------------------ HF-style accessors
    def get_model(self):
        return self.model

    def get_input_embeddings(self):
        return _WeightHolder(self.model, "embed_tokens")

    def get_output_embeddings(self):
        return _WeightHolder(self, "lm_head_w")

    @property
    def device(self):
        return self.lm_head_w.device

    @property
    def dtype(self):
        return self.lm_head_w.dtype

    def resize_token_embeddings(self, new_num_tokens: int):
        old = self.config.vocab_size
        if new_num_tokens == old:
            return self.get_input_embeddings()
        for owner, name in ((self.model, "embed_tokens"), (self, "lm_head_w")):
            w = getattr(owner, name)
            new = torch.empty(new_num_tokens, w.shape[1], dtype=w.dtype, device=w.device).normal_(
                0.0, getattr(self.config, "initializer_range", 0.02))
            n = min(old, new_num_tokens)
            new[:n] = w.data[:n]
            setattr(owner, name, nn.Parameter(new, requires_grad=w.requires_grad))
        self.config.vocab_size = new_num_tokens
        return self.get_input_embeddings()

    # ------------------------------------------------------------------ reference-format state dict
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        cfg = self.config
        D, FF = cfg.hidden_size, cfg.intermediate_size
        out = destination if destination is not None else OrderedDict()

        def put(k, v):
            out[prefix + k] = v if keep_vars else v.detach()

        m = self.model
        put("model.embed_tokens.weight", m.embed_tokens)
        for i in range(cfg.num_hidden_layers):
            p = f"model.layers.{i}."
            for n, blk in (("q", 0), ("k", 1), ("v", 2)):
                put(p + f"self_attn.{n}_proj.weight", m.qkv_w[i][blk * D:(blk + 1) * D])
            put(p + "self_attn.o_proj.weight", m.o_w[i])
            gu = m.gate_up_w[i].view(FF // 128, 2, 128, D)
            # strided views of the interleaved storage; reshape() materialises a copy for export and the
            # load path below writes through the 4-D view
            put(p + "mlp.gate_proj.weight", gu[:, 0].reshape(FF, D))
            put(p + "mlp.up_proj.weight", gu[:, 1].reshape(FF, D))
            put(p + "mlp.down_proj.weight", m.down_w[i])
            put(p + "input_layernorm.weight", m.ln1_w[i])
            put(p + "post_attention_layernorm.weight", m.ln2_w[i])
        put("model.norm.weight", m.norm_w)
        vt = getattr(m, "vision_tower", None)
        if vt is not None:
            vt.state_dict(destination=out, prefix=prefix + "model.vision_tower.", keep_vars=keep_vars)
        pj = getattr(m, "projector", None)
        if pj is not None:
            pj.state_dict(destination=out, prefix=prefix + "model.projector.", keep_vars=keep_vars)
        put("lm_head.weight", self.lm_head_w)
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        cfg = self.config
        D, FF = cfg.hidden_size, cfg.intermediate_size
        own = self.state_dict(keep_vars=True)
        skip = ("rotary_emb.inv_freq", "position_ids")
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own and not k.endswith(skip)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:4]} ({len(missing)}), "
                               f"unexpected {unexpected[:4]} ({len(unexpected)})")
        m = self.model
        with torch.no_grad():
            for k, dst in own.items():
                if k not in state_dict:
                    continue
                src = state_dict[k]
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(src.shape)} vs {tuple(dst.shape)}")
                if k.endswith("mlp.gate_proj.weight") or k.endswith("mlp.up_proj.weight"):
                    i = int(k.split("layers.")[1].split(".")[0])
                    which = 0 if "gate_proj" in k else 1
                    m.gate_up_w[i].view(FF // 128, 2, 128, D)[:, which].copy_(src.view(FF // 128, 128, D))
                else:
                    dst.copy_(src)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def save_pretrained(self, save_directory: str, state_dict=None, safe_serialization: bool = False, **kwargs):
        """Writes `pytorch_model.bin` (reference key names) + `config.json` — the files the reference's own
        loaders read back (clip_encoder.py:26-62, base_projector.py:12-48, from_pretrained).  `state_dict`
        (HF Trainer._save passes the CPU copy made by trainer.py:29-43) overrides the live weights;
        safe_serialization=True writes `model.safetensors` instead."""
        os.makedirs(save_directory, exist_ok=True)
        sd = state_dict if state_dict is not None else self.state_dict()
        sd = {k: v.detach().cpu().contiguous() for k, v in sd.items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file({k: v.clone() for k, v in sd.items()}, os.path.join(save_directory, "model.safetensors"),
                      metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, "pytorch_model.bin"))
        self.config.save_pretrained(save_directory)
        if getattr(self, "generation_config", None) is not None:
            try:
                self.generation_config.save_pretrained(save_directory)
            except Exception:  # an invalid user-edited generation config must not lose the weights
                pass

    @staticmethod
    def _read_checkpoint(path: str):
        """State dict from pytorch_model.bin, its sharded index, or safetensors files in `path`."""
        idx = os.path.join(path, "pytorch_model.bin.index.json")
        single = os.path.join(path, "pytorch_model.bin")
        sd = {}
        if os.path.exists(idx):
            for fname in sorted(set(json.load(open(idx))["weight_map"].values())):
                sd.update(torch.load(os.path.join(path, fname), map_location="cpu"))
        elif os.path.exists(single):
            sd = torch.load(single, map_location="cpu")
        else:
            st = [f for f in sorted(os.listdir(path)) if f.endswith(".safetensors")] if os.path.isdir(path) else []
            if st:
                from safetensors.torch import load_file
                for f in st:
                    sd.update(load_file(os.path.join(path, f)))
        return sd

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *model_args, config=None, torch_dtype=None, dtype=None,
                        **_):
        """Local-directory loader with the call shape of builder.py:70 (`from_pretrained(path, cache_dir=...)`):
        config.json + pytorch_model.bin[.index.json] / *.safetensors in reference key names (a plain Vicuna
        checkpoint or a saved mmgpt model)."""
        path = pretrained_model_name_or_path
        if config is None:
            raw = json.load(open(os.path.join(path, "config.json")))
            raw = {k: v for k, v in raw.items() if k not in ("model_type", "architectures", "transformers_version")}
            config = MMGPTConfig(**raw)
        model = cls(config)
        sd = cls._read_checkpoint(path)
        if sd:
            llm = {k: v for k, v in sd.items() if "vision_tower" not in k and "projector" not in k}
            vocab = llm["model.embed_tokens.weight"].shape[0]
            if vocab != config.vocab_size:           # checkpoint saved after the image tokens were added
                model.resize_token_embeddings(vocab)
            model.load_state_dict(llm, strict=False)
        want = dtype if dtype is not None else torch_dtype
        if isinstance(want, torch.dtype):
            model.to(want)
        return model

    # ------------------------------------------------------------------ forward (llama_mmgpt.py:53-112)
    def forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                past_key_values=None, inputs_embeds: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None, use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None, output_hidden_states: Optional[bool] = None,
                images: Optional[List[torch.Tensor]] = None, return_dict: Optional[bool] = None):
        if output_attentions:
            raise AssertionError("output_attentions is not supported")  # monkey_patch.py:61
        if self.lm_head_w.dtype != torch.bfloat16 or not self.lm_head_w.is_cuda:
            raise RuntimeError("MMGPTLlamaForCausalLM runs on CUDA in bfloat16: call "
                               "model.to(dtype=torch.bfloat16, device='cuda') first (builder.py:163)")
        return_dict = True if return_dict is None else return_dict
        if past_key_values is not None and not isinstance(past_key_values, MerlinKVCache):
            # HF generate() pre-allocates an EMPTY DynamicCache: that is "no past, please
            # cache" — the MerlinKVCache of this forward's output replaces it in HF's model_kwargs
            if hasattr(past_key_values, "get_seq_length") and past_key_values.get_seq_length() == 0:
                past_key_values, use_cache = None, True
            else:
                raise TypeError("past_key_values must be the MerlinKVCache returned by a previous forward")
        if (torch.is_grad_enabled() and past_key_values is None and not use_cache
                and self._wants_grad(inputs_embeds)):
            return self._forward_autograd(input_ids, attention_mask, inputs_embeds, labels, images,
                                          bool(output_hidden_states), return_dict)
        if past_key_values is not None and not past_key_values:
            past_key_values_in, past_for_mm = past_key_values, None
        else:
            past_key_values_in = past_for_mm = past_key_values

        if inputs_embeds is None:
            input_ids, attention_mask, _, inputs_embeds, labels = self.prepare_inputs_labels_for_multimodal(
                input_ids, attention_mask, past_for_mm, labels, images)
        device = self.device
        if inputs_embeds is None:  # text-only or decode step: plain embedding gather
            ids_dev = input_ids.to(device).contiguous()
            flat, _ = ops.embed_splice(ids_dev, self.model.embed_tokens, None, -1)
            inputs_embeds = flat.view(ids_dev.shape[0], ids_dev.shape[1], -1)
        B, S, D = inputs_embeds.shape

        if past_for_mm is not None:  # ---- decode step(s) against the cache (HF eager path, K12)
            if S != 1:
                raise NotImplementedError("cached forward supports one new token per step")
            logits = self._get_decoder(past_for_mm).step_logits(inputs_embeds.reshape(B, D)).view(B, 1, -1)
            out = CausalLMOutputWithPast(loss=None, logits=logits, past_key_values=past_for_mm)
            return out if return_dict else (logits, past_for_mm)

        seqlens = None
        if attention_mask is not None:
            lens = attention_mask.sum(dim=-1)
            if bool((lens != S).any()):
                seqlens = lens.to(device=device, dtype=torch.int32)
        cache = None
        if use_cache:
            cache = past_key_values_in if past_key_values_in is not None else MerlinKVCache(
                self.config.num_hidden_layers, B, max(S + 1024, getattr(self, "_cache_capacity", 0)), D, device)
        x = inputs_embeds.reshape(B * S, D)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        x = x.contiguous()
        lab = labels.to(device).contiguous() if labels is not None else None
        logits2d, hidden, all_hidden, loss = self._decoder_core(x, B, S, seqlens, cache, bool(output_hidden_states), lab)
        logits = logits2d.view(B, S, -1) if logits2d.is_contiguous() else logits2d.unflatten(0, (B, S))
        hs = None
        if output_hidden_states:
            hs = tuple(h.view(B, S, D) for h in all_hidden[:-1]) + (hidden.view(B, S, D),)
        if not return_dict:
            out = (logits, cache)
            return (loss,) + out if loss is not None else out
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=cache, hidden_states=hs,
                                      attentions=None)

    # ------------------------------------------------------------------ training forward (autograd)
    def _wants_grad(self, inputs_embeds) -> bool:
        if inputs_embeds is not None and inputs_embeds.requires_grad:
            return True
        return any(p.requires_grad for p in self.parameters())

    def _forward_autograd(self, input_ids, attention_mask, inputs_embeds, labels, images, output_hidden_states,
                          return_dict):
        """forward() with gradients: the same kernels as the inference path, sequenced through the
        torch.autograd.Functions of merlin_b200/autograd.py (one Function per decoder layer, per tower, per
        projector, for the splice and for lm_head + CE), so `out.loss.backward()` fills `param.grad` — the contract
        of HF Trainer.training_step over the reference model (mmgpt/engine/train/train.py:49-59).  With
        `gradient_checkpointing_enable()` every decoder layer keeps only its input and is re-run in backward
        (the reference trains with --gradient_checkpointing True, pretrain.sh:35)."""
        from .... import autograd as A
        m, cfg = self.model, self.config
        device = self.device
        D, H, FF = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        tower, proj = getattr(m, "vision_tower", None), getattr(m, "projector", None)
        if inputs_embeds is None:
            ids_dev = input_ids.to(device).contiguous()
            B, S = ids_dev.shape
            feats = feat_offset = feat_limit = None
            if tower is not None and images is not None and S != 1:
                if len(images) < B:
                    raise ValueError(f"images has {len(images)} entries for a batch of {B} samples")
                concat = torch.cat([im for im in images], dim=0)
                n_img = concat.shape[0]
                tower_params = [p for _, p in tower.named_parameters_reference()]
                if tower_params:
                    hidden_v = A.VisionTowerFn.apply(tower, concat, *tower_params)
                else:
                    with torch.no_grad():
                        hidden_v = tower.encode(concat)
                grid = tower.config.image_size // tower.config.patch_size
                if hasattr(proj, "project") and tower.select_feature == "patch":
                    feats = A.ConvProjectorFn.apply(hidden_v, proj.weight, proj.bias, n_img, grid, proj.in_channels,
                                                    proj.conv_stride, tower.tokens, 1)
                else:
                    rows = tower.feature_select(hidden_v, n_img).reshape(-1, hidden_v.shape[-1])
                    feats = A.LinearFn.apply(rows, proj.weight, proj.bias)
                feats = feats.reshape(-1, feats.shape[-1])
                per_image = feats.shape[0] // n_img
                ids_host = input_ids if not input_ids.is_cuda else input_ids.cpu()
                self._validate_image_grammar(ids_host, images, per_image)
                rows_per_sample = [im.shape[0] * per_image for im in images][:B]
                offs = [sum(rows_per_sample[:b]) for b in range(B)]
                feat_offset = torch.tensor(offs, dtype=torch.int32).to(device)
                feat_limit = torch.tensor(rows_per_sample, dtype=torch.int32).to(device)
                if not feats.is_contiguous():
                    feats = feats.contiguous()
            x = A.EmbedSpliceFn.apply(m.embed_tokens, feats, ids_dev, self.im_patch_token if feats is not None else -1,
                                      feat_offset, feat_limit)
        else:
            B, S, _ = inputs_embeds.shape
            x = inputs_embeds.reshape(B * S, D).to(torch.bfloat16)
        if getattr(self, "use_beam_search", False):
            raise NotImplementedError("use_beam_search is a generation-time flag (base_mmgpt.py:162-163)")
        seqlens = None
        if attention_mask is not None:
            lens = attention_mask.sum(dim=-1)
            if bool((lens != S).any()):
                seqlens = lens.to(device=device, dtype=torch.int32)
        cos, sin = m.rope_tables(S)
        recompute = bool(m.gradient_checkpointing and self.training)
        hs = [x] if output_hidden_states else None
        for i in range(cfg.num_hidden_layers):
            x = A.DecoderLayerFn.apply(x, m.qkv_w[i], m.o_w[i], m.gate_up_w[i], m.down_w[i], m.ln1_w[i], m.ln2_w[i],
                                       cos, sin, seqlens, B, S, H, FF, cfg.rms_norm_eps, recompute)
            if output_hidden_states:
                hs.append(x)
        hn = A.RMSNormFn.apply(x, m.norm_w, cfg.rms_norm_eps)
        loss = None
        if labels is not None:
            lab = labels.to(device).contiguous()
            logits2d, loss = A.LMHeadCEFn.apply(hn, self.lm_head_w, lab, B, S, IGNORE_INDEX)
        else:
            logits2d = A.LinearFn.apply(hn, self.lm_head_w, None)
        logits = logits2d.view(B, S, -1) if logits2d.is_contiguous() else logits2d.unflatten(0, (B, S))
        hidden_states = None
        if output_hidden_states:
            hidden_states = tuple(h.view(B, S, D) for h in hs[:-1]) + (hn.view(B, S, D),)
        if not return_dict:
            return (loss, logits, None) if loss is not None else (logits, None)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=hidden_states,
                                      attentions=None)

    def named_parameters_reference(self, trainable_only: bool = True):
        """(reference-style name, packed parameter) pairs in the order their gradients become final
        during backward: lm_head, final norm, decoder layers last-to-first, embeddings, projector.
        The name is the representative reference key used for the optimizer grouping rules
        (llrd_utils.py): fused q|k|v / gate|up tensors share the group of their members."""
        m = self.model
        out = [("lm_head.weight", self.lm_head_w), ("model.norm.weight", m.norm_w)]
        for i in reversed(range(self.config.num_hidden_layers)):
            p = f"model.layers.{i}."
            out += [(p + "mlp.down_proj.weight", m.down_w[i]), (p + "mlp.gate_up_proj.weight", m.gate_up_w[i]),
                    (p + "post_attention_layernorm.weight", m.ln2_w[i]),
                    (p + "self_attn.o_proj.weight", m.o_w[i]), (p + "self_attn.qkv_proj.weight", m.qkv_w[i]),
                    (p + "input_layernorm.weight", m.ln1_w[i])]
        out.append(("model.embed_tokens.weight", m.embed_tokens))
        pj = getattr(m, "projector", None)
        if pj is not None:
            out += [("model.projector.projector.weight", pj.weight), ("model.projector.projector.bias", pj.bias)]
        vt = getattr(m, "vision_tower", None)
        if vt is not None:
            out += vt.named_parameters_reference()
        if trainable_only:
            out = [(n, p) for n, p in out if p.requires_grad]
        return out

    def _decoder_core(self, x, B, S, seqlens, cache, collect_hidden=False, labels=None):
        """Device-only tail of forward: decoder stack + lm_head (+ shifted CE when labels are given) — kernel
        launches only, graph-capturable.  lm_head and the loss are ONE library call (merlin_lmhead_ce_forward)."""
        hidden, all_hidden = self.model.forward_layers(x, B, S, seqlens, cache, collect_hidden)
        logits2d, loss, _ = ops.lmhead_ce_forward(hidden, self.lm_head_w, labels, B, S, IGNORE_INDEX)
        return logits2d, hidden, all_hidden, loss          # logits [B*S, V] on a 16-byte aligned pitch

    def _prefill_core(self, ids_dev, images_bf16, feat_offset, feat_limit, cache):
        """Device-only multimodal prefill for one pre-validated batch: ViT -> projector -> splice ->
        decoder -> lm_head.  No host reads, so the whole thing can be captured in a CUDA graph."""
        m = self.model
        B, S = ids_dev.shape
        tower, proj = m.vision_tower, m.projector
        n = images_bf16.shape[0]
        hidden = tower.encode(images_bf16)
        if hasattr(proj, "project") and tower.select_feature == "patch":
            grid = tower.config.image_size // tower.config.patch_size
            feats = proj.project(hidden, n, grid, tower.tokens, 1)
        else:
            feats = proj([tower.feature_select(hidden, n)])[0]
        feats = feats.reshape(-1, feats.shape[-1])
        x, _ = ops.embed_splice(ids_dev, m.embed_tokens, feats, self.im_patch_token, feat_offset, feat_limit)
        logits2d, _, _, _ = self._decoder_core(x, B, S, None, cache)
        return logits2d

    # ------------------------------------------------------------------ generation (llama_mmgpt.py:114-134)
    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None,
                                      inputs_embeds=None, **kwargs):
        # "is there a past" by LENGTH: an empty HF DynamicCache is truthy (SURVEY.md §0 fact 9 — the reference's
        # `if past_key_values:` at llama_mmgpt.py:117 slices the PROMPT to one token under transformers 5.x)
        has_past = past_key_values is not None and (
            past_key_values.get_seq_length() > 0 if hasattr(past_key_values, "get_seq_length") else bool(past_key_values))
        if has_past:
            input_ids = input_ids[:, -1:]
        if inputs_embeds is not None and not has_past:
            model_inputs = {"inputs_embeds": inputs_embeds}
        else:
            model_inputs = {"input_ids": input_ids}
        model_inputs.update({"past_key_values": past_key_values, "use_cache": kwargs.get("use_cache"),
                             "attention_mask": attention_mask, "images": kwargs.get("images", None)})
        return model_inputs

    def _get_decoder(self, cache: MerlinKVCache):
        fp = _decode.weight_fingerprint(self)
        if self._decoder is None or self._decoder.cache is not cache or self._decoder.fingerprint != fp:
            self._decoder = _decode.DecodeRunner(self, cache)
            self._decoder.fingerprint = fp
        return self._decoder

    @torch.no_grad()
    def _generate_beam(self, input_ids, images, num_beams, max_new_tokens, eos_token_id, stopping_criteria,
                       length_penalty, early_stopping, length_norm):
        """generate(num_beams=k) (eval_box.py:246-254): ONE multimodal prefill (the reference prefills k identical
        copies, base_mmgpt.py:162-163) into the SHARED prefix cache, then one batched decode pass per position for
        all k beams (merlin_decode_step_batched: weights streamed once, per-beam suffix caches) with the beam
        bookkeeping of merlin_b200/beam.py (pinned against HF `generate`); log-softmax and the top-2k selection run
        on the GPU, and a re-ordering moves only the beams' generated-token blocks (HF re-orders the whole cache)."""
        from ....beam import BeamSearch
        if num_beams > _decode.MAX_BEAMS:
            return self._generate_beam_sequential(input_ids, images, num_beams, max_new_tokens, eos_token_id,
                                                  stopping_criteria, length_penalty, early_stopping, length_norm)
        device = self.device
        ids = input_ids.to(device)
        S = ids.shape[1]
        runner = self._beam_runner(num_beams, S, max_new_tokens)
        runner.prefix.seq_len = 0
        flag, self.use_beam_search = getattr(self, "use_beam_search", False), False
        try:
            inputs = self.prepare_inputs_for_generation(ids, None, None, use_cache=True, images=images)
            inputs["past_key_values"] = runner.prefix          # empty: the prefill fills it
            out = self.forward(**inputs)
        finally:
            self.use_beam_search = flag
        runner.begin(S)
        bs = BeamSearch(input_ids.cpu(), num_beams, eos_token_id, max_new_tokens, length_penalty, early_stopping,
                        length_norm, stopping_criteria)
        res = bs.step(out.logits[:, -1, :].float().repeat(num_beams, 1))
        while res is not None:
            tokens, parents = res
            res = bs.step(runner.step(tokens, parents))
        return bs.finalize().to(device)

    @torch.no_grad()
    def _generate_beam_sequential(self, input_ids, images, num_beams, max_new_tokens, eos_token_id, stopping_criteria,
                       length_penalty, early_stopping, length_norm):
        """More than 8 beams: k single-sequence decode steps per position, one full KV cache per beam re-gathered by
        parent (HF's `_reorder_cache` scheme).  The batched path (`_generate_beam`) covers k <= 8."""
        from ....beam import BeamSearch
        device = self.device
        ids = input_ids.to(device)
        S = ids.shape[1]
        self._cache_capacity = S + max_new_tokens + 8
        flag, self.use_beam_search = getattr(self, "use_beam_search", False), False
        try:
            out = self.forward(**self.prepare_inputs_for_generation(ids, None, None, use_cache=True, images=images))
        finally:
            self.use_beam_search = flag
        cfg = self.config
        caches = [out.past_key_values]
        for _ in range(2 * num_beams - 1):      # beams 1..k-1 and k spares for the re-gather
            caches.append(MerlinKVCache(cfg.num_hidden_layers, 1, caches[0].capacity, cfg.hidden_size, device))
        blocks = lambda n: (n + 31) // 32

        def copy_cache(dst, src, n_tokens):
            for a, b in zip(dst.kv, src.kv):
                a[:, :blocks(n_tokens)].copy_(b[:, :blocks(n_tokens)])
            dst.seq_len = src.seq_len

        for c in caches[1:num_beams]:
            copy_cache(c, caches[0], S)
        live, spare = caches[:num_beams], caches[num_beams:]
        runners = {id(c): _decode.DecodeRunner(self, c) for c in caches}
        bs = BeamSearch(input_ids.cpu(), num_beams, eos_token_id, max_new_tokens, length_penalty, early_stopping,
                        length_norm, stopping_criteria)
        res = bs.step(out.logits[:, -1, :].float().repeat(num_beams, 1))
        logits = torch.empty((num_beams, cfg.vocab_size), dtype=torch.float32, device=device)
        while res is not None:
            tokens, parents = res
            # beam j continues parent[j]: gather the caches (copies only where the parent changed)
            moved = [j for j in range(num_beams) if parents[j] != j]
            for j in moved:
                copy_cache(spare[j], live[parents[j]], live[parents[j]].seq_len)
            for j in moved:
                live[j], spare[j] = spare[j], live[j]
            tok = torch.tensor(tokens, dtype=torch.int64, device=device)
            emb = self.model.embed_tokens[tok]
            for j in range(num_beams):
                logits[j].copy_(runners[id(live[j])].step_logits(emb[j:j + 1])[0])
            res = bs.step(logits)
        return bs.finalize().to(device)

    def _beam_runner(self, num_beams: int, S: int, max_new_tokens: int):
        """Prefix cache + batched decode runner (+ its CUDA graph) per (beam count, capacity bucket)."""
        bucket = _decode.CAPACITY_BUCKET
        cap = (S + 8 + bucket - 1) // bucket * bucket
        scap = (max_new_tokens + 8 + 255) // 256 * 256
        pool = self.__dict__.setdefault("_beam_runners", {})
        fp = _decode.weight_fingerprint(self)
        for key in [k for k, r in pool.items() if r.fingerprint != fp]:
            pool.pop(key)
        fit = [k for k in pool if k[0] == num_beams and k[1] >= cap and k[2] >= scap]
        if fit:
            return pool[min(fit)]
        for key in [k for k in pool if k[0] == num_beams]:      # one runner per beam count: the larger one replaces it
            cap, scap = max(cap, key[1]), max(scap, key[2])
            pool.pop(key)
        while len(pool) >= _decode.MAX_DECODE_STATES:
            pool.pop(next(iter(pool)))
        cfg = self.config
        prefix = MerlinKVCache(cfg.num_hidden_layers, 1, cap, cfg.hidden_size, self.device)
        runner = pool[(num_beams, cap, scap)] = _decode.BeamDecodeRunner(self, prefix, num_beams, scap)
        return runner

    # generation keywords understood natively (HF names); anything else must go through use_hf_loop=True
    _GEN_KEYS = ("max_new_tokens", "max_length", "min_new_tokens", "do_sample", "temperature", "top_p", "top_k",
                 "num_beams", "eos_token_id", "pad_token_id", "bos_token_id", "length_penalty", "early_stopping",
                 "repetition_penalty", "num_return_sequences", "return_dict_in_generate", "output_scores", "use_cache")

    def _generation_settings(self, generation_config, kwargs):
        """HF semantics: self.generation_config, overridden by a passed GenerationConfig, overridden by keywords."""
        import copy
        gc = copy.deepcopy(generation_config if generation_config is not None else self.generation_config)
        explicit = {k: kwargs.pop(k) for k in list(kwargs) if k in self._GEN_KEYS}
        for k, v in explicit.items():
            setattr(gc, k, v)
        return gc, explicit

    @torch.no_grad()
    def generate(self, inputs: Optional[torch.LongTensor] = None, generation_config=None, logits_processor=None,
                 stopping_criteria=None, *, images: Optional[List[torch.Tensor]] = None, input_ids=None,
                 attention_mask=None, streamer=None, generator: Optional[torch.Generator] = None,
                 use_hf_loop: bool = False, **kwargs):
        """`generate(input_ids, images=[...], **hf_generation_kwargs)` — the call of eval_mmvet.py:101-120,
        eval_box.py:245-264, model_worker.py:169-180.  Greedy decoding, temperature / top-k / top-p sampling with
        repetition penalty and custom logits processors, and beam search (num_beams > 1, length_penalty,
        early_stopping) run on the library's KV-cache decode kernels: step 0 is the full multimodal prefill, later
        steps feed one token and skip the vision path.  Returns [1, S + new] ids, or a
        GenerateDecoderOnlyOutput with return_dict_in_generate=True.  Unknown keywords raise (nothing is silently
        swallowed); `use_hf_loop=True` runs HF's own GenerationMixin loop over this model's forward instead
        (every HF decoding mode at batch 1, without the CUDA-graph decode loop).  Calls are serialised per model
        by a lock: the serve worker runs generate() from background threads (model_worker.py:169-180)."""
        if inputs is None:
            inputs = input_ids
        if inputs is None:
            raise ValueError("generate() needs input_ids")
        if use_hf_loop:
            extra = {k: v for k, v in (("attention_mask", attention_mask), ("streamer", streamer)) if v is not None}
            with self._generate_lock():
                return GenerationMixin.generate(self, inputs, generation_config=generation_config,
                                                logits_processor=logits_processor, stopping_criteria=stopping_criteria,
                                                images=images, **extra, **kwargs)
        length_norm = kwargs.pop("length_norm", "total")
        gc, explicit = self._generation_settings(generation_config, kwargs)
        if kwargs:
            raise ValueError(f"generate(): unsupported keyword arguments {sorted(kwargs)}; pass use_hf_loop=True to run "
                             "HF's GenerationMixin loop over this model")
        input_ids = inputs
        if input_ids.shape[0] != 1:
            raise NotImplementedError("generate() handles one sequence at a time (reference eval scripts)")
        if (gc.num_return_sequences or 1) != 1:
            raise NotImplementedError("num_return_sequences > 1: use use_hf_loop=True")
        S = input_ids.shape[1]
        if "max_new_tokens" in explicit or gc.max_new_tokens is not None:
            max_new_tokens = int(gc.max_new_tokens)
        elif "max_length" in explicit:
            max_new_tokens = max(int(gc.max_length) - S, 0)
        else:
            max_new_tokens = 1024                     # the reference's eval / serve default
        eos = gc.eos_token_id if "eos_token_id" in explicit or gc.eos_token_id is not None else 2
        eos_token_id = eos[0] if isinstance(eos, (list, tuple)) and len(eos) == 1 else eos
        if isinstance(eos_token_id, (list, tuple)):
            raise NotImplementedError("several eos ids: use use_hf_loop=True")
        do_sample, temperature = bool(gc.do_sample), float(gc.temperature if gc.temperature is not None else 1.0)
        top_p = float(gc.top_p if gc.top_p is not None else 1.0)
        top_k = int(gc.top_k or 0)
        rep = float(gc.repetition_penalty or 1.0)
        min_new = int(gc.min_new_tokens or 0)
        num_beams = int(gc.num_beams or 1)
        if attention_mask is not None and not bool(attention_mask.bool().all()):
            raise NotImplementedError("generate() takes an unpadded prompt (attention_mask of ones)")
        with self._generate_lock():
            ids = self._generate_locked(input_ids, images, max_new_tokens, do_sample, temperature, top_p, top_k, rep,
                                        min_new, num_beams, eos_token_id, stopping_criteria, logits_processor, streamer,
                                        generator, gc, length_norm)
        if gc.return_dict_in_generate:
            from transformers.generation.utils import GenerateDecoderOnlyOutput
            return GenerateDecoderOnlyOutput(sequences=ids)
        return ids

    def _generate_lock(self):
        import threading
        lock = self.__dict__.get("_gen_lock")
        if lock is None:
            lock = self.__dict__.setdefault("_gen_lock", threading.RLock())
        return lock

    def _generate_locked(self, input_ids, images, max_new_tokens, do_sample, temperature, top_p, top_k, rep, min_new,
                         num_beams, eos_token_id, stopping_criteria, logits_processor, streamer, generator, gc,
                         length_norm):
        device = self.device
        if num_beams != 1:
            if do_sample or streamer is not None:
                raise NotImplementedError("beam search runs without sampling and without a streamer")
            return self._generate_beam(input_ids, images, num_beams, max_new_tokens, eos_token_id,
                                       stopping_criteria, float(gc.length_penalty if gc.length_penalty is not None else 1.0),
                                       gc.early_stopping or False, length_norm)
        plain = (not do_sample and not stopping_criteria and not logits_processor and streamer is None
                 and rep == 1.0 and min_new == 0 and max_new_tokens > 0)
        fast = (plain and images is not None and len(images) == 1
                and getattr(self, "use_cuda_graphs", True) and not getattr(self, "use_beam_search", False)
                and getattr(self.model, "vision_tower", None) is not None)
        if fast:
            # device-resident loop: prefill launches (a CUDA graph once the shape repeats) + 1 graph replay per token
            return _decode.graphed_generate(self, input_ids, images[0], max_new_tokens, eos_token_id)
        ids = input_ids.to(device)
        self._cache_capacity = ids.shape[1] + max_new_tokens + 8
        inputs = self.prepare_inputs_for_generation(ids, None, None, use_cache=True, images=images)
        out = self.forward(**inputs)
        cache = out.past_key_values
        last = out.logits[:, -1, :]
        if plain:
            # device-resident greedy loop: one CUDA-graph replay per token, no host round trips
            first = ops.argmax(last).view(1, 1)
            ids = torch.cat([ids, first], dim=1)
            if (eos_token_id is None or int(first) != eos_token_id) and max_new_tokens > 1:
                more = self._get_decoder(cache).greedy(first, max_new_tokens - 1, eos_token_id)
                ids = torch.cat([ids, more.view(1, -1)], dim=1)
            return ids
        for step in range(max_new_tokens):
            scores = last.float()
            if rep != 1.0:                                   # HF RepetitionPenaltyLogitsProcessor
                seen = scores.gather(1, ids)
                scores = scores.scatter(1, ids, torch.where(seen < 0, seen * rep, seen / rep))
            if min_new and step < min_new and eos_token_id is not None:   # MinNewTokensLengthLogitsProcessor
                scores[:, eos_token_id] = float("-inf")
            for proc in (logits_processor or ()):
                scores = proc(ids, scores)
            if do_sample and temperature > 0:
                probs = sampling_distribution(scores, temperature, top_p, top_k)
                nxt = torch.multinomial(probs, 1, generator=generator)
            elif scores is last or not (rep != 1.0 or min_new or logits_processor):
                nxt = ops.argmax(last).view(1, 1)
            else:
                nxt = scores.argmax(dim=-1, keepdim=True)
            ids = torch.cat([ids, nxt], dim=1)
            if streamer is not None:
                streamer.put(nxt.cpu())
            if eos_token_id is not None and int(nxt) == eos_token_id:
                break
            if stopping_criteria is not None and any(bool(torch.as_tensor(c(ids, last)).all()) for c in stopping_criteria):
                break
            if step + 1 == max_new_tokens:
                break
            step_in = self.prepare_inputs_for_generation(ids, cache, torch.ones_like(ids), use_cache=True,
                                                         images=images)
            last = self.forward(**step_in).logits[:, -1, :]
        if streamer is not None:
            streamer.end()
        return ids


def sampling_distribution(logits: torch.Tensor, temperature: float = 1.0, top_p: float = 1.0,
                          top_k: int = 0) -> torch.Tensor:
    """Next-token distribution of HF sampling (`generate(do_sample=True, temperature, top_p)`, the call of
    model_worker.py:169-176 / eval_box.py:256-263): logits / temperature (TemperatureLogitsWarper), top-k
    (TopKLogitsWarper), then nucleus filtering — sort ascending, drop every token whose cumulative probability is
    <= 1 - top_p, always keep the most likely one (TopPLogitsWarper, min_tokens_to_keep=1) — then softmax."""
    scores = logits.float() / temperature
    if top_k and top_k < scores.shape[-1]:
        kth = scores.topk(top_k, dim=-1).values[..., -1:]
        scores = scores.masked_fill(scores < kth, float("-inf"))
    if top_p < 1.0:
        sorted_logits, sorted_idx = torch.sort(scores, descending=False, dim=-1)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1.0 - top_p)
        remove[..., -1:] = False
        scores = scores.masked_fill(remove.scatter(-1, sorted_idx, remove), float("-inf"))
    return torch.softmax(scores, dim=-1)


MerlinForCausalLM = MMGPTLlamaForCausalLM  # the name BASELINE.json's north_star uses

# llama_mmgpt.py:137-138.  exist_ok: the reference's own module registers the same model_type when both are
# imported into one process (the parity tests do that).
AutoConfig.register("mmgpt", MMGPTConfig, exist_ok=True)
AutoModelForCausalLM.register(MMGPTConfig, MMGPTLlamaForCausalLM, exist_ok=True)
