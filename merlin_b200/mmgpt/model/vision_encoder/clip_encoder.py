"""CLIPVisionTower on B200 kernels.

Mirror of the reference's mmgpt/model/vision_encoder/clip_encoder.py:11-107 (which wraps HF
CLIPVisionModel): same constructor arguments, `forward(images: list[Tensor[n_i,3,H,W]]) ->
tuple[Tensor[n_i,P,C]]`, `feature_select`, `load_weights`, and the properties
hidden_size / num_patches / dummy_feature / dtype / device / config.  The ViT itself
(HF:clip/modeling_clip.py:202-218,363-384,676-677) runs as: im2col -> tcgen05 GEMM (patch embedding) ->
fused CLS/pos-emb/pre-LN kernel -> per layer {LayerNorm, fused QKV GEMM+bias, non-causal
flash-attention (hd 64), out_proj GEMM + bias + residual, LayerNorm, fc1 GEMM + bias + quick_gelu,
fc2 GEMM + bias + residual}.  Only the layers needed for `select_layer` are executed (the reference
runs layer 24 and throws its output away, clip_encoder.py:65 with select_layer=-2).

Parameters are stored packed for the kernels (q/k/v fused, patch kernel flattened and zero padded);
state_dict()/load_state_dict() speak the reference's key names and layouts
(`vision_tower.vision_model.encoder.layers.N.self_attn.q_proj.weight`, ...).
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict, defaultdict
from types import SimpleNamespace
from typing import List, Optional

import torch
import torch.nn as nn

from .... import ops
from ....lib import EPI_QUICKGELU, EPI_RESIDUAL


def default_clip_config(name: str = "clip-vit-large-patch14-336") -> SimpleNamespace:
    """CLIP-ViT-L/14 dims; image size is taken from the tower name."""
    image_size = 448 if "448" in name else (336 if "336" in name else 224)
    return SimpleNamespace(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                           num_attention_heads=16, image_size=image_size, patch_size=14,
                           hidden_act="quick_gelu", layer_norm_eps=1e-5, num_channels=3)


def _load_clip_config(name_or_path: str) -> SimpleNamespace:
    cfg_file = os.path.join(name_or_path or "", "config.json")
    if os.path.exists(cfg_file):
        raw = json.load(open(cfg_file))
        raw = raw.get("vision_config", raw)
        base = default_clip_config(name_or_path).__dict__
        base.update({k: raw[k] for k in base if k in raw})
        return SimpleNamespace(**base)
    return default_clip_config(name_or_path or "")


class CLIPVisionTower(nn.Module):
    def __init__(self, args, config: Optional[SimpleNamespace] = None):
        super().__init__()
        self.vision_tower_name = args.vision_tower
        self.select_layer = args.vision_select_layer
        self.select_feature = args.vision_select_feature
        self.freeze_vision_tower = args.freeze_vision_tower
        self.conv_stride = args.conv_stride
        self._config = config if config is not None else _load_clip_config(self.vision_tower_name)
        c = self._config
        if c.hidden_act != "quick_gelu":
            raise ValueError("CLIPVisionTower: only quick_gelu towers are supported (CLIP-L/14)")
        if c.hidden_size // c.num_attention_heads != 64:
            raise ValueError("CLIPVisionTower: the attention kernel needs head_dim 64")
        self.image_processor = self._build_image_processor()

        d, ff, L = c.hidden_size, c.intermediate_size, c.num_hidden_layers
        self.tokens = (c.image_size // c.patch_size) ** 2 + 1
        self.k_patch = 3 * c.patch_size * c.patch_size
        self.k_patch_padded = (self.k_patch + 7) // 8 * 8  # TMA needs 16-byte row pitch

        def P(*shape):
            return nn.Parameter(torch.zeros(*shape), requires_grad=not self.freeze_vision_tower)

        self.class_embedding = P(d)
        self.patch_weight = P(d, self.k_patch_padded)      # conv weight [d,3,p,p] flattened (c,ky,kx)
        self.position_embedding = P(self.tokens, d)
        self.pre_ln_w, self.pre_ln_b = P(d), P(d)
        self.post_ln_w, self.post_ln_b = P(d), P(d)        # unused by the hot path; kept for checkpoints
        self.qkv_w = nn.ParameterList([P(3 * d, d) for _ in range(L)])
        self.qkv_b = nn.ParameterList([P(3 * d) for _ in range(L)])
        self.out_w = nn.ParameterList([P(d, d) for _ in range(L)])
        self.out_b = nn.ParameterList([P(d) for _ in range(L)])
        self.ln1_w = nn.ParameterList([P(d) for _ in range(L)])
        self.ln1_b = nn.ParameterList([P(d) for _ in range(L)])
        self.ln2_w = nn.ParameterList([P(d) for _ in range(L)])
        self.ln2_b = nn.ParameterList([P(d) for _ in range(L)])
        self.fc1_w = nn.ParameterList([P(ff, d) for _ in range(L)])
        self.fc1_b = nn.ParameterList([P(ff) for _ in range(L)])
        self.fc2_w = nn.ParameterList([P(d, ff) for _ in range(L)])
        self.fc2_b = nn.ParameterList([P(d) for _ in range(L)])
        model_path = getattr(args, "model_name_or_path", None)
        if model_path:
            self.load_weights(model_path)

    # ------------------------------------------------------------------ reference-format state dict
    def _build_image_processor(self):
        try:
            from transformers import CLIPImageProcessor
            s = self._config.image_size
            return CLIPImageProcessor(size={"shortest_edge": s}, crop_size={"height": s, "width": s})
        except Exception:  # pragma: no cover - transformers without vision extras
            return None

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """Reference key names/layouts: `vision_tower.vision_model....` (clip_encoder.py:22)."""
        c = self._config
        d = c.hidden_size
        out = destination if destination is not None else OrderedDict()
        vm = prefix + "vision_tower.vision_model."

        def put(k, v):
            out[k] = v if keep_vars else v.detach()

        put(vm + "embeddings.class_embedding", self.class_embedding)
        put(vm + "embeddings.patch_embedding.weight",
            self.patch_weight[:, : self.k_patch].reshape(d, 3, c.patch_size, c.patch_size))
        put(vm + "embeddings.position_embedding.weight", self.position_embedding)
        put(vm + "pre_layrnorm.weight", self.pre_ln_w)
        put(vm + "pre_layrnorm.bias", self.pre_ln_b)
        for j in range(c.num_hidden_layers):
            p = vm + f"encoder.layers.{j}."
            for n, blk in (("q", 0), ("k", 1), ("v", 2)):
                put(p + f"self_attn.{n}_proj.weight", self.qkv_w[j][blk * d:(blk + 1) * d])
                put(p + f"self_attn.{n}_proj.bias", self.qkv_b[j][blk * d:(blk + 1) * d])
            put(p + "self_attn.out_proj.weight", self.out_w[j])
            put(p + "self_attn.out_proj.bias", self.out_b[j])
            put(p + "layer_norm1.weight", self.ln1_w[j])
            put(p + "layer_norm1.bias", self.ln1_b[j])
            put(p + "mlp.fc1.weight", self.fc1_w[j])
            put(p + "mlp.fc1.bias", self.fc1_b[j])
            put(p + "mlp.fc2.weight", self.fc2_w[j])
            put(p + "mlp.fc2.bias", self.fc2_b[j])
            put(p + "layer_norm2.weight", self.ln2_w[j])
            put(p + "layer_norm2.bias", self.ln2_b[j])
        put(vm + "post_layernorm.weight", self.post_ln_w)
        put(vm + "post_layernorm.bias", self.post_ln_b)
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False, interpolate: bool = False):
        """interpolate=True: a position table saved at another image size (e.g. the 577-row 336 px CLIP-L/14
        table loaded into a 448 px tower) is resized like mmgpt/utils/interpolate_model.py:11-28 does."""
        own = self.state_dict(keep_vars=True)
        pos_key = "vision_tower.vision_model.embeddings.position_embedding.weight"
        if interpolate and pos_key in state_dict and state_dict[pos_key].shape[0] != self.tokens:
            from ...utils.interpolate_model import interpolate_position_embedding
            state_dict = dict(state_dict)
            grid = self._config.image_size // self._config.patch_size
            state_dict[pos_key] = interpolate_position_embedding(state_dict[pos_key], grid)
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own and not k.endswith("position_ids")]
        if strict and (missing or unexpected):
            raise RuntimeError(f"CLIPVisionTower.load_state_dict: missing {missing[:4]}... unexpected {unexpected[:4]}...")
        with torch.no_grad():
            for k, dst in own.items():
                if k in state_dict:
                    src = state_dict[k]
                    if tuple(src.shape) != tuple(dst.shape):
                        raise RuntimeError(f"size mismatch for {k}: {tuple(src.shape)} vs {tuple(dst.shape)}")
                    dst.copy_(src)  # views write through to the packed storage
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_weights(self, model_path):
        """Same checkpoint discovery as the reference (clip_encoder.py:26-62): pull the
        `model.vision_tower.*` slice out of pytorch_model.bin[.index.json]."""
        prefix = "model.vision_tower."
        index = os.path.join(model_path, "pytorch_model.bin.index.json")
        single = os.path.join(model_path, "pytorch_model.bin")
        weights = {}
        if os.path.exists(index):
            by_file = defaultdict(list)
            for k, v in json.load(open(index))["weight_map"].items():
                if "vision_tower" in k:
                    by_file[v].append(k)
            for fname, keys in by_file.items():
                ckpt = torch.load(os.path.join(model_path, fname), map_location="cpu")
                for k in keys:
                    weights[k[len(prefix):]] = ckpt[k]
        elif os.path.exists(single):
            ckpt = torch.load(single, map_location="cpu")
            weights = {k[len(prefix):]: v for k, v in ckpt.items() if "vision_tower" in k}
        if weights:
            self.load_state_dict(weights, strict=True)

    # ------------------------------------------------------------------------------ forward
    def feature_select(self, hidden: torch.Tensor, n_img: int) -> torch.Tensor:
        """hidden [n_img*T, d] -> [n_img, P, d] view (clip_encoder.py:64-72)."""
        h = hidden.view(n_img, self.tokens, -1)
        if self.select_feature == "patch":
            return h[:, 1:]
        if self.select_feature == "cls_patch":
            return h
        raise ValueError(f"Unexpected select feature: {self.select_feature}")

    def num_layers_to_run(self) -> int:
        L = self._config.num_hidden_layers
        return L + 1 + self.select_layer if self.select_layer < 0 else self.select_layer

    def encode(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """[n,3,H,W] -> selected hidden state [n*T, d] (row-major tokens, CLS first)."""
        c = self._config
        n = pixel_values.shape[0]
        d, H = c.hidden_size, c.num_attention_heads
        px = pixel_values.to(device=self.device, dtype=torch.bfloat16).contiguous()
        col = ops.patch_im2col(px, c.patch_size, self.k_patch_padded)
        patch_out = ops.gemm(col, self.patch_weight)
        h = ops.vit_embed_ln(patch_out, self.class_embedding, self.position_embedding, self.pre_ln_w,
                             self.pre_ln_b, n, c.layer_norm_eps)
        if not ops.PROFILE.enabled:
            ws = ops.vit_layer_workspace(h.shape[0], d, c.intermediate_size, h.device)
            for j in range(self.num_layers_to_run()):
                ops.vit_layer_forward(h, ws, self, j, n)      # merlin_vit_layer_forward: 7 kernels, one call
            return h
        xn = torch.empty_like(h)
        for j in range(self.num_layers_to_run()):
            ops.layernorm(h, self.ln1_w[j], self.ln1_b[j], c.layer_norm_eps, out=xn)
            qkv = ops.gemm(xn, self.qkv_w[j], bias=self.qkv_b[j])
            att = ops.flash_attn(qkv, n, self.tokens, H, 64, causal=False)
            ops.gemm(att, self.out_w[j], bias=self.out_b[j], residual=h, out=h, epilogue=EPI_RESIDUAL)
            ops.layernorm(h, self.ln2_w[j], self.ln2_b[j], c.layer_norm_eps, out=xn)
            mid = ops.gemm(xn, self.fc1_w[j], bias=self.fc1_b[j], epilogue=EPI_QUICKGELU)
            ops.gemm(mid, self.fc2_w[j], bias=self.fc2_b[j], residual=h, out=h, epilogue=EPI_RESIDUAL)
        return h

    # ------------------------------------------------------------------ training form (explicit backward)
    VM = "model.vision_tower.vision_tower.vision_model."

    def named_parameters_reference(self):
        """(reference key, packed parameter) of every tensor that receives a gradient, in backward
        completion order: encoder layers last-to-first, then pre_layrnorm and the embeddings.  Layers beyond
        `select_layer` and post_layernorm never get gradients (clip_encoder.py:65) and are left out."""
        out = []
        for j in reversed(range(self.num_layers_to_run())):
            p = self.VM + f"encoder.layers.{j}."
            out += [(p + "mlp.fc2.weight", self.fc2_w[j]), (p + "mlp.fc2.bias", self.fc2_b[j]),
                    (p + "mlp.fc1.weight", self.fc1_w[j]), (p + "mlp.fc1.bias", self.fc1_b[j]),
                    (p + "layer_norm2.weight", self.ln2_w[j]), (p + "layer_norm2.bias", self.ln2_b[j]),
                    (p + "self_attn.out_proj.weight", self.out_w[j]), (p + "self_attn.out_proj.bias", self.out_b[j]),
                    (p + "self_attn.qkv_proj.weight", self.qkv_w[j]), (p + "self_attn.qkv_proj.bias", self.qkv_b[j]),
                    (p + "layer_norm1.weight", self.ln1_w[j]), (p + "layer_norm1.bias", self.ln1_b[j])]
        out += [(self.VM + "pre_layrnorm.weight", self.pre_ln_w), (self.VM + "pre_layrnorm.bias", self.pre_ln_b),
                (self.VM + "embeddings.position_embedding.weight", self.position_embedding),
                (self.VM + "embeddings.class_embedding", self.class_embedding),
                (self.VM + "embeddings.patch_embedding.weight", self.patch_weight)]
        return [(n, p) for n, p in out if p.requires_grad]

    def encode_train(self, pixel_values: torch.Tensor):
        """encode() that keeps what backward needs.  Returns (hidden [n*T, d], saved)."""
        c = self._config
        n, H = pixel_values.shape[0], c.num_attention_heads
        px = pixel_values.to(device=self.device, dtype=torch.bfloat16).contiguous()
        col = ops.patch_im2col(px, c.patch_size, self.k_patch_padded)
        patch_out = ops.gemm(col, self.patch_weight)
        emb = ops.vit_embed(patch_out, self.class_embedding, self.position_embedding, n)
        h = ops.layernorm(emb, self.pre_ln_w, self.pre_ln_b, c.layer_norm_eps)
        layers = []
        for j in range(self.num_layers_to_run()):
            xn1 = ops.layernorm(h, self.ln1_w[j], self.ln1_b[j], c.layer_norm_eps)
            qkv = ops.gemm(xn1, self.qkv_w[j], bias=self.qkv_b[j])
            att, lse = ops.flash_attn(qkv, n, self.tokens, H, 64, causal=False, return_lse=True)
            h_mid = ops.gemm(att, self.out_w[j], bias=self.out_b[j], residual=h, epilogue=EPI_RESIDUAL)
            xn2 = ops.layernorm(h_mid, self.ln2_w[j], self.ln2_b[j], c.layer_norm_eps)
            pre = ops.gemm(xn2, self.fc1_w[j], bias=self.fc1_b[j])
            mid = ops.quick_gelu(pre)
            h_out = ops.gemm(mid, self.fc2_w[j], bias=self.fc2_b[j], residual=h_mid, epilogue=EPI_RESIDUAL)
            layers.append((h, xn1, qkv, att, lse, h_mid, xn2, pre, mid))
            h = h_out
        return h, dict(n=n, col=col, emb=emb, layers=layers)

    def backward(self, dh: torch.Tensor, saved, grad, f32, accumulate: bool, done=None):
        """dh: gradient w.r.t. encode_train()'s hidden.  grad(name) -> bf16 arena view for weight matrices,
        f32[name] -> fp32 accumulator for biases / norm weights / position + class embeddings."""
        c = self._config
        n, H, eps = saved["n"], c.num_attention_heads, c.layer_norm_eps
        for j in reversed(range(self.num_layers_to_run())):
            h, xn1, qkv, att, lse, h_mid, xn2, pre, mid = saved["layers"][j]
            p = self.VM + f"encoder.layers.{j}."
            dmid = ops.gemm_nt(dh, self.fc2_w[j])
            ops.gemm_tn(dh, mid, out=grad(p + "mlp.fc2.weight"), accumulate=accumulate)
            ops.colsum_accumulate(dh, f32[p + "mlp.fc2.bias"])
            dpre = ops.quick_gelu(pre, dy=dmid.contiguous())
            ops.gemm_tn(dpre, xn2, out=grad(p + "mlp.fc1.weight"), accumulate=accumulate)
            ops.colsum_accumulate(dpre, f32[p + "mlp.fc1.bias"])
            dxn2 = ops.gemm_nt(dpre, self.fc1_w[j])
            dh_mid = ops.layernorm_bwd(h_mid, self.ln2_w[j], dxn2, dh, f32[p + "layer_norm2.weight"],
                                       f32[p + "layer_norm2.bias"], eps)
            datt = ops.gemm_nt(dh_mid, self.out_w[j])
            ops.gemm_tn(dh_mid, att, out=grad(p + "self_attn.out_proj.weight"), accumulate=accumulate)
            ops.colsum_accumulate(dh_mid, f32[p + "self_attn.out_proj.bias"])
            dqkv = ops.flash_attn_bwd(qkv, att, datt, lse, n, self.tokens, H, 64, causal=False)
            ops.gemm_tn(dqkv, xn1, out=grad(p + "self_attn.qkv_proj.weight"), accumulate=accumulate)
            ops.colsum_accumulate(dqkv, f32[p + "self_attn.qkv_proj.bias"])
            dxn1 = ops.gemm_nt(dqkv, self.qkv_w[j])
            dh = ops.layernorm_bwd(h, self.ln1_w[j], dxn1, dh_mid, f32[p + "layer_norm1.weight"],
                                   f32[p + "layer_norm1.bias"], eps)
            saved["layers"][j] = None
            if done is not None:
                done(p + "layer_norm1.bias")
        demb = ops.layernorm_bwd(saved["emb"], self.pre_ln_w, dh, None, f32[self.VM + "pre_layrnorm.weight"],
                                 f32[self.VM + "pre_layrnorm.bias"], eps)
        dpatch = ops.vit_embed_bwd(demb, f32[self.VM + "embeddings.position_embedding.weight"].view(self.tokens, -1),
                                   f32[self.VM + "embeddings.class_embedding"], n)
        ops.gemm_tn(dpatch, saved["col"], out=grad(self.VM + "embeddings.patch_embedding.weight"), accumulate=accumulate)

    def forward(self, images: List[torch.Tensor]):
        with torch.set_grad_enabled(False):  # forward-only kernels; training path: see engine/train
            concat = torch.cat([im for im in images], dim=0)
            sizes = [im.shape[0] for im in images]
            hidden = self.encode(concat)
            feats = self.feature_select(hidden, concat.shape[0]).to(images[0].dtype if images[0].is_floating_point() and images[0].dtype == torch.bfloat16 else torch.bfloat16)
            return torch.split(feats, sizes, dim=0)

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)

    @property
    def dtype(self):
        return self.class_embedding.dtype

    @property
    def device(self):
        return self.class_embedding.device

    @property
    def config(self):
        return self._config

    @property
    def hidden_size(self):
        return self._config.hidden_size

    @property
    def num_patches(self):
        return (self._config.image_size // self._config.patch_size // self.conv_stride) ** 2
