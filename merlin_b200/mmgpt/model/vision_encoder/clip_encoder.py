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
        # clip_encoder.py:21-24: CLIPVisionModel.from_pretrained(args.vision_tower) FIRST, then the
        # `model.vision_tower.*` slice of the mmgpt checkpoint overlays it.  The parameters above start at zero, so
        # a tower that found neither source is useless (constant features, zero gradients): say so loudly.
        self.weight_sources: List[str] = []
        if self._load_clip_checkpoint(self.vision_tower_name):
            self.weight_sources.append(f"clip:{self.vision_tower_name}")
        model_path = getattr(args, "model_name_or_path", None)
        if model_path and self.load_weights(model_path):
            self.weight_sources.append(f"mmgpt:{model_path}")
        if not self.weight_sources:
            import warnings
            warnings.warn(f"CLIPVisionTower: no pretrained weights found (vision_tower={self.vision_tower_name!r} is not "
                          f"a local CLIP checkpoint directory and model_name_or_path={model_path!r} holds no "
                          "'model.vision_tower.*' keys): every tower parameter is ZERO until load_state_dict() is "
                          "called", RuntimeWarning, stacklevel=2)

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
        return bool(weights)

    def _load_clip_checkpoint(self, name_or_path) -> bool:
        """`CLIPVisionModel.from_pretrained(vision_tower)` (clip_encoder.py:22) for a LOCAL checkpoint directory
        (there is no hub access): reads pytorch_model.bin / model.safetensors (sharded or not) of a CLIPVisionModel
        or a full CLIPModel, keeps the `vision_model.*` keys and loads them (q/k/v fused into qkv_w, the patch
        convolution flattened into patch_weight by load_state_dict's reference-layout views).  A position table of
        another resolution is interpolated (interpolate_model.py:11-28)."""
        if not name_or_path or not os.path.isdir(str(name_or_path)):
            return False
        path = str(name_or_path)
        sd = {}
        files = sorted(os.listdir(path))
        for idx in ("pytorch_model.bin.index.json", "model.safetensors.index.json"):
            if idx in files:
                files = sorted(set(json.load(open(os.path.join(path, idx)))["weight_map"].values()))
                break
        for f in files:
            full = os.path.join(path, f)
            if f.endswith(".safetensors"):
                from safetensors.torch import load_file
                sd.update(load_file(full))
            elif f.endswith(".bin") and f.startswith("pytorch_model"):
                sd.update(torch.load(full, map_location="cpu"))
        own = self.state_dict(keep_vars=True)
        picked = {}
        for k, v in sd.items():
            if k.startswith("vision_model."):
                k = "vision_tower." + k
            if k in own:
                picked[k] = v
        if not picked:
            return False
        missing = [k for k in own if k not in picked]
        if missing:
            raise RuntimeError(f"CLIP checkpoint {path} lacks {len(missing)} tower tensors, e.g. {missing[:3]}")
        self.load_state_dict(picked, strict=True, interpolate=True)
        return True

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

    def _c_weights(self):
        """struct merlin_vit_weights for the layers `select_layer` needs (rebuilt when a parameter moved)."""
        import ctypes
        from .... import lib as _lib
        c = self._config
        n_run = self.num_layers_to_run()
        key = tuple(p.data_ptr() for p in (self.class_embedding, self.patch_weight, self.position_embedding,
                                           self.qkv_w[0], self.fc2_w[n_run - 1] if n_run else self.pre_ln_w)) + (n_run,)
        cached = self.__dict__.get("_cw")
        if cached is not None and cached[0] == key:
            return cached[1]
        layers = (_lib.VitLayerWeights * max(n_run, 1))()
        for j in range(n_run):
            layers[j] = _lib.VitLayerWeights(*[t.data_ptr() for t in (
                self.qkv_w[j], self.qkv_b[j], self.out_w[j], self.out_b[j], self.ln1_w[j], self.ln1_b[j],
                self.ln2_w[j], self.ln2_b[j], self.fc1_w[j], self.fc1_b[j], self.fc2_w[j], self.fc2_b[j])])
        w = _lib.VitWeights(n_run, c.hidden_size, c.num_attention_heads, c.intermediate_size, c.image_size,
                            c.patch_size, self.k_patch_padded, float(c.layer_norm_eps),
                            self.class_embedding.data_ptr(), self.patch_weight.data_ptr(),
                            self.position_embedding.data_ptr(), self.pre_ln_w.data_ptr(), self.pre_ln_b.data_ptr(),
                            ctypes.cast(layers, ctypes.POINTER(_lib.VitLayerWeights)))
        self.__dict__["_cw"] = (key, w, layers)
        return w

    def encode(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """[n,3,H,W] -> selected hidden state [n*T, d] (row-major tokens, CLS first): ONE library call
        (merlin_vit_forward) for patch embedding, CLS / position / pre-LN and every encoder layer."""
        import ctypes
        from .... import lib as _lib
        c = self._config
        n = pixel_values.shape[0]
        d, H = c.hidden_size, c.num_attention_heads
        px = pixel_values.to(device=self.device, dtype=torch.bfloat16).contiguous()
        if not ops.PROFILE.enabled:
            L_ = _lib.load()
            w = self._c_weights()
            ws = torch.empty(L_.merlin_vit_workspace_bytes(ctypes.byref(w), n), dtype=torch.uint8, device=px.device)
            h = torch.empty((n * self.tokens, d), dtype=torch.bfloat16, device=px.device)
            ops.check(L_.merlin_vit_forward(torch.cuda.current_stream().cuda_stream, ctypes.byref(w), px.data_ptr(), n,
                                            ws.data_ptr(), h.data_ptr(), ops.DEFAULT_CTA_GROUP), "merlin_vit_forward",
                      3 + 7 * self.num_layers_to_run())
            return h
        # roofline leg of bench.py: per-op launches with a CUDA-event pair around every GEMM
        col = ops.patch_im2col(px, c.patch_size, self.k_patch_padded)
        patch_out = ops.gemm(col, self.patch_weight)
        h = ops.vit_embed_ln(patch_out, self.class_embedding, self.position_embedding, self.pre_ln_w,
                             self.pre_ln_b, n, c.layer_norm_eps)
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
        """encode() that keeps what backward needs (merlin_vit_forward_train).  Returns (hidden [n*T, d], saved)."""
        import ctypes
        from .... import lib as _lib
        n = pixel_values.shape[0]
        px = pixel_values.to(device=self.device, dtype=torch.bfloat16).contiguous()
        L_ = _lib.load()
        w = self._c_weights()
        saved = torch.empty(L_.merlin_vit_saved_bytes(ctypes.byref(w), n), dtype=torch.uint8, device=px.device)
        h = torch.empty((n * self.tokens, self._config.hidden_size), dtype=torch.bfloat16, device=px.device)
        ops.check(L_.merlin_vit_forward_train(torch.cuda.current_stream().cuda_stream, ctypes.byref(w), px.data_ptr(), n,
                                              saved.data_ptr(), h.data_ptr(), ops.DEFAULT_CTA_GROUP),
                  "merlin_vit_forward_train", 4 + 9 * self.num_layers_to_run())
        return h, dict(n=n, saved=saved)

    def backward(self, dh: torch.Tensor, saved, grad, f32, accumulate: bool, done=None):
        """dh: gradient w.r.t. encode_train()'s hidden (overwritten).  grad(name) -> bf16 gradient tensor of a
        weight matrix, f32[name] -> fp32 accumulator for biases / norm weights / position + class embeddings.
        ONE library call (merlin_vit_backward) walks the layers last-to-first."""
        import ctypes
        from .... import lib as _lib
        n, n_run = saved["n"], self.num_layers_to_run()
        L_ = _lib.load()
        w = self._c_weights()
        lg = (_lib.VitLayerGrads * max(n_run, 1))()
        for j in range(n_run):
            p = self.VM + f"encoder.layers.{j}."
            lg[j] = _lib.VitLayerGrads(
                grad(p + "self_attn.qkv_proj.weight").data_ptr(), grad(p + "self_attn.out_proj.weight").data_ptr(),
                grad(p + "mlp.fc1.weight").data_ptr(), grad(p + "mlp.fc2.weight").data_ptr(),
                f32[p + "self_attn.qkv_proj.bias"].data_ptr(), f32[p + "self_attn.out_proj.bias"].data_ptr(),
                f32[p + "layer_norm1.weight"].data_ptr(), f32[p + "layer_norm1.bias"].data_ptr(),
                f32[p + "layer_norm2.weight"].data_ptr(), f32[p + "layer_norm2.bias"].data_ptr(),
                f32[p + "mlp.fc1.bias"].data_ptr(), f32[p + "mlp.fc2.bias"].data_ptr())
        g = _lib.VitGrads(grad(self.VM + "embeddings.patch_embedding.weight").data_ptr(),
                          f32[self.VM + "embeddings.position_embedding.weight"].data_ptr(),
                          f32[self.VM + "embeddings.class_embedding"].data_ptr(),
                          f32[self.VM + "pre_layrnorm.weight"].data_ptr(), f32[self.VM + "pre_layrnorm.bias"].data_ptr(),
                          ctypes.cast(lg, ctypes.POINTER(_lib.VitLayerGrads)))
        ws = torch.empty(L_.merlin_vit_backward_workspace_bytes(ctypes.byref(w), n), dtype=torch.uint8, device=dh.device)
        if not dh.is_contiguous():
            raise ValueError("CLIPVisionTower.backward: dh must be contiguous")
        ops.check(L_.merlin_vit_backward(torch.cuda.current_stream().cuda_stream, ctypes.byref(w), ctypes.byref(g),
                                         saved["saved"].data_ptr(), dh.data_ptr(), ws.data_ptr(), n,
                                         int(bool(accumulate)), ops.DEFAULT_CTA_GROUP), "merlin_vit_backward",
                  3 + 20 * n_run)
        saved["saved"] = None
        if done is not None:
            done(self.VM + "embeddings.patch_embedding.weight")

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
