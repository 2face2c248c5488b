"""CPU oracle for the Merlin / mmgpt hot path — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 restatement of what the reference computes for
`MMGPTLlamaForCausalLM.forward` (vision tower -> projector -> embedding splice -> Llama decoder ->
lm_head -> shifted cross-entropy) and the greedy decode loop, written as pure functions over a flat
state dict that uses the reference's own parameter names.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may
import this module.  The product path (merlin_b200/) never does and has no CPU fallback.

Pinning: the reference ships no golden vectors or tests for this path, so this oracle
is pinned against the reference itself: `oracle/make_golden.py` imports the real
`mmgpt.model.mmgpt.llama_mmgpt.MMGPTLlamaForCausalLM` (+ CLIPVisionTower, ConvProjector / MLPProjector)
from /root/reference (out-of-tree copy with the one missing ')' of mmgpt/utils/constants.py:72 restored),
loads the same seeded weights, and asserts oracle == reference on CPU fp32 before writing
tests/golden/*.pt.  The third-party arithmetic the reference delegates to (transformers 4.31 Llama /
CLIP modules, flash-attn 2) is restated from the installed transformers 5.5.0 sources ("HF:" below)
with the 4.31-vs-5.5 deltas of SURVEY.md §8c neutralised (fp32 RoPE angles, eager attention).

Citations: paths are relative to /root/reference; "HF:" = transformers/models/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100  # mmgpt/utils/constants.py:7

StateDict = Dict[str, torch.Tensor]

VT = "model.vision_tower.vision_tower.vision_model."  # clip_encoder.py:22 + base_mmgpt.py:43
PROJ = "model.projector.projector."                    # conv_projector.py:14 / mlp_projector.py:16


@dataclass
class OracleConfig:
    # Llama (Vicuna) decoder — SURVEY.md Appendix A
    vocab_size: int = 32003
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_layers: int = 32
    num_heads: int = 32
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    # CLIP vision tower
    vit_hidden: int = 1024
    vit_intermediate: int = 4096
    vit_layers: int = 24
    vit_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    vit_ln_eps: float = 1e-5
    select_layer: int = -2            # mmgpt/utils/arguments.py:14
    select_feature: str = "patch"
    # projector
    projector: str = "conv"           # "conv" | "mlp"   (projector/builder.py:8-41)
    conv_stride: int = 2
    # special tokens (base_mmgpt.py:55-63, mm_utils.py:105)
    im_patch_token: int = 32000
    im_start_token: int = 32001
    im_end_token: int = 32002
    use_im_start_end: bool = True
    use_beam_search: bool = False

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def vit_tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def num_patches(self) -> int:  # clip_encoder.py:105-107
        return (self.image_size // self.patch_size // self.conv_stride) ** 2 if self.projector == "conv" \
            else (self.image_size // self.patch_size) ** 2

    @staticmethod
    def merlin_7b(**kw) -> "OracleConfig":
        return OracleConfig(**kw)

    @staticmethod
    def tiny(**kw) -> "OracleConfig":
        """Smallest config the CUDA kernels accept (decoder head_dim 128, ViT head_dim 64)."""
        base = dict(vocab_size=259, hidden_size=256, intermediate_size=384, num_layers=2, num_heads=2,
                    vit_hidden=128, vit_intermediate=256, vit_layers=4, vit_heads=2, image_size=56,
                    patch_size=14, im_patch_token=256, im_start_token=257, im_end_token=258)
        base.update(kw)
        return OracleConfig(**base)


# --------------------------------------------------------------------------- weights
def init_state_dict(cfg: OracleConfig, seed: int = 0, dtype=torch.float32, round_bf16: bool = True,
                    device="cpu") -> StateDict:
    """Seeded random-init weights under the reference's state-dict key names.  Draw order is fixed, so
    the same (cfg, seed, device type) gives identical tensors on every machine with the same torch build.
    Linear / embedding weights N(0, 0.02) as HF Llama `_init_weights`; CLIP-like scales for the tower.
    With round_bf16 the values are exactly representable in bf16 (the GPU path stores bf16).
    device="cuda" draws on the GPU (a 7B / 13B model in seconds; a different stream than the CPU generator,
    so the committed CPU goldens are unaffected); each tensor is converted to `dtype` as it is drawn."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd: StateDict = {}

    def fin(t):
        if round_bf16:
            t = t.to(torch.bfloat16)
        return t.to(dtype)

    def normal(name, shape, std):
        t = torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, std, generator=g)
        sd[name] = fin(t)

    def const(name, shape, value, jitter=0.0):
        t = torch.full(shape, float(value), dtype=torch.float32, device=device)
        if jitter:
            t += torch.empty(shape, dtype=torch.float32, device=device).normal_(0.0, jitter, generator=g)
        sd[name] = fin(t)

    D, FF, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    normal("model.embed_tokens.weight", (V, D), 0.02)
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        for n in ("q", "k", "v", "o"):
            normal(p + f"self_attn.{n}_proj.weight", (D, D), 0.02)
        normal(p + "mlp.gate_proj.weight", (FF, D), 0.02)
        normal(p + "mlp.up_proj.weight", (FF, D), 0.02)
        normal(p + "mlp.down_proj.weight", (D, FF), 0.02)
        const(p + "input_layernorm.weight", (D,), 1.0, 0.05)
        const(p + "post_attention_layernorm.weight", (D,), 1.0, 0.05)
    const("model.norm.weight", (D,), 1.0, 0.05)
    normal("lm_head.weight", (V, D), 0.02)

    d, ff = cfg.vit_hidden, cfg.vit_intermediate
    normal(VT + "embeddings.class_embedding", (d,), d ** -0.5)
    normal(VT + "embeddings.patch_embedding.weight", (d, 3, cfg.patch_size, cfg.patch_size), 0.02)
    normal(VT + "embeddings.position_embedding.weight", (cfg.vit_tokens, d), 0.02)
    const(VT + "pre_layrnorm.weight", (d,), 1.0, 0.05)
    const(VT + "pre_layrnorm.bias", (d,), 0.0, 0.02)
    for j in range(cfg.vit_layers):
        p = VT + f"encoder.layers.{j}."
        for n in ("q", "k", "v", "out"):
            normal(p + f"self_attn.{n}_proj.weight", (d, d), d ** -0.5)
            const(p + f"self_attn.{n}_proj.bias", (d,), 0.0, 0.02)
        for n in ("layer_norm1", "layer_norm2"):
            const(p + n + ".weight", (d,), 1.0, 0.05)
            const(p + n + ".bias", (d,), 0.0, 0.02)
        normal(p + "mlp.fc1.weight", (ff, d), d ** -0.5)
        const(p + "mlp.fc1.bias", (ff,), 0.0, 0.02)
        normal(p + "mlp.fc2.weight", (d, ff), ff ** -0.5)
        const(p + "mlp.fc2.bias", (d,), 0.0, 0.02)
    const(VT + "post_layernorm.weight", (d,), 1.0, 0.05)
    const(VT + "post_layernorm.bias", (d,), 0.0, 0.02)
    if cfg.projector == "conv":
        normal(PROJ + "weight", (D, d, 3, 3), (9 * d) ** -0.5)
    else:
        normal(PROJ + "weight", (D, d), d ** -0.5)
    const(PROJ + "bias", (D,), 0.0, 0.02)
    return sd


def synthetic_batch(cfg: OracleConfig, batch: int, seq_len: int, images_per_sample: Sequence[int], seed: int = 1234,
                    pad_to: Optional[int] = None):
    """Collator-shaped synthetic batch (mmgpt/data/collator.py:12-34; token grammar
    interpair_webdataset.py:47-49): bos, then per image <im_start> P'x<im_patch> <im_end> '\\n',
    then random text ids, eos.  Labels mask the visual span and the first half of the text
    (interpair_webdataset.py:121).  pad_to > seq_len right-pads with unk=0 / -100."""
    g = torch.Generator().manual_seed(seed)
    P = cfg.num_patches
    ids_rows, lab_rows, imgs = [], [], []
    text_hi = min(cfg.im_patch_token, cfg.vocab_size)
    for b in range(batch):
        n = images_per_sample[b]
        ids = [1]
        for _ in range(n):
            ids += [cfg.im_start_token] + [cfg.im_patch_token] * P + [cfg.im_end_token, 13]
        n_text = seq_len - len(ids) - 1
        if n_text < 1:
            raise ValueError("seq_len too small for the visual tokens")
        text = torch.randint(3, text_hi, (n_text,), generator=g).tolist()
        prefix = len(ids)
        ids = ids + text + [2]
        labels = list(ids)
        masked = prefix + (n_text + 1) // 2
        for t in range(masked):
            labels[t] = IGNORE_INDEX
        total = pad_to or seq_len
        ids += [0] * (total - seq_len)
        labels += [IGNORE_INDEX] * (total - seq_len)
        ids_rows.append(ids)
        lab_rows.append(labels)
        imgs.append(torch.empty(n, 3, cfg.image_size, cfg.image_size).normal_(0, 1, generator=g))
    input_ids = torch.tensor(ids_rows, dtype=torch.int64)
    labels = torch.tensor(lab_rows, dtype=torch.int64)
    attention_mask = input_ids.ne(0)
    if pad_to is None or pad_to == seq_len:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    return dict(input_ids=input_ids, labels=labels, attention_mask=attention_mask, images=imgs)


# --------------------------------------------------------------------------- vision tower
def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """HF:activations.py QuickGELUActivation — x * sigmoid(1.702 x)."""
    return x * torch.sigmoid(1.702 * x)


def clip_embeddings(sd: StateDict, cfg: OracleConfig, pixel_values: torch.Tensor) -> torch.Tensor:
    """HF:clip/modeling_clip.py:202-218 CLIPVisionEmbeddings.forward."""
    w = sd[VT + "embeddings.patch_embedding.weight"]
    patch = F.conv2d(pixel_values.to(w.dtype), w, bias=None, stride=cfg.patch_size)  # [n, d, G, G]
    patch = patch.flatten(2).transpose(1, 2)                                          # [n, G*G, d]
    cls = sd[VT + "embeddings.class_embedding"].expand(patch.shape[0], 1, -1)
    return torch.cat([cls, patch], dim=1) + sd[VT + "embeddings.position_embedding.weight"][None]


def clip_attention(sd: StateDict, cfg: OracleConfig, prefix: str, h: torch.Tensor) -> torch.Tensor:
    """HF:clip/modeling_clip.py:261-279,300-336 — bidirectional MHA, scale head_dim**-0.5, biases."""
    n, T, d = h.shape
    H = cfg.vit_heads
    hd = d // H

    def proj(name):
        return F.linear(h, sd[prefix + f"self_attn.{name}_proj.weight"], sd[prefix + f"self_attn.{name}_proj.bias"])

    q = proj("q").view(n, T, H, hd).transpose(1, 2)
    k = proj("k").view(n, T, H, hd).transpose(1, 2)
    v = proj("v").view(n, T, H, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1, dtype=torch.float32).to(q.dtype)
    o = (att @ v).transpose(1, 2).reshape(n, T, d)
    return F.linear(o, sd[prefix + "self_attn.out_proj.weight"], sd[prefix + "self_attn.out_proj.bias"])


def clip_encoder_layer(sd: StateDict, cfg: OracleConfig, j: int, h: torch.Tensor) -> torch.Tensor:
    """HF:clip/modeling_clip.py:363-384 CLIPEncoderLayer.forward (+ CLIPMLP :347-351)."""
    p = VT + f"encoder.layers.{j}."
    r = h
    h = F.layer_norm(h, (h.shape[-1],), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], cfg.vit_ln_eps)
    h = r + clip_attention(sd, cfg, p, h)
    r = h
    h = F.layer_norm(h, (h.shape[-1],), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], cfg.vit_ln_eps)
    h = F.linear(quick_gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])),
                 sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return r + h


def vision_tower_forward(sd: StateDict, cfg: OracleConfig, images: List[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
    """mmgpt/model/vision_encoder/clip_encoder.py:74-82 CLIPVisionTower.forward: cat the per-sample
    image lists, run the ViT, take hidden_states[select_layer] (:64-72), drop CLS, split back."""
    x = torch.cat([im for im in images], dim=0)
    sizes = [im.shape[0] for im in images]
    h = clip_embeddings(sd, cfg, x)
    d = h.shape[-1]
    h = F.layer_norm(h, (d,), sd[VT + "pre_layrnorm.weight"], sd[VT + "pre_layrnorm.bias"], cfg.vit_ln_eps)  # HF :676-677
    states = [h]
    # hidden_states = (embeddings, layer_1, ..., layer_L); select_layer=-2 needs layers[:L-1] only
    n_run = cfg.vit_layers + 1 + cfg.select_layer if cfg.select_layer < 0 else cfg.select_layer
    for j in range(n_run):
        h = clip_encoder_layer(sd, cfg, j, h)
        states.append(h)
    feat = states[n_run]
    if cfg.select_feature == "patch":
        feat = feat[:, 1:]
    elif cfg.select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {cfg.select_feature}")
    return torch.split(feat, sizes, dim=0)


# --------------------------------------------------------------------------- projector
def projector_forward(sd: StateDict, cfg: OracleConfig, features: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """conv: mmgpt/model/projector/conv_projector.py:23-39; mlp: mlp_projector.py:19-23."""
    out = []
    for feat in features:
        if cfg.projector == "mlp":
            out.append(F.linear(feat, sd[PROJ + "weight"], sd[PROJ + "bias"]))
            continue
        if feat.dim() == 1:  # dummy feature (conv_projector.py:27-28)
            feat = feat.view(1, 1, -1).repeat(1, 256, 1)
        B, P, C = feat.shape
        HW = int(math.sqrt(P))
        x = feat.permute(0, 2, 1).reshape(B, C, HW, HW)
        y = F.conv2d(x, sd[PROJ + "weight"], sd[PROJ + "bias"], stride=cfg.conv_stride, padding=1)
        out.append(y.reshape(B, -1, P // (cfg.conv_stride ** 2)).permute(0, 2, 1))
    return out


def encode_images(sd: StateDict, cfg: OracleConfig, images: List[torch.Tensor]) -> List[torch.Tensor]:
    """mmgpt/model/mmgpt/base_mmgpt.py:18-21."""
    return projector_forward(sd, cfg, vision_tower_forward(sd, cfg, images))


# --------------------------------------------------------------------------- splice
def prepare_inputs_embeds(sd: StateDict, cfg: OracleConfig, input_ids: torch.Tensor,
                          image_features: Optional[Sequence[torch.Tensor]]) -> torch.Tensor:
    """mmgpt/model/mmgpt/base_mmgpt.py:99-165: embed_tokens gather, then for every sample and every
    <im_start> splice that image's projected features over the following P' rows.  Restated literally
    (per-sample / per-image loop, same ValueErrors, zip truncation to len(image_features))."""
    embeds = F.embedding(input_ids, sd["model.embed_tokens.weight"])
    if image_features is None:
        return embeds
    new_rows = []
    for cur_ids, cur_emb, cur_feats in zip(input_ids, embeds, image_features):
        if (cur_ids == cfg.im_patch_token).sum() == 0:
            new_rows.append(cur_emb)  # the reference adds 0 * projector(dummy).sum(): a no-op on values
            continue
        if not cfg.use_im_start_end:
            raise NotImplementedError  # base_mmgpt.py:136-137
        if (cur_ids == cfg.im_start_token).sum() != (cur_ids == cfg.im_end_token).sum():
            raise ValueError("The number of image start tokens and image end tokens should be the same.")
        starts = torch.where(cur_ids == cfg.im_start_token)[0]
        for pos, feats in zip(starts, cur_feats):
            n = feats.shape[0]
            if cur_ids[pos + n + 1] != cfg.im_end_token:
                raise ValueError("The image end token should follow the image start token.")
            cur_emb = torch.cat((cur_emb[:pos + 1], feats.to(cur_emb.dtype), cur_emb[pos + n + 1:]), dim=0)
        new_rows.append(cur_emb)
    out = torch.stack(new_rows, dim=0)
    if cfg.use_beam_search:
        out = out.repeat_interleave(5, dim=0)  # base_mmgpt.py:162-163
    return out


# --------------------------------------------------------------------------- decoder
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """HF:llama/modeling_llama.py:62-67 LlamaRMSNorm.forward (fp32 statistics)."""
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (w.to(torch.float32) * xf).to(x.dtype)


def rope_cos_sin(cfg: OracleConfig, positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """HF:llama/modeling_llama.py:117-135 — fp32 inv_freq, angles, cos/sin; emb = cat(freqs, freqs)."""
    hd = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32, device=positions.device) / hd))
    freqs = positions.to(torch.float32)[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """HF:llama/modeling_llama.py:138-142."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def llama_attention(sd: StateDict, cfg: OracleConfig, i: int, h: torch.Tensor, key_mask: Optional[torch.Tensor],
                    positions: torch.Tensor, kv_cache: Optional[list] = None) -> torch.Tensor:
    """The patched LlamaAttention.forward (mmgpt/utils/llama_flash_attn_monkey_patch.py:20-103):
    q/k/v projections, half-split RoPE, causal softmax(QKᵀ/sqrt(hd)) V with a right-padding key mask
    (flash-attn varlen + unpad/pad: padded query rows come back as zeros), o_proj.
    With kv_cache (decode, HF eager path) keys/values of earlier positions are prepended."""
    B, S, D = h.shape
    H, hd = cfg.num_heads, cfg.head_dim
    p = f"model.layers.{i}.self_attn."
    q = F.linear(h, sd[p + "q_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
    k = F.linear(h, sd[p + "k_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
    v = F.linear(h, sd[p + "v_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
    cos, sin = rope_cos_sin(cfg, positions)                    # [S, hd]
    cos, sin = cos.to(q.dtype)[None, None], sin.to(q.dtype)[None, None]
    q = q * cos + rotate_half(q) * sin                        # HF:llama/modeling_llama.py:145-168
    k = k * cos + rotate_half(k) * sin
    if kv_cache is not None:
        if kv_cache[i] is not None:
            k = torch.cat([kv_cache[i][0], k], dim=2)
            v = torch.cat([kv_cache[i][1], v], dim=2)
        kv_cache[i] = (k, v)
    T = k.shape[2]
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    q_pos = positions.view(1, 1, S, 1)
    k_pos = torch.arange(T, device=positions.device).view(1, 1, 1, T)
    mask = k_pos > q_pos                                       # causal=True (monkey_patch.py:82-84)
    if key_mask is not None:
        mask = mask | ~key_mask.view(B, 1, 1, T).bool()        # key_padding_mask (monkey_patch.py:78)
    scores = scores.masked_fill(mask, float("-inf"))
    att = torch.softmax(scores, dim=-1, dtype=torch.float32).to(q.dtype)
    o = (att @ v).transpose(1, 2).reshape(B, S, D)
    if key_mask is not None and kv_cache is None:
        o = o * key_mask.view(B, S, 1).to(o.dtype)             # pad_input zero-fills (monkey_patch.py:97-102)
    return F.linear(o, sd[p + "o_proj.weight"])


def llama_mlp(sd: StateDict, i: int, h: torch.Tensor) -> torch.Tensor:
    """HF:llama/modeling_llama.py:182-184 — down(silu(gate(x)) * up(x))."""
    p = f"model.layers.{i}.mlp."
    return F.linear(F.silu(F.linear(h, sd[p + "gate_proj.weight"])) * F.linear(h, sd[p + "up_proj.weight"]),
                    sd[p + "down_proj.weight"])


def decoder_forward(sd: StateDict, cfg: OracleConfig, inputs_embeds: torch.Tensor,
                    attention_mask: Optional[torch.Tensor], positions: Optional[torch.Tensor] = None,
                    kv_cache: Optional[list] = None, return_layers: bool = False):
    """HF LlamaModel.forward / LlamaDecoderLayer.forward (HF:llama/modeling_llama.py:303-333,375-425):
    x += Attn(RMSNorm(x)); x += MLP(RMSNorm(x)); final RMSNorm.  positions default to arange(S)."""
    h = inputs_embeds
    S = h.shape[1]
    if positions is None:
        positions = torch.arange(S, device=h.device)
    key_mask = None
    if attention_mask is not None and not bool(attention_mask.all()):
        key_mask = attention_mask
    layers = []
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        h = h + llama_attention(sd, cfg, i, rms_norm(h, sd[p + "input_layernorm.weight"], cfg.rms_norm_eps),
                                key_mask, positions, kv_cache)
        h = h + llama_mlp(sd, i, rms_norm(h, sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps))
        if return_layers:
            layers.append(h)
    h = rms_norm(h, sd["model.norm.weight"], cfg.rms_norm_eps)
    return (h, layers) if return_layers else h


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """mmgpt/model/mmgpt/llama_mmgpt.py:89-100 — shift, flatten, CrossEntropyLoss() (mean over != -100)."""
    V = logits.shape[-1]
    return F.cross_entropy(logits[..., :-1, :].reshape(-1, V).float(), labels[..., 1:].reshape(-1),
                           ignore_index=IGNORE_INDEX)


def forward(sd: StateDict, cfg: OracleConfig, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
            labels: Optional[torch.Tensor] = None, images: Optional[List[torch.Tensor]] = None) -> dict:
    """mmgpt/model/mmgpt/llama_mmgpt.py:53-112 MMGPTLlamaForCausalLM.forward (prefill / training)."""
    feats = encode_images(sd, cfg, images) if images is not None else None
    embeds = prepare_inputs_embeds(sd, cfg, input_ids, feats)
    hidden = decoder_forward(sd, cfg, embeds, attention_mask)
    logits = F.linear(hidden, sd["lm_head.weight"])
    out = dict(logits=logits, inputs_embeds=embeds, hidden_states=hidden, image_features=feats)
    if labels is not None:
        out["loss"] = causal_lm_loss(logits, labels)
    return out


def greedy_generate(sd: StateDict, cfg: OracleConfig, input_ids: torch.Tensor, images: List[torch.Tensor],
                    max_new_tokens: int, eos_token_id: Optional[int] = 2, use_cache: bool = True) -> torch.Tensor:
    """Greedy restatement of the generate() contract: step 0 runs the full multimodal
    forward (llama_mmgpt.py:114-134 prepare_inputs_for_generation, past=None); later steps feed only the
    last token and skip the vision path (base_mmgpt.py:85-89).  use_cache=False re-runs the full forward
    each step (the cache-free loop that is valid against the real reference under transformers 5.x)."""
    assert input_ids.shape[0] == 1
    ids = input_ids.clone()
    if not use_cache:
        for _ in range(max_new_tokens):
            logits = forward(sd, cfg, ids, torch.ones_like(ids, dtype=torch.bool), None, images)["logits"]
            nxt = logits[:, -1].argmax(-1, keepdim=True)
            ids = torch.cat([ids, nxt], dim=1)
            if eos_token_id is not None and int(nxt) == eos_token_id:
                break
        return ids
    cache: list = [None] * cfg.num_layers
    feats = encode_images(sd, cfg, images)
    embeds = prepare_inputs_embeds(sd, cfg, ids, feats)
    h = decoder_forward(sd, cfg, embeds, None, torch.arange(ids.shape[1], device=ids.device), cache)
    for _ in range(max_new_tokens):
        nxt = F.linear(h[:, -1], sd["lm_head.weight"]).argmax(-1, keepdim=True)
        ids = torch.cat([ids, nxt], dim=1)
        if eos_token_id is not None and int(nxt) == eos_token_id:
            break
        e = F.embedding(nxt, sd["model.embed_tokens.weight"])
        h = decoder_forward(sd, cfg, e, None, torch.tensor([ids.shape[1] - 1], device=ids.device), cache)
    return ids


# --------------------------------------------------------------------------- optimizer grouping (N1)
def vit_lr_scale(key: str) -> float:
    """mmgpt/utils/llrd_utils.py:4-14."""
    import re
    if "vision_model.encoder.layers" in key:
        layer = int(re.findall(r"layers\.(\d+)\.", key)[0])
        return 0.9 ** (23 - layer - 1)
    if "vision_model" in key:
        return 0.1
    return 1


def param_groups(named_shapes: Sequence[Tuple[str, Tuple[int, ...]]], lr: float, wd: float, llrd: bool = True):
    """mmgpt/utils/llrd_utils.py:25-79 get_param_groups on (name, shape) pairs; returns a list of
    dict(names=[...], weight_decay=..., lr=...) in the reference's group order."""
    wd_plain, wd_scaled, nowd_plain, nowd_scaled = [], {}, [], {}
    for name, shape in named_shapes:
        no_wd = name.endswith(".bias") or len(shape) == 1
        mult = vit_lr_scale(name) if llrd else 1
        scaled = mult != 1
        if not no_wd and not scaled:
            wd_plain.append(name)
        elif not no_wd and scaled:
            wd_scaled.setdefault(mult, []).append(name)
        elif no_wd and not scaled:
            nowd_plain.append(name)
        else:
            nowd_scaled.setdefault(mult, []).append(name)
    groups = []
    if wd_plain:
        groups.append(dict(names=wd_plain, weight_decay=wd, lr=lr))
    for mult, names in wd_scaled.items():
        groups.append(dict(names=names, weight_decay=wd, lr=lr * mult))
    if nowd_plain:
        groups.append(dict(names=nowd_plain, weight_decay=0.0, lr=lr))
    for mult, names in nowd_scaled.items():
        groups.append(dict(names=names, weight_decay=0.0, lr=lr * mult))
    return groups
