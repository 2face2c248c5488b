"""Thin torch-facing wrappers over the C ABI: torch owns memory and streams, the library computes.

Every function takes CUDA bf16 tensors (unless noted), passes raw pointers + the current stream to
libmerlin_b200.so and returns torch tensors.  No op here has a PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import lib as _lib
from .lib import EPI_QUICKGELU, EPI_RESIDUAL, EPI_STORE, EPI_SWIGLU
from .lib import check as _check

DEFAULT_CTA_GROUP = 2  # CTA-pair (cta_group::2) GEMM tiles unless overridden

_BF16 = torch.bfloat16


def check(status: int, what: str, launches: int = 0) -> None:
    _check(status, what)
    PROFILE.launches += launches


class _Profile:
    """Launch counter + optional CUDA-event timing of every GEMM launch (bench.py's roofline leg)."""

    def __init__(self):
        self.enabled = False
        self.launches = 0
        self.events = []

    def reset(self, enabled: bool = False):
        self.enabled = enabled
        self.events = []

    def gemm_summary(self):
        """(algorithmic FLOPs, milliseconds, launches) of the GEMM launches recorded since reset()."""
        torch.cuda.synchronize()
        flops = sum(f for _, _, f in self.events)
        ms = sum(a.elapsed_time(b) for a, b, _ in self.events)
        return flops, ms, len(self.events)


PROFILE = _Profile()


def _gemm_call(fn, flops, *args):
    PROFILE.launches += 1
    if not PROFILE.enabled:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    PROFILE.events.append((e0, e1, flops))
    return rc


from .composite_ops import (decoder_layer_backward, decoder_layer_forward_train,  # noqa: E402
                            decoder_layer_forward as _decoder_layer_forward_c, decoder_layer_workspace,
                            vit_layer_forward, vit_layer_workspace)


def decoder_layer_forward(*args, **kw):
    return _decoder_layer_forward_c(*args, **kw)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows2d(t: torch.Tensor, name: str) -> tuple[int, int, int]:
    """(rows, cols, leading dimension) of a 2-D bf16 CUDA tensor with unit inner stride."""
    if t.dtype != _BF16 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA bfloat16 tensor, got {t.dtype} on {t.device}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with contiguous rows, got {tuple(t.shape)} strides {t.stride()}")
    return t.shape[0], t.shape[1], t.stride(0)


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         epilogue: int = EPI_STORE, cta_group: Optional[int] = None) -> torch.Tensor:
    """out = epilogue(a @ w.T): a [M,K], w [N,K] (nn.Linear layout).  See merlin_gemm_bf16."""
    M, K, lda = _rows2d(a, "a")
    N, Kw, ldb = _rows2d(w, "w")
    if K != Kw:
        raise ValueError(f"gemm: K mismatch {K} vs {Kw}")
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        ld = (n_out + 7) // 8 * 8  # TMA needs a 16-byte row pitch
        out = torch.empty((M, ld), dtype=_BF16, device=a.device)[:, :n_out]
    Mo, No, ldc = _rows2d(out, "out")
    if (Mo, No) != (M, n_out):
        raise ValueError(f"gemm: out shape {(Mo, No)} != {(M, n_out)}")
    ldr = 0
    if epilogue == EPI_RESIDUAL:
        if residual is None:
            raise ValueError("gemm: residual epilogue without residual tensor")
        Mr, Nr, ldr = _rows2d(residual, "residual")
        if (Mr, Nr) != (M, n_out):
            raise ValueError("gemm: residual shape mismatch")
    if bias is not None and (bias.dtype != _BF16 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be a contiguous bf16 vector of length N")
    if M == 0:
        return out
    L = _lib.load()
    check(_gemm_call(L.merlin_gemm_bf16, 2 * M * N * K, _stream(), a.data_ptr(), lda, w.data_ptr(), ldb,
                     out.data_ptr(), ldc, _ptr(residual), ldr, _ptr(bias), M, N, K, epilogue,
                     cta_group or DEFAULT_CTA_GROUP), "merlin_gemm_bf16")
    return out


def gemm_nt(a: torch.Tensor, b_t: torch.Tensor, *, out: Optional[torch.Tensor] = None,
            accumulate: bool = False, cta_group: Optional[int] = None) -> torch.Tensor:
    """out[M,N] (+)= a[M,K] @ b_t[K,N]   (B given transposed / "MN-major").  The dgrad contraction:
    dX = dY @ W with a = dY [tokens, out], b_t = W [out, in]."""
    M, K, lda = _rows2d(a, "a")
    Kb, N, ldb = _rows2d(b_t, "b_t")
    if K != Kb:
        raise ValueError(f"gemm_nt: K mismatch {K} vs {Kb}")
    return _gemm_ex(a, lda, 0, b_t, ldb, 1, M, N, K, out, accumulate, cta_group)


def gemm_tn(a_t: torch.Tensor, b_t: torch.Tensor, *, out: Optional[torch.Tensor] = None,
            accumulate: bool = False, cta_group: Optional[int] = None) -> torch.Tensor:
    """out[M,N] (+)= a_t[K,M].T @ b_t[K,N]   (both operands transposed).  The wgrad contraction:
    dW = dY.T @ X with a_t = dY [tokens, out], b_t = X [tokens, in]."""
    K, M, lda = _rows2d(a_t, "a_t")
    Kb, N, ldb = _rows2d(b_t, "b_t")
    if K != Kb:
        raise ValueError(f"gemm_tn: K mismatch {K} vs {Kb}")
    return _gemm_ex(a_t, lda, 1, b_t, ldb, 1, M, N, K, out, accumulate, cta_group)


def _gemm_ex(a, lda, a_mn, b, ldb, b_mn, M, N, K, out, accumulate, cta_group):
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an output tensor")
        out = torch.empty((M, (N + 7) // 8 * 8), dtype=_BF16, device=a.device)[:, :N]
    Mo, No, ldc = _rows2d(out, "out")
    if (Mo, No) != (M, N):
        raise ValueError(f"gemm: out shape {(Mo, No)} != {(M, N)}")
    L = _lib.load()
    epi = EPI_RESIDUAL if accumulate else EPI_STORE
    check(_gemm_call(L.merlin_gemm_bf16_ex, 2 * M * N * K, _stream(), a.data_ptr(), lda, a_mn, b.data_ptr(), ldb,
                     b_mn, out.data_ptr(), ldc, out.data_ptr() if accumulate else None, ldc, None, M, N, K, epi,
                     cta_group or DEFAULT_CTA_GROUP), "merlin_gemm_bf16_ex")
    return out


def gemm_qkv_rope(a: torch.Tensor, w_qkv: torch.Tensor, rope_cos: torch.Tensor, rope_sin: torch.Tensor,
                  seq_len: int, *, positions: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None, cta_group: Optional[int] = None,
                  kv_cache: Optional[torch.Tensor] = None, batch: int = 1, heads: int = 0) -> torch.Tensor:
    """Packed qkv = a @ w_qkv.T with RoPE applied to the q and k blocks in the GEMM epilogue.  With `kv_cache`
    ([batch, capacity/32, heads, 32, 256]) the rotated keys and the values are also appended to the cache from the
    same epilogue (merlin_gemm_qkv_rope_kv; batch > 1 needs seq_len % 128 == 0)."""
    M, K, lda = _rows2d(a, "a")
    N, Kw, ldb = _rows2d(w_qkv, "w_qkv")
    if K != Kw or N % 3 != 0:
        raise ValueError("gemm_qkv_rope: bad shapes")
    if rope_cos.dtype != torch.float32 or rope_sin.dtype != torch.float32 or rope_cos.shape[-1] != 64:
        raise ValueError("gemm_qkv_rope: cos/sin tables must be fp32 [max_pos, 64]")
    if out is None:
        out = torch.empty((M, N), dtype=_BF16, device=a.device)
    _, _, ldc = _rows2d(out, "out")
    if positions is not None and positions.dtype != torch.int32:
        raise ValueError("positions must be int32")
    L = _lib.load()
    if kv_cache is not None:
        if kv_cache.dim() != 5 or kv_cache.shape[0] != batch or kv_cache.shape[2] != heads or kv_cache.shape[3] != 32 \
                or kv_cache.shape[4] != 256 or not kv_cache.is_contiguous():
            raise ValueError("gemm_qkv_rope: cache must be a contiguous [batch, capacity/32, heads, 32, 256] tensor")
        check(_gemm_call(L.merlin_gemm_qkv_rope_kv, 2 * M * N * K, _stream(), a.data_ptr(), lda, w_qkv.data_ptr(), ldb,
                         out.data_ptr(), ldc, rope_cos.data_ptr(), rope_sin.data_ptr(), _ptr(positions), M, N, K,
                         2 * N // 3, seq_len, kv_cache.data_ptr(), kv_cache.shape[1] * 32, batch, heads,
                         cta_group or DEFAULT_CTA_GROUP), "merlin_gemm_qkv_rope_kv")
        return out
    check(_gemm_call(L.merlin_gemm_qkv_rope, 2 * M * N * K, _stream(), a.data_ptr(), lda, w_qkv.data_ptr(), ldb,
                     out.data_ptr(), ldc, rope_cos.data_ptr(), rope_sin.data_ptr(), _ptr(positions), M, N, K,
                     2 * N // 3, seq_len, cta_group or DEFAULT_CTA_GROUP), "merlin_gemm_qkv_rope")
    return out


def image_preprocess(pixels: torch.Tensor, size: int, mean=(0.48145466, 0.4578275, 0.40821073),
                     std=(0.26862954, 0.26130258, 0.27577711), dtype=torch.float32, return_resized: bool = False,
                     pad_to_square: bool = False):
    """uint8 RGB image [H, W, 3] on the GPU -> [3, size, size]: PIL-exact bicubic resize, /255, CLIP mean / std
    (merlin_image_preprocess; mode 'resize' of base_dataset.py:190-192).  pad_to_square=True is mode 'pad'
    (base_dataset.py:186-188): the image is first pasted at the top-left corner of a max(H, W) square filled with
    int(255 * mean) (`expand2square`, mm_utils.py:10-21) — a device-side buffer fill + copy — then resized."""
    import ctypes
    if pixels.dtype != torch.uint8 or pixels.dim() != 3 or pixels.shape[2] != 3 or not pixels.is_cuda:
        raise ValueError("image_preprocess: expected a CUDA uint8 [H, W, 3] tensor")
    if dtype not in (torch.float32, _BF16):
        raise ValueError("image_preprocess: dtype must be float32 or bfloat16")
    if pad_to_square and pixels.shape[0] != pixels.shape[1]:
        side = max(pixels.shape[0], pixels.shape[1])
        canvas = torch.empty((side, side, 3), dtype=torch.uint8, device=pixels.device)
        canvas[:] = torch.tensor([int(m * 255) for m in mean], dtype=torch.uint8, device=pixels.device)
        canvas[:pixels.shape[0], :pixels.shape[1]] = pixels
        pixels = canvas
    pixels = pixels.contiguous()
    H, W, _ = pixels.shape
    L = _lib.load()
    ws = torch.empty(L.merlin_image_preprocess_workspace_bytes(H, W, size) + 256, dtype=torch.uint8, device=pixels.device)
    ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
    out = torch.empty((3, size, size), dtype=dtype, device=pixels.device)
    resized = torch.empty((size, size, 3), dtype=torch.uint8, device=pixels.device) if return_resized else None
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    check(L.merlin_image_preprocess(_stream(), pixels.data_ptr(), W * 3, H, W, size, m, s, ws_ptr, out.data_ptr(),
                                    1 if dtype == _BF16 else 0, _ptr(resized)), "merlin_image_preprocess")
    PROFILE.launches += 2
    return (out, resized) if return_resized else out


def kv_append(qkv: torch.Tensor, kv_cache: torch.Tensor, batch: int, seq: int, hidden: int, heads: int,
              pos0: int = 0) -> None:
    """Packed qkv rows [batch*seq, 3*hidden] -> cache [batch, capacity/32, heads, 32, 256] at tokens
    pos0 .. pos0+seq-1 (k rotated | v)."""
    _, _, ldq = _rows2d(qkv, "qkv")
    if kv_cache.dim() != 5 or kv_cache.shape[0] != batch or kv_cache.shape[2] != heads or kv_cache.shape[3] != 32 \
            or kv_cache.shape[4] != 256 or not kv_cache.is_contiguous():
        raise ValueError("kv_append: cache must be a contiguous [batch, capacity/32, heads, 32, 256] tensor")
    L = _lib.load()
    check(L.merlin_kv_append(_stream(), qkv.data_ptr(), ldq, kv_cache.data_ptr(), kv_cache.shape[1] * 32, batch, seq,
                             hidden, heads, pos0), "merlin_kv_append")
    PROFILE.launches += 1


def flash_attn(qkv: torch.Tensor, batch: int, seq: int, heads: int, head_dim: int, *, causal: bool,
               seqlens: Optional[torch.Tensor] = None, return_lse: bool = False,
               softmax_scale: float = 0.0, out: Optional[torch.Tensor] = None):
    """qkv [batch*seq, 3*heads*head_dim] packed [q|k|v] -> out [batch*seq, heads*head_dim]."""
    M, N, ld = _rows2d(qkv, "qkv")
    if M != batch * seq or N != 3 * heads * head_dim:
        raise ValueError(f"flash_attn: qkv shape {tuple(qkv.shape)} != ({batch * seq}, {3 * heads * head_dim})")
    if out is None:
        out = torch.empty((M, heads * head_dim), dtype=_BF16, device=qkv.device)
    _, _, ldo = _rows2d(out, "out")
    lse = torch.empty((batch, heads, seq), dtype=torch.float32, device=qkv.device) if return_lse else None
    if seqlens is not None and (seqlens.dtype != torch.int32 or seqlens.numel() != batch):
        raise ValueError("flash_attn: seqlens must be int32 [batch]")
    L = _lib.load()
    check(L.merlin_flash_attn_fwd(_stream(), qkv.data_ptr(), ld, out.data_ptr(), ldo, _ptr(lse),
                                  _ptr(seqlens), batch, seq, heads, head_dim, int(causal),
                                  float(softmax_scale)), "merlin_flash_attn_fwd", 1)
    return (out, lse) if return_lse else out


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, D, ldx = _rows2d(x, "x")
    if out is None:
        out = torch.empty((M, D), dtype=_BF16, device=x.device)
    _, _, ldy = _rows2d(out, "out")
    check(_lib.load().merlin_rmsnorm(_stream(), x.data_ptr(), ldx, weight.data_ptr(), out.data_ptr(), ldy,
                                     M, D, float(eps)), "merlin_rmsnorm", 1)
    return out


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, D, ldx = _rows2d(x, "x")
    if out is None:
        out = torch.empty((M, D), dtype=_BF16, device=x.device)
    _, _, ldy = _rows2d(out, "out")
    check(_lib.load().merlin_layernorm(_stream(), x.data_ptr(), ldx, weight.data_ptr(), bias.data_ptr(),
                                       out.data_ptr(), ldy, M, D, float(eps)), "merlin_layernorm", 1)
    return out


def patch_im2col(pixels: torch.Tensor, patch: int, ld_col: int) -> torch.Tensor:
    """pixels [n,3,HW,HW] bf16 -> col [n*(HW/patch)^2, ld_col] (zero padded beyond 3*patch*patch)."""
    if pixels.dtype != _BF16 or not pixels.is_contiguous() or pixels.dim() != 4 or pixels.shape[1] != 3:
        raise ValueError("patch_im2col: pixels must be contiguous bf16 [n,3,H,W]")
    n, _, hw, hw2 = pixels.shape
    if hw != hw2:
        raise ValueError("patch_im2col: square images only")
    g = hw // patch
    col = torch.empty((n * g * g, ld_col), dtype=_BF16, device=pixels.device)
    check(_lib.load().merlin_patch_im2col(_stream(), pixels.data_ptr(), col.data_ptr(), n, hw, patch, ld_col),
          "merlin_patch_im2col", 1)
    return col


def vit_embed_ln(patch_out: torch.Tensor, class_embedding: torch.Tensor, position_embedding: torch.Tensor,
                 ln_w: torch.Tensor, ln_b: torch.Tensor, n_img: int, eps: float) -> torch.Tensor:
    tokens, D = position_embedding.shape
    hidden = torch.empty((n_img * tokens, D), dtype=_BF16, device=patch_out.device)
    check(_lib.load().merlin_vit_embed_ln(_stream(), patch_out.data_ptr(), class_embedding.data_ptr(),
                                          position_embedding.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(),
                                          hidden.data_ptr(), n_img, tokens, D, float(eps)),
          "merlin_vit_embed_ln", 1)
    return hidden


def conv3x3_im2col(feat: torch.Tensor, n_img: int, grid: int, C: int, stride: int, tok_stride: int,
                   tok_off: int) -> torch.Tensor:
    go = (grid + 2 - 3) // stride + 1
    col = torch.empty((n_img * go * go, 9 * C), dtype=_BF16, device=feat.device)
    check(_lib.load().merlin_conv3x3_im2col(_stream(), feat.data_ptr(), col.data_ptr(), n_img, grid, C,
                                            stride, tok_stride, tok_off), "merlin_conv3x3_im2col", 1)
    return col


def projector_forward(feat: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], n_img: int, grid: int,
                      C: int, stride: int, tok_stride: int, tok_off: int, cta_group: Optional[int] = None) -> torch.Tensor:
    """ConvProjector forward in one library call (merlin_projector_forward): [n_img * Go * Go, D_out]."""
    D_out, K9, ldw = _rows2d(weight, "weight")
    if K9 != 9 * C or ldw != K9:
        raise ValueError("projector_forward: weight must be a contiguous [D_out, 9*C] tensor")
    go = (grid + 2 - 3) // stride + 1
    L = _lib.load()
    ws = torch.empty(L.merlin_projector_workspace_bytes(n_img, grid, C, stride), dtype=torch.uint8, device=feat.device)
    out = torch.empty((n_img * go * go, D_out), dtype=_BF16, device=feat.device)
    check(_gemm_call(L.merlin_projector_forward, 2 * n_img * go * go * D_out * 9 * C, _stream(), feat.data_ptr(),
                     weight.data_ptr(), _ptr(bias), ws.data_ptr(), out.data_ptr(), D_out, n_img, grid, C, D_out, stride,
                     tok_stride, tok_off, cta_group or DEFAULT_CTA_GROUP), "merlin_projector_forward")
    PROFILE.launches += 1
    return out


def embed_splice(input_ids: torch.Tensor, embed_table: torch.Tensor, image_features: Optional[torch.Tensor],
                 im_patch_token: int, feat_offset: Optional[torch.Tensor] = None,
                 feat_limit: Optional[torch.Tensor] = None, return_rank: bool = False):
    """input_ids [B, S] int64 (CUDA) -> (inputs_embeds [B*S, D], counts [B] = #<im_patch> per sample)."""
    if input_ids.dtype != torch.int64 or not input_ids.is_contiguous() or input_ids.dim() != 2:
        raise ValueError("embed_splice: input_ids must be contiguous int64 [B, S]")
    B, S = input_ids.shape
    V, D = embed_table.shape
    out = torch.empty((B * S, D), dtype=_BF16, device=embed_table.device)
    ws = torch.empty((B * S + B,), dtype=torch.int32, device=embed_table.device)
    nfeat = 0 if image_features is None else image_features.shape[0]
    if image_features is not None and (not image_features.is_contiguous() or image_features.shape[1] != D
                                       or image_features.dtype != _BF16):
        raise ValueError("embed_splice: image_features must be contiguous bf16 [rows, D]")
    for t in (feat_offset, feat_limit):
        if t is not None and (t.dtype != torch.int32 or t.numel() != B):
            raise ValueError("embed_splice: feat_offset / feat_limit must be int32 [B]")
    if B * S == 0:
        return (out, ws[B * S:], ws[:B * S]) if return_rank else (out, ws[B * S:])
    check(_lib.load().merlin_embed_splice(_stream(), input_ids.data_ptr(), embed_table.data_ptr(),
                                          _ptr(image_features), out.data_ptr(), ws.data_ptr(), B, S, D, nfeat,
                                          _ptr(feat_offset), _ptr(feat_limit), V, im_patch_token),
          "merlin_embed_splice", 2)
    if return_rank:  # per-token feature row (-1 for text tokens): what the backward scatter needs
        return out, ws[B * S:], ws[:B * S]
    return out, ws[B * S:]


def rope_qk_inplace(qkv: torch.Tensor, rope_cos: torch.Tensor, rope_sin: torch.Tensor, heads: int,
                    seq_len: int, positions: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, _, ld = _rows2d(qkv, "qkv")
    check(_lib.load().merlin_rope_qk_inplace(_stream(), qkv.data_ptr(), ld, rope_cos.data_ptr(),
                                             rope_sin.data_ptr(), _ptr(positions), M, heads, seq_len),
          "merlin_rope_qk_inplace", 1)
    return qkv


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, batch: int, seq: int,
                  ignore_index: int = -100) -> tuple[torch.Tensor, torch.Tensor]:
    """Shifted CE over logits [batch*seq, V].  Returns (mean loss (fp32 scalar), per-row nll)."""
    M, V, ld = _rows2d(logits, "logits")
    if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != M:
        raise ValueError("cross_entropy: labels must be contiguous int64 [batch, seq]")
    row = torch.empty((M,), dtype=torch.float32, device=logits.device)
    acc = torch.empty((2,), dtype=torch.float32, device=logits.device)
    check(_lib.load().merlin_cross_entropy_fwd(_stream(), logits.data_ptr(), ld, labels.data_ptr(), batch, seq,
                                               V, ignore_index, row.data_ptr(), acc.data_ptr()),
          "merlin_cross_entropy_fwd", 1)
    return acc[0] / acc[1], row


def lmhead_ce_forward(hidden: torch.Tensor, lm_head_w: torch.Tensor, labels: Optional[torch.Tensor], batch: int,
                      seq: int, ignore_index: int = -100, cta_group: Optional[int] = None):
    """logits = hidden @ lm_head_w.T (+ shifted CE when labels are given) in one library call
    (merlin_lmhead_ce_forward).  Returns (logits [M, V] view on a 16-byte pitch, loss or None, loss_sum_count)."""
    M, D, ldh = _rows2d(hidden, "hidden")
    V, Dw, ldw = _rows2d(lm_head_w, "lm_head_w")
    if D != Dw or ldw != D or M != batch * seq:
        raise ValueError("lmhead_ce_forward: bad shapes")
    buf = torch.empty((M, (V + 7) // 8 * 8), dtype=_BF16, device=hidden.device)
    logits = buf[:, :V]
    acc = None
    if labels is not None:
        if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != M:
            raise ValueError("lmhead_ce_forward: labels must be contiguous int64 [batch, seq]")
        acc = torch.empty((2,), dtype=torch.float32, device=hidden.device)
    L = _lib.load()
    check(_gemm_call(L.merlin_lmhead_ce_forward, 2 * M * V * D, _stream(), hidden.data_ptr(), ldh, lm_head_w.data_ptr(),
                     logits.data_ptr(), buf.stride(0), _ptr(labels), batch, seq, D, V, ignore_index, None, _ptr(acc),
                     cta_group or DEFAULT_CTA_GROUP), "merlin_lmhead_ce_forward")
    if labels is not None:
        PROFILE.launches += 1
    return logits, (None if acc is None else acc[0] / acc[1]), acc


def lmhead_ce_backward(logits: torch.Tensor, labels: torch.Tensor, acc: torch.Tensor, grad_scale: float,
                       hidden: torch.Tensor, lm_head_w: torch.Tensor, d_lm_head_w: Optional[torch.Tensor],
                       accumulate: bool, batch: int, seq: int, ignore_index: int = -100,
                       cta_group: Optional[int] = None) -> torch.Tensor:
    """Backward of lmhead_ce_forward (merlin_lmhead_ce_backward): logits become dlogits in place, d_lm_head_w gets
    the weight gradient, returns dhidden [M, D]."""
    M, V, ldl = _rows2d(logits, "logits")
    _, D, ldh = _rows2d(hidden, "hidden")
    dh = torch.empty((M, D), dtype=_BF16, device=hidden.device)
    L = _lib.load()
    check(L.merlin_lmhead_ce_backward(_stream(), logits.data_ptr(), ldl, labels.data_ptr(), batch, seq, D, V, ignore_index,
                                      acc.data_ptr(), float(grad_scale), hidden.data_ptr(), ldh, lm_head_w.data_ptr(),
                                      _ptr(d_lm_head_w), 1 if accumulate else 0, dh.data_ptr(), D,
                                      cta_group or DEFAULT_CTA_GROUP), "merlin_lmhead_ce_backward")
    PROFILE.launches += 3 if d_lm_head_w is not None else 2
    return dh


def argmax(logits: torch.Tensor) -> torch.Tensor:
    M, V, ld = _rows2d(logits, "logits")
    out = torch.empty((M,), dtype=torch.int64, device=logits.device)
    check(_lib.load().merlin_argmax(_stream(), logits.data_ptr(), ld, M, V, out.data_ptr()), "merlin_argmax", 1)
    return out


def rope_tables(max_pos: int, head_dim: int = 128, theta: float = 10000.0, device="cuda"):
    """fp32 cos/sin tables [max_pos, head_dim/2] from fp32 inv_freq (SURVEY §0 fact 10:
    HF:llama/modeling_llama.py:117-135 builds inv_freq / angles in fp32; never round inv_freq to bf16)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=device) / head_dim))
    t = torch.arange(max_pos, dtype=torch.float32, device=device)
    freqs = torch.outer(t, inv_freq)
    return freqs.cos().contiguous(), freqs.sin().contiguous()


def swiglu_interleave(gate_w: torch.Tensor, up_w: torch.Tensor) -> torch.Tensor:
    """Pack gate/up weights [FF, D] into the [2*FF, D] layout the SwiGLU epilogue expects:
    per 256-row tile, 128 gate rows then the matching 128 up rows.  FF must be a multiple of 128."""
    FF, D = gate_w.shape
    if FF % 128 != 0:
        raise ValueError("swiglu_interleave: intermediate size must be a multiple of 128")
    g = gate_w.view(FF // 128, 128, D)
    u = up_w.view(FF // 128, 128, D)
    return torch.stack((g, u), dim=1).reshape(2 * FF, D).contiguous()


# ----------------------------------------------------------------------------- backward / optimizer ops
def flash_attn_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, batch: int, seq: int,
                   heads: int, head_dim: int, *, causal: bool, seqlens: Optional[torch.Tensor] = None,
                   dqkv: Optional[torch.Tensor] = None, softmax_scale: float = 0.0) -> torch.Tensor:
    """Packed gradient [batch*seq, 3*heads*head_dim] w.r.t. the (rotated) packed qkv."""
    M, N, ld = _rows2d(qkv, "qkv")
    _, _, ldo = _rows2d(out, "out")
    _, _, ldd = _rows2d(dout, "dout")
    if ldo != ldd:
        raise ValueError("flash_attn_bwd: out and dout must share a row pitch")
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    _, _, ldg = _rows2d(dqkv, "dqkv")
    L = _lib.load()
    ws = torch.empty(L.merlin_flash_attn_bwd_workspace_bytes(batch, seq, heads, head_dim), dtype=torch.uint8,
                     device=qkv.device)
    check(L.merlin_flash_attn_bwd(_stream(), qkv.data_ptr(), ld, out.data_ptr(), dout.data_ptr(), ldo,
                                  lse.data_ptr(), _ptr(seqlens), dqkv.data_ptr(), ldg, ws.data_ptr(), batch, seq,
                                  heads, head_dim, int(causal), float(softmax_scale)), "merlin_flash_attn_bwd", 4)
    return dqkv


def rope_qk_inverse_inplace(dqkv: torch.Tensor, rope_cos: torch.Tensor, rope_sin: torch.Tensor, heads: int,
                            seq_len: int) -> torch.Tensor:
    M, _, ld = _rows2d(dqkv, "dqkv")
    check(_lib.load().merlin_rope_qk_inplace_ex(_stream(), dqkv.data_ptr(), ld, rope_cos.data_ptr(),
                                                rope_sin.data_ptr(), None, M, heads, seq_len, 1),
          "merlin_rope_qk_inplace_ex", 1)
    return dqkv


def rmsnorm_bwd(x: torch.Tensor, weight: torch.Tensor, dy: torch.Tensor, dres: Optional[torch.Tensor],
                dweight: Optional[torch.Tensor], eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = rmsnorm'(dy) (+ dres); dweight (fp32 [D]) accumulates.  x, dy, dres, out: contiguous [M, D]."""
    M, D, _ = _rows2d(x, "x")
    for t, n in ((x, "x"), (dy, "dy")):
        if not t.is_contiguous():
            raise ValueError(f"rmsnorm_bwd: {n} must be contiguous")
    if out is None:
        out = torch.empty_like(x)
    if dweight is not None and dweight.dtype != torch.float32:
        raise ValueError("rmsnorm_bwd: dweight accumulates in fp32")
    check(_lib.load().merlin_rmsnorm_bwd(_stream(), x.data_ptr(), weight.data_ptr(), dy.data_ptr(), _ptr(dres),
                                         out.data_ptr(), _ptr(dweight), M, D, float(eps)), "merlin_rmsnorm_bwd", 1)
    return out


def swiglu_fwd(gate_up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, N2, _ = _rows2d(gate_up, "gate_up")
    if out is None:
        out = torch.empty((M, N2 // 2), dtype=_BF16, device=gate_up.device)
    check(_lib.load().merlin_swiglu_fwd(_stream(), gate_up.data_ptr(), out.data_ptr(), M, N2 // 2), "merlin_swiglu_fwd", 1)
    return out


def swiglu_bwd(gate_up: torch.Tensor, dmid: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    M, N2, _ = _rows2d(gate_up, "gate_up")
    if out is None:
        out = torch.empty_like(gate_up)
    check(_lib.load().merlin_swiglu_bwd(_stream(), gate_up.data_ptr(), dmid.data_ptr(), out.data_ptr(), M, N2 // 2),
          "merlin_swiglu_bwd", 1)
    return out


def cross_entropy_with_grad(logits: torch.Tensor, labels: torch.Tensor, batch: int, seq: int,
                            ignore_index: int = -100, grad_scale: float = 1.0):
    """Returns (mean loss, dlogits written IN PLACE over `logits`)."""
    M, V, ld = _rows2d(logits, "logits")
    acc = torch.empty((2,), dtype=torch.float32, device=logits.device)
    L = _lib.load()
    check(L.merlin_cross_entropy_fwd(_stream(), logits.data_ptr(), ld, labels.data_ptr(), batch, seq, V, ignore_index,
                                     None, acc.data_ptr()), "merlin_cross_entropy_fwd", 1)
    check(L.merlin_cross_entropy_bwd(_stream(), logits.data_ptr(), ld, labels.data_ptr(), batch, seq, V, ignore_index,
                                     acc.data_ptr(), float(grad_scale), logits.data_ptr(), ld),
          "merlin_cross_entropy_bwd", 1)
    return acc[0] / acc[1], logits


def embed_splice_bwd(input_ids: torch.Tensor, rank: Optional[torch.Tensor], dx: torch.Tensor,
                     dtable: Optional[torch.Tensor], dfeats: Optional[torch.Tensor], vocab: int) -> None:
    T, D, _ = _rows2d(dx, "dx")
    nfeat = 0 if dfeats is None else dfeats.shape[0]
    check(_lib.load().merlin_embed_splice_bwd(_stream(), input_ids.data_ptr(), _ptr(rank), dx.data_ptr(),
                                              _ptr(dtable), _ptr(dfeats), T, D, nfeat, vocab),
          "merlin_embed_splice_bwd", 1)


def colsum_accumulate(x: torch.Tensor, out_f32: torch.Tensor) -> None:
    M, N, ld = _rows2d(x, "x")
    check(_lib.load().merlin_colsum(_stream(), x.data_ptr(), ld, out_f32.data_ptr(), M, N), "merlin_colsum", 1)


def grad_sumsq(grads: torch.Tensor, out_f32: torch.Tensor) -> None:
    check(_lib.load().merlin_grad_sumsq(_stream(), grads.data_ptr(), grads.numel(), out_f32.data_ptr()),
          "merlin_grad_sumsq", 1)


def adamw_step(params, grads, exp_avg, exp_avg_sq, lr_mult, weight_decay, chunk: int, lr: float, beta1: float,
               beta2: float, eps: float, step: int, grad_scale: float = 1.0, max_grad_norm: float = 0.0,
               sumsq: Optional[torch.Tensor] = None) -> None:
    check(_lib.load().merlin_adamw_step(_stream(), params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(),
                                        exp_avg_sq.data_ptr(), lr_mult.data_ptr(), weight_decay.data_ptr(), chunk,
                                        params.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
                                        float(grad_scale), float(max_grad_norm), _ptr(sumsq)), "merlin_adamw_step", 1)


def f32_to_bf16(x: torch.Tensor, y: torch.Tensor) -> None:
    check(_lib.load().merlin_f32_to_bf16(_stream(), x.data_ptr(), y.data_ptr(), x.numel()), "merlin_f32_to_bf16", 1)


# ----------------------------------------------------------------------------- vision-tower backward ops
def layernorm_bwd(x, weight, dy, dres, dweight, dbias, eps, out=None):
    M, D, _ = _rows2d(x, "x")
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().merlin_layernorm_bwd(_stream(), x.data_ptr(), weight.data_ptr(), dy.data_ptr(), _ptr(dres),
                                           out.data_ptr(), _ptr(dweight), _ptr(dbias), M, D, float(eps)),
          "merlin_layernorm_bwd", 1)
    return out


def quick_gelu(x, dy=None, out=None):
    """Forward (dy None) or backward of x·σ(1.702x) over a contiguous bf16 tensor."""
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().merlin_quick_gelu(_stream(), x.data_ptr(), _ptr(dy), out.data_ptr(), x.numel()),
          "merlin_quick_gelu", 1)
    return out


def conv3x3_col2im(dcol, dfeat, n_img, grid, C, stride, tok_stride, tok_off):
    check(_lib.load().merlin_conv3x3_col2im(_stream(), dcol.data_ptr(), dfeat.data_ptr(), n_img, grid, C, stride,
                                            tok_stride, tok_off), "merlin_conv3x3_col2im", 1)
    return dfeat


def vit_embed(patch_out, class_embedding, position_embedding, n_img):
    tokens, D = position_embedding.shape
    emb = torch.empty((n_img * tokens, D), dtype=_BF16, device=patch_out.device)
    check(_lib.load().merlin_vit_embed(_stream(), patch_out.data_ptr(), class_embedding.data_ptr(),
                                       position_embedding.data_ptr(), emb.data_ptr(), n_img, tokens, D),
          "merlin_vit_embed", 1)
    return emb


def vit_embed_bwd(dhidden, dposition_f32, dclass_f32, n_img):
    tokens, D = dposition_f32.shape
    dpatch = torch.empty((n_img * (tokens - 1), D), dtype=_BF16, device=dhidden.device)
    check(_lib.load().merlin_vit_embed_bwd(_stream(), dhidden.data_ptr(), dpatch.data_ptr(), dposition_f32.data_ptr(),
                                           dclass_f32.data_ptr(), n_img, tokens, D), "merlin_vit_embed_bwd", 1)
    return dpatch
