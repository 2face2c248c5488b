"""Operator seam: plug the B200 flash-attention kernel into HF Llama attention.

Mirror of mmgpt/utils/llama_flash_attn_monkey_patch.py:115-125 (`replace_llama_attn_with_flash_attn`),
which the reference calls before importing its train module (mmgpt/engine/train/train_flash_attn.py:6-8)
to swap HF's attention for flash-attn-2 and to make the 2-D padding mask reach the attention op.

Under the installed transformers 5.x the 4.31 hooks (`LlamaAttention.forward` signature,
`LlamaModel._prepare_decoder_attention_mask`) no longer exist; the equivalent seam is
the attention-function registry that `LlamaAttention.forward` dispatches through
(HF:llama/modeling_llama.py:272-274).  `replace_llama_attn_with_flash_attn()` registers
`"merlin_b200"` there and makes it the implementation HF Llama models pick up; the attention math
then runs in merlin_flash_attn_fwd (causal, scale 1/sqrt(head_dim), no dropout, right-padding key mask,
padded query rows zero-filled — the semantics of monkey_patch.py:63-102).
"""
from __future__ import annotations

from typing import Optional

import torch

from ... import ops

IMPLEMENTATION_NAME = "merlin_b200"


def merlin_attention_forward(module, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                             attention_mask: Optional[torch.Tensor], dropout: float = 0.0,
                             scaling: Optional[float] = None, **kwargs):
    """query/key/value: [B, H, S, hd] (already rotated).  Returns ([B, S, H, hd], None)."""
    assert dropout == 0.0, "attention dropout is not supported"          # monkey_patch.py:82 passes 0.0
    B, H, S, hd = query.shape
    if key.shape[1] != H or key.shape[2] != S:
        raise AssertionError("past_key_value / GQA is not supported")    # monkey_patch.py:54
    if query.dtype != torch.bfloat16 or not query.is_cuda:
        raise TypeError("merlin_b200 attention needs CUDA bfloat16 tensors")
    seqlens = None
    if attention_mask is not None and attention_mask.dim() == 2:         # [B, S] key padding mask
        seqlens = attention_mask.sum(-1).to(torch.int32)
    # pack to the [B*S, 3*H*hd] layout the kernel reads through one TMA tensor map
    qkv = torch.stack((query, key, value), dim=1).permute(0, 3, 1, 2, 4).reshape(B * S, 3 * H * hd)
    scale = float(scaling) if scaling else 0.0
    if torch.is_grad_enabled() and qkv.requires_grad:
        # training through the seam (train_flash_attn.py:6-8 exists to TRAIN with the patched attention):
        # merlin_flash_attn_fwd saves the log-sum-exp, merlin_flash_attn_bwd returns the packed dq|dk|dv and
        # autograd carries it back through the stack / permute into HF's q, k, v projections
        from ...autograd import FlashAttnFn
        out = FlashAttnFn.apply(qkv, seqlens, B, S, H, hd, True, scale)
    else:
        out = ops.flash_attn(qkv.contiguous(), B, S, H, hd, causal=True, seqlens=seqlens, softmax_scale=scale)
    return out.view(B, S, H, hd), None


def replace_llama_attn_with_flash_attn():
    """Same entry point name as the reference's patch.  Call before building the HF model, then
    construct it with `attn_implementation="merlin_b200"` (or set config._attn_implementation)."""
    cap = torch.cuda.get_device_capability()
    if cap[0] != 10:
        raise RuntimeError(f"merlin_b200 kernels are built for sm_100a; found compute capability {cap}")
    from transformers import AttentionInterface, AttentionMaskInterface
    from transformers.masking_utils import flash_attention_mask
    AttentionInterface.register(IMPLEMENTATION_NAME, merlin_attention_forward)
    # the counterpart of the reference's `_prepare_decoder_attention_mask -> identity` patch
    # (monkey_patch.py:108-112): hand the 2-D key-padding mask to the attention op untouched
    AttentionMaskInterface.register(IMPLEMENTATION_NAME, flash_attention_mask)
    return IMPLEMENTATION_NAME
