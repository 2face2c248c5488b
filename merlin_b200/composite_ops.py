"""Wrappers of the composite per-layer entry points (merlin_decoder_layer_forward /
merlin_vit_layer_forward): one library call sequences the 7 kernels of a layer."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import lib as _lib


def decoder_layer_workspace(M: int, D: int, FF: int, device) -> torch.Tensor:
    n = _lib.load().merlin_decoder_layer_workspace_bytes(M, D, FF)
    return torch.empty(n, dtype=torch.uint8, device=device)


def decoder_layer_forward(x, ws, qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w, cos, sin, seqlens, kv, batch, seq,
                          heads, ff, eps, cta_group: int = 2, profile=None):
    from . import ops
    D = x.shape[1]
    w = _lib.LayerWeights(qkv_w.data_ptr(), o_w.data_ptr(), gate_up_w.data_ptr(), down_w.data_ptr(),
                          ln1_w.data_ptr(), ln2_w.data_ptr())
    kv_ptr, cap = (None, 0) if kv is None else (kv.data_ptr(), kv.shape[1] * 32)  # [B, capacity/32, H, 32, 256]
    fn = _lib.load().merlin_decoder_layer_forward
    args = (torch.cuda.current_stream().cuda_stream, ctypes.byref(w), x.data_ptr(), ws.data_ptr(), cos.data_ptr(),
            sin.data_ptr(), None if seqlens is None else seqlens.data_ptr(), kv_ptr, cap, batch, seq, D, heads, ff,
            float(eps), cta_group)
    ops.PROFILE.launches += 7 + (1 if (kv is not None and batch != 1 and seq % 128 != 0) else 0)   # KV append rides in the QKV epilogue
    if ops.PROFILE.enabled:
        # the roofline leg times GEMM launches individually: fall back to the per-op path for this layer
        return False
    ops.check(fn(*args), "merlin_decoder_layer_forward")
    return True


def vit_layer_workspace(M: int, d: int, ff: int, device) -> torch.Tensor:
    n = _lib.load().merlin_vit_layer_workspace_bytes(M, d, ff)
    return torch.empty(n, dtype=torch.uint8, device=device)


def vit_layer_forward(h, ws, tower, j: int, n_img: int, cta_group: int = 2):
    from . import ops
    c = tower.config
    w = _lib.VitLayerWeights(*[t.data_ptr() for t in (tower.qkv_w[j], tower.qkv_b[j], tower.out_w[j], tower.out_b[j],
                                                      tower.ln1_w[j], tower.ln1_b[j], tower.ln2_w[j], tower.ln2_b[j],
                                                      tower.fc1_w[j], tower.fc1_b[j], tower.fc2_w[j], tower.fc2_b[j])])
    ops.PROFILE.launches += 7
    ops.check(_lib.load().merlin_vit_layer_forward(torch.cuda.current_stream().cuda_stream, ctypes.byref(w),
                                                   h.data_ptr(), ws.data_ptr(), n_img, tower.tokens, c.hidden_size,
                                                   c.num_attention_heads, c.intermediate_size,
                                                   float(c.layer_norm_eps), cta_group), "merlin_vit_layer_forward")


def decoder_layer_forward_train(w, x, x_out, saved, cos, sin, seqlens, B, S, D, H, FF, eps, cta_group: int = 2):
    from . import ops
    ops.PROFILE.launches += 8
    ops.check(_lib.load().merlin_decoder_layer_forward_train(
        torch.cuda.current_stream().cuda_stream, ctypes.byref(w), x.data_ptr(), x_out.data_ptr(), saved.data_ptr(),
        cos.data_ptr(), sin.data_ptr(), None if seqlens is None else seqlens.data_ptr(), B, S, D, H, FF, float(eps),
        cta_group), "merlin_decoder_layer_forward_train")


def decoder_layer_backward(w, grads, saved, dx, ws, cos, sin, seqlens, B, S, D, H, FF, eps, accumulate,
                           cta_group: int = 2):
    from . import ops
    ops.PROFILE.launches += 17
    ops.check(_lib.load().merlin_decoder_layer_backward(
        torch.cuda.current_stream().cuda_stream, ctypes.byref(w), ctypes.byref(grads), saved.data_ptr(), dx.data_ptr(),
        ws.data_ptr(), cos.data_ptr(), sin.data_ptr(), None if seqlens is None else seqlens.data_ptr(), B, S, D, H, FF,
        float(eps), int(bool(accumulate)), cta_group), "merlin_decoder_layer_backward")
