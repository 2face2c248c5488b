"""torch.autograd.Function wrappers over the library's forward / backward entry points.

SURVEY.md §8b names "the Python shim's torch.autograd.Functions inside MMGPTLlamaForCausalLM.forward" as the
callers of the C ABI: with these, `model(**batch).loss.backward()` — what HF `Trainer.training_step` does with
the reference model (mmgpt/engine/train/train.py:49-59) — puts gradients into `param.grad`, and the operator
seam of mmgpt/utils/llama_flash_attn_monkey_patch.py:20-103 trains through `merlin_flash_attn_fwd / _bwd`.
Every FLOP still runs in libmerlin_b200.so; autograd only records the chain.  `MMGPTTrainer` keeps its own
explicit schedule (flat gradient arena, bucketed all-reduce overlap); both paths launch the same kernels.

Gradient layouts follow the packed parameters (q|k|v fused [3D, D], gate/up interleaved per 128 rows).
`recompute=True` on DecoderLayerFn keeps only the layer input and re-runs the layer forward inside backward:
the activation-recompute policy the reference gets from `--gradient_checkpointing True` (pretrain.sh:35).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import lib as _lib
from . import ops

_BF16 = torch.bfloat16


def _layer_weights(qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w):
    return _lib.LayerWeights(qkv_w.data_ptr(), o_w.data_ptr(), gate_up_w.data_ptr(), down_w.data_ptr(),
                             ln1_w.data_ptr(), ln2_w.data_ptr())


class DecoderLayerFn(Function):
    """x_out = LlamaDecoderLayer(x) (merlin_decoder_layer_forward_train / merlin_decoder_layer_backward)."""

    @staticmethod
    def forward(ctx, x, qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w, cos, sin, seqlens, B, S, H, FF, eps, recompute):
        M, D = x.shape
        x = x.contiguous()
        L = _lib.load()
        nbytes = L.merlin_decoder_layer_saved_bytes(M, D, FF, H, B, S)
        saved = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        x_out = torch.empty_like(x)
        ops.decoder_layer_forward_train(_layer_weights(qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w), x, x_out, saved,
                                        cos, sin, seqlens, B, S, D, H, FF, eps)
        ctx.save_for_backward(qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w, cos, sin)
        ctx.seqlens, ctx.dims, ctx.recompute = seqlens, (B, S, D, H, FF, float(eps)), bool(recompute)
        if recompute:
            ctx.x_in, ctx.block = x, None        # activations are rebuilt in backward
        else:
            ctx.x_in, ctx.block = None, saved
        return x_out

    @staticmethod
    def backward(ctx, dx_out):
        qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w, cos, sin = ctx.saved_tensors
        B, S, D, H, FF, eps = ctx.dims
        M = B * S
        L = _lib.load()
        w = _layer_weights(qkv_w, o_w, gate_up_w, down_w, ln1_w, ln2_w)
        saved = ctx.block
        if saved is None:   # recompute: one more layer forward, into a transient activation block
            saved = torch.empty(L.merlin_decoder_layer_saved_bytes(M, D, FF, H, B, S), dtype=torch.uint8,
                                device=dx_out.device)
            scratch = torch.empty((M, D), dtype=_BF16, device=dx_out.device)
            ops.decoder_layer_forward_train(w, ctx.x_in, scratch, saved, cos, sin, ctx.seqlens, B, S, D, H, FF, eps)
        dx = dx_out.contiguous().clone()             # the library turns dx_out into dx_in in place
        dev = dx.device
        g_qkv, g_o, g_gu, g_down = (torch.empty_like(t) for t in (qkv_w, o_w, gate_up_w, down_w))
        g_ln1 = torch.zeros(D, dtype=torch.float32, device=dev)
        g_ln2 = torch.zeros(D, dtype=torch.float32, device=dev)
        ws = torch.empty(L.merlin_decoder_layer_backward_workspace_bytes(M, D, FF, H, B, S), dtype=torch.uint8, device=dev)
        grads = _lib.LayerGrads(g_qkv.data_ptr(), g_o.data_ptr(), g_gu.data_ptr(), g_down.data_ptr(),
                                g_ln1.data_ptr(), g_ln2.data_ptr())
        ops.decoder_layer_backward(w, grads, saved, dx, ws, cos, sin, ctx.seqlens, B, S, D, H, FF, eps, False)
        ctx.block = None
        b1, b2 = torch.empty(D, dtype=_BF16, device=dev), torch.empty(D, dtype=_BF16, device=dev)
        ops.f32_to_bf16(g_ln1, b1)
        ops.f32_to_bf16(g_ln2, b2)
        return (dx, g_qkv, g_o, g_gu, g_down, b1, b2) + (None,) * 9


class RMSNormFn(Function):
    """LlamaRMSNorm (merlin_rmsnorm / merlin_rmsnorm_bwd)."""

    @staticmethod
    def forward(ctx, x, weight, eps):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.eps = float(eps)
        return ops.rmsnorm(x, weight, eps)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        gw = torch.zeros(weight.numel(), dtype=torch.float32, device=x.device)
        dx = ops.rmsnorm_bwd(x, weight, dy.contiguous(), None, gw, ctx.eps)
        out = torch.empty_like(weight)
        ops.f32_to_bf16(gw, out)
        return dx, out, None


class LMHeadCEFn(Function):
    """(logits, loss) = lm_head + shifted cross-entropy (merlin_lmhead_ce_forward / _backward).  Only the loss is
    differentiable; use LinearFn for a custom loss on the logits."""

    @staticmethod
    def forward(ctx, hidden, lm_head_w, labels, B, S, ignore_index):
        hidden = hidden.contiguous()
        logits, loss, acc = ops.lmhead_ce_forward(hidden, lm_head_w, labels, B, S, ignore_index)
        ctx.save_for_backward(hidden, lm_head_w, labels, logits, acc)
        ctx.dims = (B, S, int(ignore_index))
        ctx.mark_non_differentiable(logits)
        ctx.set_materialize_grads(False)
        return logits, loss

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        hidden, lm_head_w, labels, logits, acc = ctx.saved_tensors
        if dloss is None:
            return None, None, None, None, None, None
        B, S, ignore_index = ctx.dims
        V, D = lm_head_w.shape
        buf = torch.empty((logits.shape[0], (V + 7) // 8 * 8), dtype=_BF16, device=logits.device)[:, :V]
        buf.copy_(logits)                            # the caller still holds out.logits: work on a copy
        need_w = ctx.needs_input_grad[1]
        g_w = torch.empty_like(lm_head_w) if need_w else None
        dh = ops.lmhead_ce_backward(buf, labels, acc, 1.0, hidden, lm_head_w, g_w, False, B, S, ignore_index)
        scale = dloss.to(_BF16)
        dh.mul_(scale)                               # chain rule for an upstream gradient other than 1
        if g_w is not None:
            g_w.mul_(scale)
        return dh, g_w, None, None, None, None


class LinearFn(Function):
    """y = x · Wᵀ (+ bias): merlin_gemm_bf16 forward, dgrad / wgrad through merlin_gemm_bf16_ex."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return ops.gemm(x, weight, bias=bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        N = weight.shape[0]
        if dy.stride(1) != 1 or dy.stride(0) % 8 != 0:
            buf = torch.empty((dy.shape[0], (N + 7) // 8 * 8), dtype=_BF16, device=dy.device)[:, :N]
            buf.copy_(dy)
            dy = buf
        dx = ops.gemm_nt(dy, weight) if ctx.needs_input_grad[0] else None
        gw = ops.gemm_tn(dy, x) if ctx.needs_input_grad[1] else None
        if gw is not None and not gw.is_contiguous():
            gw = gw.contiguous()
        gb = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            acc = torch.zeros(N, dtype=torch.float32, device=dy.device)
            ops.colsum_accumulate(dy, acc)
            gb = torch.empty(N, dtype=_BF16, device=dy.device)
            ops.f32_to_bf16(acc, gb)
        return dx, gw, gb


class EmbedSpliceFn(Function):
    """inputs_embeds = embed_tokens(ids) with projected image rows spliced over <im_patch> positions
    (merlin_embed_splice / merlin_embed_splice_bwd; base_mmgpt.py:99-160)."""

    @staticmethod
    def forward(ctx, embed_table, feats, ids_dev, im_patch_token, feat_offset, feat_limit):
        x, _, rank = ops.embed_splice(ids_dev, embed_table, feats, im_patch_token, feat_offset, feat_limit,
                                      return_rank=True)
        ctx.save_for_backward(ids_dev, rank)
        ctx.shapes = (tuple(embed_table.shape), None if feats is None else tuple(feats.shape))
        return x

    @staticmethod
    def backward(ctx, dx):
        ids_dev, rank = ctx.saved_tensors
        tshape, fshape = ctx.shapes
        dev = dx.device
        dtable = torch.zeros(tshape, dtype=_BF16, device=dev) if ctx.needs_input_grad[0] else None
        dfeats = torch.zeros(fshape, dtype=_BF16, device=dev) if (fshape is not None and ctx.needs_input_grad[1]) else None
        if dtable is not None or dfeats is not None:
            ops.embed_splice_bwd(ids_dev.view(-1), rank, dx.contiguous(), dtable, dfeats, tshape[0])
        return dtable, dfeats, None, None, None, None


class ConvProjectorFn(Function):
    """ConvProjector.forward (conv_projector.py:23-39) as im2col + GEMM + bias, with its backward."""

    @staticmethod
    def forward(ctx, hidden_v, weight, bias, n_img, grid, C, stride, tok_stride, tok_off):
        col = ops.conv3x3_im2col(hidden_v, n_img, grid, C, stride, tok_stride, tok_off)
        ctx.save_for_backward(col, weight)
        ctx.geom = (n_img, grid, C, stride, tok_stride, tok_off, tuple(hidden_v.shape))
        return ops.gemm(col, weight, bias=bias)

    @staticmethod
    def backward(ctx, dfeats):
        col, weight = ctx.saved_tensors
        n_img, grid, C, stride, tok_stride, tok_off, hshape = ctx.geom
        dfeats = dfeats.contiguous()
        dev = dfeats.device
        gw = ops.gemm_tn(dfeats, col) if ctx.needs_input_grad[1] else None
        gb = None
        if ctx.needs_input_grad[2]:
            acc = torch.zeros(weight.shape[0], dtype=torch.float32, device=dev)
            ops.colsum_accumulate(dfeats, acc)
            gb = torch.empty(weight.shape[0], dtype=_BF16, device=dev)
            ops.f32_to_bf16(acc, gb)
        dh = None
        if ctx.needs_input_grad[0]:
            dcol = ops.gemm_nt(dfeats, weight)
            dh = torch.zeros(hshape, dtype=_BF16, device=dev)
            ops.conv3x3_col2im(dcol.contiguous(), dh, n_img, grid, C, stride, tok_stride, tok_off)
        return (dh, gw, gb) + (None,) * 6


class VisionTowerFn(Function):
    """hidden_states[select_layer] of the CLIP tower (merlin_vit_forward_train / merlin_vit_backward)."""

    @staticmethod
    def forward(ctx, tower, pixels, *params):
        hidden, saved = tower.encode_train(pixels)
        ctx.tower, ctx.saved = tower, saved
        ctx.names = [n for n, _ in tower.named_parameters_reference()]
        ctx.set_materialize_grads(False)
        return hidden

    @staticmethod
    def backward(ctx, dh):
        tower = ctx.tower
        named = tower.named_parameters_reference()
        if dh is None:
            return (None, None) + (None,) * len(named)
        dev = dh.device
        bf, f32 = {}, {}
        for n, p in named:
            if p.dim() == 1 or n.endswith("position_embedding.weight"):
                f32[n] = torch.zeros(p.numel(), dtype=torch.float32, device=dev)
            else:
                bf[n] = torch.empty_like(p)
        tower.backward(dh.contiguous().clone(), ctx.saved, lambda n: bf[n], f32, False)
        ctx.saved = None
        out = []
        for n, p in named:
            if n in bf:
                out.append(bf[n])
            else:
                g = torch.empty(p.shape, dtype=_BF16, device=dev)
                ops.f32_to_bf16(f32[n], g)
                out.append(g)
        return (None, None) + tuple(out)


class FlashAttnFn(Function):
    """Causal / bidirectional flash attention over a packed [B*S, 3*H*hd] qkv buffer
    (merlin_flash_attn_fwd / merlin_flash_attn_bwd) — the operator behind the HF attention seam
    (flash_attn_varlen_qkvpacked_func in monkey_patch.py:82-95)."""

    @staticmethod
    def forward(ctx, qkv, seqlens, B, S, H, hd, causal, softmax_scale):
        qkv = qkv.contiguous()
        out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=causal, seqlens=seqlens, return_lse=True,
                                  softmax_scale=softmax_scale)
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (seqlens, B, S, H, hd, bool(causal), float(softmax_scale))
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        seqlens, B, S, H, hd, causal, scale = ctx.meta
        dqkv = ops.flash_attn_bwd(qkv, out, dout.contiguous(), lse, B, S, H, hd, causal=causal, seqlens=seqlens,
                                  softmax_scale=scale)
        return (dqkv,) + (None,) * 7
