// merlin_b200 — composite entry points: one C call per decoder / CLIP encoder layer forward.
// They only sequence the kernels of gemm.cu / attn.cu / elementwise.cu on the caller's stream with a
// caller-provided workspace (SURVEY.md §8b "minimum export set": merlin_decoder_layer_forward, ...).
#include "host.cuh"

#include <cuda_bf16.h>

extern "C" {

size_t merlin_decoder_layer_workspace_bytes(int M, int D, int FF) {
  // xn [M,D] + qkv [M,3D] + att [M,D] + mid [M,FF]   (bf16)
  return (static_cast<size_t>(M) * (5 * static_cast<size_t>(D) + FF)) * 2 + 256;
}

int merlin_decoder_layer_forward(void* stream, const merlin_layer_weights* w, void* x, void* workspace,
                                 const float* rope_cos, const float* rope_sin, const int* seqlens,
                                 void* kv_cache, int64_t kv_capacity, int B, int S, int D, int H, int FF, float eps,
                                 int cta_group) {
  MB_CHECK_ARG(w && x && workspace && rope_cos && rope_sin, "decoder_layer_forward: null argument");
  MB_CHECK_ARG(D == H * 128 && FF % 128 == 0, "decoder_layer_forward: needs head_dim 128 and FF %% 128 == 0");
  const int M = B * S;
  __nv_bfloat16* xn = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* qkv = xn + static_cast<size_t>(M) * D;
  __nv_bfloat16* att = qkv + static_cast<size_t>(M) * 3 * D;
  __nv_bfloat16* mid = att + static_cast<size_t>(M) * D;
  int rc;
  // x + Attn(RMSNorm(x))          (HF:llama/modeling_llama.py:303-333, monkey_patch.py:20-103)
  if ((rc = merlin_rmsnorm(stream, x, D, w->ln1_w, xn, D, M, D, eps)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_qkv_rope(stream, xn, D, w->qkv_w, D, qkv, 3 * D, rope_cos, rope_sin, nullptr, M, 3 * D, D,
                                 2 * D, S, cta_group)) != MERLIN_OK)
    return rc;
  if (kv_cache != nullptr) {
    // rotated keys and values of the S prompt tokens -> head-major cache [B, H, kv_capacity, 256]
    if ((rc = merlin_kv_append(stream, qkv, 3 * D, kv_cache, kv_capacity, B, S, D, H, 0)) != MERLIN_OK) return rc;
  }
  if ((rc = merlin_flash_attn_fwd(stream, qkv, 3 * D, att, D, nullptr, seqlens, B, S, H, 128, 1, 0.f)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16(stream, att, D, w->o_w, D, x, D, x, D, nullptr, M, D, D, MERLIN_EPI_RESIDUAL,
                             cta_group)) != MERLIN_OK)
    return rc;
  // x + MLP(RMSNorm(x))           (HF:llama/modeling_llama.py:182-184)
  if ((rc = merlin_rmsnorm(stream, x, D, w->ln2_w, xn, D, M, D, eps)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_bf16(stream, xn, D, w->gate_up_w, D, mid, FF, nullptr, 0, nullptr, M, 2 * FF, D,
                             MERLIN_EPI_SWIGLU, cta_group)) != MERLIN_OK)
    return rc;
  return merlin_gemm_bf16(stream, mid, FF, w->down_w, FF, x, D, x, D, nullptr, M, D, FF, MERLIN_EPI_RESIDUAL,
                          cta_group);
}

size_t merlin_vit_layer_workspace_bytes(int M, int d, int ff) {
  // xn [M,d] + qkv [M,3d] + att [M,d] + mid [M,ff]
  return (static_cast<size_t>(M) * (5 * static_cast<size_t>(d) + ff)) * 2 + 256;
}

int merlin_vit_layer_forward(void* stream, const merlin_vit_layer_weights* w, void* h, void* workspace, int n_img,
                             int tokens, int d, int heads, int ff, float eps, int cta_group) {
  MB_CHECK_ARG(w && h && workspace, "vit_layer_forward: null argument");
  MB_CHECK_ARG(d == heads * 64, "vit_layer_forward: needs head_dim 64");
  const int M = n_img * tokens;
  __nv_bfloat16* xn = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* qkv = xn + static_cast<size_t>(M) * d;
  __nv_bfloat16* att = qkv + static_cast<size_t>(M) * 3 * d;
  __nv_bfloat16* mid = att + static_cast<size_t>(M) * d;
  int rc;
  // HF:clip/modeling_clip.py:363-384
  if ((rc = merlin_layernorm(stream, h, d, w->ln1_w, w->ln1_b, xn, d, M, d, eps)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_bf16(stream, xn, d, w->qkv_w, d, qkv, 3 * d, nullptr, 0, w->qkv_b, M, 3 * d, d,
                             MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_flash_attn_fwd(stream, qkv, 3 * d, att, d, nullptr, nullptr, n_img, tokens, heads, 64, 0, 0.f)) !=
      MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16(stream, att, d, w->out_w, d, h, d, h, d, w->out_b, M, d, d, MERLIN_EPI_RESIDUAL,
                             cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_layernorm(stream, h, d, w->ln2_w, w->ln2_b, xn, d, M, d, eps)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_bf16(stream, xn, d, w->fc1_w, d, mid, ff, nullptr, 0, w->fc1_b, M, ff, d,
                             MERLIN_EPI_QUICKGELU, cta_group)) != MERLIN_OK)
    return rc;
  return merlin_gemm_bf16(stream, mid, ff, w->fc2_w, ff, h, d, h, d, w->fc2_b, M, d, ff, MERLIN_EPI_RESIDUAL,
                          cta_group);
}

}  // extern "C"

// =====================================================================================================
// Training forms: one call per decoder layer for forward-with-save and for backward.  The caller owns
// the saved-activation buffers (persistent across steps) and a workspace; the schedule is the one of
// mmgpt/engine/train/trainer.py (merlin_b200) and autograd of HF LlamaDecoderLayer in the reference.
// =====================================================================================================
extern "C" {

size_t merlin_decoder_layer_saved_bytes(int M, int D, int FF, int H, int B, int S) {
  // x_in, xn1, att, x_mid, xn2 [M,D]; qkv [M,3D]; gu [M,2FF]; mid [M,FF]  (bf16) + lse [B,H,S] fp32
  const size_t bf16 = static_cast<size_t>(M) * (5 * static_cast<size_t>(D) + 3 * D + 3 * FF) * 2;
  return bf16 + static_cast<size_t>(B) * H * S * 4 + 1024;
}

static void layout_saved(void* saved, int M, int D, int FF, merlin_layer_saved* s) {
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(saved);
  const size_t md = static_cast<size_t>(M) * D;
  s->x_in = p;              p += md;
  s->xn1 = p;               p += md;
  s->qkv = p;               p += 3 * md;
  s->att = p;               p += md;
  s->x_mid = p;             p += md;
  s->xn2 = p;               p += md;
  s->gate_up = p;           p += static_cast<size_t>(M) * 2 * FF;
  s->mid = p;               p += static_cast<size_t>(M) * FF;
  s->lse = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 255) & ~static_cast<uintptr_t>(255));
}

int merlin_decoder_layer_forward_train(void* stream, const merlin_layer_weights* w, const void* x, void* x_out,
                                       void* saved, const float* rope_cos, const float* rope_sin,
                                       const int* seqlens, int B, int S, int D, int H, int FF, float eps,
                                       int cta_group) {
  MB_CHECK_ARG(w && x && x_out && saved && rope_cos && rope_sin, "decoder_layer_forward_train: null argument");
  MB_CHECK_ARG(D == H * 128 && FF % 128 == 0, "decoder_layer_forward_train: needs head_dim 128, FF %% 128 == 0");
  const int M = B * S;
  merlin_layer_saved s;
  layout_saved(saved, M, D, FF, &s);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (x != s.x_in)  // the previous layer can write its output straight into this layer's x_in slot
    MB_CHECK_CUDA(cudaMemcpyAsync(s.x_in, x, static_cast<size_t>(M) * D * 2, cudaMemcpyDeviceToDevice, st));
  if ((rc = merlin_rmsnorm(stream, s.x_in, D, w->ln1_w, s.xn1, D, M, D, eps)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_qkv_rope(stream, s.xn1, D, w->qkv_w, D, s.qkv, 3 * D, rope_cos, rope_sin, nullptr, M, 3 * D,
                                 D, 2 * D, S, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_flash_attn_fwd(stream, s.qkv, 3 * D, s.att, D, s.lse, seqlens, B, S, H, 128, 1, 0.f)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16(stream, s.att, D, w->o_w, D, s.x_mid, D, s.x_in, D, nullptr, M, D, D,
                             MERLIN_EPI_RESIDUAL, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_rmsnorm(stream, s.x_mid, D, w->ln2_w, s.xn2, D, M, D, eps)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_bf16(stream, s.xn2, D, w->gate_up_w, D, s.gate_up, 2 * FF, nullptr, 0, nullptr, M, 2 * FF, D,
                             MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_swiglu_fwd(stream, s.gate_up, s.mid, M, FF)) != MERLIN_OK) return rc;
  return merlin_gemm_bf16(stream, s.mid, FF, w->down_w, FF, x_out, D, s.x_mid, D, nullptr, M, D, FF,
                          MERLIN_EPI_RESIDUAL, cta_group);
}

size_t merlin_decoder_layer_backward_workspace_bytes(int M, int D, int FF, int H, int B, int S) {
  // dmid [M,FF] + dxn [M,D] + dx_mid [M,D] + datt [M,D] + dqkv [M,3D] (bf16) + attention-backward scratch
  return static_cast<size_t>(M) * (static_cast<size_t>(FF) + 6 * D) * 2 +
         merlin_flash_attn_bwd_workspace_bytes(B, S, H, 128) + 2048;
}

int merlin_decoder_layer_backward(void* stream, const merlin_layer_weights* w, const merlin_layer_grads* g,
                                  void* saved, void* dx, void* workspace, const float* rope_cos,
                                  const float* rope_sin, const int* seqlens, int B, int S, int D, int H, int FF,
                                  float eps, int accumulate, int cta_group) {
  MB_CHECK_ARG(w && g && saved && dx && workspace, "decoder_layer_backward: null argument");
  const int M = B * S;
  merlin_layer_saved s;
  layout_saved(saved, M, D, FF, &s);
  __nv_bfloat16* dmid = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* dxn = dmid + static_cast<size_t>(M) * FF;
  __nv_bfloat16* dx_mid = dxn + static_cast<size_t>(M) * D;
  __nv_bfloat16* datt = dx_mid + static_cast<size_t>(M) * D;
  __nv_bfloat16* dqkv = datt + static_cast<size_t>(M) * D;
  void* attn_ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(dqkv + static_cast<size_t>(M) * 3 * D) + 255) &
                                          ~static_cast<uintptr_t>(255));
  const int epi = accumulate ? MERLIN_EPI_RESIDUAL : MERLIN_EPI_STORE;
  int rc;
  // ---- MLP: x_out = x_mid + down(silu(g)·u)
  if ((rc = merlin_gemm_bf16_ex(stream, dx, D, 0, w->down_w, FF, 1, dmid, FF, nullptr, 0, nullptr, M, FF, D,
                                MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16_ex(stream, dx, D, 1, s.mid, FF, 1, g->down_w, FF, g->down_w, FF, nullptr, D, FF, M, epi,
                                cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_swiglu_bwd(stream, s.gate_up, dmid, s.gate_up, M, FF)) != MERLIN_OK) return rc;   // dgu in place
  if ((rc = merlin_gemm_bf16_ex(stream, s.gate_up, 2 * FF, 1, s.xn2, D, 1, g->gate_up_w, D, g->gate_up_w, D, nullptr,
                                2 * FF, D, M, epi, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16_ex(stream, s.gate_up, 2 * FF, 0, w->gate_up_w, D, 1, dxn, D, nullptr, 0, nullptr, M, D,
                                2 * FF, MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_rmsnorm_bwd(stream, s.x_mid, w->ln2_w, dxn, dx, dx_mid, g->ln2_w_f32, M, D, eps)) != MERLIN_OK)
    return rc;
  // ---- attention: x_mid = x_in + o_proj(attn(rope(qkv(norm(x_in)))))
  if ((rc = merlin_gemm_bf16_ex(stream, dx_mid, D, 0, w->o_w, D, 1, datt, D, nullptr, 0, nullptr, M, D, D,
                                MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16_ex(stream, dx_mid, D, 1, s.att, D, 1, g->o_w, D, g->o_w, D, nullptr, D, D, M, epi,
                                cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_flash_attn_bwd(stream, s.qkv, 3 * D, s.att, datt, D, s.lse, seqlens, dqkv, 3 * D, attn_ws, B, S, H,
                                  128, 1, 0.f)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_rope_qk_inplace_ex(stream, dqkv, 3 * D, rope_cos, rope_sin, nullptr, M, H, S, 1)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16_ex(stream, dqkv, 3 * D, 1, s.xn1, D, 1, g->qkv_w, D, g->qkv_w, D, nullptr, 3 * D, D, M,
                                epi, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_gemm_bf16_ex(stream, dqkv, 3 * D, 0, w->qkv_w, D, 1, dxn, D, nullptr, 0, nullptr, M, D, 3 * D,
                                MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  // dx (in place) = dx_mid + rmsnorm'(dxn)
  return merlin_rmsnorm_bwd(stream, s.x_in, w->ln1_w, dxn, dx_mid, dx, g->ln1_w_f32, M, D, eps);
}

// ---- lm_head + shifted cross-entropy (llama_mmgpt.py:86-100) as one call each way
int merlin_lmhead_ce_forward(void* stream, const void* hidden, int64_t ldh, const void* lm_head_w, void* logits,
                             int64_t ldl, const int64_t* labels, int B, int S, int D, int V, int64_t ignore_index,
                             float* row_nll, float* loss_sum_count, int cta_group) {
  MB_CHECK_ARG(hidden && lm_head_w && logits, "lmhead_ce_forward: null argument");
  MB_CHECK_ARG(labels == nullptr || loss_sum_count != nullptr, "lmhead_ce_forward: labels need loss_sum_count");
  const int M = B * S;
  int rc;
  if ((rc = merlin_gemm_bf16(stream, hidden, ldh, lm_head_w, D, logits, ldl, nullptr, 0, nullptr, M, V, D,
                             MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if (labels == nullptr) return MERLIN_OK;
  return merlin_cross_entropy_fwd(stream, logits, ldl, labels, B, S, V, ignore_index, row_nll, loss_sum_count);
}

// logits (in) -> dlogits (in place); d_lm_head (+)= dlogits^T . hidden; dhidden = dlogits . lm_head
int merlin_lmhead_ce_backward(void* stream, void* logits, int64_t ldl, const int64_t* labels, int B, int S, int D, int V,
                              int64_t ignore_index, const float* loss_sum_count, float grad_scale, const void* hidden,
                              int64_t ldh, const void* lm_head_w, void* d_lm_head_w, int accumulate, void* dhidden,
                              int64_t lddh, int cta_group) {
  MB_CHECK_ARG(logits && labels && loss_sum_count && hidden && lm_head_w && dhidden, "lmhead_ce_backward: null argument");
  const int M = B * S;
  int rc;
  if ((rc = merlin_cross_entropy_bwd(stream, logits, ldl, labels, B, S, V, ignore_index, loss_sum_count, grad_scale,
                                     logits, ldl)) != MERLIN_OK)
    return rc;
  if (d_lm_head_w != nullptr) {   // wgrad: [V, D] = dlogits[M, V]^T . hidden[M, D]
    if ((rc = merlin_gemm_bf16_ex(stream, logits, ldl, 1, hidden, ldh, 1, d_lm_head_w, D, accumulate ? d_lm_head_w : nullptr,
                                  D, nullptr, V, D, M, accumulate ? MERLIN_EPI_RESIDUAL : MERLIN_EPI_STORE,
                                  cta_group)) != MERLIN_OK)
      return rc;
  }
  // dgrad: [M, D] = dlogits[M, V] . lm_head[V, D]
  return merlin_gemm_bf16_ex(stream, logits, ldl, 0, lm_head_w, D, 1, dhidden, lddh, nullptr, 0, nullptr, M, D, V,
                             MERLIN_EPI_STORE, cta_group);
}

// ---- ConvProjector.forward (conv_projector.py:23-39) as one call: im2col of the CLS-dropping slice + GEMM + bias
size_t merlin_projector_workspace_bytes(int n_img, int grid, int C, int stride) {
  const int go = (grid + 2 - 3) / stride + 1;
  return static_cast<size_t>(n_img) * go * go * 9 * C * 2 + 256;
}

int merlin_projector_forward(void* stream, const void* feat, const void* weight, const void* bias, void* workspace,
                             void* out, int64_t ldo, int n_img, int grid, int C, int D_out, int stride, int tok_stride,
                             int tok_off, int cta_group) {
  MB_CHECK_ARG(feat && weight && workspace && out, "projector_forward: null argument");
  MB_CHECK_ARG(stride >= 1 && grid >= 1 && C % 8 == 0, "projector_forward: bad geometry");
  const int go = (grid + 2 - 3) / stride + 1;
  int rc;
  if ((rc = merlin_conv3x3_im2col(stream, feat, workspace, n_img, grid, C, stride, tok_stride, tok_off)) != MERLIN_OK)
    return rc;
  return merlin_gemm_bf16(stream, workspace, 9 * C, weight, 9 * C, out, ldo, nullptr, 0, bias, n_img * go * go, D_out,
                          9 * C, MERLIN_EPI_STORE, cta_group);
}

}  // extern "C"
