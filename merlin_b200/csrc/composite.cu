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
  // rotated keys and values of the S prompt tokens go to the cache from the GEMM's own epilogue (a second TMA store
  // of the staged chunk); a batch whose samples are not tile-aligned falls back to the gather kernel
  const bool fuse_kv = kv_cache != nullptr && (B == 1 || S % 128 == 0);
  if ((rc = merlin_gemm_qkv_rope_kv(stream, xn, D, w->qkv_w, D, qkv, 3 * D, rope_cos, rope_sin, nullptr, M, 3 * D, D,
                                    2 * D, S, fuse_kv ? kv_cache : nullptr, kv_capacity, B, H, cta_group)) != MERLIN_OK)
    return rc;
  if (kv_cache != nullptr && !fuse_kv) {
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

// =====================================================================================================
// Whole CLIP tower as composite entry points (SURVEY.md §8b minimum set: merlin_vit_forward / backward):
// patch embedding -> CLS / position / pre_layrnorm -> the encoder layers `select_layer` needs, and the
// explicit backward of the same schedule.  Reference: mmgpt/model/vision_encoder/clip_encoder.py:74-82 over
// HF CLIPVisionModel (HF:clip/modeling_clip.py:202-218, 363-384, 676-677) and its autograd.
// =====================================================================================================
namespace {

struct VitDims {
  int M, G, d, ff, T, kpp;
};

VitDims vit_dims(const merlin_vit_weights* w, int n_img) {
  VitDims v;
  v.G = w->image_size / w->patch_size;
  v.T = v.G * v.G + 1;
  v.M = n_img * v.T;
  v.d = w->hidden;
  v.ff = w->intermediate;
  v.kpp = w->k_patch_padded;
  return v;
}

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// saved-activation layout of merlin_vit_forward_train
struct VitSaved {
  __nv_bfloat16 *col, *emb;
  struct Layer {
    __nv_bfloat16 *h, *xn1, *qkv, *att, *h_mid, *xn2, *pre, *mid;
    float* lse;
  };
};

size_t vit_layer_saved_bytes(const VitDims& v, int n_img, int heads) {
  // h, xn1, att, h_mid, xn2 [M,d]; qkv [M,3d]; pre, mid [M,ff] (bf16); lse [n,heads,T] fp32
  return align256(static_cast<size_t>(v.M) * (8 * static_cast<size_t>(v.d) + 2 * v.ff) * 2) +
         align256(static_cast<size_t>(n_img) * heads * v.T * 4);
}

void vit_layer_layout(uint8_t* base, const VitDims& v, VitSaved::Layer* L) {
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base);
  const size_t md = static_cast<size_t>(v.M) * v.d, mf = static_cast<size_t>(v.M) * v.ff;
  L->h = p;      p += md;
  L->xn1 = p;    p += md;
  L->qkv = p;    p += 3 * md;
  L->att = p;    p += md;
  L->h_mid = p;  p += md;
  L->xn2 = p;    p += md;
  L->pre = p;    p += mf;
  L->mid = p;    p += mf;
  L->lse = reinterpret_cast<float*>(base + align256(static_cast<size_t>(v.M) * (8 * static_cast<size_t>(v.d) + 2 * v.ff) * 2));
}

size_t vit_head_saved_bytes(const VitDims& v, int n_img) {
  // col [n*G*G, kpp] + emb [M, d]
  return align256(static_cast<size_t>(n_img) * v.G * v.G * v.kpp * 2) + align256(static_cast<size_t>(v.M) * v.d * 2);
}

}  // namespace

extern "C" {

size_t merlin_vit_workspace_bytes(const merlin_vit_weights* w, int n_img) {
  const VitDims v = vit_dims(w, n_img);
  // col + patch_out + one layer's workspace
  return align256(static_cast<size_t>(n_img) * v.G * v.G * v.kpp * 2) +
         align256(static_cast<size_t>(n_img) * v.G * v.G * v.d * 2) + merlin_vit_layer_workspace_bytes(v.M, v.d, v.ff);
}

int merlin_vit_forward(void* stream, const merlin_vit_weights* w, const void* pixels, int n_img, void* workspace,
                       void* hidden, int cta_group) {
  MB_CHECK_ARG(w && pixels && workspace && hidden && n_img > 0, "vit_forward: null argument");
  MB_CHECK_ARG(w->layers && w->num_layers >= 0 && w->hidden == w->heads * 64, "vit_forward: bad weights (head_dim 64)");
  const VitDims v = vit_dims(w, n_img);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* col = ws;
  void* patch_out = ws + align256(static_cast<size_t>(n_img) * v.G * v.G * v.kpp * 2);
  void* layer_ws = static_cast<uint8_t*>(patch_out) + align256(static_cast<size_t>(n_img) * v.G * v.G * v.d * 2);
  int rc;
  // HF:clip/modeling_clip.py:202-218 (conv k=P, s=P, no bias as im2col + GEMM), :676-677 (pre_layrnorm)
  if ((rc = merlin_patch_im2col(stream, pixels, col, n_img, w->image_size, w->patch_size, v.kpp)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_bf16(stream, col, v.kpp, w->patch_weight, v.kpp, patch_out, v.d, nullptr, 0, nullptr,
                             n_img * v.G * v.G, v.d, v.kpp, MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_vit_embed_ln(stream, patch_out, w->class_embedding, w->position_embedding, w->pre_ln_w, w->pre_ln_b,
                                hidden, n_img, v.T, v.d, w->ln_eps)) != MERLIN_OK)
    return rc;
  for (int j = 0; j < w->num_layers; ++j)
    if ((rc = merlin_vit_layer_forward(stream, &w->layers[j], hidden, layer_ws, n_img, v.T, v.d, w->heads, v.ff,
                                       w->ln_eps, cta_group)) != MERLIN_OK)
      return rc;
  return MERLIN_OK;
}

size_t merlin_vit_saved_bytes(const merlin_vit_weights* w, int n_img) {
  const VitDims v = vit_dims(w, n_img);
  return vit_head_saved_bytes(v, n_img) + static_cast<size_t>(w->num_layers) * vit_layer_saved_bytes(v, n_img, w->heads) +
         align256(static_cast<size_t>(n_img) * v.G * v.G * v.d * 2);   // + patch_out scratch
}

int merlin_vit_forward_train(void* stream, const merlin_vit_weights* w, const void* pixels, int n_img, void* saved,
                             void* hidden, int cta_group) {
  MB_CHECK_ARG(w && pixels && saved && hidden && n_img > 0, "vit_forward_train: null argument");
  MB_CHECK_ARG(w->layers && w->hidden == w->heads * 64, "vit_forward_train: bad weights (head_dim 64)");
  const VitDims v = vit_dims(w, n_img);
  uint8_t* sv = static_cast<uint8_t*>(saved);
  void* col = sv;
  void* emb = sv + align256(static_cast<size_t>(n_img) * v.G * v.G * v.kpp * 2);
  uint8_t* layers = sv + vit_head_saved_bytes(v, n_img);
  const size_t per_layer = vit_layer_saved_bytes(v, n_img, w->heads);
  void* patch_out = layers + static_cast<size_t>(w->num_layers) * per_layer;
  const int M = v.M, d = v.d, ff = v.ff;
  int rc;
  if ((rc = merlin_patch_im2col(stream, pixels, col, n_img, w->image_size, w->patch_size, v.kpp)) != MERLIN_OK) return rc;
  if ((rc = merlin_gemm_bf16(stream, col, v.kpp, w->patch_weight, v.kpp, patch_out, d, nullptr, 0, nullptr,
                             n_img * v.G * v.G, d, v.kpp, MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_vit_embed(stream, patch_out, w->class_embedding, w->position_embedding, emb, n_img, v.T, d)) != MERLIN_OK)
    return rc;
  VitSaved::Layer L, Lnext;
  if (w->num_layers > 0) vit_layer_layout(layers, v, &L);
  void* h0 = w->num_layers > 0 ? static_cast<void*>(L.h) : hidden;
  if ((rc = merlin_layernorm(stream, emb, d, w->pre_ln_w, w->pre_ln_b, h0, d, M, d, w->ln_eps)) != MERLIN_OK) return rc;
  for (int j = 0; j < w->num_layers; ++j) {
    const merlin_vit_layer_weights* lw = &w->layers[j];
    vit_layer_layout(layers + j * per_layer, v, &L);
    void* h_out = hidden;
    if (j + 1 < w->num_layers) {
      vit_layer_layout(layers + (j + 1) * per_layer, v, &Lnext);
      h_out = Lnext.h;                      // layer j writes its output straight into layer j+1's input slot
    }
    if ((rc = merlin_layernorm(stream, L.h, d, lw->ln1_w, lw->ln1_b, L.xn1, d, M, d, w->ln_eps)) != MERLIN_OK) return rc;
    if ((rc = merlin_gemm_bf16(stream, L.xn1, d, lw->qkv_w, d, L.qkv, 3 * d, nullptr, 0, lw->qkv_b, M, 3 * d, d,
                               MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_flash_attn_fwd(stream, L.qkv, 3 * d, L.att, d, L.lse, nullptr, n_img, v.T, w->heads, 64, 0, 0.f)) !=
        MERLIN_OK)
      return rc;
    if ((rc = merlin_gemm_bf16(stream, L.att, d, lw->out_w, d, L.h_mid, d, L.h, d, lw->out_b, M, d, d,
                               MERLIN_EPI_RESIDUAL, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_layernorm(stream, L.h_mid, d, lw->ln2_w, lw->ln2_b, L.xn2, d, M, d, w->ln_eps)) != MERLIN_OK) return rc;
    if ((rc = merlin_gemm_bf16(stream, L.xn2, d, lw->fc1_w, d, L.pre, ff, nullptr, 0, lw->fc1_b, M, ff, d,
                               MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_quick_gelu(stream, L.pre, nullptr, L.mid, static_cast<int64_t>(M) * ff)) != MERLIN_OK) return rc;
    if ((rc = merlin_gemm_bf16(stream, L.mid, ff, lw->fc2_w, ff, h_out, d, L.h_mid, d, lw->fc2_b, M, d, ff,
                               MERLIN_EPI_RESIDUAL, cta_group)) != MERLIN_OK)
      return rc;
  }
  return MERLIN_OK;
}

size_t merlin_vit_backward_workspace_bytes(const merlin_vit_weights* w, int n_img) {
  const VitDims v = vit_dims(w, n_img);
  // dmid [M,ff] + dxn, dh_mid, datt [M,d] + dqkv [M,3d] + demb [M,d] + dpatch [n*G*G,d] + attention-backward scratch
  return align256(static_cast<size_t>(v.M) * (static_cast<size_t>(v.ff) + 8 * v.d) * 2) +
         align256(merlin_flash_attn_bwd_workspace_bytes(n_img, v.T, w->heads, 64)) + 1024;
}

int merlin_vit_backward(void* stream, const merlin_vit_weights* w, const merlin_vit_grads* g, void* saved, void* dh,
                        void* workspace, int n_img, int accumulate, int cta_group) {
  MB_CHECK_ARG(w && g && saved && dh && workspace && n_img > 0, "vit_backward: null argument");
  MB_CHECK_ARG(w->layers && g->layers, "vit_backward: null layer tables");
  const VitDims v = vit_dims(w, n_img);
  const int M = v.M, d = v.d, ff = v.ff;
  uint8_t* sv = static_cast<uint8_t*>(saved);
  void* col = sv;
  void* emb = sv + align256(static_cast<size_t>(n_img) * v.G * v.G * v.kpp * 2);
  uint8_t* layers = sv + vit_head_saved_bytes(v, n_img);
  const size_t per_layer = vit_layer_saved_bytes(v, n_img, w->heads);
  __nv_bfloat16* dmid = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* dxn = dmid + static_cast<size_t>(M) * ff;
  __nv_bfloat16* dh_mid = dxn + static_cast<size_t>(M) * d;
  __nv_bfloat16* datt = dh_mid + static_cast<size_t>(M) * d;
  __nv_bfloat16* dqkv = datt + static_cast<size_t>(M) * d;
  __nv_bfloat16* demb = dqkv + static_cast<size_t>(M) * 3 * d;
  __nv_bfloat16* dpatch = demb + static_cast<size_t>(M) * d;
  void* attn_ws = static_cast<uint8_t*>(workspace) + align256(static_cast<size_t>(M) * (static_cast<size_t>(ff) + 8 * d) * 2);
  const int epi = accumulate ? MERLIN_EPI_RESIDUAL : MERLIN_EPI_STORE;
  int rc;
  VitSaved::Layer L;
  for (int j = w->num_layers - 1; j >= 0; --j) {
    const merlin_vit_layer_weights* lw = &w->layers[j];
    const merlin_vit_layer_grads* lg = &g->layers[j];
    vit_layer_layout(layers + j * per_layer, v, &L);
    // ---- MLP: h_out = h_mid + fc2(quick_gelu(fc1(LN2(h_mid))))
    if ((rc = merlin_gemm_bf16_ex(stream, dh, d, 0, lw->fc2_w, ff, 1, dmid, ff, nullptr, 0, nullptr, M, ff, d,
                                  MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_gemm_bf16_ex(stream, dh, d, 1, L.mid, ff, 1, lg->fc2_w, ff, lg->fc2_w, ff, nullptr, d, ff, M, epi,
                                  cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_colsum(stream, dh, d, lg->fc2_b_f32, M, d)) != MERLIN_OK) return rc;
    if ((rc = merlin_quick_gelu(stream, L.pre, dmid, dmid, static_cast<int64_t>(M) * ff)) != MERLIN_OK) return rc;  // dpre
    if ((rc = merlin_gemm_bf16_ex(stream, dmid, ff, 1, L.xn2, d, 1, lg->fc1_w, d, lg->fc1_w, d, nullptr, ff, d, M, epi,
                                  cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_colsum(stream, dmid, ff, lg->fc1_b_f32, M, ff)) != MERLIN_OK) return rc;
    if ((rc = merlin_gemm_bf16_ex(stream, dmid, ff, 0, lw->fc1_w, d, 1, dxn, d, nullptr, 0, nullptr, M, d, ff,
                                  MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_layernorm_bwd(stream, L.h_mid, lw->ln2_w, dxn, dh, dh_mid, lg->ln2_w_f32, lg->ln2_b_f32, M, d,
                                   w->ln_eps)) != MERLIN_OK)
      return rc;
    // ---- attention: h_mid = h + out_proj(attn(qkv(LN1(h))))
    if ((rc = merlin_gemm_bf16_ex(stream, dh_mid, d, 0, lw->out_w, d, 1, datt, d, nullptr, 0, nullptr, M, d, d,
                                  MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_gemm_bf16_ex(stream, dh_mid, d, 1, L.att, d, 1, lg->out_w, d, lg->out_w, d, nullptr, d, d, M, epi,
                                  cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_colsum(stream, dh_mid, d, lg->out_b_f32, M, d)) != MERLIN_OK) return rc;
    if ((rc = merlin_flash_attn_bwd(stream, L.qkv, 3 * d, L.att, datt, d, L.lse, nullptr, dqkv, 3 * d, attn_ws, n_img, v.T,
                                    w->heads, 64, 0, 0.f)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_gemm_bf16_ex(stream, dqkv, 3 * d, 1, L.xn1, d, 1, lg->qkv_w, d, lg->qkv_w, d, nullptr, 3 * d, d, M,
                                  epi, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_colsum(stream, dqkv, 3 * d, lg->qkv_b_f32, M, 3 * d)) != MERLIN_OK) return rc;
    if ((rc = merlin_gemm_bf16_ex(stream, dqkv, 3 * d, 0, lw->qkv_w, d, 1, dxn, d, nullptr, 0, nullptr, M, d, 3 * d,
                                  MERLIN_EPI_STORE, cta_group)) != MERLIN_OK)
      return rc;
    if ((rc = merlin_layernorm_bwd(stream, L.h, lw->ln1_w, dxn, dh_mid, dh, lg->ln1_w_f32, lg->ln1_b_f32, M, d,
                                   w->ln_eps)) != MERLIN_OK)
      return rc;
  }
  // ---- pre_layrnorm, CLS / position embeddings, patch convolution
  if ((rc = merlin_layernorm_bwd(stream, emb, w->pre_ln_w, dh, nullptr, demb, g->pre_ln_w_f32, g->pre_ln_b_f32, M, d,
                                 w->ln_eps)) != MERLIN_OK)
    return rc;
  if ((rc = merlin_vit_embed_bwd(stream, demb, dpatch, g->position_f32, g->class_f32, n_img, v.T, d)) != MERLIN_OK)
    return rc;
  return merlin_gemm_bf16_ex(stream, dpatch, d, 1, col, v.kpp, 1, g->patch_weight, v.kpp, g->patch_weight, v.kpp, nullptr,
                             d, v.kpp, n_img * v.G * v.G, epi, cta_group);
}

}  // extern "C"
