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
    // rows of [k | v] (2D bf16) out of the packed qkv rows into cache[b, 0:S, :]
    for (int b = 0; b < B; ++b) {
      MB_CHECK_CUDA(cudaMemcpy2DAsync(static_cast<__nv_bfloat16*>(kv_cache) + static_cast<size_t>(b) * kv_capacity * 2 * D,
                                      static_cast<size_t>(2) * D * 2, qkv + static_cast<size_t>(b) * S * 3 * D + D,
                                      static_cast<size_t>(3) * D * 2, static_cast<size_t>(2) * D * 2, S,
                                      cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    }
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
