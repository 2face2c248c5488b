/* merlin_b200.h — C ABI of libmerlin_b200.so: the B200 (sm_100a) kernels behind the mmgpt hot path.
 *
 * Conventions: every entry point returns int status (MERLIN_OK == 0, negative on
 * error, text via merlin_last_error()), takes the CUDA stream as an opaque pointer followed by raw
 * device pointers and explicit sizes, never allocates device memory, never synchronises the stream
 * and holds no mutable global state besides one-time kernel attributes (re-entrant per stream).
 * All floating tensors are bf16 unless a parameter is spelled `float*`; index tensors are int64/int32
 * as spelled.  "ld*" are leading dimensions in ELEMENTS.
 *
 * Each declaration cites the reference code it replaces (paths relative to the Ahnsun/merlin tree;
 * "HF:" = transformers/models/..., the third-party library the reference calls into).
 */
#ifndef MERLIN_B200_H_
#define MERLIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MERLIN_OK 0
#define MERLIN_E_INVALID (-1) /* bad argument (shape, alignment, null pointer) */
#define MERLIN_E_CUDA (-2)    /* CUDA runtime / driver error */
#define MERLIN_E_UNSUPPORTED (-3)

/* Epilogues of merlin_gemm_bf16 */
#define MERLIN_EPI_STORE 0     /* C = A·Bᵀ (+ bias) */
#define MERLIN_EPI_RESIDUAL 1  /* C = A·Bᵀ (+ bias) + R ; R may alias C */
#define MERLIN_EPI_QUICKGELU 2 /* C = quick_gelu(A·Bᵀ + bias) */
#define MERLIN_EPI_SWIGLU 3    /* B = gate/up rows interleaved in blocks of 128; C[M,N/2] = silu(g)·u */

/* Text of the last error raised on the calling thread. */
const char* merlin_last_error(void);

/* C[M,Nout] = epilogue(A[M,K] · B[N,K]ᵀ), fp32 accumulation on tcgen05 tensor cores.
 * Replaces every nn.Linear on the path: HF:llama/modeling_llama.py q/k/v/o_proj, gate/up/down_proj
 * (LlamaMLP.forward :182-184), lm_head (mmgpt/model/mmgpt/llama_mmgpt.py:87), CLIP q/k/v/out_proj,
 * fc1/fc2 (HF:clip/modeling_clip.py:310-312,347-351), MLPProjector (mlp_projector.py:19-23), and the
 * im2col'd convolutions (patch embedding HF:clip/modeling_clip.py:202-218, ConvProjector
 * conv_projector.py:23-39).  cta_group: 1 = 128x256 tiles per CTA, 2 = 256x256 tiles per CTA pair. */
int merlin_gemm_bf16(void* stream, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                     int64_t ldc, const void* R, int64_t ldr, const void* bias, int M, int N, int K,
                     int epilogue, int cta_group);

/* qkv[M, N] = A[M,K] · Wqkv[N,K]ᵀ with half-split rotary embedding (head_dim 128, fp32 cos/sin
 * tables [max_pos,64]) applied in the epilogue to columns [0, rope_cols) — the q and k blocks.
 * Row r uses position positions[r] (or r % seq_len when positions == NULL).
 * Replaces mmgpt/utils/llama_flash_attn_monkey_patch.py:35-59 (q/k/v_proj + apply_rotary_pos_emb). */
int merlin_gemm_qkv_rope(void* stream, const void* A, int64_t lda, const void* Wqkv, int64_t ldb,
                         void* qkv, int64_t ldc, const float* rope_cos, const float* rope_sin,
                         const int* positions, int M, int N, int K, int rope_cols, int seq_len,
                         int cta_group);

/* out[B,S,H*hd] = softmax(scale·QKᵀ + mask)·V for a packed qkv buffer [B,S,3*H*hd] ([q|k|v] blocks),
 * head_dim 64 or 128.  causal != 0 adds the causal mask; seqlens[b] (NULL = S) is the number of valid
 * right-padded tokens of sample b: keys beyond it are masked, query rows beyond it are written as
 * zeros (flash-attn's pad_input).  lse (NULL or [B,H,S] fp32) receives log-sum-exp for backward.
 * softmax_scale <= 0 selects 1/sqrt(head_dim).
 * Replaces flash_attn_varlen_qkvpacked_func + unpad_input/pad_input
 * (mmgpt/utils/llama_flash_attn_monkey_patch.py:63-102) and CLIP eager attention
 * (HF:clip/modeling_clip.py:261-279). */
int merlin_flash_attn_fwd(void* stream, const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out,
                          float* lse, const int* seqlens, int B, int S, int H, int head_dim,
                          int causal, float softmax_scale);

/* y[M,D] = weight · (x · rsqrt(mean(x²) + eps)); fp32 math, bf16 I/O (D % 8 == 0, D <= 8192).
 * Replaces LlamaRMSNorm.forward (HF:llama/modeling_llama.py:62-67). */
int merlin_rmsnorm(void* stream, const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy,
                   int M, int D, float eps);

/* y[M,D] = LayerNorm(x) · weight + bias.  Replaces nn.LayerNorm in CLIPEncoderLayer
 * (HF:clip/modeling_clip.py:363-384: layer_norm1 / layer_norm2). */
int merlin_layernorm(void* stream, const void* x, int64_t ldx, const void* weight, const void* bias,
                     void* y, int64_t ldy, int M, int D, float eps);

/* col[n_img*G*G, ld_col] = im2col of pixels [n_img,3,image,image] for the stride-P patch convolution,
 * k = (c*P + ky)*P + kx, zero padded up to ld_col (ld_col % 8 == 0) so that the patch embedding
 * (HF:clip/modeling_clip.py:202-218, Conv2d(3,d,P,P,bias=False)) becomes one merlin_gemm_bf16. */
int merlin_patch_im2col(void* stream, const void* pixels, void* col, int n_img, int image_size,
                        int patch_size, int ld_col);

/* hidden[n_img*tokens, D] = pre_layrnorm(cat(class_embedding, patch_out[img]) + position_embedding).
 * Replaces the tail of CLIPVisionEmbeddings.forward and CLIPVisionTransformer.pre_layrnorm
 * (HF:clip/modeling_clip.py:210-218, 676-677). */
int merlin_vit_embed_ln(void* stream, const void* patch_out, const void* class_embedding,
                        const void* position_embedding, const void* ln_weight, const void* ln_bias,
                        void* hidden, int n_img, int tokens, int D, float eps);

/* col[n_img*Go*Go, 9*C] = im2col (3x3, pad 1, stride) of channel-last features; feature (img,y,x)
 * is the row feat + (img*tok_stride + tok_off + y*grid + x)*C, so the CLS-dropping slice
 * hidden_states[-2][:, 1:] (clip_encoder.py:64-72) is read in place (tok_stride=577, tok_off=1).
 * k = (ky*3+kx)*C + c.  Turns ConvProjector.forward (conv_projector.py:23-39) into one GEMM. */
int merlin_conv3x3_im2col(void* stream, const void* feat, void* col, int n_img, int grid, int C,
                          int stride, int tok_stride, int tok_off);

/* inputs_embeds[b,t] = image_features[feat_offset[b] + r] if input_ids[b,t] == im_patch_token, where
 * r = number of earlier <im_patch> tokens of sample b (and r < feat_limit[b]); otherwise
 * embed_table[input_ids[b,t]].  Bit-exact row copies.  feat_offset / feat_limit: int32 [B] device
 * arrays (NULL = 0 / unlimited).  rank_ws: B*S + B ints of scratch; rank_ws[B*S + b] receives the
 * number of <im_patch> tokens of sample b (the host checks it against the image count).
 * Replaces embed_tokens + the per-sample / per-image torch.cat splice of
 * prepare_inputs_labels_for_multimodal (mmgpt/model/mmgpt/base_mmgpt.py:99-160). */
int merlin_embed_splice(void* stream, const int64_t* input_ids, const void* embed_table,
                        const void* image_features, void* inputs_embeds, int* rank_ws, int B, int S, int D,
                        int n_feature_rows, const int* feat_offset, const int* feat_limit, int64_t vocab,
                        int64_t im_patch_token);

/* In-place half-split RoPE on the q and k blocks (H heads of 128 each) of packed qkv rows.
 * Stand-alone form of the merlin_gemm_qkv_rope epilogue (HF:llama/modeling_llama.py:138-168). */
int merlin_rope_qk_inplace(void* stream, void* qkv, int64_t ld, const float* rope_cos,
                           const float* rope_sin, const int* positions, int M, int H, int seq_len);

/* Shifted causal-LM cross-entropy: row (b,t) scores labels[b,t+1]; rows whose label is ignore_index
 * (or t == S-1) are skipped.  loss_sum_count[0] = Σ nll, [1] = #scored rows (mean = sum/count);
 * row_nll (NULL or [B*S]) receives per-row values.  fp32 math on bf16 logits.
 * Replaces mmgpt/model/mmgpt/llama_mmgpt.py:89-100. */
int merlin_cross_entropy_fwd(void* stream, const void* logits, int64_t ld, const int64_t* labels,
                             int B, int S, int V, int64_t ignore_index, float* row_nll,
                             float* loss_sum_count);

/* out[r] = argmax_v logits[r, v] (lowest index on ties, as torch.argmax): greedy decoding. */
int merlin_argmax(void* stream, const void* logits, int64_t ld, int rows, int V, int64_t* out);

/* ---- decode (one new token against the KV cache; HBM-bound GEMV + split-K flash-decode kernels) ----
 * Replaces the HF eager decode step the reference reaches through generate()
 * (llama_mmgpt.py:114-134 -> LlamaModel with past_key_values).  Batch 1, head_dim 128. */
typedef struct merlin_decoder_weights {
  int num_layers, hidden, heads, intermediate, vocab;
  float rms_eps;
  const void* const* qkv_w;     /* [L] -> [3*hidden, hidden]   q|k|v rows                          */
  const void* const* o_w;       /* [L] -> [hidden, hidden]                                        */
  const void* const* gate_up_w; /* [L] -> [2*intermediate, hidden], gate/up interleaved per 128   */
  const void* const* down_w;    /* [L] -> [hidden, intermediate]                                  */
  const void* const* ln1_w;     /* [L] -> [hidden] input_layernorm                                */
  const void* const* ln2_w;     /* [L] -> [hidden] post_attention_layernorm                       */
  const void* norm_w;           /* [hidden] final norm                                            */
  const void* lm_head_w;        /* [vocab, hidden]                                                */
  void* const* kv_cache;        /* [L] -> [capacity, 2*hidden] bf16, row t = [k_t (rotated) | v_t] */
  const float* rope_cos;        /* [max_pos, 64] fp32                                             */
  const float* rope_sin;
} merlin_decoder_weights;

/* Bytes of scratch merlin_decode_step needs. */
size_t merlin_decode_workspace_bytes(int hidden, int intermediate, int heads, int n_splits);

/* One decoder pass for the token whose embedding is x[hidden] (bf16, overwritten with the final
 * residual stream) at position *pos_dev: appends its k/v to every layer's cache row *pos_dev, attends
 * over rows 0..*pos_dev, and writes logits[vocab] (bf16).  The position lives in device memory so
 * the call can be captured in a CUDA graph and replayed. */
int merlin_decode_step(void* stream, const merlin_decoder_weights* w, void* x, const int* pos_dev,
                       void* workspace, void* logits, int n_splits);

/* x[D] = embed_table[*token_dev] (token id read on the device). */
int merlin_embed_token(void* stream, const void* embed_table, const int64_t* token_dev, void* x, int D,
                       int64_t vocab);

/* out_ids[*n_out_dev] = *token_dev; ++*n_out_dev; ++*pos_dev  — closes one greedy step on the device. */
int merlin_record_advance(void* stream, const int64_t* token_dev, int64_t* out_ids, int* n_out_dev,
                          int* pos_dev, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* MERLIN_B200_H_ */
