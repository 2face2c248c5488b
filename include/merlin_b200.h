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

/* Same contraction with explicit operand layouts, for the backward GEMMs.  a_mn_major != 0: A is given as
 * the row-major [K, M] matrix Aᵀ (lda = its row pitch); b_mn_major != 0: B is given as the row-major
 * [K, N] matrix Bᵀ.  dgrad  dX[M,Kin] = dY[M,N]·W[N,Kin]   -> (A=dY, B=W with b_mn_major=1, "N"=Kin, "K"=N);
 * wgrad  dW[N,Kin] = dYᵀ·X -> (A=dY with a_mn_major=1, B=X with b_mn_major=1, "M"=N, "N"=Kin, "K"=tokens);
 * EPI_RESIDUAL with R == C accumulates into an existing gradient.  (autograd of the nn.Linear calls above) */
int merlin_gemm_bf16_ex(void* stream, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                        int b_mn_major, void* C, int64_t ldc, const void* R, int64_t ldr, const void* bias,
                        int M, int N, int K, int epilogue, int cta_group);

/* qkv[M, N] = A[M,K] · Wqkv[N,K]ᵀ with half-split rotary embedding (head_dim 128, fp32 cos/sin
 * tables [max_pos,64]) applied in the epilogue to columns [0, rope_cols) — the q and k blocks.
 * Row r uses position positions[r] (or r % seq_len when positions == NULL).
 * Replaces mmgpt/utils/llama_flash_attn_monkey_patch.py:35-59 (q/k/v_proj + apply_rotary_pos_emb). */
int merlin_gemm_qkv_rope(void* stream, const void* A, int64_t lda, const void* Wqkv, int64_t ldb,
                         void* qkv, int64_t ldc, const float* rope_cos, const float* rope_sin,
                         const int* positions, int M, int N, int K, int rope_cols, int seq_len,
                         int cta_group);
/* Same GEMM with the KV-cache append fused into its epilogue (kv_cache == NULL: identical to the above): every
 * staged [128 tokens x 64 columns] chunk of the rotated k block and of the v block is also written, by a second
 * TMA store, into kv_cache[B, kv_capacity/32, H, 32, 256] at tokens 0 .. seq_len-1 (the layout of
 * merlin_decoder_weights.kv_cache / merlin_kv_append).  M == B * seq_len; B > 1 needs seq_len % 128 == 0.
 * Replaces the HF `past_key_value` tuple update of the prefill (HF:llama/modeling_llama.py, `use_cache=True`). */
int merlin_gemm_qkv_rope_kv(void* stream, const void* A, int64_t lda, const void* Wqkv, int64_t ldb, void* qkv,
                            int64_t ldc, const float* rope_cos, const float* rope_sin, const int* positions, int M,
                            int N, int K, int rope_cols, int seq_len, void* kv_cache, int64_t kv_capacity, int B,
                            int H, int cta_group);

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
/* Same with an explicit direction: inverse != 0 rotates by −θ (the backward of the rotation). */
int merlin_rope_qk_inplace_ex(void* stream, void* qkv, int64_t ld, const float* rope_cos,
                              const float* rope_sin, const int* positions, int M, int H, int seq_len,
                              int inverse);

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
  void* const* kv_cache;        /* [L] -> [kv_capacity/32, heads, 32, 256] bf16: row (h, t) = [k_t^h rotated (128) |
                                 * v_t^h (128)] at ((t/32 * heads + h) * 32 + t%32) * 256 — blocks of 32 tokens, head-major
                                 * inside a block: a (head, 32-key) tile is 16 KB contiguous, a token range is dense  */
  int64_t kv_capacity;          /* tokens the caches can hold (a multiple of 32)                                      */
  const float* rope_cos;        /* [max_pos, 64] fp32                                             */
  const float* rope_sin;
} merlin_decoder_weights;

/* Bytes of scratch merlin_decode_step needs.  The caller zero-fills the workspace ONCE before its first use (it
 * holds self-resetting arrival counters); no re-zeroing between calls. */
size_t merlin_decode_workspace_bytes(int hidden, int intermediate, int heads, int n_splits);

/* One decoder pass for the token whose embedding is x[hidden] (bf16, overwritten with the final
 * residual stream) at position *pos_dev: appends its k/v to every layer's cache row *pos_dev, attends
 * over rows 0..*pos_dev, and writes logits[vocab] (bf16).  The position lives in device memory so
 * the call can be captured in a CUDA graph and replayed. */
int merlin_decode_step(void* stream, const merlin_decoder_weights* w, void* x, const int* pos_dev,
                       void* workspace, void* logits, int n_splits);

/* The same decode step as ONE persistent kernel (one CTA per SM, cooperative launch): a producer thread per
 * CTA streams that CTA's share of every layer's weight rows and KV-cache tiles through a shared-memory ring
 * with TMA bulk copies and runs ahead across phase and layer boundaries while the consumer warps synchronise
 * grid-wide on the activations.  `workspace` (merlin_decode_stream_workspace_bytes, 256-byte aligned) must
 * first be bound to the weights and caches with merlin_decode_stream_prepare — it copies the per-layer pointer
 * tables to the device, clears the counters and synchronises `stream` (not capturable; call it again whenever
 * a pointer in `w` changes or after a failed launch).  merlin_decode_stream_step is capturable: x_in[hidden]
 * is the token embedding (read-only), logits[vocab] receives the lm_head output, the new k/v go to row
 * *pos_dev of every cache.  Same reference code as merlin_decode_step. */
size_t merlin_decode_stream_workspace_bytes(int num_layers, int hidden, int heads, int intermediate);
int merlin_decode_stream_prepare(void* stream, const merlin_decoder_weights* w, void* workspace);
int merlin_decode_stream_step(void* stream, const merlin_decoder_weights* w, const void* x_in, const int* pos_dev,
                              void* workspace, void* logits);
/* x[D] = embed_table[*token_dev] (token id read on the device). */
int merlin_embed_token(void* stream, const void* embed_table, const int64_t* token_dev, void* x, int D,
                       int64_t vocab);

/* out_ids[*n_out_dev] = *token_dev; ++*n_out_dev; ++*pos_dev  — closes one greedy step on the device. */
int merlin_record_advance(void* stream, const int64_t* token_dev, int64_t* out_ids, int* n_out_dev,
                          int* pos_dev, int capacity);

/* Beam search: ONE decode position of num_beams sequences that share a prompt (generate(num_beams=k),
 * mmgpt/engine/eval/eval_box.py:246-254; HF GenerationMixin._beam_search runs the beams as a batch of k sequences, each
 * with its own copy of the whole cache, and re-orders every cache tensor every step).  The cache is split: prefix_kv[l]
 * is the prompt's cache (merlin_decoder_weights.kv_cache layout, shared and read once for all beams), suffix_kv[l] is
 * [num_beams, suffix_capacity/32, heads, 32, 256] (same row layout, per beam: the tokens generated so far — the only
 * thing a beam re-ordering has to move).  pos_dev = device int[2] {prefix_len, t}: the new token of every beam is
 * suffix row t at position prefix_len + t (device memory, so the call can be captured in a CUDA graph). */
typedef struct merlin_beam_state {
  int num_beams;                /* 1..8 */
  int64_t suffix_capacity;      /* tokens per beam (a multiple of 32) */
  const void* const* prefix_kv; /* [L] */
  void* const* suffix_kv;       /* [L] */
  const int* pos_dev;
} merlin_beam_state;
size_t merlin_decode_batched_workspace_bytes(int hidden, int intermediate, int heads, int n_splits, int num_beams);
/* x[num_beams, hidden] (bf16 token embeddings, overwritten with the final residual stream) -> logits[num_beams,
 * ld_logits] (bf16; ld_logits >= vocab).  Every weight matrix is streamed from HBM once for all beams (mma.sync
 * m16n8k16 with the beams as the 8 columns).  w->kv_cache / kv_capacity are not used. */
int merlin_decode_step_batched(void* stream, const merlin_decoder_weights* w, const merlin_beam_state* bs, void* x,
                               void* workspace, void* logits, int64_t ld_logits, int n_splits);

/* Append S tokens to a KV cache (layout above): for b < B, s < S, h < H copies k and v of head h out of the packed
 * qkv row [b*S + s] (ldq elements per row; k at column D + h*128, v at 2D + h*128) into row (h, pos0 + s) of
 * cache block b.  cache holds B blocks of kv_capacity*H*256 elements.  (HF KV-cache update.) */
int merlin_kv_append(void* stream, const void* qkv, int64_t ldq, void* kv_cache, int64_t kv_capacity, int B, int S,
                     int D, int H, int pos0);

/* ---- data-loader image path ----
 * out[3, S, S] = normalize(rescale(resize(src))) for one uint8 RGB image src[H, W, 3] (row pitch in bytes): PIL's
 * default BICUBIC `image.resize((S, S))` reproduced bit-exactly on the uint8 pixels (third-party Pillow,
 * libImaging/Resample.c: 22-bit fixed-point weights, horizontal pass to uint8, then vertical pass), then
 * x / 255 and (x - mean) / std as the CLIP image processor does — the 'resize' mode of
 * mmgpt/data/dataset/base_dataset.py:190-192 and mmgpt/utils/mm_utils.py:41-43.  out is fp32 (out_bf16 = 0) or
 * bf16; out_resized_u8 (NULL or [S, S, 3]) also receives the resized uint8 pixels.  workspace:
 * merlin_image_preprocess_workspace_bytes, 256-byte aligned.  The filter tables are built on the host, so this
 * call synchronises `stream` once (data-loading path, not capturable). */
size_t merlin_image_preprocess_workspace_bytes(int H, int W, int S);
int merlin_image_preprocess(void* stream, const void* src_rgb_u8, int64_t src_pitch_bytes, int H, int W, int S,
                            const float* mean3, const float* std3, void* workspace, void* out, int out_bf16,
                            void* out_resized_u8);

/* ---- composite forward entry points: one call per layer, kernels sequenced on the caller's stream ---- */
typedef struct merlin_layer_weights {       /* one Vicuna decoder layer, packed layouts as above */
  const void *qkv_w, *o_w, *gate_up_w, *down_w, *ln1_w, *ln2_w;
} merlin_layer_weights;
typedef struct merlin_vit_layer_weights {   /* one CLIP encoder layer (q|k|v fused, with biases) */
  const void *qkv_w, *qkv_b, *out_w, *out_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} merlin_vit_layer_weights;

/* x[B*S, D] <- LlamaDecoderLayer(x): x += Attn(RMSNorm(x)); x += MLP(RMSNorm(x)), in place
 * (HF:llama/modeling_llama.py:303-333 with the flash-attn patch).  kv_cache (NULL = off): B blocks of
 * [kv_capacity/32, H, 32, 256] (see merlin_decoder_weights) receive the rotated keys and the values of the S prompt
 * tokens for later decode steps. */
size_t merlin_decoder_layer_workspace_bytes(int M, int D, int FF);
int merlin_decoder_layer_forward(void* stream, const merlin_layer_weights* w, void* x, void* workspace,
                                 const float* rope_cos, const float* rope_sin, const int* seqlens,
                                 void* kv_cache, int64_t kv_capacity, int B, int S, int D, int H, int FF, float eps,
                                 int cta_group);

/* h[n_img*tokens, d] <- CLIPEncoderLayer(h) in place (HF:clip/modeling_clip.py:363-384). */
size_t merlin_vit_layer_workspace_bytes(int M, int d, int ff);
int merlin_vit_layer_forward(void* stream, const merlin_vit_layer_weights* w, void* h, void* workspace, int n_img,
                             int tokens, int d, int heads, int ff, float eps, int cta_group);

/* ConvProjector.forward (mmgpt/model/projector/conv_projector.py:23-39) in one call: the 3x3 / pad 1 / stride-s
 * convolution of the channel-last patch features as merlin_conv3x3_im2col (into `workspace`,
 * merlin_projector_workspace_bytes) + merlin_gemm_bf16 with the bias epilogue.  feat addressing as in
 * merlin_conv3x3_im2col (tok_stride = 577, tok_off = 1 reads hidden_states[-2][:, 1:] in place); weight is
 * [D_out, 3, 3, C] flattened to [D_out, 9C]; out[n_img * Go * Go, D_out] with row pitch ldo. */
size_t merlin_projector_workspace_bytes(int n_img, int grid, int C, int stride);
int merlin_projector_forward(void* stream, const void* feat, const void* weight, const void* bias, void* workspace,
                             void* out, int64_t ldo, int n_img, int grid, int C, int D_out, int stride, int tok_stride,
                             int tok_off, int cta_group);

/* lm_head + shifted cross-entropy in one call (mmgpt/model/mmgpt/llama_mmgpt.py:86-100): logits[M, V] (row pitch
 * ldl elements, multiple of 8) = hidden[M, D] . lm_head[V, D]^T, and, when labels != NULL, the shifted fp32 CE of
 * merlin_cross_entropy_fwd (loss_sum_count[2], optional row_nll[M]).  M = B*S. */
int merlin_lmhead_ce_forward(void* stream, const void* hidden, int64_t ldh, const void* lm_head_w, void* logits,
                             int64_t ldl, const int64_t* labels, int B, int S, int D, int V, int64_t ignore_index,
                             float* row_nll, float* loss_sum_count, int cta_group);
/* Its backward (autograd of the two lines above): `logits` is overwritten with dlogits (merlin_cross_entropy_bwd
 * in place, scaled by grad_scale / count), d_lm_head_w[V, D] (NULL to skip) receives dlogits^T . hidden (added to
 * its contents when accumulate != 0), dhidden[M, D] receives dlogits . lm_head. */
int merlin_lmhead_ce_backward(void* stream, void* logits, int64_t ldl, const int64_t* labels, int B, int S, int D, int V,
                              int64_t ignore_index, const float* loss_sum_count, float grad_scale, const void* hidden,
                              int64_t ldh, const void* lm_head_w, void* d_lm_head_w, int accumulate, void* dhidden,
                              int64_t lddh, int cta_group);

/* Training forms of the decoder layer.  `saved` (merlin_decoder_layer_saved_bytes) keeps x_in, xn1, qkv, att,
 * x_mid, xn2, gate_up, mid and lse in that order (x_in first: a layer may write its output directly into the
 * next layer's saved block).  backward: dx [B*S, D] holds dL/dx_out on entry and dL/dx_in on return; weight
 * gradients go to the bf16 tensors of `g` (overwritten, or accumulated when accumulate != 0), the two norm
 * weights accumulate into fp32 [D] buffers.  Autograd of HF LlamaDecoderLayer + the flash-attn patch. */
typedef struct merlin_layer_saved {
  void *x_in, *xn1, *qkv, *att, *x_mid, *xn2, *gate_up, *mid;
  float* lse;
} merlin_layer_saved;
typedef struct merlin_layer_grads {
  void *qkv_w, *o_w, *gate_up_w, *down_w;
  float *ln1_w_f32, *ln2_w_f32;
} merlin_layer_grads;
size_t merlin_decoder_layer_saved_bytes(int M, int D, int FF, int H, int B, int S);
int merlin_decoder_layer_forward_train(void* stream, const merlin_layer_weights* w, const void* x, void* x_out,
                                       void* saved, const float* rope_cos, const float* rope_sin,
                                       const int* seqlens, int B, int S, int D, int H, int FF, float eps,
                                       int cta_group);
size_t merlin_decoder_layer_backward_workspace_bytes(int M, int D, int FF, int H, int B, int S);
int merlin_decoder_layer_backward(void* stream, const merlin_layer_weights* w, const merlin_layer_grads* g,
                                  void* saved, void* dx, void* workspace, const float* rope_cos,
                                  const float* rope_sin, const int* seqlens, int B, int S, int D, int H, int FF,
                                  float eps, int accumulate, int cta_group);

/* ---- backward / optimizer (training path; autograd of the ops above + the step loop of
 * mmgpt/engine/train/trainer.py:45-74 with mmgpt/utils/llrd_utils.py param groups) ---- */

/* RMSNorm backward: dx = rstd·(w⊙dy − x̂·mean(w⊙dy⊙x̂)) (+ dres, the gradient arriving through the residual
 * connection; NULL = none); dweight[D] (fp32, NULL = skip) accumulates Σ_rows dy⊙x̂. */
int merlin_rmsnorm_bwd(void* stream, const void* x, const void* weight, const void* dy, const void* dres,
                       void* dx, float* dweight, int M, int D, float eps);

/* Training form of LlamaMLP's activation: gate_up[M, 2FF] is the plain output of the packed gate/up GEMM
 * (kept for backward); mid[M, FF] = silu(gate)·up.  Backward: dgate_up (may alias gate_up) from dmid. */
int merlin_swiglu_fwd(void* stream, const void* gate_up, void* mid, int M, int FF);
int merlin_swiglu_bwd(void* stream, const void* gate_up, const void* dmid, void* dgate_up, int M, int FF);

/* dlogits = (softmax(logits) − onehot(labels shifted)) · grad_scale / count for scored rows, 0 elsewhere;
 * may run in place (dlogits == logits).  loss_sum_count from merlin_cross_entropy_fwd. */
int merlin_cross_entropy_bwd(void* stream, const void* logits, int64_t ld, const int64_t* labels, int B, int S,
                             int V, int64_t ignore_index, const float* loss_sum_count, float grad_scale,
                             void* dlogits, int64_t ldd);

/* Backward of merlin_embed_splice: rows of <im_patch> tokens (rank >= 0) are copied to
 * dimage_features[rank], all other rows are added into dembed_table[input_ids[t]] (either may be NULL). */
int merlin_embed_splice_bwd(void* stream, const int64_t* input_ids, const int* rank, const void* dx,
                            void* dembed_table, void* dimage_features, int T, int D, int n_feature_rows,
                            int64_t vocab);

/* out[N] (fp32) += column sums of x[M,N] (bias gradients). */
int merlin_colsum(void* stream, const void* x, int64_t ld, float* out, int M, int N);

/* delta[B,H,S] = rowsum(dO ∘ O): pre-pass of the flash-attention backward. */
int merlin_attn_delta(void* stream, const void* out, const void* dout, float* delta, int B, int S, int H,
                      int head_dim, int64_t ld);

/* Flash-attention backward (head_dim 128): dqkv[B*S, ld_dqkv] packed [dq|dk|dv] from the forward's packed
 * qkv, its output `out`, the incoming gradient `dout`, and the saved log-sum-exp.  dq/dk are gradients
 * w.r.t. the ROTATED q/k (apply merlin_rope_qk_inplace(..., inverse) afterwards).  workspace:
 * merlin_flash_attn_bwd_workspace_bytes() bytes (delta + fp32 dq accumulator).
 * Replaces the backward of flash_attn_varlen_qkvpacked_func (monkey_patch.py:82-95). */
size_t merlin_flash_attn_bwd_workspace_bytes(int B, int S, int H, int head_dim);
int merlin_flash_attn_bwd(void* stream, const void* qkv, int64_t ld_qkv, const void* out, const void* dout,
                          int64_t ld_out, const float* lse, const int* seqlens, void* dqkv, int64_t ld_dqkv,
                          void* workspace, int B, int S, int H, int head_dim, int causal, float softmax_scale);

/* out[0] = Σ g² over a flat bf16 gradient arena of n elements (n % 8 == 0): global grad norm.  `out` is
 * 4 floats (out[2..3] hold a 64-bit fixed-point accumulator: the sum is order-independent, hence
 * bit-identical on every data-parallel rank). */
int merlin_grad_sumsq(void* stream, const void* grads, int64_t n, float* out);

/* Fused AdamW over flat bf16 arenas (params, grads, exp_avg, exp_avg_sq; n elements).  Elements
 * [c·chunk, (c+1)·chunk) use lr·lr_mult[c] and weight_decay[c] (LLRD / no-decay groups; lr_mult 0 freezes).
 * Gradients are scaled by grad_scale and clipped to max_grad_norm (<= 0: off) using grad_sumsq[0].
 * torch.optim.AdamW update rule, fp32 math, bf16 state (the reference keeps bf16 weights and state). */
int merlin_adamw_step(void* stream, void* params, const void* grads, void* exp_avg, void* exp_avg_sq,
                      const float* lr_mult, const float* weight_decay, int chunk, int64_t n, float lr, float beta1,
                      float beta2, float eps, int step, float grad_scale, float max_grad_norm,
                      const float* grad_sumsq);

/* LayerNorm backward (CLIP layer_norm1/2, pre_layrnorm): dx (+ dres), dweight/dbias fp32 accumulators. */
int merlin_layernorm_bwd(void* stream, const void* x, const void* weight, const void* dy, const void* dres,
                         void* dx, float* dweight, float* dbias, int M, int D, float eps);

/* quick_gelu forward (dy == NULL: out = x·σ(1.702x)) or backward (out = dy·d/dx) over n elements
 * (HF QuickGELUActivation; training form of the fc1 epilogue). */
int merlin_quick_gelu(void* stream, const void* x, const void* dy, void* out, int64_t n);

/* Input gradient of ConvProjector's 3x3 convolution: gathers the im2col-layout gradient dcol[n*Go*Go, 9C]
 * back onto the feature rows dfeat + (img*tok_stride + tok_off + y*grid + x)*C (inverse of
 * merlin_conv3x3_im2col). */
int merlin_conv3x3_col2im(void* stream, const void* dcol, void* dfeat, int n_img, int grid, int C, int stride,
                          int tok_stride, int tok_off);

/* CLIP embedding assembly without the norm (training form): emb = cat(class_embedding, patch_out) + position,
 * and its backward: dpatch_out rows, fp32 dposition[tokens, D] and dclass[D] accumulators. */
int merlin_vit_embed(void* stream, const void* patch_out, const void* class_embedding,
                     const void* position_embedding, void* emb, int n_img, int tokens, int D);
int merlin_vit_embed_bwd(void* stream, const void* dhidden, void* dpatch_out, float* dposition, float* dclass,
                         int n_img, int tokens, int D);

/* y[n] (bf16) = x[n] (fp32): folds fp32 accumulators (norm-weight / bias gradients) into the arena. */
int merlin_f32_to_bf16(void* stream, const float* x, void* y, int64_t n);


/* ---- whole CLIP tower in one call each way (SURVEY.md §8b minimum set: merlin_vit_forward / merlin_vit_backward).
 * Replaces CLIPVisionTower.forward (mmgpt/model/vision_encoder/clip_encoder.py:74-82) over HF CLIPVisionModel
 * (HF:clip/modeling_clip.py:202-218 embeddings, :676-677 pre_layrnorm, :363-384 encoder layers) and its autograd.
 * `num_layers` is the number of encoder layers to RUN (select_layer already applied: 23 of 24 for -2);
 * hidden_states[select_layer] comes back as hidden[n_img*tokens, hidden] (CLS row first).  pixels: bf16
 * [n_img, 3, image_size, image_size].  Biases / norm weights / class + position embeddings accumulate their
 * gradients in fp32 side buffers (`*_f32`), weight matrices in bf16 (accumulate != 0: added to what is there). */
typedef struct merlin_vit_weights {
  int num_layers, hidden, heads, intermediate, image_size, patch_size, k_patch_padded;
  float ln_eps;
  const void *class_embedding, *patch_weight, *position_embedding, *pre_ln_w, *pre_ln_b;
  const merlin_vit_layer_weights* layers; /* [num_layers] */
} merlin_vit_weights;
typedef struct merlin_vit_layer_grads {
  void *qkv_w, *out_w, *fc1_w, *fc2_w;
  float *qkv_b_f32, *out_b_f32, *ln1_w_f32, *ln1_b_f32, *ln2_w_f32, *ln2_b_f32, *fc1_b_f32, *fc2_b_f32;
} merlin_vit_layer_grads;
typedef struct merlin_vit_grads {
  void* patch_weight;
  float *position_f32, *class_f32, *pre_ln_w_f32, *pre_ln_b_f32;
  const merlin_vit_layer_grads* layers; /* [num_layers] */
} merlin_vit_grads;
size_t merlin_vit_workspace_bytes(const merlin_vit_weights* w, int n_img);
int merlin_vit_forward(void* stream, const merlin_vit_weights* w, const void* pixels, int n_img, void* workspace,
                       void* hidden, int cta_group);
size_t merlin_vit_saved_bytes(const merlin_vit_weights* w, int n_img);
int merlin_vit_forward_train(void* stream, const merlin_vit_weights* w, const void* pixels, int n_img, void* saved,
                             void* hidden, int cta_group);
size_t merlin_vit_backward_workspace_bytes(const merlin_vit_weights* w, int n_img);
/* dh [n_img*tokens, hidden]: gradient w.r.t. merlin_vit_forward_train's `hidden`; overwritten. */
int merlin_vit_backward(void* stream, const merlin_vit_weights* w, const merlin_vit_grads* g, void* saved, void* dh,
                        void* workspace, int n_img, int accumulate, int cta_group);

/* ---- data-parallel gradient exchange (SURVEY.md §8b minimum set: merlin_dp_init / allreduce_bucket / finalize).
 * Replaces the FSDP `shard_grad_op auto_wrap` reduce-scatter / all-gather the reference trains with
 * (playground/merlin/clip-large+conv+vicuna-v15-7b/pretrain.sh:32-33 inside HF Trainer) by plain
 * data parallelism: one NCCL communicator per rank over NVLink 5 / NVSwitch and an in-place bf16 SUM
 * all-reduce per contiguous bucket of the flat gradient arena, enqueued on the caller's (side) stream.
 * The host exchanges the 128-byte unique id (rank 0 creates it) over whatever channel it has
 * (torch.distributed's store in the Python shim).  max_ctas > 0 bounds the CTAs NCCL may use per collective so
 * that the exchange co-resides with backward's persistent GEMMs (see merlin_set_sm_budget). */
#define MERLIN_DP_UNIQUE_ID_BYTES 128
/* Optional, before any other merlin_dp_* call: resolve NCCL from this shared object (the Python shim passes
 * the torch-bundled libnccl.so.2 so both use ONE library); default = the libnccl.so.2 already in the process. */
int merlin_dp_load_nccl(const char* path);
int merlin_dp_unique_id(void* id_out);
int merlin_dp_init(const void* unique_id, int rank, int world, int max_ctas, void** comm_out);
int merlin_dp_allreduce_bucket(void* comm, void* stream, void* grads_bf16, int64_t count);
/* fp32 SUM (op_max == 0) or MAX all-reduce of a small buffer: loss / timing reductions of the step loop. */
int merlin_dp_allreduce_f32(void* comm, void* stream, void* buf, int64_t count, int op_max);
/* Optional: NCCL-owned allocation (ncclMemAlloc) + user-buffer registration of the gradient arena, which lets
 * NVLS reduce in the switch straight out of the arena (no staging copies, fewer CTAs). */
int merlin_dp_mem_alloc(void** ptr_out, int64_t bytes);
int merlin_dp_mem_free(void* ptr);
int merlin_dp_register(void* comm, void* buf, int64_t bytes, void** handle_out);
int merlin_dp_deregister(void* comm, void* handle);
int merlin_dp_finalize(void* comm);
int merlin_dp_nccl_version(void); /* NCCL_VERSION_CODE of the resolved libnccl, 0 if none */

/* SM budget of the persistent kernels (the tcgen05 GEMM runs one CTA / CTA pair per SM): `sms` > 0 makes their
 * grids cover at most that many SMs, 0 restores the whole chip; returns the previous value.  Process-wide. */
int merlin_set_sm_budget(int sms);
int merlin_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MERLIN_B200_H_ */
