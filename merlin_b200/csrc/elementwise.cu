// merlin_b200 — HBM-bound kernels of the hot path: norms, rotary embedding, im2col gathers,
// CLIP embedding assembly, embedding gather + image-feature splice, cross-entropy, argmax.
// All are one-pass, 128-bit vectorised and coalesced; fp32 math, bf16 I/O.
#include "host.cuh"
#include "ptx.cuh"

namespace mb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for up to 1024 threads; `red` is 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) {
    t = warp_max(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                    pack_bf16(f[6], f[7]));
}

constexpr int kNormThreads = 256;
constexpr int kNormMaxVec = 4;  // 256 threads x 4 vectors x 8 elements = rows up to 8192 wide

// y = w · (x · rsqrt(mean(x²) + eps)), fp32 math, one rounding on output
// (HF 4.31 LlamaRMSNorm.forward; transformers/models/llama/modeling_llama.py:62-67 in 5.x).
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ y, int64_t ldy, int D, float eps) {
  __shared__ float red[32];
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(blockIdx.x) * ldx);
  uint4 v[kNormMaxVec];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / static_cast<float>(D) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(blockIdx.x) * ldy);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      float f[8], g[8];
      unpack8(v[i], f);
      unpack8(__ldg(wr + idx), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = g[e] * (f[e] * rstd);
      yr[idx] = pack8(f);
    }
  }
}

// y = (x - mean) · rsqrt(var + eps) · w + b   (nn.LayerNorm; CLIP layer_norm1/2, pre_layrnorm)
__global__ void __launch_bounds__(kNormThreads)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int64_t ldy, int D,
                 float eps) {
  __shared__ float red[32];
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(blockIdx.x) * ldx);
  uint4 v[kNormMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += f[e];
    }
  }
  const float mean = block_sum(s, red) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) sq += (f[e] - mean) * (f[e] - mean);
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / static_cast<float>(D) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(blockIdx.x) * ldy);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      float f[8], g[8], h[8];
      unpack8(v[i], f);
      unpack8(__ldg(wr + idx), g);
      unpack8(__ldg(br + idx), h);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (f[e] - mean) * rstd * g[e] + h[e];
      yr[idx] = pack8(f);
    }
  }
}

// CLIP embeddings + pre_layrnorm: row (img, t) = LN( (t == 0 ? class_embedding : patch_out[img, t-1])
// + position_embedding[t] )   (HF:clip/modeling_clip.py:202-218, 676-677).  D <= 8192.
__global__ void __launch_bounds__(kNormThreads)
vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ patch_out, const __nv_bfloat16* __restrict__ cls,
                    const __nv_bfloat16* __restrict__ pos, const __nv_bfloat16* __restrict__ w,
                    const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int tokens,
                    int D, float eps) {
  __shared__ float red[32];
  const int nvec = D >> 3;
  const int img = blockIdx.x / tokens, t = blockIdx.x % tokens;
  const uint4* src = (t == 0) ? reinterpret_cast<const uint4*>(cls)
                              : reinterpret_cast<const uint4*>(
                                    patch_out + (static_cast<size_t>(img) * (tokens - 1) + (t - 1)) * D);
  const uint4* pr = reinterpret_cast<const uint4*>(pos + static_cast<size_t>(t) * D);
  float vals[kNormMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      float f[8], g[8];
      unpack8(src[idx], f);
      unpack8(pr[idx], g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        // the reference adds in bf16: round the sum like torch does
        vals[i][e] = __bfloat162float(__float2bfloat16_rn(f[e] + g[e]));
        s += vals[i][e];
      }
    }
  }
  const float mean = block_sum(s, red) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) sq += (vals[i][e] - mean) * (vals[i][e] - mean);
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / static_cast<float>(D) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(blockIdx.x) * D);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kNormThreads;
    if (idx < nvec) {
      float g[8], h[8], o[8];
      unpack8(__ldg(wr + idx), g);
      unpack8(__ldg(br + idx), h);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (vals[i][e] - mean) * rstd * g[e] + h[e];
      yr[idx] = pack8(o);
    }
  }
}

// Patch-embedding im2col: pixels [n, 3, HW, HW] -> col [n*G*G, ld] with k = (c*P + ky)*P + kx,
// zero padded to ld (conv weight [d,3,P,P] flattens to the same k order).
__global__ void patch_im2col_kernel(const __nv_bfloat16* __restrict__ px, __nv_bfloat16* __restrict__ col,
                                    int n_img, int HW, int P, int ld) {
  const int G = HW / P;
  const int row = blockIdx.x;  // (img, gy, gx)
  const int img = row / (G * G), gy = (row / G) % G, gx = row % G;
  const int K = 3 * P * P;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) {
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (k < K) {
      const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
      v = px[((static_cast<size_t>(img) * 3 + c) * HW + (gy * P + ky)) * HW + gx * P + kx];
    }
    col[static_cast<size_t>(row) * ld + k] = v;
  }
}

// ConvProjector im2col (3x3, pad 1, stride s) over channel-last ViT features:
// feat row (img, y, x) lives at feat + ((img*tok_stride + tok_off + y*G + x) * C).
// col [n*Go*Go, 9*C] with k = (ky*3 + kx)*C + c  (weight repacked to [D, 3, 3, C]).
__global__ void conv3x3_im2col_kernel(const __nv_bfloat16* __restrict__ feat, __nv_bfloat16* __restrict__ col,
                                      int G, int Go, int C, int stride, int tok_stride, int tok_off) {
  const int row = blockIdx.x;  // (img, oy, ox)
  const int tap = blockIdx.y;  // ky*3 + kx
  const int img = row / (Go * Go), oy = (row / Go) % Go, ox = row % Go;
  const int y = oy * stride - 1 + tap / 3, x = ox * stride - 1 + tap % 3;
  const bool inside = y >= 0 && y < G && x >= 0 && x < G;
  const uint4* src = reinterpret_cast<const uint4*>(
      feat + (static_cast<size_t>(img) * tok_stride + tok_off + (inside ? y * G + x : 0)) * C);
  uint4* dst = reinterpret_cast<uint4*>(col + (static_cast<size_t>(row) * 9 + tap) * C);
  for (int i = threadIdx.x; i < (C >> 3); i += blockDim.x)
    dst[i] = inside ? src[i] : make_uint4(0, 0, 0, 0);
}

// Per sample b (one block each): exclusive rank of every <im_patch> token inside the sample, offset by
// feat_offset[b] into the flat feature buffer; tokens that are not <im_patch>, or whose rank exceeds
// feat_limit[b] (more placeholders than feature rows), get -1.  counts[b] = #<im_patch> in sample b.
__global__ void __launch_bounds__(1024)
patch_rank_kernel(const int64_t* __restrict__ ids, int* __restrict__ rank, int* __restrict__ counts, int S,
                  int64_t patch_id, const int* __restrict__ feat_offset, const int* __restrict__ feat_limit) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int64_t* row_ids = ids + static_cast<size_t>(b) * S;
  int* row_rank = rank + static_cast<size_t>(b) * S;
  const int off = feat_offset ? feat_offset[b] : 0;
  const int lim = feat_limit ? feat_limit[b] : 0x7fffffff;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < S; base += 1024) {
    const int t = base + threadIdx.x;
    const int flag = (t < S && row_ids[t] == patch_id) ? 1 : 0;
    int incl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += n;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += n;
      }
      warp_tot[threadIdx.x] = w;  // inclusive over warps
    }
    __syncthreads();
    const int warp_excl = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    const int excl = carry + warp_excl + incl - flag;
    if (t < S) row_rank[t] = (flag && excl < lim) ? off + excl : -1;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = carry;
}

// inputs_embeds[t] = (rank[t] >= 0) ? image_features[rank[t]] : embed_tokens[ids[t]]
// (mmgpt/model/mmgpt/base_mmgpt.py:99-160: embed gather + per-image torch.cat splice).
__global__ void embed_splice_kernel(const int64_t* __restrict__ ids, const int* __restrict__ rank,
                                    const __nv_bfloat16* __restrict__ table,
                                    const __nv_bfloat16* __restrict__ feats, __nv_bfloat16* __restrict__ out,
                                    int D, int n_feat_rows, int64_t vocab) {
  const int t = blockIdx.x;
  const int r = rank[t];
  const uint4* src;
  if (r >= 0 && r < n_feat_rows) {
    src = reinterpret_cast<const uint4*>(feats + static_cast<size_t>(r) * D);
  } else {
    int64_t id = ids[t];
    if (id < 0 || id >= vocab) id = 0;
    src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * D);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * D);
  for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = src[i];
}

// Half-split rotary embedding in place on the q and k blocks of packed qkv rows (head_dim 128).
__global__ void rope_qk_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, const float* __restrict__ cosT,
                               const float* __restrict__ sinT, const int* __restrict__ positions,
                               int H, int seq_len, float sin_sign) {
  const int row = blockIdx.x;
  const int pos = positions ? positions[row] : row % seq_len;
  const float* cp = cosT + static_cast<size_t>(pos) * 64;
  const float* sp = sinT + static_cast<size_t>(pos) * 64;
  __nv_bfloat16* base = qkv + static_cast<size_t>(row) * ld;
  // 2H heads (q then k), 64 pairs each
  for (int i = threadIdx.x; i < 2 * H * 64; i += blockDim.x) {
    const int head = i >> 6, j = i & 63;
    __nv_bfloat16* h = base + head * 128;
    const float a = __bfloat162float(h[j]), b = __bfloat162float(h[j + 64]);
    const float c = cp[j], s = sp[j] * sin_sign;
    h[j] = __float2bfloat16_rn(a * c - b * s);
    h[j + 64] = __float2bfloat16_rn(b * c + a * s);
  }
}

// Shifted cross-entropy rows: row t of sample b predicts labels[b, t+1]
// (mmgpt/model/mmgpt/llama_mmgpt.py:89-100).  loss_sum += nll, count += 1 for labels != ignore.
__global__ void __launch_bounds__(1024)
cross_entropy_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                     int S, int V, int64_t ignore_index, float* __restrict__ row_nll,
                     float* __restrict__ accum /* [2] = {sum, count} */) {
  __shared__ float red[32];
  const int row = blockIdx.x;  // b*S + t
  const int t = row % S;
  int64_t label = ignore_index;
  if (t + 1 < S) label = labels[row + 1];
  if (label == ignore_index || label < 0 || label >= V) {
    if (row_nll && threadIdx.x == 0) row_nll[row] = 0.f;
    return;
  }
  const __nv_bfloat16* lr = logits + static_cast<size_t>(row) * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, __bfloat162float(lr[i]));
  mx = block_max(mx, red);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) se += __expf(__bfloat162float(lr[i]) - mx);
  se = block_sum(se, red);
  if (threadIdx.x == 0) {
    const float nll = logf(se) + mx - __bfloat162float(lr[label]);
    if (row_nll) row_nll[row] = nll;
    atomicAdd(&accum[0], nll);
    atomicAdd(&accum[1], 1.f);
  }
}

__global__ void __launch_bounds__(1024)
argmax_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, int V, int64_t* __restrict__ out) {
  __shared__ float smax[32];
  __shared__ int sidx[32];
  const __nv_bfloat16* lr = logits + static_cast<size_t>(blockIdx.x) * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  // 16-byte loads when the row is aligned (all loads of a thread are issued before the first compare: the scan of
  // one 32003-entry row dropped from 17.8 to a few us); indices grow within a thread, so `>` keeps the lowest one
  const bool vec = (reinterpret_cast<uintptr_t>(lr) & 15) == 0;
  const int nvec = vec ? V >> 3 : 0;
  constexpr int kMaxIt = 4;
  for (int base = 0; base < nvec; base += kMaxIt * 1024) {
    uint4 q[kMaxIt];
#pragma unroll
    for (int u = 0; u < kMaxIt; ++u) {
      const int i = base + u * 1024 + threadIdx.x;
      q[u] = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf pairs
      if (i < nvec) q[u] = reinterpret_cast<const uint4*>(lr)[i];
    }
#pragma unroll
    for (int u = 0; u < kMaxIt; ++u) {
      const int i = base + u * 1024 + threadIdx.x;
      const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xffff0000u);
        if (lo > best) {
          best = lo;
          bi = 8 * i + 2 * e;
        }
        if (hi > best) {
          best = hi;
          bi = 8 * i + 2 * e + 1;
        }
      }
    }
  }
  for (int i = 8 * nvec + threadIdx.x; i < V; i += blockDim.x) {
    const float v = __bfloat162float(lr[i]);
    if (v > best) {  // strided scan: first maximum per thread is its smallest index
      best = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = best;
    sidx[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? smax[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < nw ? sidx[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = bi;  // torch.argmax tie-break: lowest index
  }
}

}  // namespace mb

using namespace mb;

extern "C" {

int merlin_rmsnorm(void* stream, const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy,
                   int M, int D, float eps) {
  MB_CHECK_ARG(x && weight && y, "rmsnorm: null tensor");
  MB_CHECK_ARG(M >= 0 && D > 0 && D % 8 == 0 && D <= 8 * kNormThreads * kNormMaxVec && ldx % 8 == 0 && ldy % 8 == 0,
               "rmsnorm: need D %% 8 == 0, D <= 8192 and 16-byte row pitch (D=%d)", D);
  if (M == 0) return MERLIN_OK;
  rmsnorm_kernel<<<M, kNormThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(weight),
      static_cast<__nv_bfloat16*>(y), ldy, D, eps);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_layernorm(void* stream, const void* x, int64_t ldx, const void* weight, const void* bias,
                     void* y, int64_t ldy, int M, int D, float eps) {
  MB_CHECK_ARG(x && weight && bias && y, "layernorm: null tensor");
  MB_CHECK_ARG(M >= 0 && D > 0 && D % 8 == 0 && D <= 8 * kNormThreads * kNormMaxVec && ldx % 8 == 0 && ldy % 8 == 0,
               "layernorm: need D %% 8 == 0, D <= 8192 and 16-byte row pitch (D=%d)", D);
  if (M == 0) return MERLIN_OK;
  layernorm_kernel<<<M, kNormThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), ldy, D, eps);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_patch_im2col(void* stream, const void* pixels, void* col, int n_img, int image_size,
                        int patch_size, int ld_col) {
  MB_CHECK_ARG(pixels && col, "patch_im2col: null tensor");
  MB_CHECK_ARG(n_img > 0 && patch_size > 0 && image_size % patch_size == 0 &&
                   ld_col >= 3 * patch_size * patch_size && ld_col % 8 == 0,
               "patch_im2col: bad geometry image=%d patch=%d ld=%d", image_size, patch_size, ld_col);
  const int G = image_size / patch_size;
  patch_im2col_kernel<<<n_img * G * G, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(pixels), static_cast<__nv_bfloat16*>(col), n_img, image_size,
      patch_size, ld_col);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_vit_embed_ln(void* stream, const void* patch_out, const void* class_embedding,
                        const void* position_embedding, const void* ln_weight, const void* ln_bias,
                        void* hidden, int n_img, int tokens, int D, float eps) {
  MB_CHECK_ARG(patch_out && class_embedding && position_embedding && ln_weight && ln_bias && hidden,
               "vit_embed_ln: null tensor");
  MB_CHECK_ARG(n_img > 0 && tokens > 1 && D % 8 == 0 && D <= 8 * kNormThreads * kNormMaxVec,
               "vit_embed_ln: bad shape tokens=%d D=%d", tokens, D);
  vit_embed_ln_kernel<<<n_img * tokens, kNormThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch_out), static_cast<const __nv_bfloat16*>(class_embedding),
      static_cast<const __nv_bfloat16*>(position_embedding), static_cast<const __nv_bfloat16*>(ln_weight),
      static_cast<const __nv_bfloat16*>(ln_bias), static_cast<__nv_bfloat16*>(hidden), tokens, D, eps);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_conv3x3_im2col(void* stream, const void* feat, void* col, int n_img, int grid, int C,
                          int stride, int tok_stride, int tok_off) {
  MB_CHECK_ARG(feat && col, "conv3x3_im2col: null tensor");
  MB_CHECK_ARG(n_img > 0 && grid > 0 && C % 8 == 0 && stride >= 1, "conv3x3_im2col: bad geometry");
  const int Go = (grid + 2 - 3) / stride + 1;
  dim3 g(static_cast<unsigned>(n_img * Go * Go), 9);
  conv3x3_im2col_kernel<<<g, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(feat), static_cast<__nv_bfloat16*>(col), grid, Go, C, stride,
      tok_stride, tok_off);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_embed_splice(void* stream, const int64_t* input_ids, const void* embed_table,
                        const void* image_features, void* inputs_embeds, int* rank_ws, int B, int S, int D,
                        int n_feature_rows, const int* feat_offset, const int* feat_limit, int64_t vocab,
                        int64_t im_patch_token) {
  MB_CHECK_ARG(input_ids && embed_table && inputs_embeds && rank_ws, "embed_splice: null tensor");
  MB_CHECK_ARG(B >= 0 && S >= 0 && D % 8 == 0, "embed_splice: D %% 8 != 0");
  MB_CHECK_ARG(n_feature_rows == 0 || image_features, "embed_splice: feature rows without a tensor");
  if (B * S == 0) return MERLIN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  patch_rank_kernel<<<B, 1024, 0, st>>>(input_ids, rank_ws, rank_ws + static_cast<size_t>(B) * S, S,
                                        im_patch_token, feat_offset, feat_limit);
  embed_splice_kernel<<<B * S, 256, 0, st>>>(input_ids, rank_ws, static_cast<const __nv_bfloat16*>(embed_table),
                                             static_cast<const __nv_bfloat16*>(image_features),
                                             static_cast<__nv_bfloat16*>(inputs_embeds), D, n_feature_rows,
                                             vocab);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_rope_qk_inplace_ex(void* stream, void* qkv, int64_t ld, const float* rope_cos,
                              const float* rope_sin, const int* positions, int M, int H, int seq_len,
                              int inverse) {
  MB_CHECK_ARG(qkv && rope_cos && rope_sin && M >= 0 && H > 0 && seq_len > 0, "rope: bad argument");
  if (M == 0) return MERLIN_OK;
  rope_qk_kernel<<<M, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(qkv), ld, rope_cos, rope_sin, positions, H, seq_len, inverse ? -1.f : 1.f);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_rope_qk_inplace(void* stream, void* qkv, int64_t ld, const float* rope_cos,
                           const float* rope_sin, const int* positions, int M, int H, int seq_len) {
  return merlin_rope_qk_inplace_ex(stream, qkv, ld, rope_cos, rope_sin, positions, M, H, seq_len, 0);
}

int merlin_cross_entropy_fwd(void* stream, const void* logits, int64_t ld, const int64_t* labels,
                             int B, int S, int V, int64_t ignore_index, float* row_nll,
                             float* loss_sum_count) {
  MB_CHECK_ARG(logits && labels && loss_sum_count, "cross_entropy: null tensor");
  MB_CHECK_ARG(B > 0 && S > 0 && V > 0, "cross_entropy: empty problem");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MB_CHECK_CUDA(cudaMemsetAsync(loss_sum_count, 0, 2 * sizeof(float), st));
  cross_entropy_kernel<<<B * S, 1024, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, labels, S,
                                               V, ignore_index, row_nll, loss_sum_count);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_argmax(void* stream, const void* logits, int64_t ld, int rows, int V, int64_t* out) {
  MB_CHECK_ARG(logits && out && rows > 0 && V > 0, "argmax: bad argument");
  argmax_kernel<<<rows, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), ld, V, out);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

}  // extern "C"
