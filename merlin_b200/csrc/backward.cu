// merlin_b200 — HBM-bound backward / optimizer kernels of the training path:
// RMSNorm backward (+ residual-gradient add), SwiGLU forward-with-save / backward, inverse RoPE,
// cross-entropy backward (in place over the logits), embedding / splice backward, column sums for bias
// gradients, attention delta = rowsum(dO ∘ O), gradient-norm partials and a fused AdamW over a flat
// parameter arena with per-chunk learning-rate / weight-decay tables (LLRD param groups).
#include "host.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <mutex>

namespace mb {

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum_b(float v, float* red) {
  v = warp_sum_b(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum_b(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ void unpack8b(const uint4& q, float (&f)[8]) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                    pack_bf16(f[6], f[7]));
}

// ------------------------------------------------------------------ RMSNorm backward
// y = w ⊙ x̂, x̂ = x·r, r = rsqrt(mean(x²)+eps).  g = w ⊙ dy;  dx = r·(g − x̂·mean(g ⊙ x̂)) (+ dres);
// dw[j] += Σ_rows dy[row,j]·x̂[row,j]   (fp32 atomics, one per column per CTA of the persistent grid).
constexpr int kRnbThreads = 256;
constexpr int kRnbVec = 4;    // D <= 8192
// grid = 4 CTAs per SM, rows strided over the grid: dw/db need one atomic per column per CTA only

__global__ void __launch_bounds__(kRnbThreads)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                   const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dres,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, int M, int D, float eps) {
  __shared__ float red[32];
  const int nvec = D >> 3;
  float wv[kRnbVec][8], dwacc[kRnbVec][8];
#pragma unroll
  for (int i = 0; i < kRnbVec; ++i) {
    const int idx = threadIdx.x + i * kRnbThreads;
    if (idx < nvec) unpack8b(reinterpret_cast<const uint4*>(w)[idx], wv[i]);
#pragma unroll
    for (int e = 0; e < 8; ++e) dwacc[i][e] = 0.f;
  }
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    const size_t off = static_cast<size_t>(row) * D;
    float xv[kRnbVec][8], gv[kRnbVec][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec) {
        unpack8b(reinterpret_cast<const uint4*>(x + off)[idx], xv[i]);
        unpack8b(reinterpret_cast<const uint4*>(dy + off)[idx], gv[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += xv[i][e] * xv[i][e];
      }
    }
    const float rstd = rsqrtf(block_sum_b(ss, red) / static_cast<float>(D) + eps);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = xv[i][e] * rstd;
          dwacc[i][e] += gv[i][e] * xh;        // dy·x̂
          gv[i][e] *= wv[i][e];                // g = w ⊙ dy
          dot += gv[i][e] * xh;
          xv[i][e] = xh;
        }
      }
    }
    const float mean_gx = block_sum_b(dot, red) / static_cast<float>(D);
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec) {
        float o[8];
        if (dres) unpack8b(reinterpret_cast<const uint4*>(dres + off)[idx], o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = rstd * (gv[i][e] - xv[i][e] * mean_gx);
          o[e] = dres ? o[e] + v : v;
        }
        reinterpret_cast<uint4*>(dx + off)[idx] = pack8b(o);
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&dw[idx * 8 + e], dwacc[i][e]);
    }
  }
}

// Warp-per-row variant (the default): no block-wide reductions, 2·NV 16-byte loads in flight per lane, so the
// kernel streams at HBM rate instead of waiting on two __syncthreads() reductions per row.
//   Σx² and Σ(w⊙dy)·x are reduced TOGETHER (one shuffle tree): mean(g⊙x̂) = rstd·Σ(g·x)/D, hence
//   dx = rstd·g − x·(rstd³·Σ(g·x)/D) (+ dres).
// dw: every warp accumulates dy⊙x̂ into its PRIVATE shared-memory copy ([e][vec] layout: conflict-free plain
// read-modify-write, no atomics); the CTA folds its 8 copies and issues one 16-byte red.global.add per 4 columns.
constexpr int kRnwWarps = 8;

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int NV>
__global__ void __launch_bounds__(kRnwWarps * 32)
rmsnorm_bwd_warp_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                        const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dres,
                        __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, int M, int D, float eps) {
  extern __shared__ float dwsm[];   // [kRnwWarps][8][D/8]
  const int nvec = D >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = dwsm + static_cast<size_t>(warp) * D;
  if (dw != nullptr) {
    for (int i = lane; i < D; i += 32) mine[i] = 0.f;
    __syncwarp();
  }
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  const float inv_d = 1.f / static_cast<float>(D);
  for (int row = blockIdx.x * kRnwWarps + warp; row < M; row += gridDim.x * kRnwWarps) {
    const size_t off = static_cast<size_t>(row) * D;
    const uint4* x4 = reinterpret_cast<const uint4*>(x + off);
    const uint4* g4 = reinterpret_cast<const uint4*>(dy + off);
    uint4 xq[NV], gq[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        xq[i] = ld_stream_u4(x4 + idx);
        gq[i] = ld_stream_u4(g4 + idx);
      }
    }
    float ss = 0.f, gx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        float xv[8], gv[8], wv[8];
        unpack8b(xq[i], xv);
        unpack8b(gq[i], gv);
        unpack8b(__ldg(w4 + idx), wv);
        // short independent chains (the two 128-long serial FMA chains were the latency bound of this kernel)
        const float s0 = fmaf(xv[0], xv[0], xv[1] * xv[1]), s1 = fmaf(xv[2], xv[2], xv[3] * xv[3]);
        const float s2 = fmaf(xv[4], xv[4], xv[5] * xv[5]), s3 = fmaf(xv[6], xv[6], xv[7] * xv[7]);
        const float g0 = fmaf(gv[0] * wv[0], xv[0], gv[1] * wv[1] * xv[1]);
        const float g1 = fmaf(gv[2] * wv[2], xv[2], gv[3] * wv[3] * xv[3]);
        const float g2 = fmaf(gv[4] * wv[4], xv[4], gv[5] * wv[5] * xv[5]);
        const float g3 = fmaf(gv[6] * wv[6], xv[6], gv[7] * wv[7] * xv[7]);
        ss += (s0 + s1) + (s2 + s3);
        gx += (g0 + g1) + (g2 + g3);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      gx += __shfl_xor_sync(0xffffffffu, gx, o);
    }
    const float rstd = rsqrtf(ss * inv_d + eps);
    const float c = rstd * rstd * rstd * gx * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        float xv[8], gv[8], wv[8], o[8];
        unpack8b(xq[i], xv);
        unpack8b(gq[i], gv);
        unpack8b(__ldg(w4 + idx), wv);
        if (dres != nullptr) unpack8b(ld_stream_u4(reinterpret_cast<const uint4*>(dres + off) + idx), o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = rstd * (gv[e] * wv[e]) - xv[e] * c;
          o[e] = dres != nullptr ? o[e] + v : v;
        }
        reinterpret_cast<uint4*>(dx + off)[idx] = pack8b(o);
        if (dw != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) mine[e * nvec + idx] += gv[e] * (xv[e] * rstd);
        }
      }
    }
  }
  if (dw == nullptr) return;
  __syncthreads();
  // fold the 8 private copies; element (e, idx) is column idx*8 + e: lanes take 4 consecutive columns
  for (int q = threadIdx.x; q < D / 4; q += kRnwWarps * 32) {
    const int col = q * 4, idx = col >> 3, e0 = col & 7;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ww = 0; ww < kRnwWarps; ++ww) {
      const float* src = dwsm + static_cast<size_t>(ww) * D;
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] += src[(e0 + e) * nvec + idx];
    }
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw + col), "f"(a[0]), "f"(a[1]), "f"(a[2]),
                 "f"(a[3]) : "memory");
  }
}

// ------------------------------------------------------------------ SwiGLU (training form)
// gu [M, 2FF]: per 256 columns, 128 gate then 128 up (the packed gate/up weight layout).
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ mid,
                                  int64_t total_vec /* M*FF/8 */, int FF) {
  const int vec_per_row = FF >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;     // output column
    const int blk = c >> 7, j = c & 127;
    const __nv_bfloat16* base = gu + row * 2 * FF + blk * 256 + j;
    float g[8], u[8], o[8];
    unpack8b(*reinterpret_cast<const uint4*>(base), g);
    unpack8b(*reinterpret_cast<const uint4*>(base + 128), u);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = g[e] / (1.f + __expf(-g[e])) * u[e];
    *reinterpret_cast<uint4*>(mid + row * FF + c) = pack8b(o);
  }
}
// dgu (same interleaved layout, may alias gu): dg = dmid·u·σ(g)(1 + g(1−σ(g))), du = dmid·g·σ(g)
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dmid,
                                  __nv_bfloat16* __restrict__ dgu, int64_t total_vec, int FF) {
  const int vec_per_row = FF >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const int blk = c >> 7, j = c & 127;
    const int64_t off = row * 2 * FF + blk * 256 + j;
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8b(*reinterpret_cast<const uint4*>(gu + off), g);
    unpack8b(*reinterpret_cast<const uint4*>(gu + off + 128), u);
    unpack8b(*reinterpret_cast<const uint4*>(dmid + row * FF + c), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s = 1.f / (1.f + __expf(-g[e]));
      dg[e] = d[e] * u[e] * s * (1.f + g[e] * (1.f - s));
      du[e] = d[e] * g[e] * s;
    }
    *reinterpret_cast<uint4*>(dgu + off) = pack8b(dg);
    *reinterpret_cast<uint4*>(dgu + off + 128) = pack8b(du);
  }
}

// ------------------------------------------------------------------ cross-entropy backward
// logits row (b,t) -> dlogits = (softmax − onehot(label[b,t+1])) · grad_scale / count, or 0 when the
// row is not scored.  In place (dlogits may alias logits).  count = loss_sum_count[1] of the forward.
__global__ void __launch_bounds__(1024)
cross_entropy_bwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                         int S, int V, int64_t ignore_index, const float* __restrict__ sum_count,
                         float grad_scale, __nv_bfloat16* __restrict__ dlogits, int64_t ldd) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int t = row % S;
  int64_t label = ignore_index;
  if (t + 1 < S) label = labels[row + 1];
  const __nv_bfloat16* lr = logits + static_cast<size_t>(row) * ld;
  __nv_bfloat16* dr = dlogits + static_cast<size_t>(row) * ldd;
  if (label == ignore_index || label < 0 || label >= V) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) dr[i] = __float2bfloat16(0.f);
    return;
  }
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, __bfloat162float(lr[i]));
  // block max via the sum helper's scratch
  mx = warp_sum_b(0.f) * 0.f + mx;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  float m2 = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  __syncthreads();
  if (threadIdx.x == 0) red[0] = m2;
  __syncthreads();
  mx = red[0];
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) se += __expf(__bfloat162float(lr[i]) - mx);
  se = block_sum_b(se, red);
  const float scale = grad_scale / sum_count[1];
  const float inv = 1.f / se;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    float p = __expf(__bfloat162float(lr[i]) - mx) * inv;
    if (i == label) p -= 1.f;
    dr[i] = __float2bfloat16_rn(p * scale);
  }
}

// ------------------------------------------------------------------ embedding / splice backward
// dx [T, D]: rows of <im_patch> tokens go to dfeats[rank] (unique destination), the others are
// accumulated into dtable[ids[t]] (bf16x2 atomics; duplicates are rare).
__global__ void embed_splice_bwd_kernel(const int64_t* __restrict__ ids, const int* __restrict__ rank,
                                        const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dtable,
                                        __nv_bfloat16* __restrict__ dfeats, int D, int n_feat_rows, int64_t vocab) {
  const int t = blockIdx.x;
  const int r = rank ? rank[t] : -1;
  const __nv_bfloat16* src = dx + static_cast<size_t>(t) * D;
  if (r >= 0 && r < n_feat_rows) {
    if (dfeats == nullptr) return;
    uint4* dst = reinterpret_cast<uint4*>(dfeats + static_cast<size_t>(r) * D);
    for (int i = threadIdx.x; i < (D >> 3); i += blockDim.x) dst[i] = reinterpret_cast<const uint4*>(src)[i];
    return;
  }
  if (dtable == nullptr) return;
  int64_t id = ids[t];
  if (id < 0 || id >= vocab) return;
  __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(dtable + static_cast<size_t>(id) * D);
  const __nv_bfloat162* s2 = reinterpret_cast<const __nv_bfloat162*>(src);
  for (int i = threadIdx.x; i < (D >> 1); i += blockDim.x) atomicAdd(&dst[i], s2[i]);
}

// out[j] += Σ_rows x[row, j]   (bias gradients); fp32 accumulation buffer
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ out, int M,
                              int N, int rows_per_block) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  float acc = 0.f;
  for (int r = r0; r < min(M, r0 + rows_per_block); ++r) acc += __bfloat162float(x[static_cast<size_t>(r) * ld + col]);
  atomicAdd(&out[col], acc);
}

// delta[b,h,s] = Σ_d dO[b,s,h,d]·O[b,s,h,d]     (flash-attention backward pre-pass)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  float* __restrict__ delta, int S, int H, int hd, int64_t ld) {
  // one warp per (row, head)
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row = gw / H, h = gw % H;  // row = b*S + s
  const int64_t off = static_cast<int64_t>(row) * ld + h * hd;
  float acc = 0.f;
  for (int d = lane * 2; d < hd; d += 64) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(o + off + d);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(d_o + off + d);
    acc += __low2float(a) * __low2float(b) + __high2float(a) * __high2float(b);
  }
  acc = warp_sum_b(acc);
  if (lane == 0) {
    const int b = row / S, s = row % S;
    delta[(static_cast<size_t>(b) * H + h) * S + s] = acc;
  }
}

// ------------------------------------------------------------------ optimizer
// Σ g² over the flat gradient arena.  Block partials (fixed in-block order) are accumulated as 64-bit
// FIXED-POINT integers, so the total is independent of block scheduling order: every data-parallel rank
// derives a bit-identical clip factor from bit-identical (all-reduced) gradients and replicas never drift.
constexpr float kSumsqScale = 1099511627776.f;  // 2^40: resolution 9e-13, range 1.6e7
__global__ void sumsq_finalize_kernel(float* __restrict__ out) {
  out[0] = static_cast<float>(static_cast<double>(*reinterpret_cast<unsigned long long*>(out + 2)) /
                              static_cast<double>(kSumsqScale));
}
__global__ void sumsq_kernel(const __nv_bfloat16* __restrict__ g, int64_t n_vec, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8b(reinterpret_cast<const uint4*>(g)[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += f[e] * f[e];
  }
  acc = block_sum_b(acc, red);
  if (threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(out + 2),
              static_cast<unsigned long long>(static_cast<double>(acc) * static_cast<double>(kSumsqScale) + 0.5));
}

// Fused AdamW over a flat arena: p, g, m, v are bf16 (the reference keeps bf16 weights AND optimizer
// state, mmgpt/model/builder.py:110-114,163); math in fp32.  chunk c = elements [c*chunk, (c+1)*chunk)
// carries its own lr multiplier and weight decay (LLRD groups, llrd_utils.py:25-79).
// torch.optim.AdamW semantics: p *= 1 − lr·wd; m = β1 m + (1−β1) g; v = β2 v + (1−β2) g²;
// p −= lr/bc1 · m / (sqrt(v)/sqrt(bc2) + eps).  g is first scaled by grad_scale·clip where
// clip = min(1, max_norm / (sqrt(sumsq[0])·grad_scale + 1e-6)) (torch clip_grad_norm_).
__global__ void adamw_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                             __nv_bfloat16* __restrict__ m, __nv_bfloat16* __restrict__ v,
                             const float* __restrict__ lr_mult, const float* __restrict__ wd, int chunk_vec,
                             int64_t n_vec, float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                             float grad_scale, float max_norm, const float* __restrict__ sumsq) {
  float clip = grad_scale;
  if (max_norm > 0.f && sumsq != nullptr) {
    const float norm = sqrtf(sumsq[0]) * grad_scale;
    clip = grad_scale * fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t c = i / chunk_vec;
    const float mult = lr_mult[c];
    if (mult == 0.f) continue;  // frozen segment (lr multiplier 0); lr == 0 alone (warm-up step 0) still updates m, v
    const float lr_c = lr * mult;
    const float wd_c = wd[c];
    float pf[8], gf[8], mf[8], vf[8];
    unpack8b(reinterpret_cast<const uint4*>(p)[i], pf);
    unpack8b(reinterpret_cast<const uint4*>(g)[i], gf);
    unpack8b(reinterpret_cast<const uint4*>(m)[i], mf);
    unpack8b(reinterpret_cast<const uint4*>(v)[i], vf);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gg = gf[e] * clip;
      mf[e] = beta1 * mf[e] + (1.f - beta1) * gg;
      vf[e] = beta2 * vf[e] + (1.f - beta2) * gg * gg;
      const float denom = sqrtf(vf[e]) * inv_sqrt_bc2 + eps;
      pf[e] = pf[e] * (1.f - lr_c * wd_c) - (lr_c / bc1) * (mf[e] / denom);
    }
    reinterpret_cast<uint4*>(p)[i] = pack8b(pf);
    reinterpret_cast<uint4*>(m)[i] = pack8b(mf);
    reinterpret_cast<uint4*>(v)[i] = pack8b(vf);
  }
}

// y = bf16(x_fp32)  (fp32 gradient accumulators -> bf16 gradient arena)
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

static int grid_for(int64_t work_items, int threads) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace mb

using namespace mb;

template <int NV>
static int launch_rmsnorm_bwd_warp(cudaStream_t st, const void* x, const void* weight, const void* dy, const void* dres,
                                   void* dx, float* dweight, int M, int D, float eps) {
  const size_t smem = static_cast<size_t>(kRnwWarps) * D * sizeof(float);
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(rmsnorm_bwd_warp_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kRnwWarps * NV * 256 * static_cast<int>(sizeof(float)));
  });
  MB_CHECK_CUDA(attr_rc);
  const int grid = std::min((M + kRnwWarps - 1) / kRnwWarps, sm_count());
  rmsnorm_bwd_warp_kernel<NV><<<grid, kRnwWarps * 32, smem, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(dres),
      static_cast<__nv_bfloat16*>(dx), dweight, M, D, eps);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

extern "C" {

int merlin_rmsnorm_bwd(void* stream, const void* x, const void* weight, const void* dy, const void* dres,
                       void* dx, float* dweight, int M, int D, float eps) {
  MB_CHECK_ARG(x && weight && dy && dx, "rmsnorm_bwd: null tensor");
  MB_CHECK_ARG(D % 8 == 0 && D <= 8 * kRnbThreads * kRnbVec, "rmsnorm_bwd: D=%d unsupported", D);
  if (M == 0) return MERLIN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // warp-per-row kernel for every width the decoder uses (registers hold the row: D <= 5120), D % 32 == 0 so that
  // the 16-byte dw reductions stay aligned; anything else takes the CTA-per-row kernel
  if (D % 32 == 0 && D <= 256) return launch_rmsnorm_bwd_warp<1>(st, x, weight, dy, dres, dx, dweight, M, D, eps);
  if (D % 32 == 0 && D <= 1024) return launch_rmsnorm_bwd_warp<4>(st, x, weight, dy, dres, dx, dweight, M, D, eps);
  if (D % 32 == 0 && D <= 2048) return launch_rmsnorm_bwd_warp<8>(st, x, weight, dy, dres, dx, dweight, M, D, eps);
  if (D % 32 == 0 && D <= 4096) return launch_rmsnorm_bwd_warp<16>(st, x, weight, dy, dres, dx, dweight, M, D, eps);
  if (D % 32 == 0 && D <= 5120) return launch_rmsnorm_bwd_warp<20>(st, x, weight, dy, dres, dx, dweight, M, D, eps);
  rmsnorm_bwd_kernel<<<std::min(M, sm_count() * 4), kRnbThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(dres),
      static_cast<__nv_bfloat16*>(dx), dweight, M, D, eps);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_swiglu_fwd(void* stream, const void* gate_up, void* mid, int M, int FF) {
  MB_CHECK_ARG(gate_up && mid && FF % 128 == 0, "swiglu_fwd: bad argument");
  const int64_t nvec = static_cast<int64_t>(M) * FF / 8;
  if (nvec == 0) return MERLIN_OK;
  swiglu_fwd_kernel<<<grid_for(nvec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(gate_up), static_cast<__nv_bfloat16*>(mid), nvec, FF);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_swiglu_bwd(void* stream, const void* gate_up, const void* dmid, void* dgate_up, int M, int FF) {
  MB_CHECK_ARG(gate_up && dmid && dgate_up && FF % 128 == 0, "swiglu_bwd: bad argument");
  const int64_t nvec = static_cast<int64_t>(M) * FF / 8;
  if (nvec == 0) return MERLIN_OK;
  swiglu_bwd_kernel<<<grid_for(nvec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(gate_up), static_cast<const __nv_bfloat16*>(dmid),
      static_cast<__nv_bfloat16*>(dgate_up), nvec, FF);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_cross_entropy_bwd(void* stream, const void* logits, int64_t ld, const int64_t* labels, int B, int S,
                             int V, int64_t ignore_index, const float* loss_sum_count, float grad_scale,
                             void* dlogits, int64_t ldd) {
  MB_CHECK_ARG(logits && labels && loss_sum_count && dlogits, "cross_entropy_bwd: null tensor");
  cross_entropy_bwd_kernel<<<B * S, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), ld, labels, S, V, ignore_index, loss_sum_count, grad_scale,
      static_cast<__nv_bfloat16*>(dlogits), ldd);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_embed_splice_bwd(void* stream, const int64_t* input_ids, const int* rank, const void* dx,
                            void* dembed_table, void* dimage_features, int T, int D, int n_feature_rows,
                            int64_t vocab) {
  MB_CHECK_ARG(input_ids && dx && D % 8 == 0, "embed_splice_bwd: bad argument");
  if (T == 0) return MERLIN_OK;
  embed_splice_bwd_kernel<<<T, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      input_ids, rank, static_cast<const __nv_bfloat16*>(dx), static_cast<__nv_bfloat16*>(dembed_table),
      static_cast<__nv_bfloat16*>(dimage_features), D, n_feature_rows, vocab);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_colsum(void* stream, const void* x, int64_t ld, float* out, int M, int N) {
  MB_CHECK_ARG(x && out, "colsum: null tensor");
  if (M == 0 || N == 0) return MERLIN_OK;
  const int rows_per_block = 64;
  dim3 grid((N + 255) / 256, (M + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), ld,
                                                                     out, M, N, rows_per_block);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_attn_delta(void* stream, const void* out, const void* dout, float* delta, int B, int S, int H,
                      int head_dim, int64_t ld) {
  MB_CHECK_ARG(out && dout && delta, "attn_delta: null tensor");
  const int64_t warps = static_cast<int64_t>(B) * S * H;
  if (warps == 0) return MERLIN_OK;
  const int threads = 256;
  attn_delta_kernel<<<static_cast<unsigned>((warps * 32 + threads - 1) / threads), threads, 0,
                      static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(out),
                                                           static_cast<const __nv_bfloat16*>(dout), delta, S, H,
                                                           head_dim, ld);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_grad_sumsq(void* stream, const void* grads, int64_t n, float* out) {
  MB_CHECK_ARG(grads && out && n % 8 == 0, "grad_sumsq: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 7) == 0, "grad_sumsq: out must be 8-byte aligned (4 floats)");
  MB_CHECK_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(float), st));
  if (n == 0) return MERLIN_OK;
  sumsq_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(grads), n / 8, out);
  sumsq_finalize_kernel<<<1, 1, 0, st>>>(out);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_adamw_step(void* stream, void* params, const void* grads, void* exp_avg, void* exp_avg_sq,
                      const float* lr_mult, const float* weight_decay, int chunk, int64_t n, float lr, float beta1,
                      float beta2, float eps, int step, float grad_scale, float max_grad_norm,
                      const float* grad_sumsq) {
  MB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && lr_mult && weight_decay, "adamw: null tensor");
  MB_CHECK_ARG(chunk % 8 == 0 && chunk > 0 && n % 8 == 0 && step >= 1, "adamw: bad chunking");
  if (n == 0) return MERLIN_OK;
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(params), static_cast<const __nv_bfloat16*>(grads),
      static_cast<__nv_bfloat16*>(exp_avg), static_cast<__nv_bfloat16*>(exp_avg_sq), lr_mult, weight_decay,
      chunk / 8, n / 8, lr, beta1, beta2, eps, bc1, bc2, grad_scale, max_grad_norm, grad_sumsq);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_f32_to_bf16(void* stream, const float* x, void* y, int64_t n) {
  MB_CHECK_ARG(x && y, "f32_to_bf16: null tensor");
  if (n == 0) return MERLIN_OK;
  f32_to_bf16_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

}  // extern "C"

// =====================================================================================================
// Vision-tower backward pieces (CLIP encoder layers, embeddings, conv projector input gradient)
// =====================================================================================================
namespace mb {

// LayerNorm backward: x̂ = (x−μ)·r, y = w⊙x̂ + b.  g = w⊙dy; dx = r·(g − mean(g) − x̂·mean(g⊙x̂)) (+ dres);
// dw += Σ dy⊙x̂, db += Σ dy  (fp32 atomics, one per column per CTA of the persistent grid).
__global__ void __launch_bounds__(kRnbThreads)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dres,
                     __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int M, int D,
                     float eps) {
  __shared__ float red[32];
  const int nvec = D >> 3;
  float wv[kRnbVec][8], dwacc[kRnbVec][8], dbacc[kRnbVec][8];
#pragma unroll
  for (int i = 0; i < kRnbVec; ++i) {
    const int idx = threadIdx.x + i * kRnbThreads;
    if (idx < nvec) unpack8b(reinterpret_cast<const uint4*>(w)[idx], wv[i]);
#pragma unroll
    for (int e = 0; e < 8; ++e) dwacc[i][e] = dbacc[i][e] = 0.f;
  }
  for (int row = blockIdx.x; row < M; row += gridDim.x) {
    const size_t off = static_cast<size_t>(row) * D;
    float xv[kRnbVec][8], gv[kRnbVec][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec) {
        unpack8b(reinterpret_cast<const uint4*>(x + off)[idx], xv[i]);
        unpack8b(reinterpret_cast<const uint4*>(dy + off)[idx], gv[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += xv[i][e];
      }
    }
    const float mean = block_sum_b(s, red) / static_cast<float>(D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec)
#pragma unroll
        for (int e = 0; e < 8; ++e) sq += (xv[i][e] - mean) * (xv[i][e] - mean);
    }
    const float rstd = rsqrtf(block_sum_b(sq, red) / static_cast<float>(D) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[i][e] - mean) * rstd;
          dwacc[i][e] += gv[i][e] * xh;
          dbacc[i][e] += gv[i][e];
          gv[i][e] *= wv[i][e];
          sg += gv[i][e];
          sgx += gv[i][e] * xh;
          xv[i][e] = xh;
        }
      }
    }
    const float mg = block_sum_b(sg, red) / static_cast<float>(D);
    const float mgx = block_sum_b(sgx, red) / static_cast<float>(D);
#pragma unroll
    for (int i = 0; i < kRnbVec; ++i) {
      const int idx = threadIdx.x + i * kRnbThreads;
      if (idx < nvec) {
        float o[8];
        if (dres) unpack8b(reinterpret_cast<const uint4*>(dres + off)[idx], o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = rstd * (gv[i][e] - mg - xv[i][e] * mgx);
          o[e] = dres ? o[e] + v : v;
        }
        reinterpret_cast<uint4*>(dx + off)[idx] = pack8b(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kRnbVec; ++i) {
    const int idx = threadIdx.x + i * kRnbThreads;
    if (idx < nvec)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (dw) atomicAdd(&dw[idx * 8 + e], dwacc[i][e]);
        if (db) atomicAdd(&db[idx * 8 + e], dbacc[i][e]);
      }
  }
}

// quick_gelu(x) = x·σ(1.702x);  backward: dx = dy·(σ + 1.702·x·σ·(1−σ)).  dy == nullptr selects the forward.
__global__ void quick_gelu_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                  __nv_bfloat16* __restrict__ out, int64_t n_vec) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float xv[8], d[8], o[8];
    unpack8b(reinterpret_cast<const uint4*>(x)[i], xv);
    if (dy) unpack8b(reinterpret_cast<const uint4*>(dy)[i], d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s = 1.f / (1.f + __expf(-1.702f * xv[e]));
      o[e] = dy ? d[e] * (s + 1.702f * xv[e] * s * (1.f - s)) : xv[e] * s;
    }
    reinterpret_cast<uint4*>(out)[i] = pack8b(o);
  }
}

// Input gradient of the 3x3 / pad 1 / stride s convolution from the im2col-layout gradient
// dcol [n*Go*Go, 9*C] (k = tap*C + c): dfeat(img,y,x) = Σ over the taps that read (y,x).  A gather (no
// atomics); rows are written at (img*tok_stride + tok_off + y*G + x), the layout conv3x3_im2col reads.
__global__ void conv3x3_col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dfeat,
                                      int G, int Go, int C, int stride, int tok_stride, int tok_off) {
  const int pos = blockIdx.x;  // (img, y, x)
  const int img = pos / (G * G), y = (pos / G) % G, x = pos % G;
  for (int i = threadIdx.x; i < (C >> 3); i += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || ty % stride != 0 || ty / stride >= Go) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || tx % stride != 0 || tx / stride >= Go) continue;
        const size_t row = (static_cast<size_t>(img) * Go + ty / stride) * Go + tx / stride;
        float f[8];
        unpack8b(reinterpret_cast<const uint4*>(dcol + (row * 9 + ky * 3 + kx) * C)[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
    reinterpret_cast<uint4*>(dfeat + (static_cast<size_t>(img) * tok_stride + tok_off + y * G + x) * C)[i] = pack8b(acc);
  }
}

// emb[img, t] = (t == 0 ? class_embedding : patch_out[img, t-1]) + position_embedding[t]  (bf16 add, as torch)
__global__ void vit_embed_kernel(const __nv_bfloat16* __restrict__ patch_out, const __nv_bfloat16* __restrict__ cls,
                                 const __nv_bfloat16* __restrict__ pos, __nv_bfloat16* __restrict__ emb, int tokens,
                                 int D) {
  const int img = blockIdx.x / tokens, t = blockIdx.x % tokens;
  const __nv_bfloat16* src = t == 0 ? cls : patch_out + (static_cast<size_t>(img) * (tokens - 1) + (t - 1)) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    emb[static_cast<size_t>(blockIdx.x) * D + c] =
        __float2bfloat16_rn(__bfloat162float(src[c]) + __bfloat162float(pos[static_cast<size_t>(t) * D + c]));
}

// Backward of the CLIP embedding assembly: d[img, t, :] (gradient after pre_layrnorm backward) ->
// dpatch_out[img, t-1] = d[img, t] (t >= 1); dpos[t] += Σ_img d[img, t]; dcls += Σ_img d[img, 0]  (fp32).
__global__ void vit_embed_bwd_kernel(const __nv_bfloat16* __restrict__ d, __nv_bfloat16* __restrict__ dpatch_out,
                                     float* __restrict__ dpos, float* __restrict__ dcls, int n_img, int tokens, int D) {
  const int t = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int img = 0; img < n_img; ++img) {
      const __nv_bfloat16 v = d[(static_cast<size_t>(img) * tokens + t) * D + c];
      acc += __bfloat162float(v);
      if (t > 0) dpatch_out[(static_cast<size_t>(img) * (tokens - 1) + (t - 1)) * D + c] = v;
    }
    dpos[static_cast<size_t>(t) * D + c] += acc;
    if (t == 0) dcls[c] += acc;
  }
}

}  // namespace mb

extern "C" {

int merlin_layernorm_bwd(void* stream, const void* x, const void* weight, const void* dy, const void* dres,
                         void* dx, float* dweight, float* dbias, int M, int D, float eps) {
  MB_CHECK_ARG(x && weight && dy && dx, "layernorm_bwd: null tensor");
  MB_CHECK_ARG(D % 8 == 0 && D <= 8 * kRnbThreads * kRnbVec, "layernorm_bwd: D=%d unsupported", D);
  if (M == 0) return MERLIN_OK;
  layernorm_bwd_kernel<<<std::min(M, sm_count() * 4), kRnbThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(weight),
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(dres),
      static_cast<__nv_bfloat16*>(dx), dweight, dbias, M, D, eps);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_quick_gelu(void* stream, const void* x, const void* dy, void* out, int64_t n) {
  MB_CHECK_ARG(x && out && n % 8 == 0, "quick_gelu: bad argument");
  if (n == 0) return MERLIN_OK;
  quick_gelu_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(out),
      n / 8);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_conv3x3_col2im(void* stream, const void* dcol, void* dfeat, int n_img, int grid, int C, int stride,
                          int tok_stride, int tok_off) {
  MB_CHECK_ARG(dcol && dfeat && C % 8 == 0 && stride >= 1, "conv3x3_col2im: bad argument");
  const int Go = (grid + 2 - 3) / stride + 1;
  conv3x3_col2im_kernel<<<n_img * grid * grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dcol), static_cast<__nv_bfloat16*>(dfeat), grid, Go, C, stride, tok_stride,
      tok_off);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_vit_embed(void* stream, const void* patch_out, const void* class_embedding,
                     const void* position_embedding, void* emb, int n_img, int tokens, int D) {
  MB_CHECK_ARG(patch_out && class_embedding && position_embedding && emb, "vit_embed: null tensor");
  vit_embed_kernel<<<n_img * tokens, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch_out), static_cast<const __nv_bfloat16*>(class_embedding),
      static_cast<const __nv_bfloat16*>(position_embedding), static_cast<__nv_bfloat16*>(emb), tokens, D);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_vit_embed_bwd(void* stream, const void* dhidden, void* dpatch_out, float* dposition, float* dclass,
                         int n_img, int tokens, int D) {
  MB_CHECK_ARG(dhidden && dpatch_out && dposition && dclass, "vit_embed_bwd: null tensor");
  vit_embed_bwd_kernel<<<tokens, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dhidden), static_cast<__nv_bfloat16*>(dpatch_out), dposition, dclass, n_img,
      tokens, D);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

}  // extern "C"
