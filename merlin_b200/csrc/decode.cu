// merlin_b200 — single-token decode path: HBM-bound by design.
//
// One decode step streams every weight once (13.5 GB for Vicuna-7B) plus the KV cache, so all kernels
// here are bandwidth kernels: warp-per-row-pair GEMVs with 128-bit coalesced weight loads, the RMSNorm
// fused as a prologue (each block normalises the 8 KB activation into shared memory), RoPE + KV-cache
// append / SwiGLU / residual-add fused as epilogues, and a split-K flash-decode attention.
// Every kernel reads the current position from device memory, so a whole step (and the greedy loop
// around it) can be captured once in a CUDA graph and replayed.
#include "host.cuh"

#include <mutex>
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace mb {

enum GemvMode : int {
  GEMV_PLAIN = 0,     // out[r] = dot(W[r], x) (+ residual[r])
  GEMV_QKV_ROPE = 1,  // fused q|k|v rows: rotate q,k; q -> qbuf, k/v -> cache row `pos`
  GEMV_SWIGLU = 2,    // W rows interleaved [128 gate | 128 up]; out[j] = silu(g_j) * u_j
};

struct GemvParams {
  const __nv_bfloat16* W;         // [N, K]
  const __nv_bfloat16* x;         // [K] input activation
  const __nv_bfloat16* norm_w;    // [K] RMSNorm weight applied to x first, or nullptr
  const __nv_bfloat16* residual;  // [N] or nullptr (PLAIN)
  __nv_bfloat16* out;             // PLAIN: [N]; SWIGLU: [N/2]; QKV: q buffer [D]
  __nv_bfloat16* kv_row0;         // QKV: cache base, tile-major layout (kv_row_off)
  const float* rope_cos;          // [max_pos, 64]
  const float* rope_sin;
  const int* pos;                 // device scalar: position of the token being decoded
  int N, K, D;
  float eps;
  int ksplit;                     // 1, 2 or 4 warps per row pair
};

#ifndef MERLIN_GEMV_VEC
#define MERLIN_GEMV_VEC 4
#endif
constexpr int kVec = MERLIN_GEMV_VEC;   // 16-byte vectors per row per batch (2*kVec loads in flight per lane)
constexpr int kGemvWarps = 4;
constexpr int kGemvThreads = kGemvWarps * 32;

__device__ __forceinline__ float dot8(const uint4& w, const uint4& x) {
  float s = bf16_lo(w.x) * bf16_lo(x.x);
  s = fmaf(bf16_hi(w.x), bf16_hi(x.x), s);
  s = fmaf(bf16_lo(w.y), bf16_lo(x.y), s);
  s = fmaf(bf16_hi(w.y), bf16_hi(x.y), s);
  s = fmaf(bf16_lo(w.z), bf16_lo(x.z), s);
  s = fmaf(bf16_hi(w.z), bf16_hi(x.z), s);
  s = fmaf(bf16_lo(w.w), bf16_lo(x.w), s);
  s = fmaf(bf16_hi(w.w), bf16_hi(x.w), s);
  return s;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// KV cache layout: blocks of 32 tokens, head-major inside a block: row (h, t) = [k_t^h rotated (128) | v_t^h (128)]
// at element offset (((t / 32) * H + h) * 32 + t % 32) * 256.  One (head, 32-key tile) is 16 KB contiguous for the
// streaming decode kernel's bulk copies, and all heads of a token range still form one dense region.
__device__ __forceinline__ size_t kv_row_off(int h, int t, int H) {
  return ((static_cast<size_t>(t >> 5) * H + h) * 32 + (t & 31)) * 256;
}

template <int kMode>
__global__ void __launch_bounds__(kGemvThreads)
gemv_pair_kernel(const GemvParams p) {
  extern __shared__ uint4 xs[];  // K bf16 (normalised) activation
  __shared__ float red[kGemvWarps];
  __shared__ float part[kGemvWarps][2];
  const int nvec = p.K >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- which two weight rows does this warp own?  (independent of any activation)
  // ksplit adjacent warps share one row pair and split K between them (more CTAs for small N).
  const int ksplit = p.ksplit;
  const int kpart = warp % ksplit;
  const int v0 = (nvec * kpart) / ksplit, v1 = (nvec * (kpart + 1)) / ksplit;
  const int task = (blockIdx.x * kGemvWarps + warp) / ksplit;
  int row_a, row_b;
  if (kMode == GEMV_QKV_ROPE) {
    const int qk_tasks = p.D;  // 2*D rows of q,k -> D pairs (j, j+64)
    if (task < qk_tasks) {
      const int head = task >> 6, j = task & 63;  // head index over q heads then k heads
      row_a = head * 128 + j;
      row_b = row_a + 64;
    } else {
      row_a = 2 * p.D + 2 * (task - qk_tasks);
      row_b = row_a + 1;
    }
  } else if (kMode == GEMV_SWIGLU) {
    const int blk = task >> 7, j = task & 127;
    row_a = blk * 256 + j;
    row_b = row_a + 128;
  } else {
    row_a = 2 * task;
    row_b = row_a + 1;
  }
  const bool active = row_a < p.N;
  const bool has_b = row_b < p.N;
  const uint4* wa = reinterpret_cast<const uint4*>(p.W + static_cast<size_t>(active ? row_a : 0) * p.K);
  const uint4* wb = reinterpret_cast<const uint4*>(p.W + static_cast<size_t>(has_b ? row_b : (active ? row_a : 0)) * p.K);

  // ---- prefetch the first weight batch (8 x 16 B per lane) BEFORE waiting on the producer kernel:
  // with programmatic dependent launch this overlaps the previous kernel's tail and our own prologue.
  uint4 a[kVec], b[kVec];
#pragma unroll
  for (int u = 0; u < kVec; ++u) {
    const int idx = v0 + lane + 32 * u;
    a[u] = b[u] = make_uint4(0, 0, 0, 0);
    if (active && idx < v1) {
      a[u] = ld_stream(wa + idx);
      b[u] = ld_stream(wb + idx);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");          // producer's writes (x, residual) are visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- prologue: stage x (optionally RMS-normalised, rounded to bf16 like the prefill path)
  const uint4* xg = reinterpret_cast<const uint4*>(p.x);
  if (p.norm_w != nullptr) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < nvec; i += kGemvThreads) {
      const uint4 v = xg[i];
      ss += dot8(v, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kGemvWarps; ++w) tot += red[w];
    const float rstd = rsqrtf(tot / static_cast<float>(p.K) + p.eps);
    const uint4* wg = reinterpret_cast<const uint4*>(p.norm_w);
    for (int i = threadIdx.x; i < nvec; i += kGemvThreads) {
      const uint4 v = xg[i], g = wg[i];
      uint4 o;
      o.x = pack_bf16(bf16_lo(g.x) * (bf16_lo(v.x) * rstd), bf16_hi(g.x) * (bf16_hi(v.x) * rstd));
      o.y = pack_bf16(bf16_lo(g.y) * (bf16_lo(v.y) * rstd), bf16_hi(g.y) * (bf16_hi(v.y) * rstd));
      o.z = pack_bf16(bf16_lo(g.z) * (bf16_lo(v.z) * rstd), bf16_hi(g.z) * (bf16_hi(v.z) * rstd));
      o.w = pack_bf16(bf16_lo(g.w) * (bf16_lo(v.w) * rstd), bf16_hi(g.w) * (bf16_hi(v.w) * rstd));
      xs[i] = o;
    }
  } else {
    for (int i = threadIdx.x; i < nvec; i += kGemvThreads) xs[i] = xg[i];
  }
  __syncthreads();

  // ---- software-pipelined main loop: issue batch n+1, then consume batch n
  float sa = 0.f, sb = 0.f;
  for (int i = v0 + lane; active && i < v1; i += 32 * kVec) {
    uint4 na[kVec], nb[kVec];
#pragma unroll
    for (int u = 0; u < kVec; ++u) {
      const int idx = i + 32 * kVec + 32 * u;
      na[u] = nb[u] = make_uint4(0, 0, 0, 0);
      if (idx < v1) {
        na[u] = ld_stream(wa + idx);
        nb[u] = ld_stream(wb + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < kVec; ++u) {
      const int idx = i + 32 * u;
      if (idx < v1) {
        const uint4 xv = xs[idx];
        sa += dot8(a[u], xv);
        sb += dot8(b[u], xv);
      }
    }
#pragma unroll
    for (int u = 0; u < kVec; ++u) {
      a[u] = na[u];
      b[u] = nb[u];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (ksplit > 1) {  // combine the K partitions of this row pair (adjacent warps) through shared memory
    if (lane == 0) {
      part[warp][0] = sa;
      part[warp][1] = sb;
    }
    __syncthreads();
    if (kpart == 0) {
      for (int s2 = 1; s2 < ksplit; ++s2) {
        sa += part[warp + s2][0];
        sb += part[warp + s2][1];
      }
    }
  }
  if (lane != 0 || kpart != 0 || !active) return;

  if (kMode == GEMV_QKV_ROPE) {
    const int pos = *p.pos;
    // element e of the k (which = 0) / v (which = 1) projection -> row (head e / 128, pos), column which * 128 + e % 128
    auto kv_at = [&](int e, int which) {
      return p.kv_row0 + kv_row_off(e >> 7, pos, p.D >> 7) + which * 128 + (e & 127);
    };
    if (row_a < 2 * p.D) {
      const int j = row_a & 63;
      const float c = p.rope_cos[static_cast<size_t>(pos) * 64 + j];
      const float s = p.rope_sin[static_cast<size_t>(pos) * 64 + j];
      // the prefill path rounds the projection to bf16 only after the rotation (fp32 epilogue)
      const float lo = sa * c - sb * s, hi = sb * c + sa * s;
      if (row_a < p.D) {
        p.out[row_a] = __float2bfloat16_rn(lo);
        p.out[row_b] = __float2bfloat16_rn(hi);
      } else {
        *kv_at(row_a - p.D, 0) = __float2bfloat16_rn(lo);
        *kv_at(row_b - p.D, 0) = __float2bfloat16_rn(hi);
      }
    } else {
      *kv_at(row_a - 2 * p.D, 1) = __float2bfloat16_rn(sa);
      *kv_at(row_b - 2 * p.D, 1) = __float2bfloat16_rn(sb);
    }
  } else if (kMode == GEMV_SWIGLU) {
    const int blk = row_a >> 8, j = row_a & 127;
    p.out[blk * 128 + j] = __float2bfloat16_rn(sa / (1.f + __expf(-sa)) * sb);
  } else {
    if (p.residual) {
      sa += __bfloat162float(p.residual[row_a]);
      if (has_b) sb += __bfloat162float(p.residual[row_b]);
    }
    p.out[row_a] = __float2bfloat16_rn(sa);
    if (has_b) p.out[row_b] = __float2bfloat16_rn(sb);
  }
}

// ---------------------------------------------------------------- flash-decode (head_dim 128)
constexpr int kFdThreads = 128;  // 4 warps x 4 key slots x 8 lanes (16 dims each)

__global__ void __launch_bounds__(kFdThreads)
flash_decode_partial_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                            const int* __restrict__ pos_ptr, int H, int n_splits, float scale_log2e,
                            float* __restrict__ part_o, float* __restrict__ part_ml,
                            unsigned int* __restrict__ tickets, __nv_bfloat16* __restrict__ att_out) {
  __shared__ float s_o[16][128];
  __shared__ unsigned int s_last;
  __shared__ float s_m[16], s_l[16];
  const int h = blockIdx.x, split = blockIdx.y;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int T = *pos_ptr + 1;
  const int chunk = (T + n_splits - 1) / n_splits;
  const int k0 = split * chunk;
  const int k1 = min(T, k0 + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp * 4 + (lane >> 3);  // 0..15
  const int sub = lane & 7;                 // dims [sub*16, sub*16+16)

  float qf[16];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + h * 128 + sub * 16);
    const uint4 a = qp[0], b = qp[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      qf[2 * e] = bf16_lo(w[e]) * scale_log2e;
      qf[2 * e + 1] = bf16_hi(w[e]) * scale_log2e;
    }
  }
  float m = -INFINITY, l = 0.f;
  float o[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) o[e] = 0.f;

  // software-pipelined over keys: the loads of key n+1 are in flight while key n is reduced
  auto load_key = [&](int key, uint4 (&r)[4]) {
    const bool live = key < k1;
    const __nv_bfloat16* row = kv + kv_row_off(h, live ? key : k0, H) + sub * 16;
    r[0] = ld_stream(reinterpret_cast<const uint4*>(row));
    r[1] = ld_stream(reinterpret_cast<const uint4*>(row) + 1);
    r[2] = ld_stream(reinterpret_cast<const uint4*>(row + 128));
    r[3] = ld_stream(reinterpret_cast<const uint4*>(row + 128) + 1);
  };
  uint4 cur[4], nxt[4];
  if (k0 < k1) load_key(k0 + slot, cur);
  for (int base = k0; base < k1; base += 16) {
    const int key = base + slot;
    const bool live = key < k1;
    if (base + 16 < k1) load_key(key + 16, nxt);
    const uint32_t kw[8] = {cur[0].x, cur[0].y, cur[0].z, cur[0].w, cur[1].x, cur[1].y, cur[1].z, cur[1].w};
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s = fmaf(qf[2 * e], bf16_lo(kw[e]), s);
      s = fmaf(qf[2 * e + 1], bf16_hi(kw[e]), s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live) {
      const float m_new = fmaxf(m, s);
      const float alpha = fast_exp2(m - m_new);
      const float pr = fast_exp2(s - m_new);
      l = l * alpha + pr;
      const uint32_t vw[8] = {cur[2].x, cur[2].y, cur[2].z, cur[2].w, cur[3].x, cur[3].y, cur[3].z, cur[3].w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o[2 * e] = fmaf(pr, bf16_lo(vw[e]), o[2 * e] * alpha);
        o[2 * e + 1] = fmaf(pr, bf16_hi(vw[e]), o[2 * e + 1] * alpha);
      }
      m = m_new;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) cur[e] = nxt[e];
  }
  // merge the 16 key slots of the block
#pragma unroll
  for (int e = 0; e < 16; ++e) s_o[slot][sub * 16 + e] = o[e];
  if (sub == 0) {
    s_m[slot] = m;
    s_l[slot] = l;
  }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int g = 0; g < 16; ++g) M = fmaxf(M, s_m[g]);
  const int d = threadIdx.x;  // 128 threads = 128 dims
  float acc = 0.f, L = 0.f;
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const float w = (s_m[g] == -INFINITY) ? 0.f : fast_exp2(s_m[g] - M);
    acc += s_o[g][d] * w;
    L += s_l[g] * w;
  }
  const size_t idx = static_cast<size_t>(h) * n_splits + split;
  part_o[idx * 128 + d] = acc;
  if (d == 0) {
    part_ml[idx * 2] = M;
    part_ml[idx * 2 + 1] = L;
  }
  // The last split of this head to finish merges all partials (no separate combine kernel).
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[h], 1u);
    s_last = (t == static_cast<unsigned int>(n_splits) - 1u) ? 1u : 0u;
    if (s_last) tickets[h] = 0;  // self-resetting for the next layer / token
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float Mx = -INFINITY;
  for (int sp = 0; sp < n_splits; ++sp) Mx = fmaxf(Mx, __ldcg(&part_ml[(static_cast<size_t>(h) * n_splits + sp) * 2]));
  float o_acc = 0.f, l_acc = 0.f;
  for (int sp = 0; sp < n_splits; ++sp) {
    const size_t i2 = static_cast<size_t>(h) * n_splits + sp;
    const float m = __ldcg(&part_ml[i2 * 2]);
    const float w = (m == -INFINITY) ? 0.f : fast_exp2(m - Mx);
    o_acc += __ldcg(&part_o[i2 * 128 + d]) * w;
    l_acc += __ldcg(&part_ml[i2 * 2 + 1]) * w;
  }
  att_out[h * 128 + d] = __float2bfloat16_rn(o_acc / l_acc);
}

__global__ void __launch_bounds__(128)
flash_decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                            int n_splits, __nv_bfloat16* __restrict__ out) {
  const int h = blockIdx.x, d = threadIdx.x;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  float M = -INFINITY;
  for (int s = 0; s < n_splits; ++s) M = fmaxf(M, part_ml[(static_cast<size_t>(h) * n_splits + s) * 2]);
  float acc = 0.f, L = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const size_t idx = static_cast<size_t>(h) * n_splits + s;
    const float m = part_ml[idx * 2];
    const float w = (m == -INFINITY) ? 0.f : fast_exp2(m - M);
    acc += part_o[idx * 128 + d] * w;
    L += part_ml[idx * 2 + 1] * w;
  }
  out[h * 128 + d] = __float2bfloat16_rn(acc / L);
}

// packed qkv rows -> head-major KV cache: one 16-byte chunk per thread; chunk c of token (b, s):
// c in [0, 2*D/8): which = c / (D/8) (k or v), head = (c % (D/8)) / 16, part = c % 16
__global__ void kv_append_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ldq, __nv_bfloat16* __restrict__ kv,
                                 int64_t cap, int B, int S, int D, int H, int pos0) {
  const int cpr = D >> 2;  // 16-byte chunks of [k | v] per token
  const int64_t total = static_cast<int64_t>(B) * S * cpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpr);
    const int64_t tok = i / cpr;
    const int b = static_cast<int>(tok / S), s = static_cast<int>(tok % S);
    const int which = c / (D >> 3), hc = c % (D >> 3);
    const int h = hc >> 4, part = hc & 15;
    const uint4 v = *reinterpret_cast<const uint4*>(qkv + tok * ldq + D + static_cast<size_t>(which) * D + h * 128 + part * 8);
    *reinterpret_cast<uint4*>(kv + static_cast<size_t>(b) * cap * H * 256 + kv_row_off(h, pos0 + s, H) + which * 128 + part * 8) = v;
  }
}

// token embedding row -> x ; token id read from device memory (graph-replayable greedy loop)
__global__ void embed_token_kernel(const __nv_bfloat16* __restrict__ table, const int64_t* __restrict__ token,
                                   __nv_bfloat16* __restrict__ x, int D, int64_t vocab) {
  int64_t id = *token;
  if (id < 0 || id >= vocab) id = 0;
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * D);
  uint4* dst = reinterpret_cast<uint4*>(x);
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < (D >> 3); i += blockDim.x * gridDim.x)
    dst[i] = src[i];
}
// out_ids[*n_out] = *token ; ++*n_out ; ++*pos
__global__ void record_advance_kernel(const int64_t* token, int64_t* out_ids, int* n_out, int* pos, int cap) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int n = *n_out;
    if (n < cap) out_ids[n] = *token;
    *n_out = n + 1;
    *pos = *pos + 1;
  }
}

// Launch config with programmatic dependent launch: the kernel may start while its predecessor in the
// stream drains; it calls griddepcontrol.wait before touching anything the predecessor wrote.
static bool g_use_pdl = true;
static thread_local cudaLaunchAttribute g_pdl_attr;
static cudaLaunchConfig_t pdl_config(dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  g_pdl_attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  g_pdl_attr.val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = &g_pdl_attr;
  cfg.numAttrs = 1;
  return cfg;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int kMode>
static int launch_gemv(cudaStream_t st, GemvParams p, int tasks, int ksplit) {
  p.ksplit = ksplit;
  const size_t smem = static_cast<size_t>(p.K) * 2;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] { attr_rc = cudaFuncSetAttribute(gemv_pair_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       96 * 1024); });
  MB_CHECK_CUDA(attr_rc);
  MB_CHECK_ARG(smem <= 96 * 1024 && p.K % 8 == 0, "gemv: K=%d unsupported", p.K);
  const int blocks = (tasks * p.ksplit + kGemvWarps - 1) / kGemvWarps;
  cudaLaunchConfig_t cfg = pdl_config(dim3(blocks), dim3(kGemvThreads), smem, st);
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemv_pair_kernel<kMode>, p));
  return MERLIN_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

size_t merlin_decode_workspace_bytes(int hidden, int intermediate, int heads, int n_splits) {
  size_t b = 0;
  b += static_cast<size_t>(hidden) * 2 * 3;                         // q, att, xn/spare
  b += static_cast<size_t>(intermediate) * 2;                       // mlp mid
  b += static_cast<size_t>(heads) * n_splits * (128 + 2) * 4;       // flash-decode partials
  b += static_cast<size_t>(heads) * 4;                              // per-head arrival tickets
  return (b + 255) / 256 * 256 + 1024;
}

int merlin_decode_step(void* stream, const merlin_decoder_weights* w, void* x, const int* pos_dev,
                       void* workspace, void* logits, int n_splits) {
  MB_CHECK_ARG(w && x && pos_dev && workspace && logits, "decode_step: null argument");
  MB_CHECK_ARG(w->hidden == w->heads * 128, "decode_step: head_dim must be 128");
  MB_CHECK_ARG(w->intermediate % 128 == 0 && n_splits >= 1 && n_splits <= 64, "decode_step: bad config");
  MB_CHECK_ARG(w->kv_capacity >= 32 && w->kv_capacity % 32 == 0, "decode_step: kv_capacity must be a multiple of 32");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    const char* e = getenv("MERLIN_PDL");
    g_use_pdl = !(e && e[0] == '0');
  }
  const int D = w->hidden, FF = w->intermediate, H = w->heads;
  // warps per row pair: more CTAs in flight for the small-N projections (tunable for experiments)
  static const int ks_qkv = env_int("MERLIN_KS_QKV", 1), ks_o = env_int("MERLIN_KS_O", 1),
                   ks_gu = env_int("MERLIN_KS_GU", 1), ks_down = env_int("MERLIN_KS_DOWN", 1);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* qbuf = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* att = qbuf + D;
  __nv_bfloat16* xn = att + D;
  __nv_bfloat16* mid = xn + D;
  float* part_o = reinterpret_cast<float*>(ws + ((static_cast<size_t>(3 * D + FF) * 2 + 255) / 256 * 256));
  float* part_ml = part_o + static_cast<size_t>(H) * n_splits * 128;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(part_ml + static_cast<size_t>(H) * n_splits * 2);
  // `tickets` (per-head arrival counters of the flash-decode merge) must be zero on first use — the caller
  // zero-fills the workspace once; the last arriver of every head resets its counter for the next layer / token.
  __nv_bfloat16* xb = static_cast<__nv_bfloat16*>(x);
  const float scale_log2e = 1.4426950408889634f / sqrtf(128.f);
  int rc;
  for (int i = 0; i < w->num_layers; ++i) {
    GemvParams p = {};
    // RMSNorm -> fused QKV GEMV -> RoPE -> q buffer / KV-cache append
    p.W = static_cast<const __nv_bfloat16*>(w->qkv_w[i]);
    p.x = xb;
    p.norm_w = static_cast<const __nv_bfloat16*>(w->ln1_w[i]);
    p.out = qbuf;
    p.kv_row0 = static_cast<__nv_bfloat16*>(w->kv_cache[i]);
    p.rope_cos = w->rope_cos;
    p.rope_sin = w->rope_sin;
    p.pos = pos_dev;
    p.N = 3 * D;
    p.K = D;
    p.D = D;
    p.eps = w->rms_eps;
    if ((rc = launch_gemv<GEMV_QKV_ROPE>(st, p, 3 * D / 2, ks_qkv)) != MERLIN_OK) return rc;
    // split-K attention over the cache (keys 0..pos)
    {
      cudaLaunchConfig_t cfg = pdl_config(dim3(H, n_splits), dim3(kFdThreads), 0, st);
      MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, flash_decode_partial_kernel, static_cast<const __nv_bfloat16*>(qbuf),
                                       static_cast<const __nv_bfloat16*>(w->kv_cache[i]), pos_dev, H, n_splits,
                                       scale_log2e, part_o, part_ml, tickets, att));
    }
    // x += o_proj(att)
    GemvParams po = {};
    po.W = static_cast<const __nv_bfloat16*>(w->o_w[i]);
    po.x = att;
    po.residual = xb;
    po.out = xb;
    po.N = D;
    po.K = D;
    po.D = D;
    if ((rc = launch_gemv<GEMV_PLAIN>(st, po, D / 2, ks_o)) != MERLIN_OK) return rc;
    // RMSNorm -> gate/up GEMV -> SwiGLU
    GemvParams pg = {};
    pg.W = static_cast<const __nv_bfloat16*>(w->gate_up_w[i]);
    pg.x = xb;
    pg.norm_w = static_cast<const __nv_bfloat16*>(w->ln2_w[i]);
    pg.out = mid;
    pg.N = 2 * FF;
    pg.K = D;
    pg.D = D;
    pg.eps = w->rms_eps;
    if ((rc = launch_gemv<GEMV_SWIGLU>(st, pg, FF, ks_gu)) != MERLIN_OK) return rc;
    // x += down(mid)
    GemvParams pd = {};
    pd.W = static_cast<const __nv_bfloat16*>(w->down_w[i]);
    pd.x = mid;
    pd.residual = xb;
    pd.out = xb;
    pd.N = D;
    pd.K = FF;
    pd.D = D;
    if ((rc = launch_gemv<GEMV_PLAIN>(st, pd, D / 2, ks_down)) != MERLIN_OK) return rc;
  }
  // final RMSNorm -> lm_head
  GemvParams pl = {};
  pl.W = static_cast<const __nv_bfloat16*>(w->lm_head_w);
  pl.x = xb;
  pl.norm_w = static_cast<const __nv_bfloat16*>(w->norm_w);
  pl.out = static_cast<__nv_bfloat16*>(logits);
  pl.N = w->vocab;
  pl.K = D;
  pl.D = D;
  pl.eps = w->rms_eps;
  if ((rc = launch_gemv<GEMV_PLAIN>(st, pl, (w->vocab + 1) / 2, 1)) != MERLIN_OK) return rc;
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_kv_append(void* stream, const void* qkv, int64_t ldq, void* kv_cache, int64_t kv_capacity, int B, int S,
                     int D, int H, int pos0) {
  MB_CHECK_ARG(qkv && kv_cache && D == H * 128 && ldq >= 3 * static_cast<int64_t>(D) && ldq % 8 == 0,
               "kv_append: bad argument");
  MB_CHECK_ARG(kv_capacity % 32 == 0, "kv_append: kv_capacity must be a multiple of 32");
  MB_CHECK_ARG(B >= 1 && S >= 0 && pos0 >= 0 && static_cast<int64_t>(pos0) + S <= kv_capacity,
               "kv_append: %d tokens at %d do not fit capacity %lld", S, pos0, static_cast<long long>(kv_capacity));
  if (S == 0) return MERLIN_OK;
  const int64_t total = static_cast<int64_t>(B) * S * (D >> 2);
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count()) * 16));
  kv_append_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), ldq, static_cast<__nv_bfloat16*>(kv_cache), kv_capacity, B, S, D, H, pos0);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_embed_token(void* stream, const void* embed_table, const int64_t* token_dev, void* x, int D,
                       int64_t vocab) {
  MB_CHECK_ARG(embed_table && token_dev && x && D % 8 == 0, "embed_token: bad argument");
  embed_token_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(embed_table), token_dev, static_cast<__nv_bfloat16*>(x), D, vocab);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

int merlin_record_advance(void* stream, const int64_t* token_dev, int64_t* out_ids, int* n_out_dev,
                          int* pos_dev, int capacity) {
  MB_CHECK_ARG(token_dev && out_ids && n_out_dev && pos_dev, "record_advance: null argument");
  record_advance_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(token_dev, out_ids, n_out_dev,
                                                                         pos_dev, capacity);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

}  // extern "C"
