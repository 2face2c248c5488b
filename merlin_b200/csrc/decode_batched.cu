// merlin_b200 — batched decode step for beam search (generate(num_beams = k) of
// mmgpt/engine/eval/eval_box.py:246-254 over HF `_beam_search`).
//
// The reference runs k beams as a batch of k sequences through the HF decoder, each with its own copy of the
// whole KV cache (`repeat_interleave(5)` of the prompt, base_mmgpt.py:162-163; HF's cache re-ordering copies every
// cache tensor every step).  Here one decode position of all k beams is ONE pass over the weights:
//  * every projection is a [N, K] x [K, 8] product on the tensor cores (mma.sync m16n8k16, bf16 in, fp32
//    accumulate): the 8 columns are the beams (unused ones are zero), a CTA owns 16 weight rows and its 8 warps
//    split K, so each weight byte is read from HBM once per step whatever the beam count — the step stays bound by
//    the weight bytes of a single sequence.  (tcgen05 has no use here: M = 128 rows x N = 8 would leave the
//    operand staging, not the math, as the cost; the kernel is a pure HBM stream.)
//  * the KV cache is split into the SHARED prefix (the prompt: the prefill's cache, read once for all beams) and a
//    per-beam SUFFIX (the tokens generated so far), so a beam re-ordering moves only the small suffix blocks;
//  * attention is one kernel: the prefix as two tensor-core products per 16 keys for all beams at once (Q K^T with the
//    beams as MMA rows, P V with V transposed in registers), one short CTA per (head, beam) for the suffix, and the
//    last CTA of a head merges all partials in index order.
// Same arithmetic as decode.cu (fp32 accumulation, RMSNorm rounded to bf16 as the prefill does, RoPE in fp32 before the
// bf16 rounding), so a beam's logits equal the single-sequence path's up to summation order.
#include "host.cuh"

#include <mutex>
#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mb {
namespace {

constexpr int kBWarps = 8;               // warps of a GEMV CTA: they split K
constexpr int kBThreads = kBWarps * 32;
constexpr int kMaxBeams = 8;             // the N of m16n8k16
constexpr int kMaxSplits = 16;          // prefix key splits of the tensor-core attention kernel (registers of its merge)
constexpr int kAThreads = 128;           // attention CTAs: 4 warps x 4 key slots x 8 lanes (16 dims each)

enum BMode : int { B_PLAIN = 0, B_QKV_ROPE = 1, B_SWIGLU = 2 };

struct BGemvParams {
  const __nv_bfloat16* W;         // [N, K]
  const __nv_bfloat16* x;         // [nb, ldx] input activations
  int64_t ldx;
  const __nv_bfloat16* residual;  // [nb, ldo] or nullptr (PLAIN)
  __nv_bfloat16* out;             // PLAIN / SWIGLU: [nb, ldo]; QKV: q buffer [nb, D]
  int64_t ldo;
  __nv_bfloat16* suffix_kv;       // QKV: [nb, suffix_cap/32, H, 32, 256] of this layer
  int64_t suffix_stride;          // elements between two beams' suffix caches
  const float* rope_cos;
  const float* rope_sin;
  const int* pos;                 // device {prefix_len, t}: the new token is suffix row t, position prefix_len + t
  int nb, N, K, D;
};

// 32 bytes per lane in one instruction (LDG.E.256): a warp-level load covers whole 128-byte lines
struct U8 {
  uint32_t r[8];
};
__device__ __forceinline__ U8 ldg_stream256(const void* p) {
  U8 v;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ U8 ld256(const void* p) {
  U8 v;
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ float sumsq8(const uint4& v) {
  float s = bf16_lo(v.x) * bf16_lo(v.x);
  s = fmaf(bf16_hi(v.x), bf16_hi(v.x), s);
  s = fmaf(bf16_lo(v.y), bf16_lo(v.y), s);
  s = fmaf(bf16_hi(v.y), bf16_hi(v.y), s);
  s = fmaf(bf16_lo(v.z), bf16_lo(v.z), s);
  s = fmaf(bf16_hi(v.z), bf16_hi(v.z), s);
  s = fmaf(bf16_lo(v.w), bf16_lo(v.w), s);
  s = fmaf(bf16_hi(v.w), bf16_hi(v.w), s);
  return s;
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ size_t kv_off(int h, int t, int H) {   // decode.cu: kv_row_off
  return ((static_cast<size_t>(t >> 5) * H + h) * 32 + (t & 31)) * 256;
}

// xn[b] = bf16(weight * (x[b] * rstd_b)) — the rounding of the prefill's RMSNorm kernel.  One CTA per beam; the
// projections that follow read xn through L1 instead of normalising a private copy in every CTA.
__global__ void __launch_bounds__(256)
norm_beams_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ weight,
                  __nv_bfloat16* __restrict__ xn, int K, float eps) {
  __shared__ float red[8];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int nvec = K >> 3;
  const uint4* xg = reinterpret_cast<const uint4*>(x + static_cast<size_t>(blockIdx.x) * ldx);
  const uint4* wg = reinterpret_cast<const uint4*>(weight);
  uint4* dst = reinterpret_cast<uint4*>(xn + static_cast<size_t>(blockIdx.x) * K);
  uint4 v[3];                                     // K <= 6144: up to 3 vectors per thread stay in registers
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = threadIdx.x + 256 * u;
    v[u] = make_uint4(0, 0, 0, 0);
    if (i < nvec) v[u] = xg[i];
    ss += sumsq8(v[u]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / static_cast<float>(K) + eps);
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < nvec) {
      const uint4 w = wg[i];
      uint4 o;
      o.x = pack_bf16(bf16_lo(w.x) * (bf16_lo(v[u].x) * rstd), bf16_hi(w.x) * (bf16_hi(v[u].x) * rstd));
      o.y = pack_bf16(bf16_lo(w.y) * (bf16_lo(v[u].y) * rstd), bf16_hi(w.y) * (bf16_hi(v[u].y) * rstd));
      o.z = pack_bf16(bf16_lo(w.z) * (bf16_lo(v[u].z) * rstd), bf16_hi(w.z) * (bf16_hi(v[u].z) * rstd));
      o.w = pack_bf16(bf16_lo(w.w) * (bf16_lo(v[u].w) * rstd), bf16_hi(w.w) * (bf16_hi(v[u].w) * rstd));
      dst[i] = o;
    }
  }
}

template <int kMode>
__device__ __forceinline__ void pair_rows(const BGemvParams& p, int pair, int& row_a, int& row_b) {
  if (kMode == B_QKV_ROPE) {
    if (pair < p.D) {                       // q and k blocks: RoPE pair (j, j + 64) of one head
      row_a = (pair >> 6) * 128 + (pair & 63);
      row_b = row_a + 64;
    } else {
      row_a = 2 * p.D + 2 * (pair - p.D);
      row_b = row_a + 1;
    }
  } else if (kMode == B_SWIGLU) {
    row_a = (pair >> 7) * 256 + (pair & 127);
    row_b = row_a + 128;
  } else {
    row_a = 2 * pair;
    row_b = row_a + 1;
  }
}

// lane (g, t) holds rows (row_a, row_b) of pair g for beams 2t and 2t + 1 in c[0..1] / c[2..3]
template <int kMode>
__device__ __forceinline__ void beams_epilogue(const BGemvParams& p, const float (&c)[4], int row_a, int row_b, bool has_b,
                                               int t) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int beam = 2 * t + half;
    if (beam >= p.nb) continue;
    float va = c[half], vb = c[2 + half];
    if (kMode == B_QKV_ROPE) {
      const int tok = p.pos[1];
      const int position = p.pos[0] + tok;
      const int H = p.D >> 7;
      __nv_bfloat16* kvb = p.suffix_kv + static_cast<size_t>(beam) * p.suffix_stride;
      auto kv_at = [&](int e, int which) { return kvb + kv_off(e >> 7, tok, H) + which * 128 + (e & 127); };
      if (row_a < 2 * p.D) {
        const int j = row_a & 63;
        const float cs = p.rope_cos[static_cast<size_t>(position) * 64 + j];
        const float sn = p.rope_sin[static_cast<size_t>(position) * 64 + j];
        const float lo = va * cs - vb * sn, hi = vb * cs + va * sn;
        if (row_a < p.D) {
          p.out[static_cast<size_t>(beam) * p.ldo + row_a] = __float2bfloat16_rn(lo);
          p.out[static_cast<size_t>(beam) * p.ldo + row_b] = __float2bfloat16_rn(hi);
        } else {
          *kv_at(row_a - p.D, 0) = __float2bfloat16_rn(lo);
          *kv_at(row_b - p.D, 0) = __float2bfloat16_rn(hi);
        }
      } else {
        *kv_at(row_a - 2 * p.D, 1) = __float2bfloat16_rn(va);
        *kv_at(row_b - 2 * p.D, 1) = __float2bfloat16_rn(vb);
      }
    } else if (kMode == B_SWIGLU) {
      p.out[static_cast<size_t>(beam) * p.ldo + (row_a >> 8) * 128 + (row_a & 127)] =
          __float2bfloat16_rn(va / (1.f + __expf(-va)) * vb);
    } else {
      if (p.residual) {
        va += __bfloat162float(p.residual[static_cast<size_t>(beam) * p.ldo + row_a]);
        if (has_b) vb += __bfloat162float(p.residual[static_cast<size_t>(beam) * p.ldo + row_b]);
      }
      p.out[static_cast<size_t>(beam) * p.ldo + row_a] = __float2bfloat16_rn(va);
      if (has_b) p.out[static_cast<size_t>(beam) * p.ldo + row_b] = __float2bfloat16_rn(vb);
    }
  }
}

// out[b, r] = dot(W[r], x_b) for all beams at once.  A CTA owns 8 row PAIRS (row_a -> MMA rows 0-7, row_b -> MMA rows
// 8-15; a pair is what one epilogue needs: the two halves of a RoPE rotation, gate and up of one SwiGLU output, or two
// adjacent rows) and its 8 warps split K in blocks of 64 elements.  Lane (g, t) = (lane / 4, lane % 4) loads the 32
// bytes at offset 32 t of row_a(g), row_b(g) (one 256-bit streaming load each: a warp-level load covers eight whole
// 128-byte lines) and of beam g's activation (through L1: every CTA of an SM reads the same few KB): both operands see
// the same K permutation inside the block, so four MMAs per block consume them without any shuffle.  kDepth blocks per warp are in flight in a register ring (a slot is reloaded as soon as it is consumed):
// 2 for the wide projections (four CTAs per SM), 4 for o_proj / down_proj whose 256 CTAs do not fill the chip otherwise.
template <int kMode, int kDepth>
__global__ void __launch_bounds__(kBThreads, kDepth <= 2 ? 4 : 2)
gemv_beams256_kernel(const BGemvParams p) {
  __shared__ float red[kBWarps][32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  int row_a, row_b;
  pair_rows<kMode>(p, blockIdx.x * 8 + g, row_a, row_b);
  const bool active = row_a < p.N;
  const bool has_b = row_b < p.N;
  const int nblk = p.K >> 6;
  const uint8_t* wa = reinterpret_cast<const uint8_t*>(p.W + static_cast<size_t>(active ? row_a : 0) * p.K) + t * 32;
  const uint8_t* wb = reinterpret_cast<const uint8_t*>(p.W + static_cast<size_t>(has_b ? row_b : (active ? row_a : 0)) * p.K) + t * 32;
  U8 a[kDepth], b[kDepth];
#pragma unroll
  for (int d = 0; d < kDepth; ++d) {
    const int blk = warp + kBWarps * d;
    a[d] = U8{};
    b[d] = U8{};
    if (blk < nblk) {
      a[d] = ldg_stream256(wa + blk * 128);
      b[d] = ldg_stream256(wb + blk * 128);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool beam_live = g < p.nb;
  const uint8_t* xrow = reinterpret_cast<const uint8_t*>(p.x + static_cast<size_t>(beam_live ? g : 0) * p.ldx) + t * 32;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  for (int base = warp; base < nblk; base += kBWarps * kDepth) {
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      const int blk = base + kBWarps * d;
      if (blk < nblk) {                         // warp-uniform
        U8 xv = {};
        if (beam_live) xv = ld256(xrow + blk * 128);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_16816(c, a[d].r[2 * j], b[d].r[2 * j], a[d].r[2 * j + 1], b[d].r[2 * j + 1], xv.r[2 * j], xv.r[2 * j + 1]);
        const int nxt = blk + kBWarps * kDepth;
        if (nxt < nblk) {
          a[d] = ldg_stream256(wa + nxt * 128);
          b[d] = ldg_stream256(wb + nxt * 128);
        }
      }
    }
  }
  red[warp][lane][0] = c[0];
  red[warp][lane][1] = c[1];
  red[warp][lane][2] = c[2];
  red[warp][lane][3] = c[3];
  __syncthreads();
  if (warp != 0 || !active) return;
#pragma unroll
  for (int w = 1; w < kBWarps; ++w) {
    c[0] += red[w][lane][0];
    c[1] += red[w][lane][1];
    c[2] += red[w][lane][2];
    c[3] += red[w][lane][3];
  }
  beams_epilogue<kMode>(p, c, row_a, row_b, has_b, t);
}

// ---- attention of all beams for one head.  Grid (H, NB + n_splits): CTA y < NB attends beam y over its own suffix
// rows 0 .. t (the new token included; per-beam keys: a plain FMA pass over a few keys), CTA y >= NB walks key range
// y - NB of the SHARED prefix once for ALL beams on the tensor cores — a warp takes 16 keys per step with 32 MMAs (an
// FMA version of this loop spent ~80 instructions per (key, beam) and was instruction-bound from 3 beams on: 42 us per
// layer at 5 beams against 10 us of HBM time):
//   S[beam, key] = Q K^T : A = Q (rows 0-7 = beams, rows 8-15 zero), B = K rows as they lie in the cache (a key's
//     dims are contiguous = the "col" operand layout); two 8-key tiles x eight 16-dim steps.  Lane (g, t) loads the
//     64 bytes at offset 64 t of q_g and of key g: both operands see the same dim permutation, no shuffles.
//   online softmax per beam on the C fragments (a lane owns one beam: rows are beams), P rounded to bf16 like the
//     prefill kernel's; the S fragments of the two key tiles ARE the A fragment of P for the next product (FA-2 reuse);
//   O[beam, dim] += P V : B = V^T needs two KEYS of one dim per register, the cache has two DIMS of one key:
//     movmatrix.trans on the 8x8 blocks of the loaded V registers transposes them in place (32 per step).
// Lane (g, t) ends up with O[beam g][dims 32 t .. 32 t + 31].  The four warps of a CTA take alternate 16-key groups
// and are merged through shared memory.  Partials: part_o[((h * NP + split) * NB + b) * 128 + d], part_ml[(...) * 2 +
// {m, l}], NP = n_splits + 1 (the suffix is partial n_splits); the LAST CTA of a head to finish (per-head arrival
// ticket, self-resetting) merges them in index order — deterministic — into att[b, h * 128 + d].
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

__global__ void __launch_bounds__(kAThreads, 2)
beam_attention_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ prefix_kv,
                          const __nv_bfloat16* __restrict__ suffix_kv, int64_t suffix_stride, const int* __restrict__ pos,
                          int H, int n_splits, int NB, float scale_log2e, float* __restrict__ part_o,
                          float* __restrict__ part_ml, unsigned int* __restrict__ tickets, __nv_bfloat16* __restrict__ att,
                          int64_t ld_att) {
  __shared__ float s_o[4][kMaxBeams][128];     // prefix: [warp][beam][dim]; suffix: [0..1] reused as [16 slots][128]
  __shared__ float s_m[4][kMaxBeams], s_l[4][kMaxBeams];
  __shared__ unsigned int s_last;
  // blockIdx.y < NB: the suffix CTA of beam blockIdx.y (short; scheduled first so that it never waits for a slot
  // behind the prefix CTAs); blockIdx.y >= NB: prefix split blockIdx.y - NB
  const int h = blockIdx.x, NP = n_splits + 1;
  const int split = static_cast<int>(blockIdx.y) >= NB ? static_cast<int>(blockIdx.y) - NB : n_splits + static_cast<int>(blockIdx.y);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = threadIdx.x;
  if (split < n_splits) {
    const int g = lane >> 2, t = lane & 3;
    const int prefix_len = pos[0];
    const int chunk = ((prefix_len + n_splits - 1) / n_splits + 15) / 16 * 16;
    const int k0 = min(prefix_len, split * chunk), k1 = min(prefix_len, k0 + chunk);
    auto row_ptr = [&](int key) {   // lane's 64 bytes of the k half of a cache row (v half: + 128 elements)
      return prefix_kv + kv_off(h, min(key, k1 - 1), H) + t * 32;
    };
    U8 ka[2], kb[2], va[2], vb[2];   // keys base + g (a) and base + 8 + g (b): two 32-byte halves each
    int base = k0 + warp * 16;
    if (base < k1) {                 // the prompt cache is not written by this step: load before the dependency wait
      const __nv_bfloat16* ra = row_ptr(base + g);
      const __nv_bfloat16* rb = row_ptr(base + 8 + g);
      ka[0] = ldg_stream256(ra); ka[1] = ldg_stream256(ra + 16);
      kb[0] = ldg_stream256(rb); kb[1] = ldg_stream256(rb + 16);
      va[0] = ldg_stream256(ra + 128); va[1] = ldg_stream256(ra + 144);
      vb[0] = ldg_stream256(rb + 128); vb[1] = ldg_stream256(rb + 144);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    U8 qa[2] = {};
    if (g < NB) {
      const __nv_bfloat16* qg = q + static_cast<size_t>(g) * ldq + h * 128 + t * 32;
      qa[0] = ld256(qg);
      qa[1] = ld256(qg + 16);
    }
    float o[16][4];
#pragma unroll
    for (int r = 0; r < 16; ++r) o[r][0] = o[r][1] = o[r][2] = o[r][3] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (; base < k1; base += 64) {
      // ---- S = Q K^T for key tiles [base, base + 8) and [base + 8, base + 16)
      float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t q0 = qa[j >> 2].r[(2 * j) & 7], q1 = qa[j >> 2].r[(2 * j + 1) & 7];
        mma_16816(sa, q0, 0u, q1, 0u, ka[j >> 2].r[(2 * j) & 7], ka[j >> 2].r[(2 * j + 1) & 7]);
        mma_16816(sb, q0, 0u, q1, 0u, kb[j >> 2].r[(2 * j) & 7], kb[j >> 2].r[(2 * j + 1) & 7]);
      }
      const int nxt = base + 64;
      const __nv_bfloat16* na = row_ptr(nxt + g);
      const __nv_bfloat16* nb_ = row_ptr(nxt + 8 + g);
      if (nxt < k1) {                // the next group's keys: their latency hides behind the softmax and P V below
        ka[0] = ldg_stream256(na); ka[1] = ldg_stream256(na + 16);
        kb[0] = ldg_stream256(nb_); kb[1] = ldg_stream256(nb_ + 16);
      }
      // lane (g, t): beam g, keys base + 2t, + 2t + 1 (sa[0..1]) and base + 8 + 2t, + 1 (sb[0..1])
      float s0 = sa[0] * scale_log2e, s1 = sa[1] * scale_log2e, s2 = sb[0] * scale_log2e, s3 = sb[1] * scale_log2e;
      if (base + 16 > k1) {          // ragged last group
        if (base + 2 * t >= k1) s0 = -INFINITY;
        if (base + 2 * t + 1 >= k1) s1 = -INFINITY;
        if (base + 8 + 2 * t >= k1) s2 = -INFINITY;
        if (base + 9 + 2 * t >= k1) s3 = -INFINITY;
      }
      float mx = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m, mx);                       // finite: key `base` is always live
      const float alpha = fast_exp2(m - m_new);
      const float p0 = fast_exp2(s0 - m_new), p1 = fast_exp2(s1 - m_new);
      const float p2 = fast_exp2(s2 - m_new), p3 = fast_exp2(s3 - m_new);
      const uint32_t pa = pack_bf16(p0, p1), pb = pack_bf16(p2, p3);
      // the sum runs over the ROUNDED probabilities, the ones the product below uses (as in the prefill kernel)
      l = l * alpha + (bf16_lo(pa) + bf16_hi(pa)) + (bf16_lo(pb) + bf16_hi(pb));
      m = m_new;
      if (alpha != 1.f) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          o[r][0] *= alpha;
          o[r][1] *= alpha;
        }
      }
      // ---- O += P V: B fragments = transposed 8x8 blocks of the V registers
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const uint32_t b0 = movmatrix_trans(va[r >> 3].r[r & 7]);
        const uint32_t b1 = movmatrix_trans(vb[r >> 3].r[r & 7]);
        mma_16816(o[r], pa, 0u, pb, 0u, b0, b1);
      }
      if (nxt < k1) {
        va[0] = ldg_stream256(na + 128); va[1] = ldg_stream256(na + 144);
        vb[0] = ldg_stream256(nb_ + 128); vb[1] = ldg_stream256(nb_ + 144);
      }
    }
    // ---- warp state -> shared memory; l: sum over the four lanes of a beam (they hold different keys)
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      s_o[warp][g][32 * t + 2 * r] = o[r][0];
      s_o[warp][g][32 * t + 2 * r + 1] = o[r][1];
    }
    if (t == 0) {
      s_m[warp][g] = m;
      s_l[warp][g] = l;
    }
    __syncthreads();
    for (int b = 0; b < NB; ++b) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) M = fmaxf(M, s_m[w][b]);
      float acc = 0.f, L = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float wgt = (s_m[w][b] == -INFINITY) ? 0.f : fast_exp2(s_m[w][b] - M);
        acc += s_o[w][b][d] * wgt;
        L += s_l[w][b] * wgt;
      }
      const size_t idx = (static_cast<size_t>(h) * NP + split) * NB + b;
      part_o[idx * 128 + d] = acc;
      if (d == 0) {
        part_ml[idx * 2] = M;
        part_ml[idx * 2 + 1] = L;
      }
    }
  } else {
    // ---- suffix CTA: every beam has its own keys -> per-beam FMA pass (T = generated tokens, small)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float (*so)[128] = reinterpret_cast<float (*)[128]>(&s_o[0][0][0]);   // [16 slots][128]
    float* sm = &s_m[0][0];                                                 // [16]
    float* sl = &s_l[0][0];
    const int slot = warp * 4 + (lane >> 3), sub = lane & 7;
    const int T = pos[1] + 1;
    {
      const int b = split - n_splits;          // one suffix CTA per beam
      const __nv_bfloat16* kv = suffix_kv + static_cast<size_t>(b) * suffix_stride;
      const U8 q8 = ld256(q + static_cast<size_t>(b) * ldq + h * 128 + sub * 16);
      float mm = -INFINITY, ll = 0.f, oo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) oo[e] = 0.f;
      for (int base = 0; base < T; base += 16) {
        const bool live = base + slot < T;
        const __nv_bfloat16* row = kv + kv_off(h, live ? base + slot : 0, H) + sub * 16;
        const U8 k8 = ld256(row), v8 = ld256(row + 128);
        float sc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sc = fmaf(bf16_lo(q8.r[e]), bf16_lo(k8.r[e]), sc);
          sc = fmaf(bf16_hi(q8.r[e]), bf16_hi(k8.r[e]), sc);
        }
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);
        sc += __shfl_xor_sync(0xffffffffu, sc, 2);
        sc += __shfl_xor_sync(0xffffffffu, sc, 4);
        sc *= scale_log2e;
        if (live) {
          const float m_new = fmaxf(mm, sc);
          const float alpha = fast_exp2(mm - m_new);
          const float pr = fast_exp2(sc - m_new);
          ll = ll * alpha + pr;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            oo[2 * e] = fmaf(pr, bf16_lo(v8.r[e]), oo[2 * e] * alpha);
            oo[2 * e + 1] = fmaf(pr, bf16_hi(v8.r[e]), oo[2 * e + 1] * alpha);
          }
          mm = m_new;
        }
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 16; ++e) so[slot][sub * 16 + e] = oo[e];
      if (sub == 0) {
        sm[slot] = mm;
        sl[slot] = ll;
      }
      __syncthreads();
      float M = -INFINITY;
#pragma unroll
      for (int gi = 0; gi < 16; ++gi) M = fmaxf(M, sm[gi]);
      float acc = 0.f, L = 0.f;
#pragma unroll
      for (int gi = 0; gi < 16; ++gi) {
        const float wgt = (sm[gi] == -INFINITY) ? 0.f : fast_exp2(sm[gi] - M);
        acc += so[gi][d] * wgt;
        L += sl[gi] * wgt;
      }
      const size_t idx = (static_cast<size_t>(h) * NP + n_splits) * NB + b;
      part_o[idx * 128 + d] = acc;
      if (d == 0) {
        part_ml[idx * 2] = M;
        part_ml[idx * 2 + 1] = L;
      }
    }
  }
  // ---- the last CTA of this head merges (decode.cu: flash_decode_partial_kernel): n_splits prefix CTAs + NB suffix CTAs
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(&tickets[h], 1u);
    s_last = (tk == static_cast<unsigned int>(n_splits + NB) - 1u) ? 1u : 0u;
    if (s_last) tickets[h] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // warp w merges beams w, w + 4; a lane owns four dims (one 16-byte load per partial), four partials in flight
  for (int b = warp; b < NB; b += 4) {
    float pm[kMaxSplits + 1], pl[kMaxSplits + 1];
    float M = -INFINITY;
#pragma unroll
    for (int sp = 0; sp <= kMaxSplits; ++sp) {
      if (sp < NP) {
        const float2 ml = __ldcg(reinterpret_cast<const float2*>(part_ml + ((static_cast<size_t>(h) * NP + sp) * NB + b) * 2));
        pm[sp] = ml.x;
        pl[sp] = ml.y;
        M = fmaxf(M, ml.x);
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float L = 0.f;
#pragma unroll
    for (int c0 = 0; c0 <= kMaxSplits; c0 += 4) {
      float4 po[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i < NP && c0 + i <= kMaxSplits)
          po[i] = __ldcg(reinterpret_cast<const float4*>(part_o + ((static_cast<size_t>(h) * NP + c0 + i) * NB + b) * 128) + lane);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (c0 + i < NP && c0 + i <= kMaxSplits) {
          const float wgt = (pm[c0 + i] == -INFINITY) ? 0.f : fast_exp2(pm[c0 + i] - M);
          acc.x += po[i].x * wgt;
          acc.y += po[i].y * wgt;
          acc.z += po[i].z * wgt;
          acc.w += po[i].w * wgt;
          L += pl[c0 + i] * wgt;
        }
      }
    }
    const float inv = 1.f / L;
    uint2 out;
    out.x = pack_bf16(acc.x * inv, acc.y * inv);
    out.y = pack_bf16(acc.z * inv, acc.w * inv);
    *reinterpret_cast<uint2*>(att + static_cast<size_t>(b) * ld_att + h * 128 + lane * 4) = out;
  }
}

thread_local cudaLaunchAttribute g_battr;   // per thread: generate() may run on several host threads
cudaLaunchConfig_t bpdl(dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  g_battr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  g_battr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &g_battr;
  cfg.numAttrs = 1;
  return cfg;
}

template <int kMode>
int launch_bgemv(cudaStream_t st, const BGemvParams& p, int pairs) {
  MB_CHECK_ARG(p.K % 64 == 0 && p.ldx % 16 == 0, "decode_step_batched: K=%d unsupported", p.K);
  cudaLaunchConfig_t cfg = bpdl(dim3((pairs + 7) / 8), dim3(kBThreads), 0, st);
  if ((pairs + 7) / 8 <= 2 * sm_count()) {     // a grid that does not fill the chip: more bytes in flight per warp
    MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemv_beams256_kernel<kMode, 4>, p));
  } else {
    MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemv_beams256_kernel<kMode, 2>, p));
  }
  return MERLIN_OK;
}

int launch_norm(cudaStream_t st, const __nv_bfloat16* x, int64_t ldx, const void* weight, __nv_bfloat16* xn, int K,
                float eps, int nb) {
  MB_CHECK_ARG(K % 8 == 0 && K <= 6144, "decode_step_batched: hidden size %d unsupported", K);
  cudaLaunchConfig_t cfg = bpdl(dim3(nb), dim3(256), 0, st);
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, norm_beams_kernel, x, ldx, static_cast<const __nv_bfloat16*>(weight), xn, K, eps));
  return MERLIN_OK;
}

struct AttnArgs {
  const __nv_bfloat16 *q, *prefix_kv, *suffix_kv;
  int64_t ldq, suffix_stride, ld_att;
  const int* pos;
  int H, n_splits;
  float scale;
  float *part_o, *part_ml;
  unsigned int* tickets;
  __nv_bfloat16* att;
};

int launch_attention(cudaStream_t st, const AttnArgs& a, int NB) {
  MB_CHECK_ARG(a.n_splits <= kMaxSplits, "decode_step_batched: n_splits %d > %d", a.n_splits, kMaxSplits);
  cudaLaunchConfig_t cfg = bpdl(dim3(a.H, NB + a.n_splits), dim3(kAThreads), 0, st);
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, beam_attention_kernel, a.q, a.ldq, a.prefix_kv, a.suffix_kv, a.suffix_stride,
                                   a.pos, a.H, a.n_splits, NB, a.scale, a.part_o, a.part_ml, a.tickets, a.att, a.ld_att));
  return MERLIN_OK;
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" {

size_t merlin_decode_batched_workspace_bytes(int hidden, int intermediate, int heads, int n_splits, int num_beams) {
  size_t b = (static_cast<size_t>(num_beams) * (3 * hidden + intermediate) * 2 + 255) / 256 * 256;   // q, att, xn, mid
  b += static_cast<size_t>(heads) * (n_splits + 1) * num_beams * (128 + 2) * 4;                      // partials
  b += static_cast<size_t>(heads) * 4;                                                               // tickets
  return b + 1024;
}

int merlin_decode_step_batched(void* stream, const merlin_decoder_weights* w, const merlin_beam_state* bs, void* x,
                               void* workspace, void* logits, int64_t ld_logits, int n_splits) {
  MB_CHECK_ARG(w && bs && x && workspace && logits, "decode_step_batched: null argument");
  MB_CHECK_ARG(w->hidden == w->heads * 128 && w->intermediate % 128 == 0 && w->hidden % 128 == 0,
               "decode_step_batched: head_dim must be 128");
  MB_CHECK_ARG(bs->num_beams >= 1 && bs->num_beams <= kMaxBeams, "decode_step_batched: %d beams (1..%d)", bs->num_beams,
               kMaxBeams);
  MB_CHECK_ARG(n_splits >= 1 && n_splits <= kMaxSplits, "decode_step_batched: n_splits %d (1..%d)", n_splits, kMaxSplits);
  MB_CHECK_ARG(bs->suffix_capacity >= 32 && bs->suffix_capacity % 32 == 0 && bs->suffix_kv && bs->prefix_kv && bs->pos_dev,
               "decode_step_batched: bad beam state");
  MB_CHECK_ARG(ld_logits >= w->vocab, "decode_step_batched: logits pitch %lld < vocab", static_cast<long long>(ld_logits));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = w->hidden, FF = w->intermediate, H = w->heads, NB = bs->num_beams;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* qbuf = reinterpret_cast<__nv_bfloat16*>(ws);                  // [NB, D]
  __nv_bfloat16* att = qbuf + static_cast<size_t>(NB) * D;                      // [NB, D]
  __nv_bfloat16* xn = att + static_cast<size_t>(NB) * D;                        // [NB, D]
  __nv_bfloat16* mid = xn + static_cast<size_t>(NB) * D;                        // [NB, FF]
  float* part_o = reinterpret_cast<float*>(ws + (static_cast<size_t>(NB) * (3 * D + FF) * 2 + 255) / 256 * 256);
  float* part_ml = part_o + static_cast<size_t>(H) * (n_splits + 1) * NB * 128;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(part_ml + static_cast<size_t>(H) * (n_splits + 1) * NB * 2);
  __nv_bfloat16* xb = static_cast<__nv_bfloat16*>(x);
  const int64_t suffix_stride = bs->suffix_capacity * H * 256;
  AttnArgs aa = {};
  aa.q = qbuf;
  aa.ldq = D;
  aa.suffix_stride = suffix_stride;
  aa.ld_att = D;
  aa.pos = bs->pos_dev;
  aa.H = H;
  aa.n_splits = n_splits;
  aa.scale = 1.4426950408889634f / sqrtf(128.f);
  aa.part_o = part_o;
  aa.part_ml = part_ml;
  aa.tickets = tickets;
  aa.att = att;
  int rc;
  for (int i = 0; i < w->num_layers; ++i) {
    if ((rc = launch_norm(st, xb, D, w->ln1_w[i], xn, D, w->rms_eps, NB)) != MERLIN_OK) return rc;
    BGemvParams p = {};
    p.W = static_cast<const __nv_bfloat16*>(w->qkv_w[i]);
    p.x = xn;
    p.ldx = D;
    p.out = qbuf;
    p.ldo = D;
    p.suffix_kv = static_cast<__nv_bfloat16*>(bs->suffix_kv[i]);
    p.suffix_stride = suffix_stride;
    p.rope_cos = w->rope_cos;
    p.rope_sin = w->rope_sin;
    p.pos = bs->pos_dev;
    p.nb = NB;
    p.N = 3 * D;
    p.K = D;
    p.D = D;
    if ((rc = launch_bgemv<B_QKV_ROPE>(st, p, 3 * D / 2)) != MERLIN_OK) return rc;
    aa.prefix_kv = static_cast<const __nv_bfloat16*>(bs->prefix_kv[i]);
    aa.suffix_kv = static_cast<const __nv_bfloat16*>(bs->suffix_kv[i]);
    rc = launch_attention(st, aa, NB);
    if (rc != MERLIN_OK) return rc;
    BGemvParams po = {};
    po.W = static_cast<const __nv_bfloat16*>(w->o_w[i]);
    po.x = att;
    po.ldx = D;
    po.residual = xb;
    po.out = xb;
    po.ldo = D;
    po.nb = NB;
    po.N = D;
    po.K = D;
    po.D = D;
    if ((rc = launch_bgemv<B_PLAIN>(st, po, D / 2)) != MERLIN_OK) return rc;
    if ((rc = launch_norm(st, xb, D, w->ln2_w[i], xn, D, w->rms_eps, NB)) != MERLIN_OK) return rc;
    BGemvParams pg = {};
    pg.W = static_cast<const __nv_bfloat16*>(w->gate_up_w[i]);
    pg.x = xn;
    pg.ldx = D;
    pg.out = mid;
    pg.ldo = FF;
    pg.nb = NB;
    pg.N = 2 * FF;
    pg.K = D;
    pg.D = D;
    if ((rc = launch_bgemv<B_SWIGLU>(st, pg, FF)) != MERLIN_OK) return rc;
    BGemvParams pd = {};
    pd.W = static_cast<const __nv_bfloat16*>(w->down_w[i]);
    pd.x = mid;
    pd.ldx = FF;
    pd.residual = xb;
    pd.out = xb;
    pd.ldo = D;
    pd.nb = NB;
    pd.N = D;
    pd.K = FF;
    pd.D = D;
    if ((rc = launch_bgemv<B_PLAIN>(st, pd, D / 2)) != MERLIN_OK) return rc;
  }
  if ((rc = launch_norm(st, xb, D, w->norm_w, xn, D, w->rms_eps, NB)) != MERLIN_OK) return rc;
  BGemvParams pl = {};
  pl.W = static_cast<const __nv_bfloat16*>(w->lm_head_w);
  pl.x = xn;
  pl.ldx = D;
  pl.out = static_cast<__nv_bfloat16*>(logits);
  pl.ldo = ld_logits;
  pl.nb = NB;
  pl.N = w->vocab;
  pl.K = D;
  pl.D = D;
  return launch_bgemv<B_PLAIN>(st, pl, (w->vocab + 1) / 2);
}

}  // extern "C"
