// merlin_b200 — single-launch streaming decode step (SURVEY.md K12, §3.3 hot loop #4).
//
// The per-kernel decode path (decode.cu) streams every weight once but pays a ramp-up and a tail per
// kernel: 5 kernels x 32 layers per token leave the HBM pipe idle ~25 % of the time.  Here ONE persistent
// kernel (one CTA per SM, cooperative launch) executes the whole token.
//  * Streaming.  In every CTA a producer thread walks the CTA's share of the weight rows and of the KV-cache
//    tiles of ALL layers in execution order and feeds them through a shared-memory ring with TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx): a weight-row pair, a K-chunk of a down_proj row, or 32 keys of one
//    head (16 KB contiguous in the tile-major cache).  Units go round-robin over the CTAs, so at any time the
//    chip reads one dense window of memory.  Weights and old cache rows do not depend on the activations, so the
//    producer runs ahead across phase and layer boundaries, bounded only by the ring (12 x 16 KB per SM, ~29 MB
//    in flight on the chip): HBM keeps streaming while the consumers wait at a grid-wide barrier.
//  * Consuming.  8 consumer warps take one ring stage each (no intra-CTA reductions), keep the (RMS-normalised)
//    activation vector in shared memory and apply the same fused epilogues as decode.cu: RoPE + KV-cache append,
//    SwiGLU, residual add.  Five grid barriers per layer (release arrival + polled counter).
//  * Attention.  Head h owns a group of consecutive ring warps, warp s of the group takes key tiles s, s + WPH,
//    ...; online softmax per warp, the last contributor of a head (ticket counter) merges the partials in slot
//    order — deterministic.  The unaligned tail keys and the new token are handled by one warp per head that
//    reads them straight from global memory.
// DESIGN.md §7 lists the measured variants (fence-free protocols, other cache layouts, L2 prefetch) that lost.
#include "host.cuh"

#include <mutex>
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>

#ifndef MERLIN_MERGE_UNROLL
#define MERLIN_MERGE_UNROLL 12
#endif

namespace mb {

constexpr int kMergeUnroll = MERLIN_MERGE_UNROLL;   // attention partials in flight per round of the last-arriver merge
constexpr int kSC = 8;                            // consumer warps per CTA
constexpr int kSConsThreads = kSC * 32;
constexpr int kSThreads = kSConsThreads + 32;     // + 1 producer warp (warp index kSC)
constexpr int kKT = 32;                           // keys per KV tile
constexpr int kKvTileBytes = kKT * 256 * 2;       // [32 keys][k 128 | v 128] bf16, contiguous in the cache
constexpr int kMaxParts = 96;                     // attention partials per head
constexpr int kMaxLayers = 96;
constexpr int kMaxStages = 16;
#ifdef MERLIN_DEBUG_SPIN
constexpr unsigned int kSpinLimit = 1u << 22;
#else
constexpr unsigned int kSpinLimit = 1u << 24;     // seconds; a healthy wait is microseconds
#endif

enum StreamMode : int { SM_QKV = 0, SM_SWIGLU = 1, SM_PLAIN2 = 2, SM_PLAIN1 = 3 };

struct StreamTables {  // device copy of the per-layer pointer tables (written by merlin_decode_stream_prepare)
  const __nv_bfloat16* qkv_w[kMaxLayers];
  const __nv_bfloat16* o_w[kMaxLayers];
  const __nv_bfloat16* gu_w[kMaxLayers];
  const __nv_bfloat16* down_w[kMaxLayers];
  const __nv_bfloat16* ln1_w[kMaxLayers];
  const __nv_bfloat16* ln2_w[kMaxLayers];
  __nv_bfloat16* kv[kMaxLayers];
};

struct StreamLayout {  // byte offsets inside the caller's workspace
  size_t ctrl, done, tables, xres, qbuf, att, mid, part_ml, part_o, total;
};
static StreamLayout stream_layout(int L, int D, int H, int FF) {
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  StreamLayout s;
  size_t o = 0;
  s.ctrl = o;   o += 512;                                    // grid-barrier counter @0, exit counter @128
  s.done = o;   o += up(static_cast<size_t>(H) * 4);
  s.tables = o; o += up(sizeof(StreamTables));
  s.xres = o;   o += up(static_cast<size_t>(D) * 2);
  s.qbuf = o;   o += up(static_cast<size_t>(D) * 2);
  s.att = o;    o += up(static_cast<size_t>(D) * 2);
  s.mid = o;    o += up(static_cast<size_t>(FF) * 2);
  s.part_ml = o; o += up(static_cast<size_t>(H) * kMaxParts * 2 * 4);
  s.part_o = o;  o += up(static_cast<size_t>(H) * kMaxParts * 128 * 4);
  s.total = o + 256;
  return s;
}

struct StreamParams {
  const StreamTables* tb;
  const __nv_bfloat16* norm_w;
  const __nv_bfloat16* lm_head;
  const __nv_bfloat16* x_in;   // token embedding (residual stream at layer 0)
  __nv_bfloat16* x;            // residual stream (written from layer 0's o_proj on)
  __nv_bfloat16* logits;
  __nv_bfloat16* qbuf;
  __nv_bfloat16* att;
  __nv_bfloat16* mid;
  float* part_o;
  float* part_ml;
  unsigned int* done;
  unsigned int* bar;
  unsigned int* exitc;
  const float* rope_cos;
  const float* rope_sin;
  const int* pos;
  int L, D, H, FF, V;
  float eps;
  int slot_bytes, nst, down_rows, down_split;
};

// ------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictFirst)
      : "memory");
}
__device__ __forceinline__ void ring_wait(uint64_t* bar, uint32_t parity, int what) {
  unsigned int it = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++it > kSpinLimit) {
      printf("decode_stream: ring wait timeout blk=%d thr=%d what=%d parity=%u\n", blockIdx.x, threadIdx.x, what, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float fma8(const uint4& w, const uint4& x, float s) {
  s = fmaf(bf16_lo(w.x), bf16_lo(x.x), s);
  s = fmaf(bf16_hi(w.x), bf16_hi(x.x), s);
  s = fmaf(bf16_lo(w.y), bf16_lo(x.y), s);
  s = fmaf(bf16_hi(w.y), bf16_hi(x.y), s);
  s = fmaf(bf16_lo(w.z), bf16_lo(x.z), s);
  s = fmaf(bf16_hi(w.z), bf16_hi(x.z), s);
  s = fmaf(bf16_lo(w.w), bf16_lo(x.w), s);
  s = fmaf(bf16_hi(w.w), bf16_hi(x.w), s);
  return s;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// All consumer threads of every CTA: nobody leaves before everybody (and their global writes) arrived.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target, int ctid) {
  named_bar_sync(1, kSConsThreads);
  if (ctid == 0) {
    // release: the CTA's epilogue stores (ordered before this thread by the bar.sync above) become visible
    // before the arrival; every cross-CTA read in this kernel is an L2 read, so no L1 invalidation is needed
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    // bounded: a lost arrival must end in a trap (launch failure), not in a hung GPU
    unsigned int it = 0;
    while (ld_acquire_u32(counter) < target) {
      if (++it > kSpinLimit) {
        printf("decode_stream: grid barrier timeout blk=%d target=%u have=%u\n", blockIdx.x, target,
               ld_acquire_u32(counter));
        __trap();
      }
    }
  }
  named_bar_sync(1, kSConsThreads);
}

// xs[0..K) = src (optionally RMS-normalised with norm_w, rounded to bf16 like the prefill path).
__device__ __forceinline__ void stage_vec(const __nv_bfloat16* src, const __nv_bfloat16* norm_w, int K, float eps,
                                          uint4* xs, float* red, int ctid) {
  const int nvec = K >> 3;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  if (norm_w != nullptr) {
    float ss = 0.f;
    for (int i = ctid; i < nvec; i += kSConsThreads) {
      const uint4 v = __ldcg(s4 + i);
      ss = fma8(v, v, ss);
    }
    ss = warp_sum(ss);
    if ((ctid & 31) == 0) red[ctid >> 5] = ss;
    named_bar_sync(2, kSConsThreads);
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kSC; ++w) tot += red[w];
    const float rstd = rsqrtf(tot / static_cast<float>(K) + eps);
    const uint4* g4 = reinterpret_cast<const uint4*>(norm_w);
    for (int i = ctid; i < nvec; i += kSConsThreads) {
      const uint4 v = __ldcg(s4 + i), g = __ldg(g4 + i);
      uint4 o;
      o.x = pack_bf16(bf16_lo(g.x) * (bf16_lo(v.x) * rstd), bf16_hi(g.x) * (bf16_hi(v.x) * rstd));
      o.y = pack_bf16(bf16_lo(g.y) * (bf16_lo(v.y) * rstd), bf16_hi(g.y) * (bf16_hi(v.y) * rstd));
      o.z = pack_bf16(bf16_lo(g.z) * (bf16_lo(v.z) * rstd), bf16_hi(g.z) * (bf16_hi(v.z) * rstd));
      o.w = pack_bf16(bf16_lo(g.w) * (bf16_lo(v.w) * rstd), bf16_hi(g.w) * (bf16_hi(v.w) * rstd));
      xs[i] = o;
    }
  } else {
    for (int i = ctid; i < nvec; i += kSConsThreads) xs[i] = __ldcg(s4 + i);
  }
  named_bar_sync(2, kSConsThreads);
}

// Weight rows of GEMV unit u (second row < 0: single-row unit).
__device__ __forceinline__ void unit_rows(int mode, int u, int D, int N, int& ra, int& rb) {
  if (mode == SM_QKV) {
    if (u < D) {               // q and k blocks: RoPE pair (j, j + 64) of one head
      ra = (u >> 6) * 128 + (u & 63);
      rb = ra + 64;
    } else {                   // v block: adjacent rows
      ra = 2 * D + 2 * (u - D);
      rb = ra + 1;
    }
  } else if (mode == SM_SWIGLU) {  // gate row j and up row j of one 128-row block
    ra = (u >> 7) * 256 + (u & 127);
    rb = ra + 128;
  } else if (mode == SM_PLAIN2) {
    ra = 2 * u;
    rb = (ra + 1 < N) ? ra + 1 : -1;
  } else {
    ra = u;
    rb = -1;
  }
}

struct GemvPhase {
  const __nv_bfloat16* W;
  int mode, U, K, N;
  int spu;   // ring stages per unit: single-row units (SM_PLAIN1) are split into spu K-chunks of K / spu
};

struct Ring {
  uint8_t* base;
  uint64_t* full;
  uint64_t* empty;
  volatile int* released;   // per slot: last stage whose consumer released it (-1 initially)
  int nst, slot_bytes;
};

// ---- producer side (one thread): stage `i` of this CTA goes to slot i % nst
struct ProducerState {
  int slot;
  uint32_t round;
};
__device__ __forceinline__ uint8_t* producer_acquire(const Ring& r, ProducerState& ps, uint32_t bytes) {
  if (ps.round > 0) ring_wait(&r.empty[ps.slot], (ps.round - 1) & 1, 0);
  mbar_expect_tx(&r.full[ps.slot], bytes);
  return r.base + static_cast<size_t>(ps.slot) * r.slot_bytes;
}
__device__ __forceinline__ void producer_advance(const Ring& r, ProducerState& ps) {
  if (++ps.slot == r.nst) {
    ps.slot = 0;
    ++ps.round;
  }
}

__device__ __forceinline__ void produce_gemv(const Ring& r, ProducerState& ps, const GemvPhase& ph, int first, int G,
                                             int D) {
  const int kc = ph.K / ph.spu;                      // elements per stage and row
  const uint32_t row_bytes = static_cast<uint32_t>(kc) * 2;
  for (int u = first; u < ph.U; u += G) {
    int ra, rb;
    unit_rows(ph.mode, u, D, ph.N, ra, rb);
    for (int c = 0; c < ph.spu; ++c) {
      uint8_t* dst = producer_acquire(r, ps, rb >= 0 ? 2 * row_bytes : row_bytes);
      bulk_g2s(dst, ph.W + static_cast<size_t>(ra) * ph.K + c * kc, row_bytes, &r.full[ps.slot]);
      if (rb >= 0) bulk_g2s(dst + row_bytes, ph.W + static_cast<size_t>(rb) * ph.K + c * kc, row_bytes, &r.full[ps.slot]);
      producer_advance(r, ps);
    }
  }
}

// KV cache layout (shared with decode.cu): blocks of 32 tokens, head-major inside a block —
//   element (h, t, c) at (((t / 32) * H + h) * 32 + t % 32) * 256 + c,  c < 128: k (rotated), c >= 128: v
// so one (head, 32-key tile) is 16 KB contiguous AND all heads of a token range form one dense region.
__device__ __forceinline__ size_t kv_tile_off(int h, int kt, int H) {
  return (static_cast<size_t>(kt) * H + h) * (kKT * 256);
}
// Attention work split.  Head h owns WPH consecutive ring warps; warp s of the head takes key tiles
// s, s + WPH, s + 2 WPH, ...  At any time the chip therefore streams the same narrow range of key tiles of ALL
// heads — one dense region of the cache (DRAM-friendly), not one stream per warp.  Warp s publishes partial s
// of its head, the tail warp partial min(WPH, tph): slot indices are a pure function of the split.
struct AttSplit {
  int H, tph, WPH;
};
__device__ __forceinline__ int att_head_of(const AttSplit& a, int gw) { return gw / a.WPH; }   // >= H: no work
__device__ __forceinline__ int att_count(const AttSplit& a, int gw) {
  if (gw >= a.WPH * a.H) return 0;
  const int s = gw % a.WPH;
  return a.tph > s ? (a.tph - s + a.WPH - 1) / a.WPH : 0;
}

// ---- consumer side
// Stages of one slot are consumed by DIFFERENT warps, and a warp only ever looks at its own stages, so a
// parity wait alone is ambiguous: while the slot's previous stage (stage - nst) is still in flight, the
// barrier's last completed phase (stage - 2 nst) has the parity we are about to wait for.  Waiting first for
// the previous stage to be RELEASED (which implies it landed) makes the parity wait exact.
__device__ __forceinline__ void consumer_wait(const Ring& r, uint32_t stage, int& slot, int lane) {
  slot = static_cast<int>(stage % static_cast<uint32_t>(r.nst));
  if (stage >= static_cast<uint32_t>(r.nst)) {
    if (lane == 0) {
      const int want = static_cast<int>(stage) - r.nst;
      unsigned int it = 0;
      while (r.released[slot] != want) {
        if (++it > kSpinLimit) {
          printf("decode_stream: release wait timeout blk=%d thr=%d stage=%u have=%d\n", blockIdx.x, threadIdx.x, stage,
                 r.released[slot]);
          __trap();
        }
      }
    }
    __syncwarp();
  }
  ring_wait(&r.full[slot], (stage / static_cast<uint32_t>(r.nst)) & 1, 1);
}
__device__ __forceinline__ void consumer_release(const Ring& r, int slot, uint32_t stage, int lane) {
  __syncwarp();
  if (lane == 0) {
    r.released[slot] = static_cast<int>(stage);
    mbar_arrive(&r.empty[slot]);
  }
}

struct GemvEpilogue {
  __nv_bfloat16* out;             // QKV: q buffer; SWIGLU: mid; PLAIN: out vector
  const __nv_bfloat16* residual;  // PLAIN (may be null)
  __nv_bfloat16* kv_base;         // QKV: cache of this layer (layout: kv_tile_off)
  int H, pos;
  const float* cos_row;           // QKV: rope tables at this position
  const float* sin_row;
};

__device__ __forceinline__ void consume_gemv(const Ring& r, uint32_t& stage0, const GemvPhase& ph, int first, int G,
                                             int D, const uint4* xs, const GemvEpilogue& ep, int warp, int lane) {
  const int nvec = ph.K >> 3;
  int j = 0;
  for (int u = first; u < ph.U; u += G, ++j) {
    if ((j % kSC) != warp) continue;
    int ra, rb;
    unit_rows(ph.mode, u, D, ph.N, ra, rb);
    // epilogue operands first: their latency hides behind the ring wait and the dot products
    float e0 = 0.f, e1 = 0.f;
    if (ph.mode == SM_QKV) {
      if (u < D) {
        e0 = ep.cos_row[u & 63];
        e1 = ep.sin_row[u & 63];
      }
    } else if (ph.mode != SM_SWIGLU && ep.residual != nullptr) {
      e0 = __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(ep.residual + ra))));
      if (rb >= 0)
        e1 = __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(ep.residual + rb))));
    }
    float sa, sb = 0.f;
    if (rb >= 0) {   // row pair: one stage (spu == 1)
      int slot;
      consumer_wait(r, stage0 + j, slot, lane);
      const uint4* wa = reinterpret_cast<const uint4*>(r.base + static_cast<size_t>(slot) * r.slot_bytes);
      const uint4* wb = wa + nvec;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      int i = lane;
      for (; i + 96 < nvec; i += 128) {
        const uint4 x0 = xs[i], x1 = xs[i + 32], x2 = xs[i + 64], x3 = xs[i + 96];
        const uint4 p0 = wa[i], p1 = wa[i + 32], p2 = wa[i + 64], p3 = wa[i + 96];
        const uint4 q0 = wb[i], q1 = wb[i + 32], q2 = wb[i + 64], q3 = wb[i + 96];
        a0 = fma8(p0, x0, a0);
        b0 = fma8(q0, x0, b0);
        a1 = fma8(p1, x1, a1);
        b1 = fma8(q1, x1, b1);
        a2 = fma8(p2, x2, a2);
        b2 = fma8(q2, x2, b2);
        a3 = fma8(p3, x3, a3);
        b3 = fma8(q3, x3, b3);
      }
      for (; i < nvec; i += 32) {
        const uint4 x0 = xs[i];
        a0 = fma8(wa[i], x0, a0);
        b0 = fma8(wb[i], x0, b0);
      }
      sa = warp_sum((a0 + a1) + (a2 + a3));
      sb = warp_sum((b0 + b1) + (b2 + b3));
      consumer_release(r, slot, stage0 + j, lane);   // every lane's reads fed the shuffles above
    } else {         // single row in spu K-chunks, one stage each, all consumed by this warp
      const int nvc = nvec / ph.spu;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int c = 0; c < ph.spu; ++c) {
        int slot;
        consumer_wait(r, stage0 + j * ph.spu + c, slot, lane);
        const uint4* wa = reinterpret_cast<const uint4*>(r.base + static_cast<size_t>(slot) * r.slot_bytes);
        const uint4* xc = xs + c * nvc;
        int i = lane;
        for (; i + 96 < nvc; i += 128) {
          const uint4 x0 = xc[i], x1 = xc[i + 32], x2 = xc[i + 64], x3 = xc[i + 96];
          const uint4 p0 = wa[i], p1 = wa[i + 32], p2 = wa[i + 64], p3 = wa[i + 96];
          a0 = fma8(p0, x0, a0);
          a1 = fma8(p1, x1, a1);
          a2 = fma8(p2, x2, a2);
          a3 = fma8(p3, x3, a3);
        }
        for (; i < nvc; i += 32) a0 = fma8(wa[i], xc[i], a0);
        // the partial sums carry a dependency on every lane's loads of this stage only through the final
        // reduction, so order the release behind the loads explicitly
        __syncwarp();
        consumer_release(r, slot, stage0 + j * ph.spu + c, lane);
      }
      sa = warp_sum((a0 + a1) + (a2 + a3));
    }
    if (lane != 0) continue;
    if (ph.mode == SM_QKV) {
      // element e of the k / v projection -> row (head e / 128, pos) of the cache, column (v: 128 +) e % 128
      auto kv_at = [&](int e, int which) {
        return ep.kv_base + kv_tile_off(e >> 7, ep.pos >> 5, ep.H) + (ep.pos & 31) * 256 + which * 128 + (e & 127);
      };
      if (u < D) {
        // the prefill path rounds the projection to bf16 only after the rotation (fp32 epilogue)
        const float lo = sa * e0 - sb * e1, hi = sb * e0 + sa * e1;
        if (ra < D) {
          ep.out[ra] = __float2bfloat16_rn(lo);
          ep.out[rb] = __float2bfloat16_rn(hi);
        } else {
          *kv_at(ra - D, 0) = __float2bfloat16_rn(lo);
          *kv_at(rb - D, 0) = __float2bfloat16_rn(hi);
        }
      } else {
        *kv_at(ra - 2 * D, 1) = __float2bfloat16_rn(sa);
        *kv_at(rb - 2 * D, 1) = __float2bfloat16_rn(sb);
      }
    } else if (ph.mode == SM_SWIGLU) {
      ep.out[(ra >> 8) * 128 + (ra & 127)] = __float2bfloat16_rn(sa / (1.f + __expf(-sa)) * sb);
    } else {
      ep.out[ra] = __float2bfloat16_rn(sa + e0);
      if (rb >= 0) ep.out[rb] = __float2bfloat16_rn(sb + e1);
    }
  }
  stage0 += static_cast<uint32_t>(j * ph.spu);
}

// ---- attention: one warp = 4 key slots x 8 lanes; lane `sub` owns dims [8 sub, 8 sub + 8) and [64 + 8 sub, ...)
struct AttState {
  float m, l;
  float o[16];
};
__device__ __forceinline__ void att_reset(AttState& s) {
  s.m = -INFINITY;
  s.l = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) s.o[e] = 0.f;
}
__device__ __forceinline__ void att_load_q(const __nv_bfloat16* qbuf, int h, int sub, float scale, float (&qf)[16]) {
  const uint4* qp = reinterpret_cast<const uint4*>(qbuf + h * 128);
  const uint4 a = __ldcg(qp + sub), b = __ldcg(qp + 8 + sub);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    qf[2 * e] = bf16_lo(w[e]) * scale;
    qf[2 * e + 1] = bf16_hi(w[e]) * scale;
  }
}
// one key for this lane's slot: k0,k1 / v0,v1 are the lane's two 16-byte chunks of the key and value rows
__device__ __forceinline__ void att_key(AttState& s, const float (&qf)[16], const uint4& k0, const uint4& k1,
                                        const uint4& v0, const uint4& v1, bool live) {
  const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  float sc = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc = fmaf(qf[2 * e], bf16_lo(kw[e]), sc);
    sc = fmaf(qf[2 * e + 1], bf16_hi(kw[e]), sc);
  }
  sc += __shfl_xor_sync(0xffffffffu, sc, 1);
  sc += __shfl_xor_sync(0xffffffffu, sc, 2);
  sc += __shfl_xor_sync(0xffffffffu, sc, 4);
  if (live) {
    const float m_new = fmaxf(s.m, sc);
    const float alpha = fast_exp2(s.m - m_new);
    const float pr = fast_exp2(sc - m_new);
    s.l = s.l * alpha + pr;
    const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s.o[2 * e] = fmaf(pr, bf16_lo(vw[e]), s.o[2 * e] * alpha);
      s.o[2 * e + 1] = fmaf(pr, bf16_hi(vw[e]), s.o[2 * e + 1] * alpha);
    }
    s.m = m_new;
  }
}

// Merge the warp's 4 key slots, publish them as partial `idx` of head h (covering `cnt` work items out of
// `total`); the contributor that completes the head merges partials 0..nslots-1 IN INDEX ORDER into
// att[h*128 ..] — slot indices are a pure function of the work split, so the result is deterministic.
__device__ __forceinline__ void att_flush(AttState& s, int h, unsigned int idx, unsigned int nslots, unsigned int cnt,
                                          unsigned int total, const StreamParams& p, int lane) {
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, s.m, off);
    const float l2 = __shfl_xor_sync(0xffffffffu, s.l, off);
    const float M = fmaxf(s.m, m2);
    const float w1 = (s.m == -INFINITY) ? 0.f : fast_exp2(s.m - M);
    const float w2 = (m2 == -INFINITY) ? 0.f : fast_exp2(m2 - M);
    s.l = s.l * w1 + l2 * w2;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float o2 = __shfl_xor_sync(0xffffffffu, s.o[e], off);
      s.o[e] = s.o[e] * w1 + o2 * w2;
    }
    s.m = M;
  }
  const size_t pi = static_cast<size_t>(h) * kMaxParts + idx;
  if (lane < 8) {
    float4* po = reinterpret_cast<float4*>(p.part_o + pi * 128);
    po[lane * 2] = make_float4(s.o[0], s.o[1], s.o[2], s.o[3]);
    po[lane * 2 + 1] = make_float4(s.o[4], s.o[5], s.o[6], s.o[7]);
    po[16 + lane * 2] = make_float4(s.o[8], s.o[9], s.o[10], s.o[11]);
    po[16 + lane * 2 + 1] = make_float4(s.o[12], s.o[13], s.o[14], s.o[15]);
    if (lane == 0) {
      p.part_ml[pi * 2] = s.m;
      p.part_ml[pi * 2 + 1] = s.l;
    }
  }
  __syncwarp();
  unsigned int last = 0;
  if (lane == 0) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(&p.done[h]), "r"(cnt) : "memory");
    last = (old + cnt == total) ? 1u : 0u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  const int n = static_cast<int>(nslots);
  const float* ml = p.part_ml + static_cast<size_t>(h) * kMaxParts * 2;
  const float* po = p.part_o + static_cast<size_t>(h) * kMaxParts * 128;
  float M = -INFINITY;
  for (int q = lane; q < n; q += 32) M = fmaxf(M, __ldcg(ml + 2 * q));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float L = 0.f;
  // 12 partials in flight per round: with 4 the ~37 partials of a head took ten dependent L2 round trips
#pragma unroll kMergeUnroll
  for (int q = 0; q < n; ++q) {
    const float m = __ldcg(ml + 2 * q), l = __ldcg(ml + 2 * q + 1);
    const float w = (m == -INFINITY) ? 0.f : fast_exp2(m - M);
    const float4 o = __ldcg(reinterpret_cast<const float4*>(po + static_cast<size_t>(q) * 128) + lane);
    acc.x = fmaf(o.x, w, acc.x);
    acc.y = fmaf(o.y, w, acc.y);
    acc.z = fmaf(o.z, w, acc.z);
    acc.w = fmaf(o.w, w, acc.w);
    L = fmaf(l, w, L);
  }
  const float inv = 1.f / L;
  uint2 pk;
  pk.x = pack_bf16(acc.x * inv, acc.y * inv);
  pk.y = pack_bf16(acc.z * inv, acc.w * inv);
  reinterpret_cast<uint2*>(p.att + h * 128)[lane] = pk;
  if (lane == 0) p.done[h] = 0;  // self-resetting for the next layer / token (next use is behind a grid barrier)
}

// ------------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kSThreads, 1) decode_stream_kernel(const StreamParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  Ring r;
  r.base = smem;
  r.nst = p.nst;
  r.slot_bytes = p.slot_bytes;
  const int xs_bytes = ((max(p.D, p.FF) * 2) + 127) / 128 * 128;
  uint4* xs = reinterpret_cast<uint4*>(smem + static_cast<size_t>(p.nst) * p.slot_bytes);
  r.full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + xs_bytes);
  r.empty = r.full + kMaxStages;
  r.released = reinterpret_cast<volatile int*>(r.empty + kMaxStages);
  float* red = reinterpret_cast<float*>(const_cast<int*>(r.released) + kMaxStages);

  const int b = blockIdx.x, G = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nst; ++s) {
      mbar_init(&r.full[s], 1);
      mbar_init(&r.empty[s], 1);
      r.released[s] = -1;
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int D = p.D, H = p.H, FF = p.FF;
  const int pos = *p.pos;
  const int tph = pos / kKT;                     // full KV tiles per head (keys [0, tph*32)); tail = the rest
  const int tailCTAs = (H + kSC - 1) / kSC;      // the last CTAs process the per-head tails (incl. the new key)
  const int Gr = G - tailCTAs;
  AttSplit as;
  as.H = H;
  as.tph = tph;
  as.WPH = max(min(Gr * kSC, 64 * H) / H, 1);    // ring warps per head (<= 64: bounds the partials per head)
  const StreamTables& tb = *p.tb;

  // phase list of one layer; `rot` spreads the remainder units of every phase over different CTAs
  auto gemv_phase = [&](int l, int which) {
    GemvPhase ph;
    ph.spu = 1;
    if (which == 0) { ph.W = tb.qkv_w[l]; ph.mode = SM_QKV; ph.U = 3 * D / 2; ph.K = D; ph.N = 3 * D; }
    else if (which == 1) { ph.W = tb.o_w[l]; ph.mode = SM_PLAIN2; ph.U = D / 2; ph.K = D; ph.N = D; }
    else if (which == 2) { ph.W = tb.gu_w[l]; ph.mode = SM_SWIGLU; ph.U = FF; ph.K = D; ph.N = 2 * FF; }
    else if (which == 3) {
      ph.W = tb.down_w[l]; ph.K = FF; ph.N = D;
      if (p.down_rows == 2) { ph.mode = SM_PLAIN2; ph.U = D / 2; } else { ph.mode = SM_PLAIN1; ph.U = D; ph.spu = p.down_split; }
    } else { ph.W = p.lm_head; ph.mode = SM_PLAIN2; ph.U = (p.V + 1) / 2; ph.K = D; ph.N = p.V; }
    return ph;
  };
  auto first_of = [&](int rot) {
    int f = b - rot;
    return f < 0 ? f + G : f;
  };

  if (warp == kSC) {
    // ================================================================ producer (one thread)
    if (lane != 0) return;
    ProducerState ps = {0, 0};
    int rot = 0;
    for (int l = 0; l < p.L; ++l) {
      GemvPhase ph = gemv_phase(l, 0);
      produce_gemv(r, ps, ph, first_of(rot), G, D);
      rot = (rot + ph.U) % G;
      if (b < Gr && tph > 0) {  // KV tiles of this CTA's ring warps, interleaved so stage order == consumption order
        int n[kSC], maxn = 0;
#pragma unroll
        for (int w = 0; w < kSC; ++w) {
          n[w] = att_count(as, b * kSC + w);
          maxn = max(maxn, n[w]);
        }
        const __nv_bfloat16* kvl = tb.kv[l];
        for (int jj = 0; jj < maxn; ++jj) {
#pragma unroll
          for (int w = 0; w < kSC; ++w) {
            if (jj >= n[w]) continue;
            const int gw = b * kSC + w;
            uint8_t* dst = producer_acquire(r, ps, kKvTileBytes);
            bulk_g2s(dst, kvl + kv_tile_off(att_head_of(as, gw), gw % as.WPH + jj * as.WPH, H), kKvTileBytes,
                     &r.full[ps.slot]);
            producer_advance(r, ps);
          }
        }
      }
      for (int which = 1; which <= 3; ++which) {
        ph = gemv_phase(l, which);
        produce_gemv(r, ps, ph, first_of(rot), G, D);
        rot = (rot + ph.U) % G;
      }
    }
    GemvPhase ph = gemv_phase(0, 4);
    produce_gemv(r, ps, ph, first_of(rot), G, D);
    return;
  }

  // ================================================================== consumers (8 warps)
  const int ctid = threadIdx.x;
  uint32_t stage = 0;
  unsigned int bar_target = 0;
  int rot = 0;
  const float scale_log2e = 1.4426950408889634f * rsqrtf(128.f);
  const int kslot = lane >> 3, sub = lane & 7;
  for (int l = 0; l < p.L; ++l) {
    const __nv_bfloat16* resid = (l == 0) ? p.x_in : p.x;
    __nv_bfloat16* kvl = tb.kv[l];
    // ---- RMSNorm -> fused q|k|v GEMV -> RoPE -> q buffer / KV-cache append
    stage_vec(resid, tb.ln1_w[l], D, p.eps, xs, red, ctid);
    GemvPhase ph = gemv_phase(l, 0);
    GemvEpilogue ep = {p.qbuf, nullptr, kvl, H, pos,
                       p.rope_cos + static_cast<size_t>(pos) * 64, p.rope_sin + static_cast<size_t>(pos) * 64};
    consume_gemv(r, stage, ph, first_of(rot), G, D, xs, ep, warp, lane);
    rot = (rot + ph.U) % G;
    bar_target += G;
    grid_barrier(p.bar, bar_target, ctid);

    // ---- attention over keys 0..pos
    const unsigned int nring = static_cast<unsigned int>(min(as.WPH, tph));   // ring partials per head
    if (b < Gr) {
      if (tph > 0) {
        int n[kSC], maxn = 0;
#pragma unroll
        for (int w = 0; w < kSC; ++w) {
          n[w] = att_count(as, b * kSC + w);
          maxn = max(maxn, n[w]);
        }
        const int gw = b * kSC + warp;
        const int my_n = att_count(as, gw), my_h = att_head_of(as, gw);
        AttState st;
        float qf[16];
        if (my_n > 0) {
          att_reset(st);
          att_load_q(p.qbuf, my_h, sub, scale_log2e, qf);
        }
        uint32_t c_row = stage;   // ring stage of (jj, warp 0)
        for (int jj = 0; jj < maxn; ++jj) {
          int before = 0, rowcnt = 0;
#pragma unroll
          for (int w = 0; w < kSC; ++w) {
            const int has = jj < n[w] ? 1 : 0;
            rowcnt += has;
            if (w < warp) before += has;
          }
          if (jj < my_n) {
            int slot;
            consumer_wait(r, c_row + static_cast<uint32_t>(before), slot, lane);
            const uint4* t4 = reinterpret_cast<const uint4*>(r.base + static_cast<size_t>(slot) * r.slot_bytes);
#pragma unroll 2
            for (int it = 0; it < kKT / 4; ++it) {
              const uint4* row = t4 + (it * 4 + kslot) * 32;   // [k 16 chunks | v 16 chunks]
              att_key(st, qf, row[sub], row[8 + sub], row[16 + sub], row[24 + sub], true);
            }
            consumer_release(r, slot, c_row + static_cast<uint32_t>(before), lane);
          }
          c_row += static_cast<uint32_t>(rowcnt);
        }
        stage = c_row;
        if (my_n > 0)
          att_flush(st, my_h, static_cast<unsigned int>(gw % as.WPH), nring + 1u, static_cast<unsigned int>(my_n),
                    static_cast<unsigned int>(tph) + 1u, p, lane);
      }
    } else {
      const int h = (b - Gr) * kSC + warp;
      if (h < H) {   // tail keys [tph*32, pos] straight from global memory (the new key was written above)
        AttState st;
        float qf[16];
        att_reset(st);
        att_load_q(p.qbuf, h, sub, scale_log2e, qf);
        const int k0 = tph * kKT;
        // the tail keys all live in token block `tph` of the cache
        const uint4* base = reinterpret_cast<const uint4*>(kvl + kv_tile_off(h, tph, H));
        auto load = [&](int key, uint4 (&q4)[4]) {
          const uint4* row = base + static_cast<size_t>(min(key, pos) - k0) * 32;
          q4[0] = __ldcg(row + sub);
          q4[1] = __ldcg(row + 8 + sub);
          q4[2] = __ldcg(row + 16 + sub);
          q4[3] = __ldcg(row + 24 + sub);
        };
        uint4 cur[4], nxt[4];
        load(k0 + kslot, cur);
        for (int kb = k0; kb <= pos; kb += 4) {
          const int key = kb + kslot;
          if (kb + 4 <= pos) load(key + 4, nxt);
          att_key(st, qf, cur[0], cur[1], cur[2], cur[3], key <= pos);
#pragma unroll
          for (int e = 0; e < 4; ++e) cur[e] = nxt[e];
        }
        att_flush(st, h, nring, nring + 1u, 1u, static_cast<unsigned int>(tph) + 1u, p, lane);
      }
    }
    bar_target += G;
    grid_barrier(p.bar, bar_target, ctid);

    // ---- x += o_proj(att)
    stage_vec(p.att, nullptr, D, 0.f, xs, red, ctid);
    ph = gemv_phase(l, 1);
    GemvEpilogue eo = {p.x, resid, nullptr, 0, 0, nullptr, nullptr};
    consume_gemv(r, stage, ph, first_of(rot), G, D, xs, eo, warp, lane);
    rot = (rot + ph.U) % G;
    bar_target += G;
    grid_barrier(p.bar, bar_target, ctid);

    // ---- RMSNorm -> gate|up GEMV -> SwiGLU
    stage_vec(p.x, tb.ln2_w[l], D, p.eps, xs, red, ctid);
    ph = gemv_phase(l, 2);
    GemvEpilogue eg = {p.mid, nullptr, nullptr, 0, 0, nullptr, nullptr};
    consume_gemv(r, stage, ph, first_of(rot), G, D, xs, eg, warp, lane);
    rot = (rot + ph.U) % G;
    bar_target += G;
    grid_barrier(p.bar, bar_target, ctid);

    // ---- x += down(mid)
    stage_vec(p.mid, nullptr, FF, 0.f, xs, red, ctid);
    ph = gemv_phase(l, 3);
    GemvEpilogue ed = {p.x, p.x, nullptr, 0, 0, nullptr, nullptr};
    consume_gemv(r, stage, ph, first_of(rot), G, D, xs, ed, warp, lane);
    rot = (rot + ph.U) % G;
    bar_target += G;
    grid_barrier(p.bar, bar_target, ctid);
  }
  // ---- final RMSNorm -> lm_head
  stage_vec(p.L > 0 ? p.x : p.x_in, p.norm_w, D, p.eps, xs, red, ctid);
  GemvPhase ph = gemv_phase(0, 4);
  GemvEpilogue el = {p.logits, nullptr, nullptr, 0, 0, nullptr, nullptr};
  consume_gemv(r, stage, ph, first_of(rot), G, D, xs, el, warp, lane);

  // ---- leave: the last CTA out re-arms the barrier counters for the next launch
  named_bar_sync(1, kSConsThreads);
  if (ctid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(p.exitc, 1u);
    if (t == static_cast<unsigned int>(G) - 1u) {
      *p.bar = 0;
      *p.exitc = 0;
      __threadfence();
    }
  }
}

// --------------------------------------------------------------------------------------------------- host
struct StreamGeometry {
  int slot_bytes, nst, down_rows, down_split, smem_bytes;
};
static int stream_geometry(int D, int FF, StreamGeometry* g) {
  const int budget = 227 * 1024 - 1024;  // barriers, reduction scratch, alignment slack
  const int xs_bytes = ((D > FF ? D : FF) * 2 + 127) / 128 * 128;
  int pair = 4 * D > kKvTileBytes ? 4 * D : kKvTileBytes;
  g->down_rows = (4 * FF <= pair) ? 2 : 1;
  g->down_split = 1;
  int slot = pair;
  if (g->down_rows == 1) {
    // a down_proj row (2 FF bytes) is streamed in K-chunks no larger than a row pair of the other phases, so the
    // ring slot stays small and the ring deep
    int chunks = (2 * FF + pair - 1) / pair;
    while (chunks <= 8 && FF % (8 * chunks) != 0) ++chunks;
    if (chunks <= 8) g->down_split = chunks;
    if (2 * FF / g->down_split > slot) slot = 2 * FF / g->down_split;
  }
  slot = (slot + 127) / 128 * 128;
  int nst = (budget - xs_bytes) / slot;
  if (nst > kMaxStages) nst = kMaxStages;
  if (nst < 2) return MERLIN_E_UNSUPPORTED;
  g->slot_bytes = slot;
  g->nst = nst;
  g->smem_bytes = nst * slot + xs_bytes + 2 * kMaxStages * 8 + kMaxStages * 4 + kSC * 4 + 128 + 128;
  return MERLIN_OK;
}

static int check_stream_weights(const merlin_decoder_weights* w) {
  MB_CHECK_ARG(w != nullptr, "decode_stream: null weights");
  MB_CHECK_ARG(w->hidden == w->heads * 128, "decode_stream: head_dim must be 128");
  MB_CHECK_ARG(w->num_layers >= 1 && w->num_layers <= kMaxLayers, "decode_stream: %d layers unsupported",
               w->num_layers);
  MB_CHECK_ARG(w->intermediate % 128 == 0 && w->hidden % 128 == 0 && w->vocab >= 1, "decode_stream: bad dims");
  MB_CHECK_ARG(w->kv_capacity >= 32 && w->kv_capacity % 32 == 0, "decode_stream: kv_capacity must be a multiple of 32");
  return MERLIN_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

size_t merlin_decode_stream_workspace_bytes(int num_layers, int hidden, int heads, int intermediate) {
  return stream_layout(num_layers, hidden, heads, intermediate).total;
}

int merlin_decode_stream_prepare(void* stream, const merlin_decoder_weights* w, void* workspace) {
  int rc = check_stream_weights(w);
  if (rc != MERLIN_OK) return rc;
  MB_CHECK_ARG(workspace != nullptr, "decode_stream_prepare: null workspace");
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "decode_stream_prepare: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const StreamLayout lay = stream_layout(w->num_layers, w->hidden, w->heads, w->intermediate);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  static thread_local StreamTables tb;
  memset(&tb, 0, sizeof(tb));
  for (int l = 0; l < w->num_layers; ++l) {
    tb.qkv_w[l] = static_cast<const __nv_bfloat16*>(w->qkv_w[l]);
    tb.o_w[l] = static_cast<const __nv_bfloat16*>(w->o_w[l]);
    tb.gu_w[l] = static_cast<const __nv_bfloat16*>(w->gate_up_w[l]);
    tb.down_w[l] = static_cast<const __nv_bfloat16*>(w->down_w[l]);
    tb.ln1_w[l] = static_cast<const __nv_bfloat16*>(w->ln1_w[l]);
    tb.ln2_w[l] = static_cast<const __nv_bfloat16*>(w->ln2_w[l]);
    tb.kv[l] = static_cast<__nv_bfloat16*>(w->kv_cache[l]);
    MB_CHECK_ARG((reinterpret_cast<uintptr_t>(tb.qkv_w[l]) | reinterpret_cast<uintptr_t>(tb.o_w[l]) |
                  reinterpret_cast<uintptr_t>(tb.gu_w[l]) | reinterpret_cast<uintptr_t>(tb.down_w[l]) |
                  reinterpret_cast<uintptr_t>(tb.kv[l])) % 16 == 0,
                 "decode_stream_prepare: layer %d tensors must be 16-byte aligned", l);
  }
  // synchronous copies: the staging buffers above are reused by the next call on this thread
  MB_CHECK_CUDA(cudaStreamSynchronize(st));
  MB_CHECK_CUDA(cudaMemcpy(ws + lay.tables, &tb, sizeof(tb), cudaMemcpyHostToDevice));
  MB_CHECK_CUDA(cudaMemset(ws + lay.ctrl, 0, lay.tables - lay.ctrl));  // barrier / exit / per-head counters
  return MERLIN_OK;
}

int merlin_decode_stream_step(void* stream, const merlin_decoder_weights* w, const void* x_in, const int* pos_dev,
                              void* workspace, void* logits) {
  int rc = check_stream_weights(w);
  if (rc != MERLIN_OK) return rc;
  MB_CHECK_ARG(x_in && pos_dev && workspace && logits, "decode_stream_step: null argument");
  StreamGeometry g;
  MB_CHECK_ARG(stream_geometry(w->hidden, w->intermediate, &g) == MERLIN_OK,
               "decode_stream_step: hidden=%d intermediate=%d do not fit the shared-memory ring", w->hidden,
               w->intermediate);
  const int G = sm_count();
  MB_CHECK_ARG(G > (w->heads + kSC - 1) / kSC, "decode_stream_step: %d SMs are too few for %d heads", G, w->heads);
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] { attr_rc = cudaFuncSetAttribute(decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  MB_CHECK_CUDA(attr_rc);
  const StreamLayout lay = stream_layout(w->num_layers, w->hidden, w->heads, w->intermediate);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  StreamParams p = {};
  p.tb = reinterpret_cast<const StreamTables*>(ws + lay.tables);
  p.norm_w = static_cast<const __nv_bfloat16*>(w->norm_w);
  p.lm_head = static_cast<const __nv_bfloat16*>(w->lm_head_w);
  p.x_in = static_cast<const __nv_bfloat16*>(x_in);
  p.x = reinterpret_cast<__nv_bfloat16*>(ws + lay.xres);
  p.logits = static_cast<__nv_bfloat16*>(logits);
  p.qbuf = reinterpret_cast<__nv_bfloat16*>(ws + lay.qbuf);
  p.att = reinterpret_cast<__nv_bfloat16*>(ws + lay.att);
  p.mid = reinterpret_cast<__nv_bfloat16*>(ws + lay.mid);
  p.part_o = reinterpret_cast<float*>(ws + lay.part_o);
  p.part_ml = reinterpret_cast<float*>(ws + lay.part_ml);
  p.done = reinterpret_cast<unsigned int*>(ws + lay.done);
  p.bar = reinterpret_cast<unsigned int*>(ws + lay.ctrl);
  p.exitc = reinterpret_cast<unsigned int*>(ws + lay.ctrl + 128);
  p.rope_cos = w->rope_cos;
  p.rope_sin = w->rope_sin;
  p.pos = pos_dev;
  p.L = w->num_layers;
  p.D = w->hidden;
  p.H = w->heads;
  p.FF = w->intermediate;
  p.V = w->vocab;
  p.eps = w->rms_eps;
  p.slot_bytes = g.slot_bytes;
  p.nst = g.nst;
  p.down_rows = g.down_rows;
  p.down_split = g.down_split;

  // cooperative launch: every CTA must be resident (grid-wide barriers); the runtime refuses the launch
  // instead of letting it hang when they cannot be.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(kSThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(g.smem_bytes);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeCooperative;
  attr.val.cooperative = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_stream_kernel, p));
  return MERLIN_OK;
}

}  // extern "C"
