// merlin_b200 — flash-attention forward v2 for sm_100a: two query tiles per CTA, ping-pong softmax.
//
// One CTA owns 256 query rows of one head as two 128-row tiles X in {A, B}.  Each tile has its own
// S/P accumulator (128 TMEM columns) and O accumulator (HD columns); a single MMA-issuing thread
// interleaves   S_A(j) S_B(j) | PV_A(j) S_A(j+1) | PV_B(j) S_B(j+1) | ...   so that while softmax warpgroup A
// works on S_A the tensor core runs tile B's MMAs and vice versa.  P never touches shared memory: the
// softmax threads overwrite the first 64 columns of their S buffer with packed bf16 P (tcgen05.st) and
// the PV MMA takes its A operand from TMEM.  K/V tiles of 128 keys stream through a 2-stage TMA ring and
// are shared by both query tiles (half the K/V traffic of v1).
//
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-7 softmax A, 8-11 softmax B.
#pragma once
#include "attn.cuh"

namespace mb {

constexpr int kAttn2Threads = 384;

#ifndef MERLIN_ATTN_POLY_EVERY
#define MERLIN_ATTN_POLY_EVERY 2     // every 2nd pair -> 25 % of the exponentials on the FMA pipe
#endif
constexpr int kPolyEvery = MERLIN_ATTN_POLY_EVERY;

template <int HD>
struct Attn2Cfg {
  static constexpr int kAtoms = HD / 64;
  static constexpr int kTileBytes = kAttnBlock * HD * 2;
  static constexpr int kSmemTiles = kTileBytes * 6;  // Q_A, Q_B, 2 x K, 2 x V
  static constexpr int kNumBars = 1 + 2 + 2 + 2 + 2 + 2 + 2;
  static constexpr int kSmemBytes = kSmemTiles + kNumBars * 8 + 16 + 1024;
};

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

template <int HD, bool kCausal>
__global__ void __launch_bounds__(kAttn2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                 const AttnParams p) {
  using Cfg = Attn2Cfg<HD>;
  constexpr int kAtoms = Cfg::kAtoms;
  constexpr int kAtomBytes = kAttnBlock * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sq = smem;                          // Q_A | Q_B
  uint8_t* sk = sq + 2 * Cfg::kTileBytes;      // 2 stages
  uint8_t* sv = sk + 2 * Cfg::kTileBytes;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemTiles);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;    // [2] per query tile
  uint64_t* p_full = bars + 9;    // [2] per query tile, 128 arrivals
  uint64_t* o_full = bars + 11;   // [2] per query tile
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int nq = (p.S + kAttnBlock - 1) / kAttnBlock;
  const int npairs = (nq + 1) / 2;
  // grid = (heads, tile pairs, batch): heads vary fastest, so CTAs are dispatched longest-first ACROSS heads
  const int pair = kCausal ? (npairs - 1 - static_cast<int>(blockIdx.y)) : static_cast<int>(blockIdx.y);
  const int head = blockIdx.x;
  const int batch = blockIdx.z;
  const int len = p.seqlens ? p.seqlens[batch] : p.S;
  const int n_len = (len + kAttnBlock - 1) / kAttnBlock;
  // per query tile: index, first row, number of KV tiles (0 = tile does not exist / nothing to attend)
  int qt[2], q0[2], nkv[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    qt[x] = 2 * pair + x;
    q0[x] = qt[x] * kAttnBlock;
    nkv[x] = (qt[x] < nq) ? (kCausal ? min(n_len, qt[x] + 1) : n_len) : 0;
  }
  const int n_kv = max(nkv[0], nkv[1]);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_o);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns: S_A [0,128)  S_B [128,256)  O_A [256,256+HD)  O_B [384,384+HD)

  if (warp < 4) {
    reg_dealloc<64>();
    if (warp == 0) {
      // ===================================================== TMA producer
      if (elect_one() && n_kv > 0) {
        const int q_tiles = (nkv[1] > 0) ? 2 : 1;
        mbar_expect_tx(q_full, q_tiles * Cfg::kTileBytes);
        for (int x = 0; x < q_tiles; ++x)
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            tma_load_3d(&tm_qkv, q_full, sq + x * Cfg::kTileBytes + a * kAtomBytes,
                        p.q_col0 + head * HD + a * 64, q0[x], batch, kEvictFirst);
        for (int j = 0; j < n_kv; ++j) {
          const int st = j & 1;
          mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], Cfg::kTileBytes);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            tma_load_3d(&tm_qkv, &k_full[st], sk + st * Cfg::kTileBytes + a * kAtomBytes,
                        p.k_col0 + head * HD + a * 64, j * kAttnBlock, batch, kEvictLast);
          mbar_expect_tx(&v_full[st], Cfg::kTileBytes);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            tma_load_3d(&tm_qkv, &v_full[st], sv + st * Cfg::kTileBytes + a * kAtomBytes,
                        p.v_col0 + head * HD + a * 64, j * kAttnBlock, batch, kEvictLast);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ======================================================= MMA issuer
      if (elect_one() && n_kv > 0) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
        auto issue_s = [&](int x, int j) {
          const int st = j & 1;
          mbar_wait(&k_full[st], (j >> 1) & 1);
          tc_fence_after();
          const uint32_t sq_a = smem_u32(sq + x * Cfg::kTileBytes);
          const uint32_t sk_a = smem_u32(sk + st * Cfg::kTileBytes);
          const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(x * 128);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
            umma_ss(tmem_s, umma_desc_k_sw128(sq_a + off), umma_desc_k_sw128(sk_a + off), idesc_s, kk != 0);
          }
          umma_commit(&s_full[x]);
        };
        mbar_wait(q_full, 0);
        for (int x = 0; x < 2; ++x)
          if (nkv[x] > 0) issue_s(x, 0);
        for (int j = 0; j < n_kv; ++j) {
          const int st = j & 1;
          for (int x = 0; x < 2; ++x) {
            if (j < nkv[x]) {
              mbar_wait(&p_full[x], j & 1);
              mbar_wait(&v_full[st], (j >> 1) & 1);
              tc_fence_after();
              const uint32_t sv_a = smem_u32(sv + st * Cfg::kTileBytes);
              const uint32_t tmem_p = tmem_base + static_cast<uint32_t>(x * 128);
              const uint32_t tmem_o = tmem_base + 256u + static_cast<uint32_t>(x * 128);
#pragma unroll
              for (int kk = 0; kk < kAttnBlock / 16; ++kk) {
                // A = P[128 x 16 keys] in TMEM (8 columns of packed bf16 pairs), B = V rows (MN-major)
                umma_ts(tmem_o, tmem_p + kk * 8, umma_desc_mn_sw128(sv_a + kk * 2048, kAtomBytes), idesc_o,
                        (j | kk) != 0);
              }
              if (j + 1 == nkv[x]) umma_commit(&o_full[x]);
            }
            // K/V stage j is free once the last tile that uses it has issued its PV
            if (x == 1 || nkv[1] <= j) {
              if (x == (nkv[1] > j ? 1 : 0)) umma_commit(&kv_empty[st]);
            }
            if (j + 1 < nkv[x]) issue_s(x, j + 1);
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ============================================ softmax + epilogue (two warpgroups)
    reg_alloc<216>();
    const int x = (warp - 4) >> 2;                 // query tile of this warpgroup
    const int ewarp = warp & 3;                    // TMEM lane quarter
    const int row = ewarp * 32 + static_cast<int>(lane_id());
    const int q_idx = q0[x] + row;
    const uint32_t lane_sel = static_cast<uint32_t>(ewarp * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_sel + static_cast<uint32_t>(x * 128);
    const uint32_t tmem_o = tmem_base + lane_sel + 256u + static_cast<uint32_t>(x * 128);
    const int my_nkv = nkv[x];
    float m_used = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < my_nkv; ++j) {
      const int kv0 = j * kAttnBlock;
      mbar_wait(&s_full[x], j & 1);   // also implies PV_x(j-1) finished (in-order tensor pipe)
      tc_fence_after();
      uint32_t s[128];
      tmem_ld32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
      tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
      tmem_ld_wait();
      const bool need_mask = (kCausal && j == qt[x]) || (kv0 + kAttnBlock > len);
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const int kv = kv0 + c;
          const bool dead = (kCausal && kv > q_idx) || kv >= len;
          if (dead) s[c] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
      const float mx_scaled = mx * p.scale_log2e;
      float m_new = m_used;
      float alpha = 1.f;
      if (mx_scaled > m_used + 8.f) {
        m_new = mx_scaled;
        alpha = fast_exp2(m_used - m_new);
      }
      float rowsum = 0.f;
      uint32_t pk[64];
      // kPolyEvery-th exponential runs as a cubic on the FMA pipe instead of ex2.approx: with every exponent on the
      // 16-per-clock MUFU unit the kernel measured XU = 48 % next to tensor = 48 % (profiles/r1_attn_v2_metrics.csv)
#pragma unroll
      for (int c = 0; c < 128; c += 2) {
        const float x0 = fmaf(__uint_as_float(s[c]), p.scale_log2e, -m_new);
        const float x1 = fmaf(__uint_as_float(s[c + 1]), p.scale_log2e, -m_new);
        const float p0 = fast_exp2(x0);
        const float p1 = ((c >> 1) % kPolyEvery == kPolyEvery - 1) ? exp2_poly3(x1) : fast_exp2(x1);
        rowsum += p0 + p1;
        pk[c >> 1] = pack_bf16(p0, p1);
      }
      l = l * alpha + rowsum;
      m_used = m_new;
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int ch = 0; ch < HD / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tmem_o + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st32(tmem_o + ch * 32, o);
        }
      }
      // P (bf16 pairs) overwrites the first 64 columns of this tile's S buffer
      tmem_st32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[x]);
    }

    // ------------------------------------------------------ epilogue of this query tile
    if (qt[x] < nq) {
      const bool row_valid = q_idx < len && my_nkv > 0;
      if (my_nkv > 0) {
        mbar_wait(&o_full[x], 0);
        tc_fence_after();
      }
      const float inv_l = row_valid ? 1.f / l : 0.f;
      uint8_t* stage = sq + x * Cfg::kTileBytes;  // Q_x is dead: every S_x MMA has completed
#pragma unroll
      for (int ch = 0; ch < HD / 32; ++ch) {
        uint32_t o[32];
        if (my_nkv > 0) {
          tmem_ld32(tmem_o + ch * 32, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = 0;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[q4 * 8 + 0]) * inv_l, __uint_as_float(o[q4 * 8 + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(o[q4 * 8 + 2]) * inv_l, __uint_as_float(o[q4 * 8 + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(o[q4 * 8 + 4]) * inv_l, __uint_as_float(o[q4 * 8 + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(o[q4 * 8 + 6]) * inv_l, __uint_as_float(o[q4 * 8 + 7]) * inv_l);
          const int chunk = ch * 4 + q4;
          *reinterpret_cast<uint4*>(stage + (chunk >> 3) * kAtomBytes + sw128_offset(row, chunk & 7)) = w;
        }
      }
      if (p.lse != nullptr && q_idx < p.S) {
        const float lse = row_valid ? (m_used + log2f(l)) * 0.6931471805599453f : -INFINITY;
        p.lse[(static_cast<size_t>(batch) * p.H + head) * p.S + q_idx] = lse;
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + x, 128);
      if ((threadIdx.x & 127) == 0) {
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_store_3d(&tm_o, stage + a * kAtomBytes, head * HD + a * 64, q0[x], batch);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace mb
