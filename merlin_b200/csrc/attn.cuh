// merlin_b200 — flash-attention forward for sm_100a (tcgen05 + TMEM + TMA).
//
//   O = softmax(scale · Q Kᵀ  [+ causal mask] [+ key-length mask]) V        per (batch, head)
//
// One CTA owns a 128-row query tile of one head.  K/V tiles of 128 keys stream through a 2-stage
// TMA ring; S = Q·Kᵀ is accumulated by tcgen05.mma into one of two 128-column TMEM buffers, the four
// softmax warps (one query row per thread, no shuffles) read it with tcgen05.ld, write P (bf16) to a
// 128B-swizzled shared tile, and the MMA warp accumulates O += P·V in TMEM (V consumed MN-major
// straight from its row-major TMA tile).  The running maximum is only refreshed when it grows by
// more than 2^8 (exact — the final 1/l normalisation is consistent), so O is rarely rescaled.
//
// Used causal with head_dim 128 for the Vicuna decoder (replaces flash_attn_varlen_qkvpacked_func,
// mmgpt/utils/llama_flash_attn_monkey_patch.py:82-95) and non-causal with head_dim 64 for CLIP
// (HF:clip/modeling_clip.py:261-279).
#pragma once
#include "ptx.cuh"

namespace mb {

constexpr int kAttnThreads = 256;
constexpr int kAttnBlock = 128;  // query rows per CTA == keys per KV tile

struct AttnParams {
  int B, S, H;          // batch, (padded) sequence length, heads
  int q_col0, k_col0, v_col0;  // first column of the q / k / v block inside a packed qkv row
  float scale_log2e;    // softmax scale · log2(e)
  const int* seqlens;   // [B] valid (right-padded) lengths or nullptr (= S)
  float* lse;           // [B, H, S] natural-log-sum-exp of the scaled scores, or nullptr
};

template <int HD>
struct AttnCfg {
  static constexpr int kAtoms = HD / 64;                    // 64-column swizzle atoms per row
  static constexpr int kTileBytes = kAttnBlock * HD * 2;    // one Q / K / V tile
  static constexpr int kPBytes = kAttnBlock * kAttnBlock * 2;
  static constexpr int kSmemTiles = kTileBytes * 5 + kPBytes;  // Q + 2K + 2V + P
  static constexpr int kNumBars = 1 + 2 + 2 + 2 + 2 + 1 + 1;
  static constexpr int kSmemBytes = kSmemTiles + kNumBars * 8 + 16 + 1024;
};

template <int HD, bool kCausal>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                const AttnParams p) {
  using Cfg = AttnCfg<HD>;
  constexpr int kAtoms = Cfg::kAtoms;
  constexpr int kAtomBytes = kAttnBlock * 128;  // [128 rows x 64 bf16]

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sq = smem;
  uint8_t* sk = sq + Cfg::kTileBytes;          // 2 stages
  uint8_t* sv = sk + 2 * Cfg::kTileBytes;      // 2 stages
  uint8_t* sp = sv + 2 * Cfg::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemTiles);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;    // [2]
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int nq = (p.S + kAttnBlock - 1) / kAttnBlock;
  // grid = (heads, query tiles, batch): heads vary fastest, so CTAs are dispatched longest-first ACROSS heads
  const int q_tile = kCausal ? (nq - 1 - static_cast<int>(blockIdx.y)) : static_cast<int>(blockIdx.y);
  const int head = blockIdx.x;
  const int batch = blockIdx.z;
  const int q0 = q_tile * kAttnBlock;
  const int len = p.seqlens ? p.seqlens[batch] : p.S;
  int n_kv = (len + kAttnBlock - 1) / kAttnBlock;
  if (kCausal && n_kv > q_tile + 1) n_kv = q_tile + 1;

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
    }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one() && n_kv > 0) {
      mbar_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
      for (int a = 0; a < kAtoms; ++a)
        tma_load_3d(&tm_qkv, q_full, sq + a * kAtomBytes, p.q_col0 + head * HD + a * 64, q0, batch,
                    kEvictFirst);
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
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S = Q Kᵀ : both K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);   // O += P V : V is MN-major
      const uint32_t sq_a = smem_u32(sq);
      const uint32_t sp_a = smem_u32(sp);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sk_a = smem_u32(sk + st * Cfg::kTileBytes);
        const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(st * 128);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(tmem_s, umma_desc_k_sw128(sq_a + off), umma_desc_k_sw128(sk_a + off), idesc_s,
                  kk != 0);
        }
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sv_a = smem_u32(sv + st * Cfg::kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kAttnBlock / 16; ++kk) {
          // A = P[128 x 16 keys] (K-major), B = V[16 keys x HD] (MN-major, 16 key rows = 2 KB)
          const uint32_t a_off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(tmem_o, umma_desc_k_sw128(sp_a + a_off),
                  umma_desc_mn_sw128(sv_a + kk * 2048, kAtomBytes), idesc_o, (j | kk) != 0);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================================ softmax + epilogue
    const int ewarp = warp - 4;
    const int row = ewarp * 32 + static_cast<int>(lane_id());
    const int q_idx = q0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>(ewarp * 32) << 16;
    float m_used = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const int kv0 = j * kAttnBlock;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[128];
      {
        const uint32_t ts = tmem_base + lane_sel + static_cast<uint32_t>(st * 128);
        tmem_ld32(ts, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld32(ts + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
        tmem_ld32(ts + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
        tmem_ld_wait();
      }
      const bool need_mask = (kCausal && j == q_tile) || (kv0 + kAttnBlock > len);
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const int kv = kv0 + c;
          const bool dead = (kCausal && kv > q_idx) || kv >= len;
          if (dead) s[c] = 0xff800000u;  // -inf
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
        alpha = fast_exp2(m_used - m_new);  // 0 on the first tile (m_used = -inf)
      }
      float rowsum = 0.f;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        const float pv = fast_exp2(fmaf(__uint_as_float(s[c]), p.scale_log2e, -m_new));
        rowsum += pv;
        s[c] = __float_as_uint(pv);
      }
      l = l * alpha + rowsum;
      m_used = m_new;

      if (j > 0) {
        // P buffer is free and O is stable once PV(j-1) has completed.
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int ch = 0; ch < HD / 32; ++ch) {
            uint32_t o[32];
            tmem_ld32(tmem_o + lane_sel + ch * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st32(tmem_o + lane_sel + ch * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P (bf16) -> shared, K-major [128 rows x 128 keys] as two 128B-swizzled 64-key atoms
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(s[ch * 8 + 0]), __uint_as_float(s[ch * 8 + 1]));
        w.y = pack_bf16(__uint_as_float(s[ch * 8 + 2]), __uint_as_float(s[ch * 8 + 3]));
        w.z = pack_bf16(__uint_as_float(s[ch * 8 + 4]), __uint_as_float(s[ch * 8 + 5]));
        w.w = pack_bf16(__uint_as_float(s[ch * 8 + 6]), __uint_as_float(s[ch * 8 + 7]));
        *reinterpret_cast<uint4*>(sp + (ch >> 3) * kAtomBytes + sw128_offset(row, ch & 7)) = w;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // ------------------------------------------------------ epilogue
    const bool row_valid = q_idx < len && n_kv > 0;
    if (n_kv > 0) {
      mbar_wait(pv_done, (n_kv - 1) & 1);
      tc_fence_after();
    }
    const float inv_l = row_valid ? 1.f / l : 0.f;
    // stage O (bf16) in the Q tile: every S MMA that read Q has completed by now
#pragma unroll
    for (int ch = 0; ch < HD / 32; ++ch) {
      uint32_t o[32];
      if (n_kv > 0) {
        tmem_ld32(tmem_o + lane_sel + ch * 32, o);
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
        const int chunk = ch * 4 + q4;  // 16-byte chunk index inside the HD-wide row
        *reinterpret_cast<uint4*>(sq + (chunk >> 3) * kAtomBytes + sw128_offset(row, chunk & 7)) = w;
      }
    }
    if (p.lse != nullptr && q_idx < p.S) {
      const float lse = row_valid ? (m_used + log2f(l)) * 0.6931471805599453f : -INFINITY;
      p.lse[(static_cast<size_t>(batch) * p.H + head) * p.S + q_idx] = lse;
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (threadIdx.x == 128) {
#pragma unroll
      for (int a = 0; a < kAtoms; ++a)
        tma_store_3d(&tm_o, sq + a * kAtomBytes, head * HD + a * 64, q0, batch);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace mb
