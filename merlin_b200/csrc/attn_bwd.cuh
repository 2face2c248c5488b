// merlin_b200 — flash-attention backward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64 / 128.
//
// One CTA owns a 128-key tile j of one head and walks the query tiles i that attend to it
// (i >= j when causal).  Per (i, j):
//     S  = Q_i K_jᵀ            dP = dO_i V_jᵀ                      (tcgen05 -> TMEM, fp32)
//     P  = exp2(S·c − lse_i)    dS = P ∘ (dP − δ_i) · scale         (256 threads: query row x column half)
//     dV += Pᵀ dO_i            dK += dSᵀ Q_i            dQ_i = dS K_j
// dK / dV accumulate in TMEM across the whole walk and are written once; dQ_i is reduced into an fp32
// global buffer with TMA reduce-add (the cross-CTA reduction over key tiles).  P and dS are written
// once to 128B-swizzled shared tiles and consumed both K-major (dQ) and MN-major (dV, dK: the same
// bytes, described with the M and K roles swapped), as are the Q / dO / K / V row-major TMA tiles.
// TMEM: S|dQ [0,128)  dP [128,256)  dV [256,384)  dK [384,512)  — all 512 columns, so S / dP cannot be
// double-buffered, and shared memory (K, V, Q, dO, P, dS = 192 KB) has no room for a second Q / dO stage.
// The overlap therefore happens INSIDE a step: the tensor pipe and the softmax warps alternate on halves of it
//
//   tensor pipe :  S ─────►  dP ─────────►  dV ─────────►  dK ──►  dQ ──►            (S of step s+1 ...)
//   softmax     :            P = f(S) ───►  dS = g(P, dP) ►                 drain dQ, TMA reduce-add
//   TMA loads   :                                           dO(s+1) ◄── after dV    Q(s+1) ◄── after dK
//
// i.e. P is computed while dP is being multiplied, dS while dV is, and the next step's operands arrive under
// dK / dQ and the dQ drain.  Round 1 ran these phases back to back with 4 softmax warps and measured 18 % tensor
// pipe activity (profiles/r2_train_kernels_ncu.txt); two warps per TMEM lane quarter halve the softmax time and
// keep the per-thread state at 64 columns (no local-memory spills).
#pragma once
#include "ptx.cuh"

namespace mb {

constexpr int kBwdSoftmaxWarps = 8;
constexpr int kBwdSoftmaxThreads = kBwdSoftmaxWarps * 32;
constexpr int kBwdThreads = 128 + kBwdSoftmaxThreads;   // warps 0-3: TMA, MMA, TMEM alloc, spare
constexpr int kBwdBlock = 128;

struct AttnBwdParams {
  int B, S, H;
  int q_col0, k_col0, v_col0;   // column offsets inside packed qkv / dqkv rows
  float scale;                  // softmax scale
  float scale_log2e;
  const int* seqlens;           // [B] or nullptr
  const float* lse;             // [B, H, S] natural log-sum-exp from the forward
  const float* delta;           // [B, H, S] rowsum(dO ∘ O)
  float* dq_accum;              // [B*S, H*128] fp32, zero-initialised
  __nv_bfloat16* dqkv;          // [B*S, ld_dqkv] packed gradient (k and v blocks written here)
  int64_t ld_dqkv;
};

template <int HD>
struct AttnBwdCfg {
  static constexpr int kTileBytes = kBwdBlock * HD * 2;         // [128 rows x HD] bf16: K, V, Q, dO
  static constexpr int kPBytes = kBwdBlock * kBwdBlock * 2;     // [128 x 128] bf16: P, dS
  static constexpr int kSmemTiles = kTileBytes * 4 + kPBytes * 2;
  static constexpr int kNumBars = 12;
  static constexpr int kSmemBytes = kSmemTiles + kNumBars * 8 + 16 + 1024;
};

template <int HD, bool kCausal>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const AttnBwdParams p) {
  using Cfg = AttnBwdCfg<HD>;
  constexpr int kAtoms = HD / 64;
  constexpr int kAtomBytes = kBwdBlock * 128;   // 16 KB: [128 rows x 64 bf16]

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sk = smem;
  uint8_t* sv = sk + Cfg::kTileBytes;
  uint8_t* sq = sv + Cfg::kTileBytes;
  uint8_t* sdo = sq + Cfg::kTileBytes;
  uint8_t* sp = sdo + Cfg::kTileBytes;
  uint8_t* sds = sp + Cfg::kPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemTiles);
  uint64_t* kv_full = bars;        // K_j, V_j landed (once)
  uint64_t* q_full = bars + 1;     // Q_i landed
  uint64_t* do_full = bars + 2;    // dO_i landed
  uint64_t* q_empty = bars + 3;    // S and dK MMAs of step i done: Q_i reusable
  uint64_t* do_empty = bars + 4;   // dP and dV MMAs of step i done: dO_i reusable
  uint64_t* s_full = bars + 5;     // S in TMEM
  uint64_t* dp_full = bars + 6;    // dP in TMEM
  uint64_t* p_full = bars + 7;     // P in shared memory (256 arrivals)
  uint64_t* ds_full = bars + 8;    // dS in shared memory (256 arrivals)
  uint64_t* dq_full = bars + 9;    // every MMA of step i complete (dQ in TMEM; P / dS staging reusable)
  uint64_t* dq_read = bars + 10;   // dQ drained from TMEM (256 arrivals): S region reusable
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int j = blockIdx.x;   // key tile
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int len = p.seqlens ? p.seqlens[batch] : p.S;
  const int nq = (min(len, p.S) + kBwdBlock - 1) / kBwdBlock;   // query tiles that contain valid rows
  const int kv0 = j * kBwdBlock;
  const int i_begin = kCausal ? j : 0;
  const bool has_work = kv0 < len && i_begin < nq;
  const int n_steps = has_work ? (nq - i_begin) : 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(do_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_full, kBwdSoftmaxThreads);
    mbar_init(ds_full, kBwdSoftmaxThreads);
    mbar_init(dq_full, 1);
    mbar_init(dq_read, kBwdSoftmaxThreads);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 384;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one() && n_steps > 0) {
      mbar_expect_tx(kv_full, 2 * Cfg::kTileBytes);
#pragma unroll
      for (int a = 0; a < kAtoms; ++a) {
        tma_load_3d(&tm_qkv, kv_full, sk + a * kAtomBytes, p.k_col0 + head * HD + a * 64, kv0, batch, kEvictFirst);
        tma_load_3d(&tm_qkv, kv_full, sv + a * kAtomBytes, p.v_col0 + head * HD + a * 64, kv0, batch, kEvictFirst);
      }
      for (int s = 0; s < n_steps; ++s) {
        const int q0 = (i_begin + s) * kBwdBlock;
        mbar_wait(do_empty, (s & 1) ^ 1);          // frees first: right after dV of the previous step
        mbar_expect_tx(do_full, Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_load_3d(&tm_do, do_full, sdo + a * kAtomBytes, head * HD + a * 64, q0, batch, kEvictLast);
        mbar_wait(q_empty, (s & 1) ^ 1);           // after dK of the previous step
        mbar_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_load_3d(&tm_qkv, q_full, sq + a * kAtomBytes, p.q_col0 + head * HD + a * 64, q0, batch, kEvictLast);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    if (elect_one() && n_steps > 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);   // S, dP : A, B K-major
      constexpr uint32_t idesc_mm = umma_idesc_bf16(128, HD, 1, 1);    // dV, dK: A, B MN-major
      constexpr uint32_t idesc_km = umma_idesc_bf16(128, HD, 0, 1);    // dQ    : A K-major, B MN-major
      const uint32_t a_k = smem_u32(sk), a_v = smem_u32(sv), a_q = smem_u32(sq), a_do = smem_u32(sdo),
                     a_p = smem_u32(sp), a_ds = smem_u32(sds);
      mbar_wait(kv_full, 0);
      for (int s = 0; s < n_steps; ++s) {
        const uint32_t ph = s & 1;
        // ---- S = Q Kᵀ  (the S region held dQ of the previous step until the softmax warps drained it)
        mbar_wait(q_full, ph);
        mbar_wait(dq_read, ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(t_s, umma_desc_k_sw128(a_q + off), umma_desc_k_sw128(a_k + off), idesc_kk, kk != 0);
        }
        umma_commit(s_full);
        // ---- dP = dO Vᵀ  (runs while the softmax warps turn S into P; dP of the previous step was consumed
        //      before its dS was published, which this thread waited for)
        mbar_wait(do_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(t_dp, umma_desc_k_sw128(a_do + off), umma_desc_k_sw128(a_v + off), idesc_kk, kk != 0);
        }
        umma_commit(dp_full);
        // ---- dV += Pᵀ dO  (runs while the softmax warps turn P, dP into dS)
        mbar_wait(p_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBwdBlock / 16; ++kk) {
          // reduction over 16 query rows per instruction: every operand advances 16 rows = 2 KB
          const uint32_t roff = kk * 2048;
          umma_ss(t_dv, umma_desc_mn_sw128(a_p + roff, kAtomBytes), umma_desc_mn_sw128(a_do + roff, kAtomBytes),
                  idesc_mm, (s | kk) != 0);
        }
        umma_commit(do_empty);
        // ---- dK += dSᵀ Q, then dQ = dS K into the S region (S was consumed before P was published)
        mbar_wait(ds_full, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBwdBlock / 16; ++kk) {
          const uint32_t roff = kk * 2048;
          umma_ss(t_dk, umma_desc_mn_sw128(a_ds + roff, kAtomBytes), umma_desc_mn_sw128(a_q + roff, kAtomBytes),
                  idesc_mm, (s | kk) != 0);
        }
        umma_commit(q_empty);
#pragma unroll
        for (int kk = 0; kk < kBwdBlock / 16; ++kk) {
          // dQ = dS K: reduction over 16 keys: A advances inside the key-major rows, B by 16 key rows
          const uint32_t aoff = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(t_s, umma_desc_k_sw128(a_ds + aoff), umma_desc_mn_sw128(a_k + kk * 2048, kAtomBytes), idesc_km,
                  kk != 0);
        }
        umma_commit(dq_full);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== P / dS / dQ threads: (query row) x (column half)
    const int sw = warp - 4;
    const int quarter = sw & 3;                     // TMEM lane quarter this warp may touch (warp id % 4)
    const int half = sw >> 2;                       // key columns [64 half, 64 half + 64)
    const int row = quarter * 32 + static_cast<int>(lane_id());
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const bool leader = threadIdx.x == 128;
    const float log2e = 1.4426950408889634f;
    for (int s = 0; s < n_steps; ++s) {
      const uint32_t ph = s & 1;
      const int i = i_begin + s;
      const int q_idx = i * kBwdBlock + row;
      const bool q_valid = q_idx < len;
      const size_t stat = (static_cast<size_t>(batch) * p.H + head) * p.S + (q_valid ? q_idx : 0);
      const float lse2 = q_valid ? p.lse[stat] * log2e : 0.f;
      const float dlt = q_valid ? p.delta[stat] : 0.f;
      float pv[64];
      // ---- P = exp2(S c − lse)
      mbar_wait(s_full, ph);
      tc_fence_after();
      if (s > 0) {
        // the previous step's dQ reduction must have finished READING the P|dS staging bytes
        if (leader) tma_store_wait_read<0>();
        named_bar_sync(1, kBwdSoftmaxThreads);
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;               // 32-key chunk
        uint32_t sv_[32];
        tmem_ld32(t_s + lane_sel + ch * 32, sv_);
        tmem_ld_wait();
        uint32_t pw[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int kv = kv0 + ch * 32 + c + e;
            const bool dead = !q_valid || kv >= len || (kCausal && kv > q_idx);
            pv[c2 * 32 + c + e] = dead ? 0.f : fast_exp2(fmaf(__uint_as_float(sv_[c + e]), p.scale_log2e, -lse2));
          }
          pw[c >> 1] = pack_bf16(pv[c2 * 32 + c], pv[c2 * 32 + c + 1]);
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row inside atom (ch >> 1)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = (ch >> 1) * kAtomBytes + sw128_offset(row, (ch & 1) * 4 + q4);
          *reinterpret_cast<uint4*>(sp + off) = make_uint4(pw[q4 * 4], pw[q4 * 4 + 1], pw[q4 * 4 + 2], pw[q4 * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // ---- dS = P ∘ (dP − δ) · scale
      mbar_wait(dp_full, ph);
      tc_fence_after();
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;
        uint32_t dpv[32];
        tmem_ld32(t_dp + lane_sel + ch * 32, dpv);
        tmem_ld_wait();
        uint32_t dw[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float d0 = pv[c2 * 32 + c] * (__uint_as_float(dpv[c]) - dlt) * p.scale;
          const float d1 = pv[c2 * 32 + c + 1] * (__uint_as_float(dpv[c + 1]) - dlt) * p.scale;
          dw[c >> 1] = pack_bf16(d0, d1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = (ch >> 1) * kAtomBytes + sw128_offset(row, (ch & 1) * 4 + q4);
          *reinterpret_cast<uint4*>(sds + off) = make_uint4(dw[q4 * 4], dw[q4 * 4 + 1], dw[q4 * 4 + 2], dw[q4 * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);

      // ---- dQ_i tile: TMEM -> fp32 128B-swizzled staging (aliasing P|dS, which the MMAs have finished
      // reading) -> ONE TMA reduce-add per [128 rows x 32 floats] box into the fp32 dQ accumulator
      mbar_wait(dq_full, ph);
      tc_fence_after();
#pragma unroll
      for (int c2 = 0; c2 < HD / 64; ++c2) {
        const int ch = half * (HD / 64) + c2;       // 32-float column chunk of dQ
        uint32_t v[32];
        tmem_ld32(t_s + lane_sel + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<uint4*>(sp + ch * kAtomBytes + sw128_offset(row, c16)) =
              make_uint4(v[c16 * 4], v[c16 * 4 + 1], v[c16 * 4 + 2], v[c16 * 4 + 3]);
      }
      tc_fence_before();
      mbar_arrive(dq_read);
      fence_proxy_async_smem();
      named_bar_sync(1, kBwdSoftmaxThreads);
      if (leader) {
#pragma unroll
        for (int ch = 0; ch < HD / 32; ++ch)
          tma_reduce_add_2d(&tm_dq, sp + ch * kAtomBytes, head * HD + ch * 32, batch * p.S + i * kBwdBlock);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all<0>();

    // ---- dK, dV of this key tile (thread = key row x column half).  The TMEM loads are warp-collective
    // (.sync.aligned): every lane executes them, only the global stores are predicated.
    const int kv_idx = kv0 + row;
    const bool in_range = kv_idx < p.S;
    const bool live = n_steps > 0 && kv_idx < len;
    __nv_bfloat16* out_row = p.dqkv + (static_cast<size_t>(batch) * p.S + (in_range ? kv_idx : 0)) * p.ld_dqkv;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t t_acc = which == 0 ? t_dk : t_dv;
      __nv_bfloat16* dst = out_row + (which == 0 ? p.k_col0 : p.v_col0) + head * HD;
#pragma unroll
      for (int c2 = 0; c2 < HD / 64; ++c2) {
        const int ch = half * (HD / 64) + c2;
        uint32_t v[32];
        if (n_steps > 0) {  // block-uniform
          tmem_ld32(t_acc + lane_sel + ch * 32, v);
          tmem_ld_wait();
        }
        if (in_range) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (live) {
              w.x = pack_bf16(__uint_as_float(v[q4 * 8 + 0]), __uint_as_float(v[q4 * 8 + 1]));
              w.y = pack_bf16(__uint_as_float(v[q4 * 8 + 2]), __uint_as_float(v[q4 * 8 + 3]));
              w.z = pack_bf16(__uint_as_float(v[q4 * 8 + 4]), __uint_as_float(v[q4 * 8 + 5]));
              w.w = pack_bf16(__uint_as_float(v[q4 * 8 + 6]), __uint_as_float(v[q4 * 8 + 7]));
            }
            *reinterpret_cast<uint4*>(dst + ch * 32 + q4 * 8) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// dq (fp32 accumulator, [M, H*128]) -> bf16 q block of the packed gradient rows
__global__ void dq_to_bf16_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, int64_t ld,
                                  int q_col0, int width, int64_t rows) {
  const int vec_per_row = width >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < rows * vec_per_row;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const float4 a = *reinterpret_cast<const float4*>(acc + r * width + c);
    const float4 b = *reinterpret_cast<const float4*>(acc + r * width + c + 4);
    *reinterpret_cast<uint4*>(dqkv + r * ld + q_col0 + c) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

}  // namespace mb
