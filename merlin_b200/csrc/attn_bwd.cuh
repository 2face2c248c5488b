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
// TMEM: S [0,128)  dP|dQ [128,256)  dV [256,384)  dK [384,512)  — all 512 columns, so S / dP cannot be
// double-buffered; shared memory holds K, V, TWO Q stages, dO, P, dS = 224 KB.  The overlap happens inside and
// across steps — the tensor pipe and the softmax warps alternate on halves of a step, and S of the NEXT step is
// issued as soon as P of this one has left the S columns:
//
//   tensor pipe :  dP(s) ───►  dV(s) ─► dQ(s) ─► dK(s) ─► S(s+1) ─►  dP(s+1) ...
//   softmax     :  P = f(S(s)) ─►  dS = g(P, dP) ────►  drain dQ(s), TMA reduce-add ─►  P = f(S(s+1)) ...
//   TMA loads   :  Q(s+2) after dK(s) into the other Q stage;  dO(s+1) after dV(s)
//
// dQ accumulates in the dP columns (dP is dead once dS is published), so the drain of dQ is off the path to the
// next S.  Measured by ablation (scripts/attn_bwd_bench.py): the P / dS arithmetic is ~1 % of the kernel; what
// bounds it is the chain of dependent hops per step, which this schedule shortens from S→P→dS→dK→(load Q)→S to
// S→P→dS→drain.  Round 1 ran all phases back to back with 4 softmax warps at 18 % tensor-pipe activity
// (profiles/r2_train_kernels_ncu.txt); two warps per TMEM lane quarter keep the per-thread state at 64 columns
// (no local-memory spills).
#pragma once
#include "ptx.cuh"

namespace mb {

#ifdef MERLIN_ATTN_TRACE
// bring-up only (libmerlin_b200_dbg.so): %globaltimer stamps of CTA (0,0,0), [step][event]; events 0-7 = MMA thread
// (S issued, dP issued, p_full seen, dV issued, next S issued, ds_full seen, dK+dQ issued), 8-15 = softmax leader
// (s_full seen, P published, dp_full seen, dS published, dq_full seen, dQ drained + barrier, reduce issued)
__device__ unsigned long long g_attn_bwd_trace[64 * 16];
__device__ __forceinline__ void bwd_trace(int step, int ev) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && step < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_attn_bwd_trace[step * 16 + ev] = t;
  }
}
#define BWD_TRACE(step, ev) bwd_trace(step, ev)
#else
#define BWD_TRACE(step, ev)
#endif

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
  int dbg;                      // bring-up switches (MERLIN_ATTN_BWD_DBG): 1 = skip the dQ TMA reduce, 2 = skip the dQ drain,
                                // 4 = skip P/dS math (constants)
};

template <int HD>
struct AttnBwdCfg {
  static constexpr int kTileBytes = kBwdBlock * HD * 2;         // [128 rows x HD] bf16: K, V, Q, dO
  static constexpr int kPBytes = kBwdBlock * kBwdBlock * 2;     // [128 x 128] bf16: P, dS
  static constexpr int kSmemTiles = kTileBytes * 5 + kPBytes * 2;   // K, V, Q x2, dO + P, dS
  static constexpr int kNumBars = 18;
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
  uint8_t* sq = sv + Cfg::kTileBytes;            // two stages
  uint8_t* sdo = sq + 2 * Cfg::kTileBytes;
  uint8_t* sp = sdo + Cfg::kTileBytes;
  uint8_t* sds = sp + Cfg::kPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemTiles);
  uint64_t* kv_full = bars;        // K_j, V_j landed (once)
  uint64_t* q_full = bars + 1;     // [2] Q_i landed in stage i & 1
  uint64_t* do_full = bars + 3;    // dO_i landed
  uint64_t* q_empty = bars + 11;   // [2] S and dK MMAs of step i done: stage i & 1 reusable
  uint64_t* do_empty = bars + 4;   // dP and dV MMAs of step i done: dO_i reusable
  uint64_t* s_full = bars + 5;     // S in TMEM
  uint64_t* dp_full = bars + 6;    // dP in TMEM
  uint64_t* p_full = bars + 7;     // P in shared memory (256 arrivals)
  uint64_t* ds_full = bars + 8;    // dS in shared memory (256 arrivals)
  uint64_t* dq_full = bars + 9;    // dQ of step i in TMEM; S, dP, dV MMAs complete (P staging reusable; dS: see q_empty)
  uint64_t* dq_read = bars + 10;   // dQ drained from TMEM (256 arrivals): dP columns reusable
  uint64_t* dq_staged = bars + 13; // dQ tile sits in the fp32 staging (256 arrivals): the reducer warp may issue the TMA reduce
  uint64_t* stg_free = bars + 14;  // [4] the reduce has finished reading staging chunk c (0-1 alias P, 2-3 alias dS;
                                   //     HD = 64: chunks 0-1 only, both inside P)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int j = blockIdx.y;   // key tile; grid = (heads, key tiles, batch): with causal masking tile 0 walks the most
  const int head = blockIdx.x;   // query tiles, and heads vary fastest: CTAs are dispatched longest-first across heads
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
    mbar_init(q_full + 1, 1);
    mbar_init(do_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(q_empty + 1, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_full, kBwdSoftmaxThreads);
    mbar_init(ds_full, kBwdSoftmaxThreads);
    mbar_init(dq_full, 1);
    mbar_init(dq_read, kBwdSoftmaxThreads);
    mbar_init(dq_staged, kBwdSoftmaxThreads);
    for (int c = 0; c < 4; ++c) mbar_init(stg_free + c, 1);
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
        const int st = s & 1;
        mbar_wait(q_empty + st, ((s >> 1) & 1) ^ 1);   // stage freed by dK of step s - 2
        mbar_expect_tx(q_full + st, Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_load_3d(&tm_qkv, q_full + st, sq + st * Cfg::kTileBytes + a * kAtomBytes, p.q_col0 + head * HD + a * 64, q0,
                      batch, kEvictLast);
        mbar_wait(do_empty, (s & 1) ^ 1);              // after dV of the previous step
        mbar_expect_tx(do_full, Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_load_3d(&tm_do, do_full, sdo + a * kAtomBytes, head * HD + a * 64, q0, batch, kEvictLast);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    // (a second issuing thread for S / dP was measured: 547 us vs 517 us per call at S = 4096 — the softmax
    //  warps, not the issue rate, bound the step)
    if (elect_one() && n_steps > 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, 128, 0, 0);   // S, dP : A, B K-major
      constexpr uint32_t idesc_mm = umma_idesc_bf16(128, HD, 1, 1);    // dV, dK: A, B MN-major
      constexpr uint32_t idesc_km = umma_idesc_bf16(128, HD, 0, 1);    // dQ    : A K-major, B MN-major
      const uint32_t a_k = smem_u32(sk), a_v = smem_u32(sv), a_q = smem_u32(sq), a_do = smem_u32(sdo),
                     a_p = smem_u32(sp), a_ds = smem_u32(sds);
      mbar_wait(kv_full, 0);
      auto issue_s = [&](int step) {   // S = Q Kᵀ of `step` into the S columns
        const uint32_t aq = a_q + (step & 1) * Cfg::kTileBytes;
        mbar_wait(q_full + (step & 1), (step >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(t_s, umma_desc_k_sw128(aq + off), umma_desc_k_sw128(a_k + off), idesc_kk, kk != 0);
        }
        umma_commit(s_full);
      };
      issue_s(0);
      for (int s = 0; s < n_steps; ++s) {
        const uint32_t ph = s & 1;
        const uint32_t aq = a_q + (s & 1) * Cfg::kTileBytes;
        // ---- dP = dO Vᵀ  (runs while the softmax warps turn S into P; the dP columns held dQ of the previous
        //      step until the softmax warps drained it)
        mbar_wait(do_full, ph);
        mbar_wait(dq_read, ph ^ 1);
        BWD_TRACE(s, 0);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(t_dp, umma_desc_k_sw128(a_do + off), umma_desc_k_sw128(a_v + off), idesc_kk, kk != 0);
        }
        umma_commit(dp_full);
        BWD_TRACE(s, 1);
        // ---- dV += Pᵀ dO  (runs while the softmax warps turn P, dP into dS)
        mbar_wait(p_full, ph);
        BWD_TRACE(s, 2);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBwdBlock / 16; ++kk) {
          // reduction over 16 query rows per instruction: every operand advances 16 rows = 2 KB
          const uint32_t roff = kk * 2048;
          umma_ss(t_dv, umma_desc_mn_sw128(a_p + roff, kAtomBytes), umma_desc_mn_sw128(a_do + roff, kAtomBytes),
                  idesc_mm, (s | kk) != 0);
        }
        umma_commit(do_empty);
        BWD_TRACE(s, 3);
        // ---- dQ = dS K into the dP columns (dP was consumed before dS was published) FIRST: the softmax warps
        //      are waiting to drain it; then dK += dSᵀ Q; then S of the NEXT step (P of this step left the S
        //      columns long ago, Q(s+1) sits in the other stage) — it is not needed before the drain is over
        mbar_wait(ds_full, ph);
        BWD_TRACE(s, 5);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kBwdBlock / 16; ++kk) {
          // reduction over 16 keys: A advances inside the key-major rows, B by 16 key rows
          const uint32_t aoff = (kk >> 2) * kAtomBytes + (kk & 3) * 32;
          umma_ss(t_dp, umma_desc_k_sw128(a_ds + aoff), umma_desc_mn_sw128(a_k + kk * 2048, kAtomBytes), idesc_km,
                  kk != 0);
        }
        umma_commit(dq_full);
#pragma unroll
        for (int kk = 0; kk < kBwdBlock / 16; ++kk) {
          const uint32_t roff = kk * 2048;
          umma_ss(t_dk, umma_desc_mn_sw128(a_ds + roff, kAtomBytes), umma_desc_mn_sw128(aq + roff, kAtomBytes),
                  idesc_mm, (s | kk) != 0);
        }
        umma_commit(q_empty + (s & 1));     // also: dS is no longer read (staging chunks 2-3 may be overwritten)
        BWD_TRACE(s, 6);
        if (s + 1 < n_steps) issue_s(s + 1);
        BWD_TRACE(s, 4);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================================== dQ reducer: one TMA reduce-add per [128 x 32 fp32] box
    // (kept off the softmax warps: issuing the bulk operations and waiting for their reads costs ~1 us per step)
    if (elect_one() && n_steps > 0) {
      for (int s = 0; s < n_steps; ++s) {
        const int i = i_begin + s;
        mbar_wait(dq_staged, s & 1);
        // one bulk group per 16 KB chunk, so that each chunk of the staging is released as soon as it was read
#pragma unroll
        for (int ch = 0; ch < HD / 32; ++ch) {
          if (!(p.dbg & 1))
            tma_reduce_add_2d(&tm_dq, sp + ch * kAtomBytes, head * HD + ch * 32, batch * p.S + i * kBwdBlock);
          tma_store_commit();
        }
        if (HD == 128) {
          tma_store_wait_read<3>();
          mbar_arrive(stg_free + 0);
          tma_store_wait_read<2>();
          mbar_arrive(stg_free + 1);
          tma_store_wait_read<1>();
          mbar_arrive(stg_free + 2);
          tma_store_wait_read<0>();
          mbar_arrive(stg_free + 3);
        } else {
          tma_store_wait_read<1>();
          mbar_arrive(stg_free + 0);
          tma_store_wait_read<0>();
          mbar_arrive(stg_free + 1);
          mbar_arrive(stg_free + 2);
          mbar_arrive(stg_free + 3);
        }
      }
      tma_store_wait_all<0>();
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
    // per-row statistics are fetched one step ahead: in steady state S of the next step is already waiting, so a
    // load issued at the top of the step would sit on the critical path
    const size_t stat_base = (static_cast<size_t>(batch) * p.H + head) * p.S;
    auto load_stats = [&](int step, float& l2, float& dl) {
      const int qi = (i_begin + step) * kBwdBlock + row;
      const bool ok = step < n_steps && qi < len;
      l2 = ok ? __ldg(p.lse + stat_base + qi) * log2e : 0.f;
      dl = ok ? __ldg(p.delta + stat_base + qi) : 0.f;
    };
    float lse2_next, dlt_next;
    load_stats(0, lse2_next, dlt_next);
    for (int s = 0; s < n_steps; ++s) {
      const uint32_t ph = s & 1;
      const int i = i_begin + s;
      const int q_idx = i * kBwdBlock + row;
      const bool q_valid = q_idx < len;
      const float lse2 = lse2_next, dlt = dlt_next;
      load_stats(s + 1, lse2_next, dlt_next);
      // every (query, key) pair of the tile is live: strictly below the diagonal and inside the sequence
      const bool interior = (!kCausal || i > j) && (i + 1) * kBwdBlock <= len && kv0 + kBwdBlock <= len;
      float pv[64];
      // ---- P = exp2(S c − lse)
      mbar_wait(s_full, ph);
      if (leader) BWD_TRACE(s, 8);
      tc_fence_after();
      // the previous step's dQ reduction must have finished READING the staging bytes that alias P
      // this thread's half of P aliases staging chunk `half` of the previous step's dQ tile
      mbar_wait(stg_free + half, ph ^ 1);
      if (leader) BWD_TRACE(s, 14);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;               // 32-key chunk
        uint32_t sv_[32];
        tmem_ld32(t_s + lane_sel + ch * 32, sv_);
        tmem_ld_wait();
        uint32_t pw[16];
        if (interior) {   // no causal / length masking inside this tile: the common case
#pragma unroll
          for (int c = 0; c < 32; c += 2) {   // every other exponential on the FMA pipe (MUFU does 16 per clock)
            pv[c2 * 32 + c] = fast_exp2(fmaf(__uint_as_float(sv_[c]), p.scale_log2e, -lse2));
            pv[c2 * 32 + c + 1] = exp2_poly3(fmaf(__uint_as_float(sv_[c + 1]), p.scale_log2e, -lse2));
            pw[c >> 1] = pack_bf16(pv[c2 * 32 + c], pv[c2 * 32 + c + 1]);
          }
        } else {
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
        }
        // 32 keys = 64 bytes = four 16-byte chunks of this row inside atom (ch >> 1)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = (ch >> 1) * kAtomBytes + sw128_offset(row, (ch & 1) * 4 + q4);
          *reinterpret_cast<uint4*>(sp + off) = make_uint4(pw[q4 * 4], pw[q4 * 4 + 1], pw[q4 * 4 + 2], pw[q4 * 4 + 3]);
        }
      }
      if (leader) BWD_TRACE(s, 15);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (leader) BWD_TRACE(s, 9);
      // ---- dS = P ∘ (dP − δ) · scale
      mbar_wait(dp_full, ph);
      if (leader) BWD_TRACE(s, 10);
      tc_fence_after();
      mbar_wait(stg_free + 2 + half, ph ^ 1);   // ... and its half of dS aliases chunk 2 + half
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
      if (leader) BWD_TRACE(s, 11);

      // ---- dQ_i tile: TMEM -> fp32 128B-swizzled staging (aliasing P|dS, which the MMAs have finished
      // reading) -> ONE TMA reduce-add per [128 rows x 32 floats] box into the fp32 dQ accumulator
      mbar_wait(dq_full, ph);
      if (leader) BWD_TRACE(s, 12);
      tc_fence_after();
      if (HD == 128 && half == 1) mbar_wait(q_empty + (s & 1), (s >> 1) & 1);   // dK has finished reading dS
#pragma unroll
      for (int c2 = 0; c2 < HD / 64; ++c2) {
        if (p.dbg & 2) break;
        const int ch = half * (HD / 64) + c2;       // 32-float column chunk of dQ
        uint32_t v[32];
        tmem_ld32(t_dp + lane_sel + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<uint4*>(sp + ch * kAtomBytes + sw128_offset(row, c16)) =
              make_uint4(v[c16 * 4], v[c16 * 4 + 1], v[c16 * 4 + 2], v[c16 * 4 + 3]);
      }
      tc_fence_before();
      mbar_arrive(dq_read);
      fence_proxy_async_smem();
      mbar_arrive(dq_staged);
      if (leader) BWD_TRACE(s, 13);
    }

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
