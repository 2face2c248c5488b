// merlin_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M, Nout] = epilogue( A[M, K] · B[N, K]^T )        bf16 in, fp32 accumulate in TMEM
//
// A and B are K-major (row-major activations, nn.Linear weights [out, in]), moved by TMA into
// 128B-swizzled shared-memory stages and multiplied by tcgen05.mma (UMMA 128x256x16, or
// 256x256x16 across a CTA pair with cta_group::2).  Two 256-column TMEM accumulator stages let
// the epilogue of tile i overlap the main loop of tile i+1.
//
// Warp roles (256 threads):  0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator,
//                            3 = residual-tile loader, 4..7 = epilogue (one TMEM lane per thread).
#pragma once
#include "ptx.cuh"

namespace mb {

enum GemmEpilogue : int {
  EPI_STORE = 0,      // C = acc (+ bias)
  EPI_RESIDUAL = 1,   // C = acc (+ bias) + R          (R may alias C)
  EPI_QUICKGELU = 2,  // C = v * sigmoid(1.702 v),  v = acc + bias      (CLIP quick_gelu)
  EPI_SWIGLU = 3,     // B rows interleaved [128 gate | 128 up] per 256-row tile; C[M, N/2] = silu(g)*u
  EPI_ROPE = 4,       // half-split rotary embedding on columns < rope_cols (head_dim 128), else store
};

struct GemmParams {
  int M, N, K;               // N = number of B rows (accumulator columns)
  int epilogue;              // GemmEpilogue
  const __nv_bfloat16* bias; // [N] or nullptr
  const float* rope_cos;     // [max_pos, 64] fp32
  const float* rope_sin;     // [max_pos, 64] fp32
  const int* positions;      // [M] position id of every row, or nullptr -> row % seq_len
  int rope_cols;             // columns [0, rope_cols) are rotated (q and k blocks of a packed qkv)
  int seq_len;
  // KV-cache append fused into the epilogue of the q|k|v GEMM (kv_col0 < 0: off): output columns
  // [kv_col0, kv_col0 + 2 kv_hidden) are the k and v blocks; every staged [128 tokens x 64 columns] chunk of
  // them is ALSO stored through tm_kv into the cache [B, cap/32, H, 32, 256] (k at columns [0,128), v at
  // [128,256) of a (head, token) row).  Needs kv_seq % 128 == 0 or one sample (a tile may not straddle samples).
  int kv_col0, kv_hidden, kv_seq;
};

constexpr int kGemmThreads = 256;
constexpr int kBlockM = 128;   // rows per CTA
constexpr int kBlockN = 256;   // accumulator columns per tile
constexpr int kBlockK = 64;    // one 128-byte swizzle atom of bf16
constexpr int kEpiCols = 64;   // columns per epilogue chunk / TMA store box
constexpr int kEpiChunkBytes = kBlockM * kEpiCols * 2;  // 16 KB

template <int kCtaGroup>
struct GemmCfg {
  static constexpr int kLoadN = kBlockN / kCtaGroup;  // B rows each CTA loads
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kLoadN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kCtaGroup == 2) ? 5 : 4;
  static constexpr int kEpiBufs = (kCtaGroup == 2) ? 4 : 2;
  static constexpr int kSmemTiles = kStages * kStageBytes + kEpiBufs * kEpiChunkBytes;
  static constexpr int kNumBars = 2 * kStages + 4 + 2 * kEpiBufs;
  static constexpr int kSmemBytes = kSmemTiles + kNumBars * 8 + 16 + 1024;  // +1024 alignment slack
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst,
                                                 int c0, int c1, uint64_t hint) {
  // complete_tx lands on the barrier of the even-ranked (leader) CTA of the pair.
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1),
      "l"(hint)
      : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols)
               : "memory");
}

// tile -> (m block, n block).  m runs fastest so that the B (weight) tile of one n block is
// shared through L2 by the CTAs that run concurrently.
struct TileCoord {
  int m_blk, n_blk;
};
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_m) {
  return TileCoord{t % tiles_m, t / tiles_m};
}

// kAMN / kBMN: the operand is stored "MN-major", i.e. as a row-major [K, M] (resp. [K, N]) matrix whose
// contiguous dimension is M (N) — what backward needs: dX = dY·W reads W[N_red, K_out] as B, and
// dW = dYᵀ·X reads dY[M_red, N] as A and X[M_red, K] as B.  Such tiles are staged as 64-element x 64-row
// 128B-swizzled boxes and described to tcgen05.mma with MN-major descriptors.
template <int kCtaGroup, bool kAMN = false, bool kBMN = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_r,
                 const __grid_constant__ CUtensorMap tm_kv, const GemmParams p) {
  using Cfg = GemmCfg<kCtaGroup>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kEpiBufs = Cfg::kEpiBufs;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_ab = smem;                                       // kStages x [A | B]
  uint8_t* smem_epi = smem + kStages * Cfg::kStageBytes;         // kEpiBufs x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemTiles);
  uint64_t* full_bar = bars;                   // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;    // [2]        MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;        // [2]        epilogue -> MMA
  uint64_t* res_full = tmem_empty + 2;         // [kEpiBufs] residual loader -> epilogue
  uint64_t* res_empty = res_full + kEpiBufs;   // [kEpiBufs] epilogue -> residual loader
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_empty + kEpiBufs);

  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0;
  const bool is_leader = cta_rank == 0;

  const int rows_per_tile = kBlockM * kCtaGroup;
  const int tiles_m = (p.M + rows_per_tile - 1) / rows_per_tile;
  const int tiles_n = (p.N + kBlockN - 1) / kBlockN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int worker = blockIdx.x / kCtaGroup;
  const int num_workers = gridDim.x / kCtaGroup;
  const bool has_residual = p.epilogue == EPI_RESIDUAL;
  const int out_chunks = (p.epilogue == EPI_SWIGLU) ? 2 : 4;  // 64-column output chunks per tile
  const int n_out = (p.epilogue == EPI_SWIGLU) ? p.N / 2 : p.N;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
    if (has_residual) tma_prefetch_desc(&tm_r);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4 * kCtaGroup);  // one arrive per epilogue warp of every CTA
    }
    for (int i = 0; i < kEpiBufs; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kCtaGroup == 2) tmem_alloc_pair<512>(tmem_ptr_smem);
    else tmem_alloc<512>(tmem_ptr_smem);
  }
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        const TileCoord tc = tile_coord(t, tiles_m);
        const int m0 = tc.m_blk * rows_per_tile + static_cast<int>(cta_rank) * kBlockM;
        const int n0 = tc.n_blk * kBlockN + static_cast<int>(cta_rank) * Cfg::kLoadN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_ab + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1, uint64_t hint) {
            if constexpr (kCtaGroup == 2) tma_load_2d_pair(tm, &full_bar[stage], dst, c0, c1, hint);
            else tma_load_2d(tm, &full_bar[stage], dst, c0, c1, hint);
          };
          if constexpr (kCtaGroup == 2) {
            if (is_leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          }
          if constexpr (kAMN) {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i)
              load(&tm_a, sa + i * 8192, m0 + i * 64, kb * kBlockK, kEvictNormal);
          } else {
            load(&tm_a, sa, kb * kBlockK, m0, kEvictNormal);
          }
          if constexpr (kBMN) {
#pragma unroll
            for (int i = 0; i < Cfg::kLoadN / 64; ++i)
              load(&tm_b, sb + i * 8192, n0 + i * 64, kb * kBlockK, kEvictLast);
          } else {
            load(&tm_b, sb, kb * kBlockK, n0, kEvictLast);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================= MMA issuer
    if (is_leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * kCtaGroup, kBlockN, kAMN ? 1 : 0, kBMN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBlockN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_ab + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // K-major: advance 16 bf16 = 32 bytes inside the swizzle atom.  MN-major: 16 k-rows of
            // 128 bytes = 2048 bytes; 64-element MN groups are 8 KB apart (LBO).
            const uint64_t da = kAMN ? umma_desc_mn_sw128(sa + k * 2048, 8192) : umma_desc_k_sw128(sa + k * 32);
            const uint64_t db = kBMN ? umma_desc_mn_sw128(sb + k * 2048, 8192) : umma_desc_k_sw128(sb + k * 32);
            if constexpr (kCtaGroup == 2) umma_ss_pair(tmem_d, da, db, idesc, (kb | k) != 0);
            else umma_ss(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          if constexpr (kCtaGroup == 2) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (kCtaGroup == 2) umma_commit_pair(&tmem_full[acc]);
        else umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================================ residual loader
    if (has_residual && elect_one()) {
      uint32_t it = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        const TileCoord tc = tile_coord(t, tiles_m);
        const int m0 = tc.m_blk * rows_per_tile + static_cast<int>(cta_rank) * kBlockM;
        const int n0 = tc.n_blk * kBlockN;
        for (int c = 0; c < 4; ++c, ++it) {
          const uint32_t buf = it % kEpiBufs;
          const uint32_t ph = (it / kEpiBufs) & 1;
          mbar_wait(&res_empty[buf], ph ^ 1);
          mbar_expect_tx(&res_full[buf], kEpiChunkBytes);
          tma_load_2d(&tm_r, &res_full[buf], smem_epi + buf * kEpiChunkBytes, n0 + c * kEpiCols, m0,
                      kEvictFirst);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ========================================================= epilogue
    const int ewarp = warp - 4;                       // TMEM lanes [32*ewarp, 32*ewarp+32)
    const int row = ewarp * 32 + static_cast<int>(lane_id());
    const bool issuer = threadIdx.x == 128;           // first epilogue thread issues TMA stores
    const uint32_t lane_sel = static_cast<uint32_t>(ewarp * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t it = 0;
    for (int t = worker; t < num_tiles; t += num_workers) {
      const TileCoord tc = tile_coord(t, tiles_m);
      const int m0 = tc.m_blk * rows_per_tile + static_cast<int>(cta_rank) * kBlockM;
      const int n0 = tc.n_blk * kBlockN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + lane_sel + static_cast<uint32_t>(acc * kBlockN);

      int pos = 0;
      const bool rope_tile = p.epilogue == EPI_ROPE && n0 < p.rope_cols;
      if (rope_tile) {
        const int grow = m0 + row;
        pos = (grow < p.M) ? (p.positions ? p.positions[grow] : grow % p.seq_len) : 0;
      }

      for (int c = 0; c < out_chunks; ++c, ++it) {
        const uint32_t buf = it % kEpiBufs;
        uint8_t* sbuf = smem_epi + buf * kEpiChunkBytes;
        if (has_residual) {
          mbar_wait(&res_full[buf], (it / kEpiBufs) & 1);
        } else {
          if (issuer) tma_store_wait_read<kEpiBufs - 1>();
          named_bar_sync(1, 128);
        }

        uint32_t v[64];  // fp32 bit patterns of this thread's 64 output values
        if (p.epilogue == EPI_SWIGLU) {
          uint32_t g[32], u[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(tacc + c * 64 + h * 32, g);
            tmem_ld32(tacc + 128 + c * 64 + h * 32, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float gv = __uint_as_float(g[j]);
              const float uv = __uint_as_float(u[j]);
              v[h * 32 + j] = __float_as_uint(gv / (1.0f + __expf(-gv)) * uv);
            }
          }
        } else if (rope_tile) {
          // chunk pairs (0,1) and (2,3) are the two 64-wide halves of a 128-wide head
          const int head_base = (c >> 1) * 128;
          const bool upper = (c & 1) != 0;
          const float* cosp = p.rope_cos + static_cast<size_t>(pos) * 64;
          const float* sinp = p.rope_sin + static_cast<size_t>(pos) * 64;
          uint32_t lo[32], hi[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_ld32(tacc + head_base + h * 32, lo);
            tmem_ld32(tacc + head_base + 64 + h * 32, hi);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              const float4 cs = __ldg(reinterpret_cast<const float4*>(cosp + h * 32 + j4));
              const float4 sn = __ldg(reinterpret_cast<const float4*>(sinp + h * 32 + j4));
              const float cc[4] = {cs.x, cs.y, cs.z, cs.w};
              const float ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = __uint_as_float(lo[j4 + e]);
                const float b = __uint_as_float(hi[j4 + e]);
                // rotate_half: out_lo = a*cos - b*sin ; out_hi = b*cos + a*sin
                const float o = upper ? (b * cc[e] + a * ss[e]) : (a * cc[e] - b * ss[e]);
                v[h * 32 + j4 + e] = __float_as_uint(o);
              }
            }
          }
        } else {
          uint32_t r0[32], r1[32];
          tmem_ld32(tacc + c * 64, r0);
          tmem_ld32(tacc + c * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = r0[j];
            v[32 + j] = r1[j];
          }
        }

        if (c == out_chunks - 1) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane_id() == 0) {
            if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty[acc], 0);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }

        if (p.bias != nullptr && p.epilogue != EPI_SWIGLU) {
          const int col0 = n0 + c * kEpiCols;
#pragma unroll
          for (int j8 = 0; j8 < 64; j8 += 8) {
            uint4 bq = make_uint4(0, 0, 0, 0);
            if (col0 + j8 + 8 <= p.N) {
              bq = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + j8));
            } else {
              uint32_t w[4] = {0, 0, 0, 0};
              for (int e = 0; e < 8; ++e)
                if (col0 + j8 + e < p.N) {
                  const uint32_t b = __bfloat16_as_ushort(p.bias[col0 + j8 + e]);
                  w[e >> 1] |= b << ((e & 1) * 16);
                }
              bq = make_uint4(w[0], w[1], w[2], w[3]);
            }
            const uint32_t w[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[j8 + 2 * e] = __float_as_uint(__uint_as_float(v[j8 + 2 * e]) + bf16_lo(w[e]));
              v[j8 + 2 * e + 1] = __float_as_uint(__uint_as_float(v[j8 + 2 * e + 1]) + bf16_hi(w[e]));
            }
          }
        }
        if (p.epilogue == EPI_QUICKGELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float x = __uint_as_float(v[j]);
            v[j] = __float_as_uint(x / (1.0f + __expf(-1.702f * x)));
          }
        }

        // registers -> bf16 -> 128B-swizzled staging tile (row = this thread's TMEM lane)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4* dst = reinterpret_cast<uint4*>(sbuf + sw128_offset(row, ch));
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[ch * 8 + e]);
          if (has_residual) {
            const uint4 r = *dst;
            f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x);
            f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
            f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z);
            f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
          }
          *dst = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                            pack_bf16(f[6], f[7]));
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          const int out_col = (p.epilogue == EPI_SWIGLU) ? tc.n_blk * 128 + c * kEpiCols
                                                          : n0 + c * kEpiCols;
          // boxes that start outside the tensor carry no data; partial boxes are clipped by TMA
          if (out_col < n_out && m0 < p.M) {
            tma_store_2d(&tm_c, sbuf, out_col, m0);
            if (p.kv_col0 >= 0 && out_col >= p.kv_col0) {
              // the same staged chunk, second destination: rows = 128 consecutive tokens = 4 cache blocks of 32,
              // columns = one 64-wide half of head `e / 128` of the k (which = 0) or v (which = 1) block
              const int e = out_col - p.kv_col0;
              const int which = e / p.kv_hidden, eh = e % p.kv_hidden;
              const int sample = m0 / p.kv_seq, t0 = m0 % p.kv_seq;
              tma_store_5d(&tm_kv, sbuf, which * 128 + (eh & 127), 0, eh >> 7, t0 >> 5, sample);
            }
          }
          tma_store_commit();
          if (has_residual) {
            // every store but the newest has finished reading smem: release the previous buffer
            tma_store_wait_read<1>();
            if (it > 0) mbar_arrive(&res_empty[(it - 1) % kEpiBufs]);
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if constexpr (kCtaGroup == 2) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mb
