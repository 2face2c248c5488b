// merlin_b200 — host launcher for the flash-attention forward kernel (attn.cuh).
#include "attn.cuh"
#include "attn2.cuh"
#include "attn_bwd.cuh"
#include "host.cuh"

#include <mutex>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace mb {

template <int HD, bool kCausal>
static int launch_attn(cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& to,
                       const AttnParams& p) {
  using Cfg = AttnCfg<HD>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] { attr_rc = cudaFuncSetAttribute(attn_fwd_kernel<HD, kCausal>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
  MB_CHECK_CUDA(attr_rc);
  const int nq = (p.S + kAttnBlock - 1) / kAttnBlock;
  dim3 grid(static_cast<unsigned>(p.H), static_cast<unsigned>(nq), static_cast<unsigned>(p.B));
  attn_fwd_kernel<HD, kCausal><<<grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tq, to, p);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

template <int HD, bool kCausal>
static int launch_attn2(cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& to,
                        const AttnParams& p) {
  using Cfg = Attn2Cfg<HD>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] { attr_rc = cudaFuncSetAttribute(attn_fwd2_kernel<HD, kCausal>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
  MB_CHECK_CUDA(attr_rc);
  const int nq = (p.S + kAttnBlock - 1) / kAttnBlock;
  dim3 grid(static_cast<unsigned>(p.H), static_cast<unsigned>((nq + 1) / 2), static_cast<unsigned>(p.B));
  attn_fwd2_kernel<HD, kCausal><<<grid, kAttn2Threads, Cfg::kSmemBytes, stream>>>(tq, to, p);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

// 1 = one query tile per CTA (attn.cuh), 2 = two query tiles per CTA, ping-pong softmax, P in TMEM (attn2.cuh,
// 20 % faster at S = 4096).  Default: v2 for the long causal head_dim-128 sequences of the decoder prefill /
// training forward (where it was profiled), v1 everywhere else; MERLIN_ATTN_V=1|2 forces one of them.
static int attn_version(int S, int head_dim, int causal) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MERLIN_ATTN_V");
    forced = (e && (e[0] == '1' || e[0] == '2')) ? (e[0] - '0') : 0;
  }
  if (forced) return forced;
  return (head_dim == 128 && causal && S >= 1024) ? 2 : 1;
}

}  // namespace mb

extern "C" int merlin_flash_attn_fwd(void* stream, const void* qkv, int64_t ld_qkv, void* out,
                                     int64_t ld_out, float* lse, const int* seqlens, int B, int S,
                                     int H, int head_dim, int causal, float softmax_scale) {
  using namespace mb;
  MB_CHECK_ARG(qkv && out, "flash_attn_fwd: null tensor");
  MB_CHECK_ARG(B > 0 && S > 0 && H > 0, "flash_attn_fwd: empty problem B=%d S=%d H=%d", B, S, H);
  MB_CHECK_ARG(head_dim == 64 || head_dim == 128, "flash_attn_fwd: head_dim %d not in {64,128}",
               head_dim);
  MB_CHECK_ARG(ld_qkv >= 3 * static_cast<int64_t>(H) * head_dim && ld_out >= static_cast<int64_t>(H) * head_dim,
               "flash_attn_fwd: leading dimensions too small");
  AttnParams p = {};
  p.B = B;
  p.S = S;
  p.H = H;
  p.q_col0 = 0;
  p.k_col0 = H * head_dim;
  p.v_col0 = 2 * H * head_dim;
  const float scale = softmax_scale > 0.f ? softmax_scale : 1.0f / std::sqrt(static_cast<float>(head_dim));
  p.scale_log2e = scale * 1.4426950408889634f;
  p.seqlens = seqlens;
  p.lse = lse;

  CUtensorMap tq, to;
  int rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3) * H * head_dim, static_cast<uint64_t>(S),
                        static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(ld_qkv) * 2, static_cast<uint64_t>(ld_qkv) * 2 * S};
    uint32_t box[3] = {64, kAttnBlock, 1};
    if ((rc = make_tmap_bf16(&tq, qkv, 3, dims, str, box)) != MERLIN_OK) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(H) * head_dim, static_cast<uint64_t>(S),
                        static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(ld_out) * 2, static_cast<uint64_t>(ld_out) * 2 * S};
    uint32_t box[3] = {64, kAttnBlock, 1};
    if ((rc = make_tmap_bf16(&to, out, 3, dims, str, box)) != MERLIN_OK) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (attn_version(S, head_dim, causal) == 2) {
    if (head_dim == 128)
      return causal ? launch_attn2<128, true>(st, tq, to, p) : launch_attn2<128, false>(st, tq, to, p);
    return causal ? launch_attn2<64, true>(st, tq, to, p) : launch_attn2<64, false>(st, tq, to, p);
  }
  if (head_dim == 128)
    return causal ? launch_attn<128, true>(st, tq, to, p) : launch_attn<128, false>(st, tq, to, p);
  return causal ? launch_attn<64, true>(st, tq, to, p) : launch_attn<64, false>(st, tq, to, p);
}

extern "C" int merlin_attn_delta(void* stream, const void* out, const void* dout, float* delta, int B, int S,
                                 int H, int head_dim, int64_t ld);

extern "C" size_t merlin_flash_attn_bwd_workspace_bytes(int B, int S, int H, int head_dim) {
  const size_t delta = (static_cast<size_t>(B) * H * S + 63) / 64 * 64;  // keep the dq accumulator 256-byte aligned
  return (delta + static_cast<size_t>(B) * S * H * head_dim) * sizeof(float);
}

extern "C" int merlin_flash_attn_bwd(void* stream, const void* qkv, int64_t ld_qkv, const void* out,
                                     const void* dout, int64_t ld_out, const float* lse, const int* seqlens,
                                     void* dqkv, int64_t ld_dqkv, void* workspace, int B, int S, int H,
                                     int head_dim, int causal, float softmax_scale) {
  using namespace mb;
  MB_CHECK_ARG(qkv && out && dout && lse && dqkv && workspace, "flash_attn_bwd: null tensor");
  MB_CHECK_ARG(head_dim == 128 || head_dim == 64, "flash_attn_bwd: head_dim %d not in {64,128}", head_dim);
  MB_CHECK_ARG(B > 0 && S > 0 && H > 0, "flash_attn_bwd: empty problem");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* delta = static_cast<float*>(workspace);
  float* dq_acc = delta + (static_cast<size_t>(B) * H * S + 63) / 64 * 64;
  int rc = merlin_attn_delta(stream, out, dout, delta, B, S, H, head_dim, ld_out);
  if (rc != MERLIN_OK) return rc;
  MB_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(B) * S * H * head_dim * sizeof(float), st));

  AttnBwdParams p = {};
  p.B = B;
  p.S = S;
  p.H = H;
  p.q_col0 = 0;
  p.k_col0 = H * head_dim;
  p.v_col0 = 2 * H * head_dim;
  p.scale = softmax_scale > 0.f ? softmax_scale : 1.0f / std::sqrt(static_cast<float>(head_dim));
  p.scale_log2e = p.scale * 1.4426950408889634f;
  p.seqlens = seqlens;
  p.lse = lse;
  p.delta = delta;
  p.dq_accum = dq_acc;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.ld_dqkv = ld_dqkv;
  {
    const char* e = getenv("MERLIN_ATTN_BWD_DBG");
    p.dbg = e ? atoi(e) : 0;
  }

  CUtensorMap tq, tdo;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(3) * H * head_dim, static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(ld_qkv) * 2, static_cast<uint64_t>(ld_qkv) * 2 * S};
    uint32_t box[3] = {64, kBwdBlock, 1};
    if ((rc = make_tmap_bf16(&tq, qkv, 3, dims, str, box)) != MERLIN_OK) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(H) * head_dim, static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(ld_out) * 2, static_cast<uint64_t>(ld_out) * 2 * S};
    uint32_t box[3] = {64, kBwdBlock, 1};
    if ((rc = make_tmap_bf16(&tdo, dout, 3, dims, str, box)) != MERLIN_OK) return rc;
  }
  CUtensorMap tdq;
  {
    // fp32 dQ accumulator [B*S, H*hd].  A ragged last query tile of sample b overlaps the first rows of
    // sample b+1, but its rows >= S carry P = dS = 0, i.e. exactly zero dQ: the reduce-add of zeros is a
    // no-op there, and rows beyond B*S are clipped by the tensor map.
    uint64_t dims[2] = {static_cast<uint64_t>(H) * head_dim, static_cast<uint64_t>(B) * S};
    uint64_t str[1] = {static_cast<uint64_t>(H) * head_dim * sizeof(float)};
    uint32_t box[2] = {32, kBwdBlock};
    if ((rc = make_tmap_f32(&tdq, dq_acc, 2, dims, str, box)) != MERLIN_OK) return rc;
  }
  dim3 grid(static_cast<unsigned>(H), static_cast<unsigned>((S + kBwdBlock - 1) / kBwdBlock),
            static_cast<unsigned>(B));
  auto launch = [&](auto kernel, int smem) -> int {
    MB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kernel<<<grid, kBwdThreads, smem, st>>>(tq, tdo, tdq, p);
    MB_CHECK_CUDA(cudaGetLastError());
    return MERLIN_OK;
  };
  if (head_dim == 128)
    rc = causal ? launch(attn_bwd_kernel<128, true>, AttnBwdCfg<128>::kSmemBytes)
                : launch(attn_bwd_kernel<128, false>, AttnBwdCfg<128>::kSmemBytes);
  else
    rc = causal ? launch(attn_bwd_kernel<64, true>, AttnBwdCfg<64>::kSmemBytes)
                : launch(attn_bwd_kernel<64, false>, AttnBwdCfg<64>::kSmemBytes);
  if (rc != MERLIN_OK) return rc;
  const int64_t rows = static_cast<int64_t>(B) * S;
  const int width = H * head_dim;
  dq_to_bf16_kernel<<<static_cast<unsigned>(std::min<int64_t>((rows * (width / 8) + 255) / 256, 148 * 16)), 256, 0, st>>>(
      dq_acc, static_cast<__nv_bfloat16*>(dqkv), ld_dqkv, 0, width, rows);
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

#ifdef MERLIN_ATTN_TRACE
extern "C" int merlin_debug_attn_bwd_trace(unsigned long long* host_out) {
  MB_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, mb::g_attn_bwd_trace, sizeof(unsigned long long) * 64 * 16));
  return MERLIN_OK;
}
#endif
