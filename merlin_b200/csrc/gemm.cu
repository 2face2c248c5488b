// merlin_b200 — host launcher for the tcgen05 GEMM (see gemm.cuh) and shared host helpers.
#include "gemm.cuh"
#include "host.cuh"

#include <atomic>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace mb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tmap_typed(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapDataType dtype);

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_typed(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_typed(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

// Encoded tensor maps are cached per thread, keyed on everything that goes into the encoding (base pointer,
// dims, strides, box, dtype): a training step re-issues the same ~30 GEMM shapes over the same buffers for
// every layer and every step, and cuTensorMapEncodeTiled costs a driver call each time.  thread_local keeps
// the library re-entrant without a lock (SURVEY.md §8b threading).
struct TmapKey {
  uint64_t base;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  uint32_t rank_dtype;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");

static int make_tmap_encode(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapDataType dtype);

static int make_tmap_typed(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapDataType dtype) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.base = reinterpret_cast<uint64_t>(base);
  for (int i = 0; i < rank; ++i) {
    k.dims[i] = dims[i];
    k.box[i] = box[i];
    if (i + 1 < rank) k.strides[i] = strides_bytes[i];
  }
  k.rank_dtype = (static_cast<uint32_t>(rank) << 16) | static_cast<uint32_t>(dtype);
  auto it = cache.find(k);
  if (it != cache.end()) {
    *out = it->second;
    return MERLIN_OK;
  }
  int rc = make_tmap_encode(out, base, rank, dims, strides_bytes, box, dtype);
  if (rc != MERLIN_OK) return rc;
  if (cache.size() >= 8192) cache.clear();
  cache.emplace(k, *out);
  return MERLIN_OK;
}

static int make_tmap_encode(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapDataType dtype) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return MERLIN_E_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tensor base %p is not 16-byte aligned", base);
    return MERLIN_E_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) {
      set_error("tensor stride %llu bytes (dim %d) is not a multiple of 16",
                static_cast<unsigned long long>(gstr[i]), i + 1);
      return MERLIN_E_INVALID;
    }
  CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
              static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
              static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return MERLIN_E_CUDA;
  }
  return MERLIN_OK;
}

int sm_count() {
  static int n = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  });
  return n;
}

// SM budget of the persistent kernels (merlin_set_sm_budget): 0 = the whole chip.  The data-parallel step loop
// lowers it while gradient all-reduces are in flight, so the static tile schedule of the GEMM is laid over the
// SMs the collective's CTAs leave free instead of queueing CTAs behind them.
static std::atomic<int> g_sm_budget{0};

int sm_budget() {
  const int n = sm_count();
  const int b = g_sm_budget.load(std::memory_order_relaxed);
  return (b > 0 && b < n) ? b : n;
}

template <int kCtaGroup, bool kAMN, bool kBMN>
static int launch_gemm(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb,
                       const CUtensorMap& tc, const CUtensorMap& tr, const CUtensorMap& tkv, const GemmParams& p) {
  using Cfg = GemmCfg<kCtaGroup>;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(gemm_bf16_kernel<kCtaGroup, kAMN, kBMN>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  });
  MB_CHECK_CUDA(attr_rc);
  const int rows_per_tile = kBlockM * kCtaGroup;
  const int tiles = ((p.M + rows_per_tile - 1) / rows_per_tile) * ((p.N + kBlockN - 1) / kBlockN);
  int workers = sm_budget() / kCtaGroup;
  if (tiles < workers) workers = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(workers * kCtaGroup));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<kCtaGroup, kAMN, kBMN>, ta, tb, tc, tr, tkv, p));
  return MERLIN_OK;
}

int gemm_dispatch(cudaStream_t stream, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                  int64_t ldc, const void* R, int64_t ldr, const GemmParams& p, int cta_group,
                  bool a_mn = false, bool b_mn = false, const CUtensorMap* kv_map = nullptr) {
  MB_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  MB_CHECK_ARG(cta_group == 1 || cta_group == 2, "gemm: cta_group must be 1 or 2, got %d", cta_group);
  MB_CHECK_ARG(A && B && C, "gemm: null operand");
  if (p.epilogue == EPI_SWIGLU)
    MB_CHECK_ARG(p.N % 256 == 0, "gemm: SwiGLU epilogue needs N %% 256 == 0 (interleaved gate/up), N=%d", p.N);
  if (p.epilogue == EPI_ROPE)
    MB_CHECK_ARG(p.rope_cos && p.rope_sin && p.rope_cols % 256 == 0 && p.seq_len > 0,
                 "gemm: RoPE epilogue needs cos/sin tables, rope_cols %% 256 == 0, seq_len > 0");
  if (p.epilogue == EPI_RESIDUAL) MB_CHECK_ARG(R != nullptr, "gemm: residual epilogue needs R");
  const int n_out = (p.epilogue == EPI_SWIGLU) ? p.N / 2 : p.N;

  CUtensorMap ta, tb, tc, tr;
  int rc;
  {
    // K-major: rows of A are M, inner dim K.  MN-major: stored [K, M], inner dim M, 64x64 boxes.
    uint64_t dims[2] = {static_cast<uint64_t>(a_mn ? p.M : p.K), static_cast<uint64_t>(a_mn ? p.K : p.M)};
    uint64_t str[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {kBlockK, a_mn ? 64u : static_cast<uint32_t>(kBlockM)};
    if ((rc = make_tmap_bf16(&ta, A, 2, dims, str, box)) != MERLIN_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(b_mn ? p.N : p.K), static_cast<uint64_t>(b_mn ? p.K : p.N)};
    uint64_t str[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {kBlockK, b_mn ? 64u : static_cast<uint32_t>(kBlockN / cta_group)};
    if ((rc = make_tmap_bf16(&tb, B, 2, dims, str, box)) != MERLIN_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(p.M)};
    uint64_t str[1] = {static_cast<uint64_t>(ldc) * 2};
    uint32_t box[2] = {kEpiCols, kBlockM};
    if ((rc = make_tmap_bf16(&tc, C, 2, dims, str, box)) != MERLIN_OK) return rc;
  }
  if (p.epilogue == EPI_RESIDUAL) {
    uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(p.M)};
    uint64_t str[1] = {static_cast<uint64_t>(ldr) * 2};
    uint32_t box[2] = {kEpiCols, kBlockM};
    if ((rc = make_tmap_bf16(&tr, R, 2, dims, str, box)) != MERLIN_OK) return rc;
  } else {
    tr = tc;
  }
  MB_CHECK_ARG(!(a_mn && !b_mn), "gemm: A MN-major with B K-major is not instantiated");
  const CUtensorMap& tkv = kv_map ? *kv_map : tc;
  if (cta_group == 2) {
    if (a_mn) return launch_gemm<2, true, true>(stream, ta, tb, tc, tr, tkv, p);
    if (b_mn) return launch_gemm<2, false, true>(stream, ta, tb, tc, tr, tkv, p);
    return launch_gemm<2, false, false>(stream, ta, tb, tc, tr, tkv, p);
  }
  if (a_mn) return launch_gemm<1, true, true>(stream, ta, tb, tc, tr, tkv, p);
  if (b_mn) return launch_gemm<1, false, true>(stream, ta, tb, tc, tr, tkv, p);
  return launch_gemm<1, false, false>(stream, ta, tb, tc, tr, tkv, p);
}

}  // namespace mb

extern "C" {

const char* merlin_last_error(void) { return mb::g_error; }

int merlin_set_sm_budget(int sms) {
  const int prev = mb::g_sm_budget.exchange(sms < 0 ? 0 : sms, std::memory_order_relaxed);
  return prev;
}

int merlin_sm_count(void) { return mb::sm_count(); }

int merlin_gemm_bf16(void* stream, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                     int64_t ldc, const void* R, int64_t ldr, const void* bias, int M, int N, int K,
                     int epilogue, int cta_group) {
  MB_CHECK_ARG(epilogue >= MERLIN_EPI_STORE && epilogue <= MERLIN_EPI_SWIGLU,
               "merlin_gemm_bf16: unknown epilogue %d", epilogue);
  mb::GemmParams p = {};
  p.kv_col0 = -1;
  p.M = M;
  p.N = N;
  p.K = K;
  p.epilogue = epilogue;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  return mb::gemm_dispatch(static_cast<cudaStream_t>(stream), A, lda, B, ldb, C, ldc, R, ldr, p,
                           cta_group);
}

int merlin_gemm_bf16_ex(void* stream, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                        int b_mn_major, void* C, int64_t ldc, const void* R, int64_t ldr, const void* bias,
                        int M, int N, int K, int epilogue, int cta_group) {
  MB_CHECK_ARG(epilogue >= MERLIN_EPI_STORE && epilogue <= MERLIN_EPI_SWIGLU,
               "merlin_gemm_bf16_ex: unknown epilogue %d", epilogue);
  mb::GemmParams p = {};
  p.kv_col0 = -1;
  p.M = M;
  p.N = N;
  p.K = K;
  p.epilogue = epilogue;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  return mb::gemm_dispatch(static_cast<cudaStream_t>(stream), A, lda, B, ldb, C, ldc, R, ldr, p, cta_group,
                           a_mn_major != 0, b_mn_major != 0);
}

int merlin_gemm_qkv_rope(void* stream, const void* A, int64_t lda, const void* Wqkv, int64_t ldb,
                         void* qkv, int64_t ldc, const float* rope_cos, const float* rope_sin,
                         const int* positions, int M, int N, int K, int rope_cols, int seq_len,
                         int cta_group) {
  return merlin_gemm_qkv_rope_kv(stream, A, lda, Wqkv, ldb, qkv, ldc, rope_cos, rope_sin, positions, M, N, K, rope_cols,
                                 seq_len, nullptr, 0, 0, 0, cta_group);
}

int merlin_gemm_qkv_rope_kv(void* stream, const void* A, int64_t lda, const void* Wqkv, int64_t ldb, void* qkv,
                            int64_t ldc, const float* rope_cos, const float* rope_sin, const int* positions, int M,
                            int N, int K, int rope_cols, int seq_len, void* kv_cache, int64_t kv_capacity, int B,
                            int H, int cta_group) {
  mb::GemmParams p = {};
  p.kv_col0 = -1;
  p.M = M;
  p.N = N;
  p.K = K;
  p.epilogue = mb::EPI_ROPE;
  p.rope_cos = rope_cos;
  p.rope_sin = rope_sin;
  p.positions = positions;
  p.rope_cols = rope_cols;
  p.seq_len = seq_len;
  if (kv_cache == nullptr)
    return mb::gemm_dispatch(static_cast<cudaStream_t>(stream), A, lda, Wqkv, ldb, qkv, ldc, nullptr, 0, p, cta_group);
  const int D = N / 3;
  MB_CHECK_ARG(N == 3 * D && D == H * 128 && rope_cols == 2 * D, "gemm_qkv_rope_kv: needs packed q|k|v with head_dim 128");
  MB_CHECK_ARG(B >= 1 && M == B * seq_len && (B == 1 || seq_len % 128 == 0),
               "gemm_qkv_rope_kv: M=%d must be B*seq_len and seq_len %% 128 == 0 for B > 1 (B=%d, seq_len=%d)", M, B, seq_len);
  MB_CHECK_ARG(kv_capacity % 32 == 0 && seq_len <= kv_capacity, "gemm_qkv_rope_kv: %d tokens do not fit capacity %lld",
               seq_len, static_cast<long long>(kv_capacity));
  // cache [B, cap/32, H, 32, 256] as a 5-D tensor, innermost first: column, token % 32, head, token / 32, sample;
  // box = 64 columns x 32 tokens x 1 head x 4 blocks = the [128 tokens x 64 columns] staging chunk of the epilogue
  CUtensorMap tkv;
  uint64_t dims[5] = {256, 32, static_cast<uint64_t>(H), static_cast<uint64_t>(kv_capacity / 32), static_cast<uint64_t>(B)};
  uint64_t str[4] = {512, 32 * 512, static_cast<uint64_t>(H) * 32 * 512,
                     static_cast<uint64_t>(kv_capacity / 32) * H * 32 * 512};
  uint32_t box[5] = {64, 32, 1, 4, 1};
  int rc = mb::make_tmap_bf16(&tkv, kv_cache, 5, dims, str, box);
  if (rc != MERLIN_OK) return rc;
  p.kv_col0 = D;
  p.kv_hidden = D;
  p.kv_seq = seq_len;
  return mb::gemm_dispatch(static_cast<cudaStream_t>(stream), A, lda, Wqkv, ldb, qkv, ldc, nullptr, 0, p, cta_group,
                           false, false, &tkv);
}

}  // extern "C"
