// merlin_b200 — image path of the data loader on the GPU: the 'resize' mode of
// mmgpt/data/dataset/base_dataset.py:190-192 / mmgpt/utils/mm_utils.py:41-43,
//     image = image.resize((S, S));  processor.preprocess(image, do_resize=False, do_center_crop=False)
// i.e. PIL's default BICUBIC resample of the uint8 RGB pixels to S x S followed by the CLIP processor's
// rescale (x / 255) and normalize ((x - mean) / std).
//
// PIL's resample (third-party: Pillow, src/libImaging/Resample.c) is integer work and is reproduced BIT-EXACTLY:
// per output coordinate a window [xmin, xmin + n) of source pixels and n filter weights (bicubic, a = -0.5, the
// support widened by the scale factor when shrinking), the weights normalised to sum 1 in double precision and
// rounded to 22-bit fixed point; a horizontal pass to uint8 (accumulate in int32 from 1 << 21, shift, clamp),
// then a vertical pass on that uint8 intermediate.  The weight tables are built on the host exactly like
// precompute_coeffs / normalize_coeffs_8bpc do (double arithmetic); the two passes are HBM-bound byte kernels.
#include "host.cuh"

#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

namespace mb {

constexpr int kPrecisionBits = 32 - 8 - 2;

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box [0, in_size)
static int precompute(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  kk.assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << kPrecisionBits);
      kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  return ksize;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;   // arithmetic shift, like the clip8 lookup of Resample.c
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: src [H, W, 3] uint8 (row pitch in bytes) -> tmp [H, S, 3] uint8
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, int64_t pitch, int H, int S,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                uint8_t* __restrict__ tmp) {
  const int64_t total = static_cast<int64_t>(H) * S;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % S), yy = static_cast<int>(i / S);
    const int xmin = bounds[xx * 2], n = bounds[xx * 2 + 1];
    const int* k = kk + static_cast<size_t>(xx) * ksize;
    const uint8_t* row = src + yy * pitch + static_cast<size_t>(xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
      const int w = k[x];
      s0 += row[x * 3 + 0] * w;
      s1 += row[x * 3 + 1] * w;
      s2 += row[x * 3 + 2] * w;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = static_cast<uint8_t>(clip8(s0));
    o[1] = static_cast<uint8_t>(clip8(s1));
    o[2] = static_cast<uint8_t>(clip8(s2));
  }
}

// vertical pass on the uint8 intermediate + rescale / normalize: tmp [H, S, 3] -> out [3, S, S] (fp32 or bf16)
template <bool kBf16>
__global__ void resize_v_norm_kernel(const uint8_t* __restrict__ tmp, int S, const int* __restrict__ bounds,
                                     const int* __restrict__ kk, int ksize, float m0, float m1, float m2, float d0,
                                     float d1, float d2, void* __restrict__ out, uint8_t* __restrict__ out_u8) {
  const int64_t total = static_cast<int64_t>(S) * S;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % S), yy = static_cast<int>(i / S);
    const int ymin = bounds[yy * 2], n = bounds[yy * 2 + 1];
    const int* k = kk + static_cast<size_t>(yy) * ksize;
    const uint8_t* col = tmp + (static_cast<size_t>(ymin) * S + xx) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < n; ++y) {
      const int w = k[y];
      const uint8_t* p = col + static_cast<size_t>(y) * S * 3;
      s0 += p[0] * w;
      s1 += p[1] * w;
      s2 += p[2] * w;
    }
    const int c0 = clip8(s0), c1 = clip8(s1), c2 = clip8(s2);
    if (out_u8 != nullptr) {
      out_u8[i * 3 + 0] = static_cast<uint8_t>(c0);
      out_u8[i * 3 + 1] = static_cast<uint8_t>(c1);
      out_u8[i * 3 + 2] = static_cast<uint8_t>(c2);
    }
    // HF image_transforms.rescale: float32(float64(u8) * (1 / 255)); normalize: (x - mean) / std in float32
    const float r0 = static_cast<float>(static_cast<double>(c0) * (1.0 / 255.0));
    const float r1 = static_cast<float>(static_cast<double>(c1) * (1.0 / 255.0));
    const float r2 = static_cast<float>(static_cast<double>(c2) * (1.0 / 255.0));
    const float v0 = __fdiv_rn(__fsub_rn(r0, m0), d0), v1 = __fdiv_rn(__fsub_rn(r1, m1), d1),
                v2 = __fdiv_rn(__fsub_rn(r2, m2), d2);
    if (kBf16) {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
      o[i] = __float2bfloat16_rn(v0);
      o[total + i] = __float2bfloat16_rn(v1);
      o[2 * total + i] = __float2bfloat16_rn(v2);
    } else {
      float* o = static_cast<float*>(out);
      o[i] = v0;
      o[total + i] = v1;
      o[2 * total + i] = v2;
    }
  }
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }
static int ksize_for(int in_size, int out_size) {
  double fs = static_cast<double>(in_size) / out_size;
  if (fs < 1.0) fs = 1.0;
  return static_cast<int>(std::ceil(2.0 * fs)) * 2 + 1;
}

}  // namespace mb

using namespace mb;

extern "C" {

size_t merlin_image_preprocess_workspace_bytes(int H, int W, int S) {
  if (H <= 0 || W <= 0 || S <= 0) return 0;
  size_t b = 0;
  b += align256(static_cast<size_t>(S) * 2 * 4) * 2;                       // bounds (h, v)
  b += align256(static_cast<size_t>(S) * ksize_for(W, S) * 4);             // weights h
  b += align256(static_cast<size_t>(S) * ksize_for(H, S) * 4);             // weights v
  b += align256(static_cast<size_t>(H) * S * 3);                           // uint8 intermediate
  return b + 256;
}

int merlin_image_preprocess(void* stream, const void* src_rgb_u8, int64_t src_pitch_bytes, int H, int W, int S,
                            const float* mean3, const float* std3, void* workspace, void* out, int out_bf16,
                            void* out_resized_u8) {
  MB_CHECK_ARG(src_rgb_u8 && workspace && out && mean3 && std3, "image_preprocess: null argument");
  MB_CHECK_ARG(H > 0 && W > 0 && S > 0 && src_pitch_bytes >= static_cast<int64_t>(W) * 3,
               "image_preprocess: bad shape H=%d W=%d S=%d pitch=%lld", H, W, S, static_cast<long long>(src_pitch_bytes));
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "image_preprocess: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int> bh, kh, bv, kv;
  const int ksh = precompute(W, S, bh, kh), ksv = precompute(H, S, bv, kv);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* d_bh = reinterpret_cast<int*>(ws);
  ws += align256(bh.size() * 4);
  int* d_bv = reinterpret_cast<int*>(ws);
  ws += align256(bv.size() * 4);
  int* d_kh = reinterpret_cast<int*>(ws);
  ws += align256(kh.size() * 4);
  int* d_kv = reinterpret_cast<int*>(ws);
  ws += align256(kv.size() * 4);
  uint8_t* tmp = ws;
  // the tables live in host vectors that die with this call: synchronous copies (this is the data-loading path,
  // not a capturable per-step entry point)
  MB_CHECK_CUDA(cudaStreamSynchronize(st));
  MB_CHECK_CUDA(cudaMemcpy(d_bh, bh.data(), bh.size() * 4, cudaMemcpyHostToDevice));
  MB_CHECK_CUDA(cudaMemcpy(d_bv, bv.data(), bv.size() * 4, cudaMemcpyHostToDevice));
  MB_CHECK_CUDA(cudaMemcpy(d_kh, kh.data(), kh.size() * 4, cudaMemcpyHostToDevice));
  MB_CHECK_CUDA(cudaMemcpy(d_kv, kv.data(), kv.size() * 4, cudaMemcpyHostToDevice));
  const int threads = 256;
  const int64_t n1 = static_cast<int64_t>(H) * S, n2 = static_cast<int64_t>(S) * S;
  const int cap = sm_count() * 16;
  const int g1 = static_cast<int>(std::min<int64_t>((n1 + threads - 1) / threads, cap));
  const int g2 = static_cast<int>(std::min<int64_t>((n2 + threads - 1) / threads, cap));
  // PIL skips a pass whose size does not change; a pass with the identity size has ksize 5 and weights that are
  // exactly {0, .., 1 << 22, .., 0}, so running it is the identity too — one code path.
  resize_h_kernel<<<g1, threads, 0, st>>>(static_cast<const uint8_t*>(src_rgb_u8), src_pitch_bytes, H, S, d_bh, d_kh, ksh, tmp);
  if (out_bf16)
    resize_v_norm_kernel<true><<<g2, threads, 0, st>>>(tmp, S, d_bv, d_kv, ksv, mean3[0], mean3[1], mean3[2], std3[0],
                                                     std3[1], std3[2], out, static_cast<uint8_t*>(out_resized_u8));
  else
    resize_v_norm_kernel<false><<<g2, threads, 0, st>>>(tmp, S, d_bv, d_kv, ksv, mean3[0], mean3[1], mean3[2], std3[0],
                                                      std3[1], std3[2], out, static_cast<uint8_t*>(out_resized_u8));
  MB_CHECK_CUDA(cudaGetLastError());
  return MERLIN_OK;
}

}  // extern "C"
