// Input pipeline on the GPU: the transform of the zero-shot tool, tools/zero_shot.py:202-207 -
//   Resize(S, interpolation=BICUBIC) -> CenterCrop(S) -> ToTensor() -> Normalize(mean, std)
// for a batch of decoded RGB images (uint8, HWC) of arbitrary sizes.  Byte / integer work, so the bar is bit-exactness:
// torchvision hands PIL images to Pillow's ImagingResample (third-party dependency of the reference, not vendored; this
// restates the published algorithm of Pillow's src/libImaging/Resample.c as shipped in Pillow 12.x):
//   * per axis, per output pixel: window [xmin, xmax) around centre (xx + 0.5) * scale of half-width 2 * max(scale, 1)
//     (bicubic, a = -0.5), weights normalised in double, then converted to 22-bit fixed point with round-half-away;
//   * horizontal pass first into a uint8 image (accumulate in int32 from 1 << 21, arithmetic shift by 22, clip to 0..255),
//     then the vertical pass on that rounded image with the same arithmetic.
// The coefficient tables depend only on (input size, output size) and are computed on the host in double exactly as
// Pillow does (no FMA contraction on the x86-64 baseline); the device does the integer part.  Cropping commutes with the
// per-pixel arithmetic, so only the S output columns (pass 1) and S output rows (pass 2) inside the crop are computed.
// ToTensor / Normalize follow torch: float(u8) / 255 (IEEE division), then (x - mean) / std in fp32.
// HBM-bound: every source byte is read once per pass from L2-resident rows; one launch per pass for the whole batch.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "engine.h"

namespace msclip {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

struct ImgDesc {
  const uint8_t* src;  // [H][W][3]
  uint8_t* tmp;        // [H][S][3] after the horizontal pass (crop columns only)
  int H, W;
  int row0, rows;      // source rows the vertical pass of the cropped output reads: [row0, row0 + rows)
  int kh_off, kh_size, bh_off;  // horizontal coefficients [S][kh_size] / bounds [S][2] in the pools (crop columns)
  int kv_off, kv_size, bv_off;  // vertical coefficients [S][kv_size] / bounds [S][2] (crop rows)
};

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output pixels [o0, o0 + count) of an axis resized inSize -> outSize
int precompute_coeffs(int inSize, int outSize, int o0, int count, std::vector<int>& kk, std::vector<int>& bounds) {
  const double scale = static_cast<double>(inSize) / outSize;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  kk.assign(static_cast<size_t>(count) * ksize, 0);
  bounds.assign(static_cast<size_t>(count) * 2, 0);
  std::vector<double> k(ksize);
  for (int i = 0; i < count; ++i) {
    const int xx = o0 + i;
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > inSize) xmax = inSize;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << kPrecisionBits);
      kk[static_cast<size_t>(i) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    bounds[2 * i] = xmin;
    bounds[2 * i + 1] = xmax;
  }
  return ksize;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// pass 1: tmp[r][x][c] = clip8(sum_j src[row0 + r][xmin + j][c] * k[x][j]); one thread per (r, x), grid.z = image
__global__ void __launch_bounds__(256)
resize_horizontal_kernel(const ImgDesc* __restrict__ descs, const int* __restrict__ coeffs, const int* __restrict__ bounds, int S) {
  const ImgDesc d = descs[blockIdx.z];
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= S) return;
  const int* k = coeffs + d.kh_off + x * d.kh_size;
  const int xmin = bounds[d.bh_off + 2 * x], n = bounds[d.bh_off + 2 * x + 1];
  for (int r = blockIdx.y; r < d.rows; r += gridDim.y) {
    const uint8_t* line = d.src + (static_cast<size_t>(d.row0 + r) * d.W + xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < n; ++j) {
      const int kj = __ldg(k + j);
      s0 += line[3 * j] * kj;
      s1 += line[3 * j + 1] * kj;
      s2 += line[3 * j + 2] * kj;
    }
    uint8_t* o = d.tmp + (static_cast<size_t>(r) * S + x) * 3;
    o[0] = static_cast<uint8_t>(clip8(s0));
    o[1] = static_cast<uint8_t>(clip8(s1));
    o[2] = static_cast<uint8_t>(clip8(s2));
  }
}

template <typename T>
__device__ __forceinline__ T to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }

// pass 2 + ToTensor + Normalize: out[img][c][y][x] = (clip8(sum_j tmp[ymin - row0 + j][x][c] * k[y][j]) / 255 - mean[c]) / std[c]
template <typename T>
__global__ void __launch_bounds__(256)
resize_vertical_normalize_kernel(const ImgDesc* __restrict__ descs, const int* __restrict__ coeffs, const int* __restrict__ bounds, int S,
                                 float m0, float m1, float m2, float d0, float d1, float d2, T* __restrict__ out, uint8_t* __restrict__ out_u8) {
  const ImgDesc d = descs[blockIdx.z];
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= S) return;
  const int* k = coeffs + d.kv_off + y * d.kv_size;
  const int ymin = bounds[d.bv_off + 2 * y] - d.row0, n = bounds[d.bv_off + 2 * y + 1];
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int j = 0; j < n; ++j) {
    const int kj = __ldg(k + j);
    const uint8_t* p = d.tmp + (static_cast<size_t>(ymin + j) * S + x) * 3;
    s0 += p[0] * kj;
    s1 += p[1] * kj;
    s2 += p[2] * kj;
  }
  const int v0 = clip8(s0), v1 = clip8(s1), v2 = clip8(s2);
  const size_t plane = static_cast<size_t>(S) * S;
  const size_t o = static_cast<size_t>(blockIdx.z) * 3 * plane + static_cast<size_t>(y) * S + x;
  if (out_u8 != nullptr) {  // the resized + cropped bytes themselves (parity against Pillow)
    uint8_t* q = out_u8 + (static_cast<size_t>(blockIdx.z) * plane + static_cast<size_t>(y) * S + x) * 3;
    q[0] = static_cast<uint8_t>(v0);
    q[1] = static_cast<uint8_t>(v1);
    q[2] = static_cast<uint8_t>(v2);
  }
  if (out != nullptr) {
    out[o] = to_out<T>(__fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v0), 255.0f), m0), d0));
    out[o + plane] = to_out<T>(__fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v1), 255.0f), m1), d1));
    out[o + 2 * plane] = to_out<T>(__fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v2), 255.0f), m2), d2));
  }
}

}  // namespace

// pixels: n images back to back (image i at byte offset offsets[i], heights[i] x widths[i] x 3 uint8), host or device memory.
// out [n, 3, S, S] in out_dtype (MSCLIP_F32 / BF16 / F16; may be null) and / or out_u8 [n, S, S, 3] (may be null), device memory.
int engine_preprocess(msclip_ctx* h, const uint8_t* pixels, const int64_t* offsets, const int* heights, const int* widths, int n, int S,
                      const float* mean, const float* stdv, void* out, int out_dtype, uint8_t* out_u8, cudaStream_t s) {
  MSCLIP_REQUIRE(h != nullptr && pixels && offsets && heights && widths && mean && stdv, "preprocess: null argument");
  MSCLIP_REQUIRE(n >= 1 && S >= 1 && S <= 4096 && (out != nullptr || out_u8 != nullptr), "preprocess: bad arguments");
  MSCLIP_REQUIRE(out_dtype == MSCLIP_F32 || out_dtype == MSCLIP_BF16 || out_dtype == MSCLIP_F16, "preprocess: output dtype must be f32 / bf16 / f16");
  // geometry and coefficient tables (host, double arithmetic as in Pillow); identical sizes share their tables
  std::vector<ImgDesc> descs(n);
  std::vector<int> coeffs, bounds;
  struct Key {
    int in, out, o0, koff, ksize, boff, first, last;
  };
  std::vector<Key> cache;
  auto tables = [&](int inSize, int outSize, int o0) -> const Key& {
    for (const Key& k : cache)
      if (k.in == inSize && k.out == outSize && k.o0 == o0) return k;
    std::vector<int> kk, bb;
    const int ksize = precompute_coeffs(inSize, outSize, o0, S, kk, bb);
    Key k = {inSize, outSize, o0, static_cast<int>(coeffs.size()), ksize, static_cast<int>(bounds.size()), bb[0], bb[2 * (S - 1)] + bb[2 * (S - 1) + 1]};
    coeffs.insert(coeffs.end(), kk.begin(), kk.end());
    bounds.insert(bounds.end(), bb.begin(), bb.end());
    cache.push_back(k);
    return cache.back();
  };
  size_t total_bytes = 0, tmp_bytes = 0;
  int max_rows = 0;
  std::vector<size_t> tmp_off(n);
  for (int i = 0; i < n; ++i) {
    const int H = heights[i], W = widths[i];
    MSCLIP_REQUIRE(H >= 1 && W >= 1 && H <= 32768 && W <= 32768, "preprocess: image size out of range");
    // torchvision _compute_resized_output_size: the shorter edge becomes S, the longer int(S * long / short)
    const int shorter = W <= H ? W : H, longer = W <= H ? H : W;
    const int new_long = static_cast<int>(static_cast<double>(static_cast<long long>(S) * longer) / static_cast<double>(shorter));
    const int newW = W <= H ? S : new_long, newH = W <= H ? new_long : S;
    // center_crop: int(round((size - S) / 2.0)), round half to even
    const int top = static_cast<int>(nearbyint((newH - S) / 2.0)), left = static_cast<int>(nearbyint((newW - S) / 2.0));
    const Key kh = tables(W, newW, left);
    const Key kv = tables(H, newH, top);
    ImgDesc& d = descs[i];
    d.H = H;
    d.W = W;
    d.row0 = kv.first;
    d.rows = kv.last - kv.first;
    d.kh_off = kh.koff;
    d.kh_size = kh.ksize;
    d.bh_off = kh.boff;
    d.kv_off = kv.koff;
    d.kv_size = kv.ksize;
    d.bv_off = kv.boff;
    tmp_off[i] = tmp_bytes;
    tmp_bytes += (static_cast<size_t>(d.rows) * S * 3 + 255) & ~size_t(255);
    max_rows = std::max(max_rows, d.rows);
    total_bytes = std::max(total_bytes, static_cast<size_t>(offsets[i]) + static_cast<size_t>(H) * W * 3);
  }
  // source pixels: in place when already on the device, else one host-to-device copy of the whole batch
  const uint8_t* src_dev = pixels;
  if (!is_device_pointer(pixels)) {
    WS(staged, uint8_t, "pre_pixels", total_bytes);
    MSCLIP_CHECK_CUDA(cudaMemcpyAsync(staged, pixels, total_bytes, cudaMemcpyHostToDevice, s));
    src_dev = staged;
  }
  WS(tmp, uint8_t, "pre_tmp", tmp_bytes);
  for (int i = 0; i < n; ++i) {
    descs[i].src = src_dev + offsets[i];
    descs[i].tmp = tmp + tmp_off[i];
  }
  const size_t b0 = (descs.size() * sizeof(ImgDesc) + 255) & ~size_t(255), b1 = (coeffs.size() * 4 + 255) & ~size_t(255), b2 = bounds.size() * 4;
  WS(tab, uint8_t, "pre_tables", b0 + b1 + b2);
  std::vector<uint8_t> host(b0 + b1 + b2);
  memcpy(host.data(), descs.data(), descs.size() * sizeof(ImgDesc));
  memcpy(host.data() + b0, coeffs.data(), coeffs.size() * 4);
  memcpy(host.data() + b0 + b1, bounds.data(), bounds.size() * 4);
  MSCLIP_CHECK_CUDA(cudaMemcpyAsync(tab, host.data(), host.size(), cudaMemcpyHostToDevice, s));
  MSCLIP_CHECK_CUDA(cudaStreamSynchronize(s));  // the pageable host tables must outlive the copy
  const ImgDesc* ddev = reinterpret_cast<const ImgDesc*>(tab);
  const int* cdev = reinterpret_cast<const int*>(tab + b0);
  const int* bdev = reinterpret_cast<const int*>(tab + b0 + b1);
  const int tx = S < 256 ? ((S + 31) / 32) * 32 : 256;
  const dim3 g1((S + tx - 1) / tx, std::min(max_rows, 1024), n), g2((S + tx - 1) / tx, S, n);
  resize_horizontal_kernel<<<g1, tx, 0, s>>>(ddev, cdev, bdev, S);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  if (out_dtype == MSCLIP_F32)
    resize_vertical_normalize_kernel<float><<<g2, tx, 0, s>>>(ddev, cdev, bdev, S, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2],
                                                              static_cast<float*>(out), out_u8);
  else if (out_dtype == MSCLIP_BF16)
    resize_vertical_normalize_kernel<__nv_bfloat16><<<g2, tx, 0, s>>>(ddev, cdev, bdev, S, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                                      stdv[2], static_cast<__nv_bfloat16*>(out), out_u8);
  else
    resize_vertical_normalize_kernel<__half><<<g2, tx, 0, s>>>(ddev, cdev, bdev, S, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2],
                                                               static_cast<__half*>(out), out_u8);
  MSCLIP_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

}  // namespace msclip
