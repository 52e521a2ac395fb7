// Host runtime shared by every kernel launcher: last-error storage for the C ABI and the TMA
// tensor-map encoder (driver entry point resolved at run time, so the library has no link-time
// dependency on libcuda and can be dlopen'ed on a box without a GPU driver).
#include "common.cuh"

#include <mutex>

namespace msclip {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

int make_tmap_op16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  MSCLIP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  MSCLIP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  MSCLIP_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  MSCLIP_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, MSCLIP_TMA_DTYPE, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
                   " (rows=" + std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" +
                   std::to_string(ld) + " box_rows=" + std::to_string(box_rows) + ")");
    return 1;
  }
  return 0;
}

int make_tmap_op16_3d(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  MSCLIP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  MSCLIP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, "TMA base / row pitch must be 16-byte aligned");
  MSCLIP_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, MSCLIP_TMA_DTYPE, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (3D) failed with CUresult " + std::to_string(static_cast<int>(r)) + " (batch=" +
                   std::to_string(batch) + " rows=" + std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" +
                   std::to_string(ld) + " box_rows=" + std::to_string(box_rows) + ")");
    return 1;
  }
  return 0;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_im2col_nhwc(CUtensorMap* out, const void* base, int N, int H, int W, int cpix, int c_off, int C, int ksize,
                          int stride, int pad) {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  MSCLIP_REQUIRE(fn != nullptr, "cuTensorMapEncodeIm2col is unavailable (no CUDA driver?)");
  const uint8_t* origin = static_cast<const uint8_t*>(base) + static_cast<size_t>(c_off) * 2;
  MSCLIP_REQUIRE((reinterpret_cast<uintptr_t>(origin) & 15) == 0 && (cpix * 2) % 16 == 0,
                 "im2col TMA: base and pixel pitch must be 16-byte aligned");
  MSCLIP_REQUIRE(ksize >= 1 && stride >= 1 && stride <= 8 && pad >= 0 && pad < 128 && ksize - 1 - pad < 128,
                 "im2col TMA: filter geometry out of range");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(cpix) * 2, static_cast<cuuint64_t>(W) * cpix * 2,
                           static_cast<cuuint64_t>(H) * W * cpix * 2};
  // window origins (base pixels) live in [lower, dim + upper): lower = -pad, upper = pad - (ksize - 1)
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = fn(out, MSCLIP_TMA_DTYPE, 4, const_cast<uint8_t*>(origin), dims, strides, lower, upper, 64, 128, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeIm2col failed with CUresult " + std::to_string(static_cast<int>(r)) + " (N=" +
                   std::to_string(N) + " H=" + std::to_string(H) + " W=" + std::to_string(W) + " C=" + std::to_string(C) +
                   " cpix=" + std::to_string(cpix) + " k=" + std::to_string(ksize) + " s=" + std::to_string(stride) + " p=" +
                   std::to_string(pad) + ")");
    return 1;
  }
  return 0;
}

// cuStreamWaitValue32(stream, addr, value, GEQ): the stream stalls (no SM is occupied) until the 32-bit word at
// `addr` - device memory that a peer GPU writes over NVLink - satisfies (int32)(*addr - value) >= 0.
int stream_wait_value_geq(cudaStream_t stream, const uint32_t* addr, uint32_t value) {
  typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static WaitFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WaitFn>(p);
  });
  if (fn == nullptr) return 1;
  return fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WAIT_VALUE_GEQ) ==
                 CUDA_SUCCESS
             ? 0
             : 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  return n;
}

}  // namespace msclip
