// Host-side helpers shared by the launchers: thread-local error text, CUDA error mapping,
// device properties and the cuTensorMapEncodeTiled entry point (resolved at run time so the
// library links against libcudart only).
#pragma once
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/xmm_b200.h"

namespace xmm {

inline std::string& last_error_ref() {
  thread_local std::string e;
  return e;
}

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

#define XMM_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return ::xmm::fail(XMM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                         __FILE__, __LINE__);                                          \
  } while (0)

#define XMM_REQUIRE(cond, ...)                                         \
  do {                                                                 \
    if (!(cond)) return ::xmm::fail(XMM_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int max_smem_optin = 0;
};

// Per-device info, cached (the GPU under one process never changes: one process per GPU).
// SMs the persistent kernels leave free (xmm_set_sm_reserve): while a collective's kernels share the GPU with a
// 148-CTA persistent grid whose work was divided for 148 resident CTAs, the CTAs that do not fit run as a second wave.
inline std::atomic<int>& sm_reserve() {
  static std::atomic<int> v{0};
  return v;
}

inline int device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  int dev = 0;
  XMM_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(XMM_ERR_UNSUPPORTED_DEVICE, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(mu);
  DeviceInfo& d = cache[dev];
  if (d.device != dev) {
    cudaDeviceProp p;
    XMM_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    d.device = dev;
    d.sm_count = p.multiProcessorCount;
    d.cc_major = p.major;
    d.cc_minor = p.minor;
    d.max_smem_optin = int(p.sharedMemPerBlockOptin);
  }
  *out = d;
  const int reserve = sm_reserve().load(std::memory_order_relaxed);
  if (reserve > 0 && out->sm_count - reserve >= 8) out->sm_count -= reserve;
  return XMM_OK;
}

inline int require_sm100(DeviceInfo* info) {
  int rc = device_info(info);
  if (rc != XMM_OK) return rc;
  if (info->cc_major != 10)
    return fail(XMM_ERR_UNSUPPORTED_DEVICE,
                "libxmm_b200 needs an sm_100a GPU (B200); device %d is sm_%d%d and there is no fallback path",
                info->device, info->cc_major, info->cc_minor);
  return XMM_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
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

// NHWC bf16 activation buffer [B][H][W][ctot] viewed as a 4-D tensor (c, x, y, b); boxes are
// (box_c channels) x (box_w) x (box_h) pixels of one image, zero-filled outside the image --
// that zero fill *is* the convolution's padding=1.
inline int make_nhwc_tmap(CUtensorMap* out, const void* base, int batch, int height, int width, int ctot,
                          int box_c, int box_w, int box_h, bool mn_swizzle128) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(XMM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {cuuint64_t(ctot), cuuint64_t(width), cuuint64_t(height), cuuint64_t(batch)};
  cuuint64_t strides[3] = {cuuint64_t(ctot) * 2, cuuint64_t(ctot) * 2 * width,
                           cuuint64_t(ctot) * 2 * width * height};
  cuuint32_t box[4] = {cuuint32_t(box_c), cuuint32_t(box_w), cuuint32_t(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (mn_swizzle128 || box_c * 2 == 128)
    sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (box_c * 2 == 64)
    sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (box_c * 2 == 32)
    sw = CU_TENSOR_MAP_SWIZZLE_32B;
  // L2 promotion widens every TMA request to the given sector span.  The box's contiguous run is only
  // box_c*2 bytes per pixel, so promoting past it over-fetches neighbouring channels from DRAM (measured
  // 4x on 64-byte chunks with L2_256B).  XMM_TMAP_L2PROMO=0..3 overrides for experiments.
  static const int promo_env = [] {
    const char* e = getenv("XMM_TMAP_L2PROMO");
    return e ? atoi(e) : -1;
  }();
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (box_c * 2 <= 64) promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  if (promo_env >= 0 && promo_env <= 3) promo = CUtensorMapL2promotion(promo_env);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(XMM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box [%d,%d,%d]", int(r), batch,
                height, width, ctot, box_c, box_w, box_h);
  return XMM_OK;
}

// The same NHWC buffer viewed as nbands horizontal bands of band_h rows (height == nbands * band_h): a 5-D tensor
// (c, x, row-in-band, band, b).  A box is (box_c channels) x (box_w pixels) of ONE row-in-band of box_bands
// consecutive bands -- the M = [bands][pixels] tile of the row-hop conv (conv3x3_row.cuh).  Bands / pixels outside
// the tensor are zero-filled on loads and clipped on stores.
inline int make_band_tmap(CUtensorMap* out, const void* base, int batch, int height, int width, int ctot, int band_h,
                          int nbands, int box_c, int box_w, int box_bands) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(XMM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (band_h * nbands != height) return fail(XMM_ERR_INVALID_ARGUMENT, "band view: %d x %d != height %d", nbands, band_h, height);
  cuuint64_t dims[5] = {cuuint64_t(ctot), cuuint64_t(width), cuuint64_t(band_h), cuuint64_t(nbands), cuuint64_t(batch)};
  cuuint64_t strides[4] = {cuuint64_t(ctot) * 2, cuuint64_t(ctot) * 2 * width, cuuint64_t(ctot) * 2 * width * band_h,
                           cuuint64_t(ctot) * 2 * width * height};
  cuuint32_t box[5] = {cuuint32_t(box_c), cuuint32_t(box_w), 1, cuuint32_t(box_bands), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (box_c * 2 == 128)
    sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (box_c * 2 == 64)
    sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (box_c * 2 == 32)
    sw = CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMapL2promotion promo = box_c * 2 <= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(XMM_ERR_CUDA, "cuTensorMapEncodeTiled (band view) failed (%d) for [%d,%d,%d,%d] bands %dx%d box [%d,%d,%d]",
                int(r), batch, height, width, ctot, nbands, band_h, box_c, box_w, box_bands);
  return XMM_OK;
}

}  // namespace xmm
