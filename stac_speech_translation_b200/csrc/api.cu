// Library-level entry points of libstac_b200.
#include <algorithm>
#include "common.cuh"

extern "C" int stac_version(void) { return STAC_B200_VERSION; }

extern "C" const char* stac_error_string(int code) {
  switch (code) {
    case STAC_OK: return "ok";
    case STAC_ERR_INVALID_ARGUMENT: return "stac_b200: invalid argument";
    case STAC_ERR_UNSUPPORTED_SHAPE: return "stac_b200: unsupported shape";
    case STAC_ERR_DRIVER_ENTRY: return "stac_b200: cuTensorMapEncodeTiled entry point unavailable";
    case STAC_ERR_TENSOR_MAP: return "stac_b200: cuTensorMapEncodeTiled failed";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "stac_b200: unknown error";
}

// ---- persistent-kernel grid size --------------------------------------------------------------------------------
// Every persistent kernel launches min(work items, stac_grid_limit()) CTAs.  By default that is the SM count; a host
// that runs a communication kernel beside the path (the NCCL transfers of the multi-GPU gather occupy whole SMs for
// milliseconds) lowers it with stac_set_reserved_sms(), so that no CTA of a persistent grid has to wait a full kernel
// duration for an SM that the collective is holding.
static int g_reserved_sms = 0;

int stac_grid_limit() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  return n_sm - g_reserved_sms > 8 ? n_sm - g_reserved_sms : 8;
}

extern "C" int stac_set_reserved_sms(int n) {
  if (n < 0 || n > 128) return STAC_ERR_INVALID_ARGUMENT;
  g_reserved_sms = n;
  return STAC_OK;
}

// L2 residency hint for a buffer that the kernels of one stream read and update over and over (the encoder's fp32 residual
// stream: read by both LayerNorms and reduce-added to by the out-proj and feed-forward kernels of every layer while ~175 MB
// of other activations per layer stream through the 126 MB L2).  Sets the device's persisting-L2 carve-out (at most what
// the device allows) and the stream's access-policy window; kernels launched - or captured into a CUDA graph - on that
// stream afterwards carry it.  bytes == 0 clears the window.
extern "C" int stac_l2_persist(const void* base, int64_t bytes, float hit_ratio, void* stream) {
  if (bytes < 0 || hit_ratio < 0.f || hit_ratio > 1.f || (bytes > 0 && !base)) return STAC_ERR_INVALID_ARGUMENT;
  int dev = 0, max_persist = 0, max_window = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  if (max_persist <= 0 || max_window <= 0) { (void)cudaGetLastError(); return STAC_ERR_UNSUPPORTED_SHAPE; }
  cudaStreamAttrValue v = {};
  if (bytes > 0) {
    static size_t carve = 0;                       // grown, never shrunk: other windows may be live
    const size_t want = (size_t)std::min<int64_t>(bytes, max_persist);
    if (want > carve) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      if (e != cudaSuccess) return (int)cudaGetLastError();
      carve = want;
    }
    v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    v.accessPolicyWindow.num_bytes = (size_t)std::min<int64_t>(bytes, max_window);
    v.accessPolicyWindow.hitRatio = hit_ratio;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  }
  cudaError_t e = cudaStreamSetAttribute(as_stream(stream), cudaStreamAttributeAccessPolicyWindow, &v);
  if (e != cudaSuccess) return (int)cudaGetLastError();
  return STAC_OK;
}

extern "C" int stac_l2_persist_limits(int64_t* max_persist_bytes, int64_t* max_window_bytes) {
  int dev = 0, a = 0, b = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  (void)cudaGetLastError();
  if (max_persist_bytes) *max_persist_bytes = a;
  if (max_window_bytes) *max_window_bytes = b;
  return STAC_OK;
}
