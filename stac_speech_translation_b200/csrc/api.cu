// Library-level entry points of libstac_b200.
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
