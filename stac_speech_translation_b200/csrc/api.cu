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
