// ABI version + error strings for the C-ABI (include/streammos_b200.h).
#include "common.cuh"

extern "C" {

int smos_abi_version(void) { return SMOS_ABI_VERSION; }

const char* smos_error_string(int code) {
  if (code == SMOS_OK) return "ok";
  if (code == SMOS_EINVAL) return "streammos_b200: invalid argument (shape, null pointer or alignment)";
  if (code == SMOS_EUNSUPPORTED) return "streammos_b200: unsupported size or dtype";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "streammos_b200: unknown error";
}

}  // extern "C"
