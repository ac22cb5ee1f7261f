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

int smos_stream_capture_id(void* stream, uint64_t* id_host) {
  if (id_host == nullptr) return SMOS_EINVAL;
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  const cudaError_t e = cudaStreamGetCaptureInfo(smos_stream(stream), &status, &id);
  if (e != cudaSuccess) return static_cast<int>(e);
  *id_host = status == cudaStreamCaptureStatusActive ? static_cast<uint64_t>(id) : 0;
  return SMOS_OK;
}

}  // extern "C"
