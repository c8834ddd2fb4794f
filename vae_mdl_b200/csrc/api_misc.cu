// api_misc.cu -- version / error strings / cached device properties.
#include "common.cuh"

namespace vaemdl {

const DeviceInfo& device_info() {
  static DeviceInfo cache[64];
  static bool have[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!have[dev]) {
    DeviceInfo d{};
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (d.sm_count <= 0) d.sm_count = 148;
    if (d.max_smem_optin <= 0) d.max_smem_optin = 227 * 1024;
    cache[dev] = d;
    have[dev] = true;
  }
  return cache[dev];
}

}  // namespace vaemdl

extern "C" int vaemdl_version(void) { return VAEMDL_ABI_VERSION; }

extern "C" const char* vaemdl_strerror(int code) {
  switch (code) {
    case VAEMDL_OK:
      return "ok";
    case VAEMDL_EINVAL:
      return "invalid argument";
    case VAEMDL_EALIGN:
      return "parameter/gradient pointer is not 16-byte aligned";
    case VAEMDL_EUNSUPPORTED:
      return "n_mix out of range";
    case VAEMDL_EWORKSPACE:
      return "workspace too small";
    default:
      if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
      return "unknown error";
  }
}
