// Library identity and error plumbing for lib3dvlp_b200.so (see include/pointnet2_b200.h).
#include <cstring>

#include "common.cuh"

namespace pn2 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return PN2_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

}  // namespace pn2

extern "C" int pn2_abi_version(void) { return 1; }

extern "C" const char* pn2_last_error(void) { return pn2::g_err; }

extern "C" int pn2_device_arch(int* major, int* minor) {
  int dev = 0;
  PN2_CUDA(cudaGetDevice(&dev));
  int ma = 0, mi = 0;
  PN2_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  PN2_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (major) *major = ma;
  if (minor) *minor = mi;
  return PN2_OK;
}
