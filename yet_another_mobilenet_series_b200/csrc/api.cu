// C ABI entry points (include/yamb200.h) + host utilities.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "host_util.h"

namespace yamb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int max_ctas() {
  static int cached = 0;
  if (cached != 0) return cached;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return cached = -1; }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return cached = -1;
  }
  return cached = sms;
}

PFN_encodeTiled get_encode_tiled() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return reinterpret_cast<PFN_encodeTiled>(fn);
}

}  // namespace yamb

extern "C" {

const char* yamb_last_error(void) { return yamb::g_err; }
int yamb_version(void) { return 100; }
int yamb_max_ctas(void) { return yamb::max_ctas(); }
int yamb_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(yamb_bn_fwd);
    case 1: return (int)sizeof(yamb_bn_bwd);
    case 2: return (int)sizeof(yamb_gemm);
    default: return -1;
  }
}

int yamb_pointwise_gemm(const yamb_gemm* args, yamb_stream_t stream) {
  return yamb::gemm_launch(args, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
