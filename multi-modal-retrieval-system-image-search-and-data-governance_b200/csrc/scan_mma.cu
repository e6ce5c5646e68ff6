// scan_mma.cu -- K2 placeholder (replaced by the tcgen05 kernel).
#include "common.cuh"
namespace mmrs {
int scan_mma_max_queries() { return 256; }
bool scan_mma_available() { return false; }
cudaError_t launch_scan_mma(const ScanParams&, const __nv_bfloat16*, int32_t, int, int32_t*, int,
                            cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace mmrs
