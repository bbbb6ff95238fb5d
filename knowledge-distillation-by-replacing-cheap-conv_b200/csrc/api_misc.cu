// Version / error-string / dispatch-introspection entry points of the C ABI (include/kdcc.h).
#include "dw_kernels.cuh"
#include "pw_kernels.cuh"
#include "sm100_ptx.cuh"

using namespace kdcc;

KDCC_API int kdcc_version(void) { return KDCC_VERSION; }

KDCC_API int kdcc_last_driver_status(void) { return last_driver_status(); }

KDCC_API const char *kdcc_strerror(int code) {
  switch (code) {
    case KDCC_OK: return "success";
    case KDCC_EINVAL: return "kdcc: invalid argument (null pointer, non-positive dimension or bad enum)";
    case KDCC_ESHAPE: return "kdcc: shape not supported by the sm_100a kernels (no fallback exists)";
    case KDCC_EWORKSPACE: return "kdcc: workspace too small";
    case KDCC_EALIGN: return "kdcc: pointer not 16-byte aligned";
    case KDCC_EDEVICE: return "kdcc: device/driver lacks a required sm_100 feature (TMA descriptor encoder)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "kdcc: unknown error code";
}

KDCC_API const char *kdcc_dispatch_name(int op, int N, int H, int W, int C, int Cout, int k, int dil, int pad,
                                        int dtype) {
  (void)pad;
  const bool bf16 = dtype == KDCC_BF16;
  const long M = (long)N * H * W;
  switch (op) {
    case 0: return bf16 && dw_tma_supported(C, k, dil) ? dw_tma_name(k, dil, 0) : "dw_direct";
    case 1: return bf16 && dw_tma_supported(C, k, dil) ? dw_tma_name(k, dil, 1) : "dw_wgrad_direct";
    case 2: return bf16 && pw_sm100_supported(M, C, Cout) ? "pw_gemm_sm100_tn" : "pw_simt";
    case 3: return bf16 && pw_sm100_supported(M, C, Cout) ? "pw_gemm_sm100_dx" : "pw_simt";
    case 4: return bf16 && pw_sm100_supported(M, C, Cout) ? "pw_gemm_sm100_dw" : "pw_simt";
    default: return "?";
  }
}
