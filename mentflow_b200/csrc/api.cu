// Library-level entry points of the C ABI (version, error strings, device query).
#include "common.cuh"

extern "C" {

int mfb_abi_version(void) { return MFB_ABI_VERSION; }

const char* mfb_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case MFB_E_BADARG: return "mentflow_b200: bad argument (null pointer, non-positive size or unsupported shape)";
    case MFB_E_UNSUPPORTED: return "mentflow_b200: configuration not supported by the compiled kernels";
    case MFB_E_WORKSPACE: return "mentflow_b200: workspace too small";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "mentflow_b200: unknown error";
  }
}

int mfb_sm_count(void) { return mfb::sm_count(); }

}  // extern "C"
