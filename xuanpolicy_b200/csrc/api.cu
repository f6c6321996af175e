// api.cu — version and error-string entry points of libxb200.
#include "common.cuh"

extern "C" int xb_version(void) { return XB_VERSION; }

extern "C" const char* xb_error_string(int code) {
    if (code == 0) return "ok";
    if (code == XB_E_BADARG) return "xb200: bad argument (null pointer, non-positive size or inconsistent options)";
    if (code == XB_E_UNSUPPORTED) return "xb200: unsupported configuration for this entry point";
    if (code == XB_E_DRIVER) return "xb200: CUDA driver entry point unavailable or tensor-map encode failed";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "xb200: unknown error code";
}
