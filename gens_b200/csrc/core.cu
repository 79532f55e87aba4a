// ABI bookkeeping for libgens_b200.so.
#include "common.cuh"

extern "C" int gens_abi_version(void) { return GENS_ABI_VERSION; }

extern "C" const char* gens_error_string(int code) {
    if (code == 0) return "ok";
    if (code == GENS_E_BADARG) return "gens_b200: bad argument (null pointer or non-positive size)";
    if (code == GENS_E_UNSUPPORTED) return "gens_b200: shape not supported by the sm_100a kernels";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "gens_b200: unknown error";
}
