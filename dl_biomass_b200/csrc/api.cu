// libb2pn: ABI bookkeeping entry points (see include/b2pn.h).
#include <math.h>

#include "common.cuh"

extern "C" int b2pn_abi_version(void) { return B2PN_ABI_VERSION; }

extern "C" const char *b2pn_error_string(int code)
{
    if (code == B2PN_OK) return "ok";
    if (code == B2PN_EINVAL) return "b2pn: invalid argument (null pointer, negative size or bad flag)";
    if (code == B2PN_ENOTSUP) return "b2pn: shape not supported by the sm_100a kernels";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "b2pn: unknown error";
}

extern "C" int64_t b2pn_fps_num_samples(int64_t n, float ratio)
{
    // torch_cluster: ceil(float32(n) * float32(ratio)); keep the product in fp32
    volatile float prod = (float)n * ratio;
    return (int64_t)ceilf(prod);
}
