// api.cu -- library info entry points of libvoge_b200.so.
#include "../../include/voge_b200.h"
#include "common.cuh"

extern "C" int voge_version(void) { return 1; }

extern "C" const char* voge_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

extern "C" int voge_device_sm_count(int* sm_count) {
    int dev = 0;
    VOGE_CUDA_TRY(cudaGetDevice(&dev));
    VOGE_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}
