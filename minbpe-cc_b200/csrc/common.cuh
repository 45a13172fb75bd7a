// common.cuh -- error plumbing shared by the CUDA translation units of libminbpe_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/minbpe_b200.h"

namespace mbpe {

std::string &last_error_ref(); // thread-local, defined in capi.cu

inline int set_error(int code, const std::string &msg) {
    last_error_ref() = msg;
    return code;
}

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    return set_error(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MBPE_E_NO_DEVICE : MBPE_E_CUDA, buf);
}

#define MB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return mbpe::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// select a device or fail loudly: the library has no CPU fallback
inline int use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(MBPE_E_NO_DEVICE, "no CUDA device: libminbpe_b200 has no CPU fallback");
    if (device < 0 || device >= n) return set_error(MBPE_E_INVALID, "device index out of range");
    MB_CUDA(cudaSetDevice(device));
    return MBPE_OK;
}

inline int sm_count(int device) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    return v;
}

} // namespace mbpe
