// capi.cu -- library-wide pieces of the C ABI (include/minbpe_b200.h)
#include "common.cuh"

namespace mbpe {
std::string &last_error_ref() {
    static thread_local std::string s;
    return s;
}
} // namespace mbpe

extern "C" const char *mbpe_version(void) { return "minbpe-cc_b200 0.1 (sm_100a)"; }
extern "C" const char *mbpe_last_error(void) { return mbpe::last_error_ref().c_str(); }
extern "C" int mbpe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError(); // clear the sticky "no device" so later calls report their own error
        return 0;
    }
    return n;
}
