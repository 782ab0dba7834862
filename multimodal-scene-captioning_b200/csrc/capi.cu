// capi.cu -- error plumbing and device queries of the C-ABI (include/msc_geom.h).
#include <stdarg.h>

#include "msc_common.cuh"

namespace msc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace msc

extern "C" {

int msc_abi_version(void) { return MSC_ABI_VERSION; }

const char* msc_last_error(void) { return msc::g_err; }

int msc_device_info(int32_t* sm_count, int32_t* smem_optin_bytes, int32_t* cc_major, int32_t* cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        msc::set_error("no CUDA device visible");
        return MSC_ERR_NO_DEVICE;
    }
    int dev = 0, v = 0;
    MSC_CUDA(cudaGetDevice(&dev));
    if (sm_count) { MSC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (smem_optin_bytes) { MSC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)); *smem_optin_bytes = v; }
    if (cc_major) { MSC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { MSC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return MSC_OK;
}

}  // extern "C"
