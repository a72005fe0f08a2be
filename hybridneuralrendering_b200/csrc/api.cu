// libhnr: error reporting + ABI version (see include/hnr.h).
#include "common.cuh"

static thread_local char g_err[512] = "";

extern "C" void hnr_set_error(const char* msg) {
    strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}
extern "C" const char* hnr_last_error(void) { return g_err; }
extern "C" int hnr_abi_version(void) { return 1; }

// Device sanity check used by the loader: returns the compute capability major*10+minor of the
// current device, or a negative code.  The kernels in this library are sm_100a-only.
extern "C" int hnr_device_arch(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        hnr_set_error("no CUDA device");
        return HNR_ERR_CUDA;
    }
    return p.major * 10 + p.minor;
}
