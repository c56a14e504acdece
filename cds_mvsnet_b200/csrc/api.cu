// Error reporting and library identification for the cds_b200 C ABI.
#include <cstdarg>
#include <cstdio>

#include "cds_common.cuh"

namespace {
thread_local char g_error[512] = "";
}

void cds_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cds_check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the sticky launch error so the next call starts clean
        cds_set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return CDS_OK;
}

extern "C" {

int cds_version(void) { return 100; }  // 0.1.0

const char* cds_last_error_string(void) { return g_error; }

// Device the library was built for; the host side refuses anything else loudly.
int cds_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cds_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { cds_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
    if (prop.major != 10) {
        cds_set_error("cds_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", dev, prop.major, prop.minor, prop.name);
        return CDS_EUNSUPPORTED;
    }
    return CDS_OK;
}

}  // extern "C"
