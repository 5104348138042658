// C-ABI plumbing shared by all translation units: error string, launch check, device info.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void scann_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int scann_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        scann_set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

extern "C" const char* scann_last_error(void) { return g_err; }

// Programmatic dependent launch for the kernels that support it (common.cuh); per calling thread.
// The caller switches it off around launches whose stream predecessor is not one of this library's
// PDL-aware kernels (memsets, event joins from another stream).
static thread_local bool g_pdl = false;
bool scann_pdl_enabled() { return g_pdl; }
extern "C" int scann_set_pdl(int on) {
    int prev = g_pdl ? 1 : 0;
    g_pdl = on != 0;
    return prev;
}

// Local-attention kernels with four warp groups per CTA (la_tc.cu, "tc4"): bit mask of the kernels that use them
// when the pair plan's tiles hold at most 48 rows; per calling thread, returns the previous mask.
static thread_local int g_la4 = 0;
int scann_la_tc4_mask() { return g_la4; }
bool scann_plan_unfused() { return (g_la4 & 32) != 0; }
extern "C" int scann_set_la_groups4(int mask) {
    int prev = g_la4;
    g_la4 = mask & 63;      // bit 4: stagger the start of the groups; bit 5 (development): unfused pair plan
    return prev;
}

extern "C" int scann_version(void) { return 100; }   // 0.1.0

// Number of SMs of the current device, or -1 (with the error string set) when no CUDA device
// is usable.  The product path calls this first and refuses to run without a GPU.
extern "C" int scann_device_sm_count(void) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        scann_set_error("no usable CUDA device: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return -1;
    }
    return sms;
}

extern "C" int scann_device_cc(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}
