// common.cu -- error plumbing, device checks, version.
#include "common.cuh"
#include <mutex>

namespace sedb200 {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

namespace {
struct DevInfo { int checked = 0, major = 0, minor = 0, sms = 0; };
DevInfo g_dev[64];
std::mutex g_mu;

int probe(int dev, DevInfo** out) {
    if (dev < 0 || dev >= 64) return fail(SEDB200_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevInfo& d = g_dev[dev];
    if (!d.checked) {
        SED_CUDA_OK(cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev));
        SED_CUDA_OK(cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev));
        SED_CUDA_OK(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        d.checked = 1;
    }
    *out = &d;
    return SEDB200_OK;
}
}  // namespace

int require_sm100() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail(SEDB200_ECUDA, "cudaGetDevice -> %s (no CUDA device: this library has no CPU path)",
                    cudaGetErrorString(e));
    DevInfo* d = nullptr;
    int rc = probe(dev, &d);
    if (rc) return rc;
    if (d->major != 10)
        return fail(SEDB200_EARCH, "device %d is sm_%d%d; libsedb200 is built for sm_100a only", dev,
                    d->major, d->minor);
    return SEDB200_OK;
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    DevInfo* d = nullptr;
    if (probe(dev, &d)) return 148;
    return d->sms > 0 ? d->sms : 148;
}

}  // namespace sedb200

extern "C" {

int sedb200_version(void) { return SEDB200_VERSION; }

const char* sedb200_last_error(void) { return sedb200::err_buf(); }

int sedb200_device_check(int device) {
    if (device >= 0) {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess)
            return sedb200::fail(SEDB200_ECUDA, "cudaSetDevice(%d) -> %s", device, cudaGetErrorString(e));
    }
    return sedb200::require_sm100();
}

}  // extern "C"
